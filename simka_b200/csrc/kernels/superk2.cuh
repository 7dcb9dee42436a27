// kernels/superk2.cuh -- stage 1 on super-k-mer RECORDS end to end (the default front end for 17 <= k <= 31).
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
//
// kernels/superk.cuh turns the reads into 16-byte super-k-mer records (runs of consecutive k-mers that share a minimizer bin;
// what the reference's Sequence2SuperKmer / FillPartitions::processSuperkmer write to its partition files,
// GC/kmer/impl/Sequence2SuperKmer.hpp:80-159, GC/kmer/impl/SortingCountAlgorithm.cpp:1048-1140).  Here the records themselves
// are what gets partitioned and what the counting kernel reads -- as the reference's PartitionsByVectorCommand reads the
// super-k-mer files and expands them in memory (GC/kmer/impl/PartitionsCommand.cpp:924-1150,1253-1378):
//
//   k_superk_scan      (superk.cuh) reads -> unsorted record list + per-micro-bin (k-mers | records) histogram, one 64-bit atomic per run
//   k_sk2_split        histogram -> the two u32 arrays the device-wide scans take
//   k_superk_bins      (superk.cuh) counting bins = consecutive micro-bins holding about SK2_TARGET k-mers (+ their record offsets)
//   k_sk2_scatter      records -> sorted by micro-bin (16 B per ~7 k-mers: 2.2 B per k-mer instead of 8 B keys through two levels)
//   k_sk2_count        persistent CTAs over counting bins: the bin's records arrive in shared memory by 1-D BULK COPIES
//                      (cp.async.bulk + mbarrier, double-buffered: the copy engine fetches chunk i+1 while the warps expand and
//                      insert chunk i); k-mers are rolled out of the records (items of SK2_G consecutive k-mers, laid out flat
//                      over the CTA), made canonical and counted in a shared-memory open-addressing table keyed by the canonical
//                      k-mer; read-out = abundance filter + per-sample totals (src/minikc/MiniKC.hpp:54-79); only the
//                      SURVIVORS are mixed into the hash order the merge works in.
//
// Keys cross HBM zero times: per input k-mer the path moves 0.375 B of packed reads in, ~2.2 B of records out, in, out and in
// again, and 12 B per SOLID distinct k-mer out.
#pragma once

#ifndef SIMKA_SK2_SLOTS
#define SIMKA_SK2_SLOTS 4096
#endif
#ifndef SIMKA_SK2_TARGET
#define SIMKA_SK2_TARGET 2048
#endif
#ifndef SIMKA_SK2_THREADS
#define SIMKA_SK2_THREADS 256
#endif
constexpr int SK2_THREADS = SIMKA_SK2_THREADS;
constexpr int SK2_SLOTS = SIMKA_SK2_SLOTS;          // table slots of a counting CTA (a power of two)
constexpr int sk2_log2(int v) { return v <= 1 ? 0 : 1 + sk2_log2(v >> 1); }
constexpr int SK2_SLOTS_LOG2 = sk2_log2(SK2_SLOTS);
constexpr int SK2_MAXD = SK2_SLOTS * 3 / 4;         // distinct k-mers one table pass may hold
constexpr int SK2_TARGET = SIMKA_SK2_TARGET;        // k-mers per counting bin (+ the micro-bin that crosses the mark)
#ifndef SIMKA_SK2_RCAP
#define SIMKA_SK2_RCAP 256
#endif
constexpr int SK2_RCAP = SIMKA_SK2_RCAP;            // records per shared-memory chunk (16 B each), a multiple of SK2_THREADS
static_assert(SK2_RCAP % SK2_THREADS == 0 && SK2_RCAP * 8 <= 8192, "record chunk: a multiple of the CTA size, at most 1024 records (13-bit owner ids)");
constexpr int SK2_G = 4;                            // consecutive k-mers per work item
constexpr int SK2_MAXITEMS = SK2_RCAP * (32 / SK2_G);

__global__ void __launch_bounds__(BLOCK)
k_sk2_split(const ull *__restrict__ hist64, uint32_t n, uint32_t *__restrict__ kmers, uint32_t *__restrict__ recs) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i < n) { const ull v = hist64[i]; kmers[i] = (uint32_t)v; recs[i] = (uint32_t)(v >> 32); }
}

// records -> grouped by micro-bin (order inside a micro-bin is irrelevant: counting needs grouping only)
__global__ void __launch_bounds__(BLOCK)
k_sk2_scatter(const SkRecord *__restrict__ records, const ull *__restrict__ n_records, uint64_t cap, const uint64_t *__restrict__ rec_prefix,
              uint32_t *__restrict__ cursor, SkRecord *__restrict__ out) {
    const ull nr = *n_records < cap ? *n_records : cap;
    for (ull r = (ull)blockIdx.x * BLOCK + threadIdx.x; r < nr; r += (ull)gridDim.x * BLOCK) {
        const SkRecord rec = records[r];
        const uint32_t u = (uint32_t)rec.lo & 0x1FFFFFu;
        out[rec_prefix[u] + atomicAdd(&cursor[u], 1u)] = rec;
    }
}

struct Sk2Item { uint64_t r0; uint32_t bin, nrec, last, pad; };  // a work item: records [r0, r0 + nrec) of counting bin `bin`; last chunk of the bin?
struct Sk2Smem {
    alignas(16) uint64_t tk[SK2_SLOTS];
    alignas(16) SkRecord rec[2][SK2_RCAP];
    alignas(16) uint32_t tc[SK2_SLOTS];
    alignas(16) Sk2Item item[2];
    alignas(8) uint64_t bar[2];
    uint32_t warp_tmp[SK2_THREADS / 32 + 1];
    uint32_t cursor, nclaim;
    ull tot[3];
    uint16_t claimed[SK2_SLOTS];
    uint16_t owner[SK2_MAXITEMS];   // (record << 3 | item of the record) of every work item of the chunk
};

__device__ __forceinline__ uint32_t sk2_slot(uint64_t key) {
    // 32-bit mix of both halves (the k-mers of a bin share their minimizer, i.e. a substring: multiply twice)
    uint32_t x = (uint32_t)key ^ ((uint32_t)(key >> 32) * 0x9E3779B1u);
    x *= 0x85EBCA77u;
    return x >> (32 - SK2_SLOTS_LOG2);
}

// ONE probe of the canonical-k-mer table, branches closed before returning so that the warp reconverges after every probe.
// Plain read first: most k-mers repeat one that is already there.  Returns true when the key is placed.
__device__ __forceinline__ bool sk2_probe(Sk2Smem &S, uint64_t key, uint32_t slot) {
    volatile uint64_t *vk = S.tk;
    uint64_t cur = vk[slot];
    if (cur == EMPTY) {
        cur = atomicCAS(reinterpret_cast<ull *>(&S.tk[slot]), (ull)EMPTY, (ull)key);
        if (cur == EMPTY) {
            cur = key;
            const uint32_t q = atomicAdd(&S.nclaim, 1u);
            if (q < (uint32_t)SK2_SLOTS) S.claimed[q] = (uint16_t)slot;
        }
    }
    const bool hit = cur == key;
    if (hit) atomicAdd(&S.tc[slot], 1u);
    return hit;
}
// linear probing from the slot after `slot` (the leftovers of a converged first round, or the slow path); a full table (more
// than SK2_MAXD claims) is detected by the caller
__device__ __forceinline__ void sk2_insert_from(Sk2Smem &S, uint64_t key, uint32_t slot) {
    for (int probes = 1; probes < SK2_SLOTS; probes++) {
        slot = (slot + 1) & (SK2_SLOTS - 1);
        if (sk2_probe(S, key, slot)) return;
    }
}
__device__ __forceinline__ void sk2_insert(Sk2Smem &S, uint64_t key) {
    const uint32_t slot = sk2_slot(key);
    if (!sk2_probe(S, key, slot)) sk2_insert_from(S, key, slot);
}

// the canonical k-mers j0 .. j0+cnt-1 of a record, each handed to f
template <class F>
__device__ __forceinline__ void sk2_expand(uint64_t hi, uint64_t lo, uint32_t j0, uint32_t cnt, int k, uint64_t kmask, F f) {
    const uint64_t x = j0 ? (hi << (2 * j0)) | (lo >> (64 - 2 * j0)) : hi;  // the record from base j0 on (j0 + k <= 51 bases: never the tag bits)
    uint64_t fwd = x >> (64 - 2 * k);
    uint64_t rev = simka::revcomp(fwd, k);
    f(fwd < rev ? fwd : rev);
    for (uint32_t g = 1; g < cnt; g++) {
        const int b_idx = k - 1 + (int)j0 + (int)g;  // index of the incoming base
        const uint64_t b = b_idx < 32 ? (hi >> (62 - 2 * b_idx)) & 3u : (lo >> (62 - 2 * (b_idx - 32))) & 3u;
        fwd = ((fwd << 2) | b) & kmask;
        rev = (rev >> 2) | ((b ^ 2u) << (2 * k - 2));
        f(fwd < rev ? fwd : rev);
    }
}

// Same for one work item (j0 a multiple of 4, cnt <= 4) on 32-bit halves (17 <= k <= 31), as k_extract16 rolls its windows: a
// 64-bit shift by a run-time amount costs 4-6 instructions, a funnel shift one.  The item's k + 3 bases start on a byte
// boundary of the record (j0 = 4 i), so its 96-bit window (P, Q, R) is three funnel shifts by 0 / 8 / 16 / 24 bits.
template <class F>
__device__ __forceinline__ void sk2_expand_item(uint64_t hi, uint64_t lo, uint32_t j0, uint32_t cnt, int k, F f) {
    const uint32_t A = (uint32_t)(hi >> 32), B = (uint32_t)hi, Cw = (uint32_t)(lo >> 32), D = (uint32_t)lo;
    const bool up = j0 >= 16;  // window starts in the second word
    const uint32_t w0 = up ? B : A, w1 = up ? Cw : B, w2 = up ? D : Cw, w3 = up ? 0u : D;
    const uint32_t sh = (j0 & 15u) * 2;
    const uint32_t P = __funnelshift_l(w1, w0, sh), Q = __funnelshift_l(w2, w1, sh), R = __funnelshift_l(w3, w2, sh);
    const int s1 = 64 - 2 * k, s2 = 2 * k - 34;
    uint32_t rl = 0, rh = 0;
#pragma unroll
    for (int g = 0; g < SK2_G; g++) {
        if ((uint32_t)g < cnt) {
            uint32_t x1, x0;
            if (g == 0) { x1 = P; x0 = Q; }
            else { x1 = __funnelshift_l(Q, P, 2 * g); x0 = __funnelshift_l(R, Q, 2 * g); }
            const uint32_t fh = x1 >> s1, fl = __funnelshift_r(x0, x1, s1);
            if (g == 0) {
                const uint64_t r0 = simka::revcomp(((uint64_t)fh << 32) | fl, k);
                rh = (uint32_t)(r0 >> 32); rl = (uint32_t)r0;
            } else {
                rl = __funnelshift_r(rl, rh, 2);
                rh = (rh >> 2) | (((fl & 3u) ^ 2u) << s2);
            }
            const bool fwd_smaller = fh < rh || (fh == rh && fl < rl);
            f(g, ((uint64_t)(fwd_smaller ? fh : rh) << 32) | (fwd_smaller ? fl : rl));
        }
    }
}

struct Sk2Out {
    uint64_t *stage_h;   // survivors of bin b are written at key_off[b] + i (mixed keys, the order the merge works in)
    uint32_t *stage_c;
    uint32_t *surv;      // [nbins] survivor count per bin
    ull *totals;         // [3] distinct, sum c, sum c^2
    int *error_flag;
};

// SHARDED: bucket-sharded contexts (simka_params.shard_rank / shard_count: several GPUs, or several passes of one GPU when the
// solid lists of all samples do not fit in HBM) keep the k-mers whose hash falls into their shard -- the same rule as the k-mer
// path (shard_keep on the mixed key), so a sample that falls back to that path stays consistent with the others.
#ifndef SIMKA_SK2_MINB
#define SIMKA_SK2_MINB 3
#endif
template <bool SHARDED>
__global__ void __launch_bounds__(SK2_THREADS, SIMKA_SK2_MINB)
k_sk2_count(const SkRecord *__restrict__ recs, const uint64_t *__restrict__ rec_off, const uint64_t *__restrict__ key_off, uint32_t nbins,
            KParams kp, int64_t amin, int64_t amax, Sk2Out out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sk2Smem &S = *reinterpret_cast<Sk2Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int k = kp.k;
    const uint64_t kmask = kp.mix.mask;
    for (int i = tid; i < SK2_SLOTS; i += SK2_THREADS) { S.tk[i] = EMPTY; S.tc[i] = 0; }
    if (tid == 0) {
        S.cursor = 0; S.nclaim = 0; S.tot[0] = S.tot[1] = S.tot[2] = 0;
        mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    ull t_d = 0, t_s = 0, t_q = 0;
    const bool never = amax < 0 || amin > 0xFFFFFFFFll;  // keep iff amin <= c <= amax (src/minikc/MiniKC.hpp:56), c a non-zero u32
    const uint32_t cmin = amin < 0 ? 0u : (uint32_t)amin;
    const uint32_t cmax = amax > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)(amax < 0 ? 0 : amax);

    // read-out of the claimed slots (which also leaves the table clean): survivors staged at key_off[bin] + cursor
    auto read_out = [&](uint32_t bin, bool discard, ull &b_d, ull &b_s, ull &b_q) {
        const uint32_t nc = min(S.nclaim, (uint32_t)SK2_SLOTS);
        const uint64_t beg = key_off[bin];
        for (uint32_t e0 = 0; e0 < nc; e0 += SK2_THREADS) {
            const uint32_t e = e0 + tid;
            bool keep = false;
            uint64_t key = 0;
            uint32_t c = 0;
            if (e < nc) {
                const uint32_t slot = S.claimed[e];
                key = S.tk[slot];
                c = S.tc[slot];
                S.tk[slot] = EMPTY;
                S.tc[slot] = 0;
                keep = !discard && !never && c >= cmin && c <= cmax;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                uint32_t wbase = 0;
                if (lane == 0) wbase = atomicAdd(&S.cursor, (uint32_t)__popc(m));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (keep) {
                    const uint64_t pos = beg + wbase + __popc(m & ((1u << lane) - 1));
                    out.stage_h[pos] = simka::mix_fwd(key, kp.mix);
                    out.stage_c[pos] = c;
                    b_d += 1;
                    b_s += c;
                    b_q += (ull)((double)c * (double)c);  // pow(count,2) -> u64, src/minikc/MiniKC.hpp:62
                }
            }
        }
    };

    // Work items of this CTA: (bin, chunk of <= SK2_RCAP records), bins taken with a grid stride, empty bins skipped.  Thread 0
    // plans one item ahead: it writes the item's descriptor to shared memory, arms the buffer's mbarrier with the byte count
    // and starts the bulk copy of the records -- the copy engine fills the buffer while all warps work on the previous one.
    // plan(cursor state in registers of thread 0 only) -> S.item[buf]
    uint32_t p_bin = blockIdx.x, p_chunk = 0;  // thread 0: next item to plan
    auto plan = [&](int buf) {
        Sk2Item it;
        it.bin = 0xFFFFFFFFu; it.nrec = 0; it.last = 0; it.r0 = 0; it.pad = 0;
        while (p_bin < nbins) {
            const uint64_t b0 = rec_off[p_bin], b1 = rec_off[p_bin + 1];
            if (b1 == b0) { out.surv[p_bin] = 0; p_bin += gridDim.x; p_chunk = 0; continue; }
            const uint64_t r0 = b0 + (uint64_t)p_chunk * SK2_RCAP;
            it.bin = p_bin; it.r0 = r0;
            it.nrec = (uint32_t)min((uint64_t)SK2_RCAP, b1 - r0);
            it.last = r0 + it.nrec == b1;
            if (it.last) { p_bin += gridDim.x; p_chunk = 0; } else p_chunk++;
            break;
        }
        S.item[buf] = it;
        if (it.bin != 0xFFFFFFFFu) {
            mbar_arrive_expect_tx(&S.bar[buf], it.nrec * (uint32_t)sizeof(SkRecord));
            bulk_g2s(S.rec[buf], recs + it.r0, it.nrec * (uint32_t)sizeof(SkRecord), &S.bar[buf]);
        }
    };
    if (tid == 0) plan(0);
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;
    int buf = 0;
    for (;;) {
        const Sk2Item it = S.item[buf];
        if (it.bin == 0xFFFFFFFFu) break;
        if (tid == 0) plan(buf ^ 1);  // (the other buffer and its descriptor were released by the barrier that ended the previous iteration)
        if (buf == 0) { mbar_wait(&S.bar[0], phase0); phase0 ^= 1; } else { mbar_wait(&S.bar[1], phase1); phase1 ^= 1; }
        const SkRecord *sr = S.rec[buf];
        const uint32_t nrec = it.nrec, bin = it.bin;
        // lay the work items (SK2_G consecutive k-mers of one record) out flat: thread t owns records t, t + THREADS, ...;
        // exclusive prefix of its item count by warp shuffles + one pass over the warp totals
        uint32_t my_items[SK2_RCAP / SK2_THREADS], my_sum = 0;
#pragma unroll
        for (int q = 0; q < SK2_RCAP / SK2_THREADS; q++) {
            const uint32_t r = q * SK2_THREADS + tid;
            my_items[q] = r < nrec ? ((((uint32_t)(sr[r].lo >> 21) & 31u) + 1) + SK2_G - 1) / SK2_G : 0;
            my_sum += my_items[q];
        }
        uint32_t inc = my_sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
        if (lane == 31) S.warp_tmp[tid >> 5] = inc;
        __syncthreads();
        uint32_t base = inc - my_sum, total = 0;
#pragma unroll
        for (int w = 0; w < SK2_THREADS / 32; w++) { const uint32_t v = S.warp_tmp[w]; if (w < (tid >> 5)) base += v; total += v; }
#pragma unroll
        for (int q = 0; q < SK2_RCAP / SK2_THREADS; q++) {
            const uint32_t r = q * SK2_THREADS + tid;
            for (uint32_t j = 0; j < my_items[q]; j++) S.owner[base + j] = (uint16_t)((r << 3) | j);
            base += my_items[q];
        }
        __syncthreads();
        if (*((volatile uint32_t *)&S.nclaim) <= (uint32_t)SK2_MAXD) {  // (an overflowing bin is redone below: no point filling a full table)
            for (uint32_t i = tid; i < total; i += SK2_THREADS) {
                const uint32_t ow = S.owner[i];
                const uint32_t o = ow >> 3, j0 = (ow & 7u) * SK2_G;
                const uint64_t hi = sr[o].hi, lo = sr[o].lo;
                const uint32_t n = ((uint32_t)(lo >> 21) & 31u) + 1;
                // first probes of the item's k-mers with the warp converged; the few that find another k-mer in their home slot
                // are walked afterwards
                uint64_t left_key[SK2_G];
                uint32_t left = 0;
                sk2_expand_item(hi, lo, j0, min((uint32_t)SK2_G, n - j0), k, [&](int g, uint64_t canon) {
                    left_key[g] = canon;
                    if (!SHARDED || shard_keep(simka::mix_fwd(canon, kp.mix), kp)) {
                        if (!sk2_probe(S, canon, sk2_slot(canon))) left |= 1u << g;
                    }
                });
#pragma unroll
                for (int q = 0; q < SK2_G; q++)
                    if (left & (1u << q)) sk2_insert_from(S, left_key[q], sk2_slot(left_key[q]));
            }
        }
        __syncthreads();
        if (it.last) {
            ull b_d = 0, b_s = 0, b_q = 0;
            const bool overflow = S.nclaim > (uint32_t)SK2_MAXD;
            read_out(bin, overflow, b_d, b_s, b_q);
            __syncthreads();
            if (tid == 0) S.nclaim = 0;
            if (overflow) {
                // more distinct k-mers than one table pass holds (a heavy minimizer: low-complexity sequence): R hash-selected
                // sub-passes straight from global memory, one thread per record; rare, only has to be correct
                const uint64_t rb0 = rec_off[bin], nrec_bin = rec_off[bin + 1] - rb0;
                for (uint32_t R = 2;; R *= 2) {
                    if (R > 65536) { if (tid == 0) *out.error_flag = 5; break; }
                    b_d = b_s = b_q = 0;
                    bool again = false;
                    if (tid == 0) S.cursor = 0;
                    __syncthreads();
                    for (uint32_t r = 0; r < R && !again; r++) {
                        for (uint64_t q = tid; q < nrec_bin; q += SK2_THREADS) {
                            if (*((volatile uint32_t *)&S.nclaim) > (uint32_t)SK2_MAXD) break;
                            const SkRecord rc = recs[rb0 + q];
                            const uint32_t n = ((uint32_t)(rc.lo >> 21) & 31u) + 1;
                            sk2_expand(rc.hi, rc.lo, 0, n, k, kmask, [&](uint64_t canon) {
                                // (a full avalanche: the k-mers of a heavy minimizer differ in few bits, a multiply alone leaves
                                // thousands of them with equal selector bits and the sub-passes would never split them)
                                if (((uint32_t)(simka::murmur3_fmix64(canon) >> 20) & (R - 1)) == r && (!SHARDED || shard_keep(simka::mix_fwd(canon, kp.mix), kp)))
                                    sk2_insert(S, canon);
                            });
                        }
                        __syncthreads();
                        again = S.nclaim > (uint32_t)SK2_MAXD;
                        read_out(bin, again, b_d, b_s, b_q);
                        __syncthreads();
                        if (tid == 0) S.nclaim = 0;
                        __syncthreads();
                    }
                    if (!again) break;
                }
            }
            t_d += b_d; t_s += b_s; t_q += b_q;
            if (tid == 0) { out.surv[bin] = S.cursor; S.cursor = 0; }
        }
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        t_d += __shfl_down_sync(0xffffffffu, t_d, d);
        t_s += __shfl_down_sync(0xffffffffu, t_s, d);
        t_q += __shfl_down_sync(0xffffffffu, t_q, d);
    }
    if (lane == 0 && t_d) { atomicAdd(&S.tot[0], t_d); atomicAdd(&S.tot[1], t_s); atomicAdd(&S.tot[2], t_q); }
    __syncthreads();
    if (tid == 0 && S.tot[0]) { atomicAdd(&out.totals[0], S.tot[0]); atomicAdd(&out.totals[1], S.tot[1]); atomicAdd(&out.totals[2], S.tot[2]); }
}
