// kernels/superk.cuh -- EXPERIMENTAL stage-1 front end on super-k-mers (SIMKA_B200_SUPERK=1; not the default path).
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
//
// The reference never partitions k-mers one by one: Sequence2SuperKmer (GC/kmer/impl/Sequence2SuperKmer.hpp:80-159) cuts a
// read into maximal runs of consecutive k-mers that share their minimizer and FillPartitions::processSuperkmer
// (GC/kmer/impl/SortingCountAlgorithm.cpp:1048-1140) files each run, 2 bits per base, under the partition of that minimizer.
// Same idea here, sized for the GPU: what moves through memory between the scan of the reads and the counting kernel is one
// 16-byte record per run (about 2 B per k-mer instead of 8 B twice), and the per-k-mer rank atomics of the two partition
// levels become one atomic per run.
//
//   k_superk_scan<K>   thread = one packed word = 32 start positions.  Canonical m-mers (m = 13) rolled through registers and
//                      hashed (a bijection on 32 bits, so ties are identical m-mers); minimizer VALUE of every k-mer = sliding
//                      minimum over its K-m+1 m-mers by block prefix / suffix minima (van Herk), all indices static.  A k-mer's
//                      MICRO-BIN (up to 2^21 of them) is a function of that value alone, hence of the canonical k-mer (both
//                      strands see the same set of canonical m-mers): every occurrence of a k-mer lands in the same micro-bin,
//                      which is all exact counting needs.  Runs = consecutive valid k-mers of the word with equal micro-bin (cut
//                      at the word boundary and at the record capacity); each run is appended to an unsorted record list (one
//                      atomic per warp) and adds its k-mer count to the micro-bin histogram (one atomic per run).
//   k_superk_bins      minimizer frequencies follow the coverage of the genomes they come from (bins of a fixed hash range came
//                      out at 916 +- 703 keys on the bench workload, tools/superk_balance.py), so the counting bins are cut
//                      per sample from the scanned histogram, as the reference balances its partitions from sampled minimizer
//                      frequencies (Repartitor::computeDistrib, GC/kmer/impl/PartiInfo.cpp:48-105): a bin = the consecutive
//                      micro-bins whose first key falls into the same multiple of SK_TARGET keys.
//   k_superk_expand    thread = one record: reserves its k-mers' slots in the micro-bin (one atomic), rolls the canonical
//                      k-mers out of the record, mixes them and writes the keys -- the layout k_bin_count reads today.
//
// Record (128 bits): bases of the run, first base most significant, in the top 102 bits (<= 51 bases = k-mers + K - 1),
// then 5 bits (k-mers - 1), then 21 bits of micro-bin id.
#pragma once

constexpr int SK_M = 13;          // minimizer length
constexpr int SK_MAXB = 51;       // bases a record holds
constexpr int SK_MINK = 17;       // smallest k of this path (window of >= 5 m-mers)
constexpr int SK_THREADS = 256;
constexpr int SK_MAXMICRO = 21;   // micro-bin id bits a record can carry
constexpr int SK_TARGET = 896;    // keys per counting bin (+ the micro-bin that crosses the mark)

struct __align__(16) SkRecord { uint64_t hi, lo; };

// bijections on 32 bits: the order of the m-mers (smallest hash = minimizer) and the bin of a minimizer
__device__ __forceinline__ uint32_t sk_hash(uint32_t x) {
    x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA77u;  // (the order only needs the high bits mixed; sk_bin multiplies once more)
    return x;
}
__device__ __forceinline__ uint32_t sk_bin(uint32_t minv, int bin_bits) {
    // min over a window concentrates near 0: multiply by an odd constant so that the TOP bits depend on all of it
    return bin_bits ? (minv * 0xC2B2AE3Du) >> (32 - bin_bits) : 0u;
}
__device__ __forceinline__ uint32_t sk_revcomp32(uint32_t x, int m) {  // m-mer in the low 2m bits, m <= 16
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = __byte_perm(x, 0, 0x0123);
    x ^= 0xAAAAAAAAu;
    return x >> (32 - 2 * m);
}

template <int K>
__global__ void __launch_bounds__(SK_THREADS, 2)
k_superk_scan(const uint64_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t n_words, int bin_bits,
              SkRecord *__restrict__ records, ull *__restrict__ n_records, uint64_t cap, uint32_t *__restrict__ bin_hist,
              int *__restrict__ overflow_flag, ull *__restrict__ hist64 /* != null: (records << 32 | k-mers) per micro-bin instead of bin_hist */) {
    constexpr int M = SK_M;
    constexpr int W = K - M + 1;                 // m-mers per k-mer
    constexpr int NQ = 32 + W - 1;               // m-mers the 32 k-mers of a word touch (bases 0 .. 30 + K <= 61)
    constexpr int NMAX = (SK_MAXB - K + 1) < 32 ? (SK_MAXB - K + 1) : 32;  // k-mers per record
    static_assert(K >= SK_MINK && K <= 31, "k_superk_scan: 17 <= K <= 31");
    __shared__ uint32_t s_bin[32 * SK_THREADS];  // minimizer value of start position p of thread t at [p * SK_THREADS + t]

    const int tid = threadIdx.x, lane = tid & 31;
    const uint64_t t = (uint64_t)blockIdx.x * SK_THREADS + tid;
    uint64_t hi = 0, lo = 0, bm = ~0ull;
    if (t < n_words) {
        hi = seq[t];
        lo = seq[t + 1];
        bm = ((uint64_t)bad[t] << 32) | bad[t + 1];
    }
    // valid k-mer start positions (bit p), as in k_extract
    uint32_t validmask;
    {
        uint64_t r = bm;
        int built = 1;
        while (built * 2 <= K) { r |= r << built; built *= 2; }
        if (built < K) r |= r << (K - built);
        validmask = __brev(~(uint32_t)(r >> 32));
    }
    uint32_t startmask = 0;
    if (validmask) {  // (warp-uniform enough: reads are long compared with a word)
        // hashed canonical m-mers, rolled
        uint32_t hv[NQ];
        {
            constexpr uint32_t MMASK = (1u << (2 * M)) - 1;
            uint32_t fwd = (uint32_t)(hi >> (64 - 2 * M));
            uint32_t rev = sk_revcomp32(fwd, M);
            hv[0] = sk_hash(fwd < rev ? fwd : rev);
#pragma unroll
            for (int q = 1; q < NQ; q++) {
                const int i = q + M - 1;  // index of the incoming base
                const uint32_t b = i < 32 ? (uint32_t)(hi >> (62 - 2 * i)) & 3u : (uint32_t)(lo >> (62 - 2 * (i - 32))) & 3u;
                fwd = ((fwd << 2) | b) & MMASK;
                rev = (rev >> 2) | ((b ^ 2u) << (2 * M - 2));
                hv[q] = sk_hash(fwd < rev ? fwd : rev);
            }
        }
        // sliding minimum over W m-mers: blocks of W, S[q] = min to the end of q's block, then hv[q] := min from the start of
        // q's block; window [p, p + W - 1] = min(S[p], P[p + W - 1])
        uint32_t S[32];
#pragma unroll
        for (int q = 31; q >= 0; q--) {
            // the block of q ends at e = (q / W) * W + W - 1 (clipped to NQ - 1); S[q] needs S[q + 1] unless q is a block end
            const int e = (q / W) * W + W - 1 < NQ - 1 ? (q / W) * W + W - 1 : NQ - 1;
            if (q == 31) {
                uint32_t v = hv[e];
#pragma unroll
                for (int j = e - 1; j >= 31; j--) v = hv[j] < v ? hv[j] : v;
                S[31] = v;
            } else if (q == e) {
                S[q] = hv[q];
            } else {
                S[q] = hv[q] < S[q + 1] ? hv[q] : S[q + 1];
            }
        }
#pragma unroll
        for (int q = 1; q < NQ; q++)
            if (q % W != 0) hv[q] = hv[q - 1] < hv[q] ? hv[q - 1] : hv[q];
        // runs: consecutive valid k-mers with equal bin, at most NMAX of them
        // (compared on the minimizer value: equal values have equal bins; the bin itself is computed once per run, at emission)
        uint32_t prevmv = 0;
        int len = 0;
#pragma unroll
        for (int p = 0; p < 32; p++) {
            const uint32_t mv = S[p] < hv[p + W - 1] ? S[p] : hv[p + W - 1];
            s_bin[p * SK_THREADS + tid] = mv;
            const bool v = (validmask >> p) & 1u;
            if (v && len > 0 && mv == prevmv && len < NMAX) len++;
            else if (v) { startmask |= 1u << p; len = 1; prevmv = mv; }
            else len = 0;
        }
    }
    // emit the runs: one list reservation per warp (prefix sum of the lanes' run counts, one atomic), then every lane writes its
    // runs to consecutive slots
    const uint32_t mine = (uint32_t)__popc(startmask);
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    const uint32_t warp_total = __shfl_sync(0xffffffffu, inc, 31);
    if (warp_total == 0) return;  // (warp-uniform)
    ull base = 0;
    if (lane == 31) base = atomicAdd(n_records, (ull)warp_total);
    base = __shfl_sync(0xffffffffu, base, 31);
    ull slot = base + (inc - mine);
    for (uint32_t todo = startmask; todo; todo &= todo - 1, slot++) {
        const int p = __ffs((int)todo) - 1;
        const uint32_t stop = p < 31 ? (startmask | ~validmask) >> (p + 1) : 0u;  // bit i <-> position p + 1 + i starts a run or is invalid
        const uint32_t n = stop ? (uint32_t)__ffs((int)stop) : (uint32_t)(32 - p);
        const uint32_t b = sk_bin(s_bin[p * SK_THREADS + tid], bin_bits);
        if (slot < cap) {
            SkRecord rec;
            rec.hi = p ? (hi << (2 * p)) | (lo >> (64 - 2 * p)) : hi;
            rec.lo = ((p ? lo << (2 * p) : lo) & ~((1ull << 26) - 1)) | ((uint64_t)(n - 1) << 21) | b;
            records[slot] = rec;
            if (hist64) atomicAdd(&hist64[b], (1ull << 32) | n);
            else atomicAdd(&bin_hist[b], n);
        } else {
            *overflow_flag = 1;  // the host redoes the sample on the k-mer path
        }
    }
}

typedef void (*superk_scan_fn)(const uint64_t *, const uint32_t *, uint64_t, int, SkRecord *, ull *, uint64_t, uint32_t *, int *, ull *);
static superk_scan_fn superk_scan_instance(int k) {
    static const superk_scan_fn tab[15] = {k_superk_scan<17>, k_superk_scan<18>, k_superk_scan<19>, k_superk_scan<20>, k_superk_scan<21>,
                                           k_superk_scan<22>, k_superk_scan<23>, k_superk_scan<24>, k_superk_scan<25>, k_superk_scan<26>,
                                           k_superk_scan<27>, k_superk_scan<28>, k_superk_scan<29>, k_superk_scan<30>, k_superk_scan<31>};
    return (k >= SK_MINK && k <= 31) ? tab[k - SK_MINK] : nullptr;
}

// fine_off[b] = first key of counting bin b = prefix[u] of the first micro-bin u with prefix[u] >= b * target (b <= nb; bins
// past the last key are empty).  One thread per bin: a binary search over the exclusive prefix (an earlier version let the thread
// of the last micro-bin fill every trailing empty bin by itself: 150 us per sample).
__global__ void __launch_bounds__(BLOCK)
k_superk_bins(const uint64_t *__restrict__ prefix /*[n_micro + 1]*/, uint32_t n_micro, uint64_t *__restrict__ fine_off, uint32_t nb, uint32_t target,
              const uint64_t *__restrict__ rec_prefix /*[n_micro + 1] or null*/, uint64_t *__restrict__ rec_off /*[nb + 1]: first record of bin b*/) {
    const uint32_t b = blockIdx.x * BLOCK + threadIdx.x;
    if (b > nb) return;
    const uint64_t mark = (uint64_t)b * target;
    uint32_t lo = 0, hi = n_micro;  // smallest u in [0, n_micro] with prefix[u] >= mark (prefix[0] = 0; u = n_micro when there is none below)
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (prefix[mid] >= mark) hi = mid; else lo = mid + 1;
    }
    fine_off[b] = prefix[lo];
    if (rec_prefix) rec_off[b] = rec_prefix[lo];
}

// records -> mixed canonical k-mers, written where k_bin_count expects the keys of the record's micro-bin
__global__ void __launch_bounds__(BLOCK)
k_superk_expand(const SkRecord *__restrict__ records, const ull *__restrict__ n_records, uint64_t cap, KParams kp,
                const uint64_t *__restrict__ micro_off, uint32_t *__restrict__ micro_cursor, uint64_t *__restrict__ keys_out) {
    const ull nr = *n_records < cap ? *n_records : cap;
    const int k = kp.k;
    const uint64_t kmask = kp.mix.mask;
    for (ull r = (ull)blockIdx.x * BLOCK + threadIdx.x; r < nr; r += (ull)gridDim.x * BLOCK) {
        const SkRecord rec = records[r];
        const uint32_t bin = (uint32_t)rec.lo & 0x1FFFFFu;
        const uint32_t n = ((uint32_t)(rec.lo >> 21) & 31u) + 1;
        uint64_t pos = micro_off[bin] + atomicAdd(&micro_cursor[bin], n);
        uint64_t fwd = rec.hi >> (64 - 2 * k);
        uint64_t rev = simka::revcomp(fwd, k);
        keys_out[pos] = simka::mix_fwd(fwd < rev ? fwd : rev, kp.mix);
        for (uint32_t j = 1; j < n; j++) {
            const int i = k - 1 + (int)j;  // index of the incoming base
            const uint64_t b = i < 32 ? (rec.hi >> (62 - 2 * i)) & 3u : (rec.lo >> (62 - 2 * (i - 32))) & 3u;
            fwd = ((fwd << 2) | b) & kmask;
            rev = (rev >> 2) | ((b ^ 2u) << (2 * k - 2));
            keys_out[pos + j] = simka::mix_fwd(fwd < rev ? fwd : rev, kp.mix);
        }
    }
}

// Same job with the work spread evenly: a CTA takes BLOCK records, reserves their slots (one atomic per record), cuts every record
// into items of SK_G consecutive k-mers, lays the items of the chunk out flat (block scan of the item counts + an owner byte per
// item) and every thread then takes single items: the first k-mer of an item straight from the record (shift, reverse complement,
// mix: 69 instructions), the following ones rolled (26).  Records hold 1 to 21 k-mers (7.4 on average), so one thread per record
// leaves two thirds of the lanes idle; items of 4 keep the lanes busy at about 40 instructions per k-mer.
constexpr int SK_G = 4;
__global__ void __launch_bounds__(BLOCK)
k_superk_expand_flat(const SkRecord *__restrict__ records, const ull *__restrict__ n_records, uint64_t cap, KParams kp,
                     const uint64_t *__restrict__ micro_off, uint32_t *__restrict__ micro_cursor, uint64_t *__restrict__ keys_out) {
    __shared__ SkRecord s_rec[BLOCK];
    __shared__ uint64_t s_pos[BLOCK];
    __shared__ uint32_t s_start[BLOCK], warp_tmp[BLOCK / 32 + 1];
    __shared__ uint8_t s_owner[BLOCK * (32 / SK_G)];
    const int tid = threadIdx.x;
    const ull nr = *n_records < cap ? *n_records : cap;
    const int k = kp.k;
    const uint64_t kmask = kp.mix.mask;
    for (ull c0 = (ull)blockIdx.x * BLOCK; c0 < nr; c0 += (ull)gridDim.x * BLOCK) {
        const ull r = c0 + tid;
        uint32_t items = 0;
        if (r < nr) {
            const SkRecord rec = records[r];
            const uint32_t micro = (uint32_t)rec.lo & 0x1FFFFFu;
            const uint32_t n = ((uint32_t)(rec.lo >> 21) & 31u) + 1;
            items = (n + SK_G - 1) / SK_G;
            s_rec[tid] = rec;
            s_pos[tid] = micro_off[micro] + atomicAdd(&micro_cursor[micro], n);
        }
        s_start[tid] = items;
        __syncthreads();
        block_exclusive_scan(s_start, BLOCK, warp_tmp);
        const uint32_t total = warp_tmp[BLOCK / 32];
        {
            const uint32_t st = s_start[tid];
            for (uint32_t j = 0; j < items; j++) s_owner[st + j] = (uint8_t)tid;
        }
        __syncthreads();
        for (uint32_t i = tid; i < total; i += BLOCK) {
            const uint32_t o = s_owner[i];
            const uint32_t j0 = (i - s_start[o]) * SK_G;
            const uint64_t hi = s_rec[o].hi, lo = s_rec[o].lo;
            const uint32_t n = ((uint32_t)(lo >> 21) & 31u) + 1;
            const uint64_t x = j0 ? (hi << (2 * j0)) | (lo >> (64 - 2 * j0)) : hi;  // the record from base j0 on (j0 + k <= 51 bases: never the tag bits)
            uint64_t fwd = x >> (64 - 2 * k);
            uint64_t rev = simka::revcomp(fwd, k);
            uint64_t *out = keys_out + s_pos[o] + j0;
            out[0] = simka::mix_fwd(fwd < rev ? fwd : rev, kp.mix);
#pragma unroll
            for (int g = 1; g < SK_G; g++) {
                if (j0 + g < n) {
                    const int b_idx = k - 1 + (int)j0 + g;  // index of the incoming base
                    const uint64_t b = b_idx < 32 ? (hi >> (62 - 2 * b_idx)) & 3u : (lo >> (62 - 2 * (b_idx - 32))) & 3u;
                    fwd = ((fwd << 2) | b) & kmask;
                    rev = (rev >> 2) | ((b ^ 2u) << (2 * k - 2));
                    out[g] = simka::mix_fwd(fwd < rev ? fwd : rev, kp.mix);
                }
            }
        }
        __syncthreads();
    }
}
