// kernels/merge_tiles.cuh -- stage 2 for 64 < N <= 512 samples: group once, accumulate per sample-pair tile in shared memory.
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
//
// With N > 64 the N(N-1)/2 pair accumulators no longer fit the shared memory of a CTA, and k_merge adds every pair term to
// global memory: 8 atomics per pair, 28-36 G pair-updates/s, 90 % of configuration 4's step.  Here the work is split the way
// the data wants it:
//   k_group_bins   (once per merge bin)  the grouping phases of k_merge; instead of enumerating pairs it writes, for every k-mer
//                  shared by >= 2 samples, a RECORD to a scratch buffer: the bitmap of the samples that hold it (one 32-bit
//                  word per block of 32 samples) and its abundances in sample order (the rank of a sample inside its group is
//                  the number of set bits below it, so no sort is needed).  The records of a batch are DENSE (a sub-pass claims
//                  the space of its records from a cursor; their order does not matter) and stored word-major.
//   k_merge_tiles  one CTA per (tile of 32 x 32 samples, slice of the batch's records): every thread looks at up to four records
//                  per chunk (two coalesced loads per record: the bitmap words of the tile's two sample blocks), the records
//                  with pairs in this tile become entries whose members are staged in shared memory, the tile's pairs are
//                  enumerated flat over the threads and added to tile-private accumulators in SHARED memory exactly as
//                  k_merge_sm does; the 32-bit words are swept only when 2047 entries have passed; one flush per CTA and batch.
// The scratch is the stage-1 work buffers (idle during the merge); when the records of all bins do not fit, the bins are
// processed in batches cut on the host from the per-bin entry counts.
// Reference: the same src/SimkaMerge.cpp:1164-1264,1307-1326 + src/core/SimkaAlgorithm.hpp:341-516 as k_merge.
#pragma once

constexpr int TM_MAXB = 16;        // sample blocks of 32: N <= 512
constexpr int TM_GCAP = 512;       // shared groups one sub-pass of k_group_bins may hold (their bitmaps live in shared memory)
constexpr int TM_THREADS = 512;
static_assert(TM_THREADS == MS_THREADS, "k_merge_tiles uses block_scan_u64 of k_merge_sm");
constexpr int TM_MCAP = 6144;      // members staged per chunk (a record has <= 64 members of a tile)
#ifndef SIMKA_TM_SWEEP_ENTRIES
#define SIMKA_TM_SWEEP_ENTRIES 2047  // entries between two sweeps of the shared-memory words (k_merge_tiles)
#endif
constexpr int TM_RPT = 4;          // records a thread looks at per chunk
constexpr int TM_ECAP = 1024;      // entries (records with pairs in the tile) per chunk

struct TileArgs {
    MergeArgs m;
    uint32_t bin0, bin1;           // bins [bin0, bin1) of this batch
    const uint64_t *ebin;          // [MG + 1] exclusive prefix of the entries per bin
    uint32_t *cnt;                 // scratch: abundances, region of bin b at ebin[b] - ebin[bin0]
    uint32_t *grp;                 // scratch: DENSE records of (1 + nblk) words [offset of the abundances in cnt, bitmap words]; a sub-pass of
                                   //          k_group_bins claims the space of its records from the cursor (their order does not matter)
    uint32_t *hdr;                 // [1] record cursor = records of the batch (zeroed before k_group_bins)
    uint32_t rec_cap;              // records the scratch holds: word w of record r lives at grp[w * rec_cap + r] (a tile reads two of
                                   // the words of every record: consecutive threads, consecutive addresses)
    int nblk;                      // ceil(N / 32)
    // per-CTA scratch of k_group_bins (scr_cap entries each) for bins that need several sub-passes: the bin is partitioned by
    // sub-pass once, so that every sub-pass reads its own entries instead of the whole bin
    uint64_t *scr_h;               // hashed k-mer (nullptr: no scratch, every sub-pass filters the whole bin)
    uint32_t *scr_s, *scr_c;       // sample, abundance
    uint32_t scr_cap;
};
constexpr int TM_RMAX = 512;       // sub-passes a partitioned bin may start with

static size_t group_bins_smem_bytes(int n) {
    return (size_t)MRG_SLOTS * 8 + (size_t)MRG_SLOTS * 4 * 2 + (size_t)MRG_SLOTS * 2 + (size_t)TM_GCAP * ((n + 31) / 32) * 4 + ((size_t)2 * n + 2) * 4 + 64;  // (TM_RMAX + 1 static words besides)
}

__global__ void __launch_bounds__(BLOCK)
k_bin_entries(const uint32_t *__restrict__ off, int n, uint32_t MG, uint32_t q, size_t ostride, uint32_t *__restrict__ out) {
    const uint32_t mb = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    if (mb >= MG) return;
    uint32_t acc = 0;
    for (int s = threadIdx.x & 31; s < n; s += 32) acc += off[(size_t)s * ostride + (size_t)(mb + 1) * q] - off[(size_t)s * ostride + (size_t)mb * q];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) out[mb] = acc;
}

__global__ void __launch_bounds__(BLOCK)
k_group_bins(TileArgs t) {
    const MergeArgs &a = t.m;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *tk = reinterpret_cast<uint64_t *>(smem_raw);        // [MRG_SLOTS] table keys
    uint32_t *gs = reinterpret_cast<uint32_t *>(tk + MRG_SLOTS);  // [MRG_SLOTS] group size, then packed (entry start | record index << 16)
    uint32_t *gix = gs + MRG_SLOTS;                               // [MRG_SLOTS] copy of the packed prefix
    uint32_t *bm = gix + MRG_SLOTS;                               // [TM_GCAP][nblk] sample bitmaps of the shared groups
    uint32_t *s_start = bm + (size_t)TM_GCAP * t.nblk;            // [N]
    uint32_t *s_pre = s_start + a.n;                              // [N + 1]
    uint16_t *gsz = reinterpret_cast<uint16_t *>(s_pre + a.n + 2); // [MRG_SLOTS]
    __shared__ uint32_t warp_tmp[16];
    __shared__ uint32_t sub_off[TM_RMAX + 1];  // partitioned bin: entries of sub-pass r0 at scratch [sub_off[r0], sub_off[r0 + 1])
    __shared__ uint32_t s_groups, s_shared, s_subtotal, s_base, s_ent;
    uint64_t *scr_h = t.scr_h ? t.scr_h + (size_t)blockIdx.x * t.scr_cap : nullptr;
    uint32_t *scr_s = t.scr_h ? t.scr_s + (size_t)blockIdx.x * t.scr_cap : nullptr;
    uint32_t *scr_c = t.scr_h ? t.scr_c + (size_t)blockIdx.x * t.scr_cap : nullptr;
    const int tid = threadIdx.x;
    const int N = a.n, nblk = t.nblk;
    const uint32_t q = 1u << (a.fg_bits - a.mg_bits);
    const size_t ostride = ((size_t)1 << a.fg_bits) + 1;
    ull loc_groups = 0, loc_shared = 0;

    for (uint32_t mb = t.bin0 + blockIdx.x; mb < t.bin1; mb += gridDim.x) {
        __syncthreads();
        for (int s = tid; s < N; s += BLOCK) {
            const uint32_t b0 = a.off[(size_t)s * ostride + (size_t)mb * q];
            const uint32_t b1 = a.off[(size_t)s * ostride + (size_t)(mb + 1) * q];
            s_start[s] = b0;
            s_pre[s] = b1 - b0;
        }
        if (tid == 0) s_ent = 0;
        __syncthreads();
        block_exclusive_scan(s_pre, N, warp_tmp);
        const uint32_t n_e = warp_tmp[BLOCK / 32];
        if (tid == 0) s_pre[N] = n_e;
        __syncthreads();
        const uint32_t region = (uint32_t)(t.ebin[mb] - t.ebin[t.bin0]);  // (a batch has < 2^32 entries: its abundances fit the scratch)
        uint32_t *cnt_out = t.cnt + region;
        if (n_e == 0) continue;
        uint32_t R0 = 1;
        if (n_e > MRG_CAP) { while ((uint64_t)R0 * (MRG_CAP / 2) < n_e) R0 <<= 1; }
        uint32_t R = R0, r = 0;  // hash-selected sub-passes, split when one does not fit (as in k_merge)
        // With R0 sub-passes every one of them would read (and binary-search the sample of) all entries of the bin, three times.
        // Bins that large (hundreds of samples: the offset tables cap the number of merge bins) are partitioned by their first-level
        // sub-pass into the CTA's scratch -- a counting sort, (hash, sample, abundance) per entry -- and a sub-pass then reads its own
        // range only; a split sub-pass filters that range by the further selector bits.
        const bool part = R0 > 1 && R0 <= (uint32_t)TM_RMAX && scr_h != nullptr && n_e <= t.scr_cap;
        if (part) {
            for (uint32_t i = tid; i < R0; i += BLOCK) gs[i] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < n_e; i += BLOCK) {
                int lo = 0, hi = N;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                const uint64_t hh = simka::hash2(a.eh[lo][s_start[lo] + (i - s_pre[lo])]);
                atomicAdd(&gs[(uint32_t)(hh >> 20) & (R0 - 1)], 1u);
            }
            __syncthreads();
            block_exclusive_scan(gs, (int)R0, warp_tmp);
            for (uint32_t i = tid; i < R0; i += BLOCK) sub_off[i] = gs[i];
            if (tid == 0) sub_off[R0] = n_e;
            __syncthreads();
            for (uint32_t i = tid; i < n_e; i += BLOCK) {
                int lo = 0, hi = N;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                const size_t idx = s_start[lo] + (i - s_pre[lo]);
                const uint64_t h = a.eh[lo][idx];
                const uint32_t pos = atomicAdd(&gs[(uint32_t)(simka::hash2(h) >> 20) & (R0 - 1)], 1u);
                scr_h[pos] = h;
                scr_s[pos] = (uint32_t)lo;
                scr_c[pos] = a.ec[lo][idx];
            }
        }
        bool more = true;
        while (more) {
            bool split = false;
            const uint32_t i0 = part ? sub_off[r & (R0 - 1)] : 0u, i1 = part ? sub_off[(r & (R0 - 1)) + 1] : n_e;
            __syncthreads();
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) { tk[i] = EMPTY; gs[i] = 0; }
            for (int i = tid; i < TM_GCAP * nblk; i += BLOCK) bm[i] = 0;
            if (tid == 0) { s_groups = 0; s_shared = 0; s_subtotal = 0; }
            __syncthreads();
            for (uint32_t i = i0 + tid; i < i1; i += BLOCK) {  // phase 1: insert keys, count group sizes
                uint64_t h;
                if (part) h = scr_h[i];
                else {
                    int lo = 0, hi = N;
                    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                    h = a.eh[lo][s_start[lo] + (i - s_pre[lo])];
                }
                const uint64_t hh = simka::hash2(h);
                if (((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 40) & (MRG_SLOTS - 1);
                for (int probes = 0; probes < MRG_SLOTS; probes++) {
                    ull prev = atomicCAS(reinterpret_cast<ull *>(&tk[slot]), (ull)EMPTY, (ull)h);
                    if (prev == EMPTY || prev == h) { atomicAdd(&gs[slot], 1u); break; }
                    slot = (slot + 1) & (MRG_SLOTS - 1);
                }
            }
            __syncthreads();
            uint32_t my_groups = 0, my_shared = 0, my_total = 0;  // phase 2: group statistics; only groups with >= 2 samples get a record
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) {
                const uint32_t g = gs[i];
                gsz[i] = (uint16_t)g;
                my_groups += g > 0;
                my_shared += g > 1;
                my_total += g;
                gs[i] = g > 1 ? (g | (1u << 16)) : 0;  // entries (low 16 bits) and records (high bits) scanned together
            }
            if (my_groups) atomicAdd(&s_groups, my_groups);
            if (my_shared) atomicAdd(&s_shared, my_shared);
            if (my_total) atomicAdd(&s_subtotal, my_total);
            __syncthreads();
            block_exclusive_scan(gs, MRG_SLOTS, warp_tmp);
            const uint32_t tot = warp_tmp[BLOCK / 32];
            const uint32_t n_sub = tot & 0xFFFFu, n_rec = tot >> 16;
            const uint32_t n_groups = s_groups, n_shared_groups = s_shared;
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) gix[i] = gs[i];
            __syncthreads();
#define TM_NEXT()                                                                                        \
            do {                                                                                         \
                if (split) { R <<= 1; }                                                                  \
                else {                                                                                   \
                    while (R > R0 && (r & (R >> 1))) { r -= R >> 1; R >>= 1; }                           \
                    if (R == R0) { r++; if (r >= R0) more = false; }                                     \
                    else r += R >> 1;                                                                    \
                }                                                                                        \
            } while (0)
            // split when the table is nearly full, the bitmaps of the shared groups do not fit, or the packed scan's 16 bits of entries
            // could overflow
            const bool too_big = n_groups > (uint32_t)(MRG_SLOTS * 7 / 8) || n_rec > (uint32_t)TM_GCAP || s_subtotal > 0xFFFFu;
            if (too_big && R < (1u << 24)) { split = true; TM_NEXT(); continue; }
            if (too_big) { if (tid == 0) *a.error_flag = 4; TM_NEXT(); continue; }
            if (tid == 0) { loc_groups += n_groups; loc_shared += n_shared_groups; }
            if (n_rec == 0) { TM_NEXT(); continue; }
            (void)n_sub;
            const uint32_t ent0 = s_ent;  // abundances this bin has written so far
            if (tid == 0) s_base = atomicAdd(t.hdr, n_rec);  // (read after the barrier below)
            // phase 3a: sample bitmaps of the shared groups
            for (uint32_t i = i0 + tid; i < i1; i += BLOCK) {
                uint64_t h;
                int lo = 0;
                if (part) { h = scr_h[i]; lo = (int)scr_s[i]; }
                else {
                    int hi = N;
                    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                    h = a.eh[lo][s_start[lo] + (i - s_pre[lo])];
                }
                const uint64_t hh = simka::hash2(h);
                if (((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 40) & (MRG_SLOTS - 1);
                int pr = 0;
                while (tk[slot] != h && ++pr <= MRG_SLOTS) slot = (slot + 1) & (MRG_SLOTS - 1);
                if (pr > MRG_SLOTS) { *a.error_flag = 7; continue; }
                if (gsz[slot] < 2) continue;
                atomicOr(&bm[(gix[slot] >> 16) * nblk + (lo >> 5)], 1u << (lo & 31));
            }
            __syncthreads();
            // phase 3b: abundances in sample order (rank = set bits below the sample), records
            for (uint32_t i = i0 + tid; i < i1; i += BLOCK) {
                uint64_t h;
                uint32_t c;
                int lo = 0;
                if (part) { h = scr_h[i]; lo = (int)scr_s[i]; c = scr_c[i]; }
                else {
                    int hi = N;
                    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                    const size_t idx = s_start[lo] + (i - s_pre[lo]);
                    h = a.eh[lo][idx];
                    c = a.ec[lo][idx];
                }
                const uint64_t hh = simka::hash2(h);
                if (((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 40) & (MRG_SLOTS - 1);
                int pr = 0;
                while (tk[slot] != h && ++pr <= MRG_SLOTS) slot = (slot + 1) & (MRG_SLOTS - 1);
                if (pr > MRG_SLOTS || gsz[slot] < 2) continue;
                const uint32_t pk = gix[slot];
                const uint32_t *w = bm + (pk >> 16) * nblk;
                uint32_t rank = __popc(w[lo >> 5] & ((1u << (lo & 31)) - 1));
                for (int b = 0; b < (lo >> 5); b++) rank += __popc(w[b]);
                cnt_out[ent0 + (pk & 0xFFFFu) + rank] = c;
            }
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) {
                if (gsz[i] < 2) continue;
                const uint32_t pk = gix[i];
                uint32_t *rec = t.grp + s_base + (pk >> 16);
                rec[0] = region + ent0 + (pk & 0xFFFFu);
                for (int b = 0; b < nblk; b++) rec[(uint64_t)(1 + b) * t.rec_cap] = bm[(pk >> 16) * nblk + b];
            }
            __syncthreads();
            if (tid == 0) s_ent = ent0 + (tot & 0xFFFFu);
            TM_NEXT();
        }
#undef TM_NEXT
    }
    if (tid == 0 && (loc_groups | loc_shared)) {
        atomicAdd(&a.scalars[0], loc_groups);
        atomicAdd(&a.scalars[1], loc_shared);
    }
}

// ---------------------------------------------------------------------------------------------
// tile accumulation
// ---------------------------------------------------------------------------------------------
struct TMSmem {
    alignas(16) double2 logtab[128];
    alignas(16) unsigned long long scan_tmp[TM_THREADS / 32 + 1];
    MSTotals T;                       // per-sample constants of the tile's 64 local sample slots
    uint32_t cg[TM_ECAP];             // per entry of the chunk: member start (16 bits) | members of block I (8 bits) | of block J (8 bits)
    uint32_t cgp[TM_ECAP + 1];        // pair prefix
    uint32_t me_c[TM_MCAP];           // staged members: abundance
    uint8_t me_s[TM_MCAP];            //                 local sample slot (0..31 block I, 32..63 block J)
    // followed by: uint32_t acc[MS_NACC][1024]
};
static_assert(MS_MAXN == 64, "a tile has 2 x 32 local sample slots");
static size_t merge_tiles_smem_bytes() { return ((sizeof(TMSmem) + 15) & ~(size_t)15) + (size_t)MS_NACC * 1024 * 4 + 16; }

// the work of one CTA; DIAG (tile of a block with itself: 496 pair keys instead of 1024) is a template parameter so that the number
// of keys -- the stride between the ten accumulator words of a key -- is a compile-time constant
template <bool SIMPLE, bool COMPLEX, bool DIAG>
__device__ __forceinline__ void merge_tile(const TileArgs &t, const uint32_t slices, const uint32_t I, const uint32_t J, const uint32_t slice,
                                           unsigned char *smem_raw) {
    const MergeArgs &a = t.m;
    TMSmem &S = *reinterpret_cast<TMSmem *>(smem_raw);
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(TMSmem) + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    const int N = a.n;
    constexpr bool diag = DIAG;
    constexpr int NPk = DIAG ? 496 : 1024;
    if (tid < 128) S.logtab[tid] = a.logtab[tid];
    if (tid == 0) S.T.big_tot = 0;
    __syncthreads();
    for (int i = tid; i < MS_NACC * NPk; i += TM_THREADS) acc[i] = 0;
    if (tid < 64) {
        const int gsid = tid < 32 ? (int)(32 * I) + tid : (diag ? -1 : (int)(32 * J) + tid - 32);
        const bool used = gsid >= 0 && gsid < N;
        S.T.s_gid[tid] = used ? gsid : -1;
        const double nt = used ? a.ntot[gsid] : 0.0;
        S.T.ntot[tid] = nt;
        S.T.ninv30[tid] = nt > 0.0 ? 1073741824.0 / nt : 0.0;
        S.T.nlo[tid] = nt < 4294967296.0 ? (uint32_t)nt : 0u;
        if (nt >= 4294967296.0) S.T.big_tot = 1;
    }
    __syncthreads();
    auto key_slots = [&](int key, int &li, int &lj) {
        if (!diag) { li = key >> 5; lj = 32 + (key & 31); }
        else {
            li = 0;
            while ((li + 1) * (2 * 32 - (li + 1) - 1) / 2 <= key) li++;
            lj = li + 1 + (key - li * (2 * 32 - li - 1) / 2);
        }
    };
    ull loc_pairs = 0;
    const uint32_t n_all = t.hdr[0];  // records of the batch, split evenly over the slices of the tile
    const uint32_t per = (n_all + slices - 1) / slices;
    const uint32_t r_lo = min(n_all, slice * per), n_rec = min(n_all, r_lo + per) - r_lo;
    uint32_t rpt = TM_RPT, step = TM_THREADS, since = 0;
    {
        const uint32_t *cnt_in = t.cnt;
        const uint32_t *grp_in = t.grp + r_lo;
        const uint64_t wcap = t.rec_cap;
        // A chunk is rpt * step consecutive records (step < TM_THREADS only with rpt == 1): every thread looks at up to TM_RPT records and
        // only those with pairs IN THIS TILE become entries -- with many sample blocks most records have none, and a chunk of one record
        // per thread would spend its barriers on a handful of pairs.  The shape adapts: it shrinks until the entries (<= TM_ECAP) and
        // their members (<= TM_MCAP) fit, and grows back when a chunk filled less than half of both.
        for (uint32_t c0 = 0; c0 < n_rec;) {
            uint32_t pm[TM_RPT];  // per record: pairs << 8 | members of the tile (0: not an entry)
#pragma unroll
            for (int k = 0; k < TM_RPT; k++) {
                pm[k] = 0;
                const uint32_t rix = c0 + (uint32_t)k * TM_THREADS + tid;
                if ((uint32_t)k < rpt && (uint32_t)tid < step && rix < n_rec) {
                    const uint32_t *rec = grp_in + rix;
                    const uint32_t nI = __popc(rec[(1 + I) * wcap]), nJ = diag ? 0u : __popc(rec[(1 + J) * wcap]);
                    const uint32_t pairs = diag ? nI * (nI - 1) / 2 : nI * nJ;
                    if (pairs) pm[k] = (pairs << 8) | (nI + nJ);
                }
            }
            unsigned long long total, ex;
            for (;;) {
                unsigned long long v = 0;  // pairs << 32 | entries << 20 | members
#pragma unroll
                for (int k = 0; k < TM_RPT; k++)
                    if ((uint32_t)k < rpt && (uint32_t)tid < step && pm[k]) v += ((unsigned long long)(pm[k] >> 8) << 32) | (1u << 20) | (pm[k] & 0xFF);
                ex = block_scan_u64(v, S.scan_tmp, &total);
                if (((uint32_t)total & 0xFFFFF) <= (uint32_t)TM_MCAP && (((uint32_t)total >> 20) & 0xFFF) <= (uint32_t)TM_ECAP) break;
                if (rpt > 1) rpt >>= 1; else step >>= 1;
            }
            const uint32_t P = (uint32_t)(total >> 32), E = ((uint32_t)total >> 20) & 0xFFF, M = (uint32_t)total & 0xFFFFF;
            // sweep (CTA-uniform decision): a word holds < 2^31 after a sweep and an entry adds < 2^20 to a word of a key at most once,
            // so 2047 entries may pass between two sweeps before a word could wrap (see MS_NACC)
            if (since + E > (uint32_t)SIMKA_TM_SWEEP_ENTRIES) {
                for (int i = tid; i < MS_NACC * NPk; i += TM_THREADS) {
                    const uint32_t v = acc[i];
                    if (v & 0x80000000u) {
                        int li, lj;
                        key_slots(i % NPk, li, lj);
                        ms_flush_word(a, i / NPk, S.T.s_gid[li], S.T.s_gid[lj], v);
                        acc[i] = 0;
                    }
                }
                since = 0;
            }
            since += E;
            if (tid == 0) { S.cgp[E] = P; loc_pairs += P; }
            {  // stage the entries of this thread: abundances are in sample order, the rank of a member is its bit's position among the set bits
                uint32_t e = ((uint32_t)ex >> 20) & 0xFFF, pos = (uint32_t)ex & 0xFFFFF, pp = (uint32_t)(ex >> 32);
#pragma unroll
                for (int k = 0; k < TM_RPT; k++) {
                    if (!((uint32_t)k < rpt && (uint32_t)tid < step && pm[k])) continue;
                    const uint32_t *rec = grp_in + (c0 + (uint32_t)k * TM_THREADS + tid);
                    const uint32_t e0 = rec[0];
                    uint32_t wI = 0, wJ = 0, offI = 0, offJ = 0, run = 0;
                    for (uint32_t b = 0; b <= J; b++) {
                        const uint32_t w = rec[(1 + b) * wcap];
                        if (b == I) { wI = w; offI = run; }
                        if (b == J) { wJ = w; offJ = run; }
                        run += __popc(w);
                    }
                    const uint32_t nI = __popc(wI), nJ = diag ? 0u : __popc(wJ);
                    S.cg[e] = pos | (nI << 16) | (nJ << 24);
                    S.cgp[e] = pp;
                    e++;
                    pp += pm[k] >> 8;
                    uint32_t w = wI, q = 0;
                    while (w) { const int bit = __ffs((int)w) - 1; w &= w - 1; S.me_s[pos] = (uint8_t)bit; S.me_c[pos] = cnt_in[e0 + offI + q]; pos++; q++; }
                    w = diag ? 0u : wJ; q = 0;  // (a diagonal tile has one block: its members are staged once)
                    while (w) { const int bit = __ffs((int)w) - 1; w &= w - 1; S.me_s[pos] = (uint8_t)(32 + bit); S.me_c[pos] = cnt_in[e0 + offJ + q]; pos++; q++; }
                }
            }
            __syncthreads();
            // flat pair enumeration, one contiguous run of pairs per thread (as k_merge_sm)
            const uint32_t perp = (P + TM_THREADS - 1) / TM_THREADS;
            const uint32_t l0 = min(P, tid * perp), l1 = min(P, l0 + perp);
            if (l0 < l1) {
                int lo = 0, hi = (int)E;  // cgp[lo] <= l0 < cgp[hi]
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.cgp[mid] <= l0) lo = mid; else hi = mid; }
                uint32_t g = (uint32_t)lo;
                uint32_t cgv = S.cg[g];
                uint32_t st = cgv & 0xFFFF, gI = (cgv >> 16) & 0xFF, gJ = cgv >> 24;
                uint32_t l = l0 - S.cgp[g];
                uint32_t x, y;  // indices into the entry's staged members: x in block I; y in block J (or > x in block I on the diagonal)
                if (!diag) { x = l / gJ; y = gI + l % gJ; }
                else {
                    const float tg = 2.0f * (float)gI - 1.0f;
                    int xi = (int)((tg - sqrtf(fmaxf(tg * tg - 8.0f * (float)l, 0.0f))) * 0.5f);
                    if (xi > (int)gI - 2) xi = (int)gI - 2;
                    if (xi < 0) xi = 0;
                    while ((uint32_t)(xi + 1) * (2 * gI - (xi + 1) - 1) / 2 <= l) xi++;
                    while ((uint32_t)xi * (2 * gI - xi - 1) / 2 > l) xi--;
                    x = (uint32_t)xi;
                    y = x + 1 + (l - x * (2 * gI - x - 1) / 2);
                }
                for (uint32_t p = l0; p < l1; p++) {
                    const uint32_t si = S.me_s[st + x], sj = S.me_s[st + y];  // local slots, si < sj by construction
                    const uint32_t ci = S.me_c[st + x], cj = S.me_c[st + y];
                    const uint32_t key = diag ? si * (63 - si) / 2 + (sj - si - 1) : si * 32 + (sj - 32);
                    ms_pair<SIMPLE, COMPLEX, NPk>(a, acc, NPk, key, ci, cj, si, sj, S.T, S.logtab);
                    // advance inside the entry, then to the next one (every entry has pairs)
                    bool next_rec = false;
                    if (!diag) { if (++y == gI + gJ) { y = gI; if (++x == gI) next_rec = true; } }
                    else { if (++y == gI) { if (++x == gI - 1) next_rec = true; y = x + 1; } }
                    if (next_rec && p + 1 < l1) {
                        g++;
                        cgv = S.cg[g]; st = cgv & 0xFFFF; gI = (cgv >> 16) & 0xFF; gJ = cgv >> 24;
                        x = 0; y = diag ? 1 : gI;
                    }
                }
            }
            // (no barrier here: the scan of the next chunk has three before anything staged is overwritten)
            c0 += rpt * step;
            if (2 * M <= (uint32_t)TM_MCAP && 2 * E <= (uint32_t)TM_ECAP) {
                if (step < (uint32_t)TM_THREADS) step <<= 1;
                else if (rpt < (uint32_t)TM_RPT) rpt <<= 1;
            }
        }
    }
    __syncthreads();
    for (int key = tid; key < NPk; key += TM_THREADS) {
        int li = -1, lj = -1;
        for (int w = 0; w < MS_NACC; w++) {
            const uint32_t v = acc[w * NPk + key];
            if (!v) continue;
            if (li < 0) key_slots(key, li, lj);
            ms_flush_word(a, w, S.T.s_gid[li], S.T.s_gid[lj], v);
        }
    }
    if (tid == 0 && loc_pairs) atomicAdd(&a.scalars[2], loc_pairs);
}

template <bool SIMPLE, bool COMPLEX>
__global__ void __launch_bounds__(TM_THREADS, 2)
k_merge_tiles(TileArgs t, uint32_t slices) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tile = blockIdx.x / slices, slice = blockIdx.x % slices, nblk = (uint32_t)t.nblk;
    uint32_t I = 0, J = 0;
    { uint32_t rem = tile; while (rem >= nblk - I) { rem -= nblk - I; I++; } J = I + rem; }
    if (I == J) merge_tile<SIMPLE, COMPLEX, true>(t, slices, I, J, slice, smem_raw);
    else merge_tile<SIMPLE, COMPLEX, false>(t, slices, I, J, slice, smem_raw);
}
