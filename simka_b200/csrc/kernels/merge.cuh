// kernels/merge.cuh -- stage 2 (simkaMerge): multi-sample merge + N x N accumulator update
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
#pragma once

// ---------------------------------------------------------------------------------------------
// K4: multi-sample merge + N x N accumulator update.
// Replaces the k-way merge of SimkaMergeAlgorithm::execute (src/SimkaMerge.cpp:1164-1264), the insert
// gate (:1307-1326) and SimkaCountProcessorSimple::updateDistance{Default,Simple,Complex}
// (src/core/SimkaAlgorithm.hpp:341-516).  One CTA per merge bin: the N per-sample segments of the bin
// are grouped by k-mer in a shared-memory table (group = the non-zero entries of the reference's
// abundancePerBank vector), then sample pairs of every group are enumerated flat over the threads.
//
// One-sided terms of the complex distances (pairs where exactly one of c_i,c_j is 0) are NOT
// enumerated (that is O(N) per k-mer): Canberra and Jensen-Shannon get them in closed form from
// nb_shared / the per-sample totals at read-out, Whittaker from the per-sample abundance histogram
// (k_whittaker_onesided) minus the both-present corrections accumulated here.
// ---------------------------------------------------------------------------------------------
struct MergeArgs {
    int n;                       // samples
    int fg_bits, mg_bits;        // fine-range and merge-bin granularity (log2)
    const uint64_t *const *eh;   // [n] per-sample hashed k-mers
    const uint32_t *const *ec;   // [n] per-sample counts
    const uint32_t *off;         // [n][FG+1]
    const double *ntot;          // [n] global solid k-mer totals N_i as double
    const double2 *logtab;       // [128] tab_log table
    int simple, complex_;
    int count_only;              // k_merge: only the distinct / shared k-mer counts (the tiled merge does the pairs)
    ull *nb_shared, *distinct_shared, *braycurtis, *chord, *hellinger, *whittaker;
    double *kl;
    ull *klfix;                  // Jensen-Shannon sums in fixed point (units of 2^-60), exact integer adds
    ull *scalars;                // [0] distinct k-mers, [1] shared k-mers, [2] pairs
    int *error_flag;
};

__device__ __forceinline__ long long absint32_ext(uint64_t x) {  // abs((int)x) as the reference's int -> u64 add
    int d = (int)(uint32_t)x;
    d = (int)(d < 0 ? 0u - (unsigned)d : (unsigned)d);
    return (long long)d;
}

// log(x) for normal x > 0 from a 128-entry table {1/c, log c} (interval centres c over [0.6875, 1.375), the two
// intervals around 1 use c = 1 so that the result keeps its RELATIVE accuracy near x = 1 and log(1) == 0):
// log x = k ln2 + log c + log1p(r), r = x/c - 1, |r| < 2^-7, degree-6 series.  |error| < 4e-14 relative
// (checked against logl over 2e7 points); used for the Jensen-Shannon terms whose sum is compared at 1e-10.
__device__ __forceinline__ double tab_log(double x, const double2 *T) {
    const unsigned long long OFF = 0x3fe6000000000000ull;
    const unsigned long long ix = (unsigned long long)__double_as_longlong(x);
    const unsigned long long tmp = ix - OFF;
    const int i = (int)(tmp >> 45) & 127;
    const int k = (int)((long long)tmp >> 52);
    const double z = __longlong_as_double((long long)(ix - (tmp & (0xfffull << 52))));
    const double2 t = T[i];
    const double r = fma(z, t.x, -1.0);
    const double w = fma((double)k, 0.693147180559945309417232121458, t.y);
    double p = fma(r, -1.0 / 6, 0.2);
    p = fma(r, p, -0.25);
    p = fma(r, p, 1.0 / 3);
    p = fma(r, p, -0.5);
    return w + fma(r * r, p, r);
}

// (u64)sqrt((double)prod): what the reference's u64 Hellinger accumulator gains per k-mer
// (src/core/SimkaAlgorithm.hpp:395).  Below 2^32 the exact integer square root is the same value.
__device__ __forceinline__ ull hellinger_term(ull prod) {
    if (prod >> 32) return (ull)sqrt((double)prod);
    const uint32_t p = (uint32_t)prod;
    const float pf = (float)p;
    uint32_t s = min((uint32_t)(pf * rsqrtf(pf)), 65535u);  // within +-1 of the root (prod >= 1)
    if (s * s > p) s--;
    else if (s < 65535u && (s + 1) * (s + 1) <= p) s++;
    return s;
}

__device__ __forceinline__ void accumulate_pair(const MergeArgs &a, int i, int j, uint32_t ci, uint32_t cj) {
    // i < j
    const size_t n = (size_t)a.n;
    const size_t ij = (size_t)i * n + j, ji = (size_t)j * n + i;
    const size_t sym = (size_t)j + (n - 1) * i - ((size_t)i * (i - 1) / 2);
    const uint32_t mn = ci < cj ? ci : cj;
    atomicAdd(&a.nb_shared[ij], (ull)ci);
    atomicAdd(&a.nb_shared[ji], (ull)cj);
    atomicAdd(&a.distinct_shared[sym], 1ull);
    atomicAdd(&a.braycurtis[sym], (ull)mn);
    if (a.simple) {
        const ull prod = (ull)ci * (ull)cj;
        atomicAdd(&a.chord[ij], prod);
        atomicAdd(&a.hellinger[ij], hellinger_term(prod));  // u64 += double truncates per increment
    }
    if (a.complex_) {
        const double ai = (double)ci, aj = (double)cj, Ni = a.ntot[i], Nj = a.ntot[j];
        const double yX = aj * Ni, xY = ai * Nj;
        if (xY != yX) {  // equal: both logs are log(1) = 0 exactly (same evaluation as k_merge_sm)
            const double r1 = (2 * xY) * __drcp_rn(xY + yX);
            atomicAdd(&a.kl[ij], (ai / Ni) * tab_log(r1, a.logtab) + (aj / Nj) * tab_log(2.0 - r1, a.logtab));
        }
        const uint64_t wx = (uint64_t)xY, wy = (uint64_t)yX;
        const long long two = absint32_ext(wx - wy);
        const long long sub = absint32_ext(wx) + absint32_ext(wy);  // their one-sided values, removed here
        atomicAdd(&a.whittaker[ij], (ull)(two - sub));
    }
}

__global__ void __launch_bounds__(BLOCK)
k_merge(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *tk = reinterpret_cast<uint64_t *>(smem_raw);      // [MRG_SLOTS] table keys
    uint32_t *gs = reinterpret_cast<uint32_t *>(tk + MRG_SLOTS); // [MRG_SLOTS] group size, then group start
    uint32_t *pp = gs + MRG_SLOTS;                               // [MRG_SLOTS+1] fill cursors, then pair prefix
    uint32_t *ge_c = pp + MRG_SLOTS + 1;                         // [MRG_CAP] grouped counts
    uint32_t *s_start = ge_c + MRG_CAP;                          // [N]
    uint32_t *s_pre = s_start + a.n;                             // [N+1]
    uint16_t *gsz = reinterpret_cast<uint16_t *>(s_pre + a.n + 1); // [MRG_SLOTS] group size copy
    uint16_t *ge_s = gsz + MRG_SLOTS;                            // [MRG_CAP] grouped sample ids
    __shared__ uint32_t warp_tmp[16];
    __shared__ uint32_t s_groups, s_shared;

    const int tid = threadIdx.x;
    const int N = a.n;
    const uint32_t MG = 1u << a.mg_bits;
    const uint32_t q = 1u << (a.fg_bits - a.mg_bits);
    const size_t ostride = ((size_t)1 << a.fg_bits) + 1;
    ull loc_groups = 0, loc_shared = 0;

    for (uint32_t mb = blockIdx.x; mb < MG; mb += gridDim.x) {
        __syncthreads();
        for (int s = tid; s < N; s += BLOCK) {
            const uint32_t b0 = a.off[(size_t)s * ostride + (size_t)mb * q];
            const uint32_t b1 = a.off[(size_t)s * ostride + (size_t)(mb + 1) * q];
            s_start[s] = b0;
            s_pre[s] = b1 - b0;
        }
        __syncthreads();
        block_exclusive_scan(s_pre, N, warp_tmp);
        const uint32_t n_e = warp_tmp[BLOCK / 32];
        if (tid == 0) s_pre[N] = n_e;
        __syncthreads();
        if (n_e == 0) continue;
        uint32_t R0 = 1;
        if (n_e > MRG_CAP) { while ((uint64_t)R0 * (MRG_CAP / 2) < n_e) R0 <<= 1; }
        // Sub-passes select the keys with hash bits == r (mod R).  The entries of one k-mer stay together, so with hundreds of
        // samples a sub-pass can still hold more grouped entries than fit (a few k-mers shared by everybody): such a sub-pass is
        // split in two (R doubled) BEFORE anything of it is accumulated -- a depth-first walk over the selector tree.
        uint32_t R = R0, r = 0;
        bool more = true;
        while (more) {
            bool split = false;
            __syncthreads();  // pair enumeration of the previous sub-pass still reads the group arrays
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) { tk[i] = EMPTY; gs[i] = 0; pp[i] = 0; }
            if (tid == 0) { s_groups = 0; s_shared = 0; }
            __syncthreads();
            // phase 1: insert keys, count group sizes
            for (uint32_t i = tid; i < n_e; i += BLOCK) {
                int lo = 0, hi = N;  // s_pre[lo] <= i < s_pre[hi]
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                const uint64_t h = a.eh[lo][s_start[lo] + (i - s_pre[lo])];
                const uint64_t hh = simka::hash2(h);
                if (((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 40) & (MRG_SLOTS - 1);
                int probes = 0;
                for (;;) {
                    ull prev = atomicCAS(reinterpret_cast<ull *>(&tk[slot]), (ull)EMPTY, (ull)h);
                    if (prev == EMPTY || prev == h) { atomicAdd(&gs[slot], 1u); break; }
                    slot = (slot + 1) & (MRG_SLOTS - 1);
                    if (++probes >= MRG_SLOTS) break;  // table full: this sub-pass is split below, nothing of it is kept
                }
            }
            __syncthreads();
            // phase 2: group statistics; only groups with >= 2 samples are materialised
            uint32_t my_groups = 0, my_shared = 0;
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) {
                const uint32_t g = gs[i];
                gsz[i] = (uint16_t)g;
                my_groups += g > 0;
                my_shared += g > 1;
                if (g < 2) gs[i] = 0;
            }
            if (my_groups) atomicAdd(&s_groups, my_groups);
            if (my_shared) atomicAdd(&s_shared, my_shared);
            __syncthreads();
            block_exclusive_scan(gs, MRG_SLOTS, warp_tmp);
            const uint32_t n_sub = warp_tmp[BLOCK / 32];
            const uint32_t n_shared_groups = s_shared;
            __syncthreads();
            // what follows the selection of the next sub-pass (the statements at the loop's end) for every way out of this one
#define MRG_NEXT_SUBPASS()                                                                                       \
            do {                                                                                                 \
                if (split) { R <<= 1; }                                                                          \
                else {                                                                                           \
                    while (R > R0 && (r & (R >> 1))) { r -= R >> 1; R >>= 1; }                                   \
                    if (R == R0) { r++; if (r >= R0) more = false; }                                             \
                    else r += R >> 1;                                                                            \
                }                                                                                                \
            } while (0)
            if (n_sub > MRG_CAP || s_groups > (uint32_t)(MRG_SLOTS * 7 / 8)) {
                if (R >= (1u << 24)) { if (tid == 0) *a.error_flag = 4; }
                else { split = true; MRG_NEXT_SUBPASS(); continue; }
            }
            if (tid == 0) { loc_groups += s_groups; loc_shared += s_shared; }
            if (a.count_only || n_sub > MRG_CAP || n_shared_groups == 0) { MRG_NEXT_SUBPASS(); continue; }  // (nothing shared in this bin: no accumulator changes)
            // phase 3: scatter the entries of shared k-mers into their groups
            for (uint32_t i = tid; i < n_e; i += BLOCK) {
                int lo = 0, hi = N;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                const size_t idx = s_start[lo] + (i - s_pre[lo]);
                const uint64_t h = a.eh[lo][idx];
                const uint64_t hh = simka::hash2(h);
                if (((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 40) & (MRG_SLOTS - 1);
                int pr = 0;
                while (tk[slot] != h && ++pr <= MRG_SLOTS) slot = (slot + 1) & (MRG_SLOTS - 1);
                if (pr > MRG_SLOTS) { *a.error_flag = 7; continue; }
                if (gsz[slot] < 2) continue;  // private k-mer: contributes to no pair
                const uint32_t pos = gs[slot] + atomicAdd(&pp[slot], 1u);
                ge_s[pos] = (uint16_t)lo;
                ge_c[pos] = a.ec[lo][idx];
            }
            __syncthreads();
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) { const uint32_t g = gsz[i]; pp[i] = g * (g - 1) / 2; }  // g=0 -> 0
            __syncthreads();
            block_exclusive_scan(pp, MRG_SLOTS, warp_tmp);
            const uint32_t P = warp_tmp[BLOCK / 32];
            if (tid == 0) pp[MRG_SLOTS] = P;
            __syncthreads();
            // phase 4: flat pair enumeration.  Consecutive lanes take consecutive pairs on purpose: they share the group and
            // mostly the row, so their global atomics fall into the same few L2 lines (one run of pairs per thread was
            // measured 2x slower at N = 128).
            for (uint32_t p = tid; p < P; p += BLOCK) {
                int lo = 0, hi = MRG_SLOTS;  // pp[lo] <= p < pp[hi]
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pp[mid] <= p) lo = mid; else hi = mid; }
                const uint32_t g = gsz[lo], l = p - pp[lo], base = gs[lo];
                // row-major strict upper triangle of a g x g matrix: row x starts at x(2g-x-1)/2
                const float tg = 2.0f * (float)g - 1.0f;  // g <= N <= 8192: exact in float up to the +-1 fix-up below
                int x = (int)((tg - sqrtf(fmaxf(tg * tg - 8.0f * (float)l, 0.0f))) * 0.5f);
                if (x > (int)g - 2) x = (int)g - 2;
                if (x < 0) x = 0;
                while ((uint32_t)(x + 1) * (2 * g - (x + 1) - 1) / 2 <= l) x++;
                while ((uint32_t)x * (2 * g - x - 1) / 2 > l) x--;
                const uint32_t y = x + 1 + (l - (uint32_t)x * (2 * g - x - 1) / 2);
                int si = ge_s[base + x], sj = ge_s[base + y];
                uint32_t ci = ge_c[base + x], cj = ge_c[base + y];
                if (si > sj) { int ts = si; si = sj; sj = ts; uint32_t tc = ci; ci = cj; cj = tc; }
                accumulate_pair(a, si, sj, ci, cj);
            }
            if (tid == 0) atomicAdd(&a.scalars[2], (ull)P);
            MRG_NEXT_SUBPASS();
        }
#undef MRG_NEXT_SUBPASS
    }
    if (tid == 0 && (loc_groups | loc_shared)) {
        atomicAdd(&a.scalars[0], loc_groups);
        atomicAdd(&a.scalars[1], loc_shared);
    }
}

#ifndef SIMKA_MS_THREADS
#define SIMKA_MS_THREADS 512
#endif
#ifndef SIMKA_MS_MINB
#define SIMKA_MS_MINB 2
#endif
constexpr int MS_THREADS = SIMKA_MS_THREADS;
constexpr uint32_t MS_SKIP = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned long long block_scan_u64(unsigned long long v, unsigned long long *tmp, unsigned long long *total) {
    // exclusive scan over the MS_THREADS threads of the block; tmp has MS_THREADS/32+1 entries
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    __syncthreads();  // tmp may still be read from a previous scan
    if (lane == 31) tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long w = lane < MS_THREADS / 32 ? tmp[lane] : 0, winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += u;
        }
        if (lane < MS_THREADS / 32) tmp[lane] = winc - w;
        if (lane == MS_THREADS / 32 - 1) tmp[MS_THREADS / 32] = winc;
    }
    __syncthreads();
    *total = tmp[MS_THREADS / 32];
    return tmp[wid] + inc - v;
}

// ---------------------------------------------------------------------------------------------
// K4 for N <= 64 samples, the shipped merge: same grouping as k_merge, then the sample pairs of the
// bin are enumerated FLAT (every thread takes a contiguous run of pairs: one search + decode per run, full
// lane utilisation) and added to CTA-private N(N-1)/2-entry accumulators in shared memory with 32-bit
// shared atomics (64-bit values as lo/hi words with an explicit carry).  One flush of the accumulators per
// CTA at the end.  (An earlier variant counting-sorted the pairs by key so that a thread could own a key and keep the
// accumulators in registers without atomics; it lost two thirds of the lanes to uneven per-key list lengths: 28.9 ms
// against 24.2 ms for this one)
// ---------------------------------------------------------------------------------------------
constexpr int MS_SLOTS = 2048;
constexpr int MS_ECAP = 2048;
constexpr int MS_GCAP = 1024;
constexpr int MS_MAXN = 64;
// 32-bit accumulator words per pair key in shared memory: sh_ij, sh_ji, distinct, bc, chord, hellinger, whittaker (negated) as
// (low 20 bits | units of 2^20), Jensen-Shannon (2^-60 fixed point) as (low 20 bits | units of 2^20).  Every add is below 2^20 and
// needs no return value (a 64-bit shared atomicAdd is a CAS loop, and a lo/hi pair with an explicit carry needs the old value
// back: 16 dependent atomics per pair in the first version).  A word that reaches 2^31 is moved to the global accumulator by the
// sweep after each pass (at most MS_GCAP pairs per key and pass: < 2^30 added per pass, so a word never wraps); the rare
// increment of 2^20 or more (abundances in the millions) goes straight to the global accumulator.
constexpr int MS_NACC = 10;
constexpr uint32_t MS_FAST = 1u << 20;

// per-sample constants of the (<= 64) local sample slots of a CTA, in shared memory
struct MSTotals {
    alignas(16) double ntot[MS_MAXN];      // N_i of the local sample slots
    alignas(16) double ninv30[MS_MAXN];    // 2^30 / N_i: the product of two of them scales a Jensen-Shannon term to 2^-60 units
    uint32_t nlo[MS_MAXN];                 // N_i mod 2^32
    int s_gid[MS_MAXN];                    // local sample slot -> global sample id (-1: unused)
    int big_tot;                           // some N_i >= 2^32: the 32-bit Whittaker shortcut is off
};

struct MSSmem {
    alignas(16) uint64_t tk[MS_SLOTS];
    alignas(16) double2 logtab[128];
    alignas(16) unsigned long long scan_tmp[MS_THREADS / 32 + 1];
    alignas(16) uint32_t gcnt[MS_SLOTS];   // cleaned with 128-bit stores
    uint32_t cg[MS_GCAP];
    uint32_t cgp[MS_GCAP + 1];
    uint32_t ge_c[MS_ECAP];
    uint32_t s_start[MS_MAXN], s_pre[MS_MAXN + 1];
    MSTotals T;
    uint8_t ge_s[MS_ECAP];
    // followed by: uint32_t acc[MS_NACC][NP];
};
static size_t ms_smem_bytes(int n) {
    const size_t np = n > MS_MAXN ? 1024 : (size_t)n * (n - 1) / 2;  // tiled: 32 x 32 pair keys
    return ((sizeof(MSSmem) + 15) & ~(size_t)15) + (size_t)MS_NACC * np * 4 + 16;
}

// word w of pair key `key` (local sample slots li < lj of the tile) -> the global accumulators, value v in the word's unit
__device__ __forceinline__ void ms_flush_word(const MergeArgs &a, int w, int i, int j, ull v) {
    const size_t n = (size_t)a.n;
    const size_t ij = (size_t)i * n + j, ji = (size_t)j * n + i;
    const size_t sym = (size_t)j + (n - 1) * i - ((size_t)i * (i - 1) / 2);
    switch (w) {
    case 0: atomicAdd(&a.nb_shared[ij], v); break;
    case 1: atomicAdd(&a.nb_shared[ji], v); break;
    case 2: atomicAdd(&a.distinct_shared[sym], v); break;
    case 3: atomicAdd(&a.braycurtis[sym], v); break;
    case 4: atomicAdd(&a.chord[ij], v); break;
    case 5: atomicAdd(&a.hellinger[ij], v); break;
    case 6: atomicAdd(&a.whittaker[ij], 0ull - v); break;           // accumulated negated
    case 7: atomicAdd(&a.whittaker[ij], 0ull - (v << 20)); break;
    case 8: atomicAdd(&a.klfix[ij], v); break;
    default: atomicAdd(&a.klfix[ij], v << 20); break;
    }
}

// One sample pair (local slots si < sj, abundances ci, cj) of one k-mer into the shared-memory accumulator words of its key
// (src/core/SimkaAlgorithm.hpp:341-516).  NPK > 0: the number of keys is a compile-time constant (the ten word addresses are
// immediate offsets of one base).  Common case first: both abundances below 2^20, every add but the chord product is then below
// 2^20 by construction and goes to the 32-bit words without a check.
template <bool SIMPLE, bool COMPLEX, int NPK>
__device__ __forceinline__ void ms_pair(const MergeArgs &a, uint32_t *acc, int npk_rt, uint32_t key, uint32_t ci, uint32_t cj, uint32_t si,
                                        uint32_t sj, const MSTotals &T, const double2 *logtab) {
    const int NPk = NPK > 0 ? NPK : npk_rt;
    uint32_t *p = acc + key;
    auto add = [&](int w, ull v) {
        if (v < MS_FAST) atomicAdd(p + w * NPk, (uint32_t)v);
        else ms_flush_word(a, w, T.s_gid[si], T.s_gid[sj], v);
    };
    auto add2 = [&](int w, ull v) {  // 64-bit value as (low 20 bits | units of 2^20); a term that came out negative: 64-bit modular add
        if (v < ((ull)MS_FAST << 20)) { atomicAdd(p + w * NPk, (uint32_t)v & (MS_FAST - 1)); atomicAdd(p + (w + 1) * NPk, (uint32_t)(v >> 20)); }
        else ms_flush_word(a, w, T.s_gid[si], T.s_gid[sj], v);
    };
    const bool small = ((ci | cj) >> 20) == 0;
    if (small) {
        atomicAdd(p, ci);
        atomicAdd(p + NPk, cj);
        atomicAdd(p + 2 * NPk, 1u);
        atomicAdd(p + 3 * NPk, ci < cj ? ci : cj);
        if (SIMPLE) {
            const ull prod = (ull)ci * (ull)cj;
            add(4, prod);
            atomicAdd(p + 5 * NPk, (uint32_t)hellinger_term(prod));  // floor(sqrt(prod)) < 2^20; u64 += double truncates per increment
        }
    } else {
        add(0, ci);
        add(1, cj);
        atomicAdd(p + 2 * NPk, 1u);
        add(3, ci < cj ? ci : cj);
        if (SIMPLE) {
            const ull prod = (ull)ci * (ull)cj;
            add(4, prod);
            add(5, hellinger_term(prod));
        }
    }
    if (COMPLEX) {
        const double ai = (double)ci, aj = (double)cj, Ni = T.ntot[si], Nj = T.ntot[sj];
        const double yX = aj * Ni, xY = ai * Nj;
        if (xY != yX) {  // equal: both logs are log(1) = 0 exactly
            const double r1 = (2 * xY) * __drcp_rn(xY + yX);
            // d1 + d2 = (c_i/N_i) log r1 + (c_j/N_j) log r2 = (xY log r1 + yX log r2) / (N_i N_j): no division per pair, the reciprocals
            // are per sample.  The term is >= 0 (Jensen) and < 2: exact integer accumulation in units of 2^-60 (rounding 4e-19 per term,
            // order-independent => bit-reproducible sums)
            const double num = fma(xY, tab_log(r1, logtab), yX * tab_log(2.0 - r1, logtab));
            add2(8, (ull)__double2ll_rn(num * (T.ninv30[si] * T.ninv30[sj])));
        }
        // Whittaker: abs((int)(u64(xY) - u64(yX))) minus the two one-sided values, accumulated negated (<= 0 in practice).  Only the low 32
        // bits of the u64 values enter; with abundances < 2^20 and totals < 2^32 the products are exact integers below 2^52, so
        // their low words are one 32-bit multiply each instead of two double -> u64 conversions
        if (small && !T.big_tot) {
            const uint32_t wx = ci * T.nlo[sj], wy = cj * T.nlo[si];
            add2(6, (ull)(absint32_ext(wx) + absint32_ext(wy) - absint32_ext(wx - wy)));
        } else {
            const uint64_t wx = (uint64_t)xY, wy = (uint64_t)yX;
            add2(6, (ull)(absint32_ext(wx) + absint32_ext(wy) - absint32_ext(wx - wy)));
        }
    }
}

// TILED (N > 64): the sample pairs are cut into tiles of 32 x 32 samples (block I x block J, I <= J); a CTA works on
// one tile: it gathers only the <= 64 samples of its two blocks, groups them, and accumulates the pairs that belong
// to the tile (cross pairs for I != J, all pairs for I == J) into 1024 (496) shared-memory accumulators.  Grouping
// work is repeated once per tile a block takes part in (B+1 times for B blocks); the pair work is not.
template <bool SIMPLE, bool COMPLEX, bool TILED>
__global__ void __launch_bounds__(MS_THREADS, SIMKA_MS_MINB)
k_merge_sm(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MSSmem &S = *reinterpret_cast<MSSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = a.n;
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(MSSmem) + 15) & ~(size_t)15));
    const uint32_t MG = 1u << a.mg_bits;
    const uint32_t q = 1u << (a.fg_bits - a.mg_bits);
    const size_t ostride = ((size_t)1 << a.fg_bits) + 1;
    const uint32_t G = gridDim.x, c = blockIdx.x;
    const uint32_t B = (uint32_t)(N + 31) / 32, T = TILED ? B * (B + 1) / 2 : 1;
    const bool spread = G >= T;  // several CTAs per tile (each takes a stride of the merge bins) or several tiles per CTA

    if (tid < 128) S.logtab[tid] = a.logtab[tid];
    ull loc_groups = 0, loc_shared = 0, loc_pairs = 0;

    for (uint32_t tile = c % T; tile < T; tile += spread ? T : G) {
        const uint32_t bin_first = spread ? c / T : 0, bin_stride = spread ? (G - tile + T - 1) / T : 1;
        // tile geometry: row-major upper triangle (with diagonal) over the B sample blocks
        uint32_t I = 0, J = 0;
        if (TILED) {
            uint32_t rem = tile;
            while (rem >= B - I) { rem -= B - I; I++; }
            J = I + rem;
        }
        const bool diag = !TILED || I == J;
        const int NPk = TILED ? (diag ? 496 : 1024) : N * (N - 1) / 2;  // pair keys of this tile
        if (tid == 0) S.T.big_tot = 0;  // (the flush of the previous tile does not read it)
        __syncthreads();  // the flush of the previous tile is done
        for (int i = tid; i < MS_NACC * NPk; i += MS_THREADS) acc[i] = 0;
        // pair key -> local sample slots (li < lj) of the tile
        auto key_slots = [&](int key, int &li, int &lj) {
            if (TILED && !diag) { li = key >> 5; lj = 32 + (key & 31); }
            else {
                const int n_loc = TILED ? 32 : N;
                li = 0;
                while ((li + 1) * (2 * n_loc - (li + 1) - 1) / 2 <= key) li++;
                lj = li + 1 + (key - li * (2 * n_loc - li - 1) / 2);
            }
        };
        if (tid < MS_MAXN) {  // local sample slot -> global sample id (-1: unused)
            int gs = tid;
            if (TILED) gs = tid < 32 ? (int)(32 * I) + tid : (diag ? -1 : (int)(32 * J) + tid - 32);
            const bool used = gs >= 0 && gs < N;
            S.T.s_gid[tid] = used ? gs : -1;
            const double nt = used ? a.ntot[gs] : 0.0;
            S.T.ntot[tid] = nt;
            S.T.ninv30[tid] = nt > 0.0 ? 1073741824.0 / nt : 0.0;
            S.T.nlo[tid] = nt < 4294967296.0 ? (uint32_t)nt : 0u;
            if (nt >= 4294967296.0) S.T.big_tot = 1;
        }
        __syncthreads();

        for (uint32_t mb = bin_first; mb < MG; mb += bin_stride) {
            // segment descriptors of the (<= 64) samples of this tile
            if (tid < MS_MAXN) {
                uint32_t len = 0;
                const int gs = S.T.s_gid[tid];
                if (gs >= 0) {
                    const uint32_t b0 = a.off[(size_t)gs * ostride + (size_t)mb * q];
                    const uint32_t b1 = a.off[(size_t)gs * ostride + (size_t)(mb + 1) * q];
                    S.s_start[tid] = b0;
                    len = b1 - b0;
                }
                S.s_pre[tid] = len;  // lengths first, scanned below
            }
            __syncthreads();
            if (tid < 32) {
                const uint32_t l0 = S.s_pre[tid], l1 = S.s_pre[tid + 32];
                uint32_t inc0 = l0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc0, d); if (lane >= d) inc0 += u; }
                const uint32_t tot0 = __shfl_sync(0xffffffffu, inc0, 31);
                uint32_t inc1 = l1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc1, d); if (lane >= d) inc1 += u; }
                const uint32_t tot1 = __shfl_sync(0xffffffffu, inc1, 31);
                S.s_pre[tid] = inc0 - l0;
                S.s_pre[tid + 32] = tot0 + inc1 - l1;
                if (tid == 31) S.s_pre[64] = tot0 + tot1;
            }
            __syncthreads();
            const uint32_t n_e = S.s_pre[64];
            if (n_e == 0) { __syncthreads(); continue; }
            uint32_t R = 1;
            if (n_e > MS_ECAP) { while ((uint64_t)R * (MS_ECAP / 2) < n_e) R <<= 1; }
            for (uint32_t r = 0; r < R; r++) {
                {   // clean table
                    uint4 *pk = reinterpret_cast<uint4 *>(S.tk) + tid;
                    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), z = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int qq = 0; qq < MS_SLOTS * 8 / 16 / MS_THREADS; qq++) pk[qq * MS_THREADS] = e;
                    uint4 *pg = reinterpret_cast<uint4 *>(S.gcnt) + tid;
#pragma unroll
                    for (int qq = 0; qq < MS_SLOTS * 4 / 16 / MS_THREADS; qq++) pg[qq * MS_THREADS] = z;
                }
                __syncthreads();
                // phase 1: insert keys, count group sizes
                for (uint32_t i = tid; i < n_e; i += MS_THREADS) {
                    int lo = 0, hi = 64;  // s_pre[lo] <= i < s_pre[hi]   (empty slots repeat the next prefix)
#pragma unroll
                    for (int st = 0; st < 6; st++) { const int mid = (lo + hi) >> 1; if (S.s_pre[mid] <= i) lo = mid; else hi = mid; }
                    const uint64_t h = a.eh[S.T.s_gid[lo]][S.s_start[lo] + (i - S.s_pre[lo])];
                    const uint64_t hh = simka::hash2(h);
                    if (R > 1 && (((uint32_t)(hh >> 20)) & (R - 1)) != r) continue;
                    uint32_t slot = (uint32_t)(hh >> 40) & (MS_SLOTS - 1);
                    volatile uint64_t *vk = S.tk;
                    int probes = 0;
                    for (;;) {
                        uint64_t cur = vk[slot];
                        if (cur == EMPTY) {
                            cur = atomicCAS(reinterpret_cast<ull *>(&S.tk[slot]), (ull)EMPTY, (ull)h);
                            if (cur == EMPTY) cur = h;
                        }
                        if (cur == h) { atomicAdd(&S.gcnt[slot], 1u); break; }
                        slot = (slot + 1) & (MS_SLOTS - 1);
                        if (++probes >= MS_SLOTS) { *a.error_flag = 3; break; }
                    }
                }
                __syncthreads();
                // phase 2: group statistics + compact list of the shared groups (>= 2 samples)
                unsigned long long packed = 0;  // groups (16 bits) | entries (16 bits) | pairs (32 bits)
                uint32_t g8[MS_SLOTS / MS_THREADS];
#pragma unroll
                for (int qq = 0; qq < MS_SLOTS / MS_THREADS; qq++) {
                    const uint32_t g = S.gcnt[tid * (MS_SLOTS / MS_THREADS) + qq];
                    g8[qq] = g;
                    loc_groups += g > 0;
                    if (g > 1) packed += 1ull | ((unsigned long long)g << 16) | ((unsigned long long)(g * (g - 1) / 2) << 32);
                }
                unsigned long long total;
                unsigned long long ex = block_scan_u64(packed, S.scan_tmp, &total);
                const uint32_t n_groups = (uint32_t)(total & 0xFFFF), n_sub = (uint32_t)((total >> 16) & 0xFFFF);
                const uint32_t P = (uint32_t)(total >> 32);
                if (n_groups > MS_GCAP || n_sub > MS_ECAP) {  // cannot happen for hashed keys; reported, not silent
                    if (tid == 0) *a.error_flag = 4;
                    __syncthreads();
                    continue;
                }
#pragma unroll
                for (int qq = 0; qq < MS_SLOTS / MS_THREADS; qq++) {
                    const uint32_t g = g8[qq];
                    uint32_t v = MS_SKIP;
                    if (g > 1) {
                        const uint32_t gi = (uint32_t)(ex & 0xFFFF), st = (uint32_t)((ex >> 16) & 0xFFFF), pp = (uint32_t)(ex >> 32);
                        S.cg[gi] = st | (g << 16);
                        S.cgp[gi] = pp;
                        v = st + g;  // fill cursor counts down from the end
                        ex += 1ull | ((unsigned long long)g << 16) | ((unsigned long long)(g * (g - 1) / 2) << 32);
                        loc_shared++;
                    }
                    S.gcnt[tid * (MS_SLOTS / MS_THREADS) + qq] = v;
                }
                if (tid == 0) S.cgp[n_groups] = P;
                __syncthreads();
                if (n_groups == 0) continue;  // nothing shared in this pass: no accumulator changes
                // phase 3: scatter the entries of shared k-mers into their groups
                for (uint32_t i = tid; i < n_e; i += MS_THREADS) {
                    int lo = 0, hi = 64;
#pragma unroll
                    for (int st = 0; st < 6; st++) { const int mid = (lo + hi) >> 1; if (S.s_pre[mid] <= i) lo = mid; else hi = mid; }
                    const size_t idx = S.s_start[lo] + (i - S.s_pre[lo]);
                    const int gsid = S.T.s_gid[lo];
                    const uint64_t h = a.eh[gsid][idx];
                    const uint64_t hh = simka::hash2(h);
                    if (R > 1 && (((uint32_t)(hh >> 20)) & (R - 1)) != r) continue;
                    uint32_t slot = (uint32_t)(hh >> 40) & (MS_SLOTS - 1);
                    int pr = 0;
                    while (S.tk[slot] != h && ++pr <= MS_SLOTS) slot = (slot + 1) & (MS_SLOTS - 1);
                    if (pr > MS_SLOTS) { *a.error_flag = 7; continue; }  // inconsistent input (never for lists built by this library)
                    if (S.gcnt[slot] == MS_SKIP) continue;  // private k-mer: contributes to no pair
                    const uint32_t pos = atomicSub(&S.gcnt[slot], 1u) - 1;
                    S.ge_s[pos] = (uint8_t)lo;
                    S.ge_c[pos] = a.ec[gsid][idx];
                }
                __syncthreads();
                // phase 4: flat pair enumeration, one contiguous run of pairs per thread
                {
                    const uint32_t per = (P + MS_THREADS - 1) / MS_THREADS;
                    const uint32_t l0 = min(P, tid * per), l1 = min(P, l0 + per);
                    if (l0 < l1) {
                        int lo = 0, hi = (int)n_groups;  // cgp[lo] <= l0 < cgp[hi]
                        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.cgp[mid] <= l0) lo = mid; else hi = mid; }
                        uint32_t g = (uint32_t)lo;
                        uint32_t cgv = S.cg[g];
                        uint32_t st = cgv & 0xFFFF, sz = cgv >> 16;
                        const uint32_t l = l0 - S.cgp[g];
                        // row-major strict upper triangle of sz x sz: row x starts at x(2 sz - x - 1)/2
                        const float tg = 2.0f * (float)sz - 1.0f;
                        int x = (int)((tg - sqrtf(fmaxf(tg * tg - 8.0f * (float)l, 0.0f))) * 0.5f);
                        if (x > (int)sz - 2) x = (int)sz - 2;
                        if (x < 0) x = 0;
                        while ((uint32_t)(x + 1) * (2 * sz - (x + 1) - 1) / 2 <= l) x++;
                        while ((uint32_t)x * (2 * sz - x - 1) / 2 > l) x--;
                        uint32_t y = x + 1 + (l - (uint32_t)x * (2 * sz - x - 1) / 2);
                        uint32_t sx = S.ge_s[st + x], cx = S.ge_c[st + x];
                        for (uint32_t p = l0; p < l1; p++) {
                            const uint32_t sy = S.ge_s[st + y], cy = S.ge_c[st + y];
                            // pairs of two samples of the same block belong to the diagonal tile of that block
                            if (!TILED || diag || ((sx < 32) != (sy < 32))) {
                                const bool sw = sx > sy;
                                const uint32_t si = sw ? sy : sx, sj = sw ? sx : sy;  // local slots, si < sj
                                const uint32_t ci = sw ? cy : cx, cj = sw ? cx : cy;
                                uint32_t key;
                                if (!TILED) key = si * (2 * N - si - 1) / 2 + (sj - si - 1);
                                else if (diag) key = si * (63 - si) / 2 + (sj - si - 1);
                                else key = si * 32 + (sj - 32);
                                if (TILED) loc_pairs++;
                                ms_pair<SIMPLE, COMPLEX, 0>(a, acc, NPk, key, ci, cj, si, sj, S.T, S.logtab);
                            }
                            // advance
                            if (++y == sz) {
                                if (++x == (int)sz - 1) {
                                    g++;
                                    if (g < n_groups) { cgv = S.cg[g]; st = cgv & 0xFFFF; sz = cgv >> 16; }
                                    x = 0;
                                }
                                y = x + 1;
                                if (g < n_groups) { sx = S.ge_s[st + x]; cx = S.ge_c[st + x]; }
                            }
                        }
                    }
                    if (!TILED && tid == 0) loc_pairs += P;
                }
                __syncthreads();
                // sweep: a word that has reached 2^31 moves to the global accumulator (see MS_NACC)
                for (int i = tid; i < MS_NACC * NPk; i += MS_THREADS) {
                    const uint32_t v = acc[i];
                    if (v & 0x80000000u) {
                        int li, lj;
                        key_slots(i % NPk, li, lj);
                        ms_flush_word(a, i / NPk, S.T.s_gid[li], S.T.s_gid[lj], v);
                        acc[i] = 0;
                    }
                }
            }
        }
        __syncthreads();
        // flush: one set of global atomics per CTA, tile and pair key
        for (int key = tid; key < NPk; key += MS_THREADS) {
            int li = -1, lj = -1;  // local slots of the pair, decoded for the first non-zero word
            for (int w = 0; w < MS_NACC; w++) {
                const uint32_t v = acc[w * NPk + key];
                if (!v) continue;
                if (li < 0) key_slots(key, li, lj);
                ms_flush_word(a, w, S.T.s_gid[li], S.T.s_gid[lj], v);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        loc_groups += __shfl_down_sync(0xffffffffu, loc_groups, d);
        loc_shared += __shfl_down_sync(0xffffffffu, loc_shared, d);
        loc_pairs += __shfl_down_sync(0xffffffffu, loc_pairs, d);
    }
    if (lane == 0) {
        // a tiled run groups every k-mer once per tile: its distinct / shared k-mer counts come from the count-only pass
        if (!TILED && (loc_groups | loc_shared)) { atomicAdd(&a.scalars[0], loc_groups); atomicAdd(&a.scalars[1], loc_shared); }
        if (loc_pairs) atomicAdd(&a.scalars[2], loc_pairs);
    }
}

// T[i][j] = sum over solid k-mers e of sample i of f((u64)(c_e * N_j)), f = abs((int)x): the value the
// reference's Whittaker update adds for a k-mer present in i and absent from j
// (src/core/SimkaAlgorithm.hpp:481,512).  Computed from the abundance histogram of sample i.
__global__ void __launch_bounds__(BLOCK)
k_whittaker_onesided(const ull *__restrict__ chist /*[n][CHIST]*/, const uint32_t *__restrict__ big_counts,
                     const uint32_t *__restrict__ big_sample, uint32_t big_n, const double *__restrict__ ntot, int n,
                     ull *__restrict__ T /*[n][n]*/) {
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < n; j += BLOCK) {
        if (j == i) continue;
        const double Nj = ntot[j];
        ull acc = 0;
        for (int c = 1; c < CHIST; c++) {
            const ull m = chist[(size_t)i * CHIST + c];
            if (m) acc += m * (ull)absint32_ext((uint64_t)((double)c * Nj));
        }
        for (uint32_t q = 0; q < big_n; q++)
            if (big_sample[q] == (uint32_t)i) acc += (ull)absint32_ext((uint64_t)((double)big_counts[q] * Nj));
        T[(size_t)i * n + j] = acc;
    }
}

// whittaker[i][j] (i<j) += T[i][j] + T[j][i]
__global__ void k_whittaker_fold(ull *__restrict__ whit, const ull *__restrict__ T, int n) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    int i = idx / n, j = idx % n;
    if (i < j) whit[idx] += T[idx] + T[(size_t)j * n + i];
}

// ---------------------------------------------------------------------------------------------
// Refinement of the per-sample lists for the merge.  A list comes out of stage 1 grouped by fine range (top fg bits of h),
// and fg cannot exceed the counting bins of a sample; with many samples a merge bin of that width holds far more entries
// than the on-chip tables take and the merge would re-read it once per sub-pass (16-32 times at 128-512 samples).  These two
// kernels regroup one list by the top fg2 > fg bits (a counting sort inside every fine range; the order inside a group
// stays arbitrary, the merge regroups by key), so that merge bins can be as narrow as the total number of entries asks.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
k_refine_hist(const uint64_t *__restrict__ eh, uint64_t n, int shift, uint32_t *__restrict__ hist) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * BLOCK) atomicAdd(&hist[eh[i] >> shift], 1u);
}
__global__ void __launch_bounds__(BLOCK)
k_refine_scatter(const uint64_t *__restrict__ eh, const uint32_t *__restrict__ ec, uint64_t n, int shift, const uint64_t *__restrict__ base,
                 uint32_t *__restrict__ cursor, uint64_t *__restrict__ out_h, uint32_t *__restrict__ out_c) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * BLOCK) {
        const uint64_t h = eh[i];
        const uint64_t b = h >> shift;
        const uint64_t pos = base[b] + atomicAdd(&cursor[b], 1u);
        out_h[pos] = h;
        out_c[pos] = ec[i];
    }
}
