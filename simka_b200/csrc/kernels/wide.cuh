// kernels/wide.cuh -- stage 1 and stage 2 for 32 <= k <= 63: 128-bit k-mers.
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
//
// The reference compiles its k-mer type per span: Kmer<32> holds k <= 31 in one word, Kmer<64> holds k <= 63 in LargeInt<2>
// (GC/tools/math/LargeInt2.pri; dispatch Integer::apply, GC/tools/math/Integer.hpp:76-85; KSIZE_LIST, CMakeLists.txt:66-71).
// Same split here: the kernels of stage1.cuh / merge.cuh carry 64-bit keys and are tuned for them; this file is the second
// instantiation of the same pipeline on 128-bit keys --
//   k_wide_extract   canonical k-mer per start position (128-bit roll), invertible 2k-bit mix, fine bin = top bits of the mix:
//                    a count pass and a scatter pass (one global atomic per k-mer each)
//   k_wide_bin_count per-bin exact count in a shared-memory table of 16-byte keys claimed with ATOMS.CAS.128
//   k_wide_compact   survivors -> the sample's list (16-byte keys + counts), grouped by fine range like the 64-bit lists
//   k_wide_merge     the generic merge (k_merge) on 16-byte keys: grouping table claimed with ATOMS.CAS.128, same pair enumeration
//                    and the same accumulate_pair
// -- simpler than the 64-bit path (no tile-sorted two-level partition, no shared-memory accumulators): the 128-bit sizes are
// outside every BASELINE configuration, they have to be exact, not fast.
#pragma once

struct __align__(16) K128 { uint64_t lo, hi; };
typedef unsigned __int128 u128;

__device__ __forceinline__ u128 k128_to_u(const K128 &k) { return ((u128)k.hi << 64) | k.lo; }
__device__ __forceinline__ K128 u_to_k128(u128 v) { K128 k; k.lo = (uint64_t)v; k.hi = (uint64_t)(v >> 64); return k; }
__device__ __forceinline__ bool k128_eq(const K128 &a, const K128 &b) { return a.lo == b.lo && a.hi == b.hi; }
__device__ __forceinline__ bool k128_empty(const K128 &a) { return (a.lo & a.hi) == ~0ull; }  // all ones: never a mixed k-mer (2k <= 126 bits)

// compare-and-swap of a 16-byte slot in shared memory
__device__ __forceinline__ K128 cas128_shared(K128 *addr, K128 cmp, K128 val) {
#if defined(__CUDACC__)
    K128 old;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(addr);
    asm volatile("{\n.reg .b128 c, v, o;\n"
                 "mov.b128 c, {%3, %4};\n"
                 "mov.b128 v, {%5, %6};\n"
                 "atom.shared.cas.b128 o, [%2], c, v;\n"
                 "mov.b128 {%0, %1}, o;\n}\n"
                 : "=l"(old.lo), "=l"(old.hi)
                 : "r"(a), "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi)
                 : "memory");
    return old;
#else  // kernel-logic emulator: fibers switch only at barriers
    const K128 old = *addr;
    if (k128_eq(old, cmp)) *addr = val;
    return old;
#endif
}

// invertible mix of a 2k-bit value (the 128-bit form of simka::mix_fwd): xorshift, odd multiply modulo 2^(2k), xorshift
struct Mix128 {
    uint64_t m_lo, m_hi;   // odd multiplier
    uint64_t i_lo, i_hi;   // its inverse modulo 2^128
    int shift, bits;       // k, 2k
};
__host__ __device__ __forceinline__ u128 mix128_mask(const Mix128 &p) { return p.bits >= 128 ? ~(u128)0 : ((((u128)1) << p.bits) - 1); }
__host__ __device__ __forceinline__ u128 mix128_fwd(u128 x, const Mix128 &p) {
    x ^= x >> p.shift;
    x = (x * ((((u128)p.m_hi) << 64) | p.m_lo)) & mix128_mask(p);
    x ^= x >> p.shift;
    return x;
}
__host__ __device__ __forceinline__ u128 mix128_inv(u128 x, const Mix128 &p) {
    x ^= x >> p.shift;
    x = (x * ((((u128)p.i_hi) << 64) | p.i_lo)) & mix128_mask(p);
    x ^= x >> p.shift;
    return x;
}
static Mix128 make_mix128(int k) {
    Mix128 p;
    p.bits = 2 * k;
    p.shift = k;  // ceil(2k / 2): x ^= x >> k is an involution on 2k bits
    const u128 m = ((((u128)0x9E3779B97F4A7C15ull) << 64) | 0xFF51AFD7ED558CCDull);
    u128 inv = m;  // Newton: inverse of an odd number modulo 2^128
    for (int i = 0; i < 7; i++) inv *= 2 - m * inv;
    p.m_lo = (uint64_t)m; p.m_hi = (uint64_t)(m >> 64);
    p.i_lo = (uint64_t)inv; p.i_hi = (uint64_t)(inv >> 64);
    return p;
}
__device__ __forceinline__ uint64_t wide_hash(const K128 &h) { return (h.lo ^ (h.hi * 0xC2B2AE3D27D4EB4Full)) * 0x9E3779B97F4A7C15ull; }

// ---------------------------------------------------------------------------------------------
// K1 (wide): thread = one packed word = 32 start positions; the k-mer window spans up to three packed words.
// Replaces Kmer<64>::ModelCanonical::iterate (GC/kmer/impl/Model.hpp:726-765,858-884 on LargeInt<2>).
// ---------------------------------------------------------------------------------------------
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(BLOCK)
k_wide_extract(const uint64_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t n_words, int k, Mix128 mix, int bin_bits,
               uint32_t *__restrict__ hist, const uint64_t *__restrict__ bin_off, uint32_t *__restrict__ cursor, K128 *__restrict__ keys_out) {
    const uint64_t t = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (t >= n_words) return;
    // words t, t+1, t+2 of the stream (the arrays hold n_words + 1 entries; beyond that everything is bad)
    uint64_t w[3];
    uint32_t b[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const bool in = t + q <= n_words;
        w[q] = in ? seq[t + q] : 0;
        b[q] = in ? bad[t + q] : 0xFFFFFFFFu;
    }
    // window j is valid iff none of its k bases is bad: OR of the k left shifts of the 96-bit bad mask (by doubling)
    uint32_t okmask;
    {
        u128 r = (((u128)b[0]) << 96) | (((u128)b[1]) << 64) | (((u128)b[2]) << 32) | 0xFFFFFFFFull;
        int built = 1;
        while (built * 2 <= k) { r |= r << built; built *= 2; }
        if (built < k) r |= r << (k - built);
        okmask = ~(uint32_t)(r >> 96);  // bit 31 - j <-> start position j
    }
    if (okmask == 0) return;
    const u128 kmask = mix128_mask(mix);
    u128 fwd = 0, rev = 0;
    const int last = 31 - (__ffs((int)okmask) - 1);  // last valid start position of this word
    for (int i = 0; i < k + last; i++) {
        const uint32_t c = (uint32_t)(w[i >> 5] >> (62 - 2 * (i & 31))) & 3u;
        fwd = ((fwd << 2) | c) & kmask;
        rev = (rev >> 2) | ((u128)(c ^ 2u) << (2 * (k - 1)));
        const int j = i - (k - 1);  // start position of the window that ends at base i
        if (j >= 0 && ((okmask >> (31 - j)) & 1u)) {
            const u128 h = mix128_fwd(fwd < rev ? fwd : rev, mix);
            const uint32_t bin = bin_bits ? (uint32_t)(h >> (mix.bits - bin_bits)) : 0u;
            if (COUNT_ONLY) atomicAdd(&hist[bin], 1u);
            else keys_out[bin_off[bin] + atomicAdd(&cursor[bin], 1u)] = u_to_k128(h);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3 (wide): per-bin exact multiset count + abundance filter + per-sample totals, as k_bin_count
// (GC/kmer/impl/PartitionsCommand.cpp:1390-1504,1600-1760 + src/minikc/MiniKC.hpp:54-79).
// ---------------------------------------------------------------------------------------------
struct WideOut {
    K128 *stage_h;       // survivors of bin b at fine_off[b] + i
    uint32_t *stage_c;
    uint32_t *surv;
    ull *totals;
    int *error_flag;
};
constexpr int WD_SLOTS = 2048;
constexpr int WD_MAXD = WD_SLOTS * 3 / 4;

__global__ void __launch_bounds__(BLOCK)
k_wide_bin_count(const K128 *__restrict__ keys, const uint64_t *__restrict__ fine_off, uint32_t nbins, int64_t amin, int64_t amax, WideOut out) {
    __shared__ K128 tk[WD_SLOTS];
    __shared__ uint32_t tc[WD_SLOTS];
    __shared__ uint16_t claimed[WD_SLOTS];
    __shared__ uint32_t s_cursor, s_nclaim;
    __shared__ ull s_tot[3];
    const int tid = threadIdx.x, lane = tid & 31;
    K128 empty; empty.lo = empty.hi = ~0ull;
    for (int i = tid; i < WD_SLOTS; i += BLOCK) { tk[i] = empty; tc[i] = 0; }
    if (tid == 0) { s_cursor = 0; s_nclaim = 0; s_tot[0] = s_tot[1] = s_tot[2] = 0; }
    __syncthreads();
    ull t_d = 0, t_s = 0, t_q = 0;
    const bool never = amax < 0 || amin > 0xFFFFFFFFll;
    const uint32_t cmin = amin < 0 ? 0u : (uint32_t)amin;
    const uint32_t cmax = amax > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)(amax < 0 ? 0 : amax);
    for (uint32_t bin = blockIdx.x; bin < nbins; bin += gridDim.x) {
        const uint64_t beg = fine_off[bin], end = fine_off[bin + 1];
        if (end == beg) { if (tid == 0) out.surv[bin] = 0; continue; }
        if (end - beg >= 0xFFFFFFFFull) { if (tid == 0) { *out.error_flag = 6; out.surv[bin] = 0; } continue; }
        const uint32_t n = (uint32_t)(end - beg);
        // R sub-passes (keys selected by a second hash) when the distinct keys of the bin do not fit the table
        for (uint32_t R = 1;; R *= 2) {
            if (R > 65536) { if (tid == 0) *out.error_flag = 5; break; }
            ull b_d = 0, b_s = 0, b_q = 0;
            bool overflow = false;
            for (uint32_t r = 0; r < R && !overflow; r++) {
                for (uint32_t i = tid; i < n; i += BLOCK) {
                    if (*((volatile uint32_t *)&s_nclaim) > (uint32_t)WD_MAXD) break;
                    const K128 h = keys[beg + i];
                    const uint64_t hh = wide_hash(h);
                    if (R > 1 && ((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                    uint32_t slot = (uint32_t)(hh >> 44) & (WD_SLOTS - 1);
                    for (int probes = 0; probes < WD_SLOTS; probes++) {
                        const K128 old = cas128_shared(&tk[slot], empty, h);
                        const bool won = k128_empty(old);
                        if (won) { const uint32_t q = atomicAdd(&s_nclaim, 1u); if (q < (uint32_t)WD_SLOTS) claimed[q] = (uint16_t)slot; }
                        if (won || k128_eq(old, h)) { atomicAdd(&tc[slot], 1u); break; }
                        slot = (slot + 1) & (WD_SLOTS - 1);
                    }
                }
                __syncthreads();
                const uint32_t nc = min(s_nclaim, (uint32_t)WD_SLOTS);
                overflow = s_nclaim > (uint32_t)WD_MAXD;
                for (uint32_t e0 = 0; e0 < nc; e0 += BLOCK) {  // read-out + clean of the claimed slots
                    const uint32_t e = e0 + tid;
                    bool keep = false;
                    K128 hk = empty;
                    uint32_t c = 0;
                    if (e < nc) {
                        const uint32_t slot = claimed[e];
                        hk = tk[slot]; c = tc[slot];
                        tk[slot] = empty; tc[slot] = 0;
                        keep = !overflow && !never && c >= cmin && c <= cmax;
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, keep);
                    if (m) {
                        uint32_t wbase = 0;
                        if (lane == 0) wbase = atomicAdd(&s_cursor, (uint32_t)__popc(m));
                        wbase = __shfl_sync(0xffffffffu, wbase, 0);
                        if (keep) {
                            const uint64_t pos = beg + wbase + __popc(m & ((1u << lane) - 1));
                            out.stage_h[pos] = hk;
                            out.stage_c[pos] = c;
                            b_d += 1; b_s += c;
                            b_q += (ull)((double)c * (double)c);  // pow(count,2) -> u64, src/minikc/MiniKC.hpp:62
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) s_nclaim = 0;
                __syncthreads();
            }
            if (!overflow) { t_d += b_d; t_s += b_s; t_q += b_q; break; }
            if (tid == 0) s_cursor = 0;  // forget the aborted attempt (its survivors get overwritten)
            __syncthreads();
        }
        if (tid == 0) { out.surv[bin] = s_cursor; s_cursor = 0; }
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        t_d += __shfl_down_sync(0xffffffffu, t_d, d);
        t_s += __shfl_down_sync(0xffffffffu, t_s, d);
        t_q += __shfl_down_sync(0xffffffffu, t_q, d);
    }
    if (lane == 0 && t_d) { atomicAdd(&s_tot[0], t_d); atomicAdd(&s_tot[1], t_s); atomicAdd(&s_tot[2], t_q); }
    __syncthreads();
    if (tid == 0 && s_tot[0]) { atomicAdd(&out.totals[0], s_tot[0]); atomicAdd(&out.totals[1], s_tot[1]); atomicAdd(&out.totals[2], s_tot[2]); }
}

// survivors staging -> exactly-sized per-sample list; one warp per bin
__global__ void __launch_bounds__(BLOCK)
k_wide_compact(const K128 *__restrict__ stage_h, const uint32_t *__restrict__ stage_c, const uint64_t *__restrict__ fine_off,
               const uint64_t *__restrict__ out_off, uint32_t nbins, K128 *__restrict__ eh, uint32_t *__restrict__ ec) {
    const int lane = threadIdx.x & 31;
    const uint32_t nw = gridDim.x * (BLOCK / 32);
    for (uint32_t bin = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); bin < nbins; bin += nw) {
        const uint64_t src = fine_off[bin], dst = out_off[bin];
        const uint32_t n = (uint32_t)(out_off[bin + 1] - dst);
        for (uint32_t i = lane; i < n; i += 32) { eh[dst + i] = stage_h[src + i]; ec[dst + i] = stage_c[src + i]; }
    }
}

// mixed key -> canonical k-mer as (lo, hi) words (parity hook)
__global__ void k_wide_unmix(const K128 *__restrict__ eh, uint64_t *__restrict__ lo, uint64_t *__restrict__ hi, uint64_t n, Mix128 mix) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u128 x = mix128_inv(k128_to_u(eh[i]), mix);
    lo[i] = (uint64_t)x; hi[i] = (uint64_t)(x >> 64);
}

// ---------------------------------------------------------------------------------------------
// K4 (wide): k_merge on 16-byte keys (src/SimkaMerge.cpp:1164-1264,1307-1326; src/core/SimkaAlgorithm.hpp:341-516).
// ---------------------------------------------------------------------------------------------
struct WideMergeArgs {
    MergeArgs m;                 // everything but the key arrays (m.eh is unused)
    const K128 *const *ehw;      // [n] per-sample mixed k-mers
};
static size_t wide_merge_smem_bytes(int n) {
    return (size_t)MRG_SLOTS * 16 + (size_t)MRG_SLOTS * 4 + ((size_t)MRG_SLOTS + 1) * 4 + (size_t)MRG_CAP * 4 + ((size_t)2 * n + 1) * 4 +
           (size_t)MRG_SLOTS * 2 + (size_t)MRG_CAP * 2 + 16;
}

__global__ void __launch_bounds__(BLOCK)
k_wide_merge(WideMergeArgs wa) {
    const MergeArgs &a = wa.m;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K128 *tk = reinterpret_cast<K128 *>(smem_raw);              // [MRG_SLOTS] table keys
    uint32_t *gs = reinterpret_cast<uint32_t *>(tk + MRG_SLOTS); // [MRG_SLOTS] group size, then group start
    uint32_t *pp = gs + MRG_SLOTS;                               // [MRG_SLOTS+1] fill cursors, then pair prefix
    uint32_t *ge_c = pp + MRG_SLOTS + 1;                         // [MRG_CAP] grouped counts
    uint32_t *s_start = ge_c + MRG_CAP;                          // [N]
    uint32_t *s_pre = s_start + a.n;                             // [N+1]
    uint16_t *gsz = reinterpret_cast<uint16_t *>(s_pre + a.n + 1); // [MRG_SLOTS]
    uint16_t *ge_s = gsz + MRG_SLOTS;                            // [MRG_CAP]
    __shared__ uint32_t warp_tmp[16];
    __shared__ uint32_t s_groups, s_shared;
    const int tid = threadIdx.x;
    const int N = a.n;
    const uint32_t MG = 1u << a.mg_bits;
    const uint32_t q = 1u << (a.fg_bits - a.mg_bits);
    const size_t ostride = ((size_t)1 << a.fg_bits) + 1;
    ull loc_groups = 0, loc_shared = 0;
    K128 empty; empty.lo = empty.hi = ~0ull;

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
        uint32_t R = R0, r = 0;  // sub-passes with splitting, as in k_merge
        bool more = true;
        while (more) {
            bool split = false;
            __syncthreads();
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) { tk[i] = empty; gs[i] = 0; pp[i] = 0; }
            if (tid == 0) { s_groups = 0; s_shared = 0; }
            __syncthreads();
            for (uint32_t i = tid; i < n_e; i += BLOCK) {  // phase 1: insert keys, count group sizes
                int lo = 0, hi = N;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                const K128 h = wa.ehw[lo][s_start[lo] + (i - s_pre[lo])];
                const uint64_t hh = wide_hash(h);
                if (((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 44) & (MRG_SLOTS - 1);
                for (int probes = 0; probes < MRG_SLOTS; probes++) {
                    const K128 old = cas128_shared(&tk[slot], empty, h);
                    if (k128_empty(old) || k128_eq(old, h)) { atomicAdd(&gs[slot], 1u); break; }
                    slot = (slot + 1) & (MRG_SLOTS - 1);
                }
            }
            __syncthreads();
            uint32_t my_groups = 0, my_shared = 0;  // phase 2: group statistics; only groups with >= 2 samples are materialised
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
            const uint32_t n_shared_groups = s_shared, n_groups = s_groups;
            __syncthreads();
#define WMRG_NEXT()                                                                                      \
            do {                                                                                         \
                if (split) { R <<= 1; }                                                                  \
                else {                                                                                   \
                    while (R > R0 && (r & (R >> 1))) { r -= R >> 1; R >>= 1; }                           \
                    if (R == R0) { r++; if (r >= R0) more = false; }                                     \
                    else r += R >> 1;                                                                    \
                }                                                                                        \
            } while (0)
            if (n_sub > MRG_CAP || n_groups > (uint32_t)(MRG_SLOTS * 7 / 8)) {
                if (R >= (1u << 24)) { if (tid == 0) *a.error_flag = 4; }
                else { split = true; WMRG_NEXT(); continue; }
            }
            if (tid == 0) { loc_groups += n_groups; loc_shared += n_shared_groups; }
            if (a.count_only || n_sub > MRG_CAP || n_shared_groups == 0) { WMRG_NEXT(); continue; }
            for (uint32_t i = tid; i < n_e; i += BLOCK) {  // phase 3: scatter the entries of shared k-mers into their groups
                int lo = 0, hi = N;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= i) lo = mid; else hi = mid; }
                const size_t idx = s_start[lo] + (i - s_pre[lo]);
                const K128 h = wa.ehw[lo][idx];
                const uint64_t hh = wide_hash(h);
                if (((uint32_t)(hh >> 20) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 44) & (MRG_SLOTS - 1);
                int pr = 0;
                while (!k128_eq(tk[slot], h) && ++pr <= MRG_SLOTS) slot = (slot + 1) & (MRG_SLOTS - 1);
                if (pr > MRG_SLOTS) { *a.error_flag = 7; continue; }
                if (gsz[slot] < 2) continue;
                const uint32_t pos = gs[slot] + atomicAdd(&pp[slot], 1u);
                ge_s[pos] = (uint16_t)lo;
                ge_c[pos] = a.ec[lo][idx];
            }
            __syncthreads();
            for (int i = tid; i < MRG_SLOTS; i += BLOCK) { const uint32_t g = gsz[i]; pp[i] = g * (g - 1) / 2; }
            __syncthreads();
            block_exclusive_scan(pp, MRG_SLOTS, warp_tmp);
            const uint32_t P = warp_tmp[BLOCK / 32];
            if (tid == 0) pp[MRG_SLOTS] = P;
            __syncthreads();
            for (uint32_t p = tid; p < P; p += BLOCK) {  // phase 4: flat pair enumeration
                int lo = 0, hi = MRG_SLOTS;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pp[mid] <= p) lo = mid; else hi = mid; }
                const uint32_t g = gsz[lo], l = p - pp[lo], base = gs[lo];
                const float tg = 2.0f * (float)g - 1.0f;
                int x = (int)((tg - sqrtf(fmaxf(tg * tg - 8.0f * (float)l, 0.0f))) * 0.5f);
                if (x > (int)g - 2) x = (int)g - 2;
                if (x < 0) x = 0;
                while ((uint32_t)(x + 1) * (2 * g - (x + 1) - 1) / 2 <= l) x++;
                while ((uint32_t)x * (2 * g - x - 1) / 2 > l) x--;
                const uint32_t y = x + 1 + (l - (uint32_t)x * (2 * g - x - 1) / 2);
                int si = ge_s[base + x], sj = ge_s[base + y];
                uint32_t ci = ge_c[base + x], cj = ge_c[base + y];
                if (si > sj) { int ts = si; si = sj; sj = ts; uint32_t tc2 = ci; ci = cj; cj = tc2; }
                accumulate_pair(a, si, sj, ci, cj);
            }
            if (tid == 0) atomicAdd(&a.scalars[2], (ull)P);
            WMRG_NEXT();
        }
#undef WMRG_NEXT
    }
    if (tid == 0 && (loc_groups | loc_shared)) {
        atomicAdd(&a.scalars[0], loc_groups);
        atomicAdd(&a.scalars[1], loc_shared);
    }
}
