// simka_gpu.cu -- sm_100a kernels and the C ABI of the Simka hot path (see include/simka_b200.h).
//
// Pipeline (one context per GPU, everything on one CUDA stream + one copy stream):
//
//   per sample  push_reads      H2D of the 2-bit packed base stream (+ 1 bit/base bad mask)
//               K1 extract      canonical k-mer per start position -> invertible hash h -> L1 bucket
//                               (count pass, scan, scatter pass through a shared-memory sorted tile)
//               K2 partition    L1 bucket -> fine bins (count, scan, scatter), bins ~1K keys
//               K3 bin count    one CTA per bin: shared-memory hash table, exact multiset count,
//                               abundance filter, per-sample totals, survivors compacted
//   all samples merge           one CTA per merge bin: gather the N per-sample segments, group equal
//                               k-mers in a shared-memory table (= the N-wide abundance vector, stored
//                               sparse), enumerate sample pairs, update the N x N accumulators
//
// What each stage replaces in the reference is cited at the kernels.  No host fallback exists.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/simka_b200.h"
#include "kmer_device.cuh"

using simka::Mix;

typedef unsigned long long ull;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ctx->set_error(std::string(#call) + ": " + cudaGetErrorString(e_));            \
            return SIMKA_ERR_CUDA;                                                        \
        }                                                                                 \
    } while (0)

namespace {

constexpr int BLOCK = 256;
constexpr int L1_KPT = 32;                   // K1: one packed word (32 start positions) per thread
constexpr int L1_TILE = BLOCK * L1_KPT;      // 8192 keys
constexpr int L2_KPT = 16;
constexpr int L2_TILE = BLOCK * L2_KPT;      // 4096 keys
constexpr int MAX_B = 1024;                  // max buckets per partition level
constexpr int BIN_TARGET = 768;              // fine bins hold BIN_TARGET..2*BIN_TARGET keys on average
constexpr int CNT_SLOTS = 2048;              // K3 hash table slots
constexpr int CNT_MAXDISTINCT = 1536;        // distinct keys one table pass may hold (load 0.75)
constexpr int CHIST = 1024;                  // dense abundance histogram bins (complex distances)
constexpr int MRG_SLOTS = 2048;              // merge hash table slots
constexpr int MRG_CAP = 2048;                // merge: grouped entries held in shared memory
constexpr int MRG_TARGET = 768;              // average entries per merge bin
constexpr int MRG_MAXN = 8192;               // samples supported by the shared-memory prefix arrays
constexpr uint64_t EMPTY = ~0ull;

struct KParams {
    Mix mix;
    int k;
    int shard_rank, shard_count;
};

__device__ __forceinline__ bool shard_keep(uint64_t h, const KParams &kp) {
    if (kp.shard_count <= 1) return true;
    uint32_t x = (uint32_t)(simka::hash2(h) >> 32);
    return (int)(((uint64_t)x * (uint32_t)kp.shard_count) >> 32) == kp.shard_rank;
}

// exclusive scan of a shared array of n (<= MAX_B*?) u32 values by the whole block, in place.
// Returns the total in *total (shared).  n need not be a multiple of the block size.
template <int NT = BLOCK>
__device__ void block_exclusive_scan(uint32_t *a, int n, uint32_t *warp_tmp /*[NT/32+1]*/) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + NT - 1) / NT;
    const int beg = min(tid * per, n), end = min(beg + per, n);
    uint32_t sum = 0;
    for (int i = beg; i < end; i++) sum += a[i];
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) warp_tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < NT / 32 ? warp_tmp[lane] : 0, winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += v;
        }
        if (lane < NT / 32) warp_tmp[lane] = winc - w;
        if (lane == NT / 32 - 1) warp_tmp[NT / 32] = winc;
    }
    __syncthreads();
    uint32_t run = warp_tmp[wid] + inc - sum;
    for (int i = beg; i < end; i++) {
        uint32_t v = a[i];
        a[i] = run;
        run += v;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// P0: ASCII -> 2-bit packed stream + bad mask (GPU-side packing for simka_gpu_push_ascii).
// Stream position p = read_off[r] + r + o: one separator (bad) position after every read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_pack_ascii(const char *__restrict__ seqs, const uint64_t *__restrict__ read_off,
                                                      uint64_t n_reads, uint64_t n_words, uint64_t *__restrict__ packed,
                                                      uint32_t *__restrict__ bad) {
    uint64_t t = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (t > n_words) return;
    if (t == n_words) { packed[t] = 0; bad[t] = 0xFFFFFFFFu; return; }
    const uint64_t p0 = t * 32;
    // largest r with read_off[r] + r <= p0
    uint64_t lo = 0, hi = n_reads;  // invariant: start(lo) <= p0, start(hi) > p0 (start(n_reads) = total)
    if (n_reads == 0 || read_off[n_reads] + n_reads <= p0) {
        packed[t] = 0; bad[t] = 0xFFFFFFFFu; return;
    }
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (read_off[mid] + mid <= p0) lo = mid; else hi = mid;
    }
    uint64_t r = lo;
    uint64_t rbeg = read_off[r], rend = read_off[r + 1];
    uint64_t o = p0 - (rbeg + r);  // offset inside read r (== len means separator)
    uint64_t w = 0;
    uint32_t b = 0;
#pragma unroll 4
    for (int j = 0; j < 32; j++) {
        uint32_t code = 0, isbad = 1;
        if (r < n_reads) {
            if (rbeg + o < rend) {
                unsigned char c = (unsigned char)seqs[rbeg + o];
                code = (c >> 1) & 3u;
                unsigned char u = c & 0xDF;  // upper-case
                isbad = !(u == 'A' || u == 'C' || u == 'G' || u == 'T');
                o++;
            } else {  // separator
                r++;
                if (r < n_reads) { rbeg = rend; rend = read_off[r + 1]; }
                o = 0;
            }
        }
        w = (w << 2) | code;
        b = (b << 1) | isbad;
    }
    packed[t] = w;
    bad[t] = b;
}

// ---------------------------------------------------------------------------------------------
// K1: canonical k-mer extraction + level-1 bucket partition.
// Replaces Kmer<span>::ModelCanonical::iterate (GC/kmer/impl/Model.hpp:726-765,858-884) and the
// minimizer/partition step of FillPartitions (GC/kmer/impl/SortingCountAlgorithm.cpp:1048-1140).
// Thread t owns packed word t = start positions 32t..32t+31; the 62-bit window slides through the
// (word t, word t+1) register pair, the reverse strand is rolled alongside.
// ---------------------------------------------------------------------------------------------
template <bool COUNT_ONLY, bool SHARDED, bool BIGK>
__global__ void __launch_bounds__(BLOCK, 2)
k_extract(const uint64_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t n_words, KParams kp, int b1_bits,
          uint32_t *__restrict__ g_hist, const uint64_t *__restrict__ bucket_base, uint32_t *__restrict__ bucket_cursor,
          uint64_t bucket_cap, int *__restrict__ overflow_flag, uint64_t *__restrict__ keys_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *dst = reinterpret_cast<uint64_t *>(smem_raw);            // [MAX_B] global destination of a bucket's run minus its tile-local start
    uint32_t *hist = reinterpret_cast<uint32_t *>(dst + MAX_B);        // [MAX_B] counts, then tile-local starts
    uint32_t *warp_tmp = hist + MAX_B;                                 // [16]
    uint64_t *sorted = reinterpret_cast<uint64_t *>(warp_tmp + 16);    // [L1_TILE] (scatter mode only)

    const int tid = threadIdx.x;
    const int nb = 1 << b1_bits;
    const int W = kp.mix.bits;
    const int k = kp.k;
    for (int i = tid; i < nb; i += BLOCK) hist[i] = 0;
    __syncthreads();

    const uint64_t t = (uint64_t)blockIdx.x * BLOCK + tid;
    uint64_t hi = 0, lo = 0, bm = ~0ull;
    if (t < n_words) {
        hi = seq[t];
        lo = seq[t + 1];
        bm = ((uint64_t)bad[t] << 32) | bad[t + 1];
    }
    // window j (bases j..j+k-1 of the 64 in (hi,lo)) is valid iff none of its bases is bad: OR of the k left
    // shifts of the bad mask, built by doubling; bit 63-j of the result belongs to start position j
    uint32_t okmask;
    {
        uint64_t r = bm;
        int built = 1;
        while (built * 2 <= k) { r |= r << built; built *= 2; }
        if (built < k) r |= r << (k - built);
        okmask = ~(uint32_t)(r >> 32);
    }
    uint64_t keys[L1_KPT];
    uint32_t ranks[L1_KPT / 2];
    uint32_t vmask = 0;
    const int bshift = W - b1_bits;
    if (BIGK) {
        // 17 <= k <= 31: the k-mer straddles both 32-bit halves and every run-time shift amount is < 32, so the
        // whole roll is written on explicit halves with funnel shifts (a 64-bit shift by a run-time amount costs
        // the compiler 4-6 instructions, a funnel shift one).
        const uint32_t A = (uint32_t)(hi >> 32), B = (uint32_t)hi, Cw = (uint32_t)(lo >> 32), D = (uint32_t)lo;
        const int s1 = 64 - 2 * k;        // fwd = top 2k bits of the 64-bit window: 2..30
        const int s2 = 2 * k - 34;        // position of the incoming reverse-strand base inside the high half: 0..28
        const int sx = kp.mix.shift;      // k
        const uint32_t mhi = (uint32_t)(kp.mix.mask >> 32);
        const uint32_t m1l = (uint32_t)kp.mix.m1, m1h = (uint32_t)(kp.mix.m1 >> 32);
        const int bs = bshift - 32;       // bucket = top bits of the high half (b1_bits <= 2k - 32)
        uint32_t rl = 0, rh = 0;
#pragma unroll
        for (int j = 0; j < L1_KPT; j++) {
            // 64-bit window starting at base j
            uint32_t w1, w0;
            if (j == 0) { w1 = A; w0 = B; }
            else if (j < 16) { w1 = __funnelshift_l(B, A, 2 * j); w0 = __funnelshift_l(Cw, B, 2 * j); }
            else if (j == 16) { w1 = B; w0 = Cw; }
            else { w1 = __funnelshift_l(Cw, B, 2 * j - 32); w0 = __funnelshift_l(D, Cw, 2 * j - 32); }
            const uint32_t fh = w1 >> s1, fl = __funnelshift_r(w0, w1, s1);
            if (j == 0) {
                const uint64_t r0 = simka::revcomp(((uint64_t)fh << 32) | fl, k);
                rh = (uint32_t)(r0 >> 32); rl = (uint32_t)r0;
            } else {
                rl = __funnelshift_r(rl, rh, 2);
                rh = (rh >> 2) | (((fl & 3u) ^ 2u) << s2);
            }
            const bool fwd_smaller = fh < rh || (fh == rh && fl < rl);
            uint32_t xh = fwd_smaller ? fh : rh, xl = fwd_smaller ? fl : rl;
            // mix: x ^= x >> k; x = x * m1 mod 2^2k; x ^= x >> k
            xl ^= __funnelshift_r(xl, xh, sx); xh ^= xh >> sx;
            {
                const uint64_t p = (uint64_t)xl * m1l;
                const uint32_t ph = (uint32_t)(p >> 32) + xl * m1h + xh * m1l;
                xl = (uint32_t)p; xh = ph & mhi;
            }
            xl ^= __funnelshift_r(xl, xh, sx); xh ^= xh >> sx;
            const uint64_t h = ((uint64_t)xh << 32) | xl;
            bool keep = (okmask >> (31 - j)) & 1u;
            if (SHARDED) keep = keep && shard_keep(h, kp);
            keys[j] = h;
            if (keep) {
                vmask |= 1u << j;
                uint32_t r = atomicAdd(&hist[xh >> bs], 1u);
                if (!COUNT_ONLY) {
                    if (j & 1) ranks[j >> 1] |= r << 16; else ranks[j >> 1] = r;
                }
            } else if (!COUNT_ONLY && !(j & 1)) {
                ranks[j >> 1] = 0;
            }
        }
    } else {
        uint64_t rev = 0;
#pragma unroll
        for (int j = 0; j < L1_KPT; j++) {
            const uint64_t fwd = hi >> (64 - 2 * k);
            if (j == 0) rev = simka::revcomp(fwd, k);
            else rev = (rev >> 2) | ((uint64_t)((fwd & 3) ^ 2) << (2 * k - 2));
            const uint64_t canon = fwd < rev ? fwd : rev;
            const uint64_t h = simka::mix_fwd(canon, kp.mix);
            bool keep = (okmask >> (31 - j)) & 1u;
            if (SHARDED) keep = keep && shard_keep(h, kp);
            keys[j] = h;
            if (keep) {
                vmask |= 1u << j;
                uint32_t r = atomicAdd(&hist[(uint32_t)(h >> bshift)], 1u);
                if (!COUNT_ONLY) {
                    if (j & 1) ranks[j >> 1] |= r << 16; else ranks[j >> 1] = r;
                }
            } else if (!COUNT_ONLY && !(j & 1)) {
                ranks[j >> 1] = 0;
            }
            hi = (hi << 2) | (lo >> 62);
            lo <<= 2;
        }
    }
    __syncthreads();
    if (COUNT_ONLY) {
        for (int i = tid; i < nb; i += BLOCK) {
            uint32_t c = hist[i];
            if (c) atomicAdd(&g_hist[i], c);
        }
        return;
    }
    for (int i = tid; i < nb; i += BLOCK) {
        uint32_t c = hist[i];
        uint64_t g = c ? atomicAdd(&bucket_cursor[i], c) : 0;
        // optimistic single pass: buckets have a fixed capacity; a tile that does not fit drops its keys of
        // that bucket and raises the flag (the host then redoes the sample with the exact two-pass path)
        if (c && g + c > bucket_cap) { g = ~0ull; *overflow_flag = 1; }
        dst[i] = g;
    }
    __syncthreads();
    block_exclusive_scan(hist, nb, warp_tmp);
    const uint32_t total = warp_tmp[BLOCK / 32];
    for (int i = tid; i < nb; i += BLOCK) {
        const uint64_t g = dst[i];
        if (g != ~0ull) dst[i] = bucket_base[i] + g - hist[i];
    }
#pragma unroll
    for (int j = 0; j < L1_KPT; j++) {
        if (vmask & (1u << j)) {
            uint32_t r = (j & 1) ? (ranks[j >> 1] >> 16) : (ranks[j >> 1] & 0xFFFFu);
            sorted[hist[(uint32_t)(keys[j] >> bshift)] + r] = keys[j];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < total; i += BLOCK) {
        const uint64_t h = sorted[i];
        const uint64_t d = dst[(uint32_t)(h >> bshift)];
        if (d != ~0ull) keys_out[d + i] = h;
    }
}

// K1, 512 threads x 16 positions (k >= 17): two threads share a packed word, each rolls 16 start positions.  Same
// 8192-key tile and shared-memory footprint as k_extract, but half the registers per thread, so two CTAs = 32 warps
// per SM instead of 16: the kernel is latency-bound on its shared-memory atomics, not on issue slots.
constexpr int K16_THREADS = 512;
template <bool SHARDED>
__global__ void __launch_bounds__(K16_THREADS, 2)
k_extract16(const uint64_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t n_words, KParams kp, int b1_bits,
            const uint64_t *__restrict__ bucket_base, uint32_t *__restrict__ bucket_cursor, uint64_t bucket_cap,
            int *__restrict__ overflow_flag, uint64_t *__restrict__ keys_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *dst = reinterpret_cast<uint64_t *>(smem_raw);            // [MAX_B]
    uint32_t *hist = reinterpret_cast<uint32_t *>(dst + MAX_B);        // [MAX_B]
    uint32_t *warp_tmp = hist + MAX_B;                                 // [32]
    uint64_t *sorted = reinterpret_cast<uint64_t *>(warp_tmp + 32);    // [L1_TILE]
    const int tid = threadIdx.x;
    const int nb = 1 << b1_bits;
    const int k = kp.k;
    for (int i = tid; i < nb; i += K16_THREADS) hist[i] = 0;
    __syncthreads();

    const uint64_t t = (uint64_t)blockIdx.x * (K16_THREADS / 2) + (tid >> 1);  // packed word of this thread pair
    const int half = tid & 1;
    uint64_t hi = 0, lo = 0, bm = ~0ull;
    if (t < n_words) {
        hi = seq[t];
        lo = seq[t + 1];
        bm = ((uint64_t)bad[t] << 32) | bad[t + 1];
    }
    uint32_t okmask;
    {
        uint64_t r = bm;
        int built = 1;
        while (built * 2 <= k) { r |= r << built; built *= 2; }
        if (built < k) r |= r << (k - built);
        okmask = ~(uint32_t)(r >> 32);
        okmask = half ? (okmask << 16) : okmask;  // bit 31-j <-> local position j
    }
    // the 96 bits this thread's 16 windows can touch: (P,Q,R) = the words starting at its first base
    const uint32_t A = (uint32_t)(hi >> 32), B = (uint32_t)hi, Cw = (uint32_t)(lo >> 32), D = (uint32_t)lo;
    const uint32_t P = half ? B : A, Q = half ? Cw : B, R = half ? D : Cw;
    const int s1 = 64 - 2 * k, s2 = 2 * k - 34, sx = kp.mix.shift;
    const uint32_t mhi = (uint32_t)(kp.mix.mask >> 32);
    const uint32_t m1l = (uint32_t)kp.mix.m1, m1h = (uint32_t)(kp.mix.m1 >> 32);
    const int bs = (kp.mix.bits - b1_bits) - 32;
    uint64_t keys[16];
    uint32_t ranks[8];
    uint32_t vmask = 0;
    uint32_t rl = 0, rh = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t w1, w0;
        if (j == 0) { w1 = P; w0 = Q; }
        else { w1 = __funnelshift_l(Q, P, 2 * j); w0 = __funnelshift_l(R, Q, 2 * j); }
        const uint32_t fh = w1 >> s1, fl = __funnelshift_r(w0, w1, s1);
        if (j == 0) {
            const uint64_t r0 = simka::revcomp(((uint64_t)fh << 32) | fl, k);
            rh = (uint32_t)(r0 >> 32); rl = (uint32_t)r0;
        } else {
            rl = __funnelshift_r(rl, rh, 2);
            rh = (rh >> 2) | (((fl & 3u) ^ 2u) << s2);
        }
        const bool fwd_smaller = fh < rh || (fh == rh && fl < rl);
        uint32_t xh = fwd_smaller ? fh : rh, xl = fwd_smaller ? fl : rl;
        xl ^= __funnelshift_r(xl, xh, sx); xh ^= xh >> sx;
        {
            const uint64_t pr = (uint64_t)xl * m1l;
            const uint32_t ph = (uint32_t)(pr >> 32) + xl * m1h + xh * m1l;
            xl = (uint32_t)pr; xh = ph & mhi;
        }
        xl ^= __funnelshift_r(xl, xh, sx); xh ^= xh >> sx;
        const uint64_t h = ((uint64_t)xh << 32) | xl;
        bool keep = (okmask >> (31 - j)) & 1u;
        if (SHARDED) keep = keep && shard_keep(h, kp);
        keys[j] = h;
        if (keep) {
            vmask |= 1u << j;
            const uint32_t r = atomicAdd(&hist[xh >> bs], 1u);
            if (j & 1) ranks[j >> 1] |= r << 16; else ranks[j >> 1] = r;
        } else if (!(j & 1)) ranks[j >> 1] = 0;
    }
    __syncthreads();
    for (int i = tid; i < nb; i += K16_THREADS) {
        const uint32_t c = hist[i];
        uint64_t g = c ? atomicAdd(&bucket_cursor[i], c) : 0;
        if (c && g + c > bucket_cap) { g = ~0ull; *overflow_flag = 1; }
        dst[i] = g;
    }
    __syncthreads();
    block_exclusive_scan<K16_THREADS>(hist, nb, warp_tmp);
    const uint32_t total = warp_tmp[K16_THREADS / 32];
    for (int i = tid; i < nb; i += K16_THREADS) {
        const uint64_t g = dst[i];
        if (g != ~0ull) dst[i] = bucket_base[i] + g - hist[i];
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (vmask & (1u << j)) {
            const uint32_t r = (j & 1) ? (ranks[j >> 1] >> 16) : (ranks[j >> 1] & 0xFFFFu);
            sorted[hist[(uint32_t)(keys[j] >> (kp.mix.bits - b1_bits))] + r] = keys[j];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < total; i += K16_THREADS) {
        const uint64_t h = sorted[i];
        const uint64_t d = dst[(uint32_t)(h >> (kp.mix.bits - b1_bits))];
        if (d != ~0ull) keys_out[d + i] = h;
    }
}

// bucket b of the optimistic single-pass layout starts at b * cap
__global__ void k_fill_bases(uint64_t *__restrict__ base, int n, uint64_t cap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) base[i] = (uint64_t)i * cap;
}

// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------

// out[0..n] = exclusive scan of in[0..n-1] (u32 -> u64); single block, chunked.
__global__ void __launch_bounds__(1024) k_scan_u32_u64(const uint32_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t n) {
    __shared__ uint64_t warp_sum[33];
    __shared__ uint64_t carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    constexpr int PER = 16;  // consecutive elements per thread and round
    for (uint64_t base = 0; base < n; base += 1024 * PER) {
        const uint64_t i0 = base + (uint64_t)tid * PER;
        uint32_t x[PER];
        if (i0 + PER <= n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < PER / 4; q++) {
                const uint4 t4 = *reinterpret_cast<const uint4 *>(in + i0 + 4 * q);
                x[4 * q] = t4.x; x[4 * q + 1] = t4.y; x[4 * q + 2] = t4.z; x[4 * q + 3] = t4.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < PER; q++) x[q] = (i0 + q < n) ? in[i0 + q] : 0;
        }
        uint64_t v = 0;
#pragma unroll
        for (int q = 0; q < PER; q++) v += x[q];
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        if (lane == 31) warp_sum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = warp_sum[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t u = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += u;
            }
            warp_sum[lane] = winc - w;
            if (lane == 31) warp_sum[32] = winc;
        }
        __syncthreads();
        uint64_t run = carry + warp_sum[wid] + inc - v;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (i0 + q < n) out[i0 + q] = run;
            run += x[q];
        }
        __syncthreads();
        if (tid == 0) carry += warp_sum[32];
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

// Multi-block version for the large scans (fine-bin histogram, survivors): block b scans 16384 elements and
// takes its exclusive prefix from block b-1 through a chained (epoch | running total) word.  All blocks are
// co-resident (grid <= 128), so the spin of thread 0 cannot starve its predecessor.
constexpr int SCAN_CHUNK = 1024 * 16;
__global__ void __launch_bounds__(1024) k_scan_chain(const uint32_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t n,
                                                     unsigned long long *chain, unsigned long long epoch) {
    __shared__ uint64_t warp_sum[33];
    __shared__ uint64_t s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)tid * 16;
    uint32_t x[16];
    if (i0 + 16 <= n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 t4 = *reinterpret_cast<const uint4 *>(in + i0 + 4 * q);
            x[4 * q] = t4.x; x[4 * q + 1] = t4.y; x[4 * q + 2] = t4.z; x[4 * q + 3] = t4.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) x[q] = (i0 + q < n) ? in[i0 + q] : 0;
    }
    uint64_t v = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) v += x[q];
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    if (lane == 31) warp_sum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = warp_sum[lane], winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t u = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += u;
        }
        warp_sum[lane] = winc - w;
        if (lane == 31) {
            const uint64_t total = winc;
            uint64_t prefix = 0;
            if (blockIdx.x > 0) {
                volatile unsigned long long *c = chain + (blockIdx.x - 1);
                unsigned long long pv = *c;
                while ((pv >> 40) != epoch) pv = *c;
                prefix = pv & ((1ull << 40) - 1);
            }
            *((volatile unsigned long long *)(chain + blockIdx.x)) = (epoch << 40) | (prefix + total);
            s_prefix = prefix;
            if (blockIdx.x == gridDim.x - 1) out[n] = prefix + total;
        }
    }
    __syncthreads();
    uint64_t run = s_prefix + warp_sum[wid] + inc - v;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        if (i0 + q < n) out[i0 + q] = run;
        run += x[q];
    }
}

// tiles[b] = ceil(count[b] / tile)
__global__ void k_tiles_per_bucket(const uint32_t *__restrict__ count, uint32_t *__restrict__ tiles, int n, int tile) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tiles[i] = (count[i] + tile - 1) / tile;
}

// ---------------------------------------------------------------------------------------------
// K2: level-2 partition of an L1 bucket into fine bins (count pass / scatter pass).
// Same role as the per-partition "read" phase of PartitionsByVectorCommand
// (GC/kmer/impl/PartitionsCommand.cpp:1253-1378), which splits a partition into radix bins.
// ---------------------------------------------------------------------------------------------
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(BLOCK, 4)
k_partition2(const uint64_t *__restrict__ keys_in, const uint64_t *__restrict__ bucket_base /*[B1] begin*/,
             const uint32_t *__restrict__ bucket_cnt /*[B1]*/, const uint64_t *__restrict__ tile_start /*[B1+1]*/, int b1_bits, int b2_bits, int W,
             uint32_t *__restrict__ fine_hist, const uint64_t *__restrict__ fine_off, uint32_t *__restrict__ fine_cursor,
             uint64_t *__restrict__ keys_out) {
    __shared__ uint32_t hist[MAX_B], warp_tmp[16];
    __shared__ __align__(16) uint64_t dst[COUNT_ONLY ? 1 : MAX_B];   // global destination of a bin's run minus its tile-local start
    __shared__ __align__(16) uint64_t sorted[COUNT_ONLY ? 1 : L2_TILE];
    __shared__ int s_bucket;
    const int tid = threadIdx.x;
    const int nb1 = 1 << b1_bits, nb2 = 1 << b2_bits;
    if (tid == 0) {  // which L1 bucket does this tile belong to?
        uint64_t tile = blockIdx.x;
        int lo = 0, hi = nb1;  // tile_start[lo] <= tile < tile_start[hi]
        if (tile >= tile_start[nb1]) s_bucket = -1;
        else {
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (tile_start[mid] <= tile) lo = mid; else hi = mid; }
            s_bucket = lo;
        }
    }
    for (int i = tid; i < nb2; i += BLOCK) hist[i] = 0;
    __syncthreads();
    const int b1 = s_bucket;
    if (b1 < 0) return;
    const uint64_t bbeg = bucket_base[b1], bend = bbeg + bucket_cnt[b1];
    const uint64_t tbeg = bbeg + (blockIdx.x - tile_start[b1]) * (uint64_t)L2_TILE;
    const int shift = W - b1_bits - b2_bits;
    const uint32_t m2 = (uint32_t)nb2 - 1;

    uint64_t keys[L2_KPT];
    uint32_t ranks[L2_KPT / 2];
    uint32_t vmask = 0;
#pragma unroll
    for (int j = 0; j < L2_KPT; j++) {
        uint64_t i = tbeg + (uint64_t)j * BLOCK + tid;
        keys[j] = i < bend ? keys_in[i] : 0;
        if (i < bend) vmask |= 1u << j;
    }
#pragma unroll
    for (int j = 0; j < L2_KPT; j++) {
        if (vmask & (1u << j)) {
            uint32_t r = atomicAdd(&hist[(uint32_t)(keys[j] >> shift) & m2], 1u);
            if (!COUNT_ONLY) { if (j & 1) ranks[j >> 1] |= r << 16; else ranks[j >> 1] = r; }
        } else if (!COUNT_ONLY && !(j & 1)) ranks[j >> 1] = 0;
    }
    __syncthreads();
    const uint32_t fbase = (uint32_t)b1 << b2_bits;
    if (COUNT_ONLY) {
        for (int i = tid; i < nb2; i += BLOCK) { uint32_t c = hist[i]; if (c) atomicAdd(&fine_hist[fbase + i], c); }
        return;
    }
    for (int i = tid; i < nb2; i += BLOCK) { uint32_t c = hist[i]; dst[i] = c ? atomicAdd(&fine_cursor[fbase + i], c) : 0; }
    __syncthreads();
    block_exclusive_scan(hist, nb2, warp_tmp);
    const uint32_t total = warp_tmp[BLOCK / 32];
    for (int i = tid; i < nb2; i += BLOCK) dst[i] = fine_off[fbase + i] + dst[i] - hist[i];
#pragma unroll
    for (int j = 0; j < L2_KPT; j++)
        if (vmask & (1u << j)) {
            uint32_t r = (j & 1) ? (ranks[j >> 1] >> 16) : (ranks[j >> 1] & 0xFFFFu);
            sorted[hist[(uint32_t)(keys[j] >> shift) & m2] + r] = keys[j];
        }
    __syncthreads();
    for (uint32_t i = tid; i < total; i += BLOCK) {
        const uint64_t h = sorted[i];
        keys_out[dst[(uint32_t)(h >> shift) & m2] + i] = h;
    }
}

// ---------------------------------------------------------------------------------------------
// K3: per-bin exact multiset count + abundance filter + per-sample totals.
// Replaces the sort / run-length "dump" of PartitionsByVectorCommand
// (GC/kmer/impl/PartitionsCommand.cpp:1390-1504,1600-1760) and SimkaCompressedProcessor::process
// (src/minikc/MiniKC.hpp:54-79).  Equal keys meet in a shared-memory open-addressing table; order
// inside a bin is irrelevant downstream (the merge regroups by key), so no sort is needed.
// ---------------------------------------------------------------------------------------------
struct CountOut {
    uint64_t *stage_h;      // survivors are written at fine_off[bin] + i
    uint32_t *stage_c;
    uint32_t *surv;         // [nbins] survivor count per bin
    ull *totals;            // [3] distinct, sum c, sum c^2
    int *error_flag;
};

// Insert up to four keys per thread into the open-addressing table (linear probing), the four probe
// sequences interleaved so that the shared-memory latencies overlap and the warp stays converged.
// Plain read first: most keys repeat a k-mer that is already there (one LDS + one ATOMS.POPC.INC).
// A thread that claims a slot appends it to the claimed list, which is what the read-out walks
// (the table itself is never scanned) and what counts the distinct keys of the bin.
__device__ __forceinline__ void table_insert4(uint64_t *tk, uint32_t *tc, uint16_t *claimed, uint32_t *nclaim, const uint64_t (&h)[4],
                                              uint32_t pending, int slot_shift) {
    volatile uint64_t *vk = tk;
    uint32_t slot[4];
#pragma unroll
    for (int q = 0; q < 4; q++) slot[q] = (uint32_t)(h[q] >> slot_shift) & (CNT_SLOTS - 1);
    int rounds = 0;
    while (pending) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (pending & (1u << q)) {
                uint64_t cur = vk[slot[q]];
                if (cur == EMPTY) {
                    cur = atomicCAS(reinterpret_cast<ull *>(&tk[slot[q]]), (ull)EMPTY, (ull)h[q]);
                    if (cur == EMPTY) { cur = h[q]; claimed[atomicAdd(nclaim, 1u)] = (uint16_t)slot[q]; }
                }
                if (cur == h[q]) { atomicAdd(&tc[slot[q]], 1u); pending &= ~(1u << q); }
                else slot[q] = (slot[q] + 1) & (CNT_SLOTS - 1);
            }
        }
        if (++rounds > CNT_SLOTS) return;  // table full: nclaim == CNT_SLOTS > CNT_MAXDISTINCT makes the caller retry finer
    }
}

// MULTI: the bin may hold more distinct keys than one table pass takes: stop early on overflow, R sub-passes
template <bool MULTI>
__device__ __forceinline__ void bin_insert_all(uint64_t *tk, uint32_t *tc, uint16_t *claimed, uint32_t *nclaim,
                                               const uint64_t *__restrict__ kb, uint32_t n, uint32_t R, uint32_t r, int slot_shift) {
    const int tid = threadIdx.x;
    for (uint32_t base = 0; base < n; base += BLOCK * 4) {
        uint64_t h[4];
        uint32_t pending = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = base + q * BLOCK + tid;
            h[q] = i < n ? kb[i] : EMPTY;
            if (i < n) pending |= 1u << q;
        }
        if (MULTI) {
            if (*((volatile uint32_t *)nclaim) > CNT_MAXDISTINCT) return;
            if (R > 1) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if ((((uint32_t)(simka::hash2(h[q]) >> 20)) & (R - 1)) != r) pending &= ~(1u << q);
            }
        }
        table_insert4(tk, tc, claimed, nclaim, h, pending, slot_shift);
    }
}

// Persistent CTAs, bins taken with a grid stride.  Per-bin fixed cost: two barriers; everything else is
// proportional to the keys (insert) and to the distinct keys (read-out of the claimed slots, which also
// leaves the table clean for the next bin).
__global__ void __launch_bounds__(BLOCK)
k_bin_count(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ fine_off, uint32_t nbins, int slot_shift,
            int64_t amin, int64_t amax, CountOut out) {
    __shared__ __align__(16) uint64_t tk[CNT_SLOTS];
    __shared__ __align__(16) uint32_t tc[CNT_SLOTS];
    __shared__ uint16_t claimed[CNT_SLOTS];
    __shared__ uint32_t s_cursor, s_nclaim[2];
    __shared__ ull s_tot[3];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < CNT_SLOTS; i += BLOCK) { tk[i] = EMPTY; tc[i] = 0; }
    if (tid == 0) { s_cursor = 0; s_nclaim[0] = s_nclaim[1] = 0; s_tot[0] = s_tot[1] = s_tot[2] = 0; }
    __syncthreads();
    ull t_d = 0, t_s = 0, t_q = 0;
    // keep iff amin <= c <= amax (src/minikc/MiniKC.hpp:56) with c a non-zero u32
    const bool never = amax < 0 || amin > 0xFFFFFFFFll;
    const uint32_t cmin = amin < 0 ? 0u : (uint32_t)amin;
    const uint32_t cmax = amax > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)(amax < 0 ? 0 : amax);
    uint32_t pass = 0;  // parity selects the claim counter of the current table pass

    uint64_t nbeg = 0, nend = 0;  // offsets of the next bin, loaded one bin ahead
    if (blockIdx.x < nbins) { nbeg = fine_off[blockIdx.x]; nend = fine_off[blockIdx.x + 1]; }
    for (uint32_t bin = blockIdx.x; bin < nbins; bin += gridDim.x) {
        const uint64_t beg = nbeg, end = nend;
        if (bin + gridDim.x < nbins) { nbeg = fine_off[bin + gridDim.x]; nend = fine_off[bin + gridDim.x + 1]; }
        if (end - beg >= 0xFFFFFFFFull) { if (tid == 0) { *out.error_flag = 6; out.surv[bin] = 0; } continue; }
        const uint32_t n = (uint32_t)(end - beg);
        if (n == 0) { if (tid == 0) out.surv[bin] = 0; continue; }
        const uint64_t *kb = keys + beg;
        const bool multi = n > CNT_MAXDISTINCT;
        // R sub-passes (keys selected by a second hash) only when the DISTINCT keys of the bin do not
        // fit the table; a bin that is large because of a few heavy k-mers still takes one pass.
        for (uint32_t R = 1;; R *= 2) {
            if (R > 65536) { if (tid == 0) *out.error_flag = 5; break; }
            ull b_d = 0, b_s = 0, b_q = 0;
            bool overflow = false;
            for (uint32_t r = 0; r < R && !overflow; r++, pass++) {
                uint32_t *nclaim = &s_nclaim[pass & 1];
                if (multi) bin_insert_all<true>(tk, tc, claimed, nclaim, kb, n, R, r, slot_shift);
                else bin_insert_all<false>(tk, tc, claimed, nclaim, kb, n, 1, 0, slot_shift);
                __syncthreads();
                const uint32_t nc = *nclaim;
                overflow = nc > CNT_MAXDISTINCT;
                if (tid == 0) s_nclaim[(pass + 1) & 1] = 0;
                // read-out + clean of the claimed slots
                for (uint32_t e0 = 0; e0 < nc; e0 += BLOCK) {
                    const uint32_t e = e0 + tid;
                    bool keep = false;
                    uint64_t hk = 0;
                    uint32_t c = 0;
                    if (e < nc) {
                        const uint32_t slot = claimed[e];
                        hk = tk[slot];
                        c = tc[slot];
                        tk[slot] = EMPTY;
                        tc[slot] = 0;
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
                            b_d += 1;
                            b_s += c;
                            b_q += (ull)((double)c * (double)c);  // pow(count,2) -> u64, src/minikc/MiniKC.hpp:62
                        }
                    }
                }
                __syncthreads();
            }
            if (!overflow) { t_d += b_d; t_s += b_s; t_q += b_q; break; }
            if (tid == 0) s_cursor = 0;  // forget the aborted attempt (its survivors get overwritten)
            __syncthreads();
        }
        if (tid == 0) { out.surv[bin] = s_cursor; s_cursor = 0; }
        // (s_cursor is next touched by other threads only after the next bin's first __syncthreads)
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

// survivors staging -> exactly-sized per-sample list; one warp per bin, bins taken with a grid stride
__global__ void __launch_bounds__(BLOCK)
k_compact(const uint64_t *__restrict__ stage_h, const uint32_t *__restrict__ stage_c, const uint64_t *__restrict__ fine_off,
          const uint64_t *__restrict__ out_off, uint32_t nbins, uint64_t *__restrict__ eh, uint32_t *__restrict__ ec) {
    const int lane = threadIdx.x & 31;
    const uint32_t nw = gridDim.x * (BLOCK / 32);
    for (uint32_t bin = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); bin < nbins; bin += nw) {
        const uint64_t src = fine_off[bin], dst = out_off[bin];
        const uint32_t n = (uint32_t)(out_off[bin + 1] - dst);
        for (uint32_t i = lane; i < n; i += 32) { eh[dst + i] = stage_h[src + i]; ec[dst + i] = stage_c[src + i]; }
    }
}

// per-sample abundance histogram of the solid k-mers (complex distances only): chist[c] for c < CHIST,
// larger counts go to the (count, sample) overflow list
__global__ void __launch_bounds__(BLOCK)
k_abundance_hist(const uint32_t *__restrict__ ec, uint64_t n, int sample, ull *__restrict__ chist, uint32_t *__restrict__ big_counts,
                 uint32_t *__restrict__ big_sample, uint32_t *__restrict__ big_n, uint32_t big_cap, int *__restrict__ error_flag) {
    __shared__ uint32_t hist[CHIST];
    for (int i = threadIdx.x; i < CHIST; i += BLOCK) hist[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * BLOCK) {
        const uint32_t c = ec[i];
        if (c < CHIST) atomicAdd(&hist[c], 1u);
        else {
            const uint32_t q = atomicAdd(big_n, 1u);
            if (q < big_cap) { big_counts[q] = c; big_sample[q] = (uint32_t)sample; }
            else *error_flag = 2;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CHIST; i += BLOCK) { const uint32_t c = hist[i]; if (c) atomicAdd(&chist[i], (ull)c); }
}

// sum of a u32 array (solid abundances of an imported sample range)
__global__ void __launch_bounds__(BLOCK) k_sum_u32(const uint32_t *__restrict__ v, uint64_t n, ull *__restrict__ out) {
    ull acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * BLOCK) acc += v[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// off_table[fr] = out_off[fr * (nbins / FG)], fr in [0, FG]
__global__ void k_fill_offsets(const uint64_t *__restrict__ out_off, uint32_t *__restrict__ table, uint32_t fg, uint32_t bins_per_fr) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= fg) table[i] = (uint32_t)out_off[(uint64_t)i * bins_per_fr];
}

// h -> k-mer (parity hook)
__global__ void k_unmix(const uint64_t *__restrict__ eh, uint64_t *__restrict__ out, uint64_t n, Mix mix) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = simka::mix_inv(eh[i], mix);
}

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
        if (xY != yX) {  // equal: both logs are log(1) = 0 exactly (same evaluation as k_merge32)
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
        uint32_t R = 1;
        if (n_e > MRG_CAP) { while ((uint64_t)R * (MRG_CAP / 2) < n_e) R <<= 1; }
        for (uint32_t r = 0; r < R; r++) {
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
                    if (++probes >= MRG_SLOTS) { *a.error_flag = 3; break; }
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
            if (tid == 0) { loc_groups += s_groups; loc_shared += s_shared; }
            __syncthreads();
            if (a.count_only) continue;
            if (n_sub > MRG_CAP) { if (tid == 0) *a.error_flag = 4; continue; }
            if (n_shared_groups == 0) continue;  // nothing shared in this bin: no accumulator changes
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
        }
    }
    if (tid == 0 && (loc_groups | loc_shared)) {
        atomicAdd(&a.scalars[0], loc_groups);
        atomicAdd(&a.scalars[1], loc_shared);
    }
}

// ---------------------------------------------------------------------------------------------
// K4 for N <= 32 samples: same grouping as k_merge, but the sample pairs of a bin are counting-sorted
// by pair key (i,j) in shared memory and thread t owns pair key t for the whole kernel: all N x N
// accumulators live in registers, no atomics on the accumulators at all (one flush per CTA at the end).
// ---------------------------------------------------------------------------------------------
constexpr int M32_THREADS = 512;
constexpr int M32_SLOTS = 2048;   // table slots per pass
constexpr int M32_ECAP = 2048;    // grouped entries held per pass
constexpr int M32_GCAP = 1024;    // shared groups per pass (<= ECAP / 2)
constexpr int M32_PCAP = 4096;    // sample pairs sorted per chunk
constexpr int M32_PPT = M32_PCAP / M32_THREADS;
constexpr uint32_t M32_SKIP = 0xFFFFFFFFu;

struct M32Smem {
    uint64_t tk[M32_SLOTS];
    uint2 sp[M32_PCAP];            // (count of the lower sample, count of the higher sample), sorted by pair key
    uint32_t gcnt[M32_SLOTS];      // group size -> fill cursor (end, counting down to start) / M32_SKIP
    uint32_t cg[M32_GCAP];         // compact list of shared groups: start | size << 16
    uint32_t cgp[M32_GCAP + 1];    // exclusive pair prefix over cg
    uint32_t ge_c[M32_ECAP];       // grouped counts
    uint32_t kh[M32_THREADS + 1];  // per pair key: histogram -> start
    uint32_t kcur[M32_THREADS];    // per pair key: fill cursor -> end
    uint32_t s_start[32], s_pre[33];
    unsigned long long scan_tmp[M32_THREADS / 32 + 1];
    double2 logtab[128];           // tab_log table
    uint16_t lut[1024];            // (i*32+j) -> pair key
    uint8_t ge_s[M32_ECAP];        // grouped sample ids
};

__device__ __forceinline__ unsigned long long block_scan_u64(unsigned long long v, unsigned long long *tmp, unsigned long long *total) {
    // exclusive scan over the M32_THREADS threads of the block; tmp has M32_THREADS/32+1 entries
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
        unsigned long long w = lane < M32_THREADS / 32 ? tmp[lane] : 0, winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += u;
        }
        if (lane < M32_THREADS / 32) tmp[lane] = winc - w;
        if (lane == M32_THREADS / 32 - 1) tmp[M32_THREADS / 32] = winc;
    }
    __syncthreads();
    *total = tmp[M32_THREADS / 32];
    return tmp[wid] + inc - v;
}

template <bool SIMPLE, bool COMPLEX>
__global__ void __launch_bounds__(M32_THREADS, 2)
k_merge32(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    M32Smem &S = *reinterpret_cast<M32Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = a.n;
    const int NP = N * (N - 1) / 2;
    const uint32_t MG = 1u << a.mg_bits;
    const uint32_t q = 1u << (a.fg_bits - a.mg_bits);
    const size_t ostride = ((size_t)1 << a.fg_bits) + 1;

    for (int idx = tid; idx < 1024; idx += M32_THREADS) {
        const int i = idx >> 5, j = idx & 31;
        S.lut[idx] = (i < j && j < N) ? (uint16_t)(i * (2 * N - i - 1) / 2 + (j - i - 1)) : (uint16_t)0;
    }
    S.kh[tid] = 0;
    if (tid == 0) S.kh[M32_THREADS] = 0;
    if (tid < 128) S.logtab[tid] = a.logtab[tid];
    // the pair key this thread owns
    int my_i = 0, my_j = 0;
    if (tid < NP) {
        int i = 0;
        while ((i + 1) * (2 * N - (i + 1) - 1) / 2 <= tid) i++;
        my_i = i;
        my_j = i + 1 + (tid - i * (2 * N - i - 1) / 2);
    }
    const double Ni = a.ntot[my_i], Nj = a.ntot[my_j];
    const double invNi = 1.0 / Ni, invNj = 1.0 / Nj;
    ull r_sh_ij = 0, r_sh_ji = 0, r_dist = 0, r_bc = 0, r_chord = 0, r_hell = 0, r_whit = 0;
    double r_kl = 0.0;
    ull loc_groups = 0, loc_shared = 0, loc_pairs = 0;
    __syncthreads();

    for (uint32_t mb = blockIdx.x; mb < MG; mb += gridDim.x) {
        if (tid < 32) {
            uint32_t len = 0;
            if (tid < N) {
                const uint32_t b0 = a.off[(size_t)tid * ostride + (size_t)mb * q];
                const uint32_t b1 = a.off[(size_t)tid * ostride + (size_t)(mb + 1) * q];
                S.s_start[tid] = b0;
                len = b1 - b0;
            }
            uint32_t inc = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += u;
            }
            S.s_pre[tid] = inc - len;
            if (tid == 31) S.s_pre[32] = inc;
        }
        __syncthreads();
        const uint32_t n_e = S.s_pre[32];
        if (n_e == 0) { __syncthreads(); continue; }
        uint32_t R = 1;
        if (n_e > M32_ECAP) { while ((uint64_t)R * (M32_ECAP / 2) < n_e) R <<= 1; }
        for (uint32_t r = 0; r < R; r++) {
            {   // clean table
                uint4 *pk = reinterpret_cast<uint4 *>(S.tk) + tid;
                const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), z = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int qq = 0; qq < M32_SLOTS * 8 / 16 / M32_THREADS; qq++) pk[qq * M32_THREADS] = e;
                uint4 *pg = reinterpret_cast<uint4 *>(S.gcnt) + tid;
#pragma unroll
                for (int qq = 0; qq < M32_SLOTS * 4 / 16 / M32_THREADS; qq++) pg[qq * M32_THREADS] = z;
            }
            __syncthreads();
            // phase 1: insert keys, count group sizes
            for (uint32_t i = tid; i < n_e; i += M32_THREADS) {
                int lo = 0, hi = 32;  // s_pre[lo] <= i < s_pre[hi]   (s_pre[N..32] = n_e)
#pragma unroll
                for (int st = 0; st < 5; st++) { const int mid = (lo + hi) >> 1; if (S.s_pre[mid] <= i) lo = mid; else hi = mid; }
                const uint64_t h = a.eh[lo][S.s_start[lo] + (i - S.s_pre[lo])];
                const uint64_t hh = simka::hash2(h);
                if (R > 1 && (((uint32_t)(hh >> 20)) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 40) & (M32_SLOTS - 1);
                volatile uint64_t *vk = S.tk;
                int probes = 0;
                for (;;) {
                    uint64_t cur = vk[slot];
                    if (cur == EMPTY) {
                        cur = atomicCAS(reinterpret_cast<ull *>(&S.tk[slot]), (ull)EMPTY, (ull)h);
                        if (cur == EMPTY) cur = h;
                    }
                    if (cur == h) { atomicAdd(&S.gcnt[slot], 1u); break; }
                    slot = (slot + 1) & (M32_SLOTS - 1);
                    if (++probes >= M32_SLOTS) { *a.error_flag = 3; break; }
                }
            }
            __syncthreads();
            // phase 2: group statistics + compact list of the shared groups (>= 2 samples)
            unsigned long long packed = 0;  // groups (16 bits) | entries (16 bits) | pairs (32 bits)
            uint32_t g8[M32_SLOTS / M32_THREADS];
#pragma unroll
            for (int qq = 0; qq < M32_SLOTS / M32_THREADS; qq++) {
                const uint32_t g = S.gcnt[tid * (M32_SLOTS / M32_THREADS) + qq];
                g8[qq] = g;
                loc_groups += g > 0;
                if (g > 1) packed += 1ull | ((unsigned long long)g << 16) | ((unsigned long long)(g * (g - 1) / 2) << 32);
            }
            unsigned long long total;
            unsigned long long ex = block_scan_u64(packed, S.scan_tmp, &total);
            const uint32_t n_groups = (uint32_t)(total & 0xFFFF), n_sub = (uint32_t)((total >> 16) & 0xFFFF);
            const uint32_t P = (uint32_t)(total >> 32);
            if ((total & 0xFFFF) != 0 && (n_groups > M32_GCAP || n_sub > M32_ECAP)) {  // cannot happen for hashed keys; reported, not silent
                if (tid == 0) *a.error_flag = 4;
                __syncthreads();
                continue;
            }
#pragma unroll
            for (int qq = 0; qq < M32_SLOTS / M32_THREADS; qq++) {
                const uint32_t g = g8[qq];
                uint32_t v = M32_SKIP;
                if (g > 1) {
                    const uint32_t gi = (uint32_t)(ex & 0xFFFF), st = (uint32_t)((ex >> 16) & 0xFFFF), pp = (uint32_t)(ex >> 32);
                    S.cg[gi] = st | (g << 16);
                    S.cgp[gi] = pp;
                    v = st + g;  // fill cursor counts down from the end
                    ex += 1ull | ((unsigned long long)g << 16) | ((unsigned long long)(g * (g - 1) / 2) << 32);
                    loc_shared++;
                }
                S.gcnt[tid * (M32_SLOTS / M32_THREADS) + qq] = v;
            }
            if (tid == 0) { S.cgp[n_groups] = P; loc_pairs += P; }
            __syncthreads();
            if (n_groups == 0) continue;  // nothing shared in this pass: no accumulator changes
            // phase 3: scatter the entries of shared k-mers into their groups
            for (uint32_t i = tid; i < n_e; i += M32_THREADS) {
                int lo = 0, hi = 32;
#pragma unroll
                for (int st = 0; st < 5; st++) { const int mid = (lo + hi) >> 1; if (S.s_pre[mid] <= i) lo = mid; else hi = mid; }
                const size_t idx = S.s_start[lo] + (i - S.s_pre[lo]);
                const uint64_t h = a.eh[lo][idx];
                const uint64_t hh = simka::hash2(h);
                if (R > 1 && (((uint32_t)(hh >> 20)) & (R - 1)) != r) continue;
                uint32_t slot = (uint32_t)(hh >> 40) & (M32_SLOTS - 1);
                int pr = 0;
                while (S.tk[slot] != h && ++pr <= M32_SLOTS) slot = (slot + 1) & (M32_SLOTS - 1);
                if (pr > M32_SLOTS) { *a.error_flag = 7; continue; }  // inconsistent input (never for lists built by this library)
                if (S.gcnt[slot] == M32_SKIP) continue;  // private k-mer: contributes to no pair
                const uint32_t pos = atomicSub(&S.gcnt[slot], 1u) - 1;
                S.ge_s[pos] = (uint8_t)lo;
                S.ge_c[pos] = a.ec[lo][idx];
            }
            __syncthreads();
            // phase 4: pairs, in chunks: counting sort by pair key, then every thread consumes its own key
            for (uint32_t c0 = 0; c0 < P; c0 += M32_PCAP) {
                const uint32_t c1 = min(P, c0 + M32_PCAP);
                const uint32_t per = (c1 - c0 + M32_THREADS - 1) / M32_THREADS;
                uint32_t l0 = c0 + tid * per, l1 = min(l0 + per, c1);
                uint16_t pkey[M32_PPT];
                uint32_t pa[M32_PPT], pb[M32_PPT];
                int np = 0;
                if (l0 < l1) {
                    int lo = 0, hi = (int)n_groups;  // cgp[lo] <= l0 < cgp[hi]
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.cgp[mid] <= l0) lo = mid; else hi = mid; }
                    uint32_t g = (uint32_t)lo;
                    uint32_t cgv = S.cg[g];
                    uint32_t st = cgv & 0xFFFF, sz = cgv >> 16;
                    const uint32_t l = l0 - S.cgp[g];
                    // row-major strict upper triangle of sz x sz: row x starts at x(2 sz - x - 1)/2
                    const float tg = 2.0f * (float)sz - 1.0f;
                    int x = (int)((tg - sqrtf(tg * tg - 8.0f * (float)l)) * 0.5f);
                    if (x < 0) x = 0;
                    while ((uint32_t)(x + 1) * (2 * sz - (x + 1) - 1) / 2 <= l) x++;
                    while ((uint32_t)x * (2 * sz - x - 1) / 2 > l) x--;
                    uint32_t y = x + 1 + (l - (uint32_t)x * (2 * sz - x - 1) / 2);
                    uint32_t sx = S.ge_s[st + x], cx = S.ge_c[st + x];
#pragma unroll
                    for (int k = 0; k < M32_PPT; k++) {
                        if (l0 + k < l1) {
                            const uint32_t sy = S.ge_s[st + y], cy = S.ge_c[st + y];
                            const bool sw = sx > sy;
                            pkey[k] = S.lut[sw ? (sy * 32 + sx) : (sx * 32 + sy)];
                            pa[k] = sw ? cy : cx;
                            pb[k] = sw ? cx : cy;
                            atomicAdd(&S.kh[pkey[k]], 1u);
                            np = k + 1;
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
                }
                // histogram -> start offsets
                unsigned long long tot2;
                __syncthreads();
                const uint32_t myh = S.kh[tid];
                const uint32_t exk = (uint32_t)block_scan_u64((unsigned long long)myh, S.scan_tmp, &tot2);
                S.kh[tid] = exk;
                S.kcur[tid] = exk;
                __syncthreads();
#pragma unroll
                for (int k = 0; k < M32_PPT; k++) {
                    if (k < np) {
                        const uint32_t pos = atomicAdd(&S.kcur[pkey[k]], 1u);
                        S.sp[pos] = make_uint2(pa[k], pb[k]);
                    }
                }
                __syncthreads();
                if (tid < NP) {
                    const uint32_t pbeg = S.kh[tid], pend = S.kcur[tid];
                    for (uint32_t pos = pbeg; pos < pend; pos++) {
                        const uint2 cc = S.sp[pos];
                        const uint32_t ci = cc.x, cj = cc.y;
                        r_sh_ij += ci;
                        r_sh_ji += cj;
                        r_dist += 1;
                        r_bc += ci < cj ? ci : cj;
                        if (SIMPLE) {
                            const ull prod = (ull)ci * (ull)cj;
                            r_chord += prod;
                            r_hell += hellinger_term(prod);  // u64 += double truncates per increment
                        }
                        if (COMPLEX) {
                            const double ai = (double)ci, aj = (double)cj;
                            const double yX = aj * Ni, xY = ai * Nj;
                            if (xY != yX) {  // equal: both logs are log(1) = 0 exactly
                                // r1 = 2 xY / S, r2 = 2 yX / S = 2 - r1 (absolute error eps; harmless: the term it
                                // feeds is weighted by aj/Nj ~ r2)
                                const double r1 = (2 * xY) * __drcp_rn(xY + yX);
                                r_kl += (ai * invNi) * tab_log(r1, S.logtab) + (aj * invNj) * tab_log(2.0 - r1, S.logtab);
                            }
                            const uint64_t wx = (uint64_t)xY, wy = (uint64_t)yX;
                            r_whit += (ull)(absint32_ext(wx - wy) - absint32_ext(wx) - absint32_ext(wy));
                        }
                    }
                }
                S.kh[tid] = 0;
                __syncthreads();
            }
        }
        __syncthreads();
    }
    // flush: one set of global atomics per CTA and pair key
    if (tid < NP && r_dist) {
        const size_t n = (size_t)N;
        const size_t ij = (size_t)my_i * n + my_j, ji = (size_t)my_j * n + my_i;
        const size_t sym = (size_t)my_j + (n - 1) * my_i - ((size_t)my_i * (my_i - 1) / 2);
        atomicAdd(&a.nb_shared[ij], r_sh_ij);
        atomicAdd(&a.nb_shared[ji], r_sh_ji);
        atomicAdd(&a.distinct_shared[sym], r_dist);
        atomicAdd(&a.braycurtis[sym], r_bc);
        if (SIMPLE) { atomicAdd(&a.chord[ij], r_chord); atomicAdd(&a.hellinger[ij], r_hell); }
        if (COMPLEX) { atomicAdd(&a.whittaker[ij], r_whit); atomicAdd(&a.kl[ij], r_kl); }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        loc_groups += __shfl_down_sync(0xffffffffu, loc_groups, d);
        loc_shared += __shfl_down_sync(0xffffffffu, loc_shared, d);
        loc_pairs += __shfl_down_sync(0xffffffffu, loc_pairs, d);
    }
    if (lane == 0 && (loc_groups | loc_shared | loc_pairs)) {
        atomicAdd(&a.scalars[0], loc_groups);
        atomicAdd(&a.scalars[1], loc_shared);
        atomicAdd(&a.scalars[2], loc_pairs);
    }
}

// ---------------------------------------------------------------------------------------------
// K4 for N <= 64 samples, the shipped merge: same grouping as k_merge32, then the sample pairs of the
// bin are enumerated FLAT (every thread takes a contiguous run of pairs: one search + decode per run, full
// lane utilisation) and added to CTA-private N(N-1)/2-entry accumulators in shared memory with 32-bit
// shared atomics (64-bit values as lo/hi words with an explicit carry).  One flush of the accumulators per
// CTA at the end.  k_merge32's counting sort + register accumulators avoid the atomics but lose two thirds
// of the lanes to uneven per-key list lengths; measured, this version is faster.
// ---------------------------------------------------------------------------------------------
constexpr int MS_THREADS = 512;
constexpr int MS_SLOTS = 2048;
constexpr int MS_ECAP = 2048;
constexpr int MS_GCAP = 1024;
constexpr int MS_MAXN = 64;
constexpr int MS_NACC = 8;  // u64 accumulators per pair key: sh_ij, sh_ji, distinct, bc, chord, hellinger, whittaker (negated), JS (2^-60)

struct MSSmem {
    alignas(16) uint64_t tk[MS_SLOTS];
    alignas(16) double2 logtab[128];
    alignas(16) unsigned long long scan_tmp[MS_THREADS / 32 + 1];
    alignas(16) uint32_t gcnt[MS_SLOTS];   // cleaned with 128-bit stores
    uint32_t cg[MS_GCAP];
    uint32_t cgp[MS_GCAP + 1];
    uint32_t ge_c[MS_ECAP];
    uint32_t s_start[MS_MAXN], s_pre[MS_MAXN + 1];
    int s_gid[MS_MAXN];            // local sample slot -> global sample id
    uint8_t ge_s[MS_ECAP];
    // followed by: uint32_t lo[MS_NACC][NP]; uint32_t hi[MS_NACC][NP];
};
static size_t ms_smem_bytes(int n) {
    const size_t np = n > MS_MAXN ? 1024 : (size_t)n * (n - 1) / 2;  // tiled: 32 x 32 pair keys
    return ((sizeof(MSSmem) + 15) & ~(size_t)15) + (size_t)2 * MS_NACC * np * 4 + 16;
}

__device__ __forceinline__ void ms_add64(uint32_t *lo, uint32_t *hi, uint32_t idx, unsigned long long v) {
    const uint32_t vl = (uint32_t)v;
    const uint32_t old = atomicAdd(&lo[idx], vl);
    const uint32_t hv = (uint32_t)(v >> 32) + ((uint32_t)(old + vl) < old ? 1u : 0u);
    if (hv) atomicAdd(&hi[idx], hv);
}

// TILED (N > 64): the sample pairs are cut into tiles of 32 x 32 samples (block I x block J, I <= J); a CTA works on
// one tile: it gathers only the <= 64 samples of its two blocks, groups them, and accumulates the pairs that belong
// to the tile (cross pairs for I != J, all pairs for I == J) into 1024 (496) shared-memory accumulators.  Grouping
// work is repeated once per tile a block takes part in (B+1 times for B blocks); the pair work is not.
template <bool SIMPLE, bool COMPLEX, bool TILED>
__global__ void __launch_bounds__(MS_THREADS, 2)
k_merge_sm(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MSSmem &S = *reinterpret_cast<MSSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = a.n;
    uint32_t *acc_lo = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(MSSmem) + 15) & ~(size_t)15));
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
        uint32_t *acc_hi = acc_lo + (size_t)MS_NACC * NPk;
        __syncthreads();  // the flush of the previous tile is done
        for (int i = tid; i < 2 * MS_NACC * NPk; i += MS_THREADS) acc_lo[i] = 0;  // lo and hi are contiguous
        if (tid < MS_MAXN) {  // local sample slot -> global sample id (-1: unused)
            int gs = tid;
            if (TILED) gs = tid < 32 ? (int)(32 * I) + tid : (diag ? -1 : (int)(32 * J) + tid - 32);
            S.s_gid[tid] = (gs >= 0 && gs < N) ? gs : -1;
        }
        __syncthreads();

        for (uint32_t mb = bin_first; mb < MG; mb += bin_stride) {
            // segment descriptors of the (<= 64) samples of this tile
            if (tid < MS_MAXN) {
                uint32_t len = 0;
                const int gs = S.s_gid[tid];
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
                    const uint64_t h = a.eh[S.s_gid[lo]][S.s_start[lo] + (i - S.s_pre[lo])];
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
                    uint32_t v = M32_SKIP;
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
                    const int gsid = S.s_gid[lo];
                    const uint64_t h = a.eh[gsid][idx];
                    const uint64_t hh = simka::hash2(h);
                    if (R > 1 && (((uint32_t)(hh >> 20)) & (R - 1)) != r) continue;
                    uint32_t slot = (uint32_t)(hh >> 40) & (MS_SLOTS - 1);
                    int pr = 0;
                    while (S.tk[slot] != h && ++pr <= MS_SLOTS) slot = (slot + 1) & (MS_SLOTS - 1);
                    if (pr > MS_SLOTS) { *a.error_flag = 7; continue; }  // inconsistent input (never for lists built by this library)
                    if (S.gcnt[slot] == M32_SKIP) continue;  // private k-mer: contributes to no pair
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
                                ms_add64(acc_lo + 0 * NPk, acc_hi + 0 * NPk, key, ci);
                                ms_add64(acc_lo + 1 * NPk, acc_hi + 1 * NPk, key, cj);
                                ms_add64(acc_lo + 2 * NPk, acc_hi + 2 * NPk, key, 1);
                                ms_add64(acc_lo + 3 * NPk, acc_hi + 3 * NPk, key, ci < cj ? ci : cj);
                                if (SIMPLE) {
                                    const ull prod = (ull)ci * (ull)cj;
                                    ms_add64(acc_lo + 4 * NPk, acc_hi + 4 * NPk, key, prod);
                                    ms_add64(acc_lo + 5 * NPk, acc_hi + 5 * NPk, key, hellinger_term(prod));  // u64 += double truncates per increment
                                }
                                if (COMPLEX) {
                                    const double ai = (double)ci, aj = (double)cj, Ni = a.ntot[S.s_gid[si]], Nj = a.ntot[S.s_gid[sj]];
                                    const double yX = aj * Ni, xY = ai * Nj;
                                    if (xY != yX) {  // equal: both logs are log(1) = 0 exactly
                                        const double r1 = (2 * xY) * __drcp_rn(xY + yX);
                                        // the term d1 + d2 >= 0 (Jensen), |term| < 2: exact integer accumulation in units of 2^-60
                                        // (rounding 4e-19 per term, order-independent => bit-reproducible sums)
                                        const double term = (ai / Ni) * tab_log(r1, S.logtab) + (aj / Nj) * tab_log(2.0 - r1, S.logtab);
                                        ms_add64(acc_lo + 7 * NPk, acc_hi + 7 * NPk, key, (ull)__double2ll_rn(ldexp(term, 60)));
                                    }
                                    const uint64_t wx = (uint64_t)xY, wy = (uint64_t)yX;
                                    // accumulated negated (the both-present value minus the two one-sided values is <= 0 in practice)
                                    ms_add64(acc_lo + 6 * NPk, acc_hi + 6 * NPk, key, (ull)(absint32_ext(wx) + absint32_ext(wy) - absint32_ext(wx - wy)));
                                }
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
            }
        }
        __syncthreads();
        // flush: one set of global atomics per CTA, tile and pair key
        for (int key = tid; key < NPk; key += MS_THREADS) {
            const ull dist = ((ull)acc_hi[2 * NPk + key] << 32) | acc_lo[2 * NPk + key];
            if (!dist) continue;
            int li, lj;  // local slots of the pair
            if (TILED && !diag) { li = key >> 5; lj = 32 + (key & 31); }
            else {
                const int n_loc = TILED ? 32 : N;
                li = 0;
                while ((li + 1) * (2 * n_loc - (li + 1) - 1) / 2 <= key) li++;
                lj = li + 1 + (key - li * (2 * n_loc - li - 1) / 2);
            }
            const int i = S.s_gid[li], j = S.s_gid[lj];
            const size_t n = (size_t)N;
            const size_t ij = (size_t)i * n + j, ji = (size_t)j * n + i;
            const size_t sym = (size_t)j + (n - 1) * i - ((size_t)i * (i - 1) / 2);
            auto val = [&](int w) { return ((ull)acc_hi[(size_t)w * NPk + key] << 32) | acc_lo[(size_t)w * NPk + key]; };
            atomicAdd(&a.nb_shared[ij], val(0));
            atomicAdd(&a.nb_shared[ji], val(1));
            atomicAdd(&a.distinct_shared[sym], dist);
            atomicAdd(&a.braycurtis[sym], val(3));
            if (SIMPLE) { atomicAdd(&a.chord[ij], val(4)); atomicAdd(&a.hellinger[ij], val(5)); }
            if (COMPLEX) { atomicAdd(&a.whittaker[ij], 0ull - val(6)); atomicAdd(&a.klfix[ij], val(7)); }
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
// SimkaMin: all-pairs bottom-s sketch comparison.
// Replaces ComputeDistanceManager::computeDistance_unsynch (src/simkaMin/SimkaMinDistance.hpp:190-279).
// One CTA per pair; the merged prefix that holds the s smallest distinct hashes of the union is
// located with a merge-path binary search, then counted in parallel (hash intersections by
// per-thread chunked two-pointer walks).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t *a, uint32_t n, uint64_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(BLOCK)
k_simkamin_pair(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ start,
                uint32_t n, uint32_t s, ull *__restrict__ quad /*[pairs][4]: shared, distinct, sharedK, kmers*/) {
    // pair index -> (i, j), i < j
    const uint64_t p = blockIdx.x;
    const double tn = 2.0 * n - 1.0;
    int i = (int)((tn - sqrt(tn * tn - 8.0 * (double)p)) * 0.5);
    if (i < 0) i = 0;
    while ((uint64_t)(i + 1) * (2ull * n - (i + 1) - 1) / 2 <= p) i++;
    while ((uint64_t)i * (2ull * n - i - 1) / 2 > p) i--;
    const int j = i + 1 + (int)(p - (uint64_t)i * (2ull * n - i - 1) / 2);

    const uint64_t *A = hashes + (size_t)i * s + start[i];
    const uint64_t *B = hashes + (size_t)j * s + start[j];
    const uint32_t *CA = counts + (size_t)i * s + start[i];
    const uint32_t *CB = counts + (size_t)j * s + start[j];
    const uint32_t na = s - start[i], nb = s - start[j];
    const uint32_t S = min(na, nb);  // sketchSize = min of the two stripped sizes (:204)

    __shared__ ull s_shared, s_sharedk, s_kmers;
    __shared__ uint32_t s_ea, s_eb, s_distinct;
    __shared__ uint32_t sh_cnt[BLOCK + 1];
    const int tid = threadIdx.x;
    if (tid == 0) { s_shared = s_sharedk = s_kmers = 0; }

    // The reference walk consumes the merged sequence in ascending order and stops once S distinct
    // values were seen (a side can only run out at that same step, because the union rank of a
    // side's last value is >= its length >= S).  So the consumed prefixes (ea, eb) hold exactly the
    // S smallest distinct values of A u B.
    // Phase A: per-thread chunks of A[0..min(na,S)) count their matches in B; the chunk prefix makes
    // the union rank searchable at chunk granularity.
    const uint32_t naE = min(na, S);
    const uint32_t per = (naE + BLOCK - 1) / BLOCK;
    const uint32_t a0 = min(naE, tid * per), a1 = min(naE, a0 + per);
    uint32_t my_sh = 0;
    if (a0 < a1) {
        uint32_t b = lower_bound_u64(B, nb, A[a0]);
        for (uint32_t x = a0; x < a1 && b < nb; x++) {
            const uint64_t v = A[x];
            while (b < nb && B[b] < v) b++;
            if (b < nb && B[b] == v) my_sh++;
        }
    }
    sh_cnt[tid] = my_sh;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int t = 0; t < BLOCK; t++) { uint32_t v = sh_cnt[t]; sh_cnt[t] = run; run += v; }
        sh_cnt[BLOCK] = run;
        // state after chunk t: x = end of chunk, y = #B smaller than the next A element
        uint32_t x = 0, y = 0, d = 0;
        for (int t = 0; t < BLOCK; t++) {
            const uint32_t ca0 = min(naE, (uint32_t)t * per), ca1 = min(naE, ca0 + per);
            if (ca0 >= ca1) break;
            const uint32_t yb = (ca1 < na) ? lower_bound_u64(B, nb, A[ca1]) : nb;
            const uint32_t dn = ca1 + yb - sh_cnt[t + 1];
            if (dn >= S) break;  // the stop lies inside this chunk
            x = ca1; y = yb; d = dn;
        }
        while (d < S && x < na && y < nb) {  // sequential refinement, at most one chunk long
            const uint64_t va = A[x], vb = B[y];
            x += va <= vb;
            y += vb <= va;
            d++;
        }
        // (unreachable for well-formed sketches: one side ran out before S distinct values)
        s_ea = x; s_eb = y; s_distinct = d;
    }
    __syncthreads();
    const uint32_t ea = s_ea, eb = s_eb;
    // Phase B: sums over the consumed prefixes.
    ull kmers = 0, shared = 0, sharedk = 0;
    for (uint32_t x = tid; x < ea; x += BLOCK) kmers += CA[x];
    for (uint32_t y = tid; y < eb; y += BLOCK) kmers += CB[y];
    {
        const uint32_t per2 = (ea + BLOCK - 1) / BLOCK;
        const uint32_t c0 = min(ea, tid * per2), c1 = min(ea, c0 + per2);
        if (c0 < c1) {
            uint32_t b = lower_bound_u64(B, eb, A[c0]);
            for (uint32_t x = c0; x < c1 && b < eb; x++) {
                const uint64_t v = A[x];
                while (b < eb && B[b] < v) b++;
                if (b < eb && B[b] == v) { shared++; sharedk += min(CA[x], CB[b]); }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        kmers += __shfl_down_sync(0xffffffffu, kmers, d);
        shared += __shfl_down_sync(0xffffffffu, shared, d);
        sharedk += __shfl_down_sync(0xffffffffu, sharedk, d);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_kmers, kmers); atomicAdd(&s_shared, shared); atomicAdd(&s_sharedk, sharedk); }
    __syncthreads();
    if (tid == 0) {
        quad[4 * p + 0] = s_shared;
        quad[4 * p + 1] = s_distinct;
        quad[4 * p + 2] = s_sharedk;
        quad[4 * p + 3] = s_kmers;
    }
}

// ---------------------------------------------------------------------------------------------
// SimkaMin, tiled version (the shipped one): one CTA per pair streams the two sketches through shared memory in
// tiles of SMT_TILE MERGED elements (coalesced loads, merge-path split per thread, A before B on ties), and stops at
// the tile that holds the (S+1)-th distinct value of the union -- exactly the elements the reference's two-pointer
// loop consumes (src/simkaMin/SimkaMinDistance.hpp:204-243): the S smallest distinct values, both copies of a shared one.
// ---------------------------------------------------------------------------------------------
constexpr int SMT_EPT = 6;                    // merged elements per thread and tile
constexpr int SMT_TILE = BLOCK * SMT_EPT;     // 1536

__device__ __forceinline__ void block_sum3(ull &a, ull &b, ull &c, ull *tmp /*[3*8]*/) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, d);
        b += __shfl_down_sync(0xffffffffu, b, d);
        c += __shfl_down_sync(0xffffffffu, c, d);
    }
    __syncthreads();
    if (lane == 0) { tmp[wid] = a; tmp[8 + wid] = b; tmp[16 + wid] = c; }
    __syncthreads();
    a = b = c = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) { a += tmp[w]; b += tmp[8 + w]; c += tmp[16 + w]; }
}

__global__ void __launch_bounds__(BLOCK)
k_simkamin_tiled(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ start,
                 uint32_t n, uint32_t s, ull *__restrict__ quad /*[pairs][4]: shared, distinct, sharedK, kmers*/) {
    static_assert(BLOCK == 256, "block_sum3 assumes 8 warps");
    const uint64_t p = blockIdx.x;
    const double tn = 2.0 * n - 1.0;
    int i = (int)((tn - sqrt(tn * tn - 8.0 * (double)p)) * 0.5);
    if (i < 0) i = 0;
    while ((uint64_t)(i + 1) * (2ull * n - (i + 1) - 1) / 2 <= p) i++;
    while ((uint64_t)i * (2ull * n - i - 1) / 2 > p) i--;
    const int j = i + 1 + (int)(p - (uint64_t)i * (2ull * n - i - 1) / 2);
    const uint64_t *A = hashes + (size_t)i * s + start[i];
    const uint64_t *B = hashes + (size_t)j * s + start[j];
    const uint32_t *CA = counts + (size_t)i * s + start[i];
    const uint32_t *CB = counts + (size_t)j * s + start[j];
    const uint32_t na = s - start[i], nb = s - start[j];
    const uint32_t S = min(na, nb);  // sketchSize = min of the two stripped sizes (:204)

    __shared__ uint64_t sa[SMT_TILE], sb[SMT_TILE];
    __shared__ uint32_t ca[SMT_TILE], cb[SMT_TILE];
    __shared__ ull red[24];
    __shared__ uint32_t s_scan[BLOCK / 32 + 1];
    __shared__ uint32_t s_xs, s_tstar;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    uint32_t a0 = 0, b0 = 0;            // elements of A / B consumed so far
    ull distinct = 0, shared = 0, sharedk = 0, kmers = 0;
    uint64_t prevA = ~0ull;             // last A element consumed before this tile (a duplicate B may open the tile)
    uint32_t prevCA = 0;
    for (;;) {
        const uint32_t ra = na - a0, rb = nb - b0;
        const uint32_t la = min(ra, (uint32_t)SMT_TILE), lb = min(rb, (uint32_t)SMT_TILE);
        const uint32_t tl = min((uint32_t)SMT_TILE, ra + rb);
        if (tl == 0) break;  // (distinct == S at a tile end is NOT a stop: a duplicate B of the last A may open the next tile)
        __syncthreads();  // previous tile fully consumed
        for (uint32_t q = tid; q < la; q += BLOCK) { sa[q] = A[a0 + q]; ca[q] = CA[a0 + q]; }
        for (uint32_t q = tid; q < lb; q += BLOCK) { sb[q] = B[b0 + q]; cb[q] = CB[b0 + q]; }
        __syncthreads();
        // merge-path split of this thread's diagonal: x = elements of A among the first d0 merged ones (A first on ties)
        const uint32_t d0 = min(tl, (uint32_t)tid * SMT_EPT), d1 = min(tl, d0 + SMT_EPT);
        uint32_t x;
        {
            uint32_t lo = d0 > lb ? d0 - lb : 0, hi = min(d0, la);
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sa[mid] <= sb[d0 - 1 - mid]) lo = mid + 1; else hi = mid; }
            x = lo;
        }
        uint32_t y = d0 - x;
        const uint32_t x0 = x, y0 = y;
        if (tid == BLOCK - 1 || d1 == tl) {  // (several threads may write the same value)
            uint32_t lo = tl > lb ? tl - lb : 0, hi = min(tl, la);
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sa[mid] <= sb[tl - 1 - mid]) lo = mid + 1; else hi = mid; }
            s_xs = lo;
        }
        // walk: matches are detected on the B copy (the A copy of a shared value precedes it)
        uint32_t t_match = 0;
        ull t_k = 0, t_sk = 0;
        for (uint32_t d = d0; d < d1; d++) {
            const bool takeA = y >= lb || (x < la && sa[x] <= sb[y]);
            if (takeA) { t_k += ca[x]; x++; }
            else {
                const uint64_t pa = x > 0 ? sa[x - 1] : prevA;
                if (pa == sb[y]) { t_match++; const uint32_t pc = x > 0 ? ca[x - 1] : prevCA; t_sk += min(pc, cb[y]); }
                t_k += cb[y];
                y++;
            }
        }
        const uint32_t t_dist = (d1 - d0) - t_match;
        // distinct values of the whole tile
        ull r_d = t_dist, r_m = t_match, r_k = t_k;
        block_sum3(r_d, r_m, r_k, red);
        if (distinct + r_d <= S) {  // the (S+1)-th distinct value is not in this tile: take all of it
            ull r_sk = t_sk, z0 = 0, z1 = 0;
            block_sum3(r_sk, z0, z1, red);
            distinct += r_d; shared += r_m; kmers += r_k; sharedk += r_sk;
            const uint32_t xs = s_xs;
            if (xs > 0) { prevA = sa[xs - 1]; prevCA = ca[xs - 1]; }
            a0 += xs;
            b0 += tl - xs;
            continue;
        }
        // cut inside this tile: first merged element that would be distinct value number S+1
        const uint32_t need = (uint32_t)(S - distinct);  // distinct values still to take (< r_d)
        uint32_t inc = t_dist;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += u; }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < wid; w++) wbase += s_scan[w];
        const uint32_t excl = wbase + inc - t_dist;  // distinct values before this thread's run
        // the thread whose run contains the element that makes the count exceed `need`
        const bool mine = excl <= need && excl + t_dist > need;
        if (mine) s_tstar = tid;
        __syncthreads();
        const uint32_t tstar = s_tstar;
        ull f_m = 0, f_k = 0, f_sk = 0;
        if ((uint32_t)tid < tstar) { f_m = t_match; f_k = t_k; f_sk = t_sk; }
        else if ((uint32_t)tid == tstar) {  // re-walk, stopping before distinct value number need+1
            uint32_t xx = x0, yy = y0, dcount = excl;
            for (uint32_t d = d0; d < d1; d++) {
                const bool takeA = yy >= lb || (xx < la && sa[xx] <= sb[yy]);
                bool dup = false;
                if (!takeA) { const uint64_t pa = xx > 0 ? sa[xx - 1] : prevA; dup = pa == sb[yy]; }
                if (!dup && dcount == need) break;
                if (takeA) { f_k += ca[xx]; xx++; }
                else {
                    if (dup) { f_m++; const uint32_t pc = xx > 0 ? ca[xx - 1] : prevCA; f_sk += min(pc, cb[yy]); }
                    f_k += cb[yy];
                    yy++;
                }
                if (!dup) dcount++;
            }
        }
        block_sum3(f_m, f_k, f_sk, red);
        distinct = S; shared += f_m; kmers += f_k; sharedk += f_sk;
        break;
    }
    if (tid == 0) {
        quad[4 * p + 0] = shared;
        quad[4 * p + 1] = distinct;
        quad[4 * p + 2] = sharedk;
        quad[4 * p + 3] = kmers;
    }
}

__global__ void k_sketch_starts(const uint64_t *__restrict__ hashes, uint32_t n, uint32_t s, uint32_t *__restrict__ start) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t *h = hashes + (size_t)i * s;
    uint32_t lo = 0, hi = s;  // first non-zero (zeros are a prefix: sketches are ascending)
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (h[mid] == 0) lo = mid + 1; else hi = mid; }
    start[i] = lo;
}

}  // namespace

// =============================================================================================
// context
// =============================================================================================

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// bump allocator for the per-sample solid k-mer lists (kept across simka_gpu_reset)
struct Arena {
    struct Chunk { char *p; size_t cap, used; };
    std::vector<Chunk> chunks;
    void *alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (auto &c : chunks)
            if (c.cap - c.used >= bytes) { void *r = c.p + c.used; c.used += bytes; return r; }
        size_t cap = std::max<size_t>(bytes, (size_t)256 << 20);
        void *p = nullptr;
        if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
        chunks.push_back({(char *)p, cap, bytes});
        return p;
    }
    void reset() { for (auto &c : chunks) c.used = 0; }
    void release() { for (auto &c : chunks) cudaFree(c.p); chunks.clear(); }
};

struct SampleData {
    uint64_t *eh = nullptr;   // hashed solid k-mers, grouped by fine range (arena memory)
    uint32_t *ec = nullptr;
    uint64_t n = 0;
    uint32_t *off = nullptr;  // points into ctx->off_table
    bool done = false;
    bool imported = false;    // arrays owned by the caller (simka_gpu_import_sample)
    simka_sample_totals local = {0, 0, 0, 0};
};

// input of one sample in flight; two slots so that the H2D copy of sample s+1 overlaps the kernels of s
struct InSlot {
    DevBuf packed, bad;
    const uint64_t *ext_packed = nullptr;  // caller-owned device buffers (simka_gpu_push_reads_device)
    const uint32_t *ext_bad = nullptr;
    uint64_t words = 0, reads = 0;
    int sample = -1;
    cudaEvent_t copied = nullptr, consumed = nullptr;
    bool consumed_valid = false;
};

enum { T_EXTRACT_COUNT = 0, T_EXTRACT_SCATTER, T_PART_COUNT, T_PART_SCATTER, T_BIN_COUNT, T_COMPACT, T_NSTAGE };

struct simka_ctx {
    simka_params prm;
    KParams kp;
    int W = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[2][2 * T_NSTAGE + 4] = {};  // two sets: the timings of a sample are harvested while a later one runs
    bool ev_pending[2] = {false, false};
    int ev_cur = 0;
    std::string err;
    InSlot slots[2];
    DevBuf d_ascii, d_roff;
    // work buffers
    DevBuf d_keysA, d_keysB, d_stage_c;
    DevBuf d_hist1, d_base1, d_cursor1, d_tiles, d_tile_start, d_fine_hist, d_fine_off, d_fine_cursor, d_surv, d_out_off;
    DevBuf d_small;  // totals[3] + error flag + big_n
    // per-sample results
    std::vector<SampleData> samples;
    Arena arena;
    int fg_bits = -1;    // fine-range granularity, fixed by the first sample
    DevBuf d_off_table;  // [N][FG+1]
    std::vector<simka_sample_totals> global_totals;
    bool totals_overridden = false;
    // complex-distance helpers
    DevBuf d_chist, d_big_counts, d_big_sample;
    uint32_t big_cap = 1u << 20;
    // accumulators
    DevBuf d_acc;  // one slab
    bool merged = false;
    size_t acc_bytes = 0;
    ull *a_nb_shared = nullptr, *a_distinct = nullptr, *a_bc = nullptr, *a_chord = nullptr, *a_hell = nullptr, *a_whit = nullptr, *a_T = nullptr, *a_scalars = nullptr;
    double *a_kl = nullptr, *a_ntot = nullptr;
    ull *a_klfix = nullptr;
    DevBuf d_ptrs, d_logtab, d_chain, d_isum;
    bool import_pending = false;
    unsigned long long scan_epoch = 0;
    bool use_k16 = true;        // SIMKA_B200_K1=256: the 256-thread x 32-position extraction kernel
    bool force_exact = false;   // SIMKA_B200_EXACT_COUNT=1: always run the K1 count pass (tests of the fallback)
    uint64_t exact_retries = 0;
    float stage_ms[T_NSTAGE] = {};
    float sample_ms = 0, merge_ms = 0;
    uint64_t kmers_in = 0, entries = 0;
    uint64_t launches = 0;
    int nsm = 148;
    int occ_bin_count = 0;

    void set_error(const std::string &s) { err = s; }
};

static size_t merge_smem_bytes(int n) {
    return (size_t)MRG_SLOTS * 8 + (size_t)MRG_SLOTS * 4 + ((size_t)MRG_SLOTS + 1) * 4 + (size_t)MRG_CAP * 4 + ((size_t)2 * n + 1) * 4 +
           (size_t)MRG_SLOTS * 2 + (size_t)MRG_CAP * 2 + 16;
}
// the reference expression for a k-mer present in x only (src/core/SimkaAlgorithm.hpp:450-453), evaluated
// at run time so that the degenerate empty-sample case (N_y = 0 -> 0/0) yields the same NaN
static double __attribute__((noinline)) onesided_kl_term(double c, double Nx, double Ny) {
    volatile double yX = 0.0 * Nx, xY = c * Ny;
    double xi = c / Nx;
    return xi * log((2 * xY) / (xY + yX));
}
static int ilog2_ceil(uint64_t v) { int b = 0; while (((uint64_t)1 << b) < v) b++; return b; }
static int ilog2_floor(uint64_t v) { int b = 0; while (b < 63 && ((uint64_t)2 << b) <= v) b++; return b; }

// K1 launcher: picks the template instance (count-only / sharded / k >= 17 fast roll)
typedef void (*extract_fn)(const uint64_t *, const uint32_t *, uint64_t, KParams, int, uint32_t *, const uint64_t *, uint32_t *, uint64_t, int *, uint64_t *);
static extract_fn extract_instance(bool count_only, bool sharded, bool bigk) {
    static const extract_fn tab[8] = {k_extract<false, false, false>, k_extract<false, false, true>, k_extract<false, true, false>, k_extract<false, true, true>,
                                      k_extract<true, false, false>,  k_extract<true, false, true>,  k_extract<true, true, false>,  k_extract<true, true, true>};
    return tab[(count_only ? 4 : 0) | (sharded ? 2 : 0) | (bigk ? 1 : 0)];
}
static void launch_extract(simka_ctx *ctx, bool count_only, unsigned grid, const uint64_t *seq, const uint32_t *bad, uint64_t n_words, int b1_bits,
                           uint32_t *g_hist, const uint64_t *bucket_base, uint32_t *bucket_cursor, uint64_t cap, int *overflow, uint64_t *keys_out) {
    const bool bigk = ctx->kp.k >= 17 && b1_bits <= 2 * ctx->kp.k - 32;
    if (bigk && !count_only && ctx->use_k16) {
        const unsigned g16 = (unsigned)((n_words + K16_THREADS / 2 - 1) / (K16_THREADS / 2));
        const size_t sm16 = 3 * MAX_B * 4 + 128 + L1_TILE * 8;
        if (ctx->kp.shard_count > 1) k_extract16<true><<<g16, K16_THREADS, sm16, ctx->stream>>>(seq, bad, n_words, ctx->kp, b1_bits, bucket_base, bucket_cursor, cap, overflow, keys_out);
        else k_extract16<false><<<g16, K16_THREADS, sm16, ctx->stream>>>(seq, bad, n_words, ctx->kp, b1_bits, bucket_base, bucket_cursor, cap, overflow, keys_out);
        return;
    }
    const size_t smem = 3 * MAX_B * 4 + 64 + (count_only ? 0 : L1_TILE * 8);
    extract_instance(count_only, ctx->kp.shard_count > 1, bigk)<<<grid, BLOCK, smem, ctx->stream>>>(seq, bad, n_words, ctx->kp, b1_bits, g_hist, bucket_base,
                                                                                                   bucket_cursor, cap, overflow, keys_out);
}

// exclusive scan u32[n] -> u64[n+1] on the context's stream
static void launch_scan(simka_ctx *ctx, const uint32_t *in, uint64_t *out, uint64_t n) {
    if (n <= (uint64_t)SCAN_CHUNK) { k_scan_u32_u64<<<1, 1024, 0, ctx->stream>>>(in, out, n); return; }
    const unsigned grid = (unsigned)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    if (grid > 128) { k_scan_u32_u64<<<1, 1024, 0, ctx->stream>>>(in, out, n); return; }  // co-residency bound of the chained version
    ctx->scan_epoch = (ctx->scan_epoch + 1) & 0xFFFFFFull;
    if (ctx->scan_epoch == 0) { cudaMemsetAsync(ctx->d_chain.p, 0, 128 * 8, ctx->stream); ctx->scan_epoch = 1; }
    k_scan_chain<<<grid, 1024, 0, ctx->stream>>>(in, out, n, ctx->d_chain.as<unsigned long long>(), ctx->scan_epoch);
}

// adds the per-kernel device times of a finished sample (event set `set`) to the context's totals
static void harvest_timings(simka_ctx *ctx, int set) {
    if (!ctx->ev_pending[set]) return;
    cudaEvent_t *ev = ctx->ev[set];
    cudaEventSynchronize(ev[2 * T_NSTAGE + 1]);
    float ms = 0;
    for (int t = 0; t < T_NSTAGE; t++) { if (cudaEventElapsedTime(&ms, ev[2 * t], ev[2 * t + 1]) == cudaSuccess) ctx->stage_ms[t] += ms; }
    if (cudaEventElapsedTime(&ms, ev[2 * T_NSTAGE], ev[2 * T_NSTAGE + 1]) == cudaSuccess) ctx->sample_ms += ms;
    ctx->ev_pending[set] = false;
}

static int reset_state(simka_ctx *ctx) {
    harvest_timings(ctx, 0); harvest_timings(ctx, 1);
    for (auto &s : ctx->samples) s = SampleData();
    ctx->arena.reset();
    for (auto &sl : ctx->slots) { sl.sample = -1; sl.words = sl.reads = 0; sl.ext_packed = nullptr; sl.ext_bad = nullptr; }
    ctx->global_totals.assign(ctx->prm.nb_samples, simka_sample_totals{0, 0, 0, 0});
    ctx->totals_overridden = false;
    ctx->import_pending = false;
    ctx->merged = false;
    ctx->fg_bits = -1;
    for (float &f : ctx->stage_ms) f = 0;
    ctx->sample_ms = ctx->merge_ms = 0;
    ctx->kmers_in = ctx->entries = 0;
    ctx->launches = 0;
    if (ctx->prm.complex_dist) CK(cudaMemsetAsync(ctx->d_chist.p, 0, (size_t)ctx->prm.nb_samples * CHIST * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_small.p, 0, 64, ctx->stream));
    return SIMKA_OK;
}

extern "C" {

const char *simka_b200_version(void) { return "simka_b200 0.1 (sm_100a)"; }

int simka_gpu_host_alloc(void **ptr, size_t bytes) { return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? SIMKA_OK : SIMKA_ERR_NOMEM; }
int simka_gpu_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? SIMKA_OK : SIMKA_ERR_CUDA; }

int simka_gpu_create(simka_ctx **out, const simka_params *p) {
    if (!out || !p) return SIMKA_ERR_INVALID;
    *out = nullptr;
    if (p->kmer_size < 2 || p->kmer_size > 31 || p->nb_samples < 1 || p->nb_samples > MRG_MAXN) return SIMKA_ERR_INVALID;
    if (p->shard_count > 1 && (p->shard_rank < 0 || p->shard_rank >= p->shard_count)) return SIMKA_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || p->device < 0 || p->device >= ndev) return SIMKA_ERR_CUDA;
    if (cudaSetDevice(p->device) != cudaSuccess) return SIMKA_ERR_CUDA;
    simka_ctx *ctx = new simka_ctx();
    ctx->prm = *p;
    ctx->kp.mix = simka::make_mix(p->kmer_size);
    ctx->kp.k = p->kmer_size;
    ctx->kp.shard_rank = p->shard_rank;
    ctx->kp.shard_count = p->shard_count > 1 ? p->shard_count : 1;
    ctx->W = 2 * p->kmer_size;
    { const char *e = getenv("SIMKA_B200_EXACT_COUNT"); ctx->force_exact = e && e[0] == '1'; }
    { const char *e = getenv("SIMKA_B200_K1"); ctx->use_k16 = !(e && !strcmp(e, "256")); }
    cudaDeviceGetAttribute(&ctx->nsm, cudaDevAttrMultiProcessorCount, p->device);
    ctx->samples.resize(p->nb_samples);
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto &sl : ctx->slots)
        ok = ok && cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming) == cudaSuccess;
    for (auto &set : ctx->ev) for (auto &e : set) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_extract16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * MAX_B * 4 + 128 + L1_TILE * 8)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_extract16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * MAX_B * 4 + 128 + L1_TILE * 8)) == cudaSuccess;
    for (int sh = 0; sh < 2; sh++)
        for (int bk = 0; bk < 2; bk++)
            ok = ok && cudaFuncSetAttribute(extract_instance(false, sh, bk), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(3 * MAX_B * 4 + 64 + L1_TILE * 8)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem_bytes(p->nb_samples)) == cudaSuccess;
    ok = ok && ctx->d_small.reserve(64) == 0;
    ok = ok && ctx->d_chain.reserve(128 * 8) == 0 && cudaMemset(ctx->d_chain.p, 0, 128 * 8) == cudaSuccess;
    if (ok) {  // tab_log table
        double tab[128][2];
        const uint64_t OFF = 0x3fe6000000000000ull;
        for (int i = 0; i < 128; i++) {
            uint64_t ulo = OFF + ((uint64_t)i << 45), uhi = OFF + ((uint64_t)(i + 1) << 45);
            double lo, hi;
            memcpy(&lo, &ulo, 8); memcpy(&hi, &uhi, 8);
            const double c = 0.5 * (lo + hi);
            if (i == 79 || i == 80) { tab[i][0] = 1.0; tab[i][1] = 0.0; }
            else { tab[i][0] = 1.0 / c; tab[i][1] = (double)(-logl((long double)tab[i][0])); }
        }
        ok = ctx->d_logtab.reserve(sizeof tab) == 0 && cudaMemcpy(ctx->d_logtab.p, tab, sizeof tab, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (ok && p->complex_dist)
        ok = ctx->d_chist.reserve((size_t)p->nb_samples * CHIST * 8) == 0 && ctx->d_big_counts.reserve((size_t)ctx->big_cap * 4) == 0 &&
             ctx->d_big_sample.reserve((size_t)ctx->big_cap * 4) == 0;
    if (ok) ok = reset_state(ctx) == SIMKA_OK;
    if (!ok) { simka_gpu_destroy(ctx); return SIMKA_ERR_CUDA; }
    *out = ctx;
    return SIMKA_OK;
}

int simka_gpu_reset(simka_ctx *ctx) {
    if (!ctx) return SIMKA_ERR_INVALID;
    CK(cudaSetDevice(ctx->prm.device));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return reset_state(ctx);
}

void simka_gpu_destroy(simka_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->prm.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    ctx->arena.release();
    DevBuf *bufs[] = {&ctx->slots[0].packed, &ctx->slots[0].bad, &ctx->slots[1].packed, &ctx->slots[1].bad, &ctx->d_ascii, &ctx->d_roff,
                      &ctx->d_keysA, &ctx->d_keysB, &ctx->d_stage_c, &ctx->d_hist1, &ctx->d_base1, &ctx->d_cursor1, &ctx->d_tiles,
                      &ctx->d_tile_start, &ctx->d_fine_hist, &ctx->d_fine_off, &ctx->d_fine_cursor, &ctx->d_surv, &ctx->d_out_off,
                      &ctx->d_small, &ctx->d_off_table, &ctx->d_chist, &ctx->d_big_counts, &ctx->d_big_sample, &ctx->d_acc, &ctx->d_ptrs, &ctx->d_logtab, &ctx->d_chain, &ctx->d_isum};
    for (DevBuf *b : bufs) b->release();
    for (auto &set : ctx->ev) for (auto &e : set) if (e) cudaEventDestroy(e);
    for (auto &sl : ctx->slots) { if (sl.copied) cudaEventDestroy(sl.copied); if (sl.consumed) cudaEventDestroy(sl.consumed); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

const char *simka_gpu_last_error(const simka_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void *simka_gpu_stream(simka_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// grows a device buffer while keeping its first `keep` bytes
static int grow_keep(simka_ctx *ctx, DevBuf &b, size_t bytes, size_t keep) {
    if (bytes <= b.cap) return SIMKA_OK;
    void *np = nullptr;
    size_t want = bytes + bytes / 4 + 256;
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaMalloc(&np, want));
    if (keep) { CK(cudaMemcpyAsync(np, b.p, keep, cudaMemcpyDeviceToDevice, ctx->copy_stream)); CK(cudaStreamSynchronize(ctx->copy_stream)); }
    if (b.p) cudaFree(b.p);
    b.p = np; b.cap = want;
    return SIMKA_OK;
}

static int find_slot(simka_ctx *ctx, int sample_id, InSlot **out) {
    if (!ctx) return SIMKA_ERR_INVALID;
    if (sample_id < 0 || sample_id >= ctx->prm.nb_samples || ctx->samples[sample_id].done) { ctx->set_error("bad sample id"); return SIMKA_ERR_INVALID; }
    if (cudaSetDevice(ctx->prm.device) != cudaSuccess) return SIMKA_ERR_CUDA;
    for (auto &sl : ctx->slots) if (sl.sample == sample_id) { *out = &sl; return SIMKA_OK; }
    for (auto &sl : ctx->slots)
        if (sl.sample < 0) {
            sl.sample = sample_id; sl.words = sl.reads = 0; sl.ext_packed = nullptr; sl.ext_bad = nullptr;
            // the kernels of the sample that used this slot before must have consumed it
            if (sl.consumed_valid) CK(cudaStreamWaitEvent(ctx->copy_stream, sl.consumed, 0));
            *out = &sl;
            return SIMKA_OK;
        }
    ctx->set_error("at most two samples may be in flight: finish one first");
    return SIMKA_ERR_INVALID;
}

int simka_gpu_push_reads(simka_ctx *ctx, int sample_id, const uint64_t *packed, const uint32_t *bad, uint64_t n_words, uint64_t n_reads) {
    InSlot *sl = nullptr;
    int rc = find_slot(ctx, sample_id, &sl);
    if (rc) return rc;
    if (sl->ext_packed) { ctx->set_error("sample already bound to device buffers"); return SIMKA_ERR_INVALID; }
    sl->reads += n_reads;
    if (n_words == 0) return SIMKA_OK;
    const uint64_t w0 = sl->words;
    rc = grow_keep(ctx, sl->packed, (w0 + n_words + 1) * 8, w0 * 8); if (rc) return rc;
    rc = grow_keep(ctx, sl->bad, (w0 + n_words + 1) * 4, w0 * 4); if (rc) return rc;
    CK(cudaMemcpyAsync(sl->packed.as<uint64_t>() + w0, packed, (n_words + 1) * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaMemcpyAsync(sl->bad.as<uint32_t>() + w0, bad, (n_words + 1) * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(sl->copied, ctx->copy_stream));
    sl->words = w0 + n_words;
    return SIMKA_OK;
}

int simka_gpu_push_reads_device(simka_ctx *ctx, int sample_id, const uint64_t *d_packed, const uint32_t *d_bad, uint64_t n_words, uint64_t n_reads) {
    InSlot *sl = nullptr;
    int rc = find_slot(ctx, sample_id, &sl);
    if (rc) return rc;
    if (sl->words) { ctx->set_error("sample already has host-pushed reads"); return SIMKA_ERR_INVALID; }
    sl->ext_packed = d_packed; sl->ext_bad = d_bad;
    sl->words = n_words; sl->reads = n_reads;
    CK(cudaEventRecord(sl->copied, ctx->copy_stream));
    return SIMKA_OK;
}

int simka_gpu_push_ascii(simka_ctx *ctx, int sample_id, const char *seqs, const uint64_t *read_off, uint64_t n_reads) {
    InSlot *sl = nullptr;
    int rc = find_slot(ctx, sample_id, &sl);
    if (rc) return rc;
    if (sl->ext_packed) { ctx->set_error("sample already bound to device buffers"); return SIMKA_ERR_INVALID; }
    if (n_reads == 0) return SIMKA_OK;
    const uint64_t total = read_off[n_reads] - read_off[0];
    const uint64_t n_words = (total + n_reads + 31) / 32;
    const uint64_t w0 = sl->words;
    rc = grow_keep(ctx, sl->packed, (w0 + n_words + 1) * 8, w0 * 8); if (rc) return rc;
    rc = grow_keep(ctx, sl->bad, (w0 + n_words + 1) * 4, w0 * 4); if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->copy_stream));  // the ASCII staging buffers are shared between pushes
    if (ctx->d_ascii.reserve(total + 16) || ctx->d_roff.reserve((n_reads + 1) * 8)) { ctx->set_error("out of device memory"); return SIMKA_ERR_NOMEM; }
    std::vector<uint64_t> rel;  // offsets relative to the first read
    const uint64_t *ro = read_off;
    if (read_off[0] != 0) { rel.resize(n_reads + 1); for (uint64_t i = 0; i <= n_reads; i++) rel[i] = read_off[i] - read_off[0]; ro = rel.data(); }
    CK(cudaMemcpyAsync(ctx->d_ascii.p, seqs + read_off[0], total, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaMemcpyAsync(ctx->d_roff.p, ro, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (!rel.empty()) CK(cudaStreamSynchronize(ctx->copy_stream));
    const unsigned grid = (unsigned)((n_words + 1 + BLOCK - 1) / BLOCK);
    k_pack_ascii<<<grid, BLOCK, 0, ctx->copy_stream>>>(ctx->d_ascii.as<char>(), ctx->d_roff.as<uint64_t>(), n_reads, n_words,
                                                       sl->packed.as<uint64_t>() + w0, sl->bad.as<uint32_t>() + w0);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(sl->copied, ctx->copy_stream));
    sl->words = w0 + n_words;
    sl->reads += n_reads;
    return SIMKA_OK;
}

int simka_gpu_finish_sample(simka_ctx *ctx, int sample_id, simka_sample_totals *out) {
    if (!ctx) return SIMKA_ERR_INVALID;
    if (sample_id < 0 || sample_id >= ctx->prm.nb_samples || ctx->samples[sample_id].done) { ctx->set_error("bad sample id"); return SIMKA_ERR_INVALID; }
    CK(cudaSetDevice(ctx->prm.device));
    cudaStream_t st = ctx->stream;
    SampleData &sd = ctx->samples[sample_id];
    InSlot *sl = nullptr;
    for (auto &s : ctx->slots) if (s.sample == sample_id) sl = &s;
    const uint64_t n_words = sl ? sl->words : 0;
    const uint64_t n_reads = sl ? sl->reads : 0;
    const uint64_t n_pos = n_words * 32;
    const int W = ctx->W;

    // bin geometry (fixed per sample): nbins = 2^(b1+b2) fine bins of ~BIN_TARGET keys, >= fine ranges
    if (ctx->fg_bits < 0) {
        uint64_t per_sample = ctx->prm.expected_kmers_per_sample ? ctx->prm.expected_kmers_per_sample : n_pos;
        uint64_t want = std::max<uint64_t>(1, (per_sample / 4) * (uint64_t)ctx->prm.nb_samples / MRG_TARGET);
        int fg = ilog2_ceil(want);
        // never finer than the counting bins of a sample of the expected size (merge bins may be coarser)
        const int bb = std::min(ilog2_floor(std::max<uint64_t>(1, per_sample / BIN_TARGET)), 20);
        fg = std::min(fg, std::min(W, bb));
        while (fg > 0 && ((size_t)ctx->prm.nb_samples * (((size_t)1 << fg) + 1) * 4) > ((size_t)2 << 30)) fg--;
        ctx->fg_bits = fg;
        const size_t tbl = (size_t)ctx->prm.nb_samples * (((size_t)1 << fg) + 1) * 4;
        if (ctx->d_off_table.reserve(tbl)) { ctx->set_error("out of device memory (offset table)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemsetAsync(ctx->d_off_table.p, 0, tbl, st));
    }
    int bin_bits = std::max(ilog2_floor(std::max<uint64_t>(1, n_pos / BIN_TARGET)), ctx->fg_bits);
    bin_bits = std::min(bin_bits, std::min(W, 20));
    bin_bits = std::max(bin_bits, ctx->fg_bits);
    const int b1_bits = std::min(bin_bits, bin_bits > 16 ? 10 : 8);
    const int b2_bits = bin_bits - b1_bits;
    if (b2_bits > 10) { ctx->set_error("sample too large for the bin geometry"); return SIMKA_ERR_INVALID; }
    const uint32_t nb1 = 1u << b1_bits, nbins = 1u << bin_bits;
    const size_t fgn = ((size_t)1 << ctx->fg_bits) + 1;
    sd.off = ctx->d_off_table.as<uint32_t>() + (size_t)sample_id * fgn;

    ull *d_totals = ctx->d_small.as<ull>();  // [0..2]
    int *d_error = reinterpret_cast<int *>(ctx->d_small.as<char>() + 32);
    uint32_t *d_big_n = reinterpret_cast<uint32_t *>(ctx->d_small.as<char>() + 40);
    int *d_overflow = reinterpret_cast<int *>(ctx->d_small.as<char>() + 44);
    uint64_t h_keys = 0, h_surv = 0;
    ull h_tot[3] = {0, 0, 0};

    if (n_words > 0) {
        bool okm = !ctx->d_keysA.reserve(n_pos * 8) && !ctx->d_keysB.reserve(n_pos * 8) && !ctx->d_stage_c.reserve(n_pos * 4) &&
                   !ctx->d_hist1.reserve(MAX_B * 4) && !ctx->d_base1.reserve((MAX_B + 1) * 8) && !ctx->d_cursor1.reserve(MAX_B * 4) &&
                   !ctx->d_tiles.reserve(MAX_B * 4) && !ctx->d_tile_start.reserve((MAX_B + 1) * 8) &&
                   !ctx->d_fine_hist.reserve((size_t)nbins * 4) && !ctx->d_fine_off.reserve(((size_t)nbins + 1) * 8) &&
                   !ctx->d_fine_cursor.reserve((size_t)nbins * 4) && !ctx->d_surv.reserve((size_t)nbins * 4) &&
                   !ctx->d_out_off.reserve(((size_t)nbins + 1) * 8);
        if (!okm) { ctx->set_error("out of device memory (work buffers)"); return SIMKA_ERR_NOMEM; }
        CK(cudaStreamWaitEvent(st, sl->copied, 0));
        const int evset = ctx->ev_cur;
        harvest_timings(ctx, evset);  // the sample before the previous one: long finished
        cudaEvent_t *ev = ctx->ev[evset];
        CK(cudaEventRecord(ev[2 * T_NSTAGE], st));  // sample start
      for (int attempt = 0; attempt < 2; attempt++) {
        // attempt 0: single pass over the reads into fixed-capacity L1 buckets (sizes are binomial for hashed
        // keys); attempt 1 (only after an overflow, e.g. one k-mer repeated millions of times): exact count pass first
        const bool exact = attempt == 1 || ctx->force_exact;
        const uint64_t per_bucket = n_pos / nb1;
        const uint64_t cap = per_bucket + 8 * (uint64_t)sqrt((double)per_bucket) + 1024;
        if (!exact && ctx->d_keysA.reserve((size_t)nb1 * cap * 8)) { ctx->set_error("out of device memory (work buffers)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemsetAsync(d_totals, 0, 24, st));
        CK(cudaMemsetAsync(d_overflow, 0, 4, st));
        CK(cudaMemsetAsync(ctx->d_hist1.p, 0, MAX_B * 4, st));
        CK(cudaMemsetAsync(ctx->d_cursor1.p, 0, MAX_B * 4, st));
        CK(cudaMemsetAsync(ctx->d_fine_hist.p, 0, (size_t)nbins * 4, st));
        CK(cudaMemsetAsync(ctx->d_fine_cursor.p, 0, (size_t)nbins * 4, st));

        const unsigned g1 = (unsigned)((n_words + BLOCK - 1) / BLOCK);
        const uint64_t *seq = sl->ext_packed ? sl->ext_packed : sl->packed.as<uint64_t>();
        const uint32_t *bad = sl->ext_bad ? sl->ext_bad : sl->bad.as<uint32_t>();
        // K1 (count / scan /) scatter
        CK(cudaEventRecord(ev[2 * T_EXTRACT_COUNT], st));
        if (exact) {
            launch_extract(ctx, true, g1, seq, bad, n_words, b1_bits, ctx->d_hist1.as<uint32_t>(), nullptr, nullptr, 0, nullptr, nullptr);
            ctx->launches++;
        }
        CK(cudaEventRecord(ev[2 * T_EXTRACT_COUNT + 1], st));
        if (exact) launch_scan(ctx, ctx->d_hist1.as<uint32_t>(), ctx->d_base1.as<uint64_t>(), nb1);
        else k_fill_bases<<<(nb1 + 256) / 256, 256, 0, st>>>(ctx->d_base1.as<uint64_t>(), (int)nb1, cap);
        CK(cudaEventRecord(ev[2 * T_EXTRACT_SCATTER], st));
        launch_extract(ctx, false, g1, seq, bad, n_words, b1_bits, nullptr, ctx->d_base1.as<uint64_t>(), ctx->d_cursor1.as<uint32_t>(),
                       exact ? ~0ull >> 1 : cap, d_overflow, ctx->d_keysA.as<uint64_t>());
        CK(cudaEventRecord(ev[2 * T_EXTRACT_SCATTER + 1], st));
        // per-bucket key counts: the cursors of the scatter pass (== the histogram in exact mode)
        const uint32_t *cnt1 = ctx->d_cursor1.as<uint32_t>();
        // K2 tiles, count / scan / scatter
        k_tiles_per_bucket<<<(nb1 + 255) / 256, 256, 0, st>>>(cnt1, ctx->d_tiles.as<uint32_t>(), nb1, L2_TILE);
        launch_scan(ctx, ctx->d_tiles.as<uint32_t>(), ctx->d_tile_start.as<uint64_t>(), nb1);
        const unsigned g2 = (unsigned)(n_pos / L2_TILE + nb1 + 1);  // upper bound on the number of tiles
        CK(cudaEventRecord(ev[2 * T_PART_COUNT], st));
        k_partition2<true><<<g2, BLOCK, 0, st>>>(ctx->d_keysA.as<uint64_t>(), ctx->d_base1.as<uint64_t>(), cnt1, ctx->d_tile_start.as<uint64_t>(), b1_bits, b2_bits, W,
                                                 ctx->d_fine_hist.as<uint32_t>(), nullptr, nullptr, nullptr);
        CK(cudaEventRecord(ev[2 * T_PART_COUNT + 1], st));
        launch_scan(ctx, ctx->d_fine_hist.as<uint32_t>(), ctx->d_fine_off.as<uint64_t>(), nbins);
        CK(cudaEventRecord(ev[2 * T_PART_SCATTER], st));
        k_partition2<false><<<g2, BLOCK, 0, st>>>(ctx->d_keysA.as<uint64_t>(), ctx->d_base1.as<uint64_t>(), cnt1, ctx->d_tile_start.as<uint64_t>(), b1_bits, b2_bits, W,
                                                  nullptr, ctx->d_fine_off.as<uint64_t>(), ctx->d_fine_cursor.as<uint32_t>(), ctx->d_keysB.as<uint64_t>());
        CK(cudaEventRecord(ev[2 * T_PART_SCATTER + 1], st));
        // K3 count
        CountOut co;
        co.stage_h = ctx->d_keysA.as<uint64_t>();
        co.stage_c = ctx->d_stage_c.as<uint32_t>();
        co.surv = ctx->d_surv.as<uint32_t>();
        co.totals = d_totals;
        co.error_flag = d_error;
        if (ctx->occ_bin_count <= 0) {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_bin_count, k_bin_count, BLOCK, 0));
            if (ctx->occ_bin_count <= 0) ctx->occ_bin_count = 1;
        }
        const unsigned gbin = (unsigned)std::min<uint64_t>(nbins, (uint64_t)ctx->nsm * ctx->occ_bin_count);
        const int slot_shift = std::max(0, W - bin_bits - 11);
        CK(cudaEventRecord(ev[2 * T_BIN_COUNT], st));
        k_bin_count<<<gbin, BLOCK, 0, st>>>(ctx->d_keysB.as<uint64_t>(), ctx->d_fine_off.as<uint64_t>(), nbins, slot_shift, ctx->prm.abundance_min,
                                            ctx->prm.abundance_max, co);
        CK(cudaEventRecord(ev[2 * T_BIN_COUNT + 1], st));
        launch_scan(ctx, ctx->d_surv.as<uint32_t>(), ctx->d_out_off.as<uint64_t>(), nbins);
        ctx->launches += 9;
        CK(cudaGetLastError());
        // the only host sync of the sample: sizes + totals
        int h_err = 0, h_overflow = 0;
        CK(cudaMemcpyAsync(&h_surv, ctx->d_out_off.as<uint64_t>() + nbins, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_keys, ctx->d_fine_off.as<uint64_t>() + nbins, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_tot, d_totals, 24, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_err, d_error, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_overflow, d_overflow, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_overflow && !exact) { ctx->exact_retries++; continue; }  // a bucket outgrew its capacity: redo with the exact two-pass path
        if (h_err) { ctx->set_error("on-chip table overflow in k_bin_count (code " + std::to_string(h_err) + ")"); return SIMKA_ERR_OVERFLOW; }
        break;
      }
        CK(cudaEventRecord(sl->consumed, st));  // the input slot may be overwritten from here on
        sl->consumed_valid = true;
        if (h_surv >= (1ull << 32)) { ctx->set_error("more than 2^32 solid k-mers in one sample shard"); return SIMKA_ERR_INVALID; }
        CK(cudaEventRecord(ev[2 * T_COMPACT], st));
        if (h_surv > 0) {
            sd.eh = (uint64_t *)ctx->arena.alloc(h_surv * 8);
            sd.ec = (uint32_t *)ctx->arena.alloc(h_surv * 4);
            if (!sd.eh || !sd.ec) { ctx->set_error("out of device memory (sample list)"); return SIMKA_ERR_NOMEM; }
            k_compact<<<(unsigned)std::min<uint64_t>((nbins + 7) / 8, (uint64_t)ctx->nsm * 8), BLOCK, 0, st>>>(ctx->d_keysA.as<uint64_t>(), ctx->d_stage_c.as<uint32_t>(), ctx->d_fine_off.as<uint64_t>(),
                                             ctx->d_out_off.as<uint64_t>(), nbins, sd.eh, sd.ec);
            ctx->launches++;
            if (ctx->prm.complex_dist) {
                const unsigned gh = (unsigned)std::min<uint64_t>((h_surv + BLOCK * 16 - 1) / (BLOCK * 16), (uint64_t)ctx->nsm * 4);
                k_abundance_hist<<<gh, BLOCK, 0, st>>>(sd.ec, h_surv, sample_id, ctx->d_chist.as<ull>() + (size_t)sample_id * CHIST,
                                                      ctx->d_big_counts.as<uint32_t>(), ctx->d_big_sample.as<uint32_t>(), d_big_n, ctx->big_cap, d_error);
                ctx->launches++;
            }
        }
        const uint32_t fg = 1u << ctx->fg_bits;
        k_fill_offsets<<<(fg + 1 + 255) / 256, 256, 0, st>>>(ctx->d_out_off.as<uint64_t>(), sd.off, fg, nbins >> ctx->fg_bits);
        ctx->launches++;
        CK(cudaEventRecord(ev[2 * T_COMPACT + 1], st));
        CK(cudaEventRecord(ev[2 * T_NSTAGE + 1], st));  // sample end; no host wait: the next sample's kernels queue right behind
        ctx->ev_pending[evset] = true;
        ctx->ev_cur ^= 1;
    }
    sd.n = h_surv;
    sd.done = true;
    sd.local.nb_reads = n_reads;
    sd.local.nb_distinct = h_tot[0];
    sd.local.nb_kmers = h_tot[1];
    sd.local.chord_n2 = h_tot[2];
    if (!ctx->totals_overridden) ctx->global_totals[sample_id] = sd.local;
    ctx->kmers_in += h_keys;
    ctx->entries += h_surv;
    if (sl) { sl->sample = -1; sl->words = sl->reads = 0; sl->ext_packed = nullptr; sl->ext_bad = nullptr; }
    if (out) *out = sd.local;
    return SIMKA_OK;
}

int simka_gpu_dump_sample_counts(simka_ctx *ctx, int sample_id, uint64_t *kmers, uint32_t *counts, uint64_t *n) {
    if (!ctx || !n || sample_id < 0 || sample_id >= ctx->prm.nb_samples || !ctx->samples[sample_id].done) return SIMKA_ERR_INVALID;
    CK(cudaSetDevice(ctx->prm.device));
    SampleData &sd = ctx->samples[sample_id];
    if (!kmers || !counts) { *n = sd.n; return SIMKA_OK; }
    if (*n < sd.n) { *n = sd.n; return SIMKA_ERR_INVALID; }
    *n = sd.n;
    if (sd.n == 0) return SIMKA_OK;
    uint64_t *tmp = nullptr;
    CK(cudaMalloc(&tmp, sd.n * 8));
    k_unmix<<<(unsigned)((sd.n + 255) / 256), 256, 0, ctx->stream>>>(sd.eh, tmp, sd.n, ctx->kp.mix);
    std::vector<uint64_t> hk(sd.n);
    std::vector<uint32_t> hc(sd.n);
    CK(cudaMemcpyAsync(hk.data(), tmp, sd.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hc.data(), sd.ec, sd.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    std::vector<uint64_t> order(sd.n);
    for (uint64_t i = 0; i < sd.n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return hk[a] < hk[b]; });
    for (uint64_t i = 0; i < sd.n; i++) { kmers[i] = hk[order[i]]; counts[i] = hc[order[i]]; }
    return SIMKA_OK;
}

int simka_gpu_fine_bits(simka_ctx *ctx) { return ctx ? ctx->fg_bits : -1; }

int simka_gpu_export_sample(simka_ctx *ctx, int sample_id, const uint64_t **d_kmers, const uint32_t **d_counts, const uint32_t **d_off, uint64_t *n) {
    if (!ctx || sample_id < 0 || sample_id >= ctx->prm.nb_samples || !ctx->samples[sample_id].done || ctx->fg_bits < 0) return SIMKA_ERR_INVALID;
    const SampleData &sd = ctx->samples[sample_id];
    CK(cudaSetDevice(ctx->prm.device));
    CK(cudaStreamSynchronize(ctx->stream));  // the caller reads the arrays from its own streams
    if (d_kmers) *d_kmers = sd.eh;
    if (d_counts) *d_counts = sd.ec;
    if (d_off) *d_off = sd.off ? sd.off : ctx->d_off_table.as<uint32_t>() + (size_t)sample_id * (((size_t)1 << ctx->fg_bits) + 1);
    if (n) *n = sd.n;
    return SIMKA_OK;
}

int simka_gpu_import_sample(simka_ctx *ctx, int sample_id, const uint64_t *d_kmers, const uint32_t *d_counts, uint64_t n, const uint32_t *d_off,
                            int fine_bits, uint64_t nb_reads) {
    if (!ctx) return SIMKA_ERR_INVALID;
    if (sample_id < 0 || sample_id >= ctx->prm.nb_samples || ctx->samples[sample_id].done || fine_bits < 0 || fine_bits > 24 || !d_off ||
        (n && (!d_kmers || !d_counts)) || n >= (1ull << 32)) { ctx->set_error("bad import arguments"); return SIMKA_ERR_INVALID; }
    CK(cudaSetDevice(ctx->prm.device));
    cudaStream_t st = ctx->stream;
    if (ctx->fg_bits < 0) {
        ctx->fg_bits = fine_bits;
        const size_t tbl = (size_t)ctx->prm.nb_samples * (((size_t)1 << fine_bits) + 1) * 4;
        if (ctx->d_off_table.reserve(tbl)) { ctx->set_error("out of device memory (offset table)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemsetAsync(ctx->d_off_table.p, 0, tbl, st));
    } else if (ctx->fg_bits != fine_bits) { ctx->set_error("fine_bits differs from the context's"); return SIMKA_ERR_INVALID; }
    SampleData &sd = ctx->samples[sample_id];
    const size_t fgn = ((size_t)1 << fine_bits) + 1;
    sd.off = ctx->d_off_table.as<uint32_t>() + (size_t)sample_id * fgn;
    CK(cudaMemcpyAsync(sd.off, d_off, fgn * 4, cudaMemcpyDeviceToDevice, st));
    sd.eh = const_cast<uint64_t *>(d_kmers);
    sd.ec = const_cast<uint32_t *>(d_counts);
    sd.n = n;
    int *d_error = reinterpret_cast<int *>(ctx->d_small.as<char>() + 32);
    uint32_t *d_big_n = reinterpret_cast<uint32_t *>(ctx->d_small.as<char>() + 40);
    // the sum of the imported abundances is needed at read-out only: computed asynchronously into d_isum[sample]
    // and fetched once by simka_gpu_merge_and_accumulate (no host synchronisation per imported sample)
    if (ctx->d_isum.reserve((size_t)ctx->prm.nb_samples * 8)) { ctx->set_error("out of device memory"); return SIMKA_ERR_NOMEM; }
    if (!ctx->import_pending) { CK(cudaMemsetAsync(ctx->d_isum.p, 0, (size_t)ctx->prm.nb_samples * 8, st)); ctx->import_pending = true; }
    if (n) {
        const unsigned g = (unsigned)std::min<uint64_t>((n + BLOCK * 16 - 1) / (BLOCK * 16), (uint64_t)ctx->nsm * 4);
        k_sum_u32<<<g, BLOCK, 0, st>>>(d_counts, n, ctx->d_isum.as<ull>() + sample_id);
        if (ctx->prm.complex_dist)
            k_abundance_hist<<<g, BLOCK, 0, st>>>(d_counts, n, sample_id, ctx->d_chist.as<ull>() + (size_t)sample_id * CHIST,
                                                 ctx->d_big_counts.as<uint32_t>(), ctx->d_big_sample.as<uint32_t>(), d_big_n, ctx->big_cap, d_error);
        ctx->launches += ctx->prm.complex_dist ? 2 : 1;
        CK(cudaGetLastError());
    }
    sd.done = true;
    sd.imported = true;
    sd.local.nb_reads = nb_reads;
    sd.local.nb_distinct = n;
    sd.local.nb_kmers = 0;  // filled from d_isum at merge time
    sd.local.chord_n2 = 0;
    if (!ctx->totals_overridden) ctx->global_totals[sample_id] = sd.local;
    ctx->entries += n;
    return SIMKA_OK;
}

int simka_gpu_get_totals(simka_ctx *ctx, simka_sample_totals *out) {
    if (!ctx || !out) return SIMKA_ERR_INVALID;
    for (int s = 0; s < ctx->prm.nb_samples; s++) out[s] = ctx->samples[s].local;
    return SIMKA_OK;
}

int simka_gpu_set_totals(simka_ctx *ctx, const simka_sample_totals *in) {
    if (!ctx || !in) return SIMKA_ERR_INVALID;
    for (int s = 0; s < ctx->prm.nb_samples; s++) ctx->global_totals[s] = in[s];
    ctx->totals_overridden = true;
    return SIMKA_OK;
}

int simka_gpu_merge_and_accumulate(simka_ctx *ctx) {
    if (!ctx) return SIMKA_ERR_INVALID;
    const int N = ctx->prm.nb_samples;
    for (int s = 0; s < N; s++) if (!ctx->samples[s].done) { ctx->set_error("all samples must be finished before the merge"); return SIMKA_ERR_INVALID; }
    if (ctx->merged) { ctx->set_error("merge already done"); return SIMKA_ERR_INVALID; }
    CK(cudaSetDevice(ctx->prm.device));
    cudaStream_t st = ctx->stream;
    if (ctx->import_pending) {  // abundance sums of the imported sample slices
        std::vector<ull> sums(N);
        CK(cudaMemcpyAsync(sums.data(), ctx->d_isum.p, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int s = 0; s < N; s++)
            if (ctx->samples[s].imported) {
                ctx->samples[s].local.nb_kmers = sums[s];
                if (!ctx->totals_overridden) ctx->global_totals[s].nb_kmers = sums[s];
            }
        ctx->import_pending = false;
    }
    const size_t nn = (size_t)N * N, ns = (size_t)N * (N + 1) / 2;
    // slab: nb_shared nn | distinct ns | bc ns | chord nn | hell nn | whit nn | T nn | scalars 4 | kl nn (f64) | klfix nn (i64, 2^-60 units) | ntot N (f64)
    const size_t words = nn + ns + ns + nn + nn + nn + nn + 4 + nn + nn + N;
    ctx->acc_bytes = words * 8;
    if (ctx->d_acc.reserve(ctx->acc_bytes)) { ctx->set_error("out of device memory (accumulators)"); return SIMKA_ERR_NOMEM; }
    harvest_timings(ctx, 0); harvest_timings(ctx, 1);
    cudaEvent_t e0 = ctx->ev[0][2 * T_NSTAGE + 2], e1 = ctx->ev[0][2 * T_NSTAGE + 3];
    CK(cudaEventRecord(e0, st));
    CK(cudaMemsetAsync(ctx->d_acc.p, 0, ctx->acc_bytes, st));
    ull *p = ctx->d_acc.as<ull>();
    ctx->a_nb_shared = p; p += nn;
    ctx->a_distinct = p; p += ns;
    ctx->a_bc = p; p += ns;
    ctx->a_chord = p; p += nn;
    ctx->a_hell = p; p += nn;
    ctx->a_whit = p; p += nn;
    ctx->a_T = p; p += nn;
    ctx->a_scalars = p; p += 4;
    ctx->a_kl = reinterpret_cast<double *>(p); p += nn;
    ctx->a_klfix = p; p += nn;
    ctx->a_ntot = reinterpret_cast<double *>(p);
    std::vector<double> ntot(N);
    uint64_t total_entries = 0;
    for (int s = 0; s < N; s++) { ntot[s] = (double)ctx->global_totals[s].nb_kmers; total_entries += ctx->samples[s].n; }
    CK(cudaMemcpyAsync(ctx->a_ntot, ntot.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
    // pointer tables
    std::vector<const void *> ptrs(2 * (size_t)N);
    for (int s = 0; s < N; s++) { ptrs[s] = ctx->samples[s].eh; ptrs[N + s] = ctx->samples[s].ec; }
    if (ctx->d_ptrs.reserve(ptrs.size() * sizeof(void *))) { ctx->set_error("out of device memory"); return SIMKA_ERR_NOMEM; }
    CK(cudaMemcpyAsync(ctx->d_ptrs.p, ptrs.data(), ptrs.size() * sizeof(void *), cudaMemcpyHostToDevice, st));

    int *d_error = reinterpret_cast<int *>(ctx->d_small.as<char>() + 32);
    uint32_t *d_big_n = reinterpret_cast<uint32_t *>(ctx->d_small.as<char>() + 40);
    uint32_t big_n = 0;
    if (ctx->prm.complex_dist) CK(cudaMemcpyAsync(&big_n, d_big_n, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));  // host staging vectors are reused below / big_n needed
    if (ctx->fg_bits >= 0 && total_entries > 0) {
        int mg_bits = std::min(ctx->fg_bits, ilog2_ceil(std::max<uint64_t>(1, total_entries / MRG_TARGET)));
        MergeArgs a;
        a.n = N; a.fg_bits = ctx->fg_bits; a.mg_bits = mg_bits;
        a.eh = reinterpret_cast<const uint64_t *const *>(ctx->d_ptrs.p);
        a.ec = reinterpret_cast<const uint32_t *const *>(ctx->d_ptrs.as<const void *>() + N);
        a.off = ctx->d_off_table.as<uint32_t>();
        a.ntot = ctx->a_ntot;
        a.logtab = ctx->d_logtab.as<double2>();
        a.simple = ctx->prm.simple_dist; a.complex_ = ctx->prm.complex_dist; a.count_only = 0;
        a.nb_shared = ctx->a_nb_shared; a.distinct_shared = ctx->a_distinct; a.braycurtis = ctx->a_bc;
        a.chord = ctx->a_chord; a.hellinger = ctx->a_hell; a.whittaker = ctx->a_whit; a.kl = ctx->a_kl; a.klfix = ctx->a_klfix;
        a.scalars = ctx->a_scalars; a.error_flag = d_error;
        const int nsm = ctx->nsm;
        const uint32_t MG = 1u << mg_bits;
        const unsigned grid = (unsigned)std::min<uint64_t>(MG, (uint64_t)nsm * 8);
        // default: k_merge_sm for N <= 64, k_merge (global atomics) above.  SIMKA_B200_MERGE selects the measured-and-kept
        // alternatives: "sorted" = k_merge32 (N <= 32), "generic" = k_merge for any N, "tiled" = 32 x 32 sample tiles on k_merge_sm (N > 64)
        const char *mv = getenv("SIMKA_B200_MERGE");
        const bool want_generic = mv && !strcmp(mv, "generic"), want_sorted = mv && !strcmp(mv, "sorted"), want_tiled = mv && !strcmp(mv, "tiled");
        if (N <= 32 && want_sorted) {
            void (*kern)(MergeArgs) = a.simple ? (a.complex_ ? k_merge32<true, true> : k_merge32<true, false>)
                                               : (a.complex_ ? k_merge32<false, true> : k_merge32<false, false>);
            int occ = 0;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(M32Smem)));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, M32_THREADS, sizeof(M32Smem)));
            const unsigned g32 = (unsigned)std::min<uint64_t>(MG, (uint64_t)nsm * std::max(1, occ));
            kern<<<g32, M32_THREADS, sizeof(M32Smem), st>>>(a);
        } else if (N >= 2 && N <= MS_MAXN && !want_generic) {
            void (*kern)(MergeArgs) = a.simple ? (a.complex_ ? k_merge_sm<true, true, false> : k_merge_sm<true, false, false>)
                                               : (a.complex_ ? k_merge_sm<false, true, false> : k_merge_sm<false, false, false>);
            int occ = 0;
            const size_t smem = ms_smem_bytes(N);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MS_THREADS, smem));
            const unsigned gs = (unsigned)std::min<uint64_t>(MG, (uint64_t)nsm * std::max(1, occ));
            kern<<<gs, MS_THREADS, smem, st>>>(a);
        } else if (N > MS_MAXN && want_tiled) {
            // distinct / shared k-mer counts over all samples: grouping only
            a.count_only = 1;
            k_merge<<<grid, BLOCK, merge_smem_bytes(N), st>>>(a);
            a.count_only = 0;
            ctx->launches++;
            // the pairs: tiles of 32 x 32 samples; merge bins sized for the <= 64 samples a tile gathers
            MergeArgs t = a;
            t.mg_bits = std::min(ctx->fg_bits, ilog2_ceil(std::max<uint64_t>(1, total_entries * 64 / (uint64_t)N / 1024)));
            void (*kern)(MergeArgs) = a.simple ? (a.complex_ ? k_merge_sm<true, true, true> : k_merge_sm<true, false, true>)
                                               : (a.complex_ ? k_merge_sm<false, true, true> : k_merge_sm<false, false, true>);
            int occ = 0;
            const size_t smem = ms_smem_bytes(N);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MS_THREADS, smem));
            kern<<<(unsigned)((uint64_t)nsm * std::max(1, occ)), MS_THREADS, smem, st>>>(t);
        } else {
            k_merge<<<grid, BLOCK, merge_smem_bytes(N), st>>>(a);
        }
        ctx->launches++;
        if (ctx->prm.complex_dist) {
            big_n = std::min(big_n, ctx->big_cap);
            k_whittaker_onesided<<<N, BLOCK, 0, st>>>(ctx->d_chist.as<ull>(), ctx->d_big_counts.as<uint32_t>(), ctx->d_big_sample.as<uint32_t>(), big_n,
                                                     ctx->a_ntot, N, ctx->a_T);
            k_whittaker_fold<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(ctx->a_whit, ctx->a_T, N);
            ctx->launches += 2;
        }
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(e1, st));
    int h_err = 0;
    CK(cudaMemcpyAsync(&h_err, d_error, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_err) { ctx->set_error("on-chip table overflow in k_merge (code " + std::to_string(h_err) + ")"); return SIMKA_ERR_OVERFLOW; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ctx->merge_ms = ms;
    ctx->merged = true;
    return SIMKA_OK;
}

int simka_gpu_get_stats(simka_ctx *ctx, simka_stats_view *out) {
    if (!ctx || !out || !ctx->merged) { if (ctx) ctx->set_error("merge_and_accumulate has not run"); return SIMKA_ERR_INVALID; }
    CK(cudaSetDevice(ctx->prm.device));
    const int N = ctx->prm.nb_samples;
    const size_t nn = (size_t)N * N, ns = (size_t)N * (N + 1) / 2;
    std::vector<ull> h(ctx->acc_bytes / 8);
    CK(cudaMemcpyAsync(h.data(), ctx->d_acc.p, ctx->acc_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const ull *nb_shared = h.data();
    const ull *distinct = nb_shared + nn, *bc = distinct + ns, *chord = bc + ns, *hell = chord + nn, *whit = hell + nn;
    const ull *scal = whit + nn + nn;
    const double *kl = reinterpret_cast<const double *>(scal + 4);
    const long long *klfix = reinterpret_cast<const long long *>(scal + 4 + nn);
    out->nb_samples = N; out->simple_dist = ctx->prm.simple_dist; out->complex_dist = ctx->prm.complex_dist;
    out->nb_distinct_kmers = scal[0];
    out->nb_shared_kmers = scal[1];
    if (out->totals) for (int s = 0; s < N; s++) out->totals[s] = ctx->global_totals[s];
    if (out->distinct_shared) memcpy(out->distinct_shared, distinct, ns * 8);
    if (out->braycurtis_num) memcpy(out->braycurtis_num, bc, ns * 8);
    if (out->nb_shared) memcpy(out->nb_shared, nb_shared, nn * 8);
    if (ctx->prm.simple_dist) {
        if (out->chord_ninj) memcpy(out->chord_ninj, chord, nn * 8);
        if (out->hellinger) memcpy(out->hellinger, hell, nn * 8);
        if (out->kulczynski) {  // same increments as the Bray-Curtis numerator, stored in the upper triangle
            memset(out->kulczynski, 0, nn * 8);
            for (int i = 0; i < N; i++)
                for (int j = i + 1; j < N; j++) out->kulczynski[(size_t)i * N + j] = bc[(size_t)j + (size_t)(N - 1) * i - ((size_t)i * (i - 1) / 2)];
        }
    }
    if (ctx->prm.complex_dist) {
        // one-sided closed forms use this shard's own counts so that shards still sum to the global value
        if (out->whittaker) memcpy(out->whittaker, whit, nn * 8);
        if (out->canberra) {
            memset(out->canberra, 0, nn * 8);
            for (int i = 0; i < N; i++)
                for (int j = i + 1; j < N; j++) {
                    const ull a = distinct[(size_t)j + (size_t)(N - 1) * i - ((size_t)i * (i - 1) / 2)];
                    out->canberra[(size_t)i * N + j] = ctx->samples[i].local.nb_distinct + ctx->samples[j].local.nb_distinct - 2 * a;
                }
        }
        if (out->kullback_leibler) {
            const double ln2 = log(2.0);
            for (size_t q = 0; q < nn; q++) out->kullback_leibler[q] = 0;
            for (int i = 0; i < N; i++)
                for (int j = i + 1; j < N; j++) {
                    const size_t ij = (size_t)i * N + j, ji = (size_t)j * N + i;
                    double v = kl[ij] + ldexp((double)klfix[ij], -60);
                    const double Ni = (double)ctx->global_totals[i].nb_kmers, Nj = (double)ctx->global_totals[j].nb_kmers;
                    // k-mers present in exactly one of the two samples: (c/N_x) * log(2 c N_y / (c N_y + 0)) each
                    const ull only_i = ctx->samples[i].local.nb_kmers - nb_shared[ij];
                    const ull only_j = ctx->samples[j].local.nb_kmers - nb_shared[ji];
                    if (only_i) v += (Nj > 0) ? ln2 * ((double)only_i / Ni) : onesided_kl_term(1.0, Ni, Nj);
                    if (only_j) v += (Ni > 0) ? ln2 * ((double)only_j / Nj) : onesided_kl_term(1.0, Nj, Ni);
                    out->kullback_leibler[ij] = v;
                }
        }
    }
    return SIMKA_OK;
}

int simka_gpu_get_timings(simka_ctx *ctx, simka_timings *out) {
    if (!ctx || !out) return SIMKA_ERR_INVALID;
    harvest_timings(ctx, 0); harvest_timings(ctx, 1);
    simka_timings t;
    memset(&t, 0, sizeof t);
    t.extract_count_ms = ctx->stage_ms[T_EXTRACT_COUNT];
    t.extract_scatter_ms = ctx->stage_ms[T_EXTRACT_SCATTER];
    t.part_count_ms = ctx->stage_ms[T_PART_COUNT];
    t.part_scatter_ms = ctx->stage_ms[T_PART_SCATTER];
    t.bin_count_ms = ctx->stage_ms[T_BIN_COUNT];
    t.compact_ms = ctx->stage_ms[T_COMPACT];
    t.merge_ms = ctx->merge_ms;
    t.samples_ms = ctx->sample_ms;
    t.total_ms = ctx->sample_ms + ctx->merge_ms;
    t.kmers_in = ctx->kmers_in;
    t.entries = ctx->entries;
    t.kernel_launches = ctx->launches;
    t.exact_retries = ctx->exact_retries;
    if (ctx->merged && ctx->a_scalars) {
        ull sc[4];
        if (cudaMemcpy(sc, ctx->a_scalars, 32, cudaMemcpyDeviceToHost) == cudaSuccess) { t.distinct_kmers = sc[0]; t.pairs = sc[2]; }
    }
    *out = t;
    return SIMKA_OK;
}

int simkamin_gpu_distance(int device, const uint64_t *hashes, const uint32_t *counts, uint32_t n, uint32_t s, float *jaccard,
                          float *braycurtis, float *kernel_ms) {
    if (!hashes || !counts || !jaccard || !braycurtis || n == 0 || s == 0) return SIMKA_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return SIMKA_ERR_CUDA;
    uint64_t *dh = nullptr; uint32_t *dc = nullptr, *dstart = nullptr; ull *dq = nullptr;
    const size_t ne = (size_t)n * s, nn = (size_t)n * n;
    const uint64_t pairs = (uint64_t)n * (n - 1) / 2;
    int rc = SIMKA_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<ull> hq(4 * pairs);
    do {
        if (cudaMalloc(&dh, ne * 8) || cudaMalloc(&dc, ne * 4) || cudaMalloc(&dstart, (size_t)n * 4) || cudaMalloc(&dq, (pairs + 1) * 32)) { rc = SIMKA_ERR_NOMEM; break; }
        if (cudaMemcpy(dh, hashes, ne * 8, cudaMemcpyHostToDevice) || cudaMemcpy(dc, counts, ne * 4, cudaMemcpyHostToDevice)) { rc = SIMKA_ERR_CUDA; break; }
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_sketch_starts<<<(n + 255) / 256, 256>>>(dh, n, s, dstart);
        {   // SIMKA_B200_SIMKAMIN=chunked selects the first (latency-bound) kernel, kept for comparison
            const char *v = getenv("SIMKA_B200_SIMKAMIN");
            if (pairs && v && !strcmp(v, "chunked")) k_simkamin_pair<<<(unsigned)pairs, BLOCK>>>(dh, dc, dstart, n, s, dq);
            else if (pairs) k_simkamin_tiled<<<(unsigned)pairs, BLOCK>>>(dh, dc, dstart, n, s, dq);
        }
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = SIMKA_ERR_CUDA; break; }
        if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
        if (pairs && cudaMemcpy(hq.data(), dq, pairs * 32, cudaMemcpyDeviceToHost)) { rc = SIMKA_ERR_CUDA; break; }
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(dh); cudaFree(dc); cudaFree(dstart); cudaFree(dq);
    if (rc) return rc;
    // final ratios exactly as the reference: long double division, stored as float (:244-259)
    for (size_t q = 0; q < nn; q++) { jaccard[q] = 0; braycurtis[q] = 0; }
    uint64_t p = 0;
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t j = i + 1; j < n; j++, p++) {
            const ull sh = hq[4 * p], di = hq[4 * p + 1], sk = hq[4 * p + 2], km = hq[4 * p + 3];
            float ja = di == 0 ? 1.0f : (float)(1 - (long double)sh / (long double)di);
            float b = km == 0 ? 1.0f : (float)(1 - (long double)(2 * sk) / (long double)km);
            jaccard[(size_t)i * n + j] = jaccard[(size_t)j * n + i] = ja;
            braycurtis[(size_t)i * n + j] = braycurtis[(size_t)j * n + i] = b;
        }
    return rc;
}

}  // extern "C"
