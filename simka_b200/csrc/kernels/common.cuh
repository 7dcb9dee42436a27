// kernels/common.cuh -- constants, block scans, device-wide exclusive scans
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
#pragma once


constexpr int BLOCK = 256;
constexpr int L1_KPT = 32;                   // K1: one packed word (32 start positions) per thread
constexpr int L1_TILE = BLOCK * L1_KPT;      // 8192 keys
constexpr int L2_KPT = 16;
constexpr int L2_TILE = BLOCK * L2_KPT;      // 4096 keys
constexpr int MAX_B = 1024;                  // max buckets per partition level
#ifndef SIMKA_BIN_TARGET
#define SIMKA_BIN_TARGET 768
#endif
#ifndef SIMKA_K3_BLOCK
#define SIMKA_K3_BLOCK 256
#endif
#ifndef SIMKA_CNT_SLOTS
#define SIMKA_CNT_SLOTS 2048
#endif
constexpr int BIN_TARGET = SIMKA_BIN_TARGET;              // fine bins hold BIN_TARGET..2*BIN_TARGET keys on average
constexpr int K3_BLOCK = SIMKA_K3_BLOCK;      // threads of a counting CTA (one fine bin at a time)
constexpr int CNT_SLOTS = SIMKA_CNT_SLOTS;    // K3 hash table slots
constexpr int CNT_MAXDISTINCT = CNT_SLOTS * 3 / 4;  // distinct keys one table pass may hold (load 0.75)
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
    int key_mode;        // 0: key = mix(canonical k-mer) (2k bits); 1: key = MurmurHash3(canonical k-mer) (64 bits, SimkaMin sketches)
    uint32_t seed;       // MurmurHash3 seed
    uint64_t key_max;    // key_mode 1: hashes above this bound are dropped at extraction (bottom-s pre-filter of the sketches; ~0 = keep all)
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

// Multi-block exclusive scan for the large arrays (micro-bin histograms, fine-bin histograms, survivors): three short launches --
// per-block sums, a single-block scan of the sums, per-block scan with the block's prefix.  No block ever waits for another one
// (an earlier version chained the blocks through a spin on the predecessor's word, which serialised the 128 blocks -- 0.2 ms
// per scan -- and relied on all of them being co-resident).  `stride` reads every stride-th u32 of `in` (2: one half of a
// packed u64 histogram).
constexpr int SCAN_CHUNK = 1024 * 16;
__global__ void __launch_bounds__(1024) k_scan_reduce(const uint32_t *__restrict__ in, uint64_t n, int stride, uint64_t *__restrict__ partial) {
    __shared__ uint64_t warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_CHUNK;
    uint64_t v = 0;
#pragma unroll 4
    for (int q = 0; q < 16; q++) {
        const uint64_t i = i0 + (uint64_t)q * 1024 + tid;
        if (i < n) v += in[i * stride];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) warp_sum[wid] = v;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = warp_sum[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w += __shfl_down_sync(0xffffffffu, w, d);
        if (lane == 0) partial[blockIdx.x] = w;
    }
}
// in place: partial[0..m) -> exclusive prefix, partial[m] = total; single block
__global__ void __launch_bounds__(1024) k_scan_partials(uint64_t *__restrict__ partial, uint32_t m) {
    __shared__ uint64_t warp_sum[33];
    __shared__ uint64_t carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += 1024) {
        const uint32_t i = base + tid;
        const uint64_t v = i < m ? partial[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
        if (lane == 31) warp_sum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = warp_sum[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint64_t u = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += u; }
            warp_sum[lane] = winc - w;
            if (lane == 31) warp_sum[32] = winc;
        }
        __syncthreads();
        if (i < m) partial[i] = carry + warp_sum[wid] + inc - v;
        __syncthreads();
        if (tid == 0) carry += warp_sum[32];
        __syncthreads();
    }
    if (tid == 0) partial[m] = carry;
}
__global__ void __launch_bounds__(1024) k_scan_apply(const uint32_t *__restrict__ in, uint64_t n, int stride, const uint64_t *__restrict__ partial,
                                                     uint64_t *__restrict__ out) {
    __shared__ uint64_t warp_sum[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)tid * 16;
    uint32_t x[16];
    if (stride == 1 && i0 + 16 <= n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 t4 = *reinterpret_cast<const uint4 *>(in + i0 + 4 * q);
            x[4 * q] = t4.x; x[4 * q + 1] = t4.y; x[4 * q + 2] = t4.z; x[4 * q + 3] = t4.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) x[q] = (i0 + q < n) ? in[(i0 + q) * stride] : 0;
    }
    uint64_t v = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) v += x[q];
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
    if (lane == 31) warp_sum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = warp_sum[lane], winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t u = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += u; }
        warp_sum[lane] = winc - w;
    }
    __syncthreads();
    uint64_t run = partial[blockIdx.x] + warp_sum[wid] + inc - v;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        if (i0 + q < n) out[i0 + q] = run;
        run += x[q];
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) out[n] = partial[gridDim.x];
}
