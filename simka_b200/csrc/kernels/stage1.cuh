// kernels/stage1.cuh -- stage 1 (simkaCount): packing, k-mer extraction + L1 partition, L2 partition, per-bin counting, compaction
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
#pragma once

// ---------------------------------------------------------------------------------------------
// P0: ASCII -> 2-bit packed stream + bad mask (GPU-side packing for simka_gpu_push_ascii).
// Stream position p = read_off[r] + r + o: one separator (bad) position after every read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_pack_ascii(const char *__restrict__ seqs, const uint64_t *__restrict__ read_off,
                                                      uint64_t n_reads, uint64_t n_words, uint64_t *__restrict__ packed,
                                                      uint32_t *__restrict__ bad, int badchars /* 0: k <= 15, only separators are bad */) {
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
                isbad = badchars && !(u == 'A' || u == 'C' || u == 'G' || u == 'T');
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
            const uint64_t h = kp.key_mode ? simka::murmur3_x64_128_u64(canon, kp.seed) : simka::mix_fwd(canon, kp.mix);
            bool keep = (okmask >> (31 - j)) & 1u;
            if (SHARDED) keep = keep && shard_keep(h, kp);
            if (kp.key_mode) keep = keep && h <= kp.key_max;  // sketches: only the low end of the hash range is ever selected
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
// MINB = CTAs per SM the register budget is cut for: 4 (64 registers; the scatter pass spills 40 bytes per thread) is what was
// measured; 3 (85 registers, no spill) is selectable with SIMKA_B200_K2_MINB=3 for an A/B on the GPU.
template <bool COUNT_ONLY, int MINB = 4>
__global__ void __launch_bounds__(BLOCK, MINB)
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

// One probe step of the open-addressing table (linear probing).  Plain read first: most keys repeat a k-mer that is
// already there (one LDS + one ATOMS.POPC.INC); a thread that claims a free slot appends it to the claimed list, which is
// what the read-out walks (the table itself is never scanned) and what counts the distinct keys of the bin.
// Returns true when the key is placed.
__device__ __forceinline__ bool table_probe(uint64_t *tk, uint32_t *tc, uint16_t *claimed, uint32_t *nclaim, uint64_t h, uint32_t slot) {
    volatile uint64_t *vk = tk;
    uint64_t cur = vk[slot];
    if (cur == EMPTY) {
        cur = atomicCAS(reinterpret_cast<ull *>(&tk[slot]), (ull)EMPTY, (ull)h);
        if (cur == EMPTY) { cur = h; claimed[atomicAdd(nclaim, 1u)] = (uint16_t)slot; }
    }
    if (cur == h) { atomicAdd(&tc[slot], 1u); return true; }
    return false;
}
// Insert up to four keys per thread.  Round 1 probes the home slots of the four keys one after the other with the warp
// converged: this places 9 keys out of 10.  The leftovers -- a different k-mer sits in the home slot -- are walked one key
// per thread in a tight loop: in the unrolled four-key form those later rounds ran with one or two active lanes per warp and
// were 14 % of the kernel's instructions (ncu source view of the first version).
__device__ __forceinline__ void table_insert4(uint64_t *tk, uint32_t *tc, uint16_t *claimed, uint32_t *nclaim, const uint64_t (&h)[4],
                                              uint32_t pending, int slot_shift) {
    uint32_t slot[4];
#pragma unroll
    for (int q = 0; q < 4; q++) slot[q] = (uint32_t)(h[q] >> slot_shift) & (CNT_SLOTS - 1);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (pending & (1u << q)) {
            if (table_probe(tk, tc, claimed, nclaim, h[q], slot[q])) pending &= ~(1u << q);
        }
    }
    while (pending) {
        const uint64_t hk = (pending & 1u) ? h[0] : (pending & 2u) ? h[1] : (pending & 4u) ? h[2] : h[3];
        uint32_t sl = (pending & 1u) ? slot[0] : (pending & 2u) ? slot[1] : (pending & 4u) ? slot[2] : slot[3];
        pending &= pending - 1;
        // a full table leaves nclaim == CNT_SLOTS > CNT_MAXDISTINCT, which makes the caller retry with finer sub-passes
        for (int probes = 1; probes < CNT_SLOTS; probes++) {
            sl = (sl + 1) & (CNT_SLOTS - 1);
            if (table_probe(tk, tc, claimed, nclaim, hk, sl)) break;
        }
    }
}

// MULTI: the bin may hold more distinct keys than one table pass takes: stop early on overflow, R sub-passes
template <bool MULTI>
__device__ __forceinline__ void bin_insert_all(uint64_t *tk, uint32_t *tc, uint16_t *claimed, uint32_t *nclaim,
                                               const uint64_t *__restrict__ kb, uint32_t n, uint32_t R, uint32_t r, int slot_shift) {
    const int tid = threadIdx.x;
    for (uint32_t base = 0; base < n; base += K3_BLOCK * 4) {
        uint64_t h[4];
        uint32_t pending = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = base + q * K3_BLOCK + tid;
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
__global__ void __launch_bounds__(K3_BLOCK)
k_bin_count(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ fine_off, uint32_t nbins, int slot_shift,
            int64_t amin, int64_t amax, CountOut out) {
    __shared__ __align__(16) uint64_t tk[CNT_SLOTS];
    __shared__ __align__(16) uint32_t tc[CNT_SLOTS];
    __shared__ uint16_t claimed[CNT_SLOTS];
    __shared__ uint32_t s_cursor, s_nclaim[2];
    __shared__ ull s_tot[3];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < CNT_SLOTS; i += K3_BLOCK) { tk[i] = EMPTY; tc[i] = 0; }
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
                for (uint32_t e0 = 0; e0 < nc; e0 += K3_BLOCK) {
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
// big_n counts every large abundance, also those beyond big_cap: the merge then grows the list and rebuilds it from the
// resident lists (chist == nullptr: only the list is filled, the dense histogram is already complete).
__global__ void __launch_bounds__(BLOCK)
k_abundance_hist(const uint32_t *__restrict__ ec, uint64_t n, int sample, ull *__restrict__ chist, uint32_t *__restrict__ big_counts,
                 uint32_t *__restrict__ big_sample, uint32_t *__restrict__ big_n, uint32_t big_cap, int *__restrict__ error_flag) {
    __shared__ uint32_t hist[CHIST];
    for (int i = threadIdx.x; i < CHIST; i += BLOCK) hist[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * BLOCK) {
        const uint32_t c = ec[i];
        if (c < CHIST) { if (chist) atomicAdd(&hist[c], 1u); }
        else {
            const uint32_t q = atomicAdd(big_n, 1u);
            if (q < big_cap) { big_counts[q] = c; big_sample[q] = (uint32_t)sample; }
        }
    }
    __syncthreads();
    if (chist)
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
