// kernels/simkamin.cuh -- SimkaMin sketch x sketch distances
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
#pragma once

// ---------------------------------------------------------------------------------------------
// SimkaMin: all-pairs bottom-s sketch comparison.
// Replaces ComputeDistanceManager::computeDistance_unsynch (src/simkaMin/SimkaMinDistance.hpp:190-279).  One CTA per pair streams the two sketches through shared memory in
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

// RECT = false: pair p of the upper triangle of one set of n sketches; RECT = true: pair (p / n2, p % n2) of two sets
// (hashes/counts/start: n sketches; hashes2/counts2/start2: n2 sketches), the rectangular blocks of simkaMinCore distance
template <bool RECT>
__global__ void __launch_bounds__(BLOCK)
k_simkamin_tiled(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ start,
                 uint32_t n, const uint64_t *__restrict__ hashes2, const uint32_t *__restrict__ counts2, const uint32_t *__restrict__ start2,
                 uint32_t n2, uint32_t s, ull *__restrict__ quad /*[pairs][4]: shared, distinct, sharedK, kmers*/) {
    static_assert(BLOCK == 256, "block_sum3 assumes 8 warps");
    const uint64_t p = blockIdx.x;
    int i, j;
    if (RECT) { i = (int)(p / n2); j = (int)(p % n2); }
    else {
        const double tn = 2.0 * n - 1.0;
        i = (int)((tn - sqrt(tn * tn - 8.0 * (double)p)) * 0.5);
        if (i < 0) i = 0;
        while ((uint64_t)(i + 1) * (2ull * n - (i + 1) - 1) / 2 <= p) i++;
        while ((uint64_t)i * (2ull * n - i - 1) / 2 > p) i--;
        j = i + 1 + (int)(p - (uint64_t)i * (2ull * n - i - 1) / 2);
    }
    const uint64_t *A = hashes + (size_t)i * s + start[i];
    const uint64_t *B = (RECT ? hashes2 : hashes) + (size_t)j * s + (RECT ? start2 : start)[j];
    const uint32_t *CA = counts + (size_t)i * s + start[i];
    const uint32_t *CB = (RECT ? counts2 : counts) + (size_t)j * s + (RECT ? start2 : start)[j];
    const uint32_t na = s - start[i], nb = s - (RECT ? start2 : start)[j];
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

// ---------------------------------------------------------------------------------------------
// SimkaMin sketch, second look at the reads (the multiset count itself runs through stage 1 with
// key = MurmurHash3): the reference counts occurrences of the LARGEST selected hash only until the sketch is
// complete (processUnfiltered tests `hash < top`, src/simkaMin/SimkaMinCount.hpp:324), so its count depends on the
// processing order.
//   MODE 0: first occurrence position of every sketch member (without -filter the sketch is complete at the largest
//           of them).
//   MODE 2: second occurrence position of every member, given the first ones (-filter: a k-mer enters the heap at its
//           second occurrence, processFiltered :341-380, so the sketch is complete at the largest second position).
//   MODE 1: occurrences of the largest member at positions t_from <= pos <= t_star.
// Thread = one packed word = 32 k-mer start positions, same roll as k_extract.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(BLOCK)
k_sketch_scan(const uint64_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t n_words, int k, uint32_t seed,
              const uint64_t *__restrict__ sketch, uint32_t n_sel, ull *__restrict__ first_pos, ull *__restrict__ second_pos, ull t_from,
              ull t_star, ull *__restrict__ count_out) {
    const uint64_t t = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (t >= n_words || n_sel == 0) return;
    uint64_t hi = seq[t], lo = seq[t + 1];
    const uint64_t bm = ((uint64_t)bad[t] << 32) | bad[t + 1];
    uint32_t okmask;
    {
        uint64_t r = bm;
        int built = 1;
        while (built * 2 <= k) { r |= r << built; built *= 2; }
        if (built < k) r |= r << (k - built);
        okmask = ~(uint32_t)(r >> 32);
    }
    const uint64_t top = sketch[n_sel - 1];
    uint64_t rev = 0;
    ull local = 0;
    for (int j = 0; j < 32; j++) {
        const uint64_t fwd = hi >> (64 - 2 * k);
        if (j == 0) rev = simka::revcomp(fwd, k);
        else rev = (rev >> 2) | ((uint64_t)((fwd & 3) ^ 2) << (2 * k - 2));
        if ((okmask >> (31 - j)) & 1u) {
            const uint64_t h = simka::murmur3_x64_128_u64(fwd < rev ? fwd : rev, seed);
            const ull pos = (ull)t * 32 + j;
            if (MODE == 0 || MODE == 2) {
                if (h <= top) {
                    uint32_t a = 0, b = n_sel;  // lower bound
                    while (a < b) { const uint32_t mid = (a + b) >> 1; if (sketch[mid] < h) a = mid + 1; else b = mid; }
                    if (a < n_sel && sketch[a] == h) {
                        if (MODE == 0) atomicMin(&first_pos[a], pos);
                        else if (pos != first_pos[a]) atomicMin(&second_pos[a], pos);
                    }
                }
            } else if (h == top && pos >= t_from && pos <= t_star) local++;
        }
        hi = (hi << 2) | (lo >> 62);
        lo <<= 2;
    }
    if (MODE == 1 && local) atomicAdd(count_out, local);
}
