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
#ifndef SIMKA_SMT_EPT
#define SIMKA_SMT_EPT 6
#endif
constexpr int SMT_EPT = SIMKA_SMT_EPT;        // merged elements per thread and tile
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
// (hashes/counts/start: n sketches; hashes2/counts2/start2: n2 sketches), the rectangular blocks of simkaMinCore distance.
//
// Tiles arrive by 1-D bulk copies (cp.async.bulk + mbarrier, kernels/async.cuh) into a double buffer: as soon as the merge-path
// split of tile t is known (how many elements of A and of B the tile consumes), one thread starts the copies of tile t+1, which
// then overlap the walk and the block reductions of tile t -- the per-tile chain "load -> split -> walk -> reduce" was latency
// bound at 23 % of the HBM rate.  The copies need 16-byte aligned sources: each starts at the aligned element below the
// wanted one (`o*` = elements to skip in shared memory) and the arrays carry 64 bytes of padding behind their last element
// (simkamin_gpu_distance allocates them).  The prefetch is speculative (the tile may turn out to be the last): an
// outstanding copy is always waited for before the CTA exits.
// Measured (config 5, B200): 75.1 ms with plain LDG->STS tile loads, 59.5 ms with the double-buffered bulk copies.  A tile still
// loads up to SMT_TILE elements of BOTH sketches although it consumes SMT_TILE merged ones; sizing the requests from the previous
// tile's split (8 % instead of 100 % over-read, tile length found by two more binary searches) was tried and is SLOWER (71.8 ms):
// the kernel is bound by the per-tile dependent chain, not by HBM traffic.
struct SmtBuf {
    alignas(16) uint64_t a[SMT_TILE + 2], b[SMT_TILE + 2];
    alignas(16) uint32_t ca[SMT_TILE + 4], cb[SMT_TILE + 4];
};
struct SmtSmem {
    SmtBuf buf[2];
    alignas(8) uint64_t bar[2];
    ull red[24];
    uint32_t scan[BLOCK / 32 + 1];
    uint32_t xs, tstar;
};

template <bool RECT>
__global__ void __launch_bounds__(BLOCK)
k_simkamin_tiled(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ start,
                 uint32_t n, const uint64_t *__restrict__ hashes2, const uint32_t *__restrict__ counts2, const uint32_t *__restrict__ start2,
                 uint32_t n2, uint32_t s, ull *__restrict__ quad /*[pairs][4]: shared, distinct, sharedK, kmers*/) {
    static_assert(BLOCK == 256, "block_sum3 assumes 8 warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmtSmem &M = *reinterpret_cast<SmtSmem *>(smem_raw);
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
    // element indices of the first sketch members inside their arrays (the arrays themselves are 256-byte aligned)
    const uint64_t *HA = hashes, *HB = RECT ? hashes2 : hashes;
    const uint32_t *KA = counts, *KB = RECT ? counts2 : counts;
    const uint64_t ea0 = (uint64_t)i * s + start[i], eb0 = (uint64_t)j * s + (RECT ? start2 : start)[j];
    const uint32_t na = s - start[i], nb = s - (RECT ? start2 : start)[j];
    const uint32_t S = min(na, nb);  // sketchSize = min of the two stripped sizes (:204)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { mbar_init(&M.bar[0], 1); mbar_init(&M.bar[1], 1); mbar_fence_init(); }
    __syncthreads();

    // thread 0: start the copies of the tile that begins at (a0, b0) into buffer `bf`; returns false when there is nothing to load
    auto issue = [&](uint32_t a0, uint32_t b0, int bf) -> bool {
        const uint32_t la = min(na - a0, (uint32_t)SMT_TILE), lb = min(nb - b0, (uint32_t)SMT_TILE);
        if (la + lb == 0) return false;
        const uint64_t ea = ea0 + a0, eb = eb0 + b0;
        const uint32_t ba = la ? (((uint32_t)(ea & 1) + la) * 8 + 15) & ~15u : 0, bb = lb ? (((uint32_t)(eb & 1) + lb) * 8 + 15) & ~15u : 0;
        const uint32_t bca = la ? (((uint32_t)(ea & 3) + la) * 4 + 15) & ~15u : 0, bcb = lb ? (((uint32_t)(eb & 3) + lb) * 4 + 15) & ~15u : 0;
        mbar_arrive_expect_tx(&M.bar[bf], ba + bb + bca + bcb);
        if (la) { bulk_g2s(M.buf[bf].a, HA + (ea & ~1ull), ba, &M.bar[bf]); bulk_g2s(M.buf[bf].ca, KA + (ea & ~3ull), bca, &M.bar[bf]); }
        if (lb) { bulk_g2s(M.buf[bf].b, HB + (eb & ~1ull), bb, &M.bar[bf]); bulk_g2s(M.buf[bf].cb, KB + (eb & ~3ull), bcb, &M.bar[bf]); }
        return true;
    };

    uint32_t a0 = 0, b0 = 0;            // elements of A / B consumed so far
    ull distinct = 0, shared = 0, sharedk = 0, kmers = 0;
    uint64_t prevA = ~0ull;             // last A element consumed before this tile (a duplicate B may open the tile)
    uint32_t prevCA = 0;
    uint32_t ph0 = 0, ph1 = 0;
    int cur = 0;
    bool pending_next = false;          // (thread 0) a prefetch into buffer cur ^ 1 is in flight
    if (tid == 0) issue(0, 0, 0);
    for (;;) {
        const uint32_t ra = na - a0, rb = nb - b0;
        const uint32_t la = min(ra, (uint32_t)SMT_TILE), lb = min(rb, (uint32_t)SMT_TILE);
        const uint32_t tl = min((uint32_t)SMT_TILE, ra + rb);
        if (tl == 0) break;  // (distinct == S at a tile end is NOT a stop: a duplicate B of the last A may open the next tile)
        if (cur == 0) { mbar_wait(&M.bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&M.bar[1], ph1); ph1 ^= 1; }
        pending_next = false;
        const uint64_t *sa = M.buf[cur].a + ((ea0 + a0) & 1), *sb = M.buf[cur].b + ((eb0 + b0) & 1);
        const uint32_t *ca = M.buf[cur].ca + ((ea0 + a0) & 3), *cb = M.buf[cur].cb + ((eb0 + b0) & 3);
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
        if (tid == 0) {  // elements of A the whole tile consumes -> where the next tile starts: prefetch it now
            uint32_t lo = tl > lb ? tl - lb : 0, hi = min(tl, la);
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sa[mid] <= sb[tl - 1 - mid]) lo = mid + 1; else hi = mid; }
            M.xs = lo;
            pending_next = issue(a0 + lo, b0 + (tl - lo), cur ^ 1);  // (buffer cur ^ 1 was released by the barrier that ended the previous tile)
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
        block_sum3(r_d, r_m, r_k, M.red);
        if (distinct + r_d <= S) {  // the (S+1)-th distinct value is not in this tile: take all of it
            ull r_sk = t_sk, z0 = 0, z1 = 0;
            block_sum3(r_sk, z0, z1, M.red);
            distinct += r_d; shared += r_m; kmers += r_k; sharedk += r_sk;
            const uint32_t xs = M.xs;
            if (xs > 0) { prevA = sa[xs - 1]; prevCA = ca[xs - 1]; }
            a0 += xs;
            b0 += tl - xs;
            __syncthreads();  // tile fully consumed: its buffer may be refilled by the prefetch of the tile after the next one
            cur ^= 1;
            continue;
        }
        // cut inside this tile: first merged element that would be distinct value number S+1
        const uint32_t need = (uint32_t)(S - distinct);  // distinct values still to take (< r_d)
        uint32_t inc = t_dist;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += u; }
        if (lane == 31) M.scan[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < wid; w++) wbase += M.scan[w];
        const uint32_t excl = wbase + inc - t_dist;  // distinct values before this thread's run
        // the thread whose run contains the element that makes the count exceed `need`
        const bool mine = excl <= need && excl + t_dist > need;
        if (mine) M.tstar = tid;
        __syncthreads();
        const uint32_t tstar = M.tstar;
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
        block_sum3(f_m, f_k, f_sk, M.red);
        distinct = S; shared += f_m; kmers += f_k; sharedk += f_sk;
        break;
    }
    if (tid == 0) {
        if (pending_next) { if (cur == 0) mbar_wait(&M.bar[1], ph1); else mbar_wait(&M.bar[0], ph0); }  // never exit under a copy in flight
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
