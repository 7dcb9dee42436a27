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
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/simka_b200.h"
#include "kmer_device.cuh"

// NCCL: types from its header at build time, the library itself is loaded at run time (dlopen) so that nothing but the multi-GPU
// entry points depends on it
#if !defined(SIMKA_NO_NCCL) && __has_include(<nccl.h>)
#include <nccl.h>
#define SIMKA_HAVE_NCCL_H 1
#else
#define SIMKA_HAVE_NCCL_H 0
#endif

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

#include "kernels/common.cuh"
#include "kernels/async.cuh"
#include "kernels/stage1.cuh"
#include "kernels/superk.cuh"
#include "kernels/superk2.cuh"
#include "kernels/merge.cuh"
#include "kernels/merge_tiles.cuh"
#include "kernels/simkamin.cuh"
#include "kernels/wide.cuh"

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

// bump allocator for the per-sample solid k-mer lists (kept across simka_gpu_reset).  Small requests share 256 MB chunks; a
// request of 64 MB or more gets a chunk of exactly its size (512 samples x 290 MB of lists leave no room for 40 % of slack), and
// best fit makes the next step -- same sizes -- find the same chunks again.
struct Arena {
    struct Chunk { char *p; size_t cap, used; };
    std::vector<Chunk> chunks;
    void *alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        Chunk *best = nullptr;
        for (auto &c : chunks)
            if (c.cap - c.used >= bytes && (!best || c.cap - c.used < best->cap - best->used)) best = &c;
        if (best) { void *r = best->p + best->used; best->used += bytes; return r; }
        size_t cap = bytes >= ((size_t)64 << 20) ? bytes : std::max<size_t>(bytes, (size_t)256 << 20);
        void *p = nullptr;
        if (cudaMalloc(&p, cap) != cudaSuccess) {
            cudaGetLastError();
            if (cap == bytes || cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            cap = bytes;
        }
        chunks.push_back({(char *)p, cap, bytes});
        return p;
    }
    void reset() { for (auto &c : chunks) c.used = 0; }
    void release() { for (auto &c : chunks) cudaFree(c.p); chunks.clear(); }
};

struct SampleData {
    uint64_t *eh = nullptr;   // hashed solid k-mers, grouped by fine range (arena memory)
    K128 *ehw = nullptr;      // the same for k > 31 (16-byte keys; eh is unused then)
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
    DevBuf d_off_table2; // [N][FG2+1]: offsets of the lists regrouped for the merge (many samples)
    bool no_refine = false;  // SIMKA_B200_REFINE=0: merge on the stage-1 granularity (A/B measurements)
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
    DevBuf d_ptrs, d_logtab, d_scan_tmp, d_isum;
    bool import_pending = false;
    bool use_k16 = true;        // SIMKA_B200_K1=256: the 256-thread x 32-position extraction kernel
    bool force_exact = false;   // SIMKA_B200_EXACT_COUNT=1: always run the K1 count pass (tests of the fallback)
    bool use_superk = false;    // SIMKA_B200_SUPERK=1|2: super-k-mer front end of stage 1 (kernels/superk.cuh; 2 = records all the way, kernels/superk2.cuh)
    int superk_mode = 0;
    bool wide = false;          // k > 31: 128-bit k-mers (kernels/wide.cuh)
    Mix128 mix128;
    uint32_t sk2_target = 0;    // SIMKA_B200_SK2_TARGET: k-mers per counting bin of the record path (experiments; default SK2_TARGET)
    uint64_t exact_retries = 0, superk_samples = 0, superk_fallbacks = 0;
    uint64_t superk_cap_override = 0;
    bool superk_expand_thread = false;
    DevBuf d_sk_tmpc, d_sk_prefix, d_sk_rprefix, d_sk_recoff, d_sk_hist64;
    DevBuf d_dump;  // scratch of simka_gpu_dump_sample_counts (parity hook)
    float stage_ms[T_NSTAGE] = {};
    float sample_ms = 0, merge_ms = 0;
    uint64_t kmers_in = 0, entries = 0;
    uint64_t launches = 0;
    int nsm = 148;
    int occ_bin_count = 0, occ_sk2_count = 0;
    // multi-GPU (simka_gpu_comm_init / simka_gpu_exchange_and_merge)
    void *comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    simka_ctx *mctx = nullptr;   // merge context of this rank's block of hash ranges
    bool reduced = false;        // accumulators already summed over the ranks: closed forms use the global totals
    DevBuf d_xhdr;
    cudaEvent_t xev[5] = {};
    float xms[4] = {0, 0, 0, 0};
    // the exchange runs in ROUNDS (round t = the samples t * world .. t * world + world - 1, one per rank) on a side stream, so that
    // the lists of round t cross NVLink while round t + 1 is being counted
    cudaStream_t xstream = nullptr;
    struct XRound { int state = 0; cudaEvent_t sizes_ready = nullptr; };  // 0: nothing yet, 1: sizes all-reduce enqueued, 2: data posted
    std::vector<XRound> xr;
    ull *xh = nullptr;           // pinned: [rounds][world * world + 3] entries per (owner, destination) + geometry check
    DevBuf d_xcnt, d_xids;
    std::vector<uint64_t *> x_recv_e;
    std::vector<uint32_t *> x_recv_c;
    std::vector<uint64_t> x_recv_n;
    uint32_t *x_recv_o = nullptr;
    int x_parts = 0, x_lb = 0, x_fb = -1;
    cudaEvent_t x_sample_done = nullptr, x_all_done = nullptr;
    int b1_override = 0;  // SIMKA_B200_B1BITS: level-1 bucket bits (experiments)
    bool k2_minb3 = false;  // SIMKA_B200_K2_MINB=3: scatter pass of K2 compiled for 3 CTAs per SM (no register spill; experiments)

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
    const bool bigk = !ctx->kp.key_mode && ctx->kp.k >= 17 && b1_bits <= 2 * ctx->kp.k - 32;
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

// exclusive scan u32[n] -> u64[n+1] on the context's stream (stride 2: one half of a packed u64 histogram)
static int launch_scan(simka_ctx *ctx, const uint32_t *in, uint64_t *out, uint64_t n, int stride = 1) {
    if (n <= (uint64_t)SCAN_CHUNK && stride == 1) { k_scan_u32_u64<<<1, 1024, 0, ctx->stream>>>(in, out, n); ctx->launches++; return 0; }
    const uint64_t nblk = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (ctx->d_scan_tmp.reserve((nblk + 1) * 8)) return -1;
    k_scan_reduce<<<(unsigned)nblk, 1024, 0, ctx->stream>>>(in, n, stride, ctx->d_scan_tmp.as<uint64_t>());
    k_scan_partials<<<1, 1024, 0, ctx->stream>>>(ctx->d_scan_tmp.as<uint64_t>(), (uint32_t)nblk);
    k_scan_apply<<<(unsigned)nblk, 1024, 0, ctx->stream>>>(in, n, stride, ctx->d_scan_tmp.as<uint64_t>(), out);
    ctx->launches += 3;
    return 0;
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
    ctx->reduced = false;
    ctx->fg_bits = -1;
    for (float &f : ctx->stage_ms) f = 0;
    ctx->sample_ms = ctx->merge_ms = 0;
    ctx->kmers_in = ctx->entries = 0;
    ctx->launches = 0;
    if (ctx->prm.complex_dist) CK(cudaMemsetAsync(ctx->d_chist.p, 0, (size_t)ctx->prm.nb_samples * CHIST * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_small.p, 0, 64, ctx->stream));
    return SIMKA_OK;
}

static int xchg_after_sample(simka_ctx *ctx, int sample_id);  // multi-GPU: defined with the NCCL code at the end of the file

extern "C" {

const char *simka_b200_version(void) { return "simka_b200 0.1 (sm_100a)"; }

int simka_gpu_host_alloc(void **ptr, size_t bytes) { return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? SIMKA_OK : SIMKA_ERR_NOMEM; }
int simka_gpu_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? SIMKA_OK : SIMKA_ERR_CUDA; }

int simka_gpu_create(simka_ctx **out, const simka_params *p) try {
    if (!out || !p) return SIMKA_ERR_INVALID;
    *out = nullptr;
    if (p->kmer_size < 2 || p->kmer_size > 63 || p->nb_samples < 1 || p->nb_samples > MRG_MAXN) return SIMKA_ERR_INVALID;
    if (p->kmer_size > 31 && (p->sketch_mode || p->shard_count > 1)) return SIMKA_ERR_INVALID;  // 128-bit k-mers: plain single-context runs only
    if (p->shard_count > 1 && (p->shard_rank < 0 || p->shard_rank >= p->shard_count)) return SIMKA_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || p->device < 0 || p->device >= ndev) return SIMKA_ERR_CUDA;
    if (cudaSetDevice(p->device) != cudaSuccess) return SIMKA_ERR_CUDA;
    simka_ctx *ctx = new simka_ctx();
    ctx->prm = *p;
    ctx->wide = p->kmer_size > 31;
    ctx->kp.mix = simka::make_mix(ctx->wide ? 31 : p->kmer_size);
    ctx->mix128 = make_mix128(p->kmer_size);
    ctx->kp.k = p->kmer_size;
    ctx->kp.shard_rank = p->shard_rank;
    ctx->kp.shard_count = p->shard_count > 1 ? p->shard_count : 1;
    ctx->kp.key_mode = p->sketch_mode ? 1 : 0;
    ctx->kp.seed = p->sketch_seed;
    ctx->kp.key_max = ~0ull;
    ctx->W = 2 * p->kmer_size;
    if (ctx->kp.key_mode) {  // keys are 64-bit MurmurHash3 values: partition on their top bits
        ctx->W = 64;
        ctx->kp.mix.bits = 64;
        ctx->kp.mix.mask = ~0ull;
    }
    { const char *e = getenv("SIMKA_B200_EXACT_COUNT"); ctx->force_exact = e && e[0] == '1'; }
    { const char *e = getenv("SIMKA_B200_K1"); ctx->use_k16 = !(e && !strcmp(e, "256")); }
    { const char *e = getenv("SIMKA_B200_B1BITS"); ctx->b1_override = e ? atoi(e) : 0; }
    { const char *e = getenv("SIMKA_B200_K2_MINB"); ctx->k2_minb3 = e && e[0] == '3'; }
    { const char *e = getenv("SIMKA_B200_REFINE"); ctx->no_refine = e && e[0] == '0'; }
    { const char *e = getenv("SIMKA_B200_SUPERK"); ctx->superk_mode = e ? atoi(e) : 0; ctx->use_superk = ctx->superk_mode == 1 || ctx->superk_mode == 2; }
    { const char *e = getenv("SIMKA_B200_SK2_TARGET"); ctx->sk2_target = e ? (uint32_t)std::min(std::max(atoi(e), 64), SK2_MAXD) : 0; }
    { const char *e = getenv("SIMKA_B200_SUPERK_EXPAND"); ctx->superk_expand_thread = e && !strcmp(e, "thread"); }
    { const char *e = getenv("SIMKA_B200_SUPERK_CAP"); ctx->superk_cap_override = e ? strtoull(e, nullptr, 10) : 0; }  // tests of the fallback
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
    ok = ok && cudaFuncSetAttribute(k_wide_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_merge_smem_bytes(p->nb_samples)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_sk2_count<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sk2Smem)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_sk2_count<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sk2Smem)) == cudaSuccess;
    ok = ok && ctx->d_small.reserve(64) == 0;
    ok = ok && ctx->d_scan_tmp.reserve(((size_t)1 << 17) * 8 + 64) == 0;  // block sums of the device-wide scans (up to 2^31 elements)
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
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    return SIMKA_ERR_NOMEM;
}

int simka_gpu_reset(simka_ctx *ctx) {
    if (!ctx) return SIMKA_ERR_INVALID;
    CK(cudaSetDevice(ctx->prm.device));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->xstream) CK(cudaStreamSynchronize(ctx->xstream));
    for (auto &r : ctx->xr) r.state = 0;
    ctx->x_fb = -1;
    ctx->x_recv_o = nullptr;
    if (ctx->mctx) { const int rc = simka_gpu_reset(ctx->mctx); if (rc != SIMKA_OK) { ctx->set_error(ctx->mctx->err); return rc; } }
    return reset_state(ctx);
}

static void comm_release(simka_ctx *ctx);

void simka_gpu_destroy(simka_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->prm.device);
    comm_release(ctx);
    if (ctx->mctx) { simka_gpu_destroy(ctx->mctx); ctx->mctx = nullptr; }
    for (auto &e : ctx->xev) if (e) cudaEventDestroy(e);
    ctx->d_xhdr.release(); ctx->d_xcnt.release(); ctx->d_xids.release();
    for (auto &r : ctx->xr) if (r.sizes_ready) cudaEventDestroy(r.sizes_ready);
    if (ctx->x_sample_done) cudaEventDestroy(ctx->x_sample_done);
    if (ctx->x_all_done) cudaEventDestroy(ctx->x_all_done);
    if (ctx->xh) cudaFreeHost(ctx->xh);
    if (ctx->xstream) { cudaStreamSynchronize(ctx->xstream); cudaStreamDestroy(ctx->xstream); }
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    ctx->arena.release();
    DevBuf *bufs[] = {&ctx->slots[0].packed, &ctx->slots[0].bad, &ctx->slots[1].packed, &ctx->slots[1].bad, &ctx->d_ascii, &ctx->d_roff,
                      &ctx->d_keysA, &ctx->d_keysB, &ctx->d_stage_c, &ctx->d_hist1, &ctx->d_base1, &ctx->d_cursor1, &ctx->d_tiles,
                      &ctx->d_tile_start, &ctx->d_fine_hist, &ctx->d_fine_off, &ctx->d_fine_cursor, &ctx->d_surv, &ctx->d_out_off,
                      &ctx->d_small, &ctx->d_off_table, &ctx->d_off_table2, &ctx->d_chist, &ctx->d_big_counts, &ctx->d_big_sample, &ctx->d_acc, &ctx->d_ptrs, &ctx->d_logtab, &ctx->d_scan_tmp, &ctx->d_isum, &ctx->d_sk_tmpc, &ctx->d_sk_prefix, &ctx->d_sk_rprefix, &ctx->d_sk_recoff, &ctx->d_sk_hist64, &ctx->d_dump};
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
    if (keep && (cudaMemcpyAsync(np, b.p, keep, cudaMemcpyDeviceToDevice, ctx->copy_stream) != cudaSuccess ||
                 cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess)) {
        cudaFree(np);
        ctx->set_error("device copy failed while growing an input buffer");
        return SIMKA_ERR_CUDA;
    }
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

int simka_gpu_push_ascii(simka_ctx *ctx, int sample_id, const char *seqs, const uint64_t *read_off, uint64_t n_reads) try {
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
                                                       sl->packed.as<uint64_t>() + w0, sl->bad.as<uint32_t>() + w0, ctx->prm.kmer_size > 15);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(sl->copied, ctx->copy_stream));
    sl->words = w0 + n_words;
    sl->reads += n_reads;
    return SIMKA_OK;
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
}

static int finish_sample_wide(simka_ctx *ctx, int sample_id, simka_sample_totals *out);

int simka_gpu_finish_sample(simka_ctx *ctx, int sample_id, simka_sample_totals *out) try {
    if (!ctx) return SIMKA_ERR_INVALID;
    if (sample_id < 0 || sample_id >= ctx->prm.nb_samples || ctx->samples[sample_id].done) { ctx->set_error("bad sample id"); return SIMKA_ERR_INVALID; }
    if (ctx->wide) return finish_sample_wide(ctx, sample_id, out);
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
    // level-1 buckets: 256 for up to 2^17 bins (measured on config 2: K1 18.7 ms with 256 buckets against 23.0 ms with 1024 -- longer
    // contiguous runs per bucket and a shorter in-kernel scan -- while K2 loses 1.9 ms on its 512 sub-bins), more only as far as
    // the level-2 fan-out (<= 1024) requires
    int b1_bits = std::min(bin_bits, std::min(std::max(bin_bits - 9, 8), 10));
    if (ctx->b1_override > 0) b1_bits = std::min(bin_bits, std::max(ctx->b1_override, bin_bits - 10));
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
    bool sk_used = false;
    uint32_t nbins_fin = 0;  // bins of the attempt that succeeded

    if (n_words > 0) {
        // super-k-mer mode: micro-bins of ~48 positions (at most 2^21), counting bins cut from their histogram
        const int micro_bits = std::min(SK_MAXMICRO, std::max(8, ilog2_floor(std::max<uint64_t>(1, n_pos / 48))));
        const uint32_t n_micro = 1u << micro_bits;
        const bool sk2 = ctx->superk_mode == 2;  // records all the way: scatter the records, count straight from them
        const uint32_t sk_target = sk2 ? (ctx->sk2_target ? ctx->sk2_target : SK2_TARGET) : SK_TARGET;
        const uint32_t nb_sk = (uint32_t)std::min<uint64_t>(n_pos / sk_target + 2, 0x7FFFFFFFull);
        const size_t nres = std::max<size_t>(std::max<size_t>(nbins, (size_t)1 << ctx->fg_bits),
                                             ctx->use_superk ? std::max<size_t>(n_micro, (size_t)nb_sk + 1) : 0);
        bool okm = !ctx->d_keysA.reserve(n_pos * 8) && !ctx->d_keysB.reserve(n_pos * 8) && !ctx->d_stage_c.reserve(n_pos * 4) &&
                   !ctx->d_hist1.reserve(MAX_B * 4) && !ctx->d_base1.reserve((MAX_B + 1) * 8) && !ctx->d_cursor1.reserve(MAX_B * 4) &&
                   !ctx->d_tiles.reserve(MAX_B * 4) && !ctx->d_tile_start.reserve((MAX_B + 1) * 8) &&
                   !ctx->d_fine_hist.reserve(nres * 4) && !ctx->d_fine_off.reserve((nres + 1) * 8) &&
                   !ctx->d_fine_cursor.reserve(nres * 4) && !ctx->d_surv.reserve(nres * 4) &&
                   !ctx->d_out_off.reserve((nres + 1) * 8) && (!ctx->use_superk || !ctx->d_sk_prefix.reserve(((size_t)n_micro + 1) * 8)) &&
                   (!sk2 || (!ctx->d_sk_rprefix.reserve(((size_t)n_micro + 1) * 8) && !ctx->d_sk_recoff.reserve(((size_t)nb_sk + 2) * 8) &&
                             !ctx->d_sk_hist64.reserve((size_t)n_micro * 8)));
        if (!okm) { ctx->set_error("out of device memory (work buffers)"); return SIMKA_ERR_NOMEM; }
        CK(cudaStreamWaitEvent(st, sl->copied, 0));
        const int evset = ctx->ev_cur;
        harvest_timings(ctx, evset);  // the sample before the previous one: long finished
        cudaEvent_t *ev = ctx->ev[evset];
        CK(cudaEventRecord(ev[2 * T_NSTAGE], st));  // sample start
      // experimental front end (SIMKA_B200_SUPERK=1): super-k-mer records instead of the two k-mer partition levels
      const bool sk_eligible = ctx->use_superk && !ctx->kp.key_mode && (ctx->kp.shard_count <= 1 || sk2) && ctx->kp.k >= SK_MINK && ctx->kp.k <= 31 &&
                               !ctx->force_exact;
      for (int attempt = sk_eligible ? -1 : 0; attempt < 2; attempt++) {
        // attempt -1 (super-k-mer mode only): records; falls back to attempt 0 when the record list overflows.
        // attempt 0: single pass over the reads into fixed-capacity L1 buckets (sizes are binomial for hashed
        // keys); attempt 1 (only after an overflow, e.g. one k-mer repeated millions of times): exact count pass first
        const bool sk = attempt < 0;
        const bool exact = attempt == 1 || ctx->force_exact;
        const uint64_t per_bucket = n_pos / nb1;
        const uint64_t cap = per_bucket + 8 * (uint64_t)sqrt((double)per_bucket) + 1024;
        if (!exact && ctx->d_keysA.reserve((size_t)nb1 * cap * 8)) { ctx->set_error("out of device memory (work buffers)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemsetAsync(d_totals, 0, 24, st));
        CK(cudaMemsetAsync(d_overflow, 0, 4, st));
        CK(cudaMemsetAsync(ctx->d_hist1.p, 0, MAX_B * 4, st));
        CK(cudaMemsetAsync(ctx->d_cursor1.p, 0, MAX_B * 4, st));
        CK(cudaMemsetAsync(ctx->d_fine_hist.p, 0, (size_t)(sk ? n_micro : nbins) * 4, st));
        CK(cudaMemsetAsync(ctx->d_fine_cursor.p, 0, (size_t)(sk ? n_micro : nbins) * 4, st));
        const uint32_t nbins_run = sk ? nb_sk : nbins;  // bins the counting kernel walks

        const unsigned g1 = (unsigned)((n_words + BLOCK - 1) / BLOCK);
        const uint64_t *seq = sl->ext_packed ? sl->ext_packed : sl->packed.as<uint64_t>();
        const uint32_t *bad = sl->ext_bad ? sl->ext_bad : sl->bad.as<uint32_t>();
        if (sk) {
            // K1': reads -> unsorted super-k-mer records (in d_keysA) + k-mers per bin; K2': records -> keys grouped by bin
            ull *d_nrec = reinterpret_cast<ull *>(ctx->d_small.as<char>() + 48);
            uint64_t rec_cap = n_pos / 2;  // d_keysA holds n_pos * 8 bytes
            if (ctx->superk_cap_override) rec_cap = std::min<uint64_t>(rec_cap, ctx->superk_cap_override);
            CK(cudaMemsetAsync(d_nrec, 0, 8, st));
            CK(cudaEventRecord(ev[2 * T_EXTRACT_COUNT], st));
            CK(cudaEventRecord(ev[2 * T_EXTRACT_COUNT + 1], st));
            CK(cudaEventRecord(ev[2 * T_EXTRACT_SCATTER], st));
            if (sk2) CK(cudaMemsetAsync(ctx->d_sk_hist64.p, 0, (size_t)n_micro * 8, st));
            superk_scan_instance(ctx->kp.k)<<<(unsigned)((n_words + SK_THREADS - 1) / SK_THREADS), SK_THREADS, 0, st>>>(
                seq, bad, n_words, micro_bits, ctx->d_keysA.as<SkRecord>(), d_nrec, rec_cap, ctx->d_fine_hist.as<uint32_t>(), d_overflow,
                sk2 ? ctx->d_sk_hist64.as<ull>() : nullptr);
            CK(cudaEventRecord(ev[2 * T_EXTRACT_SCATTER + 1], st));
            CK(cudaEventRecord(ev[2 * T_PART_COUNT], st));
            if (sk2) {
                // (records | k-mers) per micro-bin -> two prefix arrays; counting bins cut on the k-mer prefix, their first records alongside
                launch_scan(ctx, ctx->d_sk_hist64.as<uint32_t>(), ctx->d_sk_prefix.as<uint64_t>(), n_micro, 2);       // low halves: k-mers
                launch_scan(ctx, ctx->d_sk_hist64.as<uint32_t>() + 1, ctx->d_sk_rprefix.as<uint64_t>(), n_micro, 2);  // high halves: records
                k_superk_bins<<<(nb_sk + 1 + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(ctx->d_sk_prefix.as<uint64_t>(), n_micro, ctx->d_fine_off.as<uint64_t>(), nb_sk, sk_target,
                                                                                 ctx->d_sk_rprefix.as<uint64_t>(), ctx->d_sk_recoff.as<uint64_t>());
                CK(cudaEventRecord(ev[2 * T_PART_COUNT + 1], st));
                CK(cudaEventRecord(ev[2 * T_PART_SCATTER], st));
                k_sk2_scatter<<<(unsigned)std::min<uint64_t>((rec_cap + BLOCK - 1) / BLOCK, (uint64_t)ctx->nsm * 16), BLOCK, 0, st>>>(
                    ctx->d_keysA.as<SkRecord>(), d_nrec, rec_cap, ctx->d_sk_rprefix.as<uint64_t>(), ctx->d_fine_cursor.as<uint32_t>(), ctx->d_keysB.as<SkRecord>());
                CK(cudaEventRecord(ev[2 * T_PART_SCATTER + 1], st));
                ctx->launches += 3;
            } else {
            launch_scan(ctx, ctx->d_fine_hist.as<uint32_t>(), ctx->d_sk_prefix.as<uint64_t>(), n_micro);
            k_superk_bins<<<(nb_sk + 1 + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(ctx->d_sk_prefix.as<uint64_t>(), n_micro, ctx->d_fine_off.as<uint64_t>(), nb_sk, sk_target, nullptr, nullptr);
            CK(cudaEventRecord(ev[2 * T_PART_COUNT + 1], st));
            CK(cudaEventRecord(ev[2 * T_PART_SCATTER], st));
            // default: flat (work per k-mer); SIMKA_B200_SUPERK_EXPAND=thread: one thread rolls a whole record
            (ctx->superk_expand_thread ? k_superk_expand : k_superk_expand_flat)<<<(unsigned)std::min<uint64_t>((rec_cap + BLOCK - 1) / BLOCK, (uint64_t)ctx->nsm * 16), BLOCK, 0, st>>>(
                ctx->d_keysA.as<SkRecord>(), d_nrec, rec_cap, ctx->kp, ctx->d_sk_prefix.as<uint64_t>(), ctx->d_fine_cursor.as<uint32_t>(),
                ctx->d_keysB.as<uint64_t>());
            CK(cudaEventRecord(ev[2 * T_PART_SCATTER + 1], st));
            ctx->launches += 3;
            }
        } else {
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
        (ctx->k2_minb3 ? k_partition2<false, 3> : k_partition2<false, 4>)<<<g2, BLOCK, 0, st>>>(ctx->d_keysA.as<uint64_t>(), ctx->d_base1.as<uint64_t>(), cnt1, ctx->d_tile_start.as<uint64_t>(), b1_bits, b2_bits, W,
                                                  nullptr, ctx->d_fine_off.as<uint64_t>(), ctx->d_fine_cursor.as<uint32_t>(), ctx->d_keysB.as<uint64_t>());
        CK(cudaEventRecord(ev[2 * T_PART_SCATTER + 1], st));
        ctx->launches += exact ? 4 : 5;  // (the scans count themselves)
        }
        // K3 count
        CountOut co;
        co.stage_h = ctx->d_keysA.as<uint64_t>();
        co.stage_c = ctx->d_stage_c.as<uint32_t>();
        co.surv = ctx->d_surv.as<uint32_t>();
        co.totals = d_totals;
        co.error_flag = d_error;
        if (ctx->occ_bin_count <= 0) {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_bin_count, k_bin_count, K3_BLOCK, 0));
            if (ctx->occ_bin_count <= 0) ctx->occ_bin_count = 1;
        }
        const unsigned gbin = (unsigned)std::min<uint64_t>(nbins_run, (uint64_t)ctx->nsm * ctx->occ_bin_count);
        // table slot: the bits just below the bin bits (super-k-mer mode: bins are not key ranges, take the top bits)
        const int slot_shift = std::max(0, W - (sk ? 0 : bin_bits) - ilog2_floor(CNT_SLOTS));
        CK(cudaEventRecord(ev[2 * T_BIN_COUNT], st));
        if (sk && sk2) {
            if (ctx->occ_sk2_count <= 0) {
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_sk2_count, k_sk2_count<false>, SK2_THREADS, sizeof(Sk2Smem)));
                if (ctx->occ_sk2_count <= 0) ctx->occ_sk2_count = 1;
            }
            Sk2Out so;
            so.stage_h = co.stage_h; so.stage_c = co.stage_c; so.surv = co.surv; so.totals = co.totals; so.error_flag = co.error_flag;
            const unsigned g2c = (unsigned)std::min<uint64_t>(nbins_run, (uint64_t)ctx->nsm * ctx->occ_sk2_count);
            (ctx->kp.shard_count > 1 ? k_sk2_count<true> : k_sk2_count<false>)<<<g2c, SK2_THREADS, sizeof(Sk2Smem), st>>>(ctx->d_keysB.as<SkRecord>(), ctx->d_sk_recoff.as<uint64_t>(), ctx->d_fine_off.as<uint64_t>(), nbins_run,
                                                                 ctx->kp, ctx->prm.abundance_min, ctx->prm.abundance_max, so);
        } else
        k_bin_count<<<gbin, K3_BLOCK, 0, st>>>(ctx->d_keysB.as<uint64_t>(), ctx->d_fine_off.as<uint64_t>(), nbins_run, slot_shift, ctx->prm.abundance_min,
                                            ctx->prm.abundance_max, co);
        CK(cudaEventRecord(ev[2 * T_BIN_COUNT + 1], st));
        launch_scan(ctx, ctx->d_surv.as<uint32_t>(), ctx->d_out_off.as<uint64_t>(), nbins_run);
        ctx->launches += 1;
        CK(cudaGetLastError());
        // the only host sync of the sample: sizes + totals
        int h_err = 0, h_overflow = 0;
        CK(cudaMemcpyAsync(&h_surv, ctx->d_out_off.as<uint64_t>() + nbins_run, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_keys, ctx->d_fine_off.as<uint64_t>() + nbins_run, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_tot, d_totals, 24, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_err, d_error, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_overflow, d_overflow, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_overflow && sk) { ctx->superk_fallbacks++; continue; }  // more runs than the record list holds: redo on the k-mer path
        if (h_overflow && !exact) { ctx->exact_retries++; continue; }  // a bucket outgrew its capacity: redo with the exact two-pass path
        sk_used = sk;
        nbins_fin = nbins_run;
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
            // super-k-mer mode: the bins are minimizer bins, so the survivors are first packed into scratch (the key buffer K3
            // has finished reading) and then counting-sorted by fine range of the key into the sample's list
            uint64_t *pack_h = sd.eh;
            uint32_t *pack_c = sd.ec;
            if (sk_used) {
                if (ctx->d_sk_tmpc.reserve(h_surv * 4)) { ctx->set_error("out of device memory (work buffers)"); return SIMKA_ERR_NOMEM; }
                pack_h = ctx->d_keysB.as<uint64_t>();
                pack_c = ctx->d_sk_tmpc.as<uint32_t>();
            }
            k_compact<<<(unsigned)std::min<uint64_t>((nbins_fin + 7) / 8, (uint64_t)ctx->nsm * 8), BLOCK, 0, st>>>(ctx->d_keysA.as<uint64_t>(), ctx->d_stage_c.as<uint32_t>(), ctx->d_fine_off.as<uint64_t>(),
                                             ctx->d_out_off.as<uint64_t>(), nbins_fin, pack_h, pack_c);
            ctx->launches++;
            if (sk_used) {
                const size_t fgn1 = (size_t)1 << ctx->fg_bits;
                if (ctx->d_fine_hist.reserve(fgn1 * 4) || ctx->d_fine_cursor.reserve(fgn1 * 4) || ctx->d_fine_off.reserve((fgn1 + 1) * 8)) {
                    ctx->set_error("out of device memory (work buffers)");
                    return SIMKA_ERR_NOMEM;
                }
                CK(cudaMemsetAsync(ctx->d_fine_hist.p, 0, fgn1 * 4, st));
                CK(cudaMemsetAsync(ctx->d_fine_cursor.p, 0, fgn1 * 4, st));
                const unsigned g = (unsigned)std::min<uint64_t>((h_surv + BLOCK - 1) / BLOCK, (uint64_t)ctx->nsm * 16);
                const int shift = W - ctx->fg_bits;
                k_refine_hist<<<g, BLOCK, 0, st>>>(pack_h, h_surv, shift, ctx->d_fine_hist.as<uint32_t>());
                launch_scan(ctx, ctx->d_fine_hist.as<uint32_t>(), ctx->d_fine_off.as<uint64_t>(), fgn1);
                k_refine_scatter<<<g, BLOCK, 0, st>>>(pack_h, pack_c, h_surv, shift, ctx->d_fine_off.as<uint64_t>(), ctx->d_fine_cursor.as<uint32_t>(), sd.eh, sd.ec);
                ctx->launches += 2;
            }
            if (ctx->prm.complex_dist) {
                const unsigned gh = (unsigned)std::min<uint64_t>((h_surv + BLOCK * 16 - 1) / (BLOCK * 16), (uint64_t)ctx->nsm * 4);
                k_abundance_hist<<<gh, BLOCK, 0, st>>>(sd.ec, h_surv, sample_id, ctx->d_chist.as<ull>() + (size_t)sample_id * CHIST,
                                                      ctx->d_big_counts.as<uint32_t>(), ctx->d_big_sample.as<uint32_t>(), d_big_n, ctx->big_cap, d_error);
                ctx->launches++;
            }
        }
        const uint32_t fg = 1u << ctx->fg_bits;
        if (sk_used) {
            // offsets of the fine ranges: the scan of the regrouping above (all zero for a sample without solid k-mers)
            if (h_surv > 0) k_fill_offsets<<<(fg + 1 + 255) / 256, 256, 0, st>>>(ctx->d_fine_off.as<uint64_t>(), sd.off, fg, 1);
            else CK(cudaMemsetAsync(sd.off, 0, ((size_t)fg + 1) * 4, st));
            ctx->superk_samples++;
        } else {
            k_fill_offsets<<<(fg + 1 + 255) / 256, 256, 0, st>>>(ctx->d_out_off.as<uint64_t>(), sd.off, fg, nbins >> ctx->fg_bits);
        }
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
    if (ctx->comm && ctx->comm_world > 1) {  // multi-GPU: this sample's round starts crossing NVLink while the next one is counted
        const int rc = xchg_after_sample(ctx, sample_id);
        if (rc != SIMKA_OK) return rc;
    }
    return SIMKA_OK;
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
}

// stage 1 for k > 31 (kernels/wide.cuh): count pass, scan, scatter pass, per-bin count, compaction
static int finish_sample_wide(simka_ctx *ctx, int sample_id, simka_sample_totals *out) {
    CK(cudaSetDevice(ctx->prm.device));
    cudaStream_t st = ctx->stream;
    SampleData &sd = ctx->samples[sample_id];
    InSlot *sl = nullptr;
    for (auto &s : ctx->slots) if (s.sample == sample_id) sl = &s;
    const uint64_t n_words = sl ? sl->words : 0, n_reads = sl ? sl->reads : 0, n_pos = n_words * 32;
    const int W = 2 * ctx->prm.kmer_size;
    if (ctx->fg_bits < 0) {
        const uint64_t per_sample = ctx->prm.expected_kmers_per_sample ? ctx->prm.expected_kmers_per_sample : n_pos;
        int fg = ilog2_ceil(std::max<uint64_t>(1, (per_sample / 4) * (uint64_t)ctx->prm.nb_samples / MRG_TARGET));
        fg = std::min(fg, std::min(20, ilog2_floor(std::max<uint64_t>(1, per_sample / BIN_TARGET))));
        while (fg > 0 && ((size_t)ctx->prm.nb_samples * (((size_t)1 << fg) + 1) * 4) > ((size_t)2 << 30)) fg--;
        ctx->fg_bits = fg;
        const size_t tbl = (size_t)ctx->prm.nb_samples * (((size_t)1 << fg) + 1) * 4;
        if (ctx->d_off_table.reserve(tbl)) { ctx->set_error("out of device memory (offset table)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemsetAsync(ctx->d_off_table.p, 0, tbl, st));
    }
    int bin_bits = std::max(ilog2_floor(std::max<uint64_t>(1, n_pos / BIN_TARGET)), ctx->fg_bits);
    bin_bits = std::max(std::min(bin_bits, std::min(W, 20)), ctx->fg_bits);
    const uint32_t nbins = 1u << bin_bits;
    const size_t fgn = ((size_t)1 << ctx->fg_bits) + 1;
    sd.off = ctx->d_off_table.as<uint32_t>() + (size_t)sample_id * fgn;
    ull *d_totals = ctx->d_small.as<ull>();
    int *d_error = reinterpret_cast<int *>(ctx->d_small.as<char>() + 32);
    uint32_t *d_big_n = reinterpret_cast<uint32_t *>(ctx->d_small.as<char>() + 40);
    uint64_t h_keys = 0, h_surv = 0;
    ull h_tot[3] = {0, 0, 0};
    if (n_words > 0) {
        const size_t nres = std::max<size_t>(nbins, (size_t)1 << ctx->fg_bits);
        if (ctx->d_keysA.reserve(n_pos * 16) || ctx->d_keysB.reserve(n_pos * 16) || ctx->d_stage_c.reserve(n_pos * 4) || ctx->d_fine_hist.reserve(nres * 4) ||
            ctx->d_fine_off.reserve((nres + 1) * 8) || ctx->d_fine_cursor.reserve(nres * 4) || ctx->d_surv.reserve(nres * 4) || ctx->d_out_off.reserve((nres + 1) * 8)) {
            ctx->set_error("out of device memory (work buffers)");
            return SIMKA_ERR_NOMEM;
        }
        CK(cudaStreamWaitEvent(st, sl->copied, 0));
        const int evset = ctx->ev_cur;
        harvest_timings(ctx, evset);
        cudaEvent_t *ev = ctx->ev[evset];
        CK(cudaEventRecord(ev[2 * T_NSTAGE], st));
        CK(cudaMemsetAsync(d_totals, 0, 24, st));
        CK(cudaMemsetAsync(ctx->d_fine_hist.p, 0, (size_t)nbins * 4, st));
        CK(cudaMemsetAsync(ctx->d_fine_cursor.p, 0, (size_t)nbins * 4, st));
        const uint64_t *seq = sl->ext_packed ? sl->ext_packed : sl->packed.as<uint64_t>();
        const uint32_t *bad = sl->ext_bad ? sl->ext_bad : sl->bad.as<uint32_t>();
        const unsigned g1 = (unsigned)((n_words + BLOCK - 1) / BLOCK);
        CK(cudaEventRecord(ev[2 * T_EXTRACT_COUNT], st));
        k_wide_extract<true><<<g1, BLOCK, 0, st>>>(seq, bad, n_words, ctx->prm.kmer_size, ctx->mix128, bin_bits, ctx->d_fine_hist.as<uint32_t>(), nullptr, nullptr, nullptr);
        CK(cudaEventRecord(ev[2 * T_EXTRACT_COUNT + 1], st));
        launch_scan(ctx, ctx->d_fine_hist.as<uint32_t>(), ctx->d_fine_off.as<uint64_t>(), nbins);
        CK(cudaEventRecord(ev[2 * T_EXTRACT_SCATTER], st));
        k_wide_extract<false><<<g1, BLOCK, 0, st>>>(seq, bad, n_words, ctx->prm.kmer_size, ctx->mix128, bin_bits, nullptr, ctx->d_fine_off.as<uint64_t>(),
                                                   ctx->d_fine_cursor.as<uint32_t>(), ctx->d_keysA.as<K128>());
        CK(cudaEventRecord(ev[2 * T_EXTRACT_SCATTER + 1], st));
        CK(cudaEventRecord(ev[2 * T_PART_COUNT], st)); CK(cudaEventRecord(ev[2 * T_PART_COUNT + 1], st));
        CK(cudaEventRecord(ev[2 * T_PART_SCATTER], st)); CK(cudaEventRecord(ev[2 * T_PART_SCATTER + 1], st));
        WideOut wo;
        wo.stage_h = ctx->d_keysB.as<K128>(); wo.stage_c = ctx->d_stage_c.as<uint32_t>(); wo.surv = ctx->d_surv.as<uint32_t>();
        wo.totals = d_totals; wo.error_flag = d_error;
        CK(cudaEventRecord(ev[2 * T_BIN_COUNT], st));
        k_wide_bin_count<<<(unsigned)std::min<uint64_t>(nbins, (uint64_t)ctx->nsm * 4), BLOCK, 0, st>>>(ctx->d_keysA.as<K128>(), ctx->d_fine_off.as<uint64_t>(), nbins,
                                                                                                   ctx->prm.abundance_min, ctx->prm.abundance_max, wo);
        CK(cudaEventRecord(ev[2 * T_BIN_COUNT + 1], st));
        launch_scan(ctx, ctx->d_surv.as<uint32_t>(), ctx->d_out_off.as<uint64_t>(), nbins);
        ctx->launches += 3;
        CK(cudaGetLastError());
        int h_err = 0;
        CK(cudaMemcpyAsync(&h_surv, ctx->d_out_off.as<uint64_t>() + nbins, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_keys, ctx->d_fine_off.as<uint64_t>() + nbins, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_tot, d_totals, 24, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h_err, d_error, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_err) { ctx->set_error("on-chip table overflow in k_wide_bin_count (code " + std::to_string(h_err) + ")"); return SIMKA_ERR_OVERFLOW; }
        CK(cudaEventRecord(sl->consumed, st));
        sl->consumed_valid = true;
        if (h_surv >= (1ull << 32)) { ctx->set_error("more than 2^32 solid k-mers in one sample"); return SIMKA_ERR_INVALID; }
        CK(cudaEventRecord(ev[2 * T_COMPACT], st));
        if (h_surv > 0) {
            sd.ehw = (K128 *)ctx->arena.alloc(h_surv * 16);
            sd.ec = (uint32_t *)ctx->arena.alloc(h_surv * 4);
            if (!sd.ehw || !sd.ec) { ctx->set_error("out of device memory (sample list)"); return SIMKA_ERR_NOMEM; }
            k_wide_compact<<<(unsigned)std::min<uint64_t>((nbins + 7) / 8, (uint64_t)ctx->nsm * 8), BLOCK, 0, st>>>(ctx->d_keysB.as<K128>(), ctx->d_stage_c.as<uint32_t>(),
                                                                                                              ctx->d_fine_off.as<uint64_t>(), ctx->d_out_off.as<uint64_t>(), nbins, sd.ehw, sd.ec);
            ctx->launches++;
            if (ctx->prm.complex_dist) {
                const unsigned gh = (unsigned)std::min<uint64_t>((h_surv + BLOCK * 16 - 1) / (BLOCK * 16), (uint64_t)ctx->nsm * 4);
                k_abundance_hist<<<gh, BLOCK, 0, st>>>(sd.ec, h_surv, sample_id, ctx->d_chist.as<ull>() + (size_t)sample_id * CHIST, ctx->d_big_counts.as<uint32_t>(),
                                                      ctx->d_big_sample.as<uint32_t>(), d_big_n, ctx->big_cap, d_error);
                ctx->launches++;
            }
        }
        const uint32_t fg = 1u << ctx->fg_bits;
        k_fill_offsets<<<(fg + 1 + 255) / 256, 256, 0, st>>>(ctx->d_out_off.as<uint64_t>(), sd.off, fg, nbins >> ctx->fg_bits);
        ctx->launches++;
        CK(cudaEventRecord(ev[2 * T_COMPACT + 1], st));
        CK(cudaEventRecord(ev[2 * T_NSTAGE + 1], st));
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

// k > 31: the per-sample (canonical k-mer, count) list as (low word, high word) pairs, ascending as 128-bit integers
int simka_gpu_dump_sample_counts128(simka_ctx *ctx, int sample_id, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *counts, uint64_t *n) try {
    if (!ctx || !n || !ctx->wide || sample_id < 0 || sample_id >= ctx->prm.nb_samples || !ctx->samples[sample_id].done) return SIMKA_ERR_INVALID;
    CK(cudaSetDevice(ctx->prm.device));
    SampleData &sd = ctx->samples[sample_id];
    if (!kmers_lo || !kmers_hi || !counts) { *n = sd.n; return SIMKA_OK; }
    if (*n < sd.n) { *n = sd.n; return SIMKA_ERR_INVALID; }
    *n = sd.n;
    if (sd.n == 0) return SIMKA_OK;
    std::vector<uint64_t> lo(sd.n), hi(sd.n);
    std::vector<uint32_t> hc(sd.n);
    if (ctx->d_dump.reserve(sd.n * 16)) { ctx->set_error("out of device memory (dump)"); return SIMKA_ERR_NOMEM; }
    uint64_t *d_lo = ctx->d_dump.as<uint64_t>(), *d_hi = d_lo + sd.n;
    k_wide_unmix<<<(unsigned)((sd.n + 255) / 256), 256, 0, ctx->stream>>>(sd.ehw, d_lo, d_hi, sd.n, ctx->mix128);
    CK(cudaMemcpyAsync(lo.data(), d_lo, sd.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hi.data(), d_hi, sd.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hc.data(), sd.ec, sd.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> order(sd.n);
    for (uint64_t i = 0; i < sd.n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return hi[a] != hi[b] ? hi[a] < hi[b] : lo[a] < lo[b]; });
    for (uint64_t i = 0; i < sd.n; i++) { kmers_lo[i] = lo[order[i]]; kmers_hi[i] = hi[order[i]]; counts[i] = hc[order[i]]; }
    return SIMKA_OK;
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
}

int simka_gpu_dump_sample_counts(simka_ctx *ctx, int sample_id, uint64_t *kmers, uint32_t *counts, uint64_t *n) try {
    if (ctx && ctx->wide) { ctx->set_error("k > 31: use simka_gpu_dump_sample_counts128"); return SIMKA_ERR_INVALID; }
    if (!ctx || !n || sample_id < 0 || sample_id >= ctx->prm.nb_samples || !ctx->samples[sample_id].done) return SIMKA_ERR_INVALID;
    CK(cudaSetDevice(ctx->prm.device));
    SampleData &sd = ctx->samples[sample_id];
    if (!kmers || !counts) { *n = sd.n; return SIMKA_OK; }
    if (*n < sd.n) { *n = sd.n; return SIMKA_ERR_INVALID; }
    *n = sd.n;
    if (sd.n == 0) return SIMKA_OK;
    std::vector<uint64_t> hk(sd.n);
    std::vector<uint32_t> hc(sd.n);
    if (ctx->d_dump.reserve(sd.n * 8)) { ctx->set_error("out of device memory (dump)"); return SIMKA_ERR_NOMEM; }  // context-owned: no leak on an error return
    uint64_t *tmp = ctx->d_dump.as<uint64_t>();
    if (ctx->kp.key_mode) CK(cudaMemcpyAsync(tmp, sd.eh, sd.n * 8, cudaMemcpyDeviceToDevice, ctx->stream));  // the keys ARE the hashes
    else k_unmix<<<(unsigned)((sd.n + 255) / 256), 256, 0, ctx->stream>>>(sd.eh, tmp, sd.n, ctx->kp.mix);
    CK(cudaMemcpyAsync(hk.data(), tmp, sd.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hc.data(), sd.ec, sd.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> order(sd.n);
    for (uint64_t i = 0; i < sd.n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return hk[a] < hk[b]; });
    for (uint64_t i = 0; i < sd.n; i++) { kmers[i] = hk[order[i]]; counts[i] = hc[order[i]]; }
    return SIMKA_OK;
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
}

int simka_gpu_fine_bits(simka_ctx *ctx) { return ctx ? ctx->fg_bits : -1; }

int simka_gpu_export_sample(simka_ctx *ctx, int sample_id, const uint64_t **d_kmers, const uint32_t **d_counts, const uint32_t **d_off, uint64_t *n) {
    if (ctx && ctx->wide) { ctx->set_error("k > 31: lists of 128-bit k-mers are not exported"); return SIMKA_ERR_INVALID; }
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
    if (ctx->wide) { ctx->set_error("k > 31: lists of 128-bit k-mers are not imported"); return SIMKA_ERR_INVALID; }
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

// Registers a sample from its solid (canonical k-mer, count) list in HOST memory, e.g. the decoded
// solid/part_P/__p__I.gz records of a reference-format temp dir: hashes the k-mers, groups them by fine range and
// uploads them; after this the sample is exactly as if simka_gpu_finish_sample had produced it.
int simka_gpu_import_counts(simka_ctx *ctx, int sample_id, const uint64_t *kmers, const uint32_t *counts, uint64_t n, uint64_t nb_reads) try {
    if (!ctx) return SIMKA_ERR_INVALID;
    if (ctx->wide) { ctx->set_error("k > 31: lists of 128-bit k-mers are not imported"); return SIMKA_ERR_INVALID; }
    if (sample_id < 0 || sample_id >= ctx->prm.nb_samples || ctx->samples[sample_id].done || (n && (!kmers || !counts)) || n >= (1ull << 32)) {
        ctx->set_error("bad import arguments");
        return SIMKA_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->prm.device));
    cudaStream_t st = ctx->stream;
    const int W = ctx->W;
    if (ctx->fg_bits < 0) {
        const uint64_t per_sample = ctx->prm.expected_kmers_per_sample ? ctx->prm.expected_kmers_per_sample / 4 : n;
        int fg = ilog2_ceil(std::max<uint64_t>(1, per_sample * (uint64_t)ctx->prm.nb_samples / MRG_TARGET));
        fg = std::min(fg, std::min(W, 20));
        while (fg > 0 && ((size_t)ctx->prm.nb_samples * (((size_t)1 << fg) + 1) * 4) > ((size_t)2 << 30)) fg--;
        ctx->fg_bits = fg;
        const size_t tbl = (size_t)ctx->prm.nb_samples * (((size_t)1 << fg) + 1) * 4;
        if (ctx->d_off_table.reserve(tbl)) { ctx->set_error("out of device memory (offset table)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemsetAsync(ctx->d_off_table.p, 0, tbl, st));
    }
    const int fg = ctx->fg_bits;
    const size_t fgn = ((size_t)1 << fg) + 1;
    // counting sort by fine range on the host
    std::vector<uint32_t> off(fgn, 0);
    std::vector<uint64_t> h(n);
    const uint64_t kmask = ctx->kp.mix.mask;
    ull sum = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (kmers[i] > kmask || counts[i] == 0) { ctx->set_error("k-mer out of range or zero count in the imported list"); return SIMKA_ERR_INVALID; }
        h[i] = simka::mix_fwd(kmers[i], ctx->kp.mix);
        off[(size_t)(fg ? (h[i] >> (W - fg)) : 0) + 1]++;
        sum += counts[i];
    }
    for (size_t f = 0; f + 1 < fgn; f++) off[f + 1] += off[f];
    std::vector<uint32_t> cur(off.begin(), off.end() - 1);
    std::vector<uint64_t> eh(n);
    std::vector<uint32_t> ec(n);
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t pos = cur[(size_t)(fg ? (h[i] >> (W - fg)) : 0)]++;
        eh[pos] = h[i];
        ec[pos] = counts[i];
    }
    SampleData &sd = ctx->samples[sample_id];
    sd.off = ctx->d_off_table.as<uint32_t>() + (size_t)sample_id * fgn;
    CK(cudaMemcpyAsync(sd.off, off.data(), fgn * 4, cudaMemcpyHostToDevice, st));
    if (n) {
        sd.eh = (uint64_t *)ctx->arena.alloc(n * 8);
        sd.ec = (uint32_t *)ctx->arena.alloc(n * 4);
        if (!sd.eh || !sd.ec) { ctx->set_error("out of device memory (sample list)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemcpyAsync(sd.eh, eh.data(), n * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(sd.ec, ec.data(), n * 4, cudaMemcpyHostToDevice, st));
        if (ctx->prm.complex_dist) {
            int *d_error = reinterpret_cast<int *>(ctx->d_small.as<char>() + 32);
            uint32_t *d_big_n = reinterpret_cast<uint32_t *>(ctx->d_small.as<char>() + 40);
            const unsigned g = (unsigned)std::min<uint64_t>((n + BLOCK * 16 - 1) / (BLOCK * 16), (uint64_t)ctx->nsm * 4);
            k_abundance_hist<<<g, BLOCK, 0, st>>>(sd.ec, n, sample_id, ctx->d_chist.as<ull>() + (size_t)sample_id * CHIST,
                                                 ctx->d_big_counts.as<uint32_t>(), ctx->d_big_sample.as<uint32_t>(), d_big_n, ctx->big_cap, d_error);
            ctx->launches++;
        }
    }
    CK(cudaStreamSynchronize(st));  // the host staging vectors go out of scope
    sd.n = n;
    sd.done = true;
    sd.local.nb_reads = nb_reads;
    sd.local.nb_distinct = n;
    sd.local.nb_kmers = sum;
    sd.local.chord_n2 = 0;
    for (uint64_t i = 0; i < n; i++) sd.local.chord_n2 += (uint64_t)((double)counts[i] * (double)counts[i]);
    if (!ctx->totals_overridden) ctx->global_totals[sample_id] = sd.local;
    ctx->entries += n;
    return SIMKA_OK;
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
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

// stage 2 for 64 < N <= 512 (kernels/merge_tiles.cuh): records of the shared k-mers (sample bitmap + abundances in sample order) are written
// once per merge bin into the idle stage-1 work buffers, then one CTA per tile of 32 x 32 samples accumulates its pairs in shared
// memory.  The bins are processed in batches that fit the scratch.
static int merge_by_tiles(simka_ctx *ctx, const MergeArgs &a, uint32_t MG) {
    cudaStream_t st = ctx->stream;
    const int N = a.n, nblk = (N + 31) / 32;
    const uint32_t q = 1u << (a.fg_bits - a.mg_bits);
    const size_t ostride = ((size_t)1 << a.fg_bits) + 1;
    // entries per bin and their prefix (host copy: the batches are cut here)
    if (ctx->d_fine_hist.reserve((size_t)MG * 4) || ctx->d_fine_off.reserve(((size_t)MG + 1) * 8)) { ctx->set_error("out of device memory (tile merge)"); return SIMKA_ERR_NOMEM; }
    k_bin_entries<<<(MG + BLOCK / 32 - 1) / (BLOCK / 32), BLOCK, 0, st>>>(a.off, N, MG, q, ostride, ctx->d_fine_hist.as<uint32_t>());
    ctx->launches++;
    if (launch_scan(ctx, ctx->d_fine_hist.as<uint32_t>(), ctx->d_fine_off.as<uint64_t>(), MG)) { ctx->set_error("out of device memory (tile merge)"); return SIMKA_ERR_NOMEM; }
    std::vector<uint64_t> ebin((size_t)MG + 1);
    CK(cudaMemcpyAsync(ebin.data(), ctx->d_fine_off.p, ((size_t)MG + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // scratch: abundances (4 B per entry of the batch) in keysA, records ((1 + nblk) words per 2 entries at most) in keysB
    const size_t rec_words = 1 + (size_t)nblk;
    uint64_t max_bin = 0;
    for (uint32_t b = 0; b < MG; b++) max_bin = std::max(max_bin, ebin[b + 1] - ebin[b]);
    uint64_t cap_e = std::min<uint64_t>(ctx->d_keysA.cap / 4, (ctx->d_keysB.cap / (4 * rec_words)) * 2);
    const uint64_t want = std::max<uint64_t>(std::min<uint64_t>(ebin[MG], (uint64_t)64 << 20), max_bin + 2);
    if (cap_e < want) {
        if (ctx->d_keysA.reserve(want * 4) || ctx->d_keysB.reserve((want / 2 + 2) * rec_words * 4)) { ctx->set_error("out of device memory (tile merge scratch)"); return SIMKA_ERR_NOMEM; }
        cap_e = std::min<uint64_t>(ctx->d_keysA.cap / 4, (ctx->d_keysB.cap / (4 * rec_words)) * 2);
    }
    cap_e = std::min<uint64_t>(cap_e - 2, 0xFFFFFFF0ull);  // (abundance offsets inside a batch are 32-bit)
    { const char *e = getenv("SIMKA_B200_TILE_BATCH"); if (e) cap_e = std::min<uint64_t>(cap_e, std::max<uint64_t>(max_bin, strtoull(e, nullptr, 10))); }  // tests: force several batches
    TileArgs t;
    t.m = a;
    t.ebin = ctx->d_fine_off.as<uint64_t>();
    t.cnt = ctx->d_keysA.as<uint32_t>();
    t.grp = ctx->d_keysB.as<uint32_t>();
    t.nblk = nblk;
    t.rec_cap = (uint32_t)std::min<uint64_t>(ctx->d_keysB.cap / (4 * rec_words), 0xFFFFFFFFull);
    void (*kern)(TileArgs, uint32_t) = a.simple ? (a.complex_ ? k_merge_tiles<true, true> : k_merge_tiles<true, false>)
                                                : (a.complex_ ? k_merge_tiles<false, true> : k_merge_tiles<false, false>);
    const size_t smem_g = group_bins_smem_bytes(N), smem_t = merge_tiles_smem_bytes();
    CK(cudaFuncSetAttribute(k_group_bins, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    int occ_g = 0, occ_t = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g, k_group_bins, BLOCK, smem_g));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, kern, TM_THREADS, smem_t));
    // partition scratch of k_group_bins (bins of more than MRG_CAP entries): 16 bytes per entry of the largest bin and CTA, in the idle
    // stage-1 count buffer; without it (no memory left) every sub-pass reads its whole bin
    const unsigned gg_max = (unsigned)std::min<uint64_t>(MG, (uint64_t)ctx->nsm * std::max(1, occ_g));
    t.scr_h = nullptr; t.scr_s = t.scr_c = nullptr; t.scr_cap = 0;
    if (max_bin > (uint64_t)MRG_CAP && max_bin <= (uint64_t)TM_RMAX * (MRG_CAP / 2) && !getenv("SIMKA_B200_TILE_NOPART")) {
        const size_t cap = (size_t)((max_bin + 63) & ~(uint64_t)63);
        if (ctx->d_stage_c.reserve((size_t)gg_max * cap * 16) == 0) {
            t.scr_h = ctx->d_stage_c.as<uint64_t>();
            t.scr_s = reinterpret_cast<uint32_t *>(t.scr_h + (size_t)gg_max * cap);
            t.scr_c = t.scr_s + (size_t)gg_max * cap;
            t.scr_cap = (uint32_t)cap;
        } else {
            cudaGetLastError();
        }
    }
    const uint32_t tiles = (uint32_t)(nblk * (nblk + 1) / 2);
    const uint32_t slices = std::max<uint32_t>(1, (uint32_t)(ctx->nsm * std::max(1, occ_t)) / tiles);
    const bool timing = getenv("SIMKA_B200_TILE_TIMING") != nullptr;
    std::vector<cudaEvent_t> evs;
    for (uint32_t b0 = 0; b0 < MG;) {
        // largest b1 with ebin[b1] - ebin[b0] <= cap_e (at least one bin: every bin fits by construction)
        uint32_t b1 = (uint32_t)(std::upper_bound(ebin.begin() + b0, ebin.end(), ebin[b0] + cap_e) - ebin.begin()) - 1;
        b1 = std::max(b1, b0 + 1);
        if (ebin[b1] > ebin[b0]) {
            t.bin0 = b0; t.bin1 = b1;
            if (ctx->d_surv.reserve(64)) { ctx->set_error("out of device memory (tile merge)"); return SIMKA_ERR_NOMEM; }
            t.hdr = ctx->d_surv.as<uint32_t>();
            CK(cudaMemsetAsync(t.hdr, 0, 4, st));
            const unsigned gg = (unsigned)std::min<uint64_t>(b1 - b0, gg_max);
            if (timing) { evs.resize(evs.size() + 3); for (int i = 0; i < 3; i++) cudaEventCreate(&evs[evs.size() - 3 + i]); cudaEventRecord(evs[evs.size() - 3], st); }
            k_group_bins<<<gg, BLOCK, smem_g, st>>>(t);
            if (timing) cudaEventRecord(evs[evs.size() - 2], st);
            kern<<<tiles * slices, TM_THREADS, smem_t, st>>>(t, slices);
            if (timing) cudaEventRecord(evs[evs.size() - 1], st);
            ctx->launches += 2;
        }
        b0 = b1;
    }
    CK(cudaGetLastError());
    if (timing) {  // developer aid (SIMKA_B200_TILE_TIMING=1): where the merge's time goes, on stderr
        CK(cudaStreamSynchronize(st));
        float tg = 0, tt = 0, ms = 0;
        for (size_t i = 0; i + 2 < evs.size(); i += 3) {
            cudaEventElapsedTime(&ms, evs[i], evs[i + 1]); tg += ms;
            cudaEventElapsedTime(&ms, evs[i + 1], evs[i + 2]); tt += ms;
        }
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
        fprintf(stderr, "[simka_b200] tile merge: N=%d bins=%u entries=%llu largest bin=%llu (%s) batches=%zu slices=%u k_group_bins %.1f ms, k_merge_tiles %.1f ms\n", N, MG,
                (unsigned long long)ebin[MG], (unsigned long long)max_bin, t.scr_h ? "partitioned" : "not partitioned", evs.size() / 3, slices, tg, tt);
    }
    return SIMKA_OK;
}

int simka_gpu_merge_and_accumulate(simka_ctx *ctx) try {
    if (!ctx) return SIMKA_ERR_INVALID;
    const int N = ctx->prm.nb_samples;
    for (int s = 0; s < N; s++) if (!ctx->samples[s].done) { ctx->set_error("all samples must be finished before the merge"); return SIMKA_ERR_INVALID; }
    if (ctx->merged) { ctx->set_error("merge already done"); return SIMKA_ERR_INVALID; }
    if (ctx->kp.key_mode) { ctx->set_error("a sketch-mode context has no stage 2"); return SIMKA_ERR_INVALID; }
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
    // merge bins of about MRG_TARGET entries.  The lists are grouped by the top fg_bits of the key, which stage 1 cannot make
    // finer than the counting bins of one sample: when that leaves more entries per merge bin than the on-chip tables hold
    // (many samples), the lists are first regrouped by more bits (k_refine_*), once, instead of being re-read per sub-pass.
    const int want_bits = ilog2_ceil(std::max<uint64_t>(1, total_entries / MRG_TARGET));
    int fg_eff = ctx->fg_bits;
    const uint32_t *off_eff = ctx->d_off_table.as<uint32_t>();
    if (ctx->fg_bits >= 0 && (total_entries >> ctx->fg_bits) > (uint64_t)MRG_CAP && want_bits > ctx->fg_bits && !ctx->no_refine && !ctx->wide) {
        int fg2 = std::min(want_bits, std::min(ctx->W, 24));
        while (fg2 > ctx->fg_bits && (size_t)N * (((size_t)1 << fg2) + 1) * 4 > ((size_t)4 << 30)) fg2--;
        if (fg2 > ctx->fg_bits) {
            const size_t nb2 = (size_t)1 << fg2, row = nb2 + 1;
            uint64_t max_n = 0;
            for (int s = 0; s < N; s++) max_n = std::max<uint64_t>(max_n, ctx->samples[s].n);
            if (ctx->d_off_table2.reserve((size_t)N * row * 4) || ctx->d_fine_hist.reserve(nb2 * 4) || ctx->d_fine_cursor.reserve(nb2 * 4) ||
                ctx->d_fine_off.reserve(row * 8) || ctx->d_keysA.reserve(max_n * 8) || ctx->d_stage_c.reserve(max_n * 4)) {
                ctx->set_error("out of device memory (merge refinement)");
                return SIMKA_ERR_NOMEM;
            }
            CK(cudaMemsetAsync(ctx->d_off_table2.p, 0, (size_t)N * row * 4, st));
            const int shift = ctx->W - fg2;
            for (int s = 0; s < N; s++) {
                SampleData &sd = ctx->samples[s];
                if (sd.n == 0) continue;
                const unsigned g = (unsigned)std::min<uint64_t>((sd.n + BLOCK - 1) / BLOCK, (uint64_t)ctx->nsm * 16);
                CK(cudaMemsetAsync(ctx->d_fine_hist.p, 0, nb2 * 4, st));
                CK(cudaMemsetAsync(ctx->d_fine_cursor.p, 0, nb2 * 4, st));
                k_refine_hist<<<g, BLOCK, 0, st>>>(sd.eh, sd.n, shift, ctx->d_fine_hist.as<uint32_t>());
                launch_scan(ctx, ctx->d_fine_hist.as<uint32_t>(), ctx->d_fine_off.as<uint64_t>(), nb2);
                k_fill_offsets<<<(unsigned)((row + 255) / 256), 256, 0, st>>>(ctx->d_fine_off.as<uint64_t>(), ctx->d_off_table2.as<uint32_t>() + (size_t)s * row,
                                                                          (uint32_t)nb2, 1);
                if (sd.imported) {
                    // an imported list lives in the caller's memory (read-only for us): the regrouped copy goes to the arena
                    uint64_t *nh = (uint64_t *)ctx->arena.alloc(sd.n * 8);
                    uint32_t *nc = (uint32_t *)ctx->arena.alloc(sd.n * 4);
                    if (!nh || !nc) { ctx->set_error("out of device memory (merge refinement)"); return SIMKA_ERR_NOMEM; }
                    k_refine_scatter<<<g, BLOCK, 0, st>>>(sd.eh, sd.ec, sd.n, shift, ctx->d_fine_off.as<uint64_t>(), ctx->d_fine_cursor.as<uint32_t>(), nh, nc);
                    sd.eh = nh;
                    sd.ec = nc;
                } else {
                    k_refine_scatter<<<g, BLOCK, 0, st>>>(sd.eh, sd.ec, sd.n, shift, ctx->d_fine_off.as<uint64_t>(), ctx->d_fine_cursor.as<uint32_t>(),
                                                          ctx->d_keysA.as<uint64_t>(), ctx->d_stage_c.as<uint32_t>());
                    CK(cudaMemcpyAsync(sd.eh, ctx->d_keysA.p, sd.n * 8, cudaMemcpyDeviceToDevice, st));
                    CK(cudaMemcpyAsync(sd.ec, ctx->d_stage_c.p, sd.n * 4, cudaMemcpyDeviceToDevice, st));
                }
                ctx->launches += 3;
            }
            CK(cudaGetLastError());
            fg_eff = fg2;
            off_eff = ctx->d_off_table2.as<uint32_t>();
        }
    }
    // pointer tables
    std::vector<const void *> ptrs(2 * (size_t)N);
    for (int s = 0; s < N; s++) { ptrs[s] = ctx->wide ? (const void *)ctx->samples[s].ehw : (const void *)ctx->samples[s].eh; ptrs[N + s] = ctx->samples[s].ec; }
    if (ctx->d_ptrs.reserve(ptrs.size() * sizeof(void *))) { ctx->set_error("out of device memory"); return SIMKA_ERR_NOMEM; }
    CK(cudaMemcpyAsync(ctx->d_ptrs.p, ptrs.data(), ptrs.size() * sizeof(void *), cudaMemcpyHostToDevice, st));

    int *d_error = reinterpret_cast<int *>(ctx->d_small.as<char>() + 32);
    uint32_t *d_big_n = reinterpret_cast<uint32_t *>(ctx->d_small.as<char>() + 40);
    uint32_t big_n = 0;
    if (ctx->prm.complex_dist) CK(cudaMemcpyAsync(&big_n, d_big_n, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));  // host staging vectors are reused below / big_n needed
    if (ctx->prm.complex_dist && big_n > ctx->big_cap) {
        // more solid k-mers with an abundance >= CHIST than the list holds (deep or many samples): grow it and rebuild it from
        // the resident lists (the dense part of the histograms is complete)
        const uint32_t want = big_n + big_n / 4 + 1024;
        if (ctx->d_big_counts.reserve((size_t)want * 4) || ctx->d_big_sample.reserve((size_t)want * 4)) { ctx->set_error("out of device memory (abundance list)"); return SIMKA_ERR_NOMEM; }
        ctx->big_cap = want;
        CK(cudaMemsetAsync(d_big_n, 0, 4, st));
        for (int s = 0; s < N; s++) {
            const SampleData &sd = ctx->samples[s];
            if (!sd.n) continue;
            const unsigned g = (unsigned)std::min<uint64_t>((sd.n + BLOCK * 16 - 1) / (BLOCK * 16), (uint64_t)ctx->nsm * 4);
            k_abundance_hist<<<g, BLOCK, 0, st>>>(sd.ec, sd.n, s, nullptr, ctx->d_big_counts.as<uint32_t>(), ctx->d_big_sample.as<uint32_t>(), d_big_n, ctx->big_cap, d_error);
            ctx->launches++;
        }
        CK(cudaMemcpyAsync(&big_n, d_big_n, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    if (ctx->fg_bits >= 0 && total_entries > 0) {
        int mg_bits = std::min(fg_eff, want_bits);
        MergeArgs a;
        a.n = N; a.fg_bits = fg_eff; a.mg_bits = mg_bits;
        a.eh = reinterpret_cast<const uint64_t *const *>(ctx->d_ptrs.p);
        a.ec = reinterpret_cast<const uint32_t *const *>(ctx->d_ptrs.as<const void *>() + N);
        a.off = off_eff;
        a.ntot = ctx->a_ntot;
        a.logtab = ctx->d_logtab.as<double2>();
        a.simple = ctx->prm.simple_dist; a.complex_ = ctx->prm.complex_dist; a.count_only = 0;
        a.nb_shared = ctx->a_nb_shared; a.distinct_shared = ctx->a_distinct; a.braycurtis = ctx->a_bc;
        a.chord = ctx->a_chord; a.hellinger = ctx->a_hell; a.whittaker = ctx->a_whit; a.kl = ctx->a_kl; a.klfix = ctx->a_klfix;
        a.scalars = ctx->a_scalars; a.error_flag = d_error;
        const int nsm = ctx->nsm;
        const uint32_t MG = 1u << mg_bits;
        const unsigned grid = (unsigned)std::min<uint64_t>(MG, (uint64_t)nsm * 8);
        // default: k_merge_sm for N <= 64, k_merge (global atomics) above.  SIMKA_B200_MERGE selects the alternatives the tests
        // keep comparable: "generic" = k_merge for any N, "tiled" = 32 x 32 sample tiles on k_merge_sm (N > 64)
        const char *mv = getenv("SIMKA_B200_MERGE");
        const bool want_generic = mv && !strcmp(mv, "generic"), want_tiled = mv && !strcmp(mv, "tiled");
        const bool want_records = mv ? !strcmp(mv, "records") : true;  // 64 < N <= 512: group once, accumulate per sample tile in shared memory
        if (!ctx->wide && N > MS_MAXN && N <= 32 * TM_MAXB && want_records && !want_generic && !want_tiled) {
            const int rc = merge_by_tiles(ctx, a, MG);
            if (rc != SIMKA_OK) return rc;
        } else if (ctx->wide) {  // k > 31: the generic merge on 16-byte keys
            WideMergeArgs wa;
            wa.m = a;
            wa.ehw = reinterpret_cast<const K128 *const *>(ctx->d_ptrs.p);
            k_wide_merge<<<grid, BLOCK, wide_merge_smem_bytes(N), st>>>(wa);
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
            t.mg_bits = std::min(fg_eff, ilog2_ceil(std::max<uint64_t>(1, total_entries * 64 / (uint64_t)N / 1024)));
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
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
}

int simka_gpu_get_stats(simka_ctx *ctx, simka_stats_view *out) try {
    if (ctx && out && ctx->mctx && ctx->mctx->merged) {  // after simka_gpu_exchange_and_merge: the global statistics live in the merge context
        const int rc = simka_gpu_get_stats(ctx->mctx, out);
        if (rc != SIMKA_OK) ctx->set_error(ctx->mctx->err);
        return rc;
    }
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
                    const simka_sample_totals &ti = ctx->reduced ? ctx->global_totals[i] : ctx->samples[i].local;
                    const simka_sample_totals &tj = ctx->reduced ? ctx->global_totals[j] : ctx->samples[j].local;
                    out->canberra[(size_t)i * N + j] = ti.nb_distinct + tj.nb_distinct - 2 * a;
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
                    const ull only_i = (ctx->reduced ? ctx->global_totals[i].nb_kmers : ctx->samples[i].local.nb_kmers) - nb_shared[ij];
                    const ull only_j = (ctx->reduced ? ctx->global_totals[j].nb_kmers : ctx->samples[j].local.nb_kmers) - nb_shared[ji];
                    if (only_i) v += (Nj > 0) ? ln2 * ((double)only_i / Ni) : onesided_kl_term(1.0, Ni, Nj);
                    if (only_j) v += (Ni > 0) ? ln2 * ((double)only_j / Nj) : onesided_kl_term(1.0, Nj, Ni);
                    out->kullback_leibler[ij] = v;
                }
        }
    }
    return SIMKA_OK;
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
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
    t.superk_samples = ctx->superk_samples;
    t.superk_fallbacks = ctx->superk_fallbacks;
    const simka_ctx *mc = (ctx->mctx && ctx->mctx->merged) ? ctx->mctx : ctx;  // after simka_gpu_exchange_and_merge: stage 2 ran in the merge context
    if (mc != ctx) { t.merge_ms = mc->merge_ms; t.total_ms = ctx->sample_ms + mc->merge_ms; t.kernel_launches += mc->launches; }
    if (mc->merged && mc->a_scalars) {
        ull sc[4];
        if (cudaMemcpy(sc, mc->a_scalars, 32, cudaMemcpyDeviceToHost) == cudaSuccess) { t.distinct_kmers = sc[0]; t.pairs = sc[2]; }
    }
    *out = t;
    return SIMKA_OK;
}

static thread_local simkamin_sketch_info g_sketch_info = {0, 0, 0};  // of this thread's last simkamin_gpu_sketch* call
struct SketchCtxCache {
    std::mutex mu;
    simka_ctx *ctx = nullptr;
    int device = -1, k = 0;
    uint32_t seed = 0;
    bool filter = false;
};
static SketchCtxCache g_sketch_cache;

// The smallest hashes of the (only) sample of a sketch-mode context, ascending, with their counts: its list is grouped by fine
// range = top bits of the hash, so the `want` smallest lie in the leading ranges.  Only the ranges up to the one that holds
// the want-th entry come to the host, and each range (a few dozen entries) is sorted by itself.
static int sketch_fetch_lowest(simka_ctx *ctx, uint64_t want, std::vector<uint64_t> &hk, std::vector<uint32_t> &hc, uint64_t *fetched) {
    const SampleData &sd = ctx->samples[0];
    const uint64_t n = sd.n;
    hk.assign(1, 0);
    hc.assign(1, 0);
    *fetched = 0;
    if (n == 0) return SIMKA_OK;
    CK(cudaSetDevice(ctx->prm.device));
    const size_t fgn = ((size_t)1 << ctx->fg_bits) + 1;
    std::vector<uint32_t> off(fgn);
    CK(cudaMemcpyAsync(off.data(), sd.off, fgn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint64_t m = n;
    if (n > want) m = *std::lower_bound(off.begin(), off.end(), (uint32_t)want);  // end of the range holding the want-th entry (off[fg] = n > want)
    hk.resize(m);
    hc.resize(m);
    CK(cudaMemcpyAsync(hk.data(), sd.eh, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hc.data(), sd.ec, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<std::pair<uint64_t, uint32_t>> seg;
    for (size_t r = 0; r + 1 < fgn && off[r] < m; r++) {
        const uint64_t b = off[r], e = std::min<uint64_t>(off[r + 1], m);
        if (e - b < 2) continue;
        seg.resize(e - b);
        for (uint64_t i = b; i < e; i++) seg[i - b] = {hk[i], hc[i]};
        std::sort(seg.begin(), seg.end());
        for (uint64_t i = b; i < e; i++) { hk[i] = seg[i - b].first; hc[i] = seg[i - b].second; }
    }
    *fetched = m;
    return SIMKA_OK;
}

// filter: the -filter option (processFiltered, src/simkaMin/SimkaMinCount.hpp:341-380): only k-mers seen at least twice are
// sketched.  The reference remembers first occurrences in a Bloom filter of max-memory * 8 / cores bits; this build counts
// exactly (stage 1 with abundance-min 2), i.e. it equals the reference wherever that filter answers no false positive.
static int simkamin_sketch_impl(bool filter, int device, int kmer_size, uint32_t sketch_size, uint32_t seed, const uint64_t *packed, const uint32_t *bad,
                                uint64_t n_words, uint64_t n_reads, uint64_t *hashes_out, uint32_t *counts_out, uint32_t *n_selected) {
    if (!hashes_out || !counts_out || sketch_size == 0 || (n_words && (!packed || !bad))) return SIMKA_ERR_INVALID;
    simka_params prm;
    memset(&prm, 0, sizeof prm);
    prm.kmer_size = kmer_size;
    prm.nb_samples = 1;
    prm.abundance_min = filter ? 2 : 0;
    prm.abundance_max = 0x7FFFFFFFFFFFFFFFll;
    prm.device = device;
    prm.sketch_mode = 1;
    prm.sketch_seed = seed;
    // One stage-1 context is kept between calls (same device / k / seed / filter): a sketch run calls this once per dataset and
    // would otherwise create and free its streams, events and multi-GB work buffers every time.  Calls are serialised on it.
    std::lock_guard<std::mutex> cache_lock(g_sketch_cache.mu);
    SketchCtxCache &cc = g_sketch_cache;
    const bool use_cache = !(getenv("SIMKA_B200_SKETCH_CTX_CACHE") && getenv("SIMKA_B200_SKETCH_CTX_CACHE")[0] == '0');
    simka_ctx *ctx = nullptr;
    int rc = SIMKA_OK;
    if (cc.ctx && use_cache && cc.device == device && cc.k == kmer_size && cc.seed == seed && cc.filter == filter && simka_gpu_reset(cc.ctx) == SIMKA_OK) {
        ctx = cc.ctx;
    } else {
        if (cc.ctx) { simka_gpu_destroy(cc.ctx); cc.ctx = nullptr; }
        rc = simka_gpu_create(&ctx, &prm);
        if (rc != SIMKA_OK) return rc;
    }
    cc.ctx = nullptr;  // owned by this call until it ends well
    auto keep_or_destroy = [&](bool good) {
        if (good && use_cache) { cc.ctx = ctx; cc.device = device; cc.k = kmer_size; cc.seed = seed; cc.filter = filter; }
        else simka_gpu_destroy(ctx);
    };
    std::vector<uint64_t> hk;
    std::vector<uint32_t> hc;
    uint64_t n = 0;
    // Bottom-s pre-filter: only hashes <= key_max are extracted at all.  Every occurrence of such a hash is kept, so the counts
    // are exact, and as soon as >= sketch_size distinct (solid) hashes lie below the bound its s smallest are the s smallest of
    // the sample.  The bound starts at the fraction 16 * s / positions of the hash range (enough while a k-mer occurs <= 16 times
    // on average) and grows 8-fold until the sketch fills or the whole range is taken; small samples take the whole range at once.
    const uint64_t n_pos = n_words * 32;
    double frac = 1.0;
    { const char *e = getenv("SIMKA_B200_SKETCH_PREFILTER");  // =0: count every hash of the sample (the path before the pre-filter)
      if (!(e && e[0] == '0') && n_pos > 64ull * sketch_size) frac = 16.0 * (double)sketch_size / (double)n_pos; }
    g_sketch_info = simkamin_sketch_info{0, 0, 0};
    for (;;) {
        ctx->kp.key_max = frac >= 1.0 ? ~0ull : (uint64_t)(frac * 18446744073709551616.0);
        // exact multiset count of the hashes: stage 1 with key = MurmurHash3(canonical k-mer)
        if ((rc = simka_gpu_push_reads(ctx, 0, packed, bad, n_words, n_reads)) != SIMKA_OK) break;
        if ((rc = simka_gpu_finish_sample(ctx, 0, nullptr)) != SIMKA_OK) break;
        if ((rc = simka_gpu_dump_sample_counts(ctx, 0, nullptr, nullptr, &n)) != SIMKA_OK) break;
        g_sketch_info.attempts++;
        g_sketch_info.keys_counted += ctx->kmers_in;
        g_sketch_info.distinct_dumped = n;
        if (n < sketch_size && ctx->kp.key_max != ~0ull) {  // not enough distinct hashes below the bound: widen it
            frac *= 8.0;
            if ((rc = simka_gpu_reset(ctx)) != SIMKA_OK) break;
            continue;
        }
        uint64_t fetched = 0;
        rc = sketch_fetch_lowest(ctx, sketch_size, hk, hc, &fetched);  // ascending hashes, at least min(n, sketch_size) of them
        g_sketch_info.distinct_dumped = fetched;
        break;
    }
    if (rc != SIMKA_OK) { keep_or_destroy(false); return rc; }
    const uint32_t n_sel = (uint32_t)std::min<uint64_t>(n, sketch_size);
    // the order-dependent count of the largest member once the sketch is complete (only when it can be complete):
    // without -filter a member enters the heap at its first occurrence and the largest one counts occurrences up to the
    // last entry; with -filter members enter at their second occurrence with count 2 and the largest one adds the
    // occurrences between its own entry and the last entry
    if (n >= sketch_size && n_sel > 0) {
        uint64_t *d_sk = nullptr; ull *d_first = nullptr, *d_second = nullptr, *d_cnt = nullptr;
        uint64_t *d_packed = nullptr; uint32_t *d_bad = nullptr;
        bool ok = cudaMalloc(&d_sk, (size_t)n_sel * 8) == cudaSuccess && cudaMalloc(&d_first, (size_t)n_sel * 8) == cudaSuccess &&
                  cudaMalloc(&d_second, (size_t)n_sel * 8) == cudaSuccess &&
                  cudaMalloc(&d_cnt, 8) == cudaSuccess && cudaMalloc(&d_packed, (n_words + 1) * 8) == cudaSuccess &&
                  cudaMalloc(&d_bad, (n_words + 1) * 4) == cudaSuccess;
        ok = ok && cudaMemcpy(d_sk, hk.data(), (size_t)n_sel * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemset(d_first, 0xFF, (size_t)n_sel * 8) == cudaSuccess && cudaMemset(d_second, 0xFF, (size_t)n_sel * 8) == cudaSuccess &&
             cudaMemset(d_cnt, 0, 8) == cudaSuccess &&
             cudaMemcpy(d_packed, packed, (n_words + 1) * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(d_bad, bad, (n_words + 1) * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok) {
            const unsigned g = (unsigned)((n_words + BLOCK - 1) / BLOCK);
            k_sketch_scan<0><<<g, BLOCK>>>(d_packed, d_bad, n_words, kmer_size, seed, d_sk, n_sel, d_first, d_second, 0, 0, d_cnt);
            if (filter) k_sketch_scan<2><<<g, BLOCK>>>(d_packed, d_bad, n_words, kmer_size, seed, d_sk, n_sel, d_first, d_second, 0, 0, d_cnt);
            std::vector<ull> entry(n_sel);  // position at which each member enters the heap
            ok = cudaMemcpy(entry.data(), filter ? d_second : d_first, (size_t)n_sel * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
            ull t_star = 0;
            for (uint32_t i = 0; i < n_sel; i++) t_star = std::max(t_star, entry[i]);
            if (ok) {
                const ull t_from = filter ? entry[n_sel - 1] + 1 : 0;
                k_sketch_scan<1><<<g, BLOCK>>>(d_packed, d_bad, n_words, kmer_size, seed, d_sk, n_sel, d_first, d_second, t_from, t_star, d_cnt);
                ull cnt = 0;
                ok = cudaMemcpy(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost) == cudaSuccess && cudaGetLastError() == cudaSuccess;
                if (ok) hc[n_sel - 1] = (uint32_t)(cnt + (filter ? 2 : 0));
            }
        }
        cudaFree(d_sk); cudaFree(d_first); cudaFree(d_second); cudaFree(d_cnt); cudaFree(d_packed); cudaFree(d_bad);
        if (!ok) { keep_or_destroy(false); return SIMKA_ERR_CUDA; }
    }
    keep_or_destroy(true);
    for (uint32_t i = 0; i < sketch_size; i++) { hashes_out[i] = 0; counts_out[i] = 0; }
    for (uint32_t i = 0; i < n_sel; i++) { hashes_out[sketch_size - n_sel + i] = hk[i]; counts_out[sketch_size - n_sel + i] = hc[i]; }
    if (n_selected) *n_selected = n_sel;
    return SIMKA_OK;
}

int simkamin_gpu_sketch(int device, int kmer_size, uint32_t sketch_size, uint32_t seed, const uint64_t *packed, const uint32_t *bad,
                        uint64_t n_words, uint64_t n_reads, uint64_t *hashes_out, uint32_t *counts_out, uint32_t *n_selected) {
    return simkamin_sketch_impl(false, device, kmer_size, sketch_size, seed, packed, bad, n_words, n_reads, hashes_out, counts_out, n_selected);
}

int simkamin_gpu_sketch_filter(int device, int kmer_size, uint32_t sketch_size, uint32_t seed, const uint64_t *packed, const uint32_t *bad,
                               uint64_t n_words, uint64_t n_reads, uint64_t *hashes_out, uint32_t *counts_out, uint32_t *n_selected) {
    return simkamin_sketch_impl(true, device, kmer_size, sketch_size, seed, packed, bad, n_words, n_reads, hashes_out, counts_out, n_selected);
}

void simkamin_gpu_release_cache(void) {
    std::lock_guard<std::mutex> lk(g_sketch_cache.mu);
    if (g_sketch_cache.ctx) { simka_gpu_destroy(g_sketch_cache.ctx); g_sketch_cache.ctx = nullptr; }
}

int simkamin_gpu_last_sketch_info(simkamin_sketch_info *out) {
    if (!out) return SIMKA_ERR_INVALID;
    *out = g_sketch_info;
    return SIMKA_OK;
}

}  // extern "C"
namespace {
// Device buffers of the SimkaMin distance entry points, kept between calls (`simkaMinCore distance` walks the matrix in rectangular
// blocks, one call each; a cudaMalloc / cudaFree pair per call also serialises against anything else that holds the driver lock).
// One call at a time; simkamin_gpu_release() frees them.
struct SketchBuffers {
    std::mutex mu;
    int device = -1;
    DevBuf b[7];
    void release() {
        if (device >= 0 && cudaSetDevice(device) == cudaSuccess) for (DevBuf &x : b) x.release();
        device = -1;
    }
    int use(int dev) {  // 0: the cache now belongs to device dev (which is current)
        if (device != dev) release();
        if (cudaSetDevice(dev) != cudaSuccess) return -1;
        device = dev;
        return 0;
    }
};
SketchBuffers g_sketch_buffers;
}  // namespace
extern "C" {

int simkamin_gpu_release(void) {
    std::lock_guard<std::mutex> lock(g_sketch_buffers.mu);
    g_sketch_buffers.release();
    return SIMKA_OK;
}

int simkamin_gpu_distance(int device, const uint64_t *hashes, const uint32_t *counts, uint32_t n, uint32_t s, float *jaccard,
                          float *braycurtis, float *kernel_ms) {
    if (!hashes || !counts || !jaccard || !braycurtis || n == 0 || s == 0) return SIMKA_ERR_INVALID;
    std::lock_guard<std::mutex> lock(g_sketch_buffers.mu);
    if (g_sketch_buffers.use(device)) return SIMKA_ERR_CUDA;
    DevBuf *B = g_sketch_buffers.b;
    uint64_t *dh = nullptr; uint32_t *dc = nullptr, *dstart = nullptr; ull *dq = nullptr;
    const size_t ne = (size_t)n * s, nn = (size_t)n * n;
    const uint64_t pairs = (uint64_t)n * (n - 1) / 2;
    if (pairs > 0x7FFFFFFFull) return SIMKA_ERR_INVALID;  // one CTA per pair: the grid's x limit (n <= 65536)
    int rc = SIMKA_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<ull> hq;
    try { hq.resize(4 * pairs); } catch (const std::bad_alloc &) { return SIMKA_ERR_NOMEM; }  // nothing may throw across the C ABI
    do {
        // (+64: the bulk copies of the distance kernel round their ranges outwards to 16 bytes)
        if (B[0].reserve(ne * 8 + 64) || B[1].reserve(ne * 4 + 64) || B[2].reserve((size_t)n * 4) || B[6].reserve((pairs + 1) * 32)) { cudaGetLastError(); rc = SIMKA_ERR_NOMEM; break; }
        dh = B[0].as<uint64_t>(); dc = B[1].as<uint32_t>(); dstart = B[2].as<uint32_t>(); dq = B[6].as<ull>();
        if (cudaFuncSetAttribute(k_simkamin_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmtSmem)) != cudaSuccess) { rc = SIMKA_ERR_CUDA; break; }
        if (cudaMemcpy(dh, hashes, ne * 8, cudaMemcpyHostToDevice) || cudaMemcpy(dc, counts, ne * 4, cudaMemcpyHostToDevice)) { rc = SIMKA_ERR_CUDA; break; }
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_sketch_starts<<<(n + 255) / 256, 256>>>(dh, n, s, dstart);
        if (pairs) k_simkamin_tiled<false><<<(unsigned)pairs, BLOCK, sizeof(SmtSmem)>>>(dh, dc, dstart, n, nullptr, nullptr, nullptr, 0, s, dq);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = SIMKA_ERR_CUDA; break; }
        if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
        if (pairs && cudaMemcpy(hq.data(), dq, pairs * 32, cudaMemcpyDeviceToHost)) { rc = SIMKA_ERR_CUDA; break; }
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
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

// all n1 x n2 pairs between two sets of sketches (the rectangular blocks of `simkaMinCore distance`,
// src/simkaMin/SimkaMinDistance.hpp:642-753); outputs n1*n2 float32 row-major
int simkamin_gpu_distance_rect(int device, const uint64_t *hashes1, const uint32_t *counts1, uint32_t n1, const uint64_t *hashes2,
                               const uint32_t *counts2, uint32_t n2, uint32_t s, float *jaccard, float *braycurtis, float *kernel_ms) {
    if (!hashes1 || !counts1 || !hashes2 || !counts2 || !jaccard || !braycurtis || n1 == 0 || n2 == 0 || s == 0) return SIMKA_ERR_INVALID;
    std::lock_guard<std::mutex> lock(g_sketch_buffers.mu);
    if (g_sketch_buffers.use(device)) return SIMKA_ERR_CUDA;
    DevBuf *B = g_sketch_buffers.b;
    uint64_t *dh1 = nullptr, *dh2 = nullptr; uint32_t *dc1 = nullptr, *dc2 = nullptr, *ds1 = nullptr, *ds2 = nullptr; ull *dq = nullptr;
    const size_t e1 = (size_t)n1 * s, e2 = (size_t)n2 * s;
    const uint64_t pairs = (uint64_t)n1 * n2;
    if (pairs > 0x7FFFFFFFull) return SIMKA_ERR_INVALID;  // one CTA per pair: the grid's x limit
    int rc = SIMKA_OK;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<ull> hq;
    try { hq.resize(4 * pairs); } catch (const std::bad_alloc &) { return SIMKA_ERR_NOMEM; }
    do {
        if (cudaFuncSetAttribute(k_simkamin_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmtSmem)) != cudaSuccess) { rc = SIMKA_ERR_CUDA; break; }
        if (B[0].reserve(e1 * 8 + 64) || B[1].reserve(e1 * 4 + 64) || B[2].reserve((size_t)n1 * 4) || B[3].reserve(e2 * 8 + 64) || B[4].reserve(e2 * 4 + 64) ||
            B[5].reserve((size_t)n2 * 4) || B[6].reserve((pairs + 1) * 32)) { cudaGetLastError(); rc = SIMKA_ERR_NOMEM; break; }
        dh1 = B[0].as<uint64_t>(); dc1 = B[1].as<uint32_t>(); ds1 = B[2].as<uint32_t>(); dh2 = B[3].as<uint64_t>(); dc2 = B[4].as<uint32_t>(); ds2 = B[5].as<uint32_t>();
        dq = B[6].as<ull>();
        if (cudaMemcpy(dh1, hashes1, e1 * 8, cudaMemcpyHostToDevice) || cudaMemcpy(dc1, counts1, e1 * 4, cudaMemcpyHostToDevice) ||
            cudaMemcpy(dh2, hashes2, e2 * 8, cudaMemcpyHostToDevice) || cudaMemcpy(dc2, counts2, e2 * 4, cudaMemcpyHostToDevice)) { rc = SIMKA_ERR_CUDA; break; }
        cudaEventCreate(&ev0); cudaEventCreate(&ev1);
        cudaEventRecord(ev0);
        k_sketch_starts<<<(n1 + 255) / 256, 256>>>(dh1, n1, s, ds1);
        k_sketch_starts<<<(n2 + 255) / 256, 256>>>(dh2, n2, s, ds2);
        k_simkamin_tiled<true><<<(unsigned)pairs, BLOCK, sizeof(SmtSmem)>>>(dh1, dc1, ds1, n1, dh2, dc2, ds2, n2, s, dq);
        cudaEventRecord(ev1);
        if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = SIMKA_ERR_CUDA; break; }
        if (kernel_ms) cudaEventElapsedTime(kernel_ms, ev0, ev1);
        if (cudaMemcpy(hq.data(), dq, pairs * 32, cudaMemcpyDeviceToHost)) { rc = SIMKA_ERR_CUDA; break; }
    } while (0);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (rc) return rc;
    for (uint64_t p = 0; p < pairs; p++) {  // final ratios exactly as the reference: long double division, stored as float (:244-259)
        const ull sh = hq[4 * p], di = hq[4 * p + 1], sk = hq[4 * p + 2], km = hq[4 * p + 3];
        jaccard[p] = di == 0 ? 1.0f : (float)(1 - (long double)sh / (long double)di);
        braycurtis[p] = km == 0 ? 1.0f : (float)(1 - (long double)(2 * sk) / (long double)km);
    }
    return rc;
}

// =============================================================================================
// multi-GPU: NCCL (loaded at run time) -- totals all-reduce, list exchange, merge, accumulator all-reduce
// =============================================================================================
}  // extern "C" (helpers below have C++ linkage)

namespace {

#if SIMKA_HAVE_NCCL_H
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi &nccl() {
    static NcclApi a = [] {
        NcclApi n;
        // an already loaded libnccl.so.2 (e.g. PyTorch's) is found first; else the system one
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) { n.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (n.h) break; }
        if (!n.h) return n;
        auto sym = [&](const char *s) { return dlsym(n.h, s); };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
        n.Send = (decltype(n.Send))sym("ncclSend");
        n.Recv = (decltype(n.Recv))sym("ncclRecv");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.Send && n.Recv && n.GroupStart && n.GroupEnd && n.GetErrorString;
        return n;
    }();
    return a;
}
#define NK(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            ctx->set_error(std::string(#call) + ": " + nccl().GetErrorString(r_));                  \
            return SIMKA_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)
#endif

// one sample: row[r] = entries that fall into the block of fine ranges rank r merges; row[geo..geo+2] = geometry check (fb, 1, fb^2)
__global__ void k_xchg_counts1(const uint32_t *__restrict__ off, int world, uint32_t per, int parts, ull *__restrict__ row, ull *__restrict__ geo, int fb) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < world) row[r] = r < parts ? (ull)(off[(size_t)(r + 1) * per] - off[(size_t)r * per]) : 0ull;
    if (r == 0) { geo[0] = (ull)fb; geo[1] = 1ull; geo[2] = (ull)fb * (ull)fb; }
}
// received offset rows are slices of the owner's table: re-base every row to its first entry
__global__ void k_xchg_rebase(uint32_t *__restrict__ rows, size_t fgn, int n_rows) {
    const size_t row = blockIdx.y;
    if ((int)row >= n_rows) return;
    uint32_t *p = rows + row * fgn;
    const uint32_t base = p[0];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < fgn; i += (size_t)gridDim.x * blockDim.x)
        if (i) p[i] -= base;
    // p[0] is reset last by the kernel below (another block may still read it as the base)
}
__global__ void k_xchg_zero_first(uint32_t *__restrict__ rows, size_t fgn, int n_rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < n_rows) rows[(size_t)row * fgn] = 0;
}

}  // namespace

static void comm_release(simka_ctx *ctx) {
#if SIMKA_HAVE_NCCL_H
    if (ctx->comm && nccl().ok) nccl().CommDestroy((ncclComm_t)ctx->comm);
#endif
    ctx->comm = nullptr;
}

extern "C" {

int simka_comm_unique_id(uint8_t id[SIMKA_COMM_ID_BYTES]) {
#if SIMKA_HAVE_NCCL_H
    static_assert(sizeof(ncclUniqueId) <= SIMKA_COMM_ID_BYTES, "ncclUniqueId grew");
    if (!id || !nccl().ok) return SIMKA_ERR_INVALID;
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != ncclSuccess) return SIMKA_ERR_CUDA;
    memset(id, 0, SIMKA_COMM_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return SIMKA_OK;
#else
    (void)id;
    return SIMKA_ERR_INVALID;
#endif
}

int simka_gpu_comm_init(simka_ctx *ctx, const uint8_t id[SIMKA_COMM_ID_BYTES], int rank, int world) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return SIMKA_ERR_INVALID;
#if SIMKA_HAVE_NCCL_H
    if (!nccl().ok) { ctx->set_error("libnccl.so.2 could not be loaded"); return SIMKA_ERR_INVALID; }
    CK(cudaSetDevice(ctx->prm.device));
    comm_release(ctx);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    NK(nccl().CommInitRank(&c, world, u, rank));
    ctx->comm = c;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    for (auto &e : ctx->xev) if (!e) CK(cudaEventCreate(&e));
    return SIMKA_OK;
#else
    ctx->set_error("built without nccl.h");
    return SIMKA_ERR_INVALID;
#endif
}

int simka_gpu_exchange_timings(simka_ctx *ctx, float ms[4]) {
    if (!ctx || !ms) return SIMKA_ERR_INVALID;
    for (int i = 0; i < 4; i++) ms[i] = ctx->xms[i];
    return SIMKA_OK;
}

}  // extern "C"

#if SIMKA_HAVE_NCCL_H
// ---- the exchange, round by round -------------------------------------------------------------------------------------------
// Round t = the samples t * world + r (one per rank).  After its sample is counted a rank enqueues, on the side stream, the
// all-reduce of the round's slice sizes [owner][destination]; one sample later (the sizes have long arrived) it posts the sends
// and receives of the round, which then overlap the counting of the following samples.  Every rank issues the NCCL operations
// in the same order: sizes(0), data(0), sizes(1), data(1), ...  Only FULL rounds start early (every rank has a sample in them);
// a trailing partial round is exchanged by simka_gpu_exchange_and_merge, which all ranks call.
static int xchg_setup(simka_ctx *ctx) {
    const int N = ctx->prm.nb_samples, world = ctx->comm_world;
    const int T = (N + world - 1) / world;
    CK(cudaSetDevice(ctx->prm.device));
    if (!ctx->xstream) CK(cudaStreamCreateWithFlags(&ctx->xstream, cudaStreamNonBlocking));
    if (!ctx->x_sample_done) CK(cudaEventCreateWithFlags(&ctx->x_sample_done, cudaEventDisableTiming));
    if (!ctx->x_all_done) CK(cudaEventCreateWithFlags(&ctx->x_all_done, cudaEventDisableTiming));
    const size_t row = (size_t)world * world + 3;
    if ((int)ctx->xr.size() != T) {
        for (auto &r : ctx->xr) if (r.sizes_ready) cudaEventDestroy(r.sizes_ready);
        ctx->xr.assign(T, simka_ctx::XRound());
        for (auto &r : ctx->xr) CK(cudaEventCreateWithFlags(&r.sizes_ready, cudaEventDisableTiming));
        if (ctx->xh) { cudaFreeHost(ctx->xh); ctx->xh = nullptr; }
        CK(cudaHostAlloc((void **)&ctx->xh, (size_t)T * row * 8, cudaHostAllocDefault));
        if (ctx->d_xcnt.reserve((size_t)T * row * 8) || ctx->d_xids.reserve(64)) { ctx->set_error("out of device memory"); return SIMKA_ERR_NOMEM; }
    }
    if (!ctx->mctx) {
        simka_params mp = ctx->prm;
        mp.shard_rank = 0; mp.shard_count = 1; mp.expected_kmers_per_sample = 0;
        const int rc = simka_gpu_create(&ctx->mctx, &mp);
        if (rc != SIMKA_OK) { ctx->set_error("could not create the merge context"); return rc; }
    }
    ctx->x_recv_e.assign(N, nullptr);
    ctx->x_recv_c.assign(N, nullptr);
    ctx->x_recv_n.assign(N, 0);
    return SIMKA_OK;
}

// geometry of the slices from this rank's fine-range bits (every rank derives the same one; checked with every round)
static void xchg_geometry(simka_ctx *ctx, int fb) {
    int parts = 1;
    while (parts * 2 <= ctx->comm_world && parts * 2 <= (1 << fb)) parts *= 2;
    int lb = fb;
    for (int p = parts; p > 1; p >>= 1) lb--;
    ctx->x_fb = fb; ctx->x_parts = parts; ctx->x_lb = lb;
}

// side stream: sizes of round t -> all ranks (this rank's sample of the round is `s`, or -1 when it has none)
static int xchg_enqueue_sizes(simka_ctx *ctx, int t, int s) {
    const int world = ctx->comm_world, rank = ctx->comm_rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const size_t row = (size_t)world * world + 3;
    ull *d_row = ctx->d_xcnt.as<ull>() + (size_t)t * row;
    CK(cudaMemsetAsync(d_row, 0, row * 8, ctx->xstream));
    if (s >= 0) {
        CK(cudaEventRecord(ctx->x_sample_done, ctx->stream));  // the sample's list and offsets are final behind this point
        CK(cudaStreamWaitEvent(ctx->xstream, ctx->x_sample_done, 0));
        const int fb = ctx->fg_bits;
        if (ctx->x_fb < 0) xchg_geometry(ctx, fb);
        const uint32_t per = 1u << ctx->x_lb;
        // entries of my sample per destination = row `rank` of the [owner][destination] table
        k_xchg_counts1<<<(world + 255) / 256, 256, 0, ctx->xstream>>>(ctx->d_off_table.as<uint32_t>() + (size_t)s * (((size_t)1 << fb) + 1), world, per, ctx->x_parts,
                                                                    d_row + (size_t)rank * world, d_row + (size_t)world * world, fb);
        ctx->launches++;
    }
    NK(nccl().AllReduce(d_row, d_row, row, ncclUint64, ncclSum, comm, ctx->xstream));
    CK(cudaMemcpyAsync(ctx->xh + (size_t)t * row, d_row, row * 8, cudaMemcpyDeviceToHost, ctx->xstream));
    CK(cudaEventRecord(ctx->xr[t].sizes_ready, ctx->xstream));
    ctx->xr[t].state = 1;
    return SIMKA_OK;
}

// side stream: sends and receives of round t (its sizes must have been enqueued); waits for the sizes on the host
static int xchg_post_data(simka_ctx *ctx, int t) {
    const int N = ctx->prm.nb_samples, world = ctx->comm_world, rank = ctx->comm_rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    simka_ctx *m = ctx->mctx;
    const size_t row = (size_t)world * world + 3;
    CK(cudaEventSynchronize(ctx->xr[t].sizes_ready));
    const ull *h = ctx->xh + (size_t)t * row;  // [owner * world + destination]
    const ull *geo = h + (size_t)world * world;
    if (geo[1] == 0) { ctx->xr[t].state = 2; return SIMKA_OK; }  // nobody had a sample in this round
    if (geo[1] * geo[2] != geo[0] * geo[0]) {
        ctx->set_error("ranks disagree on the fine-range geometry: pass the same expected_kmers_per_sample to every rank");
        return SIMKA_ERR_INVALID;
    }
    const int fb = (int)(geo[0] / geo[1]);
    if (ctx->x_fb < 0) xchg_geometry(ctx, fb);
    if (ctx->x_fb != fb) { ctx->set_error("fine-range geometry changed between rounds"); return SIMKA_ERR_INVALID; }
    const int parts = ctx->x_parts, lb = ctx->x_lb;
    const uint32_t per = 1u << lb;
    const size_t fgn = ((size_t)1 << lb) + 1, fgn_src = ((size_t)1 << fb) + 1;
    if (!ctx->x_recv_o) {
        ctx->x_recv_o = (uint32_t *)m->arena.alloc((size_t)N * fgn * 4);
        if (!ctx->x_recv_o) { ctx->set_error("out of device memory (exchange buffers)"); return SIMKA_ERR_NOMEM; }
        CK(cudaMemsetAsync(ctx->x_recv_o, 0, (size_t)N * fgn * 4, ctx->xstream));
    }
    // receive buffers of the round's samples (this rank's block of hash ranges)
    for (int o = 0; o < world; o++) {
        const int s = t * world + o;
        if (s >= N) break;
        const uint64_t n = rank < parts ? h[(size_t)o * world + rank] : 0;
        ctx->x_recv_n[s] = n;
        ctx->x_recv_e[s] = (uint64_t *)m->arena.alloc(std::max<uint64_t>(n, 1) * 8);
        ctx->x_recv_c[s] = (uint32_t *)m->arena.alloc(std::max<uint64_t>(n, 1) * 4);
        if (!ctx->x_recv_e[s] || !ctx->x_recv_c[s]) { ctx->set_error("out of device memory (exchange buffers)"); return SIMKA_ERR_NOMEM; }
    }
    const int mine = t * world + rank < N ? t * world + rank : -1;
    NK(nccl().GroupStart());
    if (mine >= 0) {
        const SampleData &sd = ctx->samples[mine];
        uint64_t a = 0;
        for (int r = 0; r < parts; r++) {
            const uint64_t n = h[(size_t)rank * world + r];
            if (r != rank) {
                if (n) {
                    NK(nccl().Send(sd.eh + a, n, ncclUint64, r, comm, ctx->xstream));
                    NK(nccl().Send(sd.ec + a, n, ncclUint32, r, comm, ctx->xstream));
                }
                NK(nccl().Send(ctx->d_off_table.as<uint32_t>() + (size_t)mine * fgn_src + (size_t)r * per, fgn, ncclUint32, r, comm, ctx->xstream));
            }
            a += n;
        }
    }
    if (rank < parts) {
        for (int o = 0; o < world; o++) {
            const int s = t * world + o;
            if (s >= N) break;
            if (o == rank) continue;
            if (ctx->x_recv_n[s]) {
                NK(nccl().Recv(ctx->x_recv_e[s], ctx->x_recv_n[s], ncclUint64, o, comm, ctx->xstream));
                NK(nccl().Recv(ctx->x_recv_c[s], ctx->x_recv_n[s], ncclUint32, o, comm, ctx->xstream));
            }
            NK(nccl().Recv(ctx->x_recv_o + (size_t)s * fgn, fgn, ncclUint32, o, comm, ctx->xstream));
        }
    }
    NK(nccl().GroupEnd());
    if (mine >= 0 && rank < parts) {  // my own slice: a device copy
        const SampleData &sd = ctx->samples[mine];
        uint64_t a = 0;
        for (int q = 0; q < rank; q++) a += h[(size_t)rank * world + q];
        const uint64_t n = h[(size_t)rank * world + rank];
        if (n) {
            CK(cudaMemcpyAsync(ctx->x_recv_e[mine], sd.eh + a, n * 8, cudaMemcpyDeviceToDevice, ctx->xstream));
            CK(cudaMemcpyAsync(ctx->x_recv_c[mine], sd.ec + a, n * 4, cudaMemcpyDeviceToDevice, ctx->xstream));
        }
        CK(cudaMemcpyAsync(ctx->x_recv_o + (size_t)mine * fgn, ctx->d_off_table.as<uint32_t>() + (size_t)mine * fgn_src + (size_t)rank * per, fgn * 4,
                           cudaMemcpyDeviceToDevice, ctx->xstream));
    }
    ctx->xr[t].state = 2;
    return SIMKA_OK;
}

static int xchg_after_sample(simka_ctx *ctx, int sample_id) {
    const int N = ctx->prm.nb_samples, world = ctx->comm_world, rank = ctx->comm_rank;
    if (sample_id % world != rank || ctx->kp.shard_count > 1 || ctx->kp.key_mode || ctx->wide) return SIMKA_OK;
    const int t = sample_id / world;
    if (t >= N / world) return SIMKA_OK;  // a partial last round: exchanged by simka_gpu_exchange_and_merge (every rank is there)
    if (ctx->xr.empty() || ctx->x_recv_e.empty()) { const int rc = xchg_setup(ctx); if (rc != SIMKA_OK) return rc; }
    for (int q = 0; q < t; q++) if (ctx->xr[q].state == 0) return SIMKA_OK;  // samples finished out of order: leave everything to the final call
    if (ctx->xr[t].state != 0) return SIMKA_OK;
    if (t > 0 && ctx->xr[t - 1].state == 1) { const int rc = xchg_post_data(ctx, t - 1); if (rc != SIMKA_OK) return rc; }
    return xchg_enqueue_sizes(ctx, t, sample_id);
}
#else
static int xchg_after_sample(simka_ctx *, int) { return SIMKA_OK; }
#endif

extern "C" {

int simka_gpu_exchange_and_merge(simka_ctx *ctx) try {
    if (!ctx) return SIMKA_ERR_INVALID;
#if SIMKA_HAVE_NCCL_H
    if (!ctx->comm) { ctx->set_error("simka_gpu_comm_init has not run"); return SIMKA_ERR_INVALID; }
    if (ctx->kp.shard_count > 1 || ctx->kp.key_mode || ctx->wide) { ctx->set_error("exchange needs an unsharded counting context with k <= 31"); return SIMKA_ERR_INVALID; }
    const int N = ctx->prm.nb_samples, world = ctx->comm_world, rank = ctx->comm_rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    CK(cudaSetDevice(ctx->prm.device));
    for (int s = rank; s < N; s += world) if (!ctx->samples[s].done) { ctx->set_error("a sample of this rank is not finished"); return SIMKA_ERR_INVALID; }
    if (ctx->xr.empty() || ctx->x_recv_e.empty()) { const int rc = xchg_setup(ctx); if (rc != SIMKA_OK) return rc; }
    simka_ctx *m = ctx->mctx;
    const int T = (int)ctx->xr.size();
    CK(cudaEventRecord(ctx->xev[0], ctx->xstream));
    // ---- 1. the rounds that have not run yet (all of them when the samples were not finished in order; else the last ones)
    for (int t = 0; t < T; t++) {
        if (ctx->xr[t].state == 0) {
            const int s = t * world + rank < N ? t * world + rank : -1;
            const int rc = xchg_enqueue_sizes(ctx, t, s);
            if (rc != SIMKA_OK) return rc;
        }
        if (ctx->xr[t].state == 1) { const int rc = xchg_post_data(ctx, t); if (rc != SIMKA_OK) return rc; }
    }
    const size_t fgn = ((size_t)1 << ctx->x_lb) + 1;
    if (rank < ctx->x_parts) {
        k_xchg_rebase<<<dim3((unsigned)std::min<size_t>((fgn + 255) / 256, 64), (unsigned)N), 256, 0, ctx->xstream>>>(ctx->x_recv_o, fgn, N);
        k_xchg_zero_first<<<(N + 255) / 256, 256, 0, ctx->xstream>>>(ctx->x_recv_o, fgn, N);
        ctx->launches += 2;
    }
    CK(cudaEventRecord(ctx->xev[1], ctx->xstream));
    // ---- 2. global per-sample totals
    const size_t n_hdr = (size_t)N * 4;
    std::vector<ull> hdr(n_hdr, 0);
    for (int s = rank; s < N; s += world) {
        const simka_sample_totals &t = ctx->samples[s].local;
        hdr[(size_t)s * 4] = t.nb_reads; hdr[(size_t)s * 4 + 1] = t.nb_distinct; hdr[(size_t)s * 4 + 2] = t.nb_kmers; hdr[(size_t)s * 4 + 3] = t.chord_n2;
    }
    if (ctx->d_xhdr.reserve(n_hdr * 8)) { ctx->set_error("out of device memory"); return SIMKA_ERR_NOMEM; }
    ull *d_hdr = ctx->d_xhdr.as<ull>();
    CK(cudaMemcpyAsync(d_hdr, hdr.data(), n_hdr * 8, cudaMemcpyHostToDevice, ctx->xstream));
    NK(nccl().AllReduce(d_hdr, d_hdr, n_hdr, ncclUint64, ncclSum, comm, ctx->xstream));
    CK(cudaMemcpyAsync(hdr.data(), d_hdr, n_hdr * 8, cudaMemcpyDeviceToHost, ctx->xstream));
    CK(cudaEventRecord(ctx->xev[2], ctx->xstream));
    CK(cudaEventRecord(ctx->x_all_done, ctx->xstream));
    CK(cudaStreamSynchronize(ctx->xstream));
    std::vector<simka_sample_totals> g_tot(N);
    for (int s = 0; s < N; s++) g_tot[s] = simka_sample_totals{hdr[(size_t)s * 4], hdr[(size_t)s * 4 + 1], hdr[(size_t)s * 4 + 2], hdr[(size_t)s * 4 + 3]};
    // ---- 3. stage 2 on my block of hash ranges over all samples
    CK(cudaStreamWaitEvent(m->stream, ctx->x_all_done, 0));
    for (int s = 0; s < N; s++) {
        const int rc = simka_gpu_import_sample(m, s, ctx->x_recv_e[s], ctx->x_recv_c[s], ctx->x_recv_n[s], ctx->x_recv_o + (size_t)s * fgn, ctx->x_lb, g_tot[s].nb_reads);
        if (rc != SIMKA_OK) { ctx->set_error("import: " + m->err); return rc; }
    }
    simka_gpu_set_totals(m, g_tot.data());
    {
        const int rc = simka_gpu_merge_and_accumulate(m);
        if (rc != SIMKA_OK) { ctx->set_error("merge: " + m->err); return rc; }
    }
    // ---- 4. sum of the accumulators over the ranks: u64 block | f64 block | fixed-point block
    {
        const size_t nn = (size_t)N * N, ns = (size_t)N * (N + 1) / 2;
        const size_t n_u64 = nn + ns + ns + nn + nn + nn + nn + 4;  // up to and including the scalars
        ull *acc = m->d_acc.as<ull>();
        CK(cudaEventRecord(ctx->xev[3], m->stream));
        NK(nccl().GroupStart());
        NK(nccl().AllReduce(acc, acc, n_u64, ncclUint64, ncclSum, comm, m->stream));
        NK(nccl().AllReduce(m->a_kl, m->a_kl, nn, ncclFloat64, ncclSum, comm, m->stream));
        NK(nccl().AllReduce(m->a_klfix, m->a_klfix, nn, ncclUint64, ncclSum, comm, m->stream));
        NK(nccl().GroupEnd());
        CK(cudaEventRecord(ctx->xev[4], m->stream));
        CK(cudaStreamSynchronize(m->stream));
    }
    m->reduced = true;
    ctx->global_totals = g_tot;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->xev[1], ctx->xev[2]) == cudaSuccess) ctx->xms[0] = ms;  // totals all-reduce
    if (cudaEventElapsedTime(&ms, ctx->xev[0], ctx->xev[1]) == cudaSuccess) ctx->xms[1] = ms;  // what was left of the list exchange after the last sample
    ctx->xms[2] = m->merge_ms;
    if (cudaEventElapsedTime(&ms, ctx->xev[3], ctx->xev[4]) == cudaSuccess) ctx->xms[3] = ms;
    return SIMKA_OK;
#else
    ctx->set_error("built without nccl.h");
    return SIMKA_ERR_INVALID;
#endif
} catch (const std::bad_alloc &) {  // nothing may throw across the C ABI
    if (ctx) ctx->set_error("out of host memory");
    return SIMKA_ERR_NOMEM;
}

}  // extern "C"
