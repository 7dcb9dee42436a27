// cuda_runtime.h (tests/emu) -- a functional stand-in for the CUDA runtime and device intrinsics that lets the SAME kernel
// sources (simka_b200/csrc/simka_gpu.cu + kernels/*.cuh) be compiled with g++ and stepped on the CPU.
//
// TEST INFRASTRUCTURE ONLY.  It exists so that kernel LOGIC (index arithmetic, bit manipulation, table protocols, bin
// geometry, edge cases) is exercised by the CPU test suite and can be debugged without a GPU box.  It is built by
// tests/emu/build_emu.py into tests/emu/_build/libsimka_b200_emu.so, loaded only by tests/test_emu_kernels.py, never by
// the package, the command-line tools, bench.py or __graft_entry__; it says nothing about performance, memory ordering or
// races (one OS thread, CUDA threads are cooperative fibers, blocks run one after the other in blockIdx order).
//
// Model: a launch runs its blocks sequentially; the threads of a block are ucontext fibers scheduled round-robin.
// __syncthreads() and the warp collectives (__shfl_*_sync, __ballot_sync) park a fiber until every live thread of the block /
// warp has arrived; a scheduler pass without progress aborts with "deadlock".  `__shared__` variables become function-level
// statics (one block at a time), dynamic shared memory is one global buffer.  The launch syntax `k<<<g, b, s, st>>>(args)` is
// rewritten by build_emu.py into emu::launch(g, b, s, [&]{ k(args); }).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <mutex>
#include <type_traits>
#include <vector>

// ------------------------------------------------------------------------------------------------------------------
// language keywords
// ------------------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct alignas(16) double2 { double x, y; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline double2 make_double2(double x, double y) { double2 d; d.x = x; d.y = y; return d; }

// ------------------------------------------------------------------------------------------------------------------
// the fiber scheduler
// ------------------------------------------------------------------------------------------------------------------
namespace emu {

constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 256 << 10;
constexpr size_t DYN_SMEM_BYTES = 232 << 10;

enum State { RUNNABLE = 0, AT_BLOCK_BARRIER, AT_WARP_BARRIER, DONE };

struct Fiber {
    ucontext_t ctx;
    State state;
    unsigned warp_gen_seen;
};

struct Block {
    dim3 grid, block, bid;
    unsigned nthreads = 0;
    Fiber fib[MAX_THREADS];
    unsigned cur = 0;
    unsigned warp_gen[MAX_THREADS / 32];
    alignas(16) unsigned char warp_buf[MAX_THREADS / 32][32][16];
    unsigned char warp_pred[MAX_THREADS / 32][32];
    ucontext_t sched;
    void (*body)(void *) = nullptr;
    void *body_arg = nullptr;
};

inline Block &B() { static Block *b = new Block(); return *b; }
inline unsigned char *stacks() { static unsigned char *s = (unsigned char *)malloc(STACK_BYTES * MAX_THREADS); return s; }
inline unsigned char *dyn_smem() { alignas(16) static unsigned char s[DYN_SMEM_BYTES]; return s; }
inline uint64_t &launch_count() { static uint64_t n = 0; return n; }
inline std::mutex &launch_mutex() { static std::mutex m; return m; }  // host threads driving distinct contexts take turns

inline void yield_to_scheduler() { Block &b = B(); swapcontext(&b.fib[b.cur].ctx, &b.sched); }

inline void fiber_main() {
    Block &b = B();
    b.body(b.body_arg);
    b.fib[b.cur].state = DONE;
    swapcontext(&b.fib[b.cur].ctx, &b.sched);
}

inline void run_block() {
    Block &b = B();
    const unsigned n = b.nthreads, nw = (n + 31) / 32;
    for (unsigned w = 0; w < nw; w++) b.warp_gen[w] = 0;
    for (unsigned t = 0; t < n; t++) {
        Fiber &f = b.fib[t];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks() + (size_t)t * STACK_BYTES;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())fiber_main, 0);
        f.state = RUNNABLE;
        f.warp_gen_seen = 0;
    }
    // order in which the runnable fibers of a pass are resumed: ascending thread index by default; SIMKA_EMU_SCHED=reverse or
    // =random[:seed] changes it, so that code which only works because lower-numbered threads happen to run first (a missing
    // __syncthreads, warp-synchronous assumptions) gives different -- wrong -- results
    static const int sched_mode = [] { const char *e = getenv("SIMKA_EMU_SCHED"); return !e ? 0 : (!strncmp(e, "reverse", 7) ? 1 : (!strncmp(e, "random", 6) ? 2 : 0)); }();
    static uint64_t rng = [] { const char *e = getenv("SIMKA_EMU_SCHED"); const char *c = e ? strchr(e, ':') : nullptr; return (uint64_t)(c ? atoll(c + 1) : 1) * 0x9E3779B97F4A7C15ull + 1; }();
    static unsigned order[MAX_THREADS];
    for (unsigned t = 0; t < n; t++) order[t] = sched_mode == 1 ? n - 1 - t : t;
    for (;;) {
        bool progress = false;
        unsigned done = 0, at_block = 0;
        if (sched_mode == 2)
            for (unsigned i = n - 1; i > 0; i--) {  // Fisher-Yates with xorshift64
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                const unsigned j = (unsigned)(rng % (i + 1));
                const unsigned tmp = order[i]; order[i] = order[j]; order[j] = tmp;
            }
        for (unsigned oi = 0; oi < n; oi++) {
            const unsigned t = order[oi];
            Fiber &f = b.fib[t];
            if (f.state == AT_WARP_BARRIER && b.warp_gen[t >> 5] != f.warp_gen_seen) f.state = RUNNABLE;
            if (f.state == RUNNABLE) {
                b.cur = t;
                swapcontext(&b.sched, &f.ctx);
                progress = true;
            }
            // a warp barrier opens when every live lane of the warp has arrived
            if (f.state == AT_WARP_BARRIER || f.state == DONE) {
                const unsigned w = t >> 5, lo = w * 32, hi = lo + 32 < n ? lo + 32 : n;
                bool all = true, any = false;
                for (unsigned u = lo; u < hi; u++) {
                    const Fiber &g = b.fib[u];
                    if (g.state == AT_WARP_BARRIER && g.warp_gen_seen == b.warp_gen[w]) any = true;
                    else if (g.state != DONE) all = false;
                }
                if (all && any) { b.warp_gen[w]++; progress = true; }
            }
        }
        for (unsigned t = 0; t < n; t++) {
            if (b.fib[t].state == DONE) done++;
            else if (b.fib[t].state == AT_BLOCK_BARRIER) at_block++;
        }
        if (done == n) return;
        if (at_block && done + at_block == n) {  // the block barrier opens
            for (unsigned t = 0; t < n; t++) if (b.fib[t].state == AT_BLOCK_BARRIER) b.fib[t].state = RUNNABLE;
            progress = true;
        }
        if (!progress) {
            fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %u done, %u at __syncthreads, the rest parked in a warp collective\n",
                    b.bid.x, b.bid.y, b.bid.z, done, at_block);
            abort();
        }
    }
}

template <class F>
void trampoline(void *p) { (*static_cast<F *>(p))(); }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F &&f) {
    Block &b = B();
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > (unsigned)MAX_THREADS || smem > DYN_SMEM_BYTES || grid.x == 0 || grid.y == 0 || grid.z == 0) {
        fprintf(stderr, "cuda_emu: invalid launch configuration (grid %u, block %u, smem %zu)\n", grid.x, nthreads, smem);
        abort();  // the real runtime reports cudaErrorInvalidConfiguration: a bug either way
    }
    std::lock_guard<std::mutex> lock(launch_mutex());  // the block state and the `__shared__` statics are process-wide
    launch_count()++;
    b.grid = grid; b.block = block; b.nthreads = nthreads;
    typedef typename std::remove_reference<F>::type Fn;
    b.body = &trampoline<Fn>;
    b.body_arg = &f;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                b.bid = dim3(x, y, z);
                run_block();
            }
}
template <class F> void launch(dim3 grid, dim3 block, F &&f) { launch(grid, block, 0, static_cast<F &&>(f)); }

struct Idx {
    unsigned x, y, z;
};
inline Idx thread_idx() {
    const Block &b = B();
    const unsigned t = b.cur;
    return Idx{t % b.block.x, (t / b.block.x) % b.block.y, t / (b.block.x * b.block.y)};
}

inline void block_barrier() {
    Block &b = B();
    b.fib[b.cur].state = AT_BLOCK_BARRIER;
    yield_to_scheduler();
}
inline void warp_barrier() {
    Block &b = B();
    Fiber &f = b.fib[b.cur];
    f.warp_gen_seen = b.warp_gen[b.cur >> 5];
    f.state = AT_WARP_BARRIER;
    yield_to_scheduler();
}

// every lane deposits v, reads the deposit of lane src (its own if src is outside the warp or has exited)
template <class T>
T warp_exchange(T v, int src) {
    static_assert(sizeof(T) <= 16, "warp_exchange: value too large");
    Block &b = B();
    const unsigned t = b.cur, w = t >> 5, lane = t & 31;
    memcpy(b.warp_buf[w][lane], &v, sizeof(T));
    warp_barrier();
    T r = v;
    if (src >= 0 && src < 32 && w * 32 + (unsigned)src < b.nthreads && b.fib[w * 32 + src].state != DONE) memcpy(&r, b.warp_buf[w][src], sizeof(T));
    warp_barrier();
    return r;
}

}  // namespace emu

#define threadIdx (emu::thread_idx())
#define blockIdx (emu::B().bid)
#define blockDim (emu::B().block)
#define gridDim (emu::B().grid)

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }

template <class T> T __shfl_sync(unsigned, T v, int src, int width = 32) { (void)width; return emu::warp_exchange(v, src & 31); }
template <class T> T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    (void)width;
    const int lane = (int)(emu::B().cur & 31);
    return emu::warp_exchange(v, lane - (int)d >= 0 ? lane - (int)d : -1);
}
template <class T> T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    (void)width;
    const int lane = (int)(emu::B().cur & 31);
    return emu::warp_exchange(v, lane + (int)d < 32 ? lane + (int)d : -1);
}
template <class T> T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    (void)width;
    return emu::warp_exchange(v, (int)(emu::B().cur & 31) ^ m);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Block &b = emu::B();
    const unsigned t = b.cur, w = t >> 5, lane = t & 31;
    b.warp_pred[w][lane] = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned m = 0;
    for (unsigned l = 0; l < 32; l++)
        if (w * 32 + l < b.nthreads && b.fib[w * 32 + l].state != emu::DONE && b.warp_pred[w][l]) m |= 1u << l;
    emu::warp_barrier();
    return m;
}
inline unsigned __activemask() { return 0xffffffffu; }

// ------------------------------------------------------------------------------------------------------------------
// atomics (one OS thread: plain read-modify-write), returning the old value
// ------------------------------------------------------------------------------------------------------------------
template <class T, class U> T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> T atomicSub(T *p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U> T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> T atomicAnd(T *p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ------------------------------------------------------------------------------------------------------------------
// integer / floating-point intrinsics
// ------------------------------------------------------------------------------------------------------------------
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return (uint32_t)(((((uint64_t)hi << 32) | lo) << s) >> 32); }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return (uint32_t)((((uint64_t)hi << 32) | lo) >> s); }
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;  // (selector msb = sign replication is not used by the kernels)
}
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline uint32_t __brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
inline double __drcp_rn(double x) { return 1.0 / x; }
inline long long __double2ll_rn(double x) { return llrint(x); }  // default rounding mode: to nearest even
inline long long __double2ll_rz(double x) { return (long long)x; }
inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
inline double __dsqrt_rn(double x) { return sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
template <class T> T __ldg(const T *p) { return *p; }

// CUDA's min / max accept mixed integer types
template <class A, class B> typename std::common_type<A, B>::type min(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a < (T)b ? (T)a : (T)b; }
template <class A, class B> typename std::common_type<A, B>::type max(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a > (T)b ? (T)a : (T)b; }

// ------------------------------------------------------------------------------------------------------------------
// runtime API: device memory is host memory, streams are synchronous, events carry no time
// ------------------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct emu_stream_ *cudaStream_t;
typedef struct emu_event_ *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };

namespace emu {
inline int nsm() { const char *e = getenv("SIMKA_EMU_SMS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 3; }
}

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda_emu error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) { *v = a == cudaDevAttrMultiProcessorCount ? emu::nsm() : 0; return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { return cudaStreamCreateWithFlags(s, 0); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
template <class F> cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int v) { return (size_t)v <= emu::DYN_SMEM_BYTES ? cudaSuccess : cudaErrorInvalidValue; }
template <class F> cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }
