// cuda_emul.h — minimal fiber-based SIMT emulator (TEST INFRASTRUCTURE, dev container only).
//
// Lets the CUDA kernel sources of alice-codec_b200/csrc be compiled with plain g++
// (-DALICE_EMUL) and executed on the CPU, one thread block at a time, every CUDA thread a
// ucontext fiber, so that indexing / barrier / shuffle logic can be debugged where no GPU
// exists.  It is NOT a product path: the package never loads a library built with it,
// the -m gpu parity tests never touch it, and no performance number comes from it.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct short2 { short x, y; };
struct __attribute__((aligned(8))) short4 { short x, y, z, w; };
static inline int2 make_int2(int a, int b) { return {a, b}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
static inline short4 make_short4(short a, short b, short c, short d) { return {a, b, c, d}; }

namespace emul {
struct Fiber {
    ucontext_t uc;
    char *stack = nullptr;
    bool done = false;
};
struct Warp {
    int arrived = 0, alive = 0;
    unsigned gen = 0;
    uint64_t slot[2][32];
    unsigned char par[32];
};
struct Block {
    int nthreads = 0, alive = 0, bar_arrived = 0;
    unsigned bar_gen = 0;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    ucontext_t sched;
    int cur = 0;
    const std::function<void()> *body = nullptr;
    dim3 bdim;
};
inline Block *&cur_block() { static Block *b = nullptr; return b; }
inline unsigned char *&dyn_smem_ptr() { static unsigned char *p = nullptr; return p; }
inline unsigned char *dyn_smem() { return dyn_smem_ptr(); }
}  // namespace emul

inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emul {
inline void set_thread(Block &b, int i) {
    threadIdx.x = i % b.bdim.x;
    threadIdx.y = (i / b.bdim.x) % b.bdim.y;
    threadIdx.z = i / (b.bdim.x * b.bdim.y);
}
inline void yield() {
    Block &b = *cur_block();
    int me = b.cur;
    swapcontext(&b.fibers[me].uc, &b.sched);
    set_thread(b, me);
}
inline void block_release_if_ready(Block &b) {
    if (b.alive > 0 && b.bar_arrived == b.alive) { b.bar_arrived = 0; b.bar_gen++; }
}
inline void warp_release_if_ready(Warp &w) {
    if (w.alive > 0 && w.arrived == w.alive) { w.arrived = 0; w.gen++; }
}
inline void sync_block() {
    Block &b = *cur_block();
    unsigned my = b.bar_gen;
    b.bar_arrived++;
    block_release_if_ready(b);
    while (b.bar_gen == my) yield();
}
inline void sync_warp() {
    Block &b = *cur_block();
    Warp &w = b.warps[b.cur / 32];
    unsigned my = w.gen;
    w.arrived++;
    warp_release_if_ready(w);
    while (w.gen == my) yield();
}
inline void fiber_entry() {
    Block &b = *cur_block();
    int me = b.cur;
    set_thread(b, me);
    (*b.body)();
    b.cur = me;
    b.fibers[me].done = true;
    b.alive--;
    Warp &w = b.warps[me / 32];
    w.alive--;
    block_release_if_ready(b);
    warp_release_if_ready(w);
    swapcontext(&b.fibers[me].uc, &b.sched);
}
inline void run_block(Block &b, const std::function<void()> &body) {
    static std::vector<char *> stacks;
    const size_t kStack = 256 * 1024;
    while ((int)stacks.size() < b.nthreads) stacks.push_back((char *)malloc(kStack));
    b.fibers.assign(b.nthreads, Fiber());
    b.warps.assign((b.nthreads + 31) / 32, Warp());
    b.alive = b.nthreads;
    b.bar_arrived = 0;
    b.bar_gen = 0;
    b.body = &body;
    for (int i = 0; i < b.nthreads; i++) {
        b.warps[i / 32].alive++;
        memset(b.warps[i / 32].par, 0, 32);
        Fiber &f = b.fibers[i];
        getcontext(&f.uc);
        f.uc.uc_stack.ss_sp = stacks[i];
        f.uc.uc_stack.ss_size = kStack;
        f.uc.uc_link = nullptr;
        makecontext(&f.uc, (void (*)())fiber_entry, 0);
    }
    cur_block() = &b;
    int remaining = b.nthreads;
    long idle_passes = 0;
    while (remaining > 0) {
        int before_alive = b.alive;
        unsigned before_gen = b.bar_gen;
        unsigned wg = 0;
        for (auto &w : b.warps) wg += w.gen;
        for (int i = 0; i < b.nthreads; i++) {
            if (b.fibers[i].done) continue;
            b.cur = i;
            swapcontext(&b.sched, &b.fibers[i].uc);
        }
        remaining = b.alive;
        unsigned wg2 = 0;
        for (auto &w : b.warps) wg2 += w.gen;
        if (b.alive == before_alive && b.bar_gen == before_gen && wg == wg2) {
            if (++idle_passes > 4) {
                fprintf(stderr, "cuda_emul: deadlock (divergent barrier/shuffle) in block (%u,%u,%u)\n",
                        blockIdx.x, blockIdx.y, blockIdx.z);
                abort();
            }
        } else idle_passes = 0;
    }
    cur_block() = nullptr;
}
inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    std::vector<unsigned char> dyn(smem + 16);
    dyn_smem_ptr() = (unsigned char *)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    gridDim = grid;
    blockDim = block;
    Block b;
    b.bdim = block;
    b.nthreads = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx = {bx, by, bz};
                run_block(b, body);
            }
}
template <class T> inline uint64_t to_bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
template <class T> inline T shfl_from(T v, int src) {
    Block &b = *cur_block();
    int lane = b.cur % 32;
    Warp &w = b.warps[b.cur / 32];
    int p = w.par[lane];
    w.slot[p][lane] = to_bits(v);
    sync_warp();
    T r = from_bits<T>(w.slot[p][src & 31]);
    w.par[lane] ^= 1;
    return r;
}
inline int lane_id() { return cur_block()->cur % 32; }
}  // namespace emul

inline void __syncthreads() { emul::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emul::sync_warp(); }
template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emul::shfl_from(v, src); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int l = emul::lane_id();
    int src = l - (int)d;
    return emul::shfl_from(v, src < 0 ? l : src);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int l = emul::lane_id();
    int src = l + (int)d;
    return emul::shfl_from(v, src > 31 ? l : src);
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return emul::shfl_from(v, emul::lane_id() ^ m);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emul::Block &b = *emul::cur_block();
    int lane = b.cur % 32;
    emul::Warp &w = b.warps[b.cur / 32];
    int p = w.par[lane];
    w.slot[p][lane] = pred ? 1 : 0;
    emul::sync_warp();
    unsigned r = 0;
    int base = (b.cur / 32) * 32;
    for (int i = 0; i < 32 && base + i < b.nthreads; i++)
        if (!b.fibers[base + i].done && w.slot[p][i]) r |= 1u << i;
    w.par[lane] ^= 1;
    return r;
}
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
inline unsigned __activemask() { return 0xffffffffu; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (s & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    return (unsigned)((((((uint64_t)hi) << 32) | lo) << (s & 31)) >> 32);
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((v >> (8 * (sel & 7))) & 0xff);
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0;
        r |= byte << (8 * i);
    }
    return r;
}
template <class T> inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;
template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
inline unsigned atomicAdd(unsigned *p, int v) { unsigned o = *p; *p = o + (unsigned)v; return o; }
template <class T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; *p = o > v ? o : v; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }

// ------------------------------------------------------------------ runtime API subset
typedef int cudaError_t;
typedef struct emul_stream_t *cudaStream_t;
typedef struct emul_event { double t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0, cudaHostAllocDefault = 0 };
inline const char *cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
enum { cudaHostRegisterDefault = 0, cudaHostRegisterPortable = 1 };
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emul_event{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; int major, minor; char name[256]; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 4; p->major = 10; p->minor = 0; strcpy(p->name, "emul"); return cudaSuccess; }
