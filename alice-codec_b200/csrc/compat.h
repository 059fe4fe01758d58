// compat.h — one include for every translation unit of libalice_codec.
//
// Product build (nvcc, sm_100a): pulls in the CUDA runtime and defines the launch macro.
// ALICE_EMUL build (g++, tests/emul/ only): the same kernel sources are compiled against
// a fiber-based SIMT emulator so that kernel *logic* can be debugged in the GPU-less
// development container.  The emulator library is test infrastructure; the Python
// package and the C ABI product library never load it.
#pragma once

#ifdef ALICE_EMUL
#include "cuda_emul.h"   // found through -I tests/emul
#define ALICE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ::emul::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define ALICE_DYN_SMEM(name) unsigned char *name = ::emul::dyn_smem()
#define ALICE_LAUNCH_BOUNDS(t, b)
#else
#include <cuda_runtime.h>
#define ALICE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ALICE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define ALICE_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#endif

#include <stddef.h>
#include <stdint.h>

#define ALICE_HD __host__ __device__ __forceinline__
#define ALICE_D __device__ __forceinline__
// hide a value from the optimiser (keeps a multiply by a power of two a multiply: see the rANS decoder)
#ifdef ALICE_EMUL
#define ALICE_OPAQUE(v) ((void)0)
#else
#define ALICE_OPAQUE(v) asm volatile("" : "+r"(v))
#endif

// 32-bit shared-memory addresses for the serial rANS loops: a running address in a register and
// st.shared with an immediate offset keep the compiler from re-deriving the shared window base per store.
#ifdef ALICE_EMUL
typedef unsigned char *smem_addr_t;
inline smem_addr_t smem_addr_of(void *p) { return (unsigned char *)p; }
template <int OFF> inline void smem_st_u8(smem_addr_t a, uint32_t v) { a[OFF] = (unsigned char)v; }
template <int OFF> inline void smem_st_u16(smem_addr_t a, uint32_t v) { *(uint16_t *)(a + OFF) = (uint16_t)v; }
template <int OFF> inline uint32_t smem_ld_u32(smem_addr_t a) { return *(const uint32_t *)(a + OFF); }
template <int OFF> inline void smem_st_v2(smem_addr_t a, uint32_t v0, uint32_t v1) { ((uint32_t *)(a + OFF))[0] = v0; ((uint32_t *)(a + OFF))[1] = v1; }
template <int OFF> inline void smem_st_v4(smem_addr_t a, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    uint32_t *q = (uint32_t *)(a + OFF); q[0] = v0; q[1] = v1; q[2] = v2; q[3] = v3;
}
#else
typedef uint32_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr_of(void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int OFF> __device__ __forceinline__ void smem_st_u8(smem_addr_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v));
}
template <int OFF> __device__ __forceinline__ void smem_st_u16(smem_addr_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v));
}
template <int OFF> __device__ __forceinline__ uint32_t smem_ld_u32(smem_addr_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ void smem_st_v2(smem_addr_t a, uint32_t v0, uint32_t v1) {
    asm volatile("st.shared.v2.u32 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "r"(v0), "r"(v1));
}
template <int OFF> __device__ __forceinline__ void smem_st_v4(smem_addr_t a, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    asm volatile("st.shared.v4.u32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "r"(v0), "r"(v1), "r"(v2), "r"(v3));
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// Bulk asynchronous copies (cp.async.bulk, the 1-D form of the TMA engine) completing on an mbarrier: the fused
// wavelet kernels stage their input rows in shared memory several row pairs ahead without spending registers on
// loads in flight.  One elected thread arms the barrier with the byte count of a stage (mbar_arrive_expect_tx), any
// threads issue the copies (bulk_copy_g2s), every consumer waits on the phase parity (mbar_wait).
// The emulator has no asynchronous proxy: the shim copies at issue time and models the barrier's phase/transaction
// accounting, so the kernels' index and phase logic is exercised on the CPU; ordering and latency are not.
#ifdef ALICE_EMUL
struct mbar_t { unsigned completed; int pending, count; long long tx; };
inline void mbar_emul_check(mbar_t *b) {
    if (b->pending == 0 && b->tx == 0) { b->completed++; b->pending = b->count; }
}
inline void mbar_init(mbar_t *b, int count) { b->completed = 0; b->pending = b->count = count; b->tx = 0; }
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(mbar_t *b, unsigned bytes) { b->tx += bytes; b->pending--; mbar_emul_check(b); }
inline void bulk_copy_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, mbar_t *b) {
    if (((uintptr_t)dst_smem & 15) || ((uintptr_t)src_gmem & 15) || (bytes & 15)) {
        fprintf(stderr, "cuda_emul: cp.async.bulk needs 16-byte aligned addresses and size\n");
        abort();
    }
    memcpy(dst_smem, src_gmem, bytes);
    b->tx -= bytes;
    mbar_emul_check(b);
}
inline void mbar_wait(mbar_t *b, unsigned parity) {
    while ((b->completed & 1u) == (parity & 1u)) ::emul::yield();   // the phase with this parity has not completed yet
}
#else
typedef unsigned long long mbar_t;
__device__ __forceinline__ void mbar_init(mbar_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {   // make the initialised barriers visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(mbar_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, mbar_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(mbar_t *b, unsigned parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
#endif

namespace alice {
template <bool V> struct BoolTag { static constexpr bool value = V; };   // compile-time flag passed to generic lambdas
constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;
}  // namespace alice
