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

// 32-bit shared-memory addresses for the serial rANS loops: a running address in a register and
// st.shared with an immediate offset keep the compiler from re-deriving the shared window base per store.
#ifdef ALICE_EMUL
typedef unsigned char *smem_addr_t;
inline smem_addr_t smem_addr_of(void *p) { return (unsigned char *)p; }
template <int OFF> inline void smem_st_u8(smem_addr_t a, uint32_t v) { a[OFF] = (unsigned char)v; }
template <int OFF> inline void smem_st_u16(smem_addr_t a, uint32_t v) { *(uint16_t *)(a + OFF) = (uint16_t)v; }
template <int OFF> inline uint32_t smem_ld_u32(smem_addr_t a) { return *(const uint32_t *)(a + OFF); }
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
#endif

namespace alice {
constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;
}  // namespace alice
