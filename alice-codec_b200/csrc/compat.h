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

namespace alice {
constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;
}  // namespace alice
