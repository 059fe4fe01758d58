// k_synth.cu — integer-only synthetic RGB volumes, generated on the
// device so that benchmarks can start with their inputs resident in HBM.
//   kind 0: G0 = the reference's make_gradient fixture (src/pipeline.rs:673-683)
//   kind 1: G1 = triangle waves + 3-bit hash dither (primary benchmark input)
//   kind 2: G2 = hash noise (incompressible worst case)
#include "kernels.h"

namespace alice {

ALICE_D uint32_t hash32(uint32_t h) {
    h *= 0x9E3779B1u; h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13; h *= 0xC2B2AE3Du; h ^= h >> 16;
    return h;
}
ALICE_D uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void k_synth(int kind, uint32_t seed, int w, int h, int f, uint8_t *__restrict__ rgb) {
    const unsigned long long n = (unsigned long long)w * h * f;
    for (unsigned long long i64 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i64 < n;
         i64 += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i64 % (unsigned)w);
        const uint32_t y = (uint32_t)((i64 / (unsigned)w) % (unsigned)h);
        const uint32_t t = (uint32_t)(i64 / ((unsigned long long)w * h));
        uint8_t *p = rgb + i64 * 3;
        if (kind == 0) {
            uint32_t v = (uint32_t)((i64 * 7) % 256);
            p[0] = (uint8_t)v; p[1] = (uint8_t)(v + 30); p[2] = (uint8_t)(v + 60);
        } else {
            uint32_t hh = hash32((uint32_t)i64 ^ seed);
            if (kind == 1) {
                int a = (int)((x + 2 * t) % 128), ta = a < 64 ? a : 127 - a;
                int b = (int)((y + 3 * t) % 96), tb = b < 48 ? b : 95 - b;
                int base = 64 + 2 * ta + tb;
                p[0] = clamp255(base + (int)(hh & 7) - 4);
                p[1] = clamp255((base >> 1) + 60 + (int)((hh >> 3) & 7) - 4);
                p[2] = clamp255(255 - base + (int)((hh >> 6) & 7) - 4);
            } else {
                p[0] = (uint8_t)(hh & 255); p[1] = (uint8_t)((hh >> 8) & 255); p[2] = (uint8_t)((hh >> 16) & 255);
            }
        }
    }
}

void synth_rgb(int kind, uint32_t seed, int w, int h, int f, uint8_t *d_rgb, cudaStream_t st) {
    const unsigned long long n = (unsigned long long)w * h * f;
    if (!n) return;
    unsigned gx = (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)device_sm_count() * 32);
    ALICE_LAUNCH(k_synth, dim3(gx), dim3(256), 0, st, kind, seed, w, h, f, d_rgb);
}

}  // namespace alice
