// k_generic.cu — element-wise and per-line kernels behind the public stage API
// (Wavelet1D/2D/3D on arbitrary i32 data incl. odd lengths, Quantizer, FastQuantizer,
// to/from_symbols, build_histogram, colour transforms, AnalyticalRDO statistics).
// These follow the reference's i32-wrapping / i64-product arithmetic literally and favour
// obviousness over speed; the .alc pipeline uses the fused kernels of k_forward.cu /
// k_inverse.cu instead.
#include <string.h>

#include "kernels.h"
#include "lifting.cuh"

namespace alice {

struct LineGeom {
    long long n_lines, n, w, h;  // n = samples per line
    int axis;
};
ALICE_D long long line_pos(const LineGeom &g, long long line, long long i) {
    if (g.axis == 0) return line * g.w + i;
    if (g.axis == 1) return (line / g.w) * g.w * g.h + (line % g.w) + i * g.w;
    return line + i * g.w * g.h;
}
ALICE_D void split_idx(const LineGeom &g, long long idx, long long half, long long &line, long long &i) {
    if (g.axis == 0) { line = idx / half; i = idx % half; }
    else { i = idx / g.n_lines; line = idx % g.n_lines; }
}

// wavelet.rs:180-217 — one lifting step, in place (each step reads only the other parity)
__global__ void k_lift_step(int32_t *__restrict__ data, LineGeom g, int coeff, int predict) {
    const long long half = g.n / 2;
    const long long total = g.n_lines * half;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long line, i;
        split_idx(g, idx, half, line, i);
        if (predict) {
            int a = data[line_pos(g, line, 2 * i)];
            int b = (2 * i + 2 < g.n) ? data[line_pos(g, line, 2 * i + 2)] : a;
            long long p = line_pos(g, line, 2 * i + 1);
            data[p] = wadd(data[p], lift_delta<true>(wadd(a, b), coeff));
        } else {
            int a = (i > 0) ? data[line_pos(g, line, 2 * i - 1)] : data[line_pos(g, line, 1)];
            int b = data[line_pos(g, line, 2 * i + 1)];
            long long p = line_pos(g, line, 2 * i);
            data[p] = wadd(data[p], lift_delta<true>(wadd(a, b), coeff));
        }
    }
}

// wavelet.rs:220-248 — (de)interleave through a zeroed temporary: an odd tail sample becomes 0
__global__ void k_reorder(const int32_t *__restrict__ in, int32_t *__restrict__ out, LineGeom g, int interleave) {
    const long long half = g.n / 2;
    const long long total = g.n_lines * half;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long line, i;
        split_idx(g, idx, half, line, i);
        if (interleave) {
            out[line_pos(g, line, 2 * i)] = in[line_pos(g, line, i)];
            out[line_pos(g, line, 2 * i + 1)] = in[line_pos(g, line, half + i)];
        } else {
            out[line_pos(g, line, i)] = in[line_pos(g, line, 2 * i)];
            out[line_pos(g, line, half + i)] = in[line_pos(g, line, 2 * i + 1)];
        }
        if (i == 0 && (g.n & 1)) out[line_pos(g, line, g.n - 1)] = 0;
    }
}

static unsigned grid_for(long long total, int block) {
    long long b = (total + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(b, (long long)device_sm_count() * 16));
}

void lift_axis(int32_t *d_data, int32_t *d_tmp, int wavelet, bool inverse, int axis, long long w, long long h,
               long long d, cudaStream_t st) {
    LineGeom g;
    g.w = w; g.h = h; g.axis = axis;
    g.n = axis == 0 ? w : (axis == 1 ? h : d);
    g.n_lines = axis == 0 ? h * d : (axis == 1 ? w * d : w * h);
    if (g.n < 2 || g.n_lines <= 0) return;  // wavelet.rs:135-137
    const long long total = g.n_lines * (g.n / 2);
    const size_t bytes = (size_t)(w * h * d) * sizeof(int32_t);
    const int nst = wavelet == WT_CDF97 ? 2 : 1;
    int cp[2], cu[2];
    if (wavelet == WT_CDF97) { cp[0] = -6497; cu[0] = -217; cp[1] = 3616; cu[1] = 1817; }
    else { cp[0] = -4096; cu[0] = wavelet == WT_CDF53 ? 1024 : 2048; cp[1] = cu[1] = 0; }
    const unsigned gx = grid_for(total, 256);
    if (!inverse) {
        for (int s = 0; s < nst; s++) {
            ALICE_LAUNCH(k_lift_step, dim3(gx), dim3(256), 0, st, d_data, g, cp[s], 1);
            ALICE_LAUNCH(k_lift_step, dim3(gx), dim3(256), 0, st, d_data, g, cu[s], 0);
        }
        ALICE_LAUNCH(k_reorder, dim3(gx), dim3(256), 0, st, d_data, d_tmp, g, 0);
        cudaMemcpyAsync(d_data, d_tmp, bytes, cudaMemcpyDeviceToDevice, st);
    } else {
        ALICE_LAUNCH(k_reorder, dim3(gx), dim3(256), 0, st, d_data, d_tmp, g, 1);
        cudaMemcpyAsync(d_data, d_tmp, bytes, cudaMemcpyDeviceToDevice, st);
        for (int s = nst - 1; s >= 0; s--) {
            ALICE_LAUNCH(k_lift_step, dim3(gx), dim3(256), 0, st, d_data, g, -cu[s], 0);
            ALICE_LAUNCH(k_lift_step, dim3(gx), dim3(256), 0, st, d_data, g, -cp[s], 1);
        }
    }
}

// color.rs:199-235
__global__ void k_rgb_to_ycocg(const uint8_t *__restrict__ rgb, int16_t *__restrict__ y, int16_t *__restrict__ co,
                               int16_t *__restrict__ cg, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int16_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        int16_t c = (int16_t)(r - b);
        int16_t t = (int16_t)(b + (c >> 1));
        int16_t gg = (int16_t)(g - t);
        y[i] = (int16_t)(t + (gg >> 1));
        co[i] = c;
        cg[i] = gg;
    }
}
// color.rs:245-276 (wrapping i16, clamp)
__global__ void k_ycocg_to_rgb(const int16_t *__restrict__ y, const int16_t *__restrict__ co,
                               const int16_t *__restrict__ cg, uint8_t *__restrict__ rgb, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int16_t t = (int16_t)(y[i] - (cg[i] >> 1));
        int16_t g = (int16_t)(cg[i] + t);
        int16_t b = (int16_t)(t - (co[i] >> 1));
        int16_t r = (int16_t)(co[i] + b);
        rgb[3 * i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        rgb[3 * i + 1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
        rgb[3 * i + 2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
    }
}
void rgb_to_ycocg(const uint8_t *d_rgb, int16_t *d_y, int16_t *d_co, int16_t *d_cg, size_t n, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_rgb_to_ycocg, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_rgb, d_y, d_co, d_cg, n);
}
void ycocg_to_rgb(const int16_t *d_y, const int16_t *d_co, const int16_t *d_cg, uint8_t *d_rgb, size_t n,
                  cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_ycocg_to_rgb, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_y, d_co, d_cg, d_rgb, n);
}

// quant.rs:89-97 Quantizer::quantize, literally (true division; *panic where Rust would panic)
__global__ void k_quantize(const int32_t *__restrict__ in, int32_t *__restrict__ out, size_t n, int step, int dz,
                           int *panic) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int v = in[i];
        int a = v < 0 ? (int)(0u - (unsigned)v) : v;
        int q = 0;
        if (!(a < dz)) {
            int num = v >= 0 ? (int)((unsigned)v - (unsigned)(dz / 2)) : (int)((unsigned)v + (unsigned)(dz / 2));
            if (step == 0 || (num == INT32_MIN && step == -1)) { *panic = 1; q = 0; }
            else q = num / step;
        }
        out[i] = q;
    }
}
// quant.rs:243-264 FastQuantizer::quantize
__global__ void k_fast_quantize(const int32_t *__restrict__ in, int32_t *__restrict__ out, size_t n, int dz,
                                unsigned long long recip, unsigned shift) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int v = in[i];
        int a = v < 0 ? (int)(0u - (unsigned)v) : v;
        int q = 0;
        if (!(a < dz)) {
            unsigned adj = (unsigned)a - (unsigned)(dz >> 1);
            unsigned long long prod = (unsigned long long)adj * recip;  // wrapping u64
            int qa = (int)(unsigned)(prod >> shift);
            q = v < 0 ? (int)(0u - (unsigned)qa) : qa;
        }
        out[i] = q;
    }
}
__global__ void k_dequantize(const int32_t *__restrict__ in, int32_t *__restrict__ out, size_t n, int step) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int q = in[i];
        out[i] = q == 0 ? 0 : (int)((unsigned)q * (unsigned)step);
    }
}
// quant.rs:555-560 / 580-588
__global__ void k_to_symbols(const int32_t *__restrict__ in, uint8_t *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int c = in[i];
        out[i] = c == 0 ? 0 : (c > 0 ? (uint8_t)((unsigned)c * 2u - 1u) : (uint8_t)((0u - (unsigned)c) * 2u));
    }
}
__global__ void k_from_symbols(const uint8_t *__restrict__ in, int32_t *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int s = in[i];
        out[i] = s == 0 ? 0 : ((s & 1) ? (s + 1) / 2 : -(s / 2));
    }
}
// quant.rs:594-600
__global__ void k_histogram(const uint8_t *__restrict__ in, size_t n, unsigned *__restrict__ hist) {
    __shared__ unsigned sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[in[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// quant.rs:422 — exact integer sum
__global__ void k_sum_i64(const int32_t *__restrict__ in, size_t n, long long *sum) {
    long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc += in[i];
    atomicAdd(reinterpret_cast<unsigned long long *>(sum), (unsigned long long)acc);
}
// quant.rs:425-432 — the f64 accumulation runs in slice order; one thread keeps that order
// (round-to-nearest mul/add without FMA contraction, like the reference's scalar code).
__global__ void k_variance_seq(const int32_t *__restrict__ in, size_t n, double mean, double *acc_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double acc = 0.0;
    for (size_t i = 0; i < n; i++) {
#ifdef ALICE_EMUL
        volatile double diff = (double)in[i] - mean;
        volatile double sq = diff * diff;
        acc = acc + sq;
#else
        double diff = __dsub_rn((double)in[i], mean);
        acc = __dadd_rn(acc, __dmul_rn(diff, diff));
#endif
    }
    *acc_out = acc;
}

// metrics.rs:16-50 — sum of squared byte differences; every term is an integer <= 65025 and the total stays below
// 2^53, so the exact integer sum equals the reference's sequential f64 sum bit for bit
__global__ void k_sq_diff_sum(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, size_t n,
                              unsigned long long *sum) {
    unsigned long long acc = 0;
    const size_t nth = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n16 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0 ? n / 16 : 0;
    for (size_t i = tid; i < n16; i += nth) {
        const uint4 va = reinterpret_cast<const uint4 *>(a)[i], vb = reinterpret_cast<const uint4 *>(b)[i];
        const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
        unsigned part = 0;                                  // 16 * 65025 fits easily
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int d = (int)((wa[k] >> (8 * j)) & 0xff) - (int)((wb[k] >> (8 * j)) & 0xff);
                part += (unsigned)(d * d);
            }
        acc += part;
    }
    for (size_t i = n16 * 16 + tid; i < n; i += nth) {
        const int d = (int)a[i] - (int)b[i];
        acc += (unsigned)(d * d);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(kFullMask, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum, acc);
}
void sq_diff_sum_u8(const uint8_t *d_a, const uint8_t *d_b, size_t n, unsigned long long *d_sum, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_sq_diff_sum, dim3(grid_for((long long)((n + 15) / 16), 256)), dim3(256), 0, st, d_a, d_b, n, d_sum);
}

// rans.rs:420-430 — InterleavedRansEncoder: symbol i belongs to stream i % 4.  planes[k * stride + j] = in[4j + k].
__global__ void k_deinterleave4(const uint8_t *__restrict__ in, size_t n, uint8_t *__restrict__ planes, size_t stride) {
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) planes[(i & 3) * stride + (i >> 2)] = in[i];
}
// rans.rs:508-520 — InterleavedRansDecoder::decode_n visits the streams round-robin and skips exhausted ones, so the
// output is a sequence of segments in each of which a fixed set of `m` streams alternates: output o of segment s is
// symbol r0 + (o - base) / m of stream act[(o - base) % m].
struct RrSegments { unsigned long long base[5]; unsigned long long r0[4]; unsigned m[4]; unsigned act[4][4]; int n_seg; };
__global__ void k_interleave_rr(const uint8_t *__restrict__ planes, size_t stride, uint8_t *__restrict__ out, size_t n,
                                RrSegments sg) {
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += nth) {
        int s = 0;
        while (s + 1 < sg.n_seg && o >= sg.base[s + 1]) s++;
        const unsigned long long local = o - sg.base[s];
        const unsigned m = sg.m[s];
        out[o] = planes[(size_t)sg.act[s][local % m] * stride + (size_t)(sg.r0[s] + local / m)];
    }
}
void deinterleave4_u8(const uint8_t *d_in, size_t n, uint8_t *d_planes, size_t stride, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_deinterleave4, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, n, d_planes, stride);
}
void interleave_rr_u8(const uint8_t *d_planes, size_t stride, uint8_t *d_out, size_t n, const unsigned long long counts[4],
                      cudaStream_t st) {
    if (!n) return;
    RrSegments sg;
    memset(&sg, 0, sizeof(sg));
    unsigned long long r = 0, base = 0;
    while (sg.n_seg < 4) {                                   // rounds [r, r_next) in which the same streams are alive
        unsigned m = 0;
        unsigned long long r_next = ~0ull;
        for (unsigned k = 0; k < 4; k++)
            if (counts[k] > r) { sg.act[sg.n_seg][m++] = k; if (counts[k] < r_next) r_next = counts[k]; }
        if (m == 0) break;
        sg.base[sg.n_seg] = base; sg.r0[sg.n_seg] = r; sg.m[sg.n_seg] = m;
        base += (r_next - r) * m;
        r = r_next;
        sg.n_seg++;
    }
    sg.base[sg.n_seg] = base;
    ALICE_LAUNCH(k_interleave_rr, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_planes, stride, d_out, n, sg);
}

namespace {
struct ScratchBlock {
    void *p = nullptr;
    size_t cap = 0;
    int dev = -1;
    bool pinned;
    explicit ScratchBlock(bool pin) : pinned(pin) {}
    void *get(size_t bytes) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (p && (cur != dev || cap < bytes)) {
            if (pinned) cudaFreeHost(p); else cudaFree(p);
            p = nullptr; cap = 0;
        }
        if (!p) {
            const size_t want = bytes < 4096 ? 4096 : bytes;
            const cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
            if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
            cap = want; dev = cur;
        }
        return p;
    }
};
}  // namespace
void *scratch_device(size_t bytes) { static thread_local ScratchBlock b(false); return b.get(bytes); }
void *scratch_pinned(size_t bytes) { static thread_local ScratchBlock b(true); return b.get(bytes); }

void quantize_i32(const int32_t *d_in, int32_t *d_out, size_t n, int step, int dz, int *d_panic, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_quantize, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, d_out, n, step, dz, d_panic);
}
void fast_quantize_i32(const int32_t *d_in, int32_t *d_out, size_t n, int dz, unsigned long long recip,
                       unsigned shift, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_fast_quantize, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, d_out, n, dz, recip,
                 shift);
}
void dequantize_i32(const int32_t *d_in, int32_t *d_out, size_t n, int step, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_dequantize, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, d_out, n, step);
}
void to_symbols_u8(const int32_t *d_in, uint8_t *d_out, size_t n, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_to_symbols, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, d_out, n);
}
void from_symbols_i32(const uint8_t *d_in, int32_t *d_out, size_t n, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_from_symbols, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, d_out, n);
}
void histogram_u8(const uint8_t *d_in, size_t n, unsigned *d_hist256, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_histogram, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, n, d_hist256);
}
void sum_i64(const int32_t *d_in, size_t n, long long *d_sum, cudaStream_t st) {
    if (!n) return;
    ALICE_LAUNCH(k_sum_i64, dim3(grid_for((long long)n, 256)), dim3(256), 0, st, d_in, n, d_sum);
}
void variance_seq_f64(const int32_t *d_in, size_t n, double mean, double *d_acc, cudaStream_t st) {
    ALICE_LAUNCH(k_variance_seq, dim3(1), dim3(32), 0, st, d_in, n, mean, d_acc);
}

}  // namespace alice
