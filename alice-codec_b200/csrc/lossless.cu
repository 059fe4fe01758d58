// lossless.cu — BASELINE config 4 on the device: the reference's lossless module (src/lossless.rs) over a 64-frame set.
//
// LosslessEncoder::transform_2d / inverse_2d (lossless.rs:45-54) are Wavelet2D::cdf53 forward / inverse on one i32 image;
// the reference has no lossless *codec* around them.  Config 4 strings the reference's own stages
// together per colour channel of a w x h x f RGB volume:
//   rgb_bytes_to_ycocg_r (color.rs:199-235) -> i32 planes -> transform_2d of every frame -> to_symbols of the coefficients
//   (quant.rs:547-563: step 1, `as u8` wraps) -> build_histogram -> FrequencyTable::from_histogram -> RansEncoder
//   -> RansDecoder -> inverse_2d of the coefficients.
// One rANS stream per (frame set, channel), like the .alc pipeline.  Everything stays on the device; the stage buffers can
// be read back for the parity tests.
#include <new>

#include "engine.h"
#include "lifting.cuh"

namespace alice {

__global__ void k_widen_i16(const int16_t *__restrict__ in, int32_t *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

struct Lossless {
    uint32_t w = 0, h = 0, f = 0;
    size_t n = 0;                       // samples per channel
    cudaStream_t st = nullptr;
    int16_t *planes16 = nullptr;        // [3][n]
    int32_t *planes = nullptr;          // [3][n]  YCoCg-R as i32; after decode: the inverse transform of the coefficients
    int32_t *coefs = nullptr;           // [3][n]
    int32_t *tmp = nullptr;             // [n] scratch for shapes off the fast path
    uint8_t *symbols = nullptr;         // [3][n]
    uint8_t *decoded = nullptr;         // [3][n]
    uint8_t *payload = nullptr;         // [3][cap]
    size_t cap = 0;
    unsigned *hist = nullptr;           // [3][256]
    EncSym *enc = nullptr;
    uint32_t *dec_lut = nullptr;
    DecAux *aux = nullptr;
    RansEncJob *enc_jobs = nullptr;
    RansDecJob *dec_jobs = nullptr;
    unsigned long long *results = nullptr;
    unsigned long long h_results[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool encoded = false;
    ~Lossless() {
        cudaFree(planes16); cudaFree(planes); cudaFree(coefs); cudaFree(tmp); cudaFree(symbols); cudaFree(decoded);
        cudaFree(payload); cudaFree(hist); cudaFree(enc); cudaFree(dec_lut); cudaFree(aux); cudaFree(enc_jobs);
        cudaFree(dec_jobs); cudaFree(results);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        cudaGetLastError();
    }
};

#define LL_TRY(expr)                                                                       \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            set_error(kErrCuda, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
            return kErrCuda;                                                               \
        }                                                                                  \
    } while (0)

static bool ll_alloc(Lossless *L) {
    const size_t n = L->n;
    L->cap = rans_enc_worst_case(n);
    auto a = [&](auto &p, size_t bytes) { return cudaMalloc((void **)&p, bytes ? bytes : 16) == cudaSuccess; };
    bool ok = a(L->planes16, 3 * n * 2) && a(L->planes, 3 * n * 4) && a(L->coefs, 3 * n * 4) && a(L->tmp, n * 4) &&
              a(L->symbols, 3 * n) && a(L->decoded, 3 * n) && a(L->payload, 3 * L->cap) && a(L->hist, 3 * 256 * 4) &&
              a(L->enc, 3 * kEncTableBytes) && a(L->dec_lut, 3 * kDecLutEntries * 4) && a(L->aux, 3 * sizeof(DecAux)) &&
              a(L->enc_jobs, 3 * sizeof(RansEncJob)) && a(L->dec_jobs, 3 * sizeof(RansDecJob)) && a(L->results, 6 * 8);
    for (auto &e : L->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    if (!ok) { cudaGetLastError(); set_error(kErrCuda, "device memory allocation failed (lossless)"); }
    return ok;
}

// 2-D CDF 5/3 of every frame of one channel: src -> dst (out of place), any shape
static int ll_transform(Lossless *L, bool inverse, int32_t *src, int32_t *dst) {
    if (wavelet_fast_eligible(src, dst, L->w, L->h, L->f, 2)) {
        wavelet_xy_i32(WT_CDF53, inverse, src, dst, (int)L->w, (int)L->h, (long long)L->f, L->st);
        return kOk;
    }
    LL_TRY(cudaMemcpyAsync(dst, src, L->n * 4, cudaMemcpyDeviceToDevice, L->st));
    const size_t fs = (size_t)L->w * L->h;
    for (uint32_t t = 0; t < L->f; t++) {
        lift_axis(dst + t * fs, L->tmp, WT_CDF53, inverse, inverse ? 1 : 0, L->w, L->h, 1, L->st);
        lift_axis(dst + t * fs, L->tmp, WT_CDF53, inverse, inverse ? 0 : 1, L->w, L->h, 1, L->st);
    }
    return kOk;
}

}  // namespace alice

using namespace alice;
#pragma GCC visibility push(default)
extern "C" {

struct AliceLossless { Lossless L; };

AliceLossless *alice_codec_lossless_create(uint32_t w, uint32_t h, uint32_t f, void *cuda_stream) {
    set_error(0, "");
    if (!cuda_ready()) return nullptr;
    const unsigned __int128 tot = (unsigned __int128)w * h * f;
    if (tot == 0 || tot > 0xffffffffull) { set_error(kErrDimensions, "empty or oversized volume"); return nullptr; }
    AliceLossless *o = new (std::nothrow) AliceLossless();
    if (!o) return nullptr;
    o->L.w = w; o->L.h = h; o->L.f = f;
    o->L.n = (size_t)tot;
    o->L.st = (cudaStream_t)cuda_stream;
    if (!ll_alloc(&o->L)) { delete o; return nullptr; }
    return o;
}
void alice_codec_lossless_destroy(AliceLossless *o) { delete o; }

// d_rgb: device pointer, interleaved RGB [f][h][w][3].  Synchronises the stream before returning.
int alice_codec_lossless_encode_device(AliceLossless *o, const uint8_t *d_rgb) {
    set_error(0, "");
    if (!o || !d_rgb) { set_error(kErrNull, "null argument"); return kErrNull; }
    Lossless *L = &o->L;
    const size_t n = L->n;
    L->encoded = false;
    LL_TRY(cudaEventRecord(L->ev[0], L->st));
    rgb_to_ycocg(d_rgb, L->planes16, L->planes16 + n, L->planes16 + 2 * n, n, L->st);
    {
        const unsigned gx = (unsigned)std::min<size_t>((3 * n + 255) / 256, (size_t)device_sm_count() * 32);
        ALICE_LAUNCH(k_widen_i16, dim3(gx), dim3(256), 0, L->st, L->planes16, L->planes, 3 * n);
    }
    LL_TRY(cudaEventRecord(L->ev[1], L->st));
    for (int c = 0; c < 3; c++) {
        int rc = ll_transform(L, false, L->planes + c * n, L->coefs + c * n);
        if (rc) return rc;
    }
    LL_TRY(cudaEventRecord(L->ev[2], L->st));
    LL_TRY(cudaMemsetAsync(L->hist, 0, 3 * 256 * 4, L->st));
    for (int c = 0; c < 3; c++) {
        to_symbols_u8(L->coefs + c * n, L->symbols + c * n, n, L->st);
        histogram_u8(L->symbols + c * n, n, L->hist + c * 256, L->st);
    }
    build_tables(L->hist, 3, 256, L->enc, L->dec_lut, L->aux, nullptr, nullptr, nullptr, L->st);
    LL_TRY(cudaEventRecord(L->ev[3], L->st));
    RansEncJob jobs[3];
    for (int c = 0; c < 3; c++) jobs[c] = RansEncJob{L->symbols + c * n, (unsigned long long)n, L->payload + c * L->cap, (unsigned long long)L->cap};
    LL_TRY(cudaMemcpyAsync(L->enc_jobs, jobs, sizeof(jobs), cudaMemcpyHostToDevice, L->st));
    rans_encode(L->enc_jobs, L->enc, L->hist, L->results, 3, L->st);
    LL_TRY(cudaEventRecord(L->ev[4], L->st));
    LL_TRY(cudaMemcpyAsync(L->h_results, L->results, sizeof(L->h_results), cudaMemcpyDeviceToHost, L->st));
    LL_TRY(cudaStreamSynchronize(L->st));
    LL_TRY(cudaGetLastError());
    for (int c = 0; c < 3; c++) {
        if (L->h_results[2 * c + 1] & 2) { set_error(kErrPanic, "symbol with zero frequency in use: the reference aborts on this input"); return kErrPanic; }
        if (L->h_results[2 * c + 1]) { set_error(kErrCuda, "rANS output overflow"); return kErrCuda; }
    }
    cudaEventElapsedTime(&L->ms[0], L->ev[0], L->ev[1]);   // colour + widen
    cudaEventElapsedTime(&L->ms[1], L->ev[1], L->ev[2]);   // 2-D forward transform, 3 channels x f frames
    cudaEventElapsedTime(&L->ms[2], L->ev[2], L->ev[3]);   // symbols + histograms + tables
    cudaEventElapsedTime(&L->ms[3], L->ev[3], L->ev[4]);   // rANS encode, 3 streams
    L->encoded = true;
    return kOk;
}

// rANS-decodes the three streams of the last encode and applies inverse_2d to the coefficients.
int alice_codec_lossless_decode_device(AliceLossless *o) {
    set_error(0, "");
    if (!o) { set_error(kErrNull, "null argument"); return kErrNull; }
    Lossless *L = &o->L;
    if (!L->encoded) { set_error(kErrBufferSize, "lossless decode without a preceding encode"); return kErrBufferSize; }
    const size_t n = L->n;
    RansDecJob jobs[3];
    for (int c = 0; c < 3; c++) {
        const unsigned long long len = L->h_results[2 * c];
        jobs[c] = RansDecJob{L->payload + c * L->cap + (L->cap - len), len, L->decoded + c * n, (unsigned long long)n};
    }
    LL_TRY(cudaMemcpyAsync(L->dec_jobs, jobs, sizeof(jobs), cudaMemcpyHostToDevice, L->st));
    LL_TRY(cudaEventRecord(L->ev[5], L->st));
    rans_decode(L->dec_jobs, L->dec_lut, L->aux, 3, L->st);
    LL_TRY(cudaEventRecord(L->ev[6], L->st));
    for (int c = 0; c < 3; c++) {
        int rc = ll_transform(L, true, L->coefs + c * n, L->planes + c * n);
        if (rc) return rc;
    }
    LL_TRY(cudaEventRecord(L->ev[7], L->st));
    LL_TRY(cudaStreamSynchronize(L->st));
    LL_TRY(cudaGetLastError());
    cudaEventElapsedTime(&L->ms[4], L->ev[5], L->ev[6]);   // rANS decode, 3 streams
    cudaEventElapsedTime(&L->ms[5], L->ev[6], L->ev[7]);   // 2-D inverse transform
    return kOk;
}

// which: 0 coefficients (i32 [3][n]), 1 symbols (u8 [3][n]), 2 histograms (u32 [3][256]), 3 decoded symbols (u8 [3][n]),
// 4 inverse-transformed planes (i32 [3][n]), 5 stream c (bytes; *out_len receives its length, channel in `channel`).
int alice_codec_lossless_fetch(AliceLossless *o, int which, int channel, void *host_out, uint64_t cap, uint64_t *out_len) {
    set_error(0, "");
    if (!o || !host_out || !out_len) { set_error(kErrNull, "null argument"); return kErrNull; }
    Lossless *L = &o->L;
    const size_t n = L->n;
    const void *src = nullptr;
    size_t bytes = 0;
    switch (which) {
    case 0: src = L->coefs; bytes = 3 * n * 4; break;
    case 1: src = L->symbols; bytes = 3 * n; break;
    case 2: src = L->hist; bytes = 3 * 256 * 4; break;
    case 3: src = L->decoded; bytes = 3 * n; break;
    case 4: src = L->planes; bytes = 3 * n * 4; break;
    case 5:
        if (channel < 0 || channel > 2) { set_error(kErrDimensions, "channel out of range"); return kErrDimensions; }
        bytes = (size_t)L->h_results[2 * channel];
        src = L->payload + channel * L->cap + (L->cap - bytes);
        break;
    default: set_error(kErrDimensions, "unknown buffer"); return kErrDimensions;
    }
    *out_len = bytes;
    if (cap < bytes) { set_error(kErrBufferSize, "output buffer too small"); return kErrBufferSize; }
    LL_TRY(cudaMemcpyAsync(host_out, src, bytes, cudaMemcpyDeviceToHost, L->st));
    LL_TRY(cudaStreamSynchronize(L->st));
    return kOk;
}
// CUDA-event durations (ms) of the last encode / decode: [0] colour, [1] 2-D forward, [2] symbols + histograms + tables,
// [3] rANS encode, [4] rANS decode, [5] 2-D inverse; [6], [7] reserved
int alice_codec_lossless_timings(AliceLossless *o, float *ms8) {
    if (!o || !ms8) return kErrNull;
    for (int i = 0; i < 8; i++) ms8[i] = o->L.ms[i];
    return kOk;
}

}  // extern "C"
#pragma GCC visibility pop
