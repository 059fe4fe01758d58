// k_inverse.cu — decode back-end: three u8 symbol planes -> interleaved RGB u8.
//
// Replaces, for one chunk (reference file:line):
//   from_symbols                    src/quant.rs:572-590
//   Quantizer::dequantize_buffer    src/quant.rs:104-110, 135-146 (q * step, wrapping)
//   Wavelet3D::inverse              src/wavelet.rs:441-484 (t, then per frame y, then x)
//   crop + `as i16`                 src/pipeline.rs:602-611
//   ycocg_r_to_rgb_bytes            src/color.rs:245-276 (i16 wrapping, clamp to u8)
//
// The decoder must reproduce the reference for ANY header (arbitrary quant_step) and any
// symbol plane — including the garbage the reference decodes from its own malformed
// frequency tables — so everything here is the reference's wrapping i32
// arithmetic with the i64 lifting product (WIDE=true), and the t->y hand-off is an i32 volume.
//
//   k_inv_t     one thread per VEC adjacent coefficients: symbols -> coefficient, streaming
//               inverse temporal lifting, writes frames t < f of i32 [3][f][ph][pw].
//   k_inv_yx    mirror of k_fwd_xy: one warp per column strip marching down y; streaming
//               inverse y lifting per owned column, lane-parallel inverse x lifting with
//               shuffles, then crop, i16 truncation, inverse colour transform and RGB store.
#include <stdlib.h>

#include "kernels.h"
#include "lifting.cuh"

namespace alice {

ALICE_D int sym_to_coef(uint32_t s, int step) {
    // quant.rs:580-588 then quant.rs:104-110
    int q = (s == 0) ? 0 : ((s & 1) ? (int)((s + 1) >> 1) : -(int)(s >> 1));
    return (int)((unsigned)q * (unsigned)step);
}

// WIDE = true : the reference's wrapping i32 arithmetic with the i64 lifting product, i32 hand-off (any header).
// WIDE = false: 32-bit products and an i16 hand-off, exact while 128 * |quant_step| <= kNarrowMaxCoef (see launcher).
// hand-off type between the t pass and the y/x pass: i32 in both variants.  An i16 hand-off halves the traffic but
// measured slower twice: 2.62 vs 1.65 ms per 1080p x 64 chunk in round 1 and, with the prefetching k_inv_yx,
// 1.549 vs 1.058 ms in round 2 (profiles/r02_switches.md); it was dropped.
template <bool WIDE> struct Handoff { typedef int32_t T; };
// k_inv_yx prefetches the next row pair, takes unchecked 8-byte loads on interior strips and runs the steady-state lifting
// form after the warm-up (1.600 -> 1.060 ms for the whole back-end, profiles/r01_ab_backend.jsonl); 64-frame chunks use the
// rolled, software-pipelined compile-time-depth variant of k_inv_t (1.666 -> 1.600 ms).

template <int WT, int VEC, int PF, bool WIDE>
__global__ void ALICE_LAUNCH_BOUNDS(256, (PF ? 3 : 4))
k_inv_t(const uint8_t *__restrict__ symbols, void *__restrict__ coef_v, int pw, int ph, int f, int pf, int step0,
        int step1, int step2) {
    typedef typename Handoff<WIDE>::T HT;
    HT *coef = reinterpret_cast<HT *>(coef_v);
    constexpr int NST = WaveletTraits<WT>::NST;
    const int c = blockIdx.z;
    const int step = c == 0 ? step0 : (c == 1 ? step1 : step2);
    const int halft = PF ? PF / 2 : (pf >> 1);   // PF != 0: compile-time depth, the streaming state machine unrolls away
    const size_t frame_sz = (size_t)ph * pw;
    const uint8_t *src = symbols + (size_t)c * pf * frame_sz;
    HT *dst = coef + (size_t)c * f * frame_sz;
    const long long n_items = (long long)(frame_sz / VEC);

    auto store_vec = [&](HT *p, const int (&v)[VEC]) {
        if (sizeof(HT) == 4) {
            if (VEC == 4) *reinterpret_cast<int4 *>(p) = make_int4(v[0], v[1], v[VEC - 2], v[VEC - 1]);
            else *reinterpret_cast<int2 *>(p) = make_int2(v[0], v[1]);
        } else {
            const uint32_t w0 = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
            if (VEC == 4) {
                const uint32_t w1 = (uint32_t)(uint16_t)v[VEC - 2] | ((uint32_t)(uint16_t)v[VEC - 1] << 16);
                *reinterpret_cast<uint2 *>(p) = make_uint2(w0, w1);
            } else *reinterpret_cast<uint32_t *>(p) = w0;
        }
    };
    auto emit = [&](size_t off, int jo, const int (&ev)[VEC], const int (&od)[VEC]) {
        const int t0 = 2 * jo, t1 = 2 * jo + 1;
        if (t0 < f) store_vec(dst + (size_t)t0 * frame_sz + off, ev);
        if (t1 < f) store_vec(dst + (size_t)t1 * frame_sz + off, od);
    };

    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
         item += (long long)gridDim.x * blockDim.x) {
        const size_t off = (size_t)item * VEC;
        InvLift<WT, WIDE> L[VEC];
        int k = 0;
        struct RawPair { uint32_t a, b; };              // VEC symbols of frame j (low) and of frame halft + j (high)
        auto load_at = [&](const uint8_t *pl, const uint8_t *ph_) {
            RawPair r;
            if (VEC == 4) {
                r.a = __ldg(reinterpret_cast<const uint32_t *>(pl));
                r.b = __ldg(reinterpret_cast<const uint32_t *>(ph_));
            } else {
                r.a = __ldg(reinterpret_cast<const uint16_t *>(pl));
                r.b = __ldg(reinterpret_cast<const uint16_t *>(ph_));
            }
            return r;
        };
        if (PF != 0) {
            // compile-time depth: unrolled prologue (warm-up, mirrored left edge), then a ROLLED steady-state loop with
            // the symbols of the next two pairs in flight (the fully unrolled form spilled 1.4-1.9 KB per thread and
            // overflowed the instruction cache; the generic runtime loop below tests k and j in every iteration and
            // waits on every load: long_scoreboard was its dominant stall)
            static_assert(PF == 0 || PF / 2 >= NST + 4, "compile-time depth too short for the pipelined form");
            const uint8_t *pl = src + off, *ph_ = src + (size_t)halft * frame_sz + off;
            RawPair r0 = load_at(pl, ph_);
            pl += frame_sz; ph_ += frame_sz;
            RawPair r1 = load_at(pl, ph_);
            pl += frame_sz; ph_ += frame_sz;
            RawPair r2 = r1;
            auto step_pair = [&](int j, bool steady) {
                int ev[VEC], od[VEC];
                bool has = steady;
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    const int lo = sym_to_coef((r0.a >> (8 * i)) & 0xff, step);
                    const int hi = sym_to_coef((r0.b >> (8 * i)) & 0xff, step);
                    if (steady) L[i].push_steady(lo, hi, ev[i], od[i]);
                    else has = L[i].push(lo, hi, j, j, ev[i], od[i]);
                }
                if (has) emit(off, j - NST, ev, od);
                r0 = r1; r1 = r2;
            };
#pragma unroll
            for (int j = 0; j <= NST; j++) {
                r2 = load_at(pl, ph_);                  // pair j + 2 <= NST + 2 < halft
                pl += frame_sz; ph_ += frame_sz;
                step_pair(j, false);
            }
#pragma unroll 1
            for (int j = NST + 1; j < halft - 2; j++) {
                r2 = load_at(pl, ph_);                  // pair j + 2 <= halft - 1
                pl += frame_sz; ph_ += frame_sz;
                step_pair(j, true);
            }
            step_pair(halft - 2, true);
            step_pair(halft - 1, true);
            k = halft;
        } else
#pragma unroll
        for (int j = 0; j < halft; j++, k++) {
            const RawPair r = load_at(src + (size_t)j * frame_sz + off, src + (size_t)(halft + j) * frame_sz + off);
            const uint32_t a = r.a, b = r.b;
            int ev[VEC], od[VEC];
            bool has = false;
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                int lo = sym_to_coef((a >> (8 * i)) & 0xff, step);
                int hi = sym_to_coef((b >> (8 * i)) & 0xff, step);
                has = L[i].push(lo, hi, k, j, ev[i], od[i]);
            }
            if (has) emit(off, j - NST, ev, od);
        }
#pragma unroll
        for (int which = 0; which < NST; which++) {
            int ev[VEC], od[VEC];
            bool has = false;
#pragma unroll
            for (int i = 0; i < VEC; i++) has = L[i].flush(k, which, halft, ev[i], od[i]);
            if (has) emit(off, halft - NST + which, ev, od);
        }
    }
}

template <int M, class HT>
ALICE_D void load_group_i32(const HT *__restrict__ row, int xp, int limit, int *v) {
    if (xp >= 0 && M == 2 && xp + 2 <= limit && ((reinterpret_cast<uintptr_t>(row + xp) & (2 * sizeof(HT) - 1)) == 0)) {
        if (sizeof(HT) == 4) {
            int2 t = __ldg(reinterpret_cast<const int2 *>(row + xp));
            v[0] = t.x;
            v[1] = t.y;
        } else {
            const uint32_t t = __ldg(reinterpret_cast<const uint32_t *>(row + xp));
            v[0] = (int16_t)(t & 0xffff);
            v[1] = (int)t >> 16;
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < M; i++) {
        int x = xp + i;
        v[i] = (x >= 0 && x < limit) ? __ldg(row + x) : 0;
    }
}

ALICE_D uint32_t clamp_u8(int16_t v) { return v < 0 ? 0u : (v > 255 ? 255u : (uint32_t)v); }

template <int WT, int M, bool WIDE>
__global__ void ALICE_LAUNCH_BOUNDS(128, 3)
k_inv_yx(const void *__restrict__ coef_v, uint8_t *__restrict__ rgb, int w, int h, int f, int pw, int ph,
         int n_strips, int n_segs, int seg_pairs, int vec_ok) {
    constexpr int NST = WaveletTraits<WT>::NST;
    constexpr int PXL = 2 * M;
    constexpr int VPAIRS = 30 * M;
    const int lane = threadIdx.x & 31;
    const long long warp_g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)n_strips * n_segs * f;
    if (warp_g >= n_warps) return;
    const int sx = (int)(warp_g % n_strips);
    const int sg = (int)((warp_g / n_strips) % n_segs);
    const int t = (int)(warp_g / ((long long)n_strips * n_segs));
    const int halfx = pw >> 1, halfy = ph >> 1;
    const int p0 = sx * VPAIRS - M + lane * M;
    const int x0 = 2 * p0;
    const bool lane_ok = lane >= 1 && lane <= 30;
    const int i0 = sg * seg_pairs;
    const int i1 = min(halfy, i0 + seg_pairs);
    const int js = max(0, i0 - NST);
    const int je = min(halfy, i1 + NST);

    typedef typename Handoff<WIDE>::T HT;
    const HT *coef = reinterpret_cast<const HT *>(coef_v);
    InvLift<WT, WIDE> L[3][PXL];  // per channel: columns [0,M) = low-x, [M,2M) = high-x
    const size_t plane_sz = (size_t)f * ph * pw;
    const HT *in_t = coef + (size_t)t * ph * pw;
    uint8_t *frame = rgb + (size_t)t * w * h * 3;

    // one reconstructed image row y from its x-subband values (all lanes take part in the shuffles)
    auto emit_row = [&](int y, bool active, int (&val)[3][PXL]) {
        int px[3][PXL];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            int e[M], o[M];
#pragma unroll
            for (int i = 0; i < M; i++) { e[i] = val[c][i]; o[i] = val[c][M + i]; }
            inv_lanes<WT, WIDE, M>(e, o, p0, halfx);
#pragma unroll
            for (int i = 0; i < M; i++) { px[c][2 * i] = e[i]; px[c][2 * i + 1] = o[i]; }
        }
        if (!active || !lane_ok || y >= h || x0 < 0 || x0 >= w) return;
        uint32_t bytes[PXL * 3];
#pragma unroll
        for (int i = 0; i < PXL; i++) {
            // pipeline.rs:608 `as i16`, then color.rs:266-273 in wrapping i16
            int16_t yy = (int16_t)px[0][i], co = (int16_t)px[1][i], cg = (int16_t)px[2][i];
            int16_t tt = (int16_t)(yy - (cg >> 1));
            int16_t g = (int16_t)(cg + tt);
            int16_t b = (int16_t)(tt - (co >> 1));
            int16_t r = (int16_t)(co + b);
            bytes[3 * i] = clamp_u8(r);
            bytes[3 * i + 1] = clamp_u8(g);
            bytes[3 * i + 2] = clamp_u8(b);
        }
        uint8_t *dst = frame + ((size_t)y * w + x0) * 3;
        if (vec_ok && x0 + PXL <= w) {
            constexpr int NW = PXL * 3 / 4;
#pragma unroll
            for (int q = 0; q < NW; q++)
                reinterpret_cast<uint32_t *>(dst)[q] =
                    bytes[4 * q] | (bytes[4 * q + 1] << 8) | (bytes[4 * q + 2] << 16) | (bytes[4 * q + 3] << 24);
        } else {
#pragma unroll
            for (int i = 0; i < PXL; i++)
                if (x0 + i < w) {
                    dst[3 * i] = (uint8_t)bytes[3 * i];
                    dst[3 * i + 1] = (uint8_t)bytes[3 * i + 1];
                    dst[3 * i + 2] = (uint8_t)bytes[3 * i + 2];
                }
        }
    };

    // The row pair j + 1 is loaded before pair j is transformed (the kernel's dominant stall was long_scoreboard: every
    // iteration waited for its own loads), strips whose 32 lanes are all inside the row use plain 8-byte loads without
    // range or alignment tests, and after the warm-up the lifting state machine runs its branch-free steady form.
    const bool fast_ld = M == 2 && sx * VPAIRS - M >= 0 && sx * VPAIRS - M + 32 * M <= halfx &&
                         (halfx & 1) == 0 && (pw & 1) == 0 && (reinterpret_cast<uintptr_t>(coef) & 7) == 0;
    auto load_rows = [&](int j, int (&lo)[3][PXL], int (&hi)[3][PXL]) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const HT *row_lo = in_t + c * plane_sz + (size_t)j * pw;
            const HT *row_hi = in_t + c * plane_sz + (size_t)(halfy + j) * pw;
            if (fast_ld) {
#ifdef ALICE_EMUL
                if (((reinterpret_cast<uintptr_t>(row_lo + p0) | reinterpret_cast<uintptr_t>(row_lo + halfx + p0) |
                      reinterpret_cast<uintptr_t>(row_hi + p0) | reinterpret_cast<uintptr_t>(row_hi + halfx + p0)) & (2 * sizeof(HT) - 1)) != 0 ||
                    p0 < 0 || p0 + M > halfx || j < 0 || j >= halfy)
                    abort();   // the emulator does not fault on what the GPU would
#endif
                {
                    const int2 a = __ldg(reinterpret_cast<const int2 *>(row_lo + p0));
                    const int2 b = __ldg(reinterpret_cast<const int2 *>(row_lo + halfx + p0));
                    const int2 c2 = __ldg(reinterpret_cast<const int2 *>(row_hi + p0));
                    const int2 d = __ldg(reinterpret_cast<const int2 *>(row_hi + halfx + p0));
                    lo[c][0] = a.x; lo[c][1] = a.y; lo[c][PXL - 2] = b.x; lo[c][PXL - 1] = b.y;
                    hi[c][0] = c2.x; hi[c][1] = c2.y; hi[c][PXL - 2] = d.x; hi[c][PXL - 1] = d.y;
                }
            } else {
                int t0[M], t1[M], t2[M], t3[M];
                load_group_i32<M, HT>(row_lo, p0, halfx, t0);
                load_group_i32<M, HT>(row_lo + halfx, p0, halfx, t1);
                load_group_i32<M, HT>(row_hi, p0, halfx, t2);
                load_group_i32<M, HT>(row_hi + halfx, p0, halfx, t3);
#pragma unroll
                for (int i = 0; i < M; i++) { lo[c][i] = t0[i]; lo[c][M + i] = t1[i]; hi[c][i] = t2[i]; hi[c][M + i] = t3[i]; }
            }
        }
    };
    auto emit_pair = [&](int jo, int (&ev)[3][PXL], int (&od)[3][PXL]) {
        const bool active = jo >= i0 && jo < i1;
        emit_row(2 * jo, active, ev);
        emit_row(2 * jo + 1, active, od);
    };
    int k = 0;
    int j = js;
    int clo[3][PXL], chi[3][PXL], nlo[3][PXL], nhi[3][PXL];
    if (js < je) load_rows(js, clo, chi);
    auto rotate = [&]() {
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int i = 0; i < PXL; i++) { clo[c][i] = nlo[c][i]; chi[c][i] = nhi[c][i]; }
    };
    // warm-up: the first NST + 1 pairs go through the general push (mirrored left edge when js == 0)
    for (; j < je && k <= NST; j++, k++) {
        load_rows(min(j + 1, je - 1), nlo, nhi);
        int ev[3][PXL], od[3][PXL];
        bool has = false;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int i = 0; i < PXL; i++) has = L[c][i].push(clo[c][i], chi[c][i], k, j, ev[c][i], od[c][i]);
        if (has) emit_pair(j - NST, ev, od);
        rotate();
    }
    for (; j < je; j++, k++) {
        load_rows(min(j + 1, je - 1), nlo, nhi);
        int ev[3][PXL], od[3][PXL];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int i = 0; i < PXL; i++) L[c][i].push_steady(clo[c][i], chi[c][i], ev[c][i], od[c][i]);
        emit_pair(j - NST, ev, od);
        rotate();
    }
    if (je == halfy) {
#pragma unroll
        for (int which = 0; which < NST; which++) {
            int ev[3][PXL], od[3][PXL];
            bool has = false;
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int i = 0; i < PXL; i++) has = L[c][i].flush(k, which, halfy, ev[c][i], od[c][i]);
            if (has) {
                const int jo = halfy - NST + which;
                const bool active = jo >= i0 && jo < i1;
                emit_row(2 * jo, active, ev);
                emit_row(2 * jo + 1, active, od);
            }
        }
    }
}

// Largest |coefficient| = 128 * |quant_step| (|q| <= 128 for u8 symbols) for which the narrow variant is exact.
// Worst-case gains of the linear part of the inverse lifting (max abs row sums over every intermediate stage,
// n = 64; rounding adds < 1 per step): values 2.2744 / 2.0 / 2.0 per axis and 1.6663 / 2.0 / 2.0 at the end of an
// axis, neighbour sums (a+b) 4.549 / 2.5 / 3.0 (CDF 9/7 / CDF 5/3 / Haar).  With A = 15000: the i16 hand-off after
// the t pass holds <= 2.0 A = 30000; the largest product in the x pass is 4.549 * 1.6663^2 A * 6497 = 1.2e9
// (9/7), 3.0 * 4 A * 4096 = 7.4e8 (Haar) < 2^31.
constexpr long long kNarrowMaxCoef = kInvNarrowMaxCoef;   // kernels.h

template <int WT, bool WIDE>
static void launch_inv(const uint8_t *d_symbols, int32_t *d_coef, uint8_t *d_rgb, int w, int h, int f, int pw, int ph,
                       int pf, const int steps[3], cudaStream_t st) {
    const size_t frame_sz = (size_t)pw * ph;
    {
        const int vec = (pw % 4 == 0) ? 4 : 2;
        const long long items = frame_sz / vec;
        const unsigned gx = (unsigned)std::min<long long>((items + 255) / 256, (long long)device_sm_count() * 8);
        const dim3 tgrid(gx, 1, 3), tblock(256);
        // (a FULLY unrolled PF = 64 instantiation spilled 1.4-1.9 KB per thread and measured 2.3x slower than the
        //  runtime loop; the rolled compile-time-depth form is 4 % faster than the runtime loop on the whole back-end)
        if (vec == 4 && pf == 64) {
            auto kt = k_inv_t<WT, 4, 64, WIDE>;
            ALICE_LAUNCH(kt, tgrid, tblock, 0, st, d_symbols, d_coef, pw, ph, f, pf, steps[0], steps[1], steps[2]);
        } else if (vec == 4) {
            auto kt = k_inv_t<WT, 4, 0, WIDE>;
            ALICE_LAUNCH(kt, tgrid, tblock, 0, st, d_symbols, d_coef, pw, ph, f, pf, steps[0], steps[1], steps[2]);
        } else {
            auto kt = k_inv_t<WT, 2, 0, WIDE>;
            ALICE_LAUNCH(kt, tgrid, tblock, 0, st, d_symbols, d_coef, pw, ph, f, pf, steps[0], steps[1], steps[2]);
        }
    }
    constexpr int M = 2;
    const int halfx = pw / 2, halfy = ph / 2;
    const int n_strips = (halfx + 30 * M - 1) / (30 * M);
    long long base_warps = (long long)n_strips * f;
#ifndef ALICE_YX_TARGET_WARPS
#define ALICE_YX_TARGET_WARPS 96   // as in k_forward.cu: finer segments, one warp per block (1.88 -> 1.70 ms measured)
#endif
#ifndef ALICE_YX_WPB
#define ALICE_YX_WPB 1
#endif
    int n_segs = (int)std::min<long long>(std::max<long long>(1, (device_sm_count() * ALICE_YX_TARGET_WARPS + base_warps - 1) / base_warps),
                                          std::max(1, halfy / 16));
    int seg_pairs = (halfy + n_segs - 1) / n_segs;
    n_segs = (halfy + seg_pairs - 1) / seg_pairs;
    const long long n_warps = (long long)n_strips * n_segs * f;
    const int vec_ok = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_rgb) & 3) == 0);
    const int warps_per_block = ALICE_YX_WPB;
    dim3 grid((unsigned)((n_warps + warps_per_block - 1) / warps_per_block));
    auto kyx = k_inv_yx<WT, M, WIDE>;
    ALICE_LAUNCH(kyx, grid, dim3(32 * warps_per_block), 0, st, d_coef, d_rgb, w, h, f, pw, ph, n_strips, n_segs,
                 seg_pairs, vec_ok);
}

void inverse_backend(int wavelet, const uint8_t *d_symbols, int32_t *d_coef, uint8_t *d_rgb, int w, int h, int f,
                     int pw, int ph, int pf, const int steps[3], cudaStream_t st) {
    bool narrow = true;   // every channel's coefficients stay within the 32-bit / i16 bounds above
    for (int c = 0; c < 3; c++) {
        const long long a = steps[c] < 0 ? -(long long)steps[c] : (long long)steps[c];
        if (128 * a > kNarrowMaxCoef) narrow = false;
    }
#define ALICE_INV(WT)                                                                                            \
    do {                                                                                                         \
        if (narrow) launch_inv<WT, false>(d_symbols, d_coef, d_rgb, w, h, f, pw, ph, pf, steps, st);            \
        else launch_inv<WT, true>(d_symbols, d_coef, d_rgb, w, h, f, pw, ph, pf, steps, st);                    \
    } while (0)
    switch (wavelet) {
    case WT_CDF53: ALICE_INV(WT_CDF53); break;
    case WT_CDF97: ALICE_INV(WT_CDF97); break;
    default:       ALICE_INV(WT_HAAR); break;
    }
#undef ALICE_INV
}

}  // namespace alice
