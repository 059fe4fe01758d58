// k_wavelet_i32.cu — the public Wavelet2D / Wavelet3D transforms on i32 volumes, two passes per 3-D transform.
//
// Replaces, for volumes whose width is a multiple of 4 and whose other transformed dimensions are even (every other
// shape — odd lengths lose their last sample, wavelet.rs:220-248 — keeps the step-by-step path of k_generic.cu):
//   Wavelet2D::forward / inverse    src/wavelet.rs:292-340   (rows, then columns; inverse: columns, then rows)
//   Wavelet3D::forward / inverse    src/wavelet.rs:392-484   (x, y per frame, then t; inverse: t, then y, x)
//   LosslessEncoder::transform_2d / inverse_2d   src/lossless.rs:45-54 (= Wavelet2D::cdf53)
// Arithmetic is the reference's: wrapping i32 with the i64 lifting product (lifting.cuh, WIDE = true), for arbitrary i32.
//
//   k_wxy   2-D transform of every frame in ONE pass: a warp marches a 60-pair column strip of one frame down y; a lane
//           owns two horizontal pairs, x lifting in registers (neighbour values by warp shuffle, lanes 0/31 are halo
//           lanes), y lifting as a streaming state machine per owned column; 16-byte loads, 8-byte stores (forward) or
//           the reverse (inverse); the next row pair is loaded before the current one is transformed.
//   k_wt    temporal transform: a thread streams the temporal line of four adjacent columns (16-byte accesses, coalesced
//           across x for every frame).
// Both are out of place: a 3-D transform goes data -> tmp (xy) -> data (t), 16 B per sample of HBM traffic against the
// 8 B of an ideal in-place transform and the ~120 B of the step-by-step path.
#include "kernels.h"
#include "lifting.cuh"

namespace alice {

constexpr int kWxyVP = 60;   // valid pairs per strip (30 lanes x 2 pairs)

template <int WT, bool INV, bool EDGE>
ALICE_D void wxy_strip(const int32_t *__restrict__ src, int32_t *__restrict__ dst, int w, int halfx, int halfy, int p0, int i0,
                       int i1, bool lane_ok) {
    constexpr int NST = WaveletTraits<WT>::NST;
    const int js = max(0, i0 - NST), je = min(halfy, i1 + NST);
    const bool in_row = p0 >= 0 && p0 + 2 <= halfx;   // this lane's two pairs lie inside the row (else: halo lane outside the image)
    if (!INV) {
        FwdLift<WT, true> L[4];   // columns 0,1 = low-x, 2,3 = high-x
        auto load_pair = [&](int j, int4 (&r)[2]) {
#pragma unroll
            for (int q = 0; q < 2; q++)
                r[q] = in_row ? __ldg(reinterpret_cast<const int4 *>(src + (size_t)(2 * j + q) * w + 2 * p0)) : make_int4(0, 0, 0, 0);
        };
        auto row_x = [&](const int4 &r, int (&v)[4]) {
            int e[2] = {r.x, r.z}, o[2] = {r.y, r.w};
            fwd_lanes<WT, true, 2, EDGE>(e, o, p0, halfx);
            v[0] = e[0]; v[1] = e[1]; v[2] = o[0]; v[3] = o[1];
        };
        auto emit = [&](int jo, const int (&lo)[4], const int (&hi)[4]) {
            if (!lane_ok || jo < i0 || jo >= i1) return;
            int32_t *rl = dst + (size_t)jo * w, *rh = dst + (size_t)(halfy + jo) * w;
            *reinterpret_cast<int2 *>(rl + p0) = make_int2(lo[0], lo[1]);
            *reinterpret_cast<int2 *>(rl + halfx + p0) = make_int2(lo[2], lo[3]);
            *reinterpret_cast<int2 *>(rh + p0) = make_int2(hi[0], hi[1]);
            *reinterpret_cast<int2 *>(rh + halfx + p0) = make_int2(hi[2], hi[3]);
        };
        int4 cur[2], nxt[2];
        if (js < je) load_pair(js, cur);
        int k = 0;
        for (int j = js; j < je; j++, k++) {
            load_pair(min(j + 1, je - 1), nxt);
            int v0[4], v1[4], lo[4], hi[4];
            row_x(cur[0], v0);
            row_x(cur[1], v1);
            bool has = false;
#pragma unroll
            for (int i = 0; i < 4; i++) has = L[i].push(v0[i], v1[i], k, j, lo[i], hi[i]);
            if (has) emit(j - NST, lo, hi);
            cur[0] = nxt[0]; cur[1] = nxt[1];
        }
        if (je == halfy && k > 0) {
#pragma unroll
            for (int which = 0; which < NST; which++) {
                int lo[4], hi[4];
                bool has = false;
#pragma unroll
                for (int i = 0; i < 4; i++) has = L[i].flush(k, which, halfy, lo[i], hi[i]);
                if (has) emit(halfy - NST + which, lo, hi);
            }
        }
    } else {
        InvLift<WT, true> L[4];
        auto load_pair = [&](int j, int2 (&r)[4]) {
            const int32_t *rl = src + (size_t)j * w, *rh = src + (size_t)(halfy + j) * w;
            if (in_row) {
                r[0] = __ldg(reinterpret_cast<const int2 *>(rl + p0));
                r[1] = __ldg(reinterpret_cast<const int2 *>(rl + halfx + p0));
                r[2] = __ldg(reinterpret_cast<const int2 *>(rh + p0));
                r[3] = __ldg(reinterpret_cast<const int2 *>(rh + halfx + p0));
            } else r[0] = r[1] = r[2] = r[3] = make_int2(0, 0);
        };
        // one reconstructed image row from its x-subband values (all lanes take part in the shuffles)
        auto emit_row = [&](int y, bool active, const int (&v)[4]) {
            int e[2] = {v[0], v[1]}, o[2] = {v[2], v[3]};
            inv_lanes<WT, true, 2, EDGE>(e, o, p0, halfx);
            if (active && lane_ok) *reinterpret_cast<int4 *>(dst + (size_t)y * w + 2 * p0) = make_int4(e[0], o[0], e[1], o[1]);
        };
        auto emit = [&](int jo, const int (&ev)[4], const int (&od)[4]) {
            const bool active = jo >= i0 && jo < i1;
            emit_row(2 * jo, active, ev);
            emit_row(2 * jo + 1, active, od);
        };
        int2 cur[4], nxt[4];
        if (js < je) load_pair(js, cur);
        int k = 0;
        for (int j = js; j < je; j++, k++) {
            load_pair(min(j + 1, je - 1), nxt);
            const int lo[4] = {cur[0].x, cur[0].y, cur[1].x, cur[1].y}, hi[4] = {cur[2].x, cur[2].y, cur[3].x, cur[3].y};
            int ev[4], od[4];
            bool has = false;
#pragma unroll
            for (int i = 0; i < 4; i++) has = L[i].push(lo[i], hi[i], k, j, ev[i], od[i]);
            if (has) emit(j - NST, ev, od);
#pragma unroll
            for (int i = 0; i < 4; i++) cur[i] = nxt[i];
        }
        if (je == halfy && k > 0) {
#pragma unroll
            for (int which = 0; which < NST; which++) {
                int ev[4], od[4];
                bool has = false;
#pragma unroll
                for (int i = 0; i < 4; i++) has = L[i].flush(k, which, halfy, ev[i], od[i]);
                if (has) emit(halfy - NST + which, ev, od);
            }
        }
    }
}

template <int WT, bool INV>
__global__ void ALICE_LAUNCH_BOUNDS(128, 4)
k_wxy(const int32_t *__restrict__ src, int32_t *__restrict__ dst, int w, int h, long long n_frames, int n_strips, int n_segs,
      int seg_pairs) {
    const int lane = threadIdx.x & 31;
    const long long warp_g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)n_strips * n_segs * n_frames;
    if (warp_g >= n_warps) return;   // warp-uniform exit; the kernel has no block-level barrier
    const int sx = (int)(warp_g % n_strips);
    const int sg = (int)((warp_g / n_strips) % n_segs);
    const long long t = warp_g / ((long long)n_strips * n_segs);
    const int halfx = w >> 1, halfy = h >> 1;
    const int p0 = sx * kWxyVP - 2 + 2 * lane;
    const bool lane_ok = lane >= 1 && lane <= 30 && p0 < halfx;
    const int i0 = sg * seg_pairs, i1 = min(halfy, i0 + seg_pairs);
    const size_t fo = (size_t)t * w * h;
    const bool edge = sx == 0 || sx * kWxyVP + 62 >= halfx;   // some lane owns pair 0 or pair halfx-1 (the mirrored ones)
    if (edge) wxy_strip<WT, INV, true>(src + fo, dst + fo, w, halfx, halfy, p0, i0, i1, lane_ok);
    else wxy_strip<WT, INV, false>(src + fo, dst + fo, w, halfx, halfy, p0, i0, i1, lane_ok);
}

template <int WT, bool INV>
__global__ void ALICE_LAUNCH_BOUNDS(256, 3)
k_wt(const int32_t *__restrict__ src, int32_t *__restrict__ dst, size_t frame_sz, int halft) {
    constexpr int NST = WaveletTraits<WT>::NST;
    const long long n_items = (long long)(frame_sz / 4);
    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
         item += (long long)gridDim.x * blockDim.x) {
        const size_t off = (size_t)item * 4;
        auto ld = [&](int fr) { return __ldg(reinterpret_cast<const int4 *>(src + (size_t)fr * frame_sz + off)); };
        auto st = [&](int fr, const int (&v)[4]) {
            *reinterpret_cast<int4 *>(dst + (size_t)fr * frame_sz + off) = make_int4(v[0], v[1], v[2], v[3]);
        };
        // INV = false: pair j = frames (2j, 2j+1) in, (j, halft + j) out; INV = true: the other way round
        auto load_pair = [&](int j, int4 &a, int4 &b) {
            a = ld(INV ? j : 2 * j);
            b = ld(INV ? halft + j : 2 * j + 1);
        };
        auto emit = [&](int jo, const int (&p)[4], const int (&q)[4]) {
            st(INV ? 2 * jo : jo, p);
            st(INV ? 2 * jo + 1 : halft + jo, q);
        };
        int4 a, b, na, nb;
        load_pair(0, a, b);
        int k = 0;
        if (!INV) {
            FwdLift<WT, true> L[4];
            for (int j = 0; j < halft; j++, k++) {
                load_pair(min(j + 1, halft - 1), na, nb);
                const int e[4] = {a.x, a.y, a.z, a.w}, o[4] = {b.x, b.y, b.z, b.w};
                int lo[4], hi[4];
                bool has = false;
#pragma unroll
                for (int i = 0; i < 4; i++) has = L[i].push(e[i], o[i], k, j, lo[i], hi[i]);
                if (has) emit(j - NST, lo, hi);
                a = na; b = nb;
            }
#pragma unroll
            for (int which = 0; which < NST; which++) {
                int lo[4], hi[4];
                bool has = false;
#pragma unroll
                for (int i = 0; i < 4; i++) has = L[i].flush(k, which, halft, lo[i], hi[i]);
                if (has) emit(halft - NST + which, lo, hi);
            }
        } else {
            InvLift<WT, true> L[4];
            for (int j = 0; j < halft; j++, k++) {
                load_pair(min(j + 1, halft - 1), na, nb);
                const int lo[4] = {a.x, a.y, a.z, a.w}, hi[4] = {b.x, b.y, b.z, b.w};
                int ev[4], od[4];
                bool has = false;
#pragma unroll
                for (int i = 0; i < 4; i++) has = L[i].push(lo[i], hi[i], k, j, ev[i], od[i]);
                if (has) emit(j - NST, ev, od);
                a = na; b = nb;
            }
#pragma unroll
            for (int which = 0; which < NST; which++) {
                int ev[4], od[4];
                bool has = false;
#pragma unroll
                for (int i = 0; i < 4; i++) has = L[i].flush(k, which, halft, ev[i], od[i]);
                if (has) emit(halft - NST + which, ev, od);
            }
        }
    }
}

bool wavelet_fast_eligible(const int32_t *a, const int32_t *b, long long w, long long h, long long d, int ndim) {
    if (ndim < 2 || w < 8 || (w & 3) || h < 2 || (h & 1)) return false;
    if (ndim >= 3 && (d < 2 || (d & 1))) return false;
    if (w > (1 << 30) || h > (1 << 30) || d > (1 << 30)) return false;
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

template <int WT, bool INV>
static void launch_wxy(const int32_t *src, int32_t *dst, int w, int h, long long n_frames, cudaStream_t st) {
    const int halfx = w / 2, halfy = h / 2;
    const int n_strips = (halfx + kWxyVP - 1) / kWxyVP;
    // enough warps to fill the machine several times over, segments of >= 16 row pairs
    const long long base_warps = (long long)n_strips * n_frames;
    int n_segs = (int)std::min<long long>(std::max<long long>(1, ((long long)device_sm_count() * 96 + base_warps - 1) / base_warps),
                                          std::max(1, halfy / 16));
    const int seg_pairs = (halfy + n_segs - 1) / n_segs;
    n_segs = (halfy + seg_pairs - 1) / seg_pairs;
    const long long n_warps = base_warps * n_segs;
    auto k = k_wxy<WT, INV>;
    ALICE_LAUNCH(k, dim3((unsigned)((n_warps + 3) / 4)), dim3(128), 0, st, src, dst, w, h, n_frames, n_strips, n_segs, seg_pairs);
}
template <int WT, bool INV>
static void launch_wt(const int32_t *src, int32_t *dst, size_t frame_sz, int d, cudaStream_t st) {
    const long long items = (long long)(frame_sz / 4);
    const unsigned gx = (unsigned)std::min<long long>((items + 255) / 256, (long long)device_sm_count() * 12);
    auto k = k_wt<WT, INV>;
    ALICE_LAUNCH(k, dim3(gx), dim3(256), 0, st, src, dst, frame_sz, d / 2);
}

#define ALICE_WT_SWITCH(CALL_F, CALL_I)                                                     \
    switch (wavelet) {                                                                      \
    case WT_CDF53: if (inverse) { CALL_I(WT_CDF53); } else { CALL_F(WT_CDF53); } break;     \
    case WT_CDF97: if (inverse) { CALL_I(WT_CDF97); } else { CALL_F(WT_CDF97); } break;     \
    default:       if (inverse) { CALL_I(WT_HAAR); } else { CALL_F(WT_HAAR); } break;       \
    }

void wavelet_xy_i32(int wavelet, bool inverse, const int32_t *d_src, int32_t *d_dst, int w, int h, long long n_frames,
                    cudaStream_t st) {
#define F_(WT) launch_wxy<WT, false>(d_src, d_dst, w, h, n_frames, st)
#define I_(WT) launch_wxy<WT, true>(d_src, d_dst, w, h, n_frames, st)
    ALICE_WT_SWITCH(F_, I_)
#undef F_
#undef I_
}
void wavelet_t_i32(int wavelet, bool inverse, const int32_t *d_src, int32_t *d_dst, size_t frame_sz, int d, cudaStream_t st) {
#define F_(WT) launch_wt<WT, false>(d_src, d_dst, frame_sz, d, st)
#define I_(WT) launch_wt<WT, true>(d_src, d_dst, frame_sz, d, st)
    ALICE_WT_SWITCH(F_, I_)
#undef F_
#undef I_
}

}  // namespace alice
