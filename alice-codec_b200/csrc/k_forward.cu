// k_forward.cu — encode front-end: RGB u8 -> three u8 symbol planes + three histograms.
//
// Replaces, for one chunk (reference file:line):
//   rgb_bytes_to_ycocg_r            src/color.rs:199-235
//   pad_channel_to_i32              src/pipeline.rs:77-114 (replicate last col/row/frame)
//   Wavelet3D::forward              src/wavelet.rs:392-438 (x, then y per frame, then t)
//   Quantizer::quantize_buffer      src/quant.rs:89-128   (dead zone = step)
//   to_symbols                      src/quant.rs:547-563  (zig-zag, `as u8` wrap)
//   build_histogram                 src/quant.rs:594-600
//
// Two kernels:
//   k_fwd_xy        one warp marches one 30*M-pair-wide column strip of one frame down y.
//                   Each lane owns M horizontal pairs (2M pixels): colour transform and the
//                   x lifting stay in registers (neighbour values by warp shuffle, lanes 0/31
//                   are halo lanes); the y lifting is a streaming FwdLift per owned column.
//                   Output: i16 planes [3][f][ph][pw], de-interleaved in x and y.
//                   i16 is exact: |coef| <= 7043 after x and y for u8 input.
//   k_fwd_t_quant   one thread per VEC horizontally adjacent coefficients streams the
//                   temporal line (padded frames re-read frame f-1), then quantises, maps to
//                   symbols, stores u8 and histograms through shared-memory atomics.
// All arithmetic is 32-bit: for u8 input every product (a+b)*c stays below 2^31.
#include "kernels.h"
#include "lifting.cuh"
#include "quant.cuh"

#ifndef ALICE_FWD_M
#define ALICE_FWD_M 2   // horizontal pairs per lane in k_fwd_xy (1: 96 registers, 2: 150-205 registers)
#endif

// (Measured and dropped, profiles/r02_switches.md: one colour channel per warp -- 96 instead of 205 registers, RGB read
//  three times -- 1.173 ms against 1.110 ms for the whole front-end; unrolling the row-pair loop by 2 / 4: 1.31 / 1.82 ms.)

namespace alice {

// ------------------------------------------------------------------------------ k_fwd_xy
// Raw RGB of one lane's 2M pixels of one row, as 3M/2 little-endian words (byte 3i = R of pixel i, ...).
template <int M> struct RowRaw { uint32_t w[(6 * M + 3) / 4]; };

// EDGE = false: the caller guarantees 0 <= x0, x0 + 2M <= w and 4-byte alignment of the row -> plain loads.
template <int M, bool EDGE>
ALICE_D void load_row_raw(const uint8_t *__restrict__ row, int x0, int w, bool vec_ok, RowRaw<M> &raw) {
    constexpr int PXL = 2 * M;
    constexpr int NB = PXL * 3;          // bytes per lane: 6 (M = 1) or 12 (M = 2)
    constexpr int NW = (NB + 3) / 4;
    if (!EDGE || (vec_ok && x0 >= 0 && x0 + PXL <= w)) {
        if (NB % 4 == 0) {
            const uint32_t *p = reinterpret_cast<const uint32_t *>(row + (size_t)x0 * 3);
#pragma unroll
            for (int i = 0; i < NW; i++) raw.w[i] = __ldg(p + i);
        } else {                         // 6 bytes at a 2-byte aligned address (x0 is even)
            const uint16_t *p = reinterpret_cast<const uint16_t *>(row + (size_t)x0 * 3);
            raw.w[0] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 16);
            raw.w[1] = (uint32_t)__ldg(p + 2);
        }
    } else {
#pragma unroll
        for (int i = 0; i < NW; i++) raw.w[i] = 0;
#pragma unroll
        for (int i = 0; i < PXL; i++) {
            int x = x0 + i;
            x = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);  // replicate-pad (pipeline.rs:95-98)
            const uint8_t *p = row + (size_t)x * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) raw.w[(3 * i + ch) >> 2] |= (uint32_t)__ldg(p + ch) << (8 * ((3 * i + ch) & 3));
        }
    }
}

template <int M, bool EDGE>
ALICE_D void store_group_i16(int16_t *__restrict__ dst, const int *v, int xp, int limit, bool lane_ok) {
    // M consecutive coefficients at columns xp .. xp+M-1, valid while column < limit.
    if (!lane_ok) return;
    if (!EDGE) {
        if (M == 2) *reinterpret_cast<uint32_t *>(dst + xp) = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[M - 1] << 16);
        else dst[xp] = (int16_t)v[0];
        return;
    }
    if (xp < 0) return;
    if (M == 2 && xp + 2 <= limit && ((reinterpret_cast<uintptr_t>(dst + xp) & 3) == 0)) {
        uint32_t pk = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
        *reinterpret_cast<uint32_t *>(dst + xp) = pk;
        return;
    }
#pragma unroll
    for (int i = 0; i < M; i++)
        if (xp + i < limit) dst[xp + i] = (int16_t)v[i];
}

// One column strip of one frame, rows [i0, i1) of the y-transformed output.  EDGE = false: the strip touches neither
// the left nor the right image border, every access is in range and aligned.
// NCH = 3: the warp transforms all three colour channels.
template <int WT, int M, bool EDGE, int NCH>
ALICE_D void fwd_xy_strip(const uint8_t *__restrict__ frame, int16_t *__restrict__ out_t, size_t plane_sz, int w, int h,
                          int pw, int p0, int i0, int i1, int lane, bool vec_ok, int c0) {
    constexpr int NST = WaveletTraits<WT>::NST;
    constexpr int PXL = 2 * M;
    const int halfx = pw >> 1, halfy = (h + (h & 1)) >> 1;
    const int x0 = 2 * p0;
    constexpr int HL = (NST + M - 1) / M;     // halo lanes per side: a lane's result needs NST pairs on either side
    const bool lane_ok = lane >= HL && lane < 32 - HL;
    const int js = max(0, i0 - NST);
    const int je = min(halfy, i1 + NST);

    FwdLift<WT, false> L[NCH][PXL];  // per channel: columns [0,M) = low-x, [M,2M) = high-x

    auto load_pair = [&](int j, RowRaw<M> (&raw)[2]) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int y = min(2 * j + r, h - 1);  // replicate-pad the extra row (pipeline.rs:100-105)
            load_row_raw<M, EDGE>(frame + (size_t)y * w * 3, x0, w, vec_ok, raw[r]);
        }
    };
    // colour transform + x lifting of one row: v[c][0..M) = low-x, v[c][M..2M) = high-x
    auto row_x = [&](const RowRaw<M> &raw, int (&v)[NCH][PXL]) {
        int e[NCH][M], o[NCH][M];
#pragma unroll
        for (int i = 0; i < PXL; i++) {
            const int r = (raw.w[(3 * i) >> 2] >> (8 * ((3 * i) & 3))) & 0xff;
            const int g = (raw.w[(3 * i + 1) >> 2] >> (8 * ((3 * i + 1) & 3))) & 0xff;
            const int bb = (raw.w[(3 * i + 2) >> 2] >> (8 * ((3 * i + 2) & 3))) & 0xff;
            // color.rs:221-232 (values fit i16, so i32 arithmetic is identical)
            const int co = r - bb;
            const int tt = bb + (co >> 1);
            const int cg = g - tt;
            const int yy = tt + (cg >> 1);
            if (NCH == 3) {
                if (i & 1) { o[0][i >> 1] = yy; o[NCH - 2][i >> 1] = co; o[NCH - 1][i >> 1] = cg; }
                else       { e[0][i >> 1] = yy; e[NCH - 2][i >> 1] = co; e[NCH - 1][i >> 1] = cg; }
            } else {
                const int one = c0 == 0 ? yy : (c0 == 1 ? co : cg);
                if (i & 1) o[0][i >> 1] = one; else e[0][i >> 1] = one;
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            fwd_lanes<WT, false, M, EDGE>(e[c], o[c], p0, halfx);
#pragma unroll
            for (int i = 0; i < M; i++) { v[c][i] = e[c][i]; v[c][M + i] = o[c][i]; }
        }
    };
    auto emit = [&](int jo, const int (&lo)[NCH][PXL], const int (&hi)[NCH][PXL]) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            int16_t *row_lo = out_t + (c0 + c) * plane_sz + (size_t)jo * pw;
            int16_t *row_hi = out_t + (c0 + c) * plane_sz + (size_t)(halfy + jo) * pw;
            store_group_i16<M, EDGE>(row_lo, &lo[c][0], p0, halfx, lane_ok);
            store_group_i16<M, EDGE>(row_lo + halfx, &lo[c][M], p0, halfx, lane_ok);
            store_group_i16<M, EDGE>(row_hi, &hi[c][0], p0, halfx, lane_ok);
            store_group_i16<M, EDGE>(row_hi + halfx, &hi[c][M], p0, halfx, lane_ok);
        }
    };

    RowRaw<M> cur[2], nxt[2];
    if (js < je) load_pair(js, cur);
    // warm-up and top-edge rows: the general streaming step (outputs before i0 are inexact and skipped)
    const int j_main = min(je, max(js + NST + 1, i0 + NST));
    int k = 0, j = js;
    for (; j < j_main; j++, k++) {
        if (j + 1 < je) load_pair(j + 1, nxt);
        int v0[NCH][PXL], v1[NCH][PXL], lo[NCH][PXL], hi[NCH][PXL];
        row_x(cur[0], v0);
        row_x(cur[1], v1);
        bool has = false;
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int i = 0; i < PXL; i++) has = L[c][i].push(v0[c][i], v1[c][i], k, j, lo[c][i], hi[c][i]);
        if (has && j - NST >= i0 && j - NST < i1) emit(j - NST, lo, hi);
        cur[0] = nxt[0]; cur[1] = nxt[1];
    }
    // steady state: the next row pair is already in flight while this one is transformed
    for (; j < je; j++, k++) {
        load_pair(min(j + 1, je - 1), nxt);
        int v0[NCH][PXL], v1[NCH][PXL], lo[NCH][PXL], hi[NCH][PXL];
        row_x(cur[0], v0);
        row_x(cur[1], v1);
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int i = 0; i < PXL; i++) L[c][i].push_steady(v0[c][i], v1[c][i], lo[c][i], hi[c][i]);
        emit(j - NST, lo, hi);
        cur[0] = nxt[0]; cur[1] = nxt[1];
    }
    if (je == halfy && k > 0) {
#pragma unroll
        for (int which = 0; which < NST; which++) {
            int lo[NCH][PXL], hi[NCH][PXL];
            bool has = false;
#pragma unroll
            for (int c = 0; c < NCH; c++)
#pragma unroll
                for (int i = 0; i < PXL; i++) has = L[c][i].flush(k, which, halfy, lo[c][i], hi[c][i]);
            const int jo = halfy - NST + which;
            if (has && jo >= i0 && jo < i1) emit(jo, lo, hi);
        }
    }
}

template <int WT, int M, int NCH>
__global__ void ALICE_LAUNCH_BOUNDS(128, (NCH == 1 ? 5 : (M == 1 ? 4 : (WT == WT_CDF97 ? 2 : 3))))
k_fwd_xy(const uint8_t *__restrict__ rgb, int16_t *__restrict__ planes, int w, int h, int f, int pw, int ph,
         int n_strips, int n_segs, int seg_pairs, int vec_ok) {
    constexpr int HL = (WaveletTraits<WT>::NST + M - 1) / M;
    constexpr int VPAIRS = (32 - 2 * HL) * M;
    const int lane = threadIdx.x & 31;
    const long long warp_g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)n_strips * n_segs * f;
    if (warp_g >= n_warps) return;  // warp-uniform exit; the kernel has no block-level barrier
    const int sx = (int)(warp_g % n_strips);
    const int sg = (int)((warp_g / n_strips) % n_segs);
    const int t = (int)(warp_g / ((long long)n_strips * n_segs));
    const int halfx = pw >> 1, halfy = ph >> 1;
    const int p0 = sx * VPAIRS - HL * M + lane * M;
    const int i0 = sg * seg_pairs;
    const int i1 = min(halfy, i0 + seg_pairs);
    const uint8_t *frame = rgb + (size_t)t * w * h * 3;
    const size_t plane_sz = (size_t)f * ph * pw;
    int16_t *out_t = planes + (size_t)t * ph * pw;
    // interior strip: no lane owns pair 0 or pair halfx-1 (the mirrored ones), all 32 lanes read inside the row
    const bool interior = vec_ok && sx > 0 && (sx + 1) * VPAIRS + HL * M <= halfx - 1 && (pw == w) && ((pw & 3) == 0);  // rows 4-byte aligned
    const int c0 = NCH == 1 ? (int)blockIdx.y : 0;   // one channel per warp: the channel is the grid's y index
    if (interior) fwd_xy_strip<WT, M, false, NCH>(frame, out_t, plane_sz, w, h, pw, p0, i0, i1, lane, true, c0);
    else fwd_xy_strip<WT, M, true, NCH>(frame, out_t, plane_sz, w, h, pw, p0, i0, i1, lane, vec_ok != 0, c0);
}

// ------------------------------------------------------------------------- k_fwd_t_quant
#ifndef ALICE_T_MINBLOCKS
#define ALICE_T_MINBLOCKS 4
#endif
// The compile-time-depth variant (PF = 64) runs a rolled, software-pipelined steady-state loop (31 KB of code); the
// fully unrolled form (122 KB, four times the instruction cache) measured 1.209 ms against 1.106 ms for the whole
// front-end (profiles/r01_ab_frontend.jsonl) and was dropped.
template <int WT, int VEC, int PF>
__global__ void ALICE_LAUNCH_BOUNDS(256, PF != 0 ? 3 : ALICE_T_MINBLOCKS)
k_fwd_t_quant(const int16_t *__restrict__ planes, uint8_t *__restrict__ symbols, unsigned *__restrict__ hist,
              int pw, int ph, int f, int pf, QuantDev q, int32_t *__restrict__ coef_dump) {
    constexpr int NST = WaveletTraits<WT>::NST;
    __shared__ unsigned sh_hist[256];
    const int c = blockIdx.z;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();

    const int halft = PF ? PF / 2 : (pf >> 1);   // PF != 0: compile-time depth, the streaming state machine unrolls away
    const size_t frame_sz = (size_t)ph * pw;
    const int16_t *src = planes + (size_t)c * f * frame_sz;
    uint8_t *dst = symbols + (size_t)c * pf * frame_sz;
    const long long n_items = (long long)(frame_sz / VEC);

    auto emit = [&](size_t off, int jo, const int (&lo)[VEC], const int (&hi)[VEC]) {
        uint32_t pl = 0, phh = 0;
#pragma unroll
        for (int i = 0; i < VEC; i++) {
            uint32_t sl = quant_symbol(lo[i], q), sh = quant_symbol(hi[i], q);
            pl |= sl << (8 * i);
            phh |= sh << (8 * i);
            if (sl) atomicAdd(&sh_hist[sl], 1u);   // bin 0 is filled in afterwards: N - sum of the others
            if (sh) atomicAdd(&sh_hist[sh], 1u);
        }
        if (coef_dump) {  // parity-test stage dump of the 3-D coefficients (null in production)
            int32_t *cd = coef_dump + (size_t)c * pf * frame_sz;
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                cd[(size_t)jo * frame_sz + off + i] = lo[i];
                cd[(size_t)(halft + jo) * frame_sz + off + i] = hi[i];
            }
        }
        uint8_t *pl_dst = dst + (size_t)jo * frame_sz + off;
        uint8_t *ph_dst = dst + (size_t)(halft + jo) * frame_sz + off;
        if (VEC == 4) {
            *reinterpret_cast<uint32_t *>(pl_dst) = pl;
            *reinterpret_cast<uint32_t *>(ph_dst) = phh;
        } else {
            *reinterpret_cast<uint16_t *>(pl_dst) = (uint16_t)pl;
            *reinterpret_cast<uint16_t *>(ph_dst) = (uint16_t)phh;
        }
    };

    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
         item += (long long)gridDim.x * blockDim.x) {
        const size_t off = (size_t)item * VEC;
        FwdLift<WT, false> L[VEC];
        int k = 0;
        // rows 2j and 2j+1 of this thread's VEC columns, as raw i16 pairs (padded frames re-read frame f-1,
        // pipeline.rs:107-112)
        struct RawPair { uint32_t a[VEC / 2], b[VEC / 2]; };
        auto load_pair = [&](int j) {
            const int t0 = min(2 * j, f - 1), t1 = min(2 * j + 1, f - 1);
            RawPair r;
            if (VEC == 4) {
                const uint2 a = __ldg(reinterpret_cast<const uint2 *>(src + (size_t)t0 * frame_sz + off));
                const uint2 b = __ldg(reinterpret_cast<const uint2 *>(src + (size_t)t1 * frame_sz + off));
                r.a[0] = a.x; r.a[VEC / 2 - 1] = a.y;
                r.b[0] = b.x; r.b[VEC / 2 - 1] = b.y;
            } else {
                r.a[0] = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)t0 * frame_sz + off));
                r.b[0] = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)t1 * frame_sz + off));
            }
            return r;
        };
        auto unpack = [&](const RawPair &r, int (&e)[VEC], int (&o)[VEC]) {
#pragma unroll
            for (int i = 0; i < VEC / 2; i++) {
                e[2 * i] = (int16_t)(r.a[i] & 0xffff); e[2 * i + 1] = (int)r.a[i] >> 16;
                o[2 * i] = (int16_t)(r.b[i] & 0xffff); o[2 * i + 1] = (int)r.b[i] >> 16;
            }
        };
        if (PF != 0) {
            // Compile-time depth: a short unrolled prologue (warm-up and the mirrored left edge), a ROLLED
            // steady-state loop with the rows of the next two pairs in flight, and three peeled pairs at the end
            // (the last load is the only one that may need the frame clamp).  The fully unrolled form of this loop
            // was 122 KB of code, four times the 32 KB instruction cache, and the warps of a block sit at different
            // places in it: ncu showed as many no-instruction stalls as scoreboard stalls.
            static_assert(PF == 0 || PF / 2 >= NST + 5, "compile-time depth too short for the pipelined form");
            const int16_t *pn = src + off;                      // even row of the next pair to load
            const size_t fs2 = 2 * frame_sz;
            auto load_next = [&]() {                            // pair rows (pn, pn + frame_sz), no clamp
                RawPair r;
                if (VEC == 4) {
                    const uint2 a = __ldg(reinterpret_cast<const uint2 *>(pn));
                    const uint2 b = __ldg(reinterpret_cast<const uint2 *>(pn + frame_sz));
                    r.a[0] = a.x; r.a[VEC / 2 - 1] = a.y;
                    r.b[0] = b.x; r.b[VEC / 2 - 1] = b.y;
                } else {
                    r.a[0] = __ldg(reinterpret_cast<const uint32_t *>(pn));
                    r.b[0] = __ldg(reinterpret_cast<const uint32_t *>(pn + frame_sz));
                }
                pn += fs2;
                return r;
            };
            RawPair r0 = load_next(), r1 = load_next(), r2 = r1;
            auto steady = [&](int j) {                          // consumes r0 = pair j, emits pair j - NST
                int e[VEC], o[VEC], lo[VEC], hi[VEC];
                unpack(r0, e, o);
#pragma unroll
                for (int i = 0; i < VEC; i++) L[i].push_steady(e[i], o[i], lo[i], hi[i]);
                emit(off, j - NST, lo, hi);
                r0 = r1; r1 = r2;
            };
#pragma unroll
            for (int j = 0; j <= NST; j++) {
                r2 = load_next();                               // pair j + 2 <= NST + 2 < halft - 1
                int e[VEC], o[VEC], lo[VEC], hi[VEC];
                unpack(r0, e, o);
                bool has = false;
#pragma unroll
                for (int i = 0; i < VEC; i++) has = L[i].push(e[i], o[i], j, j, lo[i], hi[i]);
                if (has) emit(off, j - NST, lo, hi);
                r0 = r1; r1 = r2;
            }
#pragma unroll 1
            for (int j = NST + 1; j < halft - 3; j++) {
                r2 = load_next();                               // pair j + 2 <= halft - 2
                steady(j);
            }
            r2 = load_pair(halft - 1);                          // the last pair: its odd row may be frame f - 1 again
            steady(halft - 3);
            steady(halft - 2);
            steady(halft - 1);
            k = halft;
        } else
#pragma unroll
        for (int j = 0; j < halft; j++, k++) {
            int e[VEC], o[VEC];
            unpack(load_pair(j), e, o);
            int lo[VEC], hi[VEC];
            bool has = false;
#pragma unroll
            for (int i = 0; i < VEC; i++) has = L[i].push(e[i], o[i], k, j, lo[i], hi[i]);
            if (has) emit(off, j - NST, lo, hi);
        }
#pragma unroll
        for (int which = 0; which < NST; which++) {
            int lo[VEC], hi[VEC];
            bool has = false;
#pragma unroll
            for (int i = 0; i < VEC; i++) has = L[i].flush(k, which, halft, lo[i], hi[i]);
            if (has) emit(off, halft - NST + which, lo, hi);
        }
    }

    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(&hist[c * 256 + i], sh_hist[i]);
}

// bin 0 = number of symbols - sum of the other bins (the kernels above only count non-zero symbols)
__global__ void k_hist_zero_bin(unsigned *__restrict__ hist, unsigned n_symbols) {
    unsigned *h = hist + blockIdx.x * 256;
    unsigned s = 0;
    for (int i = 1 + threadIdx.x; i < 256; i += 32) s += h[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(kFullMask, s, d);
    if (threadIdx.x == 0) h[0] = n_symbols - s;
}

// ------------------------------------------------------------------------------ launchers
template <int WT>
static void launch_fwd(const uint8_t *d_rgb, int16_t *d_planes, uint8_t *d_symbols, unsigned *d_hist, int w, int h,
                       int f, int pw, int ph, int pf, int step, int32_t *d_coef_dump, cudaStream_t st) {
    constexpr int M = ALICE_FWD_M;
    const int halfx = pw / 2, halfy = ph / 2;
    constexpr int VPAIRS = (32 - 2 * ((WaveletTraits<WT>::NST + M - 1) / M)) * M;
    const int n_strips = (halfx + VPAIRS - 1) / VPAIRS;
    // enough warps to fill the machine: aim for >= 148*24 warps, segments of >= 16 row pairs
    long long base_warps = (long long)n_strips * f;
#ifndef ALICE_XY_TARGET_WARPS
#define ALICE_XY_TARGET_WARPS 96   // warps per SM's worth of segments: finer segments balance the tail (measured)
#endif
#ifndef ALICE_XY_WPB
#define ALICE_XY_WPB 1             // one warp per block: 1.37 -> 1.19 ms per 1080p x 64 chunk vs four (measured)
#endif
    int n_segs = (int)std::min<long long>(std::max<long long>(1, (device_sm_count() * ALICE_XY_TARGET_WARPS + base_warps - 1) / base_warps),
                                          std::max(1, halfy / 16));
    int seg_pairs = (halfy + n_segs - 1) / n_segs;
    n_segs = (halfy + seg_pairs - 1) / seg_pairs;
    const long long n_warps = (long long)n_strips * n_segs * f;
    const int vec_ok = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_rgb) & 3) == 0);
    const int warps_per_block = ALICE_XY_WPB;
    dim3 grid((unsigned)((n_warps + warps_per_block - 1) / warps_per_block));
    auto kxy = k_fwd_xy<WT, M, 3>;
    ALICE_LAUNCH(kxy, grid, dim3(32 * warps_per_block), 0, st, d_rgb, d_planes, w, h, f, pw, ph, n_strips, n_segs,
                 seg_pairs, vec_ok);

    const QuantDev q = make_quant_dev(step);
    const size_t frame_sz = (size_t)pw * ph;
    const int vec = (pw % 4 == 0) ? 4 : 2;
    const long long items = frame_sz / vec;
    const unsigned gx = (unsigned)std::min<long long>((items + 255) / 256, (long long)device_sm_count() * 8);
    const dim3 tgrid(gx, 1, 3), block(256);
    if (vec == 4 && pf == 64) {
        auto kt = k_fwd_t_quant<WT, 4, 64>;
        ALICE_LAUNCH(kt, tgrid, block, 0, st, d_planes, d_symbols, d_hist, pw, ph, f, pf, q, d_coef_dump);
    } else if (vec == 4) {
        auto kt = k_fwd_t_quant<WT, 4, 0>;
        ALICE_LAUNCH(kt, tgrid, block, 0, st, d_planes, d_symbols, d_hist, pw, ph, f, pf, q, d_coef_dump);
    } else {
        auto kt = k_fwd_t_quant<WT, 2, 0>;
        ALICE_LAUNCH(kt, tgrid, block, 0, st, d_planes, d_symbols, d_hist, pw, ph, f, pf, q, d_coef_dump);
    }
    ALICE_LAUNCH(k_hist_zero_bin, dim3(3), dim3(32), 0, st, d_hist, (unsigned)((size_t)pf * frame_sz));
}

void forward_frontend(int wavelet, const uint8_t *d_rgb, int16_t *d_planes, uint8_t *d_symbols, unsigned *d_hist,
                      int w, int h, int f, int pw, int ph, int pf, int step, int32_t *d_coef_dump, cudaStream_t st) {
    switch (wavelet) {
    case WT_CDF53: launch_fwd<WT_CDF53>(d_rgb, d_planes, d_symbols, d_hist, w, h, f, pw, ph, pf, step, d_coef_dump, st); break;
    case WT_CDF97: launch_fwd<WT_CDF97>(d_rgb, d_planes, d_symbols, d_hist, w, h, f, pw, ph, pf, step, d_coef_dump, st); break;
    default:       launch_fwd<WT_HAAR>(d_rgb, d_planes, d_symbols, d_hist, w, h, f, pw, ph, pf, step, d_coef_dump, st); break;
    }
}

}  // namespace alice
