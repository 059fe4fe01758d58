// k_fwd_fused.cu — encode front-end as ONE kernel: RGB u8 -> three u8 symbol planes + histograms, the x/y-subband
// planes never leave the SM (6 B per pixel of HBM traffic instead of 18).
//
// Replaces, for 64-frame chunks with even height and a width that is a multiple of 16 (reference file:line):
//   rgb_bytes_to_ycocg_r            src/color.rs:199-235
//   Wavelet3D::forward              src/wavelet.rs:392-438 (x, then y per frame, then t)
//   Quantizer::quantize_buffer      src/quant.rs:89-128   (dead zone = step)
//   to_symbols                      src/quant.rs:547-563
//   build_histogram                 src/quant.rs:594-600
// (other shapes — odd sizes, padding, depths != 64 — keep the two-kernel path of k_forward.cu).
//
// Tile = (28-pair column strip) x (segment of row pairs) x (one half of the temporal axis + its lifting halo): a block
// of NFR/2 warps, NFR = 32 + 4*NST frames... see FusedGeom.  One block per SM.
//
//   staging   The RGB rows of the next Q row pairs of all NFR frames are fetched into a shared-memory ring by bulk
//             asynchronous copies (cp.async.bulk, 208 bytes per row) that complete on one mbarrier per ring stage;
//             the copies for group g+1 are issued when group g's x/y phase has finished and land during its t phase,
//             so no thread ever waits for global memory with registers tied up.
//   x/y phase Each half-warp owns one frame; a lane owns two horizontal pairs (4 pixels x 3 channels).  Colour
//             transform and x lifting in registers (neighbour values by warp shuffle, lanes 0/15 of a half-warp are
//             halo lanes), y lifting as a streaming state machine per owned column (lifting.cuh).  Finished row
//             pairs go to the t buffer in shared memory as i16 (exact: |coef| <= 7043 after x and y for u8 input).
//   t phase   After Q row pairs: one thread per (row pair, channel, low/high row, column) streams the temporal line
//             of its column out of the t buffer (conflict-free 2-byte reads), lifts, quantises, maps to symbols,
//             stores them and histograms through shared-memory atomics.
//   t split   A block handles output t-pairs [16*twin, 16*twin + 16): the x/y lifting state of all 64 frames of a
//             strip does not fit one SM's registers, so the temporal axis is cut in two and each half recomputes
//             the x/y transform of the 2*NST halo frames its t lifting needs (+6 % / +12.5 % x/y work).
#include <stdlib.h>

#include "kernels.h"
#include "lifting.cuh"
#include "quant.cuh"

namespace alice {

constexpr int kFsVP = 28;             // valid pairs per strip (14 lanes x 2 pairs)
constexpr int kFsCols = 2 * kFsVP;    // coefficients per strip row: 28 low-x + 28 high-x
constexpr int kFsRowBytes = 208;      // bytes fetched per RGB row: 64 pixels = 192 bytes + 16 for the alignment of the start
constexpr int kFsRowPitch = 224;      // ring pitch: two rows = 112 words = 16 banks mod 32, the two frames of a warp do not collide
constexpr int kFsLines = 3 * 2 * kFsCols;   // 336 temporal lines per row pair: channel x (low / high row) x column
constexpr int kFsFrameI16 = 352;      // t-buffer frame stride in i16 (704 bytes = 176 words = 16 banks mod 32)

template <int WT> struct FusedGeom {
    static constexpr int NST = WaveletTraits<WT>::NST;
    static constexpr int NFR = 32 + 2 * NST;          // frames per block: half the temporal axis + NST halo pairs
    static constexpr int NW = NFR / 2;                // warps (two frames per warp)
    static constexpr int NT = NW * 32;                // 544 (5/3, Haar) or 576 (9/7) threads
    static constexpr int Q = NST == 2 ? 5 : 3;        // row pairs per group: Q * 336 lines ~ a whole number of rounds of NT threads
    static constexpr int IN_STAGE = NFR * 2 * kFsRowPitch;
    static constexpr int TB_SLOT = NFR * kFsFrameI16 * 2;
    static constexpr int SMEM = Q * IN_STAGE + Q * TB_SLOT + 3 * 256 * 4 + 8 * (int)sizeof(mbar_t);   // 8 >= Q barriers
};

template <int WT, bool DUMP>
__global__ void ALICE_LAUNCH_BOUNDS(FusedGeom<WT>::NT, 1)
k_fwd_fused(const FwdFusedJob *__restrict__ jobs, int w, int h, int n_strips, int seg_pairs, QuantDev q) {
    typedef FusedGeom<WT> G;
    constexpr int NST = G::NST, NFR = G::NFR, NT = G::NT, Q = G::Q;
    ALICE_DYN_SMEM(smem);
    uint8_t *in_ring = smem;
    int16_t *tbuf = reinterpret_cast<int16_t *>(smem + Q * G::IN_STAGE);
    unsigned *sh_hist = reinterpret_cast<unsigned *>(smem + Q * G::IN_STAGE + Q * G::TB_SLOT);
    mbar_t *bar = reinterpret_cast<mbar_t *>(sh_hist + 3 * 256);

    const FwdFusedJob job = jobs[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int s = (int)(blockIdx.x % (unsigned)n_strips);
    const int twin = (int)((blockIdx.x / (unsigned)n_strips) & 1u);
    const int seg = (int)(blockIdx.x / (2u * (unsigned)n_strips));
    const int halfx = w >> 1, halfy = h >> 1;
    const size_t fs = (size_t)w * h;
    const int i0 = seg * seg_pairs, i1 = min(halfy, i0 + seg_pairs);
    const int js = max(0, i0 - NST), je = min(halfy, i1 + NST);
    const int fr0 = twin ? 64 - NFR : 0;          // first frame of this block's temporal window
    const int fi = 2 * wv + (lane >> 4);          // this half-warp's frame within the window
    const int xl = lane & 15;
    const int rowbytes = w * 3;
    int a = (168 * s - 12) & ~15;                 // 16-byte aligned start of the 208-byte row window
    a = a < 0 ? 0 : (a > rowbytes - kFsRowBytes ? rowbytes - kFsRowBytes : a);
    int off = 168 * s - 12 + 12 * xl - a;         // this lane's 12 bytes inside the window (lanes outside the image read
    off = off < 0 ? 0 : (off > kFsRowBytes - 12 ? kFsRowBytes - 12 : off);   //  in-range garbage: their results are discarded)
    const int p0 = kFsVP * s - 2 + 2 * xl;        // first of this lane's two pairs
    const bool lane_ok = xl >= 1 && xl <= 14 && p0 < halfx;
    const bool edge = s == 0 || kFsVP * s + 30 >= halfx;   // some lane owns pair 0 or pair halfx-1 (the mirrored ones)

    for (int i = tid; i < 3 * 256; i += NT) sh_hist[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < Q; i++) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // Fetch the rows of pairs [jg, jg + nq) of all NFR frames into ring stages 0 .. nq-1: at most one 208-byte bulk copy
    // per thread.  (A bulk copy takes its operands from uniform registers, so a warp issues the copies of its lanes one
    // after the other, ~8 instructions each: spread over all warps they cost every warp a few hundred issue slots per
    // group; issued by one warp they made the whole block wait for it at the next barrier.)
    auto issue_group = [&](int jg, int nq) {
        if (tid == 0)
            for (int i = 0; i < nq; i++) mbar_arrive_expect_tx(&bar[i], (unsigned)(NFR * 2 * kFsRowBytes));
        for (int idx = tid; idx < nq * NFR * 2; idx += NT) {
            const int st = idx / (NFR * 2), r = idx - st * (NFR * 2);
            const int y = 2 * (jg + st) + (r & 1);
            const uint8_t *src = job.rgb + ((size_t)(fr0 + (r >> 1)) * h + y) * rowbytes + a;
            bulk_copy_g2s(in_ring + st * G::IN_STAGE + r * kFsRowPitch, src, kFsRowBytes, &bar[st]);
        }
    };

    // ---- t phase: n_out finished row pairs (first one = row pair jo_first) wait in the t buffer
    auto t_phase = [&](int n_out, int jo_first) {
        const int n_items = n_out * kFsLines;
        for (int item = tid; item < n_items; item += NT) {
            const int qq = item / kFsLines, r = item - qq * kFsLines;
            const int ch = r / (2 * kFsCols), rr = r - ch * (2 * kFsCols);
            const int row = rr / kFsCols, col = rr - row * kFsCols;
            const int xh = col / kFsVP, pair = kFsVP * s + (col - xh * kFsVP);
            if (pair >= halfx) continue;
            const int jo = jo_first + qq;
            const uint32_t pos = (uint32_t)(row ? halfy + jo : jo) * (uint32_t)w + (uint32_t)(xh ? halfx + pair : pair);
            const int16_t *src = tbuf + qq * (NFR * kFsFrameI16) + r;
            // symbol offsets fit 32 bits (3 * 64 * w * h < 2^32 is checked by the launcher)
            const uint32_t fs32 = (uint32_t)fs;
            uint32_t ol = ((uint32_t)ch * 64u + 16u * (uint32_t)twin) * fs32 + pos;   // low-t symbol of the next pair to come out
            const uint32_t hi_delta = 32u * fs32;                                     // its high-t symbol
            unsigned *hist = sh_hist + ch * 256;
            auto emit = [&](int lo, int hi) {
                const uint32_t sl = quant_symbol(lo, q), sh = quant_symbol(hi, q);
                job.symbols[ol] = (uint8_t)sl;
                job.symbols[ol + hi_delta] = (uint8_t)sh;
                if (sl) atomicAdd(&hist[sl], 1u);      // bin 0 is filled in afterwards: N - sum of the others
                if (sh) atomicAdd(&hist[sh], 1u);
                if (DUMP) { job.coef_dump[ol] = lo; job.coef_dump[ol + hi_delta] = hi; }
                ol += fs32;
            };
            FwdLift<WT, false> T;
            int lo, hi;
            if (twin == 0) {
                // pairs 0 .. 15+NST are pushed, pairs 0 .. 15 come out (the left edge is the true, mirrored one)
#pragma unroll
                for (int jt = 0; jt <= NST; jt++) {
                    const int e = src[(2 * jt) * kFsFrameI16], o = src[(2 * jt + 1) * kFsFrameI16];
                    if (T.push(e, o, jt, jt, lo, hi)) emit(lo, hi);
                }
#pragma unroll 4
                for (int jt = NST + 1; jt < 16 + NST; jt++) {
                    const int e = src[(2 * jt) * kFsFrameI16], o = src[(2 * jt + 1) * kFsFrameI16];
                    T.push_steady(e, o, lo, hi);
                    emit(lo, hi);
                }
            } else {
                // pairs 16-NST .. 31 are pushed; the first NST outputs (pairs 16-NST .. 15) are warm-up and dropped
#pragma unroll
                for (int jt = 0; jt < 2 * NST; jt++) {
                    const int e = src[(2 * jt) * kFsFrameI16], o = src[(2 * jt + 1) * kFsFrameI16];
                    if (jt <= NST) T.push(e, o, jt, 16 - NST + jt, lo, hi);
                    else T.push_steady(e, o, lo, hi);
                }
#pragma unroll 4
                for (int jt = 2 * NST; jt < 16 + NST; jt++) {
                    const int e = src[(2 * jt) * kFsFrameI16], o = src[(2 * jt + 1) * kFsFrameI16];
                    T.push_steady(e, o, lo, hi);
                    emit(lo, hi);
                }
#pragma unroll
                for (int which = 0; which < NST; which++)
                    if (T.flush(16 + NST, which, 32, lo, hi)) emit(lo, hi);
            }
        }
    };

    // ---- x/y phase of one row pair: EDGE = the strip touches the left or right image border
    FwdLift<WT, false> L[3][4];   // per channel: columns 0,1 = low-x, 2,3 = high-x
    // Colour transform (color.rs:221-232; values fit i16, so i32 arithmetic is identical) + x lifting of this lane's four
    // pixels of one row, ONE CHANNEL AT A TIME so that only four values are live next to the lifting state: Co = R - B
    // first, then Cg and Y together (Y = t + (Cg >> 1) is held while Cg goes through).  v[0,1] = low-x, v[2,3] = high-x.
    auto xlift = [&](int (&e)[2], int (&o)[2], int (&v)[4]) {
        if (edge) fwd_lanes<WT, false, 2, true>(e, o, p0, halfx);
        else fwd_lanes<WT, false, 2, false>(e, o, p0, halfx);
        v[0] = e[0]; v[1] = e[1]; v[2] = o[0]; v[3] = o[1];
    };
    auto rgb_of = [&](const uint32_t (&raw)[3], int i, int &r, int &g, int &b) {
        r = (raw[(3 * i) >> 2] >> (8 * ((3 * i) & 3))) & 0xff;
        g = (raw[(3 * i + 1) >> 2] >> (8 * ((3 * i + 1) & 3))) & 0xff;
        b = (raw[(3 * i + 2) >> 2] >> (8 * ((3 * i + 2) & 3))) & 0xff;
    };
    auto row_co = [&](const uint32_t (&raw)[3], int (&v)[4]) {
        int e[2], o[2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int r, g, b;
            rgb_of(raw, i, r, g, b);
            if (i & 1) o[i >> 1] = r - b; else e[i >> 1] = r - b;
        }
        xlift(e, o, v);
    };
    auto row_cg_y = [&](const uint32_t (&raw)[3], int (&vcg)[4], int (&ye)[2], int (&yo)[2]) {
        int e[2], o[2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int r, g, b;
            rgb_of(raw, i, r, g, b);
            const int co = r - b;
            const int tt = b + (co >> 1);
            const int cg = g - tt;
            const int yy = tt + (cg >> 1);
            if (i & 1) { o[i >> 1] = cg; yo[i >> 1] = yy; } else { e[i >> 1] = cg; ye[i >> 1] = yy; }
        }
        xlift(e, o, vcg);
    };
    // one channel of a finished row pair -> t buffer (low row and high row of the pair, low-x and high-x halves)
    auto store_ch = [&](int slot, int c, const int (&lo)[4], const int (&hi)[4]) {
        if (!lane_ok) return;
        int16_t *tb = tbuf + slot * (NFR * kFsFrameI16) + fi * kFsFrameI16 + 2 * (xl - 1);
        uint32_t *l0 = reinterpret_cast<uint32_t *>(tb + (2 * c) * kFsCols);
        uint32_t *l1 = reinterpret_cast<uint32_t *>(tb + (2 * c + 1) * kFsCols);
        l0[0] = (uint32_t)(uint16_t)lo[0] | ((uint32_t)(uint16_t)lo[1] << 16);
        l0[kFsVP / 2] = (uint32_t)(uint16_t)lo[2] | ((uint32_t)(uint16_t)lo[3] << 16);
        l1[0] = (uint32_t)(uint16_t)hi[0] | ((uint32_t)(uint16_t)hi[1] << 16);
        l1[kFsVP / 2] = (uint32_t)(uint16_t)hi[2] | ((uint32_t)(uint16_t)hi[3] << 16);
    };

    int k = 0;            // row pairs pushed so far
    int n_acc = 0;        // finished row pairs waiting in the t buffer
    int jo_first = 0;     // row pair index of t-buffer slot 0
    const int n_groups = (je - js + Q - 1) / Q;
    if (n_groups > 0) issue_group(js, min(Q, je - js));
    // One group of row pairs.  FIRST = the tile's first group: its first NST + 1 steps are the warm-up / top-edge steps of
    // the lifting state machines (general push); every later group runs the branch-free steady form only.  Two copies of
    // the body, so that the 60 registers of lifting state never flow through a join of the two forms.
    auto run_group = [&](int g, auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        const int jg = js + g * Q, nq = min(Q, je - jg);
        for (int st = 0; st < nq; st++, k++) {
            const int j = jg + st;
            mbar_wait(&bar[st], (unsigned)(g & 1));
            const uint8_t *rows = in_ring + st * G::IN_STAGE + fi * (2 * kFsRowPitch) + off;
            // The even row goes through the lifting state machines first; the odd row only enters their state
            // (lifting.cuh: push_even / set_odd), so the two rows' values are never live together.
            const int jo = j - NST;
            const bool out_ok = k >= NST && jo >= i0 && jo < i1;   // uniform: this step's output row pair is wanted
            if (out_ok && n_acc == 0) jo_first = jo;
            auto push_ch = [&](int c, const int (&v)[4]) {
                int lo[4], hi[4];
                if (FIRST && k <= NST) {
#pragma unroll
                    for (int i = 0; i < 4; i++) L[c][i].push_even(v[i], k, j, lo[i], hi[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) L[c][i].push_even_steady(v[i], lo[i], hi[i]);
                }
                if (out_ok) store_ch(n_acc, c, lo, hi);
            };
            {
                const uint32_t *p = reinterpret_cast<const uint32_t *>(rows);
                const uint32_t raw[3] = {p[0], p[1], p[2]};
                int v[4], ye[2], yo[2];
                row_co(raw, v);
                push_ch(1, v);
                row_cg_y(raw, v, ye, yo);
                push_ch(2, v);
                xlift(ye, yo, v);
                push_ch(0, v);
            }
            if (out_ok) n_acc++;
            {
                const uint32_t *p = reinterpret_cast<const uint32_t *>(rows + kFsRowPitch);
                const uint32_t raw[3] = {p[0], p[1], p[2]};
                int v[4], ye[2], yo[2];
                row_co(raw, v);
#pragma unroll
                for (int i = 0; i < 4; i++) L[1][i].set_odd(v[i]);
                row_cg_y(raw, v, ye, yo);
#pragma unroll
                for (int i = 0; i < 4; i++) L[2][i].set_odd(v[i]);
                xlift(ye, yo, v);
#pragma unroll
                for (int i = 0; i < 4; i++) L[0][i].set_odd(v[i]);
            }
        }
        __syncthreads();   // every read of the ring is done, the t buffer is complete
        if (g + 1 < n_groups) issue_group(jg + Q, min(Q, je - (jg + Q)));
        if (n_acc) t_phase(n_acc, jo_first);
        __syncthreads();   // the t buffer is free again
        n_acc = 0;
    };
    static_assert(Q > NST, "the warm-up steps must fall into the first group");
    if (n_groups > 0) run_group(0, BoolTag<true>());
    for (int g = 1; g < n_groups; g++) run_group(g, BoolTag<false>());
    if (je == halfy && k > 0) {   // bottom of the image: the last NST row pairs come out of the flush
#pragma unroll
        for (int which = 0; which < NST; which++) {
            const int jo = halfy - NST + which;
            const bool out_ok = (NST == 1 || which == 1 || k >= 2) && jo >= i0 && jo < i1;   // FwdLift::flush: no output from (which 0, k 1)
            if (out_ok && n_acc == 0) jo_first = jo;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                int lo[4], hi[4];
#pragma unroll
                for (int i = 0; i < 4; i++) L[c][i].flush(k, which, halfy, lo[i], hi[i]);
                if (out_ok) store_ch(n_acc, c, lo, hi);
            }
            if (out_ok) n_acc++;
        }
        __syncthreads();
        if (n_acc) t_phase(n_acc, jo_first);
    }
    __syncthreads();
    for (int i = tid; i < 3 * 256; i += NT)
        if (sh_hist[i]) atomicAdd(&job.hist[i], sh_hist[i]);
}

// bin 0 = number of symbols - sum of the other bins (the kernel above only counts non-zero symbols)
__global__ void k_hist_zero_bin_batch(unsigned *__restrict__ hist, unsigned n_symbols) {
    unsigned *hh = hist + (size_t)blockIdx.x * 256;
    unsigned s = 0;
    for (int i = 1 + threadIdx.x; i < 256; i += 32) s += hh[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(kFullMask, s, d);
    if (threadIdx.x == 0) hh[0] = n_symbols - s;
}

bool forward_fused_eligible(const uint8_t *d_rgb, int w, int h, int f) {
    if ((unsigned long long)w * (unsigned long long)h * 192ull >= (1ull << 32)) return false;   // 32-bit symbol offsets
    return f == 64 && (h & 1) == 0 && h >= 2 && (w & 15) == 0 && w >= 80 && (reinterpret_cast<uintptr_t>(d_rgb) & 15) == 0;
}

template <int WT>
static void launch_fused(const FwdFusedJob *d_jobs, int n_jobs, bool dump, int w, int h, int step, int n_sms, cudaStream_t st) {
    typedef FusedGeom<WT> G;
    const int halfx = w / 2, halfy = h / 2;
    const int n_strips = (halfx + kFsVP - 1) / kFsVP;
    // segments: as few as possible (every segment start recomputes NST row pairs), but enough blocks for every SM and a
    // last wave that is reasonably full
    int best_segs = 1;
    double best_cost = 1e30;
    for (int n_segs = 1; n_segs <= 16 && halfy / n_segs >= 8; n_segs++) {
        const int sp = (halfy + n_segs - 1) / n_segs;
        const long long blocks = (long long)n_jobs * 2 * n_strips * ((halfy + sp - 1) / sp);
        const long long waves = (blocks + n_sms - 1) / n_sms;
        const double cost = (double)waves * (sp + 2 * G::NST + 2);   // time ~ waves x rows per block (+ warm-up and pipeline fill)
        if (cost < best_cost) { best_cost = cost; best_segs = n_segs; }
    }
    const int seg_pairs = (halfy + best_segs - 1) / best_segs;
    const int n_segs = (halfy + seg_pairs - 1) / seg_pairs;
    const QuantDev q = make_quant_dev(step);
    const dim3 grid((unsigned)(2 * n_strips * n_segs), (unsigned)n_jobs);
#ifndef ALICE_EMUL
    static unsigned long long done_plain = 0, done_dump = 0;
    auto set_attr = [&](auto kernel, unsigned long long &mask) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (mask & (1ull << (dev & 63))) return;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
        mask |= 1ull << (dev & 63);
    };
#endif
    if (dump) {
        auto kf = k_fwd_fused<WT, true>;
#ifndef ALICE_EMUL
        set_attr(kf, done_dump);
#endif
        ALICE_LAUNCH(kf, grid, dim3(G::NT), G::SMEM, st, d_jobs, w, h, n_strips, seg_pairs, q);
    } else {
        auto kf = k_fwd_fused<WT, false>;
#ifndef ALICE_EMUL
        set_attr(kf, done_plain);
#endif
        ALICE_LAUNCH(kf, grid, dim3(G::NT), G::SMEM, st, d_jobs, w, h, n_strips, seg_pairs, q);
    }
}

void forward_frontend_fused(int wavelet, const FwdFusedJob *d_jobs, int n_jobs, bool dump, unsigned *d_hist_base, int w, int h,
                            int step, int n_sms, cudaStream_t st) {
    if (n_jobs <= 0) return;
    switch (wavelet) {
    case WT_CDF53: launch_fused<WT_CDF53>(d_jobs, n_jobs, dump, w, h, step, n_sms, st); break;
    case WT_CDF97: launch_fused<WT_CDF97>(d_jobs, n_jobs, dump, w, h, step, n_sms, st); break;
    default:       launch_fused<WT_HAAR>(d_jobs, n_jobs, dump, w, h, step, n_sms, st); break;
    }
    // histograms of the batch are contiguous: [job][3][256]
    ALICE_LAUNCH(k_hist_zero_bin_batch, dim3(3 * n_jobs), dim3(32), 0, st, d_hist_base, (unsigned)((size_t)64 * w * h));
}

}  // namespace alice
