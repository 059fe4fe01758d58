// k_rdo.cu — AnalyticalRDO sub-band statistics and the per-octant FastQuantizer pass.
//
// Replaces (reference file:line):
//   AnalyticalRDO::estimate_variance          src/quant.rs:415-435   (f64, SEQUENTIAL sum about the mean)
//   AnalyticalRDO::compute_all_quantizers     src/quant.rs:472-490   (8 octants of a 3-D decomposition, lib.rs:115-160)
//   FastQuantizer::quantize_buffer            src/quant.rs:243-299   (per octant, one launch for the volume)
//
// The variance is `sum_k fl(fl(x_k - mean)^2)` accumulated left to right in f64, so its value depends on the
// order of the additions; the quantiser step is round(sqrt(.)) of it, so a differently ordered sum can flip a
// step at a rounding tie.  This file reproduces the sequential result EXACTLY, in parallel:
//
//   While the running sum s stays inside one binade [2^e, 2^(e+1)) its ulp u = 2^(e-52) is fixed and
//   fl(s + t) = s + u * R(t/u), where R rounds to the nearest integer and only an exact tie looks at s (ties go to
//   the even multiple of u).  With S = s/u an integer in [2^52, 2^53), one addition is therefore a map
//   S -> S + c[S & 1] described by two integers (c[0], c[1]), and such maps compose associatively:
//       (g o f).c[p] = f.c[p] + g.c[(p + f.c[p]) & 1].
//   One warp composes the maps of 1024 consecutive terms (32 per lane, then an ordered shuffle reduction); a
//   single thread then walks the per-block maps until S would reach 2^53.  That block — the one in which the sum
//   changes binade — is redone with genuine f64 additions by one thread, and the scan restarts from there with
//   the new ulp.  The sum doubles about log2(n / 1024) times, so a 16.6 M-coefficient octant needs ~15 passes.
//
// Terms are computed with explicit round-to-nearest sub/mul (no FMA contraction), as the reference's scalar code.
#include <math.h>
#include <string.h>

#include "kernels.h"

namespace alice {

constexpr int kRdoBlock = 1024;          // terms per composed map (one warp: 32 lanes x 32 terms)
constexpr int kRdoWarps = 8;             // warps per CTA of k_rdo_block_maps

// x_i of a sub-box of a w x h x d volume, visited in row-major order (t, then y, then x) — the order in which a
// caller of AnalyticalRDO::compute_quantizer would gather a sub-band into a slice.
struct RdoView {
    const int32_t *base;                 // first element of the sub-box
    unsigned long long n;                // elements in the sub-box (0: empty)
    unsigned sw, sh;                     // sub-box width and height (depth = n / (sw * sh))
    unsigned long long row, plane;       // strides of the volume: W and W*H
};

ALICE_D unsigned long long rdo_offset(const RdoView &v, unsigned long long i) {
    const unsigned long long x = i % v.sw, r = i / v.sw;
    const unsigned long long y = r % v.sh, t = r / v.sh;
    return t * v.plane + y * v.row + x;
}
// iterate `count` consecutive elements starting at linear index i0, calling f(x) in order
template <class F> ALICE_D void rdo_for_each(const RdoView &v, unsigned long long i0, int count, F f) {
    unsigned long long x = i0 % v.sw, r = i0 / v.sw;
    unsigned long long y = r % v.sh, t = r / v.sh;
    const int32_t *p = v.base + t * v.plane + y * v.row;
    for (int k = 0; k < count; k++) {
        f(p[x]);
        if (++x == v.sw) {
            x = 0;
            if (++y == v.sh) { y = 0; t++; }
            p = v.base + t * v.plane + y * v.row;
        }
    }
}

ALICE_D double rdo_term(int32_t x, double mean) {
#ifdef ALICE_EMUL
    volatile double diff = (double)x - mean;
    volatile double sq = diff * diff;
    return sq;
#else
    const double diff = __dsub_rn((double)x, mean);
    return __dmul_rn(diff, diff);
#endif
}
ALICE_D double rdo_add(double a, double b) {
#ifdef ALICE_EMUL
    volatile double s = a + b;
    return s;
#else
    return __dadd_rn(a, b);
#endif
}
ALICE_D unsigned long long dbl_bits(double v) {
    unsigned long long b;
    memcpy(&b, &v, 8);
    return b;
}

struct RdoState {                        // one per view, device resident
    double s;                            // running sum after `pos` terms
    unsigned long long pos;
    double mean;
    long long isum;                      // exact integer sum of the view
    unsigned long long first_nz;         // scratch of k_rdo_first_nonzero
    unsigned long long pad;
};
struct RdoMap {                          // the composed map of one block: S -> S + c[S & 1]; over = leaves the binade for sure
    unsigned long long c0, c1;
    unsigned over, pad;
};

// ---- exact integer sum (quant.rs:422) -----------------------------------------------------------------
__global__ void k_rdo_sum(const RdoView *__restrict__ views, RdoState *__restrict__ st) {
    const RdoView v = views[blockIdx.y];
    long long acc = 0;
    const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += nth)
        acc += v.base[rdo_offset(v, i)];
    if (acc) atomicAdd(reinterpret_cast<unsigned long long *>(&st[blockIdx.y].isum), (unsigned long long)acc);
}
__global__ void k_rdo_mean(const RdoView *__restrict__ views, RdoState *__restrict__ st, int n_views) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_views) return;
    const double n = (double)views[i].n;
    const double inv_n = 1.0 / n;                       // quant.rs:421
    st[i].mean = (double)st[i].isum * inv_n;            // quant.rs:423
    st[i].s = 0.0;
    st[i].pos = 0;
}

// ---- genuine sequential additions for up to kRdoBlock terms from pos (one warp per view) --------------
// The terms are pure functions of x and the mean: all lanes compute them into shared memory, lane 0 adds them in order.
__global__ void ALICE_LAUNCH_BOUNDS(32, 1) k_rdo_seq(const RdoView *__restrict__ views, RdoState *__restrict__ st) {
    __shared__ double terms[kRdoBlock];
    const RdoView v = views[blockIdx.x];
    const RdoState s = st[blockIdx.x];
    if (s.pos >= v.n) return;                              // block-uniform
    const unsigned long long left = v.n - s.pos;
    const int count = left < (unsigned long long)kRdoBlock ? (int)left : kRdoBlock;
    const int lane = threadIdx.x;
    const double mean = s.mean;
    {
        const int i0 = lane * 32;                          // lane L owns terms [32L, 32L + 32)
        if (i0 < count) {
            int k = i0;
            rdo_for_each(v, s.pos + (unsigned long long)i0, min(32, count - i0), [&](int32_t x) { terms[k++] = rdo_term(x, mean); });
        }
    }
    __syncwarp();
    if (lane == 0) {
        double acc = s.s;
        for (int k = 0; k < count; k++) acc = rdo_add(acc, terms[k]);
        st[blockIdx.x].s = acc;
        st[blockIdx.x].pos = s.pos + (unsigned long long)count;
        st[blockIdx.x].first_nz = ~0ull;
    }
}

// ---- while the sum is still +0: skip the leading run of zero terms -----------------------------------
__global__ void k_rdo_first_nonzero(const RdoView *__restrict__ views, RdoState *__restrict__ st) {
    const RdoView v = views[blockIdx.y];
    const RdoState s = st[blockIdx.y];
    if (s.pos >= v.n || s.s != 0.0) return;
    unsigned long long best = ~0ull;
    const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = s.pos + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += nth) {
        if (rdo_term(v.base[rdo_offset(v, i)], s.mean) != 0.0) { best = i; break; }
    }
    if (best != ~0ull) atomicMin(&st[blockIdx.y].first_nz, best);
}
__global__ void k_rdo_skip_zeros(const RdoView *__restrict__ views, RdoState *__restrict__ st, int n_views) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_views) return;
    if (st[i].pos >= views[i].n || st[i].s != 0.0) return;
    st[i].pos = st[i].first_nz == ~0ull ? views[i].n : st[i].first_nz;   // adding +0 terms leaves +0
}

// ---- one composed map per block of kRdoBlock terms, for the binade the sum is in ---------------------
ALICE_D void rdo_compose(unsigned long long &c0, unsigned long long &c1, unsigned long long g0, unsigned long long g1) {
    const unsigned long long n0 = c0 + ((c0 & 1) ? g1 : g0);          // S even: parity after f is parity(c0)
    const unsigned long long n1 = c1 + (((c1 + 1) & 1) ? g1 : g0);    // S odd:  parity after f is parity(1 + c1)
    c0 = n0;
    c1 = n1;
}
__global__ void ALICE_LAUNCH_BOUNDS(32 * kRdoWarps, 1)
k_rdo_block_maps(const RdoView *__restrict__ views, const RdoState *__restrict__ st, RdoMap *__restrict__ maps,
                 unsigned long long maps_per_view) {
    const RdoView v = views[blockIdx.y];
    const RdoState s = st[blockIdx.y];
    if (s.pos >= v.n || s.s == 0.0) return;
    const unsigned long long blk = (unsigned long long)blockIdx.x * kRdoWarps + (threadIdx.x >> 5);
    const unsigned long long n_blk = (v.n - s.pos + kRdoBlock - 1) / kRdoBlock;
    if (blk >= n_blk) return;                              // warp-uniform
    const int lane = threadIdx.x & 31;
    // binade of the running sum: s = m * 2^(e_u) with m in [2^52, 2^53), ulp u = 2^(e_u)
    const unsigned long long sb = dbl_bits(s.s);
    const int e_u = (int)((sb >> 52) & 0x7ff) - 1075;      // s is positive and normal here (>= 2^-1022 in practice)
    const double top = ldexp(1.0, e_u + 53);               // 2^(e+1): a term this large leaves the binade on its own
    unsigned long long c0 = 0, c1 = 0;
    unsigned over = 0;
    const unsigned long long i0 = s.pos + blk * kRdoBlock + (unsigned long long)lane * 32;
    if (i0 < v.n) {
        const unsigned long long left = v.n - i0;
        const int count = left < 32 ? (int)left : 32;
        const double mean = s.mean;
        rdo_for_each(v, i0, count, [&](int32_t x) {
            const double t = rdo_term(x, mean);
            if (t == 0.0) return;
            if (!(t < top)) { over = 1; return; }
            const unsigned long long tb = dbl_bits(t);
            const int te = (int)((tb >> 52) & 0x7ff);
            unsigned long long m = tb & 0xfffffffffffffull;
            int q;
            if (te == 0) q = -1074;                        // subnormal term: no implicit bit
            else { m |= 1ull << 52; q = te - 1075; }
            const int d = e_u - q;                         // t = m * 2^q = (m / 2^d) * u, d >= 0 because t < 2^(e+1)
            unsigned long long g0, g1;
            if (d <= 0) { g0 = g1 = m << (-d); }           // d == 0 only (t < top), exact multiple of u
            else if (d >= 55) return;                      // t < u / 2: absorbed
            else {
                const unsigned long long a = m >> d, rem = m & ((1ull << d) - 1), half = 1ull << (d - 1);
                if (rem > half) g0 = g1 = a + 1;
                else if (rem < half) g0 = g1 = a;
                else {                                     // exact tie: the even multiple of u wins
                    g0 = a + (a & 1);                      // S even: S + a has the parity of a
                    g1 = a + ((a + 1) & 1);                // S odd
                }
            }
            rdo_compose(c0, c1, g0, g1);
        });
    }
    // ordered reduction over the lanes: lane L holds terms [32L, 32L+32), so (lane L+d) comes after (lane L)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long g0 = __shfl_down_sync(kFullMask, c0, d);
        const unsigned long long g1 = __shfl_down_sync(kFullMask, c1, d);
        const unsigned go = __shfl_down_sync(kFullMask, over, d);
        if ((lane & (2 * d - 1)) == 0) {                   // lane + d < 32 always holds for these lanes
            rdo_compose(c0, c1, g0, g1);
            over |= go;
        }
    }
    if (lane == 0) {
        RdoMap mp;
        mp.c0 = c0; mp.c1 = c1; mp.over = over; mp.pad = 0;
        maps[(unsigned long long)blockIdx.y * maps_per_view + blk] = mp;
    }
}

// ---- walk the block maps until the sum would leave its binade (one thread per view) ------------------
__global__ void k_rdo_walk(const RdoView *__restrict__ views, RdoState *__restrict__ st, const RdoMap *__restrict__ maps,
                           unsigned long long maps_per_view) {
    if (threadIdx.x != 0) return;
    const RdoView v = views[blockIdx.x];
    RdoState s = st[blockIdx.x];
    if (s.pos >= v.n || s.s == 0.0) return;
    const unsigned long long n_blk = (v.n - s.pos + kRdoBlock - 1) / kRdoBlock;
    const unsigned long long sb = dbl_bits(s.s);
    const int e_u = (int)((sb >> 52) & 0x7ff) - 1075;
    unsigned long long S = (sb & 0xfffffffffffffull) | (1ull << 52);
    const RdoMap *mp = maps + (unsigned long long)blockIdx.x * maps_per_view;
    unsigned long long b = 0;
    for (; b < n_blk; b++) {
        const RdoMap m = mp[b];
        if (m.over) break;
        const unsigned long long c = (S & 1) ? m.c1 : m.c0;
        if (c >= (1ull << 53) || S + c >= (1ull << 53)) break;     // the block changes the binade: redo it for real
        S += c;
    }
    unsigned long long pos = s.pos + b * kRdoBlock;
    if (pos > v.n) pos = v.n;
    st[blockIdx.x].s = ldexp((double)S, e_u);              // exact: S < 2^53
    st[blockIdx.x].pos = pos;
}

// ---- per-octant FastQuantizer over the whole volume (quant.rs:243-264) -------------------------------
struct RdoQuant { int dz[8]; unsigned long long recip[8]; unsigned shift[8]; };
__global__ void k_rdo_quantize(const int32_t *__restrict__ in, int32_t *__restrict__ out, unsigned w, unsigned h,
                               unsigned d, RdoQuant q) {
    const unsigned long long n = (unsigned long long)w * h * d;
    const unsigned hx = w / 2, hy = h / 2, ht = d / 2;
    const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        const unsigned x = (unsigned)(i % w);
        const unsigned long long r = i / w;
        const unsigned y = (unsigned)(r % h), t = (unsigned)(r / h);
        // SubBand3D letters are (x, y, t) (lib.rs:115-132): index = 4*[x high] + 2*[y high] + [t high]
        const int sb = (x >= hx ? 4 : 0) | (y >= hy ? 2 : 0) | (t >= ht ? 1 : 0);
        const int v = in[i];
        const int a = v < 0 ? (int)(0u - (unsigned)v) : v;
        int qv = 0;
        if (!(a < q.dz[sb])) {
            const unsigned adj = (unsigned)a - (unsigned)(q.dz[sb] >> 1);
            const unsigned long long prod = (unsigned long long)adj * q.recip[sb];   // wrapping u64
            const int qa = (int)(unsigned)(prod >> q.shift[sb]);
            qv = v < 0 ? (int)(0u - (unsigned)qa) : qa;
        }
        out[i] = qv;
    }
}

// -------------------------------------------------------------------------------------- host drivers
// variance accumulators (the sequential f64 sums, before the division by n) of n_views sub-boxes.
// h_views: host copies; d_* scratch is allocated here.  Returns cudaSuccess or the failing error.
cudaError_t rdo_seq_sums(const RdoViewHost *h_views, int n_views, double *h_acc, double *h_mean, cudaStream_t st) {
    if (n_views <= 0) return cudaSuccess;
    RdoView hv[8];
    if (n_views > 8) return cudaErrorInvalidValue;
    unsigned long long max_n = 0;
    for (int i = 0; i < n_views; i++) {
        hv[i].base = h_views[i].base; hv[i].n = h_views[i].n; hv[i].sw = h_views[i].sw; hv[i].sh = h_views[i].sh;
        hv[i].row = h_views[i].row; hv[i].plane = h_views[i].plane;
        if (hv[i].n > max_n) max_n = hv[i].n;
    }
    const unsigned long long maps_per_view = (max_n + kRdoBlock - 1) / kRdoBlock + 1;
    // one scratch block: views | states | block maps; one pinned block for the state read-back of every pass
    const size_t maps_off = 1024, maps_bytes = sizeof(RdoMap) * maps_per_view * (size_t)n_views;
    unsigned char *d_scr = (unsigned char *)scratch_device(maps_off + maps_bytes);
    RdoState *hs = (RdoState *)scratch_pinned(sizeof(RdoState) * 8);
    if (!d_scr || !hs) return cudaErrorMemoryAllocation;
    static_assert(sizeof(RdoView) * 8 + sizeof(RdoState) * 8 <= 1024, "scratch header too small");
    RdoView *d_views = (RdoView *)d_scr;
    RdoState *d_state = (RdoState *)(d_scr + sizeof(RdoView) * 8);
    RdoMap *d_maps = (RdoMap *)(d_scr + maps_off);
    cudaError_t e;
    auto fail = [&](cudaError_t err) { return err; };
    if ((e = cudaMemcpyAsync(d_views, hv, sizeof(RdoView) * n_views, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e);
    if ((e = cudaMemsetAsync(d_state, 0, sizeof(RdoState) * 8, st)) != cudaSuccess) return fail(e);
    const int gx = (int)std::min<unsigned long long>((max_n + 255) / 256, (unsigned long long)device_sm_count() * 8);
    if (max_n) {
        ALICE_LAUNCH(k_rdo_sum, dim3(gx ? gx : 1, n_views), dim3(256), 0, st, d_views, d_state);
    }
    ALICE_LAUNCH(k_rdo_mean, dim3(1), dim3(32), 0, st, d_views, d_state, n_views);
    for (int guard = 0;; guard++) {
        ALICE_LAUNCH(k_rdo_seq, dim3(n_views), dim3(32), 0, st, d_views, d_state);
        ALICE_LAUNCH(k_rdo_first_nonzero, dim3(gx ? gx : 1, n_views), dim3(256), 0, st, d_views, d_state);
        ALICE_LAUNCH(k_rdo_skip_zeros, dim3(1), dim3(32), 0, st, d_views, d_state, n_views);
        const unsigned long long grid_maps = (maps_per_view + kRdoWarps - 1) / kRdoWarps;
        ALICE_LAUNCH(k_rdo_block_maps, dim3((unsigned)grid_maps, n_views), dim3(32 * kRdoWarps), 0, st, d_views, d_state,
                     d_maps, maps_per_view);
        ALICE_LAUNCH(k_rdo_walk, dim3(n_views), dim3(32), 0, st, d_views, d_state, d_maps, maps_per_view);
        if ((e = cudaMemcpyAsync(hs, d_state, sizeof(RdoState) * n_views, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(e);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e);
        bool done = true;
        for (int i = 0; i < n_views; i++) done = done && hs[i].pos >= hv[i].n;
        if (done) break;
        if (guard > (1 << 22)) return fail(cudaErrorUnknown);   // every pass consumes at least one block per view
    }
    for (int i = 0; i < n_views; i++) { h_acc[i] = hs[i].s; h_mean[i] = hs[i].mean; }
    return cudaGetLastError();
}

void rdo_quantize_volume(const int32_t *d_in, int32_t *d_out, unsigned w, unsigned h, unsigned d, const int dz[8],
                         const unsigned long long recip[8], const unsigned shift[8], cudaStream_t st) {
    const unsigned long long n = (unsigned long long)w * h * d;
    if (!n) return;
    RdoQuant q;
    for (int i = 0; i < 8; i++) { q.dz[i] = dz[i]; q.recip[i] = recip[i]; q.shift[i] = shift[i]; }
    const int gx = (int)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)device_sm_count() * 16);
    ALICE_LAUNCH(k_rdo_quantize, dim3(gx), dim3(256), 0, st, d_in, d_out, w, h, d, q);
}

}  // namespace alice
