// lifting.cuh — integer lifting arithmetic of the reference, restated for registers.
//
// Reference: src/wavelet.rs:61-248 (Wavelet1D).  One lifting step adds
//     delta = ((a + b) * c + 4096) >> 13          (i64 product, arithmetic shift)
// to every sample of one parity, reading only the other parity (wavelet.rs:180-217):
//     predict(c): s[2i+1] += delta(s[2i]   + (2i+2 < n ? s[2i+2] : s[2i]), c)
//     update(c) : s[2i]   += delta((i > 0 ? s[2i-1] : s[1]) + s[2i+1],    c)
// The inverse applies the steps in reverse order with -c and the SAME +4096 rounding
// (wavelet.rs:157-176), so it is not the algebraic inverse; we reproduce it as is.
//
// Two evaluation forms are provided for lines of EVEN length n = 2*half:
//   * streaming (FwdLift / InvLift): pairs (s[2i], s[2i+1]) are pushed in order and the
//     transformed pair comes out NST pushes later.  State is 3 (5/3, Haar) or 5 (9/7)
//     registers per line for the forward transform, 2 or 4 for the inverse.  Used along
//     y and t, where one thread walks a whole column / temporal line.
//   * lane-parallel (fwd_lanes / inv_lanes): each lane of a warp owns M consecutive pairs
//     of one row; neighbours' boundary values travel by warp shuffle.  Lanes 0 and 31 are
//     halo lanes whose results are discarded.  Used along x.
// WIDE=false uses 32-bit products and is exact whenever |(a+b)*c| < 2^31 (proved for the
// u8-input encode path in DESIGN.md); WIDE=true is the reference's i64 arithmetic for
// arbitrary i32 data (decode path, public Wavelet API).
#pragma once
#include "compat.h"

namespace alice {

enum WaveletKind : int { WT_CDF53 = 0, WT_CDF97 = 1, WT_HAAR = 2 };  // pipeline.rs:34-41

// wavelet.rs:66-127: step lists.  A "stage" is one (predict, update) pair.
template <int WT> struct WaveletTraits { static constexpr int NST = (WT == WT_CDF97) ? 2 : 1; };
template <int WT> ALICE_HD constexpr int coef_p(int stage) {
    return WT == WT_CDF97 ? (stage == 0 ? -6497 : 3616) : -4096;
}
template <int WT> ALICE_HD constexpr int coef_u(int stage) {
    return WT == WT_CDF97 ? (stage == 0 ? -217 : 1817) : (WT == WT_CDF53 ? 1024 : 2048);
}

ALICE_HD int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }  // wrapping i32 add

template <bool WIDE> ALICE_HD int lift_delta(int sum, int c) {
    if (WIDE) return (int)(((long long)sum * (long long)c + 4096) >> 13);
    return (sum * c + 4096) >> 13;
}

// ---------------------------------------------------------------------------------------
// Streaming forward transform of one line.  Usage for global pair indices j = js .. je-1:
//     k = 0; for j: has = L.push(e, o, k++, j, lo, hi)   -> pair j-NST when has
//     at the true end of the line (je == half): NST calls of L.flush(k, which, half, lo, hi)
// Outputs are exact for pair indices >= js + NST when js > 0 (warm-up), or >= 0 when js == 0.
// ---------------------------------------------------------------------------------------
template <int WT, bool WIDE> struct FwdLift {
    static constexpr int NST = WaveletTraits<WT>::NST;
    int e, o, d;    // stage 0: previous pair (even, odd) and d of that pair's left neighbour
    int e2, d2;     // stage 1 (9/7 only): previous s1 and d2 of its left neighbour

    // one (predict, update) stage: previous pair (pe, po), new even ne, left detail dl_prev.
    static ALICE_HD void stage(int pe, int po, int ne, int dl_prev, bool first, int cp, int cu, int &s_out, int &d_out) {
        int dcur = wadd(po, lift_delta<WIDE>(wadd(pe, ne), cp));
        int dl = first ? dcur : dl_prev;
        s_out = wadd(pe, lift_delta<WIDE>(wadd(dl, dcur), cu));
        d_out = dcur;
    }

    // k = number of pairs pushed before this one (uniform), j = global index of this pair.
    ALICE_HD bool push(int en, int on, int k, int j, int &lo, int &hi) {
        if (k == 0) { e = en; o = on; d = 0; e2 = 0; d2 = 0; return false; }
        const int d_old = d;  // d1 of pair j-2 = stage 1's "odd" of its previous pair
        int s1, d1;
        stage(e, o, en, d, j - 1 == 0, coef_p<WT>(0), coef_u<WT>(0), s1, d1);
        d = d1; e = en; o = on;
        if (NST == 1) { lo = s1; hi = d1; return true; }
        if (k == 1) { e2 = s1; return false; }
        int s2, dd;
        stage(e2, d_old, s1, d2, j - 2 == 0, coef_p<WT>(1), coef_u<WT>(1), s2, dd);
        d2 = dd; e2 = s1;
        lo = s2; hi = dd;
        return true;
    }
    // push() for k > NST and j > NST: no warm-up and no left-edge mirror, the pair j-NST always comes out
    ALICE_HD void push_steady(int en, int on, int &lo, int &hi) {
        const int d_old = d;
        int s1, d1;
        stage(e, o, en, d, false, coef_p<WT>(0), coef_u<WT>(0), s1, d1);
        d = d1; e = en; o = on;
        if (NST == 1) { lo = s1; hi = d1; return; }
        int s2, dd;
        stage(e2, d_old, s1, d2, false, coef_p<WT>(1), coef_u<WT>(1), s2, dd);
        d2 = dd; e2 = s1;
        lo = s2; hi = dd;
    }
    // The odd sample of a pushed pair only enters the state (the outputs of a push depend on the new EVEN sample and on
    // older pairs), so a push can be split: push_even / push_even_steady with s[2j], then set_odd with s[2j+1].  A
    // caller that produces the two samples one after the other (two image rows) never holds both.
    ALICE_HD bool push_even(int en, int k, int j, int &lo, int &hi) { return push(en, o, k, j, lo, hi); }
    ALICE_HD void push_even_steady(int en, int &lo, int &hi) { push_steady(en, o, lo, hi); }
    ALICE_HD void set_odd(int on) { o = on; }
    // which = 0 .. NST-1; half = number of pairs in the whole line; k = pairs pushed so far (>= 1).
    ALICE_HD bool flush(int k, int which, int half, int &lo, int &hi) {
        if (NST == 1) {
            stage(e, o, e, d, half - 1 == 0, coef_p<WT>(0), coef_u<WT>(0), lo, hi);
            return true;
        }
        if (which == 0) {
            const int d_old = d;
            int s1, d1;
            stage(e, o, e, d, half - 1 == 0, coef_p<WT>(0), coef_u<WT>(0), s1, d1);
            d = d1;
            if (k == 1) { e2 = s1; return false; }
            int s2, dd;
            stage(e2, d_old, s1, d2, half - 2 == 0, coef_p<WT>(1), coef_u<WT>(1), s2, dd);
            d2 = dd; e2 = s1;
            lo = s2; hi = dd;
            return true;
        }
        stage(e2, d, e2, d2, half - 1 == 0, coef_p<WT>(1), coef_u<WT>(1), lo, hi);
        return true;
    }
};

// ---------------------------------------------------------------------------------------
// Streaming inverse transform of one line: push (low[i], high[i]); the reconstructed pair
// (s[2i], s[2i+1]) comes out NST pushes later.  Steps run in reverse order with negated
// coefficients: per stage first the update on evens, then the predict on odds.
// ---------------------------------------------------------------------------------------
template <int WT, bool WIDE> struct InvLift {
    static constexpr int NST = WaveletTraits<WT>::NST;
    int ep, op;     // first applied stage (the LAST forward stage): previous e' and odd
    int ep2, op2;   // second applied stage (9/7 only)

    // first half of a stage: e' of the new pair
    static ALICE_HD int upd(int en, int on, int o_prev, bool first, int cu) {
        int ol = first ? on : o_prev;
        return wadd(en, lift_delta<WIDE>(wadd(ol, on), -cu));
    }
    // second half: odd of the previous pair, given e' of previous and of the new pair
    static ALICE_HD int prd(int o_prev, int e_prev, int e_new, int cp) {
        return wadd(o_prev, lift_delta<WIDE>(wadd(e_prev, e_new), -cp));
    }

    ALICE_HD bool push(int en, int on, int k, int j, int &ev, int &od) {
        constexpr int SA = NST - 1;  // forward stage undone first
        int e1 = upd(en, on, k == 0 ? on : op, j == 0, coef_u<WT>(SA));
        if (k == 0) { ep = e1; op = on; ep2 = 0; op2 = 0; return false; }
        int a_e = ep, a_o = prd(op, ep, e1, coef_p<WT>(SA));  // pair j-1 after stage A
        ep = e1; op = on;
        if (NST == 1) { ev = a_e; od = a_o; return true; }
        int e2n = upd(a_e, a_o, op2, j - 1 == 0, coef_u<WT>(0));
        if (k == 1) { ep2 = e2n; op2 = a_o; return false; }
        ev = ep2; od = prd(op2, ep2, e2n, coef_p<WT>(0));     // pair j-2
        ep2 = e2n; op2 = a_o;
        return true;
    }
    // push() for k > NST and j > NST: no warm-up and no left-edge mirror, the pair j-NST always comes out
    ALICE_HD void push_steady(int en, int on, int &ev, int &od) {
        constexpr int SA = NST - 1;
        const int e1 = upd(en, on, op, false, coef_u<WT>(SA));
        const int a_e = ep, a_o = prd(op, ep, e1, coef_p<WT>(SA));
        ep = e1; op = on;
        if (NST == 1) { ev = a_e; od = a_o; return; }
        const int e2n = upd(a_e, a_o, op2, false, coef_u<WT>(0));
        ev = ep2; od = prd(op2, ep2, e2n, coef_p<WT>(0));
        ep2 = e2n; op2 = a_o;
    }
    ALICE_HD bool flush(int k, int which, int half, int &ev, int &od) {
        constexpr int SA = NST - 1;
        if (NST == 1) { ev = ep; od = prd(op, ep, ep, coef_p<WT>(SA)); return true; }
        if (which == 0) {
            int a_e = ep, a_o = prd(op, ep, ep, coef_p<WT>(SA));  // pair half-1 after stage A
            int e2n = upd(a_e, a_o, op2, half - 1 == 0, coef_u<WT>(0));
            if (k == 1) { ep2 = e2n; op2 = a_o; return false; }
            ev = ep2; od = prd(op2, ep2, e2n, coef_p<WT>(0));    // pair half-2
            ep2 = e2n; op2 = a_o;
            return true;
        }
        ev = ep2; od = prd(op2, ep2, ep2, coef_p<WT>(0));        // pair half-1
        return true;
    }
};

#if defined(__CUDACC__) || defined(ALICE_EMUL)
// ---------------------------------------------------------------------------------------
// Lane-parallel forward transform along x: this lane owns pairs p0 .. p0+M-1 of a row with
// `half` pairs.  On return e[] holds the low-pass and o[] the high-pass values of those
// pairs.  Exact for lanes 1..30 (M >= 2); lanes 0 / 31 are halo lanes.
// ---------------------------------------------------------------------------------------
template <int WT, bool WIDE, int M, bool EDGE = true>
ALICE_D void fwd_lanes(int (&e)[M], int (&o)[M], int p0, int half) {
#pragma unroll
    for (int st = 0; st < WaveletTraits<WT>::NST; st++) {
        const int cp = coef_p<WT>(st), cu = coef_u<WT>(st);
        int e_nb = __shfl_down_sync(kFullMask, e[0], 1);
#pragma unroll
        for (int k = 0; k < M; k++) {
            int er = (k + 1 < M) ? e[k + 1] : e_nb;
            if (EDGE && p0 + k == half - 1) er = e[k];              // mirror at the right edge
            o[k] = wadd(o[k], lift_delta<WIDE>(wadd(e[k], er), cp));
        }
        int d_nb = __shfl_up_sync(kFullMask, o[M - 1], 1);
#pragma unroll
        for (int k = 0; k < M; k++) {
            int dl = (k > 0) ? o[k - 1] : d_nb;
            if (EDGE && p0 + k == 0) dl = o[k];                     // mirror at the left edge
            e[k] = wadd(e[k], lift_delta<WIDE>(wadd(dl, o[k]), cu));
        }
    }
}

// Lane-parallel inverse along x: in: e[] = low[p0..], o[] = high[p0..]; out: e[] = s[2p], o[] = s[2p+1].
template <int WT, bool WIDE, int M, bool EDGE = true>
ALICE_D void inv_lanes(int (&e)[M], int (&o)[M], int p0, int half) {
#pragma unroll
    for (int st = WaveletTraits<WT>::NST - 1; st >= 0; st--) {
        const int cp = coef_p<WT>(st), cu = coef_u<WT>(st);
        int o_nb = __shfl_up_sync(kFullMask, o[M - 1], 1);
#pragma unroll
        for (int k = 0; k < M; k++) {
            int ol = (k > 0) ? o[k - 1] : o_nb;
            if (EDGE && p0 + k == 0) ol = o[k];
            e[k] = wadd(e[k], lift_delta<WIDE>(wadd(ol, o[k]), -cu));
        }
        int e_nb = __shfl_down_sync(kFullMask, e[0], 1);
#pragma unroll
        for (int k = 0; k < M; k++) {
            int er = (k + 1 < M) ? e[k + 1] : e_nb;
            if (EDGE && p0 + k == half - 1) er = e[k];
            o[k] = wadd(o[k], lift_delta<WIDE>(wadd(e[k], er), -cp));
        }
    }
}
#endif

}  // namespace alice
