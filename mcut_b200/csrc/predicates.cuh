// mcut_b200/csrc/predicates.cuh — device arithmetic of the narrowphase.
//
// Semantics (bit-for-bit, IEEE binary64 round-to-nearest, no contraction) of the reference's
//   orient3d / orient3dadapt   source/shewchuk.c:2367-2410, :1962-2365   (constants :420-433)
//   orient2d / orient2dadapt   source/shewchuk.c:1695-1729, :1611-1693
//   compute_polygon_plane_coefficients      source/math.cpp:130-239
//   compute_segment_plane_intersection      source/math.cpp:249-287
//   calculate_projection_matrix             source/math.cpp:710-793
//   compute_point_in_polygon_test (3D, 2D)  source/math.cpp:851-902, :553-704
// The expansion arithmetic follows Shewchuk (1997).  Every product/sum below is written with explicit
// __dmul_rn/__dadd_rn/__dsub_rn so no compiler flag can fuse them.
#pragma once

#include "common.cuh"

namespace pred {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

#define MCB_SPLITTER 134217729.0
#define MCB_RESULTERRBOUND 3.3306690738754706e-16
#define MCB_CCWERRBOUND_A 3.3306690738754716e-16
#define MCB_CCWERRBOUND_B 2.2204460492503146e-16
#define MCB_CCWERRBOUND_C 1.1093356479670487e-31
#define MCB_O3DERRBOUND_A 7.7715611723761027e-16
#define MCB_O3DERRBOUND_B 3.3306690738754731e-16
#define MCB_O3DERRBOUND_C 3.2047474274603644e-31

struct dd {
    double hi, lo;
};

__device__ __forceinline__ dd fast_two_sum(double a, double b)
{
    dd r;
    r.hi = add(a, b);
    r.lo = sub(b, sub(r.hi, a));
    return r;
}
__device__ __forceinline__ dd two_sum(double a, double b)
{
    dd r;
    r.hi = add(a, b);
    const double bv = sub(r.hi, a);
    const double av = sub(r.hi, bv);
    r.lo = add(sub(a, av), sub(b, bv));
    return r;
}
__device__ __forceinline__ double two_diff_tail(double a, double b, double x)
{
    const double bv = sub(a, x);
    const double av = add(x, bv);
    return add(sub(a, av), sub(bv, b));
}
__device__ __forceinline__ dd two_diff(double a, double b)
{
    dd r;
    r.hi = sub(a, b);
    r.lo = two_diff_tail(a, b, r.hi);
    return r;
}
__device__ __forceinline__ dd split(double a)
{
    dd r;
    const double c = mul(MCB_SPLITTER, a);
    r.hi = sub(c, sub(c, a));
    r.lo = sub(a, r.hi);
    return r;
}
__device__ __forceinline__ dd two_product_presplit(double a, double b, dd bs)
{
    dd r;
    r.hi = mul(a, b);
    const dd as = split(a);
    const double e1 = sub(r.hi, mul(as.hi, bs.hi));
    const double e2 = sub(e1, mul(as.lo, bs.hi));
    const double e3 = sub(e2, mul(as.hi, bs.lo));
    r.lo = sub(mul(as.lo, bs.lo), e3);
    return r;
}
__device__ __forceinline__ dd two_product(double a, double b) { return two_product_presplit(a, b, split(b)); }

__device__ __forceinline__ void two_one_diff(double a1, double a0, double b, double& x2, double& x1, double& x0)
{
    const dd i = two_diff(a0, b);
    x0 = i.lo;
    const dd j = two_sum(a1, i.hi);
    x2 = j.hi;
    x1 = j.lo;
}
__device__ __forceinline__ void two_two_diff(dd a, dd b, double* x)
{
    double j, z;
    two_one_diff(a.hi, a.lo, b.lo, j, z, x[0]);
    two_one_diff(j, z, b.hi, x[3], x[2], x[1]);
}
__device__ __forceinline__ void two_one_product(dd a, double b, double* x)
{
    const dd bs = split(b);
    const dd i = two_product_presplit(a.lo, b, bs);
    x[0] = i.lo;
    const dd j = two_product_presplit(a.hi, b, bs);
    const dd k = two_sum(i.hi, j.lo);
    x[1] = k.lo;
    const dd l = fast_two_sum(j.hi, k.hi);
    x[3] = l.hi;
    x[2] = l.lo;
}

// h = e + f (both nonoverlapping, increasing magnitude), zero components dropped; returns length of h
__device__ __noinline__ int expansion_sum(int elen, const double* e, int flen, const double* f, double* h)
{
    double Q;
    int ei = 0, fi = 0, hi = 0;
    double enow = e[0], fnow = f[0];
    if ((fnow > enow) == (fnow > -enow)) {
        Q = enow;
        if (++ei < elen) enow = e[ei];
    } else {
        Q = fnow;
        if (++fi < flen) fnow = f[fi];
    }
    if (ei < elen && fi < flen) {
        dd s;
        if ((fnow > enow) == (fnow > -enow)) {
            s = fast_two_sum(enow, Q);
            if (++ei < elen) enow = e[ei];
        } else {
            s = fast_two_sum(fnow, Q);
            if (++fi < flen) fnow = f[fi];
        }
        Q = s.hi;
        if (s.lo != 0.0) h[hi++] = s.lo;
        while (ei < elen && fi < flen) {
            if ((fnow > enow) == (fnow > -enow)) {
                s = two_sum(Q, enow);
                if (++ei < elen) enow = e[ei];
            } else {
                s = two_sum(Q, fnow);
                if (++fi < flen) fnow = f[fi];
            }
            Q = s.hi;
            if (s.lo != 0.0) h[hi++] = s.lo;
        }
    }
    while (ei < elen) {
        const dd s = two_sum(Q, enow);
        if (++ei < elen) enow = e[ei];
        Q = s.hi;
        if (s.lo != 0.0) h[hi++] = s.lo;
    }
    while (fi < flen) {
        const dd s = two_sum(Q, fnow);
        if (++fi < flen) fnow = f[fi];
        Q = s.hi;
        if (s.lo != 0.0) h[hi++] = s.lo;
    }
    if (Q != 0.0 || hi == 0) h[hi++] = Q;
    return hi;
}

// h = e * b
__device__ __noinline__ int expansion_scale(int elen, const double* e, double b, double* h)
{
    const dd bs = split(b);
    dd p = two_product_presplit(e[0], b, bs);
    double Q = p.hi;
    int hi = 0;
    if (p.lo != 0.0) h[hi++] = p.lo;
    for (int i = 1; i < elen; ++i) {
        p = two_product_presplit(e[i], b, bs);
        const dd s = two_sum(Q, p.lo);
        if (s.lo != 0.0) h[hi++] = s.lo;
        const dd t = fast_two_sum(p.hi, s.hi);
        Q = t.hi;
        if (t.lo != 0.0) h[hi++] = t.lo;
    }
    if (Q != 0.0 || hi == 0) h[hi++] = Q;
    return hi;
}

__device__ __forceinline__ double expansion_estimate(int n, const double* e)
{
    double q = e[0];
    for (int i = 1; i < n; ++i) q = add(q, e[i]);
    return q;
}

// ---- orient2d ------------------------------------------------------------------------------------------------
__device__ __noinline__ double orient2d_adapt(const double* pa, const double* pb, const double* pc, double detsum)
{
    const double acx = sub(pa[0], pc[0]), bcx = sub(pb[0], pc[0]);
    const double acy = sub(pa[1], pc[1]), bcy = sub(pb[1], pc[1]);
    double B[4];
    two_two_diff(two_product(acx, bcy), two_product(acy, bcx), B);
    double det = expansion_estimate(4, B);
    double errbound = mul(MCB_CCWERRBOUND_B, detsum);
    if (det >= errbound || -det >= errbound) return det;

    const double acxt = two_diff_tail(pa[0], pc[0], acx), bcxt = two_diff_tail(pb[0], pc[0], bcx);
    const double acyt = two_diff_tail(pa[1], pc[1], acy), bcyt = two_diff_tail(pb[1], pc[1], bcy);
    if (acxt == 0.0 && acyt == 0.0 && bcxt == 0.0 && bcyt == 0.0) return det;

    errbound = add(mul(MCB_CCWERRBOUND_C, detsum), mul(MCB_RESULTERRBOUND, fabs(det)));
    det = add(det, sub(add(mul(acx, bcyt), mul(bcy, acxt)), add(mul(acy, bcxt), mul(bcx, acyt))));
    if (det >= errbound || -det >= errbound) return det;

    double u[4], C1[8], C2[12], D[16];
    two_two_diff(two_product(acxt, bcy), two_product(acyt, bcx), u);
    const int c1 = expansion_sum(4, B, 4, u, C1);
    two_two_diff(two_product(acx, bcyt), two_product(acy, bcxt), u);
    const int c2 = expansion_sum(c1, C1, 4, u, C2);
    two_two_diff(two_product(acxt, bcyt), two_product(acyt, bcxt), u);
    const int dl = expansion_sum(c2, C2, 4, u, D);
    return D[dl - 1];
}

__device__ __forceinline__ double orient2d(const double* pa, const double* pb, const double* pc)
{
    const double detleft = mul(sub(pa[0], pc[0]), sub(pb[1], pc[1]));
    const double detright = mul(sub(pa[1], pc[1]), sub(pb[0], pc[0]));
    const double det = sub(detleft, detright);
    double detsum;
    if (detleft > 0.0) {
        if (detright <= 0.0) return det;
        detsum = add(detleft, detright);
    } else if (detleft < 0.0) {
        if (detright >= 0.0) return det;
        detsum = sub(-detleft, detright);
    } else {
        return det;
    }
    const double errbound = mul(MCB_CCWERRBOUND_A, detsum);
    if (det >= errbound || -det >= errbound) return det;
    return orient2d_adapt(pa, pb, pc, detsum);
}

// ---- orient3d ------------------------------------------------------------------------------------------------
// stage A (shewchuk.c:2367-2410).  Returns det; `certain` tells whether |det| > errbound; `permanent` is handed to
// the adaptive stages.
__device__ __forceinline__ double orient3d_stageA(const double* pa, const double* pb, const double* pc, const double* pd,
    bool& certain, double& permanent)
{
    const double adx = sub(pa[0], pd[0]), bdx = sub(pb[0], pd[0]), cdx = sub(pc[0], pd[0]);
    const double ady = sub(pa[1], pd[1]), bdy = sub(pb[1], pd[1]), cdy = sub(pc[1], pd[1]);
    const double adz = sub(pa[2], pd[2]), bdz = sub(pb[2], pd[2]), cdz = sub(pc[2], pd[2]);
    const double bdxcdy = mul(bdx, cdy), cdxbdy = mul(cdx, bdy);
    const double cdxady = mul(cdx, ady), adxcdy = mul(adx, cdy);
    const double adxbdy = mul(adx, bdy), bdxady = mul(bdx, ady);
    const double det = add(add(mul(adz, sub(bdxcdy, cdxbdy)), mul(bdz, sub(cdxady, adxcdy))), mul(cdz, sub(adxbdy, bdxady)));
    permanent = add(add(mul(add(fabs(bdxcdy), fabs(cdxbdy)), fabs(adz)), mul(add(fabs(cdxady), fabs(adxcdy)), fabs(bdz))),
        mul(add(fabs(adxbdy), fabs(bdxady)), fabs(cdz)));
    const double errbound = mul(MCB_O3DERRBOUND_A, permanent);
    certain = (det > errbound) || (-det > errbound);
    return det;
}

// t_m = xt*my - yt*mx ; t_n = yt*nx - xt*ny  as expansions of length 1, 2 or 4 (stage D of orient3dadapt)
__device__ __forceinline__ void tail_cross(double xt, double yt, double mx, double my, double nx, double ny, double* tm,
    int& tmlen, double* tn, int& tnlen)
{
    if (xt == 0.0) {
        if (yt == 0.0) {
            tm[0] = 0.0;
            tmlen = 1;
            tn[0] = 0.0;
            tnlen = 1;
        } else {
            dd p = two_product(-yt, mx);
            tm[0] = p.lo;
            tm[1] = p.hi;
            tmlen = 2;
            p = two_product(yt, nx);
            tn[0] = p.lo;
            tn[1] = p.hi;
            tnlen = 2;
        }
    } else if (yt == 0.0) {
        dd p = two_product(xt, my);
        tm[0] = p.lo;
        tm[1] = p.hi;
        tmlen = 2;
        p = two_product(-xt, ny);
        tn[0] = p.lo;
        tn[1] = p.hi;
        tnlen = 2;
    } else {
        two_two_diff(two_product(xt, my), two_product(yt, mx), tm);
        tmlen = 4;
        two_two_diff(two_product(yt, nx), two_product(xt, ny), tn);
        tnlen = 4;
    }
}

struct fin_t {
    double buf[2][192];
    int cur, len;
    __device__ __forceinline__ void accumulate(int n, const double* e)
    {
        len = expansion_sum(len, buf[cur], n, e, buf[cur ^ 1]);
        cur ^= 1;
    }
};

// stages B, C, D (shewchuk.c:1962-2365)
__device__ __noinline__ double orient3d_adapt(const double* pa, const double* pb, const double* pc, const double* pd,
    double permanent)
{
    const double adx = sub(pa[0], pd[0]), bdx = sub(pb[0], pd[0]), cdx = sub(pc[0], pd[0]);
    const double ady = sub(pa[1], pd[1]), bdy = sub(pb[1], pd[1]), cdy = sub(pc[1], pd[1]);
    const double adz = sub(pa[2], pd[2]), bdz = sub(pb[2], pd[2]), cdz = sub(pc[2], pd[2]);

    double bc[4], ca[4], ab[4], adet[8], bdet[8], cdet[8], abdet[16];
    two_two_diff(two_product(bdx, cdy), two_product(cdx, bdy), bc);
    const int alen = expansion_scale(4, bc, adz, adet);
    two_two_diff(two_product(cdx, ady), two_product(adx, cdy), ca);
    const int blen = expansion_scale(4, ca, bdz, bdet);
    two_two_diff(two_product(adx, bdy), two_product(bdx, ady), ab);
    const int clen = expansion_scale(4, ab, cdz, cdet);

    fin_t fin;
    fin.cur = 0;
    const int ablen = expansion_sum(alen, adet, blen, bdet, abdet);
    fin.len = expansion_sum(ablen, abdet, clen, cdet, fin.buf[0]);

    double det = expansion_estimate(fin.len, fin.buf[0]);
    double errbound = mul(MCB_O3DERRBOUND_B, permanent);
    if (det >= errbound || -det >= errbound) return det;

    const double adxt = two_diff_tail(pa[0], pd[0], adx), bdxt = two_diff_tail(pb[0], pd[0], bdx),
                 cdxt = two_diff_tail(pc[0], pd[0], cdx);
    const double adyt = two_diff_tail(pa[1], pd[1], ady), bdyt = two_diff_tail(pb[1], pd[1], bdy),
                 cdyt = two_diff_tail(pc[1], pd[1], cdy);
    const double adzt = two_diff_tail(pa[2], pd[2], adz), bdzt = two_diff_tail(pb[2], pd[2], bdz),
                 cdzt = two_diff_tail(pc[2], pd[2], cdz);
    if (adxt == 0.0 && bdxt == 0.0 && cdxt == 0.0 && adyt == 0.0 && bdyt == 0.0 && cdyt == 0.0 && adzt == 0.0 && bdzt == 0.0
        && cdzt == 0.0)
        return det;

    errbound = add(mul(MCB_O3DERRBOUND_C, permanent), mul(MCB_RESULTERRBOUND, fabs(det)));
    {
        const double ta = add(mul(adz, sub(add(mul(bdx, cdyt), mul(cdy, bdxt)), add(mul(bdy, cdxt), mul(cdx, bdyt)))),
            mul(adzt, sub(mul(bdx, cdy), mul(bdy, cdx))));
        const double tb = add(mul(bdz, sub(add(mul(cdx, adyt), mul(ady, cdxt)), add(mul(cdy, adxt), mul(adx, cdyt)))),
            mul(bdzt, sub(mul(cdx, ady), mul(cdy, adx))));
        const double tc = add(mul(cdz, sub(add(mul(adx, bdyt), mul(bdy, adxt)), add(mul(ady, bdxt), mul(bdx, adyt)))),
            mul(cdzt, sub(mul(adx, bdy), mul(ady, bdx))));
        det = add(det, add(add(ta, tb), tc));
    }
    if (det >= errbound || -det >= errbound) return det;

    double at_b[4], at_c[4], bt_c[4], bt_a[4], ct_a[4], ct_b[4];
    int at_bl, at_cl, bt_cl, bt_al, ct_al, ct_bl;
    tail_cross(adxt, adyt, bdx, bdy, cdx, cdy, at_b, at_bl, at_c, at_cl);
    tail_cross(bdxt, bdyt, cdx, cdy, adx, ady, bt_c, bt_cl, bt_a, bt_al);
    tail_cross(cdxt, cdyt, adx, ady, bdx, bdy, ct_a, ct_al, ct_b, ct_bl);

    double bct[8], cat[8], abt[8], w[16], v[12], u[4];
    const int bctl = expansion_sum(bt_cl, bt_c, ct_bl, ct_b, bct);
    fin.accumulate(expansion_scale(bctl, bct, adz, w), w);
    const int catl = expansion_sum(ct_al, ct_a, at_cl, at_c, cat);
    fin.accumulate(expansion_scale(catl, cat, bdz, w), w);
    const int abtl = expansion_sum(at_bl, at_b, bt_al, bt_a, abt);
    fin.accumulate(expansion_scale(abtl, abt, cdz, w), w);

    if (adzt != 0.0) fin.accumulate(expansion_scale(4, bc, adzt, v), v);
    if (bdzt != 0.0) fin.accumulate(expansion_scale(4, ca, bdzt, v), v);
    if (cdzt != 0.0) fin.accumulate(expansion_scale(4, ab, cdzt, v), v);

#define MCB_TT(xt, yt, z, zt)                   \
    do {                                        \
        const dd p_ = two_product((xt), (yt));  \
        two_one_product(p_, (z), u);            \
        fin.accumulate(4, u);                   \
        if ((zt) != 0.0) {                      \
            two_one_product(p_, (zt), u);       \
            fin.accumulate(4, u);               \
        }                                       \
    } while (0)
    if (adxt != 0.0) {
        if (bdyt != 0.0) MCB_TT(adxt, bdyt, cdz, cdzt);
        if (cdyt != 0.0) MCB_TT(-adxt, cdyt, bdz, bdzt);
    }
    if (bdxt != 0.0) {
        if (cdyt != 0.0) MCB_TT(bdxt, cdyt, adz, adzt);
        if (adyt != 0.0) MCB_TT(-bdxt, adyt, cdz, cdzt);
    }
    if (cdxt != 0.0) {
        if (adyt != 0.0) MCB_TT(cdxt, adyt, bdz, bdzt);
        if (bdyt != 0.0) MCB_TT(-cdxt, bdyt, adz, adzt);
    }
#undef MCB_TT
    if (adzt != 0.0) fin.accumulate(expansion_scale(bctl, bct, adzt, w), w);
    if (bdzt != 0.0) fin.accumulate(expansion_scale(catl, cat, bdzt, w), w);
    if (cdzt != 0.0) fin.accumulate(expansion_scale(abtl, abt, cdzt, w), w);

    return fin.buf[fin.cur][fin.len - 1];
}

// ---- plane of a polygon (math.cpp:130-239) ----------------------------------------------------------------------
// ---- device-side shortcuts that reproduce orient3d's SIGN without following its arithmetic ---------------------------
// The narrowphase consumes nothing of orient3d but the sign of its result (kernel.cpp:2483-2516 via math.cpp:416-426), and
// that sign is by construction the sign of the exact determinant (Shewchuk 1997, sec. 4.4).  Two consequences used here.

// (1) Side prefilter.  For a tested triangle T0 T1 T2 and a query point P, with u = T1-T0, v = T2-T0, w = P-T0 computed in
// binary64, det' = (u x v) . w and L >= every |u_i|, |v_i|, |w_i|:
//   |det' - D| <= 48 e L^3          (D the exact determinant, e = 2^-53: three roundings per difference, five per monomial)
//   permanent_A <= 24 L^3 (1+10e)   (rows of stage A are T0-P, T1-P, T2-P: entries <= L, 2L, 2L)
// so |det'| > 2^-43 L^3 = 1024 e L^3 implies |D| > 976 e L^3, hence stage A's own estimate exceeds
// (976 - 168) e L^3 > o3derrboundA * permanent_A: stage A certifies the sign, and the sign is sign(D) = -sign(det').
// Returns +1 / -1 = orient3d's certified sign, 0 = not decided here (the caller runs the real stage A).
__device__ __forceinline__ int orient3d_side_prefilter(const double* nrm, double luv, const double* t0, const double* p)
{
    const double wx = p[0] - t0[0], wy = p[1] - t0[1], wz = p[2] - t0[2];
    const double l = fmax(fmax(luv, fabs(wx)), fmax(fabs(wy), fabs(wz)));
    const double det = nrm[0] * wx + nrm[1] * wy + nrm[2] * wz;
    const double thr = l * l * l * 0x1p-43;
    return det > thr ? -1 : (det < -thr ? 1 : 0); // (a NaN or an overflow compares false twice: undecided)
}
__device__ __forceinline__ double side_prefilter_plane(const double* t0, const double* t1, const double* t2, double* nrm)
{
    const double ux = t1[0] - t0[0], uy = t1[1] - t0[1], uz = t1[2] - t0[2];
    const double vx = t2[0] - t0[0], vy = t2[1] - t0[1], vz = t2[2] - t0[2];
    nrm[0] = uy * vz - uz * vy;
    nrm[1] = uz * vx - ux * vz;
    nrm[2] = ux * vy - uy * vx;
    return fmax(fmax(fmax(fabs(ux), fabs(uy)), fmax(fabs(uz), fabs(vx))), fmax(fabs(vy), fabs(vz)));
}

// (2) Exact sign when the nine differences of stage A are exact (all tails zero: orient3dadapt then returns at stage B
// with an estimate of the exact determinant of those differences, shewchuk.c:2041-2075).  The determinant is written as
// 24 doubles whose sum is exact — each of the six triple products x*y*z as two_product(x,y) = (p,e), then two_product(p,z)
// and two_product(e,z), with the fused multiply-add giving each error term in one instruction — and the sign of the sum
// is found by error-free distillation passes (t[i], t[i-1] <- two_sum) until the leading term dominates what is left.
// Everything is indexed statically: the 24 terms live in registers.  Returns false if five passes did not decide (the
// caller falls back to orient3d_adapt); products that underflow are excluded by the caller's magnitude guard.
__device__ __forceinline__ void exact_triple(double x, double y, double z, double* t)
{
    const double p = __dmul_rn(x, y), e = __fma_rn(x, y, -p);
    t[0] = __dmul_rn(e, z);
    t[1] = __fma_rn(e, z, -t[0]);
    t[2] = __dmul_rn(p, z);
    t[3] = __fma_rn(p, z, -t[2]);
}
__device__ __forceinline__ bool det3_sign_exact(double adx, double ady, double adz, double bdx, double bdy, double bdz, double cdx,
    double cdy, double cdz, int& sign)
{
    double t[24];
    exact_triple(bdx, cdy, adz, t);
    exact_triple(-cdx, bdy, adz, t + 4);
    exact_triple(cdx, ady, bdz, t + 8);
    exact_triple(-adx, cdy, bdz, t + 12);
    exact_triple(adx, bdy, cdz, t + 16);
    exact_triple(-bdx, ady, cdz, t + 20);
    // guard: a term below 2^-960 in magnitude may have lost bits to underflow in its error term
    double mn = 1.0;
#pragma unroll
    for (int i = 2; i < 24; i += 4) mn = fmin(mn, t[i] == 0.0 ? 1.0 : fabs(t[i]));
    if (mn < 0x1p-800) return false;
#pragma unroll 1
    for (int pass = 0; pass < 6; ++pass) {
#pragma unroll
        for (int i = 1; i < 24; ++i) {
            const dd r = two_sum(t[i], t[i - 1]);
            t[i] = r.hi;
            t[i - 1] = r.lo;
        }
        double rest = 0.0;
#pragma unroll
        for (int i = 0; i < 23; ++i) rest += fabs(t[i]);
        if (rest == 0.0 || fabs(t[23]) > 2.0 * rest) {
            sign = (t[23] > 0.0) - (t[23] < 0.0);
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ double dot3(const double* a, const double* b)
{
    // math.h:634-642: accumulates from 0.0 in x, y, z order
    double out = 0.0;
    out = add(out, mul(a[0], b[0]));
    out = add(out, mul(a[1], b[1]));
    out = add(out, mul(a[2], b[2]));
    return out;
}

// Finishes a Newell sum: returns max_comp; `normal` becomes the unit normal (or zero when degenerate)
__device__ __forceinline__ int plane_finish(double* normal, const double* v0, double& d)
{
    d = 0.0;
    if (isnan(normal[0]) || isnan(normal[1]) || isnan(normal[2]) || dot3(normal, normal) < 1e-9) {
        normal[0] = normal[1] = normal[2] = 0.0;
        return 0;
    }
    const double len = sqrt(dot3(normal, normal));
    normal[0] = normal[0] / len;
    normal[1] = normal[1] / len;
    normal[2] = normal[2] / len;
    d = dot3(v0, normal);
    double largest = 0.0;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double t = fabs(normal[i]);
        if (t > largest) {
            largest = t;
            idx = i;
        }
    }
    return idx;
}

__device__ __forceinline__ void newell_step(double* normal, const double* c, const double* x)
{
    normal[0] = add(normal[0], mul(sub(c[1], x[1]), add(c[2], x[2])));
    normal[1] = add(normal[1], mul(sub(c[2], x[2]), add(c[0], x[0])));
    normal[2] = add(normal[2], mul(sub(c[0], x[0]), add(c[1], x[1])));
}

// triangle fast path
__device__ __forceinline__ int plane_tri(const double* v0, const double* v1, const double* v2, double* normal, double& d)
{
    normal[0] = normal[1] = normal[2] = 0.0;
    newell_step(normal, v0, v1);
    newell_step(normal, v1, v2);
    newell_step(normal, v2, v0);
    return plane_finish(normal, v0, d);
}

// ---- segment / plane (math.cpp:249-287) ----------------------------------------------------------------------------
__device__ __forceinline__ void segment_plane_point(double* p, const double* normal, double d, const double* q, const double* r)
{
    const double num = sub(d, dot3(q, normal));
    const double rq[3] = { sub(r[0], q[0]), sub(r[1], q[1]), sub(r[2], q[2]) };
    const double denom = dot3(rq, normal);
    if (denom == 0.0) { // parallel: the reference leaves p = (0,0,0) (kernel.cpp:2505)
        p[0] = p[1] = p[2] = 0.0;
        return;
    }
    const double t = num / denom;
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = add(q[i], mul(t, sub(r[i], q[i])));
}

// ---- projection matrix (math.cpp:710-793), row-major 2x3 ---------------------------------------------------------------
__device__ __forceinline__ void projection_matrix(const double* normal, int max_comp, double* P)
{
    const double len = sqrt(dot3(normal, normal));
    const double a[3] = { normal[0] / len, normal[1] / len, normal[2] / len };
    double b[3] = { 0.0, 0.0, 0.0 };
    const double nm = max_comp == 0 ? normal[0] : (max_comp == 1 ? normal[1] : normal[2]);
    const int s = (0.0 < nm) - (nm < 0.0);
    const double bv = mul(1.0, (double)s);
    if (max_comp == 0) b[0] = bv;
    else if (max_comp == 1) b[1] = bv;
    else b[2] = bv;
    double R[3][3] = { { 1.0, 0.0, 0.0 }, { 0.0, -1.0, 0.0 }, { 0.0, 0.0, 1.0 } };
    if (a[0] != b[0] || a[1] != b[1] || a[2] != b[2]) {
        const double apb[3] = { add(a[0], b[0]), add(a[1], b[1]), add(a[2], b[2]) };
        const double adb = dot3(a, b);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double I = (i == j) ? (i == 1 ? -1.0 : 1.0) : 0.0;
                R[i][j] = sub(mul(mul(apb[i], apb[j]) / adb, 2.0), I);
            }
    }
    // K selects the two kept axes; K*R through the generic triple loop that accumulates from 0.0 (math.h:450-463)
    const int r0 = (max_comp == 0) ? 1 : 0;
    const int r1 = (max_comp == 2) ? 1 : 2;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            acc0 = add(acc0, mul(k == r0 ? 1.0 : 0.0, R[k][j]));
            acc1 = add(acc1, mul(k == r1 ? 1.0 : 0.0, R[k][j]));
        }
        P[j] = acc0;
        P[3 + j] = acc1;
    }
}

__device__ __forceinline__ void project2(const double* P, const double* v, double* out)
{
    // math.h:508-533: result[row] = result[row] + P(row, col) * v[col], columns outermost
    out[0] = 0.0;
    out[1] = 0.0;
#pragma unroll
    for (int col = 0; col < 3; ++col) {
        out[0] = add(out[0], mul(P[col], v[col]));
        out[1] = add(out[1], mul(P[3 + col], v[col]));
    }
}

// O'Rourke crossing test on a triangle already shifted so the query is the origin (math.cpp:643-703)
__device__ __forceinline__ char pip2d_shifted(const double (*v)[2], int n)
{
    int rcross = 0, lcross = 0;
    for (int i = 0; i < n; ++i) {
        const double xi = v[i][0], yi = v[i][1];
        if (xi == 0.0 && yi == 0.0) return 'v';
        const int il = (i + n - 1) % n;
        const double xl = v[il][0], yl = v[il][1];
        const bool rstrad = (yi > 0.0) != (yl > 0.0);
        const bool lstrad = (yi < 0.0) != (yl < 0.0);
        if (rstrad || lstrad) {
            const double x = sub(mul(xi, yl), mul(xl, yi)) / sub(yl, yi);
            if (rstrad && x > 0.0) rcross++;
            if (lstrad && x < 0.0) lcross++;
        }
    }
    if ((rcross & 1) != (lcross & 1)) return 'e';
    return (rcross & 1) ? 'i' : 'o';
}

__device__ __forceinline__ char point_in_triangle(const double* p, const double* v0, const double* v1, const double* v2,
    const double* normal, int max_comp)
{
    double P[6], pp[2], t[3][2];
    projection_matrix(normal, max_comp, P);
    project2(P, p, pp);
    project2(P, v0, t[0]);
    project2(P, v1, t[1]);
    project2(P, v2, t[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        t[i][0] = sub(t[i][0], pp[0]);
        t[i][1] = sub(t[i][1], pp[1]);
    }
    return pip2d_shifted(t, 3);
}

} // namespace pred
