// curve.cuh -- short-Weierstrass a = 0 group law (bn256 G1: y^2 = x^3 + 3 over Fq; grumpkin:
// y^2 = x^3 - 17 over Fr) in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2).
//
// The reference gets its group law from halo2curves (Jacobian) through `best_multiexp`
// (reference src/commitment.rs:83).  The sum of points is unique, so any complete addition law gives the
// same affine result; XYZZ is used here because the mixed addition (accumulator += affine base) is the
// inner loop of the MSM and costs 8M + 2S with no inversion.
//
// Every operation handles the exceptional cases explicitly (identity operands, P + P, P + (-P)), so results
// are exact for adversarial inputs (repeated bases, cancelling scalars) -- parity is bit-exact, not generic-case.
#pragma once
#include "field.cuh"

namespace sb {

enum : int { CURVE_BN256 = 0, CURVE_GRUMPKIN = 1 };

template <class F>
struct alignas(16) Affine {  // identity encoded as (0,0) (reference src/commitment.rs:43-45)
    F x, y;
    SB_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

template <class F>
struct alignas(16) XYZZ {  // identity: zz == 0
    F x, y, zz, zzz;
    SB_HD bool is_identity() const { return zz.is_zero(); }
    static SB_HD XYZZ identity() {
        XYZZ r;
        r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero();
        return r;
    }
    static SB_HD XYZZ from_affine(const Affine<F>& p) {
        XYZZ r;
        if (p.is_identity()) return identity();
        r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one();
        return r;
    }
};

// The latency-bound kernels (bucket fix-up, reduction tree, finalisation) instantiate the group law with
// INL = false: Montgomery products become calls to one out-of-line copy, which keeps those kernels' code
// inside the instruction cache.  The accumulate kernel (throughput-bound) uses the fully inlined forms.
#if defined(__CUDACC__)
template <class P>
__device__ __noinline__ Fe<P> mul_outlined(Fe<P> a, Fe<P> b) { return mul(a, b); }
#endif
// Two independent products in one out-of-line body: ptxas interleaves the two carry chains, so a warp that is
// alone on its scheduler (the reduction tree, the fix-up) gets ~2x the issue rate of back-to-back calls.
template <class P>
struct FePair { Fe<P> a, b; };
#if defined(__CUDACC__)
template <class P>
__device__ __noinline__ FePair<P> mul2_outlined(Fe<P> a0, Fe<P> b0, Fe<P> a1, Fe<P> b1) {
    FePair<P> r;
    r.a = mul(a0, b0);
    r.b = mul(a1, b1);
    return r;
}
#endif
template <bool INL, class P>
SB_HD void mul2x(Fe<P>& r0, const Fe<P>& a0, const Fe<P>& b0, Fe<P>& r1, const Fe<P>& a1, const Fe<P>& b1) {
#if defined(__CUDA_ARCH__)
    if constexpr (INL) {
        r0 = mul(a0, b0);
        r1 = mul(a1, b1);
    } else {
        FePair<P> t = mul2_outlined(a0, b0, a1, b1);
        r0 = t.a;
        r1 = t.b;
    }
#else
    r0 = mul(a0, b0);
    r1 = mul(a1, b1);
#endif
}

template <bool INL, class P>
SB_HD Fe<P> mulx(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__)
    if constexpr (INL) return mul(a, b);
    else return mul_outlined(a, b);
#else
    return mul(a, b);
#endif
}

// 2 * (affine p), p not the identity.  y = 0 cannot happen on a prime-order curve, but is handled.
template <class F>
SB_HD XYZZ<F> xyzz_double_affine(const F& px, const F& py) {
    XYZZ<F> r;
    if (py.is_zero()) return XYZZ<F>::identity();
    F u = dbl(py);
    F v = sqr(u);
    F w = mul(u, v);
    F s = mul(px, v);
    F xx = sqr(px);
    F m = add(dbl(xx), xx);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, py));
    r.zz = v;
    r.zzz = w;
    return r;
}

template <bool INL = true, class F>
SB_HD XYZZ<F> xyzz_double(const XYZZ<F>& p) {
    if (p.is_identity() || p.y.is_zero()) return XYZZ<F>::identity();
    XYZZ<F> r;
    F u = dbl(p.y);
    F v, xx, w, s, mm, wy, t;
    mul2x<INL>(v, u, u, xx, p.x, p.x);
    mul2x<INL>(w, u, v, s, p.x, v);
    F m = add(dbl(xx), xx);
    mul2x<INL>(mm, m, m, wy, w, p.y);
    r.x = sub(mm, dbl(s));
    mul2x<INL>(t, m, sub(s, r.x), r.zz, v, p.zz);
    r.y = sub(t, wy);
    r.zzz = mulx<INL>(w, p.zzz);
    return r;
}

// acc += (negate ? -q : q), q affine.
template <class F>
SB_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q, bool negate) {
    if (q.is_identity()) return;
    F qy = negate ? neg(q.y) : q.y;
    if (acc.is_identity()) {
        acc.x = q.x; acc.y = qy; acc.zz = F::one(); acc.zzz = F::one();
        return;
    }
    F u2 = mul(q.x, acc.zz);
    F s2 = mul(qy, acc.zzz);
    F p = sub(u2, acc.x);
    F r = sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_double_affine(q.x, qy);
        else acc = XYZZ<F>::identity();
        return;
    }
    F pp = sqr(p);
    F ppp = mul(p, pp);
    F qq = mul(acc.x, pp);
    F x3 = sub(sub(sqr(r), ppp), dbl(qq));
    F y3 = sub(mul(r, sub(qq, x3)), mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = mul(acc.zz, pp);
    acc.zzz = mul(acc.zzz, ppp);
}

// acc += (negate ? -q : q) with the accumulator's coordinates kept in the LAZY domain [0, 2p) (field.cuh): the ten
// products skip their final conditional subtraction.  q is canonical (a table entry); the identity is still
// zz == 0 exactly (a non-zero residue is never stored as 0).  canon_point() brings the accumulator back to [0, p)
// before it leaves the loop.
template <class F>
SB_HD void xyzz_madd_lazy(XYZZ<F>& acc, const Affine<F>& q, bool negate) {
    if (q.is_identity()) return;
    F qy = negate ? neg(q.y) : q.y;
    if (acc.is_identity()) {
        acc.x = q.x; acc.y = qy; acc.zz = F::one(); acc.zzz = F::one();
        return;
    }
    F u2 = mul_lazy(q.x, acc.zz);
    F s2 = mul_lazy(qy, acc.zzz);
    F p = sub_lazy(u2, acc.x);
    F r = sub_lazy(s2, acc.y);
    if (is_zero_lazy(p)) {
        if (is_zero_lazy(r)) acc = xyzz_double_affine(q.x, qy);
        else acc = XYZZ<F>::identity();
        return;
    }
    F pp = mul_lazy(p, p);
    F ppp = mul_lazy(p, pp);
    F qq = mul_lazy(acc.x, pp);
    F x3 = sub_lazy(sub_lazy(mul_lazy(r, r), ppp), dbl_lazy(qq));
    F y3 = sub_lazy(mul_lazy(r, sub_lazy(qq, x3)), mul_lazy(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = mul_lazy(acc.zz, pp);
    acc.zzz = mul_lazy(acc.zzz, ppp);
}
template <class F>
SB_HD XYZZ<F> canon_point(const XYZZ<F>& a) {
    XYZZ<F> r;
    r.x = canon(a.x); r.y = canon(a.y); r.zz = canon(a.zz); r.zzz = canon(a.zzz);
    return r;
}

// acc += q, both XYZZ.
template <bool INL = true, class F>
SB_HD void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) { acc = q; return; }
    F u1, u2, s1, s2;
    mul2x<INL>(u1, acc.x, q.zz, u2, q.x, acc.zz);
    mul2x<INL>(s1, acc.y, q.zzz, s2, q.y, acc.zzz);
    F p = sub(u2, u1);
    F r = sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_double<INL>(acc);
        else acc = XYZZ<F>::identity();
        return;
    }
    F pp, zz12, ppp, qq, rr, zzz12, t1, t2;
    mul2x<INL>(pp, p, p, zz12, acc.zz, q.zz);
    mul2x<INL>(ppp, p, pp, qq, u1, pp);
    mul2x<INL>(rr, r, r, zzz12, acc.zzz, q.zzz);
    F x3 = sub(sub(rr, ppp), dbl(qq));
    mul2x<INL>(t1, r, sub(qq, x3), t2, s1, ppp);
    acc.x = x3;
    acc.y = sub(t1, t2);
    mul2x<INL>(acc.zz, zz12, pp, acc.zzz, zzz12, ppp);
}

// Out-of-line entry points for the latency-bound kernels (bucket reduction tree, fix-up, combine): a warp that
// runs alone executes each straight-line addition once, so inlined copies are instruction-fetch bound; one shared
// copy of add/double (+ the shared products) stays resident in the instruction cache.
#if defined(__CUDACC__)
template <class F>
__device__ __noinline__ void xyzz_add_call(XYZZ<F>& acc, const XYZZ<F>& q) { xyzz_add<false>(acc, q); }
template <class F>
__device__ __noinline__ void xyzz_double_call(XYZZ<F>& p) { p = xyzz_double<false>(p); }
#endif

template <bool INL = true, class F>
SB_HD Affine<F> xyzz_to_affine(const XYZZ<F>& p) {
    Affine<F> r;
    if (p.is_identity()) { r.x = F::zero(); r.y = F::zero(); return r; }
    F i = inv_safegcd(mulx<INL>(p.zz, p.zzz));  // 1/(zz*zzz)
    F izz = mulx<INL>(i, p.zzz);               // 1/zz
    F izzz = mulx<INL>(i, p.zz);               // 1/zzz
    r.x = mulx<INL>(p.x, izz);
    r.y = mulx<INL>(p.y, izzz);
    return r;
}

}  // namespace sb
