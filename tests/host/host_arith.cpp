// Host build of the PORTABLE code paths in sirius_b200/csrc/{field,curve}.cuh, so the formulas the CUDA
// kernels use can be checked against the oracle on a machine without a GPU (pytest -m "not gpu").
#include <cstddef>
#include <cstring>
#include "../../sirius_b200/csrc/curve.cuh"

using namespace sb;

template <class F>
static void t_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
    const F* A = (const F*)a; const F* B = (const F*)b; F* O = (F*)o;
    for (size_t i = 0; i < n; i++) O[i] = mul(A[i], B[i]);
}
template <class F>
static void t_addsub(const uint64_t* a, const uint64_t* b, uint64_t* o_add, uint64_t* o_sub, size_t n) {
    const F* A = (const F*)a; const F* B = (const F*)b;
    for (size_t i = 0; i < n; i++) { ((F*)o_add)[i] = add(A[i], B[i]); ((F*)o_sub)[i] = sub(A[i], B[i]); }
}
template <class F>
static void t_inv(const uint64_t* a, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; i++) ((F*)o)[i] = (i & 1) ? inv(((const F*)a)[i]) : inv_binary(((const F*)a)[i]);
}
template <class F>
static int t_inv_safegcd(const uint64_t* a, uint64_t* o, size_t n) {
    for (size_t i = 0; i < n; i++) ((F*)o)[i] = inv_safegcd(((const F*)a)[i]);
    return 0;
}
// sum_i (+/-) P_i with madd, then tree-combine two halves with xyzz_add, double once, -> affine
template <class F>
static void t_points(const uint64_t* pts, const uint8_t* negs, size_t n, uint64_t* out_sum, uint64_t* out_dbl) {
    const Affine<F>* P = (const Affine<F>*)pts;
    XYZZ<F> lo = XYZZ<F>::identity(), hi = XYZZ<F>::identity();
    for (size_t i = 0; i < n / 2; i++) xyzz_madd(lo, P[i], negs[i] != 0);
    for (size_t i = n / 2; i < n; i++) xyzz_madd(hi, P[i], negs[i] != 0);
    xyzz_add(lo, hi);
    Affine<F> s = xyzz_to_affine(lo);
    memcpy(out_sum, &s, 64);
    Affine<F> d = xyzz_to_affine(xyzz_double(lo));
    memcpy(out_dbl, &d, 64);
}

// lazy-domain twins (field.cuh "lazy domain"): raw [0, 2p) results, canonicalised by the caller
template <class F>
static void t_lazy(const uint64_t* a, const uint64_t* b, uint64_t* o_mul, uint64_t* o_sub, uint64_t* o_dbl, uint8_t* o_zero, size_t n) {
    const F* A = (const F*)a; const F* B = (const F*)b;
    for (size_t i = 0; i < n; i++) {
        ((F*)o_mul)[i] = mul_lazy(A[i], B[i]);
        ((F*)o_sub)[i] = sub_lazy(A[i], B[i]);
        ((F*)o_dbl)[i] = (i & 1) ? dbl_lazy(A[i]) : add_lazy(A[i], A[i]);
        ((F*)o_sub)[i] = (i & 1) ? sub_lazy(A[i], B[i]) : add_lazy(A[i], neg_lazy(B[i]));
        o_zero[i] = is_zero_lazy(A[i]) ? 1 : 0;
    }
}
// the same sums as t_points through xyzz_madd_lazy, with `chain` products of slack: the accumulator is fed back
// un-canonicalised for the whole run, as in k_accumulate
template <class F>
static void t_points_lazy(const uint64_t* pts, const uint8_t* negs, size_t n, uint64_t* out_sum, uint8_t* out_in_range) {
    const Affine<F>* P = (const Affine<F>*)pts;
    XYZZ<F> acc = XYZZ<F>::identity();
    uint8_t ok = 1;
    F twop;
    for (int k = 0; k < 8; k++) twop.v[k] = TwoP<typename F::Params>::limb(k);
    for (size_t i = 0; i < n; i++) {
        xyzz_madd_lazy(acc, P[i], negs[i] != 0);
        const F* c[4] = {&acc.x, &acc.y, &acc.zz, &acc.zzz};
        for (int j = 0; j < 4; j++) ok &= limbs_geq(c[j]->v, twop.v) ? 0 : 1;  // every coordinate stays below 2p
    }
    Affine<F> s = xyzz_to_affine(canon_point(acc));
    memcpy(out_sum, &s, 64);
    *out_in_range = ok;
}

extern "C" {
void ha_lazy(int field, const uint64_t* a, const uint64_t* b, uint64_t* om, uint64_t* os, uint64_t* od, uint8_t* oz, size_t n) {
    if (field == FIELD_FR) t_lazy<Fr>(a, b, om, os, od, oz, n); else t_lazy<Fq>(a, b, om, os, od, oz, n);
}
void ha_points_lazy(int curve, const uint64_t* pts, const uint8_t* negs, size_t n, uint64_t* out_sum, uint8_t* ok) {
    if (curve == CURVE_BN256) t_points_lazy<Fq>(pts, negs, n, out_sum, ok); else t_points_lazy<Fr>(pts, negs, n, out_sum, ok);
}
void ha_mul(int field, const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
    if (field == FIELD_FR) t_mul<Fr>(a, b, o, n); else t_mul<Fq>(a, b, o, n);
}
void ha_addsub(int field, const uint64_t* a, const uint64_t* b, uint64_t* oa, uint64_t* os, size_t n) {
    if (field == FIELD_FR) t_addsub<Fr>(a, b, oa, os, n); else t_addsub<Fq>(a, b, oa, os, n);
}
void ha_inv(int field, const uint64_t* a, uint64_t* o, size_t n) {
    if (field == FIELD_FR) t_inv<Fr>(a, o, n); else t_inv<Fq>(a, o, n);
}
void ha_inv_safegcd(int field, const uint64_t* a, uint64_t* o, size_t n) {
    if (field == FIELD_FR) t_inv_safegcd<Fr>(a, o, n); else t_inv_safegcd<Fq>(a, o, n);
}
void ha_points(int curve, const uint64_t* pts, const uint8_t* negs, size_t n, uint64_t* out_sum, uint64_t* out_dbl) {
    if (curve == CURVE_BN256) t_points<Fq>(pts, negs, n, out_sum, out_dbl); else t_points<Fr>(pts, negs, n, out_sum, out_dbl);
}
}
