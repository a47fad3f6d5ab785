// quad.cuh -- XYZZ addition / doubling executed by a group of 4 consecutive lanes.
//
// The tail of every commitment (weighted bucket sum, final combine) is a short chain of dependent XYZZ additions on
// a nearly idle chip; a single lane needs 14 dependent Montgomery products (13.3k cycles measured) per addition.
// Here the 14 products of one addition are issued in 4 rounds of up to 4 independent products, one per lane of the
// group, and broadcast inside the group with shuffles.  Operands and results are REPLICATED in the 4 lanes;
// shuffles use the group's own 4-lane mask, so different groups of a warp may diverge freely.
//
// Measured (profiles/r1_microbench3.txt): 12.8k cycles per quad-lane addition vs 13.3k for the single-lane one --
// the 120 partial-mask shuffles and the role selects cost what the shorter product chain saves.  The form is kept
// for the last, serial stage of the weighted bucket sum (it is correct, tested, and no slower); the real gain of
// that stage came from needing fewer dependent operations (digit sums instead of scan levels), and making these
// primitives pay off (full-mask shuffles, fewer broadcasts) is a round-2 item.
#pragma once
#include "curve.cuh"

namespace sb {

SB_D unsigned quad_mask() { return 0xFu << (threadIdx.x & 28u); }

template <class F>
SB_D F quad_bcast(const F& v, int role) {
    F r;
    const int src = (int)((threadIdx.x & 28u) | (unsigned)role);
    const unsigned m = quad_mask();
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(m, v.v[i], src);
    return r;
}

template <class F>
SB_D F quad_pick(int role, const F& a0, const F& a1, const F& a2, const F& a3) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = role == 0 ? a0.v[i] : role == 1 ? a1.v[i] : role == 2 ? a2.v[i] : a3.v[i];
    return r;
}

// acc += q   (all 4 lanes of the group call this with identical operands)
template <class F>
__device__ __noinline__ void quad_add(XYZZ<F>& acc, const XYZZ<F>& q) {
    const int role = threadIdx.x & 3;
    const bool acc_id = acc.is_identity(), q_id = q.is_identity();
    F m = mul_outlined(quad_pick(role, acc.x, q.x, acc.y, q.y), quad_pick(role, q.zz, acc.zz, q.zzz, acc.zzz));
    const F u1 = quad_bcast(m, 0), u2 = quad_bcast(m, 1), s1 = quad_bcast(m, 2), s2 = quad_bcast(m, 3);
    const F p = sub(u2, u1), r = sub(s2, s1);
    m = mul_outlined(quad_pick(role, p, acc.zz, acc.zzz, r), quad_pick(role, p, q.zz, q.zzz, r));
    const F pp = quad_bcast(m, 0), zz12 = quad_bcast(m, 1), zzz12 = quad_bcast(m, 2), rr = quad_bcast(m, 3);
    m = mul_outlined(quad_pick(role, p, u1, zz12, p), pp);
    const F ppp = quad_bcast(m, 0), qq = quad_bcast(m, 1), zz3 = quad_bcast(m, 2);
    const F x3 = sub(sub(rr, ppp), dbl(qq));
    m = mul_outlined(quad_pick(role, r, s1, zzz12, r), quad_pick(role, sub(qq, x3), ppp, ppp, ppp));
    const F t1 = quad_bcast(m, 0), t2 = quad_bcast(m, 1), zzz3 = quad_bcast(m, 2);
    // exceptional cases are uniform inside the group (replicated data) and contain no shuffles
    if (q_id) return;
    if (acc_id) { acc = q; return; }
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_double<false>(acc);
        else acc = XYZZ<F>::identity();
        return;
    }
    acc.x = x3;
    acc.y = sub(t1, t2);
    acc.zz = zz3;
    acc.zzz = zzz3;
}

// p = 2p
template <class F>
__device__ __noinline__ void quad_double(XYZZ<F>& p) {
    const int role = threadIdx.x & 3;
    const bool id = p.is_identity() || p.y.is_zero();
    const F u = dbl(p.y);
    F m = mul_outlined(quad_pick(role, u, p.x, u, p.x), quad_pick(role, u, p.x, u, p.x));
    const F v = quad_bcast(m, 0), xx = quad_bcast(m, 1);
    m = mul_outlined(quad_pick(role, u, p.x, u, p.x), v);
    const F w = quad_bcast(m, 0), s = quad_bcast(m, 1);
    const F mm3 = add(dbl(xx), xx);
    m = mul_outlined(quad_pick(role, mm3, w, v, w), quad_pick(role, mm3, p.y, p.zz, p.zzz));
    const F msq = quad_bcast(m, 0), wy = quad_bcast(m, 1), zz3 = quad_bcast(m, 2), zzz3 = quad_bcast(m, 3);
    const F x3 = sub(msq, dbl(s));
    m = mul_outlined(mm3, sub(s, x3));  // same product in all 4 lanes
    if (id) { p = XYZZ<F>::identity(); return; }
    p.x = x3;
    p.y = sub(m, wy);
    p.zz = zz3;
    p.zzz = zzz3;
}

// value held by the group `delta` groups further in the warp (full-warp shuffle: call with the warp converged)
template <class F>
SB_D XYZZ<F> group_shfl_down(const XYZZ<F>& v, int delta_groups) {
    XYZZ<F> r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 32; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta_groups * 4);
    return r;
}
template <class F>
SB_D XYZZ<F> group_shfl_xor(const XYZZ<F>& v, int mask_groups) {
    XYZZ<F> r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 32; i++) d[i] = __shfl_xor_sync(0xffffffffu, s[i], mask_groups * 4);
    return r;
}
// sum over the 8 groups of the warp; every group ends with the total
template <class F>
SB_D XYZZ<F> warp_group_sum(XYZZ<F> v) {
#pragma unroll 1
    for (int d = 4; d >= 1; d >>= 1) {
        XYZZ<F> t = group_shfl_xor(v, d);
        quad_add(v, t);
    }
    return v;
}

}  // namespace sb
