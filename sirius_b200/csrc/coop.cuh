// coop.cuh -- XYZZ addition / doubling executed by the FOUR WARPS of a 128-thread block.
//
// The tail of every commitment (bucket fix-up, row/column sums, the weighted sum of the two short vectors) is a chain
// of ~40 dependent point additions on a nearly idle chip.  A lone warp needs 13.2k cycles per addition: 14 Montgomery
// products of ~830 cycles each, one after the other, on ONE scheduler's quarter-rate IMAD.WIDE pipe.  Spreading the
// products of one addition over the lanes of a warp does not help (quad.cuh: the lanes share that one pipe and the
// shuffles cost what the shorter chain saves).  Here they are spread over the four SCHEDULERS of an SM instead:
//
//   * lane l of EVERY warp of the block holds the same operands of addition #l (32 independent additions per block,
//     operands and results replicated across the 4 warps);
//   * the 14 products are issued in 4 rounds; warp w (= its own scheduler and integer pipe) computes product w of each
//     round for all 32 lanes and publishes it through shared memory (limb-major, conflict-free), one barrier per round.
//
// Latency per addition = 4 rounds x (one product + exchange) instead of 14 products; 14 of the 16 product slots do
// useful work.  All 128 threads must call these functions together (they contain __syncthreads) with block-uniform
// trip counts; a lane with nothing to add passes the identity.
#pragma once
#include "curve.cuh"

namespace sb {

constexpr int COOP_THREADS = 128;

struct alignas(16) CoopBuf {
    uint32_t w[2][8][4][32];  // [ping-pong][limb][product slot][lane]
};

template <class F>
SB_D void coop_put(CoopBuf& sh, int buf, int slot, int lane, const F& m) {
#pragma unroll
    for (int i = 0; i < 8; i++) sh.w[buf][i][slot][lane] = m.v[i];
}
template <class F>
SB_D F coop_get(const CoopBuf& sh, int buf, int slot, int lane) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = sh.w[buf][i][slot][lane];
    return r;
}

// p = 2p for the block's 32 logical lanes
template <class F>
__device__ __noinline__ void coop4_double(XYZZ<F>& p, CoopBuf& sh) {
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool id = p.is_identity() || p.y.is_zero();
    const F u = dbl(p.y);
    F m, a, b;
    // round 1: v = u^2, xx = x^2
    if (role == 0) { a = u; b = u; }
    else { a = p.x; b = p.x; }
    if (role < 2) {   // warps 2 and 3 have no product in this round
        m = mul_outlined(a, b);
        coop_put(sh, 0, role, lane, m);
    }
    __syncthreads();
    const F v = coop_get<F>(sh, 0, 0, lane), xx = coop_get<F>(sh, 0, 1, lane);
    const F mm3 = add(dbl(xx), xx);
    // round 2: w = u v, s = x v, msq = m^2, zz3 = v zz
    if (role == 0) { a = u; b = v; }
    else if (role == 1) { a = p.x; b = v; }
    else if (role == 2) { a = mm3; b = mm3; }
    else { a = v; b = p.zz; }
    m = mul_outlined(a, b);
    coop_put(sh, 1, role, lane, m);
    __syncthreads();
    const F w = coop_get<F>(sh, 1, 0, lane), s = coop_get<F>(sh, 1, 1, lane), msq = coop_get<F>(sh, 1, 2, lane), zz3 = coop_get<F>(sh, 1, 3, lane);
    const F x3 = sub(msq, dbl(s));
    // round 3: t = m (s - x3), wy = w y, zzz3 = w zzz
    if (role == 0) { a = mm3; b = sub(s, x3); }
    else if (role == 1) { a = w; b = p.y; }
    else { a = w; b = p.zzz; }
    if (role < 3) {   // warp 3 has no product in this round
        m = mul_outlined(a, b);
        coop_put(sh, 0, role, lane, m);
    }
    __syncthreads();
    const F t = coop_get<F>(sh, 0, 0, lane), wy = coop_get<F>(sh, 0, 1, lane), zzz3 = coop_get<F>(sh, 0, 2, lane);
    // keep the ping-pong phase of coop4_add (which ends having written buffer 1): an empty round
    __syncthreads();
    if (id) {
        p = XYZZ<F>::identity();
        return;
    }
    p.x = x3;
    p.y = sub(t, wy);
    p.zz = zz3;
    p.zzz = zzz3;
}

// acc += q for the block's 32 logical lanes
template <class F>
__device__ __noinline__ void coop4_add(XYZZ<F>& acc, const XYZZ<F>& q, CoopBuf& sh) {
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool acc_id = acc.is_identity(), q_id = q.is_identity();
    F m, a, b;
    // round 1: u1 = X1 ZZ2, u2 = X2 ZZ1, s1 = Y1 ZZZ2, s2 = Y2 ZZZ1
    if (role == 0) { a = acc.x; b = q.zz; }
    else if (role == 1) { a = q.x; b = acc.zz; }
    else if (role == 2) { a = acc.y; b = q.zzz; }
    else { a = q.y; b = acc.zzz; }
    m = mul_outlined(a, b);
    coop_put(sh, 0, role, lane, m);
    __syncthreads();
    const F u1 = coop_get<F>(sh, 0, 0, lane), u2 = coop_get<F>(sh, 0, 1, lane), s1 = coop_get<F>(sh, 0, 2, lane), s2 = coop_get<F>(sh, 0, 3, lane);
    const F p = sub(u2, u1), r = sub(s2, s1);
    // round 2: pp = p^2, zz12 = ZZ1 ZZ2, zzz12 = ZZZ1 ZZZ2, rr = r^2
    if (role == 0) { a = p; b = p; }
    else if (role == 1) { a = acc.zz; b = q.zz; }
    else if (role == 2) { a = acc.zzz; b = q.zzz; }
    else { a = r; b = r; }
    m = mul_outlined(a, b);
    coop_put(sh, 1, role, lane, m);
    __syncthreads();
    const F pp = coop_get<F>(sh, 1, 0, lane), zz12 = coop_get<F>(sh, 1, 1, lane), zzz12 = coop_get<F>(sh, 1, 2, lane), rr = coop_get<F>(sh, 1, 3, lane);
    // round 3: ppp = p pp, qq = u1 pp, zz3 = zz12 pp
    if (role == 0) { a = p; b = pp; }
    else if (role == 1) { a = u1; b = pp; }
    else { a = zz12; b = pp; }
    if (role < 3) {   // warp 3 has no product in this round
        m = mul_outlined(a, b);
        coop_put(sh, 0, role, lane, m);
    }
    __syncthreads();
    const F ppp = coop_get<F>(sh, 0, 0, lane), qq = coop_get<F>(sh, 0, 1, lane), zz3 = coop_get<F>(sh, 0, 2, lane);
    const F x3 = sub(sub(rr, ppp), dbl(qq));
    // round 4: t1 = r (qq - x3), t2 = s1 ppp, zzz3 = zzz12 ppp
    if (role == 0) { a = r; b = sub(qq, x3); }
    else if (role == 1) { a = s1; b = ppp; }
    else { a = zzz12; b = ppp; }
    if (role < 3) {   // warp 3 has no product in this round
        m = mul_outlined(a, b);
        coop_put(sh, 1, role, lane, m);
    }
    __syncthreads();
    const F t1 = coop_get<F>(sh, 1, 0, lane), t2 = coop_get<F>(sh, 1, 1, lane), zzz3 = coop_get<F>(sh, 1, 2, lane);
    // exceptional cases: per logical lane, identical in the 4 warps (replicated data)
    const bool generic = !q_id && !acc_id;
    const bool same_x = generic && p.is_zero();
    const bool need_double = same_x && r.is_zero();
    if (__syncthreads_or(need_double)) {   // P + P somewhere in the block: every thread runs the doubling, the lanes that need it keep it
        XYZZ<F> d = acc;
        coop4_double(d, sh);
        if (need_double) {
            acc = d;
            return;
        }
    }
    if (q_id) return;
    if (acc_id) {
        acc = q;
        return;
    }
    if (same_x) {   // P + (-P)
        acc = XYZZ<F>::identity();
        return;
    }
    acc.x = x3;
    acc.y = sub(t1, t2);
    acc.zz = zz3;
    acc.zzz = zzz3;
}

// value of logical lane (lane ^ mask) -- every warp shuffles its own replicated copy
template <class F>
SB_D XYZZ<F> coop_shfl_xor(const XYZZ<F>& v, int mask) {
    XYZZ<F> r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 32; i++) d[i] = __shfl_xor_sync(0xffffffffu, s[i], mask);
    return r;
}
// value of logical lane (lane + delta); lanes past the end get the identity
template <class F>
SB_D XYZZ<F> coop_shfl_down(const XYZZ<F>& v, int delta) {
    XYZZ<F> r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 32; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
    if ((int)(threadIdx.x & 31) + delta > 31) r = XYZZ<F>::identity();
    return r;
}
// sum over the 32 logical lanes; every lane ends with the total
template <class F>
SB_D XYZZ<F> coop_lane_sum(XYZZ<F> v, CoopBuf& sh) {
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ<F> t = coop_shfl_xor(v, d);
        coop4_add(v, t, sh);
    }
    return v;
}

}  // namespace sb
