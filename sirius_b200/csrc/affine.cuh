// affine.cuh -- batched-affine pair additions for the MSM bucket sums (msm.cu, k_pair_round).
//
// The bucket accumulation of `CommitmentKey::commit` (reference src/commitment.rs:81-90) is n*W point additions.
// In XYZZ coordinates one mixed addition costs 8M + 2S; in affine coordinates it costs one inversion + 2M + 1S, and
// Montgomery's trick shares ONE inversion among all the additions of a thread block (3 more products each):
// 6 products per addition instead of 10.  Affine additions have no accumulator to chain through, so the sorted
// (bucket-major) entries are reduced by ROUNDS: round r adds the entries of every bucket pairwise,
//     A_{r+1}[off_{r+1}(b) + j] = A_r[off_r(b) + 2j] + A_r[off_r(b) + 2j + 1]      (an odd last entry is copied),
// halving every bucket; cnt_{r+1}(b) = ceil(cnt_r(b) / 2).  After a few rounds (they do 1/2, 1/4, 1/8 .. of all
// additions) the remaining short runs are finished by the XYZZ chunk kernel (k_accumulate, direct-source form).
// All group operations are exact, so the regrouping does not change the commitment.
//
// Everything in this header is host-compilable: the per-thread phases and the block's product tree are plain
// functions of (thread id, arrays), which tests/host/host_affine.cpp runs thread by thread on the CPU against the
// XYZZ law; the kernel in msm.cu only adds __syncthreads() between the same calls.
#pragma once
#include "curve.cuh"

namespace sb {

constexpr int PR_THREADS = 256;  // threads per block: one shared inversion per block
constexpr int PR_LEVELS = 8;     // log2(PR_THREADS)
constexpr int PR_NODES = 2 * PR_THREADS - 1;

// what an output slot does
enum : uint8_t {
    PR_NONE = 0,   // beyond the round's last output
    PR_COPY1 = 1,  // out = first operand (no partner, or the partner is the identity)
    PR_COPY2 = 2,  // out = second operand (the first is the identity)
    PR_IDENT = 3,  // out = identity (P + (-P), or doubling a point with y = 0)
    PR_ADD = 4,    // generic chord: denominator x2 - x1
    PR_DBL = 5,    // tangent: denominator 2*y1, numerator 3*x1^2
    PR_FIRST = 8   // flag: first denominator of this thread (its running prefix is 1, no product needed)
};

#if defined(__CUDA_ARCH__)
template <class T>
SB_D T pr_ld(const T* p) {  // read-only 128-bit loads
    static_assert(sizeof(T) % 16 == 0, "16-byte multiples only");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
    return r;
}
SB_D uint32_t pr_ld_u32(const uint32_t* p) { return __ldg(p); }
template <class T>
SB_D void pr_st(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}
#else
template <class T>
inline T pr_ld(const T* p) { return *p; }
inline uint32_t pr_ld_u32(const uint32_t* p) { return *p; }
template <class T>
inline void pr_st(T* p, const T& v) { *p = v; }
#endif

// Operand source of a round.  INDEXED (round 0): position -> sorted entry (table index | sign << 31) -> window
// table; otherwise the previous round's output points.
template <class F, bool INDEXED>
struct PairSrc {
    const Affine<F>* pts;
    const uint32_t* eidx;
    SB_HD F load_x(uint32_t pos) const {
        if (INDEXED) return pr_ld(&pts[pr_ld_u32(eidx + pos) & 0x7fffffffu].x);
        return pr_ld(&pts[pos].x);
    }
    SB_HD F load_y(uint32_t pos) const {
        if (INDEXED) {
            const uint32_t e = pr_ld_u32(eidx + pos);
            const F y = pr_ld(&pts[e & 0x7fffffffu].y);
            return (e >> 31) ? neg(y) : y;
        }
        return pr_ld(&pts[pos].y);
    }
};

// Forward phase of thread g: outputs [g*B, (g+1)*B) of the round.  Fills pos[] (first operand's input position),
// kind[] and cp[] (product of this thread's earlier denominators, valid where the slot has one and is not FIRST);
// returns the product of all its denominators (Montgomery 1 if it has none).
template <class F, bool INDEXED, int B>
SB_HD F pair_forward(const PairSrc<F, INDEXED>& src, const uint32_t* off_in, const uint32_t* off_out, uint32_t KB, uint32_t g,
                     uint32_t* pos, uint8_t* kind, F* cp) {
    const uint32_t total_out = pr_ld_u32(off_out + KB);
    const uint64_t o0 = (uint64_t)g * (uint64_t)B;
    F run = F::one();
    bool have = false;
    if (o0 >= total_out) {
        for (int j = 0; j < B; j++) kind[j] = PR_NONE;
        return run;
    }
    // bucket of the first output: off_out[b] <= o0 < off_out[b + 1]
    uint32_t lo = 0, hi = KB - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pr_ld_u32(off_out + mid + 1) > (uint32_t)o0) hi = mid;
        else lo = mid + 1;
    }
    uint32_t b = lo;
    uint32_t start_out = pr_ld_u32(off_out + b), end_out = pr_ld_u32(off_out + b + 1);
    uint32_t in_base = pr_ld_u32(off_in + b), cnt_in = pr_ld_u32(off_in + b + 1) - in_base;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j < B; j++) {
        const uint32_t o = (uint32_t)o0 + (uint32_t)j;
        if (o >= total_out) {
            kind[j] = PR_NONE;
            continue;
        }
        if (o >= end_out) {
            do {
                b++;
                start_out = end_out;
                end_out = pr_ld_u32(off_out + b + 1);
            } while (o >= end_out);
            in_base = pr_ld_u32(off_in + b);
            cnt_in = pr_ld_u32(off_in + b + 1) - in_base;
        }
        const uint32_t jj = o - start_out;
        const uint32_t p1 = in_base + 2u * jj;
        pos[j] = p1;
        if (2u * jj + 1u >= cnt_in) {
            kind[j] = PR_COPY1;
            continue;
        }
        const F x1 = src.load_x(p1), x2 = src.load_x(p1 + 1);
        F den = sub(x2, x1);
        uint8_t k = PR_ADD;
        if (x1.is_zero() || x2.is_zero() || den.is_zero()) {  // rare: identity operands, P + P, P + (-P)
            const F y1 = src.load_y(p1), y2 = src.load_y(p1 + 1);
            const bool id1 = x1.is_zero() && y1.is_zero(), id2 = x2.is_zero() && y2.is_zero();
            if (id1) k = id2 ? PR_IDENT : PR_COPY2;
            else if (id2) k = PR_COPY1;
            else if (!den.is_zero()) k = PR_ADD;
            else if (y1 == y2 && !y1.is_zero()) {
                k = PR_DBL;
                den = dbl(y1);
            } else k = PR_IDENT;
        }
        if (k == PR_ADD || k == PR_DBL) {
            if (!have) {
                k |= PR_FIRST;
                run = den;
                have = true;
            } else {
                cp[j] = run;
                run = mul(run, den);
            }
        }
        kind[j] = k;
    }
    return run;
}

// Backward phase: `inv_total` = 1 / (product of this thread's denominators); writes the thread's outputs.
template <class F, bool INDEXED, int B>
SB_HD void pair_backward(const PairSrc<F, INDEXED>& src, const uint32_t* pos, const uint8_t* kind, const F* cp, F inv_total,
                         Affine<F>* out /* &A_{r+1}[g*B] */) {
    F I = inv_total;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = B - 1; j >= 0; j--) {
        const uint8_t kf = kind[j];
        const uint8_t k = kf & 7;
        if (k == PR_NONE) continue;
        Affine<F> r;
        if (k == PR_IDENT) {
            r.x = F::zero();
            r.y = F::zero();
        } else if (k == PR_COPY1 || k == PR_COPY2) {
            const uint32_t p = pos[j] + (k == PR_COPY2 ? 1u : 0u);
            r.x = src.load_x(p);
            r.y = src.load_y(p);
        } else {
            const uint32_t p1 = pos[j];
            const F x1 = src.load_x(p1), y1 = src.load_y(p1);
            F x2, num, den;
            if (k == PR_ADD) {
                x2 = src.load_x(p1 + 1);
                den = sub(x2, x1);
                num = sub(src.load_y(p1 + 1), y1);
            } else {  // PR_DBL
                x2 = x1;
                den = dbl(y1);
                const F xx = sqr(x1);
                num = add(dbl(xx), xx);
            }
            F inv_den;
            if (kf & PR_FIRST) inv_den = I;  // nothing before it: I is already 1/den
            else {
                inv_den = mul(I, cp[j]);
                I = mul(I, den);
            }
            const F lam = mul(num, inv_den);
            r.x = sub(sub(sqr(lam), x1), x2);
            r.y = sub(mul(lam, sub(x1, r.x)), y1);
        }
        pr_st(out + j, r);
    }
}

// ---- the block's product tree over the 256 thread totals (Montgomery's trick, tree form) ------------------------
// node[] holds level 0 (256 thread totals) .. level 8 (the block total) back to back; ninv[] the matching inverses.
SB_HD int pr_level_off(int l) { return 2 * PR_THREADS - ((2 * PR_THREADS) >> l); }

template <class F>
SB_HD void pr_tree_up(F* node, int l, int i) {  // node[l+1][i] = node[l][2i] * node[l][2i+1],  i < PR_THREADS >> (l+1)
    const F* src = node + pr_level_off(l);
    node[pr_level_off(l + 1) + i] = mul(src[2 * i], src[2 * i + 1]);
}
template <class F>
SB_HD void pr_tree_down(const F* node, F* ninv, int l, int i) {  // ninv[l][i] = ninv[l+1][i/2] * node[l][i^1],  i < PR_THREADS >> l
    ninv[pr_level_off(l) + i] = mul(ninv[pr_level_off(l + 1) + (i >> 1)], node[pr_level_off(l) + (i ^ 1)]);
}

}  // namespace sb
