// field.cuh -- 254-bit prime-field arithmetic for bn256 Fr / Fq on sm_100a.
//
// Replaces (on the device) the halo2curves field types the reference uses everywhere on its hot path
// (re-exported at reference src/lib.rs:24-27).  Same memory format as the Rust side hands over
// (SURVEY App. A): 4 x u64 little-endian limbs == 8 x u32 little-endian limbs, Montgomery form, R = 2^256.
//
// Two implementations of the Montgomery product live here:
//   * mul_portable(): plain C++ (32x32->64 products), compiles for host and device; it is the
//     on-device cross-check for the PTX path (tests/ run both on the GPU and compare bit for bit);
//   * mul(): on the device, an even/odd-limb interleaved CIOS written as PTX carry chains
//     (mad.lo.cc / madc.hi.cc), which ptxas fuses into IMAD.WIDE.U32(.X) -- the integer pipe is the
//     roofline of every kernel in this library, so this routine is the inner loop.
//
// Values are kept fully reduced in [0, p) at every function boundary.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#define SB_D __device__ __forceinline__
#else
#define SB_HD inline
#define SB_D inline
#endif

namespace sb {

enum : int { FIELD_FR = 0, FIELD_FQ = 1 };

// bn256 scalar field (= grumpkin base field)
struct FrParams {
    static constexpr int ID = FIELD_FR;
    static constexpr uint32_t P0 = 0xf0000001u, P1 = 0x43e1f593u, P2 = 0x79b97091u, P3 = 0x2833e848u,
                              P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
    static constexpr uint32_t INV = 0xefffffffu;  // -p^-1 mod 2^32
    // R mod p
    static constexpr uint32_t R0 = 0x4ffffffbu, R1 = 0xac96341cu, R2_ = 0x9f60cd29u, R3 = 0x36fc7695u,
                              R4 = 0x7879462eu, R5 = 0x666ea36fu, R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
    // R^2 mod p
    static constexpr uint32_t RR0 = 0xae216da7u, RR1 = 0x1bb8e645u, RR2 = 0xe35c59e3u, RR3 = 0x53fe3ab1u,
                              RR4 = 0x53bb8085u, RR5 = 0x8c49833du, RR6 = 0x7f4e44a5u, RR7 = 0x0216d0b1u;
    // R^3 mod p
    static constexpr uint32_t RRR0 = 0xb4bf0040u, RRR1 = 0x5e94d8e1u, RRR2 = 0x1cfbb6b8u, RRR3 = 0x2a489cbeu,
                              RRR4 = 0xa19fcfedu, RRR5 = 0x893cc664u, RRR6 = 0x7fcc657cu, RRR7 = 0x0cf8594bu;
};

// bn256 base field (= grumpkin scalar field)
struct FqParams {
    static constexpr int ID = FIELD_FQ;
    static constexpr uint32_t P0 = 0xd87cfd47u, P1 = 0x3c208c16u, P2 = 0x6871ca8du, P3 = 0x97816a91u,
                              P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
    static constexpr uint32_t INV = 0xe4866389u;
    static constexpr uint32_t R0 = 0xc58f0d9du, R1 = 0xd35d438du, R2_ = 0xf5c70b3du, R3 = 0x0a78eb28u,
                              R4 = 0x7879462cu, R5 = 0x666ea36fu, R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
    static constexpr uint32_t RR0 = 0x538afa89u, RR1 = 0xf32cfc5bu, RR2 = 0xd44501fbu, RR3 = 0xb5e71911u,
                              RR4 = 0x0a417ff6u, RR5 = 0x47ab1effu, RR6 = 0xcab8351fu, RR7 = 0x06d89f71u;
    static constexpr uint32_t RRR0 = 0xda1530dfu, RRR1 = 0xb1cd6dafu, RRR2 = 0xa7283db6u, RRR3 = 0x62f210e6u,
                              RRR4 = 0x0ada0afbu, RRR5 = 0xef7f0b0cu, RRR6 = 0x2d592544u, RRR7 = 0x20fd6e90u;
};

template <class P>
struct alignas(16) Fe {
    uint32_t v[8];

    using Params = P;

    static SB_HD Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    static SB_HD Fe one() {  // Montgomery form of 1
        Fe r;
        r.v[0] = P::R0; r.v[1] = P::R1; r.v[2] = P::R2_; r.v[3] = P::R3;
        r.v[4] = P::R4; r.v[5] = P::R5; r.v[6] = P::R6; r.v[7] = P::R7;
        return r;
    }
    static SB_HD Fe r_squared() {
        Fe r;
        r.v[0] = P::RR0; r.v[1] = P::RR1; r.v[2] = P::RR2; r.v[3] = P::RR3;
        r.v[4] = P::RR4; r.v[5] = P::RR5; r.v[6] = P::RR6; r.v[7] = P::RR7;
        return r;
    }
    static SB_HD Fe r_cubed() {
        Fe r;
        r.v[0] = P::RRR0; r.v[1] = P::RRR1; r.v[2] = P::RRR2; r.v[3] = P::RRR3;
        r.v[4] = P::RRR4; r.v[5] = P::RRR5; r.v[6] = P::RRR6; r.v[7] = P::RRR7;
        return r;
    }
    static SB_HD uint32_t modulus_limb(int i) {
        switch (i) {
            case 0: return P::P0; case 1: return P::P1; case 2: return P::P2; case 3: return P::P3;
            case 4: return P::P4; case 5: return P::P5; case 6: return P::P6; default: return P::P7;
        }
    }

    SB_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i];
        return o == 0;
    }
    SB_HD bool operator==(const Fe& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    SB_HD bool operator!=(const Fe& b) const { return !(*this == b); }
};

// ---------------------------------------------------------------------------------------------
// portable limb helpers (host + device)
// ---------------------------------------------------------------------------------------------

// r = a - p if a >= p else a   (a < 2p assumed, 8 limbs, no carry in)
template <class P>
SB_HD void reduce_once_portable(uint32_t r[8]) {
    uint32_t t[8];
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)r[i] - Fe<P>::modulus_limb(i) - br;
        t[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    if (!br) {
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = t[i];
    }
}

template <class P>
SB_HD Fe<P> add_portable(const Fe<P>& a, const Fe<P>& b) {
    Fe<P> r;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    reduce_once_portable<P>(r.v);  // a+b < 2p < 2^255
    return r;
}

template <class P>
SB_HD Fe<P> sub_portable(const Fe<P>& a, const Fe<P>& b) {
    Fe<P> r;
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        r.v[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    if (br) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)r.v[i] + Fe<P>::modulus_limb(i);
            r.v[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    return r;
}

// CIOS Montgomery product with 32-bit limbs and 64-bit accumulators.  REDUCE = false is the lazy form (see the
// "lazy domain" section below): inputs in [0, 2p), result in [0, 2p), no final subtraction.
template <class P, bool REDUCE = true>
SB_HD Fe<P> mul_portable(const Fe<P>& a, const Fe<P>& b) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.v[j] * b.v[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * P::INV;
        c = (uint64_t)m * Fe<P>::modulus_limb(0) + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * Fe<P>::modulus_limb(j) + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    Fe<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    if (REDUCE) reduce_once_portable<P>(r.v);  // result < 2p, t[8] == 0
    return r;
}

// ---------------------------------------------------------------------------------------------
// device PTX path
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)

// r = (r >= p) ? r - p : r
template <class P>
SB_D void reduce_once_ptx(uint32_t r[8]) {
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, br;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7), "=r"(br)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(P::P0), "r"(P::P1), "r"(P::P2), "r"(P::P3), "r"(P::P4), "r"(P::P5), "r"(P::P6), "r"(P::P7));
    if (br == 0) {  // no borrow: r >= p
        r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3; r[4] = t4; r[5] = t5; r[6] = t6; r[7] = t7;
    }
}

template <class P>
SB_D Fe<P> add_ptx(const Fe<P>& a, const Fe<P>& b) {
    Fe<P> r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    reduce_once_ptx<P>(r.v);
    return r;
}

template <class P>
SB_D Fe<P> sub_ptx(const Fe<P>& a, const Fe<P>& b) {
    Fe<P> r;
    uint32_t br;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(br)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // br is 0 or 0xffffffff: add (p & br)
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
        : "r"(P::P0 & br), "r"(P::P1 & br), "r"(P::P2 & br), "r"(P::P3 & br), "r"(P::P4 & br), "r"(P::P5 & br),
          "r"(P::P6 & br), "r"(P::P7 & br));
    return r;
}

// X[0..7] += (x0,x1,x2,x3) * s as four (lo,hi) pairs on one carry chain; carry out added into X[8].
SB_D void chain_even(uint32_t X[9], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t s) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(s));
}

// Y[0..7] += (x0..x3) * s, same pairing, no carry out (bounded by the caller's invariant).
SB_D void chain_odd(uint32_t Y[9], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t s) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(s));
}

// One frame shift of the interleaved CIOS (see mul_ptx): X[0] += Y[1] and, on the same carry chain,
// Y[k] = Y[k+2] + (x0..x3)*s pairs.  Y[8] is cleared (it becomes the next even array's carry limb).
SB_D void fold_shift_chain_odd(uint32_t X[9], uint32_t Y[9], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t s) {
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %10, %14, %3;\n\t"
        "madc.hi.cc.u32 %2, %10, %14, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %14, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %14, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %14, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %14, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %14, %9;\n\t"
        "madc.hi.u32 %8, %13, %14, 0;\n\t"
        "mov.u32 %9, 0;"
        : "+r"(X[0]), "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7]), "+r"(Y[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(s));
}

// Montgomery product, even/odd interleaved CIOS.
//
// The running sum S is held as S = E + O * 2^32 with E = e[0..8] (9 limbs) and O = o[0..7].  Products
// a[j]*b_i with even j land on aligned (lo,hi) limb pairs of E, odd j on aligned pairs of O, so each
// row is two independent carry chains with no limb-by-limb ripple.  After adding m*p (m = e[0]*INV)
// e[0] == 0 and S/2^32 = (e[1] + e[2..8] * 2^32) + O: the old O becomes the new E, e[2..8] becomes
// the new O, and the stray e[1] is folded into new-E limb 0 whose carry (weight 2^32) enters the new
// O chain as its carry-in.  Invariant: S < 2p at row boundaries, S < 2^288 inside a row, hence the
// "no carry out" claims of chain_odd / fold_shift_chain_odd.
template <class P, bool REDUCE = true>
SB_D Fe<P> mul_ptx(const Fe<P>& a, const Fe<P>& b) {
    uint32_t A[9], B[9];
    const uint32_t b0 = b.v[0];
    // row 0: E = A, O = B
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(A[0]), "=r"(A[1]), "=r"(A[2]), "=r"(A[3]), "=r"(A[4]), "=r"(A[5]), "=r"(A[6]), "=r"(A[7])
        : "r"(a.v[0]), "r"(a.v[2]), "r"(a.v[4]), "r"(a.v[6]), "r"(b0));
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(B[0]), "=r"(B[1]), "=r"(B[2]), "=r"(B[3]), "=r"(B[4]), "=r"(B[5]), "=r"(B[6]), "=r"(B[7])
        : "r"(a.v[1]), "r"(a.v[3]), "r"(a.v[5]), "r"(a.v[7]), "r"(b0));
    A[8] = 0;
    B[8] = 0;
    {
        uint32_t m = A[0] * P::INV;
        chain_odd(B, P::P1, P::P3, P::P5, P::P7, m);
        chain_even(A, P::P0, P::P2, P::P4, P::P6, m);
    }
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        {  // odd row: even array = B, odd array = A
            const uint32_t bi = b.v[i];
            fold_shift_chain_odd(B, A, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            chain_even(B, a.v[0], a.v[2], a.v[4], a.v[6], bi);
            uint32_t m = B[0] * P::INV;
            chain_odd(A, P::P1, P::P3, P::P5, P::P7, m);
            chain_even(B, P::P0, P::P2, P::P4, P::P6, m);
        }
        if (i + 1 < 8) {  // even row: even array = A, odd array = B
            const uint32_t bi = b.v[i + 1];
            fold_shift_chain_odd(A, B, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            chain_even(A, a.v[0], a.v[2], a.v[4], a.v[6], bi);
            uint32_t m = A[0] * P::INV;
            chain_odd(B, P::P1, P::P3, P::P5, P::P7, m);
            chain_even(A, P::P0, P::P2, P::P4, P::P6, m);
        }
    }
    // after row 7: even array = B (B[0] == 0), odd array = A.  S/2^32 = B[1..8] + A[0..7]
    Fe<P> r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]), "r"(B[8]),
          "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]));
    if (REDUCE) reduce_once_ptx<P>(r.v);
    return r;
}
#endif  // __CUDA_ARCH__

// ---------------------------------------------------------------------------------------------
// public dispatch
// ---------------------------------------------------------------------------------------------
template <class P>
SB_HD Fe<P> add(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__)
    return add_ptx(a, b);
#else
    return add_portable(a, b);
#endif
}
template <class P>
SB_HD Fe<P> sub(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__)
    return sub_ptx(a, b);
#else
    return sub_portable(a, b);
#endif
}
template <class P>
SB_HD Fe<P> mul(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__) && !defined(SB_FORCE_PORTABLE_MUL)
    return mul_ptx(a, b);
#else
    return mul_portable(a, b);
#endif
}
template <class P>
SB_HD Fe<P> sqr(const Fe<P>& a) { return mul(a, a); }
template <class P>
SB_HD Fe<P> dbl(const Fe<P>& a) { return add(a, a); }
template <class P>
SB_HD Fe<P> neg(const Fe<P>& a) {
    return a.is_zero() ? a : sub(Fe<P>::zero(), a);
}
template <class P>
SB_HD Fe<P> from_mont(const Fe<P>& a) {
    Fe<P> o = Fe<P>::zero();
    o.v[0] = 1;
    return mul(a, o);
}
template <class P>
SB_HD Fe<P> to_mont(const Fe<P>& a) { return mul(a, Fe<P>::r_squared()); }

// ---------------------------------------------------------------------------------------------
// lazy domain: values kept in [0, 2p) inside a hot loop, canonicalised only when they leave it
// ---------------------------------------------------------------------------------------------
// p < 2^254, so 4p < 2^256 = R.  The Montgomery product t = (a*b + m*p)/R of a, b < 2p satisfies
// t < 4p^2/R + p < 2p (4p/R = 0.756 for both bn256 fields) WITHOUT the final conditional subtraction, and inside
// the CIOS rows the running sum stays below 3p(1 + 2^-32) < 2^256, so the carry-chain invariants of mul_ptx hold
// unchanged.  That saves the 17-instruction compare-and-select tail of every product of the bucket kernel.
// add/sub/double stay closed on [0, 2p) with one conditional correction by 2p; zero is 0 or p.
template <class P>
struct TwoP {  // limbs of 2p
    static SB_HD uint32_t limb(int i) {
        const uint32_t lo = i ? Fe<P>::modulus_limb(i - 1) >> 31 : 0u;
        return (Fe<P>::modulus_limb(i) << 1) | lo;
    }
};

template <class P>
SB_HD Fe<P> sub_lazy_portable(const Fe<P>& a, const Fe<P>& b) {  // a - b, + 2p if negative
    Fe<P> r;
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - br;
        r.v[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    if (br) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)r.v[i] + TwoP<P>::limb(i);
            r.v[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    return r;
}
template <class P>
SB_HD Fe<P> dbl_lazy_portable(const Fe<P>& a) {  // 2a, - 2p if >= 2p   (2a < 4p < 2^256)
    Fe<P> r, t;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + a.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)r.v[i] - TwoP<P>::limb(i) - br;
        t.v[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    return br ? r : t;
}

template <class P>
SB_HD Fe<P> add_lazy_portable(const Fe<P>& a, const Fe<P>& b) {  // a + b, - 2p if >= 2p   (a + b < 4p < 2^256)
    Fe<P> r, t;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)r.v[i] - TwoP<P>::limb(i) - br;
        t.v[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    return br ? r : t;
}

#if defined(__CUDA_ARCH__)
template <class P>
SB_D Fe<P> add_lazy_ptx(const Fe<P>& a, const Fe<P>& b) {
    uint32_t s[8], t[8], br;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(br)
        : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]),
          "r"(TwoP<P>::limb(0)), "r"(TwoP<P>::limb(1)), "r"(TwoP<P>::limb(2)), "r"(TwoP<P>::limb(3)),
          "r"(TwoP<P>::limb(4)), "r"(TwoP<P>::limb(5)), "r"(TwoP<P>::limb(6)), "r"(TwoP<P>::limb(7)));
    Fe<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = br ? s[i] : t[i];
    return r;
}
template <class P>
SB_D Fe<P> sub_lazy_ptx(const Fe<P>& a, const Fe<P>& b) {
    Fe<P> r;
    uint32_t br;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(br)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
        : "r"(TwoP<P>::limb(0) & br), "r"(TwoP<P>::limb(1) & br), "r"(TwoP<P>::limb(2) & br), "r"(TwoP<P>::limb(3) & br),
          "r"(TwoP<P>::limb(4) & br), "r"(TwoP<P>::limb(5) & br), "r"(TwoP<P>::limb(6) & br), "r"(TwoP<P>::limb(7) & br));
    return r;
}
template <class P>
SB_D Fe<P> dbl_lazy_ptx(const Fe<P>& a) {
    uint32_t s[8], t[8], br;
    asm("add.cc.u32 %0, %8, %8;\n\t"
        "addc.cc.u32 %1, %9, %9;\n\t"
        "addc.cc.u32 %2, %10, %10;\n\t"
        "addc.cc.u32 %3, %11, %11;\n\t"
        "addc.cc.u32 %4, %12, %12;\n\t"
        "addc.cc.u32 %5, %13, %13;\n\t"
        "addc.cc.u32 %6, %14, %14;\n\t"
        "addc.u32 %7, %15, %15;"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(br)
        : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]),
          "r"(TwoP<P>::limb(0)), "r"(TwoP<P>::limb(1)), "r"(TwoP<P>::limb(2)), "r"(TwoP<P>::limb(3)),
          "r"(TwoP<P>::limb(4)), "r"(TwoP<P>::limb(5)), "r"(TwoP<P>::limb(6)), "r"(TwoP<P>::limb(7)));
    Fe<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = br ? s[i] : t[i];
    return r;
}
#endif

template <class P>
SB_HD Fe<P> mul_lazy(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__) && !defined(SB_FORCE_PORTABLE_MUL)
    return mul_ptx<P, false>(a, b);
#else
    return mul_portable<P, false>(a, b);
#endif
}
template <class P>
SB_HD Fe<P> sub_lazy(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__)
    return sub_lazy_ptx(a, b);
#else
    return sub_lazy_portable(a, b);
#endif
}
template <class P>
SB_HD Fe<P> add_lazy(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__)
    return add_lazy_ptx(a, b);
#else
    return add_lazy_portable(a, b);
#endif
}
template <class P>
SB_HD Fe<P> neg_lazy(const Fe<P>& a) {  // 2p - a, and 0 for a = 0 (2p itself is outside the domain)
    Fe<P> tp;
#pragma unroll
    for (int i = 0; i < 8; i++) tp.v[i] = TwoP<P>::limb(i);
    return a.is_zero() ? a : sub_lazy(tp, a);
}
template <class P>
SB_HD Fe<P> dbl_lazy(const Fe<P>& a) {
#if defined(__CUDA_ARCH__)
    return dbl_lazy_ptx(a);
#else
    return dbl_lazy_portable(a);
#endif
}
template <class P>
SB_HD bool is_zero_lazy(const Fe<P>& a) {  // a = 0 (mod p) for a in [0, 2p): a == 0 or a == p; the low limb filters
    if (a.v[0] != 0u && a.v[0] != P::P0) return false;
    uint32_t z = 0, q = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        z |= a.v[i];
        q |= a.v[i] ^ Fe<P>::modulus_limb(i);
    }
    return z == 0 || q == 0;
}
template <class P>
SB_HD Fe<P> canon(const Fe<P>& a) {  // [0, 2p) -> [0, p)
    Fe<P> r = a;
    reduce_once_portable<P>(r.v);
    return r;
}

// ---- plain-integer helpers for the binary inversion --------------------------------------------------
SB_HD bool limbs_geq(const uint32_t a[8], const uint32_t b[8]) {
    for (int i = 7; i >= 0; i--) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
SB_HD void limbs_sub(uint32_t a[8], const uint32_t b[8]) {  // a -= b (a >= b)
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        a[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
}
SB_HD void limbs_shr1(uint32_t a[8], uint32_t top_in) {  // a = (top_in:a) >> 1
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] = (a[7] >> 1) | (top_in << 31);
}
// x = x/2 mod p  (x < p)
template <class P>
SB_HD void limbs_half_mod(uint32_t x[8]) {
    uint32_t carry = 0;
    if (x[0] & 1) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)x[i] + Fe<P>::modulus_limb(i);
            x[i] = (uint32_t)c;
            c >>= 32;
        }
        carry = (uint32_t)c;
    }
    limbs_shr1(x, carry);
}
// x = x - y mod p  (x, y < p)
template <class P>
SB_HD void limbs_sub_mod(uint32_t x[8], const uint32_t y[8]) {
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)x[i] - y[i] - br;
        x[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    if (br) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)x[i] + Fe<P>::modulus_limb(i);
            x[i] = (uint32_t)c;
            c >>= 32;
        }
    }
}

// Inverse by the binary extended Euclid on the stored integer (a != 0).  `a` holds A*R; the loop returns
// (A*R)^-1 as a plain integer and one Montgomery product by R^3 turns it into A^-1 * R.  About 6x shorter
// dependent chain than the Fermat ladder below -- it sits on the latency-bound tail of every commitment.
template <class P>
SB_HD Fe<P> inv_binary(const Fe<P>& a) {
    uint32_t u[8], v[8], x1[8], x2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u[i] = a.v[i];
        v[i] = Fe<P>::modulus_limb(i);
        x1[i] = 0;
        x2[i] = 0;
    }
    x1[0] = 1;
    auto is_one = [](const uint32_t* t) {
        uint32_t o = t[0] ^ 1u;
        for (int i = 1; i < 8; i++) o |= t[i];
        return o == 0;
    };
    while (!is_one(u) && !is_one(v)) {
        while (!(u[0] & 1)) {
            limbs_shr1(u, 0);
            limbs_half_mod<P>(x1);
        }
        while (!(v[0] & 1)) {
            limbs_shr1(v, 0);
            limbs_half_mod<P>(x2);
        }
        if (limbs_geq(u, v)) {
            limbs_sub(u, v);
            limbs_sub_mod<P>(x1, x2);
        } else {
            limbs_sub(v, u);
            limbs_sub_mod<P>(x2, x1);
        }
    }
    Fe<P> r;
    const uint32_t* src = is_one(u) ? x1 : x2;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = src[i];
    return mul(r, Fe<P>::r_cubed());
}

// ---- inverse by batched divsteps ("safegcd", Bernstein-Yang 2019, in the 30-bit-limb variable-time form that
// libsecp256k1's modinv32_var popularised) ------------------------------------------------------------------
// f, g (the gcd pair, starting at p and the input) and d, e (the Bezout coefficients mod p) live in nine signed
// 30-bit limbs.  Each round runs 30 divsteps on the low 30 bits of f, g only -- stripping trailing zeros of g with
// ctz and cancelling up to 8 low bits at once -- while collecting them in a 2x2 integer matrix t with entries below
// 2^30, and then applies t to the full (f, g) and, modulo p with one exact division by 2^30, to (d, e).  A 254-bit
// inverse takes ~18-22 rounds of ~150 multiply-adds instead of ~380 dependent whole-number shift/subtract rounds.
SB_HD int sg_ctz32(uint32_t x) {  // x != 0
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

struct SgMatrix {
    int32_t u, v, q, r;
};

// 30 divsteps on the low limbs; eta = -delta.  Returns the new eta; t maps (f, g) -> 2^30 * (f', g').
SB_HD int32_t sg_divsteps_30(int32_t eta, uint32_t f, uint32_t g, SgMatrix& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;
    int i = 30;
    for (;;) {
        const int zeros = sg_ctz32(g | (0xFFFFFFFFu << i));  // sentinel: never count past the i steps left
        g >>= zeros;
        u <<= zeros;
        v <<= zeros;
        eta -= zeros;
        i -= zeros;
        if (i == 0) break;
        if (eta < 0) {  // swap: (f, g) <- (g, -f)
            eta = -eta;
            uint32_t tmp = f; f = g; g = 0u - tmp;
            tmp = u; u = q; q = 0u - tmp;
            tmp = v; v = r; r = 0u - tmp;
        }
        // cancel the low min(eta + 1, i, 8) bits of g with a multiple of f (f is odd)
        const int limit = (eta + 1) > i ? i : (eta + 1);
        const uint32_t m = (0xFFFFFFFFu >> (32 - limit)) & 255u;
        uint32_t finv = f;                 // f * f = 1 mod 8
        finv *= 2u - f * finv;             // mod 2^6
        finv *= 2u - f * finv;             // mod 2^12
        const uint32_t w = (g * (0u - finv)) & m;
        g += f * w;
        q += u * w;
        r += v * w;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return eta;
}

// (f, g) <- t * (f, g) / 2^30   (exact)
SB_HD void sg_update_fg(int32_t f[9], int32_t g[9], const SgMatrix& t) {
    const int32_t M30 = 0x3FFFFFFF;
    int64_t cf = (int64_t)t.u * f[0] + (int64_t)t.v * g[0];
    int64_t cg = (int64_t)t.q * f[0] + (int64_t)t.r * g[0];
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cf += (int64_t)t.u * f[i] + (int64_t)t.v * g[i];
        cg += (int64_t)t.q * f[i] + (int64_t)t.r * g[i];
        f[i - 1] = (int32_t)cf & M30;
        g[i - 1] = (int32_t)cg & M30;
        cf >>= 30;
        cg >>= 30;
    }
    f[8] = (int32_t)cf;
    g[8] = (int32_t)cg;
}

// (d, e) <- t * (d, e) / 2^30 mod p, kept in (-2p, p): multiples of p make the low 30 bits vanish first
SB_HD void sg_update_de(int32_t d[9], int32_t e[9], const SgMatrix& t, const int32_t mod[9], uint32_t mod_inv30) {
    const int32_t M30 = 0x3FFFFFFF;
    const int32_t sd = d[8] >> 31, se = e[8] >> 31;
    int32_t md = (t.u & sd) + (t.v & se);
    int32_t me = (t.q & sd) + (t.r & se);
    int64_t cd = (int64_t)t.u * d[0] + (int64_t)t.v * e[0];
    int64_t ce = (int64_t)t.q * d[0] + (int64_t)t.r * e[0];
    md -= (int32_t)((mod_inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((mod_inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)mod[0] * md;
    ce += (int64_t)mod[0] * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cd += (int64_t)t.u * d[i] + (int64_t)t.v * e[i] + (int64_t)mod[i] * md;
        ce += (int64_t)t.q * d[i] + (int64_t)t.r * e[i] + (int64_t)mod[i] * me;
        d[i - 1] = (int32_t)cd & M30;
        e[i - 1] = (int32_t)ce & M30;
        cd >>= 30;
        ce >>= 30;
    }
    d[8] = (int32_t)cd;
    e[8] = (int32_t)ce;
}

SB_HD void sg_pack30(const uint32_t w[8], int32_t o[9]) {  // 8 x 32 bits -> 9 x 30 bits
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, lo = bit >> 5, sh = bit & 31;
        uint64_t x = (uint64_t)w[lo] >> sh;
        if (lo + 1 < 8 && sh > 2) x |= (uint64_t)w[lo + 1] << (32 - sh);
        o[i] = (int32_t)((uint32_t)x & 0x3FFFFFFFu);
    }
}

// Inverse of the stored integer a (Montgomery form in, Montgomery form out); zero maps to zero.
template <class P>
SB_HD Fe<P> inv_safegcd(const Fe<P>& a) {
    const int32_t M30 = 0x3FFFFFFF;
    uint32_t pw[8];
#pragma unroll
    for (int i = 0; i < 8; i++) pw[i] = Fe<P>::modulus_limb(i);
    int32_t mod[9], f[9], g[9], d[9], e[9];
    sg_pack30(pw, mod);
    sg_pack30(a.v, g);
    uint32_t minv = pw[0];  // p^-1 mod 2^30 by Newton (p * p = 1 mod 8)
    minv *= 2u - pw[0] * minv;
    minv *= 2u - pw[0] * minv;
    minv *= 2u - pw[0] * minv;
    minv *= 2u - pw[0] * minv;
    minv &= (uint32_t)M30;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        f[i] = mod[i];
        d[i] = 0;
        e[i] = 0;
    }
    e[0] = 1;
    int32_t eta = -1;
    for (int round = 0; round < 40; round++) {  // 254-bit inputs finish in about 20; the cap only guards bad input
        SgMatrix t;
        eta = sg_divsteps_30(eta, (uint32_t)f[0], (uint32_t)g[0], t);
        sg_update_de(d, e, t, mod, minv);
        sg_update_fg(f, g, t);
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) nz |= g[i];
        if (nz == 0) break;
    }
    // g = 0, f = +-gcd = +-1 (or +-p for a = 0, where d = 0); d = +-a^-1 in (-2p, p): fix the sign and the range
    const int32_t neg = f[8] >> 31;
    int32_t add = d[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        d[i] += mod[i] & add;
        d[i] = (d[i] ^ neg) - neg;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        d[i + 1] += d[i] >> 30;
        d[i] &= M30;
    }
    add = d[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) d[i] += mod[i] & add;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        d[i + 1] += d[i] >> 30;
        d[i] &= M30;
    }
    Fe<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) {  // 9 x 30 bits -> 8 x 32 bits
        const int bit = 32 * i, lo = bit / 30, sh = bit % 30;
        uint64_t x = (uint64_t)(uint32_t)d[lo] >> sh;
        x |= (uint64_t)(uint32_t)d[lo + 1] << (30 - sh);
        if (lo + 2 < 9) x |= (uint64_t)(uint32_t)d[lo + 2] << (60 - sh);
        r.v[i] = (uint32_t)x;
    }
    return mul(r, Fe<P>::r_cubed());  // (A R)^-1 * R^3 / R = A^-1 R
}

// a^(p-2) by square-and-multiply over the constant exponent (a != 0).
template <class P>
SB_HD Fe<P> inv(const Fe<P>& a) {
    Fe<P> acc = Fe<P>::one();
    // exponent p-2: only limb 0 differs from p (p is odd, P0 >= 2)
    for (int i = 7; i >= 0; i--) {
        uint32_t e = Fe<P>::modulus_limb(i) - (i == 0 ? 2u : 0u);
        for (int bit = 31; bit >= 0; bit--) {
            acc = sqr(acc);
            if ((e >> bit) & 1) acc = mul(acc, a);
        }
    }
    return acc;
}

using Fr = Fe<FrParams>;
using Fq = Fe<FqParams>;

}  // namespace sb
