// microbench.cu -- ground-truth latency / throughput of the field and group primitives on the device.
// Used from tools/microbench.py; results go into DESIGN.md (integer-pipe roofline of the MSM).
#include <vector>
#include "common.cuh"
#include "curve.cuh"
#include "quad.cuh"
#include "coop.cuh"

namespace sb {

template <class F>
__global__ void k_mb_mul_chain(F* io, int iters) {  // dependent products: latency (1 warp) or throughput (full grid)
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F x = io[2 * i], y = io[2 * i + 1];
    for (int k = 0; k < iters; k++) {
        x = mul(x, y);
        y = mul(y, x);
    }
    io[2 * i] = add(x, y);
}
template <class F>
__global__ void k_mb_mul_chain4(F* io, int iters) {  // four independent chains per thread (ILP)
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F x = io[2 * i], y = io[2 * i + 1], z = add(x, y), w = sub(x, y);
    for (int k = 0; k < iters; k++) {
        x = mul(x, x);
        y = mul(y, y);
        z = mul(z, z);
        w = mul(w, w);
    }
    io[2 * i] = add(add(x, y), add(z, w));
}
template <class F>
__global__ void k_mb_mul_outlined(F* io, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F x = io[2 * i], y = io[2 * i + 1];
    for (int k = 0; k < iters; k++) {
        x = mul_outlined(x, y);
        y = mul_outlined(y, x);
    }
    io[2 * i] = add(x, y);
}
template <class F>
__global__ void k_mb_add_call(XYZZ<F>* io, int iters) {  // serial full additions through the out-of-line entry
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> a = io[2 * i], b = io[2 * i + 1];
    for (int k = 0; k < iters; k++) xyzz_add_call(a, b);
    io[2 * i] = a;
}
template <class F>
__global__ void k_mb_add_inline(XYZZ<F>* io, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> a = io[2 * i], b = io[2 * i + 1];
    for (int k = 0; k < iters; k++) xyzz_add<true>(a, b);
    io[2 * i] = a;
}
template <class F>
__global__ void k_mb_madd(XYZZ<F>* io, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> a = io[2 * i];
    Affine<F> q;
    q.x = io[2 * i + 1].x;
    q.y = io[2 * i + 1].y;
    for (int k = 0; k < iters; k++) xyzz_madd(a, q, (k & 1) != 0);
    io[2 * i] = a;
}
template <class F>
__global__ void k_mb_madd_lazy(XYZZ<F>* io, int iters) {  // the bucket kernel's addition (lazy domain, field.cuh)
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> a = io[2 * i];
    Affine<F> q;
    q.x = io[2 * i + 1].x;
    q.y = io[2 * i + 1].y;
    for (int k = 0; k < iters; k++) xyzz_madd_lazy(a, q, (k & 1) != 0);
    io[2 * i] = canon_point(a);
}
// serial additions / doublings by the 4 warps of a block (coop.cuh): block = 128 threads = 32 logical lanes
template <class F>
__global__ void __launch_bounds__(COOP_THREADS) k_mb_coop_add(XYZZ<F>* io, int iters) {
    __shared__ CoopBuf sh;
    const size_t i = (size_t)blockIdx.x * 32 + (threadIdx.x & 31);
    XYZZ<F> a = io[2 * i], b = io[2 * i + 1];
    for (int k = 0; k < iters; k++) coop4_add(a, b, sh);
    if (threadIdx.x < 32) io[2 * i] = a;
}
template <class F>
__global__ void __launch_bounds__(COOP_THREADS) k_mb_coop_double(XYZZ<F>* io, int iters) {
    __shared__ CoopBuf sh;
    const size_t i = (size_t)blockIdx.x * 32 + (threadIdx.x & 31);
    XYZZ<F> a = io[2 * i];
    for (int k = 0; k < iters; k++) coop4_double(a, sh);
    if (threadIdx.x < 32) io[2 * i] = a;
}
// Timing experiment only (wrong results): the product with 16 of its 128 IMAD.WIDE removed and ~100 extra
// carry-chain additions, to price a Karatsuba product (48 + 64 wide multiplies + more additions) before writing it.
template <class P>
SB_D Fe<P> mul_fake112(const Fe<P>& a, const Fe<P>& b) {
    uint32_t A[9], B[9];
    const uint32_t b0 = b.v[0];
#pragma unroll
    for (int k = 0; k < 9; k++) { A[k] = a.v[k & 7] ^ b0; B[k] = b.v[k & 7] + k; }
    {
        uint32_t m = A[0] * P::INV;
        chain_odd(B, P::P1, P::P3, P::P5, P::P7, m);
        chain_even(A, P::P0, P::P2, P::P4, P::P6, m);
    }
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        {
            const uint32_t bi = b.v[i];
            fold_shift_chain_odd(B, A, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            if (i < 4) chain_even(B, a.v[0], a.v[2], a.v[4], a.v[6], bi);
            uint32_t m = B[0] * P::INV;
            chain_odd(A, P::P1, P::P3, P::P5, P::P7, m);
            chain_even(B, P::P0, P::P2, P::P4, P::P6, m);
        }
        if (i + 1 < 8) {
            const uint32_t bi = b.v[i + 1];
            fold_shift_chain_odd(A, B, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            if (i < 4) chain_even(A, a.v[0], a.v[2], a.v[4], a.v[6], bi);
            uint32_t m = A[0] * P::INV;
            chain_odd(B, P::P1, P::P3, P::P5, P::P7, m);
            chain_even(A, P::P0, P::P2, P::P4, P::P6, m);
        }
    }
    // ~100 extra additions on the ALU pipe (12 dependent 8-limb carry chains)
    Fe<P> r, t;
#pragma unroll
    for (int k = 0; k < 8; k++) { r.v[k] = B[k + 1] + A[k]; t.v[k] = A[k + 1]; }
#pragma unroll
    for (int rep = 0; rep < 6; rep++) {
        r = add_ptx(r, t);
        t = sub_ptx(t, r);
    }
    return r;
}
template <class F>
__global__ void k_mb_mul_fake(F* io, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F x = io[2 * i], y = io[2 * i + 1];
    for (int k = 0; k < iters; k++) {
        x = mul_fake112(x, y);
        y = mul_fake112(y, x);
    }
    io[2 * i] = add(x, y);
}

template <class F>
__global__ void k_mb_quad_add(XYZZ<F>* io, int iters) {  // serial quad-lane additions (operands replicated per group)
    size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    XYZZ<F> a = io[2 * g], b = io[2 * g + 1];
    for (int k = 0; k < iters; k++) quad_add(a, b);
    if ((threadIdx.x & 3) == 0) io[2 * g] = a;
}
template <class F>
__global__ void k_mb_quad_double(XYZZ<F>* io, int iters) {
    size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    XYZZ<F> a = io[2 * g];
    for (int k = 0; k < iters; k++) quad_double(a);
    if ((threadIdx.x & 3) == 0) io[2 * g] = a;
}
template <class F>
__global__ void k_mb_double_call(XYZZ<F>* io, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> a = io[2 * i];
    for (int k = 0; k < iters; k++) xyzz_double_call(a);
    io[2 * i] = a;
}
template <class F>
__global__ void k_mb_inv(F* io, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F a = io[2 * i];
    for (int k = 0; k < iters; k++) a = inv_binary(a);
    io[2 * i] = a;
}

template <class F>
__global__ void k_mb_inv_safegcd(F* io, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F a = io[2 * i];
    for (int k = 0; k < iters; k++) a = inv_safegcd(a);
    io[2 * i] = a;
}

__global__ void k_mb_imad_wide(uint32_t* io, int iters) {  // raw IMAD.WIDE.U32 issue rate, 8 independent accumulators
    uint32_t a = io[threadIdx.x], b = io[threadIdx.x + 32];
    unsigned long long acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = j;
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = (unsigned long long)a * (b + j) + acc[j];
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += acc[j];
    io[threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}


// ---- raw pipe rates (inline PTX, volatile: nothing is folded away).  Every thread runs `iters` rounds of 8
// independent dependency chains of one instruction kind; tools/microbench5.py turns the times into cycles per
// warp-instruction per scheduler.  which: 20 mad.wide.u32 (IMAD.WIDE), 21 mad.lo.cc/madc.hi.cc pairs (the
// IMAD.WIDE.X carry chain of mul_ptx), 22 mad.lo.u32 (IMAD), 23 add.cc/addc (IADD3.X), 24 fma.rz.f64 (DFMA),
// 25 DFMA + 64-bit integer add interleaved 3:2 (the FP64 limb-product recipe), 26 IMAD.WIDE and DFMA
// interleaved 1:1 in one warp, 27 even warps IMAD.WIDE / odd warps DFMA, 28 even warps carry-chain IMAD / odd warps
// the 3:2 DFMA+add recipe, 29 add.u64.
template <int KIND>
__global__ void k_mb_pipe(uint32_t* io, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = io[tid & 1023] | 1u, b = io[(tid + 7) & 1023] | 3u;
    unsigned long long acc[8];
    double dacc[8];
    uint32_t r[16];
    const double da = 4503599627370497.0 + (double)(a & 0xffff), db = 3.0 + (double)(b & 0xff);
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j] = (unsigned long long)j * 0x9e3779b97f4a7c15ull + a; dacc[j] = (double)j + 0.5; }
#pragma unroll
    for (int j = 0; j < 16; j++) r[j] = a * (j + 1) + b;
    int kind = KIND;
    if (KIND == 27) kind = ((threadIdx.x >> 5) & 1) ? 24 : 20;
    if (KIND == 28) kind = ((threadIdx.x >> 5) & 1) ? 25 : 21;
#pragma unroll 1
    for (int k = 0; k < iters; k++) {
        // every chain feeds its own result back as an operand, so ptxas cannot hoist or merge anything
        if (kind == 20) {
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = (unsigned long long)(uint32_t)acc[j] * b + acc[j];
        } else if (kind == 21) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t* q = r + 8 * h;
                const uint32_t x0 = q[1], x1 = q[3], x2 = q[5], x3 = q[7], s2 = q[0] | 1u;
                asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
                             "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                             "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
                             "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                             "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
                             "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                             "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
                             "madc.hi.u32 %7, %11, %12, %7;"
                             : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7])
                             : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(s2));
            }
        } else if (kind == 22) {
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = r[j] * b + r[j + 8];
        } else if (kind == 23) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t* q = r + 8 * h;
                asm volatile("add.cc.u32 %0, %0, %1;\n\t"
                             "addc.cc.u32 %1, %1, %2;\n\t"
                             "addc.cc.u32 %2, %2, %3;\n\t"
                             "addc.cc.u32 %3, %3, %4;\n\t"
                             "addc.cc.u32 %4, %4, %5;\n\t"
                             "addc.cc.u32 %5, %5, %6;\n\t"
                             "addc.cc.u32 %6, %6, %7;\n\t"
                             "addc.u32 %7, %7, %0;"
                             : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]));
            }
        } else if (kind == 24) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(dacc[j]) : "d"(db), "d"(da));
        } else if (kind == 25) {  // per limb product: 2 DFMA + 1 DADD on the FP64 pipe, two 64-bit integer adds
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double hi, lo, sub;
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(dacc[j]), "d"(db), "d"(da));
                asm volatile("sub.rz.f64 %0, %1, %2;" : "=d"(sub) : "d"(da), "d"(hi));
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(lo) : "d"(dacc[j]), "d"(db), "d"(sub));
                acc[2 * j] += (unsigned long long)__double_as_longlong(hi);
                acc[2 * j + 1] += (unsigned long long)__double_as_longlong(lo);
                dacc[j] = lo;
            }
        } else if (kind == 26) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                acc[j] = (unsigned long long)(uint32_t)acc[j] * b + acc[j];
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(dacc[j]) : "d"(db), "d"(da));
            }
        } else if (kind == 29) {
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] += acc[(j + 1) & 7] | 1ull;
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += acc[j] + (unsigned long long)__double_as_longlong(dacc[j]);
#pragma unroll
    for (int j = 0; j < 16; j++) s += r[j];
    if (iters == 0x7fffffff) io[tid & 1023] = (uint32_t)s ^ (uint32_t)(s >> 32);  // never true; keeps the chains live
}

}  // namespace sb

using namespace sb;

extern "C" int sb_microbench(int which, int iters, int blocks, int threads, double* out_ms) {
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    size_t n = (size_t)blocks * threads;
    char* d = nullptr;
    SB_CUDA_TRY(cudaMalloc(&d, n * 256 + 1024));
    // inputs: a valid-looking nonzero pattern (values < p: top limb small)
    std::vector<uint32_t> h(n * 64 + 256);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)(i * 2654435761u + 12345u) & ((i % 8 == 7) ? 0x0fffffffu : 0xffffffffu);
    SB_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0, rt.stream);
        switch (which) {
            case 0: k_mb_mul_chain<Fq><<<blocks, threads, 0, rt.stream>>>((Fq*)d, iters); break;
            case 1: k_mb_mul_chain4<Fq><<<blocks, threads, 0, rt.stream>>>((Fq*)d, iters); break;
            case 2: k_mb_mul_outlined<Fq><<<blocks, threads, 0, rt.stream>>>((Fq*)d, iters); break;
            case 3: k_mb_add_call<Fq><<<blocks, threads, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 4: k_mb_add_inline<Fq><<<blocks, threads, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 5: k_mb_madd<Fq><<<blocks, threads, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 6: k_mb_imad_wide<<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 7: k_mb_mul_fake<Fq><<<blocks, threads, 0, rt.stream>>>((Fq*)d, iters); break;
            case 8: k_mb_quad_add<Fq><<<blocks, threads, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 9: k_mb_quad_double<Fq><<<blocks, threads, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 10: k_mb_double_call<Fq><<<blocks, threads, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 11: k_mb_inv<Fq><<<blocks, threads, 0, rt.stream>>>((Fq*)d, iters); break;
            case 12: k_mb_inv_safegcd<Fq><<<blocks, threads, 0, rt.stream>>>((Fq*)d, iters); break;
            case 13: k_mb_madd_lazy<Fq><<<blocks, threads, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 14: k_mb_coop_add<Fq><<<blocks, COOP_THREADS, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 15: k_mb_coop_double<Fq><<<blocks, COOP_THREADS, 0, rt.stream>>>((XYZZ<Fq>*)d, iters); break;
            case 20: k_mb_pipe<20><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 21: k_mb_pipe<21><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 22: k_mb_pipe<22><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 23: k_mb_pipe<23><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 24: k_mb_pipe<24><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 25: k_mb_pipe<25><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 26: k_mb_pipe<26><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 27: k_mb_pipe<27><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 28: k_mb_pipe<28><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            case 29: k_mb_pipe<29><<<blocks, threads, 0, rt.stream>>>((uint32_t*)d, iters); break;
            default: cudaFree(d); set_error("sb_microbench: unknown test %d", which); return SB_ERR_ARG;
        }
        cudaEventRecord(e1, rt.stream);
        SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *out_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    SB_CUDA_TRY(cudaGetLastError());
    return SB_OK;
}
