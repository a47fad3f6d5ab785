// ntt.cu -- radix-2 number-theoretic transform over bn256 Fr on one B200.
//
// Replaces reference src/fft.rs: `best_fft` (:61-115, in-place radix-2 DIT, natural order in and out, caller
// supplies omega) and the wrappers fft / ifft / coset_fft / coset_ifft (:160-198) incl. the ifft divisor
// (:25-27, :177-181) and distribute_powers_zeta (:207-228).
//
// Device algorithm: Stockham auto-sort over up to three passes of radix R <= 1024.  A pass reads column
// j (stride N/R), multiplies by the inter-pass twiddle w_{Ns*R}^{(j mod Ns) * r}, runs an R-point DIT
// transform entirely in shared memory (inputs staged bit-reversed, the R/2 inner twiddles staged in shared
// memory once per block), and writes out[(j / Ns) * Ns * R + (j mod Ns) + r * Ns].  Natural order in, natural
// order out, no separate bit-reversal or transpose pass.  All arithmetic is exact, so the result equals
// best_fft's bit for bit whatever the factorisation.
//
// Roofline: algorithmic traffic 64 B/element; arithmetic (log2 N)/2 + O(1) Montgomery products
// per element.
#include <stdlib.h>
#include <string.h>
#include <map>
#include <vector>
#include <array>

#include "common.cuh"
#include "field.cuh"

namespace sb {

constexpr int NTT_MAX_LOG_R = 10;
constexpr int NTT_LO_BITS = 10;
constexpr int NTT_FULL_MAX_LOG = 21;

template <class T>
SB_D T ld16(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = s[0];
    d[1] = s[1];
    return r;
}
template <class T>
SB_D void st16(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    d[0] = s[0];
    d[1] = s[1];
}

// out[i] = g^i
template <class F>
__global__ void k_powers(F g, uint32_t count, F* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    F acc = F::one(), base = g;
    for (uint32_t e = i; e; e >>= 1) {
        if (e & 1) acc = mul(acc, base);
        base = sqr(base);
    }
    st16(out + i, acc);
}

// shared-memory element store in two 16-byte planes: consecutive threads hit consecutive 16-byte words, so the
// butterflies with stride >= 8 elements are bank-conflict free (an array of 32-byte elements is 2-way conflicted
// on every 128-bit access)
template <class F>
SB_D F sm_load(const uint4* lo, const uint4* hi, uint32_t idx) {
    F r;
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = lo[idx];
    d[1] = hi[idx];
    return r;
}
template <class F>
SB_D void sm_store(uint4* lo, uint4* hi, uint32_t idx, const F& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    lo[idx] = s[0];
    hi[idx] = s[1];
}

// out[idx] = lo[e & (2^LO - 1)] * hi[e >> LO] with e = (idx >> logR) * (idx & (R - 1)): the inter-pass twiddle of column
// jm = idx >> logR and input r, as ONE table entry (built once per (size, omega), read coalesced by the pass kernel)
template <class F>
__global__ void k_full_twiddles(const F* __restrict__ tw_lo, const F* __restrict__ tw_hi, uint32_t logR, uint32_t count, F* __restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const uint32_t e = (idx >> logR) * (idx & ((1u << logR) - 1u));
    F tw = ld16(tw_lo + (e & ((1u << NTT_LO_BITS) - 1)));
    if (e >> NTT_LO_BITS) tw = mul(tw, ld16(tw_hi + (e >> NTT_LO_BITS)));
    st16(out + idx, tw);
}

// One pass.  Values stay in the lazy domain [0, 2p) between the stages (field.cuh: no final subtraction in the products)
// and are canonicalised when they leave the block.  Inner twiddles are staged PER STAGE, contiguously (stage s reads
// entries 2^s - 1 + i, i < 2^s): consecutive lanes read consecutive 16-byte words in every stage, where one strided table
// of R/2 entries gave 8-way conflicts in the middle stages (ncu, round 2: 1.48x the conflict-free wavefronts on loads).
template <class F>
__global__ void k_ntt_pass(const F* __restrict__ in, F* __restrict__ out, uint32_t logN, uint32_t logR, uint32_t logNs,
                           uint32_t log_tj, const F* __restrict__ inner_tw, const F* __restrict__ tw_lo,
                           const F* __restrict__ tw_hi, const F* __restrict__ tw_full, F scale, int has_scale) {
    extern __shared__ uint4 smem_raw[];
    const uint32_t R = 1u << logR, halfR = R >> 1;
    const uint32_t total = R << log_tj;           // data elements in the block
    uint4* d_lo = smem_raw;
    uint4* d_hi = smem_raw + total;
    uint4* t_lo = smem_raw + 2 * total;           // inner twiddles, stage by stage: R - 1 entries (R slots)
    uint4* t_hi = t_lo + R;

    const uint32_t tj = threadIdx.x >> (logR - 1);  // which transform of this block
    const uint32_t t = threadIdx.x & (halfR - 1);
    const uint32_t j = (blockIdx.x << log_tj) + tj;
    const uint32_t col_stride_log = logN - logR;
    const uint32_t Ns_mask = (1u << logNs) - 1;
    const uint32_t my = tj << logR;                // base index of this transform

    for (uint32_t k = threadIdx.x; k + 1 < R; k += blockDim.x) {
        const uint32_t s = 31u - (uint32_t)__clz(k + 1u), i = k + 1u - (1u << s);
        sm_store(t_lo, t_hi, k, ld16(inner_tw + (i << (logR - 1 - s))));
    }

#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t r = t + h * halfR;
        F v = ld16(in + (size_t)j + ((size_t)r << col_stride_log));
        if (logNs) {
            const uint32_t jm = j & Ns_mask, e = jm * r;
            if (e) {
                F tw;
                if (tw_full) {
                    tw = ld16(tw_full + ((size_t)jm << logR) + r);
                } else {
                    tw = ld16(tw_lo + (e & ((1u << NTT_LO_BITS) - 1)));
                    if (e >> NTT_LO_BITS) tw = mul_lazy(tw, ld16(tw_hi + (e >> NTT_LO_BITS)));
                }
                v = mul_lazy(v, tw);
            }
        }
        const uint32_t br = __brev(r) >> (32 - logR);
        sm_store(d_lo, d_hi, my + br, v);
    }
    __syncthreads();

    for (uint32_t s = 0; s < logR; s++) {
        const uint32_t half = 1u << s;
        const uint32_t i = t & (half - 1);
        const uint32_t p = my + ((t >> s) << (s + 1)) + i;
        F x = sm_load<F>(d_lo, d_hi, p);
        F y = sm_load<F>(d_lo, d_hi, p + half);
        if (i) y = mul_lazy(y, sm_load<F>(t_lo, t_hi, half - 1u + i));
        sm_store(d_lo, d_hi, p, add_lazy(x, y));
        sm_store(d_lo, d_hi, p + half, sub_lazy(x, y));
        __syncthreads();
    }

    const size_t j0 = ((size_t)(j >> logNs) << (logNs + logR)) + (j & Ns_mask);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t r = t + h * halfR;
        F v = sm_load<F>(d_lo, d_hi, my + r);
        if (has_scale) v = mul_lazy(v, scale);
        st16(out + j0 + ((size_t)r << logNs), canon(v));
    }
}

// n == 2: one butterfly, no shared memory needed
template <class F>
__global__ void k_ntt_2(F* a, F scale, int has_scale) {
    if (threadIdx.x || blockIdx.x) return;
    F x = ld16(a), y = ld16(a + 1);
    F u = add(x, y), v = sub(x, y);
    if (has_scale) {
        u = mul(u, scale);
        v = mul(v, scale);
    }
    st16(a, u);
    st16(a + 1, v);
}

// a[i] *= s
template <class F>
__global__ void k_scale(F* a, size_t n, F s) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st16(a + i, mul(ld16(a + i), s));
}

// distribute_powers_zeta (src/fft.rs:207-228): a[i] *= z[i%3 - 1] for i%3 != 0; z[0], z[1] as given
template <class F>
__global__ void k_coset_scale(F* a, size_t n, F z0, F z1) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r = (uint32_t)(i % 3);
    if (r) st16(a + i, mul(ld16(a + i), r == 1 ? z0 : z1));
}

// ------------------------------------------------------------------------------------------------
struct NttTables {
    int log_n = 0;
    int passes = 0;
    int logR[3] = {0, 0, 0};
    char* dev = nullptr;       // one allocation holding everything below
    size_t inner_off[3] = {0, 0, 0};
    size_t lo_off[3] = {0, 0, 0};
    size_t hi_off[3] = {0, 0, 0};
    size_t full_off[3] = {0, 0, 0};   // full inter-pass table of the pass (0 = none: too large, or first pass)
    bool has_full[3] = {false, false, false};
};

struct NttKey {
    int log_n;
    std::array<uint64_t, 4> omega;
    bool operator<(const NttKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        return omega < o.omega;
    }
};

static std::map<NttKey, NttTables> g_ntt_cache;

template <class F>
static F host_pow(F g, uint64_t e) {  // runs on the host only to derive per-pass twiddle bases
    F acc = F::one(), base = g;
    for (; e; e >>= 1) {
        if (e & 1) acc = mul_portable(acc, base);
        base = mul_portable(base, base);
    }
    return acc;
}

template <class F>
static int build_tables(int log_n, const uint64_t omega_limbs[4], cudaStream_t st, NttTables& T) {
    T.log_n = log_n;
    T.passes = (log_n + NTT_MAX_LOG_R - 1) / NTT_MAX_LOG_R;
    if (T.passes > 3) {
        set_error("sb_ntt: log_n = %d too large (max %d)", log_n, 3 * NTT_MAX_LOG_R);
        return SB_ERR_ARG;
    }
    {  // split log_n as evenly as possible, larger radices first
        int rem = log_n;
        for (int p = 0; p < T.passes; p++) {
            int left = T.passes - p;
            T.logR[p] = (rem + left - 1) / left;
            rem -= T.logR[p];
        }
    }
    F omega;
    memcpy(omega.v, omega_limbs, 32);
    size_t off = 0;
    auto take = [&](size_t elems) {
        size_t o = off;
        off += elems * sizeof(F);
        return o;
    };
    int logNs = 0;
    size_t hi_count[3] = {0, 0, 0};
    for (int p = 0; p < T.passes; p++) {
        T.inner_off[p] = take((size_t)1 << (T.logR[p] > 0 ? T.logR[p] - 1 : 0));
        if (logNs) {
            T.lo_off[p] = take((size_t)1 << NTT_LO_BITS);
            int span = logNs + T.logR[p];
            hi_count[p] = span > NTT_LO_BITS ? ((size_t)1 << (span - NTT_LO_BITS)) : 1;
            T.hi_off[p] = take(hi_count[p]);
            // one entry per (column mod Ns, input): saves the lo x hi product per element (1 of ~7 in the pass) for 32 more
            // bytes read per element; kept up to 2^21 entries = 64 MB per cached (size, omega) -- the two-pass sizes
            if (span <= NTT_FULL_MAX_LOG) {
                T.has_full[p] = true;
                T.full_off[p] = take((size_t)1 << span);
            }
        }
        logNs += T.logR[p];
    }
    SB_CUDA_TRY(cudaMalloc(&T.dev, off));
    logNs = 0;
    for (int p = 0; p < T.passes; p++) {
        // inner transform root: w_R = omega^(N/R)
        F wR = host_pow(omega, (uint64_t)1 << (log_n - T.logR[p]));
        uint32_t cnt = 1u << (T.logR[p] - 1);
        k_powers<F><<<(cnt + 127) / 128, 128, 0, st>>>(wR, cnt, (F*)(T.dev + T.inner_off[p]));
        SB_KERNEL_CHECK();
        if (logNs) {
            // inter-pass root: w_{Ns*R} = omega^(N/(Ns*R))
            F base = host_pow(omega, (uint64_t)1 << (log_n - logNs - T.logR[p]));
            k_powers<F><<<((1u << NTT_LO_BITS) + 127) / 128, 128, 0, st>>>(base, 1u << NTT_LO_BITS, (F*)(T.dev + T.lo_off[p]));
            SB_KERNEL_CHECK();
            F base_hi = host_pow(base, (uint64_t)1 << NTT_LO_BITS);
            k_powers<F><<<((uint32_t)hi_count[p] + 127) / 128, 128, 0, st>>>(base_hi, (uint32_t)hi_count[p], (F*)(T.dev + T.hi_off[p]));
            SB_KERNEL_CHECK();
            if (T.has_full[p]) {
                const uint32_t count = 1u << (logNs + T.logR[p]);
                k_full_twiddles<F><<<(count + 255) / 256, 256, 0, st>>>((const F*)(T.dev + T.lo_off[p]), (const F*)(T.dev + T.hi_off[p]), (uint32_t)T.logR[p], count,
                                                                      (F*)(T.dev + T.full_off[p]));
                SB_KERNEL_CHECK();
            }
        }
        logNs += T.logR[p];
    }
    return SB_OK;
}

// d_a: n = 2^log_n elements on the device, transformed in place.
template <class F>
static int ntt_enqueue(F* d_a, int log_n, const uint64_t omega[4], const uint64_t* scale, cudaStream_t st) {
    F sc = F::one();
    const int has_scale = scale != nullptr;
    if (scale) memcpy(sc.v, scale, 32);
    const size_t n = (size_t)1 << log_n;
    if (log_n == 0) {  // best_fft on one element is the identity
        if (has_scale) {
            k_scale<F><<<1, 32, 0, st>>>(d_a, 1, sc);
            SB_KERNEL_CHECK();
        }
        return SB_OK;
    }
    if (log_n == 1) {
        k_ntt_2<F><<<1, 32, 0, st>>>(d_a, sc, has_scale);
        SB_KERNEL_CHECK();
        return SB_OK;
    }
    NttKey key{log_n, {omega[0], omega[1], omega[2], omega[3]}};
    auto it = g_ntt_cache.find(key);
    if (it == g_ntt_cache.end()) {
        NttTables T;
        SB_TRY(build_tables<F>(log_n, omega, st, T));
        SB_CUDA_TRY(cudaStreamSynchronize(st));   // the cached tables are read by later calls on ANY stream
        it = g_ntt_cache.emplace(key, T).first;
    }
    NttTables& T = it->second;
    Scratch& g_ntt_tmp = ws_slot(st, WS_NTT_TMP);
    SB_TRY(g_ntt_tmp.reserve(n * sizeof(F)));
    F* src = d_a;
    F* dst = (F*)g_ntt_tmp.ptr;
    int logNs = 0;
    SB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    static const int use_full = []() {
        const char* e = getenv("SB_NTT_FULL_TWIDDLES");
        return (!e || atoi(e) != 0) ? 1 : 0;
    }();
    for (int p = 0; p < T.passes; p++) {
        const int logR = T.logR[p];
        // transforms per block: fill 256 threads when R is small
        int log_tj = 9 - logR;
        if (log_tj < 0) log_tj = 0;
        if (log_tj > log_n - logR) log_tj = log_n - logR;
        const uint32_t threads = 1u << (logR - 1 + log_tj);
        const uint32_t blocks = 1u << (log_n - logR - log_tj);
        const size_t smem = (((size_t)1 << (logR + log_tj)) + ((size_t)1 << logR)) * sizeof(F);
        const bool last = (p == T.passes - 1);
        k_ntt_pass<F><<<blocks, threads, smem, st>>>(src, dst, (uint32_t)log_n, (uint32_t)logR, (uint32_t)logNs, (uint32_t)log_tj,
                                                    (const F*)(T.dev + T.inner_off[p]), logNs ? (const F*)(T.dev + T.lo_off[p]) : nullptr,
                                                    logNs ? (const F*)(T.dev + T.hi_off[p]) : nullptr,
                                                    (logNs && T.has_full[p] && use_full) ? (const F*)(T.dev + T.full_off[p]) : nullptr, sc, last ? has_scale : 0);
        SB_KERNEL_CHECK();
        std::swap(src, dst);
        logNs += logR;
    }
    if (src != d_a) SB_CUDA_TRY(cudaMemcpyAsync(d_a, src, n * sizeof(F), cudaMemcpyDeviceToDevice, st));
    return SB_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_ntt_device(int field, void* d_a, uint32_t log_n, const uint64_t omega[4], const uint64_t* scale, void* stream) {
    if (!d_a || !omega) {
        set_error("sb_ntt_device: null argument");
        return SB_ERR_ARG;
    }
    if (field != FIELD_FR) {
        set_error("sb_ntt: only bn256 Fr has a 2-adic subgroup (Fq has 2-adicity 1, SURVEY App. D)");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    return ntt_enqueue<Fr>((Fr*)d_a, (int)log_n, omega, scale, stream ? (cudaStream_t)stream : rt.stream);
}

int sb_ntt(int field, uint64_t* a, uint32_t log_n, const uint64_t omega[4], const uint64_t* scale) {
    if (!a || !omega) {
        set_error("sb_ntt: null argument");
        return SB_ERR_ARG;
    }
    if (field != FIELD_FR) {
        set_error("sb_ntt: only bn256 Fr has a 2-adic subgroup (Fq has 2-adicity 1, SURVEY App. D)");
        return SB_ERR_ARG;
    }
    if (log_n > 30) {
        set_error("sb_ntt: log_n = %u too large", log_n);
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    const size_t bytes = ((size_t)1 << log_n) * 32;
    Scratch& g_ntt_stage = ws_slot(rt.stream, WS_NTT_STAGE);
    SB_TRY(g_ntt_stage.reserve(bytes));
    SB_CUDA_TRY(cudaMemcpyAsync(g_ntt_stage.ptr, a, bytes, cudaMemcpyHostToDevice, rt.stream));
    SB_TRY(ntt_enqueue<Fr>((Fr*)g_ntt_stage.ptr, (int)log_n, omega, scale, rt.stream));
    SB_CUDA_TRY(cudaMemcpyAsync(a, g_ntt_stage.ptr, bytes, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

static int coset_scale_impl(Fr* d_a, size_t n, const uint64_t z[4], const uint64_t z2[4], cudaStream_t st) {
    Fr z0, z1;
    memcpy(z0.v, z, 32);
    memcpy(z1.v, z2, 32);
    if (n) {
        k_coset_scale<Fr><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, n, z0, z1);
        SB_KERNEL_CHECK();
    }
    return SB_OK;
}

int sb_coset_scale_device(int field, void* d_a, size_t n, const uint64_t z[4], const uint64_t z2[4], void* stream) {
    if ((!d_a && n) || !z || !z2 || field != FIELD_FR) {
        set_error("sb_coset_scale_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    return coset_scale_impl((Fr*)d_a, n, z, z2, stream ? (cudaStream_t)stream : rt.stream);
}

int sb_coset_scale(int field, uint64_t* a, size_t n, const uint64_t z[4], const uint64_t z2[4]) {
    if ((!a && n) || !z || !z2 || field != FIELD_FR) {
        set_error("sb_coset_scale: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    Scratch& g_ntt_stage = ws_slot(rt.stream, WS_NTT_STAGE);
    SB_TRY(g_ntt_stage.reserve(n * 32 + 32));
    SB_CUDA_TRY(cudaMemcpyAsync(g_ntt_stage.ptr, a, n * 32, cudaMemcpyHostToDevice, rt.stream));
    SB_TRY(coset_scale_impl((Fr*)g_ntt_stage.ptr, n, z, z2, rt.stream));
    SB_CUDA_TRY(cudaMemcpyAsync(a, g_ntt_stage.ptr, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

}  // extern "C"
