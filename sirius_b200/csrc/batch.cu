// batch.cu -- batched field inversion (Montgomery's trick), the building block of the SPS lookup columns
// h = 1/(l + r), g = m/(t + r) (reference src/plonk/lookup.rs:213-365) and of util::batch_invert_assigned
// (src/util/mod.rs:128-153) -- the "next" rows 3/4 of SURVEY 8f.  Zeros stay zero, as in ff::BatchInvert.
#include <string.h>

#include "common.cuh"
#include "field.cuh"

namespace sb {

constexpr int BI_CHUNK = 16;  // elements per thread: 3 products each + one binary-GCD inversion per chunk

template <class F>
__global__ void __launch_bounds__(128)
k_batch_invert(const F* __restrict__ in, F* __restrict__ out, size_t n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t base = t * BI_CHUNK;
    if (base >= n) return;
    const int cnt = (n - base) < (size_t)BI_CHUNK ? (int)(n - base) : BI_CHUNK;
    F pre[BI_CHUNK];
    F acc = F::one();
#pragma unroll
    for (int i = 0; i < BI_CHUNK; i++) {
        if (i < cnt) {
            F v = in[base + i];
            pre[i] = acc;
            if (!v.is_zero()) acc = mul(acc, v);
        }
    }
    F inv_all = inv_binary(acc);  // acc is a product of non-zero elements (or one)
#pragma unroll
    for (int i = BI_CHUNK - 1; i >= 0; i--) {
        if (i < cnt) {
            F v = in[base + i];
            if (v.is_zero()) {
                out[base + i] = v;
            } else {
                out[base + i] = mul(inv_all, pre[i]);
                inv_all = mul(inv_all, v);
            }
        }
    }
}

static Scratch g_bi_stage;

}  // namespace sb

using namespace sb;

extern "C" {

int sb_batch_invert_device(int field, const void* d_in, void* d_out, size_t n, void* stream) {
    if ((!d_in || !d_out) && n) {
        set_error("sb_batch_invert_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (!n) return SB_OK;
    const size_t threads = (n + BI_CHUNK - 1) / BI_CHUNK;
    const unsigned blocks = (unsigned)((threads + 127) / 128);
    if (field == FIELD_FR) k_batch_invert<Fr><<<blocks, 128, 0, st>>>((const Fr*)d_in, (Fr*)d_out, n);
    else if (field == FIELD_FQ) k_batch_invert<Fq><<<blocks, 128, 0, st>>>((const Fq*)d_in, (Fq*)d_out, n);
    else {
        set_error("sb_batch_invert_device: unknown field %d", field);
        return SB_ERR_ARG;
    }
    SB_KERNEL_CHECK();
    return SB_OK;
}

int sb_batch_invert(int field, const uint64_t* in, uint64_t* out, size_t n) {
    if ((!in || !out) && n) {
        set_error("sb_batch_invert: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    {
        std::lock_guard<std::mutex> lk(rt.mu);
        SB_TRY(g_bi_stage.reserve(n * 64 + 64));
        SB_CUDA_TRY(cudaMemcpyAsync(g_bi_stage.ptr, in, n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    char* d = (char*)g_bi_stage.ptr;
    SB_TRY(sb_batch_invert_device(field, d, d + n * 32, n, nullptr));
    std::lock_guard<std::mutex> lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, d + n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

}  // extern "C"
