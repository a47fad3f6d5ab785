// batch.cu -- plain batched field inversion (ff::BatchInvert: zeros stay zero); the kernel is lookup.cu's
// k_scaled_inverse (h = 1/(l + r), g = m/(t + r), reference src/plonk/lookup.rs:300-312) with shift 0, scale 1.
#include <string.h>

#include "common.cuh"
#include "field.cuh"

using namespace sb;

extern "C" {

int sb_batch_invert_device(int field, const void* d_in, void* d_out, size_t n, void* stream) {
    if ((!d_in || !d_out) && n) {
        set_error("sb_batch_invert_device: null argument");
        return SB_ERR_ARG;
    }
    // the warp-cooperative inversion kernel of lookup.cu with shift 0 and scale 1
    return sb_scaled_inverse_device(field, d_in, nullptr, nullptr, d_out, n, stream);
}

int sb_batch_invert(int field, const uint64_t* in, uint64_t* out, size_t n) {
    if ((!in || !out) && n) {
        set_error("sb_batch_invert: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // one critical section from staging to the synchronised download
    Scratch& g_bi_stage = ws_slot(rt.stream, WS_BI_STAGE);
    {
        SB_TRY(g_bi_stage.reserve(n * 64 + 64));
        SB_CUDA_TRY(cudaMemcpyAsync(g_bi_stage.ptr, in, n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    char* d = (char*)g_bi_stage.ptr;
    SB_TRY(sb_batch_invert_device(field, d, d + n * 32, n, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, d + n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

}  // extern "C"
