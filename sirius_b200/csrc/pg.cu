// pg.cu -- Protogalaxy prover kernels: the beta-weighted binary tree and the Lagrange witness fold.
//
// Replaces, on the device (reference src/nifs/protogalaxy):
//   * the `tree_reduce` of compute_F (poly/mod.rs:68-203), compute_G (:308-425) and evaluate_e_from_trace
//     (mod.rs:571-640): a perfect binary tree over n = 2^t leaves whose node at height h is
//     left + right * c[h]  (c[h] = beta_h + X*delta^(2^h) for F, beta*_h for G, beta_h for e), evaluated for P
//     points at once.  Leaves come from sb_pg_leaves_device (expr.cu), which evaluates the gate programs on the
//     Lagrange blend of the traces without materialising the folded witnesses (poly/folded_witness.rs:66-143).
//   * ProtoGalaxy::fold_witness (mod.rs:176-210): w = sum_j L_j(gamma) * w_j, cell by cell.
//
// The tree is evaluated level-synchronously in groups of 2^12 inputs per block (16 per thread in registers, then
// a shared-memory tree); exact arithmetic makes the result independent of the grouping.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "field.cuh"

namespace sb {

constexpr int TREE_LOCAL_LOG = 4;   // inputs per thread = 16
constexpr int TREE_GROUP_LOG = 12;  // inputs per block  = 4096

template <class T>
SB_D T ld32(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = s[0];
    d[1] = s[1];
    return r;
}
template <class T>
SB_D void st32(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    d[0] = s[0];
    d[1] = s[1];
}

// out[p][group] = tree over in[p][group * 2^log_group ...] with multipliers c[p][h0 + level]
template <class F>
__global__ void __launch_bounds__(256)
k_beta_tree(const F* __restrict__ in, size_t in_stride, uint32_t h0, const F* __restrict__ c, uint32_t c_stride,
            F* __restrict__ out, size_t out_stride, uint32_t log_group) {
    __shared__ F sh[256];
    const uint32_t p = blockIdx.y;
    const F* src = in + (size_t)p * in_stride + ((size_t)blockIdx.x << log_group);
    const F* cp = c + (size_t)p * c_stride + h0;
    const uint32_t s = log_group < (uint32_t)TREE_LOCAL_LOG ? log_group : TREE_LOCAL_LOG;
    const uint32_t tid = threadIdx.x;
    F v[1 << TREE_LOCAL_LOG];
#pragma unroll
    for (int j = 0; j < (1 << TREE_LOCAL_LOG); j++) {
        if ((uint32_t)j < (1u << s)) v[j] = ld32(src + ((size_t)tid << s) + j);
        else v[j] = F::zero();
    }
#pragma unroll
    for (int level = 0; level < TREE_LOCAL_LOG; level++) {
        if ((uint32_t)level < s) {
            const F m = ld32(cp + level);
#pragma unroll
            for (int j = 0; j < ((1 << TREE_LOCAL_LOG) >> (level + 1)); j++) v[j] = add(v[2 * j], mul(v[2 * j + 1], m));
        }
    }
    sh[tid] = v[0];
    __syncthreads();
    for (uint32_t level = s; level < log_group; level++) {
        const uint32_t d = 1u << (level - s);
        if ((tid & (2 * d - 1)) == 0) {
            F a = sh[tid], b = sh[tid + d];
            sh[tid] = add(a, mul(b, ld32(cp + level)));
        }
        __syncthreads();
    }
    if (tid == 0) st32(out + (size_t)p * out_stride + blockIdx.x, sh[0]);
}

// out[i] = sum_j coef[j] * in_j[i]
template <class F>
__global__ void k_lincomb(const F* const* __restrict__ ins, const F* __restrict__ coef, uint32_t J, F* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F acc = mul(ld32(coef), ld32(ins[0] + i));
    for (uint32_t j = 1; j < J; j++) acc = add(acc, mul(ld32(coef + j), ld32(ins[j] + i)));
    st32(out + i, acc);
}


template <class F>
static int beta_tree_enqueue(const void* d_leaves, uint32_t log_n, size_t num_points, size_t leaf_stride,
                             const uint64_t* multipliers, void* d_out, cudaStream_t st) {
    const size_t n = (size_t)1 << log_n;
    const size_t groups0 = log_n > (uint32_t)TREE_GROUP_LOG ? (n >> TREE_GROUP_LOG) : 1;
    const size_t c_bytes = align_up(32 * num_points * (log_n ? log_n : 1), 256);
    const size_t tmp_elems = num_points * groups0;
    Scratch& g_pg_ws = ws_slot(st, WS_PG);
    SB_TRY(g_pg_ws.reserve(c_bytes + 2 * align_up(tmp_elems * 32, 256)));
    char* ws = (char*)g_pg_ws.ptr;
    F* d_c = (F*)ws;
    F* tmp[2] = {(F*)(ws + c_bytes), (F*)(ws + c_bytes + align_up(tmp_elems * 32, 256))};
    if (log_n == 0) {
        for (size_t p = 0; p < num_points; p++)
            SB_CUDA_TRY(cudaMemcpyAsync((F*)d_out + p, (const F*)d_leaves + p * leaf_stride, 32, cudaMemcpyDeviceToDevice, st));
        return SB_OK;
    }
    SB_CUDA_TRY(cudaMemcpyAsync(d_c, multipliers, 32 * num_points * log_n, cudaMemcpyHostToDevice, st));
    ProfScope ps(st, PROF_PG, n * num_points);
    const F* in = (const F*)d_leaves;
    size_t in_stride = leaf_stride;
    uint32_t log_cur = log_n, h0 = 0;
    int flip = 0;
    while (log_cur > 0) {
        const uint32_t log_group = log_cur < (uint32_t)TREE_GROUP_LOG ? log_cur : TREE_GROUP_LOG;
        const size_t groups = (size_t)1 << (log_cur - log_group);
        const uint32_t s = log_group < (uint32_t)TREE_LOCAL_LOG ? log_group : TREE_LOCAL_LOG;
        const uint32_t threads = 1u << (log_group - s);
        const bool last = (log_cur == log_group);
        F* out = last ? (F*)d_out : tmp[flip];
        const size_t out_stride = last ? 1 : groups;
        dim3 grid((unsigned)groups, (unsigned)num_points);
        k_beta_tree<F><<<grid, threads, 0, st>>>(in, in_stride, h0, d_c, log_n, out, out_stride, log_group);
        SB_KERNEL_CHECK();
        in = out;
        in_stride = out_stride;
        h0 += log_group;
        log_cur -= log_group;
        flip ^= 1;
    }
    return SB_OK;
}


template <class F>
static int lincomb_enqueue(const void* const* d_inputs, const uint64_t* coef, size_t J, size_t n, void* d_out, cudaStream_t st) {
    const size_t ptr_bytes = align_up(sizeof(void*) * J, 32);
    Scratch& g_lincomb_args = ws_slot(st, WS_LINCOMB_ARGS);
    SB_TRY(g_lincomb_args.reserve(ptr_bytes + 32 * J));
    char* d = (char*)g_lincomb_args.ptr;
    SB_CUDA_TRY(cudaMemcpyAsync(d, d_inputs, sizeof(void*) * J, cudaMemcpyHostToDevice, st));
    SB_CUDA_TRY(cudaMemcpyAsync(d + ptr_bytes, coef, 32 * J, cudaMemcpyHostToDevice, st));
    ProfScope ps(st, PROF_PG, n);
    if (n) {
        k_lincomb<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const F* const*)d, (const F*)(d + ptr_bytes), (uint32_t)J, (F*)d_out, n);
        SB_KERNEL_CHECK();
    }
    return SB_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_beta_tree_device(int field, const void* d_leaves, uint32_t log_n, size_t num_points, size_t leaf_stride,
                        const uint64_t* multipliers, void* d_out, void* stream) {
    if (!d_leaves || !d_out || !num_points || (!multipliers && log_n) || log_n > 30) {
        set_error("sb_beta_tree_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (field == FIELD_FR) return beta_tree_enqueue<Fr>(d_leaves, log_n, num_points, leaf_stride, multipliers, d_out, st);
    if (field == FIELD_FQ) return beta_tree_enqueue<Fq>(d_leaves, log_n, num_points, leaf_stride, multipliers, d_out, st);
    set_error("sb_beta_tree_device: unknown field %d", field);
    return SB_ERR_ARG;
}

int sb_lincomb_device(int field, const void* const* d_inputs, const uint64_t* coef, size_t num_inputs, size_t n, void* d_out, void* stream) {
    if (!d_inputs || !coef || !num_inputs || (!d_out && n)) {
        set_error("sb_lincomb_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (field == FIELD_FR) return lincomb_enqueue<Fr>(d_inputs, coef, num_inputs, n, d_out, st);
    if (field == FIELD_FQ) return lincomb_enqueue<Fq>(d_inputs, coef, num_inputs, n, d_out, st);
    set_error("sb_lincomb_device: unknown field %d", field);
    return SB_ERR_ARG;
}


/* Host-memory front end: leaves for `num_blends` Lagrange blends of `num_traces` witnesses (single round each),
 * then the beta tree for `num_points` points (point p reads blend point_blend[p]). */
int sb_pg_tree(sb_prog_t const* gates, size_t num_gates, sb_columns_t cols, uint32_t num_advice, const uint64_t* const* traces_W,
               size_t num_traces, const uint64_t* coef, const uint64_t* challenges, size_t num_challenges, size_t num_blends,
               int row_mode, uint32_t log_leaves, const uint64_t* multipliers, size_t num_points, const uint32_t* point_blend,
               uint64_t* out) {
    if (!gates || !cols || !traces_W || !coef || !multipliers && log_leaves || !point_blend || !out || !num_traces || !num_blends || !num_points) {
        set_error("sb_pg_tree: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // one critical section from staging to the synchronised download (the stage buffer is shared)
    Scratch& g_pg_stage = ws_slot(rt.stream, WS_PG_STAGE);
    const size_t n = (size_t)1 << sb_columns_log_rows(cols);
    const size_t leaves = (size_t)1 << log_leaves;
    const size_t w_bytes = align_up((size_t)num_advice * n * 32, 256);
    const size_t leaves_off = w_bytes * num_traces;
    const size_t out_off = leaves_off + align_up(leaves * num_blends * 32, 256);
    const int field = sb_expr_field(gates[0]);
    std::vector<const void*> tables((size_t)num_advice * num_traces);
    {
        RtLock lk(rt.mu);
        SB_TRY(g_pg_stage.reserve(out_off + 32 * num_points + 256));
        char* base = (char*)g_pg_stage.ptr;
        for (size_t j = 0; j < num_traces; j++) {
            SB_CUDA_TRY(cudaMemcpyAsync(base + j * w_bytes, traces_W[j], (size_t)num_advice * n * 32, cudaMemcpyHostToDevice, rt.stream));
            for (uint32_t a = 0; a < num_advice; a++) tables[j * num_advice + a] = base + j * w_bytes + (size_t)a * n * 32;
        }
    }
    char* base = (char*)g_pg_stage.ptr;
    SB_TRY(sb_pg_leaves_device(gates, num_gates, cols, tables.data(), num_traces, num_advice, coef, challenges, num_challenges, num_blends,
                               row_mode, log_leaves, base + leaves_off, nullptr));
    // points sharing a blend are consecutive runs in practice (F: all -> 0, G: identity); run one tree per run
    size_t p = 0;
    while (p < num_points) {
        size_t q = p;
        const bool shared = (p + 1 < num_points && point_blend[p + 1] == point_blend[p]);
        if (shared) {
            while (q + 1 < num_points && point_blend[q + 1] == point_blend[p]) q++;
            SB_TRY(sb_beta_tree_device(field, base + leaves_off + (size_t)point_blend[p] * leaves * 32, log_leaves, q - p + 1, 0,
                                       multipliers + 4 * (size_t)p * log_leaves, base + out_off + 32 * p, nullptr));
        } else {
            while (q + 1 < num_points && point_blend[q + 1] == point_blend[q] + 1) q++;
            SB_TRY(sb_beta_tree_device(field, base + leaves_off + (size_t)point_blend[p] * leaves * 32, log_leaves, q - p + 1, leaves,
                                       multipliers + 4 * (size_t)p * log_leaves, base + out_off + 32 * p, nullptr));
        }
        p = q + 1;
    }
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, base + out_off, 32 * num_points, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

/* ProtoGalaxy::fold_witness on host vectors: out = sum_j coef[j] * W_j */
int sb_lincomb(int field, const uint64_t* const* inputs, const uint64_t* coef, size_t num_inputs, size_t n, uint64_t* out) {
    if (!inputs || !coef || !num_inputs || (!out && n)) {
        set_error("sb_lincomb: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);
    Scratch& g_pg_stage = ws_slot(rt.stream, WS_PG_STAGE);
    std::vector<const void*> ptrs(num_inputs);
    {
        RtLock lk(rt.mu);
        SB_TRY(g_pg_stage.reserve((num_inputs + 1) * n * 32 + 256));
        char* base = (char*)g_pg_stage.ptr;
        for (size_t j = 0; j < num_inputs; j++) {
            SB_CUDA_TRY(cudaMemcpyAsync(base + j * n * 32, inputs[j], n * 32, cudaMemcpyHostToDevice, rt.stream));
            ptrs[j] = base + j * n * 32;
        }
    }
    char* base = (char*)g_pg_stage.ptr;
    SB_TRY(sb_lincomb_device(field, ptrs.data(), coef, num_inputs, n, base + num_inputs * n * 32, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, base + num_inputs * n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

}  // extern "C"
