// expr.cu -- gate-expression evaluation over all rows, Sangria cross terms, and the witness folds.
//
// Replaces, on the device:
//   * GraphEvaluator::evaluate (reference src/polynomial/graph_evaluator.rs:361-388) driven row by row from
//     VanillaFS::commit_cross_terms (src/nifs/sangria/mod.rs:110-147), is_sat_accumulation (:334-383) and
//     PlonkStructure::is_sat (src/plonk/mod.rs:304-361).  The Rust side keeps building `Expression`s and
//     compiling them with GraphEvaluator::new; the compiled calculation list crosses the ABI (sb_expr_compile).
//   * the cross terms T_1..T_d themselves: the reference expands the homogeneous gate polynomial P into
//     degree-grouped expressions (GroupedPoly, src/polynomial/grouped_poly.rs:58-110,216-268) and evaluates each
//     with its own pass over the columns.  T_j is the coefficient of X^j in P(w1 + X*w2, c1 + X*c2) (fixed and
//     selector columns are not folded), so here ONE fused pass evaluates P at X = 0..d for every row and applies
//     the inverse Vandermonde matrix -- exact field arithmetic, hence the same bits.
//   * RelaxedPlonkWitness::fold (src/nifs/sangria/accumulator.rs:363-404): W1 + r*W2 and E + sum r^j T_j.
//
// Kernel shape: one thread per row, the calculation list is interpreted uniformly by the whole block (no
// divergence), intermediates live in shared memory slots assigned by liveness at compile time, column reads are
// coalesced (consecutive rows -> consecutive 32-byte elements).
#include <string.h>

#include <nvrtc.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "field.cuh"

namespace sb {

constexpr int EXPR_THREADS = 128;
constexpr int EXPR_MAX_DEGREE = 31;   // d + 1 <= 32 evaluation points per row (a compressed expression of g gates of degree 5 has d = 5 + g - 1)

enum : uint32_t { VS_CONSTANT = 0, VS_INTERMEDIATE = 1, VS_FIXED = 2, VS_POLY = 3, VS_CHALLENGE = 4 };
enum : uint32_t { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_SQUARE = 3, OP_DOUBLE = 4, OP_NEGATE = 5, OP_HORNER = 6, OP_STORE = 7 };

struct DevOp {       // 16 bytes, read uniformly by every thread
    uint32_t code;   // op | a_kind << 8 | b_kind << 16
    uint32_t a, b;   // constant index / slot / column | rot_idx << 24 / challenge index
    uint32_t dst;    // slot
};

struct ColumnsDev {
    uint32_t log_rows;
    uint32_t num_selectors, num_fixed;
    const uint8_t* const* selectors;  // device array of device pointers
    const void* const* fixed;
};

template <class F>
struct EvalArgs {
    const DevOp* ops;
    uint32_t num_ops;
    uint32_t num_slots;
    uint32_t result_slot;
    const F* constants;
    const int32_t* rotations;
    ColumnsDev cols;
    uint32_t num_fold_vars;        // advice + 5 * lookups of ONE instance
    const F* const* adv1;          // num_fold_vars column pointers (instance 1)
    const F* const* adv2;          // instance 2 (cross terms / two-instance expressions) or nullptr
    const F* challenges;           // single evaluation: challenge table; cross terms: [(d+1)][num_challenges]
    uint32_t num_challenges;
    // Lagrange blend (Protogalaxy FoldedWitness, poly/folded_witness.rs:66-143, never materialised here):
    // fold variable = sum_j coef[t*num_blend + j] * W_j, W_j's column table at blend_cols[j*num_fold_vars ..]
    uint32_t num_blend;            // 0: off
    const F* const* blend_cols;    // [num_blend][num_fold_vars] device column pointers
    const F* blend_coef;           // [evaluations][num_blend]
};

template <class T>
SB_D T ldg32(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = __ldg(s);
    d[1] = __ldg(s + 1);
    return r;
}
template <class T>
SB_D void stg32(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    d[0] = s[0];
    d[1] = s[1];
}

// shared-memory slot file: slot s of thread t is two 16-byte halves in separate planes (conflict-free)
template <class F>
SB_D F slot_load(const uint4* sm, uint32_t slot, uint32_t nthreads) {
    F r;
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = sm[(slot * 2) * nthreads + threadIdx.x];
    d[1] = sm[(slot * 2 + 1) * nthreads + threadIdx.x];
    return r;
}
template <class F>
SB_D void slot_store(uint4* sm, uint32_t slot, uint32_t nthreads, const F& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    sm[(slot * 2) * nthreads + threadIdx.x] = s[0];
    sm[(slot * 2 + 1) * nthreads + threadIdx.x] = s[1];
}

// Value of column variable `index` (selectors -> fixed -> fold variables, reference src/plonk/eval.rs:57-69) at
// `row`.  Fold variables of instance 1 are blended with instance 2 at the integer point t: w1 + t*w2
// (t == 0 and two-instance indices >= num_fold_vars reproduce PlonkEvalDomain::eval_advice_var, :153-228).
template <class F>
SB_D F column_value(const EvalArgs<F>& A, uint32_t index, uint32_t row, uint32_t t) {
    if (index < A.cols.num_selectors) {
        return A.cols.selectors[index][row] ? F::one() : F::zero();
    }
    index -= A.cols.num_selectors;
    if (index < A.cols.num_fixed) {
        return ldg32(reinterpret_cast<const F*>(A.cols.fixed[index]) + row);
    }
    index -= A.cols.num_fixed;
    if (A.num_blend) {
        F acc = F::zero();
        for (uint32_t j = 0; j < A.num_blend; j++) {
            F w = ldg32(A.blend_cols[(size_t)j * A.num_fold_vars + index] + row);
            acc = add(acc, mul(ldg32(A.blend_coef + (size_t)t * A.num_blend + j), w));
        }
        return acc;
    }
    if (index >= A.num_fold_vars) {  // explicit second-instance variable (grouped / folded expressions)
        return ldg32(A.adv2[index - A.num_fold_vars] + row);
    }
    F v = ldg32(A.adv1[index] + row);
    if (t) {
        F w = ldg32(A.adv2[index] + row);
        for (uint32_t k = 0; k < t; k++) v = add(v, w);
    }
    return v;
}

template <class F>
SB_D F fetch(const EvalArgs<F>& A, uint32_t kind, uint32_t v, const uint4* sm, uint32_t row, uint32_t row_mask, uint32_t t) {
    switch (kind) {
        case VS_CONSTANT: return ldg32(A.constants + v);
        case VS_INTERMEDIATE: return slot_load<F>(sm, v, blockDim.x);
        case VS_CHALLENGE: return ldg32(A.challenges + (size_t)t * A.num_challenges + v);
        default: {  // VS_POLY / VS_FIXED
            const int32_t rot = A.rotations[v >> 24];
            const uint32_t r = (uint32_t)((int32_t)row + rot) & row_mask;  // get_rotation_idx, graph_evaluator.rs:51-53
            if (kind == VS_FIXED) return ldg32(reinterpret_cast<const F*>(A.cols.fixed[v & 0xffffffu]) + r);
            return column_value(A, v & 0xffffffu, r, t);
        }
    }
}

// runs the calculation list once for (row, t); result left in registers.  Intermediates are kept in the LAZY domain
// [0, 2p) (field.cuh): products skip their final conditional subtraction; callers canonicalise what they store.
template <class F>
SB_D F run_program(const EvalArgs<F>& A, uint4* sm, uint32_t row, uint32_t row_mask, uint32_t t) {
    F last = F::zero();
    for (uint32_t i = 0; i < A.num_ops; i++) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(A.ops + i));
        const uint32_t op = raw.x & 0xff, ak = (raw.x >> 8) & 0xff, bk = (raw.x >> 16) & 0xff;
        F a = fetch(A, ak, raw.y, sm, row, row_mask, t);
        F r;
        switch (op) {
            case OP_ADD: r = add_lazy(a, fetch(A, bk, raw.z, sm, row, row_mask, t)); break;
            case OP_SUB: r = sub_lazy(a, fetch(A, bk, raw.z, sm, row, row_mask, t)); break;
            case OP_MUL: r = mul_lazy(a, fetch(A, bk, raw.z, sm, row, row_mask, t)); break;
            case OP_SQUARE: r = mul_lazy(a, a); break;
            case OP_DOUBLE: r = dbl_lazy(a); break;
            case OP_NEGATE: r = neg_lazy(a); break;
            default: r = a; break;  // OP_STORE
        }
        slot_store(sm, raw.w, blockDim.x, r);
        last = r;
    }
    return last;  // GraphEvaluator::evaluate returns the last calculation's value (:382-387)
}

template <class F>
__global__ void __launch_bounds__(EXPR_THREADS)
k_expr_eval(EvalArgs<F> A, uint32_t t, int row0_only, F* __restrict__ out) {
    extern __shared__ uint4 sm[];
    const uint32_t n = 1u << A.cols.log_rows;
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    // row0_only reproduces the reference's `index & total_row` leaf addressing (src/plonk/mod.rs:714)
    F r = run_program(A, sm, row0_only ? 0u : row, n - 1, t);
    stg32(out + row, canon(r));
}

// out[(j-1)*n + row] = T_j(row), j = 1..degree;  vinv is the (degree+1)^2 inverse Vandermonde on points 0..degree.
// One thread per (row, evaluation point t): a block holds rows_per_block rows x (degree+1) points, the evaluations
// meet in shared memory and thread (row, t) then produces T_{t+1}.  (d+1) times the parallelism of a thread per row,
// which matters when the rows are sharded over several GPUs; column loads stay coalesced (the d+1 threads of a row
// read the same address).
template <class F>
__global__ void __launch_bounds__(EXPR_THREADS)
k_cross_terms(EvalArgs<F> A, uint32_t degree, uint32_t rows_per_block, uint32_t row0, uint32_t row_end, const F* __restrict__ vinv,
              F* __restrict__ out) {
    extern __shared__ uint4 sm[];
    const uint32_t n = 1u << A.cols.log_rows;
    const uint32_t m = degree + 1;
    const uint32_t r_in = threadIdx.x / m, t = threadIdx.x - r_in * m;
    const uint32_t row = row0 + blockIdx.x * rows_per_block + r_in;   // rows [row0, row_end) of the table
    const bool live = r_in < rows_per_block && row < row_end;
    // exchange area behind the slot file: [rows_per_block][m] evaluations
    F* exch = reinterpret_cast<F*>(sm + (size_t)(A.num_slots ? A.num_slots : 1) * 2 * blockDim.x);
    if (live) {
        F e = run_program(A, sm, row, n - 1, t);
        exch[r_in * m + t] = e;
    }
    __syncthreads();
    if (live && t + 1 <= degree) {
        const uint32_t j = t + 1;
        F acc = F::zero();
        for (uint32_t s = 0; s <= degree; s++) acc = add_lazy(acc, mul_lazy(ldg32(vinv + j * m + s), exch[r_in * m + s]));
        stg32(out + (size_t)(j - 1) * n + row, canon(acc));
    }
}

// W[i] = W1[i] + r * W2[i]      (accumulator.rs:366-378)
template <class F>
__global__ void k_axpy(const F* __restrict__ w1, const F* __restrict__ w2, F r, F* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    stg32(out + i, add(ldg32(w1 + i), mul(r, ldg32(w2 + i))));
}

// E[i] = E[i] + sum_{j=1..d} r^j * T_j[i]   (accumulator.rs:386-397), T contiguous [d][n], rpow = [r, r^2, ...]
template <class F>
__global__ void k_error_fold(const F* __restrict__ e, const F* __restrict__ T, const F* __restrict__ rpow, uint32_t d,
                             F* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F acc = ldg32(e + i);
    for (uint32_t j = 0; j < d; j++) acc = add(acc, mul(ldg32(rpow + j), ldg32(T + (size_t)j * n + i)));
    stg32(out + i, acc);
}

}  // namespace sb

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using namespace sb;

struct sb_prog {
    int field;
    uint32_t num_ops, num_slots, result_slot;
    uint32_t num_constants, num_rotations;
    void* d_ops;
    void* d_constants;
    void* d_rotations;
    uint32_t vinv_degree;  // inverse Vandermonde on 0..degree, built once per (program, degree)
    void* d_vinv;
    std::vector<uint32_t> poly_indices;  // distinct column indices the program reads (for validation)
    std::vector<uint32_t> fixed_indices; // ValueSource::Fixed column indices (checked against the registered columns)
    std::vector<sb::DevOp> h_ops;        // host copy of the lowered program: input of the straight-line code generator (expr_jit)
    std::vector<int32_t> h_rotations;
    std::vector<struct sb_jit_entry*> jit;   // compiled cross-term kernels, one per (degree, column layout)
    uint32_t max_challenge;
    bool uses_challenge;
};

struct sb_columns {
    int field;
    uint32_t log_rows;
    uint32_t num_selectors, num_fixed;
    void* d_data;            // selectors then fixed, contiguous
    void* d_ptrs;            // [num_selectors + num_fixed] device pointers
};

namespace sb {

template <class F>
static void host_vandermonde_inverse(uint32_t degree, std::vector<F>& out) {
    // Lagrange basis on the points 0..d:  L_t(X) = prod_{s != t} (X - s) / (t - s); column t of V^-1 holds the
    // coefficients of L_t, i.e. coef_j = sum_t vinv[j][t] * P(t).
    const uint32_t m = degree + 1;
    auto from_u32 = [](uint32_t v) {
        F x = F::zero();
        x.v[0] = v;
        return mul_portable(x, F::r_squared());
    };
    out.assign((size_t)m * m, F::zero());
    for (uint32_t t = 0; t < m; t++) {
        std::vector<F> poly(1, F::one());  // running product prod (X - s)
        F denom = F::one();
        for (uint32_t s = 0; s < m; s++) {
            if (s == t) continue;
            std::vector<F> next(poly.size() + 1, F::zero());
            F neg_s = sub_portable(F::zero(), from_u32(s));
            for (size_t k = 0; k < poly.size(); k++) {
                next[k + 1] = add_portable(next[k + 1], poly[k]);
                next[k] = add_portable(next[k], mul_portable(poly[k], neg_s));
            }
            poly.swap(next);
            F diff = sub_portable(from_u32(t), from_u32(s));
            denom = mul_portable(denom, diff);
        }
        // host-side inverse of a tiny constant (degree <= 15): Fermat on the portable path
        F dinv = F::one();
        {
            F base = denom;
            for (int i = 0; i < 8; i++) {
                uint32_t e = F::modulus_limb(i) - (i == 0 ? 2u : 0u);
                for (int bit = 0; bit < 32; bit++) {
                    if ((e >> bit) & 1) dinv = mul_portable(dinv, base);
                    base = mul_portable(base, base);
                }
            }
        }
        for (uint32_t j = 0; j < m; j++) out[(size_t)j * m + t] = mul_portable(poly[j], dinv);
    }
}

// per-call tables (column pointer arrays, challenge table, blend coefficients) live in the launching stream's
// scratch: two streams evaluating the same program never share them
static int args_reserve(cudaStream_t st, size_t bytes, char** out) {
    Scratch& a = ws_slot(st, WS_EXPR_ARGS);
    SB_TRY(a.reserve(bytes < 4096 ? 4096 : bytes));
    *out = (char*)a.ptr;
    return SB_OK;
}

template <class F>
static int fill_args(sb_prog* prog, sb_columns* cols, const void* const* adv1, const void* const* adv2, size_t nfv,
                     EvalArgs<F>& A) {
    A.ops = (const DevOp*)prog->d_ops;
    A.num_ops = prog->num_ops;
    A.num_slots = prog->num_slots;
    A.result_slot = prog->result_slot;
    A.constants = (const F*)prog->d_constants;
    A.rotations = (const int32_t*)prog->d_rotations;
    A.cols.log_rows = cols->log_rows;
    A.cols.num_selectors = cols->num_selectors;
    A.cols.num_fixed = cols->num_fixed;
    A.cols.selectors = (const uint8_t* const*)cols->d_ptrs;
    A.cols.fixed = (const void* const*)((char*)cols->d_ptrs + sizeof(void*) * cols->num_selectors);
    A.num_fold_vars = (uint32_t)nfv;
    (void)adv1;
    (void)adv2;
    return SB_OK;
}

// validates that every column the program touches exists
static int validate_program(const sb_prog* prog, const sb_columns* cols, size_t nfv, bool have_adv2, size_t num_challenges) {
    const size_t base = (size_t)cols->num_selectors + cols->num_fixed;
    for (uint32_t idx : prog->poly_indices) {
        if (idx < base) continue;
        size_t a = idx - base;
        if (a >= nfv * (have_adv2 ? 2 : 1)) {
            set_error("expression reads column variable %u but only %zu fold variables were supplied (ColumnVariableIndexOutOfBoundary)", idx, nfv);
            return SB_ERR_ARG;
        }
    }
    for (uint32_t idx : prog->fixed_indices) {
        if (idx >= cols->num_fixed) {   // graph_evaluator.rs:104-110 -> eval::Error::ColumnVariableIndexOutOfBoundary
            set_error("expression reads fixed column %u but the structure has %u (ColumnVariableIndexOutOfBoundary)", idx, cols->num_fixed);
            return SB_ERR_ARG;
        }
    }
    if (prog->uses_challenge && prog->max_challenge >= num_challenges) {
        set_error("challenge index out of boundary: %u (have %zu)", prog->max_challenge, num_challenges);
        return SB_ERR_ARG;
    }
    return SB_OK;
}

// ------------------------------------------------------------------------------------------------
// straight-line cross-term kernels, compiled at run time (NVRTC)
// ------------------------------------------------------------------------------------------------
// The interpreter above pays, per calculation, a 16-byte op fetch, the decode, two operand-kind switches and a
// shared-memory round trip of every intermediate (ncu: issue-active 25-31 %, i.e. latency-bound, not pipe-bound).
// A compiled GraphEvaluator program is a fixed straight-line sequence, so for the cross terms (the hot use) the
// library generates CUDA source for the program -- every calculation one statement on register-resident values,
// every distinct leaf (column, rotation) loaded and blended ONCE -- compiles it with NVRTC for sm_100a and keeps
// the kernel per (program, degree, column layout).  Same lazy-domain field operations (field.cuh is compiled in
// verbatim), same evaluation points 0..d, same inverse Vandermonde: bit-identical outputs.  SB_EXPR_JIT=0 keeps
// the interpreter (also the fallback if the run-time compiler is unavailable).
struct JitArgs {
    const void* const* fixed;
    const uint8_t* const* selectors;
    const void* const* adv1;
    const void* const* adv2;
    const void* constants;
    const void* challenges;   // [(d+1)][num_challenges]
    const void* vinv;         // [(d+1)][(d+1)]
    void* out;                // [d][n]
    uint32_t n, rows_per_block, row0, row_end;   // this launch covers rows [row0, row_end)
};
struct JitEvalArgs {          // plain evaluation (GraphEvaluator::evaluate per row), optionally on a Lagrange blend of traces
    const void* const* fixed;
    const uint8_t* const* selectors;
    const void* const* adv1;
    const void* const* adv2;
    const void* constants;
    const void* challenges;        // [evaluations][num_challenges]
    const void* const* blend_cols; // [num_blend][num_fold_vars]
    const void* blend_coef;        // [evaluations][num_blend]
    void* out;                     // [n]
    uint32_t n, t, row0_only;
};

}  // namespace sb

struct sb_jit_entry {
    uint32_t degree;        // cross terms: folding degree d (d + 1 evaluation points per row); 0: plain evaluation kernel
    uint32_t num_blend;     // plain evaluation: traces blended per fold variable (Protogalaxy), 0 = none
    uint32_t num_selectors, num_fixed, nfv, nch;
    cudaLibrary_t lib;
    cudaKernel_t fn;
    bool ok;
};

namespace sb {

static const char FIELD_SRC[] =
#include "field_src.inc"
    ;

static std::string jit_source(int field, const std::vector<DevOp>& ops, const std::vector<int32_t>& rots, uint32_t degree, uint32_t num_sel,
                              uint32_t num_fixed, uint32_t nfv, uint32_t nch, uint32_t num_blend = 0) {
    const bool eval_mode = degree == 0;
    std::string s;
    s.reserve(1 << 16);
    s += "typedef unsigned char uint8_t;\ntypedef unsigned short uint16_t;\ntypedef unsigned int uint32_t;\ntypedef unsigned long long uint64_t;\n"
         "typedef signed char int8_t;\ntypedef short int16_t;\ntypedef int int32_t;\ntypedef long long int64_t;\n";
    s += FIELD_SRC;
    s += "\nusing namespace sb;\ntypedef ";
    s += field == FIELD_FR ? "Fr" : "Fq";
    s += " F;\n";
    // Long calculation lists (the gate-scaling circuits: 600-1200 calculations): the product is CALLED, not inlined -- the
    // code stays a few hundred KB and compiles in seconds instead of minutes
    static const size_t inline_max_ops = []() {
        const char* e = getenv("SB_EXPR_JIT_INLINE_MAX_OPS");
        return e ? (size_t)atol(e) : (size_t)400;
    }();
    const bool call_form = ops.size() > inline_max_ops;
    const std::string mul_name = call_form ? "mul_c(" : "mul_lazy(";
    if (call_form) s += "__device__ __noinline__ F mul_c(F a, F b) { return mul_lazy(a, b); }\n";
    s += "struct JitArgs { const void* const* fixed; const uint8_t* const* selectors; const void* const* adv1; const void* const* adv2; const void* constants;\n"
         "  const void* challenges; const void* vinv; void* out; uint32_t n, rows_per_block, row0, row_end; };\n"
         "__device__ __forceinline__ F ld(const void* p, uint32_t i) { F r; const uint4* s = reinterpret_cast<const uint4*>(p) + 2 * (size_t)i; uint4* d = reinterpret_cast<uint4*>(&r);\n"
         "  d[0] = __ldg(s); d[1] = __ldg(s + 1); return r; }\n"
         "__device__ __forceinline__ void st(void* p, size_t i, const F& v) { uint4* d = reinterpret_cast<uint4*>(p) + 2 * i; const uint4* s = reinterpret_cast<const uint4*>(&v); d[0] = s[0]; d[1] = s[1]; }\n";
    s += "struct JitEvalArgs { const void* const* fixed; const uint8_t* const* selectors; const void* const* adv1; const void* const* adv2; const void* constants;\n"
         "  const void* challenges; const void* const* blend_cols; const void* blend_coef; void* out; uint32_t n, t, row0_only; };\n";
    const uint32_t m = degree + 1;
    char buf[768];
    if (eval_mode)
        s += "extern \"C\" __global__ void __launch_bounds__(128) sb_ct(JitEvalArgs A) {\n"
             "  const uint32_t out_row = blockIdx.x * blockDim.x + threadIdx.x, mask = A.n - 1u, t = A.t;\n  (void)mask; (void)t;\n"
             "  if (out_row >= A.n) return;\n"
             "  const uint32_t row = A.row0_only ? 0u : out_row;   // `index & 2^k` leaf addressing (src/plonk/mod.rs:714, SURVEY F4)\n  {\n";
    else {
    snprintf(buf, sizeof(buf),
             "extern \"C\" __global__ void __launch_bounds__(128) sb_ct(JitArgs A) {\n"
             "  extern __shared__ uint4 sm_[];\n  F* exch = reinterpret_cast<F*>(sm_);\n"
             "  const uint32_t m = %uu, r_in = threadIdx.x / m, t = threadIdx.x - r_in * m;\n"
             "  const uint32_t row = A.row0 + blockIdx.x * A.rows_per_block + r_in, mask = A.n - 1u;\n  (void)mask;\n"
             "  const bool live = r_in < A.rows_per_block && row < A.row_end;\n  if (live) {\n", m);
    s += buf;
    }
    // leaves are materialised at their first use
    std::vector<std::string> leaf_keys, leaf_names;
    auto leaf = [&](uint32_t kind, uint32_t v) -> std::string {
        snprintf(buf, sizeof(buf), "%u:%u", kind, v);
        const std::string key = buf;
        for (size_t i = 0; i < leaf_keys.size(); i++)
            if (leaf_keys[i] == key) return leaf_names[i];
        snprintf(buf, sizeof(buf), "l%zu", leaf_keys.size());
        const std::string name = buf;
        std::string def;
        if (kind == VS_CONSTANT) {
            snprintf(buf, sizeof(buf), "    const F %s = ld(A.constants, %uu);\n", name.c_str(), v);
            def = buf;
        } else if (kind == VS_CHALLENGE) {
            snprintf(buf, sizeof(buf), "    const F %s = ld(A.challenges, t * %uu + %uu);\n", name.c_str(), nch, v);
            def = buf;
        } else {
            const uint32_t index = v & 0xffffffu;
            const int32_t rot = rots.empty() ? 0 : rots[v >> 24];
            char rexpr[64];
            if (rot == 0) snprintf(rexpr, sizeof(rexpr), "row");
            else snprintf(rexpr, sizeof(rexpr), "((uint32_t)((int32_t)row + (%d)) & mask)", rot);   // get_rotation_idx, graph_evaluator.rs:51-53
            if (kind == VS_FIXED) {
                snprintf(buf, sizeof(buf), "    const F %s = ld(A.fixed[%u], %s);\n", name.c_str(), index, rexpr);
                def = buf;
            } else if (index < num_sel) {
                snprintf(buf, sizeof(buf), "    const F %s = A.selectors[%u][%s] ? F::one() : F::zero();\n", name.c_str(), index, rexpr);
                def = buf;
            } else if (index < num_sel + num_fixed) {
                snprintf(buf, sizeof(buf), "    const F %s = ld(A.fixed[%u], %s);\n", name.c_str(), index - num_sel, rexpr);
                def = buf;
            } else {
                const uint32_t a = index - num_sel - num_fixed;
                if (eval_mode && num_blend) {   // Lagrange blend of the traces (FoldedWitness::new, poly/folded_witness.rs:66-143), ONCE per leaf
                    snprintf(buf, sizeof(buf), "    F %s = mul(ld(A.blend_coef, t * %uu), ld(A.blend_cols[%u], %s));\n", name.c_str(), num_blend, a, rexpr);
                    def = buf;
                    for (uint32_t j = 1; j < num_blend; j++) {
                        snprintf(buf, sizeof(buf), "    %s = add(%s, mul(ld(A.blend_coef, t * %uu + %uu), ld(A.blend_cols[%u], %s)));\n", name.c_str(), name.c_str(), num_blend, j,
                                 j * nfv + a, rexpr);
                        def += buf;
                    }
                } else if (eval_mode && a < nfv) {
                    snprintf(buf, sizeof(buf), "    const F %s = ld(A.adv1[%u], %s);\n", name.c_str(), a, rexpr);
                    def = buf;
                } else if (a >= nfv) {   // explicit second-instance variable
                    snprintf(buf, sizeof(buf), "    const F %s = ld(A.adv2[%u], %s);\n", name.c_str(), a - nfv, rexpr);
                    def = buf;
                } else {          // w1 + t * w2 at this thread's evaluation point (PlonkEvalDomain::eval_advice_var with the folded instance)
                    snprintf(buf, sizeof(buf), "    F %s = ld(A.adv1[%u], %s);\n    { const F w = ld(A.adv2[%u], %s); for (uint32_t k = 0; k < t; k++) %s = add(%s, w); }\n",
                             name.c_str(), a, rexpr, a, rexpr, name.c_str(), name.c_str());
                    def = buf;
                }
            }
        }
        s += def;
        leaf_keys.push_back(key);
        leaf_names.push_back(name);
        return name;
    };
    std::vector<std::string> slot_name;   // current SSA name held by each slot
    std::string last = "F::zero()";
    for (size_t i = 0; i < ops.size(); i++) {
        const DevOp& o = ops[i];
        const uint32_t op = o.code & 0xff, ak = (o.code >> 8) & 0xff, bk = (o.code >> 16) & 0xff;
        auto operand = [&](uint32_t kind, uint32_t v) -> std::string {
            if (kind == VS_INTERMEDIATE) return v < slot_name.size() ? slot_name[v] : std::string("F::zero()");
            return leaf(kind, v);
        };
        const std::string a = operand(ak, o.a);
        std::string expr;
        switch (op) {
            case OP_ADD: expr = "add_lazy(" + a + ", " + operand(bk, o.b) + ")"; break;
            case OP_SUB: expr = "sub_lazy(" + a + ", " + operand(bk, o.b) + ")"; break;
            case OP_MUL: expr = mul_name + a + ", " + operand(bk, o.b) + ")"; break;
            case OP_SQUARE: expr = mul_name + a + ", " + a + ")"; break;
            case OP_DOUBLE: expr = "dbl_lazy(" + a + ")"; break;
            case OP_NEGATE: expr = "neg_lazy(" + a + ")"; break;
            default: expr = a; break;  // OP_STORE
        }
        snprintf(buf, sizeof(buf), "v%zu", i);
        const std::string name = buf;
        s += "    const F " + name + " = " + expr + ";\n";
        if (o.dst >= slot_name.size()) slot_name.resize(o.dst + 1);
        slot_name[o.dst] = name;
        last = name;
    }
    if (eval_mode) {
        s += "    st(A.out, out_row, canon(" + last + "));\n  }\n}\n";
        return s;
    }
    s += "    exch[r_in * m + t] = " + last + ";\n  }\n  __syncthreads();\n";
    snprintf(buf, sizeof(buf),
             "  if (live && t + 1u <= %uu) {\n    const uint32_t j = t + 1u;\n    F acc = F::zero();\n"
             "    for (uint32_t q = 0; q < m; q++) acc = add_lazy(acc, mul_lazy(ld(A.vinv, j * m + q), exch[r_in * m + q]));\n"
             "    st(A.out, (size_t)(j - 1u) * A.n + row, canon(acc));\n  }\n}\n", degree);
    s += buf;
    return s;
}

// source -> CUBIN for sm_100a; `log` receives the compiler's messages
static int jit_compile_cubin(const std::string& src, std::vector<char>& cubin, std::string& log) {
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "sb_ct.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        log = "nvrtcCreateProgram failed";
        return SB_ERR_CUDA;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--ptxas-options=-v"};
    const nvrtcResult r = nvrtcCompileProgram(prog, 4, opts);
    size_t ls = 0;
    nvrtcGetProgramLogSize(prog, &ls);
    if (ls > 1) {
        log.resize(ls);
        nvrtcGetProgramLog(prog, &log[0]);
    }
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return SB_ERR_CUDA;
    }
    size_t cs = 0;
    nvrtcGetCUBINSize(prog, &cs);
    cubin.resize(cs);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    return SB_OK;
}

static int g_jit_on = []() {
    const char* e = getenv("SB_EXPR_JIT");
    return (!e || atoi(e) != 0) ? 1 : 0;
}();
static bool jit_enabled() { return g_jit_on != 0; }
// Straight-line code grows with the calculation list (every product is inlined: ~180 instructions): beyond a few hundred
// calculations the compile takes minutes, the kernel spills and no longer fits the instruction cache (measured with NVRTC
// here: 136 calculations 5 s / 128 registers, 349: 17 s / 168 registers, 619: 48 s / 255 registers + spills).  Above
// SB_EXPR_JIT_INLINE_MAX_OPS (400) calculations the generated code CALLS the product (jit_source: 619 calculations 4 s / 205
// registers / no stack, 1159: 18 s / 255 registers / 280 bytes of spills); above SB_EXPR_JIT_MAX_OPS the interpreter kernel runs.
static size_t g_jit_max_ops = []() {
    const char* e = getenv("SB_EXPR_JIT_MAX_OPS");
    return e ? (size_t)atol(e) : (size_t)2000;
}();

// the compiled kernel for (prog, degree, layout), built on first use; nullptr -> use the interpreter
static sb_jit_entry* jit_lookup(sb_prog* prog, uint32_t degree, const sb_columns* cols, uint32_t nfv, uint32_t nch, uint32_t num_blend = 0) {
    if (!jit_enabled() || prog->h_ops.size() > g_jit_max_ops) return nullptr;
    for (sb_jit_entry* e : prog->jit)
        if (e->degree == degree && e->num_blend == num_blend && e->num_selectors == cols->num_selectors && e->num_fixed == cols->num_fixed && e->nfv == nfv &&
            e->nch == nch)
            return e->ok ? e : nullptr;
    sb_jit_entry* e = new (std::nothrow) sb_jit_entry();
    if (!e) return nullptr;
    *e = sb_jit_entry{degree, num_blend, cols->num_selectors, cols->num_fixed, nfv, nch, nullptr, nullptr, false};
    prog->jit.push_back(e);
    std::vector<char> cubin;
    std::string log;
    const std::string src = jit_source(prog->field, prog->h_ops, prog->h_rotations, degree, cols->num_selectors, cols->num_fixed, nfv, nch, num_blend);
    if (jit_compile_cubin(src, cubin, log) != SB_OK) {
        fprintf(stderr, "libsirius_b200: run-time compilation of the cross-term kernel failed, using the interpreter:\n%.2000s\n", log.c_str());
        return nullptr;
    }
    if (cudaLibraryLoadData(&e->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
        cudaLibraryGetKernel(&e->fn, e->lib, "sb_ct") != cudaSuccess) {
        cudaGetLastError();
        fprintf(stderr, "libsirius_b200: loading the compiled cross-term kernel failed, using the interpreter\n");
        return nullptr;
    }
    cudaFuncSetAttribute((const void*)e->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    e->ok = true;
    return e;
}

template <class F>
static int eval_enqueue(sb_prog* prog, sb_columns* cols, const void* const* h_adv1, const void* const* h_adv2, size_t nfv,
                        const uint64_t* challenges, size_t num_challenges, void* d_out, cudaStream_t st) {
    SB_TRY(validate_program(prog, cols, nfv, h_adv2 != nullptr, num_challenges));
    const size_t ptr_bytes = align_up(sizeof(void*) * nfv * 2, 32);
    const size_t ch_bytes = align_up(32 * (num_challenges ? num_challenges : 1), 32);
    char* d = nullptr;
    SB_TRY(args_reserve(st, ptr_bytes + ch_bytes, &d));
    if (nfv) {
        SB_CUDA_TRY(cudaMemcpyAsync(d, h_adv1, sizeof(void*) * nfv, cudaMemcpyHostToDevice, st));
        if (h_adv2) SB_CUDA_TRY(cudaMemcpyAsync(d + sizeof(void*) * nfv, h_adv2, sizeof(void*) * nfv, cudaMemcpyHostToDevice, st));
    }
    if (num_challenges) SB_CUDA_TRY(cudaMemcpyAsync(d + ptr_bytes, challenges, 32 * num_challenges, cudaMemcpyHostToDevice, st));
    EvalArgs<F> A;
    SB_TRY(fill_args<F>(prog, cols, h_adv1, h_adv2, nfv, A));
    A.adv1 = (const F* const*)d;
    A.adv2 = h_adv2 ? (const F* const*)(d + sizeof(void*) * nfv) : nullptr;
    A.challenges = (const F*)(d + ptr_bytes);
    A.num_challenges = (uint32_t)num_challenges;
    const uint32_t n = 1u << cols->log_rows;
    const uint32_t threads = n < (uint32_t)EXPR_THREADS ? n : EXPR_THREADS;
    if (sb_jit_entry* je = jit_lookup(prog, 0, cols, (uint32_t)nfv, (uint32_t)num_challenges, 0)) {
        JitEvalArgs ja;
        ja.fixed = A.cols.fixed;
        ja.selectors = A.cols.selectors;
        ja.adv1 = (const void* const*)A.adv1;
        ja.adv2 = (const void* const*)A.adv2;
        ja.constants = A.constants;
        ja.challenges = A.challenges;
        ja.blend_cols = nullptr;
        ja.blend_coef = nullptr;
        ja.out = d_out;
        ja.n = n;
        ja.t = 0;
        ja.row0_only = 0;
        void* kargs[1] = {&ja};
        SB_CUDA_TRY(cudaLaunchKernel((const void*)je->fn, dim3((n + threads - 1) / threads), dim3(threads), kargs, 0, st));
        count_launch();
        return SB_OK;
    }
    const size_t smem = (size_t)(prog->num_slots ? prog->num_slots : 1) * 32 * threads;
    if (smem > 200 * 1024) {
        set_error("expression needs %u live intermediates: too many for shared memory", prog->num_slots);
        return SB_ERR_ARG;
    }
    SB_CUDA_TRY(cudaFuncSetAttribute(k_expr_eval<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    A.num_blend = 0;
    A.blend_cols = nullptr;
    A.blend_coef = nullptr;
    k_expr_eval<F><<<(n + threads - 1) / threads, threads, smem, st>>>(A, 0u, 0, (F*)d_out);
    SB_KERNEL_CHECK();
    return SB_OK;
}

// Protogalaxy leaves: for every blend b (a Lagrange combination of `num_traces` witnesses + its folded challenge
// vector) and every gate g, leaves[b][g * 2^k + row] = gate_g(row) ; padding up to 2^log_leaves is zero.
template <class F>
static int pg_leaves_enqueue(sb_prog* const* gates, size_t num_gates, sb_columns* cols, const void* const* h_cols_tables,
                             size_t num_traces, size_t nfv, const uint64_t* coef, const uint64_t* challenges,
                             size_t num_challenges, size_t num_blends, int row_mode_compat, uint32_t log_leaves, void* d_leaves,
                             cudaStream_t st) {
    const uint32_t n = 1u << cols->log_rows;
    const size_t leaves = (size_t)1 << log_leaves;
    if ((size_t)n * num_gates > leaves) {
        set_error("sb_pg_leaves: 2^%u leaves cannot hold %zu gates x %u rows", log_leaves, num_gates, n);
        return SB_ERR_ARG;
    }
    ProfScope ps(st, PROF_PG, leaves * num_blends);
    SB_CUDA_TRY(cudaMemsetAsync(d_leaves, 0, leaves * num_blends * 32, st));
    const size_t ptr_bytes = align_up(sizeof(void*) * nfv * num_traces, 32);
    const size_t coef_bytes = align_up(32 * num_blends * num_traces, 32);
    const size_t ch_bytes = align_up(32 * num_blends * (num_challenges ? num_challenges : 1), 32);
    for (size_t g = 0; g < num_gates; g++) {
        sb_prog* prog = gates[g];
        SB_TRY(validate_program(prog, cols, nfv, false, num_challenges));
        char* d = nullptr;
        SB_TRY(args_reserve(st, ptr_bytes + coef_bytes + ch_bytes, &d));
        if (nfv) SB_CUDA_TRY(cudaMemcpyAsync(d, h_cols_tables, sizeof(void*) * nfv * num_traces, cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(cudaMemcpyAsync(d + ptr_bytes, coef, 32 * num_blends * num_traces, cudaMemcpyHostToDevice, st));
        if (num_challenges) SB_CUDA_TRY(cudaMemcpyAsync(d + ptr_bytes + coef_bytes, challenges, 32 * num_blends * num_challenges, cudaMemcpyHostToDevice, st));
        EvalArgs<F> A;
        SB_TRY(fill_args<F>(prog, cols, nullptr, nullptr, nfv, A));
        A.adv1 = nullptr;
        A.adv2 = nullptr;
        A.challenges = (const F*)(d + ptr_bytes + coef_bytes);
        A.num_challenges = (uint32_t)num_challenges;
        A.num_blend = (uint32_t)num_traces;
        A.blend_cols = (const F* const*)d;
        A.blend_coef = (const F*)(d + ptr_bytes);
        const uint32_t threads = n < (uint32_t)EXPR_THREADS ? n : EXPR_THREADS;
        if (sb_jit_entry* je = jit_lookup(prog, 0, cols, (uint32_t)nfv, (uint32_t)num_challenges, (uint32_t)num_traces)) {
            for (size_t b = 0; b < num_blends; b++) {
                JitEvalArgs ja;
                ja.fixed = A.cols.fixed;
                ja.selectors = A.cols.selectors;
                ja.adv1 = nullptr;
                ja.adv2 = nullptr;
                ja.constants = A.constants;
                ja.challenges = A.challenges;
                ja.blend_cols = (const void* const*)A.blend_cols;
                ja.blend_coef = A.blend_coef;
                ja.out = (F*)d_leaves + b * leaves + g * (size_t)n;
                ja.n = n;
                ja.t = (uint32_t)b;
                ja.row0_only = row_mode_compat ? 1u : 0u;
                void* kargs[1] = {&ja};
                SB_CUDA_TRY(cudaLaunchKernel((const void*)je->fn, dim3((n + threads - 1) / threads), dim3(threads), kargs, 0, st));
                count_launch();
            }
            continue;
        }
        const size_t smem = (size_t)(prog->num_slots ? prog->num_slots : 1) * 32 * threads;
        if (smem > 200 * 1024) {
            set_error("expression needs %u live intermediates: too many for shared memory", prog->num_slots);
            return SB_ERR_ARG;
        }
        SB_CUDA_TRY(cudaFuncSetAttribute(k_expr_eval<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (size_t b = 0; b < num_blends; b++) {
            F* out = (F*)d_leaves + b * leaves + g * (size_t)n;
            k_expr_eval<F><<<(n + threads - 1) / threads, threads, smem, st>>>(A, (uint32_t)b, row_mode_compat, out);
            SB_KERNEL_CHECK();
        }
    }
    return SB_OK;
}

template <class F>
static int cross_terms_enqueue(sb_prog* prog, uint32_t degree, sb_columns* cols, const void* const* h_adv1,
                               const void* const* h_adv2, size_t nfv, const uint64_t* ch1, const uint64_t* ch2,
                               size_t num_challenges, void* d_out, cudaStream_t st, size_t row_begin = 0, size_t row_count = ~(size_t)0) {
    if (degree < 1 || degree > (uint32_t)EXPR_MAX_DEGREE) {
        set_error("sb_cross_terms: degree %u out of range [1,%d]", degree, EXPR_MAX_DEGREE);
        return SB_ERR_ARG;
    }
    SB_TRY(validate_program(prog, cols, nfv, false, num_challenges));
    const uint32_t m = degree + 1;
    const size_t ptr_bytes = align_up(sizeof(void*) * nfv * 2, 32);
    const size_t ch_bytes = (size_t)32 * m * (num_challenges ? num_challenges : 1);
    const size_t vinv_bytes = (size_t)32 * m * m;
    char* d = nullptr;
    SB_TRY(args_reserve(st, ptr_bytes + ch_bytes, &d));
    if (prog->vinv_degree != degree || !prog->d_vinv) {
        std::vector<F> vinv;
        host_vandermonde_inverse<F>(degree, vinv);
        if (prog->d_vinv) cudaFree(prog->d_vinv);
        prog->d_vinv = nullptr;
        SB_CUDA_TRY(cudaMalloc(&prog->d_vinv, vinv_bytes));
        SB_CUDA_TRY(cudaMemcpyAsync(prog->d_vinv, vinv.data(), vinv_bytes, cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(cudaStreamSynchronize(st));
        prog->vinv_degree = degree;
    }
    // challenge table: c1 + t*c2 for t = 0..degree, built on the host with the portable field code
    std::vector<F> table((size_t)m * (num_challenges ? num_challenges : 1));
    for (size_t i = 0; i < num_challenges; i++) {
        F c1, c2;
        memcpy(c1.v, ch1 + 4 * i, 32);
        memcpy(c2.v, ch2 + 4 * i, 32);
        F cur = c1;
        for (uint32_t t = 0; t < m; t++) {
            table[(size_t)t * num_challenges + i] = cur;
            cur = add_portable(cur, c2);
        }
    }
    if (nfv) {
        SB_CUDA_TRY(cudaMemcpyAsync(d, h_adv1, sizeof(void*) * nfv, cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(cudaMemcpyAsync(d + sizeof(void*) * nfv, h_adv2, sizeof(void*) * nfv, cudaMemcpyHostToDevice, st));
    }
    if (num_challenges) SB_CUDA_TRY(cudaMemcpyAsync(d + ptr_bytes, table.data(), 32 * table.size(), cudaMemcpyHostToDevice, st));
    // the staging vectors die at return: pageable cudaMemcpyAsync has already copied them out
    EvalArgs<F> A;
    SB_TRY(fill_args<F>(prog, cols, h_adv1, h_adv2, nfv, A));
    A.adv1 = (const F* const*)d;
    A.adv2 = (const F* const*)(d + sizeof(void*) * nfv);
    A.challenges = (const F*)(d + ptr_bytes);
    A.num_challenges = (uint32_t)num_challenges;
    A.num_blend = 0;
    A.blend_cols = nullptr;
    A.blend_coef = nullptr;
    const uint32_t n = 1u << cols->log_rows;
    if (row_count == ~(size_t)0) row_count = n - (row_begin < n ? row_begin : n);
    if (row_begin > n || row_count > n - row_begin) {
        set_error("sb_cross_terms: rows [%zu, %zu) outside the table of %u rows", row_begin, row_begin + row_count, n);
        return SB_ERR_ARG;
    }
    if (!row_count) return SB_OK;
    const uint32_t row0 = (uint32_t)row_begin, row_end = (uint32_t)(row_begin + row_count), nr = (uint32_t)row_count;
    uint32_t rows_per_block = (uint32_t)EXPR_THREADS / m;
    if (rows_per_block > nr) rows_per_block = nr;
    const uint32_t threads = (rows_per_block * m + 31) / 32 * 32;
    const size_t smem = (size_t)(prog->num_slots ? prog->num_slots : 1) * 32 * threads + (size_t)rows_per_block * m * 32;
    if (smem > 200 * 1024) {
        set_error("expression needs %u live intermediates: too many for shared memory", prog->num_slots);
        return SB_ERR_ARG;
    }
    if (sb_jit_entry* je = jit_lookup(prog, degree, cols, (uint32_t)nfv, (uint32_t)num_challenges)) {
        JitArgs ja;
        ja.fixed = A.cols.fixed;
        ja.selectors = A.cols.selectors;
        ja.adv1 = (const void* const*)A.adv1;
        ja.adv2 = (const void* const*)A.adv2;
        ja.constants = A.constants;
        ja.challenges = A.challenges;
        ja.vinv = prog->d_vinv;
        ja.out = d_out;
        ja.n = n;
        ja.rows_per_block = 128u / m;
        if (ja.rows_per_block > nr) ja.rows_per_block = nr;
        ja.row0 = row0;
        ja.row_end = row_end;
        void* kargs[1] = {&ja};
        const unsigned jblocks = (nr + ja.rows_per_block - 1) / ja.rows_per_block;
        ProfScope ps(st, PROF_CROSS_TERMS, nr);
        SB_CUDA_TRY(cudaLaunchKernel((const void*)je->fn, dim3(jblocks), dim3(128), kargs, (size_t)ja.rows_per_block * m * 32, st));
        count_launch();
        return SB_OK;
    }
    SB_CUDA_TRY(cudaFuncSetAttribute(k_cross_terms<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    {
        ProfScope ps(st, PROF_CROSS_TERMS, nr);
        k_cross_terms<F><<<(nr + rows_per_block - 1) / rows_per_block, threads, smem, st>>>(A, degree, rows_per_block, row0, row_end, (const F*)prog->d_vinv, (F*)d_out);
        SB_KERNEL_CHECK();
    }
    return SB_OK;
}

// fold-variable index -> (round, column) : PlonkEvalDomain::eval_advice_var's index_map (src/plonk/eval.rs:170-204)
static int fold_var_location(size_t index, size_t num_advice, size_t num_lookup, size_t num_witness, size_t* round, size_t* col) {
    if (index < num_advice) {
        *round = 0;
        *col = index;
        return SB_OK;
    }
    size_t li = (index - num_advice) / 5, sub = (index - num_advice) % 5;
    bool first = sub < 3;
    if (!first) sub -= 3;
    if (num_witness == 2) {
        if (first) { *round = 0; *col = num_advice + li * 3 + sub; }
        else { *round = 1; *col = li * 2 + sub; }
        return SB_OK;
    }
    if (num_witness == 3) {
        if (first) { *round = 1; *col = li * 3 + sub; }
        else { *round = 2; *col = li * 2 + sub; }
        return SB_OK;
    }
    set_error("Invalid witness index. num_witness: %zu, num_advice: %zu, num_lookup: %zu, index: %zu", num_witness, num_advice, num_lookup, index);
    return SB_ERR_ARG;
}


}  // namespace sb

extern "C" {

}  // extern "C"

namespace sb {

// host-only half of sb_expr_compile: validation, aliasing of leaf `Store`s, liveness-based slot assignment
struct Lowered {
    std::vector<DevOp> ops;
    uint32_t num_slots = 0, result_slot = 0, max_challenge = 0;
    bool uses_challenge = false;
    std::vector<uint32_t> poly_indices, fixed_indices;
};

static int lower_program(int field, const sb_calc* calcs, size_t n_calcs, size_t n_constants, size_t n_rotations, Lowered& L) {
    if (field != FIELD_FR && field != FIELD_FQ) {
        set_error("sb_expr_compile: unknown field %d", field);
        return SB_ERR_ARG;
    }
    // liveness: last reader of every intermediate
    std::vector<long> last_use(n_calcs, -1);
    std::vector<long> def_of(n_calcs, -1);  // target -> defining calc
    for (size_t i = 0; i < n_calcs; i++) {
        if (calcs[i].target >= n_calcs) {
            set_error("sb_expr_compile: target %u out of range", calcs[i].target);
            return SB_ERR_ARG;
        }
        def_of[calcs[i].target] = (long)i;
    }
    auto check_src = [&](uint32_t kind, uint32_t index, uint32_t rot, size_t at) -> int {
        switch (kind) {
            case VS_CONSTANT: if (index >= n_constants) { set_error("calc %zu: constant %u out of range", at, index); return SB_ERR_ARG; } break;
            case VS_INTERMEDIATE:
                if (index >= n_calcs || def_of[index] < 0 || def_of[index] >= (long)at) { set_error("calc %zu: intermediate %u used before definition", at, index); return SB_ERR_ARG; }
                last_use[index] = (long)at;
                break;
            case VS_FIXED: case VS_POLY:
                if (rot >= n_rotations || index >= (1u << 24) || n_rotations > 255) { set_error("calc %zu: rotation/column out of range", at); return SB_ERR_ARG; }
                break;
            case VS_CHALLENGE: break;
            default: set_error("calc %zu: unknown value source %u", at, kind); return SB_ERR_ARG;
        }
        return SB_OK;
    };
    for (size_t i = 0; i < n_calcs; i++) {
        const sb_calc& c = calcs[i];
        if (c.opcode == OP_HORNER) {
            set_error("sb_expr_compile: Calculation::Horner is never built by GraphEvaluator::add_expression and is not supported");
            return SB_ERR_ARG;
        }
        if (c.opcode > OP_STORE) {
            set_error("sb_expr_compile: unknown opcode %u", c.opcode);
            return SB_ERR_ARG;
        }
        SB_TRY(check_src(c.a_kind, c.a_index, c.a_rot, i));
        if (c.opcode <= OP_MUL) SB_TRY(check_src(c.b_kind, c.b_index, c.b_rot, i));
    }
    // `Store(column | challenge | constant)` calculations (one per distinct leaf, graph_evaluator.rs:264-276) are
    // not given a slot: their readers fetch the source directly (coalesced / broadcast loads).  This halves the
    // live-slot count, i.e. doubles to quadruples the resident warps of the interpreter kernels.
    struct Src { uint32_t kind, index, rot; };
    std::vector<int> is_alias(n_calcs, 0);
    std::vector<Src> alias_src(n_calcs);
    for (size_t i = 0; i + 1 < n_calcs; i++) {  // the last calculation always materialises (it is the result)
        const sb_calc& c = calcs[i];
        if (c.opcode == OP_STORE && c.a_kind != VS_INTERMEDIATE) {
            is_alias[c.target] = 1;
            alias_src[c.target] = Src{c.a_kind, c.a_index, c.a_rot};
        }
    }
    auto resolve = [&](uint32_t kind, uint32_t index, uint32_t rot) -> Src {
        if (kind == VS_INTERMEDIATE && is_alias[index]) return alias_src[index];
        return Src{kind, index, rot};
    };
    // recompute liveness on resolved operands
    std::fill(last_use.begin(), last_use.end(), -1);
    for (size_t i = 0; i < n_calcs; i++) {
        const sb_calc& c = calcs[i];
        if (is_alias[c.target] && i + 1 < n_calcs && c.opcode == OP_STORE && c.a_kind != VS_INTERMEDIATE) continue;
        Src a = resolve(c.a_kind, c.a_index, c.a_rot);
        if (a.kind == VS_INTERMEDIATE) last_use[a.index] = (long)i;
        if (c.opcode <= OP_MUL) {
            Src bsrc = resolve(c.b_kind, c.b_index, c.b_rot);
            if (bsrc.kind == VS_INTERMEDIATE) last_use[bsrc.index] = (long)i;
        }
    }
    if (n_calcs) last_use[calcs[n_calcs - 1].target] = (long)n_calcs;  // the result stays live
    // slot assignment
    std::vector<uint32_t> slot_of(n_calcs, 0);
    std::vector<uint32_t> free_slots;
    uint32_t num_slots = 0;
    L.ops.reserve(n_calcs);
    auto enc = [&](const Src& v) -> uint32_t {
        if (v.kind == VS_INTERMEDIATE) return slot_of[v.index];
        if (v.kind == VS_POLY || v.kind == VS_FIXED) {
            if (v.kind == VS_POLY) L.poly_indices.push_back(v.index);
            else L.fixed_indices.push_back(v.index);
            return v.index | (v.rot << 24);
        }
        if (v.kind == VS_CHALLENGE) {
            L.uses_challenge = true;
            if (v.index > L.max_challenge) L.max_challenge = v.index;
        }
        return v.index;
    };
    for (size_t i = 0; i < n_calcs; i++) {
        const sb_calc& c = calcs[i];
        if (is_alias[c.target] && i + 1 < n_calcs) continue;
        const bool binary = c.opcode <= OP_MUL;
        Src a = resolve(c.a_kind, c.a_index, c.a_rot);
        Src bsrc = binary ? resolve(c.b_kind, c.b_index, c.b_rot) : Src{0, 0, 0};
        DevOp o;
        o.code = c.opcode | (a.kind << 8) | ((binary ? bsrc.kind : 0u) << 16);
        o.a = enc(a);
        o.b = binary ? enc(bsrc) : 0;
        // operands dying here release their slots before the destination is chosen
        if (a.kind == VS_INTERMEDIATE && last_use[a.index] == (long)i) free_slots.push_back(slot_of[a.index]);
        if (binary && bsrc.kind == VS_INTERMEDIATE && last_use[bsrc.index] == (long)i && !(a.kind == VS_INTERMEDIATE && a.index == bsrc.index))
            free_slots.push_back(slot_of[bsrc.index]);
        uint32_t sl;
        if (!free_slots.empty()) {
            sl = free_slots.back();
            free_slots.pop_back();
        } else {
            sl = num_slots++;
        }
        slot_of[c.target] = sl;
        o.dst = sl;
        L.ops.push_back(o);
        if (last_use[c.target] < 0) free_slots.push_back(sl);  // never read (dead value)
    }
    L.num_slots = num_slots;
    L.result_slot = n_calcs ? slot_of[calcs[n_calcs - 1].target] : 0;
    return SB_OK;
}

}  // namespace sb

extern "C" {

int sb_expr_compile(int field, const sb_calc* calcs, size_t n_calcs, const uint64_t* constants_mont, size_t n_constants,
                    const int32_t* rotations, size_t n_rotations, sb_prog_t* out) {
    if (!out || (!calcs && n_calcs) || (!constants_mont && n_constants) || (!rotations && n_rotations)) {
        set_error("sb_expr_compile: null argument");
        return SB_ERR_ARG;
    }
    Lowered L;
    SB_TRY(lower_program(field, calcs, n_calcs, n_constants, n_rotations, L));
    SB_TRY(ensure_runtime());
    sb_prog* p = new (std::nothrow) sb_prog();
    if (!p) return SB_ERR_OOM;
    const size_t n_ops = L.ops.size();
    p->field = field;
    p->num_ops = (uint32_t)n_ops;
    p->num_slots = L.num_slots;
    p->result_slot = L.result_slot;
    p->num_constants = (uint32_t)n_constants;
    p->num_rotations = (uint32_t)n_rotations;
    p->max_challenge = L.max_challenge;
    p->uses_challenge = L.uses_challenge;
    p->poly_indices = L.poly_indices;
    p->fixed_indices = L.fixed_indices;
    p->h_ops = L.ops;
    p->h_rotations.assign(rotations, rotations + n_rotations);
    p->d_ops = p->d_constants = p->d_rotations = nullptr;
    p->vinv_degree = 0;
    p->d_vinv = nullptr;
    const std::vector<DevOp>& ops = L.ops;
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaError_t e = cudaMalloc(&p->d_ops, sizeof(DevOp) * (n_ops ? n_ops : 1));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_constants, 32 * (n_constants ? n_constants : 1));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_rotations, 4 * (n_rotations ? n_rotations : 1));
    if (e == cudaSuccess && n_ops) e = cudaMemcpyAsync(p->d_ops, ops.data(), sizeof(DevOp) * n_ops, cudaMemcpyHostToDevice, rt.stream);
    if (e == cudaSuccess && n_constants) e = cudaMemcpyAsync(p->d_constants, constants_mont, 32 * n_constants, cudaMemcpyHostToDevice, rt.stream);
    if (e == cudaSuccess && n_rotations) e = cudaMemcpyAsync(p->d_rotations, rotations, 4 * n_rotations, cudaMemcpyHostToDevice, rt.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rt.stream);
    if (e != cudaSuccess) {
        set_error("sb_expr_compile: %s", cudaGetErrorString(e));
        sb_expr_free(p);
        return SB_ERR_CUDA;
    }
    *out = p;
    return SB_OK;
}

/* 1 (default): cross terms run on straight-line kernels compiled at run time; 0: on the interpreter.  Same results. */
void sb_expr_jit_enable(int on) { g_jit_on = on ? 1 : 0; }

/* Code generator + run-time compiler check that needs no device (NVRTC cross-compiles for sm_100a): lowers the calculation list,
 * generates the straight-line cross-term kernel for (degree, column layout) and compiles it.  log (may be NULL) receives the
 * compiler output incl. the ptxas resource line; *cubin_bytes the size of the image.  Used by the CPU tests. */
int sb_expr_jit_selftest(int field, const sb_calc* calcs, size_t n_calcs, size_t n_constants, const int32_t* rotations, size_t n_rotations, uint32_t degree,
                         uint32_t num_selectors, uint32_t num_fixed, uint32_t num_fold_vars, uint32_t num_challenges, char* log, size_t log_cap,
                         size_t* cubin_bytes) {
    uint32_t num_blend = 0;
    if (degree >= 0x100u) {   // plain-evaluation kernel on a blend of (degree >> 8) traces; degree & 0xff must be 0
        num_blend = degree >> 8;
        degree &= 0xffu;
    }
    if ((!calcs && n_calcs) || (!rotations && n_rotations) || degree > (uint32_t)EXPR_MAX_DEGREE) {
        set_error("sb_expr_jit_selftest: bad argument");
        return SB_ERR_ARG;
    }
    Lowered L;
    SB_TRY(lower_program(field, calcs, n_calcs, n_constants, n_rotations, L));
    std::vector<int32_t> rots(rotations, rotations + n_rotations);
    const std::string src = jit_source(field, L.ops, rots, degree, num_selectors, num_fixed, num_fold_vars, num_challenges, num_blend);
    std::vector<char> cubin;
    std::string out;
    const int rc = jit_compile_cubin(src, cubin, out);
    if (log && log_cap) {
        snprintf(log, log_cap, "%s", out.c_str());
    }
    if (cubin_bytes) *cubin_bytes = cubin.size();
    if (rc != SB_OK) set_error("sb_expr_jit_selftest: NVRTC failed: %.400s", out.c_str());
    return rc;
}

void sb_expr_free(sb_prog_t p) {
    if (!p) return;
    if (p->d_ops) cudaFree(p->d_ops);
    if (p->d_constants) cudaFree(p->d_constants);
    if (p->d_rotations) cudaFree(p->d_rotations);
    if (p->d_vinv) cudaFree(p->d_vinv);
    for (sb_jit_entry* e : p->jit) {
        if (e->lib) cudaLibraryUnload(e->lib);
        delete e;
    }
    delete p;
}

uint32_t sb_expr_num_slots(sb_prog_t p) { return p ? p->num_slots : 0; }
int sb_expr_field(sb_prog_t p) { return p ? p->field : -1; }
uint32_t sb_columns_log_rows(sb_columns_t c) { return c ? c->log_rows : 0; }

int sb_columns_register(int field, uint32_t log_rows, const uint8_t* const* selectors, size_t num_selectors,
                        const uint64_t* const* fixed, size_t num_fixed, sb_columns_t* out) {
    if (!out || (!selectors && num_selectors) || (!fixed && num_fixed) || log_rows > 30) {
        set_error("sb_columns_register: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    const size_t n = (size_t)1 << log_rows;
    const size_t sel_bytes = align_up(n, 32);
    const size_t total = sel_bytes * num_selectors + n * 32 * num_fixed;
    sb_columns* c = new (std::nothrow) sb_columns();
    if (!c) return SB_ERR_OOM;
    c->field = field;
    c->log_rows = log_rows;
    c->num_selectors = (uint32_t)num_selectors;
    c->num_fixed = (uint32_t)num_fixed;
    c->d_data = c->d_ptrs = nullptr;
    cudaError_t e = cudaMalloc(&c->d_data, total ? total : 32);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_ptrs, sizeof(void*) * (num_selectors + num_fixed + 1));
    std::vector<void*> ptrs(num_selectors + num_fixed + 1, nullptr);
    char* base = (char*)c->d_data;
    for (size_t i = 0; i < num_selectors && e == cudaSuccess; i++) {
        ptrs[i] = base + sel_bytes * i;
        e = cudaMemcpyAsync(ptrs[i], selectors[i], n, cudaMemcpyHostToDevice, rt.stream);
    }
    base += sel_bytes * num_selectors;
    for (size_t i = 0; i < num_fixed && e == cudaSuccess; i++) {
        ptrs[num_selectors + i] = base + n * 32 * i;
        e = cudaMemcpyAsync(ptrs[num_selectors + i], fixed[i], n * 32, cudaMemcpyHostToDevice, rt.stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_ptrs, ptrs.data(), sizeof(void*) * ptrs.size(), cudaMemcpyHostToDevice, rt.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rt.stream);
    if (e != cudaSuccess) {
        set_error("sb_columns_register: %s", cudaGetErrorString(e));
        sb_columns_release(c);
        return e == cudaErrorMemoryAllocation ? SB_ERR_OOM : SB_ERR_CUDA;
    }
    *out = c;
    return SB_OK;
}

void sb_columns_release(sb_columns_t c) {
    if (!c) return;
    if (c->d_data) cudaFree(c->d_data);
    if (c->d_ptrs) cudaFree(c->d_ptrs);
    delete c;
}

int sb_expr_eval_device(sb_prog_t prog, sb_columns_t cols, const void* const* d_adv1_cols, const void* const* d_adv2_cols,
                        size_t num_fold_vars, const uint64_t* challenges, size_t num_challenges, void* d_out, void* stream) {
    if (!prog || !cols || !d_out || (!d_adv1_cols && num_fold_vars) || (!challenges && num_challenges)) {
        set_error("sb_expr_eval_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (prog->field == FIELD_FR) return eval_enqueue<Fr>(prog, cols, d_adv1_cols, d_adv2_cols, num_fold_vars, challenges, num_challenges, d_out, st);
    return eval_enqueue<Fq>(prog, cols, d_adv1_cols, d_adv2_cols, num_fold_vars, challenges, num_challenges, d_out, st);
}

int sb_cross_terms_device(sb_prog_t prog, uint32_t degree, sb_columns_t cols, const void* const* d_adv1_cols,
                          const void* const* d_adv2_cols, size_t num_fold_vars, const uint64_t* challenges1,
                          const uint64_t* challenges2, size_t num_challenges, void* d_out, void* stream) {
    if (!prog || !cols || !d_out || ((!d_adv1_cols || !d_adv2_cols) && num_fold_vars) || ((!challenges1 || !challenges2) && num_challenges)) {
        set_error("sb_cross_terms_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (prog->field == FIELD_FR)
        return cross_terms_enqueue<Fr>(prog, degree, cols, d_adv1_cols, d_adv2_cols, num_fold_vars, challenges1, challenges2, num_challenges, d_out, st);
    return cross_terms_enqueue<Fq>(prog, degree, cols, d_adv1_cols, d_adv2_cols, num_fold_vars, challenges1, challenges2, num_challenges, d_out, st);
}

int sb_cross_terms_rows_device(sb_prog_t prog, uint32_t degree, sb_columns_t cols, const void* const* d_adv1_cols,
                               const void* const* d_adv2_cols, size_t num_fold_vars, const uint64_t* challenges1,
                               const uint64_t* challenges2, size_t num_challenges, size_t row_begin, size_t row_count, void* d_out, void* stream) {
    if (!prog || !cols || !d_out || ((!d_adv1_cols || !d_adv2_cols) && num_fold_vars) || ((!challenges1 || !challenges2) && num_challenges)) {
        set_error("sb_cross_terms_rows_device: null argument");
        return SB_ERR_ARG;
    }
    for (int32_t r : prog->h_rotations)
        if (r != 0) {   // a rotated query reads rows outside the range: the caller could not know which rows must be resident
            set_error("sb_cross_terms_rows_device: the expression queries rotation %d; row ranges need row-local expressions", (int)r);
            return SB_ERR_ARG;
        }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (prog->field == FIELD_FR)
        return cross_terms_enqueue<Fr>(prog, degree, cols, d_adv1_cols, d_adv2_cols, num_fold_vars, challenges1, challenges2, num_challenges, d_out, st, row_begin,
                                       row_count);
    return cross_terms_enqueue<Fq>(prog, degree, cols, d_adv1_cols, d_adv2_cols, num_fold_vars, challenges1, challenges2, num_challenges, d_out, st, row_begin,
                                   row_count);
}

int sb_pg_leaves_device(sb_prog_t const* gates, size_t num_gates, sb_columns_t cols, const void* const* d_cols_tables,
                        size_t num_traces, size_t num_fold_vars, const uint64_t* coef, const uint64_t* challenges,
                        size_t num_challenges, size_t num_blends, int row_mode, uint32_t log_leaves, void* d_leaves, void* stream) {
    if (!gates || !num_gates || !cols || !d_leaves || !coef || (!d_cols_tables && num_fold_vars) || !num_traces || !num_blends ||
        (!challenges && num_challenges)) {
        set_error("sb_pg_leaves_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    const int compat = row_mode == 0;
    if (gates[0]->field == FIELD_FR)
        return pg_leaves_enqueue<Fr>(gates, num_gates, cols, d_cols_tables, num_traces, num_fold_vars, coef, challenges, num_challenges, num_blends, compat, log_leaves, d_leaves, st);
    return pg_leaves_enqueue<Fq>(gates, num_gates, cols, d_cols_tables, num_traces, num_fold_vars, coef, challenges, num_challenges, num_blends, compat, log_leaves, d_leaves, st);
}

// Host-memory front end used by the Rust shim: uploads the witness rounds, maps fold variables to columns as
// PlonkEvalDomain does, runs the fused kernel, downloads T_1..T_d.
static int stage_witness(const uint64_t* const* W, const size_t* lens, size_t rounds, size_t num_advice, size_t num_lookup,
                         size_t n, char* d_base, size_t* d_off, std::vector<const void*>& col_ptrs, cudaStream_t st) {
    std::vector<char*> round_base(rounds);
    for (size_t r = 0; r < rounds; r++) {
        round_base[r] = d_base + *d_off;
        SB_CUDA_TRY(cudaMemcpyAsync(round_base[r], W[r], lens[r] * 32, cudaMemcpyHostToDevice, st));
        *d_off += align_up(lens[r] * 32, 256);
    }
    const size_t nfv = num_advice + 5 * num_lookup;
    col_ptrs.resize(nfv);
    for (size_t i = 0; i < nfv; i++) {
        size_t round = 0, col = 0;
        if (rounds == 1 && i >= num_advice) {
            set_error("Invalid witness index. num_witness: 1, num_advice: %zu, num_lookup: %zu, index: %zu", num_advice, num_lookup, i);
            return SB_ERR_ARG;
        }
        if (rounds == 1) { round = 0; col = i; }
        else SB_TRY(fold_var_location(i, num_advice, num_lookup, rounds, &round, &col));
        if (round >= rounds || (col + 1) * n > lens[round]) {
            set_error("Invalid witness index. num_witness: %zu, num_advice: %zu, num_lookup: %zu, index: %zu", rounds, num_advice, num_lookup, i);
            return SB_ERR_ARG;
        }
        col_ptrs[i] = round_base[round] + col * n * 32;
    }
    return SB_OK;
}

int sb_cross_terms(sb_prog_t prog, uint32_t degree, sb_columns_t cols, uint32_t num_advice, uint32_t num_lookup,
                   const uint64_t* const* W1, const size_t* W1_lens, size_t W1_rounds, const uint64_t* const* W2,
                   const size_t* W2_lens, size_t W2_rounds, const uint64_t* challenges1, const uint64_t* challenges2,
                   size_t num_challenges, uint64_t* const* out_T) {
    if (!prog || !cols || !W1 || !W2 || !W1_lens || !W2_lens || !out_T) {
        set_error("sb_cross_terms: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // ONE critical section: stage -> enqueue -> download -> sync (the stage buffer is shared)
    Scratch& g_expr_stage = ws_slot(rt.stream, WS_EXPR_STAGE);
    const size_t n = (size_t)1 << cols->log_rows;
    size_t bytes = 0;
    for (size_t r = 0; r < W1_rounds; r++) bytes += align_up(W1_lens[r] * 32, 256);
    for (size_t r = 0; r < W2_rounds; r++) bytes += align_up(W2_lens[r] * 32, 256);
    const size_t out_off = bytes;
    bytes += (size_t)degree * n * 32;
    std::vector<const void*> p1, p2;
    {
        RtLock lk(rt.mu);
        SB_TRY(g_expr_stage.reserve(bytes));
        size_t off = 0;
        SB_TRY(stage_witness(W1, W1_lens, W1_rounds, num_advice, num_lookup, n, (char*)g_expr_stage.ptr, &off, p1, rt.stream));
        SB_TRY(stage_witness(W2, W2_lens, W2_rounds, num_advice, num_lookup, n, (char*)g_expr_stage.ptr, &off, p2, rt.stream));
    }
    char* d_out = (char*)g_expr_stage.ptr + out_off;
    SB_TRY(sb_cross_terms_device(prog, degree, cols, p1.data(), p2.data(), p1.size(), challenges1, challenges2, num_challenges, d_out, nullptr));
    RtLock lk(rt.mu);
    for (uint32_t j = 0; j < degree; j++)
        SB_CUDA_TRY(cudaMemcpyAsync(out_T[j], d_out + (size_t)j * n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

int sb_expr_eval(sb_prog_t prog, sb_columns_t cols, uint32_t num_advice, uint32_t num_lookup, const uint64_t* const* W1,
                 const size_t* W1_lens, size_t W1_rounds, const uint64_t* const* W2, const size_t* W2_lens, size_t W2_rounds,
                 const uint64_t* challenges, size_t num_challenges, uint64_t* out) {
    if (!prog || !cols || !W1 || !W1_lens || !out) {
        set_error("sb_expr_eval: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // ONE critical section: stage -> enqueue -> download -> sync (the stage buffer is shared)
    Scratch& g_expr_stage = ws_slot(rt.stream, WS_EXPR_STAGE);
    const size_t n = (size_t)1 << cols->log_rows;
    size_t bytes = 0;
    for (size_t r = 0; r < W1_rounds; r++) bytes += align_up(W1_lens[r] * 32, 256);
    if (W2) for (size_t r = 0; r < W2_rounds; r++) bytes += align_up(W2_lens[r] * 32, 256);
    const size_t out_off = bytes;
    bytes += n * 32;
    std::vector<const void*> p1, p2;
    {
        RtLock lk(rt.mu);
        SB_TRY(g_expr_stage.reserve(bytes));
        size_t off = 0;
        SB_TRY(stage_witness(W1, W1_lens, W1_rounds, num_advice, num_lookup, n, (char*)g_expr_stage.ptr, &off, p1, rt.stream));
        if (W2) SB_TRY(stage_witness(W2, W2_lens, W2_rounds, num_advice, num_lookup, n, (char*)g_expr_stage.ptr, &off, p2, rt.stream));
    }
    char* d_out = (char*)g_expr_stage.ptr + out_off;
    SB_TRY(sb_expr_eval_device(prog, cols, p1.data(), W2 ? p2.data() : nullptr, p1.size(), challenges, num_challenges, d_out, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, d_out, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

// ---- folds -------------------------------------------------------------------------------------------

int sb_axpy_fold_device(int field, const void* d_w1, const void* d_w2, const uint64_t r[4], void* d_out, size_t n, void* stream) {
    if ((!d_w1 || !d_w2 || !d_out) && n) {
        set_error("sb_axpy_fold_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (!n) return SB_OK;
    unsigned blocks = (unsigned)((n + 255) / 256);
    ProfScope ps(st, PROF_FOLD, n);
    if (field == FIELD_FR) {
        Fr rr;
        memcpy(rr.v, r, 32);
        k_axpy<Fr><<<blocks, 256, 0, st>>>((const Fr*)d_w1, (const Fr*)d_w2, rr, (Fr*)d_out, n);
    } else {
        Fq rr;
        memcpy(rr.v, r, 32);
        k_axpy<Fq><<<blocks, 256, 0, st>>>((const Fq*)d_w1, (const Fq*)d_w2, rr, (Fq*)d_out, n);
    }
    SB_KERNEL_CHECK();
    return SB_OK;
}


int sb_error_fold_device(int field, const void* d_e, const void* d_T, uint32_t d, const uint64_t r[4], void* d_out, size_t n, void* stream) {
    if ((!d_e || !d_T || !d_out) && n) {
        set_error("sb_error_fold_device: null argument");
        return SB_ERR_ARG;
    }
    if (d > 64) {
        set_error("sb_error_fold_device: too many cross terms (%u)", d);
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (!n) return SB_OK;
    Scratch& g_fold_consts = ws_slot(st, WS_FOLD_CONSTS);
    SB_TRY(g_fold_consts.reserve(64 * 32));
    unsigned blocks = (unsigned)((n + 255) / 256);
    ProfScope ps(st, PROF_FOLD, n);
    if (field == FIELD_FR) {
        Fr rr, pw[64];
        memcpy(rr.v, r, 32);
        Fr cur = rr;
        for (uint32_t j = 0; j < d; j++) { pw[j] = cur; cur = mul_portable(cur, rr); }  // r^1, r^2, ... (accumulator.rs:382-385)
        SB_CUDA_TRY(cudaMemcpyAsync(g_fold_consts.ptr, pw, 32 * (d ? d : 1), cudaMemcpyHostToDevice, st));
        k_error_fold<Fr><<<blocks, 256, 0, st>>>((const Fr*)d_e, (const Fr*)d_T, (const Fr*)g_fold_consts.ptr, d, (Fr*)d_out, n);
    } else {
        Fq rr, pw[64];
        memcpy(rr.v, r, 32);
        Fq cur = rr;
        for (uint32_t j = 0; j < d; j++) { pw[j] = cur; cur = mul_portable(cur, rr); }
        SB_CUDA_TRY(cudaMemcpyAsync(g_fold_consts.ptr, pw, 32 * (d ? d : 1), cudaMemcpyHostToDevice, st));
        k_error_fold<Fq><<<blocks, 256, 0, st>>>((const Fq*)d_e, (const Fq*)d_T, (const Fq*)g_fold_consts.ptr, d, (Fq*)d_out, n);
    }
    SB_KERNEL_CHECK();
    return SB_OK;
}


// RelaxedPlonkWitness::fold on host vectors: out_w = w1 + r*w2 (n_w elements); out_e = e + sum r^j T_j (n_e rows)
int sb_axpy_fold(int field, const uint64_t* w1, const uint64_t* w2, const uint64_t r[4], uint64_t* out, size_t n) {
    if ((!w1 || !w2 || !out) && n) {
        set_error("sb_axpy_fold: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);
    Scratch& g_fold_stage = ws_slot(rt.stream, WS_FOLD_STAGE);
    {
        SB_TRY(g_fold_stage.reserve(n * 96 + 96));
        SB_CUDA_TRY(cudaMemcpyAsync(g_fold_stage.ptr, w1, n * 32, cudaMemcpyHostToDevice, rt.stream));
        SB_CUDA_TRY(cudaMemcpyAsync((char*)g_fold_stage.ptr + n * 32, w2, n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    char* d = (char*)g_fold_stage.ptr;
    SB_TRY(sb_axpy_fold_device(field, d, d + n * 32, r, d + n * 64, n, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, d + n * 64, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

int sb_error_fold(int field, const uint64_t* e, const uint64_t* const* T, uint32_t d, const uint64_t r[4], uint64_t* out, size_t n) {
    if ((!e || !out || (!T && d)) && n) {
        set_error("sb_error_fold: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);
    Scratch& g_fold_stage = ws_slot(rt.stream, WS_FOLD_STAGE);
    {
        SB_TRY(g_fold_stage.reserve(n * 32 * (d + 2) + 96));
        char* dptr = (char*)g_fold_stage.ptr;
        SB_CUDA_TRY(cudaMemcpyAsync(dptr, e, n * 32, cudaMemcpyHostToDevice, rt.stream));
        for (uint32_t j = 0; j < d; j++)
            SB_CUDA_TRY(cudaMemcpyAsync(dptr + (size_t)(j + 1) * n * 32, T[j], n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    char* dptr = (char*)g_fold_stage.ptr;
    SB_TRY(sb_error_fold_device(field, dptr, dptr + n * 32, d, r, dptr + (size_t)(d + 1) * n * 32, n, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, dptr + (size_t)(d + 1) * n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

}  // extern "C"
