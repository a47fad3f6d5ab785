// lookup.cu -- the SPS lookup columns and the remaining decider/assembly pieces that sit between two commits of
// the prover step.  Replaces, on the device:
//
//   * lookup::Arguments::evaluate_m        (reference src/plonk/lookup.rs:270-298)
//         m_i = #{j : l_j = t_i} on the FIRST row holding each distinct table value, 0 on repeats.
//         The reference builds a HashMap over l and a HashSet over t; here an open-addressing table over the t rows
//         keeps, per distinct value, the smallest row index (atomicMin), the l rows count into that row, and a last
//         pass emits F::from_u128(count).  Counts are integers, so the result does not depend on atomic ordering.
//   * lookup::Arguments::evaluate_h_g      (src/plonk/lookup.rs:300-312)
//         h_i = 1/(l_i + r) (0 when l_i + r = 0),  g_i = m_i / (t_i + r): Montgomery-trick batches of 128/256 cells per
//         warp around one divstep inversion (field.cuh inv_safegcd).  The same kernel with shift 0 resolves halo2 `Assigned` fractions
//         (util::batch_invert_assigned, src/util/mod.rs:128-153: numerator * denominator^-1, zero denominator -> 0).
//   * PlonkStructure::is_sat_log_derivative (src/plonk/mod.rs:363-397): sum_i (h_i - g_i).
//   * sparse::matrix_multiply + the mismatch count of is_sat_permutation
//         (src/polynomial/sparse.rs:7-20, src/nifs/sangria/mod.rs:385-453, src/nifs/protogalaxy/mod.rs:660-689).
//   * util::concatenate_with_padding       (src/util/mod.rs:214-218): host columns -> one device round vector.
//
// Everything here is 32-byte-cell streaming work (HBM-bound) except the inversions (4.5-6 products per cell + one
// 254-bit divstep inversion per warp) and the hash probes (one random 32-byte read per probe).
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "field.cuh"

namespace sb {

namespace {

template <class T>
SB_D T ldc(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = s[0];
    d[1] = s[1];
    return r;
}
template <class T>
SB_D void stc(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    d[0] = s[0];
    d[1] = s[1];
}

constexpr uint32_t LK_EMPTY = 0xFFFFFFFFu;

template <class F>
SB_D uint32_t cell_hash(const F& a) {
    // Montgomery residues of structured values are already well spread; mix all limbs and avalanche once.
    uint32_t h = 0x9E3779B9u;
#pragma unroll
    for (int i = 0; i < 8; i++) h = (h ^ a.v[i]) * 0x85EBCA6Bu + (h >> 15);
    h ^= h >> 16;
    h *= 0xC2B2AE35u;
    h ^= h >> 13;
    return h;
}

// slot[h] <- smallest row index of t holding that value
template <class F>
__global__ void __launch_bounds__(256)
k_lookup_insert(const F* __restrict__ t, uint32_t n_t, uint32_t* __restrict__ slots, uint32_t mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_t) return;
    const F key = ldc(t + i);
    uint32_t h = cell_hash(key) & mask;
    for (;;) {
        uint32_t cur = *((volatile uint32_t*)(slots + h));
        if (cur == LK_EMPTY) {
            cur = atomicCAS(slots + h, LK_EMPTY, i);
            if (cur == LK_EMPTY) return;
        }
        if (cur == i) return;
        if (ldc(t + cur) == key) {  // the slot only ever moves between rows with this same value
            if (i < cur) atomicMin(slots + h, i);
            return;
        }
        h = (h + 1) & mask;
    }
}

// returns the first-occurrence row of `key` in t, or LK_EMPTY
template <class F>
SB_D uint32_t lookup_find(const F& key, const F* __restrict__ t, const uint32_t* __restrict__ slots, uint32_t mask) {
    uint32_t h = cell_hash(key) & mask;
    for (;;) {
        const uint32_t cur = slots[h];
        if (cur == LK_EMPTY) return LK_EMPTY;
        if (ldc(t + cur) == key) return cur;
        h = (h + 1) & mask;
    }
}

template <class F>
__global__ void __launch_bounds__(256)
k_lookup_count(const F* __restrict__ l, uint32_t n_l, const F* __restrict__ t, const uint32_t* __restrict__ slots, uint32_t mask,
               uint32_t* __restrict__ counts) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t hit = LK_EMPTY;
    if (j < n_l) hit = lookup_find(ldc(l + j), t, slots, mask);
    // unused rows usually all look up the same value: aggregate equal targets inside the warp first
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, hit);
    if (hit != LK_EMPTY && (uint32_t)(__ffs(peers) - 1) == (threadIdx.x & 31u)) atomicAdd(counts + hit, (uint32_t)__popc(peers));
}

template <class F>
__global__ void __launch_bounds__(256)
k_lookup_emit(const F* __restrict__ t, uint32_t n_t, const uint32_t* __restrict__ slots, uint32_t mask,
              const uint32_t* __restrict__ counts, F* __restrict__ m) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_t) return;
    F out = F::zero();
    const uint32_t c = counts[i];
    if (c != 0 && lookup_find(ldc(t + i), t, slots, mask) == i) {  // counts live on first occurrences only
        F x = F::zero();
        x.v[0] = c;
        out = to_mont(x);
    }
    stc(m + i, out);
}

// ---- batched inversion ---------------------------------------------------------------------------------------
// One divstep (safegcd) inversion per WARP: every lane multiplies up its CH cells (strided, so loads coalesce), the 32
// lane totals are combined by a prefix and a suffix scan through shuffles, lane 0 inverts the warp total, and each
// lane recovers the inverse of its own total as inv_total * (product of the lanes before) * (product of the lanes
// after) before unwinding its cells.  3 + 12/CH products per cell and 1/(32 CH) inversions.
struct InvJobs {
    const void* in[2];
    const void* scale[2];  // nullptr = 1
    void* out[2];
    const void* shift;     // one 32-byte cell
};

template <class F>
SB_D F shfl_up_fe(const F& v, int d) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(0xFFFFFFFFu, v.v[i], d);
    return r;
}
template <class F>
SB_D F shfl_down_fe(const F& v, int d) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xFFFFFFFFu, v.v[i], d);
    return r;
}
template <class F>
SB_D F shfl_fe(const F& v, int src) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xFFFFFFFFu, v.v[i], src);
    return r;
}

// out[i] = scale[i] * (in[i] + shift)^-1, 0 where in[i] + shift = 0;  blockIdx.y selects the job
template <class F, int CH>
__global__ void __launch_bounds__(128)
k_scaled_inverse(const InvJobs jobs, size_t n) {
    const bool second = blockIdx.y != 0;  // (selected, not indexed: keeps the parameter struct out of local memory)
    const F* __restrict__ in = (const F*)(second ? jobs.in[1] : jobs.in[0]);
    const F* __restrict__ scale = (const F*)(second ? jobs.scale[1] : jobs.scale[0]);
    F* __restrict__ out = (F*)(second ? jobs.out[1] : jobs.out[0]);
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t base = warp * (size_t)(32 * CH) + lane;
    if (warp * (size_t)(32 * CH) >= n) return;  // whole warps leave together
    const F shift = ldc((const F*)jobs.shift);
    F pre[CH];
    F acc = F::one();
#pragma unroll
    for (int i = 0; i < CH; i++) {
        const size_t idx = base + (size_t)i * 32;
        pre[i] = acc;
        if (idx < n) {
            const F v = add(ldc(in + idx), shift);
            if (!v.is_zero()) acc = mul(acc, v);
        }
    }
    F inc = acc, suf = acc;  // inclusive prefix / suffix products of the lane totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const F up = shfl_up_fe(inc, d);
        const F dn = shfl_down_fe(suf, d);
        if (lane >= d) inc = mul(up, inc);
        if (lane + d < 32) suf = mul(suf, dn);
    }
    F inv_total = shfl_fe(inc, 31);
    if (lane == 0) inv_total = inv_safegcd(inv_total);  // a product of non-zero elements (or one)
    inv_total = shfl_fe(inv_total, 0);
    F before = shfl_up_fe(inc, 1);
    F after = shfl_down_fe(suf, 1);
    if (lane == 0) before = F::one();
    if (lane == 31) after = F::one();
    F inv_all = mul(mul(inv_total, before), after);  // (this lane's total)^-1
#pragma unroll
    for (int i = CH - 1; i >= 0; i--) {
        const size_t idx = base + (size_t)i * 32;
        if (idx < n) {
            const F v = add(ldc(in + idx), shift);
            F o = F::zero();
            if (!v.is_zero()) {
                o = mul(inv_all, pre[i]);
                inv_all = mul(inv_all, v);
                if (scale) {
                    const F s = ldc(scale + idx);
                    o = s.is_zero() ? s : mul(o, s);
                }
            }
            stc(out + idx, o);
        }
    }
}

constexpr int SUM_BLOCK = 256;

// partial[b] = sum over a grid-stride slice of (a_i - b_i); b == nullptr sums a alone
template <class F>
__global__ void __launch_bounds__(SUM_BLOCK)
k_sum_diff(const F* __restrict__ a, const F* __restrict__ b, size_t n, F* __restrict__ partial) {
    __shared__ F sh[SUM_BLOCK];
    F acc = F::zero();
    for (size_t i = (size_t)blockIdx.x * SUM_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * SUM_BLOCK) {
        F v = ldc(a + i);
        if (b) v = sub(v, ldc(b + i));
        acc = add(acc, v);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SUM_BLOCK / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = add(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) stc(partial + blockIdx.x, sh[0]);
}

// count += #{i : a[i] != b[i]} (b == nullptr: a[i] != 0): the filter/count of the deciders (is_sat_accumulation compares the
// evaluated rows with E, src/nifs/sangria/mod.rs:349-370; PlonkStructure::is_sat compares them with zero, plonk/mod.rs:321-338)
template <class F>
__global__ void __launch_bounds__(256)
k_count_mismatch(const F* __restrict__ a, const F* __restrict__ b, size_t n, unsigned long long* __restrict__ count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n) {
        const F x = ldc(a + i);
        bad = b ? (x != ldc(b + i)) : !x.is_zero();
    }
    const uint32_t votes = __ballot_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31u) == 0 && votes) atomicAdd(count, (unsigned long long)__popc(votes));
}

// y = P * Z row by row (CSR), count rows with y != Z[row].  Z = head (first head_len cells) ++ tail.
template <class F>
__global__ void __launch_bounds__(256)
k_sparse_mismatch(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const F* __restrict__ val, const F* __restrict__ head,
                  size_t head_len, const F* __restrict__ tail, size_t N, unsigned long long* __restrict__ mismatches) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (row < N) {
        F y = F::zero();
        for (uint32_t e = row_ptr[row]; e < row_ptr[row + 1]; e++) {
            const size_t c = col[e];
            const F z = c < head_len ? ldc(head + c) : ldc(tail + (c - head_len));
            y = add(y, mul(ldc(val + e), z));
        }
        const F zr = row < head_len ? ldc(head + row) : ldc(tail + (row - head_len));
        bad = y != zr;
    }
    const uint32_t votes = __ballot_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31u) == 0 && votes) atomicAdd(mismatches, (unsigned long long)__popc(votes));
}

// device scratch (per stream, common.cuh): WS_LK = hash slots + counts / reduction partials / mismatch counter,
// WS_LK_STAGE = host front ends' staged inputs and outputs, WS_INV_SHIFT = the inversion kernel's shift cell

struct SparseMatrix {
    int field;
    size_t N, nnz;
    uint32_t* d_row_ptr = nullptr;
    uint32_t* d_col = nullptr;
    void* d_val = nullptr;
};

template <class F>
int multiplicity_enqueue(const void* d_l, size_t n_l, const void* d_t, size_t n_t, void* d_m, cudaStream_t st) {
    if (!n_t) return SB_OK;
    size_t slots = 1024;
    while (slots < 2 * n_t) slots <<= 1;
    const size_t slot_bytes = align_up(slots * 4, 256);
    Scratch& g_lk_ws = ws_slot(st, WS_LK);
    SB_TRY(g_lk_ws.reserve(slot_bytes + align_up(n_t * 4, 256)));
    uint32_t* d_slots = (uint32_t*)g_lk_ws.ptr;
    uint32_t* d_counts = (uint32_t*)((char*)g_lk_ws.ptr + slot_bytes);
    SB_CUDA_TRY(cudaMemsetAsync(d_slots, 0xFF, slots * 4, st));
    SB_CUDA_TRY(cudaMemsetAsync(d_counts, 0, n_t * 4, st));
    const uint32_t mask = (uint32_t)(slots - 1);
    k_lookup_insert<F><<<(unsigned)((n_t + 255) / 256), 256, 0, st>>>((const F*)d_t, (uint32_t)n_t, d_slots, mask);
    SB_KERNEL_CHECK();
    if (n_l) {
        k_lookup_count<F><<<(unsigned)((n_l + 255) / 256), 256, 0, st>>>((const F*)d_l, (uint32_t)n_l, (const F*)d_t, d_slots, mask, d_counts);
        SB_KERNEL_CHECK();
    }
    k_lookup_emit<F><<<(unsigned)((n_t + 255) / 256), 256, 0, st>>>((const F*)d_t, (uint32_t)n_t, d_slots, mask, d_counts, (F*)d_m);
    SB_KERNEL_CHECK();
    return SB_OK;
}


template <class F>
int scaled_inverse_enqueue(const void* const* d_in, const uint64_t* shift, const void* const* d_scale, void* const* d_out, int jobs, size_t n,
                           cudaStream_t st) {
    if (!n) return SB_OK;
    Scratch& g_inv_shift = ws_slot(st, WS_INV_SHIFT);
    SB_TRY(g_inv_shift.reserve(256));
    uint64_t zero[4] = {0, 0, 0, 0};
    // stream-ordered, so back-to-back calls may reuse the cell
    SB_CUDA_TRY(cudaMemcpyAsync(g_inv_shift.ptr, shift ? shift : zero, 32, cudaMemcpyHostToDevice, st));
    InvJobs j;
    for (int q = 0; q < 2; q++) {
        j.in[q] = q < jobs ? d_in[q] : nullptr;
        j.scale[q] = q < jobs ? d_scale[q] : nullptr;
        j.out[q] = q < jobs ? d_out[q] : nullptr;
    }
    j.shift = g_inv_shift.ptr;
    // few cells: short chains per lane so that every SM gets a warp; many cells: fewer inversions and scan products
    // per cell (the inversion is issue-bound at ~2 warps per scheduler, profiles/r1_microbench4.txt)
    const int ch = n <= ((size_t)1 << 18) ? 4 : (n < ((size_t)1 << 20) ? 8 : 16);
    const size_t per_warp = (size_t)32 * ch;
    const size_t warps = (n + per_warp - 1) / per_warp;
    dim3 grid((unsigned)((warps + 3) / 4), (unsigned)jobs);
    if (ch == 4) k_scaled_inverse<F, 4><<<grid, 128, 0, st>>>(j, n);
    else if (ch == 8) k_scaled_inverse<F, 8><<<grid, 128, 0, st>>>(j, n);
    else k_scaled_inverse<F, 16><<<grid, 128, 0, st>>>(j, n);
    SB_KERNEL_CHECK();
    return SB_OK;
}

template <class F>
int sum_diff_enqueue(const void* d_a, const void* d_b, size_t n, void* d_out, cudaStream_t st) {
    Runtime& rt = runtime();
    size_t blocks = (n + SUM_BLOCK - 1) / SUM_BLOCK;
    const size_t cap = (size_t)rt.sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    Scratch& g_lk_ws = ws_slot(st, WS_LK);
    SB_TRY(g_lk_ws.reserve(blocks * 32));
    F* d_part = (F*)g_lk_ws.ptr;
    k_sum_diff<F><<<(unsigned)blocks, SUM_BLOCK, 0, st>>>((const F*)d_a, (const F*)d_b, n, d_part);
    SB_KERNEL_CHECK();
    k_sum_diff<F><<<1, SUM_BLOCK, 0, st>>>(d_part, (const F*)nullptr, blocks, (F*)d_out);
    SB_KERNEL_CHECK();
    return SB_OK;
}

}  // namespace

}  // namespace sb

using namespace sb;

#define SB_FIELD_DISPATCH(field, fn, name, ...)                      \
    do {                                                             \
        if ((field) == FIELD_FR) return fn<Fr>(__VA_ARGS__);         \
        if ((field) == FIELD_FQ) return fn<Fq>(__VA_ARGS__);         \
        set_error(name ": unknown field %d", (field));               \
        return SB_ERR_ARG;                                           \
    } while (0)

extern "C" {

int sb_lookup_multiplicity_device(int field, const void* d_l, size_t n_l, const void* d_t, size_t n_t, void* d_m, void* stream) {
    if ((n_l && !d_l) || (n_t && (!d_t || !d_m)) || n_l >= 0xFFFFFFFFull || n_t > 0x40000000ull) {
        set_error("sb_lookup_multiplicity_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    SB_FIELD_DISPATCH(field, multiplicity_enqueue, "sb_lookup_multiplicity_device", d_l, n_l, d_t, n_t, d_m, st);
}

int sb_scaled_inverse_device(int field, const void* d_in, const uint64_t shift[4], const void* d_scale, void* d_out, size_t n, void* stream) {
    if (n && (!d_in || !d_out)) {
        set_error("sb_scaled_inverse_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    const void* ins[1] = {d_in};
    const void* scales[1] = {d_scale};
    void* outs[1] = {d_out};
    SB_FIELD_DISPATCH(field, scaled_inverse_enqueue, "sb_scaled_inverse_device", ins, shift, scales, outs, 1, n, st);
}

int sb_lookup_inverses_device(int field, const void* d_l, const void* d_t, const void* d_m, const uint64_t r[4], size_t n, void* d_h,
                              void* d_g, void* stream) {
    if (!r || (n && (!d_l || !d_t || !d_m || !d_h || !d_g))) {
        set_error("sb_lookup_inverses_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    const void* ins[2] = {d_l, d_t};       // h = (l + r)^-1 and g = m (t + r)^-1 in one launch
    const void* scales[2] = {nullptr, d_m};
    void* outs[2] = {d_h, d_g};
    SB_FIELD_DISPATCH(field, scaled_inverse_enqueue, "sb_lookup_inverses_device", ins, r, scales, outs, 2, n, st);
}

int sb_sum_diff_device(int field, const void* d_a, const void* d_b, size_t n, void* d_out, void* stream) {
    if (!d_out || (n && !d_a)) {
        set_error("sb_sum_diff_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    SB_FIELD_DISPATCH(field, sum_diff_enqueue, "sb_sum_diff_device", d_a, d_b, n, d_out, st);
}

/* d_count_u64 (device, 8 bytes) = #{i < n : a[i] != b[i]}; b NULL compares with zero.  Not synchronised. */
int sb_count_mismatch_device(int field, const void* d_a, const void* d_b, size_t n, void* d_count_u64, void* stream) {
    if (!d_count_u64 || (n && !d_a) || (field != FIELD_FR && field != FIELD_FQ)) {
        set_error("sb_count_mismatch_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    SB_CUDA_TRY(cudaMemsetAsync(d_count_u64, 0, 8, st));
    if (!n) return SB_OK;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (field == FIELD_FR) k_count_mismatch<Fr><<<blocks, 256, 0, st>>>((const Fr*)d_a, (const Fr*)d_b, n, (unsigned long long*)d_count_u64);
    else k_count_mismatch<Fq><<<blocks, 256, 0, st>>>((const Fq*)d_a, (const Fq*)d_b, n, (unsigned long long*)d_count_u64);
    SB_KERNEL_CHECK();
    return SB_OK;
}

int sb_sparse_register(int field, const uint64_t* rows, const uint64_t* cols, const uint64_t* values_mont, size_t nnz, size_t N,
                       sb_sparse_t* out) {
    if (!out || (nnz && (!rows || !cols || !values_mont)) || N >= 0xFFFFFFFFull || nnz >= 0xFFFFFFFFull ||
        (field != FIELD_FR && field != FIELD_FQ)) {
        set_error("sb_sparse_register: bad argument");
        return SB_ERR_ARG;
    }
    for (size_t e = 0; e < nnz; e++) {
        if (rows[e] >= N || cols[e] >= N) {  // the reference panics with "invalid matrix multiply" (sparse.rs:15-17)
            set_error("sb_sparse_register: entry %zu (%llu,%llu) outside the %zu x %zu matrix", e, (unsigned long long)rows[e],
                      (unsigned long long)cols[e], N, N);
            return SB_ERR_ARG;
        }
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    std::vector<uint32_t> row_ptr(N + 1, 0), col(nnz ? nnz : 1);
    std::vector<uint64_t> val((nnz ? nnz : 1) * 4);
    for (size_t e = 0; e < nnz; e++) row_ptr[rows[e] + 1]++;
    for (size_t r = 0; r < N; r++) row_ptr[r + 1] += row_ptr[r];
    std::vector<uint32_t> fill(row_ptr.begin(), row_ptr.end() - 1);
    for (size_t e = 0; e < nnz; e++) {  // stable: entries of one row keep the reference's accumulation order
        const uint32_t p = fill[rows[e]]++;
        col[p] = (uint32_t)cols[e];
        memcpy(&val[(size_t)p * 4], values_mont + e * 4, 32);
    }
    SparseMatrix* m = new SparseMatrix();
    m->field = field;
    m->N = N;
    m->nnz = nnz;
    RtLock lk(rt.mu);
    cudaError_t e1 = cudaMalloc(&m->d_row_ptr, (N + 1) * 4);
    cudaError_t e2 = cudaMalloc(&m->d_col, col.size() * 4);
    cudaError_t e3 = cudaMalloc(&m->d_val, val.size() * 8);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaFree(m->d_row_ptr);
        cudaFree(m->d_col);
        cudaFree(m->d_val);
        delete m;
        set_error("sb_sparse_register: cudaMalloc failed");
        return SB_ERR_OOM;
    }
    SB_CUDA_TRY(cudaMemcpy(m->d_row_ptr, row_ptr.data(), (N + 1) * 4, cudaMemcpyHostToDevice));
    SB_CUDA_TRY(cudaMemcpy(m->d_col, col.data(), col.size() * 4, cudaMemcpyHostToDevice));
    SB_CUDA_TRY(cudaMemcpy(m->d_val, val.data(), val.size() * 8, cudaMemcpyHostToDevice));
    *out = (sb_sparse_t)m;
    return SB_OK;
}

void sb_sparse_release(sb_sparse_t h) {
    SparseMatrix* m = (SparseMatrix*)h;
    if (!m) return;
    cudaDeviceSynchronize();
    cudaFree(m->d_row_ptr);
    cudaFree(m->d_col);
    cudaFree(m->d_val);
    delete m;
}

size_t sb_sparse_dim(sb_sparse_t h) { return h ? ((SparseMatrix*)h)->N : 0; }

int sb_sparse_mismatch_device(sb_sparse_t h, const uint64_t* head, size_t head_len, const void* d_tail, size_t tail_len,
                              uint64_t* mismatches, void* stream) {
    SparseMatrix* m = (SparseMatrix*)h;
    if (!m || !mismatches || (head_len && !head) || (tail_len && !d_tail)) {
        set_error("sb_sparse_mismatch_device: null argument");
        return SB_ERR_ARG;
    }
    if (head_len + tail_len != m->N) {
        set_error("sb_sparse_mismatch_device: Z has %zu cells, the matrix is %zu x %zu", head_len + tail_len, m->N, m->N);
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    *mismatches = 0;
    if (!m->N) return SB_OK;
    const size_t head_bytes = align_up(head_len * 32, 256);
    Scratch& g_lk_ws = ws_slot(st, WS_LK);
    SB_TRY(g_lk_ws.reserve(256 + head_bytes));
    unsigned long long* d_cnt = (unsigned long long*)g_lk_ws.ptr;
    char* d_head = (char*)g_lk_ws.ptr + 256;
    SB_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 8, st));
    if (head_len) SB_CUDA_TRY(cudaMemcpyAsync(d_head, head, head_len * 32, cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)((m->N + 255) / 256);
    if (m->field == FIELD_FR)
        k_sparse_mismatch<Fr><<<blocks, 256, 0, st>>>(m->d_row_ptr, m->d_col, (const Fr*)m->d_val, (const Fr*)d_head, head_len, (const Fr*)d_tail, m->N, d_cnt);
    else
        k_sparse_mismatch<Fq><<<blocks, 256, 0, st>>>(m->d_row_ptr, m->d_col, (const Fq*)m->d_val, (const Fq*)d_head, head_len, (const Fq*)d_tail, m->N, d_cnt);
    SB_KERNEL_CHECK();
    unsigned long long got = 0;
    SB_CUDA_TRY(cudaMemcpyAsync(&got, d_cnt, 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(cudaStreamSynchronize(st));
    *mismatches = got;
    return SB_OK;
}

int sb_concat_pad_device(const uint64_t* const* columns, const size_t* lens, size_t num_columns, size_t pad_size, void* d_out,
                         size_t out_capacity, size_t* out_len, void* stream) {
    if ((num_columns && (!columns || !lens)) || (!d_out && out_capacity)) {
        set_error("sb_concat_pad_device: null argument");
        return SB_ERR_ARG;
    }
    size_t total = 0;
    for (size_t c = 0; c < num_columns; c++) total += lens[c] > pad_size ? lens[c] : pad_size;  // pad_using never truncates
    if (out_len) *out_len = total;
    if (total > out_capacity) {
        set_error("sb_concat_pad_device: %zu cells do not fit the %zu-cell output", total, out_capacity);
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    char* dst = (char*)d_out;
    for (size_t c = 0; c < num_columns; c++) {
        if (lens[c]) {
            if (!columns[c]) {
                set_error("sb_concat_pad_device: column %zu is null", c);
                return SB_ERR_ARG;
            }
            SB_CUDA_TRY(cudaMemcpyAsync(dst, columns[c], lens[c] * 32, cudaMemcpyHostToDevice, st));
        }
        dst += lens[c] * 32;
        if (lens[c] < pad_size) {
            SB_CUDA_TRY(cudaMemsetAsync(dst, 0, (pad_size - lens[c]) * 32, st));
            dst += (pad_size - lens[c]) * 32;
        }
    }
    return SB_OK;
}

int sb_upload_rows_device(const uint64_t* columns, size_t num_columns, size_t column_len, size_t row_begin, size_t row_count, void* d_out,
                          void* stream) {
    if (!num_columns || !row_count) return SB_OK;
    if (!columns || !d_out || row_begin > column_len || row_count > column_len - row_begin) {
        set_error("sb_upload_rows_device: bad argument (rows [%zu, %zu) of %zu)", row_begin, row_begin + row_count, column_len);
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    // one strided copy: num_columns pieces of row_count cells, the same pitch (one column) on both sides
    SB_CUDA_TRY(cudaMemcpy2DAsync((char*)d_out + row_begin * 32, column_len * 32, (const char*)columns + row_begin * 32, column_len * 32, row_count * 32,
                                  num_columns, cudaMemcpyHostToDevice, st));
    return SB_OK;
}

/* ---- host front ends: stage through the library workspace, block until the result is back ---------------- */

int sb_lookup_multiplicity(int field, const uint64_t* l, size_t n_l, const uint64_t* t, size_t n_t, uint64_t* m) {
    if ((n_l && !l) || (n_t && (!t || !m))) {
        set_error("sb_lookup_multiplicity: null argument");
        return SB_ERR_ARG;
    }
    if (!n_t) return SB_OK;
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // one critical section from staging to the synchronised download
    Scratch& g_lk_stage = ws_slot(rt.stream, WS_LK_STAGE);
    char* d;
    {
        SB_TRY(g_lk_stage.reserve(256 + (n_l + 2 * n_t) * 32));
        d = (char*)g_lk_stage.ptr + 256;
        if (n_l) SB_CUDA_TRY(cudaMemcpyAsync(d, l, n_l * 32, cudaMemcpyHostToDevice, rt.stream));
        SB_CUDA_TRY(cudaMemcpyAsync(d + n_l * 32, t, n_t * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    SB_TRY(sb_lookup_multiplicity_device(field, d, n_l, d + n_l * 32, n_t, d + (n_l + n_t) * 32, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(m, d + (n_l + n_t) * 32, n_t * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

int sb_scaled_inverse(int field, const uint64_t* in, const uint64_t shift[4], const uint64_t* scale, uint64_t* out, size_t n) {
    if (n && (!in || !out)) {
        set_error("sb_scaled_inverse: null argument");
        return SB_ERR_ARG;
    }
    if (!n) return SB_OK;
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // one critical section from staging to the synchronised download
    Scratch& g_lk_stage = ws_slot(rt.stream, WS_LK_STAGE);
    char* d;
    {
        SB_TRY(g_lk_stage.reserve(256 + 3 * n * 32));
        d = (char*)g_lk_stage.ptr + 256;
        SB_CUDA_TRY(cudaMemcpyAsync(d, in, n * 32, cudaMemcpyHostToDevice, rt.stream));
        if (scale) SB_CUDA_TRY(cudaMemcpyAsync(d + n * 32, scale, n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    SB_TRY(sb_scaled_inverse_device(field, d, shift, scale ? d + n * 32 : nullptr, d + 2 * n * 32, n, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, d + 2 * n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

int sb_lookup_inverses(int field, const uint64_t* l, const uint64_t* t, const uint64_t* m, const uint64_t r[4], size_t n, uint64_t* h,
                       uint64_t* g) {
    if (!r || (n && (!l || !t || !m || !h || !g))) {
        set_error("sb_lookup_inverses: null argument");
        return SB_ERR_ARG;
    }
    if (!n) return SB_OK;
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // one critical section from staging to the synchronised download
    Scratch& g_lk_stage = ws_slot(rt.stream, WS_LK_STAGE);
    char* d;
    {
        SB_TRY(g_lk_stage.reserve(256 + 5 * n * 32));
        d = (char*)g_lk_stage.ptr + 256;
        SB_CUDA_TRY(cudaMemcpyAsync(d, l, n * 32, cudaMemcpyHostToDevice, rt.stream));
        SB_CUDA_TRY(cudaMemcpyAsync(d + n * 32, t, n * 32, cudaMemcpyHostToDevice, rt.stream));
        SB_CUDA_TRY(cudaMemcpyAsync(d + 2 * n * 32, m, n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    SB_TRY(sb_lookup_inverses_device(field, d, d + n * 32, d + 2 * n * 32, r, n, d + 3 * n * 32, d + 4 * n * 32, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(h, d + 3 * n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaMemcpyAsync(g, d + 4 * n * 32, n * 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

int sb_sum_diff(int field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t out[4]) {
    if (!out || (n && !a)) {
        set_error("sb_sum_diff: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // one critical section from staging to the synchronised download
    Scratch& g_lk_stage = ws_slot(rt.stream, WS_LK_STAGE);
    char* d;
    {
        SB_TRY(g_lk_stage.reserve(256 + 2 * n * 32));
        d = (char*)g_lk_stage.ptr;
        if (n) SB_CUDA_TRY(cudaMemcpyAsync(d + 256, a, n * 32, cudaMemcpyHostToDevice, rt.stream));
        if (n && b) SB_CUDA_TRY(cudaMemcpyAsync(d + 256 + n * 32, b, n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    SB_TRY(sb_sum_diff_device(field, d + 256, b ? d + 256 + n * 32 : nullptr, n, d + 64, nullptr));
    RtLock lk(rt.mu);
    SB_CUDA_TRY(cudaMemcpyAsync(out, d + 64, 32, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

int sb_sparse_mismatch(sb_sparse_t h, const uint64_t* Z, size_t N, uint64_t* mismatches) {
    if (!Z && N) {
        set_error("sb_sparse_mismatch: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk_call(rt.mu);   // one critical section from staging to the synchronised download
    Scratch& g_lk_stage = ws_slot(rt.stream, WS_LK_STAGE);
    char* d;
    {
        SB_TRY(g_lk_stage.reserve(256 + N * 32));
        d = (char*)g_lk_stage.ptr + 256;
        if (N) SB_CUDA_TRY(cudaMemcpyAsync(d, Z, N * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    return sb_sparse_mismatch_device(h, nullptr, 0, d, N, mismatches, nullptr);
}

}  // extern "C"
