/*
 * sirius_b200.h -- C ABI of libsirius_b200.so, the B200 (sm_100a) implementation of the Sirius folding-prover
 * hot path.  These are the entry points the reference's FFI for this path would bind: the reference has no
 * FFI of its own (it is pure Rust on CPU threads), so each entry point cites the Rust item whose
 * body it replaces; INTEGRATION.md shows the `extern "C"` block + shim a maintainer adds to the reference.
 *
 * Conventions
 *   - field element: 4 x uint64 little-endian limbs, Montgomery form (R = 2^256)   (halo2curves layout)
 *   - affine point : x then y, 8 x uint64; the identity is (0,0)                  (src/commitment.rs:43-45)
 *   - every function returns 0 on success or a negative SB_ERR_* code; sb_last_error() gives the text of the
 *     last failure on the calling thread.  Nothing unwinds across the boundary.
 *   - "host" entry points take caller-owned host memory valid for the call only and block until the result
 *     is in host memory.  "_device" entry points take device pointers on the current device and enqueue on
 *     the given CUDA stream (cudaStream_t passed as void*; NULL = the library's stream), returning without
 *     synchronising.
 *   - Threading: every entry point may be called from any host thread (CommitmentKey is Sync; cargo test is
 *     multi-threaded).  Host entry points are serialised on the library's own stream, each one a single critical
 *     section from staging to the synchronised download.  "_device" entry points keep their device scratch PER
 *     STREAM: calls on one stream are ordered by the stream, calls on different streams (from different threads)
 *     share nothing and may overlap; only the host-side enqueue is serialised.  sb_stream_release(stream) frees
 *     the scratch of a stream that is about to be destroyed.
 *   - There is no CPU fallback: without a CUDA device every call fails with SB_ERR_CUDA.
 */
#ifndef SIRIUS_B200_H
#define SIRIUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_CUDA (-1)
#define SB_ERR_ARG (-2)
#define SB_ERR_OOM (-3)
#define SB_ERR_TOO_LONG (-4) /* commitment::Error::TooLongInput, src/commitment.rs:24-27 */
#define SB_ERR_NCCL (-5)

#define SB_FIELD_FR 0 /* bn256 scalar field == grumpkin base field */
#define SB_FIELD_FQ 1 /* bn256 base field   == grumpkin scalar field */
#define SB_CURVE_BN256 0
#define SB_CURVE_GRUMPKIN 1

typedef struct sb_ck* sb_ck_t;           /* device-resident CommitmentKey (src/commitment.rs:29-32) */
typedef struct sb_prog* sb_prog_t;       /* uploaded GraphEvaluator program (src/polynomial/graph_evaluator.rs:164-180) */
typedef struct sb_sparse* sb_sparse_t;   /* device-resident SparseMatrix (src/polynomial/sparse.rs:5), rows in CSR order */
typedef struct sb_comm* sb_comm_t;       /* peer-memory communicator of the multi-GPU commitment (one process per GPU) */
typedef struct sb_columns* sb_columns_t; /* device-resident selectors + fixed columns of a PlonkStructure (src/plonk/mod.rs:132-133) */

/* ValueSource (graph_evaluator.rs:57-68) and Calculation (graph_evaluator.rs:72-89) discriminants */
#define SB_VS_CONSTANT 0
#define SB_VS_INTERMEDIATE 1
#define SB_VS_FIXED 2
#define SB_VS_POLY 3
#define SB_VS_CHALLENGE 4
#define SB_OP_ADD 0
#define SB_OP_SUB 1
#define SB_OP_MUL 2
#define SB_OP_SQUARE 3
#define SB_OP_DOUBLE 4
#define SB_OP_NEGATE 5
#define SB_OP_HORNER 6 /* defined by the reference but never emitted by add_expression; rejected */
#define SB_OP_STORE 7

/* One CalculationInfo (graph_evaluator.rs:152-156): `target = opcode(a, b)`.  For FIXED/POLY sources
 * *_index is the column index and *_rot the index into the program's rotation table. */
typedef struct {
    uint8_t opcode, a_kind, b_kind, _pad;
    uint32_t a_index, a_rot;
    uint32_t b_index, b_rot;
    uint32_t target;
} sb_calc;

const char* sb_last_error(void);
int sb_version(void);

/* Select / initialise the CUDA device this process uses (one process per GPU).  device < 0: keep current. */
int sb_init(int device);
/* One process, several GPUs: devices[0] is the primary device (all _device entry points, cross terms, folds, NTT run there);
 * CommitmentKeys registered from host memory afterwards are split block-cyclically over the n devices, and the plain
 * `sb_msm(ck, host scalars, n, out)` / `sb_msm_batch` -- i.e. the unchanged Rust `CommitmentKey::commit` body of INTEGRATION.md --
 * shard every commit over them inside the library: per-device upload of the rank's share of the scalars, the Pippenger
 * pipeline on every device, peer copies of the 128-byte partial sums, one combine kernel.  n = 1 is sb_init(devices[0]). */
int sb_init_devices(const int* devices, int n);
int sb_num_devices(void);
void sb_shutdown(void);
int sb_device_count(void);
/* Frees the device scratch kept for `stream` (synchronises it first). */
void sb_stream_release(void* stream);

/* ---- CommitmentKey (src/commitment.rs) ------------------------------------------------------------ */

/* Upload `n` affine generators (the bytes of `CommitmentKey::ck`, also the on-disk format of
 * save_to_file, src/commitment.rs:99-104) and build the per-window multiples table used by sb_msm.
 * window_bits = 0 picks the window from n. */
int sb_ck_register(int curve, const uint64_t* bases_xy, size_t n, int window_bits, sb_ck_t* out);
/* Same, the generators already being in device memory. */
int sb_ck_register_device(int curve, const void* d_bases_xy, size_t n, int window_bits, void* stream, sb_ck_t* out);
void sb_ck_release(sb_ck_t ck);
/* Register one more window width for this key; every commit picks the cheapest registered width for its size
 * (large commits want wide windows, the small batched cross-term commits narrow ones). */
int sb_ck_add_window(sb_ck_t ck, int window_bits, void* stream);
size_t sb_ck_len(sb_ck_t ck);
int sb_ck_window_bits(sb_ck_t ck);
/* Performance knobs of the commitment pipeline; results are bit-identical for every setting.
 * key 0: batched-affine reduction rounds before the XYZZ bucket kernel (-1 = automatic, 0 = off, <= 8);
 * key 1: outputs per thread of one round (8 or 16);
 * key 2: sort of the digit entries (0 = automatic, 1 = counting sort with one atomic per entry, 2 = two-level
 *        partition sort through shared-memory histograms whenever the bucket count allows);
 * key 3: commitment tail (1 = 4-warp cooperative group law, the default; 0 = the single-lane / quad-lane kernels;
 *        2 = cooperative fix-up as well);
 * key 4: CUDA-graph cache of the _device commitment pipelines (1 = on, the default: a commitment called again with the same
 *        key, device buffers and stream is captured on its second call and replayed as one graph launch afterwards;
 *        0 = every call launches its kernels one by one). */
int sb_msm_tune(int key, int value);

/* CommitmentKey::commit (src/commitment.rs:81-90): out = sum_{i<n} scalars[i] * ck[i], affine.
 * n > len(ck) -> SB_ERR_TOO_LONG (the Rust shim maps it to Error::TooLongInput). */
int sb_msm(sb_ck_t ck, const uint64_t* scalars_mont, size_t n, uint64_t out_xy[8]);
/* Device-resident scalars; result written to d_out (64 B device affine, and 128 B XYZZ to d_out_xyzz if
 * non-NULL -- the un-normalised partial sum used by the multi-GPU gather). */
int sb_msm_device(sb_ck_t ck, const void* d_scalars_mont, size_t n, void* d_out_xy, void* d_out_xyzz, void* stream);
/* `batch` commitments against the same key in one pipeline (the d cross-term commits of one
 * commit_cross_terms call, src/nifs/sangria/mod.rs:151-154, are independent of each other): vector b is
 * scalars[b][0..n) on the host, or d_scalars + b*stride elements on the device; out gets batch points. */
int sb_msm_batch(sb_ck_t ck, const uint64_t* const* scalars_mont, size_t n, size_t batch, uint64_t* out_xy);
int sb_msm_batch_device(sb_ck_t ck, const void* d_scalars_mont, size_t n, size_t stride, size_t batch, void* d_out_xy,
                        void* d_out_xyzz, void* stream);
/* Multi-GPU combine: sum `count` XYZZ partials (device, 128 B each) and normalise to affine. */
int sb_msm_combine_device(int curve, const void* d_partials_xyzz, int count, void* d_out_xy, void* stream);
/* Batched form: out[b] = affine(sum_g partials[g * stride + b]), g < count ranks, b < batch commitments -- reads
 * the all-gather output [rank][batch] in place. */
int sb_msm_combine_batch_device(int curve, const void* d_partials_xyzz, int count, size_t batch, size_t stride, void* d_out_xy, void* stream);

/* ---- multi-GPU, one process per GPU: exchange over NVLink / NVSwitch peer memory (csrc/comm.cu) -----------------------
 * A communicator serves ONE stream at a time (its call sequence counter lives in device memory and advances in stream order);
 * concurrent commitment groups on two streams use two communicators.
 * sb_comm_create allocates this rank's mailbox and returns its 64-byte CUDA IPC handle; the caller all-gathers the handles
 * (any transport: torch.distributed, MPI, a file) and passes the world * 64 bytes, rank-major, to sb_comm_connect on every
 * rank, followed by a barrier of its own.  max_batch = the largest commitment group (<= 4096). */
int sb_comm_create(int rank, int world, size_t max_batch, sb_comm_t* out, unsigned char ipc_handle_out[64]);
int sb_comm_connect(sb_comm_t comm, const unsigned char* all_handles);
void sb_comm_destroy(sb_comm_t comm);
int sb_comm_status(sb_comm_t comm, void* stream); /* 0 = every exchange completed, 1 = a rank timed out (4 s), < 0 = error */
/* out[b] = affine(sum over the ranks of partial[b]), the same on every rank: ONE kernel stores this rank's partials into the
 * peers' mailboxes, waits for theirs and adds them (no host round trip, no NCCL).  Every rank must make the same calls. */
int sb_comm_allsum_points_device(sb_comm_t comm, int curve, const void* d_partials_xyzz, size_t batch, void* d_out_xy, void* stream);
/* CommitmentKey::commit of `batch` row-sharded vectors: this rank's rows of the scalars against this rank's slice of the key
 * (`ck` holds ck[col * n + row] for the rank's rows); the exchange is fused into the last kernel of the pipeline. */
int sb_msm_batch_sharded_device(sb_ck_t ck, sb_comm_t comm, const void* d_scalars_mont, size_t n, size_t stride, size_t batch, void* d_out_xy,
                                void* stream);

/* Validation of a key read from the reference's cache file (raw memory dump of [C], src/commitment.rs:99-128): counts
 * the points that are neither the identity (0,0) nor on y^2 = x^3 + b, as load_or_setup_cache does (:148-157). */
int sb_points_on_curve(int curve, const uint64_t* points_xy, size_t n, uint64_t* bad_count);
int sb_points_on_curve_device(int curve, const void* d_points_xy, size_t n, void* d_bad_count_u64, void* stream);

/* Synthetic commitment key used by the benches: d_out[i] = [first + i + 1] * G, affine. */
int sb_index_multiples_device(int curve, const uint64_t gen_xy[8], uint64_t first, size_t n, void* d_out_xy, void* stream);

/* ---- gate evaluation, Sangria cross terms, folds ------------------------------------------------------ */

/* Upload a compiled GraphEvaluator (calculations, constants incl. the leading [0,1,2], rotations). */
int sb_expr_compile(int field, const sb_calc* calcs, size_t n_calcs, const uint64_t* constants_mont, size_t n_constants,
                    const int32_t* rotations, size_t n_rotations, sb_prog_t* out);
void sb_expr_free(sb_prog_t prog);
uint32_t sb_expr_num_slots(sb_prog_t prog);
int sb_expr_field(sb_prog_t prog);

/* Selector (1 byte per row, Vec<Vec<bool>>) and fixed (Vec<Vec<F>>) columns of a PlonkStructure, 2^log_rows rows. */
int sb_columns_register(int field, uint32_t log_rows, const uint8_t* const* selectors, size_t num_selectors,
                        const uint64_t* const* fixed, size_t num_fixed, sb_columns_t* out);
void sb_columns_release(sb_columns_t cols);
uint32_t sb_columns_log_rows(sb_columns_t cols);

/* GraphEvaluator::evaluate for every row (graph_evaluator.rs:361-388) with GetDataForEval::eval_column_var
 * column addressing (src/plonk/eval.rs:57-69): out[row] = expr(row).  W1 (and W2 for two-instance expressions,
 * PlonkEvalDomain, eval.rs:153-228) are the witness round vectors, column-major.  W2 may be NULL. */
int sb_expr_eval(sb_prog_t prog, sb_columns_t cols, uint32_t num_advice, uint32_t num_lookup, const uint64_t* const* W1,
                 const size_t* W1_lens, size_t W1_rounds, const uint64_t* const* W2, const size_t* W2_lens, size_t W2_rounds,
                 const uint64_t* challenges, size_t num_challenges, uint64_t* out);
int sb_expr_eval_device(sb_prog_t prog, sb_columns_t cols, const void* const* d_adv1_cols, const void* const* d_adv2_cols,
                        size_t num_fold_vars, const uint64_t* challenges, size_t num_challenges, void* d_out, void* stream);

/* Evaluation half of VanillaFS::commit_cross_terms (src/nifs/sangria/mod.rs:110-147): `prog` is the compiled
 * HOMOGENEOUS compressed gate expression (CompressedGates::homogeneous, src/plonk/mod.rs:113-115) of folding
 * degree `degree`; challenges1 = [U1.challenges.., U1.u], challenges2 = [U2.challenges.., 1] (:113-118).
 * out_T[j-1][row] = T_j(row) for j = 1..degree -- the same vectors GroupedPoly::iter_from_first yields. */
int sb_cross_terms(sb_prog_t prog, uint32_t degree, sb_columns_t cols, uint32_t num_advice, uint32_t num_lookup,
                   const uint64_t* const* W1, const size_t* W1_lens, size_t W1_rounds, const uint64_t* const* W2,
                   const size_t* W2_lens, size_t W2_rounds, const uint64_t* challenges1, const uint64_t* challenges2,
                   size_t num_challenges, uint64_t* const* out_T);
/* d_adv*_cols: HOST arrays of num_fold_vars DEVICE column pointers; d_out: degree * 2^log_rows elements.
 * The evaluation runs on a straight-line kernel generated from the calculation list and compiled with NVRTC on first use
 * per (program, degree, column layout) -- a few seconds once; sb_expr_jit_enable(0) (or SB_EXPR_JIT=0) keeps the
 * interpreter kernel.  Both give the same bits. */
int sb_cross_terms_device(sb_prog_t prog, uint32_t degree, sb_columns_t cols, const void* const* d_adv1_cols,
                          const void* const* d_adv2_cols, size_t num_fold_vars, const uint64_t* challenges1,
                          const uint64_t* challenges2, size_t num_challenges, void* d_out, void* stream);
/* The same evaluation restricted to rows [row_begin, row_begin + row_count) of the table (output positions
 * d_out[(j-1) * 2^log_rows + row] as in the full call; the other rows of d_out are not touched).  For callers that
 * stream the fresh witness columns to the device in row blocks and evaluate a block while the next one is in flight
 * (the rows of a fused cross-term sweep are independent: src/nifs/sangria/mod.rs:126-147 iterates row by row).
 * Fails with SB_ERR_ARG if the expression queries a rotated row. */
int sb_cross_terms_rows_device(sb_prog_t prog, uint32_t degree, sb_columns_t cols, const void* const* d_adv1_cols,
                               const void* const* d_adv2_cols, size_t num_fold_vars, const uint64_t* challenges1,
                               const uint64_t* challenges2, size_t num_challenges, size_t row_begin, size_t row_count,
                               void* d_out, void* stream);

void sb_expr_jit_enable(int on);

/* RelaxedPlonkWitness::fold (src/nifs/sangria/accumulator.rs:363-404):
 *   axpy : out[i] = w1[i] + r * w2[i]                        (:366-378)
 *   error: out[i] = e[i] + sum_{j=1..d} r^j * T_j[i]         (:382-397)  (device form: T contiguous [d][n]) */
int sb_axpy_fold(int field, const uint64_t* w1, const uint64_t* w2, const uint64_t r[4], uint64_t* out, size_t n);
int sb_axpy_fold_device(int field, const void* d_w1, const void* d_w2, const uint64_t r[4], void* d_out, size_t n, void* stream);
int sb_error_fold(int field, const uint64_t* e, const uint64_t* const* T, uint32_t d, const uint64_t r[4], uint64_t* out, size_t n);
int sb_error_fold_device(int field, const void* d_e, const void* d_T, uint32_t d, const uint64_t r[4], void* d_out, size_t n, void* stream);

/* ---- Protogalaxy (src/nifs/protogalaxy) ------------------------------------------------------------ */

#define SB_ROW_COMPAT 0  /* leaf i of a gate block evaluates row `i & 2^k` == row 0: the reference's behaviour (src/plonk/mod.rs:714) */
#define SB_ROW_CORRECT 1 /* leaf i evaluates row `i % 2^k` */

/* Leaves of the beta tree: leaves[b][g * 2^k + row] = gate_g evaluated on blend b, b < num_blends, where blend b
 * is the Lagrange combination sum_j coef[b][j] * trace_j of `num_traces` witnesses (FoldedWitness::new,
 * poly/folded_witness.rs:20-143, computed on the fly) with challenge vector challenges[b][..]; the rest of the
 * 2^log_leaves entries is zero (src/plonk/mod.rs:709-711).  `gates` = one compiled GraphEvaluator per S.gates entry
 * (get_evaluate_witness_fn, src/plonk/mod.rs:683-718).  d_cols_tables: host array [num_traces][num_fold_vars] of
 * device column pointers.  A single trace with coef = [1] gives the leaves of compute_F / evaluate_e_from_trace. */
int sb_pg_leaves_device(sb_prog_t const* gates, size_t num_gates, sb_columns_t cols, const void* const* d_cols_tables,
                        size_t num_traces, size_t num_fold_vars, const uint64_t* coef, const uint64_t* challenges,
                        size_t num_challenges, size_t num_blends, int row_mode, uint32_t log_leaves, void* d_leaves, void* stream);
/* tree_reduce of compute_F (poly/mod.rs:100-185), compute_G (:330-413), evaluate_e_from_trace (mod.rs:586-639):
 * out[p] = root of the binary tree over 2^log_n leaves with node(h) = left + right * multipliers[p][h].
 * Point p reads leaves at d_leaves + p * leaf_stride elements (leaf_stride 0: all points share the leaves). */
int sb_beta_tree_device(int field, const void* d_leaves, uint32_t log_n, size_t num_points, size_t leaf_stride,
                        const uint64_t* multipliers, void* d_out, void* stream);
/* Host front end (single witness round per trace): leaves + tree; point p uses blend point_blend[p]. */
int sb_pg_tree(sb_prog_t const* gates, size_t num_gates, sb_columns_t cols, uint32_t num_advice, const uint64_t* const* traces_W,
               size_t num_traces, const uint64_t* coef, const uint64_t* challenges, size_t num_challenges, size_t num_blends,
               int row_mode, uint32_t log_leaves, const uint64_t* multipliers, size_t num_points, const uint32_t* point_blend,
               uint64_t* out);
/* ProtoGalaxy::fold_witness (mod.rs:176-210) / any Lagrange fold: out[i] = sum_j coef[j] * inputs[j][i]. */
int sb_lincomb(int field, const uint64_t* const* inputs, const uint64_t* coef, size_t num_inputs, size_t n, uint64_t* out);
int sb_lincomb_device(int field, const void* const* d_inputs, const uint64_t* coef, size_t num_inputs, size_t n, void* d_out, void* stream);

/* ---- batched inversion (building block of the SPS lookup columns, src/plonk/lookup.rs:213-365, and of
 * util::batch_invert_assigned, src/util/mod.rs:128-153): out[i] = in[i]^-1, zeros stay zero (ff::BatchInvert). */
int sb_batch_invert(int field, const uint64_t* in, uint64_t* out, size_t n);
int sb_batch_invert_device(int field, const void* d_in, void* d_out, size_t n, void* stream);

/* ---- SPS lookup columns (src/plonk/lookup.rs:213-365), run between two commits of run_sps_protocol_{2,3}
 * (src/plonk/mod.rs:501-660).  l_i / t_i themselves come from sb_expr_eval on the compressed lookup / table
 * expressions (LookupEvalDomain, src/plonk/eval.rs:84-131: one witness round holding the advice columns,
 * num_lookup = 0, challenges = [r]). */

/* Arguments::evaluate_m (lookup.rs:270-298): m[i] = F::from_u128(#{j : l[j] == t[i]}) on the first row of t holding
 * that value and 0 on every later repeat.  n_t <= 2^30. */
int sb_lookup_multiplicity(int field, const uint64_t* l, size_t n_l, const uint64_t* t, size_t n_t, uint64_t* m);
int sb_lookup_multiplicity_device(int field, const void* d_l, size_t n_l, const void* d_t, size_t n_t, void* d_m, void* stream);
/* Arguments::evaluate_h_g (lookup.rs:300-312): h[i] = (l[i]+r)^-1, g[i] = m[i] * (t[i]+r)^-1, with
 * Option::from(invert()).unwrap_or(ZERO) for a zero denominator. */
int sb_lookup_inverses(int field, const uint64_t* l, const uint64_t* t, const uint64_t* m, const uint64_t r[4], size_t n, uint64_t* h,
                       uint64_t* g);
int sb_lookup_inverses_device(int field, const void* d_l, const void* d_t, const void* d_m, const uint64_t r[4], size_t n, void* d_h,
                              void* d_g, void* stream);
/* The kernel under both: out[i] = scale[i] * (in[i] + shift)^-1, 0 where in[i] + shift == 0; shift NULL = 0,
 * scale NULL = 1.  With shift NULL it is util::batch_invert_assigned's numerator * denominator^-1
 * (src/util/mod.rs:120-153; pass denominator 1 for Assigned::Trivial / Zero). */
int sb_scaled_inverse(int field, const uint64_t* in, const uint64_t shift[4], const uint64_t* scale, uint64_t* out, size_t n);
int sb_scaled_inverse_device(int field, const void* d_in, const uint64_t shift[4], const void* d_scale, void* d_out, size_t n, void* stream);
/* is_sat_log_derivative's inner sum (src/plonk/mod.rs:367-376): out = sum_i (a[i] - b[i]); b NULL sums a.
 * The _device form writes the 32-byte result to d_out without synchronising. */
int sb_sum_diff(int field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t out[4]);
int sb_sum_diff_device(int field, const void* d_a, const void* d_b, size_t n, void* d_out, void* stream);

/* Row filter of the deciders: *d_count_u64 (device, 8 bytes) = #{i < n : a[i] != b[i]}; b NULL compares with zero.
 * is_sat_accumulation compares the evaluated rows with E (src/nifs/sangria/mod.rs:349-370), PlonkStructure::is_sat with
 * zero (src/plonk/mod.rs:321-338).  Enqueues without synchronising. */
int sb_count_mismatch_device(int field, const void* d_a, const void* d_b, size_t n, void* d_count_u64, void* stream);

/* ---- permutation decider (src/nifs/sangria/mod.rs:385-453, src/nifs/protogalaxy/mod.rs:660-689) ------------ */

/* Upload a SparseMatrix = Vec<(row, col, value)> of an N x N matrix (src/polynomial/sparse.rs:5).  Entries outside
 * the matrix are rejected (the reference panics "invalid matrix multiply", sparse.rs:15-17). */
int sb_sparse_register(int field, const uint64_t* rows, const uint64_t* cols, const uint64_t* values_mont, size_t nnz, size_t N,
                       sb_sparse_t* out);
void sb_sparse_release(sb_sparse_t m);
size_t sb_sparse_dim(sb_sparse_t m);
/* mismatches = #{row : (P*Z)[row] != Z[row]} (sparse::matrix_multiply + the filter/count of is_sat_permutation).
 * Z = head (host cells: the instance part) followed by tail (device cells: W[0][.. 2^k * num_advice]).
 * Blocks until the count is on the host. */
int sb_sparse_mismatch(sb_sparse_t m, const uint64_t* Z, size_t N, uint64_t* mismatches);
int sb_sparse_mismatch_device(sb_sparse_t m, const uint64_t* head, size_t head_len, const void* d_tail, size_t tail_len,
                              uint64_t* mismatches, void* stream);

/* ---- witness assembly: util::concatenate_with_padding (src/util/mod.rs:214-218) straight into a device round
 * vector: column c occupies max(lens[c], pad_size) cells, zero-padded (pad_using never truncates).  *out_len (may
 * be NULL) receives the total; SB_ERR_ARG if it exceeds out_capacity (cells). */
int sb_concat_pad_device(const uint64_t* const* columns, const size_t* lens, size_t num_columns, size_t pad_size, void* d_out,
                         size_t out_capacity, size_t* out_len, void* stream);

/* Rows [row_begin, row_begin + row_count) of every column of a column-major witness round ([num_columns][column_len]
 * cells, the layout concatenate_with_padding produces) host -> device, as ONE strided asynchronous copy into the same
 * positions of the device round vector.  With page-locked host memory a caller streams the fresh witness in row blocks
 * and starts sb_cross_terms_rows_device on a block while the next one is in flight. */
int sb_upload_rows_device(const uint64_t* columns, size_t num_columns, size_t column_len, size_t row_begin, size_t row_count,
                          void* d_out, void* stream);

/* ---- fft (src/fft.rs) ------------------------------------------------------------------------------ */

/* best_fft (src/fft.rs:61-115): in-place radix-2 transform of n = 2^log_n elements, natural order in and
 * out, out[i] = sum_j a[j] * omega^(i*j).  `omega` (order n) comes from the Rust constants via
 * get_omega_or_inv (src/fft.rs:12-23), so no root of unity is hard-coded here.  If `scale` is non-NULL every
 * output is multiplied by it (the ifft divisor TWO_INV^k, src/fft.rs:25-27,177-181).  field must be SB_FIELD_FR
 * (Fq has 2-adicity 1). */
int sb_ntt(int field, uint64_t* a, uint32_t log_n, const uint64_t omega[4], const uint64_t* scale);
int sb_ntt_device(int field, void* d_a, uint32_t log_n, const uint64_t omega[4], const uint64_t* scale, void* stream);
/* distribute_powers_zeta (src/fft.rs:207-228): a[i] *= z if i%3==1, a[i] *= z2 if i%3==2.
 * coset_fft passes (ZETA, ZETA^2), coset_ifft passes (ZETA^2, ZETA). */
int sb_coset_scale(int field, uint64_t* a, size_t n, const uint64_t z[4], const uint64_t z2[4]);
int sb_coset_scale_device(int field, void* d_a, size_t n, const uint64_t z[4], const uint64_t z2[4], void* stream);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------------- */
uint64_t sb_launch_count(void);  /* kernels launched by this library so far */
#define SB_PROF_NUM_TAGS 10 /* decompose, sort, accumulate, fixup, reduce, finalize, cross_terms, fold, ntt, protogalaxy */
void sb_profile_enable(int on);  /* CUDA events around each kernel group, on the launching stream */
int sb_profile_collect(double* total_ms, uint64_t* total_units, uint64_t* launches); /* arrays of SB_PROF_NUM_TAGS */

/* device micro-benchmarks of the arithmetic primitives (tools/microbench.py) */
int sb_microbench(int which, int iters, int blocks, int threads, double* out_ms);

/* ---- self test hooks used by tests/ (device arithmetic vs its portable twin) ----------------------- */
int sb_selftest_field(int field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_mul_ptx,
                      uint64_t* out_mul_portable, uint64_t* out_add, uint64_t* out_sub, uint64_t* out_inv);
/* lazy-domain twins (operands anywhere in [0, 2p), raw results; out_canon = a mod p with bit 255 set where a = 0 mod p) */
int sb_selftest_lazy(int field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_mul, uint64_t* out_sub,
                     uint64_t* out_dbl, uint64_t* out_canon);

/* run-time compiled cross-term kernels (csrc/expr.cu): generate + NVRTC-compile the straight-line kernel of a calculation list for
 * (degree, column layout) without a device; log receives the compiler output (ptxas resource usage).  degree = 0 generates
 * the plain per-row evaluation kernel (sb_expr_eval), degree = num_traces << 8 the Protogalaxy leaf kernel on a blend of traces */
int sb_expr_jit_selftest(int field, const sb_calc* calcs, size_t n_calcs, size_t n_constants, const int32_t* rotations, size_t n_rotations, uint32_t degree,
                         uint32_t num_selectors, uint32_t num_fixed, uint32_t num_fold_vars, uint32_t num_challenges, char* log, size_t log_cap,
                         size_t* cubin_bytes);
/* group law: the 4-warp cooperative addition / doubling of the commitment tail (csrc/coop.cuh) against the single-lane forms,
 * on a chain that visits every exceptional case; 3 affine points per input pair in each output */
int sb_selftest_coop(int curve, const uint64_t* a_xy, const uint64_t* b_xy, size_t n, uint64_t* out_coop_xy, uint64_t* out_plain_xy);

#ifdef __cplusplus
}
#endif
#endif
