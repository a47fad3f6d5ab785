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

typedef struct sb_ck* sb_ck_t; /* device-resident CommitmentKey (src/commitment.rs:29-32) */

const char* sb_last_error(void);
int sb_version(void);

/* Select / initialise the CUDA device this process uses (one process per GPU).  device < 0: keep current. */
int sb_init(int device);
void sb_shutdown(void);
int sb_device_count(void);

/* ---- CommitmentKey (src/commitment.rs) ------------------------------------------------------------ */

/* Upload `n` affine generators (the bytes of `CommitmentKey::ck`, also the on-disk format of
 * save_to_file, src/commitment.rs:99-104) and build the per-window multiples table used by sb_msm.
 * window_bits = 0 picks the window from n. */
int sb_ck_register(int curve, const uint64_t* bases_xy, size_t n, int window_bits, sb_ck_t* out);
/* Same, the generators already being in device memory. */
int sb_ck_register_device(int curve, const void* d_bases_xy, size_t n, int window_bits, void* stream, sb_ck_t* out);
void sb_ck_release(sb_ck_t ck);
size_t sb_ck_len(sb_ck_t ck);
int sb_ck_window_bits(sb_ck_t ck);

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

/* ---- self test hooks used by tests/ (device arithmetic vs its portable twin) ----------------------- */
int sb_selftest_field(int field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_mul_ptx,
                      uint64_t* out_mul_portable, uint64_t* out_add, uint64_t* out_sub, uint64_t* out_inv);

#ifdef __cplusplus
}
#endif
#endif
