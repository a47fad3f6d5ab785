/* TEST INFRASTRUCTURE: a CPU stand-in for the handful of libsirius_b200.so entry points that include/sirius_b200.hpp
 * calls, backed by the parity oracle (oracle/sirius_oracle.c).  It exists so that the C++ mirror's full flow can be
 * executed on a machine without a GPU (tests/test_zz_cpp_mirror.py swaps it in through LD_LIBRARY_PATH); it is never
 * built into, linked with or loaded by the product. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef uint64_t u64;
int so_msm(int curve, const u64* s, const u64* b, size_t n, int threads, u64 out[8]);
int so_is_on_curve(int curve, const u64 xy[8]);
int so_best_fft(int field, u64* a, uint32_t log_n, const u64 omega[4], int threads);
int so_scale(int field, u64* a, size_t n, const u64 s[4]);
int so_coset_scale(int field, u64* a, size_t n, const u64 z[4], const u64 z2[4]);
int so_axpy(int field, const u64* w1, const u64* w2, const u64 r[4], u64* out, size_t n);
int so_error_fold(int field, const u64* e, const u64* const* T, size_t d, const u64 r[4], u64* out, size_t n);

struct sb_ck { int curve; size_t n; u64* bases; };
const char* sb_last_error(void) { return "fake libsirius_b200 (oracle-backed test stub)"; }
int sb_ck_register(int curve, const u64* bases, size_t n, int wb, struct sb_ck** out) {
    (void)wb;
    struct sb_ck* k = malloc(sizeof *k);
    k->curve = curve; k->n = n; k->bases = malloc(n * 64 + 8);
    memcpy(k->bases, bases, n * 64);
    *out = k;
    return 0;
}
void sb_ck_release(struct sb_ck* k) { if (k) { free(k->bases); free(k); } }
int sb_msm(struct sb_ck* k, const u64* s, size_t n, u64 out[8]) { return n > k->n ? -4 : so_msm(k->curve, s, k->bases, n, 0, out); }
int sb_points_on_curve(int curve, const u64* pts, size_t n, u64* bad) {
    *bad = 0;
    for (size_t i = 0; i < n; i++) {
        int z = 1;
        for (int j = 0; j < 8; j++) z &= pts[8 * i + j] == 0;
        if (!z && !so_is_on_curve(curve, pts + 8 * i)) (*bad)++;
    }
    return 0;
}
int sb_ntt(int field, u64* a, uint32_t log_n, const u64 omega[4], const u64* scale) {
    int rc = so_best_fft(field, a, log_n, omega, 0);
    if (rc == 0 && scale) rc = so_scale(field, a, (size_t)1 << log_n, scale);
    return rc;
}
int sb_coset_scale(int field, u64* a, size_t n, const u64 z[4], const u64 z2[4]) { return so_coset_scale(field, a, n, z, z2); }
int sb_axpy_fold(int field, const u64* w1, const u64* w2, const u64 r[4], u64* out, size_t n) { return so_axpy(field, w1, w2, r, out, n); }
int sb_error_fold(int field, const u64* e, const u64* const* T, uint32_t d, const u64 r[4], u64* out, size_t n) { return so_error_fold(field, e, T, d, r, out, n); }
