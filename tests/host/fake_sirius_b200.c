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

/* ---- columns, programs and the Sangria cross terms (single witness round, no lookups), for the C++ mirror's
 * VanillaFS::commit_cross_terms.  The cross terms are obtained with the ORACLE's interpreter: the homogeneous program
 * is evaluated on W1 + t*W2 (challenges likewise) for t = 0..d and the values are interpolated to the coefficients of
 * t^1..t^d -- mathematically the same vectors as the reference's degree-grouped evaluation, which the Python test computes
 * literally (oracle/expr_ref.py commit_cross_terms_eval) and compares with. */
int so_graph_evaluate(int field, const int32_t* calcs, size_t n_calcs, const u64* constants, size_t n_constants, const int32_t* rotations,
                      size_t n_rotations, const uint8_t* const* selectors, size_t n_sel, const u64* const* fixed, size_t n_fixed,
                      const u64* const* advice, size_t n_advice, const u64* challenges, size_t n_challenges, uint32_t log_rows, int threads, u64* out);
int so_field_mul(int field, const u64* a, const u64* b, u64* o, size_t n);
int so_field_add(int field, const u64* a, const u64* b, u64* o, size_t n);
int so_field_sub(int field, const u64* a, const u64* b, u64* o, size_t n);
int so_field_inv(int field, const u64* a, u64* o, size_t n);
int so_to_mont(int field, const u64* a, u64* o, size_t n);

typedef struct { uint8_t opcode, a_kind, b_kind, _pad; uint32_t a_index, a_rot, b_index, b_rot, target; } sb_calc;
struct sb_prog { int field; size_t n_calcs, n_consts, n_rots; int32_t* calcs; u64* consts; int32_t* rots; };
struct sb_columns { int field; uint32_t log_rows; size_t n_sel, n_fix; uint8_t** sel; u64** fix; };

int sb_expr_compile(int field, const sb_calc* calcs, size_t n_calcs, const u64* consts, size_t n_consts, const int32_t* rots, size_t n_rots,
                    struct sb_prog** out) {
    struct sb_prog* p = calloc(1, sizeof *p);
    p->field = field; p->n_calcs = n_calcs; p->n_consts = n_consts; p->n_rots = n_rots;
    p->calcs = malloc((n_calcs + 1) * 8 * sizeof(int32_t));
    for (size_t i = 0; i < n_calcs; i++) {
        int32_t* c = p->calcs + 8 * i;
        c[0] = calcs[i].opcode; c[1] = calcs[i].a_kind; c[2] = (int32_t)calcs[i].a_index; c[3] = (int32_t)calcs[i].a_rot;
        c[4] = calcs[i].b_kind; c[5] = (int32_t)calcs[i].b_index; c[6] = (int32_t)calcs[i].b_rot; c[7] = (int32_t)calcs[i].target;
    }
    p->consts = malloc((n_consts + 1) * 32); memcpy(p->consts, consts, n_consts * 32);
    p->rots = malloc((n_rots + 1) * sizeof(int32_t)); memcpy(p->rots, rots, n_rots * sizeof(int32_t));
    *out = p;
    return 0;
}
void sb_expr_free(struct sb_prog* p) { if (p) { free(p->calcs); free(p->consts); free(p->rots); free(p); } }
int sb_columns_register(int field, uint32_t log_rows, const uint8_t* const* sel, size_t n_sel, const u64* const* fix, size_t n_fix,
                        struct sb_columns** out) {
    const size_t n = (size_t)1 << log_rows;
    struct sb_columns* c = calloc(1, sizeof *c);
    c->field = field; c->log_rows = log_rows; c->n_sel = n_sel; c->n_fix = n_fix;
    c->sel = malloc((n_sel + 1) * sizeof(void*)); c->fix = malloc((n_fix + 1) * sizeof(void*));
    for (size_t i = 0; i < n_sel; i++) { c->sel[i] = malloc(n); memcpy(c->sel[i], sel[i], n); }
    for (size_t i = 0; i < n_fix; i++) { c->fix[i] = malloc(n * 32); memcpy(c->fix[i], fix[i], n * 32); }
    *out = c;
    return 0;
}
void sb_columns_release(struct sb_columns* c) {
    if (!c) return;
    for (size_t i = 0; i < c->n_sel; i++) free(c->sel[i]);
    for (size_t i = 0; i < c->n_fix; i++) free(c->fix[i]);
    free(c->sel); free(c->fix); free(c);
}
int sb_msm_batch(struct sb_ck* k, const u64* const* s, size_t n, size_t batch, u64* out) {
    for (size_t b = 0; b < batch; b++) {
        int rc = sb_msm(k, s[b], n, out + 8 * b);
        if (rc) return rc;
    }
    return 0;
}

static void small_mont(int field, u64 v, u64 out[4]) { u64 c[4] = {v, 0, 0, 0}; so_to_mont(field, c, out, 1); }

int sb_cross_terms(struct sb_prog* prog, uint32_t degree, struct sb_columns* cols, uint32_t num_advice, uint32_t num_lookup,
                   const u64* const* W1, const size_t* W1_lens, size_t W1_rounds, const u64* const* W2, const size_t* W2_lens, size_t W2_rounds,
                   const u64* ch1, const u64* ch2, size_t num_ch, u64* const* out_T) {
    if (num_lookup != 0 || W1_rounds != 1 || W2_rounds != 1) return -2;   /* the stub covers the single-round layout only */
    const int f = prog->field;
    const size_t n = (size_t)1 << cols->log_rows, m = degree + 1;
    if (W1_lens[0] != (size_t)num_advice * n || W2_lens[0] != W1_lens[0]) return -2;
    u64* evals = malloc(m * n * 32);          /* evals[t][row] */
    u64* blend = malloc((size_t)num_advice * n * 32);
    u64* chb = malloc((num_ch + 1) * 32);
    const u64** adv = malloc((num_advice + 1) * sizeof(void*));
    for (size_t t = 0; t < m; t++) {
        u64 tm[4];
        small_mont(f, (u64)t, tm);
        so_axpy(f, W1[0], W2[0], tm, blend, (size_t)num_advice * n);
        so_axpy(f, ch1, ch2, tm, chb, num_ch);
        for (uint32_t j = 0; j < num_advice; j++) adv[j] = blend + (size_t)j * n * 4;
        int rc = so_graph_evaluate(f, prog->calcs, prog->n_calcs, prog->consts, prog->n_consts, prog->rots, prog->n_rots,
                                   (const uint8_t* const*)cols->sel, cols->n_sel, (const u64* const*)cols->fix, cols->n_fix, adv, num_advice,
                                   chb, num_ch, cols->log_rows, 0, evals + t * n * 4);
        if (rc) return -2;
    }
    /* Lagrange interpolation on the points 0..d: coefficient vector of prod_{s != t} (X - s) / (t - s), per t */
    u64* coef = calloc(m * m * 4, sizeof(u64));   /* coef[t][j] = coefficient of X^j in the t-th Lagrange basis polynomial */
    for (size_t t = 0; t < m; t++) {
        u64* poly = calloc((m + 1) * 4, sizeof(u64));
        u64 one[4], den[4];
        small_mont(f, 1, one);
        memcpy(poly, one, 32);
        memcpy(den, one, 32);
        size_t deg = 0;
        for (size_t s = 0; s < m; s++) {
            if (s == t) continue;
            u64 sm[4], tm[4], diff[4];
            small_mont(f, (u64)s, sm);
            small_mont(f, (u64)t, tm);
            for (size_t j = deg + 1;; j--) {   /* poly *= (X - s): new[j] = old[j-1] - s*old[j], highest coefficient first */
                u64 lower[4] = {0, 0, 0, 0}, cur[4] = {0, 0, 0, 0}, prod[4];
                if (j > 0) memcpy(lower, poly + 4 * (j - 1), 32);
                if (j <= deg) memcpy(cur, poly + 4 * j, 32);
                so_field_mul(f, cur, sm, prod, 1);
                so_field_sub(f, lower, prod, poly + 4 * j, 1);
                if (j == 0) break;
            }
            deg++;
            so_field_sub(f, tm, sm, diff, 1);
            so_field_mul(f, den, diff, den, 1);
        }
        u64 inv[4];
        so_field_inv(f, den, inv, 1);
        for (size_t j = 0; j < m; j++) so_field_mul(f, poly + 4 * j, inv, coef + (t * m + j) * 4, 1);
        free(poly);
    }
    for (size_t j = 1; j <= degree; j++) {
        u64* T = out_T[j - 1];
        memset(T, 0, n * 32);
        u64* tmp = malloc(n * 32);
        for (size_t t = 0; t < m; t++) {
            for (size_t row = 0; row < n; row++) so_field_mul(f, evals + (t * n + row) * 4, coef + (t * m + j) * 4, tmp + row * 4, 1);
            so_field_add(f, T, tmp, T, n);
        }
        free(tmp);
    }
    free(coef); free(adv); free(chb); free(blend); free(evals);
    return 0;
}
