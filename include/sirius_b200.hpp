// sirius_b200.hpp -- C++17 host-side mirror, above the C ABI (sirius_b200.h), of the Rust items on the hot path whose
// bodies libsirius_b200.so replaces: same names, argument meaning and error behaviour as the reference, so a C++ caller
// (or a test) reads like the Rust call sites.  The reference is compiled code (Rust) and its toolchain is absent from this
// image, hence a compiled-language mirror beside the Python one (sirius_b200/*.py, which the parity tests drive).
//
//   sirius_b200::CommitmentKey            reference src/commitment.rs:29-170 (struct, Deref<[C]>, default_value, len,
//                                         is_empty, commit, save_to_file, load_from_file, load_or_setup_cache)
//   sirius_b200::TooLongInput             commitment::Error::TooLongInput, src/commitment.rs:24-27 (same message)
//   sirius_b200::fft::{get_omega_or_inv, get_ifft_divisor, best_fft, fft, ifft, coset_fft, coset_ifft}
//                                         src/fft.rs:12-27, 61-115, 160-198 (bn256 Fr; Fq has 2-adicity 1)
//   sirius_b200::RelaxedPlonkWitness::fold  src/nifs/sangria/accumulator.rs:363-404 (W and E folds)
//
// There is no CPU fallback here either: everything that computes goes through the C ABI and throws sirius_b200::Error
// (text of sb_last_error) when the CUDA library cannot run.  Header-only; link with libsirius_b200.so.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../sirius_b200/csrc/field.cuh"  // host-compilable portable field arithmetic (constants of the transforms)
#include "sirius_b200.h"

namespace sirius_b200 {

using Scalar = std::array<uint64_t, 4>;  // field element: 4 x u64 little-endian limbs, Montgomery form (R = 2^256)

struct Affine {  // halo2curves affine point; the identity is (0,0)  (src/commitment.rs:43-45)
    Scalar x{}, y{};
    bool is_identity() const { return (x[0] | x[1] | x[2] | x[3] | y[0] | y[1] | y[2] | y[3]) == 0; }
    bool operator==(const Affine& o) const { return x == o.x && y == o.y; }
};
static_assert(sizeof(Affine) == 64, "64-byte points: the key file format is a memory image (src/commitment.rs:99-116)");

enum class Curve : int { Bn256G1 = SB_CURVE_BN256, Grumpkin = SB_CURVE_GRUMPKIN };

struct Error : std::runtime_error {  // device / library failure: Rust has no variant for it, the shim panics
    int code;
    Error(int c, const std::string& m) : std::runtime_error("libsirius_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
struct TooLongInput : std::runtime_error {  // commitment::Error::TooLongInput { input_len, limit }
    size_t input_len, limit;
    TooLongInput(size_t n, size_t lim)
        : std::runtime_error("Can't commit too long input: input len: " + std::to_string(n) + ", but limit is " + std::to_string(lim)),
          input_len(n), limit(lim) {}
};
struct IoError : std::runtime_error {  // io::Error of the key-cache functions
    enum Kind { UnexpectedEof, InvalidData, Other } kind;
    IoError(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

inline void check(int rc) {
    if (rc != SB_OK) throw Error(rc, sb_last_error());
}

// ------------------------------------------------------------------------------------------------ CommitmentKey
class CommitmentKey {
   public:
    CommitmentKey(Curve curve, std::vector<Affine> ck) : curve_(curve), ck_(std::move(ck)) {}
    CommitmentKey(const CommitmentKey& o) : curve_(o.curve_), ck_(o.ck_) {}  // Clone: the device copy is rebuilt lazily
    CommitmentKey& operator=(const CommitmentKey& o) {
        if (this != &o) { release(); curve_ = o.curve_; ck_ = o.ck_; }
        return *this;
    }
    ~CommitmentKey() { release(); }

    static Affine default_value() { return Affine{}; }  // C::identity()
    size_t len() const { return ck_.size(); }
    bool is_empty() const { return ck_.empty(); }
    const Affine* data() const { return ck_.data(); }  // Deref<Target = [C]>
    const Affine& operator[](size_t i) const { return ck_[i]; }
    bool operator==(const CommitmentKey& o) const { return curve_ == o.curve_ && ck_ == o.ck_; }
    Curve curve() const { return curve_; }

    // `commit(&self, v: &[C::Scalar]) -> Result<C, Error>`: sum_i v[i] * ck[i], affine.  The length check comes first,
    // exactly as in the reference (src/commitment.rs:82-89): it throws before anything touches the device.
    Affine commit(const Scalar* v, size_t n) const {
        if (n > ck_.size()) throw TooLongInput(n, ck_.size());
        ensure_registered();
        Affine out;
        check(sb_msm(handle_, reinterpret_cast<const uint64_t*>(v), n, reinterpret_cast<uint64_t*>(&out)));
        return out;
    }
    Affine commit(const std::vector<Scalar>& v) const { return commit(v.data(), v.size()); }
    // [commit(v) for v in vs] as ONE device pipeline (all vectors of the same length): the d cross-term commits of
    // VanillaFS::commit_cross_terms (src/nifs/sangria/mod.rs:151-154)
    std::vector<Affine> commit_batch(const std::vector<std::vector<Scalar>>& vs) const {
        std::vector<Affine> out(vs.size());
        if (vs.empty()) return out;
        const size_t n = vs[0].size();
        std::vector<const uint64_t*> ptrs;
        for (const auto& v : vs) {
            if (v.size() != n) throw std::invalid_argument("commit_batch: vectors must have equal length");
            ptrs.push_back(reinterpret_cast<const uint64_t*>(v.data()));
        }
        if (n > ck_.size()) throw TooLongInput(n, ck_.size());
        ensure_registered();
        check(sb_msm_batch(handle_, ptrs.data(), n, vs.size(), reinterpret_cast<uint64_t*>(out.data())));
        return out;
    }

    // `save_to_file`: the key as a memory image, 64 bytes per point (src/commitment.rs:99-116)
    void save_to_file(const std::string& file_path) const {
        std::ofstream f(file_path, std::ios::binary | std::ios::trunc);
        if (!f) throw IoError(IoError::Other, "cannot create " + file_path);
        f.write(reinterpret_cast<const char*>(ck_.data()), (std::streamsize)(ck_.size() * sizeof(Affine)));
        if (!f) throw IoError(IoError::Other, "short write to " + file_path);
    }
    // `load_from_file(file_path, k)`: exactly 2^k points (`read_exact`, src/commitment.rs:118-137)
    static CommitmentKey load_from_file(Curve curve, const std::string& file_path, size_t k) {
        std::ifstream f(file_path, std::ios::binary);
        if (!f) throw IoError(IoError::Other, "cannot open " + file_path);
        std::vector<Affine> ck((size_t)1 << k);
        f.read(reinterpret_cast<char*>(ck.data()), (std::streamsize)(ck.size() * sizeof(Affine)));
        if ((size_t)f.gcount() != ck.size() * sizeof(Affine)) throw IoError(IoError::UnexpectedEof, "failed to fill whole buffer");
        return CommitmentKey(curve, std::move(ck));
    }
    // `load_or_setup_cache(cache_folder, label, k)` (src/commitment.rs:139-170): {folder}/{label}/{k}.bin, every point
    // checked on the curve (on the device); a missing file is created from `setup` (the reference's own `setup` hashes
    // to the curve with the un-vendored halo2curves: SURVEY 8f-2, so the generator is supplied by the caller).
    static CommitmentKey load_or_setup_cache(Curve curve, const std::string& cache_folder, const std::string& label, size_t k,
                                             const std::function<std::vector<Affine>(size_t, const std::string&)>& setup = nullptr) {
        const std::string dir = cache_folder + "/" + label, path = dir + "/" + std::to_string(k) + ".bin";
        if (std::ifstream(path, std::ios::binary).good()) {
            CommitmentKey key = load_from_file(curve, path, k);
            uint64_t bad = 0;
            check(sb_points_on_curve((int)curve, reinterpret_cast<const uint64_t*>(key.ck_.data()), key.ck_.size(), &bad));
            if (bad) throw IoError(IoError::InvalidData, "Wrong file in cache, some ptr out of curve");
            return key;
        }
        if (!setup) throw std::logic_error("CommitmentKey::setup needs halo2curves' hash_to_curve (un-vendored); supply `setup`");
        CommitmentKey key(curve, setup(k, label));
        std::error_code ec;
        std::filesystem::create_directories(dir, ec);   // no shell: the folder / label may contain any character
        if (ec) throw IoError(IoError::Other, "cannot create " + dir + ": " + ec.message());
        key.save_to_file(path);
        return key;
    }

   private:
    void ensure_registered() const {
        if (!handle_) check(sb_ck_register((int)curve_, reinterpret_cast<const uint64_t*>(ck_.data()), ck_.size(), 0, &handle_));
    }
    void release() {
        if (handle_) sb_ck_release(handle_);
        handle_ = nullptr;
    }
    Curve curve_;
    std::vector<Affine> ck_;          // Box<[C]>
    mutable sb_ck_t handle_ = nullptr;  // device-resident window tables, built at the first commit
};

// `setup_smallest_key(k_table_size, cs, tag)` (src/commitment.rs:172-186): the key must cover one witness round
// (advice + 5 columns per lookup) and the selector + fixed columns.  smallest_power mirrors the reference's
// `((n * 2^K) as f64).log2().ceil() as usize` (n = 0: -inf casts to 0).
inline size_t smallest_power(size_t n, uint32_t K) {
    const double v = (double)n * (double)((uint64_t)1 << K);
    if (v == 0.0) return 0;
    const double w = std::ceil(std::log2(v));
    return w < 0 ? 0 : (size_t)w;
}
inline size_t smallest_key_log2(uint32_t k_table_size, size_t num_advice_columns, size_t num_lookups, size_t num_selectors,
                                size_t num_fixed_columns) {
    const size_t p1 = smallest_power(num_advice_columns + 5 * num_lookups, k_table_size);
    const size_t p2 = smallest_power(num_selectors + num_fixed_columns, k_table_size);
    return p1 > p2 ? p1 : p2;
}
inline CommitmentKey setup_smallest_key(Curve curve, uint32_t k_table_size, size_t num_advice_columns, size_t num_lookups,
                                        size_t num_selectors, size_t num_fixed_columns, const std::string& tag,
                                        const std::function<std::vector<Affine>(size_t, const std::string&)>& setup) {
    // `setup` stands in for CommitmentKey::setup (hash_to_curve of the un-vendored halo2curves)
    return CommitmentKey(curve, setup(smallest_key_log2(k_table_size, num_advice_columns, num_lookups, num_selectors, num_fixed_columns), tag));
}

// ------------------------------------------------------------------------------------------------ fft (bn256 Fr)
namespace fft {
// halo2curves bn256::Fr associated constants (canonical integers, little-endian u64 limbs).  ROOT_OF_UNITY is pinned by
// the reference's fft known-answer test (src/fft.rs:242-251); ZETA by none: pass your own if it differs.
constexpr uint32_t S = 28;
constexpr Scalar ROOT_OF_UNITY = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull, 0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull};
constexpr Scalar TWO_INV = {0xa1f0fac9f8000001ull, 0x9419f4243cdcb848ull, 0xdc2822db40c0ac2eull, 0x183227397098d014ull};
constexpr Scalar ZETA = {0xb8ca0b2d36636f23ull, 0xcc37a73fec2bc5e9ull, 0x048b6e193fd84104ull, 0x30644e72e131a029ull};

inline sb::Fr to_fe(const Scalar& s) {
    sb::Fr r;
    for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)s[i]; r.v[2 * i + 1] = (uint32_t)(s[i] >> 32); }
    return r;
}
inline Scalar from_fe(const sb::Fr& f) {
    Scalar s;
    for (int i = 0; i < 4; i++) s[i] = (uint64_t)f.v[2 * i] | ((uint64_t)f.v[2 * i + 1] << 32);
    return s;
}
inline sb::Fr mont(const Scalar& canonical) { return sb::to_mont(to_fe(canonical)); }

// `get_omega_or_inv(k, is_inverse)` (src/fft.rs:12-23), Montgomery form: ROOT_OF_UNITY(_INV) squared S - k times
inline Scalar get_omega_or_inv(uint32_t k, bool is_inverse) {
    if (k > S) throw std::invalid_argument("k=" + std::to_string(k) + " should no larger than F::S=" + std::to_string(S));
    sb::Fr w = mont(ROOT_OF_UNITY);
    if (is_inverse) w = sb::inv_safegcd(w);  // F::ROOT_OF_UNITY_INV
    for (uint32_t i = k; i < S; i++) w = sb::sqr(w);
    return from_fe(w);
}
// `get_ifft_divisor(k)` = TWO_INV^k (src/fft.rs:25-27)
inline Scalar get_ifft_divisor(uint32_t k) {
    sb::Fr r = sb::Fr::one();
    const sb::Fr h = mont(TWO_INV);
    for (uint32_t i = 0; i < k; i++) r = sb::mul(r, h);
    return from_fe(r);
}
inline uint32_t log2_len(size_t n) {
    if (n == 0 || (n & (n - 1))) throw std::invalid_argument("a.len().is_power_of_two()");
    uint32_t k = 0;
    while (((size_t)1 << k) < n) k++;
    return k;
}
// `best_fft(a, omega, log_n)` (src/fft.rs:61-115): in place, natural order in and out
inline void best_fft(Scalar* a, const Scalar& omega, uint32_t log_n, const Scalar* scale = nullptr) {
    check(sb_ntt(SB_FIELD_FR, reinterpret_cast<uint64_t*>(a), log_n, omega.data(), scale ? scale->data() : nullptr));
}
inline void fft(std::vector<Scalar>& a) {  // src/fft.rs:160-166
    const uint32_t k = log2_len(a.size());
    best_fft(a.data(), get_omega_or_inv(k, false), k);
}
inline void ifft(std::vector<Scalar>& a) {  // src/fft.rs:168-182
    const uint32_t k = log2_len(a.size());
    const Scalar div = get_ifft_divisor(k);
    best_fft(a.data(), get_omega_or_inv(k, true), k, &div);
}
inline void distribute_powers_zeta(std::vector<Scalar>& a, const Scalar& z, const Scalar& z2) {  // src/fft.rs:207-228
    check(sb_coset_scale(SB_FIELD_FR, reinterpret_cast<uint64_t*>(a.data()), a.size(), z.data(), z2.data()));
}
inline void coset_fft(std::vector<Scalar>& a, const Scalar& zeta_canonical = ZETA) {  // src/fft.rs:184-189
    const sb::Fr z = mont(zeta_canonical);
    distribute_powers_zeta(a, from_fe(z), from_fe(sb::sqr(z)));
    fft(a);
}
inline void coset_ifft(std::vector<Scalar>& a, const Scalar& zeta_canonical = ZETA) {  // src/fft.rs:191-198 (the Rust wraps the result in UnivariatePoly)
    const sb::Fr z = mont(zeta_canonical);
    ifft(a);
    distribute_powers_zeta(a, from_fe(sb::sqr(z)), from_fe(z));
}
}  // namespace fft

// ------------------------------------------------------------------------------------------------ Sangria folds
struct RelaxedPlonkWitness {  // src/nifs/sangria/accumulator.rs:273-276, single witness round
    int field;                // SB_FIELD_FR (primary) / SB_FIELD_FQ (secondary)
    std::vector<Scalar> W, E;

    // `fold(&self, W2, cross_terms, r) -> Self` (:363-404): W + r*W2 and E + sum_{j>=1} r^j T_j
    RelaxedPlonkWitness fold(const std::vector<Scalar>& W2, const std::vector<std::vector<Scalar>>& cross_terms, const Scalar& r) const {
        if (W2.size() != W.size()) throw std::invalid_argument("fold: witness lengths differ");
        RelaxedPlonkWitness out{field, std::vector<Scalar>(W.size()), std::vector<Scalar>(E.size())};
        check(sb_axpy_fold(field, reinterpret_cast<const uint64_t*>(W.data()), reinterpret_cast<const uint64_t*>(W2.data()), r.data(),
                           reinterpret_cast<uint64_t*>(out.W.data()), W.size()));
        std::vector<const uint64_t*> T;
        for (const auto& t : cross_terms) {
            if (t.size() != E.size()) throw std::invalid_argument("fold: cross term length differs from E");
            T.push_back(reinterpret_cast<const uint64_t*>(t.data()));
        }
        check(sb_error_fold(field, reinterpret_cast<const uint64_t*>(E.data()), T.data(), (uint32_t)T.size(), r.data(),
                            reinterpret_cast<uint64_t*>(out.E.data()), E.size()));
        return out;
    }
};

}  // namespace sirius_b200
