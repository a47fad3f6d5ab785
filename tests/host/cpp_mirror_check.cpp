// Exercises include/sirius_b200.hpp (the C++ host-side mirror of the Rust interface) the way the reference's own
// tests use the Rust items: key file round trip (commitment::file_tests::consistency, src/commitment.rs:197-213),
// TooLongInput, fft / ifft / coset transforms, the Sangria witness fold.  Host-only checks are asserted here; the
// results of everything that computes are printed as hex for tests/test_zz_cpp_mirror.py to compare with the oracle.
//   cpp_mirror_check <workdir> <inputs.bin>
// inputs.bin: u64 n_pts, u64 log_fft | n_pts points (64 B) | n_pts scalars (32 B) | 2^log_fft scalars | 32 B challenge r
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/sirius_b200.hpp"

using namespace sirius_b200;

static void hex(const char* tag, const uint64_t* w, size_t n_words) {
    std::printf("%s", tag);
    for (size_t i = 0; i < n_words; i++) std::printf(" %016llx", (unsigned long long)w[i]);
    std::printf("\n");
}
#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string dir = argv[1];
    std::FILE* f = std::fopen(argv[2], "rb");
    REQUIRE(f != nullptr);
    uint64_t hdr[2];
    REQUIRE(std::fread(hdr, 8, 2, f) == 2);
    const size_t n = hdr[0], log_fft = hdr[1];
    std::vector<Affine> pts(n);
    std::vector<Scalar> sc(n), poly((size_t)1 << log_fft);
    Scalar r;
    REQUIRE(std::fread(pts.data(), 64, n, f) == n);
    REQUIRE(std::fread(sc.data(), 32, n, f) == n);
    REQUIRE(std::fread(poly.data(), 32, poly.size(), f) == poly.size());
    REQUIRE(std::fread(r.data(), 32, 1, f) == 1);
    std::fclose(f);

    // ---- host-only behaviour (no device needed)
    CommitmentKey ck(Curve::Bn256G1, pts);
    REQUIRE(ck.len() == n && !ck.is_empty() && CommitmentKey::default_value().is_identity());
    const size_t k = fft::log2_len(n);
    ck.save_to_file(dir + "/key.bin");
    CommitmentKey back = CommitmentKey::load_from_file(Curve::Bn256G1, dir + "/key.bin", k);
    REQUIRE(back == ck);  // commitment::file_tests::consistency
    try {
        CommitmentKey::load_from_file(Curve::Bn256G1, dir + "/key.bin", k + 1);
        REQUIRE(false);
    } catch (const IoError& e) { REQUIRE(e.kind == IoError::UnexpectedEof); }
    try {
        std::vector<Scalar> too_long(n + 1);
        ck.commit(too_long);
        REQUIRE(false);
    } catch (const TooLongInput& e) {
        REQUIRE(e.input_len == n + 1 && e.limit == n);
        std::printf("too_long %s\n", e.what());
    }
    for (uint32_t kk = 0; kk <= fft::S; kk++) {
        const Scalar w = fft::get_omega_or_inv(kk, false), wi = fft::get_omega_or_inv(kk, true), d = fft::get_ifft_divisor(kk);
        uint64_t line[12];
        for (int i = 0; i < 4; i++) { line[i] = w[i]; line[4 + i] = wi[i]; line[8 + i] = d[i]; }
        hex(("omega " + std::to_string(kk)).c_str(), line, 12);
    }
    try {
        fft::get_omega_or_inv(fft::S + 1, false);
        REQUIRE(false);
    } catch (const std::invalid_argument&) {}
    // setup_smallest_key's size rule (src/commitment.rs:172-186) at the benches' shapes
    REQUIRE(smallest_power(12, 17) == 21 && smallest_power(7, 17) == 20 && smallest_power(16, 17) == 21 && smallest_power(0, 17) == 0);
    REQUIRE(smallest_key_log2(17, 12, 0, 0, 26) == 22 && smallest_key_log2(17, 7, 0, 0, 15) == 21 && smallest_key_log2(10, 3, 1, 2, 1) == 13);
    std::printf("host ok\n");

    // ---- everything that computes goes through libsirius_b200.so
    try {
        const Affine c = ck.commit(sc);
        hex("commit", reinterpret_cast<const uint64_t*>(&c), 8);
        const Affine c_prefix = ck.commit(sc.data(), n / 2);  // v.len() < ck.len()
        hex("commit_prefix", reinterpret_cast<const uint64_t*>(&c_prefix), 8);
        CommitmentKey cached = CommitmentKey::load_or_setup_cache(Curve::Bn256G1, dir, "lbl", k, [&](size_t, const std::string&) { return pts; });
        CommitmentKey cached2 = CommitmentKey::load_or_setup_cache(Curve::Bn256G1, dir, "lbl", k);  // now from the file, validated on the curve
        REQUIRE(cached2 == cached && cached2 == ck);
        std::vector<Scalar> a = poly;
        fft::fft(a);
        hex("fft", reinterpret_cast<const uint64_t*>(a.data()), a.size() * 4);
        fft::ifft(a);
        REQUIRE(a == poly);
        a = poly;
        fft::coset_fft(a);
        hex("coset_fft", reinterpret_cast<const uint64_t*>(a.data()), a.size() * 4);
        fft::coset_ifft(a);
        REQUIRE(a == poly);
        RelaxedPlonkWitness acc{SB_FIELD_FR, sc, poly};
        std::vector<Scalar> W2(sc.rbegin(), sc.rend());
        std::vector<std::vector<Scalar>> T = {poly, std::vector<Scalar>(poly.rbegin(), poly.rend())};
        const RelaxedPlonkWitness folded = acc.fold(W2, T, r);
        hex("fold_W", reinterpret_cast<const uint64_t*>(folded.W.data()), folded.W.size() * 4);
        hex("fold_E", reinterpret_cast<const uint64_t*>(folded.E.data()), folded.E.size() * 4);
        std::printf("device ok\n");
    } catch (const Error& e) {
        std::printf("device_error %d %s\n", e.code, e.what());
    }
    return 0;
}
