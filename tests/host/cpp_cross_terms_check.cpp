// VanillaFS::commit_cross_terms through the C++ mirror (include/sirius_b200_expr.hpp): builds the MainGate structure,
// registers it, computes the cross terms and their commitments, prints them as hex for tests/test_zz_cpp_mirror.py.
//   cpp_cross_terms_check <inputs.bin>
// inputs.bin (u64 words): n_T, T_1..T_nT, k, n_key | key points (64 B) | fixed columns [nfix][2^k] | W1 [nadv*2^k] | W2 |
//                         U1_challenges [nch-1] | U1_u | U2_challenges [nch-1]          (field Fr / curve bn256)
#include <cstdio>
#include <vector>

#include "../../include/sirius_b200_expr.hpp"

using namespace sirius_b200;

static void hex(const char* tag, size_t j, const uint64_t* w, size_t n_words) {
    std::printf("%s %zu", tag, j);
    for (size_t i = 0; i < n_words; i++) std::printf(" %016llx", (unsigned long long)w[i]);
    std::printf("\n");
}
template <class T>
static bool rd(std::FILE* f, T* p, size_t count) { return std::fread(p, sizeof(T), count, f) == count; }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    uint64_t nT;
    if (!rd(f, &nT, 1)) return 2;
    std::vector<uint64_t> T_list(nT);
    uint64_t k, n_key;
    if (!rd(f, T_list.data(), nT) || !rd(f, &k, 1) || !rd(f, &n_key, 1)) return 2;
    const size_t n = (size_t)1 << k;
    size_t nfix = 0, nadv = 0;
    for (uint64_t T : T_list) { nfix += 2 * T + 5; nadv += T + 2; }
    std::vector<Expr> gates;
    size_t fb = 0, ab = 0;
    for (uint64_t T : T_list) {
        gates.push_back(main_gate_expression(T, fb, ab, 0, nfix));
        fb += 2 * T + 5;
        ab += T + 2;
    }
    QueryIndexContext ctx;
    ctx.num_fixed = nfix;
    ctx.num_advice = nadv;
    const CompressedGates cg = CompressedGates::create(gates, ctx);
    const size_t nch = cg.ctx.num_challenges;
    std::vector<Affine> key(n_key);
    std::vector<std::vector<Scalar>> fixed(nfix, std::vector<Scalar>(n));
    std::vector<Scalar> W1(nadv * n), W2(nadv * n), U1c(nch - 1), U2c(nch - 1);
    Scalar U1u;
    if (!rd(f, key.data(), n_key)) return 2;
    for (auto& col : fixed)
        if (!rd(f, col.data(), n)) return 2;
    if (!rd(f, W1.data(), W1.size()) || !rd(f, W2.data(), W2.size()) || !rd(f, U1c.data(), U1c.size()) || !rd(f, &U1u, 1) ||
        !rd(f, U2c.data(), U2c.size()))
        return 2;
    std::fclose(f);
    std::printf("degree %zu num_challenges %zu\n", cg.degree, nch);
    try {
        CommitmentKey ck(Curve::Bn256G1, key);
        PlonkStructure S(Modulus::fr(), (uint32_t)k, {}, fixed, nadv, 0, cg);
        auto [T, commits] = VanillaFS::commit_cross_terms(ck, S, U1c, U1u, {W1}, U2c, {W2});
        for (size_t j = 0; j < T.size(); j++) {
            hex("cross_term", j, reinterpret_cast<const uint64_t*>(T[j].data()), T[j].size() * 4);
            hex("cross_commit", j, reinterpret_cast<const uint64_t*>(&commits[j]), 8);
        }
        std::printf("device ok\n");
    } catch (const Error& e) {
        std::printf("device_error %d %s\n", e.code, e.what());
    }
    return 0;
}
