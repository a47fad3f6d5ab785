// Prints the calculation lists include/sirius_b200_expr.hpp compiles for the MainGate structures of the benches, for
// tests/test_zz_cpp_mirror.py to compare with the Python mirror and the oracle.  Host only (no library call).
//   cpp_expr_check T1[,T2...]      e.g. 5,3 = MainGate<5> + MainGate<3> (sangria_poseidon primary)
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/sirius_b200_expr.hpp"

using namespace sirius_b200;

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::vector<size_t> T_list;
    for (char* tok = std::strtok(argv[1], ","); tok; tok = std::strtok(nullptr, ",")) T_list.push_back((size_t)std::atoi(tok));
    size_t nfix = 0, nadv = 0;
    for (size_t T : T_list) { nfix += 2 * T + 5; nadv += T + 2; }
    std::vector<Expr> gates;
    size_t fb = 0, ab = 0;
    for (size_t T : T_list) {
        gates.push_back(main_gate_expression(T, fb, ab, 0, nfix));
        fb += 2 * T + 5;
        ab += T + 2;
    }
    QueryIndexContext ctx;
    ctx.num_fixed = nfix;
    ctx.num_advice = nadv;
    const CompressedGates cg = CompressedGates::create(gates, ctx);
    std::printf("degree %zu num_challenges %zu\n", cg.degree, cg.ctx.num_challenges);
    for (int which = 0; which < 2; which++) {  // the compressed gate (is_sat) and its homogeneous form (cross terms)
        const GraphEvaluator ev = GraphEvaluator::create(which ? cg.homogeneous_expr : cg.compressed, Modulus::fr());
        std::printf("program %d calcs %zu intermediates %u\n", which, ev.calculations.size(), ev.num_intermediates);
        std::printf("rotations");
        for (int32_t r : ev.rotations) std::printf(" %d", r);
        std::printf("\nconstants");
        for (const Scalar& c : ev.constants) std::printf(" %016llx%016llx%016llx%016llx", (unsigned long long)c[3], (unsigned long long)c[2], (unsigned long long)c[1], (unsigned long long)c[0]);
        std::printf("\n");
        for (const CalculationInfo& c : ev.calculations) {
            if (c.has_b) std::printf("calc %d %d %u %u %d %u %u %u\n", c.op, c.a.kind, c.a.index, c.a.rot, c.b.kind, c.b.index, c.b.rot, c.target);
            else std::printf("calc %d %d %u %u - %u\n", c.op, c.a.kind, c.a.index, c.a.rot, c.target);
        }
        // the ABI form: same list, operand b only for binary opcodes
        const std::vector<sb_calc> abi = ev.to_sb_calcs();
        const std::vector<Scalar> cm = ev.constants_mont();
        if (abi.size() != ev.calculations.size() || cm.size() != ev.constants.size()) return 1;
        std::printf("const2_mont %016llx%016llx%016llx%016llx\n", (unsigned long long)cm[2][3], (unsigned long long)cm[2][2], (unsigned long long)cm[2][1], (unsigned long long)cm[2][0]);
    }
    // a Negated constant and a Scaled node, which the MainGate does not contain
    {
        const Expr e = scaled(polynomial(3, 1) - constant(small(5)), small(7)) + (-constant(small(9))) * challenge(0);
        const GraphEvaluator ev = GraphEvaluator::create(e, Modulus::fq());
        std::printf("extra calcs %zu\n", ev.calculations.size());
        std::printf("constants");
        for (const Scalar& c : ev.constants) std::printf(" %016llx%016llx%016llx%016llx", (unsigned long long)c[3], (unsigned long long)c[2], (unsigned long long)c[1], (unsigned long long)c[0]);
        std::printf("\n");
        for (const CalculationInfo& c : ev.calculations) {
            if (c.has_b) std::printf("calc %d %d %u %u %d %u %u %u\n", c.op, c.a.kind, c.a.index, c.a.rot, c.b.kind, c.b.index, c.b.rot, c.target);
            else std::printf("calc %d %d %u %u - %u\n", c.op, c.a.kind, c.a.index, c.a.rot, c.target);
        }
    }
    return 0;
}
