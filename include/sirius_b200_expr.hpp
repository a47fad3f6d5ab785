// sirius_b200_expr.hpp -- C++17 mirror of the reference's expression IR for the part the GPU path consumes (the
// compiled-language twin of sirius_b200/polynomial.py; in the real integration the Rust crate keeps these and only the
// compiled calculation list crosses the C ABI through sb_expr_compile):
//
//   Expression, QueryIndexContext            reference src/polynomial/expression.rs:38-120
//   Expression::homogeneous                  src/polynomial/expression.rs:356-429
//   challenge_in_degree                      src/polynomial/expression.rs:503-515
//   compress_expression                      src/plonk/util.rs:35-55
//   CompressedGates::new                     src/plonk/mod.rs:68-121
//   GraphEvaluator::new (compile only)       src/polynomial/graph_evaluator.rs:57-89, 164-351
//   main_gate_expression                     src/main_gate.rs:535-583 as Expression::from_halo2_expr sees it
//   Program (upload)                         -> sb_expr_compile; VanillaFS::commit_cross_terms -> sb_cross_terms + sb_msm_batch
//
// Field constants are canonical integers (4 x u64 little-endian limbs, NOT Montgomery); conversion happens when a program
// is uploaded.  tests/test_zz_cpp_mirror.py checks that the calculation lists this compiler emits equal the ones the
// Python mirror and the oracle emit for the MainGate structures of the benches.
#pragma once
#include <memory>
#include <set>
#include <tuple>

#include "sirius_b200.hpp"

namespace sirius_b200 {

// ------------------------------------------------------------------------------------------------ canonical integers mod p
struct Modulus {
    int field;  // SB_FIELD_FR / SB_FIELD_FQ
    Scalar p;
    static Modulus fr() { return {SB_FIELD_FR, {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}}; }
    static Modulus fq() { return {SB_FIELD_FQ, {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}}; }
    Scalar neg(const Scalar& a) const {  // p - a (0 for 0); a < p
        if ((a[0] | a[1] | a[2] | a[3]) == 0) return a;
        Scalar r;
        unsigned __int128 borrow = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 d = (unsigned __int128)p[i] - a[i] - (uint64_t)borrow;
            r[i] = (uint64_t)d;
            borrow = (d >> 64) & 1;
        }
        return r;
    }
    Scalar to_mont(const Scalar& canonical) const {
        if (field == SB_FIELD_FR) return fft::from_fe(sb::to_mont(fft::to_fe(canonical)));
        sb::Fq f;
        for (int i = 0; i < 4; i++) { f.v[2 * i] = (uint32_t)canonical[i]; f.v[2 * i + 1] = (uint32_t)(canonical[i] >> 32); }
        f = sb::to_mont(f);
        Scalar s;
        for (int i = 0; i < 4; i++) s[i] = (uint64_t)f.v[2 * i] | ((uint64_t)f.v[2 * i + 1] << 32);
        return s;
    }
};
inline Scalar small(uint64_t v) { return Scalar{v, 0, 0, 0}; }

// ------------------------------------------------------------------------------------------------ Expression
struct QueryIndexContext {
    size_t num_selectors = 0, num_fixed = 0, num_advice = 0, num_challenges = 0, num_lookups = 0;
    size_t num_fold_vars() const { return num_advice + num_lookups * 5; }
};

struct Expression;
using Expr = std::shared_ptr<const Expression>;
struct Expression {
    enum Kind { Constant, Polynomial, Challenge, Negated, Sum, Product, Scaled } kind;
    Scalar value{};          // Constant value / Scaled factor (canonical)
    size_t index = 0;        // Polynomial column / Challenge index
    int32_t rotation = 0;    // Polynomial
    Expr a, b;
};
inline Expr constant(const Scalar& v) { return std::make_shared<Expression>(Expression{Expression::Constant, v, 0, 0, nullptr, nullptr}); }
inline Expr polynomial(size_t index, int32_t rotation = 0) { return std::make_shared<Expression>(Expression{Expression::Polynomial, {}, index, rotation, nullptr, nullptr}); }
inline Expr challenge(size_t index) { return std::make_shared<Expression>(Expression{Expression::Challenge, {}, index, 0, nullptr, nullptr}); }
// operator overloads exactly as expression.rs:455-499
inline Expr operator-(const Expr& e) { return std::make_shared<Expression>(Expression{Expression::Negated, {}, 0, 0, e, nullptr}); }
inline Expr operator+(const Expr& l, const Expr& r) { return std::make_shared<Expression>(Expression{Expression::Sum, {}, 0, 0, l, r}); }
inline Expr operator-(const Expr& l, const Expr& r) { return l + (-r); }
inline Expr operator*(const Expr& l, const Expr& r) { return std::make_shared<Expression>(Expression{Expression::Product, {}, 0, 0, l, r}); }
inline Expr scaled(const Expr& e, const Scalar& k) { return std::make_shared<Expression>(Expression{Expression::Scaled, k, 0, 0, e, nullptr}); }

inline void collect_challenges(const Expr& e, std::set<size_t>& out) {
    switch (e->kind) {
        case Expression::Challenge: out.insert(e->index); break;
        case Expression::Negated: case Expression::Scaled: collect_challenges(e->a, out); break;
        case Expression::Sum: case Expression::Product: collect_challenges(e->a, out); collect_challenges(e->b, out); break;
        default: break;
    }
}
inline size_t num_challenges(const Expr& e) {
    std::set<size_t> s;
    collect_challenges(e, s);
    return s.size();
}

// challenge^degree as a left-nested product (expression.rs:503-515)
inline Expr challenge_in_degree(size_t new_challenge_index, size_t degree) {
    const Expr ch = challenge(new_challenge_index);
    Expr res = ch;
    for (size_t i = 2; i <= degree; i++) res = res * ch;
    return res;
}

// expression.rs:356-429 -> (expr, degree); the homogenising challenge has index ctx.num_challenges
inline std::pair<Expr, size_t> homogeneous(const Expr& e, const QueryIndexContext& ctx) {
    const size_t new_ch = ctx.num_challenges;
    switch (e->kind) {
        case Expression::Constant: return {e, 0};
        case Expression::Polynomial: return {e, e->index >= ctx.num_selectors + ctx.num_fixed ? 1u : 0u};
        case Expression::Challenge: return {e, 1};
        case Expression::Negated: {
            auto [x, d] = homogeneous(e->a, ctx);
            return {-x, d};
        }
        case Expression::Sum: {
            auto [l, ld] = homogeneous(e->a, ctx);
            auto [r, rd] = homogeneous(e->b, ctx);
            if (ld > rd) return {l + (r * challenge_in_degree(new_ch, ld - rd)), ld};
            if (ld < rd) return {(l * challenge_in_degree(new_ch, rd - ld)) + r, rd};
            return {l + r, ld};
        }
        case Expression::Product: {
            auto [l, ld] = homogeneous(e->a, ctx);
            auto [r, rd] = homogeneous(e->b, ctx);
            return {l * r, ld + rd};
        }
        case Expression::Scaled: {
            auto [x, d] = homogeneous(e->a, ctx);
            return {scaled(x, e->value), d};
        }
    }
    throw std::logic_error("homogeneous: bad expression");
}

// plonk/util.rs:35-55: P_n + (...(P_1 + 0*y)*y...)*y for n > 1
inline Expr compress_expression(const std::vector<Expr>& exprs, size_t challenge_index) {
    const Expr y = challenge(challenge_index);
    if (exprs.size() > 1) {
        Expr acc = constant(small(0));
        for (const Expr& e : exprs) acc = e + (acc * y);
        return acc;
    }
    return exprs.empty() ? constant(small(0)) : exprs[0];
}

struct CompressedGates {  // plonk/mod.rs:68-121 (without the lazily grouped polynomial, which only the CPU reference needs)
    Expr compressed, homogeneous_expr;
    size_t degree;
    QueryIndexContext ctx;
    static CompressedGates create(const std::vector<Expr>& gates, QueryIndexContext ctx) {
        const Expr compressed = compress_expression(gates, ctx.num_challenges);
        ctx.num_challenges = num_challenges(compressed);
        auto [hom, deg] = homogeneous(compressed, ctx);
        ctx.num_challenges = num_challenges(hom);
        return CompressedGates{compressed, hom, deg, ctx};
    }
};

// ------------------------------------------------------------------------------------------------ GraphEvaluator (compile)
struct ValueSource {  // graph_evaluator.rs:57-68; the declaration order of the kinds is the PartialOrd order used to
    int kind;         // canonicalise Add / Mul operands (:304-314, :333-337)
    uint32_t index, rot;
    bool operator==(const ValueSource& o) const { return kind == o.kind && index == o.index && rot == o.rot; }
    bool operator<=(const ValueSource& o) const { return std::tie(kind, index, rot) <= std::tie(o.kind, o.index, o.rot); }
};
struct CalculationInfo {  // `target = op(a, b)` (graph_evaluator.rs:72-89, 152-156)
    int op;
    ValueSource a, b;
    bool has_b;
    uint32_t target;
};

class GraphEvaluator {
   public:
    std::vector<Scalar> constants{small(0), small(1), small(2)};  // graph_evaluator.rs:186
    std::vector<int32_t> rotations;
    std::vector<CalculationInfo> calculations;
    uint32_t num_intermediates = 0;
    Modulus modulus;

    static GraphEvaluator create(const Expr& expr, const Modulus& m) {  // GraphEvaluator::new (:196-203)
        GraphEvaluator g(m);
        const ValueSource vs = g.add_expression(expr);
        g.add_calculation(SB_OP_STORE, vs, nullptr);
        return g;
    }

    // the op list as the C ABI takes it (sb_calc), constants in Montgomery form
    std::vector<sb_calc> to_sb_calcs() const {
        std::vector<sb_calc> out(calculations.size());
        for (size_t i = 0; i < calculations.size(); i++) {
            const CalculationInfo& c = calculations[i];
            sb_calc s{};
            s.opcode = (uint8_t)c.op;
            s.a_kind = (uint8_t)c.a.kind; s.a_index = c.a.index; s.a_rot = c.a.rot;
            if (c.has_b && c.op <= SB_OP_MUL) { s.b_kind = (uint8_t)c.b.kind; s.b_index = c.b.index; s.b_rot = c.b.rot; }
            s.target = c.target;
            out[i] = s;
        }
        return out;
    }
    std::vector<Scalar> constants_mont() const {
        std::vector<Scalar> out;
        for (const Scalar& c : constants) out.push_back(modulus.to_mont(c));
        return out;
    }

   private:
    explicit GraphEvaluator(const Modulus& m) : modulus(m) {}
    uint32_t add_rotation(int32_t rot) {
        for (size_t i = 0; i < rotations.size(); i++)
            if (rotations[i] == rot) return (uint32_t)i;
        rotations.push_back(rot);
        return (uint32_t)rotations.size() - 1;
    }
    ValueSource add_constant(const Scalar& c) {  // c canonical (< p)
        for (size_t i = 0; i < constants.size(); i++)
            if (constants[i] == c) return {SB_VS_CONSTANT, (uint32_t)i, 0};
        constants.push_back(c);
        return {SB_VS_CONSTANT, (uint32_t)constants.size() - 1, 0};
    }
    ValueSource add_calculation(int op, const ValueSource& a, const ValueSource* b) {  // `find` of an equal calculation (:241-258)
        for (const CalculationInfo& c : calculations)
            if (c.op == op && c.a == a && c.has_b == (b != nullptr) && (!b || c.b == *b)) return {SB_VS_INTERMEDIATE, c.target, 0};
        const uint32_t target = num_intermediates++;
        calculations.push_back(CalculationInfo{op, a, b ? *b : ValueSource{0, 0, 0}, b != nullptr, target});
        return {SB_VS_INTERMEDIATE, target, 0};
    }
    ValueSource add_expression(const Expr& e) {  // graph_evaluator.rs:260-351
        const ValueSource ZERO{SB_VS_CONSTANT, 0, 0}, ONE{SB_VS_CONSTANT, 1, 0}, TWO{SB_VS_CONSTANT, 2, 0};
        switch (e->kind) {
            case Expression::Constant: return add_constant(e->value);
            case Expression::Polynomial: {
                const uint32_t rot_idx = add_rotation(e->rotation);
                return add_calculation(SB_OP_STORE, {SB_VS_POLY, (uint32_t)e->index, rot_idx}, nullptr);
            }
            case Expression::Challenge: return add_calculation(SB_OP_STORE, {SB_VS_CHALLENGE, (uint32_t)e->index, 0}, nullptr);
            case Expression::Negated: {
                if (e->a->kind == Expression::Constant) return add_constant(modulus.neg(e->a->value));
                const ValueSource ra = add_expression(e->a);
                if (ra == ZERO) return ra;
                return add_calculation(SB_OP_NEGATE, ra, nullptr);
            }
            case Expression::Sum: {
                if (e->b->kind == Expression::Negated) {
                    const ValueSource ra = add_expression(e->a);
                    const ValueSource rb = add_expression(e->b->a);
                    if (ra == ZERO) return add_calculation(SB_OP_NEGATE, rb, nullptr);
                    if (rb == ZERO) return ra;
                    return add_calculation(SB_OP_SUB, ra, &rb);
                }
                const ValueSource ra = add_expression(e->a);
                const ValueSource rb = add_expression(e->b);
                return ra <= rb ? add_calculation(SB_OP_ADD, ra, &rb) : add_calculation(SB_OP_ADD, rb, &ra);
            }
            case Expression::Product: {
                const ValueSource ra = add_expression(e->a);
                const ValueSource rb = add_expression(e->b);
                if (ra == ZERO || rb == ZERO) return ZERO;
                if (ra == ONE) return rb;
                if (rb == ONE) return ra;
                if (ra == TWO) return add_calculation(SB_OP_DOUBLE, rb, nullptr);
                if (rb == TWO) return add_calculation(SB_OP_DOUBLE, ra, nullptr);
                if (ra == rb) return add_calculation(SB_OP_SQUARE, ra, nullptr);
                return ra <= rb ? add_calculation(SB_OP_MUL, ra, &rb) : add_calculation(SB_OP_MUL, rb, &ra);
            }
            case Expression::Scaled: {
                const Scalar& f = e->value;
                if (f == small(0)) return ZERO;
                if (f == small(1)) return add_expression(e->a);
                const ValueSource cst = add_constant(f);
                const ValueSource ra = add_expression(e->a);
                return add_calculation(SB_OP_MUL, ra, &cst);
            }
        }
        throw std::logic_error("add_expression: bad expression");
    }
};

// The MainGate<T> custom gate as `Expression::from_halo2_expr` sees it (main_gate.rs:535-583, expression.rs:305-340).
// Fixed columns of this gate start at `fixed_base`, advice at `advice_base` (column order: q_1[T], q_5[T], q_m[2], q_i,
// q_o, rc / state[T], input, out).
inline Expr main_gate_expression(size_t T, size_t fixed_base, size_t advice_base, size_t num_selectors, size_t num_fixed_total) {
    auto fx = [&](size_t j) { return polynomial(num_selectors + fixed_base + j, 0); };
    auto ad = [&](size_t j) { return polynomial(num_selectors + num_fixed_total + advice_base + j, 0); };
    auto pow_5 = [](const Expr& v) {
        const Expr v2 = v * v;
        return v2 * v2 * v;
    };
    std::vector<Expr> state;
    for (size_t i = 0; i < T; i++) state.push_back(ad(i));
    const Expr inp = ad(T), out = ad(T + 1);
    const Expr q_i = fx(2 * T + 2), q_o = fx(2 * T + 3), rc = fx(2 * T + 4);
    Expr acc = fx(2 * T) * state[0] * state[1] + q_i * inp + rc + q_o * out;
    if (T >= 4) acc = fx(2 * T + 1) * state[2] * state[3] + acc;
    for (size_t i = 0; i < T; i++) acc = acc + (fx(i) * state[i] + fx(T + i) * pow_5(state[i]));
    return acc;
}

// ------------------------------------------------------------------------------------------------ device side
class Program {  // a GraphEvaluator uploaded through sb_expr_compile (RAII)
   public:
    explicit Program(const GraphEvaluator& ev) {
        const std::vector<sb_calc> calcs = ev.to_sb_calcs();
        const std::vector<Scalar> consts = ev.constants_mont();
        const std::vector<int32_t> rots = ev.rotations.empty() ? std::vector<int32_t>{0} : ev.rotations;
        check(sb_expr_compile(ev.modulus.field, calcs.data(), calcs.size(), reinterpret_cast<const uint64_t*>(consts.data()), consts.size(),
                              rots.data(), ev.rotations.size(), &handle_));
    }
    Program(const Program&) = delete;
    Program& operator=(const Program&) = delete;
    ~Program() { if (handle_) sb_expr_free(handle_); }
    sb_prog_t handle() const { return handle_; }

   private:
    sb_prog_t handle_ = nullptr;
};

// The part of plonk::PlonkStructure (src/plonk/mod.rs:123-160) the Sangria cross terms read: table size, selector and
// fixed columns (registered on the device once), the compressed custom gates and their compiled homogeneous form.
class PlonkStructure {
   public:
    PlonkStructure(const Modulus& m, uint32_t k, const std::vector<std::vector<uint8_t>>& selectors, const std::vector<std::vector<Scalar>>& fixed_columns,
                   size_t num_advice_columns, size_t num_lookups, const CompressedGates& gates)
        : modulus(m), k(k), num_advice_columns(num_advice_columns), num_lookups(num_lookups), degree(gates.degree),
          num_challenges(gates.ctx.num_challenges), hom_(GraphEvaluator::create(gates.homogeneous_expr, m)) {
        const size_t n = (size_t)1 << k;
        std::vector<const uint8_t*> sel;
        std::vector<const uint64_t*> fix;
        for (const auto& s : selectors) {
            if (s.size() != n) throw std::invalid_argument("PlonkStructure: selector column length != 2^k");
            sel.push_back(s.data());
        }
        for (const auto& f : fixed_columns) {
            if (f.size() != n) throw std::invalid_argument("PlonkStructure: fixed column length != 2^k");
            fix.push_back(reinterpret_cast<const uint64_t*>(f.data()));
        }
        check(sb_columns_register(m.field, k, sel.data(), sel.size(), fix.data(), fix.size(), &cols_));
    }
    PlonkStructure(const PlonkStructure&) = delete;
    PlonkStructure& operator=(const PlonkStructure&) = delete;
    ~PlonkStructure() { if (cols_) sb_columns_release(cols_); }

    Modulus modulus;
    uint32_t k;
    size_t num_advice_columns, num_lookups, degree, num_challenges;
    sb_columns_t columns() const { return cols_; }
    sb_prog_t homogeneous_program() const { return hom_.handle(); }

   private:
    Program hom_;
    sb_columns_t cols_ = nullptr;
};

struct VanillaFS {
    // `commit_cross_terms(ck, S, U1, W1, U2, W2) -> (CrossTerms, CrossTermCommits)` (src/nifs/sangria/mod.rs:102-158).
    // W1 / W2: the witness round vectors (column-major, as concatenate_with_padding lays them out); the challenge
    // vectors are U1.challenges ++ [U1.u] and U2.challenges ++ [1] (:113-118).
    static std::pair<std::vector<std::vector<Scalar>>, std::vector<Affine>> commit_cross_terms(
        const CommitmentKey& ck, const PlonkStructure& S, const std::vector<Scalar>& U1_challenges, const Scalar& U1_u,
        const std::vector<std::vector<Scalar>>& W1, const std::vector<Scalar>& U2_challenges, const std::vector<std::vector<Scalar>>& W2) {
        std::vector<Scalar> c1 = U1_challenges, c2 = U2_challenges;
        c1.push_back(U1_u);
        c2.push_back(S.modulus.to_mont(small(1)));
        if (c1.size() != c2.size()) throw std::invalid_argument("commit_cross_terms: challenge counts differ");
        auto rounds = [](const std::vector<std::vector<Scalar>>& W, std::vector<const uint64_t*>& ptrs, std::vector<size_t>& lens) {
            for (const auto& w : W) {
                ptrs.push_back(reinterpret_cast<const uint64_t*>(w.data()));
                lens.push_back(w.size());
            }
        };
        std::vector<const uint64_t*> p1, p2;
        std::vector<size_t> l1, l2;
        rounds(W1, p1, l1);
        rounds(W2, p2, l2);
        const size_t n = (size_t)1 << S.k;
        std::vector<std::vector<Scalar>> T(S.degree, std::vector<Scalar>(n));
        std::vector<uint64_t*> outp;
        for (auto& t : T) outp.push_back(reinterpret_cast<uint64_t*>(t.data()));
        check(sb_cross_terms(S.homogeneous_program(), (uint32_t)S.degree, S.columns(), (uint32_t)S.num_advice_columns, (uint32_t)S.num_lookups,
                             p1.data(), l1.data(), p1.size(), p2.data(), l2.data(), p2.size(), reinterpret_cast<const uint64_t*>(c1.data()),
                             reinterpret_cast<const uint64_t*>(c2.data()), c1.size(), outp.data()));
        std::vector<Affine> commits = ck.commit_batch(T);  // cross_terms.iter().map(|v| ck.commit(v)) (:151-154)
        return {std::move(T), std::move(commits)};
    }
};

}  // namespace sirius_b200
