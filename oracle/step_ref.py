"""CPU restatement of one Sangria `fold_step` hot path (TEST INFRASTRUCTURE: the checker of tests/ and bench.py's
`--verify`, and the timed body of bench.py's CPU arm; never imported by sirius_b200/).

Literal reference order (src/ivc/sangria/incrementally_verifiable_computation.rs:428-635):
    prove(secondary) ; commit W(primary) ; prove(primary) ; commit W(secondary)
with `prove` = VanillaFS::commit_cross_terms (src/nifs/sangria/mod.rs:102-158: d GroupedPoly sweeps of the literal
GraphEvaluator + d `CommitmentKey::commit`) followed by RelaxedPlonkWitness::fold (accumulator.rs:363-404).
"""
from __future__ import annotations

import ctypes
from typing import Dict

import numpy as np

import oracle
from oracle import expr_ref as E
from oracle import pyref as R

_u64p = ctypes.POINTER(ctypes.c_uint64)
_EV_CACHE = {}


def evaluators(side):
    """GraphEvaluator per cross term T_1..T_d of one side's compressed MainGate expression (None = zero term)."""
    key = (side["name"], side.get("kind"), tuple(side["T_list"]))
    if key not in _EV_CACHE:
        if side.get("kind") == "tiny":   # the Cyclefold support circuit's gate (tiny_gate.rs:38-84)
            nsel, nfix, nadv, gates = 1, 4, 3, [E.tiny_gate_expression()]
        else:
            nsel = 0
            nfix = sum(2 * T + 5 for T in side["T_list"])
            nadv = sum(T + 2 for T in side["T_list"])
            gates, fb, ab = [], 0, 0
            for T in side["T_list"]:
                gates.append(E.main_gate_expression(T, fb, ab, 0, nfix))
                fb += 2 * T + 5
                ab += T + 2
        cg = E.CompressedGates(gates, E.Ctx(num_selectors=nsel, num_fixed=nfix, num_advice=nadv))
        f = side["field"]
        evs = [None if ex is None else E.GraphEvaluator(ex, R.MODULUS[f]) for ex in cg.grouped()[1:]]
        _EV_CACHE[key] = (evs, cg.ctx.num_challenges - 1)
    return _EV_CACHE[key]


def prove(inp: Dict, bases: np.ndarray, threads: int) -> Dict:
    """inp: side, k, nadv, fixed[list of [n,4]], W1, E1, W2 ([A*n,4] column-major), c1, c2 ([nch,4]), u1, r ([4])."""
    side, k, nadv = inp["side"], inp["k"], inp["nadv"]
    f, curve = side["field"], side["curve"]
    n = 1 << k
    lib = oracle.lib()
    evs, nch = evaluators(side)
    one = R.to_mont_limbs([1], R.MODULUS[f])
    ch = np.ascontiguousarray(np.concatenate([inp["c1"].reshape(-1, 4), inp["u1"].reshape(1, 4), inp["c2"].reshape(-1, 4), one]), dtype=np.uint64)
    W1, W2, E1 = (np.ascontiguousarray(inp[x], dtype=np.uint64).reshape(-1, 4) for x in ("W1", "W2", "E1"))
    adv = [W1[i * n:(i + 1) * n] for i in range(nadv)] + [W2[i * n:(i + 1) * n] for i in range(nadv)]
    sel = inp.get("selectors", [])
    T = [np.zeros((n, 4), dtype=np.uint64) if ev is None else E.c_graph_evaluate(f, ev, sel, inp["fixed"], adv, ch, k, threads=threads) for ev in evs]
    commits = np.stack([oracle.msm(curve, t, bases, threads=threads) for t in T])
    r = np.ascontiguousarray(inp["r"], dtype=np.uint64).reshape(4)
    outw = np.zeros_like(W1)
    lib.so_axpy(f, W1.ctypes.data_as(_u64p), W2.ctypes.data_as(_u64p), r.ctypes.data_as(_u64p), outw.ctypes.data_as(_u64p), ctypes.c_size_t(nadv * n))
    ptrs = (_u64p * len(T))(*[t.ctypes.data_as(_u64p) for t in T])
    oute = np.zeros_like(E1)
    lib.so_error_fold(f, E1.ctypes.data_as(_u64p), ptrs, ctypes.c_size_t(len(T)), r.ctypes.data_as(_u64p), oute.ctypes.data_as(_u64p), ctypes.c_size_t(n))
    return dict(commits_T=commits, W=outw, E=oute, T=T)


def fold_step(inputs: Dict[str, Dict], bases: Dict[str, np.ndarray], threads: int = 0) -> Dict[str, Dict]:
    """inputs / bases keyed "primary" / "secondary".  Returns per side commits_T [d,8], commit_W [8], W, E (folded)."""
    out = {}
    sec, prim = inputs["secondary"], inputs["primary"]
    out["secondary"] = prove(sec, bases["secondary"], threads)
    cw_p = oracle.msm(prim["side"]["curve"], prim["W2"], bases["primary"], threads=threads)
    out["primary"] = prove(prim, bases["primary"], threads)
    out["primary"]["commit_W"] = cw_p
    out["secondary"]["commit_W"] = oracle.msm(sec["side"]["curve"], sec["W2"], bases["secondary"], threads=threads)
    return out


def synthetic_inputs(sides, k: int, seed: int) -> Dict[str, Dict]:
    """Seeded uniform inputs of the bench shapes, for the CPU arm when no GPU snapshot is at hand."""
    n = 1 << k
    st = {}
    for side in sides:
        nfix = sum(2 * T + 5 for T in side["T_list"])
        nadv = sum(T + 2 for T in side["T_list"])
        _, nch = evaluators(side)
        f = side["field"]
        st[side["name"]] = dict(
            side=side, k=k, nadv=nadv, nfix=nfix,
            fixed=[oracle.random_field(f, seed + 31 * j + side["curve"], n) for j in range(nfix)],
            W1=oracle.random_field(f, seed + 1, nadv * n), W2=oracle.random_field(f, seed + 2, nadv * n), E1=oracle.random_field(f, seed + 3, n),
            c1=oracle.random_field(f, seed + 4, nch), c2=oracle.random_field(f, seed + 5, nch), u1=oracle.random_field(f, seed + 12, 1).reshape(4),
            r=oracle.random_field(f, seed + 6, 1).reshape(4),
        )
    return st


def bases_for(inputs: Dict[str, Dict]) -> Dict[str, np.ndarray]:
    return {name: oracle.running_bases(inp["side"]["curve"], inp["nadv"] << inp["k"]) for name, inp in inputs.items()}


def compare(results_gpu: Dict[str, Dict], results_cpu: Dict[str, Dict]) -> Dict:
    """Bit-for-bit comparison of the 13 commitments and the folded W / E.  Returns {"ok": bool, "checked": [...], "bad": [...]}."""
    checked, bad = [], []
    for name in ("secondary", "primary"):
        g, c = results_gpu[name], results_cpu[name]
        for key in ("commits_T", "commit_W", "W", "E"):
            a = np.asarray(g[key], dtype=np.uint64).reshape(-1)
            b = np.asarray(c[key], dtype=np.uint64).reshape(-1)
            tag = f"{name}.{key}"
            checked.append(tag)
            if a.shape != b.shape or not np.array_equal(a, b):
                bad.append(tag)
    return {"ok": not bad, "checked": checked, "bad": bad}


# ------------------------------------------------------------------------------------------------ Cyclefold next()
class _PgCtx:
    """The PolyContext values compute_K_from_G reads (poly/mod.rs:205-269, incl. the point-count-as-log of SURVEY F5)."""

    def __init__(self, traces_len: int, max_degree: int):
        self.instances_to_fold = traces_len + 1
        p = 1
        while p < traces_len * max_degree + 1:
            p <<= 1
        self.fft_points_count_G = p

    def lagrange_domain(self):
        return self.instances_to_fold.bit_length() - 1

    def fft_log_domain_size_K(self):
        v, p = max(self.fft_points_count_G + 1 - self.instances_to_fold, 0), 1
        while p < v:
            p <<= 1
        return p


def protogalaxy_prove(inp: Dict, side: Dict, threads: int = 0) -> Dict:
    """ProtoGalaxy::prove (src/nifs/protogalaxy/mod.rs:400-481) for one accumulator + one incoming trace of the circuit `side`
    describes (gate list by MainGate widths): poly_F (poly/mod.rs:68-203), poly_G (:308-425), poly_K (:205-269) and the folded
    witness (mod.rs:176-210).  inp: k, row_mode, fixed, nadv, W_acc, W_in, betas, delta, alpha, gamma."""
    from oracle import pg_fast as PF
    from oracle import pg_ref as PG
    from sirius_b200 import workload as WL   # side descriptions only (shapes), no device code

    k, nadv = inp["k"], inp["nadv"]
    mode = "correct" if inp["row_mode"] == 1 else "compat"
    gates, nfix, _ = WL.compressed_gates(side, E)
    S = PF.Structure(k, [], inp["fixed"], nadv, gates)
    ctx = E.Ctx(num_fixed=nfix, num_advice=nadv)
    max_degree = max(PG.gate_degree(g, ctx) for g in gates)
    t = S.betas_count()
    betas, delta, alpha, gamma = inp["betas"][:t], inp["delta"], inp["alpha"], inp["gamma"]
    poly_F = PF.compute_F(S, betas, delta, inp["W_acc"], [], mode, threads)
    bs = PG.beta_stroke(betas, alpha, delta)
    poly_G = PF.compute_G(S, max_degree, bs, inp["W_acc"], [], [inp["W_in"]], [[]], mode, threads)
    pctx = _PgCtx(1, max_degree)
    poly_K = PG.compute_K_from_G(pctx, poly_G, PG.poly_eval(poly_F, alpha))
    Lg = R.eval_lagrange_polys(pctx.lagrange_domain(), gamma)
    W = PF.fold_witness(inp["W_acc"], [inp["W_in"]], Lg)
    return dict(poly_F=poly_F, poly_G=poly_G, poly_K=poly_K, W=W)


def compare_protogalaxy(g: Dict, c: Dict) -> Dict:
    checked, bad = [], []
    for key in ("poly_F", "poly_G", "poly_K"):
        checked.append(key)
        if list(g[key]) != list(c[key]):
            bad.append(key)
    checked.append("W")
    a, b = np.asarray(g["W"], dtype=np.uint64).reshape(-1), np.asarray(c["W"], dtype=np.uint64).reshape(-1)
    if a.shape != b.shape or not np.array_equal(a, b):
        bad.append("W")
    return {"ok": not bad, "checked": checked, "bad": bad}


def cyclefold_step(inp: Dict, threads: int = 0) -> Dict:
    """The prover hot path of cyclefold::IVC::next (src/ivc/cyclefold/incrementally_verifiable_computation/mod.rs:210-335) on the
    inputs of sirius_b200.workload.CyclefoldStepWorkload.snapshot_inputs(): ProtoGalaxy::prove (F, G, K, fold_witness,
    src/nifs/protogalaxy/mod.rs:400-481), fold_support_circuit (:404-473), commit of the next primary trace."""
    from sirius_b200 import workload as WL   # side descriptions only (shapes), no device code

    pg = protogalaxy_prove(inp, WL.PRIMARY, threads)
    poly_F, poly_G, poly_K, W = pg["poly_F"], pg["poly_G"], pg["poly_K"], pg["W"]
    k, nadv = inp["k"], inp["nadv"]
    sup = inp["support"]
    sup_bases = oracle.running_bases(sup["side"]["curve"], sup["nadv"] << sup["k"])
    sres = prove(sup, sup_bases, threads)
    sres["commit_W"] = oracle.msm(sup["side"]["curve"], sup["W2"], sup_bases, threads=threads)
    bases = oracle.running_bases(WL.PRIMARY["curve"], nadv << k)
    commit_W = oracle.msm(WL.PRIMARY["curve"], inp["W_in"], bases, threads=threads)
    return dict(poly_F=poly_F, poly_G=poly_G, poly_K=poly_K, W=W, commit_W=commit_W, support=sres)


def compare_cyclefold(g: Dict, c: Dict) -> Dict:
    checked, bad = [], []

    def chk(tag, a, b):
        checked.append(tag)
        if isinstance(a, list):
            ok = list(a) == list(b)
        else:
            a, b = np.asarray(a, dtype=np.uint64).reshape(-1), np.asarray(b, dtype=np.uint64).reshape(-1)
            ok = a.shape == b.shape and np.array_equal(a, b)
        if not ok:
            bad.append(tag)

    for key in ("poly_F", "poly_G", "poly_K", "W", "commit_W"):
        chk(key, g[key], c[key])
    for key in ("commits_T", "commit_W", "W", "E"):
        chk(f"support.{key}", g["support"][key], c["support"][key])
    return {"ok": not bad, "checked": checked, "bad": bad}
