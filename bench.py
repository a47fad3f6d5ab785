#!/usr/bin/env python
"""bench.py -- one Sangria IVC fold step's prover hot path at the shapes of benches/sangria_poseidon (k = 17).

A "step" is one pass of the hot path of `IVC::fold_step` (reference
src/ivc/sangria/incrementally_verifiable_computation.rs:428-635) over one batch of synthetic witness
columns, in the reference's call order:

  1. VanillaFS::prove, secondary side (grumpkin; A=7, F=15, 1 gate, d=5):
       commit_cross_terms = 5 cross-term vectors over 2^17 rows + 5 commits of 2^17 scalars, then W/E fold
  2. generate_plonk_trace, primary side (bn256): commit W, 12*2^17 = 1 572 864 scalars
  3. VanillaFS::prove, primary side (bn256; A=12, F=26, 2 gates, d=6): 6 cross terms + 6 commits + folds
  4. generate_plonk_trace, secondary side: commit W, 7*2^17 = 917 504 scalars

halo2 circuit synthesis (CPU, out of scope per SURVEY section 2) is not part of the step.  Every commitment is
copied back to the host and synchronised before the next stage, as the random oracle requires.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

`value`  : ms per step, all inputs resident in HBM.
`e2e`    : ms per step with that step's fresh witness columns copied host->device (pinned) inside the timed region.
N > 1    : rows are sharded across ranks (strong scaling); each commitment is one NCCL all-gather of 128-byte
           partial sums plus a combine kernel.
--impl reference : the CPU restatement of the reference (oracle/, OpenMP on all host cores) on the same shapes.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_TABLE = 17
PRIMARY = dict(name="primary", curve=0, field=0, T_list=[5, 3])   # bn256 / Fr : MainGate<5> + Poseidon MainGate<3>
SECONDARY = dict(name="secondary", curve=1, field=1, T_list=[5])  # grumpkin / Fq : MainGate<5> (trivial step circuit)
CK_LOG = 21                                                       # benches/sangria_poseidon.rs:26-30
METRIC = "sangria_poseidon k=17 IVC fold_step prover hot-path time"
SEED = 0x5349524955530000


def shapes(side):
    nfix = sum(2 * T + 5 for T in side["T_list"])
    nadv = sum(T + 2 for T in side["T_list"])
    return nfix, nadv


def workload_config(n_gpus):
    return {
        "workload": f"benches/sangria_poseidon k={K_TABLE} bn256/grumpkin: fold_step hot path "
                    f"(MSM {12 << K_TABLE} + 6x{1 << K_TABLE} bn256, MSM {7 << K_TABLE} + 5x{1 << K_TABLE} grumpkin, 11 cross-term vectors, W/E folds)",
        "k": K_TABLE,
        "ck_log2": CK_LOG,
        "sharding": "single GPU" if n_gpus == 1 else f"rows sharded over {n_gpus} ranks, all-gather of partial sums per commitment",
        "l2": "inputs larger than L2 (2 x 2 GiB window tables gathered at random; 0.4 GB of per-step scratch)",
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for nm, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def build_side_gpu(side, rank, world, stream):
    """Structure + key + device session for one curve, restricted to this rank's rows."""
    import numpy as np
    import torch

    import sirius_b200
    from sirius_b200 import _lib, curves, device
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG
    import ctypes

    nfix, nadv = shapes(side)
    gates, fb, ab = [], 0, 0
    for T in side["T_list"]:
        gates.append(P.main_gate_expression(T, fb, ab, 0, nfix))
        fb += 2 * T + 5
        ab += T + 2
    cg = P.CompressedGates.new(gates, P.QueryIndexContext(num_fixed=nfix, num_advice=nadv))
    from sirius_b200 import sharding

    n = 1 << K_TABLE
    row0, n_loc = sharding.row_slice(rank, world, n)
    k_loc = n_loc.bit_length() - 1
    modulus = curves.SCALAR_FIELD[side["curve"]]
    # fixed columns: synthetic uniform, generated in HBM then registered (register copies from host memory)
    fixed = [device.random_field_device(n_loc, SEED + 1000 * side["curve"] + 10 * rank + j).cpu().numpy().view(np.uint64) for j in range(nfix)]
    S = SG.PlonkStructure(side["field"], modulus, k_loc, [], fixed, nadv, 0, cg)
    if world > 1:
        sharding.check_rotations_row_local(P.GraphEvaluator.new(cg.homogeneous, modulus).rotations)
    # commitment key restricted to this rank's rows: ck[col * n + row] for row in the slice, column-major
    lib = _lib.load()
    assert nadv * n <= (1 << CK_LOG), "the 2^(k+4) key must cover W (only the prefix W needs is materialised)"
    d_bases = torch.zeros((nadv * n_loc, 8), dtype=torch.int64, device="cuda")
    g = curves.generator_limbs(side["curve"])
    for col, (first, count) in enumerate(sharding.key_segments(nadv, n, rank, world)):
        _lib.check(lib.sb_index_multiples_device(side["curve"], g.ctypes.data_as(_lib.u64p), first, count,
                                                 ctypes.c_void_p(d_bases.data_ptr() + col * n_loc * 64), ctypes.c_void_p(stream.cuda_stream)))
    stream.synchronize()
    # several window widths per key: each commit picks the cheapest one for its size (the W commits want wide
    # windows, the batched cross-term commits of 2^17/world scalars narrow ones)
    windows = [int(x) for x in os.environ.get("SB_BENCH_WINDOWS", "16,13,15,17").split(",")]
    ck = sirius_b200.CommitmentKey.from_device(side["curve"], d_bases.data_ptr(), nadv * n_loc, window_bits=windows[0], stream=stream.cuda_stream)
    for wb in windows[1:]:
        ck.add_window(wb, stream.cuda_stream)
    stream.synchronize()
    del d_bases
    sess = device.DeviceSangriaSide(S, ck, stream)
    with torch.cuda.stream(stream):
        sess.W_acc.copy_(device.random_field_device(nadv * n_loc, SEED + 7 + side["curve"] + 100 * rank))
        sess.E_acc.copy_(device.random_field_device(n_loc, SEED + 8 + side["curve"] + 100 * rank))
        sess.W_in.copy_(device.random_field_device(nadv * n_loc, SEED + 9 + side["curve"] + 100 * rank))
    host_W = sess.W_in.cpu().pin_memory()
    nch = cg.ctx.num_challenges - 1
    ch = device.random_field_device(2 * nch + 2, SEED + 11 + side["curve"]).cpu().numpy().view(np.uint64)
    extra = dict(host_W=host_W, c1=ch[:nch], c2=ch[nch:2 * nch], u1=ch[2 * nch], r=ch[2 * nch + 1], nadv=nadv, n_loc=n_loc)
    return sess, extra


class Combiner:
    """N > 1: all-gather the XYZZ partial sums of a commitment group and add them."""

    def __init__(self, world, stream):
        import torch

        self.world, self.stream, self.torch = world, stream, torch
        self.bufs = {}

    def _buffers(self, batch):
        torch = self.torch
        if batch not in self.bufs:
            self.bufs[batch] = (torch.zeros((batch, 16), dtype=torch.int64, device="cuda"),
                                torch.zeros((self.world, batch, 16), dtype=torch.int64, device="cuda"),
                                torch.zeros((batch, 8), dtype=torch.int64, device="cuda"))
        return self.bufs[batch]

    def commit(self, sess, d_scalars, n, batch, h_out):
        import ctypes

        import torch
        import torch.distributed as dist

        from sirius_b200 import _lib

        lib = _lib.load()
        part, gathered, out = self._buffers(batch)
        sess.ck.commit_batch_device(d_scalars, n, n, batch, 0, part.data_ptr(), self.stream.cuda_stream)
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(gathered, part)
        _lib.check(lib.sb_msm_combine_batch_device(sess.ck.curve, ctypes.c_void_p(gathered.data_ptr()), self.world, batch, batch,
                                                   ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(self.stream.cuda_stream)))
        with torch.cuda.stream(self.stream):
            h_out.copy_(out.view(h_out.shape), non_blocking=True)
        self.stream.synchronize()


def gpu_step(sides, extras, upload, combiner):
    """One fold_step hot path.  Returns bytes copied host->device."""
    prim, sec = sides
    ep, es = extras
    h2d = 0

    def prove(sess, ex):
        if combiner is None:
            sess.commit_cross_terms(ex["c1"], ex["u1"], ex["c2"])
        else:
            import ctypes
            import numpy as np

            from sirius_b200 import _lib

            lib = _lib.load()
            c1, c2 = sess.challenge_vectors(ex["c1"], ex["u1"], ex["c2"])
            _lib.check(lib.sb_cross_terms_device(sess.S._hom_prog._h, sess.d, sess.S._cols, sess._cols(sess.W_acc), sess._cols(sess.W_in), sess.A,
                                                 c1.ctypes.data_as(_lib.u64p), c2.ctypes.data_as(_lib.u64p), c1.shape[0],
                                                 ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(sess.stream.cuda_stream)))
            combiner.commit(sess, sess.T.data_ptr(), sess.n, sess.d, sess.h_commit_T)
        sess.fold(ex["r"])

    def commit_w(sess, ex):
        nonlocal h2d
        if upload:
            h2d += sess.upload_incoming(ex["host_W"])
        if combiner is None:
            sess.commit_incoming()
        else:
            combiner.commit(sess, sess.W_in.data_ptr(), sess.A * sess.n, 1, sess.h_commit_W)

    prove(sec, es)       # 1. fold the secondary accumulator
    commit_w(prim, ep)   # 2. primary trace: commit W
    prove(prim, ep)      # 3. fold the primary accumulator
    commit_w(sec, es)    # 4. secondary trace: commit W
    return h2d


def run_gpu(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    import sirius_b200
    from sirius_b200 import _lib

    lib = sirius_b200.load()
    _lib.check(lib.sb_init(local_rank))
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep rank 0's stdout to the one JSON line (NCCL prints its banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    prim, ep = build_side_gpu(PRIMARY, rank, world, stream)
    sec, es = build_side_gpu(SECONDARY, rank, world, stream)
    combiner = Combiner(world, stream) if world > 1 else None
    sides, extras = (prim, sec), (ep, es)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(upload):
        for _ in range(args.warmup):
            gpu_step(sides, extras, upload, combiner)
        barrier()
        import ctypes

        NT = 10
        ms_arr, un_arr, ln_arr = (ctypes.c_double * NT)(), (ctypes.c_uint64 * NT)(), (ctypes.c_uint64 * NT)()
        lib.sb_profile_enable(1)
        lib.sb_profile_collect(ms_arr, un_arr, ln_arr)
        launches0 = lib.sb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h2d = 0
        for _ in range(args.steps):
            h2d += gpu_step(sides, extras, upload, combiner)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        _lib.check(lib.sb_profile_collect(ms_arr, un_arr, ln_arr))
        lib.sb_profile_enable(0)
        tags = ["decompose", "sort", "accumulate", "fixup", "reduce", "finalize", "cross_terms", "fold", "ntt", "protogalaxy"]
        breakdown = {t: round(ms_arr[i] / args.steps, 4) for i, t in enumerate(tags) if ln_arr[i]}
        launches = lib.sb_launch_count() - launches0
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / args.steps, h2d // max(1, args.steps), launches // max(1, args.steps), (ms_arr[2], un_arr[2], ln_arr[2]), breakdown

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, _, launches, prof, breakdown = timed(upload=False)
    ms_e2e, h2d_bytes, _, _, _ = timed(upload=True)
    clocks = sampler.stop() if rank == 0 else None

    extra = {}
    if rank == 0 and world == 1:
        extra = gpu_side_metrics(stream)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        acc_ms, acc_madds, acc_n = prof
        points_per_step = (12 + 6 + 7 + 5) * (1 << K_TABLE)
        acc_pts = points_per_step * args.steps / world   # points this rank pushed through k_accumulate in the timed region
        achieved = (96.0 * acc_pts / 1e9) / (acc_ms / 1e3) if acc_ms else 0.0
        line = {
            "metric": METRIC, "value": round(ms_dev, 4), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_dev, 4), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": round(ms_e2e, 4), "unit": "ms", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 13 * 64},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {
                "kernel": "sb::k_accumulate (MSM bucket accumulation)", "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak,
                "unit": "GB/s", "frac": round(achieved / hbm_peak, 5),
                "traffic": 3.123e9 if world == 1 else None,
                "traffic_note": "dram__bytes_read+write of the largest launch (primary W commit, 1 572 864 points, 151 MB algorithmic) from "
                                "profiles/r1_accumulate_s2_ncu_full_summary.txt: each point is gathered once per window (15 x 64 B table entries, a "
                                "random 64 B gather costs a 128 B DRAM fetch) by design; the kernel sits at 10 % of DRAM throughput and is integer-pipe bound",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "launches": int(acc_n), "avg_launch_ms": round(acc_ms / acc_n, 4) if acc_n else None,
                "note": "algorithmic 96 B/point; the kernel is integer-pipe bound (SURVEY F7): see int_pipe and DESIGN.md 4.2",
                "int_pipe": {
                    "achieved_gmadd_per_s": round(acc_madds / (acc_ms * 1e6), 3) if acc_ms else None,
                    "peak_gmadd_per_s": 6.386,
                    "frac": round(acc_madds / (acc_ms * 1e6) / 6.386, 4) if acc_ms else None,
                    "peak_source": "profiles/r1_microbench6_madd_lazy.txt: serial lazy-domain XYZZ mixed additions on the full chip (IMAD.WIDE issue bound, 65.4 G Montgomery products/s)",
                },
            },
            "breakdown_ms_per_step": breakdown,
            "msm_points_per_step": points_per_step,
            "msm_mscalar_per_s_in_step": round(points_per_step / ms_dev / 1e3, 2),
            "extra": extra,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_step_baseline(1)
            except Exception as exc:  # the oracle is test infrastructure; never let it take the GPU number down
                line["cpu_baseline"] = {"error": str(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def gpu_side_metrics(stream):
    """MSM Mscalar/s (2^20, bn256) and NTT Gelt/s (k = 17, 20) -- BASELINE.json's secondary metrics."""
    import torch

    from sirius_b200 import device, fft

    out = {}
    n = 1 << 20
    ck = device.synthetic_key(0, n, stream=stream)
    s = device.random_field_device(n, SEED + 77)
    o = torch.zeros(8, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ck.commit_device(s.data_ptr(), n, o.data_ptr(), 0, stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        ck.commit_device(s.data_ptr(), n, o.data_ptr(), 0, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["msm_2^20_bn256"] = {"ms": round(ms, 3), "mscalar_per_s": round(n / ms / 1e3, 1), "gbps_algorithmic": round(96 * n / ms / 1e6, 1)}
    ck.close()
    for k in (17, 20):
        a = device.random_field_device(1 << k, SEED + k)
        w = fft.get_omega_or_inv(k, False)
        for _ in range(3):
            fft.ntt_device(a.data_ptr(), k, w, None, stream.cuda_stream)
        e0.record(stream)
        for _ in range(20):
            fft.ntt_device(a.data_ptr(), k, w, None, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[f"ntt_k{k}"] = {"us": round(ms * 1e3, 1), "gelt_per_s": round((1 << k) / ms / 1e6, 3), "gbps_algorithmic": round(64 * (1 << k) / ms / 1e6, 1)}
    return out


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
_CPU_STATE = {}


def cpu_threads() -> int:
    """Host threads for the CPU arm: the physical cores (the restatement, like rayon-based halo2, gets slower when
    it is spread over SMT siblings: 9.2 s on 128 logical vs 4.7 s on 64 physical cores of the round-1 box)."""
    try:
        import psutil

        n = psutil.cpu_count(logical=False)
        if n:
            return int(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_setup():
    """The same shapes for the CPU restatement (oracle/): test infrastructure, used only as the timed baseline."""
    if _CPU_STATE:
        return _CPU_STATE
    import numpy as np

    import oracle
    from oracle import expr_ref as E
    from oracle import pyref as R

    oracle.build()
    n = 1 << K_TABLE
    st = {}
    for side in (PRIMARY, SECONDARY):
        nfix, nadv = shapes(side)
        gates, fb, ab = [], 0, 0
        for T in side["T_list"]:
            gates.append(E.main_gate_expression(T, fb, ab, 0, nfix))
            fb += 2 * T + 5
            ab += T + 2
        cg = E.CompressedGates(gates, E.Ctx(num_fixed=nfix, num_advice=nadv))
        f = side["field"]
        evs = [None if ex is None else E.GraphEvaluator(ex, R.MODULUS[f]) for ex in cg.grouped()[1:]]
        nch = cg.ctx.num_challenges - 1
        st[side["name"]] = dict(
            side=side, nadv=nadv, evs=evs,
            fixed=[oracle.random_field(f, SEED + 31 * j + side["curve"], n) for j in range(nfix)],
            W1=oracle.random_field(f, SEED + 1, nadv * n), W2=oracle.random_field(f, SEED + 2, nadv * n),
            E1=oracle.random_field(f, SEED + 3, n),
            ch=np.concatenate([oracle.random_field(f, SEED + 4, nch + 1), oracle.random_field(f, SEED + 5, nch), R.to_mont_limbs([1], R.MODULUS[f])]),
            r=oracle.random_field(f, SEED + 6, 1).reshape(4),
            bases=oracle.running_bases(side["curve"], nadv * n),
        )
    _CPU_STATE.update(st)
    return _CPU_STATE


def cpu_step():
    """Literal reference order on the host cores: d evaluator sweeps + d commits + folds, W commits."""
    import ctypes

    import numpy as np

    import oracle
    from oracle import expr_ref as E

    st = cpu_setup()
    n = 1 << K_TABLE
    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib = oracle.lib()

    def prove(s):
        f, curve, nadv = s["side"]["field"], s["side"]["curve"], s["nadv"]
        adv = [s["W1"][i * n:(i + 1) * n] for i in range(nadv)] + [s["W2"][i * n:(i + 1) * n] for i in range(nadv)]
        T = [np.zeros((n, 4), dtype=np.uint64) if ev is None else E.c_graph_evaluate(f, ev, [], s["fixed"], adv, s["ch"], K_TABLE, threads=cpu_threads()) for ev in s["evs"]]
        for t in T:
            oracle.msm(curve, t, s["bases"], threads=cpu_threads())
        outw = np.zeros_like(s["W1"])
        lib.so_axpy(f, s["W1"].ctypes.data_as(u64p), s["W2"].ctypes.data_as(u64p), s["r"].ctypes.data_as(u64p), outw.ctypes.data_as(u64p), ctypes.c_size_t(nadv * n))
        ptrs = (u64p * len(T))(*[t.ctypes.data_as(u64p) for t in T])
        oute = np.zeros_like(s["E1"])
        lib.so_error_fold(f, s["E1"].ctypes.data_as(u64p), ptrs, ctypes.c_size_t(len(T)), s["r"].ctypes.data_as(u64p), oute.ctypes.data_as(u64p), ctypes.c_size_t(n))

    def commit_w(s):
        oracle.msm(s["side"]["curve"], s["W2"], s["bases"], threads=cpu_threads())

    prove(st["secondary"])
    commit_w(st["primary"])
    prove(st["primary"])
    commit_w(st["secondary"])


def cpu_step_baseline(steps):
    import oracle

    cpu_setup()
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    return {"value": round(ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port",
            "sample": f"{steps} full fold_step hot path(s) of the same workload (CPU restatement of the reference: halo2-style chunked Pippenger, "
                      "literal GroupedPoly/GraphEvaluator cross terms, OpenMP on all host cores)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import oracle

    cpu_setup()
    for _ in range(min(args.warmup, 1)):
        cpu_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    base = {"value": round(ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port",
            "sample": "full fold_step hot path per step (CPU restatement of the reference; the Rust crate cannot be built here: no cargo/rustc)"}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(ms, 2), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": round(ms, 2), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 limbs (254-bit modular integers)", "data": "synthetic", "config": workload_config(world),
        "cpu_baseline": base, "e2e": {"value": round(ms, 2), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    global K_TABLE, CK_LOG, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--k", type=int, default=17, help="table size 2^k rows (default 17 = BASELINE configs[1]; 20 = the north star's larger case, "
                                                       "same shapes; needs ~45 GB of window tables)")
    args = ap.parse_args()
    if args.k != K_TABLE:   # same circuit shapes at another table size (the key covers 16 * 2^k generators, as at k = 17)
        K_TABLE, CK_LOG = args.k, args.k + 4
        METRIC = f"sangria_poseidon k={args.k} IVC fold_step prover hot-path time"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
