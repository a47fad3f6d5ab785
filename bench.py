#!/usr/bin/env python
"""bench.py -- the Sangria IVC fold step's prover hot path at the shapes of benches/sangria_poseidon.

A "step" is one pass of the hot path of `IVC::fold_step` (reference
src/ivc/sangria/incrementally_verifiable_computation.rs:428-635) over one batch of synthetic witness
columns, in the reference's call order (sirius_b200/workload.py):

  1. VanillaFS::prove, secondary side (grumpkin; A=7, F=15, 1 gate, d=5): 5 cross-term vectors + 5 commits of 2^k, W/E fold
  2. generate_plonk_trace, primary side (bn256): commit W, 12*2^k scalars
  3. VanillaFS::prove, primary side (bn256; A=12, F=26, 2 gates, d=6): 6 cross terms + 6 commits + folds
  4. generate_plonk_trace, secondary side: commit W, 7*2^k scalars

halo2 circuit synthesis (CPU, out of scope per SURVEY section 2) is not part of the step.  Every commitment is copied
back to the host and synchronised before the next stage, as the random oracle requires.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload sangria_poseidon|cyclefold_poseidon|msm_sweep]

`value`        : ms per step at k = 17 (BASELINE configs[1]), all inputs resident in HBM.
`e2e`          : ms per step with that step's fresh witness columns copied host->device (pinned) inside the timed
                 region and the 13 commitments read back: the device-resident prover state of SURVEY 8f-1.
`e2e_host_abi` : the same step through the host-buffer entry points a verbatim Rust shim binds (CommitmentKey.commit,
                 VanillaFS.commit_cross_terms, RelaxedPlonkWitness.fold on PAGEABLE host arrays; T, folded W and E copied
                 back), N = 1 only.
`verified`     : before timing, one step of the exact timed objects is compared bit for bit (13 commitments, folded W
                 and E on both sides) with the CPU restatement of the reference on identical inputs.
`extra.k20`    : the same measurement at k = 20, the table size the north star quotes its target on.
`stages`       : the rest of the measured bench iteration (benches/sangria_poseidon.rs:159-175): `new` and `verify` legs.
N > 1          : rows are sharded across ranks (strong scaling); each commitment group is one exchange of 128-byte
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
SEED = 0x5349524955530000


def metric_name(k):
    return f"sangria_poseidon k={k} IVC fold_step prover hot-path time"


def workload_config(n_gpus, k):
    return {
        "workload": f"benches/sangria_poseidon k={k} bn256/grumpkin: fold_step hot path "
                    f"(MSM {12 << k} + 6x{1 << k} bn256, MSM {7 << k} + 5x{1 << k} grumpkin, 11 cross-term vectors, W/E folds)",
        "k": k,
        "ck_log2": k + 4,
        "sharding": "single GPU" if n_gpus == 1 else f"rows sharded over {n_gpus} ranks; per commitment group ONE kernel exchanges the 128-byte partial sums over NVLink peer "
                                                      f"memory and adds them ({os.environ.get('SB_BENCH_EXCHANGE', 'peer')} exchange)",
        "l2": "inputs larger than L2 (window tables of several GB gathered at random; 0.4 GB of per-step scratch at k=17)",
        "phases": ("sequential: one host synchronisation per commitment group (reference call order)" if os.environ.get("SB_BENCH_OVERLAP", "1") == "0"
                   or os.environ.get("SB_BENCH_EXCHANGE", "peer") == "nccl" and n_gpus > 1 else
                   "two per step (secondary trace, primary trace): the trace's W commitment on a second stream beside its cross terms + T commitments, "
                   "one host synchronisation per phase (sirius_b200/workload.py step())"),
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _nvml(self):
        """In-process NVML handle (the library behind nvidia-smi): a query costs microseconds.  Spawning `nvidia-smi` ten times a
        second next to phases of a millisecond was visible in the timed region (each start-up re-initialises the driver's
        management interface); it stays as the fallback when the nvidia-ml-py binding is missing."""
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    pass
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
        except Exception:
            return None, None

    def _run(self):
        nv, h = self._nvml()
        self.source = "nvml (nvidia-ml-py, in process)" if nv else "nvidia-smi"
        while not self.stop_flag:
            try:
                if nv:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    flag = lambda bit: "Active" if r & bit else "Not Active"  # noqa: E731
                    self.samples.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)),
                                         flag(nv.nvmlClocksEventReasonHwSlowdown), flag(nv.nvmlClocksEventReasonHwThermalSlowdown),
                                         flag(nv.nvmlClocksEventReasonSwThermalSlowdown), flag(nv.nvmlClocksEventReasonSwPowerCap)])
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05 if nv else 0.25)

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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "source": getattr(self, "source", "nvidia-smi")}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
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


def cpu_fold_step(inputs, bases):
    """(results, ms): the CPU restatement of the reference on `inputs` (oracle/step_ref.py), all host cores."""
    import oracle
    from oracle import step_ref

    oracle.build()
    t0 = time.perf_counter()
    res = step_ref.fold_step(inputs, bases, threads=cpu_threads())
    return res, (time.perf_counter() - t0) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import oracle
    from oracle import step_ref
    from sirius_b200 import workload as WL

    oracle.build()
    k = args.k
    inputs = step_ref.synthetic_inputs((WL.PRIMARY, WL.SECONDARY), k, SEED)
    bases = step_ref.bases_for(inputs)
    for _ in range(min(args.warmup, 1)):
        cpu_fold_step(inputs, bases)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_fold_step(inputs, bases)
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    base = {"value": round(ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port",
            "sample": "full fold_step hot path per step (CPU restatement of the reference; the Rust crate cannot be built here: no cargo/rustc)"}
    line = {
        "impl": "reference", "metric": metric_name(k), "value": round(ms, 2), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": round(ms, 2), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 limbs (254-bit modular integers)", "data": "synthetic", "config": workload_config(world, k),
        "cpu_baseline": base, "e2e": {"value": round(ms, 2), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
TAGS = ["decompose", "sort", "accumulate", "fixup", "reduce", "finalize", "cross_terms", "fold", "ntt", "protogalaxy"]


class GpuArm:
    def __init__(self, args):
        import torch

        self.args, self.torch = args, torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        import sirius_b200
        from sirius_b200 import _lib

        self._lib = _lib
        self.lib = sirius_b200.load()
        _lib.check(self.lib.sb_init(self.local_rank))
        if self.world > 1:
            import torch.distributed as dist

            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"   # keep rank 0's stdout to the one JSON line (NCCL prints its banner there)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.stream = torch.cuda.Stream()

    def barrier(self):
        torch = self.torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        import torch.distributed as dist

        t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # -------------------------------------------------------------------------- verification against the oracle
    def verify(self, wl):
        """One step of the timed objects vs the CPU restatement on identical inputs (rank 0 runs the oracle)."""
        wl.step(upload=False)    # two unchecked steps first: the library captures each commitment pipeline into a CUDA graph on its
        wl.step(upload=True)     # second call and replays it afterwards -- the step that is checked runs on the replayed graphs
        snap = wl.snapshot_inputs()
        wl.step(upload=True)     # the e2e form: the uploaded pinned copy equals W_in, so both timed variants are covered
        res = wl.snapshot_results()
        report, cpu_ms = None, None
        if self.rank == 0:
            from oracle import step_ref

            bases = step_ref.bases_for(snap)
            cpu_res, cpu_ms = cpu_fold_step(snap, bases)
            report = step_ref.compare(res, cpu_res)
            report["what"] = ("13 commitments + folded W and E of both sides, bit for bit, the third step of the timed objects (captured commitment pipelines "
                              "replaying, as in the timed region) vs oracle/step_ref.py on identical inputs")
        self.barrier()
        return report, cpu_ms

    # -------------------------------------------------------------------------- timing
    def timed(self, wl, upload, steps, warmup, profile=True):
        """profile = False: no per-kernel-group events in the timed region (two cudaEventRecord per group and commit: ~60 driver
        calls per step, which sit on the critical path once the phases are short -- the multi-GPU shards)."""
        import ctypes

        torch, lib = self.torch, self.lib
        for _ in range(warmup):
            wl.step(upload)
        self.barrier()
        NT = len(TAGS)
        ms_arr, un_arr, ln_arr = (ctypes.c_double * NT)(), (ctypes.c_uint64 * NT)(), (ctypes.c_uint64 * NT)()
        if profile:
            lib.sb_profile_enable(1)
            lib.sb_profile_collect(ms_arr, un_arr, ln_arr)
        launches0 = lib.sb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        h2d = 0
        for _ in range(steps):
            h2d += wl.step(upload)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if profile:
            self._lib.check(lib.sb_profile_collect(ms_arr, un_arr, ln_arr))
            lib.sb_profile_enable(0)
        breakdown = {t: round(ms_arr[i] / steps, 4) for i, t in enumerate(TAGS) if ln_arr[i]}
        launches = lib.sb_launch_count() - launches0
        ms = self.max_over_ranks(ms)
        return dict(ms=ms / steps, steps=steps, h2d=h2d // max(1, steps), launches=launches // max(1, steps), acc=(ms_arr[2], un_arr[2], ln_arr[2]), breakdown=breakdown)

    def timed_leg(self, fn, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps

    # -------------------------------------------------------------------------- the host-buffer ABI path (N = 1)
    def host_abi(self, wl, steps, warmup):
        """The step through the entry points a verbatim Rust shim binds: pageable host arrays in, T / folded W / E back."""
        import numpy as np

        from sirius_b200 import sangria as SG

        torch = self.torch
        sides = []
        for sess, ex in zip(wl.sides, wl.extras):
            torch.cuda.synchronize()
            sides.append(dict(sess=sess, ex=ex, acc=SG.RelaxedPlonkWitness(sess.S.field, [sess.W_acc.cpu().numpy().view(np.uint64).copy()], sess.E_acc.cpu().numpy().view(np.uint64).copy()),
                              W2=sess.W_in.cpu().numpy().view(np.uint64).copy()))
        prim, sec = sides
        h2d = d2h = 0

        def prove(s):
            nonlocal h2d, d2h
            sess, ex = s["sess"], s["ex"]
            T, commits = SG.VanillaFS.commit_cross_terms(sess.ck, sess.S, ex["c1"], ex["u1"], s["acc"].W, ex["c2"], [s["W2"]])
            s["acc"] = s["acc"].fold([s["W2"]], T, ex["r"])
            wb, tb, eb = s["W2"].nbytes, sum(t.nbytes for t in T), s["acc"].E.nbytes
            h2d += 2 * wb + tb + 2 * wb + eb + tb    # commit_cross_terms: W1, W2 ; commit(T_j) ; fold: W1, W2, E, T
            d2h += tb + wb + eb + 64 * len(T)

        def commit_w(s):
            nonlocal h2d, d2h
            s["sess"].ck.commit(s["W2"])
            h2d += s["W2"].nbytes
            d2h += 64

        def step():
            prove(sec)
            commit_w(prim)
            prove(prim)
            commit_w(sec)

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        h2d = d2h = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return {"value": round(ms, 3), "unit": "ms", "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps, "steps": steps,
                "timer": "host wall clock: every call blocks until its result is in host memory",
                "path": "CommitmentKey.commit / VanillaFS.commit_cross_terms / RelaxedPlonkWitness.fold on pageable host arrays (sb_msm, sb_cross_terms, sb_msm_batch, sb_axpy_fold, sb_error_fold)"}

    # -------------------------------------------------------------------------- one table size
    def measure(self, k, steps, warmup, with_cpu, with_host_abi, with_stages):
        from sirius_b200 import workload as WL

        wl = WL.SangriaStepWorkload(k, self.rank, self.world, self.stream)
        out = {"k": k, "steps": steps, "warmup": warmup}
        cpu_ms = None
        if not self.args.no_verify:
            report, cpu_ms = self.verify(wl)
            if self.rank == 0:
                out["verified"] = bool(report["ok"])
                out["verify"] = report
                if not report["ok"]:
                    raise SystemExit(f"bench.py: the timed path disagrees with the oracle: {report['bad']}")
        two_streams = bool(getattr(wl, "overlap", False))
        dev = self.timed(wl, False, steps, warmup, profile=not two_streams)
        e2e = self.timed(wl, True, steps, warmup, profile=not two_streams)
        # per-kernel durations (roofline, breakdown) come from a leg with the two streams serialised: in the timed region
        # above the W commitment shares the SMs with the cross terms / T commitments of the other stream, so a kernel's
        # event-to-event time there is not the kernel's own
        seq = dev
        if getattr(wl, "overlap", False):
            wl.overlap = False
            seq = self.timed(wl, False, max(3, steps // 2), 2)
            wl.overlap = True
        out.update(dev=dev, e2e=e2e, seq=seq)
        if with_stages:
            out["stages"] = {
                "new_ms": round(self.timed_leg(wl.new_leg, max(2, steps // 2), 1), 4),
                "verify_ms": round(self.timed_leg(wl.verify_leg, max(2, steps // 2), 1), 4),
                "note": "benches/sangria_poseidon.rs:159-175 times IVC::fold(.., 1) = new + fold_step + verify: new_ms = the two W commits of IVC::new; "
                        "verify_ms = is_sat_accumulation row sweep + re-commit of W_acc, E, W_in + PlonkStructure::is_sat row sweep, both sides (device-resident)",
            }
        if with_host_abi and self.world == 1:
            out["e2e_host_abi"] = self.host_abi(wl, max(2, min(steps, 3)), 1)
        if with_cpu and cpu_ms is not None and self.rank == 0:
            out["cpu_baseline"] = {"value": round(cpu_ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port",
                                   "sample": "1 full fold_step hot path of the same workload ON THE SAME INPUTS as the verified GPU step (CPU restatement of the reference: "
                                             "halo2-style chunked Pippenger, literal GroupedPoly/GraphEvaluator cross terms, OpenMP on all host cores)"}
        wl.close()
        del wl
        import gc

        gc.collect()
        self.torch.cuda.synchronize()
        self.torch.cuda.empty_cache()
        return out


def roofline_block(m, world, peaks):
    k, steps = m["k"], m["seq"]["steps"]
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    acc_ms, acc_madds, acc_n = m["seq"]["acc"]
    points_per_step = (12 + 6 + 7 + 5) * (1 << k)
    acc_pts = points_per_step * steps / world   # points this rank pushed through k_accumulate in the timed region
    achieved = (96.0 * acc_pts / 1e9) / (acc_ms / 1e3) if acc_ms else 0.0
    traffic, traffic_note = None, "no ncu capture of this kernel at this shape"
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "accumulate_traffic.json")))
        if world == 1 and k == t.get("k"):
            traffic, traffic_note = t["dram_bytes_per_launch"], t["note"]
    except Exception:
        pass
    # instruction-level bound of the integer pipe: one IMAD.WIDE warp-instruction per 4 cycles per scheduler (quarter
    # rate, profiles/r1_microbench5_pipes.txt), >= 128 of them per 254-bit Montgomery product (8x8 limb products for
    # a*b and 8x8 for the reduction), 592 schedulers at the sampled clock
    sm_clock = 1.965e9
    prod_peak = 148 * 4 * sm_clock * 32 / (128 * 4.0) / 1e9          # G products/s
    gmadd = acc_madds / (acc_ms * 1e6) if acc_ms else None
    return {
        "kernel": "sb::k_accumulate (MSM bucket accumulation)", "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak,
        "unit": "GB/s", "frac": round(achieved / hbm_peak, 5), "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
        "launches": int(acc_n), "avg_launch_ms": round(acc_ms / acc_n, 4) if acc_n else None,
        "timing": "CUDA events around the kernel on its own stream, inside bench.py, from the leg that runs the step's phases sequentially (sequential_phases_ms): "
                  "in the two-stream timed region the kernel shares the SMs with the other stream's cross terms / commitments and an event pair around it would "
                  "not measure the kernel alone (nor are the ~60 extra event records per step wanted on the critical path of the short multi-GPU phases)",
        "note": "algorithmic 96 B/point; the kernel is integer-pipe bound (SURVEY F7): see int_pipe and DESIGN.md 4.2",
        "int_pipe": {
            "achieved_gmadd_per_s": round(gmadd, 3) if gmadd else None,
            "issue_bound_gproducts_per_s": round(prod_peak, 2),
            "frac_of_issue_bound_xyzz": round(gmadd * 10 / prod_peak, 4) if gmadd else None,
            "frac_of_issue_bound_affine": round(gmadd * 6 / prod_peak, 4) if gmadd else None,
            "peak_source": "IMAD.WIDE issue rate (1 warp-instruction / 4 cycles / scheduler) x 128 IMAD.WIDE per Montgomery product x 592 schedulers x 1.965 GHz; "
                           "xyzz = 10 products per mixed addition (what the kernel executes), affine = 6 products (batched-affine addition, the algorithmic minimum): "
                           "the second fraction is the slack left to the ALGORITHM, the first to the loop",
        },
    }


def run_gpu(args):
    arm = GpuArm(args)
    rank, world = arm.rank, arm.world
    sampler = ClockSampler(arm.local_rank)
    if rank == 0:
        sampler.start()
    main = arm.measure(args.k, args.steps, args.warmup, with_cpu=not args.no_cpu_baseline and world == 1, with_host_abi=not args.no_host_abi, with_stages=True)
    clocks = sampler.stop() if rank == 0 else None
    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = gpu_side_metrics(arm.stream)
    if not args.no_k20 and args.k != 20:
        sampler20 = ClockSampler(arm.local_rank)
        if rank == 0:
            sampler20.start()
        m20 = arm.measure(20, max(2, min(args.steps, 3)), 2, with_cpu=not args.no_cpu_baseline and world == 1, with_host_abi=False, with_stages=False)
        c20 = sampler20.stop() if rank == 0 else None
        if rank == 0:
            peaks = _peaks()
            extra["k20"] = {
                "metric": metric_name(20), "value": round(m20["dev"]["ms"], 3), "unit": "ms", "n_gpus": world, "steps": m20["steps"], "warmup": m20["warmup"],
                "e2e": {"value": round(m20["e2e"]["ms"], 3), "unit": "ms", "h2d_bytes_per_step": int(m20["e2e"]["h2d"]), "d2h_bytes_per_step": 13 * 64},
                "verified": m20.get("verified"), "verify_bad": (m20.get("verify") or {}).get("bad"),
                "breakdown_ms_per_step": m20["seq"]["breakdown"], "sequential_phases_ms": round(m20["seq"]["ms"], 3), "gpu_launches": int(m20["dev"]["launches"]), "clocks": c20,
                "cpu_baseline": m20.get("cpu_baseline"), "config": workload_config(world, 20), "roofline": roofline_block(m20, world, peaks),
                "note": "the north star's target configuration (sangria_poseidon at k=20): same circuit shapes, 2^20 rows",
            }
    if rank == 0:
        peaks = _peaks()
        k = args.k
        points_per_step = (12 + 6 + 7 + 5) * (1 << k)
        line = {
            "metric": metric_name(k), "value": round(main["dev"]["ms"], 4), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(main["dev"]["ms"], 4), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic", "config": workload_config(world, k),
            "e2e": {"value": round(main["e2e"]["ms"], 4), "unit": "ms", "h2d_bytes_per_step": int(main["e2e"]["h2d"]), "d2h_bytes_per_step": 13 * 64,
                    "path": "device-resident prover state (SURVEY 8f-1): fresh witness columns streamed from pinned host memory in row blocks (the cross-term sweep of a block "
                            "runs under the transfer of the next), 13 commitments read back"},
            "gpu_launches": int(main["dev"]["launches"]), "clocks": clocks,
            "verified": main.get("verified"), "verify": main.get("verify"),
            "roofline": roofline_block(main, world, peaks),
            "breakdown_ms_per_step": main["seq"]["breakdown"],
            "breakdown_note": "kernel groups timed with the phases run sequentially (sequential_phases_ms per step); in the timed region the two streams overlap "
                              "and the groups do not add up to the step",
            "sequential_phases_ms": round(main["seq"]["ms"], 4),
            "stages": main.get("stages"),
            "msm_points_per_step": points_per_step,
            "msm_mscalar_per_s_in_step": round(points_per_step / main["dev"]["ms"] / 1e3, 2),
            "extra": extra,
        }
        if "e2e_host_abi" in main:
            line["e2e_host_abi"] = main["e2e_host_abi"]
        if "cpu_baseline" in main:
            line["cpu_baseline"] = main["cpu_baseline"]
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def gpu_side_metrics(stream):
    """MSM Mscalar/s (2^20, bn256) and NTT Gelt/s (k = 17, 20) -- BASELINE.json's secondary metrics."""
    import torch

    from sirius_b200 import device, fft

    out = {}
    n = 1 << 20
    ck = device.synthetic_key(0, n, stream=stream)
    with torch.cuda.stream(stream):
        s = device.random_field_device(n, SEED + 77)
        o = torch.zeros(8, dtype=torch.int64, device="cuda")
    stream.synchronize()
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
        with torch.cuda.stream(stream):
            a = device.random_field_device(1 << k, SEED + k)
        stream.synchronize()
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


def run_cyclefold(args):
    """--workload cyclefold_poseidon: the prover hot path of cyclefold::IVC::next (benches/cyclefold_poseidon.rs:116-125;
    sirius_b200.workload.CyclefoldStepWorkload) at k = args.k, one GPU, ROW_CORRECT leaves timed, verified against the oracle
    (and the reference-compatible ROW_COMPAT mode parity-checked once)."""
    k = args.k
    metric = f"cyclefold_poseidon k={k} IVC next() prover hot-path time"
    config = {"workload": f"benches/cyclefold_poseidon k={k}: ProtoGalaxy::prove (F over 2^{k + 1} leaves x 32 points, G over 8 Lagrange blends, K, fold_witness of {12 << k} cells) "
                          f"+ support-circuit fold (grumpkin k=15) + MSM {12 << k} bn256", "k": k, "row_mode": "correct (index % 2^k); the reference's `index & 2^k` mode is parity-checked, not timed (SURVEY F4)",
              "sharding": "single GPU (the beta tree does not shard; replicas only)"}
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.impl == "reference":
        import oracle
        from oracle import step_ref

        oracle.build()
        inp = _cyclefold_synthetic_inputs(k)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_ref.cyclefold_step(inp, threads=cpu_threads())
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        base = {"value": round(ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port", "sample": "full next() hot path per step (CPU restatement of the reference)"}
        print(json.dumps({"impl": "reference", "metric": metric, "value": round(ms, 2), "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": 0,
                          "ms_per_step": round(ms, 2), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64 limbs (254-bit modular integers)",
                          "data": "synthetic", "config": config, "cpu_baseline": base,
                          "e2e": {"value": round(ms, 2), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    arm = GpuArm(args)
    torch = arm.torch
    from sirius_b200 import workload as WL

    sampler = ClockSampler(arm.local_rank)
    sampler.start()
    out = {}
    verified, cpu_ms, reports = None, None, {}
    if not args.no_verify:
        from oracle import step_ref

        for mode, name in ((0, "compat"), (1, "correct")):
            if mode == 0 and k > 17:
                continue   # the reference-compatible mode is checked at k <= 17 (same kernels, row 0 everywhere)
            wl = WL.CyclefoldStepWorkload(k, arm.stream, row_mode=mode)
            snap = wl.snapshot_inputs()
            wl.step(upload=True)
            res = wl.snapshot_results()
            t0 = time.perf_counter()
            exp = step_ref.cyclefold_step(snap, threads=cpu_threads())
            if mode == 1:
                cpu_ms = (time.perf_counter() - t0) * 1e3
            reports[name] = step_ref.compare_cyclefold(res, exp)
            wl.close()
            del wl
            torch.cuda.empty_cache()
        verified = all(r["ok"] for r in reports.values())
        if not verified:
            raise SystemExit(f"bench.py: cyclefold path disagrees with the oracle: { {n: r['bad'] for n, r in reports.items()} }")
    wl = WL.CyclefoldStepWorkload(k, arm.stream, row_mode=1)
    dev = arm.timed(wl, False, args.steps, args.warmup)
    e2e = arm.timed(wl, True, args.steps, args.warmup)
    clocks = sampler.stop()
    line = {
        "metric": metric, "value": round(dev["ms"], 4), "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev["ms"], 4),
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic", "config": config,
        "e2e": {"value": round(e2e["ms"], 4), "unit": "ms", "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": (32 + 8) * 32 + 4 * 64},
        "gpu_launches": int(dev["launches"]), "clocks": clocks, "verified": verified, "verify": reports,
        "breakdown_ms_per_step": dev["breakdown"],
        "breakdown_note": "protogalaxy = leaf evaluation (k_expr_eval) + beta trees + lincomb; the remainder of the step is host glue (Python integers here, Rust in the integration: "
                          "ifft of 32 / 8 points, K on 256 points) and the per-stage synchronisations the random oracle imposes",
    }
    if cpu_ms is not None:
        line["cpu_baseline"] = {"value": round(cpu_ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port",
                                "sample": "1 full next() hot path on the same inputs as the verified GPU step (CPU restatement: C GraphEvaluator interpreter + array beta tree + halo2-style Pippenger, OpenMP)"}
    print(json.dumps(line), flush=True)
    wl.close()


def _cyclefold_synthetic_inputs(k):
    import random

    import oracle
    from oracle import pyref as R
    from sirius_b200 import workload as WL

    nfix, nadv = WL.shapes(WL.PRIMARY)
    n = 1 << k
    rng = random.Random(SEED)
    M = R.FR
    sfix, sadv = WL.shapes(WL.SUPPORT)
    ns = 1 << 15
    import numpy as np

    sup = dict(side=WL.SUPPORT, k=15, nadv=sadv, nfix=sfix, fixed=[oracle.random_field(1, SEED + 50 + j, ns) for j in range(sfix)],
               selectors=[(np.arange(ns) % 2).astype(np.uint8)], W1=oracle.random_field(1, SEED + 60, sadv * ns), E1=oracle.random_field(1, SEED + 61, ns),
               W2=oracle.random_field(1, SEED + 62, sadv * ns), c1=np.zeros((0, 4), dtype=np.uint64), c2=np.zeros((0, 4), dtype=np.uint64),
               u1=oracle.random_field(1, SEED + 63, 1).reshape(4), r=oracle.random_field(1, SEED + 64, 1).reshape(4))
    return dict(k=k, row_mode=1, fixed=[oracle.random_field(0, SEED + 31 * j, n) for j in range(nfix)], nadv=nadv,
                W_acc=oracle.random_field(0, SEED + 1, nadv * n), W_in=oracle.random_field(0, SEED + 2, nadv * n),
                betas=[rng.randrange(M) for _ in range(k + 1)], delta=rng.randrange(M), alpha=rng.randrange(M), gamma=rng.randrange(M), support=sup)


def run_gate_scaling(args):
    """--workload gate_scaling: BASELINE config 5 (benches/ivc_gate_scaling.rs).  The reference scales the NUMBER of parallel
    Poseidon sub-circuits (5 / 10 / 20 for sangria::IVC at k = 17, :221-223; 6 / 12 / 24 / 48 for cyclefold::IVC at k = 20,
    :229-232); there is no degree knob.  Sangria arm: the same fold_step hot path as the main line with the primary
    circuit widened to N sub-circuits (A = 7 + 5N, F = 15 + 11N, N + 1 gates, folding degree 5 + N: N + 5 cross terms and
    commitments).  Cyclefold arm: ProtoGalaxy::prove (F, G, K, fold_witness) on the same shapes in the correct row mode.
    The smallest cell of each arm is checked bit for bit against the CPU restatement before timing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k_s, k_pg = args.k, args.pg_k
    sg_counts = [int(x) for x in args.gates.split(",") if x]
    pg_counts = [int(x) for x in args.pg_gates.split(",") if x]
    metric = "ivc_gate_scaling prover hot-path time vs gate count"
    config = {"workload": f"benches/ivc_gate_scaling: sangria fold_step at k={k_s} for N={sg_counts} sub-circuits; ProtoGalaxy::prove at k={k_pg} for N={pg_counts}",
              "k": k_s, "pg_k": k_pg, "sharding": "single GPU"}
    if args.impl == "reference":
        import oracle
        from oracle import step_ref
        from sirius_b200 import workload as WL

        oracle.build()
        N = sg_counts[0]
        inp = step_ref.synthetic_inputs([WL.gate_scaling_side(N), WL.SECONDARY], k_s, SEED)
        bases = step_ref.bases_for(inp)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_ref.fold_step(inp, bases, threads=cpu_threads())
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        base = {"value": round(ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port", "sample": f"full fold_step hot path at N={N} sub-circuits per step (CPU restatement)"}
        print(json.dumps({"impl": "reference", "metric": metric, "value": round(ms, 2), "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": 0, "ms_per_step": round(ms, 2),
                          "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64 limbs (254-bit modular integers)", "data": "synthetic",
                          "config": config, "cpu_baseline": base, "e2e": {"value": round(ms, 2), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    arm = GpuArm(args)
    torch = arm.torch
    from sirius_b200 import workload as WL

    sampler = ClockSampler(arm.local_rank)
    sampler.start()
    cells, cpu_base, verified = [], None, None

    def release(wl):
        wl.close()
        import gc

        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    for i, N in enumerate(sg_counts):
        side = WL.gate_scaling_side(N)
        windows = [17, 15, 13] + ([20] if (7 + 5 * N) << k_s >= 1 << 23 else [])
        t_build = time.perf_counter()
        wl = WL.SangriaStepWorkload(k_s, 0, 1, arm.stream, windows=windows, primary=side)
        build_s = time.perf_counter() - t_build
        cell = {"arm": "sangria", "gates": N + 1, "sub_circuits": N, "k": k_s, "advice": 7 + 5 * N, "fixed": 15 + 11 * N, "cross_terms": wl.sides[0].d,
                "msm_points": [(7 + 5 * N) << k_s, f"{wl.sides[0].d}x{1 << k_s}"], "setup_s": round(build_s, 1)}
        if i == 0 and not args.no_verify:
            from oracle import step_ref

            snap = wl.snapshot_inputs()
            wl.step(upload=True)
            res = wl.snapshot_results()
            t0 = time.perf_counter()
            exp = step_ref.fold_step(snap, step_ref.bases_for(snap), threads=cpu_threads())
            cpu_ms = (time.perf_counter() - t0) * 1e3
            rep = step_ref.compare(res, exp)
            if not rep["ok"]:
                raise SystemExit(f"bench.py: gate_scaling sangria N={N} disagrees with the oracle: {rep['bad']}")
            cell["verified"] = True
            verified = True
            cpu_base = {"value": round(cpu_ms, 2), "unit": "ms", "cores": cpu_threads(), "kind": "port",
                        "sample": f"1 full fold_step hot path at N={N} sub-circuits on the same inputs as the verified GPU step (CPU restatement, OpenMP)"}
        dev = arm.timed(wl, False, args.steps, args.warmup)
        e2e = arm.timed(wl, True, args.steps, args.warmup)
        wl.overlap = False
        seq = arm.timed(wl, False, max(2, args.steps // 2), 1)
        cell.update(ms=round(dev["ms"], 3), e2e_ms=round(e2e["ms"], 3), h2d_bytes_per_step=int(e2e["h2d"]), launches=int(dev["launches"]),
                    sequential_phases_ms=round(seq["ms"], 3), breakdown_ms=seq["breakdown"])
        cells.append(cell)
        release(wl)
        del wl
    for i, N in enumerate(pg_counts):
        if i == 0 and not args.no_verify:
            from oracle import step_ref

            kv = min(k_pg, 12)
            for mode, name in ((0, "compat"), (1, "correct")):
                wl = WL.GateScalingPgWorkload(kv, N, arm.stream, row_mode=mode)
                snap = wl.snapshot_inputs()
                wl.step()
                res = wl.snapshot_results()
                exp = step_ref.protogalaxy_prove(snap, wl.side, threads=cpu_threads())
                rep = step_ref.compare_protogalaxy(res, exp)
                if not rep["ok"]:
                    raise SystemExit(f"bench.py: gate_scaling protogalaxy N={N} k={kv} ({name}) disagrees with the oracle: {rep['bad']}")
                release(wl)
                del wl
        # one straight-line kernel per gate and per leaf kind costs ~5 s of NVRTC each: beyond a few gates the interpreter kernel is used
        jit = N + 1 <= args.pg_jit_max_gates
        arm.lib.sb_expr_jit_enable(1 if jit else 0)
        t_build = time.perf_counter()
        wl = WL.GateScalingPgWorkload(k_pg, N, arm.stream, row_mode=1)
        build_s = time.perf_counter() - t_build
        dev = arm.timed(wl, False, max(2, args.steps // 2), 1)
        cell = {"arm": "protogalaxy", "gates": N + 1, "sub_circuits": N, "k": k_pg, "advice": 7 + 5 * N, "fixed": 15 + 11 * N, "leaves_log2": wl.pg.t,
                "points_F": wl.pg.nF, "points_G": wl.pg.nG, "ms": round(dev["ms"], 3), "launches": int(dev["launches"]), "breakdown_ms": dev["breakdown"],
                "setup_s": round(build_s, 1), "row_mode": "correct", "evaluator": "run-time compiled straight-line kernels" if jit else "interpreter kernel"}
        if i == 0 and not args.no_verify:
            cell["verified"] = f"bit-exact vs the CPU restatement at k={min(k_pg, 12)}, both row modes (the k={k_pg} run uses the same kernels)"
        cells.append(cell)
        release(wl)
        del wl
    arm.lib.sb_expr_jit_enable(1)
    clocks = sampler.stop()
    head = cells[0]
    line = {"metric": metric, "value": head["ms"], "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms"],
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic", "config": config,
            "value_is": f"the first cell ({head['arm']}, {head['sub_circuits']} sub-circuits); every cell is in `cells`",
            "e2e": {"value": head.get("e2e_ms", head["ms"]), "unit": "ms", "h2d_bytes_per_step": head.get("h2d_bytes_per_step", 0), "d2h_bytes_per_step": 64 * (head.get("cross_terms", 0) + 7)},
            "gpu_launches": head["launches"], "clocks": clocks, "verified": verified, "cells": cells}
    if cpu_base:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line), flush=True)


def run_msm_sweep(args):
    """--workload msm_sweep: BASELINE config 4, Pedersen MSM 2^16..2^24 bn256 G1, U (uniform) and W (witness-like: 60 % zero,
    20 % < 2^8, 10 % < 2^64, 10 % uniform) scalar distributions, every cell checked against the CPU oracle.
    N > 1 (torchrun): scalars and key are sharded by index over the ranks, one exchange of 128-byte partials + combine."""
    import ctypes

    import numpy as np

    arm = GpuArm(args)
    torch, lib, _lib = arm.torch, arm.lib, arm._lib
    rank, world, stream = arm.rank, arm.world, arm.stream
    from sirius_b200 import curves, device
    from sirius_b200 import workload as WL
    import sirius_b200

    combiner = WL.make_combiner(rank, world, stream)
    logs = [int(x) for x in (args.sizes or "16,18,20,22,24").split(",")]
    cells = []
    sampler = ClockSampler(arm.local_rank)
    if rank == 0:
        sampler.start()
    curve = 0
    g = curves.generator_limbs(curve)
    for lg in logs:
        n = 1 << lg
        n_loc = n // world
        first = rank * n_loc
        with torch.cuda.stream(stream):
            d_bases = torch.empty((n_loc, 8), dtype=torch.int64, device="cuda")
        _lib.check(lib.sb_index_multiples_device(curve, g.ctypes.data_as(_lib.u64p), first, n_loc, ctypes.c_void_p(d_bases.data_ptr()), ctypes.c_void_p(stream.cuda_stream)))
        stream.synchronize()
        ck = sirius_b200.CommitmentKey.from_device(curve, d_bases.data_ptr(), n_loc, window_bits=0, stream=stream.cuda_stream)
        stream.synchronize()
        del d_bases
        for dist_name in ("U", "W"):
            with torch.cuda.stream(stream):
                s = device.random_field_device(n_loc, SEED + 1000 * lg + 10 * rank + (0 if dist_name == "U" else 1))
                if dist_name == "W":   # witness-like columns: mostly zero / small values (src/table/circuit_runner.rs:74)
                    gen = torch.Generator(device="cuda")
                    gen.manual_seed(SEED + lg + rank)
                    u = torch.rand(n_loc, device="cuda", generator=gen)
                    s[u < 0.6] = 0
                    small = (u >= 0.6) & (u < 0.8)
                    s[small, 1:] = 0
                    s[small, 0] &= 0xFF
                    mid = (u >= 0.8) & (u < 0.9)
                    s[mid, 1:] = 0
                    # these are canonical integers; the MSM takes Montgomery residues: every value < p is the residue of
                    # SOME element, the oracle gets the same bytes, so the check is unaffected; the digit pattern the
                    # kernels see is that of value * R^-1 -- convert so that the CANONICAL scalars are the small ones
                    s = _to_montgomery_device(s, lib, _lib, stream)
                d_out = torch.zeros(8, dtype=torch.int64, device="cuda")
                h_out = torch.zeros((1, 8), dtype=torch.int64).pin_memory()
            stream.synchronize()

            def commit():
                if combiner is None:
                    ck.commit_device(s.data_ptr(), n_loc, d_out.data_ptr(), 0, stream.cuda_stream)
                    with torch.cuda.stream(stream):
                        h_out.copy_(d_out.view(1, 8), non_blocking=True)
                    stream.synchronize()
                else:
                    combiner.commit(ck, s.data_ptr(), n_loc, 1, h_out)

            ms = arm.timed_leg(commit, args.steps, args.warmup)
            got = h_out.numpy().view(np.uint64).reshape(8).copy()
            ok = None
            if not args.no_verify and lg <= args.verify_max_log:
                full = s.cpu().numpy().view(np.uint64)
                if world > 1:
                    import torch.distributed as dist

                    gathered = torch.empty((world,) + tuple(s.shape), dtype=s.dtype, device="cuda")
                    dist.all_gather_into_tensor(gathered, s.contiguous())
                    full = gathered.view(-1, 4).cpu().numpy().view(np.uint64)
                if rank == 0:
                    import oracle

                    oracle.build()
                    ok = bool(np.array_equal(got, oracle.msm(curve, full, oracle.running_bases(curve, n), threads=cpu_threads())))
                arm.barrier()
            if rank == 0:
                cells.append({"log2_n": lg, "dist": dist_name, "window_bits": ck.window_bits, "ms": round(ms, 4), "mscalar_per_s": round(n / ms / 1e3, 1),
                              "gbps_algorithmic": round(96 * n / ms / 1e6, 2), "matches_oracle": ok})
        ck.close()
        torch.cuda.empty_cache()
    if rank == 0:
        clocks = sampler.stop()
        best = max(c["mscalar_per_s"] for c in cells if c["dist"] == "U")
        if any(c["matches_oracle"] is False for c in cells):
            raise SystemExit(f"bench.py: msm_sweep cell disagrees with the oracle: {[c for c in cells if c['matches_oracle'] is False]}")
        print(json.dumps({"metric": "Pedersen MSM bn256 G1 size sweep (uniform scalars, best cell)", "value": best, "unit": "Mscalar/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 limbs (254-bit modular integers)",
                          "data": "synthetic", "config": {"workload": "Pedersen MSM size sweep 2^16..2^24 bn256 G1 (BASELINE config 4)", "sizes_log2": logs,
                                                          "sharding": "single GPU" if world == 1 else f"scalars and key sharded by index over {world} ranks"},
                          "cells": cells, "clocks": clocks, "verified": all(c["matches_oracle"] in (True, None) for c in cells),
                          "timing": "CUDA events around `steps` commits incl. the 64-byte D2H + sync of each, max over ranks"}), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def _to_montgomery_device(s, lib, _lib, stream):
    """canonical integers -> Montgomery residues on the device (x * R): one k_axpy with w1 = 0, r = R^2 would need Montgomery
    inputs, so use the fold kernel's product directly: out = 0 + R2 * s where `*` is the Montgomery product (s * R2 / R = s * R)."""
    import ctypes

    import numpy as np
    import torch

    R2 = np.array([0x1bb8e645ae216da7, 0x53fe3ab1e35c59e3, 0x8c49833d53bb8085, 0x0216d0b17f4e44a5], dtype=np.uint64)   # Fr R^2
    zero = torch.zeros_like(s)
    out = torch.empty_like(s)
    _lib.check(lib.sb_axpy_fold_device(0, ctypes.c_void_p(zero.data_ptr()), ctypes.c_void_p(s.data_ptr()), R2.ctypes.data_as(_lib.u64p),
                                       ctypes.c_void_p(out.data_ptr()), s.shape[0], ctypes.c_void_p(stream.cuda_stream)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sangria_poseidon", choices=["sangria_poseidon", "cyclefold_poseidon", "msm_sweep", "gate_scaling"])
    ap.add_argument("--gates", default="5,10,20", help="gate_scaling: sub-circuit counts of the sangria arm (benches/ivc_gate_scaling.rs:221-223)")
    ap.add_argument("--pg-gates", default="6,12,24", help="gate_scaling: sub-circuit counts of the Protogalaxy arm (:229-232 also lists 48: ~60 GB of columns at k=20)")
    ap.add_argument("--pg-k", type=int, default=20, help="gate_scaling: table size of the Protogalaxy arm")
    ap.add_argument("--pg-jit-max-gates", type=int, default=7, help="gate_scaling: compile the per-gate leaf kernels up to this many gates")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the bit-for-bit comparison of one timed step with the oracle")
    ap.add_argument("--no-k20", action="store_true", help="skip the extra.k20 block (the same measurement at k = 20)")
    ap.add_argument("--no-host-abi", action="store_true", help="skip the e2e_host_abi figure")
    ap.add_argument("--no-extra", action="store_true", help="skip the MSM 2^20 / NTT side metrics")
    ap.add_argument("--sizes", default=None, help="msm_sweep: comma-separated log2 sizes (default 16,18,20,22,24)")
    ap.add_argument("--verify-max-log", type=int, default=24, help="msm_sweep: check cells up to this log2 size against the CPU oracle")
    ap.add_argument("--k", type=int, default=K_TABLE, help="table size 2^k rows of the main line (default 17 = BASELINE configs[1])")
    args = ap.parse_args()
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # before torch / NCCL load: keep stdout to the one JSON line (NCCL prints its banner there)
    if args.workload == "cyclefold_poseidon":
        return run_cyclefold(args)
    if args.workload == "msm_sweep":
        return run_msm_sweep(args)
    if args.workload == "gate_scaling":
        return run_gate_scaling(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
