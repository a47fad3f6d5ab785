"""The Sangria `fold_step` prover hot path at the shapes of benches/sangria_poseidon, as a device-resident session.

One `step()` is the hot path of `IVC::fold_step` (reference src/ivc/sangria/incrementally_verifiable_computation.rs:428-635,
SURVEY 3.1), in the reference's call order:

  1. VanillaFS::prove, secondary side (grumpkin; A=7, F=15, 1 gate, d=5): 5 cross-term vectors + 5 commits, W/E fold
  2. generate_plonk_trace, primary side (bn256): commit W, 12 * 2^k scalars
  3. VanillaFS::prove, primary side (bn256; A=12, F=26, 2 gates, d=6): 6 cross terms + 6 commits + folds
  4. generate_plonk_trace, secondary side: commit W, 7 * 2^k scalars

plus, as separately timed stages, the rest of the measured bench iteration (benches/sangria_poseidon.rs:159-175,
`IVC::fold(.., 1)` = `new` + `fold_step` + `verify`): `new_leg()` = the two W commits of `IVC::new`, `verify_leg()` = the
deciders (src/nifs/sangria/mod.rs:334-383, 455-474; src/plonk/mod.rs:304-361).

This module is what bench.py times AND what tests/test_gpu_workload.py checks against the oracle (it never imports the
oracle itself).  N > 1: rows are sharded over the ranks (sharding.py), every commitment group is one exchange of 128-byte
XYZZ partial sums + a combine kernel.  torch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import Dict, List, Optional

import numpy as np

from . import _lib, curves, device, sharding
from . import polynomial as P
from . import sangria as SG
from .commitment import CommitmentKey

PRIMARY = dict(name="primary", curve=0, field=0, T_list=[5, 3])   # bn256 / Fr : MainGate<5> + Poseidon MainGate<3>
SECONDARY = dict(name="secondary", curve=1, field=1, T_list=[5])  # grumpkin / Fq : MainGate<5> (trivial step circuit)
# Cyclefold support circuit (src/ivc/cyclefold/support_circuit/tiny_gate.rs:38-84, k = 15): 1 selector, 4 fixed, 3 advice, d = 2
SUPPORT = dict(name="support", curve=1, field=1, kind="tiny", T_list=[])
SEED = 0x5349524955530000


def gate_scaling_side(gates_count: int):
    """Primary circuit of benches/ivc_gate_scaling.rs:131-140 (`MultiStepCircuit` of GATES_COUNT Poseidon step circuits): the
    step-folding circuit's MainGate<5> plus one MainGate<3> per sub-circuit: A = 7 + 5N advice, F = 15 + 11N
    fixed columns, N + 1 gates of degree 5 -- compressed with powers of one challenge, so the folding degree is 5 + N."""
    return dict(name="primary", curve=0, field=0, T_list=[5] + [3] * int(gates_count))


def shapes(side):
    if side.get("kind") == "tiny":
        return 4, 3
    nfix = sum(2 * T + 5 for T in side["T_list"])
    nadv = sum(T + 2 for T in side["T_list"])
    return nfix, nadv


def num_selectors(side) -> int:
    return 1 if side.get("kind") == "tiny" else 0


def compressed_gates(side, mod=P):
    """The gate expressions of one side, built with `mod` = sirius_b200.polynomial (product) or the oracle's expr_ref
    (same constructor names) -- the callers pass their own module so this file never imports oracle/."""
    nfix, nadv = shapes(side)
    if side.get("kind") == "tiny":
        return [mod.tiny_gate_expression()], nfix, nadv
    gates, fb, ab = [], 0, 0
    for T in side["T_list"]:
        gates.append(mod.main_gate_expression(T, fb, ab, 0, nfix))
        fb += 2 * T + 5
        ab += T + 2
    return gates, nfix, nadv


def default_windows(k: int) -> List[int]:
    """Window widths registered per key: the W commits want wide windows, the batched 2^k cross-term commits narrower ones."""
    env = os.environ.get("SB_BENCH_WINDOWS")
    if env:
        return [int(x) for x in env.split(",")]
    if k >= 19:
        return [17, 20]
    return [16, 13, 15, 17]


def build_structure_key(side, k, rank, world, stream, windows, seed, key_cols=None):
    """PlonkStructure + CommitmentKey of one side restricted to this rank's rows: synthetic uniform fixed columns and
    selector bits generated in HBM; key = [i+1]G for the first key_cols (default: num_advice) * 2^k indices."""
    import torch

    lib = _lib.load()
    gates, nfix, nadv = compressed_gates(side)
    nsel = num_selectors(side)
    cg = P.CompressedGates.new(gates, P.QueryIndexContext(num_selectors=nsel, num_fixed=nfix, num_advice=nadv))
    n = 1 << k
    row0, n_loc = sharding.row_slice(rank, world, n)
    k_loc = n_loc.bit_length() - 1
    modulus = curves.SCALAR_FIELD[side["curve"]]
    with torch.cuda.stream(stream):
        d_fixed = [device.random_field_device(n_loc, seed + 1000 * side["curve"] + 10 * rank + j) for j in range(nfix)]
        gsel = torch.Generator(device="cuda")
        gsel.manual_seed(seed + 555 + rank)
        d_sel = [torch.randint(0, 2, (n_loc,), dtype=torch.uint8, device="cuda", generator=gsel) for _ in range(nsel)]
    stream.synchronize()   # .cpu() below runs on the default stream
    fixed = [t.cpu().numpy().view(np.uint64) for t in d_fixed]
    selectors = [t.cpu().numpy() for t in d_sel]
    del d_fixed, d_sel
    S = SG.PlonkStructure(side["field"], modulus, k_loc, selectors, fixed, nadv, 0, cg, gates=gates)
    rotations = P.GraphEvaluator.new(cg.homogeneous, modulus).rotations
    if world > 1:
        sharding.check_rotations_row_local(rotations)
    # commitment key restricted to this rank's rows: ck[col * n + row] for row in the slice, column-major
    # (the benches' key has 2^(k+4) generators, benches/sangria_poseidon.rs:26-30; only the prefix W needs is materialised)
    kc = key_cols or nadv
    with torch.cuda.stream(stream):
        d_bases = torch.empty((kc * n_loc, 8), dtype=torch.int64, device="cuda")
    g = curves.generator_limbs(side["curve"])
    for col, (first, count) in enumerate(sharding.key_segments(kc, n, rank, world)):
        _lib.check(lib.sb_index_multiples_device(side["curve"], g.ctypes.data_as(_lib.u64p), first, count,
                                                 ctypes.c_void_p(d_bases.data_ptr() + col * n_loc * 64), ctypes.c_void_p(stream.cuda_stream)))
    stream.synchronize()
    ck = CommitmentKey.from_device(side["curve"], d_bases.data_ptr(), kc * n_loc, window_bits=windows[0], stream=stream.cuda_stream)
    for wb in windows[1:]:
        ck.add_window(wb, stream.cuda_stream)
    stream.synchronize()
    del d_bases
    return S, ck, cg, dict(side=side, nadv=nadv, nfix=nfix, n_loc=n_loc, fixed=fixed, selectors=selectors, row_local=all(int(r) == 0 for r in rotations))


def build_sangria_side(side, k, rank, world, stream, windows, seed, key_cols=None):
    """Structure + key + device session (device.DeviceSangriaSide) for one curve, restricted to this rank's rows."""
    import torch

    S, ck, cg, info = build_structure_key(side, k, rank, world, stream, windows, seed, key_cols)
    nadv, n_loc = info["nadv"], info["n_loc"]
    sess = device.DeviceSangriaSide(S, ck, stream)
    with torch.cuda.stream(stream):
        sess.W_acc.copy_(device.random_field_device(nadv * n_loc, seed + 7 + side["curve"] + 100 * rank))
        sess.E_acc.copy_(device.random_field_device(n_loc, seed + 8 + side["curve"] + 100 * rank))
        sess.W_in.copy_(device.random_field_device(nadv * n_loc, seed + 9 + side["curve"] + 100 * rank))
    stream.synchronize()
    host_W = torch.empty(sess.W_in.shape, dtype=torch.int64).pin_memory()
    host_W.copy_(sess.W_in)
    torch.cuda.synchronize()
    nch = cg.ctx.num_challenges - 1
    with torch.cuda.stream(stream):
        d_ch = device.random_field_device(2 * nch + 2, seed + 11 + side["curve"])   # identical on every rank
    stream.synchronize()
    ch = d_ch.cpu().numpy().view(np.uint64)
    extra = dict(info, host_W=host_W, c1=ch[:nch], c2=ch[nch:2 * nch], u1=ch[2 * nch], r=ch[2 * nch + 1])
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
            with torch.cuda.stream(self.stream):
                self.bufs[batch] = (torch.zeros((batch, 16), dtype=torch.int64, device="cuda"),
                                    torch.zeros((self.world, batch, 16), dtype=torch.int64, device="cuda"),
                                    torch.zeros((batch, 8), dtype=torch.int64, device="cuda"))
        return self.bufs[batch]

    def commit(self, ck: CommitmentKey, d_scalars: int, n: int, batch: int, h_out) -> None:
        import torch
        import torch.distributed as dist

        lib = _lib.load()
        part, gathered, out = self._buffers(batch)
        ck.commit_batch_device(d_scalars, n, n, batch, 0, part.data_ptr(), self.stream.cuda_stream)
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(gathered, part)
        _lib.check(lib.sb_msm_combine_batch_device(ck.curve, ctypes.c_void_p(gathered.data_ptr()), self.world, batch, batch,
                                                   ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(self.stream.cuda_stream)))
        with torch.cuda.stream(self.stream):
            h_out.copy_(out.view(h_out.shape), non_blocking=True)
        self.stream.synchronize()


class PeerCombiner:
    """N > 1, the default: the exchange is fused into the last kernel of the commitment pipeline (csrc/comm.cu): every rank
    stores its 128-byte partial sums into the peers' mailboxes over NVLink (CUDA IPC peer memory), waits for theirs, adds
    them and normalises -- no NCCL launch and no extra kernels on the critical path of a commitment group.
    torch.distributed is used once, to all-gather the 64-byte IPC handles."""

    MAX_BATCH = 64

    def __init__(self, rank, world, stream, alone: bool = False):
        """alone = True: a one-rank communicator (the rank exchanges with itself) -- tools/shard_profile.py uses it to time one
        rank's share of an N-GPU step, exchange kernel included, on a single GPU."""
        import torch

        self.rank, self.world, self.stream, self.torch = rank, world, stream, torch
        self.alone = alone
        self.lib = _lib.load()
        self._h = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        if alone:
            _lib.check(self.lib.sb_comm_create(0, 1, self.MAX_BATCH, ctypes.byref(self._h), handle))
            _lib.check(self.lib.sb_comm_connect(self._h, handle.raw))
        else:
            import torch.distributed as dist

            _lib.check(self.lib.sb_comm_create(rank, world, self.MAX_BATCH, ctypes.byref(self._h), handle))
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).cuda()
            gathered = torch.empty((world, 64), dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(gathered, mine)
            torch.cuda.synchronize()
            _lib.check(self.lib.sb_comm_connect(self._h, gathered.cpu().numpy().tobytes()))
            dist.barrier()   # every mailbox exists and is zeroed before the first store into it
        torch.cuda.synchronize()
        self.out = {}

    def commit(self, ck: CommitmentKey, d_scalars: int, n: int, batch: int, h_out, sync: bool = True) -> None:
        """sync = False: enqueue only (pipeline + exchange + D2H of the result); the caller synchronises self.stream."""
        torch = self.torch
        if batch not in self.out:
            with torch.cuda.stream(self.stream):
                self.out[batch] = torch.zeros((batch, 8), dtype=torch.int64, device="cuda")
            self.stream.synchronize()
        out = self.out[batch]
        _lib.check(self.lib.sb_msm_batch_sharded_device(ck._h, self._h, ctypes.c_void_p(d_scalars), n, n, batch, ctypes.c_void_p(out.data_ptr()),
                                                        ctypes.c_void_p(self.stream.cuda_stream)))
        with torch.cuda.stream(self.stream):
            h_out.copy_(out.view(h_out.shape), non_blocking=True)
        if sync:
            self.stream.synchronize()

    def status(self) -> int:
        return int(self.lib.sb_comm_status(self._h, ctypes.c_void_p(self.stream.cuda_stream)))

    def close(self):
        if self._h.value:
            self.lib.sb_comm_destroy(self._h)
            self._h = ctypes.c_void_p()


def make_combiner(rank, world, stream):
    """SB_BENCH_EXCHANGE=nccl selects the library path (NCCL all-gather + combine kernel) for A/B measurements."""
    if world <= 1:
        return None
    if os.environ.get("SB_BENCH_EXCHANGE", "peer") == "nccl":
        return Combiner(world, stream)
    return PeerCombiner(rank, world, stream)


class SangriaStepWorkload:
    """Both sides of the cycle, device-resident, restricted to this rank's rows."""

    def __init__(self, k: int, rank: int = 0, world: int = 1, stream=None, windows: Optional[List[int]] = None, seed: int = SEED, combiner="auto",
                 overlap: Optional[bool] = None, primary=None, secondary=None):
        import torch

        self.torch = torch
        self.k, self.rank, self.world, self.seed = k, rank, world, seed
        self.side_desc = (primary or PRIMARY, secondary or SECONDARY)
        self.stream = stream if stream is not None else torch.cuda.Stream()
        self.windows = windows or default_windows(k)
        self.lib = _lib.load()
        self.sides: List[device.DeviceSangriaSide] = []
        self.extras: List[Dict] = []
        for side in self.side_desc:
            sess, ex = self._build_side(side)
            self.sides.append(sess)
            self.extras.append(ex)
        self.combiner = make_combiner(rank, world, self.stream) if combiner == "auto" else combiner
        # Two-stream phases (see step()): the trace commitment runs on a second stream beside the cross terms and
        # their commitments.  N > 1: that stream needs a communicator of its own (mailboxes and sequence numbers are per
        # communicator); the NCCL library path (SB_BENCH_EXCHANGE=nccl) keeps the sequential phases.
        if overlap is None:
            overlap = os.environ.get("SB_BENCH_OVERLAP", "1") != "0"
        self.overlap = overlap and (self.combiner is None or isinstance(self.combiner, PeerCombiner))
        self.aux_stream, self.combiner_w, self._ev = None, None, None
        self.upload_blocks = max(1, int(os.environ.get("SB_BENCH_UPLOAD_BLOCKS", "4")))
        self.host_enqueue_s = 0.0
        self.ct_first = os.environ.get("SB_BENCH_CT_FIRST", "1") == "1"   # enqueue the cross terms before the W commitment
        if self.overlap:
            # A/B knobs.  Measured (tools/shard_profile.py, k = 17): equal priorities + cross terms enqueued first 13.72 ms on one GPU and
            # 2.69 ms for one rank of eight, against 14.47 / 3.05 ms with a high-priority W stream enqueued first (the W pipeline's
            # ~17 launches cost the host ~0.1 ms during which the GPU had nothing of this phase to run)
            prio = os.environ.get("SB_BENCH_AUX_PRIORITY", "normal")   # "normal" (default) | "high"
            self.aux_stream = torch.cuda.Stream(priority=-1 if prio == "high" else 0)
            self.copy_stream = torch.cuda.Stream()
            self._ev = torch.cuda.Event()
            self._ev_blocks = [torch.cuda.Event() for _ in range(self.upload_blocks)]
            if self.combiner is not None:
                self.combiner_w = PeerCombiner(self.combiner.rank, self.combiner.world, self.aux_stream, alone=self.combiner.alone)
        self.stream.synchronize()

    # ------------------------------------------------------------------ construction
    def _build_side(self, side):
        return build_sangria_side(side, self.k, self.rank, self.world, self.stream, self.windows, self.seed)

    # ------------------------------------------------------------------ the timed step
    def _prove(self, sess, ex):
        if self.combiner is None:
            sess.commit_cross_terms(ex["c1"], ex["u1"], ex["c2"])
        else:
            c1, c2 = sess.challenge_vectors(ex["c1"], ex["u1"], ex["c2"])
            _lib.check(self.lib.sb_cross_terms_device(sess.S._hom_prog._h, sess.d, sess.S._cols, sess._cols(sess.W_acc), sess._cols(sess.W_in), sess.A,
                                                      c1.ctypes.data_as(_lib.u64p), c2.ctypes.data_as(_lib.u64p), c1.shape[0],
                                                      ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(sess.stream.cuda_stream)))
            self.combiner.commit(sess.ck, sess.T.data_ptr(), sess.n, sess.d, sess.h_commit_T)
        sess.fold(ex["r"])

    def _commit_w(self, sess, ex, upload) -> int:
        h2d = 0
        if upload:
            h2d = sess.upload_incoming(ex["host_W"])
        if self.combiner is None:
            sess.commit_incoming()
        else:
            self.combiner.commit(sess.ck, sess.W_in.data_ptr(), sess.A * sess.n, 1, sess.h_commit_W)
        return h2d

    def _trace_and_prove(self, sess, ex, upload) -> int:
        """One trace's share of a fold step as ONE device phase with ONE host synchronisation: `generate_plonk_trace`'s
        commitment of the fresh witness (src/plonk/mod.rs:441-445) on the second stream, beside `VanillaFS::prove` of the same
        trace (cross terms + their commitments, src/nifs/sangria/mod.rs:102-158) on the main stream.  Both depend on the trace's
        witness only; the random oracle needs the W commitment and the T commitments together, when it derives r
        (src/nifs/sangria/mod.rs:162-179), and the circuits of this bench have a single witness round (no challenge is
        squeezed between W and the cross terms).  What it buys: the latency tail of the W commitment (fix-up, row/column
        sums, weighted sum, exchange, normalisation: a few SMs busy) runs under the full-chip kernels of the other stream
        instead of alone, and a host round trip disappears."""
        torch = self.torch
        main, aux = self.stream, self.aux_stream
        t_host = time.perf_counter()
        c1, c2 = sess.challenge_vectors(ex["c1"], ex["u1"], ex["c2"])
        ct_args = (sess.S._hom_prog._h, sess.d, sess.S._cols, sess._cols(sess.W_acc), sess._cols(sess.W_in), sess.A,
                   c1.ctypes.data_as(_lib.u64p), c2.ctypes.data_as(_lib.u64p), c1.shape[0])
        nb = self.upload_blocks if (upload and ex["row_local"] and sess.n >= 1024 * self.upload_blocks) else 1
        h2d = 0
        if nb > 1:
            # The fresh witness arrives in row blocks on a copy stream; the cross terms of a block (row-local: the gates
            # query rotation 0 only) start as soon as it is resident, so the sweep runs under the rest of the transfer.
            # The commitment of W needs every scalar (the digits are sorted by bucket): it waits for the last block.
            copy, rows = self.copy_stream, sess.n // nb
            self._ev.record(main)        # the previous readers of W_in (this side's fold of the last step) are on the main stream
            copy.wait_event(self._ev)
            host_W = ex["host_W"]
            for b in range(nb):
                _lib.check(self.lib.sb_upload_rows_device(ctypes.c_void_p(host_W.data_ptr()), sess.A, sess.n, b * rows, rows,
                                                          ctypes.c_void_p(sess.W_in.data_ptr()), ctypes.c_void_p(copy.cuda_stream)))
                self._ev_blocks[b].record(copy)
            h2d = host_W.numel() * 8
            for b in range(nb):
                main.wait_event(self._ev_blocks[b])
                _lib.check(self.lib.sb_cross_terms_rows_device(*ct_args, b * rows, rows, ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(main.cuda_stream)))
            aux.wait_event(self._ev_blocks[nb - 1])
        else:
            if upload:
                h2d = sess.upload_incoming(ex["host_W"])
            self._ev.record(main)        # after the upload and after everything the previous phase left on the main stream
            aux.wait_event(self._ev)
        ct_first = self.ct_first and nb == 1
        if ct_first:
            _lib.check(self.lib.sb_cross_terms_device(*ct_args, ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(main.cuda_stream)))
        if self.combiner is None:
            sess.ck.commit_device(sess.W_in.data_ptr(), sess.A * sess.n, sess.commit_W.data_ptr(), 0, aux.cuda_stream)
            with torch.cuda.stream(aux):
                sess.h_commit_W.copy_(sess.commit_W, non_blocking=True)
        else:
            self.combiner_w.commit(sess.ck, sess.W_in.data_ptr(), sess.A * sess.n, 1, sess.h_commit_W, sync=False)
        if nb == 1 and not ct_first:
            _lib.check(self.lib.sb_cross_terms_device(*ct_args, ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(main.cuda_stream)))
        if self.combiner is None:
            sess.ck.commit_batch_device(sess.T.data_ptr(), sess.n, sess.n, sess.d, sess.commit_T.data_ptr(), 0, main.cuda_stream)
            with torch.cuda.stream(main):
                sess.h_commit_T.copy_(sess.commit_T, non_blocking=True)
        else:
            self.combiner.commit(sess.ck, sess.T.data_ptr(), sess.n, sess.d, sess.h_commit_T, sync=False)
        self.host_enqueue_s += time.perf_counter() - t_host   # host time spent enqueuing (diagnostic: tools/shard_profile.py)
        aux.synchronize()
        main.synchronize()               # W and T commitments are on the host: the random oracle derives r
        sess.fold(ex["r"])
        return h2d

    def step(self, upload: bool = False) -> int:
        """One fold_step hot path.  Returns the bytes copied host -> device.

        Sequential form (SB_BENCH_OVERLAP=0, and the NCCL exchange): the reference's call order 1-4 of the module docstring,
        one host synchronisation per commitment group.  Default form: two phases, secondary trace then primary trace, each
        `_trace_and_prove`.  The work and every result are the same (tests/test_gpu_workload.py and bench.py's verification
        compare both with the oracle); what moves is the commitment of the secondary trace, which the reference issues at
        the end of fold_step i (call 4) and consumes in VanillaFS::prove at the start of fold_step i+1 (call 1): nothing
        between the two depends on it, so the session issues call 4 of step i together with call 1 of step i+1 -- a timed
        step is {4 of the previous step, 1, 2, 3}, the steady state of a run of fold steps."""
        prim, sec = self.sides
        ep, es = self.extras
        if self.overlap:
            return self._trace_and_prove(sec, es, upload) + self._trace_and_prove(prim, ep, upload)
        h2d = 0
        self._prove(sec, es)                     # 1. fold the secondary accumulator
        h2d += self._commit_w(prim, ep, upload)  # 2. primary trace: commit W
        self._prove(prim, ep)                    # 3. fold the primary accumulator
        h2d += self._commit_w(sec, es, upload)   # 4. secondary trace: commit W
        return h2d

    # ------------------------------------------------------------------ the rest of the bench iteration
    def new_leg(self) -> None:
        """`IVC::new`: the first traces' W commits on both sides (incrementally_verifiable_computation.rs:240-330)."""
        for sess, ex in zip(self.sides, self.extras):
            self._commit_w(sess, ex, False)

    def verify_leg(self) -> Dict[str, int]:
        """`IVC::verify` deciders on both sides: is_sat_accumulation (homogeneous gate polynomial on (W_acc, challenges ++ [u])
        against E, row by row), is_sat_witness_commit (re-commit W_acc and E), PlonkStructure::is_sat of the incoming trace
        (compressed gate on W_in against zero + re-commit W_in).  Returns the mismatch counts this rank saw (the synthetic
        columns do not satisfy the relation; the work is the same)."""
        torch = self.torch
        out = {}
        for sess, ex in zip(self.sides, self.extras):
            S, st = sess.S, sess.stream.cuda_stream
            if getattr(sess, "_rows", None) is None:
                with torch.cuda.stream(sess.stream):
                    sess._rows = torch.empty((sess.n, 4), dtype=torch.int64, device="cuda")
                    sess._cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
                    sess._h_cnt = torch.zeros(2, dtype=torch.int64).pin_memory()
                    sess._E_commit = torch.zeros(8, dtype=torch.int64, device="cuda")
                    sess._h_E_commit = torch.zeros(8, dtype=torch.int64).pin_memory()
                if getattr(S, "_compressed_prog", None) is None:
                    S._compressed_prog = SG.Program(S.field, P.GraphEvaluator.new(S.custom_gates_lookup_compressed.compressed, S.modulus))
            ch_acc = np.ascontiguousarray(np.concatenate([ex["c1"].reshape(-1, 4), ex["u1"].reshape(1, 4)]), dtype=np.uint64)
            ch_in = np.ascontiguousarray(ex["c2"].reshape(-1, 4), dtype=np.uint64)
            # is_sat_accumulation: evaluate, compare with E
            _lib.check(self.lib.sb_expr_eval_device(S._hom_prog._h, S._cols, sess._cols(sess.W_acc), None, sess.A, ch_acc.ctypes.data_as(_lib.u64p), ch_acc.shape[0],
                                                    ctypes.c_void_p(sess._rows.data_ptr()), ctypes.c_void_p(st)))
            _lib.check(self.lib.sb_count_mismatch_device(S.field, ctypes.c_void_p(sess._rows.data_ptr()), ctypes.c_void_p(sess.E_acc.data_ptr()), sess.n,
                                                         ctypes.c_void_p(sess._cnt.data_ptr()), ctypes.c_void_p(st)))
            # PlonkStructure::is_sat of the incoming trace: compressed gate == 0 on every row
            _lib.check(self.lib.sb_expr_eval_device(S._compressed_prog._h, S._cols, sess._cols(sess.W_in), None, sess.A,
                                                    ch_in.ctypes.data_as(_lib.u64p) if ch_in.shape[0] else None, ch_in.shape[0],
                                                    ctypes.c_void_p(sess._rows.data_ptr()), ctypes.c_void_p(st)))
            _lib.check(self.lib.sb_count_mismatch_device(S.field, ctypes.c_void_p(sess._rows.data_ptr()), None, sess.n,
                                                         ctypes.c_void_p(sess._cnt.data_ptr() + 8), ctypes.c_void_p(st)))
            with torch.cuda.stream(sess.stream):
                sess._h_cnt.copy_(sess._cnt, non_blocking=True)
            # is_sat_witness_commit: W_acc, E, and the incoming W re-open
            self._recommit(sess, sess.W_acc.data_ptr(), sess.A * sess.n, sess.h_commit_W, sess.commit_W)
            self._recommit(sess, sess.E_acc.data_ptr(), sess.n, sess._h_E_commit, sess._E_commit)
            self._recommit(sess, sess.W_in.data_ptr(), sess.A * sess.n, sess.h_commit_W, sess.commit_W)
            out[ex["side"]["name"]] = (int(sess._h_cnt[0]), int(sess._h_cnt[1]))
        return out

    def _recommit(self, sess, d_scalars, n, h_out, d_out):
        if self.combiner is None:
            sess.ck.commit_device(d_scalars, n, d_out.data_ptr(), 0, sess.stream.cuda_stream)
            with self.torch.cuda.stream(sess.stream):
                h_out.copy_(d_out, non_blocking=True)
            sess.stream.synchronize()
        else:
            self.combiner.commit(sess.ck, d_scalars, n, 1, h_out)

    # ------------------------------------------------------------------ state in/out for verification
    def _gather_cm(self, t, ncols):
        """local column-major [ncols * n_loc, 4] -> global column-major host array [ncols * n, 4] (every rank gets it)."""
        torch = self.torch
        self.stream.synchronize()
        if self.world == 1:
            return t.cpu().numpy().view(np.uint64).copy()
        import torch.distributed as dist

        n_loc = t.shape[0] // ncols
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        torch.cuda.synchronize()
        full = out.view(self.world, ncols, n_loc, 4).permute(1, 0, 2, 3).contiguous().view(ncols * self.world * n_loc, 4)
        return full.cpu().numpy().view(np.uint64).copy()

    def snapshot_inputs(self) -> Dict[str, Dict]:
        """Host copies of everything one step reads, in the reference's global layout (all ranks must call)."""
        torch = self.torch
        snap = {}
        for sess, ex in zip(self.sides, self.extras):
            fixed_loc = torch.from_numpy(np.concatenate([f.reshape(-1, 4) for f in ex["fixed"]]).view(np.int64)).cuda()
            fixed = self._gather_cm(fixed_loc, ex["nfix"])
            n = 1 << self.k
            snap[ex["side"]["name"]] = dict(
                side=ex["side"], k=self.k, nadv=ex["nadv"], nfix=ex["nfix"], selectors=[],
                fixed=[fixed[j * n:(j + 1) * n] for j in range(ex["nfix"])],
                W1=self._gather_cm(sess.W_acc, ex["nadv"]), E1=self._gather_cm(sess.E_acc, 1), W2=self._gather_cm(sess.W_in, ex["nadv"]),
                c1=ex["c1"].copy(), c2=ex["c2"].copy(), u1=ex["u1"].copy(), r=ex["r"].copy(),
            )
        return snap

    def snapshot_results(self) -> Dict[str, Dict]:
        """Host copies of what the last step produced: the commitments the host read back and the folded accumulator."""
        res = {}
        for sess, ex in zip(self.sides, self.extras):
            res[ex["side"]["name"]] = dict(
                commits_T=sess.h_commit_T.numpy().view(np.uint64).copy(), commit_W=sess.h_commit_W.numpy().view(np.uint64).copy(),
                W=self._gather_cm(sess.W_acc, ex["nadv"]), E=self._gather_cm(sess.E_acc, 1),
            )
        return res

    def close(self):
        for c in (self.combiner, self.combiner_w):
            if c is not None and hasattr(c, "close"):
                c.close()
        if self.aux_stream is not None:
            self.aux_stream.synchronize()
            self.lib.sb_stream_release(ctypes.c_void_p(self.aux_stream.cuda_stream))   # the second stream's scratch
        for sess in self.sides:
            sess.S.close()
            sess.ck.close()


class CyclefoldStepWorkload:
    """The prover hot path of `cyclefold::IVC::next` (reference src/ivc/cyclefold/incrementally_verifiable_computation/mod.rs:210-335,
    SURVEY 3.2) at the shapes of benches/cyclefold_poseidon (primary: bn256, A=12, F=26, 2 gates; support circuit: grumpkin,
    k=15, tiny gate), single GPU, device-resident:

      1. ProtoGalaxy::prove(acc, [trace]) (src/nifs/protogalaxy/mod.rs:400-481): compute_F (2^(k+1) leaves, 32 points),
         compute_G (8 Lagrange blends), K (host), fold_witness
      2. fold_support_circuit (:404-473): commit of the support trace (3 * 2^15 scalars) + SangriaFS::prove (2 cross terms,
         2 commits, W/E fold)
      3. ProtoGalaxy::generate_plonk_trace of the next primary trace: the 12 * 2^k-point MSM

    Challenges (betas, delta, alpha, gamma, r) come from the host random oracle in the reference; here they are seeded
    values.  row_mode = ROW_CORRECT by default: in the reference-compatible mode every leaf evaluates row 0
    and a timing would be meaningless."""

    def __init__(self, k: int, stream=None, windows: Optional[List[int]] = None, seed: int = SEED, support_k: int = 15, row_mode: int = 1):
        import random

        import torch

        self.torch = torch
        self.k, self.seed, self.support_k = k, seed, support_k
        self.stream = stream if stream is not None else torch.cuda.Stream()
        st = self.stream
        S, ck, cg, info = build_structure_key(PRIMARY, k, 0, 1, st, windows or ([20] if k >= 19 else [17]), seed)
        self.pg = device.DeviceProtogalaxySide(S, ck, st, row_mode)
        self.info = info
        A, n = info["nadv"], 1 << k
        with torch.cuda.stream(st):
            self.pg.W_acc.copy_(device.random_field_device(A * n, seed + 21))
            self.pg.W_in.copy_(device.random_field_device(A * n, seed + 22))
        st.synchronize()
        self.host_W = torch.empty(self.pg.W_in.shape, dtype=torch.int64).pin_memory()
        self.host_W.copy_(self.pg.W_in)
        torch.cuda.synchronize()
        self.sup, self.sup_ex = build_sangria_side(SUPPORT, support_k, 0, 1, st, [13], seed + 3)
        rng = random.Random(seed)
        M = S.modulus
        self.betas = [rng.randrange(M) for _ in range(self.pg.t)]
        self.delta, self.alpha, self.gamma = rng.randrange(M), rng.randrange(M), rng.randrange(M)
        self.last = None
        st.synchronize()

    def step(self, upload: bool = False) -> int:
        h2d = 0
        self.last = self.pg.prove(self.betas, self.delta, self.alpha, self.gamma)      # 1. ProtoGalaxy::prove
        sup, ex = self.sup, self.sup_ex                                                  # 2. fold_support_circuit
        if upload:
            h2d += sup.upload_incoming(ex["host_W"])
        sup.commit_incoming()
        sup.commit_cross_terms(ex["c1"], ex["u1"], ex["c2"])
        sup.fold(ex["r"])
        if upload:                                                                       # 3. the next primary trace
            h2d += self.pg.upload_incoming(self.host_W)
        self.pg.commit_incoming()
        return h2d

    def snapshot_inputs(self) -> Dict:
        torch = self.torch
        self.stream.synchronize()
        pg, ex = self.pg, self.sup_ex
        cpu = lambda t: t.cpu().numpy().view(np.uint64).copy()  # noqa: E731
        return dict(
            k=self.k, row_mode=pg.row_mode, fixed=self.info["fixed"], nadv=self.info["nadv"], W_acc=cpu(pg.W_acc), W_in=cpu(pg.W_in),
            betas=list(self.betas), delta=self.delta, alpha=self.alpha, gamma=self.gamma,
            support=dict(side=SUPPORT, k=self.support_k, nadv=ex["nadv"], nfix=ex["nfix"], fixed=ex["fixed"], selectors=ex["selectors"],
                         W1=cpu(self.sup.W_acc), E1=cpu(self.sup.E_acc), W2=cpu(self.sup.W_in), c1=ex["c1"].copy(), c2=ex["c2"].copy(),
                         u1=ex["u1"].copy(), r=ex["r"].copy()),
        )

    def snapshot_results(self) -> Dict:
        self.stream.synchronize()
        pg, sup = self.pg, self.sup
        cpu = lambda t: t.cpu().numpy().view(np.uint64).copy()  # noqa: E731
        poly_F, poly_G, poly_K = self.last
        return dict(poly_F=poly_F, poly_G=poly_G, poly_K=poly_K, W=cpu(pg.W_acc), commit_W=pg.h_commit_W.numpy().view(np.uint64).copy(),
                    support=dict(commits_T=sup.h_commit_T.numpy().view(np.uint64).copy(), commit_W=sup.h_commit_W.numpy().view(np.uint64).copy(),
                                 W=cpu(sup.W_acc), E=cpu(sup.E_acc)))

    def close(self):
        self.pg.S.close()
        self.pg.ck.close()
        self.sup.S.close()
        self.sup.ck.close()


class GateScalingPgWorkload:
    """BASELINE config 5, Cyclefold arm of benches/ivc_gate_scaling.rs (:183-199: `cyclefold::IVC::next` over a primary circuit
    of N parallel Poseidon sub-circuits): the Protogalaxy side of one `next()` -- compute_F over 2^t leaves
    (t = k + ceil(log2(N + 1))), compute_G over the Lagrange blends, K, fold_witness -- on the synthetic shapes
    A = 7 + 5N, F = 15 + 11N, N + 1 gates.  No commitment is timed here: at k = 20 the trace has (7 + 5N) * 2^20 >= 38.8 M
    cells, more than the bench's 2^25-generator key (the reference itself would stop with TooLongInput), so the
    key of this workload holds one column only and is never used."""

    def __init__(self, k: int, gates_count: int, stream=None, seed: int = SEED, row_mode: int = 1):
        import random

        import torch

        self.torch = torch
        self.k, self.gates_count, self.seed = k, gates_count, seed
        self.stream = stream if stream is not None else torch.cuda.Stream()
        st = self.stream
        self.side = gate_scaling_side(gates_count)
        S, ck, cg, info = build_structure_key(self.side, k, 0, 1, st, [10], seed, key_cols=1)
        self.pg = device.DeviceProtogalaxySide(S, ck, st, row_mode)
        self.info = info
        A, n = info["nadv"], 1 << k
        with torch.cuda.stream(st):
            self.pg.W_acc.copy_(device.random_field_device(A * n, seed + 21))
            self.pg.W_in.copy_(device.random_field_device(A * n, seed + 22))
        st.synchronize()
        rng = random.Random(seed)
        M = S.modulus
        self.betas = [rng.randrange(M) for _ in range(self.pg.t)]
        self.delta, self.alpha, self.gamma = rng.randrange(M), rng.randrange(M), rng.randrange(M)
        self.last = None

    def step(self, upload: bool = False) -> int:
        self.last = self.pg.prove(self.betas, self.delta, self.alpha, self.gamma)
        return 0

    def snapshot_inputs(self) -> Dict:
        self.stream.synchronize()
        cpu = lambda t: t.cpu().numpy().view(np.uint64).copy()  # noqa: E731
        return dict(k=self.k, row_mode=self.pg.row_mode, side=self.side, fixed=self.info["fixed"], nadv=self.info["nadv"], W_acc=cpu(self.pg.W_acc),
                    W_in=cpu(self.pg.W_in), betas=list(self.betas), delta=self.delta, alpha=self.alpha, gamma=self.gamma)

    def snapshot_results(self) -> Dict:
        self.stream.synchronize()
        poly_F, poly_G, poly_K = self.last
        return dict(poly_F=poly_F, poly_G=poly_G, poly_K=poly_K, W=self.pg.W_acc.cpu().numpy().view(np.uint64).copy())

    def close(self):
        self.pg.S.close()
        self.pg.ck.close()
