"""The objects bench.py times (`sirius_b200.workload.SangriaStepWorkload` over `device.DeviceSangriaSide`), checked bit for
bit against the CPU restatement of the reference at the bench's own shapes: A=12,F=26,d=6 (bn256) and A=7,F=15,d=5
(grumpkin) at k=17 with the bench's four-width key, plus the `new` / `verify` legs and the multi-GPU path (torchrun)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_step(k, windows=None, steps=1, overlap=None):
    import torch

    from oracle import step_ref
    from sirius_b200 import workload as WL

    wl = WL.SangriaStepWorkload(k, 0, 1, torch.cuda.Stream(), windows=windows, overlap=overlap)
    assert wl.overlap == (overlap is not False)
    try:
        for i in range(steps):
            snap = wl.snapshot_inputs()
            wl.step(upload=(i % 2 == 0))
            got = wl.snapshot_results()
            exp = step_ref.fold_step(snap, step_ref.bases_for(snap))
            rep = step_ref.compare(got, exp)
            assert rep["ok"], rep["bad"]
            assert len(rep["checked"]) == 8
    finally:
        wl.close()


def test_fold_step_small_two_steps(oracle):
    _check_step(10, windows=[11, 8], steps=2)   # the second step folds the first step's accumulator


def test_fold_step_small_sequential_phases(oracle):
    _check_step(10, windows=[11, 8], steps=2, overlap=False)   # the reference's call order, one sync per commitment group


def test_fold_step_two_stream_phases_repeat(oracle):
    _check_step(12, windows=[12, 9], steps=4, overlap=True)    # W commit on the second stream beside cross terms + T commits


def test_fold_step_bench_shapes_k17(oracle):
    from sirius_b200 import workload as WL

    assert WL.default_windows(17) == [16, 13, 15, 17]
    _check_step(17)


def test_new_and_verify_legs(oracle):
    import torch

    from sirius_b200 import sangria as SG
    from sirius_b200 import workload as WL
    import oracle as O

    wl = WL.SangriaStepWorkload(9, 0, 1, torch.cuda.Stream(), windows=[10])
    try:
        wl.new_leg()
        for sess, ex in zip(wl.sides, wl.extras):   # IVC::new commits the incoming W
            W = sess.W_in.cpu().numpy().view(np.uint64)
            bases = O.running_bases(ex["side"]["curve"], W.shape[0])
            assert np.array_equal(sess.h_commit_W.numpy().view(np.uint64), O.msm(ex["side"]["curve"], W, bases))
        counts = wl.verify_leg()
        n = 1 << 9
        assert counts == {"primary": (n, n), "secondary": (n, n)}   # uniform synthetic columns satisfy no row
        # make the accumulator satisfy the relaxed relation: E := the evaluated rows -> is_sat_accumulation sees 0 mismatches
        for sess, ex in zip(wl.sides, wl.extras):
            ch = np.concatenate([ex["c1"].reshape(-1, 4), ex["u1"].reshape(1, 4)])
            rows = SG.evaluate_rows(sess.S, sess.S._hom_prog, [sess.W_acc.cpu().numpy().view(np.uint64)], ch)
            sess.E_acc.copy_(torch.from_numpy(rows.view(np.int64)).cuda())
            torch.cuda.synchronize()
        counts = wl.verify_leg()
        assert counts["primary"][0] == 0 and counts["secondary"][0] == 0
        for sess, ex in zip(wl.sides, wl.extras):   # the last re-commit of the leg is W_in again; E's commitment is checked too
            E = sess.E_acc.cpu().numpy().view(np.uint64)
            bases = O.running_bases(ex["side"]["curve"], E.shape[0])
            assert np.array_equal(sess._h_E_commit.numpy().view(np.uint64), O.msm(ex["side"]["curve"], E, bases))
    finally:
        wl.close()


def test_count_mismatch_device():
    import ctypes

    import torch

    from sirius_b200 import _lib

    lib = _lib.load()
    a = torch.randint(0, 2**62, (1000, 4), dtype=torch.int64, device="cuda")
    a[:, 3] &= 0x0FFFFFFFFFFFFFFF
    b = a.clone()
    b[17, 0] += 1
    b[999, 3] ^= 1
    a[5] = 0
    b[5] = 0
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.synchronize()
    _lib.check(lib.sb_count_mismatch_device(0, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), 1000, ctypes.c_void_p(cnt.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
    st.synchronize()
    assert int(cnt.item()) == 2
    _lib.check(lib.sb_count_mismatch_device(1, ctypes.c_void_p(a.data_ptr()), None, 1000, ctypes.c_void_p(cnt.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
    st.synchronize()
    assert int(cnt.item()) == 999   # row 5 is the only zero row
    lib.sb_stream_release(ctypes.c_void_p(st.cuda_stream))


def test_multi_gpu_step_torchrun(oracle):
    """The N > 1 path of the bench (row sharding + partial-sum exchange + combine) at the bench shapes, on real GPUs."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs on the box (run with gpurun --gpus 2)")
    world = 2
    port = 29600 + (os.getpid() % 1000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "check_multi_gpu.py"), "--k", "17"],
                       capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "fold_step over 2 ranks == oracle" in r.stdout


def test_two_threads_two_streams_reentrant(oracle):
    """SURVEY 8b "Threading": CommitmentKey is Sync and cargo test is multi-threaded.  Two host threads drive two CUDA
    streams through the _device entry points (commit, cross terms, folds) and two more call the blocking host entry
    points at the same time; every result must equal the oracle's (the per-stream scratch and the one-critical-section
    host wrappers are what is under test)."""
    import ctypes
    import threading

    import torch

    import oracle as O
    import sirius_b200
    from oracle import pyref as R
    from sirius_b200 import _lib

    lib = _lib.load()
    n = 1 << 14
    errors = []
    keys = {}
    for curve in (0, 1):
        bases = O.running_bases(curve, n)
        keys[curve] = (sirius_b200.CommitmentKey(curve, bases, window_bits=11), bases)

    def device_worker(tid):
        try:
            curve = tid & 1
            ck, bases = keys[curve]
            st = torch.cuda.Stream()
            for it in range(6):
                m = n - 37 * it - tid
                s = O.random_field(curve, 1000 * tid + it, m)
                with torch.cuda.stream(st):
                    d_s = torch.from_numpy(s.view(np.int64)).cuda()
                    d_o = torch.zeros(8, dtype=torch.int64, device="cuda")
                st.synchronize()
                ck.commit_device(d_s.data_ptr(), m, d_o.data_ptr(), 0, st.cuda_stream)
                # a fold on the same stream in between (shares nothing with the other thread's stream)
                r = O.random_field(curve, 77 + it, 1).reshape(4)
                with torch.cuda.stream(st):
                    d_w = torch.empty_like(d_s)
                _lib.check(lib.sb_axpy_fold_device(curve, ctypes.c_void_p(d_s.data_ptr()), ctypes.c_void_p(d_s.data_ptr()), r.ctypes.data_as(_lib.u64p),
                                                   ctypes.c_void_p(d_w.data_ptr()), m, ctypes.c_void_p(st.cuda_stream)))
                st.synchronize()
                got = d_o.cpu().numpy().view(np.uint64)
                if not np.array_equal(got, O.msm(curve, s, bases)):
                    errors.append(f"device thread {tid} iteration {it}: commitment mismatch")
                exp_w = O.field_binop("add", curve, s, O.field_binop("mul", curve, np.tile(r, (m, 1)), s))
                if not np.array_equal(d_w.cpu().numpy().view(np.uint64), exp_w):
                    errors.append(f"device thread {tid} iteration {it}: fold mismatch")
            lib.sb_stream_release(ctypes.c_void_p(st.cuda_stream))
        except Exception as exc:  # noqa: BLE001
            errors.append(f"device thread {tid}: {exc!r}")

    def host_worker(tid):
        try:
            curve = tid & 1
            ck, bases = keys[curve]
            for it in range(6):
                m = n // 2 + 11 * it + tid
                s = O.random_field(curve, 5000 * tid + it, m)
                if not np.array_equal(ck.commit(s), O.msm(curve, s, bases)):
                    errors.append(f"host thread {tid} iteration {it}: commitment mismatch")
                r = O.random_field(curve, 9 + it, 1).reshape(4)
                out = np.zeros_like(s)
                _lib.check(lib.sb_axpy_fold(curve, s.ctypes.data_as(_lib.u64p), s.ctypes.data_as(_lib.u64p), r.ctypes.data_as(_lib.u64p), out.ctypes.data_as(_lib.u64p), m))
                exp = O.field_binop("add", curve, s, O.field_binop("mul", curve, np.tile(r, (m, 1)), s))
                if not np.array_equal(out, exp):
                    errors.append(f"host thread {tid} iteration {it}: fold mismatch")
        except Exception as exc:  # noqa: BLE001
            errors.append(f"host thread {tid}: {exc!r}")

    threads = [threading.Thread(target=device_worker, args=(t,)) for t in range(2)] + [threading.Thread(target=host_worker, args=(t,)) for t in range(2, 4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for ck, _ in keys.values():
        ck.close()
    assert not errors, errors


@pytest.mark.parametrize("k,row_mode", [(12, 0), (12, 1), (14, 1)])
def test_cyclefold_next_step_vs_c_oracle(oracle, k, row_mode):
    """Protogalaxy F / G / K / fold_witness + the support-circuit fold + the trace commitment of cyclefold::IVC::next at
    k >= 12, device-resident, against the C-interpreter oracle (oracle/pg_fast.py pinned to the literal pg_ref.py on the CPU);
    both leaf-row modes (reference-compatible `& 2^k` and corrected `% 2^k`)."""
    import torch

    from oracle import step_ref
    from sirius_b200 import workload as WL

    wl = WL.CyclefoldStepWorkload(k, torch.cuda.Stream(), windows=[13], support_k=11, row_mode=row_mode)
    try:
        for i in range(2):   # the second step folds into the first step's accumulator
            snap = wl.snapshot_inputs()
            wl.step(upload=(i == 0))
            got = wl.snapshot_results()
            exp = step_ref.cyclefold_step(snap)
            rep = step_ref.compare_cyclefold(got, exp)
            assert rep["ok"], rep["bad"]
            assert len(got["poly_F"]) == 16 and len(got["poly_G"]) == 8 and len(got["poly_K"]) == 256
    finally:
        wl.close()


def test_in_library_multi_gpu_commit(oracle):
    """Multi-GPU behind the ABI (SURVEY 8b `sb_init(devs, n)`): one process, `sb_init_devices`, then the PLAIN host entry points
    `sb_msm` / `sb_msm_batch` -- what the unchanged Rust `CommitmentKey::commit` binds -- shard every commit over the devices.
    Lengths around the block-cyclic boundaries, prefixes of the key, batches, both curves; then back to one device."""
    import ctypes

    import torch

    import oracle as O
    import sirius_b200
    from sirius_b200 import _lib

    G = torch.cuda.device_count()
    if G < 2:
        pytest.skip("needs >= 2 GPUs on the box (run with gpurun --gpus 2)")
    lib = _lib.load()
    devs = (ctypes.c_int * G)(*range(G))
    _lib.check(lib.sb_init_devices(devs, G))
    try:
        assert lib.sb_num_devices() == G
        for curve in (0, 1):
            n = 3 * 4096 * G + 4096 + 77          # several full cycles + a partial one
            bases = O.running_bases(curve, n)
            ck = sirius_b200.CommitmentKey(curve, bases)
            for m in (n, n - 1, 4096 * G, 4096 * G + 1, 4095, 1, 0, 2 * 4096 * G - 5):
                s = O.random_field(curve, 31 + m, m)
                assert np.array_equal(ck.commit(s), O.msm(curve, s, bases)), (curve, m)
            vs = [O.random_field(curve, 900 + j, 4096 * G + 9) for j in range(3)]
            got = ck.commit_batch(vs)
            for j, v in enumerate(vs):
                assert np.array_equal(got[j], O.msm(curve, v, bases)), (curve, "batch", j)
            with pytest.raises(sirius_b200.TooLongInput):
                ck.commit(O.random_field(curve, 1, n + 1))
            ck.add_window(9)
            s = O.random_field(curve, 5, 5000)
            assert np.array_equal(ck.commit(s), O.msm(curve, s, bases))
            ck.close()
    finally:
        one = (ctypes.c_int * 1)(0)
        _lib.check(lib.sb_init_devices(one, 1))
    bases = O.running_bases(0, 1000)
    ck = sirius_b200.CommitmentKey(0, bases)
    s = O.random_field(0, 3, 1000)
    assert np.array_equal(ck.commit(s), O.msm(0, s, bases))
    ck.close()


@pytest.mark.parametrize("jit", [1, 0])
def test_cross_terms_row_blocks_equal_full_call(jit):
    """sb_upload_rows_device + sb_cross_terms_rows_device over uneven row blocks == sb_cross_terms_device on the resident
    witness (compiled and interpreted kernels), and an expression that queries a rotated row is refused."""
    import ctypes

    import torch

    from sirius_b200 import _lib
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG
    from sirius_b200 import workload as WL

    lib = _lib.load()
    lib.sb_expr_jit_enable(jit)
    st = torch.cuda.Stream()
    try:
        sess, ex = WL.build_sangria_side(WL.PRIMARY, 11, 0, 1, st, [10], WL.SEED)
        n, A, d = sess.n, sess.A, sess.d
        c1, c2 = sess.challenge_vectors(ex["c1"], ex["u1"], ex["c2"])
        args = (sess.S._hom_prog._h, d, sess.S._cols, sess._cols(sess.W_acc), sess._cols(sess.W_in), A, c1.ctypes.data_as(_lib.u64p),
                c2.ctypes.data_as(_lib.u64p), c1.shape[0])
        _lib.check(lib.sb_cross_terms_device(*args, ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
        st.synchronize()
        full = sess.T.cpu().numpy().copy()
        with torch.cuda.stream(st):
            sess.T.fill_(-1)
            sess.W_in.zero_()
        st.synchronize()
        for r0, cnt in ((0, 1), (1, 700), (701, n - 701 - 3), (n - 3, 3), (5, 0)):
            _lib.check(lib.sb_upload_rows_device(ctypes.c_void_p(ex["host_W"].data_ptr()), A, n, r0, cnt, ctypes.c_void_p(sess.W_in.data_ptr()),
                                                 ctypes.c_void_p(st.cuda_stream)))
            _lib.check(lib.sb_cross_terms_rows_device(*args, r0, cnt, ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
        st.synchronize()
        assert np.array_equal(sess.W_in.cpu().numpy(), ex["host_W"].numpy())
        assert np.array_equal(sess.T.cpu().numpy(), full)
        assert lib.sb_cross_terms_rows_device(*args, n - 2, 3, ctypes.c_void_p(sess.T.data_ptr()), ctypes.c_void_p(st.cuda_stream)) == _lib.SB_ERR_ARG
        assert lib.sb_upload_rows_device(ctypes.c_void_p(ex["host_W"].data_ptr()), A, n, n, 1, ctypes.c_void_p(sess.W_in.data_ptr()),
                                         ctypes.c_void_p(st.cuda_stream)) == _lib.SB_ERR_ARG
        # a rotated query: row ranges are refused (the full call accepts it)
        rot = SG.Program(sess.S.field, P.GraphEvaluator.new(P.Expression.Polynomial(ex["nfix"], 1), sess.S.modulus))
        assert lib.sb_cross_terms_rows_device(rot._h, 1, sess.S._cols, sess._cols(sess.W_acc), sess._cols(sess.W_in), A, c1.ctypes.data_as(_lib.u64p),
                                              c2.ctypes.data_as(_lib.u64p), c1.shape[0], 0, 8, ctypes.c_void_p(sess.T.data_ptr()),
                                              ctypes.c_void_p(st.cuda_stream)) == _lib.SB_ERR_ARG
        assert b"rotation" in lib.sb_last_error()
        sess.S.close()
        sess.ck.close()
    finally:
        lib.sb_expr_jit_enable(1)
        lib.sb_stream_release(ctypes.c_void_p(st.cuda_stream))


def test_gate_scaling_sangria_step(oracle):
    """BASELINE config 5, Sangria arm at a small size: N = 3 sub-circuits (4 gates, folding degree 8 -> 8 cross terms) and
    N = 11 (degree 16, ~670 calculations: the compiled kernel that CALLS its products) against the CPU restatement."""
    import torch

    from oracle import step_ref
    from sirius_b200 import workload as WL

    for N, k in ((3, 9), (11, 7)):
        wl = WL.SangriaStepWorkload(k, 0, 1, torch.cuda.Stream(), windows=[10, 8], primary=WL.gate_scaling_side(N))
        try:
            assert wl.sides[0].d == 5 + N and wl.sides[0].A == 7 + 5 * N
            for i in range(2):
                snap = wl.snapshot_inputs()
                wl.step(upload=(i == 0))
                rep = step_ref.compare(wl.snapshot_results(), step_ref.fold_step(snap, step_ref.bases_for(snap)))
                assert rep["ok"], (N, rep["bad"])
        finally:
            wl.close()


@pytest.mark.parametrize("row_mode", [0, 1])
def test_gate_scaling_protogalaxy_prove(oracle, row_mode):
    """BASELINE config 5, Cyclefold arm: ProtoGalaxy::prove over N = 3 sub-circuits (4 gates -> 2^(k+2) leaves) == oracle."""
    import torch

    from oracle import step_ref
    from sirius_b200 import workload as WL

    wl = WL.GateScalingPgWorkload(8, 3, torch.cuda.Stream(), row_mode=row_mode)
    try:
        assert wl.pg.t == 8 + 2
        snap = wl.snapshot_inputs()
        wl.step()
        rep = step_ref.compare_protogalaxy(wl.snapshot_results(), step_ref.protogalaxy_prove(snap, wl.side))
        assert rep["ok"], rep["bad"]
    finally:
        wl.close()


def test_captured_commit_pipelines_replay_on_fresh_data(oracle):
    """The CUDA-graph cache of msm.cu: the same commitment signature (key, device buffers, stream) is run eagerly, captured on the
    second call and replayed afterwards.  The scalars are rewritten in place between the calls, single and batched commits
    alternate on one stream, a profiling run and sb_msm_tune drop back to eager launches: every result must equal the oracle's,
    and the launch counter must advance by the same amount whether a pipeline is launched or replayed."""
    import ctypes

    import torch

    import oracle as O
    import sirius_b200
    from sirius_b200 import _lib

    lib = _lib.load()
    n, batch = 5000, 3
    st = torch.cuda.Stream()
    for curve in (0, 1):
        bases = O.running_bases(curve, n)
        ck = sirius_b200.CommitmentKey(curve, bases, window_bits=9)
        with torch.cuda.stream(st):
            d_s = torch.zeros((batch * n, 4), dtype=torch.int64, device="cuda")
            d_o = torch.zeros((batch, 8), dtype=torch.int64, device="cuda")
            d_o1 = torch.zeros(8, dtype=torch.int64, device="cuda")
        per_call = []
        for it in range(6):
            if it == 4:
                lib.sb_profile_enable(1)      # event records between the kernels: eager again
            if it == 5:
                lib.sb_profile_enable(0)
                _lib.check(lib.sb_msm_tune(2, 1))   # another sort path: the captured pipelines are stale
            vs = [O.random_field(curve, 100 * it + j + curve, n) for j in range(batch)]
            with torch.cuda.stream(st):
                d_s.copy_(torch.from_numpy(np.concatenate(vs).view(np.int64)), non_blocking=False)
            l0 = lib.sb_launch_count()
            ck.commit_batch_device(d_s.data_ptr(), n, n, batch, d_o.data_ptr(), 0, st.cuda_stream)
            l1 = lib.sb_launch_count()
            ck.commit_device(d_s.data_ptr() + 32 * n, n - 7, d_o1.data_ptr(), 0, st.cuda_stream)
            st.synchronize()
            per_call.append(l1 - l0)
            got = d_o.cpu().numpy().view(np.uint64)
            for j in range(batch):
                assert np.array_equal(got[j], O.msm(curve, vs[j], bases)), (curve, it, j)
            assert np.array_equal(d_o1.cpu().numpy().view(np.uint64), O.msm(curve, vs[1][: n - 7], bases)), (curve, it)
        assert len(set(per_call[:5])) == 1 and per_call[0] > 5, per_call
        _lib.check(lib.sb_msm_tune(2, 0))
        ck.close()
    lib.sb_stream_release(ctypes.c_void_p(st.cuda_stream))
