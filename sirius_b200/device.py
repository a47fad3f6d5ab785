"""Device-resident Sangria prover state: the accumulator witness W, the error vector E, the fixed
columns and the commitment key stay in HBM across fold steps; per step only the fresh witness columns go host ->
device and the 64-byte commitments come back (each must reach the host before the next challenge can be
squeezed -- src/plonk/mod.rs:537-545, src/nifs/sangria/mod.rs:171-178).

torch is used here for device memory, pinned host buffers and streams only; every computation is a call
through the C ABI (`*_device` entry points).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import numpy as np

from . import _lib
from .commitment import CommitmentKey
from .sangria import PlonkStructure, _to_mont


def _vp_array(ptrs: List[int]):
    return (ctypes.c_void_p * max(1, len(ptrs)))(*ptrs)


class DeviceSangriaSide:
    """One curve of the cycle: relaxed accumulator (W, E) + the incoming trace's W, single witness round."""

    def __init__(self, S: PlonkStructure, ck: CommitmentKey, stream):
        import torch

        self.torch = torch
        self.S, self.ck, self.stream = S, ck, stream
        self.n = 1 << S.k
        self.A = S.num_advice_columns
        self.d = S.degree
        dev = torch.device("cuda", torch.cuda.current_device())
        i64 = torch.int64
        # allocated AND zero-filled on the session's stream: every later kernel / copy touching them runs there too
        with torch.cuda.stream(stream):
            self.W_acc = torch.zeros((self.A * self.n, 4), dtype=i64, device=dev)
            self.W_new = torch.zeros_like(self.W_acc)
            self.W_in = torch.zeros_like(self.W_acc)
            self.E_acc = torch.zeros((self.n, 4), dtype=i64, device=dev)
            self.E_new = torch.zeros_like(self.E_acc)
            self.T = torch.zeros((self.d, self.n, 4), dtype=i64, device=dev)
            self.commit_W = torch.zeros(8, dtype=i64, device=dev)
            self.commit_T = torch.zeros((self.d, 8), dtype=i64, device=dev)
        stream.synchronize()
        self.h_commit_W = torch.zeros(8, dtype=i64).pin_memory()
        self.h_commit_T = torch.zeros((self.d, 8), dtype=i64).pin_memory()
        self.one = _to_mont([1], S.modulus)
        self._ch_cache = {}   # id-keyed cache of the concatenated challenge vectors (host-side glue, built once)

    # ---- data movement
    def upload_incoming(self, host_W_pinned) -> int:
        """H2D of the fresh witness round (pinned host tensor int64 [A*n,4]); returns bytes copied."""
        with self.torch.cuda.stream(self.stream):
            self.W_in.copy_(host_W_pinned, non_blocking=True)
        return host_W_pinned.numel() * 8

    def _cols(self, t):
        base = t.data_ptr()
        hit = getattr(self, "_cols_cache", None)
        if hit is None:
            hit = self._cols_cache = {}
        arr = hit.get(base)
        if arr is None:
            arr = hit[base] = _vp_array([base + j * self.n * 32 for j in range(self.A)])
        return arr

    # ---- PlonkStructure::run_sps_protocol's commit (src/plonk/mod.rs:441-445)
    def commit_incoming(self) -> np.ndarray:
        st = self.stream.cuda_stream
        self.ck.commit_device(self.W_in.data_ptr(), self.A * self.n, self.commit_W.data_ptr(), 0, st)
        with self.torch.cuda.stream(self.stream):
            self.h_commit_W.copy_(self.commit_W, non_blocking=True)
        self.stream.synchronize()  # the commitment is absorbed by the host-side random oracle
        return self.h_commit_W.numpy().view(np.uint64).copy()

    # ---- VanillaFS::prove (src/nifs/sangria/mod.rs:253-277) split at the challenge
    def commit_cross_terms(self, U1_challenges: np.ndarray, U1_u: np.ndarray, U2_challenges: np.ndarray) -> np.ndarray:
        lib = _lib.load()
        st = self.stream.cuda_stream
        c1, c2 = self.challenge_vectors(U1_challenges, U1_u, U2_challenges)
        _lib.check(
            lib.sb_cross_terms_device(
                self.S._hom_prog._h, self.d, self.S._cols, self._cols(self.W_acc), self._cols(self.W_in), self.A,
                c1.ctypes.data_as(_lib.u64p), c2.ctypes.data_as(_lib.u64p), c1.shape[0], ctypes.c_void_p(self.T.data_ptr()), ctypes.c_void_p(st),
            )
        )
        self.ck.commit_batch_device(self.T.data_ptr(), self.n, self.n, self.d, self.commit_T.data_ptr(), 0, st)
        with self.torch.cuda.stream(self.stream):
            self.h_commit_T.copy_(self.commit_T, non_blocking=True)
        self.stream.synchronize()  # cross-term commitments feed generate_challenge (:162-179)
        return self.h_commit_T.numpy().view(np.uint64).copy()

    def challenge_vectors(self, U1_challenges, U1_u, U2_challenges):
        """[U1.challenges.., U1.u] and [U2.challenges.., 1] (src/nifs/sangria/mod.rs:113-118)."""
        key = (U1_challenges.tobytes(), U1_u.tobytes(), U2_challenges.tobytes())
        hit = self._ch_cache.get(key)
        if hit is None:
            c1 = np.ascontiguousarray(np.concatenate([U1_challenges.reshape(-1, 4), U1_u.reshape(1, 4)]), dtype=np.uint64)
            c2 = np.ascontiguousarray(np.concatenate([U2_challenges.reshape(-1, 4), self.one]), dtype=np.uint64)
            if len(self._ch_cache) > 64:
                self._ch_cache.clear()
            hit = self._ch_cache[key] = (c1, c2)
        return hit

    def fold(self, r: np.ndarray) -> None:
        """W <- W + r*W_in ; E <- E + sum r^j T_j   (accumulator.rs:363-404); buffers are swapped, not copied."""
        lib = _lib.load()
        st = self.stream.cuda_stream
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        f = self.S.field
        _lib.check(lib.sb_axpy_fold_device(f, ctypes.c_void_p(self.W_acc.data_ptr()), ctypes.c_void_p(self.W_in.data_ptr()), r.ctypes.data_as(_lib.u64p),
                                           ctypes.c_void_p(self.W_new.data_ptr()), self.A * self.n, ctypes.c_void_p(st)))
        _lib.check(lib.sb_error_fold_device(f, ctypes.c_void_p(self.E_acc.data_ptr()), ctypes.c_void_p(self.T.data_ptr()), self.d, r.ctypes.data_as(_lib.u64p),
                                            ctypes.c_void_p(self.E_new.data_ptr()), self.n, ctypes.c_void_p(st)))
        self.W_acc, self.W_new = self.W_new, self.W_acc
        self.E_acc, self.E_new = self.E_new, self.E_acc


def synthetic_key(curve: int, n: int, window_bits: int = 0, stream=None) -> CommitmentKey:
    """ck[i] = [i+1]G generated on the device (BASELINE.md section 3 synthetic bases)."""
    import torch

    from .curves import generator_limbs

    lib = _lib.load()
    d = torch.empty((n, 8), dtype=torch.int64, device="cuda")   # fully written by the kernel below (no fill on another stream)
    torch.cuda.synchronize()
    g = generator_limbs(curve)
    st = stream.cuda_stream if stream is not None else 0
    _lib.check(lib.sb_index_multiples_device(curve, g.ctypes.data_as(_lib.u64p), 0, n, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(st or None)))
    torch.cuda.synchronize()
    ck = CommitmentKey.from_device(curve, d.data_ptr(), n, window_bits=window_bits, stream=st)
    torch.cuda.synchronize()
    del d
    return ck


def random_field_device(n: int, seed: int):
    """n synthetic field elements directly in HBM: 252-bit uniform integers, read as Montgomery residues (every value
    below the modulus is the Montgomery form of exactly one element, so this is a uniform draw over a 2^252 subset)."""
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.randint(-(2**63), 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= 0x0FFFFFFFFFFFFFFF
    return t   # produced on torch's CURRENT stream: call inside `with torch.cuda.stream(s)` or synchronise before cross-stream use


def ints_to_mont(vals, modulus: int) -> np.ndarray:
    """Python ints -> uint64 [len,4] Montgomery limbs (host glue; one bytes join instead of per-limb Python loops)."""
    if not len(vals):
        return np.zeros((0, 4), dtype=np.uint64)
    buf = b"".join((((int(v) % modulus) << 256) % modulus).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def mont_to_ints(arr: np.ndarray, modulus: int):
    rinv = pow(1 << 256, -1, modulus)
    raw = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") * rinv % modulus for i in range(0, len(raw), 32)]


class DeviceProtogalaxySide:
    """Device-resident Protogalaxy prover state (src/nifs/protogalaxy): accumulator witness + one incoming trace (L = 1),
    single witness round, bn256 Fr.  compute_F / compute_G leaf evaluation and beta trees, fold_witness and the trace
    commitment run on the device; the O(#points) scalar glue (Lagrange values, ifft of 32 / 8 points, K on 256 points)
    is host integer arithmetic, as it stays host Rust in the integration."""

    def __init__(self, S: PlonkStructure, ck: CommitmentKey, stream, row_mode: int = 1):
        import torch

        from . import fft, protogalaxy as PGX

        self.torch, self.S, self.ck, self.stream, self.row_mode = torch, S, ck, stream, row_mode
        self.PGX, self.fft = PGX, fft
        self.M = S.modulus
        self.n, self.A = 1 << S.k, S.num_advice_columns
        self.ctx = PGX.PolyContext(S, 1)
        self.t = self.ctx.betas_count()
        self.nF, self.nG = self.ctx.fft_points_count_F(), self.ctx.fft_points_count_G
        self.leaves_n = 1 << self.t
        dev = torch.device("cuda", torch.cuda.current_device())
        i64 = torch.int64
        with torch.cuda.stream(stream):
            self.W_acc = torch.zeros((self.A * self.n, 4), dtype=i64, device=dev)
            self.W_in = torch.zeros_like(self.W_acc)
            self.W_new = torch.zeros_like(self.W_acc)
            self.leaves = torch.zeros((self.nG, self.leaves_n, 4), dtype=i64, device=dev)
            self.d_out = torch.zeros((max(self.nF, self.nG), 4), dtype=i64, device=dev)
            self.commit_W = torch.zeros(8, dtype=i64, device=dev)
        self.h_out = torch.zeros((max(self.nF, self.nG), 4), dtype=i64).pin_memory()
        self.h_commit_W = torch.zeros(8, dtype=i64).pin_memory()
        stream.synchronize()
        progs = S.gate_programs()
        self._gates = (ctypes.c_void_p * len(progs))(*[p._h for p in progs])
        self._ng = len(progs)
        self._XsF = PGX.iter_cyclic_subgroup(self.nF.bit_length() - 1)
        self._XsG = PGX.iter_cyclic_subgroup(self.ctx.fft_log_domain_size_G())[: self.nG]
        self._LsG = [PGX.eval_lagrange_polys(X, self.ctx.lagrange_domain()) for X in self._XsG]
        self._coefG = ints_to_mont([c for L in self._LsG for c in L[:2]], self.M)
        self._one = ints_to_mont([1], self.M)

    def _cols(self, t):
        base = t.data_ptr()
        return [base + j * self.n * 32 for j in range(self.A)]

    def _tree(self, num_points, leaf_stride, mult):
        lib = _lib.load()
        st = self.stream.cuda_stream
        m = ints_to_mont(mult, self.M)
        _lib.check(lib.sb_beta_tree_device(self.S.field, ctypes.c_void_p(self.leaves.data_ptr()), self.t, num_points, leaf_stride,
                                           m.ctypes.data_as(_lib.u64p), ctypes.c_void_p(self.d_out.data_ptr()), ctypes.c_void_p(st)))
        with self.torch.cuda.stream(self.stream):
            self.h_out[:num_points].copy_(self.d_out[:num_points], non_blocking=True)
        self.stream.synchronize()   # the polynomial is absorbed by the host-side random oracle (mod.rs:416-447)
        return mont_to_ints(self.h_out[:num_points].numpy().view(np.uint64), self.M)

    def compute_F_evals(self, betas, delta):
        """compute_F (poly/mod.rs:68-203) up to the ifft: F on the order-32 subgroup, from the ACCUMULATOR's trace."""
        lib = _lib.load()
        st = self.stream.cuda_stream
        tab = _vp_array(self._cols(self.W_acc))
        _lib.check(lib.sb_pg_leaves_device(self._gates, self._ng, self.S._cols, tab, 1, self.A, self._one.ctypes.data_as(_lib.u64p), None, 0, 1,
                                           self.row_mode, self.t, ctypes.c_void_p(self.leaves.data_ptr()), ctypes.c_void_p(st)))
        M = self.M
        deltas = [delta % M]
        for _ in range(self.t - 1):
            deltas.append(deltas[-1] * deltas[-1] % M)
        mult = [(b + X * d) % M for X in self._XsF for b, d in zip(betas[: self.t], deltas)]
        return self._tree(self.nF, 0, mult)

    def compute_G_evals(self, betas_stroke):
        """compute_G (poly/mod.rs:308-425) up to the ifft: the 8 Lagrange blends of (accumulator, incoming) are evaluated on the
        fly (no FoldedWitness copies, folded_witness.rs:66-143), one beta* tree per blend."""
        lib = _lib.load()
        st = self.stream.cuda_stream
        tab = _vp_array(self._cols(self.W_acc) + self._cols(self.W_in))
        _lib.check(lib.sb_pg_leaves_device(self._gates, self._ng, self.S._cols, tab, 2, self.A, self._coefG.ctypes.data_as(_lib.u64p), None, 0, self.nG,
                                           self.row_mode, self.t, ctypes.c_void_p(self.leaves.data_ptr()), ctypes.c_void_p(st)))
        bs = [b % self.M for b in betas_stroke[: self.t]]
        return self._tree(self.nG, self.leaves_n, bs * self.nG)

    def fold_witness(self, gamma: int) -> None:
        """ProtoGalaxy::fold_witness (mod.rs:176-210): W_acc <- L0(gamma) W_acc + L1(gamma) W_in (buffers swapped, not copied)."""
        lib = _lib.load()
        Lg = self.PGX.eval_lagrange_polys(gamma, self.ctx.lagrange_domain())[:2]
        coef = ints_to_mont(Lg, self.M)
        ins = _vp_array([self.W_acc.data_ptr(), self.W_in.data_ptr()])
        _lib.check(lib.sb_lincomb_device(self.S.field, ins, coef.ctypes.data_as(_lib.u64p), 2, self.A * self.n, ctypes.c_void_p(self.W_new.data_ptr()),
                                         ctypes.c_void_p(self.stream.cuda_stream)))
        self.W_acc, self.W_new = self.W_new, self.W_acc

    def prove(self, betas, delta: int, alpha: int, gamma: int):
        """ProtoGalaxy::prove (mod.rs:400-481) with the random oracle's challenges supplied by the caller.
        Returns (poly_F, poly_G, poly_K) as coefficient lists."""
        PGX = self.PGX
        poly_F = PGX._ifft_ints(self.compute_F_evals(betas, delta))
        bs = PGX.beta_stroke(betas[: self.t], alpha, delta)
        poly_G = PGX._ifft_ints(self.compute_G_evals(bs))
        poly_K = PGX.compute_K_from_G(self.ctx, poly_G, PGX.poly_eval(poly_F, alpha))
        self.fold_witness(gamma)
        return poly_F, poly_G, poly_K

    def upload_incoming(self, host_W_pinned) -> int:
        with self.torch.cuda.stream(self.stream):
            self.W_in.copy_(host_W_pinned, non_blocking=True)
        return host_W_pinned.numel() * 8

    def commit_incoming(self) -> np.ndarray:
        """generate_plonk_trace's commit of the new trace (src/plonk/mod.rs:441-445): the 12 * 2^k-point MSM."""
        st = self.stream.cuda_stream
        self.ck.commit_device(self.W_in.data_ptr(), self.A * self.n, self.commit_W.data_ptr(), 0, st)
        with self.torch.cuda.stream(self.stream):
            self.h_commit_W.copy_(self.commit_W, non_blocking=True)
        self.stream.synchronize()
        return self.h_commit_W.numpy().view(np.uint64).copy()
