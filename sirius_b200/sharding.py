"""Row sharding of the Sangria hot path across ranks.

Rank g of G owns rows [g*n/G, (g+1)*n/G) of every column.  A witness round vector is column-major
(`W[col*n + row]`, src/util/mod.rs:214-218), so the rank's scalars are `num_cols` separate segments of the flat
vector, and the matching commitment-key entries are `ck[col*n + row]` for the same (col, row).  Laid out
column-major again, the rank's first n/G key entries are `ck[row]`, i.e. the prefix that the cross-term commits
(`ck[..2^k]`) need -- so ONE local key serves both `commit(W)` and `commit(T_j)`.

Each rank's partial commitment is a point; the full commitment is their sum (one all-gather of 128-byte XYZZ
partials + a combine kernel on the GPU path).  Per-row kernels (cross terms with rotation 0, folds) need no exchange.
"""
from __future__ import annotations

from typing import List, Tuple


def row_slice(rank: int, world: int, n: int) -> Tuple[int, int]:
    if n % world:
        raise ValueError(f"{n} rows do not divide over {world} ranks")
    per = n // world
    return rank * per, per


def key_segments(num_cols: int, n: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """(first flat index, count) of the global key / witness entries this rank owns, in local column-major order."""
    row0, per = row_slice(rank, world, n)
    return [(col * n + row0, per) for col in range(num_cols)]


def shard_column_major(flat, num_cols: int, n: int, rank: int, world: int):
    """The rank's rows of a column-major [num_cols * n, ...] array, as a local column-major array."""
    import numpy as np

    parts = [flat[first:first + count] for first, count in key_segments(num_cols, n, rank, world)]
    return np.ascontiguousarray(np.concatenate(parts))


def check_rotations_row_local(rotations) -> None:
    """Row sharding without halo rows is only valid for programs that read the current row."""
    bad = [r for r in rotations if r != 0]
    if bad:
        raise ValueError(f"row-sharded evaluation needs rotation 0 only, program uses {bad}")
