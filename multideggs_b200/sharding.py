"""Edge sharding arithmetic shared by the host mirror, the tests and the docs.  It is exactly what the engine does
on the device side (mdeggs_engine.cu: run_shard_main / run_collective): the unique-edge list is cut into `world`
contiguous blocks of ceil(E / world) edges (network order is preserved, so re-assembly is a concatenation),
every rank fits its block, and one all-gather of equally sized blocks rebuilds the full p-value vector on
every rank before the global per-(percentile, category) adjustment (SURVEY §8 e)."""
from __future__ import annotations

import numpy as np


def edge_block(E: int, rank: int, world: int) -> tuple[int, int, int]:
    """(lo, hi, chunk): rank's half-open edge range and the padded block length used by the all-gather."""
    chunk = (E + world - 1) // world if world > 0 else E
    lo = min(rank * chunk, E)
    hi = min(lo + chunk, E)
    return lo, hi, chunk


def assemble_blocks(blocks: list[np.ndarray], E: int) -> np.ndarray:
    """Concatenate the gathered, padded per-rank blocks and drop the padding."""
    return np.concatenate(blocks)[:E]
