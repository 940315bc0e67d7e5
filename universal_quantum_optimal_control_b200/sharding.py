"""Sample-axis sharding of the Monte-Carlo average across ranks.

Every (target b, sample j) pair is independent through the propagation; the only coupling is
the pooled mean of ``SCORE.py:194`` and the sum over samples of the pulse gradient, both linear.
So rank r of R takes samples j in [j0, j0+M_r) of EVERY target, pulses and targets are
replicated, and ONE all-reduce(SUM) of the buffer ``[Fsum (B) | G (B*L*P)]`` is the whole
exchange; ``uqoc_loss_finalize`` then applies dloss/dFbar / (B*M_total) identically on every rank.
The Philox counter is the GLOBAL sample index, so the error set does not depend on R.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(M_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(j0, M_local): contiguous, balanced to within one sample, covers [0, M_total) exactly."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if M_total < world:
        raise ValueError("monte_carlo must be >= world size")
    base, rem = divmod(M_total, world)
    j0 = rank * base + min(rank, rem)
    return j0, base + (1 if rank < rem else 0)


def shard_errors(error, B: int, M_total: int, j0: int, M: int):
    """Slice a global (E, B*M_total) error tensor (sample s = b*M_total + j) to this rank's
    (E, B*M) shard with the same layout."""
    E = error.shape[0]
    return error.reshape(E, B, M_total)[:, :, j0:j0 + M].reshape(E, B * M)
