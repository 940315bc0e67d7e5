"""NVLink / NVSwitch peer-memory exchange for the multi-GPU step (include/uqoc.h
``uqoc_su2_fwdbwd_peer``).

The sample-sharded step ends in ONE exchange of ``[Fsum (B) | G (B*L*2)]``.  For MB-sized vectors NCCL's
all-reduce is the right tool (``fused_propagate_loss(group=pg)``); for the small vectors of few-target
workloads (BASELINE config 3: 2 KB) its ~45 us of launch + protocol latency is as long as the whole fused
kernel.  :class:`PeerExchange` maps every rank's exchange buffer into every process
(``torch.distributed._symmetric_memory``: CUDA VMM handles exchanged through the process group's store) and
hands the raw device pointers to ``uqoc_su2_fwdbwd_peer``, whose second kernel reduces the sample-tile
partials, pushes the result into every rank's buffer with plain stores over NVLink, raises per-block flags at
system scope, waits for the same block of every rank and sums the slots in rank order - bit-identical on all
ranks, no NCCL call on the step's critical path.

    px = PeerExchange(pg, B, L, dtype=torch.float32)             # once (collective: all ranks)
    loss, fid = fused_propagate_loss(pulses, U_target, monte_carlo=M, group=px)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

MAX_WORLD = 16
#: above this many exchanged reals the one-shot push (world-1 copies of the vector per rank) loses to NCCL
MAX_N = 1 << 18


def _symmetric_memory():
    """``torch.distributed._symmetric_memory`` is a private module (torch >= 2.5, interface as of 2.11): fail with a
    clear message - callers fall back to the NCCL all-reduce - when it is absent or has changed shape."""
    try:
        import torch.distributed._symmetric_memory as symm_mem
    except ImportError as e:                     # pragma: no cover - depends on the torch build
        raise RuntimeError(f"torch {torch.__version__} has no torch.distributed._symmetric_memory: {e}") from e
    missing = [n for n in ("empty", "rendezvous") if not hasattr(symm_mem, n)]
    if missing:                                  # pragma: no cover
        raise RuntimeError(f"torch {torch.__version__}: torch.distributed._symmetric_memory lacks {missing}; "
                           "PeerExchange needs empty() and rendezvous() -> handle.buffer_ptrs")
    return symm_mem


class PeerExchange:
    def __init__(self, group, B: int, L: int, P: int = 2, dtype: torch.dtype = torch.float32, device=None):
        import torch.distributed as dist
        symm_mem = _symmetric_memory()

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > MAX_WORLD:
            raise ValueError(f"PeerExchange supports at most {MAX_WORLD} ranks, got {self.world}")
        self.n = int(B) + int(B) * int(L) * int(P)
        if self.n > MAX_N:
            raise ValueError(f"exchange vector of {self.n} reals: use the NCCL path (group=process_group) above {MAX_N}")
        self.B, self.L, self.P, self.dtype = int(B), int(L), int(P), dtype
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.lib()
        dt = _lib.F64 if dtype == torch.float64 else _lib.F32
        data_bytes = (int(lib.uqoc_peer_data_bytes(self.n, self.world, dt)) + 255) // 256 * 256
        flag_bytes = int(lib.uqoc_peer_flag_bytes(self.world))
        self.buf = symm_mem.empty(data_bytes + flag_bytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.data_ptrs = (C.c_uint64 * self.world)(*ptrs)
        self.flag_ptrs = (C.c_uint64 * self.world)(*[p + data_bytes for p in ptrs])
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                 # every rank's flags are zero before any peer can raise one
        torch.cuda.synchronize(self.device)
        self.epoch = 0

    def next_epoch(self) -> int:
        # a free-running uint32 (0 skipped: the flags start at 0): the kernels compare flags and epoch by their SIGNED
        # difference, which stays valid across the wrap as long as the ranks are within 2^31 calls of each other
        self.epoch = ((self.epoch + 1) & 0xFFFFFFFF) or 1
        return self.epoch

    def matches(self, B: int, L: int, P: int, dtype: torch.dtype) -> bool:
        return (self.B, self.L, self.P, self.dtype) == (int(B), int(L), int(P), dtype)
