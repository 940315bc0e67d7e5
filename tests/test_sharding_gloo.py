"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): the sample-axis sharding of
sharding.py + the single all-reduce of [Fsum | G] reproduces the unsharded result.  The CUDA
kernel's place is taken by the CPU oracle here (the product itself has no CPU path); what is
under test is the partition, the Philox counter invariance and the combine step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import uqoc_oracle as orc
from universal_quantum_optimal_control_b200.sharding import shard_errors, shard_range


def test_shard_range_partitions_exactly():
    for M in (2, 3, 7, 256, 1000, 65536, 10 ** 6):
        for world in (1, 2, 3, 4, 8):
            if M < world:
                continue
            spans = [shard_range(M, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(m for _, m in spans) == M
            for (a, ma), (b, _) in zip(spans, spans[1:]):
                assert a + ma == b
            assert max(m for _, m in spans) - min(m for _, m in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(3, 0, 4)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_shard_errors_layout():
    B, M = 3, 10
    e = torch.arange(2 * B * M).reshape(2, B * M)
    j0, m = shard_range(M, 1, 3)
    s = shard_errors(e, B, M, j0, m)
    assert s.shape == (2, B * m)
    for b in range(B):
        assert torch.equal(s[0, b * m:(b + 1) * m], e[0, b * M + j0: b * M + j0 + m])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, philox, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                       # identical pulses/targets on every rank
    B, L, M_total = 3, 12, 50
    pulses = np.stack([rng.uniform(-3, 3, (B, L)), rng.uniform(0.1, 0.5, (B, L))], -1)
    T = orc.batched_unitary_generator(pulses[:, :3], np.zeros((2, B)))
    j0, M = shard_range(M_total, rank, world)
    if philox:
        err = orc.philox_errors(B, M, 0.7, 0.05, seed=42, offset=1, j0=j0)          # counter = global j
    else:
        full = np.stack([rng.normal(0, 1, B * M_total), rng.normal(0, 0.05, B * M_total)])
        err = shard_errors(torch.from_numpy(full), B, M_total, j0, M).numpy()
    Fsum, G, _ = orc.fidelity_sum_and_grad(pulses, T, err, M)                         # stands in for the kernel
    buf = torch.from_numpy(np.concatenate([G.reshape(-1), Fsum]))                     # the product's [G | Fsum] layout
    dist.all_reduce(buf)                                                              # the one exchange step
    Fbar = buf[B * L * 2:].sum().item() / (B * M_total)
    val, dval = orc.loss_and_dloss(Fbar, "sharp")
    grad = (dval / (B * M_total)) * buf[:B * L * 2].numpy().reshape(B, L, 2)
    if rank == 0:
        err_all = orc.philox_errors(B, M_total, 0.7, 0.05, 42, 1) if philox else full
        want_l, want_g, _ = orc.loss_and_grad(pulses, T, err_all, M_total, "sharp")
        ret["dl"] = abs(val - want_l)
        ret["dg"] = float(np.abs(grad - want_g).max() / np.abs(want_g).max())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("philox", [False, True])
def test_sharded_allreduce_equals_unsharded(world, philox):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), philox, ret), nprocs=world, join=True)
    assert ret["dl"] < 1e-12 and ret["dg"] < 1e-12
