#!/usr/bin/env python
"""Whole training step at the reference's shipped single-qubit config (model_params.json: d_model 512, 8 layers,
16 heads, dropout 0.1, L = 100; SCORE.py:316-328 batch 200; monte_carlo 1000): eager FusedTrainer.train_epoch
vs GraphedTrainStep (one CUDA-graph replay).  The model here is a stand-in of the same size (a stock
nn.TransformerEncoder over 9 tokens + linear head + sigmoid range map), not the reference's class."""
import os, sys, time, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200.trainer import FusedTrainer, GraphedTrainStep, SigmaSpec


class StandIn(torch.nn.Module):
    num_qubits = 1

    def __init__(self, L=100, d=512, layers=8, heads=16, drop=0.1):
        super().__init__()
        self.L = L
        self.inp = torch.nn.Linear(8, d)
        enc = torch.nn.TransformerEncoderLayer(d, heads, 4 * d, drop, batch_first=True)
        self.enc = torch.nn.TransformerEncoder(enc, layers)
        self.head = torch.nn.Linear(9 * d, 2 * L)

    def forward(self, x):                                   # x (B, 9, 8): 9 "SCORE" unitaries as real vectors
        h = self.enc(self.inp(x))
        y = torch.sigmoid(self.head(h.flatten(1))).view(-1, self.L, 2)
        return torch.stack([(y[..., 0] * 2 - 1) * 3.15, 0.1 + 0.4 * y[..., 1]], -1)


B, L, M = 200, 100, 1000
dev = "cuda"
torch.manual_seed(0)
emb = torch.randn(B, 9, 8, device=dev)
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
T = torch.matrix_exp(-1j * X[None] * (torch.rand(B) * math.pi)[:, None, None]).to(dev)
spec = SigmaSpec(0.4, 0.05)
m1, m2 = StandIn(L), StandIn(L)
print("parameters: %.1f M" % (sum(p.numel() for p in m1.parameters()) / 1e6))
tr = FusedTrainer(m1, monte_carlo=M, device=dev)
gs = GraphedTrainStep(m2, B=B, emb_shape=(9, 8), monte_carlo=M, device=dev)


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


ms_e = timed(lambda: tr.train_epoch(emb, T, spec))          # includes loss.item() per step, as trainer.py:94
ms_g = timed(lambda: gs(emb, T, spec).item())
print(f"eager FusedTrainer.train_epoch: {ms_e:.3f} ms/step   GraphedTrainStep: {ms_g:.3f} ms/step   ({ms_e / ms_g:.2f}x)")
