"""A small pulse generator shared by ``make_golden.py`` (where the UNMODIFIED reference trainer steps it on the CPU) and
the GPU trajectory test (where ``FusedTrainer`` steps the same weights through the fused op).  It is NOT reference
code: ``UniversalModelTrainer`` only needs ``model(U_emb) -> (B, L, 2)`` and ``model.num_qubits``
(model/universal_model_trainer.py:74,88), so any module with that interface exercises trainer.py:58-94 unchanged."""
import torch
import torch.nn as nn


class TinyPulseModel(nn.Module):
    num_qubits = 1

    def __init__(self, L: int = 16, hidden: int = 32, ranges=((-3.15, 3.15), (0.1, 0.5))):
        super().__init__()
        self.L = L
        self.net = nn.Sequential(nn.Linear(4, hidden), nn.Tanh(), nn.Linear(hidden, L * 2))
        self.register_buffer("lo", torch.tensor([r[0] for r in ranges]))
        self.register_buffer("hi", torch.tensor([r[1] for r in ranges]))

    def forward(self, rotation_vector: torch.Tensor) -> torch.Tensor:
        u = self.net(rotation_vector).view(rotation_vector.shape[0], self.L, 2).sigmoid()
        return self.lo + (self.hi - self.lo) * u
