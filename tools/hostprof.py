import os, sys, math, torch, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
B, M, L = 200, 1000, 100
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev)
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
T = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).to(dev)
def fused():
    p = pulses.clone().requires_grad_(True)
    loss, _ = uq.fused_propagate_loss(p, T, monte_carlo=M, sigma=(1.0, 0.05), seed=1, offset=0)
    loss.backward()
    return loss
for _ in range(20): fused()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): fused()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host time per step {(t1-t0)/200*1e6:.1f} us, incl. drain {(t2-t0)/200*1e6:.1f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): fused()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
