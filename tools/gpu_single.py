"""A few solves of one nominal problem (or a small batch) with a chosen kernel: the ncu target for the latency regime."""
import sys
import torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm

kern = sys.argv[1] if len(sys.argv) > 1 else "coop"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda", 0)
s = lm.AscentSolver(lm.Mesh(nt=nt), lm.SolverOptions(kernel=kern), device=dev)
rows = (lm.dispersed_params(B, seed=11) if B > 1 else lm.AscentParams()).rows(B, device=dev)
for _ in range(3):
    r = s.solve_rows(rows, trajectories=True)
    print(kern, B, nt, "ms", s.last_kernel_ms(), "iters", float(r["iterations"].float().mean()), "conv", int((r["status"] == 0).sum()))
