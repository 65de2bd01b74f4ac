"""Kernel time of the thread-per-problem kernel on config 4 (65 536 x nt = 200, DCOST on, batch warm start), for
A/B runs of differently built libraries (tools/gpu_variants.sh PROBE=tools/gpu_variant_thread.py).  Developer script."""
import sys
import torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
s = lm.AscentSolver(lm.Mesh(nt=200), lm.SolverOptions(kernel="thread"), device=dev)
rows = [lm.dispersed_params(B, seed=11 + 1000 * j).rows(B, device=dev) for j in range(3)]
ms = []
for i in range(6):
    r = s.solve_rows(rows[i % 3], trajectories=True)
    ms.append(s.last_kernel_ms())
print(f"  B={B}: kernel ms {['%.2f' % m for m in ms]}  min {min(ms[1:]):.2f}  mean(2..) {sum(ms[1:]) / len(ms[1:]):.2f}  "
      f"converged {int((r['status'] == 0).sum())}/{B} iters {float(r['iterations'].float().mean()):.2f}", flush=True)
s.close()
