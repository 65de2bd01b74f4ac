import sys, torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm
dev = torch.device("cuda", 0)
B = 65536
s = lm.AscentSolver(lm.Mesh(nt=200), lm.SolverOptions(kernel="thread", dcost=0.0), device=dev)
rows = [lm.dispersed_params(B, seed=11 + 1000 * j).rows(B, device=dev) for j in range(3)]
ms = []
for i in range(6):
    r = s.solve_rows(rows[i % 3], trajectories=True); ms.append(s.last_kernel_ms())
print(f"  dcost=0 B={B}: kernel ms {['%.2f' % m for m in ms]} mean(2..) {sum(ms[1:]) / 5:.2f} converged {int((r['status'] == 0).sum())} iters {float(r['iterations'].float().mean()):.2f}")
