"""Kernel times of the cooperative kernel on three batch shapes, for A/B runs of differently built libraries
(tools/gpu_variants.sh).  Developer script (GPU box)."""
import sys
import torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm

dev = torch.device("cuda", 0)
for B, nt, reps in ((1, 200, 3), (1184, 200, 3), (4096, 200, 3), (4096, 2001, 2)):
    s = lm.AscentSolver(lm.Mesh(nt=nt), lm.SolverOptions(kernel="coop", dcost=1e-5), device=dev)
    rows = (lm.dispersed_params(B, seed=11) if B > 1 else lm.AscentParams()).rows(B, device=dev)
    ms = []
    for _ in range(reps):
        r = s.solve_rows(rows, trajectories=False)
        ms.append(s.last_kernel_ms())
    print(f"  B={B} nt={nt}: {min(ms):.2f} ms, converged {int((r['status'] == 0).sum())}/{B}, iters {float(r['iterations'].float().mean()):.1f}", flush=True)
    s.close()
