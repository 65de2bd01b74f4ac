import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
rows = lm.dispersed_params(B).rows(B).cuda()
for dcost in (1e-5, 0.0):
    for mu_ref in (1e-2, 3e-3, 1e-3, 3e-4, 1e-4):
        solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(mu_ref=mu_ref, dcost=dcost), device=0)
        best = 1e9
        for rep in range(3):
            raw = solver.solve_rows(rows, trajectories=False); torch.cuda.synchronize()
            best = min(best, solver.last_kernel_ms())
        print(f'dcost {dcost:g} mu_ref {mu_ref:g}: {best:.1f} ms iters {raw["iterations"].double().mean():.2f} max {raw["iterations"].max().item()} fails {(raw["status"]!=0).sum().item()}')
