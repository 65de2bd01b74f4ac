import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 16384
rows = lm.dispersed_params(B).rows(B).cuda()
for keps in (10.0, 30.0):
    for ws in (0, 1):
        solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(kappa_eps=keps, warm_start=ws), device=0)
        best = 1e9
        for rep in range(3):
            raw = solver.solve_rows(rows); torch.cuda.synchronize(); best = min(best, solver.last_kernel_ms())
        it = raw['iterations'].double()
        print(f'keps {keps} warm {ws}: ms {best:.2f} iters mean {it.mean():.2f} max {it.max():.0f} fails {(raw["status"]!=0).sum().item()}')
