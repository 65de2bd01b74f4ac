"""Developer study: barrier parameter at which the warm start's reference solve stops."""
import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
rows = lm.dispersed_params(B).rows(B).cuda()
for mu_ref in [float(a) for a in sys.argv[1:]] or [1e-2, 1e-3, 1e-4]:
    solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(mu_ref=mu_ref), device=0)
    best = 1e9
    for rep in range(3):
        raw = solver.solve_rows(rows); torch.cuda.synchronize()
        best = min(best, solver.last_kernel_ms())
    it = raw['iterations'].double()
    print(f'mu_ref {mu_ref:.1e}: kernel ms {best:.2f} fails {(raw["status"]!=0).sum().item()} iters mean {it.mean():.2f} max {it.max():.0f}')
