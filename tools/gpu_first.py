import time, torch, sys
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
s = lm.optimise()
print('nominal', s.tf_seconds, s.iterations, s.status, s.kkt_error)
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
print('fp64 peak GF', solver.measure_fp64_peak())
for B in (1, 1024, 8192, 65536):
    p = lm.dispersed_params(B)
    rows = p.rows(B).cuda()
    for rep in range(2):
        raw = solver.solve_rows(rows); torch.cuda.synchronize()
        ms = solver.last_kernel_ms()
    it = raw['iterations'].double()
    print(f'B {B} kernel ms {ms:.2f} solves/s {B/ms*1e3:.0f} fails {(raw["status"]!=0).sum().item()} iters mean {it.mean():.1f} max {it.max():.0f}')
