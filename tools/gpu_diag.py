import sys, time, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
rows = lm.dispersed_params(B).rows(B).cuda()
def run(label, fn, n=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    print(f'{label}: {(time.perf_counter()-t0)/n*1e3:.1f} ms/step (kernel {solver.last_kernel_ms():.1f})')
run('alloc each call, traj', lambda: solver.solve_rows(rows, True))
out = solver.alloc_outputs(B, True, True)
run('reuse out, traj', lambda: solver.solve_rows(rows, True, out=out))
run('alloc each call, no traj', lambda: solver.solve_rows(rows, False))
rows_h = rows.cpu().pin_memory()
out_h = solver.alloc_outputs(B, True, False)
run('host API reuse out, traj', lambda: solver.solve_rows(rows_h, True, out=out_h))
run('host API alloc, traj', lambda: solver.solve_rows(rows_h, True))
run('host API reuse, no traj', lambda: solver.solve_rows(rows_h, False, out=out_h))
