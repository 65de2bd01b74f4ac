import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
for ws in (2,):
    solver = lm.AscentSolver(lm.Mesh(nt=nt), lm.SolverOptions(warm_start=ws), device=0)
    for B in (512, 1024, 2048, 3072, 4096, 4736, 8192):
        rows = lm.dispersed_params(B, seed=11).rows(B).cuda()
        for rep in range(2):
            raw = solver.solve_rows(rows, trajectories=False); torch.cuda.synchronize()
        it = raw['iterations'].double()
        print(f'nt {nt} B {B} ({B//32} chunks): kernel {solver.last_kernel_ms():.0f} ms, iters mean {it.mean():.1f} max {it.max():.0f}')
