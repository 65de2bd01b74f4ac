"""Developer A/B timing of one library build: B dispersions, kernel ms (CUDA events); later calls on one handle.
usage: gpu_ab.py [B] [nt] [kappa_eps]   (LMATO_LIB_OVERRIDE selects another build)"""
import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 200
keps = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
solver = lm.AscentSolver(lm.Mesh(nt=nt), lm.SolverOptions(kappa_eps=keps), device=0)
rows = lm.dispersed_params(B).rows(B).cuda()
best = 1e9
for rep in range(4):
    raw = solver.solve_rows(rows); torch.cuda.synchronize()
    ms = solver.last_kernel_ms(); best = min(best, ms)
it = raw['iterations'].double()
print(f'B {B} nt {nt} kernel ms best {best:.2f} solves/s {B/best*1e3:.0f} fails {(raw["status"]!=0).sum().item()} iters mean {it.mean():.2f} max {it.max():.0f} tf0 {raw["tf"][0].item()*470:.8f}')
