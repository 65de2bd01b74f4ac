"""Developer timing: warm vs cold start."""
import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rows = lm.dispersed_params(B).rows(B).cuda()
res = {}
for warm in (False, True):
    solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(warm_start=warm), device=0)
    best = 1e9
    for rep in range(4):
        raw = solver.solve_rows(rows); torch.cuda.synchronize()
        best = min(best, solver.last_kernel_ms())
    it = raw['iterations'].double()
    res[warm] = raw
    print(f'warm={warm}: B {B} kernel ms best {best:.2f} solves/s {B/best*1e3:.0f} fails {(raw["status"]!=0).sum().item()} iters mean {it.mean():.2f} max {it.max():.0f}')
d = (res[True]['tf'] - res[False]['tf']).abs().max().item()
dt = (res[True]['traj'] - res[False]['traj']).abs().amax(dim=(1, 2)) / res[False]['traj'].abs().amax(dim=(1, 2))
print('max |tf_warm - tf_cold|', d, 'max rel traj diff per row', [f'{x:.1e}' for x in dt.tolist()])
