import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(max_iter=400), device=0)
nom = lm.AscentParams().rows(1)
base = lm.dispersed_params(B, seed=7).rows(B)
rows = (nom + 2.0 * (base - nom))
raw = solver.solve_rows(rows.cuda()); torch.cuda.synchronize()
idx = (raw['status'] == 1).nonzero().flatten().tolist()
print('max_iter problems', idx)
for i in idx[:3]:
    print(i, 'kkt', raw['kkt'][i].item(), 'tf', raw['tf'][i].item(), 'params', ['%.10g' % v for v in rows[:, i].tolist()])
it = raw['iterations']
print('iters > 100:', int((it > 100).sum()), ' > 60:', int((it > 60).sum()), 'hist of failing iters', torch.bincount(it[raw['status'] == 2].long()).nonzero().flatten().tolist()[:40])
