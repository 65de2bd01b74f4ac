"""Developer study: which dispersed parameter predicts the iteration count (to group similar problems in a warp)."""
import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
from lunar_module_ascent_trajectory_optimiser_b200 import _cabi
B = 65536
rows = lm.dispersed_params(B).rows(B).cuda()
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
raw = solver.solve_rows(rows); raw = solver.solve_rows(rows); torch.cuda.synchronize()
it = raw['iterations'].double()
print('iters mean %.2f std %.2f; warp-max mean %.2f' % (it.mean(), it.std(), it.view(-1, 32).max(dim=1).values.mean()))
for name in ['Ft', 'M0', 'M_dot', 'angle_doubledot_max', 'r_periapsis', 'r_apoapsis']:
    x = rows[_cabi.PARAM_ROWS.index(name)]
    c = torch.corrcoef(torch.stack([x, it]))[0, 1].item()
    print(f'{name:22s} corr {c:+.3f}')
tw = rows[3] / rows[4]
print('thrust/weight corr %+.3f' % torch.corrcoef(torch.stack([tw, it]))[0, 1].item())
print('tf corr %+.3f' % torch.corrcoef(torch.stack([raw["tf"], it]))[0, 1].item())
# what grouping by a key would give
for key_name, key in [('addm', rows[7]), ('tf', raw['tf']), ('Ft/M0', tw), ('rp', rows[8])]:
    order = torch.argsort(key)
    wm = it[order].view(-1, 32).max(dim=1).values.mean().item()
    print(f'sorted by {key_name:6s}: warp-max mean {wm:.2f}')
hist = torch.bincount(it.long())
print('histogram', {i: int(h) for i, h in enumerate(hist.tolist()) if h})
