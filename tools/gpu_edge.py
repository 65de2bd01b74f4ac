"""Developer check: pathological inputs must come back as statuses quickly, never hang."""
import sys, time, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
from lunar_module_ascent_trajectory_optimiser_b200 import _cabi
solver = lm.AscentSolver(lm.Mesh(nt=200), lm.SolverOptions(), device=0)
base = lm.dispersed_params(64, seed=3).rows(64)
names = _cabi.PARAM_ROWS
cases = {
    'nan thrust': ('Ft', float('nan')), 'inf thrust': ('Ft', float('inf')), 'zero thrust': ('Ft', 0.0),
    'negative thrust': ('Ft', -15346.0), 'tiny thrust': ('Ft', 1500.0), 'huge thrust': ('Ft', 1.5e6),
    'zero mdot': ('M_dot', 0.0), 'nan mass': ('M0', float('nan')), 'zero fuel': ('fuel_mass', 0.0),
    'zero accel limit': ('angle_doubledot_max', 0.0), 'perilune 0': ('r_periapsis', 0.0),
    'apolune < perilune': ('r_apoapsis', 1000.0), 'final_time 0': ('final_time', 0.0),
    'final_time tiny': ('final_time', 10.0),
}
for label, (name, val) in cases.items():
    rows = base.clone()
    rows[names.index(name), 5] = val          # one bad problem among 64 good ones
    t0 = time.time()
    raw = solver.solve_rows(rows.cuda()); torch.cuda.synchronize()
    dt = time.time() - t0
    st = raw['status'].cpu()
    others_ok = int((st[torch.arange(64) != 5] != 0).sum()) == 0
    print(f'{label:22s} status {int(st[5])} iters {int(raw["iterations"][5])} tf {float(raw["tf"][5]):.4f} others_ok {others_ok} wall {dt*1e3:.0f} ms')
