import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
nom = lm.AscentParams().rows(1)
base = lm.dispersed_params(B, seed=7).rows(B)
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
for w in (2.0, 4.0):
    raw = solver.solve_rows((nom + w * (base - nom)).cuda()); torch.cuda.synchronize()
    it, st = raw['iterations'], raw['status']
    c = it[st == 0]
    print(f'x{w}: converged {len(c)} iters max {c.max().item()} p99.9 {torch.quantile(c.double(), 0.999).item():.0f}; >60: {(c > 60).sum().item()} >100: {(c > 100).sum().item()}; stalled iters: {sorted(it[st == 5].tolist())[-5:]}')
