"""Developer check: wide dispersions on a dense mesh, the circular model and the dcost = 0 kernel; wall time per batch."""
import sys, time, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
def run(label, solver, rows):
    t0 = time.time(); raw = solver.solve_rows(rows.cuda()); torch.cuda.synchronize(); dt = time.time() - t0
    st = raw['status']
    print(f'{label}: wall {dt*1e3:.0f} ms converged {(st == 0).double().mean().item():.4f} statuses {torch.bincount(st.long(), minlength=6).tolist()} iters mean {raw["iterations"].double().mean():.1f} max {raw["iterations"].max().item()}')
nom = lm.AscentParams().rows(1)
B = 8192
base = lm.dispersed_params(B, seed=9).rows(B)
for w in (1.0, 3.0):
    rows = nom + w * (base - nom)
    run(f'nt=801 x{w}', lm.AscentSolver(lm.Mesh(nt=801), lm.SolverOptions(), device=0), rows)
    run(f'dcost=0 x{w}', lm.AscentSolver(lm.Mesh(), lm.SolverOptions(dcost=0.0), device=0), rows)
    cn = lm.AscentParams.circular().rows(1)
    crow = cn + w * (base - nom) * (cn.abs() > 0)
    crow[8] = crow[9] = cn[8] * (1 + 0.1 * w * (2 * torch.rand(B, dtype=torch.float64, generator=torch.Generator().manual_seed(3)) - 1))
    run(f'circular x{w}', lm.AscentSolver(lm.Mesh(), lm.SolverOptions(dcost=0.0), device=0, model='circular'), crow)
