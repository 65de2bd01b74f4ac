"""Developer check: convergence over other seeds and over dispersions wider than SURVEY 8(d)."""
import sys, time, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B = 65536
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
for seed in (1, 2, 3):
    rows = lm.dispersed_params(B, seed=seed).rows(B).cuda()
    raw = solver.solve_rows(rows); torch.cuda.synchronize()
    st = raw['status']
    print(f'seed {seed}: ms {solver.last_kernel_ms():.1f} converged {(st == 0).double().mean().item():.6f} statuses {torch.bincount(st.long()).tolist()} iters mean {raw["iterations"].double().mean():.2f} max {raw["iterations"].max().item()}')
# wider: scale every deviation from nominal by w
nom = lm.AscentParams().rows(1)
base = lm.dispersed_params(B, seed=7).rows(B)
for w in (2.0, 4.0):
    rows = (nom + w * (base - nom)).cuda()
    t0 = time.time(); raw = solver.solve_rows(rows); torch.cuda.synchronize(); dt = time.time() - t0
    st = raw['status']
    print(f'width x{w}: wall {dt*1e3:.0f} ms converged {(st == 0).double().mean().item():.6f} statuses {torch.bincount(st.long(), minlength=5).tolist()} iters mean {raw["iterations"].double().mean():.2f} max {raw["iterations"].max().item()}')
