import sys, dataclasses, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
# batch of dispersed circular problems with and without the move term, both lane counts + auto at a large batch
nom = lm.AscentParams.circular()
B = 2000
base = lm.dispersed_params(B, seed=9).rows(B); en = lm.AscentParams().rows(1); cn = nom.rows(1)
crow = cn + (base - en) * (cn.abs() > 0)
crow[8] = crow[9] = cn[8] * (1 + 0.1 * (2 * torch.rand(B, dtype=torch.float64, generator=torch.Generator().manual_seed(3)) - 1))
out = {}
for dc in (0.0, 1e-5):
    for lanes in (8, 32):
        s = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(dcost=dc, coop_lanes=lanes, kernel="coop"), device=0, model="circular")
        r = s.solve_rows(crow.cuda()); torch.cuda.synchronize()
        out[(dc, lanes)] = {k: v.clone().cpu() for k, v in r.items() if v is not None}
        print(f"dcost {dc} lanes {lanes}: {s.last_kernel_ms():.1f} ms converged {int((r['status'] == 0).sum())}/{B} iters {r['iterations'].double().mean():.1f} kkt max {r['kkt'].max().item():.1e}")
        s.close()
a, b = out[(1e-5, 8)], out[(1e-5, 32)]
print("lanes 8 vs 32 with term: max rel dtf", ((a['tf'] - b['tf']).abs() / a['tf']).max().item())
a0 = out[(0.0, 8)]
print("term vs no term: max rel dtf", ((a['tf'] - a0['tf']).abs() / a0['tf']).max().item(), "max angle diff", (a['traj'][6] - a0['traj'][6]).abs().max().item())
s = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0, model="circular")
big = crow.repeat(1, 5)
r = s.solve_rows(big.cuda()); torch.cuda.synchronize()
print(f"auto kernel, B={big.shape[1]} with term: {s.last_kernel_ms():.1f} ms converged {int((r['status'] == 0).sum())}")
