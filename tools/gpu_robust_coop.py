"""Developer check: the cooperative kernel (both lane counts) against the thread-per-problem kernel on dispersions 1x, 2x
and 4x as wide as SURVEY 8(d): same statuses problem by problem, same tf on the converged ones; the circular model and
a dense mesh as well."""
import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm

def solve(rows, kernel, lanes=0, nt=200, model='elliptical', dcost=1e-5):
    s = lm.AscentSolver(lm.Mesh(nt=nt), lm.SolverOptions(kernel=kernel, coop_lanes=lanes, dcost=dcost), device=0, model=model)
    r = s.solve_rows(rows.cuda(), trajectories=False); torch.cuda.synchronize()
    out = {k: (v.clone().cpu() if v is not None else None) for k, v in r.items()}
    out['ms'] = s.last_kernel_ms(); s.close()
    return out

nom = lm.AscentParams().rows(1)
for B, nt in ((1000, 200), (4096, 200), (256, 801)):
    base = lm.dispersed_params(B, seed=21).rows(B)
    for w in (1.0, 2.0, 4.0):
        rows = nom + w * (base - nom)
        ref = solve(rows, 'thread', nt=nt)
        for lanes in (8, 32):
            got = solve(rows, 'coop', lanes, nt=nt)
            same = (got['status'] == ref['status'])
            both = (got['status'] == 0) & (ref['status'] == 0)
            dtf = ((got['tf'] - ref['tf']).abs() / ref['tf'])[both].max().item() if both.any() else float('nan')
            print(f'B={B} nt={nt} x{w} lanes {lanes}: {got["ms"]:.1f} ms (thread {ref["ms"]:.1f}) converged {int((got["status"] == 0).sum())} vs {int((ref["status"] == 0).sum())}, '
                  f'status mismatches {int((~same).sum())}, max rel dtf on common {dtf:.1e}, iters {got["iterations"].double().mean():.1f} vs {ref["iterations"].double().mean():.1f}', flush=True)
cn = lm.AscentParams.circular().rows(1)
B = 1000
base = lm.dispersed_params(B, seed=9).rows(B)
crow = cn + (base - nom) * (cn.abs() > 0)
crow[8] = crow[9] = cn[8] * (1 + 0.1 * (2 * torch.rand(B, dtype=torch.float64, generator=torch.Generator().manual_seed(3)) - 1))
ref = solve(crow, 'thread', model='circular', dcost=0.0)
for lanes in (8, 32):
    got = solve(crow, 'coop', lanes, model='circular', dcost=0.0)
    both = (got['status'] == 0) & (ref['status'] == 0)
    print(f'circular lanes {lanes}: converged {int((got["status"] == 0).sum())} vs {int((ref["status"] == 0).sum())}, max rel dtf {((got["tf"] - ref["tf"]).abs() / ref["tf"])[both].max().item():.1e}')
