"""Barrier floor / polish study with the reference's DCOST term (default kernel, config 4): time, iterations and the
worst per-variable difference to (a) the default settings and (b) a cold-start solve with the same settings, plus the
DCOST golden fixtures.  Developer script (GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rows = lm.dispersed_params(B).rows(B).cuda()
g = np.load("tests/golden/elliptical_dcost1e-5_disp4_seed11_nt200.npz")
VAR = ["y", "ydot", "ydoubledot", "x", "xdot", "xdoubledot", "angle", "angledot", "mass", "angledoubledot"]

def solve(mmf, npol, warm=True):
    s = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(warm_start=warm, n_polish=npol, mu_min_factor=mmf), device=0)
    raw = s.solve_rows(rows); torch.cuda.synchronize()
    raw = s.solve_rows(rows); torch.cuda.synchronize()
    out = {k: (v.clone() if v is not None else None) for k, v in raw.items()}
    out["ms"] = s.last_kernel_ms()
    s.close()
    return out

def rel(a, b):
    scale = b.abs().amax(dim=1, keepdim=True)
    return ((a - b).abs() / scale).amax(dim=1).amax(dim=1)      # per variable

def golden(mmf, npol):
    p = lm.dispersed_params(4, seed=11)
    sol = lm.optimise_batch(p, options=lm.SolverOptions(n_polish=npol, mu_min_factor=mmf))
    worst = np.zeros(10)
    for b in range(4):
        tr = np.stack([(sol.control if n == "angledoubledot" else sol.states[n])[b].cpu().numpy() for n in VAR])
        gt = g["traj"][b]
        worst = np.maximum(worst, (np.abs(tr - gt) / (np.abs(gt).max(axis=1, keepdims=True) + 1e-300)).max(axis=1))
    dtf = float(np.max(np.abs(sol.tf.cpu().numpy() - g["tf"]) / g["tf"]))
    return dtf, worst

base = solve(1e-3, 2)
print(f"default: {base['ms']:.1f} ms iters {base['iterations'].double().mean():.2f}")
for mmf, npol in ((1e-3, 2), (1e-3, 1), (1e-2, 2), (1e-2, 1), (1e-1, 2), (1e-1, 1), (1e-2, 3)):
    w = solve(mmf, npol)
    c = solve(mmf, npol, warm=False)
    rb = rel(w["traj"], base["traj"]); rc = rel(w["traj"], c["traj"])
    dtf, gw = golden(mmf, npol)
    print(f"mu_min_factor {mmf:g} n_polish {npol}: {w['ms']:.1f} ms iters {w['iterations'].double().mean():.2f} fails {(w['status'] != 0).sum().item()}"
          f" | vs default: states {rb[:9].max():.1e} control {rb[9]:.1e} tf {((w['tf'] - base['tf']).abs() / base['tf']).max():.1e}"
          f" | warm vs cold: states {rc[:9].max():.1e} control {rc[9]:.1e}"
          f" | golden: tf {dtf:.1e} states {gw[:9].max():.1e} control {gw[9]:.1e}", flush=True)
