"""Batch sizes at the boundaries of the kernel choice (warp, warm-start threshold, lanes per problem, cooperative /
thread-per-problem switch, one wave, beyond a wave): convergence, kernel time, and the first 500 problems compared
with the same problems solved inside the other batches."""
import sys, torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm
solver = lm.AscentSolver(lm.Mesh(), lm.SolverOptions(), device=0)
ref = None
for B in (1, 31, 33, 511, 512, 513, 1184, 1185, 6144, 6145, 37888, 37889, 65535, 131072):
    rows = lm.dispersed_params(B, seed=11).rows(B).cuda()
    raw = solver.solve_rows(rows, trajectories=(B <= 65536)); torch.cuda.synchronize()
    st, tf, it = raw["status"], raw["tf"], raw["iterations"]
    d = float("nan")
    if B >= 500:
        if ref is None: ref = tf[:500].clone()
        d = ((tf[:500] - ref) / ref).abs().max().item()
    print(f"B={B}: {solver.last_kernel_ms():.2f} ms converged {int((st == 0).sum())}/{B} iters {it.float().mean():.2f} "
          f"max rel dtf of the first 500 vs B=511: {d:.1e}", flush=True)
