"""The eight shards of the bench's weak-scaling draw (524 288 problems, seed 4011) solved one after the other on one GPU:
kernel time, iteration statistics and non-converged counts per shard (which shard makes the weak step slow?)."""
import sys, torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm
dev = torch.device("cuda", 0)
Bw, world = 8 * 65536, 8
rows = lm.dispersed_params(Bw, seed=4011).rows(Bw)
s = lm.AscentSolver(lm.Mesh(nt=200), lm.SolverOptions(), device=dev)
for r in range(world):
    lo, hi = lm.shard_bounds(Bw, world, r)
    sh = rows[:, lo:hi].contiguous().to(dev)
    out = s.solve_rows(sh, trajectories=False); out = s.solve_rows(sh, trajectories=False)
    it = out["iterations"]
    print(f"shard {r}: {s.last_kernel_ms():.1f} ms  iters mean {it.float().mean():.2f} max {int(it.max())}  >32: {int((it > 32).sum())}  status {torch.bincount(out['status'].long()).tolist()}", flush=True)
