"""B = 4096, nt = 2001, all six dispersion columns, sharded over the visible GPUs from one
process (optimise_batch(devices=...)); wall time of the call, device-resident inputs and outputs."""
import sys, time, torch
sys.path.insert(0, '.')
import dataclasses
import lunar_module_ascent_trajectory_optimiser_b200 as lm
B, nt = 4096, 2001
p = lm.dispersed_params(B, seed=11)
pc = dataclasses.replace(p, **{f.name: getattr(p, f.name).cuda() for f in dataclasses.fields(p)
                               if isinstance(getattr(p, f.name), torch.Tensor)})
n = torch.cuda.device_count()
for devs in ([0], list(range(n))):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        sol = lm.optimise_batch(pc, lm.Mesh(nt=nt), devices=devs, trajectories=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(f'devices {devs}: {dt*1e3:.0f} ms per batch = {B/dt:.0f} solves/s, converged {int((sol.status == 0).sum())}/{B}, '
          f'iterations mean {sol.iterations.double().mean():.1f}, tf[0] {float(sol.tf_seconds[0]):.6f} s')
    if len(devs) == n == 1: break

# breakdown for the multi-GPU case: launches only (all devices synchronised), then the gather
if n > 1:
    from lunar_module_ascent_trajectory_optimiser_b200.api import _get_solver, shard_bounds
    opts = lm.SolverOptions(warm_start=2, dcost=1e-5)
    solvers = [_get_solver(lm.Mesh(nt=nt), opts, d, 'elliptical') for d in range(n)]
    rows = pc.rows(B, device='cuda:0')
    shards = [rows[:, shard_bounds(B, n, g)[0]:shard_bounds(B, n, g)[1]].to(s.device).contiguous() for g, s in enumerate(solvers)]
    for rep in range(2):
        for d in range(n): torch.cuda.synchronize(d)
        t0 = time.time()
        parts = []
        for g, s in enumerate(solvers):
            with torch.cuda.device(s.device):
                parts.append(s.solve_rows(shards[g], True))
        t1 = time.time()
        for d in range(n): torch.cuda.synchronize(d)
        t2 = time.time()
    print(f'launch loop {1e3*(t1-t0):.0f} ms, until all devices done {1e3*(t2-t0):.0f} ms; kernel ms per device {[round(s.last_kernel_ms()) for s in solvers]}')
