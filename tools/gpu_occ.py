"""Developer experiment: per-iteration time of a warp as a function of how many warps share an SM
(1, 4, 8 warps per SM in one wave) -- separates memory effects from issue contention."""
import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
nt = 200
for B in (148 * 32, 148 * 2 * 32, 148 * 4 * 32, 148 * 8 * 32, 65536):
    solver = lm.AscentSolver(lm.Mesh(nt=nt), lm.SolverOptions(warm_start=False), device=0)
    rows = lm.dispersed_params(B).rows(B).cuda()
    best = 1e9
    for rep in range(3):
        raw = solver.solve_rows(rows); torch.cuda.synchronize()
        best = min(best, solver.last_kernel_ms())
    it = raw['iterations'].double()
    # iterations a warp executes = max over its 32 lanes
    wmax = it[: (B // 32) * 32].view(-1, 32).max(dim=1).values
    print(f'B {B} warps/SM {B/148/32:.2f} kernel ms {best:.2f} iters mean {it.mean():.2f} warp-max mean {wmax.mean():.2f} max {wmax.max():.0f}'
          f'  ms per warp-iteration (vs slowest warp) {best/wmax.max():.3f}')
