"""Developer experiment: kernel time with and without the batch warm start per batch size
(second call on the same handle, i.e. with the cheap re-converged reference solve)."""
import sys, torch
sys.path.insert(0, '.')
import lunar_module_ascent_trajectory_optimiser_b200 as lm
for B in (512, 1024, 2048, 4096, 8192, 16384):
    out = []
    for ws in (0, 2):
        solver = lm.AscentSolver(lm.Mesh(nt=200), lm.SolverOptions(warm_start=ws), device=0)
        best, first = 1e9, None
        for rep in range(4):
            rows = lm.dispersed_params(B, seed=11 + rep).rows(B).cuda()
            raw = solver.solve_rows(rows); torch.cuda.synchronize()
            ms = solver.last_kernel_ms()
            first = ms if first is None else first
            if rep > 0: best = min(best, ms)
        out.append(f'ws={ws}: first {first:.2f} ms, later best {best:.2f} ms, fails {(raw["status"]!=0).sum().item()}')
    print(B, ' | '.join(out))
