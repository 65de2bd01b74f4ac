import sys, torch
sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm
dev = torch.device("cuda", 0)
B = 65536
for kern in ("thread", "coop"):
    for warm in (True, False):
        s = lm.AscentSolver(lm.Mesh(nt=200), lm.SolverOptions(kernel=kern, dcost=0.0, warm_start=warm), device=dev)
        rows = lm.dispersed_params(B, seed=11).rows(B, device=dev)
        r = s.solve_rows(rows, trajectories=False); r = s.solve_rows(rows, trajectories=False)
        it = r["iterations"]
        print(f"{kern} warm {warm}: {s.last_kernel_ms():.1f} ms mean {it.float().mean():.2f} max {int(it.max())} >40: {int((it > 40).sum())} >100: {int((it > 100).sum())} status {torch.bincount(r['status'].long()).tolist()}")
        top = torch.topk(it, 5)
        print("   worst:", top.values.tolist(), top.indices.tolist())
        s.close()
