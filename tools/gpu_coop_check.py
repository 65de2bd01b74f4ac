"""Cooperative kernel (8 lanes per problem) against the one-thread-per-problem kernel on the same inputs,
and their kernel times over batch sizes.  Developer script (GPU box)."""
import dataclasses
import sys
import time

import torch

sys.path.insert(0, ".")
import lunar_module_ascent_trajectory_optimiser_b200 as lm


def run(B, nt=200, model="elliptical", dcost=1e-5, seed=11, reps=2, traj=True, lanes=0):
    dev = torch.device("cuda", 0)
    out = {}
    for kern in ("thread", "coop"):
        opts = lm.SolverOptions(kernel=kern, dcost=dcost, coop_lanes=lanes)
        s = lm.AscentSolver(lm.Mesh(nt=nt), opts, device=dev, model=model)
        if model == "circular":
            rows = lm.AscentParams.circular().rows(B, device=dev)
        else:
            rows = (lm.dispersed_params(B, seed=seed) if B > 1 else lm.AscentParams()).rows(B, device=dev)
        ms = []
        for _ in range(reps):
            r = s.solve_rows(rows, trajectories=traj)
            ms.append(s.last_kernel_ms())
        torch.cuda.synchronize()
        out[kern] = (r, min(ms))
        s.close()
    a, b = out["thread"][0], out["coop"][0]
    conv = (int((a["status"] == 0).sum()), int((b["status"] == 0).sum()))
    both = (a["status"] == 0) & (b["status"] == 0)
    dtf = float(((a["tf"] - b["tf"]).abs() / a["tf"])[both].max()) if both.any() else float("nan")
    dtr = float("nan")
    if traj:
        sc = a["traj"].abs().amax(dim=1, keepdim=True) + 1e-300
        dtr = float((((a["traj"] - b["traj"]).abs() / sc)[:, :, both]).max()) if both.any() else float("nan")
    print(f"B={B:6d} nt={nt} {model} dcost={dcost:g} lanes={lanes}: converged thread/coop {conv}, max rel dtf {dtf:.1e}, dtraj {dtr:.1e}, "
          f"iters {float(a['iterations'].float().mean()):.1f}/{float(b['iterations'].float().mean()):.1f}, "
          f"kernel ms thread {out['thread'][1]:.2f} coop {out['coop'][1]:.2f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        run(1); run(700); run(1184); run(2048); run(4096); run(512, nt=2001, traj=False); run(4096, nt=2001, traj=False, reps=1)
        sys.exit(0)
    run(1); run(1, dcost=0.0); run(1, model="circular"); run(1, nt=40)
    run(37); run(700); run(1184); run(2048); run(4096); run(4736, traj=False)
    if len(sys.argv) > 1:
        for B in (6144, 8192, 16384):
            run(B, traj=False)
        run(64, nt=2001); run(512, nt=2001, traj=False); run(4096, nt=2001, traj=False, reps=1)
