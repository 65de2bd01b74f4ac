"""GEKKO-shaped result objects: the reference's post-processing code (LO:178-202)
must run unchanged on a solution.  Uses the oracle's golden vector as the solution, no GPU."""
import math
import os

import numpy as np
import torch

import lunar_module_ascent_trajectory_optimiser_b200 as lm
from lunar_module_ascent_trajectory_optimiser_b200 import gekko_shim


def _solution_from_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "elliptical_nominal_nt200.npz"))
    names = list(g["names"])
    st = {n: torch.from_numpy(g["traj"][names.index(n)].copy()) for n in names if n != "angledoubledot"}
    return lm.AscentSolution(tf=float(g["tf"]), tf_seconds=float(g["tf"]) * 470.0, states=st,
                             control=torch.from_numpy(g["traj"][names.index("angledoubledot")].copy()),
                             final_mass=float(g["final_mass"]), status=0, iterations=25, kkt_error=1e-12,
                             time=torch.from_numpy(g["time"].copy()))


def test_reference_postprocessing_runs_on_the_shim(golden_dir):
    sol = _solution_from_golden(golden_dir)
    g = gekko_shim.as_gekko(sol)
    tf, x, y, angle, m = g.tf, g.x, g.y, g.angle, g.m
    Rfmin_py, R0_py, final_time = 17703, 1738100, 470
    # LO:187-202, verbatim logic
    ts = m.time * tf.value[0]
    y_pos_list = [0] * len(x.value)
    x_pos_list = [0] * len(x.value)
    theta_list = [0] * len(x.value)
    for i in range(len(x.value)):
        x_pos_list[i] = -x.value[i] * Rfmin_py
        y_pos_list[i] = y.value[i] * Rfmin_py + R0_py
        theta_list[i] = 3 * angle.value[i] * (180 / (np.pi))
    assert len(ts) == 200 and abs(final_time * ts[-1] - 434.0276531) < 1e-6
    assert x_pos_list[0] == 0 and abs(x_pos_list[-1] - 290134.78) < 0.1          # downrange, LO:200
    assert abs(math.hypot(x_pos_list[-1], y_pos_list[-1]) - (R0_py + Rfmin_py)) < 1e-3
    assert abs(theta_list[-1] - 88.526) < 1e-2 and max(theta_list) < 180.0
    assert g.m.options.APPSTATUS == 1 and g.m.options.SOLVER == 3 and g.m.options.NODES == 2


def test_summary_and_results_json(golden_dir):
    sol = _solution_from_golden(golden_dir)
    txt = gekko_shim.print_reference_summary(sol)
    lines = txt.splitlines()
    assert lines[0].startswith("Optimal Solution (final time): 434.02765")       # LO:178
    assert lines[1].startswith("final y -6434.33") and lines[2].startswith("final x -290134.7")
    assert lines[-1].startswith("final time 434.02765")                           # LO:194
    d = gekko_shim.results_dict(sol)
    assert set(d) >= {"time", "tf", "y", "x", "ydot", "xdot", "angle", "angledot", "mass", "angledoubledot"}
    assert all(len(v) == 200 for v in d.values())
    assert d["tf"][0] == d["tf"][-1] == sol.tf
    import json
    assert json.loads(gekko_shim.results_json(sol))["mass"][-1] == d["mass"][-1]
