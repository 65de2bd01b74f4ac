"""GEKKO-shaped view of a solution.

The reference's post-processing (LO:178-244 of /root/reference/Launch_Optimiser.py, and the
plotting code in the reference PDF p.28) reads ``tf.value[0]``, ``x.value[i]``, ``m.time`` and
``m.options.*`` from GEKKO objects.  ``as_gekko(sol)`` returns a small namespace with the same
attribute shapes, so that code runs unchanged on the CUDA solver's output:

    g = as_gekko(optimise())
    tf, x, y, angle, m = g.tf, g.x, g.y, g.angle, g.m
    ts = m.time * tf.value[0]                                   # LO:187
    x_pos = [-x.value[i] * 17703 for i in range(len(x.value))]  # LO:200

Also provides ``results_dict`` in the layout of GEKKO's ``results.json`` (variable name ->
list of nt values, plus "time").
"""
from __future__ import annotations

import json
from types import SimpleNamespace
from typing import Dict, List, Union

import numpy as np

from .api import AscentBatchSolution, AscentSolution

# GEKKO APPSTATUS / APPINFO conventions: 1 / 0 on success
_VAR_ORDER = ["mass", "y", "ydot", "ydoubledot", "x", "xdot", "xdoubledot", "angle", "angledot",
              "angledoubledot"]   # declaration order LO:83-96


class GKValue:
    """Stands in for a GEKKO Var/MV/FV after a solve: ``.value`` is a python list (LO:188-202)."""

    def __init__(self, name: str, value: List[float]):
        self.name = name
        self.value = value
        self.VALUE = value

    def __len__(self):
        return len(self.value)

    def __getitem__(self, i):
        return self.value[i]

    def __repr__(self):
        return f"GKValue({self.name!r}, n={len(self.value)})"


def _one(sol: Union[AscentSolution, AscentBatchSolution], index: int):
    if isinstance(sol, AscentSolution):
        states = {k: v.cpu().numpy() for k, v in sol.states.items()}
        ctrl = sol.control.cpu().numpy()
        states[sol.control_name] = ctrl
        return (sol.tf, states, ctrl, sol.time.cpu().numpy(), sol.status, sol.iterations, sol.final_mass,
                sol.tf_seconds)
    states = {k: v[index].cpu().numpy() for k, v in sol.states.items()}
    ctrl = sol.control[index].cpu().numpy()
    states[sol.control_name] = ctrl       # the MV: `angledoubledot` (LO:96), or `angle` in the circular model
    return (float(sol.tf[index]), states, ctrl, sol.time.cpu().numpy(), int(sol.status[index]),
            int(sol.iterations[index]), float(sol.final_mass[index]), float(sol.tf_seconds[index]))


def as_gekko(sol: Union[AscentSolution, AscentBatchSolution], index: int = 0) -> SimpleNamespace:
    tf, states, ctrl, time, status, iters, fmass, tf_s = _one(sol, index)
    nt = len(time)
    ns = SimpleNamespace()
    for name in _VAR_ORDER:
        arr = states.get(name)         # (the circular model has no angledot / angledoubledot: PDF p.27 src 54-73)
        if arr is None:
            continue
        setattr(ns, name, GKValue(name, [float(v) for v in arr]))
    ns.tf = GKValue("tf", [tf] * nt)                       # an FV carries one value at every node
    options = SimpleNamespace(APPSTATUS=1 if status == 0 else 0, APPINFO=0 if status == 0 else 1,
                              ITERATIONS=iters, OBJFCNVAL=tf, NODES=2, IMODE=6, SOLVER=3, MAX_ITER=20000)
    ns.m = SimpleNamespace(time=np.asarray(time, dtype=float), options=options)
    ns.final_mass = fmass
    ns.final_time_seconds = tf_s
    return ns


def results_dict(sol: Union[AscentSolution, AscentBatchSolution], index: int = 0) -> Dict[str, List[float]]:
    g = as_gekko(sol, index)
    out = {"time": g.m.time.tolist(), "tf": g.tf.value}
    for name in _VAR_ORDER:
        if hasattr(g, name):
            out[name] = getattr(g, name).value
    return out


def results_json(sol: Union[AscentSolution, AscentBatchSolution], index: int = 0) -> str:
    return json.dumps(results_dict(sol, index))


def print_reference_summary(sol: Union[AscentSolution, AscentBatchSolution], index: int = 0,
                            Rfmin_py: float = 17703.0, final_time: float = 470.0) -> str:
    """Reproduce the lines the reference prints at LO:178, 188-194 (same order, same scaling)."""
    g = as_gekko(sol, index)
    lines = [f"Optimal Solution (final time): {g.tf.value[0] * final_time}",
             f"final y {g.y.value[-1] * Rfmin_py}", f"final x {g.x.value[-1] * Rfmin_py}",
             f"final ydot {g.ydot.value[-1] * Rfmin_py}", f"final xdot {g.xdot.value[-1] * Rfmin_py}",
             f"final ydoubledot {g.ydoubledot.value[-1] * Rfmin_py}",
             f"final xdoubledot {g.xdoubledot.value[-1] * Rfmin_py}",
             f"final time {g.tf.value[0] * final_time}"]
    return "\n".join(lines)
