"""CPU baseline harness: the oracle (restatement, NOT GEKKO/IPOPT -- those are absent from this
image) solving one problem per process on all host cores.  TEST/BENCH INFRASTRUCTURE ONLY; used
by bench.py's ``cpu_baseline`` leg and by ``bench.py --impl reference``.

Each worker builds the NLP callables once (sympy lambdify) and then solves the problems it is
handed with the same options as the GPU arm (tol, obj_scale, nt).  Reference path being timed:
m.solve() at /root/reference/Launch_Optimiser.py:177.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import Tuple

import numpy as np

_ROWS = ["G", "M", "R0", "Ft", "M0", "M_dot", "fuel_mass", "angle_doubledot_max", "r_periapsis",
         "r_apoapsis", "final_time", "mass_scalar", "angle_ub", "u_bound"]


def _init_worker():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle.ascent_nlp import _build_node_model
    _build_node_model("elliptical")


def solve_one_gekko(args) -> Tuple[float, int, int, float]:
    """The reference's own stack (only when tools/probe_gekko.py found it usable): GEKKO(remote=False), one
    problem per process, with the reference's OTOL = RTOL = 1e-3 (LO:31-32)."""
    col, nt, _tol, _obj_scale, _dcost = args
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools.probe_gekko import declare_and_solve
    kw = dict(zip(_ROWS, [float(v) for v in col]))
    t0 = time.perf_counter()
    try:
        m, tf, _v = declare_and_solve(kw, nt=nt)
        return float(tf.value[0]), 0, int(m.options.ITERATIONS), time.perf_counter() - t0
    except Exception:
        return float("nan"), 1, 0, time.perf_counter() - t0


def solve_one(args) -> Tuple[float, int, int, float]:
    col, nt, tol, obj_scale, dcost = args
    from oracle.ascent_nlp import AscentNLP, AscentParams
    from oracle.ipm_reference import IPMOptions, solve_ipm
    kw = dict(zip(_ROWS, [float(v) for v in col]))
    p = AscentParams(G=kw["G"], M=kw["M"], R0=kw["R0"], Ft=kw["Ft"], M0=kw["M0"], M_dot=kw["M_dot"],
                     fuel_mass=kw["fuel_mass"], angle_doubledot_max=kw["angle_doubledot_max"],
                     r_periapsis=kw["r_periapsis"], r_apoapsis=kw["r_apoapsis"], final_time=kw["final_time"],
                     mass_scalar=kw["mass_scalar"], angle_ub=kw["angle_ub"], u_bound=kw["u_bound"], dcost=dcost)
    t0 = time.perf_counter()
    nlp = AscentNLP(p, nt=nt, obj_scale=obj_scale)
    r = solve_ipm(nlp, nlp.initial_guess(0.9), IPMOptions(tol=tol, max_iter=500))
    return float(r.x[nlp.i_tf]), int(r.status), int(r.iterations), time.perf_counter() - t0


class OraclePool:
    """One worker process per host core, reused across steps."""

    def __init__(self, cores: int = 0, gekko: bool = False):
        self.gekko = gekko
        self.cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_init_worker)
        # make sure every worker has finished building its model before anything is timed
        self.pool.map(_noop, range(self.cores * 2))

    def solve(self, rows: np.ndarray, nt: int, tol: float, obj_scale: float, dcost: float = 1e-5):
        """rows: [NPARAM, n].  Returns (tf[n], status[n], iters[n], wall seconds)."""
        n = rows.shape[1]
        t0 = time.perf_counter()
        res = self.pool.map(solve_one_gekko if self.gekko else solve_one, [(rows[:, i].copy(), nt, tol, obj_scale, dcost) for i in range(n)], chunksize=1)
        wall = time.perf_counter() - t0
        tf = np.array([r[0] for r in res])
        st = np.array([r[1] for r in res])
        it = np.array([r[2] for r in res])
        return tf, st, it, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def _noop(i):
    time.sleep(0.05)
    return i
