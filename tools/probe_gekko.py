"""Probe for the reference's own solver stack (GEKKO + its bundled `apm` + IPOPT) and, when it is there, solve
the reference's model with it.  BASELINE.md section 2 asks for this probe before every benchmark; SURVEY.md
section 8(c) found the stack absent from this image (no network, not in /opt/wheelhouse), in which case the
benchmark's CPU arm is the oracle port and says so.

    python tools/probe_gekko.py                 # prints what it found; exit 0 either way
    python tools/probe_gekko.py --fixtures      # with GEKKO present: also writes tests/golden/gekko_*.npz

The model below is the parameterised DECLARATION of /root/reference/Launch_Optimiser.py (LO:19-176) through
GEKKO's public API: that script has no function to call (it builds, solves and plots on import), so a harness
that wants to run it on other parameters has to declare it again.  It is test/bench infrastructure only.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NOMINAL = dict(G=6.674e-11, M=7.346e22, R0=1738100.0, Ft=15346.0, M0=4821.0, M_dot=5.053, fuel_mass=2376.0,
               angle_doubledot_max=5e-4, r_periapsis=17703.0, r_apoapsis=88615.0, final_time=470.0)


def declare_and_solve(p: dict, nt: int = 200, nodes: int = 2, otol: float = 1e-3, rtol: float = 1e-3, disp: bool = False):
    """Returns (m, variables) after m.solve().  Raises whatever GEKKO raises ("@error: Solution Not Found")."""
    from gekko import GEKKO
    m = GEKKO(remote=False)                                     # BASELINE.json config 1: the local apm executable
    m.time = np.linspace(0, 1, nt)                              # LO:20-21
    o = m.options
    o.NODES, o.SOLVER, o.IMODE, o.MAX_ITER, o.MV_TYPE = nodes, 3, 6, 20000, 0     # LO:25-29
    o.OTOL, o.RTOL = otol, rtol                                 # LO:31-32
    T = p["final_time"]
    tf = m.FV(value=0, lb=0, ub=1); tf.STATUS = 1               # LO:39-40
    S, R0, GM = p["r_periapsis"], p["R0"], p["G"] * p["M"]      # LO:50-52, 73, 107
    mflow = p["M_dot"] / p["fuel_mass"]                         # LO:65
    asc = p["angle_doubledot_max"] / 3.0                        # LO:109
    vp = np.sqrt(GM / (R0 + 0.5 * (p["r_periapsis"] + p["r_apoapsis"])))          # LO:72-75
    v = {n: m.Var(value=0) for n in ("y", "ydot", "ydoubledot", "x", "xdot", "xdoubledot", "angledot")}   # LO:84-95
    v["mass"] = m.Var(value=0, lb=0, ub=1)                      # LO:83
    v["angle"] = m.Var(value=0, lb=0, ub=np.pi / 3)             # LO:94
    u = m.MV(value=0, lb=-1, ub=1); u.STATUS = 1; u.DCOST = 1e-5; u.REQONCTRL = 3   # LO:96-100
    y, x, a, ms = v["y"], v["x"], v["angle"], v["mass"]
    m.Equations([y.dt() == v["ydot"] * tf * T, v["ydot"].dt() == v["ydoubledot"] * tf * T,       # LO:114-118
                 x.dt() == v["xdot"] * tf * T, v["xdot"].dt() == v["xdoubledot"] * tf * T,
                 a.dt() == v["angledot"] * tf * T, v["angledot"].dt() == u * asc * tf * T,      # LO:120-121
                 ms.dt() == mflow * T * tf])                                                     # LO:123
    X, Y = x * S, y * S + R0
    r = (X ** 2 + Y ** 2) ** 0.5
    k = p["Ft"] / ((p["M0"] - p["fuel_mass"] * ms) * r)
    m.Equation(v["ydoubledot"] == (k * (Y * m.cos(3 * a) + X * m.sin(3 * a)) - Y * GM / r ** 3) / S)   # LO:127-130
    m.Equation(v["xdoubledot"] == (k * (X * m.cos(3 * a) - Y * m.sin(3 * a)) - X * GM / r ** 3) / S)   # LO:133-136
    for name in ("y", "x", "ydot", "xdot", "angle", "mass"):
        m.fix(v[name], pos=0, val=0)                            # LO:145-151
    slack_everywhere = np.full(nt, S + R0 + 1.0); slack_everywhere[-1] = 0.0                    # LO:158-160
    last_only = np.zeros(nt); last_only[-1] = 1.0                                                # LO:166-168
    pr, pv = m.Param(value=slack_everywhere), m.Param(value=last_only)
    m.Equation(((y + R0 / S) ** 2 + x ** 2) ** 0.5 + pr >= (R0 + S) / S)                         # LO:161
    m.Equation(v["xdot"] ** 2 + v["ydot"] ** 2 >= (vp / S) ** 2 * pv)                            # LO:169
    m.Equation(((y * S + R0) * (v["ydot"] * S) + (x * S) * (v["xdot"] * S)) * pv == 0)           # LO:173
    m.Minimize(tf)                                              # LO:176
    m.solve(disp=disp)                                          # LO:177
    v["angledoubledot"] = u
    return m, tf, v


def solve_reference_model(p: dict | None = None, nt: int = 200, **kw):
    """(tf, wall seconds) of one local GEKKO solve."""
    t0 = time.perf_counter()
    _m, tf, _v = declare_and_solve(dict(NOMINAL, **(p or {})), nt=nt, **kw)
    return float(tf.value[0]), time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fixtures", action="store_true")
    args = ap.parse_args()
    try:
        import gekko
    except Exception as e:
        print(f"gekko: NOT importable ({type(e).__name__}: {e}); the CPU arm of bench.py uses the oracle port")
        return 0
    print("gekko", getattr(gekko, "__version__", "?"), "found at", os.path.dirname(gekko.__file__))
    for otol in (1e-3, 1e-8):
        try:
            m, tf, v = declare_and_solve(NOMINAL, otol=otol, rtol=otol)
        except Exception as e:
            print(f"GEKKO(remote=False) solve failed at OTOL=RTOL={otol:g}: {type(e).__name__}: {e}")
            continue
        print(f"OTOL=RTOL={otol:g}: tf*470 = {tf.value[0] * 470.0:.9f} s, iterations {m.options.ITERATIONS}, "
              f"solve time {m.options.SOLVETIME:.2f} s, objective {m.options.OBJFCNVAL}")
        if args.fixtures:
            names = ["y", "ydot", "ydoubledot", "x", "xdot", "xdoubledot", "angle", "angledot", "mass", "angledoubledot"]
            out = os.path.join(ROOT, "tests", "golden", f"gekko_nominal_nt200_tol{otol:g}.npz")
            np.savez(out, tf=tf.value[0], traj=np.array([v[n].value for n in names]), names=np.array(names),
                     iterations=m.options.ITERATIONS, gekko_version=getattr(gekko, "__version__", "?"))
            print("wrote", out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
