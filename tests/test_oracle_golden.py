"""Pin the CPU oracle: against the reference's two published outputs (the only known-answer
vectors that exist for this path -- SURVEY.md Appendix D), against its own committed fixtures,
and against finite differences.  No GPU needed."""
import math
import os

import numpy as np
import pytest

from oracle.ascent_nlp import AscentNLP, AscentParams, collocation_matrix
from oracle.ipm_reference import IPMOptions, solve_ipm

# /root/reference/Numerical_results.png, printed by LO:188-194 (SI units = scaled value * 17703)
PNG = dict(y=-6430.82513705478, x=-290117.041689258, ydot=-273.361084935561, xdot=-1631.655147319155,
           ydoubledot=-2.38232397650751, xdoubledot=-5.52005887268682, tf_s=434.03530607609997)
# reference PDF p.30 (x-quantities printed with the sign flipped by that script, PDF p.28 src 154-160)
PDF = dict(tf=0.92616537474, tf_s=435.29772612780005, y=28716.160349635127, x=-294598.36483519967,
           ydot=-272.0993356840796, xdot=-1631.8810994353596, ydoubledot=-4.689807224722499,
           xdoubledot=-5.184464927832862)


def test_screenshot_self_consistency():
    """The published final state satisfies the terminal rows it was solved for (LO:161,169,173)."""
    p = AscentParams()
    Y = PNG["y"] + p.R0
    r = math.hypot(PNG["x"], Y)
    assert abs(r - (p.R0 + p.r_periapsis)) < 1.0                       # radius, 0.53 m of slack
    v = math.hypot(PNG["xdot"], PNG["ydot"])
    assert abs(v - p.v_target) / p.v_target < 1e-7
    assert abs(Y * PNG["ydot"] + PNG["x"] * PNG["xdot"]) / (r * v) < 1e-9
    m = p.M0 - p.M_dot * PNG["tf_s"]
    assert abs(math.hypot(PNG["ydoubledot"] + p.GM * Y / r ** 3, PNG["xdoubledot"] + p.GM * PNG["x"] / r ** 3)
               - p.Ft / m) < 1e-5                                       # thrust acceleration = Ft/m


# oracle (converged to 1e-12) minus screenshot, relative to the variable's largest magnitude on the trajectory:
# the per-variable residuals DESIGN.md section 6 tabulates.  The screenshot is an IPOPT iterate accepted at
# OTOL = RTOL = 1e-3 (LO:31-32), i.e. a point ON THE WAY to the optimum, not the optimum.
PNG_RESIDUAL = dict(tf_s=-1.8e-5, y=-2.5e-4, x=-6.1e-5, ydot=-6.1e-5, xdot=+1.7e-6, ydoubledot=-1.1e-3, xdoubledot=+8.7e-5)


def _final_state_residuals(nlp, x, S=17703.0):
    nv = nlp.node_values(x)
    out = {"tf_s": (x[nlp.i_tf] * 470.0 - PNG["tf_s"]) / PNG["tf_s"]}
    for k in ("y", "x", "ydot", "xdot", "ydoubledot", "xdoubledot"):
        out[k] = (nv[k][-1] * S - PNG[k]) / (np.abs(nv[k]).max() * S)
    return out


def test_oracle_reproduces_elliptical_screenshot(golden_dir):
    g = np.load(os.path.join(golden_dir, "elliptical_nominal_nt200.npz"))
    names = list(g["names"])
    tf_s = float(g["tf"]) * 470.0
    # the screenshot was produced at OTOL=RTOL=1e-3 (LO:31-32): good to ~2e-5, the bar is 1e-4
    assert abs(tf_s - PNG["tf_s"]) / PNG["tf_s"] < 1e-4
    assert abs(tf_s - 434.02765337) < 1e-6                             # converged optimum of the same NLP
    S = 17703.0
    for k in ("y", "x", "ydot", "xdot", "ydoubledot", "xdoubledot"):
        val = g["traj"][names.index(k), -1] * S
        scale = np.abs(g["traj"][names.index(k)]).max() * S
        res = (val - PNG[k]) / scale
        # each variable against its own documented residual (no blanket loosening): x, ydot, xdot, xdoubledot
        # are inside north_star's 1e-4; y (2.5e-4) and ydoubledot (1.1e-3) are not, see the next test for why
        assert abs(res - PNG_RESIDUAL[k]) < 0.15 * abs(PNG_RESIDUAL[k]) + 2e-6, (k, res)
        if k in ("x", "ydot", "xdot", "xdoubledot"):
            assert abs(res) < 1e-4, (k, res)
    fm = float(g["final_mass"])
    assert abs(fm - (4821 - 5.053 * tf_s)) < 1e-9


def test_screenshot_lies_on_the_oracles_central_path():
    """Why `y` and `ydoubledot` of the converged oracle miss the screenshot by more than 1e-4: the screenshot is
    not a converged point.  Stopping the oracle's own interior-point iteration early, at a scaled KKT error of
    3e-5 (barrier parameter 3e-6) on the reference's objective (DCOST included), lands on it: tf to 2e-6, y, x,
    ydot, xdot, xdoubledot to < 1e-4, and ydoubledot -- which follows the final pitch angle, the weakest
    determined quantity of this NLP (DESIGN.md "Tolerance") -- to 7e-4, changing sign between 1e-4 and 3e-5.
    Continuing to 1e-8 moves every quantity monotonically to the converged values of PNG_RESIDUAL."""
    nlp = AscentNLP(AscentParams(dcost=1e-5), nt=200, obj_scale=10.0)
    x0 = nlp.initial_guess(0.9)
    early = _final_state_residuals(nlp, solve_ipm(nlp, x0, IPMOptions(tol=3e-5)).x)
    assert abs(early["tf_s"]) < 5e-6, early
    for k in ("y", "x", "ydot", "xdot", "xdoubledot"):
        assert abs(early[k]) < 1e-4, (k, early)
    assert abs(early["ydoubledot"]) < 8e-4, early
    before = _final_state_residuals(nlp, solve_ipm(nlp, x0, IPMOptions(tol=1e-4)).x)
    late = _final_state_residuals(nlp, solve_ipm(nlp, x0, IPMOptions(tol=1e-8)).x)
    for k in ("tf_s", "y", "ydoubledot"):
        assert before[k] > early[k] > late[k], (k, before[k], early[k], late[k])       # the path crosses the screenshot
        assert abs(late[k] - PNG_RESIDUAL[k]) < 0.15 * abs(PNG_RESIDUAL[k]) + 2e-6, (k, late[k])


def test_oracle_reproduces_circular_pdf_output(golden_dir):
    g = np.load(os.path.join(golden_dir, "circular_nominal_nt200.npz"))
    names = list(g["names"])
    assert abs(float(g["tf"]) - PDF["tf"]) / PDF["tf"] < 1e-5
    S = 53108.4
    for k in ("y", "x", "ydot", "xdot", "ydoubledot", "xdoubledot"):
        val = g["traj"][names.index(k), -1] * S
        assert abs(val - PDF[k]) / abs(PDF[k]) < 1e-4, (k, val, PDF[k])


def test_live_oracle_matches_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "elliptical_nominal_nt40.npz"))
    nlp = AscentNLP(AscentParams(), nt=40, obj_scale=10.0)
    r = solve_ipm(nlp, nlp.initial_guess(0.9), IPMOptions(tol=1e-12))
    assert r.status == 0
    assert abs(r.x[nlp.i_tf] - float(g["tf"])) < 1e-10
    nv = nlp.node_values(r.x)
    names = list(g["names"])
    for n in names:
        assert np.allclose(nv[n], g["traj"][names.index(n)], rtol=0, atol=1e-8)


def test_alternative_transcriptions_do_not_match():
    """only backward Euler with the control acting on the step that ends at
    its node reproduces the reference; a coarse check that the restatement is not accidental."""
    nlp = AscentNLP(AscentParams(), nt=60, obj_scale=10.0)
    r = solve_ipm(nlp, nlp.initial_guess(0.9), IPMOptions(tol=1e-8))
    assert r.status == 0
    # backward Euler under-estimates the burn on a coarse mesh; it approaches 435.2 s from below
    assert 430.0 < r.x[nlp.i_tf] * 470 < 434.03


@pytest.mark.parametrize("model", ["elliptical", "circular"])
def test_oracle_derivatives_fd(model):
    p = AscentParams() if model == "elliptical" else AscentParams.circular()
    nlp = AscentNLP(p, nt=7)
    rng = np.random.default_rng(0)
    x = nlp.initial_guess(0.9) + 1e-3 * rng.standard_normal(nlp.n)
    J = nlp.jac(x).toarray()
    eps = 1e-6
    Jfd = np.zeros_like(J)
    for i in range(nlp.n):
        e = np.zeros(nlp.n); e[i] = eps
        Jfd[:, i] = (nlp.c(x + e) - nlp.c(x - e)) / (2 * eps)
    assert np.abs(J - Jfd).max() < 1e-6 * max(1.0, np.abs(J).max())
    lam = rng.standard_normal(nlp.m)
    H = nlp.hess(x, lam).toarray()
    Hfd = np.zeros_like(H)
    for i in range(nlp.n):
        e = np.zeros(nlp.n); e[i] = eps
        Hfd[:, i] = (nlp.jac(x + e).T @ lam - nlp.jac(x - e).T @ lam) / (2 * eps)
    assert np.abs(H - Hfd).max() < 1e-5 * max(1.0, np.abs(H).max())
    assert np.abs(H - H.T).max() == 0.0


def test_collocation_matrices_match_apmonitor_tables():
    """SURVEY Appendix B.2 (values published in APMonitor's course notes)."""
    _, N2 = collocation_matrix(2)
    assert N2.tolist() == [[1.0]]
    tau3, N3 = collocation_matrix(3)
    assert np.allclose(tau3, [0, .5, 1]) and np.allclose(N3, [[.75, -.25], [1.0, 0.0]], atol=1e-12)
    tau4, N4 = collocation_matrix(4)
    assert np.allclose(tau4, [0, .276393, .723607, 1], atol=1e-6)
    assert np.allclose(N4, [[.436339, -.280547, .120601], [.613880, .063661, .046066], [.603006, .230328, .166667]], atol=2e-6)
    tau5, N5 = collocation_matrix(5)
    assert np.allclose(N5[3], [.388889, .222222, .388889, 0.0], atol=2e-6)


def test_higher_order_collocation_converges_to_continuous_optimum():
    """NODES=3 on a coarse mesh is already close to the continuous-time optimum (~435.1 s) that
    backward Euler only reaches on dense meshes (SURVEY Appendix D.3 side note)."""
    nlp = AscentNLP(AscentParams(), nt=30, nodes=3, obj_scale=10.0)
    r = solve_ipm(nlp, nlp.initial_guess(0.9), IPMOptions(tol=1e-8))
    assert r.status == 0
    assert 434.5 < r.x[nlp.i_tf] * 470 < 435.8


def test_sensitivity_fixture_is_consistent():
    """tests/golden/sens_nominal_nt40.npz (oracle finite differences, `make_golden.py --sens`): signs and
    magnitudes follow the physics (more thrust / flow / pitch authority shorten the burn, more mass
    lengthens it) and the first-order mass relation d tf/d M0 ~ -(Ft/M0) d tf/d Ft holds to 10 %."""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sens_nominal_nt40.npz"))
    d = dict(zip([str(n) for n in g["names"]], g["dtf"]))
    assert d["Ft"] < 0 and d["M_dot"] < 0 and d["angle_doubledot_max"] < 0 and d["M0"] > 0
    assert abs(d["M0"] / (-(15346.0 / 4821.0) * d["Ft"]) - 1.0) < 0.4
