"""The device solver's SOURCE, compiled for the host.

`csrc/ascent_model.cuh`, `ascent_ipm.cuh` and `ascent_ipm_dc.cuh` are written so that g++ accepts them
(tools/hostsim): the three sweeps, the IPM driver and the model derivatives below are the very lines the
CUDA kernel inlines, executed by one CPU thread.  This gives the CPU-only test tier a check of the
solver's logic against the oracle's golden vectors and of the hand-derived derivatives against finite
differences.  It is test infrastructure only: the product has no CPU path (`_cabi.lib()` raises without
the CUDA library), and nothing here is imported by the package.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
NOMINAL = np.array([6.674e-11, 7.346e22, 1738100.0, 15346.0, 4821.0, 5.053, 2376.0, 5e-4, 17703.0, 88615.0, 470.0,
                    2376.0, np.pi / 3, 1.0])


@pytest.fixture(scope="module")
def hostsim(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("hostsim") / "libhostsim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I",
                           os.path.join(ROOT, "lunar_module_ascent_trajectory_optimiser_b200", "csrc"),
                           os.path.join(ROOT, "tools", "hostsim", "hostsim.cpp"), "-o", so])
    L = C.CDLL(so)
    L.hostsim_check_derivatives.restype = C.c_double
    L.hostsim_check_derivatives.argtypes = [C.c_void_p] + [C.c_double] * 4
    return L


def _solve(L, raw, nt=200, time=None, tol=1e-10, fn="hostsim_solve"):
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    traj = np.empty((10, nt))
    tf, it, kkt = C.c_double(), C.c_int(), C.c_double()
    tp = None if time is None else np.ascontiguousarray(time, dtype=np.float64).ctypes.data_as(C.c_void_p)
    st = getattr(L, fn)(raw.ctypes.data_as(C.c_void_p), nt, tp, C.c_double(tol), C.c_double(10.0), C.c_double(1e-3),
                         traj.ctypes.data_as(C.c_void_p), C.byref(tf), C.byref(it), C.byref(kkt))
    return st, tf.value, it.value, traj


def _rel(traj, gold):
    return (np.abs(traj - gold) / (np.abs(gold).max(axis=1, keepdims=True) + 1e-300)).max(axis=1)


def test_model_derivatives_match_finite_differences(hostsim):
    """accel_first / accel_second (LO:127-136 differentiated by hand) at points along a trajectory."""
    raw = NOMINAL.copy()
    for y, x, a, m in [(0.0, 0.0, 0.05, 0.0), (0.4, -3.0, 0.3, 0.3), (-0.36, -16.4, 0.51, 0.92), (1.2, -9.0, 0.9, 0.6)]:
        err = hostsim.hostsim_check_derivatives(raw.ctypes.data_as(C.c_void_p), y, x, a, m)
        assert err < 2e-6, (y, x, a, m, err)


def test_device_ipm_source_matches_goldens(hostsim, monkeypatch):
    monkeypatch.delenv("WDC", raising=False)
    monkeypatch.delenv("CIRCULAR", raising=False)
    # config 1 (7-state sweeps, no DCOST) and a non-uniform mesh
    g = np.load(os.path.join(GOLDEN, "elliptical_nominal_nt200.npz"))
    st, tf, it, traj = _solve(hostsim, NOMINAL)
    assert st == 0 and abs(tf - float(g["tf"])) / float(g["tf"]) < 1e-8
    err = _rel(traj, g["traj"])
    assert err[:9].max() < 1e-4 and err[9] < 2e-4, err
    g = np.load(os.path.join(GOLDEN, "elliptical_nominal_nonuniform60.npz"))
    st, tf, it, traj = _solve(hostsim, NOMINAL, nt=60, time=g["time"])
    assert st == 0 and abs(tf - float(g["tf"])) / float(g["tf"]) < 1e-8
    assert _rel(traj, g["traj"])[:9].max() < 1e-4
    # the reference's objective with DCOST (8-state sweeps)
    g = np.load(os.path.join(GOLDEN, "elliptical_dcost1e-5_disp4_seed11_nt200.npz"))
    monkeypatch.setenv("WDC", repr(10.0 * 1e-5 / 199))
    for b in range(2):
        st, tf, it, traj = _solve(hostsim, g["rows"][:, b])
        assert st == 0 and abs(tf - g["tf"][b]) / g["tf"][b] < 1e-8
        err = _rel(traj, g["traj"][b])
        assert err[:9].max() < 1e-4 and err[9] < 2e-4, err
    monkeypatch.delenv("WDC")
    # config 2: the circular model (structural flag coup5 = 0)
    g = np.load(os.path.join(GOLDEN, "circular_nominal_nt200.npz"))
    monkeypatch.setenv("CIRCULAR", "1")
    raw = NOMINAL.copy(); raw[8] = raw[9] = 53108.4; raw[11] = 2576.0
    st, tf, it, traj = _solve(hostsim, raw)
    assert st == 0 and abs(tf - float(g["tf"])) / float(g["tf"]) < 1e-8


def test_cooperative_sweeps_source_matches_goldens_and_thread_sweeps(hostsim, monkeypatch):
    """csrc/ascent_coop.cuh run by one lane (G = 1: the group primitives are identities): the stored-model
    phases, the row-wise congruence with its symmetrisation, the forward / adjoint recursions.  Same optimum
    and the same number of iterations as the one-thread-per-problem sweeps, on every golden case."""
    monkeypatch.delenv("WDC", raising=False)
    monkeypatch.delenv("CIRCULAR", raising=False)
    monkeypatch.delenv("NPOL", raising=False)

    def both(raw, gold_tf, gold_traj=None, **kw):
        a = _solve(hostsim, raw, fn="hostsim_solve", **kw)
        b = _solve(hostsim, raw, fn="hostsim_solve_coop", **kw)
        assert a[0] == 0 and b[0] == 0
        assert abs(b[1] - gold_tf) / gold_tf < 1e-8 and abs(a[1] - b[1]) / a[1] < 1e-11
        assert abs(a[2] - b[2]) <= 1, (a[2], b[2])                    # iterations
        if gold_traj is not None:
            err = _rel(b[3], gold_traj)
            assert err[:9].max() < 1e-4 and err[9] < 2e-4, err
        assert _rel(b[3], a[3] + 0.0)[:9].max() < 1e-5

    g = np.load(os.path.join(GOLDEN, "elliptical_nominal_nt200.npz"))
    both(NOMINAL, float(g["tf"]), g["traj"])
    g = np.load(os.path.join(GOLDEN, "elliptical_nominal_nt40.npz"))
    both(NOMINAL, float(g["tf"]), g["traj"], nt=40)
    g = np.load(os.path.join(GOLDEN, "elliptical_nominal_nonuniform60.npz"))
    both(NOMINAL, float(g["tf"]), g["traj"], nt=60, time=g["time"])
    g = np.load(os.path.join(GOLDEN, "elliptical_dcost1e-5_disp4_seed11_nt200.npz"))
    monkeypatch.setenv("WDC", repr(10.0 * 1e-5 / 199))
    monkeypatch.setenv("NPOL", "2")
    for b in range(4):
        both(g["rows"][:, b], g["tf"][b], g["traj"][b])
    monkeypatch.delenv("WDC"); monkeypatch.delenv("NPOL")
    g = np.load(os.path.join(GOLDEN, "circular_nominal_nt200.npz"))
    monkeypatch.setenv("CIRCULAR", "1")
    raw = NOMINAL.copy(); raw[8] = raw[9] = 53108.4; raw[11] = 2576.0
    both(raw, float(g["tf"]))
    # the circular model WITH its move suppression (PDF p.27 src 69-73: DCOST = 1e-5 on the MV `angle`): the MV slot
    # of the 8-state cooperative sweeps holds the pitch angle; the oracle carries the term as slack pairs on the MV
    g = np.load(os.path.join(GOLDEN, "circular_dcost1e-5_nt200.npz"))
    monkeypatch.setenv("WDC", repr(10.0 * 1e-5 / 199)); monkeypatch.setenv("NPOL", "2")
    st, tf, it, traj = _solve(hostsim, raw, fn="hostsim_solve_coop")
    assert st == 0 and abs(tf - float(g["tf"])) / float(g["tf"]) < 1e-10
    names = list(g["names"])
    for n in names:
        ref = g["traj"][names.index(n)]
        mine = traj[["y", "ydot", "ydoubledot", "x", "xdot", "xdoubledot", "angle", "angledot", "mass"].index(n)]
        assert np.abs(mine - ref).max() / np.abs(ref).max() < 1e-7, n
    assert np.array_equal(traj[9], traj[6]) and np.abs(traj[7]).max() == 0.0     # MV slot == angle, angledot pinned at 0


def test_higher_order_collocation_source_matches_goldens(hostsim):
    """csrc/ascent_colloc.cuh (NODES = 3..6, LO:25) run by one lane against the oracle's fixtures: per-step LU
    condensation, dense Riccati on the coupling unknowns, multipliers from the stored factors.  The collocation
    rule comes from the C ABI (lmato_collocation_rule) and must equal the oracle's (APMonitor's tables)."""
    from lunar_module_ascent_trajectory_optimiser_b200 import _cabi
    from oracle.ascent_nlp import collocation_matrix
    g = np.load(os.path.join(GOLDEN, "elliptical_higher_order_nodes3to6.npz"))
    fn = hostsim.hostsim_solve_colloc

    def solve(raw, nt, nodes):
        m = nodes - 1
        tau = np.zeros(m); Nc = np.zeros((m, m))
        assert _cabi.lib().lmato_collocation_rule(nodes, tau.ctypes.data_as(C.c_void_p), Nc.ctypes.data_as(C.c_void_p)) == 0
        t_or, N_or = collocation_matrix(nodes)
        assert np.abs(tau - t_or[1:]).max() < 1e-14 and np.abs(Nc - N_or).max() < 1e-12
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        traj = np.empty((10, nt)); tf, it, kkt = C.c_double(), C.c_int(), C.c_double()
        st = fn(raw.ctypes.data_as(C.c_void_p), nt, None, nodes, Nc.ctypes.data_as(C.c_void_p), tau.ctypes.data_as(C.c_void_p),
                C.c_double(1e-10), C.c_double(10.0), C.c_double(1e-3), traj.ctypes.data_as(C.c_void_p), C.byref(tf),
                C.byref(it), C.byref(kkt))
        return st, tf.value, it.value, traj

    for nodes in (3, 4, 5, 6):
        nt = int(g[f"nt_n{nodes}"])
        st, tf, it, traj = solve(NOMINAL, nt, nodes)
        assert st == 0 and abs(tf - float(g[f"tf_n{nodes}"])) / tf < 1e-9, (nodes, st, tf)
        err = _rel(traj, g[f"traj_n{nodes}"])
        assert err[:9].max() < 1e-6 and err[9] < 1e-4, (nodes, err)
    for b in range(3):
        st, tf, it, traj = solve(g["rows"][:, b], int(g["disp_nt"]), 3)
        assert st == 0 and abs(tf - g["disp_tf"][b]) / tf < 1e-9
        assert _rel(traj, g["disp_traj"][b])[:9].max() < 1e-5
