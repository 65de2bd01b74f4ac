"""Host-side logic that needs no GPU: parameter packing, meshes, index sharding, and the
multi-rank gather (world_size 2 over gloo with a stub in place of the CUDA solve)."""
import math
import os

import pytest
import torch
import torch.multiprocessing as mp

import lunar_module_ascent_trajectory_optimiser_b200 as lm
from lunar_module_ascent_trajectory_optimiser_b200 import _cabi
from lunar_module_ascent_trajectory_optimiser_b200.api import package_solution


def test_defaults_are_the_reference_literals():
    p = lm.AscentParams()
    r = p.rows()
    assert r.shape == (14, 1) and r.dtype == torch.float64
    lit = [6.674e-11, 7.346e22, 1738100.0, 15346.0, 4821.0, 5.053, 2376.0, 5e-4, 17703.0, 88615.0, 470.0,
           2376.0, math.pi / 3, 1.0]                                      # LO:38, 50-71, 94, 96, 108
    assert r[:, 0].tolist() == lit
    c = lm.AscentParams.circular().rows()
    assert c[8, 0] == 53108.4 and c[9, 0] == 53108.4 and c[11, 0] == 2576.0   # PDF p.26 src 40, p.27 src 67


def test_batched_params_broadcast_and_isp():
    Ft = torch.tensor([15000.0, 15346.0, 16000.0], dtype=torch.float64)
    p = lm.AscentParams(Ft=Ft, Isp=torch.full((3,), 310.0, dtype=torch.float64))
    r = p.rows()
    assert r.shape == (14, 3)
    assert torch.allclose(r[5], Ft / (310.0 * 9.807))                     # PDF p.6
    assert torch.all(r[4] == 4821.0)
    with pytest.raises(ValueError):
        lm.AscentParams(Ft=Ft, M0=torch.ones(4, dtype=torch.float64)).rows()
    with pytest.raises(ValueError):
        lm.AscentParams(Ft=Ft).rows(5)


def test_dispersions_are_seeded_and_bounded():
    a = lm.dispersed_params(257, seed=11).rows()
    b = lm.dispersed_params(257, seed=11).rows()
    assert torch.equal(a, b)
    assert a[:, 0].tolist() == lm.AscentParams().rows()[:, 0].tolist() or abs(a[5, 0] - 5.053) < 1e-12
    assert float((a[3] / 15346 - 1).abs().max()) <= 0.02 + 1e-12
    assert float((a[8] / 17703 - 1).abs().max()) <= 0.10 + 1e-12
    assert 2.5e-4 - 1e-12 <= float(a[7].min()) and float(a[7].max()) <= 1e-3 + 1e-12
    c3 = lm.dispersed_params(64, columns=(0, 1, 2, 3)).rows()
    assert torch.all(c3[8] == 17703.0) and torch.all(c3[9] == 88615.0)


def test_mesh_grid():
    g = lm.Mesh().grid()
    assert g.shape == (200,) and g[0] == 0 and g[-1] == 1                 # LO:20-21
    assert torch.allclose(torch.diff(g), torch.full((199,), 1 / 199, dtype=torch.float64))
    g2 = lm.Mesh(time=[0, 0.1, 0.5, 1.0]).grid()
    assert g2.tolist() == [0, 0.1, 0.5, 1.0]


def test_shard_bounds_cover_and_balance():
    for B in (0, 1, 7, 8, 65536, 65537):
        for G in (1, 2, 3, 4, 8):
            spans = [lm.shard_bounds(B, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
            for i in range(B if B < 100 else 0):
                r = [k for k, (l, h) in enumerate(spans) if l <= i < h][0]
                assert r == (i * G) // B or spans[(i * G) // B][0] <= i   # problem i -> rank floor(i*G/B)


def _stub_solve(nt):
    def f(rows):
        B = rows.shape[1]
        tf = rows[3] * 1e-5 + rows[8] * 1e-6
        traj = torch.arange(10 * nt, dtype=torch.float64).reshape(10, nt, 1) + tf.reshape(1, 1, B)
        return {"traj": traj.contiguous(), "tf": tf.clone(), "final_mass": rows[4] - tf,
                "status": torch.zeros(B, dtype=torch.int32), "iterations": (rows[3] % 7).to(torch.int32),
                "kkt": tf * 1e-9}
    return f


def _worker(rank, world, port, B, nt, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = lm.dispersed_params(B, seed=3).rows()
    out = lm.sharded_solve(rows, _stub_solve(nt), None)
    ref = _stub_solve(nt)(rows)
    ok = all(torch.equal(out[k], ref[k]) for k in ref) and out["status"].dtype == torch.int32
    # scalars gathered, trajectories left sharded (what bench.py's multi-GPU step does)
    part = lm.sharded_solve(rows, _stub_solve(nt), None, gather_traj=False)
    lo, hi = lm.shard_bounds(B, world, rank)
    ok = ok and part["shard"] == (lo, hi) and all(torch.equal(part[k], ref[k]) for k in ref if k != "traj")
    ok = ok and torch.equal(part["traj"], ref["traj"][:, :, lo:hi])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 64])
def test_sharded_solve_gloo_world2(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 6, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_package_solution_shapes():
    nt, B = 5, 3
    raw = _stub_solve(nt)(lm.dispersed_params(B).rows())
    sol = package_solution(raw, lm.dispersed_params(B).rows(), torch.linspace(0, 1, nt, dtype=torch.float64))
    assert set(sol.states) == set(_cabi.VAR_ROWS) - {"angledoubledot"}
    assert sol.states["y"].shape == (B, nt) and sol.control.shape == (B, nt)
    assert torch.allclose(sol.tf_seconds, sol.tf * 470.0)
    assert len(sol) == B and bool(sol.converged.all())


def test_c_abi_default_options_and_kernel_ids():
    """lmato_default_options through ctypes (no GPU needed): the struct layouts agree and the defaults are the
    documented ones, incl. the reference's OTOL = RTOL = 1e-3 (LO:31-32) and MAX_ITER = 20000 (LO:28)."""
    import ctypes as C
    o = _cabi.LmatoOptions()
    _cabi.lib().lmato_default_options(C.byref(o))
    assert (o.tol, o.mu_init, o.obj_scale, o.tf_guess) == (1e-10, 0.1, 10.0, 0.9)
    assert (o.max_iter, o.max_ls, o.n_polish, o.warm_start) == (20000, 40, -1, 1)
    assert (o.dcost, o.kappa_eps, o.objective_nodes) == (1e-5, 30.0, 0)
    assert (o.kernel, o.coop_lanes, o.otol, o.rtol) == (0, 0, 1e-3, 1e-3)
    assert _cabi.KERNEL_IDS == {"auto": 0, "thread": 1, "coop": 2}
    s = lm.SolverOptions()
    assert (s.otol, s.rtol, s.max_iter, s.kernel) == (1e-3, 1e-3, 20000, "auto")


def test_circular_solution_names_its_control():
    nt, B = 5, 2
    raw = _stub_solve(nt)(lm.dispersed_params(B).rows())
    sol = package_solution(raw, lm.dispersed_params(B).rows(), torch.linspace(0, 1, nt, dtype=torch.float64), "circular")
    assert sol.control_name == "angle" and "angle" not in sol.states and "angledot" not in sol.states
    from lunar_module_ascent_trajectory_optimiser_b200 import gekko_shim
    g = gekko_shim.as_gekko(sol, 1)
    assert g.angle.value == sol.control[1].tolist() and not hasattr(g, "angledoubledot")
