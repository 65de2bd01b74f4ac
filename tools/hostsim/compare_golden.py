"""Developer tool: run the g++ build of the device solver core on the golden dispersion inputs
and report deviations from tests/golden (no GPU needed).  Not part of the product or the tests."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = "/tmp/libhostsim.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I",
                       os.path.join(ROOT, "lunar_module_ascent_trajectory_optimiser_b200", "csrc"),
                       os.path.join(ROOT, "tools", "hostsim", "hostsim.cpp"), "-o", so])
L = C.CDLL(so)
tol = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-10
mmf = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
g = np.load(os.path.join(ROOT, "tests", "golden", "elliptical_dispersions8_seed11_nt200.npz"))
nt = 200
worst = np.zeros(10)
for b in range(8):
    raw = np.ascontiguousarray(g["rows"][:, b])
    traj = np.empty((10, nt)); tf = C.c_double(); it = C.c_int(); kkt = C.c_double()
    st = L.hostsim_solve(raw.ctypes.data_as(C.c_void_p), nt, None, C.c_double(tol), C.c_double(10.0), C.c_double(mmf),
                         traj.ctypes.data_as(C.c_void_p), C.byref(tf), C.byref(it), C.byref(kkt))
    gt = g["traj"][b]
    err = (np.abs(traj - gt) / np.abs(gt).max(axis=1, keepdims=True)).max(axis=1)
    worst = np.maximum(worst, err)
    print(b, "status", st, "iters", it.value, "kkt %.1e" % kkt.value, "tf rel %.1e" % (abs(tf.value - g["tf"][b]) / g["tf"][b]),
          "angledot %.1e control %.1e others %.1e" % (err[7], err[9], np.delete(err, [7, 9]).max()))
print("worst per row:", " ".join("%.1e" % e for e in worst))

# circular model (config 2): same sweeps, coup5 = 0
os.environ["CIRCULAR"] = "1"
g = np.load(os.path.join(ROOT, "tests", "golden", "circular_nominal_nt200.npz"))
raw = np.array([6.674e-11, 7.346e22, 1738100.0, 15346.0, 4821.0, 5.053, 2376.0, 5e-4, 53108.4, 53108.4, 470.0, 2576.0,
                np.pi / 3, 1.0])
traj = np.empty((10, nt)); tf = C.c_double(); it = C.c_int(); kkt = C.c_double()
st = L.hostsim_solve(raw.ctypes.data_as(C.c_void_p), nt, None, C.c_double(tol), C.c_double(10.0), C.c_double(mmf),
                     traj.ctypes.data_as(C.c_void_p), C.byref(tf), C.byref(it), C.byref(kkt))
names = list(g["names"])
rows = {"y": 0, "ydot": 1, "ydoubledot": 2, "x": 3, "xdot": 4, "xdoubledot": 5, "angle": 6, "mass": 8}
err = {n: np.abs(traj[rows[n]] - g["traj"][names.index(n)]).max() / np.abs(g["traj"][names.index(n)]).max() for n in rows}
print("circular: status", st, "iters", it.value, "kkt %.1e" % kkt.value, "tf_s %.8f (golden %.8f)" % (tf.value * 470, float(g["tf"]) * 470))
print("circular rel err:", {k: "%.1e" % v for k, v in err.items()})

# DCOST (8-state path): weight relative to obj_scale*tf = obj_scale*dcost/(nt-1)
os.environ.pop("CIRCULAR", None)
os.environ["WDC"] = repr(10.0 * 1e-5 / 199)
g = np.load(os.path.join(ROOT, "tests", "golden", "elliptical_dcost1e-5_disp4_seed11_nt200.npz"))
worst = np.zeros(10)
for b in range(g["rows"].shape[1]):
    raw = np.ascontiguousarray(g["rows"][:, b])
    traj = np.empty((10, nt)); tf = C.c_double(); it = C.c_int(); kkt = C.c_double()
    st = L.hostsim_solve(raw.ctypes.data_as(C.c_void_p), nt, None, C.c_double(tol), C.c_double(10.0), C.c_double(mmf),
                         traj.ctypes.data_as(C.c_void_p), C.byref(tf), C.byref(it), C.byref(kkt))
    gt = g["traj"][b]
    err = (np.abs(traj - gt) / np.abs(gt).max(axis=1, keepdims=True)).max(axis=1)
    worst = np.maximum(worst, err)
    print("dcost", b, "status", st, "iters", it.value, "kkt %.1e" % kkt.value, "tf rel %.1e" % (abs(tf.value - g["tf"][b]) / g["tf"][b]),
          "angledot %.1e control %.1e others %.1e" % (err[7], err[9], np.delete(err, [7, 9]).max()))
print("dcost worst per row:", " ".join("%.1e" % e for e in worst))
