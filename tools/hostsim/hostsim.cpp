// DEVELOPER TOOL -- NOT PART OF THE PRODUCT.
//
// Compiles the device solver core (csrc/ascent_ipm.cuh) with g++ so that the IPM control
// logic can be stepped through on a box without a GPU.  It is not built by
// __graft_entry__.build(), not loaded by the Python package, not used by any test as the
// thing under test, and bench.py never calls it.  The product path is the CUDA library
// only and fails loudly without it.
//
//   g++ -O2 -std=c++17 -I lunar_module_ascent_trajectory_optimiser_b200/csrc tools/hostsim/hostsim.cpp -o /tmp/hostsim
//   /tmp/hostsim [nt] [seed dispersions...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include "ascent_ipm_dc.cuh"

using namespace lmato;

static Params make_params(double Ft, double M0, double Mdot, double addm, double rp, double ra) {
  Params P;
  const double G = 6.674e-11, Mm = 7.346e22, R0 = 1738100.0, fuel = 2376.0, T = 470.0;
  P.GM = G * Mm; P.R0 = R0; P.Ft = Ft; P.M0 = M0; P.S = rp; P.ms = fuel; P.mflow = Mdot / fuel;
  P.asc = addm / 3.0; P.T = T; P.a_ub = M_PI / 3.0; P.u_ub = 1.0;
  const double vt = std::sqrt(P.GM / (R0 + 0.5 * (rp + ra)));
  P.vt2 = (vt / P.S) * (vt / P.S);
  P.rt = (R0 + P.S) / P.S; P.R0S = R0 / P.S;
  P.tf_ub = std::fmin(1.0, 1.0 / (P.mflow * T));
  P.fuel = fuel; P.Sinv = 1.0 / P.S; P.coup5 = 1.0; P.mT = P.mflow * P.T;
  return P;
}

int main(int argc, char** argv) {
  int nt = argc > 1 ? atoi(argv[1]) : 200;
  int nprob = argc > 2 ? atoi(argv[2]) : 1;
  int verbose = argc > 3 ? atoi(argv[3]) : 0;
  int only = argc > 4 ? atoi(argv[4]) : -1;
  const int N = nt - 1;
  std::vector<double> h(N + 1), tau(N + 1);
  for (int k = 0; k <= N; ++k) { tau[k] = (double)k / N; h[k] = k ? tau[k] - tau[k - 1] : 0.0; }
  Mesh M{N, h.data(), tau.data()};
  Options O;
  O.tol = 1e-10; O.mu_init = 0.1; O.obj_scale = 10.0; O.kappa_eps = 30.0; O.kappa_mu = 0.2; O.theta_mu = 1.5; O.theta_mu_warm = 2.0;
  O.tau_min = 0.99; O.delta_c = 1e-8; O.tf_guess = 0.9; O.max_iter = 500; O.max_ls = 40; O.mu_min_factor = 1e-3; O.n_polish = 4;
  if (getenv("MMF")) O.mu_min_factor = atof(getenv("MMF"));
  if (getenv("NPOL")) O.n_polish = atoi(getenv("NPOL"));
  O.w_dcost = getenv("WDC") ? atof(getenv("WDC")) : 0.0;
  if (getenv("THMU")) O.theta_mu = atof(getenv("THMU"));
  if (getenv("KMU")) O.kappa_mu = atof(getenv("KMU"));
  if (getenv("KEPS")) O.kappa_eps = atof(getenv("KEPS"));
  if (getenv("OBJ")) O.obj_scale = atof(getenv("OBJ"));
  if (getenv("MU0")) O.mu_init = atof(getenv("MU0"));
  if (getenv("TOL")) O.tol = atof(getenv("TOL"));
  if (getenv("TF0")) O.tf_guess = atof(getenv("TF0"));
  if (getenv("DC")) O.delta_c = atof(getenv("DC"));
  const int nfields = O.w_dcost > 0 ? (int)dc::N_FIELDS : (int)N_FIELDS;
  std::vector<double> ws((size_t)nfields * LANES * (N + 1), 0.0);   // one warp block, lane 0 used
  std::vector<double> tile(2 * TILE_ROWS, 0.0);
  Ws W{ws.data(), (long)nfields * LANES, tile.data()};
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  int nfail = 0, itsum = 0, itmax = 0;
  for (int p = 0; p < nprob; ++p) {
    double u[6];
    for (int i = 0; i < 6; ++i) u[i] = (p == 0) ? 0.5 : U(rng);
    const double Ft = 15346.0 * (1 + 0.02 * (2 * u[0] - 1));
    const double Isp = 309.7 * (1 + 0.01 * (2 * u[1] - 1));
    const double Mdot = (p == 0) ? 5.053 : Ft / (Isp * 9.807);
    const double M0 = 4821.0 * (1 + 0.02 * (2 * u[2] - 1));
    const double addm = 5e-4 * std::pow(2.0, 2 * u[3] - 1);
    const double rp = 17703.0 * (1 + 0.10 * (2 * u[4] - 1));
    const double ra = 88615.0 * (1 + 0.10 * (2 * u[5] - 1));
    if (only >= 0 && p != only) continue;
    Params P = make_params(Ft, M0, Mdot, addm, rp, ra);
    std::fill(ws.begin(), ws.end(), 0.0);
    SolveOut out;
    if (O.w_dcost > 0) {
      IpmState S;
      dc::init_guess(P, M, O, W, S.cur);
      ipm_begin(O, S);
      while (!ipm_iterate_t<Sweeps8>(P, M, O, W, S)) {}
      ipm_result(S, out);
    } else {
      ipm_solve(P, M, O, W, false, out);
    }
    itsum += out.iters; if (out.iters > itmax) itmax = out.iters;
    if (out.status != 0) ++nfail;
    if (verbose || nprob <= 20 || out.status != 0)
      printf("p %4d status %d iters %3d kkt %.2e mu %.1e tf %.10f tf_s %.8f  (Ft %.1f M0 %.1f Mdot %.4f addm %.2e rp %.0f ra %.0f)\n",
             p, out.status, out.iters, out.kkt, out.mu, out.tf, out.tf * P.T, Ft, M0, Mdot, addm, rp, ra);
    if (p == 0 && nprob == 1) {
      const int b = out.cur;
      const double* sp = W.stage(N);
      printf("final y %.9f x %.9f vy %.9f vx %.9f\n", WS_AT(sp, b * N_ITER + F_Z + 0) * P.S, WS_AT(sp, b * N_ITER + F_Z + 2) * P.S,
             WS_AT(sp, b * N_ITER + F_Z + 1) * P.S, WS_AT(sp, b * N_ITER + F_Z + 3) * P.S);
      printf("final mass %.9f\n", P.M0 - P.fuel * P.mflow * P.T * out.tf);
    }
  }
  printf("solved %d problems: %d failures, mean iters %.1f, max %d\n", nprob, nfail, (double)itsum / nprob, itmax);
  return 0;
}

// ctypes entry for developer-side comparisons against tests/golden (NOT the product path).
extern "C" int hostsim_solve(const double* raw14, int nt, const double* time, double tol, double obj_scale,
                             double mu_min_factor, double* traj /* [10][nt] */, double* tf_out, int* iters, double* kkt) {
  const int N = nt - 1;
  std::vector<double> h(N + 1), tau(N + 1);
  for (int k = 0; k <= N; ++k) { tau[k] = time ? time[k] : (double)k / N; h[k] = k ? tau[k] - tau[k - 1] : 0.0; }
  Mesh M{N, h.data(), tau.data()};
  Options O;
  O.tol = tol; O.mu_init = 0.1; O.obj_scale = obj_scale; O.kappa_eps = 30.0; O.kappa_mu = 0.2; O.theta_mu = 1.5; O.theta_mu_warm = 2.0;
  O.tau_min = 0.99; O.delta_c = 1e-8; O.tf_guess = 0.9; O.max_iter = 500; O.max_ls = 40; O.mu_min_factor = mu_min_factor; O.n_polish = getenv("NPOL") ? atoi(getenv("NPOL")) : 4;
  O.w_dcost = getenv("WDC") ? atof(getenv("WDC")) : 0.0;
  Params P;
  const double* r = raw14;
  P.GM = r[0] * r[1]; P.R0 = r[2]; P.Ft = r[3]; P.M0 = r[4]; P.S = r[8]; P.ms = r[11]; P.mflow = r[5] / r[6];
  P.asc = r[7] / 3.0; P.T = r[10]; P.a_ub = r[12]; P.u_ub = r[13];
  const double vt = std::sqrt(P.GM / (P.R0 + 0.5 * (r[8] + r[9])));
  P.vt2 = (vt / P.S) * (vt / P.S); P.rt = (P.R0 + P.S) / P.S; P.R0S = P.R0 / P.S;
  P.tf_ub = std::fmin(1.0, 1.0 / (P.mflow * P.T)); P.fuel = r[6]; P.Sinv = 1.0 / P.S; P.coup5 = 1.0; P.mT = P.mflow * P.T;
  if (getenv("CIRCULAR")) { P.coup5 = 0.0; P.asc = 1.0; P.u_ub = 1e20; }
  const bool DC = O.w_dcost > 0;
  const int nfields = DC ? (int)dc::N_FIELDS : (int)N_FIELDS;
  const int niter = DC ? (int)dc::N_ITER : (int)N_ITER;
  std::vector<double> ws((size_t)nfields * LANES * (N + 1), 0.0);
  std::vector<double> tile(2 * TILE_ROWS, 0.0);
  Ws W{ws.data(), (long)nfields * LANES, tile.data()};
  SolveOut out;
  if (DC) {
    IpmState S;
    dc::init_guess(P, M, O, W, S.cur);
    ipm_begin(O, S);
    while (!ipm_iterate_t<Sweeps8>(P, M, O, W, S)) {}
    ipm_result(S, out);
  } else {
    ipm_solve(P, M, O, W, false, out);
  }
  *tf_out = out.tf; *iters = out.iters; *kkt = out.kkt;
  for (int v = 0; v < 10; ++v) traj[v * nt] = 0.0;
  for (int k = 1; k <= N; ++k) {
    const double* sp = W.stage(k);
    double z[6];
    for (int i = 0; i < 6; ++i) z[i] = WS_AT(sp, out.cur * niter + F_Z + i);
    const double m = P.mflow * P.T * tau[k] * out.tf;
    double ay, ax;
    accel_value(P, z[0], z[2], z[4], m, ay, ax);
    const double vals[10] = {z[0], z[1], ay, z[2], z[3], ax, z[4], z[5], m, WS_AT(sp, out.cur * niter + F_U)};
    for (int v = 0; v < 10; ++v) traj[v * nt + k] = vals[v];
  }
  return out.status;
}

// Finite-difference check of the hand-derived model derivatives (ascent_model.cuh) at one point:
// returns the largest relative error of accel_first's eight first derivatives and of the
// second-derivative contraction used by the Hessian (accel_second) against central differences.
extern "C" double hostsim_check_derivatives(const double* raw14, double y, double x, double a, double m) {
  Params P;
  const double* r = raw14;
  P.GM = r[0] * r[1]; P.R0 = r[2]; P.Ft = r[3]; P.M0 = r[4]; P.S = r[8]; P.ms = r[11]; P.mflow = r[5] / r[6];
  P.asc = r[7] / 3.0; P.T = r[10]; P.a_ub = r[12]; P.u_ub = r[13];
  P.fuel = r[6]; P.Sinv = 1.0 / P.S; P.coup5 = 1.0; P.mT = P.mflow * P.T;
  Accel1 f;
  accel_first(P, y, x, a, m, f);
  const double v[4] = {y, x, a, m};
  const double an1[2][4] = {{f.ay_y, f.ay_x, f.ay_a, f.ay_m}, {f.ax_y, f.ax_x, f.ax_a, f.ax_m}};
  double worst = 0.0;
  Accel1 fp[4], fm[4];
  double hs[4];
  for (int j = 0; j < 4; ++j) {
    double p[4] = {v[0], v[1], v[2], v[3]}, q[4] = {v[0], v[1], v[2], v[3]};
    hs[j] = 1e-6 * std::fmax(1.0, std::fabs(v[j]));
    p[j] += hs[j]; q[j] -= hs[j];
    accel_first(P, p[0], p[1], p[2], p[3], fp[j]);
    accel_first(P, q[0], q[1], q[2], q[3], fm[j]);
    const double fd[2] = {(fp[j].ay - fm[j].ay) / (2 * hs[j]), (fp[j].ax - fm[j].ax) / (2 * hs[j])};
    for (int i = 0; i < 2; ++i)
      worst = std::fmax(worst, std::fabs(fd[i] - an1[i][j]) / std::fmax(1e-12, std::fabs(an1[i][j])));
  }
  // second derivatives: H = wy * d2 ay + wx * d2 ax over (y, x, a, m), against differences of the first ones
  const double wy = 0.7, wx = -1.3;
  Accel2 h2;
  accel_second(P, f, wy, wx, h2);
  const double an2[4][4] = {{h2.yy, h2.yx, h2.ya, h2.ym}, {h2.yx, h2.xx, h2.xa, h2.xm},
                            {h2.ya, h2.xa, h2.aa, h2.am}, {h2.ym, h2.xm, h2.am, h2.mm}};
  for (int j = 0; j < 4; ++j) {
    const double gp[4] = {wy * fp[j].ay_y + wx * fp[j].ax_y, wy * fp[j].ay_x + wx * fp[j].ax_x,
                          wy * fp[j].ay_a + wx * fp[j].ax_a, wy * fp[j].ay_m + wx * fp[j].ax_m};
    const double gm[4] = {wy * fm[j].ay_y + wx * fm[j].ax_y, wy * fm[j].ay_x + wx * fm[j].ax_x,
                          wy * fm[j].ay_a + wx * fm[j].ax_a, wy * fm[j].ay_m + wx * fm[j].ax_m};
    for (int i = 0; i < 4; ++i) {
      const double fd = (gp[i] - gm[i]) / (2 * hs[j]);
      double scale = 0.0;
      for (int a2 = 0; a2 < 4; ++a2) scale = std::fmax(scale, std::fabs(an2[i][a2]));
      worst = std::fmax(worst, std::fabs(fd - an2[i][j]) / std::fmax(1e-12, scale));
    }
  }
  return worst;
}

// The cooperative sweeps (csrc/ascent_coop.cuh) run by one lane (G = 1): same entry as hostsim_solve.
#include "ascent_coop.cuh"
extern "C" int hostsim_solve_coop(const double* raw14, int nt, const double* time, double tol, double obj_scale,
                                  double mu_min_factor, double* traj /* [10][nt] */, double* tf_out, int* iters, double* kkt) {
  const int N = nt - 1;
  std::vector<double> h(N + 1), tau(N + 1);
  for (int k = 0; k <= N; ++k) { tau[k] = time ? time[k] : (double)k / N; h[k] = k ? tau[k] - tau[k - 1] : 0.0; }
  Mesh M{N, h.data(), tau.data()};
  Options O;
  O.tol = tol; O.mu_init = 0.1; O.obj_scale = obj_scale; O.kappa_eps = 30.0; O.kappa_mu = 0.2; O.theta_mu = 1.5; O.theta_mu_warm = 2.0;
  O.tau_min = 0.99; O.delta_c = 1e-8; O.tf_guess = 0.9; O.max_iter = 500; O.max_ls = 40; O.mu_min_factor = mu_min_factor; O.n_polish = getenv("NPOL") ? atoi(getenv("NPOL")) : 4;
  O.w_dcost = getenv("WDC") ? atof(getenv("WDC")) : 0.0;
  Params P;
  const double* r = raw14;
  P.GM = r[0] * r[1]; P.R0 = r[2]; P.Ft = r[3]; P.M0 = r[4]; P.S = r[8]; P.ms = r[11]; P.mflow = r[5] / r[6];
  P.asc = r[7] / 3.0; P.T = r[10]; P.a_ub = r[12]; P.u_ub = r[13];
  const double vt = std::sqrt(P.GM / (P.R0 + 0.5 * (r[8] + r[9])));
  P.vt2 = (vt / P.S) * (vt / P.S); P.rt = (P.R0 + P.S) / P.S; P.R0S = P.R0 / P.S;
  P.tf_ub = std::fmin(1.0, 1.0 / (P.mflow * P.T)); P.fuel = r[6]; P.Sinv = 1.0 / P.S; P.coup5 = 1.0; P.mT = P.mflow * P.T;
  if (getenv("CIRCULAR")) { P.coup5 = 0.0; P.asc = 1.0; P.u_ub = 1e20; }
  const bool DC = O.w_dcost > 0;
  if (getenv("CIRCULAR") && DC) P.asc = 0.0;       // the circular model with its move term: the MV slot is the pitch angle
  std::vector<double> ws((size_t)coop::coop_doubles_per_problem(nt), 0.0);
  std::vector<double> scr(coop::SCR_DOUBLES, 0.0);
  coop::Cws W{ws.data(), nt, scr.data(), 0, 1u, 1u, 0.0, 0.0, 0, 0u};
  IpmState S;
  ipm_begin(O, S);
  const bool vrec = getenv("VREC") != nullptr;     // the variant that keeps W_k, g_k instead of running the adjoint recursion
  if (DC && getenv("CIRCULAR")) { SweepsCoop<1, 1, 2, false>::guess(P, M, O, W, S.cur); while (!ipm_iterate_t<SweepsCoop<1, 1, 2, false>>(P, M, O, W, S)) {} }
  else if (DC && vrec)  { SweepsCoop<1, 1, true, true>::guess(P, M, O, W, S.cur); while (!ipm_iterate_t<SweepsCoop<1, 1, true, true>>(P, M, O, W, S)) {} }
  else if (DC)     { SweepsCoop<1, 1, true, false>::guess(P, M, O, W, S.cur); while (!ipm_iterate_t<SweepsCoop<1, 1, true, false>>(P, M, O, W, S)) {} }
  else if (vrec)   { SweepsCoop<1, 1, false, true>::guess(P, M, O, W, S.cur); while (!ipm_iterate_t<SweepsCoop<1, 1, false, true>>(P, M, O, W, S)) {} }
  else             { SweepsCoop<1, 1, false, false>::guess(P, M, O, W, S.cur); while (!ipm_iterate_t<SweepsCoop<1, 1, false, false>>(P, M, O, W, S)) {} }
  SolveOut out;
  ipm_result(S, out);
  *tf_out = out.tf; *iters = out.iters; *kkt = out.kkt;
  for (int v = 0; v < 10; ++v) traj[v * nt] = 0.0;
  for (int k = 1; k <= N; ++k) {
    const double* x = W.X(out.cur, k);
    const double m = P.mflow * P.T * tau[k] * out.tf;
    double ay, ax;
    accel_value(P, x[0], x[2], x[4], m, ay, ax);
    const double vals[10] = {x[0], x[1], ay, x[2], x[3], ax, x[4], x[5], m, x[coop::X_U]};
    for (int v = 0; v < 10; ++v) traj[v * nt + k] = vals[v];
  }
  return out.status;
}

// Higher-order collocation (csrc/ascent_colloc.cuh), GP = 1.  Ncol: row-major m x m, tauc: m points.
#include "ascent_colloc.cuh"
extern "C" int hostsim_solve_colloc(const double* raw14, int nt, const double* time, int nodes, const double* Ncol,
                                    const double* tauc, double tol, double obj_scale, double mu_min_factor,
                                    double* traj /* [10][nt] */, double* tf_out, int* iters, double* kkt) {
  const int N = nt - 1, m = nodes - 1;
  std::vector<double> h(N + 1), tau(N + 1);
  for (int k = 0; k <= N; ++k) { tau[k] = time ? time[k] : (double)k / N; h[k] = k ? tau[k] - tau[k - 1] : 0.0; }
  Mesh M{N, h.data(), tau.data()};
  Options O;
  O.tol = tol; O.mu_init = 0.1; O.obj_scale = obj_scale; O.kappa_eps = 30.0; O.kappa_mu = 0.2; O.theta_mu = 1.5; O.theta_mu_warm = 2.0;
  O.tau_min = 0.99; O.delta_c = 1e-8; O.tf_guess = 0.9; O.max_iter = 500; O.max_ls = 40; O.mu_min_factor = mu_min_factor; O.n_polish = getenv("NPOL") ? atoi(getenv("NPOL")) : 4;
  O.w_dcost = 0.0;
  Params P;
  const double* r = raw14;
  P.GM = r[0] * r[1]; P.R0 = r[2]; P.Ft = r[3]; P.M0 = r[4]; P.S = r[8]; P.ms = r[11]; P.mflow = r[5] / r[6];
  P.asc = r[7] / 3.0; P.T = r[10]; P.a_ub = r[12]; P.u_ub = r[13];
  const double vt = std::sqrt(P.GM / (P.R0 + 0.5 * (r[8] + r[9])));
  P.vt2 = (vt / P.S) * (vt / P.S); P.rt = (P.R0 + P.S) / P.S; P.R0S = P.R0 / P.S;
  P.tf_ub = std::fmin(1.0, 1.0 / (P.mflow * P.T)); P.fuel = r[6]; P.Sinv = 1.0 / P.S; P.coup5 = 1.0; P.mT = P.mflow * P.T;
  colloc::Coll C;
  C.m = m;
  for (int i = 0; i < m; ++i) { C.tau[i] = tauc[i]; for (int j = 0; j < m; ++j) C.N[i][j] = Ncol[i * m + j]; }
  std::vector<double> ws((size_t)colloc::colloc_doubles_per_problem(nt, m), 0.0);
  colloc::Nws W;
  W.base = ws.data(); W.N1 = nt; W.L.init(m); W.C = &C; W.g = 0; W.mask = 1u; W.dw = 0.0; W.pimax = 0.0; W.ls_flag = 0;
  IpmState S;
  SolveOut out;
  const int v0 = getenv("VARIANT") ? atoi(getenv("VARIANT")) : 0;
  const int nv = getenv("VARIANT") ? v0 + 1 : (int)SweepsColloc<1>::N_STARTS;
  for (int variant = v0; variant < nv; ++variant) {        // the kernel's restart ladder
    ipm_begin(O, S);
    SweepsColloc<1>::guess_variant(P, M, O, W, S.cur, variant);
    while (!ipm_iterate_t<SweepsColloc<1>>(P, M, O, W, S)) {}
    ipm_result(S, out);
    if (out.status == ST_CONVERGED) break;
  }
  *tf_out = out.tf; *iters = out.iters; *kkt = out.kkt;
  for (int v = 0; v < 10; ++v) traj[v * nt] = 0.0;
  for (int k = 1; k <= N; ++k) {
    const double* x = W.X(out.cur, k) + 6 * (m - 1);
    const double ms = P.mflow * P.T * tau[k] * out.tf;
    double ay, ax;
    accel_value(P, x[0], x[2], x[4], ms, ay, ax);
    const double vals[10] = {x[0], x[1], ay, x[2], x[3], ax, x[4], x[5], ms, W.X(out.cur, k)[W.L.x_u]};
    for (int v = 0; v < 10; ++v) traj[v * nt + k] = vals[v];
  }
  return out.status;
}

// Batch warm start on the host (what lmato_solve_batch does on the device): the batch-mean problem is solved to
// mu_ref by the cooperative sweeps without the move term and stored as the reference; every dispersed problem
// then starts from it.  Returns the mean iteration count; used to study barrier-update strategies on the CPU.
//   hostsim_warm_batch(nprob, nt, wdc, mu_ref, out_iters[nprob], out_status[nprob], out_tf[nprob])
extern "C" double hostsim_warm_batch(int nprob, int nt, double wdc, double mu_ref, int* out_iters, int* out_status, double* out_tf) {
  const int N = nt - 1;
  std::vector<double> h(N + 1), tau(N + 1);
  for (int k = 0; k <= N; ++k) { tau[k] = (double)k / N; h[k] = k ? tau[k] - tau[k - 1] : 0.0; }
  Mesh M{N, h.data(), tau.data()};
  Options O;
  O.tol = 1e-10; O.mu_init = 0.1; O.obj_scale = 10.0; O.kappa_eps = 30.0; O.kappa_mu = 0.2; O.theta_mu = 1.5; O.theta_mu_warm = 2.0;
  O.tau_min = 0.99; O.delta_c = 1e-8; O.tf_guess = 0.9; O.max_iter = 500; O.max_ls = 40; O.mu_min_factor = 1e-3; O.n_polish = 2;
  O.w_dcost = wdc;
  if (getenv("THMUW")) O.theta_mu_warm = atof(getenv("THMUW"));
  if (getenv("KEPS")) O.kappa_eps = atof(getenv("KEPS"));
  if (getenv("KMU")) O.kappa_mu = atof(getenv("KMU"));
  if (getenv("NPOL")) O.n_polish = atoi(getenv("NPOL"));
  if (getenv("MMF")) O.mu_min_factor = atof(getenv("MMF"));
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::vector<Params> Ps;
  double mean[6] = {0, 0, 0, 0, 0, 0};
  for (int p = 0; p < nprob; ++p) {
    double u[6];
    for (int i = 0; i < 6; ++i) u[i] = (p == 0) ? 0.5 : U(rng);
    const double Ft = 15346.0 * (1 + 0.02 * (2 * u[0] - 1));
    const double Isp = 309.7 * (1 + 0.01 * (2 * u[1] - 1));
    const double Mdot = (p == 0) ? 5.053 : Ft / (Isp * 9.807);
    const double M0 = 4821.0 * (1 + 0.02 * (2 * u[2] - 1));
    const double addm = 5e-4 * std::pow(2.0, 2 * u[3] - 1);
    const double rp = 17703.0 * (1 + 0.10 * (2 * u[4] - 1));
    const double ra = 88615.0 * (1 + 0.10 * (2 * u[5] - 1));
    Ps.push_back(make_params(Ft, M0, Mdot, addm, rp, ra));
    const double v[6] = {Ft, M0, Mdot, addm, rp, ra};
    for (int i = 0; i < 6; ++i) mean[i] += v[i] / nprob;
  }
  std::vector<double> ws((size_t)coop::coop_doubles_per_problem(nt), 0.0), scr(coop::SCR_DOUBLES, 0.0);
  std::vector<double> ref((size_t)dc::REF_ROWS * nt, 0.0);
  coop::Cws W{ws.data(), nt, scr.data(), 0, 1u, 1u, 0.0, 0.0, 0, 0u};
  {   // reference solve
    Params Pm = make_params(mean[0], mean[1], mean[2], mean[3], mean[4], mean[5]);
    Options R = O;
    R.tol = 10.0 * mu_ref; R.mu_min_factor = 0.1; R.n_polish = 0; R.w_dcost = 0.0;
    IpmState S;
    ipm_begin(R, S);
    SweepsCoop<1, 1, false, false>::guess(Pm, M, R, W, S.cur);
    while (!ipm_iterate_t<SweepsCoop<1, 1, false, false>>(Pm, M, R, W, S)) {}
    SolveOut out; ipm_result(S, out);
    SweepsCoop<1, 1, false, false>::store_ref(Pm, M, W, out.cur, S.cur, S.ctl.mu, out.status == ST_CONVERGED, ref.data());
  }
  double itsum = 0;
  for (int p = 0; p < nprob; ++p) {
    IpmState S;
    ipm_begin(O, S);
    double mu0 = 0.0;
    SolveOut out;
    if (wdc > 0) {
      SweepsCoop<1, 1, true, false>::load_ref(Ps[p], M, O, W, ref.data(), S.cur, &mu0);
      S.warm = true; S.ctl.mu = mu0; S.ctl.tau = dmax(O.tau_min, 1.0 - mu0);
      while (!ipm_iterate_t<SweepsCoop<1, 1, true, false>>(Ps[p], M, O, W, S)) {}
    } else {
      SweepsCoop<1, 1, false, false>::load_ref(Ps[p], M, O, W, ref.data(), S.cur, &mu0);
      S.warm = true; S.ctl.mu = mu0; S.ctl.tau = dmax(O.tau_min, 1.0 - mu0);
      while (!ipm_iterate_t<SweepsCoop<1, 1, false, false>>(Ps[p], M, O, W, S)) {}
    }
    ipm_result(S, out);
    out_iters[p] = out.iters; out_status[p] = out.status; out_tf[p] = out.tf;
    itsum += out.iters;
  }
  return itsum / nprob;
}
