// Batched primal-dual interior-point method for the ascent NLP, one problem per thread.
//
// Replaces, for this one model family, what the reference reaches through
// m.solve() (LO:177): APMonitor's collocation transcription + AD + IPOPT/MA27.
//   subsystem (1) transcription ......... stage defects below (backward Euler == GEKKO NODES=2, LO:25)
//   subsystem (2) residual/Jacobian/Hessian  ascent_model.cuh + stage_build()
//   subsystem (3) KKT factorisation ...... riccati_backward() / riccati_forward(): stage-wise
//                 block-tridiagonal LDL^T (Riccati recursion) with inertia read from the pivots
//   subsystem (4) barrier / fraction-to-boundary / filter line search: ipm_iterate()
//
// IPM = Waechter & Biegler, Math. Prog. 106 (2006) (the algorithm behind SOLVER=3, LO:26).
//
// Data layout: struct-of-arrays  ws[field][stage][slot]  with slot (= problem in flight)
// fastest, so that the 32 lanes of a warp touch 32 consecutive doubles (256 B) for every
// field of every stage.  All sweeps stream stage by stage through HBM.
#pragma once
#include "ascent_model.cuh"

namespace lmato {

// ---------------------------------------------------------------------------------------
// options / status
// ---------------------------------------------------------------------------------------
struct Options {
  double tol;            // scaled KKT error (IPOPT `tol`)
  double mu_init;        // 0.1
  double obj_scale;      // objective = obj_scale * tf
  double kappa_eps;      // 10
  double kappa_mu;       // 0.2
  double theta_mu;       // 1.5
  double tau_min;        // 0.99
  double delta_c;        // dual regularisation of the terminal equality row
  double tf_guess;       // initial tf (scaled, 0..1)
  int max_iter;          // LO:28 MAX_ITER
  int max_ls;            // max backtracking steps
};

enum Status : int {
  ST_CONVERGED = 0,
  ST_MAX_ITER = 1,
  ST_LINESEARCH_FAIL = 2,
  ST_INERTIA_FAIL = 3,
  ST_NUMERICAL = 4,
  ST_RUNNING = -1,
};

// ---------------------------------------------------------------------------------------
// workspace fields
// ---------------------------------------------------------------------------------------
enum : int {
  // iterate (two ping-pong copies)
  F_Z = 0,          // 6: y, vy, x, vx, angle, angledot
  F_U = 6,          // 1
  F_LAM = 7,        // 6: defect multipliers
  F_ZLA = 13, F_ZUA = 14, F_ZLU = 15, F_ZUU = 16,
  N_ITER = 17,
  // step
  F_DS = 0,         // 6
  F_DU = 6,
  F_PI = 7,         // 6: new defect multipliers
  N_STEP = 13,
  // factor
  F_K = 0,          // 7 feedback gains
  F_KFF = 7,
  F_P = 8,          // 28 packed symmetric cost-to-go Hessian P_{k-1}
  F_PV = 36,        // 7 cost-to-go gradient p_{k-1}
  N_FACT = 43,
  N_FIELDS = 2 * N_ITER + N_STEP + N_FACT   // 90 doubles per stage per problem
};

struct Mesh {
  int N;                 // number of steps (nt-1)
  const double* h;       // h[k]  = time[k]-time[k-1], k=1..N (index 0 unused)   LO:21
  const double* tau;     // tau[k] = time[k]
};

struct Ws {
  double* base;          // [N_FIELDS][N+1][B]
  long B;                // slots
  int N1;                // N+1
  long j;                // this thread's slot
  LM_HD double& it(int buf, int f, int k) const { return base[((long)(buf * N_ITER + f) * N1 + k) * B + j]; }
  LM_HD double& st(int f, int k) const { return base[((long)(2 * N_ITER + f) * N1 + k) * B + j]; }
  LM_HD double& fa(int f, int k) const { return base[((long)(2 * N_ITER + N_STEP + f) * N1 + k) * B + j]; }
};

constexpr int NFILT = 12;

// Per-problem scalar state (registers / local memory).
struct Scal {
  double tf, zLt, zUt;            // global final time and its bound multipliers
  double sg1, sg2, zs1, zs2, nu3; // terminal slacks, their multipliers, equality multiplier
  double theta, fobj, sumlog;     // constraint violation (l1), objective, sum of log slacks
  double dual_inf, prim_inf, cmin, cmax, sum_lam, sum_z;
};

struct Ctl {
  double mu, tau;
  double theta_max, theta_min;
  double dw_last;
  double ft[NFILT], fp[NFILT];
  int nf;
  int iter;
  int status;
};

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
LM_HD int pidx(int i, int j) {  // packed lower-triangular index, i >= j
  return i * (i + 1) / 2 + j;
}
LM_HD double dmax(double a, double b) { return a > b ? a : b; }
LM_HD double dmin(double a, double b) { return a < b ? a : b; }

// Stage Jacobian data: everything needed to apply E^{-1} and E^{-T}.
//   E = d defect_k / d s_k  for s = (y, vy, x, vx, a, w, tf); d defect_k / d s_{k-1} = -I;
//   d defect_k / d u_k = -beta e_5.
struct StageJac {
  double al;                  // alpha = h*T*tf
  double a2a, a2b, a2c, a2d;  // alpha^2 * (ay_y, ay_x, ax_y, ax_x)
  double ala, alb, alc, ald;  // alpha   * (ay_y, ay_x, ax_y, ax_x)
  double m11, m13, m31, m33;  // inverse of the 2x2 velocity block
  double ga1, ga3;            // alpha * (ay_a, ax_a)
  double e0, e1, e2, e3, e4, e5;  // tf column (negated entries of E)
  double beta;                // alpha * asc
};

LM_HD void stagejac_build(const Params& P, double kap, double tf, double taum /* d mass/d tf */,
                          const Accel1& f, double vy, double vx, double w, double u, StageJac& J) {
  const double al = kap * tf;
  J.al = al;
  J.ala = al * f.ay_y; J.alb = al * f.ay_x; J.alc = al * f.ax_y; J.ald = al * f.ax_x;
  J.a2a = al * J.ala; J.a2b = al * J.alb; J.a2c = al * J.alc; J.a2d = al * J.ald;
  const double d11 = 1.0 - J.a2a, d33 = 1.0 - J.a2d;
  const double Dinv = 1.0 / (d11 * d33 - J.a2b * J.a2c);
  J.m11 = d33 * Dinv; J.m13 = J.a2b * Dinv; J.m31 = J.a2c * Dinv; J.m33 = d11 * Dinv;
  J.ga1 = al * f.ay_a; J.ga3 = al * f.ax_a;
  J.e0 = kap * vy;
  J.e1 = kap * f.ay + al * f.ay_m * taum;
  J.e2 = kap * vx;
  J.e3 = kap * f.ax + al * f.ax_m * taum;
  J.e4 = kap * w;
  J.e5 = kap * P.asc * u;
  J.beta = al * P.asc;
}

// v <- E^{-1} v
LM_HD void solveE(const StageJac& J, double* v) {
  const double v6 = v[6];
  const double v5 = fma(J.e5, v6, v[5]);
  const double v4 = v[4] + J.al * v5 + J.e4 * v6;
  const double r0 = fma(J.e0, v6, v[0]);
  const double r2 = fma(J.e2, v6, v[2]);
  const double r1 = v[1] + J.ga1 * v4 + J.e1 * v6;
  const double r3 = v[3] + J.ga3 * v4 + J.e3 * v6;
  const double t1 = r1 + J.ala * r0 + J.alb * r2;
  const double t3 = r3 + J.alc * r0 + J.ald * r2;
  const double v1 = J.m11 * t1 + J.m13 * t3;
  const double v3 = J.m31 * t1 + J.m33 * t3;
  v[0] = fma(J.al, v1, r0);
  v[1] = v1;
  v[2] = fma(J.al, v3, r2);
  v[3] = v3;
  v[4] = v4;
  v[5] = v5;
}

// g <- E^{-T} g
LM_HD void solveET(const StageJac& J, double* g) {
  // velocity/position block:  M^T w = g_p
  const double t0 = g[0] + J.ala * g[1] + J.alc * g[3];
  const double t2 = g[2] + J.alb * g[1] + J.ald * g[3];
  // [1-a2a, -a2c; -a2b, 1-a2d] [w0; w2] = [t0; t2]  (transpose of the 2x2 in solveE)
  const double w0 = J.m11 * t0 + J.m31 * t2;
  const double w2 = J.m13 * t0 + J.m33 * t2;
  const double w1 = fma(J.al, w0, g[1]);
  const double w3 = fma(J.al, w2, g[3]);
  const double gaw = J.ga1 * w1 + J.ga3 * w3;
  const double gew = J.e0 * w0 + J.e1 * w1 + J.e2 * w2 + J.e3 * w3;
  const double g4 = g[4], g5 = g[5];
  const double w4 = g4 + gaw;
  const double w5 = g5 + J.al * w4;
  const double w6 = g[6] + gew + J.e4 * w4 + J.e5 * w5;
  g[0] = w0; g[1] = w1; g[2] = w2; g[3] = w3; g[4] = w4; g[5] = w5; g[6] = w6;
}

// E^T lam restricted to the six real states (lam[6] multiplies the trivial tf row).
LM_HD void applyET6(const StageJac& J, const double* l, double* out) {
  out[0] = l[0] - J.ala * l[1] - J.alc * l[3];
  out[1] = l[1] - J.al * l[0];
  out[2] = l[2] - J.alb * l[1] - J.ald * l[3];
  out[3] = l[3] - J.al * l[2];
  out[4] = l[4] - J.ga1 * l[1] - J.ga3 * l[3];
  out[5] = l[5] - J.al * l[4];
}

// ---------------------------------------------------------------------------------------
// initial guess (restated independently in oracle/ascent_nlp.py: initial_guess)
// ---------------------------------------------------------------------------------------
// Start point: a dynamically consistent roll-out of a bang-bang pitch-acceleration profile
// (u = +0.9 until t1, -0.9 until t1+t2, then 0), i.e. the shape of the known optimum
// (Angle_vs_Time.png).  Only the three terminal rows and the interior push of `angle`
// are infeasible at the start.
struct GuessProfile { double t1, t2, ulev; };

LM_HD GuessProfile guess_profile(const Params& P) {
  GuessProfile g;
  const double a_tgt = dmin(0.40, 0.8 * P.a_ub);   // ~69 deg of physical pitch
  const double w_rem = 3.5e-4;                      // residual pitch rate [rad/s of `angle`]
  g.ulev = 0.9 * P.u_ub;
  const double ueff = g.ulev * P.asc;
  g.t1 = sqrt(a_tgt / ueff);
  g.t2 = dmax(g.t1 - w_rem / ueff, 0.0);
  return g;
}

LM_HD void init_guess(const Params& P, const Mesh& M, const Options& O, const Ws& W, Scal& s) {
  const int N = M.N;
  const double tf0 = dmin(dmax(O.tf_guess, 1e-2 * P.tf_ub), 0.99 * P.tf_ub);
  const GuessProfile g = guess_profile(P);
  double y = 0, vy = 0, x = 0, vx = 0, a = 0, w = 0, t_prev = 0;
  const double a_lo = 1e-2 * P.a_ub, a_hi = 0.99 * P.a_ub;
  for (int k = 1; k <= N; ++k) {
    const double t = M.tau[k] * tf0 * P.T;
    const double dt = t - t_prev;
    const double tm = 0.5 * (t + t_prev);
    const double u = tm < g.t1 ? g.ulev : (tm < g.t1 + g.t2 ? -g.ulev : 0.0);
    w += dt * P.asc * u;
    a += dt * w;
    const double ac = dmin(dmax(a, a_lo), a_hi);
    const double m = P.mflow * P.T * M.tau[k] * tf0;
    // backward-Euler step for (y, vy, x, vx): two fixed-point sweeps
    double yn = y + dt * vy, xn = x + dt * vx, vyn = vy, vxn = vx;
    for (int itr = 0; itr < 3; ++itr) {
      Accel1 f;
      accel_first(P, yn, xn, ac, m, f);
      vyn = vy + dt * f.ay; vxn = vx + dt * f.ax;
      yn = y + dt * vyn;    xn = x + dt * vxn;
    }
    y = yn; vy = vyn; x = xn; vx = vxn;
    W.it(0, F_Z + 0, k) = y;  W.it(0, F_Z + 1, k) = vy;
    W.it(0, F_Z + 2, k) = x;  W.it(0, F_Z + 3, k) = vx;
    W.it(0, F_Z + 4, k) = ac; W.it(0, F_Z + 5, k) = w;
    W.it(0, F_U, k) = u;
    for (int i = 0; i < 6; ++i) W.it(0, F_LAM + i, k) = 0.0;
    W.it(0, F_ZLA, k) = 1.0; W.it(0, F_ZUA, k) = 1.0;
    W.it(0, F_ZLU, k) = 1.0; W.it(0, F_ZUU, k) = 1.0;
    for (int i = 0; i < N_STEP; ++i) W.st(i, k) = 0.0;
    t_prev = t;
  }
  s.tf = tf0;
  s.zLt = 1.0; s.zUt = 1.0;
  s.sg1 = 1e-2; s.sg2 = 1e-2; s.zs1 = 1.0; s.zs2 = 1.0; s.nu3 = 0.0;
}

// ---------------------------------------------------------------------------------------
// evaluation pass: trial point  x + alpha*dx  (alpha=0: the current point).
// Writes the trial iterate into buffer `dst`, returns its merit / error ingredients.
// ---------------------------------------------------------------------------------------
struct TermStep { double dtf, dsg1, dsg2, dzs1, dzs2, dnu3, dzLt, dzUt; };

LM_HD void eval_pass(const Params& P, const Mesh& M, const Options& O, const Ws& W, int src, int dst,
                     const Scal& c0, const TermStep& ts, double mu, double alpha, double alpha_z,
                     double alpha_lam, Scal& t) {
  const int N = M.N;
  const double kS = 1e10;  // kappa_Sigma, IPOPT eq. (16)
  t.tf = c0.tf + alpha * ts.dtf;
  const double tf = t.tf;
  double zp[6] = {0, 0, 0, 0, 0, 0};      // previous node's trial state
  double pend[6] = {0, 0, 0, 0, 0, 0};    // E_{k-1}^T lam_{k-1} + bound terms, awaiting -lam_k
  double theta = 0, prim = 0, dual = 0, sumlog = 0, cmin = 1e300, cmax = 0, slam = 0, sz = 0;
  double gtf = 0;                          // d Lagrangian / d tf accumulated over stages
  bool bad = false;
  for (int k = 1; k <= N; ++k) {
    double z[6], lam[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) z[i] = fma(alpha, W.st(F_DS + i, k), W.it(src, F_Z + i, k));
    const double u_old = W.it(src, F_U, k);
    const double du = W.st(F_DU, k);
    const double u = fma(alpha, du, u_old);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double l0 = W.it(src, F_LAM + i, k);
      lam[i] = fma(alpha_lam, W.st(F_PI + i, k) - l0, l0);
    }
    // bound multipliers: dz = mu/d - z - (z/d) dx  (old d, old z), then kappa_Sigma clip
    const double a_old = W.it(src, F_Z + 4, k), da = W.st(F_DS + 4, k);
    double zla = W.it(src, F_ZLA, k), zua = W.it(src, F_ZUA, k);
    double zlu = W.it(src, F_ZLU, k), zuu = W.it(src, F_ZUU, k);
    {
      const double dLa = a_old, dUa = P.a_ub - a_old, dLu = u_old + P.u_ub, dUu = P.u_ub - u_old;
      zla += alpha_z * (mu / dLa - zla - zla / dLa * da);
      zua += alpha_z * (mu / dUa - zua + zua / dUa * da);
      zlu += alpha_z * (mu / dLu - zlu - zlu / dLu * du);
      zuu += alpha_z * (mu / dUu - zuu + zuu / dUu * du);
    }
    const double dLa = z[4], dUa = P.a_ub - z[4], dLu = u + P.u_ub, dUu = P.u_ub - u;
    if (!(dLa > 0 && dUa > 0 && dLu > 0 && dUu > 0)) bad = true;
    zla = dmax(dmin(zla, kS * mu / dLa), mu / (kS * dLa));
    zua = dmax(dmin(zua, kS * mu / dUa), mu / (kS * dUa));
    zlu = dmax(dmin(zlu, kS * mu / dLu), mu / (kS * dLu));
    zuu = dmax(dmin(zuu, kS * mu / dUu), mu / (kS * dUu));
    sumlog += log(dLa) + log(dUa) + log(dLu) + log(dUu);
    {
      const double c1 = dLa * zla, c2 = dUa * zua, c3 = dLu * zlu, c4 = dUu * zuu;
      cmin = dmin(cmin, dmin(dmin(c1, c2), dmin(c3, c4)));
      cmax = dmax(cmax, dmax(dmax(c1, c2), dmax(c3, c4)));
    }
    sz += zla + zua + zlu + zuu;
    // dynamics
    const double kap = M.h[k] * P.T;
    const double taum = P.mflow * P.T * M.tau[k];
    Accel1 f;
    accel_first(P, z[0], z[2], z[4], taum * tf, f);
    StageJac J;
    stagejac_build(P, kap, tf, taum, f, z[1], z[3], z[5], u, J);
    const double al = J.al;
    double c[6];
    c[0] = z[0] - zp[0] - al * z[1];
    c[1] = z[1] - zp[1] - al * f.ay;
    c[2] = z[2] - zp[2] - al * z[3];
    c[3] = z[3] - zp[3] - al * f.ax;
    c[4] = z[4] - zp[4] - al * z[5];
    c[5] = z[5] - zp[5] - J.beta * u;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double ac = fabs(c[i]);
      theta += ac;
      prim = dmax(prim, ac);
      slam += fabs(lam[i]);
    }
    // dual residual of the previous stage is complete once lam_k is known
    if (k > 1) {
#pragma unroll
      for (int i = 0; i < 6; ++i) dual = dmax(dual, fabs(pend[i] - lam[i]));
    }
    applyET6(J, lam, pend);
    pend[4] += zua - zla;
    dual = dmax(dual, fabs(-J.beta * lam[5] - zlu + zuu));     // d L / d u_k
    gtf -= J.e0 * lam[0] + J.e1 * lam[1] + J.e2 * lam[2] + J.e3 * lam[3] + J.e4 * lam[4] + J.e5 * lam[5];
    // write trial iterate
#pragma unroll
    for (int i = 0; i < 6; ++i) { W.it(dst, F_Z + i, k) = z[i]; W.it(dst, F_LAM + i, k) = lam[i]; zp[i] = z[i]; }
    W.it(dst, F_U, k) = u;
    W.it(dst, F_ZLA, k) = zla; W.it(dst, F_ZUA, k) = zua;
    W.it(dst, F_ZLU, k) = zlu; W.it(dst, F_ZUU, k) = zuu;
  }
  // terminal node
  t.sg1 = c0.sg1 + alpha * ts.dsg1;
  t.sg2 = c0.sg2 + alpha * ts.dsg2;
  t.nu3 = c0.nu3 + alpha_lam * ts.dnu3;
  t.zs1 = c0.zs1 + alpha_z * ts.dzs1;
  t.zs2 = c0.zs2 + alpha_z * ts.dzs2;
  t.zLt = c0.zLt + alpha_z * ts.dzLt;
  t.zUt = c0.zUt + alpha_z * ts.dzUt;
  const double dLt = tf, dUt = P.tf_ub - tf;
  if (!(t.sg1 > 0 && t.sg2 > 0 && dLt > 0 && dUt > 0)) bad = true;
  t.zs1 = dmax(dmin(t.zs1, kS * mu / t.sg1), mu / (kS * t.sg1));
  t.zs2 = dmax(dmin(t.zs2, kS * mu / t.sg2), mu / (kS * t.sg2));
  t.zLt = dmax(dmin(t.zLt, kS * mu / dLt), mu / (kS * dLt));
  t.zUt = dmax(dmin(t.zUt, kS * mu / dUt), mu / (kS * dUt));
  Terminal T;
  terminal_eval(P, zp[0], zp[1], zp[2], zp[3], T);
  const double c1 = T.g1 - t.sg1, c2 = T.g2 - t.sg2, c3 = T.g3;
  theta += fabs(c1) + fabs(c2) + fabs(c3);
  prim = dmax(prim, dmax(fabs(c1), dmax(fabs(c2), fabs(c3))));
  sumlog += log(t.sg1) + log(t.sg2) + log(dLt) + log(dUt);
  {
    const double q1 = t.sg1 * t.zs1, q2 = t.sg2 * t.zs2, q3 = dLt * t.zLt, q4 = dUt * t.zUt;
    cmin = dmin(cmin, dmin(dmin(q1, q2), dmin(q3, q4)));
    cmax = dmax(cmax, dmax(dmax(q1, q2), dmax(q3, q4)));
  }
  sz += t.zs1 + t.zs2 + t.zLt + t.zUt;
  slam += t.zs1 + t.zs2 + fabs(t.nu3);
  // Lagrangian gradient at the last node: multipliers of g1,g2 are -zs1,-zs2 (slack stationarity)
  {
    const double rinv = 1.0 / T.rT;
    pend[0] += -t.zs1 * T.Yb * rinv + t.nu3 * zp[1];
    pend[2] += -t.zs1 * zp[2] * rinv + t.nu3 * zp[3];
    pend[1] += -t.zs2 * 2.0 * zp[1] + t.nu3 * T.Yb;
    pend[3] += -t.zs2 * 2.0 * zp[3] + t.nu3 * zp[2];
#pragma unroll
    for (int i = 0; i < 6; ++i) dual = dmax(dual, fabs(pend[i]));
  }
  gtf += O.obj_scale - t.zLt + t.zUt;
  dual = dmax(dual, fabs(gtf));
  t.theta = theta;
  t.fobj = O.obj_scale * tf;
  t.sumlog = bad ? -1e300 : sumlog;
  t.prim_inf = prim; t.dual_inf = dual; t.cmin = cmin; t.cmax = cmax; t.sum_lam = slam; t.sum_z = sz;
  if (bad || !(theta == theta)) t.theta = 1e300;
}

// ---------------------------------------------------------------------------------------
// backward Riccati sweep = block LDL^T of the KKT matrix in stage order.
// Returns false if a pivot shows wrong inertia (caller increases delta_w and retries).
// ---------------------------------------------------------------------------------------
LM_HD bool riccati_backward(const Params& P, const Mesh& M, const Options& O, const Ws& W, int src,
                            const Scal& c0, double mu, double dw, bool ls, double* dtf_out, double* p0_out) {
  // ls == true: least-squares multiplier estimate (IPOPT section 3.6): Hessian := I, defects := 0,
  // gradient := grad f - zL + zU; the forward sweep then returns the multipliers in F_PI.
  const int N = M.N;
  const double tf = c0.tf;
  double Pm[28];   // packed lower triangle of the cost-to-go Hessian
  double pv[7];
#pragma unroll
  for (int i = 0; i < 28; ++i) Pm[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 7; ++i) pv[i] = 0.0;
  double zn[6];    // state at node k
#pragma unroll
  for (int i = 0; i < 6; ++i) zn[i] = W.it(src, F_Z + i, N);
  // ---- terminal contributions (added to stage N's Q and q) ----
  {
    Terminal T;
    terminal_eval(P, zn[0], zn[1], zn[2], zn[3], T);
    const double rinv = 1.0 / T.rT;
    const double g1y = T.Yb * rinv, g1x = zn[2] * rinv;
    const double w1 = ls ? 1.0 : c0.zs1 / c0.sg1, w2 = ls ? 1.0 : c0.zs2 / c0.sg2, w3 = 1.0 / O.delta_c;
    const double c1 = T.g1 - c0.sg1, c2 = T.g2 - c0.sg2;
    // gradient terms: grad g_i * (w_i c_i - mu/sg_i), grad g3 * (nu3 + g3/delta_c)
    const double q1 = ls ? -c0.zs1 : w1 * c1 - mu / c0.sg1;
    const double q2 = ls ? -c0.zs2 : w2 * c2 - mu / c0.sg2;
    const double q3 = ls ? 0.0 : c0.nu3 + T.g3 * w3;
    const double hz1 = ls ? 0.0 : c0.zs1, hz2 = ls ? 0.0 : c0.zs2, hn3 = ls ? 0.0 : c0.nu3;
    const double G2y = 2.0 * zn[1], G2x = 2.0 * zn[3];
    // g3 gradient wrt (y, vy, x, vx) = (vy, Yb, vx, x)
    const double G3[4] = {zn[1], T.Yb, zn[3], zn[2]};
    pv[0] = g1y * q1 + G3[0] * q3;
    pv[1] = G2y * q2 + G3[1] * q3;
    pv[2] = g1x * q1 + G3[2] * q3;
    pv[3] = G2x * q2 + G3[3] * q3;
    // Hessian: w_i grad grad^T + multipliers * second derivatives (mult of g1,g2 = -zs)
    const double r3 = rinv * rinv * rinv;
    const double h1yy = zn[2] * zn[2] * r3, h1xx = T.Yb * T.Yb * r3, h1yx = -T.Yb * zn[2] * r3;
    Pm[pidx(0, 0)] = w1 * g1y * g1y - hz1 * h1yy + w3 * G3[0] * G3[0];
    Pm[pidx(2, 0)] = w1 * g1x * g1y - hz1 * h1yx + w3 * G3[2] * G3[0];
    Pm[pidx(2, 2)] = w1 * g1x * g1x - hz1 * h1xx + w3 * G3[2] * G3[2];
    Pm[pidx(1, 1)] = w2 * G2y * G2y - 2.0 * hz2 + w3 * G3[1] * G3[1];
    Pm[pidx(3, 1)] = w2 * G2x * G2y + w3 * G3[3] * G3[1];
    Pm[pidx(3, 3)] = w2 * G2x * G2x - 2.0 * hz2 + w3 * G3[3] * G3[3];
    Pm[pidx(1, 0)] = w3 * G3[1] * G3[0] + hn3;
    Pm[pidx(3, 0)] = w3 * G3[3] * G3[0];
    Pm[pidx(2, 1)] = w3 * G3[2] * G3[1];
    Pm[pidx(3, 2)] = w3 * G3[3] * G3[2] + hn3;
    // tf: objective, bound barrier, regularisation
    const double dLt = tf, dUt = P.tf_ub - tf;
    pv[6] = ls ? O.obj_scale - c0.zLt + c0.zUt : O.obj_scale - mu / dLt + mu / dUt;
    Pm[pidx(6, 6)] = ls ? 1.0 : c0.zLt / dLt + c0.zUt / dUt + dw;
  }
  bool ok = true;
  for (int k = N; k >= 1; --k) {
    double lam[6], zm[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) lam[i] = W.it(src, F_LAM + i, k);
    if (k > 1) {
#pragma unroll
      for (int i = 0; i < 6; ++i) zm[i] = W.it(src, F_Z + i, k - 1);
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) zm[i] = 0.0;
    }
    const double u = W.it(src, F_U, k);
    const double zla = W.it(src, F_ZLA, k), zua = W.it(src, F_ZUA, k);
    const double zlu = W.it(src, F_ZLU, k), zuu = W.it(src, F_ZUU, k);
    const double kap = M.h[k] * P.T;
    const double taum = P.mflow * P.T * M.tau[k];
    Accel1 f;
    accel_first(P, zn[0], zn[2], zn[4], taum * tf, f);
    StageJac J;
    stagejac_build(P, kap, tf, taum, f, zn[1], zn[3], zn[5], u, J);
    const double al = J.al;
    // ---- W = Q_k + P_k (in place), g = q_k + p_k ----
    const double dLa = zn[4], dUa = P.a_ub - zn[4];
    const double dLu = u + P.u_ub, dUu = P.u_ub - u;
    double R, r, sig;
    if (!ls) {
      Accel2 h2;
      accel_second(P, f, lam[1], lam[3], h2);
      Pm[pidx(0, 0)] += -al * h2.yy + dw;
      Pm[pidx(2, 0)] += -al * h2.yx;
      Pm[pidx(2, 2)] += -al * h2.xx + dw;
      Pm[pidx(4, 0)] += -al * h2.ya;
      Pm[pidx(4, 2)] += -al * h2.xa;
      Pm[pidx(4, 4)] += -al * h2.aa + zla / dLa + zua / dUa + dw;
      Pm[pidx(1, 1)] += dw; Pm[pidx(3, 3)] += dw; Pm[pidx(5, 5)] += dw;
      const double Phy = lam[1] * f.ay_y + lam[3] * f.ax_y;
      const double Phx = lam[1] * f.ay_x + lam[3] * f.ax_x;
      const double Pha = lam[1] * f.ay_a + lam[3] * f.ax_a;
      const double Phm = lam[1] * f.ay_m + lam[3] * f.ax_m;
      Pm[pidx(6, 0)] += -kap * Phy - al * h2.ym * taum;
      Pm[pidx(6, 2)] += -kap * Phx - al * h2.xm * taum;
      Pm[pidx(6, 4)] += -kap * Pha - al * h2.am * taum;
      Pm[pidx(6, 1)] += -kap * lam[0];
      Pm[pidx(6, 3)] += -kap * lam[2];
      Pm[pidx(6, 5)] += -kap * lam[4];
      Pm[pidx(6, 6)] += -2.0 * kap * Phm * taum - al * h2.mm * taum * taum;
      pv[4] += -mu / dLa + mu / dUa;
      R = zlu / dLu + zuu / dUu + dw;
      r = -mu / dLu + mu / dUu;
      sig = -kap * lam[5] * P.asc;   // u-tf cross term
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) Pm[pidx(i, i)] += 1.0;
      pv[4] += -zla + zua;
      R = 1.0; r = -zlu + zuu; sig = 0.0;
    }
    // ---- defect ----
    double c[7];
    c[0] = zn[0] - zm[0] - al * zn[1];
    c[1] = zn[1] - zm[1] - al * f.ay;
    c[2] = zn[2] - zm[2] - al * zn[3];
    c[3] = zn[3] - zm[3] - al * f.ax;
    c[4] = zn[4] - zm[4] - al * zn[5];
    c[5] = zn[5] - zm[5] - J.beta * u;
    c[6] = 0.0;
    if (ls) {
#pragma unroll
      for (int i = 0; i < 6; ++i) c[i] = 0.0;
    }
    // ---- congruence  Wt = E^{-T} W E^{-1} ----
    double Wf[7][7];
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) { Wf[i][j] = Pm[pidx(i, j)]; Wf[j][i] = Wf[i][j]; }
#pragma unroll
    for (int j = 0; j < 7; ++j) {        // columns: E^{-T} W
      double col[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) col[i] = Wf[i][j];
      solveET(J, col);
#pragma unroll
      for (int i = 0; i < 7; ++i) Wf[i][j] = col[i];
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) {        // rows: (.) E^{-1}  ==  E^{-T} applied to the row as a vector
      double row[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) row[j] = Wf[i][j];
      solveET(J, row);
#pragma unroll
      for (int j = 0; j < 7; ++j) Wf[i][j] = row[j];
    }
    solveET(J, pv);                      // g~ = E^{-T} (q + p)
    // ---- condense the control ----
    const double beta = J.beta;
    const double Ruu = R + beta * beta * Wf[5][5];
    double Rux[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) Rux[i] = beta * 0.5 * (Wf[5][i] + Wf[i][5]);
    Rux[6] += sig;
    double rx[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      double acc = pv[i];
#pragma unroll
      for (int j = 0; j < 6; ++j) acc -= 0.5 * (Wf[i][j] + Wf[j][i]) * c[j];
      rx[i] = acc;
    }
    const double ru = r + beta * rx[5];
    if (!(Ruu > 0.0) || !(Ruu < 1e300)) ok = false;
    const double Rinv = 1.0 / Ruu;
#pragma unroll
    for (int i = 0; i < 7; ++i) W.fa(F_K + i, k) = -Rux[i] * Rinv;
    W.fa(F_KFF, k) = -ru * Rinv;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const double v = 0.5 * (Wf[i][j] + Wf[j][i]) - Rux[i] * Rux[j] * Rinv;
        Pm[pidx(i, j)] = v;
        W.fa(F_P + pidx(i, j), k) = v;
      }
      pv[i] = rx[i] - Rux[i] * ru * Rinv;
      W.fa(F_PV + i, k) = pv[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) zn[i] = zm[i];
    if (!ok) return false;
  }
  // free initial tf: minimise 0.5 P66 dtf^2 + p6 dtf
  const double P66 = Pm[pidx(6, 6)];
  if (!(P66 > 0.0)) return false;
  *dtf_out = -pv[6] / P66;
  *p0_out = P66;
  return true;
}

// ---------------------------------------------------------------------------------------
// forward sweep: recover the Newton step, fraction-to-boundary limits and d(phi)/d(alpha)
// ---------------------------------------------------------------------------------------
struct StepInfo { double a_max, a_z, dphi, dxmax, pimax; };

LM_HD void riccati_forward(const Params& P, const Mesh& M, const Options& O, const Ws& W, int src,
                           const Scal& c0, double mu, double tau, double dtf, bool ls, TermStep& ts, StepInfo& si) {
  const int N = M.N;
  const double tf = c0.tf;
  double ds[7] = {0, 0, 0, 0, 0, 0, dtf};
  double zm[6] = {0, 0, 0, 0, 0, 0};
  double amax = 1.0, az = 1.0, dphi = 0.0, dxmax = fabs(dtf), pimax = 0.0;
  const double cw = ls ? 0.0 : 1.0;   // defects are dropped in the least-squares mode
  for (int k = 1; k <= N; ++k) {
    double zn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) zn[i] = W.it(src, F_Z + i, k);
    const double u = W.it(src, F_U, k);
    const double kap = M.h[k] * P.T;
    const double taum = P.mflow * P.T * M.tau[k];
    Accel1 f;
    accel_first(P, zn[0], zn[2], zn[4], taum * tf, f);
    StageJac J;
    stagejac_build(P, kap, tf, taum, f, zn[1], zn[3], zn[5], u, J);
    const double al = J.al;
    // new multipliers  pi_k = -(P_{k-1} ds_{k-1} + p_{k-1})
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double acc = W.fa(F_PV + i, k);
#pragma unroll
      for (int j = 0; j < 7; ++j) acc = fma(W.fa(F_P + (i >= j ? pidx(i, j) : pidx(j, i)), k), ds[j], acc);
      W.st(F_PI + i, k) = -acc;
      pimax = dmax(pimax, fabs(acc));
    }
    double du = W.fa(F_KFF, k);
#pragma unroll
    for (int i = 0; i < 7; ++i) du = fma(W.fa(F_K + i, k), ds[i], du);
    double xi[7];
    xi[0] = ds[0] - cw * (zn[0] - zm[0] - al * zn[1]);
    xi[1] = ds[1] - cw * (zn[1] - zm[1] - al * f.ay);
    xi[2] = ds[2] - cw * (zn[2] - zm[2] - al * zn[3]);
    xi[3] = ds[3] - cw * (zn[3] - zm[3] - al * f.ax);
    xi[4] = ds[4] - cw * (zn[4] - zm[4] - al * zn[5]);
    xi[5] = ds[5] - cw * (zn[5] - zm[5] - J.beta * u) + J.beta * du;
    xi[6] = dtf;
    solveE(J, xi);
#pragma unroll
    for (int i = 0; i < 6; ++i) { ds[i] = xi[i]; W.st(F_DS + i, k) = xi[i]; zm[i] = zn[i]; dxmax = dmax(dxmax, fabs(xi[i])); }
    W.st(F_DU, k) = du;
    dxmax = dmax(dxmax, fabs(du));
    // fraction to the boundary (IPOPT eq. 15) for angle and control, and their multipliers
    const double da = ds[4];
    const double dLa = zn[4], dUa = P.a_ub - zn[4], dLu = u + P.u_ub, dUu = P.u_ub - u;
    if (da < 0) amax = dmin(amax, -tau * dLa / da);
    if (da > 0) amax = dmin(amax, tau * dUa / da);
    if (du < 0) amax = dmin(amax, -tau * dLu / du);
    if (du > 0) amax = dmin(amax, tau * dUu / du);
    const double zla = W.it(src, F_ZLA, k), zua = W.it(src, F_ZUA, k);
    const double zlu = W.it(src, F_ZLU, k), zuu = W.it(src, F_ZUU, k);
    const double d1 = mu / dLa - zla - zla / dLa * da;
    const double d2 = mu / dUa - zua + zua / dUa * da;
    const double d3 = mu / dLu - zlu - zlu / dLu * du;
    const double d4 = mu / dUu - zuu + zuu / dUu * du;
    if (d1 < 0) az = dmin(az, -tau * zla / d1);
    if (d2 < 0) az = dmin(az, -tau * zua / d2);
    if (d3 < 0) az = dmin(az, -tau * zlu / d3);
    if (d4 < 0) az = dmin(az, -tau * zuu / d4);
    dphi += (-mu / dLa + mu / dUa) * da + (-mu / dLu + mu / dUu) * du;
  }
  // terminal slacks and multipliers
  {
    Terminal T;
    terminal_eval(P, zm[0], zm[1], zm[2], zm[3], T);
    const double rinv = 1.0 / T.rT;
    const double dg1 = T.Yb * rinv * ds[0] + zm[2] * rinv * ds[2];
    const double dg2 = 2.0 * zm[1] * ds[1] + 2.0 * zm[3] * ds[3];
    const double dg3 = zm[1] * ds[0] + T.Yb * ds[1] + zm[3] * ds[2] + zm[2] * ds[3];
    ts.dtf = dtf;
    ts.dsg1 = dg1 + (T.g1 - c0.sg1);
    ts.dsg2 = dg2 + (T.g2 - c0.sg2);
    ts.dnu3 = (dg3 + cw * T.g3) / O.delta_c;
    ts.dzs1 = mu / c0.sg1 - c0.zs1 - c0.zs1 / c0.sg1 * ts.dsg1;
    ts.dzs2 = mu / c0.sg2 - c0.zs2 - c0.zs2 / c0.sg2 * ts.dsg2;
    const double dLt = tf, dUt = P.tf_ub - tf;
    ts.dzLt = mu / dLt - c0.zLt - c0.zLt / dLt * dtf;
    ts.dzUt = mu / dUt - c0.zUt + c0.zUt / dUt * dtf;
    if (ts.dsg1 < 0) amax = dmin(amax, -tau * c0.sg1 / ts.dsg1);
    if (ts.dsg2 < 0) amax = dmin(amax, -tau * c0.sg2 / ts.dsg2);
    if (dtf < 0) amax = dmin(amax, -tau * dLt / dtf);
    if (dtf > 0) amax = dmin(amax, tau * dUt / dtf);
    if (ts.dzs1 < 0) az = dmin(az, -tau * c0.zs1 / ts.dzs1);
    if (ts.dzs2 < 0) az = dmin(az, -tau * c0.zs2 / ts.dzs2);
    if (ts.dzLt < 0) az = dmin(az, -tau * c0.zLt / ts.dzLt);
    if (ts.dzUt < 0) az = dmin(az, -tau * c0.zUt / ts.dzUt);
    dphi += (O.obj_scale - mu / dLt + mu / dUt) * dtf - mu / c0.sg1 * ts.dsg1 - mu / c0.sg2 * ts.dsg2;
    dxmax = dmax(dxmax, dmax(fabs(ts.dsg1), fabs(ts.dsg2)));
  }
  si.a_max = amax; si.a_z = az; si.dphi = dphi; si.dxmax = dxmax; si.pimax = pimax;
}

// ---------------------------------------------------------------------------------------
// IPM driver pieces
// ---------------------------------------------------------------------------------------
LM_HD double kkt_error(const Scal& s, double mu, int n_eq, int n_bd) {
  const double smax = 100.0;
  const double sd = dmax(smax, (s.sum_lam + s.sum_z) / (double)(n_eq + n_bd)) / smax;
  const double sc = dmax(smax, s.sum_z / (double)n_bd) / smax;
  const double comp = dmax(fabs(s.cmax - mu), fabs(s.cmin - mu));
  return dmax(dmax(s.dual_inf / sd, s.prim_inf), comp / sc);
}

LM_HD bool filter_ok(const Ctl& c, double th, double ph) {
  for (int i = 0; i < c.nf; ++i)
    if (th >= c.ft[i] && ph >= c.fp[i]) return false;
  return true;
}

LM_HD void filter_add(Ctl& c, double th, double ph) {
  // drop dominated entries, then append (overwrite the weakest if full)
  int n = 0;
  for (int i = 0; i < c.nf; ++i)
    if (!(c.ft[i] >= th && c.fp[i] >= ph)) { c.ft[n] = c.ft[i]; c.fp[n] = c.fp[i]; ++n; }
  if (n == NFILT) {   // merge: keep the envelope conservative by replacing the largest-theta entry
    int w = 0;
    for (int i = 1; i < n; ++i) if (c.ft[i] > c.ft[w]) w = i;
    c.ft[w] = th; c.fp[w] = ph;
  } else { c.ft[n] = th; c.fp[n] = ph; ++n; }
  c.nf = n;
}

// One complete solve of one problem.  `W.j` selects the slot.  Output goes to the caller.
struct SolveOut { double tf; int status; int iters; double kkt; double mu; int cur; };

LM_HD void ipm_solve(const Params& P, const Mesh& M, const Options& O, const Ws& W, bool have_guess,
                     SolveOut& out) {
  Scal cur, trial;
  Ctl ctl;
  TermStep ts;
  ts.dtf = ts.dsg1 = ts.dsg2 = ts.dzs1 = ts.dzs2 = ts.dnu3 = ts.dzLt = ts.dzUt = 0.0;
  if (!have_guess) init_guess(P, M, O, W, cur);
  const int N = M.N;
  const int n_eq = 6 * N + 3;
  const int n_bd = 4 * N + 4;
  ctl.mu = O.mu_init;
  ctl.tau = dmax(O.tau_min, 1.0 - ctl.mu);
  ctl.nf = 0; ctl.dw_last = 0.0; ctl.iter = 0; ctl.status = ST_RUNNING;
  int src = 0;
  // evaluate the starting point (alpha = 0 copies buffer 0 -> 1 with the safeguards applied)
  {
    // least-squares multipliers for the defect rows (IPOPT 3.6); discarded if too large
    double dtf0 = 0.0, p00 = 0.0;
    StepInfo s0;
    double al = 0.0;
    if (riccati_backward(P, M, O, W, src, cur, ctl.mu, 0.0, true, &dtf0, &p00)) {
      riccati_forward(P, M, O, W, src, cur, ctl.mu, ctl.tau, dtf0, true, ts, s0);
      if (s0.pimax <= 1e3 * dmax(1.0, O.obj_scale) && fabs(ts.dnu3) <= 1e3 * dmax(1.0, O.obj_scale)) al = 1.0;
#if defined(LMATO_TRACE) && !defined(__CUDA_ARCH__)
      printf("LS multipliers: pimax %.3e dnu3 %.3e dtf %.3e -> %s\n", s0.pimax, ts.dnu3, dtf0, al > 0 ? "used" : "discarded");
#endif
    }
    eval_pass(P, M, O, W, src, 1 - src, cur, ts, ctl.mu, 0.0, 0.0, al, trial);
  }
  cur = trial; src = 1 - src;
  ctl.theta_max = 1e4 * dmax(1.0, cur.theta);
  ctl.theta_min = 1e-4 * dmax(1.0, cur.theta);
  double err0 = 1e300;
  while (true) {
    err0 = kkt_error(cur, 0.0, n_eq, n_bd);
    if (err0 <= O.tol) { ctl.status = ST_CONVERGED; break; }
    if (ctl.iter >= O.max_iter) { ctl.status = ST_MAX_ITER; break; }
    // barrier parameter update (IPOPT eq. 7)
    bool mu_changed = false;
    while (kkt_error(cur, ctl.mu, n_eq, n_bd) <= O.kappa_eps * ctl.mu && ctl.mu > O.tol / 10.0 * (1.0 + 1e-12)) {
      ctl.mu = dmax(O.tol / 10.0, dmin(O.kappa_mu * ctl.mu, pow(ctl.mu, O.theta_mu)));
      ctl.tau = dmax(O.tau_min, 1.0 - ctl.mu);
      ctl.nf = 0;
      mu_changed = true;
    }
    (void)mu_changed;
    // factorisation with inertia correction (IPOPT Algorithm IC)
    double dw = 0.0, dtf = 0.0, p0 = 0.0;
    bool fact_ok = false;
    for (int attempt = 0; attempt < 40; ++attempt) {
      if (riccati_backward(P, M, O, W, src, cur, ctl.mu, dw, false, &dtf, &p0)) { fact_ok = true; break; }
      if (dw == 0.0) dw = (ctl.dw_last == 0.0) ? 1e-4 : dmax(1e-20, ctl.dw_last / 3.0);
      else dw *= (ctl.dw_last == 0.0) ? 100.0 : 8.0;
      if (dw > 1e40) break;
    }
    if (!fact_ok) { ctl.status = ST_INERTIA_FAIL; break; }
    if (dw > 0.0) ctl.dw_last = dw;
    StepInfo si;
    riccati_forward(P, M, O, W, src, cur, ctl.mu, ctl.tau, dtf, false, ts, si);
    if (!(si.dphi == si.dphi) || !(si.dxmax < 1e300)) { ctl.status = ST_NUMERICAL; break; }
    // filter line search (IPOPT section 2.3)
    const double theta = cur.theta;
    const double phi = cur.fobj - ctl.mu * cur.sumlog;
    const double dphi = si.dphi;
    const double g_th = 1e-5, g_ph = 1e-8, s_th = 1.1, s_ph = 2.3, eta = 1e-8, delta = 1.0;
    double alpha = si.a_max;
    bool accepted = false, ftype = false;
    for (int ls = 0; ls < O.max_ls; ++ls) {
      eval_pass(P, M, O, W, src, 1 - src, cur, ts, ctl.mu, alpha, si.a_z, alpha, trial);
      const double th_t = trial.theta;
      const double ph_t = trial.fobj - ctl.mu * trial.sumlog;
      bool ok = (th_t <= ctl.theta_max) && (ph_t == ph_t) && (ph_t < 1e299) && filter_ok(ctl, th_t, ph_t);
      if (ok) {
        ftype = (theta <= ctl.theta_min) && (dphi < 0.0) &&
                (alpha * pow(-dphi, s_ph) > delta * pow(theta, s_th));
        if (ftype) ok = ph_t <= phi + eta * alpha * dphi + 2.2e-15 * fabs(phi);
        else ok = (th_t <= (1.0 - g_th) * theta) || (ph_t <= phi - g_ph * theta);
      }
      if (ok) { accepted = true; break; }
      alpha *= 0.5;
    }
#if defined(LMATO_TRACE) && !defined(__CUDA_ARCH__)
    if (!accepted) printf("LS FAIL theta %.3e phi %.6e dphi %.3e amax %.3e dx %.2e th_t %.3e ph_t %.6e\n", theta, phi, dphi, si.a_max, si.dxmax, trial.theta, trial.fobj - ctl.mu * trial.sumlog);
#endif
    if (!accepted) { ctl.status = ST_LINESEARCH_FAIL; break; }
    if (!ftype) filter_add(ctl, (1.0 - g_th) * theta, phi - g_ph * theta);
    cur = trial; src = 1 - src;
    ++ctl.iter;
#if defined(LMATO_TRACE) && !defined(__CUDA_ARCH__)
    printf("%3d tf %.8f th %.3e err %.3e mu %.1e a %.3e az %.3e dw %.1e dphi %.2e dx %.2e nf %d %s\n", ctl.iter, cur.tf,
           cur.theta, err0, ctl.mu, alpha, si.a_z, dw, dphi, si.dxmax, ctl.nf, ftype ? "f" : "h");
#endif
  }
  out.tf = cur.tf; out.status = ctl.status; out.iters = ctl.iter; out.kkt = err0; out.mu = ctl.mu; out.cur = src;
}

}  // namespace lmato
