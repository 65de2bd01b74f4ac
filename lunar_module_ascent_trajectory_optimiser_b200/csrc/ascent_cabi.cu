// C ABI + kernels of the batched ascent solver (sm_100a).  See include/lmato_b200.h.
//
// Two kernels solve the same NLP with the same IPM driver (ipm_iterate_t, ascent_ipm.cuh):
//   ascent_ipm_kernel   ONE PROBLEM PER THREAD, persistent warps, for batches that fill the GPU (>= ~1 wave of
//                       148 x 256 problems).  The stage data of a problem (55-67 doubles x nt) streams through a
//                       workspace in HBM laid out [stage][warp][field][lane]: every access of a warp is one
//                       coalesced 256-byte row, staged one stage ahead in shared memory with cp.async.  Each sweep
//                       re-evaluates the model, so the HBM traffic stays near 3x the algorithmic minimum.
//   ascent_coop_kernel  EIGHT LANES PER PROBLEM (ascent_coop.cuh), for everything smaller: a single solve, the
//                       1 024 / 4 096-problem configurations, a 65 536 batch strong-scaled over 8 GPUs.  Model
//                       evaluation is stage-parallel over the group and stored per stage; the Riccati recursion
//                       runs on the stored records with the cost-to-go distributed by rows over the group.
// Both claim problems from a device-side queue and never return to the host during a solve (barrier updates,
// fraction-to-boundary, filter line search and convergence masks live in the owning lanes).  Tensor cores are
// not used: the stage blocks are 8x8 FP64 and the recursion over stages is sequential.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <new>
#include <vector>

#include "../../include/lmato_b200.h"
#include "ascent_ipm_dc.cuh"
#include "ascent_coop.cuh"
#include "ascent_colloc.cuh"

using namespace lmato;

namespace {

thread_local char g_err[512] = "";

void set_err(const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
}

#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      set_err("CUDA error: %s at %s", cudaGetErrorString(e__), #expr);        \
      return LMATO_ERR_CUDA;                                                  \
    }                                                                         \
  } while (0)

// One CTA of 256 threads (8 warps) per SM: 255 registers x 256 threads fills the register file.
#ifndef LMATO_SYNC_PERIOD
#define LMATO_SYNC_PERIOD 1
#endif
constexpr int kBlock = 256;
constexpr int kBlocksPerSM = 1;
// dynamic shared memory: the staging tiles, then one workspace view per thread
constexpr size_t kTileSmem = (size_t)(kBlock / 32) * 2 * TILE_ROWS * LANES * sizeof(double) + kBlock * sizeof(Ws);
// Batch warm start: one reference solve of the batch-mean problem (cooperative kernel, one group) against ~6 saved
// iterations per problem.  Measured with the reference solve on the cooperative kernel (1.8 ms on the first call of a
// handle, 0.2 ms afterwards): 512 problems 7.2 ms cold, 7.3 ms warm on the first call, 5.5 ms on later calls; from
// 1 024 problems up the first call wins as well (tools/gpu_warm_threshold.py).
constexpr long kWarmStartMinBatch = 512;

struct KArgs {
  const double* params;  // [NPARAM][B]
  long B;
  double* traj;          // [NVAR][nt][B] or null
  double* tf;
  double* fmass;
  int* status;
  int* iters;
  double* kkt;
  double* sens;          // [LMATO_NSENS][B] d tf / d parameter, or null
  const double* guess_traj;  // caller-supplied start point [NVAR][nt][B] and [B], or null
  const double* guess_tf;
  double* ws;            // [N+1][slots/32][N_FIELDS][32]
  long slots;
  int N;
  const double* h;
  const double* tau;
  int* counter;
  int model;             // lmato_model_t
  double* ref;           // reference column (warm start), or null
  int ref_mode;          // 0 cold start; 1 solve and store the reference (starting from the previous one,
                         // if any); 2 start from the reference
  Options O;
  const colloc::Coll* coll;   // collocation rule in device memory (NODES >= 3 only)
};

// Mean of every parameter row over the batch: the reference problem of the warm start.
__global__ void __launch_bounds__(256) mean_params_kernel(const double* __restrict__ p, long B, double* out) {
  __shared__ double red[256];
  const int row = blockIdx.x;
  double acc = 0.0;
  for (long i = threadIdx.x; i < B; i += blockDim.x) acc += p[row * B + i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[row] = red[0] / (double)B;
}

__device__ __forceinline__ Params derive_params(const double* __restrict__ p, long B, long b, int model, bool move = false) {
  // LO:50-75, 107-109
  Params P;
  const double G = p[LMATO_P_G * B + b], Mm = p[LMATO_P_M * B + b];
  const double R0 = p[LMATO_P_R0 * B + b];
  const double fuel = p[LMATO_P_FUEL_MASS * B + b];
  const double rp = p[LMATO_P_R_PERIAPSIS * B + b], ra = p[LMATO_P_R_APOAPSIS * B + b];
  P.GM = G * Mm;
  P.R0 = R0;
  P.Ft = p[LMATO_P_FT * B + b];
  P.M0 = p[LMATO_P_M0 * B + b];
  P.S = rp;
  P.ms = p[LMATO_P_MASS_SCALAR * B + b];
  P.mflow = p[LMATO_P_M_DOT * B + b] / fuel;
  P.asc = p[LMATO_P_ANGLE_DOUBLEDOT_MAX * B + b] / 3.0;
  P.T = p[LMATO_P_FINAL_TIME * B + b];
  P.a_ub = p[LMATO_P_ANGLE_UB * B + b];
  P.u_ub = p[LMATO_P_U_BOUND * B + b];
  const double vt = sqrt(P.GM / (R0 + 0.5 * (rp + ra)));
  P.vt2 = (vt / P.S) * (vt / P.S);
  P.rt = (R0 + P.S) / P.S;
  P.R0S = R0 / P.S;
  P.tf_ub = fmin(1.0, 1.0 / (P.mflow * P.T));
  P.fuel = fuel;
  P.Sinv = 1.0 / P.S;
  P.coup5 = 1.0;
  P.mT = P.mflow * P.T;
  if (model == LMATO_MODEL_CIRCULAR) {
    // PDF p.27 src 69-73: the MV is the pitch angle itself; no rate or acceleration limit exists
    P.coup5 = 0.0;
    P.asc = move ? 0.0 : 1.0;     // with the move term the MV slot holds the pitch angle itself (mv_is_angle())
    P.u_ub = 1e20;
  }
  return P;
}

// Params padded to an odd number of doubles per thread so that the per-thread structs in shared memory
// are bank-conflict free (Params = 19 doubles -> stride 21 doubles = 42 words; 42 mod 32 = 10).
struct alignas(8) ParamsSlot { Params p; double pad[(sizeof(Params) / 8) % 2 == 0 ? 1 : 2]; };

static_assert(GuessSrc::V_Y == LMATO_V_Y && GuessSrc::V_YDOT == LMATO_V_YDOT && GuessSrc::V_X == LMATO_V_X &&
              GuessSrc::V_XDOT == LMATO_V_XDOT && GuessSrc::V_ANGLE == LMATO_V_ANGLE &&
              GuessSrc::V_ANGLEDOT == LMATO_V_ANGLEDOT && GuessSrc::V_U == LMATO_V_ANGLEDOUBLEDOT,
              "the guess uses the output layout of include/lmato_b200.h");

// Results of one finished problem (LO:178-202: tf, the per-node values of all ten variables, final
// mass, status).  The iterate is read back from the thread's workspace column; ydoubledot,
// xdoubledot and mass, which the device formulation eliminates, are recomputed (LO:123, 127-136).
template <class SW>
__device__ __noinline__ void write_results(const KArgs& a, const Params& P, const Ws& W, const IpmState& S, long b) {
  SolveOut out;
  ipm_result(S, out);
  const int nt = a.N + 1;
  a.tf[b] = out.tf;
  a.fmass[b] = P.M0 - P.fuel * (P.mflow * P.T * out.tf);   // mass(nt-1) = mflow*T*tf (LO:123)
  a.status[b] = out.status;
  a.iters[b] = out.iters;
  if (a.kkt) a.kkt[b] = out.kkt;
  if (!a.traj && !a.sens) return;
  double* __restrict__ t = a.traj;
  const long B = a.B;
  if (t) {
#pragma unroll
    for (int v = 0; v < LMATO_NVAR; ++v) t[((long)v * nt) * B + b] = 0.0;   // node 0 pinned
  }
  // Parametric sensitivity of the optimum (envelope theorem): d objective* / d theta = d Lagrangian / d theta
  // = sum_k lambda_k . d c_k / d theta at the solution, with c_k = s_k - s_{k-1} - h_k T tf F(s_k, u_k; theta).
  // theta enters only the velocity rows (through the thrust acceleration Ft / (M0 - ms*mass), LO:127-136,
  // and mass = mflow*T*tau*tf, LO:123) and the angledot row (angle_scalar, LO:121).
  double s_ft = 0.0, s_m0 = 0.0, s_mf = 0.0, s_asc = 0.0;
  for (int k = 1; k <= a.N; ++k) {
    const double* sp = W.stage(k);
    double z[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) z[i] = WS_AT(sp, out.cur * SW::NITER + SW::FZ + i);
    const double u = WS_AT(sp, out.cur * SW::NITER + SW::FU);
    const double m = P.mflow * P.T * a.tau[k] * out.tf;
    double ay, ax;
    if (a.sens) {
      Accel1 f;
      accel_first(P, z[0], z[2], z[4], m, f);
      ay = f.ay; ax = f.ax;
      const double al = a.h[k] * P.T * out.tf;
      const double l1 = WS_AT(sp, out.cur * SW::NITER + SW::FLAM + 1);     // multiplier of the ydot row
      const double l3 = WS_AT(sp, out.cur * SW::NITER + SW::FLAM + 3);     // xdot row
      const double l5 = WS_AT(sp, out.cur * SW::NITER + SW::FLAM + 5);     // angledot row
      const double thrust = l1 * (f.AT * f.Ty * P.Sinv) + l3 * (f.AT * f.Tx * P.Sinv);
      s_ft -= al * thrust / P.Ft;
      s_m0 += al * thrust * (f.AT / P.Ft);                                  // d AT / d M0 = -AT / (M0 - ms*mass)
      s_mf -= al * (l1 * f.ay_m + l3 * f.ax_m) * (P.T * a.tau[k] * out.tf);
      s_asc -= al * u * l5;
    } else {
      accel_value(P, z[0], z[2], z[4], m, ay, ax);
    }
    if (!t) continue;
    t[((long)LMATO_V_Y * nt + k) * B + b] = z[0];
    t[((long)LMATO_V_YDOT * nt + k) * B + b] = z[1];
    t[((long)LMATO_V_YDOUBLEDOT * nt + k) * B + b] = ay;
    t[((long)LMATO_V_X * nt + k) * B + b] = z[2];
    t[((long)LMATO_V_XDOT * nt + k) * B + b] = z[3];
    t[((long)LMATO_V_XDOUBLEDOT * nt + k) * B + b] = ax;
    t[((long)LMATO_V_ANGLE * nt + k) * B + b] = z[4];
    t[((long)LMATO_V_ANGLEDOT * nt + k) * B + b] = z[5];
    t[((long)LMATO_V_MASS * nt + k) * B + b] = m;
    t[((long)LMATO_V_ANGLEDOUBLEDOT * nt + k) * B + b] = u;
  }
  if (a.sens) {
    // per unit of the RAW parameters of include/lmato_b200.h: M_dot = mflow * fuel_mass (LO:65),
    // angle_doubledot_max = 3 * angle_scalar (LO:109); tf in the reference's scaled units (0..1)
    const double inv = 1.0 / a.O.obj_scale;
    // the mass bound 0 <= mass <= 1 (LO:83) is carried as tf <= tf_ub = 1 / (mflow T): where that is the binding
    // upper bound on tf, its multiplier contributes  -zUt * d tf_ub / d mflow = zUt / (mflow^2 T)
    if (P.tf_ub < 1.0) s_mf += S.cur.zUt / (P.mflow * P.mflow * P.T);
    a.sens[(long)LMATO_S_FT * B + b] = s_ft * inv;
    a.sens[(long)LMATO_S_M0 * B + b] = s_m0 * inv;
    a.sens[(long)LMATO_S_M_DOT * B + b] = s_mf * inv / P.fuel;
    a.sens[(long)LMATO_S_ANGLE_DOUBLEDOT_MAX * B + b] = s_asc * inv / 3.0;
  }
}

// SW = Sweeps7 (dcost = 0) or Sweeps8 (with the reference's move-suppression term, LO:99)
template <class SW>
__global__ void __launch_bounds__(kBlock, kBlocksPerSM) ascent_ipm_kernel(KArgs a) {
  // Per-problem constants are read in every stage of every sweep.  They live in shared memory: as
  // registers they would be spilled (255 are already in use), and a spill is reloaded from local
  // memory, which misses the thrashed L1 every stage.
  __shared__ ParamsSlot sP[kBlock];
  __shared__ Options sO;
  __shared__ Mesh sM;
  if (threadIdx.x == 0) { sO = a.O; sM = Mesh{a.N, a.h, a.tau}; }
  __syncthreads();
  const Options& O = sO;
  const Mesh& M = sM;
  Params& P = sP[threadIdx.x].p;
  const long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const long nwarps = a.slots / LANES;
  // staging tiles: [warp][2 buffers][TILE_ROWS][32 lanes] doubles of dynamic shared memory
  extern __shared__ __align__(16) double sTiles[];
  const unsigned tile0 = (unsigned)__cvta_generic_to_shared(sTiles) +
                         (unsigned)(((threadIdx.x / 32) * 2 * TILE_ROWS * LANES + lane) * sizeof(double));
  // the workspace view is read inside every stage as well: shared memory for the same reason
  // (24-byte stride: conflict-free per half warp)
  Ws* sW = reinterpret_cast<Ws*>(sTiles + (size_t)(kBlock / 32) * 2 * TILE_ROWS * LANES);   // after the tiles
  sW[threadIdx.x] = Ws{a.ws + ((slot / LANES) * SW::NFIELDS) * LANES + lane, nwarps * SW::NFIELDS * LANES, tile0};
  const Ws& W = sW[threadIdx.x];
  IpmState S;
  bool active = false;        // this lane holds an unfinished problem
  bool pending = false;       // this lane holds a finished problem whose results are not written yet
  bool exhausted = false;     // the queue is empty (warp-uniform)
  bool first = true;
  long b = -1;
  const int warps_per_block = kBlock / 32;
  unsigned round = 0;
  while (true) {
    // ---- a warp whose 32 problems are all finished claims the next 32 (warp-uniform branch) ----
    const bool chunk_done = !__any_sync(0xffffffffu, active);
    // ---- a finished chunk writes its results with all 32 lanes together: every row of the
    //      [variable][node][problem] output is then one coalesced 256-byte store per warp (the output
    //      may be pinned host memory written over PCIe while the other warps keep computing) ----
    if (chunk_done && __any_sync(0xffffffffu, pending)) {
      if (pending) write_results<SW>(a, P, W, S, b);
      pending = false;
    }
    if (!exhausted && chunk_done) {
      // first chunk: static round-robin over CTAs (spreads a small batch over all SMs);
      // afterwards: the device-wide queue
      int chunk = 0;
      if (first) {
        chunk = (threadIdx.x / 32) * gridDim.x + blockIdx.x;
        first = false;
      } else {
        if (lane == 0) chunk = atomicAdd(a.counter, 1) + gridDim.x * warps_per_block;
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
      }
      if ((long)chunk * 32 >= a.B) {
        exhausted = true;
      } else {
        b = (long)chunk * 32 + lane;
        if (b < a.B) {
          P = derive_params(a.params, a.B, b, a.model);
          ipm_begin(O, S);
          double mu0 = 0.0;
          // ref_mode 2: start from the batch reference.  ref_mode 1 (the reference solve itself):
          // start from the previous call's reference if the handle has one (consecutive batches of a
          // campaign have nearly the same mean, so it re-converges in one or two iterations).
          if (a.guess_traj) {
            SW::guess_from(P, M, O, W, GuessSrc{a.guess_traj, a.guess_tf, a.B, b}, S.cur);
            S.from_guess = true;
          } else if (a.ref_mode != 0 && SW::load_ref(P, M, O, W, a.ref, S.cur, &mu0)) {
            S.warm = true;
            S.ctl.mu = mu0;
            S.ctl.tau = dmax(O.tau_min, 1.0 - mu0);
          } else {
            SW::guess(P, M, O, W, S.cur);
          }
          active = true;
        }
      }
    }
    // ---- all eight warps of the SM start each iteration together: they then run the same sweep
    //      (same code) at the same time, which keeps the 32 KB instruction cache effective ----
    // (every LMATO_SYNC_PERIOD-th round: a barrier per iteration costs more in waiting than it
    //  gains, because line-search retries differ between warps; the imbalance averages out over
    //  a few iterations while the warps stay close enough to share the instruction cache)
    if ((round++ % LMATO_SYNC_PERIOD) == 0 && !__syncthreads_or((active || !exhausted) ? 1 : 0)) break;
    if (active && ipm_iterate_t<SW>(P, M, O, W, S)) {
      if ((S.warm || S.from_guess) && S.ctl.status != ST_CONVERGED) {
        // a warm start, or a caller's start point (lmato_set_initial_guess), that did not work out: this lane
        // restarts the same problem from the built-in roll-out.  (IPOPT would enter its feasibility restoration
        // phase from such a point, e.g. from the reference's all-zero start values LO:39, 83-96; this solver's
        // substitute is a start point that is dynamically feasible by construction.)
        { const int prior = S.iters_prior + S.ctl.iter; ipm_begin(O, S); S.iters_prior = prior; }
        SW::guess(P, M, O, W, S.cur);
        continue;
      }
      active = false;
      pending = true;      // results are written when the whole chunk is done (see above)
      // (the reference solve of the batch warm start, ref_mode 1, runs on the cooperative kernel)
    }
  }
}

// ---------------------------------------------------------------------------------------
// cooperative kernel: G lanes per problem (ascent_coop.cuh)
// ---------------------------------------------------------------------------------------
constexpr int kCoopBlock = 256;
constexpr int kCoopBlocksPerSM = 2;
constexpr int kCoopG = 8;
constexpr int kCoopScrStride = coop::SCR_DOUBLES;       // per group: record ring + transpose tiles (even: 16-byte aligned)
static size_t coop_smem(int gp) { return sizeof(double) * (size_t)(kCoopBlock / gp) * kCoopScrStride; }

template <int G, class CW>
__device__ __noinline__ void coop_write_results(const KArgs& a, const Params& P, const CW& W, const IpmState& S, long b) {
  SolveOut out;
  ipm_result(S, out);
  const int nt = a.N + 1;
  const long B = a.B;
  if (W.g == 0) {
    a.tf[b] = out.tf;
    a.fmass[b] = P.M0 - P.fuel * (P.mflow * P.T * out.tf);   // mass(nt-1) = mflow*T*tf (LO:123)
    a.status[b] = out.status;
    a.iters[b] = out.iters;
    if (a.kkt) a.kkt[b] = out.kkt;
  }
  if (!a.traj && !a.sens) return;
  double* __restrict__ t = a.traj;
  if (t) {
    for (int v = W.g; v < LMATO_NVAR; v += G) t[((long)v * nt) * B + b] = 0.0;   // node 0 pinned
  }
  // (see write_results above for the sensitivity formulas)
  double s_ft = 0.0, s_m0 = 0.0, s_mf = 0.0, s_asc = 0.0;
  for (int k = 1 + W.g; k <= a.N; k += G) {
    double x[coop::XR];
    coop::ldv<coop::XR>(W.X(out.cur, k), x);
    const double* z = x + coop::X_Z;
    const double u = x[coop::X_U];
    const double m = P.mflow * P.T * a.tau[k] * out.tf;
    double ay, ax;
    if (a.sens) {
      Accel1 f;
      accel_first(P, z[0], z[2], z[4], m, f);
      ay = f.ay; ax = f.ax;
      const double al = a.h[k] * P.T * out.tf;
      const double l1 = x[coop::X_LAM + 1], l3 = x[coop::X_LAM + 3], l5 = x[coop::X_LAM + 5];
      const double thrust = l1 * (f.AT * f.Ty * P.Sinv) + l3 * (f.AT * f.Tx * P.Sinv);
      s_ft -= al * thrust / P.Ft;
      s_m0 += al * thrust * (f.AT / P.Ft);
      s_mf -= al * (l1 * f.ay_m + l3 * f.ax_m) * (P.T * a.tau[k] * out.tf);
      s_asc -= al * u * l5;
    } else {
      accel_value(P, z[0], z[2], z[4], m, ay, ax);
    }
    if (!t) continue;
    t[((long)LMATO_V_Y * nt + k) * B + b] = z[0];
    t[((long)LMATO_V_YDOT * nt + k) * B + b] = z[1];
    t[((long)LMATO_V_YDOUBLEDOT * nt + k) * B + b] = ay;
    t[((long)LMATO_V_X * nt + k) * B + b] = z[2];
    t[((long)LMATO_V_XDOT * nt + k) * B + b] = z[3];
    t[((long)LMATO_V_XDOUBLEDOT * nt + k) * B + b] = ax;
    t[((long)LMATO_V_ANGLE * nt + k) * B + b] = z[4];
    t[((long)LMATO_V_ANGLEDOT * nt + k) * B + b] = z[5];
    t[((long)LMATO_V_MASS * nt + k) * B + b] = m;
    t[((long)LMATO_V_ANGLEDOUBLEDOT * nt + k) * B + b] = u;
  }
  if (a.sens) {
    s_ft = coop::Grp<G>::sum(W.mask, s_ft); s_m0 = coop::Grp<G>::sum(W.mask, s_m0);
    s_mf = coop::Grp<G>::sum(W.mask, s_mf); s_asc = coop::Grp<G>::sum(W.mask, s_asc);
    if (P.tf_ub < 1.0) s_mf += S.cur.zUt / (P.mflow * P.mflow * P.T);     // see write_results()
    if (W.g == 0) {
      const double inv = 1.0 / a.O.obj_scale;
      a.sens[(long)LMATO_S_FT * B + b] = s_ft * inv;
      a.sens[(long)LMATO_S_M0 * B + b] = s_m0 * inv;
      a.sens[(long)LMATO_S_M_DOT * B + b] = s_mf * inv / P.fuel;
      a.sens[(long)LMATO_S_ANGLE_DOUBLEDOT_MAX * B + b] = s_asc * inv / 3.0;
    }
  }
}

// A warp holds 32/GP problems; its groups run the IPM driver in lock step (same sweep at the same time, each
// with its own lane mask), finished groups idle until the warp's chunk is done, then the warp claims the next
// chunk from the device-wide queue.  IpmState (filter, barrier parameter, ...) is carried redundantly by the
// GP lanes of a group: every decision of the driver is a function of group-uniform values.
//   GP = 8 : four problems per warp, for batches that can fill the GPU that way;
//   GP = 32: the whole warp works on one problem (the stage-parallel phases are four times shorter), for
//            the latency regime: up to a few problems per SM.
//   BPS = CTAs per SM: 1 leaves 255 registers per thread (no spills), 2 halves them for twice the problems in flight.
template <int G, int GP, int MOVE, int BPS>      // MOVE: 0 no move term, 1 move term (LO:99), 2 the circular model's (MV = angle)
__global__ void __launch_bounds__(kCoopBlock, BPS) ascent_coop_kernel(KArgs a) {
  using SW = SweepsCoop<G, GP, MOVE>;
  constexpr int PPW = 32 / GP;                 // problems per warp
  constexpr int GPB = kCoopBlock / GP;         // groups per block
  __shared__ ParamsSlot sP[GPB];
  __shared__ Options sO;
  __shared__ Mesh sM;
  extern __shared__ __align__(16) double sScr[];    // [GPB][kCoopScrStride]
  if (threadIdx.x == 0) { sO = a.O; sM = Mesh{a.N, a.h, a.tau}; }
  __syncthreads();
  const Options& O = sO;
  const Mesh& M = sM;
  const int lane = threadIdx.x & 31;
  const int grp = threadIdx.x / GP;
  const unsigned gmask = GP == 32 ? 0xffffffffu : (((1u << GP) - 1u) << ((lane / GP) * GP));
  const unsigned smask = ((1u << G) - 1u) << ((lane / GP) * GP);
  Params& P = sP[grp].p;
  const long slot = (long)blockIdx.x * GPB + grp;
  // workspace layout by regime (ascent_coop.cuh, CwsT): stage-major blocks when a whole warp walks one problem
  const coop::CwsT<(GP == 32)> W{a.ws + slot * coop::coop_doubles_per_problem(a.N + 1), a.N + 1, sScr + grp * kCoopScrStride,
                                 lane % GP, gmask, smask, 0.0, 0.0, 0, 0u};
  if (W.g == 0) coop::ring_init_barriers(W.bars());      // the mbarriers of this group's record ring
  __syncthreads();
  IpmState S;
  bool active = false, pending = false, exhausted = false, first = true;
  long b = -1;
  const int warps_per_block = kCoopBlock / 32;
  while (true) {
    const bool chunk_done = !__any_sync(0xffffffffu, active);
    if (chunk_done && __any_sync(0xffffffffu, pending)) {
      if (pending) coop_write_results<GP>(a, P, W, S, b);
      pending = false;
    }
    if (!exhausted && chunk_done) {
      int chunk = 0;
      if (first) {
        chunk = (threadIdx.x / 32) * gridDim.x + blockIdx.x;     // static round-robin: spreads a small batch over all SMs
        first = false;
      } else {
        if (lane == 0) chunk = atomicAdd(a.counter, 1) + gridDim.x * warps_per_block;
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
      }
      if ((long)chunk * PPW >= a.B) {
        exhausted = true;
      } else {
        b = (long)chunk * PPW + lane / GP;
        if (b < a.B) {
          if (W.g == 0) P = derive_params(a.params, a.B, b, a.model, MOVE == 2);
          coop::Grp<GP>::sync(gmask);
          ipm_begin(O, S);
          double mu0 = 0.0;
          if (a.guess_traj) {
            SW::guess_from(P, M, O, W, GuessSrc{a.guess_traj, a.guess_tf, a.B, b}, S.cur);
            S.from_guess = true;
          } else if (a.ref_mode != 0 && SW::load_ref(P, M, O, W, a.ref, S.cur, &mu0)) {
            S.warm = true;
            S.ctl.mu = mu0;
            S.ctl.tau = dmax(O.tau_min, 1.0 - mu0);
          } else {
            SW::guess(P, M, O, W, S.cur);
          }
          active = true;
        }
      }
    }
    // (no block barrier per iteration here: the warps of an SM are few and their sweeps are short; a warp that
    //  waits for its neighbours' line-search retries loses more than the instruction cache gains)
    if (exhausted && !__any_sync(0xffffffffu, active || pending)) break;
    if (active && ipm_iterate_t<SW>(P, M, O, W, S)) {
      if ((S.warm || S.from_guess) && S.ctl.status != ST_CONVERGED) {
        // a warm start or a caller's guess that did not work out: restart from the built-in roll-out
        { const int prior = S.iters_prior + S.ctl.iter; ipm_begin(O, S); S.iters_prior = prior; }
        SW::guess(P, M, O, W, S.cur);
        continue;
      }
      active = false;
      if (a.ref_mode == 1) {
        SolveOut out;
        ipm_result(S, out);
        SW::store_ref(P, M, W, out.cur, S.cur, S.ctl.mu, out.status == ST_CONVERGED, a.ref);
        continue;
      }
      pending = true;
    }
  }
}

template <int GP, int BPS>
static void coop_launch(int move, long grid, cudaStream_t st, const KArgs& a) {
  if (move == 2) ascent_coop_kernel<kCoopG, GP, 2, BPS><<<(int)grid, kCoopBlock, coop_smem(GP), st>>>(a);
  else if (move == 1) ascent_coop_kernel<kCoopG, GP, 1, BPS><<<(int)grid, kCoopBlock, coop_smem(GP), st>>>(a);
  else ascent_coop_kernel<kCoopG, GP, 0, BPS><<<(int)grid, kCoopBlock, coop_smem(GP), st>>>(a);
}
template <int GP, int BPS>
static cudaError_t coop_optin() {
  cudaError_t e = cudaFuncSetAttribute(ascent_coop_kernel<kCoopG, GP, 2, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coop_smem(GP));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(ascent_coop_kernel<kCoopG, GP, 1, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coop_smem(GP));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ascent_coop_kernel<kCoopG, GP, 0, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coop_smem(GP));
}

// ---------------------------------------------------------------------------------------
// higher-order collocation kernel (NODES = 3..6, ascent_colloc.cuh): GP lanes per problem
// ---------------------------------------------------------------------------------------
template <int GP>
__device__ __noinline__ void colloc_write_results(const KArgs& a, const Params& P, const colloc::Nws& W, const IpmState& S, long b) {
  SolveOut out;
  ipm_result(S, out);
  const int nt = a.N + 1, m = W.L.m;
  const long B = a.B;
  if (W.g == 0) {
    a.tf[b] = out.tf;
    a.fmass[b] = P.M0 - P.fuel * (P.mflow * P.T * out.tf);
    a.status[b] = out.status;
    a.iters[b] = out.iters;
    if (a.kkt) a.kkt[b] = out.kkt;
  }
  double* __restrict__ t = a.traj;
  if (!t) return;
  for (int v = W.g; v < LMATO_NVAR; v += GP) t[((long)v * nt) * B + b] = 0.0;   // node 0 pinned
  for (int k = 1 + W.g; k <= a.N; k += GP) {
    const double* xr = W.X(out.cur, k);
    const double* z = xr + 6 * (m - 1);                  // the mesh node is the last collocation point of its step
    const double ms = P.mflow * P.T * a.tau[k] * out.tf;
    double ay, ax;
    accel_value(P, z[0], z[2], z[4], ms, ay, ax);
    t[((long)LMATO_V_Y * nt + k) * B + b] = z[0];
    t[((long)LMATO_V_YDOT * nt + k) * B + b] = z[1];
    t[((long)LMATO_V_YDOUBLEDOT * nt + k) * B + b] = ay;
    t[((long)LMATO_V_X * nt + k) * B + b] = z[2];
    t[((long)LMATO_V_XDOT * nt + k) * B + b] = z[3];
    t[((long)LMATO_V_XDOUBLEDOT * nt + k) * B + b] = ax;
    t[((long)LMATO_V_ANGLE * nt + k) * B + b] = z[4];
    t[((long)LMATO_V_ANGLEDOT * nt + k) * B + b] = z[5];
    t[((long)LMATO_V_MASS * nt + k) * B + b] = ms;
    t[((long)LMATO_V_ANGLEDOUBLEDOT * nt + k) * B + b] = xr[W.L.x_u];
  }
}

template <int GP>
__global__ void __launch_bounds__(kCoopBlock, 1) ascent_colloc_kernel(KArgs a) {
  using SW = SweepsColloc<GP>;
  constexpr int PPW = 32 / GP;
  constexpr int GPB = kCoopBlock / GP;
  __shared__ ParamsSlot sP[GPB];
  __shared__ Options sO;
  __shared__ Mesh sM;
  __shared__ colloc::Coll sC;
  if (threadIdx.x == 0) { sO = a.O; sM = Mesh{a.N, a.h, a.tau}; sC = *a.coll; }
  __syncthreads();
  const Options& O = sO;
  const Mesh& M = sM;
  const int lane = threadIdx.x & 31;
  const int grp = threadIdx.x / GP;
  const unsigned gmask = GP == 32 ? 0xffffffffu : (((1u << GP) - 1u) << ((lane / GP) * GP));
  Params& P = sP[grp].p;
  const long slot = (long)blockIdx.x * GPB + grp;
  colloc::Nws W;
  W.L.init(sC.m);
  W.base = a.ws + slot * colloc::colloc_doubles_per_problem(a.N + 1, sC.m);
  W.N1 = a.N + 1; W.C = &sC; W.g = lane % GP; W.mask = gmask; W.dw = 0.0; W.pimax = 0.0; W.ls_flag = 0;
  IpmState S;
  int variant = 0;
  bool active = false, pending = false, exhausted = false, first = true;
  long b = -1;
  const int warps_per_block = kCoopBlock / 32;
  while (true) {
    const bool chunk_done = !__any_sync(0xffffffffu, active);
    if (chunk_done && __any_sync(0xffffffffu, pending)) {
      if (pending) colloc_write_results<GP>(a, P, W, S, b);
      pending = false;
    }
    if (!exhausted && chunk_done) {
      int chunk = 0;
      if (first) {
        chunk = (threadIdx.x / 32) * gridDim.x + blockIdx.x;
        first = false;
      } else {
        if (lane == 0) chunk = atomicAdd(a.counter, 1) + gridDim.x * warps_per_block;
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
      }
      if ((long)chunk * PPW >= a.B) {
        exhausted = true;
      } else {
        b = (long)chunk * PPW + lane / GP;
        if (b < a.B) {
          if (W.g == 0) P = derive_params(a.params, a.B, b, a.model);
          coop::Grp<GP>::sync(gmask);
          ipm_begin(O, S);
          variant = 0;
          SW::guess_variant(P, M, O, W, S.cur, variant);
          active = true;
        }
      }
    }
    if (exhausted && !__any_sync(0xffffffffu, active || pending)) break;
    if (active && ipm_iterate_t<SW>(P, M, O, W, S)) {
      if (S.ctl.status != ST_CONVERGED && S.ctl.status != ST_MAX_ITER && variant + 1 < (int)SW::N_STARTS) {
        // no restoration phase: a problem that fails from one start point is restarted from the next
        const int it_used = S.ctl.iter;
        ipm_begin(O, S);
        S.ctl.iter = it_used; S.ctl.iter_best = it_used;      // MAX_ITER (LO:28) bounds the whole ladder
        SW::guess_variant(P, M, O, W, S.cur, ++variant);
        continue;
      }
      active = false;
      pending = true;
    }
  }
}

// FP64 FMA peak: 8 independent chains per thread, no memory traffic.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Post-solve orbit check of the reference PDF (p.28-29 src 185-237): explicit Euler two-body coast,
// one problem per thread, everything in registers.  state/out: [4][B] / [6][B], SI, Moon-centred.
__global__ void __launch_bounds__(128) coast_orbit_kernel(const double* __restrict__ state, long B, double gm,
                                                          double dt, long nsteps, double* __restrict__ out) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x = state[0 * B + b], y = state[1 * B + b], vx = state[2 * B + b], vy = state[3 * B + b];
  double r2 = fma(x, x, y * y);
  double r2min = r2, r2max = r2;
  for (long i = 0; i < nsteps; ++i) {
    const double inv = lm_rsqrt(r2);
    const double k = gm * inv * inv * inv;
    const double ax = -k * x, ay = -k * y;            // src 207-208
    x = fma(vx, dt, x);                               // src 231-232 (old velocity)
    y = fma(vy, dt, y);
    vx = fma(ax, dt, vx);                             // src 233-234
    vy = fma(ay, dt, vy);
    r2 = fma(x, x, y * y);
    r2min = fmin(r2min, r2);
    r2max = fmax(r2max, r2);
  }
  out[0 * B + b] = sqrt(r2min); out[1 * B + b] = sqrt(r2max);
  out[2 * B + b] = x; out[3 * B + b] = y; out[4 * B + b] = vx; out[5 * B + b] = vy;
}

// Self-test of the branch-free math against the CUDA library: max relative errors of
// rcp, rsqrt, log over [1e-40, 1e3] and max absolute errors of sin, cos over [0, 3.5].
__global__ void math_selftest_kernel(double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = (double)i / (double)(n - 1);
  const double x = exp(log(1e-40) + t * (log(1e3) - log(1e-40)));
  const double a = 3.5 * t;
  double s, c;
  lm_sincos_small(a, &s, &c);
  const double lx = log(x);
  out[i * 5 + 0] = fabs(lm_rcp(x) * x - 1.0);
  out[i * 5 + 1] = fabs(lm_rsqrt(x) * sqrt(x) - 1.0);
  out[i * 5 + 2] = fabs(lm_log_pos(x) - lx) / fmax(fabs(lx), 1e-3);
  out[i * 5 + 3] = fabs(s - sin(a));
  out[i * 5 + 4] = fabs(c - cos(a));
}

}  // namespace

struct lmato_handle {
  int device = 0;
  int nt = 0;
  int nodes = 2;
  int model = 0;
  int sm_count = 0;
  double* d_h = nullptr;
  double* d_tau = nullptr;
  double* d_ws = nullptr;
  size_t ws_bytes = 0;
  long ws_slots = 0;
  int* d_counter = nullptr;
  // staging for the host-buffer entry point
  double* d_params = nullptr; size_t params_bytes = 0;
  double* d_out = nullptr; size_t out_bytes = 0;
  double* d_ref = nullptr;        // reference column of the warm start: [REF_ROWS][nt]
  double* d_refparams = nullptr;  // [NPARAM] batch-mean parameters + scratch outputs of the reference solve
  lmato_options opt;
  double* sens_out = nullptr;     // optional extra output of the next solves (lmato_set_sensitivity_output)
  const double* guess_traj = nullptr;   // optional start point of the next solves (lmato_set_initial_guess)
  const double* guess_tf = nullptr;
  double* d_guess = nullptr; size_t guess_bytes = 0;   // staging of a host guess
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t last_stream = nullptr;
  bool timed = false;
  bool last_coop = false;         // which kernel the last solve used
  colloc::Coll coll;              // collocation rule for NODES >= 3 ...
  colloc::Coll* d_coll = nullptr; // ... and its device copy
};

// Collocation rule of GEKKO NODES = n: Lobatto points on [0,1] including both ends;
// h N f_{1..m} = z_{1..m} - z_0 with N_ij = int_0^{tau_i} l_j(s) ds, l_j the Lagrange basis on tau_1..tau_m.
static bool collocation_rule(int nodes, colloc::Coll* C) {
  if (nodes < 2 || nodes > 6) return false;
  const int m = nodes - 1;
  double pts[6];                         // Lobatto points on [-1, 1]
  switch (nodes) {
    case 2: pts[0] = -1; pts[1] = 1; break;
    case 3: pts[0] = -1; pts[1] = 0; pts[2] = 1; break;
    case 4: { const double a = sqrt(1.0 / 5.0); pts[0] = -1; pts[1] = -a; pts[2] = a; pts[3] = 1; break; }
    case 5: { const double a = sqrt(3.0 / 7.0); pts[0] = -1; pts[1] = -a; pts[2] = 0; pts[3] = a; pts[4] = 1; break; }
    default: {
      const double a = sqrt(1.0 / 3.0 - 2.0 * sqrt(7.0) / 21.0), b = sqrt(1.0 / 3.0 + 2.0 * sqrt(7.0) / 21.0);
      pts[0] = -1; pts[1] = -b; pts[2] = -a; pts[3] = a; pts[4] = b; pts[5] = 1; break;
    }
  }
  double tau[6];
  for (int i = 0; i < nodes; ++i) tau[i] = 0.5 * (pts[i] + 1.0);
  C->m = m;
  for (int i = 0; i < colloc::MAXM; ++i) { C->tau[i] = 0.0; for (int j = 0; j < colloc::MAXM; ++j) C->N[i][j] = 0.0; }
  for (int i = 0; i < m; ++i) C->tau[i] = tau[i + 1];
  for (int j = 0; j < m; ++j) {
    double c[6] = {1, 0, 0, 0, 0, 0};    // coefficients of l_j, ascending powers
    int deg = 0;
    for (int q = 0; q < m; ++q) {
      if (q == j) continue;
      const double den = tau[j + 1] - tau[q + 1];
      double n[6] = {0, 0, 0, 0, 0, 0};
      for (int d = 0; d <= deg; ++d) { n[d + 1] += c[d] / den; n[d] -= c[d] * tau[q + 1] / den; }
      ++deg;
      for (int d = 0; d <= deg; ++d) c[d] = n[d];
    }
    for (int i = 0; i < m; ++i) {
      double s = 0.0, tp = tau[i + 1];
      for (int d = 0; d <= deg; ++d) { s += c[d] * tp / (double)(d + 1); tp *= tau[i + 1]; }
      C->N[i][j] = s;
    }
  }
  return true;
}

// device-side part of lmato_create (every failure path returns through the caller's lmato_destroy)
static lmato_status_t create_device_state(lmato_handle* H, const std::vector<double>& h, const std::vector<double>& tau) {
  const int nt = H->nt;
  CUDA_TRY(cudaSetDevice(H->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, H->device));
  H->sm_count = prop.multiProcessorCount;
  // the staging tiles need the opt-in shared-memory size (158 KB dynamic + 42 KB static per CTA)
  CUDA_TRY(cudaFuncSetAttribute(ascent_ipm_kernel<Sweeps7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem));
  CUDA_TRY(cudaFuncSetAttribute(ascent_ipm_kernel<Sweeps8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem));
  CUDA_TRY((coop_optin<8, 1>()));
  CUDA_TRY((coop_optin<8, 2>()));
  CUDA_TRY((coop_optin<32, 1>()));
  CUDA_TRY((coop_optin<32, 2>()));
  CUDA_TRY(cudaMalloc(&H->d_h, sizeof(double) * nt));
  CUDA_TRY(cudaMalloc(&H->d_tau, sizeof(double) * nt));
  CUDA_TRY(cudaMemcpy(H->d_h, h.data(), sizeof(double) * nt, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(H->d_tau, tau.data(), sizeof(double) * nt, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&H->d_counter, sizeof(int)));
  CUDA_TRY(cudaMalloc(&H->d_ref, sizeof(double) * (size_t)dc::REF_ROWS * nt));   // dc:: is the larger layout
  CUDA_TRY(cudaMemset(H->d_ref, 0, sizeof(double) * (size_t)dc::REF_ROWS * nt));   // REF_OK = 0: no reference yet
  CUDA_TRY(cudaMalloc(&H->d_refparams, sizeof(double) * (LMATO_NPARAM + 8)));
  CUDA_TRY(cudaMalloc(&H->d_coll, sizeof(colloc::Coll)));
  CUDA_TRY(cudaMemcpy(H->d_coll, &H->coll, sizeof(colloc::Coll), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaEventCreate(&H->ev0));
  CUDA_TRY(cudaEventCreate(&H->ev1));
  return LMATO_OK;
}

extern "C" {

const char* lmato_last_error(void) { return g_err; }
const char* lmato_version(void) { return "lmato_b200 0.1 (sm_100a)"; }

void lmato_default_options(lmato_options* o) {
  if (!o) return;
  o->tol = 1e-10;
  o->mu_init = 0.1;
  o->obj_scale = 10.0;
  o->tf_guess = 0.9;
  o->delta_c = 1e-8;
  o->mu_min_factor = 1e-3;
  o->n_polish = -1;           // automatic: 2 with the DCOST term, 4 without (DESIGN.md "Tolerance")
  o->warm_start = 1;
  o->mu_ref = 1e-3;
  o->dcost = 1e-5;            // LO:99
  o->kappa_eps = 30.0;
  o->objective_nodes = 0;     // 0 = nt - 1
  o->kernel = LMATO_KERNEL_AUTO;
  o->coop_lanes = 0;
  o->otol = 1e-3;             // LO:31
  o->rtol = 1e-3;             // LO:32
  o->max_iter = 20000;   // LO:28
  o->max_ls = 40;
}

lmato_status_t lmato_create(lmato_handle** out, int32_t device, int32_t nt, const double* time,
                            int32_t nodes, int32_t model) {
  if (!out) { set_err("lmato_create: out is NULL"); return LMATO_ERR_INVALID; }
  *out = nullptr;
  if (nt < 2) { set_err("lmato_create: nt must be >= 2"); return LMATO_ERR_INVALID; }
  if (nodes < 2 || nodes > 6) { set_err("lmato_create: NODES must be 2..6 (GEKKO's range, LO:25)"); return LMATO_ERR_INVALID; }
  if (model != LMATO_MODEL_ELLIPTICAL && model != LMATO_MODEL_CIRCULAR) {
    set_err("lmato_create: unknown model");
    return LMATO_ERR_INVALID;
  }
  if (nodes > 2 && model != LMATO_MODEL_ELLIPTICAL) {
    set_err("lmato_create: NODES > 2 is implemented for the elliptical model (Launch_Optimiser.py) only");
    return LMATO_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_err("lmato_create: no CUDA device available (%s); there is no CPU fallback",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return LMATO_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) { set_err("lmato_create: bad device ordinal"); return LMATO_ERR_INVALID; }
  std::vector<double> h(nt), tau(nt);
  for (int k = 0; k < nt; ++k) tau[k] = time ? time[k] : (double)k / (double)(nt - 1);   // LO:21
  if (tau[0] != 0.0) { set_err("lmato_create: time[0] must be 0"); return LMATO_ERR_INVALID; }
  h[0] = 0.0;
  for (int k = 1; k < nt; ++k) {
    h[k] = tau[k] - tau[k - 1];
    if (!(h[k] > 0.0)) { set_err("lmato_create: time must be strictly increasing"); return LMATO_ERR_INVALID; }
  }
  lmato_handle* H = new (std::nothrow) lmato_handle();
  if (!H) { set_err("lmato_create: out of host memory"); return LMATO_ERR_INVALID; }
  H->device = device; H->nt = nt; H->nodes = nodes; H->model = model;
  lmato_default_options(&H->opt);
  collocation_rule(nodes, &H->coll);
  const lmato_status_t rc = create_device_state(H, h, tau);
  if (rc != LMATO_OK) { lmato_destroy(H); return rc; }      // frees whatever was allocated before the failure
  *out = H;
  return LMATO_OK;
}

lmato_status_t lmato_destroy(lmato_handle* h) {
  if (!h) return LMATO_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_h); cudaFree(h->d_tau); cudaFree(h->d_ws); cudaFree(h->d_counter);
  cudaFree(h->d_params); cudaFree(h->d_out); cudaFree(h->d_ref); cudaFree(h->d_refparams); cudaFree(h->d_guess); cudaFree(h->d_coll);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return LMATO_OK;
}

lmato_status_t lmato_set_options(lmato_handle* h, const lmato_options* o) {
  if (!h || !o) { set_err("lmato_set_options: NULL argument"); return LMATO_ERR_INVALID; }
  if (!(o->tol > 0) || !(o->mu_init > 0) || !(o->obj_scale > 0) || !(o->delta_c > 0) ||
      !(o->tf_guess > 0 && o->tf_guess < 1) || o->max_iter < 0 || o->max_ls < 1 ||
      !(o->mu_min_factor > 0 && o->mu_min_factor <= 1) || o->n_polish < -1 ||
      (o->warm_start < 0 || o->warm_start > 2) || !(o->mu_ref > 0 && o->mu_ref <= o->mu_init) ||
      !(o->dcost >= 0) || o->objective_nodes < 0 || !(o->kappa_eps >= 1.0) ||
      o->kernel < LMATO_KERNEL_AUTO || o->kernel > LMATO_KERNEL_COOP || !(o->otol >= 0) || !(o->rtol >= 0) ||
      (o->coop_lanes != 0 && o->coop_lanes != 8 && o->coop_lanes != 32)) {
    set_err("lmato_set_options: option out of range");
    return LMATO_ERR_INVALID;
  }
  h->opt = *o;
  return LMATO_OK;
}

// The move-suppression term: DCOST on the MV angledoubledot of the elliptical model (LO:99), on the MV angle of the
// circular model (PDF p.27 src 69-73).  The circular model with the term is carried by the cooperative kernel only.
static bool dcost_active(const lmato_handle* h) { return h->opt.dcost > 0.0; }
static bool circular_move(const lmato_handle* h) { return h->model == LMATO_MODEL_CIRCULAR && dcost_active(h); }

// Which kernel solves a batch of B problems.  One thread per problem needs ~1 wave (SMs x 256 problems) to fill
// the GPU and streams the least HBM traffic per problem; eight lanes per problem fill it with an eighth of that
// and shorten the critical path of every problem, at ~2.5x the traffic.  Measured cross-over on B200 at nt = 200:
// between 16 384 and 32 768 problems (profiles/README.md).
constexpr int64_t kCoopMaxBatch = 6144;
static bool use_coop(const lmato_handle* h, int64_t B) {
  if (circular_move(h)) return true;
  if (h->opt.kernel == LMATO_KERNEL_COOP) return true;
  if (h->opt.kernel == LMATO_KERNEL_THREAD) return false;
  return B <= kCoopMaxBatch;
}
static int fields_for(const lmato_handle* h) { return dcost_active(h) ? (int)dc::N_FIELDS : (int)N_FIELDS; }

static long slots_for(const lmato_handle* h, int64_t B) {
  // one CTA per SM, but never more CTAs than 32-problem chunks
  const long chunks = (B + 31) / 32;
  const long grid = chunks < (long)h->sm_count * kBlocksPerSM ? (chunks > 0 ? chunks : 1) : (long)h->sm_count * kBlocksPerSM;
  return grid * kBlock;
}
// cooperative kernel: lanes per problem in the stage-parallel phases.  A whole warp per problem while that still
// gives every SM at most ~8 warps (the latency regime), otherwise 8 lanes (four problems per warp).
static int coop_gp_for(const lmato_handle* h, int64_t B) {
  if (h->opt.coop_lanes == 8 || h->opt.coop_lanes == 32) return h->opt.coop_lanes;
  return B <= (int64_t)h->sm_count * 8 ? 32 : 8;
}
// CTAs per SM: one (255 registers per thread) while one CTA per SM holds the whole batch, else two
static int coop_bps_for(const lmato_handle* h, int64_t B, int gp) {
  return B <= (int64_t)h->sm_count * (kCoopBlock / gp) ? 1 : kCoopBlocksPerSM;
}
// CTAs of 256 threads = 256/GP groups; a warp's chunk is 32/GP problems
static long coop_grid_for(const lmato_handle* h, int64_t B, int gp) {
  const long chunks = (B + (32 / gp) - 1) / (32 / gp);
  const long cap = (long)h->sm_count * coop_bps_for(h, B, gp);
  return chunks < cap ? (chunks > 0 ? chunks : 1) : cap;
}
static size_t coop_ws_bytes(const lmato_handle* h, long grid, int gp) {
  return sizeof(double) * (size_t)coop::coop_doubles_per_problem(h->nt) * (size_t)(grid * (kCoopBlock / gp));
}
static int colloc_gp_for(const lmato_handle* h, int64_t B) { return B <= (int64_t)h->sm_count * 8 ? 32 : 8; }
static long colloc_grid_for(const lmato_handle* h, int64_t B, int gp) {
  const long chunks = (B + (32 / gp) - 1) / (32 / gp);
  const long cap = (long)h->sm_count;
  return chunks < cap ? (chunks > 0 ? chunks : 1) : cap;
}
static size_t ws_bytes_for(const lmato_handle* h, int64_t B) {
  if (h->nodes > 2) {
    const int gp = colloc_gp_for(h, B);
    return sizeof(double) * (size_t)colloc::colloc_doubles_per_problem(h->nt, h->nodes - 1) *
           (size_t)(colloc_grid_for(h, B, gp) * (kCoopBlock / gp));
  }
  if (use_coop(h, B)) { const int gp = coop_gp_for(h, B); return coop_ws_bytes(h, coop_grid_for(h, B, gp), gp); }
  return sizeof(double) * (size_t)fields_for(h) * (size_t)h->nt * (size_t)slots_for(h, B);
}

lmato_status_t lmato_workspace_bytes(lmato_handle* h, int64_t B, int64_t* bytes) {
  if (!h || !bytes || B < 0) { set_err("lmato_workspace_bytes: bad argument"); return LMATO_ERR_INVALID; }
  *bytes = (int64_t)ws_bytes_for(h, B);
  return LMATO_OK;
}

lmato_status_t lmato_kernel_launches(lmato_handle* h, int64_t* n) {
  if (!h || !n) { set_err("lmato_kernel_launches: bad argument"); return LMATO_ERR_INVALID; }
  *n = h->launches;
  return LMATO_OK;
}

// What kind of memory a registered optional buffer (sensitivity output, start point) points at.  The device entry
// point needs memory the kernel can address (device, managed, or pinned host memory); the host entry point needs
// memory the host can address.  Answers true when the runtime cannot tell (plain host memory is "unregistered").
static bool kernel_can_address(const void* p) {
  if (!p) return true;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged ||
         (attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr);
}
static bool host_can_address(const void* p) {
  if (!p) return true;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return attr.type != cudaMemoryTypeDevice;
}

lmato_status_t lmato_solve_batch(lmato_handle* h, const double* params, int64_t B,
                                 double* out_traj, double* out_tf, double* out_final_mass,
                                 int32_t* out_status, int32_t* out_iters, double* out_kkt,
                                 void* stream) {
  if (!h) { set_err("lmato_solve_batch: NULL handle"); return LMATO_ERR_INVALID; }
  if (B < 0) { set_err("lmato_solve_batch: negative batch"); return LMATO_ERR_INVALID; }
  if (B == 0) return LMATO_OK;
  if (!params || !out_tf || !out_final_mass || !out_status || !out_iters) {
    set_err("lmato_solve_batch: NULL buffer");
    return LMATO_ERR_INVALID;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (!kernel_can_address(h->sens_out) || !kernel_can_address(h->guess_traj) || !kernel_can_address(h->guess_tf)) {
    set_err("lmato_solve_batch: the registered sensitivity output / start point is not device-addressable memory "
            "(lmato_solve_batch takes DEVICE pointers; use lmato_solve_batch_host for host buffers)");
    return LMATO_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // Every solve on a handle shares its workspace, work-queue counter and warm-start reference.  A solve issued on
  // a different stream than the previous one first waits (on the device) for that one to finish.
  if (h->timed && st != h->last_stream) CUDA_TRY(cudaStreamWaitEvent(st, h->ev1, 0));
  const bool hi_order = h->nodes > 2;        // NODES >= 3: the general collocation kernel
  if (hi_order && (h->sens_out || h->guess_traj)) {
    set_err("lmato_solve_batch: sensitivities and caller-supplied start points are implemented for NODES = 2 only");
    return LMATO_ERR_UNSUPPORTED;
  }
  const bool coopk = !hi_order && use_coop(h, B);
  const long slots = slots_for(h, B);
  const int gp = coop_gp_for(h, B);
  const long cgrid = coop_grid_for(h, B, gp);
  const bool use_dc = dcost_active(h) && !hi_order;   // the l1 move term is carried by the NODES = 2 kernels only
  // (no batch warm start for the circular model with its move term: the reference column of the 7-state solve
  //  carries a different MV)
  const bool warm = !hi_order && !h->guess_traj && !circular_move(h) &&
                    (h->opt.warm_start == 2 || (h->opt.warm_start == 1 && B >= kWarmStartMinBatch));
  size_t need = ws_bytes_for(h, B);
  if (warm) { const size_t r = coop_ws_bytes(h, 1, 32); if (r > need) need = r; }     // the reference solve: one warp of one CTA
  if (need > h->ws_bytes) {
    if (h->d_ws) {
      if (h->timed) CUDA_TRY(cudaEventSynchronize(h->ev1));     // the previous solve may still be using it
      CUDA_TRY(cudaFree(h->d_ws)); h->d_ws = nullptr; h->ws_bytes = 0;
    }
    CUDA_TRY(cudaMalloc(&h->d_ws, need));
    h->ws_bytes = need;
  }
  h->ws_slots = slots;
  CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
  KArgs a;
  a.params = params; a.B = B; a.traj = out_traj; a.tf = out_tf; a.fmass = out_final_mass;
  a.status = out_status; a.iters = out_iters; a.kkt = out_kkt; a.sens = h->sens_out;
  a.guess_traj = h->guess_traj; a.guess_tf = h->guess_tf;
  a.ws = h->d_ws; a.slots = slots; a.N = h->nt - 1; a.h = h->d_h; a.tau = h->d_tau;
  a.counter = h->d_counter;
  a.model = h->model;
  // OTOL / RTOL (LO:31-32): how APMonitor maps them onto IPOPT's `tol` is not verifiable here, so the rule is
  // the conservative one: the solve runs to the tightest of tol, otol and rtol (defaults 1e-10, 1e-3, 1e-3).
  double tol = h->opt.tol;
  if (h->opt.otol > 0.0 && h->opt.otol < tol) tol = h->opt.otol;
  if (h->opt.rtol > 0.0 && h->opt.rtol < tol) tol = h->opt.rtol;
  a.O.tol = tol; a.O.mu_init = h->opt.mu_init; a.O.obj_scale = h->opt.obj_scale;
  // barrier decrease exponent: IPOPT's 1.5 from the cold start, 2 from the batch warm start.  Measured on the
  // benchmark batch (warm): 1.3 / 1.5 / 1.7 / 2.0 / 2.5 / 3.0 -> 91 / 96 / 95 / 88 / 88 / 115 ms; from the cold
  // start 2.0 costs 0.7 iterations more than 1.5.  kappa_mu and tau_min do not matter.
  a.O.kappa_eps = h->opt.kappa_eps; a.O.kappa_mu = 0.2; a.O.theta_mu = 1.5; a.O.tau_min = 0.99;
  a.O.delta_c = h->opt.delta_c; a.O.tf_guess = h->opt.tf_guess;
  a.O.max_iter = h->opt.max_iter; a.O.max_ls = h->opt.max_ls;
  a.O.mu_min_factor = h->opt.mu_min_factor;
  a.O.n_polish = h->opt.n_polish >= 0 ? h->opt.n_polish : (use_dc ? 2 : 4);
  // ... 2 from the warm start WITH the move term only.  Without it the control on the singular arc is a nearly flat
  // direction, and the jump 1e-6 -> 1e-12 leaves a few problems in 10 000 on an error plateau until the stall guard
  // gives the warm attempt up -- each holding its warp: 65 536 problems, dcost = 0: 260-560 ms with 2, 66 ms with 1.5.
  a.O.theta_mu_warm = use_dc ? 2.0 : 1.5;
  {
    const int on = h->opt.objective_nodes > 0 ? h->opt.objective_nodes : h->nt - 1;
    a.O.w_dcost = use_dc ? h->opt.obj_scale * h->opt.dcost / (double)on : 0.0;
  }
  a.ref = nullptr; a.ref_mode = 0;
  CUDA_TRY(cudaEventRecord(h->ev0, st));
  // (a caller-supplied guess replaces the batch warm start)
  if (warm) {
    // reference problem = batch mean, solved down to mu_ref only by ONE group of the cooperative kernel without
    // the move term (at mu_ref >> w it is immaterial): ~10 iterations from the cold start on the first call,
    // 1-2 from the previous call's reference afterwards
    mean_params_kernel<<<LMATO_NPARAM, 256, 0, st>>>(params, B, h->d_refparams);
    CUDA_TRY(cudaGetLastError());
    KArgs r = a;
    double* scratch = h->d_refparams + LMATO_NPARAM;
    r.params = h->d_refparams; r.B = 1; r.traj = nullptr; r.tf = scratch; r.fmass = scratch + 1;
    r.status = (int*)(scratch + 2); r.iters = (int*)(scratch + 3); r.kkt = scratch + 4; r.sens = nullptr;
    r.guess_traj = nullptr; r.guess_tf = nullptr;
    r.ref = h->d_ref; r.ref_mode = 1;
    r.O.tol = 10.0 * h->opt.mu_ref; r.O.mu_min_factor = 0.1; r.O.n_polish = 0;
    r.O.w_dcost = 0.0;
    coop_launch<32, 1>(0, 1, st, r);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
    a.ref = h->d_ref; a.ref_mode = 2;
    h->launches += 2;
  }
  a.coll = h->d_coll;
  if (hi_order) {
    const int cgp = colloc_gp_for(h, B);
    const int cg = (int)colloc_grid_for(h, B, cgp);
    if (cgp == 32) ascent_colloc_kernel<32><<<cg, kCoopBlock, 0, st>>>(a);
    else ascent_colloc_kernel<8><<<cg, kCoopBlock, 0, st>>>(a);
  } else if (coopk) {
    const int bps = coop_bps_for(h, B, gp);
    const int mv = !use_dc ? 0 : circular_move(h) ? 2 : 1;
    if (gp == 32) { if (bps == 1) coop_launch<32, 1>(mv, cgrid, st, a); else coop_launch<32, 2>(mv, cgrid, st, a); }
    else          { if (bps == 1) coop_launch<8, 1>(mv, cgrid, st, a); else coop_launch<8, 2>(mv, cgrid, st, a); }
  } else {
    const int grid = (int)(slots / kBlock);
    if (use_dc) ascent_ipm_kernel<Sweeps8><<<grid, kBlock, kTileSmem, st>>>(a);
    else ascent_ipm_kernel<Sweeps7><<<grid, kBlock, kTileSmem, st>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(h->ev1, st));
  h->last_stream = st; h->timed = true;
  h->last_coop = coopk;
  h->launches += 1;
  return LMATO_OK;
}

lmato_status_t lmato_last_kernel_ms(lmato_handle* h, double* ms) {
  if (!h || !ms) { set_err("lmato_last_kernel_ms: bad argument"); return LMATO_ERR_INVALID; }
  if (!h->timed) { set_err("lmato_last_kernel_ms: no solve has been launched"); return LMATO_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  float f = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&f, h->ev0, h->ev1));
  *ms = (double)f;
  return LMATO_OK;
}

// Device-usable alias of a host buffer if the caller pinned it (cudaHostAlloc / cudaHostRegister /
// torch pin_memory), else nullptr.  With unified addressing pinned host memory is mapped into the
// device's address space, so the kernel can store results straight into it.
static double* pinned_alias(double* host) {
  if (!host) return nullptr;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return nullptr;
  return static_cast<double*>(attr.devicePointer);
}

lmato_status_t lmato_solve_batch_host(lmato_handle* h, const double* params, int64_t B,
                                      double* out_traj, double* out_tf, double* out_final_mass,
                                      int32_t* out_status, int32_t* out_iters, double* out_kkt) {
  if (!h) { set_err("lmato_solve_batch_host: NULL handle"); return LMATO_ERR_INVALID; }
  if (B < 0) { set_err("lmato_solve_batch_host: negative batch"); return LMATO_ERR_INVALID; }
  if (B == 0) return LMATO_OK;
  if (!params || !out_tf || !out_final_mass || !out_status || !out_iters) {
    set_err("lmato_solve_batch_host: NULL buffer");
    return LMATO_ERR_INVALID;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t pbytes = sizeof(double) * LMATO_NPARAM * (size_t)B;
  if (pbytes > h->params_bytes) {
    if (h->d_params) CUDA_TRY(cudaFree(h->d_params));
    h->d_params = nullptr; h->params_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->d_params, pbytes));
    h->params_bytes = pbytes;
  }
  // The per-node trajectories are 16 KB per problem (1 GB for 65 536 problems).  If the caller's
  // buffer is pinned the kernel writes them there directly, chunk by chunk as problems finish, so the
  // transfer over PCIe overlaps the solve; a pageable buffer is filled by a copy after the kernel.
  double* traj_alias = pinned_alias(out_traj);
  const size_t traj_n = out_traj ? (size_t)LMATO_NVAR * h->nt * (size_t)B : 0;
  const size_t traj_stage = traj_alias ? 0 : traj_n;
  if (!host_can_address(h->sens_out) || !host_can_address(h->guess_traj) || !host_can_address(h->guess_tf)) {
    set_err("lmato_solve_batch_host: the registered sensitivity output / start point is device memory "
            "(lmato_solve_batch_host takes HOST pointers)");
    return LMATO_ERR_INVALID;
  }
  double* sens_host = h->sens_out;       // for this entry point the registered pointer is a HOST buffer
  const size_t sens_n = sens_host ? (size_t)LMATO_NSENS * (size_t)B : 0;
  // layout of d_out: traj (only when staged) | sens | tf | fmass | kkt | status | iters
  const size_t obytes = sizeof(double) * (traj_stage + sens_n + 3 * (size_t)B) + sizeof(int32_t) * 2 * (size_t)B;
  if (obytes > h->out_bytes) {
    if (h->d_out) CUDA_TRY(cudaFree(h->d_out));
    h->d_out = nullptr; h->out_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->d_out, obytes));
    h->out_bytes = obytes;
  }
  double* d_traj = out_traj ? (traj_alias ? traj_alias : h->d_out) : nullptr;
  double* d_sens = sens_host ? h->d_out + traj_stage : nullptr;
  double* d_tf = h->d_out + traj_stage + sens_n;
  double* d_fm = d_tf + B;
  double* d_kkt = d_fm + B;
  int32_t* d_st = (int32_t*)(d_kkt + B);
  int32_t* d_it = d_st + B;
  cudaStream_t st = nullptr;
  CUDA_TRY(cudaMemcpyAsync(h->d_params, params, pbytes, cudaMemcpyHostToDevice, st));
  // a registered start point is a HOST buffer for this entry point: stage it
  const double* guess_traj_host = h->guess_traj;
  const double* guess_tf_host = h->guess_tf;
  if (guess_traj_host) {
    const size_t gn = (size_t)LMATO_NVAR * h->nt * (size_t)B;
    const size_t gbytes = sizeof(double) * (gn + (size_t)B);
    if (gbytes > h->guess_bytes) {
      if (h->d_guess) CUDA_TRY(cudaFree(h->d_guess));
      h->d_guess = nullptr; h->guess_bytes = 0;
      CUDA_TRY(cudaMalloc(&h->d_guess, gbytes));
      h->guess_bytes = gbytes;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_guess, guess_traj_host, sizeof(double) * gn, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->d_guess + gn, guess_tf_host, sizeof(double) * (size_t)B, cudaMemcpyHostToDevice, st));
    h->guess_traj = h->d_guess; h->guess_tf = h->d_guess + gn;
  }
  h->sens_out = d_sens;
  lmato_status_t rc = lmato_solve_batch(h, h->d_params, B, d_traj, d_tf, d_fm, d_st, d_it, d_kkt, st);
  h->sens_out = sens_host;
  h->guess_traj = guess_traj_host; h->guess_tf = guess_tf_host;
  if (rc != LMATO_OK) return rc;
  if (sens_host) CUDA_TRY(cudaMemcpyAsync(sens_host, d_sens, sizeof(double) * sens_n, cudaMemcpyDeviceToHost, st));
  if (out_traj && !traj_alias)
    CUDA_TRY(cudaMemcpyAsync(out_traj, d_traj, sizeof(double) * traj_n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_tf, d_tf, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_final_mass, d_fm, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  if (out_kkt) CUDA_TRY(cudaMemcpyAsync(out_kkt, d_kkt, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_status, d_st, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_iters, d_it, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return LMATO_OK;
}

lmato_status_t lmato_collocation_rule(int32_t nodes, double* tau, double* N) {
  colloc::Coll C;
  if (!tau || !N || !collocation_rule(nodes, &C)) { set_err("lmato_collocation_rule: NODES must be 2..6"); return LMATO_ERR_INVALID; }
  for (int i = 0; i < C.m; ++i) { tau[i] = C.tau[i]; for (int j = 0; j < C.m; ++j) N[i * C.m + j] = C.N[i][j]; }
  return LMATO_OK;
}

lmato_status_t lmato_set_initial_guess(lmato_handle* h, const double* guess_traj, const double* guess_tf) {
  if (!h) { set_err("lmato_set_initial_guess: NULL handle"); return LMATO_ERR_INVALID; }
  if ((guess_traj == nullptr) != (guess_tf == nullptr)) {
    set_err("lmato_set_initial_guess: pass both the trajectories and tf, or NULL for both");
    return LMATO_ERR_INVALID;
  }
  h->guess_traj = guess_traj; h->guess_tf = guess_tf;
  return LMATO_OK;
}

lmato_status_t lmato_set_sensitivity_output(lmato_handle* h, double* out_dtf) {
  if (!h) { set_err("lmato_set_sensitivity_output: NULL handle"); return LMATO_ERR_INVALID; }
  h->sens_out = out_dtf;
  return LMATO_OK;
}

lmato_status_t lmato_coast_orbit(lmato_handle* h, const double* state, int64_t B, double gm, double dt,
                                 int64_t nsteps, double* out, void* stream) {
  if (!h || B < 0 || nsteps < 0 || !(dt > 0) || !(gm > 0)) { set_err("lmato_coast_orbit: bad argument"); return LMATO_ERR_INVALID; }
  if (B == 0) return LMATO_OK;
  if (!state || !out) { set_err("lmato_coast_orbit: NULL buffer"); return LMATO_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  coast_orbit_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(state, (long)B, gm, dt, (long)nsteps, out);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return LMATO_OK;
}

lmato_status_t lmato_selftest_math(lmato_handle* h, double* max_err5) {
  if (!h || !max_err5) { set_err("lmato_selftest_math: bad argument"); return LMATO_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(h->device));
  const int n = 1 << 18;
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double) * 5 * n));
  math_selftest_kernel<<<(n + 255) / 256, 256>>>(d, n);
  CUDA_TRY(cudaGetLastError());
  std::vector<double> hbuf((size_t)5 * n);
  CUDA_TRY(cudaMemcpy(hbuf.data(), d, sizeof(double) * 5 * n, cudaMemcpyDeviceToHost));
  cudaFree(d);
  h->launches += 1;
  for (int j = 0; j < 5; ++j) max_err5[j] = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 5; ++j)
      if (!(hbuf[(size_t)i * 5 + j] <= max_err5[j])) max_err5[j] = hbuf[(size_t)i * 5 + j];   // NaN-propagating max
  return LMATO_OK;
}

lmato_status_t lmato_measure_fp64_peak(lmato_handle* h, double* gflops) {
  if (!h || !gflops) { set_err("lmato_measure_fp64_peak: bad argument"); return LMATO_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(h->device));
  const int blocks = h->sm_count * 8, threads = 256, iters = 20000;
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, 0));
    dfma_peak_kernel<<<blocks, threads>>>(d, iters);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    const double g = fl / (ms * 1e-3) * 1e-9;
    if (rep > 0 && g > best) best = g;
  }
  h->launches += 5;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *gflops = best;
  return LMATO_OK;
}


// ---------------------------------------------------------------------------------------
// device list: one host process, several GPUs.  Problem i goes to device floor(i*G/B)
// (contiguous index ranges); every device solves its shard with its own handle, concurrently (all launches
// are asynchronous), and writes its slice of the caller's HOST result arrays.  The problems are independent
// and the results land on the host, so there is no collective to own here; the one-process-per-GPU path with
// its single NCCL allgather is `sharded_solve` in the Python host layer.
// ---------------------------------------------------------------------------------------
struct lmato_multi {
  std::vector<lmato_handle*> h;
  std::vector<cudaStream_t> st;
  std::vector<double*> d_params;   // per device [NPARAM][Bshard]
  std::vector<double*> d_out;      // per device: traj | tf | fmass | kkt | status | iters
  std::vector<size_t> params_bytes, out_bytes;
};

lmato_status_t lmato_multi_destroy(lmato_multi* m) {
  if (!m) return LMATO_OK;
  for (size_t g = 0; g < m->h.size(); ++g) {
    if (m->h[g]) cudaSetDevice(m->h[g]->device);
    if (g < m->st.size() && m->st[g]) cudaStreamDestroy(m->st[g]);
    if (g < m->d_params.size()) cudaFree(m->d_params[g]);
    if (g < m->d_out.size()) cudaFree(m->d_out[g]);
    lmato_destroy(m->h[g]);
  }
  delete m;
  return LMATO_OK;
}

lmato_status_t lmato_multi_create(lmato_multi** out, const int32_t* devices, int32_t ndev, int32_t nt,
                                  const double* time, int32_t nodes, int32_t model) {
  if (!out) { set_err("lmato_multi_create: out is NULL"); return LMATO_ERR_INVALID; }
  *out = nullptr;
  if (!devices || ndev < 1) { set_err("lmato_multi_create: empty device list"); return LMATO_ERR_INVALID; }
  lmato_multi* m = new (std::nothrow) lmato_multi();
  if (!m) { set_err("lmato_multi_create: out of host memory"); return LMATO_ERR_INVALID; }
  m->h.assign(ndev, nullptr); m->st.assign(ndev, nullptr);
  m->d_params.assign(ndev, nullptr); m->d_out.assign(ndev, nullptr);
  m->params_bytes.assign(ndev, 0); m->out_bytes.assign(ndev, 0);
  for (int g = 0; g < ndev; ++g) {
    lmato_status_t rc = lmato_create(&m->h[g], devices[g], nt, time, nodes, model);
    if (rc == LMATO_OK && (cudaSetDevice(devices[g]) != cudaSuccess ||
                           cudaStreamCreateWithFlags(&m->st[g], cudaStreamNonBlocking) != cudaSuccess)) {
      set_err("lmato_multi_create: could not create a stream on device %s", g == 0 ? "0" : "n");
      rc = LMATO_ERR_CUDA;
    }
    if (rc != LMATO_OK) { lmato_multi_destroy(m); return rc; }
  }
  *out = m;
  return LMATO_OK;
}

lmato_status_t lmato_multi_set_options(lmato_multi* m, const lmato_options* o) {
  if (!m || !o) { set_err("lmato_multi_set_options: NULL argument"); return LMATO_ERR_INVALID; }
  for (lmato_handle* h : m->h) {
    const lmato_status_t rc = lmato_set_options(h, o);
    if (rc != LMATO_OK) return rc;
  }
  return LMATO_OK;
}

lmato_status_t lmato_multi_device_count(lmato_multi* m, int32_t* n) {
  if (!m || !n) { set_err("lmato_multi_device_count: NULL argument"); return LMATO_ERR_INVALID; }
  *n = (int32_t)m->h.size();
  return LMATO_OK;
}

lmato_status_t lmato_multi_solve_host(lmato_multi* m, const double* params, int64_t B,
                                      double* out_traj, double* out_tf, double* out_final_mass,
                                      int32_t* out_status, int32_t* out_iters, double* out_kkt) {
  if (!m) { set_err("lmato_multi_solve_host: NULL handle"); return LMATO_ERR_INVALID; }
  if (B < 0) { set_err("lmato_multi_solve_host: negative batch"); return LMATO_ERR_INVALID; }
  if (B == 0) return LMATO_OK;
  if (!params || !out_tf || !out_final_mass || !out_status || !out_iters) {
    set_err("lmato_multi_solve_host: NULL buffer");
    return LMATO_ERR_INVALID;
  }
  const int G = (int)m->h.size();
  const int nt = m->h[0]->nt;
  struct Shard { int64_t lo, n; double *traj, *tf, *fm, *kkt; int32_t *st, *it; };
  std::vector<Shard> sh(G);
  // phase 1: every device gets its parameter columns and its launch (nothing here waits for a device)
  for (int g = 0; g < G; ++g) {
    lmato_handle* h = m->h[g];
    Shard& s = sh[g];
    s.lo = (B * g) / G;
    s.n = (B * (g + 1)) / G - s.lo;
    if (s.n == 0) continue;
    CUDA_TRY(cudaSetDevice(h->device));
    // the shards of one batch keep the batch warm start even when a shard alone is small
    const int32_t ws_saved = h->opt.warm_start;
    if (ws_saved == 1 && B >= kWarmStartMinBatch) h->opt.warm_start = 2;
    const size_t pbytes = sizeof(double) * LMATO_NPARAM * (size_t)s.n;
    if (pbytes > m->params_bytes[g]) {
      CUDA_TRY(cudaStreamSynchronize(m->st[g]));
      cudaFree(m->d_params[g]); m->d_params[g] = nullptr; m->params_bytes[g] = 0;
      CUDA_TRY(cudaMalloc(&m->d_params[g], pbytes));
      m->params_bytes[g] = pbytes;
    }
    const size_t traj_n = out_traj ? (size_t)LMATO_NVAR * nt * (size_t)s.n : 0;
    const size_t obytes = sizeof(double) * (traj_n + 3 * (size_t)s.n) + sizeof(int32_t) * 2 * (size_t)s.n;
    if (obytes > m->out_bytes[g]) {
      CUDA_TRY(cudaStreamSynchronize(m->st[g]));
      cudaFree(m->d_out[g]); m->d_out[g] = nullptr; m->out_bytes[g] = 0;
      CUDA_TRY(cudaMalloc(&m->d_out[g], obytes));
      m->out_bytes[g] = obytes;
    }
    s.traj = out_traj ? m->d_out[g] : nullptr;
    s.tf = m->d_out[g] + traj_n; s.fm = s.tf + s.n; s.kkt = s.fm + s.n;
    s.st = (int32_t*)(s.kkt + s.n); s.it = s.st + s.n;
    // columns [lo, lo+n) of the row-major [NPARAM][B] block
    CUDA_TRY(cudaMemcpy2DAsync(m->d_params[g], sizeof(double) * s.n, params + s.lo, sizeof(double) * B,
                               sizeof(double) * s.n, LMATO_NPARAM, cudaMemcpyHostToDevice, m->st[g]));
    const lmato_status_t rc = lmato_solve_batch(h, m->d_params[g], s.n, s.traj, s.tf, s.fm, s.st, s.it, s.kkt, m->st[g]);
    h->opt.warm_start = ws_saved;
    if (rc != LMATO_OK) return rc;
  }
  // phase 2: every device writes its slice of the host arrays
  for (int g = 0; g < G; ++g) {
    const Shard& s = sh[g];
    if (s.n == 0) continue;
    CUDA_TRY(cudaSetDevice(m->h[g]->device));
    cudaStream_t st = m->st[g];
    if (out_traj)
      CUDA_TRY(cudaMemcpy2DAsync(out_traj + s.lo, sizeof(double) * B, s.traj, sizeof(double) * s.n, sizeof(double) * s.n,
                                 (size_t)LMATO_NVAR * nt, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_tf + s.lo, s.tf, sizeof(double) * s.n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_final_mass + s.lo, s.fm, sizeof(double) * s.n, cudaMemcpyDeviceToHost, st));
    if (out_kkt) CUDA_TRY(cudaMemcpyAsync(out_kkt + s.lo, s.kkt, sizeof(double) * s.n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_status + s.lo, s.st, sizeof(int32_t) * s.n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_iters + s.lo, s.it, sizeof(int32_t) * s.n, cudaMemcpyDeviceToHost, st));
  }
  for (int g = 0; g < G; ++g) {
    if (sh[g].n == 0) continue;
    CUDA_TRY(cudaSetDevice(m->h[g]->device));
    CUDA_TRY(cudaStreamSynchronize(m->st[g]));
  }
  return LMATO_OK;
}

}  // extern "C"
