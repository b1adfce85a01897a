// Shared device helpers and parameter blocks for the kdot kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kdot.h"

namespace kdot {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegBig = -1.0e30f;  // "minus infinity" that survives (a - b) without NaN
constexpr float kLogZeroWeight = -100000.0f;  // geomloss log_weights(): log(0) stand-in

// Per-round constants of the eps-scaling loop, derived in float64 and rounded once.
//   v_ij (log2 domain) = h_j + coef * |x_i - y_j|^2          coef  = -0.5*log2(e)/eps      (p = 2)
//   new potential      = scale * log2-sum-exp                 scale = -lambda(eps)*eps*ln 2
//   h_j for NEXT round = lw2_j + pot_j * hmul                 hmul  = log2(e)/eps_next
struct RoundConst {
  float coef;
  float scale;
  float hmul;
  float eps;
};

struct SinkhornParams {
  float* xs;
  const float* ws;
  float* xt;
  const float* wt;
  const int32_t* cu_n;
  const int32_t* cu_m;
  int nimg, B;
  // element strides (in cells) of a (cell, slot) pair inside xs/ws and xt/wt
  long long s_cell_n, s_slot_n, s_cell_m, s_slot_m;
  double p, blur, scaling;
  double rho;  // reach^p, < 0 -> balanced
  float w, h;
  int normalize;
  float* loss_per_img;
  float* loss_per_slot;
  int32_t* valid;
  float* grad_xs;
  float* grad_ws;
  int32_t* nits_per_img;
  // tiled path scratch
  RoundConst* sched;     // [nimg][KDOT_MAX_ROUNDS]
  int32_t* sched_rounds; // [nimg] number of rounds (nits + 2) or <0 status
  float* slot_loss;      // [nimg][B]
  unsigned int* done_ctr; // [nimg]
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// geomloss max_diameter on fp32 data: |maxs - mins|_2 evaluated without FMA contraction.
__device__ __forceinline__ float bbox_diameter(float minx, float miny, float maxx, float maxy) {
  const float ex = __fsub_rn(maxx, minx), ey = __fsub_rn(maxy, miny);
  return sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
}

// Number of geomloss schedule entries: len([diam^p] + arange(p ln diam, p ln blur, p ln scaling) + [blur^p]).
__device__ __forceinline__ int schedule_len(double diam, double p, double blur, double scaling, double* start,
                                            double* delta) {
  const double a = p * log(diam), b = p * log(blur), s = p * log(scaling);
  double len = ceil((b - a) / s);
  if (!(len > 0.0)) len = 0.0;
  *start = a;
  *delta = (a + s) - a;  // numpy arange fills with start + i * ((start + step) - start)
  return (int)len + 2;
}

// eps of geomloss' eps_s[t], t in [0, nits)
__device__ __forceinline__ double schedule_eps(int t, int nits, double diam, double p, double blur, double start,
                                               double delta) {
  if (t <= 0) return pow(diam, p);
  if (t >= nits - 1) return pow(blur, p);
  return exp(start + (double)(t - 1) * delta);
}

// Round r of the kernel's flattened loop: r = 0 init (eps_s[0]), 1..nits the loop, nits+1 the last extrapolation.
__device__ __forceinline__ int round_to_sched(int r, int nits) {
  if (r == 0) return 0;
  if (r <= nits) return r - 1;
  return nits - 1;
}

__device__ __forceinline__ RoundConst make_round_const(int r, int nits, double diam, double p, double blur,
                                                       double start, double delta, double rho) {
  const double eps = schedule_eps(round_to_sched(r, nits), nits, diam, p, blur, start, delta);
  const double eps_next = schedule_eps(round_to_sched(r + 1, nits), nits, diam, p, blur, start, delta);
  const double lam = rho < 0.0 ? 1.0 : 1.0 / (1.0 + eps / rho);
  RoundConst rc;
  rc.coef = (float)(-0.5 * 1.4426950408889634 / eps);
  rc.scale = (float)(-lam * eps * 0.6931471805599453);
  rc.hmul = (float)(1.4426950408889634 / eps_next);
  rc.eps = (float)eps;
  return rc;
}

// Loss / gradient factors of one row once both final potentials are known.
//   S = potential against the row's own cloud (a_x or b_y), C = against the other cloud (b_x or a_y).
//   unbalanced:  term = (rho + eps/2) * (exp(-S/rho) - exp(-C/rho))      balanced: term = C - S
// exp(-S/rho) - exp(-C/rho) is evaluated as -exp(-S/rho) * expm1((S - C)/rho) to avoid cancellation.
struct RowFinal {
  float term;  // d loss / d weight  (loss contribution = weight * term)
  float eS;    // exp(-S/rho)  (1 when balanced)
  float eC;    // exp(-C/rho)
};
__device__ __forceinline__ RowFinal row_final(float S, float C, double rho, float eps) {
  RowFinal f;
  if (rho < 0.0) {
    f.term = C - S;
    f.eS = 1.f;
    f.eC = 1.f;
  } else {
    const float r = (float)rho, k = (float)(rho + 0.5 * (double)eps);
    f.eS = expf(-S / r);
    f.eC = expf(-C / r);
    f.term = -k * f.eS * expm1f((S - C) / r);
  }
  return f;
}

}  // namespace kdot
