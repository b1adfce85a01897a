// Shared device helpers and parameter blocks for the kdot kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kdot.h"

namespace kdot {

constexpr int KDOT_SCHED_TABLE = 48;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegBig = -1.0e30f;  // "minus infinity" that survives (a - b) without NaN
constexpr float kLogZeroWeight = -100000.0f;  // geomloss log_weights(): log(0) stand-in

// Per-round constants of the eps-scaling loop, derived in float64; the fp32 copies are rounded once.
//   v_ij (log2 domain) = h_j + coef * |x_i - y_j|^2          coef  = -0.5*log2(e)/eps      (p = 2)
//                      = h_j + coef * |x_i - y_j|            coef  = -log2(e)/eps          (p = 1)
//   new potential      = scale * log2-sum-exp                 scale = -lambda(eps)*eps*ln 2
//   h_j for NEXT round = lw2_j + pot_j * hmul                 hmul  = log2(e)/eps_next
// Precision plan: potentials, h and the log-sum-exp of EVERY round are carried in float64
// (per-row work, a few DP operations per row and round); the per-pair soft-min arguments are evaluated in fp32 except
// in the last KDOT_HI_ROUNDS rounds, where |h| and |coef * d^2| reach 1e3..1e4 log2-units and an fp32 argument
// (ulp ~1e-3) would leave ~7e-4 relative noise on the soft-max weights of the gradient: there the argument
// h_j + coef * d^2 - ref is formed in float64 (DADD/DFMA at 64 lanes/clk/SM on B200) and only the small difference is
// rounded to fp32 for the one ex2 per pair.
struct RoundConst {
  float coef;
  float scale;
  float hmul;
  float eps;
  double coefd;
  double scaled;
  double hmuld;
};

#ifndef KDOT_HI_ROUNDS
#define KDOT_HI_ROUNDS 6
#endif
// Rounds [nrounds - KDOT_HI_ROUNDS, nrounds) (the last one is the gradient round) evaluate their pair arguments in
// float64 -- but only while they are COLD (eps < eps_0 / 256): in a warm round |coef * d^2| <= 0.72 * eps_0 / eps < 185
// log2-units, where an fp32 argument is accurate to ~1e-5 anyway.  Errors made in earlier rounds are halved by every
// averaged update that follows, so six trailing rounds leave less than 2^-6 of an fp32 potential's rounding error.
#ifndef KDOT_HI_EPS_RATIO
#define KDOT_HI_EPS_RATIO 256.0f
#endif
__device__ __forceinline__ bool is_hi_round(int r, int nrounds, float eps, float eps0, int last = KDOT_HI_ROUNDS) {
  return r >= nrounds - last && eps * KDOT_HI_EPS_RATIO < eps0;
}

// Magnitude test that goes with is_hi_round: an fp32 soft-min argument carries an error of ~2^-24 max(|h_j|, |coef d^2|),
// and an error made k rounds before the last one reaches the final h amplified by about 2^(k-1) (eps shrinks by
// scaling^-p per round, the averaged update halves; fitted on the fp64 oracle).  Pairs whose
// offsets stay below  kHiMagnitude * 2^(k+1)  are accurate enough in fp32:  hi iff  max|h| * hi_mag_factor(k) > 1.
#ifndef KDOT_HI_MAGNITUDE
#define KDOT_HI_MAGNITUDE 64.0f
#endif
constexpr float kHiMagnitude = KDOT_HI_MAGNITUDE;
__device__ __forceinline__ float hi_mag_factor(int r, int nrounds) {
  return exp2f(-(float)(nrounds - r)) / kHiMagnitude;  // k = nrounds - 1 - r rounds follow this one
}

// float64 -> fp32 of a soft-min argument difference without the XU pipe (F2F.F32.F64 shares it with MUFU.EX2 at 16
// lanes/clk/SM): a 64-bit integer add of half an fp32 ulp (round to nearest, ties away) and three integer-pipe
// operations that re-pack sign, exponent and the top 23 mantissa bits.  Valid for 2^-126 <= |u| < 2^127; smaller |u|
// (in particular an exact 0) returns 0.
__device__ __forceinline__ float f64_to_f32_trunc(double u) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(u) + 0x10000000ull;  // + 2^28: half of the dropped 29 bits
  const unsigned int hi = (unsigned int)(b >> 32), lo = (unsigned int)b;
  const unsigned int packed = __funnelshift_l(lo, hi, 3) ^ 0x40000000u;          // exponent re-biased: 11 -> 8 bits
  const unsigned int bits = (packed & 0x7fffffffu) | (hi & 0x80000000u);
  return ((hi & 0x7ff00000u) < (897u << 20)) ? 0.f : __uint_as_float(bits);
}
// truncating variant without the small-|u| guard, for callers that keep u away from 0 by construction (u <= -1)
__device__ __forceinline__ float f64_to_f32_trunc_nz(double u) {
  const unsigned int hi = (unsigned int)__double2hiint(u), lo = (unsigned int)__double2loint(u);
  const unsigned int packed = __funnelshift_l(lo, hi, 3) ^ 0x40000000u;
  return __uint_as_float((packed & 0x7fffffffu) | (hi & 0x80000000u));
}

// data-independent schedule inputs, precomputed in float64 on the host
struct SchedParams {
  double p;              // 1 or 2
  double log_blur_p;     // p * ln(blur)
  double log_scaling_p;  // p * ln(scaling)
  double eps_final;      // blur^p
  double rho;            // reach^p, < 0 -> balanced
  double pow_table[KDOT_SCHED_TABLE];  // scaling^(p k) = exp(k * p ln scaling), k = 0..KDOT_SCHED_TABLE-1
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
  SchedParams sp;
  double rho;  // reach^p, < 0 -> balanced   (copy of sp.rho)
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
  int32_t* order;         // [nimg] tiled path: images by descending size (rank -> image), written by the prep kernel
  long long* dbg_clk;     // optional [nimg][16] SM-clock stamps / counters (kdot_debug_set_clock_buffer), NULL in production
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void dbg_stamp(const SinkhornParams& p, int img, int k) {
  if (p.dbg_clk && threadIdx.x == 0) p.dbg_clk[(size_t)img * 16 + k] = clock64();
}

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

// log2 on the SFU.  Used only for log2(sum of exps) with the sum in [1, #columns]: the 2^-22 relative error of
// lg2.approx is far below the fp32 rounding of the running max it is added to (|max| ~ 1e2..1e4).
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// log2 of a running exp-sum in float64.  lg2.approx is accurate to 2^-22 ABSOLUTE only near 1; for a sum kept against
// a stale reference exponent (up to 2^24 in the tiled kernel, 2^80 in the streaming kernel) its 2^-22 RELATIVE error
// would put up to 2e-5 log2-units into the potential -- times eps_r / eps_next = 4 in the next round's h.  Splitting
// off the binary exponent first keeps the SFU argument in [1, 2).
__device__ __forceinline__ double lg2_sum(float s) {
  const int bits = __float_as_int(s);
  const int e = (bits >> 23) - 127;
  const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
  return (double)e + (double)lg2_approx(m);
}

// Same to float64 accuracy, for the kernels whose rows hold many pairs (per-row cost: ~15 DP operations): one Newton
// step on the SFU estimate y0, log2(m) = y0 + log2(m * 2^-y0), with 2^-y0 from a degree-10 float64 polynomial.  In the
// dense configurations an lse error of 2^-22 per row and round (lg2.approx) alone costs 2e-5 on d/dx.
__device__ __forceinline__ double lg2_sum_exact(float s) {
  const int bits = __float_as_int(s);
  const int e = (bits >> 23) - 127;
  const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
  const float y0 = lg2_approx(m);
  const double t = (0.5 - (double)y0) * 0.6931471805599453;  // |t| <= 0.3466
  double q = 2.7557319223985893e-07;                         // 1/10!
  q = fma(q, t, 2.7557319223985888e-06);
  q = fma(q, t, 2.4801587301587302e-05);
  q = fma(q, t, 1.9841269841269841e-04);
  q = fma(q, t, 1.3888888888888889e-03);
  q = fma(q, t, 8.3333333333333332e-03);
  q = fma(q, t, 4.1666666666666664e-02);
  q = fma(q, t, 1.6666666666666666e-01);
  q = fma(q, t, 0.5);
  q = fma(q, t, 1.0);
  q = fma(q, t, 1.0);                                        // e^t = 2^(0.5 - y0)
  const double r = fma((double)m * 0.7071067811865476, q, -1.0);  // m * 2^-y0 - 1  (~1e-7)
  return ((double)e + (double)y0) + r * 1.4426950408889634 * (1.0 - 0.5 * r);
}

// geomloss max_diameter on fp32 data: |maxs - mins|_2 evaluated without FMA contraction.
__device__ __forceinline__ float bbox_diameter(float minx, float miny, float maxx, float maxy) {
  const float ex = __fsub_rn(maxx, minx), ey = __fsub_rn(maxy, miny);
  return sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
}

// ---- geomloss' epsilon schedule -------------------------------------------------------------------------
//   eps_s = [diam^p] + [exp(e) for e in arange(p ln diam, p ln blur, p ln scaling)] + [blur^p]     (float64)
// i.e. eps_s[1 + k] = diam^p * scaling^(p k) for every k >= 0 with diam^p * scaling^(p k) > blur^p.
// The data-independent factors scaling^(p k) (k < KDOT_SCHED_TABLE), p ln blur, p ln scaling and blur^p are
// precomputed in float64 on the host (SchedParams), so the common case needs no fp64 transcendental on the
// device: a few DMUL/DSETP per round.  Longer schedules (scaling close to 1) take the log/exp route below.
// Both routes agree with numpy's arange/exp to a few ulps of float64 (1e-15 relative on eps); the schedule
// LENGTH can differ only if diam^p * scaling^(p k) hits blur^p to within that error.
struct ImgSched {
  int nits;       // len(eps_s)
  int slow;       // 1: schedule longer than the host table, use start/delta with exp()
  double start;   // p ln diam                                   (slow route only)
  double delta;   // numpy arange step (start + step) - start     (slow route only)
  double eps0;    // diam^p
};

static __device__ __noinline__ ImgSched image_schedule_slow(double diam, double p, double log_blur_p,
                                                            double log_scaling_p, double eps0) {
  ImgSched is;
  const double a = p * log(diam);
  double len = ceil((log_blur_p - a) / log_scaling_p);
  if (!(len > 0.0)) len = 0.0;
  is.nits = (int)len + 2;
  is.slow = 1;
  is.start = a;
  is.delta = (a + log_scaling_p) - a;
  is.eps0 = eps0;
  return is;
}

__device__ __forceinline__ ImgSched image_schedule(float diam_f, const SchedParams& sp) {
  const double diam = (double)diam_f;
  const double eps0 = sp.p == 2.0 ? diam * diam : diam;  // p in {1, 2}
  if (eps0 * sp.pow_table[KDOT_SCHED_TABLE - 1] > sp.eps_final) return image_schedule_slow(diam, sp.p, sp.log_blur_p, sp.log_scaling_p, eps0);
  ImgSched is;
  int len = 0;
  while (len < KDOT_SCHED_TABLE && eps0 * sp.pow_table[len] > sp.eps_final) ++len;
  // Exact ties -- diam^p * scaling^(p k) == blur^p, e.g. scaling = 0.5 with diam / blur a power of two -- are decided
  // by rounding: numpy's arange length is ceil((p ln blur - p ln diam) / (p ln scaling)), the loop above compares
  // products of exponentials.  Within 1e-9 of a tie the length is taken from numpy's own formula (log route); nits is an
  // integer output and has to be bit-exact.
  {
    const double last_in = len > 0 ? eps0 * sp.pow_table[len - 1] / sp.eps_final : 2.0;
    const double first_out = len < KDOT_SCHED_TABLE ? eps0 * sp.pow_table[len] / sp.eps_final : 0.0;
    if (last_in - 1.0 < 1e-9 || 1.0 - first_out < 1e-9)
      return image_schedule_slow(diam, sp.p, sp.log_blur_p, sp.log_scaling_p, eps0);
  }
  is.nits = len + 2;
  is.slow = 0;
  is.start = 0.0;
  is.delta = 0.0;
  is.eps0 = eps0;
  return is;
}

// image_schedule evaluated by a whole warp (every lane gets the result): the table is scanned by two ballots instead of a
// serial loop of dependent constant-bank loads (the predicate is monotonic in k, so the length is its population count).
__device__ __forceinline__ ImgSched image_schedule_warp(float diam_f, const SchedParams& sp, int lane) {
  static_assert(KDOT_SCHED_TABLE <= 64, "two ballots cover the table");
  const double diam = (double)diam_f;
  const double eps0 = sp.p == 2.0 ? diam * diam : diam;  // p in {1, 2}
  const bool in0 = lane < KDOT_SCHED_TABLE && eps0 * sp.pow_table[lane < KDOT_SCHED_TABLE ? lane : 0] > sp.eps_final;
  const bool in1 = lane + 32 < KDOT_SCHED_TABLE && eps0 * sp.pow_table[lane + 32 < KDOT_SCHED_TABLE ? lane + 32 : 0] > sp.eps_final;
  const int len = __popc(__ballot_sync(0xffffffffu, in0)) + __popc(__ballot_sync(0xffffffffu, in1));
  if (len >= KDOT_SCHED_TABLE) return image_schedule_slow(diam, sp.p, sp.log_blur_p, sp.log_scaling_p, eps0);
  // exact ties (see image_schedule): the same 1e-9 band around eps_final, written without divisions
  const double last_in = len > 0 ? eps0 * sp.pow_table[len - 1] : 2.0 * sp.eps_final;
  const double first_out = eps0 * sp.pow_table[len];
  if (last_in < sp.eps_final * (1.0 + 1e-9) || first_out > sp.eps_final * (1.0 - 1e-9))
    return image_schedule_slow(diam, sp.p, sp.log_blur_p, sp.log_scaling_p, eps0);
  ImgSched is;
  is.nits = len + 2;
  is.slow = 0;
  is.start = 0.0;
  is.delta = 0.0;
  is.eps0 = eps0;
  return is;
}

// Round r of the kernel's flattened loop: r = 0 init (eps_s[0]), 1..nits the loop, nits+1 the last extrapolation.
__device__ __forceinline__ int round_to_sched(int r, int nits) {
  if (r == 0) return 0;
  if (r <= nits) return r - 1;
  return nits - 1;
}

static __device__ __noinline__ double schedule_eps_slow(int k, double start, double delta) {
  return exp(start + (double)k * delta);
}

__device__ __forceinline__ double schedule_eps(int t, const ImgSched& is, const SchedParams& sp) {
  if (t <= 0) return is.eps0;
  if (t >= is.nits - 1) return sp.eps_final;
  if (!is.slow) return is.eps0 * sp.pow_table[t - 1];
  return schedule_eps_slow(t - 1, is.start, is.delta);
}

__device__ __forceinline__ RoundConst make_round_const(int r, const ImgSched& is, const SchedParams& sp) {
  const double eps = schedule_eps(round_to_sched(r, is.nits), is, sp);
  const double eps_next = schedule_eps(round_to_sched(r + 1, is.nits), is, sp);
  RoundConst rc;
  rc.coefd = (sp.p == 2.0 ? -0.5 : -1.0) * 1.4426950408889634 / eps;  // cost |d|^2/2 (p = 2) or |d| (p = 1)
  // -lambda eps ln 2 with lambda = 1 / (1 + eps/rho) = rho / (rho + eps): one division
  rc.scaled = sp.rho < 0.0 ? -eps * 0.6931471805599453 : -(sp.rho * eps * 0.6931471805599453) / (sp.rho + eps);
  rc.hmuld = 1.4426950408889634 / eps_next;
  rc.coef = (float)rc.coefd;
  rc.scale = (float)rc.scaled;
  rc.hmul = (float)rc.hmuld;
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
// same with the final potentials in float64: the difference S - C is formed before rounding
__device__ __forceinline__ RowFinal row_final(double S, double C, double rho, float eps) {
  RowFinal f;
  if (rho < 0.0) {
    f.term = (float)(C - S);
    f.eS = 1.f;
    f.eC = 1.f;
  } else {
    const float r = (float)rho, k = (float)(rho + 0.5 * (double)eps);
    f.eS = expf(-(float)S / r);
    f.eC = expf(-(float)C / r);
    f.term = -k * f.eS * expm1f((float)(S - C) / r);
  }
  return f;
}

}  // namespace kdot
