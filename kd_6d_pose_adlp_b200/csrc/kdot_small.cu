// Fused OT distillation loss, small-problem kernel (N_i + M_i <= 32 cells per image, D = 2: the shipped ape shape).
//
// One CTA per image, one warp per OT slot (keypoint); warps are independent until the final sum over
// slots.  Everything the reference does for one image -- losses/loss_libs.py:8-12 (normalise), :22-50
// (per-image split / transposes) and geomloss' tensorized Sinkhorn divergence with its backward -- runs
// in this single launch for the whole mini-batch; the N x M cost matrices live only in registers.
//
// Every point i carries two potentials: S_i against its own cloud (geomloss a_x / b_y) and C_i against
// the other cloud (b_x / a_y).  With h^S_j = log w_j + S_j/eps and h^C_j = log w_j + C_j/eps the four
// softmins of a Sinkhorn round collapse to one rule for every row i:
//   S_i <- lambda * softmin_{j in own cloud}(h^S_j),   C_i <- lambda * softmin_{j in other cloud}(h^C_j).
//
// Fast path (N_i + M_i <= 32, the shipped ape shape): lane = row; the columns sit in shared memory as
// SoA float4 chunks [student cols padded to 4 | teacher cols padded to 4]; the round body is fully
// unrolled for the image's chunk count (template CH), evaluates 2 columns per instruction with packed
// f32x2 math, keeps all soft-min arguments in registers (exact max, one exp2 per pair) and needs one
// __syncwarp per round.  The eps schedule is computed in float64 with one lane per round.
#include <cooperative_groups.h>

#include "kdot_common.cuh"

namespace cg = cooperative_groups;

namespace kdot {

#ifndef KDOT_SMALL_HI_ROUNDS
#define KDOT_SMALL_HI_ROUNDS KDOT_HI_ROUNDS
#endif
constexpr int kSmallHiRounds = KDOT_SMALL_HI_ROUNDS;  // trailing rounds that may take float64 pair arguments (kdot_common.cuh)
constexpr int kFastMaxCols = 40;  // padded columns of the fast path (<= 32 points + padding)
constexpr int kFastMaxCH = kFastMaxCols / 4;
constexpr int kFastWarpDoubles = (4 + 32 + 3) * kFastMaxCols + 4;  // per-warp shared memory in units of 8 bytes (see the kernel)

// =========================================================================================================
// fast path
// =========================================================================================================
template <int CH, bool GRAD>
struct RoundOut {
  double lseX, lseY;       // log2-sum-exp over the student / teacher columns: (double)max + (double)log2(sum)
  float gXx, gXy, sX;      // sum e * (p_j - p_i) over the student columns and sum e   (GRAD only)
  float gYx, gYy, sY;
};

template <int CH, bool GRAD>
__device__ __forceinline__ RoundOut<CH, GRAD> fast_round(const float* __restrict__ cx, const float* __restrict__ cy,
                                                         const float* __restrict__ hp, int nchx, float px, float py,
                                                         float coef) {
  const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py), coef2 = make_float2(coef, coef);
  float2 v[2 * CH];
  float mX = kNegBig, mY = kNegBig;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const float4 X = reinterpret_cast<const float4*>(cx)[c];
    const float4 Y = reinterpret_cast<const float4*>(cy)[c];
    const float4 H = reinterpret_cast<const float4*>(hp)[c];
    const float2 d0 = __fadd2_rn(make_float2(X.x, X.y), npx), d1 = __fadd2_rn(make_float2(X.z, X.w), npx);
    const float2 e0 = __fadd2_rn(make_float2(Y.x, Y.y), npy), e1 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
    v[2 * c] = __ffma2_rn(__ffma2_rn(e0, e0, __fmul2_rn(d0, d0)), coef2, make_float2(H.x, H.y));
    v[2 * c + 1] = __ffma2_rn(__ffma2_rn(e1, e1, __fmul2_rn(d1, d1)), coef2, make_float2(H.z, H.w));
    const float m4 = fmaxf(fmaxf(v[2 * c].x, v[2 * c].y), fmaxf(v[2 * c + 1].x, v[2 * c + 1].y));
    if (c < nchx) mX = fmaxf(mX, m4); else mY = fmaxf(mY, m4);  // warp-uniform
  }
  float2 sX = make_float2(0.f, 0.f), sY = sX, gXx = sX, gXy = sX, gYx = sX, gYy = sX;
  const float2 nmX = make_float2(-mX, -mX), nmY = make_float2(-mY, -mY);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const bool isx = c < nchx;  // warp-uniform
    const float2 nm = isx ? nmX : nmY;
    const float2 a0 = __fadd2_rn(v[2 * c], nm), a1 = __fadd2_rn(v[2 * c + 1], nm);
    const float2 p0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
    const float2 p1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
    const float2 ps = __fadd2_rn(p0, p1);
    if (isx) sX = __fadd2_rn(sX, ps); else sY = __fadd2_rn(sY, ps);
    if (GRAD) {
      const float4 X = reinterpret_cast<const float4*>(cx)[c];
      const float4 Y = reinterpret_cast<const float4*>(cy)[c];
      const float2 d0 = __fadd2_rn(make_float2(X.x, X.y), npx), d1 = __fadd2_rn(make_float2(X.z, X.w), npx);
      const float2 e0 = __fadd2_rn(make_float2(Y.x, Y.y), npy), e1 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
      const float2 gx = __ffma2_rn(p1, d1, __fmul2_rn(p0, d0)), gy = __ffma2_rn(p1, e1, __fmul2_rn(p0, e0));
      if (isx) { gXx = __fadd2_rn(gXx, gx); gXy = __fadd2_rn(gXy, gy); }
      else     { gYx = __fadd2_rn(gYx, gx); gYy = __fadd2_rn(gYy, gy); }
    }
  }
  RoundOut<CH, GRAD> o;
  o.sX = sX.x + sX.y; o.sY = sY.x + sY.y;
  o.lseX = (double)mX + lg2_sum(o.sX);  // the sum is formed in float64: |max| reaches 1e4 in the cold rounds
  o.lseY = (double)mY + lg2_sum(o.sY);
  o.gXx = gXx.x + gXx.y; o.gXy = gXy.x + gXy.y;
  o.gYx = gYx.x + gYx.y; o.gYy = gYy.x + gYy.y;
  return o;
}

// ---- rolled forms (runtime chunk count) ---------------------------------------------------------------------------------
// ncu on the fully unrolled, per-chunk-count templated rounds showed `no_instruction` as the top stall (3 cycles per issued
// instruction, 13 % issue utilisation): one warp per sub-partition streams ~500 straight-line instructions per round out of
// the L1.5 instruction cache and never reuses them.  The rolled forms run the same arithmetic (bit-identical sums: same
// accumulators, same order) from loops of ~40 instructions that stay in the L0 instruction cache, at the price of
// recomputing the arguments in the second pass (packed FMAs, cheap next to the ex2) instead of holding them in registers.
__device__ __forceinline__ float rolled_max(const float* __restrict__ cx, const float* __restrict__ cy,
                                            const float* __restrict__ hp, int c0, int c1, float2 npx, float2 npy, float2 coef2) {
  float m = kNegBig;
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    const float4 X = reinterpret_cast<const float4*>(cx)[c];
    const float4 Y = reinterpret_cast<const float4*>(cy)[c];
    const float4 H = reinterpret_cast<const float4*>(hp)[c];
    const float2 d0 = __fadd2_rn(make_float2(X.x, X.y), npx), d1 = __fadd2_rn(make_float2(X.z, X.w), npx);
    const float2 e0 = __fadd2_rn(make_float2(Y.x, Y.y), npy), e1 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
    const float2 v0 = __ffma2_rn(__ffma2_rn(e0, e0, __fmul2_rn(d0, d0)), coef2, make_float2(H.x, H.y));
    const float2 v1 = __ffma2_rn(__ffma2_rn(e1, e1, __fmul2_rn(d1, d1)), coef2, make_float2(H.z, H.w));
    m = fmaxf(m, fmaxf(fmaxf(v0.x, v0.y), fmaxf(v1.x, v1.y)));
  }
  return m;
}

template <bool GRAD>
__device__ __forceinline__ void rolled_sum(const float* __restrict__ cx, const float* __restrict__ cy,
                                           const float* __restrict__ hp, int c0, int c1, float2 npx, float2 npy, float2 coef2,
                                           float m, float& ssum, float& gxo, float& gyo) {
  const float2 nm = make_float2(-m, -m);
  float2 s = make_float2(0.f, 0.f), gx = s, gy = s;
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    const float4 X = reinterpret_cast<const float4*>(cx)[c];
    const float4 Y = reinterpret_cast<const float4*>(cy)[c];
    const float4 H = reinterpret_cast<const float4*>(hp)[c];
    const float2 d0 = __fadd2_rn(make_float2(X.x, X.y), npx), d1 = __fadd2_rn(make_float2(X.z, X.w), npx);
    const float2 e0 = __fadd2_rn(make_float2(Y.x, Y.y), npy), e1 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
    const float2 a0 = __fadd2_rn(__ffma2_rn(__ffma2_rn(e0, e0, __fmul2_rn(d0, d0)), coef2, make_float2(H.x, H.y)), nm);
    const float2 a1 = __fadd2_rn(__ffma2_rn(__ffma2_rn(e1, e1, __fmul2_rn(d1, d1)), coef2, make_float2(H.z, H.w)), nm);
    const float2 p0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
    const float2 p1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
    s = __fadd2_rn(s, __fadd2_rn(p0, p1));
    if (GRAD) {
      gx = __fadd2_rn(gx, __ffma2_rn(p1, d1, __fmul2_rn(p0, d0)));
      gy = __fadd2_rn(gy, __ffma2_rn(p1, e1, __fmul2_rn(p0, e0)));
    }
  }
  ssum = s.x + s.y; gxo = gx.x + gx.y; gyo = gy.x + gy.y;
}

struct RoundOutRt {
  double lseX, lseY;
  float gXx, gXy, sX, gYx, gYy, sY;
};

template <bool GRAD>
__device__ __forceinline__ RoundOutRt fast_round_rt(const float* __restrict__ cx, const float* __restrict__ cy,
                                                    const float* __restrict__ hp, int nchx, int nch, float px, float py,
                                                    float coef) {
  const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py), coef2 = make_float2(coef, coef);
  const float mX = rolled_max(cx, cy, hp, 0, nchx, npx, npy, coef2);
  const float mY = rolled_max(cx, cy, hp, nchx, nch, npx, npy, coef2);
  RoundOutRt o;
  rolled_sum<GRAD>(cx, cy, hp, 0, nchx, npx, npy, coef2, mX, o.sX, o.gXx, o.gXy);
  rolled_sum<GRAD>(cx, cy, hp, nchx, nch, npx, npy, coef2, mY, o.sY, o.gYx, o.gYy);
  o.lseX = (double)mX + lg2_sum(o.sX);
  o.lseY = (double)mY + lg2_sum(o.sY);
  return o;
}

// rolled high-precision sweep over chunks [c0, c1) relative to ref (see hi_round)
template <bool GRAD>
__device__ __forceinline__ void hi_sum_rt(const double* __restrict__ d2s, const double* __restrict__ hd,
                                          const float* __restrict__ cx, const float* __restrict__ cy, int c0, int c1,
                                          float pxf, float pyf, double coef, double ref, float& ssum, float& gxo, float& gyo) {
  float s = 0.f, gx = 0.f, gy = 0.f;
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    const int j = 4 * c;
    const double2 H0 = *reinterpret_cast<const double2*>(hd + j), H1 = *reinterpret_cast<const double2*>(hd + j + 2);
    const double u0 = fma(coef, d2s[(j + 0) * 32], H0.x - ref), u1 = fma(coef, d2s[(j + 1) * 32], H0.y - ref);
    const double u2 = fma(coef, d2s[(j + 2) * 32], H1.x - ref), u3 = fma(coef, d2s[(j + 3) * 32], H1.y - ref);
    const float e0 = ex2_approx(f64_to_f32_trunc_nz(u0)), e1 = ex2_approx(f64_to_f32_trunc_nz(u1));
    const float e2 = ex2_approx(f64_to_f32_trunc_nz(u2)), e3 = ex2_approx(f64_to_f32_trunc_nz(u3));
    s += (e0 + e1) + (e2 + e3);
    if (GRAD) {
      const float4 XF = *reinterpret_cast<const float4*>(cx + j);
      const float4 YF = *reinterpret_cast<const float4*>(cy + j);
      gx += fmaf(e0, XF.x - pxf, fmaf(e1, XF.y - pxf, fmaf(e2, XF.z - pxf, e3 * (XF.w - pxf))));
      gy += fmaf(e0, YF.x - pyf, fmaf(e1, YF.y - pyf, fmaf(e2, YF.z - pyf, e3 * (YF.w - pyf))));
    }
  }
  ssum = s; gxo = gx; gyo = gy;
}

// High-precision round (the cold ones among the last KDOT_HI_ROUNDS rounds, see kdot_common.cuh): the soft-min argument
//   t_ij = h_j + coef * |p_i - p_j|^2   (|h|, |coef d^2| ~ 1e3..1e4, t - max_j t = O(1) for the pairs that matter)
// is formed in float64 from float64 copies of the columns and offsets; only t - ref is re-packed to fp32 for the
// exponential.  The reference is the fp32 estimate of the row maximum plus 2 (fast_max: one packed fp32 pass, error
// ~1e-3), so t - ref <= -1 for every pair: a single float64 sweep, and the difference is never 0 (f64_to_f32_trunc).
template <int CH>
__device__ __forceinline__ void fast_max(const float* __restrict__ cx, const float* __restrict__ cy,
                                         const float* __restrict__ hp, int nchx, float px, float py, float coef,
                                         float& mX, float& mY) {
  const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py), coef2 = make_float2(coef, coef);
  mX = kNegBig; mY = kNegBig;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const float4 X = reinterpret_cast<const float4*>(cx)[c];
    const float4 Y = reinterpret_cast<const float4*>(cy)[c];
    const float4 H = reinterpret_cast<const float4*>(hp)[c];
    const float2 d0 = __fadd2_rn(make_float2(X.x, X.y), npx), d1 = __fadd2_rn(make_float2(X.z, X.w), npx);
    const float2 e0 = __fadd2_rn(make_float2(Y.x, Y.y), npy), e1 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
    const float2 v0 = __ffma2_rn(__ffma2_rn(e0, e0, __fmul2_rn(d0, d0)), coef2, make_float2(H.x, H.y));
    const float2 v1 = __ffma2_rn(__ffma2_rn(e1, e1, __fmul2_rn(d1, d1)), coef2, make_float2(H.z, H.w));
    const float m4 = fmaxf(fmaxf(v0.x, v0.y), fmaxf(v1.x, v1.y));
    if (c < nchx) mX = fmaxf(mX, m4); else mY = fmaxf(mY, m4);  // warp-uniform
  }
}

// Offsets beyond ~1e6 log2-units (un-normalised pixel coordinates at blur = 1e-3): the fp32 estimate of the row maximum
// is no longer good to +-2, the maxima are taken in float64.  Rare and slow (fmax on doubles): kept out of line so that it
// does not weigh on the register allocation and scheduling of the rounds.
static __device__ __noinline__ void hi_reference_f64(const double* d2s, const double* hd, int ncol, int ncolx, double coef,
                                                     double& refX, double& refY) {
  double mx = -1.0e300, my = -1.0e300;
  for (int j = 0; j < ncol; ++j) {
    const double t = fma(coef, d2s[j * 32], hd[j]);
    if (j < ncolx) mx = fmax(mx, t); else my = fmax(my, t);
  }
  refX = mx + 2.0; refY = my + 2.0;
}

// d2s: this lane's squared distances to every column in float64, [column][lane] in shared memory (computed once per
// (image, slot): they do not change between rounds), so a pair costs one LDS.64, one DADD and one DFMA on the DP pipe,
// three integer operations for the re-packing and the one ex2.
template <int CH, bool GRAD>
__device__ __forceinline__ RoundOut<CH, GRAD> hi_round(const double* __restrict__ d2s, const double* __restrict__ hd,
                                                       const float* __restrict__ cx, const float* __restrict__ cy,
                                                       int nchx, float pxf, float pyf, double coef, float mX, float mY,
                                                       float hmag) {
  double refX = (double)mX + 2.0, refY = (double)mY + 2.0;
  if ((hmag + fabsf(mX) + fabsf(mY)) * 9.5e-7f > 1.0f) hi_reference_f64(d2s, hd, 4 * CH, 4 * nchx, coef, refX, refY);
  float sX = 0.f, sY = 0.f, gXx = 0.f, gXy = 0.f, gYx = 0.f, gYy = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int j = 4 * c;
    const bool isx = c < nchx;  // warp-uniform
    const double ref = isx ? refX : refY;
    const double2 H0 = *reinterpret_cast<const double2*>(hd + j), H1 = *reinterpret_cast<const double2*>(hd + j + 2);
    const double u0 = fma(coef, d2s[(j + 0) * 32], H0.x - ref), u1 = fma(coef, d2s[(j + 1) * 32], H0.y - ref);
    const double u2 = fma(coef, d2s[(j + 2) * 32], H1.x - ref), u3 = fma(coef, d2s[(j + 3) * 32], H1.y - ref);
    const float e0 = ex2_approx(f64_to_f32_trunc_nz(u0)), e1 = ex2_approx(f64_to_f32_trunc_nz(u1));
    const float e2 = ex2_approx(f64_to_f32_trunc_nz(u2)), e3 = ex2_approx(f64_to_f32_trunc_nz(u3));
    const float es = (e0 + e1) + (e2 + e3);
    if (isx) sX += es; else sY += es;
    if (GRAD) {
      const float4 XF = *reinterpret_cast<const float4*>(cx + j);
      const float4 YF = *reinterpret_cast<const float4*>(cy + j);
      const float gx = fmaf(e0, XF.x - pxf, fmaf(e1, XF.y - pxf, fmaf(e2, XF.z - pxf, e3 * (XF.w - pxf))));
      const float gy = fmaf(e0, YF.x - pyf, fmaf(e1, YF.y - pyf, fmaf(e2, YF.z - pyf, e3 * (YF.w - pyf))));
      if (isx) { gXx += gx; gXy += gy; } else { gYx += gx; gYy += gy; }
    }
  }
  RoundOut<CH, GRAD> o;
  o.sX = sX; o.sY = sY;
  o.lseX = refX + lg2_sum(sX);
  o.lseY = refY + lg2_sum(sY);
  o.gXx = gXx; o.gXy = gXy; o.gYx = gYx; o.gYy = gYy;
  return o;
}

struct FastCtx {
  float* cx; float* cy; float* hb;  // per-warp smem: cx[40], cy[40], hb[2 buffers][2 views][40]
  double* hbd;   // float64 copy of hb (high-precision rounds)
  double* d2s;   // [40 columns][32 lanes] float64 squared distances of this lane's point to every column
  double* ctr;   // centres of the published offsets: (potS, potC) of the first student point | of the first teacher point
  int nchx;                         // student chunks (padded student columns / 4)
  int nstu;                         // student points N: lanes [0, N) are student rows, [N, N + M) teacher rows
  bool act, isx;
  int col;                          // padded column slot of this lane's point
  float px, py, wgt, lw2;
  double lw2d;
};

// All rounds of one (image, slot) for a compile-time chunk count.  Returns via references the final
// potentials' outputs.  `mine` holds the constants of round (lane) -- broadcast with shuffles.
// Potentials live in float64 registers for every round; h is published as an fp32 and a float64 copy.
template <int CH>
__device__ __forceinline__ void fast_solve(const FastCtx& c, int nrounds, const ImgSched& is, const SinkhornParams& prm,
                                           RoundConst mine, double& S_out, double& C_out,
                                           float& gSx, float& gSy, float& gCx, float& gCy, RoundConst& rc_last) {
  const int lane = threadIdx.x & 31;
  const float eps0 = (float)is.eps0;
  double potS = 0.0, potC = 0.0;
  float hmag = 0.f;  // max |h| over this slot's columns, as published for the current round (hi_mag_factor test)
  double addX = 0.0, addY = 0.0;  // centres of the consumed student / teacher column offsets (0 in the init round)
  int cur = 0;
  for (int r = 0; r < nrounds - 1; ++r) {
    if (r >= 32 && (r & 31) == 0)  // schedules longer than 32 rounds: next block of constants
      mine = make_round_const(r + lane, is, prm.sp);
#ifdef KDOT_SMALL_ROUND_STAMPS
    {  // profiling aid (tools/microbench_small.py): SM clock at the start of 7 rounds, slots 8..14 of the debug buffer
#ifdef KDOT_SMALL_ROUND_STAMPS_TAIL
      const int k = r - (nrounds - 8);   // the last 7 rounds before the gradient round; slot 15 = start of the gradient round
#else
      const int k = r;
#endif
      if (prm.dbg_clk && threadIdx.x == 0 && k >= 0 && k < 7) prm.dbg_clk[(size_t)(blockIdx.x / (gridDim.x / prm.nimg)) * 16 + 8 + k] = clock64();
    }
#endif
    const double scaled = __shfl_sync(0xffffffffu, mine.scaled, r & 31);
    const double hmuld = __shfl_sync(0xffffffffu, mine.hmuld, r & 31);
    double lseX, lseY;
    const float coef = __shfl_sync(0xffffffffu, mine.coef, r & 31);
    const float eps = __shfl_sync(0xffffffffu, mine.eps, r & 31);
    // centres for the h this round publishes (potentials as they stand after round r - 1; 0 before the first update)
#ifndef KDOT_SMALL_NO_CENTRE
    // (potS, potC) of the first student point and of the first teacher point, left in shared memory by their lanes at
    // the end of the previous round: two broadcast LDS.128 instead of eight shuffles (330 cycles per round, measured)
    const double2 own = *reinterpret_cast<const double2*>(c.ctr + (c.isx ? 0 : 2));   // this lane's own set
    const double2 oth = *reinterpret_cast<const double2*>(c.ctr + (c.isx ? 2 : 0));
    __syncwarp();   // every lane holds the centres before lanes 0 / N overwrite them at the end of this round
    const double ownS = own.x, ownC = own.y;
    const double refX = c.isx ? own.x : oth.y;   // student columns: student rows read h^S[X], teacher rows h^C[X]
    const double refY = c.isx ? oth.y : own.x;   // teacher columns: student rows read h^C[Y], teacher rows h^S[Y]
#else
    const double ownS = 0.0, ownC = 0.0, refX = 0.0, refY = 0.0;
#endif
    const float* hp = c.hb + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
    if (!(is_hi_round(r, nrounds, eps, eps0, kSmallHiRounds) && hmag * hi_mag_factor(r, nrounds) > 1.0f)) {
      const RoundOut<CH, false> o = fast_round<CH, false>(c.cx, c.cy, hp, c.nchx, c.px, c.py, coef);
      lseX = o.lseX; lseY = o.lseY;
    } else {
      const double coefd = __shfl_sync(0xffffffffu, mine.coefd, r & 31);
      const double* hpd = c.hbd + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
      float mX, mY;
      fast_max<CH>(c.cx, c.cy, hp, c.nchx, c.px, c.py, coef, mX, mY);
      const RoundOut<CH, false> o = hi_round<CH, false>(c.d2s, hpd, c.cx, c.cy, c.nchx, c.px, c.py, coefd, mX, mY, hmag);
      lseX = o.lseX; lseY = o.lseY;
    }
    lseX += addX; lseY += addY;   // the centres the consumed h was published relative to (see below)
    const double nS = scaled * (c.isx ? lseX : lseY);
    const double nC = scaled * (c.isx ? lseY : lseX);
    potS = r == 0 ? nS : 0.5 * (potS + nS);
    potC = r == 0 ? nC : 0.5 * (potC + nC);
    // Centred offsets: with unequal total masses every potential of a cloud carries a common term of order
    // rho * log(mass ratio) / eps (1e4 log2-units at eps = 1e-6) that cancels in h_j - max_j h_j; each of the four
    // (type S/C, cloud) sets is published relative to the potential its first point had ONE ROUND EARLIER (read at the top
    // of the round, off the critical path) and the same constant is added
    // back to the log-sum-exp in float64.  What is left in |h| is the variation across the cloud -- which is what decides
    // whether fp32 pair arguments are accurate enough (hi_mag_factor).
    addX = refX * hmuld;
    addY = refY * hmuld;
    hmag = 0.f;
    if (c.act) {
      const double hS = fma(potS - ownS, hmuld, c.lw2d), hC = fma(potC - ownC, hmuld, c.lw2d);
      hmag = fmaxf(fabsf((float)hS), fabsf((float)hC));
      float* hn = c.hb + (cur ^ 1) * (2 * kFastMaxCols);
      double* hnd = c.hbd + (cur ^ 1) * (2 * kFastMaxCols);
      hn[c.col] = (float)(c.isx ? hS : hC);                  // view 0: what student rows read for this column
      hn[kFastMaxCols + c.col] = (float)(c.isx ? hC : hS);   // view 1: what teacher rows read
      hnd[c.col] = c.isx ? hS : hC;
      hnd[kFastMaxCols + c.col] = c.isx ? hC : hS;
    }
    if (lane == 0 || lane == c.nstu)   // next round's centres (all lanes read the current ones at the top of this round)
      *reinterpret_cast<double2*>(c.ctr + (lane == 0 ? 0 : 2)) = make_double2(potS, potC);
    __syncwarp();
    hmag = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hmag)));  // non-negative floats order like their bits
    cur ^= 1;
  }
  const int r = nrounds - 1;
#ifdef KDOT_SMALL_ROUND_STAMPS
  if (prm.dbg_clk && threadIdx.x == 0) prm.dbg_clk[(size_t)(blockIdx.x / (gridDim.x / prm.nimg)) * 16 + 15] = clock64();
#endif
  if (r >= 32 && (r & 31) == 0)
    mine = make_round_const(r + lane, is, prm.sp);
  rc_last.coef = __shfl_sync(0xffffffffu, mine.coef, r & 31);
  rc_last.scale = __shfl_sync(0xffffffffu, mine.scale, r & 31);
  rc_last.hmul = 0.f;
  rc_last.eps = __shfl_sync(0xffffffffu, mine.eps, r & 31);
  rc_last.coefd = __shfl_sync(0xffffffffu, mine.coefd, r & 31);
  rc_last.scaled = __shfl_sync(0xffffffffu, mine.scaled, r & 31);
  rc_last.hmuld = 0.0;
  double lseX, lseY;
  float sX, sY, gXx, gXy, gYx, gYy;
  const float* hp = c.hb + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
  if (!(is_hi_round(r, nrounds, rc_last.eps, eps0, kSmallHiRounds) && hmag * hi_mag_factor(r, nrounds) > 1.0f)) {
    const RoundOut<CH, true> o = fast_round<CH, true>(c.cx, c.cy, hp, c.nchx, c.px, c.py, rc_last.coef);
    lseX = o.lseX; lseY = o.lseY; sX = o.sX; sY = o.sY;
    gXx = o.gXx; gXy = o.gXy; gYx = o.gYx; gYy = o.gYy;
  } else {
    const double* hpd = c.hbd + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
    float mX, mY;
    fast_max<CH>(c.cx, c.cy, hp, c.nchx, c.px, c.py, rc_last.coef, mX, mY);
    const RoundOut<CH, true> o = hi_round<CH, true>(c.d2s, hpd, c.cx, c.cy, c.nchx, c.px, c.py, rc_last.coefd, mX, mY, hmag);
    lseX = o.lseX; lseY = o.lseY; sX = o.sX; sY = o.sY;
    gXx = o.gXx; gXy = o.gXy; gYx = o.gYx; gYy = o.gYy;
  }
  lseX += addX; lseY += addY;
  S_out = rc_last.scaled * (c.isx ? lseX : lseY);
  C_out = rc_last.scaled * (c.isx ? lseY : lseX);
  // barycentric displacements  sum_j W_ij (p_j - p_i)  against own / other cloud (student rows use them)
  gSx = gXx / sX; gSy = gXy / sX;
  gCx = gYx / sY; gCy = gYy / sY;
}

// rolled variant of fast_solve (runtime chunk count, one instantiation): see the note at rolled_max
__device__ __forceinline__ void fast_solve_rt(const FastCtx& c, int nch, int nrounds, const ImgSched& is, const SinkhornParams& prm,
                                           RoundConst mine, double& S_out, double& C_out,
                                           float& gSx, float& gSy, float& gCx, float& gCy, RoundConst& rc_last) {
  const int lane = threadIdx.x & 31;
  const float eps0 = (float)is.eps0;
  double potS = 0.0, potC = 0.0;
  float hmag = 0.f;  // max |h| over this slot's columns, as published for the current round (hi_mag_factor test)
  double addX = 0.0, addY = 0.0;  // centres of the consumed student / teacher column offsets (0 in the init round)
  int cur = 0;
  for (int r = 0; r < nrounds - 1; ++r) {
    if (r >= 32 && (r & 31) == 0)  // schedules longer than 32 rounds: next block of constants
      mine = make_round_const(r + lane, is, prm.sp);
    const double scaled = __shfl_sync(0xffffffffu, mine.scaled, r & 31);
    const double hmuld = __shfl_sync(0xffffffffu, mine.hmuld, r & 31);
    double lseX, lseY;
    const float coef = __shfl_sync(0xffffffffu, mine.coef, r & 31);
    const float eps = __shfl_sync(0xffffffffu, mine.eps, r & 31);
    // centres for the h this round publishes (potentials as they stand after round r - 1; 0 before the first update)
#ifndef KDOT_SMALL_NO_CENTRE
    // (potS, potC) of the first student point and of the first teacher point, left in shared memory by their lanes at
    // the end of the previous round: two broadcast LDS.128 instead of eight shuffles
    const double2 own = *reinterpret_cast<const double2*>(c.ctr + (c.isx ? 0 : 2));   // this lane's own set
    const double2 oth = *reinterpret_cast<const double2*>(c.ctr + (c.isx ? 2 : 0));
    __syncwarp();   // every lane holds the centres before lanes 0 / N overwrite them at the end of this round
    const double ownS = own.x, ownC = own.y;
    const double refX = c.isx ? own.x : oth.y;   // student columns: student rows read h^S[X], teacher rows h^C[X]
    const double refY = c.isx ? oth.y : own.x;   // teacher columns: student rows read h^C[Y], teacher rows h^S[Y]
#else
    const double ownS = 0.0, ownC = 0.0, refX = 0.0, refY = 0.0;
#endif
    const float* hp = c.hb + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
    if (!(is_hi_round(r, nrounds, eps, eps0, kSmallHiRounds) && hmag * hi_mag_factor(r, nrounds) > 1.0f)) {
      const RoundOutRt o = fast_round_rt<false>(c.cx, c.cy, hp, c.nchx, nch, c.px, c.py, coef);
      lseX = o.lseX; lseY = o.lseY;
    } else {
      const double coefd = __shfl_sync(0xffffffffu, mine.coefd, r & 31);
      const double* hpd = c.hbd + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
      const float2 npx = make_float2(-c.px, -c.px), npy = make_float2(-c.py, -c.py), coef2 = make_float2(coef, coef);
      double refX = (double)rolled_max(c.cx, c.cy, hp, 0, c.nchx, npx, npy, coef2) + 2.0;
      double refY = (double)rolled_max(c.cx, c.cy, hp, c.nchx, nch, npx, npy, coef2) + 2.0;
      if ((hmag + (float)fabs(refX) + (float)fabs(refY)) * 9.5e-7f > 1.0f) hi_reference_f64(c.d2s, hpd, 4 * nch, 4 * c.nchx, coefd, refX, refY);
      float sx, sy, g0, g1;
      hi_sum_rt<false>(c.d2s, hpd, c.cx, c.cy, 0, c.nchx, c.px, c.py, coefd, refX, sx, g0, g1);
      hi_sum_rt<false>(c.d2s, hpd, c.cx, c.cy, c.nchx, nch, c.px, c.py, coefd, refY, sy, g0, g1);
      lseX = refX + lg2_sum(sx); lseY = refY + lg2_sum(sy);
    }
    lseX += addX; lseY += addY;   // the centres the consumed h was published relative to (see below)
    const double nS = scaled * (c.isx ? lseX : lseY);
    const double nC = scaled * (c.isx ? lseY : lseX);
    potS = r == 0 ? nS : 0.5 * (potS + nS);
    potC = r == 0 ? nC : 0.5 * (potC + nC);
    // Centred offsets: with unequal total masses every potential of a cloud carries a common term of order
    // rho * log(mass ratio) / eps (1e4 log2-units at eps = 1e-6) that cancels in h_j - max_j h_j; each of the four
    // (type S/C, cloud) sets is published relative to the potential its first point had ONE ROUND EARLIER (read at the top
    // of the round, off the critical path) and the same constant is added
    // back to the log-sum-exp in float64.  What is left in |h| is the variation across the cloud -- which is what decides
    // whether fp32 pair arguments are accurate enough (hi_mag_factor).
    addX = refX * hmuld;
    addY = refY * hmuld;
    hmag = 0.f;
    if (c.act) {
      const double hS = fma(potS - ownS, hmuld, c.lw2d), hC = fma(potC - ownC, hmuld, c.lw2d);
      hmag = fmaxf(fabsf((float)hS), fabsf((float)hC));
      float* hn = c.hb + (cur ^ 1) * (2 * kFastMaxCols);
      double* hnd = c.hbd + (cur ^ 1) * (2 * kFastMaxCols);
      hn[c.col] = (float)(c.isx ? hS : hC);                  // view 0: what student rows read for this column
      hn[kFastMaxCols + c.col] = (float)(c.isx ? hC : hS);   // view 1: what teacher rows read
      hnd[c.col] = c.isx ? hS : hC;
      hnd[kFastMaxCols + c.col] = c.isx ? hC : hS;
    }
    if (lane == 0 || lane == c.nstu)   // next round's centres (all lanes read the current ones at the top of this round)
      *reinterpret_cast<double2*>(c.ctr + (lane == 0 ? 0 : 2)) = make_double2(potS, potC);
    __syncwarp();
    hmag = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hmag)));  // non-negative floats order like their bits
    cur ^= 1;
  }
  const int r = nrounds - 1;
  if (r >= 32 && (r & 31) == 0)
    mine = make_round_const(r + lane, is, prm.sp);
  rc_last.coef = __shfl_sync(0xffffffffu, mine.coef, r & 31);
  rc_last.scale = __shfl_sync(0xffffffffu, mine.scale, r & 31);
  rc_last.hmul = 0.f;
  rc_last.eps = __shfl_sync(0xffffffffu, mine.eps, r & 31);
  rc_last.coefd = __shfl_sync(0xffffffffu, mine.coefd, r & 31);
  rc_last.scaled = __shfl_sync(0xffffffffu, mine.scaled, r & 31);
  rc_last.hmuld = 0.0;
  double lseX, lseY;
  float sX, sY, gXx, gXy, gYx, gYy;
  const float* hp = c.hb + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
  if (!(is_hi_round(r, nrounds, rc_last.eps, eps0, kSmallHiRounds) && hmag * hi_mag_factor(r, nrounds) > 1.0f)) {
    const RoundOutRt o = fast_round_rt<true>(c.cx, c.cy, hp, c.nchx, nch, c.px, c.py, rc_last.coef);
    lseX = o.lseX; lseY = o.lseY; sX = o.sX; sY = o.sY;
    gXx = o.gXx; gXy = o.gXy; gYx = o.gYx; gYy = o.gYy;
  } else {
    const double* hpd = c.hbd + cur * (2 * kFastMaxCols) + (c.isx ? 0 : kFastMaxCols);
    const float2 npx = make_float2(-c.px, -c.px), npy = make_float2(-c.py, -c.py), coef2 = make_float2(rc_last.coef, rc_last.coef);
    double refX = (double)rolled_max(c.cx, c.cy, hp, 0, c.nchx, npx, npy, coef2) + 2.0;
    double refY = (double)rolled_max(c.cx, c.cy, hp, c.nchx, nch, npx, npy, coef2) + 2.0;
    if ((hmag + (float)fabs(refX) + (float)fabs(refY)) * 9.5e-7f > 1.0f) hi_reference_f64(c.d2s, hpd, 4 * nch, 4 * c.nchx, rc_last.coefd, refX, refY);
    hi_sum_rt<true>(c.d2s, hpd, c.cx, c.cy, 0, c.nchx, c.px, c.py, rc_last.coefd, refX, sX, gXx, gXy);
    hi_sum_rt<true>(c.d2s, hpd, c.cx, c.cy, c.nchx, nch, c.px, c.py, rc_last.coefd, refY, sY, gYx, gYy);
    lseX = refX + lg2_sum(sX); lseY = refY + lg2_sum(sY);
  }
  lseX += addX; lseY += addY;
  S_out = rc_last.scaled * (c.isx ? lseX : lseY);
  C_out = rc_last.scaled * (c.isx ? lseY : lseX);
  // barycentric displacements  sum_j W_ij (p_j - p_i)  against own / other cloud (student rows use them)
  gSx = gXx / sX; gSy = gXy / sX;
  gCx = gYx / sY; gCy = gYy / sY;
}


// The B slots of an image are spread over a thread-block cluster of `split` CTAs (B/split warps each) so that a
// small batch still covers the whole chip with about one warp per SM sub-partition: the rounds are bound by the
// per-sub-partition SFU / FP32 pipes, not by occupancy.  The cluster is only needed twice, and never waited on before the
// rounds: a split barrier around the in-place normalisation (arrive once this warp has read the raw points of all B
// slots for the bounding box, wait -- long since complete -- before the normalised store at the end), and the fixed-order
// sum over slots, which rank 0 performs on values its peers wrote into its shared memory (DSMEM).
// ROLLED = false: per-chunk-count unrolled rounds, arguments held in registers (176 registers, one CTA per SM): lowest latency
// when the batch gives every SM sub-partition at most one warp (ape_b64: 26.6 vs 28.9 us).  ROLLED = true: the rolled rounds
// (120 registers, two CTAs per SM, L0-resident loops): 20 % more throughput once the grid exceeds one CTA per SM (1024 images:
// 111 vs 133 us).  launch_small picks by grid size.
template <bool ROLLED>
__global__ void __launch_bounds__(256, ROLLED ? 2 : 1) kdot_small_fast_kernel(SinkhornParams prm, int split) {
  cg::cluster_group cluster = cg::this_cluster();
  const int img = blockIdx.x / split, part = blockIdx.x - img * split;
  const int wpc = prm.B / split;  // warps (slots) per CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = part * wpc + warp;
  const int B = prm.B;
  // per warp: float64 h[2][2][40] | d2[40][32], then fp32 cx[40] | cy[40] | h[2][2][40], then the 4 float64 centres
  extern __shared__ __align__(16) double s_dynd[];
  __shared__ double s_slot_loss[16];               // rank 0's copy collects all B slots
  double* wbase_d = s_dynd + (size_t)warp * kFastWarpDoubles;
  float* wbase_s = reinterpret_cast<float*>(wbase_d + (4 + 32) * kFastMaxCols);

  const int n0 = prm.cu_n[img], N = prm.cu_n[img + 1] - n0;
  const int m0 = prm.cu_m[img], M = prm.cu_m[img + 1] - m0;
  const int P = N + M;
  const int Nq = (N + 3) & ~3, Mq = (M + 3) & ~3;

  dbg_stamp(prm, img, 0);
  if (prm.dbg_clk && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    prm.dbg_clk[(size_t)img * 16 + 7] = (long long)t;
  }
  FastCtx c;
  c.cx = wbase_s; c.cy = wbase_s + kFastMaxCols; c.hb = wbase_s + 2 * kFastMaxCols;
  c.hbd = wbase_d; c.d2s = wbase_d + 4 * kFastMaxCols + lane;
  c.ctr = wbase_d + (4 + 32 + 3) * kFastMaxCols;
  c.nchx = Nq >> 2;
  c.nstu = N;
  c.act = lane < P;
  c.isx = lane < N;
  c.col = c.isx ? lane : Nq + (lane - N);
  c.px = c.py = 0.f; c.wgt = 0.f; c.lw2 = 0.f; c.lw2d = 0.0;

  // ---- every warp loads ITS slot's points; the image-wide bounding box (geomloss' diameter spans all B slots) is
  //      taken by every warp on its own from the raw points of all slots -- B loads per lane that hit the lines its
  //      sibling warps fetch anyway -- instead of an exchange through (distributed) shared memory behind a cluster
  //      barrier.  Division by a positive width is monotonic, so the box of the normalised points is the normalised box.
  //      The in-place normalisation of the caller's buffer (loss_libs.py:8-12) is deferred until every warp of the
  //      cluster has read the raw values: barrier arrive here, wait before the store at the end. ----
  float minx = 3.0e38f, miny = 3.0e38f, maxx = -3.0e38f, maxy = -3.0e38f;
  long long gidx = 0;
  float* base = nullptr;
  float2 vnorm = make_float2(0.f, 0.f);
  if (c.act) {
    long long cell;
    long long s_cell, s_slot;
    const float* wsrc;
    if (c.isx) { base = prm.xs; cell = n0 + lane; s_cell = prm.s_cell_n; s_slot = prm.s_slot_n; wsrc = prm.ws; }
    else       { base = prm.xt; cell = m0 + lane - N; s_cell = prm.s_cell_m; s_slot = prm.s_slot_m; wsrc = prm.wt; }
    gidx = cell * s_cell + (long long)slot * s_slot;
    float2 v = *reinterpret_cast<const float2*>(base + 2 * gidx);
    c.wgt = wsrc ? wsrc[gidx] : __fdiv_rn(1.0f, (float)(c.isx ? N : M));
    const long long g0 = cell * s_cell;
    for (int sl = 0; sl < B; ++sl) {
      const float2 q = *reinterpret_cast<const float2*>(base + 2 * (g0 + (long long)sl * s_slot));
      minx = fminf(minx, q.x); maxx = fmaxf(maxx, q.x);
      miny = fminf(miny, q.y); maxy = fmaxf(maxy, q.y);
    }
    if (prm.normalize) {
      v.x = __fdiv_rn(v.x, prm.w);
      v.y = __fdiv_rn(v.y, prm.h);
    }
    vnorm = v;
    c.px = v.x; c.py = v.y;
    c.lw2 = (c.wgt > 0.f ? logf(c.wgt) : kLogZeroWeight) * kLog2e;
    c.lw2d = (double)c.lw2;
  }
  const bool store_norm = c.act && prm.normalize == 1;   // 2: keep the caller's buffer raw
  float* gx_out = prm.grad_xs + 2 * gidx;
  dbg_stamp(prm, img, 1);
  minx = warp_min(minx); miny = warp_min(miny); maxx = warp_max(maxx); maxy = warp_max(maxy);
  cluster.barrier_arrive();   // this warp's raw reads are done (the reductions above consumed them)
  if (prm.normalize) {
    minx = __fdiv_rn(minx, prm.w); maxx = __fdiv_rn(maxx, prm.w);
    miny = __fdiv_rn(miny, prm.h); maxy = __fdiv_rn(maxy, prm.h);
  }

  if (N == 0 || M == 0) {  // skipped image (loss_libs.py:25-28); uniform over the cluster
    if (c.act && c.isx) {
      *reinterpret_cast<float2*>(gx_out) = make_float2(0.f, 0.f);
      if (prm.grad_ws) prm.grad_ws[gidx] = 0.f;
    }
    if (lane == 0 && prm.loss_per_slot) prm.loss_per_slot[(size_t)img * B + slot] = 0.f;
    if (threadIdx.x == 0 && part == 0) {
      prm.loss_per_img[img] = 0.f;
      prm.valid[img] = KDOT_IMG_SKIPPED;
      if (prm.nits_per_img) prm.nits_per_img[img] = 0;
    }
    cluster.barrier_wait();
    if (store_norm) *reinterpret_cast<float2*>(base + 2 * gidx) = vnorm;
    return;
  }
  const float diam_f = bbox_diameter(minx, miny, maxx, maxy);
  dbg_stamp(prm, img, 2);
  int status = KDOT_IMG_OK, nits = 0, nrounds = 0;
  ImgSched is;
  is.nits = 0; is.start = 0.0; is.delta = 0.0; is.eps0 = 0.0;
  if (!(diam_f > 0.f) || !isfinite(diam_f)) {
    status = KDOT_IMG_DEGENERATE;
  } else {
    is = image_schedule_warp(diam_f, prm.sp, lane);
    nits = is.nits;
    nrounds = nits + 2;
    if (nrounds > KDOT_MAX_ROUNDS) status = KDOT_IMG_TOO_MANY_ROUNDS;
  }
  if (status != KDOT_IMG_OK) {  // uniform over the cluster
    const float nan = __int_as_float(0x7fc00000);
    if (c.act && c.isx) {
      *reinterpret_cast<float2*>(gx_out) = make_float2(nan, nan);
      if (prm.grad_ws) prm.grad_ws[gidx] = nan;
    }
    if (lane == 0 && prm.loss_per_slot) prm.loss_per_slot[(size_t)img * B + slot] = nan;
    if (threadIdx.x == 0 && part == 0) {
      prm.loss_per_img[img] = nan;
      prm.valid[img] = status;
      if (prm.nits_per_img) prm.nits_per_img[img] = nits;
    }
    cluster.barrier_wait();
    if (store_norm) *reinterpret_cast<float2*>(base + 2 * gidx) = vnorm;
    return;
  }

  dbg_stamp(prm, img, 3);
  // ---- stage this slot's columns: [student | pad | teacher | pad], pads carry h = -big (exp2 -> 0) ----
  // The float64 squared distances come first: the columns' coordinates are parked as doubles in the second h buffer
  // (not read before round 0 publishes into it), so the table costs two broadcast LDS.64, three DP operations and one
  // STS.64 per column -- no fp32 -> fp64 conversions (XU pipe) in the loop.
  double* cxd = c.hbd + 2 * kFastMaxCols;
  double* cyd = c.hbd + 3 * kFastMaxCols;
  for (int j = lane; j < kFastMaxCols; j += 32) {
    c.cx[j] = 0.f; c.cy[j] = 0.f;
    cxd[j] = 0.0; cyd[j] = 0.0;
  }
  if (lane < 4) c.ctr[lane] = 0.0;   // round 0 publishes uncentred offsets (potentials start at 0)
  __syncwarp();
  const double pxd = (double)c.px, pyd = (double)c.py;
  if (c.act) {
    c.cx[c.col] = c.px; c.cy[c.col] = c.py;
    cxd[c.col] = pxd; cyd[c.col] = pyd;
  }
  __syncwarp();
  {  // pads: columns at the origin, h = -big
    const int ncol = Nq + Mq;  // multiple of 4
#pragma unroll 2
    for (int j = 0; j < ncol; j += 4) {
      const double2 X0 = *reinterpret_cast<const double2*>(cxd + j), X1 = *reinterpret_cast<const double2*>(cxd + j + 2);
      const double2 Y0 = *reinterpret_cast<const double2*>(cyd + j), Y1 = *reinterpret_cast<const double2*>(cyd + j + 2);
      const double ax0 = X0.x - pxd, ax1 = X0.y - pxd, ax2 = X1.x - pxd, ax3 = X1.y - pxd;
      const double ay0 = Y0.x - pyd, ay1 = Y0.y - pyd, ay2 = Y1.x - pyd, ay3 = Y1.y - pyd;
      c.d2s[(j + 0) * 32] = fma(ay0, ay0, ax0 * ax0);
      c.d2s[(j + 1) * 32] = fma(ay1, ay1, ax1 * ax1);
      c.d2s[(j + 2) * 32] = fma(ay2, ay2, ax2 * ax2);
      c.d2s[(j + 3) * 32] = fma(ay3, ay3, ax3 * ax3);
    }
  }
  __syncwarp();
  for (int j = lane; j < kFastMaxCols; j += 32) {
#pragma unroll
    for (int v = 0; v < 4; ++v) { c.hb[v * kFastMaxCols + j] = kNegBig; c.hbd[v * kFastMaxCols + j] = (double)kNegBig; }
  }
  __syncwarp();
  if (c.act) {
    c.hb[c.col] = c.lw2; c.hb[kFastMaxCols + c.col] = c.lw2;  // init round: h = log w for both views
    c.hbd[c.col] = c.lw2d; c.hbd[kFastMaxCols + c.col] = c.lw2d;
  }
  __syncwarp();
  double S = 0.0, C = 0.0;
  float gSx = 0.f, gSy = 0.f, gCx = 0.f, gCy = 0.f;
  RoundConst rc;
  const RoundConst mine = make_round_const(lane, is, prm.sp);  // lane r holds the constants of round r
  dbg_stamp(prm, img, 4);
  const int ch = (Nq + Mq) >> 2;
  if (!ROLLED) {
    switch (ch) {
#define KDOT_CASE(K) case K: fast_solve<K>(c, nrounds, is, prm, mine, S, C, gSx, gSy, gCx, gCy, rc); break;
      KDOT_CASE(2) KDOT_CASE(3) KDOT_CASE(4) KDOT_CASE(5) KDOT_CASE(6) KDOT_CASE(7) KDOT_CASE(8) KDOT_CASE(9)
      default: fast_solve<kFastMaxCH>(c, nrounds, is, prm, mine, S, C, gSx, gSy, gCx, gCy, rc); break;
#undef KDOT_CASE
    }
  } else {
    fast_solve_rt(c, ch, nrounds, is, prm, mine, S, C, gSx, gSy, gCx, gCy, rc);
  }

  dbg_stamp(prm, img, 5);
  // ---- loss + analytic backward ----
  const double rho = prm.rho;
  // (rho + eps/2) / rho * lambda with lambda = 1 / (1 + eps/rho): one division
  const float gfac = rho < 0.0 ? 1.f : (float)((rho + 0.5 * (double)rc.eps) / (rho + (double)rc.eps));
  double loss = 0.0;
  if (c.act) {
    const RowFinal f = row_final(S, C, rho, rc.eps);
    loss = (double)c.wgt * (double)f.term;
    if (c.isx) {
      float gx = c.wgt * gfac * (f.eS * gSx - f.eC * gCx);
      float gy = c.wgt * gfac * (f.eS * gSy - f.eC * gCy);
      if (prm.normalize) {
        gx = __fdiv_rn(gx, prm.w);
        gy = __fdiv_rn(gy, prm.h);
      }
      *reinterpret_cast<float2*>(gx_out) = make_float2(gx, gy);
      if (prm.grad_ws) prm.grad_ws[gidx] = f.term;
    }
  }
  loss = warp_sum(loss);
  cluster.barrier_wait();   // completes the arrive of the prologue (long done): every CTA of the cluster is running and has
                            // read its raw points -> its shared memory may be written, the caller's buffer normalised
  if (store_norm) *reinterpret_cast<float2*>(base + 2 * gidx) = vnorm;
  if (lane == 0) {
    double* dst = split > 1 ? cluster.map_shared_rank(s_slot_loss, 0) : s_slot_loss;  // rank 0 collects
    dst[slot] = loss;
    if (prm.loss_per_slot) prm.loss_per_slot[(size_t)img * B + slot] = (float)loss;
  }
  cluster.sync();
  if (threadIdx.x == 0 && part == 0) {
    double tot = 0.0;
    for (int s = 0; s < B; ++s) tot += s_slot_loss[s];  // fixed order: deterministic
    if (prm.dbg_clk) prm.dbg_clk[(size_t)img * 16 + 6] = clock64();
    prm.loss_per_img[img] = (float)tot;
    prm.valid[img] = KDOT_IMG_OK;
    if (prm.nits_per_img) prm.nits_per_img[img] = nits;
  }
}

cudaError_t launch_small(const SinkhornParams& prm, int max_n, int max_m, cudaStream_t stream) {
  {
    // spread the slots of an image over a cluster until the grid has about one warp per SM sub-partition
    static int sm_count = 0;
    if (sm_count == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      if (sm_count <= 0) sm_count = 148;
    }
    int split = prm.B > 8 ? 2 : 1;  // at most 8 warps (slots) per CTA: the kernel is compiled for 256 threads
    while (split < 8 && prm.B % (split * 2) == 0 && (long long)prm.nimg * split * 2 <= (long long)sm_count) split *= 2;
    const int wpc = prm.B / split;
    const size_t smem = (size_t)wpc * kFastWarpDoubles * sizeof(double);
    const bool rolled = (long long)prm.nimg * split > (long long)sm_count;  // more than one CTA per SM: throughput variant
    static size_t configured[2] = {48 * 1024, 48 * 1024};
    if (smem > configured[rolled]) {
      cudaError_t e = rolled ? cudaFuncSetAttribute(kdot_small_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                             : cudaFuncSetAttribute(kdot_small_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured[rolled] = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(prm.nimg * split);
    cfg.blockDim = dim3(32 * wpc);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = split;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return rolled ? cudaLaunchKernelEx(&cfg, kdot_small_fast_kernel<true>, prm, split)
                  : cudaLaunchKernelEx(&cfg, kdot_small_fast_kernel<false>, prm, split);
  }
}

}  // namespace kdot
