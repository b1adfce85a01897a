// Fused OT distillation loss, CTA-resident ("tiled") kernel for mid-size clouds (33..256 points per image, D = 2;
// anything that fits 227 KB of shared memory with KDOT_FORCE_PATH=tiled).
//
// Replaces the same reference code as kdot_small.cu (losses/loss_libs.py:8-12,22-50 + geomloss'
// tensorized Sinkhorn divergence + its autograd backward).
//
// Launch 1 (kdot_prep_kernel, one CTA per image): in-place normalisation, image-wide bounding box,
//   geomloss' float64 epsilon schedule -> per-round fp32 constants in HBM scratch, rank of the image by size.
// Launch 2 (kdot_tiled_kernel, one CTA per (image, slot), largest images first): the whole cloud
//   [student | teacher] of that slot is staged ONCE in shared memory as SoA columns (x[], y[], h^S[], h^C[]); every
//   lane owns kRows rows and streams all columns with broadcast LDS.128, evaluating |x_i - y_j|^2, the log2-domain
//   soft-min argument, a lazily rescaled online max and the exp2 sum entirely in registers with packed f32x2
//   arithmetic; one __syncthreads per Sinkhorn round.  The N x M cost matrix is never materialised; HBM traffic is
//   the inputs once and the gradients once.  Bound: SFU ex2 (1 per pair) / FP32 pipe -- see DESIGN.md.
#include "kdot_common.cuh"

namespace kdot {

#ifndef KDOT_TILED_THREADS
#define KDOT_TILED_THREADS 256
#endif
// tuning (tools/size_sweep.py on B200, 64 images): this kernel serves the 65..256-point range, where one row per lane
// (32-row warp units: 16 units per round at N = M = 128 for the CTA's 8 warps) at 64 registers / 4 CTAs per SM beats
// 2 and 4 rows per lane; above ~256 points the streaming kernel's chip-wide balance wins (kdot_api.cu: choose_path)
#ifndef KDOT_TILED_ROWS
#define KDOT_TILED_ROWS 1
#endif
#ifndef KDOT_TILED_MINBLOCKS
#define KDOT_TILED_MINBLOCKS 4
#endif
constexpr int kTiledThreads = KDOT_TILED_THREADS;
constexpr int kRows = KDOT_TILED_ROWS;       // rows per lane
constexpr int kUnitRows = 32 * kRows;        // rows per warp unit
constexpr float kTau = 24.f;                 // lazy-rescale threshold (log2 units)

__device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// ---------------------------------------------------------------------------------------------------------
// prep: normalise, bounding box, schedule
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kdot_prep_kernel(SinkhornParams prm) {
  const int img = blockIdx.x;
  const int B = prm.B;
  const int n0 = prm.cu_n[img], N = prm.cu_n[img + 1] - n0;
  const int m0 = prm.cu_m[img], M = prm.cu_m[img + 1] - m0;
  __shared__ float s_box[8][4];
  __shared__ int s_info[2];
  __shared__ ImgSched s_is;

  float minx = 3.0e38f, miny = 3.0e38f, maxx = -3.0e38f, maxy = -3.0e38f;
  for (int t = threadIdx.x; t < (N + M) * B; t += blockDim.x) {
    const int q = t / B, slot = t - q * B;
    float* base;
    long long g;
    if (q < N) {
      base = prm.xs;
      g = (long long)(n0 + q) * prm.s_cell_n + (long long)slot * prm.s_slot_n;
    } else {
      base = prm.xt;
      g = (long long)(m0 + q - N) * prm.s_cell_m + (long long)slot * prm.s_slot_m;
    }
    float2 v = *reinterpret_cast<const float2*>(base + 2 * g);
    if (prm.normalize) {
      v.x = __fdiv_rn(v.x, prm.w);
      v.y = __fdiv_rn(v.y, prm.h);
      *reinterpret_cast<float2*>(base + 2 * g) = v;
    }
    minx = fminf(minx, v.x); maxx = fmaxf(maxx, v.x);
    miny = fminf(miny, v.y); maxy = fmaxf(maxy, v.y);
  }
  // Rank of this image by pair count, descending (ties: lower index first).  The main kernel maps CTA k to the image
  // of rank k / B: the hardware hands consecutive CTAs to different SMs, so every SM receives one problem from each
  // size quantile and the single wave of resident CTAs is balanced even when cloud sizes vary several-fold.
  {
    const long long mine = (long long)(N + M) * (N + M);
    int ahead = 0;
    for (int j = threadIdx.x; j < prm.nimg; j += blockDim.x) {
      const long long pj = (long long)(prm.cu_n[j + 1] - prm.cu_n[j]) + (long long)(prm.cu_m[j + 1] - prm.cu_m[j]);
      const long long other = pj * pj;
      ahead += (other > mine || (other == mine && j < img)) ? 1 : 0;
    }
    __shared__ int s_rank;
    if (threadIdx.x == 0) s_rank = 0;
    __syncthreads();
    ahead = __reduce_add_sync(0xffffffffu, ahead);
    if ((threadIdx.x & 31) == 0 && ahead) atomicAdd(&s_rank, ahead);
    __syncthreads();
    if (threadIdx.x == 0) prm.order[s_rank] = img;
  }
  const bool skipped = (N == 0 || M == 0);
  minx = warp_min(minx); miny = warp_min(miny); maxx = warp_max(maxx); maxy = warp_max(maxy);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_box[warp][0] = minx; s_box[warp][1] = miny; s_box[warp][2] = maxx; s_box[warp][3] = maxy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wi = 1; wi < (int)(blockDim.x >> 5); ++wi) {
      minx = fminf(minx, s_box[wi][0]); miny = fminf(miny, s_box[wi][1]);
      maxx = fmaxf(maxx, s_box[wi][2]); maxy = fmaxf(maxy, s_box[wi][3]);
    }
    int status = KDOT_IMG_OK, nits = 0;
    float diam = 0.f;
    if (skipped) {
      status = KDOT_IMG_SKIPPED;
    } else {
      diam = bbox_diameter(minx, miny, maxx, maxy);
      if (!(diam > 0.f) || !isfinite(diam)) {
        status = KDOT_IMG_DEGENERATE;
      } else {
        s_is = image_schedule(diam, prm.sp);
        nits = s_is.nits;
        if (nits + 2 > KDOT_MAX_ROUNDS) status = KDOT_IMG_TOO_MANY_ROUNDS;
      }
    }
    s_info[0] = status;
    s_info[1] = nits;
    prm.sched_rounds[img] = status == KDOT_IMG_OK ? nits + 2 : 0;
    prm.done_ctr[img] = 0u;
    prm.valid[img] = status;
    if (prm.nits_per_img) prm.nits_per_img[img] = nits;
    if (status != KDOT_IMG_OK) prm.loss_per_img[img] = status == KDOT_IMG_SKIPPED ? 0.f : __int_as_float(0x7fc00000);
  }
  __syncthreads();
  const int status = s_info[0], nits = s_info[1];
  if (status == KDOT_IMG_OK) {
    for (int r = threadIdx.x; r < nits + 2; r += blockDim.x)
      prm.sched[(size_t)img * KDOT_MAX_ROUNDS + r] = make_round_const(r, s_is, prm.sp);
  } else {
    const float fill = status == KDOT_IMG_SKIPPED ? 0.f : __int_as_float(0x7fc00000);
    for (int t = threadIdx.x; t < N * B; t += blockDim.x) {
      const int q = t / B, slot = t - q * B;
      const long long g = (long long)(n0 + q) * prm.s_cell_n + (long long)slot * prm.s_slot_n;
      *reinterpret_cast<float2*>(prm.grad_xs + 2 * g) = make_float2(fill, fill);
      if (prm.grad_ws) prm.grad_ws[g] = fill;
    }
    if (prm.loss_per_slot)
      for (int s = threadIdx.x; s < B; s += blockDim.x) prm.loss_per_slot[(size_t)img * B + s] = fill;
  }
}

// ---------------------------------------------------------------------------------------------------------
// inner loop: log2-sum-exp of R rows against columns [c0, c1) (c0, c1 multiples of 4)
// ---------------------------------------------------------------------------------------------------------
template <bool kGrad>
struct RowState {
  float2 nx, ny;    // (-px, -px), (-py, -py)
  float2 nm;        // (-mref, -mref)
  float mref;       // reference exponent of the running sums (may be stale by up to ~kTau)
  float2 s;         // two partial exp sums of the current 32-column block
  float tot, comp;  // compensated (Kahan) total of the finished blocks, same scale; row sum = tot - comp (see row_fold)
  float2 gx, gy;    // sum e * (p_j - p_i), two partials each   (kGrad only)
};

// The running sums of a row are kept relative to a reference exponent `mref` that may be STALE: instead of tracking
// the exact running max (one FMNMX per pair plus compares), the hot loop just evaluates p = 2^(v - mref) and
// looks at the chunk's partial sum.  If it exceeds 2^kTau (in particular +inf: the very first chunk, where
// mref = -big) some v is more than ~kTau above the reference; only then the cold path recomputes that row's chunk
// with the exact max and re-bases the sums.  Values far BELOW the reference need no care (they underflow to the
// correct negligible contribution).  Hot path per 4 pairs: 4 FADD2 + 2 FMUL2 + 4 FFMA2 + 2 FADD2 + 4 MUFU + 2 FADD2
// + FADD + FSETP = 5.0 issue slots per pair against the SFU's 8 cycles per warp-wide ex2.
// Row sums: fp32 accumulators span 32 columns only and are folded into a Kahan pair -- two running fp32 accumulators over a
// 128-column row cost d/dx ~1e-5 in a solver that is otherwise exact (tools/rowsum_study.py).
template <bool kGrad>
__device__ __forceinline__ void row_fold(RowState<kGrad>& st) {
  const float y = __fsub_rn(__fadd_rn(st.s.x, st.s.y), st.comp);
  const float t = __fadd_rn(st.tot, y);
  st.comp = __fsub_rn(__fsub_rn(t, st.tot), y);
  st.tot = t;
  st.s = make_float2(0.f, 0.f);
}
template <bool kGrad>
__device__ __forceinline__ double row_lg2_sum(const RowState<kGrad>& st) {   // log2(tot - comp)
  return lg2_sum_exact(st.tot) - (st.tot > 0.f ? (double)__fdiv_rn(st.comp, st.tot) * 1.4426950408889634 : 0.0);
}

template <bool kGrad>
__device__ __forceinline__ void rows_vs_columns(RowState<kGrad> (&st)[kRows], const float* __restrict__ cx,
                                                const float* __restrict__ cy, const float* __restrict__ ch, int c0,
                                                int c1, float coef) {
  const float2 coef2 = make_float2(coef, coef);
  const float big = exp2f(kTau);
#pragma unroll 1
  for (int jb = c0; jb < c1; jb += 32) {
  const int je = min(jb + 32, c1);
#pragma unroll 1
  for (int j = jb; j < je; j += 4) {
    const float4 X = *reinterpret_cast<const float4*>(cx + j);
    const float4 Y = *reinterpret_cast<const float4*>(cy + j);
    const float4 H = *reinterpret_cast<const float4*>(ch + j);
    float2 ps[kRows], pgx[kRows], pgy[kRows];
    bool rebase = false;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const float2 d0 = __fadd2_rn(make_float2(X.x, X.y), st[k].nx);
      const float2 d1 = __fadd2_rn(make_float2(X.z, X.w), st[k].nx);
      const float2 e0 = __fadd2_rn(make_float2(Y.x, Y.y), st[k].ny);
      const float2 e1 = __fadd2_rn(make_float2(Y.z, Y.w), st[k].ny);
      const float2 q0 = __ffma2_rn(e0, e0, __fmul2_rn(d0, d0));
      const float2 q1 = __ffma2_rn(e1, e1, __fmul2_rn(d1, d1));
      // v - mref with the subtraction folded into the per-row offset hm = (H - mref) would cost the same FADD2
      const float2 a0 = __fadd2_rn(__ffma2_rn(q0, coef2, make_float2(H.x, H.y)), st[k].nm);
      const float2 a1 = __fadd2_rn(__ffma2_rn(q1, coef2, make_float2(H.z, H.w)), st[k].nm);
      const float2 p0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
      const float2 p1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
      ps[k] = __fadd2_rn(p0, p1);
      if (kGrad) {
        pgx[k] = __ffma2_rn(p1, d1, __fmul2_rn(p0, d0));
        pgy[k] = __ffma2_rn(p1, e1, __fmul2_rn(p0, e0));
      }
      rebase |= !(ps[k].x + ps[k].y <= big);
    }
    if (rebase) {  // cold: recompute the offending rows' chunk against its exact max and re-base their sums
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        if (ps[k].x + ps[k].y <= big) continue;
        const float2 d0 = __fadd2_rn(make_float2(X.x, X.y), st[k].nx);
        const float2 d1 = __fadd2_rn(make_float2(X.z, X.w), st[k].nx);
        const float2 e0 = __fadd2_rn(make_float2(Y.x, Y.y), st[k].ny);
        const float2 e1 = __fadd2_rn(make_float2(Y.z, Y.w), st[k].ny);
        const float2 v0 = __ffma2_rn(__ffma2_rn(e0, e0, __fmul2_rn(d0, d0)), coef2, make_float2(H.x, H.y));
        const float2 v1 = __ffma2_rn(__ffma2_rn(e1, e1, __fmul2_rn(d1, d1)), coef2, make_float2(H.z, H.w));
        const float vm = fmaxf(fmaxf(v0.x, v0.y), fmaxf(v1.x, v1.y));
        const float sc = ex2_approx(st[k].mref - vm);  // 0 for the first chunk (mref = -big)
        st[k].s.x *= sc; st[k].s.y *= sc;
        st[k].tot *= sc; st[k].comp *= sc;
        if (kGrad) {
          st[k].gx.x *= sc; st[k].gx.y *= sc;
          st[k].gy.x *= sc; st[k].gy.y *= sc;
        }
        st[k].mref = vm;
        st[k].nm = make_float2(-vm, -vm);
        const float2 p0 = make_float2(ex2_approx(v0.x - vm), ex2_approx(v0.y - vm));
        const float2 p1 = make_float2(ex2_approx(v1.x - vm), ex2_approx(v1.y - vm));
        ps[k] = __fadd2_rn(p0, p1);
        if (kGrad) {
          pgx[k] = __ffma2_rn(p1, d1, __fmul2_rn(p0, d0));
          pgy[k] = __ffma2_rn(p1, e1, __fmul2_rn(p0, e0));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      st[k].s = __fadd2_rn(st[k].s, ps[k]);
      if (kGrad) {
        st[k].gx = __fadd2_rn(st[k].gx, pgx[k]);
        st[k].gy = __fadd2_rn(st[k].gy, pgy[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kRows; ++k) row_fold(st[k]);
  }
}

// High-precision sweep (last KDOT_HI_ROUNDS rounds, kdot_common.cuh): same structure, but the soft-min argument
// t = h_j + coef * |p_i - p_j|^2 - mref is formed in float64 from float64 copies of the columns, offsets and reference;
// only the small difference is rounded to fp32 for the exponential.  Sums and gradient accumulators stay fp32 (they are
// sums of positive terms / of terms weighted by exact fp32 coordinate differences).
template <bool kGrad>
struct RowStateHi {
  double px, py;
  double mref;
  float pxf, pyf;
  float2 s;
  float2 gx, gy;
};

template <bool kGrad>
__device__ __forceinline__ void rows_vs_columns_hi(RowStateHi<kGrad>& st, const double* __restrict__ cxd,
                                                   const double* __restrict__ cyd, const double* __restrict__ chd,
                                                   const float* __restrict__ cx, const float* __restrict__ cy, int c0,
                                                   int c1, double coef) {
  const float big = exp2f(kTau);
#pragma unroll 1
  for (int j = c0; j < c1; j += 4) {
    const double2 X0 = *reinterpret_cast<const double2*>(cxd + j), X1 = *reinterpret_cast<const double2*>(cxd + j + 2);
    const double2 Y0 = *reinterpret_cast<const double2*>(cyd + j), Y1 = *reinterpret_cast<const double2*>(cyd + j + 2);
    const double2 H0 = *reinterpret_cast<const double2*>(chd + j), H1 = *reinterpret_cast<const double2*>(chd + j + 2);
    const double ax0 = X0.x - st.px, ax1 = X0.y - st.px, ax2 = X1.x - st.px, ax3 = X1.y - st.px;
    const double ay0 = Y0.x - st.py, ay1 = Y0.y - st.py, ay2 = Y1.x - st.py, ay3 = Y1.y - st.py;
    const double v0 = fma(coef, fma(ay0, ay0, ax0 * ax0), H0.x), v1 = fma(coef, fma(ay1, ay1, ax1 * ax1), H0.y);
    const double v2 = fma(coef, fma(ay2, ay2, ax2 * ax2), H1.x), v3 = fma(coef, fma(ay3, ay3, ax3 * ax3), H1.y);
    float p0 = ex2_approx(f64_to_f32_trunc(v0 - st.mref)), p1 = ex2_approx(f64_to_f32_trunc(v1 - st.mref));
    float p2 = ex2_approx(f64_to_f32_trunc(v2 - st.mref)), p3 = ex2_approx(f64_to_f32_trunc(v3 - st.mref));
    if (!((p0 + p1) + (p2 + p3) <= big)) {  // cold: re-base on the exact max of this chunk
      const double vm = fmax(fmax(v0, v1), fmax(v2, v3));
      const float sc = ex2_approx((float)(st.mref - vm));  // 0 for the first chunk (mref = -big)
      st.s.x *= sc; st.s.y *= sc;
      if (kGrad) { st.gx.x *= sc; st.gx.y *= sc; st.gy.x *= sc; st.gy.y *= sc; }
      st.mref = vm;
      p0 = ex2_approx(f64_to_f32_trunc(v0 - vm)); p1 = ex2_approx(f64_to_f32_trunc(v1 - vm));
      p2 = ex2_approx(f64_to_f32_trunc(v2 - vm)); p3 = ex2_approx(f64_to_f32_trunc(v3 - vm));
    }
    st.s.x += p0 + p2; st.s.y += p1 + p3;
    if (kGrad) {
      const float4 XF = *reinterpret_cast<const float4*>(cx + j);
      const float4 YF = *reinterpret_cast<const float4*>(cy + j);
      st.gx.x = fmaf(p0, XF.x - st.pxf, fmaf(p2, XF.z - st.pxf, st.gx.x));
      st.gx.y = fmaf(p1, XF.y - st.pxf, fmaf(p3, XF.w - st.pxf, st.gx.y));
      st.gy.x = fmaf(p0, YF.x - st.pyf, fmaf(p2, YF.z - st.pyf, st.gy.x));
      st.gy.y = fmaf(p1, YF.y - st.pyf, fmaf(p3, YF.w - st.pyf, st.gy.y));
    }
  }
}

template <bool kGrad>
__device__ __forceinline__ void rows_reset_hi(RowStateHi<kGrad>& st, float pxf, float pyf) {
  st.px = (double)pxf; st.py = (double)pyf; st.pxf = pxf; st.pyf = pyf;
  st.mref = (double)kNegBig;
  st.s = make_float2(0.f, 0.f);
  st.gx = make_float2(0.f, 0.f);
  st.gy = make_float2(0.f, 0.f);
}

template <bool kGrad>
__device__ __forceinline__ void rows_reset(RowState<kGrad> (&st)[kRows]) {
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    st[k].mref = kNegBig;
    st[k].nm = make_float2(-kNegBig, -kNegBig);
    st[k].s = make_float2(0.f, 0.f);
    st[k].tot = 0.f; st[k].comp = 0.f;
    st[k].gx = make_float2(0.f, 0.f);
    st[k].gy = make_float2(0.f, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// main kernel: one CTA per (image, slot)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTiledThreads, KDOT_TILED_MINBLOCKS) kdot_tiled_kernel(SinkhornParams prm) {
  const int B = prm.B;
  const int rank = blockIdx.x / B, slot = blockIdx.x - rank * B;
  const int img = prm.order[rank];  // largest images first (see the prep kernel)
  const int nrounds = prm.sched_rounds[img];
  if (nrounds <= 0) return;  // skipped / degenerate image: prep kernel wrote the outputs
  const int nits = nrounds - 2;
  const int n0 = prm.cu_n[img], N = prm.cu_n[img + 1] - n0;
  const int m0 = prm.cu_m[img], M = prm.cu_m[img + 1] - m0;
  const int Nq = round4(N), Mq = round4(M), Pq = Nq + Mq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  extern __shared__ __align__(16) double smem_d[];
  // float64 arrays first (alignment), then the fp32 ones: 8 doubles + 7 floats = 92 bytes per padded point
  double* cxd = smem_d;
  double* cyd = cxd + Pq;
  double* potS = cyd + Pq;
  double* potC = potS + Pq;
  double* hSd = potC + Pq;        // [2][Pq]
  double* hCd = hSd + 2 * Pq;     // [2][Pq]
  float* cx = reinterpret_cast<float*>(hCd + 2 * Pq);
  float* cy = cx + Pq;
  float* lw2 = cy + Pq;
  float* hS = lw2 + Pq;           // [2][Pq]
  float* hC = hS + 2 * Pq;        // [2][Pq]
  const RoundConst* __restrict__ rcs = prm.sched + (size_t)img * KDOT_MAX_ROUNDS;  // L1/L2-resident, 1 load / round
  __shared__ unsigned int s_ctr[2];
  __shared__ double s_pref[3][4];     // centres of the published offsets (see kdot_stream.cu: stream_unit), slot r % 3
  __shared__ unsigned int s_hmag[3];  // max |h| (fp32 bits) of the values consumed in round r, slot r % 3 (hi_mag_factor test)
  __shared__ double s_part[kTiledThreads / 32];

  // ---- stage the cloud of this slot (already normalised by the prep kernel) ----
  for (int q = threadIdx.x; q < Pq; q += blockDim.x) {
    const bool in_x = q < Nq;
    const int i = in_x ? q : q - Nq;
    const bool real = in_x ? (i < N) : (i < M);
    float2 v = make_float2(0.f, 0.f);
    float l2 = kNegBig;
    if (real) {
      long long g;
      const float* base;
      const float* wbase;
      if (in_x) {
        g = (long long)(n0 + i) * prm.s_cell_n + (long long)slot * prm.s_slot_n;
        base = prm.xs; wbase = prm.ws;
      } else {
        g = (long long)(m0 + i) * prm.s_cell_m + (long long)slot * prm.s_slot_m;
        base = prm.xt; wbase = prm.wt;
      }
      v = *reinterpret_cast<const float2*>(base + 2 * g);
      const float wg = wbase ? wbase[g] : __fdiv_rn(1.0f, (float)(in_x ? N : M));
      l2 = (wg > 0.f ? logf(wg) : kLogZeroWeight) * kLog2e;
    }
    cx[q] = v.x; cy[q] = v.y; lw2[q] = l2;
    cxd[q] = (double)v.x; cyd[q] = (double)v.y;
    potS[q] = 0.0; potC[q] = 0.0;
    hS[q] = l2; hC[q] = l2;             // init round: h = log w  (pads: -big => exp2 -> 0)
    hS[Pq + q] = l2; hC[Pq + q] = l2;   // pads of the second buffer stay -big forever
    hSd[q] = (double)l2; hCd[q] = (double)l2;
    hSd[Pq + q] = (double)l2; hCd[Pq + q] = (double)l2;
  }
  if (threadIdx.x == 0) { s_ctr[0] = 0u; s_ctr[1] = 0u; s_hmag[0] = 0u; s_hmag[1] = 0u; s_hmag[2] = 0u; }
  if (threadIdx.x < 12) s_pref[threadIdx.x >> 2][threadIdx.x & 3] = 0.0;
  __syncthreads();

  const int nbx = (N + kUnitRows - 1) / kUnitRows, nby = (M + kUnitRows - 1) / kUnitRows;

  // ---- init + loop rounds ----
  int cur = 0;
  for (int r = 0; r < nrounds - 1; ++r) {
    const RoundConst rc = rcs[r];
    const double hmul_prev = r > 0 ? rcs[r - 1].hmuld : 0.0;
    const bool hi = is_hi_round(r, nrounds, rc.eps, rcs[0].eps) && __uint_as_float(s_hmag[r % 3]) * hi_mag_factor(r, nrounds) > 1.0f;
    const float* hSc = hS + cur * Pq;
    const float* hCc = hC + cur * Pq;
    float* hSn = hS + (cur ^ 1) * Pq;
    float* hCn = hC + (cur ^ 1) * Pq;
    const double* hScd = hSd + cur * Pq;
    const double* hCcd = hCd + cur * Pq;
    double* hSnd = hSd + (cur ^ 1) * Pq;
    double* hCnd = hCd + (cur ^ 1) * Pq;
    const int nunits = 2 * (nbx + nby);
    for (;;) {
      unsigned int u = 0;
      if (lane == 0) u = atomicAdd(&s_ctr[r & 1], 1u);
      u = __shfl_sync(0xffffffffu, u, 0);
      if ((int)u >= nunits) break;
      const bool rows_x = (int)u < 2 * nbx;
      const int ub = rows_x ? (int)u : (int)u - 2 * nbx;
      const int blk = ub >> 1;
      const bool own = (ub & 1) == 0;
      const int rbase = rows_x ? 0 : Nq, rcount = rows_x ? N : M;
      const bool cols_x = (rows_x == own);
      const int c0 = cols_x ? 0 : Nq, c1 = cols_x ? Nq : Pq;
      // centred offsets (the common rho * log(mass ratio) / eps term of a cloud's potentials stays out of the fp32 heads):
      // consumed h is relative to c_in, published h to p_out; the owner of a set's first row records its new potential
      // for the h consumed two rounds later (three slots: no reader and writer ever share one within a round)
      const double c_in = s_pref[r % 3][(own ? 0 : 2) + (cols_x ? 0 : 1)] * hmul_prev;
      const double p_out = s_pref[(r + 1) % 3][(own ? 0 : 2) + (rows_x ? 0 : 1)];

      int ridx[kRows];
      double lse[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const int i = blk * kUnitRows + lane + 32 * k;
        ridx[k] = i < rcount ? rbase + i : -1;
      }
      if (!hi) {
        RowState<false> st[kRows];
        rows_reset(st);
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const int src = ridx[k] >= 0 ? ridx[k] : rbase;
          st[k].nx = make_float2(-cx[src], -cx[src]);
          st[k].ny = make_float2(-cy[src], -cy[src]);
        }
        rows_vs_columns<false>(st, cx, cy, own ? hSc : hCc, c0, c1, rc.coef);
#pragma unroll
        for (int k = 0; k < kRows; ++k) lse[k] = (double)st[k].mref + row_lg2_sum(st[k]);
      } else {
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const int src = ridx[k] >= 0 ? ridx[k] : rbase;
          RowStateHi<false> sh;
          rows_reset_hi(sh, cx[src], cy[src]);
          rows_vs_columns_hi<false>(sh, cxd, cyd, own ? hScd : hCcd, cx, cy, c0, c1, rc.coefd);
          lse[k] = sh.mref + lg2_sum_exact(sh.s.x + sh.s.y);
        }
      }
      float hm = 0.f;
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        if (ridx[k] < 0) continue;
        double* pot = own ? potS : potC;
        const double nv = rc.scaled * (lse[k] + c_in);
        const double pv = r == 0 ? nv : 0.5 * (pot[ridx[k]] + nv);
        pot[ridx[k]] = pv;
        if (ridx[k] == rbase) s_pref[(r + 2) % 3][(own ? 0 : 2) + (rows_x ? 0 : 1)] = pv;
        const double hv = fma(pv - p_out, rc.hmuld, (double)lw2[ridx[k]]);
        (own ? hSn : hCn)[ridx[k]] = (float)hv;
        (own ? hSnd : hCnd)[ridx[k]] = hv;
        hm = fmaxf(hm, fabsf((float)hv));
      }
      hm = warp_max(hm);
      if (lane == 0) atomicMax(&s_hmag[(r + 1) % 3], __float_as_uint(hm));  // non-negative floats order like their bits
    }
    if (threadIdx.x == 0) { s_ctr[(r & 1) ^ 1] = 0u; s_hmag[(r + 2) % 3] = 0u; }
    __syncthreads();
    cur ^= 1;
  }

  // ---- last extrapolation + loss + analytic backward ----
  {
    const int r = nrounds - 1;
    const RoundConst rc = rcs[r];
    const double hmul_prev = r > 0 ? rcs[r - 1].hmuld : 0.0;
    const double cSX = s_pref[r % 3][0] * hmul_prev, cSY = s_pref[r % 3][1] * hmul_prev;   // centres of h^S[X], h^S[Y]
    const double cCX = s_pref[r % 3][2] * hmul_prev, cCY = s_pref[r % 3][3] * hmul_prev;   // centres of h^C[X], h^C[Y]
    const bool hi = is_hi_round(r, nrounds, rc.eps, rcs[0].eps) && __uint_as_float(s_hmag[r % 3]) * hi_mag_factor(r, nrounds) > 1.0f;
    const float* hSc = hS + cur * Pq;
    const float* hCc = hC + cur * Pq;
    const double* hScd = hSd + cur * Pq;
    const double* hCcd = hCd + cur * Pq;
    float* term = hS + (cur ^ 1) * Pq;  // free buffer: per-row loss terms (weight * term)
    const double rho = prm.rho;
    const float lam = rho < 0.0 ? 1.f : (float)(1.0 / (1.0 + (double)rc.eps / rho));
    const float gfac = rho < 0.0 ? 1.f : (float)((rho + 0.5 * (double)rc.eps) / rho) * lam;
    const int nunits = nbx + 2 * nby;
    for (;;) {
      unsigned int u = 0;
      if (lane == 0) u = atomicAdd(&s_ctr[r & 1], 1u);
      u = __shfl_sync(0xffffffffu, u, 0);
      if ((int)u >= nunits) break;
      if ((int)u < nbx) {
        // student rows: both column sets in one unit so the gradient is finished in registers
        int ridx[kRows];
        double S[kRows], C[kRows];
        float gSx[kRows], gSy[kRows], gCx[kRows], gCy[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const int i = (int)u * kUnitRows + lane + 32 * k;
          ridx[k] = i < N ? i : -1;
        }
        if (!hi) {
          RowState<true> st[kRows];
          rows_reset(st);
#pragma unroll
          for (int k = 0; k < kRows; ++k) {
            const int src = ridx[k] >= 0 ? ridx[k] : 0;
            st[k].nx = make_float2(-cx[src], -cx[src]);
            st[k].ny = make_float2(-cy[src], -cy[src]);
          }
          rows_vs_columns<true>(st, cx, cy, hSc, 0, Nq, rc.coef);
#pragma unroll
          for (int k = 0; k < kRows; ++k) {
            const float s = st[k].tot;
            S[k] = rc.scaled * ((double)st[k].mref + row_lg2_sum(st[k]) + cSX);
            gSx[k] = (st[k].gx.x + st[k].gx.y) / s;
            gSy[k] = (st[k].gy.x + st[k].gy.y) / s;
          }
          rows_reset(st);
          rows_vs_columns<true>(st, cx, cy, hCc, Nq, Pq, rc.coef);
#pragma unroll
          for (int k = 0; k < kRows; ++k) {
            const float s = st[k].tot;
            C[k] = rc.scaled * ((double)st[k].mref + row_lg2_sum(st[k]) + cCY);
            gCx[k] = (st[k].gx.x + st[k].gx.y) / s;
            gCy[k] = (st[k].gy.x + st[k].gy.y) / s;
          }
        } else {
#pragma unroll
          for (int k = 0; k < kRows; ++k) {
            const int src = ridx[k] >= 0 ? ridx[k] : 0;
            RowStateHi<true> sh;
            rows_reset_hi(sh, cx[src], cy[src]);
            rows_vs_columns_hi<true>(sh, cxd, cyd, hScd, cx, cy, 0, Nq, rc.coefd);
            float s = sh.s.x + sh.s.y;
            S[k] = rc.scaled * (sh.mref + lg2_sum_exact(s) + cSX);
            gSx[k] = (sh.gx.x + sh.gx.y) / s;
            gSy[k] = (sh.gy.x + sh.gy.y) / s;
            rows_reset_hi(sh, cx[src], cy[src]);
            rows_vs_columns_hi<true>(sh, cxd, cyd, hCcd, cx, cy, Nq, Pq, rc.coefd);
            s = sh.s.x + sh.s.y;
            C[k] = rc.scaled * (sh.mref + lg2_sum_exact(s) + cCY);
            gCx[k] = (sh.gx.x + sh.gx.y) / s;
            gCy[k] = (sh.gy.x + sh.gy.y) / s;
          }
        }
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          if (ridx[k] < 0) continue;
          const RowFinal f = row_final(S[k], C[k], rho, rc.eps);
          const long long g = (long long)(n0 + ridx[k]) * prm.s_cell_n + (long long)slot * prm.s_slot_n;
          const float wg = prm.ws ? prm.ws[g] : __fdiv_rn(1.0f, (float)N);
          term[ridx[k]] = wg * f.term;
          float gx = wg * gfac * (f.eS * gSx[k] - f.eC * gCx[k]);
          float gy = wg * gfac * (f.eS * gSy[k] - f.eC * gCy[k]);
          if (prm.normalize) {
            gx = __fdiv_rn(gx, prm.w);
            gy = __fdiv_rn(gy, prm.h);
          }
          *reinterpret_cast<float2*>(prm.grad_xs + 2 * g) = make_float2(gx, gy);
          if (prm.grad_ws) prm.grad_ws[g] = f.term;
        }
      } else {
        // teacher rows: potentials only (no gradient flows to the teacher)
        const int ub = (int)u - nbx;
        const int blk = ub >> 1;
        const bool own = (ub & 1) == 0;
        const int c0 = own ? Nq : 0, c1 = own ? Pq : Nq;
        int ridx[kRows];
        double lse[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const int i = blk * kUnitRows + lane + 32 * k;
          ridx[k] = i < M ? Nq + i : -1;
        }
        if (!hi) {
          RowState<false> st[kRows];
          rows_reset(st);
#pragma unroll
          for (int k = 0; k < kRows; ++k) {
            const int src = ridx[k] >= 0 ? ridx[k] : Nq;
            st[k].nx = make_float2(-cx[src], -cx[src]);
            st[k].ny = make_float2(-cy[src], -cy[src]);
          }
          rows_vs_columns<false>(st, cx, cy, own ? hSc : hCc, c0, c1, rc.coef);
#pragma unroll
          for (int k = 0; k < kRows; ++k) lse[k] = (double)st[k].mref + row_lg2_sum(st[k]);
        } else {
#pragma unroll
          for (int k = 0; k < kRows; ++k) {
            const int src = ridx[k] >= 0 ? ridx[k] : Nq;
            RowStateHi<false> sh;
            rows_reset_hi(sh, cx[src], cy[src]);
            rows_vs_columns_hi<false>(sh, cxd, cyd, own ? hScd : hCcd, cx, cy, c0, c1, rc.coefd);
            lse[k] = sh.mref + lg2_sum_exact(sh.s.x + sh.s.y);
          }
        }
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          if (ridx[k] < 0) continue;
          (own ? potS : potC)[ridx[k]] = rc.scaled * (lse[k] + (own ? cSY : cCX));
        }
      }
    }
    __syncthreads();

    // deterministic reduction of the slot's loss: fixed thread-strided order, fp64 accumulation
    double acc = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) acc += (double)term[i];
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      const RowFinal f = row_final(potS[Nq + j], potC[Nq + j], rho, rc.eps);
      const long long g = (long long)(m0 + j) * prm.s_cell_m + (long long)slot * prm.s_slot_m;
      const float wg = prm.wt ? prm.wt[g] : __fdiv_rn(1.0f, (float)M);
      acc += (double)wg * (double)f.term;
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int wi = 0; wi < nwarps; ++wi) tot += s_part[wi];
      prm.slot_loss[(size_t)img * B + slot] = (float)tot;
      if (prm.loss_per_slot) prm.loss_per_slot[(size_t)img * B + slot] = (float)tot;
      __threadfence();
      const unsigned int done = atomicAdd(&prm.done_ctr[img], 1u);
      if (done == (unsigned int)(B - 1)) {  // last slot of the image: fixed-order sum over slots
        __threadfence();
        double t2 = 0.0;
        for (int s = 0; s < B; ++s) t2 += (double)((volatile float*)prm.slot_loss)[(size_t)img * B + s];
        prm.loss_per_img[img] = (float)t2;
      }
    }
  }
  (void)nits;
}

size_t tiled_smem_bytes(int max_n, int max_m) {
  const size_t pq = (size_t)((max_n + 3) & ~3) + (size_t)((max_m + 3) & ~3);
  return pq * (8 * sizeof(double) + 7 * sizeof(float));
}

cudaError_t launch_tiled(const SinkhornParams& prm, int max_n, int max_m, cudaStream_t stream, size_t smem_limit,
                         bool* too_large) {
  *too_large = false;
  const size_t smem = tiled_smem_bytes(max_n, max_m);
  if (smem > smem_limit) {
    *too_large = true;
    return cudaSuccess;
  }
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kdot_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  kdot_prep_kernel<<<prm.nimg, 256, 0, stream>>>(prm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  kdot_tiled_kernel<<<prm.nimg * prm.B, kTiledThreads, smem, stream>>>(prm);
  return cudaGetLastError();
}

}  // namespace kdot
