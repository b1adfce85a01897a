// Fused OT distillation loss, small-problem kernel (N_i + M_i <= 64 cells per image, D = 2).
//
// One CTA per image, one warp per OT slot (keypoint).  Lane l owns point l (and l+32 when the image
// has more than 32 cells) of the concatenated cloud [student cells | teacher cells].  Everything the
// reference does for one image -- losses/loss_libs.py:8-12 (normalise), :22-50 (per-image split) and
// geomloss' tensorized Sinkhorn divergence with its backward -- runs in this single launch for the
// whole mini-batch; the N x M cost matrices live only in registers.
//
// Every point i carries two potentials: S_i against its own cloud (geomloss a_x / b_y) and C_i against
// the other cloud (b_x / a_y).  With h^S_j = log w_j + S_j/eps and h^C_j = log w_j + C_j/eps the four
// softmins of a Sinkhorn round collapse to one rule for every row i:
//   S_i <- lambda * softmin_{j in own cloud}(h^S_j),   C_i <- lambda * softmin_{j in other cloud}(h^C_j).
#include "kdot_common.cuh"

namespace kdot {

constexpr int kSmallMaxPts = 64;

struct RowAcc {
  float lse_x, lse_y;          // log2-sum-exp over the student / teacher columns
  float gxx, gxy, sx;          // sum e*(p_j - p_i) and sum e over the student columns   (grad rounds only)
  float gyx, gyy, sy;          // same over the teacher columns
};

// Two-pass (exact max) log2-sum-exp of row (px,py) against all P columns; the first N columns are the
// student cloud.  hx / hy select which of (hS,hC) applies to student / teacher columns for this row.
template <bool kGrad>
__device__ __forceinline__ RowAcc row_pass(const float2* __restrict__ pts, const float2* __restrict__ hb, int N, int P,
                                           float px, float py, bool row_is_x, float coef) {
  float mx = kNegBig, my = kNegBig;
  for (int j = 0; j < N; ++j) {
    const float2 q = pts[j], hh = hb[j];
    const float dx = q.x - px, dy = q.y - py;
    const float c = fmaf(dy, dy, dx * dx);
    mx = fmaxf(mx, fmaf(c, coef, row_is_x ? hh.x : hh.y));
  }
  for (int j = N; j < P; ++j) {
    const float2 q = pts[j], hh = hb[j];
    const float dx = q.x - px, dy = q.y - py;
    const float c = fmaf(dy, dy, dx * dx);
    my = fmaxf(my, fmaf(c, coef, row_is_x ? hh.y : hh.x));
  }
  RowAcc a;
  float sx = 0.f, sy = 0.f, gxx = 0.f, gxy = 0.f, gyx = 0.f, gyy = 0.f;
  for (int j = 0; j < N; ++j) {
    const float2 q = pts[j], hh = hb[j];
    const float dx = q.x - px, dy = q.y - py;
    const float c = fmaf(dy, dy, dx * dx);
    const float e = ex2_approx(fmaf(c, coef, row_is_x ? hh.x : hh.y) - mx);
    sx += e;
    if (kGrad) {
      gxx = fmaf(e, dx, gxx);
      gxy = fmaf(e, dy, gxy);
    }
  }
  for (int j = N; j < P; ++j) {
    const float2 q = pts[j], hh = hb[j];
    const float dx = q.x - px, dy = q.y - py;
    const float c = fmaf(dy, dy, dx * dx);
    const float e = ex2_approx(fmaf(c, coef, row_is_x ? hh.y : hh.x) - my);
    sy += e;
    if (kGrad) {
      gyx = fmaf(e, dx, gyx);
      gyy = fmaf(e, dy, gyy);
    }
  }
  a.lse_x = mx + log2f(sx);
  a.lse_y = my + log2f(sy);
  a.gxx = gxx; a.gxy = gxy; a.sx = sx;
  a.gyx = gyx; a.gyy = gyy; a.sy = sy;
  return a;
}

__global__ void __launch_bounds__(512) kdot_small_kernel(SinkhornParams prm) {
  const int img = blockIdx.x;
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = prm.B;

  __shared__ RoundConst s_rc[KDOT_MAX_ROUNDS];
  __shared__ float s_box[16][4];
  __shared__ double s_slot_loss[16];
  extern __shared__ float2 s_dyn[];  // per warp: pts[64] | h[2][64]
  float2* pts = s_dyn + (size_t)slot * (3 * kSmallMaxPts);
  float2* hbuf = pts + kSmallMaxPts;

  const int n0 = prm.cu_n[img], N = prm.cu_n[img + 1] - n0;
  const int m0 = prm.cu_m[img], M = prm.cu_m[img + 1] - m0;
  const int P = N + M;
  const bool two = P > 32;

  // ---- load, normalise in place, log-weights ------------------------------------------------------------
  float px[2], py[2], wgt[2], lw2[2];
  bool act[2], isx[2];
  long long gidx[2];  // element index of this (cell, slot) in xs/ws (student) or xt/wt (teacher)
  float minx = 3.0e38f, miny = 3.0e38f, maxx = -3.0e38f, maxy = -3.0e38f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int q = lane + 32 * k;
    act[k] = q < P;
    isx[k] = q < N;
    px[k] = py[k] = 0.f;
    wgt[k] = 0.f;
    lw2[k] = 0.f;
    gidx[k] = 0;
    if (act[k]) {
      float* base;
      const float* wbase;
      if (isx[k]) {
        gidx[k] = (long long)(n0 + q) * prm.s_cell_n + (long long)slot * prm.s_slot_n;
        base = prm.xs;
        wbase = prm.ws;
      } else {
        gidx[k] = (long long)(m0 + q - N) * prm.s_cell_m + (long long)slot * prm.s_slot_m;
        base = prm.xt;
        wbase = prm.wt;
      }
      float2 v = *reinterpret_cast<const float2*>(base + 2 * gidx[k]);
      if (prm.normalize) {
        v.x = __fdiv_rn(v.x, prm.w);
        v.y = __fdiv_rn(v.y, prm.h);
        *reinterpret_cast<float2*>(base + 2 * gidx[k]) = v;
      }
      px[k] = v.x;
      py[k] = v.y;
      wgt[k] = wbase ? wbase[gidx[k]] : __fdiv_rn(1.0f, (float)(isx[k] ? N : M));
      lw2[k] = (wgt[k] > 0.f ? logf(wgt[k]) : kLogZeroWeight) * kLog2e;
      minx = fminf(minx, v.x); maxx = fmaxf(maxx, v.x);
      miny = fminf(miny, v.y); maxy = fmaxf(maxy, v.y);
      pts[q] = v;
      hbuf[q] = make_float2(lw2[k], lw2[k]);  // init round: h = log w
    }
  }

  if (N == 0 || M == 0) {  // skipped image (loss_libs.py:25-28); uniform over the CTA
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (act[k] && isx[k]) {
        *reinterpret_cast<float2*>(prm.grad_xs + 2 * gidx[k]) = make_float2(0.f, 0.f);
        if (prm.grad_ws) prm.grad_ws[gidx[k]] = 0.f;
      }
    if (lane == 0 && prm.loss_per_slot) prm.loss_per_slot[(size_t)img * B + slot] = 0.f;
    if (threadIdx.x == 0) {
      prm.loss_per_img[img] = 0.f;
      prm.valid[img] = KDOT_IMG_SKIPPED;
      if (prm.nits_per_img) prm.nits_per_img[img] = 0;
    }
    return;
  }

  // ---- image-wide bounding box -> diameter -> eps schedule (float64, as numpy does) ---------------------
  minx = warp_min(minx); miny = warp_min(miny); maxx = warp_max(maxx); maxy = warp_max(maxy);
  if (lane == 0) {
    s_box[slot][0] = minx; s_box[slot][1] = miny; s_box[slot][2] = maxx; s_box[slot][3] = maxy;
  }
  __syncthreads();
  for (int s = 0; s < B; ++s) {
    minx = fminf(minx, s_box[s][0]); miny = fminf(miny, s_box[s][1]);
    maxx = fmaxf(maxx, s_box[s][2]); maxy = fmaxf(maxy, s_box[s][3]);
  }
  const float diam_f = bbox_diameter(minx, miny, maxx, maxy);
  int status = KDOT_IMG_OK;
  int nits = 0, nrounds = 0;
  if (!(diam_f > 0.f) || !isfinite(diam_f)) {
    status = KDOT_IMG_DEGENERATE;
  } else {
    double start, delta;
    nits = schedule_len((double)diam_f, prm.p, prm.blur, prm.scaling, &start, &delta);
    nrounds = nits + 2;
    if (nrounds > KDOT_MAX_ROUNDS) {
      status = KDOT_IMG_TOO_MANY_ROUNDS;
    } else {
      for (int r = threadIdx.x; r < nrounds; r += blockDim.x)
        s_rc[r] = make_round_const(r, nits, (double)diam_f, prm.p, prm.blur, start, delta, prm.rho);
    }
  }
  if (status != KDOT_IMG_OK) {  // uniform over the CTA
    const float nan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (act[k] && isx[k]) {
        *reinterpret_cast<float2*>(prm.grad_xs + 2 * gidx[k]) = make_float2(nan, nan);
        if (prm.grad_ws) prm.grad_ws[gidx[k]] = nan;
      }
    if (lane == 0 && prm.loss_per_slot) prm.loss_per_slot[(size_t)img * B + slot] = nan;
    if (threadIdx.x == 0) {
      prm.loss_per_img[img] = nan;
      prm.valid[img] = status;
      if (prm.nits_per_img) prm.nits_per_img[img] = nits;
    }
    return;
  }
  __syncthreads();  // schedule visible (also orders the pts/hbuf writes of this warp)

  // ---- Sinkhorn rounds ------------------------------------------------------------------------------------
  float potS[2] = {0.f, 0.f}, potC[2] = {0.f, 0.f};
  int cur = 0;
  for (int r = 0; r < nrounds - 1; ++r) {
    const RoundConst rc = s_rc[r];
    const float2* hb = hbuf + cur * kSmallMaxPts;
    float2* hn = hbuf + (cur ^ 1) * kSmallMaxPts;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !two) break;
      if (act[k]) {
        const RowAcc a = row_pass<false>(pts, hb, N, P, px[k], py[k], isx[k], rc.coef);
        const float nS = rc.scale * (isx[k] ? a.lse_x : a.lse_y);
        const float nC = rc.scale * (isx[k] ? a.lse_y : a.lse_x);
        if (r == 0) {
          potS[k] = nS;
          potC[k] = nC;
        } else {
          potS[k] = 0.5f * (potS[k] + nS);
          potC[k] = 0.5f * (potC[k] + nC);
        }
        hn[lane + 32 * k] = make_float2(fmaf(potS[k], rc.hmul, lw2[k]), fmaf(potC[k], rc.hmul, lw2[k]));
      }
    }
    __syncwarp();
    cur ^= 1;
  }

  // ---- last extrapolation (all four potentials from the same old ones) + loss + analytic backward ---------
  const RoundConst rc = s_rc[nrounds - 1];
  const float2* hb = hbuf + cur * kSmallMaxPts;
  const double rho = prm.rho;
  const float lam = rho < 0.0 ? 1.f : (float)(1.0 / (1.0 + (double)rc.eps / rho));
  const float gfac = rho < 0.0 ? 1.f : (float)((rho + 0.5 * (double)rc.eps) / rho) * lam;
  double loss = 0.0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (k == 1 && !two) break;
    if (act[k]) {
      const RowAcc a = row_pass<true>(pts, hb, N, P, px[k], py[k], isx[k], rc.coef);
      const float S = rc.scale * (isx[k] ? a.lse_x : a.lse_y);
      const float C = rc.scale * (isx[k] ? a.lse_y : a.lse_x);
      const RowFinal f = row_final(S, C, rho, rc.eps);
      loss += (double)wgt[k] * (double)f.term;
      if (isx[k]) {
        // own cloud = student columns (gx*, sx), other cloud = teacher columns (gy*, sy)
        float gx = wgt[k] * gfac * (f.eS * (a.gxx / a.sx) - f.eC * (a.gyx / a.sy));
        float gy = wgt[k] * gfac * (f.eS * (a.gxy / a.sx) - f.eC * (a.gyy / a.sy));
        if (prm.normalize) {
          gx = __fdiv_rn(gx, prm.w);
          gy = __fdiv_rn(gy, prm.h);
        }
        *reinterpret_cast<float2*>(prm.grad_xs + 2 * gidx[k]) = make_float2(gx, gy);
        if (prm.grad_ws) prm.grad_ws[gidx[k]] = f.term;
      }
    }
  }
  loss = warp_sum(loss);
  if (lane == 0) {
    s_slot_loss[slot] = loss;
    if (prm.loss_per_slot) prm.loss_per_slot[(size_t)img * B + slot] = (float)loss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int s = 0; s < B; ++s) tot += s_slot_loss[s];
    prm.loss_per_img[img] = (float)tot;
    prm.valid[img] = KDOT_IMG_OK;
    if (prm.nits_per_img) prm.nits_per_img[img] = nits;
  }
}

cudaError_t launch_small(const SinkhornParams& prm, cudaStream_t stream) {
  const size_t smem = (size_t)prm.B * 3 * kSmallMaxPts * sizeof(float2);
  kdot_small_kernel<<<prm.nimg, 32 * prm.B, smem, stream>>>(prm);
  return cudaGetLastError();
}

}  // namespace kdot
