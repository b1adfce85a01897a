// The two dense losses that sit beside the OT term in KDPoseLoss.__call__:
// after the OT kernel they were what the step consisted of -- ~250 small torch launches for ~0.02 ms of arithmetic.
//
//   kdot_focal_loss_fwd_bwd   SigmoidFocalLoss (reference losses/loss.py:12-40) over EVERY cell and class, forward and
//                             gradient in one pass, reading the per-level (nimg, C, H, W) logits in place -- no
//                             permute / reshape / cat of the class maps (losses/loss.py:62-96), no boolean-index copies
//                             (kd_loss.py:133-134): ignored cells (label -1) simply contribute nothing.
//   kdot_reg3d_loss_fwd_bwd   the 3-D object-space SmoothL1 regression loss (reference losses/kd_loss.py:57-71 =
//                             losses/loss.py:129-162 after the decode): per positive cell and key-point, project the 3-D
//                             target onto the viewing ray of the predicted 2-D key-point, SmoothL1 at 0.02 diameters,
//                             forward and d/d(key-point) in one launch.
//
// Both are HBM-trivial (1.3 M logits / 5 k key-points at batch 64) and launch-bound: one launch each instead of ~60.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kdot.h"

namespace kdot {

constexpr int kLossMaxLevels = 8;
constexpr int kLossThreads = 256;

struct FocalParams {
  const float* cls[kLossMaxLevels];
  float* gcls[kLossMaxLevels];
  int hw[kLossMaxLevels];
  int off[kLossMaxLevels + 1];             // prefix sums of hw (cells)
  long long eoff[kLossMaxLevels + 1];      // prefix sums of nimg * C * hw (elements)
  int nlvl, nimg, C;
  const int64_t* labels;                   // [nimg * cells]: -1 ignored, 0 background, c + 1 positive of class c
  float gamma, alpha, eps;
  double* partial;                         // [gridDim.x]
  unsigned int* ticket;                    // [1], zero on entry, zero on exit
  float* loss;                             // [1]
  int write_grad;
};

__device__ __forceinline__ float pow_gamma(float b, float gamma) { return gamma == 2.0f ? b * b : powf(b, gamma); }

__global__ void __launch_bounds__(kLossThreads) kdot_focal_kernel(FocalParams p) {
  const int cells = p.off[p.nlvl];
  const long long total = p.eoff[p.nlvl];
  double acc = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < p.nlvl && e >= p.eoff[l + 1]) ++l;
    const long long r = e - p.eoff[l];
    const int hw = p.hw[l];
    const int cell = (int)(r % hw);
    const int c = (int)((r / hw) % p.C);
    const int img = (int)(r / ((long long)hw * p.C));
    const long long t = p.labels[(long long)img * cells + p.off[l] + cell];
    float g = 0.f;
    if (t >= 0) {
      const float x = p.cls[l][r];
      const float s = 1.0f / (1.0f + expf(-x));
      const bool clamped = s < p.eps || s > 1.0f - p.eps;
      const float q = fminf(fmaxf(s, p.eps), 1.0f - p.eps);
      float dq;  // d loss / d q
      if (t == c + 1) {
        const float w = pow_gamma(1.0f - q, p.gamma), lg = logf(q);
        acc -= (double)(p.alpha * w * lg);
        dq = p.alpha * (p.gamma * pow_gamma(1.0f - q, p.gamma - 1.0f) * lg - w / q);
      } else {
        const float w = pow_gamma(q, p.gamma), lg = logf(1.0f - q);
        acc -= (double)((1.0f - p.alpha) * w * lg);
        dq = -(1.0f - p.alpha) * (p.gamma * pow_gamma(q, p.gamma - 1.0f) * lg - w / (1.0f - q));
      }
      g = clamped ? 0.f : dq * s * (1.0f - s);
    }
    if (p.write_grad) p.gcls[l][r] = g;
  }
  // deterministic reduction: fixed per-block order, then the last block sums the partials in index order
  __shared__ double s_red[kLossThreads / 32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) tsum += s_red[w];
    p.partial[blockIdx.x] = tsum;
    __threadfence();
    if (atomicAdd(p.ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      double tot = 0.0;
      for (unsigned int b = 0; b < gridDim.x; ++b) tot += ((volatile double*)p.partial)[b];
      *p.loss = (float)tot;
      *p.ticket = 0u;
    }
  }
}

struct Reg3dParams {
  const float* xy;        // [npos * 8][2] decoded key-points (full-image pixels)
  const float* target;    // [npos][8][3] 3-D key-points in the camera frame
  const float* diam;      // [npos] mesh diameter of the cell's class
  float kinv[9];          // inverse intrinsics, row-major
  int npos;
  float* loss_cell;       // [npos]
  float* g_xy;            // [npos * 8][2]
};

// one thread per (cell, key-point); the 8 key-points of a cell sit in 8 consecutive lanes
__global__ void __launch_bounds__(kLossThreads) kdot_reg3d_kernel(Reg3dParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = idx < p.npos * 8;
  float loss = 0.f;
  if (act) {
    const int cell = idx >> 3;
    const float2 v = *reinterpret_cast<const float2*>(p.xy + 2 * (size_t)idx);
    const float* k = p.kinv;
    const float r0 = k[0] * v.x + k[1] * v.y + k[2], r1 = k[3] * v.x + k[4] * v.y + k[5], r2 = k[6] * v.x + k[7] * v.y + k[8];
    const float t0 = p.target[3 * (size_t)idx], t1 = p.target[3 * (size_t)idx + 1], t2 = p.target[3 * (size_t)idx + 2];
    const float rr = r0 * r0 + r1 * r1 + r2 * r2, rt = r0 * t0 + r1 * t1 + r2 * t2;
    const float s = rt / rr;
    const float inv_d = 1.0f / p.diam[cell];
    const float kk = 50.0f;  // SmoothL1 at 0.02 diameters (kd_loss.py:65)
    const float u0 = kk * (s * r0 - t0) * inv_d, u1 = kk * (s * r1 - t1) * inv_d, u2 = kk * (s * r2 - t2) * inv_d;
    auto sl1 = [](float u) { const float a = fabsf(u); return a < 1.0f ? 0.5f * u * u : a - 0.5f; };
    auto dsl1 = [](float u) { return fabsf(u) < 1.0f ? u : (u > 0.f ? 1.0f : -1.0f); };
    loss = (sl1(u0) + sl1(u1) + sl1(u2)) / (24.0f * kk);
    // d loss_cell / d q_c = dsl1(u_c) / (24 * diam);  q = s r,  ds = (t . dr) / rr - 2 rt (r . dr) / rr^2
    const float c = inv_d / 24.0f;
    const float a0 = dsl1(u0) * c, a1 = dsl1(u1) * c, a2 = dsl1(u2) * c;
    const float ar = a0 * r0 + a1 * r1 + a2 * r2;          // (a . r): multiplies ds
    float g[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float d0 = k[j], d1 = k[3 + j], d2 = k[6 + j];  // dr / d(x or y) = column j of K^-1
      const float ds = (t0 * d0 + t1 * d1 + t2 * d2) / rr - 2.0f * rt * (r0 * d0 + r1 * d1 + r2 * d2) / (rr * rr);
      g[j] = ar * ds + s * (a0 * d0 + a1 * d1 + a2 * d2);
    }
    *reinterpret_cast<float2*>(p.g_xy + 2 * (size_t)idx) = make_float2(g[0], g[1]);
  }
  // per-cell sum over its 8 key-points (fixed shuffle order: deterministic)
  loss += __shfl_xor_sync(0xffffffffu, loss, 1);
  loss += __shfl_xor_sync(0xffffffffu, loss, 2);
  loss += __shfl_xor_sync(0xffffffffu, loss, 4);
  if (act && (idx & 7) == 0) p.loss_cell[idx >> 3] = loss;
}

void count_launches(unsigned n);
}  // namespace kdot

using namespace kdot;

extern "C" {

int kdot_focal_loss_fwd_bwd(const float* const* cls_lvl, const int32_t* hw_lvl, int nlvl, int nimg, int C,
                            const int64_t* labels, float gamma, float alpha, float* loss, float* const* g_cls_lvl,
                            void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (nlvl <= 0 || nlvl > kLossMaxLevels || nimg <= 0 || C <= 0 || !cls_lvl || !hw_lvl || !labels || !loss || !workspace)
    return KDOT_E_BADARG;
  FocalParams p;
  p.nlvl = nlvl; p.nimg = nimg; p.C = C;
  p.off[0] = 0; p.eoff[0] = 0;
  for (int l = 0; l < nlvl; ++l) {
    p.cls[l] = cls_lvl[l];
    p.gcls[l] = g_cls_lvl ? g_cls_lvl[l] : nullptr;
    p.hw[l] = hw_lvl[l];
    p.off[l + 1] = p.off[l] + hw_lvl[l];
    p.eoff[l + 1] = p.eoff[l] + (long long)nimg * C * hw_lvl[l];
  }
  p.labels = labels; p.gamma = gamma; p.alpha = alpha; p.eps = 1e-4f;
  p.write_grad = g_cls_lvl != nullptr;
  const long long total = p.eoff[nlvl];
  int blocks = (int)((total + kLossThreads * 4 - 1) / (kLossThreads * 4));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (workspace_bytes < kdot_focal_workspace_bytes()) return KDOT_E_WORKSPACE;
  p.ticket = (unsigned int*)workspace;            // zeroed once by the caller, left zero by the kernel
  p.partial = (double*)((char*)workspace + 16);
  p.loss = loss;
  kdot_focal_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)cuda_stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}

size_t kdot_focal_workspace_bytes(void) { return 16 + (size_t)148 * 8 * sizeof(double); }

int kdot_reg3d_loss_fwd_bwd(const float* xy, const float* target3d, const float* diam_cell, const float* kinv9_host,
                            int npos, float* loss_cell, float* g_xy, void* cuda_stream) {
  if (npos < 0 || !kinv9_host) return KDOT_E_BADARG;
  if (npos == 0) return KDOT_OK;
  if (!xy || !target3d || !diam_cell || !loss_cell || !g_xy) return KDOT_E_BADARG;
  Reg3dParams p;
  p.xy = xy; p.target = target3d; p.diam = diam_cell; p.npos = npos; p.loss_cell = loss_cell; p.g_xy = g_xy;
  for (int i = 0; i < 9; ++i) p.kinv[i] = kinv9_host[i];
  const int blocks = (npos * 8 + kLossThreads - 1) / kLossThreads;
  kdot_reg3d_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)cuda_stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}

}  // extern "C"
