// Kernel-MMD sample losses ("gaussian", "laplacian", "energy"), fused forward + analytic backward.
//
// The reference exposes them through --gtype (arguments/argument_kd.py:41) -> SamplesLoss(GTYPE, blur=GBLUR)
// (losses/kd_loss.py:26-30); geomloss evaluates them with kernel_tensorized (geomloss/kernel_samples.py):
//   K_xx = k(x, x), K_yy = k(y, y), K_xy = k(x, y)
//   loss = 1/2 <alpha, K_xx alpha> + 1/2 <beta, K_yy beta> - <alpha, K_xy beta>
//   gaussian  k = exp(-|x-y|^2 / (2 blur^2)),  laplacian  k = exp(-r / blur),  energy  k = -r,
//   r = sqrt(max(|x-y|^2 [/ blur^2], 1e-8))   (the clamp also zeroes the gradient there).
// SURVEY.md section 8(f) item 4.  One CTA per image, slots in sequence, one row per thread (strided), columns read
// with uniform (broadcast) loads; the N x M kernel matrices are never materialised.  O(N*M) once -- no iterations.
#include "kdot_common.cuh"

namespace kdot {

enum { KDOT_MMD_GAUSSIAN = 0, KDOT_MMD_LAPLACIAN = 1, KDOT_MMD_ENERGY = 2 };
constexpr int kMmdThreads = 256;
constexpr int kMmdMaxD = 16;

struct MmdParams {
  SinkhornParams b;
  int D, kind;
  float blur;
};

// value and d/d(first argument) scale of the kernel for squared distance q (already divided by blur^2 for the
// gaussian / laplacian kernels): returns k and writes g such that  d k / d x_i = g * (x_i - y_j)
__device__ __forceinline__ float mmd_kernel(int kind, float q, float inv_blur, float& g) {
  if (kind == KDOT_MMD_GAUSSIAN) {
    const float k = expf(-0.5f * q);
    g = -k * inv_blur * inv_blur;
    return k;
  }
  const bool clamped = !(q > 1e-8f);
  const float r = sqrtf(fmaxf(q, 1e-8f));
  if (kind == KDOT_MMD_LAPLACIAN) {
    const float k = expf(-r);
    g = clamped ? 0.f : -k * inv_blur * inv_blur / r;
    return k;
  }
  g = clamped ? 0.f : -1.0f / r;
  return -r;
}

__global__ void __launch_bounds__(kMmdThreads) kdot_mmd_kernel(MmdParams p) {
  const SinkhornParams& b = p.b;
  const int img = blockIdx.x, B = b.B, D = p.D;
  const int n0 = b.cu_n[img], N = b.cu_n[img + 1] - n0;
  const int m0 = b.cu_m[img], M = b.cu_m[img + 1] - m0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  __shared__ double s_part[kMmdThreads / 32];
  __shared__ double s_slot;

  if (b.normalize) {  // D == 2 (checked on the host): xy[:,0] /= w ; xy[:,1] /= h in place (loss_libs.py:8-12)
    for (int t = threadIdx.x; t < (N + M) * B; t += blockDim.x) {
      const int q = t / B, slot = t - q * B;
      const bool stu = q < N;
      float* base = stu ? b.xs : b.xt;
      const long long g = stu ? (long long)(n0 + q) * b.s_cell_n + (long long)slot * b.s_slot_n
                              : (long long)(m0 + q - N) * b.s_cell_m + (long long)slot * b.s_slot_m;
      base[2 * g] = __fdiv_rn(base[2 * g], b.w);
      base[2 * g + 1] = __fdiv_rn(base[2 * g + 1], b.h);
    }
    __syncthreads();
  }
  const bool skipped = (N == 0 || M == 0);
  // energy works on raw coordinates; gaussian / laplacian on coordinates divided by blur
  const float inv_blur = p.kind == KDOT_MMD_ENERGY ? 1.0f : 1.0f / p.blur;
  double img_loss = 0.0;
  for (int slot = 0; slot < B; ++slot) {
    double acc = 0.0;
    for (int r = threadIdx.x; r < N + M && !skipped; r += blockDim.x) {
      const bool stu = r < N;
      const long long gi = stu ? (long long)(n0 + r) * b.s_cell_n + (long long)slot * b.s_slot_n
                               : (long long)(m0 + r - N) * b.s_cell_m + (long long)slot * b.s_slot_m;
      const float* rbase = stu ? b.xs : b.xt;
      const float* rw = stu ? b.ws : b.wt;
      float xi[kMmdMaxD];
      // scaled coordinates are formed by an explicit (non-contracted) multiply on both sides of every difference: an FMA
      // xi - x_k * inv_blur would leave the rounding residual of x_i * inv_blur on the diagonal (k == i), and the Gaussian's
      // -1/blur^2 turns that into a spurious 1e-4 gradient
      for (int d = 0; d < D; ++d) xi[d] = __fmul_rn(rbase[(size_t)D * gi + d], inv_blur);
      const float wi = rw ? rw[gi] : __fdiv_rn(1.0f, (float)(stu ? N : M));
      float same = 0.f, cross = 0.f;  // (K alpha)_i over the own cloud, (K beta)_i over the other cloud
      float gs[kMmdMaxD], gc[kMmdMaxD];
      for (int d = 0; d < D; ++d) { gs[d] = 0.f; gc[d] = 0.f; }
      // own cloud
      {
        const int cnt = stu ? N : M, c0 = stu ? n0 : m0;
        const long long sc = stu ? b.s_cell_n : b.s_cell_m, ss = stu ? b.s_slot_n : b.s_slot_m;
        for (int k = 0; k < cnt; ++k) {
          const long long gk = (long long)(c0 + k) * sc + (long long)slot * ss;
          float q = 0.f, df[kMmdMaxD];
          for (int d = 0; d < D; ++d) { df[d] = __fsub_rn(xi[d], __fmul_rn(rbase[(size_t)D * gk + d], inv_blur)); q = fmaf(df[d], df[d], q); }
          const float wk = rw ? __ldg(rw + gk) : __fdiv_rn(1.0f, (float)cnt);
          float g;
          const float kv = mmd_kernel(p.kind, q, inv_blur, g);
          same = fmaf(kv, wk, same);
          if (stu) for (int d = 0; d < D; ++d) gs[d] = fmaf(g * wk, df[d], gs[d]);
        }
      }
      if (stu) {  // student rows also need the cross term against the teacher
        for (int j = 0; j < M; ++j) {
          const long long gj = (long long)(m0 + j) * b.s_cell_m + (long long)slot * b.s_slot_m;
          float q = 0.f, df[kMmdMaxD];
          for (int d = 0; d < D; ++d) { df[d] = __fsub_rn(xi[d], __fmul_rn(b.xt[(size_t)D * gj + d], inv_blur)); q = fmaf(df[d], df[d], q); }
          const float wj = b.wt ? __ldg(b.wt + gj) : __fdiv_rn(1.0f, (float)M);
          float g;
          const float kv = mmd_kernel(p.kind, q, inv_blur, g);
          cross = fmaf(kv, wj, cross);
          for (int d = 0; d < D; ++d) gc[d] = fmaf(g * wj, df[d], gc[d]);
        }
        acc += (double)wi * (0.5 * (double)same - (double)cross);
        // d/dx_i: the 1/blur of the scaled coordinates is already inside g for gaussian / laplacian (df is scaled)
        const float unscale = p.kind == KDOT_MMD_ENERGY ? 1.0f : p.blur;
        for (int d = 0; d < D; ++d) {
          float gv = wi * (gs[d] - gc[d]) * unscale;
          if (b.normalize) gv = __fdiv_rn(gv, d == 0 ? b.w : b.h);
          b.grad_xs[(size_t)D * gi + d] = gv;
        }
        if (b.grad_ws) b.grad_ws[gi] = same - cross;
      } else {
        acc += 0.5 * (double)wi * (double)same;
      }
    }
    if (skipped) {
      for (int r = threadIdx.x; r < N; r += blockDim.x) {
        const long long gi = (long long)(n0 + r) * b.s_cell_n + (long long)slot * b.s_slot_n;
        for (int d = 0; d < D; ++d) b.grad_xs[(size_t)D * gi + d] = 0.f;
        if (b.grad_ws) b.grad_ws[gi] = 0.f;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int wi = 0; wi < nwarps; ++wi) t += s_part[wi];
      s_slot = t;
      if (b.loss_per_slot) b.loss_per_slot[(size_t)img * B + slot] = (float)t;
    }
    __syncthreads();
    img_loss += s_slot;
  }
  if (threadIdx.x == 0) {
    b.loss_per_img[img] = (float)img_loss;
    b.valid[img] = skipped ? KDOT_IMG_SKIPPED : KDOT_IMG_OK;
    if (b.nits_per_img) b.nits_per_img[img] = 0;
  }
}

cudaError_t launch_mmd(const SinkhornParams& prm, int D, int kind, float blur, cudaStream_t stream) {
  MmdParams p;
  p.b = prm;
  p.D = D;
  p.kind = kind;
  p.blur = blur;
  kdot_mmd_kernel<<<prm.nimg, kMmdThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace kdot
