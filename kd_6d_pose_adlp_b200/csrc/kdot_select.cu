// Teacher knowledge extraction: segmentation-weighted cell selection, all images x classes in ONE launch.
//
// Replaces the index-producing part of the reference's PostProcessorKD
//   forward_for_single_feature_map  (postprocess/postprocess_kd.py:22-59)  sigmoid > th candidates, decode
//   pose_infer_ml                   (postprocess/postprocess_kd.py:99-156) per-level arg-max, running-best box
//                                   size, per-level budget nk, per-level top-min(valid, nk)
// which the reference runs as Python loops over images x levels x labels with a host sync per step.
//
// One CTA per (image, class): the class plane of every FPN level is contiguous in the NCHW head output, so the
// CTA streams  sum_l H_l*W_l  logits once with coalesced loads into shared memory (HBM-bound single pass),
// then does the tiny arg-max / top-k work on chip.  Ordering key is the logit (sigmoid and sqrt are monotone),
// ties broken towards the lower cell index; floating-point outputs (scores, key-points) follow the reference's
// op order (mul then add, no FMA contraction) so that they are reproducible against it.
#include <cstring>

#include "kdot_common.cuh"

namespace kdot {

constexpr int kSelMaxLevels = 8;
constexpr int kSelThreads = 128;
constexpr int kSelMaxSizes = 8;

struct SelectParams {
  const float* cls[kSelMaxLevels];
  const float* reg[kSelMaxLevels];
  int hw[kSelMaxLevels];
  int wd[kSelMaxLevels];
  int off[kSelMaxLevels + 1];  // prefix sums of hw
  float stride[kSelMaxLevels];
  float sizes[kSelMaxSizes];
  int nlvl, nsizes, nimg, C, positive_num, cap;
  float th, lambda;
  int32_t* sel_count; int32_t* sel_level; int32_t* sel_loc; float* sel_score; float* sel_kpts;
  int32_t* nk; int32_t* valid_cnt; int32_t* best;
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// block-wide arg-max of (value, index) with ties to the lower index; result broadcast to all threads
__device__ __forceinline__ void block_argmax(float& v, int& idx, float* s_v, int* s_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { s_v[warp] = v; s_i[warp] = idx; }
  __syncthreads();
  v = s_v[0]; idx = s_i[0];
#pragma unroll
  for (int w = 1; w < kSelThreads / 32; ++w) {
    const float ov = s_v[w];
    const int oi = s_i[w];
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

// decode the 16 offsets of cell `loc` on level l -> [x0..x7, y0..y7] crop pixels (models/model.py:144-154)
__device__ __forceinline__ void decode_cell(const SelectParams& p, int img, int cls, int l, int loc, float* out16) {
  const int hw = p.hw[l], wd = p.wd[l];
  const float size = p.sizes[l], stride = p.stride[l];
  const int hy = loc / wd, wx = loc - hy * wd;
  const float acx = __fadd_rn(__fmul_rn((float)wx, stride), 0.5f * stride);
  const float acy = __fadd_rn(__fmul_rn((float)hy, stride), 0.5f * stride);
  const float* base = p.reg[l] + ((size_t)img * p.C * 16 + (size_t)cls * 16) * hw + loc;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    out16[k] = __fadd_rn(__fmul_rn(base[(size_t)k * hw], size), acx);
    out16[8 + k] = __fadd_rn(__fmul_rn(base[(size_t)(8 + k) * hw], size), acy);
  }
}

// same, one offset channel per thread (threads 0..15 of the CTA): the 16 strided loads of a cell are in flight together
__device__ __forceinline__ void decode_cell_par(const SelectParams& p, int img, int cls, int l, int loc, float* out16, int k) {
  const int hw = p.hw[l], wd = p.wd[l];
  const float size = p.sizes[l], stride = p.stride[l];
  const int hy = loc / wd, wx = loc - hy * wd;
  const float ac = __fadd_rn(__fmul_rn((float)(k < 8 ? wx : hy), stride), 0.5f * stride);
  const float* base = p.reg[l] + ((size_t)img * p.C * 16 + (size_t)cls * 16) * hw + loc;
  out16[k] = __fadd_rn(__fmul_rn(base[(size_t)k * hw], size), ac);
}

__global__ void __launch_bounds__(kSelThreads) kdot_select_kernel(SelectParams p) {
  const int q = blockIdx.x;
  const int img = q / p.C, cls = q - img * p.C;
  extern __shared__ float s_logit[];  // all levels of this (image, class) plane; non-candidates -> -big
  __shared__ float s_v[kSelThreads / 32];
  __shared__ int s_i[kSelThreads / 32];
  __shared__ int s_cnt[kSelMaxLevels];
  __shared__ float s_best16[16];
  __shared__ int s_pick[64];  // (level << 24) | cell of every selected cell (cap <= 64)
  const float neg = -3.0e38f;

  if (threadIdx.x < kSelMaxLevels) s_cnt[threadIdx.x] = 0;
  __syncthreads();

  // ---- pass 1: stream the plane, threshold (postprocess_kd.py:35), count per level ----
  for (int l = 0; l < p.nlvl; ++l) {
    const float* plane = p.cls[l] + ((size_t)img * p.C + cls) * p.hw[l];
    int local = 0;
    for (int c = threadIdx.x; c < p.hw[l]; c += kSelThreads) {
      const float x = plane[c];
      const bool cand = sigmoidf_ref(x) > p.th;
      s_logit[p.off[l] + c] = cand ? x : neg;
      local += cand ? 1 : 0;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt[l], local);
  }
  __syncthreads();
  int total = 0;
  for (int l = 0; l < p.nlvl; ++l) total += s_cnt[l];
  if (threadIdx.x < p.nlvl) p.valid_cnt[(size_t)q * p.nlvl + threadIdx.x] = s_cnt[threadIdx.x];
  if (total == 0) {  // class absent from this image (postprocess_kd.py:104-109)
    if (threadIdx.x == 0) {
      p.sel_count[q] = 0;
      p.best[2 * q] = -1; p.best[2 * q + 1] = -1;
    }
    if (threadIdx.x < p.nsizes) p.nk[(size_t)q * p.nsizes + threadIdx.x] = 0;
    return;
  }

  // ---- per-level arg-max, running best confidence / box size (postprocess_kd.py:123-141) ----
  float box_conf = 0.f, box_size = 0.f;
  int best_l = -1, best_loc = -1;
  for (int l = 0; l < p.nlvl; ++l) {
    if (s_cnt[l] == 0) continue;  // uniform
    float v = neg;
    int idx = 0x7fffffff;
    for (int c = threadIdx.x; c < p.hw[l]; c += kSelThreads) {
      const float x = s_logit[p.off[l] + c];
      if (x > v) { v = x; idx = c; }
    }
    block_argmax(v, idx, s_v, s_i);
    const float score = sqrtf(sigmoidf_ref(v));
    if (score > box_conf) {  // uniform: every thread holds the same (v, idx)
      box_conf = score;
      best_l = l; best_loc = idx;
      __syncthreads();
      if (threadIdx.x < 16) decode_cell_par(p, img, cls, l, idx, s_best16, threadIdx.x);
      __syncthreads();
      float xmin = s_best16[0], xmax = s_best16[0], ymin = s_best16[8], ymax = s_best16[8];
#pragma unroll
      for (int k = 1; k < 8; ++k) {
        xmin = fminf(xmin, s_best16[k]); xmax = fmaxf(xmax, s_best16[k]);
        ymin = fminf(ymin, s_best16[8 + k]); ymax = fmaxf(ymax, s_best16[8 + k]);
      }
      const float sz = fmaxf(__fsub_rn(xmax, xmin), __fsub_rn(ymax, ymin));
      if (sz > box_size) box_size = sz;
    }
  }

  // ---- per-level budget nk (postprocess_kd.py:143-146), fp32 in the reference's op order ----
  int nk_l[kSelMaxSizes];
  {
    float e[kSelMaxSizes];
    float sum = 0.f;
    for (int s = 0; s < p.nsizes; ++s) {
      const float dk = log2f(__fdiv_rn(box_size, p.sizes[s]));
      e[s] = expf(__fmul_rn(-p.lambda, __fmul_rn(dk, dk)));
      sum = __fadd_rn(sum, e[s]);
    }
    for (int s = 0; s < p.nsizes; ++s) {
      const float v = __fdiv_rn(__fmul_rn((float)p.positive_num, e[s]), sum);
      nk_l[s] = (int)__fadd_rn(v, 0.5f);
      if (threadIdx.x == 0) p.nk[(size_t)q * p.nsizes + s] = nk_l[s];
    }
  }

  // ---- per-level top-min(valid, nk) by score, descending (postprocess_kd.py:151-156) ----
  int count = 0;
  for (int l = 0; l < p.nlvl; ++l) {
    const int want = min(s_cnt[l], l < p.nsizes ? nk_l[l] : 0);
    for (int t = 0; t < want; ++t) {
      float v = neg;
      int idx = 0x7fffffff;
      for (int c = threadIdx.x; c < p.hw[l]; c += kSelThreads) {
        const float x = s_logit[p.off[l] + c];
        if (x > v) { v = x; idx = c; }
      }
      block_argmax(v, idx, s_v, s_i);
      if (count < p.cap) {
        const size_t o = (size_t)q * p.cap + count;
        if (threadIdx.x == 0) {
          p.sel_level[o] = l;
          p.sel_loc[o] = idx;
          p.sel_score[o] = sqrtf(sigmoidf_ref(v));
          s_pick[count] = (l << 24) | idx;
          s_logit[p.off[l] + idx] = neg;  // remove from the pool
        }
        ++count;
      }
      __syncthreads();
    }
  }
  // decode of all picked cells at the end: count x 16 independent strided loads spread over the CTA
  for (int t = threadIdx.x; t < count * 16; t += kSelThreads) {
    const int k = t >> 4, ch = t & 15, pk = s_pick[k];
    decode_cell_par(p, img, cls, pk >> 24, pk & 0xffffff, p.sel_kpts + ((size_t)q * p.cap + k) * 16, ch);
  }
  if (threadIdx.x == 0) {
    p.sel_count[q] = count;
    p.best[2 * q] = best_l;
    p.best[2 * q + 1] = best_loc;
  }
}

void count_launches(unsigned n);
}  // namespace kdot

using namespace kdot;

extern "C" int kdot_select_cells(const float* const* cls_lvl, const float* const* reg_lvl, const int32_t* hw_lvl,
                                 const int32_t* w_lvl, const float* stride_lvl, int nlvl,
                                 const float* anchor_sizes_all, int nsizes, int nimg, int C, float th,
                                 int positive_num, float positive_lambda, int cap, int32_t* sel_count,
                                 int32_t* sel_level, int32_t* sel_loc, float* sel_score, float* sel_kpts, int32_t* nk,
                                 int32_t* valid_cnt, int32_t* best, void* cuda_stream) {
  if (nimg == 0) return KDOT_OK;
  if (!cls_lvl || !reg_lvl || !hw_lvl || !w_lvl || !stride_lvl || !anchor_sizes_all || !sel_count || !sel_level ||
      !sel_loc || !sel_score || !sel_kpts || !nk || !valid_cnt || !best)
    return KDOT_E_BADARG;
  if (nlvl <= 0 || nlvl > kSelMaxLevels || nsizes < nlvl || nsizes > kSelMaxSizes || nimg < 0 || C <= 0 || cap <= 0 || cap > 64)
    return KDOT_E_BADARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return KDOT_E_NODEVICE;
  SelectParams p;
  memset(&p, 0, sizeof(p));
  int off = 0;
  for (int l = 0; l < nlvl; ++l) {
    if (!cls_lvl[l] || !reg_lvl[l] || hw_lvl[l] <= 0 || w_lvl[l] <= 0) return KDOT_E_BADARG;
    p.cls[l] = cls_lvl[l]; p.reg[l] = reg_lvl[l];
    p.hw[l] = hw_lvl[l]; p.wd[l] = w_lvl[l]; p.stride[l] = stride_lvl[l];
    p.off[l] = off;
    off += hw_lvl[l];
  }
  p.off[nlvl] = off;
  for (int s = 0; s < nsizes; ++s) p.sizes[s] = anchor_sizes_all[s];
  p.nlvl = nlvl; p.nsizes = nsizes; p.nimg = nimg; p.C = C; p.positive_num = positive_num; p.cap = cap;
  p.th = th; p.lambda = positive_lambda;
  p.sel_count = sel_count; p.sel_level = sel_level; p.sel_loc = sel_loc; p.sel_score = sel_score;
  p.sel_kpts = sel_kpts; p.nk = nk; p.valid_cnt = valid_cnt; p.best = best;
  const size_t smem = (size_t)off * sizeof(float);
  if (smem > 200 * 1024) return KDOT_E_TOOLARGE;
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    if (cudaFuncSetAttribute(kdot_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return KDOT_E_CUDA;
    configured = smem;
  }
  kdot_select_kernel<<<nimg * C, kSelThreads, smem, (cudaStream_t)cuda_stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}
