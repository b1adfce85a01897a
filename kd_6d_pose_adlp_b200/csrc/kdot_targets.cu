// Device-side SSC positive sampling: the label assignment of PoseLossDzi.prepare_targets
// (reference losses/loss.py:164-268, POSITIVE_TYPE == 'SSC'), SURVEY.md section 8(f) item 3.
//
// The reference walks images x levels x ground-truth objects in Python: per object a full-mask `nonzero` (box_span via
// to_object_boxlist, libs/poses.py:264-304), per (level, object) a `nonzero` + `torch.randperm` + `min(tensor, len)` host
// sync (loss.py:218-229), then an `roi.max` over a (cells, num_gt) matrix -- a few hundred launches and ~4 syncs per image.
// Here:
//   kdot_ssc_count   one CTA per image: which object (if any) owns every anchor centre (mask look-up, loss.py:194-203),
//                    the objects' reprojected 2-D boxes and spans (poses.py:280-300, boxlist.py:229-233), the per-level
//                    budget nk = int(P * exp(-lambda * log2(span / size)^2) / sum + 0.5) (loss.py:211-215) and the number
//                    of candidate cells per (level, object);
//   kdot_ssc_pick    (optional) a uniform draw WITHOUT replacement of min(nk, count) candidates per (level, object) from
//                    a counter-based generator -- the device replacement of loss.py:227's CPU `torch.randperm`;
//   kdot_ssc_assign  labels per cell: class + 1 for the drawn cells, -1 for in-mask cells that were not drawn, 0 for
//                    background (loss.py:236-252), plus the owning object of every cell.
// For bit-exact parity with a reference run the draw can instead come from the host: the caller reads `count`, replays
// `torch.randperm(count)[:k]` in the reference's (image, level, object) order on the CPU generator and passes the picks in.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kdot.h"

namespace kdot {

constexpr int kSscMaxLevels = 8;
constexpr int kSscMaxGt = 8;
constexpr int kSscThreads = 256;

struct SscParams {
  const float* mask;        // [nimg][mh][mw] object-index map: 0 background, g + 1 object g
  int mh, mw;
  const float* anchors;     // [cells][4] xyxy of ONE image's anchors (identical for every image)
  int hw[kSscMaxLevels];
  int off[kSscMaxLevels + 1];
  float size[kSscMaxLevels];   // anchor size per level
  int nlvl, nimg, maxgt;
  const int32_t* num_gt;    // [nimg]
  const float* rot;         // [nimg][maxgt][3][3]
  const float* trans;       // [nimg][maxgt][3]
  const float* kp3d;        // [nimg][maxgt][8][3]  key-points of each object's class
  const float* K;           // [nimg][3][3]
  const float* bbox_trans;  // [nimg][2][3] or null
  int positive_num;
  float positive_lambda;
  uint8_t* gtid;            // out [nimg][cells]: 0 none, g + 1
  int32_t* count;           // out [nimg][nlvl][maxgt]
  int32_t* nk;              // out [nimg][nlvl][maxgt]
  float* span;              // out [nimg][maxgt]
};

__global__ void __launch_bounds__(kSscThreads) kdot_ssc_count_kernel(SscParams p) {
  const int img = blockIdx.x;
  const int G = min(p.num_gt[img], p.maxgt);
  const int cells = p.off[p.nlvl];
  __shared__ int s_has[kSscMaxGt];
  __shared__ float s_span[kSscMaxGt];
  __shared__ int s_cnt[kSscMaxLevels][kSscMaxGt];
  for (int t = threadIdx.x; t < kSscMaxGt; t += blockDim.x) s_has[t] = 0;
  for (int t = threadIdx.x; t < kSscMaxLevels * kSscMaxGt; t += blockDim.x) s_cnt[t / kSscMaxGt][t % kSscMaxGt] = 0;
  __syncthreads();
  // objects with at least one mask pixel (poses.py:270-278: otherwise the box is [0, 0, 0, 0])
  const float* m = p.mask + (size_t)img * p.mh * p.mw;
  unsigned int seen = 0u;
  for (int t = threadIdx.x; t < p.mh * p.mw; t += blockDim.x) {
    const int g = (int)m[t];
    if (g >= 1 && g <= G && (float)g == m[t]) seen |= 1u << (g - 1);
  }
  for (int o = 16; o > 0; o >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, o);
  if ((threadIdx.x & 31) == 0 && seen) {
    for (int g = 0; g < G; ++g) if (seen >> g & 1u) atomicOr(&s_has[g], 1);
  }
  __syncthreads();
  // reprojected boxes: one thread per object (8 key-points each)
  if ((int)threadIdx.x < G) {
    const int g = threadIdx.x;
    float span = 1.0f;  // box [0,0,0,0]: max(0 - 0 + 1, 0 - 0 + 1)
    if (s_has[g]) {
      const float* R = p.rot + ((size_t)img * p.maxgt + g) * 9;
      const float* T = p.trans + ((size_t)img * p.maxgt + g) * 3;
      const float* P3 = p.kp3d + ((size_t)img * p.maxgt + g) * 24;
      const float* K = p.K + (size_t)img * 9;
      float x0 = 3.0e38f, y0 = 3.0e38f, x1 = -3.0e38f, y1 = -3.0e38f;
      for (int k = 0; k < 8; ++k) {
        float c[3];
        for (int r = 0; r < 3; ++r) c[r] = R[3 * r] * P3[3 * k] + R[3 * r + 1] * P3[3 * k + 1] + R[3 * r + 2] * P3[3 * k + 2] + T[r];
        const float u = K[0] * c[0] + K[1] * c[1] + K[2] * c[2], v = K[3] * c[0] + K[4] * c[1] + K[5] * c[2];
        const float w = K[6] * c[0] + K[7] * c[1] + K[8] * c[2];
        float x = u / (w + 1e-8f), y = v / (w + 1e-8f);
        if (p.bbox_trans) {
          const float* A = p.bbox_trans + (size_t)img * 6;
          const float xx = A[0] * x + A[1] * y + A[2], yy = A[3] * x + A[4] * y + A[5];
          x = xx; y = yy;
        }
        x0 = fminf(x0, x); x1 = fmaxf(x1, x); y0 = fminf(y0, y); y1 = fmaxf(y1, y);
      }
      span = fmaxf(x1 - x0 + 1.0f, y1 - y0 + 1.0f);
    }
    s_span[g] = span;
    p.span[(size_t)img * p.maxgt + g] = span;
  }
  // owner of every anchor centre + candidates per (level, object)
  for (int l = 0; l < p.nlvl; ++l) {
    for (int c = threadIdx.x; c < p.hw[l]; c += blockDim.x) {
      const float4 a = *reinterpret_cast<const float4*>(p.anchors + (size_t)(p.off[l] + c) * 4);
      // loss.py:194-198: centre, clamped to the mask, truncated towards zero by .long()
      const float cx = fminf(fmaxf((a.z + a.x) / 2.0f, 0.f), (float)(p.mw - 1));
      const float cy = fminf(fmaxf((a.w + a.y) / 2.0f, 0.f), (float)(p.mh - 1));
      const float mv = m[(size_t)(long long)cy * p.mw + (long long)cx];
      int g = 0;
      for (int q = 1; q <= G; ++q) if (mv == (float)q) g = q;
      p.gtid[(size_t)img * cells + p.off[l] + c] = (uint8_t)g;
      if (g) atomicAdd(&s_cnt[l][g - 1], 1);
    }
  }
  __syncthreads();
  // per-level budget (loss.py:207-215): fp32, the op order of the reference
  if ((int)threadIdx.x < G) {
    const int g = threadIdx.x;
    float w[kSscMaxLevels], sum = 0.f;
    for (int l = 0; l < p.nlvl; ++l) {
      const float dk = fabsf(log2f(s_span[g] / p.size[l]));
      w[l] = expf(-p.positive_lambda * (dk * dk));
      sum += w[l];
    }
    for (int l = 0; l < p.nlvl; ++l) {
      const float v = (float)p.positive_num * w[l] / sum;
      p.nk[((size_t)img * p.nlvl + l) * p.maxgt + g] = (int)(v + 0.5f);
      p.count[((size_t)img * p.nlvl + l) * p.maxgt + g] = s_cnt[l][g];
    }
  }
  for (int t = threadIdx.x; t < p.nlvl * p.maxgt; t += blockDim.x) {
    if (t % p.maxgt >= G) {
      p.nk[(size_t)img * p.nlvl * p.maxgt + t] = 0;
      p.count[(size_t)img * p.nlvl * p.maxgt + t] = 0;
    }
  }
}

// counter-based generator: 64-bit mix (splitmix64 finaliser) of (seed, image, level, object, draw)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// one thread per (image, level, object): Floyd's algorithm -- a uniformly random k-subset of {0 .. count-1}
__global__ void kdot_ssc_pick_kernel(const int32_t* count, const int32_t* nk, int n, int cap, unsigned long long seed,
                                     int32_t* picks) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int cnt = count[q], k = min(min(nk[q], cnt), cap);
  int32_t* out = picks + (size_t)q * cap;
  int m = 0;
  for (int j = cnt - k; j < cnt; ++j) {
    const unsigned long long r = mix64(seed ^ mix64(((unsigned long long)q << 20) + (unsigned long long)j + 0x9e3779b97f4a7c15ull));
    int t = (int)(r % (unsigned long long)(j + 1));
    bool dup = false;
    for (int i = 0; i < m; ++i) dup |= out[i] == t;
    out[m++] = dup ? j : t;
  }
  for (; m < cap; ++m) out[m] = -1;
}

struct AssignParams {
  const uint8_t* gtid;      // [nimg][cells]
  const int32_t* picks;     // [nimg][nlvl][maxgt][cap] ordinals among the (level, object) candidates in ascending cell order; -1 pads
  const int64_t* cls_plus1; // [nimg][maxgt] class id + 1 of every object
  int hw[kSscMaxLevels];
  int off[kSscMaxLevels + 1];
  int nlvl, nimg, maxgt, cap;
  int64_t* labels;          // out [nimg][cells]
  int32_t* owner;           // out [nimg][cells]: object index of a positive cell, 0 elsewhere (anchors_to_gt_indexs)
  int32_t* npos;            // out [nimg]
};

__global__ void __launch_bounds__(kSscThreads) kdot_ssc_assign_kernel(AssignParams p) {
  const int img = blockIdx.x;
  const int cells = p.off[p.nlvl];
  const uint8_t* gt = p.gtid + (size_t)img * cells;
  __shared__ int s_warp[kSscThreads / 32];
  __shared__ int s_base[kSscMaxGt];
  __shared__ int s_npos;
  if (threadIdx.x == 0) s_npos = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int l = 0; l < p.nlvl; ++l) {
    if (threadIdx.x < kSscMaxGt) s_base[threadIdx.x] = 0;
    __syncthreads();
    for (int c0 = 0; c0 < p.hw[l]; c0 += blockDim.x) {
      const int c = c0 + threadIdx.x;
      const int g = c < p.hw[l] ? gt[p.off[l] + c] : 0;
      int64_t label = 0;
      int owner = 0;
      // ordinal of this cell among its object's candidates of the level: block-wide exclusive count, object by object
      for (int q = 1; q <= p.maxgt; ++q) {
        const unsigned int bal = __ballot_sync(0xffffffffu, g == q);
        if (!__syncthreads_or(bal != 0u)) continue;   // uniform: nobody in the block belongs to object q in this chunk
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_base[q - 1];
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const int ord = before + __popc(bal & ((1u << lane) - 1u));
        if (g == q) {
          const int32_t* pk = p.picks + (((size_t)img * p.nlvl + l) * p.maxgt + (q - 1)) * p.cap;
          bool hit = false;
          for (int i = 0; i < p.cap; ++i) hit |= pk[i] == ord;
          label = hit ? p.cls_plus1[(size_t)img * p.maxgt + q - 1] : -1;   // in the mask but not drawn: ignored
          owner = hit ? q - 1 : 0;
          if (hit) atomicAdd(&s_npos, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          int tot = 0;
          for (int w = 0; w < kSscThreads / 32; ++w) tot += s_warp[w];
          s_base[q - 1] += tot;
        }
        __syncthreads();
      }
      if (c < p.hw[l]) {
        p.labels[(size_t)img * cells + p.off[l] + c] = label;
        p.owner[(size_t)img * cells + p.off[l] + c] = owner;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) p.npos[img] = s_npos;
}

void count_launches(unsigned n);
}  // namespace kdot

using namespace kdot;

extern "C" {

int kdot_ssc_count(const float* mask, int mh, int mw, const float* anchors, const int32_t* hw_lvl, const float* size_lvl,
                   int nlvl, int nimg, int maxgt, const int32_t* num_gt, const float* rot, const float* trans,
                   const float* kp3d, const float* K, const float* bbox_trans, int positive_num, float positive_lambda,
                   uint8_t* gtid, int32_t* count, int32_t* nk, float* span, void* cuda_stream) {
  if (nimg <= 0 || nlvl <= 0 || nlvl > kSscMaxLevels || maxgt <= 0 || maxgt > kSscMaxGt || !mask || !anchors || !hw_lvl ||
      !size_lvl || !num_gt || !rot || !trans || !kp3d || !K || !gtid || !count || !nk || !span)
    return KDOT_E_BADARG;
  SscParams p;
  p.mask = mask; p.mh = mh; p.mw = mw; p.anchors = anchors; p.nlvl = nlvl; p.nimg = nimg; p.maxgt = maxgt;
  p.off[0] = 0;
  for (int l = 0; l < nlvl; ++l) { p.hw[l] = hw_lvl[l]; p.off[l + 1] = p.off[l] + hw_lvl[l]; p.size[l] = size_lvl[l]; }
  p.num_gt = num_gt; p.rot = rot; p.trans = trans; p.kp3d = kp3d; p.K = K; p.bbox_trans = bbox_trans;
  p.positive_num = positive_num; p.positive_lambda = positive_lambda;
  p.gtid = gtid; p.count = count; p.nk = nk; p.span = span;
  kdot_ssc_count_kernel<<<nimg, kSscThreads, 0, (cudaStream_t)cuda_stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}

int kdot_ssc_pick(const int32_t* count, const int32_t* nk, int nimg, int nlvl, int maxgt, int cap, uint64_t seed,
                  int32_t* picks, void* cuda_stream) {
  if (nimg <= 0 || !count || !nk || !picks || cap <= 0) return KDOT_E_BADARG;
  const int n = nimg * nlvl * maxgt;
  kdot_ssc_pick_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(count, nk, n, cap, (unsigned long long)seed, picks);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}

int kdot_ssc_assign(const uint8_t* gtid, const int32_t* picks, const int64_t* cls_plus1, const int32_t* hw_lvl, int nlvl,
                    int nimg, int maxgt, int cap, int64_t* labels, int32_t* owner, int32_t* npos, void* cuda_stream) {
  if (nimg <= 0 || nlvl <= 0 || nlvl > kSscMaxLevels || maxgt <= 0 || maxgt > kSscMaxGt || !gtid || !picks || !cls_plus1 ||
      !hw_lvl || !labels || !owner || !npos)
    return KDOT_E_BADARG;
  AssignParams p;
  p.gtid = gtid; p.picks = picks; p.cls_plus1 = cls_plus1; p.nlvl = nlvl; p.nimg = nimg; p.maxgt = maxgt; p.cap = cap;
  p.off[0] = 0;
  for (int l = 0; l < nlvl; ++l) { p.hw[l] = hw_lvl[l]; p.off[l + 1] = p.off[l] + hw_lvl[l]; }
  p.labels = labels; p.owner = owner; p.npos = npos;
  kdot_ssc_assign_kernel<<<nimg, kSscThreads, 0, (cudaStream_t)cuda_stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}

}  // extern "C"
