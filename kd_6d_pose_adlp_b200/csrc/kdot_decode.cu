// Student-side prologue / epilogue of the KD loss:
//
//   forward   gather the 16 key-point offsets of every positive cell straight from the per-level head outputs
//             (nimg, C*16, H, W) at (image, level, cell, class), decode them (offset * anchor size + anchor centre,
//             then the inverse of the 2x3 crop affine) and write the (npos, 8, 2) pixel key-points the loss consumes;
//   backward  the transpose: d/d(key-points) -> d/d(offsets), scattered into the (zero-filled) per-level gradients.
//
// Replaces, in the reference, `permute/reshape/cat` of all of pred_reg (`losses/loss.py:62-96`, ~83 MB at batch 64),
// `pred_reg_flatten[pos_inds]` (`kd_loss.py:156`), `pred.view(n,-1,16)[arange, cls]` (`kd_loss.py:47`),
// `TargetCoder.decode` (`models/model.py:144-166`) and `view(-1,2,8).transpose(1,2)` (`kd_loss.py:50`), and their
// autograd backward (index_put into an 83 MB zero tensor, cat / permute backward).
//
// One thread per (positive cell, offset channel k = 0..15); the x / y halves of a key-point meet through one shuffle.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/kdot.h"

namespace kdot {

constexpr int kDecMaxLevels = 8;
constexpr int kDecThreads = 256;

struct DecodeParams {
  const float* reg[kDecMaxLevels];  // forward: head outputs; backward: unused
  float* greg[kDecMaxLevels];       // backward: per-level gradients (zero-filled by the caller)
  int hw[kDecMaxLevels];
  int off[kDecMaxLevels + 1];  // prefix sums of hw
  int nlvl, nimg, C, npos;
  const int64_t* pos_inds;   // [npos] flat cell index = img * cells + off[level] + h * W + w
  const int64_t* cls_label;  // [npos] class of the cell (0-based)
  const float* anchors;      // [npos][4] xyxy
  const float* bbox_trans;   // [npos][2][3] or null
  float* xy;                 // forward out [npos][8][2]
  const float* g_xy;         // backward in  [npos][8][2]
};

struct CellRef {
  int lvl;
  size_t base;   // element offset of channel (cls * 16 + 0) at this cell inside level `lvl`
  size_t chan;   // channel stride (= H * W)
  float aw, ah, acx, acy;
  float i00, i01, i10, i11, t0, t1;  // inverse of the crop affine's linear part, and its offset
  bool affine;
};

__device__ __forceinline__ CellRef locate(const DecodeParams& p, int pos) {
  CellRef c;
  const int cells = p.off[p.nlvl];
  const long long idx = p.pos_inds[pos];
  const int img = (int)(idx / cells);
  const int rem = (int)(idx - (long long)img * cells);
  int l = 0;
  while (l + 1 < p.nlvl && rem >= p.off[l + 1]) ++l;
  c.lvl = l;
  c.chan = (size_t)p.hw[l];
  c.base = ((size_t)img * p.C * 16 + (size_t)p.cls_label[pos] * 16) * c.chan + (size_t)(rem - p.off[l]);
  const float4 a = *reinterpret_cast<const float4*>(p.anchors + (size_t)pos * 4);
  c.aw = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f);
  c.ah = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
  c.acx = __fmul_rn(__fadd_rn(a.z, a.x), 0.5f);
  c.acy = __fmul_rn(__fadd_rn(a.w, a.y), 0.5f);
  c.affine = p.bbox_trans != nullptr;
  if (c.affine) {
    const float* b = p.bbox_trans + (size_t)pos * 6;
    const float a00 = b[0], a01 = b[1], a10 = b[3], a11 = b[4];
    const float det = __fsub_rn(__fmul_rn(a00, a11), __fmul_rn(a01, a10));
    c.i00 = __fdiv_rn(a11, det);
    c.i01 = __fdiv_rn(-a01, det);
    c.i10 = __fdiv_rn(-a10, det);
    c.i11 = __fdiv_rn(a00, det);
    c.t0 = b[2];
    c.t1 = b[5];
  }
  return c;
}

__global__ void __launch_bounds__(kDecThreads) kdot_gather_decode_fwd_kernel(DecodeParams p) {
  const int t = blockIdx.x * kDecThreads + threadIdx.x;
  const int pos = t >> 4, k = t & 15;
  const bool live = pos < p.npos;
  float v = 0.0f;
  CellRef c;
  if (live) {
    c = locate(p, pos);
    const float off = __ldg(p.reg[c.lvl] + c.base + (size_t)k * c.chan);
    v = (k < 8) ? __fadd_rn(__fmul_rn(off, c.aw), c.acx) : __fadd_rn(__fmul_rn(off, c.ah), c.acy);
  }
  const float other = __shfl_xor_sync(0xffffffffu, v, 8);  // x <-> y of the same key-point
  if (!live) return;
  float out = v;
  if (c.affine) {
    const float dx = __fsub_rn((k < 8) ? v : other, c.t0);
    const float dy = __fsub_rn((k < 8) ? other : v, c.t1);
    out = (k < 8) ? __fadd_rn(__fmul_rn(c.i00, dx), __fmul_rn(c.i01, dy))
                  : __fadd_rn(__fmul_rn(c.i10, dx), __fmul_rn(c.i11, dy));
  }
  p.xy[(size_t)pos * 16 + (size_t)(k & 7) * 2 + (k >> 3)] = out;
}

__global__ void __launch_bounds__(kDecThreads) kdot_gather_decode_bwd_kernel(DecodeParams p) {
  const int t = blockIdx.x * kDecThreads + threadIdx.x;
  const int pos = t >> 4, k = t & 15;
  if (pos >= p.npos) return;
  const CellRef c = locate(p, pos);
  const float2 g = *reinterpret_cast<const float2*>(p.g_xy + (size_t)pos * 16 + (size_t)(k & 7) * 2);
  float gx = g.x, gy = g.y;
  if (c.affine) {  // transpose of the inverse linear map
    const float tx = __fadd_rn(__fmul_rn(c.i00, g.x), __fmul_rn(c.i10, g.y));
    const float ty = __fadd_rn(__fmul_rn(c.i01, g.x), __fmul_rn(c.i11, g.y));
    gx = tx;
    gy = ty;
  }
  p.greg[c.lvl][c.base + (size_t)k * c.chan] = (k < 8) ? __fmul_rn(gx, c.aw) : __fmul_rn(gy, c.ah);
}

void count_launches(unsigned n);

static int fill_params(DecodeParams& p, const int32_t* hw_lvl, int nlvl, int nimg, int C, const int64_t* pos_inds,
                       const int64_t* cls_label, const float* anchors, const float* bbox_trans, int npos) {
  if (!hw_lvl || !pos_inds || !cls_label || !anchors) return KDOT_E_BADARG;
  if (nlvl <= 0 || nlvl > kDecMaxLevels || nimg <= 0 || C <= 0 || npos < 0) return KDOT_E_BADARG;
  if (npos > (1 << 26)) return KDOT_E_TOOLARGE;
  memset(&p, 0, sizeof(p));
  int off = 0;
  for (int l = 0; l < nlvl; ++l) {
    if (hw_lvl[l] <= 0) return KDOT_E_BADARG;
    p.hw[l] = hw_lvl[l];
    p.off[l] = off;
    off += hw_lvl[l];
  }
  p.off[nlvl] = off;
  p.nlvl = nlvl; p.nimg = nimg; p.C = C; p.npos = npos;
  p.pos_inds = pos_inds; p.cls_label = cls_label; p.anchors = anchors; p.bbox_trans = bbox_trans;
  return KDOT_OK;
}

}  // namespace kdot

using namespace kdot;

extern "C" int kdot_gather_decode_fwd(const float* const* reg_lvl, const int32_t* hw_lvl, int nlvl, int nimg, int C,
                                      const int64_t* pos_inds, const int64_t* cls_label, const float* anchors,
                                      const float* bbox_trans, int npos, float* xy, void* cuda_stream) {
  if (npos == 0) return KDOT_OK;
  DecodeParams p;
  const int rc = fill_params(p, hw_lvl, nlvl, nimg, C, pos_inds, cls_label, anchors, bbox_trans, npos);
  if (rc != KDOT_OK) return rc;
  if (!reg_lvl || !xy) return KDOT_E_BADARG;
  for (int l = 0; l < nlvl; ++l) {
    if (!reg_lvl[l]) return KDOT_E_BADARG;
    p.reg[l] = reg_lvl[l];
  }
  p.xy = xy;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return KDOT_E_NODEVICE;
  const int blocks = (npos * 16 + kDecThreads - 1) / kDecThreads;
  kdot_gather_decode_fwd_kernel<<<blocks, kDecThreads, 0, (cudaStream_t)cuda_stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}

extern "C" int kdot_gather_decode_bwd(const float* g_xy, const int32_t* hw_lvl, int nlvl, int nimg, int C,
                                      const int64_t* pos_inds, const int64_t* cls_label, const float* anchors,
                                      const float* bbox_trans, int npos, float* const* g_reg_lvl, void* cuda_stream) {
  if (npos == 0) return KDOT_OK;
  DecodeParams p;
  const int rc = fill_params(p, hw_lvl, nlvl, nimg, C, pos_inds, cls_label, anchors, bbox_trans, npos);
  if (rc != KDOT_OK) return rc;
  if (!g_reg_lvl || !g_xy) return KDOT_E_BADARG;
  for (int l = 0; l < nlvl; ++l) {
    if (!g_reg_lvl[l]) return KDOT_E_BADARG;
    p.greg[l] = g_reg_lvl[l];
  }
  p.g_xy = g_xy;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return KDOT_E_NODEVICE;
  const int blocks = (npos * 16 + kDecThreads - 1) / kDecThreads;
  kdot_gather_decode_bwd_kernel<<<blocks, kDecThreads, 0, (cudaStream_t)cuda_stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return KDOT_E_CUDA;
  count_launches(1);
  return KDOT_OK;
}
