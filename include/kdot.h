/*
 * kdot.h -- C ABI of libkdot.so: the B200-native optimal-transport knowledge-distillation path.
 *
 * This is the drop-in boundary.  The reference has no FFI: its boundary is the
 * Python call chain  KDPoseLoss.__call__ -> KDObjectSpaceLoss -> kd_loss_2d -> geomloss.SamplesLoss
 * (/root/reference/losses/kd_loss.py:111,40,86 ; losses/loss_libs.py:1,47) and
 * PostProcessorKD.forward (/root/reference/postprocess/postprocess_kd.py:61).  Each entry point below
 * names the reference code it replaces.  Signatures use plain pointers and sizes only (no torch types);
 * the Python host side (kd_6d_pose_adlp_b200/) binds them with ctypes.
 *
 * Conventions
 *   - all device pointers are caller-owned, contiguous, 16-byte aligned fp32 / int32 buffers;
 *   - every call is asynchronous on `cuda_stream` (a cudaStream_t passed as void*), allocates nothing
 *     and returns 0 on success or a KDOT_E_* code (kdot_last_error() gives the text);
 *   - B = OT slots per cell (8 keypoints for WDRNet+), D = dims of one local prediction (2).
 */
#ifndef KDOT_H_
#define KDOT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDOT_VERSION 100

enum {
  KDOT_OK = 0,
  KDOT_E_BADARG = 1,      /* NULL pointer, negative size, unsupported p / D / layout            */
  KDOT_E_TOOLARGE = 2,    /* problem does not fit the kernel's shared-memory plan                */
  KDOT_E_WORKSPACE = 3,   /* workspace smaller than kdot_workspace_bytes()                        */
  KDOT_E_CUDA = 4,        /* launch / runtime failure                                             */
  KDOT_E_NODEVICE = 5     /* no CUDA device -- there is NO CPU fallback                           */
};

/* Layout of the cell arrays. */
enum {
  KDOT_LAYOUT_CELL_MAJOR = 0, /* xs[cell][slot][D], ws[cell][slot]: the reference's flat layout,
                                 losses/kd_loss.py:50,83 (pred_xy.view(-1,8,2), s_cls (sumN,8))   */
  KDOT_LAYOUT_SLOT_MAJOR = 1  /* xs[slot][cell][D], ws[slot][cell]: SamplesLoss' (B,N,D)/(B,N)
                                 batch layout, losses/loss_libs.py:41-47; requires nimg == 1      */
};

/* per-image status written to valid[] */
enum {
  KDOT_IMG_SKIPPED = 0,     /* N == 0 or M == 0: skipped like losses/loss_libs.py:25-28          */
  KDOT_IMG_OK = 1,
  KDOT_IMG_DEGENERATE = -1, /* zero diameter: geomloss' epsilon_schedule would raise; loss = NaN */
  KDOT_IMG_TOO_MANY_ROUNDS = -2 /* eps schedule longer than KDOT_MAX_ROUNDS; loss = NaN           */
};

#define KDOT_MAX_ROUNDS 1024

/*
 * Fused optimal-transport distillation loss, forward + analytic backward, whole mini-batch.
 *
 * Replaces, per image i with N_i = cu_n[i+1]-cu_n[i] student and M_i teacher cells:
 *   - the in-place normalisation  xy[:,0] /= w ; xy[:,1] /= h        (losses/loss_libs.py:8-12)
 *     when normalize != 0 (requires D == 2); xs AND xt are overwritten with the normalised values
 *     (normalize == 2: normalise in registers WITHOUT the write-back -- small fused path only; used by the
 *     host-buffer entry point when it lets the kernel read the caller's staging area directly);
 *   - the per-image split / transposes / SamplesLoss(...).sum() loop   (losses/loss_libs.py:22-50);
 *   - geomloss.SamplesLoss("sinkhorn", p, blur, scaling, reach) on the tensorized backend, i.e. the
 *     four cost matrices, max_diameter, the float64 epsilon schedule, the symmetrised log-domain
 *     Sinkhorn loop, the last extrapolation and the debiased (un)balanced cost (geomloss 0.2.4;
 *     constructed at losses/kd_loss.py:26-30) -- none of the N x M matrices is materialised;
 *   - autograd's backward of all of the above w.r.t. the student points and masses.
 *
 * ws / wt may be NULL: uniform 1/N_i, 1/M_i masses (losses/loss_libs.py:49, --weightedOT false).
 * reach < 0 selects balanced OT (geomloss reach=None).  p == 2 (cost |x-y|^2/2) or p == 1 (cost |x-y|, D == 2 only).
 *
 * Outputs
 *   loss_per_img [nimg]    sum over the B slots of the Sinkhorn divergence (the `.sum()` of
 *                          loss_libs.py:47); 0 for skipped images
 *   loss_per_slot[nimg*B]  optional (may be NULL): the (B,) vector SamplesLoss itself returns
 *   valid        [nimg]    KDOT_IMG_* status
 *   grad_xs, grad_ws       d(loss_per_img)/d(xs), d/d(ws), same layout as xs / ws; the xs gradient
 *                          is w.r.t. the UN-normalised input when normalize != 0 (includes 1/w, 1/h)
 *                          (grad_ws is written even when ws == NULL; it may be NULL)
 *   nits_per_img [nimg]    optional: len(eps_s) of geomloss' schedule (parity / debugging)
 * The caller applies the reduction over images (losses/kd_loss.py:99-101 divides the sum of the
 * non-skipped images by their count).
 */
int kdot_sinkhorn_fwd_bwd(float* xs, const float* ws, float* xt, const float* wt,
                          const int32_t* cu_n, const int32_t* cu_m, int nimg, int B, int D,
                          int max_n, int max_m, int layout,
                          float p, float blur, float reach, float scaling,
                          float w, float h, int normalize,
                          float* loss_per_img, float* loss_per_slot, int32_t* valid,
                          float* grad_xs, float* grad_ws, int32_t* nits_per_img,
                          void* workspace, size_t workspace_bytes, void* cuda_stream);

/*
 * Kernel-MMD sample losses, fused forward + backward: SamplesLoss("gaussian" | "laplacian" | "energy", blur)
 * (--gtype of arguments/argument_kd.py:41 -> losses/kd_loss.py:26-30; geomloss kernel_tensorized):
 *   loss = 1/2 <a, K_xx a> + 1/2 <b, K_yy b> - <a, K_xy b>,  kind 0: exp(-|x-y|^2 / 2 blur^2), 1: exp(-|x-y| / blur),
 *   2: -|x-y|.  Same layouts, normalisation, skip rule and outputs as kdot_sinkhorn_fwd_bwd (no workspace, no rounds).
 */
int kdot_kernel_mmd_fwd_bwd(float* xs, const float* ws, float* xt, const float* wt,
                            const int32_t* cu_n, const int32_t* cu_m, int nimg, int B, int D,
                            int max_n, int max_m, int layout, int kind, float blur,
                            float w, float h, int normalize,
                            float* loss_per_img, float* loss_per_slot, int32_t* valid,
                            float* grad_xs, float* grad_ws, void* cuda_stream);

/* Bytes of device scratch kdot_sinkhorn_fwd_bwd needs for these bounds (0 is a valid answer). */
size_t kdot_workspace_bytes(int nimg, int max_n, int max_m, int B, int D);
/* same, for a given p (p == 1 always runs on the streaming kernel and needs its scratch even for small clouds) */
size_t kdot_workspace_bytes_ex(int nimg, int max_n, int max_m, int B, int D, float p);

/*
 * Host-buffer convenience around kdot_sinkhorn_fwd_bwd for callers that hold NumPy / C arrays
 * (what `bench.py` times as the end-to-end number): packs the inputs into a pinned staging area,
 * one H2D copy, the fused kernel, one D2H copy, stream synchronise, unpack.  All pointers are HOST
 * pointers; pos_n / pos_m are the per-image cell counts (pos_per_img, pos_per_img_t of
 * losses/kd_loss.py:74-76).  xs_h / xt_h are overwritten with the normalised values only when
 * write_back_normalized != 0.  The context owns device + pinned memory sized at creation.
 */
typedef struct kdot_host_ctx kdot_host_ctx;
kdot_host_ctx* kdot_host_ctx_create(int device, int max_img, int max_cells_s, int max_cells_t, int B, int D);
void kdot_host_ctx_destroy(kdot_host_ctx* ctx);
int kdot_sinkhorn_fwd_bwd_host(kdot_host_ctx* ctx, float* xs_h, const float* ws_h, float* xt_h,
                               const float* wt_h, const int32_t* pos_n, const int32_t* pos_m, int nimg,
                               float p, float blur, float reach, float scaling, float w, float h,
                               int normalize, int write_back_normalized,
                               float* loss_per_img_h, int32_t* valid_h, float* grad_xs_h,
                               float* grad_ws_h, int32_t* nits_h);
/* host wall-clock (microseconds) of the last host call's phases: pack, enqueue, synchronise, unpack */
void kdot_host_ctx_last_timing(const kdot_host_ctx* ctx, double* us4);
/* bytes moved by the last host call (for bench.py's e2e accounting) */
void kdot_host_ctx_last_traffic(const kdot_host_ctx* ctx, size_t* h2d_bytes, size_t* d2h_bytes);

/*
 * Teacher knowledge extraction: segmentation-weighted cell voting / selection, all images and all
 * classes in one launch.  Replaces the index-producing part of
 *   PostProcessorKD.forward_for_single_feature_map  (postprocess/postprocess_kd.py:22-59)
 *   PostProcessorKD.pose_infer_ml                   (postprocess/postprocess_kd.py:99-156)
 * i.e. sigmoid > th candidates, per-level arg-max, box size of the running-best cell, the per-level
 * cell budget nk = int(positive_num * softmax-like(-lambda * log2(size/anchor)^2) + 0.5), and the
 * per-level top-min(valid, nk) by score.  RANSAC-EPnP (postprocess_kd.py:191) stays on the host.
 *
 * Inputs (device): cls_lvl[l] -> (nimg, C, H_l, W_l) logits, reg_lvl[l] -> (nimg, C*16, H_l, W_l);
 * hw_lvl[l] = H_l * W_l and stride_lvl[l], anchor size per level (host arrays); anchor_sizes_all
 * is the FULL cfg list (its sum normalises nk even when nlvl is shorter: postprocess_kd.py:143).
 * Outputs (device), per (image, class) pair q = img * C + cls, capacity `cap` cells:
 *   sel_count[q]                number of selected cells (0 when the class has no candidate)
 *   sel_level[q*cap + k], sel_loc[q*cap + k]   level and h*W+w index of the k-th selected cell,
 *                               level by level, descending score inside a level (topk order)
 *   sel_score[q*cap + k]        sqrt(sigmoid(logit))  (postprocess_kd.py:57)
 *   sel_kpts [(q*cap + k)*16]   decoded keypoints [x0..x7, y0..y7] in crop pixels (model.py:144-154)
 *   nk[q*nsizes + l]            per-level budget; valid_cnt[q*nlvl + l] candidates per level
 *   best[q*2 + {0,1}]           level / loc of the highest-confidence cell
 */
int kdot_select_cells(const float* const* cls_lvl, const float* const* reg_lvl, const int32_t* hw_lvl,
                      const int32_t* w_lvl, const float* stride_lvl, int nlvl,
                      const float* anchor_sizes_all, int nsizes,
                      int nimg, int C, float th, int positive_num, float positive_lambda, int cap,
                      int32_t* sel_count, int32_t* sel_level, int32_t* sel_loc, float* sel_score,
                      float* sel_kpts, int32_t* nk, int32_t* valid_cnt, int32_t* best,
                      void* cuda_stream);

/*
 * Student-side prologue / epilogue of the loss: gather the 16 key-point
 * offsets of every positive cell straight from the per-level head outputs and decode them to full-image
 * pixels, and the transposed scatter for the backward pass.  Replaces
 *   flatten of pred_reg                    (losses/loss.py:62-96; permute + reshape + cat of ~83 MB at batch 64)
 *   pred_reg_flatten[pos_inds]             (losses/kd_loss.py:156)
 *   pred.view(n,-1,16)[arange(n), cls]     (losses/kd_loss.py:47)
 *   TargetCoder.decode                     (models/model.py:144-166: offset * anchor size + centre, inverse crop affine)
 *   .view(-1,2,8).transpose(1,2)           (losses/kd_loss.py:50)
 * and their autograd backward.
 *
 * reg_lvl[l] -> (nimg, C*16, H_l, W_l) fp32 device tensors (host array of device pointers), hw_lvl[l] = H_l*W_l
 * (host); pos_inds[npos] int64 flat cell indices img*cells + level_offset + h*W + w (the reference's label
 * order); cls_label[npos] int64 0-based class per cell; anchors[npos][4] xyxy of those cells;
 * bbox_trans[npos][2][3] crop affines or NULL.  Forward writes xy[npos][8][2] (key-point k: x, y).
 * Backward takes g_xy[npos][8][2] and writes d/d(offset) at the same 16 addresses of g_reg_lvl[l], which the
 * caller has zero-filled (pos_inds are unique, so there are no atomics).
 */
int kdot_gather_decode_fwd(const float* const* reg_lvl, const int32_t* hw_lvl, int nlvl, int nimg, int C,
                           const int64_t* pos_inds, const int64_t* cls_label, const float* anchors,
                           const float* bbox_trans, int npos, float* xy, void* cuda_stream);
int kdot_gather_decode_bwd(const float* g_xy, const int32_t* hw_lvl, int nlvl, int nimg, int C,
                           const int64_t* pos_inds, const int64_t* cls_label, const float* anchors,
                           const float* bbox_trans, int npos, float* const* g_reg_lvl, void* cuda_stream);

/*
 * The two dense losses beside the OT term in KDPoseLoss.__call__, forward + gradient fused.
 *
 * kdot_focal_loss_fwd_bwd: SigmoidFocalLoss.forward(pred_cls_flatten[valid], labels[valid]) of losses/loss.py:12-40
 * (called at losses/kd_loss.py:134) without flattening anything: cls_lvl[l] -> (nimg, C, H_l, W_l) logits (host array
 * of device pointers), labels[nimg * cells] int64 in the reference's label order (-1 ignored, 0 background, c + 1
 * positive of class c; losses/loss.py:246-252).  Writes loss[0] = sum over all non-ignored (cell, class) pairs and,
 * when g_cls_lvl != NULL, d loss / d logit into per-level gradients of the same shape (0 at ignored cells).
 * workspace: kdot_focal_workspace_bytes() bytes whose first 16 are ZERO on entry (the kernel leaves them zero).
 * The sum is reduced in a fixed order: bit-reproducible run to run.
 *
 * kdot_reg3d_loss_fwd_bwd: the 3-D object-space regression loss of losses/kd_loss.py:57-71 on the decoded key-points:
 * xy[npos*8][2] (kd_loss.py:50), target3d[npos][8][3] (aux_3D_in_camera_frame of the positive cells), diam_cell[npos]
 * (MESH_DIAMETERS[class]), kinv9_host = inverse(INTERNAL_K) row-major (HOST array).  Writes loss_cell[npos]
 * (the per-cell `losses` of kd_loss.py:69-70; the reference returns their sum) and g_xy = d(sum)/d(xy).
 */
int kdot_focal_loss_fwd_bwd(const float* const* cls_lvl, const int32_t* hw_lvl, int nlvl, int nimg, int C,
                            const int64_t* labels, float gamma, float alpha, float* loss, float* const* g_cls_lvl,
                            void* workspace, size_t workspace_bytes, void* cuda_stream);
size_t kdot_focal_workspace_bytes(void);
int kdot_reg3d_loss_fwd_bwd(const float* xy, const float* target3d, const float* diam_cell, const float* kinv9_host,
                            int npos, float* loss_cell, float* g_xy, void* cuda_stream);

/*
 * Device-side SSC positive sampling: the label assignment of PoseLossDzi.prepare_targets (losses/loss.py:164-268,
 * POSITIVE_TYPE == 'SSC').  All images in one launch per stage.
 *
 * kdot_ssc_count: mask[nimg][mh][mw] object-index maps (0 background, g + 1 object g: PoseAnnot.mask), anchors[cells][4]
 * of ONE image (xyxy, level-major), hw_lvl / size_lvl host arrays, per-image objects padded to maxgt (<= 8):
 * num_gt[nimg], rot[nimg][maxgt][3][3], trans[nimg][maxgt][3], kp3d[nimg][maxgt][8][3] (key-points of the object's class),
 * K[nimg][3][3], bbox_trans[nimg][2][3] or NULL.  Outputs: gtid[nimg][cells] (owner of every anchor centre, loss.py:194-203),
 * span[nimg][maxgt] (box_span of the reprojected 3-D box, libs/poses.py:280-300 + libs/boxlist.py:229-233),
 * nk[nimg][nlvl][maxgt] (per-level budget, loss.py:207-215), count[nimg][nlvl][maxgt] (candidates, loss.py:224).
 *
 * kdot_ssc_pick: picks[nimg][nlvl][maxgt][cap] = a uniform draw without replacement of min(nk, count) ordinals in
 * [0, count) from a counter-based generator keyed by `seed` (-1 pads) -- the device replacement of loss.py:227's CPU
 * torch.randperm.  For bit-exact parity with a reference run the caller fills `picks` on the host instead
 * (torch.randperm(count)[:k] in the reference's image / level / object order).
 *
 * kdot_ssc_assign: labels[nimg][cells] int64 (class + 1 drawn, -1 in mask but not drawn, 0 background: loss.py:236-252),
 * owner[nimg][cells] (anchors_to_gt_indexs), npos[nimg].  cls_plus1[nimg][maxgt] = class_ids + 1.
 */
int kdot_ssc_count(const float* mask, int mh, int mw, const float* anchors, const int32_t* hw_lvl, const float* size_lvl,
                   int nlvl, int nimg, int maxgt, const int32_t* num_gt, const float* rot, const float* trans,
                   const float* kp3d, const float* K, const float* bbox_trans, int positive_num, float positive_lambda,
                   uint8_t* gtid, int32_t* count, int32_t* nk, float* span, void* cuda_stream);
int kdot_ssc_pick(const int32_t* count, const int32_t* nk, int nimg, int nlvl, int maxgt, int cap, uint64_t seed,
                  int32_t* picks, void* cuda_stream);
int kdot_ssc_assign(const uint8_t* gtid, const int32_t* picks, const int64_t* cls_plus1, const int32_t* hw_lvl, int nlvl,
                    int nimg, int maxgt, int cap, int64_t* labels, int32_t* owner, int32_t* npos, void* cuda_stream);

/* Diagnostics */
const char* kdot_last_error(void);
int kdot_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
unsigned long long kdot_launch_count(void);
/* profiling aid: when non-NULL, CTA 0-thread [nimg][16] int64 SM-clock stamps (0-6), globaltimer (7) and exact-fallback count (8) are written by the small kernel */
void kdot_debug_set_clock_buffer(void* dev_ptr);
/* FP32 FMA-chain micro-benchmark: returns measured TFLOP/s of the device (roofline denominator) */
double kdot_measure_fp32_peak_tflops(int device, int iters);

#ifdef __cplusplus
}
#endif
#endif /* KDOT_H_ */
