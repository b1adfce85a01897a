// Fused OT distillation loss, streaming kernel: any cloud size, any feature dimension D in {1,2,3,4,8,16}.
//
// Same reference code as the other Sinkhorn kernels (losses/loss_libs.py:8-12,22-50 + geomloss' tensorized
// Sinkhorn divergence + autograd backward), for problems that do not fit one SM's shared memory and for the dense
// ZebraPose-style 16-D configuration (BASELINE.json configs[3]).  ONE cooperative launch for the whole mini-batch:
//
//   phase 0a  per image: in-place normalisation, bounding box, float64 epsilon schedule
//   phase 0b  stage every (image, slot) problem once into an L2-resident SoA scratch  pts[prob][d][P], lw[prob][P],
//             h^S/h^C[prob][2][P]  (32-point tiles, pads carry h = -big); D = 2 clouds in Morton order with per-tile
//             bounding boxes, which the cold rounds use to skip -- exactly -- tiles whose exponentials all flush to 0
//   rounds    dataflow, no barrier between rounds: one global FIFO of WARP units (round, problem, row cloud,
//             32*R-row block, column set); a unit of round r starts when every unit of round r-1 of ITS problem has
//             finished (per-problem counters).  A warp streams the unit's columns through a private double-buffered
//             cp.async tile and evaluates the pairs like the tiled kernel: direct differences, log2-domain soft-min
//             argument, lazily re-based reference exponent, one ex2 per pair, packed f32x2 math.
//   final     fixed-order fp64 reduction of the per-row loss terms (bit-reproducible).
//
// Potentials never leave the chip's L2 between rounds; the N x M cost matrix is never materialised.
#include <cooperative_groups.h>

#include "kdot_common.cuh"

namespace cg = cooperative_groups;

namespace kdot {

#ifndef KDOT_STREAM_THREADS
#define KDOT_STREAM_THREADS 256
#endif
// tuning (tools/bench_variants.sh, dense_b32 on B200): D <= 2 runs best with 2 rows per lane at 64 registers and
// 32 resident warps per SM (8.72 ms) -- more warps hide the SFU / FMA pipe contention better than more rows per lane
// (4 rows, 124 registers, 16 warps: 9.44 ms)
#ifndef KDOT_STREAM_MINBLOCKS
#define KDOT_STREAM_MINBLOCKS 4
#endif
#ifndef KDOT_STREAM_R2
#define KDOT_STREAM_R2 2
#endif
constexpr int kStreamThreads = KDOT_STREAM_THREADS;
constexpr float kTauS = 80.f;  // re-base when a chunk sum exceeds 2^80 of the reference: rare, and fp32 keeps full precision below 2^127

struct StreamParams {
  SinkhornParams b;
  int D;
  int strideP;   // padded points per problem (NqMax + MqMax)
  int nqMax;     // offset of the teacher block inside a problem
  int nbx, nby;  // row blocks per cloud (upper bound over the batch), block = 32 * R rows
  int upp;       // units per problem per round = 2 * (nbx + nby)
  float* pts;    // [nprob][D][strideP]
  float* lw2;    // [nprob][strideP]
  double* pot;   // [nprob][2][strideP]          S, C  (float64: kdot_common.cuh precision plan)
  float* h;      // [nprob][2 buffers][2][strideP]   fp32 head of h (what the fp32 sweeps and the skip test read)
  float* hlo;    // [nprob][2 buffers][2][strideP]   h - (float)h: float64 h = h + hlo for the high-precision sub-tiles
  int* perm;     // [nprob][strideP]             staged position -> cell index inside its cloud (-1 for pads)
  float4* tbox;  // [nprob][strideP/32]          bounding box (min x, min y, max x, max y) of every 32-point tile (D = 2)
  float* hmax;   // [nprob][2 buffers][2][strideP/32]  max of h over the tile, maintained by the units that write h
  unsigned short* jb;  // [nprob][2][strideP]  (value = sub-tile index = column / 32)          per row and column set (own, cross): first column of the chunk that held
                 //                              the row's running max in the previous round (seeds the next sweep)
  double* pref;       // [nprob][3][4]            reference potentials the published h is centred on (see stream_unit), slot r % 3
  unsigned int* hmag; // [nprob][3]               max |h| (fp32 bits) of the offsets consumed in round r: slot r % 3; gates the float64 path
  unsigned int* ctr;  // [KDOT_MAX_ROUNDS]       first 8 bytes: 64-bit head of the global unit FIFO
  unsigned int* done; // [nprob]                 finished units per problem (all rounds)
};

template <int D, int R, bool GRAD>
struct URow {
  float nx[D];          // -p_i
  float mref;           // reference exponent of the running sums (may be stale by up to ~kTauS)
  float2 mu;            // mref / (-coef), duplicated: folded into the squared-distance FMA chain
  float2 s;             // partial sums of the CURRENT 32-column sub-tile (two interleaved accumulators)
  float tot, comp;      // compensated (Kahan) total of the finished sub-tiles, same scale; row sum = tot - comp
  float2 g[GRAD ? D : 1];
  // cold rounds (SKIP): the 32-column sub-tile that added the most to the running sum so far.  Its maximum is within a
  // factor 32 (5 log2 units) of the row maximum, so it seeds the next round's sweep (urow_seed) with a reference that
  // needs no re-basing and puts every far tile below the skip threshold from the first column on.
  int jb;
  float gbest;
};

// Row sums.  A row of a dense cloud adds up thousands of exponentials of comparable size in the early rounds; two
// running fp32 accumulators lose ~sqrt(n) 2^-24 there (1e-6 at n = 1360), the potentials inherit eps_r times that, and
// the last round divides what survives by eps_final = 1e-6: emulated in numpy, sequential fp32 row sums alone put 4e-3
// into the final offsets and 4e-5..3e-4 into d/dx of the dense 1360 x 1364 problem, a compensated sum over 32-column
// sub-tiles 2e-6 (float64 pair arguments and exact exponentials changed nothing -- DESIGN.md section 3).  So the fp32
// accumulators only ever span one sub-tile (16 terms each) and are folded into a Kahan pair: 4 FADD per row and sub-tile.
template <int D, int R, bool GRAD>
__device__ __forceinline__ void urow_fold(URow<D, R, GRAD>& st) {
  const float y = __fsub_rn(__fadd_rn(st.s.x, st.s.y), st.comp);
  const float t = __fadd_rn(st.tot, y);
  st.comp = __fsub_rn(__fsub_rn(t, st.tot), y);
  st.tot = t;
  st.s = make_float2(0.f, 0.f);
}
// log2 of the row sum tot - comp
template <int D, int R, bool GRAD>
__device__ __forceinline__ double urow_lg2_sum(const URow<D, R, GRAD>& st) {
  return lg2_sum_exact(st.tot) - (st.tot > 0.f ? (double)__fdiv_rn(st.comp, st.tot) * 1.4426950408889634 : 0.0);
}

__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// One 4-column chunk (X[d] = coordinates, H = soft-min offsets) against the R rows of this lane.
// Running sums are kept relative to a possibly stale reference exponent mref; a chunk whose partial sum exceeds
// 2^kTauS (or is +inf: first chunk, mref = -big) takes the cold path that re-bases on the exact max (see
// kdot_tiled.cu for the full argument).
// Hot-path arithmetic per pair: the reference exponent is folded into the first FMA of the squared distance,
//   a = v - mref = coef * (mu + dx^2 + dy^2 + ...) + h   with   mu = mref / (-coef)   (row constant),
// so a pair costs D FADD + D FFMA (distance) + 1 FFMA (scale + offset): no separate subtraction of the reference.
// A rounding error in mu shifts every exponent of the row by the same amount and cancels in the soft-min.
// FOLD is used only while the exponents are small (early, warm rounds: eps >= eps_0 / 256, |h| < ~200): there the
// two extra roundings it introduces are < 2e-5.  In the cold rounds (|h| ~ 1e3..1e4) the unfolded form
// (one rounding of v, then an exact subtraction) is kept, because those rounds set the accuracy of the result.
// P1: cost |x - y| instead of |x - y|^2 / 2 (geomloss p = 1: sqrt(max(|d|^2, 1e-8)); the clamp also zeroes the
// gradient there).  Never combined with FOLD.
__device__ __forceinline__ float2 p1_norm(float2 q) {
  return make_float2(sqrtf(fmaxf(q.x, 1e-8f)), sqrtf(fmaxf(q.y, 1e-8f)));
}
__device__ __forceinline__ float2 p1_grad_weight(float2 p, float2 q) {  // p / r, 0 where the clamp is active
  return make_float2(q.x > 1e-8f ? p.x * rsqrtf(q.x) : 0.f, q.y > 1e-8f ? p.y * rsqrtf(q.y) : 0.f);
}

template <int D, int R, bool GRAD, bool FOLD, bool P1>
__device__ __forceinline__ void stream_chunk(URow<D, R, GRAD> (&st)[R], const float4 (&X)[D], const float4& H,
                                             const float2 coef2, const float inv_ncoef, const float big) {
  static_assert(!(FOLD && P1), "the folded form is only defined for the squared cost");
  float2 ps[R], p0s[GRAD ? R : 1], p1s[GRAD ? R : 1];
  bool rebase = false;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    float2 q0 = st[k].mu, q1 = st[k].mu;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float2 nd = make_float2(st[k].nx[d], st[k].nx[d]);
      const float2 a0 = __fadd2_rn(make_float2(X[d].x, X[d].y), nd);
      const float2 a1 = __fadd2_rn(make_float2(X[d].z, X[d].w), nd);
      q0 = (!FOLD && d == 0) ? __fmul2_rn(a0, a0) : __ffma2_rn(a0, a0, q0);
      q1 = (!FOLD && d == 0) ? __fmul2_rn(a1, a1) : __ffma2_rn(a1, a1, q1);
    }
    const float2 sq0 = q0, sq1 = q1;
    if (P1) { q0 = p1_norm(q0); q1 = p1_norm(q1); }
    float2 e0 = __ffma2_rn(q0, coef2, make_float2(H.x, H.y));
    float2 e1 = __ffma2_rn(q1, coef2, make_float2(H.z, H.w));
    if (!FOLD) {
      const float2 nm = make_float2(-st[k].mref, -st[k].mref);
      e0 = __fadd2_rn(e0, nm);
      e1 = __fadd2_rn(e1, nm);
    }
    const float2 p0 = make_float2(ex2_approx(e0.x), ex2_approx(e0.y));
    const float2 p1 = make_float2(ex2_approx(e1.x), ex2_approx(e1.y));
    ps[k] = __fadd2_rn(p0, p1);
    if (GRAD) { p0s[k] = P1 ? p1_grad_weight(p0, sq0) : p0; p1s[k] = P1 ? p1_grad_weight(p1, sq1) : p1; }
    rebase |= !(ps[k].x + ps[k].y <= big);
  }
  if (rebase) {  // cold: exact max of the offending rows' chunk, re-base their running sums
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (ps[k].x + ps[k].y <= big) continue;
      float2 q0 = make_float2(0.f, 0.f), q1 = q0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float2 nd = make_float2(st[k].nx[d], st[k].nx[d]);
        const float2 a0 = __fadd2_rn(make_float2(X[d].x, X[d].y), nd);
        const float2 a1 = __fadd2_rn(make_float2(X[d].z, X[d].w), nd);
        q0 = d == 0 ? __fmul2_rn(a0, a0) : __ffma2_rn(a0, a0, q0);
        q1 = d == 0 ? __fmul2_rn(a1, a1) : __ffma2_rn(a1, a1, q1);
      }
      const float2 sq0 = q0, sq1 = q1;
      if (P1) { q0 = p1_norm(q0); q1 = p1_norm(q1); }
      const float2 v0 = __ffma2_rn(q0, coef2, make_float2(H.x, H.y));
      const float2 v1 = __ffma2_rn(q1, coef2, make_float2(H.z, H.w));
      const float vm = fmaxf(fmaxf(v0.x, v0.y), fmaxf(v1.x, v1.y));
      const float sc = ex2_approx(st[k].mref - vm);  // 0 for the first chunk (mref = -big)
      st[k].s.x *= sc; st[k].s.y *= sc;
      if (GRAD) {
#pragma unroll
        for (int d = 0; d < D; ++d) { st[k].g[d].x *= sc; st[k].g[d].y *= sc; }
      }
      st[k].mref = vm;
      st[k].gbest *= sc;   // the sub-tile bookkeeping of stream_rows lives on the same scale as the sums
      st[k].tot *= sc; st[k].comp *= sc;
      const float mu = vm * inv_ncoef;
      st[k].mu = make_float2(mu, mu);
      const float2 p0 = make_float2(ex2_approx(v0.x - vm), ex2_approx(v0.y - vm));
      const float2 p1 = make_float2(ex2_approx(v1.x - vm), ex2_approx(v1.y - vm));
      ps[k] = __fadd2_rn(p0, p1);
      if (GRAD) { p0s[k] = P1 ? p1_grad_weight(p0, sq0) : p0; p1s[k] = P1 ? p1_grad_weight(p1, sq1) : p1; }
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    st[k].s = __fadd2_rn(st[k].s, ps[k]);
    if (GRAD) {  // sum_j e_ij (p_j - p_i): differences recomputed (last round only) to keep registers flat in D
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float2 nd = make_float2(st[k].nx[d], st[k].nx[d]);
        st[k].g[d] = __ffma2_rn(p0s[k], __fadd2_rn(make_float2(X[d].x, X[d].y), nd), st[k].g[d]);
        st[k].g[d] = __ffma2_rn(p1s[k], __fadd2_rn(make_float2(X[d].z, X[d].w), nd), st[k].g[d]);
      }
    }
  }
}

// Speculative form of stream_chunk for the potential-only sweeps: no per-chunk overflow test, the sums are checked once
// per 32-column sub-tile by the caller (which re-runs the sub-tile through stream_chunk when a sum left the safe range).
template <int D, int R, bool FOLD>
__device__ __forceinline__ void stream_chunk_spec(URow<D, R, false> (&st)[R], const float4 (&X)[D], const float4& H,
                                                  const float2 coef2) {
#pragma unroll
  for (int k = 0; k < R; ++k) {
    float2 q0 = st[k].mu, q1 = st[k].mu;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float2 nd = make_float2(st[k].nx[d], st[k].nx[d]);
      const float2 a0 = __fadd2_rn(make_float2(X[d].x, X[d].y), nd);
      const float2 a1 = __fadd2_rn(make_float2(X[d].z, X[d].w), nd);
      q0 = (!FOLD && d == 0) ? __fmul2_rn(a0, a0) : __ffma2_rn(a0, a0, q0);
      q1 = (!FOLD && d == 0) ? __fmul2_rn(a1, a1) : __ffma2_rn(a1, a1, q1);
    }
    float2 e0 = __ffma2_rn(q0, coef2, make_float2(H.x, H.y));
    float2 e1 = __ffma2_rn(q1, coef2, make_float2(H.z, H.w));
    if (!FOLD) {
      const float2 nm = make_float2(-st[k].mref, -st[k].mref);
      e0 = __fadd2_rn(e0, nm);
      e1 = __fadd2_rn(e1, nm);
    }
    st[k].s = __fadd2_rn(st[k].s, make_float2(ex2_approx(e0.x), ex2_approx(e0.y)));
    st[k].s = __fadd2_rn(st[k].s, make_float2(ex2_approx(e1.x), ex2_approx(e1.y)));
  }
}

// Column tile staged per warp in shared memory with cp.async (L2 -> smem, no registers, no L1): T columns of the D
// coordinate arrays plus the soft-min offsets, double buffered, so the loads of tile t+1 are in flight during the
// whole evaluation of tile t (thousands of cycles: the L2 latency is fully hidden).
template <int D>
__host__ __device__ constexpr int stream_tile_cols() { return D <= 4 ? 128 : (D <= 8 ? 64 : 32); }
// ... plus a float64 staging area for one 32-column sub-tile (D coordinates + h), see stream_subtile_hi
template <int D>
__host__ __device__ constexpr int stream_warp_smem_floats() { return 2 * (D + 1) * stream_tile_cols<D>() + 2 * (D + 1) * 32 + 8; }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// R rows of this lane against columns [0, ncols) of one column set (ncols multiple of 4; SoA global arrays
// pts[d][strideP], ch[]).  The h arrays are rewritten every round by other SMs: cp.async.cg reads them from L2.
// Tile-level skip data of one column set: bounding boxes and h maxima of its 32-column tiles.  For row i and a tile,
//   hmax_tile + coef * dist^2(p_i, tile box)       (coef < 0)
// bounds every soft-min argument of the row against the tile's columns from above; when it lies more than 140 (plus a
// magnitude-proportional rounding allowance, derived at the test itself) below the row's reference exponent for EVERY
// row of the warp, each ex2 of the tile flushes to exactly +0 in every lane (-126 = ex2.approx.ftz flush point), so the
// tile is skipped without evaluating a single pair: only exact zeros are dropped, the result is bit-identical to
// evaluating everything.  tests/test_sinkhorn_gpu.py::test_tile_skipping_is_bit_exact checks that against the
// -DKDOT_NO_TILE_SKIP build (libkdot_noskip.so: same order, seeds and arithmetic, every tile evaluated) on adversarial
// clouds (clusters + far outliers, masses 1e-3 / 0.999, blur 1e-3..5e-2, N >> M).
struct TileSkip {
  const float4* tbox;  // first tile of the column set: (min x, min y, max x, max y)
  const float* hmax;   // same tiles, current h buffer
};

// ---- high-precision sub-tiles (cold rounds among the last KDOT_HI_ROUNDS, kdot_common.cuh) -------------------------
// In those rounds |h| and |coef d^2| reach 1e3..1e4 log2-units and an fp32 soft-min argument (ulp ~1e-3) is too coarse
// for the pairs that carry the sum.  Only few pairs do: the fp32 evaluation of a 32-column sub-tile is kept as a SCREEN,
// and a sub-tile that adds more than kSigThr of a row's running sum (for any row of the warp) is re-evaluated with the
// argument  h_j + coef |p_i - p_j|^2 - ref  formed in float64 -- float64 copies of its 32 columns are staged in a small
// per-warp buffer (coordinates are exact fp32 -> float64 conversions, h = fp32 head + fp32 tail read from L2).
// What stays fp32 adds up to < 32 * kSigThr per row with ~1e-3 relative error each: < 1e-6 of the sum.
// The gradient round evaluates every sub-tile it does not skip this way (its weights enter d/dx directly).
#ifndef KDOT_SIG_THR
#define KDOT_SIG_THR (1.0f / 16384.0f)
#endif
constexpr float kSigThr = KDOT_SIG_THR;
// Stream-kernel value of kHiMagnitude (kdot_common.cuh).  With the compensated row sums (urow_fold) in place the fp32
// pair arguments are what is left: tools/accuracy_stream_scan.py (16 streaming cases, max-norm of d/dx against float64)
// gives 1.4e-4 / 1.1e-4 / 8.6e-5 at a gate of 4096 / 512 / 64 log2-units of CENTRED offset magnitude -- the maximum is set
// by the wide, sparse clouds (sigma = 0.3) -- for 8.43 / 8.64 / 8.82 ms on dense_b32, whose own max-norm (1.1e-4) is ONE
// knife-edge cell of 21 760 that sits at 1.0e-4..1.1e-4 even with float64 arguments everywhere (99.9 % quantile: 2e-6).
#ifndef KDOT_STREAM_HI_MAG
#define KDOT_STREAM_HI_MAG 4096.0f
#endif
constexpr float kStreamHiMagnitude = KDOT_STREAM_HI_MAG;
// Problems of up to kStreamSmallPoints staged points (both clouds) take the register / CTA-resident kernels' gate instead
// (kHiMagnitude = 64): there the float64 sub-tiles cost next to nothing in absolute terms and bring the wide sparse clouds of
// the scan from 1.4e-4 to <= 8.6e-5, while on dense_b32 (2752 staged points) the lower gate costs 4.6 % for one cell moving
// from 1.14e-4 to 1.01e-4.
#ifndef KDOT_STREAM_SMALL_POINTS
#define KDOT_STREAM_SMALL_POINTS 2048
#endif
constexpr int kStreamSmallPoints = KDOT_STREAM_SMALL_POINTS;
#ifndef KDOT_STREAM_HI_ROUNDS
#define KDOT_STREAM_HI_ROUNDS KDOT_HI_ROUNDS
#endif
constexpr int kStreamHiRounds = KDOT_STREAM_HI_ROUNDS;  // trailing rounds that may take float64 sub-tiles

// Arguments of the high-precision path.  They live in the warp's shared memory (the last 32 bytes of its region, written
// by lane 0 before a sweep), not in registers: the kernel runs at 64 registers per thread for four CTAs per SM, and nine
// registers held across the tile loop for a path that is taken by a few sub-tiles per sweep were spilled -- 89 M local
// loads / stores per dense_b32 launch, 105 M sectors of write-through traffic to L2 (ncu, profiles/r02_prof_stream_*).
struct HiArgs {
  const float* chlo;  // fp32 tail of h for the column set (same indexing as ch)
  double coefd;
  long long* dbg;     // profiling aid (kdot_debug_set_clock_buffer): [5] screened, [6] float64 sub-tiles, [7] gradient-round float64 sub-tiles
  float mag_fac;      // hi_mag_factor of this round: a sub-tile needs float64 arguments iff max|h| * mag_fac > 1
  float pad;
};
static_assert(sizeof(HiArgs) == 32, "HiArgs occupies the 8 trailing floats of a warp's shared-memory region");

template <int D>
__device__ __forceinline__ HiArgs* stream_hi_args(float* wsm) {
  return reinterpret_cast<HiArgs*>(wsm + 2 * (D + 1) * stream_tile_cols<D>() + 2 * (D + 1) * 32);
}
template <int D>
__device__ __forceinline__ double* stream_hi_staging(float* wsm) {  // [D + 1][32] float64
  return reinterpret_cast<double*>(wsm + 2 * (D + 1) * stream_tile_cols<D>());
}
template <int D>
__device__ __forceinline__ void stream_hi_publish(float* wsm, int lane, const float* chlo, double coefd, float mag_fac, long long* dbg) {
  __syncwarp();
  if (lane == 0) {
    HiArgs* ha = stream_hi_args<D>(wsm);
    ha->chlo = chlo; ha->coefd = coefd; ha->dbg = dbg; ha->mag_fac = mag_fac; ha->pad = 0.f;
  }
  __syncwarp();
}

// Staging of one sub-tile for the float64 evaluation.  The argument is evaluated in the EXPANDED form
//   u_ij = h_j + coef |x_j|^2  +  coef |p_i|^2 - ref_i  -  2 coef <x_j, p_i>        (exact enough in float64: 1e-10)
// so a pair costs D + 1 DP operations instead of 2 D + 2: the column part G_j = h_j + coef |x_j|^2 is staged, the row
// part K_i and the vector -2 coef p_i are formed once per sub-tile.  Everything is scaled by kRnd = 1 + 2^-25, which turns
// the truncation of f64_to_f32_trunc_nz-style re-packing into (nearly) round-to-nearest at no cost per pair.
constexpr double kRnd = 1.0 + 1.0 / 33554432.0;

template <int D>
__device__ __forceinline__ void stage_hi(double* hsm, const float* tb, int T, int sb, int lane, const float* chlo_sub,
                                         bool valid, double coefd) {
  __syncwarp();  // previous readers of the staging buffer are done
  double xx = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double x = (double)tb[d * T + sb + lane];
    hsm[d * 32 + lane] = x;
    xx = fma(x, x, xx);
  }
  const float lo = valid ? __ldcg(chlo_sub + lane) : 0.f;
  const double h = (double)tb[D * T + sb + lane] + (double)lo;
  hsm[D * 32 + lane] = fma(coefd, xx, h) * kRnd;
  __syncwarp();
}

// max |h| over the real columns of a staged 32-column sub-tile (pads carry -big)
template <int D>
__device__ __forceinline__ float subtile_hmag(const float* tb, int T, int sb, int n, int lane) {
  float v = 0.f;
  if (sb + lane < n) {
    const float h = tb[D * T + sb + lane];
    v = h > -1.0e29f ? fabsf(h) : 0.f;
  }
  return warp_max(v);
}

// float64 -> fp32 re-packing of an argument difference: integer pipe (exact 0 / tiny |u| -> 0), see kdot_common.cuh
__device__ __forceinline__ float hi_pack(double u) {
  const unsigned int hi = (unsigned int)__double2hiint(u), lo = (unsigned int)__double2loint(u);
  const unsigned int packed = __funnelshift_l(lo, hi, 3) ^ 0x40000000u;
  const unsigned int bits = (packed & 0x7fffffffu) | (hi & 0x80000000u);
  return ((hi & 0x7ff00000u) < (897u << 20)) ? 0.f : __uint_as_float(bits);
}

// The R rows of this lane against the ncol (multiple of 4, <= 32) columns of one sub-tile in float64; same lazily
// re-based running sums as stream_chunk (the reference stays an fp32 number: exactly representable in float64).  Three of
// four conversions go through the integer pipe, the fourth through F2F on the XU pipe (which also serves the ex2).
// Deliberately NOT inlined, state passed and returned by value: the float64 code needs ~2x the registers of the fp32
// sweep, and inlined it cost the fp32 hot loop of the same instantiation 10 % (measured on dense_b32) although it runs
// for a few per cent of the sub-tiles only.
template <int D, int R, bool GRAD>
struct HiIO {
  float nx[R][D];   // -p_i
  float mref[R];    // reference exponent (in / out)
  float2 s[R];      // running sums (in / out)
  float2 g[GRAD ? R * D : 1];  // gradient accumulators (in / out)
  float sc[R];      // out: product of the re-basing factors applied to s (1 when the reference did not move)
};

template <int D, int R, bool GRAD>
__device__ __noinline__ HiIO<D, R, GRAD> subtile_hi(HiIO<D, R, GRAD> io, double* hsm, const float* tb, int T, int sb,
                                                    int ncol, int lane, const float* chlo_sub, bool valid, double coefd,
                                                    float big) {
  stage_hi<D>(hsm, tb, T, sb, lane, chlo_sub, valid, coefd);
  double ax[R][D], K[R];
  const double c2 = -2.0 * coefd * kRnd;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    double pp = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double pd = -(double)io.nx[k][d];
      ax[k][d] = c2 * pd;
      pp = fma(pd, pd, pp);
    }
    K[k] = fma(coefd, pp, -(double)io.mref[k]) * kRnd;
    io.sc[k] = 1.0f;
  }
#pragma unroll 1
  for (int j = 0; j < ncol; j += 4) {
    const double2 G0 = *reinterpret_cast<const double2*>(hsm + D * 32 + j);
    const double2 G1 = *reinterpret_cast<const double2*>(hsm + D * 32 + j + 2);
    double u[R][4];
#pragma unroll
    for (int k = 0; k < R; ++k) { u[k][0] = G0.x + K[k]; u[k][1] = G0.y + K[k]; u[k][2] = G1.x + K[k]; u[k][3] = G1.y + K[k]; }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double2 A = *reinterpret_cast<const double2*>(hsm + d * 32 + j);
      const double2 Bv = *reinterpret_cast<const double2*>(hsm + d * 32 + j + 2);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        u[k][0] = fma(A.x, ax[k][d], u[k][0]); u[k][1] = fma(A.y, ax[k][d], u[k][1]);
        u[k][2] = fma(Bv.x, ax[k][d], u[k][2]); u[k][3] = fma(Bv.y, ax[k][d], u[k][3]);
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      // (correctly rounded exponentials here -- exp2() in float64 -- change no digit of the worst cases: DESIGN.md section 3)
      float p0 = ex2_approx(hi_pack(u[k][0])), p1 = ex2_approx(hi_pack(u[k][1]));
      float p2 = ex2_approx(hi_pack(u[k][2])), p3 = ex2_approx((float)u[k][3]);
      if (!((p0 + p1) + (p2 + p3) <= big)) {  // cold: re-base on the max of this chunk
        const float um = (float)fmax(fmax(u[k][0], u[k][1]), fmax(u[k][2], u[k][3]));  // relative to the old reference
        const float vm = io.mref[k] + um;                                                  // any fp32 number near the max will do
        const float sc = ex2_approx(io.mref[k] - vm);  // 0 for the first chunk (mref = -big)
        const double shift = ((double)vm - (double)io.mref[k]) * kRnd;  // exact difference of two fp32 numbers
        io.s[k].x *= sc; io.s[k].y *= sc;
        if (GRAD) {
#pragma unroll
          for (int d = 0; d < D; ++d) { io.g[k * D + d].x *= sc; io.g[k * D + d].y *= sc; }
        }
        io.mref[k] = vm;
        io.sc[k] *= sc;
        K[k] -= shift;
        p0 = ex2_approx(hi_pack(u[k][0] - shift)); p1 = ex2_approx(hi_pack(u[k][1] - shift));
        p2 = ex2_approx(hi_pack(u[k][2] - shift)); p3 = ex2_approx((float)(u[k][3] - shift));
      }
      io.s[k].x += p0 + p2; io.s[k].y += p1 + p3;
      if (GRAD) {  // weights from the float64 argument, coordinate differences in fp32 (exact enough: they are O(cloud size))
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float4 X = *reinterpret_cast<const float4*>(tb + d * T + sb + j);
          const float nd = io.nx[k][d];
          io.g[k * D + d].x = fmaf(p0, X.x + nd, fmaf(p2, X.z + nd, io.g[k * D + d].x));
          io.g[k * D + d].y = fmaf(p1, X.y + nd, fmaf(p3, X.w + nd, io.g[k * D + d].y));
        }
      }
    }
  }
  return io;
}

// glue: URow state <-> HiIO
template <int D, int R, bool GRAD>
__device__ __forceinline__ void stream_subtile_hi(URow<D, R, GRAD> (&st)[R], double* hsm, const float* tb, int T, int sb,
                                                  int ncol, int lane, const float* chlo_sub, bool valid, double coefd,
                                                  float inv_ncoef, float big) {
  HiIO<D, R, GRAD> io;
#pragma unroll
  for (int k = 0; k < R; ++k) {
#pragma unroll
    for (int d = 0; d < D; ++d) io.nx[k][d] = st[k].nx[d];
    io.mref[k] = st[k].mref;
    io.s[k] = st[k].s;
    if (GRAD) {
#pragma unroll
      for (int d = 0; d < D; ++d) io.g[k * D + d] = st[k].g[d];
    }
  }
  io = subtile_hi<D, R, GRAD>(io, hsm, tb, T, sb, ncol, lane, chlo_sub, valid, coefd, big);
#pragma unroll
  for (int k = 0; k < R; ++k) {
    st[k].s = io.s[k];
    if (GRAD) {
#pragma unroll
      for (int d = 0; d < D; ++d) st[k].g[d] = io.g[k * D + d];
    }
    if (io.mref[k] != st[k].mref) {
      st[k].mref = io.mref[k];
      st[k].gbest *= io.sc[k];
      st[k].tot *= io.sc[k]; st[k].comp *= io.sc[k];
      const float mu = io.mref[k] * inv_ncoef;
      st[k].mu = make_float2(mu, mu);
    }
  }
}

// SEED: cold rounds -- sub-tile bookkeeping that seeds the next round's reference exponent.  TSKIP: D = 2 problems staged
// in Morton order -- tile-level exact skipping (TileSkip).  HI: high-precision sub-tiles (see above; never with FOLD / P1).
template <int D, int R, bool GRAD, bool FOLD, bool P1, bool SEED, bool TSKIP, bool HI>
__device__ __forceinline__ void stream_rows(URow<D, R, GRAD> (&st)[R], const float* __restrict__ pts, int strideP,
                                            const float* ch, int ncols, float coef, float* wsm, int lane,
                                            const TileSkip ts = TileSkip{}) {
  static_assert(!(HI && (FOLD || P1)), "high-precision sub-tiles exist for the unfolded squared cost only");
  static_assert(!TSKIP || (D == 2 && SEED), "tile skipping needs the D = 2 tile boxes and a seeded reference");
  constexpr int T = stream_tile_cols<D>();
  const float2 coef2 = make_float2(coef, coef);
  const float inv_ncoef = -1.0f / coef;
  const float big = exp2f(kTauS);
  const int ntiles = (ncols + T - 1) / T;
  auto issue = [&](int t) {
    const int c = t * T + lane * 4;
    if (lane * 4 < T && c < ncols) {
      float* dst = wsm + (t & 1) * (D + 1) * T + lane * 4;
#pragma unroll
      for (int d = 0; d < D; ++d) cp_async16(dst + d * T, pts + (size_t)d * strideP + c);
      cp_async16(dst + D * T, ch + c);
    }
    cp_async_commit();
  };
  issue(0);
  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) issue(t + 1); else cp_async_commit();  // keep one group per iteration
    cp_async_wait1();
    __syncwarp();
    const float* tb = wsm + (t & 1) * (D + 1) * T;
    const int n = min(T, ncols - t * T);
#ifndef KDOT_NO_TILE_SKIP
    constexpr bool kTileSkip = true;
#else
    constexpr bool kTileSkip = false;  // A/B build: same order, seeds and arithmetic, every tile evaluated
#endif
    unsigned int dead = 0u;  // bit s: 32-column sub-tile s of this tile contributes exactly nothing to any row of the warp
    if (TSKIP && kTileSkip) {
#pragma unroll
      for (int sb = 0; sb < T / 32; ++sb) {
        if (sb * 32 >= n) break;
        const float4 bx = ts.tbox[t * (T / 32) + sb];          // warp-uniform loads
        const float hm = __ldcg(ts.hmax + t * (T / 32) + sb);
        bool far = true;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const float px = -st[k].nx[0], py = -st[k].nx[D - 1];
          const float dx = fmaxf(fmaxf(bx.x - px, px - bx.z), 0.f);
          const float dy = fmaxf(fmaxf(bx.y - py, py - bx.w), 0.f);
          // b = fl(coef * fl(dist^2) + hmax) bounds every true argument h_j + coef |p_i - p_j|^2 of the sub-tile from above
          // (coef < 0, |p_i - p_j| >= dist(p_i, box), h_j <= hmax) up to its own rounding, <= 2^-22 (|hmax| + |coef dist^2|);
          // an evaluated argument fl(coef * fl(d^2) + h_j) exceeds its true value by at most the same expression with
          // d in place of dist.  ex2.approx.ftz returns exactly +0 below -126, so "b - mref < -140 - 2^-20 (|hmax| +
          // |coef dist^2|)" leaves 14 units plus twice the worst-case rounding: nothing but exact zeros is dropped.
          const float cd = coef * fmaf(dx, dx, dy * dy);
          far = far && ((cd + hm) - st[k].mref < -140.f - 9.5367e-7f * (fabsf(hm) + fabsf(cd)));
        }
        if (__all_sync(0xffffffffu, far)) dead |= 1u << sb;
      }
    }
#pragma unroll 1
    for (int sb = 0; sb < n; sb += 32) {
      if (TSKIP && ((dead >> (sb >> 5)) & 1u)) continue;
      const int je = min(sb + 32, n);   // st[k].s == 0 here: the previous sub-tile was folded into (tot, comp)
      bool grad_hi = false;
#ifdef KDOT_NO_GRAD_HI
      constexpr bool kGradHi = false;
#else
      constexpr bool kGradHi = true;
#endif
      if (kGradHi && HI && GRAD && subtile_hmag<D>(tb, T, sb, n, lane) * stream_hi_args<D>(wsm)->mag_fac > 1.0f) {
        // gradient round: float64 arguments for the sub-tiles that carry a visible share of a row's weights.  Screen with
        // a sums-only fp32 pass on copies (no state is touched); an overflow counts as "visible".
        URow<D, R, false> tmp[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
#pragma unroll
          for (int d = 0; d < D; ++d) tmp[k].nx[d] = st[k].nx[d];
          tmp[k].mref = st[k].mref; tmp[k].mu = st[k].mu; tmp[k].s = make_float2(0.f, 0.f);
        }
#pragma unroll 2
        for (int j = sb; j < je; j += 4) {
          float4 X[D];
#pragma unroll
          for (int d = 0; d < D; ++d) X[d] = *reinterpret_cast<const float4*>(tb + d * T + j);
          const float4 H = *reinterpret_cast<const float4*>(tb + D * T + j);
          stream_chunk_spec<D, R, false>(tmp, X, H, coef2);
        }
        bool sig = false;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const float gain = tmp[k].s.x + tmp[k].s.y;
          sig |= !(gain <= kSigThr * fmaxf(st[k].tot + gain, 1.0f));
        }
        grad_hi = __any_sync(0xffffffffu, sig);
        if (lane == 0) if (long long* dbg = stream_hi_args<D>(wsm)->dbg) { atomicAdd((unsigned long long*)dbg + 5, 1ull); if (grad_hi) atomicAdd((unsigned long long*)dbg + 7, 1ull); }
      }
      if (HI && GRAD && grad_hi) {
        stream_subtile_hi<D, R, GRAD>(st, stream_hi_staging<D>(wsm), tb, T, sb, je - sb, lane, stream_hi_args<D>(wsm)->chlo + t * T + sb, sb + lane < n, stream_hi_args<D>(wsm)->coefd, inv_ncoef, big);
      } else {
        bool redo = true;
        if (!GRAD && !P1 && D <= 2) {  // speculate: whole sub-tile without overflow tests, one check at the end (larger D: registers)
#pragma unroll 4
          for (int j = sb; j < je; j += 4) {
            float4 X[D];
#pragma unroll
            for (int d = 0; d < D; ++d) X[d] = *reinterpret_cast<const float4*>(tb + d * T + j);
            const float4 H = *reinterpret_cast<const float4*>(tb + D * T + j);
            stream_chunk_spec<D, R, FOLD>(reinterpret_cast<URow<D, R, false>(&)[R]>(st), X, H, coef2);
          }
          redo = false;
#pragma unroll
          for (int k = 0; k < R; ++k) redo |= !(st[k].s.x + st[k].s.y <= big);
          if (redo) {
#pragma unroll
            for (int k = 0; k < R; ++k) st[k].s = make_float2(0.f, 0.f);
          }
        }
        if (redo) {
#pragma unroll 1
          for (int j = sb; j < je; j += 4) {
            float4 X[D];
#pragma unroll
            for (int d = 0; d < D; ++d) X[d] = *reinterpret_cast<const float4*>(tb + d * T + j);
            const float4 H = *reinterpret_cast<const float4*>(tb + D * T + j);
            stream_chunk<D, R, GRAD, FOLD, P1>(st, X, H, coef2, inv_ncoef, big);
          }
        }
#ifdef KDOT_NO_POT_HI
        constexpr bool kPotHi = false;
#else
        constexpr bool kPotHi = true;
#endif
        if (kPotHi && HI && !GRAD) {  // screen: does this sub-tile carry a visible share of any row's sum?
          bool sig = false;
#pragma unroll
          for (int k = 0; k < R; ++k) {
            const float gain = st[k].s.x + st[k].s.y;
            sig |= gain > kSigThr * fmaxf(st[k].tot + gain, 1.0f);
          }
          if (__any_sync(0xffffffffu, sig) && subtile_hmag<D>(tb, T, sb, n, lane) * stream_hi_args<D>(wsm)->mag_fac > 1.0f) {
            if (lane == 0) if (long long* dbg = stream_hi_args<D>(wsm)->dbg) atomicAdd((unsigned long long*)dbg + 6, 1ull);
#pragma unroll
            for (int k = 0; k < R; ++k) st[k].s = make_float2(0.f, 0.f);  // drop the fp32 contribution
            stream_subtile_hi<D, R, GRAD>(st, stream_hi_staging<D>(wsm), tb, T, sb, je - sb, lane, stream_hi_args<D>(wsm)->chlo + t * T + sb, sb + lane < n, stream_hi_args<D>(wsm)->coefd, inv_ncoef, big);
          }
        }
      }
      if (SEED) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const float gain = st[k].s.x + st[k].s.y;
          if (gain > st[k].gbest) { st[k].gbest = gain; st[k].jb = t * T + sb; }
        }
      }
#pragma unroll
      for (int k = 0; k < R; ++k) urow_fold(st[k]);
    }
    __syncwarp();  // every lane is done with this buffer before tile t+2 overwrites it
  }
}

template <int D, int R, bool GRAD>
__device__ __forceinline__ void urow_reset(URow<D, R, GRAD> (&st)[R], float inv_ncoef) {
#pragma unroll
  for (int k = 0; k < R; ++k) {
    st[k].mref = kNegBig;
    st[k].jb = 0;
    st[k].gbest = 0.f;
    st[k].tot = 0.f; st[k].comp = 0.f;
    st[k].mu = make_float2(kNegBig * inv_ncoef, kNegBig * inv_ncoef);
    st[k].s = make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 0; d < (GRAD ? D : 1); ++d) st[k].g[d] = make_float2(0.f, 0.f);
  }
}

// Cold rounds: start the sweep of row k with the reference exponent set to the max of the 4-column chunk that held
// the row's running max in the PREVIOUS round (jb[k]).  It is a true value of this row, hence a valid lower bound of the
// row max, and it is almost always within a few units of it: the far chunks of the (Morton-ordered) sweep are then
// below the skip threshold from the first column on, and the climb towards the near region needs no re-basing.
template <int D, int R, bool GRAD, bool P1>
__device__ __forceinline__ void urow_seed(URow<D, R, GRAD> (&st)[R], const float* __restrict__ pts, int strideP,
                                          const float* ch, float coef, const int (&jb)[R]) {
  const float2 coef2 = make_float2(coef, coef);
#pragma unroll
  for (int k = 0; k < R; ++k) {
    float vm = kNegBig;
#pragma unroll 2
    for (int c = 0; c < 32; c += 4) {  // the clouds are padded to 32-point tiles: the whole sub-tile is readable
      float2 q0 = make_float2(0.f, 0.f), q1 = q0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4 X = *reinterpret_cast<const float4*>(pts + (size_t)d * strideP + jb[k] + c);
        const float2 nd = make_float2(st[k].nx[d], st[k].nx[d]);
        const float2 a0 = __fadd2_rn(make_float2(X.x, X.y), nd);
        const float2 a1 = __fadd2_rn(make_float2(X.z, X.w), nd);
        q0 = d == 0 ? __fmul2_rn(a0, a0) : __ffma2_rn(a0, a0, q0);
        q1 = d == 0 ? __fmul2_rn(a1, a1) : __ffma2_rn(a1, a1, q1);
      }
      if (P1) { q0 = p1_norm(q0); q1 = p1_norm(q1); }
      const float4 H = ld_cg4(ch + jb[k] + c);
      const float2 v0 = __ffma2_rn(q0, coef2, make_float2(H.x, H.y));
      const float2 v1 = __ffma2_rn(q1, coef2, make_float2(H.z, H.w));
      vm = fmaxf(vm, fmaxf(fmaxf(v0.x, v0.y), fmaxf(v1.x, v1.y)));
    }
    st[k].mref = vm;
    st[k].jb = jb[k];
  }
}

__device__ __forceinline__ long long cell_index(const SinkhornParams& b, bool student, int cell, int slot) {
  return student ? (long long)cell * b.s_cell_n + (long long)slot * b.s_slot_n
                 : (long long)cell * b.s_cell_m + (long long)slot * b.s_slot_m;
}

// One warp unit: rows [blk*32R, blk*32R + 32R) of one cloud of problem `prob` against one column set, round r.
template <int D, int R, bool P1>
__device__ __forceinline__ void stream_unit(const StreamParams& p, int r, int prob, int uu, int lane, float* wsm) {
  constexpr bool kSeed = !P1;                // cold rounds: seeded reference exponent (urow_seed) + sub-tile bookkeeping
  constexpr bool kTSkip = (D == 2) && !P1;   // exact skipping of all-underflow sub-tiles in the cold rounds (TileSkip)
  constexpr int T = stream_tile_cols<D>();
  const SinkhornParams& b = p.b;
  const int B = b.B, strideP = p.strideP;
  const int img = prob / B, slot = prob - img * B;
  const int nrounds = b.sched_rounds[img];
  if (r >= nrounds) return;  // image finished (or skipped)
  const bool last = r == nrounds - 1;
  const int cur = r & 1;
  const int N = b.cu_n[img + 1] - b.cu_n[img], M = b.cu_m[img + 1] - b.cu_m[img];
  const int Nq = (N + 3) & ~3, Mq = (M + 3) & ~3;
  const bool rows_x = uu < 2 * p.nbx;
  const int ub = rows_x ? uu : uu - 2 * p.nbx;
  const int blk = ub >> 1;
  const bool own = (ub & 1) == 0;
  const int rcount = rows_x ? N : M;
  if (blk * 32 * R >= rcount) return;  // padded unit of a smaller image
  const int rbase = rows_x ? 0 : p.nqMax;
  const RoundConst rc = b.sched[(size_t)img * KDOT_MAX_ROUNDS + r];
  const float eps0 = b.sched[(size_t)img * KDOT_MAX_ROUNDS].eps;
  const bool warm = !P1 && rc.eps * 256.0f >= eps0;   // reference exponent folded into the distance chain (one op less per pair)
  // float64 pair arguments only for problems whose centred offsets are large enough for fp32 to hurt: above
  // kStreamHiMagnitude log2-units for the large batches, above kHiMagnitude for clouds of up to kStreamSmallPoints points
  const float gate_ratio = p.strideP <= kStreamSmallPoints ? 1.0f : kHiMagnitude / kStreamHiMagnitude;
  // Centred offsets.  Unbalanced OT with unequal total masses puts a common term of order rho * log(mass ratio) / eps
  // (1e4 log2-units at eps = 1e-6) into every potential of a cloud; it cancels in h_j - max_j, but an fp32 head of h
  // would spend its mantissa on it.  The h of every (type S/C, cloud X/Y) set is therefore published relative to
  // c = pot(row 0 of that set, two rounds earlier) * hmul -- a value every unit of the round can read without a race --
  // and the same c is added back to the log-sum-exp in float64.  pref[slot q % 3] holds the raw potentials for the h
  // consumed in round q; it is written in round q - 2 by the unit that owns row 0.
  double* pref = p.pref + (size_t)prob * 12;
  const double hmul_prev = r > 0 ? b.sched[(size_t)img * KDOT_MAX_ROUNDS + r - 1].hmuld : 0.0;
  unsigned int* hmag = p.hmag + (size_t)prob * 3;
  const bool hi = !P1 && is_hi_round(r, nrounds, rc.eps, eps0, kStreamHiRounds) &&
                  __uint_as_float(__ldcg(hmag + r % 3)) * hi_mag_factor(r, nrounds) * gate_ratio > 1.0f;
  if (uu == 0 && lane == 0) hmag[(r + 2) % 3] = 0u;  // slot of round r + 2: last read in round r - 1, next written in round r + 1
  const double rho = b.rho;
  const float* pts = p.pts + (size_t)prob * D * strideP;
  const float* lw2 = p.lw2 + (size_t)prob * strideP;
  double* potS = p.pot + (size_t)prob * 2 * strideP;
  double* potC = potS + strideP;
  const float* hSc = p.h + ((size_t)prob * 4 + cur * 2) * strideP;
  const float* hCc = hSc + strideP;
  float* hSn = p.h + ((size_t)prob * 4 + (cur ^ 1) * 2) * strideP;
  float* hCn = hSn + strideP;
  const float* lSc = p.hlo + ((size_t)prob * 4 + cur * 2) * strideP;
  const float* lCc = lSc + strideP;
  float* lSn = p.hlo + ((size_t)prob * 4 + (cur ^ 1) * 2) * strideP;
  float* lCn = lSn + strideP;
  unsigned short* jbS = p.jb + (size_t)prob * 2 * strideP;  // indexed by staged row position, like the potentials
  unsigned short* jbC = jbS + strideP;
  const int ntile = strideP >> 5;
  const float4* tbox = p.tbox + (size_t)prob * ntile;
  const float* hmSc = p.hmax + ((size_t)prob * 4 + cur * 2) * ntile;  // same layout as h, one value per 32 points
  const float* hmCc = hmSc + ntile;
  float* hmSn = p.hmax + ((size_t)prob * 4 + (cur ^ 1) * 2) * ntile;
  float* hmCn = hmSn + ntile;
  const float hi_mag_fac = hi_mag_factor(r, nrounds) * gate_ratio;

  if (last && rows_x) {
    if (!own) return;  // the student's last round is done by the "own" unit for both column sets
    // one row per pass keeps the gradient accumulators in registers for every D
    for (int k0 = 0; k0 < R; ++k0) {
      const int i = blk * 32 * R + lane + 32 * k0;
      const bool act = i < N;
      const int src = act ? i : 0;
      URow<D, 1, true> st[1];
      urow_reset<D, 1, true>(st, -1.0f / rc.coef);
#pragma unroll
      for (int d = 0; d < D; ++d) st[0].nx[d] = -pts[(size_t)d * strideP + src];
      TileSkip ts{};
      if (kSeed && !warm) {
        const int jb1[1] = {32 * (int)__ldcg(jbS + src)};
        urow_seed<D, 1, true, P1>(st, pts, strideP, hSc, rc.coef, jb1);
        ts.tbox = tbox; ts.hmax = hmSc;
      }
      if (hi) stream_hi_publish<D>(wsm, lane, lSc, rc.coefd, hi_mag_fac, b.dbg_clk);
      if (hi) stream_rows<D, 1, true, false, P1, kSeed, kTSkip, !P1>(st, pts, strideP, hSc, Nq, rc.coef, wsm, lane, ts);
      else if (!warm) stream_rows<D, 1, true, false, P1, kSeed, kTSkip, false>(st, pts, strideP, hSc, Nq, rc.coef, wsm, lane, ts);
      else stream_rows<D, 1, true, false, P1, false, false, false>(st, pts, strideP, hSc, Nq, rc.coef, wsm, lane);
      const float sS = st[0].tot;
      const double S = rc.scaled * ((double)st[0].mref + urow_lg2_sum(st[0]) + __ldcg(pref + (r % 3) * 4 + 0) * hmul_prev);  // type S, cloud X
      float gS[D];
#pragma unroll
      for (int d = 0; d < D; ++d) gS[d] = (st[0].g[d].x + st[0].g[d].y) / sS;
      urow_reset<D, 1, true>(st, -1.0f / rc.coef);
      if (kSeed && !warm) {
        const int jb1[1] = {32 * (int)__ldcg(jbC + src)};
        urow_seed<D, 1, true, P1>(st, pts + p.nqMax, strideP, hCc + p.nqMax, rc.coef, jb1);
        ts.tbox = tbox + (p.nqMax >> 5); ts.hmax = hmCc + (p.nqMax >> 5);
      }
      if (hi) stream_hi_publish<D>(wsm, lane, lCc + p.nqMax, rc.coefd, hi_mag_fac, b.dbg_clk);
      if (hi) stream_rows<D, 1, true, false, P1, kSeed, kTSkip, !P1>(st, pts + p.nqMax, strideP, hCc + p.nqMax, Mq, rc.coef, wsm, lane, ts);
      else if (!warm) stream_rows<D, 1, true, false, P1, kSeed, kTSkip, false>(st, pts + p.nqMax, strideP, hCc + p.nqMax, Mq, rc.coef, wsm, lane, ts);
      else stream_rows<D, 1, true, false, P1, false, false, false>(st, pts + p.nqMax, strideP, hCc + p.nqMax, Mq, rc.coef, wsm, lane);
      if (!act) continue;
      const float sC = st[0].tot;
      const double C = rc.scaled * ((double)st[0].mref + urow_lg2_sum(st[0]) + __ldcg(pref + (r % 3) * 4 + 3) * hmul_prev);  // type C, cloud Y
      const RowFinal f = row_final(S, C, rho, rc.eps);
      const float lam = rho < 0.0 ? 1.f : (float)(1.0 / (1.0 + (double)rc.eps / rho));
      const float gfac = rho < 0.0 ? 1.f : (float)((rho + 0.5 * (double)rc.eps) / rho) * lam;
      const long long g = cell_index(b, true, b.cu_n[img] + p.perm[(size_t)prob * strideP + i], slot);
      const float wg = b.ws ? b.ws[g] : __fdiv_rn(1.0f, (float)N);
      hSn[i] = wg * f.term;  // per-row loss terms go to the h buffer the last round no longer needs
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float gv = wg * gfac * (f.eS * gS[d] - f.eC * ((st[0].g[d].x + st[0].g[d].y) / sC));
        if (b.normalize && D == 2) gv = __fdiv_rn(gv, d == 0 ? b.w : b.h);
        b.grad_xs[(size_t)D * g + d] = gv;
      }
      if (b.grad_ws) b.grad_ws[g] = f.term;
    }
    return;
  }

  const bool cols_x = (rows_x == own);
  const int coff = cols_x ? 0 : p.nqMax;
  const float* cpts = pts + coff;
  const float* ch = (own ? hSc : hCc) + coff;
  const int ncols = cols_x ? Nq : Mq;
  URow<D, R, false> st[R];
  urow_reset<D, R, false>(st, -1.0f / rc.coef);
  int ridx[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = blk * 32 * R + lane + 32 * k;
    ridx[k] = i < rcount ? rbase + i : -1;
    const int src = ridx[k] >= 0 ? ridx[k] : rbase;
#pragma unroll
    for (int d = 0; d < D; ++d) st[k].nx[d] = -pts[(size_t)d * strideP + src];
  }
  unsigned short* jbArr = own ? jbS : jbC;
  if (warm) {
#ifdef KDOT_NO_FOLD
    constexpr bool kFold = false;  // A/B build
#else
    constexpr bool kFold = !P1;
#endif
    stream_rows<D, R, false, kFold, false, false, false, false>(st, cpts, strideP, ch, ncols, rc.coef, wsm, lane);
  } else {
    TileSkip ts{};
    if (kSeed) {
      int jbv[R];
#pragma unroll
      for (int k = 0; k < R; ++k) jbv[k] = 32 * (int)__ldcg(jbArr + (ridx[k] >= 0 ? ridx[k] : rbase));
      urow_seed<D, R, false, P1>(st, cpts, strideP, ch, rc.coef, jbv);
      ts.tbox = tbox + (coff >> 5);
      ts.hmax = (own ? hmSc : hmCc) + (coff >> 5);
    }
    if (hi) stream_hi_publish<D>(wsm, lane, (own ? lSc : lCc) + coff, rc.coefd, hi_mag_fac, b.dbg_clk);
    if (hi) stream_rows<D, R, false, false, P1, kSeed, kTSkip, !P1>(st, cpts, strideP, ch, ncols, rc.coef, wsm, lane, ts);
    else stream_rows<D, R, false, false, P1, kSeed, kTSkip, false>(st, cpts, strideP, ch, ncols, rc.coef, wsm, lane, ts);
  }
  // new potentials / next round's h in float64; fp32 head + tail of h, per-tile maximum of the head (skip test)
  const double c_in = __ldcg(pref + (r % 3) * 4 + (own ? 0 : 2) + (cols_x ? 0 : 1)) * hmul_prev;      // consumed h: its centre
  const double p_out = last ? 0.0 : __ldcg(pref + ((r + 1) % 3) * 4 + (own ? 0 : 2) + (rows_x ? 0 : 1));  // published h: centred on this
  double hv[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    hv[k] = (double)kNegBig;
    if (ridx[k] < 0) continue;
    if (kSeed && !warm) jbArr[ridx[k]] = (unsigned short)(st[k].jb >> 5);
    // warm rounds fold the reference into the distance chain as mu = fl(mref / -coef): the sums are relative to
    // -coef * mu, which differs from mref by the rounding of mu (6e-8 |mref|: up to 1e-5) -- add back what was subtracted
    const double ref = warm ? -(double)rc.coef * (double)st[k].mu.x : (double)st[k].mref;
    const double lse = ref + urow_lg2_sum(st[k]) + c_in;
    double* pot = own ? potS : potC;
    const double nv = rc.scaled * lse;
    const double pv = (r == 0 || last) ? nv : 0.5 * (__ldcg(pot + ridx[k]) + nv);
    pot[ridx[k]] = pv;
    if (!last) {
      hv[k] = fma(pv - p_out, rc.hmuld, (double)__ldcg(lw2 + ridx[k]));
      if (ridx[k] == rbase) pref[((r + 2) % 3) * 4 + (own ? 0 : 2) + (rows_x ? 0 : 1)] = pv;  // row 0 of this set: centre of the h consumed in round r + 2
      const float hh = (float)hv[k];
      (own ? hSn : hCn)[ridx[k]] = hh;
      (own ? lSn : lCn)[ridx[k]] = (float)(hv[k] - (double)hh);
    }
  }
  if (!last) {
    float am = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k) am = fmaxf(am, ridx[k] >= 0 ? fabsf((float)hv[k]) : 0.f);
    am = warp_max(am);
    if (lane == 0) atomicMax(hmag + (r + 1) % 3, __float_as_uint(am));  // non-negative floats order like their bits
  }
  if (kTSkip && !last) {  // per-tile maximum of the h values this unit publishes (rows of pass k = one tile)
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float hm = warp_max((float)hv[k]);
      const int tile = (rbase + blk * 32 * R + 32 * k) >> 5;
      if (lane == 0 && tile < ntile && blk * 32 * R + 32 * k < rcount) (own ? hmSn : hmCn)[tile] = hm;
    }
  }
}

template <int D, int R, bool P1>
__global__ void __launch_bounds__(kStreamThreads, (D <= 2 ? KDOT_STREAM_MINBLOCKS : 2)) kdot_stream_kernel(StreamParams p) {
  cg::grid_group grid = cg::this_grid();
  const SinkhornParams& b = p.b;
  const int B = b.B;
  const int nprob = b.nimg * B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int strideP = p.strideP;
  if (b.dbg_clk && blockIdx.x == 0 && threadIdx.x == 0) b.dbg_clk[0] = (long long)global_ns();
  extern __shared__ __align__(16) float s_tiles[];  // per warp: 2 x (D+1) x T floats (cp.async column tiles)
  float* wsm = s_tiles + (size_t)warp * stream_warp_smem_floats<D>();
  __shared__ float s_box[kStreamThreads / 32][2 * D];
  __shared__ int s_info[2];
  __shared__ ImgSched s_is;

  // ---------------- phase 0a: per image normalise / bbox / schedule ----------------
  for (int img = blockIdx.x; img < b.nimg; img += gridDim.x) {
    const int n0 = b.cu_n[img], N = b.cu_n[img + 1] - n0;
    const int m0 = b.cu_m[img], M = b.cu_m[img + 1] - m0;
    float mn[D], mx[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { mn[d] = 3.0e38f; mx[d] = -3.0e38f; }
    for (int t = threadIdx.x; t < (N + M) * B; t += blockDim.x) {
      const int q = t / B, slot = t - q * B;
      const bool stu = q < N;
      float* base = stu ? b.xs : b.xt;
      const long long g = cell_index(b, stu, stu ? n0 + q : m0 + q - N, slot);
      float v[D];
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = base[(size_t)D * g + d];
      if (b.normalize && D == 2) {
        v[0] = __fdiv_rn(v[0], b.w);
        v[D - 1] = __fdiv_rn(v[D - 1], b.h);
        base[(size_t)D * g] = v[0];
        base[(size_t)D * g + D - 1] = v[D - 1];
      }
#pragma unroll
      for (int d = 0; d < D; ++d) { mn[d] = fminf(mn[d], v[d]); mx[d] = fmaxf(mx[d], v[d]); }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) { mn[d] = warp_min(mn[d]); mx[d] = warp_max(mx[d]); }
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < D; ++d) { s_box[warp][d] = mn[d]; s_box[warp][D + d] = mx[d]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float lo = s_box[0][d], hi = s_box[0][D + d];
        for (int wi = 1; wi < nwarps; ++wi) { lo = fminf(lo, s_box[wi][d]); hi = fmaxf(hi, s_box[wi][D + d]); }
        const float e = __fsub_rn(hi, lo);
        acc = __fadd_rn(acc, __fmul_rn(e, e));
      }
      int status = KDOT_IMG_OK, nits = 0;
      if (N == 0 || M == 0) {
        status = KDOT_IMG_SKIPPED;
      } else {
        const float diam = sqrtf(acc);
        if (!(diam > 0.f) || !isfinite(diam)) {
          status = KDOT_IMG_DEGENERATE;
        } else {
          s_is = image_schedule(diam, b.sp);
          nits = s_is.nits;
          if (nits + 2 > KDOT_MAX_ROUNDS) status = KDOT_IMG_TOO_MANY_ROUNDS;
        }
      }
      s_info[0] = status;
      s_info[1] = nits;
      b.sched_rounds[img] = status == KDOT_IMG_OK ? nits + 2 : 0;
      b.valid[img] = status;
      if (b.nits_per_img) b.nits_per_img[img] = nits;
      if (status != KDOT_IMG_OK) b.loss_per_img[img] = status == KDOT_IMG_SKIPPED ? 0.f : __int_as_float(0x7fc00000);
    }
    __syncthreads();
    const int status = s_info[0], nits = s_info[1];
    if (status == KDOT_IMG_OK) {
      for (int r = threadIdx.x; r < nits + 2; r += blockDim.x)
        b.sched[(size_t)img * KDOT_MAX_ROUNDS + r] = make_round_const(r, s_is, b.sp);
    } else {
      const float fill = status == KDOT_IMG_SKIPPED ? 0.f : __int_as_float(0x7fc00000);
      for (int t = threadIdx.x; t < N * B; t += blockDim.x) {
        const int q = t / B, slot = t - q * B;
        const long long g = cell_index(b, true, n0 + q, slot);
#pragma unroll
        for (int d = 0; d < D; ++d) b.grad_xs[(size_t)D * g + d] = fill;
        if (b.grad_ws) b.grad_ws[g] = fill;
      }
      if (b.loss_per_slot)
        for (int s = threadIdx.x; s < B; s += blockDim.x) b.loss_per_slot[(size_t)img * B + s] = fill;
    }
    __syncthreads();
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < KDOT_MAX_ROUNDS; r += gridDim.x * blockDim.x) p.ctr[r] = 0u;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nprob; q += gridDim.x * blockDim.x) {
    p.done[q] = 0u;
    p.hmag[3 * q] = 0u; p.hmag[3 * q + 1] = 0u; p.hmag[3 * q + 2] = 0u;
    for (int a = 0; a < 12; ++a) p.pref[12 * (size_t)q + a] = 0.0;
  }
  grid.sync();

  if (b.dbg_clk && blockIdx.x == 0 && threadIdx.x == 0) b.dbg_clk[1] = (long long)global_ns();
  // ---------------- phase 0b: stage problems (SoA, padded) ----------------
  int max_rounds = 0;
  for (int i = 0; i < b.nimg; ++i) max_rounds = max(max_rounds, b.sched_rounds[i]);
  // One CTA per (problem, cloud).  D = 2 clouds of up to kSortMax points are staged in MORTON ORDER of their own
  // bounding box (rank by counting over 32-bit keys <morton16 | index>: deterministic, O(n^2 / 256) per thread, a few
  // microseconds): rows of a warp unit and columns of a chunk are then spatial neighbours, which is what lets the cold
  // rounds skip whole chunks exactly (stream_chunk SKIP).  perm[] maps a staged position back to its cell.
  {
    constexpr int kSortMax = 4096;
    unsigned int* keys = reinterpret_cast<unsigned int*>(s_tiles);
    __shared__ float s_mm[kStreamThreads / 32][4];
    for (int pc = blockIdx.x; pc < 2 * nprob; pc += gridDim.x) {
      const int prob = pc >> 1;
      const bool in_x = (pc & 1) == 0;
      const int img = prob / B, slot = prob - img * B;
      if (b.sched_rounds[img] <= 0) continue;  // uniform per CTA
      const int n0 = b.cu_n[img], N = b.cu_n[img + 1] - n0;
      const int m0 = b.cu_m[img], M = b.cu_m[img + 1] - m0;
      const int n = in_x ? N : M;
      const int q0 = in_x ? 0 : p.nqMax;
      const int cap = in_x ? p.nqMax : strideP - p.nqMax;
      const float* base = in_x ? b.xs : b.xt;
      const float* wb = in_x ? b.ws : b.wt;
      const int c0 = in_x ? n0 : m0;
      const bool sorted = D == 2 && n > 32 && n <= kSortMax;
      if (sorted) {
        float mnx = 3.0e38f, mny = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const long long g = cell_index(b, in_x, c0 + i, slot);
          const float vx = base[(size_t)D * g], vy = base[(size_t)D * g + D - 1];
          mnx = fminf(mnx, vx); mxx = fmaxf(mxx, vx); mny = fminf(mny, vy); mxy = fmaxf(mxy, vy);
        }
        mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
        if (lane == 0) { s_mm[warp][0] = mnx; s_mm[warp][1] = mny; s_mm[warp][2] = mxx; s_mm[warp][3] = mxy; }
        __syncthreads();
        for (int wi = 0; wi < nwarps; ++wi) {
          mnx = fminf(mnx, s_mm[wi][0]); mny = fminf(mny, s_mm[wi][1]);
          mxx = fmaxf(mxx, s_mm[wi][2]); mxy = fmaxf(mxy, s_mm[wi][3]);
        }
        const float sx = mxx > mnx ? 255.0f / (mxx - mnx) : 0.f, sy = mxy > mny ? 255.0f / (mxy - mny) : 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const long long g = cell_index(b, in_x, c0 + i, slot);
          unsigned int qx = (unsigned int)fminf(fmaxf((base[(size_t)D * g] - mnx) * sx, 0.f), 255.f);
          unsigned int qy = (unsigned int)fminf(fmaxf((base[(size_t)D * g + D - 1] - mny) * sy, 0.f), 255.f);
          qx = (qx | (qx << 4)) & 0x0F0Fu; qx = (qx | (qx << 2)) & 0x3333u; qx = (qx | (qx << 1)) & 0x5555u;
          qy = (qy | (qy << 4)) & 0x0F0Fu; qy = (qy | (qy << 2)) & 0x3333u; qy = (qy | (qy << 1)) & 0x5555u;
          keys[i] = ((qx | (qy << 1)) << 16) | (unsigned int)i;  // unique: ties broken by the cell index
        }
        for (int i = n + threadIdx.x; i < ((n + 3) & ~3); i += blockDim.x) keys[i] = 0xffffffffu;
        __syncthreads();
      }
      for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        const bool real = i < n;
        int rank = i;
        if (sorted && real) {
          const unsigned int ki = keys[i];
          int c = 0;
          for (int j = 0; j < n; j += 4) {
            const uint4 kj = *reinterpret_cast<const uint4*>(keys + j);
            c += (kj.x < ki) + (kj.y < ki) + (kj.z < ki) + (kj.w < ki);
          }
          rank = c;
        }
        const int q = q0 + rank;
        float l2 = kNegBig;
        float v[D];
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = 0.f;
        if (real) {
          const long long g = cell_index(b, in_x, c0 + i, slot);
#pragma unroll
          for (int d = 0; d < D; ++d) v[d] = base[(size_t)D * g + d];
          const float wg = wb ? wb[g] : __fdiv_rn(1.0f, (float)n);
          l2 = (wg > 0.f ? logf(wg) : kLogZeroWeight) * kLog2e;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) p.pts[((size_t)prob * D + d) * strideP + q] = v[d];
        p.lw2[(size_t)prob * strideP + q] = l2;
        float* hb = p.h + (size_t)prob * 4 * strideP;
        hb[q] = l2; hb[strideP + q] = l2; hb[2 * strideP + q] = l2; hb[3 * strideP + q] = l2;
        p.pot[(size_t)prob * 2 * strideP + q] = 0.0;
        p.pot[(size_t)prob * 2 * strideP + strideP + q] = 0.0;
        float* lb = p.hlo + (size_t)prob * 4 * strideP;
        lb[q] = 0.f; lb[strideP + q] = 0.f; lb[2 * strideP + q] = 0.f; lb[3 * strideP + q] = 0.f;
        p.perm[(size_t)prob * strideP + q] = real ? i : -1;
        p.jb[(size_t)prob * 2 * strideP + q] = 0;
        p.jb[(size_t)prob * 2 * strideP + strideP + q] = 0;
      }
      __syncthreads();  // staged values of this cloud are visible to the whole CTA; keys may be reused
      if (D == 2) {  // per 32-point tile: bounding box of its real points and the maximum of the initial h (= log weight)
        const int ntile = strideP >> 5;
        for (int tile = warp; tile * 32 < cap; tile += nwarps) {
          const int q = q0 + tile * 32 + lane;
          const bool real = tile * 32 + lane < cap && p.perm[(size_t)prob * strideP + q] >= 0;
          const float vx = real ? p.pts[((size_t)prob * D) * strideP + q] : 0.f;
          const float vy = real ? p.pts[((size_t)prob * D + D - 1) * strideP + q] : 0.f;
          const float lo_x = warp_min(real ? vx : 3.0e38f), lo_y = warp_min(real ? vy : 3.0e38f);
          const float hi_x = warp_max(real ? vx : -3.0e38f), hi_y = warp_max(real ? vy : -3.0e38f);
          const float hm = warp_max(real ? p.lw2[(size_t)prob * strideP + q] : kNegBig);
          if (lane == 0) {
            const int gt = (q0 >> 5) + tile;
            p.tbox[(size_t)prob * ntile + gt] = make_float4(lo_x, lo_y, hi_x, hi_y);
#pragma unroll
            for (int a = 0; a < 4; ++a) p.hmax[((size_t)prob * 4 + a) * ntile + gt] = hm;
          }
        }
      }
    }
  }
  grid.sync();

  if (b.dbg_clk && blockIdx.x == 0 && threadIdx.x == 0) b.dbg_clk[2] = (long long)global_ns();
  // ---------------- rounds: dataflow over a single global FIFO of warp units ----------------
  // Unit ids enumerate (round, problem, unit-in-problem) in that order.  A unit of round r may start once all units
  // of round r-1 of ITS problem have finished (done[prob] == r * upp); because ids are handed out in FIFO order and
  // the whole grid is co-resident (cooperative launch), the units it waits for are already running, so there is no
  // deadlock -- and with more than a few problems in the batch the previous round of a problem finished long ago,
  // so nobody actually waits and there is no global barrier between rounds.
  const long long upr = (long long)nprob * p.upp;
  const unsigned long long total_ids = (unsigned long long)max_rounds * (unsigned long long)upr;
  unsigned long long* fifo = reinterpret_cast<unsigned long long*>(p.ctr);
  for (;;) {
    unsigned long long id = 0;
    if (lane == 0) id = atomicAdd(fifo, 1ull);
    id = __shfl_sync(0xffffffffu, id, 0);
    if (id >= total_ids) break;
    const int r = (int)(id / (unsigned long long)upr);
    const long long u = (long long)(id - (unsigned long long)r * (unsigned long long)upr);
    const int prob = (int)(u / p.upp), uu = (int)(u - (long long)prob * p.upp);
    // profiling aid (kdot_debug_set_clock_buffer): globaltimer stamps of problem 0's units, [8 + (r*upp+uu)*4 + k]
    long long* stamp = (b.dbg_clk && prob == 0 && lane == 0 && r < 64) ? b.dbg_clk + 8 + ((size_t)r * p.upp + uu) * 4 : nullptr;
    if (stamp) stamp[0] = (long long)global_ns();
    if (r > 0) {
      if (lane == 0) {
        const unsigned int need = (unsigned int)r * (unsigned int)p.upp;
        unsigned int seen;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.done + prob) : "memory");
          if (seen >= need) break;
          __nanosleep(100);
        }
      }
      __syncwarp();
    }
    if (stamp) stamp[1] = (long long)global_ns();
    stream_unit<D, R, P1>(p, r, prob, uu, lane, wsm);
    if (stamp) stamp[2] = (long long)global_ns();
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(p.done + prob, 1u);
    if (stamp) stamp[3] = (long long)global_ns();
  }
  if (b.dbg_clk && blockIdx.x == 0 && threadIdx.x == 0) b.dbg_clk[3] = (long long)global_ns();
  grid.sync();
  if (b.dbg_clk && blockIdx.x == 0 && threadIdx.x == 0) b.dbg_clk[4] = (long long)global_ns();

  // ---------------- final: fixed-order loss reduction, one warp per image ----------------
  for (int img = blockIdx.x * nwarps + warp; img < b.nimg; img += gridDim.x * nwarps) {
    const int nrounds = b.sched_rounds[img];
    if (nrounds <= 0) continue;
    const int N = b.cu_n[img + 1] - b.cu_n[img], M = b.cu_m[img + 1] - b.cu_m[img];
    const float eps_last = b.sched[(size_t)img * KDOT_MAX_ROUNDS + nrounds - 1].eps;
    double tot = 0.0;
    for (int slot = 0; slot < B; ++slot) {
      const int prob = img * B + slot;
      const float* term = p.h + ((size_t)prob * 4 + (((nrounds - 1) & 1) ^ 1) * 2) * strideP;  // see stream_unit (last round)
      const double* potS = p.pot + (size_t)prob * 2 * strideP;
      const double* potC = potS + strideP;
      double acc = 0.0;
      for (int i = lane; i < N; i += 32) acc += (double)__ldcg(term + i);
      for (int j = lane; j < M; j += 32) {
        const RowFinal f = row_final(__ldcg(potS + p.nqMax + j), __ldcg(potC + p.nqMax + j), b.rho, eps_last);
        const long long g = cell_index(b, false, b.cu_m[img] + __ldcg(p.perm + (size_t)prob * strideP + p.nqMax + j), slot);
        const float wg = b.wt ? b.wt[g] : __fdiv_rn(1.0f, (float)M);
        acc += (double)wg * (double)f.term;
      }
      acc = warp_sum(acc);
      if (lane == 0 && b.loss_per_slot) b.loss_per_slot[(size_t)prob] = (float)acc;
      tot += acc;
    }
    if (lane == 0) b.loss_per_img[img] = (float)tot;
  }
}

struct StreamPlan {
  int strideP, nqMax, nbx, nby, upp, R;
  size_t off_pts, off_lw, off_pot, off_h, off_hlo, off_perm, off_jb, off_tbox, off_hmax, off_hmag, off_pref, off_ctr, off_done, off_sched, off_rounds, total;
};

static int rows_per_lane(int D) { return D <= 2 ? KDOT_STREAM_R2 : (D <= 8 ? 2 : 1); }

StreamPlan plan_stream(int nimg, int max_n, int max_m, int B, int D) {
  StreamPlan s;
  s.R = rows_per_lane(D);
  s.nqMax = (max_n + 31) & ~31;  // both clouds start on a 32-point tile boundary (tile boxes / tile maxima of h)
  const int mq = (max_m + 31) & ~31;
  s.strideP = s.nqMax + mq;
  s.nbx = (max_n + 32 * s.R - 1) / (32 * s.R);
  s.nby = (max_m + 32 * s.R - 1) / (32 * s.R);
  s.upp = 2 * (s.nbx + s.nby);
  const size_t nprob = (size_t)nimg * B, P = (size_t)s.strideP;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o = 0;
  s.off_pts = o;    o = up(o + nprob * D * P * 4);
  s.off_lw = o;     o = up(o + nprob * P * 4);
  s.off_pot = o;    o = up(o + nprob * 2 * P * 8);
  s.off_h = o;      o = up(o + nprob * 4 * P * 4);
  s.off_hlo = o;    o = up(o + nprob * 4 * P * 4);
  s.off_perm = o;   o = up(o + nprob * P * 4);
  s.off_jb = o;     o = up(o + nprob * 2 * P * 2);
  s.off_tbox = o;   o = up(o + nprob * (P / 32) * 16);
  s.off_hmax = o;   o = up(o + nprob * 4 * (P / 32) * 4);
  s.off_hmag = o;   o = up(o + nprob * 3 * 4);
  s.off_pref = o;   o = up(o + nprob * 12 * 8);
  s.off_ctr = o;    o = up(o + (size_t)KDOT_MAX_ROUNDS * 4);
  s.off_done = o;   o = up(o + nprob * 4);
  s.off_sched = o;  o = up(o + (size_t)nimg * KDOT_MAX_ROUNDS * sizeof(RoundConst));
  s.off_rounds = o; o = up(o + (size_t)nimg * 4);
  s.total = o;
  return s;
}

size_t stream_workspace_bytes(int nimg, int max_n, int max_m, int B, int D) {
  return plan_stream(nimg, max_n, max_m, B, D).total;
}

bool stream_supports_dim(int D) { return D == 1 || D == 2 || D == 3 || D == 4 || D == 8 || D == 16; }

template <int D, int R, bool P1 = false>
static cudaError_t launch_stream_t(StreamParams& sp, cudaStream_t stream) {
  static int blocks_per_sm = 0, sms = 0;
  if (blocks_per_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem_bytes = (kStreamThreads / 32) * stream_warp_smem_floats<D>() * (int)sizeof(float);
    if (smem_bytes > 48 * 1024) {
      cudaError_t ea = cudaFuncSetAttribute(kdot_stream_kernel<D, R, P1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (ea != cudaSuccess) return ea;
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kdot_stream_kernel<D, R, P1>, kStreamThreads,
                                                                  (kStreamThreads / 32) * stream_warp_smem_floats<D>() * sizeof(float));
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
  }
  void* args[] = {&sp};
  return cudaLaunchCooperativeKernel((const void*)kdot_stream_kernel<D, R, P1>, dim3(blocks_per_sm * sms), dim3(kStreamThreads),
                                     args, (kStreamThreads / 32) * stream_warp_smem_floats<D>() * sizeof(float), stream);
}

cudaError_t launch_stream(const SinkhornParams& prm, int D, int max_n, int max_m, void* workspace, cudaStream_t stream) {
  const StreamPlan pl = plan_stream(prm.nimg, max_n, max_m, prm.B, D);
  char* base = (char*)workspace;
  StreamParams sp;
  sp.b = prm;
  sp.b.sched = (RoundConst*)(base + pl.off_sched);
  sp.b.sched_rounds = (int32_t*)(base + pl.off_rounds);
  sp.D = D;
  sp.strideP = pl.strideP; sp.nqMax = pl.nqMax; sp.nbx = pl.nbx; sp.nby = pl.nby; sp.upp = pl.upp;
  sp.pts = (float*)(base + pl.off_pts);
  sp.lw2 = (float*)(base + pl.off_lw);
  sp.pot = (double*)(base + pl.off_pot);
  sp.h = (float*)(base + pl.off_h);
  sp.hlo = (float*)(base + pl.off_hlo);
  sp.perm = (int*)(base + pl.off_perm);
  sp.jb = (unsigned short*)(base + pl.off_jb);
  sp.tbox = (float4*)(base + pl.off_tbox);
  sp.hmax = (float*)(base + pl.off_hmax);
  sp.hmag = (unsigned int*)(base + pl.off_hmag);
  sp.pref = (double*)(base + pl.off_pref);
  sp.ctr = (unsigned int*)(base + pl.off_ctr);
  sp.done = (unsigned int*)(base + pl.off_done);
  if (prm.sp.p == 1.0) return D == 2 ? launch_stream_t<2, KDOT_STREAM_R2, true>(sp, stream) : cudaErrorInvalidValue;
  switch (D) {
    case 1: return launch_stream_t<1, KDOT_STREAM_R2>(sp, stream);
    case 2: return launch_stream_t<2, KDOT_STREAM_R2>(sp, stream);
    case 3: return launch_stream_t<3, 2>(sp, stream);
    case 4: return launch_stream_t<4, 2>(sp, stream);
    case 8: return launch_stream_t<8, 2>(sp, stream);
    case 16: return launch_stream_t<16, 1>(sp, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace kdot

