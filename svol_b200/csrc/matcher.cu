// Batched on-GPU Hungarian matching for every decoder layer in one launch (sm_100a).
//
// Replaces PerFrameMatcher.forward / HungarianMatcher.forward (lib/modeling/matcher.py:38-119,
// 131-159): the reference builds the full cross-batch (B*Q) x sum(n) cost matrix on the device,
// copies it to the host and calls scipy.optimize.linear_sum_assignment once per frame.  Here one
// warp owns one assignment problem: it forms only that problem's cost block (the block-diagonal
// entries the reference reads, matcher.py:92-93) and solves it in place.
//
// Bit-exactness: the cost is evaluated in fp32 with the reference's operation order and with
// explicit round-to-nearest intrinsics so the compiler cannot contract a*b+c into an FMA
// (matcher.py:59-85; box_utils.py:9-13,24-37,55-61).  The solver is the shortest-augmenting-path
// algorithm scipy uses (Crouse 2016) on costs promoted to fp64, including its candidate order
// (columns scanned last-to-first), its preference for unassigned columns among equal minima, the
// transpose for tall problems and rows returned in ascending order.
#include <math_constants.h>

#include "common.cuh"
#include "svol_internal.h"

namespace svol {

struct Box { float cx, cy, w, h; };

__device__ __forceinline__ float pair_cost(float p_fg, const Box& a, const Box& t, float w_class, float w_bbox,
                                           float w_giou) {
  // cdist(p=1): sum of absolute differences, accumulated left to right        matcher.py:79
  float l1 = fabsf(__fsub_rn(a.cx, t.cx));
  l1 = __fadd_rn(l1, fabsf(__fsub_rn(a.cy, t.cy)));
  l1 = __fadd_rn(l1, fabsf(__fsub_rn(a.w, t.w)));
  l1 = __fadd_rn(l1, fabsf(__fsub_rn(a.h, t.h)));
  // cxcywh -> xyxy                                                             box_utils.py:9-13
  const float ax0 = __fsub_rn(a.cx, __fmul_rn(0.5f, a.w)), ay0 = __fsub_rn(a.cy, __fmul_rn(0.5f, a.h));
  const float ax1 = __fadd_rn(a.cx, __fmul_rn(0.5f, a.w)), ay1 = __fadd_rn(a.cy, __fmul_rn(0.5f, a.h));
  const float tx0 = __fsub_rn(t.cx, __fmul_rn(0.5f, t.w)), ty0 = __fsub_rn(t.cy, __fmul_rn(0.5f, t.h));
  const float tx1 = __fadd_rn(t.cx, __fmul_rn(0.5f, t.w)), ty1 = __fadd_rn(t.cy, __fmul_rn(0.5f, t.h));
  // IoU                                                                        box_utils.py:24-37
  const float area_a = __fmul_rn(__fsub_rn(ax1, ax0), __fsub_rn(ay1, ay0));
  const float area_t = __fmul_rn(__fsub_rn(tx1, tx0), __fsub_rn(ty1, ty0));
  const float iw = fmaxf(__fsub_rn(fminf(ax1, tx1), fmaxf(ax0, tx0)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(ay1, ty1), fmaxf(ay0, ty0)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_t), inter);
  const float iou = __fdiv_rn(inter, uni);
  // enclosing box                                                              box_utils.py:55-61
  const float cw = fmaxf(__fsub_rn(fmaxf(ax1, tx1), fminf(ax0, tx0)), 0.f);
  const float ch = fmaxf(__fsub_rn(fmaxf(ay1, ty1), fminf(ay0, ty0)), 0.f);
  const float hull = __fmul_rn(cw, ch);
  const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
  // C = w_bbox*L1 + w_giou*(-GIoU) + w_class*(-p_fg)                            matcher.py:85
  return __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_giou, -giou)), __fmul_rn(w_class, -p_fg));
}

__device__ __forceinline__ float fg_prob(float l0, float l1) {   // softmax(logits)[0]   matcher.py:59
  const float m = fmaxf(l0, l1);
  const float e0 = expf(__fsub_rn(l0, m)), e1 = expf(__fsub_rn(l1, m));
  return __fdiv_rn(e0, __fadd_rn(e0, e1));
}

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per problem.  Dynamic shared memory per warp (n_small = min(rows, cols), n_big = max):
//   double u[n_small], v[n_big], spc[n_big]; int path[n_big], col4row[n_small], row4col[n_big],
//   remaining[n_big]; uint8 SR[n_small], SC[n_big]
// When a problem's cost block fits (cost_smem_floats > 0) it is also kept in shared memory: the solver re-reads
// costs once per scanned column per path step, and from the global workspace every such read was an exposed L2
// round trip (C5 stress: 482 us -> see profiles; the default 10 x n_f problems are launch-latency bound either way).
__global__ void __launch_bounds__(32) match_kernel(const MatchArgs a, int max_small, int max_big, int cost_smem_floats) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int lane = threadIdx.x;
  const int P = a.B * a.problems_per_video;
  const int layer = blockIdx.x / P, p = blockIdx.x - layer * P;
  const int video = p / a.problems_per_video, local = p - video * a.problems_per_video;
  const int nrows = a.rows_per_problem;
  const int t0 = a.tgt_off[p], ncols = a.tgt_off[p + 1] - t0;
  if (ncols <= 0 || nrows <= 0) return;

  double* u = reinterpret_cast<double*>(sm_raw);
  double* v = u + max_small;
  double* spc = v + max_big;
  int* path = reinterpret_cast<int*>(spc + max_big);
  int* col4row = path + max_big;
  int* row4col = col4row + max_small;
  int* remaining = row4col + max_big;
  uint8_t* SR = reinterpret_cast<uint8_t*>(remaining + max_big);
  uint8_t* SC = SR + max_small;
  float* Cs = reinterpret_cast<float*>(sm_raw + ((sizeof(double) * (max_small + 2 * max_big) + sizeof(int) * (3 * max_big + max_small) +
                                                   (max_small + max_big) + 15) & ~static_cast<size_t>(15)));
  const bool cost_in_smem = nrows * ncols <= cost_smem_floats;

  // ---- cost block, row-major nrows x ncols, in this layer's slab of the workspace
  const size_t q0 = (static_cast<size_t>(layer) * a.B + video) * a.Q + static_cast<size_t>(local) * nrows;
  float* C = a.cost_ws + static_cast<size_t>(layer) * a.cost_off[P] + a.cost_off[p];
  bool bad = false;
  for (int e = lane; e < nrows * ncols; e += 32) {
    const int r = e / ncols, c = e - r * ncols;
    const float2 lg = reinterpret_cast<const float2*>(a.logits)[q0 + r];
    const float4 pb = reinterpret_cast<const float4*>(a.boxes)[q0 + r];
    const float4 tb = reinterpret_cast<const float4*>(a.tgt_boxes)[t0 + c];
    const Box pa{pb.x, pb.y, pb.z, pb.w}, ta{tb.x, tb.y, tb.z, tb.w};
    const float cost = pair_cost(fg_prob(lg.x, lg.y), pa, ta, a.w_class, a.w_bbox, a.w_giou);
    bad |= (cost != cost) || (cost == -CUDART_INF_F);
    C[e] = cost;
    if (cost_in_smem) Cs[e] = cost;
  }
  if (__any_sync(0xffffffffu, bad)) {      // scipy: "matrix contains invalid numeric entries"
    if (lane == 0) atomicOr(a.status, 1);
    return;
  }
  __syncwarp();

  // ---- rectangular LSAP on the short side (scipy transposes tall problems)
  const bool transposed = ncols < nrows;
  const int nr = transposed ? ncols : nrows;      // rows of the working problem
  const int nc = transposed ? nrows : ncols;
  const float* Cr = cost_in_smem ? Cs : C;
  auto cost_at = [&](int i, int j) -> double {
    return static_cast<double>(transposed ? Cr[j * ncols + i] : Cr[i * ncols + j]);
  };
  for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = lane; j < nc; j += 32) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  __syncwarp();

  for (int cur = 0; cur < nr; ++cur) {
    for (int j = lane; j < nc; j += 32) { remaining[j] = nc - j - 1; SC[j] = 0; spc[j] = CUDART_INF; }
    for (int i = lane; i < nr; i += 32) SR[i] = 0;
    __syncwarp();
    int n_rem = nc, i = cur, sink = -1;
    double best = 0.0;
    while (sink == -1) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      double lowest = CUDART_INF;
      int first_pos = INT_MAX, last_free = -1;
      for (int t = lane; t < n_rem; t += 32) {
        const int j = remaining[t];
        const double r = ((best + cost_at(i, j)) - ui) - v[j];
        double s = spc[j];
        if (r < s) { path[j] = i; spc[j] = r; s = r; }
        const bool free_col = row4col[j] == -1;
        if (s < lowest) { lowest = s; first_pos = t; last_free = free_col ? t : -1; }
        else if (s == lowest && free_col) last_free = t;
      }
      const double gl = warp_min_d(lowest);
      const bool mine = lowest == gl && first_pos != INT_MAX;
      const int fp = warp_min_i(mine ? first_pos : INT_MAX);
      const int lf = warp_max_i(mine ? last_free : -1);
      if (!(gl < CUDART_INF)) {             // infeasible (cannot happen with finite costs)
        if (lane == 0) atomicOr(a.status, 2);
        return;
      }
      // sequential rule: the first minimum wins unless a later (or the same) equal minimum is free
      const int pick = lf >= 0 ? lf : fp;
      best = gl;
      __syncwarp();
      const int j = remaining[pick];
      const int owner = row4col[j];
      if (owner == -1) sink = j; else i = owner;
      __syncwarp();
      if (lane == 0) { SC[j] = 1; remaining[pick] = remaining[n_rem - 1]; }
      --n_rem;
      __syncwarp();
    }
    // dual update
    for (int r = lane; r < nr; r += 32) {
      if (r == cur) u[r] += best;
      else if (SR[r]) u[r] += best - spc[col4row[r]];
    }
    for (int j = lane; j < nc; j += 32)
      if (SC[j]) v[j] -= best - spc[j];
    __syncwarp();
    // augment along the path (sequential)
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = path[j];
        row4col[j] = r;
        const int prev = col4row[r];
        col4row[r] = j;
        j = prev;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }

  // ---- emit (query index within the video, global target index), query ascending
  const int m0 = a.match_off[p];
  int64_t* po = a.pred_idx + static_cast<size_t>(layer) * a.match_off[P] + m0;
  int64_t* to = a.tgt_idx + static_cast<size_t>(layer) * a.match_off[P] + m0;
  if (!transposed) {
    for (int r = lane; r < nr; r += 32) { po[r] = static_cast<int64_t>(local) * nrows + r; to[r] = t0 + col4row[r]; }
  } else {
    // working columns are the original rows (queries); compact the assigned ones in order
    int written = 0;
    for (int base = 0; base < nc; base += 32) {
      const int j = base + lane;
      const bool has = j < nc && row4col[j] != -1;
      const unsigned m = __ballot_sync(0xffffffffu, has);
      if (has) {
        const int k = written + __popc(m & ((1u << lane) - 1));
        po[k] = static_cast<int64_t>(local) * nrows + j;
        to[k] = t0 + row4col[j];
      }
      written += __popc(m);
    }
  }
}

int launch_match(const MatchArgs& a, cudaStream_t stream) {
  if (a.NL <= 0 || a.B <= 0 || a.Q <= 0 || a.problems_per_video <= 0 || a.rows_per_problem <= 0 ||
      a.problems_per_video * a.rows_per_problem != a.Q)      // matcher.py:56
    return svol_fail(SVOL_ERR_SHAPE, "match: Q must equal problems_per_video * rows_per_problem");
  const int max_small = a.rows_per_problem < a.max_cols ? a.rows_per_problem : a.max_cols;
  const int max_big = a.rows_per_problem > a.max_cols ? a.rows_per_problem : a.max_cols;
  const int ms = (max_small + 1) & ~1, mb = (max_big + 1) & ~1;     // keep the int arrays 8-byte aligned
  const size_t smem_solver = (sizeof(double) * (ms + 2 * mb) + sizeof(int) * (3 * mb + ms) + (ms + mb) + 15) & ~static_cast<size_t>(15);
  if (smem_solver > 200 * 1024) return svol_fail(SVOL_ERR_SHAPE, "match: problem too large for one warp's shared memory");
  // cost block in shared memory when it fits in ~64 KB per warp (keeps >= 3 problems resident per SM)
  const size_t cost_bytes = static_cast<size_t>(a.rows_per_problem) * static_cast<size_t>(a.max_cols) * sizeof(float);
  const int cost_smem_floats = cost_bytes <= 64 * 1024 ? a.rows_per_problem * a.max_cols : 0;
  const size_t smem = smem_solver + static_cast<size_t>(cost_smem_floats) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "match: cudaFuncSetAttribute");
    configured = smem;
  }
  const int P = a.B * a.problems_per_video;
  match_kernel<<<a.NL * P, 32, smem, stream>>>(a, ms, mb, cost_smem_floats);
  return svol_check_launch("match");
}

// matcher.py:114-115: tgt_idx -= min(tgt_idx) per video (and per layer)
__global__ void __launch_bounds__(32) match_localize_kernel(int64_t* tgt_idx, const int32_t* video_match_off, int B, int K) {
  const int layer = blockIdx.x / B, b = blockIdx.x - layer * B, lane = threadIdx.x;
  const int k0 = video_match_off[b], k1 = video_match_off[b + 1];
  int64_t* t = tgt_idx + static_cast<size_t>(layer) * K;
  long long mn = LLONG_MAX;
  for (int k = k0 + lane; k < k1; k += 32) mn = min(mn, static_cast<long long>(t[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  for (int k = k0 + lane; k < k1; k += 32) t[k] -= mn;
}

int launch_match_localize(int64_t* tgt_idx, const int32_t* video_match_off, int NL, int B, int K, cudaStream_t stream) {
  if (NL <= 0 || B <= 0) return svol_fail(SVOL_ERR_SHAPE, "match_localize: bad sizes");
  match_localize_kernel<<<NL * B, 32, 0, stream>>>(tgt_idx, video_match_off, B, K);
  return svol_check_launch("match_localize");
}

}  // namespace svol
