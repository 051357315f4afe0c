// Batched on-GPU Hungarian matching for every decoder layer in one launch (sm_100a).
//
// Replaces PerFrameMatcher.forward / HungarianMatcher.forward (lib/modeling/matcher.py:38-119,
// 131-159): the reference builds the full cross-batch (B*Q) x sum(n) cost matrix on the device,
// copies it to the host and calls scipy.optimize.linear_sum_assignment once per frame.  Here one
// warp owns one assignment problem: it forms only that problem's cost block (the block-diagonal
// entries the reference reads, matcher.py:92-93) in shared memory and solves it in place.
//
// Bit-exactness: the cost is evaluated in fp32 with the reference's operation order and with
// explicit round-to-nearest intrinsics so the compiler cannot contract a*b+c into an FMA
// (matcher.py:59-85; box_utils.py:9-13,24-37,55-61).  The solver is the shortest-augmenting-path
// algorithm scipy uses (Crouse 2016) on costs promoted to fp64, including its candidate order
// (the `remaining` list: columns last-to-first, the last entry moved into a removed slot), its
// preference for unassigned columns among equal minima, the transpose for tall problems and rows
// returned in ascending order.
//
// Solver layout (round 2): the per-column state of a problem (dual v, shortest-path cost, position
// in scipy's `remaining` list, scanned / free flags) lives in REGISTERS, column j on lane j % 32 --
// up to 10 columns per lane (320 columns); only what is addressed by row (u, col4row) or walked
// sequentially (path, row4col) stays in shared memory.  One path step is then: one conflict-free
// shared-memory load of the cost row per owned column, three fp64 adds, and THREE `redux.sync`
// warp reductions -- the fp64 minimum as two 32-bit halves of an order-preserving key, then one
// packed (free?, position) key that encodes scipy's tie rule -- instead of 25 shuffles and seven
// dependent shared-memory round trips.  Problems wider than 320 columns keep the shared-memory
// state solver (`lsap_warp_smem`), which is also selectable for tests.
#include <math_constants.h>

#include <climits>

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
__device__ __forceinline__ int warp_min_i(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max_i(int v) { return __reduce_max_sync(0xffffffffu, v); }

// ------------------------------------------------------------------------------------------------------------------
// Shared-memory workspace of one warp's problem.  `ms` / `mb` = even upper bounds of min / max(rows, cols).
struct LsapSmem {
  double* u;        // [ms]   row duals
  double* v;        // [mb]   (shared-memory-state solver only)
  double* spc;      // [mb]   (shared-memory-state solver only)
  int* col4row;     // [ms]
  int* row4col;     // [mb]
  int* path;        // [mb]
  int* remaining;   // [mb]   (shared-memory-state solver only)
  uint8_t* SR;      // [ms]   (shared-memory-state solver only)
  uint8_t* SC;      // [mb]   (shared-memory-state solver only)
};
__host__ __device__ inline size_t lsap_smem_bytes(int ms, int mb) {
  return (sizeof(double) * (ms + 2 * mb) + sizeof(int) * (3 * mb + ms) + (ms + mb) + 15) & ~static_cast<size_t>(15);
}
__device__ __forceinline__ LsapSmem lsap_carve(uint8_t* raw, int ms, int mb) {
  LsapSmem s;
  s.u = reinterpret_cast<double*>(raw);
  s.v = s.u + ms;
  s.spc = s.v + mb;
  s.col4row = reinterpret_cast<int*>(s.spc + mb);
  s.row4col = s.col4row + ms;
  s.path = s.row4col + mb;
  s.remaining = s.path + mb;
  s.SR = reinterpret_cast<uint8_t*>(s.remaining + mb);
  s.SC = s.SR + ms;
  return s;
}

// Order-preserving 64-bit key of a double (no NaN; -0 folded into +0 first).
__device__ __forceinline__ unsigned long long ordered_key(double x) {
  const long long b = __double_as_longlong(x + 0.0);
  return static_cast<unsigned long long>(b) ^ (static_cast<unsigned long long>(b >> 63) | 0x8000000000000000ull);
}

// Register-state solver.  C: cost block of the WORKING problem (nr <= nc <= 32 * CPL), entry (i, j) at C[i * si + j * sj]
// (shared memory: si = nc, sj = 1; a tall block read in place from the global workspace: si = 1, sj = its row pitch).
// Returns 0, or 2 when the problem is infeasible (every remaining entry +inf).  On return col4row / row4col hold the
// assignment of the working problem.
template <int CPL>
__device__ __forceinline__ int lsap_warp_regs(const float* __restrict__ C, int si, int sj, int nr, int nc, const LsapSmem& s,
                                              int lane) {
  double v[CPL], spc[CPL], uo[CPL];   // per owned column: dual, shortest-path cost, u of the row it is assigned to
  int pos[CPL], own[CPL];             // position in scipy's `remaining` list; row4col (constant during one search)
#pragma unroll
  for (int k = 0; k < CPL; ++k) { v[k] = 0.0; uo[k] = 0.0; own[k] = -1; }
  unsigned valid = 0;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int j = lane + 32 * k;
    if (j < nc) { valid |= 1u << k; s.row4col[j] = -1; s.path[j] = -1; }
  }
  for (int i = lane; i < nr; i += 32) { s.u[i] = 0.0; s.col4row[i] = -1; }
  __syncwarp();

  for (int cur = 0; cur < nr; ++cur) {
    unsigned scm = 0;                                   // scanned columns (SC) among mine
#pragma unroll
    for (int k = 0; k < CPL; ++k) { spc[k] = CUDART_INF; pos[k] = nc - 1 - (lane + 32 * k); }   // remaining[t] = nc - t - 1
    int n_rem = nc, i = cur, sink = -1;
    double best = 0.0, ui = s.u[cur];
    while (sink == -1) {
      const float* crow = C + static_cast<size_t>(i) * si;
      // branch-free over the owned columns, so that their load -> three fp64 adds -> compare chains overlap
      float c[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) c[k] = crow[static_cast<size_t>(min(lane + 32 * k, nc - 1)) * sj];
      double lw = CUDART_INF;
      unsigned kw = 0;
      const unsigned live = valid & ~scm;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int j = lane + 32 * k;
        const bool on = (live >> k) & 1u;
        const double r = ((best + static_cast<double>(c[k])) - ui) - v[k];
        const bool upd = on && r < spc[k];
        if (upd) s.path[j] = i;
        spc[k] = upd ? r : spc[k];
        const double sv = on ? spc[k] : CUDART_INF;
        // scipy: a strictly lower value wins; among equal values a free column found later in `remaining` order wins.
        // As a set function: the maximum over the minima of (free, free ? position : -position); the column index rides
        // in the low 10 bits (positions are unique, so it never decides).
        const unsigned p = static_cast<unsigned>(pos[k]);
        const unsigned key = on ? ((own[k] < 0 ? (0x80000000u | (p << 10)) : ((0x1fffffu - p) << 10)) | static_cast<unsigned>(j)) : 0u;
        const bool take = sv < lw || (sv == lw && key > kw);
        lw = take ? sv : lw;
        kw = take ? key : kw;
      }
      // fp64 minimum as two 32-bit reductions of an order-preserving key, then the tie rule as a third
      const unsigned long long ok = ordered_key(lw);
      const unsigned hi = static_cast<unsigned>(ok >> 32), lo = static_cast<unsigned>(ok);
      const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
      const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
      const unsigned kmax = __reduce_max_sync(0xffffffffu, (hi == mh && lo == ml) ? kw : 0u);
      if (mh == 0xfff00000u && ml == 0u) return 2;      // minimum is +inf: infeasible
      const unsigned long long okm = (static_cast<unsigned long long>(mh) << 32) | ml;
      best = __longlong_as_double(static_cast<long long>((okm >> 63) ? (okm ^ 0x8000000000000000ull) : ~okm));
      const int j = static_cast<int>(kmax & 0x3ffu);
      const unsigned pf = (kmax >> 10) & 0x1fffffu;
      const int pick = (kmax & 0x80000000u) ? static_cast<int>(pf) : static_cast<int>(0x1fffffu - pf);
      // owner row of column j and its dual, from the registers of the lane that holds the column
      const int slot = j >> 5;
      int own_s = own[0];
      double uo_s = uo[0];
#pragma unroll
      for (int k = 1; k < CPL; ++k) { own_s = slot == k ? own[k] : own_s; uo_s = slot == k ? uo[k] : uo_s; }
      const int owner = __shfl_sync(0xffffffffu, own_s, j & 31);
      ui = __shfl_sync(0xffffffffu, uo_s, j & 31);
      if (lane == (j & 31)) scm |= 1u << slot;
      --n_rem;                                          // remaining[pick] = remaining[--n_rem]
#pragma unroll
      for (int k = 0; k < CPL; ++k) pos[k] = pos[k] == n_rem ? pick : pos[k];
      if (owner == -1) sink = j; else i = owner;
    }
    // dual update: rows reached through a scanned column (their col4row is that column), then the scanned columns
    if (lane == 0) s.u[cur] += best;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      if ((scm >> k) & 1u) {
        const int j = lane + 32 * k;
        const double dlt = best - spc[k];
        if (j != sink) s.u[own[k]] += dlt;
        v[k] -= dlt;
      }
    }
    __syncwarp();
    // augment along the path (sequential)
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = s.path[j];
        s.row4col[j] = r;
        const int prev = s.col4row[r];
        s.col4row[r] = j;
        j = prev;
        if (r == cur) break;
      }
    }
    __syncwarp();
    // refresh the per-column copies of row4col / u[row4col] for the next search
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      if ((valid >> k) & 1u) {
        own[k] = s.row4col[lane + 32 * k];
        uo[k] = own[k] >= 0 ? s.u[own[k]] : 0.0;
      }
    }
  }
  return 0;
}

// Shared-memory-state solver (any width).  Same algorithm, all state in shared memory.
__device__ __noinline__ int lsap_warp_smem(const float* __restrict__ C, int si, int sj, int nr, int nc, const LsapSmem& s,
                                           int lane) {
  double* u = s.u; double* v = s.v; double* spc = s.spc;
  int* path = s.path; int* col4row = s.col4row; int* row4col = s.row4col; int* remaining = s.remaining;
  uint8_t* SR = s.SR; uint8_t* SC = s.SC;
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
        const double r = ((best + static_cast<double>(C[static_cast<size_t>(i) * si + static_cast<size_t>(j) * sj])) - ui) - v[j];
        double sv = spc[j];
        if (r < sv) { path[j] = i; spc[j] = r; sv = r; }
        const bool free_col = row4col[j] == -1;
        if (sv < lowest) { lowest = sv; first_pos = t; last_free = free_col ? t : -1; }
        else if (sv == lowest && free_col) last_free = t;
      }
      const double gl = warp_min_d(lowest);
      const bool mine = lowest == gl && first_pos != INT_MAX;
      const int fp = warp_min_i(mine ? first_pos : INT_MAX);
      const int lf = warp_max_i(mine ? last_free : -1);
      if (!(gl < CUDART_INF)) return 2;
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
    for (int r = lane; r < nr; r += 32) {
      if (r == cur) u[r] += best;
      else if (SR[r]) u[r] += best - spc[col4row[r]];
    }
    for (int j = lane; j < nc; j += 32)
      if (SC[j]) v[j] -= best - spc[j];
    __syncwarp();
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
  return 0;
}

__device__ __forceinline__ int lsap_dispatch(const float* C, int si, int sj, int nr, int nc, const LsapSmem& s, int lane,
                                             bool force_smem_state) {
  if (force_smem_state || nc > 320) return lsap_warp_smem(C, si, sj, nr, nc, s, lane);
  if (nc <= 32) return lsap_warp_regs<1>(C, si, sj, nr, nc, s, lane);
  if (nc <= 64) return lsap_warp_regs<2>(C, si, sj, nr, nc, s, lane);
  if (nc <= 128) return lsap_warp_regs<4>(C, si, sj, nr, nc, s, lane);
  return lsap_warp_regs<10>(C, si, sj, nr, nc, s, lane);
}

// Emits the assignment of the ORIGINAL problem (rows ascending, like scipy) through `put(k, row, col)`.
template <typename Put>
__device__ __forceinline__ void lsap_emit(bool transposed, int nr, int nc, const LsapSmem& s, int lane, Put put) {
  if (!transposed) {
    for (int r = lane; r < nr; r += 32) put(r, r, s.col4row[r]);
  } else {
    // working columns are the original rows; compact the assigned ones in order
    int written = 0;
    for (int base = 0; base < nc; base += 32) {
      const int j = base + lane;
      const bool has = j < nc && s.row4col[j] != -1;
      const unsigned m = __ballot_sync(0xffffffffu, has);
      if (has) put(written + __popc(m & ((1u << lane) - 1)), j, s.row4col[j]);
      written += __popc(m);
    }
  }
}

// Per-box quantities of the cost formula that do not depend on the partner box, staged once per problem in shared
// memory (structure of arrays: consecutive lanes read consecutive boxes).  Computed with the same fp32 operations as
// pair_cost(), so a cost assembled from them is bit-identical.
constexpr int QF = 10, TF = 9;     // floats per query (p_fg, cx, cy, w, h, x0, y0, x1, y1, area) / per target (no p_fg)
__device__ __forceinline__ void box_derive(const float4& b, float* dst, int pitch) {
  dst[0 * pitch] = b.x; dst[1 * pitch] = b.y; dst[2 * pitch] = b.z; dst[3 * pitch] = b.w;
  const float x0 = __fsub_rn(b.x, __fmul_rn(0.5f, b.z)), y0 = __fsub_rn(b.y, __fmul_rn(0.5f, b.w));
  const float x1 = __fadd_rn(b.x, __fmul_rn(0.5f, b.z)), y1 = __fadd_rn(b.y, __fmul_rn(0.5f, b.w));
  dst[4 * pitch] = x0; dst[5 * pitch] = y0; dst[6 * pitch] = x1; dst[7 * pitch] = y1;
  dst[8 * pitch] = __fmul_rn(__fsub_rn(x1, x0), __fsub_rn(y1, y0));
}
__device__ __forceinline__ float staged_cost(const float* q, int qp, int r, const float* t, int tp, int c, float w_class,
                                             float w_bbox, float w_giou) {
  const float p_fg = q[r];
  const float* qb = q + qp + r;      // box fields of query r start after the p_fg plane
  const float* tb = t + c;
  float l1 = fabsf(__fsub_rn(qb[0], tb[0]));
  l1 = __fadd_rn(l1, fabsf(__fsub_rn(qb[qp], tb[tp])));
  l1 = __fadd_rn(l1, fabsf(__fsub_rn(qb[2 * qp], tb[2 * tp])));
  l1 = __fadd_rn(l1, fabsf(__fsub_rn(qb[3 * qp], tb[3 * tp])));
  const float ax0 = qb[4 * qp], ay0 = qb[5 * qp], ax1 = qb[6 * qp], ay1 = qb[7 * qp], area_a = qb[8 * qp];
  const float tx0 = tb[4 * tp], ty0 = tb[5 * tp], tx1 = tb[6 * tp], ty1 = tb[7 * tp], area_t = tb[8 * tp];
  const float iw = fmaxf(__fsub_rn(fminf(ax1, tx1), fmaxf(ax0, tx0)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(ay1, ty1), fmaxf(ay0, ty0)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_t), inter);
  const float iou = __fdiv_rn(inter, uni);
  const float cw = fmaxf(__fsub_rn(fmaxf(ax1, tx1), fminf(ax0, tx0)), 0.f);
  const float ch = fmaxf(__fsub_rn(fmaxf(ay1, ty1), fminf(ay0, ty0)), 0.f);
  const float hull = __fmul_rn(cw, ch);
  const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
  return __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_giou, -giou)), __fmul_rn(w_class, -p_fg));
}

// One CTA of 1, 2 or 4 warps per (layer, problem): all of them stage the problem's boxes and fill its cost block, warp 0
// solves (registers are allocated per thread, so extra fill warps cost solver occupancy: launch_match picks the width).
// Problems are taken in the caller's `order` (largest first: the launch ends with the short ones) when one is given.
// mode 0: cost block + solve; 1: cost blocks only (into cost_ws); 2: solve from cost_ws.
// Dynamic shared memory: solver state (lsap_smem_bytes) | staged boxes (QF * rows + TF * max_cols floats) | the cost
// block in the solver's working orientation when it fits (cost_smem_floats > 0); larger blocks live in the global
// workspace, also in working orientation (transposed when the problem is tall) so that the solver's row reads coalesce.
template <int MATCH_THREADS>
__global__ void __launch_bounds__(MATCH_THREADS) match_kernel(const MatchArgs a, int ms, int mb, int cost_smem_floats) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.B * a.problems_per_video;
  const int slot_p = blockIdx.x / a.NL, layer = blockIdx.x - slot_p * a.NL;     // layers of one problem run side by side
  const int p = a.order ? a.order[slot_p] : slot_p;
  const int video = p / a.problems_per_video, local = p - video * a.problems_per_video;
  const int nrows = a.rows_per_problem;
  const int t0 = a.tgt_off[p], ncols = a.tgt_off[p + 1] - t0;
  if (ncols <= 0 || nrows <= 0) return;
  int32_t* status_acc = a.status + 1;

  const LsapSmem s = lsap_carve(sm_raw, ms, mb);
  const int qp = (nrows + 3) & ~3, tp = (a.max_cols + 3) & ~3;           // plane pitches of the staged boxes
  float* qs = reinterpret_cast<float*>(sm_raw + lsap_smem_bytes(ms, mb));
  float* ts = qs + QF * qp;
  float* Cs = ts + TF * tp;
  const bool cost_in_smem = nrows * ncols <= cost_smem_floats;
  const bool transposed = ncols < nrows;          // scipy transposes tall problems
  const int nr = transposed ? ncols : nrows;      // rows of the working problem
  const int nc = transposed ? nrows : ncols;

  const int m0 = a.match_off[p];
  int64_t* po = a.pred_idx + static_cast<size_t>(layer) * a.K + m0;       // a.K: row pitch of the index arrays (>= match_off[P])
  int64_t* to = a.tgt_idx + static_cast<size_t>(layer) * a.K + m0;
  // indices that downstream kernels can always dereference (row r <-> column min(r, ncols - 1)): written whenever the
  // problem has no solution, next to the status bit -- the criterion must never see uninitialised indices
  auto safe_indices = [&]() {
    for (int r = lane; r < nr; r += 32) {
      po[r] = static_cast<int64_t>(local) * nrows + r;
      to[r] = t0 + (r < ncols ? r : ncols - 1);
    }
  };

  // ---- cost block
  const size_t q0 = (static_cast<size_t>(layer) * a.B + video) * a.Q + static_cast<size_t>(local) * nrows;
  float* Cg = a.cost_ws ? a.cost_ws + static_cast<size_t>(layer) * a.cost_off[P] + a.cost_off[p] : nullptr;
  const bool ws_working = cost_smem_floats == 0;  // workspace blocks in working orientation (see above)
  bool bad = false;
  if (a.mode != 2) {
    for (int r = tid; r < nrows; r += MATCH_THREADS) {
      const float2 lg = reinterpret_cast<const float2*>(a.logits)[q0 + r];
      qs[r] = fg_prob(lg.x, lg.y);
      box_derive(reinterpret_cast<const float4*>(a.boxes)[q0 + r], qs + qp + r, qp);
    }
    for (int c = tid; c < ncols; c += MATCH_THREADS)
      box_derive(reinterpret_cast<const float4*>(a.tgt_boxes)[t0 + c], ts + c, tp);
    if (MATCH_THREADS > 32) __syncthreads(); else __syncwarp();
    // entries visited in working order (conflict-free shared-memory stores, coalesced workspace stores)
    for (int e = tid; e < nr * nc; e += MATCH_THREADS) {
      const int i = e / nc, j = e - i * nc;
      const int r = transposed ? j : i, c = transposed ? i : j;
      const float cost = staged_cost(qs, qp, r, ts, tp, c, a.w_class, a.w_bbox, a.w_giou);
      bad |= (cost != cost) || (cost == -CUDART_INF_F);
      if (Cg) Cg[ws_working ? e : r * ncols + c] = cost;
      if (cost_in_smem) Cs[e] = cost;
    }
    if (a.mode == 1) return;
  } else {
    for (int e = tid; e < nr * nc; e += MATCH_THREADS) {
      const int i = e / nc, j = e - i * nc;
      const float cost = Cg[(ws_working || !transposed) ? e : j * ncols + i];
      bad |= (cost != cost) || (cost == -CUDART_INF_F);
      if (cost_in_smem) Cs[e] = cost;
    }
  }
  bool any_bad;
  if (MATCH_THREADS > 32) {
    any_bad = __syncthreads_or(bad);
    if (warp != 0) return;                 // the assignment itself is one warp's sequential work
  } else {
    any_bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
  }
  if (any_bad) {                           // scipy: "matrix contains invalid numeric entries"
    if (lane == 0) atomicOr(status_acc, 1);
    safe_indices();
    return;
  }

  const int rc = lsap_dispatch(cost_in_smem ? Cs : Cg, nc, 1, nr, nc, s, lane, a.solver == 1);
  if (rc != 0) {                           // infeasible: every candidate +inf
    if (lane == 0) atomicOr(status_acc, 2);
    safe_indices();
    return;
  }
  // ---- emit (query index within the video, global target index), query ascending
  lsap_emit(transposed, nr, nc, s, lane,
            [&](int k, int r, int c) { po[k] = static_cast<int64_t>(local) * nrows + r; to[k] = t0 + c; });
}

// After match_kernel on the same stream: (1) global -> video-local target indices.  localize 1 = PerFrameMatcher's quirk
// (matcher.py:114-115): subtract the minimum matched global index of the video; 2 = HungarianMatcher (:158): columns are
// local to the video's split, i.e. global minus the video's first target.  (2) publish the status word of this call and
// clear the accumulator for the next one (status[0] = status[1]; status[1] = 0), so no host-side memset is needed and the
// whole matcher can sit inside a captured CUDA graph.
__global__ void __launch_bounds__(32) match_finalize_kernel(int64_t* tgt_idx, const int32_t* video_match_off,
                                                            const int32_t* video_tgt_off, int32_t* status, int B, int K,
                                                            int localize) {
  const int layer = blockIdx.x / B, b = blockIdx.x - layer * B, lane = threadIdx.x;
  if (blockIdx.x == 0 && lane == 0 && status) { status[0] = status[1]; status[1] = 0; }
  if (!localize) return;
  const int k0 = video_match_off[b], k1 = video_match_off[b + 1];
  int64_t* t = tgt_idx + static_cast<size_t>(layer) * K;
  long long mn = LLONG_MAX;
  if (localize == 1) {
    for (int k = k0 + lane; k < k1; k += 32) mn = min(mn, static_cast<long long>(t[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  } else {
    mn = video_tgt_off[b];
  }
  for (int k = k0 + lane; k < k1; k += 32) t[k] -= mn;
}

static int match_smem(const MatchArgs& a, int* ms_out, int* mb_out, int* cost_floats_out, size_t* smem_out) {
  const int max_small = a.rows_per_problem < a.max_cols ? a.rows_per_problem : a.max_cols;
  const int max_big = a.rows_per_problem > a.max_cols ? a.rows_per_problem : a.max_cols;
  const int ms = (max_small + 1) & ~1, mb = (max_big + 1) & ~1;     // keep the int arrays 8-byte aligned
  const size_t smem_solver = lsap_smem_bytes(ms, mb) +
      sizeof(float) * (QF * ((a.rows_per_problem + 3) & ~3) + TF * ((a.max_cols + 3) & ~3));     // + staged boxes
  if (smem_solver > 120 * 1024) return svol_fail(SVOL_ERR_SHAPE, "match: problem too large for one CTA's shared memory");
  // cost block in shared memory when it fits next to the solver state (<= 96 KB keeps >= 2 problems resident per SM)
  const size_t cost_bytes = static_cast<size_t>(a.rows_per_problem) * static_cast<size_t>(a.max_cols) * sizeof(float);
  const bool fits = cost_bytes <= 96 * 1024 && smem_solver + cost_bytes <= 200 * 1024;
  if (!fits && a.cost_ws == nullptr)
    return svol_fail(SVOL_ERR_NULL, "match: problems this large need the global cost workspace (cost_ws)");
  *ms_out = ms; *mb_out = mb;
  *cost_floats_out = fits ? a.rows_per_problem * a.max_cols : 0;
  *smem_out = smem_solver + (fits ? cost_bytes : 0);
  return SVOL_OK;
}

int launch_match(const MatchArgs& a, cudaStream_t stream) {
  if (a.NL <= 0 || a.B <= 0 || a.Q <= 0 || a.problems_per_video <= 0 || a.rows_per_problem <= 0 ||
      a.problems_per_video * a.rows_per_problem != a.Q)      // matcher.py:56
    return svol_fail(SVOL_ERR_SHAPE, "match: Q must equal problems_per_video * rows_per_problem");
  if (a.mode < 0 || a.mode > 2 || a.solver < 0 || a.solver > 1 || a.localize < 0 || a.localize > 2)
    return svol_fail(SVOL_ERR_SHAPE, "match: mode in 0..2, solver in 0..1, localize in 0..2");
  if (a.K <= 0) return svol_fail(SVOL_ERR_SHAPE, "match: K (row pitch of pred_idx / tgt_idx, >= match_off[P]) must be > 0");
  if (a.mode != 0 && a.cost_ws == nullptr) return svol_fail(SVOL_ERR_NULL, "match: modes 1 and 2 need cost_ws");
  if (a.localize && (a.video_match_off == nullptr || (a.localize == 2 && a.video_tgt_off == nullptr)))
    return svol_fail(SVOL_ERR_NULL, "match: localize needs video_match_off (and video_tgt_off for mode 2)");
  int ms, mb, cost_floats;
  size_t smem;
  if (int rc = match_smem(a, &ms, &mb, &cost_floats, &smem)) return rc;
  const int P = a.B * a.problems_per_video;
  // width of a CTA: one warp for tiny blocks; two when there are many problems (the solver's ~128 registers per thread are
  // allocated for every thread of the CTA, so more fill warps would cut the number of resident solvers); four otherwise
  const long long entries = static_cast<long long>(a.rows_per_problem) * a.max_cols;
  const int threads = entries <= 512 ? 32 : (static_cast<long long>(a.NL) * P >= 4LL * sm_count() ? 64 : 128);
  auto launch = [&](auto kernel, size_t* configured) -> int {
    if (smem > 48 * 1024 && smem > *configured) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return svol_fail_cuda(e, "match: cudaFuncSetAttribute");
      *configured = smem;
    }
    kernel<<<a.NL * P, threads, smem, stream>>>(a, ms, mb, cost_floats);
    return SVOL_OK;
  };
  static size_t conf32 = 0, conf64 = 0, conf128 = 0;
  int lrc;
  if (threads == 32) lrc = launch(match_kernel<32>, &conf32);
  else if (threads == 64) lrc = launch(match_kernel<64>, &conf64);
  else lrc = launch(match_kernel<128>, &conf128);
  if (lrc) return lrc;
  if (int rc = svol_check_launch("match")) return rc;
  if (a.mode == 1) return SVOL_OK;
  match_finalize_kernel<<<a.NL * a.B, 32, 0, stream>>>(a.tgt_idx, a.video_match_off, a.video_tgt_off, a.status, a.B,
                                                        a.K, a.localize);
  return svol_check_launch("match_finalize");
}

// matcher.py:114-115 as a stand-alone call (kept for callers that ran svol_match with localize = 0)
int launch_match_localize(int64_t* tgt_idx, const int32_t* video_match_off, int NL, int B, int K, cudaStream_t stream) {
  if (NL <= 0 || B <= 0) return svol_fail(SVOL_ERR_SHAPE, "match_localize: bad sizes");
  match_finalize_kernel<<<NL * B, 32, 0, stream>>>(tgt_idx, video_match_off, nullptr, nullptr, B, K, 1);
  return svol_check_launch("match_localize");
}

// ------------------------------------------------------------------------------------------------------------------
// Stand-alone batched LSAP on caller-supplied fp32 cost matrices: scipy.optimize.linear_sum_assignment semantics
// (matcher.py:93,158), one warp per problem, the same solver code as match_kernel.
__global__ void __launch_bounds__(32) lsap_kernel(const float* __restrict__ cost, const int64_t* __restrict__ cost_off,
                                                  const int32_t* __restrict__ shape, int64_t* __restrict__ rows_out,
                                                  int64_t* __restrict__ cols_out, const int64_t* __restrict__ out_off,
                                                  int32_t* __restrict__ status, int ms, int mb, int cost_smem_floats,
                                                  int solver) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int lane = threadIdx.x, p = blockIdx.x;
  const int nrows = shape[2 * p], ncols = shape[2 * p + 1];
  if (lane == 0) status[p] = 0;
  if (nrows <= 0 || ncols <= 0) return;
  const LsapSmem s = lsap_carve(sm_raw, ms, mb);
  float* Cs = reinterpret_cast<float*>(sm_raw + lsap_smem_bytes(ms, mb));
  const bool transposed = ncols < nrows;
  const int nr = transposed ? ncols : nrows, nc = transposed ? nrows : ncols;
  const float* Cg = cost + cost_off[p];
  bool bad = false;
  for (int e = lane; e < nr * nc; e += 32) {
    const int i = e / nc, j = e - i * nc;
    const float c = Cg[transposed ? j * ncols + i : e];
    bad |= (c != c) || (c == -CUDART_INF_F);
    Cs[e] = c;
  }
  int64_t* ro = rows_out + out_off[p];
  int64_t* co = cols_out + out_off[p];
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) status[p] = 1;
    return;
  }
  __syncwarp();
  (void)cost_smem_floats;
  const int rc = lsap_dispatch(Cs, nc, 1, nr, nc, s, lane, solver == 1);
  if (rc != 0) {
    if (lane == 0) status[p] = 2;
    return;
  }
  lsap_emit(transposed, nr, nc, s, lane, [&](int k, int r, int c) { ro[k] = r; co[k] = c; });
}

int launch_lsap_f32(const float* cost, const int64_t* cost_off, const int32_t* shape, int n_problems, int max_small,
                    int max_big, int max_entries, int64_t* rows_out, int64_t* cols_out, const int64_t* out_off, int32_t* status,
                    int solver, cudaStream_t stream) {
  if (n_problems <= 0 || max_small <= 0 || max_big < max_small || max_entries <= 0 || solver < 0 || solver > 1)
    return svol_fail(SVOL_ERR_SHAPE, "lsap: n_problems, max_small <= max_big, max_entries > 0; solver in 0..1");
  const int ms = (max_small + 1) & ~1, mb = (max_big + 1) & ~1;
  const size_t cost_bytes = static_cast<size_t>(max_entries) * sizeof(float);
  const size_t smem = lsap_smem_bytes(ms, mb) + cost_bytes;
  if (smem > 200 * 1024) return svol_fail(SVOL_ERR_SHAPE, "lsap: max_entries * 4 bytes must fit in shared memory");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "lsap: cudaFuncSetAttribute");
    configured = smem;
  }
  lsap_kernel<<<n_problems, 32, smem, stream>>>(cost, cost_off, shape, rows_out, cols_out, out_off, status, ms, mb,
                                                max_entries, solver);
  return svol_check_launch("lsap");
}

}  // namespace svol
