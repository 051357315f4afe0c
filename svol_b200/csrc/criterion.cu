// SetCriterion losses for every decoder layer in one launch (sm_100a), forward and backward.
//
// Replaces SetCriterion.loss_labels / loss_boxes (lib/modeling/loss.py:39-60,76-103), which the
// reference evaluates per layer with ~20 small PyTorch kernels, an advanced-indexing scatter for the
// target classes and a K x K generalized_box_iou whose diagonal is the only part used.  One CTA per
// decoder layer: matched (video, query) pairs are marked in a shared-memory bitmap, the weighted
// cross-entropy runs over all B*Q logits with coalesced float2 loads, the matched pairs are gathered
// for L1 / GIoU, and fp64 block reductions produce the four scalars.
#include <climits>

#include "common.cuh"
#include "svol_internal.h"

namespace svol {

constexpr int CRIT_THREADS = 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}

struct PairGeom {
  float x0, y0, x1, y1, tx0, ty0, tx1, ty1, iw, ih, inter, uni, cw, ch, hull, giou;
};

__device__ __forceinline__ PairGeom pair_geom(const float4& s, const float4& t) {
  PairGeom g;
  g.x0 = s.x - 0.5f * s.z; g.y0 = s.y - 0.5f * s.w; g.x1 = s.x + 0.5f * s.z; g.y1 = s.y + 0.5f * s.w;
  g.tx0 = t.x - 0.5f * t.z; g.ty0 = t.y - 0.5f * t.w; g.tx1 = t.x + 0.5f * t.z; g.ty1 = t.y + 0.5f * t.w;
  const float area_s = (g.x1 - g.x0) * (g.y1 - g.y0), area_t = (g.tx1 - g.tx0) * (g.ty1 - g.ty0);
  g.iw = fmaxf(fminf(g.x1, g.tx1) - fmaxf(g.x0, g.tx0), 0.f);
  g.ih = fmaxf(fminf(g.y1, g.ty1) - fmaxf(g.y0, g.ty0), 0.f);
  g.inter = g.iw * g.ih;
  g.uni = area_s + area_t - g.inter;
  g.cw = fmaxf(fmaxf(g.x1, g.tx1) - fminf(g.x0, g.tx0), 0.f);
  g.ch = fmaxf(fmaxf(g.y1, g.ty1) - fminf(g.y0, g.ty0), 0.f);
  g.hull = g.cw * g.ch;
  g.giou = g.inter / g.uni - (g.hull - g.uni) / g.hull;
  return g;
}

// Sizes that may live on the device (meta[0] = K of the batch in the static target buffer) so that one captured launch
// serves every batch; the index arrays then have a fixed row pitch.
struct CritDims { int K, pitch, S; };
__device__ __forceinline__ CritDims crit_dims(const CriterionArgs& a) {
  CritDims d;
  d.K = a.meta ? a.meta[0] : a.K;
  d.pitch = a.idx_pitch > 0 ? a.idx_pitch : d.K;
  d.S = a.video_tgt_off[a.B];
  return d;
}
// video of matched pair k: the caller's table, or a binary search in the per-video output ranges
__device__ __forceinline__ int pair_video(const CriterionArgs& a, int k) {
  if (a.match_video) return a.match_video[k];
  int lo = 0, hi = a.B;               // video_match_off[lo] <= k < video_match_off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.video_match_off[mid] <= k) lo = mid; else hi = mid;
  }
  return lo;
}
// Indices are clamped into range: a problem without a solution (NaN costs; reported through the matcher's status word,
// matcher.cu) must not turn into out-of-bounds shared / global accesses here.
__device__ __forceinline__ int pred_entry(const CriterionArgs& a, int b, int64_t q) {
  return b * a.Q + static_cast<int>(q < 0 ? 0 : (q >= a.Q ? a.Q - 1 : q));
}
__device__ __forceinline__ int tgt_entry(const CriterionArgs& a, const CritDims& d, int b, int64_t t) {
  const long long g = static_cast<long long>(a.video_tgt_off[b]) + t;
  return static_cast<int>(g < 0 ? 0 : (g >= d.S ? d.S - 1 : g));
}

// A CTA works on the videos [b0, b1) of one decoder layer: queries [b0*Q, b1*Q), matched pairs
// [video_match_off[b0], video_match_off[b1]) (the whole batch when launched with one CTA per layer).
struct CritRange { int b0, b1, q0, q1, k0, k1; };
__device__ __forceinline__ CritRange crit_range(const CriterionArgs& a, const CritDims& d, bool chunked) {
  CritRange r;
  if (chunked) {
    r.b0 = blockIdx.x; r.b1 = r.b0 + 1;
    r.k0 = a.video_match_off[r.b0]; r.k1 = a.video_match_off[r.b1];
  } else {
    r.b0 = 0; r.b1 = a.B; r.k0 = 0; r.k1 = d.K;
  }
  r.q0 = r.b0 * a.Q; r.q1 = r.b1 * a.Q;
  return r;
}

__device__ __forceinline__ void mark_matched(uint32_t* bitmap, const CriterionArgs& a, const CritDims& d, const CritRange& r,
                                             bool chunked, int layer) {
  const int words = (r.q1 - r.q0 + 31) >> 5;
  for (int i = threadIdx.x; i < words; i += blockDim.x) bitmap[i] = 0u;
  __syncthreads();
  const int64_t* pi = a.pred_idx + static_cast<size_t>(layer) * d.pitch;
  for (int k = r.k0 + threadIdx.x; k < r.k1; k += blockDim.x) {
    const int e = pred_entry(a, chunked ? r.b0 : pair_video(a, k), pi[k]) - r.q0;
    atomicOr(&bitmap[e >> 5], 1u << (e & 31));
  }
  __syncthreads();
}

// Launched as (1, NL) x 1024 threads (one CTA per layer), or -- with a scratch buffer -- as (B, NL) x 256: one CTA per
// (video, layer) writes four fp64 partial sums, the last CTA of a layer to arrive adds them up in video order
// (deterministic) and resets the arrival counter for the next launch.
__global__ void __launch_bounds__(CRIT_THREADS) criterion_kernel(const CriterionArgs a) {
  extern __shared__ uint32_t bitmap[];
  __shared__ double red[32];
  __shared__ int is_last;
  const bool chunked = gridDim.x > 1;
  const int layer = blockIdx.y, tid = threadIdx.x;
  const int n = a.B * a.Q;
  const CritDims d = crit_dims(a);
  const CritRange r = crit_range(a, d, chunked);
  mark_matched(bitmap, a, d, r, chunked, layer);

  // weighted cross-entropy over every query (loss.py:50-55): sum(w * nll) / (B*Q)
  const float2* lg = reinterpret_cast<const float2*>(a.logits) + static_cast<size_t>(layer) * n;
  double ce = 0.0;
  for (int e = r.q0 + tid; e < r.q1; e += blockDim.x) {
    const float2 l = __ldg(lg + e);
    const int w = e - r.q0;
    const bool fg = (bitmap[w >> 5] >> (w & 31)) & 1u;
    const float m = fmaxf(l.x, l.y);
    const float lse = m + logf(expf(l.x - m) + expf(l.y - m));
    const float nll = lse - (fg ? l.x : l.y);
    ce += static_cast<double>((fg ? 1.0f : a.eos_coef) * nll);
  }
  // matched pairs: class_error (loss.py:57-59), L1 and GIoU (loss.py:92-102)
  const int64_t* pi = a.pred_idx + static_cast<size_t>(layer) * d.pitch;
  const int64_t* ti = a.tgt_idx + static_cast<size_t>(layer) * d.pitch;
  const float4* bx = reinterpret_cast<const float4*>(a.boxes) + static_cast<size_t>(layer) * n;
  double correct = 0.0, l1 = 0.0, gl = 0.0;
  for (int k = r.k0 + tid; k < r.k1; k += blockDim.x) {
    const int b = chunked ? r.b0 : pair_video(a, k);
    const int e = pred_entry(a, b, pi[k]);
    const float2 l = __ldg(lg + e);
    correct += (l.x >= l.y) ? 1.0 : 0.0;        // top-1 == foreground (index 0 wins ties)
    const float4 s = __ldg(bx + e);
    const float4 t = __ldg(reinterpret_cast<const float4*>(a.tgt_boxes) + tgt_entry(a, d, b, ti[k]));
    l1 += static_cast<double>(fabsf(s.x - t.x) + fabsf(s.y - t.y) + fabsf(s.z - t.z) + fabsf(s.w - t.w));
    gl += static_cast<double>(1.0f - pair_geom(s, t).giou);
  }
  ce = block_sum(ce, red);
  correct = block_sum(correct, red);
  l1 = block_sum(l1, red);
  gl = block_sum(gl, red);
  auto publish = [&](double ce_, double correct_, double l1_, double gl_) {
    float* o = a.losses + layer * 4;
    o[0] = static_cast<float>(ce_ / n);
    o[1] = static_cast<float>(100.0 - correct_ * (100.0 / d.K));
    o[2] = static_cast<float>(l1_ / (4.0 * d.K));
    o[3] = static_cast<float>(gl_ / d.K);
  };
  if (!chunked) {
    if (tid == 0) publish(ce, correct, l1, gl);
    return;
  }
  double* partial = reinterpret_cast<double*>(a.scratch) + (static_cast<size_t>(layer) * a.B) * 4;
  int* counter = reinterpret_cast<int*>(reinterpret_cast<double*>(a.scratch) + static_cast<size_t>(gridDim.y) * a.B * 4) + layer;
  if (tid == 0) {
    double* p = partial + blockIdx.x * 4;
    p[0] = ce; p[1] = correct; p[2] = l1; p[3] = gl;
    __threadfence();
    is_last = atomicAdd(counter, 1) == a.B - 1;
  }
  __syncthreads();
  if (is_last && tid < 32) {
    __threadfence();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (tid < 4) {
      for (int b = 0; b < a.B; ++b) acc[0] += __ldcg(partial + b * 4 + tid);       // lane c sums component c, in video order
    }
    const double c0 = __shfl_sync(0xffffffffu, acc[0], 0), c1 = __shfl_sync(0xffffffffu, acc[0], 1);
    const double c2 = __shfl_sync(0xffffffffu, acc[0], 2), c3 = __shfl_sync(0xffffffffu, acc[0], 3);
    if (tid == 0) {
      publish(c0, c1, c2, c3);
      *counter = 0;
    }
  }
}

// d(sum_layer w_label*loss_label + w_bbox*loss_bbox + w_giou*loss_giou) / d(logits, boxes); same two launch shapes.
__global__ void __launch_bounds__(CRIT_THREADS) criterion_backward_kernel(const CriterionArgs a,
                                                                          const float* __restrict__ grad_w,
                                                                          float* __restrict__ grad_logits,
                                                                          float* __restrict__ grad_boxes) {
  extern __shared__ uint32_t bitmap[];
  const bool chunked = gridDim.x > 1;
  const int layer = blockIdx.y, tid = threadIdx.x;
  const int n = a.B * a.Q;
  const CritDims d = crit_dims(a);
  const CritRange r = crit_range(a, d, chunked);
  mark_matched(bitmap, a, d, r, chunked, layer);
  const float w_label = grad_w[layer * 3 + 0], w_bbox = grad_w[layer * 3 + 1], w_giou = grad_w[layer * 3 + 2];
  const float2* lg = reinterpret_cast<const float2*>(a.logits) + static_cast<size_t>(layer) * n;
  float2* glg = reinterpret_cast<float2*>(grad_logits) + static_cast<size_t>(layer) * n;
  float4* gbx = reinterpret_cast<float4*>(grad_boxes) + static_cast<size_t>(layer) * n;
  const float ce_scale = w_label / n;
  for (int e = r.q0 + tid; e < r.q1; e += blockDim.x) {
    const float2 l = __ldg(lg + e);
    const int w_ = e - r.q0;
    const bool fg = (bitmap[w_ >> 5] >> (w_ & 31)) & 1u;
    const float m = fmaxf(l.x, l.y);
    const float e0 = expf(l.x - m), e1 = expf(l.y - m);
    const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
    const float w = (fg ? 1.0f : a.eos_coef) * ce_scale;
    glg[e] = make_float2(w * (p0 - (fg ? 1.f : 0.f)), w * (p1 - (fg ? 0.f : 1.f)));
    if (!fg) gbx[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t* pi = a.pred_idx + static_cast<size_t>(layer) * d.pitch;
  const int64_t* ti = a.tgt_idx + static_cast<size_t>(layer) * d.pitch;
  const float4* bx = reinterpret_cast<const float4*>(a.boxes) + static_cast<size_t>(layer) * n;
  const float s_l1 = w_bbox / (4.0f * d.K), s_g = -w_giou / d.K;       // loss_giou = mean(1 - giou)
  for (int k = r.k0 + tid; k < r.k1; k += blockDim.x) {
    const int b = chunked ? r.b0 : pair_video(a, k);
    const int e = pred_entry(a, b, pi[k]);
    const float4 s = __ldg(bx + e);
    const float4 t = __ldg(reinterpret_cast<const float4*>(a.tgt_boxes) + tgt_entry(a, d, b, ti[k]));
    const PairGeom g = pair_geom(s, t);
    // d inter, d hull, d area w.r.t. the four corners of the prediction
    const float di_x1 = (g.iw > 0.f && g.x1 < g.tx1) ? g.ih : 0.f, di_x0 = (g.iw > 0.f && g.x0 > g.tx0) ? -g.ih : 0.f;
    const float di_y1 = (g.ih > 0.f && g.y1 < g.ty1) ? g.iw : 0.f, di_y0 = (g.ih > 0.f && g.y0 > g.ty0) ? -g.iw : 0.f;
    const float dh_x1 = (g.cw > 0.f && g.x1 > g.tx1) ? g.ch : 0.f, dh_x0 = (g.cw > 0.f && g.x0 < g.tx0) ? -g.ch : 0.f;
    const float dh_y1 = (g.ch > 0.f && g.y1 > g.ty1) ? g.cw : 0.f, dh_y0 = (g.ch > 0.f && g.y0 < g.ty0) ? -g.cw : 0.f;
    const float da_x1 = g.y1 - g.y0, da_x0 = -(g.y1 - g.y0), da_y1 = g.x1 - g.x0, da_y0 = -(g.x1 - g.x0);
    const float iu2 = 1.f / (g.uni * g.uni), ih2 = 1.f / (g.hull * g.hull);
    auto dgiou = [&](float di, float da, float dh) {
      const float du = da - di;
      return (di * g.uni - g.inter * du) * iu2 + (du * g.hull - g.uni * dh) * ih2;
    };
    const float gx0 = dgiou(di_x0, da_x0, dh_x0), gx1 = dgiou(di_x1, da_x1, dh_x1);
    const float gy0 = dgiou(di_y0, da_y0, dh_y0), gy1 = dgiou(di_y1, da_y1, dh_y1);
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    float4 o;
    o.x = s_l1 * sgn(s.x - t.x) + s_g * (gx0 + gx1);
    o.y = s_l1 * sgn(s.y - t.y) + s_g * (gy0 + gy1);
    o.z = s_l1 * sgn(s.z - t.z) + s_g * 0.5f * (gx1 - gx0);
    o.w = s_l1 * sgn(s.w - t.w) + s_g * 0.5f * (gy1 - gy0);
    gbx[e] = o;
  }
}

static int check_criterion(const CriterionArgs& a, size_t* smem, dim3* grid, int* threads) {
  if (a.NL <= 0 || a.B <= 0 || a.Q <= 0 || (a.meta == nullptr && a.K <= 0))
    return svol_fail(SVOL_ERR_SHAPE, "criterion: bad sizes (K must be > 0)");
  if (a.match_video == nullptr && a.video_match_off == nullptr)
    return svol_fail(SVOL_ERR_NULL, "criterion: match_video or video_match_off is required");
  if (a.meta != nullptr && a.idx_pitch <= 0) return svol_fail(SVOL_ERR_SHAPE, "criterion: a device-side K needs idx_pitch");
  const bool chunked = a.scratch != nullptr && a.video_match_off != nullptr && a.B > 1;
  *grid = dim3(chunked ? a.B : 1, a.NL);
  *threads = chunked ? 256 : CRIT_THREADS;
  *smem = static_cast<size_t>(((chunked ? 1 : a.B) * a.Q + 31) / 32) * 4;
  if (*smem > 160 * 1024) return svol_fail(SVOL_ERR_SHAPE, "criterion: B*Q too large for the shared-memory bitmap");
  return SVOL_OK;
}

int launch_criterion(const CriterionArgs& a, cudaStream_t stream) {
  size_t smem; dim3 grid; int threads;
  if (int rc = check_criterion(a, &smem, &grid, &threads)) return rc;
  static size_t configured = 0;
  if (smem > 40 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(criterion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "criterion: cudaFuncSetAttribute");
    configured = smem;
  }
  criterion_kernel<<<grid, threads, smem, stream>>>(a);
  return svol_check_launch("criterion");
}

int launch_criterion_backward(const CriterionArgs& a, const float* grad_w, float* grad_logits, float* grad_boxes,
                              cudaStream_t stream) {
  size_t smem; dim3 grid; int threads;
  if (int rc = check_criterion(a, &smem, &grid, &threads)) return rc;
  static size_t configured = 0;
  if (smem > 40 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(criterion_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "criterion_backward: cudaFuncSetAttribute");
    configured = smem;
  }
  criterion_backward_kernel<<<grid, threads, smem, stream>>>(a, grad_w, grad_logits, grad_boxes);
  return svol_check_launch("criterion_backward");
}

}  // namespace svol
