// Evaluation metrics of the reference on the GPU: best IoU per ground-truth box among a frame's top-1 /
// top-5 predictions (lib/evaluate/eval.py:73-99 -> SVOL-R1 / R5, mIoU) and the VOC-style average precision of every
// (video, sketch) unit at 10 IoU thresholds (eval.py:20-70, utils.py:118-202, :98-115 -> SVOL-mAP).
//
// Everything that decides a comparison is float64 in the reference's operation order (explicit __d*_rn: no FMA
// contraction) on predictions rounded to 4 decimals as test.py:161 writes them, so IoU >= threshold decisions and the
// greedy matching are bit-exact with the numpy path.  The inputs are the device arrays the forward already produced
// (svol_postprocess: per-frame score-sorted rows) plus the flat ground truth; the reference's list of per-frame dicts,
// its JSONL round trip and its per-prediction numpy calls never exist.
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace {
__device__ __forceinline__ double round4(float x) { return __ddiv_rn(rint(__dmul_rn(static_cast<double>(x), 1e4)), 1e4); }

struct Box { double x0, y0, x1, y1; };

__device__ __forceinline__ Box load_pred(const float* p) { return Box{round4(p[0]), round4(p[1]), round4(p[2]), round4(p[3])}; }
__device__ __forceinline__ Box load_gt(const float* g) {
  return Box{static_cast<double>(g[0]), static_cast<double>(g[1]), static_cast<double>(g[2]), static_cast<double>(g[3])};
}
// compute_iou_batch_paired (utils.py:36-73), float64, same operation order
__device__ __forceinline__ double iou_pair(const Box& a, const Box& b) {
  const double xmin = fmax(a.x0, b.x0), ymin = fmax(a.y0, b.y0), xmax = fmin(a.x1, b.x1), ymax = fmin(a.y1, b.y1);
  const double inter = __dmul_rn(__dsub_rn(xmax, xmin), __dsub_rn(ymax, ymin));
  const double a1 = __dmul_rn(__dsub_rn(a.x1, a.x0), __dsub_rn(a.y1, a.y0));
  const double a2 = __dmul_rn(__dsub_rn(b.x1, b.x0), __dsub_rn(b.y1, b.y0));
  const double uni = __dsub_rn(__dadd_rn(a1, a2), inter);
  return (xmin <= xmax && ymin <= ymax) ? __ddiv_rn(inter, uni) : 0.0;
}
}  // namespace

// One thread per ground-truth column.  compute_iou_batch_cross (utils.py:76-96) pairs tile(box1) with repeat(box2) and
// reshapes the flat result to (N, M): entry [n, m] is the IoU of (pred[(n*M+m) % N], gt[(n*M+m) / N]) -- for N > 1 the
// columns mix ground-truth boxes.  eval.py:88 takes the column maximum of exactly that array; reproduced literally.
__global__ void __launch_bounds__(128) eval_max_iou_kernel(const float* __restrict__ pred, const int* __restrict__ frame_index,
                                                           const float* __restrict__ gt, const int* __restrict__ gt_off, int F, int S,
                                                           int qf, double* __restrict__ max1, double* __restrict__ max5) {
  const int s = blockIdx.x * 128 + threadIdx.x;
  if (s >= S) return;
  int lo = 0, hi = F;                                   // frame f with gt_off[f] <= s < gt_off[f+1]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (gt_off[mid] <= s) lo = mid; else hi = mid; }
  const int f = lo, g0 = gt_off[f], M = gt_off[f + 1] - g0, m = s - g0;
  const float* pf = pred + static_cast<size_t>(frame_index[f]) * qf * 5;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int N = min(pass == 0 ? 1 : 5, qf);
    double best = 0.0;
    for (int n = 0; n < N; ++n) {
      const int i = n * M + m;
      const double v = iou_pair(load_pred(pf + (i % N) * 5), load_gt(gt + static_cast<size_t>(g0 + i / N) * 4));
      best = n == 0 ? v : fmax(best, v);
    }
    (pass == 0 ? max1 : max5)[s] = best;
  }
}

// One CTA per evaluation unit.  All threads rank the unit's predictions by (rounded) score, descending and stable
// (list.sort(key=-score), utils.py:149); then thread t < 10 runs the greedy matching of IoU threshold t over the sorted
// predictions (utils.py:163-187) and the interpolated AP (utils.py:189-201, :98-115) from the true-positive flags.
__global__ void __launch_bounds__(256) eval_ap_kernel(const float* __restrict__ pred, const int* __restrict__ frame_index,
                                                      const float* __restrict__ gt, const int* __restrict__ gt_off,
                                                      const int* __restrict__ frame_off, int qf, int max_gt, double* __restrict__ ap) {
  extern __shared__ unsigned char sm_raw[];
  const int v = blockIdx.x, f0 = frame_off[v], f1 = frame_off[v + 1];
  const int M = (f1 - f0) * qf, n_gt = gt_off[f1] - gt_off[f0];
  double* score = reinterpret_cast<double*>(sm_raw);                                // [M]
  int* order = reinterpret_cast<int*>(score + M);                                   // [M] sorted position -> prediction
  int* lock = order + M;                                                            // [10][max_gt]
  unsigned char* tp = reinterpret_cast<unsigned char*>(lock + 10 * max_gt);         // [10][M]
  for (int i = threadIdx.x; i < M; i += 256)
    score[i] = round4(pred[(static_cast<size_t>(frame_index[f0 + i / qf]) * qf + i % qf) * 5 + 4]);
  for (int i = threadIdx.x; i < 10 * max_gt; i += 256) lock[i] = -1;
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += 256) {
    const double si = score[i];
    int rank = 0;
    for (int j = 0; j < M; ++j) rank += (score[j] > si) || (score[j] == si && j < i);
    order[rank] = i;
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= 10) return;
  // np.linspace(0.5, 0.95, 10) formatted to 2 decimals (eval.py:22)
  const double thds[10] = {0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95};
  const double thd = thds[t];
  int* lk = lock + t * max_gt;
  unsigned char* tpt = tp + static_cast<size_t>(t) * M;
  int total_tp = 0;
  for (int i = 0; i < M; ++i) {
    const int p = order[i], f = f0 + p / qf;
    const int g0 = gt_off[f], g1 = gt_off[f + 1];
    unsigned char hit = 0;
    if (g1 > g0) {
      const Box pb = load_pred(pred + (static_cast<size_t>(frame_index[f]) * qf + p % qf) * 5);
      // candidates in descending IoU (argsort()[::-1]: among equal IoUs the larger index first); the first one below
      // the threshold ends the search, locked ones are skipped: = the best unlocked ground truth with IoU >= thd
      int best = -1;
      double best_iou = -1.0;
      for (int g = g0; g < g1; ++g) {
        const double u = iou_pair(pb, load_gt(gt + static_cast<size_t>(g) * 4));
        if (u >= thd && lk[g - gt_off[f0]] < 0 && u >= best_iou) { best = g; best_iou = u; }
      }
      if (best >= 0) { hit = 1; lk[best - gt_off[f0]] = i; }
    }
    tpt[i] = hit;
    total_tp += hit;
  }
  // AP: recall changes exactly at the true positives; precision_i = tp_cumsum_i / (i + 1); interpolation = suffix maximum
  const double npos = static_cast<double>(n_gt);
  double acc = 0.0, sufmax = 0.0;
  int tpc = total_tp;
  for (int i = M - 1; i >= 0; --i) {
    const double prec = __ddiv_rn(static_cast<double>(tpc), static_cast<double>(i + 1));
    sufmax = fmax(sufmax, prec);
    if (tpt[i]) {
      const double dr = __dsub_rn(__ddiv_rn(static_cast<double>(tpc), npos), __ddiv_rn(static_cast<double>(tpc - 1), npos));
      acc = __dadd_rn(acc, __dmul_rn(dr, sufmax));
      --tpc;
    }
  }
  ap[static_cast<size_t>(v) * 10 + t] = M > 0 ? acc : 0.0;
}

int launch_eval_max_iou(const float* pred, const int* frame_index, const float* gt, const int* gt_off, int F, int S, int qf,
                        double* max1, double* max5, cudaStream_t stream) {
  if (F <= 0 || S <= 0 || qf <= 0) return svol_fail(SVOL_ERR_SHAPE, "eval_max_iou: bad sizes");
  eval_max_iou_kernel<<<(S + 127) / 128, 128, 0, stream>>>(pred, frame_index, gt, gt_off, F, S, qf, max1, max5);
  return svol_check_launch("eval_max_iou");
}

int launch_eval_average_precision(const float* pred, const int* frame_index, const float* gt, const int* gt_off, const int* frame_off,
                                  int V, int qf, int max_frames, int max_gt, double* ap, cudaStream_t stream) {
  if (V <= 0 || qf <= 0 || max_frames <= 0 || max_gt <= 0) return svol_fail(SVOL_ERR_SHAPE, "eval_average_precision: bad sizes");
  const size_t M = static_cast<size_t>(max_frames) * qf;
  const size_t smem = M * 8 + M * 4 + 10 * static_cast<size_t>(max_gt) * 4 + 10 * M + 16;
  if (smem > 200 * 1024) return svol_fail(SVOL_ERR_SHAPE, "eval_average_precision: unit too large for shared memory");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(eval_ap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "eval_average_precision: cudaFuncSetAttribute");
    configured = smem;
  }
  eval_ap_kernel<<<V, 256, smem, stream>>>(pred, frame_index, gt, gt_off, frame_off, qf, max_gt, ap);
  return svol_check_launch("eval_average_precision");
}

}  // namespace svol
