// Plain SIMT versions of the two tensor-core kernels, same C-ABI contracts (sm_100a build, but no
// tcgen05 / TMA).  They exist so the GPU tests can triangulate: tcgen05 kernel vs plain kernel vs CPU
// oracle on the same device buffers, also at sizes where the numpy oracle is too slow.  The product
// path (svol_b200/engine.py) never calls them unless SVOL_B200_PLAIN=1 is set for debugging.
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

// One warp per output row; lane n handles columns n, n+32, ...  Row values are staged in shared
// memory so the LayerNorm epilogue can see the whole row.
__global__ void __launch_bounds__(256) gemm_plain_kernel(const __nv_bfloat16* __restrict__ A,
                                                         const __nv_bfloat16* __restrict__ W, const GemmEpilogue ep,
                                                         int M, int N, int K, int lda, int ldw) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  float* a_row = sm + warp * (K + N);
  float* o_row = a_row + K;
  if (row >= M) return;
  for (int k = lane; k < K; k += 32) a_row[k] = __bfloat162float(A[static_cast<size_t>(row) * lda + k]);
  __syncwarp();
  for (int n = lane; n < N; n += 32) {
    const __nv_bfloat16* w = W + static_cast<size_t>(n) * ldw;
    float acc = 0.f;
    for (int k = 0; k < K; k += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(w + k));
      acc = fmaf(a_row[k + 0], bf16_lo(q.x), acc); acc = fmaf(a_row[k + 1], bf16_hi(q.x), acc);
      acc = fmaf(a_row[k + 2], bf16_lo(q.y), acc); acc = fmaf(a_row[k + 3], bf16_hi(q.y), acc);
      acc = fmaf(a_row[k + 4], bf16_lo(q.z), acc); acc = fmaf(a_row[k + 5], bf16_hi(q.z), acc);
      acc = fmaf(a_row[k + 6], bf16_lo(q.w), acc); acc = fmaf(a_row[k + 7], bf16_hi(q.w), acc);
    }
    if (ep.bias) acc += ep.bias[n];
    if (ep.act == SVOL_ACT_RELU) acc = fmaxf(acc, 0.f);
    else if (ep.act == SVOL_ACT_GELU) acc = gelu_erf(acc);
    if (ep.residual)
      acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.residual)[static_cast<size_t>(row) * ep.ld_res + n]);
    o_row[n] = acc;
  }
  __syncwarp();
  if (ep.ln_weight) {
    float s = 0.f;
    for (int n = lane; n < N; n += 32) s += o_row[n];
    const float mean = warp_sum(s) / N;
    float ss = 0.f;
    for (int n = lane; n < N; n += 32) { const float d = o_row[n] - mean; ss += d * d; }
    const float rstd = rsqrtf(warp_sum(ss) / N + ep.ln_eps);
    for (int n = lane; n < N; n += 32) o_row[n] = (o_row[n] - mean) * rstd * ep.ln_weight[n] + ep.ln_bias[n];
    __syncwarp();
  }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out);
  __nv_bfloat16* out_pos = reinterpret_cast<__nv_bfloat16*>(ep.out_pos);
  __nv_bfloat16* out_vt = reinterpret_cast<__nv_bfloat16*>(ep.out_vt);
  for (int n = lane; n < N; n += 32) {
    const float v = o_row[n];
    if (out) out[static_cast<size_t>(row) * ep.ld_out + n] = __float2bfloat16_rn(v);
    if (out_pos) {
      const int prow = ep.pos_row_mod > 0 ? row % ep.pos_row_mod : row;
      out_pos[static_cast<size_t>(row) * ep.ld_out + n] = __float2bfloat16_rn(v + ep.pos[static_cast<size_t>(prow) * ep.ld_pos + n]);
    }
    if (out_vt) {
      const int b = row / ep.vt_len, l = row - b * ep.vt_len;
      out_vt[(static_cast<size_t>(b) * N + n) * ep.vt_pitch + l] = __float2bfloat16_rn(v);
    }
  }
}

int launch_gemm_bf16_plain(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.K % 8 != 0 || a.ldw % 8 != 0)
    return svol_fail(SVOL_ERR_SHAPE, "gemm_plain: K and ldw must be multiples of 8");
  if (a.out_f32 || a.ep.out_pre || a.ep.dact_src)
    return svol_fail(SVOL_ERR_SHAPE, "gemm_plain: the training-step options (out_f32, out_pre, dact_src) exist on the tcgen05 kernel only");
  const size_t smem = 8 * static_cast<size_t>(a.K + a.N) * sizeof(float);
  if (smem > 200 * 1024) return svol_fail(SVOL_ERR_SHAPE, "gemm_plain: K + N too large");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "gemm_plain: cudaFuncSetAttribute");
    configured = smem;
  }
  gemm_plain_kernel<<<(a.M + 7) / 8, 256, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a.A),
                                                         reinterpret_cast<const __nv_bfloat16*>(a.W), a.ep, a.M, a.N,
                                                         a.K, a.lda, a.ldw);
  return svol_check_launch("gemm_plain");
}

// One warp per (sample, head, query); lanes stride over keys; two passes (max, then exp2 / PV).
__global__ void __launch_bounds__(256) attention_plain_kernel(const AttnArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = static_cast<long long>(blockIdx.x) * 8 + warp;
  const long long total = static_cast<long long>(a.B) * a.H * a.Lq;
  if (item >= total) return;
  const int q = static_cast<int>(item % a.Lq);
  const int h = static_cast<int>((item / a.Lq) % a.H);
  const int b = static_cast<int>(item / (static_cast<long long>(a.Lq) * a.H));
  const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(a.q) + (static_cast<size_t>(b) * a.Lq + q) * a.ldq + h * 32;
  const __nv_bfloat16* Kp = reinterpret_cast<const __nv_bfloat16*>(a.k) + static_cast<size_t>(b) * a.Lk * a.ldk + h * 32;
  const __nv_bfloat16* Vt = reinterpret_cast<const __nv_bfloat16*>(a.vt) + (static_cast<size_t>(b) * a.H + h) * 32 * a.vt_pitch;
  const float* mrow = a.key_mask ? a.key_mask + static_cast<size_t>(b) * a.Lk : nullptr;
  float qv[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) qv[i] = __bfloat162float(Q[i]);
  auto score = [&](int kv) {
    const __nv_bfloat16* kr = Kp + static_cast<size_t>(kv) * a.ldk;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s = fmaf(qv[i], __bfloat162float(kr[i]), s);
    return s;
  };
  float m = -INFINITY;
  for (int kv = lane; kv < a.Lk; kv += 32)
    if (!mrow || mrow[kv] != 0.f) m = fmaxf(m, score(kv));
  m = warp_max(m);
  float l = 0.f, acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int kv = lane; kv < a.Lk; kv += 32) {
    if (mrow && mrow[kv] == 0.f) continue;
    const float p = exp2f(score(kv) - m);
    l += p;
    // the tensor-core kernel rounds P to bf16 before P V; mirror that so both agree closely
    const float pb = __bfloat162float(__float2bfloat16_rn(p));
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(pb, __bfloat162float(Vt[static_cast<size_t>(i) * a.vt_pitch + kv]), acc[i]);
  }
  l = warp_sum(l);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + (static_cast<size_t>(b) * a.Lq + q) * a.ldo + h * 32;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) o[i] = __float2bfloat16_rn(v / l);
  }
}

int launch_attention_plain(const AttnArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.H <= 0 || a.Lq <= 0 || a.Lk <= 0) return svol_fail(SVOL_ERR_SHAPE, "attention_plain: bad sizes");
  const long long total = static_cast<long long>(a.B) * a.H * a.Lq;
  attention_plain_kernel<<<static_cast<unsigned>((total + 7) / 8), 256, 0, stream>>>(a);
  return svol_check_launch("attention_plain");
}

}  // namespace svol
