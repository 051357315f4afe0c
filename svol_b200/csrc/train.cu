// Memory-bound kernels of the TRAINING step of the SVOL head (sm_100a): the pieces of the backward pass that are
// not dense contractions.  The reference gets all of this from torch.autograd over lib/modeling/svanet.py and
// lib/modeling/cross_modal_transformer.py (train.py:222-232: forward, criterion, loss.backward(), optimizer.step());
// here every backward op is an explicit kernel.  Dense gradients (dgrad / wgrad of every nn.Linear) run on the
// tcgen05 GEMM of gemm_tc.cu with transposed operands produced by transpose_bf16 below; the attention backward is
// attn_bwd_tc.cu.
//
// All kernels are warp-per-row / coalesced, fp32 arithmetic, bf16 activations and activation gradients, fp32
// parameter gradients accumulated with atomics (sums over ~10^4..10^5 rows are first reduced per warp and per CTA).
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace {
constexpr int TD = 256;            // hidden_dim
constexpr int TH = 8;              // heads

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
  v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), c = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// LayerNorm forward, bf16 -> bf16, 256 columns, one warp per row; optional second output y + pos with the same
// three position sources as the GEMM epilogue (fp32 table, table repeating every `mod` rows, or sine angles).
// Training forward of norm1..norm6 and of the second input-projection LayerNorm: the GEMM in front of it stores
// the pre-normalisation sum z, which the backward needs (cross_modal_transformer.py:127,141,143,149,156,158).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ w,
                                                             const float* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                             __nv_bfloat16* __restrict__ y_pos, const float* __restrict__ pos,
                                                             int pos_mod, const float* __restrict__ theta, int rows, float eps,
                                                             DropoutCfg drop) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[8], g[8], o[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(z + static_cast<size_t>(row) * TD) + lane), v);
  load8_f32(w + lane * 8, g);
  load8_f32(b + lane * 8, o);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / TD);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; ss += d * d; }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / TD) + eps);
  float yv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) yv[i] = (v[i] - mean) * rstd * g[i] + o[i];
  if (drop.p > 0.f) {     // train-mode Dropout after the LayerNorm (svanet.py:168-170); y_pos is not combined with it
    const unsigned long long key = dropout_key(drop), e0 = static_cast<unsigned long long>(row) * TD + lane * 8;
    const uint32_t thr = dropout_threshold(drop.p);
    const float sc = 1.0f / (1.0f - drop.p);
#pragma unroll
    for (int i = 0; i < 8; ++i) yv[i] = dropout_keep(e0 + i, key, thr) ? yv[i] * sc : 0.f;
  }
  reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * TD)[lane] = pack8(yv);
  if (y_pos) {
    float pp[8];
    if (theta) {
      const float th = __ldg(theta + row);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = th * (1.0f / powf(10000.f, __fdiv_rn(__fmul_rn(2.f, static_cast<float>(lane * 4 + i)), static_cast<float>(TD))));
        a = a > 3.14159265358979f ? a - 6.28318530717959f : a;
        pp[2 * i] = __sinf(a);
        pp[2 * i + 1] = __cosf(a);
      }
    } else {
      const int prow = pos_mod > 0 ? row % pos_mod : row;
      load8_f32(pos + static_cast<size_t>(prow) * TD + lane * 8, pp);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) yv[i] += pp[i];
    reinterpret_cast<uint4*>(y_pos + static_cast<size_t>(row) * TD)[lane] = pack8(yv);
  }
}

int launch_layernorm_bf16(const svol_bf16* z, const float* w, const float* b, svol_bf16* y, svol_bf16* y_pos,
                          const float* pos, int pos_mod, const float* theta, int rows, int cols, float eps, float drop_p,
                          const long long* seed, int site, cudaStream_t stream) {
  if (cols != TD || rows <= 0) return svol_fail(SVOL_ERR_SHAPE, "layernorm_bf16: 256 columns only");
  if (y_pos && !pos && !theta) return svol_fail(SVOL_ERR_NULL, "layernorm_bf16: y_pos needs pos or theta");
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && (!seed || y_pos)))
    return svol_fail(SVOL_ERR_SHAPE, "layernorm_bf16: 0 <= drop_p < 1; dropout needs seed and excludes y_pos");
  layernorm_bf16_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(z), w, b, reinterpret_cast<__nv_bfloat16*>(y),
      reinterpret_cast<__nv_bfloat16*>(y_pos), pos, pos_mod, theta, rows, eps, DropoutCfg{drop_p, seed, site});
  return svol_check_launch("layernorm_bf16");
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  z is the forward input (bf16, or fp32 for the first LayerNorm over the caller's
// features), optionally multiplied by (1 + att[row]) -- the sketch gate of cross_modal_transformer.py:124-127,
// whose product is never stored.  dy = dy1 (+ dy2 + dy3): gradients arriving from up to three consumers.
//   xhat = (z - mean) * rstd;  dgamma += dy * xhat;  dbeta += dy
//   dz   = rstd * (gamma*dy - mean(gamma*dy) - xhat * mean(gamma*dy*xhat))
// without att:  dx = dz.    with att:  dx = dz * (1 + att[row]),  datt[row] = sum_c dz_c * x_c.
// kV = float4/uint4 groups per lane: cols = 256 * kV.
// ---------------------------------------------------------------------------------------------
template <int kV, bool kF32>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ z_, const float* __restrict__ att,
                                                            const __nv_bfloat16* __restrict__ dy1,
                                                            const __nv_bfloat16* __restrict__ dy2,
                                                            const __nv_bfloat16* __restrict__ dy3,
                                                            const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                                                            float* __restrict__ datt, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int rows, float eps, DropoutCfg drop) {
  constexpr int COLS = 256 * kV;
  __shared__ float red[8][32 * 8 * kV + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[kV][8], ag[kV][8], ab[kV][8];
#pragma unroll
  for (int k = 0; k < kV; ++k) {
    load8_f32(gamma + (k * 32 + lane) * 8, g[k]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[k][i] = 0.f; ab[k][i] = 0.f; }
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    float x[kV][8], d[kV][8];
    const float sc = att ? 1.0f + __ldg(att + row) : 1.0f;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kV; ++k) {
      if (kF32) load8_f32(reinterpret_cast<const float*>(z_) + static_cast<size_t>(row) * COLS + (k * 32 + lane) * 8, x[k]);
      else unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(z_) + static_cast<size_t>(row) * COLS) + k * 32 + lane), x[k]);
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy1 + static_cast<size_t>(row) * COLS) + k * 32 + lane), d[k]);
      if (dy2) {
        float t[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy2 + static_cast<size_t>(row) * COLS) + k * 32 + lane), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[k][i] += t[i];
      }
      if (dy3) {
        float t[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy3 + static_cast<size_t>(row) * COLS) + k * 32 + lane), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[k][i] += t[i];
      }
      if (drop.p > 0.f) {      // the forward dropped this LayerNorm's output: dy <- mask * dy / (1 - p), mask recomputed
        const unsigned long long key = dropout_key(drop), e0 = static_cast<unsigned long long>(row) * COLS + (k * 32 + lane) * 8;
        const uint32_t thr = dropout_threshold(drop.p);
        const float dsc = 1.0f / (1.0f - drop.p);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[k][i] = dropout_keep(e0 + i, key, thr) ? d[k][i] * dsc : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += x[k][i];
    }
    const float mean = warp_sum(s) * sc * (1.0f / COLS);          // mean of z = sc * x
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kV; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float c = x[k][i] * sc - mean; ss += c * c; }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / COLS) + eps);
    float s1 = 0.f, s2 = 0.f;
    float xh[kV][8];
#pragma unroll
    for (int k = 0; k < kV; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[k][i] = (x[k][i] * sc - mean) * rstd;
        ag[k][i] += d[k][i] * xh[k][i];
        ab[k][i] += d[k][i];
        const float gd = g[k][i] * d[k][i];
        s1 += gd;
        s2 += gd * xh[k][i];
      }
    s1 = warp_sum(s1) * (1.0f / COLS);
    s2 = warp_sum(s2) * (1.0f / COLS);
    float da = 0.f;
#pragma unroll
    for (int k = 0; k < kV; ++k) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dz = rstd * (g[k][i] * d[k][i] - s1 - xh[k][i] * s2);
        da += dz * x[k][i];
        o[i] = dz * sc;
      }
      if (dx) reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * COLS)[k * 32 + lane] = pack8(o);
    }
    if (datt) {
      da = warp_sum(da);
      if (lane == 0) datt[row] = da;
    }
  }
  // per-CTA reduction of the parameter gradients, then one atomic per column
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kV; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][(k * 32 + lane) * 8 + i] = pass == 0 ? ag[k][i] : ab[k][i];
    __syncthreads();
    for (int c = threadIdx.x; c < COLS; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][c];
      atomicAdd((pass == 0 ? dgamma : dbeta) + c, t);
    }
  }
}

int launch_layernorm_backward(const void* z, int z_is_f32, const float* att, const svol_bf16* dy1, const svol_bf16* dy2,
                              const svol_bf16* dy3, const float* gamma, svol_bf16* dx, float* datt, float* dgamma,
                              float* dbeta, int rows, int cols, float eps, float drop_p, const long long* seed, int site,
                              cudaStream_t stream) {
  if (rows <= 0 || cols % 256 != 0 || cols < 256 || cols > 1024)
    return svol_fail(SVOL_ERR_SHAPE, "layernorm_backward: cols must be 256, 512, 768 or 1024");
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !seed)) return svol_fail(SVOL_ERR_SHAPE, "layernorm_backward: 0 <= drop_p < 1, seed required");
  const DropoutCfg drop{drop_p, seed, site};
  const int grid = min((rows + 7) / 8, sm_count() * 4);
  auto a1 = reinterpret_cast<const __nv_bfloat16*>(dy1);
  auto a2 = reinterpret_cast<const __nv_bfloat16*>(dy2);
  auto a3 = reinterpret_cast<const __nv_bfloat16*>(dy3);
  auto o = reinterpret_cast<__nv_bfloat16*>(dx);
#define SVOL_LNB(KV, F32) layernorm_bwd_kernel<KV, F32><<<grid, 256, 0, stream>>>(z, att, a1, a2, a3, gamma, o, datt, dgamma, dbeta, rows, eps, drop)
  switch (cols / 256 * 2 + (z_is_f32 ? 1 : 0)) {
    case 2: SVOL_LNB(1, false); break;
    case 3: SVOL_LNB(1, true); break;
    case 4: SVOL_LNB(2, false); break;
    case 5: SVOL_LNB(2, true); break;
    case 6: SVOL_LNB(3, false); break;
    case 7: SVOL_LNB(3, true); break;
    case 8: SVOL_LNB(4, false); break;
    default: SVOL_LNB(4, true); break;
  }
#undef SVOL_LNB
  return svol_check_launch("layernorm_backward");
}

// ---------------------------------------------------------------------------------------------
// Elementwise: GELU forward (bf16 -> bf16) and activation backward out = dy * f'(saved).
//   mode RELU: saved = the activation OUTPUT (mask = saved > 0);  mode GELU: saved = the pre-activation,
//   gelu'(x) = Phi(x) + x phi(x)  (F.gelu, erf form; cross_modal_transformer.py:163-179).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gelu_bf16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long n8) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  unpack8(__ldg(x + i), v);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = gelu_erf_fast(v[k]);
  y[i] = pack8(v);
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ saved,
                                                      uint4* __restrict__ out, long long n8, int mode) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= n8) return;
  float d[8], s[8];
  unpack8(__ldg(dy + i), d);
  unpack8(__ldg(saved + i), s);
  if (mode == SVOL_ACT_RELU) {
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = s[k] > 0.f ? d[k] : 0.f;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float x = s[k];
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
      d[k] *= cdf + x * pdf;
    }
  }
  out[i] = pack8(d);
}
int launch_gelu_bf16(const svol_bf16* x, svol_bf16* y, long long n, cudaStream_t stream) {
  if (n <= 0 || n % 8) return svol_fail(SVOL_ERR_SHAPE, "gelu_bf16: n % 8 == 0");
  gelu_bf16_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), n / 8);
  return svol_check_launch("gelu_bf16");
}
int launch_act_backward(const svol_bf16* dy, const svol_bf16* saved, svol_bf16* out, long long n, int mode, cudaStream_t stream) {
  if (n <= 0 || n % 8 || (mode != SVOL_ACT_RELU && mode != SVOL_ACT_GELU)) return svol_fail(SVOL_ERR_SHAPE, "act_backward: n % 8 == 0, mode RELU | GELU");
  act_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(saved), reinterpret_cast<uint4*>(out), n / 8, mode);
  return svol_check_launch("act_backward");
}

// ---------------------------------------------------------------------------------------------
// out[c, r] = in[r, c]  (bf16, 64 x 64 tiles through shared memory) and, optionally, colsum[c] += sum_r in[r, c]
// (the bias gradient of the nn.Linear whose output gradient `in` is).  The transposed copies are the K-major
// operands of the weight-gradient GEMMs: dW[N,K] = dY^T[N, rows] x X^T[K, rows]^T contracts over the token rows.
// Columns [rows, ld_out) of `out` are left untouched (the buffers are zero-initialised once).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, int rows, int cols,
                                                             __nv_bfloat16* __restrict__ out, int ld_out, float* __restrict__ colsum) {
  __shared__ __nv_bfloat16 tile[64][66];
  __shared__ float csum[4][64];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;          // 64 x 4
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r0 + ty + i * 4, c = c0 + tx;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (r < rows && c < cols) v = in[static_cast<size_t>(r) * ld_in + c];
    tile[ty + i * 4][tx] = v;
    acc += __bfloat162float(v);
  }
  if (colsum) csum[ty][tx] = acc;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = c0 + ty + i * 4, r = r0 + tx;
    if (r < rows && c < cols) out[static_cast<size_t>(c) * ld_out + r] = tile[tx][ty + i * 4];
  }
  if (colsum && threadIdx.x < 64 && c0 + tx < cols)
    atomicAdd(colsum + c0 + tx, (csum[0][tx] + csum[1][tx]) + (csum[2][tx] + csum[3][tx]));
}
int launch_transpose_bf16(const svol_bf16* in, int ld_in, int rows, int cols, svol_bf16* out, int ld_out, float* colsum,
                          cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || ld_out < rows || ld_in < cols) return svol_fail(SVOL_ERR_SHAPE, "transpose_bf16: bad sizes");
  transpose_bf16_kernel<<<dim3((rows + 63) / 64, (cols + 63) / 64), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), ld_in, rows, cols, reinterpret_cast<__nv_bfloat16*>(out), ld_out, colsum);
  return svol_check_launch("transpose_bf16");
}

// colsum[c] += sum_r in[r, c] (bias gradient of an nn.Linear from its output gradient).  A CTA covers 256 columns:
// warp w reads rows w, w + 8, ... 16 bytes per lane (512 contiguous bytes per warp and row), partial sums are combined
// through shared memory and added with one atomic per column.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, int rows, int cols,
                                                          float* __restrict__ colsum) {
  __shared__ float red[8][256 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < cols) {
    uint4 q[4];
    int r = blockIdx.x * 8 + warp;
    const int step = gridDim.x * 8;
    for (; r + 3 * step < rows; r += 4 * step) {           // four independent 16-byte loads in flight per lane
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(in + static_cast<size_t>(r + u * step) * ld_in + c0));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
        unpack8(q[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
    }
    for (; r < rows; r += step) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(in + static_cast<size_t>(r) * ld_in + c0)), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(colsum + c, t);
  }
}
int launch_colsum_bf16(const svol_bf16* in, int ld_in, int rows, int cols, float* colsum, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || cols % 8 != 0 || ld_in % 8 != 0 || (reinterpret_cast<uintptr_t>(in) & 15))
    return svol_fail(SVOL_ERR_SHAPE, "colsum_bf16: cols and ld_in must be multiples of 8, in 16-byte aligned");
  const int cblocks = (cols + 255) / 256;
  int gx = (rows + 31) / 32;
  const int cap = max(1, 4 * sm_count() / cblocks);
  if (gx > cap) gx = cap;
  colsum_bf16_kernel<<<dim3(gx, cblocks), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld_in, rows, cols, colsum);
  return svol_check_launch("colsum_bf16");
}

// ---------------------------------------------------------------------------------------------
// delta[b,h,q] = sum_d dO[b,q,h*32+d] * O[b,q,h*32+d]: the softmax-backward row term of flash attention
// (dS = P * (dP - delta)).  Layout [B, H, pitch], pitch >= Lq.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                                         float* __restrict__ delta, int B, int Lq, int pitch) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * Lq) return;
  float a[8], c[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(dO + row * TD) + lane), a);
  unpack8(__ldg(reinterpret_cast<const uint4*>(O + row * TD) + lane), c);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(a[i], c[i], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if ((lane & 3) == 0) {
    const int b = static_cast<int>(row / Lq), q = static_cast<int>(row - static_cast<long long>(b) * Lq);
    delta[(static_cast<size_t>(b) * TH + (lane >> 2)) * pitch + q] = s;
  }
}
int launch_attn_delta(const svol_bf16* dO, const svol_bf16* O, float* delta, int B, int H, int Lq, int pitch, cudaStream_t stream) {
  if (H != TH || B <= 0 || Lq <= 0 || pitch < Lq) return svol_fail(SVOL_ERR_SHAPE, "attn_delta: 8 heads of 32");
  const long long rows = static_cast<long long>(B) * Lq;
  attn_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dO), reinterpret_cast<const __nv_bfloat16*>(O), delta, B, Lq, pitch);
  return svol_check_launch("attn_delta");
}

// ---------------------------------------------------------------------------------------------
// Backward of svol_heads (svanet.py:125-127): logits = Wc hs + bc, boxes = sigmoid(Wb h2 + bb).
//   dhs_cls[row,:] = sum_j dlogits[row,j] Wc[j,:]
//   dh2[row,:]     = relu'(h2) * sum_j (dboxes * boxes * (1 - boxes))[row,j] Wb[j,:]      (h2 = ReLU output)
//   dWc, dbc, dWb, dbb accumulated in fp32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) heads_bwd_kernel(const __nv_bfloat16* __restrict__ hs, const __nv_bfloat16* __restrict__ h2,
                                                        const float* __restrict__ wc, const float* __restrict__ wb,
                                                        const float* __restrict__ boxes, const float* __restrict__ dlogits,
                                                        const float* __restrict__ dboxes, __nv_bfloat16* __restrict__ dhs,
                                                        __nv_bfloat16* __restrict__ dh2, float* __restrict__ dwc, float* __restrict__ dbc,
                                                        float* __restrict__ dwb, float* __restrict__ dbb, int rows) {
  __shared__ float red[8][3 * TD + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w[6][8], aw[6][8], abias[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    load8_f32((j < 2 ? wc + j * TD : wb + (j - 2) * TD) + lane * 8, w[j]);
    abias[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[j][i] = 0.f;
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    float a[8], c[8], g[6];
    unpack8(__ldg(reinterpret_cast<const uint4*>(hs + static_cast<size_t>(row) * TD) + lane), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(h2 + static_cast<size_t>(row) * TD) + lane), c);
    g[0] = __ldg(dlogits + static_cast<size_t>(row) * 2);
    g[1] = __ldg(dlogits + static_cast<size_t>(row) * 2 + 1);
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + row), db = __ldg(reinterpret_cast<const float4*>(dboxes) + row);
    g[2] = db.x * bx.x * (1.f - bx.x); g[3] = db.y * bx.y * (1.f - bx.y);
    g[4] = db.z * bx.z * (1.f - bx.z); g[5] = db.w * bx.w * (1.f - bx.w);
    float o1[8], o2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o1[i] = g[0] * w[0][i] + g[1] * w[1][i];
      const float t = (g[2] * w[2][i] + g[3] * w[3][i]) + (g[4] * w[4][i] + g[5] * w[5][i]);
      o2[i] = c[i] > 0.f ? t : 0.f;
    }
    reinterpret_cast<uint4*>(dhs + static_cast<size_t>(row) * TD)[lane] = pack8(o1);
    reinterpret_cast<uint4*>(dh2 + static_cast<size_t>(row) * TD)[lane] = pack8(o2);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      abias[j] += g[j];
#pragma unroll
      for (int i = 0; i < 8; ++i) aw[j][i] += g[j] * (j < 2 ? a[i] : c[i]);
    }
  }
  // two passes of three weight rows each (static shared memory stays below 48 KB)
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][jj * TD + lane * 8 + i] = aw[pass * 3 + jj][i];
      if (lane == 0) red[warp][3 * TD + jj] = abias[pass * 3 + jj];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * TD + 3; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][c];
      const int j = pass * 3 + (c < 3 * TD ? c / TD : c - 3 * TD);      // weight row 0..5 (0,1: class; 2..5: box)
      if (c < 3 * TD) {
        const int col = c % TD;
        if (j < 2) atomicAdd(dwc + j * TD + col, t); else atomicAdd(dwb + (j - 2) * TD + col, t);
      } else {
        if (j < 2) atomicAdd(dbc + j, t); else atomicAdd(dbb + (j - 2), t);
      }
    }
  }
}
int launch_heads_backward(const svol_bf16* hs, const svol_bf16* h2, const float* wc, const float* wb, const float* boxes,
                          const float* dlogits, const float* dboxes, svol_bf16* dhs, svol_bf16* dh2, float* dwc, float* dbc,
                          float* dwb, float* dbb, int rows, int d, cudaStream_t stream) {
  if (d != TD || rows <= 0) return svol_fail(SVOL_ERR_SHAPE, "heads_backward: hidden_dim 256 only");
  const int grid = min((rows + 7) / 8, sm_count() * 2);
  heads_bwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(hs), reinterpret_cast<const __nv_bfloat16*>(h2), wc, wb,
                                             boxes, dlogits, dboxes, reinterpret_cast<__nv_bfloat16*>(dhs),
                                             reinterpret_cast<__nv_bfloat16*>(dh2), dwc, dbc, dwb, dbb, rows);
  return svol_check_launch("heads_backward");
}

// ---------------------------------------------------------------------------------------------
// Backward of the sketch gate (cross_modal_transformer.py:122-125):  att[b,l] = mean_h softmax_l(scores[b,h,:]).
//   gate_softmax_bwd:  dscores[b,h,l] = p_h[l] * (datt[l] - sum_l' p_h[l'] datt[l']) / H        (one CTA per (h, b))
//   gate_scores_bwd:   dx_out[b,l,:]  = dx_in[b,l,:] + sum_h dscores[b,h,l] u[b,h,:]           (scores = (x+pos) . u)
//                      du[b,h,:]     += sum_l dscores[b,h,l] (x+pos)[b,l,:]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_softmax_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ datt,
                                                               float* __restrict__ dscores, int L) {
  __shared__ float red[8];
  __shared__ float bc[3];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* sr = scores + (static_cast<size_t>(b) * TH + h) * L;
  const float* da = datt + static_cast<size_t>(b) * L;
  auto block_reduce = [&](float v, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = red[0];
    for (int i = 1; i < 8; ++i) t = is_max ? fmaxf(t, red[i]) : t + red[i];
    return t;
  };
  float m = -INFINITY;
  for (int l = tid; l < L; l += 256) m = fmaxf(m, sr[l]);
  m = block_reduce(m, true);
  float s = 0.f, t = 0.f;
  for (int l = tid; l < L; l += 256) { const float e = expf(sr[l] - m); s += e; t += e * da[l]; }
  s = block_reduce(s, false);
  t = block_reduce(t, false);
  if (tid == 0) { bc[0] = m; bc[1] = 1.f / s; bc[2] = t / s; }
  __syncthreads();
  const float inv = bc[1], tbar = bc[2];
  float* out = dscores + (static_cast<size_t>(b) * TH + h) * L;
  for (int l = tid; l < L; l += 256) out[l] = expf(sr[l] - m) * inv * (da[l] - tbar) * (1.0f / TH);
}

__global__ void __launch_bounds__(256) gate_scores_bwd_kernel(const __nv_bfloat16* __restrict__ xpos, const float* __restrict__ u,
                                                              const float* __restrict__ dscores, const __nv_bfloat16* __restrict__ dx_in,
                                                              __nv_bfloat16* __restrict__ dx_out, float* __restrict__ du, int L,
                                                              int rows_per_cta) {
  __shared__ float red[8][TH * TD / 8 + 1];   // one head at a time: 8 warps x 256 columns
  const int b = blockIdx.y, l0 = blockIdx.x * rows_per_cta, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float ur[TH][8], acc[TH][8];
#pragma unroll
  for (int h = 0; h < TH; ++h) {
    load8_f32(u + (static_cast<size_t>(b) * TH + h) * TD + lane * 8, ur[h]);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[h][i] = 0.f;
  }
  const int l_end = min(L, l0 + rows_per_cta);
  for (int l = l0 + warp; l < l_end; l += 8) {
    const size_t row = static_cast<size_t>(b) * L + l;
    float xp[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(xpos + row * TD) + lane), xp);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dx_in + row * TD) + lane), g);
#pragma unroll
    for (int h = 0; h < TH; ++h) {
      const float ds = __ldg(dscores + (static_cast<size_t>(b) * TH + h) * L + l);
#pragma unroll
      for (int i = 0; i < 8; ++i) { g[i] = fmaf(ds, ur[h][i], g[i]); acc[h][i] = fmaf(ds, xp[i], acc[h][i]); }
    }
    reinterpret_cast<uint4*>(dx_out + row * TD)[lane] = pack8(g);
  }
#pragma unroll
  for (int h = 0; h < TH; ++h) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[h][i];
    __syncthreads();
    {
      const int c = tid;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][c];
      atomicAdd(du + (static_cast<size_t>(b) * TH + h) * TD + c, t);
    }
  }
}
int launch_gate_backward(const svol_bf16* xpos, const float* u, const float* scores, const float* datt, const svol_bf16* dx_in,
                         svol_bf16* dx_out, float* dscores, float* du, int B, int L, int d, int H, cudaStream_t stream) {
  if (d != TD || H != TH || B <= 0 || L <= 0) return svol_fail(SVOL_ERR_SHAPE, "gate_backward: hidden_dim 256 / 8 heads only");
  cudaError_t e = cudaMemsetAsync(du, 0, static_cast<size_t>(B) * TH * TD * sizeof(float), stream);     // du is written, not accumulated
  if (e != cudaSuccess) return svol_fail_cuda(e, "gate_backward: memset");
  gate_softmax_bwd_kernel<<<dim3(TH, B), 256, 0, stream>>>(scores, datt, dscores, L);
  int rc = svol_check_launch("gate_softmax_bwd");
  if (rc) return rc;
  const int rows_per_cta = 128;
  gate_scores_bwd_kernel<<<dim3((L + rows_per_cta - 1) / rows_per_cta, B), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(xpos), u, dscores, reinterpret_cast<const __nv_bfloat16*>(dx_in),
      reinterpret_cast<__nv_bfloat16*>(dx_out), du, L, rows_per_cta);
  return svol_check_launch("gate_scores_bwd");
}

// Backward of svol_gate_vectors:  qs_j = (Wq[hj,:] . s_b + bq[hj]) / sqrt(dh);  u[b,h,c] = sum_j qs_j Wk[hj,c].
// One CTA per (h, b); w / dw are the fp32 in_proj_weight [3d,d] of sketch_video_cross_attn and its gradient.
__global__ void __launch_bounds__(256) gate_vectors_bwd_kernel(const float* __restrict__ sketch, const float* __restrict__ w,
                                                               const float* __restrict__ bias, const float* __restrict__ du,
                                                               float* __restrict__ dw, float* __restrict__ dbias,
                                                               float* __restrict__ dsketch, int d, int H) {
  extern __shared__ float sm[];     // s[d], du[d], qs[dh], dacc[dh]
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dh = d / H;
  float* s = sm;
  float* g = sm + d;
  float* qs = g + d;
  float* dacc = qs + dh;
  const float rs = rsqrtf(static_cast<float>(dh));
  for (int i = tid; i < d; i += blockDim.x) {
    s[i] = sketch[static_cast<size_t>(b) * d + i];
    g[i] = du[(static_cast<size_t>(b) * H + h) * d + i];
  }
  __syncthreads();
  for (int j = warp; j < dh; j += 8) {
    const float* wq = w + static_cast<size_t>(h * dh + j) * d;
    const float* wk = w + static_cast<size_t>(d + h * dh + j) * d;
    float a = 0.f, c = 0.f;
    for (int i = lane; i < d; i += 32) { a = fmaf(s[i], __ldg(wq + i), a); c = fmaf(g[i], __ldg(wk + i), c); }
    a = warp_sum(a); c = warp_sum(c);
    if (lane == 0) { qs[j] = (a + bias[h * dh + j]) * rs; dacc[j] = c * rs; }
  }
  __syncthreads();
  for (int j = 0; j < dh; ++j) {
    const float q = qs[j], da = dacc[j];
    for (int c = tid; c < d; c += blockDim.x) {
      atomicAdd(dw + static_cast<size_t>(d + h * dh + j) * d + c, q * g[c]);        // dWk
      atomicAdd(dw + static_cast<size_t>(h * dh + j) * d + c, da * s[c]);           // dWq
    }
    if (tid == 0) atomicAdd(dbias + h * dh + j, da);
  }
  for (int c = tid; c < d; c += blockDim.x) {
    float t = 0.f;
    for (int j = 0; j < dh; ++j) t = fmaf(dacc[j], __ldg(w + static_cast<size_t>(h * dh + j) * d + c), t);
    atomicAdd(dsketch + static_cast<size_t>(b) * d + c, t);
  }
}
int launch_gate_vectors_backward(const float* sketch, const float* w, const float* bias, const float* du, float* dw, float* dbias,
                                 float* dsketch, int B, int d, int H, cudaStream_t stream) {
  if (B <= 0 || d <= 0 || H <= 0 || d % H != 0) return svol_fail(SVOL_ERR_SHAPE, "gate_vectors_backward: bad sizes");
  gate_vectors_bwd_kernel<<<dim3(H, B), 256, (2 * d + 2 * (d / H)) * sizeof(float), stream>>>(sketch, w, bias, du, dw, dbias, dsketch, d, H);
  return svol_check_launch("gate_vectors_backward");
}

// ---------------------------------------------------------------------------------------------
// Backward of svol_ln_linear_f32 (one LinearLayer of the sketch branch, svanet.py:56-60,159-181), fp32:
//   y = [ReLU](W xn + b), xn = LayerNorm(x) * gamma + beta.
// Two kernels: (rows x out_dim/32) CTAs accumulate dW / db and their share of dxn = dy W into the dx buffer (zeroed
// first), then one CTA per row turns dxn into dx (LayerNorm backward) and dgamma / dbeta.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}
// xh[i] = (x - mean) * rstd of one row (in shared memory); returns rstd
__device__ __forceinline__ float row_normalise(const float* __restrict__ xr, float* xh, float* red, int in_dim, float eps) {
  float s = 0.f;
  for (int i = threadIdx.x; i < in_dim; i += 256) { xh[i] = xr[i]; s += xr[i]; }
  const float mean = block_sum_256(s, red) / in_dim;
  float ss = 0.f;
  for (int i = threadIdx.x; i < in_dim; i += 256) { const float c = xh[i] - mean; ss += c * c; }
  const float rstd = rsqrtf(block_sum_256(ss, red) / in_dim + eps);
  for (int i = threadIdx.x; i < in_dim; i += 256) xh[i] = (xh[i] - mean) * rstd;
  __syncthreads();
  return rstd;
}

__global__ void __launch_bounds__(256) ln_linear_f32_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ lw,
                                                                  const float* __restrict__ lb, const float* __restrict__ w,
                                                                  const float* __restrict__ y, const float* __restrict__ dy, int relu,
                                                                  float* __restrict__ dxn, float* __restrict__ dw, float* __restrict__ db,
                                                                  int in_dim, int out_dim, float eps, DropoutCfg drop) {
  extern __shared__ float sm[];     // xhat[in], dyr[32], red[8]
  float* xh = sm;
  float* dyr = xh + in_dim;
  float* red = dyr + 32;
  const int row = blockIdx.x, o0 = blockIdx.y * 32, tid = threadIdx.x;
  row_normalise(x + static_cast<size_t>(row) * in_dim, xh, red, in_dim, eps);
  if (tid < 32) {
    const int o = o0 + tid;
    float g = 0.f;
    if (o < out_dim) {
      g = dy[static_cast<size_t>(row) * out_dim + o];
      if (relu && !(y[static_cast<size_t>(row) * out_dim + o] > 0.f)) g = 0.f;
      atomicAdd(db + o, g);
    }
    dyr[tid] = g;
  }
  __syncthreads();
  const unsigned long long dkey = drop.p > 0.f ? dropout_key(drop) : 0ull;
  const uint32_t dthr = dropout_threshold(drop.p);
  const float dsc = 1.0f / (1.0f - drop.p);
  for (int i = tid; i < in_dim; i += 256) {
    // the Linear saw dropout(xn): keep ? xn / (1 - p) : 0 (mask recomputed); its input gradient passes the same mask
    const float keep = (drop.p > 0.f && !dropout_keep(static_cast<unsigned long long>(row) * in_dim + i, dkey, dthr)) ? 0.f
                       : (drop.p > 0.f ? dsc : 1.0f);
    const float xn = (xh[i] * lw[i] + lb[i]) * keep;
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const int o = o0 + j;
      if (o >= out_dim) break;
      const float g = dyr[j];
      acc = fmaf(g, __ldg(w + static_cast<size_t>(o) * in_dim + i), acc);
      if (g != 0.f) atomicAdd(dw + static_cast<size_t>(o) * in_dim + i, g * xn);
    }
    atomicAdd(dxn + static_cast<size_t>(row) * in_dim + i, acc * keep);
  }
}

__global__ void __launch_bounds__(256) ln_linear_f32_bwd_x_kernel(const float* __restrict__ x, const float* __restrict__ lw,
                                                                  float* __restrict__ dx /* in: dxn, out: dx */,
                                                                  float* __restrict__ dlw, float* __restrict__ dlb, int in_dim, float eps) {
  extern __shared__ float sm[];     // xhat[in], red[8]
  float* xh = sm;
  float* red = xh + in_dim;
  const int row = blockIdx.x, tid = threadIdx.x;
  const float rstd = row_normalise(x + static_cast<size_t>(row) * in_dim, xh, red, in_dim, eps);
  float* dr = dx + static_cast<size_t>(row) * in_dim;
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid; i < in_dim; i += 256) {
    const float d = dr[i], gd = lw[i] * d;
    s1 += gd; s2 += gd * xh[i];
    atomicAdd(dlw + i, d * xh[i]);
    atomicAdd(dlb + i, d);
  }
  s1 = block_sum_256(s1, red) / in_dim;
  s2 = block_sum_256(s2, red) / in_dim;
  for (int i = tid; i < in_dim; i += 256) dr[i] = rstd * (lw[i] * dr[i] - s1 - xh[i] * s2);
}

int launch_ln_linear_f32_backward(const float* x, const float* lw, const float* lb, const float* w, const float* y, const float* dy,
                                  int relu, float* dx, float* dlw, float* dlb, float* dw, float* db, int rows, int in_dim,
                                  int out_dim, float eps, float drop_p, const long long* seed, int site, cudaStream_t stream) {
  if (rows <= 0 || in_dim <= 0 || out_dim <= 0 || in_dim > 8192) return svol_fail(SVOL_ERR_SHAPE, "ln_linear_backward: bad sizes");
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !seed)) return svol_fail(SVOL_ERR_SHAPE, "ln_linear_backward: 0 <= drop_p < 1, seed required");
  cudaError_t e = cudaMemsetAsync(dx, 0, static_cast<size_t>(rows) * in_dim * sizeof(float), stream);
  if (e != cudaSuccess) return svol_fail_cuda(e, "ln_linear_backward: memset");
  ln_linear_f32_bwd_w_kernel<<<dim3(rows, (out_dim + 31) / 32), 256, (in_dim + 48) * sizeof(float), stream>>>(
      x, lw, lb, w, y, dy, relu, dx, dw, db, in_dim, out_dim, eps, DropoutCfg{drop_p, seed, site});
  int rc = svol_check_launch("ln_linear_f32_backward (weights)");
  if (rc) return rc;
  ln_linear_f32_bwd_x_kernel<<<rows, 256, (in_dim + 16) * sizeof(float), stream>>>(x, lw, dx, dlw, dlb, in_dim, eps);
  return svol_check_launch("ln_linear_f32_backward (input)");
}

// ---------------------------------------------------------------------------------------------
// acc[r % mod, :] += sum over the rows r of g (bf16 -> fp32): gradient of the query embedding, which the
// forward broadcasts over the batch (cross_modal_transformer.py:52-56; rows = B*Q, mod = Q).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) batch_sum_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ acc, int rows, int mod) {
  const int q = blockIdx.x, c = threadIdx.x;       // 256 columns
  float t = 0.f;
  for (int r = q; r < rows; r += mod) t += __bfloat162float(g[static_cast<size_t>(r) * TD + c]);
  acc[static_cast<size_t>(q) * TD + c] += t;
}
int launch_batch_sum(const svol_bf16* g, float* acc, int rows, int cols, int mod, cudaStream_t stream) {
  if (cols != TD || rows <= 0 || mod <= 0 || rows % mod) return svol_fail(SVOL_ERR_SHAPE, "batch_sum: 256 columns, rows % mod == 0");
  batch_sum_kernel<<<mod, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(g), acc, rows, mod);
  return svol_check_launch("batch_sum");
}

// dst[i] (+)= scale * src[i]   (bf16 weight-gradient GEMM output -> fp32 parameter gradient)
__global__ void __launch_bounds__(256) accum_bf16_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n,
                                                         float scale, int accumulate) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= n) return;
  const float v = scale * __bfloat162float(src[i]);
  dst[i] = accumulate ? dst[i] + v : v;
}
int launch_accum_bf16(const svol_bf16* src, float* dst, long long n, float scale, int accumulate, cudaStream_t stream) {
  if (n <= 0) return svol_fail(SVOL_ERR_SHAPE, "accum_bf16: n > 0");
  accum_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n, scale, accumulate);
  return svol_check_launch("accum_bf16");
}

// ---------------------------------------------------------------------------------------------
// Weight packing: every bf16 / fp32 operand copy the launch plans read (projection weights with the query rows
// pre-scaled by log2(e)/sqrt(dh), transposed weights for the dgrad GEMMs, biases) is refreshed from the fp32
// parameters in ONE launch driven by a job table -- after an optimizer step the per-parameter torch casts /
// concatenations this replaces cost more GPU time than the AdamW update itself.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const svol_pack_job* __restrict__ jobs) {
  const svol_pack_job j = jobs[blockIdx.y];
  const long long n = static_cast<long long>(j.rows) * j.cols;
  const bool to_bf16 = j.flags & SVOL_PACK_BF16, transpose = j.flags & SVOL_PACK_TRANSPOSE;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const int r = static_cast<int>(i / j.cols), c = static_cast<int>(i - static_cast<long long>(r) * j.cols);
    float v = j.src[i];
    if (r < j.scaled_rows) v *= j.scale;
    const long long o = transpose ? static_cast<long long>(c) * j.rows + r : i;
    if (to_bf16) reinterpret_cast<__nv_bfloat16*>(j.dst)[o] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(j.dst)[o] = v;
  }
}
int launch_pack_weights(const svol_pack_job* jobs, int n_jobs, cudaStream_t stream) {
  if (n_jobs <= 0) return svol_fail(SVOL_ERR_SHAPE, "pack_weights: n_jobs > 0");
  pack_weights_kernel<<<dim3(64, n_jobs), 256, 0, stream>>>(jobs);
  return svol_check_launch("pack_weights");
}

// ---------------------------------------------------------------------------------------------
// Fused AdamW over one flat fp32 parameter buffer (torch.optim.AdamW semantics, train.py:71-78):
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// grad_scale multiplies g first (1 / world_size after a sum all-reduce).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, float grad_scale) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  float pi = p[i] * (1.0f - lr * wd);
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  p[i] = pi;
}
int launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
                 int step, float grad_scale, cudaStream_t stream) {
  if (n <= 0 || step <= 0) return svol_fail(SVOL_ERR_SHAPE, "adamw: n > 0, step >= 1");
  const float bc1 = 1.0f - powf(b1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(b2, static_cast<float>(step));
  adamw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, sqrtf(bc2), grad_scale);
  return svol_check_launch("adamw");
}

// Same update over a flat buffer that is a concatenation of parameter SEGMENTS (4-element aligned): segment s covers
// [seg_end[s-1], seg_end[s]) and belongs to hyper-parameter group seg_group[s], or is skipped when seg_group[s] < 0 --
// torch.optim.AdamW leaves a parameter whose .grad is None untouched (no weight decay, no moment update).
struct AdamwGroups { svol_adamw_group g[SVOL_ADAMW_MAX_GROUPS]; float bc1[SVOL_ADAMW_MAX_GROUPS]; float bc2_sqrt[SVOL_ADAMW_MAX_GROUPS]; };

__global__ void __launch_bounds__(256) adamw_segments_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                             float* __restrict__ m, float* __restrict__ v, long long n4,
                                                             const long long* __restrict__ seg_end,
                                                             const int* __restrict__ seg_group, int n_seg, AdamwGroups G,
                                                             float grad_scale) {
  const long long i4 = blockIdx.x * 256LL + threadIdx.x;
  if (i4 >= n4) return;
  const long long i = i4 * 4;
  int lo = 0, hi = n_seg - 1;                       // first segment with seg_end > i
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seg_end + mid) > i) hi = mid; else lo = mid + 1;
  }
  const int gi = __ldg(seg_group + lo);
  if (gi < 0) return;
  const svol_adamw_group h = G.g[gi];
  const float bc1 = G.bc1[gi], bc2s = G.bc2_sqrt[gi];
  float4 pv = reinterpret_cast<float4*>(p)[i4], mv = reinterpret_cast<float4*>(m)[i4], vv = reinterpret_cast<float4*>(v)[i4];
  const float4 gv = reinterpret_cast<const float4*>(g)[i4];
  auto upd = [&](float& pi, float& mi, float& vi, float gr) {
    gr *= grad_scale;
    pi *= 1.0f - h.lr * h.weight_decay;
    mi = h.beta1 * mi + (1.0f - h.beta1) * gr;
    vi = h.beta2 * vi + (1.0f - h.beta2) * gr * gr;
    pi -= (h.lr / bc1) * mi / (sqrtf(vi) / bc2s + h.eps);
  };
  upd(pv.x, mv.x, vv.x, gv.x); upd(pv.y, mv.y, vv.y, gv.y); upd(pv.z, mv.z, vv.z, gv.z); upd(pv.w, mv.w, vv.w, gv.w);
  reinterpret_cast<float4*>(p)[i4] = pv; reinterpret_cast<float4*>(m)[i4] = mv; reinterpret_cast<float4*>(v)[i4] = vv;
}

int launch_adamw_segments(float* p, const float* g, float* m, float* v, long long n, const long long* seg_end,
                          const int* seg_group, int n_seg, const svol_adamw_group* groups, int n_groups, float grad_scale,
                          cudaStream_t stream) {
  if (n <= 0 || (n & 3) || n_seg <= 0 || n_groups <= 0 || n_groups > SVOL_ADAMW_MAX_GROUPS)
    return svol_fail(SVOL_ERR_SHAPE, "adamw_segments: n > 0 and a multiple of 4, 1..8 groups, >= 1 segment");
  AdamwGroups G;
  for (int i = 0; i < n_groups; ++i) {
    if (groups[i].step <= 0) return svol_fail(SVOL_ERR_SHAPE, "adamw_segments: group step >= 1");
    G.g[i] = groups[i];
    G.bc1[i] = 1.0f - powf(groups[i].beta1, static_cast<float>(groups[i].step));
    G.bc2_sqrt[i] = sqrtf(1.0f - powf(groups[i].beta2, static_cast<float>(groups[i].step)));
  }
  const long long n4 = n / 4;
  adamw_segments_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(p, g, m, v, n4, seg_end, seg_group, n_seg, G,
                                                                                      grad_scale);
  return svol_check_launch("adamw_segments");
}

}  // namespace svol
