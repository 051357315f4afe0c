// Memory-bound row-wise kernels of the SVOL head (sm_100a): LayerNorm of the input features, the
// fp32 sketch branch, the sine positional table, the sketch-conditioned gate, the output heads and
// the inference post-processing.  All are coalesced, 16-byte vectorised, warp-per-row kernels with
// shuffle reductions; none of them has data reuse that would justify tensor cores.
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

// ---------------------------------------------------------------------------------------------
// LayerNorm fp32 -> bf16, one warp per row (first op of LinearLayer, svanet.py:174-176)
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;   // float4 per lane -> cols <= 1024

// 16-byte load of four consecutive inputs (fp32) / eight consecutive inputs (bf16 feature caches) as floats
__device__ __forceinline__ float4 ln_load4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ln_load4(const __nv_bfloat16* p) {
  const uint2 q = __ldcs(reinterpret_cast<const uint2*>(p));
  return make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
}

// Persistent warps: a warp keeps its 4 * kV columns of gamma / beta in registers and walks rows with a stride of the whole grid,
// TWO rows per iteration (both rows' 16-byte loads are issued before either is reduced).  The one-row-per-warp form re-read
// gamma and beta (fp32, 2 x the row's own bytes) through the LSU for every row and was half instruction-bound
// (ncu: issue slots 56 % busy at 3.8 TB/s); kV is a template parameter so the common widths have no per-element predicates.
template <bool kDrop, typename TIn, int kV>
__global__ void __launch_bounds__(256, kV > 6 ? 1 : 2) layernorm_f32_to_bf16_kernel(const TIn* __restrict__ x,
                                                                        const float* __restrict__ w,
                                                                        const float* __restrict__ b,
                                                                        __nv_bfloat16* __restrict__ y, int rows,
                                                                        int cols, float eps, DropoutCfg drop) {
  const int lane = threadIdx.x & 31;
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  const int nv = cols >> 2;
  const bool exact = nv == kV * 32;
  float4 g[kV], o[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int idx = i * 32 + lane;
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f); o[i] = g[i];
    if (exact || idx < nv) { g[i] = __ldg(reinterpret_cast<const float4*>(w) + idx); o[i] = __ldg(reinterpret_cast<const float4*>(b) + idx); }
  }
  const float inv_cols = 1.0f / cols;
  // The loads of the NEXT pair of rows are issued before the current pair is reduced, normalised and written (software
  // pipelining in registers): with load -> reduce -> write strictly in sequence a warp has nothing in flight for half of
  // every iteration (ncu: 4.4 TB/s = 68 % of the measured HBM peak at 15 resident warps per SM).
  auto load_rows = [&](int row0, float4 (&dst)[2][kV]) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row0 + r * n_warps;
      const TIn* xr = x + static_cast<size_t>(row) * cols;
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        const int idx = i * 32 + lane;
        dst[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows && (exact || idx < nv)) dst[r][i] = ln_load4(xr + idx * 4);       // streamed once
      }
    }
  };
  float4 nxt[2][kV];
  const int row_first = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  load_rows(row_first, nxt);
  for (int row0 = row_first; row0 < rows; row0 += 2 * n_warps) {
    float4 buf[2][kV];
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < kV; ++i) buf[r][i] = nxt[r][i];
    load_rows(row0 + 2 * n_warps, nxt);            // (rows beyond the end: predicated off)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < kV; ++i) s[r] += (buf[r][i].x + buf[r][i].y) + (buf[r][i].z + buf[r][i].w);
    float mean[2], ss[2] = {0.f, 0.f};
#pragma unroll
    for (int o_ = 16; o_ > 0; o_ >>= 1) {
      s[0] += __shfl_xor_sync(0xffffffffu, s[0], o_);
      s[1] += __shfl_xor_sync(0xffffffffu, s[1], o_);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mean[r] = s[r] * inv_cols;
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        if (exact || i * 32 + lane < nv) {
          const float a = buf[r][i].x - mean[r], c = buf[r][i].y - mean[r], d = buf[r][i].z - mean[r], e = buf[r][i].w - mean[r];
          ss[r] += (a * a + c * c) + (d * d + e * e);
        }
      }
    }
#pragma unroll
    for (int o_ = 16; o_ > 0; o_ >>= 1) {
      ss[0] += __shfl_xor_sync(0xffffffffu, ss[0], o_);
      ss[1] += __shfl_xor_sync(0xffffffffu, ss[1], o_);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row0 + r * n_warps;
      if (row >= rows) break;
      const float rstd = rsqrtf(ss[r] * inv_cols + eps);
      uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * cols);
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        const int idx = i * 32 + lane;
        if (exact || idx < nv) {
          float v0 = (buf[r][i].x - mean[r]) * rstd * g[i].x + o[i].x, v1 = (buf[r][i].y - mean[r]) * rstd * g[i].y + o[i].y;
          float v2 = (buf[r][i].z - mean[r]) * rstd * g[i].z + o[i].z, v3 = (buf[r][i].w - mean[r]) * rstd * g[i].w + o[i].w;
          if (kDrop) {       // train-mode Dropout after the LayerNorm (svanet.py:168-170)
            const unsigned long long key = dropout_key(drop), e0 = static_cast<unsigned long long>(row) * cols + idx * 4;
            const uint32_t thr = dropout_threshold(drop.p);
            const float sc = 1.0f / (1.0f - drop.p);
            v0 = dropout_keep(e0, key, thr) ? v0 * sc : 0.f; v1 = dropout_keep(e0 + 1, key, thr) ? v1 * sc : 0.f;
            v2 = dropout_keep(e0 + 2, key, thr) ? v2 * sc : 0.f; v3 = dropout_keep(e0 + 3, key, thr) ? v3 * sc : 0.f;
          }
          uint2 q;
          q.x = pack_bf16x2(v0, v1);
          q.y = pack_bf16x2(v2, v3);
          yr[idx] = q;
        }
      }
    }
  }
}

template <bool kDrop, typename TIn>
static void launch_ln_rows(const TIn* x, const float* w, const float* b, __nv_bfloat16* y, int rows, int cols, float eps, DropoutCfg drop,
                           cudaStream_t stream) {
  const int wpb = 8;
  const int want = (rows + 2 * wpb - 1) / (2 * wpb);            // one pass of two rows per warp
  const int grid = std::min(want, 2 * sm_count());              // two resident CTAs per SM (launch bounds)
  const int kv = (cols + 127) / 128;
  if (kv <= 2) layernorm_f32_to_bf16_kernel<kDrop, TIn, 2><<<grid, wpb * 32, 0, stream>>>(x, w, b, y, rows, cols, eps, drop);
  else if (kv <= 4) layernorm_f32_to_bf16_kernel<kDrop, TIn, 4><<<grid, wpb * 32, 0, stream>>>(x, w, b, y, rows, cols, eps, drop);
  else if (kv <= 6) layernorm_f32_to_bf16_kernel<kDrop, TIn, 6><<<grid, wpb * 32, 0, stream>>>(x, w, b, y, rows, cols, eps, drop);
  else layernorm_f32_to_bf16_kernel<kDrop, TIn, 8><<<grid, wpb * 32, 0, stream>>>(x, w, b, y, rows, cols, eps, drop);
}

int launch_layernorm_f32_to_bf16(const float* x, const float* w, const float* b, svol_bf16* y, int rows, int cols,
                                 float eps, float drop_p, const long long* seed, int site, cudaStream_t stream) {
  if (cols % 4 != 0 || cols > LN_MAXV * 128 || rows <= 0) return svol_fail(SVOL_ERR_SHAPE, "layernorm: cols % 4 == 0, cols <= 1024");
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !seed)) return svol_fail(SVOL_ERR_SHAPE, "layernorm: 0 <= drop_p < 1, seed required");
  const DropoutCfg drop{drop_p, seed, site};
  if (drop_p > 0.f) launch_ln_rows<true, float>(x, w, b, reinterpret_cast<__nv_bfloat16*>(y), rows, cols, eps, drop, stream);
  else launch_ln_rows<false, float>(x, w, b, reinterpret_cast<__nv_bfloat16*>(y), rows, cols, eps, drop, stream);
  return svol_check_launch("layernorm_f32_to_bf16");
}

// Same LayerNorm for frame features the caller keeps in bf16 (a precomputed feature cache): half the bytes to read and,
// when they come from the host, half the PCIe traffic of the step.
int launch_layernorm_bf16_to_bf16(const svol_bf16* x, const float* w, const float* b, svol_bf16* y, int rows, int cols,
                                  float eps, cudaStream_t stream) {
  if (cols % 4 != 0 || cols > LN_MAXV * 128 || rows <= 0) return svol_fail(SVOL_ERR_SHAPE, "layernorm: cols % 4 == 0, cols <= 1024");
  launch_ln_rows<false, __nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(x), w, b, reinterpret_cast<__nv_bfloat16*>(y), rows, cols, eps,
                                       DropoutCfg{0.f, nullptr, 0}, stream);
  return svol_check_launch("layernorm_bf16_to_bf16");
}

// ---------------------------------------------------------------------------------------------
// Backbone hand-off (backbone.py:72-89, model.py:18-22): the ResNet trunk emits (N*T, C, h, w); the reference
// reshapes / transposes it twice into (N, T*h*w, C) tokens (a strided copy of the whole 100 MB tensor) before the
// head's first LayerNorm reads it again.  This kernel reads the channel-major feature map of one frame per CTA
// (one fully coalesced pass into shared memory) and writes the LayerNorm-ed bf16 TOKEN rows directly: the permuted
// fp32 copy never exists.  x [F, C, S] fp32 (S = h*w), y [F*S, C] bf16; token row = f*S + s, as the reference's
// flatten(2).transpose(1,2).reshape(N, -1, C) orders them.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_nchw_to_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                                     int C, int S, float eps) {
  extern __shared__ __align__(16) float tile[];    // [C][S] of this frame
  __shared__ uint64_t bar;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t n = static_cast<size_t>(C) * S;
  const float* xf = x + static_cast<size_t>(f) * n;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(xf) & 15) == 0) {
    // the frame's C*S floats are contiguous: bulk-async copies (no register staging, completion on an mbarrier); the
    // second resident CTA of the SM computes while this one's copy is in flight
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(n * 4);
      mbar_arrive_expect_tx(&bar, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768)
        bulk_load_1d(reinterpret_cast<uint8_t*>(tile) + off, reinterpret_cast<const uint8_t*>(xf) + off, min(32768u, bytes - off), &bar);
    }
    mbar_wait(&bar, 0);
  } else {
    for (size_t i = tid; i < n; i += 256) tile[i] = __ldcs(xf + i);
    __syncthreads();
  }
  const int per_lane = (C + 31) / 32;               // <= 32 (C <= 1024)
  for (int s = warp; s < S; s += 8) {
    float v[32];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int c = lane + 32 * k;
      v[k] = (k < per_lane && c < C) ? tile[static_cast<size_t>(c) * S + s] : 0.f;     // stride S floats: S odd -> conflict-free
      sum += v[k];
    }
    const float mean = warp_sum(sum) / C;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int c = lane + 32 * k;
      if (k < per_lane && c < C) { const float d = v[k] - mean; ss += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    __nv_bfloat16* yr = y + (static_cast<size_t>(f) * S + s) * C;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int c = lane + 32 * k;
      if (k < per_lane && c < C) yr[c] = __float2bfloat16_rn((v[k] - mean) * rstd * __ldg(w + c) + __ldg(b + c));
    }
  }
}

int launch_layernorm_nchw_to_bf16(const float* x, const float* w, const float* b, svol_bf16* y, int frames, int C, int S, float eps,
                                  cudaStream_t stream) {
  if (frames <= 0 || C <= 0 || C > 1024 || S <= 0) return svol_fail(SVOL_ERR_SHAPE, "layernorm_nchw: C <= 1024");
  const size_t smem = static_cast<size_t>(C) * S * sizeof(float);
  if (smem > 200 * 1024) return svol_fail(SVOL_ERR_SHAPE, "layernorm_nchw: C * h * w * 4 bytes must fit in shared memory (200 KB)");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(layernorm_nchw_to_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "layernorm_nchw: cudaFuncSetAttribute");
    configured = smem;
  }
  layernorm_nchw_to_bf16_kernel<<<frames, 256, smem, stream>>>(x, w, b, reinterpret_cast<__nv_bfloat16*>(y), C, S, eps);
  return svol_check_launch("layernorm_nchw_to_bf16");
}

// ---------------------------------------------------------------------------------------------
// y = [ReLU](Linear(LayerNorm(x))) in fp32, one CTA per row (sketch branch, B rows only)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ lw,
                                                            const float* __restrict__ lb, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int relu,
                                                            float* __restrict__ y, int in_dim, int out_dim, float eps,
                                                            DropoutCfg drop) {
  extern __shared__ float xs[];   // in_dim normalised inputs + 16 scratch
  float* red = xs + in_dim;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xr = x + static_cast<size_t>(row) * in_dim;
  float s = 0.f;
  for (int i = tid; i < in_dim; i += blockDim.x) { const float v = xr[i]; xs[i] = v; s += v; }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float mean = tot / in_dim;
  __syncthreads();
  float ss = 0.f;
  for (int i = tid; i < in_dim; i += blockDim.x) { const float d = xs[i] - mean; ss += d * d; }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / in_dim + eps);
  for (int i = tid; i < in_dim; i += blockDim.x) xs[i] = (xs[i] - mean) * rstd * lw[i] + lb[i];
  if (drop.p > 0.f) {   // train-mode Dropout between the LayerNorm and the Linear (svanet.py:168-170)
    const unsigned long long key = dropout_key(drop);
    const uint32_t thr = dropout_threshold(drop.p);
    const float sc = 1.0f / (1.0f - drop.p);
    for (int i = tid; i < in_dim; i += blockDim.x)
      xs[i] = dropout_keep(static_cast<unsigned long long>(row) * in_dim + i, key, thr) ? xs[i] * sc : 0.f;
  }
  __syncthreads();
  const int o_end = min(out_dim, static_cast<int>(blockIdx.y + 1) * 32);
  for (int o = blockIdx.y * 32 + warp; o < o_end; o += 8) {
    const float* wr = w + static_cast<size_t>(o) * in_dim;
    float acc = 0.f;
    for (int i = lane; i < in_dim; i += 32) acc = fmaf(xs[i], __ldg(wr + i), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += bias[o];
      y[static_cast<size_t>(row) * out_dim + o] = relu ? fmaxf(acc, 0.f) : acc;
    }
  }
}

int launch_ln_linear_f32(const float* x, const float* lw, const float* lb, const float* w, const float* b, int relu,
                         float* y, int rows, int in_dim, int out_dim, float eps, float drop_p, const long long* seed, int site,
                         cudaStream_t stream) {
  if (rows <= 0 || in_dim <= 0 || in_dim > 8192) return svol_fail(SVOL_ERR_SHAPE, "ln_linear: bad sizes");
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !seed)) return svol_fail(SVOL_ERR_SHAPE, "ln_linear: 0 <= drop_p < 1, seed required");
  const DropoutCfg drop{drop_p, seed, site};
  ln_linear_f32_kernel<<<dim3(rows, (out_dim + 31) / 32), 256, (in_dim + 16) * sizeof(float), stream>>>(x, lw, lb, w, b, relu, y, in_dim, out_dim, eps, drop);
  return svol_check_launch("ln_linear_f32");
}

// ---------------------------------------------------------------------------------------------
// Sine positional table (position_encoding.py:51-71).  Each CTA owns 32 token rows of one sample
// and recomputes the mask prefix sum it needs from the (L2-resident) mask row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) posenc_sine_kernel(const float* __restrict__ mask, float* __restrict__ pos,
                                                          int L, int d) {
  __shared__ float red[2][8];
  __shared__ float xrow[32];
  const int b = blockIdx.y, l0 = blockIdx.x * 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* mrow = mask + static_cast<size_t>(b) * L;
  float before = 0.f, total = 0.f;
  for (int l = tid; l < L; l += blockDim.x) {
    const float m = mrow[l] != 0.f ? 1.f : 0.f;
    total += m;
    if (l < l0) before += m;
  }
  before = warp_sum(before); total = warp_sum(total);
  if (lane == 0) { red[0][warp] = before; red[1][warp] = total; }
  __syncthreads();
  if (tid < 32) {
    float bsum = 0.f, tsum = 0.f;
    for (int i = 0; i < 8; ++i) { bsum += red[0][i]; tsum += red[1][i]; }
    float cum = bsum;                                      // exact: sums of 0/1 below 2^24
    for (int t = 0; t <= tid; ++t) if (l0 + t < L) cum += (mrow[l0 + t] != 0.f ? 1.f : 0.f);
    // x_embed / (x_embed[:, -1:] + eps) * scale, left to right in fp32
    xrow[tid] = __fmul_rn(__fdiv_rn(cum, __fadd_rn(tsum, 1e-6f)), 6.283185307179586f);
  }
  __syncthreads();
  // dim_t[i] = 10000^(2*(i/2)/d): one powf per column per CTA instead of one per element
  extern __shared__ float dim_t[];
  for (int i = tid; i < d; i += blockDim.x)
    dim_t[i] = powf(10000.f, __fdiv_rn(__fmul_rn(2.f, static_cast<float>(i >> 1)), static_cast<float>(d)));
  __syncthreads();
  const int rows = min(32, L - l0);
  for (int e = tid; e < rows * d; e += blockDim.x) {
    const int r = e / d, i = e - r * d;
    const float a = __fdiv_rn(xrow[r], dim_t[i]);
    pos[(static_cast<size_t>(b) * L + l0 + r) * d + i] = (i & 1) ? cosf(a) : sinf(a);
  }
}

int launch_posenc_sine(const float* mask, float* pos, int B, int L, int d, cudaStream_t stream) {
  if (B <= 0 || L <= 0 || d <= 0) return svol_fail(SVOL_ERR_SHAPE, "posenc: bad sizes");
  posenc_sine_kernel<<<dim3((L + 31) / 32, B), 256, d * sizeof(float), stream>>>(mask, pos, L, d);
  return svol_check_launch("posenc_sine");
}

// ---------------------------------------------------------------------------------------------
// theta[b,l] = cumsum(mask)[b,l] / (sum(mask[b]) + 1e-6) * 2 pi  (position_encoding.py:55-61), bit-identical to the
// xrow values of posenc_sine_kernel.  One CTA per sample: per-thread runs of consecutive tokens + a block scan.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) posenc_theta_kernel(const float* __restrict__ mask, float* __restrict__ theta, int L) {
  __shared__ float part[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* mrow = mask + static_cast<size_t>(b) * L;
  const int per = (L + 255) / 256, l0 = tid * per, l1 = min(L, l0 + per);
  float s = 0.f;
  for (int l = l0; l < l1; ++l) s += mrow[l] != 0.f ? 1.f : 0.f;
  part[tid] = s;
  __syncthreads();
  float before = 0.f, total = 0.f;                       // exact: sums of 0/1 below 2^24
  for (int i = 0; i < 256; ++i) { const float v = part[i]; total += v; if (i < tid) before += v; }
  float cum = before;
  for (int l = l0; l < l1; ++l) {
    cum += mrow[l] != 0.f ? 1.f : 0.f;
    theta[static_cast<size_t>(b) * L + l] = __fmul_rn(__fdiv_rn(cum, __fadd_rn(total, 1e-6f)), 6.283185307179586f);
  }
}

int launch_posenc_theta(const float* mask, float* theta, int B, int L, cudaStream_t stream) {
  if (B <= 0 || L <= 0) return svol_fail(SVOL_ERR_SHAPE, "posenc_theta: bad sizes");
  posenc_theta_kernel<<<B, 256, 0, stream>>>(mask, theta, L);
  return svol_check_launch("posenc_theta");
}

// ---------------------------------------------------------------------------------------------
// out[r,:] = bf16(x[r % mod,:] + pos[r % mod,:])
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_pos_bf16_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                           __nv_bfloat16* __restrict__ out, long long total4,
                                                           int cols4, int mod) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total4) return;
  const long long r = i / cols4;
  const int c = static_cast<int>(i - r * cols4);
  const long long src = (mod > 0 ? r % mod : r) * cols4 + c;
  float4 v = __ldg(reinterpret_cast<const float4*>(x) + src);
  if (pos) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + src);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
  }
  uint2 q;
  q.x = pack_bf16x2(v.x, v.y); q.y = pack_bf16x2(v.z, v.w);
  reinterpret_cast<uint2*>(out)[i] = q;
}

int launch_add_pos_bf16(const float* x, const float* pos, svol_bf16* out, int rows, int cols, int mod,
                        cudaStream_t stream) {
  if (cols % 4 != 0 || rows <= 0) return svol_fail(SVOL_ERR_SHAPE, "add_pos: cols % 4 == 0");
  const long long total4 = static_cast<long long>(rows) * (cols / 4);
  add_pos_bf16_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(
      x, pos, reinterpret_cast<__nv_bfloat16*>(out), total4, cols / 4, mod);
  return svol_check_launch("add_pos_bf16");
}

// ---------------------------------------------------------------------------------------------
// Sketch-conditioned gate (cross_modal_transformer.py:122-127)
// ---------------------------------------------------------------------------------------------
// u[b,h,:] = (1/sqrt(dh)) * sum_j (Wq[h*dh+j,:] . s_b + bq[h*dh+j]) * Wk[h*dh+j,:]
// (the key bias adds the same constant to every token's score and cancels in the softmax)
__global__ void __launch_bounds__(256) gate_vectors_kernel(const float* __restrict__ sketch,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ u, int d, int H) {
  extern __shared__ float sm[];          // d sketch values + dh q values
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dh = d / H;
  float* s = sm;
  float* q = sm + d;
  for (int i = tid; i < d; i += blockDim.x) s[i] = sketch[static_cast<size_t>(b) * d + i];
  __syncthreads();
  for (int j = warp; j < dh; j += 8) {
    const float* wr = w + static_cast<size_t>(h * dh + j) * d;
    float acc = 0.f;
    for (int i = lane; i < d; i += 32) acc = fmaf(s[i], __ldg(wr + i), acc);
    acc = warp_sum(acc);
    if (lane == 0) q[j] = (acc + bias[h * dh + j]) * rsqrtf(static_cast<float>(dh));
  }
  __syncthreads();
  for (int c = tid; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < dh; ++j) acc = fmaf(q[j], __ldg(w + static_cast<size_t>(d + h * dh + j) * d + c), acc);
    u[(static_cast<size_t>(b) * H + h) * d + c] = acc;
  }
}

int launch_gate_vectors(const float* sketch, const float* w, const float* b, float* u, int B, int d, int H,
                        cudaStream_t stream) {
  if (B <= 0 || d <= 0 || H <= 0 || d % H != 0) return svol_fail(SVOL_ERR_SHAPE, "gate_vectors: bad sizes");
  gate_vectors_kernel<<<dim3(H, B), 256, (d + d / H) * sizeof(float), stream>>>(sketch, w, b, u, d, H);
  return svol_check_launch("gate_vectors");
}

// scores[b,h,l] = (x+pos)[b,l,:] . u[b,h,:]; one warp per token, d = 256 (8 bf16 per lane), H = 8
constexpr int GATE_D = 256, GATE_H = 8, GATE_ROWS = 64;
constexpr int GATE_APPLY_ROWS = 128;   // rows per CTA of gate_apply: amortises the per-CTA softmax statistics pass

__global__ void __launch_bounds__(256) gate_scores_kernel(const __nv_bfloat16* __restrict__ xpos,
                                                          const float* __restrict__ u, float* __restrict__ scores,
                                                          int L) {
  const int b = blockIdx.y, l0 = blockIdx.x * GATE_ROWS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this lane's 8 columns of every head's gate vector stay in registers for the CTA's 64 rows
  float ur[GATE_H][8];
#pragma unroll
  for (int h = 0; h < GATE_H; ++h) {
    const float4* up = reinterpret_cast<const float4*>(u + (static_cast<size_t>(b) * GATE_H + h) * GATE_D) + lane * 2;
    const float4 a = __ldg(up), c = __ldg(up + 1);
    ur[h][0] = a.x; ur[h][1] = a.y; ur[h][2] = a.z; ur[h][3] = a.w;
    ur[h][4] = c.x; ur[h][5] = c.y; ur[h][6] = c.z; ur[h][7] = c.w;
  }
  // the loads of all of this warp's rows are issued up front (8 independent 16-byte loads per lane in flight):
  // one row at a time the kernel was bound by the load -> reduce -> store latency chain, not by bandwidth
  uint4 qrow[GATE_ROWS / 8];
#pragma unroll
  for (int i = 0; i < GATE_ROWS / 8; ++i) {
    const int l = l0 + warp + i * 8;
    qrow[i] = l < L ? __ldg(reinterpret_cast<const uint4*>(xpos + (static_cast<size_t>(b) * L + l) * GATE_D) + lane)
                    : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int ri = 0; ri < GATE_ROWS / 8; ++ri) {
    const int l = l0 + warp + ri * 8;
    if (l >= L) break;
    const uint4 q = qrow[ri];
    const float xv[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y), bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
    float acc[GATE_H];
#pragma unroll
    for (int h = 0; h < GATE_H; ++h) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a = fmaf(xv[i], ur[h][i], a);
      acc[h] = a;
    }
    // reduce-scatter butterfly: 8 heads x 32 lanes -> one fully reduced head per lane in 4+2+1+1+1 = 9 shuffles
    // (a plain butterfly per head needs 40): at offset 16 / 8 / 4 each lane hands the half of its values the partner
    // keeps and keeps the other half; offsets 2 and 1 finish the one remaining value.
    float a4[4], a2[2], a1;
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = b16 ? acc[i] : acc[i + 4], keep = b16 ? acc[i + 4] : acc[i];
      a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = b8 ? a4[i] : a4[i + 2], keep = b8 ? a4[i + 2] : a4[i];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
      const float send = b4 ? a2[0] : a2[1], keep = b4 ? a2[1] : a2[0];
      a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    const int head = (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0);      // every group of 4 lanes holds one head's score
    if ((lane & 3) == 0) scores[(static_cast<size_t>(b) * GATE_H + head) * L + l] = a1;
  }
}

int launch_gate_scores(const svol_bf16* xpos, const float* u, float* scores, int B, int L, int d, int H,
                       cudaStream_t stream) {
  if (d != GATE_D || H != GATE_H || B <= 0 || L <= 0) return svol_fail(SVOL_ERR_SHAPE, "gate_scores: hidden_dim 256 / 8 heads only");
  gate_scores_kernel<<<dim3((L + GATE_ROWS - 1) / GATE_ROWS, B), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(xpos), u, scores, L);
  return svol_check_launch("gate_scores");
}

// att[b,l] = mean_h softmax_l(scores[b,h,:]); mem = LN1(x + att*x); mem_pos = mem + pos
// kTheta: `pos` is the [B*L] angle array of svol_posenc_theta and the sine encoding is evaluated in place
// (columns (2k, 2k+1) = (sin, cos)(theta / 10000^(2k/d))) instead of reading the 1 KB/token fp32 table.
template <bool kTheta>
__global__ void __launch_bounds__(256) gate_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                         const float* __restrict__ scores,
                                                         const float* __restrict__ lw, const float* __restrict__ lb,
                                                         const float* __restrict__ pos, __nv_bfloat16* __restrict__ mem,
                                                         __nv_bfloat16* __restrict__ mem_pos, float* __restrict__ att_out,
                                                         int L, float eps) {
  __shared__ float smax[GATE_H], sinv[GATE_H];
  const int b = blockIdx.y, l0 = blockIdx.x * GATE_APPLY_ROWS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    // warp h: softmax statistics of head h over all L tokens of this sample (L2-resident re-read)
    const float* sr = scores + (static_cast<size_t>(b) * GATE_H + warp) * L;
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) m = fmaxf(m, sr[l]);
    m = warp_max(m);
    float s = 0.f;
    for (int l = lane; l < L; l += 32) s += expf(sr[l] - m);
    s = warp_sum(s);
    if (lane == 0) { smax[warp] = m; sinv[warp] = 1.f / s; }
  }
  __syncthreads();
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(lw) + lane * 2), g1 = __ldg(reinterpret_cast<const float4*>(lw) + lane * 2 + 1);
  const float4 o0 = __ldg(reinterpret_cast<const float4*>(lb) + lane * 2), o1 = __ldg(reinterpret_cast<const float4*>(lb) + lane * 2 + 1);
  const float gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float gb[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
  float idt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    idt[i] = 1.0f / powf(10000.f, __fdiv_rn(__fmul_rn(2.f, static_cast<float>(lane * 4 + i)), static_cast<float>(GATE_D)));
  // software pipeline: the next row's token and score loads are in flight while this row is normalised
  auto load_row = [&](int l, uint4& q, float& sc) {
    q = make_uint4(0u, 0u, 0u, 0u); sc = 0.f;
    if (l < L) {
      q = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * L + l) * GATE_D) + lane);
      if (lane < GATE_H) sc = __ldg(scores + (static_cast<size_t>(b) * GATE_H + lane) * L + l);
    }
  };
  uint4 q_next; float sc_next;
  load_row(l0 + warp, q_next, sc_next);
  for (int r = warp; r < GATE_APPLY_ROWS; r += 8) {
    const int l = l0 + r;
    if (l >= L) break;
    const size_t row = static_cast<size_t>(b) * L + l;
    const uint4 q = q_next;
    const float sc = sc_next;
    load_row(l + 8 < l0 + GATE_APPLY_ROWS ? l + 8 : L, q_next, sc_next);
    float a = 0.f;
    if (lane < GATE_H) a = expf(sc - smax[lane]) * sinv[lane];
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    const float att = __shfl_sync(0xffffffffu, a, 0) * (1.0f / GATE_H);
    if (att_out && lane == 0) att_out[row] = att;
    float v[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y), bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = v[i] + att * v[i]; s += v[i]; }
    const float mean = warp_sum(s) * (1.0f / GATE_D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float dlt = v[i] - mean; ss += dlt * dlt; }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / GATE_D) + eps);
    float pp[8];
    if (kTheta) {
      const float theta = __ldg(pos + row);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = theta * idt[i];
        a = a > 3.14159265358979f ? a - 6.28318530717959f : a;     // theta in [0, 2 pi] -> [-pi, pi] for MUFU sin / cos
        pp[2 * i] = __sinf(a);
        pp[2 * i + 1] = __cosf(a);
      }
    } else {
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + row * GATE_D) + lane * 2);
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + row * GATE_D) + lane * 2 + 1);
      pp[0] = p0.x; pp[1] = p0.y; pp[2] = p0.z; pp[3] = p0.w; pp[4] = p1.x; pp[5] = p1.y; pp[6] = p1.z; pp[7] = p1.w;
    }
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (v[i] - mean) * rstd * gw[i] + gb[i];
    uint4 o, op;
    o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]); o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
    op.x = pack_bf16x2(y[0] + pp[0], y[1] + pp[1]); op.y = pack_bf16x2(y[2] + pp[2], y[3] + pp[3]);
    op.z = pack_bf16x2(y[4] + pp[4], y[5] + pp[5]); op.w = pack_bf16x2(y[6] + pp[6], y[7] + pp[7]);
    reinterpret_cast<uint4*>(mem + row * GATE_D)[lane] = o;
    reinterpret_cast<uint4*>(mem_pos + row * GATE_D)[lane] = op;
  }
}

int launch_gate_apply(const svol_bf16* x, const float* scores, const float* lw, const float* lb, const float* pos,
                      svol_bf16* mem, svol_bf16* mem_pos, float* att_out, int B, int L, int d, int H, float eps,
                      bool pos_is_theta, cudaStream_t stream) {
  if (d != GATE_D || H != GATE_H || B <= 0 || L <= 0) return svol_fail(SVOL_ERR_SHAPE, "gate_apply: hidden_dim 256 / 8 heads only");
  const dim3 grid((L + GATE_APPLY_ROWS - 1) / GATE_APPLY_ROWS, B);
  if (pos_is_theta)
    gate_apply_kernel<true><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), scores, lw, lb, pos, reinterpret_cast<__nv_bfloat16*>(mem),
        reinterpret_cast<__nv_bfloat16*>(mem_pos), att_out, L, eps);
  else
    gate_apply_kernel<false><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), scores, lw, lb, pos, reinterpret_cast<__nv_bfloat16*>(mem),
        reinterpret_cast<__nv_bfloat16*>(mem_pos), att_out, L, eps);
  return svol_check_launch("gate_apply");
}

// ---------------------------------------------------------------------------------------------
// The whole gate in ONE launch (inference path): scores -> per-sample softmax statistics -> att -> LN1 -> mem, mem + pos.
// A cluster of GF_CL CTAs owns one sample; each CTA bulk-copies its contiguous slice of token rows into shared memory
// ONCE (cp.async.bulk, one mbarrier per 32-row chunk, everything in flight from the first cycle), computes the 8 head
// scores per row from x + pos (positions evaluated from the angle, fp32, never rounded to bf16), and the softmax over the
// sample's L tokens is closed with one exchange of (max, sum) pairs through distributed shared memory.  Against
// gate_scores + gate_apply this drops the second read of the token rows, the (x + pos) operand, the score round trip
// through HBM and one launch: compulsory traffic = B*L*d*2 read + 2*B*L*d*2 written.
// ---------------------------------------------------------------------------------------------
constexpr int GF_CHUNK = 32, GF_MAX_CHUNKS = 16, GF_ILP = 2, GF_ILP2 = 4;   // rows in flight per warp: score pass / LayerNorm pass
constexpr int GF_ROW_BYTES = GATE_D * 2;
constexpr int GF_MAX_ROWS = GF_CHUNK * GF_MAX_CHUNKS;                 // token rows per CTA
constexpr int GF_HDR_BYTES = 256;                                      // chunk barriers [16], local stats [8] float2, global [8] float2

__host__ __device__ inline int gf_pad4(int n) { return (n + 3) & ~3; }
__host__ __device__ inline size_t gf_smem_bytes(int rpc) {
  return GF_HDR_BYTES + static_cast<size_t>(gf_pad4(rpc)) * (1 + GATE_H) * 4 + static_cast<size_t>(rpc) * GF_ROW_BYTES;
}

// columns (2k, 2k+1) = (sin, cos)(theta / 10000^(2k/d)) for this lane's four k (position_encoding.py:62-70)
__device__ __forceinline__ void gate_sincos(float theta, const float (&idt)[4], float (&pp)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a = theta * idt[i];
    a = a > 3.14159265358979f ? a - 6.28318530717959f : a;     // theta in [0, 2 pi] -> [-pi, pi] for MUFU sin / cos
    pp[2 * i] = __sinf(a);
    pp[2 * i + 1] = __cosf(a);
  }
}

// kCL CTAs per sample (cluster), kThreads threads per CTA: <8, 256> runs two CTAs per SM (<= 113 KB of rows each),
// <4, 512> one CTA per SM; both keep 16 warps and the same number of token rows per SM.
template <int kCL, int kThreads>
__global__ void __launch_bounds__(kThreads, 512 / kThreads)
gate_fused_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ u, const float* __restrict__ lw,
                  const float* __restrict__ lb, const float* __restrict__ theta, __nv_bfloat16* __restrict__ mem,
                  __nv_bfloat16* __restrict__ mem_pos, float* __restrict__ att_out, float* __restrict__ scores_out,
                  int L, int rpc, float eps) {
  constexpr int NW = kThreads / 32;
  extern __shared__ __align__(128) uint8_t gf_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(gf_smem);                    // [GF_MAX_CHUNKS]
  float2* stat_local = reinterpret_cast<float2*>(gf_smem + 128);            // [GATE_H] (max, sum exp) over this CTA's rows
  float2* stat_all = reinterpret_cast<float2*>(gf_smem + 192);              // [GATE_H] (max, 1 / sum exp) over the sample
  const int rpc_pad = gf_pad4(rpc);
  float* th_s = reinterpret_cast<float*>(gf_smem + GF_HDR_BYTES);           // [rpc_pad]
  float* sc_s = th_s + rpc_pad;                                             // [GATE_H][rpc_pad]
  uint4* rows_s = reinterpret_cast<uint4*>(sc_s + GATE_H * rpc_pad);        // [rpc][32]

  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = static_cast<int>(cluster_ctarank());
  const int l_begin = rank * rpc;
  const int n_rows = max(0, min(L - l_begin, rpc));
  const size_t row0 = static_cast<size_t>(b) * L + l_begin;
  const int n_chunks = (n_rows + GF_CHUNK - 1) / GF_CHUNK;

  if (tid == 0) {
    for (int c = 0; c < n_chunks; ++c) mbar_init(&bars[c], 1);
    fence_barrier_init();
    for (int c = 0; c < n_chunks; ++c) {
      const int nr = min(GF_CHUNK, n_rows - c * GF_CHUNK);
      mbar_arrive_expect_tx(&bars[c], nr * GF_ROW_BYTES);
      bulk_load_1d(rows_s + c * GF_CHUNK * 32, x + (row0 + c * GF_CHUNK) * GATE_D, nr * GF_ROW_BYTES, &bars[c]);
    }
  }
  for (int i = tid; i < n_rows; i += kThreads) th_s[i] = __ldg(theta + row0 + i);
  // this lane's 8 columns of every head's gate vector stay in registers
  float ur[GATE_H][8];
#pragma unroll
  for (int h = 0; h < GATE_H; ++h) {
    const float4* up = reinterpret_cast<const float4*>(u + (static_cast<size_t>(b) * GATE_H + h) * GATE_D) + lane * 2;
    const float4 a = __ldg(up), c = __ldg(up + 1);
    ur[h][0] = a.x; ur[h][1] = a.y; ur[h][2] = a.z; ur[h][3] = a.w;
    ur[h][4] = c.x; ur[h][5] = c.y; ur[h][6] = c.z; ur[h][7] = c.w;
  }
  float idt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)     // 10000^(-2k/d) = 2^(-k * log2(10000) * 2 / d), k = 4 lane + i
    idt[i] = exp2f(-static_cast<float>(lane * 4 + i) * (13.287712379549449f * 2.0f / GATE_D));
  __syncthreads();          // barrier inits and th_s visible

  // ---- phase 1: scores of this CTA's rows (warp per row, GF_ILP independent rows in flight per warp)
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
  const int head = (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0);
  uint32_t chunks_seen = 0;
  // rows of iteration i + 1 are fetched from shared memory before iteration i is computed
  auto fetch = [&](int lr0, uint4 (&q)[GF_ILP], float (&th)[GF_ILP]) {
#pragma unroll
    for (int r = 0; r < GF_ILP; ++r) {
      const int lr = lr0 + r * NW;
      q[r] = make_uint4(0u, 0u, 0u, 0u); th[r] = 0.f;
      if (lr < n_rows) {
        const int c = lr / GF_CHUNK;
        if (!((chunks_seen >> c) & 1u)) { mbar_wait(&bars[c], 0); chunks_seen |= 1u << c; }
        q[r] = rows_s[lr * 32 + lane];
        th[r] = th_s[lr];
      }
    }
  };
  uint4 q_next[GF_ILP];
  float th_next[GF_ILP];
  fetch(warp, q_next, th_next);
  for (int lr0 = warp; lr0 < n_rows; lr0 += GF_ILP * NW) {
    uint4 q[GF_ILP];
    float th[GF_ILP];
#pragma unroll
    for (int r = 0; r < GF_ILP; ++r) { q[r] = q_next[r]; th[r] = th_next[r]; }
    fetch(lr0 + GF_ILP * NW, q_next, th_next);
    float a1[GF_ILP];
#pragma unroll
    for (int r = 0; r < GF_ILP; ++r) {
      float pp[8];
      gate_sincos(th[r], idt, pp);
      const float xv[8] = {bf16_lo(q[r].x) + pp[0], bf16_hi(q[r].x) + pp[1], bf16_lo(q[r].y) + pp[2], bf16_hi(q[r].y) + pp[3],
                           bf16_lo(q[r].z) + pp[4], bf16_hi(q[r].z) + pp[5], bf16_lo(q[r].w) + pp[6], bf16_hi(q[r].w) + pp[7]};
      float acc[GATE_H];
#pragma unroll
      for (int h = 0; h < GATE_H; ++h) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a = fmaf(xv[i], ur[h][i], a);
        acc[h] = a;
      }
      // reduce-scatter butterfly (see gate_scores_kernel): every group of 4 lanes ends up with one head's score
      float a4[4], a2[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = b16 ? acc[i] : acc[i + 4], keep = b16 ? acc[i + 4] : acc[i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = b8 ? a4[i] : a4[i + 2], keep = b8 ? a4[i + 2] : a4[i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      const float send = b4 ? a2[0] : a2[1], keep = b4 ? a2[1] : a2[0];
      float t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      a1[r] = t;
    }
#pragma unroll
    for (int r = 0; r < GF_ILP; ++r) {
      const int lr = lr0 + r * NW;
      if (lr < n_rows && (lane & 3) == 0) {
        sc_s[head * rpc_pad + lr] = a1[r];
        if (scores_out) scores_out[(static_cast<size_t>(b) * GATE_H + head) * L + l_begin + lr] = a1[r];
      }
    }
  }
  __syncthreads();

  // ---- softmax statistics: warp h reduces head h over this CTA's rows, then the cluster exchanges (max, sum) pairs
  if (warp < GATE_H) {
    const float* sr = sc_s + warp * rpc_pad;
    float m = -INFINITY;
    for (int i = lane; i < n_rows; i += 32) m = fmaxf(m, sr[i]);
    m = warp_max(m);
    float sum = 0.f;
    for (int i = lane; i < n_rows; i += 32) sum += expf(sr[i] - m);
    sum = warp_sum(sum);
    if (lane == 0) stat_local[warp] = make_float2(m, n_rows > 0 ? sum : 0.f);
  }
  cluster_sync_all();
  if (tid < GATE_H * kCL) {
    const int h = tid / kCL, c = tid % kCL;
    uint32_t remote;
    float m_c, s_c;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&stat_local[h])), "r"(c));
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(m_c), "=f"(s_c) : "r"(remote) : "memory");
    float m = m_c;
#pragma unroll
    for (int o = kCL / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = m_c == -INFINITY ? 0.f : s_c * expf(m_c - m);
#pragma unroll
    for (int o = kCL / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (c == 0) stat_all[h] = make_float2(m, 1.f / sum);
  }
  // second cluster phase: arrive now (the remote reads above are done), wait before exit, so that no CTA's shared
  // memory disappears while a peer may still be reading its statistics
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  __syncthreads();

  // ---- phase 2: att, LN1(x (1 + att)), + pos
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(lw) + lane * 2), g1 = __ldg(reinterpret_cast<const float4*>(lw) + lane * 2 + 1);
  const float4 o0 = __ldg(reinterpret_cast<const float4*>(lb) + lane * 2), o1 = __ldg(reinterpret_cast<const float4*>(lb) + lane * 2 + 1);
  const float gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float gb[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
  const float2 st = lane < GATE_H ? stat_all[lane] : make_float2(0.f, 0.f);
  // GF_ILP2 rows per warp in flight, every stage (att, mean, variance) written across the rows so that the shuffle
  // chains of different rows overlap; a tail slot recomputes the last row and skips the stores
  for (int lr0 = warp; lr0 < n_rows; lr0 += GF_ILP2 * NW) {
    uint4 q[GF_ILP2];
    float a[GF_ILP2], thv[GF_ILP2];
#pragma unroll
    for (int r = 0; r < GF_ILP2; ++r) {
      const int lr = min(lr0 + r * NW, n_rows - 1);
      q[r] = rows_s[lr * 32 + lane];
      thv[r] = th_s[lr];
      a[r] = lane < GATE_H ? expf(sc_s[lane * rpc_pad + lr] - st.x) * st.y : 0.f;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < GF_ILP2; ++r) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
    float v[GF_ILP2][8], red[GF_ILP2];
#pragma unroll
    for (int r = 0; r < GF_ILP2; ++r) {
      const float att = __shfl_sync(0xffffffffu, a[r], 0) * (1.0f / GATE_H);
      if (att_out && lane == 0 && lr0 + r * NW < n_rows) att_out[row0 + lr0 + r * NW] = att;
      const float x8[8] = {bf16_lo(q[r].x), bf16_hi(q[r].x), bf16_lo(q[r].y), bf16_hi(q[r].y),
                           bf16_lo(q[r].z), bf16_hi(q[r].z), bf16_lo(q[r].w), bf16_hi(q[r].w)};
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[r][i] = x8[i] + att * x8[i]; s += v[r][i]; }
      red[r] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < GF_ILP2; ++r) red[r] += __shfl_xor_sync(0xffffffffu, red[r], o);
    float mean[GF_ILP2];
#pragma unroll
    for (int r = 0; r < GF_ILP2; ++r) {
      mean[r] = red[r] * (1.0f / GATE_D);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float dlt = v[r][i] - mean[r]; ss += dlt * dlt; }
      red[r] = ss;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < GF_ILP2; ++r) red[r] += __shfl_xor_sync(0xffffffffu, red[r], o);
#pragma unroll
    for (int r = 0; r < GF_ILP2; ++r) {
      const float rstd = rsqrtf(red[r] * (1.0f / GATE_D) + eps);
      float pp[8];
      gate_sincos(thv[r], idt, pp);
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = (v[r][i] - mean[r]) * rstd * gw[i] + gb[i];
      uint4 o, op;
      o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]); o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
      op.x = pack_bf16x2(y[0] + pp[0], y[1] + pp[1]); op.y = pack_bf16x2(y[2] + pp[2], y[3] + pp[3]);
      op.z = pack_bf16x2(y[4] + pp[4], y[5] + pp[5]); op.w = pack_bf16x2(y[6] + pp[6], y[7] + pp[7]);
      if (lr0 + r * NW < n_rows) {
        const size_t row = row0 + lr0 + r * NW;
        reinterpret_cast<uint4*>(mem + row * GATE_D)[lane] = o;
        reinterpret_cast<uint4*>(mem_pos + row * GATE_D)[lane] = op;
      }
    }
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// cluster size for a launch: 4 CTAs of 512 threads (one per SM) when a quarter of a sample fits one CTA's shared memory
// and the batch still fills the GPU, else 8 CTAs of 256 threads (two per SM); 0 = the rows do not fit (L > 3384)
static int gate_fused_cluster(int B, int L) {
  if (L <= 0) return 0;
  static const int forced = [] { const char* e = getenv("SVOL_GATE_CL"); return e ? atoi(e) : 0; }();
  for (int cl : {4, 8}) {
    const int rpc = (L + cl - 1) / cl;
    const bool fits = rpc <= GF_MAX_ROWS && gf_smem_bytes(rpc) <= (cl == 4 ? 227u : 113u) * 1024u;
    if (forced == cl && rpc <= GF_MAX_ROWS && gf_smem_bytes(rpc) <= 227u * 1024u) return cl;
    if (forced == 0 && fits && (cl == 8 || B * 4 >= 96)) return cl;
  }
  const int rpc = (L + 7) / 8;       // 8 CTAs, one per SM
  return (rpc <= GF_MAX_ROWS && gf_smem_bytes(rpc) <= 227u * 1024u) ? 8 : 0;
}

// 1 when the fused kernel can keep a sample's token rows in the shared memory of one cluster (L <= 3384)
int gate_fused_supported(int L) { return gate_fused_cluster(1, L) != 0; }

template <int kCL, int kThreads>
static int launch_gate_fused_t(const svol_bf16* x, const float* u, const float* lw, const float* lb, const float* theta,
                               svol_bf16* mem, svol_bf16* mem_pos, float* att_out, float* scores_out, int B, int L, float eps,
                               cudaStream_t stream) {
  const int rpc = (L + kCL - 1) / kCL;
  const size_t smem = gf_smem_bytes(rpc);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gate_fused_kernel<kCL, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return svol_fail_cuda(e, "gate_fused: cudaFuncSetAttribute");
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCL, B);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (getenv("SVOL_DEBUG_OCC")) {
    int n = -1;
    cudaOccupancyMaxActiveClusters(&n, gate_fused_kernel<kCL, kThreads>, &cfg);
    fprintf(stderr, "gate_fused<%d,%d>: %d clusters requested, %d co-resident, %zu B smem\n", kCL, kThreads, B, n, smem);
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gate_fused_kernel<kCL, kThreads>, reinterpret_cast<const __nv_bfloat16*>(x), u, lw, lb,
                                     theta, reinterpret_cast<__nv_bfloat16*>(mem), reinterpret_cast<__nv_bfloat16*>(mem_pos),
                                     att_out, scores_out, L, rpc, eps);
  if (e != cudaSuccess) return svol_fail_cuda(e, "gate_fused: cluster launch");
  return svol_check_launch("gate_fused");
}

int launch_gate_fused(const svol_bf16* x, const float* u, const float* lw, const float* lb, const float* theta, svol_bf16* mem,
                      svol_bf16* mem_pos, float* att_out, float* scores_out, int B, int L, int d, int H, float eps,
                      cudaStream_t stream) {
  if (d != GATE_D || H != GATE_H || B <= 0 || L <= 0) return svol_fail(SVOL_ERR_SHAPE, "gate_fused: hidden_dim 256 / 8 heads only");
  const int cl = gate_fused_cluster(B, L);
  if (cl == 0) return svol_fail(SVOL_ERR_SHAPE, "gate_fused: a sample's tokens do not fit one cluster's shared memory (use gate_scores + gate_apply)");
  if (reinterpret_cast<uintptr_t>(x) % 16 != 0) return svol_fail(SVOL_ERR_SHAPE, "gate_fused: x must be 16-byte aligned");
  if (cl == 4) return launch_gate_fused_t<4, 512>(x, u, lw, lb, theta, mem, mem_pos, att_out, scores_out, B, L, eps, stream);
  return launch_gate_fused_t<8, 256>(x, u, lw, lb, theta, mem, mem_pos, att_out, scores_out, B, L, eps, stream);
}

// ---------------------------------------------------------------------------------------------
// Output heads (svanet.py:125-127): class logits and sigmoid of the last box-MLP layer
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) heads_kernel(const __nv_bfloat16* __restrict__ hs,
                                                    const __nv_bfloat16* __restrict__ h2, const float* __restrict__ wc,
                                                    const float* __restrict__ bc, const float* __restrict__ wb,
                                                    const float* __restrict__ bb, float* __restrict__ logits,
                                                    float* __restrict__ boxes, int rows) {
  // Persistent warps: a lane keeps its 8 columns of the six weight rows in registers (the one-row-per-warp form staged the
  // 6 KB of weights through shared memory in every CTA of 8 rows: 15 MB of L2 reads and a block barrier per 8 rows), two
  // rows in flight per warp, and the six dot products are reduced together (8 values -> reduce-scatter butterfly).
  const int lane = threadIdx.x & 31;
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  float w[6][8];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4* wp = reinterpret_cast<const float4*>((j < 2 ? wc + j * GATE_D : wb + (j - 2) * GATE_D)) + lane * 2;
    const float4 p = __ldg(wp), q = __ldg(wp + 1);
    w[j][0] = p.x; w[j][1] = p.y; w[j][2] = p.z; w[j][3] = p.w; w[j][4] = q.x; w[j][5] = q.y; w[j][6] = q.z; w[j][7] = q.w;
  }
  const float bias_c0 = __ldg(bc), bias_c1 = __ldg(bc + 1);
  const float4 bias_b = make_float4(__ldg(bb), __ldg(bb + 1), __ldg(bb + 2), __ldg(bb + 3));   // (no alignment assumed)
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
  for (int row0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row0 < rows; row0 += 2 * n_warps) {
    uint4 qa[2], qb[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = min(row0 + r * n_warps, rows - 1);
      qa[r] = __ldg(reinterpret_cast<const uint4*>(hs + static_cast<size_t>(row) * GATE_D) + lane);
      qb[r] = __ldg(reinterpret_cast<const uint4*>(h2 + static_cast<size_t>(row) * GATE_D) + lane);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float a[8] = {bf16_lo(qa[r].x), bf16_hi(qa[r].x), bf16_lo(qa[r].y), bf16_hi(qa[r].y), bf16_lo(qa[r].z), bf16_hi(qa[r].z), bf16_lo(qa[r].w), bf16_hi(qa[r].w)};
      const float c[8] = {bf16_lo(qb[r].x), bf16_hi(qb[r].x), bf16_lo(qb[r].y), bf16_hi(qb[r].y), bf16_lo(qb[r].z), bf16_hi(qb[r].z), bf16_lo(qb[r].w), bf16_hi(qb[r].w)};
      float acc[8];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(j < 2 ? a[i] : c[i], w[j][i], s);
        acc[j] = s;
      }
      acc[6] = 0.f; acc[7] = 0.f;
      // reduce-scatter: after three exchange steps every group of 4 lanes holds one output's partial sum
      float a4[4], a2[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = b16 ? acc[i] : acc[i + 4], keep = b16 ? acc[i + 4] : acc[i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = b8 ? a4[i] : a4[i + 2], keep = b8 ? a4[i + 2] : a4[i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      const float send = b4 ? a2[0] : a2[1], keep = b4 ? a2[1] : a2[0];
      float t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      // lane group g = lane >> 2 holds output (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0); gather the six on lane 0
      const float o0 = t, o1 = __shfl_sync(0xffffffffu, t, 4), o2 = __shfl_sync(0xffffffffu, t, 8), o3 = __shfl_sync(0xffffffffu, t, 12),
                  o4 = __shfl_sync(0xffffffffu, t, 16), o5 = __shfl_sync(0xffffffffu, t, 20);
      const int row = row0 + r * n_warps;
      if (lane == 0 && row < rows) {
        reinterpret_cast<float2*>(logits)[row] = make_float2(o0 + bias_c0, o1 + bias_c1);
        float4 o;
        o.x = 1.f / (1.f + expf(-(o2 + bias_b.x))); o.y = 1.f / (1.f + expf(-(o3 + bias_b.y)));
        o.z = 1.f / (1.f + expf(-(o4 + bias_b.z))); o.w = 1.f / (1.f + expf(-(o5 + bias_b.w)));
        reinterpret_cast<float4*>(boxes)[row] = o;
      }
    }
  }
}

int launch_heads(const svol_bf16* hs, const svol_bf16* h2, const float* wc, const float* bc, const float* wb,
                 const float* bb, float* logits, float* boxes, int rows, int d, cudaStream_t stream) {
  if (d != GATE_D || rows <= 0) return svol_fail(SVOL_ERR_SHAPE, "heads: hidden_dim 256 only");
  const int want = (rows + 15) / 16;                         // 8 warps x 2 rows per pass
  const int grid = std::min(want, 4 * sm_count());
  heads_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(hs),
                                         reinterpret_cast<const __nv_bfloat16*>(h2), wc, bc, wb, bb, logits,
                                         boxes, rows);
  return svol_check_launch("heads");
}

// ---------------------------------------------------------------------------------------------
// Inference post-processing (test.py:133-158): one CTA per (video, frame)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) postprocess_kernel(const float* __restrict__ logits,
                                                          const float* __restrict__ boxes, float* __restrict__ out,
                                                          int32_t* __restrict__ order, int Q, int qf) {
  extern __shared__ float sc[];   // qf scores
  const int frame = blockIdx.x, b = blockIdx.y;
  const size_t base = static_cast<size_t>(b) * Q + static_cast<size_t>(frame) * qf;
  for (int i = threadIdx.x; i < qf; i += blockDim.x) {
    const float l0 = logits[(base + i) * 2], l1 = logits[(base + i) * 2 + 1];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    sc[i] = __fdiv_rn(e0, __fadd_rn(e0, e1));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < qf; i += blockDim.x) {
    const float s = sc[i];
    int rank = 0;
    for (int j = 0; j < qf; ++j) rank += (sc[j] > s) || (sc[j] == s && j < i);   // stable, descending
    const float4 bx = reinterpret_cast<const float4*>(boxes)[base + i];
    const float hw = __fmul_rn(0.5f, bx.z), hh = __fmul_rn(0.5f, bx.w);
    float* o = out + (base + rank) * 5;
    o[0] = fminf(fmaxf(__fsub_rn(bx.x, hw), 0.f), 1.f);
    o[1] = fminf(fmaxf(__fsub_rn(bx.y, hh), 0.f), 1.f);
    o[2] = fminf(fmaxf(__fadd_rn(bx.x, hw), 0.f), 1.f);
    o[3] = fminf(fmaxf(__fadd_rn(bx.y, hh), 0.f), 1.f);
    o[4] = s;
    order[base + rank] = i;
  }
}

int launch_postprocess(const float* logits, const float* boxes, float* out, int32_t* order, int B, int Q, int qf,
                       cudaStream_t stream) {
  if (B <= 0 || Q <= 0 || qf <= 0 || Q % qf != 0 || qf > 8192) return svol_fail(SVOL_ERR_SHAPE, "postprocess: Q % q_per_frame == 0");
  postprocess_kernel<<<dim3(Q / qf, B), 128, qf * sizeof(float), stream>>>(logits, boxes, out, order, Q, qf);
  return svol_check_launch("postprocess");
}

}  // namespace svol
