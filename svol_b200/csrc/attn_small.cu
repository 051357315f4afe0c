// Attention core for SHORT key sequences (Lk <= 512, no key-padding mask, inference): the object queries' self-attention,
// nn.MultiheadAttention at lib/modeling/cross_modal_transformer.py:147 (Q = 320 queries at the headline config: 3 key tiles).
//
// Why a second kernel.  attention_tc_kernel (attn_tc.cu) is organised around a long key loop: a CTA owns an SM (640
// threads, 210 KB of shared memory, all 512 tensor-memory columns), allocates tensor memory, fills a TMA ring and pays a
// ~2 k-clk pipeline fill, a 1.65 k-clk warpgroup stagger, a merge of split-key partial results and a 1.2 us launch gap per
// CTA.  With 13 key tiles per CTA that is 20 % of its life; with the 3 key tiles of the query self-attention it is most of
// it: 29 us per launch for 5 us of MUFU work (102 TFLOP/s).  The work per (sample, head) here is 320 x 320 scores --
// 13 MFLOP, nowhere near the tensor pipe's reach -- so this kernel drops the tensor-memory machinery altogether:
//   * one CTA = 64 query rows of one (sample, head), 4 warps x 16 rows, 128 threads; K and V^T of the head (<= 73 KB)
//     and the Q tile are copied to shared memory ONCE with plain 16-byte loads (rows padded by 16 bytes: ldmatrix reads
//     are conflict-free); 4-5 CTAs are resident per SM, so the block scheduler levels the tail at a 64-row granularity
//     and other CTAs' exponentials cover a CTA's load phase;
//   * scores and probabilities live in REGISTERS: S = Q K^T with mma.sync.m16n8k16 (bf16 in, fp32 out) per 64-key chunk,
//     online softmax (running maximum / sum per row, ex2 on pre-scaled scores), the accumulator fragments of S are
//     re-packed in place as the A operand of P V (the m16n8 C layout of two adjacent key blocks IS the m16n8k16 A
//     layout), O += P V with V^T rows as the col-major B operand -- no shared-memory or tensor-memory round trip
//     between the two products, no barrier inside the key loop.
// Numerics follow attention_tc_kernel: Q arrives pre-scaled by log2(e) / sqrt(dh), probabilities are rounded to bf16
// before P V, the row sum is accumulated in fp32 from the unrounded probabilities, the output is O / l in bf16.
#include <type_traits>

#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace attn_small {
constexpr int DH = 32, KC = 64;
constexpr int ROW_PITCH = DH * 2 + 16;          // bytes per Q / K row in shared memory (64 + 16: ldmatrix conflict-free)
constexpr int MAX_LK = 512;
__host__ __device__ inline int vt_pitch_bytes(int lk_pad) { return lk_pad * 2 + 16; }
__host__ __device__ inline size_t smem_bytes(int qb, int lk_pad) {
  return static_cast<size_t>(qb) * ROW_PITCH + static_cast<size_t>(lk_pad) * ROW_PITCH + static_cast<size_t>(DH) * vt_pitch_bytes(lk_pad);
}
}  // namespace attn_small

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
attention_small_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ vt,
                       __nv_bfloat16* __restrict__ out, int H, int Lq, int Lk, int ldq, int ldk, int vt_pitch, int ldo, int lk_pad) {
  using namespace attn_small;
  constexpr int QB = kWarps * 16, THREADS = kWarps * 32;
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* q_s = sm;                                        // [QB][ROW_PITCH]
  uint8_t* k_s = q_s + QB * ROW_PITCH;                      // [lk_pad][ROW_PITCH]
  uint8_t* v_s = k_s + static_cast<size_t>(lk_pad) * ROW_PITCH;   // [DH][vpb]: V^T, one row per head dimension
  const int vpb = vt_pitch_bytes(lk_pad);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * QB, h = blockIdx.y, b = blockIdx.z;
  griddep_wait();

  // ---- Q tile, K and V^T of this (sample, head) -> shared memory with cp.async (16 bytes each, all of a thread's ~22 copies
  // in flight at once: with load -> store pairs through registers the phase was a chain of exposed L2 round trips);
  // rows / keys out of range are zero-filled (source size 0)
  auto copy16 = [](uint8_t* dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
  };
  for (int i = tid; i < QB * 4; i += THREADS) {
    const int r = i >> 2, c = i & 3, row = min(q0 + r, Lq - 1);
    copy16(q_s + r * ROW_PITCH + c * 16, q + (static_cast<size_t>(b) * Lq + row) * ldq + h * DH + c * 8, q0 + r < Lq);
  }
  for (int i = tid; i < lk_pad * 4; i += THREADS) {
    const int r = i >> 2, c = i & 3, row = min(r, Lk - 1);
    copy16(k_s + r * ROW_PITCH + c * 16, k + (static_cast<size_t>(b) * Lk + row) * ldk + h * DH + c * 8, r < Lk);
  }
  const int chunks = lk_pad / 8;                             // 16-byte chunks (8 keys) per V^T row
  {
    const __nv_bfloat16* vbase = vt + (static_cast<size_t>(b) * H + h) * DH * vt_pitch;
    for (int i = tid; i < DH * chunks; i += THREADS) {
      const int d = i / chunks, c = i - d * chunks, key0 = c * 8;
      // (a chunk that starts below Lk ends at or below vt_pitch, a multiple of 8)
      copy16(v_s + d * vpb + c * 16, vbase + static_cast<size_t>(d) * vt_pitch + (key0 < Lk ? key0 : 0), key0 < Lk);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (Lk & 7) {                                              // V^T chunk straddling Lk: zero the keys beyond it (P is zero there too)
    __syncthreads();
    if (tid < DH) {
      const int c = Lk >> 3;
      uint32_t* w = reinterpret_cast<uint32_t*>(v_s + tid * vpb + c * 16);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c * 8 + e >= Lk) w[e >> 1] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
    }
  }
  __syncthreads();
  griddep_launch_dependents();

  // ---- this warp's 16 query rows: A fragments of Q for the two 16-wide k steps of the head dimension
  const int g = lane >> 2, t4 = lane & 3;                    // fragment row / column-pair index
  uint32_t qa[2][4];
  {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ldmatrix_x4(smem_u32(q_s + row * ROW_PITCH + (ks * 2 + (lane >> 4)) * 16), qa[ks]);
  }
  float o[4][4];                                             // O accumulators: 4 blocks of 8 head dimensions
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};                   // rows g, g + 8

  const uint32_t k_lane = smem_u32(k_s + (lane & 7) * ROW_PITCH + (lane >> 3) * 16);
  const uint32_t v_lane = smem_u32(v_s + ((lane & 7) + (lane >> 4) * 8) * vpb + ((lane >> 3) & 1) * 16);
  float2 l2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};   // row sums as (even, odd) column halves: packed adds
  // one 64-key chunk; kRagged (the last chunk when Lk % 64 != 0) masks the keys >= Lk out of the softmax -- a separate
  // instantiation, so that full chunks do not pay 50 predicated selects each
  auto chunk = [&](int kc, auto ragged_tag) {
    constexpr bool kRagged = decltype(ragged_tag)::value;
    // S = Q K^T for 64 keys: 8 blocks of 8 keys
    float s[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
      uint32_t kb[4];
      ldmatrix_x4(k_lane + (kc + nb * 8) * ROW_PITCH, kb);   // keys kc + 8 nb .. + 7, head dimensions 0-7 | 8-15 | 16-23 | 24-31
      mma_bf16_16816(s[nb], qa[0], kb[0], kb[1]);
      mma_bf16_16816(s[nb], qa[1], kb[2], kb[3]);
    }
    if (kRagged) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int key = kc + nb * 8 + t4 * 2;
        if (key >= Lk) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
        if (key + 1 >= Lk) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
      }
    }
    // online softmax: rows g (values 0, 1 of every block) and g + 8 (values 2, 3)
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float alpha[2];
    float2 neg_m[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_use = mx[r] == -INFINITY ? 0.f : mx[r];  // (a row without any key yet: keep everything zero)
      alpha[r] = ex2f(m_run[r] - m_use);                     // first chunk: 2^(-inf) = 0
      m_run[r] = mx[r];
      l2[r].x *= alpha[r]; l2[r].y *= alpha[r];
      neg_m[r] = make_float2(-m_use, -m_use);
    }
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) { o[dn][0] *= alpha[0]; o[dn][1] *= alpha[0]; o[dn][2] *= alpha[1]; o[dn][3] *= alpha[1]; }
    uint32_t pa[4][4];                                       // P as A fragments: key step j covers key blocks 2j, 2j + 1
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float2 x0 = __fadd2_rn(make_float2(s[nb][0], s[nb][1]), neg_m[0]);
      const float2 x1 = __fadd2_rn(make_float2(s[nb][2], s[nb][3]), neg_m[1]);
      const float2 p0 = make_float2(ex2f(x0.x), ex2f(x0.y));
      const float2 p1 = make_float2(ex2f(x1.x), ex2f(x1.y));
      l2[0] = __fadd2_rn(l2[0], p0);
      l2[1] = __fadd2_rn(l2[1], p1);
      pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0.x, p0.y);
      pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p1.x, p1.y);
    }
    // O += P V: per 16-key step, B fragments from V^T rows (head dimension = n, keys = k)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {                       // two blocks of 8 head dimensions per ldmatrix.x4
        uint32_t vb[4];
        ldmatrix_x4(v_lane + dp * 16 * vpb + (kc + j * 16) * 2, vb);
        mma_bf16_16816(o[dp * 2], pa[j], vb[0], vb[1]);
        mma_bf16_16816(o[dp * 2 + 1], pa[j], vb[2], vb[3]);
      }
    }
  };
  const int full_end = Lk / KC * KC;
  for (int kc = 0; kc < full_end; kc += KC) chunk(kc, std::false_type{});
  if (full_end < lk_pad) chunk(full_end, std::true_type{});
  const float l_run[2] = {l2[0].x + l2[0].y, l2[1].x + l2[1].y};

  // ---- O / l -> bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.0f / l;
    const int row = q0 + warp * 16 + g + r * 8;
    if (row < Lq) {
      __nv_bfloat16* op = out + (static_cast<size_t>(b) * Lq + row) * ldo + h * DH + t4 * 2;
#pragma unroll
      for (int dn = 0; dn < 4; ++dn)
        *reinterpret_cast<uint32_t*>(op + dn * 8) = pack_bf16x2(o[dn][r * 2] * inv, o[dn][r * 2 + 1] * inv);
    }
  }
}

template <int kWarps>
static int launch_small_t(const AttnArgs& a, int lk_pad, cudaStream_t stream) {
  using namespace attn_small;
  constexpr int QB = kWarps * 16;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_small_kernel<kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes(QB, MAX_LK)));
    if (e != cudaSuccess) return svol_fail_cuda(e, "attention (short keys): cudaFuncSetAttribute");
    configured = true;
  }
  const dim3 grid((a.Lq + QB - 1) / QB, a.H, a.B);
  cudaError_t e = launch_kernel_pdl(attention_small_kernel<kWarps>, grid, dim3(kWarps * 32), smem_bytes(QB, lk_pad), stream,
                                    reinterpret_cast<const __nv_bfloat16*>(a.q), reinterpret_cast<const __nv_bfloat16*>(a.k),
                                    reinterpret_cast<const __nv_bfloat16*>(a.vt), reinterpret_cast<__nv_bfloat16*>(a.out), a.H, a.Lq, a.Lk,
                                    a.ldq, a.ldk, a.vt_pitch, a.ldo, lk_pad);
  if (e != cudaSuccess) return svol_fail_cuda(e, "attention (short keys) launch");
  return svol_check_launch("attention_small");
}

// Returns -1 when the launch does not qualify (the caller falls through to attention_tc_kernel).
int launch_attention_small(const AttnArgs& a, cudaStream_t stream) {
  using namespace attn_small;
  const char* env = getenv("SVOL_ATTN_SMALL");               // read per launch: tests and A/B measurements flip it
  const bool enabled = env == nullptr || env[0] != '0';
  if (!enabled || a.lse != nullptr || a.key_mask != nullptr || a.Lk > MAX_LK) return -1;
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.vt)) & 15) return -1;
  if ((reinterpret_cast<uintptr_t>(a.out) & 3) || (a.ldo & 1)) return -1;
  const int lk_pad = (a.Lk + KC - 1) / KC * KC;
  // Query rows per CTA (16 per warp): the CTAs of a launch run in rounds of (SMs x CTAs resident per SM), every round costs a
  // load phase plus a softmax phase that grows with the rows per SM -- pick the block size with the cheapest schedule
  // (SVOL_ATTN_SMALL_WARPS=4|5|8 forces one).  320 queries x 8 heads x 32 samples on 148 SMs: 5 warps = 1024 CTAs in 2 rounds,
  // 4 warps = 1280 CTAs in 3.
  const char* env_w = getenv("SVOL_ATTN_SMALL_WARPS");
  int best_w = env_w ? atoi(env_w) : 0;
  if (best_w != 4 && best_w != 5 && best_w != 8) {
    double best = 1e30;
    for (int w : {4, 5, 8}) {
      const int qb = 16 * w;
      const long ctas = static_cast<long>((a.Lq + qb - 1) / qb) * a.H * a.B;
      const int per_sm = static_cast<int>((227 * 1024) / (smem_bytes(qb, lk_pad) + 1024));
      if (per_sm < 1) continue;
      const long slots = static_cast<long>(sm_count()) * per_sm;
      const long rounds = (ctas + slots - 1) / slots;
      // per round: key / value copy (~3 k clk) + MUFU time of the rows resident on an SM (rows x keys / 16 per clk)
      const double cost = rounds * (3000.0 + static_cast<double>(per_sm) * qb * lk_pad / 16.0);
      if (cost < best) { best = cost; best_w = w; }
    }
    if (best_w == 0) return -1;
  }
  switch (best_w) {
    case 4: return launch_small_t<4>(a, lk_pad, stream);
    case 5: return launch_small_t<5>(a, lk_pad, stream);
    default: return launch_small_t<8>(a, lk_pad, stream);
  }
}

}  // namespace svol
