// Flash-style multi-head attention core on tcgen05 / TMEM / TMA (sm_100a), head_dim = 32.
//
// Replaces softmax(Q K^T / sqrt(dh) + key_padding_mask) V inside nn.MultiheadAttention at
// lib/modeling/cross_modal_transformer.py:139 (video self-attention, L x L), :147 (query
// self-attention, Q x Q) and :154 (query -> video cross-attention, Q x L, padded keys masked).
// The reference materialises the (B*8, Lq, Lk) score tensor (2.5 GB at the headline config); here
// scores only ever exist as one 128 x 128 fp32 tile in tensor memory.
//
// One CTA per (128-query tile, head, sample); 192 threads; two CTAs are co-resident per SM so one
// CTA's softmax overlaps the other's MMAs:
//   warps 0..3  softmax: thread = query row.  S tile TMEM -> registers (two passes: row max, then
//               exp2), P tile -> shared memory as bf16 in the 128B-swizzled K-major layout the MMA
//               reads, running max / sum / 32-wide output accumulator in registers.
//   warp 4      TMA producer: Q tile once, then a 3-stage ring of K tiles (128 x 32, 64B swizzle)
//               and V^T tiles (32 x 128, 128B swizzle).
//   warp 5      MMA issuer: S = Q K^T (M128 N128 K32) and O_tile = P V (M128 N32 K128), fp32 in TMEM.
// Q is pre-scaled by log2(e)/sqrt(dh) when it is produced, so the softmax is a bare ex2.
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace attn {
constexpr int BQ = 128, BKV = 128, DH = 32, STAGES = 3;
constexpr int Q_BYTES = BQ * DH * 2;            // 8192
constexpr int K_BYTES = BKV * DH * 2;           // 8192
constexpr int VT_BYTES = DH * BKV * 2;          // 8192 = 2 k-blocks of 32 rows x 128 B
constexpr int P_BYTES = BQ * BKV * 2;           // 32768 = 2 k-blocks of 128 rows x 128 B
constexpr int OFF_K = Q_BYTES, OFF_VT = OFF_K + STAGES * K_BYTES, OFF_P = OFF_VT + STAGES * VT_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 256, TMEM_O = 128;
}  // namespace attn

struct AttnBars {
  uint64_t q_full, s_full, s_free, p_ready, o_full;
  uint64_t kv_full[attn::STAGES], kv_empty[attn::STAGES];
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(attn::THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, const float* __restrict__ key_mask,
                    __nv_bfloat16* __restrict__ out, int H, int Lq, int Lk, int ldo) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + OFF_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (Lk + BKV - 1) / BKV;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_free, 4);
    mbar_init(&bars->p_ready, 4);
    mbar_init(&bars->o_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->q_full, Q_BYTES);
      tma_load_2d(smem, &tmQ, &bars->q_full, h * DH, b * Lq + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % STAGES;
        const uint32_t use = j / STAGES;
        mbar_wait(&bars->kv_empty[s], (use & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->kv_full[s], K_BYTES + VT_BYTES);
        tma_load_2d(smem + OFF_K + s * K_BYTES, &tmK, &bars->kv_full[s], h * DH, b * Lk + j * BKV);
        const int vrow = (b * H + h) * DH;
        tma_load_2d(smem + OFF_VT + s * VT_BYTES, &tmVt, &bars->kv_full[s], j * BKV, vrow);
        tma_load_2d(smem + OFF_VT + s * VT_BYTES + VT_BYTES / 2, &tmVt, &bars->kv_full[s], j * BKV + 64, vrow);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, DH);
      const uint32_t sQ = smem_u32(smem), sP = smem_u32(smem + OFF_P);
      auto issue_pv = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(&bars->p_ready, j & 1);
        tcgen05_fence_after();
        const uint32_t sV = smem_u32(smem + OFF_VT + s * VT_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t a_desc = make_kmajor_desc<128>(sP + (k >> 2) * (P_BYTES / 2) + (k & 3) * 32);
          const uint64_t b_desc = make_kmajor_desc<128>(sV + (k >> 2) * (VT_BYTES / 2) + (k & 3) * 32);
          umma_bf16_ss(tmem_base + TMEM_O, a_desc, b_desc, idesc_o, k != 0);
        }
        umma_commit(&bars->o_full);
        umma_commit(&bars->kv_empty[s]);
      };
      mbar_wait(&bars->q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % STAGES;
        mbar_wait(&bars->kv_full[s], (j / STAGES) & 1);
        if (j > 0) mbar_wait(&bars->s_free, (j - 1) & 1);
        tcgen05_fence_after();
        const uint32_t sK = smem_u32(smem + OFF_K + s * K_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss(tmem_base, make_kmajor_desc<64>(sQ + k * 32), make_kmajor_desc<64>(sK + k * 32), idesc_s, k != 0);
        umma_commit(&bars->s_full);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(n_tiles - 1);
    }
  } else {
    // ------------------------------------------------------------------ softmax (4 warps)
    const int r = warp * 32 + lane;                    // row inside the tile == TMEM lane
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* p_row = smem + OFF_P + (r >> 3) * 1024 + (r & 7) * 128;
    const float* mrow = key_mask ? key_mask + static_cast<size_t>(b) * Lk : nullptr;
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 0.f;
    float acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = j * BKV;
      const bool masked_tile = (mrow != nullptr) || (kv0 + BKV > Lk);
      mbar_wait(&bars->s_full, j & 1);
      tcgen05_fence_after();
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c * 32, v);
        tmem_ld_wait();
        if (masked_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kv = kv0 + c * 32 + i;
            const bool ok = kv < Lk && (mrow == nullptr || __ldg(mrow + kv) != 0.f);
            mx = fmaxf(mx, ok ? __uint_as_float(v[i]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = ex2_approx(m_run - m_use);
      // pass 2: p = 2^(s - m), packed to bf16
      uint32_t p[BKV / 2];
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0, e1;
          if (masked_tile) {
            const int kv = kv0 + c * 32 + i;
            const bool ok0 = kv < Lk && (mrow == nullptr || __ldg(mrow + kv) != 0.f);
            const bool ok1 = kv + 1 < Lk && (mrow == nullptr || __ldg(mrow + kv + 1) != 0.f);
            e0 = ok0 ? ex2_approx(__uint_as_float(v[i]) - m_use) : 0.f;
            e1 = ok1 ? ex2_approx(__uint_as_float(v[i + 1]) - m_use) : 0.f;
          } else {
            e0 = ex2_approx(__uint_as_float(v[i]) - m_use);
            e1 = ex2_approx(__uint_as_float(v[i + 1]) - m_use);
          }
          rs += e0 + e1;
          p[c * 16 + (i >> 1)] = pack_bf16x2(e0, e1);
        }
      }
      // S has been consumed: the MMA warp may overwrite it with the next tile's scores
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free);
      l_run = l_run * alpha + rs;
      m_run = m_new;

      if (j > 0) {
        // fold in the previous tile's P V (this also guarantees the MMA is done reading P)
        mbar_wait(&bars->o_full, (j - 1) & 1);
        tcgen05_fence_after();
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_row + TMEM_O, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < DH; ++i) acc[i] = acc[i] * alpha_pending + __uint_as_float(o[i]);
      }
      alpha_pending = alpha;
      // P tile -> shared memory (K-major, 128B swizzle: 16B chunk index XOR (row & 7))
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        const int kb = ch >> 3, jj = ch & 7;
        uint4 q = make_uint4(p[ch * 4], p[ch * 4 + 1], p[ch * 4 + 2], p[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(p_row + kb * (P_BYTES / 2) + ((jj ^ (r & 7)) << 4)) = q;
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_ready);
    }
    // last tile
    mbar_wait(&bars->o_full, (n_tiles - 1) & 1);
    tcgen05_fence_after();
    {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_row + TMEM_O, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < DH; ++i) acc[i] = acc[i] * alpha_pending + __uint_as_float(o[i]);
    }
    const int q = q0 + r;
    if (q < Lq) {
      const float inv = 1.0f / l_run;
      uint4* op = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * Lq + q) * ldo + h * DH);
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        uint4 w;
        w.x = pack_bf16x2(acc[8 * i + 0] * inv, acc[8 * i + 1] * inv);
        w.y = pack_bf16x2(acc[8 * i + 2] * inv, acc[8 * i + 3] * inv);
        w.z = pack_bf16x2(acc[8 * i + 4] * inv, acc[8 * i + 5] * inv);
        w.w = pack_bf16x2(acc[8 * i + 6] * inv, acc[8 * i + 7] * inv);
        op[i] = w;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc<attn::TMEM_COLS>(tmem_base);
  }
}

int launch_attention_tc(const AttnArgs& a, cudaStream_t stream) {
  using namespace attn;
  if (a.B <= 0 || a.H <= 0 || a.Lq <= 0 || a.Lk <= 0) return svol_fail(SVOL_ERR_SHAPE, "attention: bad sizes");
  if (a.vt_pitch < a.Lk || a.vt_pitch % 8 != 0 || a.ldq % 8 || a.ldk % 8 || a.ldo % 8)
    return svol_fail(SVOL_ERR_SHAPE, "attention: pitches must be multiples of 8 elements and vt_pitch >= Lk");
  CUtensorMap tmQ, tmK, tmVt;
  int rc = make_tensor_map_2d(&tmQ, a.q, a.H * DH, static_cast<int64_t>(a.B) * a.Lq, a.ldq, DH, BQ, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmK, a.k, a.H * DH, static_cast<int64_t>(a.B) * a.Lk, a.ldk, DH, BKV, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmVt, a.vt, a.vt_pitch, static_cast<int64_t>(a.B) * a.H * DH, a.vt_pitch, 64, DH, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return svol_fail_cuda(e, "attention: cudaFuncSetAttribute");
    configured = true;
  }
  dim3 grid((a.Lq + BQ - 1) / BQ, a.H, a.B);
  attention_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmVt, a.key_mask,
                                                            reinterpret_cast<__nv_bfloat16*>(a.out), a.H, a.Lq, a.Lk, a.ldo);
  return svol_check_launch("attention_tc");
}

}  // namespace svol
