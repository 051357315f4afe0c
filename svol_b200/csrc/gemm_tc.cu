// Persistent warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   out[M,N] = epilogue( A[M,K] (bf16, row-major) x W[N,K]^T (bf16, nn.Linear layout) )
//
// replaces the cuBLAS GEMM + separate bias / activation / residual / LayerNorm kernels that the
// reference's nn.Linear / nn.LayerNorm / F.gelu / F.relu calls turn into
// (lib/modeling/svanet.py:159-181, lib/modeling/cross_modal_transformer.py:88-100,137-158,163-179).
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer : A tile 128x64 and W tile 256x64 per k-block, 4-stage smem ring
//   warp 1      MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, fp32
//                              accumulators in TMEM, two accumulator stages (2 x 256 columns)
//   warps 2..9  epilogue     : tcgen05.ld -> registers (TMEM stage released immediately) ->
//                              +bias -> ReLU | GELU(erf) -> +residual -> LayerNorm over the
//                              256-wide row -> bf16 store, optional second output (x + pos),
//                              optional per-head transposed store for the attention V operand.
// The epilogue of tile i overlaps the mainloop of tile i+1.
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace gemm {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8, THREADS = 64 + EPI_WARPS * 32;
constexpr int COLS_PER_THREAD = BN / (EPI_WARPS / 4);   // 128
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * BM * 4 * 2 /*LN exchange*/;
}  // namespace gemm

struct GemmSmemTail {
  uint64_t full[gemm::STAGES];
  uint64_t empty[gemm::STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(gemm::THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmEpilogue ep, int M, int N, int K) {
  using namespace gemm;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmSmemTail* tail = reinterpret_cast<GemmSmemTail*>(smem + STAGES * STAGE_BYTES);
  float* ln_x = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // [2 halves][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blocks = (M + BM - 1) / BM;
  const int n_blocks = N / BN;
  const int num_tiles = m_blocks * n_blocks;
  const int num_kb = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tail->tmem_full[s], 1); mbar_init(&tail->tmem_empty[s], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tail->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&tail->full[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d(sa, &tmA, &tail->full[stage], kb * BK, m_blk * BM);
          tma_load_2d(sa + A_BYTES, &tmB, &tail->full[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&tail->full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_desc = make_kmajor_desc<128>(sa);
          const uint64_t b_desc = make_kmajor_desc<128>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&tail->empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tail->tmem_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int quarter = warp & 3;               // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;           // which 128 of the tile's 256 columns
    const int r_in_tile = quarter * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_blk * BM + r_in_tile;
      const int col0 = n_blk * BN + half * COLS_PER_THREAD;
      const bool row_ok = row < M;

      mbar_wait(&tail->tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      float v[COLS_PER_THREAD];
      {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * COLS_PER_THREAD;
#pragma unroll
        for (int c = 0; c < COLS_PER_THREAD / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c * 32 + i] = __uint_as_float(r[i]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);

      // bias
      if (ep.bias) {
        const float4* bp = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD / 4; ++i) {
          const float4 b = __ldg(bp + i);
          v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
      }
      if (ep.act == SVOL_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (ep.act == SVOL_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) v[i] = gelu_erf(v[i]);
      }
      if (ep.residual && row_ok) {
        const uint4* rp = reinterpret_cast<const uint4*>(ep.residual + static_cast<size_t>(row) * ep.ld_res + col0);
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD / 8; ++i) {
          const uint4 q = __ldg(rp + i);
          v[8 * i + 0] += bf16_lo(q.x); v[8 * i + 1] += bf16_hi(q.x);
          v[8 * i + 2] += bf16_lo(q.y); v[8 * i + 3] += bf16_hi(q.y);
          v[8 * i + 4] += bf16_lo(q.z); v[8 * i + 5] += bf16_hi(q.z);
          v[8 * i + 6] += bf16_lo(q.w); v[8 * i + 7] += bf16_hi(q.w);
        }
      }
      if (ep.ln_weight) {
        // LayerNorm over the full 256-wide row: the two warps that share a row exchange partial
        // sums through shared memory (two-pass: mean, then centred second moment).
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) s += v[i];
        ln_x[half * BM + r_in_tile] = s;
        named_bar_sync(1 + quarter, 64);
        const float mean = (ln_x[r_in_tile] + ln_x[BM + r_in_tile]) * (1.0f / BN);
        named_bar_sync(1 + quarter, 64);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) { const float d = v[i] - mean; ss += d * d; }
        ln_x[half * BM + r_in_tile] = ss;
        named_bar_sync(1 + quarter, 64);
        const float var = (ln_x[r_in_tile] + ln_x[BM + r_in_tile]) * (1.0f / BN);
        named_bar_sync(1 + quarter, 64);
        const float rstd = rsqrtf(var + ep.ln_eps);
        const float4* gp = reinterpret_cast<const float4*>(ep.ln_weight + col0);
        const float4* bp = reinterpret_cast<const float4*>(ep.ln_bias + col0);
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD / 4; ++i) {
          const float4 g = __ldg(gp + i), b = __ldg(bp + i);
          v[4 * i + 0] = (v[4 * i + 0] - mean) * rstd * g.x + b.x;
          v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * g.y + b.y;
          v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * g.z + b.z;
          v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * g.w + b.w;
        }
      }
      if (row_ok && ep.out) {
        uint4* op = reinterpret_cast<uint4*>(ep.out + static_cast<size_t>(row) * ep.ld_out + col0);
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD / 8; ++i) {
          uint4 q;
          q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          op[i] = q;
        }
      }
      if (row_ok && ep.out_pos) {
        // second output: x + pos (the q/k operand of the next attention block)
        const int prow = ep.pos_row_mod > 0 ? row % ep.pos_row_mod : row;
        const float4* pp = reinterpret_cast<const float4*>(ep.pos + static_cast<size_t>(prow) * ep.ld_pos + col0);
        uint4* op = reinterpret_cast<uint4*>(ep.out_pos + static_cast<size_t>(row) * ep.ld_out + col0);
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD / 8; ++i) {
          const float4 p0 = __ldg(pp + 2 * i), p1 = __ldg(pp + 2 * i + 1);
          uint4 q;
          q.x = pack_bf16x2(v[8 * i + 0] + p0.x, v[8 * i + 1] + p0.y);
          q.y = pack_bf16x2(v[8 * i + 2] + p0.z, v[8 * i + 3] + p0.w);
          q.z = pack_bf16x2(v[8 * i + 4] + p1.x, v[8 * i + 5] + p1.y);
          q.w = pack_bf16x2(v[8 * i + 6] + p1.z, v[8 * i + 7] + p1.w);
          op[i] = q;
        }
      }
      if (row_ok && ep.out_vt) {
        // per-head transposed store: Vt[(b*H + h)*dh + d][l], row = b*L + l, col = h*dh + d
        const int b = row / ep.vt_len, l = row - b * ep.vt_len;
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ep.out_vt) + (static_cast<size_t>(b) * N + col0) * ep.vt_pitch + l;
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) base[static_cast<size_t>(i) * ep.vt_pitch] = __float2bfloat16_rn(v[i]);
      }
      __syncwarp();   // reconverge before the next tile's warp-collective tcgen05.ld
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int launch_gemm_bf16_tc(const GemmArgs& a, cudaStream_t stream) {
  using namespace gemm;
  if (a.N % BN != 0 || a.K % BK != 0 || a.M <= 0) return svol_fail(SVOL_ERR_SHAPE, "gemm: need N % 256 == 0, K % 64 == 0, M > 0");
  if (a.ep.ln_weight && a.N != BN) return svol_fail(SVOL_ERR_SHAPE, "gemm: fused LayerNorm needs N == 256");
  if (a.ep.out_vt && a.N != BN) return svol_fail(SVOL_ERR_SHAPE, "gemm: transposed-V store needs N == 256");
  CUtensorMap tmA, tmB;
  int rc = make_tensor_map_2d(&tmA, a.A, a.K, a.M, a.lda, BK, BM, 128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmB, a.W, a.K, a.N, a.ldw, BK, BN, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return svol_fail_cuda(e, "gemm: cudaFuncSetAttribute");
    configured = true;
  }
  const int m_blocks = (a.M + BM - 1) / BM;
  const int tiles = m_blocks * (a.N / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_bf16_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, a.ep, a.M, a.N, a.K);
  return svol_check_launch("gemm_bf16_tc");
}

}  // namespace svol
