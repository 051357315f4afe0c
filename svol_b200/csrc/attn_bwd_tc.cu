// Backward of the multi-head attention core on tcgen05 / TMEM / TMA (sm_100a), head_dim = 32.
//
// The reference differentiates softmax(Q K^T / sqrt(dh) + mask) V with torch.autograd through
// nn.MultiheadAttention (lib/modeling/cross_modal_transformer.py:139,147,154; train.py:229 loss.backward()), which
// re-reads the materialised (B*8, Lq, Lk) probability tensor.  Here, as in the forward (attn_tc.cu), scores and
// probabilities only exist as tiles in tensor memory: they are recomputed from Q, K and the saved base-2
// log-sum-exp of every row,
//     P = 2^(S - lse),   dP = dO V^T,   dS = P * (dP - delta),   delta = rowsum(dO * O),
//     dQ = dS K / sqrt(dh),   dK = ln2 * dS^T Q',   dV = P^T dO          (Q' = Q * log2(e)/sqrt(dh) as stored)
// Two kernels, so that every accumulator stays in tensor memory of ONE CTA and no gradient needs atomics:
//   attn_bwd_dq_kernel    one CTA per (128-query tile, head, sample), loops over 64-key blocks:
//                         S = Q K_j^T, dP = dO V_j^T (operands in shared memory) -> dS (bf16) to TENSOR MEMORY ->
//                         dQ += dS K_j  (A = dS in TMEM, B = the K^T block in shared memory)
//   attn_bwd_dkdv_kernel  one CTA per (128-key tile, head, sample), loops over 64-query blocks, with the roles of
//                         rows and columns swapped: S^T = K Q_j^T, dP^T = V dO_j^T -> P^T, dS^T (bf16) to TMEM ->
//                         dV += P^T dO_j, dK += dS^T Q_j  (B = the transposed dO / Q blocks in shared memory)
// Every operand is K-major, exactly the layouts the forward uses: Q, K, V, dO row-major [B*L, ld] with head h at
// columns [32h, 32h+32) (TMA boxes of 32 columns, 64B swizzle) and per-head transposed copies K^T, Q^T, dO^T
// [B*8*32, pitch] (boxes of 64 columns x 32 rows, 128B swizzle) that the projection GEMMs' epilogues write
// (gemm_tc.cu: out_vt).  Each CTA uses 256 TMEM columns and ~80 KB of shared memory, so TWO CTAs share an SM: while
// one CTA's 128 softmax threads turn scores into gradients, the other CTA's MMAs run -- the overlap the forward
// builds by hand with four staggered warpgroups comes from occupancy here.
//   warps 0..7  compute: warp w works on TMEM lane quarter w % 4 (thread = one row of the tile) and on columns
//               [32 (w / 4), +32) of every 64-column block, so 16 compute warps per SM keep the MUFU pipe fed
//   warp 8      TMA producer          warp 9  MMA issuer
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace abwd {
constexpr int BM = 128, BN = 64, DH = 32;
constexpr int COMPUTE_WARPS = 8;                 // two warps per TMEM lane quarter, 32 of the 64 block columns each
constexpr int THREADS = (COMPUTE_WARPS + 2) * 32;
constexpr int ROW_TILE_BYTES = BM * DH * 2;      // 8192: [128 x 32] bf16, 64B swizzle
constexpr int COL_TILE_BYTES = BN * DH * 2;      // 4096: [64 x 32] bf16 (64B swizzle) or its transpose [32 x 64] (128B swizzle)
constexpr uint32_t TMEM_COLS = 256;
// dQ kernel
constexpr int DQ_STAGES = 5, DQ_STAGE_BYTES = 3 * COL_TILE_BYTES;                   // K_j, V_j, K_j^T
constexpr int DQ_OFF_STAGES = 2 * ROW_TILE_BYTES;
constexpr int DQ_OFF_BAR = DQ_OFF_STAGES + DQ_STAGES * DQ_STAGE_BYTES;
constexpr int DQ_SMEM = DQ_OFF_BAR + 256 + 1024;
constexpr uint32_t DQ_T_S = 0, DQ_T_DP = 64, DQ_T_DS = 128, DQ_T_ACC = 160;
// dK / dV kernel
constexpr int KV_STAGES = 4, KV_STAGE_BYTES = 4 * COL_TILE_BYTES;                   // Q_j, dO_j, Q_j^T, dO_j^T
constexpr int KV_OFF_STAGES = 2 * ROW_TILE_BYTES;
constexpr int KV_OFF_STAT = KV_OFF_STAGES + KV_STAGES * KV_STAGE_BYTES;             // [2 buffers][lse 64 | delta 64] floats
constexpr int KV_OFF_BAR = KV_OFF_STAT + 2 * 2 * BN * 4;
constexpr int KV_SMEM = KV_OFF_BAR + 256 + 1024;
constexpr uint32_t KV_T_S = 0, KV_T_DP = 64, KV_T_P = 128, KV_T_DS = 160, KV_T_DV = 192, KV_T_DK = 224;
static_assert(DQ_SMEM > 232448 / 3 && KV_SMEM > 232448 / 3, "at most two CTAs per SM (2 x 256 TMEM columns)");
constexpr int MAX_STAGES = 5;
}  // namespace abwd

struct AttnBwdBars {
  uint64_t once_full;                       // the CTA's own row tiles
  uint64_t full[abwd::MAX_STAGES], empty[abwd::MAX_STAGES];
  uint64_t sdp_full[2], ds_ready[2], acc_full;   // per 32-column half of the block: two independent MMA <-> compute chains
  uint32_t tmem_base, pad;
};
static_assert(sizeof(AttnBwdBars) <= 256, "barrier block");

namespace {
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// validity of the 32 keys [kv0, kv0 + 32) of sample b as a ballot word (ragged tail and key_padding_mask)
__device__ __forceinline__ uint32_t key_word(const float* mrow, int kv0, int Lk, int lane) {
  const int kv = kv0 + lane;
  return __ballot_sync(0xffffffffu, kv < Lk && (mrow == nullptr || __ldg(mrow + kv) != 0.f));
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(abwd::THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmKt, const float* __restrict__ lse, const float* __restrict__ delta,
                   const float* __restrict__ key_mask, __nv_bfloat16* __restrict__ dq, int H, int Lq, int Lk, int stat_pitch,
                   int ld_dq) {
  using namespace abwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBwdBars* bars = reinterpret_cast<AttnBwdBars*>(smem + DQ_OFF_BAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blk = (Lk + BN - 1) / BN;

  if (warp == COMPUTE_WARPS && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmKt);
    mbar_init(&bars->once_full, 1);
    for (int s = 0; s < DQ_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int hf = 0; hf < 2; ++hf) { mbar_init(&bars->sdp_full[hf], 1); mbar_init(&bars->ds_ready[hf], COMPUTE_WARPS / 2); }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == COMPUTE_WARPS + 1) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == COMPUTE_WARPS) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->once_full, 2 * ROW_TILE_BYTES);
      tma_load_2d(smem, &tmQ, &bars->once_full, h * DH, b * Lq + q0);
      tma_load_2d(smem + ROW_TILE_BYTES, &tmdO, &bars->once_full, h * DH, b * Lq + q0);
      for (int j = 0; j < n_blk; ++j) {
        const int s = j % DQ_STAGES;
        mbar_wait(&bars->empty[s], ((j / DQ_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[s], DQ_STAGE_BYTES);
        uint8_t* st = smem + DQ_OFF_STAGES + s * DQ_STAGE_BYTES;
        tma_load_2d(st, &tmK, &bars->full[s], h * DH, b * Lk + j * BN);
        tma_load_2d(st + COL_TILE_BYTES, &tmV, &bars->full[s], h * DH, b * Lk + j * BN);
        tma_load_2d(st + 2 * COL_TILE_BYTES, &tmKt, &bars->full[s], j * BN, (b * H + h) * DH);
      }
    }
  } else if (warp == COMPUTE_WARPS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_acc = make_idesc_bf16(BM, DH);
      const uint64_t dQd = make_kmajor_desc<64>(smem_u32(smem));
      const uint64_t dOd = make_kmajor_desc<64>(smem_u32(smem + ROW_TILE_BYTES));
      // Each 64-key block is processed as two 32-key halves with their own barriers: while the compute warps of one half
      // turn scores into dS, the MMAs of the other half run -- two MMA <-> compute chains per CTA, four per SM.
      auto issue_acc = [&](int j, int hf) {         // dQ += dS_hf(j) K_j[32 hf : 32 hf + 32]
        const int s = j % DQ_STAGES;
        mbar_wait(&bars->ds_ready[hf], j & 1);
        tcgen05_fence_after();
        const uint64_t dKt = make_kmajor_desc<128>(smem_u32(smem + DQ_OFF_STAGES + s * DQ_STAGE_BYTES + 2 * COL_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int kk = hf * 2 + k;
          umma_ts(tmem_base + DQ_T_ACC, tmem_base + DQ_T_DS + kk * 8, dKt + 2 * kk, idesc_acc, (j > 0 || kk > 0) ? 1u : 0u);
        }
        if (hf == 1) umma_commit(&bars->empty[s]);
      };
      mbar_wait(&bars->once_full, 0);
      for (int j = 0; j < n_blk; ++j) {
        const int s = j % DQ_STAGES;
        mbar_wait(&bars->full[s], (j / DQ_STAGES) & 1);
        const uint8_t* st = smem + DQ_OFF_STAGES + s * DQ_STAGE_BYTES;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (j > 0) issue_acc(j - 1, hf);
          tcgen05_fence_after();
          // rows [32 hf, +32) of the 64B-swizzled [64 x 32] K / V tiles start 2048 bytes in
          const uint64_t dK = make_kmajor_desc<64>(smem_u32(st) + hf * 2048);
          const uint64_t dV = make_kmajor_desc<64>(smem_u32(st + COL_TILE_BYTES) + hf * 2048);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + DQ_T_S + hf * 32, dQd + 2 * k, dK + 2 * k, idesc_acc, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + DQ_T_DP + hf * 32, dOd + 2 * k, dV + 2 * k, idesc_acc, k != 0);
          umma_commit(&bars->sdp_full[hf]);
        }
      }
      issue_acc(n_blk - 1, 0);
      issue_acc(n_blk - 1, 1);
      umma_commit(&bars->acc_full);
    }
  } else {
    // ------------------------------------------------------------------ compute: thread = query row
    const int quarter = warp & 3, c = warp >> 2;      // TMEM lane quarter, column half of each 64-key block
    const int r = quarter * 32 + lane;
    const int q = q0 + r;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const size_t stat = (static_cast<size_t>(b) * H + h) * stat_pitch;
    const float lse_r = q < Lq ? __ldg(lse + stat + q) : INFINITY;
    const float delta_r = q < Lq ? __ldg(delta + stat + q) : 0.f;
    const float* mrow = key_mask ? key_mask + static_cast<size_t>(b) * Lk : nullptr;
    for (int j = 0; j < n_blk; ++j) {
      uint32_t word = 0xffffffffu;                  // validity of this warp's 32 keys of the block
      if (mrow != nullptr || (j + 1) * BN > Lk) word = key_word(mrow, j * BN + c * 32, Lk, lane);
      mbar_wait(&bars->sdp_full[c], j & 1);
      tcgen05_fence_after();
      {
        uint32_t s[32], dp[32], packed[16];
        tmem_ld_32x32b_x32(t_lane + DQ_T_S + c * 32, s);
        tmem_ld_32x32b_x32(t_lane + DQ_T_DP + c * 32, dp);
        tmem_ld_wait();
        // per pair of scores: one packed subtract (S - lse), two MUFU ex2, one packed subtract (dP - delta), one packed
        // multiply, one pack -- 3 issue slots per element; blocks without invalid keys skip the mask selects
        const float2 nl = make_float2(-lse_r, -lse_r), nd = make_float2(-delta_r, -delta_r);
        if (word == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = __fadd2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), nl);
            const float2 pr = make_float2(ex2f(x.x), ex2f(x.y));
            const float2 g = __fmul2_rn(pr, __fadd2_rn(make_float2(__uint_as_float(dp[i]), __uint_as_float(dp[i + 1])), nd));
            packed[i >> 1] = pack_bf16x2(g.x, g.y);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = __fadd2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), nl);
            float2 pr = make_float2(ex2f(x.x), ex2f(x.y));
            if (!((word >> i) & 1u)) pr.x = 0.f;
            if (!((word >> (i + 1)) & 1u)) pr.y = 0.f;
            const float2 g = __fmul2_rn(pr, __fadd2_rn(make_float2(__uint_as_float(dp[i]), __uint_as_float(dp[i + 1])), nd));
            packed[i >> 1] = pack_bf16x2(g.x, g.y);
          }
        }
        tmem_st_x16(t_lane + DQ_T_DS + c * 16, packed);
      }
      tmem_st_wait_();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ds_ready[c]);
    }
    mbar_wait(&bars->acc_full, 0);
    tcgen05_fence_after();
    uint32_t acc[16];                               // this warp's half of the 32 accumulator columns
    tmem_ld_32x32b_x16(t_lane + DQ_T_ACC + c * 16, acc);
    tmem_ld_wait();
    if (q < Lq) {
      const float sc = 0.17677669529663687f;       // 1 / sqrt(32)
      uint4* op = reinterpret_cast<uint4*>(dq + (static_cast<size_t>(b) * Lq + q) * ld_dq + h * DH + c * 16);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(acc[8 * i + 0]) * sc, __uint_as_float(acc[8 * i + 1]) * sc);
        w.y = pack_bf16x2(__uint_as_float(acc[8 * i + 2]) * sc, __uint_as_float(acc[8 * i + 3]) * sc);
        w.z = pack_bf16x2(__uint_as_float(acc[8 * i + 4]) * sc, __uint_as_float(acc[8 * i + 5]) * sc);
        w.w = pack_bf16x2(__uint_as_float(acc[8 * i + 6]) * sc, __uint_as_float(acc[8 * i + 7]) * sc);
        op[i] = w;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == abwd::COMPUTE_WARPS + 1) {
    tcgen05_fence_after();
    tmem_dealloc<abwd::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dK, dV
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(abwd::THREADS, 2)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmdOt,
                     const float* __restrict__ lse, const float* __restrict__ delta, const float* __restrict__ key_mask,
                     __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int H, int Lq, int Lk, int stat_pitch,
                     int ld_dk, int ld_dv) {
  using namespace abwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBwdBars* bars = reinterpret_cast<AttnBwdBars*>(smem + KV_OFF_BAR);
  float* stat_s = reinterpret_cast<float*>(smem + KV_OFF_STAT);       // [buf][lse 64 | delta 64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blk = (Lq + BN - 1) / BN;

  if (warp == COMPUTE_WARPS && lane == 0) {
    tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmdOt);
    mbar_init(&bars->once_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int hf = 0; hf < 2; ++hf) { mbar_init(&bars->sdp_full[hf], 1); mbar_init(&bars->ds_ready[hf], COMPUTE_WARPS / 2); }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == COMPUTE_WARPS + 1) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == COMPUTE_WARPS) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->once_full, 2 * ROW_TILE_BYTES);
      tma_load_2d(smem, &tmK, &bars->once_full, h * DH, b * Lk + k0);
      tma_load_2d(smem + ROW_TILE_BYTES, &tmV, &bars->once_full, h * DH, b * Lk + k0);
      for (int j = 0; j < n_blk; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(&bars->empty[s], ((j / KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[s], KV_STAGE_BYTES);
        uint8_t* st = smem + KV_OFF_STAGES + s * KV_STAGE_BYTES;
        tma_load_2d(st, &tmQ, &bars->full[s], h * DH, b * Lq + j * BN);
        tma_load_2d(st + COL_TILE_BYTES, &tmdO, &bars->full[s], h * DH, b * Lq + j * BN);
        tma_load_2d(st + 2 * COL_TILE_BYTES, &tmQt, &bars->full[s], j * BN, (b * H + h) * DH);
        tma_load_2d(st + 3 * COL_TILE_BYTES, &tmdOt, &bars->full[s], j * BN, (b * H + h) * DH);
      }
    }
  } else if (warp == COMPUTE_WARPS + 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_acc = make_idesc_bf16(BM, DH);
      const uint64_t dKd = make_kmajor_desc<64>(smem_u32(smem));
      const uint64_t dVd = make_kmajor_desc<64>(smem_u32(smem + ROW_TILE_BYTES));
      auto issue_acc = [&](int j, int hf) {         // dV += P^T_hf(j) dO_j[32 hf : +32] ;  dK += dS^T_hf(j) Q_j[32 hf : +32]
        const int s = j % KV_STAGES;
        mbar_wait(&bars->ds_ready[hf], j & 1);
        tcgen05_fence_after();
        const uint8_t* st = smem + KV_OFF_STAGES + s * KV_STAGE_BYTES;
        const uint64_t dQt = make_kmajor_desc<128>(smem_u32(st + 2 * COL_TILE_BYTES));
        const uint64_t dOt = make_kmajor_desc<128>(smem_u32(st + 3 * COL_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int kk = hf * 2 + k;
          umma_ts(tmem_base + KV_T_DV, tmem_base + KV_T_P + kk * 8, dOt + 2 * kk, idesc_acc, (j > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int kk = hf * 2 + k;
          umma_ts(tmem_base + KV_T_DK, tmem_base + KV_T_DS + kk * 8, dQt + 2 * kk, idesc_acc, (j > 0 || kk > 0) ? 1u : 0u);
        }
        if (hf == 1) umma_commit(&bars->empty[s]);
      };
      mbar_wait(&bars->once_full, 0);
      for (int j = 0; j < n_blk; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(&bars->full[s], (j / KV_STAGES) & 1);
        const uint8_t* st = smem + KV_OFF_STAGES + s * KV_STAGE_BYTES;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (j > 0) issue_acc(j - 1, hf);
          tcgen05_fence_after();
          const uint64_t dQ = make_kmajor_desc<64>(smem_u32(st) + hf * 2048);
          const uint64_t dO = make_kmajor_desc<64>(smem_u32(st + COL_TILE_BYTES) + hf * 2048);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + KV_T_S + hf * 32, dKd + 2 * k, dQ + 2 * k, idesc_acc, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + KV_T_DP + hf * 32, dVd + 2 * k, dO + 2 * k, idesc_acc, k != 0);
          umma_commit(&bars->sdp_full[hf]);
        }
      }
      issue_acc(n_blk - 1, 0);
      issue_acc(n_blk - 1, 1);
      umma_commit(&bars->acc_full);
    }
  } else {
    // ------------------------------------------------------------------ compute: thread = key row
    const int quarter = warp & 3, c = warp >> 2;      // TMEM lane quarter, column half of each 64-query block
    const int r = quarter * 32 + lane;
    const int key = k0 + r;
    const bool key_ok = key < Lk && (key_mask == nullptr || __ldg(key_mask + static_cast<size_t>(b) * Lk + key) != 0.f);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float2 kb2 = key_ok ? make_float2(0.f, 0.f) : make_float2(-INFINITY, -INFINITY);
    // Each half (4 warps) stages the lse | delta of ITS 32 queries of the block and synchronises on its own named
    // barrier, so the two halves never wait for each other.  hid = thread index inside the half: 0..31 load lse,
    // 32..63 load delta.  Layout of a buffer: [lse 64 | delta 64] floats, half c owns entries [32 c, 32 c + 32) of each.
    const int hid = quarter * 32 + lane;
    const float* stat_g = (hid < 32 ? lse : delta) + (static_cast<size_t>(b) * H + h) * stat_pitch + c * 32 + (hid & 31);
    float nxt = hid < 64 ? __ldg(stat_g) : 0.f;       // stat_pitch is a multiple of 64: every block read is in bounds
    for (int j = 0; j < n_blk; ++j) {
      float* st = stat_s + (j & 1) * 2 * BN;
      if (hid < 64) st[(hid < 32 ? 0 : BN) + c * 32 + (hid & 31)] = nxt;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + c) : "memory");
      if (hid < 64 && j + 1 < n_blk) nxt = __ldg(stat_g + (j + 1) * BN);
      mbar_wait(&bars->sdp_full[c], j & 1);
      tcgen05_fence_after();
      const uint32_t st_addr = smem_u32(st);
      {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {           // 16 columns at a time: keeps the live registers under the 2-CTA budget
          uint32_t s[16], dp[16], pp[8], dsp[8];
          tmem_ld_32x32b_x16(t_lane + KV_T_S + c * 32 + sub * 16, s);
          tmem_ld_32x32b_x16(t_lane + KV_T_DP + c * 32 + sub * 16, dp);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 l4 = lds_f4(st_addr + (c * 32 + sub * 16 + i) * 4);
            const float4 d4 = lds_f4(st_addr + (BN + c * 32 + sub * 16 + i) * 4);
            // packed subtracts / multiplies (3 issue slots per element besides the two shared-memory loads); an invalid
            // key row adds -inf to every exponent (key_bias), so its probabilities are exactly 0 without per-element selects
            const float2 x0 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(s[i + 0]), __uint_as_float(s[i + 1])), make_float2(-l4.x, -l4.y)), kb2);
            const float2 x1 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), make_float2(-l4.z, -l4.w)), kb2);
            const float2 p01 = make_float2(ex2f(x0.x), ex2f(x0.y)), p23 = make_float2(ex2f(x1.x), ex2f(x1.y));
            const float2 g01 = __fmul2_rn(p01, __fadd2_rn(make_float2(__uint_as_float(dp[i + 0]), __uint_as_float(dp[i + 1])), make_float2(-d4.x, -d4.y)));
            const float2 g23 = __fmul2_rn(p23, __fadd2_rn(make_float2(__uint_as_float(dp[i + 2]), __uint_as_float(dp[i + 3])), make_float2(-d4.z, -d4.w)));
            const int o = i >> 1;
            pp[o] = pack_bf16x2(p01.x, p01.y);
            pp[o + 1] = pack_bf16x2(p23.x, p23.y);
            dsp[o] = pack_bf16x2(g01.x, g01.y);
            dsp[o + 1] = pack_bf16x2(g23.x, g23.y);
          }
          tmem_st_x8(t_lane + KV_T_P + c * 16 + sub * 8, pp);
          tmem_st_x8(t_lane + KV_T_DS + c * 16 + sub * 8, dsp);
        }
      }
      tmem_st_wait_();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ds_ready[c]);
    }
    mbar_wait(&bars->acc_full, 0);
    tcgen05_fence_after();
    uint32_t av[16], ak[16];                        // this warp's half of the 32 + 32 accumulator columns
    tmem_ld_32x32b_x16(t_lane + KV_T_DV + c * 16, av);
    tmem_ld_32x32b_x16(t_lane + KV_T_DK + c * 16, ak);
    tmem_ld_wait();
    if (key < Lk) {
      const float ln2 = 0.6931471805599453f;
      uint4* ov = reinterpret_cast<uint4*>(dv + (static_cast<size_t>(b) * Lk + key) * ld_dv + h * DH + c * 16);
      uint4* ok = reinterpret_cast<uint4*>(dk + (static_cast<size_t>(b) * Lk + key) * ld_dk + h * DH + c * 16);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(av[8 * i + 0]), __uint_as_float(av[8 * i + 1]));
        w.y = pack_bf16x2(__uint_as_float(av[8 * i + 2]), __uint_as_float(av[8 * i + 3]));
        w.z = pack_bf16x2(__uint_as_float(av[8 * i + 4]), __uint_as_float(av[8 * i + 5]));
        w.w = pack_bf16x2(__uint_as_float(av[8 * i + 6]), __uint_as_float(av[8 * i + 7]));
        ov[i] = w;
        w.x = pack_bf16x2(__uint_as_float(ak[8 * i + 0]) * ln2, __uint_as_float(ak[8 * i + 1]) * ln2);
        w.y = pack_bf16x2(__uint_as_float(ak[8 * i + 2]) * ln2, __uint_as_float(ak[8 * i + 3]) * ln2);
        w.z = pack_bf16x2(__uint_as_float(ak[8 * i + 4]) * ln2, __uint_as_float(ak[8 * i + 5]) * ln2);
        w.w = pack_bf16x2(__uint_as_float(ak[8 * i + 6]) * ln2, __uint_as_float(ak[8 * i + 7]) * ln2);
        ok[i] = w;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == abwd::COMPUTE_WARPS + 1) {
    tcgen05_fence_after();
    tmem_dealloc<abwd::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
int launch_attn_delta(const svol_bf16* dO, const svol_bf16* O, float* delta, int B, int H, int Lq, int pitch, cudaStream_t stream);

int launch_attention_backward_tc(const svol_attn_bwd_args& a, cudaStream_t stream) {
  using namespace abwd;
  if (a.B <= 0 || a.H != 8 || a.Lq <= 0 || a.Lk <= 0) return svol_fail(SVOL_ERR_SHAPE, "attention_backward: bad sizes (8 heads of 32)");
  if (a.stat_pitch < a.Lq || a.stat_pitch % BN != 0) return svol_fail(SVOL_ERR_SHAPE, "attention_backward: stat_pitch must be a multiple of 64 and >= Lq");
  if (a.ld_dq % 8 || a.ld_dk % 8 || a.ld_dv % 8 || a.ld_o != a.H * DH || a.ld_do != a.H * DH)
    return svol_fail(SVOL_ERR_SHAPE, "attention_backward: gradient pitches must be multiples of 8; o / d_o are [B*Lq, 256]");
  int rc = launch_attn_delta(a.d_o, a.o, a.delta, a.B, a.H, a.Lq, a.stat_pitch, stream);
  if (rc) return rc;
  const int64_t rq = static_cast<int64_t>(a.B) * a.Lq, rk = static_cast<int64_t>(a.B) * a.Lk, rt = static_cast<int64_t>(a.B) * a.H * DH;
  const int W = a.H * DH;
  CUtensorMap q128, do128, k64, v64, kt, k128, v128, q64, do64, qt, dot;
  if ((rc = make_tensor_map_2d(&q128, a.q, W, rq, a.ldq, DH, BM, 64))) return rc;
  if ((rc = make_tensor_map_2d(&do128, a.d_o, W, rq, a.ld_do, DH, BM, 64))) return rc;
  if ((rc = make_tensor_map_2d(&k64, a.k, W, rk, a.ldk, DH, BN, 64))) return rc;
  if ((rc = make_tensor_map_2d(&v64, a.v, W, rk, a.ldv, DH, BN, 64))) return rc;
  if ((rc = make_tensor_map_2d(&kt, a.kt, a.kt_pitch, rt, a.kt_pitch, BN, DH, 128))) return rc;
  if ((rc = make_tensor_map_2d(&k128, a.k, W, rk, a.ldk, DH, BM, 64))) return rc;
  if ((rc = make_tensor_map_2d(&v128, a.v, W, rk, a.ldv, DH, BM, 64))) return rc;
  if ((rc = make_tensor_map_2d(&q64, a.q, W, rq, a.ldq, DH, BN, 64))) return rc;
  if ((rc = make_tensor_map_2d(&do64, a.d_o, W, rq, a.ld_do, DH, BN, 64))) return rc;
  if ((rc = make_tensor_map_2d(&qt, a.qt, a.qt_pitch, rt, a.qt_pitch, BN, DH, 128))) return rc;
  if ((rc = make_tensor_map_2d(&dot, a.d_ot, a.qt_pitch, rt, a.qt_pitch, BN, DH, 128))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM);
    if (e != cudaSuccess) return svol_fail_cuda(e, "attention_backward: cudaFuncSetAttribute");
    configured = true;
  }
  attn_bwd_dq_kernel<<<dim3((a.Lq + BM - 1) / BM, a.H, a.B), THREADS, DQ_SMEM, stream>>>(
      q128, do128, k64, v64, kt, a.lse, a.delta, a.key_mask, reinterpret_cast<__nv_bfloat16*>(a.dq), a.H, a.Lq, a.Lk,
      a.stat_pitch, a.ld_dq);
  if ((rc = svol_check_launch("attn_bwd_dq"))) return rc;
  attn_bwd_dkdv_kernel<<<dim3((a.Lk + BM - 1) / BM, a.H, a.B), THREADS, KV_SMEM, stream>>>(
      k128, v128, q64, do64, qt, dot, a.lse, a.delta, a.key_mask, reinterpret_cast<__nv_bfloat16*>(a.dk),
      reinterpret_cast<__nv_bfloat16*>(a.dv), a.H, a.Lq, a.Lk, a.stat_pitch, a.ld_dk, a.ld_dv);
  return svol_check_launch("attn_bwd_dkdv");
}

}  // namespace svol
