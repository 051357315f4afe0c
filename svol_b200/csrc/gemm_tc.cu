// Persistent warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   out[M,N] = epilogue( A[M,K] (bf16, row-major) x W[N,K]^T (bf16, nn.Linear layout) )
//
// replaces the cuBLAS GEMM + separate bias / activation / residual / LayerNorm kernels that the
// reference's nn.Linear / nn.LayerNorm / F.gelu / F.relu calls turn into
// (lib/modeling/svanet.py:159-181, lib/modeling/cross_modal_transformer.py:88-100,137-158,163-179).
//
// Structure (one CTA per SM, 384 threads = 3 warpgroups; warps 2-3 idle so that setmaxnreg can re-split
// the register file per warpgroup):
//   warp 0      TMA producer
//   warp 1      MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, fp32
//                              accumulators in TMEM, two accumulator stages (2 x 256 columns)
//   warps 4..11 epilogue     : tcgen05.ld -> registers (TMEM stage released immediately) ->
//                              +bias -> ReLU | GELU(erf) -> +residual -> LayerNorm over the
//                              256-wide row -> bf16 store, optional second output (x + pos),
//                              optional per-head transposed store for the attention V operand.
// The epilogue of tile i overlaps the mainloop of tile i+1.
//
// Two operand-feed variants (measured: the 128x256 tile streamed from L2 is bound by L2->SM
// bandwidth, ~48 KB per 64-wide k-block per SM, long before the tensor pipe):
//   kResidentW = true  (K == 256, every projection except three): each CTA owns one 256-row block
//                of W, loads its 128 KB ONCE into shared memory and then streams only A tiles
//                (16 KB per k-block, 4-stage ring) -- 3x less L2 traffic per tile.
//   kResidentW = false (K = 512 input projection, K = 2048 FFN down-projection): A and W tiles are
//                both streamed through a 4-stage ring of 48 KB.
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace gemm {
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int EPI_WARPS = 8, FIRST_EPI_WARP = 4, THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;   // 3 warpgroups
constexpr int COLS_PER_THREAD = BN / (EPI_WARPS / 4);   // 128
constexpr int RES_K = 256, RES_KB = RES_K / BK;         // resident-W variant: K fixed at 256
constexpr int MAX_STAGES = 4;
constexpr int STG_BYTES = 32 * 128;                     // per epilogue warp: 32 rows x 128 B

template <bool kResidentW>
struct Cfg {
  static constexpr int STAGES = kResidentW ? 3 : 4;
  static constexpr int STAGE_BYTES = kResidentW ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int W_BYTES = kResidentW ? RES_KB * B_BYTES : 0;                  // 128 KB
  static constexpr int OFF_STAGES = W_BYTES;
  static constexpr int OFF_STG = OFF_STAGES + STAGES * STAGE_BYTES;                    // epilogue staging
  static constexpr int OFF_TAIL = OFF_STG + EPI_WARPS * STG_BYTES;
  static constexpr int PAR_BYTES = kResidentW ? 3 * BN * 4 : 0;     // bias, ln_w, ln_b of the CTA's fixed column block
  static constexpr int SMEM_BYTES = OFF_TAIL + 256 /*barriers*/ + 2 * BM * 4 /*LN exchange*/ + 512 /*1/dim_t*/ + PAR_BYTES + 1024 /*align*/;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};
}  // namespace gemm

struct GemmSmemTail {
  uint64_t full[gemm::MAX_STAGES];
  uint64_t empty[gemm::MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t w_full[gemm::RES_KB];     // resident weights: one barrier per 64-wide k block, so the first MMAs start after 32 KB, not 128
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(GemmSmemTail) <= 256, "barrier block");

#ifdef SVOL_GEMM_TRACE
// Debug build only (-DSVOL_GEMM_TRACE): CTA 0 records clock64() at the phase boundaries of its first 32 tiles.
// role 0: epilogue warp 4; role 1: MMA issuer.
__device__ long long g_gemm_trace[2][32][8];
#define SVOL_GTR(role, it, slot)                                                           \
  do {                                                                                     \
    if (gtrace_on && (it) < 32) {                                                          \
      long long c_;                                                                        \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)::"memory");                          \
      g_gemm_trace[role][it][slot] = c_;                                                   \
    }                                                                                      \
  } while (0)
#else
#define SVOL_GTR(role, it, slot) do {} while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Global <-> register transposition of a [32 rows x 128 B] block through a per-warp staging buffer.
// In the epilogue a thread owns one ROW (that is how tcgen05.ld hands out the accumulator), so a direct
// 16-byte access per thread touches 32 different cache lines per warp instruction; measured, that made
// the epilogue L1-wavefront bound (~8 us per tile).  Staged, every global instruction covers 4 rows x 128
// contiguous bytes.  16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4): conflict-free both ways.
// The staging buffer is addressed as SHARED memory explicitly (a 32-bit address kept in a register, sts_u4 / lds_u4 of
// common.cuh): through the 1024-byte aligned generic pointer ptxas emitted generic LD.E / ST.E for every staged 16-byte chunk
// (cuobjdump: 156 generic against 101 shared accesses in the resident-weight kernel).
__device__ __forceinline__ void block_store(uint32_t stg, const uint4 (&q)[8], uint8_t* gbase, size_t pitch,
                                            int rows_valid, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) sts_u4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), q[j]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = k * 4 + (lane >> 3), ch = lane & 7;
    const uint4 val = lds_u4(stg + row * 128 + ((ch ^ (row & 7)) << 4));
    if (row < rows_valid) *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(row) * pitch + ch * 16) = val;
  }
  __syncwarp();
}
__device__ __forceinline__ void block_load(uint32_t stg, uint4 (&q)[8], const uint8_t* gbase, size_t pitch,
                                           int rows_valid, int lane) {
  if (lane == 0) tma_store_wait_read<0>();      // a TMA store may still be reading this warp's staging buffer
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = k * 4 + (lane >> 3), ch = lane & 7;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (row < rows_valid) val = __ldg(reinterpret_cast<const uint4*>(gbase + static_cast<size_t>(row) * pitch + ch * 16));
    sts_u4(stg + row * 128 + ((ch ^ (row & 7)) << 4), val);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = lds_u4(stg + lane * 128 + ((j ^ (lane & 7)) << 4));
  __syncwarp();
}

// [32 rows x 64 bf16 columns] of this warp -> global through its staging buffer and ONE asynchronous TMA store
// (box 64 x 32, 128B swizzle == the staging layout above): 8 STS per lane instead of 8 STS + 8 LDS + 8 STG, and the
// warp does not wait for the global writes.  Rows beyond M are clipped by the tensor map.
__device__ __forceinline__ void block_store_tma(uint32_t stg, const uint4 (&q)[8], const CUtensorMap* tm, int col, int row0,
                                                int lane) {
  if (lane == 0) tma_store_wait_read<0>();      // the previous store has finished reading the staging buffer
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) sts_u4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), q[j]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d_a(tm, stg, col, row0);
    tma_store_commit();
  }
}

// kTrain compiles in the training-step options (split-K fp32 accumulation, out_pre, fused activation backward); the
// inference instantiations are exactly the kernels the forward plan was tuned with.
template <bool kResidentW, bool kTrain>
__global__ void __launch_bounds__(gemm::THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOutPos,
                    const __grid_constant__ CUtensorMap tmA2, const GemmEpilogue ep, int M, int N, int K, int split_block,
                    int k_splits, float* __restrict__ out_f32, int ld_f32, int mn_major, int l2_prefetch, int pdl) {
  using namespace gemm;
  using C = Cfg<kResidentW>;
  // mn_major (training, streaming variant): the operands are given untransposed, A = dY [K rows, M columns] and
  // W = X [K rows, N columns] (row-major, the contraction index is the ROW): tiles are loaded as 64-column x 64-row TMA
  // boxes and consumed through MN-major shared-memory descriptors -- dW = dY^T X without a transposition pass.
  const bool mn = kTrain && !kResidentW && mn_major != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem + C::OFF_STAGES;
  GemmSmemTail* tail = reinterpret_cast<GemmSmemTail*>(smem + C::OFF_TAIL);
  float* ln_x = reinterpret_cast<float*>(smem + C::OFF_TAIL + 256);   // [2 halves][128 rows]
  float* s_idt = ln_x + 2 * BM;                                       // [128] 1 / dim_t of the sine positional encoding
  float* s_par = s_idt + BN / 2;                                      // resident variant: bias | ln_w | ln_b of column block my_n

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef SVOL_GEMM_TRACE
  const bool gtrace_on = blockIdx.x == 0 && lane == 0 && (warp == gemm::FIRST_EPI_WARP || warp == 1);
#endif
  const int m_blocks = (M + BM - 1) / BM;
  const int n_blocks = N / BN;
  const int num_kb = (K + BK - 1) / BK;        // K % 64 != 0 only with mn_major operands (rows beyond K read as zero)
  // tile schedule.  streaming: tile = blockIdx.x + i*grid, n fastest.  resident: this CTA's n block is
  // fixed (grid is a multiple of n_blocks) and it walks m blocks with stride grid / n_blocks.
  const int my_n = kResidentW ? static_cast<int>(blockIdx.x) % n_blocks : 0;
  const int m_first = kResidentW ? static_cast<int>(blockIdx.x) / n_blocks : 0;
  const int m_stride = kResidentW ? static_cast<int>(gridDim.x) / n_blocks : 0;
  // split-K (streaming variant, weight gradients: few output tiles, a contraction over all token rows): a work item
  // is (tile, k slice); slices accumulate into out_f32 with fp32 atomics.  k_splits == 1 otherwise.
  if (!kTrain) k_splits = 1;
  const int kb_per_split = (num_kb + k_splits - 1) / k_splits;
  const int num_tiles = m_blocks * n_blocks * k_splits;
  const int my_tiles = kResidentW ? (m_blocks > m_first ? (m_blocks - m_first + m_stride - 1) / m_stride : 0)
                                  : (num_tiles > static_cast<int>(blockIdx.x)
                                         ? (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                                         : 0);
  // split launch (resident variant): column blocks >= split_block read A2 and write the transposed output
  const bool second_part = kResidentW && split_block > 0 && my_n >= split_block;
  const bool do_out = !second_part;                                  // out / out_pos columns belong to the first part
  const bool do_vt = ep.out_vt != nullptr && (split_block == 0 || second_part);
  const int vt_col_shift = second_part ? split_block * BN : 0;
  auto tile_coords = [&](int it, int& m_blk, int& n_blk) {
    if (kResidentW) { m_blk = m_first + it * m_stride; n_blk = my_n; }
    else { const int tile = (blockIdx.x + it * gridDim.x) / k_splits; m_blk = tile / n_blocks; n_blk = tile % n_blocks; }
  };
  auto k_range = [&](int it, int& kb0, int& kb1) {      // k-blocks [kb0, kb1) of work item `it`
    if (!kTrain || kResidentW || k_splits == 1) { kb0 = 0; kb1 = num_kb; return; }
    const int split = (blockIdx.x + it * gridDim.x) % k_splits;
    kb0 = split * kb_per_split;
    kb1 = min(num_kb, kb0 + kb_per_split);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tail->tmem_full[s], 1); mbar_init(&tail->tmem_empty[s], EPI_WARPS); }
    for (int kb = 0; kb < RES_KB; ++kb) mbar_init(&tail->w_full[kb], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tail->tmem_base);
  if (kResidentW) {                 // per-column parameters of this CTA's (fixed) column block -> shared memory
    for (int i = threadIdx.x; i < BN; i += THREADS) {
      s_par[i] = ep.bias ? __ldg(ep.bias + my_n * BN + i) : 0.f;
      s_par[BN + i] = ep.ln_weight ? __ldg(ep.ln_weight + my_n * BN + i) : 1.f;
      s_par[2 * BN + i] = ep.ln_weight ? __ldg(ep.ln_bias + my_n * BN + i) : 0.f;
    }
  }
  if (ep.pos_theta != nullptr)      // 1 / dim_t[2k] = 10000^(-2k/256)  (position_encoding.py:64-65)
    for (int i = threadIdx.x; i < BN / 2; i += THREADS)
      s_idt[i] = 1.0f / powf(10000.f, __fdiv_rn(__fmul_rn(2.f, static_cast<float>(i)), static_cast<float>(BN)));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  // Register re-split between warpgroups (each SM sub-partition holds one warp of warpgroup 0 and two
  // epilogue warps): the producer / MMA warpgroup keeps 40 registers, the epilogue warps get 232 so the
  // 128-wide accumulator row, its residual and its pos operand stay in registers.
  // (the setmaxnreg must dominate the role's code, so each role branch starts with its own).
  if (warp < FIRST_EPI_WARP) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      // Programmatic dependent launch (pdl): this CTA may have started under the tail of the kernel that produces A / the
      // residual.  The resident weight block does not depend on it, so all of it is requested first and only then does the
      // producer wait for the predecessor (otherwise the weight k blocks stay interleaved with the first A tile's).
      const bool w_first = kResidentW && pdl && my_tiles > 0;
      if (w_first)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_arrive_expect_tx(&tail->w_full[kb], B_BYTES);
          tma_load_2d(smem + kb * B_BYTES, &tmB, &tail->w_full[kb], kb * BK, my_n * BN);
        }
      griddep_wait();
      for (int it = 0; it < my_tiles; ++it) {
        int m_blk, n_blk, kb0, kb1;
        tile_coords(it, m_blk, n_blk);
        k_range(it, kb0, kb1);
        // last tile of this CTA: the next kernel on the stream may take the SMs that free up from here on
        if (it == my_tiles - 1) griddep_launch_dependents();
        if (kResidentW && l2_prefetch) {
          // One CTA per SM and a 3-stage ring (48 KB) keep less than one A tile in flight: the loads were latency-bound
          // (phase trace: 3.9 k clk per tile in the MMA issuer against 2.3 k of tensor work) and the epilogue's residual read
          // paid a full DRAM round trip per tile (4.6 k of 11.1 k clk).  The producer therefore pulls the NEXT tile's A
          // block and THIS tile's residual rows (consumed two tile periods from now) into L2 ahead of time.
          if (it + 1 < my_tiles) {
            const int m_next = m_first + (it + 1) * m_stride;
            for (int kb = kb0; kb < kb1; ++kb) tma_prefetch_2d(second_part ? &tmA2 : &tmA, kb * BK, m_next * BM);
          }
          if (ep.residual != nullptr) {
            const int rows = min(BM, M - m_blk * BM);
            const svol_bf16* rp = ep.residual + static_cast<size_t>(m_blk) * BM * ep.ld_res + n_blk * BN;
            if (ep.ld_res == BN) bulk_prefetch_l2(rp, static_cast<uint32_t>(rows) * BN * 2);
            else for (int r = 0; r < rows; ++r) bulk_prefetch_l2(rp + static_cast<size_t>(r) * ep.ld_res, BN * 2);
          }
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kResidentW && it == 0 && !w_first) {      // this CTA's weight block, k block by k block, interleaved with the first A tiles
            mbar_arrive_expect_tx(&tail->w_full[kb], B_BYTES);
            tma_load_2d(smem + kb * B_BYTES, &tmB, &tail->w_full[kb], kb * BK, my_n * BN);
          }
          mbar_wait(&tail->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&tail->full[stage], C::STAGE_BYTES);
          uint8_t* sa = stages + stage * C::STAGE_BYTES;
          if (mn) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, &tail->full[stage], m_blk * BM + c * 64, kb * BK);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sa + A_BYTES + c * 8192, &tmB, &tail->full[stage], n_blk * BN + c * 64, kb * BK);
          } else {
            tma_load_2d(sa, second_part ? &tmA2 : &tmA, &tail->full[stage], kb * BK, m_blk * BM);
            if (!kResidentW) tma_load_2d(sa + A_BYTES, &tmB, &tail->full[stage], kb * BK, n_blk * BN);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        SVOL_GTR(1, it, 0);
        mbar_wait(&tail->tmem_empty[acc], acc_phase ^ 1);
        SVOL_GTR(1, it, 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        int kb0, kb1;
        k_range(it, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kResidentW && it == 0) mbar_wait(&tail->w_full[kb], 0);
          mbar_wait(&tail->full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(stages + stage * C::STAGE_BYTES);
          if (mn) {
            // [64 MN x 64 K] boxes of 8 KB; a K16 step advances 16 rows of 128 bytes
            constexpr uint32_t idesc_mn = make_idesc_bf16_mn(BM, BN);
            const uint64_t a_desc = make_mnmajor_desc_sw128(sa, 8192, 1024);
            const uint64_t b_desc = make_mnmajor_desc_sw128(sa + A_BYTES, 8192, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss(d_tmem, a_desc + 128 * k, b_desc + 128 * k, idesc_mn, ((kb - kb0) | k) != 0);
          } else {
            const uint64_t a_desc = make_kmajor_desc<128>(sa);
            const uint64_t b_desc = make_kmajor_desc<128>(kResidentW ? smem_u32(smem + kb * B_BYTES) : sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, ((kb - kb0) | k) != 0);
          }
          umma_commit(&tail->empty[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tail->tmem_full[acc]);
        SVOL_GTR(1, it, 2);
      }
    }
   }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    griddep_wait();                             // residual / pos operands are read with plain loads below
    const int quarter = warp & 3;               // TMEM lane quarter this warp may read
    const int half = (warp - FIRST_EPI_WARP) >> 2;           // which 128 of the tile's 256 columns
    const int r_in_tile = quarter * 32 + lane;
    uint32_t stg = smem_u32(smem + C::OFF_STG + (warp - FIRST_EPI_WARP) * STG_BYTES);
    asm volatile("" : "+r"(stg));               // computed once, kept in a register
    for (int it = 0; it < my_tiles; ++it) {
      int m_blk, n_blk;
      tile_coords(it, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_blk * BM + r_in_tile;
      const int col0 = n_blk * BN + half * COLS_PER_THREAD;
      const bool row_ok = row < M;
      const int slab_row0 = m_blk * BM + quarter * 32;             // first row of this warp's 32-row slab
      const int rows_valid = min(32, max(0, M - slab_row0));

      SVOL_GTR(0, it, 0);
      mbar_wait(&tail->tmem_full[acc], acc_phase);
      SVOL_GTR(0, it, 1);
      tcgen05_fence_after();
      float v[COLS_PER_THREAD];
      {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * COLS_PER_THREAD;
        uint32_t raw[COLS_PER_THREAD / 32][32];
#pragma unroll
        for (int c = 0; c < COLS_PER_THREAD / 32; ++c) tmem_ld_32x32b_x32(taddr + c * 32, raw[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < COLS_PER_THREAD / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c * 32 + i] = __uint_as_float(raw[c][i]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tmem_empty[acc]);
      SVOL_GTR(0, it, 2);

      if (kTrain && out_f32 != nullptr) {
        // weight-gradient mode: out_f32[row, col] += partial sum of this k slice (no bias / activation / bf16 output).
        // An empty slice (kb0 >= kb1) issued no MMA: its accumulator is stale, nothing may be added.
        int kb0, kb1;
        k_range(it, kb0, kb1);
        if (row_ok && kb0 < kb1) {
          float* orow = out_f32 + static_cast<size_t>(row) * ld_f32 + col0;
#pragma unroll
          for (int i = 0; i < COLS_PER_THREAD; i += 4)      // 16-byte vector reductions: 4x fewer L2 atomic operations
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]),
                         "f"(v[i + 3])
                         : "memory");
        }
        __syncwarp();
        continue;
      }

      if (ep.bias) {
        if (kResidentW) {
          const uint32_t bp = smem_u32(s_par + half * COLS_PER_THREAD);
#pragma unroll
          for (int i = 0; i < COLS_PER_THREAD / 4; ++i) {
            const float4 b = lds_f4(bp + i * 16);
            v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
          }
        } else {
          const float4* bp = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
          for (int i = 0; i < COLS_PER_THREAD / 4; ++i) {
            const float4 b = __ldg(bp + i);
            v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
          }
        }
      }
      if (kTrain && ep.out_pre && do_out) {
        // training forward: keep the pre-activation (second bf16 output through the out_pos tensor map)
#pragma unroll
        for (int blk = 0; blk < COLS_PER_THREAD / 64; ++blk) {
          uint4 q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float* vv = &v[blk * 64 + j * 8];
            q[j] = make_uint4(pack_bf16x2(vv[0], vv[1]), pack_bf16x2(vv[2], vv[3]), pack_bf16x2(vv[4], vv[5]), pack_bf16x2(vv[6], vv[7]));
          }
          block_store_tma(stg, q, &tmOutPos, col0 + blk * 64, slab_row0, lane);
        }
      }
      if (ep.act == SVOL_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (ep.act == SVOL_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) v[i] = gelu_erf_fast(v[i]);
      }
      if (kTrain && ep.dact_src) {
        // training backward: dX = (dY W) * f'(saved), the activation backward fused into the dgrad GEMM
        const uint8_t* dbase = reinterpret_cast<const uint8_t*>(ep.dact_src + static_cast<size_t>(slab_row0) * ep.ld_dact + col0);
#pragma unroll
        for (int blk = 0; blk < COLS_PER_THREAD / 64; ++blk) {
          uint4 q[8];
          block_load(stg, q, dbase + blk * 128, static_cast<size_t>(ep.ld_dact) * 2, rows_valid, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float* vv = &v[blk * 64 + j * 8];
            const float sv[8] = {bf16_lo(q[j].x), bf16_hi(q[j].x), bf16_lo(q[j].y), bf16_hi(q[j].y),
                                 bf16_lo(q[j].z), bf16_hi(q[j].z), bf16_lo(q[j].w), bf16_hi(q[j].w)};
            if (ep.dact_mode == SVOL_ACT_GELU) {
#pragma unroll
              for (int e = 0; e < 8; ++e) vv[e] *= gelu_grad_fast(sv[e]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) vv[e] = sv[e] > 0.f ? vv[e] : 0.f;
            }
          }
        }
      }
      SVOL_GTR(0, it, 3);
      if (ep.residual) {
        // both 64-column blocks of the residual are requested before either is consumed: one DRAM round trip per tile
        // instead of two dependent ones (phase trace: 6.5 k of the 11.8 k-clock tile period of the attention output projection)
        const uint8_t* rbase = reinterpret_cast<const uint8_t*>(ep.residual + static_cast<size_t>(slab_row0) * ep.ld_res + col0);
        uint4 raw[COLS_PER_THREAD / 64][8];
#pragma unroll
        for (int blk = 0; blk < COLS_PER_THREAD / 64; ++blk) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = k * 4 + (lane >> 3), ch = lane & 7;
            raw[blk][k] = make_uint4(0u, 0u, 0u, 0u);
            if (rr < rows_valid)
              raw[blk][k] = __ldg(reinterpret_cast<const uint4*>(rbase + blk * 128 + static_cast<size_t>(rr) * ep.ld_res * 2 + ch * 16));
          }
        }
#pragma unroll
        for (int blk = 0; blk < COLS_PER_THREAD / 64; ++blk) {
          uint4 q[8];
          if (lane == 0) tma_store_wait_read<0>();      // a TMA store may still be reading this warp's staging buffer
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = k * 4 + (lane >> 3), ch = lane & 7;
            sts_u4(stg + rr * 128 + ((ch ^ (rr & 7)) << 4), raw[blk][k]);
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) q[j] = lds_u4(stg + lane * 128 + ((j ^ (lane & 7)) << 4));
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float* vv = &v[blk * 64 + j * 8];
            vv[0] += bf16_lo(q[j].x); vv[1] += bf16_hi(q[j].x); vv[2] += bf16_lo(q[j].y); vv[3] += bf16_hi(q[j].y);
            vv[4] += bf16_lo(q[j].z); vv[5] += bf16_hi(q[j].z); vv[6] += bf16_lo(q[j].w); vv[7] += bf16_hi(q[j].w);
          }
        }
      }
      SVOL_GTR(0, it, 4);
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
        const uint32_t sgp = smem_u32(s_par + BN + half * COLS_PER_THREAD), sbp = smem_u32(s_par + 2 * BN + half * COLS_PER_THREAD);
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD / 4; ++i) {
          const float4 g = kResidentW ? lds_f4(sgp + i * 16) : __ldg(gp + i);
          const float4 b = kResidentW ? lds_f4(sbp + i * 16) : __ldg(bp + i);
          v[4 * i + 0] = (v[4 * i + 0] - mean) * rstd * g.x + b.x;
          v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * g.y + b.y;
          v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * g.z + b.z;
          v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * g.w + b.w;
        }
      }
      SVOL_GTR(0, it, 5);
      if (ep.out && do_out) {
#pragma unroll
        for (int blk = 0; blk < COLS_PER_THREAD / 64; ++blk) {
          uint4 q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float* vv = &v[blk * 64 + j * 8];
            q[j] = make_uint4(pack_bf16x2(vv[0], vv[1]), pack_bf16x2(vv[2], vv[3]), pack_bf16x2(vv[4], vv[5]), pack_bf16x2(vv[6], vv[7]));
          }
          block_store_tma(stg, q, &tmOut, col0 + blk * 64, slab_row0, lane);
        }
      }
      SVOL_GTR(0, it, 6);
      if (ep.out_pos && do_out) {
        // second output: x + pos (the q/k operand of the next attention block).  pos rows follow the output
        // rows (pos_row_mod == 0) or repeat with period pos_row_mod (query embedding broadcast over the batch);
        // the broadcast case is read directly (its 320 x 256 table stays in L1/L2).
        const int prow = ep.pos_row_mod > 0 ? row % ep.pos_row_mod : row;
        const float theta = (ep.pos_theta != nullptr && row_ok) ? __ldg(ep.pos_theta + row) : 0.f;
#pragma unroll
        for (int blk = 0; blk < COLS_PER_THREAD / 64; ++blk) {
          uint4 q[8];
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            uint4 pq[8];                                              // 32 fp32 pos values of this row
            if (ep.pos_theta != nullptr) {
              // sine encoding in place: columns (2k, 2k+1) = (sin, cos)(theta / dim_t[2k]); theta in [0, 2 pi] is folded
              // to [-pi, pi] for the MUFU sin / cos
              const float* it = s_idt + (half * COLS_PER_THREAD + blk * 64 + sub * 32) / 2;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float a0 = theta * it[2 * j], a1 = theta * it[2 * j + 1];
                a0 = a0 > 3.14159265358979f ? a0 - 6.28318530717959f : a0;
                a1 = a1 > 3.14159265358979f ? a1 - 6.28318530717959f : a1;
                pq[j] = make_uint4(__float_as_uint(__sinf(a0)), __float_as_uint(__cosf(a0)), __float_as_uint(__sinf(a1)),
                                   __float_as_uint(__cosf(a1)));
              }
            } else if (ep.pos_row_mod > 0) {
              const uint4* pp = reinterpret_cast<const uint4*>(ep.pos + static_cast<size_t>(prow) * ep.ld_pos + col0 + blk * 64 + sub * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) pq[j] = row_ok ? __ldg(pp + j) : make_uint4(0u, 0u, 0u, 0u);
            } else {
              block_load(stg, pq, reinterpret_cast<const uint8_t*>(ep.pos + static_cast<size_t>(slab_row0) * ep.ld_pos + col0 + blk * 64 + sub * 32),
                         static_cast<size_t>(ep.ld_pos) * 4, rows_valid, lane);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float* vv = &v[blk * 64 + sub * 32 + j * 8];
              const uint4 a = pq[2 * j], c = pq[2 * j + 1];
              q[sub * 4 + j] = make_uint4(pack_bf16x2(vv[0] + __uint_as_float(a.x), vv[1] + __uint_as_float(a.y)),
                                          pack_bf16x2(vv[2] + __uint_as_float(a.z), vv[3] + __uint_as_float(a.w)),
                                          pack_bf16x2(vv[4] + __uint_as_float(c.x), vv[5] + __uint_as_float(c.y)),
                                          pack_bf16x2(vv[6] + __uint_as_float(c.z), vv[7] + __uint_as_float(c.w)));
            }
          }
          block_store_tma(stg, q, &tmOutPos, col0 + blk * 64, slab_row0, lane);
        }
      }
      if (row_ok && do_vt) {
        // per-head transposed store: Vt[(b*H + h)*dh + d][l], row = b*L + l, col = h*dh + d.
        // Consecutive lanes hold consecutive tokens l, so each store instruction writes 64
        // contiguous bytes per output row.
        const int b = row / ep.vt_len, l = row - b * ep.vt_len;
        const int n_vt = N - vt_col_shift;                       // columns of the transposed output
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ep.out_vt) +
                              (static_cast<size_t>(b) * n_vt + (col0 - vt_col_shift)) * ep.vt_pitch + l;
#pragma unroll
        for (int i = 0; i < COLS_PER_THREAD; ++i) base[static_cast<size_t>(i) * ep.vt_pitch] = __float2bfloat16_rn(v[i]);
      }
      __syncwarp();   // reconverge before the next tile's warp-collective tcgen05.ld
      SVOL_GTR(0, it, 7);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's TMA stores are complete
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
template <bool kResidentW, bool kTrain>
static int launch_variant(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                          const CUtensorMap& tmOutPos, const CUtensorMap& tmA2, int k_splits, cudaStream_t stream) {
  using namespace gemm;
  using C = Cfg<kResidentW>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel<kResidentW, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return svol_fail_cuda(e, "gemm: cudaFuncSetAttribute");
    configured = true;
  }
  const int m_blocks = (a.M + BM - 1) / BM, n_blocks = a.N / BN;
  const int tiles = m_blocks * n_blocks * k_splits;
  int grid = tiles < sm_count() ? tiles : sm_count();
  if (kResidentW) grid = grid / n_blocks * n_blocks;       // every CTA owns one n block
  if (kResidentW && k_splits == 1 && grid >= n_blocks) {
    // only as many CTAs as the slowest one needs rounds (392 row tiles on 148 SMs: 131 CTAs x 3 tiles finish when 148 CTAs
    // with 3 / 2 tiles would, and leave 17 SMs to concurrent graph branches)
    const int per_n = grid / n_blocks;
    const int rounds = (m_blocks + per_n - 1) / per_n;
    grid = ((m_blocks + rounds - 1) / rounds) * n_blocks;
  } else if (!kResidentW && grid > 0) {
    const int rounds = (tiles + grid - 1) / grid;
    grid = (tiles + rounds - 1) / rounds;
  }
  const char* env_pf = getenv("SVOL_GEMM_L2_PREFETCH");       // read per launch (A/B measurements); default on
  const int l2_prefetch = env_pf ? atoi(env_pf) : 1;
  // inference instantiations: programmatic dependent launch (set-up and the resident weight block overlap the tail of the
  // kernel before this one on the stream); the training kernels are launched the ordinary way
  if (!kTrain) {
    cudaError_t e = launch_kernel_pdl(gemm_bf16_tc_kernel<kResidentW, kTrain>, dim3(grid), dim3(THREADS), C::SMEM_BYTES, stream, tmA, tmB,
                                      tmOut, tmOutPos, tmA2, a.ep, a.M, a.N, a.K, a.split_block, k_splits, a.out_f32, a.ld_f32,
                                      a.mn_major, l2_prefetch, pdl_enabled() ? 1 : 0);
    if (e != cudaSuccess) return svol_fail_cuda(e, "gemm_bf16_tc launch");
    return svol_check_launch("gemm_bf16_tc");
  }
  gemm_bf16_tc_kernel<kResidentW, kTrain><<<grid, THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, tmOut, tmOutPos, tmA2, a.ep, a.M, a.N, a.K,
                                                                            a.split_block, k_splits, a.out_f32, a.ld_f32, a.mn_major,
                                                                            l2_prefetch, 0);
  return svol_check_launch("gemm_bf16_tc");
}

int launch_gemm_bf16_tc(const GemmArgs& a, cudaStream_t stream) {
  using namespace gemm;
  if (a.mn_major) {
    if (!a.out_f32) return svol_fail(SVOL_ERR_SHAPE, "gemm: mn_major operands are the weight-gradient mode (out_f32)");
    if (a.N % BN != 0 || a.M <= 0 || a.K <= 0 || a.M % 8 != 0) return svol_fail(SVOL_ERR_SHAPE, "gemm (mn_major): need N % 256 == 0, M % 8 == 0");
  } else if (a.N % BN != 0 || a.K % BK != 0 || a.M <= 0) return svol_fail(SVOL_ERR_SHAPE, "gemm: need N % 256 == 0, K % 64 == 0, M > 0");
  if (a.ep.ln_weight && a.N != BN) return svol_fail(SVOL_ERR_SHAPE, "gemm: fused LayerNorm needs N == 256");
  const int split = a.split_block;
  if (split != 0 && (split < 0 || split >= a.N / BN || a.K != RES_K || !a.A2 || !a.ep.out_vt || a.ep.ln_weight))
    return svol_fail(SVOL_ERR_SHAPE, "gemm: split launch needs K == 256, 0 < split_block < N/256, A2 and out_vt, no LayerNorm");
  if (a.ep.out_vt && (a.N - split * BN) != BN) return svol_fail(SVOL_ERR_SHAPE, "gemm: transposed-V store needs 256 columns");
  if (a.ep.pos_theta && a.N != BN) return svol_fail(SVOL_ERR_SHAPE, "gemm: in-epilogue sine positions need N == 256");
  if (a.ep.out_pos && !a.ep.pos && !a.ep.pos_theta) return svol_fail(SVOL_ERR_NULL, "gemm: out_pos needs pos or pos_theta");
  if (a.ep.out_pre && a.ep.out_pos) return svol_fail(SVOL_ERR_SHAPE, "gemm: out_pre and out_pos are mutually exclusive");
  if (a.ep.dact_src && (a.ep.ld_dact % 8 != 0 || (a.ep.dact_mode != SVOL_ACT_GELU && a.ep.dact_mode != SVOL_ACT_RELU)))
    return svol_fail(SVOL_ERR_SHAPE, "gemm: dact_src needs ld_dact % 8 == 0 and dact_mode RELU | GELU");
  CUtensorMap tmA, tmB;
  int rc;
  if (a.mn_major) {      // A [K rows, M cols], W [K rows, N cols]: boxes of 64 columns x 64 rows (rows beyond K read as zero)
    rc = make_tensor_map_2d(&tmA, a.A, a.M, a.K, a.lda, 64, BK, 128);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmB, a.W, a.N, a.K, a.ldw, 64, BK, 128);
  } else {
    rc = make_tensor_map_2d(&tmA, a.A, a.K, a.M, a.lda, BK, BM, 128);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmB, a.W, a.K, a.N, a.ldw, BK, BN, 128);
  }
  if (rc) return rc;
  // outputs are written by TMA stores of [32 rows x 64 columns] boxes (rows beyond M clipped by the map)
  CUtensorMap tmOut = tmA, tmOutPos = tmA;
  const int n_out = split > 0 ? split * BN : a.N;          // columns of out / out_pos
  CUtensorMap tmA2 = tmA;
  if (a.ep.out) {
    rc = make_tensor_map_2d(&tmOut, a.ep.out, n_out, a.M, a.ep.ld_out, 64, 32, 128);
    if (rc) return rc;
  }
  if (a.ep.out_pos) {
    rc = make_tensor_map_2d(&tmOutPos, a.ep.out_pos, n_out, a.M, a.ep.ld_out, 64, 32, 128);
    if (rc) return rc;
  } else if (a.ep.out_pre) {
    rc = make_tensor_map_2d(&tmOutPos, a.ep.out_pre, n_out, a.M, a.ep.ld_out, 64, 32, 128);
    if (rc) return rc;
  }
  if (split > 0) {
    rc = make_tensor_map_2d(&tmA2, a.A2, a.K, a.M, a.lda2, BK, BM, 128);
    if (rc) return rc;
  }
  const bool resident = a.K == RES_K && a.N / BN <= sm_count() && a.out_f32 == nullptr;
  if (split > 0 && !resident) return svol_fail(SVOL_ERR_SHAPE, "gemm: split launch needs the resident-weight variant");
  int k_splits = 1;
  if (a.out_f32 != nullptr) {
    // accumulate-into-fp32 mode (weight gradients): split the contraction so that ~all SMs get a work item, at
    // least 4 k-blocks per slice
    if (a.ep.bias || a.ep.act != SVOL_ACT_NONE || a.ep.residual || a.ep.ln_weight || a.ep.out || a.ep.out_pos || a.ep.out_vt || split ||
        a.ep.out_pre || a.ep.dact_src)
      return svol_fail(SVOL_ERR_SHAPE, "gemm: out_f32 (accumulating fp32 output) excludes every other epilogue option");
    if (a.ld_f32 < a.N || a.ld_f32 % 4 != 0 || (reinterpret_cast<uintptr_t>(a.out_f32) & 15))
      return svol_fail(SVOL_ERR_SHAPE, "gemm: out_f32 must be 16-byte aligned with ld_f32 >= N, ld_f32 % 4 == 0");
    const int tiles = ((a.M + BM - 1) / BM) * (a.N / BN), num_kb = (a.K + BK - 1) / BK;
    k_splits = sm_count() / tiles;
    if (k_splits > num_kb / 4) k_splits = num_kb / 4;
    if (k_splits < 1) k_splits = 1;
    const int per = (num_kb + k_splits - 1) / k_splits;
    k_splits = (num_kb + per - 1) / per;                     // no empty slices
  }
  const bool train = a.out_f32 != nullptr || a.ep.out_pre != nullptr || a.ep.dact_src != nullptr;
  if (train)
    return resident ? launch_variant<true, true>(a, tmA, tmB, tmOut, tmOutPos, tmA2, k_splits, stream)
                    : launch_variant<false, true>(a, tmA, tmB, tmOut, tmOutPos, tmA2, k_splits, stream);
  return resident ? launch_variant<true, false>(a, tmA, tmB, tmOut, tmOutPos, tmA2, k_splits, stream)
                  : launch_variant<false, false>(a, tmA, tmB, tmOut, tmOutPos, tmA2, k_splits, stream);
}

}  // namespace svol

#ifdef SVOL_GEMM_TRACE
extern "C" int svol_debug_gemm_trace(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, svol::g_gemm_trace, sizeof(svol::g_gemm_trace)));
}
#endif
