// Flash-style multi-head attention core on tcgen05 / TMEM / TMA (sm_100a), head_dim = 32.
//
// Replaces softmax(Q K^T / sqrt(dh) + key_padding_mask) V inside nn.MultiheadAttention at
// lib/modeling/cross_modal_transformer.py:139 (video self-attention, L x L), :147 (query
// self-attention, Q x Q) and :154 (query -> video cross-attention, Q x L, padded keys masked).
// The reference materialises the (B*8, Lq, Lk) score tensor (2.5 GB at the headline config); here
// scores only ever exist as 128 x 128 fp32 tiles in tensor memory.
//
// One CTA per (pair of 128-query tiles, head, sample), one CTA per SM, 384 threads (warps 10-11 idle, so
// that setmaxnreg can move registers from the TMA/MMA warpgroup to the softmax warpgroups):
//   warps 0..3  softmax warpgroup A: thread = query row of tile A
//   warps 4..7  softmax warpgroup B: thread = query row of tile B
//               per key tile: S tile TMEM -> 128 registers in ONE pass (the TMEM buffer is released
//               to the MMA warp immediately, so the next QK^T overlaps this tile's exponentials),
//               row max, ex2, bf16 P tile -> TENSOR MEMORY (tcgen05.st; the P V MMA reads its A operand
//               from TMEM, so probabilities never touch shared memory and no generic->async proxy fence
//               is needed); running max / sum / 32-wide output accumulator stay in registers.
//   warp 8      TMA producer: both Q tiles once, then a 3-stage ring of K tiles (128 x 32, 64B swizzle)
//               and V^T tiles (32 x 128, 128B swizzle) shared by the two query tiles.
//   warp 9      MMA issuer: S = Q K^T (M128 N128 K32, both operands in shared memory) and
//               O_tile = P V (M128 N48 K128, A = P in TMEM, B = V^T in shared memory) for both tiles,
//               fp32 accumulators in TMEM.
// TMEM map (512 columns): S_A [0,128) S_B [128,256) P_A [256,320) P_B [320,384) O_A [384,432) O_B [448,496).
// Each SM sub-partition hosts one warp of A and one of B; while one of them is in its ex2 section the other
// loads / reduces / stores, which keeps the MUFU pipe (16 ex2/clk/SM) -- the binding unit at head_dim 32:
// 128 tensor FLOPs per exponential -- busy.  See DESIGN.md.
// Q is pre-scaled by log2(e)/sqrt(dh) when it is produced, so the softmax is a bare ex2.
#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace attn {
constexpr int BQ = 128, BKV = 128, DH = 32, STAGES = 3;
constexpr int Q_BYTES = BQ * DH * 2;            // 8192 per query tile
constexpr int K_BYTES = BKV * DH * 2;           // 8192
constexpr int NV = DH + 16;                     // V^T rows fed to the P V MMA: 32 value rows, one row of ones
                                                // (-> column 32 of the product is the softmax row sum), 15 zero rows
constexpr int VT_KB_BYTES = NV * 128;           // one 64-key k-block of V^T: 48 rows x 128 B (TMA fills rows 0..31)
constexpr int VT_BYTES = 2 * VT_KB_BYTES;       // 12288
constexpr int OFF_K = 2 * Q_BYTES, OFF_VT = OFF_K + STAGES * K_BYTES;
constexpr int OFF_BAR = OFF_VT + STAGES * VT_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 384;   // 3 warpgroups: softmax A, softmax B, {TMA, MMA, 2 idle warps}
constexpr uint32_t TMEM_COLS = 512, TMEM_P = 256, TMEM_P_STRIDE = 64, TMEM_O = 384, TMEM_O_STRIDE = 64;
constexpr int TMA_VT_BYTES = 2 * DH * 128;      // bytes the two V^T TMA boxes deliver per stage
}  // namespace attn

struct AttnBars {
  uint64_t q_full;
  uint64_t s_full[2], s_free[2], p_ready[2], o_full[2];
  uint64_t kv_full[attn::STAGES], kv_empty[attn::STAGES];
  uint32_t tmem_base, pad;
};

#ifdef SVOL_ATTN_TRACE
// Debug build only (-DSVOL_ATTN_TRACE): CTA (0,0,0) records clock64() at phase boundaries of every key tile.
__device__ long long g_attn_trace[4][64][8];
#define SVOL_TR(role, j, slot)                                                             \
  do {                                                                                     \
    if (trace_on && (j) < 64) {                                                            \
      long long c_;                                                                        \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)::"memory");                          \
      g_attn_trace[role][j][slot] = c_;                                                    \
    }                                                                                      \
  } while (0)
#else
#define SVOL_TR(role, j, slot) do {} while (0)
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld_32x32b_x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

// registers -> TMEM: this warp's 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T : A is a 128-lane x (K/2)-column block of packed bf16 pairs; ONE thread issues.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Row maximum of the 128 scores a thread holds; kMasked additionally overwrites invalid keys with -inf.
template <bool kMasked>
__device__ __forceinline__ float tile_row_max(uint32_t (&s)[attn::BKV], const uint32_t (&words)[attn::BKV / 32]) {
  if (kMasked) {
#pragma unroll
    for (int c = 0; c < attn::BKV / 32; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (!((words[c] >> i) & 1u)) s[c * 32 + i] = 0xff800000u;   // -inf
  }
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < attn::BKV; i += 4) {
    m0 = fmaxf(m0, __uint_as_float(s[i + 0])); m1 = fmaxf(m1, __uint_as_float(s[i + 1]));
    m2 = fmaxf(m2, __uint_as_float(s[i + 2])); m3 = fmaxf(m3, __uint_as_float(s[i + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

__global__ void __launch_bounds__(attn::THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, const float* __restrict__ key_mask,
                    __nv_bfloat16* __restrict__ out, int H, int Lq, int Lk, int ldo) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + OFF_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ), h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (Lk + BKV - 1) / BKV;
  const int n_q = (q0 + BQ < Lq) ? 2 : 1;          // is the second query tile of this CTA populated?
#ifdef SVOL_ATTN_TRACE
  const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
#endif

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    mbar_init(&bars->q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars->s_full[t], 1);
      mbar_init(&bars->s_free[t], 4);
      mbar_init(&bars->p_ready[t], 4);
      mbar_init(&bars->o_full[t], 1);
    }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  {
    // rows 32..47 of every V^T k-block buffer are constant: row 32 = 1.0 (bf16), rows 33..47 = 0.  TMA only ever
    // writes rows 0..31, so this is done once.  (A row of equal values is invariant under the 128B swizzle.)
    constexpr int kRegions = STAGES * 2, kChunks = (NV - DH) * 128 / 16;      // 16-byte chunks per region
    for (int idx = threadIdx.x; idx < kRegions * kChunks; idx += THREADS) {
      const int region = idx / kChunks, off = idx % kChunks;
      const uint32_t v = off < 8 ? 0x3F803F80u : 0u;
      *reinterpret_cast<uint4*>(smem + OFF_VT + (region >> 1) * VT_BYTES + (region & 1) * VT_KB_BYTES + DH * 128 + off * 16) =
          make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // register re-split: the softmax warpgroups hold a 128-wide score row + accumulators per thread
  // (each role branch starts with its own setmaxnreg so that it dominates the role's code)
  if (warp >= 8) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->q_full, n_q * Q_BYTES);
      for (int t = 0; t < n_q; ++t)
        tma_load_2d(smem + t * Q_BYTES, &tmQ, &bars->q_full, h * DH, b * Lq + q0 + t * BQ);
      const int vrow = (b * H + h) * DH;
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % STAGES;
        mbar_wait(&bars->kv_empty[s], ((j / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->kv_full[s], K_BYTES + TMA_VT_BYTES);
        tma_load_2d(smem + OFF_K + s * K_BYTES, &tmK, &bars->kv_full[s], h * DH, b * Lk + j * BKV);
        tma_load_2d(smem + OFF_VT + s * VT_BYTES, &tmVt, &bars->kv_full[s], j * BKV, vrow);
        tma_load_2d(smem + OFF_VT + s * VT_BYTES + VT_KB_BYTES, &tmVt, &bars->kv_full[s], j * BKV + 64, vrow);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, NV);
      // descriptors differ only in their 14-bit start-address field (bytes >> 4): plain integer adds below
      const uint64_t dQ = make_kmajor_desc<64>(smem_u32(smem));
      const uint64_t dK = make_kmajor_desc<64>(smem_u32(smem + OFF_K));
      const uint64_t dV = make_kmajor_desc<128>(smem_u32(smem + OFF_VT));
      mbar_wait(&bars->q_full, 0);
      for (int j = 0; j <= n_tiles; ++j) {
        if (j < n_tiles) {
          const int s = j % STAGES;
          SVOL_TR(2, j, 0);
          mbar_wait(&bars->kv_full[s], (j / STAGES) & 1);
          SVOL_TR(2, j, 1);
          const uint64_t dKs = dK + static_cast<uint64_t>(s * (K_BYTES >> 4));
          for (int t = 0; t < n_q; ++t) {
            if (j > 0) mbar_wait(&bars->s_free[t], (j - 1) & 1);
            SVOL_TR(2, j, 2 + t);
            tcgen05_fence_after();
            const uint64_t dQt = dQ + static_cast<uint64_t>(t * (Q_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_bf16_ss(tmem_base + t * BKV, dQt + 2 * k, dKs + 2 * k, idesc_s, k != 0);
            umma_commit(&bars->s_full[t]);
          }
        }
        if (j > 0) {
          const int jp = j - 1, sp = jp % STAGES;
          const uint64_t dVs = dV + static_cast<uint64_t>(sp * (VT_BYTES >> 4));
          for (int t = 0; t < n_q; ++t) {
            mbar_wait(&bars->p_ready[t], jp & 1);
            SVOL_TR(2, j, 4 + t);
            tcgen05_fence_after();
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              umma_bf16_ts(tmem_base + TMEM_O + t * TMEM_O_STRIDE, tmem_base + TMEM_P + t * TMEM_P_STRIDE + k * 8,
                           dVs + static_cast<uint64_t>((k >> 2) * (VT_KB_BYTES >> 4) + (k & 3) * 2), idesc_o, k != 0);
            umma_commit(&bars->o_full[t]);
          }
          umma_commit(&bars->kv_empty[sp]);
        }
      }
    }
   }
  } else {
   asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
   if ((warp >> 2) < n_q) {
    // ------------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;                            // query tile of this warpgroup
    const int quarter = warp & 3;                       // TMEM lane quarter == warp % 4
    const int r = quarter * 32 + lane;                  // row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + t * BKV, t_o = t_lane + TMEM_O + t * TMEM_O_STRIDE;
    const uint32_t t_p = t_lane + TMEM_P + t * TMEM_P_STRIDE;
    const float* mrow = key_mask ? key_mask + static_cast<size_t>(b) * Lk : nullptr;
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 0.f;
    float acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.f;

    // Stagger: warpgroup B starts half a period late (after A's first ex2 section), so that on every SM
    // sub-partition one warp's ex2 section runs under the other's load / max / fold / store phases.  Both
    // warpgroups have the same period, so the offset persists without further synchronisation.
    if (n_q == 2 && t == 1) asm volatile("bar.sync 1, 256;" ::: "memory");

    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = j * BKV;
#ifdef SVOL_ATTN_TRACE
      const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && quarter == 0;
#endif
      SVOL_TR(t, j, 0);
      mbar_wait(&bars->s_full[t], j & 1);
      SVOL_TR(t, j, 1);
      tcgen05_fence_after();
      uint32_t s[BKV];
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32b_x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      SVOL_TR(t, j, 2);
      // scores are in registers: hand the TMEM buffer back so the next QK^T can start now
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free[t]);

      // validity of this tile's 128 keys as four 32-bit words (ragged tail and key_padding_mask); the
      // masked variant of the row-max code is a separate instantiation so full tiles pay nothing for it
      float mx;
      bool masked = false;
      uint32_t words[BKV / 32];
      if (mrow != nullptr || kv0 + BKV > Lk) {
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          const int kv = kv0 + c * 32 + lane;
          const bool ok = kv < Lk && (mrow == nullptr || __ldg(mrow + kv) != 0.f);
          words[c] = __ballot_sync(0xffffffffu, ok);
          masked |= words[c] != 0xffffffffu;
        }
      }
      if (masked) mx = tile_row_max<true>(s, words);
      else mx = tile_row_max<false>(s, words);
      const float m_new = fmaxf(m_run, mx);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = ex2_approx(m_run - m_use);
      m_run = m_new;
      SVOL_TR(t, j, 3);

      // one MUFU ex2 per probability (ex2.approx.ftz.bf16x2 was tried: on sm_100 it is issued as two
      // MUFU.EX2.BF16 ops plus a PRMT, so it saves nothing and only costs precision); the subtraction
      // of the row max is a packed FADD2
      const float2 neg_m = make_float2(-m_use, -m_use);
#pragma unroll
      for (int i = 0; i < BKV; i += 2) {
        const float2 x = __fadd2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), neg_m);
        s[i >> 1] = pack_bf16x2(ex2_approx(x.x), ex2_approx(x.y));
      }
      SVOL_TR(t, j, 5);
      if (j == 0 && n_q == 2 && t == 0)   // (register inputs keep the ex2 section in front of the arrive)
        asm volatile("bar.arrive 1, 256;" ::"r"(s[15]), "r"(s[31]), "r"(s[47]), "r"(s[63]) : "memory");

      if (j > 0) {
        // fold in the previous tile's P V (this also guarantees the MMA is done reading the P columns);
        // column 32 of the product is that tile's row sum of the SAME bf16 probabilities (ones row of V^T)
        mbar_wait(&bars->o_full[t], (j - 1) & 1);
        SVOL_TR(t, j, 6);
        tcgen05_fence_after();
        uint32_t o[32], rsum;
        tmem_ld_32x32b_x32(t_o, o);
        tmem_ld_32x32b_x1(t_o + DH, rsum);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < DH; ++i) acc[i] = acc[i] * alpha_pending + __uint_as_float(o[i]);
        l_run = l_run * alpha_pending + __uint_as_float(rsum);
      }
      alpha_pending = alpha;
      // P tile -> tensor memory: lane = query row, 64 columns of packed bf16 pairs (the A operand of P V)
      tmem_st_32x32b_x32(t_p, &s[0]);
      tmem_st_32x32b_x32(t_p + 32, &s[32]);
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_ready[t]);
      SVOL_TR(t, j, 7);
    }
    // last tile
    mbar_wait(&bars->o_full[t], (n_tiles - 1) & 1);
    tcgen05_fence_after();
    {
      uint32_t o[32], rsum;
      tmem_ld_32x32b_x32(t_o, o);
      tmem_ld_32x32b_x1(t_o + DH, rsum);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < DH; ++i) acc[i] = acc[i] * alpha_pending + __uint_as_float(o[i]);
      l_run = l_run * alpha_pending + __uint_as_float(rsum);
    }
    const int q = q0 + t * BQ + r;
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
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) {
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
  dim3 grid((a.Lq + 2 * BQ - 1) / (2 * BQ), a.H, a.B);
  attention_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmVt, a.key_mask,
                                                            reinterpret_cast<__nv_bfloat16*>(a.out), a.H, a.Lq, a.Lk, a.ldo);
  return svol_check_launch("attention_tc");
}

}  // namespace svol

#ifdef SVOL_ATTN_TRACE
extern "C" int svol_debug_attn_trace(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, svol::g_attn_trace, sizeof(svol::g_attn_trace)));
}
#endif
