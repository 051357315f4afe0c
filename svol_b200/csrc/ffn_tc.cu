// Fused transformer FFN block on tcgen05 / TMEM / TMA (sm_100a):
//
//   y = LayerNorm( x + fc2( GELU_erf( fc1(x) ) ) )          (+ optional second output y + pos)
//
// replaces MLP.forward + residual + norm3 / norm6 of the reference
// (lib/modeling/cross_modal_transformer.py:142-143,157-158,163-179): there it is two cuBLAS GEMMs, a GELU
// kernel, an add and a LayerNorm, with the [tokens, 2048] hidden activation written to and re-read from HBM
// (411 MB per layer at the headline config).  Here the hidden activation never leaves the SM: a CTA owns a
// 128-token tile, walks the 2048 hidden units in 8 chunks of 256 and, per chunk,
//     H_c   = GELU(x W1_c^T + b1_c)      MMA 1: M128 N256 K256  -> fp32 in TMEM -> epilogue -> bf16 in shared memory
//     O    += H_c W2_c^T                 MMA 2: M128 N256 K256, accumulated in TMEM across the 8 chunks
// and finishes with +b2, +x (the residual is the x tile that is already in shared memory), LayerNorm, bf16 stores
// through TMA.  Tensor memory: H accumulator [0,256), O accumulator [256,512).
//
// Warp roles (384 threads = 3 warpgroups, one CTA per SM, persistent over token tiles):
//   warp 0      weight producer: streams W1 / W2 as [256 x 32] bf16 blocks (16 KB, 64B swizzle) through a 5-stage ring
//               in exactly the order the MMA warp consumes them: W1_0, then (W1_{c+1}, W2_c) for c = 0..7
//   warp 1      MMA issuer
//   warp 2      x-tile producer (4 TMA boxes of 128 x 64, 128B swizzle) -- separate from warp 0 so that weight
//               prefetch for the next tile never waits for the previous tile's epilogue
//   warp 3      idle
//   warps 4-11  epilogue: thread = (token row, 128-column half).  Chunk epilogue: TMEM -> registers (accumulator
//               released at once, so MMA 1 of the next chunk overlaps the GELU), +b1, GELU, bf16, swizzled store
//               into the H operand buffer.  Tile epilogue: +b2, +x, LayerNorm (two-pass, halves exchanged through
//               shared memory), bf16 tile staged in the (now free) H / x buffers and written with TMA stores.
// GELU is the exact-erf form evaluated branch-free as  relu(t) -+ 0.5 t * 2^(|t| q(|t|))  with q a degree-4 minimax
// polynomial of log2(erfc(|t|/sqrt 2))/|t| (max |error| 1.2e-6, far below the bf16 rounding of the hidden units):
// 8 FMA-pipe, 3.5 ALU-pipe and 1 MUFU instruction per element.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace ffn {
constexpr int BM = 128, D = 256, CH = 256, BKX = 64, BKW = 32;
constexpr int XKB_BYTES = BM * BKX * 2;                 // one 64-wide k-block of the x / H tile: 16 KB
constexpr int X_BYTES = (D / BKX) * XKB_BYTES;          // 64 KB
constexpr int H_BYTES = (CH / BKX) * XKB_BYTES;         // 64 KB
constexpr int W_STAGE_BYTES = 256 * BKW * 2;            // 16 KB
constexpr int W_STAGES = 5;
constexpr int STAGES_PER_GEMM = 256 / BKW;              // 8 weight blocks per MMA 1 / MMA 2 of one chunk
constexpr int MAX_FF = 2048;
constexpr int OFF_X = 0, OFF_H = OFF_X + X_BYTES, OFF_W = OFF_H + H_BYTES;
constexpr int OFF_B1 = OFF_W + W_STAGES * W_STAGE_BYTES;       // fp32 [MAX_FF]
constexpr int OFF_P2 = OFF_B1 + MAX_FF * 4;                    // fp32 b2, ln_w, ln_b [3][256]
constexpr int OFF_LN = OFF_P2 + 3 * D * 4;                     // float2 [4][128] LayerNorm exchange: (sum, M2) per column group
constexpr int OFF_IDT = OFF_LN + 8 * BM * 4;                   // fp32 [128] 1 / dim_t of the sine positional encoding
constexpr int OFF_BAR = OFF_IDT + (D / 2) * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
// Epilogue warps: 8 (two column groups of 128 per row, 384 threads) or 16 (four groups of 64, 640 threads) -- template
// parameter kEW of the kernel.  A warp may only read its own 32-lane quarter of tensor memory, so the groups split columns.
constexpr int FIRST_EPI_WARP = 4;
constexpr uint32_t TMEM_H = 0, TMEM_O = 256;
}  // namespace ffn

struct FfnBars {
  uint64_t w_full[ffn::W_STAGES], w_empty[ffn::W_STAGES];
  uint64_t x_full, x_free;
  uint64_t hacc_full, hacc_free;      // H accumulator (TMEM) written by MMA 1 / read out by the epilogue
  uint64_t h_ready, h_free;           // H operand (shared memory) written by the epilogue / consumed by MMA 2
  uint64_t oacc_full, oacc_free;      // O accumulator complete / read out
  uint32_t tmem_base, pad;
};
static_assert(sizeof(FfnBars) <= 256, "barrier block");

struct FfnParams {
  const float* b1;
  const float* b2;
  const float* ln_w;
  const float* ln_b;
  const float* pos;         // fp32 [*, ld_pos] or nullptr
  const float* pos_theta;   // fp32 [M] angles of the sine positional encoding, or nullptr (then `pos` is a table)
  int ld_pos, pos_row_mod;
  float ln_eps;
  int M, FF;
  int has_out_pos;
};

#ifdef SVOL_FFN_TRACE
// Debug build only (-DSVOL_FFN_TRACE): CTA 0 records clock64() per chunk of its first tiles.
// role 0: epilogue warp 4 (slots: top, hacc_full, acc->reg, gelu done, h_free, stored); role 1: MMA issuer
// (slots: mma1 start, mma1 issued, h_ready, mma2 issued)
__device__ long long g_ffn_trace[3][64][8];   // [2]: tile epilogue of the first epilogue warp, per tile
#define SVOL_FTR(role, idx, slot)                                                          \
  do {                                                                                     \
    if (ftrace_on && (idx) < 64) {                                                         \
      long long c_;                                                                        \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)::"memory");                          \
      g_ffn_trace[role][idx][slot] = c_;                                                   \
    }                                                                                      \
  } while (0)
#else
#define SVOL_FTR(role, idx, slot) do {} while (0)
#endif

// erf-form GELU of two values (t = accumulator + bias).  For either sign of t
//     gelu(t) = relu(t) - |t|/2 * erfc(|t| / sqrt 2)
// (t > 0: t - t/2 erfc; t < 0: t/2 erfc(-t/sqrt 2) ... = -|t|/2 erfc(|t|/sqrt 2)), so no sign handling is needed.  With
// n = -min(|t|, 4 sqrt 2) (beyond the clamp the correction is below 5e-8) and erfc(a/sqrt 2)/2 = 2^(n Q(n) - 1), Q a degree-3
// minimax fit of -log2(erfc(a/sqrt 2))/a weighted by the GELU's sensitivity (max |gelu error| 8.6e-6, far below the bf16
// rounding of the hidden activation this feeds):  gelu = fma(n, 2^fma(n, Q(n), -1), relu(t)).
// Per PAIR of values: 6 packed f32x2 FMA-pipe instructions (bias add, 3 Horner steps, exponent, final), 4 FMNMX, 2 MUFU.EX2 --
// the chunk epilogue is bound by the FP32 / ALU pipes of its sub-partition (phase trace: ~4 k clk per chunk with 8 OR 16
// epilogue warps, against 4.8 k of MMA), so every instruction removed here shortens the chunk period.  (Round 1 used a
// degree-4 fit with an explicit sign flip: 8 packed + 6 ALU instructions per pair.)
__device__ __forceinline__ float2 gelu_erf_q3_x2(float2 t) {
  constexpr float kClamp = 5.65685424949238f;                   // |t| clamped at z = |t| / sqrt 2 = 4
  const float2 n = make_float2(fmaxf(-fabsf(t.x), -kClamp), fmaxf(-fabsf(t.y), -kClamp));
  float2 q = __ffma2_rn(make_float2(0.004160920158f, 0.004160920158f), n, make_float2(0.04573254287f, 0.04573254287f));
  q = __ffma2_rn(q, n, make_float2(-0.4649338424f, -0.4649338424f));
  q = __ffma2_rn(q, n, make_float2(1.149565816f, 1.149565816f));
  const float2 u = __ffma2_rn(n, q, make_float2(-1.0f, -1.0f));
  float ex, ey;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(u.x));       // erfc(|t| / sqrt 2) / 2
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(u.y));
  return __ffma2_rn(n, make_float2(ex, ey), make_float2(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f)));
}

// kMC: the kernel runs as clusters of two CTAs that work on different token tiles but consume the SAME weight stream.
// Each CTA loads half of every [256 x 32] weight block (128 rows) and multicasts it into both CTAs' rings, so the L2 ->
// SM weight traffic per SM halves (the non-multicast kernel is bound by exactly that feed: ~43 B/clk/SM delivered against
// the 64 B/clk/SM the tensor pipe could consume).  A ring slot is refilled only after BOTH CTAs' MMAs released it
// (tcgen05.commit multicast onto both w_empty barriers).
template <bool kMC, int kEW>
__global__ void __launch_bounds__((ffn::FIRST_EPI_WARP + kEW) * 32, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
              const __grid_constant__ CUtensorMap tmOutPos, const FfnParams p) {
  using namespace ffn;
  constexpr int EPI_WARPS = kEW, THREADS = (FIRST_EPI_WARP + kEW) * 32;
  constexpr int NCG = kEW / 4;                 // column groups (threads per token row)
  constexpr int COLS = D / NCG;                // columns per epilogue thread
  constexpr int KBT = COLS / BKX;              // 64-wide k-blocks of the operand tile a thread owns
  static_assert(kEW == 8 || kEW == 16, "8 or 16 epilogue warps");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  FfnBars* bars = reinterpret_cast<FfnBars*>(smem + OFF_BAR);
  float* s_b1 = reinterpret_cast<float*>(smem + OFF_B1);
  float* s_p2 = reinterpret_cast<float*>(smem + OFF_P2);
  float* ln_x = reinterpret_cast<float*>(smem + OFF_LN);
  float* s_idt = reinterpret_cast<float*>(smem + OFF_IDT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef SVOL_FFN_TRACE
  const bool ftrace_on = blockIdx.x == 0 && lane == 0 && (warp == ffn::FIRST_EPI_WARP || warp == 1);
#endif
  const int m_blocks = (p.M + BM - 1) / BM;
  const int n_chunks = p.FF / CH;
  // tile schedule: group g (a cluster, or a single CTA) takes tile groups g, g + n_groups, ...; inside a group the CTA
  // of rank r takes tile 2 * group + r.  Both CTAs of a cluster run the same number of tiles (the weight stream is shared);
  // a tile index >= m_blocks is a phantom tile: its loads read zeros, its stores are clipped by the tensor maps.
  constexpr int kStride = kMC ? 2 : 1;
  const int rank = kMC ? static_cast<int>(cluster_ctarank()) : 0;
  const int group = static_cast<int>(blockIdx.x) / kStride, n_groups = static_cast<int>(gridDim.x) / kStride;
  const int tile_groups = (m_blocks + kStride - 1) / kStride;
  const int my_tiles = tile_groups > group ? (tile_groups - group + n_groups - 1) / n_groups : 0;
  auto tile_of = [&](int it) { return (group + it * n_groups) * kStride + rank; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmOut);
    if (p.has_out_pos) tma_prefetch_desc(&tmOutPos);
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(&bars->w_full[s], 1); mbar_init(&bars->w_empty[s], kMC ? 2 : 1); }
    mbar_init(&bars->x_full, 1);    mbar_init(&bars->x_free, 1);
    mbar_init(&bars->hacc_full, 1); mbar_init(&bars->hacc_free, EPI_WARPS);
    mbar_init(&bars->h_ready, EPI_WARPS); mbar_init(&bars->h_free, 1);
    mbar_init(&bars->oacc_full, 1); mbar_init(&bars->oacc_free, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
  // parameters used by every tile -> shared memory (read back as broadcasts)
  for (int i = threadIdx.x; i < p.FF; i += THREADS) s_b1[i] = __ldg(p.b1 + i);
  for (int i = threadIdx.x; i < D; i += THREADS) {
    s_p2[i] = __ldg(p.b2 + i);
    s_p2[D + i] = __ldg(p.ln_w + i);
    s_p2[2 * D + i] = __ldg(p.ln_b + i);
  }
  if (p.pos_theta != nullptr)       // 1 / dim_t[2k] = 10000^(-2k/d)  (position_encoding.py:64-65)
    for (int i = threadIdx.x; i < D / 2; i += THREADS)
      s_idt[i] = 1.0f / powf(10000.f, __fdiv_rn(__fmul_rn(2.f, static_cast<float>(i)), static_cast<float>(D)));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (kMC) cluster_sync_all();       // the peer's barriers are initialised before any multicast traffic targets them
  // the tensor-memory base is re-read inside each role (opaque to common-subexpression elimination): as a value of the
  // common prologue it was spilled to local memory and reloaded before every MMA in the issuer's 88-register budget
  auto tmem_base_of = [](const FfnBars* b) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&b->tmem_base)) : "memory"); return v; };

  // Register split of the 384 x 168 launch allocation: 88 for the producer / MMA-issuer warpgroup, 208 for the epilogue
  // warps (128 x 88 + 256 x 208 = 384 x 168).  With 40 / 232 the MMA issuer's descriptors, phases and chunk counters did
  // not fit: ptxas spilled 400 bytes, 130 local-memory instructions INSIDE the issue loop (ncu: 171 k local loads per
  // launch) -- latency on the one thread that feeds the tensor pipe.  88 / 208: 0 bytes.
  if (warp < FIRST_EPI_WARP) {
    // (16 epilogue warps: 640 x 96 at launch = 128 x 64 + 512 x 104)
    if (kEW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ------------------------------------------------------------------ weight producer
      if (elect_one()) {
        int stage = 0; uint32_t phase = 0;
        auto load_blocks = [&](const CUtensorMap* tm, int k0, int n0) {
          for (int kb = 0; kb < STAGES_PER_GEMM; ++kb) {
            mbar_wait(&bars->w_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bars->w_full[stage], W_STAGE_BYTES);
            if (kMC)      // this CTA's 128 rows of the block, delivered to both CTAs (the peer sends the other 128)
              tma_load_2d_multicast(smem + OFF_W + stage * W_STAGE_BYTES + rank * (W_STAGE_BYTES / 2), tm, &bars->w_full[stage],
                                    k0 + kb * BKW, n0 + rank * 128, 0x3);
            else
              tma_load_2d(smem + OFF_W + stage * W_STAGE_BYTES, tm, &bars->w_full[stage], k0 + kb * BKW, n0);
            if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
          }
        };
        for (int it = 0; it < my_tiles; ++it) {
          load_blocks(&tmW1, 0, 0);                                   // W1 chunk 0
          for (int c = 0; c < n_chunks; ++c) {
            if (c + 1 < n_chunks) load_blocks(&tmW1, 0, (c + 1) * CH);  // W1 chunk c+1: rows = hidden units
            load_blocks(&tmW2, c * CH, 0);                            // W2 chunk c: k = hidden units
          }
        }
      }
    } else if (warp == 2) {
      // ------------------------------------------------------------------ x-tile producer
      if (elect_one()) {
        // (programmatic dependent launch: the weight producer above already fills its ring under the tail of the kernel that
        // produces x; the token tiles wait for that kernel to complete)
        griddep_wait();
        for (int it = 0; it < my_tiles; ++it) {
          const int m_blk = tile_of(it);
          if (it == my_tiles - 1) griddep_launch_dependents();   // last tile: the next kernel may take the SMs that free up
          mbar_wait(&bars->x_free, (it & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->x_full, X_BYTES);
          for (int kb = 0; kb < D / BKX; ++kb)
            tma_load_2d(smem + OFF_X + kb * XKB_BYTES, &tmX, &bars->x_full, kb * BKX, m_blk * BM);
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(BM, 256);
        const uint64_t dX = make_kmajor_desc<128>(smem_u32(smem + OFF_X));
        const uint64_t dH = make_kmajor_desc<128>(smem_u32(smem + OFF_H));
        const uint64_t dW = make_kmajor_desc<64>(smem_u32(smem + OFF_W));
        int stage = 0; uint32_t phase = 0;
        // one 128 x 256 x 256 GEMM: A = 4 k-blocks of 64 in shared memory, B = 8 streamed 32-wide weight blocks
        auto gemm256 = [&](uint32_t d_tmem, uint64_t dA, bool fresh) {
          for (int s = 0; s < STAGES_PER_GEMM; ++s) {
            mbar_wait(&bars->w_full[stage], phase);
            tcgen05_fence_after();
            const uint64_t a = dA + static_cast<uint64_t>((s >> 1) * (XKB_BYTES >> 4) + (s & 1) * 4);
            const uint64_t b = dW + static_cast<uint64_t>(stage * (W_STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BKW / 16; ++k) umma_bf16_ss(d_tmem, a + 2 * k, b + 2 * k, idesc, (fresh && s == 0 && k == 0) ? 0u : 1u);
            if (kMC) umma_commit_multicast(&bars->w_empty[stage], 0x3);      // the slot is shared: release it in both CTAs
            else umma_commit(&bars->w_empty[stage]);
            if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
          }
        };
        uint32_t n_h = 0;          // chunks whose MMA 1 has been issued (H accumulator uses)
        uint32_t n_o = 0;          // chunks whose MMA 2 has been issued (H operand uses)
        const uint32_t tmem_base = tmem_base_of(bars);
        for (int it = 0; it < my_tiles; ++it) {
          mbar_wait(&bars->x_full, it & 1);
          auto mma1 = [&]() {
            mbar_wait(&bars->hacc_free, (n_h & 1) ^ 1);       // epilogue has read out the previous chunk's accumulator
            SVOL_FTR(1, n_h, 0);
            tcgen05_fence_after();
            gemm256(tmem_base + TMEM_H, dX, true);
            umma_commit(&bars->hacc_full);
            SVOL_FTR(1, n_h, 1);
            ++n_h;
          };
          mma1();
          for (int c = 0; c < n_chunks; ++c) {
            if (c + 1 < n_chunks) mma1();
            mbar_wait(&bars->h_ready, n_o & 1);                // H_c is in shared memory
            if (c == 0) mbar_wait(&bars->oacc_free, (it & 1) ^ 1);   // previous tile's O has been read out
            SVOL_FTR(1, n_o, 2);
            tcgen05_fence_after();
            gemm256(tmem_base + TMEM_O, dH, c == 0);
            umma_commit(&bars->h_free);
            SVOL_FTR(1, n_o, 3);
            ++n_o;
          }
          umma_commit(&bars->oacc_full);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    if (kEW == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    griddep_wait();
    const int quarter = warp & 3;                             // TMEM lane quarter this warp may read
    const int half = (warp - FIRST_EPI_WARP) >> 2;            // column group: which COLS of the 256 columns
    const int r = quarter * 32 + lane;                        // row inside the tile
    const uint32_t t_lane = tmem_base_of(bars) + (static_cast<uint32_t>(quarter * 32) << 16);
    // this thread's row inside a [4 k-blocks][128 rows][128 B] operand tile: 16-byte chunk j of k-block kb lives at
    // kb * 16 KB + r * 128 + ((j ^ (r & 7)) << 4); the thread owns k-blocks KBT*half .. KBT*half + KBT - 1
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    uint32_t n_h = 0, n_hs = 0;     // H accumulator read-outs / H operand stores so far
    auto store_tile_half = [&](uint8_t* buf, const float (&v)[COLS]) {
#pragma unroll
      for (int kb = 0; kb < KBT; ++kb)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float* vv = &v[kb * 64 + j * 8];
          const uint4 q = make_uint4(pack_bf16x2(vv[0], vv[1]), pack_bf16x2(vv[2], vv[3]), pack_bf16x2(vv[4], vv[5]),
                                     pack_bf16x2(vv[6], vv[7]));
          *reinterpret_cast<uint4*>(buf + (KBT * half + kb) * XKB_BYTES + row_off + ((static_cast<uint32_t>(j) ^ swz) << 4)) = q;
        }
    };
    auto load_acc = [&](uint32_t col0, float (&v)[COLS]) {
      uint32_t raw[COLS / 32][32];
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c) tmem_ld_32x32b_x32(t_lane + col0 + half * COLS + c * 32, raw[c]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[c * 32 + i] = __uint_as_float(raw[c][i]);
    };

    for (int it = 0; it < my_tiles; ++it) {
      const int m_blk = tile_of(it);
      const int row = m_blk * BM + r;
      for (int c = 0; c < n_chunks; ++c) {
        // ---- chunk epilogue: H_c = GELU(acc + b1_c) -> bf16 -> shared memory operand of MMA 2
        SVOL_FTR(0, n_h, 0);
        mbar_wait(&bars->hacc_full, n_h & 1);
        SVOL_FTR(0, n_h, 1);
        tcgen05_fence_after();
        float v[COLS];
        load_acc(TMEM_H, v);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->hacc_free);
        SVOL_FTR(0, n_h, 2);
        ++n_h;
        const uint32_t bp = smem_u32(s_b1 + c * CH + half * COLS);
#pragma unroll
        for (int i = 0; i < COLS / 4; ++i) {
          const float4 b = lds_f4(bp + i * 16);
          const float2 g0 = gelu_erf_q3_x2(__fadd2_rn(make_float2(v[4 * i + 0], v[4 * i + 1]), make_float2(b.x, b.y)));
          const float2 g1 = gelu_erf_q3_x2(__fadd2_rn(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w)));
          v[4 * i + 0] = g0.x; v[4 * i + 1] = g0.y; v[4 * i + 2] = g1.x; v[4 * i + 3] = g1.y;
        }
        // the H buffer is free once MMA 2 of the previous chunk has consumed it; at the start of a tile it also served as
        // the staging buffer of the previous tile's output store (x_free is signalled only after those stores are read)
        SVOL_FTR(0, n_hs, 3);
        if (n_hs > 0) mbar_wait(&bars->h_free, (n_hs - 1) & 1);
        if (c == 0 && it > 0) {
          // the previous tile's last output store is still being read out of this buffer by the TMA engine: waited for
          // HERE, after this chunk's GELU, instead of at the end of the tile epilogue (2 k clk that nothing else covered)
          if (warp == FIRST_EPI_WARP && lane == 0) tma_store_wait_read<0>();
          asm volatile("bar.sync 5, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        }
        SVOL_FTR(0, n_hs, 4);
        store_tile_half(smem + OFF_H, v);
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->h_ready);
        SVOL_FTR(0, n_hs, 5);
        ++n_hs;
      }

      // ---- tile epilogue: y = LayerNorm(O + b2 + x)
      mbar_wait(&bars->oacc_full, it & 1);
      SVOL_FTR(2, it, 0);
      tcgen05_fence_after();
      float v[COLS];
      load_acc(TMEM_O, v);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->oacc_free);
      SVOL_FTR(2, it, 1);
      {
        const uint32_t bp = smem_u32(s_p2 + half * COLS);
#pragma unroll
        for (int kb = 0; kb < KBT; ++kb)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 xq = *reinterpret_cast<const uint4*>(smem + OFF_X + (KBT * half + kb) * XKB_BYTES + row_off +
                                                             ((static_cast<uint32_t>(j) ^ swz) << 4));
            const float4 b0 = lds_f4(bp + (kb * 16 + j * 2) * 16), b1 = lds_f4(bp + (kb * 16 + j * 2 + 1) * 16);
            float* vv = &v[kb * 64 + j * 8];
            vv[0] += b0.x + bf16_lo(xq.x); vv[1] += b0.y + bf16_hi(xq.x); vv[2] += b0.z + bf16_lo(xq.y); vv[3] += b0.w + bf16_hi(xq.y);
            vv[4] += b1.x + bf16_lo(xq.z); vv[5] += b1.y + bf16_hi(xq.z); vv[6] += b1.z + bf16_lo(xq.w); vv[7] += b1.w + bf16_hi(xq.w);
          }
      }
      // every epilogue warp has read its residual: the x buffer can be refilled with the next tile while this one finishes
      asm volatile("bar.sync 5, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (warp == FIRST_EPI_WARP && lane == 0) mbar_arrive(&bars->x_free);
      SVOL_FTR(2, it, 2);
      // LayerNorm over the 256-wide row.  Each of the NCG threads that share a row reduces its own columns to (sum, M2 about
      // its own mean) in registers; ONE exchange through shared memory, then the partials are combined exactly
      // (Chan et al.: M2 = sum_i M2_i + n_i (mean_i - mean)^2) -- two named barriers fewer than a two-pass exchange.
      // (four independent partial sums each: one 128-long dependent chain of adds is ~500 clk of pure latency)
      float sp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < COLS; ++i) sp[i & 3] += v[i];
      const float s = (sp[0] + sp[1]) + (sp[2] + sp[3]);
      const float mean_own = s * (1.0f / COLS);
      float mp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < COLS; ++i) { const float d = v[i] - mean_own; mp[i & 3] = fmaf(d, d, mp[i & 3]); }
      const float m2 = (mp[0] + mp[1]) + (mp[2] + mp[3]);
      reinterpret_cast<float2*>(ln_x)[half * BM + r] = make_float2(s, m2);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(NCG * 32) : "memory");
      float2 part[NCG];
      float tot = 0.f;
#pragma unroll
      for (int c = 0; c < NCG; ++c) { part[c] = reinterpret_cast<const float2*>(ln_x)[c * BM + r]; tot += part[c].x; }
      const float mean = tot * (1.0f / D);
      float m2_tot = 0.f;
#pragma unroll
      for (int c = 0; c < NCG; ++c) { const float dm = part[c].x * (1.0f / COLS) - mean; m2_tot += part[c].y + COLS * dm * dm; }
      const float var = m2_tot * (1.0f / D);
      const float rstd = rsqrtf(var + p.ln_eps);
      {
        const uint32_t gp = smem_u32(s_p2 + D + half * COLS), bp = smem_u32(s_p2 + 2 * D + half * COLS);
#pragma unroll
        for (int i = 0; i < COLS / 4; ++i) {
          const float4 g = lds_f4(gp + i * 16), b = lds_f4(bp + i * 16);
          v[4 * i + 0] = (v[4 * i + 0] - mean) * rstd * g.x + b.x;
          v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * g.y + b.y;
          v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * g.z + b.z;
          v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * g.w + b.w;
        }
      }
      // y -> H buffer (free: oacc_full implies the last MMA 2 of this tile has completed) -> TMA store
      SVOL_FTR(2, it, 3);
      store_tile_half(smem + OFF_H, v);
      fence_proxy_async_smem();
      asm volatile("bar.sync 5, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (warp == FIRST_EPI_WARP && lane == 0) {
        for (int kb = 0; kb < D / BKX; ++kb) tma_store_2d(&tmOut, smem + OFF_H + kb * XKB_BYTES, kb * BKX, m_blk * BM);
        tma_store_commit();
      }
      SVOL_FTR(2, it, 4);
      if (p.has_out_pos) {
        // second output y + pos, staged in the same buffer once the first store has been read out of it
        if (p.pos_theta != nullptr) {
          // sine positional encoding evaluated in place (position_encoding.py:62-71): columns (2k, 2k+1) =
          // (sin, cos)(theta_row / dim_t[2k]).  theta is in [0, 2 pi]: folded to [-pi, pi] for the MUFU sin / cos.
          const float theta = row < p.M ? __ldg(p.pos_theta + row) : 0.f;
          const uint32_t ip = smem_u32(s_idt + half * (COLS / 2));
#pragma unroll
          for (int i = 0; i < COLS / 8; ++i) {
            const float4 w = lds_f4(ip + i * 16);
            const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = theta * ws[e];
              a = a > 3.14159265358979f ? a - 6.28318530717959f : a;
              v[8 * i + 2 * e] += __sinf(a);
              v[8 * i + 2 * e + 1] += __cosf(a);
            }
          }
        } else if (row < p.M) {
          const int prow = p.pos_row_mod > 0 ? row % p.pos_row_mod : row;
          const float4* pp = reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(prow) * p.ld_pos + half * COLS);
#pragma unroll
          for (int i = 0; i < COLS / 4; ++i) {
            const float4 q = __ldg(pp + i);
            v[4 * i + 0] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
          }
        }
        if (warp == FIRST_EPI_WARP && lane == 0) tma_store_wait_read<0>();
        SVOL_FTR(2, it, 5);
        asm volatile("bar.sync 5, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        store_tile_half(smem + OFF_H, v);
        fence_proxy_async_smem();
        asm volatile("bar.sync 5, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        if (warp == FIRST_EPI_WARP && lane == 0) {
          for (int kb = 0; kb < D / BKX; ++kb) tma_store_2d(&tmOutPos, smem + OFF_H + kb * XKB_BYTES, kb * BKX, m_blk * BM);
          tma_store_commit();
        }
        SVOL_FTR(2, it, 6);
      }
      SVOL_FTR(2, it, 7);
    }
    if (warp == FIRST_EPI_WARP && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  if (kMC) cluster_sync_all();       // no CTA exits while its peer can still multicast into it / arrive on its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base_of(bars));
  }
}

int launch_ffn_tc(const svol_ffn_args& a, cudaStream_t stream) {
  using namespace ffn;
  if (a.M <= 0 || a.d != D || a.ff <= 0 || a.ff % CH != 0 || a.ff > MAX_FF)
    return svol_fail(SVOL_ERR_SHAPE, "ffn: need d == 256 and ff a multiple of 256 (<= 2048)");
  if (!a.x || !a.w1 || !a.b1 || !a.w2 || !a.b2 || !a.ln_weight || !a.ln_bias || !a.out)
    return svol_fail(SVOL_ERR_NULL, "ffn: required pointer is NULL");
  if (a.out_pos && !a.pos && !a.pos_theta) return svol_fail(SVOL_ERR_NULL, "ffn: out_pos needs pos or pos_theta");
  CUtensorMap tmX, tmW1, tmW2, tmOut, tmOutPos;
  int rc = make_tensor_map_2d(&tmX, a.x, D, a.M, a.ldx, BKX, BM, 128);
  if (rc) return rc;
  const int m_blocks = (a.M + BM - 1) / BM;
  static const bool mc_enabled = [] { const char* e = getenv("SVOL_FFN_MULTICAST"); return !(e && e[0] == '0'); }();
  // measured (B200, C2): 50176 tokens (392 tiles, persistent, 3 waves) 129.7 -> 122.1 us with multicast; 10240 tokens (80 tiles,
  // one wave) 53.5 -> 55.4 us: the pairing only pays when every SM streams the weights several times
  const bool multicast = mc_enabled && m_blocks > sm_count() && sm_count() >= 2;
  const int w_box_rows = multicast ? 128 : 256;        // multicast: each CTA of the pair loads half of a weight block
  rc = make_tensor_map_2d(&tmW1, a.w1, D, a.ff, a.ldw1, BKW, w_box_rows, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmW2, a.w2, a.ff, D, a.ldw2, BKW, w_box_rows, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmOut, a.out, D, a.M, a.ld_out, BKX, BM, 128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmOutPos, a.out_pos ? a.out_pos : a.out, D, a.M, a.ld_out, BKX, BM, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ffn_tc_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ffn_tc_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ffn_tc_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return svol_fail_cuda(e, "ffn: cudaFuncSetAttribute");
    configured = true;
  }
  const char* env_ew = getenv("SVOL_FFN_EPI_WARPS");            // read per launch (A/B measurements)
  const bool wide = env_ew ? atoi(env_ew) == 16 : false;
  const int threads = (FIRST_EPI_WARP + (wide ? 16 : 8)) * 32;
  FfnParams p;
  p.b1 = a.b1; p.b2 = a.b2; p.ln_w = a.ln_weight; p.ln_b = a.ln_bias; p.pos = a.pos; p.pos_theta = a.pos_theta;
  p.ld_pos = a.ld_pos; p.pos_row_mod = a.pos_row_mod; p.ln_eps = a.ln_eps; p.M = a.M; p.FF = a.ff;
  p.has_out_pos = a.out_pos != nullptr;
  // Launch only as many CTAs as the slowest one needs rounds: 392 token tiles on 148 SMs are three rounds whether 148
  // CTAs (96 of them with three tiles, 52 with two) or 131 (all with three) run them -- the second form finishes at the same
  // time and leaves 17 SMs to the kernels of the other graph branches / the other batch in flight.
  const int rounds = (m_blocks + sm_count() - 1) / sm_count();
  const int ctas_needed = (m_blocks + rounds - 1) / rounds;
  if (multicast) {
    const int pairs = std::min((ctas_needed + 1) / 2, sm_count() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = wide ? cudaLaunchKernelEx(&cfg, ffn_tc_kernel<true, 16>, tmX, tmW1, tmW2, tmOut, tmOutPos, p)
                         : cudaLaunchKernelEx(&cfg, ffn_tc_kernel<true, 8>, tmX, tmW1, tmW2, tmOut, tmOutPos, p);
    if (e != cudaSuccess) return svol_fail_cuda(e, "ffn: cluster launch");
    return svol_check_launch("ffn_tc (2-CTA multicast)");
  }
  const int grid = ctas_needed < sm_count() ? ctas_needed : sm_count();
  cudaError_t e = wide ? launch_kernel_pdl(ffn_tc_kernel<false, 16>, dim3(grid), dim3(threads), SMEM_BYTES, stream, tmX, tmW1, tmW2, tmOut, tmOutPos, p)
                       : launch_kernel_pdl(ffn_tc_kernel<false, 8>, dim3(grid), dim3(threads), SMEM_BYTES, stream, tmX, tmW1, tmW2, tmOut, tmOutPos, p);
  if (e != cudaSuccess) return svol_fail_cuda(e, "ffn: launch");
  return svol_check_launch("ffn_tc");
}

}  // namespace svol

#ifdef SVOL_FFN_TRACE
extern "C" int svol_debug_ffn_trace(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, svol::g_ffn_trace, sizeof(svol::g_ffn_trace)));
}
#endif
