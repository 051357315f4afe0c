// Flash-style multi-head attention core on tcgen05 / TMEM / TMA (sm_100a), head_dim = 32.
//
// Replaces softmax(Q K^T / sqrt(dh) + key_padding_mask) V inside nn.MultiheadAttention at
// lib/modeling/cross_modal_transformer.py:139 (video self-attention, L x L), :147 (query
// self-attention, Q x Q) and :154 (query -> video cross-attention, Q x L, padded keys masked).
// The reference materialises the (B*8, Lq, Lk) score tensor (2.5 GB at the headline config); here
// scores only ever exist as 128 x 128 fp32 tiles in tensor memory.
//
// With head_dim 32 every probability costs 128 tensor FLOPs and one ex2: the binding unit is the MUFU
// pipe (16 ex2/clk/SM, measured: one warp per SM sub-partition already saturates it), not the tensor pipe.
// The kernel is therefore organised around keeping MUFU busy: FOUR independent softmax warpgroups per CTA
// (four warps per sub-partition), so that while one warp waits on a barrier, loads scores, reduces its row
// maximum or stores probabilities, the others issue exponentials.
//
// One CTA per (pair of 128-query tiles, head, sample), one CTA per SM, 640 threads:
//   warps 0..15  softmax warpgroups g = 0..3; g = 2*t + half handles query tile t (0/1) against the
//                `half`-th 64 keys of every 128-key tile -- split-key flash attention: each warpgroup keeps
//                its own running maximum, row sum and output accumulator, merged once at the end, so the
//                warpgroups never synchronise with each other inside the key loop.
//                Per key tile and thread (= one query row, 64 keys): scores TMEM -> registers (the S buffer
//                is released to the MMA warp immediately), row max, ex2 (row-max subtraction as packed
//                FADD2), fp32 row sum, bf16 probabilities -> TENSOR MEMORY (tcgen05.st): the P V MMA reads its
//                A operand from TMEM, so probabilities never touch shared memory and no proxy fence is needed.
//                The output accumulator stays in TMEM across key tiles (MMA accumulate); it is rescaled
//                (TMEM -> registers -> TMEM) only when some row's maximum grew by more than 2^8 since the
//                reference maximum was fixed ("lazy rescaling": probabilities are then at most 256, exact in
//                bf16's exponent range), which after the first tiles practically never happens.
//   warp 16      TMA producer: both Q tiles once, then a 3-stage ring of K tiles (128 x 32, 64B swizzle)
//                and V^T tiles (32 x 128 as two 64-key blocks, 128B swizzle) shared by all warpgroups.
//   warps 17,18  MMA issuers, one per query tile (independent instruction streams, so a late barrier of one
//                tile never delays the other): S_g = Q_t K_half^T (M128 N64 K32, operands in shared memory) and
//                O_g += P_g V_half (M128 N32 K64, A = P in TMEM, B = V^T block in shared memory), issued in the
//                order the staggered warpgroups make them ready: QK_lo(i), PV_hi(i-2), QK_hi(i), PV_lo(i-1).
//   warp 19      idle (setmaxnreg works on whole warpgroups).
// The four warpgroups are fully independent pipelines (own S / P / O columns and barriers).  They are started
// a quarter of a key-tile period apart (g0, g2, g1, g3) so that on every sub-partition the ex2 sections of the
// four resident warps interleave instead of colliding; equal periods keep the offsets.
// TMEM map (512 columns): S_g [64g, +64) P_g [256+32g, +32) O_g [384+32g, +32).
// A CTA whose second query tile is empty (the tail of Lq: the last CTA of every (sample, head) when Lq % 256 is in
// (0, 128]) would leave warpgroups 2, 3 and one MMA issuer idle while the other two walk all key tiles at the latency-bound
// pace of a single chain (2043 clk per key tile, measured).  Such a CTA instead runs its ONE query tile on all four
// warpgroups: "tile slot" t = 0 / 1 takes the even / odd key tiles, each slot with its own issuer, S / P / O columns and
// running maxima, and the four partial results per row are merged at the end.
// Q is pre-scaled by log2(e)/sqrt(dh) when it is produced, so the softmax is a bare ex2.
#include <atomic>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "svol_internal.h"

namespace svol {

namespace attn {
// K / V^T ring depth.  A stage is released by the P V of its tile's upper half, which an issuer reaches two of ITS OWN
// key tiles later; in a single-tile CTA consecutive key tiles alternate between the two issuers, so a stage stays
// occupied for four tiles and a 5-deep ring left no room to prefetch (measured: 2-3 k clk of load latency exposed on
// every key tile).  8 stages = 4 tiles of lag + 4 of prefetch.
constexpr int BQ = 128, BKV = 128, HALF = 64, DH = 32, STAGES = 8;
constexpr int Q_BYTES = BQ * DH * 2;            // 8192 per query tile
constexpr int K_BYTES = BKV * DH * 2;           // 8192
constexpr int VT_KB_BYTES = DH * 128;           // one 64-key block of V^T: 32 rows x 128 B
constexpr int VT_BYTES = 2 * VT_KB_BYTES;       // 8192
constexpr int CMB_STRIDE = 36;                  // floats per row of the merge buffer: m, l, O[32], 2 unused -- nine 16-byte accesses per row,
                                                // conflict-free (a quarter warp's rows start 4 banks apart)
// merge buffer: up to three partial results per row of a 128-row tile (single-tile CTAs), or seven per row of a 64-row tile
// (single-tile CTAs with at most 64 query rows: two key parts per half tile, run_softmax_dup)
constexpr int CMB_BYTES = 7 * 64 * CMB_STRIDE * 4;
constexpr int CMB_BYTES_3 = 3 * BQ * CMB_STRIDE * 4;           // three partials per 128-row tile only (persistent variant)
constexpr int OFF_K = 2 * Q_BYTES, OFF_VT = OFF_K + STAGES * K_BYTES, OFF_CMB = OFF_VT + STAGES * VT_BYTES;
constexpr int KMASK_WORDS = 256;                // key-padding bitmask of one sample: up to 8192 keys
constexpr int OFF_KMASK = OFF_CMB + CMB_BYTES;
constexpr int OFF_BAR = OFF_KMASK + KMASK_WORDS * 4;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr int THREADS = 640;   // 5 warpgroups: 4 x softmax, {TMA, MMA 0, MMA 1, idle}
constexpr uint32_t TMEM_COLS = 512, TMEM_P = 256, TMEM_O = 384;
constexpr float RESCALE_THRESHOLD = 8.0f;       // log2 units
#ifndef SVOL_ATTN_STAGGER
#define SVOL_ATTN_STAGGER 550
#endif
#ifndef SVOL_ATTN_REINIT_MODE
#define SVOL_ATTN_REINIT_MODE 0
#endif
constexpr int STAGGER_CLK = SVOL_ATTN_STAGGER;  // start offset between consecutive warpgroups (~ period / 4)
}  // namespace attn

struct AttnBars {
  uint64_t q_full;
  uint64_t s_full[4], s_free[4];
  uint64_t p_ready[4], o_full[4];
  uint64_t kv_full[attn::STAGES], kv_empty[attn::STAGES];
  uint32_t tmem_base, pad;
  int next_item, cur_item;   // looping CTAs (kLoop): the work item after the current one / the current one
};
static_assert(offsetof(AttnBars, tmem_base) == (17 + 2 * attn::STAGES) * 8, "the 17 + 2 x STAGES barriers are re-initialised as one array");
static_assert(offsetof(AttnBars, s_free) - offsetof(AttnBars, s_full) == 32 && offsetof(AttnBars, p_ready) - offsetof(AttnBars, s_full) == 64 &&
              offsetof(AttnBars, o_full) - offsetof(AttnBars, s_full) == 96, "softmax warps address their four barriers from one register");

#ifdef SVOL_ATTN_TRACE
// Debug build only (-DSVOL_ATTN_TRACE): CTA (0,0,0) records clock64() at phase boundaries of every key tile.
// roles 0..3: softmax warpgroup g (its first warp); 4, 5: MMA issuer of tile 0 / 1.
__device__ long long g_attn_trace[6][64][8];
#define SVOL_TR(role, j, slot)                                                             \
  do {                                                                                     \
    if (trace_on && (j) < 64) {                                                            \
      long long c_;                                                                        \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)::"memory");                          \
      g_attn_trace[role][j][slot] = c_;                                                    \
    }                                                                                      \
  } while (0)
// same, but ordered after the computation of `val` (a float register)
#define SVOL_TR_AFTER(role, j, slot, val)                                                  \
  do {                                                                                     \
    if (trace_on && (j) < 64) {                                                            \
      long long c_;                                                                        \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_), "+f"(val)::"memory");               \
      g_attn_trace[role][j][slot] = c_;                                                    \
    }                                                                                      \
  } while (0)
// per-CTA timeline: {SM id, first instruction, -, last instruction} in globaltimer ns, indexed by the linear block id
// (tools/attn_cta_timeline.py)
__device__ long long g_attn_cta[8192][4];
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory"); return t; }
#else
#define SVOL_TR(role, j, slot) do {} while (0)
#define SVOL_TR_AFTER(role, j, slot, val) do {} while (0)
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max relative error 1.5e-4): x = n + r with
// n = round(x), |r| <= 0.5; 2^r from the polynomial, 2^n by adding n to the exponent field.  EXPERIMENT, off by default:
// with -DSVOL_ATTN_POLY_PER4=1|2 that many of every 4 probabilities bypass the MUFU pipe.  Measured on B200 (video
// self-attention, B=32): 265.7 us (0) / 266.0 us (1) / 304.5 us (2) -- the kernel is not limited by raw MUFU
// throughput but by the queueing of every other MIO-class instruction (tcgen05.ld/st, mbarrier ops) behind the
// exponentials, so moving exponentials to the FMA pipe buys nothing; see DESIGN.md 4.1.
#ifndef SVOL_ATTN_POLY_PER4
#define SVOL_ATTN_POLY_PER4 0
#endif
#ifndef SVOL_ATTN_SPEC_MAX
#define SVOL_ATTN_SPEC_MAX 0        // fold the row maximum of tiles 1.. into the exponential loop (see the softmax loop).
// EXPERIMENT, off by default: measured 303.8 / 82.4 / 33.7 us against 261.7 / 72.7 / 29.9 (video self / cross / query self): with
// all 64 scores, the running maxima and the packed results live at once the loop no longer fits the 112 registers of a
// softmax thread (140 B of spills INSIDE the key loop), which costs far more than the ~360 clk the max phase took.
#endif
#ifndef SVOL_ATTN_POLY_PAIRS
#define SVOL_ATTN_POLY_PAIRS 0      // of every four pairs of probabilities, this many use the packed polynomial (ex2_poly2).
// EXPERIMENT, off by default.  Measured on B200 (video self / cross / query self-attention, us per launch, kernels alone):
// 0: 262.7 / 72.3 / 30.0;  1: 265.7 / 76.8 / 29.7;  2: 293.3 / 82.6 / 31.9 -- a quarter fewer MUFU operations at 4 packed issue
// slots per probability buys nothing: the key-tile period is set by the four warps' serial chains queueing on the shared
// pipes, not by MUFU or issue throughput.
#endif
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);                                   // also maps masked scores (-inf) to 2^-126 ~ 0
  const float t = __fadd_rn(x, 12582912.f);               // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float r = __fsub_rn(x, __fsub_rn(t, 12582912.f));
  float p = fmaf(0.055170271545648575f, r, 0.2426079511642456f);
  p = fmaf(p, r, 0.693260908126831f);
  p = fmaf(p, r, 0.9999282956123352f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// registers <-> TMEM: this warp's 32 lanes x N consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
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

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Two probabilities 2^(s - m) on the FMA pipe with PACKED f32x2 instructions (add.f32x2 / fma.rn.f32x2): Cody-Waite split
// n = round(x), r = x - n in [-0.5, 0.5], degree-3 minimax polynomial for 2^r (max relative error 1.5e-4, an order of magnitude
// below the bf16 rounding of P), 2^n added into the exponent field.  3 packed adds + 3 packed FMAs + 2 integer ops for two
// values: 4 issue slots per probability instead of 9 for the scalar form above -- what makes moving a quarter of the
// exponentials off the MUFU pipe pay (the kernel is MUFU-bound at 8 clk per warp instruction but uses ~60 % of its issue
// slots).  Requires finite s - m > -126 (not used on tiles with masked keys).
//   magic_m = 1.5 * 2^23 - m (both lanes), neg_m = -m.
__device__ __forceinline__ float2 ex2_poly2(float2 sc, float2 magic_m, float2 neg_m) {
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = __fadd2_rn(sc, magic_m);                         // x + magic: round(x) lands in the low mantissa bits
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 x = __fadd2_rn(sc, neg_m);
  const float2 r = __fadd2_rn(x, make_float2(-n.x, -n.y));
  (void)magic;
  float2 p = __ffma2_rn(make_float2(0.055170271545648575f, 0.055170271545648575f), r, make_float2(0.2426079511642456f, 0.2426079511642456f));
  p = __ffma2_rn(p, r, make_float2(0.693260908126831f, 0.693260908126831f));
  p = __ffma2_rn(p, r, make_float2(0.9999282956123352f, 0.9999282956123352f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// Row maximum of the 64 scores a thread holds; kMasked additionally overwrites invalid keys with -inf.
// mbarrier operations on a 32-bit shared-memory ADDRESS (the softmax warps keep one opaque address register per thread:
// given pointers, ptxas re-derived the shared-window conversion and the warp index from special registers -- S2R / S2UR,
// tens of cycles each -- in front of every barrier operation of the latency-bound key loop)
// Explicit shared-memory accesses for the merge buffer and the key-mask words: through the 1024-byte-aligned GENERIC pointer the
// compiler emitted generic LD.E / ST.E (34 scalar stores + 34 scalar loads per row of a partial result, two loads per key tile
// for the mask words) -- the merge of a two-tile CTA took 2.4 k clk (tools/attn_ends_trace.py)
__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// one partial result (m, l, O[32]) of a row <-> merge buffer
__device__ __forceinline__ void cmb_store(uint32_t a, float m, float l, const uint32_t (&o)[attn::DH]) {
  sts_f4(a, m, l, __uint_as_float(o[0]), __uint_as_float(o[1]));
#pragma unroll
  for (int i = 0; i < 7; ++i)
    sts_f4(a + 16 + 16 * i, __uint_as_float(o[2 + 4 * i]), __uint_as_float(o[3 + 4 * i]), __uint_as_float(o[4 + 4 * i]), __uint_as_float(o[5 + 4 * i]));
  sts_f4(a + 128, __uint_as_float(o[30]), __uint_as_float(o[31]), 0.f, 0.f);
}
// acc += 2^(m_part - m_safe) * O_part, l_tot += the same factor * l_part (element order as before: bit-identical results)
__device__ __forceinline__ void cmb_accumulate(uint32_t a, float m_safe, float& l_tot, float (&acc)[attn::DH]) {
  const float4 h = lds_f4(a);
  float ap;
  {
    const float x = h.x - m_safe;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ap) : "f"(x));
  }
  l_tot = fmaf(ap, h.y, l_tot);
  acc[0] = fmaf(ap, h.z, acc[0]);
  acc[1] = fmaf(ap, h.w, acc[1]);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float4 v = lds_f4(a + 16 + 16 * i);
    acc[2 + 4 * i] = fmaf(ap, v.x, acc[2 + 4 * i]);
    acc[3 + 4 * i] = fmaf(ap, v.y, acc[3 + 4 * i]);
    acc[4 + 4 * i] = fmaf(ap, v.z, acc[4 + 4 * i]);
    acc[5 + 4 * i] = fmaf(ap, v.w, acc[5 + 4 * i]);
  }
  const float4 v = lds_f4(a + 128);
  acc[30] = fmaf(ap, v.x, acc[30]);
  acc[31] = fmaf(ap, v.y, acc[31]);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > SVOL_SPIN_LIMIT) __trap();
  }
}

template <bool kMasked>
__device__ __forceinline__ float half_row_max(uint32_t (&s)[attn::HALF], const uint32_t (&words)[2]) {
  if (kMasked) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (!((words[c] >> i) & 1u)) s[c * 32 + i] = 0xff800000u;   // -inf
  }
  // three-input maxima (FMNMX3 on sm_100): 32 instructions for 64 scores instead of 64
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < attn::HALF; i += 8) {
    m0 = fmax3(m0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
    m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
    m2 = fmax3(m2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
    m3 = fmax3(m3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
  }
  return fmax3(fmaxf(m0, m1), m2, m3);
}

// row maximum of one 32-key part (single-tile CTAs with at most 64 query rows, see run_softmax_dup)
template <bool kMasked>
__device__ __forceinline__ float part32_row_max(uint32_t (&s)[32], uint32_t word) {
  if (kMasked) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!((word >> i) & 1u)) s[i] = 0xff800000u;   // -inf
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    m0 = fmax3(m0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
    m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
  }
  return fmaxf(m0, m1);
}

// kLse: the training forward also stores each row's base-2 log-sum-exp (a separate instantiation so that the inference
// kernel keeps the register allocation it was tuned with).
//
// kLoop (inference, more work items than SMs): ONE CTA per SM walks work items instead of exiting after one -- the first is
// its block index, the following ones come from a global counter (fetched one item ahead by the TMA thread).  Nothing is
// carried across items except the tensor-memory allocation: after an item all 640 threads meet at a CTA-wide barrier, warp 16
// re-initialises the mbarriers (so every phase count below starts from zero again, exactly as in a fresh CTA) and requests
// the next item's Q / K / V^T tiles, and a second barrier releases the roles.  That replaces the exit / block-scheduler /
// launch / tensor-memory-allocation gap between two CTAs on an SM (measured 1.2 us, 5 % of the video self-attention) by two
// barriers; the key loop is the same code with the same register budget (unlike attention_p_kernel below, which overlaps
// items and pays for the cross-item state inside the loop).
// MEASURED (B200, kernels alone, us per launch: video self / cross / query self-attention), OFF by default (SVOL_ATTN_LOOP=1):
//   one CTA per item (default)                 245.4 / 71.0 / 28.7
//   this variant's code, one CTA per item      259.5 / 77.1 / 33.7     (SVOL_ATTN_LOOP=2)
//   looping CTAs                               248.5 / 81.2 / 31.7
// i.e. the hand-over does save what the timeline predicted on the video shape (11 us of 259) but the same source inside an
// item loop compiles to a slower item (547 instead of 537 instructions per key tile and more special-register reads on the
// latency-bound chain; bit-identical results), and the heterogeneous cross-attention items (a two-tile item and a
// 64-row duplicated-row item per head) lose on top of that.  Outputs are bit-identical to the default kernel's
// (tests/test_kernels_gpu.py::test_looping_attention_ctas_match_one_cta_per_item).
template <bool kLse, bool kLoop>
__global__ void __launch_bounds__(attn::THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmQ64,
                    const float* __restrict__ key_mask,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int lse_pitch, int H, int Lq, int Lk, int ldo,
                    int n_items, unsigned int* __restrict__ item_counter) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + OFF_BAR);

  // `lane` and the tensor-memory base are re-derived inside each role (opaque to common-subexpression elimination): as
  // values of the common prologue they would have to live in the 32-register budget of the producer / issuer warpgroup
  // and were spilled to local memory there
  const int warp = threadIdx.x >> 5;
  auto lane_id = []() { int l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; };
  auto tmem_base_of = [](const AttnBars* b) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&b->tmem_base)) : "memory"); return v; };
  // work item = (pair of query tiles, head, sample); kLoop: linear index, query pair fastest (the CTAs running at the same
  // time share a head's K / V^T in L2, as the 3-D grid's launch order does)
  int q0, h, b, n_q;
  bool split, dup;
  const int n_tiles = (Lk + BKV - 1) / BKV;
  auto decode_item = [&](int item) {
    if (kLoop) {
      const int nx = (Lq + 2 * BQ - 1) / (2 * BQ);
      q0 = (item % nx) * (2 * BQ); h = (item / nx) % H; b = item / (nx * H);
    } else {
      q0 = blockIdx.x * (2 * BQ); h = blockIdx.y; b = blockIdx.z;
    }
    n_q = (q0 + BQ < Lq) ? 2 : 1;          // is the second query tile of this item populated?
    // single-tile item: both tile slots work on query tile 0, slot t on key tiles t, t + 2, ... (see the header)
    // Both variants of the role code below are separate instantiations (generic lambdas on a compile-time flag): the
    // two-tile path keeps exactly the code it was tuned with (key tile == iteration, no extra live values).
    // (with fewer than four key tiles the extra merge costs more than the shorter walk saves: query self-attention, 3 tiles)
    split = n_q == 1 && n_tiles >= 4;
    // ... and when that one query tile holds at most 64 rows (cross-attention: 320 = 2 x 128 + 64 queries; video self-attention:
    // 1568 = 12 x 128 + 32), the rows are loaded TWICE (tile rows 64..127 = rows 0..63 again) and the two copies split every
    // half tile's 64 keys: copy 0 exponentiates keys [0, 32), copy 1 keys [32, 64), each writing zeros for the other part's
    // probabilities once.  The MMAs are unchanged; every softmax warp issues half the MUFU / row-max work per key tile
    // instead of spending it on rows that do not exist, and eight partial results per row are merged at the end.
    dup = split && Lq - q0 <= 64;
  };
  // kLoop: the current item lives in SHARED memory (bars->cur_item) and is re-read where it is needed (item set-up, output
  // addresses in the epilogue) instead of being carried in registers through the key loop: with the item, its coordinates and
  // an iteration count live across the loop ptxas rematerialised the shared-memory base (S2UR / S2R of special registers) in
  // front of every barrier operation of the key loop -- 7 % more instructions on the latency-bound chain, 10 % slower.
  auto current_item = [&]() -> int {
    int v = 0;
    if (kLoop) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&bars->cur_item)) : "memory");
    return v;
  };
  decode_item(kLoop ? static_cast<int>(blockIdx.x) : 0);
#ifdef SVOL_ATTN_TRACE
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (threadIdx.x == 0 && cta_lin < 8192) {
    uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_attn_cta[cta_lin][0] = smid;
    g_attn_cta[cta_lin][1] = global_ns();
  }
#endif
  // key tiles whose upper 64 keys hold at least one in-range key: when Lk % 128 is in (0, 64] the upper half of the last
  // tile is empty and its QK^T / softmax / PV are skipped altogether (1568 keys: 1 of 26 half tiles; 320 keys: 1 of 6)
  const int n_hi = Lk > HALF ? (Lk - HALF + BKV - 1) / BKV : 0;
#ifdef SVOL_ATTN_TRACE
#ifdef SVOL_ATTN_TRACE_LAST      // trace the LAST query-tile pair of (sample 0, head 0): the single-tile CTA when Lq % 256 is in (0, 128]
  const bool trace_on = blockIdx.x == gridDim.x - 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane_id() == 0 && ((warp & 3) == 0 || warp >= 16);
#else
  const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane_id() == 0 && ((warp & 3) == 0 || warp >= 16);
#endif
  if (warp == 0) SVOL_TR(0, 60, 3);                   // kernel entry (warp 0)
#endif

  // Requests the Q tile(s) and the first fills of the K / V^T ring of the current item (one thread; the barriers are fresh)
  auto request_first_tiles = [&]() {
    mbar_arrive_expect_tx(&bars->q_full, n_q * Q_BYTES);
    if (dup) {          // the same 64 rows into both halves of the tile (64 rows x 64 B = 4 KB: a whole number of swizzle atoms)
      tma_load_2d(smem, &tmQ64, &bars->q_full, h * DH, b * Lq + q0);
      tma_load_2d(smem + Q_BYTES / 2, &tmQ64, &bars->q_full, h * DH, b * Lq + q0);
    } else {
      for (int t = 0; t < n_q; ++t)
        tma_load_2d(smem + t * Q_BYTES, &tmQ, &bars->q_full, h * DH, b * Lq + q0 + t * BQ);
    }
    const int vrow = (b * H + h) * DH;
    for (int j = 0; j < min(n_tiles, STAGES); ++j) {
      mbar_arrive_expect_tx(&bars->kv_full[j], K_BYTES + VT_BYTES);
      tma_load_2d(smem + OFF_K + j * K_BYTES, &tmK, &bars->kv_full[j], h * DH, b * Lk + j * BKV);
      tma_load_2d(smem + OFF_VT + j * VT_BYTES, &tmVt, &bars->kv_full[j], j * BKV, vrow);
      tma_load_2d(smem + OFF_VT + j * VT_BYTES + VT_KB_BYTES, &tmVt, &bars->kv_full[j], j * BKV + HALF, vrow);
    }
  };
  if (warp == 16 && lane_id() == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    mbar_init(&bars->q_full, 1);
    for (int g = 0; g < 4; ++g) {
      mbar_init(&bars->s_full[g], 1);
      mbar_init(&bars->s_free[g], 4);
      mbar_init(&bars->p_ready[g], 4);
      mbar_init(&bars->o_full[g], 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], n_q);          // one commit per active MMA issuer
    }
    fence_barrier_init();
    // The Q tile(s) and the first fills of the K / V^T ring are requested HERE, before the CTA-wide barrier below: the loads
    // only need the barriers this thread has just initialised, and their L2 / DRAM latency then overlaps the tensor-memory
    // allocation, the barrier and the role set-up instead of following them (a CTA lives for only ~13 key tiles).
    // (Programmatic dependent launch: Q / K / V^T are the previous kernel's outputs -- this thread waits for it here, the
    // other warps at the CTA-wide barrier below, i.e. barrier set-up and tensor-memory allocation run under its tail.)
    griddep_wait();
    request_first_tiles();
    if (kLoop) bars->cur_item = static_cast<int>(blockIdx.x);
    if (!kLoop && n_tiles <= STAGES) griddep_launch_dependents();     // every load of this CTA is requested (see the producer)
  }
  SVOL_TR(0, 60, 0);                                  // (trace build, tools/attn_ends_trace.py: warp 0 reaches the set-up barrier)
  if (warp == 17) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  SVOL_TR(0, 60, 1);                                  // set-up barrier passed (tensor memory allocated)

  // kLoop: end of a work item, executed by all 640 threads.  Returns the next item (>= n_items: none).
  auto next_item_sync = [&]() -> int {
    __syncwarp();                                    // (single-thread roles: the whole warp arrives together)
    tcgen05_fence_before();
    asm volatile("bar.sync 0;" ::: "memory");        // every role is done with the item: tiles consumed, O read out, P / S idle
    // (the producer writes next_item for the FOLLOWING item only after the second barrier below)
    const int nxt = *reinterpret_cast<volatile int*>(&bars->next_item);
    if (warp == 16 && nxt < n_items) {
      const int l = lane_id();
      // The only asynchronous arrivals nobody has waited for are the issuers' last releases of the ring stages: stage l
      // was released once per key tile l, l + STAGES, ... that has an upper half
      if (l < STAGES) {
        const int rel = n_hi > l ? (n_hi - l + STAGES - 1) / STAGES : 0;
        if (rel > 0) mbar_wait(&bars->kv_empty[l], (rel - 1) & 1);
      }
      __syncwarp();
      decode_item(nxt);
      // fresh barriers (one or two per lane), so that every phase count of the next item starts from zero
      uint64_t* bar_array = reinterpret_cast<uint64_t*>(bars);
#if SVOL_ATTN_REINIT_MODE == 0
      if (l == 0) {
        bars->cur_item = nxt;
        for (int i = 0; i < 17 + 2 * STAGES; ++i) {
          const uint32_t count = (i >= 5 && i < 13) ? 4u : (i >= 17 + STAGES ? static_cast<uint32_t>(n_q) : 1u);
          mbar_init(bar_array + i, count);
        }
        fence_barrier_init();
        request_first_tiles();
      }
#else
      for (int i = l; i < 17 + 2 * STAGES; i += 32) {
        const uint32_t count = (i >= 5 && i < 13) ? 4u : (i >= 17 + STAGES ? static_cast<uint32_t>(n_q) : 1u);
        if (SVOL_ATTN_REINIT_MODE == 2) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar_array + i)) : "memory");
        mbar_init(bar_array + i, count);
      }
      fence_barrier_init();
      __syncwarp();
      if (l == 0) { bars->cur_item = nxt; request_first_tiles(); }
#endif
    }
    asm volatile("bar.sync 0;" ::: "memory");
    tcgen05_fence_after();
    return nxt;
  };

  // register re-split (each role branch starts with its own setmaxnreg so that it dominates the role's code).
  // The CTA is launched with 640 x 96 registers and setmaxnreg only redistributes them: 4 x 112 + 32 = 5 x 96.  (With 24
  // for the fifth warpgroup the issuers' descriptors and counters were spilled INSIDE their issue loops.)
  if (warp >= 16) {
    // (giving these registers back at kernel entry instead, so that the softmax warps' setmaxnreg.inc -- 660 clk after the set-up
    // barrier in the trace -- finds them free, changes nothing: 243.2 vs 243.0 us; the first tiles' load latency covers it)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    for (;;) {
    if (kLoop) decode_item(current_item());
    if (warp == 16) {
      // ------------------------------------------------------------------ TMA producer
      if (elect_one()) {
        // kLoop: the item after this one is fetched now and published when this item's loads are all requested
        unsigned int fetched = 0;
        if (kLoop) fetched = atomicAdd(item_counter, 1u);
        // (Q and the first STAGES fills were requested in the prologue / at the end of the previous item)
        const int vrow = (b * H + h) * DH;
        for (int j = STAGES; j < n_tiles; ++j) {
          const int s = j % STAGES;
          mbar_wait(&bars->kv_empty[s], ((j / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->kv_full[s], K_BYTES + VT_BYTES);
          tma_load_2d(smem + OFF_K + s * K_BYTES, &tmK, &bars->kv_full[s], h * DH, b * Lk + j * BKV);
          tma_load_2d(smem + OFF_VT + s * VT_BYTES, &tmVt, &bars->kv_full[s], j * BKV, vrow);
          tma_load_2d(smem + OFF_VT + s * VT_BYTES + VT_KB_BYTES, &tmVt, &bars->kv_full[s], j * BKV + HALF, vrow);
          // last load requested: once every CTA of the grid is this far (or gone) the next kernel on the stream may start
          // on the SMs that free up and run its set-up under this kernel's tail
          if (!kLoop && j == n_tiles - 1) griddep_launch_dependents();
        }
        if (kLoop) {
          bars->next_item = static_cast<int>(fetched + gridDim.x);
          // n_items fetches per launch (one per item processed); the last one leaves the counter at zero for the next launch
          if (fetched == static_cast<unsigned int>(n_items - 1)) atomicExch(item_counter, 0u);
        }
      }
    } else if (warp == 17 || warp == 18) {
      // ------------------------------------------------------------------ MMA issuer of tile slot t
      const int t = warp - 17;
      const uint32_t tmem_base = tmem_base_of(bars);
      auto run_issuer = [&](auto split_tag) {
      constexpr bool kSplit = decltype(split_tag)::value;
      constexpr int j_step = kSplit ? 2 : 1;
      if (!kSplit && t >= n_q) return;                    // second query tile empty and not worth splitting: this issuer idles
      const int j0 = kSplit ? t : 0;
      // key tiles of this slot: all / those with a populated upper half
      const int cnt_lo = n_tiles > j0 ? (n_tiles - j0 + j_step - 1) / j_step : 0;
      const int cnt_hi = n_hi > j0 ? (n_hi - j0 + j_step - 1) / j_step : 0;
      if (elect_one()) {
        constexpr uint32_t idesc_s = make_idesc_bf16(BQ, HALF);
        constexpr uint32_t idesc_o = make_idesc_bf16(BQ, DH);
        // descriptors differ only in their 14-bit start-address field (bytes >> 4): plain integer adds below
        const uint64_t dQ = make_kmajor_desc<64>(smem_u32(smem + (kSplit ? 0 : t) * Q_BYTES));
        const uint64_t dK = make_kmajor_desc<64>(smem_u32(smem + OFF_K));
        const uint64_t dV = make_kmajor_desc<128>(smem_u32(smem + OFF_VT));
        auto issue_qk = [&](int i, int half) {          // i-th key tile of this slot: S_g = Q K_half(j)^T
          const int j = j0 + j_step * i;
          const int g = 2 * t + half, s = j % STAGES;
          if (half == 0) {
            // single-tile CTA: this issuer consumes every other key tile only, but it still observes EVERY fill of the ring
            // in order (the skipped tile's barrier first), so that a parity wait can never alias a fill two phases back
            if (kSplit && j > 0) mbar_wait(&bars->kv_full[(j - 1) % STAGES], ((j - 1) / STAGES) & 1);
            mbar_wait(&bars->kv_full[s], (j / STAGES) & 1);
          }
          if (i > 0) mbar_wait(&bars->s_free[g], (i - 1) & 1);
          SVOL_TR(4 + t, i, half);
          tcgen05_fence_after();
          const uint64_t dKs = dK + static_cast<uint64_t>(s * (K_BYTES >> 4) + half * (HALF * DH * 2 >> 4));
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + g * HALF, dQ + 2 * k, dKs + 2 * k, idesc_s, k != 0);
          umma_commit(&bars->s_full[g]);
        };
        auto issue_pv = [&](int i, int half) {          // O_g += P_g V_half(j)
          const int j = j0 + j_step * i;
          const int g = 2 * t + half, s = j % STAGES;
          mbar_wait(&bars->p_ready[g], i & 1);
          SVOL_TR(4 + t, i, 2 + half);
          tcgen05_fence_after();
          const uint64_t dVs = dV + static_cast<uint64_t>(s * (VT_BYTES >> 4) + half * (VT_KB_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < HALF / 16; ++k)
            umma_bf16_ts(tmem_base + TMEM_O + g * DH, tmem_base + TMEM_P + g * 32 + k * 8, dVs + 2 * k, idesc_o,
                         (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&bars->o_full[g]);
          // last reader of stage s (covers every earlier MMA); a last tile without an upper half is never reloaded
          if (half == 1) umma_commit(&bars->kv_empty[s]);
        };
        mbar_wait(&bars->q_full, 0);
        for (int i = 0; i <= cnt_lo + 1; ++i) {
          if (i < cnt_lo) issue_qk(i, 0);
          if (i >= 2 && i - 2 < cnt_hi) issue_pv(i - 2, 1);
          if (i < cnt_hi) issue_qk(i, 1);
          if (i >= 1 && i <= cnt_lo) issue_pv(i - 1, 0);
        }
      }
      };
      if (split) run_issuer(std::true_type{}); else run_issuer(std::false_type{});
    }
    if (!kLoop) break;
    if (next_item_sync() >= n_items) break;
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // Barrier probes are software-pipelined (see the key loop); their predicate registers are declared once
    asm volatile(".reg .pred p_of, p_sf;");
    SVOL_TR(0, 60, 2);                                // registers re-split
    for (;;) {
    if (kLoop) decode_item(current_item());
    const int g = warp >> 2;                            // softmax warpgroup
    const int t = g >> 1, half = g & 1;                 // query tile, key half
    const int lane = lane_id();
    const uint32_t tmem_base = tmem_base_of(bars);
    // key_padding_mask of this sample as a bitmask in shared memory (built once, under the TMA / first-QK^T latency):
    // reading the float mask inside the key loop put an L2 round trip on every tile's critical path
    uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + OFF_KMASK);
    const uint32_t a_kmask = smem_u32(kmask);             // read back with ld.shared (through the generic pointer: LD.E in the key loop)
    const bool mask_in_smem = key_mask != nullptr && (Lk + 31) / 32 <= KMASK_WORDS;
    if (mask_in_smem) {
      const float* mr = key_mask + static_cast<size_t>(b) * Lk;
      for (int w = warp; w < (Lk + 31) / 32; w += 16) {
        const int kv = w * 32 + lane;
        const uint32_t m = __ballot_sync(0xffffffffu, kv < Lk && __ldg(mr + kv) != 0.f);
        if (lane == 0) kmask[w] = m;
      }
      asm volatile("bar.sync 3, 512;" ::: "memory");     // all 16 softmax warps
    }
    {
      // ------------------------------------------------------------------ softmax warpgroups: two query tiles x two key
      // halves, or -- single-tile CTA with at least four key tiles -- one query tile x two key halves x even / odd key tiles
      const int quarter = warp & 3;                       // TMEM lane quarter == warp % 4
      const int r = quarter * 32 + lane;                  // row inside the tile == TMEM lane
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const uint32_t t_s = t_lane + g * HALF;
      const uint32_t t_p = t_lane + TMEM_P + g * 32, t_o = t_lane + TMEM_O + g * DH;
      const float* mrow = key_mask ? key_mask + static_cast<size_t>(b) * Lk : nullptr;
      float m_ref = -INFINITY;      // reference maximum the stored probabilities / O / l are relative to
      float mx_seen = -INFINITY;    // largest row maximum seen so far (drives the lazy rescaling)
      bool refs_finite = false;     // warp-uniform: every row of this warp has a finite reference maximum
      float2 l2 = make_float2(0.f, 0.f);

      // start offsets: g0, g2, g1, g3 a quarter period apart (lo before hi inside a tile, as the MMA issue order assumes)
      {
        // (no stagger for a handful of key tiles -- query self-attention, 3 tiles: there is no steady state to protect and
        // the last warpgroup would start 1650 clk late in a CTA that lives ~16 k clk)
        const int slot = n_tiles >= 6 ? half * 2 + t : 0;
        const long long t_start = clock64();
        while (clock64() - t_start < static_cast<long long>(slot) * STAGGER_CLK) {}
      }

      // Barrier probes are software-pipelined: a (non-blocking) mbarrier.test_wait is issued well before its result
      // is needed and consumed after independent work; the blocking wait is only the fallback.
      // a_g = address of s_full[g]; s_free[g] / p_ready[g] / o_full[g] follow at +32 / +64 / +96 (AttnBars).  Opaque, so that it
      // is computed once per item and kept in a register (see mbar_wait_a); likewise the lane-0 flag of the arrivals.
      uint32_t a_g = smem_u32(&bars->s_full[g]);
      uint32_t leader = lane == 0 ? 1u : 0u;
      asm volatile("" : "+r"(a_g), "+r"(leader));
      const uint32_t a_sfull = a_g, a_ofull = a_g + 96;

      auto run_softmax = [&](auto split_tag) {
      constexpr bool kSplit = decltype(split_tag)::value;
      constexpr int j_step = kSplit ? 2 : 1;
      if (!kSplit && t >= n_q) return;                    // (see the issuer)
      uint32_t s_ready = 0;
      const int j0 = kSplit ? t : 0;
      const int n_half = half ? n_hi : n_tiles;           // key tiles that have keys in this warpgroup's half ...
      const int n_mine = n_half > j0 ? (n_half - j0 + j_step - 1) / j_step : 0;    // ... of which this tile slot takes these
      for (int i = 0; i < n_mine; ++i) {
        const int j = j0 + j_step * i;                    // key tile; barrier phases count i, this warpgroup's own iterations
        const int kv0 = j * BKV + half * HALF;
        SVOL_TR(g, i, 0);
        if (!s_ready) mbar_wait_a(a_g, i & 1);
        SVOL_TR(g, i, 1);
        tcgen05_fence_after();
        uint32_t s[HALF];
        tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld_wait();
        SVOL_TR(g, i, 2);
        // scores are in registers: hand the TMEM buffer back so the next QK^T can start now
        tcgen05_fence_before();
        __syncwarp();
        if (leader) mbar_arrive_a(a_g + 32);
        SVOL_TR(g, i, 4);

        // validity of this half tile's 64 keys as two 32-bit words (ragged tail and key_padding_mask); the
        // masked variant of the row-max code is a separate instantiation so full tiles pay nothing for it
        float mx;
        bool masked = false;
        uint32_t words[2];
        if (mask_in_smem) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int w = (kv0 >> 5) + c;
            words[c] = w * 32 < Lk ? lds_u32(a_kmask + w * 4) : 0u;
            masked |= words[c] != 0xffffffffu;
          }
        } else if (mrow != nullptr || kv0 + HALF > Lk) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int kv = kv0 + c * 32 + lane;
            const bool ok = kv < Lk && (mrow == nullptr || __ldg(mrow + kv) != 0.f);
            words[c] = __ballot_sync(0xffffffffu, ok);
            masked |= words[c] != 0xffffffffu;
          }
        }
        // Speculative tiles (every unmasked tile after the first, once all rows have a finite reference): the exponentials
        // below run against the CURRENT reference maximum and this tile's row maximum is folded into their loop (FMNMX3 on
        // the ALU pipe between the MUFU instructions) instead of preceding it -- ~360 clk off each warp's serial chain per
        // key tile.  The lazy-rescaling test then looks at the maximum seen up to the PREVIOUS tile; a tile whose scores
        // outgrow the reference only produces probabilities above 2^8 for that one tile (fp32 / bf16 exponent range is not
        // at risk below 2^127), and the reference is raised before the next one.
        const bool spec = SVOL_ATTN_SPEC_MAX && i > 0 && !masked && refs_finite;
        if (!spec) {
          if (masked) mx = half_row_max<true>(s, words);
          else mx = half_row_max<false>(s, words);
          mx_seen = fmaxf(mx_seen, mx);
        }
        SVOL_TR_AFTER(g, i, 3, mx_seen);

        if (i == 0) {
          m_ref = mx;
          refs_finite = !__any_sync(0xffffffffu, m_ref == -INFINITY);
        } else {
          // lazy rescaling: only when some row's maximum outgrew the reference by more than 2^8
          const bool need = mx_seen > m_ref + RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2_approx(m_ref - mx_seen) : 1.0f;   // (m_ref = -inf, finite maximum) -> 0
            mbar_wait_a(a_g + 96, (i - 1) & 1);                  // every earlier P V has landed in O_g
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(t_o + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x16(t_o + c * 16, o);
            }
            tmem_st_wait();
            l2.x *= alpha; l2.y *= alpha;
            if (need) m_ref = mx_seen;
            refs_finite = !__any_sync(0xffffffffu, m_ref == -INFINITY);
          }
        }
        float m_use = m_ref == -INFINITY ? 0.f : m_ref;
        if (i > 0)   // probe: has the previous P V landed (P columns reusable)?  consumed after the exponentials
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 p_of, [%0], %1;" ::"r"(a_ofull), "r"((i - 1) & 1) : "memory");

        // one MUFU ex2 per probability (ex2.approx.ftz.bf16x2 was tried: on sm_100 it is issued as two
        // MUFU.EX2.BF16 ops plus a PRMT, so it saves nothing and only costs precision); the subtraction
        // of the reference maximum and the row sum are packed FADD2s
        const float2 neg_m = make_float2(-m_use, -m_use);
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
        if (spec) {
          // scores -> probabilities with the row maximum of THIS tile accumulated on the side (see above)
          float m0 = mx_seen, m1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < HALF; i += 4) {
            const float s0 = __uint_as_float(s[i]), s1 = __uint_as_float(s[i + 1]), s2 = __uint_as_float(s[i + 2]), s3 = __uint_as_float(s[i + 3]);
            m0 = fmax3(m0, s0, s1);
            m1 = fmax3(m1, s2, s3);
            const float2 x0 = __fadd2_rn(make_float2(s0, s1), neg_m);
            const float2 x1 = __fadd2_rn(make_float2(s2, s3), neg_m);
            const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
            const float2 p1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
            la = __fadd2_rn(la, p0);
            lb = __fadd2_rn(lb, p1);
            s[i >> 1] = pack_bf16x2(p0.x, p0.y);
            s[(i >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          }
          mx_seen = fmaxf(m0, m1);
        } else if (SVOL_ATTN_POLY_PAIRS > 0 && !masked) {
          // unmasked tile: SVOL_ATTN_POLY_PAIRS of every four pairs go through the packed FMA-pipe polynomial
          const float2 magic_m = make_float2(12582912.f - m_use, 12582912.f - m_use);
#pragma unroll
          for (int i = 0; i < HALF; i += 8) {
            float2 p[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 sc = make_float2(__uint_as_float(s[i + 2 * k]), __uint_as_float(s[i + 2 * k + 1]));
              if (k >= 4 - SVOL_ATTN_POLY_PAIRS) {
                p[k] = ex2_poly2(sc, magic_m, neg_m);
              } else {
                const float2 x = __fadd2_rn(sc, neg_m);
                p[k] = make_float2(ex2_approx(x.x), ex2_approx(x.y));
              }
            }
            la = __fadd2_rn(la, __fadd2_rn(p[0], p[1]));
            lb = __fadd2_rn(lb, __fadd2_rn(p[2], p[3]));
#pragma unroll
            for (int k = 0; k < 4; ++k) s[(i >> 1) + k] = pack_bf16x2(p[k].x, p[k].y);
          }
        } else {
#pragma unroll
        for (int i = 0; i < HALF; i += 4) {
          const float2 x0 = __fadd2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), neg_m);
          const float2 x1 = __fadd2_rn(make_float2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), neg_m);
          const float2 p0 = make_float2(ex2_approx(x0.x), SVOL_ATTN_POLY_PER4 >= 2 ? ex2_poly(x0.y) : ex2_approx(x0.y));
          const float2 p1 = make_float2(ex2_approx(x1.x), SVOL_ATTN_POLY_PER4 >= 1 ? ex2_poly(x1.y) : ex2_approx(x1.y));
          la = __fadd2_rn(la, p0);
          lb = __fadd2_rn(lb, p1);
          s[i >> 1] = pack_bf16x2(p0.x, p0.y);
          s[(i >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
        }
        }
        l2 = __fadd2_rn(l2, __fadd2_rn(la, lb));
        SVOL_TR(g, i, 5);

        // the previous P V must be done reading the P columns before they are overwritten
        if (i > 0) {
          uint32_t ok;
          asm volatile("selp.u32 %0, 1, 0, p_of;" : "=r"(ok));
          if (!ok) mbar_wait_a(a_g + 96, (i - 1) & 1);
          tcgen05_fence_after();
        }
        if (i + 1 < n_mine)   // probe the next score tile (its QK^T was issued when this tile's scores were read out)
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 p_sf, [%0], %1;" ::"r"(a_sfull), "r"((i + 1) & 1) : "memory");
        SVOL_TR(g, i, 6);
        // P half tile -> tensor memory: lane = query row, 32 columns of packed bf16 pairs (the A operand of P V)
        tmem_st_32x32b_x32(t_p, &s[0]);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (leader) mbar_arrive_a(a_g + 64);
        s_ready = 0;
        if (i + 1 < n_mine) asm volatile("selp.u32 %0, 1, 0, p_sf;" : "=r"(s_ready));
        SVOL_TR(g, i, 7);
      }

      // ---- epilogue: O_g is complete once the last P V has landed; merge the partial results of each row
      uint32_t o[DH];
      if (n_mine > 0) {
        mbar_wait_a(a_g + 96, (n_mine - 1) & 1);
        tcgen05_fence_after();
        tmem_ld_32x32b_x32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld_wait();
      } else {                                            // this warpgroup had no key tile (m = -inf, l = 0)
#pragma unroll
        for (int i = 0; i < DH; ++i) o[i] = 0u;
      }
      SVOL_TR(g, 61, 0);                                 // last P V landed, O in registers
      const float l_mine = l2.x + l2.y;
      // two query tiles: warpgroup (t, hi) hands its partial to (t, lo) through slot t.  Single-tile CTA: warpgroups 1, 2, 3
      // hand theirs to warpgroup 0 through slots 0, 1, 2.
      const uint32_t a_cmb = smem_u32(smem + OFF_CMB);    // (shared-space address: see cmb_store)
      const bool writer = kSplit ? (g != 0) : (half == 1);
      if (writer) cmb_store(a_cmb + (((kSplit ? g - 1 : t) * BQ + r) * CMB_STRIDE) * 4, m_ref, l_mine, o);
      if (kSplit) asm volatile("bar.sync 4, 512;" ::: "memory");      // all four warpgroups
      else asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");   // the two warpgroups of query tile t
      if (!writer) {
        constexpr int n_parts = kSplit ? 3 : 1;
        const int first = kSplit ? 0 : t;
        float m = m_ref;
#pragma unroll
        for (int pi = 0; pi < n_parts; ++pi) m = fmaxf(m, lds_f32(a_cmb + (((first + pi) * BQ + r) * CMB_STRIDE) * 4));
        const float m_safe = m == -INFINITY ? 0.f : m;
        const float a_own = ex2_approx(m_ref - m_safe);
        float l_tot = a_own * l_mine;
        float acc[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i) acc[i] = __uint_as_float(o[i]) * a_own;
#pragma unroll
        for (int pi = 0; pi < n_parts; ++pi) cmb_accumulate(a_cmb + (((first + pi) * BQ + r) * CMB_STRIDE) * 4, m_safe, l_tot, acc);
        const float inv = 1.0f / l_tot;
        if (kLoop) decode_item(current_item());             // (not carried through the key loop: see current_item)
        const int q = q0 + (kSplit ? 0 : t) * BQ + r;
        if (q < Lq) {
          // training forward: base-2 log-sum-exp of the (pre-scaled) scores, so that the backward recomputes
          // P = 2^(S - lse) without a second softmax pass
          if (kLse) lse[(static_cast<size_t>(b) * H + h) * lse_pitch + q] = m_safe + __log2f(l_tot);
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
      SVOL_TR(g, 61, 1);                                 // merged and stored (or partial handed over)
      };
      // ---- single-tile CTA with at most 64 query rows: the split walk above with the two row copies sharing each half tile
      auto run_softmax_dup = [&]() {
      const int copy = quarter >> 1;                      // rows 0..63: keys [0, 32) of the half tile; rows 64..127: keys [32, 64)
      const int r64 = r & 63;                             // query row of this thread
      uint32_t s_ready = 0;
      const int j0 = t;
      const int n_half = half ? n_hi : n_tiles;
      const int n_mine = n_half > j0 ? (n_half - j0 + 1) / 2 : 0;
      {
        // probabilities of the OTHER copy's keys are zero in this warp's rows for the whole walk: written once (ordered before
        // the first p_ready arrival by the fence in front of it)
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        tmem_st_32x32b_x16(t_p + (1 - copy) * 16, z);
        tmem_st_wait();
      }
      for (int i = 0; i < n_mine; ++i) {
        const int j = j0 + 2 * i;
        const int kv0 = j * BKV + half * HALF + copy * 32;
        if (!s_ready) mbar_wait_a(a_g, i & 1);
        tcgen05_fence_after();
        uint32_t s[32];
        tmem_ld_32x32b_x32(t_s + copy * 32, s);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (leader) mbar_arrive_a(a_g + 32);

        uint32_t word = 0xffffffffu;
        if (mask_in_smem) {
          const int w = kv0 >> 5;
          word = w * 32 < Lk ? lds_u32(a_kmask + w * 4) : 0u;
        } else if (mrow != nullptr || kv0 + 32 > Lk) {
          const int kv = kv0 + lane;
          word = __ballot_sync(0xffffffffu, kv < Lk && (mrow == nullptr || __ldg(mrow + kv) != 0.f));
        }
        float mx;
        if (word != 0xffffffffu) mx = part32_row_max<true>(s, word);
        else mx = part32_row_max<false>(s, word);
        mx_seen = fmaxf(mx_seen, mx);

        if (i == 0) {
          m_ref = mx;
        } else {
          const bool need = mx_seen > m_ref + RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2_approx(m_ref - mx_seen) : 1.0f;
            mbar_wait_a(a_g + 96, (i - 1) & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(t_o + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st_32x32b_x16(t_o + c * 16, o);
            }
            tmem_st_wait();
            l2.x *= alpha; l2.y *= alpha;
            if (need) m_ref = mx_seen;
          }
        }
        const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
        if (i > 0)
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 p_of, [%0], %1;" ::"r"(a_ofull), "r"((i - 1) & 1) : "memory");
        const float2 neg_m = make_float2(-m_use, -m_use);
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float2 x0 = __fadd2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), neg_m);
          const float2 x1 = __fadd2_rn(make_float2(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])), neg_m);
          const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
          const float2 p1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
          la = __fadd2_rn(la, p0);
          lb = __fadd2_rn(lb, p1);
          s[e >> 1] = pack_bf16x2(p0.x, p0.y);
          s[(e >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
        }
        l2 = __fadd2_rn(l2, __fadd2_rn(la, lb));
        if (i > 0) {
          uint32_t ok;
          asm volatile("selp.u32 %0, 1, 0, p_of;" : "=r"(ok));
          if (!ok) mbar_wait_a(a_g + 96, (i - 1) & 1);
          tcgen05_fence_after();
        }
        if (i + 1 < n_mine)
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 p_sf, [%0], %1;" ::"r"(a_sfull), "r"((i + 1) & 1) : "memory");
        tmem_st_32x32b_x16(t_p + copy * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (leader) mbar_arrive_a(a_g + 64);
        s_ready = 0;
        if (i + 1 < n_mine) asm volatile("selp.u32 %0, 1, 0, p_sf;" : "=r"(s_ready));
      }

      // ---- epilogue: eight partial results per query row (4 warpgroups x 2 copies); (warpgroup 0, copy 0) merges
      uint32_t o[DH];
      if (n_mine > 0) {
        mbar_wait_a(a_g + 96, (n_mine - 1) & 1);
        tcgen05_fence_after();
        tmem_ld_32x32b_x32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < DH; ++i) o[i] = 0u;
      }
      const float l_mine = l2.x + l2.y;
      const uint32_t a_cmb = smem_u32(smem + OFF_CMB);
      const int slot = g * 2 + copy;
      if (slot != 0) cmb_store(a_cmb + (((slot - 1) * 64 + r64) * CMB_STRIDE) * 4, m_ref, l_mine, o);
      asm volatile("bar.sync 4, 512;" ::: "memory");      // all four warpgroups
      if (slot == 0) {
        float m = m_ref;
#pragma unroll
        for (int pi = 0; pi < 7; ++pi) m = fmaxf(m, lds_f32(a_cmb + ((pi * 64 + r64) * CMB_STRIDE) * 4));
        const float m_safe = m == -INFINITY ? 0.f : m;
        const float a_own = ex2_approx(m_ref - m_safe);
        float l_tot = a_own * l_mine;
        float acc[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i) acc[i] = __uint_as_float(o[i]) * a_own;
#pragma unroll
        for (int pi = 0; pi < 7; ++pi) cmb_accumulate(a_cmb + ((pi * 64 + r64) * CMB_STRIDE) * 4, m_safe, l_tot, acc);
        const float inv = 1.0f / l_tot;
        if (kLoop) decode_item(current_item());
        const int q = q0 + r64;
        if (q < Lq) {
          if (kLse) lse[(static_cast<size_t>(b) * H + h) * lse_pitch + q] = m_safe + __log2f(l_tot);
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
      };
      if (dup) run_softmax_dup();
      else if (split) run_softmax(std::true_type{});
      else run_softmax(std::false_type{});
    }
    if (!kLoop) break;
    if (next_item_sync() >= n_items) break;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  SVOL_TR(warp < 16 ? (warp >> 2) : 4, 61, 2);         // every role done
#ifdef SVOL_ATTN_TRACE
  if (threadIdx.x == 0 && cta_lin < 8192) g_attn_cta[cta_lin][3] = global_ns();
#endif
  if (warp == 17) {
    tcgen05_fence_after();
    tmem_dealloc<attn::TMEM_COLS>(tmem_base_of(bars));
  }
}

// =====================================================================================================================
// Persistent variant (round 2).  Same tiles, roles and inner loops as attention_tc_kernel, but ONE CTA per SM walks a list
// of work items (query-tile pair, head, sample) instead of exiting after one: tensor memory, barriers and descriptors are set
// up once, the TMA producer runs ahead across item boundaries (Q is double-buffered, the K / V^T ring simply keeps
// filling), and an issuer starts the next item's first Q K^T as soon as its score columns are free -- so the next item's
// pipeline fills while the current item's last tiles and epilogue drain.  At the headline shape a CTA lives for only 13 key
// tiles; the same kernel on the long clip (49 key tiles per CTA) reaches 416 TFLOP/s against 302 here, i.e. ~27 % of the
// one-item-per-CTA kernel's time was prologue, drain and wave quantisation (1792 CTAs on 148 SMs).
// Items are numbered query-pair-major, so the cheaper single-tile items (the last pair of every (sample, head)) come last
// and level the tail.  Every mbarrier phase is a toggling parity bit instead of the per-item iteration count.
//
// MEASURED (B200, kernels alone, us per launch; bit-identical outputs over 1000 launches, tools/attn_p_check.py), and the
// reason this kernel is NOT the default:
//                                 one item per CTA    persistent
//   video self (1792 items x 13 key tiles)   261.8         280.6
//   cross      ( 512 items x 13 key tiles)    72.1          95.6
//   long clip  ( 800 items x 49 key tiles)   410.7         479.3
//   112 items x 13 tiles, ONE item per CTA in both kernels:  27.9 vs 31.1  (tools/attn_p_time.py)
// The last line isolates the cause: with identical scheduling the persistent kernel's per-key-tile code is ~11 % slower --
// the cross-item state (phase bits, ring position, item decode) does not fit the 32 registers that the 4 x 112 softmax
// allocation leaves the issuer warpgroup (23 local-memory instructions in the issue loop) nor the softmax threads' 112
// (3 per key tile) -- and that costs more than the overlapped prologue / drain wins (~4 %); static round-robin assignment
// adds a 7 % imbalance (12.2 vs 11.4 two-tile units per CTA) that a work counter would remove.  Net: not better than the
// hardware's own CTA scheduler refilling an SM every ~20 us.
// Per-stage / per-Q-buffer release barriers always expect TWO arrivals, one per issuer: in a split single-tile item an
// issuer also passes and releases the key tiles it does not consume (the issuers drift apart at item boundaries, and a
// stage released by its consumer alone could be refilled twice before the other issuer looked at it: parity aliasing,
// found as a deadlock after ~40 launches); in an item with a single active issuer that issuer arrives twice.
// =====================================================================================================================
namespace attn_p {
using namespace attn;
constexpr int KMW = 128;                         // key-validity words per warp (keys up to 8192 -> else the in-loop fallback)
constexpr int OFF_Q2 = 0;                        // two Q buffers of two tiles each
constexpr int OFF_K = 2 * 2 * Q_BYTES, OFF_VT = OFF_K + STAGES * K_BYTES, OFF_CMB = OFF_VT + STAGES * VT_BYTES;
constexpr int OFF_KMASK = OFF_CMB + CMB_BYTES_3;
constexpr int OFF_BAR = OFF_KMASK + 16 * KMW * 4;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
static_assert(SMEM_BYTES <= 232448, "attention (persistent): shared memory budget");
}  // namespace attn_p

struct AttnPBars {
  uint64_t q_full[2], q_free[2];
  uint64_t s_full[4], s_free[4];
  uint64_t p_ready[4], o_full[4], o_free[4];
  uint64_t kv_full[attn::STAGES], kv_empty[attn::STAGES];
  uint32_t tmem_base, pad;
};
static_assert(sizeof(AttnPBars) <= 512, "barrier block");

template <bool kLse>
__global__ void __launch_bounds__(attn::THREADS, 1)
attention_p_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmVt, const float* __restrict__ key_mask,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int lse_pitch, int B, int H, int Lq, int Lk, int ldo) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnPBars* bars = reinterpret_cast<AttnPBars*>(smem + attn_p::OFF_BAR);

  const int warp = threadIdx.x >> 5;
  auto lane_id = []() { int l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; };
  auto tmem_base_of = [](const AttnPBars* b) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&b->tmem_base)) : "memory"); return v; };
  const int n_tiles = (Lk + BKV - 1) / BKV;
  const int n_hi = Lk > HALF ? (Lk - HALF + BKV - 1) / BKV : 0;
  const int nx = (Lq + 2 * BQ - 1) / (2 * BQ);          // query-tile pairs per (sample, head)
  const int BH = B * H;
  const int n_items = nx * BH;
  const int my_items = n_items > static_cast<int>(blockIdx.x) ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  // Items are numbered query-pair-major: every item but those of the last pair has two populated tiles, and whether a
  // single-tile item is split over both tile slots only depends on the key length -- so the item kind follows from the
  // pair index alone.
  const bool last_single = (nx - 1) * (2 * BQ) + BQ >= Lq;          // the last pair of every (sample, head) holds one tile
  const bool split_single = n_tiles >= 4;
  auto pair_of = [&](int n) { return (static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x)) / BH; };

  if (warp == 16 && lane_id() == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->q_full[i], 1); mbar_init(&bars->q_free[i], 2); }
    for (int g = 0; g < 4; ++g) {
      mbar_init(&bars->s_full[g], 1);
      mbar_init(&bars->s_free[g], 4);
      mbar_init(&bars->p_ready[g], 4);
      mbar_init(&bars->o_full[g], 1);
      mbar_init(&bars->o_free[g], 4);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 2);
    }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 16) {
      // ------------------------------------------------------------------ TMA producer (runs ahead across items)
      if (elect_one()) {
        int jt = 0;                                             // ring fills so far
        for (int n = 0; n < my_items; ++n) {
          const int idx = blockIdx.x + n * gridDim.x;
          const int qp = idx / BH, bh = idx - qp * BH;
          const int b = bh / H, h = bh - b * H;
          const int q0 = qp * (2 * BQ);
          const int n_q = (q0 + BQ < Lq) ? 2 : 1;
          const int qb = n & 1;
          mbar_wait(&bars->q_free[qb], ((n >> 1) & 1) ^ 1);      // both issuers are done with this Q buffer (passes on a fresh barrier)
          mbar_arrive_expect_tx(&bars->q_full[qb], n_q * Q_BYTES);
          for (int t = 0; t < n_q; ++t)
            tma_load_2d(smem + attn_p::OFF_Q2 + (qb * 2 + t) * Q_BYTES, &tmQ, &bars->q_full[qb], h * DH, b * Lq + q0 + t * BQ);
          for (int j = 0; j < n_tiles; ++j, ++jt) {
            const int s = jt % STAGES;
            mbar_wait(&bars->kv_empty[s], ((jt / STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(&bars->kv_full[s], K_BYTES + VT_BYTES);
            tma_load_2d(smem + attn_p::OFF_K + s * K_BYTES, &tmK, &bars->kv_full[s], h * DH, b * Lk + j * BKV);
            tma_load_2d(smem + attn_p::OFF_VT + s * VT_BYTES, &tmVt, &bars->kv_full[s], j * BKV, bh * DH);
            tma_load_2d(smem + attn_p::OFF_VT + s * VT_BYTES + VT_KB_BYTES, &tmVt, &bars->kv_full[s], j * BKV + HALF, bh * DH);
          }
        }
      }
    } else if (warp == 17 || warp == 18) {
      // ------------------------------------------------------------------ MMA issuer of tile slot t
      const int t = warp - 17;
      const uint32_t tmem_base = tmem_base_of(bars);
      if (elect_one()) {
        constexpr uint32_t idesc_s = make_idesc_bf16(BQ, HALF);
        constexpr uint32_t idesc_o = make_idesc_bf16(BQ, DH);
        const uint64_t dK = make_kmajor_desc<64>(smem_u32(smem + attn_p::OFF_K));
        const uint64_t dV = make_kmajor_desc<128>(smem_u32(smem + attn_p::OFF_VT));
        // Phase bits of this issuer's barriers (bit k toggles at every use): 0 / 1 s_free (the parity of the NEXT Q K^T use
        // of key half 0 / 1: the release it needs is the previous phase, and a fresh barrier passes), 2 / 3 p_ready,
        // 4 / 5 o_free (items completed by half 0 / 1).
        uint32_t ph = 0;
        int jj0 = 0;                                            // ring fill number of this item's key tile 0
        auto item = [&](auto split_tag, int qb, bool alone) {
          constexpr bool kSplit = decltype(split_tag)::value;
          constexpr int j_step = kSplit ? 2 : 1;
          const int j0 = kSplit ? t : 0;
          const int cnt_lo = n_tiles > j0 ? (n_tiles - j0 + j_step - 1) / j_step : 0;
          const int cnt_hi = n_hi > j0 ? (n_hi - j0 + j_step - 1) / j_step : 0;
          const uint64_t dQ = make_kmajor_desc<64>(smem_u32(smem + attn_p::OFF_Q2 + (qb * 2 + (kSplit ? 0 : t)) * Q_BYTES));
          auto issue_qk = [&](int i, int half) {
            const int jj = jj0 + j0 + j_step * i;               // ring fill number of this key tile
            const int g = 2 * t + half, s = jj % STAGES;
            if (half == 0) {
              // A split issuer consumes every other key tile, but it passes EVERY fill of the ring in order and releases the
              // tiles it skips as well (one arrival per issuer and stage, like in a two-tile item).  With the consumer
              // releasing a stage alone, an issuer that enters the item late (the two issuers drift apart at item
              // boundaries) found its first stage refilled twice and waited on an aliased parity forever.
              if (kSplit && j0 + j_step * i > 0) {
                mbar_wait(&bars->kv_full[(jj - 1) % STAGES], ((jj - 1) / STAGES) & 1);
                umma_commit(&bars->kv_empty[(jj - 1) % STAGES]);
              }
              mbar_wait(&bars->kv_full[s], (jj / STAGES) & 1);
            }
            mbar_wait(&bars->s_free[g], ((ph >> half) & 1u) ^ 1u);
            ph ^= 1u << half;
            tcgen05_fence_after();
            const uint64_t dKs = dK + static_cast<uint64_t>(s * (K_BYTES >> 4) + half * (HALF * DH * 2 >> 4));
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + g * HALF, dQ + 2 * k, dKs + 2 * k, idesc_s, k != 0);
            umma_commit(&bars->s_full[g]);
          };
          auto issue_pv = [&](int i, int half) {
            const int jj = jj0 + j0 + j_step * i;
            const int g = 2 * t + half, s = jj % STAGES;
            mbar_wait(&bars->p_ready[g], (ph >> (2 + half)) & 1u);
            ph ^= 4u << half;
            // the first P V of an item overwrites O_g: the previous item's accumulator must have been read out
            if (i == 0) mbar_wait(&bars->o_free[g], ((ph >> (4 + half)) & 1u) ^ 1u);
            tcgen05_fence_after();
            const uint64_t dVs = dV + static_cast<uint64_t>(s * (VT_BYTES >> 4) + half * (VT_KB_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < HALF / 16; ++k)
              umma_bf16_ts(tmem_base + TMEM_O + g * DH, tmem_base + TMEM_P + g * 32 + k * 8, dVs + 2 * k, idesc_o,
                           (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bars->o_full[g]);
            // last reader of the stage among this issuer's MMAs: the upper half's P V, or the lower half's when the
            // tile has no upper half (the last tile of a ragged key length); covers every earlier MMA of this thread
            if (half == 1 || i >= cnt_hi) {
              umma_commit(&bars->kv_empty[s]);
              if (alone) umma_commit(&bars->kv_empty[s]);
            }
          };
          for (int i = 0; i <= cnt_lo + 1; ++i) {
            if (i < cnt_lo) issue_qk(i, 0);
            if (i >= 2 && i - 2 < cnt_hi) issue_pv(i - 2, 1);
            if (i < cnt_hi) issue_qk(i, 1);
            if (i >= 1 && i <= cnt_lo) issue_pv(i - 1, 0);
          }
          if (kSplit && ((n_tiles - 1 - j0) & 1)) {
            // the item's last key tile belongs to the other issuer: pass and release it too
            const int jl = jj0 + n_tiles - 1;
            mbar_wait(&bars->kv_full[jl % STAGES], (jl / STAGES) & 1);
            umma_commit(&bars->kv_empty[jl % STAGES]);
          }
          // this item's Q tile(s) are no longer read (covers every Q K^T above)
          umma_commit(&bars->q_free[qb]);
          if (alone) umma_commit(&bars->q_free[qb]);
          if (cnt_lo > 0) ph ^= 16u;
          if (cnt_hi > 0) ph ^= 32u;
        };
        for (int n = 0; n < my_items; ++n, jj0 += n_tiles) {
          const bool single = last_single && pair_of(n) == nx - 1;
          const bool split = single && split_single;
          if (single && !split && t == 1) continue;             // (the active issuer releases stages / Q for both)
          mbar_wait(&bars->q_full[n & 1], (n >> 1) & 1);
          if (split) item(std::true_type{}, n & 1, false);
          else item(std::false_type{}, n & 1, single);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int g = warp >> 2;                            // softmax warpgroup
    const int t = g >> 1, half = g & 1;                 // tile slot, key half
    const int lane = lane_id();
    const uint32_t tmem_base = tmem_base_of(bars);
    uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + attn_p::OFF_KMASK) + warp * attn_p::KMW;   // this warp's own words
    const bool mask_words = key_mask != nullptr && (Lk + 31) / 32 <= attn_p::KMW;
    const int quarter = warp & 3;                       // TMEM lane quarter == warp % 4
    const int r = quarter * 32 + lane;                  // row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + g * HALF;
    const uint32_t t_p = t_lane + TMEM_P + g * 32, t_o = t_lane + TMEM_O + g * DH;
    asm volatile(".reg .pred pp_of, pp_sf;");
    const uint32_t a_sfull = smem_u32(&bars->s_full[g]), a_ofull = smem_u32(&bars->o_full[g]);
    uint32_t u = 0;                                     // parity of this warpgroup's next key-tile iteration (all items)
    int b_mask = -1;

    // one item: the key loop of the one-item kernel, phases taken from `u`
    bool guard_armed = false;                           // a reader's arrival on the merge-buffer guard is pending for this writer
    auto item = [&](auto split_tag, int n, bool more_of_kind) {
      constexpr bool kSplit = decltype(split_tag)::value;
      constexpr int j_step = kSplit ? 2 : 1;
      const int j0 = kSplit ? t : 0;
      const int n_half = half ? n_hi : n_tiles;
      const int n_mine = n_half > j0 ? (n_half - j0 + j_step - 1) / j_step : 0;
      const float* mrow = nullptr;
      if (key_mask != nullptr) {
        const int idx = blockIdx.x + n * gridDim.x;
        const int b = (idx % BH) / H;
        mrow = key_mask + static_cast<size_t>(b) * Lk;
        if (mask_words && b != b_mask) {
          // key_padding_mask of this sample as a bitmask, one private copy per warp (no cross-warp synchronisation at
          // item boundaries); eight independent loads in flight per lane
          const int n_words = (Lk + 31) / 32;
          for (int w0 = 0; w0 < n_words; w0 += 8) {
            float mv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int kv = (w0 + e) * 32 + lane;
              mv[e] = (w0 + e < n_words && kv < Lk) ? __ldg(mrow + kv) : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t m = __ballot_sync(0xffffffffu, mv[e] != 0.f);
              if (lane == 0 && w0 + e < n_words) kmask[w0 + e] = m;
            }
          }
          __syncwarp();
          b_mask = b;
        }
      }
      float m_ref = -INFINITY;
      float2 l2 = make_float2(0.f, 0.f);
      uint32_t s_ready = 0;
      for (int i = 0; i < n_mine; ++i, u ^= 1u) {
        const int kv0 = (j0 + j_step * i) * BKV + half * HALF;
        if (!s_ready) mbar_wait(&bars->s_full[g], u);
        tcgen05_fence_after();
        uint32_t s[HALF];
        tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_free[g]);

        float mx;
        bool masked = false;
        uint32_t words[2];
        if (mask_words) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int w = (kv0 >> 5) + c;
            words[c] = w * 32 < Lk ? kmask[w] : 0u;
            masked |= words[c] != 0xffffffffu;
          }
        } else if (mrow != nullptr || kv0 + HALF > Lk) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int kv = kv0 + c * 32 + lane;
            const bool ok = kv < Lk && (mrow == nullptr || __ldg(mrow + kv) != 0.f);
            words[c] = __ballot_sync(0xffffffffu, ok);
            masked |= words[c] != 0xffffffffu;
          }
        }
        if (masked) mx = half_row_max<true>(s, words);
        else mx = half_row_max<false>(s, words);

        if (i == 0) {
          m_ref = mx;
        } else {
          const bool need = mx > m_ref + RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2_approx(m_ref - mx) : 1.0f;
            mbar_wait(&bars->o_full[g], u ^ 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(t_o + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st_32x32b_x16(t_o + c * 16, o);
            }
            tmem_st_wait();
            l2.x *= alpha; l2.y *= alpha;
            if (need) m_ref = mx;
          }
        }
        const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
        if (i > 0)
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 pp_of, [%0], %1;" ::"r"(a_ofull), "r"(u ^ 1u) : "memory");

        const float2 neg_m = make_float2(-m_use, -m_use);
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < HALF; e += 4) {
          const float2 x0 = __fadd2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), neg_m);
          const float2 x1 = __fadd2_rn(make_float2(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])), neg_m);
          const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
          const float2 p1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
          la = __fadd2_rn(la, p0);
          lb = __fadd2_rn(lb, p1);
          s[e >> 1] = pack_bf16x2(p0.x, p0.y);
          s[(e >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
        }
        l2 = __fadd2_rn(l2, __fadd2_rn(la, lb));

        if (i > 0) {
          uint32_t ok;
          asm volatile("selp.u32 %0, 1, 0, pp_of;" : "=r"(ok));
          if (!ok) mbar_wait(&bars->o_full[g], u ^ 1u);
          tcgen05_fence_after();
        }
        if (i + 1 < n_mine)
          asm volatile("mbarrier.test_wait.parity.shared::cta.b64 pp_sf, [%0], %1;" ::"r"(a_sfull), "r"(u ^ 1u) : "memory");
        tmem_st_32x32b_x32(t_p, &s[0]);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready[g]);
        s_ready = 0;
        if (i + 1 < n_mine) asm volatile("selp.u32 %0, 1, 0, pp_sf;" : "=r"(s_ready));
      }

      // ---- item epilogue: O_g is complete once the last P V has landed; the accumulator is handed back to the issuer
      // (the next item's first P V overwrites it) as soon as it is in registers
      uint32_t o[DH];
      if (n_mine > 0) {
        mbar_wait(&bars->o_full[g], u ^ 1u);
        tcgen05_fence_after();
        tmem_ld_32x32b_x32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[g]);
      } else {
#pragma unroll
        for (int e = 0; e < DH; ++e) o[e] = 0u;
      }
      const float l_mine = l2.x + l2.y;
      float* cmb_base = reinterpret_cast<float*>(smem + attn_p::OFF_CMB);
      const bool writer = kSplit ? (g != 0) : (half == 1);
      if (writer) {
        // the slot's previous contents (the previous item of the same kind) have been read: the reader ARRIVED on the
        // guard barrier when it was through, long ago by now -- the writer never waits for the reader's epilogue, so the two
        // warpgroups of a tile do not fall into lockstep at item boundaries (lockstep halves the MUFU overlap)
        if (guard_armed) {
          if (kSplit) asm volatile("bar.sync 7, 512;" ::: "memory");
          else asm volatile("bar.sync %0, 256;" ::"r"(5 + t) : "memory");
        }
        float* cmb = cmb_base + ((kSplit ? g - 1 : t) * BQ + r) * CMB_STRIDE;
        cmb[0] = m_ref;
        cmb[1] = l_mine;
#pragma unroll
        for (int e = 0; e < DH; ++e) cmb[2 + e] = __uint_as_float(o[e]);
      }
      // partial results ready: the writers only ARRIVE (PTX's producer / consumer form of the named barrier) and move on
      // to the next item; the reader waits
      if (writer) {
        if (kSplit) asm volatile("bar.arrive 4, 512;" ::: "memory");
        else asm volatile("bar.arrive %0, 256;" ::"r"(1 + t) : "memory");
      } else {
        if (kSplit) asm volatile("bar.sync 4, 512;" ::: "memory");
        else asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      }
      if (!writer) {
        constexpr int n_parts = kSplit ? 3 : 1;
        const int first = kSplit ? 0 : t;
        float m = m_ref;
#pragma unroll
        for (int pi = 0; pi < n_parts; ++pi) m = fmaxf(m, cmb_base[((first + pi) * BQ + r) * CMB_STRIDE]);
        const float m_safe = m == -INFINITY ? 0.f : m;
        const float a_own = ex2_approx(m_ref - m_safe);
        float l_tot = a_own * l_mine;
        float acc[DH];
#pragma unroll
        for (int e = 0; e < DH; ++e) acc[e] = __uint_as_float(o[e]) * a_own;
#pragma unroll
        for (int pi = 0; pi < n_parts; ++pi) {
          const float* cmb = cmb_base + ((first + pi) * BQ + r) * CMB_STRIDE;
          const float a_p = ex2_approx(cmb[0] - m_safe);
          l_tot = fmaf(a_p, cmb[1], l_tot);
#pragma unroll
          for (int e = 0; e < DH; ++e) acc[e] = fmaf(a_p, cmb[2 + e], acc[e]);
        }
        const float inv = 1.0f / l_tot;
        // item coordinates are only needed here: recomputed instead of being kept live through the key loop
        const int idx = blockIdx.x + n * gridDim.x;
        const int qp = idx / BH, bh = idx - qp * BH;
        const int q = qp * (2 * BQ) + (kSplit ? 0 : t) * BQ + r;
        if (q < Lq) {
          const int b = bh / H, h = bh - b * H;
          if (kLse) lse[static_cast<size_t>(bh) * lse_pitch + q] = m_safe + __log2f(l_tot);
          uint4* op = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * Lq + q) * ldo + h * DH);
#pragma unroll
          for (int e = 0; e < DH / 8; ++e) {
            uint4 w;
            w.x = pack_bf16x2(acc[8 * e + 0] * inv, acc[8 * e + 1] * inv);
            w.y = pack_bf16x2(acc[8 * e + 2] * inv, acc[8 * e + 3] * inv);
            w.z = pack_bf16x2(acc[8 * e + 4] * inv, acc[8 * e + 5] * inv);
            w.w = pack_bf16x2(acc[8 * e + 6] * inv, acc[8 * e + 7] * inv);
            op[e] = w;
          }
        }
      }
      // the merge buffer is rewritten by the next item's writers only after its readers are through (see above)
      if (!writer && more_of_kind) {
        if (kSplit) asm volatile("bar.arrive 7, 512;" ::: "memory");
        else asm volatile("bar.arrive %0, 256;" ::"r"(5 + t) : "memory");
      }
      guard_armed = more_of_kind;
    };

    {
      // start offsets of the first item: g0, g2, g1, g3 a quarter period apart (see the one-item kernel)
      const int slot = n_tiles >= 6 ? half * 2 + t : 0;
      const long long t_start = clock64();
      while (clock64() - t_start < static_cast<long long>(slot) * STAGGER_CLK) {}
    }
    bool was_split = false;
    for (int n = 0; n < my_items; ++n) {
      const bool single = last_single && pair_of(n) == nx - 1;
      const bool split = single && split_single;
      // the NEXT item of this CTA has the same writer / reader roles (items are pair-major: kinds change at most once)
      const bool next_single = n + 1 < my_items && last_single && pair_of(n + 1) == nx - 1;
      const bool more = n + 1 < my_items && (next_single == single);
      if (split && !was_split && n > 0) asm volatile("bar.sync 8, 512;" ::: "memory");   // roles change: everyone is through with the two-tile items
      was_split = split;
      if (single && !split && t == 1) continue;
      if (split) item(std::true_type{}, n, more);
      else item(std::false_type{}, n, more);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 17) {
    tcgen05_fence_after();
    tmem_dealloc<attn::TMEM_COLS>(tmem_base_of(bars));
  }
}

int launch_attention_tc(const AttnArgs& a, cudaStream_t stream) {
  using namespace attn;
  if (a.B <= 0 || a.H <= 0 || a.Lq <= 0 || a.Lk <= 0) return svol_fail(SVOL_ERR_SHAPE, "attention: bad sizes");
  if (a.vt_pitch < a.Lk || a.vt_pitch % 8 != 0 || a.ldq % 8 || a.ldk % 8 || a.ldo % 8)
    return svol_fail(SVOL_ERR_SHAPE, "attention: pitches must be multiples of 8 elements and vt_pitch >= Lk");
  {      // short key sequences without a mask (the object queries' self-attention): attn_small.cu
    const int rc_small = launch_attention_small(a, stream);
    if (rc_small >= 0) return rc_small;
  }
  CUtensorMap tmQ, tmK, tmVt;
  int rc = make_tensor_map_2d(&tmQ, a.q, a.H * DH, static_cast<int64_t>(a.B) * a.Lq, a.ldq, DH, BQ, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmK, a.k, a.H * DH, static_cast<int64_t>(a.B) * a.Lk, a.ldk, DH, BKV, 64);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmVt, a.vt, a.vt_pitch, static_cast<int64_t>(a.B) * a.H * DH, a.vt_pitch, 64, DH, 128);
  if (rc) return rc;
  CUtensorMap tmQ64;       // 64-row box of Q: single-tile CTAs with at most 64 query rows load their rows twice
  rc = make_tensor_map_2d(&tmQ64, a.q, a.H * DH, static_cast<int64_t>(a.B) * a.Lq, a.ldq, DH, BQ / 2, 64);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return svol_fail_cuda(e, "attention: cudaFuncSetAttribute");
    configured = true;
  }
  if (a.lse != nullptr && a.lse_pitch < a.Lq) return svol_fail(SVOL_ERR_SHAPE, "attention: lse_pitch < Lq");
  // Persistent CTAs (one per SM walking a work list): OFF by default -- measured slower, see the kernel's header.
  // SVOL_ATTN_PERSISTENT=1 enables it when there is more than one wave of work, =2 always (tests, experiments).
  const char* env_p = getenv("SVOL_ATTN_PERSISTENT");          // read per launch: tests flip it between calls
  const int persistent = env_p ? atoi(env_p) : 0;
  const int n_items = ((a.Lq + 2 * BQ - 1) / (2 * BQ)) * a.H * a.B;
  if (persistent == 2 || (persistent && n_items > sm_count())) {      // 2: force (experiments)
    static bool configured_p = false;
    if (!configured_p) {
      cudaError_t e = cudaFuncSetAttribute(attention_p_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_p::SMEM_BYTES);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_p_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_p::SMEM_BYTES);
      if (e != cudaSuccess) return svol_fail_cuda(e, "attention (persistent): cudaFuncSetAttribute");
      configured_p = true;
    }
    const int grid_p = n_items < sm_count() ? n_items : sm_count();
    if (a.lse != nullptr)
      attention_p_kernel<true><<<grid_p, THREADS, attn_p::SMEM_BYTES, stream>>>(tmQ, tmK, tmVt, a.key_mask, reinterpret_cast<__nv_bfloat16*>(a.out),
                                                                              a.lse, a.lse_pitch, a.B, a.H, a.Lq, a.Lk, a.ldo);
    else
      attention_p_kernel<false><<<grid_p, THREADS, attn_p::SMEM_BYTES, stream>>>(tmQ, tmK, tmVt, a.key_mask, reinterpret_cast<__nv_bfloat16*>(a.out),
                                                                               nullptr, 0, a.B, a.H, a.Lq, a.Lk, a.ldo);
    return svol_check_launch("attention_tc (persistent)");
  }
  dim3 grid((a.Lq + 2 * BQ - 1) / (2 * BQ), a.H, a.B);
  if (a.lse != nullptr) {
    attention_tc_kernel<true, false><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmVt, tmQ64, a.key_mask, reinterpret_cast<__nv_bfloat16*>(a.out),
                                                                           a.lse, a.lse_pitch, a.H, a.Lq, a.Lk, a.ldo, n_items, nullptr);
    return svol_check_launch("attention_tc");
  }
  // Inference.  SVOL_ATTN_LOOP=1 and more work items than SMs: one looping CTA per SM (kLoop, see the kernel; off by
  // default -- measured slower).  Every such launch gets its own item counter out of a pool (launches on different streams
  // run concurrently; a CUDA graph keeps the slot it was captured with); a launch leaves its counter at zero.
  const char* env_l = getenv("SVOL_ATTN_LOOP");                // read per launch (A/B measurements, tests)
  bool loop = (env_l ? atoi(env_l) != 0 : false) && n_items > sm_count();
  cudaError_t e;
  // a launch being captured into a CUDA graph keeps its counter for the life of the graph: those slots are never handed
  // out again (2048 captured attention launches per process; beyond that the one-CTA-per-item kernel is used), eager
  // launches rotate through the other half of the pool
  constexpr unsigned int kSlots = 4096, kGraphSlots = 2048;
  static std::atomic<unsigned int> next_slot{0}, next_graph_slot{0};
  unsigned int slot = 0;
  if (loop) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusNone;
    if (cap == cudaStreamCaptureStatusActive) {
      slot = next_graph_slot.fetch_add(1);
      if (slot >= kGraphSlots) loop = false;
    } else {
      slot = kGraphSlots + next_slot.fetch_add(1) % (kSlots - kGraphSlots);
    }
  }
  if (loop) {
    static unsigned int* counters = nullptr;
    static std::mutex mu;
    {
      std::lock_guard<std::mutex> lock(mu);
      if (counters == nullptr) {
        // (allocation and clearing are synchronous, legacy-stream operations: not captured even if `stream` is capturing)
        cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
        cudaThreadExchangeStreamCaptureMode(&mode);
        e = cudaMalloc(&counters, kSlots * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemset(counters, 0, kSlots * sizeof(unsigned int));
        cudaThreadExchangeStreamCaptureMode(&mode);
        if (e != cudaSuccess) { counters = nullptr; return svol_fail_cuda(e, "attention: item counters"); }
      }
    }
    unsigned int* counter = counters + slot;
    // (SVOL_ATTN_LOOP=2, measurements: the looping kernel with one CTA per item, i.e. its code without any item hand-over)
    const int grid_l = (env_l && atoi(env_l) == 2) ? n_items : sm_count();
    e = launch_kernel_pdl(attention_tc_kernel<false, true>, dim3(grid_l), dim3(THREADS), SMEM_BYTES, stream, tmQ, tmK, tmVt, tmQ64,
                          a.key_mask, reinterpret_cast<__nv_bfloat16*>(a.out), nullptr, 0, a.H, a.Lq, a.Lk, a.ldo, n_items, counter);
  } else {       // programmatic dependent launch: common.cuh
    e = launch_kernel_pdl(attention_tc_kernel<false, false>, grid, dim3(THREADS), SMEM_BYTES, stream, tmQ, tmK, tmVt, tmQ64, a.key_mask,
                          reinterpret_cast<__nv_bfloat16*>(a.out), nullptr, 0, a.H, a.Lq, a.Lk, a.ldo, n_items, nullptr);
  }
  if (e != cudaSuccess) return svol_fail_cuda(e, "attention_tc launch");
  return svol_check_launch("attention_tc");
}

}  // namespace svol

#ifdef SVOL_ATTN_TRACE
extern "C" int svol_debug_attn_cta_timeline(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, svol::g_attn_cta, sizeof(svol::g_attn_cta)));
}
extern "C" int svol_debug_attn_trace(long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, svol::g_attn_trace, sizeof(svol::g_attn_trace)));
}
#endif
