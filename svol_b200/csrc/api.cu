// extern "C" entry points of libsvol_b200.so, error plumbing and TMA descriptor construction.
#include <cstdio>
#include <cstring>
#include <mutex>

#include "svol_internal.h"

namespace svol {

static thread_local char g_err[512] = "";

int svol_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int svol_fail_cuda(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return static_cast<int>(e);
}
int svol_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return svol_fail_cuda(e, what);
  return SVOL_OK;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

// cuTensorMapEncodeTiled is a driver API; resolve it through the runtime so the library does not
// link against libcuda (absent on the build machine).
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld,
                       int box_inner, int box_outer, int swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return svol_fail(SVOL_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 != 0)
    return svol_fail(SVOL_ERR_SHAPE, "tensor map: base must be 16-byte aligned and the row pitch a multiple of 8 elements");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d (inner %lld outer %lld ld %lld box %dx%d)",
             static_cast<int>(r), (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer);
    return svol_fail(SVOL_ERR_DRIVER, msg);
  }
  return SVOL_OK;
}

// rowwise.cu
int launch_layernorm_f32_to_bf16(const float*, const float*, const float*, svol_bf16*, int, int, float, float, const long long*, int,
                                 cudaStream_t);
int launch_ln_linear_f32(const float*, const float*, const float*, const float*, const float*, int, float*, int, int, int,
                         float, float, const long long*, int, cudaStream_t);
int launch_layernorm_bf16_to_bf16(const svol_bf16*, const float*, const float*, svol_bf16*, int, int, float, cudaStream_t);
int launch_layernorm_nchw_to_bf16(const float*, const float*, const float*, svol_bf16*, int, int, int, float, cudaStream_t);
int launch_posenc_sine(const float*, float*, int, int, int, cudaStream_t);
int launch_posenc_theta(const float*, float*, int, int, cudaStream_t);
int launch_add_pos_bf16(const float*, const float*, svol_bf16*, int, int, int, cudaStream_t);
int launch_gate_vectors(const float*, const float*, const float*, float*, int, int, int, cudaStream_t);
int launch_gate_scores(const svol_bf16*, const float*, float*, int, int, int, int, cudaStream_t);
int launch_gate_apply(const svol_bf16*, const float*, const float*, const float*, const float*, svol_bf16*, svol_bf16*,
                      float*, int, int, int, int, float, bool, cudaStream_t);
int launch_heads(const svol_bf16*, const svol_bf16*, const float*, const float*, const float*, const float*, float*,
                 float*, int, int, cudaStream_t);
int launch_postprocess(const float*, const float*, float*, int32_t*, int, int, int, cudaStream_t);
// train.cu / attn_bwd_tc.cu
int launch_layernorm_bf16(const svol_bf16*, const float*, const float*, svol_bf16*, svol_bf16*, const float*, int, const float*,
                          int, int, float, float, const long long*, int, cudaStream_t);
int launch_layernorm_backward(const void*, int, const float*, const svol_bf16*, const svol_bf16*, const svol_bf16*, const float*,
                              svol_bf16*, float*, float*, float*, int, int, float, float, const long long*, int, cudaStream_t);
int launch_gelu_bf16(const svol_bf16*, svol_bf16*, long long, cudaStream_t);
int launch_act_backward(const svol_bf16*, const svol_bf16*, svol_bf16*, long long, int, cudaStream_t);
int launch_transpose_bf16(const svol_bf16*, int, int, int, svol_bf16*, int, float*, cudaStream_t);
int launch_colsum_bf16(const svol_bf16*, int, int, int, float*, cudaStream_t);
int launch_attention_backward_tc(const svol_attn_bwd_args&, cudaStream_t);
int launch_heads_backward(const svol_bf16*, const svol_bf16*, const float*, const float*, const float*, const float*, const float*,
                          svol_bf16*, svol_bf16*, float*, float*, float*, float*, int, int, cudaStream_t);
int launch_gate_fused(const svol_bf16*, const float*, const float*, const float*, const float*, svol_bf16*, svol_bf16*, float*,
                      float*, int, int, int, int, float, cudaStream_t);
int gate_fused_supported(int);
int launch_gate_backward(const svol_bf16*, const float*, const float*, const float*, const svol_bf16*, svol_bf16*, float*, float*,
                         int, int, int, int, cudaStream_t);
int launch_gate_vectors_backward(const float*, const float*, const float*, const float*, float*, float*, float*, int, int, int,
                                 cudaStream_t);
int launch_ln_linear_f32_backward(const float*, const float*, const float*, const float*, const float*, const float*, int, float*,
                                  float*, float*, float*, float*, int, int, int, float, float, const long long*, int, cudaStream_t);
int launch_batch_sum(const svol_bf16*, float*, int, int, int, cudaStream_t);
int launch_accum_bf16(const svol_bf16*, float*, long long, float, int, cudaStream_t);
int launch_adamw(float*, const float*, float*, float*, long long, float, float, float, float, float, int, float, cudaStream_t);
int launch_adamw_segments(float*, const float*, float*, float*, long long, const long long*, const int*, int,
                          const svol_adamw_group*, int, float, cudaStream_t);
int launch_pack_weights(const svol_pack_job*, int, cudaStream_t);
// evaluate.cu
int launch_eval_max_iou(const float*, const int*, const float*, const int*, int, int, int, double*, double*, cudaStream_t);
int launch_eval_average_precision(const float*, const int*, const float*, const int*, const int*, int, int, int, int, double*,
                                  cudaStream_t);

}  // namespace svol

using namespace svol;

#define SVOL_STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define SVOL_REQUIRE(p)                                                   \
  do {                                                                    \
    if (!(p)) return svol_fail(SVOL_ERR_NULL, "required pointer is NULL: " #p); \
  } while (0)

extern "C" {

int svol_abi_version(void) { return SVOL_ABI_VERSION; }
const char* svol_last_error(void) { return g_err; }

int svol_sizeof_args(int which) {
  switch (which) {
    case 0: return static_cast<int>(sizeof(svol_gemm_args));
    case 1: return static_cast<int>(sizeof(svol_attn_args));
    case 2: return static_cast<int>(sizeof(svol_match_args));
    case 3: return static_cast<int>(sizeof(svol_criterion_args));
    case 4: return static_cast<int>(sizeof(svol_gemm_epilogue));
    case 5: return static_cast<int>(sizeof(svol_ffn_args));
    case 6: return static_cast<int>(sizeof(svol_attn_bwd_args));
    case 7: return static_cast<int>(sizeof(svol_pack_job));
    default: return -1;
  }
}

int svol_device_check(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return svol_fail_cuda(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return svol_fail_cuda(e, "cudaDeviceGetAttribute");
  if (major != 10) return svol_fail(SVOL_ERR_DEVICE, "svol_b200 needs a compute capability 10.x (B200) device");
  return SVOL_OK;
}

int svol_gemm_bf16(const svol_gemm_args* a, void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(a->A); SVOL_REQUIRE(a->W);
  return launch_gemm_bf16_tc(*a, SVOL_STREAM(stream));
}
int svol_gemm_bf16_plain(const svol_gemm_args* a, void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(a->A); SVOL_REQUIRE(a->W);
  return launch_gemm_bf16_plain(*a, SVOL_STREAM(stream));
}
int svol_ffn_bf16(const svol_ffn_args* a, void* stream) {
  SVOL_REQUIRE(a);
  return launch_ffn_tc(*a, SVOL_STREAM(stream));
}
int svol_attention_bf16(const svol_attn_args* a, void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(a->q); SVOL_REQUIRE(a->k); SVOL_REQUIRE(a->vt); SVOL_REQUIRE(a->out);
  return launch_attention_tc(*a, SVOL_STREAM(stream));
}
int svol_attention_bf16_plain(const svol_attn_args* a, void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(a->q); SVOL_REQUIRE(a->k); SVOL_REQUIRE(a->vt); SVOL_REQUIRE(a->out);
  return launch_attention_plain(*a, SVOL_STREAM(stream));
}

int svol_layernorm_f32_to_bf16(const float* x, const float* w, const float* b, svol_bf16* y, int32_t rows,
                               int32_t cols, float eps, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(y);
  return launch_layernorm_f32_to_bf16(x, w, b, y, rows, cols, eps, 0.f, nullptr, 0, SVOL_STREAM(stream));
}
int svol_layernorm_f32_to_bf16_dropout(const float* x, const float* w, const float* b, svol_bf16* y, int32_t rows, int32_t cols,
                                       float eps, float drop_p, const int64_t* seed, int32_t site, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(y);
  return launch_layernorm_f32_to_bf16(x, w, b, y, rows, cols, eps, drop_p, reinterpret_cast<const long long*>(seed), site,
                                      SVOL_STREAM(stream));
}
int svol_ln_linear_f32(const float* x, const float* lw, const float* lb, const float* w, const float* b, int32_t relu,
                       float* y, int32_t rows, int32_t in_dim, int32_t out_dim, float eps, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(lw); SVOL_REQUIRE(lb); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(y);
  return launch_ln_linear_f32(x, lw, lb, w, b, relu, y, rows, in_dim, out_dim, eps, 0.f, nullptr, 0, SVOL_STREAM(stream));
}
int svol_ln_linear_f32_dropout(const float* x, const float* lw, const float* lb, const float* w, const float* b, int32_t relu,
                               float* y, int32_t rows, int32_t in_dim, int32_t out_dim, float eps, float drop_p,
                               const int64_t* seed, int32_t site, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(lw); SVOL_REQUIRE(lb); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(y);
  return launch_ln_linear_f32(x, lw, lb, w, b, relu, y, rows, in_dim, out_dim, eps, drop_p, reinterpret_cast<const long long*>(seed),
                              site, SVOL_STREAM(stream));
}
int svol_layernorm_bf16_to_bf16(const svol_bf16* x, const float* w, const float* b, svol_bf16* y, int32_t rows, int32_t cols,
                                float eps, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(y);
  return launch_layernorm_bf16_to_bf16(x, w, b, y, rows, cols, eps, SVOL_STREAM(stream));
}
int svol_layernorm_nchw_to_bf16(const float* x, const float* w, const float* b, svol_bf16* y, int32_t frames, int32_t channels,
                                int32_t spatial, float eps, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(y);
  return launch_layernorm_nchw_to_bf16(x, w, b, y, frames, channels, spatial, eps, SVOL_STREAM(stream));
}
int svol_posenc_sine(const float* mask, float* pos, int32_t B, int32_t L, int32_t d, void* stream) {
  SVOL_REQUIRE(mask); SVOL_REQUIRE(pos);
  return launch_posenc_sine(mask, pos, B, L, d, SVOL_STREAM(stream));
}
int svol_posenc_theta(const float* mask, float* theta, int32_t B, int32_t L, void* stream) {
  SVOL_REQUIRE(mask); SVOL_REQUIRE(theta);
  return launch_posenc_theta(mask, theta, B, L, SVOL_STREAM(stream));
}
int svol_add_pos_bf16(const float* x, const float* pos, svol_bf16* out, int32_t rows, int32_t cols, int32_t mod,
                      void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(out);
  return launch_add_pos_bf16(x, pos, out, rows, cols, mod, SVOL_STREAM(stream));
}
int svol_gate_vectors(const float* sketch, const float* w, const float* b, float* u, int32_t B, int32_t d, int32_t H,
                      void* stream) {
  SVOL_REQUIRE(sketch); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(u);
  return launch_gate_vectors(sketch, w, b, u, B, d, H, SVOL_STREAM(stream));
}
int svol_gate_scores(const svol_bf16* xpos, const float* u, float* scores, int32_t B, int32_t L, int32_t d, int32_t H,
                     void* stream) {
  SVOL_REQUIRE(xpos); SVOL_REQUIRE(u); SVOL_REQUIRE(scores);
  return launch_gate_scores(xpos, u, scores, B, L, d, H, SVOL_STREAM(stream));
}
int svol_gate_apply(const svol_bf16* x, const float* scores, const float* lw, const float* lb, const float* pos,
                    svol_bf16* mem, svol_bf16* mem_pos, float* att_out, int32_t B, int32_t L, int32_t d, int32_t H,
                    float eps, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(scores); SVOL_REQUIRE(lw); SVOL_REQUIRE(lb); SVOL_REQUIRE(pos); SVOL_REQUIRE(mem);
  SVOL_REQUIRE(mem_pos);
  return launch_gate_apply(x, scores, lw, lb, pos, mem, mem_pos, att_out, B, L, d, H, eps, false, SVOL_STREAM(stream));
}
int svol_gate_apply_theta(const svol_bf16* x, const float* scores, const float* lw, const float* lb, const float* theta,
                          svol_bf16* mem, svol_bf16* mem_pos, float* att_out, int32_t B, int32_t L, int32_t d, int32_t H,
                          float eps, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(scores); SVOL_REQUIRE(lw); SVOL_REQUIRE(lb); SVOL_REQUIRE(theta); SVOL_REQUIRE(mem);
  SVOL_REQUIRE(mem_pos);
  return launch_gate_apply(x, scores, lw, lb, theta, mem, mem_pos, att_out, B, L, d, H, eps, true, SVOL_STREAM(stream));
}
int svol_gate_fused_supported(int32_t L) { return gate_fused_supported(L); }
int svol_gate_fused(const svol_bf16* x, const float* u, const float* lw, const float* lb, const float* theta, svol_bf16* mem,
                    svol_bf16* mem_pos, float* att_out, float* scores_out, int32_t B, int32_t L, int32_t d, int32_t H, float eps,
                    void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(u); SVOL_REQUIRE(lw); SVOL_REQUIRE(lb); SVOL_REQUIRE(theta); SVOL_REQUIRE(mem);
  SVOL_REQUIRE(mem_pos);
  return launch_gate_fused(x, u, lw, lb, theta, mem, mem_pos, att_out, scores_out, B, L, d, H, eps, SVOL_STREAM(stream));
}
int svol_heads(const svol_bf16* hs, const svol_bf16* h2, const float* wc, const float* bc, const float* wb,
               const float* bb, float* logits, float* boxes, int32_t rows, int32_t d, void* stream) {
  SVOL_REQUIRE(hs); SVOL_REQUIRE(h2); SVOL_REQUIRE(wc); SVOL_REQUIRE(bc); SVOL_REQUIRE(wb); SVOL_REQUIRE(bb);
  SVOL_REQUIRE(logits); SVOL_REQUIRE(boxes);
  return launch_heads(hs, h2, wc, bc, wb, bb, logits, boxes, rows, d, SVOL_STREAM(stream));
}

int svol_match(const svol_match_args* a, void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(a->logits); SVOL_REQUIRE(a->boxes); SVOL_REQUIRE(a->tgt_boxes); SVOL_REQUIRE(a->tgt_off);
  SVOL_REQUIRE(a->match_off); SVOL_REQUIRE(a->cost_off); SVOL_REQUIRE(a->pred_idx);
  SVOL_REQUIRE(a->tgt_idx); SVOL_REQUIRE(a->status);
  return launch_match(*a, SVOL_STREAM(stream));
}
int svol_lsap_f32(const float* cost, const int64_t* cost_off, const int32_t* shape, int32_t n_problems, int32_t max_small,
                  int32_t max_big, int32_t max_entries, int64_t* rows_out, int64_t* cols_out, const int64_t* out_off, int32_t* status,
                  int32_t solver, void* stream) {
  SVOL_REQUIRE(cost); SVOL_REQUIRE(cost_off); SVOL_REQUIRE(shape); SVOL_REQUIRE(rows_out); SVOL_REQUIRE(cols_out);
  SVOL_REQUIRE(out_off); SVOL_REQUIRE(status);
  return launch_lsap_f32(cost, cost_off, shape, n_problems, max_small, max_big, max_entries, rows_out, cols_out, out_off, status, solver,
                         SVOL_STREAM(stream));
}
int svol_match_localize(int64_t* tgt_idx, const int32_t* video_match_off, int32_t NL, int32_t B, int32_t K,
                        void* stream) {
  SVOL_REQUIRE(tgt_idx); SVOL_REQUIRE(video_match_off);
  return launch_match_localize(tgt_idx, video_match_off, NL, B, K, SVOL_STREAM(stream));
}
int64_t svol_criterion_scratch_bytes(int32_t NL, int32_t B) {
  return (static_cast<int64_t>(NL) * B * 4 * 8 + static_cast<int64_t>(NL) * 4 + 15) / 16 * 16;
}
int svol_criterion(const svol_criterion_args* a, void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(a->logits); SVOL_REQUIRE(a->boxes); SVOL_REQUIRE(a->tgt_boxes); SVOL_REQUIRE(a->pred_idx);
  SVOL_REQUIRE(a->tgt_idx); SVOL_REQUIRE(a->video_tgt_off); SVOL_REQUIRE(a->losses);
  return launch_criterion(*a, SVOL_STREAM(stream));
}
int svol_criterion_backward(const svol_criterion_args* a, const float* grad_w, float* grad_logits, float* grad_boxes,
                            void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(grad_w); SVOL_REQUIRE(grad_logits); SVOL_REQUIRE(grad_boxes);
  return launch_criterion_backward(*a, grad_w, grad_logits, grad_boxes, SVOL_STREAM(stream));
}
int svol_postprocess(const float* logits, const float* boxes, float* out, int32_t* order, int32_t B, int32_t Q,
                     int32_t q_per_frame, void* stream) {
  SVOL_REQUIRE(logits); SVOL_REQUIRE(boxes); SVOL_REQUIRE(out); SVOL_REQUIRE(order);
  return launch_postprocess(logits, boxes, out, order, B, Q, q_per_frame, SVOL_STREAM(stream));
}


// ---- training step
int svol_layernorm_bf16(const svol_bf16* z, const float* w, const float* b, svol_bf16* y, svol_bf16* y_pos, const float* pos,
                        int32_t pos_mod, const float* theta, int32_t rows, int32_t cols, float eps, float drop_p,
                        const int64_t* seed, int32_t site, void* stream) {
  SVOL_REQUIRE(z); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(y);
  return launch_layernorm_bf16(z, w, b, y, y_pos, pos, pos_mod, theta, rows, cols, eps, drop_p,
                               reinterpret_cast<const long long*>(seed), site, SVOL_STREAM(stream));
}
int svol_layernorm_backward(const void* z, int32_t z_is_f32, const float* att, const svol_bf16* dy1, const svol_bf16* dy2,
                            const svol_bf16* dy3, const float* gamma, svol_bf16* dx, float* datt, float* dgamma, float* dbeta,
                            int32_t rows, int32_t cols, float eps, float drop_p, const int64_t* seed, int32_t site, void* stream) {
  SVOL_REQUIRE(z); SVOL_REQUIRE(dy1); SVOL_REQUIRE(gamma); SVOL_REQUIRE(dgamma); SVOL_REQUIRE(dbeta);
  if (att && !datt) return svol_fail(SVOL_ERR_NULL, "layernorm_backward: att needs datt");
  return launch_layernorm_backward(z, z_is_f32, att, dy1, dy2, dy3, gamma, dx, datt, dgamma, dbeta, rows, cols, eps, drop_p,
                                   reinterpret_cast<const long long*>(seed), site, SVOL_STREAM(stream));
}
int svol_gelu_bf16(const svol_bf16* x, svol_bf16* y, int64_t n, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(y);
  return launch_gelu_bf16(x, y, n, SVOL_STREAM(stream));
}
int svol_act_backward(const svol_bf16* dy, const svol_bf16* saved, svol_bf16* out, int64_t n, int32_t mode, void* stream) {
  SVOL_REQUIRE(dy); SVOL_REQUIRE(saved); SVOL_REQUIRE(out);
  return launch_act_backward(dy, saved, out, n, mode, SVOL_STREAM(stream));
}
int svol_transpose_bf16(const svol_bf16* in, int32_t ld_in, int32_t rows, int32_t cols, svol_bf16* out, int32_t ld_out,
                        float* colsum, void* stream) {
  SVOL_REQUIRE(in); SVOL_REQUIRE(out);
  return launch_transpose_bf16(in, ld_in, rows, cols, out, ld_out, colsum, SVOL_STREAM(stream));
}
int svol_colsum_bf16(const svol_bf16* in, int32_t ld_in, int32_t rows, int32_t cols, float* colsum, void* stream) {
  SVOL_REQUIRE(in); SVOL_REQUIRE(colsum);
  return launch_colsum_bf16(in, ld_in, rows, cols, colsum, SVOL_STREAM(stream));
}
int svol_attention_backward_bf16(const svol_attn_bwd_args* a, void* stream) {
  SVOL_REQUIRE(a); SVOL_REQUIRE(a->q); SVOL_REQUIRE(a->k); SVOL_REQUIRE(a->v); SVOL_REQUIRE(a->kt); SVOL_REQUIRE(a->qt);
  SVOL_REQUIRE(a->o); SVOL_REQUIRE(a->d_o); SVOL_REQUIRE(a->d_ot); SVOL_REQUIRE(a->lse); SVOL_REQUIRE(a->delta);
  SVOL_REQUIRE(a->dq); SVOL_REQUIRE(a->dk); SVOL_REQUIRE(a->dv);
  return launch_attention_backward_tc(*a, SVOL_STREAM(stream));
}
int svol_heads_backward(const svol_bf16* hs, const svol_bf16* h2, const float* wc, const float* wb, const float* boxes,
                        const float* dlogits, const float* dboxes, svol_bf16* dhs_cls, svol_bf16* dh2, float* dwc, float* dbc,
                        float* dwb, float* dbb, int32_t rows, int32_t d, void* stream) {
  SVOL_REQUIRE(hs); SVOL_REQUIRE(h2); SVOL_REQUIRE(wc); SVOL_REQUIRE(wb); SVOL_REQUIRE(boxes); SVOL_REQUIRE(dlogits);
  SVOL_REQUIRE(dboxes); SVOL_REQUIRE(dhs_cls); SVOL_REQUIRE(dh2); SVOL_REQUIRE(dwc); SVOL_REQUIRE(dbc); SVOL_REQUIRE(dwb);
  SVOL_REQUIRE(dbb);
  return launch_heads_backward(hs, h2, wc, wb, boxes, dlogits, dboxes, dhs_cls, dh2, dwc, dbc, dwb, dbb, rows, d,
                               SVOL_STREAM(stream));
}
int svol_gate_backward(const svol_bf16* xpos, const float* u, const float* scores, const float* datt, const svol_bf16* dx_in,
                       svol_bf16* dx_out, float* dscores, float* du, int32_t B, int32_t L, int32_t d, int32_t H, void* stream) {
  SVOL_REQUIRE(xpos); SVOL_REQUIRE(u); SVOL_REQUIRE(scores); SVOL_REQUIRE(datt); SVOL_REQUIRE(dx_in); SVOL_REQUIRE(dx_out);
  SVOL_REQUIRE(dscores); SVOL_REQUIRE(du);
  return launch_gate_backward(xpos, u, scores, datt, dx_in, dx_out, dscores, du, B, L, d, H, SVOL_STREAM(stream));
}
int svol_gate_vectors_backward(const float* sketch, const float* w, const float* b, const float* du, float* dw, float* db,
                               float* dsketch, int32_t B, int32_t d, int32_t H, void* stream) {
  SVOL_REQUIRE(sketch); SVOL_REQUIRE(w); SVOL_REQUIRE(b); SVOL_REQUIRE(du); SVOL_REQUIRE(dw); SVOL_REQUIRE(db);
  SVOL_REQUIRE(dsketch);
  return launch_gate_vectors_backward(sketch, w, b, du, dw, db, dsketch, B, d, H, SVOL_STREAM(stream));
}
int svol_ln_linear_f32_backward(const float* x, const float* lw, const float* lb, const float* w, const float* y,
                                const float* dy, int32_t relu, float* dx, float* dlw, float* dlb, float* dw, float* db,
                                int32_t rows, int32_t in_dim, int32_t out_dim, float eps, float drop_p, const int64_t* seed,
                                int32_t site, void* stream) {
  SVOL_REQUIRE(x); SVOL_REQUIRE(lw); SVOL_REQUIRE(lb); SVOL_REQUIRE(w); SVOL_REQUIRE(y); SVOL_REQUIRE(dy); SVOL_REQUIRE(dx);
  SVOL_REQUIRE(dlw); SVOL_REQUIRE(dlb); SVOL_REQUIRE(dw); SVOL_REQUIRE(db);
  return launch_ln_linear_f32_backward(x, lw, lb, w, y, dy, relu, dx, dlw, dlb, dw, db, rows, in_dim, out_dim, eps, drop_p,
                                       reinterpret_cast<const long long*>(seed), site, SVOL_STREAM(stream));
}
int svol_batch_sum(const svol_bf16* g, float* acc, int32_t rows, int32_t cols, int32_t mod, void* stream) {
  SVOL_REQUIRE(g); SVOL_REQUIRE(acc);
  return launch_batch_sum(g, acc, rows, cols, mod, SVOL_STREAM(stream));
}
int svol_accum_bf16(const svol_bf16* src, float* dst, int64_t n, float scale, int32_t accumulate, void* stream) {
  SVOL_REQUIRE(src); SVOL_REQUIRE(dst);
  return launch_accum_bf16(src, dst, n, scale, accumulate, SVOL_STREAM(stream));
}
int svol_eval_max_iou(const float* pred, const int32_t* frame_index, const float* gt, const int32_t* gt_off, int32_t frames,
                      int32_t n_gt, int32_t q_per_frame, double* max1, double* max5, void* stream) {
  SVOL_REQUIRE(pred); SVOL_REQUIRE(frame_index); SVOL_REQUIRE(gt); SVOL_REQUIRE(gt_off); SVOL_REQUIRE(max1); SVOL_REQUIRE(max5);
  return launch_eval_max_iou(pred, frame_index, gt, gt_off, frames, n_gt, q_per_frame, max1, max5, SVOL_STREAM(stream));
}
int svol_eval_average_precision(const float* pred, const int32_t* frame_index, const float* gt, const int32_t* gt_off,
                                const int32_t* frame_off, int32_t units, int32_t q_per_frame, int32_t max_frames, int32_t max_gt,
                                double* ap, void* stream) {
  SVOL_REQUIRE(pred); SVOL_REQUIRE(frame_index); SVOL_REQUIRE(gt); SVOL_REQUIRE(gt_off); SVOL_REQUIRE(frame_off); SVOL_REQUIRE(ap);
  return launch_eval_average_precision(pred, frame_index, gt, gt_off, frame_off, units, q_per_frame, max_frames, max_gt, ap,
                                       SVOL_STREAM(stream));
}
int svol_pack_weights(const svol_pack_job* jobs, int32_t n_jobs, void* stream) {
  SVOL_REQUIRE(jobs);
  return launch_pack_weights(jobs, n_jobs, SVOL_STREAM(stream));
}
int svol_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int32_t step, float grad_scale, void* stream) {
  SVOL_REQUIRE(p); SVOL_REQUIRE(g); SVOL_REQUIRE(m); SVOL_REQUIRE(v);
  return launch_adamw(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, SVOL_STREAM(stream));
}
int svol_adamw_segments(float* p, const float* g, float* m, float* v, int64_t n, const int64_t* seg_end,
                        const int32_t* seg_group, int32_t n_seg, const svol_adamw_group* groups, int32_t n_groups,
                        float grad_scale, void* stream) {
  SVOL_REQUIRE(p); SVOL_REQUIRE(g); SVOL_REQUIRE(m); SVOL_REQUIRE(v); SVOL_REQUIRE(seg_end); SVOL_REQUIRE(seg_group);
  SVOL_REQUIRE(groups);
  return launch_adamw_segments(p, g, m, v, n, reinterpret_cast<const long long*>(seg_end), seg_group, n_seg, groups, n_groups,
                               grad_scale, SVOL_STREAM(stream));
}

}  // extern "C"
