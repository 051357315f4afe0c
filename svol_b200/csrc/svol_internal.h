// Internal declarations shared by the translation units of libsvol_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/svol_b200.h"

namespace svol {

using GemmEpilogue = svol_gemm_epilogue;
using GemmArgs = svol_gemm_args;
using AttnArgs = svol_attn_args;
using MatchArgs = svol_match_args;
using CriterionArgs = svol_criterion_args;

// error plumbing (api.cu): records a thread-local message and returns `code`
int svol_fail(int code, const char* msg);
int svol_fail_cuda(cudaError_t e, const char* what);
int svol_check_launch(const char* what);   // cudaGetLastError() after a launch

int sm_count();

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows, row pitch `ld` elements,
// box (box_inner x box_outer), swizzle 128 / 64 / 32 bytes.  Out-of-bounds reads return zero.
int make_tensor_map_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld,
                       int box_inner, int box_outer, int swizzle_bytes);

int launch_gemm_bf16_tc(const GemmArgs& a, cudaStream_t stream);
int launch_gemm_bf16_plain(const GemmArgs& a, cudaStream_t stream);
int launch_attention_tc(const AttnArgs& a, cudaStream_t stream);
int launch_attention_small(const AttnArgs& a, cudaStream_t stream);   // short key sequences; -1: does not qualify
int launch_ffn_tc(const svol_ffn_args& a, cudaStream_t stream);
int launch_attention_plain(const AttnArgs& a, cudaStream_t stream);
int launch_match(const MatchArgs& a, cudaStream_t stream);
int launch_match_localize(int64_t* tgt_idx, const int32_t* video_match_off, int NL, int B, int K, cudaStream_t stream);
int launch_lsap_f32(const float* cost, const int64_t* cost_off, const int32_t* shape, int n_problems, int max_small,
                    int max_big, int max_entries, int64_t* rows_out, int64_t* cols_out, const int64_t* out_off, int32_t* status, int solver,
                    cudaStream_t stream);
int launch_criterion(const CriterionArgs& a, cudaStream_t stream);
int launch_criterion_backward(const CriterionArgs& a, const float* grad_w, float* grad_logits, float* grad_boxes,
                              cudaStream_t stream);

}  // namespace svol
