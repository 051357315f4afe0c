// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition for 1..4 warps per SMSP, alone and mixed with the
// softmax companions (FADD2 subtract, F2FP pack), and the PACKED forms ex2.approx.f16x2 / ex2.approx.ftz.bf16x2 (modes 3, 4:
// two results per instruction -- on sm_100a they compile to two MUFU.EX2.{F16,BF16} + PRMT, this measures what that costs).
//   nvcc -arch=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2_f16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = seed * (i + 1) + threadIdx.x * 1e-3f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = ex2(v[i]);
    } else if (MODE == 1) {        // sub (packed) + ex2
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float2 x = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(-seed, -seed));
        v[i] = ex2(x.x); v[i + 1] = ex2(x.y);
      }
    } else if (MODE == 3 || MODE == 4) {   // packed: 64 registers = 128 half-precision values
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const unsigned u = __float_as_uint(v[i]);
        v[i] = __uint_as_float(MODE == 3 ? ex2_f16x2(u) : ex2_bf16x2(u));
      }
    } else {                       // sub + ex2 + pack (result fed back through unpack to keep the chain)
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float2 x = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(-seed, -seed));
        __nv_bfloat162 p = __floats2bfloat162_rn(ex2(x.x), ex2(x.y));
        unsigned u = *reinterpret_cast<unsigned*>(&p);
        v[i] = __uint_as_float(u << 16); v[i + 1] = __uint_as_float(u & 0xffff0000u);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  for (int mode = 0; mode < 5; ++mode)
    for (int warps : {1, 2, 4, 8, 12, 16}) {   // per SM; warps/SMSP = warps/4 (1 -> a single SMSP)
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, 0.001f);
        if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, 0.001f);
        if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, 0.001f);
        if (mode == 3) k<3><<<1, warps * 32>>>(out, cyc, 0.001f);
        if (mode == 4) k<4><<<1, warps * 32>>>(out, cyc, 0.001f);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      double per = double(h) / (64.0 * 64.0);
      const double results = mode >= 3 ? 2.0 : 1.0;      // results per lane and instruction
      printf("mode %d warps/SM %2d: %lld cycles, %.2f cyc per ex2 warp-instr (per warp), SM rate %.1f results/clk\n", mode, warps,
             h, per, results * warps * 32.0 / per);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
