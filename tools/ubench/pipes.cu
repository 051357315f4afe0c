// Micro-benchmark: issue rate (cycles per warp instruction per SM sub-partition) of the instructions the
// softmax / GELU epilogues are made of.   nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, float a, float b) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = a * (i + 1) + threadIdx.x * 1e-3f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 128; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      if (MODE == 0) { v[i] = fmaf(v[i], a, b); v[i + 1] = fmaf(v[i + 1], a, b); }                  // FFMA 3-reg
      if (MODE == 1) { v[i] = fmaf(v[i], a, 0.1234f); v[i + 1] = fmaf(v[i + 1], a, 0.1234f); }      // FFMA imm
      if (MODE == 2) { float2 r = __ffma2_rn(make_float2(v[i], v[i + 1]), make_float2(a, a), make_float2(b, b)); v[i] = r.x; v[i + 1] = r.y; }
      if (MODE == 3) { float2 r = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(b, b)); v[i] = r.x; v[i + 1] = r.y; }
      if (MODE == 4) { v[i] = fmaxf(v[i], b); v[i + 1] = fminf(v[i + 1], a); }                       // FMNMX
      if (MODE == 5) { __nv_bfloat162 p = __floats2bfloat162_rn(v[i], v[i + 1]); unsigned u = *reinterpret_cast<unsigned*>(&p); v[i] = __uint_as_float(u); }  // F2FP (1 per pair)
      if (MODE == 6) { v[i] = v[i] + b; v[i + 1] = v[i + 1] + b; }                                   // FADD
      if (MODE == 7) { v[i] = v[i] * a; v[i + 1] = v[i + 1] * a; }                                   // FMUL
      if (MODE == 8) { v[i] = __uint_as_float((__float_as_uint(v[i]) & 0x7fffffffu) | (__float_as_uint(v[i + 1]) & 0x80000000u)); }  // LOP3
    }
  }
  __syncthreads();                 // all warps done (unfair arbitration would let warp 0 finish early)
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[] = {"FFMA reg (2/pair)", "FFMA imm (2/pair)", "FFMA2 (1/pair)", "FADD2 (1/pair)", "FMNMX (2/pair)",
                         "F2FP pack (1/pair)", "FADD (2/pair)", "FMUL (2/pair)", "LOP3 (1/pair)"};
  for (int mode = 0; mode < 9; ++mode)
    for (int warps : {4, 8, 16}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
          case 0: k<0><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 1: k<1><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 2: k<2><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 3: k<3><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 4: k<4><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 5: k<5><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 6: k<6><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 7: k<7><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
          case 8: k<8><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f); break;
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      double pairs = 128.0 * 16;   // pair-steps per warp
      printf("%-20s warps/SMSP %d: %.2f cycles per element-pair per SMSP\n", names[mode], warps / 4, double(h) / pairs / (warps / 4));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
