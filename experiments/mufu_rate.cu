// Microbenchmark: sustained MUFU.EX2 rate per SM, alone and with the attention softmax's companion
// instructions (FFMA2 affine, FADD2 row sum, F2FP bf16x2 pack).   nvcc -arch=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, int iters, float c, float m) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  float2 acc = make_float2(0.f, 0.f);
  uint32_t pk = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float a = x[i], b = x[i + 1];
      if (MODE >= 1) {
        float2 t = __ffma2_rn(make_float2(a, b), make_float2(c, c), make_float2(m, m));
        a = t.x; b = t.y;
      }
      float p0, p1;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(b));
      if (MODE >= 1) {
        acc = __fadd2_rn(acc, make_float2(p0, p1));
        __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
        pk ^= *reinterpret_cast<uint32_t*>(&h);
        x[i] = p0 * 0.5f; x[i + 1] = p1 * 0.5f;
      } else {
        x[i] = p0; x[i + 1] = p1;
      }
    }
  }
  float s = acc.x + acc.y + __uint_as_float(pk);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(int warps_per_sm, const char* name) {
  int sms = 148;
  float* out;
  cudaMalloc(&out, sms * warps_per_sm * 32 * sizeof(float));
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, 100, 0.9f, -0.1f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 0.9f, -0.1f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double mufu = (double)sms * warps_per_sm * 32 * iters * 16;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s warps/SM=%2d  %.3f ms  %.2f MUFU/ns/SM  = %.2f per clk per SM at %.0f MHz (max boost)\n", name, warps_per_sm, ms,
         mufu / (ms * 1e6) / sms, mufu / (ms * 1e6) / sms / (clk * 1e-6), clk * 1e-3);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 12, 16, 32}) run<0>(w, "MUFU.EX2 only");
  for (int w : {4, 8, 12, 16, 32}) run<1>(w, "MUFU + FFMA2/FADD2/F2FP");
  return 0;
}
