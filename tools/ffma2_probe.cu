// Throughput probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b,
                                                    unsigned long long c) {
  unsigned long long r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, long long* cyc, int iters) {
  float a[8];
  unsigned long long p[8];
  const float m = 1.0000001f, c = 1e-9f;
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-3f + i;
    float2 v = make_float2(a[i], a[i] + 1.f);
    p[i] = *reinterpret_cast<unsigned long long*>(&v);
  }
  float2 mv = make_float2(m, m), cv = make_float2(c, c);
  const unsigned long long mp = *reinterpret_cast<unsigned long long*>(&mv);
  const unsigned long long cp = *reinterpret_cast<unsigned long long*>(&cv);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], m, c);
      else p[i] = ffma2(p[i], mp, cp);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) {
    float2 v = *reinterpret_cast<float2*>(&p[i]);
    s += a[i] + v.x + v.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
  cudaMalloc(&cyc, 148 * 4 * sizeof(long long));
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    for (int bps = 1; bps <= 4; bps *= 2) {  // blocks per SM (256 threads each)
      const int grid = 148 * bps;
      if (mode == 0) probe<0><<<grid, 256>>>(out, cyc, iters);
      else probe<1><<<grid, 256>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long h[148 * 4];
      cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < grid; ++i) avg += h[i];
      avg /= grid;
      const double warp_instr = (double)iters * 8 * 8 * bps;  // per SM
      printf("%s  %d warps/SM: %.3f cycles per warp-instruction per SM  (%.1f fp32 FMA lanes/clk/SM)\n",
             mode ? "FFMA2" : "FFMA ", 8 * bps, avg / warp_instr,
             warp_instr * 32 * (mode ? 2 : 1) / avg);
    }
  }
  return 0;
}
