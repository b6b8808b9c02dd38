// Probe of B200's FP64 pipe: dependent-chain latency and throughput of DFMA and F2F.F64.F32.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_dfma(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, b, a);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
__global__ void lat_f2f_dfma(double* out, long long* cyc, float a, double b) {
  float f = a;
  double acc = 0.0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc = fma((double)f, b, acc); f += 1.0f; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = acc;
}
template <int CH>
__global__ void thr_dfma(double* out, long long* cyc, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = a + c;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < CH; ++c) x[c] = fma(x[c], b, a);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void thr_f2f(double* out, long long* cyc, float a) {
  float f[CH];
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { f[c] = a + c; x[c] = 0; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < CH; ++c) { x[c] = (double)f[c]; f[c] = __double2float_rn(x[c]) + 1.0f; }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  lat_dfma<<<1, 32>>>(out, cyc, 1.0, 1.0000001); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.1f cycles\n", h / 4096.0);
  lat_f2f_dfma<<<1, 32>>>(out, cyc, 1.0f, 1.0000001); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent (F2F -> DFMA acc) per step: %.1f cycles\n", h / 4096.0);
  for (int warps = 1; warps <= 32; warps *= 2) {
    thr_dfma<8><<<1, 32 * warps>>>(out, cyc, 1.0, 1.0000001); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ins = 256.0 * 4 * 8 * warps;
    printf("DFMA throughput, %2d warps x 8 chains on one SM: %.2f cycles per warp-instruction (%.1f lanes/clk)\n", warps, h / ins, 32.0 * ins / h);
  }
  for (int warps = 1; warps <= 32; warps *= 4) {
    thr_f2f<8><<<1, 32 * warps>>>(out, cyc, 1.0f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ins = 256.0 * 4 * 8 * warps * 2;
    printf("F2F (f32<->f64) throughput, %2d warps: %.2f cycles per warp-instruction\n", warps, h / ins);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
