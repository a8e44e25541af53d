// FP64 MMA pipe with the operand pattern of the 3M complex GEMM inner loop (12 m8n8k4 on 12 accumulators, operands
// a[2] x {re, im, sum}, b[2] x {re, im, sum} refreshed from shared memory every step), with and without the loads.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: registers only, operands rotate; 1: + LDS.128 per step like the GEMM
__global__ void __launch_bounds__(256, 3) k(double *out, int iters) {
  __shared__ double2 sm[64 * 20 + 32 * 20];
  for (int i = threadIdx.x; i < 64 * 20 + 32 * 20; i += blockDim.x) sm[i] = make_double2(1e-3 * i, 2e-3 * i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 16;
  double p1[2][2][2] = {}, p2[2][2][2] = {}, p3[2][2][2] = {};
  double2 a[2] = {make_double2(1.0 + lane, 0.5), make_double2(0.25, 0.125)}, b[2] = {make_double2(1e-3 * lane, 2e-3), make_double2(3e-3, 4e-3)};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4) {
      if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = sm[(wm + i * 8 + g) * 20 + kk + t];
#pragma unroll
        for (int j = 0; j < 2; ++j) b[j] = sm[64 * 20 + (wn + j * 8 + g) * 20 + kk + t];
      } else {
        a[0].x += 1e-9; a[1].y += 1e-9; b[0].x += 1e-9; b[1].y += 1e-9;
      }
      double asum[2], bsum[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) asum[i] = a[i].x + a[i].y;
#pragma unroll
      for (int j = 0; j < 2; ++j) bsum[j] = b[j].x + b[j].y;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(p1[i][j][0], p1[i][j][1], a[i].x, b[j].x);
          dmma(p2[i][j][0], p2[i][j][1], a[i].y, b[j].y);
          dmma(p3[i][j][0], p3[i][j][1], asum[i], bsum[j]);
        }
    }
  }
  double s = 0;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int c = 0; c < 2; ++c) s += p1[i][j][c] + p2[i][j][c] + p3[i][j][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int blocks) {
  double *out; cudaMalloc(&out, sizeof(double) * blocks * 256);
  const int iters = 4000;
  k<MODE><<<blocks, 256>>>(out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fl = 512.0 * 48 * (double)iters * blocks * 8;
  printf("%s blocks %d: %.2f TFLOP/s real (%.2f ms)\n", name, blocks, fl / (ms * 1e-3) / 1e12, ms);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sm = p.multiProcessorCount;
  run<0>("regs only ", sm * 3);
  run<1>("with LDS  ", sm * 3);
  run<1>("with LDS  ", sm * 2);
  run<1>("with LDS  ", sm * 1);
  return 0;
}
