// Peak issue rate of the FP64 MMA pipe on this GPU: register-only loops of mma.sync f64 (m8n8k4 and m16n8k8), no memory
// traffic.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void k(double *out, int iters) {
  double c[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[4] = {1.0 + threadIdx.x, 0.5, 0.25, 0.125}, b[2] = {1e-3 * threadIdx.x, 2e-3};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (SHAPE == 8) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      }
    }
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
void run(const char *name, int blocks, int threads) {
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const int iters = 20000;
  k<SHAPE, NACC><<<blocks, threads>>>(out, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<SHAPE, NACC><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fl = (SHAPE == 8 ? 512.0 : 2048.0) * NACC * (double)iters * blocks * (threads / 32);
  printf("%s blocks %d threads %d acc %d: %.2f TFLOP/s (%.2f ms)\n", name, blocks, threads, NACC, fl / (ms * 1e-3) / 1e12, ms);
  cudaFree(out);
}

int main() {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sm = p.multiProcessorCount;
  run<8, 8>("m8n8k4 ", sm * 2, 256);
  run<8, 8>("m8n8k4 ", sm * 4, 256);
  run<8, 12>("m8n8k4 ", sm * 3, 256);
  run<16, 6>("m16n8k8", sm * 2, 256);
  run<16, 6>("m16n8k8", sm * 4, 256);
  run<16, 6>("m16n8k8", sm * 3, 256);
  return 0;
}
