// fp64_peak.cu — measures the FP64 vector (DFMA) throughput of the GPU, the compute roofline of the particle step
// ("B200 FP64 vector peak is not in MEASURED_PEAKS — measure it").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_peak scripts/fp64_peak.cu && gpurun_out/fp64_peak
// Prints one JSON line: {"fp64_tflops": ..., "dfma_per_clk_per_sm": ..., "sm_count": ..., "sm_mhz": ...}
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);   // independent chains: ILP per thread
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the loop alive, never true in practice
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { std::printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 1024);
  constexpr int ILP = 8;
  const int iters = 1 << 16, blocks = prop.multiProcessorCount * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.;
  for (int rep = 0; rep < 5; ++rep) {   // first repetitions warm up clocks
    cudaEventRecord(e0);
    k_dfma<ILP><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)blocks * threads * ILP * (double)iters;
    const double tf = 2.0 * fma / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  const double per_clk_sm = best * 1e12 / 2.0 / ((double)clk_khz * 1e3) / prop.multiProcessorCount;
  std::printf("{\"fp64_tflops\": %.3f, \"dfma_per_clk_per_sm\": %.2f, \"sm_count\": %d, \"sm_mhz\": %.0f}\n", best, per_clk_sm,
              prop.multiProcessorCount, clk_khz / 1e3);
  cudaFree(out);
  return 0;
}
