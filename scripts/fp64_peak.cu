// fp64_peak.cu — measured FP64 throughput of the B200 without FMA contraction (the element tapes are bit-faithful to
// the reference: separate DADD / DMUL, never DFMA) and with it, for the K1 "FP64 pipe utilisation" line of bench.py.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false scripts/fp64_peak.cu -o scripts/fp64_peak && scripts/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: dadd + dmul alternating (no FMA), 1: dfma
__global__ void __launch_bounds__(256) peak_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        x[i] = __dadd_rn(x[i], b);
        x[i] = __dmul_rn(x[i], a);
      } else {
        x[i] = __fma_rn(x[i], a, b);
        x[i] = __fma_rn(x[i], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d;
  cudaMalloc(&d, 8);
  const int iters = 20000, grid = sms * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best[2] = {0, 0};
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0)
        peak_kernel<0><<<grid, block>>>(d, iters, 1.0000001, 1e-7);
      else
        peak_kernel<1><<<grid, block>>>(d, iters, 1.0000001, 1e-7);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double ops = double(grid) * block * double(iters) * 8 * 2;  // instructions (lane-ops); an FMA counts 2 flops
      const double tf = ops * (mode == 1 ? 2.0 : 1.0) / (ms * 1e-3) / 1e12;
      if (tf > best[mode]) best[mode] = tf;
    }
  std::printf("{\"fp64_nofma_tflops\": %.3f, \"fp64_fma_tflops\": %.3f, \"sms\": %d, \"how\": \"8 independent chains per thread, "
              "256 threads x 8 CTAs per SM, alternating DADD/DMUL (no FMA) resp. DFMA, best of 5, CUDA events\"}\n",
              best[0], best[1], sms);
  return 0;
}
