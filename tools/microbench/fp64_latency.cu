// Dependent-issue latency of the fp64 pipe on this GPU (one warp, no contention) and with 4 warps per scheduler -- the
// number that bounds the contact kernel's Gauss-Seidel chain.   nvcc -arch=sm_100a -O3 -fmad=false -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cycles, int iters, double a, double b) {
  double x = a + threadIdx.x * 1e-9, y = b;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (OP == 0) x = x + y;            // DADD
      else if (OP == 1) x = x * y;       // DMUL
      else if (OP == 2) x = fma(x, y, y); // DFMA
      else if (OP == 3) x = x * y + y;   // DMUL + DADD (no contraction with -fmad=false)
      else if (OP == 4) x = sqrt(x + y); // DADD + sqrt
      else x = y / (x + y);              // DADD + division
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
  const char* names[6] = {"DADD", "DMUL", "DFMA", "DMUL+DADD", "DADD+sqrt", "DADD+div"};
  const int ops_per[6] = {1, 1, 1, 2, 1, 1};
  for (int threads : {32, 128, 512}) {
    for (int op = 0; op < 6; ++op) {
      const int iters = 2000;
      long long h[1];
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: chain<0><<<1, threads>>>(out, cyc, iters, 1.0, 1e-9); break;
          case 1: chain<1><<<1, threads>>>(out, cyc, iters, 1.0, 1.0000001); break;
          case 2: chain<2><<<1, threads>>>(out, cyc, iters, 1.0, 1e-9); break;
          case 3: chain<3><<<1, threads>>>(out, cyc, iters, 1.0, 1e-9); break;
          case 4: chain<4><<<1, threads>>>(out, cyc, iters, 1.0, 1e-9); break;
          default: chain<5><<<1, threads>>>(out, cyc, iters, 1.0, 1e-9); break;
        }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
      printf("%3d threads/SM  %-10s %.1f cycles per dependent step (%d fp64 op%s each)\n", threads, names[op], (double)h[0] / (iters * 16.0), ops_per[op],
             ops_per[op] > 1 ? "s" : "");
    }
  }
  return 0;
}
