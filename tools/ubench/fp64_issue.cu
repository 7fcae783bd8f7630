// Microbenchmark: how the FP64 pipe of one SM sub-partition shares its issue port on sm_100a (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_issue fp64_issue.cu && ./fp64_issue
// Each kernel runs W warps per SMSP on ONE SM (block = 4*W warps) for ITER iterations of an unrolled body and
// reports cycles per iteration per SMSP.  Bodies: D = 48 independent-ish DADD/DMUL (6 chains), I = 48 integer ops,
// L = 12 conflict-free LDS.64, S = 12 SHFL pairs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2000

template <int MODE>
__global__ void __launch_bounds__(1024) k(double* out, long long* cyc, double c) {
  extern __shared__ double sh[];
  const int t = threadIdx.x;
  for (int i = t; i < 4096; i += blockDim.x) sh[i] = i * 1e-3;
  __syncthreads();
  double a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5;
  unsigned x0 = t, x1 = t * 3, x2 = t * 5, x3 = t * 7, x4 = t * 11, x5 = t * 13;
  double l = 0;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh) + (t & 31) * 8 + (t >> 5) * 256;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE & 1) {
        a0 = __dsub_rn(a0, __dmul_rn(c, a1)); a1 = __dsub_rn(a1, __dmul_rn(c, a2)); a2 = __dsub_rn(a2, __dmul_rn(c, a3));
        a3 = __dsub_rn(a3, __dmul_rn(c, a4)); a4 = __dsub_rn(a4, __dmul_rn(c, a5)); a5 = __dsub_rn(a5, __dmul_rn(c, a0));
      }
      if (MODE & 2) {
        x0 = x0 * 5 + x1; x1 = (x1 ^ x2) + 7; x2 = x2 * 3 + x3; x3 = (x3 ^ x4) + 1; x4 = x4 * 9 + x5; x5 = (x5 ^ x0) + 3;
        x0 ^= x3; x1 += x4; x2 ^= x5; x3 += x0; x4 ^= x1; x5 += x2;
      }
      if (MODE & 4) {
        double v0, v1, v2;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v0) : "r"(base + u * 1024));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v1) : "r"(base + u * 1024 + 8192));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v2) : "r"(base + u * 1024 + 16384));
        l += v0 + v1 + v2;
      }
      if (MODE & 8) {
        double v0, v1, v2;  // 16-byte lane stride: the k-1 neighbour loads of the fused kernels
        const uint32_t b2 = base + (t & 31) * 8;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v0) : "r"(b2 + u * 1024));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v1) : "r"(b2 + u * 1024 + 8192));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v2) : "r"(b2 + u * 1024 + 16384));
        l += v0 + v1 + v2;
      }
      if (MODE & 16) {
        a0 += __shfl_up_sync(0xffffffffu, a3, 1);
        a1 += __shfl_up_sync(0xffffffffu, a4, 1);
        a2 += __shfl_up_sync(0xffffffffu, a5, 1);
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + t] = a0 + a1 + a2 + a3 + a4 + a5 + (double)(x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5) + l;
  if (t == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  double* out; long long* cyc;
  const int threads = 128 * warps_per_smsp;
  cudaMalloc(&out, threads * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE><<<1, threads, 65536>>>(out, cyc, 0.1);
  k<MODE><<<1, threads, 65536>>>(out, cyc, 0.1);
  long long h = 0; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-34s warps/SMSP=%d  cycles/iter=%8.1f  (per warp-iter: %d FP64, %d int, %d lds, %d shfl)%s\n", name, warps_per_smsp,
         (double)h / ITER, (MODE & 1) ? 96 : 0, (MODE & 2) ? 96 : 0, (MODE & 12) ? 24 : 0, (MODE & 16) ? 48 : 0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<1>("FP64 only", w);
    run<2>("INT only", w);
    run<3>("FP64 + INT", w);
    run<4>("LDS.64 contiguous only", w);
    run<8>("LDS.64 16-byte lane stride only", w);
    run<5>("FP64 + LDS.64 contiguous", w);
    run<9>("FP64 + LDS.64 strided", w);
    run<16>("SHFL.f64 only", w);
    run<17>("FP64 + SHFL.f64", w);
  }
  return 0;
}
