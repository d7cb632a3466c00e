// microbenchmark: mma.sync m16n8k8 tf32 (and m16n8k16 bf16) issue rate per SM sub-core on sm_100a
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/experiments/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int NACC, bool BF16>
__global__ void k(float* out, int iters, long long* cyc) {
  float acc[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < NACC; ++i) for (int e = 0; e < 4; ++e) s += acc[i][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC, bool BF16>
void run(int warps_per_sm) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<NACC, BF16><<<148, warps_per_sm * 32>>>(out, iters, cyc);
  k<NACC, BF16><<<148, warps_per_sm * 32>>>(out, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_subcore = (double)h / ((double)iters * NACC * (warps_per_sm / 4.0 < 1 ? 1 : warps_per_sm / 4.0));
  printf("%s NACC %2d warps/SM %2d: %.1f cycles per MMA per sub-core (%.1f per warp)\n", BF16 ? "bf16 k16" : "tf32 k8 ", NACC,
         warps_per_sm, per_subcore, (double)h / ((double)iters * NACC));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, false>(4); run<4, false>(4); run<8, false>(4); run<8, false>(8); run<8, false>(16); run<16, false>(16);
  run<1, true>(4); run<8, true>(4); run<8, true>(16);
  return 0;
}
