// Dependent-issue latency probe for the fp64 pipe on sm_100a (one warp, one CTA).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double a, double b) {
  const int N = 4096;
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  double y = a + threadIdx.x, z = b + threadIdx.x;
#pragma unroll 8
  for (int i = 0; i < N; ++i) { y = fma(y, b, a); z = fma(z, b, a); }
  long long t2 = clock64();
  double s = a + threadIdx.x;
#pragma unroll 16
  for (int i = 0; i < N; ++i) s += __shfl_xor_sync(0xffffffffu, s, 1);
  long long t3 = clock64();
  double r = a + threadIdx.x + 1.0;
#pragma unroll 4
  for (int i = 0; i < 512; ++i) r = rsqrt(r) + a;
  long long t4 = clock64();
  double c = a + threadIdx.x + 1.0;
#pragma unroll 4
  for (int i = 0; i < 512; ++i) c = __drcp_rn(c) + a;
  long long t5 = clock64();
  float f = (float)a + threadIdx.x;
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = fmaf(f, (float)b, (float)a);
  long long t6 = clock64();
  double m0 = a, m1 = b;
#pragma unroll 8
  for (int i = 0; i < 1024; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(m0), "+d"(m1) : "d"(a), "d"(b));
  long long t7 = clock64();
  __shared__ double sm[64];
  sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
  __syncwarp();
  long long t8 = clock64();
  double l = a; int idx = threadIdx.x;
#pragma unroll 8
  for (int i = 0; i < 1024; ++i) { l += sm[idx]; idx = (idx + (int)l) & 63; }
  long long t9 = clock64();
  double gsum = 0.0;
  const double* gp = out;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { gsum += __ldcg(gp + ((int)gsum & 31)); }
  long long t10 = clock64();
  out[threadIdx.x] = x + y + z + s + r + c + f + m0 + m1 + l + gsum;
  if (threadIdx.x == 0) { cyc[7] = t9 - t8; cyc[8] = t10 - t9; }
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = (t4 - t3); cyc[4] = (t5 - t4); cyc[5] = t6 - t5; cyc[6] = t7 - t6;
  }
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 16 * 8);
  for (int rep = 0; rep < 2; ++rep) { probe<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); }
  printf("DFMA dependent: %.1f cyc\n", cyc[0] / 4096.0);
  printf("2 interleaved DFMA chains: %.1f cyc per pair\n", cyc[1] / 4096.0);
  printf("SHFL64 + DADD dependent: %.1f cyc\n", cyc[2] / 4096.0);
  printf("rsqrt(double)+DADD: %.1f cyc\n", cyc[3] / 512.0);
  printf("__drcp_rn+DADD: %.1f cyc\n", cyc[4] / 512.0);
  printf("FFMA dependent: %.1f cyc\n", cyc[5] / 4096.0);
  printf("DMMA dependent (same accumulator): %.1f cyc\n", cyc[6] / 1024.0);
  printf("LDS + DADD + index dependent: %.1f cyc\n", cyc[7] / 1024.0);
  printf("ld.cg (L2 hit) + DADD dependent: %.1f cyc\n", cyc[8] / 256.0);
  return 0;
}
