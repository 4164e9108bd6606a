// FP64 peak probe for B200 (sm_100a): DFMA stream, DMMA (mma.sync f64) shapes,
// plus latency probes for sqrt/div/shfl. Output: one JSON object on stdout.
// This is measurement tooling for the roofline denominator (SURVEY.md §8d:
// "FP64 peak measured by the builder"); it is not on the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C/D 2 regs per lane
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters) {
  double c0[ILP], c1[ILP];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C/D 4 regs per lane (sm_90+)
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters) {
  double c[ILP][4];
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double b0 = 1.0 + threadIdx.x * 1e-4, b1 = b0 * 0.5;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C/D 4 regs
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters) {
  double c[ILP][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 * (i + 1);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 1; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// dependent-chain latency probes (single warp), cycles per op via clock64
__global__ void k_lat(double* out, long long* cyc, int iters) {
  double x = 1.5 + threadIdx.x * 1e-6;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64(); cyc[0] = t1 - t0;
  // sqrt chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = sqrt(x + 2.0);
  t1 = clock64(); cyc[1] = t1 - t0;
  // div chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = 3.0 / (x + 1.0);
  t1 = clock64(); cyc[2] = t1 - t0;
  // rsqrt chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = rsqrt(x + 2.0);
  t1 = clock64(); cyc[3] = t1 - t0;
  // shfl (64-bit) chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1e-9;
  t1 = clock64(); cyc[4] = t1 - t0;
  // sincos chain
  double s, c;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { sincos(x, &s, &c); x = s + c; }
  t1 = clock64(); cyc[5] = t1 - t0;
  // dmma 884 dependent chain
  double c0 = x, c1 = x;
  t0 = clock64();
  for (int i = 0; i < iters; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(1e-3), "d"(1e-3));
  t1 = clock64(); cyc[6] = t1 - t0;
  out[threadIdx.x] = x + c0 + c1;
}

// shuffle throughput: many warps, independent 32-bit shuffles
__global__ void __launch_bounds__(256) k_shfl(double* out, int iters) {
  int v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (i & 3));
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 123456789) out[0] = s;
}

template <typename F>
static double time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 4096));
  long long* cyc; CK(cudaMalloc(&cyc, 64 * sizeof(long long)));
  const int iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);

  for (int bps = 2; bps <= 8; bps *= 2) {  // blocks per SM of 256 threads
    int grid = sms * bps;
    double ms = time_ms([&] { k_dfma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    double tf = 2.0 * 8 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dfma_ilp8_bps%d_tflops\": %.3f", bps, tf);
  }
  {
    int grid = sms * 4;
    double ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_ilp16_bps4_tflops\": %.3f", 2.0 * 16 * iters * 256.0 * grid / (ms * 1e-3) / 1e12);
  }
  for (int bps = 1; bps <= 8; bps *= 2) {
    int grid = sms * bps;
    double ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, iters); });
    double tf = 2.0 * 8 * 8 * 4 * 8.0 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;  // 8 warps/block
    printf(", \"dmma_m8n8k4_ilp8_bps%d_tflops\": %.3f", bps, tf);
  }
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps;
    double ms = time_ms([&] { k_dmma1688<4><<<grid, 256>>>(out, iters); });
    double tf = 2.0 * 16 * 8 * 8 * 4.0 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dmma_m16n8k8_ilp4_bps%d_tflops\": %.3f", bps, tf);
  }
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps;
    double ms = time_ms([&] { k_dmma16816<4><<<grid, 256>>>(out, iters); });
    double tf = 2.0 * 16 * 8 * 16 * 4.0 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    printf(", \"dmma_m16n8k16_ilp4_bps%d_tflops\": %.3f", bps, tf);
  }
  {
    int grid = sms * 4;
    double ms = time_ms([&] { k_shfl<<<grid, 256>>>(out, iters); });
    // warp-instr per clk per SM = total warp shfls / (time * clock * sms)
    double shfl = 8.0 * iters * 8.0 * grid;
    printf(", \"shfl32_warp_instr_per_s\": %.4e, \"shfl_ms\": %.4f", shfl / (ms * 1e-3), ms);
  }
  {
    const int li = 2000;
    k_lat<<<1, 32>>>(out, cyc, li); CK(cudaDeviceSynchronize());
    k_lat<<<1, 32>>>(out, cyc, li); CK(cudaDeviceSynchronize());
    long long h[8]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    const char* names[] = {"dfma", "sqrt", "div", "rsqrt", "shfl64", "sincos", "dmma884"};
    for (int i = 0; i < 7; ++i) printf(", \"lat_%s_cyc\": %.1f", names[i], (double)h[i] / li);
  }
  printf("}\n");
  return 0;
}
