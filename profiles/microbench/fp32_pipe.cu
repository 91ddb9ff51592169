// FP32-pipe micro-benchmark for B200 (sm_100a).
// Measures warp-instruction issue rates of the instruction mix the EKF step kernel is made of, so
// that the roofline denominator ("FP32 peak") and the scalar-vs-packed (FFMA2) design choice are
// measured, not assumed.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipe fp32_pipe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NACC = 16;     // independent accumulators per thread (covers the 4-cycle latency)
constexpr int ITERS = 4096;

// mode 0: FFMA, 3 distinct register sources (a = a*b + c)
// mode 1: FFMA, accumulate form with all-varying operands (a[i] = a[j]*a[k] + a[i])
// mode 2: FFMA with an immediate multiplier
// mode 3: FFMA2 (packed f32x2) 3 distinct register-pair sources
// mode 4: FMUL  (a = a*b)
// mode 5: FADD  (a = a+b)
// mode 6: FMUL2 / FADD2 alternating
// mode 7: MUFU.RSQ
// mode 8: 8 FFMA : 1 MUFU.RSQ mix
// mode 9: FFMA2 accumulate form with varying operands
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float b, float c, long long* clk) {
  float a[NACC];
  float2 a2[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { a[i] = threadIdx.x * 1e-3f + i; a2[i] = make_float2(a[i], a[i] + 0.5f); }
  float2 b2 = make_float2(b, b * 1.0001f), c2 = make_float2(c, c * 0.999f);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], b, c);
      if (MODE == 1) a[i] = fmaf(a[(i + 5) % NACC], a[(i + 11) % NACC], a[i]);
      if (MODE == 2) a[i] = fmaf(a[i], 0.99993f, c);
      if (MODE == 3) a2[i] = __ffma2_rn(a2[i], b2, c2);
      if (MODE == 4) a[i] = a[i] * b;
      if (MODE == 5) a[i] = a[i] + b;
      if (MODE == 6) a2[i] = (i & 1) ? __fmul2_rn(a2[i], b2) : __fadd2_rn(a2[i], c2);
      if (MODE == 7) a[i] = rsqrtf(a[i]);
      if (MODE == 8) { a[i] = fmaf(a[i], b, c); if ((i & 7) == 7) a[i] = rsqrtf(a[i]); }
      if (MODE == 9) a2[i] = __ffma2_rn(a2[(i + 5) % NACC], a2[(i + 11) % NACC], a2[i]);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i] + a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char* name, double flops_per_inst, int sms) {
  int blocks = sms * 8;       // 8 CTAs x 256 threads = 2048 threads/SM = 64 warps/SM (full)
  float* out; long long* clk;
  CK(cudaMalloc(&out, (size_t)blocks * 256 * 4)); CK(cudaMalloc(&clk, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, clk);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int r = 0; r < reps; ++r) k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, clk);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  long long cyc; CK(cudaMemcpy(&cyc, clk, 8, cudaMemcpyDeviceToHost));
  double insts_per_thread = (double)ITERS * NACC;
  if (MODE == 8) insts_per_thread *= 1.125;
  double warp_insts = insts_per_thread * blocks * 256 / 32.0;
  double mhz = cyc / (ms * 1e3);                 // cycles of CTA 0 / elapsed(us) (approx; CTA0 spans ~ whole kernel)
  double per_smsp_clk = warp_insts / (sms * 4.0) / (double)cyc;   // warp-inst / cycle / SMSP
  double tflops = insts_per_thread * blocks * 256 * flops_per_inst / (ms * 1e-3) / 1e12;
  printf("{\"mode\": \"%s\", \"ms\": %.4f, \"cta0_cycles\": %lld, \"approx_sm_mhz\": %.0f, \"warp_inst_per_clk_per_smsp\": %.3f, \"tflops\": %.2f}\n",
         name, ms, cyc, mhz, per_smsp_clk, tflops);
  CK(cudaFree(out)); CK(cudaFree(clk));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  int sms = p.multiProcessorCount;
  run<0>("ffma_3reg", 2, sms);
  run<1>("ffma_varying", 2, sms);
  run<2>("ffma_imm", 2, sms);
  run<3>("ffma2_3reg", 4, sms);
  run<9>("ffma2_varying", 4, sms);
  run<4>("fmul", 1, sms);
  run<5>("fadd", 1, sms);
  run<6>("fmul2_fadd2", 2, sms);
  run<7>("mufu_rsq", 1, sms);
  run<8>("ffma8_mufu1", 2.0 * 8 / 9 + 1.0 / 9, sms);
  return 0;
}
