// r02_pipes.cu -- round-2 pipe / register-file micro-benchmarks for B200 (sm_100a).
// Questions:
//   (1) what limits an all-distinct-operand FFMA2 stream at 66 % of the FP32 peak: register BANK conflicts (depends on
//       which registers an instruction names) or register-file BANDWIDTH (depends only on how many it reads)?
//   (2) does the FP64 pipe run beside the FP32 pipe on sm_100 (DFMA interleaved with FFMA2), and at what rate?
// One CTA of 1024 threads per SM (8 warps per scheduler, 16 independent chains per thread); the SM clock is
// measured in the kernel (clock64 against %globaltimer), so every rate is in lane-operations per SM per cycle.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o r02_pipes r02_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NACC = 16;
constexpr int ITERS = 2048;

enum Mode {
  FFMA2_ACC,     // a[i] = fma(a[i+J], a[i+K], a[i])        three varying register pairs
  FFMA2_2V,      // a[i] = fma(a[i+J], b, a[i])             two varying + one constant (reuse candidate)
  FMUL2_2V,      // a[i] = a[i+J] * a[i+K]                  two varying
  FADD2_2V,      // a[i] = a[i+J] + a[i+K]
  FFMA_ACC,      // scalar: a[i] = fma(a[i+J], a[i+K], a[i])
  FMUL_2V,       // scalar: a[i] = a[i+J] * a[i+K]
  DFMA_ACC,      // double: d[i] = fma(d[i+J], d[i+K], d[i])
  DFMA_2V,       // double: d[i] = fma(d[i+J], c, d[i])
  MIX_F2_D,      // per i: two FFMA2_2V-style + one DFMA_2V (independent streams)
  MIX_F2_D_ACC,  // per i: two FFMA2_ACC + one DFMA_ACC
  MIX_FMA_ADD2,  // alternate FFMA2_ACC and FADD2_2V (3-read and 2-read instructions)
  CVT_F2D,       // float -> double -> float round trips (F2F throughput)
  MIX_F2_CVT,    // FFMA2_2V with one F2F.F64.F32 per 4 packed instructions
};

template <int MODE, int J, int K>
__global__ void __launch_bounds__(1024) kern(float* out, float b, float c, unsigned long long* meas) {
  float2 a[NACC];
  float s[NACC];
  double d[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    s[i] = threadIdx.x * 1e-3f + i;
    a[i] = make_float2(s[i], s[i] + 0.5f);
    d[i] = (double)s[i] * 1.000001;
  }
  const float2 b2 = make_float2(b, b * 1.0001f);
  const double cd = (double)c;
  unsigned long long g0, g1;
  long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const int j = (i + J) % NACC, k = (i + K) % NACC;
      if (MODE == FFMA2_ACC) a[i] = __ffma2_rn(a[j], a[k], a[i]);
      if (MODE == FFMA2_2V) a[i] = __ffma2_rn(a[j], b2, a[i]);
      if (MODE == FMUL2_2V) a[i] = __fmul2_rn(a[j], a[k]);
      if (MODE == FADD2_2V) a[i] = __fadd2_rn(a[j], a[k]);
      if (MODE == FFMA_ACC) s[i] = fmaf(s[j], s[k], s[i]);
      if (MODE == FMUL_2V) s[i] = s[j] * s[k];
      if (MODE == DFMA_ACC) d[i] = fma(d[j], d[k], d[i]);
      if (MODE == DFMA_2V) d[i] = fma(d[j], cd, d[i]);
      if (MODE == MIX_F2_D) {
        a[i] = __ffma2_rn(a[j], b2, a[i]);
        if ((i & 1) == 0) d[i / 2] = fma(d[(i / 2 + 3) % (NACC / 2)], cd, d[i / 2]);
      }
      if (MODE == MIX_F2_D_ACC) {
        a[i] = __ffma2_rn(a[j], a[k], a[i]);
        if ((i & 1) == 0) d[i / 2] = fma(d[(i / 2 + 3) % (NACC / 2)], d[(i / 2 + 5) % (NACC / 2)], d[i / 2]);
      }
      if (MODE == MIX_FMA_ADD2) a[i] = (i & 1) ? __ffma2_rn(a[j], a[k], a[i]) : __fadd2_rn(a[j], a[k]);
      if (MODE == CVT_F2D) s[i] = (float)((double)s[i] * cd);      // F2F.F64.F32, DMUL, F2F.F32.F64
      if (MODE == MIX_F2_CVT) {
        a[i] = __ffma2_rn(a[j], b2, a[i]);
        if ((i & 3) == 0) d[i / 4] = (double)a[i].x;
      }
    }
  }
  t1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc += a[i].x + a[i].y + s[i] + (float)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) { meas[0] = (unsigned long long)(t1 - t0); meas[1] = g1 - g0; }
}

struct Result { double ms, cycles, mhz; };

template <int MODE, int J, int K>
void run(const char* name, double f32_lane_ops_per_i, double f64_lane_ops_per_i, int sms, int threads = 1024) {
  const int blocks = sms;          // ONE CTA per SM: `threads`/128 warps per scheduler, all co-resident
  float* out; unsigned long long* meas;
  CK(cudaMalloc(&out, (size_t)blocks * 1024 * 4)); CK(cudaMalloc(&meas, 16));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) kern<MODE, J, K><<<blocks, threads>>>(out, 1.0001f, 1e-4f, meas);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  kern<MODE, J, K><<<blocks, threads>>>(out, 1.0001f, 1e-4f, meas);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long m[2]; CK(cudaMemcpy(m, meas, 16, cudaMemcpyDeviceToHost));
  const double cyc = (double)m[0], mhz = cyc / (double)m[1] * 1e3;
  // rates from the event time of the whole launch (one wave) and the SM clock measured in the kernel
  const double n_i = (double)ITERS * NACC * threads;       // loop bodies per SM
  const double clk = ms * 1e-3 * mhz * 1e6;                          // SM cycles of the launch
  printf("{\"mode\": \"%s\", \"J\": %d, \"K\": %d, \"warps_per_scheduler\": %d, \"ms\": %.4f, \"sm_mhz\": %.0f, "
         "\"f32_lane_ops_per_sm_clk\": %.1f, \"f32_frac_of_128\": %.3f, \"f64_lane_ops_per_sm_clk\": %.1f}\n",
         name, J, K, threads / 128, ms, mhz, n_i * f32_lane_ops_per_i / clk, n_i * f32_lane_ops_per_i / clk / 128.0,
         n_i * f64_lane_ops_per_i / clk);
  CK(cudaFree(out)); CK(cudaFree(meas));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  // (1) bank structure: the same three-operand stream with different register distances
  run<FFMA2_ACC, 1, 2>("ffma2_acc", 2, 0, sms);
  run<FFMA2_ACC, 2, 4>("ffma2_acc", 2, 0, sms);
  run<FFMA2_ACC, 4, 8>("ffma2_acc", 2, 0, sms);
  run<FFMA2_ACC, 5, 11>("ffma2_acc", 2, 0, sms);
  run<FFMA2_ACC, 1, 1>("ffma2_acc_sq", 2, 0, sms);       // a[i] += a[j]*a[j]: two distinct registers
  run<FFMA_ACC, 1, 2>("ffma_acc", 1, 0, sms);
  run<FFMA_ACC, 2, 4>("ffma_acc", 1, 0, sms);
  run<FFMA_ACC, 4, 8>("ffma_acc", 1, 0, sms);
  run<FFMA_ACC, 5, 11>("ffma_acc", 1, 0, sms);
  run<FFMA_ACC, 1, 1>("ffma_acc_sq", 1, 0, sms);
  // two-operand streams
  run<FFMA2_2V, 1, 0>("ffma2_2v_const", 2, 0, sms);
  run<FFMA2_2V, 5, 0>("ffma2_2v_const", 2, 0, sms);
  run<FMUL2_2V, 1, 2>("fmul2_2v", 2, 0, sms);
  run<FMUL2_2V, 5, 11>("fmul2_2v", 2, 0, sms);
  run<FADD2_2V, 5, 11>("fadd2_2v", 2, 0, sms);
  run<FMUL_2V, 5, 11>("fmul_2v", 1, 0, sms);
  run<MIX_FMA_ADD2, 5, 11>("mix_ffma2_fadd2", 2, 0, sms);
  // (2) FP64 pipe
  run<DFMA_ACC, 5, 11>("dfma_acc", 0, 1, sms);
  run<DFMA_2V, 5, 0>("dfma_2v_const", 0, 1, sms);
  run<MIX_F2_D, 5, 0>("mix_ffma2_dfma_2v", 2, 0.5, sms);
  run<MIX_F2_D_ACC, 5, 11>("mix_ffma2_dfma_acc", 2, 0.5, sms);
  run<CVT_F2D, 0, 0>("cvt_f2d_dmul_d2f", 0, 1, sms);
  run<MIX_F2_CVT, 5, 0>("mix_ffma2_cvt", 2, 0.25, sms);
  // (3) does the reuse cache survive warp interleaving?  one reused operand, 2 / 4 / 8 / 16 warps per scheduler... and
  // the all-distinct stream for comparison
  for (int t = 128; t <= 1024; t *= 2) {
    run<FFMA2_2V, 5, 0>("ffma2_2v_const", 2, 0, sms, t);
    run<FFMA2_ACC, 5, 11>("ffma2_acc", 2, 0, sms, t);
    run<FMUL2_2V, 5, 11>("fmul2_2v", 2, 0, sms, t);
  }
  return 0;
}
