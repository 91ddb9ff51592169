// posekf_kernels.cu -- sm_100a kernels + C ABI (include/posekf.h) of the batched quaternion EKF.
//
// Design (DESIGN.md has the long form):
//   * one filter per thread; X (4), upper-triangular P (10), the Wahba reference frame (12) and the
//     Q/R scalars live in registers for the whole launch -- a launch covers MANY timesteps;
//   * IMU samples are a structure-of-arrays stream [T][9][N] (filter index fastest), so a warp's
//     load of one channel of one step is one 128-byte line;
//   * staging: a TMA (cp.async.bulk.tensor.3d) ring of [TC][9][128] tiles in shared memory driven by
//     mbarriers (default), or plain coalesced LDG with a one-step register prefetch (unaligned input);
//   * template variants: Wahba solver (rank-2 QR / Jacobi), low-pass stage, auxiliary outputs
//     (trajectory, flip mask, tuning loss), compensated two-float state;
//   * no tensor cores (per-filter matrices are 4x4), no inter-thread communication on the step;
//   * filters are independent: multi-GPU = shard N, no collective on this path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "../../include/posekf.h"
#include "ekf_math.cuh"

using namespace pkf;

namespace {

// tunables (overridable with -D for experiments; the defaults are what ships)
#ifndef PKF_TMA_STEPS
#define PKF_TMA_STEPS 4
#endif
#ifndef PKF_TMA_STAGES
#define PKF_TMA_STAGES 2
#endif
#ifndef PKF_MIN_CTAS
#define PKF_MIN_CTAS 6
#endif
constexpr int kThreads = 128;                  // filters per CTA (4 warps)
constexpr int kMinCtasPerSm = PKF_MIN_CTAS;    // 6 CTAs/SM (24 warps) <=> <=80 registers per thread
constexpr int kTmaSteps = PKF_TMA_STEPS;       // TC: timesteps per TMA tile
constexpr int kTmaStages = PKF_TMA_STAGES;     // ring depth
constexpr int kChannels = 9;
// packed (two filters per thread) variant
#ifndef PKF_TMA2_STEPS
#define PKF_TMA2_STEPS 4
#endif
#ifndef PKF_TMA2_STAGES
#define PKF_TMA2_STAGES 2
#endif
#ifndef PKF_MIN_CTAS2
#define PKF_MIN_CTAS2 6
#endif
#ifndef PKF_THREADS2
#define PKF_THREADS2 64
#endif
constexpr int kThreads2 = PKF_THREADS2;        // threads per CTA of the packed kernel
constexpr int kTile2 = 2 * kThreads2;          // filters per CTA (two per thread); TMA box width, <= 256
constexpr int kTma2Steps = PKF_TMA2_STEPS;
constexpr int kTma2Stages = PKF_TMA2_STAGES;
#ifndef PKF_AUTO_PACKED
#define PKF_AUTO_PACKED 1
#endif
constexpr bool kAutoPrefersPacked = PKF_AUTO_PACKED != 0;   // whether POSEKF_STAGE_AUTO picks the packed kernel

#define PKF_CUDA_TRY(expr)                           \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldg_stream(const float* p) {
  // streamed once: read-only path, do not keep in L1
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ f32x2 ld2(const float* p) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  return f32x2(v.x, v.y);
}
__device__ __forceinline__ void st2(float* p, const f32x2& v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }

struct ReplayParams {
  int64_t N, T, Ns;
  const float* streams;
  const float* dt;
  int dt_per_step;
  const float* acc_ref;
  const float* mag_ref;
  const float* q_scale;
  const float* r_scale;
  float alpha_acc, alpha_mag;
  float* state_x;
  float* state_x_lo;    // [4][N] low-order part of the two-float state (compensated variant), or null
  float* state_p;
  float* state_lpf;
  float* out_traj;
  uint8_t* out_flip;
  const float* truth;   // [T][Ns][4] reference track for the tuning objective, or null
  float* loss_acc;      // [N] in/out: sum over steps of 1 - (x . truth)^2
};

struct FilterRegs {
  Quat<float> x;
  Quat<float> xlo;      // used by the compensated variant only
  Sym4<float> P;
  FilterConst<float> fc;
  Vec3<float> la, lm;   // low-pass state
};

template <bool LPF, bool COMP>
__device__ __forceinline__ void load_filter(const ReplayParams& p, int64_t n, int64_t col, FilterRegs& f) {
  const int64_t N = p.N, Ns = p.Ns;
  Vec3<float> ra = {p.acc_ref[col], p.acc_ref[Ns + col], p.acc_ref[2 * Ns + col]};
  Vec3<float> rm = {p.mag_ref[col], p.mag_ref[Ns + col], p.mag_ref[2 * Ns + col]};
  f.fc = make_filter_const<float>(ra, rm, p.q_scale[n], p.r_scale[n]);
  f.x = {p.state_x[n], p.state_x[N + n], p.state_x[2 * N + n], p.state_x[3 * N + n]};
  f.xlo = {0.f, 0.f, 0.f, 0.f};
  if (COMP) f.xlo = {p.state_x_lo[n], p.state_x_lo[N + n], p.state_x_lo[2 * N + n], p.state_x_lo[3 * N + n]};
  const float* sp = p.state_p + n;   // P/r: the step works in units of r (see ekf_step), so does the state buffer
  f.P = {sp[0], sp[N], sp[2 * N], sp[3 * N], sp[4 * N], sp[5 * N], sp[6 * N], sp[7 * N], sp[8 * N], sp[9 * N]};
  if (LPF) {
    const float* sl = p.state_lpf + n;
    f.la = {sl[0], sl[N], sl[2 * N]};
    f.lm = {sl[3 * N], sl[4 * N], sl[5 * N]};
  }
}

template <bool LPF, bool COMP>
__device__ __forceinline__ void store_filter(const ReplayParams& p, int64_t n, const FilterRegs& f) {
  const int64_t N = p.N;
  p.state_x[n] = f.x.w; p.state_x[N + n] = f.x.x; p.state_x[2 * N + n] = f.x.y; p.state_x[3 * N + n] = f.x.z;
  if (COMP) {
    p.state_x_lo[n] = f.xlo.w; p.state_x_lo[N + n] = f.xlo.x; p.state_x_lo[2 * N + n] = f.xlo.y; p.state_x_lo[3 * N + n] = f.xlo.z;
  }
  float* sp = p.state_p + n;
  sp[0] = f.P.a00; sp[N] = f.P.a01; sp[2 * N] = f.P.a02; sp[3 * N] = f.P.a03; sp[4 * N] = f.P.a11;
  sp[5 * N] = f.P.a12; sp[6 * N] = f.P.a13; sp[7 * N] = f.P.a22; sp[8 * N] = f.P.a23; sp[9 * N] = f.P.a33;
  if (LPF) {
    float* sl = p.state_lpf + n;
    sl[0] = f.la.x; sl[N] = f.la.y; sl[2 * N] = f.la.z; sl[3 * N] = f.lm.x; sl[4 * N] = f.lm.y; sl[5 * N] = f.lm.z;
  }
}

// One filter step + optional outputs.  AUX = the launch has a trajectory and/or flip output.
struct AuxPtrs {
  float4* traj;          // this filter's slot in out_traj, advanced by N per step
  uint8_t* flips;
  const float4* truth;   // this filter's column in the reference track, advanced by Ns per step
  float loss;
};

template <bool AUX> __device__ __forceinline__ AuxPtrs make_aux(const ReplayParams& p, int64_t n, int64_t col, bool valid) {
  AuxPtrs a = {nullptr, nullptr, nullptr, 0.f};
  if (AUX && valid) {
    if (p.out_traj) a.traj = reinterpret_cast<float4*>(p.out_traj) + n;
    if (p.out_flip) a.flips = p.out_flip + n;
    if (p.truth) { a.truth = reinterpret_cast<const float4*>(p.truth) + col; a.loss = p.loss_acc[n]; }
  }
  return a;
}

template <int ALGO, bool LPF, bool AUX, bool COMP>
__device__ __forceinline__ void filter_step(const ReplayParams& p, FilterRegs& f, const float (&s)[kChannels], float h,
                                            AuxPtrs& aux) {
  Vec3<float> w = {s[0], s[1], s[2]}, a = {s[3], s[4], s[5]}, m = {s[6], s[7], s[8]};
  if (LPF) {   // SRV/KalmanFilter.cpp:285,298 -- filtered values feed Wahba, not renormalised
    if (p.alpha_acc >= 0.f) { lowpass<float>(f.la, a, p.alpha_acc, 1.f - p.alpha_acc); a = f.la; }
    if (p.alpha_mag >= 0.f) { lowpass<float>(f.lm, m, p.alpha_mag, 1.f - p.alpha_mag); m = f.lm; }
  }
  bool flip;
  ekf_step<float, ALGO, AUX, COMP>(f.x, f.xlo, f.P, f.fc, w, a, m, h, flip);
  if (AUX) {
    if (aux.traj) {   // [T][N][4]: one 16-byte store per filter-step, consecutive filters consecutive
      asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(aux.traj), "f"(f.x.w), "f"(f.x.x), "f"(f.x.y),
                   "f"(f.x.z)
                   : "memory");
      aux.traj += p.N;
    }
    if (aux.flips) { *aux.flips = flip ? 1 : 0; aux.flips += p.N; }
    if (aux.truth) {  // tuning objective: sin^2 of the angle between the estimate and the reference track
      const float4 qt = __ldg(aux.truth);
      aux.truth += p.Ns;
      // sin^2 of the angle as the squared 4-D wedge product |X ^ q_ref|^2 = |X|^2 |q_ref|^2 - (X.q_ref)^2
      // (Lagrange identity): the six 2x2 minors are small numbers computed without the 1 - d^2
      // cancellation and without sensitivity to the 1e-7 norm error of either quaternion.
      const float xw = f.x.w, xx = f.x.x, xy = f.x.y, xz = f.x.z, qw = qt.x, qx = qt.y, qy = qt.z, qz = qt.w;
      const float m01 = fmaf(xw, qx, -(xx * qw)), m02 = fmaf(xw, qy, -(xy * qw)), m03 = fmaf(xw, qz, -(xz * qw));
      const float m12 = fmaf(xx, qy, -(xy * qx)), m13 = fmaf(xx, qz, -(xz * qx)), m23 = fmaf(xy, qz, -(xz * qy));
      aux.loss += fmaf(m23, m23, fmaf(m13, m13, fmaf(m12, m12, fmaf(m03, m03, fmaf(m02, m02, m01 * m01)))));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Replay, LDG staging: coalesced loads straight to registers, next step prefetched while the
// current one is computed.
// ---------------------------------------------------------------------------------------------
template <int ALGO, bool LPF, bool AUX, bool COMP>
__global__ void __launch_bounds__(kThreads, ALGO == WAHBA_JACOBI ? 4 : ((COMP || LPF) ? (kMinCtasPerSm > 5 ? 5 : kMinCtasPerSm) : kMinCtasPerSm))
    replay_ldg_kernel(const ReplayParams p) {
  const int64_t n = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (n >= p.N) return;
  const int64_t Ns = p.Ns;
  const int64_t col = (Ns == p.N) ? n : (n % Ns);
  FilterRegs f;
  load_filter<LPF, COMP>(p, n, col, f);
  AuxPtrs aux = make_aux<AUX>(p, n, col, true);
  const float* s = p.streams + col;
  const int64_t step_stride = kChannels * Ns;
  float cur[kChannels], nxt[kChannels];
#pragma unroll
  for (int c = 0; c < kChannels; ++c) cur[c] = ldg_stream(s + c * Ns);
  const float dt0 = p.dt[0];
  const int T = (int)p.T;
  for (int t = 0; t < T; ++t) {
    if (t + 1 < T) s += step_stride;
#pragma unroll
    for (int c = 0; c < kChannels; ++c) nxt[c] = ldg_stream(s + c * Ns);
    const float h = p.dt_per_step ? __ldg(p.dt + t) : dt0;
    filter_step<ALGO, LPF, AUX, COMP>(p, f, cur, h, aux);
#pragma unroll
    for (int c = 0; c < kChannels; ++c) cur[c] = nxt[c];
  }
  store_filter<LPF, COMP>(p, n, f);
  if (AUX && aux.truth) p.loss_acc[n] = aux.loss;
}

// ---------------------------------------------------------------------------------------------
// Replay, TMA staging: a ring of kTmaStages tiles [kTmaSteps][9][128] in shared memory, each filled
// by ONE cp.async.bulk.tensor.3d issued by thread 0 and signalled through an mbarrier; consumers
// release a tile with one mbarrier arrive per warp.  Out-of-range columns/steps are zero-filled by
// the TMA unit, so ragged N and T need no special casing on the load side.
// ---------------------------------------------------------------------------------------------
struct __align__(128) TmaSmem {
  float tile[kTmaStages][kTmaSteps][kChannels][kThreads];
  uint64_t full[kTmaStages];
  uint64_t empty[kTmaStages];
};
constexpr uint32_t kTileBytes = kTmaSteps * kChannels * kThreads * sizeof(float);

template <int ALGO, bool LPF, bool AUX, bool COMP>
__global__ void __launch_bounds__(kThreads, (COMP && (LPF || AUX)) ? (kMinCtasPerSm > 5 ? 5 : kMinCtasPerSm) : kMinCtasPerSm)
    replay_tma_kernel(const ReplayParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TmaSmem& sm = *reinterpret_cast<TmaSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int64_t n0 = (int64_t)blockIdx.x * kThreads;
  const int64_t n = n0 + tid;
  const bool valid = n < p.N;
  const int col0 = (int)((p.Ns == p.N) ? n0 : (n0 % p.Ns));
  const int T = (int)p.T;
  const int n_chunks = (T + kTmaSteps - 1) / kTmaSteps;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) {
      if (s < n_chunks) {
        mbar_expect_tx(&sm.full[s], kTileBytes);
        tma_load_3d(&sm.tile[s][0][0][0], &tmap, &sm.full[s], col0, 0, s * kTmaSteps);
      }
    }
  }

  FilterRegs f;
  if (valid) load_filter<LPF, COMP>(p, n, (int64_t)col0 + tid, f);
  AuxPtrs aux = make_aux<AUX>(p, n, (int64_t)col0 + tid, valid);
  const float dt0 = p.dt[0];

  int stage = 0;
  uint32_t parity = 0;
  for (int k = 0; k < n_chunks; ++k) {
    // producer: refill the tile that every warp released in the previous iteration
    if (tid == 0 && k >= 1 && (k - 1 + kTmaStages) < n_chunks) {
      const int ps = (stage == 0) ? kTmaStages - 1 : stage - 1;
      const uint32_t pp = (stage == 0) ? (parity ^ 1u) : parity;     // parity of iteration k-1
      mbar_wait(&sm.empty[ps], pp);
      mbar_expect_tx(&sm.full[ps], kTileBytes);
      tma_load_3d(&sm.tile[ps][0][0][0], &tmap, &sm.full[ps], col0, 0, (k - 1 + kTmaStages) * kTmaSteps);
    }
    mbar_wait(&sm.full[stage], parity);
    if (valid) {
      const int steps = min(kTmaSteps, T - k * kTmaSteps);     // < kTmaSteps only in the last chunk
#pragma unroll
      for (int tt = 0; tt < kTmaSteps; ++tt) {
        if (tt < steps) {
          float s[kChannels];
#pragma unroll
          for (int c = 0; c < kChannels; ++c) s[c] = sm.tile[stage][tt][c][tid];
          const float h = p.dt_per_step ? __ldg(p.dt + k * kTmaSteps + tt) : dt0;
          filter_step<ALGO, LPF, AUX, COMP>(p, f, s, h, aux);
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sm.empty[stage]);
    if (++stage == kTmaStages) { stage = 0; parity ^= 1; }
  }
  if (valid) store_filter<LPF, COMP>(p, n, f);
  if (AUX && valid && aux.truth) p.loss_acc[n] = aux.loss;
}

// ---------------------------------------------------------------------------------------------
// Replay, TMA staging, PACKED: each thread advances TWO filters in the lanes of f32x2 values, so
// every FP32 operation of the step is one FFMA2 / FMUL2 / FADD2.  Same tiles, same barriers and
// the same arithmetic per filter as replay_tma_kernel (results are bit-identical); half the
// threads.  Rank-2 Wahba only (the Jacobi variant uses the scalar kernel).
// ---------------------------------------------------------------------------------------------
struct __align__(128) Tma2Smem {
  float tile[kTma2Stages][kTma2Steps][kChannels][kTile2];
  uint64_t full[kTma2Stages];
  uint64_t empty[kTma2Stages];
};
constexpr uint32_t kTile2Bytes = kTma2Steps * kChannels * kTile2 * sizeof(float);

struct FilterRegs2 {
  Quat<f32x2> x, xlo;
  Sym4<f32x2> P;
  FilterConst<f32x2> fc;
  Vec3<f32x2> la, lm;
};

template <bool LPF, bool AUX, bool COMP>
__global__ void __launch_bounds__(kThreads2, (LPF && COMP) ? (PKF_MIN_CTAS2 > 6 ? 6 : PKF_MIN_CTAS2) : PKF_MIN_CTAS2)
    replay_tma2_kernel(const ReplayParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Tma2Smem& sm = *reinterpret_cast<Tma2Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int64_t n0 = (int64_t)blockIdx.x * kTile2;
  const int64_t n = n0 + 2 * tid;                  // this thread owns filters n and n + 1 (N is even)
  const bool valid = n < p.N;
  const int col0 = (int)((p.Ns == p.N) ? n0 : (n0 % p.Ns));
  const int T = (int)p.T;
  const int n_chunks = (T + kTma2Steps - 1) / kTma2Steps;
  const int64_t N = p.N, Ns = p.Ns;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTma2Stages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kThreads2 / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
#pragma unroll
    for (int s = 0; s < kTma2Stages; ++s) {
      if (s < n_chunks) {
        mbar_expect_tx(&sm.full[s], kTile2Bytes);
        tma_load_3d(&sm.tile[s][0][0][0], &tmap, &sm.full[s], col0, 0, s * kTma2Steps);
      }
    }
  }

  FilterRegs2 f;
  if (valid) {
    const int64_t col = (int64_t)col0 + 2 * tid;
    Vec3<f32x2> ra = {ld2(p.acc_ref + col), ld2(p.acc_ref + Ns + col), ld2(p.acc_ref + 2 * Ns + col)};
    Vec3<f32x2> rm = {ld2(p.mag_ref + col), ld2(p.mag_ref + Ns + col), ld2(p.mag_ref + 2 * Ns + col)};
    f.fc = make_filter_const<f32x2>(ra, rm, ld2(p.q_scale + n), ld2(p.r_scale + n));
    f.x = {ld2(p.state_x + n), ld2(p.state_x + N + n), ld2(p.state_x + 2 * N + n), ld2(p.state_x + 3 * N + n)};
    f.xlo = {f32x2(0.f), f32x2(0.f), f32x2(0.f), f32x2(0.f)};
    if (COMP) f.xlo = {ld2(p.state_x_lo + n), ld2(p.state_x_lo + N + n), ld2(p.state_x_lo + 2 * N + n), ld2(p.state_x_lo + 3 * N + n)};
    const float* sp = p.state_p + n;
    f.P = {ld2(sp), ld2(sp + N), ld2(sp + 2 * N), ld2(sp + 3 * N), ld2(sp + 4 * N), ld2(sp + 5 * N), ld2(sp + 6 * N),
           ld2(sp + 7 * N), ld2(sp + 8 * N), ld2(sp + 9 * N)};
    if (LPF) {
      const float* sl = p.state_lpf + n;
      f.la = {ld2(sl), ld2(sl + N), ld2(sl + 2 * N)};
      f.lm = {ld2(sl + 3 * N), ld2(sl + 4 * N), ld2(sl + 5 * N)};
    }
  }
  const float dt0 = p.dt[0];
  // auxiliary outputs: this thread's pair of adjacent slots
  float4* traj = nullptr;
  uint8_t* flips = nullptr;
  const float4* truth = nullptr;
  f32x2 loss(0.f);
  if (AUX && valid) {
    if (p.out_traj) traj = reinterpret_cast<float4*>(p.out_traj) + n;
    if (p.out_flip) flips = p.out_flip + n;
    if (p.truth) { truth = reinterpret_cast<const float4*>(p.truth) + ((int64_t)col0 + 2 * tid); loss = ld2(p.loss_acc + n); }
  }

  int stage = 0;
  uint32_t parity = 0;
  for (int k = 0; k < n_chunks; ++k) {
    if (tid == 0 && k >= 1 && (k - 1 + kTma2Stages) < n_chunks) {
      const int ps = (stage == 0) ? kTma2Stages - 1 : stage - 1;
      const uint32_t pp = (stage == 0) ? (parity ^ 1u) : parity;
      mbar_wait(&sm.empty[ps], pp);
      mbar_expect_tx(&sm.full[ps], kTile2Bytes);
      tma_load_3d(&sm.tile[ps][0][0][0], &tmap, &sm.full[ps], col0, 0, (k - 1 + kTma2Stages) * kTma2Steps);
    }
    mbar_wait(&sm.full[stage], parity);
    if (valid) {
      const int steps = min(kTma2Steps, T - k * kTma2Steps);
#pragma unroll
      for (int tt = 0; tt < kTma2Steps; ++tt) {
        if (tt < steps) {
          f32x2 s[kChannels];
#pragma unroll
          for (int c = 0; c < kChannels; ++c) s[c] = ld2(&sm.tile[stage][tt][c][2 * tid]);
          const float h = p.dt_per_step ? __ldg(p.dt + k * kTma2Steps + tt) : dt0;
          Vec3<f32x2> w = {s[0], s[1], s[2]}, a = {s[3], s[4], s[5]}, m = {s[6], s[7], s[8]};
          if (LPF) {
            if (p.alpha_acc >= 0.f) { lowpass<f32x2>(f.la, a, f32x2(p.alpha_acc), f32x2(1.f - p.alpha_acc)); a = f.la; }
            if (p.alpha_mag >= 0.f) { lowpass<f32x2>(f.lm, m, f32x2(p.alpha_mag), f32x2(1.f - p.alpha_mag)); m = f.lm; }
          }
          mask2 flip;
          ekf_step<f32x2, WAHBA_QR2, AUX, COMP>(f.x, f.xlo, f.P, f.fc, w, a, m, f32x2(h), flip);
          if (AUX) {
            if (traj) {   // [T][N][4]: two adjacent 16-byte quaternions
              asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(traj), "f"(f.x.w.x), "f"(f.x.x.x), "f"(f.x.y.x),
                           "f"(f.x.z.x) : "memory");
              asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(traj + 1), "f"(f.x.w.y), "f"(f.x.x.y),
                           "f"(f.x.y.y), "f"(f.x.z.y) : "memory");
              traj += N;
            }
            if (flips) { *reinterpret_cast<uchar2*>(flips) = make_uchar2(flip.x ? 1 : 0, flip.y ? 1 : 0); flips += N; }
            if (truth) {  // squared wedge product |X ^ q_ref|^2 per lane (see the scalar kernel)
              const float4 t0 = __ldg(truth), t1 = __ldg(truth + 1);
              truth += Ns;
              const f32x2 qw(t0.x, t1.x), qx(t0.y, t1.y), qy(t0.z, t1.z), qz(t0.w, t1.w);
              const f32x2 xw = f.x.w, xx = f.x.x, xy = f.x.y, xz = f.x.z;
              const f32x2 m01 = fma_(xw, qx, -(xx * qw)), m02 = fma_(xw, qy, -(xy * qw)), m03 = fma_(xw, qz, -(xz * qw));
              const f32x2 m12 = fma_(xx, qy, -(xy * qx)), m13 = fma_(xx, qz, -(xz * qx)), m23 = fma_(xy, qz, -(xz * qy));
              loss = loss + fma_(m23, m23, fma_(m13, m13, fma_(m12, m12, fma_(m03, m03, fma_(m02, m02, m01 * m01)))));
            }
          }
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sm.empty[stage]);
    if (++stage == kTma2Stages) { stage = 0; parity ^= 1; }
  }
  if (valid) {
    st2(p.state_x + n, f.x.w); st2(p.state_x + N + n, f.x.x); st2(p.state_x + 2 * N + n, f.x.y); st2(p.state_x + 3 * N + n, f.x.z);
    if (COMP) {
      st2(p.state_x_lo + n, f.xlo.w); st2(p.state_x_lo + N + n, f.xlo.x); st2(p.state_x_lo + 2 * N + n, f.xlo.y);
      st2(p.state_x_lo + 3 * N + n, f.xlo.z);
    }
    float* sp = p.state_p + n;
    st2(sp, f.P.a00); st2(sp + N, f.P.a01); st2(sp + 2 * N, f.P.a02); st2(sp + 3 * N, f.P.a03); st2(sp + 4 * N, f.P.a11);
    st2(sp + 5 * N, f.P.a12); st2(sp + 6 * N, f.P.a13); st2(sp + 7 * N, f.P.a22); st2(sp + 8 * N, f.P.a23); st2(sp + 9 * N, f.P.a33);
    if (LPF) {
      float* sl = p.state_lpf + n;
      st2(sl, f.la.x); st2(sl + N, f.la.y); st2(sl + 2 * N, f.la.z); st2(sl + 3 * N, f.lm.x); st2(sl + 4 * N, f.lm.y); st2(sl + 5 * N, f.lm.z);
    }
    if (AUX && p.truth) st2(p.loss_acc + n, loss);
  }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone Wahba (config "Wahba-only batched 3x3 SVD + R->quat").
// The Jacobi variant runs sweeps until every lane of the warp has converged (warp vote), at most
// `max_sweeps`.
// ---------------------------------------------------------------------------------------------
struct WahbaParams {
  int64_t N;
  const float *acc_ref, *mag_ref;
  int ref_shared;
  const float *acc, *mag, *k_acc, *k_mag;
  float k_acc_s, k_mag_s;
  int weights_from_acc;
  float *out_rot, *out_quat;
  int max_sweeps;
};

template <int ALGO> __global__ void __launch_bounds__(256) wahba_kernel(const WahbaParams p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n < p.N;
  const int64_t i = valid ? n : 0;   // tail lanes recompute element 0 so the warp vote stays full
  const int64_t N = p.N;
  Vec3<float> ra, rm;
  if (p.ref_shared) {
    ra = {__ldg(p.acc_ref), __ldg(p.acc_ref + 1), __ldg(p.acc_ref + 2)};
    rm = {__ldg(p.mag_ref), __ldg(p.mag_ref + 1), __ldg(p.mag_ref + 2)};
  } else {
    ra = {ldg_stream(p.acc_ref + i), ldg_stream(p.acc_ref + N + i), ldg_stream(p.acc_ref + 2 * N + i)};
    rm = {ldg_stream(p.mag_ref + i), ldg_stream(p.mag_ref + N + i), ldg_stream(p.mag_ref + 2 * N + i)};
  }
  Vec3<float> a = {ldg_stream(p.acc + i), ldg_stream(p.acc + N + i), ldg_stream(p.acc + 2 * N + i)};
  Vec3<float> m = {ldg_stream(p.mag + i), ldg_stream(p.mag + N + i), ldg_stream(p.mag + 2 * N + i)};
  float ka, km;
  if (p.k_acc) { ka = ldg_stream(p.k_acc + i); km = ldg_stream(p.k_mag + i); }
  else if (p.weights_from_acc) { ka = fabsf(a.z); km = 1.f - ka; }           // PKF/ExtendedKalmanFilter.py:71
  else { ka = p.k_acc_s; km = p.k_mag_s; }
  Mat3<float> R;
  if (ALGO == WAHBA_QR2) {
    R = wahba_qr2<float>(frame_from_pair<float>(ra, rm), a, m, ka, km);
  } else {
    Vec3<float> g0, g1, g2;
    wahba_form_b<float>(ra, rm, a, m, ka, km, g0, g1, g2);
    Vec3<float> v0 = {1.f, 0.f, 0.f}, v1 = {0.f, 1.f, 0.f}, v2 = {0.f, 0.f, 1.f};
    for (int s = 0; s < p.max_sweeps; ++s) {
      jacobi_pair(g0, g1, v0, v1);
      jacobi_pair(g0, g2, v0, v2);
      jacobi_pair(g1, g2, v1, v2);
      // converged when every pairwise column dot product is below eps * (largest column norm)^2
      float nmax = fmaxf(dot3(g0, g0), fmaxf(dot3(g1, g1), dot3(g2, g2)));
      bool more = jacobi_offdiag(g0, g1, g2) > 6e-8f * nmax;
      if (!__any_sync(0xffffffffu, more)) break;
    }
    R = rotation_from_svd_pairs<float>(g0, g1, g2, v0, v1, v2);
  }
  if (!valid) return;
  if (p.out_rot) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) stg_stream(p.out_rot + (3 * r + c) * N + n, R.m[r][c]);
  }
  if (p.out_quat) {
    Quat<float> q = rotation_to_quat_ref<float>(R);
    stg_stream(p.out_quat + n, q.w); stg_stream(p.out_quat + N + n, q.x);
    stg_stream(p.out_quat + 2 * N + n, q.y); stg_stream(p.out_quat + 3 * N + n, q.z);
  }
}

// ---------------------------------------------------------------------------------------------
// Comparison tracks of the tuning workflow (what Results/*.png overlays): the gyro-only attitude
// (RK4 without correction: SRV/KalmanFilter.cpp:149, `Quarternion_Gyro_pure`) and the Wahba-only
// attitude per sample (PKF/main_file.py:40: getQuarternion(acc, mag, .5, .5), raw sign convention).
// ---------------------------------------------------------------------------------------------
struct TracksParams {
  int64_t N, T, Ns;
  const float *streams, *dt;
  int dt_per_step;
  const float *acc_ref, *mag_ref;
  float k_acc, k_mag;
  int weights_from_acc;
  float* gyro_state;   // [4][N] in/out, or null
  float* out_gyro;     // [T][N][4] or null
  float* out_wahba;    // [T][N][4] or null
};

template <int ALGO> __global__ void __launch_bounds__(128) tracks_kernel(const TracksParams p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int64_t N = p.N, Ns = p.Ns, col = (Ns == N) ? n : (n % Ns);
  Vec3<float> ra = {p.acc_ref[col], p.acc_ref[Ns + col], p.acc_ref[2 * Ns + col]};
  Vec3<float> rm = {p.mag_ref[col], p.mag_ref[Ns + col], p.mag_ref[2 * Ns + col]};
  const RefFrame<float> E = frame_from_pair<float>(ra, rm);
  Quat<float> g = {1.f, 0.f, 0.f, 0.f};
  if (p.gyro_state) g = {p.gyro_state[n], p.gyro_state[N + n], p.gyro_state[2 * N + n], p.gyro_state[3 * N + n]};
  const float* s = p.streams + col;
  float4* og = p.out_gyro ? reinterpret_cast<float4*>(p.out_gyro) + n : nullptr;
  float4* ow = p.out_wahba ? reinterpret_cast<float4*>(p.out_wahba) + n : nullptr;
  const float dt0 = p.dt[0];
  for (int64_t t = 0; t < p.T; ++t, s += kChannels * Ns) {
    const float h = p.dt_per_step ? __ldg(p.dt + t) : dt0;
    if (og || p.gyro_state) {
      Vec3<float> hw = {0.5f * ldg_stream(s), 0.5f * ldg_stream(s + Ns), 0.5f * ldg_stream(s + 2 * Ns)};
      g = rk4_step<float>(g, hw, h);
      if (og) { *og = make_float4(g.w, g.x, g.y, g.z); og += N; }
    }
    if (ow) {
      Vec3<float> a = {ldg_stream(s + 3 * Ns), ldg_stream(s + 4 * Ns), ldg_stream(s + 5 * Ns)};
      Vec3<float> m = {ldg_stream(s + 6 * Ns), ldg_stream(s + 7 * Ns), ldg_stream(s + 8 * Ns)};
      float ka = p.k_acc, km = p.k_mag;
      if (p.weights_from_acc) { ka = fabsf(a.z); km = 1.f - ka; }
      Mat3<float> R = (ALGO == WAHBA_QR2) ? wahba_qr2<float>(E, a, m, ka, km) : wahba_jacobi<float>(ra, rm, a, m, ka, km, 6);
      Quat<float> q = rotation_to_quat_ref<float>(R);
      *ow = make_float4(q.w, q.x, q.y, q.z);
      ow += N;
    }
  }
  if (p.gyro_state) { p.gyro_state[n] = g.w; p.gyro_state[N + n] = g.x; p.gyro_state[2 * N + n] = g.y; p.gyro_state[3 * N + n] = g.z; }
}

// ---------------------------------------------------------------------------------------------
// Raw-sensor pre-processing of the online pipeline (the step in front of the filter): linear
// interpolation of the accel / mag samples that bracket the gyro timestamp, normalisation, and the
// optional alpha low-pass -- SRV/Parser.cpp:229-267 (ExecuteKalmanFilter, LinearInterpolationSensor),
// :221-228 (NormalizeValues), SRV/KalmanFilter.cpp:279-303 (low-pass inside Set*Measurements).
// Writes the [T][9][N] stream the replay kernel consumes.
// ---------------------------------------------------------------------------------------------
struct PreprocessParams {
  int64_t N, T;
  const float* gyro;        // [T][3][N]
  const float* raw_prev;    // [T][6][N]  acc xyz, mag xyz : sample before the gyro timestamp (y1)
  const float* raw_next;    // [T][6][N]  sample after (y2)
  const float* tspan;       // [T][4][N]  seconds: acc (t2-t1), acc (t3-t1), mag (t2-t1), mag (t3-t1)
  float alpha_acc, alpha_mag;
  float* lpf_state;         // [6][N] in/out or null
  float* out_streams;       // [T][9][N]
};

__global__ void __launch_bounds__(256) preprocess_kernel(const PreprocessParams p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int64_t N = p.N;
  const bool lpa = p.alpha_acc >= 0.f, lpm = p.alpha_mag >= 0.f;
  Vec3<float> la = {0.f, 0.f, 0.f}, lm = {0.f, 0.f, 0.f};
  if (p.lpf_state) {
    la = {p.lpf_state[n], p.lpf_state[N + n], p.lpf_state[2 * N + n]};
    lm = {p.lpf_state[3 * N + n], p.lpf_state[4 * N + n], p.lpf_state[5 * N + n]};
  }
  for (int64_t t = 0; t < p.T; ++t) {
    const float* y1 = p.raw_prev + t * 6 * N + n;
    const float* y2 = p.raw_next + t * 6 * N + n;
    const float* ts = p.tspan + t * 4 * N + n;
    float* o = p.out_streams + t * 9 * N + n;
    const float* g = p.gyro + t * 3 * N + n;
    o[0] = ldg_stream(g); o[N] = ldg_stream(g + N); o[2 * N] = ldg_stream(g + 2 * N);
#pragma unroll
    for (int s = 0; s < 2; ++s) {                      // s = 0 accel, 1 mag
      const float t21 = ldg_stream(ts + (2 * s) * N), t31 = ldg_stream(ts + (2 * s + 1) * N);
      float v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = ldg_stream(y1 + (3 * s + c) * N), b = ldg_stream(y2 + (3 * s + c) * N);
        v[c] = (b - a) / t21 * t31 + a;                // Parser.cpp:264, same operation order
      }
      const float den = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);     // Parser.cpp:223-227
      Vec3<float> u = {v[0] / den, v[1] / den, v[2] / den};
      if (s == 0 && lpa) { lowpass<float>(la, u, p.alpha_acc, 1.f - p.alpha_acc); u = la; }
      if (s == 1 && lpm) { lowpass<float>(lm, u, p.alpha_mag, 1.f - p.alpha_mag); u = lm; }
      o[(3 + 3 * s) * N] = u.x; o[(4 + 3 * s) * N] = u.y; o[(5 + 3 * s) * N] = u.z;
    }
  }
  if (p.lpf_state) {
    p.lpf_state[n] = la.x; p.lpf_state[N + n] = la.y; p.lpf_state[2 * N + n] = la.z;
    p.lpf_state[3 * N + n] = lm.x; p.lpf_state[4 * N + n] = lm.y; p.lpf_state[5 * N + n] = lm.z;
  }
}

// trajectory [M][4] -> roll/pitch/yaw degrees [M][3]   (PKF/UtilityFunctions.py:3-14 per row)
__global__ void __launch_bounds__(256) traj2rpy_kernel(int64_t M, const float4* __restrict__ q, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float4 v = q[i];
  const float w = v.x, x = v.y, y = v.z, z = v.w, k = 57.29577951308232f;
  out[3 * i] = atan2f(2.f * (w * x + y * z), 1.f - 2.f * (x * x + y * y)) * k;
  out[3 * i + 1] = asinf(2.f * (w * y - z * x)) * k;
  out[3 * i + 2] = atan2f(2.f * (w * z + x * y), 1.f - 2.f * (y * y + z * z)) * k;
}

// Packed form of the rank-2 Wahba kernel: two solves per thread in f32x2 lanes (N even, per-pair or
// shared references).  Same arithmetic per solve as wahba_kernel<WAHBA_QR2>.
__global__ void __launch_bounds__(256) wahba2_kernel(const WahbaParams p) {
  const int64_t n = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (n >= p.N) return;
  const int64_t N = p.N;
  Vec3<f32x2> ra, rm;
  if (p.ref_shared) {
    ra = {f32x2(__ldg(p.acc_ref)), f32x2(__ldg(p.acc_ref + 1)), f32x2(__ldg(p.acc_ref + 2))};
    rm = {f32x2(__ldg(p.mag_ref)), f32x2(__ldg(p.mag_ref + 1)), f32x2(__ldg(p.mag_ref + 2))};
  } else {
    ra = {ld2(p.acc_ref + n), ld2(p.acc_ref + N + n), ld2(p.acc_ref + 2 * N + n)};
    rm = {ld2(p.mag_ref + n), ld2(p.mag_ref + N + n), ld2(p.mag_ref + 2 * N + n)};
  }
  auto ld2s = [](const float* q) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(q));
    return f32x2(v.x, v.y);
  };
  Vec3<f32x2> a = {ld2s(p.acc + n), ld2s(p.acc + N + n), ld2s(p.acc + 2 * N + n)};
  Vec3<f32x2> m = {ld2s(p.mag + n), ld2s(p.mag + N + n), ld2s(p.mag + 2 * N + n)};
  f32x2 ka, km;
  if (p.k_acc) { ka = ld2s(p.k_acc + n); km = ld2s(p.k_mag + n); }
  else if (p.weights_from_acc) { ka = abs_<f32x2>(a.z); km = f32x2(1.f) - ka; }
  else { ka = f32x2(p.k_acc_s); km = f32x2(p.k_mag_s); }
  const Mat3<f32x2> R = wahba_qr2<f32x2>(frame_from_pair<f32x2>(ra, rm), a, m, ka, km);
  auto st2s = [](float* q, const f32x2& v) { asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(q), "f"(v.x), "f"(v.y) : "memory"); };
  if (p.out_rot) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) st2s(p.out_rot + (3 * r + c) * N + n, R.m[r][c]);
  }
  if (p.out_quat) {
    const Quat<f32x2> q = rotation_to_quat_ref<f32x2>(R);
    st2s(p.out_quat + n, q.w); st2s(p.out_quat + N + n, q.x); st2s(p.out_quat + 2 * N + n, q.y); st2s(p.out_quat + 3 * N + n, q.z);
  }
}

__global__ void __launch_bounds__(256) rot2quat_kernel(int64_t N, const float* __restrict__ rot, float* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Mat3<float> R;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R.m[r][c] = rot[(3 * r + c) * N + n];
  Quat<float> q = rotation_to_quat_ref<float>(R);
  out[n] = q.w; out[N + n] = q.x; out[2 * N + n] = q.y; out[3 * N + n] = q.z;
}

// ---------------------------------------------------------------------------------------------
// General Prediction / Correction (full matrices, exactly the reference's operations).
// ---------------------------------------------------------------------------------------------
struct PredictParams {
  int64_t N;
  const float *gyro, *dt;
  int dt_shared;
  const float *x, *p, *q_mat, *r_mat, *q_scale, *r_scale;
  float *out_z, *out_p, *out_k;
};

__global__ void __launch_bounds__(128) predict_kernel(const PredictParams a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int64_t N = a.N;
  Vec3<float> w = {a.gyro[n], a.gyro[N + n], a.gyro[2 * N + n]};
  Quat<float> x = {a.x[n], a.x[N + n], a.x[2 * N + n], a.x[3 * N + n]};
  Mat4<float> P, A, AP, S, Si;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) P.m[i][j] = a.p[(4 * i + j) * N + n];
  half_omega<float>(w, A);                                      // GetJacobian_A  :43-48
  // GetJacobian_B(x)  :51-56
  const float B[4][3] = {{-0.5f * x.x, -0.5f * x.y, -0.5f * x.z},
                         {0.5f * x.w, 0.5f * x.z, -0.5f * x.y},
                         {-0.5f * x.z, 0.5f * x.w, 0.5f * x.x},
                         {0.5f * x.y, -0.5f * x.x, 0.5f * x.w}};
  const float qs = a.q_scale ? a.q_scale[n] : 1.f, rs = a.r_scale ? a.r_scale[n] : 1.f;
  float BQ[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc = fmaf(B[i][k], qs * __ldg(a.q_mat + 3 * k + j), acc);
      BQ[i][j] = acc;
    }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(A.m[i][k], P.m[k][j], acc);
      AP.m[i][j] = acc;
    }
  Mat4<float> Pn;                                               // P = A P A^T + B Q B^T   :61
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(AP.m[i][k], A.m[j][k], acc);
#pragma unroll
      for (int k = 0; k < 3; ++k) acc = fmaf(BQ[i][k], B[j][k], acc);
      Pn.m[i][j] = acc;
      S.m[i][j] = acc + rs * __ldg(a.r_mat + 4 * i + j);        // S = P + R   :63
    }
  const float h = a.dt_shared ? a.dt[0] : a.dt[n];
  Vec3<float> hw = {0.5f * w.x, 0.5f * w.y, 0.5f * w.z};
  Quat<float> z = rk4_step<float>(x, hw, h);                    // :62
  inverse4<float>(S, Si);                                       // :65
  a.out_z[n] = z.w; a.out_z[N + n] = z.x; a.out_z[2 * N + n] = z.y; a.out_z[3 * N + n] = z.z;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;                                          // K = P S^-1   :66
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(Pn.m[i][k], Si.m[k][j], acc);
      a.out_k[(4 * i + j) * N + n] = acc;
      a.out_p[(4 * i + j) * N + n] = Pn.m[i][j];
    }
}

struct CorrectParams {
  int64_t N;
  const float *mag, *acc, *acc_ref, *mag_ref;
  int ref_shared;
  const float *z, *p, *k;
  float *out_x, *out_p;
  uint8_t* out_flip;
  float* out_meas;
};

template <int ALGO> __global__ void __launch_bounds__(128) correct_kernel(const CorrectParams a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int64_t N = a.N;
  Vec3<float> ra, rm;
  if (a.ref_shared) {
    ra = {a.acc_ref[0], a.acc_ref[1], a.acc_ref[2]};
    rm = {a.mag_ref[0], a.mag_ref[1], a.mag_ref[2]};
  } else {
    ra = {a.acc_ref[n], a.acc_ref[N + n], a.acc_ref[2 * N + n]};
    rm = {a.mag_ref[n], a.mag_ref[N + n], a.mag_ref[2 * N + n]};
  }
  Vec3<float> ac = {a.acc[n], a.acc[N + n], a.acc[2 * N + n]};
  Vec3<float> mg = {a.mag[n], a.mag[N + n], a.mag[2 * N + n]};
  Quat<float> z = {a.z[n], a.z[N + n], a.z[2 * N + n], a.z[3 * N + n]};
  const float ka = fabsf(ac.z), km = 1.f - ka;                                     // :71
  Mat3<float> R = (ALGO == WAHBA_QR2) ? wahba_qr2<float>(frame_from_pair<float>(ra, rm), ac, mg, ka, km)
                                      : wahba_jacobi<float>(ra, rm, ac, mg, ka, km, 6);
  Quat<float> y = rotation_to_quat_ref<float>(R);
  const bool flip = dot4(y, z) < 0.f;                                              // :73-74 (Comparator[0] == dot)
  if (flip) { y.w = -y.w; y.x = -y.x; y.y = -y.y; y.z = -y.z; }
  const float e[4] = {y.w - z.w, y.x - z.x, y.y - z.y, y.z - z.z};                 // :76
  float K[4][4], P[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { K[i][j] = a.k[(4 * i + j) * N + n]; P[i][j] = a.p[(4 * i + j) * N + n]; }
  const float zz[4] = {z.w, z.x, z.y, z.z};
  float X[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float acc = zz[i];                                                             // X = z + K e   :77
#pragma unroll
    for (int j = 0; j < 4; ++j) acc = fmaf(K[i][j], e[j], acc);
    X[i] = acc;
  }
  const float inv = rsqrtf(fmaf(X[3], X[3], fmaf(X[2], X[2], fmaf(X[1], X[1], X[0] * X[0]))));
#pragma unroll
  for (int i = 0; i < 4; ++i) a.out_x[i * N + n] = X[i] * inv;                     // :79
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = P[i][j];                                                         // P = P - K P   :78
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(-K[i][k], P[k][j], acc);
      a.out_p[(4 * i + j) * N + n] = acc;
    }
  if (a.out_flip) a.out_flip[n] = flip ? 1 : 0;
  if (a.out_meas) { a.out_meas[n] = y.w; a.out_meas[N + n] = y.x; a.out_meas[2 * N + n] = y.y; a.out_meas[3 * N + n] = y.z; }
}

__global__ void __launch_bounds__(256)
    rk4_kernel(int64_t N, const float* q, const float* dt, int dt_shared, const float* w, float* out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Quat<float> x = {q[n], q[N + n], q[2 * N + n], q[3 * N + n]};
  Vec3<float> hw = {0.5f * w[n], 0.5f * w[N + n], 0.5f * w[2 * N + n]};
  Quat<float> z = rk4_step<float>(x, hw, dt_shared ? dt[0] : dt[n]);
  out[n] = z.w; out[N + n] = z.x; out[2 * N + n] = z.y; out[3 * N + n] = z.z;
}

__global__ void __launch_bounds__(256)
    jacobians_kernel(int64_t N, const float* w, float* out_a, const float* q, float* out_b) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (w && out_a) {
    Mat4<float> A;
    half_omega<float>({w[n], w[N + n], w[2 * N + n]}, A);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) out_a[(4 * i + j) * N + n] = A.m[i][j];
  }
  if (q && out_b) {
    const float q0 = 0.5f * q[n], q1 = 0.5f * q[N + n], q2 = 0.5f * q[2 * N + n], q3 = 0.5f * q[3 * N + n];
    const float B[12] = {-q1, -q2, -q3, q0, q3, -q2, -q3, q0, q1, q2, -q1, q0};
#pragma unroll
    for (int i = 0; i < 12; ++i) out_b[i * N + n] = B[i];
  }
}

__global__ void __launch_bounds__(256) comparator_kernel(int64_t N, const float* q1, const float* q2, float* out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  // conj(q1) (x) q2, written as the reference's 4x4 mat-vec (PKF/ExtendedKalmanFilter.py:17-23)
  const float c0 = q1[n], c1 = -q1[N + n], c2 = -q1[2 * N + n], c3 = -q1[3 * N + n];
  const float b0 = q2[n], b1 = q2[N + n], b2 = q2[2 * N + n], b3 = q2[3 * N + n];
  out[n] = fmaf(-c3, b3, fmaf(-c2, b2, fmaf(-c1, b1, c0 * b0)));
  out[N + n] = fmaf(c2, b3, fmaf(-c3, b2, fmaf(c0, b1, c1 * b0)));
  out[2 * N + n] = fmaf(-c1, b3, fmaf(c0, b2, fmaf(c3, b1, c2 * b0)));
  out[3 * N + n] = fmaf(c0, b3, fmaf(c1, b2, fmaf(-c2, b1, c3 * b0)));
}

__global__ void __launch_bounds__(256)
    lowpass_kernel(int64_t N, int64_t T, const float* x, float alpha, float* state, float* out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Vec3<float> y = {state[n], state[N + n], state[2 * N + n]};
  const float oma = 1.f - alpha;
  for (int64_t t = 0; t < T; ++t) {
    const float* xi = x + t * 3 * N + n;
    Vec3<float> v = {ldg_stream(xi), ldg_stream(xi + N), ldg_stream(xi + 2 * N)};
    lowpass<float>(y, v, alpha, oma);
    float* o = out + t * 3 * N + n;
    o[0] = y.x; o[N] = y.y; o[2 * N] = y.z;
  }
  state[n] = y.x; state[N + n] = y.y; state[2 * N + n] = y.z;
}

__global__ void __launch_bounds__(256) quat2rpy_kernel(int64_t N, const float* q, float* out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float w = q[n], x = q[N + n], y = q[2 * N + n], z = q[3 * N + n];
  const float k = 57.29577951308232f;
  out[n] = atan2f(2.f * (w * x + y * z), 1.f - 2.f * (x * x + y * y)) * k;
  out[N + n] = asinf(2.f * (w * y - z * x)) * k;
  out[2 * N + n] = atan2f(2.f * (w * z + x * y), 1.f - 2.f * (y * y + z * z)) * k;
}

__global__ void __launch_bounds__(256) norm_kernel(int64_t N, int k, const float* v, float* out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int i = 0; i < k; ++i) { const float e = v[(int64_t)i * N + n]; acc = fmaf(e, e, acc); }   // left to right, :18-19
  out[n] = sqrtf(acc);
}

// host-replay helpers: initial state and P <-> P/r conversion on the device
__global__ void __launch_bounds__(256)
    host_init_state_kernel(int64_t N, int have_x0, int have_p0, const float* __restrict__ r, float* x, float* p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (!have_x0) { x[n] = 1.f; x[N + n] = 0.f; x[2 * N + n] = 0.f; x[3 * N + n] = 0.f; }     // PKF/main_file.py:26
  const float ir = 1.f / r[n];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const bool diag = (k == 0 || k == 4 || k == 7 || k == 9);
    const float p0 = have_p0 ? p[k * N + n] : (diag ? 1.f : 0.f);                            // PKF/main_file.py:23
    p[k * N + n] = p0 * ir;
  }
}
__global__ void __launch_bounds__(256) host_unscale_p_kernel(int64_t N, const float* __restrict__ r, float* p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float rr = r[n];
#pragma unroll
  for (int k = 0; k < 10; ++k) p[k * N + n] *= rr;
}

// FP32 peak probe: 16 independent FFMA chains per thread, all SMs full.
constexpr int kProbeIters = 8192, kProbeAcc = 16;
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* out, float b, float c) {
  float a[kProbeAcc];
#pragma unroll
  for (int i = 0; i < kProbeAcc; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < kProbeIters; ++it) {
#pragma unroll
    for (int i = 0; i < kProbeAcc; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kProbeAcc; ++i) s += a[i];
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

bool tma_eligible(const ReplayParams& p) {
  if ((reinterpret_cast<uintptr_t>(p.streams) & 15) != 0) return false;
  if (p.Ns % 4 != 0) return false;                        // global strides must be multiples of 16 bytes
  if (p.Ns != p.N && (p.Ns % kThreads) != 0) return false;   // a CTA's 128 columns must not wrap
  if (p.T > INT32_MAX || p.Ns > INT32_MAX) return false;
  return true;
}

template <int ALGO, bool LPF, bool AUX, bool COMP> int launch_replay(const ReplayParams& p, bool use_tma, cudaStream_t st) {
  const unsigned grid = (unsigned)((p.N + kThreads - 1) / kThreads);
  if (!use_tma) {
    replay_ldg_kernel<ALGO, LPF, AUX, COMP><<<grid, kThreads, 0, st>>>(p);
    return launch_status();
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return POSEKF_ENODEV;
  CUtensorMap tmap;
  const cuuint64_t dims[3] = {(cuuint64_t)p.Ns, (cuuint64_t)kChannels, (cuuint64_t)p.T};
  const cuuint64_t strides[2] = {(cuuint64_t)p.Ns * sizeof(float), (cuuint64_t)p.Ns * kChannels * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)kThreads, (cuuint32_t)kChannels, (cuuint32_t)kTmaSteps};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.streams), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return POSEKF_EALIGN;
  auto kern = replay_tma_kernel<ALGO, LPF, AUX, COMP>;
  // idempotent; set on every launch (cheap) so that it holds on every device of a multi-GPU process
  PKF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TmaSmem)));
  kern<<<grid, kThreads, sizeof(TmaSmem), st>>>(p, tmap);
  return launch_status();
}

bool packed_eligible(const ReplayParams& p) {
  // float2 accesses to the [k][N] state / constant arrays need N even and 8-byte aligned bases
  if ((p.N & 1) != 0) return false;
  if (p.Ns != p.N && (p.Ns % kTile2) != 0) return false;      // a CTA's columns must not wrap
  const void* ptrs[] = {p.acc_ref, p.mag_ref, p.q_scale, p.r_scale, p.state_x, p.state_x_lo, p.state_p, p.state_lpf};
  for (const void* q : ptrs) if ((reinterpret_cast<uintptr_t>(q) & 7) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.loss_acc) & 7) != 0 || (reinterpret_cast<uintptr_t>(p.out_flip) & 1) != 0) return false;
  return true;      // out_traj / truth are already required to be 16-byte aligned
}

template <bool LPF, bool AUX, bool COMP> int launch_replay_packed(const ReplayParams& p, cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return POSEKF_ENODEV;
  CUtensorMap tmap;
  const cuuint64_t dims[3] = {(cuuint64_t)p.Ns, (cuuint64_t)kChannels, (cuuint64_t)p.T};
  const cuuint64_t strides[2] = {(cuuint64_t)p.Ns * sizeof(float), (cuuint64_t)p.Ns * kChannels * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)kTile2, (cuuint32_t)kChannels, (cuuint32_t)kTma2Steps};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.streams), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return POSEKF_EALIGN;
  auto kern = replay_tma2_kernel<LPF, AUX, COMP>;
  PKF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tma2Smem)));
  const unsigned grid = (unsigned)((p.N + kTile2 - 1) / kTile2);
  kern<<<grid, kThreads2, sizeof(Tma2Smem), st>>>(p, tmap);
  return launch_status();
}

template <int ALGO, bool LPF> int launch_replay_aux(const ReplayParams& p, bool use_tma, cudaStream_t st) {
  const bool aux = p.out_traj != nullptr || p.out_flip != nullptr || p.truth != nullptr;
  const bool comp = p.state_x_lo != nullptr;
  if (comp) return aux ? launch_replay<ALGO, LPF, true, true>(p, use_tma, st) : launch_replay<ALGO, LPF, false, true>(p, use_tma, st);
  return aux ? launch_replay<ALGO, LPF, true, false>(p, use_tma, st) : launch_replay<ALGO, LPF, false, false>(p, use_tma, st);
}

int replay_dispatch(const ReplayParams& p, int algo, int staging, cudaStream_t st) {
  const bool lpf = (p.alpha_acc >= 0.f) || (p.alpha_mag >= 0.f);
  bool use_tma, packed = false;
  if (staging == POSEKF_STAGE_LDG) use_tma = false;
  else if (staging == POSEKF_STAGE_TMA) { if (!tma_eligible(p)) return POSEKF_EALIGN; use_tma = true; }
  else if (staging == POSEKF_STAGE_TMA_PACKED) {
    if (!tma_eligible(p) || !packed_eligible(p) || algo != POSEKF_WAHBA_QR2) return POSEKF_EALIGN;
    use_tma = packed = true;
  } else if (staging == POSEKF_STAGE_AUTO) {
    use_tma = tma_eligible(p);
    packed = use_tma && kAutoPrefersPacked && algo == POSEKF_WAHBA_QR2 && packed_eligible(p);
  } else return POSEKF_EINVAL;
  if (packed) {
    const bool comp = p.state_x_lo != nullptr;
    const bool aux = p.out_traj != nullptr || p.out_flip != nullptr || p.truth != nullptr;
    if (lpf) {
      if (aux) return comp ? launch_replay_packed<true, true, true>(p, st) : launch_replay_packed<true, true, false>(p, st);
      return comp ? launch_replay_packed<true, false, true>(p, st) : launch_replay_packed<true, false, false>(p, st);
    }
    if (aux) return comp ? launch_replay_packed<false, true, true>(p, st) : launch_replay_packed<false, true, false>(p, st);
    return comp ? launch_replay_packed<false, false, true>(p, st) : launch_replay_packed<false, false, false>(p, st);
  }
  if (algo == POSEKF_WAHBA_QR2) return lpf ? launch_replay_aux<WAHBA_QR2, true>(p, use_tma, st) : launch_replay_aux<WAHBA_QR2, false>(p, use_tma, st);
  if (algo == POSEKF_WAHBA_JACOBI) return lpf ? launch_replay_aux<WAHBA_JACOBI, true>(p, use_tma, st) : launch_replay_aux<WAHBA_JACOBI, false>(p, use_tma, st);
  return POSEKF_EINVAL;
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* posekf_version(void) { return "posekf_b200 0.1 sm_100a"; }

int posekf_replay_f32(int64_t n_filters, int64_t n_steps, const float* streams, int64_t n_streams, const float* dt,
                      int dt_per_step, const float* acc_ref, const float* mag_ref, const float* q_scale,
                      const float* r_scale, float lpf_alpha_acc, float lpf_alpha_mag, float* state_x, float* state_x_lo,
                      float* state_p, float* state_lpf, float* out_traj, uint8_t* out_flip, const float* truth, float* loss_acc,
                      int wahba_algo, int staging, void* stream) {
  if (n_filters < 0 || n_steps < 0 || n_streams <= 0 && n_filters > 0) return POSEKF_EINVAL;
  if (n_filters == 0 || n_steps == 0) return 0;
  if (!streams || !dt || !acc_ref || !mag_ref || !q_scale || !r_scale || !state_x || !state_p) return POSEKF_EINVAL;
  if (n_streams > n_filters || (n_filters % n_streams) != 0) return POSEKF_EINVAL;
  const bool lpf = lpf_alpha_acc >= 0.f || lpf_alpha_mag >= 0.f;
  if (lpf && !state_lpf) return POSEKF_EINVAL;
  if ((n_filters + kThreads - 1) / kThreads > 0x7fffffffLL || n_steps > 0x7fffffffLL) return POSEKF_EINVAL;
  if (out_traj && (reinterpret_cast<uintptr_t>(out_traj) & 15) != 0) return POSEKF_EALIGN;
  if (truth && (!loss_acc || (reinterpret_cast<uintptr_t>(truth) & 15) != 0)) return truth && !loss_acc ? POSEKF_EINVAL : POSEKF_EALIGN;
  ReplayParams p;
  p.N = n_filters; p.T = n_steps; p.Ns = n_streams; p.streams = streams; p.dt = dt; p.dt_per_step = dt_per_step;
  p.acc_ref = acc_ref; p.mag_ref = mag_ref; p.q_scale = q_scale; p.r_scale = r_scale;
  p.alpha_acc = lpf_alpha_acc; p.alpha_mag = lpf_alpha_mag;
  p.state_x = state_x; p.state_x_lo = state_x_lo; p.state_p = state_p; p.state_lpf = state_lpf; p.out_traj = out_traj; p.out_flip = out_flip;
  p.truth = truth; p.loss_acc = loss_acc;
  return replay_dispatch(p, wahba_algo, staging, (cudaStream_t)stream);
}

// ---- host-buffer replay: workspace (device staging buffers, streams, events) ----------------------
struct HostWorkspace {
  int device = 0;
  int64_t N = 0, chunk_steps = 0;
  bool traj = false;
  cudaStream_t s_copy = nullptr, s_comp = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_traj[2] = {nullptr, nullptr},
              ev_tfree[2] = {nullptr, nullptr};
  float *d_in[2] = {nullptr, nullptr}, *d_traj[2] = {nullptr, nullptr};
  float *d_ref = nullptr, *d_qr = nullptr, *d_x = nullptr, *d_xlo = nullptr, *d_p = nullptr, *d_lpf = nullptr, *d_dt = nullptr;
};

static void host_ws_free(HostWorkspace* w) {
  if (!w) return;
  cudaSetDevice(w->device);
  for (int i = 0; i < 2; ++i) {
    if (w->d_in[i]) cudaFree(w->d_in[i]);
    if (w->d_traj[i]) cudaFree(w->d_traj[i]);
    if (w->ev_in[i]) cudaEventDestroy(w->ev_in[i]);
    if (w->ev_free[i]) cudaEventDestroy(w->ev_free[i]);
    if (w->ev_traj[i]) cudaEventDestroy(w->ev_traj[i]);
    if (w->ev_tfree[i]) cudaEventDestroy(w->ev_tfree[i]);
  }
  float* ptrs[] = {w->d_ref, w->d_qr, w->d_x, w->d_xlo, w->d_p, w->d_lpf, w->d_dt};
  for (float* q : ptrs) if (q) cudaFree(q);
  if (w->s_copy) cudaStreamDestroy(w->s_copy);
  if (w->s_comp) cudaStreamDestroy(w->s_comp);
  if (w->s_out) cudaStreamDestroy(w->s_out);
  delete w;
}

int posekf_host_workspace_create(int device, int64_t n_filters, int64_t chunk_steps, int with_trajectory, void** out_ws) {
  if (!out_ws || n_filters <= 0) return POSEKF_EINVAL;
  *out_ws = nullptr;
  PKF_CUDA_TRY(cudaSetDevice(device));
  const int64_t N = n_filters;
  if (chunk_steps <= 0) {   // ~256 MiB per staging buffer: small enough that the first kernel starts after ~5 ms
    const int64_t bytes_per_step = (int64_t)kChannels * N * sizeof(float);
    chunk_steps = std::max<int64_t>(1, (int64_t)(256ll << 20) / bytes_per_step);
  }
  HostWorkspace* w = new HostWorkspace();
  w->device = device; w->N = N; w->chunk_steps = chunk_steps; w->traj = with_trajectory != 0;
#define WS_TRY(expr)                                                   \
  do {                                                                 \
    cudaError_t _e = (expr);                                           \
    if (_e != cudaSuccess) { host_ws_free(w); return (int)_e; }        \
  } while (0)
  WS_TRY(cudaStreamCreateWithFlags(&w->s_copy, cudaStreamNonBlocking));
  WS_TRY(cudaStreamCreateWithFlags(&w->s_comp, cudaStreamNonBlocking));
  WS_TRY(cudaStreamCreateWithFlags(&w->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    WS_TRY(cudaEventCreateWithFlags(&w->ev_in[i], cudaEventDisableTiming));
    WS_TRY(cudaEventCreateWithFlags(&w->ev_free[i], cudaEventDisableTiming));
    WS_TRY(cudaEventCreateWithFlags(&w->ev_traj[i], cudaEventDisableTiming));
    WS_TRY(cudaEventCreateWithFlags(&w->ev_tfree[i], cudaEventDisableTiming));
    WS_TRY(cudaMalloc(&w->d_in[i], (size_t)chunk_steps * kChannels * N * sizeof(float)));
    if (w->traj) WS_TRY(cudaMalloc(&w->d_traj[i], (size_t)chunk_steps * 4 * N * sizeof(float)));
  }
  WS_TRY(cudaMalloc(&w->d_ref, (size_t)6 * N * sizeof(float)));
  WS_TRY(cudaMalloc(&w->d_qr, (size_t)2 * N * sizeof(float)));
  WS_TRY(cudaMalloc(&w->d_x, (size_t)4 * N * sizeof(float)));
  WS_TRY(cudaMalloc(&w->d_xlo, (size_t)4 * N * sizeof(float)));
  WS_TRY(cudaMalloc(&w->d_p, (size_t)10 * N * sizeof(float)));
  WS_TRY(cudaMalloc(&w->d_lpf, (size_t)6 * N * sizeof(float)));
  WS_TRY(cudaMalloc(&w->d_dt, sizeof(float)));
#undef WS_TRY
  *out_ws = w;
  return 0;
}

int posekf_host_workspace_destroy(void* ws) {
  host_ws_free(static_cast<HostWorkspace*>(ws));
  return 0;
}

int posekf_replay_host_f32(int64_t N, int64_t T, const float* streams_host, float dt, const float* acc_ref_host,
                           const float* mag_ref_host, const float* q_scale_host, const float* r_scale_host,
                           float lpf_alpha_acc, float lpf_alpha_mag, const float* x0_host, const float* p0_host,
                           float* out_x_host, float* out_p_host, float* out_traj_host, int64_t chunk_steps,
                           int wahba_algo, int precise, int device, void* workspace) {
  if (N <= 0 || T < 0 || !streams_host || !acc_ref_host || !mag_ref_host || !q_scale_host || !r_scale_host || !out_x_host)
    return POSEKF_EINVAL;
  const bool traj = out_traj_host != nullptr;
  HostWorkspace* w = static_cast<HostWorkspace*>(workspace);
  bool own = false;
  if (w) {
    if (w->N != N || w->device != device || (traj && !w->traj)) return POSEKF_EINVAL;
  } else {
    void* tmp = nullptr;
    int rc0 = posekf_host_workspace_create(device, N, chunk_steps, traj ? 1 : 0, &tmp);
    if (rc0 != 0) return rc0;
    w = static_cast<HostWorkspace*>(tmp);
    own = true;
  }
  PKF_CUDA_TRY(cudaSetDevice(device));
  chunk_steps = w->chunk_steps;
  const bool lpf = lpf_alpha_acc >= 0.f || lpf_alpha_mag >= 0.f;
  int rc = 0;
#define TRY(expr)                                                                  \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) { rc = (int)_e; if (own) host_ws_free(w); return rc; }  \
  } while (0)
  cudaStream_t s_copy = w->s_copy, s_comp = w->s_comp, s_out = w->s_out;
  float* d_ref = w->d_ref; float* d_qr = w->d_qr; float* d_x = w->d_x; float* d_p = w->d_p; float* d_dt = w->d_dt;
  float* d_lpf = lpf ? w->d_lpf : nullptr;
  float* d_xlo = precise ? w->d_xlo : nullptr;
  if (lpf) TRY(cudaMemsetAsync(d_lpf, 0, (size_t)6 * N * sizeof(float), s_comp));
  if (precise) TRY(cudaMemsetAsync(d_xlo, 0, (size_t)4 * N * sizeof(float), s_comp));
  TRY(cudaMemcpyAsync(d_ref, acc_ref_host, (size_t)3 * N * sizeof(float), cudaMemcpyHostToDevice, s_comp));
  TRY(cudaMemcpyAsync(d_ref + 3 * N, mag_ref_host, (size_t)3 * N * sizeof(float), cudaMemcpyHostToDevice, s_comp));
  TRY(cudaMemcpyAsync(d_qr, q_scale_host, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, s_comp));
  TRY(cudaMemcpyAsync(d_qr + N, r_scale_host, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, s_comp));
  TRY(cudaMemcpyAsync(d_dt, &dt, sizeof(float), cudaMemcpyHostToDevice, s_comp));
  // the first stream chunk is independent of the state set-up: start it right away
  const int64_t n_chunks = (T + chunk_steps - 1) / chunk_steps;
  auto issue_copy = [&](int64_t c) -> cudaError_t {
    const int b = (int)(c & 1);
    const int64_t t0 = c * chunk_steps, tc = std::min<int64_t>(chunk_steps, T - t0);
    cudaError_t e;
    if (c >= 2 && (e = cudaStreamWaitEvent(s_copy, w->ev_free[b], 0)) != cudaSuccess) return e;   // kernel of chunk c-2 done
    if ((e = cudaMemcpyAsync(w->d_in[b], streams_host + (size_t)t0 * kChannels * N, (size_t)tc * kChannels * N * sizeof(float),
                             cudaMemcpyHostToDevice, s_copy)) != cudaSuccess) return e;
    return cudaEventRecord(w->ev_in[b], s_copy);
  };
  if (n_chunks > 0) TRY(issue_copy(0));
  // initial state: X = [1,0,0,0], P = I4 (PKF/main_file.py:23,26) unless given; the device state holds P/r
  if (x0_host) TRY(cudaMemcpyAsync(d_x, x0_host, (size_t)4 * N * sizeof(float), cudaMemcpyHostToDevice, s_comp));
  if (p0_host) TRY(cudaMemcpyAsync(d_p, p0_host, (size_t)10 * N * sizeof(float), cudaMemcpyHostToDevice, s_comp));
  host_init_state_kernel<<<blocks_for(N, 256), 256, 0, s_comp>>>(N, x0_host != nullptr, p0_host != nullptr, d_qr + N, d_x, d_p);
  TRY(cudaPeekAtLastError());
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int b = (int)(c & 1);
    const int64_t t0 = c * chunk_steps, tc = std::min<int64_t>(chunk_steps, T - t0);
    if (c + 1 < n_chunks) TRY(issue_copy(c + 1));                             // keep the copy engine one chunk ahead
    TRY(cudaStreamWaitEvent(s_comp, w->ev_in[b], 0));
    if (traj && c >= 2) TRY(cudaStreamWaitEvent(s_comp, w->ev_tfree[b], 0));  // D2H of chunk c-2 done with d_traj[b]
    rc = posekf_replay_f32(N, tc, w->d_in[b], N, d_dt, 0, d_ref, d_ref + 3 * N, d_qr, d_qr + N, lpf_alpha_acc,
                           lpf_alpha_mag, d_x, d_xlo, d_p, d_lpf, traj ? w->d_traj[b] : nullptr, nullptr, nullptr, nullptr,
                           wahba_algo, POSEKF_STAGE_AUTO, s_comp);
    if (rc != 0) { if (own) host_ws_free(w); return rc; }
    TRY(cudaEventRecord(w->ev_free[b], s_comp));
    if (traj) {
      TRY(cudaEventRecord(w->ev_traj[b], s_comp));
      TRY(cudaStreamWaitEvent(s_out, w->ev_traj[b], 0));
      TRY(cudaMemcpyAsync(out_traj_host + (size_t)t0 * 4 * N, w->d_traj[b], (size_t)tc * 4 * N * sizeof(float),
                          cudaMemcpyDeviceToHost, s_out));
      TRY(cudaEventRecord(w->ev_tfree[b], s_out));
    }
  }
  TRY(cudaMemcpyAsync(out_x_host, d_x, (size_t)4 * N * sizeof(float), cudaMemcpyDeviceToHost, s_comp));
  if (out_p_host) {
    host_unscale_p_kernel<<<blocks_for(N, 256), 256, 0, s_comp>>>(N, d_qr + N, d_p);     // P/r -> P
    TRY(cudaPeekAtLastError());
    TRY(cudaMemcpyAsync(out_p_host, d_p, (size_t)10 * N * sizeof(float), cudaMemcpyDeviceToHost, s_comp));
  }
  TRY(cudaStreamSynchronize(s_comp));
  TRY(cudaStreamSynchronize(s_out));
  TRY(cudaStreamSynchronize(s_copy));
#undef TRY
  if (own) host_ws_free(w);
  return 0;
}

int posekf_wahba_f32(int64_t n, const float* acc_ref, const float* mag_ref, int ref_shared, const float* acc,
                     const float* mag, const float* k_acc, const float* k_mag, float k_acc_s, float k_mag_s,
                     int weights_from_acc, float* out_rot, float* out_quat, int wahba_algo, int jacobi_sweeps,
                     void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!acc_ref || !mag_ref || !acc || !mag || (!out_rot && !out_quat) || ((k_acc == nullptr) != (k_mag == nullptr)))
    return POSEKF_EINVAL;
  WahbaParams p{n, acc_ref, mag_ref, ref_shared, acc, mag, k_acc, k_mag, k_acc_s, k_mag_s, weights_from_acc,
                out_rot, out_quat, jacobi_sweeps > 0 ? jacobi_sweeps : 6};
  cudaStream_t st = (cudaStream_t)stream;
  if (wahba_algo == POSEKF_WAHBA_QR2) {
    // two solves per thread (packed f32x2) when every [k][N] array can be read as float2
    bool packed = (n & 1) == 0;
    const void* ptrs[] = {acc, mag, k_acc, k_mag, out_rot, out_quat, ref_shared ? nullptr : acc_ref, ref_shared ? nullptr : mag_ref};
    for (const void* q : ptrs) packed = packed && (reinterpret_cast<uintptr_t>(q) & 7) == 0;
    if (packed) wahba2_kernel<<<blocks_for(n / 2, 256), 256, 0, st>>>(p);
    else wahba_kernel<WAHBA_QR2><<<blocks_for(n, 256), 256, 0, st>>>(p);
  }
  else if (wahba_algo == POSEKF_WAHBA_JACOBI) wahba_kernel<WAHBA_JACOBI><<<blocks_for(n, 256), 256, 0, st>>>(p);
  else return POSEKF_EINVAL;
  return launch_status();
}

int posekf_tracks_f32(int64_t n_filters, int64_t n_steps, const float* streams, int64_t n_streams, const float* dt,
                      int dt_per_step, const float* acc_ref, const float* mag_ref, float k_acc, float k_mag,
                      int weights_from_acc, float* gyro_state, float* out_gyro, float* out_wahba, int wahba_algo,
                      void* stream) {
  if (n_filters < 0 || n_steps < 0) return POSEKF_EINVAL;
  if (n_filters == 0 || n_steps == 0) return 0;
  if (!streams || !dt || n_streams <= 0 || n_streams > n_filters || (n_filters % n_streams) != 0) return POSEKF_EINVAL;
  if (!out_gyro && !out_wahba && !gyro_state) return POSEKF_EINVAL;
  if (out_wahba && (!acc_ref || !mag_ref)) return POSEKF_EINVAL;
  if (((reinterpret_cast<uintptr_t>(out_gyro) | reinterpret_cast<uintptr_t>(out_wahba)) & 15) != 0) return POSEKF_EALIGN;
  TracksParams p{n_filters, n_steps, n_streams, streams, dt, dt_per_step, acc_ref ? acc_ref : streams,
                 mag_ref ? mag_ref : streams, k_acc, k_mag, weights_from_acc, gyro_state, out_gyro, out_wahba};
  cudaStream_t st = (cudaStream_t)stream;
  if (wahba_algo == POSEKF_WAHBA_QR2) tracks_kernel<WAHBA_QR2><<<blocks_for(n_filters, 128), 128, 0, st>>>(p);
  else if (wahba_algo == POSEKF_WAHBA_JACOBI) tracks_kernel<WAHBA_JACOBI><<<blocks_for(n_filters, 128), 128, 0, st>>>(p);
  else return POSEKF_EINVAL;
  return launch_status();
}

int posekf_preprocess_f32(int64_t n_filters, int64_t n_steps, const float* gyro, const float* raw_prev,
                          const float* raw_next, const float* tspan, float lpf_alpha_acc, float lpf_alpha_mag,
                          float* lpf_state, float* out_streams, void* stream) {
  if (n_filters < 0 || n_steps < 0) return POSEKF_EINVAL;
  if (n_filters == 0 || n_steps == 0) return 0;
  if (!gyro || !raw_prev || !raw_next || !tspan || !out_streams) return POSEKF_EINVAL;
  if ((lpf_alpha_acc >= 0.f || lpf_alpha_mag >= 0.f) && !lpf_state) return POSEKF_EINVAL;
  PreprocessParams p{n_filters, n_steps, gyro, raw_prev, raw_next, tspan, lpf_alpha_acc, lpf_alpha_mag, lpf_state, out_streams};
  preprocess_kernel<<<blocks_for(n_filters, 256), 256, 0, (cudaStream_t)stream>>>(p);
  return launch_status();
}

int posekf_traj2rpy_f32(int64_t m, const float* traj, float* out_rpy_deg, void* stream) {
  if (m < 0) return POSEKF_EINVAL;
  if (m == 0) return 0;
  if (!traj || !out_rpy_deg) return POSEKF_EINVAL;
  if ((reinterpret_cast<uintptr_t>(traj) & 15) != 0) return POSEKF_EALIGN;
  traj2rpy_kernel<<<blocks_for(m, 256), 256, 0, (cudaStream_t)stream>>>(m, reinterpret_cast<const float4*>(traj), out_rpy_deg);
  return launch_status();
}

int posekf_rot2quat_f32(int64_t n, const float* rot, float* out_quat, void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!rot || !out_quat) return POSEKF_EINVAL;
  rot2quat_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, rot, out_quat);
  return launch_status();
}

int posekf_predict_f32(int64_t n, const float* gyro, const float* dt, int dt_shared, const float* x, const float* p,
                       const float* q_mat, const float* r_mat, const float* q_scale, const float* r_scale, float* out_z,
                       float* out_p, float* out_k, void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!gyro || !dt || !x || !p || !q_mat || !r_mat || !out_z || !out_p || !out_k) return POSEKF_EINVAL;
  PredictParams a{n, gyro, dt, dt_shared, x, p, q_mat, r_mat, q_scale, r_scale, out_z, out_p, out_k};
  predict_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(a);
  return launch_status();
}

int posekf_correct_f32(int64_t n, const float* mag, const float* acc, const float* acc_ref, const float* mag_ref,
                       int ref_shared, const float* z, const float* p, const float* k, float* out_x, float* out_p,
                       uint8_t* out_flip, float* out_meas, int wahba_algo, void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!mag || !acc || !acc_ref || !mag_ref || !z || !p || !k || !out_x || !out_p) return POSEKF_EINVAL;
  CorrectParams a{n, mag, acc, acc_ref, mag_ref, ref_shared, z, p, k, out_x, out_p, out_flip, out_meas};
  cudaStream_t st = (cudaStream_t)stream;
  if (wahba_algo == POSEKF_WAHBA_QR2) correct_kernel<WAHBA_QR2><<<blocks_for(n, 128), 128, 0, st>>>(a);
  else if (wahba_algo == POSEKF_WAHBA_JACOBI) correct_kernel<WAHBA_JACOBI><<<blocks_for(n, 128), 128, 0, st>>>(a);
  else return POSEKF_EINVAL;
  return launch_status();
}

int posekf_rk4_f32(int64_t n, const float* q, const float* dt, int dt_shared, const float* w, float* out_q, void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!q || !dt || !w || !out_q) return POSEKF_EINVAL;
  rk4_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, q, dt, dt_shared, w, out_q);
  return launch_status();
}

int posekf_jacobians_f32(int64_t n, const float* w, float* out_a, const float* q, float* out_b, void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!((w && out_a) || (q && out_b))) return POSEKF_EINVAL;
  jacobians_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, w, out_a, q, out_b);
  return launch_status();
}

int posekf_comparator_f32(int64_t n, const float* q1, const float* q2, float* out, void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!q1 || !q2 || !out) return POSEKF_EINVAL;
  comparator_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, q1, q2, out);
  return launch_status();
}

int posekf_lowpass_f32(int64_t n, int64_t n_steps, const float* x, float alpha, float* state, float* out, void* stream) {
  if (n < 0 || n_steps < 0) return POSEKF_EINVAL;
  if (n == 0 || n_steps == 0) return 0;
  if (!x || !state || !out) return POSEKF_EINVAL;
  lowpass_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, n_steps, x, alpha, state, out);
  return launch_status();
}

int posekf_quat2rpy_f32(int64_t n, const float* q, float* out_rpy_deg, void* stream) {
  if (n < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!q || !out_rpy_deg) return POSEKF_EINVAL;
  quat2rpy_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, q, out_rpy_deg);
  return launch_status();
}

int posekf_norm_f32(int64_t n, int k, const float* v, float* out, void* stream) {
  if (n < 0 || k < 0) return POSEKF_EINVAL;
  if (n == 0) return 0;
  if (!v || !out) return POSEKF_EINVAL;
  norm_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, k, v, out);
  return launch_status();
}

int posekf_copy_async(void* dst, const void* src, int64_t bytes, int to_device, void* stream) {
  if (bytes < 0 || (bytes > 0 && (!dst || !src))) return POSEKF_EINVAL;
  if (bytes == 0) return 0;
  PKF_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
  return 0;
}

int posekf_stream_sync(void* stream) {
  PKF_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int posekf_fp32_peak_tflops(int device, double* out_tflops, double* out_ms) {
  if (!out_tflops) return POSEKF_EINVAL;
  PKF_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  PKF_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8;
  float* out = nullptr;
  PKF_CUDA_TRY(cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) fp32_probe_kernel<<<blocks, 256>>>(out, 1.0001f, 1e-4f);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    fp32_probe_kernel<<<blocks, 256>>>(out, 1.0001f, 1e-4f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const double flops = 2.0 * kProbeIters * kProbeAcc * (double)blocks * 256;
  *out_tflops = flops / (best * 1e-3) / 1e12;
  if (out_ms) *out_ms = best;
  return 0;
}

}  // extern "C"
