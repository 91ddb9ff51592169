// hostsim.cpp -- CPU build of poseestimationkf_b200/csrc/ekf_math.cuh for numerics probing.
//
// TEST TOOL ONLY.  The product has no CPU path; this file exists so that the arithmetic of the
// device header can be exercised in float32 (and float64) inside the GPU-less build container and
// by the `-m "not gpu"` test-suite.  It is compiled by tests/hostsim/build.py with g++ and loaded
// with ctypes by tests only.
#include <cstdint>
#include <cstddef>
#include "../../poseestimationkf_b200/csrc/ekf_math.cuh"

using namespace pkf;

template <typename F, int ALGO, bool COMP>
static void replay_t(int64_t N, int64_t T, const float* streams, const double* dt, int dt_per_step,
                     const float* acc_ref, const float* mag_ref, const float* q, const float* r,
                     float lpf_acc, float lpf_mag, double* out_traj, uint8_t* out_flip, double* out_P,
                     const float* x0 = nullptr, const float* p0 = nullptr) {
  for (int64_t n = 0; n < N; ++n) {
    Vec3<F> ra = {(F)acc_ref[0 * N + n], (F)acc_ref[1 * N + n], (F)acc_ref[2 * N + n]};
    Vec3<F> rm = {(F)mag_ref[0 * N + n], (F)mag_ref[1 * N + n], (F)mag_ref[2 * N + n]};
    FilterConst<F> fc = make_filter_const<F>(ra, rm, (F)q[n], (F)r[n]);
    Quat<F> x = {F(1), F(0), F(0), F(0)}, xlo = {F(0), F(0), F(0), F(0)};
    const F ir = F(1) / (F)r[n];   // the step carries P/r
    Sym4<F> P = {ir, F(0), F(0), F(0), ir, F(0), F(0), ir, F(0), ir};
    Vec3<F> la = {F(0), F(0), F(0)}, lm = {F(0), F(0), F(0)};
    if (x0) x = {(F)x0[n], (F)x0[N + n], (F)x0[2 * N + n], (F)x0[3 * N + n]};          // state buffers as the kernels take them:
    if (p0) P = {(F)p0[n], (F)p0[N + n], (F)p0[2 * N + n], (F)p0[3 * N + n], (F)p0[4 * N + n], (F)p0[5 * N + n],   // X [4][N], P/r [10][N]
                 (F)p0[6 * N + n], (F)p0[7 * N + n], (F)p0[8 * N + n], (F)p0[9 * N + n]};
    adopt_state(fc, x, xlo);
    enter_filter_frame(fc, x, xlo, P, COMP);     // as the kernels do at the start of a launch
    for (int64_t t = 0; t < T; ++t) {
      const float* s = streams + (size_t)t * 9 * N + n;
      Vec3<F> w = {(F)s[0 * N], (F)s[1 * N], (F)s[2 * N]};
      Vec3<F> a = {(F)s[3 * N], (F)s[4 * N], (F)s[5 * N]};
      Vec3<F> m = {(F)s[6 * N], (F)s[7 * N], (F)s[8 * N]};
      if (lpf_acc >= 0.f) { lowpass<F>(la, a, (F)lpf_acc, F(1) - (F)lpf_acc); a = la; }
      if (lpf_mag >= 0.f) { lowpass<F>(lm, m, (F)lpf_mag, F(1) - (F)lpf_mag); m = lm; }
      F h = (F)(dt_per_step ? dt[t] : dt[0]);
      bool flip;
      ekf_step<F, ALGO, true, COMP>(x, xlo, P, fc, w, a, m, StepH<F>(h), flip);
      if (out_traj) {
        const Quat<F> xr = state_in_reference_frame(fc, x);
        double* o = out_traj + (size_t)t * 4 * N + n;
        o[0 * N] = xr.w; o[1 * N] = xr.x; o[2 * N] = xr.y; o[3 * N] = xr.z;
      }
      if (out_flip) out_flip[(size_t)t * N + n] = flip;
    }
    leave_filter_frame(fc, x, xlo, P, COMP);
    if (out_P) {
      const F p[10] = {P.a00, P.a01, P.a02, P.a03, P.a11, P.a12, P.a13, P.a22, P.a23, P.a33};
      for (int k = 0; k < 10; ++k) out_P[(size_t)k * N + n] = (F)r[n] * p[k];
    }
  }
}

// Packed (two filters per "thread", F = f32x2) form of the same replay: exercises the lane logic of the
// packed device kernel on the CPU.  N must be even.
template <bool COMP>
static void replay_packed_t(int64_t N, int64_t T, const float* streams, const double* dt, int dt_per_step,
                            const float* acc_ref, const float* mag_ref, const float* q, const float* r, float lpf_acc,
                            float lpf_mag, double* out_traj, double* out_P, uint8_t* out_flip) {
  typedef f32x2 F;
  for (int64_t n = 0; n < N; n += 2) {
    Vec3<F> ra = {F(acc_ref[n], acc_ref[n + 1]), F(acc_ref[N + n], acc_ref[N + n + 1]), F(acc_ref[2 * N + n], acc_ref[2 * N + n + 1])};
    Vec3<F> rm = {F(mag_ref[n], mag_ref[n + 1]), F(mag_ref[N + n], mag_ref[N + n + 1]), F(mag_ref[2 * N + n], mag_ref[2 * N + n + 1])};
    FilterConst<F> fc = make_filter_const<F>(ra, rm, F(q[n], q[n + 1]), F(r[n], r[n + 1]));
    Quat<F> x = {F(1.f), F(0.f), F(0.f), F(0.f)}, xlo = {F(0.f), F(0.f), F(0.f), F(0.f)};
    const F ir = F(1.f / r[n], 1.f / r[n + 1]);
    Sym4<F> P = {ir, F(0.f), F(0.f), F(0.f), ir, F(0.f), F(0.f), ir, F(0.f), ir};
    Vec3<F> la = {F(0.f), F(0.f), F(0.f)}, lm = {F(0.f), F(0.f), F(0.f)};
    enter_filter_frame(fc, x, xlo, P, COMP);
    for (int64_t t = 0; t < T; ++t) {
      const float* s = streams + (size_t)t * 9 * N + n;
      Vec3<F> w = {F(s[0], s[1]), F(s[N], s[N + 1]), F(s[2 * N], s[2 * N + 1])};
      Vec3<F> a = {F(s[3 * N], s[3 * N + 1]), F(s[4 * N], s[4 * N + 1]), F(s[5 * N], s[5 * N + 1])};
      Vec3<F> m = {F(s[6 * N], s[6 * N + 1]), F(s[7 * N], s[7 * N + 1]), F(s[8 * N], s[8 * N + 1])};
      if (lpf_acc >= 0.f) { lowpass<F>(la, a, F(lpf_acc), F(1.f - lpf_acc)); a = la; }
      if (lpf_mag >= 0.f) { lowpass<F>(lm, m, F(lpf_mag), F(1.f - lpf_mag)); m = lm; }
      F h = F((float)(dt_per_step ? dt[t] : dt[0]));
      mask2 flip;
      ekf_step<F, WAHBA_QR2, true, COMP>(x, xlo, P, fc, w, a, m, StepH<F>(h), flip);
      if (out_flip) { out_flip[(size_t)t * N + n] = flip.x; out_flip[(size_t)t * N + n + 1] = flip.y; }
      if (out_traj) {
        const Quat<F> xr = state_in_reference_frame(fc, x);
        double* o = out_traj + (size_t)t * 4 * N + n;
        o[0] = xr.w.x; o[1] = xr.w.y; o[N] = xr.x.x; o[N + 1] = xr.x.y; o[2 * N] = xr.y.x; o[2 * N + 1] = xr.y.y;
        o[3 * N] = xr.z.x; o[3 * N + 1] = xr.z.y;
      }
    }
    leave_filter_frame(fc, x, xlo, P, COMP);
    if (out_P) {
      const F p[10] = {P.a00, P.a01, P.a02, P.a03, P.a11, P.a12, P.a13, P.a22, P.a23, P.a33};
      for (int k = 0; k < 10; ++k) { out_P[(size_t)k * N + n] = r[n] * p[k].x; out_P[(size_t)k * N + n + 1] = r[n + 1] * p[k].y; }
    }
  }
}

extern "C" {

int hostsim_replay_packed(int comp, int64_t N, int64_t T, const float* streams, const double* dt, int dt_per_step,
                          const float* acc_ref, const float* mag_ref, const float* q, const float* r, float lpf_acc,
                          float lpf_mag, double* out_traj, double* out_P, uint8_t* out_flip) {
  if (N % 2) return 1;
  if (comp) replay_packed_t<true>(N, T, streams, dt, dt_per_step, acc_ref, mag_ref, q, r, lpf_acc, lpf_mag, out_traj, out_P, out_flip);
  else replay_packed_t<false>(N, T, streams, dt, dt_per_step, acc_ref, mag_ref, q, r, lpf_acc, lpf_mag, out_traj, out_P, out_flip);
  return 0;
}

// precision: 0 = float32, 1 = float64 ; algo: 0 = QR2, 1 = Jacobi
int hostsim_replay(int precision, int algo, int comp, int64_t N, int64_t T, const float* streams, const double* dt,
                   int dt_per_step, const float* acc_ref, const float* mag_ref, const float* q, const float* r,
                   float lpf_acc, float lpf_mag, double* out_traj, uint8_t* out_flip, double* out_P, const float* x0,
                   const float* p0) {
#define GO(F, A) if (comp) replay_t<F, A, true>(N, T, streams, dt, dt_per_step, acc_ref, mag_ref, q, r, lpf_acc, lpf_mag, out_traj, out_flip, out_P, x0, p0); else replay_t<F, A, false>(N, T, streams, dt, dt_per_step, acc_ref, mag_ref, q, r, lpf_acc, lpf_mag, out_traj, out_flip, out_P, x0, p0)
  if (precision == 0 && algo == 0) GO(float, WAHBA_QR2);
  else if (precision == 0 && algo == 1) GO(float, WAHBA_JACOBI);
  else if (precision == 1 && algo == 0) GO(double, WAHBA_QR2);
  else if (precision == 1 && algo == 1) GO(double, WAHBA_JACOBI);
  else return 1;
#undef GO
  return 0;
}

// Wahba only: inputs [3][N] each; out_q [4][N], out_R [9][N] (row-major entries) or null
int hostsim_wahba(int precision, int algo, int sweeps, int64_t N, const float* acc_ref, const float* mag_ref,
                  const float* acc, const float* mag, const float* ka, const float* km, double* out_q,
                  double* out_R) {
  for (int64_t n = 0; n < N; ++n) {
    auto run = [&](auto tag) {
      using F = decltype(tag);
      Vec3<F> ra = {(F)acc_ref[n], (F)acc_ref[N + n], (F)acc_ref[2 * N + n]};
      Vec3<F> rm = {(F)mag_ref[n], (F)mag_ref[N + n], (F)mag_ref[2 * N + n]};
      Vec3<F> a = {(F)acc[n], (F)acc[N + n], (F)acc[2 * N + n]};
      Vec3<F> m = {(F)mag[n], (F)mag[N + n], (F)mag[2 * N + n]};
      if (algo == 2) {
        // the fused step's measurement: closed-form quaternion in the filter frame, mapped back by qE (sign arbitrary)
        FilterConst<F> fc = make_filter_const<F>(ra, rm, F(1), F(1));
        F inv;
        Quat<F> yl = wahba_quat2_local<F>(fc.E, a, m, (F)ka[n], (F)km[n], inv);
        yl.w *= inv; yl.x *= inv; yl.y *= inv; yl.z *= inv;
        Quat<F> qq = qmul(fc.qE, yl);
        out_q[n] = qq.w; out_q[N + n] = qq.x; out_q[2 * N + n] = qq.y; out_q[3 * N + n] = qq.z;
        return;
      }
      Mat3<F> R = algo == 0 ? wahba_qr2<F>(frame_from_pair<F>(ra, rm), a, m, (F)ka[n], (F)km[n])
                            : wahba_jacobi<F>(ra, rm, a, m, (F)ka[n], (F)km[n], sweeps);
      Quat<F> qq = rotation_to_quat_ref<F>(R);
      out_q[n] = qq.w; out_q[N + n] = qq.x; out_q[2 * N + n] = qq.y; out_q[3 * N + n] = qq.z;
      if (out_R) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out_R[(size_t)(3 * i + j) * N + n] = R.m[i][j];
    };
    if (precision == 0) run(float(0)); else run(double(0));
  }
  return 0;
}

}  // extern "C"
