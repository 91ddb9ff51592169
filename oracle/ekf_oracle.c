/* ekf_oracle.c -- float64 C restatement of the reference replay path.  TEST INFRASTRUCTURE.
 *
 * Same role as oracle/ekf_oracle.py (the pinned numpy oracle), compiled so that thousands of
 * filters x thousands of steps can be checked in seconds and so that an optimised multi-core CPU
 * rate can be quoted beside the GPU number.  Only tests/, __graft_entry__ and bench.py's CPU-baseline
 * leg load it.  It deliberately does NOT share code with the device header (csrc/ekf_math.cuh): the
 * matrices are dense 4x4 as in the reference, the Wahba rotation comes from a Jacobi SVD of the
 * 3x3 B as the reference forms it, the gain from a pivoted Gauss-Jordan inverse.
 *
 * Parity status: pinned -- tests/test_oracle_c.py checks it against the frozen outputs of the
 * unmodified reference (tests/golden) to 1e-12 rad.
 *
 * Reference lines (PKF = "Python Kalman Filter"):
 *   half_omega .... PKF/ExtendedKalmanFilter.py:27-30,44-47     jacobian_b .. :51-56
 *   rk4 ........... PKF/ExtendedKalmanFilter.py:25-41            predict ..... :58-68
 *   correct ....... PKF/ExtendedKalmanFilter.py:70-80            wahba ....... PKF/Wahba.py:8-17
 *   rot2quat ...... PKF/Wahba.py:20-47                           norm ........ PKF/UtilityFunctions.py:16-21
 *   driver loop ... PKF/main_file.py:19-47
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

static void half_omega(const double w[3], double A[4][4]) {
  const double x = 0.5 * w[0], y = 0.5 * w[1], z = 0.5 * w[2];
  const double M[4][4] = {{0, -x, -y, -z}, {x, 0, z, -y}, {y, -z, 0, x}, {z, y, -x, 0}};
  memcpy(A, M, sizeof(M));
}

static void jacobian_b(const double q[4], double B[4][3]) {
  const double M[4][3] = {{-q[1], -q[2], -q[3]}, {q[0], q[3], -q[2]}, {-q[3], q[0], q[1]}, {q[2], -q[1], q[0]}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) B[i][j] = 0.5 * M[i][j];
}

static double norm4(const double a[4]) {
  double s = 0.0;
  for (int i = 0; i < 4; ++i) s += a[i] * a[i];
  return sqrt(s);
}

static void matvec4(const double A[4][4], const double v[4], double o[4]) {
  for (int i = 0; i < 4; ++i) {
    double s = 0.0;
    for (int j = 0; j < 4; ++j) s += A[i][j] * v[j];
    o[i] = s;
  }
}

static void rk4(const double q0[4], double dt_ns, const double w[3], double out[4]) {
  double W[4][4], k1[4], k2[4], k3[4], k4[4], t[4];
  half_omega(w, W);
  const double h = dt_ns * 1e-9;
  matvec4(W, q0, k1);
  for (int i = 0; i < 4; ++i) t[i] = q0[i] + h / 2 * k1[i];
  matvec4(W, t, k2);
  for (int i = 0; i < 4; ++i) t[i] = q0[i] + h / 2 * k2[i];
  matvec4(W, t, k3);
  for (int i = 0; i < 4; ++i) t[i] = q0[i] + h * k3[i];
  matvec4(W, t, k4);
  for (int i = 0; i < 4; ++i) out[i] = q0[i] + 1.0 / 6 * h * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
  const double n = norm4(out);
  for (int i = 0; i < 4; ++i) out[i] /= n;
}

/* Gauss-Jordan inverse with partial pivoting; returns 0 on success */
static int inverse4(const double S[4][4], double Inv[4][4]) {
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a[i][j] = S[i][j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int p = c;
    for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    if (a[p][c] == 0.0) return 1;
    if (p != c) for (int j = 0; j < 8; ++j) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
    const double d = a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] /= d;
    for (int r = 0; r < 4; ++r) if (r != c) {
      const double f = a[r][c];
      if (f != 0.0) for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) Inv[i][j] = a[i][4 + j];
  return 0;
}

static double det3(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

/* SVD of a 3x3 by one-sided Jacobi run to convergence, singular values sorted descending,
 * U completed to an orthogonal matrix.  B = U diag(s) V^T. */
static void svd3(const double B[3][3], double U[3][3], double s[3], double V[3][3]) {
  double G[3][3];
  memcpy(G, B, sizeof(G));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < 3; ++i) { al += G[i][p] * G[i][p]; be += G[i][q] * G[i][q]; ga += G[i][p] * G[i][q]; }
        if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
        off = fmax(off, fabs(ga) / sqrt(al * be));
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int i = 0; i < 3; ++i) {
          double gp = G[i][p], gq = G[i][q];
          G[i][p] = c * gp - sn * gq; G[i][q] = sn * gp + c * gq;
          double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - sn * vq; V[i][q] = sn * vp + c * vq;
        }
      }
    if (off < 1e-16) break;
  }
  double n[3];
  int idx[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) { n[j] = sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j]); }
  for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (n[idx[b]] > n[idx[a]]) { int t = idx[a]; idx[a] = idx[b]; idx[b] = t; }
  double Vs[3][3];
  for (int j = 0; j < 3; ++j) {
    s[j] = n[idx[j]];
    for (int i = 0; i < 3; ++i) { Vs[i][j] = V[i][idx[j]]; U[i][j] = (s[j] > 0) ? G[i][idx[j]] / s[j] : 0.0; }
  }
  memcpy(V, Vs, sizeof(Vs));
  /* rank(B) = 2 for two observations: complete the third left vector orthogonally */
  if (s[2] <= 1e-13 * s[0]) {
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
}

static void wahba_rotation(const double ra[3], const double rm[3], const double a[3], const double m[3], double ka,
                           double km, double R[3][3]) {
  double B[3][3], U[3][3], V[3][3], s[3], Vt[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[i][j] = ka * ra[i] * a[j] + km * rm[i] * m[j];
  svd3(B, U, s, V);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vt[i][j] = V[j][i];
  const double d = det3(U) * det3(Vt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = U[i][0] * Vt[0][j] + U[i][1] * Vt[1][j] + d * U[i][2] * Vt[2][j];
}

static void rot2quat(const double M[3][3], double q[4]) {
  const double t1 = 1.0 + M[0][0] - M[1][1] - M[2][2];
  const double t2 = 1.0 - M[0][0] + M[1][1] - M[2][2];
  const double t3 = 1.0 - M[0][0] - M[1][1] + M[2][2];
  if (t1 > t2 && t1 > t3) {
    const double S = sqrt(t1) * 2;
    q[0] = (M[2][1] - M[1][2]) / S; q[1] = 0.25 * S; q[2] = (M[0][1] + M[1][0]) / S; q[3] = (M[0][2] + M[2][0]) / S;
  } else if (t2 > t1 && t2 > t3) {
    const double S = sqrt(t2) * 2;
    q[0] = (M[0][2] - M[2][0]) / S; q[1] = (M[0][1] + M[1][0]) / S; q[2] = 0.25 * S; q[3] = (M[1][2] + M[2][1]) / S;
  } else {
    const double S = sqrt(t3) * 2;
    q[0] = (M[1][0] - M[0][1]) / S; q[1] = (M[0][2] + M[2][0]) / S; q[2] = (M[1][2] + M[2][1]) / S; q[3] = 0.25 * S;
  }
}

static void mm4(const double A[4][4], const double B[4][4], double C[4][4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[i][k] * B[k][j];
      C[i][j] = s;
    }
}

/* One filter, T steps.  streams [T][9][N] float32 (column n), dt_ns [T] float64.
 * out_traj: [T][4] doubles with stride traj_stride between steps (or NULL). */
static void replay_one(int64_t T, int64_t N, int64_t n, const float* streams, const double* dt_ns, const float* acc_ref,
                       const float* mag_ref, double qs, double rs, double* out_traj, int64_t traj_stride, double* out_x,
                       double* out_P, uint8_t* out_flip, int64_t flip_stride) {
  double X[4] = {1, 0, 0, 0}, P[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  const double ra[3] = {acc_ref[n], acc_ref[N + n], acc_ref[2 * N + n]};
  const double rm[3] = {mag_ref[n], mag_ref[N + n], mag_ref[2 * N + n]};
  for (int64_t t = 0; t < T; ++t) {
    const float* s = streams + (size_t)t * 9 * N + n;
    const double w[3] = {s[0], s[N], s[2 * N]}, a[3] = {s[3 * N], s[4 * N], s[5 * N]}, m[3] = {s[6 * N], s[7 * N], s[8 * N]};
    /* Prediction */
    double A[4][4], At[4][4], Bn[4][3], AP[4][4], APAt[4][4], Pn[4][4], S[4][4], Si[4][4], K[4][4], z[4];
    half_omega(w, A);
    jacobian_b(X, Bn);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) At[i][j] = A[j][i];
    mm4(A, P, AP);
    mm4(AP, At, APAt);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double bq = 0.0;
        for (int k = 0; k < 3; ++k) bq += (Bn[i][k] * qs) * Bn[j][k];     /* B (q I) B^T */
        Pn[i][j] = APAt[i][j] + bq;
        S[i][j] = Pn[i][j] + ((i == j) ? rs : 0.0);
      }
    rk4(X, dt_ns[t], w, z);
    inverse4(S, Si);
    mm4(Pn, Si, K);
    /* Correction */
    double R[3][3], y[4];
    const double ka = fabs(a[2]);
    wahba_rotation(ra, rm, a, m, ka, 1.0 - ka, R);
    rot2quat(R, y);
    const double cmp = y[0] * z[0] + y[1] * z[1] + y[2] * z[2] + y[3] * z[3];
    const int flip = cmp < 0.0;
    if (flip) for (int i = 0; i < 4; ++i) y[i] = -y[i];
    double e[4], Ke[4], KP[4][4];
    for (int i = 0; i < 4; ++i) e[i] = y[i] - z[i];
    matvec4(K, e, Ke);
    for (int i = 0; i < 4; ++i) X[i] = z[i] + Ke[i];
    mm4(K, Pn, KP);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) P[i][j] = Pn[i][j] - KP[i][j];
    const double nx = norm4(X);
    for (int i = 0; i < 4; ++i) X[i] /= nx;
    if (out_traj) for (int i = 0; i < 4; ++i) out_traj[(size_t)t * traj_stride + i] = X[i];
    if (out_flip) out_flip[(size_t)t * flip_stride] = (uint8_t)flip;
  }
  if (out_x) for (int i = 0; i < 4; ++i) out_x[i] = X[i];
  if (out_P) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out_P[4 * i + j] = P[i][j];
}

/* N filters, split over `n_threads` pthreads (contiguous slices of the filter axis).
 * out_traj [T][N][4] or NULL; out_x [N][4]; out_P [N][16] or NULL; out_flip [T][N] or NULL. */
typedef struct {
  int64_t N, T, n0, n1;
  const float *streams, *acc_ref, *mag_ref;
  const double *dt_ns, *q, *r;
  double *out_traj, *out_x, *out_P;
  uint8_t* out_flip;
} replay_job;

static void* replay_worker(void* arg) {
  replay_job* j = (replay_job*)arg;
  for (int64_t n = j->n0; n < j->n1; ++n)
    replay_one(j->T, j->N, n, j->streams, j->dt_ns, j->acc_ref, j->mag_ref, j->q[n], j->r[n],
               j->out_traj ? j->out_traj + (size_t)n * 4 : 0, j->N * 4, j->out_x ? j->out_x + (size_t)n * 4 : 0,
               j->out_P ? j->out_P + (size_t)n * 16 : 0, j->out_flip ? j->out_flip + n : 0, j->N);
  return 0;
}

int oracle_replay_f64(int64_t N, int64_t T, const float* streams, const double* dt_ns, const float* acc_ref,
                      const float* mag_ref, const double* q, const double* r, double* out_traj, double* out_x,
                      double* out_P, uint8_t* out_flip, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  if ((int64_t)n_threads > N) n_threads = (int)(N > 0 ? N : 1);
  pthread_t th[256];
  replay_job jobs[256];
  for (int k = 0; k < n_threads; ++k) {
    replay_job j = {N, T, N * k / n_threads, N * (k + 1) / n_threads, streams, acc_ref, mag_ref, dt_ns, q, r,
                    out_traj, out_x, out_P, out_flip};
    jobs[k] = j;
    if (n_threads == 1) replay_worker(&jobs[k]);
    else if (pthread_create(&th[k], 0, replay_worker, &jobs[k]) != 0) return 1;
  }
  if (n_threads > 1) for (int k = 0; k < n_threads; ++k) pthread_join(th[k], 0);
  return 0;
}

/* Wahba only: inputs [3][N] float32, weights [N] float64 -> out_R [N][9], out_q [N][4] */
int oracle_wahba_f64(int64_t N, const float* acc_ref, const float* mag_ref, const float* acc, const float* mag,
                     const double* ka, const double* km, double* out_R, double* out_q) {
  for (int64_t n = 0; n < N; ++n) {
    const double ra[3] = {acc_ref[n], acc_ref[N + n], acc_ref[2 * N + n]}, rm[3] = {mag_ref[n], mag_ref[N + n], mag_ref[2 * N + n]};
    const double a[3] = {acc[n], acc[N + n], acc[2 * N + n]}, m[3] = {mag[n], mag[N + n], mag[2 * N + n]};
    double R[3][3], qq[4];
    wahba_rotation(ra, rm, a, m, ka[n], km[n], R);
    rot2quat(R, qq);
    if (out_R) for (int i = 0; i < 9; ++i) out_R[(size_t)n * 9 + i] = R[i / 3][i % 3];
    if (out_q) for (int i = 0; i < 4; ++i) out_q[(size_t)n * 4 + i] = qq[i];
  }
  return 0;
}
