/* replay_capi.c -- the drop-in boundary used from plain C: no Python, no torch, no CUDA headers.
 *
 * Replays N independent filters over T samples that live in HOST memory through
 * posekf_replay_host_f32 (include/posekf.h) and writes the final states and the trajectory.
 * This is the loop of `Python Kalman Filter/main_file.py:38-47` (Prediction + Correction per sample,
 * Q = 1, R = 0.1 as in main_file.py:21-22) for a whole batch in one call.
 *
 *   gcc -O2 -Iinclude examples/replay_capi.c -o replay_capi -Lposeestimationkf_b200 -lposekf_b200 \
 *       -Wl,-rpath,$PWD/poseestimationkf_b200
 *   ./replay_capi N T in.bin out.bin
 *
 *   in.bin   float32: streams [T][9][N] (gyro xyz, acc xyz, mag xyz; filter index fastest),
 *                     acc_ref [3][N], mag_ref [3][N]
 *   out.bin  float32: final X [4][N], final P [10][N] (upper triangle), trajectory [T][N][4]
 */
#include <stdio.h>
#include <stdlib.h>

#include "posekf.h"

static float* read_floats(FILE* f, size_t n) {
  float* p = (float*)malloc(n * sizeof(float));
  if (!p || fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: %s N T in.bin out.bin\n", argv[0]); return 2; }
  const int64_t N = atoll(argv[1]), T = atoll(argv[2]);
  FILE* fin = fopen(argv[3], "rb");
  if (!fin) { perror(argv[3]); return 2; }
  float* streams = read_floats(fin, (size_t)T * 9 * N);
  float* acc_ref = read_floats(fin, (size_t)3 * N);
  float* mag_ref = read_floats(fin, (size_t)3 * N);
  fclose(fin);

  float* q = (float*)malloc(N * sizeof(float));
  float* r = (float*)malloc(N * sizeof(float));
  for (int64_t n = 0; n < N; ++n) { q[n] = 1.0f; r[n] = 0.1f; }     /* setQ(1), setR(0.1) */
  float* x = (float*)malloc((size_t)4 * N * sizeof(float));
  float* p = (float*)malloc((size_t)10 * N * sizeof(float));
  float* traj = (float*)malloc((size_t)T * N * 4 * sizeof(float));

  /* X0 = [1,0,0,0], P0 = I (main_file.py:23,26): pass NULL.  No low-pass (alpha < 0), default chunking,
   * rank-2 Wahba solver, plain float32 state, device 0, temporary workspace. */
  const int rc = posekf_replay_host_f32(N, T, streams, 0.01f, acc_ref, mag_ref, q, r, -1.0f, -1.0f, NULL, NULL, x, p, traj,
                                        0, POSEKF_WAHBA_QR2, 0, 0, NULL);
  if (rc != 0) { fprintf(stderr, "posekf_replay_host_f32 failed: %d (%s)\n", rc, posekf_version()); return 1; }

  FILE* fout = fopen(argv[4], "wb");
  if (!fout) { perror(argv[4]); return 2; }
  fwrite(x, sizeof(float), (size_t)4 * N, fout);
  fwrite(p, sizeof(float), (size_t)10 * N, fout);
  fwrite(traj, sizeof(float), (size_t)T * N * 4, fout);
  fclose(fout);
  printf("%s: %lld filters x %lld steps; filter 0 ends at [%.6f %.6f %.6f %.6f]\n", posekf_version(), (long long)N, (long long)T,
         x[0], x[N], x[2 * N], x[3 * N]);
  return 0;
}
