// ref_harness.cpp -- C entry points around the reference's OWN C++ for the two "next" rows of SURVEY.md section 8f
// that the Python reference does not cover (TEST INFRASTRUCTURE, built into oracle/_ref/ by oracle/Makefile):
//   * InitialValues.cpp (Kalman Filter Server/PoseEstimator) is compiled from where it lies, unmodified;
//   * Parser::NormalizeValues / Parser::LinearInterpolationSensor are included from parser_extract.inc, which
//     extract_parser.py writes from the reference's Parser.cpp at build time.
// Used to freeze tests/golden/preprocess_ref.npz and initial_values_ref.npz (tests/golden/make_golden_cpp.py).
#include <array>
#include <cmath>
#include <cstdio>
#include <unistd.h>
#include <fcntl.h>

#include "InitialValues.h"          // the reference's header (-I to its directory)
#include "parser_extract.inc"       // generated: ref_parser_NormalizeValues, ref_parser_LinearInterpolationSensor

extern "C" {

// InitialValues(K), K calls of setValuesforAverage, then getAverageValues / getVariance.      (InitialValues.cpp:4-66)
// The class prints its result with printf: stdout is pointed at /dev/null for the duration of the call.
int ref_initial_values(const double* samples_xyz, int K, double* avg3, double* var3) {
  fflush(stdout);
  const int saved = dup(1), devnull = open("/dev/null", O_WRONLY);
  dup2(devnull, 1);
  InitialValues iv(K);
  for (int i = 0; i < K; ++i) iv.setValuesforAverage(samples_xyz[3 * i], samples_xyz[3 * i + 1], samples_xyz[3 * i + 2]);
  fflush(stdout);
  dup2(saved, 1);
  close(saved); close(devnull);
  if (!iv.sensorCalibrated()) return 1;
  const std::array<double, 3> a = iv.getAverageValues(), v = iv.getVariance();
  for (int k = 0; k < 3; ++k) { avg3[k] = a[k]; var3[k] = v[k]; }
  return 0;
}

// Parser::NormalizeValues in place.                                                              (Parser.cpp:221-228)
void ref_normalize(double* v3) {
  std::array<double, 3> a = {v3[0], v3[1], v3[2]};
  ref_parser_NormalizeValues(a);
  v3[0] = a[0]; v3[1] = a[1]; v3[2] = a[2];
}

// Parser::LinearInterpolationSensor(t1, t2, t3, y1, y2).                                         (Parser.cpp:259-267)
void ref_interpolate(long long t1, long long t2, long long t3, const double* y1, const double* y2, double* out3) {
  const std::array<double, 3> a = {y1[0], y1[1], y1[2]}, b = {y2[0], y2[1], y2[2]};
  const std::array<double, 3> r = ref_parser_LinearInterpolationSensor(t1, t2, t3, a, b);
  out3[0] = r[0]; out3[1] = r[1]; out3[2] = r[2];
}

// ExecuteKalmanFilter's sequence for one sensor: interpolate to the gyro timestamp, then normalise.  (Parser.cpp:232-242)
void ref_interpolate_normalise(long long t1, long long t2, long long t3, const double* y1, const double* y2, double* out3) {
  ref_interpolate(t1, t2, t3, y1, y2, out3);
  ref_normalize(out3);
}

}  // extern "C"
