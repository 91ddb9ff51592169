// Developer probe: instantiate ONE replay kernel so that a source experiment compiles in seconds.
#include "../../poseestimationkf_b200/csrc/ekf_math.cuh"
#include "../../poseestimationkf_b200/csrc/device_util.cuh"
using namespace pkf;
#include "../../poseestimationkf_b200/csrc/replay_kernels.cuh"
#ifndef ONE_COMP
#define ONE_COMP false
#endif
#ifndef ONE_AUX
#define ONE_AUX false
#endif
namespace pkf_dev {
template __global__ void replay_tma2_kernel<WAHBA_QR2, false, ONE_AUX, ONE_COMP>(const ReplayParams, const __grid_constant__ CUtensorMap);
}
