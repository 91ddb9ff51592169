"""Builds tests/hostsim/libhostsim.so (g++, CPU) -- a test-only float32/float64 build of the device
math header.  Not used by the product."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhostsim.so")
SRC = os.path.join(HERE, "hostsim.cpp")
HDR = os.path.join(HERE, "..", "..", "poseestimationkf_b200", "csrc", "ekf_math.cuh")


def build(force=False):
    if (not force and os.path.exists(SO)
            and os.path.getmtime(SO) > max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        return SO
    # -ffp-contract=off: only the explicit fma_ calls fuse, like the nvcc build of the same header
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-ffp-contract=off", "-fPIC", "-shared",
                           "-o", SO, SRC])
    return SO


if __name__ == "__main__":
    print(build(force=True))
