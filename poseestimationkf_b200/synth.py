"""Synthetic IMU sequence generator.

The recorded dataset of the reference is not in its repository (`Python Kalman Filter/ReadFile.py:24`
opens an un-shipped file), so every parity and throughput input is synthetic.  The generator is
plain torch (it runs on the CPU for the tests and on the GPU for the benchmark) and is *plumbing*:
it produces the `[T, 9, N]` float32 stream tensor the replay kernel consumes, it is never timed.

Conventions (validated against the reference by tests/test_synth.py): scalar-first quaternion
[w,x,y,z]; q maps body -> initial frame; omega is the body rate, q_dot = 0.5*q (x) (0,omega);
measurements are `meas = R(q)^T ref`, so that Wahba's `ref ~= R meas` recovers R(q).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

GYRO, ACC, MAG = slice(0, 3), slice(3, 6), slice(6, 9)
N_CHANNELS = 9


@dataclass
class SyntheticIMU:
    """streams  : [T, 9, N] float32   rows 0-2 gyro (rad/s, body), 3-5 acc, 6-8 mag (unit vectors, body)
    acc_ref  : [3, N] float32      sample 0 of acc  (the log's acc_0)
    mag_ref  : [3, N] float32      sample 0 of mag  (the log's mag_0)
    q_true   : [T, 4, N] float64 or None   ground-truth attitude after each step
    dt       : float               seconds between samples"""
    streams: torch.Tensor
    acc_ref: torch.Tensor
    mag_ref: torch.Tensor
    q_true: torch.Tensor | None
    dt: float


def _quat_mul(a, b):
    """Hamilton product of [4,N] tensors (scalar first)."""
    aw, ax, ay, az = a[0], a[1], a[2], a[3]
    bw, bx, by, bz = b[0], b[1], b[2], b[3]
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw))


def _rotate_inverse(q, v):
    """R(q)^T v for q [4,N], v [3,N]."""
    w, x, y, z = q[0], q[1], q[2], q[3]
    vx, vy, vz = v[0], v[1], v[2]
    # rows of R(q)^T are the columns of R(q)
    r00 = 1 - 2 * (y * y + z * z); r01 = 2 * (x * y - w * z); r02 = 2 * (x * z + w * y)
    r10 = 2 * (x * y + w * z); r11 = 1 - 2 * (x * x + z * z); r12 = 2 * (y * z - w * x)
    r20 = 2 * (x * z - w * y); r21 = 2 * (y * z + w * x); r22 = 1 - 2 * (x * x + y * y)
    return torch.stack((r00 * vx + r10 * vy + r20 * vz,
                        r01 * vx + r11 * vy + r21 * vz,
                        r02 * vx + r12 * vy + r22 * vz))


def _unit(v):
    return v / torch.linalg.vector_norm(v, dim=0, keepdim=True)


def make_imu(n_filters: int, n_steps: int, *, seed: int = 0, sigma: float = 0.0, dt: float = 0.01,
             device="cpu", keep_truth: bool = False, out: torch.Tensor | None = None,
             az_range=(0.02, 0.98)) -> SyntheticIMU:
    """Generate `n_filters` independent trajectories of `n_steps` samples.

    Body rate: omega_i(t) = sum_{j<3} A_ij sin(2 pi f_ij t + phi_ij), A~U(0.2,1) rad/s,
    f~U(0.05,0.5) Hz, phi~U(0,2pi).  True attitude: per-sample exact exponential of the mid-point
    rate.  Reference vectors: gravity (0,0,1) and field normalize(0.4,0,-0.9165) seen through a
    random initial tilt drawn so that |acc_z| starts inside `az_range`.  Noise sigma is added to
    all three sensors, acc/mag are renormalised afterwards.  Everything is computed in float64 and
    rounded to float32 once; the oracle consumes exactly those float32 values.
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=dev)
    N, T = n_filters, n_steps

    def U(lo, hi, *shape):
        return lo + (hi - lo) * torch.rand(*shape, generator=g, **f64)

    amp, freq, phase = U(0.2, 1.0, 3, 3, N), U(0.05, 0.5, 3, 3, N), U(0.0, 2 * math.pi, 3, 3, N)

    # initial tilt: gravity in the body frame has |z| = cos(tilt) in az_range, random azimuth
    az = U(az_range[0] + 0.03, az_range[1] - 0.03, N) * torch.where(U(0, 1, N) < 0.5, -1.0, 1.0)
    psi = U(0.0, 2 * math.pi, N)
    s = torch.sqrt(1 - az * az)
    acc0 = torch.stack((s * torch.cos(psi), s * torch.sin(psi), az))
    # field: fixed angle to gravity (world: g=(0,0,1), m=normalize(0.4,0,-0.9165)); pick the
    # horizontal direction at a random heading around gravity
    mw = torch.tensor([0.4, 0.0, -0.9165], **f64)
    mw = mw / torch.linalg.vector_norm(mw)
    hdg = U(0.0, 2 * math.pi, N)
    helper = torch.where((acc0[2].abs() < 0.9).unsqueeze(0),
                         torch.tensor([0.0, 0.0, 1.0], **f64).unsqueeze(1).expand(3, N),
                         torch.tensor([1.0, 0.0, 0.0], **f64).unsqueeze(1).expand(3, N))
    e1 = _unit(torch.linalg.cross(acc0, helper, dim=0))
    e2 = torch.linalg.cross(acc0, e1, dim=0)
    horiz = torch.cos(hdg) * e1 + torch.sin(hdg) * e2
    mag0 = _unit(mw[2] * acc0 + mw[0] * horiz)

    if out is None:
        out = torch.empty((T, N_CHANNELS, N), dtype=torch.float32, device=dev)
    assert out.shape == (T, N_CHANNELS, N) and out.dtype == torch.float32
    q_true = torch.empty((T, 4, N), **f64) if keep_truth else None

    def add_noise(v):
        if sigma == 0.0:
            return v
        return v + sigma * torch.randn(v.shape, generator=g, **f64)

    if sigma != 0.0:   # the references themselves are noisy samples, like a real log's sample 0
        acc0_meas, mag0_meas = _unit(add_noise(acc0)), _unit(add_noise(mag0))
    else:
        acc0_meas, mag0_meas = acc0, mag0

    q = torch.zeros((4, N), **f64)
    q[0] = 1.0
    two_pi_f = 2 * math.pi * freq
    for i in range(T):
        t_mid = (i + 0.5) * dt
        t_end = (i + 1.0) * dt
        w_mid = (amp * torch.sin(two_pi_f * t_mid + phase)).sum(dim=1)        # [3,N]
        w_end = (amp * torch.sin(two_pi_f * t_end + phase)).sum(dim=1)
        ang = torch.linalg.vector_norm(w_mid, dim=0) * dt
        half = 0.5 * ang
        sinc = torch.where(ang > 1e-12, torch.sin(half) / ang.clamp_min(1e-300), torch.full_like(ang, 0.5))
        dq = torch.cat((torch.cos(half).unsqueeze(0), w_mid * dt * sinc), dim=0)
        q = _quat_mul(q, dq)
        q = q / torch.linalg.vector_norm(q, dim=0, keepdim=True)
        # the gyro sample logged with step i is the body rate at that sample's timestamp
        out[i, GYRO] = add_noise(w_end).to(torch.float32)
        out[i, ACC] = _unit(add_noise(_rotate_inverse(q, acc0))).to(torch.float32)
        out[i, MAG] = _unit(add_noise(_rotate_inverse(q, mag0))).to(torch.float32)
        if keep_truth:
            q_true[i] = q
    return SyntheticIMU(out, acc0_meas.to(torch.float32), mag0_meas.to(torch.float32), q_true, dt)
