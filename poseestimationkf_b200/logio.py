"""Reader / writer of the reference's text log (`KalmanFilter.txt`), SURVEY.md section 8(f-1).

The C++ server appends one `key : v,v,v` line per event (`Kalman Filter Server/PoseEstimator/
KalmanFilter.cpp:28,32,60-67,138,150-153,180-183,274,287,300`, numbers via `std::to_string`, i.e.
`%f`), and the Python replay parses it with substring tests in a fixed order
(`Python Kalman Filter/ReadFile.py:23-45`).  `read_log` follows those parse rules exactly (the order
matters: `'q_gyro'` is tested before `'gyro'`, and `'T'` is a bare substring test); `write_log`
emits the server's line sequence.  This is host-side text I/O, as in the reference; the arrays it
yields feed the device through `LogData.to_streams`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class LogData:
    """Same attribute names as the reference's `getData` (ReadFile.py:3-11)."""
    mag_0: list = field(default_factory=list)
    mag_1: list = field(default_factory=list)
    acc_0: list = field(default_factory=list)
    acc_1: list = field(default_factory=list)
    gyro: list = field(default_factory=list)
    timestamp: list = field(default_factory=list)
    quart_wahba: list = field(default_factory=list)
    quart_xk: list = field(default_factory=list)
    quart_gyro: list = field(default_factory=list)

    # ---- device hand-off ---------------------------------------------------------------------
    def n_steps(self) -> int:
        return len(self.acc_1)

    def to_streams(self, device="cuda"):
        """-> (streams [T,9,1] float32, acc_ref [3,1], mag_ref [3,1], dt [T] float32 seconds), the
        arguments of `batched.replay`.  Timestamps are differenced in integer/float64 ns on the host
        (the first one only seeds previousT: main_file.py:19,25)."""
        import torch
        T = self.n_steps()
        s = np.concatenate([np.asarray(self.gyro[:T], dtype=np.float64), np.asarray(self.acc_1, dtype=np.float64),
                            np.asarray(self.mag_1[:T], dtype=np.float64)], axis=1)           # [T, 9]
        t_ns = np.asarray(self.timestamp, dtype=np.float64)[:, 0]
        dt = np.diff(t_ns)[:T] * 1e-9
        dev = torch.device(device)
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        return (f32(s[:, :, None]), f32(np.asarray(self.acc_0, dtype=np.float64)[:, None]),
                f32(np.asarray(self.mag_0, dtype=np.float64)[:, None]), f32(dt))


def _values(line: str):
    """ReadFile.py:14-21 (`getArray`): text after the first ':' split on ','."""
    return [float(v) for v in line.split(":")[1].split(",")]


def parse_lines(lines) -> LogData:
    d = LogData()
    for line in lines:                                   # ReadFile.py:27-45, same test order
        if "mag_0" in line:
            d.mag_0 = _values(line)
        elif "acc_0" in line:
            d.acc_0 = _values(line)
        elif "Acc_1" in line:
            d.acc_1.append(_values(line))
        elif "Mag_1" in line:
            d.mag_1.append(_values(line))
        elif "q_gyro" in line:
            d.quart_gyro.append(_values(line))
        elif "gyro" in line:
            d.gyro.append(_values(line))
        elif "T" in line:
            d.timestamp.append(_values(line))
        elif "Wahba_quart" in line:
            d.quart_wahba.append(_values(line))
        elif "X_k" in line:
            d.quart_xk.append(_values(line))
    return d


def read_log(path: str) -> LogData:
    with open(path) as fh:
        return parse_lines(fh.readlines())


def _f(v) -> str:
    return "%f" % float(v)          # std::to_string(double)


def _vec(key: str, v) -> str:
    return key + " : " + ",".join(_f(x) for x in v)


def format_log(acc_0, mag_0, t0_ns: int, t_ns, gyro, mag_1, acc_1, x_k, wahba_quart, q_gyro) -> list[str]:
    """The server's line sequence: header (`set_mag_0`, `set_acc_0`, `compute_initial_params`:
    KalmanFilter.cpp:26-33,56-67), the first `T` line (`Prediction` :136-141), then per sample
    gyro (:274), T, q_gyro (:150-153), Mag_1 (:287), Acc_1 (:300), X_k, Wahba_quart (:180-183)."""
    out = [_vec("mag_0", mag_0), _vec("acc_0", acc_0), "q_gyro : 1.0, 0.0, 0.0, 0.0", "X_k : 1.0, 0.0, 0.0, 0.0",
           "Wahba_quart : 1.0, 0.0, 0.0, 0.0"]
    for i in range(len(gyro)):
        out.append(_vec("gyro", gyro[i]))
        if i == 0:
            out.append("T : %d" % int(t0_ns))
        out.append("T : %d" % int(t_ns[i]))
        out.append(_vec("q_gyro", q_gyro[i]))
        out.append(_vec("Mag_1", mag_1[i]))
        out.append(_vec("Acc_1", acc_1[i]))
        out.append(_vec("X_k", x_k[i]))
        out.append(_vec("Wahba_quart", wahba_quart[i]))
    return out


def write_log(path: str, **kw) -> None:
    with open(path, "w") as fh:
        fh.write("\n".join(format_log(**kw)) + "\n")
