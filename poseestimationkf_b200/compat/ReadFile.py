"""`ReadFile.getData` with the reference's attribute names (reference: Python Kalman Filter/
ReadFile.py:1-45).  The reference opens a hard-coded Windows path (:24); here the path comes from
the constructor argument or the POSEKF_LOG environment variable (default ./KalmanFilter.txt)."""
import os

from poseestimationkf_b200 import logio


class getData(logio.LogData):
    def __init__(self, path=None):
        super().__init__()
        self.path = path or os.environ.get("POSEKF_LOG", "KalmanFilter.txt")
        self.readFile()

    @staticmethod
    def getArray(line, n):                       # reference :14-21
        return logio._values(line)

    def readFile(self):                          # reference :23-45
        parsed = logio.read_log(self.path)
        for name in ("mag_0", "mag_1", "acc_0", "acc_1", "gyro", "timestamp", "quart_wahba", "quart_xk", "quart_gyro"):
            setattr(self, name, getattr(parsed, name))
