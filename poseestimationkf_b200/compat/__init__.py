"""Drop-in modules with the reference's import names.

Put this directory on sys.path and the reference's replay script imports resolve to the B200
implementation:

    sys.path.insert(0, poseestimationkf_b200.compat.PATH)
    from ExtendedKalmanFilter import KalmanFilter      # Python Kalman Filter/main_file.py:6
    from Wahba import Wahba                            # :2
    from UtilityFunctions import DimensionalSplit, norm  # :4
"""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
