"""poseestimationkf_b200 -- B200-native batched quaternion EKF (drop-in for the offline replay path of
varunbachalli/PoseEstimationKF).  See DESIGN.md."""
