"""curvepointslam_b200 -- B200-native hot path of CurvePointSLAM (batched Bezier LM fits, EKF
measurement update, Mahalanobis gate) behind a C ABI (include/*.h, libcps.so)."""
