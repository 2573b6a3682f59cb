"""Drop-in for reference src/pipeline/estimators/online_pose_estimator.py."""
from freepose_b200.pipeline.estimators.online_pose_estimator import DinoOnlinePoseEstimator  # noqa: F401
