"""Drop-in for reference src/pipeline/estimators/pose_estimator.py."""
from freepose_b200.pipeline.estimators.pose_estimator import DinoPoseEstimator  # noqa: F401
