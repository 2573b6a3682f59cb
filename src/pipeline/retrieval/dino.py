"""Drop-in for reference src/pipeline/retrieval/dino.py (B200 engine)."""
from freepose_b200.pipeline.retrieval.dino import DINOv2FeatureExtractor  # noqa: F401
