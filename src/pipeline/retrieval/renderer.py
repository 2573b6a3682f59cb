"""Drop-in for reference src/pipeline/retrieval/renderer.py (CUDA rasteriser)."""
from freepose_b200.pipeline.retrieval.renderer import MeshRenderer  # noqa: F401
