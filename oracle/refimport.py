"""ORACLE support (test infrastructure): import the reference's OWN modules from /root/reference, unmodified,
with ``sys.modules`` stubs for dependencies that are not installed (sam2/hydra, skimage, pyrender, trimesh,
loguru) and ``torch.hub.load`` patched to return a hub-shaped oracle ViT.  Only used in this build container
to validate the restatements and to mint tests/golden fixtures (tests/golden/make_golden.py);
/root/reference does not exist on the GPU box, so nothing at run time may depend on this module."""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "src" / "pipeline" / "estimators" / "pose_estimator.py").exists()


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def install_stubs():
    _stub("sam2"); _stub("sam2.utils")
    _stub("sam2.utils.amg", mask_to_rle_pytorch=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError()))
    _stub("skimage")
    _stub("skimage.measure", regionprops=None)
    _stub("skimage.morphology", isotropic_erosion=None)
    try:
        import loguru  # noqa: F401
    except ImportError:
        import logging
        _stub("loguru", logger=logging.getLogger("reference"))
    rf = types.SimpleNamespace(SKIP_CULL_FACES=1024)
    _stub("pyrender", IntrinsicsCamera=None, OffscreenRenderer=None, Mesh=None, Scene=None)
    _stub("pyrender.constants", RenderFlags=rf)
    _stub("trimesh", Trimesh=type("Trimesh", (), {}), PointCloud=type("PointCloud", (), {}))
    _stub("open3d")


class _RefPath:
    """Puts /root/reference first on sys.path while importing ``src.*`` and hides this repo's own ``src``
    shim package, restoring everything afterwards."""

    def __enter__(self):
        self._saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in self._saved:
            del sys.modules[k]
        # The reference's ``src`` has no __init__.py (a namespace package), so ANY regular ``src`` package on sys.path
        # -- this repository's overlay -- would win regardless of order: take those entries off the path meanwhile.
        self._path = list(sys.path)
        import os
        sys.path[:] = [str(REFERENCE_ROOT)] + [p for p in sys.path
                                               if not os.path.exists(os.path.join(p or os.getcwd(), "src", "__init__.py"))]
        importlib.invalidate_caches()
        return self

    def __exit__(self, *exc):
        sys.path[:] = self._path
        self.loaded = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in self.loaded:
            del sys.modules[k]
        sys.modules.update(self._saved)
        return False


def import_reference(module: str):
    """e.g. import_reference('src.utils.bbox_utils') -> the reference's module object."""
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?)")
    install_stubs()
    with _RefPath():
        mod = importlib.import_module(module)
    if not str(getattr(mod, "__file__", "")).startswith(str(REFERENCE_ROOT)):
        raise RuntimeError(f"{module} resolved to {mod.__file__}, not to the reference checkout")
    return mod


def reference_feature_extractor(hub_model):
    """The reference's DINOv2FeatureExtractor with torch.hub.load returning `hub_model` (an oracle ViT)."""
    import torch
    ref = import_reference("src.pipeline.retrieval.dino")
    orig = torch.hub.load
    torch.hub.load = lambda *a, **k: hub_model
    try:
        fe = ref.DINOv2FeatureExtractor()
    finally:
        torch.hub.load = orig
    return fe
