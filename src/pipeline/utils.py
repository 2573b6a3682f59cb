"""Drop-in for reference src/pipeline/utils.py: Proposals / depthmap_to_pointcloud / get_z_from_pointcloud /
mask_to_bbox from the B200 package; other names (point-cloud helpers of the scale stage) fall through to the
reference module when $FREEPOSE_REFERENCE_ROOT is set."""
import importlib.util as _ilu
import os as _os

from freepose_b200.pipeline.proposals import Proposals  # noqa: F401
from freepose_b200.pipeline.utils import (depthmap_to_pointcloud, get_z_from_pointcloud,  # noqa: F401
                                          mask_to_bbox)

_ref_mod = None


def __getattr__(name):
    global _ref_mod
    ref = _os.environ.get("FREEPOSE_REFERENCE_ROOT")
    path = _os.path.join(ref, "src", "pipeline", "utils.py") if ref else None
    if path and _os.path.exists(path):
        if _ref_mod is None:
            spec = _ilu.spec_from_file_location("_freepose_ref_pipeline_utils", path)
            _ref_mod = _ilu.module_from_spec(spec)
            spec.loader.exec_module(_ref_mod)
        return getattr(_ref_mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r} (set FREEPOSE_REFERENCE_ROOT for the rest "
                         "of the reference module)")
