"""Drop-in for reference src/utils/bbox_utils.py: CropResizePad runs on the GPU; the small box helpers are
re-exported from the reference module when a reference checkout is configured."""
import importlib.util as _ilu
import os as _os

from freepose_b200.pipeline.bbox_utils import CropResizePad  # noqa: F401

_ref = _os.environ.get("FREEPOSE_REFERENCE_ROOT")
if _ref and _os.path.exists(_os.path.join(_ref, "src", "utils", "bbox_utils.py")):
    _spec = _ilu.spec_from_file_location("_freepose_ref_bbox_utils", _os.path.join(_ref, "src", "utils", "bbox_utils.py"))
    _mod = _ilu.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    for _k, _v in vars(_mod).items():
        if not _k.startswith("_") and _k != "CropResizePad":
            globals().setdefault(_k, _v)
