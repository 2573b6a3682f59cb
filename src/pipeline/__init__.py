"""Overlay package: modules present here (the B200 hot path) shadow the reference's; everything else resolves from
the reference checkout named by $FREEPOSE_REFERENCE_ROOT."""
import os as _os

_ref = _os.environ.get("FREEPOSE_REFERENCE_ROOT")
if _ref:
    _p = _os.path.join(_ref, *__name__.split("."))
    if _os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)
