"""bdpose — B200-native (sm_100a) implementation of the bin-and-delta pose hot path of
JHUVisionLab/multi-modal-regression.  The flat modules one directory up (axisAngle, quaternion,
binDeltaLosses, binDeltaGenerators, binDeltaModels, poseModels) mirror the reference's import names
and call signatures; this package holds the C-ABI binding and the tensor-level wrappers."""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "kmeans", "metrics"]
