# same export as the reference's lib/pose/models/__init__.py:1 (the hot path covers the deconv head only)
from .pose_deconv import DeconvResnet, deconv  # noqa: F401
