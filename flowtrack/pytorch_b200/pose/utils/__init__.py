"""Reference module layout lib/pose/utils/{evaluation,transforms}.py for the functions on the hot path."""
from . import evaluation, transforms  # noqa: F401
