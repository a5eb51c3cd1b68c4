from . import models, utils  # noqa: F401
