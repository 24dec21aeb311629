"""B200-native per-grid-cell GP regression (the hot path of GPR_CS2S3.py) behind a C ABI."""
from .gpr import GPRDay, Handle, OIError  # noqa: F401
from . import synthetic  # noqa: F401
