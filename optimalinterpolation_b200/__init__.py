"""B200-native per-grid-cell GP regression (the hot path of GPR_CS2S3.py) behind a C ABI."""
import os as _os

# the lockstep engine overlaps 8+ streams: give them their own hardware queues (must be set before CUDA initialises)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .gpr import GPRDay, Handle, OIError  # noqa: F401,E402
from . import synthetic  # noqa: F401,E402
