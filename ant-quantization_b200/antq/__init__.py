"""Core of the B200 fake-quant path: ctypes binding to libantq.so + torch-facing ops."""
from . import _lib  # noqa: F401
from .ops import (Codebook, HostPipeline, absmax, fakequant, fakequant_grouped, fakequant_plan,  # noqa: F401
                  lut_nearest, mse_sweep, prepare_codebook)
