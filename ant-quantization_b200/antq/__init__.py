"""Core of the B200 fake-quant path: ctypes binding to libantq.so + torch-facing ops."""
from . import _lib  # noqa: F401
from .ops import (Codebook, HostPipeline, absmax, calibrate, decode_p4, encode_p4, fakequant, fakequant_dynamic,  # noqa: F401
                  fakequant_backward, fakequant_grouped, fakequant_plan, linear_p4, linear_p4_fp8, lut_nearest, mse_sweep,
                  prepare_codebook)
