"""yolo-compression-and-deployment-in-fpga_b200 — B200-native fixed-point slim_yolo_v2 forward pass.

Host-side mirror of the reference's interface for its one hot path:
  export ... checkpoint -> int8 weights + exponent tables (+ weight.h reader/writer)
  lib ...... ctypes binding of the C-ABI in include/yolo_b200.h (libyolo_b200.so, hand-written CUDA)
  model .... SlimYOLOv2_quantize_bnfuse drop-in (models/slim_yolo_v2.py:40-382)
  runner ... batch sharding over the GPUs of one box + detection gather

The directory name contains '-' (it is the reference's name), so import it through the `yolo_b200`
alias module at the repository root, or with importlib.import_module().
"""
from . import export  # noqa: F401  (pure host logic; importable without a GPU or the built library)

__all__ = ["export"]
