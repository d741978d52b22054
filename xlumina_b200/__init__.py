"""
xlumina_b200 -- B200-native (sm_100a) light propagation for XLuminA's hot path.

Mirrors the reference's public surface for that path only (paths relative to the XLuminA repo):
  xlumina/__init__.py:15-21            units um, nm, mm, cm (base unit = micron)
  xlumina/wave_optics.py               ScalarLight, LightSource, RS_propagation / CZT (+ the *_jit seam functions)
  xlumina/vectorized_optics.py         VectorizedLight, PolarizedLightSource, VRS_propagation / VCZT
  xlumina/optical_elements.py:515-672  high_NA_objective_lens is fused into VCZT_objective_lens
  xlumina/toolbox.py:49-61             space

Arrays are torch CUDA tensors; every propagation runs in libxlprop.so (hand-written CUDA behind the C ABI in
include/xlprop.h).  There is no CPU fallback.
"""
import math

um = 1
nm = 1e-3
mm = 1e3
cm = 1e4

radians = 1
degrees = 180 / math.pi

from .toolbox import space  # noqa: E402
from .wave_optics import ScalarLight, LightSource  # noqa: E402
from .vectorized_optics import VectorizedLight, PolarizedLightSource  # noqa: E402
from .optical_elements import VCZT_objective_lens  # noqa: E402
from . import ops  # noqa: E402

__all__ = ["um", "nm", "mm", "cm", "radians", "degrees", "space", "ScalarLight", "LightSource", "VectorizedLight",
           "PolarizedLightSource", "VCZT_objective_lens", "ops"]
