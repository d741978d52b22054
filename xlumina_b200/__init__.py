"""
xlumina_b200 -- B200-native (sm_100a) light propagation for XLuminA's hot path.

Mirrors the reference's public surface for that path only (paths relative to the XLuminA repo):
  xlumina/__init__.py:15-21            units um, nm, mm, cm (base unit = micron)
  xlumina/wave_optics.py               ScalarLight, LightSource, RS_propagation / CZT (+ the *_jit seam functions)
  xlumina/vectorized_optics.py         VectorizedLight, PolarizedLightSource, VRS_propagation / VCZT
  xlumina/optical_elements.py:515-672  high_NA_objective_lens is fused into VCZT_objective_lens
  xlumina/toolbox.py:49-61             space
and, as the first "next" rows of SURVEY.md 8f (callers either side of the path, torch arithmetic on the field planes):
  xlumina/optical_elements.py:76-392, 678-703, 894-939, 1503-1649   SLM, sSLM, LCD, linear_polarizer, BS_symmetric, lens,
                                       building_block, hybrid_setup_sharp_focus   (optical_elements.py)
  xlumina/loss_functions.py            small_area_hybrid, MSE_* ...               (loss_functions.py)
  experiments/four_f_optical_table.py  the dual-SLM 4f table and its batch loss   (four_f.py)

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
from .optical_elements import (VCZT_objective_lens, SLM, sSLM, sSLM_with_amplitude, LCD, linear_polarizer, BS_symmetric,  # noqa: E402
                               lens, cylindrical_lens, axicon_lens, building_block, bb_amplitude_and_phase_mod,
                               hybrid_setup_sharp_focus)
from . import ops, loss_functions, four_f  # noqa: E402

__all__ = ["um", "nm", "mm", "cm", "radians", "degrees", "space", "ScalarLight", "LightSource", "VectorizedLight",
           "PolarizedLightSource", "VCZT_objective_lens", "SLM", "sSLM", "sSLM_with_amplitude", "LCD", "linear_polarizer",
           "BS_symmetric", "lens", "cylindrical_lens", "axicon_lens", "building_block", "bb_amplitude_and_phase_mod", "hybrid_setup_sharp_focus", "ops",
           "loss_functions", "four_f"]
