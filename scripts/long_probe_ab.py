"""A/B of the inverse radix step of the split kernels: cluster kernels (distributed shared memory) against the two-launch form."""
import sys, os, subprocess
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(here))
from xlumina_b200 import _lib
L = _lib.lib()
for on in (1, 0, -1):
    L.xl_debug_set_long_cluster(on)
    print(f"===== cluster kernels { {1: 'on', 0: 'off', -1: 'auto (R <= 4)'}[on] }", flush=True)
    sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["16384"])
    exec(open(os.path.join(here, "long_probe.py")).read())
