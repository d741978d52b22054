"""A/B of the inverse radix step of the split-line chain on one GPU: cluster kernels (distributed shared memory) against the
two-launch form through the scratch buffer, and the default (clusters for split factors <= 4).
usage: python scripts/long_probe_ab.py [N ...]"""
import sys, os
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(here))
from xlumina_b200 import _lib
L = _lib.lib()
for cluster in (1, 0, -1):
    L.xl_debug_set_long_cluster(cluster)
    print(f"===== cluster kernels { {1: 'on', 0: 'off', -1: 'auto (R <= 4)'}[cluster] }", flush=True)
    sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["16384"])
    exec(open(os.path.join(here, "long_probe.py")).read())
