// xl_api.cu -- the whole library as ONE translation unit.  The product build compiles xl_core.cu, xl_rs.cu, xl_slab.cu,
// xl_czt.cu and xl_elements.cu separately (xlumina_b200/build.py: one nvcc process per kernel family, reproducible and parallel); this file
// is what the test-only host emulation (g++ -DXL_HOST_EMU, tests/emu) and single-command builds compile.
#include "xl_core.cu"
#include "xl_rs.cu"
#include "xl_slab.cu"
#include "xl_czt.cu"
#include "xl_elements.cu"
