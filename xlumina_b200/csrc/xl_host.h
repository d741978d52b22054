// xl_host.h -- host helpers that need the kernels' parameter structs (included after xl_kernels.cuh).
#pragma once
#include "xl_common.h"
#include "xl_kernels.cuh"

static inline int rs_base_params(XlRsParams& p, int N, double dx, double dy, double k) {
    memset(&p, 0, sizeof(p));
    p.N = N;
    p.L = xl_rs_padded_length(N);
    if (!p.L) return xl_fail(XL_E_UNSUPPORTED, "RS: N=%s%lld unsupported (padded length must be in [32,4096])", "", N);
    p.rows = N; p.chunk_rows = N;
    p.dx = dx; p.dy = dy; p.k = k;
    p.hscale = (float)(dx * dy / ((double)p.L * (double)p.L));
    p.tw = xl_twiddles();
    if (!p.tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    return XL_OK;
}
