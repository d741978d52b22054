// xl_async.cuh -- bulk-asynchronous tile movement (cp.async.bulk + mbarrier, the 1-D form of TMA) and the persistent
// column kernels built on it.
//
// The column kernels of the RS path read one CONTIGUOUS tile per work item (the blocked layouts of xl_kernels.cuh were
// chosen for that: a column pair of the row spectra is 16*N bytes, a column pair of the transfer function (L/2+1)*16
// bytes).  Instead of LDG -> registers -> STS by all threads in the first pass (memory phase and butterfly phase of a CTA
// serialise), ONE thread issues a bulk copy into a staging buffer in shared memory and an mbarrier counts the bytes:
//   * the CTAs are persistent (grid = resident CTAs per SM x SM count) and walk the work items;
//   * the input tile of item i+1 is requested as soon as the spectrum phase of item i has finished with the staging
//     buffer, and lands while item i runs its inverse passes and stores;
//   * the transfer-function tile of item i is requested right after the first pass has consumed the input tile, and lands
//     while the middle forward pass runs: the spectrum multiply reads shared memory instead of waiting on L2.
// One staging buffer alternates between the two roles, so the kernels keep two CTAs per SM.  The d/dz column kernel works
// the same way on single columns: its input tile is the interleaved (cotangent, conj-field) column written by rs_rows_dual,
// its spectrum factors are the contiguous column copies of H and of the reduced dH/dz (xl_h_colcopy).
// SASS: UBLKCP (bulk copy) + SYNCS (mbarrier arrive.expect_tx / try_wait).
// Host emulation (tests/emu): a bulk copy is a memcpy at issue time and waiting is a no-op; the ordering on the device is
// argued next to each barrier below and checked by compute-sanitizer racecheck on the device (profiles/).
#pragma once
#include "xl_kernels.cuh"
#ifdef XL_HOST_EMU
#include <string.h>
#endif

typedef unsigned long long xl_mbar_t;

#ifdef XL_HOST_EMU
static inline void xl_mbar_init(xl_mbar_t* b, int) { *b = 0; }
static inline void xl_mbar_init_fence() {}
static inline void xl_mbar_expect_tx(xl_mbar_t*, unsigned) {}
static inline void xl_bulk_g2s(void* dst, const void* src, unsigned bytes, xl_mbar_t*) { memcpy(dst, src, bytes); }
static inline void xl_mbar_wait(xl_mbar_t*, unsigned) {}
#else
XL_DEV unsigned xl_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
XL_DEV void xl_mbar_init(xl_mbar_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xl_saddr(b)), "r"(count) : "memory");
}
// makes the initialised barrier visible to the asynchronous proxy (followed by a CTA barrier)
XL_DEV void xl_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival + the number of bytes the copies issued next will deliver
XL_DEV void xl_mbar_expect_tx(xl_mbar_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xl_saddr(b)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (16-byte aligned on both sides, bytes a multiple of 16); completes on the mbarrier
XL_DEV void xl_bulk_g2s(void* dst, const void* src, unsigned bytes, xl_mbar_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(xl_saddr(dst)), "l"(src), "r"(bytes), "r"(xl_saddr(b)) : "memory");
}
// wait for the phase with the given parity to complete (hardware-suspended try_wait, not a spin on shared memory)
XL_DEV void xl_mbar_wait(xl_mbar_t* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "XL_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra XL_MBAR_DONE;\n"
        "bra XL_MBAR_WAIT;\n"
        "XL_MBAR_DONE:\n"
        "}\n" ::"r"(xl_saddr(b)), "r"(parity) : "memory");
}
#endif

// one tile as bulk copies of at most 32 KB; the caller has announced the bytes (xl_mbar_expect_tx)
XL_DEV void xl_bulk_chunks(void* dst, const void* src, size_t bytes, xl_mbar_t* b) {
    const size_t CH = 32768;
    for (size_t o = 0; o < bytes; o += CH)
        xl_bulk_g2s((char*)dst + o, (const char*)src + o, (unsigned)(bytes - o < CH ? bytes - o : CH), b);
}
// announce + copy one tile (an mbarrier phase counts at most 2^20 - 1 bytes: far above any tile here)
XL_DEV void xl_bulk_tile(void* dst, const void* src, size_t bytes, xl_mbar_t* b) {
    xl_mbar_expect_tx(b, (unsigned)bytes);
    xl_bulk_chunks(dst, src, bytes, b);
}

// ==================================================================================================================
// K2 (persistent, bulk-asynchronous): column FFT of the row spectra x transfer function -> inverse column FFT, keep rows
// [0,N).  Replaces wave_optics.py:288.  Work item = (column pair G, field f), fields of one pair back to back so that they
// share its transfer-function tile in L2.  Transfer function of a pair: the 16-byte interleaved pair tile when both columns
// live in one stored pair (hmode 0 / 1 = straight / swapped), else the two contiguous column copies (hmode 2).
// ==================================================================================================================
template <int L> struct XlRsColsAsyncOp : XlOpBase {
    static constexpr bool kInLoHalf = true, kOutLoHalf = true;
    static constexpr bool kSpecSyncCta = true;   // after_spec_sync() overwrites the staging buffer every thread has just read
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    static constexpr int HCR = xl_hc_rows(L);
    const XlRsParams& p;
    cf* tile;            // this item's column-pair tile in the spectra buffer (in place)
    cf* stage;           // shared staging buffer: input tile [N][2], then the transfer function of the pair
    xl_mbar_t* bar;      // bar[0]: input tile landed, bar[1]: transfer function landed
    unsigned parity;     // phase parity of both barriers for this item
    const cf* Ha; const cf* Hb; int hmode;   // hmode 0/1: Ha = pair tile; hmode 2: Ha, Hb = the two column copies
    const cf* next;      // input tile of this CTA's next item, or null
    XL_DEV void load(int i, cf* v, int stride) const {
        if (i < p.N) xl_ld4(stage + (size_t)i * XL_V, v, v + stride);
        else { v[0] = cf_zero(); v[stride] = cf_zero(); }
    }
    XL_DEV void before_first() const { xl_mbar_wait(bar, parity); }
    // after the barrier that ends the first pass: every thread has consumed the staged input -> the buffer takes H
    XL_DEV void after_first_sync(int tid) const {
        if (tid != 0) return;
        if (hmode != 2) {
            xl_mbar_expect_tx(bar + 1, (unsigned)((L / 2 + 1) * XL_V * sizeof(cf)));
            xl_bulk_chunks(stage, Ha, (size_t)(L / 2 + 1) * XL_V * sizeof(cf), bar + 1);
        } else {
            xl_mbar_expect_tx(bar + 1, (unsigned)(2 * HCR * sizeof(cf)));
            xl_bulk_chunks(stage, Ha, (size_t)HCR * sizeof(cf), bar + 1);
            xl_bulk_chunks(stage + HCR, Hb, (size_t)HCR * sizeof(cf), bar + 1);
        }
    }
    XL_DEV void before_spec() const { xl_mbar_wait(bar + 1, parity); }
    XL_DEV void spec(int beta, cf* v) const {
        const XlHRow<L> hr(beta);
        if (hmode == 2) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int r = hr.row(q);
                v[q] = cf_mul(v[q], stage[r]);
                v[16 + q] = cf_mul(v[16 + q], stage[HCR + r]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                cf lo, hi;
                xl_ld4(stage + (size_t)hr.row(q) * XL_V, &lo, &hi);
                v[q] = cf_mul(v[q], hmode == 0 ? lo : hi);
                v[16 + q] = cf_mul(v[16 + q], hmode == 0 ? hi : lo);
            }
        }
    }
    // after the barrier that ends the spectrum phase: the buffer is free again -> request the next item's input
    XL_DEV void after_spec_sync(int tid) const {
        if (tid == 0 && next) xl_bulk_tile(stage, next, (size_t)p.N * XL_V * sizeof(cf), bar);
    }
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1 / 2; ++j) {
            const int i = n + S1 * j;
            if (i < p.N) xl_st4(tile + (size_t)i * XL_V, v[j], v[R1 + j]);
        }
    }
};
template <int L> struct XlRsColsAsync {
    static const char* name() { return "rs_cols"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static constexpr int STAGE = 2 * xl_hc_rows(L);       // cf elements: max(2N, 2(L/2+1), 2 column copies), 16-byte multiple
    static size_t smem() { return xl_smem_bytes(L, XL_V) + (size_t)STAGE * sizeof(cf) + 2 * sizeof(xl_mbar_t); }
    XL_DEV static cf* item_tile(const Params& p, int it) {
        const int G = it / p.nfields, f = p.f0 + it % p.nfields;
        return p.spec + (size_t)f * L * p.N + (size_t)G * p.N * XL_V;
    }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* stage = s + xl_tile_elems(L, XL_V);           // 16-byte aligned: the tile holds an even number of cf
        cf* t = stage + STAGE;
        xl_mbar_t* bar = (xl_mbar_t*)(t + xl_tw_total(L));
        const int items = (L / XL_V) * p.nfields;
        int it = XL_BLOCK_X;
        if (it >= items) return;
        XL_THREADS(tid, NT) {
            if (tid == 0) {
                xl_mbar_init(bar, 1);
                xl_mbar_init(bar + 1, 1);
                xl_mbar_init_fence();
            }
        }
        XlFft<L, XL_V>::init_tw(t, p.tw);                 // ends with a CTA barrier: the mbarriers are initialised for all
        XL_THREADS(tid, NT) {
            if (tid == 0) xl_bulk_tile(stage, item_tile(p, it), (size_t)p.N * XL_V * sizeof(cf), bar);
        }
        unsigned parity = 0;
        for (; it < items; it += XL_GRID_X, parity ^= 1) {
            const int G = it / p.nfields;
            const cf* H0 = xl_h_column<L>(p.H, XL_V * G);
            const cf* H1 = xl_h_column<L>(p.H, XL_V * G + 1);
            const bool a0 = (((size_t)(H0 - p.H)) & 1) == 0, a1 = (((size_t)(H1 - p.H)) & 1) == 0;
            const int hmode = (a0 && H1 == H0 + 1) ? 0 : ((a1 && H0 == H1 + 1) ? 1 : 2);
            const cf* Ha = hmode == 0 ? H0 : (hmode == 1 ? H1 : xl_h_colcopy<L>(p.H, XL_V * G));
            const cf* Hb = xl_h_colcopy<L>(p.H, XL_V * G + 1);
            const int nx = it + XL_GRID_X;
            XlRsColsAsyncOp<L> op{{}, p, item_tile(p, it), stage, bar, parity, Ha, Hb, hmode,
                                  nx < items ? item_tile(p, nx) : (const cf*)0};
            XlFft<L, XL_V>::conv(s, t, op);
            XL_SYNC();   // the last pass has read the tile: the next item's first pass may overwrite it
        }
    }
};

// ==================================================================================================================
// K4 (persistent, bulk-asynchronous factors): backward column kernel with d/dz.  Work item = (x-slot g, field f).  The two
// lines of the forward transform are the cotangent spectra column C and the column W of the spectra of conj(U) -- both in
// one 16-byte element per row of the interleaved tile written by rs_rows_dual -- so both column spectra meet in the
// registers of the same thread:
//     gz += Re sum conj(W) * C * Hz'        (Parseval form of ct_z, SURVEY.md A.1, with the REDUCED kernel Hz' of xl_rs_h)
// C*H goes through a one-line inverse FFT (in place on line 0 of the tile) for ct_field.  Three FFTs per column, nothing
// parked in HBM.  The column copies of H and Hz' are bulk-staged at the start of the item (they land during the first two
// passes); the (C,W) tile of the NEXT item is pulled into L2 by one bulk prefetch.
// ==================================================================================================================
template <int L> struct XlRsColsGzAsyncOp : XlOpBase {
    static constexpr bool kInLoHalf = true;
    static constexpr int HCR = xl_hc_rows(L);
    const XlRsParams& p; const cf* in; const cf* stage; xl_mbar_t* bar; unsigned parity; cf* tile2; float* red;
    XL_DEV void load(int i, cf* v, int stride) const {
        xl_ldg4(in + (size_t)(i < p.N ? i : 0) * 4, v, v + stride);      // {cotangent, conj-field} of row i
        if (i >= p.N) { v[0] = cf_zero(); v[stride] = cf_zero(); }
    }
    XL_DEV void before_spec() const { xl_mbar_wait(bar, parity); }
    XL_DEV void spec(int beta, const cf* v) const {
        float acc = 0.f;
        cf u[16];
        const XlHRow<L> hr(beta);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int r = hr.row(q);
            const cf t = cf_mul(v[q], stage[HCR + r]);
            acc += v[16 + q].x * t.x + v[16 + q].y * t.y;  // Re(conj(w) * t)
            u[q] = cf_mul(v[q], stage[r]);
        }
        red[beta] += acc;
        if (p.flags & XL_F_NOFIELD) return;         // d/dz only (kernel-uniform): no field cotangent, no inverse transform
        XlBfly<16, +1, false, false>::run(u);       // first inverse pass, fused; back into line 0 of the slots just read
#pragma unroll
        for (int j = 0; j < 16; ++j) XlTileLine0Of2::st(tile2, 16 * beta + j, u + j, 16);
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L> struct XlRsColsGzAsync {
    static const char* name() { return "rs_cols_gz"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static constexpr int NB = L / 16;                     // butterflies per line == entries of the partial-sum array
    static constexpr int HCR = xl_hc_rows(L);
    static constexpr int STAGE = 2 * HCR;
    static size_t smem() {
        return (size_t)(xl_tile_elems(L, 2) + STAGE + xl_tw_total(L)) * sizeof(cf) + 2 * sizeof(xl_mbar_t) + (size_t)NB * sizeof(float) + NT * sizeof(double);
    }
    // p.spec2: interleaved (C,W) spectra [f][L/2][N][4] (rs_rows_dual); p.spec: result, blocked pair layout [f][L/2][N][2]
    XL_DEV static const cf* item_in(const Params& p, int it) {
        const int g = it / p.nfields, f = p.f0 + it % p.nfields;
        return p.spec2 + ((size_t)f * (L / 2) + (g >> 1)) * p.N * 4 + (g & 1) * 2;
    }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* stage = s + xl_tile_elems(L, 2);
        cf* t = stage + STAGE;
        xl_mbar_t* bar = (xl_mbar_t*)(t + xl_tw_total(L));
        double* dred = (double*)(bar + 2);
        float* red = (float*)(dred + NT);
        const int items = L * p.nfields;
        int it = XL_BLOCK_X;
        if (it >= items) return;
        XL_THREADS(tid, NT) {
            if (tid == 0) { xl_mbar_init(bar, 1); xl_mbar_init_fence(); }
            for (int i = tid; i < NB; i += NT) red[i] = 0.f;
        }
        XlFft<L, 2>::init_tw(t, p.tw);
        unsigned parity = 0;
        for (; it < items; it += XL_GRID_X, parity ^= 1) {
            const int g = it / p.nfields, f = p.f0 + it % p.nfields, nx = it + XL_GRID_X;
            XL_THREADS(tid, NT) {
                if (tid == 0) {   // the staging buffer is free (barrier at the end of the previous item)
                    xl_mbar_expect_tx(bar, (unsigned)(2 * HCR * sizeof(cf)));
                    xl_bulk_chunks(stage, xl_h_colcopy<L>(p.H, g), (size_t)HCR * sizeof(cf), bar);
                    xl_bulk_chunks(stage + HCR, xl_h_colcopy<L>(p.H2, g), (size_t)HCR * sizeof(cf), bar);
                    if (nx < items && !((nx / p.nfields) & 1))      // the even column of a pair fetches the pair's tile
                        for (size_t o = 0; o < (size_t)p.N * 4 * sizeof(cf); o += 32768) {
                            const size_t left = (size_t)p.N * 4 * sizeof(cf) - o;
                            xl_prefetch_l2_bulk((const char*)item_in(p, nx) + o, (unsigned)(left < 32768 ? left : 32768));
                        }
                }
            }
            XlRsColsGzAsyncOp<L> op{{}, p, item_in(p, it), stage, bar, parity, s, red};
            XlFft<L, 2>::forward(s, t, op);
            if (!(p.flags & XL_F_NOFIELD)) {
                // spectrum phase done.  The first inverse pass went back into the slots this thread had read, and the B = 256
                // level that follows stays inside the half-warp's block of 256 positions: a warp-level barrier is enough
                // (the staging buffer is not reused before the CTA barrier that ends the item)
                XlFft<L, 2>::template sync_local<XlFft<L, 2>::kHasMid>();
                XlRsColsGzOutOp<L> oo{{}, p, p.spec + (size_t)f * L * p.N + (size_t)(g >> 1) * p.N * XL_V, g & 1};
                XlFft<L, 1, XlTileLine0Of2>::inverse_tail(s, t, oo);
            }
            XL_SYNC();                                   // the last pass has read the tile; the staging buffer is free
        }
        // one reduction and one atomic per CTA
        XL_THREADS(tid, NT) {
            double a = 0.0;
            for (int i = tid; i < NB; i += NT) a += (double)red[i];
            dred[tid] = a;
        }
        XL_SYNC();
        xl_block_sum<NT>(dred);
        XL_THREADS(tid, NT) {
            if (tid == 0) xl_atomic_add(p.gz, dred[0]);
        }
    }
};
