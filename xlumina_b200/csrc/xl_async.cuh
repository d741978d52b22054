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
// One staging buffer alternates between the two roles, so the kernel keeps two CTAs per SM.
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

// bulk copies are limited by the mbarrier's transaction count (2^20 - 1 bytes per phase); split large tiles
XL_DEV void xl_bulk_tile(void* dst, const void* src, size_t bytes, xl_mbar_t* b) {
    xl_mbar_expect_tx(b, (unsigned)bytes);
    const size_t CH = 32768;
    for (size_t o = 0; o < bytes; o += CH)
        xl_bulk_g2s((char*)dst + o, (const char*)src + o, (unsigned)(bytes - o < CH ? bytes - o : CH), b);
}

// ==================================================================================================================
// K2 (persistent, bulk-asynchronous): column FFT of the row spectra x transfer function -> inverse column FFT.
// Same arithmetic as XlRsCols (wave_optics.py:288); work item = (column pair G, field f), fields of one pair back to back
// so that they share its transfer-function tile in L2.
// ==================================================================================================================
template <int L> struct XlRsColsAsyncOp : XlOpBase {
    static constexpr bool kInLoHalf = true, kOutLoHalf = true;
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    const XlRsParams& p;
    cf* tile;            // this item's column-pair tile in the spectra buffer (in place)
    cf* stage;           // shared staging buffer: input tile [N][2], then transfer-function tile [(L/2+1)][2]
    xl_mbar_t* bar;      // bar[0]: input tile landed, bar[1]: transfer-function tile landed
    unsigned parity, parity_h;   // phase parities of the two barriers for this item
    const cf* Hp;        // global address of the transfer-function pair tile (hmode 0/1), or null
    const cf* H0; const cf* H1; int hmode;   // as in XlRsColsOp (hmode 2: unrelated columns, read through L2)
    const cf* next;      // input tile of this CTA's next item, or null
    XL_DEV void load(int i, cf* v, int stride) const {
        if (i < p.N) xl_ld4(stage + (size_t)i * XL_V, v, v + stride);
        else { v[0] = cf_zero(); v[stride] = cf_zero(); }
    }
    XL_DEV void before_first() const { xl_mbar_wait(bar, parity); }
    // after the barrier that ends the first pass: every thread has consumed the staged input -> the buffer takes H
    XL_DEV void after_first_sync(int tid) const {
        if (tid == 0 && hmode != 2) xl_bulk_tile(stage, Hp, (size_t)(L / 2 + 1) * XL_V * sizeof(cf), bar + 1);
    }
    XL_DEV void before_spec() const { if (hmode != 2) xl_mbar_wait(bar + 1, parity_h); }
    XL_DEV void spec(int beta, cf* v) const {
        const XlHRow<L> hr(beta);
        if (hmode == 2) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const size_t o = (size_t)hr.row(q) * XL_V;
                v[q] = cf_mul(v[q], xl_ldg(H0 + o));
                v[16 + q] = cf_mul(v[16 + q], xl_ldg(H1 + o));
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
    static constexpr int STAGE = (L / 2 + 2) * XL_V;      // cf elements: max(N, L/2 + 1) rows of two columns, 16-byte multiple
    static constexpr bool kPersistent = true;
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
        if (p.stagger_ns && XL_BLOCK_X >= XL_GRID_X / 2) xl_nanosleep(p.stagger_ns);
        unsigned parity = 0, parity_h = 0;
        for (; it < items; it += XL_GRID_X, parity ^= 1) {
            const int G = it / p.nfields;
            const cf* H0 = xl_h_column<L>(p.H, XL_V * G);
            const cf* H1 = xl_h_column<L>(p.H, XL_V * G + 1);
            const bool a0 = (((size_t)(H0 - p.H)) & 1) == 0, a1 = (((size_t)(H1 - p.H)) & 1) == 0;
            const int hmode = (a0 && H1 == H0 + 1) ? 0 : ((a1 && H0 == H1 + 1) ? 1 : 2);
            const int nx = it + XL_GRID_X;
            XlRsColsAsyncOp<L> op{{}, p, item_tile(p, it), stage, bar, parity, parity_h, hmode == 1 ? H1 : H0, H0, H1, hmode,
                                  nx < items ? item_tile(p, nx) : (const cf*)0};
            XlFft<L, XL_V>::conv(s, t, op);
            if (hmode != 2) parity_h ^= 1;
            XL_SYNC();   // the last pass has read the tile: the next item's first pass may overwrite it
        }
    }
};

// ==================================================================================================================
// K1 (persistent, bulk-asynchronous): rows of the zero-padded field -> blocked row spectra (XlRsRowsFwd's arithmetic).
// Work item = (row pair, field), field-major; the two rows of a pair are one contiguous 16*N-byte block of the field, so
// one bulk copy stages them (line-major: stage[l*N + i]) while the previous pair is being transformed.  The Ez items of
// the vectorial path (formed from Ex, Ey at load) come last in the item order and use the direct-load functor.
// ==================================================================================================================
template <int L> struct XlRsRowsFwdAsyncOp : XlOpBase {
    static constexpr bool kInLoHalf = true;
    const XlRsParams& p; int f, yb; const cf* stage; xl_mbar_t* bar; unsigned parity; const cf* next; unsigned next_bytes;
    XL_DEV void before_first() const { xl_mbar_wait(bar, parity); }
    XL_DEV void load(int i, cf* v, int stride) const {
        const bool oki = i < p.N;
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const bool ok = oki && yb + l < p.rows;
            cf x = stage[ok ? (size_t)l * p.N + i : 0];
            if (p.flags & XL_F_CONJ_IN) x = cf_conj(x);
            v[l * stride] = ok ? x : cf_zero();
        }
    }
    XL_DEV void after_first_sync(int tid) const {
        if (tid == 0 && next) xl_bulk_tile(const_cast<cf*>(stage), next, next_bytes, bar);
    }
    XL_DEV void spec(int beta, const cf* v) const {
        cf* base = p.spec + (size_t)f * L * p.rows;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            xl_blocked_store2(base + (size_t)(g / 2) * p.rows * 2, yb, p.rows, g, v[q], v[16 + q]);
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L> struct XlRsRowsFwdAsync {
    static const char* name() { return "rs_rows_fwd"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static constexpr int STAGE = L;                       // cf elements: two rows of N <= L/2 samples
    static constexpr bool kPersistent = true;
    static size_t smem() { return xl_smem_bytes(L, XL_V) + (size_t)STAGE * sizeof(cf) + sizeof(xl_mbar_t); }
    // the bulk path needs 16-byte aligned row pairs of 16-byte multiples: even N (checked by the host, p.chunk_rows carries
    // the number of STAGED items: all of them, or those of the first two fields when field 2 is Ez)
    XL_DEV static const cf* item_src(const Params& p, int it, int groups, unsigned* bytes) {
        const int f = p.f0 + it / groups, yb = (it % groups) * XL_V;
        const int nr = p.rows - yb < XL_V ? p.rows - yb : XL_V;
        *bytes = (unsigned)((size_t)nr * p.N * sizeof(cf));
        return p.in + ((size_t)f * p.rows + yb) * p.N;
    }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* stage = s + xl_tile_elems(L, XL_V);
        cf* t = stage + STAGE;
        xl_mbar_t* bar = (xl_mbar_t*)(t + xl_tw_total(L));
        const int groups = (p.rows + XL_V - 1) / XL_V;
        const int items = groups * p.nfields, staged = p.chunk_rows;
        int it = XL_BLOCK_X;
        if (it >= items) return;
        XL_THREADS(tid, NT) {
            if (tid == 0) { xl_mbar_init(bar, 1); xl_mbar_init_fence(); }
        }
        XlFft<L, XL_V>::init_tw(t, p.tw);
        if (it < staged) {
            XL_THREADS(tid, NT) {
                if (tid == 0) { unsigned b; const cf* src = item_src(p, it, groups, &b); xl_bulk_tile(stage, src, b, bar); }
            }
        }
        const double z = (p.flags & XL_F_VRS) ? xl_ldg(p.z) : 0.0;
        if (p.stagger_ns && XL_BLOCK_X >= XL_GRID_X / 2) xl_nanosleep(p.stagger_ns);
        unsigned parity = 0;
        for (; it < items; it += XL_GRID_X) {
            const int f = p.f0 + it / groups, yb = (it % groups) * XL_V;
            if (it < staged) {
                const int nx = it + XL_GRID_X;
                unsigned nb = 0;
                const cf* next = nx < staged ? item_src(p, nx, groups, &nb) : (const cf*)0;
                XlRsRowsFwdAsyncOp<L> op{{}, p, f, yb, stage, bar, parity, next, nb};
                XlFft<L, XL_V>::forward(s, t, op);
                parity ^= 1;
            } else {   // Ez = (Ex X + Ey Y)/r formed while loading (vectorized_optics.py:258-261): direct loads
                XlRsRowsFwdOp<L, true> op{{}, p, f, yb, z * z};
                XlFft<L, XL_V>::forward(s, t, op);
            }
            XL_SYNC();   // the last pass has read the tile: the next item's first pass may overwrite it
        }
    }
};
