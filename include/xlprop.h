/*
 * xlprop.h -- C ABI of libxlprop.so: B200-native (sm_100a) light propagation for XLuminA's hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces one of the reference's jitted seam functions
 * (paths relative to the XLuminA repository):
 *
 *   xl_rs_fwd / xl_rs_bwd          RS_propagation_jit            xlumina/wave_optics.py:281-289  (+ build_grid :265-279,
 *                                                                transfer_function_RS :291-297) and its JAX VJP
 *   xl_vrs_fwd / xl_vrs_bwd        VRS_propagation_jit           xlumina/vectorized_optics.py:364-373 (+ Ez, :258-261)
 *   xl_czt_fwd / xl_czt_bwd        CZT_jit / VCZT_jit            xlumina/wave_optics.py:333-357, vectorized_optics.py:375-384
 *                                  (+ build_CZT_grid :299-331, Bluestein_method :412-460, compute_fft :385-410)
 *   xl_highna_fwd / xl_highna_bwd  _high_NA_objective_lens_ + vectorized_CZT_for_high_NA + cte
 *                                  xlumina/optical_elements.py:515-638, vectorized_optics.py:386-394, wave_optics.py:359-371
 *
 * Conventions
 *   - All array arguments are DEVICE pointers; complex data is interleaved float32 (re,im) = complex64, row-major
 *     [..., y, x] exactly like the reference's arrays; geometry scalars are float64.
 *   - `z` (propagation distance) is a traced value in the reference (wave_optics.py:281 marks only nx,ny,dx,dy,k static),
 *     so it is a pointer to ONE float64 in device memory.  Everything that is static in the reference is passed by value.
 *   - Coordinate grids (Xext/Yext/X/Y/Xout/Yout) are never passed: they are regenerated analytically from
 *     (first coordinate, spacing); the input and output grids must be uniformly spaced (true for toolbox.space()).
 *   - Backward functions compute the JAX-convention VJP (transpose, no conjugation).  XL_CONJ_IN / XL_CONJ_OUT conjugate
 *     the cotangent on load / the result on store, which turns the call into torch's convention at no cost.
 *   - Nothing here allocates, frees, or synchronises in steady state: scratch is a caller-provided workspace sized by the
 *     matching *_workspace_bytes query, work is enqueued on `stream` (a cudaStream_t) only, and the calls are
 *     CUDA-graph-capturable after the first (warm-up) call on a device, which uploads the twiddle table.
 *   - Return value: 0 on success, a negative XL_E_* code otherwise; xl_last_error() gives a thread-local message.
 *     No call aborts or throws across the ABI.  NaNs are produced only where the reference produces them
 *     (high-NA lens with a sample at rho == 0, optical_elements.py:538).
 */
#ifndef XLPROP_H
#define XLPROP_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XLPROP_VERSION 200

enum {
    XL_OK = 0,
    XL_E_BAD_ARG = -1,      /* null pointer / non-positive size */
    XL_E_UNSUPPORTED = -2,  /* padded length outside [32, 4096], or m+M-1 a power of two (reference raises too) */
    XL_E_WORKSPACE = -3,    /* workspace too small */
    XL_E_CUDA = -4          /* CUDA runtime error (message in xl_last_error) */
};

enum {
    XL_CONJ_IN = 1,   /* conjugate the (co)tangent operand while loading it */
    XL_CONJ_OUT = 2,  /* conjugate results while storing them */
    XL_REUSE_H = 16,  /* forward only: `H` already holds the transfer function for this z (cache hit) */
    XL_PHASE_BLIND = 64,   /* xl_rs_bwd_fused: the caller guarantees that the loss does not change when the output of THIS
                              propagation is multiplied by a global phase (every table whose paths through this plane end
                              in intensity detectors).  Then Im sum ct_out*out is exactly zero and the i k out part of d/dz
                              is dropped instead of being evaluated as a complex64 cancellation residue. */
    XL_WITH_HZ = 128, /* RS / VRS, forward AND backward call of one propagation: `H` is twice xl_rs_transfer_bytes(N) and also
                         holds the reduced dH/dz of the same distance.  The forward call generates both together (the
                         impulse response and its derivative share r, exp(i k r) and the powers of 1/r: one launch pair
                         instead of two), the backward call with grad_z takes dH/dz from there instead of generating it.
                         For callers that know at forward time that z needs a gradient. */
    XL_REUSE_TABLES = 32   /* CZT family: `tables` already holds the tables of these sizes, grids and z (e.g. the backward
                              call of a propagation whose forward call filled them) */
};

int xl_version(void);
const char* xl_last_error(void);

/* Padded FFT length used for an N-sample RS axis: smallest power of two >= 2N-1 (the reference pads to 2N-1,
 * wave_optics.py:286; any length >= 2N-1 gives the identical linear convolution). Returns 0 if unsupported. */
int xl_rs_padded_length(int N);
/* Bluestein length for m inputs -> M outputs: 2^ceil(log2(m+M-1)), wave_optics.py:373-383,440-441. */
int xl_czt_padded_length(int m, int M);

/* ---------------------------------------------------------------- RS / VRS ---------------------------------- */
/* Bytes of one transfer-function buffer H (L*L complex64, private layout). */
size_t xl_rs_transfer_bytes(int N);
size_t xl_rs_workspace_bytes(int N, int nfields, int want_grad_z);

/* H <- FFT2 of the sampled Rayleigh-Sommerfeld impulse response (deriv=0) or of dh/dz (deriv=1), times dx*dy/L^2.
 * Replaces transfer_function_RS + fft2(H), wave_optics.py:285,288,291-297. */
int xl_rs_transfer(void* H, const double* z, int N, double dx, double dy, double k, int deriv, void* stream);

/* `count` transfer functions in ONE launch pair (the fixed per-step cost of an optimizer whose distances are all known at the
 * start of a step): buffer i lives at H + i*h_stride_bytes (>= xl_rs_transfer_bytes(N), a multiple of 16).  pairs == 0:
 * buffer i = H(z[i]) (deriv as in xl_rs_transfer).  pairs == 1 (count even): buffer 2j = H(z[j]), buffer 2j+1 = the reduced
 * dH/dz of z[j] that xl_rs_bwd_fused takes through xl_rs_fuse.Hz. */
int xl_rs_transfer_multi(void* H, size_t h_stride_bytes, const double* z, int count, int pairs, int N,
                         double dx, double dy, double k, int deriv, void* stream);

/* out[f] = (ifft2(fft2(pad(in[f])) * fft2(H)) * dx*dy)[N-1:, N-1:]  for f < nfields.      wave_optics.py:286-288
 * H is written unless XL_REUSE_H is set (keep it for the backward call). */
int xl_rs_fwd(const void* in, void* out, void* H, const double* z, int N, int nfields,
              double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream);

/* VJP of xl_rs_fwd.  ct_in[f] = A^T ct_out[f] (A is complex-symmetric, so this is the forward operator);
 * if grad_z != NULL:  *grad_z += Re sum_f sum ct_out[f] * d out[f]/dz, which needs the primal `in` AND the primal result
 * `out` of the forward call: d out/dz = i k out + (reduced kernel) * in, and the first term -- which cancels identically
 * for intensity-type losses and would otherwise drown the rest in complex64 rounding -- is evaluated exactly in real
 * space as -k Im sum ct_out*out. */
int xl_rs_bwd(const void* in, const void* out, const void* ct_out, void* ct_in, double* grad_z, const void* H,
              const double* z, int N, int nfields, double dx, double dy, double k, int flags,
              void* ws, size_t ws_bytes, void* stream);

/* (Ex, Ey) (N,N) each -> out = [Ex', Ey', Ez'] (3,N,N); Ez = (Ex X + Ey Y)/sqrt(X^2+Y^2+z^2) is formed while loading
 * (vectorized_optics.py:258-261); x0,y0 = first grid coordinates.  Replaces VRS_propagation_jit (:364-373).
 * `ex` and `ey` are the two input planes wherever they live (the elements of an optical table produce them separately, so no
 * stacking copy is needed); ey == NULL means ey = ex + N*N (a stacked (2,N,N) pair).  The same convention holds for every
 * vectorial entry point below.  ct_exy is one (2,N,N) buffer = the cotangents of Ex and Ey. */
int xl_vrs_fwd(const void* ex, const void* ey, void* out, void* H, const double* z, int N, double x0, double y0,
               double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream);
int xl_vrs_bwd(const void* ex, const void* ey, const void* out, const void* ct_out, void* ct_exy, double* grad_z, const void* H,
               const double* z, int N, double x0, double y0, double dx, double dy, double k, int flags,
               void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- fused pointwise elements (scalar RS) ----- */
/* The elements that bracket a scalar propagation in an optical table, folded into its first and last pass (SURVEY.md 8f-1,
 * 8f-2) so that they cost no extra pass over HBM:
 *   mod      shared complex64 [N][N] plane multiplied into EVERY field while it is loaded: a phase-only SLM exp(i phi)
 *            (xlumina/optical_elements.py:87-103), an amplitude mask, or the beam under a batch of object masks
 *            (experiments/four_f_optical_table.py:58-65); NULL: none
 *   in_real  the input planes are float32 (binary object masks), not complex64
 *   target   detection: float32 [nfields][N][N] target intensities; the last pass also accumulates
 *            mse[f] += sum (|out[f]|^2 - target[f])^2 / N^2   (MSE of intensities, four_f_optical_table.py:129-141);
 *            the caller zeroes `mse` (nfields float64); NULL: no detection
 * `out` (complex) is always written: the backward pass needs it. */
typedef struct xl_rs_fuse {
    const void* mod;
    int in_real;
    const float* target;
    double* mse;
    const void* Hz;   /* backward only: the reduced dH/dz transfer function of this z, generated ahead (xl_rs_transfer_multi);
                         NULL: the backward call generates it */
} xl_rs_fuse;
int xl_rs_fwd_fused(const void* in, void* out, void* H, const double* z, int N, int nfields,
                    double dx, double dy, double k, int flags, const xl_rs_fuse* fuse,
                    void* ws, size_t ws_bytes, void* stream);
/* VJP of xl_rs_fwd_fused (conventions of xl_rs_bwd).  Output cotangent: `ct_out` (complex planes), or -- when fuse->target
 * is set -- ct_mse[f] = dL/dmse[f] (device float64): the cotangent 4 ct_mse[f]/N^2 (|out|^2 - target) out is then formed from
 * the primal `out` while it is loaded, and the i k out part of d/dz is dropped because it vanishes identically.
 *   ct_in   [nfields][N][N] cotangent of the (unmodulated) inputs = (A^T ct) * mod;  NULL: the inputs are constants
 *   ct_mod  [N][N] cotangent of `mod` = sum_f (A^T ct)[f] * in[f] (overwritten);     NULL: `mod` is a constant
 *   grad_z  += d/dz (needs `in` and `out`; see XL_PHASE_BLIND);                       NULL: z is a constant
 * With ct_in == ct_mod == NULL only the d/dz pipeline runs (no inverse transforms). */
int xl_rs_bwd_fused(const void* in, const void* out, const void* ct_out, const double* ct_mse, void* ct_in, void* ct_mod,
                    double* grad_z, const void* H, const double* z, int N, int nfields,
                    double dx, double dy, double k, int flags, const xl_rs_fuse* fuse,
                    void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- slab-decomposed RS (multi-GPU) ------------ */
/* One N x N field split into row slabs over G ranks (N a multiple of 2G, L/2 a multiple of G, L = xl_rs_padded_length(N)):
 * rank g owns field rows [g N/G, (g+1) N/G).  These are the per-rank STAGES; the caller moves data between them with an
 * all-to-all (xlumina_b200/slab.py uses torch.distributed: NCCL over NVLink on GPUs).  Forward:
 *     xl_slab_h_rows -> all-to-all -> xl_slab_h_cols                       (transfer-function slab, once per z)
 *     xl_slab_rows_fwd -> all-to-all -> xl_slab_cols -> all-to-all -> xl_slab_rows_inv
 * Buffers: a spectra buffer on the row side is [L/2 slot pairs][rows of this rank][2] complex64; after an all-to-all that
 * sends peer r the slot pairs [r P, (r+1) P), P = (L/2)/G, it is [source rank][P][rows of that rank][2] on the column side,
 * and the second all-to-all is the exact inverse.  The VJP with respect to the field is the same chain applied to the
 * cotangent (the operator is complex-symmetric).  d/dz (JAX differentiates wave_optics.py:285-297 with respect to z):
 *     d out/dz = i k out + (in conv h_red),   h_red = dh/dz - i k h  (the reduced kernel, see xl_rs_bwd)
 * -- xl_slab_h_rows_dz -> all-to-all -> xl_slab_h_cols give the transfer-function slab of h_red, the same forward chain
 * applies it to the primal input, and the two dot products with the cotangent are the caller's (xlumina_b200/slab.py:
 * rs_slab_grad_z, the first one in fp64 from the saved output, summed over the ranks with one all-reduce).
 * Replaces the same reference lines as xl_rs_fwd (wave_optics.py:281-297). */
int xl_slab_padded_length(int N);            /* 2^ceil(log2(2N-1)) up to 32768 (N <= 16384); 0 if unsupported */
int xl_slab_h_rows_per_rank(int N, int G);   /* rows of the y >= 0 half of the impulse response each rank transforms */
size_t xl_slab_scratch_bytes(int N, int G);  /* scratch of the split kernels (padded length > 4096); small otherwise */
int xl_slab_h_rows(void* R, const double* z, int N, int G, int rank, double dx, double dy, double k, void* scratch, void* stream);
int xl_slab_h_rows_dz(void* R, const double* z, int N, int G, int rank, double dx, double dy, double k, void* scratch, void* stream);
int xl_slab_h_cols(const void* Th, void* Hloc, int N, int G, double dx, double dy, void* stream);
int xl_slab_rows_fwd(const void* in_local, void* S, int N, int G, int flags, void* stream);
int xl_slab_cols(void* T, const void* Hloc, int N, int G, void* scratch, void* stream);
int xl_slab_rows_inv(const void* S, void* out_local, int N, int G, int flags, void* scratch, void* stream);
/* Test hook: sub-line length of the split kernels (32 or 4096, default 4096), so that tests reach them at small N. */
void xl_debug_set_max_line(int sub_line_length);
/* Test / A-B hook: inverse radix step of the split kernels inside a thread-block cluster (1) or as a second launch through
 * the scratch buffer (0); -1 (default): clusters when the split factor is <= 4 (padded length <= 16384). */
void xl_debug_set_long_cluster(int on);

/* ---------------------------------------------------------------- CZT / VCZT -------------------------------- */
/* vectorial = 0: in (N,N) -> out (My,Mx); ey ignored         CZT_jit,  wave_optics.py:333-357
 * vectorial = 1: in = Ex, ey = Ey (NULL: in + N*N) -> out (3,My,Mx); Ez = ((Ex X + Ey Y)/r) z/r   VCZT, vectorized_optics.py:341-344,375-384
 * Input grid: x_j = x0 + j dx, y_i = y0 + i dy.  Output grid: Mx samples from xout0 to xoutl, My from yout0 to youtl.
 * Dm = lambda*z/dx (wave_optics.py:322).
 * `tables` is a caller-owned buffer of xl_czt_tables_bytes() bytes (opaque, like H of the RS path): the Bluestein chirps and
 * kernel spectra of both axes (compute_fft, wave_optics.py:385-410) and the RS factors F, F0 (:340-341) sampled on the
 * input and output grids.  A call fills it unless XL_REUSE_TABLES is set; pass the buffer of the forward call with
 * XL_REUSE_TABLES to the backward call (same sizes, grids, z).  `ws` is scratch (xl_czt_workspace_bytes()). */
size_t xl_czt_workspace_bytes(int N, int Mx, int My, int vectorial);
size_t xl_czt_tables_bytes(int N, int Mx, int My);
int xl_czt_fwd(const void* in, const void* ey, void* out, const double* z, double lambda, int N, int Mx, int My, int vectorial,
               double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
               int flags, void* tables, void* ws, size_t ws_bytes, void* stream);
/* VJP with respect to the input field(s) (z, lambda and the grids are static in every reference caller; xl_czt_bwd_z below
 * also returns the gradient with respect to z). */
int xl_czt_bwd(const void* ct_out, void* ct_in, const double* z, double lambda, int N, int Mx, int My, int vectorial,
               double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
               int flags, void* tables, void* ws, size_t ws_bytes, void* stream);

/* Field VJP AND the gradient with respect to the distance: JAX differentiates CZT_jit / VCZT_jit through F, F0,
 * the constant z dx dy lambda and, via Dm = lambda z/dx, every Bluestein chirp (wave_optics.py:322, 340-355, 393-403).
 * `in` / `out` are the primal input and result of the forward call; *grad_z += Re sum ct_out * d out/dz.  The workspace is
 * larger (xl_czt_workspace_bytes_z): two more forward chains run on index-weighted inputs (the chirp-z kernel of an axis is
 * exp(i Phi(l,k)) with dPhi/dDm bilinear in the output and input index). */
size_t xl_czt_workspace_bytes_z(int N, int Mx, int My, int vectorial);
int xl_czt_bwd_z(const void* in, const void* ey, const void* out, const void* ct_out, void* ct_in, double* grad_z,
                 const double* z, double lambda, int N, int Mx, int My, int vectorial,
                 double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                 int flags, void* tables, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- high-NA objective ------------------------- */
/* (Ex, Ey) (ey == NULL: ex + N*N) -> out = [Ex,Ey,Ez] (3,My,Mx) in the focal plane:
 *   -i sin^2(theta_max)/(f lambda) * Bluestein_x(Bluestein_y( apod*G*RL(theta,phi) (Ex,Ey,Ez)^T )),  Dm = f lambda (N-1)/(2R).
 * optical_elements.py:515-672.  `tables` (xl_highna_tables_bytes()) holds the Bluestein tables and the lens matrix
 * apod*G*RL sampled on the input grid; it depends on the sizes, grids, radius, f and lambda only, so one buffer serves
 * every call of an optical table that uses the same objective (XL_REUSE_TABLES). */
size_t xl_highna_workspace_bytes(int N, int Mx, int My);
size_t xl_highna_tables_bytes(int N, int Mx, int My);
int xl_highna_fwd(const void* ex, const void* ey, void* out, int N, int Mx, int My, double radius, double f, double lambda,
                  double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                  int flags, void* tables, void* ws, size_t ws_bytes, void* stream);
int xl_highna_bwd(const void* ct_out, void* ct_exy, int N, int Mx, int My, double radius, double f, double lambda,
                  double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                  int flags, void* tables, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- batched entry points --------------------- */
/* The reference vmaps its seam functions over masks, candidate set-ups and noisy distances (experiments/four_f_optical_table.py:98,
 * examples/noisy_optimization.ipynb cell 7): `nbatch` independent propagations in ONE library call.
 *   - outputs, cotangents and gradients gain a contiguous leading batch axis; vectorial inputs keep the (ex, ey) convention of
 *     the single-item calls with `in_bstride` ELEMENTS between the Ex planes (and between the Ey planes) of consecutive items
 *     (2*N*N for a stacked (B,2,N,N) array with ey == NULL, N*N for two (B,N,N) arrays);
 *   - z[b * z_stride] is item b's distance.  z_stride == 0: ONE distance -- one transfer function / one set of tables, and for the
 *     scalar operators single launches over all the planes of the batch; grad_z then receives the sum over the batch.
 *     z_stride != 0: one distance per item -- `H` holds nbatch buffers xl_rs_transfer_bytes(N) apart (2 * that with XL_WITH_HZ),
 *     all generated in ONE launch pair; `tables` holds nbatch buffers xl_czt_tables_bytes() apart; grad_z[b * gz_stride] per item.
 *   - `ws`: xl_rs_workspace_bytes(N, nfields * nbatch, .) for the scalar RS calls with z_stride == 0, xl_czt_workspace_bytes_batch()
 *     for the CZT calls, otherwise the single-item size (the items reuse it in stream order).
 * Flags as in the single-item calls. */
int xl_rs_fwd_batch(const void* in, void* out, void* H, const double* z, int z_stride, int N, int nfields, int nbatch,
                    double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream);
int xl_rs_bwd_batch(const void* in, const void* out, const void* ct_out, void* ct_in, double* grad_z, int gz_stride, const void* H,
                    const double* z, int z_stride, int N, int nfields, int nbatch, double dx, double dy, double k, int flags,
                    void* ws, size_t ws_bytes, void* stream);
int xl_vrs_fwd_batch(const void* ex, const void* ey, long long in_bstride, void* out, void* H, const double* z, int z_stride,
                     int N, int nbatch, double x0, double y0, double dx, double dy, double k, int flags,
                     void* ws, size_t ws_bytes, void* stream);
int xl_vrs_bwd_batch(const void* ex, const void* ey, long long in_bstride, const void* out, const void* ct_out, void* ct_exy,
                     double* grad_z, int gz_stride, const void* H, const double* z, int z_stride, int N, int nbatch,
                     double x0, double y0, double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream);
size_t xl_czt_workspace_bytes_batch(int N, int Mx, int My, int vectorial, int nbatch);
int xl_czt_fwd_batch(const void* in, const void* ey, long long in_bstride, void* out, const double* z, int z_stride, double lambda,
                     int N, int Mx, int My, int vectorial, int nbatch,
                     double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                     int flags, void* tables, void* ws, size_t ws_bytes, void* stream);
int xl_czt_bwd_batch(const void* ct_out, void* ct_in, const double* z, int z_stride, double lambda,
                     int N, int Mx, int My, int vectorial, int nbatch,
                     double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                     int flags, void* tables, void* ws, size_t ws_bytes, void* stream);
int xl_highna_fwd_batch(const void* ex, const void* ey, long long in_bstride, void* out, int N, int Mx, int My, int nbatch,
                        double radius, double f, double lambda, double x0, double dx, double y0, double dy,
                        double xout0, double xoutl, double yout0, double youtl, int flags, void* tables, void* ws, size_t ws_bytes, void* stream);
int xl_highna_bwd_batch(const void* ct_out, void* ct_exy, int N, int Mx, int My, int nbatch, double radius, double f, double lambda,
                        double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                        int flags, void* tables, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- pointwise Jones elements (vectorial tables) */
/* The elements that sit between the propagations of a vectorial optical table, each as ONE pass over the
 * planes it changes, forward and VJP, with the reference's parameter maps and the reductions of the scalar-parameter
 * gradients inside the kernels.  Planes are complex64 [n] (n = N*N pixels); cotangents follow torch's convention
 * (g = dL/dRe + i dL/dIm of a real loss), any cotangent pointer may be NULL (a zero cotangent on input, an unwanted
 * gradient on output; the Ex and Ey planes of one beam come together).  Scalar parameters are float64 in device memory and
 * enter as the optimizer's raw values p: angle = scale * p + offset (the reference maps p in (0,1) to p*2pi - pi,
 * optical_elements.py:1556-1592; scale = 1, offset = 0 passes angles through).  Parameter gradients are ACCUMULATED (+=).
 * `scratch`: xl_el_scratch_bytes() bytes of device memory (needed when a scalar-parameter gradient is requested). */
size_t xl_el_scratch_bytes(void);
/* sSLM: ox = ex * exp(i (scale*alpha + offset)), oy = ey * exp(i (scale*phi + offset)); alpha, phi float32 [n] planes.
 * optical_elements.py:186-222.  VJP: g_ex, g_ey and the per-pixel g_alpha, g_phi (float32, overwritten). */
int xl_el_sslm(const void* ex, const void* ey, const float* alpha, const float* phi, double scale, double offset,
               void* ox, void* oy, size_t n, void* stream);
int xl_el_sslm_bwd(const void* ex, const void* ey, const float* alpha, const float* phi, double scale, double offset,
                   const void* g_ox, const void* g_oy, void* g_ex, void* g_ey, float* g_alpha, float* g_phi,
                   size_t n, void* stream);
/* LCD: uniform retarder eta with its fast axis at theta: (ox, oy) = [[a, b], [b, d]] (ex, ey), a = cos(eta/2) - i sin(eta/2) cos 2theta,
 * b = -i sin(eta/2) sin 2theta, d = conj-mirror of a.  optical_elements.py:123-140, 170-180, 266-305. */
int xl_el_lcd(const void* ex, const void* ey, const double* eta, const double* theta, double scale, double offset,
              void* ox, void* oy, size_t n, void* stream);
int xl_el_lcd_bwd(const void* ex, const void* ey, const double* eta, const double* theta, double scale, double offset,
                  const void* g_ox, const void* g_oy, void* g_ex, void* g_ey, double* g_eta, double* g_theta,
                  void* scratch, size_t n, void* stream);
/* BS_symmetric: c = R a + i T b, d = i T a + R b, T = 0.99 |cos theta|, R = |sin theta| - 0.01 |cos theta|.
 * optical_elements.py:334-392. */
int xl_el_bs(const void* a_ex, const void* a_ey, const void* b_ex, const void* b_ey, const double* theta, double scale, double offset,
             void* c_ex, void* c_ey, void* d_ex, void* d_ey, size_t n, void* stream);
int xl_el_bs_bwd(const void* a_ex, const void* a_ey, const void* b_ex, const void* b_ey, const double* theta, double scale, double offset,
                 const void* g_c_ex, const void* g_c_ey, const void* g_d_ex, const void* g_d_ey,
                 void* g_a_ex, void* g_a_ey, void* g_b_ex, void* g_b_ey, double* g_theta,
                 void* scratch, size_t n, void* stream);

/* ---------------------------------------------------------------- instrumentation (bench.py) ---------------- */
/* Number of kernels this library has launched in this process. */
long long xl_launch_count(void);
/* Per-kernel timing with CUDA events recorded on the launching stream: xl_prof_enable(1) clears and starts recording,
 * xl_prof_report() synchronises the recorded events and writes "name count total_ms" lines; xl_prof_enable(0) stops. */
void xl_prof_enable(int on);
int xl_prof_report(char* buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* XLPROP_H */
