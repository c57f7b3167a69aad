/*
 * pdephi_b200.h -- C ABI of the B200-native spectral-convolution hot path of PDE-Fluid-Phi.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer on the
 * plan's / current CUDA device unless stated otherwise; complex tensors are interleaved
 * (re, im) float pairs, i.e. the memory layout of torch.complex64.  All functions are
 * asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise the host,
 * never allocate persistent memory except the plan's constant twiddle tables, and return
 * PDEPHI_OK or an error code; pdephi_last_error() gives the message of the calling thread's
 * last failure.
 *
 * The reference is 100 % Python and has no FFI: the seam these entry points replace is the
 * body of two nn.Module.forward methods and their autograd backward (paths below are relative
 * to /root/reference/src/pde_fluid_phi):
 *
 *   RationalFourierLayer.forward   operators/rational_fourier.py:150-172
 *   SpectralConv3D.forward         operators/spectral_layers.py:46-78
 *
 * Notation: B batch, I/O in/out channels, grid Nx,Ny,Nz (z contiguous), retained modes
 * mx,my,mz, mzh = mz/2+1, M = mx*my*mzh.
 */
#ifndef PDEPHI_B200_H
#define PDEPHI_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEPHI_VERSION 100 /* 0.1.0 */

typedef struct pdephi_plan_s* pdephi_plan_t;
typedef void* pdephi_stream_t; /* cudaStream_t */

enum pdephi_status {
  PDEPHI_OK = 0,
  PDEPHI_ERR_INVALID = 1,     /* bad argument (null pointer, size mismatch, modes > grid ...) */
  PDEPHI_ERR_UNSUPPORTED = 2, /* valid but not implemented (e.g. a rational order that is not instantiated) */
  PDEPHI_ERR_CUDA = 3,        /* a CUDA runtime call or launch failed */
  PDEPHI_ERR_WORKSPACE = 4    /* workspace too small */
};

/* Arithmetic of the transform stages.  The channel mix, R(k) evaluation and projection always
 * run in IEEE fp32 with the reference's rounding order (bit-identical P, Q, R). */
enum pdephi_precision {
  PDEPHI_PREC_FP32 = 0, /* fp32 FFMA contractions; parity gate 1e-5 */
  PDEPHI_PREC_TF32 = 1, /* tcgen05 kind::tf32 contractions, fp32 accumulate; parity gate 2e-3 */
  PDEPHI_PREC_TF32X3 = 2 /* tcgen05 split contractions: every fp32 operand as a (hi, lo) TF32 pair, three MMAs per
                            product, fp32 accumulate -- fp32-grade results at tensor-core speed; parity gate 1e-5 */
};

/* Which member of a transform pair.
 *  analysis : FORWARD = pruned rfftn                      ADJOINT = backward of synthesis(FORWARD)
 *  synthesis: FORWARD = zero-padded irfftn (1/N, C2R)     ADJOINT = backward of analysis(FORWARD) */
enum pdephi_direction { PDEPHI_FORWARD = 0, PDEPHI_ADJOINT = 1 };

/* Which stages of a transform one call runs.  The (z,y) stages act on every x plane independently and the x stage on
 * every (ky,kz) column independently: the seam of the slab-decomposed transform, where an all-to-all re-partitions the
 * small intermediate T [BC, Nx, my, mzh] complex between "x slab, all ky" and "all x, ky subset" (SURVEY.md section 8e). */
enum pdephi_stages { PDEPHI_STAGES_ZY = 1, PDEPHI_STAGES_X = 2, PDEPHI_STAGES_ALL = 3 };

int pdephi_version(void);
const char* pdephi_last_error(void);

/* ---- plan: constant DFT twiddle tables for one (grid, modes) pair on one device ------------- */
int pdephi_plan_create(pdephi_plan_t* plan, int device, int Nx, int Ny, int Nz, int mx, int my, int mzh);
/* Plan for an x slab [x_offset, x_offset+Nx) of a grid with Nx_total planes (slab-decomposed transform over
 * several GPUs): K1 then returns the slab's partial sums of the retained modes (summing them over the slabs
 * gives the transform of the whole volume) and K3 returns the slab's rows of the synthesis; the 1/N of the
 * normalised directions uses Nx_total. */
int pdephi_plan_create_slab(pdephi_plan_t* plan, int device, int Nx, int Ny, int Nz, int mx, int my, int mzh,
                            int Nx_total, int x_offset);
int pdephi_plan_destroy(pdephi_plan_t plan);
/* 1 if the plan's (grid, modes) shape is covered by the kernels of `precision` (FP32 covers every shape; the fused TF32 /
 * TF32X3 kernels need Ny = 128, Nz in {64,128}, my <= 32, mzh <= 20; both additionally cover, through the per-axis
 * kernels, axes that are multiples of 32 up to 256 with up to 64 retained modes each and 2-D grids [Nx,Ny,1]), else 0 */
int pdephi_plan_supports(pdephi_plan_t plan, int precision);
/* Which kernel family a transform of `precision` takes on this plan's shape.  The host side uses it to choose between the
 * fp32-grade routes ('auto'): the fused kernels win at every size they cover, the per-axis split kernels (three launches
 * per axis stage) only once the volume is large enough to hide their launch count. */
enum pdephi_route {
  PDEPHI_ROUTE_NONE = 0,     /* precision not available for this shape */
  PDEPHI_ROUTE_FFMA = 1,     /* transforms_simt.cu */
  PDEPHI_ROUTE_FUSED_TC = 2, /* fused (z,y) + x stage tcgen05 kernels (transforms_tc.cu / transforms_tc3.cu / xstage_tc.cu) */
  PDEPHI_ROUTE_PER_AXIS_TC = 3 /* one tcgen05 contraction per axis (transforms_gen.cu) */
};
int pdephi_plan_route(pdephi_plan_t plan, int precision);
/* bytes of scratch the two transforms need for `BC` (= batch*channels) volumes */
size_t pdephi_plan_workspace_bytes(pdephi_plan_t plan, long long BC);

/* ---- K1: pruned real-to-complex analysis transform ------------------------------------------
 * replaces torch.fft.rfftn(x, dim=[-3,-2,-1])[:, :, :mx, :my, :mzh]
 *          (rational_fourier.py:161 + :107, spectral_layers.py:59 + :63)
 *   x [BC, Nx, Ny, Nz] float  ->  X [BC, mx, my, mzh] complex
 * ADJOINT additionally multiplies by c_kz/(Nx*Ny*Nz): the backward of K3(FORWARD). */
int pdephi_analysis_r2c(pdephi_plan_t plan, const float* x, float* X, long long BC, int direction,
                        int precision, void* workspace, size_t workspace_bytes, pdephi_stream_t stream);

/* K1 by stages.  ZY: in = x [BC,Nx,Ny,Nz] float -> out = T [BC,Nx,my,mzh] complex (ADJOINT carries the c_kz/N weights);
 * X: in = T [BC,Nx,my,mzh] complex -> out = X [BC,mx,my,mzh] complex (the plan's Nx planes and my*mzh columns: a plan
 * built for the whole x extent and this rank's ky subset);  ALL = pdephi_analysis_r2c.  2-D grids: ALL only. */
int pdephi_analysis_stages(pdephi_plan_t plan, const float* in, float* out, long long BC, int direction, int precision,
                           int stages, void* workspace, size_t workspace_bytes, pdephi_stream_t stream);

/* ---- K3: pruned complex-to-real synthesis transform, fused scaling and skip adds ------------
 * replaces zeros_like/zeros + slice-assign + torch.fft.irfftn(out_ft, s=x.shape[-3:])
 *          (rational_fourier.py:116-117,170, stability.py:85-86, spectral_layers.py:69-76)
 *   Z [BC, mx, my, mzh] complex  ->  out [BC, Nx, Ny, Nz] float
 *   out = S(Z * mode_scale[k] * row_scale[bc]) + addend0 + addend1
 * mode_scale [M] (StabilityProjection decay mask), row_scale (element bc at
 * row_scale[bc*row_scale_stride], the energy-conservation factor), addend0/1 [BC,Nx,Ny,Nz]
 * (the block's "+ x_local + x", rational_fourier.py:271) may each be NULL.
 * FORWARD has irfftn semantics (imaginary part of the kz=0 / Nyquist bins ignored, 1/N);
 * ADJOINT is the plain unnormalised Re sum: the backward of K1(FORWARD). */
int pdephi_synthesis_c2r(pdephi_plan_t plan, const float* Z, float* out, const float* addend0,
                         const float* addend1, const float* mode_scale, const float* row_scale,
                         long long row_scale_stride, long long BC, int direction, int precision,
                         void* workspace, size_t workspace_bytes, pdephi_stream_t stream);

/* K3 by stages.  X: in = Z [BC,mx,my,mzh] -> out = T [BC,Nx,my,mzh] complex, scales applied, addends must be NULL;
 * ZY: in = T -> out [BC,Nx,Ny,Nz] float + addends, scales must be NULL (FORWARD carries c_kz/N);  ALL = pdephi_synthesis_c2r. */
int pdephi_synthesis_stages(pdephi_plan_t plan, const float* in, float* out, const float* addend0, const float* addend1,
                            const float* mode_scale, const float* row_scale, long long row_scale_stride, long long BC,
                            int direction, int precision, int stages, void* workspace, size_t workspace_bytes,
                            pdephi_stream_t stream);

/* ---- K2: rational channel mix + StabilityProjection statistics ------------------------------
 * replaces RationalFourierLayer.rational_multiply (rational_fourier.py:73-119) and the two
 * energy reductions of StabilityProjection (stability.py:90-103).
 *   X [B, I, M] complex, P [O, I, op,op,op], Q [O, I, oq,oq,oq] float,
 *   monoP [TP, M], monoQ [TQ, M]: monomial tables kx^a ky^b kz^c in the reference's loop order
 *   (rational_fourier.py:135-144), mask [M] decay mask (stability.py:40-57).
 *   Y [B, O, M] complex   = sum_i R[o,i,k] X[b,i,k],  R = P(k) / (Q(k) + eps)  (unmasked, unscaled)
 *   stats [B*O, 4] float  = { a = sum|Y|^2 + proj_eps, b = sum mask^2|Y|^2 + proj_eps, s = sqrt(a/b), 0 }
 * The projected modes are Z = mask * s * Y; K3 applies that scaling on the fly.
 * R_out / Qe_out [O, I, M] float (both or neither; may be NULL for inference): the transfer function and
 * Q(k)+eps, materialised once per forward so the backward streams them instead of re-evaluating. */
int pdephi_rational_mix_fwd(const float* X, const float* P, const float* Q, const float* monoP,
                            const float* monoQ, int order_p, int order_q, float eps, const float* mask,
                            float proj_eps, float* Y, float* stats, float* R_out, float* Qe_out, int B, int I,
                            int O, long long M, pdephi_stream_t stream);

/* ---- K2 with a materialised transfer function (the cached forward) ---------------------------------------------
 * R(k) = P(k)/(Q(k)+eps) depends on the coefficients only (rational_fourier.py:92-104,135-146), not on the input: a forward
 * pass with unchanged coefficients (inference, the four evaluations of an RK4 rollout step, gradient accumulation) can reuse
 * the copies pdephi_rational_mix_fwd materialised.
 *  pdephi_rational_transfer      R_out, Qe_out [O,I,M] alone, bit-identical to what pdephi_rational_mix_fwd writes.  With
 *                                snapP / snapQ (device copies of P / Q the caller keeps next to R, Qe) and `changed` (one device
 *                                int) it is a REFRESH: a guard kernel compares the coefficients with the snapshots bit for bit
 *                                (updating them) and the evaluation runs only if something differs -- two few-microsecond
 *                                launches when nothing has changed.  (A host-side version counter alone misses writes through
 *                                `param.data`, e.g. the reference's broadcast_parameters, training/distributed.py:230.)
 *  pdephi_rational_mix_apply     Y = sum_i R X and the projection statistics from a materialised R: same accumulation order
 *                                as pdephi_rational_mix_fwd, so Y and stats are bit-identical to the fused evaluation. */
int pdephi_rational_transfer(const float* P, const float* Q, const float* monoP, const float* monoQ, int order_p, int order_q,
                             float eps, float* R_out, float* Qe_out, int I, int O, long long M, float* snapP, float* snapQ,
                             int* changed, pdephi_stream_t stream);
int pdephi_rational_mix_apply(const float* X, const float* R, const float* mask, float proj_eps, float* Y, float* stats, int B,
                              int I, int O, long long M, pdephi_stream_t stream);

/* Backward of Z = mask*s*Y (stability.py:59-103) :  dZ [B*O, M] complex -> dY [B*O, M] complex */
int pdephi_projection_bwd(const float* dZ, const float* Y, const float* mask, const float* stats,
                          float* dY, long long rows, long long M, pdephi_stream_t stream);

/* The same backward when the M modes of a row are sharded over several GPUs (slab-decomposed transform, all-to-all
 * route): phase 1 writes this shard's partial G[row] = sum_k Re(conj(dZ) mask Y) (dY untouched); the caller sums G over
 * the shards (NCCL all-reduce of `rows` floats); phase 2 applies dY = m s dZ + G s (1/a - m^2/b) Y with the summed G.
 * `stats` holds the GLOBAL a, b, s of each row. */
int pdephi_projection_bwd_sharded(const float* dZ, const float* Y, const float* mask, const float* stats, float* G,
                                  float* dY, long long rows, long long M, int phase, pdephi_stream_t stream);

/* Backward of the rational mix from the materialised R, Qe of the forward: dX [B,I,M] complex,
 * dP [O,I,op^3], dQ [O,I,oq^3] (entries with a+b+c >= order are written as zeros, as autograd does).
 * Deterministic two-stage reduction. */
size_t pdephi_rational_mix_bwd_workspace_bytes(int order_p, int order_q, int I, int O, long long M);
int pdephi_rational_mix_bwd(const float* X, const float* dY, const float* R, const float* Qe,
                            const float* monoP, const float* monoQ, int order_p, int order_q, float* dX,
                            float* dP, float* dQ, int B, int I, int O, long long M, void* workspace,
                            size_t workspace_bytes, pdephi_stream_t stream);

/* ---- K2': complex channel mix of SpectralConv3D (spectral_layers.py:66) ---------------------
 *   W [O, I, mxy, mz] complex parameter; only kz < mzh is used (oracle convention 2).
 *   Y[b,o,k] = sum_i W[o,i,k] X[b,i,k];   dX = sum_o conj(W) dY;   dW = sum_b conj(X) dY (0 for kz>=mzh) */
int pdephi_complex_mix_fwd(const float* X, const float* W, float* Y, int B, int I, int O, long long mxy,
                           int mzh, int mz, pdephi_stream_t stream);
int pdephi_complex_mix_bwd(const float* X, const float* W, const float* dY, float* dX, float* dW, int B,
                           int I, int O, long long mxy, int mzh, int mz, pdephi_stream_t stream);

/* ---- channels-first pointwise linear maps: the models' lift / project ------------------------
 * replaces rearrange -> nn.Linear -> rearrange (rational_fourier.py:258-260,274-276; fno3d.py:86-88,100-102)
 *   out[b,o,s] = bias[o] + sum_c W[o,c] x[b,c,s]      x [B,Cin,S], W [Cout,Cin], out [B,Cout,S]
 * transpose_w != 0 uses W as [Cin,Cout] (the input-gradient map).  One of Cin, Cout must be <= 8, S % 4 == 0.
 * wgrad: dW[o,c] = sum_{b,s} g[b,o,s] x[b,c,s], dbias[o] = sum_{b,s} g[b,o,s]; deterministic two-stage sum. */
int pdephi_pointwise_linear(const float* x, const float* W, const float* bias, float* out, int B, int Cin,
                            int Cout, long long S, int transpose_w, pdephi_stream_t stream);
size_t pdephi_pointwise_wgrad_workspace_bytes(int Cin, int Cout, long long S);
int pdephi_pointwise_wgrad(const float* x, const float* g, float* dW, float* dbias, int B, int Cin, int Cout,
                           long long S, void* workspace, size_t workspace_bytes, pdephi_stream_t stream);

/* ---- fused block tail of a width-64 operator model (tcgen05 TF32 channel GEMM + epilogue) -----
 * replaces Conv3d 1x1x1 + the two residual adds + exact GELU of rational_fourier.py:268-271 / fno3d.py:93-97
 *   fwd:  u = spec + (Wc h + bc) + h,  hout = gelu(u)            h, spec, u, hout [B,64,S];  Wc [64,64]
 *   bwd:  gu = g * gelu'(u),  t = gu + Wc^T gu,  dWc = sum_{b,s} gu (x) h,  dbc = sum_{b,s} gu   (t may be NULL: not stored)
 * Needs 64 channels and S % 128 == 0 (PDEPHI_ERR_UNSUPPORTED otherwise). */
int pdephi_block_tail_fwd(const float* h, const float* spec, const float* Wc, const float* bc, float* u,
                          float* hout, int B, int C, long long S, pdephi_stream_t stream);
size_t pdephi_block_tail_bwd_workspace_bytes(int C);
int pdephi_block_tail_bwd(const float* g, const float* u, const float* h, const float* Wc, float* gu, float* t,
                          float* dWc, float* dbc, int B, int C, long long S, void* workspace,
                          size_t workspace_bytes, pdephi_stream_t stream);

/* ---- whole spectral branch + block tail, forward, with the spectral output never materialised ---------------
 * u = S(Y * mode_scale * row_scale) + (Wc h + bc) + h,  hout = gelu(u)       (rational_fourier.py:265-271)
 * Y [B*64, mx, my, mzh] complex: the mixed, un-projected modes (output of pdephi_rational_mix_fwd / complex_mix_fwd).
 * The x and y stages of the synthesis run as per-axis tensor-core contractions and leave V [B,Nx,Ny,64,ldz] (28 % of an
 * activation at mzh = 17); the z stage is two more MMA groups inside the block-tail kernel, accumulated into the same
 * TMEM tile as the 1x1x1 convolution -- the 4.3 GB write and read-back of `spec` per layer disappear.
 * Needs width 64, Nz = 128 and a (grid, modes) shape of the per-axis route (pdephi_block_spectral_supported). */
int pdephi_block_spectral_supported(pdephi_plan_t plan, int C);
size_t pdephi_block_spectral_fwd_workspace_bytes(pdephi_plan_t plan, int B, int C);
int pdephi_block_spectral_fwd(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                              long long row_scale_stride, const float* h, const float* Wc, const float* bc, float* u,
                              float* hout, int B, int C, void* workspace, size_t workspace_bytes, pdephi_stream_t stream);

/* ---- the LAST block of a model together with the output projection (nn.Linear 64 -> n_out <= 4) --------------
 * replaces rational_fourier.py:274-276 / fno3d.py:100-102 on top of the block above.
 *   forward : additionally out[b,n,s] = bp[n] + sum_c Wp[n,c] hout[b,c,s], computed in the block-tail epilogue (the separate
 *             projection pass and its read of hout disappear; hout is still written for the projection's weight gradient)
 *   backward: the block's incoming gradient g' = Wp^T gout is generated inside pdephi_block_tail_proj_bwd from
 *             gout [B,n_out,S] instead of being written by one kernel and read by the next.
 * Wp [n_out,64] row-major, bp [n_out] or NULL. */
int pdephi_block_tail_proj_fwd(const float* h, const float* spec, const float* Wc, const float* bc, float* u, float* hout,
                               const float* Wp, const float* bp, int n_out, float* out, int B, int C, long long S,
                               pdephi_stream_t stream);
int pdephi_block_spectral_proj_fwd(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                                   long long row_scale_stride, const float* h, const float* Wc, const float* bc, float* u,
                                   float* hout, const float* Wp, const float* bp, int n_out, float* out, int B, int C,
                                   void* workspace, size_t workspace_bytes, pdephi_stream_t stream);
int pdephi_block_tail_proj_bwd(const float* gout, const float* Wp, int n_out, const float* u, const float* h, const float* Wc,
                               float* gu, float* t, float* dWc, float* dbc, int B, int C, long long S, void* workspace,
                               size_t workspace_bytes, pdephi_stream_t stream);

/* ---- the FIRST block of a model together with the input lift (nn.Linear n_in <= 4 -> 64) -------------------
 * replaces rational_fourier.py:258-260 / fno3d.py:86-88 under the block above: the block's input h = Win x + bin is never
 * written to or read from HBM.  The fused forward folds it into the epilogue (u = spec + (Wc + I) Win x + (Wc + I) bin + bc,
 * n_in FMAs per element, no 64-channel operand tile); the unfused forward and the backward with t generate it tile by tile.
 * Win [64,n_in] row-major, bin [64] or NULL.  (The block's modes follow from the modes of x: A(h) = Win A(x) + bin N delta_k0.) */
int pdephi_block_tail_lift_fwd(const float* x, const float* Win, const float* bin, int n_in, const float* spec, const float* Wc,
                               const float* bc, float* u, float* hout, int B, int C, long long S, pdephi_stream_t stream);
int pdephi_block_spectral_lift_fwd(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                                   long long row_scale_stride, const float* x, const float* Win, const float* bin, int n_in,
                                   const float* Wc, const float* bc, float* u, float* hout, int B, int C, void* workspace,
                                   size_t workspace_bytes, pdephi_stream_t stream);
/* Backward.  t == NULL (the model input needs no gradient): the block needs neither t nor an h0 tile -- dWc = G Win^T +
 * dbc (x) bin with G [64,n_in] = sum_s gu (x) x accumulated on the tensor cores against the staged x rows; G (may be NULL) is
 * also returned: it is what the lift's own weight gradient needs (dWin = (I + Wc^T) G + the mode-space term). */
int pdephi_block_tail_lift_bwd(const float* g, const float* u, const float* x, const float* Win, const float* bin, int n_in,
                               const float* Wc, float* gu, float* t, float* dWc, float* dbc, float* G, int B, int C, long long S,
                               void* workspace, size_t workspace_bytes, pdephi_stream_t stream);

/* ---- fused block BACKWARD (mirror of pdephi_block_spectral_fwd) ----------------------------------------------------------
 * The backward of `gelu(spectral + conv1x1 + x)` (rational_fourier.py:263-271) needs gu = g gelu'(u) only as the input of the
 * ADJOINT analysis that follows it.  Here the z stage of that analysis runs inside the block-tail backward kernel (one more
 * tcgen05 contraction on the staged gu tile, a tile being one z line), gu is never written to or read from HBM, and the y / x
 * stages run as per-axis contractions on the pruned channel-innermost intermediate.  Results: t = gu + Wc^T gu (may be NULL:
 * first block of a model whose input needs no gradient), dWc, dbc (may be NULL), and dZ [B,C,mx,my,mzh] complex64 = the
 * ADJOINT analysis of gu (what pdephi_analysis_r2c(gu, PDEPHI_ADJOINT) would return).  Optional ends of the model as in
 * pdephi_block_tail_proj_bwd (Wp, n_out: `g` is then the projected output's gradient) and pdephi_block_tail_lift_bwd (Win, bin,
 * n_in, G: `h` is then the model input).  Same shapes as pdephi_block_spectral_supported.  tail_workspace:
 * pdephi_block_tail_bwd_workspace_bytes(C); workspace: pdephi_block_spectral_bwd_workspace_bytes. */
size_t pdephi_block_spectral_bwd_workspace_bytes(pdephi_plan_t plan, int B, int C);
int pdephi_block_spectral_bwd(pdephi_plan_t plan, const float* g, const float* u, const float* h, const float* Wc, float* t,
                              float* dWc, float* dbc, float* dZ, int B, int C, const float* Wp, int n_out, const float* Win,
                              const float* bin, int n_in, float* G, void* tail_workspace, size_t tail_workspace_bytes,
                              void* workspace, size_t workspace_bytes, pdephi_stream_t stream);

/* ---- fused block forward that also starts the NEXT layer's analysis ---------------------------------------------------------
 * pdephi_block_spectral_fwd_z = pdephi_block_spectral_fwd (Win == NULL) or pdephi_block_spectral_lift_fwd, plus: the z stage of
 * the FORWARD analysis of the block's output hout runs inside the kernel (the epilogue also writes hout, rounded to TF32, into a
 * K-major tile; one more tcgen05 contraction against a TMEM-resident twiddle operand) and leaves vz, the per-axis route's
 * z-analysed intermediate [B*C][Nx*Ny][ldz] (pdephi_block_spectral_vz_bytes).  pdephi_analysis_from_z finishes that analysis
 * (y stage, split x stage): X [BC,mx,my,mzh] complex64 = what pdephi_analysis_r2c(hout, PDEPHI_FORWARD, TF32) returns to the
 * TF32 gate -- the next block reads 1.2 GB instead of the 4.3 GB activation.  workspace of both: as pdephi_block_spectral_fwd. */
size_t pdephi_block_spectral_vz_bytes(pdephi_plan_t plan, int B, int C);
int pdephi_block_spectral_fwd_z(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                                long long row_scale_stride, const float* h, const float* Win, const float* bin, int n_in,
                                const float* Wc, const float* bc, float* u, float* hout, float* vz, int B, int C, void* workspace,
                                size_t workspace_bytes, pdephi_stream_t stream);
int pdephi_analysis_from_z(pdephi_plan_t plan, const float* vz, float* X, long long BC, void* workspace, size_t workspace_bytes,
                           pdephi_stream_t stream);

/* ---- pointwise maps with up to 64 channels on BOTH sides, fp32 FFMA, any spatial size (pointwise_wide.cu) --------------
 * wide_linear: out[b,o,s] = bias[o] + sum_i W(o,i) x[b,i,s] (+ add0 + add1, each [B,Cout,S] or NULL);
 * W(o,i) = transpose_w ? W[i*Cout + o] : W[o*Cin + i].  Replaces rearrange -> nn.Linear -> rearrange for the lift / projection
 * of operators built with in_channels = out_channels = width (models/multiscale_fno.py:100-127), and its backward
 * (dx: the same call with transpose_w = 1; dW, db: pdephi_wide_wgrad). */
int pdephi_wide_linear(const float* x, const float* W, const float* bias, const float* add0, const float* add1, float* out, int B,
                       int Cin, int Cout, long long S, int transpose_w, pdephi_stream_t stream);
size_t pdephi_wide_wgrad_workspace_bytes(int B, long long S);
int pdephi_wide_wgrad(const float* x, const float* g, float* dW, float* dbias, int B, int Cin, int Cout, long long S, void* workspace,
                      size_t workspace_bytes, pdephi_stream_t stream);
/* The block tail (see pdephi_block_tail_fwd / _bwd) for ANY width C <= 64, any spatial size and the activations the models are
 * usually built with (fno3d.py:57 `getattr(F, activation)`), fp32: u = spec + (Wc h + bc) + h, hout = act(u);
 * gu = g act'(u), t = gu + Wc^T gu, dWc = sum gu (x) h, dbc = sum gu.  workspace: pdephi_wide_wgrad_workspace_bytes(B, S). */
enum pdephi_activation {
  PDEPHI_ACT_IDENTITY = 0, PDEPHI_ACT_GELU = 1 /* exact (erf) */, PDEPHI_ACT_RELU = 2, PDEPHI_ACT_TANH = 3, PDEPHI_ACT_SILU = 4,
  PDEPHI_ACT_SIGMOID = 5, PDEPHI_ACT_LEAKY_RELU = 6 /* slope 0.01 */, PDEPHI_ACT_ELU = 7 /* alpha 1 */
};
int pdephi_wide_block_tail_fwd(const float* h, const float* spec, const float* Wc, const float* bc, float* u, float* hout, int B,
                               int C, long long S, int activation, pdephi_stream_t stream);
int pdephi_wide_block_tail_bwd(const float* g, const float* u, const float* h, const float* Wc, float* gu, float* t, float* dWc,
                               float* dbc, int B, int C, long long S, int activation, void* workspace, size_t workspace_bytes,
                               pdephi_stream_t stream);

/* ---- run-time options ------------------------------------------------------------------------------
 * PDEPHI_OPT_TF32_XSTAGE_SPLIT (default 1): on the single-pass TF32 route the x stage -- which works on the pruned
 * intermediate, 6.6 % of an activation -- runs split (fp32-grade), removing one of the transform's three TF32 roundings.
 * PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE (default 1): the `u` output of the width-64 block-tail FORWARD kernels (pdephi_block_tail_*_fwd,
 * pdephi_block_spectral_*_fwd) holds gelu'(pre-activation) instead of the pre-activation, and the backward kernels read it as
 * such: the backward needs u for nothing else, the forward has the Gaussian CDF / density in registers anyway (two more
 * instructions per element), and the backward's compute warps -- its bound -- drop ~25 instructions per element.  Same bits
 * in gu either way.  Set 0 to get the pre-activation itself (both directions must run under the same setting).
 * Value 2 (the default): the derivative is stored as IEEE binary16, one 32-bit word per channel PAIR and point -- `u` is then
 * an opaque buffer of B*(C/2)*S words, laid out [B][C/2][S], low half = channel 2p, high half = channel 2p+1 -- which takes
 * half an activation off the forward's writes and the backward's reads.  These kernels contract in single-pass TF32
 * whichever route calls them; binary16 carries the 11 significant bits of a TF32 operand, gelu' lies in [-0.13, 1.13], and
 * gu = g gelu'(u) is a TF32 operand of every product of the backward (the residual term of t sees the 2^-12 rounding
 * directly).  Value 1 keeps it in fp32. */
enum pdephi_option { PDEPHI_OPT_TF32_XSTAGE_SPLIT = 0, PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE = 1 };
int pdephi_set_option(int key, int value);
int pdephi_get_option(int key);

/* ---- per-kernel launch counters and CUDA-event timing (measurement support for bench.py) ------
 * Launches are always counted.  With timing enabled every kernel launch is bracketed by a pair of
 * cudaEvents on the launching stream; pdephi_profile_query synchronises on them and returns the
 * launch count and summed device time since the last reset. */
int pdephi_profile_enable(int on);
int pdephi_profile_reset(void);
int pdephi_profile_num_kernels(void);
const char* pdephi_profile_kernel_name(int kernel_id);
int pdephi_profile_query(int kernel_id, long long* launches, double* total_ms);

#ifdef __cplusplus
}
#endif
#endif /* PDEPHI_B200_H */
