// Shared helpers for the pdephi_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pdephi_b200.h"

namespace pdephi {

// thread-local error message returned by pdephi_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);

#define PDEPHI_CUDA(call)                                                                         \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return ::pdephi::fail(PDEPHI_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                  \
  } while (0)

#define PDEPHI_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess)                                                                       \
      return ::pdephi::fail(PDEPHI_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define PDEPHI_REQUIRE(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return ::pdephi::fail(PDEPHI_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// general per-axis tensor-core route (transforms_gen.cu)
struct GenPlan {
  int has_mid;              // 1: 3-D problem (z, y, x stages); 0: 2-D problem [Nx, Ny, 1] (y takes the role of z)
  int zN, zm, ldz, nbm;     // contiguous axis length, its retained modes, padded row length, MMA N of the z analysis
  int ym;                   // rows of the middle axis carried through the x stage (my, or 1 for 2-D)
  int mky, mkx;             // K of the axis-synthesis contractions (32 or 64)
  int za_ring, zs_ring, xa_st, xs_st, ya_st, ys_st;   // pipeline depths that fit shared memory
  float* za[2];             // z analysis  B images [direction]
  float* zs[2];             // z synthesis B images [direction]
  float* za_plain;          // adjoint z analysis as a plain [128][128] TMEM-resident A operand (fused block backward; zN = 128 only)
  float* za_plain_fwd;      // the same for the FORWARD direction (plain e^{-i theta}): z analysis of a block's output in the fused forward
  float* ya; float* ysc; float* yss;   // y axis: analysis A (plain [128][Ny]), synthesis cos / sin images
  float* xa; float* xsc; float* xss;   // x axis
  // split-TF32 (fp32-grade) variant of the route: lo parts of the axis images, stacked [hi | lo] z images per column block
  int x3_ok;
  float* ya_lo; float* ysc_lo; float* yss_lo;
  float* xa_lo; float* xsc_lo; float* xss_lo;
  int za3_np, za3_nc, za3_ring, za3_lo;          // z analysis: np launches over ranges of k-blocks, nc = nbm output columns
  float* za3[2];                                 // [direction] stacked (hi rows | lo rows) image
  int zs3_np, zs3_c0[8], zs3_nc[8], zs3_ring[8]; // z synthesis: blocks of output positions
  float* zs3[2][8];
};

struct Plan {
  int device;
  int Nx, Ny, Nz, mx, my, mzh;
  int Nx_total, x_offset;   // x slab of a larger grid (Nx_total == Nx, x_offset == 0 for a whole volume)
  // e^{-i theta} tables for the analysis direction, e^{+i theta} for synthesis; (re, im) pairs.
  float2* az[2];  // [Nz][mzh]  analysis z twiddles; [1] carries c_kz/(NxNyNz)
  float2* ay;     // [Ny][my]
  float2* ax;     // [Nx][mx]
  float2* sx;     // [mx][Nx]
  float2* sy;     // [my][Ny]
  float2* sz[2];  // [mzh][Nz]  synthesis z twiddles (w cos, w sin); [0] carries c_kz/(NxNyNz)
  void* arena;    // one allocation backing all tables
  int num_sms;
  size_t smem_optin;
  // tcgen05 path: twiddle operands as swizzled K-major shared-memory images (transforms_tc.cu)
  int tc_ok;
  float tc_comp;    // epilogue factor undoing the TF32 truncation bias of raw data operands (1 + 2^-11)
  void* tc_arena;
  float* tc_b1[2];  // analysis z  [48][Nz]
  float* tc_a2;     // analysis y  [64][Ny]
  float* tc_ay;     // synthesis y [Ny][64]
  float* tc_bz[2];  // synthesis z [Nz][64]
  // tcgen05 x stage (xstage_tc.cu)
  int xtc_a_ok, xtc_s_ok;
  void* xtc_arena;
  float* xtc_ax;    // analysis  [64][Nx]   (cos rows | sin rows)
  float* xtc_sc;    // synthesis [Nx][32]   cos
  float* xtc_ss;    // synthesis [Nx][32]   sin
  // split (hi, lo) TF32 operands of the fp32-grade tensor-core path (transforms_tc3.cu, xstage_tc.cu)
  int tc3_ok;
  void* tc3_arena;
  float* tc3_b1[2];     // analysis z  [dir]  per k-block: hi rows 0..39 | lo rows 40..79, [Nz] columns
  float* tc3_a2;        // analysis y  plain [128][Ny]: rows 0..63 hi, 64..127 lo (TMEM-resident)
  float* tc3_ay[2];     // synthesis y [hi|lo] [Ny][64]
  float* tc3_bz[2][2];  // synthesis z [dir][hi|lo] [Nz][64]
  int xtc3_a_ok, xtc3_s_ok;
  void* xtc3_arena;
  float* xtc3_ax[2];
  float* xtc3_sc[2];
  float* xtc3_ss[2];
  int gen_ok;
  void* gen_arena;
  void* gen_arena3;   // operand images of the split-TF32 variant
  GenPlan g;
};
int tc_plan_init(Plan* p);
void tc_plan_destroy(Plan* p);
int xtc_plan_init(Plan* p);
void xtc_plan_destroy(Plan* p);
int tc3_plan_init(Plan* p);
void tc3_plan_destroy(Plan* p);
int gen_plan_init(Plan* p);
void gen_plan_destroy(Plan* p);
size_t gen_workspace_bytes(const Plan* p, long long BC);

enum KernelId {
  K_ANALYSIS_ZY = 0, K_ANALYSIS_X, K_SYNTHESIS_X, K_SYNTHESIS_ZY, K_ANALYSIS_ZY_TC, K_SYNTHESIS_ZY_TC,
  K_RATIONAL_MIX, K_RATIONAL_MIX_T, K_PROJ_STATS, K_PROJ_BWD, K_RATIONAL_DPQ, K_DPQ_REDUCE,
  K_COMPLEX_MIX, K_COMPLEX_MIX_T, K_COMPLEX_DW, K_POINTWISE_LINEAR, K_POINTWISE_WGRAD, K_BLOCK_TAIL_FWD, K_BLOCK_TAIL_BWD, K_ANALYSIS_X_TC, K_SYNTHESIS_X_TC, K_SCALE_MODES, K_RATIONAL_G,
  K_ANALYSIS_ZY_TC3, K_SYNTHESIS_ZY_TC3, K_ANALYSIS_X_TC3, K_SYNTHESIS_X_TC3,
  K_GEN_Z_ANALYSIS, K_GEN_Z_SYNTHESIS, K_GEN_AXIS_ANALYSIS, K_GEN_AXIS_SYNTHESIS, K_BLOCK_SPECTRAL_FWD, K_BLOCK_SPECTRAL_BWD,
  K_RATIONAL_TRANSFER, K_RATIONAL_MIX_APPLY, K_COUNT
};
// run-time options (pdephi_set_option)
enum { OPT_TF32_XSTAGE_SPLIT = 0, OPT_BLOCK_SAVES_DERIVATIVE = 1, OPT_COUNT };
int option_get(int key);
void prof_begin(int id, cudaStream_t st);
void prof_end(int id, cudaStream_t st);
struct ProfScope {
  int id; cudaStream_t st;
  ProfScope(int i, cudaStream_t s) : id(i), st(s) { prof_begin(id, st); }
  ~ProfScope() { prof_end(id, st); }
};

// which stages of a transform a call runs (include/pdephi_b200.h: pdephi_stages): the (z,y) stages act on x planes
// independently, the x stage on (ky,kz) columns independently -- the seam of the slab-decomposed transform
enum { ST_ZY = 1, ST_X = 2, ST_ALL = 3 };

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(float2& acc, float2 a, float2 b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y);
  acc.y = fmaf(a.y, b.x, acc.y);
}

// fixed-order block reduction of a double; result valid in thread 0 (and broadcast through smem)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red /* [THREADS/32] */) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) t += red[i];
    red[0] = t;
  }
  __syncthreads();
  t = red[0];
  return t;
}

}  // namespace pdephi
