// K1 / K3, fp32 path: pruned analysis and synthesis transforms as dense DFT contractions on the
// CUDA cores (FFMA), any grid / mode shape.  The (z,y) stages are fused per (volume, x) plane so the
// full-size activation crosses HBM exactly once; only the 15x smaller [BC,Nx,my,mzh] intermediate
// is staged through the workspace before / after the x stage.
//
// Replaces torch.fft.rfftn + slice (rational_fourier.py:161,107) and zero-pad + torch.fft.irfftn
// (rational_fourier.py:116-117,170) of the reference.
#include "common.cuh"

namespace pdephi {

constexpr int ZY_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// analysis, fused z+y stage.  One CTA per (bc, x) plane.
//   T1[y][kz] = sum_z x[y][z] * az[z][kz]           (real x complex)
//   T2[ky][kz] = sum_y ay[y][ky] * T1[y][kz]        (complex x complex)
// ---------------------------------------------------------------------------------------------
template <int G, bool VEC4>
__global__ void __launch_bounds__(ZY_THREADS)
analysis_zy_kernel(const float* __restrict__ x, float2* __restrict__ T2, const float2* __restrict__ az,
                   const float2* __restrict__ ay, int Ny, int Nz, int my, int mzh, int YC, int ldx, int tw_smem) {
  extern __shared__ __align__(16) float smem[];
  float2* T1s = reinterpret_cast<float2*>(smem);                         // [Ny][mzh]
  float* Xs = smem + ((2 * (size_t)Ny * mzh + 3) / 4) * 4;                // [YC][ldx]
  // twiddle tables staged in shared memory when they fit (tw_smem): warp-wide broadcast LDS instead of LDG
  const float2* azs = az;
  const float2* ays = ay;
  if (tw_smem) {
    float2* a = reinterpret_cast<float2*>(Xs + (size_t)YC * ldx + ((size_t)YC * ldx & 1));
    float2* b = a + (size_t)Nz * mzh;
    for (int i = threadIdx.x; i < Nz * mzh; i += ZY_THREADS) a[i] = __ldg(az + i);
    for (int i = threadIdx.x; i < Ny * my; i += ZY_THREADS) b[i] = __ldg(ay + i);
    azs = a;
    ays = b;
  }
  const size_t plane = blockIdx.x;
  const float* xp = x + plane * (size_t)Ny * Nz;
  const int NG = (mzh + G - 1) / G;

  for (int y0 = 0; y0 < Ny; y0 += YC) {
    const int rows = min(YC, Ny - y0);
    __syncthreads();
    if (VEC4) {
      const int nz4 = Nz >> 2;
      for (int i = threadIdx.x; i < rows * nz4; i += ZY_THREADS) {
        const int r = i / nz4, c = i - r * nz4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(xp + (size_t)(y0 + r) * Nz) + c);
        *reinterpret_cast<float4*>(&Xs[r * ldx + 4 * c]) = v;
      }
    } else {
      for (int i = threadIdx.x; i < rows * Nz; i += ZY_THREADS) {
        const int r = i / Nz, c = i - r * Nz;
        Xs[r * ldx + c] = __ldg(xp + (size_t)(y0 + r) * Nz + c);
      }
    }
    __syncthreads();
    for (int it = threadIdx.x; it < rows * NG; it += ZY_THREADS) {
      const int g = it / rows, y = it - g * rows;
      const int kz0 = g * G;
      int kzi[G];
#pragma unroll
      for (int j = 0; j < G; ++j) kzi[j] = min(kz0 + j, mzh - 1);
      float2 acc[G];
#pragma unroll
      for (int j = 0; j < G; ++j) acc[j] = make_float2(0.f, 0.f);
      const float* xr = &Xs[y * ldx];
      if (VEC4) {
        for (int z = 0; z < Nz; z += 4) {
          const float4 xv = *reinterpret_cast<const float4*>(xr + z);
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int dz = 0; dz < 4; ++dz) {
            const float2* trow = azs + (size_t)(z + dz) * mzh;
#pragma unroll
            for (int j = 0; j < G; ++j) {
              const float2 tw = trow[kzi[j]];
              acc[j].x = fmaf(xs[dz], tw.x, acc[j].x);
              acc[j].y = fmaf(xs[dz], tw.y, acc[j].y);
            }
          }
        }
      } else {
        for (int z = 0; z < Nz; ++z) {
          const float xv = xr[z];
          const float2* trow = azs + (size_t)z * mzh;
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const float2 tw = trow[kzi[j]];
            acc[j].x = fmaf(xv, tw.x, acc[j].x);
            acc[j].y = fmaf(xv, tw.y, acc[j].y);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < G; ++j)
        if (kz0 + j < mzh) T1s[(size_t)(y0 + y) * mzh + kz0 + j] = acc[j];
    }
  }
  __syncthreads();
  float2* outp = T2 + plane * (size_t)my * mzh;
  for (int it = threadIdx.x; it < my * mzh; it += ZY_THREADS) {
    const int ky = it / mzh, kz = it - ky * mzh;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int y = 0; y < Ny; ++y) cfma(acc, ays[(size_t)y * my + ky], T1s[(size_t)y * mzh + kz]);
    outp[it] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// x stages: per volume bc a [K x P] complex contraction against a twiddle matrix, P = my*mzh.
//   analysis : out[kx][p] = sum_x tw[x][kx] * in[x][p]        (tw = ax [Nx][mx])
//   synthesis: out[x][p]  = sum_kx tw[kx][x] * in[kx][p]*scale (tw = sx [mx][Nx])
// Both are "out[r][p] = sum_k tw[k][r] * in[k][p]" with R output rows and K contracted rows.
// A warp owns XT output rows x 64 modes (2 per lane, one 128-bit load per k); the twiddles of the
// CTA's row range sit in shared memory and are read as warp-wide broadcasts, so the inner loop is
// 4 LDS.128 + 1 LDG.128 per 64 FFMA.
// ---------------------------------------------------------------------------------------------
constexpr int XT = 8;   // output rows per warp
constexpr int XW = 8;   // warps (row groups) per CTA
template <bool SCALED>
__global__ void __launch_bounds__(32 * XW)
xstage_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, int K, int R,
              int P, const float* __restrict__ mode_scale, const float* __restrict__ row_scale,
              long long row_scale_stride) {
  extern __shared__ __align__(16) float2 tws[];      // [K][XT*XW] twiddles of this CTA's rows
  const int lane = threadIdx.x, wy = threadIdx.y;
  const size_t bc = blockIdx.y;
  const int rblk = blockIdx.z * (XT * XW);
  for (int i = wy * 32 + lane; i < K * XT * XW; i += 32 * blockDim.y) {
    const int k = i / (XT * XW), j = i - k * (XT * XW);
    tws[i] = (rblk + j < R) ? __ldg(tw + (size_t)k * R + rblk + j) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int r0 = rblk + wy * XT;
  if (r0 >= R) return;
  const int p0 = (blockIdx.x * 32 + lane) * 2;
  const bool pair = ((P & 1) == 0);                   // 128-bit path needs 16-byte aligned rows
  const bool v0 = p0 < P, v1 = p0 + 1 < P;
  float2 acc0[XT], acc1[XT];
#pragma unroll
  for (int j = 0; j < XT; ++j) acc0[j] = acc1[j] = make_float2(0.f, 0.f);
  const float2* inp = in + bc * (size_t)K * P;
  float rs = 1.f;
  if (SCALED && row_scale) rs = __ldg(row_scale + bc * row_scale_stride);
  const int pc0 = v0 ? p0 : 0, pc1 = v1 ? p0 + 1 : pc0;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float2 a, b;
    if (pair && v1) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(inp + (size_t)k * P + p0));
      a = make_float2(q.x, q.y);
      b = make_float2(q.z, q.w);
    } else {
      a = __ldg(inp + (size_t)k * P + pc0);
      b = __ldg(inp + (size_t)k * P + pc1);
    }
    if (SCALED) {
      float s0 = rs, s1 = rs;
      if (mode_scale) {
        s0 *= __ldg(mode_scale + (size_t)k * P + pc0);
        s1 *= __ldg(mode_scale + (size_t)k * P + pc1);
      }
      a.x *= s0; a.y *= s0; b.x *= s1; b.y *= s1;
    }
    const float4* t4 = reinterpret_cast<const float4*>(tws + (size_t)k * (XT * XW) + wy * XT);
#pragma unroll
    for (int j = 0; j < XT / 2; ++j) {
      const float4 t = t4[j];
      const float2 ta = make_float2(t.x, t.y), tb = make_float2(t.z, t.w);
      cfma(acc0[2 * j], ta, a);
      cfma(acc1[2 * j], ta, b);
      cfma(acc0[2 * j + 1], tb, a);
      cfma(acc1[2 * j + 1], tb, b);
    }
  }
  float2* outp = out + bc * (size_t)R * P;
#pragma unroll
  for (int j = 0; j < XT; ++j) {
    if (r0 + j < R) {
      if (pair && v1) {
        *reinterpret_cast<float4*>(outp + (size_t)(r0 + j) * P + p0) = make_float4(acc0[j].x, acc0[j].y, acc1[j].x, acc1[j].y);
      } else {
        if (v0) outp[(size_t)(r0 + j) * P + p0] = acc0[j];
        if (v1) outp[(size_t)(r0 + j) * P + p0 + 1] = acc1[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// synthesis, fused y+z stage.  One CTA per (bc, x) plane.
//   V[y][kz]  = sum_ky sy[ky][y] * U[ky][kz]                       (complex x complex)
//   out[y][z] = sum_kz V.re*sz[kz][z].re - V.im*sz[kz][z].im  (+ addend0 + addend1)
// ---------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(ZY_THREADS)
synthesis_zy_kernel(const float2* __restrict__ U, float* __restrict__ out, const float* __restrict__ add0,
                    const float* __restrict__ add1, const float2* __restrict__ sy_g, const float2* __restrict__ sz_g,
                    int Ny, int Nz, int my, int mzh, int tw_smem) {
  extern __shared__ __align__(16) float smem[];
  float2* Us = reinterpret_cast<float2*>(smem);        // [my][mzh]
  float2* Vs = Us + (size_t)my * mzh;                  // [Ny][mzh]
  const float2* sy = sy_g;
  const float2* sz = sz_g;
  if (tw_smem) {                                       // twiddles in shared memory (16-byte aligned for the float4 reads)
    float2* a = Vs + (((size_t)Ny * mzh + (size_t)my * mzh + 1) & ~(size_t)1) - (size_t)my * mzh;
    float2* b = a + (((size_t)mzh * Nz + 1) & ~(size_t)1);
    for (int i = threadIdx.x; i < mzh * Nz; i += ZY_THREADS) a[i] = __ldg(sz_g + i);
    for (int i = threadIdx.x; i < my * Ny; i += ZY_THREADS) b[i] = __ldg(sy_g + i);
    sz = a;
    sy = b;
  }
  const size_t plane = blockIdx.x;
  const float2* up = U + plane * (size_t)my * mzh;
  for (int i = threadIdx.x; i < my * mzh; i += ZY_THREADS) Us[i] = __ldg(up + i);
  __syncthreads();
  for (int it = threadIdx.x; it < Ny * mzh; it += ZY_THREADS) {
    const int y = it / mzh, kz = it - y * mzh;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int ky = 0; ky < my; ++ky) cfma(acc, sy[(size_t)ky * Ny + y], Us[ky * mzh + kz]);
    Vs[it] = acc;
  }
  __syncthreads();
  const size_t pbase = plane * (size_t)Ny * Nz;
  if (VEC4) {
    const int nz4 = Nz >> 2;
    const int nyp = (Ny + 1) >> 1;
    for (int it = threadIdx.x; it < nyp * nz4; it += ZY_THREADS) {
      const int yp = it / nz4, zq = it - yp * nz4;
      const int y0 = 2 * yp, y1 = min(y0 + 1, Ny - 1), z0 = 4 * zq;
      float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int kz = 0; kz < mzh; ++kz) {
        const float2 v0 = Vs[(size_t)y0 * mzh + kz];
        const float2 v1 = Vs[(size_t)y1 * mzh + kz];
        const float4* t4 = reinterpret_cast<const float4*>(sz + (size_t)kz * Nz + z0);
        const float4 ta = t4[0], tb = t4[1];
        const float2 t[4] = {make_float2(ta.x, ta.y), make_float2(ta.z, ta.w), make_float2(tb.x, tb.y),
                             make_float2(tb.z, tb.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          a0[j] = fmaf(v0.x, t[j].x, a0[j]);
          a0[j] = fmaf(-v0.y, t[j].y, a0[j]);
          a1[j] = fmaf(v1.x, t[j].x, a1[j]);
          a1[j] = fmaf(-v1.y, t[j].y, a1[j]);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int y = y0 + r;
        if (y >= Ny) break;
        const float* a = r ? a1 : a0;
        const size_t o = pbase + (size_t)y * Nz + z0;
        float4 v = make_float4(a[0], a[1], a[2], a[3]);
        if (add0) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(add0 + o));
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        if (add1) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(add1 + o));
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        *reinterpret_cast<float4*>(out + o) = v;
      }
    }
  } else {
    for (int it = threadIdx.x; it < Ny * Nz; it += ZY_THREADS) {
      const int y = it / Nz, z = it - y * Nz;
      float a = 0.f;
      for (int kz = 0; kz < mzh; ++kz) {
        const float2 v = Vs[(size_t)y * mzh + kz];
        const float2 t = sz[(size_t)kz * Nz + z];
        a = fmaf(v.x, t.x, a);
        a = fmaf(-v.y, t.y, a);
      }
      const size_t o = pbase + it;
      if (add0) a += __ldg(add0 + o);
      if (add1) a += __ldg(add1 + o);
      out[o] = a;
    }
  }
}

// ------------------------------- host-side launchers ------------------------------------------
static int pick_group(int mzh) {
  // kz values per z-stage work item: the group size in {2,3,4} with the least padding (ties -> larger)
  int best = 4, waste = (4 - mzh % 4) % 4;
  for (int g = 3; g >= 2; --g) {
    const int w = (g - mzh % g) % g;
    if (w < waste) { waste = w; best = g; }
  }
  return best;
}

static int launch_xstage(bool scaled, const float2* in, float2* out, const float2* tw, int K, int R, int P,
                         const float* mode_scale, const float* row_scale, long long rs_stride, long long BC,
                         size_t smem_optin, cudaStream_t st) {
  if (BC > 65535) return fail(PDEPHI_ERR_UNSUPPORTED, "x stage: batch*channels > 65535");
  const size_t smem = (size_t)K * XT * XW * sizeof(float2);
  if (smem > smem_optin) return fail(PDEPHI_ERR_UNSUPPORTED, "x stage: %d contracted rows exceed shared memory", K);
  dim3 grid((P + 63) / 64, (unsigned)BC, (R + XT * XW - 1) / (XT * XW));
  const int rows_here = R < XT * XW ? R : XT * XW;
  dim3 block(32, (rows_here + XT - 1) / XT);
  if (scaled) {
    auto k = xstage_kernel<true>;
    if (smem > 48 * 1024) PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_SYNTHESIS_X, st);
    k<<<grid, block, smem, st>>>(in, out, tw, K, R, P, mode_scale, row_scale, rs_stride);
  } else {
    auto k = xstage_kernel<false>;
    if (smem > 48 * 1024) PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_ANALYSIS_X, st);
    k<<<grid, block, smem, st>>>(in, out, tw, K, R, P, nullptr, nullptr, 0);
  }
  PDEPHI_LAUNCH_CHECK("xstage_kernel");
  return PDEPHI_OK;
}

int xstage_analysis(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st) {
  return launch_xstage(false, ws, X, p->ax, p->Nx, p->mx, p->my * p->mzh, nullptr, nullptr, 0, BC, p->smem_optin, st);
}

int xstage_synthesis(const Plan* p, const float2* Z, float2* ws, const float* mode_scale, const float* row_scale,
                     long long rs_stride, long long BC, cudaStream_t st) {
  return launch_xstage(true, Z, ws, p->sx, p->mx, p->Nx, p->my * p->mzh, mode_scale, row_scale, rs_stride, BC,
                       p->smem_optin, st);
}

int analysis_simt(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws,
                  cudaStream_t st, int stages, float2* T) {
  if (stages != ST_ALL) ws = T;                 // the exchange tensor takes the place of the intermediate
  if (!(stages & ST_ZY)) return xstage_analysis(p, ws, X, BC, st);
  const int Ny = p->Ny, Nz = p->Nz, my = p->my, mzh = p->mzh;
  const bool vec4 = (Nz % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int ldx = vec4 ? Nz + 4 : Nz + 1;
  const size_t t1_floats = ((2 * (size_t)Ny * mzh + 3) / 4) * 4;
  // chunk of y rows staged at a time: sized so that ~4 CTAs share an SM, else as many rows as fit
  const size_t budget = 56 * 1024;
  if (t1_floats * 4 + 32 * (size_t)ldx * 4 > p->smem_optin)
    return fail(PDEPHI_ERR_UNSUPPORTED, "analysis: plane %dx%d with %d kz modes does not fit shared memory", Ny, Nz, mzh);
  int YC = Ny;
  while (YC > 32 && t1_floats * 4 + (size_t)YC * ldx * 4 > budget) YC = (YC + 1) / 2;
  YC = min(Ny, (YC + 31) / 32 * 32);
  while (YC > 1 && t1_floats * 4 + (size_t)YC * ldx * 4 > p->smem_optin) YC -= 1;
  size_t smem = t1_floats * 4 + (size_t)YC * ldx * 4;
  const size_t tw_bytes = 8 + ((size_t)Nz * mzh + (size_t)Ny * my) * sizeof(float2);
  const int tw_smem = (smem + tw_bytes <= 110 * 1024) ? 1 : 0;      // keep two CTAs per SM
  if (tw_smem) smem += tw_bytes;
  const size_t planes = (size_t)BC * p->Nx;
  if (planes > 0x7fffffffULL) return fail(PDEPHI_ERR_UNSUPPORTED, "analysis: too many planes");
  const int G = pick_group(mzh);
  const float2* az = p->az[direction];
#define LAUNCH_A(GG, VV)                                                                                   \
  do {                                                                                                     \
    auto k = analysis_zy_kernel<GG, VV>;                                                                   \
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    ProfScope ps(K_ANALYSIS_ZY, st);                                                                       \
    k<<<(unsigned)planes, ZY_THREADS, smem, st>>>(x, ws, az, p->ay, Ny, Nz, my, mzh, YC, ldx, tw_smem);    \
  } while (0)
  if (vec4) {
    if (G == 2) LAUNCH_A(2, true); else if (G == 3) LAUNCH_A(3, true); else LAUNCH_A(4, true);
  } else {
    if (G == 2) LAUNCH_A(2, false); else if (G == 3) LAUNCH_A(3, false); else LAUNCH_A(4, false);
  }
#undef LAUNCH_A
  PDEPHI_LAUNCH_CHECK("analysis_zy_kernel");
  if (!(stages & ST_X)) return PDEPHI_OK;
  return xstage_analysis(p, ws, X, BC, st);
}

int synthesis_simt(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                   const float* mode_scale, const float* row_scale, long long rs_stride, long long BC,
                   int direction, float2* ws, cudaStream_t st, int stages, float2* T) {
  const int Ny = p->Ny, Nz = p->Nz, my = p->my, mzh = p->mzh;
  if (stages != ST_ALL) ws = T;
  if (stages & ST_X) {
    int rc = xstage_synthesis(p, Z, ws, mode_scale, row_scale, rs_stride, BC, st);
    if (rc) return rc;
  }
  if (!(stages & ST_ZY)) return PDEPHI_OK;
  size_t smem = ((size_t)my * mzh + (size_t)Ny * mzh) * sizeof(float2);
  if (smem > p->smem_optin)
    return fail(PDEPHI_ERR_UNSUPPORTED, "synthesis: %d y rows x %d kz modes do not fit shared memory", Ny, mzh);
  const size_t tw_bytes = 32 + ((size_t)mzh * Nz + (size_t)my * Ny) * sizeof(float2);
  const int tw_smem = (smem + tw_bytes <= 110 * 1024) ? 1 : 0;
  if (tw_smem) smem += tw_bytes;
  const size_t planes = (size_t)BC * p->Nx;
  if (planes > 0x7fffffffULL) return fail(PDEPHI_ERR_UNSUPPORTED, "synthesis: too many planes");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec4 = (Nz % 4 == 0) && al16(out) && al16(add0) && al16(add1);
  const float2* sz = p->sz[direction];
  if (vec4) {
    auto k = synthesis_zy_kernel<true>;
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_SYNTHESIS_ZY, st);
    k<<<(unsigned)planes, ZY_THREADS, smem, st>>>(ws, out, add0, add1, p->sy, sz, Ny, Nz, my, mzh, tw_smem);
  } else {
    auto k = synthesis_zy_kernel<false>;
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_SYNTHESIS_ZY, st);
    k<<<(unsigned)planes, ZY_THREADS, smem, st>>>(ws, out, add0, add1, p->sy, sz, Ny, Nz, my, mzh, tw_smem);
  }
  PDEPHI_LAUNCH_CHECK("synthesis_zy_kernel");
  return PDEPHI_OK;
}

}  // namespace pdephi
