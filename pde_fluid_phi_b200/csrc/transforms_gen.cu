// K1 / K3, TF32 path for the shapes the fused (z,y) kernels of transforms_tc.cu do not cover (grids up to 256 per
// axis, up to 64 retained modes per axis, 2-D problems [Nx, Ny, 1]): the pruned transforms as ONE tcgen05
// contraction per axis, every stage a persistent warp-specialised kernel with TMA-staged data operands.
//
//   analysis    x [rows][zN] --(z)--> W1 [rows][ldz] --(y)--> W2 [BC*Nx][my][mzh] --(x)--> X [BC][mx][my][mzh]
//   synthesis   Z --pad/scale--> Zp [BC][mx][ym][ldz/2] --(x)--> U [BC*Nx][ym*ldz] --(y)--> V [rows][ldz] --(z)--> out
//
//   z stages   the contracted index is contiguous: plain K-major GEMMs over 128-row tiles,
//              analysis   D[128 x NB] = tile[128 x zN] . Bz[NB x zN]^T          (columns 2k = re, 2k+1 = im)
//              synthesis  D[128 x zN] = tile[128 x ldz] . Bs[zN x ldz]^T  (+ addends in the epilogue)
//   y/x stages the contracted index is the ROW index of a row-major [N rows][cols] matrix per batch item, so a TMA box
//              [N rows][32 floats] is an MN-major B operand (128-byte swizzle with 32-byte atoms):
//              analysis   D[128 x 64] = A[(cos k | sin k) x N] . tile[N x 64],  A TMEM-resident, combined through smem
//              synthesis  D1 = C[128 x mk] . tile[mk x 64],  D2 = S . tile,  combined in registers
//
// A 2-D problem (Nz = 1) is the same pipeline with the y axis in the role of the contiguous axis and no middle stage.
// The intermediate row length ldz = 2*zm rounded up to 4 floats keeps every TMA row stride a multiple of 16 bytes.
// Replaces torch.fft.rfftn / irfftn of rational_fourier.py:161,170 and spectral_layers.py:59,76 at those shapes.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pdephi {
namespace gen {
using namespace tcptx;

constexpr int KBYTES = 16384;   // one [128 rows][32 floats] k-block
// The tensor core truncates the raw fp32 data operand of every stage to TF32 (a mean relative bias of -2^-11).  Each
// stage undoes it on its fp32 accumulator in the epilogue.  Folding the factor into the constant operand (as
// transforms_tc.cu does) is only right on average: twiddles that are exactly 1.0 -- the whole DC mode, i.e. every mean
// and every bias gradient -- round 1 + 2^-11 up to 1 + 2^-10 in TF32 and over-compensate by 4.9e-4 per stage.
#define PDEPHI_GEN_COMP 1.00048828125f
constexpr int ST_LD = 36;       // padded staging row (32 floats + 4)

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {   // MN-major TF32 operand, 128B swizzle / 32B atoms
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// constant operand images (plan creation), generated in fp64 and rounded once to TF32
//   kind 0  z analysis  B  K-major [R = NB rows][K = zN]   row r: mode r>>1, (r&1 ? -w sin : w cos)(2 pi mode k / N)
//   kind 1  z synthesis B  K-major [R = zN rows][K]        col k: mode k>>1, (k&1 ? -w sin : w cos)(2 pi mode r / N)
//   kind 2  axis analysis A  plain [128][N]                row r: mode r&63, (r>>6 ? sin : cos)(2 pi mode (k+off) / Nt)
//   kind 3/4 axis synthesis A K-major [R = N rows][K = mk]  (cos / sin)(2 pi k (r+off) / Nt)
//   kind 5  z analysis as a plain [R][K = zN] A operand (TMEM-resident, fused block backward): rows as kind 0; the LAST row
//           is all ones (the tile sums of gu = the bias gradient of the block's convolution ride in the same product)
// wmode 0: w = 1, 1: w = c_k / total (c_k = 1 for k = 0 and Nyquist, else 2), 2: w = 1 / total
// ------------------------------------------------------------------------------------------------
__global__ void gen_image_kernel(float* __restrict__ img, int kind, int R, int K, int Nt, int off, int m, int wmode, int nyq,
                                 double inv_total, double comp, int part, int c0, int nc) {
  // part 0: the TF32 (round-to-nearest) value of the entry; 1: its lo part, rna_tf32(v - hi).
  // nc > 0 (kinds 0 and 1): the STACKED image of a column block of the split-TF32 z kernels -- rows [0, nc) hold the hi
  // parts of columns c0 .. c0+nc-1 (kind 0: output columns; kind 1: output positions), rows [nc, 2 nc) their lo parts.
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * K) return;
  int r = idx / K;
  const int k = idx % K;
  const int rimg = r;
  if (nc > 0) { part = r >= nc ? 1 : 0; r = c0 + (r >= nc ? r - nc : r); }
  int mode = -1, pos = 0, sine = 0;
  if (kind == 0 || kind == 5) { mode = r >> 1; sine = r & 1; pos = k; }
  if (kind == 1) { mode = k >> 1; sine = k & 1; pos = r; }
  if (kind == 2) { mode = r & 63; sine = r >> 6; pos = k + off; }
  if (kind == 3 || kind == 4) { mode = k; sine = (kind == 4); pos = r + off; }
  double v = 0.0;
  if (mode < m && (kind != 1 || nc == 0 || r < Nt)) {
    double s, c;
    sincospi(2.0 * (double)(((long long)mode * pos) % Nt) / (double)Nt, &s, &c);
    double w = 1.0;
    if (wmode == 1) w = ((mode == 0 || mode == nyq) ? 1.0 : 2.0) * inv_total;
    if (wmode == 2) w = inv_total;
    v = (kind <= 1 || kind == 5) ? (sine ? -w * s : w * c) : (sine ? s : c);
    v *= comp;
  }
  if (kind == 5 && rimg == R - 1) v = 1.0;
  float f = to_tf32((float)v);
  if (part == 1) f = to_tf32((float)(v - (double)f));
  if (kind == 2 || kind == 5) img[(size_t)rimg * K + k] = f;
  else img[swz_off(R, rimg, k) / 4] = f;
}

// ------------------------------------------------------------------------------------------------
// z analysis: W[row][0..ldz) = sum_z x[row][z] . Bz[c][z]
//   warp 0 TMA producer (k-block ring), warp 1 MMA issuer, warps 2..5 epilogue (TMEM quarter = warp % 4)
// ------------------------------------------------------------------------------------------------
constexpr int ZA_THREADS = 192;
constexpr int MAX_RING = 8;
struct ZASmem { uint32_t ring, bimg, stg, bars, tmem_slot, total; int nring; };
__host__ __device__ inline ZASmem za_smem(int NKB, int NBm, int nring) {
  ZASmem s; uint32_t o = 0;
  s.nring = nring;
  s.ring = o; o += (uint32_t)nring * KBYTES;
  s.bimg = o; o += (uint32_t)NKB * NBm * 128;
  s.stg = o; o += 4 * 32 * ST_LD * 4;
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { ZB_FULL = 0, ZB_EMPTY = MAX_RING, ZB_DFULL = 2 * MAX_RING, ZB_DEMPTY = 2 * MAX_RING + 2 };

__global__ void __launch_bounds__(ZA_THREADS, 1)
gen_z_analysis_kernel(const __grid_constant__ CUtensorMap tmap_x, float* __restrict__ W, const float* __restrict__ Bimg,
                      long long rows, int ntiles, int NKB, int NBm, int ldz, int nring) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const ZASmem L = za_smem(NKB, NBm, nring);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float4* s1 = reinterpret_cast<const float4*>(Bimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.bimg);
    for (int i = threadIdx.x; i < NKB * NBm * 128 / 16; i += ZA_THREADS) d1[i] = __ldg(s1 + i);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RING; ++i) { mbar_init(bar(ZB_FULL + i), 1); mbar_init(bar(ZB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(ZB_DFULL + i), 1); mbar_init(bar(ZB_DEMPTY + i), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 1;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
        for (int kb = 0; kb < NKB; ++kb) {
          mbar_wait(bar(ZB_EMPTY + s), ph);
          mbar_arrive_expect_tx(bar(ZB_FULL + s), KBYTES);
          tma_load_2d(base + L.ring + s * KBYTES, &tmap_x, bar(ZB_FULL + s), kb * 32, t * 128);
          if (++s == nring) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t id = idesc(128, NBm, 0, 0);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int d = it & 1;
        mbar_wait(bar(ZB_DEMPTY + d), ((it >> 1) & 1) ^ 1);
        for (int kb = 0; kb < NKB; ++kb) {
          mbar_wait(bar(ZB_FULL + s), ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ss(tmem + d * 128, umma_desc(base + L.ring + s * KBYTES + k * 32, 1024),
                   umma_desc(base + L.bimg + kb * NBm * 128 + k * 32, 1024), id, (kb | k) ? 1u : 0u);
          mma_commit(bar(ZB_EMPTY + s));
          if (++s == nring) { s = 0; ph ^= 1; }
        }
        mma_commit(bar(ZB_DFULL + d));
      }
    }
  } else {
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stg) + (warp - 2) * 32 * ST_LD;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int d = it & 1;
      mbar_wait(bar(ZB_DFULL + d), (it >> 1) & 1);
      tc_fence_after();
      float* wrow = W + ((size_t)t * 128 + 32 * q) * ldz;
      const int nvalid = (int)min(32LL, rows - ((long long)t * 128 + 32 * q));   // the last tile may be partial
      for (int c0 = 0; c0 < ldz; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + d * 128 + lane_base + c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= ldz) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(ZB_DEMPTY + d));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + lane * ST_LD + 4 * j) =
              make_float4(v[4 * j] * PDEPHI_GEN_COMP, v[4 * j + 1] * PDEPHI_GEN_COMP, v[4 * j + 2] * PDEPHI_GEN_COMP,
                          v[4 * j + 3] * PDEPHI_GEN_COMP);
        __syncwarp();
        const int lpr = min(32, ldz - c0) >> 2;            // float4 lanes per row of this chunk
        for (int idx = lane; idx < 32 * lpr; idx += 32) {
          const int row = idx / lpr, c4 = idx - row * lpr;
          if (row < nvalid)
            *reinterpret_cast<float4*>(wrow + (size_t)row * ldz + c0 + 4 * c4) =
                *reinterpret_cast<const float4*>(st + row * ST_LD + 4 * c4);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// z synthesis: out[row][z] = sum_c V[row][c] . Bs[z][c]  (+ addends)
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue: two groups of four on alternate tiles
// ------------------------------------------------------------------------------------------------
constexpr int ZS_EPI = 8;
constexpr int ZS_THREADS = 32 * (2 + ZS_EPI);
struct ZSSmem { uint32_t ring, bimg, stg, bars, tmem_slot, total; int nring; };
__host__ __device__ inline ZSSmem zs_smem(int zN, int NKBK, int nring) {
  ZSSmem s; uint32_t o = 0;
  s.nring = nring;
  s.ring = o; o += (uint32_t)nring * KBYTES;
  s.bimg = o; o += (uint32_t)NKBK * zN * 128;
  s.stg = o; o += ZS_EPI * 32 * ST_LD * 4;
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}

template <int NADD>
__global__ void __launch_bounds__(ZS_THREADS, 1)
gen_z_synthesis_kernel(const __grid_constant__ CUtensorMap tmap_v, float* __restrict__ out, const float* __restrict__ add0,
                       const float* __restrict__ add1, const float* __restrict__ Bimg, long long rows, int ntiles, int zN,
                       int NKBK, int ksteps, int nring, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const ZSSmem L = zs_smem(zN, NKBK, nring);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float4* s1 = reinterpret_cast<const float4*>(Bimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.bimg);
    for (int i = threadIdx.x; i < NKBK * zN * 128 / 16; i += ZS_THREADS) d1[i] = __ldg(s1 + i);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RING; ++i) { mbar_init(bar(ZB_FULL + i), 1); mbar_init(bar(ZB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(ZB_DFULL + i), 1); mbar_init(bar(ZB_DEMPTY + i), ZS_EPI / 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, (uint32_t)tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 1;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
        for (int kb = 0; kb < NKBK; ++kb) {
          mbar_wait(bar(ZB_EMPTY + s), ph);
          mbar_arrive_expect_tx(bar(ZB_FULL + s), KBYTES);
          tma_load_2d(base + L.ring + s * KBYTES, &tmap_v, bar(ZB_FULL + s), kb * 32, t * 128);
          if (++s == nring) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t id = idesc(128, zN, 0, 0);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int d = it & 1;
        mbar_wait(bar(ZB_DEMPTY + d), ((it >> 1) & 1) ^ 1);
        for (int kb = 0; kb < NKBK; ++kb) {
          mbar_wait(bar(ZB_FULL + s), ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (kb * 4 + k < ksteps)
              mma_ss(tmem + d * zN, umma_desc(base + L.ring + s * KBYTES + k * 32, 1024),
                     umma_desc(base + L.bimg + kb * zN * 128 + k * 32, 1024), id, (kb | k) ? 1u : 0u);
          mma_commit(bar(ZB_EMPTY + s));
          if (++s == nring) { s = 0; ph ^= 1; }
        }
        mma_commit(bar(ZB_DFULL + d));
      }
    }
  } else {
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stg) + (warp - 2) * 32 * ST_LD;
    const int rsub = lane >> 3, c4 = lane & 7;
    const int n_my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nrounds = zN >> 5;
    for (int it = grp; it < n_my; it += 2) {
      const int t = blockIdx.x + it * gridDim.x;
      mbar_wait(bar(ZB_DFULL + grp), (it >> 1) & 1);
      tc_fence_after();
      const int nvalid = (int)min(32LL, rows - ((long long)t * 128 + 32 * q)) - rsub;   // rows 4j + rsub < valid count
      for (int r = 0; r < nrounds; ++r) {
        const size_t gbase = ((size_t)t * 128 + 32 * q + rsub) * zN + r * 32 + 4 * c4;
        float4 a0v[NADD >= 1 ? 8 : 1], a1v[NADD == 2 ? 8 : 1];
        if (NADD >= 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < nvalid) a0v[j] = __ldcs(reinterpret_cast<const float4*>(add0 + gbase + (size_t)(4 * j) * zN));
        }
        if (NADD == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < nvalid) a1v[j] = __ldcs(reinterpret_cast<const float4*>(add1 + gbase + (size_t)(4 * j) * zN));
        }
        float v[32];
        tmem_ld32(tmem + grp * zN + lane_base + r * 32, v);
        tmem_ld_wait();
        if (r == nrounds - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(ZB_DEMPTY + grp));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + lane * ST_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = *reinterpret_cast<const float4*>(st + (4 * j + rsub) * ST_LD + 4 * c4);
          if (NADD >= 1) {
            o.x = fmaf(o.x, PDEPHI_GEN_COMP, a0v[j].x); o.y = fmaf(o.y, PDEPHI_GEN_COMP, a0v[j].y);
            o.z = fmaf(o.z, PDEPHI_GEN_COMP, a0v[j].z); o.w = fmaf(o.w, PDEPHI_GEN_COMP, a0v[j].w);
          } else {
            o.x *= PDEPHI_GEN_COMP; o.y *= PDEPHI_GEN_COMP; o.z *= PDEPHI_GEN_COMP; o.w *= PDEPHI_GEN_COMP;
          }
          if (NADD == 2) { o.x += a1v[j].x; o.y += a1v[j].y; o.z += a1v[j].z; o.w += a1v[j].w; }
          if (4 * j < nvalid) __stcs(reinterpret_cast<float4*>(out + gbase + (size_t)(4 * j) * zN), o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// split-TF32 (fp32-grade) z stages.  A.B ~ Ahi.Bhi + (Ahi.Blo + Alo.Bhi): the raw fp32 data operand IS its hi part (the
// tensor core drops the low 13 mantissa bits), lo = rna_tf32(v - hi).  The constant operand comes as a STACKED image --
// hi rows then lo rows of one column block -- so one MMA of twice the width yields data_hi.Bhi and data_hi.Blo side by
// side in TMEM; a second MMA adds data_lo.Bhi to the small half.  Large and small partial products keep separate
// accumulators and meet in fp32 in the epilogue (the tensor core truncates after every accumulation step: one shared
// accumulator leaves a systematic 2e-6 shrink, transforms_tc3.cu).  Where the stacked image does not fit beside the data
// ring (256-point axes, 64 retained modes) the z analysis splits the CONTRACTED index over launches -- launch j reads only
// its k-blocks of x, so the activation still crosses HBM once, and accumulates into the small W -- and the z synthesis
// splits the output positions (each launch re-reads the small V).
//
// z analysis: W[row][0 .. nc) (+)= sum_{z in this launch's k-blocks} x[row][z] . Bz[c][z];  Wlo = lo part of the final W
// (data operand of the next stage)
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue, warps 6..9 splitter (landed k-block -> lo k-block)
// ------------------------------------------------------------------------------------------------
constexpr int ZA3_THREADS = 320;
constexpr int ZA3_MAXLO = 4;
struct ZA3Smem { uint32_t ring, lo, bimg, stg, bars, tmem_slot, total; };
__host__ __device__ inline ZA3Smem za3_smem(int NKB, int nc, int nring, int nlo) {
  ZA3Smem s; uint32_t o = 0;
  s.ring = o; o += (uint32_t)nring * KBYTES;
  s.lo = o; o += (uint32_t)nlo * KBYTES;
  s.bimg = o; o += (uint32_t)NKB * 2 * nc * 128;
  s.stg = o; o += 4 * 32 * ST_LD * 4;
  s.bars = o; o += 512;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { Z3_FULL = 0, Z3_EMPTY = MAX_RING, Z3_LFULL = 2 * MAX_RING, Z3_LEMPTY = 2 * MAX_RING + ZA3_MAXLO,
       Z3_DFULL = 2 * MAX_RING + 2 * ZA3_MAXLO, Z3_DEMPTY = 2 * MAX_RING + 2 * ZA3_MAXLO + 2 };

__device__ __forceinline__ float split_lo(float v) { return to_tf32(v - tf32_trunc(v)); }

__global__ void __launch_bounds__(ZA3_THREADS, 1)
gen_z_analysis_x3_kernel(const __grid_constant__ CUtensorMap tmap_x, float* __restrict__ W, float* __restrict__ Wlo,
                         const float* __restrict__ Bimg, long long rows, int ntiles, int NKB, int nc, int ldz, int kb0, int accumulate,
                         int nring, int nlo, int tmem_cols) {
  // NKB k-blocks starting at kb0 (Bimg points at their slice of the stacked image); accumulate: W += (an earlier launch
  // covered other k-blocks); Wlo non-null on the last launch only
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const ZA3Smem L = za3_smem(NKB, nc, nring, nlo);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float4* s1 = reinterpret_cast<const float4*>(Bimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.bimg);
    for (int i = threadIdx.x; i < NKB * 2 * nc * 128 / 16; i += ZA3_THREADS) d1[i] = __ldg(s1 + i);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RING; ++i) { mbar_init(bar(Z3_FULL + i), 1); mbar_init(bar(Z3_EMPTY + i), 1); }
    for (int i = 0; i < ZA3_MAXLO; ++i) { mbar_init(bar(Z3_LFULL + i), 4); mbar_init(bar(Z3_LEMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(Z3_DFULL + i), 1); mbar_init(bar(Z3_DEMPTY + i), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, (uint32_t)tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const int DW = 2 * nc;                       // TMEM columns of one accumulator pair: [large | small]

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 1;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
        for (int kb = 0; kb < NKB; ++kb) {
          mbar_wait(bar(Z3_EMPTY + s), ph);
          mbar_arrive_expect_tx(bar(Z3_FULL + s), KBYTES);
          tma_load_2d(base + L.ring + s * KBYTES, &tmap_x, bar(Z3_FULL + s), (kb0 + kb) * 32, t * 128);
          if (++s == nring) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t id2 = idesc(128, 2 * nc, 0, 0), id1 = idesc(128, nc, 0, 0);
      int s = 0, l = 0; uint32_t ph = 0, lph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int d = it & 1;
        mbar_wait(bar(Z3_DEMPTY + d), ((it >> 1) & 1) ^ 1);
        for (int kb = 0; kb < NKB; ++kb) {
          mbar_wait(bar(Z3_FULL + s), ph);
          mbar_wait(bar(Z3_LFULL + l), lph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = umma_desc(base + L.bimg + kb * 2 * nc * 128 + k * 32, 1024);
            mma_ss(tmem + d * DW, umma_desc(base + L.ring + s * KBYTES + k * 32, 1024), bd, id2, (kb | k) ? 1u : 0u);
            mma_ss(tmem + d * DW + nc, umma_desc(base + L.lo + l * KBYTES + k * 32, 1024), bd, id1, 1u);
          }
          mma_commit(bar(Z3_EMPTY + s));
          mma_commit(bar(Z3_LEMPTY + l));
          if (++s == nring) { s = 0; ph ^= 1; }
          if (++l == nlo) { l = 0; lph ^= 1; }
        }
        mma_commit(bar(Z3_DFULL + d));
      }
    }
  } else if (warp >= 6) {
    // splitter: the landed k-block (raw fp32 = hi + lo) -> its lo part at the same (swizzled) offsets of a lo slot
    const int tid = threadIdx.x - 6 * 32;
    int s = 0, l = 0; uint32_t ph = 0, lph = 1;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
      for (int kb = 0; kb < NKB; ++kb) {
        mbar_wait(bar(Z3_FULL + s), ph);
        mbar_wait(bar(Z3_LEMPTY + l), lph);
        const float4* src = reinterpret_cast<const float4*>(sm + L.ring + s * KBYTES);
        float4* dst = reinterpret_cast<float4*>(sm + L.lo + l * KBYTES);
#pragma unroll
        for (int j = 0; j < KBYTES / 16 / 128; ++j) {
          float4 v = src[tid + 128 * j];
          v.x = split_lo(v.x); v.y = split_lo(v.y); v.z = split_lo(v.z); v.w = split_lo(v.w);
          dst[tid + 128 * j] = v;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Z3_LFULL + l));
        if (++s == nring) { s = 0; ph ^= 1; }
        if (++l == nlo) { l = 0; lph ^= 1; }
      }
  } else {
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stg) + (warp - 2) * 32 * ST_LD;
    const int ncv = min(nc, ldz);              // columns that exist in W
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int d = it & 1;
      const size_t rbase = ((size_t)t * 128 + 32 * q) * ldz;
      const int nvalid = (int)min(32LL, rows - ((long long)t * 128 + 32 * q));   // the last tile may be partial
      for (int cc = 0; cc < nc; cc += 32) {
        const int lpr = max(0, min(32, ncv - cc)) >> 2;            // float4 lanes per row of this chunk
        float4 prev[8];          // accumulating launch: the earlier partial sums, fetched before the TMEM read
        if (accumulate) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = lane + 32 * i, row = idx / max(lpr, 1), c4 = idx - row * lpr;
            prev[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < lpr && row < nvalid) prev[i] = __ldcs(reinterpret_cast<const float4*>(W + rbase + (size_t)row * ldz + cc + 4 * c4));
          }
        }
        if (cc == 0) {
          mbar_wait(bar(Z3_DFULL + d), (it >> 1) & 1);
          tc_fence_after();
        }
        float v[32], w[32];
        if (cc + 32 <= nc) {
          tmem_ld32(tmem + d * DW + lane_base + cc, v);
          tmem_ld32(tmem + d * DW + nc + lane_base + cc, w);
        } else {                                 // nc is a multiple of 16: a 16-column tail
          tmem_ld16(tmem + d * DW + lane_base + cc, v);
          tmem_ld16(tmem + d * DW + nc + lane_base + cc, w);
#pragma unroll
          for (int j = 16; j < 32; ++j) { v[j] = 0.f; w[j] = 0.f; }
        }
        tmem_ld_wait();
        if (cc + 32 >= nc) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(Z3_DEMPTY + d));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + lane * ST_LD + 4 * j) =
              make_float4(v[4 * j] + w[4 * j], v[4 * j + 1] + w[4 * j + 1], v[4 * j + 2] + w[4 * j + 2], v[4 * j + 3] + w[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = lane + 32 * i, row = idx / max(lpr, 1), c4 = idx - row * lpr;
          if (i < lpr && row < nvalid) {
            float4 o = *reinterpret_cast<const float4*>(st + row * ST_LD + 4 * c4);
            if (accumulate) { o.x += prev[i].x; o.y += prev[i].y; o.z += prev[i].z; o.w += prev[i].w; }
            *reinterpret_cast<float4*>(W + rbase + (size_t)row * ldz + cc + 4 * c4) = o;
            if (Wlo)
              *reinterpret_cast<float4*>(Wlo + rbase + (size_t)row * ldz + cc + 4 * c4) =
                  make_float4(split_lo(o.x), split_lo(o.y), split_lo(o.z), split_lo(o.w));
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// split-TF32 z synthesis: out[row][c0 .. c0+nc) = sum_c V[row][c] . Bs[z][c]  (+ addends), V = Vhi (raw) + Vlo (its own array,
// written by the producing stage): two TMA streams, no splitter.  warps as in gen_z_synthesis_kernel.
// ------------------------------------------------------------------------------------------------
struct ZS3Smem { uint32_t ring, bimg, stg, bars, tmem_slot, total; };
__host__ __device__ inline ZS3Smem zs3_smem(int nc, int NKBK, int nring) {
  ZS3Smem s; uint32_t o = 0;
  s.ring = o; o += (uint32_t)nring * 2 * KBYTES;           // a slot = the k-block of V and the same k-block of Vlo
  s.bimg = o; o += (uint32_t)NKBK * 2 * nc * 128;
  s.stg = o; o += ZS_EPI * 32 * ST_LD * 4;
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}

template <int NADD>
__global__ void __launch_bounds__(ZS_THREADS, 1)
gen_z_synthesis_x3_kernel(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_vlo,
                          float* __restrict__ out, const float* __restrict__ add0, const float* __restrict__ add1,
                          const float* __restrict__ Bimg, long long rows, int ntiles, int zN, int c0, int nc, int NKBK, int ksteps,
                          int nring, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const ZS3Smem L = zs3_smem(nc, NKBK, nring);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float4* s1 = reinterpret_cast<const float4*>(Bimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.bimg);
    for (int i = threadIdx.x; i < NKBK * 2 * nc * 128 / 16; i += ZS_THREADS) d1[i] = __ldg(s1 + i);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RING; ++i) { mbar_init(bar(ZB_FULL + i), 1); mbar_init(bar(ZB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(ZB_DFULL + i), 1); mbar_init(bar(ZB_DEMPTY + i), ZS_EPI / 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, (uint32_t)tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const int DW = 2 * nc;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 1;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
        for (int kb = 0; kb < NKBK; ++kb) {
          mbar_wait(bar(ZB_EMPTY + s), ph);
          mbar_arrive_expect_tx(bar(ZB_FULL + s), 2 * KBYTES);
          tma_load_2d(base + L.ring + s * 2 * KBYTES, &tmap_v, bar(ZB_FULL + s), kb * 32, t * 128);
          tma_load_2d(base + L.ring + s * 2 * KBYTES + KBYTES, &tmap_vlo, bar(ZB_FULL + s), kb * 32, t * 128);
          if (++s == nring) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t id2 = idesc(128, 2 * nc, 0, 0), id1 = idesc(128, nc, 0, 0);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int d = it & 1;
        mbar_wait(bar(ZB_DEMPTY + d), ((it >> 1) & 1) ^ 1);
        for (int kb = 0; kb < NKBK; ++kb) {
          mbar_wait(bar(ZB_FULL + s), ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (kb * 4 + k < ksteps) {
              const uint64_t bd = umma_desc(base + L.bimg + kb * 2 * nc * 128 + k * 32, 1024);
              mma_ss(tmem + d * DW, umma_desc(base + L.ring + s * 2 * KBYTES + k * 32, 1024), bd, id2, (kb | k) ? 1u : 0u);
              mma_ss(tmem + d * DW + nc, umma_desc(base + L.ring + s * 2 * KBYTES + KBYTES + k * 32, 1024), bd, id1, 1u);
            }
          mma_commit(bar(ZB_EMPTY + s));
          if (++s == nring) { s = 0; ph ^= 1; }
        }
        mma_commit(bar(ZB_DFULL + d));
      }
    }
  } else {
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stg) + (warp - 2) * 32 * ST_LD;
    const int rsub = lane >> 3, c4 = lane & 7;
    const int n_my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nrounds = nc >> 5;               // nc is a multiple of 32 here
    for (int it = grp; it < n_my; it += 2) {
      const int t = blockIdx.x + it * gridDim.x;
      mbar_wait(bar(ZB_DFULL + grp), (it >> 1) & 1);
      tc_fence_after();
      const int nvalid = (int)min(32LL, rows - ((long long)t * 128 + 32 * q)) - rsub;   // rows 4j + rsub < valid count
      for (int r = 0; r < nrounds; ++r) {
        const size_t gbase = ((size_t)t * 128 + 32 * q + rsub) * zN + c0 + r * 32 + 4 * c4;
        float4 a0v[NADD >= 1 ? 8 : 1], a1v[NADD == 2 ? 8 : 1];
        if (NADD >= 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < nvalid) a0v[j] = __ldcs(reinterpret_cast<const float4*>(add0 + gbase + (size_t)(4 * j) * zN));
        }
        if (NADD == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < nvalid) a1v[j] = __ldcs(reinterpret_cast<const float4*>(add1 + gbase + (size_t)(4 * j) * zN));
        }
        float v[32], w[32];
        tmem_ld32(tmem + grp * DW + lane_base + r * 32, v);
        tmem_ld32(tmem + grp * DW + nc + lane_base + r * 32, w);
        tmem_ld_wait();
        if (r == nrounds - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(ZB_DEMPTY + grp));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + lane * ST_LD + 4 * j) =
              make_float4(v[4 * j] + w[4 * j], v[4 * j + 1] + w[4 * j + 1], v[4 * j + 2] + w[4 * j + 2], v[4 * j + 3] + w[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = *reinterpret_cast<const float4*>(st + (4 * j + rsub) * ST_LD + 4 * c4);
          if (NADD >= 1) { o.x += a0v[j].x; o.y += a0v[j].y; o.z += a0v[j].z; o.w += a0v[j].w; }
          if (NADD == 2) { o.x += a1v[j].x; o.y += a1v[j].y; o.z += a1v[j].z; o.w += a1v[j].w; }
          if (4 * j < nvalid) __stcs(reinterpret_cast<float4*>(out + gbase + (size_t)(4 * j) * zN), o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// axis analysis (contracted index = row): out[b][k][j] = sum_n e^{-i theta(k,n)} in[b][n][j]
//   warp 0 TMA, warp 1 MMA, warps 2..5 epilogue.  A = [cos k (rows 0..63) | sin k (rows 64..127)] x N lives in TMEM.
// ------------------------------------------------------------------------------------------------
constexpr int AA_THREADS = 192;
constexpr int AA_MAXST = 4;
constexpr int AA_ELD = 66;
constexpr int NT = 64;          // floats per column tile (32 complex)
struct AASmem { uint32_t stage, e, bars, tmem_slot, total; int nst; };
__host__ __device__ inline AASmem aa_smem(int N, int nst) {
  AASmem s; uint32_t o = 0;
  s.nst = nst;
  s.stage = o; o += (uint32_t)nst * 2 * N * 128;
  s.e = o; o += 128 * AA_ELD * 4;
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { AB_FULL = 0, AB_EMPTY = AA_MAXST, AB_DFULL = 2 * AA_MAXST, AB_DEMPTY = 2 * AA_MAXST + 2 };

__global__ void __launch_bounds__(AA_THREADS, 1)
gen_axis_analysis_kernel(const __grid_constant__ CUtensorMap tmap_in, float2* __restrict__ out, const float* __restrict__ Aplain,
                         int N, int m, int Pout, int ntn, long long ntiles, int nst, float comp, int accumulate,
                         float2* __restrict__ out_lo) {
  // comp: epilogue factor (PDEPHI_GEN_COMP on the single-pass route, 1 on the split route).  accumulate: out += result -- the
  // split route runs a stage as three launches (data.Ahi, data_lo.Ahi, data.Alo) that meet in fp32 here.  out_lo: also
  // store the lo part of the final value (the data_lo operand of the next stage).
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const AASmem L = aa_smem(N, nst);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box = (uint32_t)N * 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < AA_MAXST; ++i) { mbar_init(bar(AB_FULL + i), 1); mbar_init(bar(AB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(AB_DFULL + i), 1); mbar_init(bar(AB_DEMPTY + i), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const uint32_t tmem_a = tmem, tmem_d = tmem + 384;     // A: N <= 256 columns, D: 2 x 64 columns
  if (warp >= 2) {
    const int q = warp & 3;
    const float* row = Aplain + (size_t)(32 * q + lane) * N;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(row + c0) + j);
        r[4 * j] = t.x; r[4 * j + 1] = t.y; r[4 * j + 2] = t.z; r[4 * j + 3] = t.w;
      }
      tmem_st32(tmem_a + ((uint32_t)(q * 32) << 16) + c0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 1;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(bar(AB_EMPTY + s), ph);
        mbar_arrive_expect_tx(bar(AB_FULL + s), 2 * box);
        const long long b = t / ntn;
        const int n0 = (int)(t - b * ntn) * NT;
        tma_load_2d(base + L.stage + s * 2 * box, &tmap_in, bar(AB_FULL + s), n0, (int)(b * N));
        tma_load_2d(base + L.stage + s * 2 * box + box, &tmap_in, bar(AB_FULL + s), n0 + 32, (int)(b * N));
        if (++s == nst) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t id = idesc(128, NT, 0, 1);
      int s = 0; uint32_t ph = 0;
      long long it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int d = (int)(it & 1);
        mbar_wait(bar(AB_FULL + s), ph);
        mbar_wait(bar(AB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        for (int k = 0; k < N / 8; ++k)
          mma_ts(tmem_d + d * NT, tmem_a + 8 * k, desc_mn(base + L.stage + s * 2 * box + k * 1024, box), id, k ? 1u : 0u);
        mma_commit(bar(AB_EMPTY + s));
        mma_commit(bar(AB_DFULL + d));
        if (++s == nst) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int et = (warp - 2) * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* E = reinterpret_cast<float*>(sm + L.e);
    const int niter = (m * 32 + 127) / 128;
    long long it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int d = (int)(it & 1);
      const long long b = t / ntn;
      const int p0 = (int)(t - b * ntn) * (NT / 2);
      // accumulating launch: the previous partial result comes from an earlier launch, so fetch it before waiting for this
      // tile's MMAs (a load -> add -> store chain per element inside the store loop made these launches latency-bound)
      float2 prev[16];
      if (accumulate) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = et + 128 * i, k = idx >> 5, j = idx & 31;
          prev[i] = make_float2(0.f, 0.f);
          if (i < niter && k < m && p0 + j < Pout) prev[i] = __ldcs(out + (size_t)b * m * Pout + (size_t)k * Pout + p0 + j);
        }
      }
      mbar_wait(bar(AB_DFULL + d), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[64];
      tmem_ld32(tmem_d + d * NT + lane_base, v);
      tmem_ld32(tmem_d + d * NT + lane_base + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(AB_DEMPTY + d));
      epi_bar_sync();                                  // previous tile's readers of E are done
      {
        float* er = E + (32 * q + lane) * AA_ELD;
#pragma unroll
        for (int c = 0; c < 64; c += 2) *reinterpret_cast<float2*>(er + c) = make_float2(v[c], v[c + 1]);
      }
      epi_bar_sync();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = et + 128 * i, k = idx >> 5, j = idx & 31;
        if (i < niter && k < m && p0 + j < Pout) {
          const float2 c = *reinterpret_cast<const float2*>(E + k * AA_ELD + 2 * j);
          const float2 s = *reinterpret_cast<const float2*>(E + (64 + k) * AA_ELD + 2 * j);
          float2 o = make_float2((c.x + s.y) * comp, (c.y - s.x) * comp);   // (C - iS)(re + i im)
          const size_t oi = (size_t)b * m * Pout + (size_t)k * Pout + p0 + j;
          if (accumulate) { o.x += prev[i].x; o.y += prev[i].y; }
          out[oi] = o;
          if (out_lo) out_lo[oi] = make_float2(split_lo(o.x), split_lo(o.y));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// y analysis of whole [N rows][ld <= 64 floats] blocks (one batch item = one contiguous block): the stage after the z
// analysis when a row is short (ld = 36 at 17 retained kz).  gen_axis_analysis_kernel stages such a block as two TMA boxes of
// N rows -- 128-byte and 16-byte pieces -- and the TMA engine pays per ROW request, not per byte (1.4 us for an 18 KB block:
// 0.62 ms for the y stage of the fused block backward at the cfg-2 shape, 2.7 TB/s).  Here four loader warps stage the block
// with cp.async in 16-byte chunks, applying the MN-major swizzle (128-byte rows, 32-byte atoms) in software; the contraction
// and the epilogue are those of gen_axis_analysis_kernel with the whole row as the one column tile.
//   warp 0 MMA issuer + TMEM owner, warps 1..4 loaders, warps 5..8 epilogue
// ------------------------------------------------------------------------------------------------
constexpr int PA_THREADS = 32 * 13;   // warp 0 MMA, warps 1..4 loaders, warps 5..12 epilogue: two groups of four on alternate blocks
constexpr int PA_MAXST = 6;
constexpr int PA_ELD = 66;            // exchange-buffer row (64 accumulator columns + 2)
constexpr int PA_BOX = 128 * 128;      // one [128 rows][32 floats] box
struct PASmem { uint32_t stage, e, bars, tmem_slot, total; int nst; };
__host__ __device__ inline PASmem pa_smem(int nst) {
  PASmem s; uint32_t o = 0;
  s.nst = nst;
  s.stage = o; o += (uint32_t)nst * 2 * PA_BOX;
  s.e = o; o += 2 * 128 * PA_ELD * 4;                  // one exchange buffer per epilogue group
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { PB_FULL = 0, PB_EMPTY = PA_MAXST, PB_DFULL = 2 * PA_MAXST, PB_DEMPTY = 2 * PA_MAXST + 2 };
__device__ __forceinline__ void pa_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(PA_THREADS, 1)
gen_plane_analysis_kernel(const float* __restrict__ in, float2* __restrict__ out, const float* __restrict__ Aplain, int m, int ld,
                          int nbn, int Pout, long long nblocks, int nst, float comp) {
  constexpr int N = 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const PASmem L = pa_smem(nst);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {   // columns >= ld of a staged block are never written: zero every stage once
    float4* z = reinterpret_cast<float4*>(sm + L.stage);
    for (int i = threadIdx.x; i < nst * 2 * PA_BOX / 16; i += PA_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < PA_MAXST; ++i) { mbar_init(bar(PB_FULL + i), 128); mbar_init(bar(PB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(PB_DFULL + i), 1); mbar_init(bar(PB_DEMPTY + i), 4); }   // D[i] <-> epilogue group i
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(base + L.tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const uint32_t tmem_a = tmem, tmem_d = tmem + 128;     // A: 128 columns, D: 2 x 64 columns
  if (warp >= 5 && warp < 9) {
    const int q = warp & 3;
    const float* row = Aplain + (size_t)(32 * q + lane) * N;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(row + c0) + j);
        r[4 * j] = t.x; r[4 * j + 1] = t.y; r[4 * j + 2] = t.z; r[4 * j + 3] = t.w;
      }
      tmem_st32(tmem_a + ((uint32_t)(q * 32) << 16) + c0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < nblocks ? (nblocks - first + step - 1) / step : 0;
  if (warp >= 1 && warp < 5) {
    // ===== loaders: block `it` -> stage it % nst, up to nst - 1 blocks in flight; a landed block is published before the
    //       loader blocks on the stage of the next one =====
    const int lt = threadIdx.x - 32;                 // 0..127
    const int cpr = ld >> 2, nchunk = N * cpr;       // 16-byte chunks per row / per block
    // this thread's chunks qd = lt + 128 i are the same in every block: their swizzled offsets are formed once (the division
    // by the chunks-per-row inside the copy loop cost more than the copies)
    constexpr int PA_MAXCH = 16;                     // ld <= 64 -> at most 128 * 16 / 128 chunks per thread
    uint32_t doff[PA_MAXCH];
#pragma unroll
    for (int i = 0; i < PA_MAXCH; ++i) {
      const int qd = lt + 128 * i;
      const int r = qd / cpr, c9 = qd - r * cpr, j = c9 >> 3, cc = c9 & 7;
      doff[i] = (uint32_t)(j * PA_BOX + r * 128 + ((((cc >> 1) ^ (r & 3)) & 3) << 5) + (cc & 1) * 16);
    }
    auto issue = [&](long long it) {
      const int s = (int)(it % nst);
      mbar_wait(bar(PB_EMPTY + s), (uint32_t)(((it / nst) & 1) ^ 1));
      const float* src = in + (size_t)(first + it * step) * N * ld + (size_t)lt * 4;
      const uint32_t dst = base + L.stage + (uint32_t)s * 2 * PA_BOX;
#pragma unroll
      for (int i = 0; i < PA_MAXCH; ++i)
        if (lt + 128 * i < nchunk) pa_cp_async16(dst + doff[i], src + (size_t)i * 512);
    };
    // groups are committed one per block (possibly empty) so that wait_group<K> always means "block it has landed"
    auto wait_landed = [&](int depth) {              // all but the newest depth - 1 groups complete
      switch (depth) {
        case 1: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
      }
    };
    const int ahead = nst - 1;
    for (int pre = 0; pre < ahead; ++pre) {
      if (pre < mine) issue(pre);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (long long it = 0; it < mine; ++it) {
      wait_landed(ahead);
      fence_proxy_async();
      mbar_arrive(bar(PB_FULL + (int)(it % nst)));
      if (it + ahead < mine) issue(it + ahead);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  } else if (warp == 0) {
    if (elect_one()) {
      const uint32_t id = idesc(128, nbn, 0, 1);
      for (long long it = 0; it < mine; ++it) {
        const int s = (int)(it % nst), d = (int)(it & 1);
        mbar_wait(bar(PB_FULL + s), (uint32_t)((it / nst) & 1));
        mbar_wait(bar(PB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t st = base + L.stage + (uint32_t)s * 2 * PA_BOX;
#pragma unroll
        for (int k = 0; k < N / 8; ++k) mma_ts(tmem_d + d * 64, tmem_a + 8 * k, desc_mn(st + k * 1024, PA_BOX), id, k ? 1u : 0u);
        mma_commit(bar(PB_EMPTY + s));
        mma_commit(bar(PB_DFULL + d));
      }
    }
  } else {
    // Two epilogue groups (warps 5..8 and 9..12) take alternate blocks: block `it` uses accumulator D[it & 1], which belongs
    // to group it & 1 -- with one group its TMEM read, exchange through shared memory (two named-barrier syncs) and stores
    // of block it ran after those of block it - 1, 0.79 us per block against 0.52 us of HBM time (ncu: the loaders waiting
    // for an empty stage and the tensor pipe 28 % busy).
    const int q = warp & 3, grp = (warp - 5) >> 2;
    const int et = ((warp - 5) & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* E = reinterpret_cast<float*>(sm + L.e) + grp * (128 * PA_ELD);
    const int nout = m * Pout;
    // this thread's outputs idx = et + 128 i = (ky k, column j) are the same in every block
    constexpr int PA_MAXOUT = 16;                    // m <= 64, Pout <= 32 -> at most 2048 / 128 outputs per thread
    int eoff[PA_MAXOUT];
#pragma unroll
    for (int i = 0; i < PA_MAXOUT; ++i) {
      const int idx = et + 128 * i, k = idx / Pout, j = idx - k * Pout;
      eoff[i] = k * PA_ELD + 2 * j;
    }
    for (long long it = grp; it < mine; it += 2) {
      const int d = grp;
      const long long b = first + it * step;
      mbar_wait(bar(PB_DFULL + d), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[64];
      tmem_ld32(tmem_d + d * 64 + lane_base, v);
      if (nbn > 32) tmem_ld32(tmem_d + d * 64 + lane_base + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PB_DEMPTY + d));
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");   // previous readers of E done
      {
        float* er = E + (32 * q + lane) * PA_ELD;
#pragma unroll
        for (int c = 0; c < 64; c += 2)
          if (c < nbn) *reinterpret_cast<float2*>(er + c) = make_float2(v[c], v[c + 1]);
      }
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      float2* ob = out + (size_t)b * nout + et;
#pragma unroll
      for (int i = 0; i < PA_MAXOUT; ++i) {
        if (et + 128 * i < nout) {
          const float2 c = *reinterpret_cast<const float2*>(E + eoff[i]);
          const float2 sn = *reinterpret_cast<const float2*>(E + 64 * PA_ELD + eoff[i]);
          ob[128 * i] = make_float2((c.x + sn.y) * comp, (c.y - sn.x) * comp);   // (C - iS)(re + i im)
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// axis synthesis (contracted index = row): out[b][n][j] = sum_k e^{+i theta(k,n)} in[b][k][j]
//   warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (column half = (warp-2)/4)
// ------------------------------------------------------------------------------------------------
constexpr int AS_EPI = 8;
constexpr int AS_THREADS = 32 * (2 + AS_EPI);
constexpr int AS_MAXST = 4;
struct ASSmem { uint32_t stage, ac, as, stg, bars, tmem_slot, total; int nst; };
__host__ __device__ inline ASSmem as_smem(int N, int mk, int nst) {
  ASSmem s; uint32_t o = 0;
  s.nst = nst;
  s.stage = o; o += (uint32_t)nst * 2 * mk * 128;
  o = (o + 1023u) & ~1023u;
  s.ac = o; o += (uint32_t)N * mk * 4;       // N here = output rows padded to a multiple of 128
  s.as = o; o += (uint32_t)N * mk * 4;
  s.stg = o; o += AS_EPI * 32 * ST_LD * 4;
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}

__global__ void __launch_bounds__(AS_THREADS, 1)
gen_axis_synthesis_kernel(const __grid_constant__ CUtensorMap tmap_in, float* __restrict__ U, const float* __restrict__ Acimg,
                          const float* __restrict__ Asimg, int N, int Npad, int m, int mk, int ld_out, int ntn, int nmt,
                          long long ntiles, int nst, float comp, int accumulate, float* __restrict__ U_lo) {
  // comp / accumulate / U_lo: as in gen_axis_analysis_kernel
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const ASSmem L = as_smem(Npad, mk, nst);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box = (uint32_t)mk * 128;
  {
    const float4* s1 = reinterpret_cast<const float4*>(Acimg);
    const float4* s2 = reinterpret_cast<const float4*>(Asimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.ac);
    float4* d2 = reinterpret_cast<float4*>(sm + L.as);
    for (int i = threadIdx.x; i < Npad * mk / 4; i += AS_THREADS) { d1[i] = __ldg(s1 + i); d2[i] = __ldg(s2 + i); }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < AS_MAXST; ++i) { mbar_init(bar(AB_FULL + i), 1); mbar_init(bar(AB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(AB_DFULL + i), 1); mbar_init(bar(AB_DEMPTY + i), AS_EPI); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  // a tile is (batch item, 64-column block): its operand is staged once and contracted against every 128-row block of
  // the output (nmt of them), so the input crosses HBM exactly once

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 1;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(bar(AB_EMPTY + s), ph);
        mbar_arrive_expect_tx(bar(AB_FULL + s), 2 * box);
        const long long b = t / ntn;
        const int n0 = (int)(t - b * ntn) * NT;
        tma_load_2d(base + L.stage + s * 2 * box, &tmap_in, bar(AB_FULL + s), n0, (int)(b * m));
        tma_load_2d(base + L.stage + s * 2 * box + box, &tmap_in, bar(AB_FULL + s), n0 + 32, (int)(b * m));
        if (++s == nst) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t id = idesc(128, NT, 0, 1);
      int s = 0; uint32_t ph = 0;
      long long it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(bar(AB_FULL + s), ph);
        for (int mt = 0; mt < nmt; ++mt, ++it) {
          const int d = (int)(it & 1);
          mbar_wait(bar(AB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
          tc_fence_after();
          for (int ks = 0; ks < mk / 8; ++ks) {
            const uint32_t aoff = (uint32_t)(ks >> 2) * (uint32_t)Npad * 128 + (uint32_t)mt * 16384 + (uint32_t)(ks & 3) * 32;
            const uint64_t bd = desc_mn(base + L.stage + s * 2 * box + ks * 1024, box);
            mma_ss(tmem + d * 128, umma_desc(base + L.ac + aoff, 1024), bd, id, ks ? 1u : 0u);
            mma_ss(tmem + d * 128 + NT, umma_desc(base + L.as + aoff, 1024), bd, id, ks ? 1u : 0u);
          }
          mma_commit(bar(AB_DFULL + d));
        }
        mma_commit(bar(AB_EMPTY + s));
        if (++s == nst) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stg) + (warp - 2) * 32 * ST_LD;
    const int rsub = lane >> 3, c4 = lane & 7;
    long long it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
     for (int mt = 0; mt < nmt; ++mt, ++it) {
      const int d = (int)(it & 1);
      const long long b = t / ntn;
      const int n0 = (int)(t - b * ntn) * NT + half * 32;
      float* ub = U + ((size_t)b * N + (size_t)mt * 128 + 32 * q) * ld_out + n0;
      float4 prev[8];            // accumulating launch: fetch the earlier partial result before waiting for the MMAs
      if (accumulate) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int row = 4 * g + rsub;
          prev[g] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n0 + 4 * c4 < ld_out && mt * 128 + 32 * q + row < N)
            prev[g] = __ldcs(reinterpret_cast<const float4*>(ub + (size_t)row * ld_out + 4 * c4));
        }
      }
      mbar_wait(bar(AB_DFULL + d), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float c[32], s[32];
      tmem_ld32(tmem + d * 128 + lane_base + half * 32, c);
      tmem_ld32(tmem + d * 128 + NT + lane_base + half * 32, s);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(AB_DEMPTY + d));
      // (cos + i sin)(re + i im): re' = C re - S im, im' = C im + S re   (columns 2j = re, 2j+1 = im)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 o;
        o.x = (c[4 * j] - s[4 * j + 1]) * comp;
        o.y = (c[4 * j + 1] + s[4 * j]) * comp;
        o.z = (c[4 * j + 2] - s[4 * j + 3]) * comp;
        o.w = (c[4 * j + 3] + s[4 * j + 2]) * comp;
        *reinterpret_cast<float4*>(st + lane * ST_LD + 4 * j) = o;
      }
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int row = 4 * g + rsub;
        if (n0 + 4 * c4 < ld_out && mt * 128 + 32 * q + row < N) {
          float4 o = *reinterpret_cast<const float4*>(st + row * ST_LD + 4 * c4);
          float4* dst = reinterpret_cast<float4*>(ub + (size_t)row * ld_out + 4 * c4);
          if (accumulate) { o.x += prev[g].x; o.y += prev[g].y; o.z += prev[g].z; o.w += prev[g].w; }
          *dst = o;
          if (U_lo)
            *reinterpret_cast<float4*>(U_lo + (ub - U) + (size_t)row * ld_out + 4 * c4) =
                make_float4(split_lo(o.x), split_lo(o.y), split_lo(o.z), split_lo(o.w));
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// Zp[bc][kx][ky][0..ldz/2) = Z[bc][kx][ky][kz] * mode_scale[k] * row_scale[bc]   (zero in the padding columns)
__global__ void gen_pad_scale_kernel(const float2* __restrict__ Z, float2* __restrict__ Zp, const float* __restrict__ mode_scale,
                                     const float* __restrict__ row_scale, long long rs_stride, int zm, int ldh, long long Mrows,
                                     long long total, float2* __restrict__ Zp_lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / ldh;            // (bc, kx, ky)
  const int kz = (int)(i - row * ldh);
  float2 v = make_float2(0.f, 0.f);
  if (kz < zm) {
    const long long bc = row / Mrows, kr = row - bc * Mrows;
    float s = mode_scale ? __ldg(mode_scale + kr * zm + kz) : 1.f;
    if (row_scale) s *= __ldg(row_scale + bc * rs_stride);
    const float2 z = __ldg(Z + row * zm + kz);
    v = make_float2(z.x * s, z.y * s);
  }
  Zp[i] = v;
  if (Zp_lo) Zp_lo[i] = make_float2(split_lo(v.x), split_lo(v.y));
}

// lo[i] = lo part of src[i]: the data_lo operand of a split stage whose input arrives from outside (the exchange tensor of
// the slab-decomposed transform)
__global__ void gen_split_lo_kernel(const float4* __restrict__ src, float4* __restrict__ lo, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(src + i);
  lo[i] = make_float4(split_lo(v.x), split_lo(v.y), split_lo(v.z), split_lo(v.w));
}

// The same with the channel index moved inside: Zc[b][kx][ky][c][0..ldh) = Z[b*C + c][kx][ky][kz] * scales.  With the
// channels innermost the x and y synthesis stages treat (c, kz) as columns, and the rows of V = [b, x, y][c][ldz] that one
// block-tail tile reads (block_tc.cu, fused forward) are 64*ldz contiguous floats.
__global__ void gen_pad_scale_cinner_kernel(const float2* __restrict__ Z, float2* __restrict__ Zc, const float* __restrict__ mode_scale,
                                            const float* __restrict__ row_scale, long long rs_stride, int zm, int ldh, int C,
                                            long long Mrows, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kz = (int)(i % ldh);
  long long r = i / ldh;
  const int c = (int)(r % C);
  r /= C;
  const long long kr = r % Mrows, b = r / Mrows;      // kr = kx*my + ky
  float2 v = make_float2(0.f, 0.f);
  if (kz < zm) {
    const long long bc = b * C + c;
    float s = mode_scale ? __ldg(mode_scale + kr * zm + kz) : 1.f;
    if (row_scale) s *= __ldg(row_scale + bc * rs_stride);
    const float2 z = __ldg(Z + (bc * Mrows + kr) * zm + kz);
    v = make_float2(z.x * s, z.y * s);
  }
  Zc[i] = v;
}

}  // namespace gen

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder(const void* device_ptr);

static int pow2_at_least(int v) { int p = 32; while (p < v) p <<= 1; return p; }

static int gen_make_map(CUtensorMap* m, const void* ptr, long long rows, long long cols, int box_rows, bool mn_major) {
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  if (rows > 0x7fffffffLL) return fail(PDEPHI_ERR_UNSUPPORTED, "general transform route: too many rows for one tensor map");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (cols & 3))
    return fail(PDEPHI_ERR_INVALID, "general transform route: buffer must be 16-byte aligned with a row length multiple of 4");
  CUresult r = tensor_map_encoder(ptr)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PDEPHI_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return PDEPHI_OK;
}

bool gen_supported(const Plan* p) { return p->gen_ok != 0; }
bool gen_x3_supported(const Plan* p) { return p->gen_ok != 0 && p->g.x3_ok != 0; }

static size_t smem_budget(const Plan* p) { return p->smem_optin - 1024; }

int gen_plan_init(Plan* p) {
  p->gen_ok = 0;
  p->gen_arena = nullptr;
  p->gen_arena3 = nullptr;
  p->g.x3_ok = 0;
  const bool two_d = (p->Nz == 1 && p->mzh == 1);
  GenPlan& g = p->g;
  g.has_mid = two_d ? 0 : 1;
  g.zN = two_d ? p->Ny : p->Nz;
  g.zm = two_d ? p->my : p->mzh;
  g.ym = two_d ? 1 : p->my;
  g.ldz = (2 * g.zm + 3) / 4 * 4;
  g.nbm = (g.ldz + 15) / 16 * 16;
  g.mky = p->my <= 32 ? 32 : 64;
  g.mkx = p->mx <= 32 ? 32 : 64;
  auto axis_ok = [](int N, int m) { return N % 32 == 0 && N <= 256 && m <= 64; };
  bool ok = g.zN % 32 == 0 && g.zN >= 32 && g.zN <= 256 && g.nbm <= 128 && axis_ok(p->Nx, p->mx);
  if (!two_d) ok = ok && axis_ok(p->Ny, p->my) && (p->my * p->mzh) % 2 == 0;
  if (!ok || !tensor_map_encoder(nullptr)) return PDEPHI_OK;
  // shared-memory plans of the four kernels
  const int NKB = g.zN / 32, NKBK = (g.ldz + 31) / 32;
  {
    const size_t fixed = gen::za_smem(NKB, g.nbm, 0).total;
    const int nr = (int)((smem_budget(p) - fixed) / gen::KBYTES);
    if (smem_budget(p) < fixed || nr < 2) return PDEPHI_OK;
    g.za_ring = nr > gen::MAX_RING ? gen::MAX_RING : nr;
  }
  {
    const size_t fixed = gen::zs_smem(g.zN, NKBK, 0).total;
    const int nr = (int)((smem_budget(p) - fixed) / gen::KBYTES);
    if (smem_budget(p) < fixed || nr < 2) return PDEPHI_OK;
    g.zs_ring = nr > gen::MAX_RING ? gen::MAX_RING : nr;
  }
  auto aa_stages = [&](int N) {
    const size_t fixed = gen::aa_smem(N, 0).total;
    if (smem_budget(p) < fixed) return 0;
    const int ns = (int)((smem_budget(p) - fixed) / (2 * (size_t)N * 128));
    return ns > gen::AA_MAXST ? gen::AA_MAXST : ns;
  };
  auto pad128 = [](int N) { return (N + 127) / 128 * 128; };
  auto as_stages = [&](int N, int mk) {
    const size_t fixed = gen::as_smem(pad128(N), mk, 0).total + 1024;
    if (smem_budget(p) < fixed) return 0;
    const int ns = (int)((smem_budget(p) - fixed) / (2 * (size_t)mk * 128));
    return ns > gen::AS_MAXST ? gen::AS_MAXST : ns;
  };
  g.xa_st = aa_stages(p->Nx);
  g.xs_st = as_stages(p->Nx, g.mkx);
  g.ya_st = two_d ? 2 : aa_stages(p->Ny);
  g.ys_st = two_d ? 2 : as_stages(p->Ny, g.mky);
  if (g.xa_st < 2 || g.xs_st < 2 || g.ya_st < 2 || g.ys_st < 2) return PDEPHI_OK;

  const size_t n_za = (size_t)g.nbm * g.zN, n_zs = (size_t)g.zN * NKBK * 32;
  const size_t n_ya = two_d ? 0 : (size_t)128 * p->Ny, n_ys = two_d ? 0 : (size_t)pad128(p->Ny) * g.mky;
  const size_t n_xa = (size_t)128 * p->Nx, n_xs = (size_t)pad128(p->Nx) * g.mkx;
  const size_t n_zap = (!two_d && g.zN == 128) ? (size_t)128 * 128 : 0;
  const size_t total = (2 * n_za + 2 * n_zs + n_ya + 2 * n_ys + n_xa + 2 * n_xs + 2 * n_zap) * sizeof(float);
  PDEPHI_CUDA(cudaMalloc(&p->gen_arena, total));
  PDEPHI_CUDA(cudaMemset(p->gen_arena, 0, total));
  float* f = (float*)p->gen_arena;
  g.za[0] = f; f += n_za; g.za[1] = f; f += n_za;
  g.zs[0] = f; f += n_zs; g.zs[1] = f; f += n_zs;
  g.ya = f; f += n_ya; g.ysc = f; f += n_ys; g.yss = f; f += n_ys;
  g.xa = f; f += n_xa; g.xsc = f; f += n_xs; g.xss = f; f += n_xs;
  g.za_plain = n_zap ? f : nullptr; f += n_zap;
  g.za_plain_fwd = n_zap ? f : nullptr; f += n_zap;
  const double inv_total = 1.0 / ((double)p->Nx_total * p->Ny * p->Nz);
  const int nyq = (!two_d && p->Nz % 2 == 0 && p->mzh - 1 == p->Nz / 2) ? p->Nz / 2 : -1;
  const double comp = 1.0;                    // truncation bias of the data operands: undone in the epilogues (PDEPHI_GEN_COMP)
  const int wnorm = two_d ? 2 : 1;
  auto genimg = [&](float* img, int kind, int R, int K, int Nt, int off, int m, int wmode, int part = 0, int c0 = 0, int nc = 0) {
    const int n = R * K;
    gen::gen_image_kernel<<<(n + 255) / 256, 256>>>(img, kind, R, K, Nt, off, m, wmode, nyq, inv_total, comp, part, c0, nc);
  };
  genimg(g.za[0], 0, g.nbm, g.zN, g.zN, 0, g.zm, 0);          // analysis FORWARD: plain e^{-i}
  genimg(g.za[1], 0, g.nbm, g.zN, g.zN, 0, g.zm, wnorm);      // analysis ADJOINT: c_k / total
  genimg(g.zs[0], 1, g.zN, NKBK * 32, g.zN, 0, g.zm, wnorm);  // synthesis FORWARD: c_k / total
  genimg(g.zs[1], 1, g.zN, NKBK * 32, g.zN, 0, g.zm, 0);      // synthesis ADJOINT: plain
  if (g.za_plain) genimg(g.za_plain, 5, 128, 128, g.zN, 0, g.zm, wnorm);   // analysis ADJOINT, rows >= 2 zm zero
  if (g.za_plain_fwd) genimg(g.za_plain_fwd, 5, 128, 128, g.zN, 0, g.zm, 0);   // analysis FORWARD
  if (!two_d) {
    genimg(g.ya, 2, 128, p->Ny, p->Ny, 0, p->my, 0);
    genimg(g.ysc, 3, pad128(p->Ny), g.mky, p->Ny, 0, p->my, 0);
    genimg(g.yss, 4, pad128(p->Ny), g.mky, p->Ny, 0, p->my, 0);
  }
  genimg(g.xa, 2, 128, p->Nx, p->Nx_total, p->x_offset, p->mx, 0);
  genimg(g.xsc, 3, pad128(p->Nx), g.mkx, p->Nx_total, p->x_offset, p->mx, 0);
  genimg(g.xss, 4, pad128(p->Nx), g.mkx, p->Nx_total, p->x_offset, p->mx, 0);
  PDEPHI_LAUNCH_CHECK("gen_image_kernel");
  p->gen_ok = 1;

  // ---- split-TF32 variant: geometry of the column blocks, then the lo / stacked images --------------------------
  g.x3_ok = 0;
  {
    // z analysis: the stacked image of all nbm output columns takes zN * nbm * 8 bytes; the contracted index is split over
    // np launches (np divides the number of k-blocks) until a launch's slice leaves room for 4 data + 2 lo slots
    const int nc = g.nbm;
    int np = 1;
    while (np <= NKB && (NKB % np != 0 || gen::za3_smem(NKB / np, nc, 6, 0).total > smem_budget(p))) ++np;
    if (np > NKB || np > 8) return PDEPHI_OK;
    const size_t fixed = gen::za3_smem(NKB / np, nc, 0, 0).total;
    const int slots = (int)((smem_budget(p) - fixed) / gen::KBYTES);
    g.za3_np = np; g.za3_nc = nc; g.za3_lo = 2;
    g.za3_ring = slots - 2 > gen::MAX_RING ? gen::MAX_RING : slots - 2;
    // z synthesis: blocks of at most 128 output positions (TMEM: two accumulator pairs of 2 * nc columns)
    int nblk = 0, c0 = 0;
    int cap = 128;
    while (cap >= 32 && gen::zs3_smem(cap, NKBK, 2).total > smem_budget(p)) cap -= 32;
    if (cap < 32) return PDEPHI_OK;
    while (c0 < g.zN) {
      if (nblk == 8) return PDEPHI_OK;
      const int nb = g.zN - c0 < cap ? g.zN - c0 : cap;
      const size_t fx = gen::zs3_smem(nb, NKBK, 0).total;
      int nr = (int)((smem_budget(p) - fx) / (2 * gen::KBYTES));
      g.zs3_c0[nblk] = c0; g.zs3_nc[nblk] = nb; g.zs3_ring[nblk] = nr > gen::MAX_RING ? gen::MAX_RING : nr;
      c0 += nb; ++nblk;
    }
    g.zs3_np = nblk;
  }
  {
    size_t n3 = n_ya + 2 * n_ys + n_xa + 2 * n_xs;                          // lo parts of the six axis images
    n3 += 2 * (size_t)2 * g.za3_nc * g.zN;                                     // [direction] stacked z analysis image
    for (int j = 0; j < g.zs3_np; ++j) n3 += 2 * (size_t)2 * g.zs3_nc[j] * NKBK * 32;
    float* f3 = nullptr;
    PDEPHI_CUDA(cudaMalloc(&f3, n3 * sizeof(float)));
    p->gen_arena3 = f3;
    PDEPHI_CUDA(cudaMemset(f3, 0, n3 * sizeof(float)));
    g.ya_lo = f3; f3 += n_ya; g.ysc_lo = f3; f3 += n_ys; g.yss_lo = f3; f3 += n_ys;
    g.xa_lo = f3; f3 += n_xa; g.xsc_lo = f3; f3 += n_xs; g.xss_lo = f3; f3 += n_xs;
    for (int d = 0; d < 2; ++d) {
      g.za3[d] = f3; f3 += (size_t)2 * g.za3_nc * g.zN;
      genimg(g.za3[d], 0, 2 * g.za3_nc, g.zN, g.zN, 0, g.zm, d == 0 ? 0 : wnorm, 0, 0, g.za3_nc);
      for (int j = 0; j < g.zs3_np; ++j) {
        g.zs3[d][j] = f3; f3 += (size_t)2 * g.zs3_nc[j] * NKBK * 32;
        genimg(g.zs3[d][j], 1, 2 * g.zs3_nc[j], NKBK * 32, g.zN, 0, g.zm, d == 0 ? wnorm : 0, 0, g.zs3_c0[j], g.zs3_nc[j]);
      }
    }
    if (!two_d) {
      genimg(g.ya_lo, 2, 128, p->Ny, p->Ny, 0, p->my, 0, 1);
      genimg(g.ysc_lo, 3, pad128(p->Ny), g.mky, p->Ny, 0, p->my, 0, 1);
      genimg(g.yss_lo, 4, pad128(p->Ny), g.mky, p->Ny, 0, p->my, 0, 1);
    }
    genimg(g.xa_lo, 2, 128, p->Nx, p->Nx_total, p->x_offset, p->mx, 0, 1);
    genimg(g.xsc_lo, 3, pad128(p->Nx), g.mkx, p->Nx_total, p->x_offset, p->mx, 0, 1);
    genimg(g.xss_lo, 4, pad128(p->Nx), g.mkx, p->Nx_total, p->x_offset, p->mx, 0, 1);
    PDEPHI_LAUNCH_CHECK("gen_image_kernel(lo)");
    g.x3_ok = 1;
  }
  return PDEPHI_OK;
}

void gen_plan_destroy(Plan* p) {
  if (p->gen_arena) cudaFree(p->gen_arena);
  if (p->gen_arena3) cudaFree(p->gen_arena3);
  p->gen_arena = p->gen_arena3 = nullptr;
}

// workspace layout of the general route: [W1 | W2 | Zp | W1lo | W2lo | Zplo]  (the lo arrays: split-TF32 variant only)
struct GenWs { size_t w1, w2, zp, w1lo, w2lo, zplo, total; };
static GenWs gen_ws(const Plan* p, long long BC) {
  const GenPlan& g = p->g;
  const size_t rows = (size_t)BC * p->Nx * (g.has_mid ? p->Ny : 1);
  const size_t b1 = align_up(rows * g.ldz * sizeof(float), 1024);
  const size_t b2 = align_up((size_t)BC * p->Nx * g.ym * g.ldz * sizeof(float), 1024);
  const size_t b3 = align_up((size_t)BC * p->mx * g.ym * g.ldz * sizeof(float), 1024);
  GenWs w;
  w.w1 = 0; w.w2 = b1; w.zp = b1 + b2;
  size_t o = b1 + b2 + b3;
  w.w1lo = w.w2lo = w.zplo = o;
  if (g.x3_ok) { w.w1lo = o; w.w2lo = o + b1; w.zplo = o + b1 + b2; o += b1 + b2 + b3; }
  w.total = o;
  return w;
}
size_t gen_workspace_bytes(const Plan* p, long long BC) { return gen_ws(p, BC).total; }

static int launch_axis_analysis(const Plan* p, const float* in, long long nb, int N, int m, int ld_in, int Pout, float2* out,
                                const float* Aplain, int nst, int kid, cudaStream_t st, float comp = PDEPHI_GEN_COMP,
                                int accumulate = 0, float2* out_lo = nullptr) {
  CUtensorMap map;
  int rc = gen_make_map(&map, in, nb * N, ld_in, N, true);
  if (rc) return rc;
  const int ntn = (ld_in + gen::NT - 1) / gen::NT;
  const long long ntiles = nb * ntn;
  const size_t smem = gen::aa_smem(N, nst).total + 1024;
  const int grid = (int)(ntiles < p->num_sms ? ntiles : p->num_sms);
  auto k = gen::gen_axis_analysis_kernel;
  PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProfScope ps(kid, st);
    k<<<grid, gen::AA_THREADS, smem, st>>>(map, out, Aplain, N, m, Pout, ntn, ntiles, nst, comp, accumulate, out_lo);
  }
  PDEPHI_LAUNCH_CHECK("gen_axis_analysis_kernel");
  return PDEPHI_OK;
}

// y analysis of contiguous [128][ld] blocks with cp.async staging (gen_plane_analysis_kernel); same result as
// launch_axis_analysis(p, in, nb, 128, m, ld, Pout, ...)
static bool plane_analysis_ok(const Plan* p, int N, int ld, int Pout) {
  return N == 128 && ld % 4 == 0 && ld >= 4 && ld <= 64 && 2 * Pout <= ld && smem_budget(p) >= gen::pa_smem(2).total;
}
static int launch_plane_analysis(const Plan* p, const float* in, long long nb, int m, int ld, int Pout, float2* out,
                                 const float* Aplain, int kid, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(in) & 15) return fail(PDEPHI_ERR_INVALID, "plane analysis: input must be 16-byte aligned");
  int nst = (int)((smem_budget(p) - gen::pa_smem(0).total) / (2 * (size_t)gen::PA_BOX));
  if (nst > gen::PA_MAXST) nst = gen::PA_MAXST;
  const int nbn = (ld + 15) / 16 * 16;
  const size_t smem = gen::pa_smem(nst).total + 1024;
  const int grid = (int)(nb < p->num_sms ? nb : p->num_sms);
  auto k = gen::gen_plane_analysis_kernel;
  PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProfScope ps(kid, st);
    k<<<grid, gen::PA_THREADS, smem, st>>>(in, out, Aplain, m, ld, nbn, Pout, nb, nst, PDEPHI_GEN_COMP);
  }
  PDEPHI_LAUNCH_CHECK("gen_plane_analysis_kernel");
  return PDEPHI_OK;
}

static int launch_axis_synthesis(const Plan* p, const float* in, long long nb, int N, int m, int mk, int ld, float* out,
                                 const float* Ac, const float* As, int nst, int kid, cudaStream_t st,
                                 float comp = PDEPHI_GEN_COMP, int accumulate = 0, float* out_lo = nullptr) {
  CUtensorMap map;
  int rc = gen_make_map(&map, in, nb * m, ld, mk, true);
  if (rc) return rc;
  const int ntn = (ld + gen::NT - 1) / gen::NT;
  const int Npad = (N + 127) / 128 * 128;
  const int nmt = Npad / 128;
  const long long ntiles = nb * ntn;
  const size_t smem = gen::as_smem(Npad, mk, nst).total + 1024;
  const int grid = (int)(ntiles < p->num_sms ? ntiles : p->num_sms);
  auto k = gen::gen_axis_synthesis_kernel;
  PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProfScope ps(kid, st);
    k<<<grid, gen::AS_THREADS, smem, st>>>(map, out, Ac, As, N, Npad, m, mk, ld, ntn, nmt, ntiles, nst, comp, accumulate, out_lo);
  }
  PDEPHI_LAUNCH_CHECK("gen_axis_synthesis_kernel");
  return PDEPHI_OK;
}

// One axis stage on either route.  Split route: data.A ~ data.Ahi + data_lo.Ahi + data.Alo as three launches that meet in
// fp32 in the epilogue (the tensor core reads `data` as its hi part); the last one can emit the lo part of the result.
static int axis_analysis_stage(const Plan* p, bool x3, const float* in, const float* in_lo, long long nb, int N, int m, int ld_in,
                               int Pout, float2* out, float2* out_lo, const float* Ahi, const float* Alo, int nst,
                               cudaStream_t st) {
  if (!x3) return launch_axis_analysis(p, in, nb, N, m, ld_in, Pout, out, Ahi, nst, K_GEN_AXIS_ANALYSIS, st);
  int rc = launch_axis_analysis(p, in, nb, N, m, ld_in, Pout, out, Ahi, nst, K_GEN_AXIS_ANALYSIS, st, 1.0f, 0, nullptr);
  if (rc) return rc;
  rc = launch_axis_analysis(p, in_lo, nb, N, m, ld_in, Pout, out, Ahi, nst, K_GEN_AXIS_ANALYSIS, st, 1.0f, 1, nullptr);
  if (rc) return rc;
  return launch_axis_analysis(p, in, nb, N, m, ld_in, Pout, out, Alo, nst, K_GEN_AXIS_ANALYSIS, st, 1.0f, 1, out_lo);
}

static int axis_synthesis_stage(const Plan* p, bool x3, const float* in, const float* in_lo, long long nb, int N, int m, int mk,
                                int ld, float* out, float* out_lo, const float* Ac, const float* As, const float* Ac_lo,
                                const float* As_lo, int nst, cudaStream_t st) {
  if (!x3) return launch_axis_synthesis(p, in, nb, N, m, mk, ld, out, Ac, As, nst, K_GEN_AXIS_SYNTHESIS, st);
  int rc = launch_axis_synthesis(p, in, nb, N, m, mk, ld, out, Ac, As, nst, K_GEN_AXIS_SYNTHESIS, st, 1.0f, 0, nullptr);
  if (rc) return rc;
  rc = launch_axis_synthesis(p, in_lo, nb, N, m, mk, ld, out, Ac, As, nst, K_GEN_AXIS_SYNTHESIS, st, 1.0f, 1, nullptr);
  if (rc) return rc;
  return launch_axis_synthesis(p, in, nb, N, m, mk, ld, out, Ac_lo, As_lo, nst, K_GEN_AXIS_SYNTHESIS, st, 1.0f, 1, out_lo);
}

static int launch_split_lo(const float* src, float* lo, size_t nfloats, cudaStream_t st) {
  if ((nfloats & 3) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15))
    return fail(PDEPHI_ERR_INVALID, "general transform route: exchange tensor must be 16-byte aligned with a multiple of 4 floats");
  const long long n4 = (long long)(nfloats / 4);
  {
    ProfScope ps(K_SCALE_MODES, st);
    gen::gen_split_lo_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>((const float4*)src, (float4*)lo, n4);
  }
  PDEPHI_LAUNCH_CHECK("gen_split_lo_kernel");
  return PDEPHI_OK;
}

int analysis_gen(const Plan* p, const float* x, float2* X, long long BC, int direction, void* ws, cudaStream_t st, int stages,
                 float2* T, bool x3) {
  const GenPlan& g = p->g;
  if (x3 && !g.x3_ok) return fail(PDEPHI_ERR_UNSUPPORTED, "split-TF32 per-axis route is not available for this shape");
  const GenWs w = gen_ws(p, BC);
  char* wb = (char*)ws;
  float* W1 = (float*)(wb + w.w1);
  float* W2 = (float*)(wb + w.w2);
  float* W1lo = (float*)(wb + w.w1lo);
  float* W2lo = (float*)(wb + w.w2lo);
  if (stages != ST_ALL) {
    if (!g.has_mid) return fail(PDEPHI_ERR_UNSUPPORTED, "stage-split transforms are not available for 2-D grids");
    W2 = (float*)T;        // the exchange tensor [BC, Nx, my, mzh] takes the place of the (y -> x) intermediate
  }
  const long long rows = BC * p->Nx * (g.has_mid ? p->Ny : 1);
  if (stages & ST_ZY) {
    CUtensorMap map;
    int rc = gen_make_map(&map, x, rows, g.zN, 128, false);
    if (rc) return rc;
    const int NKB = g.zN / 32;
    const int ntiles = (int)((rows + 127) / 128);
    const int grid = ntiles < p->num_sms ? ntiles : p->num_sms;
    if (!x3) {
      const size_t smem = gen::za_smem(NKB, g.nbm, g.za_ring).total + 1024;
      auto k = gen::gen_z_analysis_kernel;
      PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      {
        ProfScope ps(K_GEN_Z_ANALYSIS, st);
        k<<<grid, gen::ZA_THREADS, smem, st>>>(map, W1, g.za[direction], rows, ntiles, NKB, g.nbm, g.ldz, g.za_ring);
      }
      PDEPHI_LAUNCH_CHECK("gen_z_analysis_kernel");
    } else {
      const int nc = g.za3_nc, np = g.za3_np, nkb = NKB / np;
      const size_t smem = gen::za3_smem(nkb, nc, g.za3_ring, g.za3_lo).total + 1024;
      const int tmem_cols = pow2_at_least(4 * nc + ((nc & 31) ? 32 : 0));
      auto k = gen::gen_z_analysis_x3_kernel;
      PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int j = 0; j < np; ++j) {             // one launch per range of k-blocks (the contracted index)
        ProfScope ps(K_GEN_Z_ANALYSIS, st);
        k<<<grid, gen::ZA3_THREADS, smem, st>>>(map, W1, j == np - 1 ? W1lo : nullptr,
                                               g.za3[direction] + (size_t)j * nkb * 2 * nc * 32, rows, ntiles, nkb, nc, g.ldz,
                                               j * nkb, j > 0 ? 1 : 0, g.za3_ring, g.za3_lo, tmem_cols);
      }
      PDEPHI_LAUNCH_CHECK("gen_z_analysis_x3_kernel");
    }
    if (g.has_mid) {
      rc = axis_analysis_stage(p, x3, W1, W1lo, BC * p->Nx, p->Ny, p->my, g.ldz, p->mzh, (float2*)W2,
                               (stages & ST_X) ? (float2*)W2lo : nullptr, g.ya, g.ya_lo, g.ya_st, st);
      if (rc) return rc;
    }
  }
  if (!(stages & ST_X)) return PDEPHI_OK;
  if (g.has_mid) {
    if (x3 && stages == ST_X) {       // the exchange tensor arrives from outside: form its lo part here
      int rc = launch_split_lo(W2, W2lo, (size_t)BC * p->Nx * 2 * p->my * p->mzh, st);
      if (rc) return rc;
    }
    return axis_analysis_stage(p, x3, W2, W2lo, BC, p->Nx, p->mx, 2 * p->my * p->mzh, p->my * p->mzh, X, nullptr, g.xa, g.xa_lo,
                               g.xa_st, st);
  }
  return axis_analysis_stage(p, x3, W1, W1lo, BC, p->Nx, p->mx, g.ldz, g.zm, X, nullptr, g.xa, g.xa_lo, g.xa_st, st);
}

static int launch_pad_scale(const float2* Z, float2* Zp, const float* mode_scale, const float* row_scale, long long rs_stride,
                            int zm, int ldh, long long Mrows, long long total, cudaStream_t st, float2* Zp_lo = nullptr) {
  {
    ProfScope ps(K_SCALE_MODES, st);
    gen::gen_pad_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Z, Zp, mode_scale, row_scale, rs_stride, zm, ldh,
                                                                             Mrows, total, Zp_lo);
  }
  PDEPHI_LAUNCH_CHECK("gen_pad_scale_kernel");
  return PDEPHI_OK;
}

int synthesis_gen(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1, const float* mode_scale,
                  const float* row_scale, long long rs_stride, long long BC, int direction, void* ws, cudaStream_t st, int stages,
                  float2* T, bool x3) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(out) || !al16(add0) || !al16(add1)) return fail(PDEPHI_ERR_INVALID, "synthesis: buffers must be 16-byte aligned");
  const GenPlan& g = p->g;
  if (x3 && !g.x3_ok) return fail(PDEPHI_ERR_UNSUPPORTED, "split-TF32 per-axis route is not available for this shape");
  const GenWs w = gen_ws(p, BC);
  char* wb = (char*)ws;
  float* W1 = (float*)(wb + w.w1);
  float* W2 = (float*)(wb + w.w2);
  float* Zp = (float*)(wb + w.zp);
  float* W1lo = (float*)(wb + w.w1lo);
  float* W2lo = (float*)(wb + w.w2lo);
  float* Zplo = (float*)(wb + w.zplo);
  const int ldh = g.ldz / 2;
  const long long Mrows = (long long)p->mx * g.ym;
  const long long rows = BC * p->Nx * (g.has_mid ? p->Ny : 1);
  if (stages != ST_ALL && !g.has_mid) return fail(PDEPHI_ERR_UNSUPPORTED, "stage-split transforms are not available for 2-D grids");
  if (stages == ST_X) {
    // x stage alone: scaled modes -> unpadded exchange tensor T [BC, Nx, my, mzh]
    const float* src = (const float*)Z;
    if (mode_scale || row_scale || x3) {
      int rc = launch_pad_scale(Z, (float2*)Zp, mode_scale, row_scale, rs_stride, g.zm, g.zm, Mrows, BC * Mrows * g.zm, st,
                                x3 ? (float2*)Zplo : nullptr);
      if (rc) return rc;
      src = Zp;
    }
    return axis_synthesis_stage(p, x3, src, Zplo, BC, p->Nx, p->mx, g.mkx, 2 * p->my * p->mzh, (float*)T, nullptr, g.xsc, g.xss,
                                g.xsc_lo, g.xss_lo, g.xs_st, st);
  }
  const float* zin;
  if (stages == ST_ZY) {
    // (y,z) stages alone from the unpadded exchange tensor: re-pad its rows when 2*mzh is not a multiple of 4
    const float* src = (const float*)T;
    if (g.ldz != 2 * g.zm) {
      int rc = launch_pad_scale(T, (float2*)W2, nullptr, nullptr, 0, g.zm, ldh, 1, BC * p->Nx * (long long)p->my * ldh, st,
                                x3 ? (float2*)W2lo : nullptr);
      if (rc) return rc;
      src = W2;
    } else if (x3) {
      int rc = launch_split_lo(src, W2lo, (size_t)BC * p->Nx * p->my * g.ldz, st);
      if (rc) return rc;
    }
    int rc = axis_synthesis_stage(p, x3, src, W2lo, BC * p->Nx, p->Ny, p->my, g.mky, g.ldz, W1, W1lo, g.ysc, g.yss, g.ysc_lo,
                                  g.yss_lo, g.ys_st, st);
    if (rc) return rc;
    zin = W1;
  } else {
    const float* src = (const float*)Z;
    if (mode_scale || row_scale || g.ldz != 2 * g.zm || x3) {
      int rc = launch_pad_scale(Z, (float2*)Zp, mode_scale, row_scale, rs_stride, g.zm, ldh, Mrows, BC * Mrows * ldh, st,
                                x3 ? (float2*)Zplo : nullptr);
      if (rc) return rc;
      src = Zp;
    }
    if (g.has_mid) {
      int rc = axis_synthesis_stage(p, x3, src, Zplo, BC, p->Nx, p->mx, g.mkx, g.ym * g.ldz, W2, W2lo, g.xsc, g.xss, g.xsc_lo,
                                    g.xss_lo, g.xs_st, st);
      if (rc) return rc;
      rc = axis_synthesis_stage(p, x3, W2, W2lo, BC * p->Nx, p->Ny, p->my, g.mky, g.ldz, W1, W1lo, g.ysc, g.yss, g.ysc_lo,
                                g.yss_lo, g.ys_st, st);
      if (rc) return rc;
    } else {
      int rc = axis_synthesis_stage(p, x3, src, Zplo, BC, p->Nx, p->mx, g.mkx, g.ldz, W1, W1lo, g.xsc, g.xss, g.xsc_lo, g.xss_lo,
                                    g.xs_st, st);
      if (rc) return rc;
    }
    zin = W1;
  }
  CUtensorMap map;
  int rc = gen_make_map(&map, zin, rows, g.ldz, 128, false);
  if (rc) return rc;
  const int NKBK = (g.ldz + 31) / 32, ksteps = (g.ldz + 7) / 8;
  const int ntiles = (int)((rows + 127) / 128);
  const int grid = ntiles < p->num_sms ? ntiles : p->num_sms;
  if (add1 && !add0) { add0 = add1; add1 = nullptr; }
  const int nadd = (add0 ? 1 : 0) + (add1 ? 1 : 0);
  if (x3) {
    CUtensorMap maplo;
    rc = gen_make_map(&maplo, W1lo, rows, g.ldz, 128, false);
    if (rc) return rc;
    for (int j = 0; j < g.zs3_np; ++j) {         // one launch per block of output positions
      const int nc = g.zs3_nc[j], c0 = g.zs3_c0[j];
      const size_t smem = gen::zs3_smem(nc, NKBK, g.zs3_ring[j]).total + 1024;
      const int tmem_cols = pow2_at_least(4 * nc);
#define PDEPHI_GZS3_LAUNCH(NA)                                                                                           \
  {                                                                                                                      \
    auto k = gen::gen_z_synthesis_x3_kernel<NA>;                                                                         \
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
    ProfScope ps(K_GEN_Z_SYNTHESIS, st);                                                                                 \
    k<<<grid, gen::ZS_THREADS, smem, st>>>(map, maplo, out, add0, add1, g.zs3[direction][j], rows, ntiles, g.zN, c0, nc, \
                                           NKBK, ksteps, g.zs3_ring[j], tmem_cols);                                      \
  }
      if (nadd == 0) PDEPHI_GZS3_LAUNCH(0) else if (nadd == 1) PDEPHI_GZS3_LAUNCH(1) else PDEPHI_GZS3_LAUNCH(2)
#undef PDEPHI_GZS3_LAUNCH
    }
    PDEPHI_LAUNCH_CHECK("gen_z_synthesis_x3_kernel");
    return PDEPHI_OK;
  }
  const size_t smem = gen::zs_smem(g.zN, NKBK, g.zs_ring).total + 1024;
  const int tmem_cols = pow2_at_least(2 * g.zN);
#define PDEPHI_GZS_LAUNCH(NA)                                                                                            \
  {                                                                                                                      \
    auto k = gen::gen_z_synthesis_kernel<NA>;                                                                            \
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
    ProfScope ps(K_GEN_Z_SYNTHESIS, st);                                                                                 \
    k<<<grid, gen::ZS_THREADS, smem, st>>>(map, out, add0, add1, g.zs[direction], rows, ntiles, g.zN, NKBK, ksteps,     \
                                           g.zs_ring, tmem_cols);                                                        \
  }
  if (nadd == 0) PDEPHI_GZS_LAUNCH(0) else if (nadd == 1) PDEPHI_GZS_LAUNCH(1) else PDEPHI_GZS_LAUNCH(2)
#undef PDEPHI_GZS_LAUNCH
  PDEPHI_LAUNCH_CHECK("gen_z_synthesis_kernel");
  return PDEPHI_OK;
}

int block_tail_fwd_fused_launch(const float* h, const float* V, int ldz, const float* bzimg, const float* Wc, const float* bc,
                                float* u, float* hout, int B, long long S, const float* Wp, const float* bp, float* proj_out,
                                int n_out, const float* Win, const float* bin, int n_in, cudaStream_t st,
                                const float* za_img, float* vh);

// The fused block forward is available when the per-axis kernels cover the shape and a 128-point tile of the block
// tail is exactly one z line.
bool block_spectral_ok(const Plan* p, int C) {
  return p->gen_ok && p->g.has_mid && p->Nz == 128 && C == 64 && p->g.ldz <= 64 && p->Nx == p->Nx_total;
}

}  // namespace pdephi

using namespace pdephi;

extern "C" int pdephi_block_spectral_supported(pdephi_plan_t plan, int C) {
  const Plan* p = (const Plan*)plan;
  return (p && block_spectral_ok(p, C)) ? 1 : 0;
}

extern "C" size_t pdephi_block_spectral_fwd_workspace_bytes(pdephi_plan_t plan, int B, int C) {
  const Plan* p = (const Plan*)plan;
  if (!p || !block_spectral_ok(p, C) || B <= 0) return 0;
  return gen_workspace_bytes(p, (long long)B * C);
}

static int block_spectral_run(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                              long long row_scale_stride, const float* h, const float* Wc, const float* bc, float* u, float* hout,
                              const float* Wp, const float* bp, int n_out, float* proj_out, const float* Win, const float* bin,
                              int n_in, int B, int C, void* workspace, size_t workspace_bytes, pdephi_stream_t stream,
                              float* vz = nullptr);

extern "C" int pdephi_block_spectral_fwd(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                                         long long row_scale_stride, const float* h, const float* Wc, const float* bc, float* u,
                                         float* hout, int B, int C, void* workspace, size_t workspace_bytes,
                                         pdephi_stream_t stream) {
  return block_spectral_run(plan, Y, mode_scale, row_scale, row_scale_stride, h, Wc, bc, u, hout, nullptr, nullptr, 0, nullptr,
                            nullptr, nullptr, 0, B, C, workspace, workspace_bytes, stream);
}

extern "C" int pdephi_block_spectral_lift_fwd(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                                              long long row_scale_stride, const float* x, const float* Win, const float* bin,
                                              int n_in, const float* Wc, const float* bc, float* u, float* hout, int B, int C,
                                              void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(x && Win, "block_spectral_lift_fwd: null pointer");
  return block_spectral_run(plan, Y, mode_scale, row_scale, row_scale_stride, x, Wc, bc, u, hout, nullptr, nullptr, 0, nullptr, Win,
                            bin, n_in, B, C, workspace, workspace_bytes, stream);
}

extern "C" int pdephi_block_spectral_proj_fwd(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                                              long long row_scale_stride, const float* h, const float* Wc, const float* bc,
                                              float* u, float* hout, const float* Wp, const float* bp, int n_out, float* out,
                                              int B, int C, void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(Wp && out, "block_spectral_proj_fwd: null pointer");
  return block_spectral_run(plan, Y, mode_scale, row_scale, row_scale_stride, h, Wc, bc, u, hout, Wp, bp, n_out, out, nullptr,
                            nullptr, 0, B, C, workspace, workspace_bytes, stream);
}

// ---- fused block backward: block-tail backward with the z stage of the ADJOINT analysis of gu inside it (it leaves the
// per-axis route's z-analysed intermediate), then that route's y stage and the split x stage ----
namespace pdephi {
bool xstage_analysis_tc_ok(const Plan* p);
bool xstage_analysis_tc3_ok(const Plan* p);
int xstage_analysis_cols(const Plan* p, const float2* in, float2* X, long long nb, int Pcols, bool x3, cudaStream_t st);
int block_tail_bwd_za_launch(const float* g, const float* u, const float* h, const float* Wc, float* t, float* dWc, float* dbc,
                             int B, long long S, void* workspace, size_t workspace_bytes, const float* Wp, int n_out,
                             const float* Win, const float* bin, int n_in, float* G, const float* za_img, float* vg, int ldz,
                             cudaStream_t st);
}

extern "C" size_t pdephi_block_spectral_bwd_workspace_bytes(pdephi_plan_t plan, int B, int C) {
  const Plan* p = (const Plan*)plan;
  if (!p || !block_spectral_ok(p, C) || !p->g.za_plain || B <= 0) return 0;
  return gen_workspace_bytes(p, (long long)B * C);
}

extern "C" int pdephi_block_spectral_bwd(pdephi_plan_t plan, const float* g, const float* u, const float* h, const float* Wc,
                                         float* t, float* dWc, float* dbc, float* dZ, int B, int C, const float* Wp, int n_out,
                                         const float* Win, const float* bin, int n_in, float* G, void* tail_workspace,
                                         size_t tail_workspace_bytes, void* workspace, size_t workspace_bytes,
                                         pdephi_stream_t stream) {
  const Plan* p = (const Plan*)plan;
  PDEPHI_REQUIRE(p && g && u && h && Wc && dWc && dZ, "block_spectral_bwd: null pointer");     /* t, dbc, G may be NULL */
  if (!block_spectral_ok(p, C) || !p->g.za_plain)
    return fail(PDEPHI_ERR_UNSUPPORTED, "block_spectral_bwd: needs width 64, Nz = 128 and a shape of the per-axis tensor-core route");
  const long long BC = (long long)B * C;
  const size_t need = gen_workspace_bytes(p, BC);
  if (!workspace || workspace_bytes < need)
    return fail(PDEPHI_ERR_WORKSPACE, "block_spectral_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const GenPlan& gp = p->g;
  const GenWs w = gen_ws(p, BC);
  char* wb = (char*)workspace;
  float* Vg = (float*)(wb + w.w1);       // [BC, Nx, Ny, ldz]   z-analysed gu (the per-axis route's W1)
  float* W2 = (float*)(wb + w.w2);       // [BC, Nx, my, mzh]   after the y stage
  float* W2lo = (float*)(wb + w.w2lo);
  int rc = block_tail_bwd_za_launch(g, u, h, Wc, t, dWc, dbc, B, (long long)p->Nx * p->Ny * p->Nz, tail_workspace,
                                    tail_workspace_bytes, Wp, n_out, Win, bin, n_in, G, gp.za_plain, Vg, gp.ldz, st);
  if (rc) return rc;
  // y stage: batch = (volume, x), rows = y, columns = kz (one contiguous [Ny][ldz] block per batch item).  The x stage is small
  // and runs split (fp32-grade): its TF32 noise is what the coefficient gradients of the headline model are most sensitive to
  // (OPT_TF32_XSTAGE_SPLIT of the unfused route) -- the single-kernel split x stage of xstage_tc.cu where it applies (mx <= 32),
  // else three accumulating launches of the per-axis kernel.
  const bool x_fused = xstage_analysis_tc3_ok(p);
  const bool xsplit = !x_fused && gp.x3_ok != 0;
  if (!xsplit && plane_analysis_ok(p, p->Ny, gp.ldz, p->mzh))
    rc = launch_plane_analysis(p, Vg, BC * p->Nx, p->my, gp.ldz, p->mzh, (float2*)W2, gp.ya, K_GEN_AXIS_ANALYSIS, st);
  else
    rc = launch_axis_analysis(p, Vg, BC * p->Nx, p->Ny, p->my, gp.ldz, p->mzh, (float2*)W2, gp.ya, gp.ya_st, K_GEN_AXIS_ANALYSIS, st,
                              PDEPHI_GEN_COMP, 0, xsplit ? (float2*)W2lo : nullptr);
  if (rc) return rc;
  if (x_fused) return xstage_analysis_cols(p, (const float2*)W2, (float2*)dZ, BC, p->my * p->mzh, true, st);
  return axis_analysis_stage(p, xsplit, W2, W2lo, BC, p->Nx, p->mx, 2 * p->my * p->mzh, p->my * p->mzh, (float2*)dZ, nullptr, gp.xa,
                             gp.xa_lo, gp.xa_st, st);
}

// ---- the z stage of the NEXT layer's forward analysis inside the fused forward (vz), and the analysis from it ----
extern "C" size_t pdephi_block_spectral_vz_bytes(pdephi_plan_t plan, int B, int C) {
  const Plan* p = (const Plan*)plan;
  if (!p || !block_spectral_ok(p, C) || !p->g.za_plain_fwd || B <= 0) return 0;
  return (size_t)B * C * p->Nx * p->Ny * p->g.ldz * sizeof(float);
}

extern "C" int pdephi_block_spectral_fwd_z(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                                           long long row_scale_stride, const float* h, const float* Win, const float* bin, int n_in,
                                           const float* Wc, const float* bc, float* u, float* hout, float* vz, int B, int C,
                                           void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  const Plan* p = (const Plan*)plan;
  PDEPHI_REQUIRE(p && vz, "block_spectral_fwd_z: null pointer");
  if (!block_spectral_ok(p, C) || !p->g.za_plain_fwd)
    return fail(PDEPHI_ERR_UNSUPPORTED, "block_spectral_fwd_z: needs width 64, Nz = 128 and a shape of the per-axis tensor-core route");
  return block_spectral_run(plan, Y, mode_scale, row_scale, row_scale_stride, h, Wc, bc, u, hout, nullptr, nullptr, 0, nullptr, Win,
                            bin, n_in, B, C, workspace, workspace_bytes, stream, vz);
}

extern "C" int pdephi_analysis_from_z(pdephi_plan_t plan, const float* vz, float* X, long long BC, void* workspace,
                                      size_t workspace_bytes, pdephi_stream_t stream) {
  const Plan* p = (const Plan*)plan;
  PDEPHI_REQUIRE(p && vz && X && BC > 0, "analysis_from_z: null pointer");
  if (!p->gen_ok || !p->g.has_mid) return fail(PDEPHI_ERR_UNSUPPORTED, "analysis_from_z: not a shape of the per-axis tensor-core route");
  const size_t need = gen_workspace_bytes(p, BC);
  if (!workspace || workspace_bytes < need)
    return fail(PDEPHI_ERR_WORKSPACE, "analysis_from_z: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const GenPlan& gp = p->g;
  const GenWs w = gen_ws(p, BC);
  char* wb = (char*)workspace;
  float* W2 = (float*)(wb + w.w2);       // [BC, Nx, my, mzh] after the y stage
  float* W2lo = (float*)(wb + w.w2lo);
  const bool x_fused = xstage_analysis_tc3_ok(p);
  const bool xsplit = !x_fused && gp.x3_ok != 0;
  int rc;
  if (!xsplit && plane_analysis_ok(p, p->Ny, gp.ldz, p->mzh))
    rc = launch_plane_analysis(p, vz, BC * p->Nx, p->my, gp.ldz, p->mzh, (float2*)W2, gp.ya, K_GEN_AXIS_ANALYSIS, st);
  else
    rc = launch_axis_analysis(p, vz, BC * p->Nx, p->Ny, p->my, gp.ldz, p->mzh, (float2*)W2, gp.ya, gp.ya_st, K_GEN_AXIS_ANALYSIS, st,
                              PDEPHI_GEN_COMP, 0, xsplit ? (float2*)W2lo : nullptr);
  if (rc) return rc;
  if (x_fused) return xstage_analysis_cols(p, (const float2*)W2, (float2*)X, BC, p->my * p->mzh, true, st);
  return axis_analysis_stage(p, xsplit, W2, W2lo, BC, p->Nx, p->mx, 2 * p->my * p->mzh, p->my * p->mzh, (float2*)X, nullptr, gp.xa,
                             gp.xa_lo, gp.xa_st, st);
}

static int block_spectral_run(pdephi_plan_t plan, const float* Y, const float* mode_scale, const float* row_scale,
                              long long row_scale_stride, const float* h, const float* Wc, const float* bc, float* u, float* hout,
                              const float* Wp, const float* bp, int n_out, float* proj_out, const float* Win, const float* bin,
                              int n_in, int B, int C, void* workspace, size_t workspace_bytes, pdephi_stream_t stream, float* vz) {
  const Plan* p = (const Plan*)plan;
  PDEPHI_REQUIRE(p && Y && h && Wc && u && hout, "block_spectral_fwd: null pointer");
  if (!block_spectral_ok(p, C))
    return fail(PDEPHI_ERR_UNSUPPORTED, "block_spectral_fwd: needs width 64, Nz = 128 and a shape of the per-axis tensor-core route");
  const long long BC = (long long)B * C;
  const size_t need = gen_workspace_bytes(p, BC);
  if (!workspace || workspace_bytes < need)
    return fail(PDEPHI_ERR_WORKSPACE, "block_spectral_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const GenPlan& g = p->g;
  const GenWs w = gen_ws(p, BC);
  char* wb = (char*)workspace;
  float* V = (float*)(wb + w.w1);      // [B, Nx, Ny, C, ldz]
  float* W2 = (float*)(wb + w.w2);     // [BC, Nx, my, ldz]
  float* Zp = (float*)(wb + w.zp);     // [BC, mx, my, ldz]
  // modes -> channel-innermost padded layout [B, mx, my, C, ldz] (always: it is also the transposition)
  {
    const long long Mrows = (long long)p->mx * p->my;
    const long long total = (long long)B * Mrows * C * (g.ldz / 2);
    {
      ProfScope ps(K_SCALE_MODES, st);
      gen::gen_pad_scale_cinner_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float2*)Y, (float2*)Zp, mode_scale,
                                                                                      row_scale, row_scale_stride, g.zm, g.ldz / 2,
                                                                                      C, Mrows, total);
    }
    PDEPHI_LAUNCH_CHECK("gen_pad_scale_cinner_kernel");
  }
  // x stage: batch = volume b, columns = (ky, c, kz);  y stage: batch = (b, x), columns = (c, kz) -> V [b, x, y][c][ldz]
  int rc = launch_axis_synthesis(p, Zp, B, p->Nx, p->mx, g.mkx, p->my * C * g.ldz, W2, g.xsc, g.xss, g.xs_st, K_GEN_AXIS_SYNTHESIS, st);
  if (rc) return rc;
  rc = launch_axis_synthesis(p, W2, (long long)B * p->Nx, p->Ny, p->my, g.mky, C * g.ldz, V, g.ysc, g.yss, g.ys_st,
                             K_GEN_AXIS_SYNTHESIS, st);
  if (rc) return rc;
  return block_tail_fwd_fused_launch(h, V, g.ldz, g.zs[PDEPHI_FORWARD], Wc, bc, u, hout, B, (long long)p->Nx * p->Ny * p->Nz, Wp, bp,
                                     proj_out, n_out, Win, bin, n_in, st, vz ? g.za_plain_fwd : nullptr, vz);
}
