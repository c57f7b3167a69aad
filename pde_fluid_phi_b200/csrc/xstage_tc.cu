// K1 / K3, TF32 path: the x stage of the pruned transforms as tcgen05 contractions.
//
// The intermediate between the fused (z,y) stage and the x stage is [BC, Nx, P] complex = [BC*Nx rows][2P
// floats] (P = my*mzh); the retained modes are [BC, mx, P] complex.  Both are row-major with the contracted
// index (x or kx) on the ROWS, so a TMA box [K rows][32 floats] is an MN-major B operand (128-byte swizzle
// with 32-byte atoms, the only MN-major layout 32-bit operands have) and no transposition is ever needed:
//
//   analysis   D[64 x 64]   = [cos_x ; sin_x][64 x Nx] . T2tile[Nx x 64]          (rows: kx | 32+kx)
//              X[kx][p]     = (Dc[2p] + Ds[2p+1],  Dc[2p+1] - Ds[2p])              via a shared-memory exchange
//   synthesis  D1[128 x 64] = cos_x[128 x 32] . Ztile[32 x 64],  D2 = sin_x . Ztile   (rows: x)
//              U[x][p]      = (D1[2p] - D2[2p+1],  D1[2p+1] + D2[2p])              in registers (same TMEM lane)
//
// Persistent warp-specialised CTAs: warp 0 TMA producer, warp 1 MMA issuer, epilogue warps after that.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pdephi {
namespace xtc {
using namespace tcptx;

constexpr int SIN0 = 32;   // first sine row (analysis A operand)
constexpr int NT = 64;     // floats per n tile (32 complex modes)

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

// kind 0: analysis A [64 rows][K = Nx]: row m<mx -> cos(2 pi m (k+off)/Nt), row 32+m -> sin
// kind 1/2: synthesis A [Nx rows][K = 32]: (x, k<mx) -> cos / sin (2 pi k (x+off)/Nt)
__global__ void ximage_kernel(float* __restrict__ img, int kind, int R, int K, int Nt, int off, int m, double comp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * K) return;
  const int r = idx / K, k = idx % K;
  int mode = -1, pos = 0, sine = 0;
  if (kind == 0) { pos = k + off; if (r < m) mode = r; else if (r >= SIN0 && r < SIN0 + m) { mode = r - SIN0; sine = 1; } }
  else { pos = r + off; sine = (kind == 2); if (k < m) mode = k; }
  double v = 0.0;
  if (mode >= 0) {
    double s, c;
    sincospi(2.0 * (double)(((long long)mode * pos) % Nt) / (double)Nt, &s, &c);
    v = (sine ? s : c) * comp;
  }
  img[swz_off(R, r, k) / 4] = to_tf32((float)v);
}

// same images as (hi, lo) TF32 pairs for the split (fp32-grade) contractions of transforms_tc3.cu
__global__ void ximage_hilo_kernel(float* __restrict__ hi, float* __restrict__ lo, int kind, int R, int K, int Nt, int off, int m) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * K) return;
  const int r = idx / K, k = idx % K;
  int mode = -1, pos = 0, sine = 0;
  if (kind == 0) { pos = k + off; if (r < m) mode = r; else if (r >= SIN0 && r < SIN0 + m) { mode = r - SIN0; sine = 1; } }
  else { pos = r + off; sine = (kind == 2); if (k < m) mode = k; }
  double v = 0.0;
  if (mode >= 0) {
    double s, c;
    sincospi(2.0 * (double)(((long long)mode * pos) % Nt) / (double)Nt, &s, &c);
    v = sine ? s : c;
  }
  const float h = to_tf32((float)v);
  hi[swz_off(R, r, k) / 4] = h;
  lo[swz_off(R, r, k) / 4] = to_tf32((float)(v - (double)h));
}

// in-place split of a staged data operand: every element becomes its lo part (see transforms_tc3.cu)
__device__ __forceinline__ void split_in_place(float4* blk, int n4, int tid, int nthr) {
  for (int i = tid; i < n4; i += nthr) {
    float4 v = blk[i];
    v.x = to_tf32(v.x - tf32_trunc(v.x));
    v.y = to_tf32(v.y - tf32_trunc(v.y));
    v.z = to_tf32(v.z - tf32_trunc(v.z));
    v.w = to_tf32(v.w - tf32_trunc(v.w));
    blk[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// analysis x stage
// ------------------------------------------------------------------------------------------------
constexpr int AX_THREADS = 192;
constexpr int X3_SPLIT_THREADS = 128;
constexpr int AX_STAGES = 4;
constexpr int E_LD = 66;
struct AXSmem { uint32_t stage[AX_STAGES], a, alo, e, bars, tmem_slot, total; };
__host__ __device__ inline AXSmem ax_smem(int Nx, bool x3 = false) {
  AXSmem s; uint32_t o = 0;
  for (int i = 0; i < AX_STAGES; ++i) { s.stage[i] = o; o += 2 * Nx * 128; }
  s.a = o; o += (Nx / 32) * 64 * 128;
  s.alo = o; if (x3) o += (Nx / 32) * 64 * 128;
  s.e = o; o += 64 * E_LD * 4 + 128;
  o = (o + 127) & ~127u;
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { AXB_FULL = 0, AXB_EMPTY = AX_STAGES, AXB_DFULL = 2 * AX_STAGES, AXB_DEMPTY = 2 * AX_STAGES + 2,
       AXB_HIDONE = 2 * AX_STAGES + 4, AXB_LOREADY = 3 * AX_STAGES + 4 };

// X3: split contraction A.B = Ahi.Braw + Alo.Braw + Ahi.Blo; the data stage is rewritten in place as its lo part by
// four extra warps between the second and third pass.
template <bool X3>
__global__ void __launch_bounds__(AX_THREADS + (X3 ? X3_SPLIT_THREADS : 0), 1)
analysis_x_tc_kernel(const __grid_constant__ CUtensorMap tmap_t2, float2* __restrict__ X, const float* __restrict__ Aimg,
                     const float* __restrict__ Aloimg, int Nx, int mx, int P, int ntn, long long ntiles, float comp) {
  constexpr int NTHR = AX_THREADS + (X3 ? X3_SPLIT_THREADS : 0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const AXSmem L = ax_smem(Nx, X3);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box = (uint32_t)Nx * 128;

  {
    const float4* s1 = reinterpret_cast<const float4*>(Aimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.a);
    for (int i = threadIdx.x; i < (Nx / 32) * 64 * 128 / 16; i += NTHR) d1[i] = __ldg(s1 + i);
    if (X3) {
      const float4* s2 = reinterpret_cast<const float4*>(Aloimg);
      float4* d2 = reinterpret_cast<float4*>(sm + L.alo);
      for (int i = threadIdx.x; i < (Nx / 32) * 64 * 128 / 16; i += NTHR) d2[i] = __ldg(s2 + i);
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < AX_STAGES; ++i) {
      mbar_init(bar(AXB_FULL + i), 1); mbar_init(bar(AXB_EMPTY + i), 1);
      mbar_init(bar(AXB_HIDONE + i), 1); mbar_init(bar(AXB_LOREADY + i), X3_SPLIT_THREADS);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(AXB_DFULL + i), 1); mbar_init(bar(AXB_DEMPTY + i), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);

  if (warp == 0) {
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      long long it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = (int)(it % AX_STAGES);
        mbar_wait(bar(AXB_EMPTY + s), (uint32_t)(((it / AX_STAGES) & 1) ^ 1));
        mbar_arrive_expect_tx(bar(AXB_FULL + s), 2 * box);
        const long long bc = t / ntn;
        const int n0 = (int)(t - bc * ntn) * NT;
        tma_load_2d(base + L.stage[s], &tmap_t2, bar(AXB_FULL + s), n0, (int)(bc * Nx));
        tma_load_2d(base + L.stage[s] + box, &tmap_t2, bar(AXB_FULL + s), n0 + 32, (int)(bc * Nx));
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      constexpr uint32_t id = idesc(64, NT, 0, 1);
      if (!X3) {
        long long it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
          const int s = (int)(it % AX_STAGES), d = (int)(it & 1);
          mbar_wait(bar(AXB_FULL + s), (uint32_t)((it / AX_STAGES) & 1));
          mbar_wait(bar(AXB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
          tc_fence_after();
          for (int k = 0; k < Nx / 8; ++k)
            mma_ss(tmem + d * NT, umma_desc(base + L.a + (k >> 2) * (64 * 128) + (k & 3) * 32, 1024),
                   desc_mn(base + L.stage[s] + k * 1024, box), id, k ? 1u : 0u);
          mma_commit(bar(AXB_EMPTY + s));
          mma_commit(bar(AXB_DFULL + d));
        }
      } else {
        const long long n_my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        auto hi_pass = [&](long long it) {
          const int s = (int)(it % AX_STAGES), d = (int)(it & 1);
          mbar_wait(bar(AXB_FULL + s), (uint32_t)((it / AX_STAGES) & 1));
          mbar_wait(bar(AXB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
          tc_fence_after();
          for (int k = 0; k < Nx / 8; ++k)
            mma_ss(tmem + d * NT, umma_desc(base + L.a + (k >> 2) * (64 * 128) + (k & 3) * 32, 1024),
                   desc_mn(base + L.stage[s] + k * 1024, box), id, k ? 1u : 0u);
          for (int k = 0; k < Nx / 8; ++k)
            mma_ss(tmem + d * NT, umma_desc(base + L.alo + (k >> 2) * (64 * 128) + (k & 3) * 32, 1024),
                   desc_mn(base + L.stage[s] + k * 1024, box), id, 1u);
          mma_commit(bar(AXB_HIDONE + s));
        };
        auto lo_pass = [&](long long it) {
          const int s = (int)(it % AX_STAGES), d = (int)(it & 1);
          mbar_wait(bar(AXB_LOREADY + s), (uint32_t)((it / AX_STAGES) & 1));
          tc_fence_after();
          for (int k = 0; k < Nx / 8; ++k)
            mma_ss(tmem + d * NT, umma_desc(base + L.a + (k >> 2) * (64 * 128) + (k & 3) * 32, 1024),
                   desc_mn(base + L.stage[s] + k * 1024, box), id, 1u);
          mma_commit(bar(AXB_EMPTY + s));
          mma_commit(bar(AXB_DFULL + d));
        };
        if (n_my > 0) hi_pass(0);
        for (long long it = 0; it < n_my; ++it) {
          if (it + 1 < n_my) hi_pass(it + 1);
          lo_pass(it);
        }
      }
    }
  } else if (X3 && warp >= AX_THREADS / 32) {
    const int tid = threadIdx.x - AX_THREADS;
    long long it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = (int)(it % AX_STAGES);
      mbar_wait(bar(AXB_HIDONE + s), (uint32_t)((it / AX_STAGES) & 1));
      split_in_place(reinterpret_cast<float4*>(sm + L.stage[s]), (int)(2 * box / 16), tid, X3_SPLIT_THREADS);
      fence_proxy_async();
      mbar_arrive(bar(AXB_LOREADY + s));
    }
  } else {
    const int q = warp & 3;
    const int et = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* E = reinterpret_cast<float*>(sm + L.e);
    long long it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int d = (int)(it & 1);
      mbar_wait(bar(AXB_DFULL + d), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[64];
      tmem_ld32(tmem + d * NT + lane_base, v);
      tmem_ld32(tmem + d * NT + lane_base + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(AXB_DEMPTY + d));
      epi_bar_sync();                                  // previous tile's readers of E are done
      if (lane < 16) {
        float* er = E + (16 * q + lane) * E_LD;
#pragma unroll
        for (int c = 0; c < 64; c += 2) *reinterpret_cast<float2*>(er + c) = make_float2(v[c], v[c + 1]);
      }
      epi_bar_sync();
      const long long bc = t / ntn;
      const int p0 = (int)(t - bc * ntn) * (NT / 2);
      float2* outp = X + (size_t)bc * mx * P;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = et + 128 * i, kx = idx >> 5, j = idx & 31;
        if (kx < mx && p0 + j < P) {
          const float2 c = *reinterpret_cast<const float2*>(E + kx * E_LD + 2 * j);
          const float2 s = *reinterpret_cast<const float2*>(E + (SIN0 + kx) * E_LD + 2 * j);
          outp[(size_t)kx * P + p0 + j] = make_float2((c.x + s.y) * comp, (c.y - s.x) * comp);   // comp = 1 on the split route
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------
// synthesis x stage
// ------------------------------------------------------------------------------------------------
constexpr int SX_EPI = 8;
constexpr int SX_THREADS = 32 * (2 + SX_EPI);
constexpr int SX_STAGES = 4;
constexpr int SX_LD = 36;
struct SXSmem { uint32_t stage[SX_STAGES], ac, as, aclo, aslo, stg, bars, tmem_slot, total; };
__host__ __device__ inline SXSmem sx_smem(bool x3 = false) {
  SXSmem s; uint32_t o = 0;
  for (int i = 0; i < SX_STAGES; ++i) { s.stage[i] = o; o += 2 * 32 * 128; }
  s.ac = o; o += 128 * 128;
  s.as = o; o += 128 * 128;
  s.aclo = o; if (x3) o += 128 * 128;
  s.aslo = o; if (x3) o += 128 * 128;
  s.stg = o; o += SX_EPI * 32 * SX_LD * 4;
  o = (o + 127) & ~127u;
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { SXB_FULL = 0, SXB_EMPTY = SX_STAGES, SXB_DFULL = 2 * SX_STAGES, SXB_DEMPTY = 2 * SX_STAGES + 2,
       SXB_HIDONE = 2 * SX_STAGES + 4, SXB_LOREADY = 3 * SX_STAGES + 4 };

// Z' = Y * mode_scale[k] * row_scale[bc]   (StabilityProjection's mask and energy factor), into the workspace
__global__ void scale_modes_kernel(const float2* __restrict__ Y, float2* __restrict__ Z, const float* __restrict__ mode_scale,
                                   const float* __restrict__ row_scale, long long rs_stride, long long M, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long bc = i / M, k = i - bc * M;
  float s = mode_scale ? __ldg(mode_scale + k) : 1.f;
  if (row_scale) s *= __ldg(row_scale + bc * rs_stride);
  const float2 v = __ldg(Y + i);
  Z[i] = make_float2(v.x * s, v.y * s);
}

template <bool X3>
__global__ void __launch_bounds__(SX_THREADS + (X3 ? X3_SPLIT_THREADS : 0), 1)
synthesis_x_tc_kernel(const __grid_constant__ CUtensorMap tmap_z, float* __restrict__ U, const float* __restrict__ Acimg,
                      const float* __restrict__ Asimg, const float* __restrict__ Acloimg, const float* __restrict__ Asloimg,
                      int Nx, int mx, int P, int ntn, int nmt, long long ntiles, float comp) {
  constexpr int NTHR = SX_THREADS + (X3 ? X3_SPLIT_THREADS : 0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const SXSmem L = sx_smem(X3);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t box = 32 * 128;
  {
    const float4* s1 = reinterpret_cast<const float4*>(Acimg);
    const float4* s2 = reinterpret_cast<const float4*>(Asimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.ac);
    float4* d2 = reinterpret_cast<float4*>(sm + L.as);
    for (int i = threadIdx.x; i < 128 * 128 / 16; i += NTHR) { d1[i] = __ldg(s1 + i); d2[i] = __ldg(s2 + i); }
    if (X3) {
      const float4* s3 = reinterpret_cast<const float4*>(Acloimg);
      const float4* s4 = reinterpret_cast<const float4*>(Asloimg);
      float4* d3 = reinterpret_cast<float4*>(sm + L.aclo);
      float4* d4 = reinterpret_cast<float4*>(sm + L.aslo);
      for (int i = threadIdx.x; i < 128 * 128 / 16; i += NTHR) { d3[i] = __ldg(s3 + i); d4[i] = __ldg(s4 + i); }
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < SX_STAGES; ++i) {
      mbar_init(bar(SXB_FULL + i), 1); mbar_init(bar(SXB_EMPTY + i), 1);
      mbar_init(bar(SXB_HIDONE + i), 1); mbar_init(bar(SXB_LOREADY + i), X3_SPLIT_THREADS);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(SXB_DFULL + i), 1); mbar_init(bar(SXB_DEMPTY + i), SX_EPI); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const long long per_mt = ntiles / nmt;

  if (warp == 0) {
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      long long it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = (int)(it % SX_STAGES);
        mbar_wait(bar(SXB_EMPTY + s), (uint32_t)(((it / SX_STAGES) & 1) ^ 1));
        mbar_arrive_expect_tx(bar(SXB_FULL + s), 2 * box);
        const long long r = t % per_mt, bc = r / ntn;
        const int n0 = (int)(r - bc * ntn) * NT;
        tma_load_2d(base + L.stage[s], &tmap_z, bar(SXB_FULL + s), n0, (int)(bc * mx));
        tma_load_2d(base + L.stage[s] + box, &tmap_z, bar(SXB_FULL + s), n0 + 32, (int)(bc * mx));
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      constexpr uint32_t id = idesc(128, NT, 0, 1);
      if (!X3) {
        long long it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
          const int s = (int)(it % SX_STAGES), d = (int)(it & 1);
          mbar_wait(bar(SXB_FULL + s), (uint32_t)((it / SX_STAGES) & 1));
          mbar_wait(bar(SXB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = desc_mn(base + L.stage[s] + k * 1024, box);
            mma_ss(tmem + d * 128, umma_desc(base + L.ac + k * 32, 1024), bd, id, k ? 1u : 0u);
            mma_ss(tmem + d * 128 + NT, umma_desc(base + L.as + k * 32, 1024), bd, id, k ? 1u : 0u);
          }
          mma_commit(bar(SXB_EMPTY + s));
          mma_commit(bar(SXB_DFULL + d));
        }
      } else {
        const long long n_my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        auto hi_pass = [&](long long it) {
          const int s = (int)(it % SX_STAGES), d = (int)(it & 1);
          mbar_wait(bar(SXB_FULL + s), (uint32_t)((it / SX_STAGES) & 1));
          mbar_wait(bar(SXB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = desc_mn(base + L.stage[s] + k * 1024, box);
            mma_ss(tmem + d * 128, umma_desc(base + L.ac + k * 32, 1024), bd, id, k ? 1u : 0u);
            mma_ss(tmem + d * 128 + NT, umma_desc(base + L.as + k * 32, 1024), bd, id, k ? 1u : 0u);
            mma_ss(tmem + d * 128, umma_desc(base + L.aclo + k * 32, 1024), bd, id, 1u);
            mma_ss(tmem + d * 128 + NT, umma_desc(base + L.aslo + k * 32, 1024), bd, id, 1u);
          }
          mma_commit(bar(SXB_HIDONE + s));
        };
        auto lo_pass = [&](long long it) {
          const int s = (int)(it % SX_STAGES), d = (int)(it & 1);
          mbar_wait(bar(SXB_LOREADY + s), (uint32_t)((it / SX_STAGES) & 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = desc_mn(base + L.stage[s] + k * 1024, box);
            mma_ss(tmem + d * 128, umma_desc(base + L.ac + k * 32, 1024), bd, id, 1u);
            mma_ss(tmem + d * 128 + NT, umma_desc(base + L.as + k * 32, 1024), bd, id, 1u);
          }
          mma_commit(bar(SXB_EMPTY + s));
          mma_commit(bar(SXB_DFULL + d));
        };
        if (n_my > 0) hi_pass(0);
        for (long long it = 0; it < n_my; ++it) {
          if (it + 1 < n_my) hi_pass(it + 1);
          lo_pass(it);
        }
      }
    }
  } else if (X3 && warp >= SX_THREADS / 32) {
    const int tid = threadIdx.x - SX_THREADS;
    long long it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = (int)(it % SX_STAGES);
      mbar_wait(bar(SXB_HIDONE + s), (uint32_t)((it / SX_STAGES) & 1));
      split_in_place(reinterpret_cast<float4*>(sm + L.stage[s]), (int)(2 * box / 16), tid, X3_SPLIT_THREADS);
      fence_proxy_async();
      mbar_arrive(bar(SXB_LOREADY + s));
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stg) + (warp - 2) * 32 * SX_LD;
    const int rsub = lane >> 3, c4 = lane & 7;
    const long long rowlen = 2LL * P;
    long long it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int d = (int)(it & 1);
      mbar_wait(bar(SXB_DFULL + d), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float c[32], s[32];
      tmem_ld32(tmem + d * 128 + lane_base + half * 32, c);
      tmem_ld32(tmem + d * 128 + NT + lane_base + half * 32, s);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(SXB_DEMPTY + d));
      // (cos + i sin)(re + i im): re' = C re - S im, im' = C im + S re   (columns 2j = re, 2j+1 = im)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 o;
        o.x = (c[4 * j] - s[4 * j + 1]) * comp;
        o.y = (c[4 * j + 1] + s[4 * j]) * comp;
        o.z = (c[4 * j + 2] - s[4 * j + 3]) * comp;
        o.w = (c[4 * j + 3] + s[4 * j + 2]) * comp;
        *reinterpret_cast<float4*>(st + lane * SX_LD + 4 * j) = o;
      }
      __syncwarp();
      const int mt = (int)(t / per_mt);
      const long long r = t % per_mt, bc = r / ntn;
      const int n0 = (int)(r - bc * ntn) * NT + half * 32;
      float* ub = U + ((size_t)bc * Nx + (size_t)mt * 128 + 32 * q) * rowlen + n0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int row = 4 * g + rsub;
        if (n0 + 4 * c4 < rowlen)
          *reinterpret_cast<float4*>(ub + (size_t)row * rowlen + 4 * c4) = *reinterpret_cast<const float4*>(st + row * SX_LD + 4 * c4);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace xtc

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder(const void* device_ptr);

int xtc_plan_init(Plan* p) {
  p->xtc_a_ok = p->xtc_s_ok = 0;
  p->xtc_arena = nullptr;
  const int P = p->my * p->mzh;
  const bool a_ok = p->Nx % 32 == 0 && p->Nx <= 256 && p->mx <= 32 && P % 2 == 0;
  const bool s_ok = p->Nx == 128 && p->mx <= 32 && P % 2 == 0;     // one 128-row block of x per tile
  if (!a_ok && !s_ok) return PDEPHI_OK;
  if (xtc::ax_smem(p->Nx).total + 1024 > p->smem_optin) return PDEPHI_OK;
  const size_t nA = (size_t)64 * p->Nx, nS = (size_t)p->Nx * 32;
  PDEPHI_CUDA(cudaMalloc(&p->xtc_arena, (nA + 2 * nS) * sizeof(float)));
  PDEPHI_CUDA(cudaMemset(p->xtc_arena, 0, (nA + 2 * nS) * sizeof(float)));
  float* f = (float*)p->xtc_arena;
  p->xtc_ax = f; f += nA;
  p->xtc_sc = f; f += nS;
  p->xtc_ss = f;
  const double comp = 1.0;   // the truncation bias of the raw data operand is undone in the epilogues (Plan::tc_comp)
  auto gen = [&](float* img, int kind, int R, int K) {
    const int n = R * K;
    xtc::ximage_kernel<<<(n + 255) / 256, 256>>>(img, kind, R, K, p->Nx_total, p->x_offset, p->mx, comp);
  };
  gen(p->xtc_ax, 0, 64, p->Nx);
  gen(p->xtc_sc, 1, p->Nx, 32);
  gen(p->xtc_ss, 2, p->Nx, 32);
  PDEPHI_LAUNCH_CHECK("ximage_kernel");
  if (!tensor_map_encoder(nullptr)) return PDEPHI_OK;
  p->xtc_a_ok = a_ok;
  p->xtc_s_ok = s_ok;
  // (hi, lo) images of the same twiddles for the split fp32-grade contractions
  p->xtc3_a_ok = p->xtc3_s_ok = 0;
  p->xtc3_arena = nullptr;
  const bool a3_ok = a_ok && xtc::ax_smem(p->Nx, true).total + 1024 <= p->smem_optin;
  const bool s3_ok = s_ok && xtc::sx_smem(true).total + 1024 <= p->smem_optin;
  if (!a3_ok && !s3_ok) return PDEPHI_OK;
  PDEPHI_CUDA(cudaMalloc(&p->xtc3_arena, 2 * (nA + 2 * nS) * sizeof(float)));
  PDEPHI_CUDA(cudaMemset(p->xtc3_arena, 0, 2 * (nA + 2 * nS) * sizeof(float)));
  f = (float*)p->xtc3_arena;
  for (int h = 0; h < 2; ++h) { p->xtc3_ax[h] = f; f += nA; }
  for (int h = 0; h < 2; ++h) { p->xtc3_sc[h] = f; f += nS; }
  for (int h = 0; h < 2; ++h) { p->xtc3_ss[h] = f; f += nS; }
  auto gen3 = [&](float* hi, float* lo, int kind, int R, int K) {
    const int n = R * K;
    xtc::ximage_hilo_kernel<<<(n + 255) / 256, 256>>>(hi, lo, kind, R, K, p->Nx_total, p->x_offset, p->mx);
  };
  gen3(p->xtc3_ax[0], p->xtc3_ax[1], 0, 64, p->Nx);
  gen3(p->xtc3_sc[0], p->xtc3_sc[1], 1, p->Nx, 32);
  gen3(p->xtc3_ss[0], p->xtc3_ss[1], 2, p->Nx, 32);
  PDEPHI_LAUNCH_CHECK("ximage_hilo_kernel");
  p->xtc3_a_ok = a3_ok;
  p->xtc3_s_ok = s3_ok;
  return PDEPHI_OK;
}

void xtc_plan_destroy(Plan* p) {
  if (p->xtc_arena) cudaFree(p->xtc_arena);
  p->xtc_arena = nullptr;
  if (p->xtc3_arena) cudaFree(p->xtc3_arena);
  p->xtc3_arena = nullptr;
}

static int make_row_map(CUtensorMap* m, const void* ptr, long long rows, long long cols, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = tensor_map_encoder(ptr)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PDEPHI_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return PDEPHI_OK;
}

bool xstage_analysis_tc_ok(const Plan* p) { return p->xtc_a_ok != 0; }
bool xstage_synthesis_tc_ok(const Plan* p) { return p->xtc_s_ok != 0; }
bool xstage_analysis_tc3_ok(const Plan* p) { return p->xtc3_a_ok != 0; }
bool xstage_synthesis_tc3_ok(const Plan* p) { return p->xtc3_s_ok != 0; }

// Pcols > 0: complex columns per batch item instead of my * mzh (the channel-innermost intermediate of the fused block
// backward: batch item = volume b, columns = (ky, c, kz))
static int xstage_analysis_launch(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st, bool x3, int Pcols = 0) {
  const int P = Pcols > 0 ? Pcols : p->my * p->mzh;
  const long long rows = BC * p->Nx;
  if (rows > 0x7fffffffLL) return fail(PDEPHI_ERR_UNSUPPORTED, "x stage: too many rows");
  CUtensorMap map;
  int rc = make_row_map(&map, ws, rows, 2LL * P, p->Nx);
  if (rc) return rc;
  const int ntn = (2 * P + xtc::NT - 1) / xtc::NT;
  const long long ntiles = BC * ntn;
  const size_t smem = xtc::ax_smem(p->Nx, x3).total + 1024;
  const int grid = (int)(ntiles < p->num_sms ? ntiles : p->num_sms);
  if (x3) {
    auto k = xtc::analysis_x_tc_kernel<true>;
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_ANALYSIS_X_TC3, st);
    k<<<grid, xtc::AX_THREADS + xtc::X3_SPLIT_THREADS, smem, st>>>(map, X, p->xtc3_ax[0], p->xtc3_ax[1], p->Nx, p->mx, P, ntn,
                                                                  ntiles, 1.0f);
  } else {
    auto k = xtc::analysis_x_tc_kernel<false>;
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_ANALYSIS_X_TC, st);
    k<<<grid, xtc::AX_THREADS, smem, st>>>(map, X, p->xtc_ax, nullptr, p->Nx, p->mx, P, ntn, ntiles, p->tc_comp);
  }
  PDEPHI_LAUNCH_CHECK("analysis_x_tc_kernel");
  return PDEPHI_OK;
}
int xstage_analysis_tc(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st) {
  return xstage_analysis_launch(p, ws, X, BC, st, false);
}
int xstage_analysis_tc3(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st) {
  return xstage_analysis_launch(p, ws, X, BC, st, true);
}
int xstage_analysis_cols(const Plan* p, const float2* in, float2* X, long long nb, int Pcols, bool x3, cudaStream_t st) {
  if (Pcols <= 0 || (Pcols & 1)) return fail(PDEPHI_ERR_INVALID, "x stage: column count %d", Pcols);
  return xstage_analysis_launch(p, in, X, nb, st, x3, Pcols);
}

// Z -> (optional scaling into `scratch`) -> U (workspace)
static int xstage_synthesis_launch(const Plan* p, const float2* Z, float2* ws, float2* scratch, const float* mode_scale,
                                   const float* row_scale, long long rs_stride, long long BC, cudaStream_t st, bool x3) {
  const int P = p->my * p->mzh;
  const long long M = (long long)p->mx * P;
  const float2* src = Z;
  if (mode_scale || row_scale) {
    const long long total = BC * M;
    {
      ProfScope ps(K_SCALE_MODES, st);
      xtc::scale_modes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Z, scratch, mode_scale, row_scale, rs_stride, M, total);
    }
    PDEPHI_LAUNCH_CHECK("scale_modes_kernel");
    src = scratch;
  }
  const long long rows = BC * p->mx;
  if (rows > 0x7fffffffLL) return fail(PDEPHI_ERR_UNSUPPORTED, "x stage: too many rows");
  CUtensorMap map;
  int rc = make_row_map(&map, src, rows, 2LL * P, 32);
  if (rc) return rc;
  const int ntn = (2 * P + xtc::NT - 1) / xtc::NT;
  const int nmt = p->Nx / 128;
  const long long ntiles = BC * ntn * nmt;
  const size_t smem = xtc::sx_smem(x3).total + 1024;
  const int grid = (int)(ntiles < p->num_sms ? ntiles : p->num_sms);
  if (x3) {
    auto k = xtc::synthesis_x_tc_kernel<true>;
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_SYNTHESIS_X_TC3, st);
    k<<<grid, xtc::SX_THREADS + xtc::X3_SPLIT_THREADS, smem, st>>>(map, (float*)ws, p->xtc3_sc[0], p->xtc3_ss[0], p->xtc3_sc[1],
                                                                  p->xtc3_ss[1], p->Nx, p->mx, P, ntn, nmt, ntiles, 1.0f);
  } else {
    auto k = xtc::synthesis_x_tc_kernel<false>;
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(K_SYNTHESIS_X_TC, st);
    k<<<grid, xtc::SX_THREADS, smem, st>>>(map, (float*)ws, p->xtc_sc, p->xtc_ss, nullptr, nullptr, p->Nx, p->mx, P, ntn, nmt,
                                           ntiles, p->tc_comp);
  }
  PDEPHI_LAUNCH_CHECK("synthesis_x_tc_kernel");
  return PDEPHI_OK;
}
int xstage_synthesis_tc(const Plan* p, const float2* Z, float2* ws, float2* scratch, const float* mode_scale,
                        const float* row_scale, long long rs_stride, long long BC, cudaStream_t st) {
  return xstage_synthesis_launch(p, Z, ws, scratch, mode_scale, row_scale, rs_stride, BC, st, false);
}
int xstage_synthesis_tc3(const Plan* p, const float2* Z, float2* ws, float2* scratch, const float* mode_scale,
                         const float* row_scale, long long rs_stride, long long BC, cudaStream_t st) {
  return xstage_synthesis_launch(p, Z, ws, scratch, mode_scale, row_scale, rs_stride, BC, st, true);
}

}  // namespace pdephi
