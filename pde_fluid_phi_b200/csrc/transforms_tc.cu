// K1 / K3, TF32 path: the fused (z,y) stages as tcgen05 tensor-core contractions (sm_100a).
//
// One persistent, warp-specialised CTA per SM walks the (volume, x) planes.
//
//  analysis  (plane [Ny=128][Nz] fp32, TMA-staged, double buffered)
//     GEMM1  D1[y, c]   = plane[y, z] . B1[c, z]^T        M=128 N=48 K=Nz    (c: re kz | im kz)
//     epi-1  D1 (TMEM) -> registers -> cvt.rna.tf32 -> B2 operand tile in shared memory
//     GEMM2  D2[m, c]   = A2[m, y] . B2[c, y]^T           M=64  N=40 K=128   (m: cos ky | sin ky)
//     epi-2  D2 (TMEM) -> registers -> complex combine -> T2[ky, kz] to the workspace
//
//  synthesis (plane of U[ky, kz] complex from the workspace)
//     build  By operand tile [c, k] from U (re/-im/im/re blocks), cvt.rna.tf32
//     GEMM-Y V[y, c]    = Ay[y, k] . By[c, k]^T           M=128 N=48 K=64    (k: cos ky | sin ky)
//     GEMM-Z out[y, z]  = V[y, c] (A from TMEM) . Bz[z, c]^T   M=128 N=Nz K=40
//     epi    out (TMEM) -> registers -> shared-memory transpose -> (+ addends) -> coalesced stores
//
// Twiddle operands are built once per plan, already in the 128-byte-swizzled K-major shared-memory
// image the UMMA descriptors expect, rounded to TF32 with round-to-nearest.  The x stage and the
// channel mix stay fp32.  Replaces torch.fft.rfftn / irfftn of rational_fourier.py:161,170.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pdephi {
namespace tc {
using namespace tcptx;

constexpr int ROWS_B = 48;     // rows reserved for the (re kz | im kz) operand tiles
constexpr int N1 = 48;         // GEMM1 / GEMM-Y N
constexpr int N2 = 40;         // GEMM2 N and GEMM-Z K (>= 2*mzh)
constexpr int M2 = 64;         // GEMM2 M
constexpr int SIN0 = 32;       // first sine row / column of the y-stage twiddle operands
constexpr int E_LD = 41;

// ------------------------------------------------------------------------------------------------
// constant operand images (plan creation)
// ------------------------------------------------------------------------------------------------
// kind 0: analysis B1 [48 rows c][K=Nz]      c<mzh: w cos(2 pi c z/Nz), mzh<=c<2mzh: -w sin
// kind 1: analysis A2 [64 rows m][K=Ny]      m<my: cos(2 pi m y/Ny), 32<=m<32+my: sin
// kind 2: synthesis Ay [128 rows y][K=64]    k<my: cos(2 pi k y/Ny), 32<=k<32+my: sin
// kind 3: synthesis Bz [Nz rows z][K=64]     c<mzh: w cos(2 pi c z/Nz), mzh<=c<2mzh: -w sin, else 0
__global__ void tc_image_kernel(float* __restrict__ img, int kind, int R, int K, int N, int m, int weight_mode, int nyq,
                                double inv_total, double comp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * K) return;
  const int r = idx / K, k = idx % K;
  int mode = -1, pos = 0, sine = 0;
  if (kind == 0) { pos = k; if (r < m) mode = r; else if (r < 2 * m) { mode = r - m; sine = 1; } }
  if (kind == 1) { pos = k; if (r < m) mode = r; else if (r >= SIN0 && r < SIN0 + m) { mode = r - SIN0; sine = 1; } }
  if (kind == 2) { pos = r; if (k < m) mode = k; else if (k >= SIN0 && k < SIN0 + m) { mode = k - SIN0; sine = 1; } }
  if (kind == 3) { pos = r; if (k < m) mode = k; else if (k < 2 * m) { mode = k - m; sine = 1; } }
  double v = 0.0;
  if (mode >= 0) {
    double s, c;
    sincospi(2.0 * (double)(((long long)mode * pos) % N) / (double)N, &s, &c);
    double w = 1.0;
    if (weight_mode == 1) w = ((mode == 0 || mode == nyq) ? 1.0 : 2.0) * inv_total;
    v = sine ? ((kind == 0 || kind == 3) ? -w * s : s) : ((kind == 0 || kind == 3) ? w * c : c);
    v *= comp;
  }
  img[swz_off(R, r, k) / 4] = to_tf32((float)v);
}

// ------------------------------------------------------------------------------------------------
// analysis kernel
// ------------------------------------------------------------------------------------------------
constexpr int A_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..5 epi-1 (D1 -> B2), warps 6..9 epi-2 (D2 -> T2)
struct ASmem {
  uint32_t a1[2], b1, a2, b2, e, bars, tmem_slot, total;
};
__host__ __device__ inline ASmem analysis_smem(int NKB) {
  ASmem s;
  uint32_t o = 0;
  s.a1[0] = o; o += NKB * 16384;
  s.a1[1] = o; o += NKB * 16384;
  s.b1 = o; o += NKB * ROWS_B * 128;
  s.a2 = o; o += 4 * M2 * 128;
  s.b2 = o; o += 4 * ROWS_B * 128;
  s.e = o; o += 10752;
  s.bars = o; o += 128;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { AB_FULL0 = 0, AB_FULL1, AB_EMPTY0, AB_EMPTY1, AB_D1FULL0, AB_D1FULL1, AB_D1EMPTY0, AB_D1EMPTY1, AB_B2FULL, AB_B2EMPTY,
       AB_D2FULL0, AB_D2FULL1, AB_D2EMPTY0, AB_D2EMPTY1 };

__global__ void __launch_bounds__(A_THREADS, 1)
analysis_zy_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, float2* __restrict__ T2, const float* __restrict__ B1img,
                      const float* __restrict__ A2img, int planes, int NKB, int my, int mzh, float comp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const ASmem L = analysis_smem(NKB);
  const uint32_t bar0 = base + L.bars;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup ---------------------------------------------------------------------------
  {
    const float4* s1 = reinterpret_cast<const float4*>(B1img);
    float4* d1 = reinterpret_cast<float4*>(sm + L.b1);
    for (int i = threadIdx.x; i < NKB * ROWS_B * 128 / 16; i += A_THREADS) d1[i] = __ldg(s1 + i);
    const float4* s2 = reinterpret_cast<const float4*>(A2img);
    float4* d2 = reinterpret_cast<float4*>(sm + L.a2);
    for (int i = threadIdx.x; i < 4 * M2 * 128 / 16; i += A_THREADS) d2[i] = __ldg(s2 + i);
    float4* z = reinterpret_cast<float4*>(sm + L.b2);
    for (int i = threadIdx.x; i < 4 * ROWS_B * 128 / 16; i += A_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar(AB_FULL0), 1); mbar_init(bar(AB_FULL1), 1);
    mbar_init(bar(AB_EMPTY0), 1); mbar_init(bar(AB_EMPTY1), 1);
    mbar_init(bar(AB_D1FULL0), 1); mbar_init(bar(AB_D1FULL1), 1);
    mbar_init(bar(AB_D1EMPTY0), 128); mbar_init(bar(AB_D1EMPTY1), 128);
    mbar_init(bar(AB_B2FULL), 128); mbar_init(bar(AB_B2EMPTY), 1);
    mbar_init(bar(AB_D2FULL0), 1); mbar_init(bar(AB_D2FULL1), 1);
    mbar_init(bar(AB_D2EMPTY0), 128); mbar_init(bar(AB_D2EMPTY1), 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  // D1 and D2 double buffered: GEMM1 runs a plane ahead, and the two epilogue warpgroups work on different planes
  const uint32_t tmem_d1[2] = {tmem, tmem + 64}, tmem_d2[2] = {tmem + 128, tmem + 192};

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      int it = 0;
      for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(bar(AB_EMPTY0 + s), ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(AB_FULL0 + s), (uint32_t)NKB * 16384u);
        for (int kb = 0; kb < NKB; ++kb)
          tma_load_2d(base + L.a1[s] + kb * 16384, &tmap_x, bar(AB_FULL0 + s), kb * 32, p * 128);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      constexpr uint32_t id1 = umma_idesc(128, N1), id2 = umma_idesc(M2, N2);
      const int n_my = blockIdx.x < planes ? (planes - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      auto gemm1 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(bar(AB_FULL0 + s), ph);
        mbar_wait(bar(AB_D1EMPTY0 + s), ph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ss(tmem_d1[s], umma_desc(base + L.a1[s] + kb * 16384 + k * 32, 1024),
                   umma_desc(base + L.b1 + kb * ROWS_B * 128 + k * 32, 1024), id1, (kb | k) ? 1u : 0u);
        mma_commit(bar(AB_EMPTY0 + s));
        mma_commit(bar(AB_D1FULL0 + s));
      };
      if (n_my > 0) gemm1(0);
      for (int it = 0; it < n_my; ++it) {
        if (it + 1 < n_my) gemm1(it + 1);            // z stage of the next plane overlaps this plane's epilogues
        mbar_wait(bar(AB_B2FULL), it & 1);
        mbar_wait(bar(AB_D2EMPTY0 + (it & 1)), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ss(tmem_d2[it & 1], umma_desc(base + L.a2 + kb * M2 * 128 + k * 32, 1024),
                   umma_desc(base + L.b2 + kb * ROWS_B * 128 + k * 32, 1024), id2, (kb | k) ? 1u : 0u);
        mma_commit(bar(AB_B2EMPTY));                  // B2 consumed: epi-1 of the next plane may overwrite it
        mma_commit(bar(AB_D2FULL0 + (it & 1)));
      }
    }
  } else if (warp < 6) {
    // ===== epi-1 warps (2..5), TMEM quarter q = warp % 4: D1 row y = 32q+lane -> B2[c][y] =====
    // (a separate warpgroup from epi-2: with one group doing both, epi-1 -> GEMM2 -> epi-2 of a plane was a serial chain
    //  of ~2 us per plane on four warps, longer than the plane's 1.5 us of HBM time)
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int ncol = 2 * mzh;
    int it = 0;
    for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
      mbar_wait(bar(AB_D1FULL0 + (it & 1)), (it >> 1) & 1);
      tc_fence_after();
      float v[40];
      tmem_ld32(tmem_d1[it & 1] + lane_base, v);
      tmem_ld8(tmem_d1[it & 1] + lane_base + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(AB_D1EMPTY0 + (it & 1)));
      mbar_wait(bar(AB_B2EMPTY), (it & 1) ^ 1);       // GEMM2 of the previous plane has read B2
      {
        uint8_t* b2 = sm + L.b2 + q * (ROWS_B * 128);   // k-block q holds y in [32q, 32q+32)
        const uint32_t kin = (uint32_t)(lane & 3) * 4;
#pragma unroll
        for (int c = 0; c < 40; ++c) {
          if (c < ncol) {
            const uint32_t off = (uint32_t)((c >> 3) * 1024 + (c & 7) * 128 + ((((lane >> 2) ^ (c & 7)) & 7) << 4)) + kin;
            *reinterpret_cast<float*>(b2 + off) = to_tf32(v[c] * comp);     // undo the truncation bias of the raw plane operand
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(AB_B2FULL));
    }
  } else {
    // ===== epi-2 warps (6..9): D2 rows 16q..16q+15 (lanes 0..15 of this quarter) -> E -> combine -> T2 =====
    const int q = warp & 3;
    const int et = (warp - 6) * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* E = reinterpret_cast<float*>(sm + L.e);
    int it = 0;
    for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
      mbar_wait(bar(AB_D2FULL0 + (it & 1)), (it >> 1) & 1);
      tc_fence_after();
      float v[40];
      tmem_ld32(tmem_d2[it & 1] + lane_base, v);
      tmem_ld8(tmem_d2[it & 1] + lane_base + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(AB_D2EMPTY0 + (it & 1)));
      epi_bar_sync();                                   // previous plane's readers of E are done
      if (lane < 16) {
        float* er = E + (16 * q + lane) * E_LD;
#pragma unroll
        for (int c = 0; c < 40; ++c) er[c] = v[c];
      }
      epi_bar_sync();
      float2* outp = T2 + (size_t)p * my * mzh;
      for (int idx = et; idx < my * mzh; idx += 128) {
        const int ky = idx / mzh, kz = idx - ky * mzh;
        const float cr = E[ky * E_LD + kz], ci = E[ky * E_LD + mzh + kz];
        const float sr = E[(SIN0 + ky) * E_LD + kz], si = E[(SIN0 + ky) * E_LD + mzh + kz];
        outp[idx] = make_float2(cr + si, ci - sr);      // (C - iS)(re + i im)
      }
    }
  }
  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// synthesis kernel
//   warp 0: MMA issuer + TMEM owner, warps 2..9: epilogue (TMEM quarter = warp % 4, plane parity = (warp-2)/4),
//   warps 10..13: By builders.  The builders issue all their loads of a plane before the first store (a serial
//   load->store loop was the kernel's bound).
//   Per plane the epilogue moves 64 KB out of TMEM (64 B/clk), through the shared-memory transpose (64 KB in + 64 KB
//   out at 128 B/clk) and into the store path: ~2600 clocks of SM-internal data movement against ~2900 clocks of HBM
//   time.  With all eight warps on the same plane those phases ran one after the other (measured: MMA pipeline alone
//   0.39 ms, + staging 0.69 ms, + stores 0.91 ms); two groups of four warps on alternate planes (one output
//   accumulator each) keep the TMEM port, the shared-memory port and the store path busy at the same time.
// ------------------------------------------------------------------------------------------------
constexpr int S_EPI_WARPS = 8;
constexpr int S_BUILD_WARPS = 4;
constexpr int S_BUILD_NE = 5;    // elements of U per builder thread: my*mzh <= 32*20 = 640 = 128*5
constexpr int S_THREADS = 32 * (2 + S_EPI_WARPS + S_BUILD_WARPS);
struct SSmem {
  uint32_t ay, bz, by[2], stage, bars, tmem_slot, total;
};
__host__ __device__ inline SSmem synthesis_smem(int Nz) {
  SSmem s;
  uint32_t o = 0;
  s.ay = o; o += 2 * 128 * 128;
  s.bz = o; o += 2 * Nz * 128;
  s.by[0] = o; o += 2 * ROWS_B * 128;
  s.by[1] = o; o += 2 * ROWS_B * 128;
  s.stage = o; o += S_EPI_WARPS * 32 * (Nz / 2 + 4) * 4;
  s.bars = o; o += 128;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { SB_BYFULL0 = 0, SB_BYFULL1, SB_BYEMPTY0, SB_BYEMPTY1, SB_DVFULL, SB_OUTFULL0, SB_OUTFULL1, SB_OUTEMPTY0 /*+1*/ };

__device__ __forceinline__ float4 ldg_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

template <int NZ, int NADD>   // NADD: number of fused addends (0, 1: add0 only, 2)
__global__ void __launch_bounds__(S_THREADS, 1)
synthesis_zy_tc_kernel(const float2* __restrict__ U, float* __restrict__ out, const float* __restrict__ add0,
                       const float* __restrict__ add1, const float* __restrict__ Ayimg, const float* __restrict__ Bzimg,
                       int planes, int my, int mzh, float comp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const SSmem L = synthesis_smem(NZ);
  const uint32_t bar0 = base + L.bars;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    const float4* s1 = reinterpret_cast<const float4*>(Ayimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.ay);
    for (int i = threadIdx.x; i < 2 * 128 * 128 / 16; i += S_THREADS) d1[i] = __ldg(s1 + i);
    const float4* s2 = reinterpret_cast<const float4*>(Bzimg);
    float4* d2 = reinterpret_cast<float4*>(sm + L.bz);
    for (int i = threadIdx.x; i < 2 * NZ * 128 / 16; i += S_THREADS) d2[i] = __ldg(s2 + i);
    float4* z = reinterpret_cast<float4*>(sm + L.by[0]);
    for (int i = threadIdx.x; i < 2 * 2 * ROWS_B * 128 / 16; i += S_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar(SB_BYFULL0), S_BUILD_WARPS); mbar_init(bar(SB_BYFULL1), S_BUILD_WARPS);   // one arrival per warp
    mbar_init(bar(SB_BYEMPTY0), 1); mbar_init(bar(SB_BYEMPTY1), 1);
    mbar_init(bar(SB_DVFULL), 1);
    mbar_init(bar(SB_OUTFULL0), 1); mbar_init(bar(SB_OUTFULL1), 1);
    mbar_init(bar(SB_OUTEMPTY0), S_EPI_WARPS / 2); mbar_init(bar(SB_OUTEMPTY0 + 1), S_EPI_WARPS / 2);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(base + L.tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const uint32_t tmem_v = tmem;                       // 48 columns
  const uint32_t tmem_out[2] = {tmem + 128, tmem + 256};

  if (warp == 0) {
    // ===== MMA issuer =====
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      constexpr uint32_t idy = umma_idesc(128, N1), idz = umma_idesc(128, NZ);
      int it = 0;
      for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(bar(SB_BYFULL0 + b), ph);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ss(tmem_v, umma_desc(base + L.ay + kb * 16384 + k * 32, 1024),
                   umma_desc(base + L.by[b] + kb * ROWS_B * 128 + k * 32, 1024), idy, (kb | k) ? 1u : 0u);
        mma_commit(bar(SB_BYEMPTY0 + b));
        mma_commit(bar(SB_DVFULL));
        mbar_wait(bar(SB_DVFULL), it & 1);            // V complete in TMEM before it is read as the A operand
        mbar_wait(bar(SB_OUTEMPTY0 + b), ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < N2 / 8; ++j)
          mma_ts(tmem_out[b], tmem_v + 8 * j, umma_desc(base + L.bz + (j >> 2) * NZ * 128 + (j & 3) * 32, 1024), idz,
                 j ? 1u : 0u);
        mma_commit(bar(SB_OUTFULL0 + b));
      }
    }
  } else if (warp >= 2 + S_EPI_WARPS) {
    // ===== By builders (128 threads): U[ky][kz] -> [[re, im], [-im, re]] operand tile =====
    const int bt = (warp - 2 - S_EPI_WARPS) * 32 + lane;
    const int nel = my * mzh;
    // the element -> operand-offset map is the same for every plane: compute it once
    uint32_t o_re[S_BUILD_NE], o_im[S_BUILD_NE], o_nim[S_BUILD_NE], o_re2[S_BUILD_NE];
#pragma unroll
    for (int j = 0; j < S_BUILD_NE; ++j) {
      const int idx = min(bt + j * 32 * S_BUILD_WARPS, nel - 1);
      const int ky = idx / mzh, kz = idx - ky * mzh;
      o_re[j] = swz_off(ROWS_B, kz, ky);
      o_im[j] = swz_off(ROWS_B, mzh + kz, ky);
      o_nim[j] = swz_off(ROWS_B, kz, SIN0 + ky);
      o_re2[j] = swz_off(ROWS_B, mzh + kz, SIN0 + ky);
    }
    int it = 0;
    for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
      const int b = it & 1;
      const float2* up = U + (size_t)p * nel;
      float2 u[S_BUILD_NE];
#pragma unroll
      for (int j = 0; j < S_BUILD_NE; ++j) u[j] = __ldg(up + min(bt + j * 32 * S_BUILD_WARPS, nel - 1));   // before the wait
      mbar_wait(bar(SB_BYEMPTY0 + b), ((it >> 1) & 1) ^ 1);
      uint8_t* by = sm + L.by[b];
#pragma unroll
      for (int j = 0; j < S_BUILD_NE; ++j) {
        if (bt + j * 32 * S_BUILD_WARPS < nel) {
          const float re = to_tf32(u[j].x), im = to_tf32(u[j].y);
          *reinterpret_cast<float*>(by + o_re[j]) = re;
          *reinterpret_cast<float*>(by + o_im[j]) = im;
          *reinterpret_cast<float*>(by + o_nim[j]) = -im;
          *reinterpret_cast<float*>(by + o_re2[j]) = re;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(SB_BYFULL0 + b));
    }
  } else if (warp >= 2) {
    // ===== epilogue warps: quarter q = warp % 4 (32 rows), group = (warp - 2) / 4 takes the planes of that parity =====
    constexpr int CW = NZ / 2;                 // columns per round (64 or 32); two rounds per plane
    constexpr int LD = CW + 4;                 // padded staging row
    constexpr int LPR = CW / 4;                // lanes per output row (one float4 each)
    constexpr int RPI = 32 / LPR;              // rows per warp-wide instruction (2 or 4)
    constexpr int NIT = 32 / RPI;              // 16 or 8
    constexpr int G = 8;                       // row instructions whose addend loads are issued together
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stage) + (warp - 2) * 32 * LD;
    const int rsub = lane / LPR, c4 = lane % LPR;
    const int n_my = blockIdx.x < planes ? (planes - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (int it = grp; it < n_my; it += 2) {
      const int p = blockIdx.x + it * gridDim.x;
      mbar_wait(bar(SB_OUTFULL0 + grp), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const size_t gbase = (size_t)p * 128 * NZ + (size_t)(32 * q + rsub) * NZ + r * CW + 4 * c4;
        // addend loads of the first group of rows are issued before the TMEM read, those of the next group before the
        // current group is consumed (NADD = 1: two register sets): their latency overlaps the on-chip data movement
        float4 a0v[NADD == 1 ? 2 : 1][G], a1v[G];
        auto load_addends = [&](int g0, int set) {
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const size_t g = gbase + (size_t)((g0 + j) * RPI) * NZ;
            if (NADD >= 1) a0v[set][j] = ldg_stream(add0 + g);
            if (NADD == 2) a1v[j] = ldg_stream(add1 + g);
          }
        };
        if (NADD) load_addends(0, 0);
        float v[CW];
        tmem_ld32(tmem_out[grp] + lane_base + r * CW, v);
        if (CW == 64) tmem_ld32(tmem_out[grp] + lane_base + r * CW + 32, v + (CW == 64 ? 32 : 0));
        tmem_ld_wait();
        if (r == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(SB_OUTEMPTY0 + grp));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CW / 4; ++j)
          *reinterpret_cast<float4*>(st + lane * LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int g0 = 0; g0 < NIT; g0 += G) {
          const int set = (NADD == 1) ? ((g0 / G) & 1) : 0;
          if (NADD == 1 && g0 + G < NIT) load_addends(g0 + G, set ^ 1);
          if (NADD == 2 && g0) load_addends(g0, 0);
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const int row = (g0 + j) * RPI + rsub;
            float4 o = *reinterpret_cast<const float4*>(st + row * LD + 4 * c4);
            // comp undoes the truncation bias of the TMEM-resident V operand of GEMM-Z, on the fp32 result
            if (NADD >= 1) {
              o.x = fmaf(o.x, comp, a0v[set][j].x); o.y = fmaf(o.y, comp, a0v[set][j].y);
              o.z = fmaf(o.z, comp, a0v[set][j].z); o.w = fmaf(o.w, comp, a0v[set][j].w);
            } else {
              o.x *= comp; o.y *= comp; o.z *= comp; o.w *= comp;
            }
            if (NADD == 2) { o.x += a1v[j].x; o.y += a1v[j].y; o.z += a1v[j].z; o.w += a1v[j].w; }
            __stcs(reinterpret_cast<float4*>(out + gbase + (size_t)((g0 + j) * RPI) * NZ), o);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder(const void* device_ptr);
static EncodeTiledFn encode_fn(const void* device_ptr = nullptr) { return tensor_map_encoder(device_ptr); }
// cuTensorMapEncodeTiled is a driver-API call and needs the primary context current on the calling thread; an
// autograd worker thread may not have touched the runtime yet (CUDA_ERROR_INVALID_CONTEXT).  cudaSetDevice binds it.
EncodeTiledFn tensor_map_encoder(const void* device_ptr) {
  cudaPointerAttributes attr;
  if (device_ptr && cudaPointerGetAttributes(&attr, device_ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice)
    cudaSetDevice(attr.device);
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

bool tc_supported(const Plan* p) {
  return p->tc_ok != 0;
}

// called from pdephi_plan_create (device already selected)
int tc_plan_init(Plan* p) {
  p->tc_ok = 0;
  p->tc_comp = 1.0f;
  const bool shape_ok = p->Ny == 128 && (p->Nz == 64 || p->Nz == 128) && p->my <= 32 && 2 * p->mzh <= tc::N2;
  if (!shape_ok) return PDEPHI_OK;
  const int NKB = p->Nz / 32;
  const size_t nB1 = (size_t)NKB * tc::ROWS_B * 32, nA2 = 4 * tc::M2 * 32, nAy = 2 * 128 * 32, nBz = (size_t)2 * p->Nz * 32;
  const size_t total = (2 * nB1 + nA2 + nAy + 2 * nBz) * sizeof(float);
  PDEPHI_CUDA(cudaMalloc(&p->tc_arena, total));
  PDEPHI_CUDA(cudaMemset(p->tc_arena, 0, total));
  float* f = (float*)p->tc_arena;
  p->tc_b1[0] = f; f += nB1;
  p->tc_b1[1] = f; f += nB1;
  p->tc_a2 = f; f += nA2;
  p->tc_ay = f; f += nAy;
  p->tc_bz[0] = f; f += nBz;
  p->tc_bz[1] = f; f += nBz;
  const double inv_total = 1.0 / ((double)p->Nx_total * p->Ny * p->Nz);
  const int nyq = (p->Nz % 2 == 0 && p->mzh - 1 == p->Nz / 2) ? p->Nz / 2 : -1;
  // The tensor core drops the low 13 mantissa bits of an fp32 operand it reads itself (the TMA-staged
  // plane, the TMEM-resident V): a mean relative bias of -2^-11.  The kernels undo it on the fp32 accumulator in
  // their epilogues (p->tc_comp).  Folding the factor into the constant operand instead is only right on average:
  // twiddles that are exactly 1.0 -- the whole DC mode, i.e. every mean -- round 1 + 2^-11 up to 1 + 2^-10 in TF32
  // and over-compensate by 4.9e-4 per stage, and every other entry picks up extra rounding noise (measured per
  // transform: 4.5e-4 -> see DESIGN.md).  PDEPHI_TF32_COMP=0 disables.
  const char* env = getenv("PDEPHI_TF32_COMP");
  p->tc_comp = (env && env[0] == '0') ? 1.0f : 1.0f + 1.0f / 2048.0f;
  const double comp = 1.0;
  auto gen = [&](float* img, int kind, int R, int K, int N, int m, int wm, double c) {
    const int n = R * K;
    tc::tc_image_kernel<<<(n + 255) / 256, 256>>>(img, kind, R, K, N, m, wm, nyq, inv_total, c);
  };
  gen(p->tc_b1[0], 0, tc::ROWS_B, p->Nz, p->Nz, p->mzh, 0, comp);
  gen(p->tc_b1[1], 0, tc::ROWS_B, p->Nz, p->Nz, p->mzh, 1, comp);
  gen(p->tc_a2, 1, tc::M2, 128, p->Ny, p->my, 0, 1.0);
  gen(p->tc_ay, 2, 128, 64, p->Ny, p->my, 0, 1.0);
  gen(p->tc_bz[0], 3, p->Nz, 64, p->Nz, p->mzh, 1, comp);
  gen(p->tc_bz[1], 3, p->Nz, 64, p->Nz, p->mzh, 0, comp);
  PDEPHI_LAUNCH_CHECK("tc_image_kernel");
  if (!encode_fn()) return PDEPHI_OK;   // no driver entry point: TF32 path stays off
  p->tc_ok = 1;
  return PDEPHI_OK;
}

void tc_plan_destroy(Plan* p) {
  if (p->tc_arena) cudaFree(p->tc_arena);
  p->tc_arena = nullptr;
}

int xstage_analysis(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st);
int xstage_synthesis(const Plan* p, const float2* Z, float2* ws, const float* mode_scale, const float* row_scale,
                     long long rs_stride, long long BC, cudaStream_t st);
bool xstage_analysis_tc_ok(const Plan* p);
bool xstage_synthesis_tc_ok(const Plan* p);
bool xstage_analysis_tc3_ok(const Plan* p);
bool xstage_synthesis_tc3_ok(const Plan* p);
int xstage_analysis_tc3(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st);
int xstage_synthesis_tc3(const Plan* p, const float2* Z, float2* ws, float2* scratch, const float* mode_scale,
                         const float* row_scale, long long rs_stride, long long BC, cudaStream_t st);
int xstage_analysis_tc(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st);
int xstage_synthesis_tc(const Plan* p, const float2* Z, float2* ws, float2* scratch, const float* mode_scale,
                        const float* row_scale, long long rs_stride, long long BC, cudaStream_t st);

// The x stage works on the pruned intermediate (6.6 % of an activation): running it split (three MMAs per product, the
// kernels of the fp32-grade route) costs ~20 us per call and removes one of the three TF32 roundings of the transform
// (5.8e-4 -> 4.7e-4 per transform).  OPT_TF32_XSTAGE_SPLIT (default on) selects it.
static int xstage_analysis_for_tf32(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st) {
  if (option_get(OPT_TF32_XSTAGE_SPLIT) && xstage_analysis_tc3_ok(p)) return xstage_analysis_tc3(p, ws, X, BC, st);
  return xstage_analysis_tc_ok(p) ? xstage_analysis_tc(p, ws, X, BC, st) : xstage_analysis(p, ws, X, BC, st);
}

int analysis_tc(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st,
                int stages, float2* T) {
  if (stages != ST_ALL) ws = T;
  if (!(stages & ST_ZY)) return xstage_analysis_for_tf32(p, ws, X, BC, st);
  const size_t planes = (size_t)BC * p->Nx;
  if (planes * 128 > 0x7fffffffULL) return fail(PDEPHI_ERR_UNSUPPORTED, "analysis_tc: too many rows for the tensor map");
  if (reinterpret_cast<uintptr_t>(x) & 15) return fail(PDEPHI_ERR_INVALID, "analysis_tc: x must be 16-byte aligned");
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)p->Nz, (cuuint64_t)planes * 128};
  const cuuint64_t strides[1] = {(cuuint64_t)p->Nz * sizeof(float)};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn(x)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PDEPHI_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  const int NKB = p->Nz / 32;
  const tc::ASmem L = tc::analysis_smem(NKB);
  const size_t smem = L.total + 1024;
  auto k = tc::analysis_zy_tc_kernel;
  PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(planes < (size_t)p->num_sms ? planes : (size_t)p->num_sms);
  {
    ProfScope ps(K_ANALYSIS_ZY_TC, st);
    k<<<grid, tc::A_THREADS, smem, st>>>(tmap, ws, p->tc_b1[direction], p->tc_a2, (int)planes, NKB, p->my, p->mzh, p->tc_comp);
  }
  PDEPHI_LAUNCH_CHECK("analysis_zy_tc_kernel");
  if (!(stages & ST_X)) return PDEPHI_OK;
  return xstage_analysis_for_tf32(p, ws, X, BC, st);
}

int synthesis_tc(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                 const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                 float2* ws, cudaStream_t st, int stages, float2* T) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(out) || !al16(add0) || !al16(add1)) return fail(PDEPHI_ERR_INVALID, "synthesis_tc: buffers must be 16-byte aligned");
  float2* scratch = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) +
                                              align_up((size_t)BC * p->Nx * p->my * p->mzh * sizeof(float2), 1024));
  if (stages != ST_ALL) ws = T;
  if (stages & ST_X) {
    int rc;
    if (option_get(OPT_TF32_XSTAGE_SPLIT) && xstage_synthesis_tc3_ok(p))
      rc = xstage_synthesis_tc3(p, Z, ws, scratch, mode_scale, row_scale, rs_stride, BC, st);
    else if (xstage_synthesis_tc_ok(p)) rc = xstage_synthesis_tc(p, Z, ws, scratch, mode_scale, row_scale, rs_stride, BC, st);
    else rc = xstage_synthesis(p, Z, ws, mode_scale, row_scale, rs_stride, BC, st);
    if (rc) return rc;
  }
  if (!(stages & ST_ZY)) return PDEPHI_OK;
  const size_t planes = (size_t)BC * p->Nx;
  if (planes > 0x7fffffffULL) return fail(PDEPHI_ERR_UNSUPPORTED, "synthesis_tc: too many planes");
  const tc::SSmem L = tc::synthesis_smem(p->Nz);
  const size_t smem = L.total + 1024;
  const int grid = (int)(planes < (size_t)p->num_sms ? planes : (size_t)p->num_sms);
  if (add1 && !add0) { add0 = add1; add1 = nullptr; }
  const int nadd = (add0 ? 1 : 0) + (add1 ? 1 : 0);
#define PDEPHI_SYN_LAUNCH(NZV, NA)                                                                                          \
  {                                                                                                                         \
    auto k = tc::synthesis_zy_tc_kernel<NZV, NA>;                                                                           \
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                           \
    ProfScope ps(K_SYNTHESIS_ZY_TC, st);                                                                                    \
    k<<<grid, tc::S_THREADS, smem, st>>>(ws, out, add0, add1, p->tc_ay, p->tc_bz[direction], (int)planes, p->my, p->mzh, \
                                         p->tc_comp);                                                                     \
  }
  if (p->Nz == 128) {
    if (nadd == 0) PDEPHI_SYN_LAUNCH(128, 0) else if (nadd == 1) PDEPHI_SYN_LAUNCH(128, 1) else PDEPHI_SYN_LAUNCH(128, 2)
  } else {
    if (nadd == 0) PDEPHI_SYN_LAUNCH(64, 0) else if (nadd == 1) PDEPHI_SYN_LAUNCH(64, 1) else PDEPHI_SYN_LAUNCH(64, 2)
  }
#undef PDEPHI_SYN_LAUNCH
  PDEPHI_LAUNCH_CHECK("synthesis_zy_tc_kernel");
  return PDEPHI_OK;
}

}  // namespace pdephi
