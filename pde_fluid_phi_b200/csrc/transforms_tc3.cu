// K1 / K3, fp32-grade tensor-core path: the fused (z,y) stages as *split* TF32 contractions (sm_100a).
//
// A TF32 operand keeps 10 explicit mantissa bits, so a single kind::tf32 MMA loses ~2^-11 per operand
// (transforms_tc.cu, parity gate 2e-3).  Here every fp32 operand v is carried as v = hi + lo,
//   hi = v with the low 13 mantissa bits cleared  (exactly what the tensor core reads from a raw fp32 operand),
//   lo = rna_tf32(v - hi)                          (exact difference, rounded once to TF32),
// and every contraction is issued as three MMAs accumulating into the same fp32 TMEM accumulator:
//   A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo       (dropped: A_lo.B_lo ~ 2^-21 relative)
// which is fp32-grade (parity gate 1e-5, measured ~1e-6) at tensor-core speed; the kernels stay HBM-bound.
//
//  analysis  (plane [Ny=128][Nz] fp32; TMA streams it as [128 rows][32 floats] k-blocks through a ring)
//     split  landed k-block -> its lo part, into a second small ring (splitter warps)
//     GEMM1  D1a[y, c] += blk_raw . B1hi^T                              (raw operand = hi part)
//            D1b[y, c] += blk_raw . B1lo^T + blk_lo . B1hi^T            M=128 N=48 K=32 per block
//     epi-1  D1a + D1b (TMEM) -> registers -> (hi, lo) -> two B2 operand images in shared memory
//     GEMM2  D2a[r, c]  = A2[r, y] . B2hi[c, y]^T,  D2b = A2 . B2lo^T   (A2 TMEM-resident, rows 0..63 hi | 64..127 lo)
//     epi-2  D2a + D2b, rows r and r+64 summed, complex combine -> T2[ky, kz]
//
//  synthesis (plane of U[ky, kz] complex from the workspace)
//     build  By hi / lo operand images from U
//     GEMM-Y V[y, c]    = Ayhi.Byhi + Aylo.Byhi + Ayhi.Bylo            M=128 N=48 K=64
//     split  V (TMEM) -> registers -> lo part -> Vlo (TMEM)
//     GEMM-Z out[y, z]  = V.Bzhi + V.Bzlo + Vlo.Bzhi   (A operands from TMEM)   M=128 N=Nz K=40
//     epi    out (TMEM) -> registers -> shared-memory transpose -> (+ addends) -> coalesced stores
//
// Replaces torch.fft.rfftn / irfftn of rational_fourier.py:161,170 on the fp32 parity route.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pdephi {
namespace tc3 {
using namespace tcptx;

constexpr int RB = 40;      // rows of one (re kz | im kz) operand image per k-block (>= 2*mzh)
constexpr int RC = 2 * RB;  // analysis operands keep hi (rows 0..39) and lo (rows 40..79) of a k-block together: one N = 80 MMA
                            // gives raw.hi in accumulator columns 0..39 and raw.lo in columns 40..79
constexpr int NB = 48;      // MMA N for a single 40-row image: M = 128 needs N % 16 == 0; rows 40..47 are the next image's
constexpr int NC = 40;      // accumulator columns that are read back per image
constexpr int SIN0 = 32;    // first sine row / column of the y-stage twiddle operands
constexpr int E_LD = 41;

// ------------------------------------------------------------------------------------------------
// constant operand images (plan creation): value in fp64 -> (hi, lo) TF32 pair
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double twiddle_value(int kind, int r, int k, int N, int m, int weight_mode, int nyq,
                                                double inv_total) {
  int mode = -1, pos = 0, sine = 0;
  if (kind == 0) { pos = k; if (r < m) mode = r; else if (r < 2 * m) { mode = r - m; sine = 1; } }               // B1 [c][z]
  if (kind == 1) { pos = k; if (r < m) mode = r; else if (r >= SIN0 && r < SIN0 + m) { mode = r - SIN0; sine = 1; } }   // A2 [m][y]
  if (kind == 2) { pos = r; if (k < m) mode = k; else if (k >= SIN0 && k < SIN0 + m) { mode = k - SIN0; sine = 1; } }   // Ay [y][k]
  if (kind == 3) { pos = r; if (k < m) mode = k; else if (k < 2 * m) { mode = k - m; sine = 1; } }               // Bz [z][c]
  if (mode < 0) return 0.0;
  double s, c;
  sincospi(2.0 * (double)(((long long)mode * pos) % N) / (double)N, &s, &c);
  double w = 1.0;
  if (weight_mode == 1) w = ((mode == 0 || mode == nyq) ? 1.0 : 2.0) * inv_total;
  const bool zdir = (kind == 0 || kind == 3);
  return sine ? (zdir ? -w * s : s) : (zdir ? w * c : c);
}
// swizzled K-major images with RT rows per k-block; logical rows r < R, the lo part lands `lo_row` rows below the hi part
// (lo_row = 0: separate hi / lo images)
__global__ void image_hilo_kernel(float* __restrict__ hi, float* __restrict__ lo, int kind, int R, int RT, int lo_row, int K,
                                  int N, int m, int weight_mode, int nyq, double inv_total) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * K) return;
  const int r = idx / K, k = idx % K;
  const double v = twiddle_value(kind, r, k, N, m, weight_mode, nyq, inv_total);
  const float h = to_tf32((float)v);
  const float l = to_tf32((float)(v - (double)h));
  hi[swz_off(RT, r, k) / 4] = h;
  lo[swz_off(RT, lo_row + r, k) / 4] = l;
}
// analysis y-stage operand for TMEM residency: plain row-major [128][Ny], rows 0..63 = hi of A2, rows 64..127 = lo
__global__ void a2_plain_kernel(float* __restrict__ img, int Ny, int my) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * Ny) return;
  const int r = idx / Ny, k = idx % Ny;
  const double v = twiddle_value(1, r, k, Ny, my, 0, -1, 1.0);
  const float h = to_tf32((float)v);
  img[(size_t)r * Ny + k] = h;
  img[(size_t)(64 + r) * Ny + k] = to_tf32((float)(v - (double)h));
}

// ------------------------------------------------------------------------------------------------
// analysis kernel
//   warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue, warps 6..9: splitters
//   The splitters turn every landed k-block into its lo part in a second, small ring, independently of the MMA
//   warp (an in-place split between two MMA passes put a commit -> split -> issue round trip on the critical path).
//   Large and small partial products go to separate accumulators (D1a/D1b, D2a/D2b), summed in fp32 by the
//   epilogue: the tensor core truncates after every accumulation step, and the bias that leaves is proportional
//   to the number of steps taken on the large accumulator.
// ------------------------------------------------------------------------------------------------
constexpr int RING = 6;          // raw k-block stages of 16 KB
constexpr int LRING = 3;         // lo k-block stages
constexpr int GLAG = 2;          // GEMM2 of a plane is issued GLAG k-blocks after the plane's last GEMM1 block
constexpr int A3_THREADS = 320;
constexpr int KBYTES = 128 * 128;   // one k-block: 128 rows x 32 floats
struct A3Smem {
  uint32_t ring, lring, b1, b2, e, bars, tmem_slot, total;
};
__host__ __device__ inline A3Smem a3_smem(int NKB) {
  A3Smem s;
  uint32_t o = 0;
  s.ring = o; o += RING * KBYTES;
  s.lring = o; o += LRING * KBYTES;
  s.b1 = o; o += NKB * RC * 128;            // [k-block][hi 40 rows | lo 40 rows]
  s.b2 = o; o += 2 * RC * 128;              // k-blocks 0, 1 of the y-stage data operand
  s.e = o;                                  // k-blocks 2, 3; the epi-2 exchange buffer E aliases them (B2 is dead between
  o += 21 * 1024;                           // GEMM2 and the next epi-1): max(2*RC*128, 128*E_LD*4)
  s.bars = o; o += 512;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { A3_FULL = 0, A3_EMPTY = RING, A3_LOREADY = 2 * RING, A3_LOEMPTY = 2 * RING + LRING, A3_D1FULL = 2 * RING + 2 * LRING,
       A3_D1EMPTY = A3_D1FULL + 2, A3_B2FULL = A3_D1FULL + 4, A3_D2FULL, A3_D2EMPTY };

__device__ __forceinline__ float lo_part(float v) {
  // v - hi is exact; round it to TF32 (half away from zero) with integer ops -- cvt.rna.tf32 is a slow-pipe instruction
  const float d = v - tf32_trunc(v);
  return __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(A3_THREADS, 1)
analysis_zy_tc3_kernel(const __grid_constant__ CUtensorMap tmap_x, float2* __restrict__ T2, const float* __restrict__ B1img,
                       const float* __restrict__ A2plain, int planes, int NKB, int my, int mzh) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const A3Smem L = a3_smem(NKB);
  const uint32_t bar0 = base + L.bars;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup ---------------------------------------------------------------------------
  {
    const int n1 = NKB * RC * 128 / 16;
    const float4* s1 = reinterpret_cast<const float4*>(B1img);
    float4* d1 = reinterpret_cast<float4*>(sm + L.b1);
    for (int i = threadIdx.x; i < n1; i += A3_THREADS) d1[i] = __ldg(s1 + i);
    float4* z = reinterpret_cast<float4*>(sm + L.b2);       // B2 (with E behind it): zero all of it
    const int nz = (2 * RC * 128 + 21 * 1024) / 16;
    for (int i = threadIdx.x; i < nz; i += A3_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(bar(A3_FULL + i), 1); mbar_init(bar(A3_EMPTY + i), 1); }
    for (int i = 0; i < LRING; ++i) { mbar_init(bar(A3_LOREADY + i), 128); mbar_init(bar(A3_LOEMPTY + i), 1); }
    mbar_init(bar(A3_D1FULL), 1); mbar_init(bar(A3_D1FULL + 1), 1);
    mbar_init(bar(A3_D1EMPTY), 128); mbar_init(bar(A3_D1EMPTY + 1), 128);
    mbar_init(bar(A3_B2FULL), 128);
    mbar_init(bar(A3_D2FULL), 1); mbar_init(bar(A3_D2EMPTY), 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(base + L.tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  // columns: D1[2] at 0 / 96 (a: +0..39, b: +40..79, +80..87 scratch), D2 at 192 (a: +0..39, b: +40..79), A2 at 288..415
  const uint32_t tmem_d1[2] = {tmem, tmem + 96}, tmem_d2 = tmem + 192, tmem_a2 = tmem + 288;
  if (warp >= 2 && warp < 6) {
    // y-stage twiddles become TMEM-resident: lane = row (hi rows 0..63, lo rows 64..127), column = y
    const int q = warp & 3;
    const float* row = A2plain + (size_t)(32 * q + lane) * 128;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(row + c0) + j);
        r[4 * j] = t.x; r[4 * j + 1] = t.y; r[4 * j + 2] = t.z; r[4 * j + 3] = t.w;
      }
      tmem_st32(tmem_a2 + ((uint32_t)(q * 32) << 16) + c0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int n_my = blockIdx.x < planes ? (planes - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total = (long long)n_my * NKB;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA / UTMALDG back to back
      long long j = 0;
      for (int p = blockIdx.x; p < planes; p += gridDim.x)
        for (int kb = 0; kb < NKB; ++kb, ++j) {
          const int s = (int)(j % RING);
          mbar_wait(bar(A3_EMPTY + s), (uint32_t)(((j / RING) & 1) ^ 1));
          mbar_arrive_expect_tx(bar(A3_FULL + s), (uint32_t)KBYTES);
          tma_load_2d(base + L.ring + s * KBYTES, &tmap_x, bar(A3_FULL + s), kb * 32, p * 128);
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one elected lane =====
    if (elect_one()) {
      const uint32_t issue = 1;
      constexpr uint32_t id_cat = umma_idesc(128, RC), id_one = umma_idesc(128, NB);
      const uint32_t tm = tmem;
      const uint32_t dhi = umma_desc_hi(1024);
      const uint32_t ring_lo = umma_desc_lo(base + L.ring), lring_lo = umma_desc_lo(base + L.lring);
      const uint32_t b1_lo = umma_desc_lo(base + L.b1), b2_lo = umma_desc_lo(base + L.b2);
      constexpr uint32_t KB16 = KBYTES >> 4, RC16 = (RC * 128) >> 4;
      int s = 0, sph = 0, l = 0, lph = 0, kb = 0, d = 0, dph = 1;    // ring stage / lo stage / k-block / D1 buffer and phases
      int gk = -GLAG, gph = 0;                                      // GEMM2 trails GEMM1 by GLAG k-blocks
      const int nsteps = (int)total + GLAG;
      for (int t = 0; t < nsteps; ++t) {
        if (t < (int)total) {
          mbar_wait(bar(A3_FULL + s), (uint32_t)sph);
          if (kb == 0) mbar_wait(bar(A3_D1EMPTY + d), (uint32_t)dph);
          tc_fence_after();
          const uint32_t a = ring_lo + s * KB16, al = lring_lo + l * KB16, b = b1_lo + kb * RC16;
          const uint32_t dd = tm + d * 96;
#pragma unroll
          for (int k = 0; k < 4; ++k)      // columns 0..39 += raw.hi, 40..79 += raw.lo
            mma_ss_w(dd, umma_desc_join(dhi, a + 2 * k), umma_desc_join(dhi, b + 2 * k), id_cat, (kb | k) ? 1u : 0u, issue);
          mbar_wait(bar(A3_LOREADY + l), (uint32_t)lph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)      // columns 40..79 += lo.hi (N = 48: columns 80..87 collect lo.lo rows 0..7, never read)
            mma_ss_w(dd + RB, umma_desc_join(dhi, al + 2 * k), umma_desc_join(dhi, b + 2 * k), id_one, 1u, issue);
          mma_commit_w(bar(A3_EMPTY + s), issue);
          mma_commit_w(bar(A3_LOEMPTY + l), issue);
          if (kb == NKB - 1) mma_commit_w(bar(A3_D1FULL + d), issue);
          if (++s == RING) { s = 0; sph ^= 1; }
          if (++l == LRING) { l = 0; lph ^= 1; }
          if (++kb == NKB) { kb = 0; d ^= 1; if (d == 0) dph ^= 1; }
        }
        if (gk == NKB - 1) {
          mbar_wait(bar(A3_B2FULL), (uint32_t)gph);
          mbar_wait(bar(A3_D2EMPTY), (uint32_t)(gph ^ 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 16; ++k)     // columns 0..39 = A2.B2hi, 40..79 = A2.B2lo
            mma_ts_w(tm + 192, tm + 288 + 8 * k, umma_desc_join(dhi, b2_lo + (k >> 2) * RC16 + 2 * (k & 3)), id_cat, k ? 1u : 0u, issue);
          mma_commit_w(bar(A3_D2FULL), issue);
          gph ^= 1;
          gk = 0;
        } else {
          ++gk;
        }
      }
    }
  } else if (warp >= 6) {
    // ===== splitters: landed k-block -> its lo part in the lo ring (element-wise: same swizzled layout) =====
    const int tid = threadIdx.x - 6 * 32;
    for (long long j = 0; j < total; ++j) {
      const int s = (int)(j % RING), l = (int)(j % LRING);
      mbar_wait(bar(A3_FULL + s), (uint32_t)((j / RING) & 1));
      const float4* src = reinterpret_cast<const float4*>(sm + L.ring + s * KBYTES);
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = src[tid + 128 * i];
      mbar_wait(bar(A3_LOEMPTY + l), (uint32_t)(((j / LRING) & 1) ^ 1));
      float4* dst = reinterpret_cast<float4*>(sm + L.lring + l * KBYTES);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x = lo_part(v[i].x); v[i].y = lo_part(v[i].y); v[i].z = lo_part(v[i].z); v[i].w = lo_part(v[i].w);
        dst[tid + 128 * i] = v[i];
      }
      fence_proxy_async();
      mbar_arrive(bar(A3_LOREADY + l));
    }
  } else {
    // ===== epilogue warps (2..5): TMEM quarter q = warp % 4 =====
    const int q = warp & 3;
    const int et = (warp - 2) * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* E = reinterpret_cast<float*>(sm + L.e);
    const int ncol = 2 * mzh;
    int it = 0;
    for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
      // ---- epi-1: D1a + D1b, row y = 32q+lane -> (hi, lo) -> B2hi[c][y], B2lo[c][y] --------------------
      mbar_wait(bar(A3_D1FULL + (it & 1)), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[NC], w[NC];
      tmem_ld32(tmem_d1[it & 1] + lane_base, v);
      tmem_ld8(tmem_d1[it & 1] + lane_base + 32, v + 32);
      tmem_ld32(tmem_d1[it & 1] + lane_base + RB, w);
      tmem_ld8(tmem_d1[it & 1] + lane_base + RB + 32, w + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(A3_D1EMPTY + (it & 1)));
      epi_bar_sync();                                   // previous plane's readers of E (= B2 k-blocks 2, 3) are done
      {
        uint8_t* bh = sm + L.b2 + q * (RC * 128);       // k-block q holds y in [32q, 32q+32): hi rows 0..39, lo rows 40..79
        uint8_t* bl = bh + RB * 128;
        const uint32_t kin = (uint32_t)(lane & 3) * 4;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (c < ncol) {
            const uint32_t off = (uint32_t)((c >> 3) * 1024 + (c & 7) * 128 + ((((lane >> 2) ^ (c & 7)) & 7) << 4)) + kin;
            const float x = v[c] + w[c];
            const float h = tf32_trunc(x);
            *reinterpret_cast<float*>(bh + off) = h;
            *reinterpret_cast<float*>(bl + off) = lo_part(x);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(A3_B2FULL));
      // ---- epi-2: D2a + D2b, row 32q+lane (rows 0..63: hi twiddles, 64..127: lo twiddles) -> E -> combine -> T2 ----
      mbar_wait(bar(A3_D2FULL), (uint32_t)(it & 1));    // every warp's epi-1 stores were consumed: E may overwrite B2
      tc_fence_after();
      tmem_ld32(tmem_d2 + lane_base, v);
      tmem_ld8(tmem_d2 + lane_base + 32, v + 32);
      tmem_ld32(tmem_d2 + lane_base + RB, w);
      tmem_ld8(tmem_d2 + lane_base + RB + 32, w + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(A3_D2EMPTY));
      {
        float* er = E + (32 * q + lane) * E_LD;
#pragma unroll
        for (int c = 0; c < NC; ++c) er[c] = v[c] + w[c];
      }
      epi_bar_sync();
      float2* outp = T2 + (size_t)p * my * mzh;
      for (int idx = et; idx < my * mzh; idx += 128) {
        const int ky = idx / mzh, kz = idx - ky * mzh;
        const float* ec = E + ky * E_LD;                 // cos rows: ky (hi), 64+ky (lo)
        const float* es = E + (SIN0 + ky) * E_LD;        // sin rows: 32+ky (hi), 96+ky (lo)
        const float cr = ec[kz] + ec[64 * E_LD + kz], ci = ec[mzh + kz] + ec[64 * E_LD + mzh + kz];
        const float sr = es[kz] + es[64 * E_LD + kz], si = es[mzh + kz] + es[64 * E_LD + mzh + kz];
        outp[idx] = make_float2(cr + si, ci - sr);      // (C - iS)(re + i im)
      }
    }
  }
  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// synthesis kernel
//   warp 0: MMA issuer + TMEM owner, warps 2..9: epilogue (TMEM quarter = warp % 4, column half = (warp-2)/4),
//   warps 10..13: By builders, warps 14..17: V splitters (TMEM -> lo part -> TMEM)
// ------------------------------------------------------------------------------------------------
constexpr int S_EPI_WARPS = 8;
constexpr int S_BUILD_WARPS = 4;
constexpr int S_SPLIT_WARPS = 4;
constexpr int S_BUILD_NE = 5;    // elements of U per builder thread: my*mzh <= 32*20 = 640 = 128*5
constexpr int S3_THREADS = 32 * (2 + S_EPI_WARPS + S_BUILD_WARPS + S_SPLIT_WARPS);
constexpr int ST_LD = 36;        // staging row: 32 columns + pad
struct S3Smem {
  uint32_t ayh, ayl, bzh, bzl, byh[2], byl[2], stage, bars, tmem_slot, total;
};
__host__ __device__ inline S3Smem s3_smem(int Nz) {
  S3Smem s;
  uint32_t o = 0;
  s.ayh = o; o += 2 * 128 * 128;
  s.ayl = o; o += 2 * 128 * 128;
  s.bzh = o; o += 2 * Nz * 128;
  s.bzl = o; o += 2 * Nz * 128;
  for (int i = 0; i < 2; ++i) { s.byh[i] = o; o += 2 * NB * 128; s.byl[i] = o; o += 2 * NB * 128; }
  s.stage = o; o += S_EPI_WARPS * 32 * ST_LD * 4;
  o = (o + 127u) & ~127u;
  s.bars = o; o += 128;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { S3_BYFULL = 0, S3_BYEMPTY = 2, S3_VFULL = 4, S3_VLOFULL = 6, S3_OUTFULL = 8, S3_OUTEMPTY = 10 };

__device__ __forceinline__ float4 ldg_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

template <int NZ, int NADD>   // NADD: number of fused addends (0, 1: add0 only, 2)
__global__ void __launch_bounds__(S3_THREADS, 1)
synthesis_zy_tc3_kernel(const float2* __restrict__ U, float* __restrict__ out, const float* __restrict__ add0,
                        const float* __restrict__ add1, const float* __restrict__ Ayhi, const float* __restrict__ Aylo,
                        const float* __restrict__ Bzhi, const float* __restrict__ Bzlo, int planes, int my, int mzh) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const S3Smem L = s3_smem(NZ);
  const uint32_t bar0 = base + L.bars;
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    const float4* s1 = reinterpret_cast<const float4*>(Ayhi);
    const float4* s2 = reinterpret_cast<const float4*>(Aylo);
    float4* d1 = reinterpret_cast<float4*>(sm + L.ayh);
    float4* d2 = reinterpret_cast<float4*>(sm + L.ayl);
    for (int i = threadIdx.x; i < 2 * 128 * 128 / 16; i += S3_THREADS) { d1[i] = __ldg(s1 + i); d2[i] = __ldg(s2 + i); }
    const float4* s3 = reinterpret_cast<const float4*>(Bzhi);
    const float4* s4 = reinterpret_cast<const float4*>(Bzlo);
    float4* d3 = reinterpret_cast<float4*>(sm + L.bzh);
    float4* d4 = reinterpret_cast<float4*>(sm + L.bzl);
    for (int i = threadIdx.x; i < 2 * NZ * 128 / 16; i += S3_THREADS) { d3[i] = __ldg(s3 + i); d4[i] = __ldg(s4 + i); }
    float4* z = reinterpret_cast<float4*>(sm + L.byh[0]);   // the four By images are contiguous
    for (int i = threadIdx.x; i < 4 * 2 * NB * 128 / 16; i += S3_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(S3_BYFULL + i), S_BUILD_WARPS);
      mbar_init(bar(S3_BYEMPTY + i), 1);
      mbar_init(bar(S3_VFULL + i), 1);
      mbar_init(bar(S3_VLOFULL + i), 32 * S_SPLIT_WARPS);
      mbar_init(bar(S3_OUTFULL + i), 1);
      mbar_init(bar(S3_OUTEMPTY + i), S_EPI_WARPS / 2);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(base + L.tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const uint32_t tmem_v[2] = {tmem, tmem + 64}, tmem_vlo[2] = {tmem + 128, tmem + 192};
  const uint32_t tmem_out[2] = {tmem + 256, tmem + 384};

  if (warp == 0) {
    // ===== MMA issuer: one elected lane =====
    if (elect_one()) {
      const uint32_t issue = 1;
      constexpr uint32_t idy = umma_idesc(128, NB), idz = umma_idesc(128, NZ);
      const int n_my = blockIdx.x < planes ? (planes - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      auto gemm_y = [&](int it) {
        const int b = it & 1;
        mbar_wait(bar(S3_BYFULL + b), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        const uint32_t ah[3] = {base + L.ayh, base + L.ayl, base + L.ayh};
        const uint32_t bh[3] = {base + L.byh[b], base + L.byh[b], base + L.byl[b]};
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_ss_w(tmem_v[b], umma_desc(ah[c] + kb * 16384 + k * 32, 1024), umma_desc(bh[c] + kb * NB * 128 + k * 32, 1024),
                     idy, (c | kb | k) ? 1u : 0u, issue);
        mma_commit_w(bar(S3_BYEMPTY + b), issue);
        mma_commit_w(bar(S3_VFULL + b), issue);
      };
      auto gemm_z = [&](int it) {
        const int b = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        mbar_wait(bar(S3_VLOFULL + b), ph);             // V complete and its lo part written back to TMEM
        mbar_wait(bar(S3_OUTEMPTY + b), ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < NC / 8; ++j)
          mma_ts_w(tmem_out[b], tmem_v[b] + 8 * j, umma_desc(base + L.bzh + (j >> 2) * NZ * 128 + (j & 3) * 32, 1024), idz,
                 j ? 1u : 0u, issue);
#pragma unroll
        for (int j = 0; j < NC / 8; ++j)
          mma_ts_w(tmem_out[b], tmem_v[b] + 8 * j, umma_desc(base + L.bzl + (j >> 2) * NZ * 128 + (j & 3) * 32, 1024), idz, 1u, issue);
#pragma unroll
        for (int j = 0; j < NC / 8; ++j)
          mma_ts_w(tmem_out[b], tmem_vlo[b] + 8 * j, umma_desc(base + L.bzh + (j >> 2) * NZ * 128 + (j & 3) * 32, 1024), idz, 1u, issue);
        mma_commit_w(bar(S3_OUTFULL + b), issue);
      };
      if (n_my > 0) gemm_y(0);
      for (int it = 0; it < n_my; ++it) {
        if (it + 1 < n_my) gemm_y(it + 1);              // y stage of the next plane overlaps this plane's split
        gemm_z(it);
      }
    }
  } else if (warp >= 2 + S_EPI_WARPS + S_BUILD_WARPS) {
    // ===== V splitters: lane = row y, 48 columns =====
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(bar(S3_VFULL + b), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[NB];
      tmem_ld32(tmem_v[b] + lane_base, v);
      tmem_ld16(tmem_v[b] + lane_base + 32, v + 32);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NB; ++c) v[c] = lo_part(v[c]);
      tmem_st32(tmem_vlo[b] + lane_base, v);
      tmem_st16(tmem_vlo[b] + lane_base + 32, v + 32);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(S3_VLOFULL + b));
    }
  } else if (warp >= 2 + S_EPI_WARPS) {
    // ===== By builders (128 threads): U[ky][kz] -> [[re, im], [-im, re]] operand tiles, hi and lo =====
    const int bt = (warp - 2 - S_EPI_WARPS) * 32 + lane;
    const int nel = my * mzh;
    uint32_t o_re[S_BUILD_NE], o_im[S_BUILD_NE], o_nim[S_BUILD_NE], o_re2[S_BUILD_NE];
#pragma unroll
    for (int j = 0; j < S_BUILD_NE; ++j) {
      const int idx = min(bt + j * 32 * S_BUILD_WARPS, nel - 1);
      const int ky = idx / mzh, kz = idx - ky * mzh;
      o_re[j] = swz_off(NB, kz, ky);
      o_im[j] = swz_off(NB, mzh + kz, ky);
      o_nim[j] = swz_off(NB, kz, SIN0 + ky);
      o_re2[j] = swz_off(NB, mzh + kz, SIN0 + ky);
    }
    int it = 0;
    for (int p = blockIdx.x; p < planes; p += gridDim.x, ++it) {
      const int b = it & 1;
      const float2* up = U + (size_t)p * nel;
      float2 u[S_BUILD_NE];
#pragma unroll
      for (int j = 0; j < S_BUILD_NE; ++j) u[j] = __ldg(up + min(bt + j * 32 * S_BUILD_WARPS, nel - 1));   // before the wait
      mbar_wait(bar(S3_BYEMPTY + b), (uint32_t)(((it >> 1) & 1) ^ 1));
      uint8_t* byh = sm + L.byh[b];
      uint8_t* byl = sm + L.byl[b];
#pragma unroll
      for (int j = 0; j < S_BUILD_NE; ++j) {
        if (bt + j * 32 * S_BUILD_WARPS < nel) {
          const float re = to_tf32(u[j].x), im = to_tf32(u[j].y);
          const float rl = to_tf32(u[j].x - re), il = to_tf32(u[j].y - im);
          *reinterpret_cast<float*>(byh + o_re[j]) = re;
          *reinterpret_cast<float*>(byh + o_im[j]) = im;
          *reinterpret_cast<float*>(byh + o_nim[j]) = -im;
          *reinterpret_cast<float*>(byh + o_re2[j]) = re;
          *reinterpret_cast<float*>(byl + o_re[j]) = rl;
          *reinterpret_cast<float*>(byl + o_im[j]) = il;
          *reinterpret_cast<float*>(byl + o_nim[j]) = -il;
          *reinterpret_cast<float*>(byl + o_re2[j]) = rl;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(S3_BYFULL + b));
    }
  } else if (warp >= 2) {
    // ===== epilogue warps: quarter q = warp % 4 (32 rows), group = (warp - 2) / 4 takes the planes of that parity; a plane
    //       goes out in rounds of 32 columns (see transforms_tc.cu for why the two groups work on different planes) =====
    constexpr int ROUNDS = NZ / 32;
    constexpr int G = 8;                         // 8 lanes (one float4 each) per 32-column row segment, 4 rows per instruction
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* st = reinterpret_cast<float*>(sm + L.stage) + (warp - 2) * 32 * ST_LD;
    const int rsub = lane >> 3, c4 = lane & 7;
    const int n_my = blockIdx.x < planes ? (planes - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (int it = grp; it < n_my; it += 2) {
      const int p = blockIdx.x + it * gridDim.x;
      mbar_wait(bar(S3_OUTFULL + grp), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r) {
        const size_t gbase = (size_t)p * 128 * NZ + (size_t)(32 * q + rsub) * NZ + r * 32 + 4 * c4;
        float4 a0v[G], a1v[G];
        if (NADD) {                              // addend loads are in flight during the TMEM read and the transpose
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const size_t g = gbase + (size_t)(j * 4) * NZ;
            a0v[j] = ldg_stream(add0 + g);
            if (NADD == 2) a1v[j] = ldg_stream(add1 + g);
          }
        }
        float v[32];
        tmem_ld32(tmem_out[grp] + lane_base + r * 32, v);
        tmem_ld_wait();
        if (r == ROUNDS - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(S3_OUTEMPTY + grp));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + lane * ST_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int row = j * 4 + rsub;
          float4 o = *reinterpret_cast<const float4*>(st + row * ST_LD + 4 * c4);
          if (NADD >= 1) { o.x += a0v[j].x; o.y += a0v[j].y; o.z += a0v[j].z; o.w += a0v[j].w; }
          if (NADD == 2) { o.x += a1v[j].x; o.y += a1v[j].y; o.z += a1v[j].z; o.w += a1v[j].w; }
          __stcs(reinterpret_cast<float4*>(out + gbase + (size_t)(j * 4) * NZ), o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace tc3

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder(const void* device_ptr);

int xstage_analysis(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st);
int xstage_synthesis(const Plan* p, const float2* Z, float2* ws, const float* mode_scale, const float* row_scale,
                     long long rs_stride, long long BC, cudaStream_t st);
bool xstage_analysis_tc3_ok(const Plan* p);
bool xstage_synthesis_tc3_ok(const Plan* p);
int xstage_analysis_tc3(const Plan* p, const float2* ws, float2* X, long long BC, cudaStream_t st);
int xstage_synthesis_tc3(const Plan* p, const float2* Z, float2* ws, float2* scratch, const float* mode_scale,
                         const float* row_scale, long long rs_stride, long long BC, cudaStream_t st);

bool tc3_supported(const Plan* p) { return p->tc3_ok != 0; }

// called from pdephi_plan_create (device already selected)
int tc3_plan_init(Plan* p) {
  p->tc3_ok = 0;
  p->tc3_arena = nullptr;
  const bool shape_ok = p->Ny == 128 && (p->Nz == 64 || p->Nz == 128) && p->my <= 32 && 2 * p->mzh <= tc3::NC;
  if (!shape_ok) return PDEPHI_OK;
  if (tc3::a3_smem(p->Nz / 32).total + 1024 > p->smem_optin || tc3::s3_smem(p->Nz).total + 1024 > p->smem_optin) return PDEPHI_OK;
  const int NKB = p->Nz / 32;
  const size_t nB1 = (size_t)NKB * tc3::RC * 32, nA2 = (size_t)128 * 128, nAy = 2 * 128 * 32, nBz = (size_t)2 * p->Nz * 32;
  const size_t total = (2 * nB1 + nA2 + 2 * nAy + 4 * nBz) * sizeof(float);
  PDEPHI_CUDA(cudaMalloc(&p->tc3_arena, total));
  PDEPHI_CUDA(cudaMemset(p->tc3_arena, 0, total));
  float* f = (float*)p->tc3_arena;
  for (int d = 0; d < 2; ++d) { p->tc3_b1[d] = f; f += nB1; }
  p->tc3_a2 = f; f += nA2;
  for (int h = 0; h < 2; ++h) { p->tc3_ay[h] = f; f += nAy; }
  for (int d = 0; d < 2; ++d)
    for (int h = 0; h < 2; ++h) { p->tc3_bz[d][h] = f; f += nBz; }
  const double inv_total = 1.0 / ((double)p->Nx_total * p->Ny * p->Nz);
  const int nyq = (p->Nz % 2 == 0 && p->mzh - 1 == p->Nz / 2) ? p->Nz / 2 : -1;
  auto gen = [&](float* hi, float* lo, int kind, int R, int K, int N, int m, int wm, int RT = 0, int lo_row = 0) {
    const int n = R * K;
    tc3::image_hilo_kernel<<<(n + 255) / 256, 256>>>(hi, lo, kind, R, RT ? RT : R, lo_row, K, N, m, wm, nyq, inv_total);
  };
  // analysis z operand (hi rows 0..39 | lo rows 40..79 per k-block): direction 0 = plain, 1 = c_kz / N weights
  gen(p->tc3_b1[0], p->tc3_b1[0], 0, tc3::RB, p->Nz, p->Nz, p->mzh, 0, tc3::RC, tc3::RB);
  gen(p->tc3_b1[1], p->tc3_b1[1], 0, tc3::RB, p->Nz, p->Nz, p->mzh, 1, tc3::RC, tc3::RB);
  tc3::a2_plain_kernel<<<(64 * 128 + 255) / 256, 256>>>(p->tc3_a2, p->Ny, p->my);
  gen(p->tc3_ay[0], p->tc3_ay[1], 2, 128, 64, p->Ny, p->my, 0);
  // synthesis z operand: direction 0 (FORWARD) carries c_kz / N, 1 (ADJOINT) is plain
  gen(p->tc3_bz[0][0], p->tc3_bz[0][1], 3, p->Nz, 64, p->Nz, p->mzh, 1);
  gen(p->tc3_bz[1][0], p->tc3_bz[1][1], 3, p->Nz, 64, p->Nz, p->mzh, 0);
  PDEPHI_LAUNCH_CHECK("tc3 image kernels");
  if (!tensor_map_encoder(nullptr)) return PDEPHI_OK;
  p->tc3_ok = 1;
  return PDEPHI_OK;
}

void tc3_plan_destroy(Plan* p) {
  if (p->tc3_arena) cudaFree(p->tc3_arena);
  p->tc3_arena = nullptr;
}

int analysis_tc3(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st,
                 int stages, float2* T) {
  if (stages != ST_ALL) ws = T;
  if (!(stages & ST_ZY)) return xstage_analysis_tc3_ok(p) ? xstage_analysis_tc3(p, ws, X, BC, st) : xstage_analysis(p, ws, X, BC, st);
  const size_t planes = (size_t)BC * p->Nx;
  if (planes * 128 > 0x7fffffffULL) return fail(PDEPHI_ERR_UNSUPPORTED, "analysis_tc3: too many rows for the tensor map");
  if (reinterpret_cast<uintptr_t>(x) & 15) return fail(PDEPHI_ERR_INVALID, "analysis_tc3: x must be 16-byte aligned");
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)p->Nz, (cuuint64_t)planes * 128};
  const cuuint64_t strides[1] = {(cuuint64_t)p->Nz * sizeof(float)};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = tensor_map_encoder(x)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PDEPHI_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  const int NKB = p->Nz / 32;
  const size_t smem = tc3::a3_smem(NKB).total + 1024;
  auto k = tc3::analysis_zy_tc3_kernel;
  PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(planes < (size_t)p->num_sms ? planes : (size_t)p->num_sms);
  {
    ProfScope ps(K_ANALYSIS_ZY_TC3, st);
    k<<<grid, tc3::A3_THREADS, smem, st>>>(tmap, ws, p->tc3_b1[direction], p->tc3_a2, (int)planes, NKB, p->my, p->mzh);
  }
  PDEPHI_LAUNCH_CHECK("analysis_zy_tc3_kernel");
  if (!(stages & ST_X)) return PDEPHI_OK;
  if (xstage_analysis_tc3_ok(p)) return xstage_analysis_tc3(p, ws, X, BC, st);
  return xstage_analysis(p, ws, X, BC, st);
}

int synthesis_tc3(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                  const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                  float2* ws, cudaStream_t st, int stages, float2* T) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(out) || !al16(add0) || !al16(add1)) return fail(PDEPHI_ERR_INVALID, "synthesis_tc3: buffers must be 16-byte aligned");
  float2* scratch = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) +
                                              align_up((size_t)BC * p->Nx * p->my * p->mzh * sizeof(float2), 1024));
  if (stages != ST_ALL) ws = T;
  if (stages & ST_X) {
    int rc;
    if (xstage_synthesis_tc3_ok(p)) rc = xstage_synthesis_tc3(p, Z, ws, scratch, mode_scale, row_scale, rs_stride, BC, st);
    else rc = xstage_synthesis(p, Z, ws, mode_scale, row_scale, rs_stride, BC, st);
    if (rc) return rc;
  }
  if (!(stages & ST_ZY)) return PDEPHI_OK;
  const size_t planes = (size_t)BC * p->Nx;
  if (planes > 0x7fffffffULL) return fail(PDEPHI_ERR_UNSUPPORTED, "synthesis_tc3: too many planes");
  const size_t smem = tc3::s3_smem(p->Nz).total + 1024;
  const int grid = (int)(planes < (size_t)p->num_sms ? planes : (size_t)p->num_sms);
  if (add1 && !add0) { add0 = add1; add1 = nullptr; }
  const int nadd = (add0 ? 1 : 0) + (add1 ? 1 : 0);
#define PDEPHI_SYN3_LAUNCH(NZV, NA)                                                                                     \
  {                                                                                                                     \
    auto k = tc3::synthesis_zy_tc3_kernel<NZV, NA>;                                                                     \
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                       \
    ProfScope ps(K_SYNTHESIS_ZY_TC3, st);                                                                               \
    k<<<grid, tc3::S3_THREADS, smem, st>>>(ws, out, add0, add1, p->tc3_ay[0], p->tc3_ay[1], p->tc3_bz[direction][0],    \
                                           p->tc3_bz[direction][1], (int)planes, p->my, p->mzh);                        \
  }
  if (p->Nz == 128) {
    if (nadd == 0) PDEPHI_SYN3_LAUNCH(128, 0) else if (nadd == 1) PDEPHI_SYN3_LAUNCH(128, 1) else PDEPHI_SYN3_LAUNCH(128, 2)
  } else {
    if (nadd == 0) PDEPHI_SYN3_LAUNCH(64, 0) else if (nadd == 1) PDEPHI_SYN3_LAUNCH(64, 1) else PDEPHI_SYN3_LAUNCH(64, 2)
  }
#undef PDEPHI_SYN3_LAUNCH
  PDEPHI_LAUNCH_CHECK("synthesis_zy_tc3_kernel");
  return PDEPHI_OK;
}

}  // namespace pdephi
