// Block tail of the operator models, fused and on the tensor cores:
//   forward   u = spec + (Wc h + bc) + h ;  h' = gelu(u)            (rational_fourier.py:268-271, fno3d.py:93-97)
//   backward  gu = g' gelu'(u) ;  t = gu + Wc^T gu ;  dWc = sum_s gu (x) h ;  dbc = sum_s gu
// for a width-64 model on the channels-first [B, 64, S] layout.  A tile is 128 spatial points x 64 channels,
// staged as four [64 channel rows][32 points] sub-tiles with 128-byte swizzle: that one shared-memory image is
// the MN-major A operand of the channel GEMM  D[128 points x 64] = tile^T . Wc^T  (tcgen05 kind::tf32, TMEM
// accumulator) and, in the backward, also the K-major operands of the weight-gradient GEMM accumulated in
// TMEM across all tiles of a persistent CTA.
// The tiles are staged by four loader warps with cp.async (16-byte chunks, software swizzle), not by TMA: the 64
// rows of a tile are 4*S bytes (8 MB) apart, and a TMA box that touches 64 pages for 128 bytes each tops out at
// 4.6 TB/s on a B200 whatever the swizzle mode, ring depth or CTA count (scripts/probes/tma_probe.cu), while
// warp-wide 512-byte row reads of the same tile reach the 7.3 TB/s read peak (scripts/probes/bw_probe.cu).
// Replaces cuDNN conv fwd/bwd + bias reductions + GELU fwd/bwd +
// residual / gradient-accumulation adds (about 56 GB of HBM traffic per layer) with two passes (21.5 GB each).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pdephi {
namespace blk {
using namespace tcptx;

constexpr int C = 64;            // channels (model width)
constexpr int TS = 128;          // spatial points per tile
constexpr int BOX = C * 128;     // bytes of one [64 rows][32 floats] box
constexpr int TILE = 4 * BOX;    // 32 KB
constexpr int EPI_WARPS = 16;    // 4 TMEM quarters x 4 channel groups: enough warps to hide LDS / MUFU latency
constexpr int CPW = C / (EPI_WARPS / 4);   // channels per epilogue warp (16)
#ifndef PDEPHI_LD_WARPS
#define PDEPHI_LD_WARPS 4
#endif
constexpr int LD_WARPS = PDEPHI_LD_WARPS;   // cp.async loader warps: warp w stages channel rows w, w + LD_WARPS, ... of every tile
constexpr int LD_ROWS = (64 + LD_WARPS - 1) / LD_WARPS;
constexpr int LD_THREADS = 32 * LD_WARPS;
constexpr int EPI0 = 1 + LD_WARPS;         // first epilogue warp (warp 0 issues the MMAs)
constexpr int THREADS = 32 * (EPI0 + EPI_WARPS);
constexpr int MAXP_IN = 4;       // most input channels the fused lift takes

// MN-major TF32 operand (M or N contiguous).  32-bit MN-major operands exist only in the "128B swizzle with
// 32-byte atoms" layout (UMMA LayoutType 1 = SWIZZLE_128B_BASE32B, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
// atoms of 32 elements along MN are `lbo` bytes apart, K rows are 128 B apart in groups of 4 (512 B).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of (channel row r, point s in [0,128)) inside a tile staged with 32-byte-atom swizzle (MN-major use)
__device__ __forceinline__ uint32_t tile_off(int r, int s) {
  return (uint32_t)((s >> 5) * BOX + r * 128 + (((((s & 31) >> 3) ^ (r & 3)) & 3) << 5) + (s & 7) * 4);
}
// same element in a tile with the standard 16-byte-chunk 128B swizzle (K-major use)
__device__ __forceinline__ uint32_t tile_off_k(int r, int s) {
  return (uint32_t)((s >> 5) * BOX + r * 128 + (((((s & 31) >> 2) ^ (r & 7)) & 7) << 4) + (s & 3) * 4);
}
// Exact (erf) GELU and its derivative from one exponential: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7,
// i.e. fp32 rounding level) with e^{-x^2} = e^{-u^2/2} shared with the Gaussian density of the derivative.  The
// epilogues are issue-bound (the fused forward spends 1.5 ms of issue slots on 2.7 ms of HBM time), so the two
// transcendental steps are single MUFU instructions (ex2 / rcp .approx.ftz: no denormal range fix-ups, which
// __expf / __fdividef add as 6 more instructions per element) and the 0.5 of 0.5*erfc is folded into the polynomial.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void gelu_parts(float u, float& cdf, float& e) {
  e = ex2_ftz(u * u * -0.72134752044448170368f);               // e^{-u^2/2}
  const float t = rcp_ftz(fmaf(0.23164189591177988f, fabsf(u), 1.f));   // 1 / (1 + 0.3275911 |u| / sqrt 2)
  float p = fmaf(0.5307027145f, t, -0.7265760135f);            // half the A&S coefficients: p t e = 0.5 erfc(|u| / sqrt 2)
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float half_erfc = p * t * e;
  cdf = u >= 0.f ? 1.f - half_erfc : half_erfc;                // Phi(u) = 0.5 (1 + erf(u / sqrt 2))
}
__device__ __forceinline__ float gelu_f(float u) {
  float cdf, e;
  gelu_parts(u, cdf, e);
  return u * cdf;
}
// gelu(u) and gelu'(u) from the same CDF / density
__device__ __forceinline__ float gelu_both_f(float u, float& grad) {
  float cdf, e;
  gelu_parts(u, cdf, e);
  grad = fmaf(u * 0.3989422804014327f, e, cdf);
  return u * cdf;
}
// PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE = 2: gelu'(u) of a channel PAIR (2p, 2p+1) at one point travels as one 32-bit word of two
// binary16 values, [B][32 pairs][S] words.  binary16 has the 11 significant bits of a TF32 operand -- and the one product the
// derivative enters, gu = g' gelu'(u), is a TF32 operand of every contraction of the backward -- and gelu' lies in
// [-0.13, 1.13], far inside its range (below 6e-5 in magnitude the spacing is 6e-8 absolute).  The word of pair p at a point
// sits, in the backward's staged K-major tile, in the 4-byte slot of row 2p at that point: the slot the same thread later
// overwrites with gu, so the unpacking needs no exchange between threads.
__device__ __forceinline__ uint32_t pack_du(float even, float odd) {
  const __half2 h = __floats2half2_rn(even, odd);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_du(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ float gelu_grad_f(float u) {
  float cdf, e;
  gelu_parts(u, cdf, e);
  return fmaf(u * 0.3989422804014327f, e, cdf);
}
// W image: K-major [64 rows][K = 64], element (r, k) = TRANSPOSE ? W[k][r] : W[r][k], RN-rounded to TF32
constexpr float TF32_COMP = 1.00048828125f;   // 1 + 2^-11: undoes the mean bias of the tensor core's operand truncation
__device__ __forceinline__ void build_w_image(uint8_t* dst, const float* __restrict__ W, bool transpose, int tid, int nthr,
                                              float comp) {
  for (int i = tid; i < C * C; i += nthr) {
    const int r = i >> 6, k = i & 63;
    const float v = (transpose ? __ldg(W + k * C + r) : __ldg(W + i)) * comp;
    *reinterpret_cast<float*>(dst + swz_off(C, r, k)) = to_tf32(v);
  }
}

// ------------------------------------------------------------------------------------------------
// cp.async staging
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// A warp instruction stages one channel row of a tile: lane = 16-byte chunk (4 points) of the 128-point row, i.e.
// 512 contiguous bytes of global memory.  Offset of chunk `lane` of row r inside the staged tile:
//   MN-major use: 32-byte-atom swizzle, the image CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B would produce
__device__ __forceinline__ uint32_t chunk_off_mn(int r, int lane) {
  const int j = lane >> 3, cc = lane & 7;
  return (uint32_t)(j * BOX + r * 128 + ((((cc >> 1) ^ (r & 3)) & 3) << 5) + (cc & 1) * 16);
}
//   K-major use: 16-byte-chunk swizzle (CU_TENSOR_MAP_SWIZZLE_128B)
__device__ __forceinline__ uint32_t chunk_off_k(int r, int lane) {
  const int j = lane >> 3, cc = lane & 7;
  return (uint32_t)(j * BOX + r * 128 + (((cc ^ (r & 7)) & 7) << 4));
}
// rows [16*lw, 16*lw+16) of the tile of tensor `t` ([B*64 rows][S]) at (batch b, first point s0)
template <bool KMAJOR, int NW = LD_WARPS>
__device__ __forceinline__ void stage_rows(uint32_t dst, const float* __restrict__ t, long long b, long long s0, long long S,
                                           int lw, int lane) {
  const float* src = t + ((size_t)b * C + lw) * S + s0 + lane * 4;
#pragma unroll
  for (int i = 0; i < (C + NW - 1) / NW; ++i) {
    const int r = lw + NW * i;
    if (C % NW == 0 || r < C)
      cp_async16(dst + (KMAJOR ? chunk_off_k(r, lane) : chunk_off_mn(r, lane)), src + (size_t)i * NW * S);
  }
}

// the packed derivative ([B*32 pair rows][S] words, see pack_du): pair row p lands in row 2p of the K-major tile
template <int NW>
__device__ __forceinline__ void stage_pair_rows(uint32_t dst, const float* __restrict__ t, long long b, long long s0, long long S,
                                                int lw, int lane) {
  const float* src = t + ((size_t)b * (C / 2) + lw) * S + s0 + lane * 4;
#pragma unroll
  for (int i = 0; i < (C / 2 + NW - 1) / NW; ++i) {
    const int p = lw + NW * i;
    if ((C / 2) % NW == 0 || p < C / 2) cp_async16(dst + chunk_off_k(2 * p, lane), src + (size_t)i * NW * S);
  }
}

// GENH: the block's input h = W_in x + b_in (the model's lift, rational_fourier.py:258-260) is generated by the loader warps
// from the n_in <= 4 channels of the model input instead of being staged from HBM: h0 is never materialised.  wtab: [64]
// float4 (W_in[c][0..3]) followed by [64] bias floats in shared memory.  Same arithmetic in the forward and the backward.
// this lane's four points of every input channel of tile (b, s0)
__device__ __forceinline__ void load_x(float4* xv, const float* __restrict__ xin, int n_in, long long b, long long s0, long long S,
                                       int lane) {
#pragma unroll
  for (int n = 0; n < MAXP_IN; ++n)
    xv[n] = (n < n_in) ? __ldg(reinterpret_cast<const float4*>(xin + ((size_t)b * n_in + n) * S + s0) + lane)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <bool KMAJOR, int NW = LD_WARPS>
__device__ __forceinline__ void generate_rows(uint8_t* dst, const float4* xv, const uint8_t* wtab, int lw, int lane) {
  const float* bias = reinterpret_cast<const float*>(wtab + C * 16);
#pragma unroll
  for (int i = 0; i < (C + NW - 1) / NW; ++i) {
    const int r = lw + NW * i;
    if (C % NW != 0 && r >= C) break;
    const float4 w = reinterpret_cast<const float4*>(wtab)[r];
    const float bb = bias[r];
    float4 o;
    o.x = fmaf(w.w, xv[3].x, fmaf(w.z, xv[2].x, fmaf(w.y, xv[1].x, fmaf(w.x, xv[0].x, bb))));
    o.y = fmaf(w.w, xv[3].y, fmaf(w.z, xv[2].y, fmaf(w.y, xv[1].y, fmaf(w.x, xv[0].y, bb))));
    o.z = fmaf(w.w, xv[3].z, fmaf(w.z, xv[2].z, fmaf(w.y, xv[1].z, fmaf(w.x, xv[0].z, bb))));
    o.w = fmaf(w.w, xv[3].w, fmaf(w.z, xv[2].w, fmaf(w.y, xv[1].w, fmaf(w.x, xv[0].w, bb))));
    *reinterpret_cast<float4*>(dst + (KMAJOR ? chunk_off_k(r, lane) : chunk_off_mn(r, lane))) = o;
  }
}
__device__ __forceinline__ void build_lift_table(uint8_t* wtab, const float* __restrict__ Win, const float* __restrict__ bin, int n_in,
                                                 int tid, int nthr) {
  float* w = reinterpret_cast<float*>(wtab);
  for (int i = tid; i < C * 4; i += nthr) w[i] = ((i & 3) < n_in) ? __ldg(Win + (i >> 2) * n_in + (i & 3)) : 0.f;
  for (int i = tid; i < C; i += nthr) w[C * 4 + i] = bin ? __ldg(bin + i) : 0.f;
}
// The first block needs no h0 tile at all:  u = spec + Wc h0 + bc + h0  with  h0 = W_in x + b_in  is
//   u = spec + W_eff x + b_eff,   W_eff = (Wc + I) W_in  [64 x n_in],   b_eff = (Wc + I) b_in + bc,
// three FMAs per element in the epilogue instead of a generated 64-channel operand tile and a 64 x 64 contraction (the
// generating loader warps, not HBM, set the pace of that launch: 3.7 TB/s).  Same table layout as build_lift_table; fp32.
__device__ __forceinline__ void build_lift_eff_table(uint8_t* wtab, const float* __restrict__ Win, const float* __restrict__ bin,
                                                     int n_in, const float* __restrict__ Wc, const float* __restrict__ bc, int tid,
                                                     int nthr) {
  float* w = reinterpret_cast<float*>(wtab);
  for (int i = tid; i < C * 4 + C; i += nthr) {
    const int c = i < C * 4 ? i >> 2 : i - C * 4, n = i & 3;
    float acc = 0.f;
    if (i < C * 4) {
      if (n < n_in) {
        acc = __ldg(Win + c * n_in + n);
        for (int k = 0; k < C; ++k) acc = fmaf(__ldg(Wc + c * C + k), __ldg(Win + k * n_in + n), acc);
      }
    } else {
      acc = bc ? __ldg(bc + c) : 0.f;
      if (bin) {
        acc += __ldg(bin + c);
        for (int k = 0; k < C; ++k) acc = fmaf(__ldg(Wc + c * C + k), __ldg(bin + k), acc);
      }
    }
    w[i] = acc;
  }
}
// one row of a small side tensor ([B, n, S], n <= 4: the model input x / the output gradient gout) staged plain:
// lane = 16-byte chunk of the 128-point row
__device__ __forceinline__ void stage_small_rows(uint32_t dst, const float* __restrict__ t, int n, long long b, long long s0,
                                                 long long S, int lane) {
  for (int r = 0; r < n; ++r) cp_async16(dst + r * 512 + lane * 16, t + ((size_t)b * n + r) * S + s0 + lane * 4);
}

// ------------------------------------------------------------------------------------------------
// forward
//   warp 0: MMA issuer + TMEM owner, warps 1..4: loaders, warps 5..20: epilogue
// ------------------------------------------------------------------------------------------------
constexpr int F_STAGES = 3;
// FUSED: the spectral branch is never materialised.  Instead of a staged `spec` tile the stage holds the tile's rows of
// V = (x,y)-synthesised modes, [64 channel rows][ldz <= 64 floats] (columns 2k = re, 2k+1 = im of kz = k), and the z stage
// of the synthesis runs as extra MMAs into the same accumulator:  D[128 z x 64 ch] += Bz[128 z x K] . Vtile[64 ch x K]^T.
// A tile is then one z line of the grid (Nz = 128), and V is laid out [tile][channel][ldz] so its rows are contiguous.
// PROJ: the model's output projection (nn.Linear 64 -> n_out <= 4, rational_fourier.py:274-276) of the LAST block rides in
// the epilogue: every thread reduces its 16 channels, the four channel groups meet through shared memory.  The separate
// projection kernel and its 4.3 GB read of h' disappear (h' is still written: the projection's weight gradient reads it).
constexpr int MAXP = 4;
// ZF: the z stage of the NEXT layer's forward analysis runs inside this kernel, on the block's output h' = gelu(u): the
// epilogue threads also write h' (rounded to TF32) into a K-major tile, and  D'[128 x 64 ch] = A (TMEM-resident, rows
// 2 kz + (re | im) of e^{-i theta}) . h'^T  leaves the per-axis route's z-analysed intermediate Vh [B*64][Nx*Ny][ldz] -- the
// next block's analysis then reads 1.2 GB instead of the 4.3 GB activation (mirror of the fused backward, ZA).
struct FSmem { uint32_t stage[F_STAGES], w, bz, bias, wp, red, lift, hk, bars, tmem_slot, total; };
__host__ __device__ inline FSmem fwd_smem(bool fused, bool proj = false, bool genh = false, bool zf = false) {
  FSmem s; uint32_t o = 0;
  for (int i = 0; i < F_STAGES; ++i) { s.stage[i] = o; o += fused ? TILE + 2 * BOX : 2 * TILE; }
  s.w = o; o += 2 * BOX;
  s.bz = o; o += fused ? 2 * TS * 128 : 0;
  s.hk = o; o += zf ? TILE : 0;                            // ZF: h' as a K-major [64 ch][128 points] operand tile (1024-aligned here)
  s.bias = o; o += 256;
  s.wp = o; o += proj ? C * 16 + 16 : 0;                   // [64] float4 (w_0..w_3 of channel c) + bias float4
  s.red = o; o += proj ? 2 * 4 * MAXP * TS * 4 : 0;        // [tile parity][channel group][output][point]
  s.lift = o; o += genh ? C * 16 + C * 4 : 0;              // GENH: [64] float4 lift weights + [64] bias
  s.bars = o; o += 128;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
enum { FB_FULL = 0, FB_EMPTY = F_STAGES, FB_DFULL = 2 * F_STAGES, FB_DEMPTY = 2 * F_STAGES + 2, FB_HREADY = 2 * F_STAGES + 4,
       FB_ZFFULL = 2 * F_STAGES + 5 };

// PACKDU: the instantiation for PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE = 2 (a template parameter, not a run-time test: a branch per
// element in the epilogue loop kept ptxas from interleaving the sixteen elements' erf / exp chains)
template <bool FUSED, bool PROJ, bool GENH, bool ZF = false, bool PACKDU = false>
__global__ void __launch_bounds__(THREADS, 1)
block_tail_fwd_kernel(const float* __restrict__ h_in, const float* __restrict__ spec_in, const float* __restrict__ Wc,
                      const float* __restrict__ bc, float* __restrict__ u_out, float* __restrict__ h_out,
                      long long tiles_per_batch, long long ntiles, long long S, const float* __restrict__ bzimg, int ldz,
                      int contiguous, const float* __restrict__ Wp, const float* __restrict__ bp, float* __restrict__ proj_out,
                      int NP, const float* __restrict__ Win, const float* __restrict__ bin, int n_in, int save_du,
                      const float* __restrict__ za_img, float* __restrict__ vh_out) {
  // save_du: u_out receives gelu'(u) -- all the backward needs u for -- instead of u (PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const FSmem L = fwd_smem(FUSED, PROJ, GENH, ZF);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // the truncation bias of the data operands (h; V) is undone on the accumulator in the epilogue, not in the weights
  build_w_image(sm + L.w, Wc, false, threadIdx.x, THREADS, 1.0f);
  constexpr bool LIFT_EFF = GENH && FUSED;     // first block, fused forward: no h0 tile (build_lift_eff_table)
  if (LIFT_EFF) build_lift_eff_table(sm + L.lift, Win, bin, n_in, Wc, bc, threadIdx.x, THREADS);
  else if (GENH) build_lift_table(sm + L.lift, Win, bin, n_in, threadIdx.x, THREADS);
  if (PROJ) {
    float* wp = reinterpret_cast<float*>(sm + L.wp);
    for (int i = threadIdx.x; i < C * MAXP + MAXP; i += THREADS) {
      const int c = i >> 2, n = i & 3;
      wp[i] = (n < NP) ? (c < C ? __ldg(Wp + n * C + c) : (bp ? __ldg(bp + n) : 0.f)) : 0.f;
    }
  }
  if (FUSED) {
    const float4* s1 = reinterpret_cast<const float4*>(bzimg);
    float4* d1 = reinterpret_cast<float4*>(sm + L.bz);
    const int nkb = (ldz + 31) / 32;
    for (int i = threadIdx.x; i < nkb * TS * 128 / 16; i += THREADS) d1[i] = __ldg(s1 + i);
    for (int st = 0; st < F_STAGES; ++st) {        // the chunks of a V stage beyond ldz are never written: zero them once
      float4* z = reinterpret_cast<float4*>(sm + L.stage[st] + TILE);
      for (int i = threadIdx.x; i < 2 * BOX / 16; i += THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (threadIdx.x < C) reinterpret_cast<float*>(sm + L.bias)[threadIdx.x] = bc ? __ldg(bc + threadIdx.x) : 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < F_STAGES; ++i) { mbar_init(bar(FB_FULL + i), LD_THREADS); mbar_init(bar(FB_EMPTY + i), 1 + EPI_WARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar(FB_DFULL + i), 1); mbar_init(bar(FB_DEMPTY + i), EPI_WARPS); }
    mbar_init(bar(FB_HREADY), EPI_WARPS);
    mbar_init(bar(FB_ZFFULL), 1);
    fence_barrier_init();
  }
  constexpr uint32_t TMEM_COLS = ZF ? 512 : 128;
  if (warp == 0) tmem_alloc(base + L.tmem_slot, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const uint32_t tmem_zf = tmem + 2 * C, tmem_a = tmem + 4 * C;      // ZF: accumulator D' [128 x 64], A image [128 x 128]
  if (ZF) {
    if (warp >= EPI0 && warp < EPI0 + 4) {       // four consecutive warps cover the four TMEM lane quarters
      const int q = warp & 3;
      const float* row = za_img + (size_t)(32 * q + lane) * TS;
      for (int c0 = 0; c0 < TS; c0 += 32) {
        float r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(row + c0) + j);
          r[4 * j] = t4.x; r[4 * j + 1] = t4.y; r[4 * j + 2] = t4.z; r[4 * j + 3] = t4.w;
        }
        tmem_st32(tmem_a + ((uint32_t)(q * 32) << 16) + c0, r);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // every CTA walks a CONTIGUOUS range of tiles (PDEPHI_TILE_ORDER=0: round-robin): consecutive tiles extend the same 64
  // channel rows, so the 512-byte pieces a tile writes to each row meet their neighbours in L2 within one tile time
  const long long tq = ntiles / gridDim.x, tr = ntiles % gridDim.x;
  const long long my_tiles = contiguous ? tq + (blockIdx.x < tr ? 1 : 0)
                                        : (blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  const long long tile0 = contiguous ? blockIdx.x * tq + (blockIdx.x < tr ? blockIdx.x : tr) : blockIdx.x;
  const long long tstep = contiguous ? 1 : gridDim.x;

  if (warp >= 1 && warp < EPI0) {
    // ===== loaders: h and spec (or V) tiles, up to F_STAGES - 1 tiles in flight =====
    // Order matters: a landed tile is published (FULL) BEFORE the loader blocks on the stage the next tile needs.  The
    // first version waited for that stage first, so FULL(k) could not be signalled until the stage of tile k-1 had been
    // released by its epilogue -- MMA(k) never overlapped epilogue(k-1).
    const int lw = warp - 1;
    const int cpr = ldz >> 2, nchunk = C * cpr;        // FUSED: 16-byte chunks per V row / per V tile
    // GENH: the x values of the NEXT tile are fetched while this one is generated (with the loads issued where they are
    // consumed the loader, not HBM, set the pace of the first block: 2.5 -> 3.25 ms)
    float4 xnext[MAXP_IN];
    auto fetch_x = [&](long long it) {
      const long long t = tile0 + it * tstep;
      const long long b = t / tiles_per_batch, s0 = (t - b * tiles_per_batch) * TS;
      load_x(xnext, h_in, n_in, b, s0, S, lane);
    };
    if (GENH && !LIFT_EFF && my_tiles > 0) fetch_x(0);
    auto issue = [&](long long it) {
      const long long t = tile0 + it * tstep;
      const int s = (int)(it % F_STAGES);
      float4 xv[MAXP_IN];
      if (GENH && !LIFT_EFF) {
#pragma unroll
        for (int n = 0; n < MAXP_IN; ++n) xv[n] = xnext[n];
        if (it + 1 < my_tiles) fetch_x(it + 1);
      }
      mbar_wait(bar(FB_EMPTY + s), (uint32_t)(((it / F_STAGES) & 1) ^ 1));
      const long long b = t / tiles_per_batch, s0 = (t - b * tiles_per_batch) * TS;
      if (LIFT_EFF) { if (lw == 0) stage_small_rows(base + L.stage[s], h_in, n_in, b, s0, S, lane); }   // x rows, plain
      else if (GENH) generate_rows<false>(sm + L.stage[s], xv, sm + L.lift, lw, lane);   // h_in = model input x
      else stage_rows<false>(base + L.stage[s], h_in, b, s0, S, lw, lane);
      if (FUSED) {
        // V rows of tile t are contiguous: chunk q = (channel c, chunk j of its row) -> K-major 128B-swizzled B operand
        const float* vsrc = spec_in + (size_t)t * C * ldz;
        const uint32_t vdst = base + L.stage[s] + TILE;
        for (int q = lw * 32 + lane; q < nchunk; q += LD_THREADS) {
          const int c = q / cpr, j = q - c * cpr;
          cp_async16(vdst + (uint32_t)((j >> 3) * BOX + c * 128 + ((((j & 7) ^ (c & 7)) & 7) << 4)), vsrc + (size_t)q * 4);
        }
      } else {
        stage_rows<false>(base + L.stage[s] + TILE, spec_in, b, s0, S, lw, lane);
      }
    };
    for (int pre = 0; pre < F_STAGES - 1; ++pre) {
      if (pre < my_tiles) issue(pre);
      cp_async_commit();
    }
    for (long long it = 0; it < my_tiles; ++it) {
      cp_async_wait<F_STAGES - 2>();             // every group but the newest F_STAGES - 2: tile `it` has landed
      fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar(FB_FULL + (int)(it % F_STAGES)));
      if (it + F_STAGES - 1 < my_tiles) issue(it + F_STAGES - 1);
      cp_async_commit();                         // (possibly empty: keeps the group count uniform)
    }
  } else if (warp == 0) {
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA back to back
      constexpr uint32_t id = idesc_tf32(TS, C, 1, 0);
      constexpr uint32_t idz = idesc_tf32(TS, C, 0, 0);
      const int ksteps = (ldz + 7) / 8;
      for (long long it = 0; it < my_tiles; ++it) {
        const int s = (int)(it % F_STAGES), d = (int)(it & 1);
        mbar_wait(bar(FB_FULL + s), (uint32_t)((it / F_STAGES) & 1));
        mbar_wait(bar(FB_DEMPTY + d), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        if (!LIFT_EFF) {
#pragma unroll
          for (int k = 0; k < C / 8; ++k)
            mma_ss(tmem + d * C, umma_desc_mn(base + L.stage[s] + k * 1024, BOX),
                   umma_desc(base + L.w + (k >> 2) * BOX + (k & 3) * 32, 1024), id, k ? 1u : 0u);
        }
        if (FUSED) {
          for (int k = 0; k < ksteps; ++k)         // z stage of the synthesis: += Bz[z][k] . V[ch][k]
            mma_ss(tmem + d * C, umma_desc(base + L.bz + (k >> 2) * (TS * 128) + (k & 3) * 32, 1024),
                   umma_desc(base + L.stage[s] + TILE + (k >> 2) * BOX + (k & 3) * 32, 1024), idz, (LIFT_EFF && k == 0) ? 0u : 1u);
        }
        mma_commit(bar(FB_EMPTY + s));
        mma_commit(bar(FB_DFULL + d));
        if (ZF && it > 0) {      // z analysis of the PREVIOUS tile's h' (its epilogue has written the operand tile by now)
          constexpr uint32_t id_z = idesc_tf32(TS, C, 0, 0);
          mbar_wait(bar(FB_HREADY), (uint32_t)((it - 1) & 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < TS / 8; ++k)
            mma_ts(tmem_zf, tmem_a + 8 * k, umma_desc(base + L.hk + (k >> 2) * BOX + (k & 3) * 32, 1024), id_z, k ? 1u : 0u);
          mma_commit(bar(FB_ZFFULL));
        }
      }
      if (ZF && my_tiles > 0) {
        constexpr uint32_t id_z = idesc_tf32(TS, C, 0, 0);
        mbar_wait(bar(FB_HREADY), (uint32_t)((my_tiles - 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < TS / 8; ++k)
          mma_ts(tmem_zf, tmem_a + 8 * k, umma_desc(base + L.hk + (k >> 2) * BOX + (k & 3) * 32, 1024), id_z, k ? 1u : 0u);
        mma_commit(bar(FB_ZFFULL));
      }
    }
  } else {
    const int q = warp & 3, part = (warp - EPI0) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float bsr[CPW];
#pragma unroll
    for (int j = 0; j < CPW; ++j)
      bsr[j] = LIFT_EFF ? reinterpret_cast<const float*>(sm + L.lift)[C * 4 + part * CPW + j]      // b_eff
                        : reinterpret_cast<const float*>(sm + L.bias)[part * CPW + j];
    // element (channel row o = part*CPW + j, point 32q + lane): the swizzle term depends on o & 3 = j & 3 only
    uint32_t xc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) xc[r] = (uint32_t)(q * BOX + part * CPW * 128 + ((((lane >> 3) ^ r) & 3) << 5) + (lane & 7) * 4);
    // ZF: offsets of this thread's elements in the K-major (16-byte-chunk swizzle) h' tile; the swizzle term depends on o & 7
    uint32_t xk[ZF ? 8 : 1];
    if (ZF) {
#pragma unroll
      for (int r = 0; r < 8; ++r) xk[r] = (uint32_t)(q * BOX + part * CPW * 128 + ((((lane >> 2) ^ r) & 7) << 4) + (lane & 3) * 4);
    }
    // rows kz' = 32 q + lane of D' (quarters 0 and 1: ldz <= 64) of tile `tz` -> Vh [B*64 volumes][tiles of a volume][ldz]
    auto store_vh = [&](long long tz, uint32_t phase) {
      mbar_wait(bar(FB_ZFFULL), phase);
      tc_fence_after();
      float z[CPW];
      if (q < 2) {
        tmem_ld16(tmem_zf + lane_base + part * CPW, z);
        tmem_ld_wait();
      }
      tc_fence_before();
      if (q < 2 && 32 * q + lane < ldz) {
        const long long b = tz / tiles_per_batch, tib = tz - b * tiles_per_batch;
        float* vp = vh_out + (((size_t)b * C + part * CPW) * tiles_per_batch + tib) * ldz + 32 * q + lane;
        const size_t cs = (size_t)tiles_per_batch * ldz;
#pragma unroll
        for (int j = 0; j < CPW; ++j) vp[(size_t)j * cs] = z[j];
      }
    };
    for (long long it = 0; it < my_tiles; ++it) {
      const long long t = tile0 + it * tstep;
      const int s = (int)(it % F_STAGES), d = (int)(it & 1);
      mbar_wait(bar(FB_DFULL + d), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[CPW];
      tmem_ld16(tmem + d * C + lane_base + part * CPW, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(FB_DEMPTY + d));
      const long long b = t / tiles_per_batch, s0 = (t - b * tiles_per_batch) * TS;
      const uint8_t* th = sm + L.stage[s];
      const uint8_t* tsp = th + TILE;
      // The residual (and spec) operands leave the stage first and the stage is released BEFORE the GELU and the
      // stores: its lifetime is then load latency + MMA instead of load latency + MMA + the whole epilogue, which is
      // what the three-deep ring has to cover (measured: the kernel took 2.0 ms with an empty epilogue and 3.2 ms with
      // the real one at any byte count -- it was bound by this chain, not by HBM).
      float pre[CPW];
      if (LIFT_EFF) {      // this thread's point of the n_in <= 4 input channels (rows beyond n_in are never written: weight 0)
        const float* xs = reinterpret_cast<const float*>(th) + 32 * q + lane;
        const float x0 = xs[0], x1 = n_in > 1 ? xs[128] : 0.f, x2 = n_in > 2 ? xs[256] : 0.f, x3 = n_in > 3 ? xs[384] : 0.f;
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const float4 w = reinterpret_cast<const float4*>(sm + L.lift)[part * CPW + j];
          pre[j] = fmaf(w.w, x3, fmaf(w.z, x2, fmaf(w.y, x1, w.x * x0)));
        }
      } else {
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          const uint32_t off = xc[j & 3] + j * 128;
          pre[j] = *reinterpret_cast<const float*>(th + off);
          if (!FUSED) pre[j] += *reinterpret_cast<const float*>(tsp + off);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(FB_EMPTY + s));
      // ZF: the previous tile's D' must be out of TMEM (and its MMAs done with the h' tile) before this tile overwrites the tile
      if (ZF && it > 0) store_vh(t - tstep, (uint32_t)((it - 1) & 1));
      const size_t g0 = ((size_t)b * C + part * CPW) * S + s0 + 32 * q + lane;
      float* up = u_out + g0;
      float* hp = h_out + g0;
      // save_du == 2: one word per channel PAIR and point (see pack_du) -- this thread holds both channels of each of its pairs
      uint32_t* up2 = reinterpret_cast<uint32_t*>(u_out) + ((size_t)b * (C / 2) + part * (CPW / 2)) * S + s0 + 32 * q + lane;
      float du_even = 0.f;
      float pj[MAXP] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < CPW; ++j) {
        const float u = fmaf(v[j], TF32_COMP, bsr[j]) + pre[j];
        float du;
        const float hv = gelu_both_f(u, du);
        if (PACKDU) {
          if (j & 1) { __stcs(up2, pack_du(du_even, du)); up2 += S; }
          else du_even = du;
        } else {
          __stcs(up, save_du ? du : u);
        }
        *hp = hv;
        if (ZF) *reinterpret_cast<float*>(sm + L.hk + xk[j & 7] + j * 128) = to_tf32(hv);
        if (PROJ) {
          const float4 w = reinterpret_cast<const float4*>(sm + L.wp)[part * CPW + j];
          pj[0] = fmaf(w.x, hv, pj[0]); pj[1] = fmaf(w.y, hv, pj[1]);
          pj[2] = fmaf(w.z, hv, pj[2]); pj[3] = fmaf(w.w, hv, pj[3]);
        }
        up += S;
        hp += S;
      }
      if (ZF) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(FB_HREADY));
      }
      if (PROJ) {
        // two buffers by tile parity and one barrier per tile: a buffer is rewritten two tiles later, after every warp has
        // passed the next tile's barrier, i.e. after its readers are done
        float* red = reinterpret_cast<float*>(sm + L.red) + (it & 1) * (4 * MAXP * TS);
        const int pt = 32 * q + lane;
#pragma unroll
        for (int n = 0; n < MAXP; ++n) red[(part * MAXP + n) * TS + pt] = pj[n];
        asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        if (part == 0) {
          const float4 bias4 = reinterpret_cast<const float4*>(sm + L.wp)[C];
          const float bb[MAXP] = {bias4.x, bias4.y, bias4.z, bias4.w};
#pragma unroll
          for (int n = 0; n < MAXP; ++n)
            if (n < NP)
              proj_out[((size_t)b * NP + n) * S + s0 + pt] =
                  bb[n] + ((red[n * TS + pt] + red[(MAXP + n) * TS + pt]) + (red[(2 * MAXP + n) * TS + pt] + red[(3 * MAXP + n) * TS + pt]));
        }
      }
    }
  }
  if (ZF && warp >= EPI0 && my_tiles > 0) {       // the last tile's D'
    const int q = warp & 3, part = (warp - EPI0) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    mbar_wait(bar(FB_ZFFULL), (uint32_t)((my_tiles - 1) & 1));
    tc_fence_after();
    float z[CPW];
    if (q < 2) {
      tmem_ld16(tmem_zf + lane_base + part * CPW, z);
      tmem_ld_wait();
    }
    if (q < 2 && 32 * q + lane < ldz) {
      const long long tz = tile0 + (my_tiles - 1) * tstep;
      const long long b = tz / tiles_per_batch, tib = tz - b * tiles_per_batch;
      float* vp = vh_out + (((size_t)b * C + part * CPW) * tiles_per_batch + tib) * ldz + 32 * q + lane;
      const size_t cs = (size_t)tiles_per_batch * ldz;
#pragma unroll
      for (int j = 0; j < CPW; ++j) vp[(size_t)j * cs] = z[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// backward
//   warp 0: MMA issuer + TMEM owner, warps 1..4: loaders, warps 5..20: compute / epilogue
// ------------------------------------------------------------------------------------------------
constexpr int B_STAGES = 2;
// The ZA variants run three loader warps instead of four: 20 warps fit the 96-register allocation, 21 round up to 24 and cap
// the kernel at 80, where the extra epilogue of the fused z analysis spilled ~300 bytes per thread onto the critical path of
// the compute warps (3.67 / 4.04 / 3.97 ms for the first / middle / last block at 80 registers, 2.85 / 3.67 / 3.53 ms at 96).
// The forward kernels and the plain backward keep four loaders (they are loader-bound: 2.49 -> 2.62 ms with three).
// The last block's variant (GENG: its loader warps GENERATE the g' tile from gout) keeps four -- there the loaders set the pace
// (3.57 ms with three, 3.35 ms with four at 80 registers, same box; forming g' in step 1 of the compute warps instead, with
// three loaders and 96 registers: 3.43 ms).
// The last block's variant (GENG: the g' tile is generated from gout, not staged) has two forms, chosen at compile time:
// PDEPHI_GENG_STEP1 = 0: its LOADER warps generate the tile -- they then set the pace, so the variant keeps four loaders and
// the 80-register allocation; = 1: the loaders stage the gout rows plain and the compute warps form g' = Wp^T gout in step 1
// (4 FMAs per element), three loaders and 96 registers like the other ZA variants.  Same-box timings, cfg-2 shape: loaders x3
// 3.57 ms, loaders x4 3.35-3.60 ms depending on what the register allocator spills at 80, step 1 3.43 ms; prefetching gout a
// tile ahead in the loaders: 3.72-3.94 ms (its 16 registers spill in the compute warps).
#ifndef PDEPHI_GENG_STEP1
#define PDEPHI_GENG_STEP1 1
#endif
constexpr bool GENG_STEP1 = PDEPHI_GENG_STEP1 != 0;
__host__ __device__ constexpr int bwd_ld_warps(bool za, bool geng) { return (za && (!geng || GENG_STEP1)) ? 3 : LD_WARPS; }
__host__ __device__ constexpr int bwd_threads(bool za, bool geng) { return 32 * (1 + bwd_ld_warps(za, geng) + EPI_WARPS); }
constexpr int AUX = 4096;        // per stage: the x rows as a K-major [8 rows][128 points] operand (LIFT_G)
// lift_g (first block of a model whose input needs no gradient: no h tile, no t): a stage is two tiles instead of three and the
// Wc^T image is not needed, so the ring runs THREE stages -- that variant moves 10 GB per launch and the two-stage chain
// (load latency -> step 1 -> MMAs -> release) paced it at 4.0 TB/s.
constexpr int B_MAXST = 3;
// last3 (PDEPHI_LAST3; the LAST block on the fused route, whose g' is generated, not loaded): the MN-major copy of gu exists only
// as the A operand of the Wc^T gu product, and a tcgen05 A operand may live in TMEM in exactly the layout the compute threads
// hold gu in (lane = point, columns = channels).  They store it there (64 columns, single-buffered), the stage shrinks to
// [h][u -> gu K-major] = 64 KB and the ring runs three stages: two tiles' loads in flight instead of one.
#ifndef PDEPHI_LAST3
#define PDEPHI_LAST3 1
#endif
constexpr bool LAST3_ON = PDEPHI_LAST3 != 0 && GENG_STEP1;
// (The same for a MIDDLE block -- g' landing in the K-major tile, gu overwriting it slot by slot, the packed derivative fetched
// by the compute threads themselves one tile ahead in registers, stage = [h][g' -> gu] = 64 KB x 3 -- was built and measured:
// bit-identical, but 3.27 ms against 3.03 ms: the sixteen registers of the two word sets spill (44 / 124 bytes) and the
// threads' own loads cost more than the third stage gains.  Not kept.)
struct BSmem { uint32_t stage[B_MAXST], aux[B_MAXST], wt, red, wp, lift, bars, tmem_slot, total; int nst, uoff; };
__host__ __device__ inline BSmem bwd_smem(bool lift_g = false, bool last3 = false) {
  BSmem s; uint32_t o = 0;
  s.nst = (lift_g || last3) ? 3 : B_STAGES;
  s.uoff = (lift_g || last3) ? TILE : 2 * TILE;                            // offset of the u (-> gu, K layout) tile in a stage
  for (int i = 0; i < B_MAXST; ++i) { s.stage[i] = o; if (i < s.nst) o += (lift_g || last3) ? 2 * TILE : 3 * TILE; }   // g' (-> gu, MN), [h (K),] u (-> gu, K); last3: h, u
  for (int i = 0; i < B_MAXST; ++i) { s.aux[i] = o; if (i < s.nst) o += AUX; }
  s.wt = o; o += lift_g ? 0 : 2 * BOX;
  s.red = o; o += 4 * C * 4;
  s.wp = o; o += C * 16;                                                   // GENG: [64] float4 columns of the projection weight
  s.lift = o; o += C * 16 + C * 4;                                         // GENH: lift weights + bias
  s.bars = o; o += 256;
  s.tmem_slot = o; o += 64;
  s.total = o;
  return s;
}
// GUREADY / TFULL / TEMPTY are per parity of the tile index (two each): the compute warps run step 1 of tile i+1
// before the epilogue of tile i, so two tiles are in flight between them and the MMA warp.
enum { BB_FULL = 0, BB_EMPTY = B_MAXST, BB_GUREADY = 2 * B_MAXST, BB_TFULL = 2 * B_MAXST + 2, BB_TEMPTY = 2 * B_MAXST + 4,
       BB_DWDONE = 2 * B_MAXST + 6, BB_ZAFULL = 2 * B_MAXST + 7, BB_ZAEMPTY = 2 * B_MAXST + 9 };

// GENG: the incoming gradient g' = Wp^T gout of the LAST block (the output projection's input gradient) is generated by the
// loader warps from gout [B, n_out <= 4, S] instead of being read: the 4.3 GB g' is neither written nor read.  (Forming it
// in step 1 from staged gout rows instead was measured slower, 3.62 vs 3.52 ms: the compute warps are this kernel's bound.)
// WT = false: t is not needed (first block of a model whose input needs no gradient: the lift's weight gradient is formed
// from gu and the retained modes instead, functional.py::_lift_backward) and is not stored.
// GENH && !WT ("LIFT_G"): that first block needs no h0 tile either.  dWc = sum gu (x) h0 with h0 = W_in x + b_in is
//   dWc = G W_in^T + dbc (x) b_in,   G = sum_s gu (x) x  [64 x n_in],
// so the weight-gradient GEMM runs against the staged x rows (N = 8 instead of 64) and the reduce kernel finishes dWc;
// G is also what the lift's own weight gradient needs, so the separate pass over gu (pointwise_wgrad, 4.3 GB) goes too.
// ZA: the z stage of the ADJOINT analysis of gu (the next kernel of the backward pass, 4.3 GB written here and read straight
// back there) runs inside this kernel: a tile is one z line (Nz = 128), so  Vg[tile][ch][kz'] = sum_z A[kz'][z] gu[ch][z]  is
// one more GEMM on the staged K-layout gu tile -- D'[128 x 64 ch] = A (TMEM-resident, rows kz' = 2 kz + (re | im), c_kz / N
// folded in) . gu^T -- and gu itself is never stored.  Vg has the layout of the per-axis route's z-analysed intermediate
// ([B*64][Nx][Ny][ldz]), so the y and x stages follow unchanged (transforms_gen.cu: pdephi_block_spectral_bwd).
// PACKDU: the instantiation for PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE = 2 (u_in = packed binary16 derivative)
template <bool GENG, bool WT, bool GENH, bool ZA, bool PACKDU = false>
__global__ void __launch_bounds__(bwd_threads(ZA, GENG), 1)
block_tail_bwd_kernel(const float* __restrict__ g_in, const float* __restrict__ h_in, const float* __restrict__ u_in,
                      const float* __restrict__ Wc, float* __restrict__ gu_out, float* __restrict__ t_out,
                      float* __restrict__ partial, long long tiles_per_batch, long long ntiles, long long S, int contiguous,
                      const float* __restrict__ Wp, int NP, const float* __restrict__ Win, const float* __restrict__ bin, int n_in,
                      const float* __restrict__ za_img, float* __restrict__ vg_out, int ldz, int u_is_du) {
  // u_is_du: u_in holds gelu'(u) as the forward kernels store it under PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  constexpr bool LAST3 = LAST3_ON && GENG && WT && !GENH;            // see bwd_smem (with or without the fused z analysis)
  constexpr bool TMEMA = LAST3;                                      // gu^T as the TMEM A operand of the Wc^T gu product
  const BSmem L = bwd_smem(GENH && !WT, TMEMA);
  auto bar = [&](int i) { return base + L.bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  constexpr bool LIFT_G = GENH && !WT;
  constexpr int NST = (LIFT_G || TMEMA) ? 3 : B_STAGES, UOFF = (LIFT_G || TMEMA) ? TILE : 2 * TILE;     // ring depth, offset of the u tile in a stage
  constexpr int HOFF = TMEMA ? 0 : TILE;                               // offset of the h tile in a stage
  constexpr int NLW = bwd_ld_warps(ZA, GENG), NTHR = bwd_threads(ZA, GENG), BEPI0 = 1 + NLW;      // loader warps / threads / first compute warp
  if (WT) build_w_image(sm + L.wt, Wc, true, threadIdx.x, NTHR, TF32_COMP);      // rows c, K = o : (Wc^T)
  if (GENG) {
    float* wp = reinterpret_cast<float*>(sm + L.wp);
    for (int i = threadIdx.x; i < C * MAXP; i += NTHR) wp[i] = ((i & 3) < NP) ? __ldg(Wp + (i & 3) * C + (i >> 2)) : 0.f;
  }
  if (GENH && !LIFT_G) build_lift_table(sm + L.lift, Win, bin, n_in, threadIdx.x, NTHR);
  // The bias gradient dbc = sum_s gu costs sixteen running sums per thread when it is kept in registers -- the registers that
  // pushed this kernel into spilling.  Where a tensor-core product over the staged gu tile exists anyway it comes out of that
  // product instead, against a row of ONES in the constant operand:
  //   LIFT_G  row 7 of both aux blocks (K-major [8 rows][128 points]; rows < n_in receive the x rows of every tile): column 7
  //           of G = sum gu (x) x accumulates it in TMEM over all tiles of the CTA;
  //   ZA      row 127 of the z-analysis operand (gen_image_kernel kind 5): row 127 of D' holds the tile's sums, lane 31 of the
  //           four quarter-3 compute warps adds them into shared memory (DB_ZA);
  //   else    the running sums in registers (DB_REGS; not on the fused route).
  // (A separate [64 x 8] product against a ones block was measured too: its 16 MMAs delay the release of the stage, which is
  // what paces the last block's generating loaders -- 3.6 -> 4.0 ms.)
  constexpr bool DB_ZA = ZA && !LIFT_G, DB_REGS = !ZA && !LIFT_G;
  if (LIFT_G) {
    float* z = reinterpret_cast<float*>(sm + L.aux[0]);
    for (int i = threadIdx.x; i < NST * AUX / 4; i += NTHR) z[i] = (((i >> 5) & 7) == 7) ? 1.f : 0.f;
  }

  if (GENG && GENG_STEP1) {  // gout rows beyond NP are never staged and meet zero weights: keep them finite
    float* z = reinterpret_cast<float*>(sm + L.aux[0]);
    for (int i = threadIdx.x; i < NST * AUX / 4; i += NTHR) z[i] = 0.f;
  }
  if (threadIdx.x < 4 * C) reinterpret_cast<float*>(sm + L.red)[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(bar(BB_FULL + i), 32 * NLW); mbar_init(bar(BB_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(BB_GUREADY + i), EPI_WARPS);
      mbar_init(bar(BB_TFULL + i), 1);
      mbar_init(bar(BB_TEMPTY + i), EPI_WARPS);
      mbar_init(bar(BB_ZAFULL + i), 1);
      mbar_init(bar(BB_ZAEMPTY + i), EPI_WARPS);
    }
    mbar_init(bar(BB_DWDONE), 1);
    fence_barrier_init();
  }
  constexpr uint32_t TMEM_COLS = ZA ? 512 : 256;
  if (warp == 0) tmem_alloc(base + L.tmem_slot, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_slot);
  const uint32_t tmem_t[2] = {tmem, tmem + C}, tmem_dw = tmem + 2 * C;
  const uint32_t tmem_za[2] = {tmem + 3 * C, tmem + 4 * C}, tmem_a = tmem + 6 * C;   // ZA: accumulators, A image [128 x 128]
  const uint32_t tmem_gu = tmem + (ZA ? 5 : 3) * C;                                  // TMEMA: gu^T [128 points x 64 channels]
  if (ZA) {
    if (warp >= BEPI0 && warp < BEPI0 + 4) {       // four consecutive warps cover the four TMEM lane quarters
      const int q = warp & 3;
      const float* row = za_img + (size_t)(32 * q + lane) * TS;
      for (int c0 = 0; c0 < TS; c0 += 32) {
        float r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(row + c0) + j);
          r[4 * j] = t4.x; r[4 * j + 1] = t4.y; r[4 * j + 2] = t4.z; r[4 * j + 3] = t4.w;
        }
        tmem_st32(tmem_a + ((uint32_t)(q * 32) << 16) + c0, r);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // every CTA walks a CONTIGUOUS range of tiles (PDEPHI_TILE_ORDER=0: round-robin): consecutive tiles extend the same 64
  // channel rows, so the 512-byte pieces a tile writes to each row meet their neighbours in L2 within one tile time
  const long long tq = ntiles / gridDim.x, tr = ntiles % gridDim.x;
  const long long my_tiles = contiguous ? tq + (blockIdx.x < tr ? 1 : 0)
                                        : (blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  const long long tile0 = contiguous ? blockIdx.x * tq + (blockIdx.x < tr ? blockIdx.x : tr) : blockIdx.x;
  const long long tstep = contiguous ? 1 : gridDim.x;

  if (warp >= 1 && warp < BEPI0) {
    // ===== loaders: g' (MN layout), h (K layout), u (K layout); a landed tile is published before the loader blocks on
    //       the stage of the next one (see the forward kernel) =====
    const int lw = warp - 1;
    auto issue = [&](long long it) {
      const long long t = tile0 + it * tstep;
      const int s = (int)(it % NST);
      mbar_wait(bar(BB_EMPTY + s), (uint32_t)(((it / NST) & 1) ^ 1));
      const long long b = t / tiles_per_batch, s0 = (t - b * tiles_per_batch) * TS;
      float4 gv[MAXP];
      if (GENG && GENG_STEP1) {   // g_in = gout [B, NP <= 4, S]: its rows go plain into aux[s]; the compute warps form g' = Wp^T gout
        if (lw == 0) stage_small_rows(base + L.aux[s], g_in, NP, b, s0, S, lane);
      } else if (GENG) {          // ... or the loader warps do: this lane's four points of every output channel, issued ahead of the copies
#pragma unroll
        for (int n = 0; n < MAXP; ++n)
          gv[n] = (n < NP) ? __ldg(reinterpret_cast<const float4*>(g_in + ((size_t)b * NP + n) * S + s0) + lane)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        stage_rows<false, NLW>(base + L.stage[s], g_in, b, s0, S, lw, lane);
      }
      if (LIFT_G) {      // h_in = model input x: its n_in rows as a K-major operand [8 rows][128 points], 128-byte swizzle
        if (lw == 0) {
          const int j = lane >> 3, cc = lane & 7;
          for (int r = 0; r < n_in; ++r)
            cp_async16(base + L.aux[s] + (uint32_t)(j * 1024 + r * 128 + (((cc ^ (r & 7)) & 7) << 4)),
                       h_in + ((size_t)b * n_in + r) * S + s0 + lane * 4);
        }
      } else if (GENH) {                                                           // h_in = model input x
        float4 xv[MAXP_IN];
        load_x(xv, h_in, n_in, b, s0, S, lane);
        generate_rows<true, NLW>(sm + L.stage[s] + TILE, xv, sm + L.lift, lw, lane);
      } else {
        stage_rows<true, NLW>(base + L.stage[s] + HOFF, h_in, b, s0, S, lw, lane);
      }
      if (PACKDU) stage_pair_rows<NLW>(base + L.stage[s] + UOFF, u_in, b, s0, S, lw, lane);
      else stage_rows<true, NLW>(base + L.stage[s] + UOFF, u_in, b, s0, S, lw, lane);
      if (GENG && !GENG_STEP1) {
        uint8_t* dst = sm + L.stage[s];
#pragma unroll
        for (int i = 0; i < (C + NLW - 1) / NLW; ++i) {
          const int r = lw + NLW * i;
          if (C % NLW != 0 && r >= C) break;
          const float4 w = reinterpret_cast<const float4*>(sm + L.wp)[r];
          float4 o;
          o.x = fmaf(w.w, gv[3].x, fmaf(w.z, gv[2].x, fmaf(w.y, gv[1].x, w.x * gv[0].x)));
          o.y = fmaf(w.w, gv[3].y, fmaf(w.z, gv[2].y, fmaf(w.y, gv[1].y, w.x * gv[0].y)));
          o.z = fmaf(w.w, gv[3].z, fmaf(w.z, gv[2].z, fmaf(w.y, gv[1].z, w.x * gv[0].z)));
          o.w = fmaf(w.w, gv[3].w, fmaf(w.z, gv[2].w, fmaf(w.y, gv[1].w, w.x * gv[0].w)));
          *reinterpret_cast<float4*>(dst + chunk_off_mn(r, lane)) = o;
        }
      }
    };
    for (int pre = 0; pre < NST - 1; ++pre) {
      if (pre < my_tiles) issue(pre);
      cp_async_commit();
    }
    for (long long it = 0; it < my_tiles; ++it) {
      cp_async_wait<NST - 2>();
      fence_proxy_async();
      mbar_arrive(bar(BB_FULL + (int)(it % NST)));
      if (it + NST - 1 < my_tiles) issue(it + NST - 1);
      cp_async_commit();
    }
  } else if (warp == 0) {
    if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows one lane is active and issues UTCHMMA back to back
      constexpr uint32_t id_t = idesc_tf32(TS, C, 1, 0);     // T[128 x 64] = GU^T (MN-major A) . WcT (K-major B)
      constexpr uint32_t id_w = idesc_tf32(C, C, 0, 0);      // dW[64 x 64] += GU (K-major A) . H (K-major B)
      for (long long it = 0; it < my_tiles; ++it) {
        const int s = (int)(it % NST);
        const int d = (int)(it & 1);
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        mbar_wait(bar(BB_GUREADY + d), ph);                  // gu written in place over g' (implies the tile landed)
        mbar_wait(bar(BB_TEMPTY + d), ph ^ 1);
        tc_fence_after();
        const uint32_t gu = base + L.stage[s], hh = gu + HOFF, guk = gu + UOFF;
        if (WT && TMEMA) {       // A = gu^T from TMEM (K-major by construction: lane = point, column = channel o)
          constexpr uint32_t id_ta = idesc_tf32(TS, C, 0, 0);
#pragma unroll
          for (int k = 0; k < C / 8; ++k)
            mma_ts(tmem_t[d], tmem_gu + 8 * k, umma_desc(base + L.wt + (k >> 2) * BOX + (k & 3) * 32, 1024), id_ta, k ? 1u : 0u);
        } else if (WT) {
#pragma unroll
          for (int k = 0; k < C / 8; ++k)
            mma_ss(tmem_t[d], umma_desc_mn(gu + k * 1024, BOX), umma_desc(base + L.wt + (k >> 2) * BOX + (k & 3) * 32, 1024), id_t,
                   k ? 1u : 0u);
        }
        mma_commit(bar(BB_TFULL + d));
        if (LIFT_G) {      // G[64 x 8] += GU (K-major A) . X (K-major B, rows >= n_in are zero)
          constexpr uint32_t id_g = idesc_tf32(C, 8, 0, 0);
          const uint32_t xx = base + L.aux[s];
#pragma unroll
          for (int k = 0; k < TS / 8; ++k)
            mma_ss(tmem_dw, umma_desc(guk + (k >> 2) * BOX + (k & 3) * 32, 1024), umma_desc(xx + (k >> 2) * 1024 + (k & 3) * 32, 1024),
                   id_g, (it | k) ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < TS / 8; ++k)
            mma_ss(tmem_dw, umma_desc(guk + (k >> 2) * BOX + (k & 3) * 32, 1024), umma_desc(hh + (k >> 2) * BOX + (k & 3) * 32, 1024),
                   id_w, (it | k) ? 1u : 0u);
        }
        if (ZA) {          // D'[128 (kz') x 64 ch] = A[128 x 128 z] (TMEM) . gu[64 ch x 128 z]^T (K-major B, the dWc GEMM's A tile)
          constexpr uint32_t id_z = idesc_tf32(TS, C, 0, 0);
          mbar_wait(bar(BB_ZAEMPTY + d), ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < TS / 8; ++k)
            mma_ts(tmem_za[d], tmem_a + 8 * k, umma_desc(guk + (k >> 2) * BOX + (k & 3) * 32, 1024), id_z, k ? 1u : 0u);
          mma_commit(bar(BB_ZAFULL + d));
        }
        mma_commit(bar(BB_EMPTY + s));
      }
      mma_commit(bar(BB_DWDONE));
    }
  } else {
    const int q = warp & 3, part = (warp - BEPI0) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float db[DB_REGS ? CPW : 1];
#pragma unroll
    for (int j = 0; j < (DB_REGS ? CPW : 1); ++j) db[j] = 0.f;
    // static shared-memory offsets of this thread's elements: row o = part*CPW + j, point 32q + lane.  The swizzle term
    // depends on o & 3 (32-byte-atom layout) or o & 7 (16-byte-chunk layout), i.e. on j only (CPW is a multiple of 8).
    uint32_t xc[4], xk[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) xc[r] = (uint32_t)(q * BOX + part * CPW * 128 + ((((lane >> 3) ^ r) & 3) << 5) + (lane & 7) * 4);
#pragma unroll
    for (int r = 0; r < 8; ++r) xk[r] = (uint32_t)(q * BOX + part * CPW * 128 + ((((lane >> 2) ^ r) & 7) << 4) + (lane & 3) * 4);
    // ---- step 1: gu = g' * gelu'(u) (this warp: channels [16*part, +16), points [32q, +32)).  gu replaces g' in place
    //      (MN layout, operand of the Wc^T gu GEMM) and u in place (K layout, operand of the dWc GEMM), goes to HBM, and
    //      stays in registers (r) for the epilogue ----
    auto step1 = [&](long long it, float* r) {
      const long long t = tile0 + it * tstep;
      const int s = (int)(it % NST), d = (int)(it & 1);
      const long long b = t / tiles_per_batch, s0 = (t - b * tiles_per_batch) * TS;
      uint8_t* tg = sm + L.stage[s];
      uint8_t* tgk = tg + UOFF;
      mbar_wait(bar(BB_FULL + s), (uint32_t)((it / NST) & 1));
      float* gp = gu_out + ((size_t)b * C + part * CPW) * S + s0 + 32 * q + lane;
      float go[MAXP];
      if (GENG && GENG_STEP1) {
        const float* gs = reinterpret_cast<const float*>(sm + L.aux[s]) + 32 * q + lane;
#pragma unroll
        for (int n = 0; n < MAXP; ++n) go[n] = gs[n * TS];
      }
      auto element = [&](int j, float dact) {     // dact = gelu'(u) of this element
        const uint32_t off = xc[j & 3] + j * 128, offk = xk[j & 7] + j * 128;
        float gin;
        if (GENG && GENG_STEP1) {
          const float4 w = reinterpret_cast<const float4*>(sm + L.wp)[part * CPW + j];
          gin = fmaf(w.w, go[3], fmaf(w.z, go[2], fmaf(w.y, go[1], w.x * go[0])));
        } else {
          gin = *reinterpret_cast<const float*>(tg + off);
        }
        r[j] = gin * dact;
        if (WT && !TMEMA) *reinterpret_cast<float*>(tg + off) = r[j];
        *reinterpret_cast<float*>(tgk + offk) = to_tf32(r[j]);
        if (!ZA) { *gp = r[j]; gp += S; }
        if (DB_REGS) db[j] += r[j];
      };
      if (PACKDU) {        // ... as binary16 pairs in the even rows' slots (pack_du): read the word, then overwrite both rows' slots
#pragma unroll
        for (int j = 0; j < CPW; j += 2) {
          const float2 dd = unpack_du(*reinterpret_cast<const uint32_t*>(tgk + xk[j & 7] + j * 128));
          element(j, dd.x);
          element(j + 1, dd.y);
        }
      } else if (u_is_du) {       // the forward stored the derivative: one multiply per element instead of the erf / density evaluation
#pragma unroll
        for (int j = 0; j < CPW; ++j) element(j, *reinterpret_cast<const float*>(tgk + xk[j & 7] + j * 128));
      } else {
#pragma unroll
        for (int j = 0; j < CPW; ++j) element(j, gelu_grad_f(*reinterpret_cast<const float*>(tgk + xk[j & 7] + j * 128)));
      }
      if (TMEMA) {         // gu^T -> TMEM, once the previous tile's Wc^T gu product has read the (single) buffer -- it was issued a
                           // whole step 1 ago, so this wait does not block in steady state
        if (it > 0) mbar_wait(bar(BB_TFULL + (int)((it - 1) & 1)), (uint32_t)(((it - 1) >> 1) & 1));
        tc_fence_after();
        tmem_st16(tmem_gu + lane_base + part * CPW, r);
        tmem_st_wait();
        tc_fence_before();
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BB_GUREADY + d));
    };
    // ---- epilogue: t[c][s] = (Wc^T gu)[c][s] + gu[c][s] ----
    auto epilogue = [&](long long it, const float* r) {
      const long long t = tile0 + it * tstep;
      const int d = (int)(it & 1);
      const long long b = t / tiles_per_batch, s0 = (t - b * tiles_per_batch) * TS;
      mbar_wait(bar(BB_TFULL + d), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[CPW];
      if (WT) {
        tmem_ld16(tmem_t[d] + lane_base + part * CPW, v);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BB_TEMPTY + d));
      if (WT) {
        float* tp = t_out + ((size_t)b * C + part * CPW) * S + s0 + 32 * q + lane;
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
          __stcs(tp, v[j] + r[j]);
          tp += S;
        }
      }
      if (ZA) {            // rows kz' = 32 q + lane of D' exist in quarters 0 and 1 only (ldz <= 64)
        mbar_wait(bar(BB_ZAFULL + d), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        if (q < 2 || (DB_ZA && q == 3)) {
          tmem_ld16(tmem_za[d] + lane_base + part * CPW, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BB_ZAEMPTY + d));
        if (DB_ZA && q == 3 && lane == 31) {     // row 127 of D' = sum over the tile's points of gu (the operand's ones row)
          float* acc = reinterpret_cast<float*>(sm + L.red) + part * CPW;
#pragma unroll
          for (int j = 0; j < CPW; ++j) acc[j] += v[j];
        }
        if (q < 2 && 32 * q + lane < ldz) {       // Vg [B*64 volumes][tiles of a volume = (x, y)][ldz]: the W1 layout of the per-axis route
          const long long b = t / tiles_per_batch, tib = t - b * tiles_per_batch;
          float* vp = vg_out + (((size_t)b * C + part * CPW) * tiles_per_batch + tib) * ldz + 32 * q + lane;
          const size_t cs = (size_t)tiles_per_batch * ldz;
#pragma unroll
          for (int j = 0; j < CPW; ++j) vp[(size_t)j * cs] = v[j];
        }
      }
    };
    // Step 1 of tile i+1 runs BEFORE the epilogue of tile i: the Wc^T gu GEMM of tile i (and its commit / barrier round
    // trip, ~1000 clocks) is then hidden behind useful work instead of being waited for by all sixteen warps.
    // (order s0 s1 e0 s2 e1 s3 e2 ...; written so that each of the two register sets has ONE inlined copy of step 1 and of the
    // epilogue -- the kernel is ~100 KB of code with a third copy of step 1)
    float ra[CPW], rb[CPW];
    for (long long it = 0; it <= my_tiles; it += 2) {
      if (it < my_tiles) step1(it, ra);
      if (it >= 1) epilogue(it - 1, rb);
      if (it + 1 < my_tiles) step1(it + 1, rb);
      if (it < my_tiles) epilogue(it, ra);
    }
    // ---- per-CTA partials: the bias gradient (see DB_ZA / DB_REGS / LIFT_G above) and dW or G (TMEM) ----
    float* red = reinterpret_cast<float*>(sm + L.red);      // DB_REGS: [4 quarters][64];  DB_ZA: [64]
    if (DB_REGS) {
#pragma unroll
      for (int j = 0; j < CPW; ++j) {
        float x = db[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) red[q * C + part * CPW + j] = x;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    float* pout = partial + (size_t)blockIdx.x * (C * C + C);
    if (!LIFT_G && warp == BEPI0) {
      for (int o = lane; o < C; o += 32)
        pout[C * C + o] = DB_REGS ? (red[o] + red[C + o]) + (red[2 * C + o] + red[3 * C + o]) : red[o];
    }
    if (my_tiles > 0) {
      mbar_wait(bar(BB_DWDONE), 0);
      tc_fence_after();
      if (part == 0) {                 // M = 64 accumulator: rows 16q..16q+15 live in lanes 0..15 of quarter q
        if (LIFT_G) {                  // G: 8 columns, column 7 = sum gu (the ones row)
          float w0[8];
          tmem_ld8(tmem_dw + lane_base, w0);
          tmem_ld_wait();
          if (lane < 16) {
            float* row = pout + (size_t)(16 * q + lane) * C;
#pragma unroll
            for (int c = 0; c < 7; ++c) row[c] = w0[c];
            row[7] = 0.f;
            pout[C * C + 16 * q + lane] = w0[7];
          }
        } else {
          float w0[32], w1[32];
          tmem_ld32(tmem_dw + lane_base, w0);
          tmem_ld32(tmem_dw + lane_base + 32, w1);
          tmem_ld_wait();
          if (lane < 16) {
            float* row = pout + (size_t)(16 * q + lane) * C;
#pragma unroll
            for (int c = 0; c < 32; ++c) { row[c] = w0[c]; row[32 + c] = w1[c]; }
          }
        }
      }
    } else if (part == 0 && lane < 16) {
      float* row = pout + (size_t)(16 * q + lane) * C;
      for (int c = 0; c < C; ++c) row[c] = 0.f;
      if (LIFT_G) pout[C * C + 16 * q + lane] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

__global__ void block_tail_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ db,
                                         int nparts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * C + C) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)partial[(size_t)p * (C * C + C) + idx];
  if (idx < C * C) dW[idx] = (float)(s * 1.00048828125);   // h is truncated to TF32 by the tensor core in the dWc GEMM
  else if (db) db[idx - C * C] = (float)s;
}

// LIFT_G: G[c][n] = sum of the per-CTA partials (first 8 columns of each 64-wide row), dbc likewise, then
// dWc[c][k] = sum_n G[c][n] W_in[k][n] + dbc[c] b_in[k]   (one block of 1024 threads)
__global__ void block_tail_lift_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ Win,
                                              const float* __restrict__ bin, int n_in, float* __restrict__ dW, float* __restrict__ db,
                                              float* __restrict__ G, int nparts) {
  __shared__ float sG[C * 8];
  __shared__ float sdb[C];
  const int tid = threadIdx.x;
  if (tid < C * 8 + C) {
    const int idx = tid < C * 8 ? (tid >> 3) * C + (tid & 7) : C * C + (tid - C * 8);
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)partial[(size_t)p * (C * C + C) + idx];
    if (tid < C * 8) sG[tid] = (float)(s * 1.00048828125);       // x is truncated to TF32 by the tensor core in the G GEMM
    else sdb[tid - C * 8] = (float)s;
  }
  __syncthreads();
  if (tid < C) { if (db) db[tid] = sdb[tid]; }
  if (G && tid < C * n_in) G[tid] = sG[(tid / n_in) * 8 + tid % n_in];
  for (int i = tid; i < C * C; i += blockDim.x) {
    const int c = i >> 6, k = i & 63;
    float acc = bin ? sdb[c] * __ldg(bin + k) : 0.f;
    for (int n = 0; n < n_in; ++n) acc = fmaf(sG[c * 8 + n], __ldg(Win + k * n_in + n), acc);
    dW[i] = acc;
  }
}

}  // namespace blk

static int check_shape(int B, int Cc, long long S, const char* who) {
  PDEPHI_REQUIRE(B > 0 && S > 0, "%s: non-positive size", who);
  if (Cc != blk::C || S % blk::TS != 0 || (long long)B * Cc > 0x7fffffffLL || S > 0x7fffffffLL)
    return fail(PDEPHI_ERR_UNSUPPORTED, "%s: needs 64 channels and a spatial size divisible by 128 (got C=%d, S=%lld)", who, Cc, S);
  return PDEPHI_OK;
}

static int tile_order() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PDEPHI_TILE_ORDER"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

static int num_sms() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

}  // namespace pdephi

using namespace pdephi;

namespace pdephi {
struct ProjArgs { const float* Wp; const float* bp; float* out; int n; };   // output projection fused into the last block
struct LiftArgs { const float* Win; const float* bin; int n; };              // input lift fused into the first block (h = x then)

static int check_lift(const LiftArgs& lf, const ProjArgs* pr, const char* who) {
  if (!lf.Win) return PDEPHI_OK;
  if (lf.n < 1 || lf.n > blk::MAXP_IN) return fail(PDEPHI_ERR_UNSUPPORTED, "%s: the fused lift takes 1..%d input channels, got %d", who, blk::MAXP_IN, lf.n);
  if (pr && pr->Wp) return fail(PDEPHI_ERR_UNSUPPORTED, "%s: lift and projection in the same block are not combined", who);
  return PDEPHI_OK;
}

static int check_proj(const ProjArgs& pr, const char* who) {
  if (!pr.Wp) return PDEPHI_OK;
  PDEPHI_REQUIRE(pr.out != nullptr, "%s: projection output is null", who);
  if (pr.n < 1 || pr.n > blk::MAXP) return fail(PDEPHI_ERR_UNSUPPORTED, "%s: the fused projection takes 1..%d outputs, got %d", who, blk::MAXP, pr.n);
  return PDEPHI_OK;
}

#define PDEPHI_FWD_LAUNCH_Z(FF, PP, HH, ZZ, SPEC, BZ, LDZ, ZAIMG, VH)                                                       \
  {                                                                                                                        \
    const int du_mode = option_get(OPT_BLOCK_SAVES_DERIVATIVE);                                                            \
    auto k = du_mode == 2 ? blk::block_tail_fwd_kernel<FF, PP, HH, ZZ, true> : blk::block_tail_fwd_kernel<FF, PP, HH, ZZ, false>; \
    PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                         \
    k<<<grid, blk::THREADS, smem, st>>>(h, SPEC, Wc, bc, u, hout, tpb, ntiles, S, BZ, LDZ, tile_order(), pr.Wp, pr.bp, pr.out, \
                                        pr.n, lf.Win, lf.bin, lf.n, du_mode, ZAIMG, VH);                                   \
  }
#define PDEPHI_FWD_LAUNCH(FF, PP, HH, SPEC, BZ, LDZ) PDEPHI_FWD_LAUNCH_Z(FF, PP, HH, false, SPEC, BZ, LDZ, nullptr, nullptr)

static int block_tail_fwd_launch(const float* h, const float* spec, const float* Wc, const float* bc, float* u, float* hout,
                                 int B, long long S, const ProjArgs& pr, const LiftArgs& lf, cudaStream_t st) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(h) || !al16(spec)) return fail(PDEPHI_ERR_INVALID, "block_tail_fwd: h and spec must be 16-byte aligned");
  int rc = check_proj(pr, "block_tail_fwd");
  if (rc) return rc;
  rc = check_lift(lf, &pr, "block_tail_fwd");
  if (rc) return rc;
  const long long tpb = S / blk::TS, ntiles = tpb * B;
  const bool proj = pr.Wp != nullptr, genh = lf.Win != nullptr;
  const size_t smem = blk::fwd_smem(false, proj, genh).total + 1024;
  const int sms = num_sms();
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  ProfScope ps(K_BLOCK_TAIL_FWD, st);
  if (genh) PDEPHI_FWD_LAUNCH(false, false, true, spec, nullptr, 0)
  else if (proj) PDEPHI_FWD_LAUNCH(false, true, false, spec, nullptr, 0)
  else PDEPHI_FWD_LAUNCH(false, false, false, spec, nullptr, 0)
  PDEPHI_LAUNCH_CHECK("block_tail_fwd_kernel");
  return PDEPHI_OK;
}

// Block tail with the z stage of the synthesis fused in (transforms_gen.cu: pdephi_block_spectral_fwd).
// V [B*S/128 tiles][64][ldz] float, bzimg = the plan's z-synthesis operand image for Nz = 128.
// za_img / vh: additionally run the z stage of the next layer's FORWARD analysis on h' (ZF variants; not with the projection)
int block_tail_fwd_fused_launch(const float* h, const float* V, int ldz, const float* bzimg, const float* Wc, const float* bc,
                                float* u, float* hout, int B, long long S, const float* Wp, const float* bp, float* proj_out,
                                int n_out, const float* Win, const float* bin, int n_in, cudaStream_t st,
                                const float* za_img, float* vh) {
  int rc = check_shape(B, blk::C, S, "block_spectral_fwd");
  if (rc) return rc;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(h) || !al16(V)) return fail(PDEPHI_ERR_INVALID, "block_spectral_fwd: h and the workspace must be 16-byte aligned");
  if (ldz % 4 != 0 || ldz > 64) return fail(PDEPHI_ERR_UNSUPPORTED, "block_spectral_fwd: ldz %d", ldz);
  const ProjArgs pr{Wp, bp, proj_out, n_out};
  const LiftArgs lf{Win, bin, n_in};
  rc = check_proj(pr, "block_spectral_fwd");
  if (rc) return rc;
  rc = check_lift(lf, &pr, "block_spectral_fwd");
  if (rc) return rc;
  const long long tpb = S / blk::TS, ntiles = tpb * B;
  const bool proj = Wp != nullptr, genh = Win != nullptr, zf = vh != nullptr;
  if (zf && (proj || !za_img || (reinterpret_cast<uintptr_t>(za_img) & 15)))
    return fail(PDEPHI_ERR_INVALID, "block_spectral_fwd: the fused z analysis of the output needs an operand image and no projection");
  const size_t smem = blk::fwd_smem(true, proj, genh, zf).total + 1024;
  const int sms = num_sms();
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  ProfScope ps(K_BLOCK_SPECTRAL_FWD, st);
  if (zf && genh) PDEPHI_FWD_LAUNCH_Z(true, false, true, true, V, bzimg, ldz, za_img, vh)
  else if (zf) PDEPHI_FWD_LAUNCH_Z(true, false, false, true, V, bzimg, ldz, za_img, vh)
  else if (genh) PDEPHI_FWD_LAUNCH(true, false, true, V, bzimg, ldz)
  else if (proj) PDEPHI_FWD_LAUNCH(true, true, false, V, bzimg, ldz)
  else PDEPHI_FWD_LAUNCH(true, false, false, V, bzimg, ldz)
  PDEPHI_LAUNCH_CHECK("block_tail_fwd_kernel<fused>");
  return PDEPHI_OK;
}

// za_img / vg / ldz: the ZA variant (z stage of the adjoint analysis of gu fused in; gu is not stored then and may be NULL)
int block_tail_bwd_launch(const float* g, const float* u, const float* h, const float* Wc, float* gu, float* t, float* dWc,
                          float* dbc, int B, int Cc, long long S, void* workspace, size_t workspace_bytes,
                          const float* Wp, int n_out, const LiftArgs& lf, cudaStream_t st, float* G = nullptr,
                          const float* za_img = nullptr, float* vg = nullptr, int ldz = 0) {
  int rc = check_shape(B, Cc, S, "block_tail_bwd");
  if (rc) return rc;
  const size_t need = pdephi_block_tail_bwd_workspace_bytes(Cc);
  if (!workspace || workspace_bytes < need) return fail(PDEPHI_ERR_WORKSPACE, "block_tail_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(g) || !al16(h) || !al16(u)) return fail(PDEPHI_ERR_INVALID, "block_tail_bwd: g, h and u must be 16-byte aligned");
  if (Wp && (n_out < 1 || n_out > blk::MAXP))
    return fail(PDEPHI_ERR_UNSUPPORTED, "block_tail_bwd: the fused projection takes 1..%d outputs, got %d", blk::MAXP, n_out);
  if (vg && (!za_img || ldz < 4 || ldz > 64 || (reinterpret_cast<uintptr_t>(za_img) & 15)))
    return fail(PDEPHI_ERR_INVALID, "block_tail_bwd: fused z analysis needs an operand image and 4 <= ldz <= 64");
  if (!vg && !gu) return fail(PDEPHI_ERR_INVALID, "block_tail_bwd: gu is null");
  {
    const ProjArgs pr{Wp, nullptr, nullptr, n_out};
    rc = check_lift(lf, &pr, "block_tail_bwd");
    if (rc) return rc;
  }
  const long long tpb = S / blk::TS, ntiles = tpb * B;
  const int du_mode = option_get(OPT_BLOCK_SAVES_DERIVATIVE);
  const bool last3 = blk::LAST3_ON && Wp != nullptr && t != nullptr && lf.Win == nullptr;
  const size_t smem = blk::bwd_smem(lf.Win != nullptr && t == nullptr, last3).total + 1024;
  int sms = num_sms();
  if (sms > 1024) sms = 1024;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  const int kid = vg ? K_BLOCK_SPECTRAL_BWD : K_BLOCK_TAIL_BWD;
  {
    ProfScope ps(kid, st);
#define PDEPHI_BWD_LAUNCH_Z(GG, WW, HH, ZZ)                                                                                \
    {                                                                                                                      \
      auto k = du_mode == 2 ? blk::block_tail_bwd_kernel<GG, WW, HH, ZZ, true> : blk::block_tail_bwd_kernel<GG, WW, HH, ZZ, false>; \
      PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                       \
      k<<<grid, blk::bwd_threads(ZZ, GG), smem, st>>>(g, h, u, Wc, gu, t, (float*)workspace, tpb, ntiles, S, tile_order(), Wp, n_out,  \
                                          lf.Win, lf.bin, lf.n, za_img, vg, ldz, du_mode);                                 \
    }
#define PDEPHI_BWD_LAUNCH(GG, WW, HH) { if (vg) PDEPHI_BWD_LAUNCH_Z(GG, WW, HH, true) else PDEPHI_BWD_LAUNCH_Z(GG, WW, HH, false) }
    if (lf.Win) { if (t) PDEPHI_BWD_LAUNCH(false, true, true) else PDEPHI_BWD_LAUNCH(false, false, true) }
    else if (Wp) { if (t) PDEPHI_BWD_LAUNCH(true, true, false) else PDEPHI_BWD_LAUNCH(true, false, false) }
    else { if (t) PDEPHI_BWD_LAUNCH(false, true, false) else PDEPHI_BWD_LAUNCH(false, false, false) }
#undef PDEPHI_BWD_LAUNCH
#undef PDEPHI_BWD_LAUNCH_Z
  }
  PDEPHI_LAUNCH_CHECK("block_tail_bwd_kernel");
  {
    ProfScope ps(kid, st);
    if (lf.Win && !t)      // LIFT_G: the partials hold G = sum gu (x) x; dWc = G W_in^T + dbc (x) b_in
      blk::block_tail_lift_reduce_kernel<<<1, 1024, 0, st>>>((const float*)workspace, lf.Win, lf.bin, lf.n, dWc, dbc, G, grid);
    else
      blk::block_tail_reduce_kernel<<<(blk::C * blk::C + blk::C + 127) / 128, 128, 0, st>>>((const float*)workspace, dWc, dbc, grid);
  }
  PDEPHI_LAUNCH_CHECK("block_tail_reduce_kernel");
  return PDEPHI_OK;
}

// Block-tail backward with the z stage of the adjoint analysis of gu fused in (transforms_gen.cu: pdephi_block_spectral_bwd).
// vg [B*64][S/128][ldz] float, za_img = the plan's plain [128][128] adjoint z-analysis operand for Nz = 128.
int block_tail_bwd_za_launch(const float* g, const float* u, const float* h, const float* Wc, float* t, float* dWc, float* dbc,
                             int B, long long S, void* workspace, size_t workspace_bytes, const float* Wp, int n_out,
                             const float* Win, const float* bin, int n_in, float* G, const float* za_img, float* vg, int ldz,
                             cudaStream_t st) {
  if (!vg) return fail(PDEPHI_ERR_INVALID, "block_spectral_bwd: null workspace");
  return block_tail_bwd_launch(g, u, h, Wc, nullptr, t, dWc, dbc, B, blk::C, S, workspace, workspace_bytes, Wp, n_out,
                               LiftArgs{Win, bin, n_in}, st, G, za_img, vg, ldz);
}
}  // namespace pdephi

extern "C" size_t pdephi_block_tail_bwd_workspace_bytes(int Cc) {
  if (Cc != blk::C) return 0;
  return align_up((size_t)1024 * (blk::C * blk::C + blk::C) * sizeof(float), 256);   // up to 1024 CTAs of partials
}

extern "C" int pdephi_block_tail_fwd(const float* h, const float* spec, const float* Wc, const float* bc, float* u,
                                     float* hout, int B, int Cc, long long S, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(h && spec && Wc && u && hout, "block_tail_fwd: null pointer");
  int rc = check_shape(B, Cc, S, "block_tail_fwd");
  if (rc) return rc;
  return block_tail_fwd_launch(h, spec, Wc, bc, u, hout, B, S, ProjArgs{nullptr, nullptr, nullptr, 0}, LiftArgs{nullptr, nullptr, 0},
                               (cudaStream_t)stream);
}

extern "C" int pdephi_block_tail_lift_fwd(const float* x, const float* Win, const float* bin, int n_in, const float* spec,
                                          const float* Wc, const float* bc, float* u, float* hout, int B, int Cc, long long S,
                                          pdephi_stream_t stream) {
  PDEPHI_REQUIRE(x && Win && spec && Wc && u && hout, "block_tail_lift_fwd: null pointer");
  int rc = check_shape(B, Cc, S, "block_tail_lift_fwd");
  if (rc) return rc;
  return block_tail_fwd_launch(x, spec, Wc, bc, u, hout, B, S, ProjArgs{nullptr, nullptr, nullptr, 0}, LiftArgs{Win, bin, n_in},
                               (cudaStream_t)stream);
}

extern "C" int pdephi_block_tail_proj_fwd(const float* h, const float* spec, const float* Wc, const float* bc, float* u,
                                          float* hout, const float* Wp, const float* bp, int n_out, float* out, int B, int Cc,
                                          long long S, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(h && spec && Wc && u && hout && Wp && out, "block_tail_proj_fwd: null pointer");
  int rc = check_shape(B, Cc, S, "block_tail_proj_fwd");
  if (rc) return rc;
  return block_tail_fwd_launch(h, spec, Wc, bc, u, hout, B, S, ProjArgs{Wp, bp, out, n_out}, LiftArgs{nullptr, nullptr, 0},
                               (cudaStream_t)stream);
}

extern "C" int pdephi_block_tail_bwd(const float* g, const float* u, const float* h, const float* Wc, float* gu, float* t,
                                     float* dWc, float* dbc, int B, int Cc, long long S, void* workspace,
                                     size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(g && u && h && Wc && gu && dWc, "block_tail_bwd: null pointer");     /* t may be NULL: not stored */
  return block_tail_bwd_launch(g, u, h, Wc, gu, t, dWc, dbc, B, Cc, S, workspace, workspace_bytes, nullptr, 0,
                               LiftArgs{nullptr, nullptr, 0}, (cudaStream_t)stream);
}

extern "C" int pdephi_block_tail_lift_bwd(const float* g, const float* u, const float* x, const float* Win, const float* bin,
                                          int n_in, const float* Wc, float* gu, float* t, float* dWc, float* dbc, float* G, int B,
                                          int Cc, long long S, void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(g && u && x && Win && Wc && gu && dWc, "block_tail_lift_bwd: null pointer");     /* t, G may be NULL */
  return block_tail_bwd_launch(g, u, x, Wc, gu, t, dWc, dbc, B, Cc, S, workspace, workspace_bytes, nullptr, 0,
                               LiftArgs{Win, bin, n_in}, (cudaStream_t)stream, G);
}

extern "C" int pdephi_block_tail_proj_bwd(const float* gout, const float* Wp, int n_out, const float* u, const float* h,
                                          const float* Wc, float* gu, float* t, float* dWc, float* dbc, int B, int Cc, long long S,
                                          void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(gout && Wp && u && h && Wc && gu && dWc, "block_tail_proj_bwd: null pointer");   /* t may be NULL */
  return block_tail_bwd_launch(gout, u, h, Wc, gu, t, dWc, dbc, B, Cc, S, workspace, workspace_bytes, Wp, n_out,
                               LiftArgs{nullptr, nullptr, 0}, (cudaStream_t)stream);
}
