// Channels-first pointwise linear maps of the operator models:
//   lift    input_proj  : [B, 3, S] -> [B, W, S]     (rational_fourier.py:258-260, fno3d.py:86-88)
//   project output_proj : [B, W, S] -> [B, 3, S]     (rational_fourier.py:274-276, fno3d.py:100-102)
// The reference permutes to channels-last, calls nn.Linear and permutes back (two 4.3 GB transposed
// copies at the 128^3 config); here the contraction runs directly on the [B, C, S] layout, one pass,
// fp32.  One side of the map has at most 8 channels ("small"), the other is streamed ("big").
#include "common.cuh"

namespace pdephi {

constexpr int PW_THREADS = 256;
constexpr int PW_MAXS = 8;     // max channels on the small side

// out[b,o,s] = bias[o] + sum_c W(o,c) x[b,c,s];  W(o,c) = TRANS ? Wt[c*Cout + o] : Wt[o*Cin + c].
// SMALL_IN: Cin <= 8 (inputs live in registers, outputs streamed); otherwise Cout <= 8 (accumulators in registers).
template <bool SMALL_IN, bool TRANS>
__global__ void __launch_bounds__(PW_THREADS)
pointwise_linear_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                        float* __restrict__ out, int Cin, int Cout, long long S4) {
  extern __shared__ float wsm[];   // [Cout][Cin] row-major (o, c) + bias[Cout]
  for (int i = threadIdx.x; i < Cin * Cout; i += PW_THREADS) {
    const int o = i / Cin, c = i - o * Cin;
    wsm[i] = TRANS ? __ldg(W + (size_t)c * Cout + o) : __ldg(W + i);
  }
  for (int i = threadIdx.x; i < Cout; i += PW_THREADS) wsm[Cin * Cout + i] = bias ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const float* bs = wsm + Cin * Cout;
  const size_t b = blockIdx.y;
  const float4* xb = reinterpret_cast<const float4*>(x) + b * (size_t)Cin * S4;
  float4* ob = reinterpret_cast<float4*>(out) + b * (size_t)Cout * S4;
  for (long long s = (long long)blockIdx.x * PW_THREADS + threadIdx.x; s < S4; s += (long long)gridDim.x * PW_THREADS) {
    if (SMALL_IN) {
      float4 xin[PW_MAXS];
#pragma unroll
      for (int c = 0; c < PW_MAXS; ++c)
        if (c < Cin) xin[c] = __ldcs(xb + (size_t)c * S4 + s);
      for (int o = 0; o < Cout; ++o) {
        const float bo = bs[o];
        float4 a = make_float4(bo, bo, bo, bo);
#pragma unroll
        for (int c = 0; c < PW_MAXS; ++c)
          if (c < Cin) {
            const float w = wsm[o * Cin + c];
            a.x = fmaf(w, xin[c].x, a.x); a.y = fmaf(w, xin[c].y, a.y);
            a.z = fmaf(w, xin[c].z, a.z); a.w = fmaf(w, xin[c].w, a.w);
          }
        ob[(size_t)o * S4 + s] = a;
      }
    } else {
      float4 acc[PW_MAXS];
#pragma unroll
      for (int o = 0; o < PW_MAXS; ++o) {
        const float bo = o < Cout ? bs[o] : 0.f;
        acc[o] = make_float4(bo, bo, bo, bo);
      }
#pragma unroll 4
      for (int c = 0; c < Cin; ++c) {
        const float4 v = __ldcs(xb + (size_t)c * S4 + s);
#pragma unroll
        for (int o = 0; o < PW_MAXS; ++o)
          if (o < Cout) {
            const float w = wsm[o * Cin + c];
            acc[o].x = fmaf(w, v.x, acc[o].x); acc[o].y = fmaf(w, v.y, acc[o].y);
            acc[o].z = fmaf(w, v.z, acc[o].z); acc[o].w = fmaf(w, v.w, acc[o].w);
          }
      }
#pragma unroll
      for (int o = 0; o < PW_MAXS; ++o)
        if (o < Cout) ob[(size_t)o * S4 + s] = acc[o];
    }
  }
}

// Weight / bias gradient: outer[big_ch][small_ch] = sum_{b,s} big[b,big_ch,s] * small[b,small_ch,s],
// sum_big[big_ch], sum_small[small_ch].  Stage 1: per-CTA partials (fixed order), stage 2: fixed-order sum.
constexpr int PG_BT = 8;       // big channels per CTA
template <int CS>              // channels on the small side, compile time so no predicated-off FMAs are issued
__global__ void __launch_bounds__(PW_THREADS, 2)   // 160 registers unbounded = one 8-warp CTA per SM: latency-bound at 3.5 TB/s
pointwise_wgrad_kernel(const float* __restrict__ big, const float* __restrict__ small, float* __restrict__ partial,
                       int Cbig, long long S4, int B, long long chunk4) {
  constexpr int Csmall = CS;
  __shared__ float red[PW_THREADS / 32][PG_BT * (PW_MAXS + 1) + PW_MAXS];
  const int g0 = blockIdx.y * PG_BT;
  const long long s_lo = (long long)blockIdx.x * chunk4, s_hi = min(S4, s_lo + chunk4);
  float acc[PG_BT][PW_MAXS + 1];
  float ssum[PW_MAXS];
#pragma unroll
  for (int i = 0; i < PG_BT; ++i)
#pragma unroll
    for (int j = 0; j <= PW_MAXS; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < PW_MAXS; ++j) ssum[j] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float4* bb = reinterpret_cast<const float4*>(big) + (size_t)b * Cbig * S4;
    const float4* sb = reinterpret_cast<const float4*>(small) + (size_t)b * Csmall * S4;
    for (long long s = s_lo + threadIdx.x; s < s_hi; s += PW_THREADS) {
      float4 sv[PW_MAXS];
#pragma unroll
      for (int j = 0; j < PW_MAXS; ++j)
        if (j < Csmall) {
          sv[j] = __ldg(sb + (size_t)j * S4 + s);
          ssum[j] += (sv[j].x + sv[j].y) + (sv[j].z + sv[j].w);
        }
      float4 bv[PG_BT];
#pragma unroll
      for (int i = 0; i < PG_BT; ++i) bv[i] = __ldcs(bb + (size_t)min(g0 + i, Cbig - 1) * S4 + s);   // all loads first
#pragma unroll
      for (int i = 0; i < PG_BT; ++i) {
        const float4 v = bv[i];
#pragma unroll
        for (int j = 0; j < PW_MAXS; ++j)
          if (j < Csmall) {
            float a = acc[i][j];
            a = fmaf(v.x, sv[j].x, a); a = fmaf(v.y, sv[j].y, a);
            a = fmaf(v.z, sv[j].z, a); a = fmaf(v.w, sv[j].w, a);
            acc[i][j] = a;
          }
        acc[i][PW_MAXS] += (v.x + v.y) + (v.z + v.w);
      }
    }
  }
  // block reduction, fixed order: shuffle tree inside a warp, then warp 0 sums the warps in order
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NV = PG_BT * (PW_MAXS + 1) + PW_MAXS;
#pragma unroll
  for (int i = 0; i < PG_BT; ++i)
#pragma unroll
    for (int j = 0; j <= PW_MAXS; ++j) {
      float v = acc[i][j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) red[w][i * (PW_MAXS + 1) + j] = v;
    }
#pragma unroll
  for (int j = 0; j < PW_MAXS; ++j) {
    float v = ssum[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) red[w][PG_BT * (PW_MAXS + 1) + j] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < PW_THREADS / 32; ++ww) t += red[ww][threadIdx.x];
    partial[((size_t)blockIdx.x * gridDim.y + blockIdx.y) * NV + threadIdx.x] = t;
  }
}

// dW[(big,small) or (small,big)], d_bias for whichever side is the output of the forward map
__global__ void pointwise_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW,
                                              float* __restrict__ dbias, int Cbig, int Csmall, int nchunks, int ngroups,
                                              int big_is_out) {
  constexpr int NV = PG_BT * (PW_MAXS + 1) + PW_MAXS;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = Cbig * (Csmall + 1) + Csmall;
  if (idx >= total) return;
  if (idx < Cbig * (Csmall + 1)) {
    const int cb = idx / (Csmall + 1), j = idx - cb * (Csmall + 1);
    const int grp = cb / PG_BT, i = cb - grp * PG_BT;
    const int slot = i * (PW_MAXS + 1) + (j < Csmall ? j : PW_MAXS);
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += (double)partial[((size_t)c * ngroups + grp) * NV + slot];
    if (j < Csmall) {
      // forward weight is [Cout][Cin]
      if (big_is_out) dW[(size_t)cb * Csmall + j] = (float)s;
      else dW[(size_t)j * Cbig + cb] = (float)s;
    } else if (big_is_out && dbias) {
      dbias[cb] = (float)s;
    }
  } else {
    const int j = idx - Cbig * (Csmall + 1);
    if (!big_is_out && dbias) {
      double s = 0.0;   // every big-channel group accumulated the same small-side sums; take group 0
      for (int c = 0; c < nchunks; ++c) s += (double)partial[((size_t)c * ngroups + 0) * NV + PG_BT * (PW_MAXS + 1) + j];
      dbias[j] = (float)s;
    }
  }
}

static void wgrad_geometry(int Cbig, long long S4, int* nchunks, long long* chunk4, int* ngroups) {
  *ngroups = (Cbig + PG_BT - 1) / PG_BT;
  long long want = (4 * 148 + *ngroups - 1) / *ngroups;
  const long long maxc = (S4 + PW_THREADS - 1) / PW_THREADS;
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  *chunk4 = (S4 + want - 1) / want;
  *nchunks = (int)((S4 + *chunk4 - 1) / *chunk4);
}

}  // namespace pdephi

using namespace pdephi;

extern "C" int pdephi_pointwise_linear(const float* x, const float* W, const float* bias, float* out, int B, int Cin,
                                       int Cout, long long S, int transpose_w, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(x && W && out, "pointwise_linear: null pointer");
  PDEPHI_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && S > 0, "pointwise_linear: non-positive size");
  PDEPHI_REQUIRE(B <= 65535, "pointwise_linear: batch > 65535");
  if (S % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(PDEPHI_ERR_UNSUPPORTED, "pointwise_linear: needs 16-byte aligned buffers and a spatial size divisible by 4");
  if (Cin > PW_MAXS && Cout > PW_MAXS)
    return fail(PDEPHI_ERR_UNSUPPORTED, "pointwise_linear: one side must have at most %d channels (got %d -> %d)", PW_MAXS, Cin, Cout);
  const long long S4 = S / 4;
  const size_t smem = ((size_t)Cin * Cout + Cout) * sizeof(float);
  long long gx = (S4 + PW_THREADS - 1) / PW_THREADS;
  if (gx > 148 * 16) gx = 148 * 16;
  dim3 grid((unsigned)gx, B);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(K_POINTWISE_LINEAR, st);
  if (Cin <= PW_MAXS) {
    if (transpose_w) pointwise_linear_kernel<true, true><<<grid, PW_THREADS, smem, st>>>(x, W, bias, out, Cin, Cout, S4);
    else pointwise_linear_kernel<true, false><<<grid, PW_THREADS, smem, st>>>(x, W, bias, out, Cin, Cout, S4);
  } else {
    if (transpose_w) pointwise_linear_kernel<false, true><<<grid, PW_THREADS, smem, st>>>(x, W, bias, out, Cin, Cout, S4);
    else pointwise_linear_kernel<false, false><<<grid, PW_THREADS, smem, st>>>(x, W, bias, out, Cin, Cout, S4);
  }
  PDEPHI_LAUNCH_CHECK("pointwise_linear_kernel");
  return PDEPHI_OK;
}

extern "C" size_t pdephi_pointwise_wgrad_workspace_bytes(int Cin, int Cout, long long S) {
  if (Cin <= 0 || Cout <= 0 || S <= 0) return 0;
  const int Cbig = Cin > Cout ? Cin : Cout;
  int nchunks, ngroups;
  long long chunk4;
  wgrad_geometry(Cbig, S / 4, &nchunks, &chunk4, &ngroups);
  return align_up((size_t)nchunks * ngroups * (PG_BT * (PW_MAXS + 1) + PW_MAXS) * sizeof(float), 256);
}

extern "C" int pdephi_pointwise_wgrad(const float* x, const float* g, float* dW, float* dbias, int B, int Cin, int Cout,
                                      long long S, void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(x && g && dW, "pointwise_wgrad: null pointer");
  PDEPHI_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && S > 0, "pointwise_wgrad: non-positive size");
  if (S % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(g) & 15))
    return fail(PDEPHI_ERR_UNSUPPORTED, "pointwise_wgrad: needs 16-byte aligned buffers and a spatial size divisible by 4");
  const bool big_is_out = Cout >= Cin;
  const int Cbig = big_is_out ? Cout : Cin, Csmall = big_is_out ? Cin : Cout;
  if (Csmall > PW_MAXS) return fail(PDEPHI_ERR_UNSUPPORTED, "pointwise_wgrad: one side must have at most %d channels", PW_MAXS);
  const size_t need = pdephi_pointwise_wgrad_workspace_bytes(Cin, Cout, S);
  if (!workspace || workspace_bytes < need) return fail(PDEPHI_ERR_WORKSPACE, "pointwise_wgrad: workspace %zu < %zu bytes", workspace_bytes, need);
  int nchunks, ngroups;
  long long chunk4;
  wgrad_geometry(Cbig, S / 4, &nchunks, &chunk4, &ngroups);
  cudaStream_t st = (cudaStream_t)stream;
  const float* big = big_is_out ? g : x;
  const float* small = big_is_out ? x : g;
  {
    ProfScope ps(K_POINTWISE_WGRAD, st);
    dim3 grid(nchunks, ngroups);
    float* part = (float*)workspace;
    switch (Csmall) {
#define CASE(n) case n: pointwise_wgrad_kernel<n><<<grid, PW_THREADS, 0, st>>>(big, small, part, Cbig, S / 4, B, chunk4); break;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
  }
  PDEPHI_LAUNCH_CHECK("pointwise_wgrad_kernel");
  const int total = Cbig * (Csmall + 1) + Csmall;
  {
    ProfScope ps(K_POINTWISE_WGRAD, st);
    pointwise_wgrad_reduce_kernel<<<(total + 127) / 128, 128, 0, st>>>((const float*)workspace, dW, dbias, Cbig, Csmall, nchunks, ngroups, big_is_out ? 1 : 0);
  }
  PDEPHI_LAUNCH_CHECK("pointwise_wgrad_reduce_kernel");
  return PDEPHI_OK;
}
