// Channels-first pointwise maps with up to 64 channels on BOTH sides, fp32 FFMA, any spatial size (SURVEY.md section 8f,
// rows N1 / N2 beyond the width-64 tensor-core kernels of block_tc.cu and the narrow maps of pointwise.cu):
//   wide_linear   out = W x + bias (+ addends)                         lift / project of MultiScaleFNO's inner operators
//                                                                      (in = out = width, models/multiscale_fno.py:100-127)
//   block tail    u = spec + (Wc h + bc) + h ;  h' = act(u)            rational_fourier.py:268-271, fno3d.py:93-97 at any
//   and backward  gu = g act'(u) ;  t = gu + Wc^T gu                   width <= 64 and for the activations getattr(F, name)
//                 dWc = sum gu (x) h ;  dbc = sum gu                   usually names (fno3d.py:57)
// One thread owns two neighbouring points and all output channels (accumulators in registers, weights broadcast from
// shared memory as float4); the residual is folded into the weight image as + I.  This is the fp32-exact route: it replaces
// cuDNN's convolution + eager activation / bias kernels (and their autograd) wherever the tcgen05 block kernels do not
// apply, at about half the FFMA peak -- those shapes are small or not on the headline path.
#include "common.cuh"

namespace pdephi {
namespace wide {

constexpr int THREADS = 128;
enum { ACT_IDENTITY = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_TANH = 3, ACT_SILU = 4, ACT_SIGMOID = 5, ACT_LEAKY_RELU = 6, ACT_ELU = 7,
       ACT_COUNT = 8 };

__device__ __forceinline__ float act_f(int a, float u) {
  switch (a) {
    case ACT_GELU: return 0.5f * u * (1.f + erff(u * 0.70710678118654752440f));
    case ACT_RELU: return fmaxf(u, 0.f);
    case ACT_TANH: return tanhf(u);
    case ACT_SILU: return u / (1.f + expf(-u));
    case ACT_SIGMOID: return 1.f / (1.f + expf(-u));
    case ACT_LEAKY_RELU: return u > 0.f ? u : 0.01f * u;
    case ACT_ELU: return u > 0.f ? u : expm1f(u);
    default: return u;
  }
}
__device__ __forceinline__ float act_grad_f(int a, float u) {
  switch (a) {
    case ACT_GELU: return 0.5f * (1.f + erff(u * 0.70710678118654752440f)) + u * 0.39894228040143267794f * expf(-0.5f * u * u);
    case ACT_RELU: return u > 0.f ? 1.f : 0.f;
    case ACT_TANH: { const float t = tanhf(u); return 1.f - t * t; }
    case ACT_SILU: { const float s = 1.f / (1.f + expf(-u)); return s * (1.f + u * (1.f - s)); }
    case ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-u)); return s * (1.f - s); }
    case ACT_LEAKY_RELU: return u > 0.f ? 1.f : 0.01f;
    case ACT_ELU: return u > 0.f ? 1.f : expf(u);
    default: return 1.f;
  }
}

// MODE 0: out = W x + bias + add0 + add1
// MODE 1: out = u = W' x + bias + add0 (W' = W + I: the residual), out2 = act(u)                       [block tail forward]
// MODE 2: gu = x * act'(uin) -> out2,  out = t = W'^T-style map of gu (W' = W + I, caller passes trans) [block tail backward]
// W(o,i) = trans ? Wt[i*Cout + o] : Wt[o*Cin + i].  CO >= Cout: accumulators per point.
template <int CO, int MODE>
__global__ void __launch_bounds__(THREADS)
wide_linear_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                   const float* __restrict__ add0, const float* __restrict__ add1, const float* __restrict__ uin,
                   float* __restrict__ out, float* __restrict__ out2, int Cin, int Cout, long long S, int trans, int act) {
  extern __shared__ float wsm[];          // [Cin][CO] (i, o), zero beyond Cout, then bias[CO]
  for (int idx = threadIdx.x; idx < Cin * CO; idx += THREADS) {
    const int i = idx / CO, o = idx - i * CO;
    float w = 0.f;
    if (o < Cout) {
      w = trans ? __ldg(W + (size_t)i * Cout + o) : __ldg(W + (size_t)o * Cin + i);
      if (MODE != 0 && i == o) w += 1.f;
    }
    wsm[idx] = w;
  }
  for (int o = threadIdx.x; o < CO; o += THREADS) wsm[Cin * CO + o] = (bias && o < Cout) ? __ldg(bias + o) : 0.f;
  __syncthreads();
  const float* bs = wsm + Cin * CO;
  const size_t b = blockIdx.y;
  const float* xb = x + b * (size_t)Cin * S;
  const float* ub = MODE == 2 ? uin + b * (size_t)Cin * S : nullptr;
  float* ob = out + b * (size_t)Cout * S;
  float* o2b = MODE == 1 ? out2 + b * (size_t)Cout * S : (MODE == 2 ? out2 + b * (size_t)Cin * S : nullptr);
  const float* a0b = add0 ? add0 + b * (size_t)Cout * S : nullptr;
  const float* a1b = add1 ? add1 + b * (size_t)Cout * S : nullptr;
  const bool vec = (S & 1) == 0;          // rows stay 8-byte aligned: float2 accesses
  const long long npair = (S + 1) / 2;
  for (long long p = (long long)blockIdx.x * THREADS + threadIdx.x; p < npair; p += (long long)gridDim.x * THREADS) {
    const long long s = 2 * p;
    const bool two = s + 1 < S;
    float a0[CO], a1[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) { a0[o] = bs[o]; a1[o] = bs[o]; }
    for (int i = 0; i < Cin; ++i) {
      float v0, v1 = 0.f;
      if (vec) { const float2 t = __ldg(reinterpret_cast<const float2*>(xb + (size_t)i * S + s)); v0 = t.x; v1 = t.y; }
      else { v0 = __ldg(xb + (size_t)i * S + s); if (two) v1 = __ldg(xb + (size_t)i * S + s + 1); }
      if (MODE == 2) {
        float u0, u1 = 0.f;
        if (vec) { const float2 t = __ldg(reinterpret_cast<const float2*>(ub + (size_t)i * S + s)); u0 = t.x; u1 = t.y; }
        else { u0 = __ldg(ub + (size_t)i * S + s); if (two) u1 = __ldg(ub + (size_t)i * S + s + 1); }
        v0 *= act_grad_f(act, u0);
        v1 *= act_grad_f(act, u1);
        if (vec) *reinterpret_cast<float2*>(o2b + (size_t)i * S + s) = make_float2(v0, v1);
        else { o2b[(size_t)i * S + s] = v0; if (two) o2b[(size_t)i * S + s + 1] = v1; }
      }
      const float4* wr = reinterpret_cast<const float4*>(wsm + i * CO);
#pragma unroll
      for (int o4 = 0; o4 < CO / 4; ++o4) {
        const float4 w = wr[o4];
        a0[4 * o4] = fmaf(w.x, v0, a0[4 * o4]);         a1[4 * o4] = fmaf(w.x, v1, a1[4 * o4]);
        a0[4 * o4 + 1] = fmaf(w.y, v0, a0[4 * o4 + 1]); a1[4 * o4 + 1] = fmaf(w.y, v1, a1[4 * o4 + 1]);
        a0[4 * o4 + 2] = fmaf(w.z, v0, a0[4 * o4 + 2]); a1[4 * o4 + 2] = fmaf(w.z, v1, a1[4 * o4 + 2]);
        a0[4 * o4 + 3] = fmaf(w.w, v0, a0[4 * o4 + 3]); a1[4 * o4 + 3] = fmaf(w.w, v1, a1[4 * o4 + 3]);
      }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      if (o < Cout) {
        float r0 = a0[o], r1 = a1[o];
        const size_t off = (size_t)o * S + s;
        if (a0b) { r0 += __ldg(a0b + off); if (two) r1 += __ldg(a0b + off + 1); }
        if (a1b) { r0 += __ldg(a1b + off); if (two) r1 += __ldg(a1b + off + 1); }
        if (vec) *reinterpret_cast<float2*>(ob + off) = make_float2(r0, r1);
        else { ob[off] = r0; if (two) ob[off + 1] = r1; }
        if (MODE == 1) {
          const float h0 = act_f(act, r0), h1 = act_f(act, r1);
          if (vec) *reinterpret_cast<float2*>(o2b + off) = make_float2(h0, h1);
          else { o2b[off] = h0; if (two) o2b[off + 1] = h1; }
        }
      }
    }
  }
}

// dW[o][i] = sum_{b,s} g[b,o,s] x[b,i,s],  db[o] = sum g.  A CTA of 256 threads walks chunks of PT points of one batch item:
// the g and x tiles go through shared memory, thread (to, ti) owns the 4 x 4 block (o = 4 to .., i = 4 ti ..) of a 64 x 64 tile.
constexpr int WG_THREADS = 256;
constexpr int PT = 32;
constexpr int PLD = PT + 4;
__global__ void __launch_bounds__(WG_THREADS)
wide_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ partial, int Cin, int Cout,
                  long long S, int B, long long chunk) {
  __shared__ __align__(16) float gs[64 * PLD];
  __shared__ __align__(16) float xs[64 * PLD];
  const int to = threadIdx.x >> 4, ti = threadIdx.x & 15;
  float acc[4][4];
  float gsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  const int b = blockIdx.y;
  const long long s_lo = (long long)blockIdx.x * chunk, s_hi = min(S, s_lo + chunk);
  for (long long s0 = s_lo; s0 < s_hi; s0 += PT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * PT; idx += WG_THREADS) {
      const int r = idx / PT, p = idx - r * PT;
      const long long s = s0 + p;
      gs[r * PLD + p] = (r < Cout && s < s_hi) ? __ldg(g + ((size_t)b * Cout + r) * S + s) : 0.f;
      xs[r * PLD + p] = (r < Cin && s < s_hi) ? __ldg(x + ((size_t)b * Cin + r) * S + s) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < PT; p += 4) {
      float4 gv[4], xv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) gv[a] = *reinterpret_cast<const float4*>(gs + (4 * to + a) * PLD + p);
#pragma unroll
      for (int c = 0; c < 4; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (4 * ti + c) * PLD + p);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t = acc[a][c];
          t = fmaf(gv[a].x, xv[c].x, t); t = fmaf(gv[a].y, xv[c].y, t);
          t = fmaf(gv[a].z, xv[c].z, t); t = fmaf(gv[a].w, xv[c].w, t);
          acc[a][c] = t;
        }
        if (ti == 0) gsum[a] += (gv[a].x + gv[a].y) + (gv[a].z + gv[a].w);
      }
    }
  }
  float* pout = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (64 * 64 + 64);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int c = 0; c < 4; ++c) pout[(4 * to + a) * 64 + 4 * ti + c] = acc[a][c];
    if (ti == 0) pout[64 * 64 + 4 * to + a] = gsum[a];
  }
}

__global__ void wide_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, float* __restrict__ db, int Cin,
                                         int Cout, int nparts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 64 + 64) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)partial[(size_t)p * (64 * 64 + 64) + idx];
  if (idx < 64 * 64) {
    const int o = idx >> 6, i = idx & 63;
    if (o < Cout && i < Cin) dW[(size_t)o * Cin + i] = (float)s;
  } else if (db && idx - 64 * 64 < Cout) {
    db[idx - 64 * 64] = (float)s;
  }
}

static void wgrad_geometry(long long S, int B, int* nchunks, long long* chunk) {
  long long want = (4 * 148 + B - 1) / B;
  const long long maxc = (S + PT - 1) / PT;
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  long long c = (S + want - 1) / want;
  c = (c + PT - 1) / PT * PT;
  *chunk = c;
  *nchunks = (int)((S + c - 1) / c);
}

template <int MODE>
static int launch_linear(const float* x, const float* W, const float* bias, const float* add0, const float* add1, const float* uin,
                         float* out, float* out2, int B, int Cin, int Cout, long long S, int trans, int act, int kid,
                         cudaStream_t st, const char* who) {
  PDEPHI_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && S > 0, "%s: non-positive size", who);
  PDEPHI_REQUIRE(B <= 65535, "%s: batch > 65535", who);
  if (Cin > 64 || Cout > 64) return fail(PDEPHI_ERR_UNSUPPORTED, "%s: at most 64 channels per side (got %d -> %d)", who, Cin, Cout);
  if (act < 0 || act >= ACT_COUNT) return fail(PDEPHI_ERR_UNSUPPORTED, "%s: unknown activation code %d", who, act);
  if (MODE != 0 && Cin != Cout) return fail(PDEPHI_ERR_INVALID, "%s: the block tail needs Cin == Cout", who);
  const int CO = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
  const size_t smem = ((size_t)Cin * CO + CO) * sizeof(float);
  long long gx = ((S + 1) / 2 + THREADS - 1) / THREADS;
  const long long cap = (148LL * 16 + B - 1) / B;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, B);
  ProfScope ps(kid, st);
  switch (CO) {
    case 16: wide_linear_kernel<16, MODE><<<grid, THREADS, smem, st>>>(x, W, bias, add0, add1, uin, out, out2, Cin, Cout, S, trans, act); break;
    case 32: wide_linear_kernel<32, MODE><<<grid, THREADS, smem, st>>>(x, W, bias, add0, add1, uin, out, out2, Cin, Cout, S, trans, act); break;
    default: wide_linear_kernel<64, MODE><<<grid, THREADS, smem, st>>>(x, W, bias, add0, add1, uin, out, out2, Cin, Cout, S, trans, act); break;
  }
  PDEPHI_LAUNCH_CHECK("wide_linear_kernel");
  return PDEPHI_OK;
}

}  // namespace wide
}  // namespace pdephi

using namespace pdephi;

extern "C" int pdephi_wide_linear(const float* x, const float* W, const float* bias, const float* add0, const float* add1,
                                  float* out, int B, int Cin, int Cout, long long S, int transpose_w, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(x && W && out, "wide_linear: null pointer");
  return wide::launch_linear<0>(x, W, bias, add0, add1, nullptr, out, nullptr, B, Cin, Cout, S, transpose_w, 0, K_POINTWISE_LINEAR,
                                (cudaStream_t)stream, "wide_linear");
}

extern "C" int pdephi_wide_block_tail_fwd(const float* h, const float* spec, const float* Wc, const float* bc, float* u, float* hout,
                                          int B, int C, long long S, int activation, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(h && spec && Wc && u && hout, "wide_block_tail_fwd: null pointer");
  return wide::launch_linear<1>(h, Wc, bc, spec, nullptr, nullptr, u, hout, B, C, C, S, 0, activation, K_BLOCK_TAIL_FWD,
                                (cudaStream_t)stream, "wide_block_tail_fwd");
}

extern "C" size_t pdephi_wide_wgrad_workspace_bytes(int B, long long S) {
  if (B <= 0 || S <= 0) return 0;
  int nchunks;
  long long chunk;
  wide::wgrad_geometry(S, B, &nchunks, &chunk);
  return align_up((size_t)nchunks * B * (64 * 64 + 64) * sizeof(float), 256);
}

extern "C" int pdephi_wide_wgrad(const float* x, const float* g, float* dW, float* dbias, int B, int Cin, int Cout, long long S,
                                 void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(x && g && dW, "wide_wgrad: null pointer");
  PDEPHI_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && S > 0, "wide_wgrad: non-positive size");
  PDEPHI_REQUIRE(B <= 65535, "wide_wgrad: batch > 65535");
  if (Cin > 64 || Cout > 64) return fail(PDEPHI_ERR_UNSUPPORTED, "wide_wgrad: at most 64 channels per side (got %d -> %d)", Cin, Cout);
  const size_t need = pdephi_wide_wgrad_workspace_bytes(B, S);
  if (!workspace || workspace_bytes < need) return fail(PDEPHI_ERR_WORKSPACE, "wide_wgrad: workspace %zu < %zu bytes", workspace_bytes, need);
  int nchunks;
  long long chunk;
  wide::wgrad_geometry(S, B, &nchunks, &chunk);
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope ps(K_POINTWISE_WGRAD, st);
    wide::wide_wgrad_kernel<<<dim3(nchunks, B), wide::WG_THREADS, 0, st>>>(x, g, (float*)workspace, Cin, Cout, S, B, chunk);
  }
  PDEPHI_LAUNCH_CHECK("wide_wgrad_kernel");
  {
    ProfScope ps(K_POINTWISE_WGRAD, st);
    wide::wide_wgrad_reduce_kernel<<<(64 * 64 + 64 + 127) / 128, 128, 0, st>>>((const float*)workspace, dW, dbias, Cin, Cout, nchunks * B);
  }
  PDEPHI_LAUNCH_CHECK("wide_wgrad_reduce_kernel");
  return PDEPHI_OK;
}

extern "C" int pdephi_wide_block_tail_bwd(const float* g, const float* u, const float* h, const float* Wc, float* gu, float* t,
                                          float* dWc, float* dbc, int B, int C, long long S, int activation, void* workspace,
                                          size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(g && u && h && Wc && gu && t && dWc, "wide_block_tail_bwd: null pointer");
  // gu = g act'(u) (stored), t = gu + Wc^T gu: the map gu -> t has weight W(c, o) = Wc[o][c], i.e. the transposed read
  int rc = wide::launch_linear<2>(g, Wc, nullptr, nullptr, nullptr, u, t, gu, B, C, C, S, 1, activation, K_BLOCK_TAIL_BWD,
                                  (cudaStream_t)stream, "wide_block_tail_bwd");
  if (rc) return rc;
  return pdephi_wide_wgrad(h, gu, dWc, dbc, B, C, C, S, workspace, workspace_bytes, stream);
}
