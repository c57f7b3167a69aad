// K2: per-mode channel mix kernels.
//
//  * rational mix (RationalFourierLayer.rational_multiply, rational_fourier.py:73-119): the real
//    transfer function R[o,i,k] = P(k)/(Q(k)+eps) is evaluated on the fly, in registers, with the
//    reference's exact rounding order (one rounded product and one rounded add per monomial, IEEE
//    division; rational_fourier.py:135-146,101,104), so P, Q and R are bit-identical to the
//    reference's and the 285 MB R tensor is never materialised.
//  * StabilityProjection statistics and backward (stability.py:59-103).
//  * complex mix of SpectralConv3D (spectral_layers.py:66).
//
// All reductions are fixed-order (no float atomics): the reference's tests require bit-repeatable
// losses (tests/test_comprehensive.py:575-647).
#include "common.cuh"

namespace pdephi {

constexpr int MAX_T = 56;  // monomials of order 6
struct MonoOffsets {
  int p[MAX_T];
  int q[MAX_T];
};
static inline int tri(int ord) { return ord * (ord + 1) * (ord + 2) / 6; }
template <int ORD>
struct Tri {
  static constexpr int T = ORD * (ORD + 1) * (ORD + 2) / 6;
  static constexpr int TP4 = (T + 3) / 4 * 4;
};
static void fill_offsets(int ord, int* off) {
  int n = 0;
  for (int a = 0; a < ord; ++a)
    for (int b = 0; b < ord; ++b)
      for (int c = 0; c < ord; ++c)
        if (a + b + c < ord) off[n++] = (a * ord + b) * ord + c;
}

// R = P/(Q+eps) with the reference's stepwise fp32 rounding.  cp/cq: coefficients, mp/mq: monomials.
template <int TPN, int TQN>
__device__ __forceinline__ float eval_R(const float* cp, const float* cq, const float* mp, const float* mq, float eps,
                                        float* qe_out) {
  float p = 0.f, q = 0.f;
#pragma unroll
  for (int t = 0; t < TPN; ++t) p = __fadd_rn(p, __fmul_rn(cp[t], mp[t]));
#pragma unroll
  for (int t = 0; t < TQN; ++t) q = __fadd_rn(q, __fmul_rn(cq[t], mq[t]));
  const float qe = __fadd_rn(q, eps);
  if (qe_out) *qe_out = qe;
  return __fdiv_rn(p, qe);
}

// Two transfer functions at once (output channels u and u+1, same mode k) on Blackwell's packed fp32 pipe.  The
// reference's rounding sequence needs a separately rounded product and a separately rounded add per monomial, and
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (also under -fmad=false; it even folds fma(a, 1, b) back
// to an add first).  So both steps are written as FMAs whose constant operands are RUNTIME values the compiler cannot
// see through:  term = fma(c, m, -0) rounds c*m once (and keeps the sign of a zero product), acc = fma(acc, 1, term)
// rounds acc + term once -- bit for bit __fmul_rn / __fadd_rn, at half the instruction count.
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// cp2 / cq2: coefficient pairs (c_u[t], c_{u+1}[t]); one2 = (1, 1), nz2 = (-0, -0) from kernel arguments
template <int TPN, int TQN>
__device__ __forceinline__ void eval_R2(const uint64_t* cp2, const uint64_t* cq2, const float* mp, const float* mq, float eps,
                                        uint64_t one2, uint64_t nz2, float* R, float* qe) {
  uint64_t p = pk2(0.f, 0.f), q = pk2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < TPN; ++t) p = fma2(p, one2, fma2(cp2[t], pk2(mp[t], mp[t]), nz2));
#pragma unroll
  for (int t = 0; t < TQN; ++t) q = fma2(q, one2, fma2(cq2[t], pk2(mq[t], mq[t]), nz2));
  q = fma2(q, one2, pk2(eps, eps));
  float p0, p1;
  upk2(p, p0, p1);
  upk2(q, qe[0], qe[1]);
  R[0] = __fdiv_rn(p0, qe[0]);
  R[1] = __fdiv_rn(p1, qe[1]);
}

// ---------------------------------------------------------------------------------------------
// out[b][u][k] = sum_v R(u,v)[k] * in[b][v][k]     (TR=false: (o,i)=(u,v);  TR=true: (o,i)=(v,u))
// thread = one mode k (coalesced along k), UT output channels x BT batch entries per thread.
// ---------------------------------------------------------------------------------------------
constexpr int MIX_THREADS = 128;
constexpr int UT = 4;
constexpr int BT = 8;

// (Forcing four or five CTAs per SM with __launch_bounds__ was measured: 128 / 96 registers, 573 / 933 us against 497 us at
// the compiler's own 158 registers and three CTAs.)
template <int OP, int OQ, bool TR, int BTV = BT>   // BTV: batch entries per thread (fewer for small batches)
__global__ void __launch_bounds__(MIX_THREADS)
rational_mix_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float* __restrict__ P,
                    const float* __restrict__ Q, const float* __restrict__ monoP, const float* __restrict__ monoQ,
                    float eps, float* __restrict__ Rout, float* __restrict__ Qeout, int B, int U, int V, long long M,
                    MonoOffsets offs, float one, float negzero) {
  constexpr int TPN = Tri<OP>::T, TQN = Tri<OQ>::T;
  constexpr int TP4 = Tri<OP>::TP4, TQ4 = Tri<OQ>::TP4;
  extern __shared__ __align__(16) float smem[];
  float* cPs = smem;                            // [UT/2][V][TP4][2]: coefficient pairs of output channels (2p, 2p+1)
  float* cQs = smem + (size_t)UT * V * TP4;     // [UT/2][V][TQ4][2]
  const uint64_t one2 = pk2(one, one), nz2 = pk2(negzero, negzero);
  const int u0 = blockIdx.y * UT;
  const int O = TR ? V : U, I = TR ? U : V;
  (void)O;
  for (int idx = threadIdx.x; idx < UT * V * TP4; idx += MIX_THREADS) {
    const int t = idx % TP4, uv = idx / TP4, v = uv % V, ul = uv / V;
    const int u = min(u0 + ul, U - 1);
    const int o = TR ? v : u, i = TR ? u : v;
    cPs[((((size_t)(ul >> 1) * V + v) * TP4 + t) << 1) + (ul & 1)] =
        (t < TPN) ? __ldg(P + ((size_t)o * I + i) * (OP * OP * OP) + offs.p[t]) : 0.f;
  }
  for (int idx = threadIdx.x; idx < UT * V * TQ4; idx += MIX_THREADS) {
    const int t = idx % TQ4, uv = idx / TQ4, v = uv % V, ul = uv / V;
    const int u = min(u0 + ul, U - 1);
    const int o = TR ? v : u, i = TR ? u : v;
    cQs[((((size_t)(ul >> 1) * V + v) * TQ4 + t) << 1) + (ul & 1)] =
        (t < TQN) ? __ldg(Q + ((size_t)o * I + i) * (OQ * OQ * OQ) + offs.q[t]) : 0.f;
  }
  __syncthreads();

  const long long k = (long long)blockIdx.x * MIX_THREADS + threadIdx.x;
  const long long kc = k < M ? k : M - 1;
  float mp[TP4], mqs[(OP == OQ) ? 1 : TQ4];
#pragma unroll
  for (int t = 0; t < TP4; ++t) mp[t] = (t < TPN) ? __ldg(monoP + (size_t)t * M + kc) : 0.f;
  if (OP != OQ) {
#pragma unroll
    for (int t = 0; t < TQ4; ++t) mqs[t] = (t < TQN) ? __ldg(monoQ + (size_t)t * M + kc) : 0.f;
  }
  const float* mq = (OP == OQ) ? mp : mqs;

  for (int b0 = 0; b0 < B; b0 += BTV) {
    float2 acc[UT][BTV];
#pragma unroll
    for (int ul = 0; ul < UT; ++ul)
#pragma unroll
      for (int j = 0; j < BTV; ++j) acc[ul][j] = make_float2(0.f, 0.f);
    for (int v = 0; v < V; ++v) {
      float2 xin[BTV];
#pragma unroll
      for (int j = 0; j < BTV; ++j)
        xin[j] = (b0 + j < B) ? __ldg(in + ((size_t)(b0 + j) * V + v) * M + kc) : make_float2(0.f, 0.f);
#pragma unroll
      for (int up = 0; up < UT / 2; ++up) {
        uint64_t cp2[TP4], cq2[TQ4];
        const float4* p4 = reinterpret_cast<const float4*>(cPs + (((size_t)up * V + v) * TP4) * 2);
        const float4* q4 = reinterpret_cast<const float4*>(cQs + (((size_t)up * V + v) * TQ4) * 2);
#pragma unroll
        for (int t = 0; t < TP4 / 2; ++t) {
          const float4 c = p4[t];
          cp2[2 * t] = pk2(c.x, c.y); cp2[2 * t + 1] = pk2(c.z, c.w);
        }
#pragma unroll
        for (int t = 0; t < TQ4 / 2; ++t) {
          const float4 c = q4[t];
          cq2[2 * t] = pk2(c.x, c.y); cq2[2 * t + 1] = pk2(c.z, c.w);
        }
        float R2[2], qe2[2];
        eval_R2<TPN, TQN>(cp2, cq2, mp, mq, eps, one2, nz2, R2, qe2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ul = 2 * up + h;
          if (!TR && Rout && b0 == 0 && k < M && u0 + ul < U) {   // materialise R and Q+eps once for the backward
            const size_t ri = ((size_t)(u0 + ul) * V + v) * M + k;
            Rout[ri] = R2[h];
            Qeout[ri] = qe2[h];
          }
#pragma unroll
          for (int j = 0; j < BTV; ++j) {
            acc[ul][j].x = fmaf(R2[h], xin[j].x, acc[ul][j].x);
            acc[ul][j].y = fmaf(R2[h], xin[j].y, acc[ul][j].y);
          }
        }
      }
    }
    if (k < M) {
#pragma unroll
      for (int ul = 0; ul < UT; ++ul)
#pragma unroll
        for (int j = 0; j < BTV; ++j)
          if (u0 + ul < U && b0 + j < B) out[((size_t)(b0 + j) * U + u0 + ul) * M + k] = acc[ul][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// R and Q+eps alone (no mix): the transfer function depends on the coefficients only, not on the input, so a forward with
// unchanged coefficients (inference, the 4 evaluations of an RK4 rollout step, gradient accumulation) reuses the materialised
// copies (functional.py: TransferCache) and runs the streaming mix below instead of re-evaluating 2 x 20 monomials per
// (o, i, k).  Same evaluation, bit for bit, as rational_mix_kernel.  `changed`: device flag written by params_changed_kernel;
// 0 = the cached copies are still valid, nothing to do (the guard costs two few-microsecond launches).
// ---------------------------------------------------------------------------------------------
template <int OP, int OQ>
__global__ void __launch_bounds__(MIX_THREADS)
rational_transfer_kernel(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ monoP,
                         const float* __restrict__ monoQ, float eps, float* __restrict__ Rout, float* __restrict__ Qeout, int U,
                         int V, long long M, MonoOffsets offs, float one, float negzero, const int* __restrict__ changed) {
  if (changed && *changed == 0) return;
  constexpr int TPN = Tri<OP>::T, TQN = Tri<OQ>::T;
  constexpr int TP4 = Tri<OP>::TP4, TQ4 = Tri<OQ>::TP4;
  extern __shared__ __align__(16) float smem[];
  float* cPs = smem;                            // [UT/2][V][TP4][2]
  float* cQs = smem + (size_t)UT * V * TP4;     // [UT/2][V][TQ4][2]
  const uint64_t one2 = pk2(one, one), nz2 = pk2(negzero, negzero);
  const int u0 = blockIdx.y * UT;
  for (int idx = threadIdx.x; idx < UT * V * TP4; idx += MIX_THREADS) {
    const int t = idx % TP4, uv = idx / TP4, v = uv % V, ul = uv / V;
    const int u = min(u0 + ul, U - 1);
    cPs[((((size_t)(ul >> 1) * V + v) * TP4 + t) << 1) + (ul & 1)] =
        (t < TPN) ? __ldg(P + ((size_t)u * V + v) * (OP * OP * OP) + offs.p[t]) : 0.f;
  }
  for (int idx = threadIdx.x; idx < UT * V * TQ4; idx += MIX_THREADS) {
    const int t = idx % TQ4, uv = idx / TQ4, v = uv % V, ul = uv / V;
    const int u = min(u0 + ul, U - 1);
    cQs[((((size_t)(ul >> 1) * V + v) * TQ4 + t) << 1) + (ul & 1)] =
        (t < TQN) ? __ldg(Q + ((size_t)u * V + v) * (OQ * OQ * OQ) + offs.q[t]) : 0.f;
  }
  __syncthreads();
  const long long k = (long long)blockIdx.x * MIX_THREADS + threadIdx.x;
  const long long kc = k < M ? k : M - 1;
  float mp[TP4], mqs[(OP == OQ) ? 1 : TQ4];
#pragma unroll
  for (int t = 0; t < TP4; ++t) mp[t] = (t < TPN) ? __ldg(monoP + (size_t)t * M + kc) : 0.f;
  if (OP != OQ) {
#pragma unroll
    for (int t = 0; t < TQ4; ++t) mqs[t] = (t < TQN) ? __ldg(monoQ + (size_t)t * M + kc) : 0.f;
  }
  const float* mq = (OP == OQ) ? mp : mqs;
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int up = 0; up < UT / 2; ++up) {
      uint64_t cp2[TP4], cq2[TQ4];
      const float4* p4 = reinterpret_cast<const float4*>(cPs + (((size_t)up * V + v) * TP4) * 2);
      const float4* q4 = reinterpret_cast<const float4*>(cQs + (((size_t)up * V + v) * TQ4) * 2);
#pragma unroll
      for (int t = 0; t < TP4 / 2; ++t) {
        const float4 c = p4[t];
        cp2[2 * t] = pk2(c.x, c.y); cp2[2 * t + 1] = pk2(c.z, c.w);
      }
#pragma unroll
      for (int t = 0; t < TQ4 / 2; ++t) {
        const float4 c = q4[t];
        cq2[2 * t] = pk2(c.x, c.y); cq2[2 * t + 1] = pk2(c.z, c.w);
      }
      float R2[2], qe2[2];
      eval_R2<TPN, TQN>(cp2, cq2, mp, mq, eps, one2, nz2, R2, qe2);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ul = 2 * up + h;
        if (k < M && u0 + ul < U) {
          const size_t ri = ((size_t)(u0 + ul) * V + v) * M + k;
          Rout[ri] = R2[h];
          Qeout[ri] = qe2[h];
        }
      }
    }
  }
}

// flag |= (cur != snap) bitwise, snap = cur: has a coefficient tensor changed since the transfer function was materialised?
// (A host-side version counter misses `param.data` writes such as the reference's broadcast_parameters,
// training/distributed.py:230.)
__global__ void params_changed_kernel(const uint32_t* __restrict__ curP, uint32_t* __restrict__ snapP, long long nP,
                                      const uint32_t* __restrict__ curQ, uint32_t* __restrict__ snapQ, long long nQ,
                                      int* __restrict__ changed) {
  bool diff = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nP + nQ; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t c = i < nP ? __ldg(curP + i) : __ldg(curQ + (i - nP));
    uint32_t* s = i < nP ? snapP + i : snapQ + (i - nP);
    if (*s != c) { diff = true; *s = c; }
  }
  if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(changed, 1);
}

// ---------------------------------------------------------------------------------------------
// StabilityProjection statistics: one CTA per (b,o) row.  stats = {a, b, s, 0}
// ---------------------------------------------------------------------------------------------
constexpr int ROW_THREADS = 256;
__global__ void __launch_bounds__(ROW_THREADS)
projection_stats_kernel(const float2* __restrict__ Y, const float* __restrict__ mask, float proj_eps,
                        float* __restrict__ stats, long long M) {
  __shared__ double red[ROW_THREADS / 32];
  const size_t row = blockIdx.x;
  const float2* y = Y + row * (size_t)M;
  float e0 = 0.f, e1 = 0.f;
  double E0 = 0.0, E1 = 0.0;
  int n = 0;
  for (long long k = threadIdx.x; k < M; k += ROW_THREADS) {
    const float2 v = __ldg(y + k);
    const float m = __ldg(mask + k);
    const float a2 = fmaf(v.x, v.x, v.y * v.y);
    e0 += a2;
    e1 = fmaf(m * m, a2, e1);
    if (++n == 64) { E0 += e0; E1 += e1; e0 = e1 = 0.f; n = 0; }  // short fp32 runs, fp64 carry
  }
  E0 += e0;
  E1 += e1;
  E0 = block_sum<ROW_THREADS>(E0, red);
  E1 = block_sum<ROW_THREADS>(E1, red);
  if (threadIdx.x == 0) {
    const float a = (float)E0 + proj_eps, b = (float)E1 + proj_eps;
    stats[row * 4 + 0] = a;
    stats[row * 4 + 1] = b;
    stats[row * 4 + 2] = sqrtf(a / b);
    stats[row * 4 + 3] = 0.f;
  }
}

// backward of Z = mask*s*Y:  dY = m s dZ + G s (1/a - m^2/b) Y,  G = sum_k Re(conj(dZ) m Y)
__global__ void __launch_bounds__(ROW_THREADS)
projection_bwd_kernel(const float2* __restrict__ dZ, const float2* __restrict__ Y, const float* __restrict__ mask,
                      const float* __restrict__ stats, float2* __restrict__ dY, long long M, float* __restrict__ Gext, int phase) {
  // phase 0: whole row on this GPU;  1: partial G of this mode shard -> Gext[row], nothing else;  2: G = Gext[row] (summed)
  __shared__ double red[ROW_THREADS / 32];
  const size_t row = blockIdx.x;
  const float2* y = Y + row * (size_t)M;
  const float2* dz = dZ + row * (size_t)M;
  double Gd = 0.0;
  if (phase != 2) {
    float g = 0.f;
    int n = 0;
    for (long long k = threadIdx.x; k < M; k += ROW_THREADS) {
      const float2 v = __ldg(y + k), d = __ldg(dz + k);
      g = fmaf(__ldg(mask + k), fmaf(d.x, v.x, d.y * v.y), g);
      if (++n == 64) { Gd += g; g = 0.f; n = 0; }
    }
    Gd += g;
    Gd = block_sum<ROW_THREADS>(Gd, red);
    if (phase == 1) {
      if (threadIdx.x == 0) Gext[row] = (float)Gd;
      return;
    }
  } else {
    Gd = (double)Gext[row];
  }
  const float a = stats[row * 4 + 0], b = stats[row * 4 + 1], s = stats[row * 4 + 2];
  const float Gs = (float)Gd * s;
  const float ia = 1.f / a, ib = 1.f / b;
  float2* o = dY + row * (size_t)M;
  for (long long k = threadIdx.x; k < M; k += ROW_THREADS) {
    const float2 v = __ldg(y + k), d = __ldg(dz + k);
    const float m = __ldg(mask + k);
    const float c0 = m * s, c1 = Gs * (ia - m * m * ib);
    o[k] = make_float2(fmaf(c0, d.x, c1 * v.x), fmaf(c0, d.y, c1 * v.y));
  }
}

// ---------------------------------------------------------------------------------------------
// dX[b][i][k] = sum_o R[o][i][k] dY[b][o][k] from the materialised R: a streaming kernel, R is read
// exactly once (thread = mode k, IT input channels x BT batch entries per thread).
// ---------------------------------------------------------------------------------------------
constexpr int IT = 4;
// BTV: batch entries per thread (8 for training batches; 1 / 2 / 4 for the small batches of the 256^3 configurations,
// where a fixed 8 spent 7/8 of the FMAs and registers on zeros)
// FWD: the forward mix from a materialised (cached) transfer function, Y[b][o][k] = sum_i R[o][i][k] X[b][i][k] -- the same
// kernel with the roles of the channel indices exchanged (dY = X, dX = Y, "I" = outputs, "O" = contracted inputs); the
// accumulation order over the contracted index is that of rational_mix_kernel, so Y is bit-identical to the fused evaluation.
template <int BTV, bool FWD = false>
__global__ void __launch_bounds__(MIX_THREADS, 4)
mix_from_R_kernel(const float2* __restrict__ dY, float2* __restrict__ dX, const float* __restrict__ R, int B, int I,
                  int O, long long M) {
  const long long k = (long long)blockIdx.x * MIX_THREADS + threadIdx.x;
  const long long kc = k < M ? k : M - 1;
  const int i0 = blockIdx.y * IT;
  for (int b0 = 0; b0 < B; b0 += BTV) {
    float2 acc[IT][BTV];
#pragma unroll
    for (int a = 0; a < IT; ++a)
#pragma unroll
      for (int j = 0; j < BTV; ++j) acc[a][j] = make_float2(0.f, 0.f);
#pragma unroll(BTV >= 8 ? 2 : 8)      // few batch entries: more o steps in flight instead
    for (int o = 0; o < O; ++o) {
      float r[IT];
#pragma unroll
      for (int a = 0; a < IT; ++a)
        r[a] = __ldg(R + (FWD ? ((size_t)min(i0 + a, I - 1) * O + o) : ((size_t)o * I + min(i0 + a, I - 1))) * M + kc);
      float2 g[BTV];
#pragma unroll
      for (int j = 0; j < BTV; ++j)
        g[j] = (b0 + j < B) ? __ldg(dY + ((size_t)(b0 + j) * O + o) * M + kc) : make_float2(0.f, 0.f);
#pragma unroll
      for (int a = 0; a < IT; ++a)
#pragma unroll
        for (int j = 0; j < BTV; ++j) {
          acc[a][j].x = fmaf(r[a], g[j].x, acc[a][j].x);
          acc[a][j].y = fmaf(r[a], g[j].y, acc[a][j].y);
        }
    }
    if (k < M) {
#pragma unroll
      for (int a = 0; a < IT; ++a)
#pragma unroll
        for (int j = 0; j < BTV; ++j)
          if (i0 + a < I && b0 + j < B) dX[((size_t)(b0 + j) * I + i0 + a) * M + k] = acc[a][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dP / dQ in two phases (both deterministic, fixed-order sums):
//   phase A  G[o,i,k] = dR / Qe,  dR[o,i,k] = sum_b Re(conj(X[b,i,k]) dY[b,o,k])      (thread = mode k, streaming)
//   phase B  dP[o,i,t] = sum_k G mono_t(k),   dQ[o,i,t] = sum_k (-G R) mono_t(k)        (thread = (o,i) pair)
// Phase B keeps the 2T accumulators of a pair in registers, stages [256 pairs x KT] tiles of G and R and the
// monomial tile through shared memory, writes per-k-split partials; dpq_reduce_kernel sums them in fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int GO = 4;   // output channels per CTA of phase A
constexpr int GI = 2;   // input channels per inner step
// SINGLE: B <= BTV, the dY values of the CTA's output channels stay in registers; BTV: batch entries per inner step
// (8 for training batches, fewer for the small batches of the 256^3 configurations)
template <bool SINGLE, int BTV = BT>
__global__ void __launch_bounds__(MIX_THREADS, 4)
rational_g_kernel(const float2* __restrict__ X, const float2* __restrict__ dY, const float* __restrict__ Qe,
                  float* __restrict__ G, int B, int I, int O, long long M) {
  const long long k = (long long)blockIdx.x * MIX_THREADS + threadIdx.x;
  const long long kc = k < M ? k : M - 1;
  const int o0 = blockIdx.y * GO;
  float2 y[GO][BTV];
  if (SINGLE) {
#pragma unroll
    for (int a = 0; a < GO; ++a)
#pragma unroll
      for (int j = 0; j < BTV; ++j)
        y[a][j] = (j < B) ? __ldg(dY + ((size_t)j * O + min(o0 + a, O - 1)) * M + kc) : make_float2(0.f, 0.f);
  }
  const size_t iM = (size_t)I * M;
  const float2* xk = X + kc;                            // + (b*I + i)*M
  const size_t orow = (size_t)o0 * iM + (size_t)k;      // G / Qe index of (o0, i=0, k)
  // Qe is loaded unconditionally (clamped address) at the top of each step: behind the k < M guard of the store block
  // the compiler could not hoist it, and every step stalled a full memory latency on the division's operand
  // (0.68 ms -> 0.26 ms at the 128^3 / width-64 shape).
  const size_t qrow = (size_t)o0 * iM + (size_t)kc;
#pragma unroll 2
  for (int i0 = 0; i0 < I; i0 += GI) {
    float dR[GO][GI], qe[GO][GI];
#pragma unroll
    for (int a = 0; a < GO; ++a)
#pragma unroll
      for (int c = 0; c < GI; ++c) {
        dR[a][c] = 0.f;
        qe[a][c] = __ldg(Qe + qrow + (size_t)min(a, O - 1 - o0) * iM + (size_t)min(i0 + c, I - 1) * M);
      }
    for (int b0 = 0; b0 < B; b0 += BTV) {
      if (!SINGLE) {
#pragma unroll
        for (int a = 0; a < GO; ++a)
#pragma unroll
          for (int j = 0; j < BTV; ++j)
            y[a][j] = (b0 + j < B) ? __ldg(dY + ((size_t)(b0 + j) * O + min(o0 + a, O - 1)) * M + kc) : make_float2(0.f, 0.f);
      }
      float2 xv[GI][BTV];
#pragma unroll
      for (int c = 0; c < GI; ++c) {
        const float2* xi = xk + (size_t)min(i0 + c, I - 1) * M + (size_t)b0 * iM;
#pragma unroll
        for (int j = 0; j < BTV; ++j)   // entries beyond B multiply a zero dY: clamp the address, skip the predicate
          xv[c][j] = __ldg(xi + (size_t)min(j, B - 1 - b0) * iM);
      }
#pragma unroll
      for (int a = 0; a < GO; ++a)
#pragma unroll
        for (int c = 0; c < GI; ++c)
#pragma unroll
          for (int j = 0; j < BTV; ++j) {
            dR[a][c] = fmaf(xv[c][j].x, y[a][j].x, dR[a][c]);
            dR[a][c] = fmaf(xv[c][j].y, y[a][j].y, dR[a][c]);
          }
      if (SINGLE) break;
    }
    if (k < M) {
#pragma unroll
      for (int a = 0; a < GO; ++a)
#pragma unroll
        for (int c = 0; c < GI; ++c)
          if (o0 + a < O && i0 + c < I) {
            const size_t ri = orow + (size_t)a * iM + (size_t)(i0 + c) * M;
            G[ri] = __fdiv_rn(dR[a][c], qe[a][c]);
          }
    }
  }
}

constexpr int DPQ_THREADS = 256;
constexpr int KT = 16;   // modes per staged tile (staging index math assumes a power of two)
constexpr int KTP = KT + 1;

template <int OP, int OQ>
__global__ void __launch_bounds__(DPQ_THREADS, 2)
rational_dpq_kernel(const float* __restrict__ G, const float* __restrict__ R, const float* __restrict__ monoP,
                    const float* __restrict__ monoQ, float* __restrict__ partial, long long npairs, long long M,
                    long long kchunk) {
  constexpr int TPN = Tri<OP>::T, TQN = Tri<OQ>::T;
  constexpr int TP4 = Tri<OP>::TP4, TQ4 = Tri<OQ>::TP4;
  constexpr bool SAME = (OP == OQ);
  __shared__ __align__(16) float Gs[DPQ_THREADS * KTP];
  __shared__ __align__(16) float Rs[DPQ_THREADS * KTP];
  __shared__ __align__(16) float mPs[KT * TP4];
  __shared__ __align__(16) float mQs[SAME ? 4 : KT * TQ4];

  const long long q0 = (long long)blockIdx.x * DPQ_THREADS;
  const long long qraw = q0 + threadIdx.x;
  const bool valid = qraw < npairs;
  float ap[TPN], aq[TQN];
#pragma unroll
  for (int t = 0; t < TPN; ++t) ap[t] = 0.f;
#pragma unroll
  for (int t = 0; t < TQN; ++t) aq[t] = 0.f;

  const long long k_lo = (long long)blockIdx.y * kchunk;
  const long long k_hi = min(M, k_lo + kchunk);
  for (long long k0 = k_lo; k0 < k_hi; k0 += KT) {
    __syncthreads();
    {
      float gv[KT], rv[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) {                 // DPQ_THREADS*KT elements, KT per thread, all loads first
        const int idx = threadIdx.x + j * DPQ_THREADS;
        const int kk = idx & (KT - 1), pr = idx / KT;
        const long long qq = min(q0 + pr, npairs - 1);
        const long long k = k0 + kk;
        const bool ok = k < k_hi;
        gv[j] = ok ? __ldg(G + (size_t)qq * M + k) : 0.f;   // 0 beyond the chunk: contributes nothing
        rv[j] = ok ? __ldg(R + (size_t)qq * M + k) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const int idx = threadIdx.x + j * DPQ_THREADS;
        const int kk = idx & (KT - 1), pr = idx / KT;
        Gs[pr * KTP + kk] = gv[j];
        Rs[pr * KTP + kk] = rv[j];
      }
      for (int idx = threadIdx.x; idx < KT * (TP4 + (SAME ? 0 : TQ4)); idx += DPQ_THREADS) {
        const bool isq = idx >= KT * TP4;
        const int j = isq ? idx - KT * TP4 : idx;
        const int T4 = isq ? TQ4 : TP4, TN = isq ? TQN : TPN;
        const int kk = j / T4, t = j % T4;
        const long long k = min(k0 + kk, M - 1);
        const float v = (t < TN) ? __ldg((isq ? monoQ : monoP) + (size_t)t * M + k) : 0.f;
        (isq ? mQs : mPs)[j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const float g = Gs[threadIdx.x * KTP + kk];
      const float h = -g * Rs[threadIdx.x * KTP + kk];
      const float4* p4 = reinterpret_cast<const float4*>(mPs + kk * TP4);
#pragma unroll
      for (int t4 = 0; t4 < TP4 / 4; ++t4) {
        const float4 m = p4[t4];
        const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (4 * t4 + e < TPN) ap[4 * t4 + e] = fmaf(g, mm[e], ap[4 * t4 + e]);
          if (SAME && 4 * t4 + e < TQN) aq[4 * t4 + e] = fmaf(h, mm[e], aq[4 * t4 + e]);
        }
      }
      if (!SAME) {
        const float4* q4 = reinterpret_cast<const float4*>(mQs + kk * TQ4);
#pragma unroll
        for (int t4 = 0; t4 < TQ4 / 4; ++t4) {
          const float4 m = q4[t4];
          const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (4 * t4 + e < TQN) aq[4 * t4 + e] = fmaf(h, mm[e], aq[4 * t4 + e]);
        }
      }
    }
  }
  if (valid) {
    float* dst = partial + ((size_t)blockIdx.y * npairs + qraw) * (TPN + TQN);
#pragma unroll
    for (int t = 0; t < TPN; ++t) dst[t] = ap[t];
#pragma unroll
    for (int t = 0; t < TQN; ++t) dst[TPN + t] = aq[t];
  }
}

// ---------------------------------------------------------------------------------------------
// Phase B, pipelined: the same sums with the G / R tiles ([256 pairs][32 modes], 128-byte row pieces) and the monomial tile
// staged by cp.async two tiles deep, so the loads of tile n+1 run under the 1280 FMAs per thread of tile n.  The version above
// loads a tile into registers, synchronises, computes, synchronises: 0.33 ms for 0.57 GB at the 128^3 / width-64 shape and
// 2.5 ms for 4.4 GB at 64^3 modes (1.7 TB/s) against ~0.1 / 0.75 ms of HBM or FMA time.  Needs 16-byte aligned rows
// (M % 4 == 0) and a k-major monomial table (monoT [M][T4], transposed once per call by mono_transpose_kernel).
// Rows are padded to 36 floats: thread = pair reads its row as float4, and 16-byte chunk index 9 p mod 8 = p mod 8 is
// conflict-free per quarter warp.
// ---------------------------------------------------------------------------------------------
constexpr int KT2 = 32;
constexpr int ROWF = KT2 + 4;
__device__ __forceinline__ void dpq_cp16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__global__ void mono_transpose_kernel(const float* __restrict__ mono, float* __restrict__ monoT, int TN, int T4, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over M * T4
  if (i >= M * T4) return;
  const long long k = i / T4;
  const int t = (int)(i - k * T4);
  monoT[i] = t < TN ? __ldg(mono + (size_t)t * M + k) : 0.f;
}
template <int OP, int OQ>
struct DpqSmem {
  static constexpr int TP4 = Tri<OP>::TP4, TQ4 = Tri<OQ>::TP4;
  static constexpr bool SAME = (OP == OQ);
  static constexpr int G = 0, R = DPQ_THREADS * ROWF * 4, MP = 2 * DPQ_THREADS * ROWF * 4, MQ = MP + KT2 * TP4 * 4;
  static constexpr int STAGE = MQ + (SAME ? 0 : KT2 * TQ4 * 4);
  static constexpr int TOTAL = 2 * STAGE;
};

template <int OP, int OQ>
__global__ void __launch_bounds__(DPQ_THREADS, 1)
rational_dpq_pipe_kernel(const float* __restrict__ G, const float* __restrict__ R, const float* __restrict__ monoPT,
                         const float* __restrict__ monoQT, float* __restrict__ partial, long long npairs, long long M,
                         long long kchunk) {
  constexpr int TPN = Tri<OP>::T, TQN = Tri<OQ>::T;
  constexpr int TP4 = Tri<OP>::TP4, TQ4 = Tri<OQ>::TP4;
  constexpr bool SAME = (OP == OQ);
  using L = DpqSmem<OP, OQ>;
  extern __shared__ __align__(16) uint8_t dsm[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dsm);
  const long long q0 = (long long)blockIdx.x * DPQ_THREADS;
  const long long qraw = q0 + threadIdx.x;
  const long long k_lo = (long long)blockIdx.y * kchunk;
  const long long k_hi = min(M, k_lo + kchunk);
  const int ntiles = k_hi > k_lo ? (int)((k_hi - k_lo + KT2 - 1) / KT2) : 0;

  auto issue = [&](int tile) {
    const long long k0 = k_lo + (long long)tile * KT2;
    const uint32_t st = sbase + (uint32_t)(tile & 1) * L::STAGE;
#pragma unroll
    for (int j = 0; j < 8; ++j) {                      // 256 rows x 8 chunks per tensor, 8 per thread: a warp = 4 rows x 128 B
      const int c = threadIdx.x + j * DPQ_THREADS, row = c >> 3, ch = c & 7;
      const long long qq = min(q0 + row, npairs - 1);
      const long long k = k0 + ch * 4;
      const int nb = k >= k_hi ? 0 : (int)min(4LL, k_hi - k) * 4;      // bytes of this chunk inside the k range (rest zero-filled)
      const size_t off = (size_t)qq * M + (nb ? k : k_lo);
      dpq_cp16(st + L::G + (uint32_t)(row * ROWF + ch * 4) * 4, G + off, nb);
      dpq_cp16(st + L::R + (uint32_t)(row * ROWF + ch * 4) * 4, R + off, nb);
    }
    for (int c = threadIdx.x; c < KT2 * TP4 / 4; c += DPQ_THREADS) {       // monoT rows k0 .. k0+31: contiguous
      const long long k = k0 + (c * 4) / TP4;
      const int nb = k < M ? 16 : 0;
      dpq_cp16(st + L::MP + c * 16, monoPT + (nb ? (size_t)k0 * TP4 + c * 4 : 0), nb);
    }
    if (!SAME) {
      for (int c = threadIdx.x; c < KT2 * TQ4 / 4; c += DPQ_THREADS) {
        const long long k = k0 + (c * 4) / TQ4;
        const int nb = k < M ? 16 : 0;
        dpq_cp16(st + L::MQ + c * 16, monoQT + (nb ? (size_t)k0 * TQ4 + c * 4 : 0), nb);
      }
    }
  };

  float ap[TPN], aq[TQN];
#pragma unroll
  for (int t = 0; t < TPN; ++t) ap[t] = 0.f;
#pragma unroll
  for (int t = 0; t < TQN; ++t) aq[t] = 0.f;
  if (ntiles > 0) issue(0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int tile = 0; tile < ntiles; ++tile) {
    if (tile + 1 < ntiles) issue(tile + 1);            // its buffer was released by the barrier that ended tile - 1
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // tile `tile` has landed (this thread's copies) ...
    __syncthreads();                                   // ... and everybody else's
    const uint8_t* st = dsm + (tile & 1) * L::STAGE;
    const float* Gs = reinterpret_cast<const float*>(st + L::G) + threadIdx.x * ROWF;
    const float* Rs = reinterpret_cast<const float*>(st + L::R) + threadIdx.x * ROWF;
    const float* mPs = reinterpret_cast<const float*>(st + L::MP);
    const float* mQs = reinterpret_cast<const float*>(st + L::MQ);
#pragma unroll 2
    for (int k4 = 0; k4 < KT2 / 4; ++k4) {
      const float4 g4 = *reinterpret_cast<const float4*>(Gs + 4 * k4);
      const float4 r4 = *reinterpret_cast<const float4*>(Rs + 4 * k4);
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kk = 4 * k4 + e;
        const float g = gg[e];
        const float h = -g * rr[e];
        const float4* p4 = reinterpret_cast<const float4*>(mPs + kk * TP4);
#pragma unroll
        for (int t4 = 0; t4 < TP4 / 4; ++t4) {
          const float4 m = p4[t4];
          const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            if (4 * t4 + x < TPN) ap[4 * t4 + x] = fmaf(g, mm[x], ap[4 * t4 + x]);
            if (SAME && 4 * t4 + x < TQN) aq[4 * t4 + x] = fmaf(h, mm[x], aq[4 * t4 + x]);
          }
        }
        if (!SAME) {
          const float4* q4 = reinterpret_cast<const float4*>(mQs + kk * TQ4);
#pragma unroll
          for (int t4 = 0; t4 < TQ4 / 4; ++t4) {
            const float4 m = q4[t4];
            const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
              if (4 * t4 + x < TQN) aq[4 * t4 + x] = fmaf(h, mm[x], aq[4 * t4 + x]);
          }
        }
      }
    }
    __syncthreads();                                   // this buffer may be refilled (by issue(tile + 2) next iteration)
  }
  if (qraw < npairs) {
    float* dst = partial + ((size_t)blockIdx.y * npairs + qraw) * (TPN + TQN);
#pragma unroll
    for (int t = 0; t < TPN; ++t) dst[t] = ap[t];
#pragma unroll
    for (int t = 0; t < TQN; ++t) dst[TPN + t] = aq[t];
  }
}

// one thread per (pair, monomial): coalesced over the 2T partials of a pair, fixed-order sum over the k splits;
// dP / dQ are zero-filled beforehand so the unused (a+b+c >= order) entries read 0 like autograd's
__global__ void dpq_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dP, float* __restrict__ dQ,
                                  int KS, long long npairs, int op, int oq, int TPN, int TQN, MonoOffsets offs) {
  const int TT = TPN + TQN;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npairs * TT) return;
  const long long q = idx / TT;
  const int t = (int)(idx - q * TT);
  double s = 0.0;
  for (int ks = 0; ks < KS; ++ks) s += (double)__ldg(partial + (size_t)ks * npairs * TT + idx);
  if (t < TPN) dP[(size_t)q * (op * op * op) + offs.p[t]] = (float)s;
  else dQ[(size_t)q * (oq * oq * oq) + offs.q[t - TPN]] = (float)s;
}

// ---------------------------------------------------------------------------------------------
// complex mix (SpectralConv3D).  W [O][I][mxy][mz] complex, only kz < mzh used.
//   fwd  : out[b][u][k] = sum_v W(u,v)[k] in[b][v][k]            (TR=false, CONJ=false)
//   dX   : out[b][u][k] = sum_v conj(W(v,u)[k]) in[b][v][k]      (TR=true,  CONJ=true)
// ---------------------------------------------------------------------------------------------
template <bool TR>
__global__ void __launch_bounds__(MIX_THREADS)
complex_mix_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ W, int B, int U,
                   int V, long long mxy, int mzh, int mz) {
  const long long M = mxy * mzh;
  const long long k = (long long)blockIdx.x * MIX_THREADS + threadIdx.x;
  const long long kc = k < M ? k : M - 1;
  const long long kxy = kc / mzh;
  const int kz = (int)(kc - kxy * mzh);
  const size_t wk = (size_t)kxy * mz + kz;
  const size_t Wm = (size_t)mxy * mz;
  const int u0 = blockIdx.y * UT;
  const int I = TR ? U : V;
  for (int b0 = 0; b0 < B; b0 += BT) {
    float2 acc[UT][BT];
#pragma unroll
    for (int ul = 0; ul < UT; ++ul)
#pragma unroll
      for (int j = 0; j < BT; ++j) acc[ul][j] = make_float2(0.f, 0.f);
    for (int v = 0; v < V; ++v) {
      float2 xin[BT];
#pragma unroll
      for (int j = 0; j < BT; ++j)
        xin[j] = (b0 + j < B) ? __ldg(in + ((size_t)(b0 + j) * V + v) * M + kc) : make_float2(0.f, 0.f);
#pragma unroll
      for (int ul = 0; ul < UT; ++ul) {
        const int u = min(u0 + ul, U - 1);
        const int o = TR ? v : u, i = TR ? u : v;
        float2 w = __ldg(W + ((size_t)o * I + i) * Wm + wk);
        if (TR) w.y = -w.y;
#pragma unroll
        for (int j = 0; j < BT; ++j) cfma(acc[ul][j], w, xin[j]);
      }
    }
    if (k < M) {
#pragma unroll
      for (int ul = 0; ul < UT; ++ul)
#pragma unroll
        for (int j = 0; j < BT; ++j)
          if (u0 + ul < U && b0 + j < B) out[((size_t)(b0 + j) * U + u0 + ul) * M + k] = acc[ul][j];
    }
  }
}

// dW[o][i][kxy][kz] = sum_b conj(X[b,i,k]) dY[b,o,k] for kz < mzh, 0 for kz >= mzh.
constexpr int WT = 4;
__global__ void __launch_bounds__(MIX_THREADS)
complex_dw_kernel(const float2* __restrict__ X, const float2* __restrict__ dY, float2* __restrict__ dW, int B, int I,
                  int O, long long mxy, int mzh, int mz) {
  const long long Mw = mxy * mz, M = mxy * mzh;
  const long long kw = (long long)blockIdx.x * MIX_THREADS + threadIdx.x;
  if (kw >= Mw) return;
  const long long kxy = kw / mz;
  const int kz = (int)(kw - kxy * mz);
  const int o0 = blockIdx.y * WT, i0 = blockIdx.z * WT;
  float2 acc[WT][WT];
#pragma unroll
  for (int a = 0; a < WT; ++a)
#pragma unroll
    for (int c = 0; c < WT; ++c) acc[a][c] = make_float2(0.f, 0.f);
  if (kz < mzh) {
    const long long k = kxy * mzh + kz;
    for (int b = 0; b < B; ++b) {
      float2 xv[WT], yv[WT];
#pragma unroll
      for (int c = 0; c < WT; ++c) {
        xv[c] = __ldg(X + ((size_t)b * I + min(i0 + c, I - 1)) * M + k);
        xv[c].y = -xv[c].y;
      }
#pragma unroll
      for (int a = 0; a < WT; ++a) yv[a] = __ldg(dY + ((size_t)b * O + min(o0 + a, O - 1)) * M + k);
#pragma unroll
      for (int a = 0; a < WT; ++a)
#pragma unroll
        for (int c = 0; c < WT; ++c) cfma(acc[a][c], xv[c], yv[a]);
    }
  }
#pragma unroll
  for (int a = 0; a < WT; ++a)
#pragma unroll
    for (int c = 0; c < WT; ++c)
      if (o0 + a < O && i0 + c < I) dW[((size_t)(o0 + a) * I + i0 + c) * Mw + kw] = acc[a][c];
}

// ------------------------------- host-side dispatch -------------------------------------------
template <int OP, int OQ>
static int launch_mix(const float2* in, float2* out, const float* P, const float* Q, const float* monoP,
                      const float* monoQ, float eps, float* Rout, float* Qeout, int B, int U, int V, long long M,
                      const MonoOffsets& offs, cudaStream_t st) {
  const size_t smem = (size_t)UT * V * (Tri<OP>::TP4 + Tri<OQ>::TP4) * sizeof(float);
  auto k = B >= 5 ? rational_mix_kernel<OP, OQ, false, 8>
                  : (B >= 3 ? rational_mix_kernel<OP, OQ, false, 4> : (B == 2 ? rational_mix_kernel<OP, OQ, false, 2>
                                                                              : rational_mix_kernel<OP, OQ, false, 1>));
  if (smem > 200 * 1024) return fail(PDEPHI_ERR_UNSUPPORTED, "rational_mix_fwd: %d input channels exceed shared memory", V);
  if (smem > 48 * 1024) PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((M + MIX_THREADS - 1) / MIX_THREADS), (U + UT - 1) / UT);
  {
    ProfScope ps(K_RATIONAL_MIX, st);
    // 1 and -0 travel as kernel arguments: the packed evaluation needs them opaque to the compiler (see eval_R2)
    k<<<grid, MIX_THREADS, smem, st>>>(in, out, P, Q, monoP, monoQ, eps, Rout, Qeout, B, U, V, M, offs, 1.0f, -0.0f);
  }
  PDEPHI_LAUNCH_CHECK("rational_mix_kernel");
  return PDEPHI_OK;
}

static void dpq_geometry(int I, int O, long long M, int* pair_blocks, int* KS, long long* kchunk) {
  const long long npairs = (long long)O * I;
  *pair_blocks = (int)((npairs + DPQ_THREADS - 1) / DPQ_THREADS);
  long long ks = (6 * 148 + *pair_blocks - 1) / *pair_blocks;
  const long long max_ks = (M + 63) / 64;
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  long long chunk = (M + ks - 1) / ks;
  chunk = (chunk + KT - 1) / KT * KT;
  *KS = (int)((M + chunk - 1) / chunk);
  *kchunk = chunk;
}
// pipelined kernel: one CTA per SM (152 KB of shared memory), so one wave of pair_blocks x KS CTAs; chunks are multiples of 32
static bool dpq_pipe_ok(long long M) { return M % 4 == 0 && M >= 4 * KT2; }
static void dpq_pipe_geometry(int I, int O, long long M, int* pair_blocks, int* KS, long long* kchunk) {
  const long long npairs = (long long)O * I;
  *pair_blocks = (int)((npairs + DPQ_THREADS - 1) / DPQ_THREADS);
  long long ks = 148 / *pair_blocks;
  if (ks < 1) ks = 1;
  const long long max_ks = (M + 4 * KT2 - 1) / (4 * KT2);      // at least four tiles per CTA
  if (ks > max_ks) ks = max_ks;
  long long chunk = (M + ks - 1) / ks;
  chunk = (chunk + KT2 - 1) / KT2 * KT2;
  *KS = (int)((M + chunk - 1) / chunk);
  *kchunk = chunk;
}
static size_t dpq_partial_bytes(int op, int oq, int I, int O, long long M) {
  int pair_blocks, KS;
  long long kchunk;
  dpq_geometry(I, O, M, &pair_blocks, &KS, &kchunk);
  return align_up((size_t)KS * O * I * (tri(op) + tri(oq)) * sizeof(float), 256);
}

template <int OP, int OQ>
static int launch_dpq(const float2* X, const float2* dY, const float* R, const float* Qe, const float* monoP,
                      const float* monoQ, float* dP, float* dQ, int B, int I, int O, long long M, float* partial,
                      float* G, float* monoT, const MonoOffsets& offs, cudaStream_t st) {
  int pair_blocks, KS;
  long long kchunk;
  dpq_geometry(I, O, M, &pair_blocks, &KS, &kchunk);
  constexpr int TPN = Tri<OP>::T, TQN = Tri<OQ>::T;
  constexpr int TP4 = Tri<OP>::TP4, TQ4 = Tri<OQ>::TP4;
  const long long npairs = (long long)O * I;
  if ((O + GO - 1) / GO > 65535) return fail(PDEPHI_ERR_UNSUPPORTED, "rational_mix_bwd: too many output channels");
  {
    dim3 grid((unsigned)((M + MIX_THREADS - 1) / MIX_THREADS), (O + GO - 1) / GO);
    ProfScope ps(K_RATIONAL_G, st);
    if (B == 1) rational_g_kernel<true, 1><<<grid, MIX_THREADS, 0, st>>>(X, dY, Qe, G, B, I, O, M);
    else if (B == 2) rational_g_kernel<true, 2><<<grid, MIX_THREADS, 0, st>>>(X, dY, Qe, G, B, I, O, M);
    else if (B <= 4) rational_g_kernel<true, 4><<<grid, MIX_THREADS, 0, st>>>(X, dY, Qe, G, B, I, O, M);
    else if (B <= BT) rational_g_kernel<true><<<grid, MIX_THREADS, 0, st>>>(X, dY, Qe, G, B, I, O, M);
    else rational_g_kernel<false><<<grid, MIX_THREADS, 0, st>>>(X, dY, Qe, G, B, I, O, M);
  }
  PDEPHI_LAUNCH_CHECK("rational_g_kernel");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (dpq_pipe_ok(M) && al16(G) && al16(R) && monoT && al16(monoT)) {
    // pipelined reduction: k-major monomial tables first (a few MB), then one wave of CTAs with cp.async double buffering
    dpq_pipe_geometry(I, O, M, &pair_blocks, &KS, &kchunk);
    float* monoPT = monoT;
    float* monoQT = monoT + (size_t)M * TP4;
    {
      ProfScope ps(K_RATIONAL_DPQ, st);
      mono_transpose_kernel<<<(unsigned)((M * TP4 + 255) / 256), 256, 0, st>>>(monoP, monoPT, TPN, TP4, M);
      if (OP != OQ) mono_transpose_kernel<<<(unsigned)((M * TQ4 + 255) / 256), 256, 0, st>>>(monoQ, monoQT, TQN, TQ4, M);
      auto k = rational_dpq_pipe_kernel<OP, OQ>;
      PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, DpqSmem<OP, OQ>::TOTAL));
      k<<<dim3(pair_blocks, KS), DPQ_THREADS, DpqSmem<OP, OQ>::TOTAL, st>>>(G, R, monoPT, monoQT, partial, npairs, M, kchunk);
    }
    PDEPHI_LAUNCH_CHECK("rational_dpq_pipe_kernel");
  } else {
    ProfScope ps(K_RATIONAL_DPQ, st);
    rational_dpq_kernel<OP, OQ><<<dim3(pair_blocks, KS), DPQ_THREADS, 0, st>>>(G, R, monoP, monoQ, partial, npairs, M, kchunk);
    PDEPHI_LAUNCH_CHECK("rational_dpq_kernel");
  }
  PDEPHI_CUDA(cudaMemsetAsync(dP, 0, (size_t)npairs * OP * OP * OP * sizeof(float), st));
  PDEPHI_CUDA(cudaMemsetAsync(dQ, 0, (size_t)npairs * OQ * OQ * OQ * sizeof(float), st));
  {
    ProfScope ps(K_DPQ_REDUCE, st);
    const long long n = npairs * (TPN + TQN);
    dpq_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, dP, dQ, KS, npairs, OP, OQ, TPN, TQN, offs);
  }
  PDEPHI_LAUNCH_CHECK("dpq_reduce_kernel");
  return PDEPHI_OK;
}

}  // namespace pdephi

using namespace pdephi;

// every (p, q) in 2..4 -- the orders the reference's models construct (rational_fourier.py:33, rfno.py:66,
// multiscale_fno.py:100-127) -- plus the symmetric orders 1, 5 and 6 (quantum_rational_fourier.py:392 asks for (6,6))
#define ORDER_CASES CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 2) CASE(3, 3) CASE(3, 4) CASE(4, 2) CASE(4, 3) CASE(4, 4) \
  CASE(1, 1) CASE(5, 5) CASE(6, 6)

extern "C" int pdephi_rational_mix_fwd(const float* X, const float* P, const float* Q, const float* monoP,
                                       const float* monoQ, int order_p, int order_q, float eps, const float* mask,
                                       float proj_eps, float* Y, float* stats, float* R_out, float* Qe_out, int B,
                                       int I, int O, long long M, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(X && P && Q && monoP && monoQ && mask && Y && stats, "rational_mix_fwd: null pointer");
  PDEPHI_REQUIRE((R_out == nullptr) == (Qe_out == nullptr), "rational_mix_fwd: R_out and Qe_out must both be given or both be NULL");
  PDEPHI_REQUIRE(B > 0 && I > 0 && O > 0 && M > 0, "rational_mix_fwd: non-positive size");
  PDEPHI_REQUIRE((long long)B * O <= 0x7fffffffLL, "rational_mix_fwd: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  MonoOffsets offs;
  memset(&offs, 0, sizeof(offs));
  fill_offsets(order_p, offs.p);
  fill_offsets(order_q, offs.q);
  int rc = -1;
#define CASE(a, b)                                                                                                   \
  if (order_p == a && order_q == b)                                                                                  \
    rc = launch_mix<a, b>((const float2*)X, (float2*)Y, P, Q, monoP, monoQ, eps, R_out, Qe_out, B, O, I, M, offs, st);
  ORDER_CASES
#undef CASE
  if (rc == -1) return fail(PDEPHI_ERR_UNSUPPORTED, "rational order (%d,%d) is not instantiated (supported: every (p,q) in 2..4, and (1,1), (5,5), (6,6))", order_p, order_q);
  if (rc) return rc;
  {
    ProfScope ps(K_PROJ_STATS, st);
    projection_stats_kernel<<<(unsigned)(B * O), ROW_THREADS, 0, st>>>((const float2*)Y, mask, proj_eps, stats, M);
  }
  PDEPHI_LAUNCH_CHECK("projection_stats_kernel");
  return PDEPHI_OK;
}

// ---- transfer function alone / mix from a materialised transfer function (the cached forward) ----
template <int OP, int OQ>
static int launch_transfer(const float* P, const float* Q, const float* monoP, const float* monoQ, float eps, float* R, float* Qe,
                           int U, int V, long long M, const MonoOffsets& offs, const int* changed, cudaStream_t st) {
  const size_t smem = (size_t)UT * V * (Tri<OP>::TP4 + Tri<OQ>::TP4) * sizeof(float);
  auto k = rational_transfer_kernel<OP, OQ>;
  if (smem > 200 * 1024) return fail(PDEPHI_ERR_UNSUPPORTED, "rational_transfer: %d input channels exceed shared memory", V);
  if (smem > 48 * 1024) PDEPHI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((M + MIX_THREADS - 1) / MIX_THREADS), (U + UT - 1) / UT);
  {
    ProfScope ps(K_RATIONAL_TRANSFER, st);
    k<<<grid, MIX_THREADS, smem, st>>>(P, Q, monoP, monoQ, eps, R, Qe, U, V, M, offs, 1.0f, -0.0f, changed);
  }
  PDEPHI_LAUNCH_CHECK("rational_transfer_kernel");
  return PDEPHI_OK;
}

extern "C" int pdephi_rational_transfer(const float* P, const float* Q, const float* monoP, const float* monoQ, int order_p,
                                        int order_q, float eps, float* R_out, float* Qe_out, int I, int O, long long M,
                                        float* snapP, float* snapQ, int* changed, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(P && Q && monoP && monoQ && R_out && Qe_out, "rational_transfer: null pointer");
  PDEPHI_REQUIRE(I > 0 && O > 0 && M > 0, "rational_transfer: non-positive size");
  PDEPHI_REQUIRE((snapP == nullptr) == (snapQ == nullptr) && (snapP == nullptr) == (changed == nullptr),
                 "rational_transfer: snapP, snapQ and changed must all be given or all be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (changed) {     // refresh: re-evaluate only if a coefficient differs from its snapshot
    PDEPHI_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
    const long long nP = (long long)O * I * order_p * order_p * order_p, nQ = (long long)O * I * order_q * order_q * order_q;
    {
      ProfScope ps(K_RATIONAL_TRANSFER, st);
      params_changed_kernel<<<(unsigned)((nP + nQ + 1023) / 1024 < 296 ? (nP + nQ + 1023) / 1024 : 296), 256, 0, st>>>(
          (const uint32_t*)P, (uint32_t*)snapP, nP, (const uint32_t*)Q, (uint32_t*)snapQ, nQ, changed);
    }
    PDEPHI_LAUNCH_CHECK("params_changed_kernel");
  }
  MonoOffsets offs;
  memset(&offs, 0, sizeof(offs));
  fill_offsets(order_p, offs.p);
  fill_offsets(order_q, offs.q);
#define CASE(a, b) \
  if (order_p == a && order_q == b) return launch_transfer<a, b>(P, Q, monoP, monoQ, eps, R_out, Qe_out, O, I, M, offs, changed, st);
  ORDER_CASES
#undef CASE
  return fail(PDEPHI_ERR_UNSUPPORTED, "rational order (%d,%d) is not instantiated (supported: every (p,q) in 2..4, and (1,1), (5,5), (6,6))", order_p, order_q);
}

extern "C" int pdephi_rational_mix_apply(const float* X, const float* R, const float* mask, float proj_eps, float* Y, float* stats,
                                         int B, int I, int O, long long M, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(X && R && mask && Y && stats, "rational_mix_apply: null pointer");
  PDEPHI_REQUIRE(B > 0 && I > 0 && O > 0 && M > 0, "rational_mix_apply: non-positive size");
  PDEPHI_REQUIRE((long long)B * O <= 0x7fffffffLL, "rational_mix_apply: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // mix_from_R_kernel<.., FWD>: its "I" = output channels, its "O" = contracted input channels
    dim3 grid((unsigned)((M + MIX_THREADS - 1) / MIX_THREADS), (O + IT - 1) / IT);
    ProfScope ps(K_RATIONAL_MIX_APPLY, st);
    if (B >= 5) mix_from_R_kernel<8, true><<<grid, MIX_THREADS, 0, st>>>((const float2*)X, (float2*)Y, R, B, O, I, M);
    else if (B >= 3) mix_from_R_kernel<4, true><<<grid, MIX_THREADS, 0, st>>>((const float2*)X, (float2*)Y, R, B, O, I, M);
    else if (B == 2) mix_from_R_kernel<2, true><<<grid, MIX_THREADS, 0, st>>>((const float2*)X, (float2*)Y, R, B, O, I, M);
    else mix_from_R_kernel<1, true><<<grid, MIX_THREADS, 0, st>>>((const float2*)X, (float2*)Y, R, B, O, I, M);
  }
  PDEPHI_LAUNCH_CHECK("mix_from_R_kernel<fwd>");
  {
    ProfScope ps(K_PROJ_STATS, st);
    projection_stats_kernel<<<(unsigned)(B * O), ROW_THREADS, 0, st>>>((const float2*)Y, mask, proj_eps, stats, M);
  }
  PDEPHI_LAUNCH_CHECK("projection_stats_kernel");
  return PDEPHI_OK;
}

extern "C" int pdephi_projection_bwd(const float* dZ, const float* Y, const float* mask, const float* stats, float* dY,
                                     long long rows, long long M, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(dZ && Y && mask && stats && dY, "projection_bwd: null pointer");
  PDEPHI_REQUIRE(rows > 0 && M > 0 && rows <= 0x7fffffffLL, "projection_bwd: bad size");
  {
    ProfScope ps(K_PROJ_BWD, (cudaStream_t)stream);
    projection_bwd_kernel<<<(unsigned)rows, ROW_THREADS, 0, (cudaStream_t)stream>>>((const float2*)dZ, (const float2*)Y,
                                                                                     mask, stats, (float2*)dY, M, nullptr, 0);
  }
  PDEPHI_LAUNCH_CHECK("projection_bwd_kernel");
  return PDEPHI_OK;
}

extern "C" int pdephi_projection_bwd_sharded(const float* dZ, const float* Y, const float* mask, const float* stats, float* G,
                                             float* dY, long long rows, long long M, int phase, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(dZ && Y && mask && stats && G, "projection_bwd_sharded: null pointer");
  PDEPHI_REQUIRE(phase == 1 || (phase == 2 && dY), "projection_bwd_sharded: phase must be 1 or 2 (2 needs dY)");
  PDEPHI_REQUIRE(rows > 0 && M > 0 && rows <= 0x7fffffffLL, "projection_bwd_sharded: bad size");
  {
    ProfScope ps(K_PROJ_BWD, (cudaStream_t)stream);
    projection_bwd_kernel<<<(unsigned)rows, ROW_THREADS, 0, (cudaStream_t)stream>>>((const float2*)dZ, (const float2*)Y,
                                                                                     mask, stats, (float2*)dY, M, G, phase);
  }
  PDEPHI_LAUNCH_CHECK("projection_bwd_kernel");
  return PDEPHI_OK;
}

extern "C" size_t pdephi_rational_mix_bwd_workspace_bytes(int order_p, int order_q, int I, int O, long long M) {
  if (I <= 0 || O <= 0 || M <= 0 || order_p < 1 || order_q < 1) return 0;
  // per-k-split partials of dP/dQ, the G = dR/Qe matrix [O, I, M], the k-major monomial tables of the pipelined reduction
  const size_t t4p = (size_t)(tri(order_p) + 3) / 4 * 4, t4q = (size_t)(tri(order_q) + 3) / 4 * 4;
  return dpq_partial_bytes(order_p, order_q, I, O, M) + align_up((size_t)O * I * M * sizeof(float), 256) +
         align_up((size_t)M * (t4p + t4q) * sizeof(float), 256);
}

extern "C" int pdephi_rational_mix_bwd(const float* X, const float* dY, const float* R, const float* Qe,
                                       const float* monoP, const float* monoQ, int order_p, int order_q, float* dX,
                                       float* dP, float* dQ, int B, int I, int O, long long M, void* workspace,
                                       size_t workspace_bytes, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(X && dY && R && Qe && monoP && monoQ && dX && dP && dQ, "rational_mix_bwd: null pointer");
  PDEPHI_REQUIRE(B > 0 && I > 0 && O > 0 && M > 0, "rational_mix_bwd: non-positive size");
  const size_t need = pdephi_rational_mix_bwd_workspace_bytes(order_p, order_q, I, O, M);
  if (!workspace || workspace_bytes < need)
    return fail(PDEPHI_ERR_WORKSPACE, "rational_mix_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  {
    dim3 grid((unsigned)((M + MIX_THREADS - 1) / MIX_THREADS), (I + IT - 1) / IT);
    ProfScope ps(K_RATIONAL_MIX_T, st);
    if (B >= 5) mix_from_R_kernel<8><<<grid, MIX_THREADS, 0, st>>>((const float2*)dY, (float2*)dX, R, B, I, O, M);
    else if (B >= 3) mix_from_R_kernel<4><<<grid, MIX_THREADS, 0, st>>>((const float2*)dY, (float2*)dX, R, B, I, O, M);
    else if (B == 2) mix_from_R_kernel<2><<<grid, MIX_THREADS, 0, st>>>((const float2*)dY, (float2*)dX, R, B, I, O, M);
    else mix_from_R_kernel<1><<<grid, MIX_THREADS, 0, st>>>((const float2*)dY, (float2*)dX, R, B, I, O, M);
  }
  PDEPHI_LAUNCH_CHECK("mix_from_R_kernel");
  MonoOffsets offs;
  memset(&offs, 0, sizeof(offs));
  fill_offsets(order_p, offs.p);
  fill_offsets(order_q, offs.q);
  float* Gbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + dpq_partial_bytes(order_p, order_q, I, O, M));
  float* monoT = reinterpret_cast<float*>(reinterpret_cast<char*>(Gbuf) + align_up((size_t)O * I * M * sizeof(float), 256));
#define CASE(a, b)                                                                                          \
  if (order_p == a && order_q == b)                                                                         \
    return launch_dpq<a, b>((const float2*)X, (const float2*)dY, R, Qe, monoP, monoQ, dP, dQ, B, I, O, M,    \
                            (float*)workspace, Gbuf, monoT, offs, st);
  ORDER_CASES
#undef CASE
  return fail(PDEPHI_ERR_UNSUPPORTED, "rational order (%d,%d) is not instantiated (supported: every (p,q) in 2..4, and (1,1), (5,5), (6,6))", order_p, order_q);
}

extern "C" int pdephi_complex_mix_fwd(const float* X, const float* W, float* Y, int B, int I, int O, long long mxy,
                                      int mzh, int mz, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(X && W && Y, "complex_mix_fwd: null pointer");
  PDEPHI_REQUIRE(B > 0 && I > 0 && O > 0 && mxy > 0 && mzh > 0 && mz >= mzh, "complex_mix_fwd: bad size");
  const long long M = mxy * mzh;
  dim3 grid((unsigned)((M + MIX_THREADS - 1) / MIX_THREADS), (O + UT - 1) / UT);
  {
    ProfScope ps(K_COMPLEX_MIX, (cudaStream_t)stream);
    complex_mix_kernel<false><<<grid, MIX_THREADS, 0, (cudaStream_t)stream>>>((const float2*)X, (float2*)Y,
                                                                             (const float2*)W, B, O, I, mxy, mzh, mz);
  }
  PDEPHI_LAUNCH_CHECK("complex_mix_kernel");
  return PDEPHI_OK;
}

extern "C" int pdephi_complex_mix_bwd(const float* X, const float* W, const float* dY, float* dX, float* dW, int B,
                                      int I, int O, long long mxy, int mzh, int mz, pdephi_stream_t stream) {
  PDEPHI_REQUIRE(X && W && dY && dX && dW, "complex_mix_bwd: null pointer");
  PDEPHI_REQUIRE(B > 0 && I > 0 && O > 0 && mxy > 0 && mzh > 0 && mz >= mzh, "complex_mix_bwd: bad size");
  PDEPHI_REQUIRE((O + WT - 1) / WT <= 65535 && (I + WT - 1) / WT <= 65535, "complex_mix_bwd: too many channels");
  const long long M = mxy * mzh, Mw = mxy * mz;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((M + MIX_THREADS - 1) / MIX_THREADS), (I + UT - 1) / UT);
  {
    ProfScope ps(K_COMPLEX_MIX_T, st);
    complex_mix_kernel<true><<<grid, MIX_THREADS, 0, st>>>((const float2*)dY, (float2*)dX, (const float2*)W, B, I, O, mxy,
                                                          mzh, mz);
  }
  PDEPHI_LAUNCH_CHECK("complex_mix_kernel(dX)");
  dim3 gw((unsigned)((Mw + MIX_THREADS - 1) / MIX_THREADS), (O + WT - 1) / WT, (I + WT - 1) / WT);
  {
    ProfScope ps(K_COMPLEX_DW, st);
    complex_dw_kernel<<<gw, MIX_THREADS, 0, st>>>((const float2*)X, (const float2*)dY, (float2*)dW, B, I, O, mxy, mzh, mz);
  }
  PDEPHI_LAUNCH_CHECK("complex_dw_kernel");
  return PDEPHI_OK;
}
