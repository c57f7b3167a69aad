// GPU probe: achievable HBM bandwidth of the block-tail access pattern -- 2 reads + 2 writes of [B*64 rows][S] fp32
// tensors walked as tiles of 64 channel rows x TP points (rows 8 MB apart), against a flat streaming copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bw_probe bw_probe.cu && ./bw_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int TP, int NR, int NW>   // points per tile, tensors read, tensors written
__global__ void __launch_bounds__(512, 1) tile_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c,
                                                      float* __restrict__ d, long long S, long long ntiles, long long tiles_per_batch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int V = TP / 128;   // float4 per lane per row segment
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long bb = t / tiles_per_batch, s0 = (t - bb * tiles_per_batch) * TP;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long row = bb * 64 + warp * 4 + r;
      const size_t off = (size_t)row * S + s0;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const size_t o = off + (size_t)(v * 32 + lane) * 4;
        float4 x = __ldcs(reinterpret_cast<const float4*>(a + o));
        if (NR > 1) { float4 y = __ldcs(reinterpret_cast<const float4*>(b + o)); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
        __stcs(reinterpret_cast<float4*>(c + o), x);
        if (NW > 1) __stcs(reinterpret_cast<float4*>(d + o), x);
      }
    }
  }
}
// read-only variants: SEG = contiguous bytes a warp instruction takes from one row (512 = one row per instruction,
// 128 = four rows x 128 B, the request granularity of a [64 rows][32 floats] TMA box); the sum keeps the loads alive
template <int SEG, int DEPTH>
__global__ void __launch_bounds__(512, 1) read_kernel(const float* __restrict__ a, float* __restrict__ sink, long long S, long long ntiles,
                                                      long long tiles_per_batch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long bb = t / tiles_per_batch, s0 = (t - bb * tiles_per_batch) * 128;
    float4 x[DEPTH];
#pragma unroll
    for (int r = 0; r < DEPTH; ++r) {     // 16 warps x DEPTH(4) instructions x 512 B = 32 KB = one [64 rows][128 points] tile
      size_t o;
      if (SEG == 512) o = (size_t)(bb * 64 + warp * 4 + r) * S + s0 + lane * 4;
      else o = (size_t)(bb * 64 + warp * 4 + (lane >> 3)) * S + s0 + r * 32 + (lane & 7) * 4;
      x[r] = __ldcs(reinterpret_cast<const float4*>(a + o));
    }
#pragma unroll
    for (int r = 0; r < DEPTH; ++r) acc += x[r].x + x[r].y + x[r].z + x[r].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}
// write-only: THREADS per CTA, each warp instruction stores 512 contiguous bytes (the synthesis epilogue's store shape)
__global__ void write_kernel(float4* __restrict__ c, size_t n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(c + i, v);
}
// plane-wise: CTA b writes plane p = b, b + grid, ... of 64 KB; each of the 8 warps writes 2 rows x 256 B per instruction
__global__ void __launch_bounds__(256, 1) write_plane_kernel(float* __restrict__ c, long long planes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2, rsub = lane >> 4, c4 = lane & 15;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (long long p = blockIdx.x; p < planes; p += gridDim.x) {
    float* base = c + (size_t)p * 128 * 128 + (size_t)(32 * q + rsub) * 128 + half * 64 + 4 * c4;
#pragma unroll
    for (int g = 0; g < 16; ++g) __stcs(reinterpret_cast<float4*>(base + (size_t)(2 * g) * 128), v);
  }
}
// write-only, the block tail's tile pattern (64 channel rows 8 MB apart x 128 points, two output tensors):
//   SEG = 128: a warp instruction stores 128 B of one row (lane = point, 4 B each) -- the epilogue's TMEM-lane mapping
//   SEG = 512: a warp instruction stores 512 B of one row (lane = 4 consecutive points)
template <int SEG>
__global__ void __launch_bounds__(512, 1) tile_write_kernel(float* __restrict__ c, float* __restrict__ d, long long S, long long ntiles,
                                                            long long tiles_per_batch, int contiguous) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tq = ntiles / gridDim.x;
  for (long long it = 0; it < tq; ++it) {
    const long long t = contiguous ? blockIdx.x * tq + it : blockIdx.x + it * gridDim.x;
    const long long bb = t / tiles_per_batch, s0 = (t - bb * tiles_per_batch) * 128;
    if (SEG == 128) {
      const int q = warp & 3, part = warp >> 2;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const size_t o = (size_t)(bb * 64 + part * 16 + j) * S + s0 + 32 * q + lane;
        __stcs(c + o, 1.f);
        d[o] = 2.f;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const size_t o = (size_t)(bb * 64 + warp * 4 + r) * S + s0 + lane * 4;
        __stcs(reinterpret_cast<float4*>(c + o), make_float4(1.f, 2.f, 3.f, 4.f));
        *reinterpret_cast<float4*>(d + o) = make_float4(1.f, 2.f, 3.f, 4.f);
      }
    }
  }
}
template <int NR, int NW>
__global__ void flat_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ c, float4* __restrict__ d, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 x = __ldcs(a + i);
    if (NR > 1) { float4 y = __ldcs(b + i); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
    __stcs(c + i, x);
    if (NW > 1) __stcs(d + i, x);
  }
}
template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}
int main() {
  const int B = 8; const long long S = 128LL * 128 * 128; const size_t n = (size_t)B * 64 * S;
  float *a, *b, *c, *d;
  cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&c, n * 4); cudaMalloc(&d, n * 4);
  cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4);
  const double gb = n * 4 / 1e9;
#define TILE(TP, NR, NW, GRID) { long long tpb = S / TP, nt = tpb * B; float ms = time_ms([&] { tile_kernel<TP, NR, NW><<<GRID, 512>>>(a, b, c, d, S, nt, tpb); }); \
    printf("tile TP=%4d reads=%d writes=%d grid=%4d: %.3f ms  %.0f GB/s\n", TP, NR, NW, GRID, ms, gb * (NR + NW) / ms * 1e3); }
#define FLAT(NR, NW, GRID) { float ms = time_ms([&] { flat_kernel<NR, NW><<<GRID, 512>>>((float4*)a, (float4*)b, (float4*)c, (float4*)d, n / 4); }); \
    printf("flat reads=%d writes=%d grid=%4d: %.3f ms  %.0f GB/s\n", NR, NW, GRID, ms, gb * (NR + NW) / ms * 1e3); }
  FLAT(1, 1, 148 * 4) FLAT(2, 2, 148 * 4) FLAT(2, 2, 148 * 16) FLAT(1, 1, 148 * 16)
  TILE(128, 2, 2, 148) TILE(256, 2, 2, 148) TILE(512, 2, 2, 148) TILE(1024, 2, 2, 148)
  TILE(128, 2, 2, 296) TILE(128, 2, 2, 592) TILE(256, 2, 2, 592) TILE(512, 2, 2, 592)
  TILE(128, 1, 1, 592) TILE(512, 1, 1, 592) TILE(128, 2, 1, 592)
#define READ(SEG, GRID) { long long tpb = S / 128, nt = tpb * B; float ms = time_ms([&] { read_kernel<SEG, 4><<<GRID, 512>>>(a, c, S, nt, tpb); }); \
    printf("read-only SEG=%3d grid=%4d: %.3f ms  %.0f GB/s\n", SEG, GRID, ms, gb / ms * 1e3); }
  READ(512, 148) READ(128, 148) READ(512, 296) READ(128, 296) READ(512, 592) READ(128, 592)
  { const size_t n4 = n / 4; for (int g : {148, 296, 592, 2368}) for (int th : {256, 512, 1024}) {
      float ms = time_ms([&] { write_kernel<<<g, th>>>((float4*)c, n4); });
      printf("write-only flat grid=%4d threads=%4d: %.3f ms  %.0f GB/s\n", g, th, ms, gb / ms * 1e3); } }
  { const long long planes = (long long)(n / (128 * 128)); for (int g : {148, 296}) {
      float ms = time_ms([&] { write_plane_kernel<<<g, 256>>>(c, planes); });
      printf("write-only planes grid=%4d: %.3f ms  %.0f GB/s\n", g, ms, gb / ms * 1e3); } }
  { long long tpb = S / 128, nt = tpb * B;
    for (int cont : {0, 1}) {
      float ms = time_ms([&] { tile_write_kernel<128><<<148, 512>>>(c, d, S, nt, tpb, cont); });
      printf("tile write-only SEG=128 contiguous=%d: %.3f ms  %.0f GB/s\n", cont, ms, 2 * gb / ms * 1e3);
      ms = time_ms([&] { tile_write_kernel<512><<<148, 512>>>(c, d, S, nt, tpb, cont); });
      printf("tile write-only SEG=512 contiguous=%d: %.3f ms  %.0f GB/s\n", cont, ms, 2 * gb / ms * 1e3);
      ms = time_ms([&] { tile_write_kernel<512><<<296, 512>>>(c, d, S, nt, tpb, cont); });
      printf("tile write-only SEG=512 grid 296 contiguous=%d: %.3f ms  %.0f GB/s\n", cont, ms, 2 * gb / ms * 1e3);
    } }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
