// GPU probe: what a persistent one-CTA-per-SM TMA load pipeline delivers (no MMA, no stores), by box shape,
// swizzle mode, row stride and ring depth.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../pde_fluid_phi_b200/csrc -o tma_probe.bin tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "tc_ptx.cuh"
using namespace pdephi::tcptx;

// tile = NBOX boxes of [ROWS rows][32 floats]; STAGES ring; consumers: NCONS warps that (optionally) read the tile
template <int ROWS, int NBOX, int STAGES, int NCONS, bool TOUCH, int NPROD>
__global__ void __launch_bounds__(32 * (NPROD + NCONS), 1)
tma_kernel(const __grid_constant__ CUtensorMap map, float* sink, long long ntiles, long long tiles_per_group, int rows_per_group) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  constexpr uint32_t BOXB = ROWS * 128, TILEB = NBOX * BOXB;
  const uint32_t bars = base + STAGES * TILEB;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (STAGES + i); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full(i), NPROD); mbar_init(empty(i), NCONS); }
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (warp < NPROD) {
    if (elect_one()) {
      long long it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = (int)(it % STAGES);
        mbar_wait(empty(s), (uint32_t)(((it / STAGES) & 1) ^ 1));
        mbar_arrive_expect_tx(full(s), TILEB / NPROD);
        const long long g = t / tiles_per_group, c0 = (t - g * tiles_per_group) * (32 * NBOX);
        for (int j = warp; j < NBOX; j += NPROD)
          tma_load_2d(base + s * TILEB + j * BOXB, &map, full(s), (int)(c0 + 32 * j), (int)(g * rows_per_group));
      }
    }
  } else {
    float acc = 0.f;
    long long it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = (int)(it % STAGES);
      mbar_wait(full(s), (uint32_t)((it / STAGES) & 1));
      if (TOUCH) {
        const float4* p = reinterpret_cast<const float4*>(sm + s * TILEB);
        for (int i = (warp - NPROD) * 32 + lane; i < (int)(TILEB / 16); i += NCONS * 32) { float4 v = p[i]; acc += v.x + v.w; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(s));
    }
    if (acc == 123.456f) sink[0] = acc;
  }
}

// the same ring fed by cp.async loader warps (block_tc.cu): NLW warps, each warp instruction = one 512-byte row
template <int STAGES, int NLW, int NCONS, int HINT>
__global__ void __launch_bounds__(32 * (NLW + NCONS), 1)
cpasync_kernel(const float* __restrict__ a, float* sink, long long S, long long ntiles, long long tiles_per_batch) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t TILEB = 64 * 512;      // 64 rows x 128 points
  const uint32_t bars = base + STAGES * TILEB;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (STAGES + i); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full(i), NLW * 32); mbar_init(empty(i), NCONS); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (warp < NLW) {
    constexpr int RPW = 64 / NLW;
    for (long long it = 0; it < my_tiles + (STAGES - 1); ++it) {
      if (it < my_tiles) {
        const long long t = blockIdx.x + it * gridDim.x;
        const int s = (int)(it % STAGES);
        mbar_wait(empty(s), (uint32_t)(((it / STAGES) & 1) ^ 1));
        const long long b = t / tiles_per_batch, s0 = (t - b * tiles_per_batch) * 128;
        const float* src = a + ((size_t)b * 64 + warp * RPW) * S + s0 + lane * 4;
        const uint32_t dst = base + s * TILEB + (warp * RPW) * 512 + lane * 16;
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          if (HINT == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 512), "l"(src + (size_t)i * S) : "memory");
          if (HINT == 1) asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(dst + i * 512), "l"(src + (size_t)i * S) : "memory");
          if (HINT == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + i * 512), "l"(src + (size_t)i * S) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (it >= STAGES - 1) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        fence_proxy_async();
        mbar_arrive(full((int)((it - (STAGES - 1)) % STAGES)));
      }
    }
  } else {
    for (long long it = 0; it < my_tiles; ++it) {
      const int s = (int)(it % STAGES);
      mbar_wait(full(s), (uint32_t)((it / STAGES) & 1));
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(s));
    }
  }
  if (threadIdx.x == 12345) sink[0] = 0.f;
}
template <int STAGES, int NLW, int NCONS, int HINT>
void run_cp(const char* name, const float* a, float* sink, long long rows, long long S, int grid = 148) {
  const long long tpb = S / 128, ntiles = tpb * (rows / 64);
  auto k = cpasync_kernel<STAGES, NLW, NCONS, HINT>;
  const int smem = STAGES * 64 * 512 + 1024 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float ms = time_ms([&] { k<<<grid, 32 * (NLW + NCONS), smem>>>(a, sink, S, ntiles, tpb); });
  cudaError_t e = cudaGetLastError();
  printf("%-58s loaders %2d stages %d: %.3f ms  %.0f GB/s %s\n", name, NLW, STAGES, ms, rows * S * 4 / 1e9 / ms * 1e3,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (EncodeTiledFn)p;
}
static CUtensorMap make_map(const float* ptr, long long rows, long long cols, int box_rows, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
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
template <int ROWS, int NBOX, int STAGES, int NCONS, bool TOUCH, int NPROD>
void run(const char* name, const float* a, float* sink, long long rows, long long cols, int rows_per_group, CUtensorMapSwizzle sw,
         CUtensorMapL2promotion l2, int grid = 148) {
  CUtensorMap m = make_map(a, rows, cols, ROWS, sw, l2);
  const long long tpg = cols / (32 * NBOX), ntiles = tpg * (rows / rows_per_group);
  auto k = tma_kernel<ROWS, NBOX, STAGES, NCONS, TOUCH, NPROD>;
  const int smem = STAGES * NBOX * ROWS * 128 + 1024 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float ms = time_ms([&] { k<<<grid, 32 * (NPROD + NCONS), smem>>>(m, sink, ntiles, tpg, rows_per_group); });
  cudaError_t e = cudaGetLastError();
  printf("%-58s rows/box %3d boxes %d stages %d: %.3f ms  %.0f GB/s %s\n", name, ROWS, NBOX, STAGES, ms, rows * cols * 4 / 1e9 / ms * 1e3,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  const long long S = 128LL * 128 * 128; const long long rows = 8 * 64;
  float *a, *sink;
  cudaMalloc(&a, rows * S * 4); cudaMalloc(&sink, 16);
  cudaMemset(a, 0, rows * S * 4);
  const auto SW = CU_TENSOR_MAP_SWIZZLE_128B, SWA = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, NOSW = CU_TENSOR_MAP_SWIZZLE_NONE;
  const auto L2 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B, L1 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B, L0 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  // block-tail shape: [512 rows][S], tile = 64 rows x 128 points
  run<64, 4, 3, 4, false, 1>("block-tail tile, sw128 atom32, L2 256B", a, sink, rows, S, 64, SWA, L2);
  run<64, 4, 3, 4, false, 1>("block-tail tile, sw128, L2 256B", a, sink, rows, S, 64, SW, L2);
  run<64, 4, 3, 4, false, 1>("block-tail tile, no swizzle, L2 256B", a, sink, rows, S, 64, NOSW, L2);
  run<64, 4, 3, 4, false, 1>("block-tail tile, sw128, L2 128B", a, sink, rows, S, 64, SW, L1);
  run<64, 4, 3, 4, false, 1>("block-tail tile, sw128, no L2 promotion", a, sink, rows, S, 64, SW, L0);
  run<64, 4, 6, 4, false, 1>("block-tail tile, 6 stages", a, sink, rows, S, 64, SW, L2);
  run<64, 4, 6, 4, false, 4>("block-tail tile, 6 stages, 4 producer warps", a, sink, rows, S, 64, SW, L2);
  run<64, 4, 3, 4, true, 1>("block-tail tile, consumers read smem", a, sink, rows, S, 64, SW, L2);
  run<64, 4, 3, 4, false, 1>("block-tail tile, grid 296 (2 CTAs/SM)", a, sink, rows, S, 64, SW, L2, 296);
  run<64, 2, 3, 4, false, 1>("half tile (64 pts), grid 592 (4 CTAs/SM)", a, sink, rows, S, 64, SW, L2, 592);
  run_cp<3, 4, 4, 1>("cp.async ring, .cg L2::256B", a, sink, rows, S);
  run_cp<6, 4, 4, 1>("cp.async ring, .cg L2::256B", a, sink, rows, S);
  run_cp<6, 8, 4, 1>("cp.async ring, .cg L2::256B", a, sink, rows, S);
  run_cp<6, 16, 4, 1>("cp.async ring, .cg L2::256B", a, sink, rows, S);
  run_cp<6, 8, 4, 0>("cp.async ring, .cg", a, sink, rows, S);
  run_cp<6, 8, 4, 2>("cp.async ring, .ca", a, sink, rows, S);
  run_cp<3, 8, 4, 1>("cp.async ring, .cg L2::256B", a, sink, rows, S);
  run_cp<3, 16, 4, 1>("cp.async ring, .cg L2::256B", a, sink, rows, S);
  run_cp<3, 8, 4, 1>("cp.async ring, .cg L2::256B, 2 CTAs/SM", a, sink, rows, S, 296);
  // analysis shape: [BC*Nx*128 rows][128], plane = 128 rows x 128 floats contiguous
  run<128, 4, 2, 4, false, 1>("analysis plane, sw128, 2 stages", a, sink, rows * S / 128, 128, 128, SW, L2);
  run<128, 4, 3, 4, false, 1>("analysis plane, sw128, 3 stages", a, sink, rows * S / 128, 128, 128, SW, L2);
  run<128, 1, 8, 4, false, 1>("analysis k-block ring (16 KB x 8)", a, sink, rows * S / 128, 128, 128, SW, L2);
  run<128, 1, 12, 4, false, 1>("analysis k-block ring (16 KB x 12)", a, sink, rows * S / 128, 128, 128, SW, L2);
  run<128, 4, 3, 4, false, 4>("analysis plane, 3 stages, 4 producer warps", a, sink, rows * S / 128, 128, 128, SW, L2);
  run<256, 2, 3, 4, false, 1>("256-row boxes x2 (64 KB), 3 stages", a, sink, rows * S / 128, 128, 256, SW, L2);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
