// Plan object: constant DFT twiddle tables, generated on the device in fp64 and rounded once.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace pdephi {

static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- launch counters / event timing -------------------------------------------------------------
static const char* const kKernelNames[K_COUNT] = {
    "analysis_zy_kernel", "xstage_kernel(analysis)", "xstage_kernel(synthesis)", "synthesis_zy_kernel",
    "analysis_zy_tc_kernel", "synthesis_zy_tc_kernel", "rational_mix_kernel", "mix_from_R_kernel",
    "projection_stats_kernel", "projection_bwd_kernel", "rational_dpq_kernel", "dpq_reduce_kernel",
    "complex_mix_kernel", "complex_mix_kernel(transposed)", "complex_dw_kernel", "pointwise_linear_kernel",
    "pointwise_wgrad_kernel", "block_tail_fwd_kernel", "block_tail_bwd_kernel", "analysis_x_tc_kernel",
    "synthesis_x_tc_kernel", "scale_modes_kernel", "rational_g_kernel", "analysis_zy_tc3_kernel",
    "synthesis_zy_tc3_kernel", "analysis_x_tc3_kernel", "synthesis_x_tc3_kernel", "gen_z_analysis_kernel",
    "gen_z_synthesis_kernel", "gen_axis_analysis_kernel", "gen_axis_synthesis_kernel", "block_spectral_fwd_kernel",
    "block_spectral_bwd_kernel", "rational_transfer_kernel", "mix_apply_kernel"};
struct ProfRec { int id; cudaEvent_t a, b; };
// option defaults: the x stage of the single-pass TF32 route runs split (fp32-grade) -- see pdephi_set_option
static std::atomic<int> g_options[OPT_COUNT] = {{1}, {2}};
static std::atomic<long long> g_launches[K_COUNT];
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_open[K_COUNT];
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
int option_get(int key) { return (key >= 0 && key < OPT_COUNT) ? g_options[key].load() : 0; }

void prof_begin(int id, cudaStream_t st) {
  g_launches[id].fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  g_open[id].push_back(e);
}
void prof_end(int id, cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_open[id].empty()) return;
  cudaEvent_t a = g_open[id].back();
  g_open[id].pop_back();
  cudaEvent_t b = get_event();
  cudaEventRecord(b, st);
  g_recs.push_back({id, a, b});
}

// table[a][b] = w(k) * (cos, sgn*sin)(2 pi k n / N); k_is_col selects which index is the mode.
// The product k*n is reduced mod N in integers so the angle is exact to fp64 rounding.
__global__ void twiddle_kernel(float2* __restrict__ out, int rows, int cols, int N, int n_offset, bool k_is_col,
                               double sgn, int weight_mode /*0 none, 1 hermitian c_k/total*/, int Nz_even_nyq,
                               double inv_total) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx % cols;
  const int k = k_is_col ? c : r;
  const int n = (k_is_col ? r : c) + n_offset;
  const long long kn = ((long long)k * n) % N;
  double s, co;
  sincospi(2.0 * (double)kn / (double)N, &s, &co);
  double w = 1.0;
  if (weight_mode == 1) {
    const double ck = (k == 0 || k == Nz_even_nyq) ? 1.0 : 2.0;
    w = ck * inv_total;
  }
  out[idx] = make_float2((float)(w * co), (float)(w * sgn * s));
}

}  // namespace pdephi

using namespace pdephi;

extern "C" int pdephi_version(void) { return PDEPHI_VERSION; }
extern "C" const char* pdephi_last_error(void) { return err_buf(); }

static int plan_build(Plan* p, int device, int Nx, int Ny, int Nz, int mx, int my, int mzh, int Nx_total, int x_offset);

extern "C" int pdephi_plan_create(pdephi_plan_t* out, int device, int Nx, int Ny, int Nz, int mx, int my, int mzh) {
  return pdephi_plan_create_slab(out, device, Nx, Ny, Nz, mx, my, mzh, Nx, 0);
}

extern "C" int pdephi_plan_create_slab(pdephi_plan_t* out, int device, int Nx, int Ny, int Nz, int mx, int my, int mzh,
                                       int Nx_total, int x_offset) {
  PDEPHI_REQUIRE(out != nullptr, "plan_create: null output handle");
  PDEPHI_REQUIRE(Nx > 0 && Ny > 0 && Nz > 0 && mx > 0 && my > 0 && mzh > 0, "plan_create: non-positive size");
  PDEPHI_REQUIRE(x_offset >= 0 && x_offset + Nx <= Nx_total, "plan_create: x slab [%d, %d) outside [0, %d)", x_offset,
                 x_offset + Nx, Nx_total);
  PDEPHI_REQUIRE(mx <= Nx_total && my <= Ny && mzh <= Nz / 2 + 1,
                 "plan_create: retained modes (%d,%d,%d) exceed the half-spectrum of grid (%d,%d,%d)", mx, my, mzh,
                 Nx_total, Ny, Nz);
  int cur = -1;
  PDEPHI_CUDA(cudaGetDevice(&cur));
  PDEPHI_CUDA(cudaSetDevice(device));
  Plan* p = new Plan();
  memset(p, 0, sizeof(Plan));
  // every failure below unwinds through here: the partially built plan and its device arenas are released and the
  // caller's current device is restored (an out-of-memory error on one GPU must not leave the thread on another)
  const int rc = plan_build(p, device, Nx, Ny, Nz, mx, my, mzh, Nx_total, x_offset);
  if (rc != PDEPHI_OK) pdephi_plan_destroy((pdephi_plan_t)p);
  cudaSetDevice(cur);
  if (rc != PDEPHI_OK) return rc;
  *out = (pdephi_plan_t)p;
  return PDEPHI_OK;
}

static int plan_build(Plan* p, int device, int Nx, int Ny, int Nz, int mx, int my, int mzh, int Nx_total, int x_offset) {
  p->device = device;
  p->Nx = Nx; p->Ny = Ny; p->Nz = Nz; p->mx = mx; p->my = my; p->mzh = mzh;
  p->Nx_total = Nx_total; p->x_offset = x_offset;
  cudaDeviceProp prop;
  PDEPHI_CUDA(cudaGetDeviceProperties(&prop, device));
  p->num_sms = prop.multiProcessorCount;
  p->smem_optin = prop.sharedMemPerBlockOptin;

  const size_t n_az = (size_t)Nz * mzh, n_ay = (size_t)Ny * my, n_ax = (size_t)Nx * mx;
  const size_t sizes[8] = {n_az, n_az, n_ay, n_ax, n_ax, n_ay, n_az, n_az};
  size_t off[9];
  off[0] = 0;
  for (int i = 0; i < 8; ++i) off[i + 1] = off[i] + align_up(sizes[i] * sizeof(float2), 256);
  PDEPHI_CUDA(cudaMalloc(&p->arena, off[8]));
  char* base = (char*)p->arena;
  p->az[0] = (float2*)(base + off[0]);
  p->az[1] = (float2*)(base + off[1]);
  p->ay = (float2*)(base + off[2]);
  p->ax = (float2*)(base + off[3]);
  p->sx = (float2*)(base + off[4]);
  p->sy = (float2*)(base + off[5]);
  p->sz[0] = (float2*)(base + off[6]);
  p->sz[1] = (float2*)(base + off[7]);

  const double inv_total = 1.0 / ((double)Nx_total * (double)Ny * (double)Nz);
  const int nyq = (Nz % 2 == 0 && mzh - 1 == Nz / 2) ? Nz / 2 : -1;
  auto gen = [&](float2* t, int rows, int cols, int N, int off, bool k_is_col, double sgn, int wm) {
    const int n = rows * cols;
    twiddle_kernel<<<(n + 255) / 256, 256>>>(t, rows, cols, N, off, k_is_col, sgn, wm, nyq, inv_total);
  };
  gen(p->az[0], Nz, mzh, Nz, 0, true, -1.0, 0);
  gen(p->az[1], Nz, mzh, Nz, 0, true, -1.0, 1);
  gen(p->ay, Ny, my, Ny, 0, true, -1.0, 0);
  gen(p->ax, Nx, mx, Nx_total, x_offset, true, -1.0, 0);      // slab rows carry their global x index
  gen(p->sx, mx, Nx, Nx_total, x_offset, false, +1.0, 0);
  gen(p->sy, my, Ny, Ny, 0, false, +1.0, 0);
  gen(p->sz[0], mzh, Nz, Nz, 0, false, +1.0, 1);
  gen(p->sz[1], mzh, Nz, Nz, 0, false, +1.0, 0);
  PDEPHI_LAUNCH_CHECK("twiddle_kernel");
  {
    int rc = tc_plan_init(p);
    if (rc) return rc;
    rc = xtc_plan_init(p);
    if (rc) return rc;
    rc = tc3_plan_init(p);
    if (rc) return rc;
    rc = gen_plan_init(p);
    if (rc) return rc;
  }
  PDEPHI_CUDA(cudaDeviceSynchronize());  // plan creation is the one synchronous call
  return PDEPHI_OK;
}

extern "C" int pdephi_plan_destroy(pdephi_plan_t h) {
  if (!h) return PDEPHI_OK;
  Plan* p = (Plan*)h;
  if (p->arena) cudaFree(p->arena);
  tc_plan_destroy(p);
  xtc_plan_destroy(p);
  tc3_plan_destroy(p);
  gen_plan_destroy(p);
  delete p;
  return PDEPHI_OK;
}

extern "C" size_t pdephi_plan_workspace_bytes(pdephi_plan_t h, long long BC) {
  if (!h || BC <= 0) return 0;
  const Plan* p = (const Plan*)h;
  // intermediate between the fused (z,y) stage and the x stage: [BC, Nx, my, mzh] complex, followed by a
  // [BC, mx, my, mzh] complex scratch for the projected modes of the tensor-core synthesis x stage
  const size_t fused = align_up((size_t)BC * p->Nx * p->my * p->mzh * sizeof(float2), 1024) +
                       align_up((size_t)BC * p->mx * p->my * p->mzh * sizeof(float2), 1024);
  // the per-axis tensor-core route (shapes the fused kernels do not cover) stages one more, larger intermediate
  const size_t general = (p->gen_ok && !p->tc_ok) ? gen_workspace_bytes(p, BC) : 0;
  return fused > general ? fused : general;
}

extern "C" int pdephi_set_option(int key, int value) {
  PDEPHI_REQUIRE(key >= 0 && key < OPT_COUNT, "set_option: unknown option %d", key);
  g_options[key].store(value);
  return PDEPHI_OK;
}
extern "C" int pdephi_get_option(int key) { return option_get(key); }

extern "C" int pdephi_profile_enable(int on) {
  g_prof_on.store(on != 0);
  return PDEPHI_OK;
}
extern "C" int pdephi_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < K_COUNT; ++i) g_launches[i].store(0);
  for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_recs.clear();
  return PDEPHI_OK;
}
extern "C" int pdephi_profile_num_kernels(void) { return K_COUNT; }
extern "C" const char* pdephi_profile_kernel_name(int id) { return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : ""; }
extern "C" int pdephi_profile_query(int id, long long* launches, double* total_ms) {
  PDEPHI_REQUIRE(id >= 0 && id < K_COUNT, "profile_query: bad kernel id %d", id);
  if (launches) *launches = g_launches[id].load();
  if (total_ms) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double tot = 0.0;
    for (auto& r : g_recs) {
      if (r.id != id) continue;
      PDEPHI_CUDA(cudaEventSynchronize(r.b));
      float ms = 0.f;
      PDEPHI_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      tot += ms;
    }
    *total_ms = tot;
  }
  return PDEPHI_OK;
}
