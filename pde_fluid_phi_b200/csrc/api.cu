// extern "C" entry points of the two transforms: argument checks and dispatch between the
// fp32 FFMA path (any shape) and the tcgen05 TF32 path (tensor-core shapes).
#include "common.cuh"

namespace pdephi {
int analysis_simt(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st, int stages, float2* T);
int synthesis_simt(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                   const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                   float2* ws, cudaStream_t st, int stages, float2* T);
bool tc_supported(const Plan* p);
int analysis_tc(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st, int stages, float2* T);
int synthesis_tc(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                 const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                 float2* ws, cudaStream_t st, int stages, float2* T);
bool tc3_supported(const Plan* p);
int analysis_tc3(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st, int stages, float2* T);
int synthesis_tc3(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                  const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                  float2* ws, cudaStream_t st, int stages, float2* T);
bool gen_supported(const Plan* p);
bool gen_x3_supported(const Plan* p);
int analysis_gen(const Plan* p, const float* x, float2* X, long long BC, int direction, void* ws, cudaStream_t st, int stages, float2* T,
                 bool x3);
int synthesis_gen(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1, const float* mode_scale,
                  const float* row_scale, long long rs_stride, long long BC, int direction, void* ws, cudaStream_t st, int stages, float2* T,
                  bool x3);
}  // namespace pdephi

using namespace pdephi;

extern "C" int pdephi_plan_supports(pdephi_plan_t plan, int precision) {
  const Plan* p = (const Plan*)plan;
  if (!p) return 0;
  if (precision == PDEPHI_PREC_FP32) return 1;
  if (precision == PDEPHI_PREC_TF32) return (tc_supported(p) || gen_supported(p)) ? 1 : 0;
  if (precision == PDEPHI_PREC_TF32X3) return (tc3_supported(p) || gen_x3_supported(p)) ? 1 : 0;
  return 0;
}

extern "C" int pdephi_plan_route(pdephi_plan_t plan, int precision) {
  const Plan* p = (const Plan*)plan;
  if (!p) return PDEPHI_ROUTE_NONE;
  if (precision == PDEPHI_PREC_FP32) return PDEPHI_ROUTE_FFMA;
  if (precision == PDEPHI_PREC_TF32)
    return tc_supported(p) ? PDEPHI_ROUTE_FUSED_TC : (gen_supported(p) ? PDEPHI_ROUTE_PER_AXIS_TC : PDEPHI_ROUTE_NONE);
  if (precision == PDEPHI_PREC_TF32X3)
    return tc3_supported(p) ? PDEPHI_ROUTE_FUSED_TC : (gen_x3_supported(p) ? PDEPHI_ROUTE_PER_AXIS_TC : PDEPHI_ROUTE_NONE);
  return PDEPHI_ROUTE_NONE;
}

static int check_common(const Plan* p, long long BC, int direction, int precision, int stages, void* ws, size_t ws_bytes,
                        const char* who) {
  PDEPHI_REQUIRE(p != nullptr, "%s: null plan", who);
  PDEPHI_REQUIRE(BC > 0, "%s: batch*channels must be positive", who);
  PDEPHI_REQUIRE(direction == PDEPHI_FORWARD || direction == PDEPHI_ADJOINT, "%s: bad direction %d", who, direction);
  PDEPHI_REQUIRE(precision == PDEPHI_PREC_FP32 || precision == PDEPHI_PREC_TF32 || precision == PDEPHI_PREC_TF32X3,
                 "%s: bad precision %d", who, precision);
  PDEPHI_REQUIRE(stages == PDEPHI_STAGES_ALL || stages == PDEPHI_STAGES_ZY || stages == PDEPHI_STAGES_X, "%s: bad stages %d", who,
                 stages);
  int dev = -1;
  PDEPHI_CUDA(cudaGetDevice(&dev));
  PDEPHI_REQUIRE(dev == p->device, "%s: plan was created on device %d but the current device is %d", who, p->device, dev);
  const size_t need = pdephi_plan_workspace_bytes((pdephi_plan_t)p, BC);
  if (!ws || ws_bytes < need) return fail(PDEPHI_ERR_WORKSPACE, "%s: workspace %zu < %zu bytes", who, ws_bytes, need);
  if (precision == PDEPHI_PREC_TF32 && !tc_supported(p) && !gen_supported(p))
    return fail(PDEPHI_ERR_UNSUPPORTED, "%s: TF32 tensor-core path does not support grid (%d,%d,%d) modes (%d,%d,%d)", who, p->Nx,
                p->Ny, p->Nz, p->mx, p->my, p->mzh);
  if (precision == PDEPHI_PREC_TF32X3 && !tc3_supported(p) && !gen_x3_supported(p))
    return fail(PDEPHI_ERR_UNSUPPORTED, "%s: split-TF32 tensor-core path does not support grid (%d,%d,%d) modes (%d,%d,%d)", who,
                p->Nx, p->Ny, p->Nz, p->mx, p->my, p->mzh);
  return PDEPHI_OK;
}

// in / out by stage selection:  ALL: x -> X;  ZY: x -> T;  X: T -> X   (T = [BC, Nx, my, mzh] complex)
static int run_analysis(const Plan* p, const float* in, float* out, long long BC, int direction, int precision, int stages,
                        void* workspace, cudaStream_t st) {
  const float* x = (stages & ST_ZY) ? in : nullptr;
  float2* X = (stages & ST_X) ? (float2*)out : nullptr;
  float2* T = stages == ST_ZY ? (float2*)out : (stages == ST_X ? (float2*)const_cast<float*>(in) : nullptr);
  if (precision == PDEPHI_PREC_TF32) {
    if (tc_supported(p)) return analysis_tc(p, x, X, BC, direction, (float2*)workspace, st, stages, T);
    return analysis_gen(p, x, X, BC, direction, workspace, st, stages, T, false);
  }
  if (precision == PDEPHI_PREC_TF32X3) {
    if (tc3_supported(p)) return analysis_tc3(p, x, X, BC, direction, (float2*)workspace, st, stages, T);
    return analysis_gen(p, x, X, BC, direction, workspace, st, stages, T, true);
  }
  return analysis_simt(p, x, X, BC, direction, (float2*)workspace, st, stages, T);
}

// ALL: Z -> out;  X: Z -> T (scales applied);  ZY: T -> out (+ addends)
static int run_synthesis(const Plan* p, const float* in, float* out, const float* a0, const float* a1, const float* mode_scale,
                         const float* row_scale, long long rs_stride, long long BC, int direction, int precision, int stages,
                         void* workspace, cudaStream_t st) {
  const float2* Z = (stages & ST_X) ? (const float2*)in : nullptr;
  float* o = (stages & ST_ZY) ? out : nullptr;
  float2* T = stages == ST_X ? (float2*)out : (stages == ST_ZY ? (float2*)const_cast<float*>(in) : nullptr);
  if (precision == PDEPHI_PREC_TF32) {
    if (tc_supported(p))
      return synthesis_tc(p, Z, o, a0, a1, mode_scale, row_scale, rs_stride, BC, direction, (float2*)workspace, st, stages, T);
    return synthesis_gen(p, Z, o, a0, a1, mode_scale, row_scale, rs_stride, BC, direction, workspace, st, stages, T, false);
  }
  if (precision == PDEPHI_PREC_TF32X3) {
    if (tc3_supported(p))
      return synthesis_tc3(p, Z, o, a0, a1, mode_scale, row_scale, rs_stride, BC, direction, (float2*)workspace, st, stages, T);
    return synthesis_gen(p, Z, o, a0, a1, mode_scale, row_scale, rs_stride, BC, direction, workspace, st, stages, T, true);
  }
  return synthesis_simt(p, Z, o, a0, a1, mode_scale, row_scale, rs_stride, BC, direction, (float2*)workspace, st, stages, T);
}

extern "C" int pdephi_analysis_stages(pdephi_plan_t plan, const float* in, float* out, long long BC, int direction, int precision,
                                      int stages, void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  const Plan* p = (const Plan*)plan;
  int rc = check_common(p, BC, direction, precision, stages, workspace, workspace_bytes, "analysis");
  if (rc) return rc;
  PDEPHI_REQUIRE(in && out, "analysis: null pointer");
  return run_analysis(p, in, out, BC, direction, precision, stages, workspace, (cudaStream_t)stream);
}

extern "C" int pdephi_analysis_r2c(pdephi_plan_t plan, const float* x, float* X, long long BC, int direction,
                                   int precision, void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  return pdephi_analysis_stages(plan, x, X, BC, direction, precision, PDEPHI_STAGES_ALL, workspace, workspace_bytes, stream);
}

extern "C" int pdephi_synthesis_stages(pdephi_plan_t plan, const float* in, float* out, const float* addend0,
                                       const float* addend1, const float* mode_scale, const float* row_scale,
                                       long long row_scale_stride, long long BC, int direction, int precision, int stages,
                                       void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  const Plan* p = (const Plan*)plan;
  int rc = check_common(p, BC, direction, precision, stages, workspace, workspace_bytes, "synthesis");
  if (rc) return rc;
  PDEPHI_REQUIRE(in && out, "synthesis: null pointer");
  PDEPHI_REQUIRE(stages != PDEPHI_STAGES_X || (!addend0 && !addend1), "synthesis: the x stage alone takes no addends");
  PDEPHI_REQUIRE(stages != PDEPHI_STAGES_ZY || (!mode_scale && !row_scale), "synthesis: the (z,y) stages alone take no mode scaling");
  return run_synthesis(p, in, out, addend0, addend1, mode_scale, row_scale, row_scale_stride, BC, direction, precision, stages,
                       workspace, (cudaStream_t)stream);
}

extern "C" int pdephi_synthesis_c2r(pdephi_plan_t plan, const float* Z, float* out, const float* addend0,
                                    const float* addend1, const float* mode_scale, const float* row_scale,
                                    long long row_scale_stride, long long BC, int direction, int precision,
                                    void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  return pdephi_synthesis_stages(plan, Z, out, addend0, addend1, mode_scale, row_scale, row_scale_stride, BC, direction,
                                 precision, PDEPHI_STAGES_ALL, workspace, workspace_bytes, stream);
}
