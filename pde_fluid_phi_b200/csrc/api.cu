// extern "C" entry points of the two transforms: argument checks and dispatch between the
// fp32 FFMA path (any shape) and the tcgen05 TF32 path (tensor-core shapes).
#include "common.cuh"

namespace pdephi {
int analysis_simt(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st);
int synthesis_simt(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                   const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                   float2* ws, cudaStream_t st);
bool tc_supported(const Plan* p);
int analysis_tc(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st);
int synthesis_tc(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                 const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                 float2* ws, cudaStream_t st);
bool tc3_supported(const Plan* p);
int analysis_tc3(const Plan* p, const float* x, float2* X, long long BC, int direction, float2* ws, cudaStream_t st);
int synthesis_tc3(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1,
                  const float* mode_scale, const float* row_scale, long long rs_stride, long long BC, int direction,
                  float2* ws, cudaStream_t st);
bool gen_supported(const Plan* p);
int analysis_gen(const Plan* p, const float* x, float2* X, long long BC, int direction, void* ws, cudaStream_t st);
int synthesis_gen(const Plan* p, const float2* Z, float* out, const float* add0, const float* add1, const float* mode_scale,
                  const float* row_scale, long long rs_stride, long long BC, int direction, void* ws, cudaStream_t st);
}  // namespace pdephi

using namespace pdephi;

extern "C" int pdephi_plan_supports(pdephi_plan_t plan, int precision) {
  const Plan* p = (const Plan*)plan;
  if (!p) return 0;
  if (precision == PDEPHI_PREC_FP32) return 1;
  if (precision == PDEPHI_PREC_TF32) return (tc_supported(p) || gen_supported(p)) ? 1 : 0;
  if (precision == PDEPHI_PREC_TF32X3) return tc3_supported(p) ? 1 : 0;
  return 0;
}

static int check_common(const Plan* p, long long BC, int direction, int precision, void* ws, size_t ws_bytes,
                        const char* who) {
  PDEPHI_REQUIRE(p != nullptr, "%s: null plan", who);
  PDEPHI_REQUIRE(BC > 0, "%s: batch*channels must be positive", who);
  PDEPHI_REQUIRE(direction == PDEPHI_FORWARD || direction == PDEPHI_ADJOINT, "%s: bad direction %d", who, direction);
  PDEPHI_REQUIRE(precision == PDEPHI_PREC_FP32 || precision == PDEPHI_PREC_TF32 || precision == PDEPHI_PREC_TF32X3,
                 "%s: bad precision %d", who, precision);
  int dev = -1;
  PDEPHI_CUDA(cudaGetDevice(&dev));
  PDEPHI_REQUIRE(dev == p->device, "%s: plan was created on device %d but the current device is %d", who, p->device, dev);
  const size_t need = pdephi_plan_workspace_bytes((pdephi_plan_t)p, BC);
  if (!ws || ws_bytes < need) return fail(PDEPHI_ERR_WORKSPACE, "%s: workspace %zu < %zu bytes", who, ws_bytes, need);
  return PDEPHI_OK;
}

extern "C" int pdephi_analysis_r2c(pdephi_plan_t plan, const float* x, float* X, long long BC, int direction,
                                   int precision, void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  const Plan* p = (const Plan*)plan;
  int rc = check_common(p, BC, direction, precision, workspace, workspace_bytes, "analysis_r2c");
  if (rc) return rc;
  PDEPHI_REQUIRE(x && X, "analysis_r2c: null pointer");
  if (precision == PDEPHI_PREC_TF32) {
    if (!tc_supported(p) && gen_supported(p))
      return analysis_gen(p, x, (float2*)X, BC, direction, workspace, (cudaStream_t)stream);
    if (!tc_supported(p))
      return fail(PDEPHI_ERR_UNSUPPORTED, "analysis_r2c: TF32 tensor-core path does not support grid (%d,%d,%d) modes (%d,%d,%d)",
                  p->Nx, p->Ny, p->Nz, p->mx, p->my, p->mzh);
    return analysis_tc(p, x, (float2*)X, BC, direction, (float2*)workspace, (cudaStream_t)stream);
  }
  if (precision == PDEPHI_PREC_TF32X3) {
    if (!tc3_supported(p))
      return fail(PDEPHI_ERR_UNSUPPORTED, "analysis_r2c: split-TF32 tensor-core path does not support grid (%d,%d,%d) modes (%d,%d,%d)",
                  p->Nx, p->Ny, p->Nz, p->mx, p->my, p->mzh);
    return analysis_tc3(p, x, (float2*)X, BC, direction, (float2*)workspace, (cudaStream_t)stream);
  }
  return analysis_simt(p, x, (float2*)X, BC, direction, (float2*)workspace, (cudaStream_t)stream);
}

extern "C" int pdephi_synthesis_c2r(pdephi_plan_t plan, const float* Z, float* out, const float* addend0,
                                    const float* addend1, const float* mode_scale, const float* row_scale,
                                    long long row_scale_stride, long long BC, int direction, int precision,
                                    void* workspace, size_t workspace_bytes, pdephi_stream_t stream) {
  const Plan* p = (const Plan*)plan;
  int rc = check_common(p, BC, direction, precision, workspace, workspace_bytes, "synthesis_c2r");
  if (rc) return rc;
  PDEPHI_REQUIRE(Z && out, "synthesis_c2r: null pointer");
  if (precision == PDEPHI_PREC_TF32) {
    if (!tc_supported(p) && gen_supported(p))
      return synthesis_gen(p, (const float2*)Z, out, addend0, addend1, mode_scale, row_scale, row_scale_stride, BC, direction,
                           workspace, (cudaStream_t)stream);
    if (!tc_supported(p))
      return fail(PDEPHI_ERR_UNSUPPORTED, "synthesis_c2r: TF32 tensor-core path does not support grid (%d,%d,%d) modes (%d,%d,%d)",
                  p->Nx, p->Ny, p->Nz, p->mx, p->my, p->mzh);
    return synthesis_tc(p, (const float2*)Z, out, addend0, addend1, mode_scale, row_scale, row_scale_stride, BC,
                        direction, (float2*)workspace, (cudaStream_t)stream);
  }
  if (precision == PDEPHI_PREC_TF32X3) {
    if (!tc3_supported(p))
      return fail(PDEPHI_ERR_UNSUPPORTED, "synthesis_c2r: split-TF32 tensor-core path does not support grid (%d,%d,%d) modes (%d,%d,%d)",
                  p->Nx, p->Ny, p->Nz, p->mx, p->my, p->mzh);
    return synthesis_tc3(p, (const float2*)Z, out, addend0, addend1, mode_scale, row_scale, row_scale_stride, BC,
                         direction, (float2*)workspace, (cudaStream_t)stream);
  }
  return synthesis_simt(p, (const float2*)Z, out, addend0, addend1, mode_scale, row_scale, row_scale_stride, BC,
                        direction, (float2*)workspace, (cudaStream_t)stream);
}
