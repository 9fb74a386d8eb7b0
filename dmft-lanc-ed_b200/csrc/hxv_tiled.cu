// hxv_tiled.cu -- shared-memory staged two-pass H*v (placeholder until the tiled kernels land).
#include "engine.h"
bool tiled_supported(const edgpu_ctx *) { return false; }
int tiled_plan_build(edgpu_ctx *) { return EDGPU_OK; }
int tiled_plan_free(edgpu_ctx *) { return EDGPU_OK; }
int tiled_apply_local(edgpu_ctx *, const double *, double *) { return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "tiled H*v not built"); }
