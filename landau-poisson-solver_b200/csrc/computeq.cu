// Tiled ComputeQ kernel (placeholder until the register-tiled version lands): returning -1
// tells lp_launch_computeQ to use the simple per-xi kernel.
#include "lpgpu_internal.h"
int lp_launch_computeQ_tiled(lpgpu_ctx *, const double *, double *, int) { return -1; }
