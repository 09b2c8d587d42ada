// Backward kernels (placeholder translation unit: entry points are filled in by the backward milestone).
#include "l3d_common.cuh"

#define L3D_NOT_YET(name) do { l3d_set_error(name ": backward kernel not built yet"); return 99; } while (0)

extern "C" int l3d_merge_bwd(const l3d_act *, const l3d_act *, const l3d_act *, const l3d_act *, const l3d_act *, const l3d_norm *,
                             const l3d_act *, const l3d_norm *, int, int, int, int, float, const float *, int, const float *,
                             const float *, float *, float *, const l3d_act *, double *, double *, void *) { L3D_NOT_YET("l3d_merge_bwd"); }
extern "C" int l3d_pw_bwd(const l3d_act *, const l3d_act *, const l3d_norm *, const double *, const l3d_act *, const l3d_norm *,
                          int, int, int, int, const float *, float *, const l3d_act *, int, void *) { L3D_NOT_YET("l3d_pw_bwd"); }
extern "C" int l3d_dw_bwd(const l3d_act *, const l3d_act *, const l3d_norm *, int, int, int, int, const float *, float *,
                          const l3d_act *, int, double *, void *) { L3D_NOT_YET("l3d_dw_bwd"); }
extern "C" int l3d_conv3_bwd(const l3d_act *, const l3d_act *, const l3d_norm *, const double *, const l3d_act *, const l3d_norm *,
                             int, int, int, int, const float *, int, float *, const l3d_act *, int, double *, void *) { L3D_NOT_YET("l3d_conv3_bwd"); }
extern "C" int l3d_convt_bwd(const l3d_act *, int, int, int, int, int, int, const l3d_act *, int, int, int, int, const float *,
                             float *, float *, const l3d_act *, int, void *) { L3D_NOT_YET("l3d_convt_bwd"); }
extern "C" int l3d_norm_param_grad(const double *, int, int, float *, float *, void *) { L3D_NOT_YET("l3d_norm_param_grad"); }
