// C-ABI glue: weight-table enumeration, version and struct-size probes (include/graspnerf_b200.h).
#include "gn_weights.cuh"
#include "../../include/graspnerf_b200.h"

extern "C" int gn_weight_entry_count(void) { return GN_W_COUNT; }
extern "C" int gn_weight_blob_floats(void) { return GN_W_TOTAL; }
extern "C" int gn_weight_entry(int idx, const char** name, int* offset, int* rows, int* cols, int* cols_padded)
{
    if (idx < 0 || idx >= GN_W_COUNT) return -1;
    if (name) *name = kGnW[idx].name;
    if (offset) *offset = gn_w_off(idx);
    if (rows) *rows = kGnW[idx].rows;
    if (cols) *cols = kGnW[idx].cols;
    if (cols_padded) *cols_padded = kGnW[idx].cols_pad;
    return 0;
}
extern "C" const char* gn_version(void) { return "graspnerf_b200 0.1.0 (sm_100a)"; }
extern "C" int gn_sizeof_k1_params(void) { return (int)sizeof(GnK1Params); }
extern "C" int gn_sizeof_k2a_params(void) { return (int)sizeof(GnK2aParams); }
extern "C" int gn_sizeof_k2b_params(void) { return (int)sizeof(GnK2bParams); }
extern "C" int gn_sizeof_k3_params(void) { return (int)sizeof(GnK3Params); }
extern "C" int gn_sizeof_k2b_bwd_params(void) { return (int)sizeof(GnK2bBwdParams); }
extern "C" int gn_sizeof_k2a_bwd_params(void) { return (int)sizeof(GnK2aBwdParams); }
extern "C" int gn_sizeof_k1_bwd_params(void) { return (int)sizeof(GnK1BwdParams); }
extern "C" int gn_sizeof_ray_setup_params(void) { return (int)sizeof(GnRaySetupParams); }
extern "C" int gn_sizeof_depth_mean_params(void) { return (int)sizeof(GnDepthMeanParams); }
extern "C" int gn_sizeof_grasp_post_params(void) { return (int)sizeof(GnGraspPostParams); }
extern "C" int gn_sizeof_vgn_params(void) { return (int)sizeof(GnVgnParams); }
extern "C" int gn_sizeof_norm_act_pad_params(void) { return (int)sizeof(GnNormActPadParams); }
extern "C" int gn_sizeof_conv_params(void) { return (int)sizeof(GnConvParams); }
