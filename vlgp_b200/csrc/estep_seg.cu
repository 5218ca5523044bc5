// K2: SMEM-resident E-step for window-length segments (placeholder until the specialised kernel lands: reports
// "not handled" so that the general kernel in estep.cu runs).
#include "common.cuh"

int vlgp_launch_estep_segments(vlgp_ctx *ctx, TrialSet *ts, int n_iter, double dmu_bound, int method_vb, bool *handled) {
    (void)ctx; (void)ts; (void)n_iter; (void)dmu_bound; (void)method_vb;
    *handled = false;
    return VLGP_OK;
}
