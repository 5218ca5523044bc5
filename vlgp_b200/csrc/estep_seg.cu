// K2 host side: geometry of the SMEM-resident segment E-step and the choice of kernel variant (estep_seg_impl.cuh).
#include "estep_seg_impl.cuh"

using namespace segk;

namespace segk {
template <> int launch_seg_variant<2, 1>(vlgp_ctx *, TrialSet *, SegArgs &, size_t, bool *);
template <> int launch_seg_variant<2, 0>(vlgp_ctx *, TrialSet *, SegArgs &, size_t, bool *);
template <> int launch_seg_variant<4, 0>(vlgp_ctx *, TrialSet *, SegArgs &, size_t, bool *);
template <> int launch_seg_variant<4, 1>(vlgp_ctx *, TrialSet *, SegArgs &, size_t, bool *);
template <> int launch_seg_variant<2, 2>(vlgp_ctx *, TrialSet *, SegArgs &, size_t, bool *);
template <> int launch_seg_variant<4, 2>(vlgp_ctx *, TrialSet *, SegArgs &, size_t, bool *);
}

int vlgp_launch_estep_segments(vlgp_ctx *ctx, TrialSet *ts, int n_iter, double dmu_bound, int method_vb, bool *handled,
                               const int32_t *d_subset, int n_subset) {
    *handled = false;
    if (getenv("VLGP_FORCE_GENERIC_ESTEP")) return VLGP_OK;
    if (ts->min_len != ts->max_len || ts->max_len > VLGP_MAX_W || ts->factors.size() != 1) return VLGP_OK;
    if (ts->d_x) return VLGP_OK;             // general regressors: the any-length kernel takes the offsets einsum(x, b)
    const int W = ts->max_len, L = ctx->L, N = ctx->N;
    SegArgs p{};
    p.n_seg = d_subset ? n_subset : ts->n_trials; p.subset = d_subset; p.W = W; p.N = N; p.rank = ctx->rank;
    p.G = ts->factors[0].d_G;
    int goff = 0, moff = 0, po = 0, co = 0;
    p.use_dmma = getenv("VLGP_NO_DMMA_ESTEP") ? 0 : 1;
    for (int l = 0; l < L; ++l)
        if (ts->factors[0].h_ncol[l] > 32) p.use_dmma = 0;
    for (int l = 0; l < L; ++l) {
        const int nc = ts->factors[0].h_ncol[l];
        if (nc < 1) return VLGP_OK;          // degenerate factor: let the general kernel handle it
        p.nc[l] = nc;
        p.goff[l] = goff;
        p.moff[l] = moff;
        p.pairoff[l] = po;
        p.coloff[l] = co;
        const int nb8 = 8 * ((nc + 7) / 8);
        p.ldm[l] = p.use_dmma ? nb8 + 4 : (nc | 1);
        goff += W * (nc | 1);
        moff += p.use_dmma ? nb8 * (nb8 + 4) : nc * (nc | 1);
        po += nc * (nc + 1) / 2;
        co += nc;
    }
    for (int l = 0; l < L; ++l) p.ldg[l] = p.nc[l] | 1;
    p.pairoff[L] = po; p.coloff[L] = co;
    p.pair_total = po; p.col_total = co;
    p.g_total = goff; p.m_total = moff;
    p.y = ts->d_y; p.ydtype = ts->ydtype;
    p.mu = ts->d_mu; p.v = ts->d_v; p.w = ts->d_w; p.dmu = ts->d_dmu;
    p.a = ctx->d_a; p.b = ctx->d_b; p.noise = ctx->d_noise; p.poisson = ctx->d_poisson;
    p.n_iter = n_iter; p.dmu_bound = dmu_bound; p.method_vb = method_vb; p.flags = ctx->d_flags;
    p.skip = getenv("VLGP_DEBUG_SKIP") ? atoi(getenv("VLGP_DEBUG_SKIP")) : 0;
#ifdef VLGP_ESTEP_TWO_BINS
    p.tpb = NT / ((W + 1) / 2);          // threads per PAIR of bins (estep_seg_impl.cuh: rate_pass_two_bins)
#else
    p.tpb = NT / W;
#endif
    if (p.tpb < 1) return VLGP_OK;
    if (p.tpb > N) p.tpb = N;
    p.chunk = (N + p.tpb - 1) / p.tpb;
    p.np = 8 * ((N + 7) / 8);
    if (p.np % 16 == 0) p.np += 8;               // = 8 mod 16: the B-operand fragment loads are bank-conflict-free
    p.kp = 4 * ((2 * L + 1 + 3) / 4);
    const bool fast = !ctx->any_gauss && ts->ydtype == VLGP_Y_U8 && !getenv("VLGP_NO_FAST_ESTEP");
    p.fused = fast && p.use_dmma && !getenv("VLGP_ESTEP_NO_FUSED");
    p.f32 = p.fused && ctx->estep_f32;                            // vlgp_set_precision(ctx, 32)
    const bool stage_y = ts->ydtype == VLGP_Y_U8 && !p.fused;     // the fused pipeline reads the counts once, in place
    size_t smem = seg_smem_bytes(L, N, W, p.g_total, p.m_total, p.tpb, stage_y, p.kp, p.np, p.col_total, p.f32, p.fused);
    bool big = false;
    for (int l = 0; l < L; ++l)
        if (p.nc[l] > 16) big = true;
    if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin && p.use_dmma) {
        // many latents x many neurons x wide factors (config 3 in its first EM iterations: 10 x 200, 29 columns): leave
        // the factors where they are (L x W x rank doubles, L2-resident, shared by every segment) and keep the rest in SMEM
        p.g_global = 1;
        p.g_total = 0;
        for (int l = 0; l < L; ++l) {
            p.goff[l] = l * W * ctx->rank;
            p.ldg[l] = ctx->rank;
        }
        big = true;                          // the NBMAX = 4 instantiations carry the in-place path
        smem = seg_smem_bytes(L, N, W, 0, p.m_total, p.tpb, stage_y, p.kp, p.np, p.col_total, p.f32, p.fused);
    }
    if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin) return VLGP_OK;
    int rc = VLGP_OK;
    if (p.f32) rc = big ? launch_seg_variant<4, 2>(ctx, ts, p, smem, handled) : launch_seg_variant<2, 2>(ctx, ts, p, smem, handled);
    else if (big && p.use_dmma && fast) rc = launch_seg_variant<4, 1>(ctx, ts, p, smem, handled);
    else if (big && p.use_dmma) rc = launch_seg_variant<4, 0>(ctx, ts, p, smem, handled);
    else if (fast) rc = launch_seg_variant<2, 1>(ctx, ts, p, smem, handled);
    else rc = launch_seg_variant<2, 0>(ctx, ts, p, smem, handled);
    return rc;
}
