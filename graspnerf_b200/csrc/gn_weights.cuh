// Weight-blob layout shared by the K2 kernels and the host packer (enumerated through the C ABI:
// gn_weight_entry*).  Every matrix is stored k-major, i.e. blob[off + k*cols_padded + n] = torch_weight[n][k]
// (nn.Linear keeps [out][in]); vectors (biases, 1-row matrices) are rows=1.
#pragma once

struct GnWEntry { const char* name; int rows; int cols; int cols_pad; };

// Row (input) orders that differ from the reference's concatenation order are noted; the host packer
// (graspnerf_b200/weights.py) performs the permutation.
//   f35 order here  : [img_feats 32 | rgb 3]           (reference ibrnet.py:458-459: [rgb 3 | img_feats 32])
#define GN_W_TABLE(X) \
    /* ray_dir_fc (ibrnet.py:382-385): 4 -> 16 -> 35, output columns in record order [img 32 | rgb 3 | pad] */ \
    X(RD_W0, "rd.w0", 4, 16, 16) X(RD_B0, "rd.b0", 1, 16, 16) X(RD_W1, "rd.w1", 16, 35, 36) X(RD_B1, "rd.b1", 1, 35, 36) \
    X(DD_MEAN_W0, "dd.mean.w0", 32, 32, 32) X(DD_MEAN_B0, "dd.mean.b0", 1, 32, 32) \
    X(DD_MEAN_W2, "dd.mean.w2", 32, 32, 32) X(DD_MEAN_B2, "dd.mean.b2", 1, 32, 32) \
    X(DD_MEAN_W4, "dd.mean.w4", 32, 2, 4)   X(DD_MEAN_B4, "dd.mean.b4", 1, 2, 4)   \
    X(DD_VAR_W0,  "dd.var.w0", 32, 32, 32)  X(DD_VAR_B0,  "dd.var.b0", 1, 32, 32)  \
    X(DD_VAR_W2,  "dd.var.w2", 32, 32, 32)  X(DD_VAR_B2,  "dd.var.b2", 1, 32, 32)  \
    X(DD_VAR_W4,  "dd.var.w4", 32, 2, 4)    X(DD_VAR_B4,  "dd.var.b4", 1, 2, 4)    \
    X(DD_AW_W0,   "dd.aw.w0", 32, 32, 32)   X(DD_AW_B0,   "dd.aw.b0", 1, 32, 32)   \
    X(DD_AW_W2,   "dd.aw.w2", 32, 32, 32)   X(DD_AW_B2,   "dd.aw.b2", 1, 32, 32)   \
    X(DD_AW_W4,   "dd.aw.w4", 32, 1, 4)     X(DD_AW_B4,   "dd.aw.b4", 1, 1, 4)     \
    /* prob_embed (aggregate_net.py:29-33): rows = [ray_feats 32, hit, vis] */ \
    X(PE_W0, "pe.w0", 34, 32, 32) X(PE_B0, "pe.b0", 1, 32, 32) X(PE_W2, "pe.w2", 32, 32, 32) X(PE_B2, "pe.b2", 1, 32, 32) \
    /* neuray_fc (ibrnet.py:419-423) */ \
    X(NF_W0, "nf.w0", 32, 8, 8) X(NF_B0, "nf.b0", 1, 8, 8) X(NF_W2, "nf.w2", 1, 8, 8) X(NF_B2, "nf.b2", 1, 1, 4) \
    /* base_fc (ibrnet.py:387-390) split by input block: global [mean0,var0,mean1,var1] (each img32|rgb3|pad -> 36), f35(+pad), prob_emb */ \
    X(BF_WG, "bf.wg", 144, 64, 64) X(BF_WF, "bf.wf", 36, 64, 64) X(BF_WP, "bf.wp", 32, 64, 64) X(BF_B0, "bf.b0", 1, 64, 64) \
    X(BF_W2, "bf.w2", 64, 32, 32) X(BF_B2, "bf.b2", 1, 32, 32) \
    /* vis_fc (ibrnet.py:392-396), vis_fc2 (398-402) */ \
    X(VF_W0, "vf.w0", 32, 32, 32) X(VF_B0, "vf.b0", 1, 32, 32) X(VF_W2, "vf.w2", 32, 33, 36) X(VF_B2, "vf.b2", 1, 33, 36) \
    X(V2_W0, "v2.w0", 32, 32, 32) X(V2_B0, "v2.b0", 1, 32, 32) X(V2_W2, "v2.w2", 1, 32, 32) X(V2_B2, "v2.b2", 1, 1, 4) \
    /* rgb_fc (ibrnet.py:413-417): rows = [x 32, vis 1, dir_diff 4] */ \
    X(RF_W0, "rf.w0", 37, 16, 16) X(RF_B0, "rf.b0", 1, 16, 16) X(RF_W2, "rf.w2", 16, 8, 8) X(RF_B2, "rf.b2", 1, 8, 8) \
    X(RF_W4, "rf.w4", 1, 8, 8) X(RF_B4, "rf.b4", 1, 1, 4) \
    /* geometry_fc (ibrnet.py:404-407): rows = [mean 32, var 32, wmean 1, embed 21] */ \
    X(GF_W0, "gf.w0", 86, 64, 64) X(GF_B0, "gf.b0", 1, 64, 64) X(GF_W2, "gf.w2", 64, 16, 16) X(GF_B2, "gf.b2", 1, 16, 16) \
    /* ray_attention (ibrnet.py:409, 52-102): k-major projections, LayerNorm */ \
    X(AT_WQ, "at.wq", 16, 16, 16) X(AT_WK, "at.wk", 16, 16, 16) X(AT_WV, "at.wv", 16, 16, 16) X(AT_FC, "at.fc", 16, 16, 16) \
    X(AT_LNW, "at.ln_w", 1, 16, 16) X(AT_LNB, "at.ln_b", 1, 16, 16) \
    /* out_geometry_fc (ibrnet.py:410-412) */ \
    X(OG_W0, "og.w0", 16, 16, 16) X(OG_B0, "og.b0", 1, 16, 16) X(OG_W1, "og.w1", 1, 16, 16) X(OG_B1, "og.b1", 1, 1, 4) \
    /* host-side algebraic fusions used by the tensor-core K2a: prob_embed.2 has no activation, so its consumers take the  \
       ReLU'd hidden e1 directly: neuray_fc.0 o prob_embed.2 (32 -> 8) and the prob_embed block of base_fc.0 o prob_embed.2 */ \
    X(NFC_W0, "nfc.w0", 32, 8, 8) X(NFC_B0, "nfc.b0", 1, 8, 8) X(BF_WPC, "bf.wpc", 32, 64, 64) X(BF_B0C, "bf.b0c", 1, 64, 64)

enum GnWIdx {
#define X(id, name, r, c, cp) GN_W_##id,
    GN_W_TABLE(X)
#undef X
    GN_W_COUNT
};

static constexpr GnWEntry kGnW[] = {
#define X(id, name, r, c, cp) {name, r, c, cp},
    GN_W_TABLE(X)
#undef X
};

constexpr int gn_w_size(int i) { return kGnW[i].rows * kGnW[i].cols_pad; }
constexpr int gn_w_off(int idx) {
    int o = 0;
    for (int i = 0; i < idx; ++i) o += gn_w_size(i);
    return o;
}
constexpr int GN_W_TOTAL = gn_w_off(GN_W_COUNT);
// K2a uses entries [DD_MEAN_W0, GF_W0); K2b uses [GF_W0, COUNT)
constexpr int GN_W_K2A_FLOATS = gn_w_off(GN_W_GF_W0);
constexpr int GN_W_K2B_OFF = gn_w_off(GN_W_GF_W0);
constexpr int GN_W_K2B_FLOATS = gn_w_off(GN_W_NFC_W0) - GN_W_K2B_OFF;       // geometry_fc .. out_geometry_fc
template <int I> struct GnOffT { static constexpr int value = gn_w_off(I); };
#define GN_OFF(id) (GnOffT<GN_W_##id>::value)
