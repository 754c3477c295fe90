// K2a backward: reverse pass of the per-(point,view) head + cross-view poolings for training (first order).
//
// Forward chain (recomputed per row from the saved K1 record; same thread <-> (point, view) layout and shuffle poolings as
// the fp32 forward k2a_head_simt.cu; reference lines in that file's header):
//   dist decoder (3 MLPs) -> compute_prob -> hit, vis -> prob_embed -> neuray_fc -> w0
//   ray_dir_fc -> f = [img_feats | rgb] + dfeat ;  (mean0,var0) = pool(f, w0) ; (mean1,var1) = pool(f, wgt)
//   base_fc([mean0,var0,mean1,var1, f, prob_emb]) -> x0 ; vis_fc(x0*wgt) -> x = x0 + res, vis1 ; vis_fc2(x*vis1) -> vis2
//   w2 = vis2 / (sum_v vis2 + 1e-8) ; pooled = [pool(x, w2), mean_v(w2)]
// Reverse: d pooled[65] -> d rec.ray_feats[32], d rec.img_feats[32] per row, and the gradient of every weight of the chain
// (dd.*, pe.*, nf.*, rd.*, bf.*, vf.*, v2.*), accumulated with coalesced atomics in the blob layout.
// Activations live in thread-local arrays (registers / stack); weights sit in shared memory; the per-layer weight gradient
// is a CTA-wide rows^T x rows product staged through shared memory (gn_bwd.cuh).  With d_colors (RGB head) the reverse of
// rgb_fc + the softmax colour blend (ibrnet.py:507-511) is added: d colours -> d x, d vis2 and the rf.* weight gradients.
#include "gn_bwd.cuh"
#include "gn_weights.cuh"
#include "../../include/graspnerf_b200.h"

// Same fast-math activations as the inference kernels (a full-precision variant changed no gradient beyond the last digit:
// the residual fp32 noise of the compute_prob chain comes from ReLU gates at the switching point, see tests/test_gpu_backward.py)
__device__ __forceinline__ float kb_elu(float x) { return gn_elu(x); }
__device__ __forceinline__ float kb_sigmoid(float x) { return gn_sigmoid(x); }
__device__ __forceinline__ float kb_softplus(float x) { return gn_softplus(x); }

#define KB_THREADS 256
#define KB_WARPS (KB_THREADS / 32)
#define FULL 0xffffffffu
static_assert(64 * KB_THREADS <= KB_THREADS * GN_BWD_LDZ, "scratch columns must fit in the dZ tile");
static_assert(((size_t)GN_W_K2A_FLOATS + (size_t)KB_THREADS * (GN_BWD_LDX + GN_BWD_LDZ)) * 4 <= 227 * 1024, "K2a backward shared-memory budget");

template <int NP>
__device__ __forceinline__ void load_bias(const float* __restrict__ b, float* y)
{
#pragma unroll
    for (int n = 0; n < NP; ++n) y[n] = b[n];
}
__device__ __forceinline__ float gsum(float t, int gb, int V)
{
    float s = 0.f;
    for (int jv = 0; jv < V; ++jv) s += __shfl_sync(FULL, t, (gb + jv) & 31);
    return s;
}

// ---- dist-decoder MLP 32 -> 32 -> 32 -> NO (dist_decoder.py:62-86), activations kept for the reverse pass
template <int W0, int B0, int W2, int B2, int W4, int B4>
__device__ __forceinline__ void dd_fwd(const float* __restrict__ sw, const float* ray, float* h1, float* h2, float* o, float* scr)
{
    load_bias<32>(sw + B0, h1); mv_acc_rolled<32, 32, KB_THREADS>(sw + W0, ray, h1, scr);
#pragma unroll
    for (int c = 0; c < 32; ++c) h1[c] = kb_elu(h1[c]);
    load_bias<32>(sw + B2, h2); mv_acc_rolled<32, 32, KB_THREADS>(sw + W2, h1, h2, scr);
#pragma unroll
    for (int c = 0; c < 32; ++c) h2[c] = kb_elu(h2[c]);
    load_bias<4>(sw + B4, o); mv_acc_rolled<32, 4, KB_THREADS>(sw + W4, h2, o, scr);
}
template <int W0, int B0, int W2, int B2, int W4, int B4>
__device__ __forceinline__ void dd_bwd(const float* __restrict__ sw, double* gw, const float* ray, const float* h1, const float* h2,
                                       const float* d_o, float* d_ray, float* sX, float* sZ, float* scr)
{
    dw_layer<32, 4, 4>(gw + W4, gw + B4, h2, d_o, sX, sZ, KB_THREADS);
    float dh2[32], dh1[32];
    mv_bwd_rolled<32, 4, false, KB_THREADS>(sw + W4, d_o, dh2, scr);
#pragma unroll
    for (int c = 0; c < 32; ++c) dh2[c] *= gn_delu(h2[c]);
    dw_layer<32, 32, 32>(gw + W2, gw + B2, h1, dh2, sX, sZ, KB_THREADS);
    mv_bwd_rolled<32, 32, false, KB_THREADS>(sw + W2, dh2, dh1, scr);
#pragma unroll
    for (int c = 0; c < 32; ++c) dh1[c] *= gn_delu(h1[c]);
    dw_layer<32, 32, 32>(gw + W0, gw + B0, ray, dh1, sX, sZ, KB_THREADS);
    mv_bwd_rolled<32, 32, true, KB_THREADS>(sw + W0, dh1, d_ray, scr);
}
#define DD_IDS(n) GN_OFF(DD_##n##_W0), GN_OFF(DD_##n##_B0), GN_OFF(DD_##n##_W2), GN_OFF(DD_##n##_B2), GN_OFF(DD_##n##_W4), GN_OFF(DD_##n##_B4)

// weighted mean / variance over the V views of a point (ibrnet.py:112-116), 36-wide with a zero pad lane
__device__ __forceinline__ void pool_fwd36(const float* f, float w, int gb, int V, float* mu, float* var)
{
#pragma unroll
    for (int c = 0; c < 35; ++c) mu[c] = gsum(w * f[c], gb, V);
#pragma unroll
    for (int c = 0; c < 35; ++c) { const float d = f[c] - mu[c]; var[c] = gsum(w * d * d, gb, V); }
    mu[35] = 0.f; var[35] = 0.f;
}
// reverse of mean = sum_v w_v x_v, var = sum_v w_v (x_v - mean)^2 for this row: dx += ..., returns d w (this row)
template <int C>
__device__ __forceinline__ float pool_bwd(const float* x, const float* mu, const float* dmu, const float* dvar, float w, float wsum, float* dx)
{
    float dw = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float dl = x[c] - mu[c];
        const float dm_tot = dmu[c] - 2.f * dvar[c] * mu[c] * (1.f - wsum);     // d mean incl. its use inside var
        dx[c] += w * (dm_tot + 2.f * dvar[c] * dl);
        dw = fmaf(dm_tot, x[c], fmaf(dvar[c] * dl, dl, dw));
    }
    return dw;
}

__global__ void __launch_bounds__(KB_THREADS, 1)
gn_k2a_backward_kernel(const __grid_constant__ GnK2aBwdParams p, int num_tiles, int G)
{
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                                   // weights [GN_W_K2A_FLOATS]
    float* sX = sw + GN_W_K2A_FLOATS;                   // [256][GN_BWD_LDX]
    float* sZ = sX + KB_THREADS * GN_BWD_LDX;           // [256][GN_BWD_LDZ]
    for (int i = threadIdx.x * 4; i < GN_W_K2A_FLOATS; i += KB_THREADS * 4)
        *reinterpret_cast<float4*>(sw + i) = ldg4(p.weights + i);
    __syncthreads();
    float* scr = sZ + threadIdx.x;                      // private scratch column scr[k * KB_THREADS], k < 64 (aliases the dZ tile)
    double* gw = p.d_weights;

    const int V = p.V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lane_active = lane < G * V;
    const int g = lane_active ? lane / V : 0;
    const int v = lane_active ? lane - g * V : 0;
    const int gb = g * V;
    const long long total_pts = (long long)p.B * p.N;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        long long pidx = ((long long)tile * KB_WARPS + warp) * G + g;
        const bool valid = lane_active && pidx < total_pts;
        pidx = pidx < total_pts ? pidx : total_pts - 1;
        const int b = (int)(pidx / p.N);
        const int n = (int)(pidx - (long long)b * p.N);
        const float* row = p.rec + ((size_t)pidx * V + v) * GN_REC_STRIDE;
        const float2 ptv = __ldg(reinterpret_cast<const float2*>(p.pt + (size_t)pidx * GN_PT_STRIDE));
        const float4 tail = ldg4(row + GN_REC_RGB);
        const float mask = (valid && ((__float_as_uint(ptv.y) >> v) & 1u)) ? 1.f : 0.f;
        const float depth = tail.w;
        const float nvalid = ptv.x;
        const float wgt = __fdiv_rn(mask, nvalid + 1e-8f);

        // ========================================= forward =========================================================
        float ray[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            const float4 t = ldg4(row + GN_REC_RAYF + c);
            ray[c] = t.x; ray[c + 1] = t.y; ray[c + 2] = t.z; ray[c + 3] = t.w;
        }
        float h1m[32], h2m[32], h1v[32], h2v[32], h1a[32], h2a[32], om[4], ov[4], oa[4];
        dd_fwd<DD_IDS(MEAN)>(sw, ray, h1m, h2m, om, scr);
        dd_fwd<DD_IDS(VAR)>(sw, ray, h1v, h2v, ov, scr);
        dd_fwd<DD_IDS(AW)>(sw, ray, h1a, h2a, oa, scr);
        const float mean0 = kb_softplus(om[0]), mean1 = kb_softplus(om[1]);
        const float var0 = kb_softplus(ov[0]) + 0.05f, var1 = kb_softplus(ov[1]) + 0.05f;
        const float aw = kb_sigmoid(oa[0]);
        float nearp, farp;
        {
            const float* dr = p.depth_range + ((size_t)b * V + v) * 2;
            const float rnear = __fdiv_rn(-1.f, __ldg(dr)), rfar = __fdiv_rn(-1.f, __ldg(dr + 1));
            float d = __fdiv_rn(-1.f, fmaxf(depth, 1e-5f));
            d = __fdiv_rn(d - rnear, rfar - rnear);
            if (p.que_dists == nullptr) { nearp = d - 0.005f; farp = d + 0.005f; }
            else {
                const int smp = n % p.dn;
                const float* qd = p.que_dists + (size_t)b * p.N + n;
                const float h_cur = __ldg(qd) * 0.5f;
                const float h_prev = smp > 0 ? __ldg(qd - 1) * 0.5f : h_cur;
                nearp = d - h_prev; farp = d + h_cur;
            }
        }
        const float c00 = 0.5f + 0.5f * tanhf((nearp - mean0) * var0), c10 = 0.5f + 0.5f * tanhf((farp - mean0) * var0);
        const float c01 = 0.5f + 0.5f * tanhf((nearp - mean1) * var1), c11 = 0.5f + 0.5f * tanhf((farp - mean1) * var1);
        const float mix1 = 1.f - aw;
        const float vis = ((1.f - c00) * aw + (1.f - c01) * mix1) * mask;
        const float hit = ((c10 - c00) * aw + (c11 - c01) * mix1) * mask;
        // prob_embed
        float xin[34], e1[32], pe[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) xin[c] = ray[c];
        xin[32] = (hit - 0.5f) * 2.f; xin[33] = (vis - 0.5f) * 2.f;
        load_bias<32>(sw + GN_OFF(PE_B0), e1);
        mv_acc_rolled<34, 32, KB_THREADS>(sw + GN_OFF(PE_W0), xin, e1, scr);
#pragma unroll
        for (int c = 0; c < 32; ++c) e1[c] = fmaxf(e1[c], 0.f);
        load_bias<32>(sw + GN_OFF(PE_B2), pe);
        mv_acc_rolled<32, 32, KB_THREADS>(sw + GN_OFF(PE_W2), e1, pe, scr);
        // neuray_fc -> w0
        float t8[8], sig0, w0;
        load_bias<8>(sw + GN_OFF(NF_B0), t8);
        mv_acc_rolled<32, 8, KB_THREADS>(sw + GN_OFF(NF_W0), pe, t8, scr);
        {
            float s = sw[GN_OFF(NF_B2)];
#pragma unroll
            for (int c = 0; c < 8; ++c) { t8[c] = kb_elu(t8[c]); s = fmaf(sw[GN_OFF(NF_W2) + c], t8[c], s); }
            sig0 = kb_sigmoid(s);
            w0 = sig0 * wgt;
        }
        // ray_dir_fc and f
        float dd[4], hid[16], dfe[36], f[36];
        {
            const float4 ddv = ldg4(row + GN_REC_DD);
            dd[0] = ddv.x; dd[1] = ddv.y; dd[2] = ddv.z; dd[3] = ddv.w;
            load_bias<16>(sw + GN_OFF(RD_B0), hid);
            mv_acc_rolled<4, 16, KB_THREADS>(sw + GN_OFF(RD_W0), dd, hid, scr);
#pragma unroll
            for (int c = 0; c < 16; ++c) hid[c] = kb_elu(hid[c]);
            load_bias<36>(sw + GN_OFF(RD_B1), dfe);
            mv_acc_rolled<16, 36, KB_THREADS>(sw + GN_OFF(RD_W1), hid, dfe, scr);
#pragma unroll
            for (int c = 0; c < 36; ++c) dfe[c] = kb_elu(dfe[c]);
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = ldg4(row + GN_REC_IMGF + c);
                f[c] = dfe[c] + t.x; f[c + 1] = dfe[c + 1] + t.y; f[c + 2] = dfe[c + 2] + t.z; f[c + 3] = dfe[c + 3] + t.w;
            }
            f[32] = dfe[32] + tail.x; f[33] = dfe[33] + tail.y; f[34] = dfe[34] + tail.z; f[35] = 0.f;
        }
        float m0[36], v0[36], m1[36], v1[36];
        pool_fwd36(f, w0, gb, V, m0, v0);
        pool_fwd36(f, wgt, gb, V, m1, v1);
        const float S0 = gsum(w0, gb, V), S1 = gsum(wgt, gb, V);
        // base_fc
        float bh[64], x0[32];
        load_bias<64>(sw + GN_OFF(BF_B0), bh);
        mv_acc_rolled<36, 64, KB_THREADS>(sw + GN_OFF(BF_WG), m0, bh, scr);
        mv_acc_rolled<36, 64, KB_THREADS>(sw + GN_OFF(BF_WG) + 36 * 64, v0, bh, scr);
        mv_acc_rolled<36, 64, KB_THREADS>(sw + GN_OFF(BF_WG) + 72 * 64, m1, bh, scr);
        mv_acc_rolled<36, 64, KB_THREADS>(sw + GN_OFF(BF_WG) + 108 * 64, v1, bh, scr);
        mv_acc_rolled<36, 64, KB_THREADS>(sw + GN_OFF(BF_WF), f, bh, scr);
        mv_acc_rolled<32, 64, KB_THREADS>(sw + GN_OFF(BF_WP), pe, bh, scr);
#pragma unroll
        for (int c = 0; c < 64; ++c) bh[c] = kb_elu(bh[c]);
        load_bias<32>(sw + GN_OFF(BF_B2), x0);
        mv_acc_rolled<64, 32, KB_THREADS>(sw + GN_OFF(BF_W2), bh, x0, scr);
#pragma unroll
        for (int c = 0; c < 32; ++c) x0[c] = kb_elu(x0[c]);
        // vis_fc
        float xw[32], vh[32], xv[36], x[32], sig1, vis1;
#pragma unroll
        for (int c = 0; c < 32; ++c) xw[c] = x0[c] * wgt;
        load_bias<32>(sw + GN_OFF(VF_B0), vh);
        mv_acc_rolled<32, 32, KB_THREADS>(sw + GN_OFF(VF_W0), xw, vh, scr);
#pragma unroll
        for (int c = 0; c < 32; ++c) vh[c] = kb_elu(vh[c]);
        load_bias<36>(sw + GN_OFF(VF_B2), xv);
        mv_acc_rolled<32, 36, KB_THREADS>(sw + GN_OFF(VF_W2), vh, xv, scr);
#pragma unroll
        for (int c = 0; c < 36; ++c) xv[c] = kb_elu(xv[c]);
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = x0[c] + xv[c];
        sig1 = kb_sigmoid(xv[32]);
        vis1 = sig1 * mask;
        // vis_fc2
        float xs[32], v2h[32], sig2, vis2;
#pragma unroll
        for (int c = 0; c < 32; ++c) xs[c] = x[c] * vis1;
        load_bias<32>(sw + GN_OFF(V2_B0), v2h);
        mv_acc_rolled<32, 32, KB_THREADS>(sw + GN_OFF(V2_W0), xs, v2h, scr);
        {
            float s = sw[GN_OFF(V2_B2)];
#pragma unroll
            for (int c = 0; c < 32; ++c) { v2h[c] = kb_elu(v2h[c]); s = fmaf(sw[GN_OFF(V2_W2) + c], v2h[c], s); }
            sig2 = kb_sigmoid(s);
            vis2 = sig2 * mask;
        }
        // rgb_fc + softmax blend over the views (ibrnet.py:507-511), RGB head only
        float rin[40], r16[16], r8[8], bw = 0.f;
        const bool with_rgb = p.d_colors != nullptr;
        if (with_rgb) {
#pragma unroll
            for (int c = 0; c < 32; ++c) rin[c] = x[c];
            rin[32] = vis2; rin[33] = dd[0]; rin[34] = dd[1]; rin[35] = dd[2]; rin[36] = dd[3]; rin[37] = rin[38] = rin[39] = 0.f;
            load_bias<16>(sw + GN_OFF(RF_B0), r16);
            mv_acc_rolled<37, 16, KB_THREADS>(sw + GN_OFF(RF_W0), rin, r16, scr);
#pragma unroll
            for (int c = 0; c < 16; ++c) r16[c] = kb_elu(r16[c]);
            load_bias<8>(sw + GN_OFF(RF_B2), r8);
            mv_acc_rolled<16, 8, KB_THREADS>(sw + GN_OFF(RF_W2), r16, r8, scr);
            float logit = sw[GN_OFF(RF_B4)];
#pragma unroll
            for (int c = 0; c < 8; ++c) { r8[c] = kb_elu(r8[c]); logit = fmaf(sw[GN_OFF(RF_W4) + c], r8[c], logit); }
            if (mask == 0.f) logit = -1e9f;
            float mx = -INFINITY;
            for (int jv = 0; jv < V; ++jv) mx = fmaxf(mx, __shfl_sync(FULL, logit, (gb + jv) & 31));
            const float e = __expf(logit - mx);
            bw = __fdiv_rn(e, gsum(e, gb, V));
        }
        const float Sv = gsum(vis2, gb, V) + 1e-8f;
        const float w2 = __fdiv_rn(vis2, Sv);
        const float S2 = gsum(w2, gb, V);
        float mu[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) mu[c] = gsum(w2 * x[c], gb, V);

        // ========================================= reverse =========================================================
        float dmu[32], dvar[32];
        float dwm;
        {
            const float* dp = p.d_pooled + (size_t)pidx * GN_POOL_STRIDE;
            const float live = valid ? 1.f : 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 a = ldg4(dp + c), bq = ldg4(dp + 32 + c);
                dmu[c] = a.x * live; dmu[c + 1] = a.y * live; dmu[c + 2] = a.z * live; dmu[c + 3] = a.w * live;
                dvar[c] = bq.x * live; dvar[c + 1] = bq.y * live; dvar[c + 2] = bq.z * live; dvar[c + 3] = bq.w * live;
            }
            dwm = __ldg(dp + 64) * live;
        }
        float dx[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) dx[c] = 0.f;
        float dw2 = pool_bwd<32>(x, mu, dmu, dvar, w2, S2, dx) + dwm / (float)V;        // pooled[64] = mean_v(w2)
        const float one = 1.f;
        float dvis2_rgb = 0.f;
        if (with_rgb) {
            // colours = sum_v bw_v * rgb_v (rgb carries the mask); softmax over the views; rgb_fc 37 -> 16 -> 8 -> 1
            const float4 dc = ldg4(p.d_colors + (size_t)pidx * 4);
            const float dbw = valid ? (dc.x * tail.x + dc.y * tail.y + dc.z * tail.z) : 0.f;
            const float dlogit = mask * bw * (dbw - gsum(bw * dbw, gb, V));             // a masked logit is the constant -1e9
            dw_layer<1, 8, 8>(gw + GN_OFF(RF_W4), nullptr, &dlogit, r8, sX, sZ, KB_THREADS);
            dw_layer<1, 1, 4>(gw + GN_OFF(RF_B4), nullptr, &one, &dlogit, sX, sZ, KB_THREADS);
            float dr8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) dr8[c] = sw[GN_OFF(RF_W4) + c] * dlogit * gn_delu(r8[c]);
            dw_layer<16, 8, 8>(gw + GN_OFF(RF_W2), gw + GN_OFF(RF_B2), r16, dr8, sX, sZ, KB_THREADS);
            float dr16[16];
            mv_bwd_rolled<16, 8, false, KB_THREADS>(sw + GN_OFF(RF_W2), dr8, dr16, scr);
#pragma unroll
            for (int c = 0; c < 16; ++c) dr16[c] *= gn_delu(r16[c]);
            dw_layer<37, 16, 16>(gw + GN_OFF(RF_W0), gw + GN_OFF(RF_B0), rin, dr16, sX, sZ, KB_THREADS);
            float drin[37];
            mv_bwd_rolled<37, 16, false, KB_THREADS>(sw + GN_OFF(RF_W0), dr16, drin, scr);
#pragma unroll
            for (int c = 0; c < 32; ++c) dx[c] += drin[c];
            dvis2_rgb = drin[32];                                                       // dir_diff carries no gradient
        }
        // w2 = vis2 / Sv
        const float dvis2 = (dw2 - gsum(dw2 * w2, gb, V)) / Sv + dvis2_rgb;
        const float ds2 = dvis2 * mask * sig2 * (1.f - sig2);
        // vis_fc2.2 (row vector) and bias
        dw_layer<1, 32, 32>(gw + GN_OFF(V2_W2), nullptr, &ds2, v2h, sX, sZ, KB_THREADS);
        dw_layer<1, 1, 4>(gw + GN_OFF(V2_B2), nullptr, &one, &ds2, sX, sZ, KB_THREADS);
        float dt[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) dt[c] = sw[GN_OFF(V2_W2) + c] * ds2 * gn_delu(v2h[c]);
        dw_layer<32, 32, 32>(gw + GN_OFF(V2_W0), gw + GN_OFF(V2_B0), xs, dt, sX, sZ, KB_THREADS);
        float dxs[32];
        mv_bwd_rolled<32, 32, false, KB_THREADS>(sw + GN_OFF(V2_W0), dt, dxs, scr);
        float dvis1 = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) { dx[c] = fmaf(dxs[c], vis1, dx[c]); dvis1 = fmaf(dxs[c], x[c], dvis1); }
        // x = x0 + xv[0..31] ; vis1 = sigmoid(xv[32]) * mask ; xv = elu(vf.2(vh))
        float dxv[36];
#pragma unroll
        for (int c = 0; c < 32; ++c) dxv[c] = dx[c] * gn_delu(xv[c]);
        dxv[32] = dvis1 * mask * sig1 * (1.f - sig1) * gn_delu(xv[32]);
        dxv[33] = 0.f; dxv[34] = 0.f; dxv[35] = 0.f;
        dw_layer<32, 36, 36>(gw + GN_OFF(VF_W2), gw + GN_OFF(VF_B2), vh, dxv, sX, sZ, KB_THREADS);
        float dvh[32];
        mv_bwd_rolled<32, 36, false, KB_THREADS>(sw + GN_OFF(VF_W2), dxv, dvh, scr);
#pragma unroll
        for (int c = 0; c < 32; ++c) dvh[c] *= gn_delu(vh[c]);
        dw_layer<32, 32, 32>(gw + GN_OFF(VF_W0), gw + GN_OFF(VF_B0), xw, dvh, sX, sZ, KB_THREADS);
        float dx0[32];
        mv_bwd_rolled<32, 32, false, KB_THREADS>(sw + GN_OFF(VF_W0), dvh, dx0, scr);
#pragma unroll
        for (int c = 0; c < 32; ++c) dx0[c] = (fmaf(dx0[c], wgt, dx[c])) * gn_delu(x0[c]);      // d pre-activation of base_fc.2
        dw_layer<64, 32, 32>(gw + GN_OFF(BF_W2), gw + GN_OFF(BF_B2), bh, dx0, sX, sZ, KB_THREADS);
        float dbh[64];
        mv_bwd_rolled<64, 32, false, KB_THREADS>(sw + GN_OFF(BF_W2), dx0, dbh, scr);
#pragma unroll
        for (int c = 0; c < 64; ++c) dbh[c] *= gn_delu(bh[c]);
        // base_fc.0 : inputs [m0 36 | v0 36 | m1 36 | v1 36 | f 36 | pe 32]
        dw_layer<36, 64, 64>(gw + GN_OFF(BF_WG), nullptr, m0, dbh, sX, sZ, KB_THREADS);
        dw_layer<36, 64, 64>(gw + GN_OFF(BF_WG) + 36 * 64, nullptr, v0, dbh, sX, sZ, KB_THREADS);
        dw_layer<36, 64, 64>(gw + GN_OFF(BF_WG) + 72 * 64, nullptr, m1, dbh, sX, sZ, KB_THREADS);
        dw_layer<36, 64, 64>(gw + GN_OFF(BF_WG) + 108 * 64, nullptr, v1, dbh, sX, sZ, KB_THREADS);
        dw_layer<36, 64, 64>(gw + GN_OFF(BF_WF), gw + GN_OFF(BF_B0), f, dbh, sX, sZ, KB_THREADS);
        dw_layer<32, 64, 64>(gw + GN_OFF(BF_WP), nullptr, pe, dbh, sX, sZ, KB_THREADS);
        float df[36], dpe[32];
        mv_bwd_rolled<36, 64, false, KB_THREADS>(sw + GN_OFF(BF_WF), dbh, df, scr);
        mv_bwd_rolled<32, 64, false, KB_THREADS>(sw + GN_OFF(BF_WP), dbh, dpe, scr);
        float dw0;
        {
            // pooled statistics are shared by the V rows of the point: sum the per-row cotangents over the group
            float dm[36], dv_[36];
            mv_bwd_rolled<36, 64, false, KB_THREADS>(sw + GN_OFF(BF_WG), dbh, dm, scr);
            mv_bwd_rolled<36, 64, false, KB_THREADS>(sw + GN_OFF(BF_WG) + 36 * 64, dbh, dv_, scr);
#pragma unroll
            for (int c = 0; c < 35; ++c) { dm[c] = gsum(dm[c], gb, V); dv_[c] = gsum(dv_[c], gb, V); }
            dw0 = pool_bwd<35>(f, m0, dm, dv_, w0, S0, df);
            mv_bwd_rolled<36, 64, false, KB_THREADS>(sw + GN_OFF(BF_WG) + 72 * 64, dbh, dm, scr);
            mv_bwd_rolled<36, 64, false, KB_THREADS>(sw + GN_OFF(BF_WG) + 108 * 64, dbh, dv_, scr);
#pragma unroll
            for (int c = 0; c < 35; ++c) { dm[c] = gsum(dm[c], gb, V); dv_[c] = gsum(dv_[c], gb, V); }
            (void)pool_bwd<35>(f, m1, dm, dv_, wgt, S1, df);                    // wgt carries no gradient
        }
        // w0 = sigmoid(nf) * wgt ; neuray_fc
        {
            const float dsn = dw0 * wgt * sig0 * (1.f - sig0);
            dw_layer<1, 8, 8>(gw + GN_OFF(NF_W2), nullptr, &dsn, t8, sX, sZ, KB_THREADS);
            dw_layer<1, 1, 4>(gw + GN_OFF(NF_B2), nullptr, &one, &dsn, sX, sZ, KB_THREADS);
            float dt8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) dt8[c] = sw[GN_OFF(NF_W2) + c] * dsn * gn_delu(t8[c]);
            dw_layer<32, 8, 8>(gw + GN_OFF(NF_W0), gw + GN_OFF(NF_B0), pe, dt8, sX, sZ, KB_THREADS);
            mv_bwd_rolled<32, 8, true, KB_THREADS>(sw + GN_OFF(NF_W0), dt8, dpe, scr);
        }
        // f = dfeat + [img_feats | rgb] : img_feats gradient leaves here; ray_dir_fc weights
        float* drow = p.d_rec + ((size_t)pidx * V + v) * 64;
        if (valid) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) st4(drow + 32 + c, make_float4(df[c], df[c + 1], df[c + 2], df[c + 3]));
        }
        {
            float ddf[36];
#pragma unroll
            for (int c = 0; c < 35; ++c) ddf[c] = df[c] * gn_delu(dfe[c]);
            ddf[35] = 0.f;
            dw_layer<16, 36, 36>(gw + GN_OFF(RD_W1), gw + GN_OFF(RD_B1), hid, ddf, sX, sZ, KB_THREADS);
            float dhid[16];
            mv_bwd_rolled<16, 36, false, KB_THREADS>(sw + GN_OFF(RD_W1), ddf, dhid, scr);
#pragma unroll
            for (int c = 0; c < 16; ++c) dhid[c] *= gn_delu(hid[c]);
            dw_layer<4, 16, 16>(gw + GN_OFF(RD_W0), gw + GN_OFF(RD_B0), dd, dhid, sX, sZ, KB_THREADS);
        }
        // prob_embed
        float dray[32];
        float dhit, dvis;
        {
            dw_layer<32, 32, 32>(gw + GN_OFF(PE_W2), gw + GN_OFF(PE_B2), e1, dpe, sX, sZ, KB_THREADS);
            float de1[32];
            mv_bwd_rolled<32, 32, false, KB_THREADS>(sw + GN_OFF(PE_W2), dpe, de1, scr);
#pragma unroll
            for (int c = 0; c < 32; ++c) de1[c] = e1[c] > 0.f ? de1[c] : 0.f;
            dw_layer<34, 32, 32>(gw + GN_OFF(PE_W0), gw + GN_OFF(PE_B0), xin, de1, sX, sZ, KB_THREADS);
            float dxin[34];
            mv_bwd_rolled<34, 32, false, KB_THREADS>(sw + GN_OFF(PE_W0), de1, dxin, scr);
#pragma unroll
            for (int c = 0; c < 32; ++c) dray[c] = dxin[c];
            dhit = 2.f * dxin[32] * mask; dvis = 2.f * dxin[33] * mask;       // hv = (x - 0.5) * 2 ; hit, vis carry the mask
        }
        // compute_prob (dist_decoder.py:109-142): c = 0.5 + 0.5 tanh(z), dc/dz = 2 c (1 - c)
        float dom[4], dov[4], doa[4];
        {
            const float dc00 = -aw * (dvis + dhit), dc01 = -mix1 * (dvis + dhit);
            const float dc10 = aw * dhit, dc11 = mix1 * dhit;
            const float daw = ((1.f - c00) - (1.f - c01)) * dvis + ((c10 - c00) - (c11 - c01)) * dhit;
            const float dz00 = dc00 * 2.f * c00 * (1.f - c00), dz10 = dc10 * 2.f * c10 * (1.f - c10);
            const float dz01 = dc01 * 2.f * c01 * (1.f - c01), dz11 = dc11 * 2.f * c11 * (1.f - c11);
            const float dmean0 = -var0 * (dz00 + dz10), dmean1 = -var1 * (dz01 + dz11);
            const float dvar0 = (nearp - mean0) * dz00 + (farp - mean0) * dz10;
            const float dvar1 = (nearp - mean1) * dz01 + (farp - mean1) * dz11;
            // softplus'(u) = sigmoid(u) (threshold 20: identity above)
            dom[0] = dmean0 * (om[0] > 20.f ? 1.f : kb_sigmoid(om[0])); dom[1] = dmean1 * (om[1] > 20.f ? 1.f : kb_sigmoid(om[1]));
            dov[0] = dvar0 * (ov[0] > 20.f ? 1.f : kb_sigmoid(ov[0])); dov[1] = dvar1 * (ov[1] > 20.f ? 1.f : kb_sigmoid(ov[1]));
            doa[0] = daw * aw * (1.f - aw);
            dom[2] = dom[3] = dov[2] = dov[3] = doa[1] = doa[2] = doa[3] = 0.f;
        }
        dd_bwd<DD_IDS(MEAN)>(sw, gw, ray, h1m, h2m, dom, dray, sX, sZ, scr);
        dd_bwd<DD_IDS(VAR)>(sw, gw, ray, h1v, h2v, dov, dray, sX, sZ, scr);
        dd_bwd<DD_IDS(AW)>(sw, gw, ray, h1a, h2a, doa, dray, sX, sZ, scr);
        if (valid) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) st4(drow + c, make_float4(dray[c], dray[c + 1], dray[c + 2], dray[c + 3]));
        }
    }
}

extern "C" int gn_k2a_backward(const GnK2aBwdParams* hp, void* stream)
{
    const GnK2aBwdParams& p = *hp;
    if (p.V < 1 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (!p.rec || !p.pt || !p.weights || !p.depth_range || !p.d_pooled || !p.d_rec || !p.d_weights) return -2;
    if (p.que_dists && (p.dn < 1 || (p.N % p.dn) != 0)) return -4;
    const int G = 32 / p.V;
    const long long total = (long long)p.B * p.N;
    const long long per_tile = (long long)KB_WARPS * G;
    const long long tiles = (total + per_tile - 1) / per_tile;
    if (tiles > 0x7fffffffLL) return -6;
    const size_t smem = ((size_t)GN_W_K2A_FLOATS + (size_t)KB_THREADS * (GN_BWD_LDX + GN_BWD_LDZ)) * sizeof(float);
    if (smem > 227 * 1024) return -5;
    static size_t cache[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k2a_backward_kernel, smem, cache);
    if (e != cudaSuccess) return (int)e;
    const int sms = gn_sm_count();
    const int grid = (int)(tiles < sms ? tiles : sms);
    gn_k2a_backward_kernel<<<grid, KB_THREADS, smem, (cudaStream_t)stream>>>(p, (int)tiles, G);
    return (int)cudaGetLastError();
}
