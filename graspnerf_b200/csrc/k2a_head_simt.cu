// K2a (fp32 SIMT version): per-(point,view) head + cross-view pooling.
//
// One thread owns one (point, view) row and walks the whole MLP chain in registers; the weights live in shared
// memory in k-major order so every lane of a warp reads the SAME float4 (broadcast, conflict free).  The V views
// of a point sit in adjacent lanes, so every cross-view reduction (the three weighted mean/variance poolings of
// ibrnet.py:470-471,484 and the softmax of 509-510) is a short warp-shuffle loop.
//
// Reference op chain replaced (per row):
//   ray_dir_fc + add, mean1/var1            ibrnet.py:457-459,471
//   dist_decoder MLPs + compute_prob       dist_decoder.py:99-142, 6-51   (via renderer.py:62-78)
//   prob_embed                             aggregate_net.py:47-54
//   neuray_fc, weight0, mean0/var0         ibrnet.py:469-470
//   base_fc on cat[globalfeat, f, prob]    ibrnet.py:472-475   (140 view-invariant columns computed once per point)
//   vis_fc, vis_fc2, re-weighting          ibrnet.py:477-482
//   final weighted mean/var                ibrnet.py:484 (+ weight.mean, 487)
//   rgb_fc + masked softmax blend          ibrnet.py:507-511   (only when `colors` is requested)
#include "gn_common.cuh"
#include "gn_weights.cuh"
#include "../../include/graspnerf_b200.h"

#define K2A_THREADS 256
#define K2A_WARPS (K2A_THREADS / 32)

// y[0..NP) += sum_k x[k] * W[k][0..NP)      (W in shared memory, row stride NP, NP % 4 == 0)
template <int K, int NP>
__device__ __forceinline__ void mv_acc(const float* __restrict__ W, const float* x, float* y)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float xk = x[k];
#pragma unroll
        for (int n = 0; n < NP; n += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + k * NP + n);
            y[n + 0] = fmaf(xk, w.x, y[n + 0]); y[n + 1] = fmaf(xk, w.y, y[n + 1]);
            y[n + 2] = fmaf(xk, w.z, y[n + 2]); y[n + 3] = fmaf(xk, w.w, y[n + 3]);
        }
    }
}
template <int NP>
__device__ __forceinline__ void load_bias(const float* __restrict__ b, float* y)
{
#pragma unroll
    for (int n = 0; n < NP; n += 4) {
        const float4 w = *reinterpret_cast<const float4*>(b + n);
        y[n] = w.x; y[n + 1] = w.y; y[n + 2] = w.z; y[n + 3] = w.w;
    }
}
template <int N>
__device__ __forceinline__ void elu_inplace(float* y)
{
#pragma unroll
    for (int n = 0; n < N; ++n) y[n] = gn_elu(y[n]);
}
// dot of a register vector with one k-major "row vector" entry (rows = 1)
template <int K>
__device__ __forceinline__ float dot_row(const float* __restrict__ w, const float* x, float acc)
{
#pragma unroll
    for (int k = 0; k < K; ++k) acc = fmaf(w[k], x[k], acc);
    return acc;
}

// 32 -> 32 -> 32 -> NO decoder (dist_decoder.py:62-86): returns pre-activation outputs in o[0..3]
__device__ __forceinline__ void dist_mlp(const float* __restrict__ sw, int w0, int b0, int w2, int b2, int w4, int b4,
                                         const float* ray, float* o)
{
    float h1[32], h2[32];
    load_bias<32>(sw + b0, h1);
    mv_acc<32, 32>(sw + w0, ray, h1);
    elu_inplace<32>(h1);
    load_bias<32>(sw + b2, h2);
    mv_acc<32, 32>(sw + w2, h1, h2);
    elu_inplace<32>(h2);
    load_bias<4>(sw + b4, o);
    mv_acc<32, 4>(sw + w4, h2, o);
}

__global__ void __launch_bounds__(K2A_THREADS, 1)
gn_k2a_simt_kernel(const __grid_constant__ GnK2aParams p, int num_tiles, int G)
{
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                                  // weights [GN_W_K2A_FLOATS]
    float* s_ypart = smem + GN_W_K2A_FLOATS;           // [K2A_WARPS][G][64]

    for (int i = threadIdx.x * 4; i < GN_W_K2A_FLOATS; i += K2A_THREADS * 4)
        *reinterpret_cast<float4*>(sw + i) = ldg4(p.weights + i);
    __syncthreads();

    const int V = p.V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lane_active = lane < G * V;
    const int g = lane_active ? lane / V : 0;
    const int v = lane_active ? lane - g * V : 0;
    const int gb = g * V;
    const long long total_pts = (long long)p.B * p.N;
    float* ypart = s_ypart + (warp * G + g) * 64;
    const unsigned FULL = 0xffffffffu;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        long long pidx = ((long long)tile * K2A_WARPS + warp) * G + g;
        const bool valid = lane_active && pidx < total_pts;
        pidx = pidx < total_pts ? pidx : total_pts - 1;
        const int b = (int)(pidx / p.N);
        const int n = (int)(pidx - (long long)b * p.N);
        const float* row = p.rec + ((size_t)pidx * V + v) * GN_REC_STRIDE;
        const float2 ptv = __ldg(reinterpret_cast<const float2*>(p.pt + (size_t)pidx * GN_PT_STRIDE));

        const float4 tail = ldg4(row + GN_REC_RGB);            // rgb0..2 (masked), depth
        const float mask = (valid && ((__float_as_uint(ptv.y) >> v) & 1u)) ? 1.f : 0.f;
        const float depth = tail.w;
        const float nvalid = ptv.x;
        const float wgt = __fdiv_rn(mask, nvalid + 1e-8f);      // ibrnet.py:466

        float pe[32];
        float hit, vis;
        {
            float ray[32];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = ldg4(row + GN_REC_RAYF + c);
                ray[c] = t.x; ray[c + 1] = t.y; ray[c + 2] = t.z; ray[c + 3] = t.w;
            }
            // ---- dist decoder (dist_decoder.py:99-107, use_vis False)
            float om[4], ov[4], oa[4];
            dist_mlp(sw, GN_OFF(DD_MEAN_W0), GN_OFF(DD_MEAN_B0), GN_OFF(DD_MEAN_W2), GN_OFF(DD_MEAN_B2), GN_OFF(DD_MEAN_W4), GN_OFF(DD_MEAN_B4), ray, om);
            dist_mlp(sw, GN_OFF(DD_VAR_W0), GN_OFF(DD_VAR_B0), GN_OFF(DD_VAR_W2), GN_OFF(DD_VAR_B2), GN_OFF(DD_VAR_W4), GN_OFF(DD_VAR_B4), ray, ov);
            dist_mlp(sw, GN_OFF(DD_AW_W0), GN_OFF(DD_AW_B0), GN_OFF(DD_AW_W2), GN_OFF(DD_AW_B2), GN_OFF(DD_AW_W4), GN_OFF(DD_AW_B4), ray, oa);
            const float mean0 = gn_softplus(om[0]), mean1 = gn_softplus(om[1]);
            const float var0 = gn_softplus(ov[0]) + 0.05f, var1 = gn_softplus(ov[1]) + 0.05f;   // AddBias(0.05)
            const float aw = gn_sigmoid(oa[0]);
            // ---- compute_prob (dist_decoder.py:109-142), is_ref branch of get_near_far_points (17-24)
            const float* dr = p.depth_range + ((size_t)b * V + v) * 2;
            const float rnear = __fdiv_rn(-1.f, __ldg(dr)), rfar = __fdiv_rn(-1.f, __ldg(dr + 1));
            float d = __fdiv_rn(-1.f, fmaxf(depth, 1e-5f));
            d = __fdiv_rn(d - rnear, rfar - rnear);
            float nearp, farp;
            if (p.que_dists == nullptr) {              // fixed interval 0.01 (dist_decoder.py:47-49,121-124)
                nearp = d - 0.005f; farp = d + 0.005f;
            } else {                                    // half intervals (dist_decoder.py:33-38)
                const int smp = n % p.dn;
                const float* qd = p.que_dists + (size_t)b * p.N + n;
                const float h_cur = __ldg(qd) * 0.5f;
                const float h_prev = smp > 0 ? __ldg(qd - 1) * 0.5f : h_cur;
                nearp = d - h_prev; farp = d + h_cur;
            }
            const float c00 = 0.5f + 0.5f * tanhf((nearp - mean0) * var0), c10 = 0.5f + 0.5f * tanhf((farp - mean0) * var0);
            const float c01 = 0.5f + 0.5f * tanhf((nearp - mean1) * var1), c11 = 0.5f + 0.5f * tanhf((farp - mean1) * var1);
            const float mix1 = 1.f - aw;
            vis = ((1.f - c00) * aw + (1.f - c01) * mix1) * mask;          // renderer.py:76
            hit = ((c10 - c00) * aw + (c11 - c01) * mix1) * mask;          // renderer.py:77
            // ---- prob_embed (aggregate_net.py:47-54)
            float e1[32];
            load_bias<32>(sw + GN_OFF(PE_B0), e1);
            mv_acc<32, 32>(sw + GN_OFF(PE_W0), ray, e1);
            const float hv[2] = { (hit - 0.5f) * 2.f, (vis - 0.5f) * 2.f };
            mv_acc<2, 32>(sw + GN_OFF(PE_W0) + 32 * 32, hv, e1);
#pragma unroll
            for (int c = 0; c < 32; ++c) e1[c] = fmaxf(e1[c], 0.f);
            load_bias<32>(sw + GN_OFF(PE_B2), pe);
            mv_acc<32, 32>(sw + GN_OFF(PE_W2), e1, pe);
        }
        // ---- neuray_fc -> weight0 (ibrnet.py:469)
        float w0;
        {
            float t[8];
            load_bias<8>(sw + GN_OFF(NF_B0), t);
            mv_acc<32, 8>(sw + GN_OFF(NF_W0), pe, t);
            elu_inplace<8>(t);
            const float s = dot_row<8>(sw + GN_OFF(NF_W2), t, sw[GN_OFF(NF_B2)]);
            w0 = gn_sigmoid(s) * wgt;
        }
        // ---- ray_dir_fc (ibrnet.py:457) and f = [img_feats | rgb] + direction feature (ibrnet.py:459), record order
        const float4 ddv = ldg4(row + GN_REC_DD);
        float f[36];
        {
            const float dd[4] = { ddv.x, ddv.y, ddv.z, ddv.w };
            float hid[16];
            load_bias<16>(sw + GN_OFF(RD_B0), hid);
            mv_acc<4, 16>(sw + GN_OFF(RD_W0), dd, hid);
            elu_inplace<16>(hid);
            load_bias<36>(sw + GN_OFF(RD_B1), f);
            mv_acc<16, 36>(sw + GN_OFF(RD_W1), hid, f);
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = ldg4(row + GN_REC_IMGF + c);
                f[c] = gn_elu(f[c]) + t.x; f[c + 1] = gn_elu(f[c + 1]) + t.y;
                f[c + 2] = gn_elu(f[c + 2]) + t.z; f[c + 3] = gn_elu(f[c + 3]) + t.w;
            }
            f[32] = gn_elu(f[32]) + tail.x; f[33] = gn_elu(f[33]) + tail.y; f[34] = gn_elu(f[34]) + tail.z; f[35] = 0.f;
        }

        float y[64];
        {
            // view-invariant part of base_fc.0: 144 inputs [mean0|var0|mean1|var1] (each 36 wide, pad row = 0 weight).
            // Lane (g,v) computes outputs n = v, v+V, ... and parks them in shared memory for its group.
            float g0[36], g1[36];                          // means first, then overwritten in place by the variances
#pragma unroll
            for (int c = 0; c < 35; ++c) {                 // ibrnet.py:470-471
                const float t0 = w0 * f[c], t1 = wgt * f[c];
                float s0 = 0.f, s1 = 0.f;
                for (int jv = 0; jv < V; ++jv) {
                    s0 += __shfl_sync(FULL, t0, (gb + jv) & 31);
                    s1 += __shfl_sync(FULL, t1, (gb + jv) & 31);
                }
                g0[c] = s0; g1[c] = s1;
            }
            g0[35] = 0.f; g1[35] = 0.f;
            const float* wg = sw + GN_OFF(BF_WG);
            for (int nn = v; nn < 64; nn += V) {            // rows 0..35 (mean0), 72..107 (mean1)
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 36; ++c) acc = fmaf(g0[c], wg[c * 64 + nn], acc);
#pragma unroll
                for (int c = 0; c < 36; ++c) acc = fmaf(g1[c], wg[(72 + c) * 64 + nn], acc);
                if (lane_active) ypart[nn] = acc;
            }
#pragma unroll
            for (int c = 0; c < 35; ++c) {
                const float d0 = f[c] - g0[c], d1 = f[c] - g1[c];
                const float t0 = w0 * d0 * d0, t1 = wgt * d1 * d1;
                float s0 = 0.f, s1 = 0.f;
                for (int jv = 0; jv < V; ++jv) {
                    s0 += __shfl_sync(FULL, t0, (gb + jv) & 31);
                    s1 += __shfl_sync(FULL, t1, (gb + jv) & 31);
                }
                g0[c] = s0; g1[c] = s1;
            }
            for (int nn = v; nn < 64; nn += V) {            // rows 36..71 (var0), 108..143 (var1)
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 36; ++c) acc = fmaf(g0[c], wg[(36 + c) * 64 + nn], acc);
#pragma unroll
                for (int c = 0; c < 36; ++c) acc = fmaf(g1[c], wg[(108 + c) * 64 + nn], acc);
                if (lane_active) ypart[nn] += acc;
            }
            __syncwarp();
            // ---- base_fc (ibrnet.py:472-475)
            load_bias<64>(sw + GN_OFF(BF_B0), y);
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(ypart + c);
                y[c] += t.x; y[c + 1] += t.y; y[c + 2] += t.z; y[c + 3] += t.w;
            }
            __syncwarp();
        }
        mv_acc<36, 64>(sw + GN_OFF(BF_WF), f, y);
        mv_acc<32, 64>(sw + GN_OFF(BF_WP), pe, y);
        elu_inplace<64>(y);
        float x[32];
        load_bias<32>(sw + GN_OFF(BF_B2), x);
        mv_acc<64, 32>(sw + GN_OFF(BF_W2), y, x);
        elu_inplace<32>(x);
        // ---- vis_fc (ibrnet.py:477-480)
        float visw;
        {
            float xi[32], t[32], xv[36];
#pragma unroll
            for (int c = 0; c < 32; ++c) xi[c] = x[c] * wgt;
            load_bias<32>(sw + GN_OFF(VF_B0), t);
            mv_acc<32, 32>(sw + GN_OFF(VF_W0), xi, t);
            elu_inplace<32>(t);
            load_bias<36>(sw + GN_OFF(VF_B2), xv);
            mv_acc<32, 36>(sw + GN_OFF(VF_W2), t, xv);
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] += gn_elu(xv[c]);
            visw = gn_sigmoid(gn_elu(xv[32])) * mask;
        }
        // ---- vis_fc2 (ibrnet.py:481)
        float vis2;
        {
            float xi[32], t[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) xi[c] = x[c] * visw;
            load_bias<32>(sw + GN_OFF(V2_B0), t);
            mv_acc<32, 32>(sw + GN_OFF(V2_W0), xi, t);
            elu_inplace<32>(t);
            const float s = dot_row<32>(sw + GN_OFF(V2_W2), t, sw[GN_OFF(V2_B2)]);
            vis2 = gn_sigmoid(s) * mask;
        }
        // ---- final pooling (ibrnet.py:482-484,487)
        float ssum = 0.f;
        for (int jv = 0; jv < V; ++jv) ssum += __shfl_sync(FULL, vis2, (gb + jv) & 31);
        const float w2 = __fdiv_rn(vis2, ssum + 1e-8f);
        float w2sum = 0.f;
        for (int jv = 0; jv < V; ++jv) w2sum += __shfl_sync(FULL, w2, (gb + jv) & 31);
        float* out = p.pooled + (size_t)pidx * GN_POOL_STRIDE;
        const bool writer = valid && v == 0;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            float mu[4], vr[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float t = w2 * x[c + q];
                float s = 0.f;
                for (int jv = 0; jv < V; ++jv) s += __shfl_sync(FULL, t, (gb + jv) & 31);
                mu[q] = s;
                const float dlt = x[c + q] - s;
                const float t2 = w2 * dlt * dlt;
                float s2 = 0.f;
                for (int jv = 0; jv < V; ++jv) s2 += __shfl_sync(FULL, t2, (gb + jv) & 31);
                vr[q] = s2;
            }
            if (writer) {
                st4(out + c, make_float4(mu[0], mu[1], mu[2], mu[3]));
                st4(out + 32 + c, make_float4(vr[0], vr[1], vr[2], vr[3]));
            }
        }
        if (writer) st4(out + 64, make_float4(w2sum / (float)V, nvalid, 0.f, 0.f));

        if (p.dbg_rows && valid) {
            float* dr = p.dbg_rows + ((size_t)pidx * V + v) * 8;
            st4(dr, make_float4(hit, vis, w0, vis2));
            st4(dr + 4, make_float4(x[0], x[1], pe[0], pe[1]));
        }

        // ---- rgb_fc + masked softmax over views (ibrnet.py:507-511)
        if (p.with_rgb && p.colors) {
            float in[40], r16[16], r8[8];
#pragma unroll
            for (int c = 0; c < 32; ++c) in[c] = x[c];
            in[32] = vis2;
            in[33] = ddv.x; in[34] = ddv.y; in[35] = ddv.z; in[36] = ddv.w;
            load_bias<16>(sw + GN_OFF(RF_B0), r16);
            mv_acc<37, 16>(sw + GN_OFF(RF_W0), in, r16);
            elu_inplace<16>(r16);
            load_bias<8>(sw + GN_OFF(RF_B2), r8);
            mv_acc<16, 8>(sw + GN_OFF(RF_W2), r16, r8);
            elu_inplace<8>(r8);
            float logit = dot_row<8>(sw + GN_OFF(RF_W4), r8, sw[GN_OFF(RF_B4)]);
            if (mask == 0.f) logit = -1e9f;
            float mx = -INFINITY;
            for (int jv = 0; jv < V; ++jv) mx = fmaxf(mx, __shfl_sync(FULL, logit, (gb + jv) & 31));
            const float e = __expf(logit - mx);
            float es = 0.f;
            for (int jv = 0; jv < V; ++jv) es += __shfl_sync(FULL, e, (gb + jv) & 31);
            const float bw = __fdiv_rn(e, es);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f;                // rgb_in = rgb * mask (ibrnet.py:458)
            for (int jv = 0; jv < V; ++jv) {
                c0 += __shfl_sync(FULL, bw * tail.x, (gb + jv) & 31);
                c1 += __shfl_sync(FULL, bw * tail.y, (gb + jv) & 31);
                c2 += __shfl_sync(FULL, bw * tail.z, (gb + jv) & 31);
            }
            if (writer) st4(p.colors + (size_t)pidx * 4, make_float4(c0, c1, c2, 0.f));
        }
    }
}

extern "C" int gn_k2a_forward(const GnK2aParams* hp, void* stream)
{
    const GnK2aParams& p = *hp;
    if (p.V < 1 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (p.que_dists && (p.dn < 1 || (p.N % p.dn) != 0)) return -4;
    const int G = 32 / p.V;
    const long long total = (long long)p.B * p.N;
    const long long per_tile = (long long)K2A_WARPS * G;
    const long long tiles = (total + per_tile - 1) / per_tile;
    if (tiles > 0x7fffffffLL) return -6;
    const size_t smem = ((size_t)GN_W_K2A_FLOATS + (size_t)K2A_WARPS * G * 64) * sizeof(float);
    if (smem > 227 * 1024) return -5;
    static size_t smem_cache_gn_k2a_simt_kernel[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k2a_simt_kernel, smem, smem_cache_gn_k2a_simt_kernel);
    if (e != cudaSuccess) return (int)e;
    const int sms = gn_sm_count();
    const int grid = (int)(tiles < sms ? tiles : sms);
    gn_k2a_simt_kernel<<<grid, K2A_THREADS, smem, (cudaStream_t)stream>>>(p, (int)tiles, G);
    return (int)cudaGetLastError();
}
