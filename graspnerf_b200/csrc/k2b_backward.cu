// K2b backward: reverse pass of the per-ray geometry head (ibrnet.py:485-495) for training.
//
// Forward (recomputed here from the saved `pooled`):  in = [pooled 65 | embed(pts) 21] -> geometry_fc (ELU, ELU) -> + pos
//   -> 4-head attention over the ray's samples (query-row mask) -> fc + residual -> LayerNorm -> out_geometry_fc (2 Linears)
//   -> clip(-1,1) -> invalid (nvalid < 1) := 1.
// Reverse: d sdf (volume layout in volume mode) -> d pooled[0..64] and the gradients of every weight of the chain
//   (gf.*, at.*, og.*), accumulated into the blob-layout buffer with coalesced atomics.
// The cotangent chain through attention / LayerNorm is the one gn_k2b_kernel<true> uses for d sdf / d pts
// (k2b_ray_head.cu), seeded with the upstream gradient instead of ones.  One thread per sample, CTA = floor(128/dn) rays.
#include "gn_bwd.cuh"
#include "gn_weights.cuh"
#include "../../include/graspnerf_b200.h"

#define K2B_THREADS 128
#define WB(id) (GN_OFF(id) - GN_W_K2B_OFF)

__global__ void __launch_bounds__(K2B_THREADS, 1)
gn_k2b_backward_kernel(const __grid_constant__ GnK2bBwdParams p, int rpb)
{
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                                   // [GN_W_K2B_FLOATS]
    float* sK = sw + GN_W_K2B_FLOATS;                   // [128][16]
    float* sV = sK + K2B_THREADS * 16;
    float* sQ = sV + K2B_THREADS * 16;                  // q / temperature
    float* sDO = sQ + K2B_THREADS * 16;                 // d(o)
    float* sMZD = sDO + K2B_THREADS * 16;               // [128][12] per head: max, Z (<=0: masked row), D
    float* sG1 = sMZD + K2B_THREADS * 12;               // [128][65] geometry_fc hidden (post-ELU)
    float* sX = sG1 + K2B_THREADS * 65;                 // [128][GN_BWD_LDX] staging
    float* sZ = sX + K2B_THREADS * GN_BWD_LDX;          // [128][GN_BWD_LDZ]

    for (int i = threadIdx.x * 4; i < GN_W_K2B_FLOATS; i += K2B_THREADS * 4)
        *reinterpret_cast<float4*>(sw + i) = ldg4(p.weights + GN_W_K2B_OFF + i);
    double* gw = p.d_weights + GN_W_K2B_OFF;             // gradient blob, same offsets as sw

    const int t = threadIdx.x;
    const int dn = p.dn;
    const int rn = p.N / dn;
    const int rl = t / dn, d = t - rl * dn;
    const long long ray = (long long)blockIdx.x * rpb + rl;
    const bool valid = (rl < rpb) && (ray < (long long)p.B * rn);
    const long long rayc = valid ? ray : 0;
    const int b = (int)(rayc / rn), r = (int)(rayc - (long long)b * rn);
    const size_t pidx = (size_t)b * p.N + (size_t)r * dn + d;
    const int t0 = (rl < rpb) ? rl * dn : 0;
    __syncthreads();

    // ================================ forward (recompute) ==========================================================
    float in[88];
    float nvalid;
    {
        const float* pp = p.pooled + pidx * GN_POOL_STRIDE;
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
            const float4 q = ldg4(pp + c);
            in[c] = q.x; in[c + 1] = q.y; in[c + 2] = q.z; in[c + 3] = q.w;
        }
        const float4 q = ldg4(pp + 64);
        in[64] = q.x; nvalid = q.y;
    }
    {
        float px, py, pz;
        if (p.volume_mode) {
            const int R = p.R;
            const int i = r / R, j = r - i * R, k = R - 1 - d;
            px = __fadd_rn(__ldg(p.axis + i), __ldg(p.bbox_min + b * 3 + 0));
            py = __fadd_rn(__ldg(p.axis + j), __ldg(p.bbox_min + b * 3 + 1));
            pz = __fadd_rn(__ldg(p.axis + k), __ldg(p.bbox_min + b * 3 + 2));
        } else {
            const float* q = p.pts + pidx * 3;
            px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
        }
        const float pv[3] = { px, py, pz };
        in[65] = px; in[66] = py; in[67] = pz;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float fr = (float)(1 << q);
#pragma unroll
            for (int a = 0; a < 3; ++a) sincosf(pv[a] * fr, &in[68 + 6 * q + a], &in[68 + 6 * q + 3 + a]);
        }
        in[86] = 0.f; in[87] = 0.f;
    }
    float tok[16], g2[16];
    {
        float g1[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) g1[c] = sw[WB(GF_B0) + c];
#pragma unroll
        for (int k = 0; k < 86; ++k) {
            const float xk = in[k];
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const float4 w = *reinterpret_cast<const float4*>(sw + WB(GF_W0) + k * 64 + c);
                g1[c] = fmaf(xk, w.x, g1[c]); g1[c + 1] = fmaf(xk, w.y, g1[c + 1]);
                g1[c + 2] = fmaf(xk, w.z, g1[c + 2]); g1[c + 3] = fmaf(xk, w.w, g1[c + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < 64; ++c) { g1[c] = gn_elu(g1[c]); sG1[t * 65 + c] = g1[c]; }
#pragma unroll
        for (int c = 0; c < 16; ++c) tok[c] = sw[WB(GF_B2) + c];
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const float xk = g1[k];
#pragma unroll
            for (int c = 0; c < 16; ++c) tok[c] = fmaf(xk, sw[WB(GF_W2) + k * 16 + c], tok[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) { g2[c] = gn_elu(tok[c]); tok[c] = g2[c] + __ldg(p.pos_table + d * 16 + c); }

    float q[16], kk[16], vv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { q[c] = 0.f; kk[c] = 0.f; vv[c] = 0.f; }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float xk = tok[k];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            q[c] = fmaf(xk, sw[WB(AT_WQ) + k * 16 + c], q[c]);
            kk[c] = fmaf(xk, sw[WB(AT_WK) + k * 16 + c], kk[c]);
            vv[c] = fmaf(xk, sw[WB(AT_WV) + k * 16 + c], vv[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) { q[c] = q[c] / 2.0f; sK[t * 16 + c] = kk[c]; sV[t * 16 + c] = vv[c]; }
    __syncthreads();

    const bool qmask = nvalid > 1.f;
    float o[16], mx[4], zs[4];
    const float puni = 1.0f / (float)dn;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (!qmask) {
            for (int jj = 0; jj < dn; ++jj) {
                const float4 vj = *reinterpret_cast<const float4*>(sV + (t0 + jj) * 16 + 4 * h);
                a0 = fmaf(puni, vj.x, a0); a1 = fmaf(puni, vj.y, a1); a2 = fmaf(puni, vj.z, a2); a3 = fmaf(puni, vj.w, a3);
            }
            mx[h] = 0.f; zs[h] = -1.f;
        } else {
            float m = -INFINITY;
            for (int jj = 0; jj < dn; ++jj) {
                const float4 kj = *reinterpret_cast<const float4*>(sK + (t0 + jj) * 16 + 4 * h);
                m = fmaxf(m, fmaf(q[4 * h + 3], kj.w, fmaf(q[4 * h + 2], kj.z, fmaf(q[4 * h + 1], kj.y, q[4 * h] * kj.x))));
            }
            float z = 0.f;
            for (int jj = 0; jj < dn; ++jj) {
                const float4 kj = *reinterpret_cast<const float4*>(sK + (t0 + jj) * 16 + 4 * h);
                const float4 vj = *reinterpret_cast<const float4*>(sV + (t0 + jj) * 16 + 4 * h);
                const float s = fmaf(q[4 * h + 3], kj.w, fmaf(q[4 * h + 2], kj.z, fmaf(q[4 * h + 1], kj.y, q[4 * h] * kj.x)));
                const float e = __expf(s - m);
                z += e;
                a0 = fmaf(e, vj.x, a0); a1 = fmaf(e, vj.y, a1); a2 = fmaf(e, vj.z, a2); a3 = fmaf(e, vj.w, a3);
            }
            const float iz = 1.f / z;
            a0 *= iz; a1 *= iz; a2 *= iz; a3 *= iz;
            mx[h] = m; zs[h] = z;
        }
        o[4 * h] = a0; o[4 * h + 1] = a1; o[4 * h + 2] = a2; o[4 * h + 3] = a3;
    }
    float xh[16], ln[16], z16[16], rstd;
    {
        float a[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) a[c] = tok[c];
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = fmaf(o[k], sw[WB(AT_FC) + k * 16 + c], a[c]);
        float mu = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) mu += a[c];
        mu *= (1.f / 16.f);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) { const float dl = a[c] - mu; var = fmaf(dl, dl, var); }
        var *= (1.f / 16.f);
        rstd = rsqrtf(var + 1e-6f);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            xh[c] = (a[c] - mu) * rstd;
            ln[c] = fmaf(xh[c], sw[WB(AT_LNW) + c], sw[WB(AT_LNB) + c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) z16[c] = sw[WB(OG_B0) + c];
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int c = 0; c < 16; ++c) z16[c] = fmaf(ln[k], sw[WB(OG_W0) + k * 16 + c], z16[c]);
    float s = sw[WB(OG_B1)];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(z16[c], sw[WB(OG_W1) + c], s);

    // ================================ reverse ======================================================================
    float ds = 0.f;
    if (valid && nvalid >= 1.f && s >= -1.f && s <= 1.f) {                // clip passes the gradient on [-1,1]; masked fill kills it
        const size_t oi = p.volume_mode ? ((size_t)b * p.N + (size_t)r * p.R + (p.R - 1 - d)) : pidx;   // renderer.py:195-198 flip
        ds = __ldg(p.d_sdf + oi);
    }
    const float one = 1.f;
    // out_geometry_fc.1: s = b1 + z16 . w1
    dw_layer<1, 16, 16>(gw + WB(OG_W1), nullptr, &ds, z16, sX, sZ, K2B_THREADS);
    dw_layer<1, 1, 4>(gw + WB(OG_B1), nullptr, &one, &ds, sX, sZ, K2B_THREADS);
    float dz16[16], dln[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) dz16[c] = sw[WB(OG_W1) + c] * ds;
    dw_layer<16, 16, 16>(gw + WB(OG_W0), gw + WB(OG_B0), ln, dz16, sX, sZ, K2B_THREADS);
    mv_bwd<16, 16, 16, false>(sw + WB(OG_W0), dz16, dln);
    // LayerNorm: ln = xh * gamma + beta
    {
        float dgam[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) dgam[c] = dln[c] * xh[c];
        dw_layer<1, 16, 16>(gw + WB(AT_LNW), nullptr, &one, dgam, sX, sZ, K2B_THREADS);
        dw_layer<1, 16, 16>(gw + WB(AT_LNB), nullptr, &one, dln, sX, sZ, K2B_THREADS);
    }
    float da[16], dO[16], dtok[16];
    {
        float m1 = 0.f, m2 = 0.f, dxh[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { dxh[i] = dln[i] * sw[WB(AT_LNW) + i]; m1 += dxh[i]; m2 = fmaf(dxh[i], xh[i], m2); }
        m1 *= (1.f / 16.f); m2 *= (1.f / 16.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) { da[i] = rstd * (dxh[i] - m1 - xh[i] * m2); dtok[i] = da[i]; }
    }
    dw_layer<16, 16, 16>(gw + WB(AT_FC), nullptr, o, da, sX, sZ, K2B_THREADS);
    mv_bwd<16, 16, 16, false>(sw + WB(AT_FC), da, dO);
#pragma unroll
    for (int c = 0; c < 16; ++c) { sQ[t * 16 + c] = q[c]; sDO[t * 16 + c] = dO[c]; }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        float D = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) D = fmaf(dO[4 * h + c], o[4 * h + c], D);
        sMZD[t * 12 + 3 * h] = mx[h]; sMZD[t * 12 + 3 * h + 1] = zs[h]; sMZD[t * 12 + 3 * h + 2] = D;
    }
    __syncthreads();
    float dq[16], dk[16], dv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { dq[c] = 0.f; dk[c] = 0.f; dv[c] = 0.f; }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        if (qmask) {                                     // this thread as query i
            const float iz = 1.f / zs[h];
            const float D = sMZD[t * 12 + 3 * h + 2];
            for (int jj = 0; jj < dn; ++jj) {
                const float4 kj = *reinterpret_cast<const float4*>(sK + (t0 + jj) * 16 + 4 * h);
                const float4 vj = *reinterpret_cast<const float4*>(sV + (t0 + jj) * 16 + 4 * h);
                const float sc = fmaf(q[4 * h + 3], kj.w, fmaf(q[4 * h + 2], kj.z, fmaf(q[4 * h + 1], kj.y, q[4 * h] * kj.x)));
                const float P = __expf(sc - mx[h]) * iz;
                const float dP = fmaf(dO[4 * h + 3], vj.w, fmaf(dO[4 * h + 2], vj.z, fmaf(dO[4 * h + 1], vj.y, dO[4 * h] * vj.x)));
                const float dS = P * (dP - D);
                dq[4 * h] = fmaf(dS, kj.x, dq[4 * h]); dq[4 * h + 1] = fmaf(dS, kj.y, dq[4 * h + 1]);
                dq[4 * h + 2] = fmaf(dS, kj.z, dq[4 * h + 2]); dq[4 * h + 3] = fmaf(dS, kj.w, dq[4 * h + 3]);
            }
        }
        for (int ii = 0; ii < dn; ++ii) {                // this thread as key/value j over the queries i of its ray
            const int ti = t0 + ii;
            const float4 qi = *reinterpret_cast<const float4*>(sQ + ti * 16 + 4 * h);
            const float4 doi = *reinterpret_cast<const float4*>(sDO + ti * 16 + 4 * h);
            const float mi = sMZD[ti * 12 + 3 * h], zi = sMZD[ti * 12 + 3 * h + 1], Di = sMZD[ti * 12 + 3 * h + 2];
            float P, dS;
            if (zi <= 0.f) { P = puni; dS = 0.f; }
            else {
                const float sc = fmaf(qi.w, kk[4 * h + 3], fmaf(qi.z, kk[4 * h + 2], fmaf(qi.y, kk[4 * h + 1], qi.x * kk[4 * h])));
                P = __expf(sc - mi) / zi;
                const float dP = fmaf(doi.w, vv[4 * h + 3], fmaf(doi.z, vv[4 * h + 2], fmaf(doi.y, vv[4 * h + 1], doi.x * vv[4 * h])));
                dS = P * (dP - Di);
            }
            dk[4 * h] = fmaf(dS, qi.x, dk[4 * h]); dk[4 * h + 1] = fmaf(dS, qi.y, dk[4 * h + 1]);
            dk[4 * h + 2] = fmaf(dS, qi.z, dk[4 * h + 2]); dk[4 * h + 3] = fmaf(dS, qi.w, dk[4 * h + 3]);
            dv[4 * h] = fmaf(P, doi.x, dv[4 * h]); dv[4 * h + 1] = fmaf(P, doi.y, dv[4 * h + 1]);
            dv[4 * h + 2] = fmaf(P, doi.z, dv[4 * h + 2]); dv[4 * h + 3] = fmaf(P, doi.w, dv[4 * h + 3]);
        }
    }
    const float live = valid ? 1.f : 0.f;                // idle threads alias ray 0: they must not contribute
#pragma unroll
    for (int c = 0; c < 16; ++c) { dq[c] *= 0.5f * live; dk[c] *= live; dv[c] *= live; }      // q = (W_q tok) / 2
    dw_layer<16, 16, 16>(gw + WB(AT_WQ), nullptr, tok, dq, sX, sZ, K2B_THREADS);
    dw_layer<16, 16, 16>(gw + WB(AT_WK), nullptr, tok, dk, sX, sZ, K2B_THREADS);
    dw_layer<16, 16, 16>(gw + WB(AT_WV), nullptr, tok, dv, sX, sZ, K2B_THREADS);
    mv_bwd<16, 16, 16, true>(sw + WB(AT_WQ), dq, dtok);
    mv_bwd<16, 16, 16, true>(sw + WB(AT_WK), dk, dtok);
    mv_bwd<16, 16, 16, true>(sw + WB(AT_WV), dv, dtok);
    // tok = elu(u2) + pos ; u2 = b2 + W2 g1 ; g1 = elu(u1) ; u1 = b0 + W0 in
    float du2[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) du2[m] = dtok[m] * gn_delu(g2[m]);
    float du1[64];
    {
        float g1[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) g1[c] = sG1[t * 65 + c];
        dw_layer<64, 16, 16>(gw + WB(GF_W2), gw + WB(GF_B2), g1, du2, sX, sZ, K2B_THREADS);
        mv_bwd<64, 16, 16, false>(sw + WB(GF_W2), du2, du1);
#pragma unroll
        for (int c = 0; c < 64; ++c) du1[c] *= gn_delu(g1[c]);
    }
    dw_layer<86, 64, 64>(gw + WB(GF_W0), gw + WB(GF_B0), in, du1, sX, sZ, K2B_THREADS);
    if (valid) {
        float* dp = p.d_pooled + pidx * GN_POOL_STRIDE;
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                a.x = fmaf(sw[WB(GF_W0) + (k + 0) * 64 + i], du1[i], a.x); a.y = fmaf(sw[WB(GF_W0) + (k + 1) * 64 + i], du1[i], a.y);
                a.z = fmaf(sw[WB(GF_W0) + (k + 2) * 64 + i], du1[i], a.z); a.w = fmaf(sw[WB(GF_W0) + (k + 3) * 64 + i], du1[i], a.w);
            }
            st4(dp + k, a);
        }
        float a64 = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) a64 = fmaf(sw[WB(GF_W0) + 64 * 64 + i], du1[i], a64);
        st4(dp + 64, make_float4(a64, 0.f, 0.f, 0.f));
    }
}

extern "C" int gn_k2b_backward(const GnK2bBwdParams* hp, void* stream)
{
    const GnK2bBwdParams& p = *hp;
    if (p.B < 1 || p.N < 1 || p.dn < 1 || p.dn > K2B_THREADS || (p.N % p.dn) != 0) return -1;
    if (!p.pooled || !p.weights || !p.pos_table || !p.d_sdf || !p.d_pooled || !p.d_weights) return -2;
    if (p.volume_mode) { if (!p.axis || !p.bbox_min || p.R != p.dn || p.N != p.R * p.R * p.R) return -3; }
    else if (!p.pts) return -4;
    const int rpb = K2B_THREADS / p.dn;
    const long long rays = (long long)p.B * (p.N / p.dn);
    const long long grid = (rays + rpb - 1) / rpb;
    if (grid > 0x7fffffffLL) return -6;
    const size_t smem = ((size_t)GN_W_K2B_FLOATS + (size_t)K2B_THREADS * (16 * 4 + 12 + 65 + GN_BWD_LDX + GN_BWD_LDZ)) * sizeof(float);
    if (smem > 227 * 1024) return -5;
    static size_t cache[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k2b_backward_kernel, smem, cache);
    if (e != cudaSuccess) return (int)e;
    gn_k2b_backward_kernel<<<(unsigned)grid, K2B_THREADS, smem, (cudaStream_t)stream>>>(p, rpb);
    return (int)cudaGetLastError();
}
