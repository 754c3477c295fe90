// K2b: per-ray geometry head.  One thread per sample, one CTA = floor(128/dn) rays.
//
// Replaces ibrnet.py:485-504:
//   embed_fn (neus.py:21-66, multires 3)            -> 21 dims
//   geometry_fc 86 -> 64 -> 16 (ELU)                 ibrnet.py:404-407,487-489
//   + pos_encoding[d]                                ibrnet.py:437-445,491
//   MultiHeadAttention(4 heads, d_k 4) over the dn samples of the ray, query-row mask, residual, LayerNorm(1e-6)
//                                                    ibrnet.py:52-102 (ScaledDotProductAttention 7-27)
//   out_geometry_fc 16 -> 16 -> 1 (no activation), clip(-1,1), invalid -> 1.0     ibrnet.py:410-412,494-495
//   gradients = autograd.grad(sdf, que_pts, ones)    ibrnet.py:497-504  (GRAD=true: hand-derived reverse pass)
// In volume mode the final z flip of renderer.py:198 is fused into the store.
#include "gn_common.cuh"
#include "gn_weights.cuh"
#include "../../include/graspnerf_b200.h"

#define K2B_THREADS 128
#define WB(id) (GN_OFF(id) - GN_W_K2B_OFF)

template <bool GRAD>
__global__ void __launch_bounds__(K2B_THREADS)
gn_k2b_kernel(const __grid_constant__ GnK2bParams p, int rpb)
{
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                                   // [GN_W_K2B_FLOATS]
    float* sK = sw + GN_W_K2B_FLOATS;                   // [128][16]
    float* sV = sK + K2B_THREADS * 16;                  // [128][16]
    // GRAD only:
    float* sQ = sV + K2B_THREADS * 16;                  // [128][16]  q / temperature
    float* sDO = sQ + K2B_THREADS * 16;                 // [128][16]  d(o)
    float* sMZD = sDO + K2B_THREADS * 16;               // [128][12]  per head: max, Z (<=0: masked row), D
    float* sG1 = sMZD + K2B_THREADS * 12;               // [128][65]  geometry_fc hidden (post-ELU)

    for (int i = threadIdx.x * 4; i < GN_W_K2B_FLOATS; i += K2B_THREADS * 4)
        *reinterpret_cast<float4*>(sw + i) = ldg4(p.weights + GN_W_K2B_OFF + i);

    const int t = threadIdx.x;
    const int dn = p.dn;
    const int rn = p.N / dn;
    const int rl = t / dn, d = t - rl * dn;
    const long long ray = (long long)blockIdx.x * rpb + rl;
    const bool valid = (rl < rpb) && (ray < (long long)p.B * rn);
    const long long rayc = valid ? ray : 0;
    const int b = (int)(rayc / rn), r = (int)(rayc - (long long)b * rn);
    const size_t pidx = (size_t)b * p.N + (size_t)r * dn + d;
    const int t0 = (rl < rpb) ? rl * dn : 0;             // first thread of this ray (idle threads alias ray 0)
    __syncthreads();

    // ---- inputs: pooled[65] + embed(pts)[21]
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
    float px, py, pz;
    if (p.volume_mode) {       // same arithmetic as K1 (field_utils.py:17-27 + bbox3d[0]); sample d <-> k = R-1-d
        const int R = p.R;
        const int i = r / R, j = r - i * R, k = R - 1 - d;
        px = __fadd_rn(__ldg(p.axis + i), __ldg(p.bbox_min + b * 3 + 0));
        py = __fadd_rn(__ldg(p.axis + j), __ldg(p.bbox_min + b * 3 + 1));
        pz = __fadd_rn(__ldg(p.axis + k), __ldg(p.bbox_min + b * 3 + 2));
    } else {
        const float* q = p.pts + pidx * 3;
        px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
    }
    float sn[9], cs[9];        // sin/cos(f * p_a), f in {1,2,4}
    {
        const float pv[3] = { px, py, pz };
        in[65] = px; in[66] = py; in[67] = pz;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float fr = (float)(1 << q);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                sincosf(pv[a] * fr, &sn[q * 3 + a], &cs[q * 3 + a]);
                in[68 + 6 * q + a] = sn[q * 3 + a];
                in[68 + 6 * q + 3 + a] = cs[q * 3 + a];
            }
        }
        in[86] = 0.f; in[87] = 0.f;
    }
    // ---- geometry_fc
    float tok[16];
    {
        float g1[64];
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
            const float4 q = *reinterpret_cast<const float4*>(sw + WB(GF_B0) + c);
            g1[c] = q.x; g1[c + 1] = q.y; g1[c + 2] = q.z; g1[c + 3] = q.w;
        }
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
        for (int c = 0; c < 64; ++c) {
            g1[c] = gn_elu(g1[c]);
            if (GRAD) sG1[t * 65 + c] = g1[c];
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            const float4 q = *reinterpret_cast<const float4*>(sw + WB(GF_B2) + c);
            tok[c] = q.x; tok[c + 1] = q.y; tok[c + 2] = q.z; tok[c + 3] = q.w;
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const float xk = g1[k];
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 w = *reinterpret_cast<const float4*>(sw + WB(GF_W2) + k * 16 + c);
                tok[c] = fmaf(xk, w.x, tok[c]); tok[c + 1] = fmaf(xk, w.y, tok[c + 1]);
                tok[c + 2] = fmaf(xk, w.z, tok[c + 2]); tok[c + 3] = fmaf(xk, w.w, tok[c + 3]);
            }
        }
    }
    float g2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { g2[c] = gn_elu(tok[c]); tok[c] = g2[c] + __ldg(p.pos_table + d * 16 + c); }  // ibrnet.py:491

    // ---- q, k, v projections (bias-free, ibrnet.py:62-64)
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
    for (int c = 0; c < 16; ++c) { q[c] = q[c] / 2.0f; sK[t * 16 + c] = kk[c]; sV[t * 16 + c] = vv[c]; }   // temperature d_k^0.5 = 2
    __syncthreads();

    const bool qmask = nvalid > 1.f;                      // ibrnet.py:492-493: mask=(num_valid_obs > 1) on the query row
    float o[16], mx[4], zs[4];
    const float puni = 1.0f / (float)dn;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (!qmask) {                                     // whole row filled with -1e9 -> uniform softmax
            for (int jj = 0; jj < dn; ++jj) {
                const float4 vj = *reinterpret_cast<const float4*>(sV + (t0 + jj) * 16 + 4 * h);
                a0 = fmaf(puni, vj.x, a0); a1 = fmaf(puni, vj.y, a1); a2 = fmaf(puni, vj.z, a2); a3 = fmaf(puni, vj.w, a3);
            }
            mx[h] = 0.f; zs[h] = -1.f;
        } else {
            float m = -INFINITY;
            for (int jj = 0; jj < dn; ++jj) {
                const float4 kj = *reinterpret_cast<const float4*>(sK + (t0 + jj) * 16 + 4 * h);
                const float s = fmaf(q[4 * h + 3], kj.w, fmaf(q[4 * h + 2], kj.z, fmaf(q[4 * h + 1], kj.y, q[4 * h] * kj.x)));
                m = fmaxf(m, s);
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
    // ---- fc + residual + LayerNorm (ibrnet.py:96-100)
    float a[16], xh[16], ln[16];
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
    const float rstd = rsqrtf(var + 1e-6f);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        xh[c] = (a[c] - mu) * rstd;
        ln[c] = fmaf(xh[c], sw[WB(AT_LNW) + c], sw[WB(AT_LNB) + c]);
    }
    // ---- out_geometry_fc (two Linears, no activation), clip, invalid -> 1
    float z16[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) z16[c] = sw[WB(OG_B0) + c];
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int c = 0; c < 16; ++c) z16[c] = fmaf(ln[k], sw[WB(OG_W0) + k * 16 + c], z16[c]);
    float s = sw[WB(OG_B1)];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(z16[c], sw[WB(OG_W1) + c], s);
    float sdf = fminf(fmaxf(s, -1.f), 1.f);
    if (nvalid < 1.f) sdf = 1.f;
    if (valid) {
        if (p.volume_mode) {
            const int R = p.R;                            // volume[b][i][j][k], k = R-1-d   (renderer.py:195-198)
            p.sdf[(size_t)b * p.N + (size_t)r * R + (R - 1 - d)] = sdf;
        } else {
            p.sdf[pidx] = sdf;
        }
    }

    if (GRAD) {
        // ================= reverse pass: cotangent 1 on every sdf of the ray =========================
        const float ds = (valid && nvalid >= 1.f && s >= -1.f && s <= 1.f) ? 1.f : 0.f;
        float dln[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < 16; ++m) acc = fmaf(sw[WB(OG_W0) + i * 16 + m], sw[WB(OG_W1) + m], acc);
            dln[i] = acc * ds;
        }
        float m1 = 0.f, m2 = 0.f, dxh[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { dxh[i] = dln[i] * sw[WB(AT_LNW) + i]; m1 += dxh[i]; m2 = fmaf(dxh[i], xh[i], m2); }
        m1 *= (1.f / 16.f); m2 *= (1.f / 16.f);
        float da[16], dO[16], dtok[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { da[i] = rstd * (dxh[i] - m1 - xh[i] * m2); dtok[i] = da[i]; }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < 16; ++m) acc = fmaf(sw[WB(AT_FC) + i * 16 + m], da[m], acc);
            dO[i] = acc;
        }
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
            // (a) this thread as query i
            if (qmask) {
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
            // (b) this thread as key/value j, looping over the queries i of the ray
            for (int ii = 0; ii < dn; ++ii) {
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
#pragma unroll
        for (int c = 0; c < 16; ++c) dq[c] *= 0.5f;       // q = (W_q tok) / 2
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float acc = dtok[i];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                acc = fmaf(sw[WB(AT_WQ) + i * 16 + m], dq[m], acc);
                acc = fmaf(sw[WB(AT_WK) + i * 16 + m], dk[m], acc);
                acc = fmaf(sw[WB(AT_WV) + i * 16 + m], dv[m], acc);
            }
            dtok[i] = acc;
        }
        // tok = elu(u2) + pos ; u2 = b2 + W2 g1 ; g1 = elu(u1) ; u1 = b0 + W0 [pooled, embed]
        float du2[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) du2[m] = dtok[m] * (g2[m] > 0.f ? 1.f : g2[m] + 1.f);
        float de[21];
#pragma unroll
        for (int c = 0; c < 21; ++c) de[c] = 0.f;
        for (int i = 0; i < 64; ++i) {
            const float g1i = sG1[t * 65 + i];
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < 16; ++m) acc = fmaf(sw[WB(GF_W2) + i * 16 + m], du2[m], acc);
            const float du1 = acc * (g1i > 0.f ? 1.f : g1i + 1.f);
#pragma unroll
            for (int c = 0; c < 21; ++c) de[c] = fmaf(sw[WB(GF_W0) + (65 + c) * 64 + i], du1, de[c]);
        }
        if (valid && p.grad) {
            float dp[3];
#pragma unroll
            for (int a2 = 0; a2 < 3; ++a2) {
                float acc = de[a2];
#pragma unroll
                for (int qf = 0; qf < 3; ++qf) {
                    const float fr = (float)(1 << qf);
                    acc = fmaf(fr * cs[qf * 3 + a2], de[3 + 6 * qf + a2], acc);
                    acc = fmaf(-fr * sn[qf * 3 + a2], de[3 + 6 * qf + 3 + a2], acc);
                }
                dp[a2] = acc;
            }
            float* go = p.grad + pidx * 3;
            go[0] = dp[0]; go[1] = dp[1]; go[2] = dp[2];
        }
    }
}

// Attention-only variant: the geometry_fc output (post-ELU, before the positional table) arrives from K2a's GEMM chain
// as tok[B,N,20]; this kernel does +pos, q/k/v, 4-head attention over the ray's samples, fc + residual, LayerNorm,
// out_geometry_fc, clip and the invalid fill (ibrnet.py:491-495).  No gradient path (volume mode / sdf_only).
#define K2B_AT_OFF (GN_OFF(AT_WQ) - GN_W_K2B_OFF)
#define K2B_AT_FLOATS (GN_W_K2B_FLOATS - K2B_AT_OFF)
#define WA(id) (GN_OFF(id) - GN_OFF(AT_WQ))
__global__ void __launch_bounds__(K2B_THREADS, 4)
gn_k2b_attn_kernel(const __grid_constant__ GnK2bParams p, int rpb)
{
    __shared__ __align__(16) float sw[K2B_AT_FLOATS];
    __shared__ __align__(16) float sK[K2B_THREADS * 16];
    __shared__ __align__(16) float sV[K2B_THREADS * 16];
    for (int i = threadIdx.x * 4; i < K2B_AT_FLOATS; i += K2B_THREADS * 4)
        *reinterpret_cast<float4*>(sw + i) = ldg4(p.weights + GN_OFF(AT_WQ) + i);
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
    float tok[16];
    float nvalid;
    {
        const float* tp = p.tok + pidx * GN_TOK_STRIDE;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            const float4 q = ldg4(tp + c);
            const float4 ps = ldg4(p.pos_table + d * 16 + c);
            tok[c] = q.x + ps.x; tok[c + 1] = q.y + ps.y; tok[c + 2] = q.z + ps.z; tok[c + 3] = q.w + ps.w;   // ibrnet.py:491
        }
        nvalid = __ldg(tp + 16);
    }
    __syncthreads();
    float q[16], kk[16], vv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { q[c] = 0.f; kk[c] = 0.f; vv[c] = 0.f; }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float xk = tok[k];
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            const float4 wq = *reinterpret_cast<const float4*>(sw + WA(AT_WQ) + k * 16 + c);
            const float4 wk = *reinterpret_cast<const float4*>(sw + WA(AT_WK) + k * 16 + c);
            const float4 wv = *reinterpret_cast<const float4*>(sw + WA(AT_WV) + k * 16 + c);
            q[c] = fmaf(xk, wq.x, q[c]); q[c + 1] = fmaf(xk, wq.y, q[c + 1]); q[c + 2] = fmaf(xk, wq.z, q[c + 2]); q[c + 3] = fmaf(xk, wq.w, q[c + 3]);
            kk[c] = fmaf(xk, wk.x, kk[c]); kk[c + 1] = fmaf(xk, wk.y, kk[c + 1]); kk[c + 2] = fmaf(xk, wk.z, kk[c + 2]); kk[c + 3] = fmaf(xk, wk.w, kk[c + 3]);
            vv[c] = fmaf(xk, wv.x, vv[c]); vv[c + 1] = fmaf(xk, wv.y, vv[c + 1]); vv[c + 2] = fmaf(xk, wv.z, vv[c + 2]); vv[c + 3] = fmaf(xk, wv.w, vv[c + 3]);
        }
    }
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
        q[c] = q[c] / 2.0f; q[c + 1] = q[c + 1] / 2.0f; q[c + 2] = q[c + 2] / 2.0f; q[c + 3] = q[c + 3] / 2.0f;     // temperature d_k^0.5 = 2
        st4(sK + t * 16 + c, make_float4(kk[c], kk[c + 1], kk[c + 2], kk[c + 3]));
        st4(sV + t * 16 + c, make_float4(vv[c], vv[c + 1], vv[c + 2], vv[c + 3]));
    }
    __syncthreads();
    const bool qmask = nvalid > 1.f;
    float o[16];
    const float puni = 1.0f / (float)dn;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (!qmask) {
            for (int jj = 0; jj < dn; ++jj) {
                const float4 vj = *reinterpret_cast<const float4*>(sV + (t0 + jj) * 16 + 4 * h);
                a0 = fmaf(puni, vj.x, a0); a1 = fmaf(puni, vj.y, a1); a2 = fmaf(puni, vj.z, a2); a3 = fmaf(puni, vj.w, a3);
            }
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
        }
        o[4 * h] = a0; o[4 * h + 1] = a1; o[4 * h + 2] = a2; o[4 * h + 3] = a3;
    }
    float a[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) a[c] = tok[c];
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sw + WA(AT_FC) + k * 16 + c);
            a[c] = fmaf(o[k], w.x, a[c]); a[c + 1] = fmaf(o[k], w.y, a[c + 1]); a[c + 2] = fmaf(o[k], w.z, a[c + 2]); a[c + 3] = fmaf(o[k], w.w, a[c + 3]);
        }
    float mu = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) mu += a[c];
    mu *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { const float dl = a[c] - mu; var = fmaf(dl, dl, var); }
    var *= (1.f / 16.f);
    const float rstd = rsqrtf(var + 1e-6f);
    float ln[16], z16[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) ln[c] = fmaf((a[c] - mu) * rstd, sw[WA(AT_LNW) + c], sw[WA(AT_LNB) + c]);
#pragma unroll
    for (int c = 0; c < 16; ++c) z16[c] = sw[WA(OG_B0) + c];
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sw + WA(OG_W0) + k * 16 + c);
            z16[c] = fmaf(ln[k], w.x, z16[c]); z16[c + 1] = fmaf(ln[k], w.y, z16[c + 1]); z16[c + 2] = fmaf(ln[k], w.z, z16[c + 2]); z16[c + 3] = fmaf(ln[k], w.w, z16[c + 3]);
        }
    float s = sw[WA(OG_B1)];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(z16[c], sw[WA(OG_W1) + c], s);
    float sdf = fminf(fmaxf(s, -1.f), 1.f);
    if (nvalid < 1.f) sdf = 1.f;
    if (valid) {
        if (p.volume_mode) {
            const int R = p.R;
            p.sdf[(size_t)b * p.N + (size_t)r * R + (R - 1 - d)] = sdf;
        } else {
            p.sdf[pidx] = sdf;
        }
    }
}

extern "C" int gn_k2b_forward(const GnK2bParams* hp, void* stream)
{
    const GnK2bParams& p = *hp;
    if (p.B < 1 || p.N < 1 || p.dn < 1 || p.dn > K2B_THREADS || (p.N % p.dn) != 0) return -1;
    const int rpb = K2B_THREADS / p.dn;
    const long long rays = (long long)p.B * (p.N / p.dn);
    const long long grid = (rays + rpb - 1) / rpb;
    if (grid > 0x7fffffffLL) return -6;
    if (!p.pooled) {                       // attention-only path on K2a's tokens
        if (!p.tok || p.grad) return -2;
        if (p.volume_mode && (p.N != p.R * p.R * p.R || p.dn != p.R)) return -3;
        gn_k2b_attn_kernel<<<(unsigned)grid, K2B_THREADS, 0, (cudaStream_t)stream>>>(p, rpb);
        return (int)cudaGetLastError();
    }
    if (p.volume_mode && (p.N != p.R * p.R * p.R || p.dn != p.R || !p.axis || !p.bbox_min)) return -3;
    if (!p.volume_mode && !p.pts) return -4;
    const bool wg = p.grad != nullptr;
    size_t smem = ((size_t)GN_W_K2B_FLOATS + 2 * K2B_THREADS * 16) * sizeof(float);
    if (wg) smem += (size_t)K2B_THREADS * (16 + 16 + 12 + 65) * sizeof(float);
    cudaError_t e;
    if (wg) {
        static size_t cache_t[16] = {0};
        e = gn_ensure_smem(gn_k2b_kernel<true>, smem, cache_t);
        if (e != cudaSuccess) return (int)e;
        gn_k2b_kernel<true><<<(unsigned)grid, K2B_THREADS, smem, (cudaStream_t)stream>>>(p, rpb);
    } else {
        static size_t cache_f[16] = {0};
        e = gn_ensure_smem(gn_k2b_kernel<false>, smem, cache_f);
        if (e != cudaSuccess) return (int)e;
        gn_k2b_kernel<false><<<(unsigned)grid, K2B_THREADS, smem, (cudaStream_t)stream>>>(p, rpb);
    }
    return (int)cudaGetLastError();
}
