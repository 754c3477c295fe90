// K3: NeuS alpha, alpha compositing and the depth samplers of the RGB head (latency-bound, one thread per ray).
//
//   gn_k3_composite     _get_alpha_from_sdf (aggregate_net.py:105-123), alpha_values2hit_prob (render_ops.py:72-80),
//                       pixel colour / depth sums (renderer.py:105-106,136), eikonal partial sums (aggregate_net.py:139)
//   gn_k3_coarse_depths sample_depth, deterministic branch (render_ops.py:146-170)
//   gn_k3_fine_depths   sample_fine_depth (render_ops.py:172-229) + the sort of renderer.py:148.
// Sums and scans run left to right in fp32 (the order oracle/nr_oracle.py states), so the searchsorted index table is
// bit-exact against the oracle.
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"

__device__ __forceinline__ float k3_sigmoid(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

__global__ void gn_k3_composite_kernel(const GnK3Params p)
{
    const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= (long long)p.B * p.rn) return;
    const int dn = p.dn;
    const float* sdf = p.sdf + ray * dn;
    const float* grad = p.grad + ray * dn * 3;
    const float* col = p.colors + ray * dn * 4;
    const float* dep = p.depth + ray * dn;
    const float qx = p.que_dir[ray * 3], qy = p.que_dir[ray * 3 + 1], qz = p.que_dir[ray * 3 + 2];
    const float car = p.cos_anneal_ratio;
    float T = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, rd = 0.f, eik = 0.f;
    for (int d = 0; d < dn; ++d) {
        const float gx = grad[d * 3], gy = grad[d * 3 + 1], gz = grad[d * 3 + 2];
        const float dist = (d + 1 < dn) ? __fsub_rn(dep[d + 1], dep[d]) : 1e6f;      // depth2dists render_ops.py:41-44
        const float true_cos = ((-qx * gx) + (-qy * gy)) + (-qz * gz);
        const float iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.f - car) + fmaxf(-true_cos, 0.f) * car);
        const float s = sdf[d];
        const float nxt = s + iter_cos * dist * 0.5f, prv = s - iter_cos * dist * 0.5f;
        const float pc = k3_sigmoid(prv * p.inv_s), nc = k3_sigmoid(nxt * p.inv_s);
        const float alpha = fminf(fmaxf(__fdiv_rn((pc - nc) + 1e-5f, pc + 1e-5f), 0.f), 1.f);
        const float hit = alpha * T;                                                 // render_ops.py:77-79
        T *= (1.f - alpha) + 1e-10f;
        p.alpha[ray * dn + d] = alpha;
        p.hit_prob[ray * dn + d] = hit;
        c0 = fmaf(hit, col[d * 4], c0); c1 = fmaf(hit, col[d * 4 + 1], c1); c2 = fmaf(hit, col[d * 4 + 2], c2);
        rd = fmaf(hit, dep[d], rd);
        const float gn = sqrtf(gx * gx + gy * gy + gz * gz) - 1.f;
        eik = fmaf(gn, gn, eik);
    }
    p.pixel_colors[ray * 3] = c0; p.pixel_colors[ray * 3 + 1] = c1; p.pixel_colors[ray * 3 + 2] = c2;
    p.render_depth[ray] = rd;
    p.eik_partial[ray] = eik;
}

extern "C" int gn_k3_composite(const GnK3Params* hp, void* stream)
{
    const GnK3Params& p = *hp;
    if (p.B < 1 || p.rn < 1 || p.dn < 1) return -1;
    const long long rays = (long long)p.B * p.rn;
    const int threads = 128;
    gn_k3_composite_kernel<<<(unsigned)((rays + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

__global__ void gn_k3_coarse_kernel(const float* __restrict__ depth_range, float* __restrict__ depth, int B, int rn, int dn)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * rn * dn) return;
    const int d = (int)(i % dn);
    const int b = (int)(i / ((long long)rn * dn));
    const float near = depth_range[b * 2], far = depth_range[b * 2 + 1];
    const float inear = __fdiv_rn(1.f, near), span = __fsub_rn(__fdiv_rn(1.f, far), inear);
    const float interval = __fdiv_rn(span, (float)(dn - 1));
    float tick;
    if (d == 0) tick = 0.f;
    else if (d == dn - 1) tick = span;
    else tick = __fmul_rn(interval, (float)d);
    depth[i] = __fdiv_rn(1.f, __fadd_rn(inear, tick));
}

extern "C" int gn_k3_coarse_depths(const float* depth_range, float* depth, int B, int rn, int dn, void* stream)
{
    if (B < 1 || rn < 1 || dn < 3) return -1;
    const long long n = (long long)B * rn * dn;
    gn_k3_coarse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(depth_range, depth, B, rn, dn);
    return (int)cudaGetLastError();
}

#define K3F_THREADS 64
__global__ void gn_k3_fine_kernel(const float* __restrict__ depth, const float* __restrict__ hit_prob,
                                  const float* __restrict__ depth_range, const float* __restrict__ u,
                                  float* __restrict__ fine_depth, long long* __restrict__ inds,
                                  int B, int rn, int dn, int fdn)
{
    extern __shared__ float k3s[];
    const int per = 2 * (dn + 1) + fdn;
    float* cdf = k3s + threadIdx.x * per;
    float* centre = cdf + dn + 1;
    float* out = centre + dn + 1;
    const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= (long long)B * rn) return;
    const int b = (int)(ray / rn);
    const float near = __fdiv_rn(-1.f, depth_range[b * 2]), far = __fdiv_rn(-1.f, depth_range[b * 2 + 1]);
    const float span = __fsub_rn(far, near);
    const float* dep = depth + ray * dn;
    const float* hp = hit_prob + ray * dn;
    float prev = 0.f, sum = 0.f;
    for (int i = 0; i < dn; ++i) {
        const float di = __fdiv_rn(__fsub_rn(__fdiv_rn(-1.f, dep[i]), near), span);
        if (i == 0) centre[0] = di;
        else centre[i] = __fdiv_rn(__fadd_rn(di, prev), 2.f);
        prev = di;
        const float h = __fadd_rn(hp[i], 1e-5f);
        sum = (i == 0) ? h : __fadd_rn(sum, h);
    }
    centre[dn] = prev;
    cdf[0] = 0.f;
    float run = 0.f;
    for (int i = 0; i < dn; ++i) {
        const float pdf = __fdiv_rn(__fadd_rn(hp[i], 1e-5f), sum);
        run = (i == 0) ? pdf : __fadd_rn(run, pdf);
        cdf[i + 1] = run;
    }
    for (int j = 0; j < fdn; ++j) {
        const float uj = u[ray * fdn + j];
        int idx = 0;                                         // searchsorted(right=True): first i with cdf[i] > u
        while (idx <= dn && !(cdf[idx] > uj)) ++idx;
        const int below = max(idx - 1, 0), above = min(idx, dn);
        const float c0 = cdf[below], c1 = cdf[above], b0 = centre[below], b1 = centre[above];
        float denom = __fsub_rn(c1, c0);
        if (denom < 1e-5f) denom = 1.f;
        const float tt = __fdiv_rn(__fsub_rn(uj, c0), denom);
        float fd = __fadd_rn(b0, __fmul_rn(tt, __fsub_rn(b1, b0)));
        fd = __fdiv_rn(-1.f, __fadd_rn(__fmul_rn(fd, span), near));
        if (inds) inds[ray * fdn + j] = idx;
        // insertion sort (ascending), renderer.py:148
        int k = j;
        while (k > 0 && out[k - 1] > fd) { out[k] = out[k - 1]; --k; }
        out[k] = fd;
    }
    for (int j = 0; j < fdn; ++j) fine_depth[ray * fdn + j] = out[j];
}

extern "C" int gn_k3_fine_depths(const float* depth, const float* hit_prob, const float* depth_range, const float* u,
                                 float* fine_depth, int64_t* inds, int B, int rn, int dn, int fdn, void* stream)
{
    if (B < 1 || rn < 1 || dn < 2 || fdn < 1) return -1;
    const size_t smem = (size_t)K3F_THREADS * (2 * (dn + 1) + fdn) * sizeof(float);
    if (smem > 227 * 1024) return -5;
    static size_t smem_cache_gn_k3_fine_kernel[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k3_fine_kernel, smem, smem_cache_gn_k3_fine_kernel);
    if (e != cudaSuccess) return (int)e;
    const long long rays = (long long)B * rn;
    gn_k3_fine_kernel<<<(unsigned)((rays + K3F_THREADS - 1) / K3F_THREADS), K3F_THREADS, smem, (cudaStream_t)stream>>>(
        depth, hit_prob, depth_range, u, fine_depth, (long long*)inds, B, rn, dn, fdn);
    return (int)cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------------
// Ray set-up of the RGB head: coords2rays (render_ops.py:4-25), depth2points (27-39), depth2inv_dists (46-52) in ONE launch
// (the reference - and round 1 of this repo - runs a 3x3 torch.inverse, three bmm and ~10 element-wise kernels per chunk).
// thread <-> ray.  K^-1 is the adjugate / determinant in fp32 (torch.inverse factorises; both are exact to ~1 ulp for the
// upper-triangular pinhole K), the world direction is (R^T (K^-1 [x,y,1]) + c) - c like render_ops.py:22-23.
__global__ void gn_k3_ray_setup_kernel(const GnRaySetupParams p)
{
    const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= (long long)p.B * p.rn) return;
    const int b = (int)(ray / p.rn);
    const float* K = p.Ks + b * 9;
    const float* P = p.poses + b * 12;
    const float a = K[0], bb = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
    const float A = e * i - f * h, Bc = -(d * i - f * g), Cc = d * h - e * g;
    const float det = a * A + bb * Bc + c * Cc;
    const float id = __fdiv_rn(1.f, det);
    const float inv[9] = { A * id, -(bb * i - c * h) * id, (bb * f - c * e) * id,
                           Bc * id, (a * i - c * g) * id, -(a * f - c * d) * id,
                           Cc * id, -(a * h - bb * g) * id, (a * e - bb * d) * id };
    const float x = p.coords[ray * 2], y = p.coords[ray * 2 + 1];
    const float cx = inv[0] * x + inv[1] * y + inv[2], cy = inv[3] * x + inv[4] * y + inv[5], cz = inv[6] * x + inv[7] * y + inv[8];
    // rot = R^T (render_ops.py:14), centre = -R^T t (15)
    float ctr[3], dir[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        ctr[r] = -(P[0 * 4 + r] * P[3] + P[1 * 4 + r] * P[7] + P[2 * 4 + r] * P[11]);
        const float w = (P[0 * 4 + r] * cx + P[1 * 4 + r] * cy + P[2 * 4 + r] * cz) + ctr[r];
        dir[r] = w - ctr[r];
    }
    if (p.centers) { p.centers[ray * 3] = ctr[0]; p.centers[ray * 3 + 1] = ctr[1]; p.centers[ray * 3 + 2] = ctr[2]; }
    if (p.dirs) { p.dirs[ray * 3] = dir[0]; p.dirs[ray * 3 + 1] = dir[1]; p.dirs[ray * 3 + 2] = dir[2]; }
    const float nrm = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    p.que_dir[ray * 3] = __fdiv_rn(-dir[0], nrm); p.que_dir[ray * 3 + 1] = __fdiv_rn(-dir[1], nrm); p.que_dir[ray * 3 + 2] = __fdiv_rn(-dir[2], nrm);
    const float rnear = __fdiv_rn(-1.f, p.depth_range[b * 2]), rfar = __fdiv_rn(-1.f, p.depth_range[b * 2 + 1]);
    const float* dep = p.depth + ray * p.dn;
    float* pts = p.pts + ray * p.dn * 3;
    float* qd = p.inv_dists + ray * p.dn;
    float cur = __fdiv_rn(__fdiv_rn(-1.f, dep[0]) - rnear, rfar - rnear);
    for (int s = 0; s < p.dn; ++s) {
        const float z = dep[s];
        pts[s * 3] = ctr[0] + dir[0] * z; pts[s * 3 + 1] = ctr[1] + dir[1] * z; pts[s * 3 + 2] = ctr[2] + dir[2] * z;
        float nxt = 0.f;
        if (s + 1 < p.dn) nxt = __fdiv_rn(__fdiv_rn(-1.f, dep[s + 1]) - rnear, rfar - rnear);
        qd[s] = (s + 1 < p.dn) ? __fsub_rn(nxt, cur) : 1e6f;                        // depth2dists: last spacing 1e6
        cur = nxt;
    }
}

extern "C" int gn_k3_ray_setup(const GnRaySetupParams* hp, void* stream)
{
    const GnRaySetupParams& p = *hp;
    if (p.B < 1 || p.rn < 1 || p.dn < 1) return -1;
    if (!p.coords || !p.poses || !p.Ks || !p.depth || !p.depth_range || !p.pts || !p.que_dir || !p.inv_dists) return -2;
    const long long rays = (long long)p.B * p.rn;
    gn_k3_ray_setup_kernel<<<(unsigned)((rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ depth-mean head
// renderer.py:222-266 (predict_mean_for_depth_loss) in one launch: bilinear sample of ray_feats at `num` pixels per reference
// view (interpolate_feats -> F.grid_sample, border padding, ops.py:14-34) and the mean decoders of dist_decoder /
// fine_dist_decoder (32 -> 32 -> 32 -> 2, ELU ELU Softplus; dist_decoder.py:64-70,148-150).  thread <-> (view, pixel); the
// six Linear layers (17 KB) sit in shared memory and are read as broadcasts.
#define DM_THREADS 128
#define DM_DEC_FLOATS (32 * 32 + 32 + 32 * 32 + 32 + 2 * 32 + 2)

__device__ __forceinline__ float dm_elu(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float dm_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// grid_sampler_unnormalize + clip_coordinates (at::native GridSampler.cuh), padding_mode='border'
__device__ __forceinline__ float dm_source_index(float g, int size, int align)
{
    float x = align ? ((g + 1.f) / 2.f) * (float)(size - 1) : ((g + 1.f) * (float)size - 1.f) / 2.f;
    return fminf((float)(size - 1), fmaxf(x, 0.f));
}

__device__ __forceinline__ void dm_decode(const float* __restrict__ w, const float* x, float* out2)
{
    float h1[32], h2[32];
#pragma unroll
    for (int n = 0; n < 32; ++n) {
        float a = w[1024 + n];
#pragma unroll
        for (int k = 0; k < 32; ++k) a = fmaf(w[n * 32 + k], x[k], a);
        h1[n] = dm_elu(a);
    }
    const float* w1 = w + 1056;
#pragma unroll
    for (int n = 0; n < 32; ++n) {
        float a = w1[1024 + n];
#pragma unroll
        for (int k = 0; k < 32; ++k) a = fmaf(w1[n * 32 + k], h1[k], a);
        h2[n] = dm_elu(a);
    }
    const float* w2 = w + 2112;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        float a = w2[64 + n];
#pragma unroll
        for (int k = 0; k < 32; ++k) a = fmaf(w2[n * 32 + k], h2[k], a);
        out2[n] = dm_softplus(a);
    }
}

__global__ void __launch_bounds__(DM_THREADS)
gn_k3_depth_mean_kernel(const GnDepthMeanParams p)
{
    __shared__ __align__(16) float sw[2 * DM_DEC_FLOATS];
    const int ndec = p.w_fine[0] ? 2 : 1;
    for (int d = 0; d < ndec; ++d) {
        const float* const* src = d ? p.w_fine : p.w_coarse;
        const int off[7] = {0, 1024, 1056, 2080, 2112, 2176, 2178};
        for (int t = 0; t < 6; ++t)
            for (int i = threadIdx.x; i < off[t + 1] - off[t]; i += DM_THREADS) sw[d * DM_DEC_FLOATS + off[t] + i] = __ldg(src[t] + i);
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * DM_THREADS + threadIdx.x;
    if (i >= (long long)p.V * p.num) return;
    const int v = (int)(i / p.num);
    // the reference stacks (row, col) and feeds it to interpolate_feats as (x, y) (renderer.py:229-243): kept as is
    const float c0 = (float)p.coords[2 * i], c1 = (float)p.coords[2 * i + 1];
    const float gx = c0 / (float)(p.W - 1) * 2.f - 1.f, gy = c1 / (float)(p.H - 1) * 2.f - 1.f;
    const float ix = dm_source_index(gx, p.fw, p.align_corners), iy = dm_source_index(gy, p.fh, p.align_corners);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = (fx + 1.f - ix) * (fy + 1.f - iy), wne = (ix - fx) * (fy + 1.f - iy);
    const float wsw = (fx + 1.f - ix) * (iy - fy), wse = (ix - fx) * (iy - fy);
    const bool xin = x1 < p.fw, yin = y1 < p.fh;               // (x0, y0) is always inside after the border clip
    const float* f = p.feats + (long long)v * p.stride_v;
    float x[32];
    if (p.stride_c == 1 && (p.stride_x & 3) == 0 && (p.stride_y & 3) == 0 && (p.stride_v & 3) == 0 && (reinterpret_cast<uintptr_t>(p.feats) & 15) == 0) {
        // channels-last map (the fused K1 feature buffer): a tap's 32 channels are one 128-byte run -> 8 float4 loads per tap
        const float4* t00 = reinterpret_cast<const float4*>(f + (long long)y0 * p.stride_y + (long long)x0 * p.stride_x);
        const float4* t01 = reinterpret_cast<const float4*>(f + (long long)y0 * p.stride_y + (long long)(xin ? x1 : x0) * p.stride_x);
        const float4* t10 = reinterpret_cast<const float4*>(f + (long long)(yin ? y1 : y0) * p.stride_y + (long long)x0 * p.stride_x);
        const float4* t11 = reinterpret_cast<const float4*>(f + (long long)(yin ? y1 : y0) * p.stride_y + (long long)(xin ? x1 : x0) * p.stride_x);
        const float w01 = xin ? wne : 0.f, w10 = yin ? wsw : 0.f, w11 = (xin && yin) ? wse : 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            const float4 a = __ldg(t00 + c4), b = __ldg(t01 + c4), c = __ldg(t10 + c4), d = __ldg(t11 + c4);
            // same order of operations as the scalar path: a*wnw (+ b*wne) (+ c*wsw) (+ d*wse); a skipped tap has weight 0 there
            float r0 = a.x * wnw, r1 = a.y * wnw, r2 = a.z * wnw, r3 = a.w * wnw;
            if (xin) { r0 += b.x * w01; r1 += b.y * w01; r2 += b.z * w01; r3 += b.w * w01; }
            if (yin) { r0 += c.x * w10; r1 += c.y * w10; r2 += c.z * w10; r3 += c.w * w10; }
            if (xin && yin) { r0 += d.x * w11; r1 += d.y * w11; r2 += d.z * w11; r3 += d.w * w11; }
            x[4 * c4] = r0; x[4 * c4 + 1] = r1; x[4 * c4 + 2] = r2; x[4 * c4 + 3] = r3;
        }
    } else
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const float* fc = f + (long long)c * p.stride_c;
        float a = __ldg(fc + (long long)y0 * p.stride_y + (long long)x0 * p.stride_x) * wnw;
        if (xin) a += __ldg(fc + (long long)y0 * p.stride_y + (long long)x1 * p.stride_x) * wne;
        if (yin) a += __ldg(fc + (long long)y1 * p.stride_y + (long long)x0 * p.stride_x) * wsw;
        if (xin && yin) a += __ldg(fc + (long long)y1 * p.stride_y + (long long)x1 * p.stride_x) * wse;
        x[c] = a;
    }
    float o[2];
    dm_decode(sw, x, o);
    p.mean[2 * i] = o[0]; p.mean[2 * i + 1] = o[1];
    if (ndec == 2) {
        dm_decode(sw + DM_DEC_FLOATS, x, o);
        p.mean_fine[2 * i] = o[0]; p.mean_fine[2 * i + 1] = o[1];
    }
}

extern "C" int gn_k3_depth_mean(const GnDepthMeanParams* hp, void* stream)
{
    const GnDepthMeanParams& p = *hp;
    if (p.V < 1 || p.num < 1 || p.H < 2 || p.W < 2 || p.fh < 1 || p.fw < 1) return -1;
    if (!p.feats || !p.coords || !p.mean) return -2;
    for (int t = 0; t < 6; ++t) {
        if (!p.w_coarse[t]) return -2;
        if ((p.w_fine[t] == nullptr) != (p.w_fine[0] == nullptr)) return -2;
    }
    if (p.w_fine[0] && !p.mean_fine) return -2;
    const long long n = (long long)p.V * p.num;
    gn_k3_depth_mean_kernel<<<(unsigned)((n + DM_THREADS - 1) / DM_THREADS), DM_THREADS, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
