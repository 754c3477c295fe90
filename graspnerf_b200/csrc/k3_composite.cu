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
