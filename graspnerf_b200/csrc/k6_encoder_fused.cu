// K6: fused element-wise stages of the 2-D encoders (src/nr/network/ops.py:78-230, init_net.py:8-35, vis_encoder.py:6-21).
// The reference's ResUNetLight / residual blocks run, per 3x3 convolution, a reflection-pad kernel, the convolution, an
// InstanceNorm (two kernels + a `repeat` of the affine parameters), an activation and often a residual add: ~6 launches
// whose cost at batch 6 is launch / latency, not bytes (profiles/profile_forward_r02e.txt: pad 12 %, instance norm + its
// repeat / copy 18 %, bilinear upsampling 11 %, ReLU 2 % of the encoders' GPU time).  Here ONE kernel per layer does
//   out = reflect_pad( act( InstanceNorm(x) * gamma + beta  [+ residual | + InstanceNorm(residual) * g_r + b_r] ), p )
// and writes the tensor already padded for the NEXT convolution (which then runs with padding 0), plus the un-padded copy
// when a 1x1 / skip consumer needs it.  The convolutions themselves stay in cuDNN (fp32).
//   gn_k6_norm_act_pad   one CTA per (image, channel) plane; mean / biased variance in two passes (fp32, like at::native).
//   gn_k6_upsample2x_pad F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) (ops.py:142-150) + reflect pad.
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"

#define K6_THREADS 256
#define K6_CACHE_FLOATS 9216      // a CTA's share of the plane stays in shared memory between the three passes (72x128 map: 36 KB)

__device__ __forceinline__ float k6_block_sum(float v, float* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < K6_THREADS / 32; ++w) t += s_red[w];
    return t;
}

__device__ __forceinline__ int k6_reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// Sum over the CLUSTER of CTAs that share a plane (cluster size CS: 1 or 8).  Large planes (the 144x256 / 72x128 maps of the
// first stages: 6 images x 16..32 channels = only 96..192 planes for 148 SMs) are split over a thread-block cluster; the
// partial sums meet through distributed shared memory (each CTA reads its peers' partials after a cluster barrier).
template <int CS>
__device__ __forceinline__ float k6_cluster_sum(float v, float* s_red, float* s_part, int phase)
{
    float t = k6_block_sum(v, s_red);
    if (CS == 1) return t;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) s_part[phase] = t;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < CS; ++r) {                        // same order in every CTA: identical mean / rstd across the cluster
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(&s_part[phase])), "r"(r));
        float pv;
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pv) : "r"(remote) : "memory");
        tot += pv;
    }
    return tot;
}

// statistics of rows [r0, r1) of this CTA's share of a plane, combined over the cluster
// x as the sum of `splits` partial tensors (split-K convolution output), always added in the same order
__device__ __forceinline__ float k6_ld(const float* __restrict__ x, int off, int splits, long long stride) {
    float v = x[off];
    for (int s = 1; s < splits; ++s) v += x[(long long)s * stride + off];
    return v;
}

// cache: optional shared-memory copy [(r1-r0)*W] of the rows read here (filled in the first pass, re-read in the second - each
// thread re-reads exactly what it wrote - and by the caller's output pass after the block-wide barriers of the reductions)
template <int CS>
__device__ __forceinline__ void k6_plane_stats_cl(const float* __restrict__ x, int W, int xs, int r0, int r1, int n_total, float eps,
                                                  float* s_red, float* s_part, int phase0, float& mean, float& rstd,
                                                  int splits = 1, long long sstride = 0, float* cache = nullptr)
{
    // a warp walks a row, lanes stride over the columns: coalesced, no integer division in the loops
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float s = 0.f;
    if (cache) {
        for (int r = r0 + warp; r < r1; r += K6_THREADS / 32)
            for (int c = lane; c < W; c += 32) { const float v = k6_ld(x, r * xs + c, splits, sstride); cache[(r - r0) * W + c] = v; s += v; }
    } else {
        for (int r = r0 + warp; r < r1; r += K6_THREADS / 32)
            for (int c = lane; c < W; c += 32) s += k6_ld(x, r * xs + c, splits, sstride);
    }
    mean = k6_cluster_sum<CS>(s, s_red, s_part, phase0) / (float)n_total;
    float q = 0.f;
    if (cache) {
        for (int r = r0 + warp; r < r1; r += K6_THREADS / 32)
            for (int c = lane; c < W; c += 32) { const float d = cache[(r - r0) * W + c] - mean; q = fmaf(d, d, q); }
    } else {
        for (int r = r0 + warp; r < r1; r += K6_THREADS / 32)
            for (int c = lane; c < W; c += 32) { const float d = k6_ld(x, r * xs + c, splits, sstride) - mean; q = fmaf(d, d, q); }
    }
    rstd = rsqrtf(k6_cluster_sum<CS>(q, s_red, s_part, phase0 + 1) / (float)n_total + eps);      // biased variance, like InstanceNorm
}

template <int CS>
__global__ void __launch_bounds__(K6_THREADS)
gn_k6_norm_act_pad_kernel(const GnNormActPadParams p)
{
    __shared__ float s_red[K6_THREADS / 32];
    __shared__ float s_part[4];
    __shared__ float s_plane[K6_CACHE_FLOATS];
    const int plane = blockIdx.x / CS;                   // n * C + c
    const int rank = blockIdx.x - plane * CS;
    const int c = plane % p.C;
    const int H = p.H, W = p.W, hw = H * W;
    const int xs = W + 2 * p.x_pad;                      // row stride of the (possibly padded) input
    const float* x = p.x + (size_t)plane * (size_t)(H + 2 * p.x_pad) * xs + (size_t)p.x_pad * xs + p.x_pad;
    // this CTA's rows for the statistics (32-bit arithmetic: H * CS < 2^31; a 64-bit division costs ~100 instructions and these
    // kernels run 2-40 elements per thread)
    const int r0 = CS == 1 ? 0 : (H * rank) / CS, r1 = CS == 1 ? H : (H * (rank + 1)) / CS;
    float g = 1.f, b = 0.f;
    const bool cached = p.gamma && (r1 - r0) * W <= K6_CACHE_FLOATS;
    if (p.gamma) {
        float mean, rstd;
        k6_plane_stats_cl<CS>(x, W, xs, r0, r1, hw, p.eps, s_red, s_part, 0, mean, rstd, p.x_splits, p.x_split_stride, cached ? s_plane : nullptr);
        g = __ldg(p.gamma + c) * rstd; b = __ldg(p.beta + c) - mean * g;          // y = (x - mean) * rstd * gamma + beta
    }
    const float* r = nullptr;
    int rs = 0;
    float rg = 1.f, rb = 0.f;
    if (p.res) {
        rs = W + 2 * p.res_pad;
        r = p.res + (size_t)plane * (size_t)(H + 2 * p.res_pad) * rs + (size_t)p.res_pad * rs + p.res_pad;
        if (p.res_gamma) {
            float rm, rr;
            k6_plane_stats_cl<CS>(r, W, rs, r0, r1, hw, p.eps, s_red, s_part, 2, rm, rr);
            rg = __ldg(p.res_gamma + c) * rr; rb = __ldg(p.res_beta + c) - rm * rg;
        }
    }
    const int P = p.pad, Hp = H + 2 * P, Wp = W + 2 * P;
    float* op = p.out_padded ? p.out_padded + (size_t)plane * Hp * Wp : nullptr;
    float* ou = p.out_unpadded ? p.out_unpadded + (size_t)plane * hw : nullptr;
    const int h0 = CS == 1 ? 0 : (Hp * rank) / CS, h1 = CS == 1 ? Hp : (Hp * (rank + 1)) / CS;        // this CTA's padded output rows
    for (int hp = h0 + (int)(threadIdx.x >> 5); hp < h1; hp += K6_THREADS / 32) {
        const int h = k6_reflect(hp - P, H);
        const bool hin = hp >= P && hp < P + H;
        for (int wp = threadIdx.x & 31; wp < Wp; wp += 32) {
            const int w = k6_reflect(wp - P, W);
            const float xv = (cached && h >= r0 && h < r1) ? s_plane[(h - r0) * W + w] : k6_ld(x, h * xs + w, p.x_splits, p.x_split_stride);
            float v = fmaf(xv, g, b);
            if (r) v += fmaf(r[h * rs + w], rg, rb);
            if (p.act == 1) v = fmaxf(v, 0.f);
            else if (p.act == 2) v = v > 0.f ? v : expm1f(v);                      // F.elu
            if (op) op[hp * Wp + wp] = v;
            if (ou && hin && wp >= P && wp < P + W) ou[(hp - P) * W + (wp - P)] = v;
        }
    }
    if (CS > 1)       // a CTA must not exit while peers may still read its partial sums
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// no statistics at all (reflection pad / copy / activation / residual add only): plain grid-stride element-wise kernel
__global__ void __launch_bounds__(K6_THREADS)
gn_k6_pad_only_kernel(const GnNormActPadParams p)
{
    const int H = p.H, W = p.W, P = p.pad, Hp = H + 2 * P, Wp = W + 2 * P;
    const int xs = W + 2 * p.x_pad, rs = W + 2 * p.res_pad;
    const size_t xplane = (size_t)(H + 2 * p.x_pad) * xs, rplane = (size_t)(H + 2 * p.res_pad) * rs;
    // grid: x covers one padded plane (32-bit index, one 32-bit division per element), y strides over the planes
    const int per_plane = Hp * Wp, planes = p.N * p.C;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const float* xp = p.x + (size_t)plane * xplane + (size_t)p.x_pad * xs + p.x_pad;
        const float* rp = p.res ? p.res + (size_t)plane * rplane + (size_t)p.res_pad * rs + p.res_pad : nullptr;
        float* op = p.out_padded ? p.out_padded + (size_t)plane * per_plane : nullptr;
        float* ou = p.out_unpadded ? p.out_unpadded + (size_t)plane * H * W : nullptr;
        for (int i = blockIdx.x * K6_THREADS + threadIdx.x; i < per_plane; i += gridDim.x * K6_THREADS) {
            const int hp = i / Wp, wp = i - hp * Wp;
            const int h = k6_reflect(hp - P, H), w = k6_reflect(wp - P, W);
            float v = xp[h * xs + w];
            if (rp) v += rp[h * rs + w];
            if (p.act == 1) v = fmaxf(v, 0.f);
            else if (p.act == 2) v = v > 0.f ? v : expm1f(v);
            if (op) op[i] = v;
            if (ou && hp >= P && hp < P + H && wp >= P && wp < P + W) ou[(hp - P) * W + (wp - P)] = v;
        }
    }
}

extern "C" int gn_k6_norm_act_pad(const GnNormActPadParams* hp, void* stream)
{
    const GnNormActPadParams& p = *hp;
    if (p.N < 1 || p.C < 1 || p.H < 1 || p.W < 1 || p.pad < 0 || p.pad >= p.H || p.pad >= p.W || p.x_pad < 0 || p.res_pad < 0) return -1;
    if (!p.x || (!p.out_padded && !p.out_unpadded) || ((p.gamma == nullptr) != (p.beta == nullptr)) || ((p.res_gamma == nullptr) != (p.res_beta == nullptr))) return -2;
    if (p.res_gamma && !p.res) return -2;
    if (p.act < 0 || p.act > 2) return -3;
    if (!p.out_padded && p.pad != 0) return -3;
    if (p.x_splits > 1 && (!p.gamma || p.x_pad != 0)) return -4;         // partial sums only feed a normalising stage
    cudaStream_t st = (cudaStream_t)stream;
    if (!p.gamma && !p.res_gamma) {
        const long long per_plane = (long long)(p.H + 2 * p.pad) * (p.W + 2 * p.pad);
        if (per_plane > 0x7fffffffLL) return -6;
        const long long planes = (long long)p.N * p.C;
        const unsigned bx = (unsigned)((per_plane + K6_THREADS - 1) / K6_THREADS < 64 ? (per_plane + K6_THREADS - 1) / K6_THREADS : 64);
        gn_k6_pad_only_kernel<<<dim3(bx, (unsigned)(planes < 65535 ? planes : 65535), 1), K6_THREADS, 0, st>>>(p);
        return (int)cudaGetLastError();
    }
    const int planes = p.N * p.C;
    if ((long long)p.H * p.W >= 16384 && p.H >= 16) {      // large planes (>= 128x128): a cluster of 8 CTAs per plane (DSMEM reduction)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)planes * 8, 1, 1);
        cfg.blockDim = dim3(K6_THREADS, 1, 1);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gn_k6_norm_act_pad_kernel<8>, p);
        return (int)(e != cudaSuccess ? e : cudaGetLastError());
    }
    gn_k6_norm_act_pad_kernel<1><<<(unsigned)planes, K6_THREADS, 0, st>>>(p);
    return (int)cudaGetLastError();
}

// F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True) then reflection padding by `pad`
// (at::native upsample_bilinear2d: scale = (in-1)/(out-1) in fp32, src = scale*dst, lambda1 = src - floor(src)).
__global__ void __launch_bounds__(K6_THREADS)
gn_k6_upsample2x_pad_kernel(const float* __restrict__ x, float* __restrict__ out, int planes, int H, int W, int pad)
{
    const int Ho = 2 * H, Wo = 2 * W, Hp = Ho + 2 * pad, Wp = Wo + 2 * pad;
    const int per_plane = Hp * Wp;
    const float sh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, sw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {                  // x: inside one padded plane (32-bit index math), y: planes
        const float* xp = x + (size_t)pl * H * W;
        float* op = out + (size_t)pl * per_plane;
        for (int i = blockIdx.x * K6_THREADS + threadIdx.x; i < per_plane; i += gridDim.x * K6_THREADS) {
            const int hp = i / Wp, wp = i - hp * Wp;
            const int ho = k6_reflect(hp - pad, Ho), wo = k6_reflect(wp - pad, Wo);
            const float fh = sh * (float)ho, fw = sw * (float)wo;
            const int h0 = (int)fh, w0 = (int)fw;
            const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
            const float lh1 = fh - (float)h0, lw1 = fw - (float)w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
            op[i] = lh0 * (lw0 * __ldg(xp + h0 * W + w0) + lw1 * __ldg(xp + h0 * W + w1)) + lh1 * (lw0 * __ldg(xp + h1 * W + w0) + lw1 * __ldg(xp + h1 * W + w1));
        }
    }
}

extern "C" int gn_k6_upsample2x_pad(const float* x, float* out, int planes, int H, int W, int pad, void* stream)
{
    if (!x || !out || planes < 1 || H < 1 || W < 1 || pad < 0 || pad >= 2 * H || pad >= 2 * W) return -1;
    const long long per_plane = (long long)(2 * H + 2 * pad) * (2 * W + 2 * pad);
    if (per_plane > 0x7fffffffLL) return -6;
    const long long nb = (per_plane + K6_THREADS - 1) / K6_THREADS;
    gn_k6_upsample2x_pad_kernel<<<dim3((unsigned)(nb < 64 ? nb : 64), (unsigned)(planes < 65535 ? planes : 65535), 1), K6_THREADS, 0, (cudaStream_t)stream>>>(x, out, planes, H, W, pad);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ layout glue around the encoders
// gn_k6_fuse_features: the encoders' two NCHW maps [P,32,h,w] -> the fused channels-last buffer [P,h,w,64] (ray | img) K1 gathers
// from (ops.Scene).  A 64-channel x 32-pixel tile goes through shared memory: reads coalesced along w, writes along the channels.
__global__ void __launch_bounds__(256)
gn_k6_fuse_features_kernel(const float* __restrict__ ray, const float* __restrict__ img, float* __restrict__ out, int hw)
{
    __shared__ float tile[64][33];
    const int plane = blockIdx.y, p0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t src = (size_t)plane * 32 * hw;
#pragma unroll
    for (int c = warp; c < 64; c += 8) {
        const float* s = (c < 32 ? ray + src + (size_t)c * hw : img + src + (size_t)(c - 32) * hw);
        tile[c][lane] = p0 + lane < hw ? __ldg(s + p0 + lane) : 0.f;
    }
    __syncthreads();
    float* o = out + ((size_t)plane * hw + p0) * 64;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + 256 * i, px = idx >> 6, ch = idx & 63;
        if (p0 + px < hw) o[idx] = tile[ch][px];
    }
}

extern "C" int gn_k6_fuse_features(const float* ray_feats, const float* img_feats, float* out, int planes, int hw, void* stream)
{
    if (!ray_feats || !img_feats || !out || planes < 1 || planes > 65535 || hw < 1) return -1;
    gn_k6_fuse_features_kernel<<<dim3((unsigned)((hw + 31) / 32), (unsigned)planes, 1), 256, 0, (cudaStream_t)stream>>>(ray_feats, img_feats, out, hw);
    return (int)cudaGetLastError();
}

// gn_k6_images_u8: uint8 images [V,H,W,C] (C = 3 or 4, as imread gives them) -> fp32 [V,3,H,W] = u8 / 255 (color_map_forward,
// main.py:170: a true division, like numpy's) for the encoders and, optionally, the uint8 RGBA texel buffer [V,H,W,4] of K1.
__global__ void __launch_bounds__(256)
gn_k6_images_u8_kernel(const unsigned char* __restrict__ in, float* __restrict__ out_f, unsigned char* __restrict__ out_rgba, int V, int HW, int C)
{
    const int v = blockIdx.y;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
        const unsigned char* px = in + ((size_t)v * HW + i) * C;
        const unsigned char r = px[0], g = px[1], b = px[2];
        float* o = out_f + (size_t)v * 3 * HW + i;
        o[0] = __fdiv_rn((float)r, 255.f); o[HW] = __fdiv_rn((float)g, 255.f); o[2 * (size_t)HW] = __fdiv_rn((float)b, 255.f);
        if (out_rgba) reinterpret_cast<uchar4*>(out_rgba)[(size_t)v * HW + i] = make_uchar4(r, g, b, 0);
    }
}

extern "C" int gn_k6_images_u8(const unsigned char* in, float* out_f, unsigned char* out_rgba, int V, int H, int W, int C, void* stream)
{
    if (!in || !out_f || V < 1 || V > 65535 || H < 1 || W < 1 || (C != 3 && C != 4)) return -1;
    const long long hw = (long long)H * W;
    if (hw > 0x7fffffffLL) return -6;
    const long long nb = (hw + 255) / 256;
    gn_k6_images_u8_kernel<<<dim3((unsigned)(nb < 1024 ? nb : 1024), (unsigned)V, 1), 256, 0, (cudaStream_t)stream>>>(in, out_f, out_rgba, V, (int)hw, C);
    return (int)cudaGetLastError();
}
