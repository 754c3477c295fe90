// Common device helpers for the graspnerf_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GN_FEAT_C 32          // channels of img_feats / ray_feats (renderer.py:53, init_net.py:24)
#define GN_REC_STRIDE 72      // floats per (point,view) record
#define GN_PT_STRIDE 2        // floats per point written by K1: nvalid, view bit mask (uint32 bits)
#define GN_POOL_STRIDE 68     // floats per point written by K2a: mean32 | var32 | wmean, nvalid, 0, 0
#define GN_TOK_STRIDE 20      // floats per point written by K2a (tensor-core path): geometry_fc output 16 | nvalid, 0, 0, 0

// record layout (floats), see DESIGN.md "Data layout in HBM".  Two halves of 36 floats = the two TMA boxes K2a stages
// (csrc/k2a_head_tc3.cu): [ray_feats | dir_diff] is what the first GEMM rounds need, [rgb, depth | img_feats] the later ones.
#define GN_REC_RAYF 0         // [0,32)  ray_feats * mask
#define GN_REC_DD 32          // [32,36) dir_diff = (dir - que_dir, dir . que_dir)   (aggregate_net.py:11-17)
#define GN_REC_RGB 36         // [36,39) rgb * mask ; [39] projection depth
#define GN_REC_DEPTH 39
#define GN_REC_IMGF 40        // [40,72) img_feats * mask
#define GN_REC_HALF 36        // floats per half / TMA box

__device__ __forceinline__ float gn_elu(float x) {
    // nn.ELU (alpha 1): x>0 ? x : exp(x)-1.  __expf = ex2.approx(x*log2e): rel. err ~2^-21.
    return x > 0.f ? x : (__expf(x) - 1.f);
}
__device__ __forceinline__ float gn_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float gn_softplus(float x) {
    // nn.Softplus(beta=1, threshold=20)
    return x > 20.f ? x : log1pf(__expf(x));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// record store: written once by K1, read once by K2a right after.  Default: streaming (evict-first) store.
// GN_K1_STORE_DEFAULT (experiment): normal write-back policy, so the 110 MB record of a 40^3 scene can stay in the 126 MB L2
// until K2a reads it.
#ifdef GN_K1_STORE_DEFAULT
__device__ __forceinline__ void st4_cs(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
#else
__device__ __forceinline__ void st4_cs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
#endif

__device__ __forceinline__ float4 f4_fma(float4 a, float w, float4 acc) {
    acc.x = fmaf(a.x, w, acc.x); acc.y = fmaf(a.y, w, acc.y);
    acc.z = fmaf(a.z, w, acc.z); acc.w = fmaf(a.w, w, acc.w);
    return acc;
}
__device__ __forceinline__ float4 f4_mul(float4 a, float w) { return make_float4(a.x * w, a.y * w, a.z * w, a.w * w); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ---- bilinear tap arithmetic, FIXED op order without FMA (mirrors oracle/nr_oracle.py:bilinear_coords;
// reference ops.py:29-30 + F.grid_sample(padding_mode='border')).
struct GnTap1D { int i0, i1; float w0, w1; };
__device__ __forceinline__ GnTap1D gn_tap1d(float u, int size_img, int size_map, bool align_corners) {
    float xn = __fsub_rn(__fmul_rn(__fdiv_rn(u, (float)(size_img - 1)), 2.f), 1.f);
    float ix;
    if (align_corners) ix = __fmul_rn(__fdiv_rn(__fadd_rn(xn, 1.f), 2.f), (float)(size_map - 1));
    else               ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(xn, 1.f), (float)size_map), 1.f), 2.f);
    ix = fminf(fmaxf(ix, 0.f), (float)(size_map - 1));
    float f0 = floorf(ix);
    GnTap1D t;
    t.w1 = __fsub_rn(ix, f0);
    t.w0 = __fsub_rn(__fadd_rn(f0, 1.f), ix);
    t.i0 = (int)f0;
    t.i1 = min(t.i0 + 1, size_map - 1);
    return t;
}

// Opt a kernel in to `bytes` of dynamic shared memory once per device (cached: later launches - including launches made
// while a stream is being captured into a CUDA graph - do not call into the runtime again).
template <typename K>
static inline cudaError_t gn_ensure_smem(K kernel, size_t bytes, size_t (&cache)[16]) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if (bytes <= cache[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cache[dev] = bytes;
    return e;
}
static inline int gn_sm_count() {
    static int cache[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!cache[dev & 15]) cudaDeviceGetAttribute(&cache[dev & 15], cudaDevAttrMultiProcessorCount, dev);
    return cache[dev & 15];
}
