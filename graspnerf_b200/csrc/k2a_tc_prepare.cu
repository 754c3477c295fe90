// fp32 weight blob -> [fp16 hi/lo K-major operand images | small fp32 constants] (global buffer copied verbatim into the
// shared memory of the tensor-core K2a kernels).  Run once per weight update (gn_k2a_tc_prepare).
#include "k2a_tc_common.cuh"

// ---- prepare: fp32 blob -> [fp16 hi/lo images | small fp32 constants] in global memory ---------------------------------
__device__ void tc_fill(__half* img, int N, int K, const float* __restrict__ src, int ksrc, int nsrc, int cp, int k_dst, int n_dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ksrc * nsrc; i += gridDim.x * blockDim.x) {
        const int k = i / nsrc, n = i - k * nsrc;
        const float w = __ldg(src + k * cp + n);
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const int kk = k + k_dst, nn = n + n_dst;
        const int off = (kk >> 3) * (N * 8) + nn * 8 + (kk & 7);
        img[off] = hi;
        img[N * K + off] = lo;
    }
}
struct TcSmallPlan { int src[kTcSmallCount], n[kTcSmallCount], dst[kTcSmallCount]; };
__global__ void gn_k2a_tc_prepare_kernel(const float* __restrict__ W, unsigned char* __restrict__ out, const TcSmallPlan plan)
{
    __half* s_img = reinterpret_cast<__half*>(out);               // caller zero-fills `out` first
#define IMG(L) (s_img + tc_img_off(L))
    tc_fill(IMG(L_DD1), 96, 32, W + GN_OFF(DD_MEAN_W0), 32, 32, 32, 0, 0);
    tc_fill(IMG(L_DD1), 96, 32, W + GN_OFF(DD_VAR_W0), 32, 32, 32, 0, 32);
    tc_fill(IMG(L_DD1), 96, 32, W + GN_OFF(DD_AW_W0), 32, 32, 32, 0, 64);
    tc_fill(IMG(L_DD2M), 32, 32, W + GN_OFF(DD_MEAN_W2), 32, 32, 32, 0, 0);
    tc_fill(IMG(L_DD2V), 32, 32, W + GN_OFF(DD_VAR_W2), 32, 32, 32, 0, 0);
    tc_fill(IMG(L_DD2A), 32, 32, W + GN_OFF(DD_AW_W2), 32, 32, 32, 0, 0);
    tc_fill(IMG(L_PE0), 32, 48, W + GN_OFF(PE_W0), 34, 32, 32, 0, 0);
    tc_fill(IMG(L_RD0), 16, 16, W + GN_OFF(RD_W0), 4, 16, 16, 0, 0);
    tc_fill(IMG(L_RD1), 48, 16, W + GN_OFF(RD_W1), 16, 36, 36, 0, 0);
    // bf.wg rows are [mean0 36 | var0 36 | mean1 36 | var1 36]; image k order: m0[0..31] m1[0..31] v0[0..31] v1[0..31] tails
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 0 * 64, 32, 64, 64, 0, 0);
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 72 * 64, 32, 64, 64, 32, 0);
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 36 * 64, 32, 64, 64, 64, 0);
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 108 * 64, 32, 64, 64, 96, 0);
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 32 * 64, 3, 64, 64, 128, 0);
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 104 * 64, 3, 64, 64, 131, 0);
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 68 * 64, 3, 64, 64, 134, 0);
    tc_fill(IMG(L_BF0B), 64, 144, W + GN_OFF(BF_WG) + 140 * 64, 3, 64, 64, 137, 0);
    tc_fill(IMG(L_BF2), 32, 64, W + GN_OFF(BF_W2), 64, 32, 32, 0, 0);
    tc_fill(IMG(L_VF0), 32, 32, W + GN_OFF(VF_W0), 32, 32, 32, 0, 0);
    tc_fill(IMG(L_VF2), 48, 32, W + GN_OFF(VF_W2), 32, 36, 36, 0, 0);
    tc_fill(IMG(L_V20), 32, 32, W + GN_OFF(V2_W0), 32, 32, 32, 0, 0);
    tc_fill(IMG(L_GF0), 64, 96, W + GN_OFF(GF_W0), 86, 64, 64, 0, 0);
    tc_fill(IMG(L_GF2), 16, 64, W + GN_OFF(GF_W2), 64, 16, 16, 0, 0);
    tc_fill(IMG(L_NFC), 16, 32, W + GN_OFF(NFC_W0), 32, 8, 8, 0, 0);         // neuray_fc.0 o prob_embed.2 on the ReLU'd hidden
    tc_fill(IMG(L_BF0C), 64, 80, W + GN_OFF(BF_WF), 36, 64, 64, 0, 0);       // base_fc.0: f block (k 0..47) ...
    tc_fill(IMG(L_BF0C), 64, 80, W + GN_OFF(BF_WPC), 32, 64, 64, 48, 0);     // ... | (prob_embed block o prob_embed.2) (k 48..79)
#undef IMG
    float* small = reinterpret_cast<float*>(out + (size_t)TC_IMG_HALVES * 2);
    for (int e = 0; e < kTcSmallCount; ++e)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plan.n[e]; i += gridDim.x * blockDim.x)
            small[plan.dst[e] + i] = __ldg(W + plan.src[e] + i);
}

extern "C" int gn_k2a_tc_const_bytes(void) { return TC_CONST_BYTES; }

extern "C" int gn_k2a_tc_prepare(const float* weights, void* tc_const, void* stream)
{
    cudaError_t e = cudaMemsetAsync(tc_const, 0, TC_CONST_BYTES, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    TcSmallPlan plan;
    for (int i = 0; i < kTcSmallCount; ++i) { plan.src[i] = gn_w_off(kTcSmall[i]); plan.n[i] = gn_w_size(kTcSmall[i]); plan.dst[i] = ts_off_idx(i); }
    gn_k2a_tc_prepare_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(weights, reinterpret_cast<unsigned char*>(tc_const), plan);
    return (int)cudaGetLastError();
}

