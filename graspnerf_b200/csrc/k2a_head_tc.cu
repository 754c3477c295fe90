// K2a (tensor-core version): per-(point,view) head + cross-view pooling on tcgen05 / TMEM.
//
// Same reference op chain and same math as k2a_head_simt.cu (ibrnet.py:457-484,507-511; dist_decoder.py:99-142;
// aggregate_net.py:47-54), re-organised as a chain of small GEMMs D[128 x N] = A[128 x K] * W[N x K]^T:
//   * a tile is 128 rows = 4 warps x (G points x V views); thread t owns row t = TMEM lane t for the whole chain;
//   * the A operand of every layer lives in TENSOR MEMORY (tcgen05.mma "TS" form): the epilogue of layer i writes
//     the activated output straight back with tcgen05.st, it never touches shared memory;
//   * every fp32 operand is split into fp16 hi + lo (a = hi + lo to ~2^-22) and each product is three MMAs
//     (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM) - kind::f16, K = 16 per instruction;
//   * weights are converted once per CTA into K-major SWIZZLE_NONE fp16 images in shared memory (B operand);
//   * two tiles ("slots", 4 warps each) are in flight per CTA so one slot's MMAs run under the other slot's epilogue;
//   * cross-view poolings are warp-shuffle loops exactly as in the SIMT kernel.
// Operand layouts were validated on B200 with tools/tc_probe.cu (see profiles/tc_probe_r01.txt).
#include "gn_common.cuh"
#include "gn_weights.cuh"
#include "../../include/graspnerf_b200.h"
#include <cuda_fp16.h>

#define TC_THREADS 256
#define TC_SLOTS 2

// ---- TMEM column map per slot (256 columns each) -------------------------------------------------------------
#define TM_D 0            // accumulator, up to 96 columns
#define TM_AHI 96         // A operand hi halves, 72 columns (K <= 144)
#define TM_ALO 168        // A operand lo halves
#define TM_SLOT 256

// ---- shared-memory B images (fp16, element (n,k) at (k/8)*(N*8) + n*8 + k%8) ------------------------------------
struct TcLayer { int N, K; };
enum { L_DD1, L_DD2M, L_DD2V, L_DD2A, L_PE0, L_PE2, L_NF0, L_RD0, L_RD1, L_BF0A, L_BF0B, L_BF2, L_VF0, L_VF2, L_V20, L_COUNT };
__host__ __device__ constexpr TcLayer tc_layer(int i) {
    return i == L_DD1 ? TcLayer{96, 32} : i == L_DD2M ? TcLayer{32, 32} : i == L_DD2V ? TcLayer{32, 32} : i == L_DD2A ? TcLayer{32, 32}
         : i == L_PE0 ? TcLayer{32, 48} : i == L_PE2 ? TcLayer{32, 32} : i == L_NF0 ? TcLayer{16, 32} : i == L_RD0 ? TcLayer{16, 16}
         : i == L_RD1 ? TcLayer{48, 16} : i == L_BF0A ? TcLayer{64, 80} : i == L_BF0B ? TcLayer{64, 144} : i == L_BF2 ? TcLayer{32, 64}
         : i == L_VF0 ? TcLayer{32, 32} : i == L_VF2 ? TcLayer{48, 32} : TcLayer{32, 32};
}
__host__ __device__ constexpr int tc_img_off(int i) {          // offset in halves of the HI image; LO follows at +N*K
    int o = 0;
    for (int j = 0; j < i; ++j) o += 2 * tc_layer(j).N * tc_layer(j).K;
    return o;
}
constexpr int TC_IMG_HALVES = tc_img_off(L_COUNT);
// biases and the few CUDA-core weights (third dist-decoder layers, neuray_fc.2, vis_fc2.2, rgb_fc) are read straight from
// the fp32 blob in global memory: warp-uniform addresses, a couple of KB that stay in L1.
constexpr size_t TC_SMEM_BYTES = (size_t)TC_IMG_HALVES * 2 + 64;
static_assert(TC_SMEM_BYTES <= 227 * 1024, "K2a-TC shared memory budget");

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // tcgen05 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), SWIZZLE_NONE
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends the thread for a HW-bounded time per attempt; the attempt counter turns a lost arrival
    // (a bug) into a trap instead of a hung GPU.
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (it > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float* y) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(r[i]);
}
template <int N> __device__ __forceinline__ void tm_ld(uint32_t taddr, float* y) {
#pragma unroll
    for (int c = 0; c < N; c += 16) tm_ld16(taddr + c, y + c);
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// split K fp32 values into fp16 hi/lo pairs and store them as the A operand (k0 = first k index, multiple of 16)
template <int K> __device__ __forceinline__ void tm_store_a(uint32_t slot_lane_addr, int k0, const float* a) {
#pragma unroll
    for (int c = 0; c < K / 2; c += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a0 = a[2 * (c + i)], a1 = a[2 * (c + i) + 1];
            const __half2 h = __floats2half2_rn(a0, a1);                 // .x (low 16 bits) = even k  (tc_probe variant 0)
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
            hi[i] = *reinterpret_cast<const uint32_t*>(&h);
            lo[i] = *reinterpret_cast<const uint32_t*>(&l);
        }
        tm_st8(slot_lane_addr + TM_AHI + k0 / 2 + c, hi);
        tm_st8(slot_lane_addr + TM_ALO + k0 / 2 + c, lo);
    }
}

struct TcCtx {
    uint32_t tmem_slot;       // TMEM base of this slot (lane 0)
    uint32_t lane_addr;       // tmem_slot + (warp%4 * 32 << 16)
    uint32_t img_base;        // shared address of the image area
    uint32_t bar;             // shared address of this slot's mbarrier
    uint32_t parity;
    int bar_id;               // named barrier id of this slot
    bool leader;
};

// One GEMM (three fp16 passes).  Issued by the slot leader only.
template <int LAYER>
__device__ __forceinline__ void tc_issue(const TcCtx& cx, int d_col, int a_k0, bool accumulate) {
    constexpr int N = tc_layer(LAYER).N, K = tc_layer(LAYER).K;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F32 accum, F16 x F16, M=128
    const uint32_t bhi = cx.img_base + tc_img_off(LAYER) * 2, blo = bhi + N * K * 2;
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {                 // small terms first: lo*hi, hi*lo, hi*hi
        const uint32_t a_col = (pass == 0 ? TM_ALO : TM_AHI) + a_k0 / 2;
        const uint32_t b = (pass == 1) ? blo : bhi;
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            tc_mma(cx.tmem_slot + TM_D + d_col, cx.tmem_slot + a_col + ks * 8, tc_desc(b + ks * 2 * (N * 16), N * 16, 128), idesc, acc);
            acc = 1u;
        }
    }
}
// all 128 threads of the slot: A operand written -> (leader issues) -> accumulator ready
#define TC_GEMM_BEGIN(cx)                                                            \
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");                     \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                 \
    asm volatile("bar.sync %0, 128;" :: "r"((cx).bar_id) : "memory");                \
    if ((cx).leader) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#define TC_GEMM_END(cx)                                                              \
        tc_commit((cx).bar); }                                                       \
    mbar_wait((cx).bar, (cx).parity); (cx).parity ^= 1u;                             \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

// fp32 (k-major blob) -> fp16 hi/lo image block: rows [k_dst, k_dst+ksrc), cols [n_dst, n_dst+nsrc)
__device__ void tc_fill(__half* img, int N, int K, const float* __restrict__ src, int ksrc, int nsrc, int cp, int k_dst, int n_dst) {
    for (int i = threadIdx.x; i < ksrc * nsrc; i += TC_THREADS) {
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

template <int N> __device__ __forceinline__ void add_bias(const float* __restrict__ b, float* y) {
#pragma unroll
    for (int n = 0; n < N; n += 4) {
        const float4 w = *reinterpret_cast<const float4*>(b + n);
        y[n] += w.x; y[n + 1] += w.y; y[n + 2] += w.z; y[n + 3] += w.w;
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gn_k2a_tc_kernel(const __grid_constant__ GnK2aParams p, int num_tiles, int G)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __half* s_img = reinterpret_cast<__half*>(smem_raw);
    const float* __restrict__ sw = p.weights;                                              // fp32 blob (global, L1-resident constants)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TC_IMG_HALVES * 2);   // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + TC_SLOTS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = warp >> 2;
    // ---- one-time setup: fp32 blob copy, fp16 images, TMEM, mbarriers
    for (int i = tid; i < TC_IMG_HALVES / 2; i += TC_THREADS) reinterpret_cast<uint32_t*>(s_img)[i] = 0u;
    __syncthreads();
    {
        const float* W = p.weights;
#define IMG(L) (s_img + tc_img_off(L))
        tc_fill(IMG(L_DD1), 96, 32, W + GN_OFF(DD_MEAN_W0), 32, 32, 32, 0, 0);
        tc_fill(IMG(L_DD1), 96, 32, W + GN_OFF(DD_VAR_W0), 32, 32, 32, 0, 32);
        tc_fill(IMG(L_DD1), 96, 32, W + GN_OFF(DD_AW_W0), 32, 32, 32, 0, 64);
        tc_fill(IMG(L_DD2M), 32, 32, W + GN_OFF(DD_MEAN_W2), 32, 32, 32, 0, 0);
        tc_fill(IMG(L_DD2V), 32, 32, W + GN_OFF(DD_VAR_W2), 32, 32, 32, 0, 0);
        tc_fill(IMG(L_DD2A), 32, 32, W + GN_OFF(DD_AW_W2), 32, 32, 32, 0, 0);
        tc_fill(IMG(L_PE0), 32, 48, W + GN_OFF(PE_W0), 34, 32, 32, 0, 0);
        tc_fill(IMG(L_PE2), 32, 32, W + GN_OFF(PE_W2), 32, 32, 32, 0, 0);
        tc_fill(IMG(L_NF0), 16, 32, W + GN_OFF(NF_W0), 32, 8, 8, 0, 0);
        tc_fill(IMG(L_RD0), 16, 16, W + GN_OFF(RD_W0), 4, 16, 16, 0, 0);
        tc_fill(IMG(L_RD1), 48, 16, W + GN_OFF(RD_W1), 16, 36, 36, 0, 0);
        tc_fill(IMG(L_BF0A), 64, 80, W + GN_OFF(BF_WF), 36, 64, 64, 0, 0);
        tc_fill(IMG(L_BF0A), 64, 80, W + GN_OFF(BF_WP), 32, 64, 64, 48, 0);
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
#undef IMG
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to tcgen05.mma
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    TcCtx cx;
    cx.tmem_slot = *s_tmem + slot * TM_SLOT;
    cx.lane_addr = cx.tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    cx.img_base = smem_u32(s_img);
    cx.bar = smem_u32(&s_bar[slot]);
    cx.parity = 0u;
    cx.bar_id = 1 + slot;
    cx.leader = (tid & 127) == 0;

    const int V = p.V;
    const bool lane_active = lane < G * V;
    const int g = lane_active ? lane / V : 0;
    const int v = lane_active ? lane - g * V : 0;
    const int gb = g * V;
    const long long total_pts = (long long)p.B * p.N;
    const unsigned FULL = 0xffffffffu;

    for (int tile = blockIdx.x * TC_SLOTS + slot; tile < num_tiles; tile += gridDim.x * TC_SLOTS) {
        long long pidx = ((long long)tile * 4 + (warp & 3)) * G + g;
        const bool valid = lane_active && pidx < total_pts;
        pidx = pidx < total_pts ? pidx : total_pts - 1;
        const int b = (int)(pidx / p.N);
        const int n = (int)(pidx - (long long)b * p.N);
        const float* row = p.rec + ((size_t)pidx * V + v) * GN_REC_STRIDE;
        const float2 ptv = __ldg(reinterpret_cast<const float2*>(p.pt + (size_t)pidx * GN_PT_STRIDE));
        const float4 tail = ldg4(row + GN_REC_RGB);            // rgb0..2 (masked), depth
        const float4 ddv = ldg4(row + GN_REC_DD);
        const float mask = (valid && ((__float_as_uint(ptv.y) >> v) & 1u)) ? 1.f : 0.f;
        const float depth = tail.w;
        const float nvalid = ptv.x;
        const float wgt = __fdiv_rn(mask, nvalid + 1e-8f);      // ibrnet.py:466

        // ================= S1: dist-decoder first layers (3 x 32 -> 32 as one N=96 GEMM) =================
        {
            float ray[32];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = ldg4(row + GN_REC_RAYF + c);
                ray[c] = t.x; ray[c + 1] = t.y; ray[c + 2] = t.z; ray[c + 3] = t.w;
            }
            tm_store_a<32>(cx.lane_addr, 0, ray);               // A[k 0..31] = ray_feats (kept for S3)
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_DD1>(cx, 0, 0, false); TC_GEMM_END(cx)
        // ================= S2: second layers (block diagonal: three N=32,K=32 GEMMs) =====================
        {
            float h[32];
            tm_ld<32>(cx.lane_addr + TM_D + 0, h);  add_bias<32>(sw + GN_OFF(DD_MEAN_B0), h);
#pragma unroll
            for (int c = 0; c < 32; ++c) h[c] = gn_elu(h[c]);
            tm_store_a<32>(cx.lane_addr, 48, h);
            tm_ld<32>(cx.lane_addr + TM_D + 32, h); add_bias<32>(sw + GN_OFF(DD_VAR_B0), h);
#pragma unroll
            for (int c = 0; c < 32; ++c) h[c] = gn_elu(h[c]);
            tm_store_a<32>(cx.lane_addr, 80, h);
            tm_ld<32>(cx.lane_addr + TM_D + 64, h); add_bias<32>(sw + GN_OFF(DD_AW_B0), h);
#pragma unroll
            for (int c = 0; c < 32; ++c) h[c] = gn_elu(h[c]);
            tm_store_a<32>(cx.lane_addr, 112, h);
        }
        TC_GEMM_BEGIN(cx)
            tc_issue<L_DD2M>(cx, 0, 48, false); tc_issue<L_DD2V>(cx, 32, 80, false); tc_issue<L_DD2A>(cx, 64, 112, false);
        TC_GEMM_END(cx)
        // ================= third layers on CUDA cores, compute_prob (dist_decoder.py:109-142) ============
        float hit, vis;
        {
            float h[32], om[4], ov[4], oa[4];
            tm_ld<32>(cx.lane_addr + TM_D + 0, h);  add_bias<32>(sw + GN_OFF(DD_MEAN_B2), h);
#pragma unroll
            for (int c = 0; c < 32; ++c) h[c] = gn_elu(h[c]);
#pragma unroll
            for (int q = 0; q < 4; ++q) om[q] = sw[GN_OFF(DD_MEAN_B4) + q];
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float4 w = *reinterpret_cast<const float4*>(sw + GN_OFF(DD_MEAN_W4) + k * 4); om[0] = fmaf(h[k], w.x, om[0]); om[1] = fmaf(h[k], w.y, om[1]); }
            tm_ld<32>(cx.lane_addr + TM_D + 32, h); add_bias<32>(sw + GN_OFF(DD_VAR_B2), h);
#pragma unroll
            for (int c = 0; c < 32; ++c) h[c] = gn_elu(h[c]);
#pragma unroll
            for (int q = 0; q < 4; ++q) ov[q] = sw[GN_OFF(DD_VAR_B4) + q];
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float4 w = *reinterpret_cast<const float4*>(sw + GN_OFF(DD_VAR_W4) + k * 4); ov[0] = fmaf(h[k], w.x, ov[0]); ov[1] = fmaf(h[k], w.y, ov[1]); }
            tm_ld<32>(cx.lane_addr + TM_D + 64, h); add_bias<32>(sw + GN_OFF(DD_AW_B2), h);
#pragma unroll
            for (int c = 0; c < 32; ++c) h[c] = gn_elu(h[c]);
            oa[0] = sw[GN_OFF(DD_AW_B4)];
#pragma unroll
            for (int k = 0; k < 32; ++k) oa[0] = fmaf(h[k], sw[GN_OFF(DD_AW_W4) + k * 4], oa[0]);
            const float mean0 = gn_softplus(om[0]), mean1 = gn_softplus(om[1]);
            const float var0 = gn_softplus(ov[0]) + 0.05f, var1 = gn_softplus(ov[1]) + 0.05f;
            const float aw = gn_sigmoid(oa[0]);
            const float* dr = p.depth_range + ((size_t)b * V + v) * 2;
            const float rnear = __fdiv_rn(-1.f, __ldg(dr)), rfar = __fdiv_rn(-1.f, __ldg(dr + 1));
            float d = __fdiv_rn(-1.f, fmaxf(depth, 1e-5f));
            d = __fdiv_rn(d - rnear, rfar - rnear);
            float nearp, farp;
            if (p.que_dists == nullptr) { nearp = d - 0.005f; farp = d + 0.005f; }
            else {
                const int smp = n % p.dn;
                const float* qd = p.que_dists + (size_t)b * p.N + n;
                const float h_cur = __ldg(qd) * 0.5f;
                const float h_prev = smp > 0 ? __ldg(qd - 1) * 0.5f : h_cur;
                nearp = d - h_prev; farp = d + h_cur;
            }
            const float c00 = 0.5f + 0.5f * tanhf((nearp - mean0) * var0), c10 = 0.5f + 0.5f * tanhf((farp - mean0) * var0);
            const float c01 = 0.5f + 0.5f * tanhf((nearp - mean1) * var1), c11 = 0.5f + 0.5f * tanhf((farp - mean1) * var1);
            const float mix1 = 1.f - aw;
            vis = ((1.f - c00) * aw + (1.f - c01) * mix1) * mask;
            hit = ((c10 - c00) * aw + (c11 - c01) * mix1) * mask;
        }
        // ================= S3: prob_embed.0 on [ray | 2hit-1 | 2vis-1]  (K = 34 -> 48) ====================
        {
            float hv[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) hv[c] = 0.f;
            hv[0] = (hit - 0.5f) * 2.f; hv[1] = (vis - 0.5f) * 2.f;
            tm_store_a<16>(cx.lane_addr, 32, hv);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_PE0>(cx, 0, 0, false); TC_GEMM_END(cx)
        // ================= S4: prob_embed.2 ===============================================================
        {
            float e1[32];
            tm_ld<32>(cx.lane_addr + TM_D, e1); add_bias<32>(sw + GN_OFF(PE_B0), e1);
#pragma unroll
            for (int c = 0; c < 32; ++c) e1[c] = fmaxf(e1[c], 0.f);
            tm_store_a<32>(cx.lane_addr, 80, e1);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_PE2>(cx, 0, 80, false); TC_GEMM_END(cx)
        // ================= S5: neuray_fc.0 on prob_emb, ray_dir_fc.0 on dir_diff ==========================
        float pe01[2];
        {
            float pe[32];
            tm_ld<32>(cx.lane_addr + TM_D, pe); add_bias<32>(sw + GN_OFF(PE_B2), pe);
            pe01[0] = pe[0]; pe01[1] = pe[1];
            tm_store_a<32>(cx.lane_addr, 48, pe);               // stays at k 48..79 for base_fc (S7a)
            float dd[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) dd[c] = 0.f;
            dd[0] = ddv.x; dd[1] = ddv.y; dd[2] = ddv.z; dd[3] = ddv.w;
            tm_store_a<16>(cx.lane_addr, 112, dd);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_NF0>(cx, 0, 48, false); tc_issue<L_RD0>(cx, 16, 112, false); TC_GEMM_END(cx)
        // ================= S6: ray_dir_fc.2 ; weight0 =======================================================
        float w0;
        {
            float t[32];
            tm_ld<32>(cx.lane_addr + TM_D, t);
            float s = sw[GN_OFF(NF_B2)];
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(gn_elu(t[k] + sw[GN_OFF(NF_B0) + k]), sw[GN_OFF(NF_W2) + k], s);
            w0 = gn_sigmoid(s) * wgt;                           // ibrnet.py:469
            float hid[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) hid[k] = gn_elu(t[16 + k] + sw[GN_OFF(RD_B0) + k]);
            tm_store_a<16>(cx.lane_addr, 128, hid);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_RD1>(cx, 0, 128, false); TC_GEMM_END(cx)
        // ================= f = feats + direction feature; mean/var poolings; S7a ==========================
        float g0[36], g1[36];
        {
            float f[48];
            tm_ld<48>(cx.lane_addr + TM_D, f); add_bias<36>(sw + GN_OFF(RD_B1), f);
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = ldg4(row + GN_REC_IMGF + c);
                f[c] = gn_elu(f[c]) + t.x; f[c + 1] = gn_elu(f[c + 1]) + t.y; f[c + 2] = gn_elu(f[c + 2]) + t.z; f[c + 3] = gn_elu(f[c + 3]) + t.w;
            }
            f[32] = gn_elu(f[32]) + tail.x; f[33] = gn_elu(f[33]) + tail.y; f[34] = gn_elu(f[34]) + tail.z;
#pragma unroll
            for (int c = 35; c < 48; ++c) f[c] = 0.f;
            tm_store_a<48>(cx.lane_addr, 0, f);                 // A[k 0..47] = f (ray_feats no longer needed)
#pragma unroll
            for (int c = 0; c < 35; ++c) {                      // ibrnet.py:470-471 means
                const float t0 = w0 * f[c], t1 = wgt * f[c];
                float s0 = 0.f, s1 = 0.f;
                for (int jv = 0; jv < V; ++jv) { s0 += __shfl_sync(FULL, t0, (gb + jv) & 31); s1 += __shfl_sync(FULL, t1, (gb + jv) & 31); }
                g0[c] = s0; g1[c] = s1;
            }
            g0[35] = 0.f; g1[35] = 0.f;
            TC_GEMM_BEGIN(cx) tc_issue<L_BF0A>(cx, 0, 0, false); TC_GEMM_END(cx)
            // S7b operand, k layout: mean0[0..31] | mean1[0..31] | var0[0..31] | var1[0..31] | tails (channels 32..34 of the four)
            tm_store_a<32>(cx.lane_addr, 0, g0);
            tm_store_a<32>(cx.lane_addr, 32, g1);
            float tl[16];
#pragma unroll
            for (int c = 0; c < 3; ++c) { tl[c] = g0[32 + c]; tl[3 + c] = g1[32 + c]; }
#pragma unroll
            for (int c = 0; c < 35; ++c) {                      // variances in place (ibrnet.py:115)
                const float d0 = f[c] - g0[c], d1 = f[c] - g1[c];
                const float t0 = w0 * d0 * d0, t1 = wgt * d1 * d1;
                float s0 = 0.f, s1 = 0.f;
                for (int jv = 0; jv < V; ++jv) { s0 += __shfl_sync(FULL, t0, (gb + jv) & 31); s1 += __shfl_sync(FULL, t1, (gb + jv) & 31); }
                g0[c] = s0; g1[c] = s1;
            }
            tm_store_a<32>(cx.lane_addr, 64, g0);
            tm_store_a<32>(cx.lane_addr, 96, g1);
#pragma unroll
            for (int c = 0; c < 3; ++c) { tl[6 + c] = g0[32 + c]; tl[9 + c] = g1[32 + c]; }
            tl[12] = 0.f; tl[13] = 0.f; tl[14] = 0.f; tl[15] = 0.f;
            tm_store_a<16>(cx.lane_addr, 128, tl);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_BF0B>(cx, 0, 0, true); TC_GEMM_END(cx)
        // ================= S8: base_fc.2 ====================================================================
        {
            float y[64];
            tm_ld<64>(cx.lane_addr + TM_D, y); add_bias<64>(sw + GN_OFF(BF_B0), y);
#pragma unroll
            for (int c = 0; c < 64; ++c) y[c] = gn_elu(y[c]);
            tm_store_a<64>(cx.lane_addr, 0, y);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_BF2>(cx, 0, 0, false); TC_GEMM_END(cx)
        // ================= S9/S10: vis_fc ====================================================================
        float x[32];
        tm_ld<32>(cx.lane_addr + TM_D, x); add_bias<32>(sw + GN_OFF(BF_B2), x);
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = gn_elu(x[c]);
        {
            float xi[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) xi[c] = x[c] * wgt;
            tm_store_a<32>(cx.lane_addr, 64, xi);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_VF0>(cx, 0, 64, false); TC_GEMM_END(cx)
        {
            float t[32];
            tm_ld<32>(cx.lane_addr + TM_D, t); add_bias<32>(sw + GN_OFF(VF_B0), t);
#pragma unroll
            for (int c = 0; c < 32; ++c) t[c] = gn_elu(t[c]);
            tm_store_a<32>(cx.lane_addr, 96, t);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_VF2>(cx, 0, 96, false); TC_GEMM_END(cx)
        float visw;
        {
            float xv[48];
            tm_ld<48>(cx.lane_addr + TM_D, xv); add_bias<36>(sw + GN_OFF(VF_B2), xv);
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] += gn_elu(xv[c]);
            visw = gn_sigmoid(gn_elu(xv[32])) * mask;           // ibrnet.py:478-479
            float xi[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) xi[c] = x[c] * visw;
            tm_store_a<32>(cx.lane_addr, 64, xi);
        }
        // ================= S11: vis_fc2 =======================================================================
        TC_GEMM_BEGIN(cx) tc_issue<L_V20>(cx, 0, 64, false); TC_GEMM_END(cx)
        float vis2;
        {
            float t[32];
            tm_ld<32>(cx.lane_addr + TM_D, t); add_bias<32>(sw + GN_OFF(V2_B0), t);
            float s = sw[GN_OFF(V2_B2)];
#pragma unroll
            for (int k = 0; k < 32; ++k) s = fmaf(gn_elu(t[k]), sw[GN_OFF(V2_W2) + k], s);
            vis2 = gn_sigmoid(s) * mask;
        }
        // ================= final pooling (ibrnet.py:482-484,487) ===============================================
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
            st4(dr + 4, make_float4(x[0], x[1], pe01[0], pe01[1]));
        }
        // ================= rgb_fc + masked softmax over views (ibrnet.py:507-511), CUDA cores ================
        if (p.with_rgb && p.colors) {
            float r16[16], r8[8];
#pragma unroll
            for (int c = 0; c < 16; ++c) r16[c] = sw[GN_OFF(RF_B0) + c];
            const float dd4[5] = { vis2, ddv.x, ddv.y, ddv.z, ddv.w };
#pragma unroll
            for (int k = 0; k < 37; ++k) {
                const float xk = k < 32 ? x[k] : dd4[k - 32];
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(sw + GN_OFF(RF_W0) + k * 16 + c);
                    r16[c] = fmaf(xk, w.x, r16[c]); r16[c + 1] = fmaf(xk, w.y, r16[c + 1]); r16[c + 2] = fmaf(xk, w.z, r16[c + 2]); r16[c + 3] = fmaf(xk, w.w, r16[c + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) r8[c] = sw[GN_OFF(RF_B2) + c];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float xk = gn_elu(r16[k]);
#pragma unroll
                for (int c = 0; c < 8; ++c) r8[c] = fmaf(xk, sw[GN_OFF(RF_W2) + k * 8 + c], r8[c]);
            }
            float logit = sw[GN_OFF(RF_B4)];
#pragma unroll
            for (int k = 0; k < 8; ++k) logit = fmaf(gn_elu(r8[k]), sw[GN_OFF(RF_W4) + k], logit);
            if (mask == 0.f) logit = -1e9f;
            float mx = -INFINITY;
            for (int jv = 0; jv < V; ++jv) mx = fmaxf(mx, __shfl_sync(FULL, logit, (gb + jv) & 31));
            const float e = __expf(logit - mx);
            float es = 0.f;
            for (int jv = 0; jv < V; ++jv) es += __shfl_sync(FULL, e, (gb + jv) & 31);
            const float bw = __fdiv_rn(e, es);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f;
            for (int jv = 0; jv < V; ++jv) {
                c0 += __shfl_sync(FULL, bw * tail.x, (gb + jv) & 31);
                c1 += __shfl_sync(FULL, bw * tail.y, (gb + jv) & 31);
                c2 += __shfl_sync(FULL, bw * tail.z, (gb + jv) & 31);
            }
            if (writer) st4(p.colors + (size_t)pidx * 4, make_float4(c0, c1, c2, 0.f));
        }
    }
    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*s_tmem), "r"(512));
}

extern "C" int gn_k2a_forward_tc(const GnK2aParams* hp, void* stream)
{
    const GnK2aParams& p = *hp;
    if (p.V < 1 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (p.que_dists && (p.dn < 1 || (p.N % p.dn) != 0)) return -4;
    const int G = 32 / p.V;
    const long long total = (long long)p.B * p.N;
    const long long per_tile = 4LL * G;
    const long long tiles = (total + per_tile - 1) / per_tile;
    if (tiles > 0x7fffffffLL) return -6;
    cudaError_t e = cudaFuncSetAttribute(gn_k2a_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (tiles + TC_SLOTS - 1) / TC_SLOTS;
    const int grid = (int)(want < sms ? want : sms);
    gn_k2a_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(p, (int)tiles, G);
    return (int)cudaGetLastError();
}
