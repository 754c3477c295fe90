// K2a (tensor-core version): per-(point,view) head + cross-view pooling (+ geometry_fc) on tcgen05 / TMEM.
//
// Same reference op chain and same math as k2a_head_simt.cu (ibrnet.py:457-489,507-511; dist_decoder.py:99-142;
// aggregate_net.py:47-54), re-organised as a chain of small GEMMs D[128 x N] = A[128 x K] * W[N x K]^T:
//   * a tile is 128 rows = 4 warps x (G points x V views); thread t owns row t = TMEM lane t for the whole chain;
//   * the A operand of every layer lives in TENSOR MEMORY (tcgen05.mma "TS" form): the epilogue of layer i writes
//     the activated output straight back with tcgen05.st, it never touches shared memory;
//   * every fp32 operand is split into fp16 hi + lo (a = hi + lo to ~2^-22) and each product is three MMAs
//     (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM) - kind::f16, K = 16 per instruction;
//   * weights: K-major SWIZZLE_NONE fp16 hi/lo images + the small fp32 constants, prepared ONCE per weight update by
//     gn_k2a_tc_prepare_kernel into a global buffer that each CTA copies verbatim into shared memory;
//   * two tiles ("slots", 4 warps each) are in flight per CTA so one slot's MMAs run under the other slot's epilogue;
//   * cross-view poolings go through a per-warp shared-memory scratch (write 36 values, V-strided partial sums, read back);
//   * optionally (tok != NULL) geometry_fc (ibrnet.py:487-489: 86 -> 64 -> 16 on [mean, var, mean_v(w), embed(pts)]) runs
//     as two more GEMMs on the pooled rows, so the per-ray kernel K2b only does attention + LayerNorm + output MLP.
// Operand layouts were validated on B200 with tools/tc_probe.cu (profiles/tc_probe_r01.txt).
#include "k2a_tc_common.cuh"
#include <cstdlib>

#define TC_THREADS 256
#define TC_SLOTS 2

#define TC_POOL_STRIDE 44                                                     // floats per row of the pooling scratch
__host__ __device__ constexpr size_t tc_smem_bytes(int G) {
    return (size_t)TC_CONST_BYTES + (size_t)(TC_THREADS / 32) * (32 + G) * TC_POOL_STRIDE * 4 + 64;   // + barriers, tmem ptr
}
static_assert(tc_smem_bytes(5) <= 227 * 1024, "K2a-TC shared memory budget at V = 6");

// Generic layer epilogue, shared by every "plain" layer (NOT inlined: one copy of the code serves ~60 % of all
// activations, which keeps the kernel's instruction footprint cache-friendly):
//   A[k0 + 0 .. 16*nchunk) = act(D[d_col + 0 .. 16*nchunk) + bias),  ACT 0 none, 1 ELU, 2 ReLU.
// The tcgen05.ld of chunk c+1 is in flight while chunk c is activated, split and stored.
template <int ACT>
__device__ __noinline__ void tc_epilogue(uint32_t lane_addr, int d_col, int nchunk, const float* __restrict__ bias, int k0)
{
    uint32_t r[16];
    tm_ld16_issue(lane_addr + TM_D + d_col, r);
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tm_ld16_fence(r);
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 w = *reinterpret_cast<const float4*>(bias + c * 16 + i);
            upk2(add2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), pk2(w.x, w.y)), y[i], y[i + 1]);
            upk2(add2(pk2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), pk2(w.z, w.w)), y[i + 2], y[i + 3]);
        }
        if (c + 1 < nchunk) tm_ld16_issue(lane_addr + TM_D + d_col + (c + 1) * 16, r);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (ACT == 1) tc_elu2(y[i], y[i + 1]);
            else if (ACT == 2) { y[i] = fmaxf(y[i], 0.f); y[i + 1] = fmaxf(y[i + 1], 0.f); }
        }
        tm_store_a<16>(lane_addr, k0 + c * 16, y);
    }
}

// Cross-view sums of the 36 values every lane has parked in its scratch row: sums row of group g <- sum over its V rows.
// (summation order v = 0..V-1, like the shuffle loops of the SIMT kernel)
__device__ __noinline__ void tc_pool_rows(float* scr, int g, int v, int gb, int V, bool lane_active)
{
    __syncwarp();
    float* sums = scr + (32 + g) * TC_POOL_STRIDE;
    const float* base = scr + gb * TC_POOL_STRIDE;
#pragma unroll 1
    for (int ch = v; ch < 9; ch += V) {                 // lane (g,v) sums the float4 chunks v, v+V, ... of its point's V rows
        const float* q = base + 4 * ch;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int jv = 0; jv < V; ++jv) {
            const float4 t = *reinterpret_cast<const float4*>(q + jv * TC_POOL_STRIDE);
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        if (lane_active) st4(sums + 4 * ch, s);
    }
    __syncwarp();
}

struct TcCtx {
    uint32_t tmem_slot;       // TMEM base of this slot (lane 0)
    uint32_t lane_addr;       // tmem_slot + (warp%4 * 32 << 16)
    uint32_t img_base16;      // (shared address of the image area) >> 4
    uint32_t bar;             // shared address of this slot's mbarrier
    uint32_t parity;
    int bar_id;               // named barrier id of this slot
    bool issuer;              // this warp issues the slot's MMAs (warp-uniform)
    uint32_t elected;         // 1 on the one lane of the issuing warp that actually issues
};

// One GEMM (three fp16 passes).  Executed by the slot's issuing warp; one elected lane issues.
template <int LAYER>
__device__ __forceinline__ void tc_issue(const TcCtx& cx, int d_col, int a_k0, bool accumulate) {
    constexpr int N = tc_layer(LAYER).N, K = tc_layer(LAYER).K;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F32 accum, F16 x F16, M=128
    // low descriptor word = (start address >> 4) | (LBO >> 4) << 16 with LBO = N*16 B (k-chunk stride); the start address of
    // every image / k-step is the (runtime) image base plus a compile-time constant, so each MMA costs one add
    constexpr uint32_t lbo_field = (uint32_t)((N * 16) >> 4) << 16;
    constexpr uint32_t hi_off = (uint32_t)(tc_img_off(LAYER) * 2) >> 4, lo_off = hi_off + (uint32_t)((N * K * 2) >> 4);
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {                 // small terms first: lo*hi, hi*lo, hi*hi
        const uint32_t a_col = (pass == 0 ? TM_ALO : TM_AHI) + a_k0 / 2;
        const uint32_t boff = (pass == 1) ? lo_off : hi_off;
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            tc_mma(cx.tmem_slot + TM_D + d_col, cx.tmem_slot + a_col + ks * 8,
                   cx.img_base16 + (boff + (uint32_t)(ks * 2 * N) + lbo_field), idesc, acc, cx.elected);
            acc = 1u;
        }
    }
}
// all 128 threads of the slot: A operand written -> (leader issues) -> accumulator ready
#define TC_GEMM_BEGIN(cx)                                                            \
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");                     \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                 \
    asm volatile("bar.sync %0, 128;" :: "r"((cx).bar_id) : "memory");                \
    if ((cx).issuer) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#define TC_GEMM_COMMIT(cx)                                                           \
        tc_commit((cx).bar, (cx).elected); __syncwarp(); }
#define TC_GEMM_WAIT(cx)                                                             \
    mbar_wait((cx).bar, (cx).parity); (cx).parity ^= 1u;                             \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#define TC_GEMM_END(cx) TC_GEMM_COMMIT(cx) TC_GEMM_WAIT(cx)

// Cross-view sum of 36 per-row values through the warp's scratch: out[c] = sum over the V rows of my point of vals[c].
__device__ __forceinline__ void pool36(float* scr, int lane, int g, int v, int gb, int V, bool lane_active, const float* vals, float* out)
{
    float* mine = scr + lane * TC_POOL_STRIDE;
#pragma unroll
    for (int c = 0; c < 36; c += 4) st4(mine + c, make_float4(vals[c], vals[c + 1], vals[c + 2], vals[c + 3]));
    tc_pool_rows(scr, g, v, gb, V, lane_active);
    const float* sums = scr + (32 + g) * TC_POOL_STRIDE;
#pragma unroll
    for (int c = 0; c < 36; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(sums + c);
        out[c] = t.x; out[c + 1] = t.y; out[c + 2] = t.z; out[c + 3] = t.w;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gn_k2a_tc_kernel(const __grid_constant__ GnK2aParams p, int num_tiles, int G, int nslots)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __half* s_img = reinterpret_cast<__half*>(smem_raw);
    const float* sw = reinterpret_cast<const float*>(smem_raw + (size_t)TC_IMG_HALVES * 2);   // small fp32 constants, index with TS()
    float* s_pool = reinterpret_cast<float*>(smem_raw + TC_CONST_BYTES);                       // [8 warps][(32+G)][44]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pool + (TC_THREADS / 32) * (32 + G) * TC_POOL_STRIDE);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + TC_SLOTS);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction
    const int slot = warp >> 2;
    // ---- one-time setup: constants (TMA bulk copy global -> shared, completion on an mbarrier), TMEM, mbarriers
    uint64_t* s_cbar = reinterpret_cast<uint64_t*>(s_tmem + 2);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(s_cbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(s_cbar)), "r"((uint32_t)TC_CONST_BYTES) : "memory");
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.tc_const);
        for (uint32_t off = 0; off < (uint32_t)TC_CONST_BYTES; off += 32768u) {
            const uint32_t n = min(32768u, (uint32_t)TC_CONST_BYTES - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(smem_raw + off)), "l"(src + off), "r"(n), "r"(smem_u32(s_cbar)) : "memory");
        }
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(smem_u32(s_cbar), 0u);                                  // operand images + constants have landed (async proxy)

    TcCtx cx;
    cx.tmem_slot = *s_tmem + slot * TM_SLOT;
    cx.lane_addr = cx.tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    cx.img_base16 = smem_u32(s_img) >> 4;
    cx.bar = smem_u32(&s_bar[slot]);
    cx.parity = 0u;
    cx.bar_id = 1 + slot;
    cx.issuer = (warp & 3) == 0;
    cx.elected = cx.issuer ? elect_one() : 0u;

    const int V = p.V;
    const bool lane_active = lane < G * V;
    const int g = lane_active ? lane / V : 0;
    const int v = lane_active ? lane - g * V : 0;
    const int gb = g * V;
    const long long total_pts = (long long)p.B * p.N;
    const unsigned FULL = 0xffffffffu;
    float* scr = s_pool + warp * (32 + G) * TC_POOL_STRIDE;

    // nslots = TC_SLOTS in production; nslots = 1 (GN_K2A_SLOTS=1, measurement aid) leaves the second slot's warps idle
    for (int tile = blockIdx.x * nslots + slot; slot < nslots && tile < num_tiles; tile += gridDim.x * nslots) {
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

        // ================= S1: dist-decoder first layers (N = 96) + ray_dir_fc.0 on dir_diff (N = 16) ====================
        {
            float ray[32];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = ldg4(row + GN_REC_RAYF + c);
                ray[c] = t.x; ray[c + 1] = t.y; ray[c + 2] = t.z; ray[c + 3] = t.w;
            }
            tm_store_a<32>(cx.lane_addr, 0, ray);               // A[k 0..31] = ray_feats (kept for S3)
            float dd[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) dd[c] = 0.f;
            dd[0] = ddv.x; dd[1] = ddv.y; dd[2] = ddv.z; dd[3] = ddv.w;
            tm_store_a<16>(cx.lane_addr, 112, dd);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_DD1>(cx, 0, 0, false); tc_issue<L_RD0>(cx, 96, 112, false); TC_GEMM_END(cx)
        // ================= S2: dist-decoder second layers (block diagonal: three N=32,K=32 GEMMs) ========================
        static_assert(TS(DD_VAR_B0) == TS(DD_MEAN_B0) + 32 && TS(DD_AW_B0) == TS(DD_MEAN_B0) + 64, "dist-decoder biases must be contiguous");
        tc_epilogue<1>(cx.lane_addr, 0, 6, sw + TS(DD_MEAN_B0), 48);      // D[0..95] -> ELU -> A[k 48..143]  (dir_diff no longer needed)
        TC_GEMM_BEGIN(cx)
            tc_issue<L_DD2M>(cx, 0, 48, false); tc_issue<L_DD2V>(cx, 32, 80, false); tc_issue<L_DD2A>(cx, 64, 112, false);
        TC_GEMM_END(cx)
        // ================= third layers on CUDA cores, compute_prob (dist_decoder.py:109-142) ============
        float hit, vis;
        {
            float h[32], om0, om1, ov0, ov1, oa;
            tm_ld<32>(cx.lane_addr + TM_D + 0, h);  bias_elu<32>(sw + TS(DD_MEAN_B2), h);
            om0 = sw[TS(DD_MEAN_B4)]; om1 = sw[TS(DD_MEAN_B4) + 1];
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float2 w = *reinterpret_cast<const float2*>(sw + TS(DD_MEAN_W4) + k * 4); om0 = fmaf(h[k], w.x, om0); om1 = fmaf(h[k], w.y, om1); }
            tm_ld<32>(cx.lane_addr + TM_D + 32, h); bias_elu<32>(sw + TS(DD_VAR_B2), h);
            ov0 = sw[TS(DD_VAR_B4)]; ov1 = sw[TS(DD_VAR_B4) + 1];
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float2 w = *reinterpret_cast<const float2*>(sw + TS(DD_VAR_W4) + k * 4); ov0 = fmaf(h[k], w.x, ov0); ov1 = fmaf(h[k], w.y, ov1); }
            tm_ld<32>(cx.lane_addr + TM_D + 64, h); bias_elu<32>(sw + TS(DD_AW_B2), h);
            oa = sw[TS(DD_AW_B4)];
#pragma unroll
            for (int k = 0; k < 32; ++k) oa = fmaf(h[k], sw[TS(DD_AW_W4) + k * 4], oa);
            const float mean0 = gn_softplus(om0), mean1 = gn_softplus(om1);
            const float var0 = gn_softplus(ov0) + 0.05f, var1 = gn_softplus(ov1) + 0.05f;
            const float aw = gn_sigmoid(oa);
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
            // 0.5 + 0.5*tanh(d) == sigmoid(2d)   (dist_decoder.py:129-130)
            const float c00 = gn_sigmoid(2.f * ((nearp - mean0) * var0)), c10 = gn_sigmoid(2.f * ((farp - mean0) * var0));
            const float c01 = gn_sigmoid(2.f * ((nearp - mean1) * var1)), c11 = gn_sigmoid(2.f * ((farp - mean1) * var1));
            const float mix1 = 1.f - aw;
            vis = ((1.f - c00) * aw + (1.f - c01) * mix1) * mask;
            hit = ((c10 - c00) * aw + (c11 - c01) * mix1) * mask;
        }
        asm volatile("prefetch.global.L1 [%0];" :: "l"(row + GN_REC_IMGF));      // needed after S3
        // ================= S3: prob_embed.0 on [ray | 2hit-1 | 2vis-1] (K = 34 -> 48)  +  ray_dir_fc.2 (16 -> 35) ==========
        {
            float hv[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) hv[c] = 0.f;
            hv[0] = (hit - 0.5f) * 2.f; hv[1] = (vis - 0.5f) * 2.f;
            tm_store_a<16>(cx.lane_addr, 32, hv);
        }
        tc_epilogue<1>(cx.lane_addr, 96, 1, sw + TS(RD_B0), 128);        // ray_dir_fc hidden: D[96..111] -> ELU -> A[k 128..143]
        TC_GEMM_BEGIN(cx) tc_issue<L_PE0>(cx, 0, 0, false); tc_issue<L_RD1>(cx, 32, 128, false); TC_GEMM_END(cx)
        // ================= S4: {neuray_fc.0 o prob_embed.2} and base_fc.0's per-view part, both on [f | e1] ===================
        // prob_embed.2 has no activation, so its two consumers are pre-multiplied on the host (weights.py) and read the
        // ReLU'd hidden e1 directly: prob_embed itself is never materialised.
        tc_epilogue<2>(cx.lane_addr, 0, 2, sw + TS(PE_B0), 48);           // e1 = ReLU(.) -> A[k 48..79]
        float w0;
        {
            float f[48], g0[36], g1[36], tmp[36];
            tm_ld<48>(cx.lane_addr + TM_D + 32, f); bias_elu<36>(sw + TS(RD_B1), f);
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = ldg4(row + GN_REC_IMGF + c);
                f[c] += t.x; f[c + 1] += t.y; f[c + 2] += t.z; f[c + 3] += t.w;     // ibrnet.py:459
            }
            f[32] += tail.x; f[33] += tail.y; f[34] += tail.z;
#pragma unroll
            for (int c = 35; c < 48; ++c) f[c] = 0.f;
            tm_store_a<48>(cx.lane_addr, 0, f);                 // A[k 0..47] = f (ray_feats no longer needed)
            TC_GEMM_BEGIN(cx) tc_issue<L_NFC>(cx, 96, 48, false); tc_issue<L_BF0C>(cx, 0, 0, false); TC_GEMM_COMMIT(cx)
            // mean1 (weights w = mask / sum mask) does not need weight0: pooled while the MMAs run (shared memory only)
#pragma unroll
            for (int c = 0; c < 36; ++c) tmp[c] = wgt * f[c];
            pool36(scr, lane, g, v, gb, V, lane_active, tmp, g1);
            TC_GEMM_WAIT(cx)
            {
                float t[16];
                tm_ld<16>(cx.lane_addr + TM_D + 96, t);
                float s = sw[TS(NF_B2)];
#pragma unroll
                for (int k = 0; k < 8; ++k) s = fmaf(tc_elu(t[k] + sw[TS(NFC_B0) + k]), sw[TS(NF_W2) + k], s);
                w0 = gn_sigmoid(s) * wgt;                       // ibrnet.py:469
            }
#pragma unroll
            for (int c = 0; c < 36; ++c) tmp[c] = w0 * f[c];
            pool36(scr, lane, g, v, gb, V, lane_active, tmp, g0);
            // S7b operand, k layout: mean0[0..31] | mean1[0..31] | var0[0..31] | var1[0..31] | tails (channels 32..34 of the four)
            tm_store_a<32>(cx.lane_addr, 0, g0);
            tm_store_a<32>(cx.lane_addr, 32, g1);
            float tl[16];
#pragma unroll
            for (int c = 0; c < 3; ++c) { tl[c] = g0[32 + c]; tl[3 + c] = g1[32 + c]; }
#pragma unroll
            for (int c = 0; c < 36; ++c) { const float d0 = f[c] - g0[c]; tmp[c] = w0 * d0 * d0; }      // ibrnet.py:115
            pool36(scr, lane, g, v, gb, V, lane_active, tmp, g0);
#pragma unroll
            for (int c = 0; c < 36; ++c) { const float d1 = f[c] - g1[c]; tmp[c] = wgt * d1 * d1; }
            pool36(scr, lane, g, v, gb, V, lane_active, tmp, g1);
            tm_store_a<32>(cx.lane_addr, 64, g0);
            tm_store_a<32>(cx.lane_addr, 96, g1);
#pragma unroll
            for (int c = 0; c < 3; ++c) { tl[6 + c] = g0[32 + c]; tl[9 + c] = g1[32 + c]; }
            tl[12] = 0.f; tl[13] = 0.f; tl[14] = 0.f; tl[15] = 0.f;
            tm_store_a<16>(cx.lane_addr, 128, tl);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_BF0B>(cx, 0, 0, true); TC_GEMM_END(cx)
        {   // next tile's first record line (ray_feats) and per-point word: hide their latency under S8..S11
            const long long npidx = ((long long)(tile + gridDim.x * nslots) * 4 + (warp & 3)) * G + g;
            if (npidx < total_pts) {
                asm volatile("prefetch.global.L1 [%0];" :: "l"(p.rec + ((size_t)npidx * V + v) * GN_REC_STRIDE));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(p.rec + ((size_t)npidx * V + v) * GN_REC_STRIDE + GN_REC_RGB));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(p.pt + (size_t)npidx * GN_PT_STRIDE));
            }
        }
        // ================= S8: base_fc.2 ====================================================================
        tc_epilogue<1>(cx.lane_addr, 0, 4, sw + TS(BF_B0C), 0);          // bias includes the folded prob_embed.2 bias
        TC_GEMM_BEGIN(cx) tc_issue<L_BF2>(cx, 0, 0, false); TC_GEMM_END(cx)
        // ================= S9/S10: vis_fc ====================================================================
        float x[36];
        tm_ld<32>(cx.lane_addr + TM_D, x); bias_elu<32>(sw + TS(BF_B2), x);
        x[32] = 0.f; x[33] = 0.f; x[34] = 0.f; x[35] = 0.f;
        {
            float xi[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) xi[c] = x[c] * wgt;
            tm_store_a<32>(cx.lane_addr, 64, xi);
        }
        TC_GEMM_BEGIN(cx) tc_issue<L_VF0>(cx, 0, 64, false); TC_GEMM_END(cx)
        tc_epilogue<1>(cx.lane_addr, 0, 2, sw + TS(VF_B0), 96);
        TC_GEMM_BEGIN(cx) tc_issue<L_VF2>(cx, 0, 96, false); TC_GEMM_END(cx)
        {
            float xv[48];
            tm_ld<48>(cx.lane_addr + TM_D, xv); bias_elu<36>(sw + TS(VF_B2), xv);
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] += xv[c];
            const float visw = gn_sigmoid(xv[32]) * mask;       // ibrnet.py:478-479
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
            tm_ld<32>(cx.lane_addr + TM_D, t); bias_elu<32>(sw + TS(V2_B0), t);
            float s = sw[TS(V2_B2)];
#pragma unroll
            for (int k = 0; k < 32; ++k) s = fmaf(t[k], sw[TS(V2_W2) + k], s);
            vis2 = gn_sigmoid(s) * mask;
        }
        // ================= final pooling (ibrnet.py:482-484,487) ===============================================
        float ssum = 0.f;
        for (int jv = 0; jv < V; ++jv) ssum += __shfl_sync(FULL, vis2, (gb + jv) & 31);
        const float w2 = __fdiv_rn(vis2, ssum + 1e-8f);
        float w2sum = 0.f;
        for (int jv = 0; jv < V; ++jv) w2sum += __shfl_sync(FULL, w2, (gb + jv) & 31);
        const float wmean = w2sum / (float)V;
        const bool writer = valid && v == 0;
        float mu[36], vr[36];
        {
            float tmp[36];
#pragma unroll
            for (int c = 0; c < 36; ++c) tmp[c] = w2 * x[c];
            pool36(scr, lane, g, v, gb, V, lane_active, tmp, mu);
#pragma unroll
            for (int c = 0; c < 36; ++c) { const float dl = x[c] - mu[c]; tmp[c] = w2 * dl * dl; }
            pool36(scr, lane, g, v, gb, V, lane_active, tmp, vr);
        }
        if (p.pooled && writer) {
            float* out = p.pooled + (size_t)pidx * GN_POOL_STRIDE;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                st4(out + c, make_float4(mu[c], mu[c + 1], mu[c + 2], mu[c + 3]));
                st4(out + 32 + c, make_float4(vr[c], vr[c + 1], vr[c + 2], vr[c + 3]));
            }
            st4(out + 64, make_float4(wmean, nvalid, 0.f, 0.f));
        }
        if (p.dbg_rows && valid) {
            float* dr = p.dbg_rows + ((size_t)pidx * V + v) * 8;
            st4(dr, make_float4(hit, vis, w0, vis2));
            st4(dr + 4, make_float4(x[0], x[1], 0.f, 0.f));      // prob_embed is fused away in this kernel (not materialised)
        }
        // ================= geometry_fc on the pooled rows (ibrnet.py:487-489) -> per-point token =================
        if (p.tok) {            // uniform branch
            tm_store_a<32>(cx.lane_addr, 0, mu);
            tm_store_a<32>(cx.lane_addr, 32, vr);
            {
                float e[32];
                float px, py, pz;
                if (p.volume_mode) {   // same arithmetic as K1 (field_utils.py:17-27 + bbox3d[0]); n = (i*R+j)*R + (R-1-k)
                    const int R = p.R;
                    const int r = n / R, dsm = n - r * R;
                    const int i = r / R, j = r - i * R, k = R - 1 - dsm;
                    px = __fadd_rn(__ldg(p.axis + i), __ldg(p.bbox_min + b * 3 + 0));
                    py = __fadd_rn(__ldg(p.axis + j), __ldg(p.bbox_min + b * 3 + 1));
                    pz = __fadd_rn(__ldg(p.axis + k), __ldg(p.bbox_min + b * 3 + 2));
                } else {
                    const float* q = p.pts + (size_t)pidx * 3;
                    px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
                }
                const float pv[3] = { px, py, pz };
                e[0] = wmean; e[1] = px; e[2] = py; e[3] = pz;                 // k 64 | embed (neus.py:21-66): p, sin/cos(p*{1,2,4})
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int a = 0; a < 3; ++a) __sincosf(pv[a] * (float)(1 << q), &e[4 + 6 * q + a], &e[4 + 6 * q + 3 + a]);   // |arg| < 4: abs err ~1e-6
#pragma unroll
                for (int c = 22; c < 32; ++c) e[c] = 0.f;
                tm_store_a<32>(cx.lane_addr, 64, e);
            }
            TC_GEMM_BEGIN(cx) tc_issue<L_GF0>(cx, 0, 0, false); TC_GEMM_END(cx)
            tc_epilogue<1>(cx.lane_addr, 0, 4, sw + TS(GF_B0), 0);
            TC_GEMM_BEGIN(cx) tc_issue<L_GF2>(cx, 0, 0, false); TC_GEMM_END(cx)
            {
                float t[16];
                tm_ld<16>(cx.lane_addr + TM_D, t); bias_elu<16>(sw + TS(GF_B2), t);
                if (writer) {
                    float* out = p.tok + (size_t)pidx * GN_TOK_STRIDE;
                    st4(out, make_float4(t[0], t[1], t[2], t[3]));       st4(out + 4, make_float4(t[4], t[5], t[6], t[7]));
                    st4(out + 8, make_float4(t[8], t[9], t[10], t[11])); st4(out + 12, make_float4(t[12], t[13], t[14], t[15]));
                    st4(out + 16, make_float4(nvalid, 0.f, 0.f, 0.f));
                }
            }
        }
        // ================= rgb_fc + masked softmax over views (ibrnet.py:507-511), CUDA cores ================
        if (p.with_rgb && p.colors) {
            float r16[16], r8[8];
#pragma unroll
            for (int c = 0; c < 16; ++c) r16[c] = sw[TS(RF_B0) + c];
            const float dd4[5] = { vis2, ddv.x, ddv.y, ddv.z, ddv.w };
#pragma unroll
            for (int k = 0; k < 37; ++k) {
                const float xk = k < 32 ? x[k] : dd4[k - 32];
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(sw + TS(RF_W0) + k * 16 + c);
                    r16[c] = fmaf(xk, w.x, r16[c]); r16[c + 1] = fmaf(xk, w.y, r16[c + 1]); r16[c + 2] = fmaf(xk, w.z, r16[c + 2]); r16[c + 3] = fmaf(xk, w.w, r16[c + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) r8[c] = sw[TS(RF_B2) + c];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float xk = tc_elu(r16[k]);
#pragma unroll
                for (int c = 0; c < 8; ++c) r8[c] = fmaf(xk, sw[TS(RF_W2) + k * 8 + c], r8[c]);
            }
            float logit = sw[TS(RF_B4)];
#pragma unroll
            for (int k = 0; k < 8; ++k) logit = fmaf(tc_elu(r8[k]), sw[TS(RF_W4) + k], logit);
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
    if (p.V < 2 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (p.que_dists && (p.dn < 1 || (p.N % p.dn) != 0)) return -4;
    if (!p.tc_const) return -7;
    if (p.tok) {
        if (p.volume_mode && (p.R < 1 || p.N != p.R * p.R * p.R || !p.axis || !p.bbox_min)) return -3;
        if (!p.volume_mode && !p.pts) return -4;
    }
    const int G = 32 / p.V;
    const long long total = (long long)p.B * p.N;
    const long long per_tile = 4LL * G;
    const long long tiles = (total + per_tile - 1) / per_tile;
    if (tiles > 0x7fffffffLL) return -6;
    const size_t smem = tc_smem_bytes(G);
    if (smem > 227 * 1024) return -5;
    static size_t smem_cache_gn_k2a_tc_kernel[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k2a_tc_kernel, smem, smem_cache_gn_k2a_tc_kernel);
    if (e != cudaSuccess) return (int)e;
    const int sms = gn_sm_count();
    const char* es = getenv("GN_K2A_SLOTS");
    const int nslots = (es && atoi(es) == 1) ? 1 : TC_SLOTS;
    const long long want = (tiles + nslots - 1) / nslots;
    const int grid = (int)(want < sms ? want : sms);
    gn_k2a_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p, (int)tiles, G, nslots);
    return (int)cudaGetLastError();
}
