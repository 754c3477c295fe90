// K2a, tensor-core version with EIGHT warps per 128-row tile ("tc2").
//
// Same math and the same tcgen05 / TMEM machinery as k2a_head_tc.cu (see its header and k2a_tc_common.cuh), re-balanced
// for latency hiding.  ncu on the 4-warp kernel (profiles/k2a_tc_r01*_source_summary.txt) showed 39 % issue-active with
// two warps per scheduler: every stall of one warp (TMEM load, mbarrier, MUFU chain) idles half of the scheduler.
// Here a tile is still 128 rows = 128 TMEM lanes, but every row is served by TWO threads in two different warps (same
// lane quadrant, hf = 0 / 1): each takes half of the 16-column chunks of every layer's output, so a thread carries half
// the activations (<= 128 registers) and an SM holds 16 warps (2 tiles x 8).  Row-scalar quantities (hit_prob,
// visibility, the three re-weighting factors) are needed by both halves; the 1- and 2-column layers that produce them
// run as tiny N=16 GEMMs so that both halves simply read the accumulator.
#include "k2a_tc_common.cuh"

#define T2_THREADS 512
#define T2_SLOTS 2
#define T2_WPS 8                                   // warps per slot
#define T2_POOL_STRIDE 20                          // floats per scratch row (16 values + pad: conflict-free float4 rows)
__host__ __device__ constexpr size_t t2_smem_bytes(int G) {
    return (size_t)TC_CONST_BYTES + (size_t)(T2_THREADS / 32) * (32 + G) * T2_POOL_STRIDE * 4 + 64;
}
static_assert(t2_smem_bytes(5) <= 227 * 1024, "K2a-TC2 shared memory budget at V = 6");
static_assert(TS(DD_VAR_B0) == TS(DD_MEAN_B0) + 32 && TS(DD_AW_B0) == TS(DD_MEAN_B0) + 64, "dist-decoder .0 biases must be contiguous");
static_assert(TS(DD_VAR_B2) == TS(DD_MEAN_B2) + 32 && TS(DD_AW_B2) == TS(DD_MEAN_B2) + 64, "dist-decoder .2 biases must be contiguous");

struct T2Ctx {
    uint32_t tmem_slot, lane_addr, img_base16, bar, parity, elected;
    int bar_id;
    bool issuer;
};

template <int LAYER>
__device__ __forceinline__ void t2_issue(const T2Ctx& cx, int d_col, int a_k0, bool accumulate) {
    constexpr int N = tc_layer(LAYER).N, K = tc_layer(LAYER).K;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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
#define T2_GEMM_BEGIN(cx)                                                            \
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");                     \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                 \
    asm volatile("bar.sync %0, 256;" :: "r"((cx).bar_id) : "memory");                \
    if ((cx).issuer) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#define T2_GEMM_COMMIT(cx)                                                           \
        tc_commit((cx).bar, (cx).elected); __syncwarp(); }
#define T2_GEMM_WAIT(cx)                                                             \
    mbar_wait((cx).bar, (cx).parity); (cx).parity ^= 1u;                             \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#define T2_GEMM_END(cx) T2_GEMM_COMMIT(cx) T2_GEMM_WAIT(cx)

// A[k0 + 0 .. 16*nchunk) = act(D[d_col + 0 .. 16*nchunk) + bias)   (ACT 0 none, 1 ELU, 2 ReLU); one copy for all layers
template <int ACT>
__device__ __noinline__ void t2_epilogue(uint32_t lane_addr, int d_col, int nchunk, const float* __restrict__ bias, int k0)
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
            y[i] = __uint_as_float(r[i]) + w.x; y[i + 1] = __uint_as_float(r[i + 1]) + w.y;
            y[i + 2] = __uint_as_float(r[i + 2]) + w.z; y[i + 3] = __uint_as_float(r[i + 3]) + w.w;
        }
        if (c + 1 < nchunk) tm_ld16_issue(lane_addr + TM_D + d_col + (c + 1) * 16, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = ACT == 1 ? tc_elu(y[i]) : (ACT == 2 ? fmaxf(y[i], 0.f) : y[i]);
        tm_store_a<16>(lane_addr, k0 + c * 16, y);
    }
}

// cross-view sums of the 16 values every lane parked in its scratch row (order v = 0..V-1)
__device__ __noinline__ void t2_pool_rows(float* scr, int g, int v, int gb, int V, bool lane_active)
{
    __syncwarp();
    float* sums = scr + (32 + g) * T2_POOL_STRIDE;
    const float* base = scr + gb * T2_POOL_STRIDE;
#pragma unroll 1
    for (int ch = v; ch < 4; ch += V) {
        const float* q = base + 4 * ch;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int jv = 0; jv < V; ++jv) {
            const float4 t = *reinterpret_cast<const float4*>(q + jv * T2_POOL_STRIDE);
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        if (lane_active) st4(sums + 4 * ch, s);
    }
    __syncwarp();
}
__device__ __forceinline__ void pool16(float* scr, int lane, int g, int v, int gb, int V, bool lane_active, const float* vals, float* out)
{
    float* mine = scr + lane * T2_POOL_STRIDE;
#pragma unroll
    for (int c = 0; c < 16; c += 4) st4(mine + c, make_float4(vals[c], vals[c + 1], vals[c + 2], vals[c + 3]));
    t2_pool_rows(scr, g, v, gb, V, lane_active);
    const float* sums = scr + (32 + g) * T2_POOL_STRIDE;
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(sums + c);
        out[c] = t.x; out[c + 1] = t.y; out[c + 2] = t.z; out[c + 3] = t.w;
    }
    __syncwarp();
}
__device__ __forceinline__ float gsum(float x, int gb, int V) {
    float s = 0.f;
    for (int jv = 0; jv < V; ++jv) s += __shfl_sync(0xffffffffu, x, (gb + jv) & 31);
    return s;
}

__global__ void __launch_bounds__(T2_THREADS, 1)
gn_k2a_tc2_kernel(const __grid_constant__ GnK2aParams p, int num_tiles, int G)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __half* s_img = reinterpret_cast<__half*>(smem_raw);
    const float* sw = reinterpret_cast<const float*>(smem_raw + (size_t)TC_IMG_HALVES * 2);
    float* s_pool = reinterpret_cast<float*>(smem_raw + TC_CONST_BYTES);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pool + (T2_THREADS / 32) * (32 + G) * T2_POOL_STRIDE);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + T2_SLOTS);
    uint64_t* s_cbar = reinterpret_cast<uint64_t*>(s_tmem + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction
    const int slot = warp >> 3, ws = warp & 7, q = ws & 3, hf = ws >> 2;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(s_cbar)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(s_cbar)), "r"((uint32_t)TC_CONST_BYTES) : "memory");
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.tc_const);
        for (uint32_t off = 0; off < (uint32_t)TC_CONST_BYTES; off += 32768u) {     // TMA bulk copies global -> shared
            const uint32_t n = min(32768u, (uint32_t)TC_CONST_BYTES - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(smem_raw + off)), "l"(src + off), "r"(n), "r"(smem_u32(s_cbar)) : "memory");
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(smem_u32(s_cbar), 0u);

    T2Ctx cx;
    cx.tmem_slot = *s_tmem + slot * TM_SLOT;
    cx.lane_addr = cx.tmem_slot + ((uint32_t)(q * 32) << 16);
    cx.img_base16 = smem_u32(s_img) >> 4;
    cx.bar = smem_u32(&s_bar[slot]);
    cx.parity = 0u;
    cx.bar_id = 1 + slot;
    cx.issuer = ws == 0;
    cx.elected = cx.issuer ? elect_one() : 0u;

    const int V = p.V;
    const bool lane_active = lane < G * V;
    const int g = lane_active ? lane / V : 0;
    const int v = lane_active ? lane - g * V : 0;
    const int gb = g * V;
    const long long total_pts = (long long)p.B * p.N;
    float* scr = s_pool + warp * (32 + G) * T2_POOL_STRIDE;

    for (int tile = blockIdx.x * T2_SLOTS + slot; tile < num_tiles; tile += gridDim.x * T2_SLOTS) {
        long long pidx = ((long long)tile * 4 + q) * G + g;
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
        const bool writer = valid && v == 0;

        // ===== S1: dist-decoder first layers (N = 96); each half stores 16 of the 32 ray_feats columns
        {
            float r[16];
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 t = ldg4(row + GN_REC_RAYF + 16 * hf + c);
                r[c] = t.x; r[c + 1] = t.y; r[c + 2] = t.z; r[c + 3] = t.w;
            }
            tm_store_a<16>(cx.lane_addr, 16 * hf, r);
        }
        T2_GEMM_BEGIN(cx) t2_issue<L_DD1>(cx, 0, 0, false); T2_GEMM_END(cx)
        // ===== S2: second layers
        t2_epilogue<1>(cx.lane_addr, 48 * hf, 3, sw + TS(DD_MEAN_B0) + 48 * hf, 48 + 48 * hf);
        T2_GEMM_BEGIN(cx)
            t2_issue<L_DD2M>(cx, 0, 48, false); t2_issue<L_DD2V>(cx, 32, 80, false); t2_issue<L_DD2A>(cx, 64, 112, false);
        T2_GEMM_END(cx)
        // ===== S2b: third layers (32 -> 2, 2, 1) as one N=16 block GEMM so both halves can read the five outputs
        t2_epilogue<1>(cx.lane_addr, 48 * hf, 3, sw + TS(DD_MEAN_B2) + 48 * hf, 48 + 48 * hf);
        asm volatile("prefetch.global.L1 [%0];" :: "l"(row + GN_REC_IMGF));
        T2_GEMM_BEGIN(cx) t2_issue<L_DD3>(cx, 0, 48, false); T2_GEMM_END(cx)
        float hit, vis;
        {
            float o[16];
            tm_ld<16>(cx.lane_addr + TM_D, o);
            const float mean0 = gn_softplus(o[0] + sw[TS(DD_MEAN_B4)]), mean1 = gn_softplus(o[1] + sw[TS(DD_MEAN_B4) + 1]);
            const float var0 = gn_softplus(o[2] + sw[TS(DD_VAR_B4)]) + 0.05f, var1 = gn_softplus(o[3] + sw[TS(DD_VAR_B4) + 1]) + 0.05f;
            const float aw = gn_sigmoid(o[4] + sw[TS(DD_AW_B4)]);
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
            const float c00 = gn_sigmoid(2.f * ((nearp - mean0) * var0)), c10 = gn_sigmoid(2.f * ((farp - mean0) * var0));
            const float c01 = gn_sigmoid(2.f * ((nearp - mean1) * var1)), c11 = gn_sigmoid(2.f * ((farp - mean1) * var1));
            const float mix1 = 1.f - aw;
            vis = ((1.f - c00) * aw + (1.f - c01) * mix1) * mask;
            hit = ((c10 - c00) * aw + (c11 - c01) * mix1) * mask;
        }
        // ===== S3: prob_embed.0 on [ray | 2hit-1 | 2vis-1]
        if (hf == 0) {
            float hv[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) hv[c] = 0.f;
            hv[0] = (hit - 0.5f) * 2.f; hv[1] = (vis - 0.5f) * 2.f;
            tm_store_a<16>(cx.lane_addr, 32, hv);
        }
        T2_GEMM_BEGIN(cx) t2_issue<L_PE0>(cx, 0, 0, false); T2_GEMM_END(cx)
        // ===== S4: prob_embed.2
        t2_epilogue<2>(cx.lane_addr, 16 * hf, 1, sw + TS(PE_B0) + 16 * hf, 80 + 16 * hf);
        T2_GEMM_BEGIN(cx) t2_issue<L_PE2>(cx, 0, 80, false); T2_GEMM_END(cx)
        // ===== S5: neuray_fc.0 on prob_emb (k 48..79), ray_dir_fc.0 on dir_diff (k 112..127)
        float pe01[2] = { 0.f, 0.f };
        if (p.dbg_rows && hf == 0) {
            float t16[16];
            tm_ld<16>(cx.lane_addr + TM_D, t16);
            pe01[0] = t16[0] + sw[TS(PE_B2)]; pe01[1] = t16[1] + sw[TS(PE_B2) + 1];
        }
        t2_epilogue<0>(cx.lane_addr, 16 * hf, 1, sw + TS(PE_B2) + 16 * hf, 48 + 16 * hf);
        if (hf == 0) {
            float dd[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) dd[c] = 0.f;
            dd[0] = ddv.x; dd[1] = ddv.y; dd[2] = ddv.z; dd[3] = ddv.w;
            tm_store_a<16>(cx.lane_addr, 112, dd);
        }
        T2_GEMM_BEGIN(cx) t2_issue<L_NF0>(cx, 0, 48, false); t2_issue<L_RD0>(cx, 16, 112, false); T2_GEMM_END(cx)
        // ===== S6: weight0 (both halves), ray_dir_fc hidden (half 1)
        float w0;
        {
            float t[16];
            tm_ld<16>(cx.lane_addr + TM_D, t);
            float s = sw[TS(NF_B2)];
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(tc_elu(t[k] + sw[TS(NF_B0) + k]), sw[TS(NF_W2) + k], s);
            w0 = gn_sigmoid(s) * wgt;                           // ibrnet.py:469
        }
        if (hf == 1) t2_epilogue<1>(cx.lane_addr, 16, 1, sw + TS(RD_B0), 128);
        T2_GEMM_BEGIN(cx) t2_issue<L_RD1>(cx, 0, 128, false); T2_GEMM_END(cx)
        // ===== S7: f = feats + direction feature (half hf owns image-feature channels 16hf..16hf+15, half 0 also the 3 rgb
        //       channels); S7a runs under the mean poolings; then [mean0 | mean1 | var0 | var1 | tails] for S7b
        {
            float f[16], fr0 = 0.f, fr1 = 0.f, fr2 = 0.f;
            tm_ld<16>(cx.lane_addr + TM_D + 16 * hf, f); bias_elu<16>(sw + TS(RD_B1) + 16 * hf, f);
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 t = ldg4(row + GN_REC_IMGF + 16 * hf + c);
                f[c] += t.x; f[c + 1] += t.y; f[c + 2] += t.z; f[c + 3] += t.w;     // ibrnet.py:459
            }
            tm_store_a<16>(cx.lane_addr, 16 * hf, f);
            if (hf == 0) {
                float t[16];
                tm_ld<16>(cx.lane_addr + TM_D + 32, t);
                fr0 = tc_elu(t[0] + sw[TS(RD_B1) + 32]) + tail.x;
                fr1 = tc_elu(t[1] + sw[TS(RD_B1) + 33]) + tail.y;
                fr2 = tc_elu(t[2] + sw[TS(RD_B1) + 34]) + tail.z;
#pragma unroll
                for (int c = 3; c < 16; ++c) t[c] = 0.f;
                t[0] = fr0; t[1] = fr1; t[2] = fr2;
                tm_store_a<16>(cx.lane_addr, 32, t);
            }
            T2_GEMM_BEGIN(cx) t2_issue<L_BF0A>(cx, 0, 0, false); T2_GEMM_COMMIT(cx)
            float m0[16], m1[16], tmp[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) tmp[c] = w0 * f[c];
            pool16(scr, lane, g, v, gb, V, lane_active, tmp, m0);
#pragma unroll
            for (int c = 0; c < 16; ++c) tmp[c] = wgt * f[c];
            pool16(scr, lane, g, v, gb, V, lane_active, tmp, m1);
            float tl[16];
            if (hf == 0) {                                      // rgb channels: 3 values, warp shuffles
                tl[0] = gsum(w0 * fr0, gb, V); tl[1] = gsum(w0 * fr1, gb, V); tl[2] = gsum(w0 * fr2, gb, V);
                tl[3] = gsum(wgt * fr0, gb, V); tl[4] = gsum(wgt * fr1, gb, V); tl[5] = gsum(wgt * fr2, gb, V);
                float d;
                d = fr0 - tl[0]; tl[6] = gsum(w0 * d * d, gb, V); d = fr1 - tl[1]; tl[7] = gsum(w0 * d * d, gb, V); d = fr2 - tl[2]; tl[8] = gsum(w0 * d * d, gb, V);
                d = fr0 - tl[3]; tl[9] = gsum(wgt * d * d, gb, V); d = fr1 - tl[4]; tl[10] = gsum(wgt * d * d, gb, V); d = fr2 - tl[5]; tl[11] = gsum(wgt * d * d, gb, V);
                tl[12] = 0.f; tl[13] = 0.f; tl[14] = 0.f; tl[15] = 0.f;
            }
            T2_GEMM_WAIT(cx)
            tm_store_a<16>(cx.lane_addr, 16 * hf, m0);
            tm_store_a<16>(cx.lane_addr, 32 + 16 * hf, m1);
#pragma unroll
            for (int c = 0; c < 16; ++c) { const float d0 = f[c] - m0[c]; tmp[c] = w0 * d0 * d0; }      // ibrnet.py:115
            pool16(scr, lane, g, v, gb, V, lane_active, tmp, m0);
            tm_store_a<16>(cx.lane_addr, 64 + 16 * hf, m0);
#pragma unroll
            for (int c = 0; c < 16; ++c) { const float d1 = f[c] - m1[c]; tmp[c] = wgt * d1 * d1; }
            pool16(scr, lane, g, v, gb, V, lane_active, tmp, m1);
            tm_store_a<16>(cx.lane_addr, 96 + 16 * hf, m1);
            if (hf == 0) tm_store_a<16>(cx.lane_addr, 128, tl);
        }
        T2_GEMM_BEGIN(cx) t2_issue<L_BF0B>(cx, 0, 0, true); T2_GEMM_END(cx)
        {   // next tile's first record lines: hide their latency under S8..S11
            const long long npidx = ((long long)(tile + gridDim.x * T2_SLOTS) * 4 + q) * G + g;
            if (npidx < total_pts) {
                asm volatile("prefetch.global.L1 [%0];" :: "l"(p.rec + ((size_t)npidx * V + v) * GN_REC_STRIDE));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(p.rec + ((size_t)npidx * V + v) * GN_REC_STRIDE + GN_REC_RGB));
            }
        }
        // ===== S8: base_fc.2
        t2_epilogue<1>(cx.lane_addr, 32 * hf, 2, sw + TS(BF_B0) + 32 * hf, 32 * hf);
        T2_GEMM_BEGIN(cx) t2_issue<L_BF2>(cx, 0, 0, false); T2_GEMM_END(cx)
        // ===== S9 / S10: vis_fc  (each half keeps its 16 columns of x to the end of the tile)
        float x[16];
        tm_ld<16>(cx.lane_addr + TM_D + 16 * hf, x); bias_elu<16>(sw + TS(BF_B2) + 16 * hf, x);
        {
            float xi[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) xi[c] = x[c] * wgt;
            tm_store_a<16>(cx.lane_addr, 64 + 16 * hf, xi);
        }
        T2_GEMM_BEGIN(cx) t2_issue<L_VF0>(cx, 0, 64, false); T2_GEMM_END(cx)
        t2_epilogue<1>(cx.lane_addr, 16 * hf, 1, sw + TS(VF_B0) + 16 * hf, 96 + 16 * hf);
        T2_GEMM_BEGIN(cx) t2_issue<L_VF2>(cx, 0, 96, false); T2_GEMM_END(cx)
        {
            float xv[16], t[16];
            tm_ld<16>(cx.lane_addr + TM_D + 16 * hf, xv); bias_elu<16>(sw + TS(VF_B2) + 16 * hf, xv);
            tm_ld<16>(cx.lane_addr + TM_D + 32, t);
            const float visw = gn_sigmoid(tc_elu(t[0] + sw[TS(VF_B2) + 32])) * mask;     // ibrnet.py:478-479
#pragma unroll
            for (int c = 0; c < 16; ++c) { x[c] += xv[c]; xv[c] = x[c] * visw; }
            tm_store_a<16>(cx.lane_addr, 64 + 16 * hf, xv);
        }
        // ===== S11: vis_fc2 (32 -> 32 -> 1; the 1-column layer as an N=16 GEMM)
        T2_GEMM_BEGIN(cx) t2_issue<L_V20>(cx, 0, 64, false); T2_GEMM_END(cx)
        t2_epilogue<1>(cx.lane_addr, 16 * hf, 1, sw + TS(V2_B0) + 16 * hf, 96 + 16 * hf);
        T2_GEMM_BEGIN(cx) t2_issue<L_V22>(cx, 0, 96, false); T2_GEMM_END(cx)
        float vis2;
        {
            float t[16];
            tm_ld<16>(cx.lane_addr + TM_D, t);
            vis2 = gn_sigmoid(t[0] + sw[TS(V2_B2)]) * mask;
        }
        // ===== final pooling (ibrnet.py:482-484,487)
        const float ssum = gsum(vis2, gb, V);
        const float w2 = __fdiv_rn(vis2, ssum + 1e-8f);
        const float wmean = gsum(w2, gb, V) / (float)V;
        float mu[16], vr[16];
        {
            float tmp[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) tmp[c] = w2 * x[c];
            pool16(scr, lane, g, v, gb, V, lane_active, tmp, mu);
#pragma unroll
            for (int c = 0; c < 16; ++c) { const float dl = x[c] - mu[c]; tmp[c] = w2 * dl * dl; }
            pool16(scr, lane, g, v, gb, V, lane_active, tmp, vr);
        }
        if (p.pooled && writer) {
            float* out = p.pooled + (size_t)pidx * GN_POOL_STRIDE;
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                st4(out + 16 * hf + c, make_float4(mu[c], mu[c + 1], mu[c + 2], mu[c + 3]));
                st4(out + 32 + 16 * hf + c, make_float4(vr[c], vr[c + 1], vr[c + 2], vr[c + 3]));
            }
            if (hf == 0) st4(out + 64, make_float4(wmean, nvalid, 0.f, 0.f));
        }
        if (p.dbg_rows && valid && hf == 0) {
            float* dr = p.dbg_rows + ((size_t)pidx * V + v) * 8;
            st4(dr, make_float4(hit, vis, w0, vis2));
            st4(dr + 4, make_float4(x[0], x[1], pe01[0], pe01[1]));
        }
        // ===== geometry_fc on the pooled rows (ibrnet.py:487-489) -> per-point token
        if (p.tok) {
            tm_store_a<16>(cx.lane_addr, 16 * hf, mu);
            tm_store_a<16>(cx.lane_addr, 32 + 16 * hf, vr);
            {
                float px, py, pz;
                if (p.volume_mode) {
                    const int R = p.R;
                    const int r = n / R, dsm = n - r * R;
                    const int i = r / R, j = r - i * R, k = R - 1 - dsm;
                    px = __fadd_rn(__ldg(p.axis + i), __ldg(p.bbox_min + b * 3 + 0));
                    py = __fadd_rn(__ldg(p.axis + j), __ldg(p.bbox_min + b * 3 + 1));
                    pz = __fadd_rn(__ldg(p.axis + k), __ldg(p.bbox_min + b * 3 + 2));
                } else {
                    const float* qp = p.pts + (size_t)pidx * 3;
                    px = __ldg(qp); py = __ldg(qp + 1); pz = __ldg(qp + 2);
                }
                const float pv[3] = { px, py, pz };
                float e[32];                                        // k 64.. : wmean | p | sin/cos(p*{1,2,4})   (neus.py:21-66)
                e[0] = wmean; e[1] = px; e[2] = py; e[3] = pz;
#pragma unroll
                for (int fq = 0; fq < 3; ++fq)
#pragma unroll
                    for (int a = 0; a < 3; ++a) __sincosf(pv[a] * (float)(1 << fq), &e[4 + 6 * fq + a], &e[4 + 6 * fq + 3 + a]);
#pragma unroll
                for (int c = 22; c < 32; ++c) e[c] = 0.f;
                if (hf == 0) tm_store_a<16>(cx.lane_addr, 64, e); else tm_store_a<16>(cx.lane_addr, 80, e + 16);
            }
            T2_GEMM_BEGIN(cx) t2_issue<L_GF0>(cx, 0, 0, false); T2_GEMM_END(cx)
            t2_epilogue<1>(cx.lane_addr, 32 * hf, 2, sw + TS(GF_B0) + 32 * hf, 32 * hf);
            T2_GEMM_BEGIN(cx) t2_issue<L_GF2>(cx, 0, 0, false); T2_GEMM_END(cx)
            if (hf == 0) {
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
        // ===== rgb_fc (ibrnet.py:507-511): first layer (37 -> 16) as a GEMM on [x | vis, dir_diff], the rest on half 0
        if (p.with_rgb && p.colors) {
            tm_store_a<16>(cx.lane_addr, 16 * hf, x);
            if (hf == 1) {
                float t[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) t[c] = 0.f;
                t[0] = vis2; t[1] = ddv.x; t[2] = ddv.y; t[3] = ddv.z; t[4] = ddv.w;
                tm_store_a<16>(cx.lane_addr, 32, t);
            }
            T2_GEMM_BEGIN(cx) t2_issue<L_RF0>(cx, 0, 0, false); T2_GEMM_END(cx)
            if (hf == 0) {
                float r16[16], r8[8];
                tm_ld<16>(cx.lane_addr + TM_D, r16); bias_elu<16>(sw + TS(RF_B0), r16);
#pragma unroll
                for (int c = 0; c < 8; ++c) r8[c] = sw[TS(RF_B2) + c];
#pragma unroll
                for (int k = 0; k < 16; ++k)
#pragma unroll
                    for (int c = 0; c < 8; ++c) r8[c] = fmaf(r16[k], sw[TS(RF_W2) + k * 8 + c], r8[c]);
                float logit = sw[TS(RF_B4)];
#pragma unroll
                for (int k = 0; k < 8; ++k) logit = fmaf(tc_elu(r8[k]), sw[TS(RF_W4) + k], logit);
                if (mask == 0.f) logit = -1e9f;
                float mx = -INFINITY;
                for (int jv = 0; jv < V; ++jv) mx = fmaxf(mx, __shfl_sync(0xffffffffu, logit, (gb + jv) & 31));
                const float e = __expf(logit - mx);
                const float bw = __fdiv_rn(e, gsum(e, gb, V));
                const float c0 = gsum(bw * tail.x, gb, V), c1 = gsum(bw * tail.y, gb, V), c2 = gsum(bw * tail.z, gb, V);
                if (writer) st4(p.colors + (size_t)pidx * 4, make_float4(c0, c1, c2, 0.f));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*s_tmem), "r"(512));
}

extern "C" int gn_k2a_forward_tc2(const GnK2aParams* hp, void* stream)
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
    const size_t smem = t2_smem_bytes(G);
    if (smem > 227 * 1024) return -5;
    static size_t smem_cache[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k2a_tc2_kernel, smem, smem_cache);
    if (e != cudaSuccess) return (int)e;
    const int sms = gn_sm_count();
    const long long want = (tiles + T2_SLOTS - 1) / T2_SLOTS;
    const int grid = (int)(want < sms ? want : sms);
    gn_k2a_tc2_kernel<<<grid, T2_THREADS, smem, (cudaStream_t)stream>>>(p, (int)tiles, G);
    return (int)cudaGetLastError();
}
