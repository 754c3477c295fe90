// K2a on tcgen05 / TMEM, three resident 128-row tiles per SM: the GEMM chain is scheduled so that a tile needs 160
// tensor-memory columns, which lets THREE tiles (12 warps, 3 per scheduler) share an SM.  (The first version, two tiles of
// 256 columns, was 15 % slower and is gone; profiles/ keeps its measurements.)
//
// Why: the kernel is bound by exposed latency per resident warp, not by the tensor pipe or the instruction count
// (profiles/: one tile per SM 455 us, two tiles 275 us per 40^3 volume; T(n) ~ 95 + 360/n us).  Registers allow 12 warps
// (168 per thread); tensor memory (512 columns) and shared memory (153 KB of weight images) are what limit the tile count.
//
// Operand images: gn_k2a_tc_prepare (k2a_tc_prepare.cu).  Layout of a tile:
//   * TMEM map per tile: D 80 columns | A hi 40 | A lo 40  (K <= 80 per GEMM issue);
//   * the GEMM chain is cut into smaller rounds that fit that map - sub-ranges of the prepared images are addressed through
//     the descriptor (row offset n0, k offset k0), so no new images are needed:
//       R1  [mean,var first layers N=64 | ray_dir_fc.0]          R2  [mean,var second layers]
//       R3  [aw first layer | ray_dir_fc.2]  (ray_feats re-stored) R4  [aw second layer]
//       R5  [prob_embed.0]                                        R6  [neuray_fc.0 o prob_embed.2 | base_fc.0 per-view part]
//       R7a [base_fc.0 on mean0|mean1]  R7b [base_fc.0 on var0|var1|tails]   R8 [base_fc.2]
//       R9  [vis_fc.0]  R10 [vis_fc.2]  R11 [vis_fc2.0]
//   * geometry_fc (per POINT, not per view) runs in a second phase of the same launch (t3_geometry_phase): the CTA re-reads
//     the pooled rows of its own points and runs [geometry_fc.0 on mean|var, then on embed], [geometry_fc.2], one row per point;
//   * tiles are handed out inside the CTA from a shared-memory counter (each CTA owns a contiguous range of tiles), so
//     the three slots stay busy although a slot only sees ~7 tiles of a 40^3 volume.
#include "k2a_tc_common.cuh"
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

#define T3_THREADS 384
#define T3_SLOTS 3
// TMEM column map per slot
#define T3_D 0
#define T3_AHI 80
#define T3_ALO 120
#define T3_SLOT 160
#define T3_POOL_STRIDE 36                                                     // floats per row of the pooling scratch (144 B: conflict-free float4 rows)

// per-warp pooling scratch [(32+G) rows][36 floats]; its first 32 rows double as the landing zone of the warp's TMA boxes
// (a record half is 36 floats = one scratch row), so the stride is rounded up to the 128-byte alignment TMA wants
__host__ __device__ constexpr size_t t3_warp_scratch_bytes(int G) { return (((size_t)(32 + G) * T3_POOL_STRIDE * 4) + 127) / 128 * 128; }
__host__ __device__ constexpr size_t t3_smem_bytes(int G) {
    return (size_t)TC_CONST_BYTES + 128 + (size_t)(T3_THREADS / 32) * t3_warp_scratch_bytes(G) + 256;   // + alignment slack, barriers, tmem ptr, tile counter
}
static_assert(t3_smem_bytes(5) <= 227 * 1024, "K2a-TC3 shared memory budget at V = 6");
#define T3_MAX_G 8                                                             // points per warp: min(32 / V, 8) - the scratch of 2- and 3-view scenes must fit too
static_assert(t3_smem_bytes(T3_MAX_G) <= 227 * 1024, "K2a-TC3 shared memory budget at G = 8");

template <int K> __device__ __forceinline__ void t3_store_a(uint32_t lane_addr, int k0, const float* a) { tm_store_a<K, T3_AHI, T3_ALO>(lane_addr, k0, a); }

// Generic layer epilogue: A[k0 ..) = act(D[d_col ..) + bias), 16 columns per step, next load in flight.
template <int ACT>
__device__ __noinline__ void t3_epilogue(uint32_t lane_addr, int d_col, int nchunk, const float* __restrict__ bias, int k0)
{
    uint32_t r[16];
    tm_ld16_issue(lane_addr + T3_D + d_col, r);
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
        if (c + 1 < nchunk) tm_ld16_issue(lane_addr + T3_D + d_col + (c + 1) * 16, r);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (ACT == 1) tc_elu2(y[i], y[i + 1]);
            else if (ACT == 2) { y[i] = y[i] > 0.f ? y[i] : y[i] * 0.f; y[i + 1] = y[i + 1] > 0.f ? y[i + 1] : y[i + 1] * 0.f; }   // ReLU that keeps NaN (fmaxf would hide an overflow)
        }
        t3_store_a<16>(lane_addr, k0 + c * 16, y);
    }
}

// sum of the V rows of a point, float4 chunk `ch` (packed FADD2 accumulation)
template <int VV>
__device__ __forceinline__ void t3_pool_chunk(const float* q, float* dst, int V, bool lane_active)
{
    unsigned long long s0 = 0ull, s1 = 0ull;              // (+0, +0) pairs
    if (VV > 0) {
#pragma unroll
        for (int jv = 0; jv < VV; ++jv) {
            const float4 t = *reinterpret_cast<const float4*>(q + jv * T3_POOL_STRIDE);
            s0 = add2(s0, pk2(t.x, t.y)); s1 = add2(s1, pk2(t.z, t.w));
        }
    } else {
#pragma unroll 2
        for (int jv = 0; jv < V; ++jv) {
            const float4 t = *reinterpret_cast<const float4*>(q + jv * T3_POOL_STRIDE);
            s0 = add2(s0, pk2(t.x, t.y)); s1 = add2(s1, pk2(t.z, t.w));
        }
    }
    float4 s;
    upk2(s0, s.x, s.y); upk2(s1, s.z, s.w);
    if (lane_active) st4(dst, s);
}
__device__ __noinline__ void t3_pool_rows(float* scr, int g, int v, int gb, int V, bool lane_active)
{
    __syncwarp();
    float* sums = scr + (32 + g) * T3_POOL_STRIDE;
    const float* base = scr + gb * T3_POOL_STRIDE;
    // lane (g,v) sums the float4 chunks v, v+V, ... of its point's V rows.  V = 6 (the shipped configuration): chunk v for
    // every lane, chunk v+6 for v < 3, fully unrolled.
    if (V == 6) {
        t3_pool_chunk<6>(base + 4 * v, sums + 4 * v, V, lane_active);
        if (v < 3) t3_pool_chunk<6>(base + 4 * (v + 6), sums + 4 * (v + 6), V, lane_active);
    } else {
#pragma unroll 1
        for (int ch = v; ch < 9; ch += V) t3_pool_chunk<0>(base + 4 * ch, sums + 4 * ch, V, lane_active);
    }
    __syncwarp();
}
__device__ __forceinline__ void t3_pool36(float* scr, int lane, int g, int v, int gb, int V, bool lane_active, const float* vals, float* out)
{
    float* mine = scr + lane * T3_POOL_STRIDE;
#pragma unroll
    for (int c = 0; c < 36; c += 4) st4(mine + c, make_float4(vals[c], vals[c + 1], vals[c + 2], vals[c + 3]));
    t3_pool_rows(scr, g, v, gb, V, lane_active);
    const float* sums = scr + (32 + g) * T3_POOL_STRIDE;
#pragma unroll
    for (int c = 0; c < 36; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(sums + c);
        out[c] = t.x; out[c + 1] = t.y; out[c + 2] = t.z; out[c + 3] = t.w;
    }
    __syncwarp();
}

// One GEMM (three fp16 passes) on a SUB-RANGE of a prepared image: rows [N0, N0+NS) of the layer's N outputs, k range
// [K0, K0+KS) of its K inputs.  Image element (n,k) sits at 16-byte unit (k/8)*N + n  (k-chunk stride LBO = N*16 B).
template <int LAYER, int N0, int NS, int K0, int KS>
__device__ __forceinline__ void t3_issue(const TcCtx& cx, int d_col, int a_k0, bool accumulate) {
    constexpr int N = tc_layer(LAYER).N, K = tc_layer(LAYER).K;
    static_assert(N0 % 8 == 0 && NS % 16 == 0 && N0 + NS <= N && K0 % 16 == 0 && KS % 16 == 0 && K0 + KS <= K && KS <= 80, "sub-GEMM range");
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F32 accum, F16 x F16, M=128
    constexpr uint32_t lbo_field = (uint32_t)((N * 16) >> 4) << 16;
    constexpr uint32_t sub = (uint32_t)((K0 / 8) * N + N0);
    constexpr uint32_t hi_off = ((uint32_t)(tc_img_off(LAYER) * 2) >> 4) + sub, lo_off = hi_off + (uint32_t)((N * K * 2) >> 4);
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {                 // small terms first: lo*hi, hi*lo, hi*hi
        const uint32_t a_col = (pass == 0 ? T3_ALO : T3_AHI) + a_k0 / 2;
        const uint32_t boff = (pass == 1) ? lo_off : hi_off;
#pragma unroll
        for (int ks = 0; ks < KS / 16; ++ks) {
            tc_mma(cx.tmem_slot + T3_D + d_col, cx.tmem_slot + a_col + ks * 8,
                   cx.img_base16 + (boff + (uint32_t)(ks * 2 * N) + lbo_field), idesc, acc, cx.elected);
            acc = 1u;
        }
    }
}
template <int LAYER>
__device__ __forceinline__ void t3_issue_full(const TcCtx& cx, int d_col, int a_k0, bool accumulate) {
    t3_issue<LAYER, 0, tc_layer(LAYER).N, 0, tc_layer(LAYER).K>(cx, d_col, a_k0, accumulate);
}

// Phase 2 of the launch: geometry_fc (ibrnet.py:487-489) once per POINT.  In the tile loop every point occupies V rows, so
// running geometry_fc there repeats it V times (15 % of the kernel's instructions at V = 6).  The pooled features of this
// CTA's points were written to `pooled` by the tile loop (L2 resident: ~120 KB per CTA); here a row is one point, 128 points
// per tile.  NOT inlined: its register allocation must not disturb the tile loop's (which sits right at the 168-register
// limit of 12 warps per SM).
__device__ __noinline__ void t3_geometry_phase(const GnK2aParams& p, uint32_t tmem_slot, uint32_t img_base16, uint32_t bar, uint32_t parity,
                                               const float* __restrict__ sw, int* s_ctr, int* s_tile, int slot, int warp, int lane,
                                               long long pt_lo, long long pt_hi)
{
    TcCtx cx;                                           // rebuilt from scalars: a reference would pin the caller's copy in local memory
    cx.tmem_slot = tmem_slot;
    cx.lane_addr = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    cx.img_base16 = img_base16;
    cx.bar = bar;
    cx.parity = parity;
    cx.bar_id = 1 + slot;
    cx.issuer = (warp & 3) == 0;
    cx.elected = cx.issuer ? elect_one() : 0u;
    const int n2 = (int)((pt_hi - pt_lo + 127) / 128);
    for (;;) {
        if ((warp & 3) == 0 && lane == 0) s_tile[slot] = atomicAdd(s_ctr, 1);
        asm volatile("bar.sync %0, 128;" :: "r"(cx.bar_id) : "memory");
        const int t2 = s_tile[slot];
        asm volatile("bar.sync %0, 128;" :: "r"(cx.bar_id) : "memory");
        if (t2 >= n2) break;
        long long pidx = pt_lo + (long long)t2 * 128 + (warp & 3) * 32 + lane;
        const bool valid2 = pidx < pt_hi;
        pidx = valid2 ? pidx : pt_hi - 1;
        const int b = (int)(pidx / p.N);
        const int n = (int)(pidx - (long long)b * p.N);
        const float* pr = p.pooled + (size_t)pidx * GN_POOL_STRIDE;
        {
            float mu[32];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 a4 = __ldcg(reinterpret_cast<const float4*>(pr + c));
                mu[c] = a4.x; mu[c + 1] = a4.y; mu[c + 2] = a4.z; mu[c + 3] = a4.w;
            }
            t3_store_a<32>(cx.lane_addr, 0, mu);
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 b4 = __ldcg(reinterpret_cast<const float4*>(pr + 32 + c));
                mu[c] = b4.x; mu[c + 1] = b4.y; mu[c + 2] = b4.z; mu[c + 3] = b4.w;
            }
            t3_store_a<32>(cx.lane_addr, 32, mu);
        }
        const float4 wn = __ldcg(reinterpret_cast<const float4*>(pr + 64));       // wmean, nvalid
        TC_GEMM_BEGIN(cx) t3_issue<L_GF0, 0, 64, 0, 64>(cx, 0, 0, false); TC_GEMM_COMMIT(cx)
        float e[32];
        {
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
            e[0] = wn.x; e[1] = px; e[2] = py; e[3] = pz;                  // k 64 | embed (neus.py:21-66): p, sin/cos(p*{1,2,4})
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int a = 0; a < 3; ++a) __sincosf(pv[a] * (float)(1 << q), &e[4 + 6 * q + a], &e[4 + 6 * q + 3 + a]);   // |arg| < 4: abs err ~1e-6
#pragma unroll
            for (int c = 22; c < 32; ++c) e[c] = 0.f;
        }
        TC_GEMM_WAIT(cx)
        t3_store_a<32>(cx.lane_addr, 0, e);
        TC_GEMM_BEGIN(cx) t3_issue<L_GF0, 0, 64, 64, 32>(cx, 0, 0, true); TC_GEMM_END(cx)
        t3_epilogue<1>(cx.lane_addr, 0, 4, sw + TS(GF_B0), 0);
        TC_GEMM_BEGIN(cx) t3_issue_full<L_GF2>(cx, 0, 0, false); TC_GEMM_END(cx)
        {
            float tk[16];
            tm_ld<16>(cx.lane_addr + T3_D, tk); bias_elu<16>(sw + TS(GF_B2), tk);
            if (valid2) {
                float chk = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) chk += fabsf(tk[c]);
                if (p.status && !(chk <= 3.0e38f)) atomicOr(p.status, 1);       // inf / NaN reached the tokens (fp16 operand overflow upstream)
                float* out = p.tok + (size_t)pidx * GN_TOK_STRIDE;
                st4(out, make_float4(tk[0], tk[1], tk[2], tk[3]));       st4(out + 4, make_float4(tk[4], tk[5], tk[6], tk[7]));
                st4(out + 8, make_float4(tk[8], tk[9], tk[10], tk[11])); st4(out + 12, make_float4(tk[12], tk[13], tk[14], tk[15]));
                st4(out + 16, make_float4(wn.y, 0.f, 0.f, 0.f));
            }
        }
    }
}

__global__ void __launch_bounds__(T3_THREADS, 1)
gn_k2a_tc3_kernel(const __grid_constant__ GnK2aParams p, const __grid_constant__ CUtensorMap rec_map, int num_tiles, int G)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __half* s_img = reinterpret_cast<__half*>(smem_raw);
    const float* sw = reinterpret_cast<const float*>(smem_raw + (size_t)TC_IMG_HALVES * 2);   // small fp32 constants, index with TS()
    unsigned char* pool_raw = smem_raw + (((size_t)TC_CONST_BYTES + 127) / 128) * 128;         // 128-byte aligned (smem_raw is 1024-aligned)
    const size_t wscr = t3_warp_scratch_bytes(G);                                               // [12 warps][(32+G)][36], 128-byte multiples
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(pool_raw + (T3_THREADS / 32) * wscr);        // [3] slot barriers, [1] constants, [12] per-warp TMA barriers
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + T3_SLOTS + 1 + T3_THREADS / 32);
    int* s_ctr = reinterpret_cast<int*>(s_tmem + 1);                   // next tile of this CTA's range
    int* s_tile = s_ctr + 1;                                           // [3] tile handed to each slot

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction
    const int slot = warp >> 2;
    // this CTA's contiguous tile range
    const int tile_lo = (int)((long long)blockIdx.x * num_tiles / gridDim.x);
    const int tile_hi = (int)((long long)(blockIdx.x + 1) * num_tiles / gridDim.x);
    uint64_t* s_cbar = s_bar + T3_SLOTS;
    if (tid == 0) {
        *s_ctr = tile_lo;
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
        for (int s = 0; s < T3_SLOTS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[s])), "r"(1));
        for (int w = 0; w < T3_THREADS / 32; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[T3_SLOTS + 1 + w])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(smem_u32(s_cbar), 0u);                                  // operand images + constants have landed (async proxy)

    TcCtx cx;
    cx.tmem_slot = *s_tmem + slot * T3_SLOT;
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
    float* scr = reinterpret_cast<float*>(pool_raw + warp * wscr);

    // ---- record staging by TMA (round 2).  A warp's G*V rows of a tile are contiguous in HBM; each row is two halves of 36
    // floats ([ray_feats | dir_diff], [rgb, depth | img_feats]).  One cp.async.bulk.tensor box {36 floats, G*V rows} lands a
    // half of all the warp's rows in its pooling scratch (row pitch 36 floats = the box's dense layout), where lane l reads
    // row l.  Half A of the NEXT tile is requested at the end of the current one, half B after round 3 has consumed half A:
    // the record's latency no longer sits in front of the first use, and no thread issues 18 strided LDG.128 per tile any more.
    const uint32_t tma_bar = smem_u32(&s_bar[T3_SLOTS + 1 + warp]);
    uint32_t tma_par = 0u;
    const uint32_t scr_s = smem_u32(scr);
    const uint32_t box_bytes = (uint32_t)(G * V) * GN_REC_HALF * 4u;
    const float* mine = scr + lane * T3_POOL_STRIDE;
    auto tma_rows = [&](int col, long long first_row) {           // whole warp calls; lane 0 issues
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the scratch was last touched through the generic proxy
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tma_bar), "r"(box_bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(scr_s), "l"(&rec_map), "r"(col), "r"((int)first_row), "r"(tma_bar) : "memory");
        }
    };
    auto warp_first_row = [&](int t) { return ((long long)t * 4 + (warp & 3)) * G * V; };
    if (tile_lo + slot < tile_hi) tma_rows(0, warp_first_row(tile_lo + slot));

    // tiles of the CTA's range go round-robin to the three slots (all tiles cost the same; a static order lets a slot
    // request its next tile's rows early)
    for (int tile = tile_lo + slot; tile < tile_hi; tile += T3_SLOTS) {
        long long pidx = ((long long)tile * 4 + (warp & 3)) * G + g;
        const bool valid = lane_active && pidx < total_pts;
        pidx = pidx < total_pts ? pidx : total_pts - 1;
        const int b = (int)(pidx / p.N);
        const int n = (int)(pidx - (long long)b * p.N);
        const float2 ptv = __ldg(reinterpret_cast<const float2*>(p.pt + (size_t)pidx * GN_PT_STRIDE));
        mbar_wait(tma_bar, tma_par); tma_par ^= 1u;            // half A of this tile's rows: [ray_feats 32 | dir_diff 4] in scratch row `lane`
        const float4 ddv = *reinterpret_cast<const float4*>(mine + GN_REC_DD);

        // ================= R1: mean / var first layers (N = 64) on ray_feats, ray_dir_fc.0 (N = 16) on dir_diff ==============
        {
            float ray[32];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(mine + GN_REC_RAYF + c);
                ray[c] = t.x; ray[c + 1] = t.y; ray[c + 2] = t.z; ray[c + 3] = t.w;
            }
            t3_store_a<32>(cx.lane_addr, 0, ray);               // A[k 0..31] = ray_feats
            float dd[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) dd[c] = 0.f;
            dd[0] = ddv.x; dd[1] = ddv.y; dd[2] = ddv.z; dd[3] = ddv.w;
            t3_store_a<16>(cx.lane_addr, 32, dd);               // A[k 32..47] = dir_diff
        }
        TC_GEMM_BEGIN(cx) t3_issue<L_DD1, 0, 64, 0, 32>(cx, 0, 0, false); t3_issue_full<L_RD0>(cx, 64, 32, false); TC_GEMM_END(cx)
        // ================= R2: mean / var second layers ========================================================================
        static_assert(TS(DD_VAR_B0) == TS(DD_MEAN_B0) + 32 && TS(DD_AW_B0) == TS(DD_MEAN_B0) + 64, "dist-decoder biases must be contiguous");
        t3_epilogue<1>(cx.lane_addr, 0, 4, sw + TS(DD_MEAN_B0), 0);        // D[0..63] -> ELU -> A[k 0..63]
        t3_epilogue<1>(cx.lane_addr, 64, 1, sw + TS(RD_B0), 64);           // ray_dir_fc hidden: D[64..79] -> ELU -> A[k 64..79]
        TC_GEMM_BEGIN(cx) t3_issue_full<L_DD2M>(cx, 0, 0, false); t3_issue_full<L_DD2V>(cx, 32, 32, false); TC_GEMM_END(cx)
        // ================= R3: aw first layer (rows 64..95 of the fused first-layer image), ray_dir_fc.2 (16 -> 35) ==========
        // The second-layer outputs of mean / var are pulled into registers first; their third layers (CUDA cores) then run
        // UNDER this round's MMAs instead of in front of them.
        float om0, om1, ov0, ov1;
        {
            float hm[32], hvv[32];
            tm_ld<32>(cx.lane_addr + T3_D + 0, hm);
            tm_ld<32>(cx.lane_addr + T3_D + 32, hvv);
            {
                float ray[32];
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(mine + GN_REC_RAYF + c);   // still in the scratch row
                    ray[c] = t.x; ray[c + 1] = t.y; ray[c + 2] = t.z; ray[c + 3] = t.w;
                }
                t3_store_a<32>(cx.lane_addr, 0, ray);               // A[k 0..31] = ray_feats again (kept for R5)
            }
            tma_rows(GN_REC_HALF, warp_first_row(tile));            // half A is consumed: request half B [rgb, depth | img_feats] into the same rows
            TC_GEMM_BEGIN(cx) t3_issue<L_DD1, 64, 32, 0, 32>(cx, 0, 0, false); t3_issue_full<L_RD1>(cx, 32, 64, false); TC_GEMM_COMMIT(cx)
            bias_elu<32>(sw + TS(DD_MEAN_B2), hm);
            om0 = sw[TS(DD_MEAN_B4)]; om1 = sw[TS(DD_MEAN_B4) + 1];
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float2 w = *reinterpret_cast<const float2*>(sw + TS(DD_MEAN_W4) + k * 4); om0 = fmaf(hm[k], w.x, om0); om1 = fmaf(hm[k], w.y, om1); }
            bias_elu<32>(sw + TS(DD_VAR_B2), hvv);
            ov0 = sw[TS(DD_VAR_B4)]; ov1 = sw[TS(DD_VAR_B4) + 1];
#pragma unroll
            for (int k = 0; k < 32; ++k) { const float2 w = *reinterpret_cast<const float2*>(sw + TS(DD_VAR_W4) + k * 4); ov0 = fmaf(hvv[k], w.x, ov0); ov1 = fmaf(hvv[k], w.y, ov1); }
            TC_GEMM_WAIT(cx)
        }
        // ================= R4: aw second layer; f = [img_feats | rgb] + dir feature ==============================================
        t3_epilogue<1>(cx.lane_addr, 0, 2, sw + TS(DD_AW_B0), 32);          // D[0..31] -> ELU -> A[k 32..63]
        TC_GEMM_BEGIN(cx) t3_issue_full<L_DD2A>(cx, 0, 32, false); TC_GEMM_COMMIT(cx)
        float f[48];                                                        // D[32..79] (other columns than the running MMA's)
        tm_ld<48>(cx.lane_addr + T3_D + 32, f); bias_elu<36>(sw + TS(RD_B1), f);
        // per-point mask word / valid count: loaded at the top of the tile, first used here (its latency sits under rounds 1-3)
        const float mask = (valid && ((__float_as_uint(ptv.y) >> v) & 1u)) ? 1.f : 0.f;
        const float nvalid = ptv.x;
        const float wgt = __fdiv_rn(mask, nvalid + 1e-8f);      // ibrnet.py:466
        mbar_wait(tma_bar, tma_par); tma_par ^= 1u;                         // half B has landed
        const float4 tail = *reinterpret_cast<const float4*>(mine + (GN_REC_RGB - GN_REC_HALF));     // rgb0..2 (masked), depth
        const float depth = tail.w;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(mine + (GN_REC_IMGF - GN_REC_HALF) + c);
            f[c] += t.x; f[c + 1] += t.y; f[c + 2] += t.z; f[c + 3] += t.w;     // ibrnet.py:459
        }
        f[32] += tail.x; f[33] += tail.y; f[34] += tail.z;
        __syncwarp();                                                        // every lane has read its row: the scratch is free for the poolings
#pragma unroll
        for (int c = 35; c < 48; ++c) f[c] = 0.f;
        TC_GEMM_WAIT(cx)
        // aw third layer, compute_prob (dist_decoder.py:109-142)
        float hit, vis;
        {
            float h[32], oa;
            tm_ld<32>(cx.lane_addr + T3_D + 0, h); bias_elu<32>(sw + TS(DD_AW_B2), h);
            oa = sw[TS(DD_AW_B4)];
#pragma unroll
            for (int k = 0; k < 32; ++k) oa = fmaf(h[k], sw[TS(DD_AW_W4) + k * 4], oa);
            const float mean0 = tc_softplus(om0), mean1 = tc_softplus(om1);
            const float var0 = tc_softplus(ov0) + 0.05f, var1 = tc_softplus(ov1) + 0.05f;
            const float aw = tc_sigmoid(oa);
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
            const float c00 = tc_sigmoid(2.f * ((nearp - mean0) * var0)), c10 = tc_sigmoid(2.f * ((farp - mean0) * var0));
            const float c01 = tc_sigmoid(2.f * ((nearp - mean1) * var1)), c11 = tc_sigmoid(2.f * ((farp - mean1) * var1));
            const float mix1 = 1.f - aw;
            vis = ((1.f - c00) * aw + (1.f - c01) * mix1) * mask;
            hit = ((c10 - c00) * aw + (c11 - c01) * mix1) * mask;
        }
        // ================= R5: prob_embed.0 on [ray | 2hit-1 | 2vis-1] (K = 34 -> 48) ==========================================
        {
            float hv[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) hv[c] = 0.f;
            hv[0] = (hit - 0.5f) * 2.f; hv[1] = (vis - 0.5f) * 2.f;
            t3_store_a<16>(cx.lane_addr, 32, hv);
        }
        TC_GEMM_BEGIN(cx) t3_issue_full<L_PE0>(cx, 0, 0, false); TC_GEMM_END(cx)
        // ================= R6: {neuray_fc.0 o prob_embed.2} and base_fc.0's per-view part, both on [f | e1] =====================
        t3_epilogue<2>(cx.lane_addr, 0, 2, sw + TS(PE_B0), 48);           // e1 = ReLU(.) -> A[k 48..79]
        t3_store_a<48>(cx.lane_addr, 0, f);                               // A[k 0..47] = f
        float w0;
        {
            float g0[36], g1[36], tmp[36];
            TC_GEMM_BEGIN(cx) t3_issue_full<L_NFC>(cx, 64, 48, false); t3_issue_full<L_BF0C>(cx, 0, 0, false); TC_GEMM_COMMIT(cx)
            // mean1 (weights w = mask / sum mask) does not need weight0: pooled while the MMAs run (shared memory only)
#pragma unroll
            for (int c = 0; c < 36; ++c) tmp[c] = wgt * f[c];
            t3_pool36(scr, lane, g, v, gb, V, lane_active, tmp, g1);
            TC_GEMM_WAIT(cx)
            {
                float t[16];
                tm_ld<16>(cx.lane_addr + T3_D + 64, t);
                float s = sw[TS(NF_B2)];
#pragma unroll
                for (int k = 0; k < 8; ++k) s = fmaf(tc_elu(t[k] + sw[TS(NFC_B0) + k]), sw[TS(NF_W2) + k], s);
                w0 = tc_sigmoid(s) * wgt;                       // ibrnet.py:469
            }
#pragma unroll
            for (int c = 0; c < 36; ++c) tmp[c] = w0 * f[c];
            t3_pool36(scr, lane, g, v, gb, V, lane_active, tmp, g0);
            // ---- R7a: base_fc.0 on [mean0 | mean1]  (image k 0..63)
            t3_store_a<32>(cx.lane_addr, 0, g0);
            t3_store_a<32>(cx.lane_addr, 32, g1);
            float tl[16];
#pragma unroll
            for (int c = 0; c < 3; ++c) { tl[c] = g0[32 + c]; tl[3 + c] = g1[32 + c]; }
            TC_GEMM_BEGIN(cx) t3_issue<L_BF0B, 0, 64, 0, 64>(cx, 0, 0, true); TC_GEMM_COMMIT(cx)
            // variances, pooled while R7a runs
#pragma unroll
            for (int c = 0; c < 36; ++c) { const float d0 = f[c] - g0[c]; tmp[c] = w0 * d0 * d0; }      // ibrnet.py:115
            t3_pool36(scr, lane, g, v, gb, V, lane_active, tmp, g0);
#pragma unroll
            for (int c = 0; c < 36; ++c) { const float d1 = f[c] - g1[c]; tmp[c] = wgt * d1 * d1; }
            t3_pool36(scr, lane, g, v, gb, V, lane_active, tmp, g1);
            TC_GEMM_WAIT(cx)
            // ---- R7b: base_fc.0 on [var0 | var1 | tails]  (image k 64..143)
            t3_store_a<32>(cx.lane_addr, 0, g0);
            t3_store_a<32>(cx.lane_addr, 32, g1);
#pragma unroll
            for (int c = 0; c < 3; ++c) { tl[6 + c] = g0[32 + c]; tl[9 + c] = g1[32 + c]; }
            tl[12] = 0.f; tl[13] = 0.f; tl[14] = 0.f; tl[15] = 0.f;
            t3_store_a<16>(cx.lane_addr, 64, tl);
        }
        TC_GEMM_BEGIN(cx) t3_issue<L_BF0B, 0, 64, 64, 80>(cx, 0, 0, true); TC_GEMM_END(cx)
        // ================= R8: base_fc.2 ====================================================================
        t3_epilogue<1>(cx.lane_addr, 0, 4, sw + TS(BF_B0C), 0);          // bias includes the folded prob_embed.2 bias
        TC_GEMM_BEGIN(cx) t3_issue_full<L_BF2>(cx, 0, 0, false); TC_GEMM_END(cx)
        // ================= R9 / R10: vis_fc ====================================================================
        float x[36];
        tm_ld<32>(cx.lane_addr + T3_D, x); bias_elu<32>(sw + TS(BF_B2), x);
        x[32] = 0.f; x[33] = 0.f; x[34] = 0.f; x[35] = 0.f;
        {
            float xi[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) xi[c] = x[c] * wgt;
            t3_store_a<32>(cx.lane_addr, 0, xi);
        }
        TC_GEMM_BEGIN(cx) t3_issue_full<L_VF0>(cx, 0, 0, false); TC_GEMM_END(cx)
        t3_epilogue<1>(cx.lane_addr, 0, 2, sw + TS(VF_B0), 32);
        TC_GEMM_BEGIN(cx) t3_issue_full<L_VF2>(cx, 0, 32, false); TC_GEMM_END(cx)
        {
            float xv[48];
            tm_ld<48>(cx.lane_addr + T3_D, xv); bias_elu<36>(sw + TS(VF_B2), xv);
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] += xv[c];
            const float visw = tc_sigmoid(xv[32]) * mask;       // ibrnet.py:478-479
            float xi[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) xi[c] = x[c] * visw;
            t3_store_a<32>(cx.lane_addr, 0, xi);
        }
        // ================= R11: vis_fc2 =======================================================================
        TC_GEMM_BEGIN(cx) t3_issue_full<L_V20>(cx, 0, 0, false); TC_GEMM_END(cx)
        float vis2;
        {
            float t[32];
            tm_ld<32>(cx.lane_addr + T3_D, t); bias_elu<32>(sw + TS(V2_B0), t);
            float s = sw[TS(V2_B2)];
#pragma unroll
            for (int k = 0; k < 32; ++k) s = fmaf(t[k], sw[TS(V2_W2) + k], s);
            vis2 = tc_sigmoid(s) * mask;
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
            t3_pool36(scr, lane, g, v, gb, V, lane_active, tmp, mu);
#pragma unroll
            for (int c = 0; c < 36; ++c) { const float dl = x[c] - mu[c]; tmp[c] = w2 * dl * dl; }
            t3_pool36(scr, lane, g, v, gb, V, lane_active, tmp, vr);
        }
        if (p.status) {
            float chk = fabsf(hit) + fabsf(vis);
            if (writer) {
#pragma unroll
                for (int c = 0; c < 32; ++c) chk += fabsf(mu[c]) + fabsf(vr[c]);
            }
            if (valid && !(chk <= 3.0e38f)) atomicOr(p.status, 1);
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
        if (tile + T3_SLOTS < tile_hi) tma_rows(0, warp_first_row(tile + T3_SLOTS));   // next tile's half A, while this warp wraps up
    }
    if (p.tok) {            // uniform branch: phase 2, geometry_fc per point
        __syncthreads();                                                // every slot is done: all pooled rows of the CTA are written
        if (tid == 0) *s_ctr = 0;
        __syncthreads();
        const long long pt_lo = (long long)tile_lo * 4 * G;
        const long long pt_end = (long long)tile_hi * 4 * G;
        t3_geometry_phase(p, cx.tmem_slot, cx.img_base16, cx.bar, cx.parity, sw, s_ctr, s_tile, slot, warp, lane, pt_lo,
                          pt_end < total_pts ? pt_end : total_pts);
    }
    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*s_tmem), "r"(512));
}

typedef CUresult (*GnEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static GnEncodeTiledFn gn_encode_tiled() {
    static GnEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<GnEncodeTiledFn>(p);
    }
    return fn;
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
        if (!p.pooled) return -9;                     // phase 2 (geometry_fc per point) reads the pooled rows back
    }
    const int G = (32 / p.V) < T3_MAX_G ? (32 / p.V) : T3_MAX_G;
    const long long total = (long long)p.B * p.N;
    const long long per_tile = 4LL * G;
    const long long tiles = (total + per_tile - 1) / per_tile;
    if (tiles > 0x7fffffffLL || total * p.V > 0x7fffffffLL) return -6;
    const size_t smem = t3_smem_bytes(G);
    if (smem > 227 * 1024) return -5;
    // tensor map of the record as a 2-D fp32 tensor [B*N*V rows][72]; box = one 36-float half of a warp's G*V rows
    GnEncodeTiledFn enc = gn_encode_tiled();
    if (!enc) return -10;
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = { (cuuint64_t)GN_REC_STRIDE, (cuuint64_t)(total * p.V) };
        const cuuint64_t gstride[1] = { (cuuint64_t)GN_REC_STRIDE * 4 };
        const cuuint32_t box[2] = { (cuuint32_t)GN_REC_HALF, (cuuint32_t)(G * p.V) };
        const cuuint32_t estr[2] = { 1, 1 };
        if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.rec), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -11;
    }
    static size_t smem_cache[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k2a_tc3_kernel, smem, smem_cache);
    if (e != cudaSuccess) return (int)e;
    const int sms = gn_sm_count();
    const long long want = (tiles + T3_SLOTS - 1) / T3_SLOTS;
    const int grid = (int)(want < sms ? want : sms);
    gn_k2a_tc3_kernel<<<grid, T3_THREADS, smem, (cudaStream_t)stream>>>(p, tmap, (int)tiles, G);
    return (int)cudaGetLastError();
}
