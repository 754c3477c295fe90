// K1: fused project - sample kernel (HBM-bound).
//
// Replaces, per (point, view), the reference op chain
//   project_points_dict        render_ops.py:82-144   (K@Rt projection, validity mask, view dirs,
//                                                      bilinear taps of ray_feats and imgs)
//   get_img_feats              renderer.py:80-88      (bilinear tap of img_feats)
//   get_dir_diff               aggregate_net.py:11-17
//   num_valid_obs / mask       ibrnet.py:466,490      (per-point valid count + view bit mask)
// and writes one 288-byte record per (point, view) plus 8 bytes per point.  (ray_dir_fc and the weighted
// mean/variance poolings of ibrnet.py:457-471 are GEMM / epilogue work and live in K2a: with them inside this
// kernel it was issue-bound at 25 % of the HBM roofline, see profiles/.)
//
// Two implementations of the same arithmetic (cross-checked on the GPU, tests/test_gpu_volume.py):
//
//  gn_k1_kernel (default)  one CTA = 256 threads = a 2x2x8 voxel tile (32 points) x V views.
//    phase A  thread <-> (point, view): projection, mask, view direction, dir_diff, bilinear tap offsets/weights -> shared memory.
//    phase B  8 lanes <-> one point, lane j <-> channels 4j..4j+3: each bilinear tap is ONE 256-byte texel of the fused
//             channels-last feature buffer (2 x LDG.128 per lane, the second at an immediate +128 B), the 4 image taps are
//             4 RGBA texels on lanes 0..3, records leave as coalesced float4 streaming stores.
//
//  gn_k1_walk_kernel (GN_K1_IMPL=3, experiment kept as the on-GPU cross-check)  one CTA = a tile of 16 CONSECUTIVE points
//    of the record order (two z-runs of 8 samples) x V views; an 8-lane group WALKS one (run, view) and keeps the 2x2 tap
//    window of both maps in registers, stored by texel parity so that a window that moved by one texel reloads only the
//    column/row that changed (4.1 instead of 8 gathered lines per (point,view) on the bench scene).  GN_K1_STAGE=1 also
//    assembles the tile's records (16 x V x 288 B, contiguous in HBM) in shared memory and writes them with ONE bulk async
//    copy (cp.async.bulk.global.shared::cta, `UBLKCP` in SASS).  It executes 28 % fewer instructions and half the gathers,
//    but measured SLOWER on B200 (49 / 53 us vs 45 us per 40^3 volume, profiles/k1_variants_r01e.txt): the register window
//    (and, staged, the 27 KB tile buffer) cuts the resident warps from 32 to 24 (18) per SM and the kernel is bound by
//    exposed latency per resident warp, not by the number of gathers.
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"
#include <cstdlib>

// ----------------------------------------------------------------------------------------------------------------------
// per-(point,view) set-up shared by both kernels
struct K1Pair {
    int   fo[4];              // feature-map tap texel indices (y*fw + x) within the view's [fh,fw] map
    float fw_[4];             // feature tap weights * mask   (nw, ne, sw, se)
    int   io[4];              // image tap offsets (pixels) within one H*W plane
    float iw[4];              // image tap weights * mask
    float dd[4];              // dir_diff
    float mask, depth;
    int   x0, y0;             // feature-map corner (index-table dump)
};

// H = K@[R|t] row-major 3x4, c = camera centre.  Fixed op order, no FMA (index-table parity with the oracle).
__device__ __forceinline__ void k1_pair_setup(const float (&Hm)[12], const float (&cc)[3], float px, float py, float pz,
                                              float qx, float qy, float qz, bool live, int H, int W, int fh, int fw, K1Pair& o)
{
    // render_ops.py:94-99, fixed order ((h0*x + h1*y) + h2*z) + h3
    const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hm[0], px), __fmul_rn(Hm[1], py)), __fmul_rn(Hm[2], pz)), Hm[3]);
    const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hm[4], px), __fmul_rn(Hm[5], py)), __fmul_rn(Hm[6], pz)), Hm[7]);
    const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hm[8], px), __fmul_rn(Hm[9], py)), __fmul_rn(Hm[10], pz)), Hm[11]);
    const bool near_zero = fabsf(zc) < 1e-4f;                 // render_ops.py:101
    const float depth = near_zero ? 1e-3f : zc;               // render_ops.py:102
    const float u = __fdiv_rn(xc, depth), w_ = __fdiv_rn(yc, depth);   // render_ops.py:103
    const bool outside = (u < -0.5f) | (u >= (float)W - 0.5f) | (w_ < -0.5f) | (w_ >= (float)H - 0.5f);
    const float mask = (live && !near_zero && !outside) ? 1.f : 0.f;   // render_ops.py:126-128 (no z>0 test)
    o.mask = mask; o.depth = depth;

    // view direction, render_ops.py:112-114
    const float dx = __fsub_rn(px, cc[0]), dy = __fsub_rn(py, cc[1]), dz = __fsub_rn(pz, cc[2]);
    const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-5f);
    const float ex = __fdiv_rn(-dx, nrm), ey = __fdiv_rn(-dy, nrm), ez = __fdiv_rn(-dz, nrm);
    // aggregate_net.py:13-15
    o.dd[0] = __fsub_rn(ex, qx); o.dd[1] = __fsub_rn(ey, qy); o.dd[2] = __fsub_rn(ez, qz);
    o.dd[3] = __fadd_rn(__fadd_rn(__fmul_rn(ex, qx), __fmul_rn(ey, qy)), __fmul_rn(ez, qz));

    // bilinear taps.  feature maps: align_corners=False (map size != image size), images: True
    // (render_ops.py:64-68); both normalised by the IMAGE size (ops.py:29-30); border padding.
    {
        const bool ac = (fh == H) && (fw == W);
        const GnTap1D tx = gn_tap1d(u, W, fw, ac), ty = gn_tap1d(w_, H, fh, ac);
        o.fo[0] = ty.i0 * fw + tx.i0; o.fo[1] = ty.i0 * fw + tx.i1;
        o.fo[2] = ty.i1 * fw + tx.i0; o.fo[3] = ty.i1 * fw + tx.i1;
        o.fw_[0] = __fmul_rn(__fmul_rn(tx.w0, ty.w0), mask); o.fw_[1] = __fmul_rn(__fmul_rn(tx.w1, ty.w0), mask);
        o.fw_[2] = __fmul_rn(__fmul_rn(tx.w0, ty.w1), mask); o.fw_[3] = __fmul_rn(__fmul_rn(tx.w1, ty.w1), mask);
        o.x0 = tx.i0; o.y0 = ty.i0;
    }
    {
        const GnTap1D tx = gn_tap1d(u, W, W, true), ty = gn_tap1d(w_, H, H, true);
        o.io[0] = ty.i0 * W + tx.i0; o.io[1] = ty.i0 * W + tx.i1;
        o.io[2] = ty.i1 * W + tx.i0; o.io[3] = ty.i1 * W + tx.i1;
        o.iw[0] = __fmul_rn(__fmul_rn(tx.w0, ty.w0), mask); o.iw[1] = __fmul_rn(__fmul_rn(tx.w1, ty.w0), mask);
        o.iw[2] = __fmul_rn(__fmul_rn(tx.w0, ty.w1), mask); o.iw[3] = __fmul_rn(__fmul_rn(tx.w1, ty.w1), mask);
    }
}

// bilinear blend in the fixed order both kernels share: ((t0*w0) then fma t1, t2, t3)
__device__ __forceinline__ float k1_blend(float t0, float t1, float t2, float t3, const float4 w) {
    float a = __fmul_rn(t0, w.x);
    a = __fmaf_rn(t1, w.y, a); a = __fmaf_rn(t2, w.z, a); a = __fmaf_rn(t3, w.w, a);
    return a;
}

// ======================================================================================================================
// walking kernel
#define K1W_TILE_P 16          // points per CTA (two runs of 8)
#define K1W_RUN 8
#define K1W_CST_VIEW 20        // floats per view in the constant block: KRt 12 | cam 3 | pad (80-byte stride: conflict-free float4 reads)

// Per (point,view) tap table.  The 2x2 bilinear window is stored by TEXEL PARITY: slot s = (y&1)<<1 | (x&1).  A window that
// moves by one texel replaces exactly the slots of the column/row that left; the others keep their registers - no data moves.
struct K1WInfo {
    unsigned fo[4];           // byte offset of the slot's texel within the scene's feature buffer (view offset included)
    float    fw_[4];          // slot weight * mask
    unsigned io[4];           // byte offset of image tap t's RGBA texel within the scene's images (view offset included)
    float    iw[4];           // image tap weight * mask
    float    dd[4];           // dir_diff            (read back by phase B only in the direct-store variant)
    float    depth, pad[3];
};

__device__ __forceinline__ float4 k1_ld4(const char* base, unsigned off) { return __ldg(reinterpret_cast<const float4*>(base + (size_t)off)); }
__device__ __forceinline__ float4 k1_blend4(const float4 a, const float4 b, const float4 c, const float4 d, const float4 w) {
    float4 o;
    o.x = k1_blend(a.x, b.x, c.x, d.x, w); o.y = k1_blend(a.y, b.y, c.y, d.y, w);
    o.z = k1_blend(a.z, b.z, c.z, d.z, w); o.w = k1_blend(a.w, b.w, c.w, d.w, w);
    return o;
}

// FUSED: both feature maps live in one [B,V,fh,fw,64] buffer (ray_feats | img_feats per texel): one address per tap.
// blockDim.x = 8 * (2V rounded up to a multiple of 4): one 8-lane group per (run, view) unit; NT = its compile-time bound.
// STAGE: records are assembled in shared memory and leave as one bulk async copy per tile; !STAGE: float4 streaming stores
// straight from registers (no staging buffer: more resident CTAs, larger L1).
template <int NT, int MINB, bool FUSED, bool STAGE>
__global__ void __launch_bounds__(NT, MINB)
gn_k1_walk_kernel(const __grid_constant__ GnK1Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int V = p.V, R = p.R;
    const int npair = K1W_TILE_P * V;
    float*   rec_s  = reinterpret_cast<float*>(smem_raw);                                  // [npair][72]  (= the HBM layout)
    K1WInfo* s_info = reinterpret_cast<K1WInfo*>(rec_s + (STAGE ? (size_t)npair * GN_REC_STRIDE : 0));   // [npair]
    float*   s_mask = reinterpret_cast<float*>(s_info + npair);                            // [npair]
    float*   s_cst  = s_mask + npair;                                                      // [V][20] | bbox 4 | axis R

    const int tiles_per_scene = p.tiles_per_scene;
    const int b = blockIdx.x / tiles_per_scene;
    const int tile = blockIdx.x - b * tiles_per_scene;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n0 = tile * K1W_TILE_P;

    // ---- constants to shared memory ---------------------------------------------------------------------------------
    for (int i = tid; i < V * 12; i += nthr) {
        const int v = i / 12;
        s_cst[v * K1W_CST_VIEW + (i - v * 12)] = __ldg(p.KRt + (size_t)b * V * 12 + i);
    }
    for (int i = tid; i < V * 3; i += nthr) {
        const int v = i / 3;
        s_cst[v * K1W_CST_VIEW + 12 + (i - v * 3)] = __ldg(p.cam + (size_t)b * V * 3 + i);
    }
    if (p.volume_mode) {
        if (tid < 3) s_cst[V * K1W_CST_VIEW + tid] = __ldg(p.bbox_min + b * 3 + tid);
        for (int i = tid; i < R; i += nthr) s_cst[V * K1W_CST_VIEW + 4 + i] = __ldg(p.axis + i);
    }
    __syncthreads();

    // =============================== phase A ==========================================================================
    for (int q = tid; q < npair; q += nthr) {
        const int pl = q / V;
        const int v = q - pl * V;
        int n = n0 + pl;
        const bool live = n < p.N;
        n = min(n, p.N - 1);
        float px, py, pz, qx = 0.f, qy = 0.f, qz = 1.f;   // que_dir = (0,0,1) in volume mode, renderer.py:179
        if (p.volume_mode) {
            // record index n = (i*R + j)*R + (R-1-k)   (renderer.py:169-170: reshape (1,R*R,R,3), flip the sample axis)
            const int ij = n / R, kk = n - ij * R;
            const int i = ij / R, j = ij - i * R, k = R - 1 - kk;
            const float* ax = s_cst + V * K1W_CST_VIEW + 4;
            // field_utils.py:17-27 table (host-built, fp32) + bbox3d[0] in fp32 (renderer.py:167-168)
            px = __fadd_rn(ax[i], s_cst[V * K1W_CST_VIEW + 0]);
            py = __fadd_rn(ax[j], s_cst[V * K1W_CST_VIEW + 1]);
            pz = __fadd_rn(ax[k], s_cst[V * K1W_CST_VIEW + 2]);
        } else {
            const float* pp = p.pts + ((size_t)b * p.N + n) * 3;
            px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
            const float* d = p.que_dir + ((size_t)b * (p.N / p.dn) + n / p.dn) * 3;
            qx = __ldg(d); qy = __ldg(d + 1); qz = __ldg(d + 2);
        }
        float Hm[12], cc[3];
        {
            const float4* c4 = reinterpret_cast<const float4*>(s_cst + v * K1W_CST_VIEW);
            const float4 a0 = c4[0], a1 = c4[1], a2 = c4[2], a3 = c4[3];
            Hm[0] = a0.x; Hm[1] = a0.y; Hm[2] = a0.z; Hm[3] = a0.w; Hm[4] = a1.x; Hm[5] = a1.y; Hm[6] = a1.z; Hm[7] = a1.w;
            Hm[8] = a2.x; Hm[9] = a2.y; Hm[10] = a2.z; Hm[11] = a2.w; cc[0] = a3.x; cc[1] = a3.y; cc[2] = a3.z;
        }
        K1Pair o;
        k1_pair_setup(Hm, cc, px, py, pz, qx, qy, qz, live, p.H, p.W, p.fh, p.fw, o);
        {
            // tap t -> parity slot t ^ s0, s0 = parity of the window corner.  (At the border a clamped tap repeats its
            // neighbour's texel with weight exactly 0; it still gets its own slot, so the mapping stays a permutation.)
            const unsigned tb = (unsigned)p.feat_stride * 4u, vb = (unsigned)v * (unsigned)(p.fh * p.fw) * tb;
            const bool sx = o.x0 & 1, sy = o.y0 & 1;
            unsigned f0 = vb + (unsigned)o.fo[0] * tb, f1 = vb + (unsigned)o.fo[1] * tb, f2 = vb + (unsigned)o.fo[2] * tb, f3 = vb + (unsigned)o.fo[3] * tb;
            float w0 = o.fw_[0], w1 = o.fw_[1], w2 = o.fw_[2], w3 = o.fw_[3];
            if (sx) { unsigned t; float u; t = f0; f0 = f1; f1 = t; t = f2; f2 = f3; f3 = t; u = w0; w0 = w1; w1 = u; u = w2; w2 = w3; w3 = u; }
            if (sy) { unsigned t; float u; t = f0; f0 = f2; f2 = t; t = f1; f1 = f3; f3 = t; u = w0; w0 = w2; w2 = u; u = w1; w1 = w3; w3 = u; }
            *reinterpret_cast<uint4*>(s_info[q].fo) = make_uint4(f0, f1, f2, f3);
            *reinterpret_cast<float4*>(s_info[q].fw_) = make_float4(w0, w1, w2, w3);
            const unsigned ib = (unsigned)v * (unsigned)(p.H * p.W) * 16u;
            *reinterpret_cast<uint4*>(s_info[q].io) = make_uint4(ib + (unsigned)o.io[0] * 16u, ib + (unsigned)o.io[1] * 16u,
                                                                 ib + (unsigned)o.io[2] * 16u, ib + (unsigned)o.io[3] * 16u);
            *reinterpret_cast<float4*>(s_info[q].iw) = *reinterpret_cast<const float4*>(o.iw);
        }
        s_mask[q] = o.mask;
        if (STAGE) {
            float* row = rec_s + (size_t)q * GN_REC_STRIDE;
            st4(row + GN_REC_RGB, make_float4(0.f, 0.f, 0.f, o.depth));       // rgb is filled in by phase B
            st4(row + GN_REC_DD, make_float4(o.dd[0], o.dd[1], o.dd[2], o.dd[3]));
        } else {
            *reinterpret_cast<float4*>(s_info[q].dd) = make_float4(o.dd[0], o.dd[1], o.dd[2], o.dd[3]);
            s_info[q].depth = o.depth;
        }
        if (p.dbg_feat_idx && live) {   // optional index-table dump for the bit-exactness tests
            int* od = p.dbg_feat_idx + (((size_t)b * p.N + n) * V + v) * 2;
            od[0] = o.x0; od[1] = o.y0;
        }
    }
    __syncthreads();

    // per-point valid count / view bit mask (ibrnet.py:466,490)
    if (tid < K1W_TILE_P && n0 + tid < p.N) {
        float nvalid = 0.f;
        unsigned bits = 0u;
        for (int v = 0; v < V; ++v) {
            const float m = s_mask[tid * V + v];
            nvalid += m;
            bits |= (m != 0.f ? 1u : 0u) << v;
        }
        float2 o2; o2.x = nvalid; o2.y = __uint_as_float(bits);
        *reinterpret_cast<float2*>(p.pt + ((size_t)b * p.N + n0 + tid) * GN_PT_STRIDE) = o2;
    }

    // =============================== phase B ==========================================================================
    // 8-lane group <-> unit u = (view v = u >> 1, run = u & 1); lane j <-> channels 4j..4j+3 of both maps.
    // The tap table and the image texels of sample z+1 are fetched while sample z is blended.
    {
        const int unit = tid >> 3, j = tid & 7;
        const bool on = unit < 2 * V;
        const int v = on ? (unit >> 1) : 0, run = unit & 1;
        const size_t fmap_bytes = (size_t)p.fh * p.fw * p.feat_stride * 4;
        const char* rf_scene = reinterpret_cast<const char*>(p.ray_feats) + (size_t)b * V * fmap_bytes + j * 16;
        const char* if_scene = FUSED ? rf_scene + 128 : reinterpret_cast<const char*>(p.img_feats) + (size_t)b * V * fmap_bytes + j * 16;
        const char* im_scene = reinterpret_cast<const char*>(p.imgs) + (size_t)b * V * p.H * p.W * 16;   // RGBA-interleaved [B,V,H,W,4]
        int q = run * K1W_RUN * V + v;
        const int nrow0 = n0 + run * K1W_RUN;                      // first point of the run
        float* grow = p.rec + (((size_t)b * p.N + nrow0) * V + v) * GN_REC_STRIDE;      // (!STAGE) HBM row of the current sample
        unsigned c0 = 0xffffffffu, c1 = 0xffffffffu, c2 = 0xffffffffu, c3 = 0xffffffffu;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0, g0 = r0, g1 = r0, g2 = r0, g3 = r0;
        uint4 fo = *reinterpret_cast<const uint4*>(s_info[q].fo);
        float4 w = *reinterpret_cast<const float4*>(s_info[q].fw_);
        float4 tx = make_float4(0.f, 0.f, 0.f, 0.f);
        float iw = 0.f;
        if (j < 4) { iw = s_info[q].iw[j]; tx = k1_ld4(im_scene, s_info[q].io[j]); }
#pragma unroll 2
        for (int z = 0; z < K1W_RUN; ++z) {
            // this sample's window: load only the slots whose texel changed
            if (fo.x != c0) { r0 = k1_ld4(rf_scene, fo.x); g0 = k1_ld4(if_scene, fo.x); c0 = fo.x; }
            if (fo.y != c1) { r1 = k1_ld4(rf_scene, fo.y); g1 = k1_ld4(if_scene, fo.y); c1 = fo.y; }
            if (fo.z != c2) { r2 = k1_ld4(rf_scene, fo.z); g2 = k1_ld4(if_scene, fo.z); c2 = fo.z; }
            if (fo.w != c3) { r3 = k1_ld4(rf_scene, fo.w); g3 = k1_ld4(if_scene, fo.w); c3 = fo.w; }
            // next sample's tap table and image texels
            const int qn = q + (z + 1 < K1W_RUN ? V : 0);
            const uint4 fo_n = *reinterpret_cast<const uint4*>(s_info[qn].fo);
            const float4 w_n = *reinterpret_cast<const float4*>(s_info[qn].fw_);
            float4 tx_n = make_float4(0.f, 0.f, 0.f, 0.f);
            float iw_n = 0.f;
            if (j < 4) { iw_n = s_info[qn].iw[j]; tx_n = k1_ld4(im_scene, s_info[qn].io[j]); }
            // image taps: lane j<4 holds tap j as one RGBA texel; scaled and summed over lanes 0..3 with two xor-shuffles
            float cr = __fmul_rn(tx.x, iw), cg = __fmul_rn(tx.y, iw), cb = __fmul_rn(tx.z, iw);
            cr = __fadd_rn(cr, __shfl_xor_sync(0xffffffffu, cr, 1)); cg = __fadd_rn(cg, __shfl_xor_sync(0xffffffffu, cg, 1)); cb = __fadd_rn(cb, __shfl_xor_sync(0xffffffffu, cb, 1));
            cr = __fadd_rn(cr, __shfl_xor_sync(0xffffffffu, cr, 2)); cg = __fadd_rn(cg, __shfl_xor_sync(0xffffffffu, cg, 2)); cb = __fadd_rn(cb, __shfl_xor_sync(0xffffffffu, cb, 2));
            const float4 ray = k1_blend4(r0, r1, r2, r3, w), img = k1_blend4(g0, g1, g2, g3, w);
            if (STAGE) {
                if (on) {
                    float* row = rec_s + (size_t)q * GN_REC_STRIDE;
                    st4(row + GN_REC_RAYF + 4 * j, ray);
                    st4(row + GN_REC_IMGF + 4 * j, img);
                    if (j == 0) { *reinterpret_cast<float2*>(row + GN_REC_RGB) = make_float2(cr, cg); row[GN_REC_RGB + 2] = cb; }
                }
            } else if (on && nrow0 + z < p.N) {
                st4_cs(grow + GN_REC_RAYF + 4 * j, ray);
                st4_cs(grow + GN_REC_IMGF + 4 * j, img);
                // tail: lanes 0 and 1 write the two adjacent 16-byte chunks [64,68) and [68,72) with ONE store instruction
                if (j < 2) st4_cs(grow + GN_REC_RGB + 4 * j, j == 0 ? make_float4(cr, cg, cb, s_info[q].depth) : *reinterpret_cast<const float4*>(s_info[q].dd));
            }
            grow += (size_t)V * GN_REC_STRIDE;
            q = qn; fo = fo_n; w = w_n; tx = tx_n; iw = iw_n;
        }
    }
    if (!STAGE) return;

    // =============================== store ============================================================================
    // generic-proxy writes to shared memory -> visible to the async proxy, then one bulk copy of the live prefix
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const int nlive = min(K1W_TILE_P, p.N - n0);
        const unsigned bytes = (unsigned)nlive * (unsigned)V * GN_REC_STRIDE * 4u;
        float* dst = p.rec + ((size_t)b * p.N + n0) * V * GN_REC_STRIDE;
        const unsigned src = (unsigned)__cvta_generic_to_shared(rec_s);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // shared memory must outlive the copy's reads
    }
}

// ======================================================================================================================
// round-1 kernel (GN_K1_IMPL=2): one CTA = 256 threads = a 2x2x8 voxel tile (32 points) x V views
//   phase A  thread <-> (point, view) -> shared memory;  phase B  8 lanes <-> one point, lane j <-> channels 4j..4j+3
#define K1_THREADS 256
#define K1_TILE_P 32

struct K1PairInfo {           // 80-byte stride: the four groups of a warp read pairs V apart; 64 B put them on the same banks
    int   fo[4];
    float fw_[4];
    int   io[4];
    float iw[4];
    float pad[4];
};
#define K1_MISC 12            // floats per pair in s_misc: dd0..3, mask, depth, pad

template <bool FUSED, bool TEXPF>
__global__ void __launch_bounds__(K1_THREADS, 4)
gn_k1_kernel(const __grid_constant__ GnK1Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int V = p.V;
    const int npair = K1_TILE_P * V;
    K1PairInfo* s_info = reinterpret_cast<K1PairInfo*>(smem_raw);                    // [npair]
    float* s_misc  = reinterpret_cast<float*>(s_info + npair);                       // [npair][12]: dd0..3, mask, depth, pad

    const int tiles_per_scene = p.tiles_per_scene;          // CTAs per scene
    const int b = blockIdx.x / tiles_per_scene;
    const int tile = blockIdx.x - b * tiles_per_scene;
    const int tid = threadIdx.x;
    const int R = p.R;

    // volume mode: tile = 2x2x8 block of voxels (i,j,k); ray mode: tile = 32 consecutive points n of the explicit pts array
    int tk = 0, tj = 0, ti = 0;
    if (p.volume_mode) {
        const int nz = R >> 3, ny = R >> 1;
        tk = tile % nz; tj = (tile / nz) % ny; ti = tile / (nz * ny);
    }

    for (int pair = tid; pair < npair; pair += K1_THREADS) {
        const int pl = pair / V;
        const int v = pair - pl * V;
        float px, py, pz, qx = 0.f, qy = 0.f, qz = 1.f;
        bool live = true;
        int n;
        if (p.volume_mode) {
            const int i = ti * 2 + (pl >> 4), j = tj * 2 + ((pl >> 3) & 1), k = tk * 8 + (pl & 7);
            px = __fadd_rn(__ldg(p.axis + i), __ldg(p.bbox_min + b * 3 + 0));
            py = __fadd_rn(__ldg(p.axis + j), __ldg(p.bbox_min + b * 3 + 1));
            pz = __fadd_rn(__ldg(p.axis + k), __ldg(p.bbox_min + b * 3 + 2));
            n = (i * R + j) * R + (R - 1 - k);
        } else {
            n = tile * K1_TILE_P + pl;
            live = n < p.N;
            n = min(n, p.N - 1);
            const float* q = p.pts + ((size_t)b * p.N + n) * 3;
            px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
            const float* d = p.que_dir + ((size_t)b * (p.N / p.dn) + n / p.dn) * 3;
            qx = __ldg(d); qy = __ldg(d + 1); qz = __ldg(d + 2);
        }
        float Hm[12], cc[3];
#pragma unroll
        for (int i = 0; i < 12; ++i) Hm[i] = __ldg(p.KRt + ((size_t)b * V + v) * 12 + i);
#pragma unroll
        for (int i = 0; i < 3; ++i) cc[i] = __ldg(p.cam + ((size_t)b * V + v) * 3 + i);
        K1Pair o;
        k1_pair_setup(Hm, cc, px, py, pz, qx, qy, qz, live, p.H, p.W, p.fh, p.fw, o);
        if (p.dbg_feat_idx && live) {
            int* od = p.dbg_feat_idx + (((size_t)b * p.N + n) * V + v) * 2;
            od[0] = o.x0; od[1] = o.y0;
        }
        *reinterpret_cast<int4*>(s_info[pair].fo) = *reinterpret_cast<const int4*>(o.fo);
        *reinterpret_cast<float4*>(s_info[pair].fw_) = *reinterpret_cast<const float4*>(o.fw_);
        *reinterpret_cast<int4*>(s_info[pair].io) = *reinterpret_cast<const int4*>(o.io);
        *reinterpret_cast<float4*>(s_info[pair].iw) = *reinterpret_cast<const float4*>(o.iw);
        float* ms = s_misc + pair * K1_MISC;
        st4(ms, make_float4(o.dd[0], o.dd[1], o.dd[2], o.dd[3]));
        ms[4] = o.mask; ms[5] = o.depth;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, j = lane & 7;
    const int pl = warp * 4 + grp;                 // local point 0..31
    int n;
    bool live = true;
    if (p.volume_mode) {
        const int i = ti * 2 + (pl >> 4), jj = tj * 2 + ((pl >> 3) & 1), k = tk * 8 + (pl & 7);
        n = (i * R + jj) * R + (R - 1 - k);
    } else {
        n = tile * K1_TILE_P + pl;
        live = n < p.N;
        n = min(n, p.N - 1);
    }
    float nvalid = 0.f;
    unsigned bits = 0u;
    for (int v = 0; v < V; ++v) {
        const float m = s_misc[(pl * V + v) * K1_MISC + 4];
        nvalid += m;
        bits |= (m != 0.f ? 1u : 0u) << v;
    }
    float* rec = p.rec + ((size_t)b * p.N + n) * V * GN_REC_STRIDE;
    const int fs = p.feat_stride;
    const size_t fmap_sz = (size_t)p.fh * p.fw * fs;
    const size_t plane = (size_t)p.H * p.W;
    const float* rf_base = p.ray_feats + (size_t)b * V * fmap_sz + 4 * j;
    const float* if_base = p.img_feats + (size_t)b * V * fmap_sz + 4 * j;
    const float* im_base = p.imgs + (size_t)b * V * plane * 4;

    // image taps: lane j<4 fetches tap j as one RGBA texel and scales it; summed over lanes 0..3 with two xor-shuffles.
    // The texels of view v+1 are requested while view v is blended: they are the gathers that miss to DRAM (the images are
    // read once per scene), one full iteration ahead hides their latency behind the feature gathers of the current view.
    float4 px = make_float4(0.f, 0.f, 0.f, 0.f);
    float iw = 0.f;
    if (TEXPF && j < 4) {
        iw = s_info[pl * V].iw[j];
        px = ldg4(im_base + (size_t)s_info[pl * V].io[j] * 4);
    }
    for (int v = 0; v < V; ++v) {
        const int pair = pl * V + v;
        int4 fo = *reinterpret_cast<const int4*>(s_info[pair].fo);
        fo.x *= fs; fo.y *= fs; fo.z *= fs; fo.w *= fs;
        const float4 fwt = *reinterpret_cast<const float4*>(s_info[pair].fw_);
        const float* rf = rf_base + (size_t)v * fmap_sz;
        const float* imf = FUSED ? rf + GN_FEAT_C : if_base + (size_t)v * fmap_sz;
        const float4 r0 = ldg4(rf + fo.x), r1 = ldg4(rf + fo.y), r2 = ldg4(rf + fo.z), r3 = ldg4(rf + fo.w);
        const float4 g0 = ldg4(imf + fo.x), g1 = ldg4(imf + fo.y), g2 = ldg4(imf + fo.z), g3 = ldg4(imf + fo.w);
        float4 px_n = make_float4(0.f, 0.f, 0.f, 0.f);
        float iw_n = 0.f;
        if (TEXPF) {
            if (j < 4 && v + 1 < V) {
                iw_n = s_info[pair + 1].iw[j];
                px_n = ldg4(im_base + ((size_t)(v + 1) * plane + s_info[pair + 1].io[j]) * 4);
            }
        } else if (j < 4) {
            iw = s_info[pair].iw[j];
            px = ldg4(im_base + ((size_t)v * plane + s_info[pair].io[j]) * 4);
        }
        float cr = __fmul_rn(px.x, iw), cg = __fmul_rn(px.y, iw), cb = __fmul_rn(px.z, iw);
        if (TEXPF) { px = px_n; iw = iw_n; }
        float4 ray, img;
        ray.x = k1_blend(r0.x, r1.x, r2.x, r3.x, fwt); ray.y = k1_blend(r0.y, r1.y, r2.y, r3.y, fwt);
        ray.z = k1_blend(r0.z, r1.z, r2.z, r3.z, fwt); ray.w = k1_blend(r0.w, r1.w, r2.w, r3.w, fwt);
        img.x = k1_blend(g0.x, g1.x, g2.x, g3.x, fwt); img.y = k1_blend(g0.y, g1.y, g2.y, g3.y, fwt);
        img.z = k1_blend(g0.z, g1.z, g2.z, g3.z, fwt); img.w = k1_blend(g0.w, g1.w, g2.w, g3.w, fwt);
        float* row = rec + (size_t)v * GN_REC_STRIDE;
        cr = __fadd_rn(cr, __shfl_xor_sync(0xffffffffu, cr, 1)); cg = __fadd_rn(cg, __shfl_xor_sync(0xffffffffu, cg, 1)); cb = __fadd_rn(cb, __shfl_xor_sync(0xffffffffu, cb, 1));
        cr = __fadd_rn(cr, __shfl_xor_sync(0xffffffffu, cr, 2)); cg = __fadd_rn(cg, __shfl_xor_sync(0xffffffffu, cg, 2)); cb = __fadd_rn(cb, __shfl_xor_sync(0xffffffffu, cb, 2));
        if (live) {
            st4_cs(row + GN_REC_RAYF + 4 * j, ray);
            st4_cs(row + GN_REC_IMGF + 4 * j, img);
            if (j < 2) {
                const float4 ddq = *reinterpret_cast<const float4*>(s_misc + pair * K1_MISC);
                st4_cs(row + GN_REC_RGB + 4 * j, j == 0 ? make_float4(cr, cg, cb, s_misc[pair * K1_MISC + 5]) : ddq);
            }
        }
    }
    if (live && j == 0) {
        float2 o; o.x = nvalid; o.y = __uint_as_float(bits);
        *reinterpret_cast<float2*>(p.pt + ((size_t)b * p.N + n) * GN_PT_STRIDE) = o;
    }
}

// ----------------------------------------------------------------------------------------------------------------------
static int k1_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

extern "C" int gn_k1_forward(const GnK1Params* hp, void* stream)
{
    GnK1Params p = *hp;
    if (p.V < 1 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (p.volume_mode) {
        if ((p.R % 8) != 0 || p.N != p.R * p.R * p.R || !p.axis || !p.bbox_min) return -3;
    } else {
        if (!p.pts || !p.que_dir || p.dn < 1 || (p.N % p.dn) != 0) return -4;
    }
    if (p.feat_stride == 0) p.feat_stride = GN_FEAT_C;
    const bool fused = p.feat_stride == 2 * GN_FEAT_C;
    if (p.feat_stride != GN_FEAT_C && !(fused && p.img_feats == p.ray_feats + GN_FEAT_C)) return -7;
    const int impl = k1_env_int("GN_K1_IMPL", 2) == 3 ? 3 : 2;          // read per call: tests switch it at run time
    cudaError_t e;
    if (impl == 3) {
        p.tiles_per_scene = (p.N + K1W_TILE_P - 1) / K1W_TILE_P;
        const int npair = K1W_TILE_P * p.V;
        const bool stage = k1_env_int("GN_K1_STAGE", 0) != 0;
        const size_t smem = (size_t)npair * ((stage ? GN_REC_STRIDE * 4 : 0) + sizeof(K1WInfo) + 4)
                          + (size_t)(p.V * K1W_CST_VIEW + 4 + (p.volume_mode ? p.R : 0)) * 4 + 16;
        const unsigned nthr = 32u * (unsigned)((2 * p.V + 3) / 4);          // one 8-lane group per (run, view) unit
        if (smem > 227 * 1024) return -5;
        if ((double)p.V * p.fh * p.fw * p.feat_stride * 4.0 >= 4294967296.0 || (double)p.V * p.H * p.W * 16.0 >= 4294967296.0) return -8;   // 32-bit tap byte offsets
        const long long grid = (long long)p.B * p.tiles_per_scene;
        if (grid > 0x7fffffffLL) return -6;
#define K1W_LAUNCH(NT, MINB, FU, ST) { static size_t cache[16] = {0}; \
            e = gn_ensure_smem(gn_k1_walk_kernel<NT, MINB, FU, ST>, smem, cache); if (e != cudaSuccess) return (int)e; \
            gn_k1_walk_kernel<NT, MINB, FU, ST><<<(unsigned)grid, nthr, smem, (cudaStream_t)stream>>>(p); }
#define K1W_LAUNCH_F(NT, MINB, ST) { if (fused) K1W_LAUNCH(NT, MINB, true, ST) else K1W_LAUNCH(NT, MINB, false, ST) }
        if (stage) {
            if (nthr <= 96)       K1W_LAUNCH_F(96, 6, true)
            else if (nthr <= 192) K1W_LAUNCH_F(192, 3, true)
            else                  K1W_LAUNCH_F(512, 1, true)
        } else {
            if (nthr <= 96)       K1W_LAUNCH_F(96, 8, false)
            else if (nthr <= 192) K1W_LAUNCH_F(192, 4, false)
            else                  K1W_LAUNCH_F(512, 1, false)
        }
#undef K1W_LAUNCH_F
#undef K1W_LAUNCH
        return (int)cudaGetLastError();
    }
    if (p.volume_mode) p.tiles_per_scene = (p.R / 2) * (p.R / 2) * (p.R / 8);
    else               p.tiles_per_scene = (p.N + K1_TILE_P - 1) / K1_TILE_P;
    const int npair = K1_TILE_P * p.V;
    const size_t smem = (size_t)npair * (sizeof(K1PairInfo) + K1_MISC * sizeof(float));
    if (smem > 227 * 1024) return -5;
    const long long grid = (long long)p.B * p.tiles_per_scene;
    if (grid > 0x7fffffffLL) return -6;
    const bool texpf = k1_env_int("GN_K1_TEXPF", 1) != 0;
#define K1T_LAUNCH(FU, PF) { static size_t cache[16] = {0}; \
        e = gn_ensure_smem(gn_k1_kernel<FU, PF>, smem, cache); if (e != cudaSuccess) return (int)e; \
        gn_k1_kernel<FU, PF><<<(unsigned)grid, K1_THREADS, smem, (cudaStream_t)stream>>>(p); }
    if (fused) { if (texpf) K1T_LAUNCH(true, true) else K1T_LAUNCH(true, false) }
    else       { if (texpf) K1T_LAUNCH(false, true) else K1T_LAUNCH(false, false) }
#undef K1T_LAUNCH
    return (int)cudaGetLastError();
}
