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
//  gn_k1_kernel  one CTA = 256 threads = a 2x2x8 voxel tile (32 points) x V views.
//    phase A  thread <-> (point, view): projection, mask, view direction, dir_diff, bilinear tap offsets/weights -> shared memory.
//    phase B  8 lanes <-> one point, lane j <-> channels 4j..4j+3: each bilinear tap is ONE 256-byte texel of the fused
//             channels-last feature buffer (2 x LDG.128 per lane, the second at an immediate +128 B), the 4 image taps are
//             4 RGBA texels on lanes 0..3 (fp32, or uint8 divided by 255 in the kernel: main.py:170 color_map_forward),
//             records leave as coalesced float4 streaming stores.
//
// Variants that were built, measured on B200 and found slower are kept as source records under profiles/experiments/
// (k1_project_sample_bulk_walk_r02a.cu.txt): a register-window "walk" along z (4.1 instead of 8 gathered lines per (point,view),
// 49 us vs 45 us: occupancy), records staged in shared memory and written by cp.async.bulk (53 us per tile / 49 us per warp row).
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"

// ----------------------------------------------------------------------------------------------------------------------
// per-(point,view) set-up
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

// bilinear blend in a fixed order: ((t0*w0) then fma t1, t2, t3)
__device__ __forceinline__ float k1_blend(float t0, float t1, float t2, float t3, const float4 w) {
    float a = __fmul_rn(t0, w.x);
    a = __fmaf_rn(t1, w.y, a); a = __fmaf_rn(t2, w.z, a); a = __fmaf_rn(t3, w.w, a);
    return a;
}

// ======================================================================================================================
// one CTA = 256 threads = a 2x2x8 voxel tile (32 points) x V views
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

// one image tap as (r,g,b,-): fp32 RGBA texel, or uint8 RGBA texel / 255 (exactly np.float32(u8) / 255, main.py:170)
template <bool U8>
__device__ __forceinline__ float4 k1_img_texel(const void* imgs, size_t texel) {
    if (U8) {
        const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(imgs) + texel);
        return make_float4(__fdiv_rn((float)t.x, 255.f), __fdiv_rn((float)t.y, 255.f), __fdiv_rn((float)t.z, 255.f), 0.f);
    }
    return __ldg(reinterpret_cast<const float4*>(imgs) + texel);
}

template <bool FUSED, bool U8>
__global__ void __launch_bounds__(K1_THREADS, 4)
gn_k1_kernel(const __grid_constant__ GnK1Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int V = p.V;
    const int npair = K1_TILE_P * V;
    K1PairInfo* s_info = reinterpret_cast<K1PairInfo*>(smem_raw);                    // [npair]
    float* s_misc  = reinterpret_cast<float*>(s_info + npair);                       // [npair][12]: dd0..3, mask, depth, pad
    __shared__ int s_cnt[32];                                                        // valid projections per view (optional diagnostic)
    if (p.valid_count && threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
    if (p.valid_count) __syncthreads();

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
        float px, py, pz, qx = 0.f, qy = 0.f, qz = 1.f;   // que_dir = (0,0,1) in volume mode, renderer.py:179
        bool live = true;
        int n;
        if (p.volume_mode) {
            // record index n = (i*R + j)*R + (R-1-k)   (renderer.py:169-170: reshape (1,R*R,R,3), flip the sample axis);
            // point = field_utils.py:17-27 table (host-built, fp32) + bbox3d[0] in fp32 (renderer.py:167-168)
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
        if (p.dbg_feat_idx && live) {   // optional index-table dump for the bit-exactness tests
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
        if (p.valid_count && o.mask != 0.f) atomicAdd(&s_cnt[v], 1);
    }
    __syncthreads();
    if (p.valid_count && tid < V && s_cnt[tid]) atomicAdd(p.valid_count + b * V + tid, s_cnt[tid]);

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
    // per-point valid count / view bit mask (ibrnet.py:466,490)
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
    const size_t im_base = (size_t)b * V * plane;      // texel index of this scene's first view

    // image taps: lane j<4 fetches tap j as one RGBA texel and scales it; summed over lanes 0..3 with two xor-shuffles.
    // The texels of view v+1 are requested while view v is blended: they are the gathers that miss to DRAM (the images are
    // read once per scene), one full iteration ahead hides their latency behind the feature gathers of the current view.
    float4 px = make_float4(0.f, 0.f, 0.f, 0.f);
    float iw = 0.f;
    if (j < 4) {
        iw = s_info[pl * V].iw[j];
        px = k1_img_texel<U8>(p.imgs, im_base + (size_t)s_info[pl * V].io[j]);
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
        if (j < 4 && v + 1 < V) {
            iw_n = s_info[pair + 1].iw[j];
            px_n = k1_img_texel<U8>(p.imgs, im_base + (size_t)(v + 1) * plane + (size_t)s_info[pair + 1].io[j]);
        }
        float cr = __fmul_rn(px.x, iw), cg = __fmul_rn(px.y, iw), cb = __fmul_rn(px.z, iw);
        px = px_n; iw = iw_n;
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
            if (j < 2) {       // lanes 0 and 1 write the two adjacent 16-byte chunks dir_diff [32,36) and rgb|depth [36,40) with ONE store instruction
                const float4 ddq = *reinterpret_cast<const float4*>(s_misc + pair * K1_MISC);
                st4_cs(row + GN_REC_DD + 4 * j, j == 0 ? ddq : make_float4(cr, cg, cb, s_misc[pair * K1_MISC + 5]));
            }
        }
    }
    if (live && j == 0) {
        float2 o; o.x = nvalid; o.y = __uint_as_float(bits);
        *reinterpret_cast<float2*>(p.pt + ((size_t)b * p.N + n) * GN_PT_STRIDE) = o;
    }
}

// ----------------------------------------------------------------------------------------------------------------------
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
    if (p.img_u8 != 0 && p.img_u8 != 1) return -8;
    if (p.volume_mode) p.tiles_per_scene = (p.R / 2) * (p.R / 2) * (p.R / 8);
    else               p.tiles_per_scene = (p.N + K1_TILE_P - 1) / K1_TILE_P;
    const int npair = K1_TILE_P * p.V;
    const size_t smem = (size_t)npair * (sizeof(K1PairInfo) + K1_MISC * sizeof(float));
    if (smem > 227 * 1024) return -5;
    const long long grid = (long long)p.B * p.tiles_per_scene;
    if (grid > 0x7fffffffLL) return -6;
    cudaError_t e;
#define K1T_LAUNCH(FU, U8) { static size_t cache[16] = {0}; \
        e = gn_ensure_smem(gn_k1_kernel<FU, U8>, smem, cache); if (e != cudaSuccess) return (int)e; \
        gn_k1_kernel<FU, U8><<<(unsigned)grid, K1_THREADS, smem, (cudaStream_t)stream>>>(p); }
    if (fused) { if (p.img_u8) K1T_LAUNCH(true, true) else K1T_LAUNCH(true, false) }
    else       { if (p.img_u8) K1T_LAUNCH(false, true) else K1T_LAUNCH(false, false) }
#undef K1T_LAUNCH
    return (int)cudaGetLastError();
}
