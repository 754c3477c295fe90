// K1: fused project - sample - aggregate kernel (HBM-bound).
//
// Replaces, per (point, view), the reference op chain
//   project_points_dict        render_ops.py:82-144   (K@Rt projection, validity mask, view dirs,
//                                                      bilinear taps of ray_feats and imgs)
//   get_img_feats              renderer.py:80-88      (bilinear tap of img_feats)
//   get_dir_diff               aggregate_net.py:11-17
//   ray_dir_fc + add           ibrnet.py:457-459
//   mask-weighted mean/var     ibrnet.py:466,471 (mean1,var1), fused_mean_variance ibrnet.py:112-116
// and writes one 288-byte record per (point, view) plus one 288-byte record per point.
//
// Work decomposition (one CTA = 256 threads = one tile of 32 points x V views):
//   phase A  thread <-> (point, view): projection, mask, view direction, dir_diff, ray_dir_fc (4->16->35),
//            bilinear tap offsets/weights; results parked in shared memory.
//   phase B  8 lanes <-> one point, lane j <-> channels 4j..4j+3: each bilinear tap is ONE 128-byte line of
//            the channels-last feature map (8 x LDG.128), records leave as coalesced float4 stores, the
//            cross-view weighted mean/variance is accumulated in registers over the view loop.
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"

#define K1_THREADS 256
#define K1_TILE_P 32
#define K1_PAIR_F 36          // floats of dir-feature per pair in smem: [img 32 | rgb 3 | pad]

struct K1PairInfo {           // 64 bytes, written in phase A, read (broadcast) in phase B
    int   fo[4];              // feature-map tap offsets (floats) within the view's [fh,fw,32] map
    float fw_[4];             // feature tap weights * mask   (nw, ne, sw, se)
    int   io[4];              // image tap offsets (pixels) within one H*W plane
    float iw[4];              // image tap weights * mask
};

__global__ void __launch_bounds__(K1_THREADS, 3)
gn_k1_kernel(const __grid_constant__ GnK1Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int V = p.V;
    const int npair = K1_TILE_P * V;
    K1PairInfo* s_info = reinterpret_cast<K1PairInfo*>(smem_raw);                    // [npair]
    float* s_dfeat = reinterpret_cast<float*>(s_info + npair);                       // [npair][36]
    float* s_f     = s_dfeat + npair * K1_PAIR_F;                                    // [npair][36] stash of f for the variance pass
    float* s_misc  = s_f + npair * K1_PAIR_F;                                        // [npair][8]: mask, depth, dd0..3, pad

    const int tiles_per_scene = p.tiles_per_scene;
    const int b = blockIdx.x / tiles_per_scene;
    const int tile = blockIdx.x - b * tiles_per_scene;
    const int tid = threadIdx.x;

    // ---- tile -> point mapping -------------------------------------------------------------------
    // volume mode: tile = 2x2x8 block of voxels (i,j,k); record index n = (i*R+j)*R + (R-1-k)
    //              (renderer.py:169-170: reshape (1,R*R,R,3) then flip the sample axis)
    // ray mode   : tile = 32 consecutive points n of the explicit pts array
    const int R = p.R;
    int tk = 0, tj = 0, ti = 0;
    if (p.volume_mode) {
        const int nz = R >> 3, ny = R >> 1;
        tk = tile % nz; tj = (tile / nz) % ny; ti = tile / (nz * ny);
    }

    // =============================== phase A ======================================================
    for (int pair = tid; pair < npair; pair += K1_THREADS) {
        const int pl = pair / V;
        const int v = pair - pl * V;
        float px, py, pz, qx = 0.f, qy = 0.f, qz = 1.f;   // que_dir = (0,0,1) in volume mode, renderer.py:179
        bool live = true;
        if (p.volume_mode) {
            const int i = ti * 2 + (pl >> 4), j = tj * 2 + ((pl >> 3) & 1), k = tk * 8 + (pl & 7);
            // field_utils.py:17-27 table (host-built, fp32) + bbox3d[0] in fp32 (renderer.py:167-168)
            px = __fadd_rn(__ldg(p.axis + i), __ldg(p.bbox_min + b * 3 + 0));
            py = __fadd_rn(__ldg(p.axis + j), __ldg(p.bbox_min + b * 3 + 1));
            pz = __fadd_rn(__ldg(p.axis + k), __ldg(p.bbox_min + b * 3 + 2));
        } else {
            int n = tile * K1_TILE_P + pl;
            live = n < p.N;
            n = min(n, p.N - 1);
            const float* q = p.pts + ((size_t)b * p.N + n) * 3;
            px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
            const float* d = p.que_dir + ((size_t)b * (p.N / p.dn) + n / p.dn) * 3;
            qx = __ldg(d); qy = __ldg(d + 1); qz = __ldg(d + 2);
        }
        const float* Hm = p.KRt + ((size_t)b * V + v) * 12;
        // render_ops.py:94-99, fixed order ((h0*x + h1*y) + h2*z) + h3, no FMA (index-table parity)
        float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(Hm + 0), px), __fmul_rn(__ldg(Hm + 1), py)), __fmul_rn(__ldg(Hm + 2), pz)), __ldg(Hm + 3));
        float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(Hm + 4), px), __fmul_rn(__ldg(Hm + 5), py)), __fmul_rn(__ldg(Hm + 6), pz)), __ldg(Hm + 7));
        float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(Hm + 8), px), __fmul_rn(__ldg(Hm + 9), py)), __fmul_rn(__ldg(Hm + 10), pz)), __ldg(Hm + 11));
        const bool near_zero = fabsf(zc) < 1e-4f;                 // render_ops.py:101
        const float depth = near_zero ? 1e-3f : zc;               // render_ops.py:102
        const float u = __fdiv_rn(xc, depth), w_ = __fdiv_rn(yc, depth);   // render_ops.py:103
        const bool outside = (u < -0.5f) | (u >= (float)p.W - 0.5f) | (w_ < -0.5f) | (w_ >= (float)p.H - 0.5f);
        const float mask = (live && !near_zero && !outside) ? 1.f : 0.f;   // render_ops.py:126-128 (no z>0 test)

        // view direction, render_ops.py:112-114
        const float* cc = p.cam + ((size_t)b * V + v) * 3;
        const float dx = __fsub_rn(px, __ldg(cc)), dy = __fsub_rn(py, __ldg(cc + 1)), dz = __fsub_rn(pz, __ldg(cc + 2));
        const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-5f);
        const float ex = __fdiv_rn(-dx, nrm), ey = __fdiv_rn(-dy, nrm), ez = __fdiv_rn(-dz, nrm);
        // aggregate_net.py:13-15
        float dd[4] = { ex - qx, ey - qy, ez - qz, (ex * qx + ey * qy) + ez * qz };

        // ray_dir_fc: Linear(4,16) ELU Linear(16,35) ELU   (ibrnet.py:382-385,457)
        float hid[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a = p.rdfc.b0[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) a = fmaf(p.rdfc.w0[k][i], dd[i], a);
            hid[k] = gn_elu(a);
        }
        float* df = s_dfeat + pair * K1_PAIR_F;
#pragma unroll
        for (int c = 0; c < 35; ++c) {          // output order permuted on the host: [img 32 | rgb 3]
            float a = p.rdfc.b1[c];
#pragma unroll
            for (int k = 0; k < 16; ++k) a = fmaf(p.rdfc.w1[c][k], hid[k], a);
            df[c] = gn_elu(a);
        }
        df[35] = 0.f;

        // bilinear taps.  feature maps: align_corners=False (map size != image size), images: True
        // (render_ops.py:64-68); both normalised by the IMAGE size (ops.py:29-30); border padding.
        K1PairInfo inf;
        {
            const bool ac = (p.fh == p.H) && (p.fw == p.W);
            GnTap1D tx = gn_tap1d(u, p.W, p.fw, ac), ty = gn_tap1d(w_, p.H, p.fh, ac);
            inf.fo[0] = (ty.i0 * p.fw + tx.i0) * GN_FEAT_C; inf.fo[1] = (ty.i0 * p.fw + tx.i1) * GN_FEAT_C;
            inf.fo[2] = (ty.i1 * p.fw + tx.i0) * GN_FEAT_C; inf.fo[3] = (ty.i1 * p.fw + tx.i1) * GN_FEAT_C;
            inf.fw_[0] = __fmul_rn(tx.w0, ty.w0) * mask; inf.fw_[1] = __fmul_rn(tx.w1, ty.w0) * mask;
            inf.fw_[2] = __fmul_rn(tx.w0, ty.w1) * mask; inf.fw_[3] = __fmul_rn(tx.w1, ty.w1) * mask;
            if (p.dbg_feat_idx && live) {   // optional index-table dump for the bit-exactness tests
                int n_dbg;
                if (p.volume_mode) {
                    const int i = ti * 2 + (pl >> 4), j = tj * 2 + ((pl >> 3) & 1), k = tk * 8 + (pl & 7);
                    n_dbg = (i * R + j) * R + (R - 1 - k);
                } else n_dbg = tile * K1_TILE_P + pl;
                int* o = p.dbg_feat_idx + (((size_t)b * p.N + n_dbg) * V + v) * 2;
                o[0] = tx.i0; o[1] = ty.i0;
            }
        }
        {
            GnTap1D tx = gn_tap1d(u, p.W, p.W, true), ty = gn_tap1d(w_, p.H, p.H, true);
            inf.io[0] = ty.i0 * p.W + tx.i0; inf.io[1] = ty.i0 * p.W + tx.i1;
            inf.io[2] = ty.i1 * p.W + tx.i0; inf.io[3] = ty.i1 * p.W + tx.i1;
            inf.iw[0] = __fmul_rn(tx.w0, ty.w0) * mask; inf.iw[1] = __fmul_rn(tx.w1, ty.w0) * mask;
            inf.iw[2] = __fmul_rn(tx.w0, ty.w1) * mask; inf.iw[3] = __fmul_rn(tx.w1, ty.w1) * mask;
        }
        s_info[pair] = inf;
        float* ms = s_misc + pair * 8;
        ms[0] = mask; ms[1] = depth; ms[2] = dd[0]; ms[3] = dd[1]; ms[4] = dd[2]; ms[5] = dd[3];
    }
    __syncthreads();

    // =============================== phase B ======================================================
    const int lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, j = lane & 7;
    const int pl = warp * 4 + grp;                 // local point 0..31
    const unsigned gbase = lane & ~7u;
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
    for (int v = 0; v < V; ++v) nvalid += s_misc[(pl * V + v) * 8];
    const float wden = nvalid + 1e-8f;             // ibrnet.py:466
    const int S = p.S;
    float* rec = p.rec + ((size_t)b * p.N + n) * V * S;
    const size_t fmap_sz = (size_t)p.fh * p.fw * GN_FEAT_C;
    const size_t plane = (size_t)p.H * p.W;

    float4 m_img = make_float4(0.f, 0.f, 0.f, 0.f);
    float m_rgb = 0.f;
    for (int v = 0; v < V; ++v) {
        const int pair = pl * V + v;
        const int4 fo = *reinterpret_cast<const int4*>(s_info[pair].fo);
        const float4 fwt = *reinterpret_cast<const float4*>(s_info[pair].fw_);
        const float* rf = p.ray_feats + ((size_t)b * V + v) * fmap_sz + 4 * j;
        const float* imf = p.img_feats + ((size_t)b * V + v) * fmap_sz + 4 * j;
        // issue all eight 128-bit gathers before use
        const float4 r0 = ldg4(rf + fo.x), r1 = ldg4(rf + fo.y), r2 = ldg4(rf + fo.z), r3 = ldg4(rf + fo.w);
        const float4 g0 = ldg4(imf + fo.x), g1 = ldg4(imf + fo.y), g2 = ldg4(imf + fo.z), g3 = ldg4(imf + fo.w);
        float rgbv = 0.f;
        if (j < 3) {
            const int4 io = *reinterpret_cast<const int4*>(s_info[pair].io);
            const float4 iw = *reinterpret_cast<const float4*>(s_info[pair].iw);
            const float* im = p.imgs + (((size_t)b * V + v) * 3 + j) * plane;
            rgbv = __ldg(im + io.x) * iw.x;
            rgbv = fmaf(__ldg(im + io.y), iw.y, rgbv);
            rgbv = fmaf(__ldg(im + io.z), iw.z, rgbv);
            rgbv = fmaf(__ldg(im + io.w), iw.w, rgbv);
        }
        const float mask = s_misc[pair * 8], depth = s_misc[pair * 8 + 1];
        const float wv = __fdiv_rn(mask, wden);
        float4 ray = f4_mul(r0, fwt.x); ray = f4_fma(r1, fwt.y, ray); ray = f4_fma(r2, fwt.z, ray); ray = f4_fma(r3, fwt.w, ray);
        float4 img = f4_mul(g0, fwt.x); img = f4_fma(g1, fwt.y, img); img = f4_fma(g2, fwt.z, img); img = f4_fma(g3, fwt.w, img);
        const float4 dfe = *reinterpret_cast<const float4*>(s_dfeat + pair * K1_PAIR_F + 4 * j);
        const float4 fimg = f4_add(img, dfe);                                  // ibrnet.py:459
        const float frgb = rgbv + s_dfeat[pair * K1_PAIR_F + 32 + (j < 3 ? j : 3)];
        float* row = rec + (size_t)v * S;
        if (live) {
            st4_cs(row + GN_REC_RAYF + 4 * j, ray);
            st4_cs(row + GN_REC_FIMG + 4 * j, fimg);
        }
        st4(s_f + pair * K1_PAIR_F + 4 * j, fimg);
        if (j < 3) s_f[pair * K1_PAIR_F + 32 + j] = frgb;
        m_img = f4_fma(fimg, wv, m_img);
        m_rgb = fmaf(frgb, wv, m_rgb);
        // tail chunks: lane 0 <- (frgb0,frgb1,frgb2,mask), lane 1 <- (depth, rgb0, rgb1, rgb2)
        const float f1 = __shfl_sync(0xffffffffu, frgb, gbase + 1), f2 = __shfl_sync(0xffffffffu, frgb, gbase + 2);
        const float c0 = __shfl_sync(0xffffffffu, rgbv, gbase + 0), c2 = __shfl_sync(0xffffffffu, rgbv, gbase + 2);
        if (live) {
            if (j == 0) st4_cs(row + GN_REC_FRGB, make_float4(frgb, f1, f2, mask));
            else if (j == 1) st4_cs(row + GN_REC_DEPTH, make_float4(depth, c0, rgbv, c2));
            else if (j == 2 && S > GN_REC_DD) st4_cs(row + GN_REC_DD, *reinterpret_cast<const float4*>(s_misc + pair * 8 + 2));
        }
    }
    __syncwarp();
    // variance pass over the stashed f (ibrnet.py:115: sum_v w * (x - mean)^2)
    float4 v_img = make_float4(0.f, 0.f, 0.f, 0.f);
    float v_rgb = 0.f;
    for (int v = 0; v < V; ++v) {
        const int pair = pl * V + v;
        const float wv = __fdiv_rn(s_misc[pair * 8], wden);
        const float4 f = *reinterpret_cast<const float4*>(s_f + pair * K1_PAIR_F + 4 * j);
        float d;
        d = f.x - m_img.x; v_img.x = fmaf(wv * d, d, v_img.x);
        d = f.y - m_img.y; v_img.y = fmaf(wv * d, d, v_img.y);
        d = f.z - m_img.z; v_img.z = fmaf(wv * d, d, v_img.z);
        d = f.w - m_img.w; v_img.w = fmaf(wv * d, d, v_img.w);
        if (j < 3) { d = s_f[pair * K1_PAIR_F + 32 + j] - m_rgb; v_rgb = fmaf(wv * d, d, v_rgb); }
    }
    const float mr1 = __shfl_sync(0xffffffffu, m_rgb, gbase + 1), mr2 = __shfl_sync(0xffffffffu, m_rgb, gbase + 2);
    const float vr1 = __shfl_sync(0xffffffffu, v_rgb, gbase + 1), vr2 = __shfl_sync(0xffffffffu, v_rgb, gbase + 2);
    if (live) {
        float* pt = p.pt + ((size_t)b * p.N + n) * GN_PT_STRIDE;
        st4_cs(pt + 4 * j, m_img);
        st4_cs(pt + 36 + 4 * j, v_img);
        if (j == 0) {
            st4_cs(pt + 32, make_float4(m_rgb, mr1, mr2, nvalid));
            st4_cs(pt + 68, make_float4(v_rgb, vr1, vr2, 0.f));
        }
    }
}

extern "C" int gn_k1_forward(const GnK1Params* hp, void* stream)
{
    GnK1Params p = *hp;
    if (p.V < 1 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (p.S != GN_REC_VOL && p.S != GN_REC_RAY) return -2;
    if (p.volume_mode) {
        if ((p.R % 8) != 0 || p.N != p.R * p.R * p.R || !p.axis || !p.bbox_min) return -3;
        p.tiles_per_scene = (p.R / 2) * (p.R / 2) * (p.R / 8);
    } else {
        if (!p.pts || !p.que_dir || p.dn < 1 || (p.N % p.dn) != 0) return -4;
        p.tiles_per_scene = (p.N + K1_TILE_P - 1) / K1_TILE_P;
    }
    const int npair = K1_TILE_P * p.V;
    const size_t smem = (size_t)npair * (sizeof(K1PairInfo) + (2 * K1_PAIR_F + 8) * sizeof(float));
    if (smem > 227 * 1024) return -5;
    cudaError_t e = cudaFuncSetAttribute(gn_k1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long grid = (long long)p.B * p.tiles_per_scene;
    if (grid > 0x7fffffffLL) return -6;
    gn_k1_kernel<<<(unsigned)grid, K1_THREADS, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
