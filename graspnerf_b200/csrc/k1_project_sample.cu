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
// Work decomposition (one CTA = 256 threads = one tile of 32 points x V views):
//   phase A  thread <-> (point, view): projection, mask, view direction, dir_diff, bilinear tap offsets/weights;
//            results parked in shared memory.
//   phase B  8 lanes <-> one point, lane j <-> channels 4j..4j+3: each bilinear tap is ONE 128-byte line of
//            the channels-last feature map (8 x LDG.128), the 4 image taps are 4 RGBA texels on lanes 0..3, records leave
//            as coalesced float4 streaming stores.
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"
#include <cstdlib>

#define K1_THREADS 256
#define K1_TILE_P 32

struct K1PairInfo {           // written in phase A, read (broadcast within an 8-lane group) in phase B.  80-byte stride: the four
                              // groups of a warp read pairs V apart; 64 B (and 32 B for s_misc) strides put them on the same banks
    int   fo[4];              // feature-map tap offsets (floats) within the view's [fh,fw,32] map
    float fw_[4];             // feature tap weights * mask   (nw, ne, sw, se)
    int   io[4];              // image tap offsets (pixels) within one H*W plane
    float iw[4];              // image tap weights * mask
    float pad[4];
};
#define K1_MISC 12            // floats per pair in s_misc: dd0..3, mask, depth, pad

template <int MINB>
__global__ void __launch_bounds__(K1_THREADS, MINB)
gn_k1_kernel(const __grid_constant__ GnK1Params p, const int nsub)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int V = p.V;
    const int npair = K1_TILE_P * V;
    K1PairInfo* s_info = reinterpret_cast<K1PairInfo*>(smem_raw);                    // [npair]
    float* s_misc  = reinterpret_cast<float*>(s_info + npair);                       // [npair][12]: dd0..3, mask, depth, pad

    const int tiles_per_scene = p.tiles_per_scene;          // CTAs per scene
    const int b = blockIdx.x / tiles_per_scene;
    const int cta_tile = blockIdx.x - b * tiles_per_scene;
    const int tid = threadIdx.x;
    const int R = p.R;

    // ---- tile -> point mapping -------------------------------------------------------------------
    // volume mode: tile = 2x2x8 block of voxels (i,j,k); record index n = (i*R+j)*R + (R-1-k)
    //              (renderer.py:169-170: reshape (1,R*R,R,3) then flip the sample axis).  With nsub = 4 a CTA walks the
    //              four x/y-adjacent tiles of a 4x4x8 block one after the other, so their shared taps hit in L1.
    // ray mode   : tile = 32 consecutive points n of the explicit pts array
    for (int sub = 0; sub < nsub; ++sub) {
    int tile = cta_tile;
    int tk = 0, tj = 0, ti = 0;
    if (p.volume_mode) {
        const int nz = R >> 3;
        if (nsub == 4) {
            const int ny2 = R >> 2;
            tk = cta_tile % nz; tj = ((cta_tile / nz) % ny2) * 2 + (sub & 1); ti = (cta_tile / (nz * ny2)) * 2 + (sub >> 1);
        } else {
            const int ny = R >> 1;
            tk = tile % nz; tj = (tile / nz) % ny; ti = tile / (nz * ny);
        }
    }
    if (sub) __syncthreads();                                // shared-memory records of the previous sub-tile are done

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
        *reinterpret_cast<int4*>(s_info[pair].fo) = *reinterpret_cast<const int4*>(inf.fo);
        *reinterpret_cast<float4*>(s_info[pair].fw_) = *reinterpret_cast<const float4*>(inf.fw_);
        *reinterpret_cast<int4*>(s_info[pair].io) = *reinterpret_cast<const int4*>(inf.io);
        *reinterpret_cast<float4*>(s_info[pair].iw) = *reinterpret_cast<const float4*>(inf.iw);
        float* ms = s_misc + pair * K1_MISC;
        st4(ms, make_float4(dd[0], dd[1], dd[2], dd[3]));
        ms[4] = mask; ms[5] = depth;
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
    unsigned bits = 0u;
    for (int v = 0; v < V; ++v) {
        const float m = s_misc[(pl * V + v) * K1_MISC + 4];
        nvalid += m;
        bits |= (m != 0.f ? 1u : 0u) << v;
    }
    float* rec = p.rec + ((size_t)b * p.N + n) * V * GN_REC_STRIDE;
    const size_t fmap_sz = (size_t)p.fh * p.fw * GN_FEAT_C;
    const size_t plane = (size_t)p.H * p.W;
    const float* rf_base = p.ray_feats + (size_t)b * V * fmap_sz + 4 * j;
    const float* if_base = p.img_feats + (size_t)b * V * fmap_sz + 4 * j;
    const float* im_base = p.imgs + (size_t)b * V * plane * 4;            // RGBA-interleaved [B,V,H,W,4]

    for (int v = 0; v < V; ++v) {
        const int pair = pl * V + v;
        const int4 fo = *reinterpret_cast<const int4*>(s_info[pair].fo);
        const float4 fwt = *reinterpret_cast<const float4*>(s_info[pair].fw_);
        const float* rf = rf_base + (size_t)v * fmap_sz;
        const float* imf = if_base + (size_t)v * fmap_sz;
        // issue all eight 128-bit gathers before use
        const float4 r0 = ldg4(rf + fo.x), r1 = ldg4(rf + fo.y), r2 = ldg4(rf + fo.z), r3 = ldg4(rf + fo.w);
        const float4 g0 = ldg4(imf + fo.x), g1 = ldg4(imf + fo.y), g2 = ldg4(imf + fo.z), g3 = ldg4(imf + fo.w);
        // image taps: lane j<4 fetches tap j as one RGBA texel (LDG.128) and scales it; the 4 partial colours are
        // summed over lanes 0..3 of the group with two xor-shuffles (order (t0+t1)+(t2+t3))
        float cr = 0.f, cg = 0.f, cb = 0.f;
        if (j < 4) {
            const int io = s_info[pair].io[j];
            const float iw = s_info[pair].iw[j];
            const float4 px = ldg4(im_base + ((size_t)v * plane + io) * 4);
            cr = px.x * iw; cg = px.y * iw; cb = px.z * iw;
        }
        float4 ray = f4_mul(r0, fwt.x); ray = f4_fma(r1, fwt.y, ray); ray = f4_fma(r2, fwt.z, ray); ray = f4_fma(r3, fwt.w, ray);
        float4 img = f4_mul(g0, fwt.x); img = f4_fma(g1, fwt.y, img); img = f4_fma(g2, fwt.z, img); img = f4_fma(g3, fwt.w, img);
        float* row = rec + (size_t)v * GN_REC_STRIDE;
        // tail chunks: lane 0 <- (rgb0, rgb1, rgb2, depth), lane 1 <- dir_diff
        cr += __shfl_xor_sync(0xffffffffu, cr, 1); cg += __shfl_xor_sync(0xffffffffu, cg, 1); cb += __shfl_xor_sync(0xffffffffu, cb, 1);
        cr += __shfl_xor_sync(0xffffffffu, cr, 2); cg += __shfl_xor_sync(0xffffffffu, cg, 2); cb += __shfl_xor_sync(0xffffffffu, cb, 2);
        if (live) {
            st4_cs(row + GN_REC_RAYF + 4 * j, ray);
            st4_cs(row + GN_REC_IMGF + 4 * j, img);
            // tail: lanes 0 and 1 write the two adjacent 16-byte chunks [64,68) and [68,72) with ONE store instruction
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
    }   // sub-tile loop
}

extern "C" int gn_k1_forward(const GnK1Params* hp, void* stream)
{
    GnK1Params p = *hp;
    if (p.V < 1 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (p.volume_mode) {
        if ((p.R % 8) != 0 || p.N != p.R * p.R * p.R || !p.axis || !p.bbox_min) return -3;
        p.tiles_per_scene = (p.R / 2) * (p.R / 2) * (p.R / 8);
    } else {
        if (!p.pts || !p.que_dir || p.dn < 1 || (p.N % p.dn) != 0) return -4;
        p.tiles_per_scene = (p.N + K1_TILE_P - 1) / K1_TILE_P;
    }
    const int npair = K1_TILE_P * p.V;
    const size_t smem = (size_t)npair * (sizeof(K1PairInfo) + K1_MISC * sizeof(float));
    if (smem > 227 * 1024) return -5;
    // resident CTAs per SM the kernel is compiled for (register budget): 4 by default; GN_K1_MINB=4|5|6 selects another
    // instantiation (tuning aid)
    static int nsub_env = -1;
    if (nsub_env < 0) { const char* e = getenv("GN_K1_SUPER"); nsub_env = (e && atoi(e) == 1) ? 4 : 1; }
    int nsub = 1;
    if (p.volume_mode && nsub_env == 4 && (p.R % 4) == 0) { nsub = 4; p.tiles_per_scene /= 4; }
    static int minb = 0;
    if (!minb) { const char* e = getenv("GN_K1_MINB"); minb = e ? atoi(e) : 4; if (minb < 4 || minb > 6) minb = 4; }
    const long long grid = (long long)p.B * p.tiles_per_scene;
    if (grid > 0x7fffffffLL) return -6;
    cudaError_t e;
    static size_t c4[16] = {0}, c5[16] = {0}, c6[16] = {0};
    if (minb == 5) {
        e = gn_ensure_smem(gn_k1_kernel<5>, smem, c5); if (e != cudaSuccess) return (int)e;
        gn_k1_kernel<5><<<(unsigned)grid, K1_THREADS, smem, (cudaStream_t)stream>>>(p, nsub);
    } else if (minb == 6) {
        e = gn_ensure_smem(gn_k1_kernel<6>, smem, c6); if (e != cudaSuccess) return (int)e;
        gn_k1_kernel<6><<<(unsigned)grid, K1_THREADS, smem, (cudaStream_t)stream>>>(p, nsub);
    } else {
        e = gn_ensure_smem(gn_k1_kernel<4>, smem, c4); if (e != cudaSuccess) return (int)e;
        gn_k1_kernel<4><<<(unsigned)grid, K1_THREADS, smem, (cudaStream_t)stream>>>(p, nsub);
    }
    return (int)cudaGetLastError();
}
