// K1 backward: scatter of the record gradient into the two feature maps.
//
// Reverse of the two feature gathers of K1 (render_ops.py:64-70 / renderer.py:80-88 through F.grid_sample, bilinear,
// border padding): rec.ray_feats = mask * sum_t w_t * ray_feats[tap_t]  =>  d ray_feats[tap_t] += mask * w_t * d rec.ray_feats,
// same for img_feats.  Tap indices / weights are recomputed with the forward's arithmetic (they carry no gradient: the
// query points and cameras are constants of the step).  The images and the geometric record entries (depth, dir_diff)
// need no gradient.
//
// CTA = 32 points x V views, as the forward: phase A thread <-> (point, view) recomputes offsets + weights into shared
// memory; phase B 8 lanes <-> one point, lane j <-> channels 4j..4j+3: one 128-bit vector atomic per tap and map.
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"

#define K1B_THREADS 256
#define K1B_TILE_P 32

struct K1BPair { int fo[4]; float fw_[4]; float pad[4]; };      // 48-byte stride (bank spread)

__global__ void __launch_bounds__(K1B_THREADS, 4)
gn_k1_backward_kernel(const __grid_constant__ GnK1BwdParams p, int tiles_per_scene)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K1BPair* s_info = reinterpret_cast<K1BPair*>(smem_raw);
    const int V = p.V;
    const int npair = K1B_TILE_P * V;
    const int b = blockIdx.x / tiles_per_scene;
    const int tile = blockIdx.x - b * tiles_per_scene;
    const int tid = threadIdx.x;
    const int R = p.R;
    int tk = 0, tj = 0, ti = 0;
    if (p.volume_mode) {
        const int nz = R >> 3, ny = R >> 1;
        tk = tile % nz; tj = (tile / nz) % ny; ti = tile / (nz * ny);
    }
    for (int pair = tid; pair < npair; pair += K1B_THREADS) {
        const int pl = pair / V;
        const int v = pair - pl * V;
        float px, py, pz;
        bool live = true;
        if (p.volume_mode) {
            const int i = ti * 2 + (pl >> 4), j = tj * 2 + ((pl >> 3) & 1), k = tk * 8 + (pl & 7);
            px = __fadd_rn(__ldg(p.axis + i), __ldg(p.bbox_min + b * 3 + 0));
            py = __fadd_rn(__ldg(p.axis + j), __ldg(p.bbox_min + b * 3 + 1));
            pz = __fadd_rn(__ldg(p.axis + k), __ldg(p.bbox_min + b * 3 + 2));
        } else {
            int n = tile * K1B_TILE_P + pl;
            live = n < p.N;
            n = min(n, p.N - 1);
            const float* q = p.pts + ((size_t)b * p.N + n) * 3;
            px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
        }
        const float* Hm = p.KRt + ((size_t)b * V + v) * 12;
        // identical op order to the forward (render_ops.py:94-103), no FMA
        const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(Hm + 0), px), __fmul_rn(__ldg(Hm + 1), py)), __fmul_rn(__ldg(Hm + 2), pz)), __ldg(Hm + 3));
        const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(Hm + 4), px), __fmul_rn(__ldg(Hm + 5), py)), __fmul_rn(__ldg(Hm + 6), pz)), __ldg(Hm + 7));
        const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(Hm + 8), px), __fmul_rn(__ldg(Hm + 9), py)), __fmul_rn(__ldg(Hm + 10), pz)), __ldg(Hm + 11));
        const bool near_zero = fabsf(zc) < 1e-4f;
        const float depth = near_zero ? 1e-3f : zc;
        const float u = __fdiv_rn(xc, depth), w_ = __fdiv_rn(yc, depth);
        const bool outside = (u < -0.5f) | (u >= (float)p.W - 0.5f) | (w_ < -0.5f) | (w_ >= (float)p.H - 0.5f);
        const float mask = (live && !near_zero && !outside) ? 1.f : 0.f;
        const bool ac = (p.fh == p.H) && (p.fw == p.W);
        const GnTap1D tx = gn_tap1d(u, p.W, p.fw, ac), ty = gn_tap1d(w_, p.H, p.fh, ac);
        K1BPair inf;
        inf.fo[0] = (ty.i0 * p.fw + tx.i0) * GN_FEAT_C; inf.fo[1] = (ty.i0 * p.fw + tx.i1) * GN_FEAT_C;
        inf.fo[2] = (ty.i1 * p.fw + tx.i0) * GN_FEAT_C; inf.fo[3] = (ty.i1 * p.fw + tx.i1) * GN_FEAT_C;
        inf.fw_[0] = __fmul_rn(tx.w0, ty.w0) * mask; inf.fw_[1] = __fmul_rn(tx.w1, ty.w0) * mask;
        inf.fw_[2] = __fmul_rn(tx.w0, ty.w1) * mask; inf.fw_[3] = __fmul_rn(tx.w1, ty.w1) * mask;
        *reinterpret_cast<int4*>(s_info[pair].fo) = *reinterpret_cast<const int4*>(inf.fo);
        *reinterpret_cast<float4*>(s_info[pair].fw_) = *reinterpret_cast<const float4*>(inf.fw_);
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, j = lane & 7;
    const int pl = warp * 4 + grp;
    int n;
    bool live = true;
    if (p.volume_mode) {
        const int i = ti * 2 + (pl >> 4), jj = tj * 2 + ((pl >> 3) & 1), k = tk * 8 + (pl & 7);
        n = (i * R + jj) * R + (R - 1 - k);
    } else {
        n = tile * K1B_TILE_P + pl;
        live = n < p.N;
        n = min(n, p.N - 1);
    }
    if (!live) return;
    const size_t fmap_sz = (size_t)p.fh * p.fw * GN_FEAT_C;
    const float* drec = p.d_rec + ((size_t)b * p.N + n) * V * 64;
    for (int v = 0; v < V; ++v) {
        const int pair = pl * V + v;
        const int4 fo = *reinterpret_cast<const int4*>(s_info[pair].fo);
        const float4 fwt = *reinterpret_cast<const float4*>(s_info[pair].fw_);
        if (fwt.x == 0.f && fwt.y == 0.f && fwt.z == 0.f && fwt.w == 0.f) continue;        // masked pair (or zero weights)
        const float4 dr = ldg4(drec + v * 64 + 4 * j);
        const float4 di = ldg4(drec + v * 64 + 32 + 4 * j);
        float* rf = p.d_ray_feats + ((size_t)b * V + v) * fmap_sz + 4 * j;
        float* imf = p.d_img_feats + ((size_t)b * V + v) * fmap_sz + 4 * j;
        atomicAdd(reinterpret_cast<float4*>(rf + fo.x), f4_mul(dr, fwt.x));
        atomicAdd(reinterpret_cast<float4*>(rf + fo.y), f4_mul(dr, fwt.y));
        atomicAdd(reinterpret_cast<float4*>(rf + fo.z), f4_mul(dr, fwt.z));
        atomicAdd(reinterpret_cast<float4*>(rf + fo.w), f4_mul(dr, fwt.w));
        atomicAdd(reinterpret_cast<float4*>(imf + fo.x), f4_mul(di, fwt.x));
        atomicAdd(reinterpret_cast<float4*>(imf + fo.y), f4_mul(di, fwt.y));
        atomicAdd(reinterpret_cast<float4*>(imf + fo.z), f4_mul(di, fwt.z));
        atomicAdd(reinterpret_cast<float4*>(imf + fo.w), f4_mul(di, fwt.w));
    }
}

extern "C" int gn_k1_backward(const GnK1BwdParams* hp, void* stream)
{
    const GnK1BwdParams& p = *hp;
    if (p.V < 1 || p.V > 32 || p.B < 1 || p.N < 1) return -1;
    if (!p.d_rec || !p.d_img_feats || !p.d_ray_feats || !p.KRt) return -2;
    int tiles;
    if (p.volume_mode) {
        if ((p.R % 8) != 0 || p.N != p.R * p.R * p.R || !p.axis || !p.bbox_min) return -3;
        tiles = (p.R / 2) * (p.R / 2) * (p.R / 8);
    } else {
        if (!p.pts) return -4;
        tiles = (p.N + K1B_TILE_P - 1) / K1B_TILE_P;
    }
    const size_t smem = (size_t)K1B_TILE_P * p.V * sizeof(K1BPair);
    if (smem > 227 * 1024) return -5;
    const long long grid = (long long)p.B * tiles;
    if (grid > 0x7fffffffLL) return -6;
    static size_t cache[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k1_backward_kernel, smem, cache);
    if (e != cudaSuccess) return (int)e;
    gn_k1_backward_kernel<<<(unsigned)grid, K1B_THREADS, smem, (cudaStream_t)stream>>>(p, tiles);
    return (int)cudaGetLastError();
}
