// K5: the VGN 3-D ConvNet that consumes the TSDF volume on every call (src/gd/networks.py:39-97, renderer.py:323-330):
//   encoder  Conv3d(1,16,5,s2) - Conv3d(16,32,3,s2) - Conv3d(32,64,3,s2), ReLU after each        R^3 -> (R/8)^3 x 64
//   decoder  Conv3d(64,64,3) ReLU, nearest x2, Conv3d(64,32,3) ReLU, nearest x2, Conv3d(32,16,5) ReLU, nearest x2
//   heads    Conv3d(16,{1,4,1},5) on the upsampled grid: sigmoid / F.normalize / identity
// as seven direct fp32 convolutions (CUDA cores; cuDNN needs 2.7 ms for this 1.4 GMAC net at batch 1 on B200).
//
// Nearest-neighbour x2 upsampling followed by a KxK xK convolution is folded EXACTLY into 8 parity classes of 3x3x3
// convolutions on the low-resolution grid (the taps of a high-res output voxel o = 2m + parity that fall into the same
// low-res cell share their input value, so their weights are summed once, in fp64, when the weights are prepared): the three
// layers that run on upsampled grids need 27 instead of 125 / 27 taps per input channel and never materialise the upsampled
// tensors.  1.40 GMAC -> 0.34 GMAC, seven launches.  Skinny outputs (6 / 16 channels) do not fill an MMA tile; fp32 FFMA
// keeps the result within 1e-6 of the fp32 reference.
//
// thread <-> one output voxel x COUT_T output channels (register accumulators); weights of the CTA's parity class are staged
// in shared memory k-major ([cin*tap][cout]) and read as broadcast float4; the input is read through L1 (__ldg).
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"

#define K5_THREADS 128
#ifndef K5_SPLIT_D3
#define K5_SPLIT_D3 2          // lanes per output of the 20^3 -> 40^3 decoder layer / of the heads (A/B knobs, see KSPLIT below)
#endif
#ifndef K5_SPLIT_HEADS
#define K5_SPLIT_HEADS 1
#endif

// ACT: 0 none, 1 ReLU, 2 VGN heads (channel 0 sigmoid, 1..4 L2-normalised, 5 identity)
// KSPLIT: the input channels of one output are split over KSPLIT ADJACENT LANES (each runs CIN/KSPLIT channels, the partial sums
// meet in xor-shuffles, always in the same order).  The coarse layers have 125 / 1 000 output cells and a reduction of 864 / 1 728
// terms per output: with one thread per output that is a serial chain of ~1 700 dependent load+FMA steps on a handful of warps
// (45-65 us per layer, r02z launch list); split 8 ways the chain is ~200 steps and the layer has 8x the threads.
template <int CIN, int COUT_T, int KS, int STRIDE, bool FOLD, int ACT, int KSPLIT>
__global__ void __launch_bounds__(K5_THREADS)
gn_k5_conv_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias, float* __restrict__ out,
                  int Din, int Dout, int cout_total)
{
    constexpr int TAPS = FOLD ? 27 : KS * KS * KS;
    constexpr int PAD = FOLD ? 1 : KS / 2;
    constexpr int KD = FOLD ? 3 : KS;
    extern __shared__ __align__(16) float s_w[];                     // [CIN*TAPS][COUT_T]
    const int cls = FOLD ? (int)blockIdx.y : 0;
    const int cblk = (int)blockIdx.z;
    const int nblk = (int)gridDim.z;
    // stage this (class, cout block)'s weights: global layout [class][cout_block][CIN*TAPS][COUT_T]
    {
        const float* src = wgt + ((size_t)cls * nblk + cblk) * (size_t)(CIN * TAPS * COUT_T);
        if ((CIN * TAPS * COUT_T) % 4 == 0) {
            for (int i = threadIdx.x; i < CIN * TAPS * COUT_T / 4; i += K5_THREADS)
                reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        } else {
            for (int i = threadIdx.x; i < CIN * TAPS * COUT_T; i += K5_THREADS) s_w[i] = __ldg(src + i);
        }
    }
    __syncthreads();
    const int Dc = FOLD ? Din : Dout;                                // grid of "cells" the threads enumerate
    const int gt = blockIdx.x * K5_THREADS + threadIdx.x;
    const int ks = gt % KSPLIT;                                      // this lane's slice of the input channels
    const int ncell = Dc * Dc * Dc;
    // lanes past the last cell still take part in the shuffles (KSPLIT > 1): they compute on the last cell and do not store
    const bool live = gt / KSPLIT < ncell;
    if (KSPLIT == 1 && !live) return;
    const int cell = live ? gt / KSPLIT : ncell - 1;
    const int cz = cell % Dc, cy = (cell / Dc) % Dc, cx = cell / (Dc * Dc);
    // first input coordinate of the window
    const int bx = (FOLD ? cx : cx * STRIDE) - PAD, by = (FOLD ? cy : cy * STRIDE) - PAD, bz = (FOLD ? cz : cz * STRIDE) - PAD;
    float acc[COUT_T];
#pragma unroll
    for (int c = 0; c < COUT_T; ++c) acc[c] = ks == 0 ? __ldg(bias + cblk * COUT_T + c) : 0.f;
    const size_t plane = (size_t)Din * Din * Din;
    static_assert(CIN % KSPLIT == 0 && (KSPLIT & (KSPLIT - 1)) == 0 && KSPLIT <= 32, "KSPLIT: power of two dividing CIN");
#pragma unroll 1
    for (int ci = ks * (CIN / KSPLIT); ci < (ks + 1) * (CIN / KSPLIT); ++ci) {
        const float* ip = in + (size_t)ci * plane;
        const float* wp = s_w + (size_t)ci * TAPS * COUT_T;
#pragma unroll 1
        for (int dx = 0; dx < KD; ++dx) {
            const int x = bx + dx;
            if ((unsigned)x >= (unsigned)Din) continue;              // zero padding (warp-divergent only at the grid border)
#pragma unroll
            for (int dy = 0; dy < KD; ++dy) {
                const int y = by + dy;
                const bool yin = (unsigned)y < (unsigned)Din;
#pragma unroll
                for (int dz = 0; dz < KD; ++dz) {
                    const int z = bz + dz;
                    float v = 0.f;
                    if (yin && (unsigned)z < (unsigned)Din) v = __ldg(ip + ((size_t)x * Din + y) * Din + z);
                    const float* w = wp + ((dx * KD + dy) * KD + dz) * COUT_T;
                    if (COUT_T % 4 == 0) {
#pragma unroll
                        for (int c = 0; c < COUT_T; c += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(w + c);
                            acc[c] = fmaf(v, w4.x, acc[c]); acc[c + 1] = fmaf(v, w4.y, acc[c + 1]);
                            acc[c + 2] = fmaf(v, w4.z, acc[c + 2]); acc[c + 3] = fmaf(v, w4.w, acc[c + 3]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < COUT_T; ++c) acc[c] = fmaf(v, w[c], acc[c]);
                    }
                }
            }
        }
    }
    if (KSPLIT > 1) {
#pragma unroll
        for (int o = 1; o < KSPLIT; o <<= 1) {
#pragma unroll
            for (int c = 0; c < COUT_T; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        }
        if (ks != 0 || !live) return;
    }
    // output voxel
    int ox = cx, oy = cy, oz = cz;
    if (FOLD) { ox = 2 * cx + ((cls >> 2) & 1); oy = 2 * cy + ((cls >> 1) & 1); oz = 2 * cz + (cls & 1); }
    const size_t oplane = (size_t)Dout * Dout * Dout;
    const size_t o = ((size_t)ox * Dout + oy) * Dout + oz;
    if (ACT == 2) {
        // networks.py:49-53: sigmoid(qual), F.normalize(rot, dim=1) = x / max(||x||, 1e-12), width
        const float q = 1.f / (1.f + expf(-acc[0]));
        const float nrm = fmaxf(sqrtf(acc[1] * acc[1] + acc[2] * acc[2] + acc[3] * acc[3] + acc[4] * acc[4]), 1e-12f);
        out[o] = q;
#pragma unroll
        for (int c = 1; c < 5; ++c) out[(size_t)c * oplane + o] = acc[c] / nrm;
        out[5 * oplane + o] = acc[5];
    } else {
#pragma unroll
        for (int c = 0; c < COUT_T; ++c) {
            const int co = cblk * COUT_T + c;
            if (co < cout_total) out[(size_t)co * oplane + o] = ACT == 1 ? fmaxf(acc[c], 0.f) : acc[c];
        }
    }
}

template <int CIN, int COUT_T, int KS, int STRIDE, bool FOLD, int ACT, int KSPLIT = 1>
static cudaError_t k5_launch(const float* in, const float* w, const float* b, float* out, int Din, int Dout, int cout_total,
                             cudaStream_t st)
{
    constexpr int TAPS = FOLD ? 27 : KS * KS * KS;
    const size_t smem = (size_t)CIN * TAPS * COUT_T * sizeof(float);
    static size_t cache[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k5_conv_kernel<CIN, COUT_T, KS, STRIDE, FOLD, ACT, KSPLIT>, smem, cache);
    if (e != cudaSuccess) return e;
    const int Dc = FOLD ? Din : Dout;
    const int cells = Dc * Dc * Dc;
    dim3 grid((cells * KSPLIT + K5_THREADS - 1) / K5_THREADS, FOLD ? 8 : 1, (cout_total + COUT_T - 1) / COUT_T);
    gn_k5_conv_kernel<CIN, COUT_T, KS, STRIDE, FOLD, ACT, KSPLIT><<<grid, K5_THREADS, smem, st>>>(in, w, b, out, Din, Dout, cout_total);
    return cudaGetLastError();
}

// prepared-weight blob layout (floats), in layer order; per layer [classes][cout blocks][CIN*TAPS][COUT_T] then the biases
// (padded to the block size).  Sizes are exported so the host packer (weights.py:pack_vgn) cannot disagree.
struct K5Layer { int cin, cout, cout_t, taps, classes; };
// cout_t = output channels per thread.  The coarse layers have only 125 / 1 000 output voxels: there a thread owns ONE (or 4)
// output channel(s) and the grid's z dimension runs over the channels, otherwise 4 CTAs would do the whole layer (r02n: 117 us
// for the 5^3 decoder convolution with 16 channels per thread).
static const K5Layer kK5[7] = {
    {1, 16, 16, 125, 1}, {16, 32, 4, 27, 1}, {32, 64, 1, 27, 1}, {64, 64, 1, 27, 1},
    {64, 32, 1, 27, 8}, {32, 16, 4, 27, 8}, {16, 6, 8, 27, 8} };
static int k5_wfloats(int l) { const K5Layer& L = kK5[l]; const int nb = (L.cout + L.cout_t - 1) / L.cout_t; return L.classes * nb * L.cin * L.taps * L.cout_t; }
static int k5_bfloats(int l) { const K5Layer& L = kK5[l]; const int nb = (L.cout + L.cout_t - 1) / L.cout_t; return nb * L.cout_t; }

extern "C" int gn_vgn_layer_info(int layer, int* cin, int* cout, int* cout_t, int* taps, int* classes, int* w_offset, int* b_offset)
{
    if (layer < 0 || layer >= 7) return -1;
    int off = 0;
    for (int l = 0; l < layer; ++l) off += k5_wfloats(l) + k5_bfloats(l);
    const K5Layer& L = kK5[layer];
    if (cin) *cin = L.cin; if (cout) *cout = L.cout; if (cout_t) *cout_t = L.cout_t; if (taps) *taps = L.taps; if (classes) *classes = L.classes;
    if (w_offset) *w_offset = off;
    if (b_offset) *b_offset = off + k5_wfloats(layer);
    return 0;
}
extern "C" int gn_vgn_blob_floats(void) { int n = 0; for (int l = 0; l < 7; ++l) n += k5_wfloats(l) + k5_bfloats(l); return n; }
extern "C" int gn_vgn_workspace_floats(int R) { const int a = R / 2, b = R / 4, c = R / 8; return 16 * a * a * a + 32 * b * b * b + 2 * 64 * c * c * c + 32 * b * b * b + 16 * a * a * a; }

extern "C" int gn_vgn_forward(const GnVgnParams* hp, void* stream)
{
    const GnVgnParams& p = *hp;
    if (p.R < 8 || (p.R % 8) != 0 || p.B < 1) return -1;
    if (!p.volume || !p.weights || !p.workspace || !p.out) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = p.R, a = R / 2, b = R / 4, c = R / 8;
    int wo[7], bo[7];
    for (int l = 0; l < 7; ++l) gn_vgn_layer_info(l, 0, 0, 0, 0, 0, &wo[l], &bo[l]);
    const float* W = p.weights;
    cudaError_t e = cudaSuccess;
    for (int s = 0; s < p.B && e == cudaSuccess; ++s) {
        const float* vol = p.volume + (size_t)s * R * R * R;
        float* ws = p.workspace;                                   // one scene at a time reuses the workspace (stream-ordered)
        float* e1 = ws;                    ws += 16 * a * a * a;
        float* e2 = ws;                    ws += 32 * b * b * b;
        float* e3 = ws;                    ws += 64 * c * c * c;
        float* d1 = ws;                    ws += 64 * c * c * c;
        float* d2 = ws;                    ws += 32 * b * b * b;
        float* d3 = ws;
        float* out = p.out + (size_t)s * p.out_scene_stride;
        e = k5_launch<1, 16, 5, 2, false, 1>(vol, W + wo[0], W + bo[0], e1, R, a, 16, st);                 if (e) break;   // networks.py:66-67
        e = k5_launch<16, 4, 3, 2, false, 1, 4>(e1, W + wo[1], W + bo[1], e2, a, b, 32, st);               if (e) break;   // 69-70
        e = k5_launch<32, 1, 3, 2, false, 1, 8>(e2, W + wo[2], W + bo[2], e3, b, c, 64, st);               if (e) break;   // 72-73
        e = k5_launch<64, 1, 3, 1, false, 1, 8>(e3, W + wo[3], W + bo[3], d1, c, c, 64, st);               if (e) break;   // 85-86
        e = k5_launch<64, 1, 3, 1, true, 1, 8>(d1, W + wo[4], W + bo[4], d2, c, b, 32, st);                if (e) break;   // 88-90 (upsample folded)
        e = k5_launch<32, 4, 5, 1, true, 1, K5_SPLIT_D3>(d2, W + wo[5], W + bo[5], d3, b, a, 16, st);      if (e) break;   // 92-94
        e = k5_launch<16, 8, 5, 1, true, 2, K5_SPLIT_HEADS>(d3, W + wo[6], W + bo[6], out, a, R, 6, st);                   // 96 + 47-53
    }
    return (int)e;
}
