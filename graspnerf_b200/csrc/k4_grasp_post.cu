// K4: grasp post-processing on the device - the step right after the path in the planner (GraspNeRFPlanner.__call__,
// main.py:188-209): `process` (main.py:23-55) and `select` (main.py:58-74) on the 40^3 volumes, so that a grasp attempt
// never leaves the GPU between image upload and grasp list (the reference copies four volumes to the host and runs
// scipy.ndimage there).  Latency-bound: 64 000 voxels; four tiny launches.
//
//   process : qual = gaussian_filter(qual, sigma, mode='nearest')          scipy: separable, truncate 4 -> radius int(4 sigma + .5),
//                                                                          float64 accumulation per axis pass, fp32 between passes
//             valid = binary_dilation(tsdf > hi, iterations=2, mask=~(lo < tsdf < hi))   6-neighbour cross, border 0, masked-off
//                                                                          voxels keep their value at every iteration
//             qual[~valid] = 0 ; qual[width < min | width > max] = 0
//   select  : qual[qual < thr] = 0 ; max = maximum_filter(qual, size=s)    window [i - s/2, i + s - s/2 - 1], mode 'reflect'
//             keep qual == max (and != 0) ; np.argwhere order (i, j, k lexicographic)
// Arithmetic follows scipy's order (ni_filters.c NI_Correlate1D, symmetric branch) so the filtered volume is bit-identical.
#include "gn_common.cuh"
#include "../../include/graspnerf_b200.h"

#define K4_MAXRAD 16

struct K4Gauss { double w[2 * K4_MAXRAD + 1]; int rad; };

// one axis pass of gaussian_filter1d(mode='nearest'): out[l] = c*w[0] + sum_{j=1..rad} (in[l-j] + in[l+j]) * w[j], far taps first
template <int AXIS>
__global__ void gn_k4_gauss_kernel(const float* __restrict__ in, float* __restrict__ out, int R, K4Gauss g)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = R * R * R;
    if (idx >= n) return;
    const int k = idx % R, j = (idx / R) % R, i = idx / (R * R);
    const int stride = AXIS == 0 ? R * R : (AXIS == 1 ? R : 1);
    const int l = AXIS == 0 ? i : (AXIS == 1 ? j : k);
    const float* line = in + (idx - l * stride);
    double tmp = (double)line[l * stride] * g.w[g.rad];
    for (int jj = -g.rad; jj < 0; ++jj) {
        const int a = min(max(l + jj, 0), R - 1), b = min(max(l - jj, 0), R - 1);
        tmp += ((double)line[a * stride] + (double)line[b * stride]) * g.w[jj + g.rad];
    }
    out[idx] = (float)tmp;
}

__device__ __forceinline__ bool k4_x0(const float* tsdf, int R, int i, int j, int k, float hi) {
    if ((unsigned)i >= (unsigned)R || (unsigned)j >= (unsigned)R || (unsigned)k >= (unsigned)R) return false;   // border_value 0
    return tsdf[(i * R + j) * R + k] > hi;
}
__device__ __forceinline__ bool k4_band(const float* tsdf, int R, int i, int j, int k, float lo, float hi) {
    const float t = tsdf[(i * R + j) * R + k];
    return !(lo < t && t < hi);                                                   // mask = logical_not(inside_voxels)
}
__device__ __forceinline__ bool k4_x1(const float* tsdf, int R, int i, int j, int k, float lo, float hi) {
    if ((unsigned)i >= (unsigned)R || (unsigned)j >= (unsigned)R || (unsigned)k >= (unsigned)R) return false;
    const bool self = k4_x0(tsdf, R, i, j, k, hi);
    if (!k4_band(tsdf, R, i, j, k, lo, hi)) return self;
    return self | k4_x0(tsdf, R, i - 1, j, k, hi) | k4_x0(tsdf, R, i + 1, j, k, hi) | k4_x0(tsdf, R, i, j - 1, k, hi)
                | k4_x0(tsdf, R, i, j + 1, k, hi) | k4_x0(tsdf, R, i, j, k - 1, hi) | k4_x0(tsdf, R, i, j, k + 1, hi);
}

// process(): band mask (two masked dilations evaluated directly from the TSDF), width limits; then select()'s threshold
__global__ void gn_k4_mask_kernel(const GnGraspPostParams p, const float* __restrict__ qual_f, float* __restrict__ qual_thr)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = p.R;
    if (idx >= R * R * R) return;
    const int k = idx % R, j = (idx / R) % R, i = idx / (R * R);
    const float lo = p.tsdf_thres_low, hi = p.tsdf_thres_high;
    bool valid = k4_x1(p.tsdf, R, i, j, k, lo, hi);
    if (k4_band(p.tsdf, R, i, j, k, lo, hi))
        valid = valid | k4_x1(p.tsdf, R, i - 1, j, k, lo, hi) | k4_x1(p.tsdf, R, i + 1, j, k, lo, hi) | k4_x1(p.tsdf, R, i, j - 1, k, lo, hi)
                      | k4_x1(p.tsdf, R, i, j + 1, k, lo, hi) | k4_x1(p.tsdf, R, i, j, k - 1, lo, hi) | k4_x1(p.tsdf, R, i, j, k + 1, lo, hi);
    float q = qual_f[idx];
    if (!valid) q = 0.f;
    const float w = p.width[idx];
    if (w < p.min_width || w > p.max_width) q = 0.f;
    p.qual_out[idx] = q;                                   // what process() returns
    qual_thr[idx] = q < p.threshold ? 0.f : q;             // select(): qual_vol[qual_vol < threshold] = 0
}

// select(): maximum_filter(size, mode='reflect') + non-maximum suppression flag
__global__ void gn_k4_nms_kernel(const float* __restrict__ q, float* __restrict__ keep, int R, int size)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * R * R) return;
    const int k = idx % R, j = (idx / R) % R, i = idx / (R * R);
    const int lo = -(size / 2), hi = size - size / 2 - 1;
    float m = -INFINITY;
    for (int a = lo; a <= hi; ++a) {
        int ii = i + a; ii = ii < 0 ? -ii - 1 : (ii >= R ? 2 * R - 1 - ii : ii);          // reflect: d c b a | a b c d | d c b a
        for (int b = lo; b <= hi; ++b) {
            int jj = j + b; jj = jj < 0 ? -jj - 1 : (jj >= R ? 2 * R - 1 - jj : jj);
            for (int c = lo; c <= hi; ++c) {
                int kk = k + c; kk = kk < 0 ? -kk - 1 : (kk >= R ? 2 * R - 1 - kk : kk);
                m = fmaxf(m, q[(ii * R + jj) * R + kk]);
            }
        }
    }
    const float v = q[idx];
    keep[idx] = (v == m && v != 0.f) ? v : 0.f;            // np.where(qual == max, qual, 0); mask = qual != 0
}

// np.argwhere order: one CTA, thread t owns a contiguous chunk, exclusive scan of the chunk counts, ordered writes
#define K4_CT 1024
__global__ void __launch_bounds__(K4_CT, 1)
gn_k4_compact_kernel(const GnGraspPostParams p, const float* __restrict__ keep)
{
    __shared__ int s_cnt[K4_CT];
    const int R = p.R, n = R * R * R;
    const int per = (n + K4_CT - 1) / K4_CT;
    const int t = threadIdx.x;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    int c = 0;
    for (int x = lo; x < hi; ++x) c += keep[x] != 0.f;
    s_cnt[t] = c;
    __syncthreads();
    for (int off = 1; off < K4_CT; off <<= 1) {            // Hillis-Steele inclusive scan
        const int v = t >= off ? s_cnt[t - off] : 0;
        __syncthreads();
        s_cnt[t] += v;
        __syncthreads();
    }
    int pos = s_cnt[t] - c;
    if (t == K4_CT - 1) *p.count = s_cnt[t];
    for (int x = lo; x < hi; ++x) {
        const float v = keep[x];
        if (v != 0.f) {
            if (pos < p.max_grasps) {
                float* g = p.grasps + (size_t)pos * 9;
                const int k = x % R, j = (x / R) % R, i = x / (R * R);
                g[0] = (float)i; g[1] = (float)j; g[2] = (float)k; g[3] = v;                           // select_index (main.py:77-84)
                g[4] = p.rot[x]; g[5] = p.rot[n + x]; g[6] = p.rot[2 * n + x]; g[7] = p.rot[3 * n + x];
                g[8] = p.width[x];
            }
            ++pos;
        }
    }
}

extern "C" int gn_k4_grasp_post(const GnGraspPostParams* hp, void* stream)
{
    const GnGraspPostParams& p = *hp;
    if (p.R < 1 || p.R > 256 || p.max_filter_size < 1 || p.max_filter_size > 15 || p.max_grasps < 0) return -1;
    if (!p.tsdf || !p.qual || !p.rot || !p.width || !p.qual_out || !p.scratch || !p.count || (p.max_grasps > 0 && !p.grasps)) return -2;
    if (!(p.sigma > 0.f)) return -3;
    K4Gauss g;
    g.rad = (int)(4.0 * (double)p.sigma + 0.5);            // scipy: lw = int(truncate * sd + 0.5), truncate = 4.0
    if (g.rad > K4_MAXRAD) return -3;
    {
        // scipy.ndimage._filters._gaussian_kernel1d: exp(-0.5 / sigma^2 * x^2) / sum, float64
        const double s2 = (double)p.sigma * (double)p.sigma;
        double sum = 0.0;
        for (int x = -g.rad; x <= g.rad; ++x) { g.w[x + g.rad] = exp(-0.5 / s2 * (double)(x * x)); sum += g.w[x + g.rad]; }
        for (int x = 0; x <= 2 * g.rad; ++x) g.w[x] /= sum;
    }
    const int n = p.R * p.R * p.R;
    const int threads = 256, blocks = (n + threads - 1) / threads;
    cudaStream_t st = (cudaStream_t)stream;
    float* s0 = p.scratch; float* s1 = p.scratch + n; float* s2 = p.scratch + 2 * n;
    gn_k4_gauss_kernel<0><<<blocks, threads, 0, st>>>(p.qual, s0, p.R, g);
    gn_k4_gauss_kernel<1><<<blocks, threads, 0, st>>>(s0, s1, p.R, g);
    gn_k4_gauss_kernel<2><<<blocks, threads, 0, st>>>(s1, s0, p.R, g);
    gn_k4_mask_kernel<<<blocks, threads, 0, st>>>(p, s0, s1);
    gn_k4_nms_kernel<<<blocks, threads, 0, st>>>(s1, s2, p.R, p.max_filter_size);
    gn_k4_compact_kernel<<<1, K4_CT, 0, st>>>(p, s2);
    return (int)cudaGetLastError();
}
