// Helpers shared by the backward (training) kernels: ELU derivative, transposed mat-vec, and the CTA-wide
// weight-gradient accumulation  dW[k][n] += sum_rows x[row][k] * dz[row][n]  staged through shared memory.
#pragma once
#include "gn_common.cuh"

#define GN_BWD_LDX 33          // floats per row of the X staging tile (<= 32 inputs per chunk, +1 against bank conflicts)
#define GN_BWD_LDZ 68          // floats per row of the dZ staging tile (<= 64 outputs; 16-byte aligned rows, 4-bank skew)

// ---- packed fp32x2 arithmetic (Blackwell FFMA2: two IEEE fp32 FMAs per issued instruction, bit-identical to scalar fmaf)
typedef unsigned long long gn_f2;
__device__ __forceinline__ gn_f2 gn_pk2(float lo, float hi) { gn_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void gn_upk2(gn_f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ gn_f2 gn_fma2(gn_f2 a, gn_f2 b, gn_f2 c) { gn_f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// derivative of ELU expressed with its OUTPUT y = elu(u):  u > 0 ? 1 : exp(u) = y + 1
__device__ __forceinline__ float gn_delu(float y) { return y > 0.f ? 1.f : y + 1.f; }

// dx[k] (+)= sum_{n<N} W[k*NP + n] * dz[n]      (W k-major in shared memory; the transpose of mv_acc)
// Fully unrolled form (small kernels only: the code grows with K*N).
template <int K, int N, int NP, bool ACC>
__device__ __forceinline__ void mv_bwd(const float* __restrict__ W, const float* dz, float* dx)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float a = ACC ? dx[k] : 0.f;
#pragma unroll
        for (int n = 0; n < N; ++n) a = fmaf(W[k * NP + n], dz[n], a);
        dx[k] = a;
    }
}
// Compact forms for the big per-row kernels: the k loop stays ROLLED (straight-line code for every layer would be megabytes of
// SASS and the kernel becomes instruction-fetch bound), the dynamically indexed operand goes through the thread's private
// column of a shared-memory scratch area: scr[k * SCR_STRIDE] (stride = threads per CTA, conflict free).
template <int K, int NP, int SCR_STRIDE>
__device__ __forceinline__ void mv_acc_rolled(const float* __restrict__ W, const float* x, float* y, float* scr)
{
#pragma unroll
    for (int k = 0; k < K; ++k) scr[k * SCR_STRIDE] = x[k];
    gn_f2 acc[NP / 2];
#pragma unroll
    for (int n = 0; n < NP; n += 2) acc[n / 2] = gn_pk2(y[n], y[n + 1]);
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float xk = scr[k * SCR_STRIDE];
        const gn_f2 xx = gn_pk2(xk, xk);
#pragma unroll
        for (int n = 0; n < NP; n += 4) {
            const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(W + k * NP + n);      // (w[n],w[n+1]) , (w[n+2],w[n+3])
            acc[n / 2] = gn_fma2(w.x, xx, acc[n / 2]);
            acc[n / 2 + 1] = gn_fma2(w.y, xx, acc[n / 2 + 1]);
        }
    }
#pragma unroll
    for (int n = 0; n < NP; n += 2) gn_upk2(acc[n / 2], y[n], y[n + 1]);
}
template <int K, int NP, bool ACC, int SCR_STRIDE>       // dz has NP entries (pad entries zero)
__device__ __forceinline__ void mv_bwd_rolled(const float* __restrict__ W, const float* dz, float* dx, float* scr)
{
    gn_f2 dzp[NP / 2];
#pragma unroll
    for (int n = 0; n < NP; n += 2) dzp[n / 2] = gn_pk2(dz[n], dz[n + 1]);
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        gn_f2 a = 0ull, b = 0ull;                        // +0.0f, +0.0f
#pragma unroll
        for (int n = 0; n < NP; n += 4) {
            const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(W + k * NP + n);
            a = gn_fma2(w.x, dzp[n / 2], a);
            b = gn_fma2(w.y, dzp[n / 2 + 1], b);
        }
        float a0, a1, b0, b1;
        gn_upk2(a, a0, a1); gn_upk2(b, b0, b1);
        scr[k * SCR_STRIDE] = (a0 + b0) + (a1 + b1);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) dx[k] = ACC ? dx[k] + scr[k * SCR_STRIDE] : scr[k * SCR_STRIDE];
}

// gW[k*NP + n] += sum_{r<nrows} sX[r][k] * sZ[r][n]   for k < KC, n < N.  N % 4 == 0: one (k, 4 n) strip per thread and pass
// (1 LDS.32 + 1 LDS.128 + 4 FMA per row); otherwise one (k,n) per thread.  Coalesced fp64 atomics.
static __device__ __noinline__ void gn_dw_flush(double* __restrict__ gW, int NP, int KC, int N, const float* __restrict__ sX,
                                         const float* __restrict__ sZ, int nrows, int nthreads)
{
    // small layers: also split the rows, so that every thread of the CTA has a strip to reduce (a few more atomics)
    const bool vec = (N & 3) == 0;
    const int items = vec ? KC * (N >> 2) : KC * N;
    int nsplit = 1;
    if (items * 8 <= nthreads) { nsplit = nthreads / (items * 4); if (nsplit > (nrows >> 1)) nsplit = nrows >> 1; }
    const int rpp = (((nrows + nsplit - 1) / nsplit) + 1) & ~1;              // rows per part, even
    if (vec) {
        const int n4 = N >> 2;
        for (int idx = threadIdx.x; idx < items * nsplit; idx += nthreads) {
            const int part = idx / items, rem = idx - part * items;
            const int k = rem / n4, n = (rem - k * n4) * 4;
            const int r0 = part * rpp, r1 = min(nrows, r0 + rpp);
            gn_f2 alo = 0ull, ahi = 0ull, blo = 0ull, bhi = 0ull;
            for (int r = r0; r < r1; r += 2) {
                const float x0 = sX[r * GN_BWD_LDX + k], x1 = sX[(r + 1) * GN_BWD_LDX + k];
                const ulonglong2 z0 = *reinterpret_cast<const ulonglong2*>(sZ + r * GN_BWD_LDZ + n);
                const ulonglong2 z1 = *reinterpret_cast<const ulonglong2*>(sZ + (r + 1) * GN_BWD_LDZ + n);
                const gn_f2 xx0 = gn_pk2(x0, x0), xx1 = gn_pk2(x1, x1);
                alo = gn_fma2(z0.x, xx0, alo); ahi = gn_fma2(z0.y, xx0, ahi);
                blo = gn_fma2(z1.x, xx1, blo); bhi = gn_fma2(z1.y, xx1, bhi);
            }
            float4 a, b;
            gn_upk2(alo, a.x, a.y); gn_upk2(ahi, a.z, a.w); gn_upk2(blo, b.x, b.y); gn_upk2(bhi, b.z, b.w);
            double* o = gW + k * NP + n;
            atomicAdd(o, (double)a.x + (double)b.x); atomicAdd(o + 1, (double)a.y + (double)b.y);
            atomicAdd(o + 2, (double)a.z + (double)b.z); atomicAdd(o + 3, (double)a.w + (double)b.w);
        }
        return;
    }
    for (int idx = threadIdx.x; idx < items * nsplit; idx += nthreads) {
        const int part = idx / items, rem = idx - part * items;
        const int k = rem / N, n = rem - k * N;
        const int r0 = part * rpp, r1 = min(nrows, r0 + rpp);
        float a0 = 0.f, a1 = 0.f;
        for (int r = r0; r < r1; r += 2) {
            a0 = fmaf(sX[r * GN_BWD_LDX + k], sZ[r * GN_BWD_LDZ + n], a0);
            a1 = fmaf(sX[(r + 1) * GN_BWD_LDX + k], sZ[(r + 1) * GN_BWD_LDZ + n], a1);
        }
        atomicAdd(gW + k * NP + n, (double)a0 + (double)a1);
    }
}
static __device__ __noinline__ void gn_db_flush(double* __restrict__ gB, int N, const float* __restrict__ sZ, int nrows, int nthreads)
{
    for (int n = threadIdx.x; n < N; n += nthreads) {
        float a = 0.f;
        for (int r = 0; r < nrows; ++r) a += sZ[r * GN_BWD_LDZ + n];
        atomicAdd(gB + n, (double)a);
    }
}
// Every thread of the CTA contributes its row (x[K], dz[N]); nthreads (even) rows are reduced.  gB may be NULL.
// Contains __syncthreads(): must be reached by all threads of the CTA.
template <int K, int N, int NP>
__device__ __forceinline__ void dw_layer(double* gW, double* gB, const float* x, const float* dz, float* sX, float* sZ, int nthreads)
{
    static_assert(N <= 64, "dZ staging tile holds 64 outputs");
    const int t = threadIdx.x;
    __syncthreads();                                     // the staging tiles may alias per-thread scratch columns still being read
#pragma unroll
    for (int n = 0; n < N; ++n) sZ[t * GN_BWD_LDZ + n] = dz[n];
#pragma unroll
    for (int kc = 0; kc < K; kc += 32) {
        const int kn = (K - kc) < 32 ? (K - kc) : 32;
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (kc + k < K) sX[t * GN_BWD_LDX + k] = x[kc + k];
        __syncthreads();
        gn_dw_flush(gW + kc * NP, NP, kn, N, sX, sZ, nthreads, nthreads);
        if (kc == 0 && gB) gn_db_flush(gB, N, sZ, nthreads, nthreads);
        __syncthreads();
    }
}
