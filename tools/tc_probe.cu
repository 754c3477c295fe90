// tcgen05 probe (development aid, not part of the product): validates on real sm_100a hardware the operand-layout
// assumptions K2a's tensor-core path relies on:
//   * B operand in shared memory, K-major, SWIZZLE_NONE: element (n,k) at (k/8)*LBO + (n/8)*SBO + (n%8)*16 + (k%8)*2 bytes
//   * A operand in tensor memory (TS form): row m in lane m, fp16 pairs packed two per 32-bit column
//   * accumulator D (M=128, cta_group::1): row m in lane m, column n
//   * fp16 hi/lo split, 3 MMAs: error vs fp64
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu ; run: ./tc_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

template <int N> __device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void tmem_ld32<32>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

// variant bit0: fp16 pair packing order in TMEM (0: low half = even k) ; bit1: swap LBO/SBO ; bit2: SS mode (A from smem)
// mode 0: exact test (inputs representable in fp16) ; mode 1: fp16 hi/lo split, 3 MMAs
template <int N, int K>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                    int variant, int mode)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __half* sBhi = reinterpret_cast<__half*>(smem);                 // [K/8][N][8]
    __half* sBlo = sBhi + N * K;
    __half* sAhi = sBlo + N * K;                                    // [K/8][128][8]  (SS mode)
    __half* sAlo = sAhi + 128 * K;
    const int t = threadIdx.x, warp = t >> 5;
    const bool swap = variant & 2, ss = variant & 4, odd_first = variant & 1;

    for (int i = t; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const float w = B[n * K + k];
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const int off = (k / 8) * (N * 8) + n * 8 + (k % 8);
        sBhi[off] = hi; sBlo[off] = lo;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t colD = 0, colAhi = 128, colAlo = 128 + K / 2;

    // ---- A operand: row t
    uint32_t ahi[K / 2], alo[K / 2];
#pragma unroll
    for (int c = 0; c < K / 2; ++c) {
        const float a0 = A[t * K + 2 * c], a1 = A[t * K + 2 * c + 1];
        const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
        const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
        const uint32_t uh0 = __half_as_ushort(h0), uh1 = __half_as_ushort(h1), ul0 = __half_as_ushort(l0), ul1 = __half_as_ushort(l1);
        ahi[c] = odd_first ? (uh1 | (uh0 << 16)) : (uh0 | (uh1 << 16));
        alo[c] = odd_first ? (ul1 | (ul0 << 16)) : (ul0 | (ul1 << 16));
        if (ss) {
            const int k0 = 2 * c;
            const int off = (k0 / 8) * (128 * 8) + t * 8 + (k0 % 8);
            sAhi[off] = h0; sAhi[off + 1] = h1; sAlo[off] = l0; sAlo[off + 1] = l1;
        }
    }
    if (!ss) {
#pragma unroll
        for (int c = 0; c < K / 2; c += 16) {
            tmem_st16(tb + lane_base + colAhi + c, ahi + c);
            tmem_st16(tb + lane_base + colAlo + c, alo + c);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t lboB = swap ? 128 : N * 16, sboB = swap ? N * 16 : 128;
        const uint32_t lboA = swap ? 128 : 128 * 16, sboA = swap ? 128 * 16 : 128;
        const int npass = mode == 0 ? 1 : 3;
        uint32_t acc = 0;
        for (int pass = 0; pass < npass; ++pass) {
            // pass 0: Ahi*Bhi ; pass 1: Alo*Bhi ; pass 2: Ahi*Blo
            const __half* sb = (pass == 2) ? sBlo : sBhi;
            const __half* sa = (pass == 1) ? sAlo : sAhi;
            const uint32_t colA = (pass == 1) ? colAlo : colAhi;
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * (N * 16), lboB, sboB);
                if (ss) {
                    const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * (128 * 16), lboA, sboA);
                    mma_ss(tb + colD, ad, bd, idesc, acc);
                } else {
                    mma_ts(tb + colD, tb + colA + ks * 8, bd, idesc, acc);
                }
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    mbar_wait(smem_u32(&mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32<32>(tb + lane_base + colD + c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int c = 0; c < 32; ++c) D[t * N + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(256));
}

template <int N, int K>
static void run(int variant, int mode)
{
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(1234 + N * 7 + K);
    for (auto& x : A) x = mode == 0 ? (float)((rand() % 33) - 16) / 8.0f : (float)rand() / RAND_MAX * 4.f - 2.f;
    for (auto& x : B) x = mode == 0 ? (float)((rand() % 33) - 16) / 16.0f : (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, D.size() * 4));
    const size_t smem = (size_t)(2 * N * K + 2 * 128 * K) * 2 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD, variant, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d variant=%d mode=%d: KERNEL ERROR %s\n", N, K, variant, mode, cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
            const double err = fabs(ref - (double)D[m * N + n]);
            if (!(err <= maxerr)) maxerr = err;
            if (fabs(ref) > maxref) maxref = fabs(ref);
        }
    printf("N=%3d K=%3d variant=%d (%s, pack %s, %s) mode=%d: max|err| = %.3e  (max|ref| %.3e)  %s\n", N, K, variant,
           (variant & 4) ? "SS" : "TS", (variant & 1) ? "odd-first" : "even-first", (variant & 2) ? "LBO/SBO swapped" : "LBO=k-chunk,SBO=8-row",
           mode, maxerr, maxref, maxerr < 1e-5 * (maxref + 1) ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main()
{
    for (int v = 0; v < 8; ++v) {
        if ((v & 4) && (v & 1)) continue;          // packing order is irrelevant in SS mode
        run<32, 32>(v, 0);
    }
    for (int v = 0; v < 8; ++v) {
        if ((v & 4) && (v & 1)) continue;
        run<96, 32>(v, 0);
        run<64, 64>(v, 0);
        run<16, 32>(v, 0);
    }
    for (int v = 0; v < 8; ++v) {
        if ((v & 4) && (v & 1)) continue;
        run<32, 32>(v, 1);
        run<64, 64>(v, 1);
    }
    return 0;
}
