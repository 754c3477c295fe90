// K7: 2-D convolution of the encoders as an implicit GEMM on tcgen05 / TMEM (src/nr/network/ops.py:78-230: every
// nn.Conv2d of ResUNetLight / BasicBlock / conv / upconv; init_net.py:21-26; vis_encoder.py:9-14), inference only.
//
// cuDNN runs these fp32 convolutions as SIMT `implicit_convolve_sgemm` at ~21 TFLOP/s (3.1 of the encoders' 4.1 ms per
// 6x288x512 scene, profiles/profile_forward_r02h.txt).  Here:  out[m][n] = sum_k A[m][k] * W[n][k]
//   m = output pixel (image, oy, ox), 128 per CTA = the 128 TMEM lanes;   n = output channel (16..128);
//   k = (ci, dy, dx), consumed in chunks of 32.
// A is gathered from the NCHW fp32 input - ALREADY reflection-padded by the producing K6 launch, so the convolution is a
// plain "valid" one and a pixel's k-th input is in[base(m) + koff[k]] with a per-layer offset table - split into fp16 hi / lo
// halves exactly like K2a's operands (4 instructions per pair) and written to shared memory in the tcgen05 K-major
// SWIZZLE_NONE layout (element (m,k) at 16-byte unit (k/8)*128 + m; layout validated by tools/tc_probe.cu, SS form).  The
// weights are prepared once per weight update (fp16 hi / lo images per k-chunk) and arrive by cp.async.bulk.  Three MMAs per
// product (lo*hi, hi*lo, hi*hi, fp32 accumulation in TMEM) keep fp32 accuracy.  A ring of stages decouples the gather
// (512 threads = 16 gather warps, which signal a per-stage `ready` mbarrier) from the MMAs and the weight TMA (a 17th, issuer
// warp; completion tracked by tcgen05.commit on the stage's `empty` mbarrier): no block-wide barrier inside the chunk loop - with
// thread 0 issuing between two __syncthreads, 33 % of the stall samples were the other 15 warps waiting for it (r02zc capture).  Epilogue: warp w reads TMEM lanes 32*(w%4).., column chunks w/4, w/4+4, ..
#include "k2a_tc_common.cuh"

#define K7_THREADS 512                 // gather threads, 4 per output pixel: thread (m, q) gathers k = 8q..8q+7 of every chunk (one 16-byte operand unit)
#define K7_BLOCK (K7_THREADS + 32)     // + one issuer warp: weight TMA + tcgen05.mma + commits (lane 0); the gather warps never wait for it
#define K7_KC 32                       // k per chunk (two MMA k-steps of 16)
#ifndef K7_STAGES
#define K7_STAGES 3
#endif
#define K7_A_BYTES (128 * K7_KC * 2)   // one half (hi or lo) of the A chunk: 8 KB

__host__ __device__ constexpr size_t k7_stage_bytes(int N) { return (size_t)2 * K7_A_BYTES + (size_t)2 * N * K7_KC * 2; }
__host__ __device__ constexpr int k7_tmem_cols(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 128); }     // power of two >= 32

__device__ __forceinline__ void k7_mma_ss(uint32_t d_tmem, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t idesc, uint32_t acc) {
    // both descriptors: SBO = 128 B (>>4 = 8), version 1 (bit 46) in the high word
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "mov.b64 ad, {%1, %5};\n\tmov.b64 bd, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(adesc_lo), "r"(bdesc_lo), "r"(idesc), "r"(acc), "n"(0x4008) : "memory");
}

template <int N>
__global__ void __launch_bounds__(K7_BLOCK)
gn_k7_conv_kernel(const GnConvParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr size_t STAGE = k7_stage_bytes(N);
    int* s_koff = reinterpret_cast<int*>(smem + K7_STAGES * STAGE);                    // [Kpad]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_koff + p.Kpad);                      // full[S] | empty[S] | done | ready[S]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 3 * K7_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < p.Kpad; i += K7_BLOCK) s_koff[i] = __ldg(p.koff + i);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(k7_tmem_cols(N)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < 2 * K7_STAGES + 1; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[s])), "r"(1));
        for (int s = 0; s < K7_STAGES; ++s)              // ready[s]: one arrival per gather warp
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[2 * K7_STAGES + 1 + s])), "r"(K7_THREADS / 32));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;

    // this thread's output pixel m and its slice q of every k chunk
    const int m = tid & 127, q = tid >> 7;
    const long long gm = (long long)blockIdx.x * 128 + m;
    const bool live = gm < p.M;
    const long long gmc = live ? gm : p.M - 1;
    const int hw = p.Ho * p.Wo;
    const int img = p.M <= 0x7fffffffLL ? (int)gmc / hw : (int)(gmc / hw);          // (a 64-bit division is ~100 instructions per thread)
    const int r = (int)(gmc - (long long)img * hw);
    const int oy = r / p.Wo, ox = r - oy * p.Wo;
    const float* base = p.in + ((size_t)img * p.Cin * p.Hp + (size_t)oy * p.stride) * p.Wp + (size_t)ox * p.stride;

    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F32 accum, F16 x F16, M = 128
    // split-K: this CTA runs chunks [c_lo, c_hi) of the layer's k chunks and writes partial output `split`
    const int nchunk_all = p.Kpad / K7_KC;
    const int nsplit = p.ksplit > 1 ? p.ksplit : 1, split = (int)blockIdx.y;
    const int c_lo = nchunk_all * split / nsplit, c_hi = nchunk_all * (split + 1) / nsplit;     // nchunk_all * 8 fits an int by far
    const int nchunk = c_hi - c_lo;
    const unsigned char* wimg = reinterpret_cast<const unsigned char*>(p.wimg) + (size_t)c_lo * (2 * N * K7_KC * 2);
    constexpr uint32_t B_BYTES = (uint32_t)(2 * N * K7_KC * 2);                          // hi + lo of one chunk

    // ---- issuer warp: weight chunks by TMA (two chunks ahead of the MMAs), MMAs as soon as a stage's A image is ready
    if (warp == K7_THREADS / 32) {
        if ((tid & 31) == 0) {
            auto load_weights = [&](int k) {              // chunk k -> the B half of stage k % STAGES (its previous MMAs must be done)
                const int s1 = k % K7_STAGES;
                const uint32_t fb = smem_u32(&s_bar[s1]);
                if (k >= K7_STAGES) mbar_wait(smem_u32(&s_bar[K7_STAGES + s1]), (uint32_t)((k / K7_STAGES - 1) & 1));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(fb), "r"(B_BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(smem + s1 * STAGE + 2 * K7_A_BYTES)), "l"(wimg + (size_t)k * B_BYTES), "r"(B_BYTES), "r"(fb) : "memory");
            };
            for (int k = 0; k < K7_STAGES - 1 && k < nchunk; ++k) load_weights(k);
            for (int c = 0; c < nchunk; ++c) {
                const int st = c % K7_STAGES;
                if (c + K7_STAGES - 1 < nchunk) load_weights(c + K7_STAGES - 1);         // into the stage chunk c-1 has just been issued from
                unsigned char* sA = smem + st * STAGE;                                   // A hi | A lo | B hi | B lo
                unsigned char* sB = sA + 2 * K7_A_BYTES;
                mbar_wait(smem_u32(&s_bar[2 * K7_STAGES + 1 + st]), (uint32_t)((c / K7_STAGES) & 1));     // the 16 gather warps have written A
                mbar_wait(smem_u32(&s_bar[st]), (uint32_t)((c / K7_STAGES) & 1));                          // the weight chunk has landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(sA) >> 4, a_lo = a_hi + (K7_A_BYTES >> 4);
                const uint32_t b_hi = smem_u32(sB) >> 4, b_lo = b_hi + ((N * K7_KC * 2) >> 4);
                constexpr uint32_t lboA = (uint32_t)((128 * 16) >> 4) << 16, lboB = (uint32_t)((N * 16) >> 4) << 16;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {                                   // small terms first: lo*hi, hi*lo, hi*hi
                    const uint32_t ao = pass == 0 ? a_lo : a_hi, bo = pass == 1 ? b_lo : b_hi;
#pragma unroll
                    for (int ks = 0; ks < K7_KC / 16; ++ks)
                        k7_mma_ss(tmem, ((ao + (uint32_t)(ks * 2 * 128)) & 0x3FFF) | lboA, ((bo + (uint32_t)(ks * 2 * N)) & 0x3FFF) | lboB, idesc,
                                  (c > 0 || pass > 0 || ks > 0) ? 1u : 0u);
                }
                // completion of everything issued so far -> this stage may be overwritten; after the last chunk -> accumulator ready
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&s_bar[K7_STAGES + st])) : "memory");
                if (c == nchunk - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&s_bar[2 * K7_STAGES])) : "memory");
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                                 // (the gather warps' final barrier, below)
        return;
    }
    // ---- gather warps: chunk after chunk into the ring; they only wait for a stage's previous MMAs (empty), never for the issuer
    for (int c = 0; c < nchunk; ++c) {
        const int st = c % K7_STAGES;
        unsigned char* sA = smem + st * STAGE;
        if (c >= K7_STAGES) mbar_wait(smem_u32(&s_bar[K7_STAGES + st]), (uint32_t)((c / K7_STAGES - 1) & 1));   // the MMAs that read this stage have completed
        // gather + split this thread's 8 k values of the chunk -> one 16-byte unit of the hi image and one of the lo image
        const int* ko = s_koff + (c_lo + c) * K7_KC;
        {
            const int g8 = q;
            float a[8];
            const int4 k0 = *reinterpret_cast<const int4*>(ko + g8 * 8), k1 = *reinterpret_cast<const int4*>(ko + g8 * 8 + 4);
            a[0] = __ldg(base + k0.x); a[1] = __ldg(base + k0.y); a[2] = __ldg(base + k0.z); a[3] = __ldg(base + k0.w);
            a[4] = __ldg(base + k1.x); a[5] = __ldg(base + k1.y); a[6] = __ldg(base + k1.z); a[7] = __ldg(base + k1.w);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                asm("{\n\t.reg .b16 h0, h1, m1;\n\t.reg .f32 d0, d1;\n\t"
                    "cvt.rn.f16x2.f32 %0, %3, %2;\n\t"
                    "mov.b32 {h0, h1}, %0;\n\tmov.b16 m1, 0xBC00;\n\t"
                    "fma.rn.f32.f16 d0, h0, m1, %2;\n\tfma.rn.f32.f16 d1, h1, m1, %3;\n\t"
                    "cvt.rn.f16x2.f32 %1, d1, d0;\n\t}" : "=&r"(hi[i]), "=r"(lo[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
            }
            uint4* dh = reinterpret_cast<uint4*>(sA + ((size_t)g8 * 128 + m) * 16);
            uint4* dl = reinterpret_cast<uint4*>(sA + K7_A_BYTES + ((size_t)g8 * 128 + m) * 16);
            *dh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *dl = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                    // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if ((tid & 31) == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&s_bar[2 * K7_STAGES + 1 + st])) : "memory");
    }
    mbar_wait(smem_u32(&s_bar[2 * K7_STAGES]), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: row m of D -> out[img][co][oy][ox] (+ bias); for a fixed co the 32 lanes of a warp write 32 consecutive pixels
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* op = p.out + (size_t)split * ((size_t)p.Nimg * p.Cout * p.Ho * p.Wo) + ((size_t)img * p.Cout * p.Ho + oy) * p.Wo + ox;
    const size_t cstride = (size_t)p.Ho * p.Wo;
    const bool add_bias = p.bias != nullptr && split == 0;
    const bool full_n = p.Cout == N;                       // no padded output channels: no per-element channel test
#pragma unroll
    for (int c0 = 16 * q; c0 < N; c0 += 64) {
        float y[16];
        tm_ld<16>(lane_addr + c0, y);
        if (add_bias) {                                    // CTA-uniform branch; the bias has Cout entries (not Npad): the tail is guarded
#pragma unroll
            for (int i = 0; i < 16; ++i) if (full_n || c0 + i < p.Cout) y[i] += __ldg(p.bias + c0 + i);
        }
        if (live) {
            float* o = op + (size_t)c0 * cstride;
            if (full_n) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { *o = y[i]; o += cstride; }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) { if (c0 + i < p.Cout) *o = y[i]; o += cstride; }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(k7_tmem_cols(N)));
}

template <int N>
static cudaError_t k7_launch(const GnConvParams& p, cudaStream_t st)
{
    const size_t smem = K7_STAGES * k7_stage_bytes(N) + (size_t)p.Kpad * 4 + (3 * K7_STAGES + 1) * 8 + 16;
    static size_t cache[16] = {0};
    cudaError_t e = gn_ensure_smem(gn_k7_conv_kernel<N>, smem, cache);
    if (e != cudaSuccess) return e;
    const long long tiles = (p.M + 127) / 128;
    gn_k7_conv_kernel<N><<<dim3((unsigned)tiles, (unsigned)(p.ksplit > 1 ? p.ksplit : 1), 1), K7_BLOCK, smem, st>>>(p);
    return cudaGetLastError();
}

extern "C" int gn_k7_conv_forward(const GnConvParams* hp, void* stream)
{
    GnConvParams p = *hp;
    if (p.Nimg < 1 || p.Cin < 1 || p.Cout < 1 || p.Ho < 1 || p.Wo < 1 || p.stride < 1 || p.Kpad < K7_KC || (p.Kpad % K7_KC) != 0) return -1;
    if (!p.in || !p.wimg || !p.koff || !p.out) return -2;
    p.M = (long long)p.Nimg * p.Ho * p.Wo;
    if (p.ksplit < 0 || p.ksplit > 8 || (p.ksplit > 1 && p.Kpad / K7_KC < p.ksplit)) return -5;
    if ((p.M + 127) / 128 > 0x7fffffffLL) return -6;
    const int npad = (p.Cout + 15) / 16 * 16;
    if (npad != p.Npad) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (p.Npad) {
        case 16: e = k7_launch<16>(p, st); break;
        case 32: e = k7_launch<32>(p, st); break;
        case 48: e = k7_launch<48>(p, st); break;
        case 64: e = k7_launch<64>(p, st); break;
        case 128: e = k7_launch<128>(p, st); break;
        default: return -4;
    }
    return (int)e;
}
