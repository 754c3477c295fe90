// Shared pieces of the tensor-core K2a kernel (k2a_head_tc3.cu) and its prepare step (k2a_tc_prepare.cu):
// TMEM column map, fp16 operand-image table, small-constant table, tcgen05 / mbarrier PTX wrappers, the prepare kernel.
#pragma once
#include "gn_common.cuh"
#include "gn_weights.cuh"
#include "../../include/graspnerf_b200.h"
#include <cuda_fp16.h>

// ---- TMEM column map per slot (256 columns each) -------------------------------------------------------------
#define TM_D 0            // accumulator, up to 112 columns
#define TM_AHI 112        // A operand hi halves, 72 columns (K <= 144)
#define TM_ALO 184        // A operand lo halves
#define TM_SLOT 256

// ---- shared-memory B images (fp16, element (n,k) at (k/8)*(N*8) + n*8 + k%8) ------------------------------------
struct TcLayer { int N, K; };
enum { L_DD1, L_DD2M, L_DD2V, L_DD2A, L_PE0, L_NFC, L_RD0, L_RD1, L_BF0C, L_BF0B, L_BF2, L_VF0, L_VF2, L_V20, L_GF0, L_GF2, L_COUNT };
__host__ __device__ constexpr TcLayer tc_layer(int i) {
    return i == L_DD1 ? TcLayer{96, 32} : i == L_DD2M ? TcLayer{32, 32} : i == L_DD2V ? TcLayer{32, 32} : i == L_DD2A ? TcLayer{32, 32}
         : i == L_PE0 ? TcLayer{32, 48} : i == L_NFC ? TcLayer{16, 32} : i == L_RD0 ? TcLayer{16, 16} : i == L_RD1 ? TcLayer{48, 16}
         : i == L_BF0C ? TcLayer{64, 80} : i == L_BF0B ? TcLayer{64, 144} : i == L_BF2 ? TcLayer{32, 64}
         : i == L_VF0 ? TcLayer{32, 32} : i == L_VF2 ? TcLayer{48, 32} : i == L_V20 ? TcLayer{32, 32}
         : i == L_GF0 ? TcLayer{64, 96} : TcLayer{16, 64};
}
__host__ __device__ constexpr int tc_img_off(int i) {          // offset in halves of the HI image; LO follows at +N*K
    int o = 0;
    for (int j = 0; j < i; ++j) o += 2 * tc_layer(j).N * tc_layer(j).K;
    return o;
}
constexpr int TC_IMG_HALVES = tc_img_off(L_COUNT);

// ---- small fp32 constants (biases + the CUDA-core layers), stored right after the images --------------------------
constexpr int kTcSmall[] = {
    GN_W_DD_MEAN_B0, GN_W_DD_VAR_B0, GN_W_DD_AW_B0, GN_W_DD_MEAN_B2, GN_W_DD_VAR_B2, GN_W_DD_AW_B2,
    GN_W_DD_MEAN_W4, GN_W_DD_VAR_W4, GN_W_DD_AW_W4, GN_W_DD_MEAN_B4, GN_W_DD_VAR_B4, GN_W_DD_AW_B4,
    GN_W_PE_B0, GN_W_NF_W2, GN_W_NF_B2, GN_W_RD_B0, GN_W_RD_B1, GN_W_BF_B2,
    GN_W_VF_B0, GN_W_VF_B2, GN_W_V2_B0, GN_W_V2_W2, GN_W_V2_B2,
    GN_W_RF_W0, GN_W_RF_B0, GN_W_RF_W2, GN_W_RF_B2, GN_W_RF_W4, GN_W_RF_B4, GN_W_GF_B0, GN_W_GF_B2,
    GN_W_NFC_B0, GN_W_BF_B0C };
constexpr int kTcSmallCount = sizeof(kTcSmall) / sizeof(int);
constexpr int ts_off_idx(int j) { int o = 0; for (int i = 0; i < j; ++i) o += gn_w_size(kTcSmall[i]); return o; }
constexpr int ts_find(int id) { for (int i = 0; i < kTcSmallCount; ++i) if (kTcSmall[i] == id) return i; return -1; }
template <int ID> struct TsOffT {
    static_assert(ts_find(ID) >= 0, "entry is not in the small-constant list");
    static constexpr int value = ts_off_idx(ts_find(ID));
};
constexpr int TC_SMALL_FLOATS = ts_off_idx(kTcSmallCount);
#define TS(id) (TsOffT<GN_W_##id>::value)
constexpr int TC_CONST_BYTES = TC_IMG_HALVES * 2 + TC_SMALL_FLOATS * 4;      // global "tc_const" buffer == its smem image
static_assert(TC_CONST_BYTES % 16 == 0, "tc_const must be copyable with 16-byte loads");
// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // tcgen05 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), SWIZZLE_NONE
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t idesc, uint32_t acc, uint32_t elected) {
    // executed by the whole issuing warp with warp-uniform operands; only the elected lane issues (CUTLASS idiom), so the
    // operands live in uniform registers.  High descriptor word is constant: SBO = 128 B (>>4 = 8), version 1 (bit 46).
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "mov.b64 bd, {%2, %6};\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "r"(bdesc_lo), "r"(idesc), "r"(acc), "r"(elected), "n"(0x4008) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar, uint32_t elected) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" :: "r"(bar), "r"(elected) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends the thread for a HW-bounded time per attempt; the attempt counter turns a lost arrival
    // (a bug) into a trap instead of a hung GPU.
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (it > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tm_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tm_ld16_fence(uint32_t* r) {
    // ties the loaded registers to a point AFTER tcgen05.wait::ld so no use can be scheduled above the wait
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                      "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
// N accumulator columns -> registers: all loads in flight, ONE wait
template <int N> __device__ __forceinline__ void tm_ld(uint32_t taddr, float* y) {
    uint32_t r[N];
#pragma unroll
    for (int c = 0; c < N; c += 16) tm_ld16_issue(taddr + c, r + c);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < N; c += 16) tm_ld16_fence(r + c);
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: two fp32 lanes per instruction) -------------------------
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// split K fp32 values into fp16 hi/lo pairs and store them as the A operand (k0 = first k index, multiple of 16)
template <int K, int AHI = TM_AHI, int ALO = TM_ALO> __device__ __forceinline__ void tm_store_a(uint32_t slot_lane_addr, int k0, const float* a) {
#pragma unroll
    for (int c = 0; c < K / 2; c += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a0 = a[2 * (c + i)], a1 = a[2 * (c + i) + 1];
#ifndef GN_SPLIT_CVT
            // hi = rn_f16(a) packed (low 16 bits = even k, tc_probe variant 0); lo = rn_f16(a - hi) with the residual taken by
            // the mixed-precision FMA (fma.rn.f32.f16: a - 1*hi in ONE instruction, `FHFMA` in SASS) - 4 instructions per pair
            asm("{\n\t.reg .b16 h0, h1, m1;\n\t.reg .f32 d0, d1;\n\t"
                "cvt.rn.f16x2.f32 %0, %3, %2;\n\t"
                "mov.b32 {h0, h1}, %0;\n\tmov.b16 m1, 0xBC00;\n\t"
                "fma.rn.f32.f16 d0, h0, m1, %2;\n\tfma.rn.f32.f16 d1, h1, m1, %3;\n\t"
                "cvt.rn.f16x2.f32 %1, d1, d0;\n\t}" : "=&r"(hi[i]), "=r"(lo[i]) : "f"(a0), "f"(a1));
#else
            const __half2 h = __floats2half2_rn(a0, a1);                 // .x (low 16 bits) = even k  (tc_probe variant 0)
            const float2 hf = __half22float2(h);
            float d0, d1;
            upk2(sub2(pk2(a0, a1), pk2(hf.x, hf.y)), d0, d1);            // one FADD2 for both residuals
            const __half2 l = __floats2half2_rn(d0, d1);
            hi[i] = *reinterpret_cast<const uint32_t*>(&h);
            lo[i] = *reinterpret_cast<const uint32_t*>(&l);
#endif
        }
        tm_st8(slot_lane_addr + AHI + k0 / 2 + c, hi);
        tm_st8(slot_lane_addr + ALO + k0 / 2 + c, lo);
    }
}
// sigmoid with two MUFUs (ex2, rcp): 1 / (1 + e^-x), __fdividef = rcp.approx * x (<= 2 ulp); the IEEE division of
// gn_sigmoid costs ~10 instructions and this kernel evaluates 8 sigmoids per row
__device__ __forceinline__ float tc_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// nn.Softplus(beta=1, threshold=20) = log1p(e^x) without the ~30-instruction log1pf: log(u) * e / (u - 1) with u = 1 + e
// cancels the rounding of 1 + e (error ~1e-7 relative for every e), two more MUFUs
__device__ __forceinline__ float tc_softplus(float x) {
    const float e = __expf(x), u = 1.f + e;
    const float r = (u == 1.f) ? e : __logf(u) * __fdividef(e, u - 1.f);
    return x > 20.f ? x : r;
}
// ELU with one MUFU: ex2.approx.ftz (rel. err 2^-22)
__device__ __forceinline__ float tc_elu(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.f;
}
// two ELUs: the scale and the "-1" are packed (FMUL2 / FADD2), the exponentials stay scalar MUFUs
__device__ __forceinline__ void tc_elu2(float& x0, float& x1) {
    float m0, m1, e0, e1, r0, r1;
    upk2(mul2(pk2(x0, x1), pk2(1.4426950408889634f, 1.4426950408889634f)), m0, m1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(m0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(m1));
    upk2(add2(pk2(e0, e1), pk2(-1.f, -1.f)), r0, r1);
    x0 = x0 > 0.f ? x0 : r0;
    x1 = x1 > 0.f ? x1 : r1;
}
// y = elu(y + b)
template <int N> __device__ __forceinline__ void bias_elu(const float* __restrict__ b, float* y) {
#pragma unroll
    for (int n = 0; n < N; n += 4) {
        const float4 w = *reinterpret_cast<const float4*>(b + n);
        upk2(add2(pk2(y[n], y[n + 1]), pk2(w.x, w.y)), y[n], y[n + 1]);
        upk2(add2(pk2(y[n + 2], y[n + 3]), pk2(w.z, w.w)), y[n + 2], y[n + 3]);
        tc_elu2(y[n], y[n + 1]); tc_elu2(y[n + 2], y[n + 3]);
    }
}
template <int N> __device__ __forceinline__ void add_bias(const float* __restrict__ b, float* y) {
#pragma unroll
    for (int n = 0; n < N; n += 4) {
        const float4 w = *reinterpret_cast<const float4*>(b + n);
        y[n] += w.x; y[n + 1] += w.y; y[n + 2] += w.z; y[n + 3] += w.w;
    }
}

// ---- per-slot context and the GEMM round protocol shared by the tensor-core K2a kernels ----------------------------------
struct TcCtx {
    uint32_t tmem_slot;       // TMEM base of this slot (lane 0)
    uint32_t lane_addr;       // tmem_slot + (warp%4 * 32 << 16)
    uint32_t img_base16;      // (shared address of the image area) >> 4
    uint32_t bar;             // shared address of this slot's mbarrier
    uint32_t parity;
    int bar_id;               // named barrier id of this slot
    bool issuer;              // this warp issues the slot's MMAs (warp-uniform)
    uint32_t elected;         // 1 on the one lane of the issuing warp that actually issues
};

// all 128 threads of the slot: A operand written -> (leader issues) -> accumulator ready
#define TC_GEMM_BEGIN(cx)                                                            \
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");                     \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                 \
    asm volatile("bar.sync %0, 128;" :: "r"((cx).bar_id) : "memory");                \
    if ((cx).issuer) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#define TC_GEMM_COMMIT(cx)                                                           \
        tc_commit((cx).bar, (cx).elected); __syncwarp(); }
#define TC_GEMM_WAIT(cx)                                                             \
    mbar_wait((cx).bar, (cx).parity); (cx).parity ^= 1u;                             \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#define TC_GEMM_END(cx) TC_GEMM_COMMIT(cx) TC_GEMM_WAIT(cx)

