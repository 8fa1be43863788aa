// Device helpers shared by the fused conv kernels (conv_fused.cu, conv_pipe.cu): mbarrier / tcgen05 wrappers,
// packed fp32x2 math, shared-memory vector accesses and the AdaIN-affine + activation transform of 4 channels.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace st2 {

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// elect.sync: exactly one lane of the (converged) warp gets true -- unlike `lane == 0` the compiler then knows the
// guarded region is single-threaded, so UTCHMMA / UTMALDG operands go to uniform registers without a waterfall loop
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Every lane waits on the mbarrier itself.  (Polling from lane 0 only and re-joining with __syncwarp() leaves the warp
// permanently split into {lane 0} and {lanes 1..31} -- ncu showed 16 active threads per instruction and every
// instruction of the role loop issued twice.)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    mbar_wait(bar, parity);
    __syncwarp();
}

// tcgen05.mma with the two 64-bit shared-memory descriptors given as (lo, hi) 32-bit halves: the issue loop only
// does 32-bit adds on the start-address field
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                              uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accum)
        : "memory");
}
// the same with a descriptor hi word per operand (operands of different layouts / strides)
__device__ __forceinline__ void umma_f16_lohi2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
        : "memory");
}
// hi word of the K-major SWIZZLE_128B descriptor: SBO = 1024 B (>>4) | version 1 @ bit 46 | layout 2 @ bit 61
static constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack16(float lo, float hi, int is_bf16) {
    if (is_bf16) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
    }
    __half2 t = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
// ---- small PTX helpers for the hot loops: 32-bit shared addresses, packed fp32x2 math (FFMA2/FADD2/FMUL2) ----
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

// streaming 128-bit global load that does not allocate in L1 (residual / accumulate rows are read exactly once)
__device__ __forceinline__ float4 ldg_stream(const float* ptr) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
    return v;
}

__device__ __forceinline__ float4 unpack16x4(uint2 u, int is_bf16) {
    float4 v;
    if (is_bf16) {
        v.x = __uint_as_float(u.x << 16); v.y = __uint_as_float(u.x & 0xffff0000u);
        v.z = __uint_as_float(u.y << 16); v.w = __uint_as_float(u.y & 0xffff0000u);
    } else {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        v = make_float4(a.x, a.y, b.x, b.y);
    }
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// per-thread transform constants: y = act(a*x + b) for 4 consecutive channels, packed as two float2
struct XfCoef { float2 a01, a23, b01, b23, al01, al23, ia01, ia23; };

template <int ACT>
__device__ __forceinline__ uint2 transform4(const float4 v, const XfCoef& c, int is_bf16) {
    float2 y01 = ffma2(c.a01, make_float2(v.x, v.y), c.b01);
    float2 y23 = ffma2(c.a23, make_float2(v.z, v.w), c.b23);
    if (ACT == ACT_SNAKE) {
        const float2 t01 = fmul2(c.al01, y01), t23 = fmul2(c.al23, y23);
        const float2 s01 = make_float2(__sinf(t01.x), __sinf(t01.y)), s23 = make_float2(__sinf(t23.x), __sinf(t23.y));
        y01 = ffma2(fmul2(c.ia01, s01), s01, y01);
        y23 = ffma2(fmul2(c.ia23, s23), s23, y23);
    } else if (ACT == ACT_LRELU) {
        y01.x = y01.x >= 0.f ? y01.x : y01.x * c.al01.x; y01.y = y01.y >= 0.f ? y01.y : y01.y * c.al01.x;
        y23.x = y23.x >= 0.f ? y23.x : y23.x * c.al01.x; y23.y = y23.y >= 0.f ? y23.y : y23.y * c.al01.x;
    }
    return make_uint2(pack16(y01.x, y01.y, is_bf16), pack16(y23.x, y23.y, is_bf16));
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still running; pdl_wait() blocks until the predecessor grid has completed and its writes are
// visible (a no-op for a plain launch), pdl_trigger() lets the successor's CTAs be scheduled as resources free up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace st2
