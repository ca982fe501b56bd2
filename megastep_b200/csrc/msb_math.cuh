// msb_math.cuh — pinned fp32 arithmetic for the sm_100a kernels.
//
// Results must match the reference's own CUDA build bit-for-bit on hit/collision decisions, so every operation on
// the decision path is spelled with a rounding-explicit intrinsic (never contracted or re-associated by the
// compiler) in exactly the order the reference compiles to (docs/REFERENCE_ARITHMETIC.md, decoded from its SASS).
// The translation unit is built with -ftz=true so these become the .FTZ forms the reference's --use_fast_math
// build uses; approximate reciprocal / square root are the raw MUFU ops via PTX.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace msb {

__device__ __forceinline__ float rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// a*b - c*d as the reference compiles every 2-D cross product: ffma(a, b, -(c*d))
__device__ __forceinline__ float cross2(float a, float b, float c, float d) { return ffma(a, b, -fmul(c, d)); }
// a*b + c*d as the reference compiles every 2-D dot product: ffma(a, b, c*d)
__device__ __forceinline__ float dot2(float a, float b, float c, float d) { return ffma(a, b, fmul(c, d)); }

constexpr float K180 = 0.0055555556900799274445f;   // angle/180.f after ptxas constant-folds the reciprocal
constexpr float PARALLEL_EPS = 1.e-3f;               // kernels.cu:77

// sincos of an angle in degrees exactly as draw_kernel / raycast_kernel get them (kernels.cu:304-306, 335-337).
__device__ __forceinline__ void sincos_deg(float angle, float& s, float& c) {
    const float a = fmul(angle, K180);
    c = cospif(a);
    s = sinpif(a);
}

struct Hit { float s, t; };

// intersect(P, U, Q, V) (kernels.cu:67-88) with the line-only terms hoisted by the caller:
//   V = b - a, PQ = a - P, snum = cross2(Vy, PQx, Vx, PQy)
__device__ __forceinline__ Hit intersect_pre(float Ux, float Uy, float Vx, float Vy, float PQx, float PQy, float snum) {
    const float UxV = cross2(Ux, Vy, Uy, Vx);
    Hit h;
    if (fabsf(UxV) < PARALLEL_EPS) {
        h.s = CUDART_INF_F;
        h.t = CUDART_INF_F;
    } else {
        const float rc = rcp(UxV);
        h.s = fmul(snum, rc);
        h.t = fmul(cross2(Uy, PQx, Ux, PQy), rc);
    }
    return h;
}

__device__ __forceinline__ Hit intersect(float Px, float Py, float Ux, float Uy, float4 l) {
    const float Vx = fsub(l.z, l.x), Vy = fsub(l.w, l.y);
    const float PQx = fsub(l.x, Px), PQy = fsub(l.y, Py);
    return intersect_pre(Ux, Uy, Vx, Vy, PQx, PQy, cross2(Vy, PQx, Vx, PQy));
}

// sensibilize() (kernels.cu:109-118): NaN -> 0, else clamp(0.99 p, 0, 1)
__device__ __forceinline__ float sens(float p) { return isnan(p) ? 0.f : __saturatef(ffma(p, 0.99f, 0.f)); }

// at::remainder on floats (Python-style modulo) as used by normalize_degrees (kernels.cu:173-175)
__device__ __forceinline__ float remainder_(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.f && ((b < 0.f) != (m < 0.f))) m = __fadd_rn(m, b);
    return m;
}

__device__ __forceinline__ float warp_min(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fminf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// ---- mbarrier + 1-D bulk (TMA) copy ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP), completion counted in bytes on `bar`. 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// hint: bring `bytes` (a multiple of 16, from a 16-byte aligned address) into L2; one instruction for the whole range
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// hint: bring the 128-byte line at p into L2 (no register result, nothing to wait for)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// release-ordered publication at GPU scope: every write that happens-before (this thread's, and those of threads it
// synchronised with through a warp / CTA barrier) is visible to whoever observes the stored value
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release_add(int* p, int v) { asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while the
// kernel ahead of it in the stream is still running; it must not touch anything that kernel writes before
// griddep_wait() returns (= the whole primary grid has completed and its writes are visible). griddep_launch() is the
// primary's side: "my dependents may be scheduled as soon as every CTA of mine got here". Both are no-ops otherwise.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace msb
