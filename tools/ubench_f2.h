// F2: two fp32 lanes in one 64-bit register pair, computed with Blackwell's packed FP32 instructions
// (PTX add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2, sm_100+).
//
// Why: on B200 a scalar 3-register FFMA issues at ~0.58 warp-instructions/clk/SMSP (operand bandwidth) while an FFMA2
// issues at ~0.45 and does two FMAs per lane -- 1.54x the FP32 rate and half the issue slots
// (profiles/r2_ubench_fp32_issue.txt).  The rollout kernels are bound by exactly that rate, so the throughput layout
// gives every thread the same rigid body of TWO environments and instantiates the scalar-templated stage functions of
// ppr_body.h with T = F2: lane x = environment A, lane y = environment B.  ptxas folds negations and scalar broadcasts
// into operand modifiers (-R4.F32x2, R0.F32) and fuses mul + add into FFMA2, so the operators below cost one
// instruction each.  Value-dependent choices are masks (M2) consumed by sel().
#pragma once
#include <cuda_runtime.h>

#include "../ppr_diffphys_b200/csrc/ppr_math.h"

namespace ppr {

struct M2 { bool x, y; };
__device__ __forceinline__ M2 operator&&(M2 a, M2 b) { M2 r; r.x = a.x && b.x; r.y = a.y && b.y; return r; }
__device__ __forceinline__ M2 operator||(M2 a, M2 b) { M2 r; r.x = a.x || b.x; r.y = a.y || b.y; return r; }
__device__ __forceinline__ M2 operator!(M2 a) { M2 r; r.x = !a.x; r.y = !a.y; return r; }

// sign flip that ptxas can fold into the consumer's operand modifier even when the build flushes denormals
// (-ftz=true turns a C++ `-x` into FADD.FTZ -x, -0, which stays a separate scalar instruction per lane)
__device__ __forceinline__ float negf(float x) { float r; asm("neg.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(negf(a.x), negf(a.y)); }

struct F2 {
    float2 v;
    F2() = default;
    __device__ __forceinline__ F2(float s) : v(make_float2(s, s)) {}
    __device__ __forceinline__ F2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ explicit F2(float2 p) : v(p) {}
};
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return F2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 operator-(F2 a) { return F2(neg2(a.v)); }
__device__ __forceinline__ F2 fma_(F2 a, F2 b, F2 c) { return F2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ F2 fnma_(F2 a, F2 b, F2 c) { return F2(__ffma2_rn(neg2(a.v), b.v, c.v)); }
__device__ __forceinline__ F2 fms_(F2 a, F2 b, F2 c) { return F2(__ffma2_rn(a.v, b.v, neg2(c.v))); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return F2(__fadd2_rn(a.v, neg2(b.v))); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return F2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 operator/(F2 a, F2 b) { return F2(a.v.x / b.v.x, a.v.y / b.v.y); }
__device__ __forceinline__ void operator+=(F2& a, F2 b) { a = a + b; }
__device__ __forceinline__ void operator-=(F2& a, F2 b) { a = a - b; }
__device__ __forceinline__ void operator*=(F2& a, F2 b) { a = a * b; }
__device__ __forceinline__ M2 operator<(F2 a, F2 b) { M2 r; r.x = a.v.x < b.v.x; r.y = a.v.y < b.v.y; return r; }
__device__ __forceinline__ M2 operator>(F2 a, F2 b) { M2 r; r.x = a.v.x > b.v.x; r.y = a.v.y > b.v.y; return r; }
__device__ __forceinline__ M2 operator<=(F2 a, F2 b) { M2 r; r.x = a.v.x <= b.v.x; r.y = a.v.y <= b.v.y; return r; }
__device__ __forceinline__ M2 operator>=(F2 a, F2 b) { M2 r; r.x = a.v.x >= b.v.x; r.y = a.v.y >= b.v.y; return r; }
__device__ __forceinline__ M2 operator!=(F2 a, F2 b) { M2 r; r.x = a.v.x != b.v.x; r.y = a.v.y != b.v.y; return r; }
__device__ __forceinline__ M2 operator==(F2 a, F2 b) { M2 r; r.x = a.v.x == b.v.x; r.y = a.v.y == b.v.y; return r; }
__device__ __forceinline__ F2 sel(M2 c, F2 a, F2 b) { return F2(c.x ? a.v.x : b.v.x, c.y ? a.v.y : b.v.y); }
__device__ __forceinline__ F2 sqrt(F2 a) { return F2(sqrtf(a.v.x), sqrtf(a.v.y)); }
__device__ __forceinline__ F2 atan2(F2 a, F2 b) { return F2(atan2f(a.v.x, b.v.x), atan2f(a.v.y, b.v.y)); }
__device__ __forceinline__ F2 asin(F2 a) { return F2(asinf(a.v.x), asinf(a.v.y)); }
__device__ __forceinline__ F2 acos(F2 a) { return F2(acosf(a.v.x), acosf(a.v.y)); }
__device__ __forceinline__ F2 sin(F2 a) { return F2(sinf(a.v.x), sinf(a.v.y)); }
__device__ __forceinline__ F2 cos(F2 a) { return F2(cosf(a.v.x), cosf(a.v.y)); }

// lane access: h = 0 -> environment A, 1 -> environment B
template <int H> __device__ __forceinline__ float lane(F2 a) { return H ? a.v.y : a.v.x; }
__device__ __forceinline__ float lane(F2 a, int h) { return h ? a.v.y : a.v.x; }
template <int H> __device__ __forceinline__ V3<float> lane(const V3<F2>& a) {
    return v3<float>(lane<H>(a.x), lane<H>(a.y), lane<H>(a.z));
}
template <int H> __device__ __forceinline__ Q4<float> lane(const Q4<F2>& a) {
    return q4<float>(lane<H>(a.x), lane<H>(a.y), lane<H>(a.z), lane<H>(a.w));
}
__device__ __forceinline__ V3<F2> pack(V3<float> a, V3<float> b) {
    return v3<F2>(F2(a.x, b.x), F2(a.y, b.y), F2(a.z, b.z));
}
__device__ __forceinline__ Q4<F2> pack(Q4<float> a, Q4<float> b) {
    return q4<F2>(F2(a.x, b.x), F2(a.y, b.y), F2(a.z, b.z), F2(a.w, b.w));
}

}  // namespace ppr
