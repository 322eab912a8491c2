// Small vector / quaternion algebra with hand-written reverse-mode adjoints.
//
// Semantics follow the Warp built-ins the reference kernels call (quaternions xyzw, quat_rotate is the
// NON-normalising formula v(2w^2-1) + 2w(u x v) + 2u(u.v); normalize/acos/asin return 0 adjoint at their
// singular points -- the reference scrubs the resulting NaNs to 0, diffphys/dp_utils.py:53).
// The header is scalar-templated and host/device so that the sm_100a kernels (float) and the CPU port used
// as the timed CPU baseline (float / double) share one statement of the arithmetic.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PPR_HD __host__ __device__ __forceinline__
#define PPR_UNROLL _Pragma("unroll")
#else
#define PPR_HD inline
#define PPR_UNROLL
#endif

namespace ppr {

template <class T> struct V3 { T x, y, z; };
template <class T> struct Q4 { T x, y, z, w; };

template <class T> PPR_HD V3<T> v3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <class T> PPR_HD Q4<T> q4(T x, T y, T z, T w) { Q4<T> r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
template <class T> PPR_HD V3<T> vzero() { return v3<T>(T(0), T(0), T(0)); }
template <class T> PPR_HD Q4<T> qzero() { return q4<T>(T(0), T(0), T(0), T(0)); }

template <class T> PPR_HD V3<T> operator+(V3<T> a, V3<T> b) { return v3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> PPR_HD V3<T> operator-(V3<T> a, V3<T> b) { return v3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> PPR_HD V3<T> operator-(V3<T> a) { return v3<T>(-a.x, -a.y, -a.z); }
template <class T> PPR_HD V3<T> operator*(V3<T> a, T s) { return v3<T>(a.x * s, a.y * s, a.z * s); }
template <class T> PPR_HD V3<T> operator*(T s, V3<T> a) { return v3<T>(a.x * s, a.y * s, a.z * s); }
template <class T> PPR_HD void operator+=(V3<T>& a, V3<T> b) { a.x += b.x; a.y += b.y; a.z += b.z; }
template <class T> PPR_HD void operator-=(V3<T>& a, V3<T> b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
template <class T> PPR_HD T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> PPR_HD V3<T> cross(V3<T> a, V3<T> b) {
    return v3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

template <class T> PPR_HD Q4<T> operator+(Q4<T> a, Q4<T> b) { return q4<T>(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
template <class T> PPR_HD Q4<T> operator*(Q4<T> a, T s) { return q4<T>(a.x * s, a.y * s, a.z * s, a.w * s); }
template <class T> PPR_HD void operator+=(Q4<T>& a, Q4<T> b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
template <class T> PPR_HD T qdot(Q4<T> a, Q4<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
template <class T> PPR_HD V3<T> qvec(Q4<T> q) { return v3<T>(q.x, q.y, q.z); }
template <class T> PPR_HD Q4<T> qconj(Q4<T> q) { return q4<T>(-q.x, -q.y, -q.z, q.w); }

// Hamilton product (xyzw). Adjoint: adj_a += adj_c * conj(b); adj_b += conj(a) * adj_c.
template <class T> PPR_HD Q4<T> qmul(Q4<T> a, Q4<T> b) {
    return q4<T>(a.w * b.x + b.w * a.x + a.y * b.z - a.z * b.y,
                 a.w * b.y + b.w * a.y + a.z * b.x - a.x * b.z,
                 a.w * b.z + b.w * a.z + a.x * b.y - a.y * b.x,
                 a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}

// Warp quat_rotate / quat_rotate_inv (linear in v; transposes of each other).
template <class T> PPR_HD V3<T> qrot(Q4<T> q, V3<T> v) {
    V3<T> u = qvec(q);
    T a = T(2) * q.w * q.w - T(1), b = T(2) * q.w, c = T(2) * dot(u, v);
    V3<T> uxv = cross(u, v);
    return v3<T>(v.x * a + uxv.x * b + u.x * c, v.y * a + uxv.y * b + u.y * c, v.z * a + uxv.z * b + u.z * c);
}
template <class T> PPR_HD V3<T> qrot_inv(Q4<T> q, V3<T> v) {
    V3<T> u = qvec(q);
    T a = T(2) * q.w * q.w - T(1), b = T(2) * q.w, c = T(2) * dot(u, v);
    V3<T> uxv = cross(u, v);
    return v3<T>(v.x * a - uxv.x * b + u.x * c, v.y * a - uxv.y * b + u.y * c, v.z * a - uxv.z * b + u.z * c);
}
// d(g . qrot(q,v))/dq
template <class T> PPR_HD Q4<T> qrot_adj_q(Q4<T> q, V3<T> v, V3<T> g) {
    V3<T> u = qvec(q);
    V3<T> uxv = cross(u, v), vxg = cross(v, g);
    T uv = dot(u, v), ug = dot(u, g);
    T aw = T(4) * q.w * dot(v, g) + T(2) * dot(uxv, g);
    T tw = T(2) * q.w;
    return q4<T>(tw * vxg.x + T(2) * (uv * g.x + ug * v.x), tw * vxg.y + T(2) * (uv * g.y + ug * v.y),
                 tw * vxg.z + T(2) * (uv * g.z + ug * v.z), aw);
}
// d(g . qrot_inv(q,v))/dq
template <class T> PPR_HD Q4<T> qrotinv_adj_q(Q4<T> q, V3<T> v, V3<T> g) {
    V3<T> u = qvec(q);
    V3<T> uxv = cross(u, v), vxg = cross(v, g);
    T uv = dot(u, v), ug = dot(u, g);
    T aw = T(4) * q.w * dot(v, g) - T(2) * dot(uxv, g);
    T tw = T(2) * q.w;
    return q4<T>(-tw * vxg.x + T(2) * (uv * g.x + ug * v.x), -tw * vxg.y + T(2) * (uv * g.y + ug * v.y),
                 -tw * vxg.z + T(2) * (uv * g.z + ug * v.z), aw);
}

// ---- the same rotation as a 3x3 matrix -------------------------------------------------------------------
// quat_rotate(q, v) is LINEAR in v for any (also non-unit) q:  M(q) = (2w^2-1) I + 2w [u]x + 2 u u^T.
// A body that rotates many vectors by its own quaternion per substep builds M once (18 flops) and pays 9 flops per
// rotation instead of ~21; in the adjoint every d/dq of such a rotation becomes a rank-1 update of G = dL/dM
// (9 flops) and ONE conversion G -> dL/dq per body and substep (qmat_adj) instead of ~35 flops per rotation.
template <class T> struct M3 { T m[9]; };  // row-major
template <class T> PPR_HD M3<T> m3_zero() { M3<T> r; PPR_UNROLL for (int i = 0; i < 9; ++i) r.m[i] = T(0); return r; }
template <class T> PPR_HD M3<T> qmat(Q4<T> q) {
    T a = T(2) * q.w * q.w - T(1), tw = T(2) * q.w;
    T xx = T(2) * q.x * q.x, yy = T(2) * q.y * q.y, zz = T(2) * q.z * q.z;
    T xy = T(2) * q.x * q.y, xz = T(2) * q.x * q.z, yz = T(2) * q.y * q.z;
    M3<T> r;
    r.m[0] = a + xx;        r.m[1] = xy - tw * q.z; r.m[2] = xz + tw * q.y;
    r.m[3] = xy + tw * q.z; r.m[4] = a + yy;        r.m[5] = yz - tw * q.x;
    r.m[6] = xz - tw * q.y; r.m[7] = yz + tw * q.x; r.m[8] = a + zz;
    return r;
}
template <class T> PPR_HD V3<T> mrot(const M3<T>& M, V3<T> v) {   // = quat_rotate(q, v)
    return v3<T>(M.m[0] * v.x + M.m[1] * v.y + M.m[2] * v.z, M.m[3] * v.x + M.m[4] * v.y + M.m[5] * v.z,
                 M.m[6] * v.x + M.m[7] * v.y + M.m[8] * v.z);
}
template <class T> PPR_HD V3<T> mrot_t(const M3<T>& M, V3<T> v) { // = quat_rotate_inv(q, v)
    return v3<T>(M.m[0] * v.x + M.m[3] * v.y + M.m[6] * v.z, M.m[1] * v.x + M.m[4] * v.y + M.m[7] * v.z,
                 M.m[2] * v.x + M.m[5] * v.y + M.m[8] * v.z);
}
// adjoint bookkeeping: y = M v  with adjoint g  ->  G += g v^T ;   y = M^T v with adjoint g  ->  G += v g^T
template <class T> PPR_HD void m3_acc(M3<T>& G, V3<T> a, V3<T> b) {  // G += a b^T
    G.m[0] += a.x * b.x; G.m[1] += a.x * b.y; G.m[2] += a.x * b.z;
    G.m[3] += a.y * b.x; G.m[4] += a.y * b.y; G.m[5] += a.y * b.z;
    G.m[6] += a.z * b.x; G.m[7] += a.z * b.y; G.m[8] += a.z * b.z;
}
// dL/dq from G = dL/dM(q)
template <class T> PPR_HD Q4<T> qmat_adj(Q4<T> q, const M3<T>& G) {
    T ax = G.m[7] - G.m[5], ay = G.m[2] - G.m[6], az = G.m[3] - G.m[1];   // a_k = sum_ij eps_ikj G_ij
    T tr = G.m[0] + G.m[4] + G.m[8];
    T sx = (G.m[0] + G.m[0]) * q.x + (G.m[1] + G.m[3]) * q.y + (G.m[2] + G.m[6]) * q.z;  // ((G + G^T) u)
    T sy = (G.m[3] + G.m[1]) * q.x + (G.m[4] + G.m[4]) * q.y + (G.m[5] + G.m[7]) * q.z;
    T sz = (G.m[6] + G.m[2]) * q.x + (G.m[7] + G.m[5]) * q.y + (G.m[8] + G.m[8]) * q.z;
    T tw = T(2) * q.w;
    return q4<T>(tw * ax + T(2) * sx, tw * ay + T(2) * sy, tw * az + T(2) * sz,
                 T(4) * q.w * tr + T(2) * (q.x * ax + q.y * ay + q.z * az));
}

template <class T> PPR_HD Q4<T> q_axis_angle(V3<T> a, T ang) {
    T h = T(0.5) * ang, s = sin(h), c = cos(h);
    return q4<T>(a.x * s, a.y * s, a.z * s, c);
}
// adjoint of q = (a sin(h), cos(h)), h = ang/2 : returns adj_ang, accumulates adj_a
template <class T> PPR_HD T q_axis_angle_adj(V3<T> a, T ang, Q4<T> g, V3<T>& adj_a) {
    T h = T(0.5) * ang, s = sin(h), c = cos(h);
    adj_a += qvec(g) * s;
    return T(0.5) * (c * dot(a, qvec(g)) - s * g.w);
}

template <class T> PPR_HD Q4<T> qnormalize(Q4<T> q, T& len) {
    len = sqrt(qdot(q, q));
    T inv = len > T(0) ? T(1) / len : T(0);
    return q * inv;
}
// y = q/|q| (already computed), len = |q|
template <class T> PPR_HD Q4<T> qnormalize_adj(Q4<T> y, T len, Q4<T> g) {
    if (!(len > T(0))) return qzero<T>();
    T inv = T(1) / len, yg = qdot(y, g);
    return q4<T>((g.x - y.x * yg) * inv, (g.y - y.y * yg) * inv, (g.z - y.z * yg) * inv, (g.w - y.w * yg) * inv);
}

template <class T> PPR_HD T clampT(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <class T> PPR_HD T clamp_mask(T x, T lo, T hi) { return (x < lo || x > hi) ? T(0) : T(1); }
template <class T> PPR_HD V3<T> clamp3(V3<T> a, T lim) {
    return v3<T>(clampT(a.x, -lim, lim), clampT(a.y, -lim, lim), clampT(a.z, -lim, lim));
}
template <class T> PPR_HD V3<T> clamp3_mask(V3<T> a, T lim, V3<T> g) {
    return v3<T>(g.x * clamp_mask(a.x, -lim, lim), g.y * clamp_mask(a.y, -lim, lim), g.z * clamp_mask(a.z, -lim, lim));
}
template <class T> PPR_HD T safe_acos(T x) { return acos(clampT(x, T(-1), T(1))); }
template <class T> PPR_HD T safe_acos_adj(T x) {  // d acos / dx, 0 at saturation
    T d = T(1) - x * x;
    return d > T(0) ? T(-1) / sqrt(d) : T(0);
}
template <class T> PPR_HD T safe_asin(T x) { return asin(clampT(x, T(-1), T(1))); }
template <class T> PPR_HD T safe_asin_adj(T x) {
    T d = T(1) - x * x;
    return d > T(0) ? T(1) / sqrt(d) : T(0);
}

// 3x3 row-major helpers (I[3*i+j])
template <class T> PPR_HD V3<T> matvec(const T* M, V3<T> v) {
    return v3<T>(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z,
                 M[6] * v.x + M[7] * v.y + M[8] * v.z);
}
template <class T> PPR_HD V3<T> matTvec(const T* M, V3<T> v) {
    return v3<T>(M[0] * v.x + M[3] * v.y + M[6] * v.z, M[1] * v.x + M[4] * v.y + M[7] * v.z,
                 M[2] * v.x + M[5] * v.y + M[8] * v.z);
}
template <class T> PPR_HD void outer_acc(T* M, V3<T> a, V3<T> b, T s) {  // M += s * a b^T
    M[0] += s * a.x * b.x; M[1] += s * a.x * b.y; M[2] += s * a.x * b.z;
    M[3] += s * a.y * b.x; M[4] += s * a.y * b.y; M[5] += s * a.y * b.z;
    M[6] += s * a.z * b.x; M[7] += s * a.z * b.y; M[8] += s * a.z * b.z;
}

}  // namespace ppr
