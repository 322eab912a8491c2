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
