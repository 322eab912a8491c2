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

// Lane-generic conditionals: every value-dependent choice in the stage functions goes through sel(mask, a, b) so that
// the same templates instantiate for float / double (mask = bool) and for the packed two-environment type F2 of the
// sm_100a kernels (mask = M2, ppr_f2.h), where a ternary on a pair of lanes cannot be written.
template <class T> PPR_HD T sel(bool c, T a, T b) { return c ? a : b; }
// Fused forms written out explicitly: for float / double the compiler would contract a * b + c anyway; for F2 they map
// to ONE packed FFMA2 with the negation folded into an operand modifier (ppr_f2.h) instead of FMUL2 + FADD2 + 2 negates.
template <class T> PPR_HD T fma_(T a, T b, T c) { return a * b + c; }    //  a b + c
template <class T> PPR_HD T fnma_(T a, T b, T c) { return c - a * b; }   // -a b + c
template <class T> PPR_HD T fms_(T a, T b, T c) { return a * b - c; }    //  a b - c

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
template <class T> PPR_HD T dot(V3<T> a, V3<T> b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
template <class T> PPR_HD V3<T> cross(V3<T> a, V3<T> b) {
    return v3<T>(fnma_(a.z, b.y, a.y * b.z), fnma_(a.x, b.z, a.z * b.x), fnma_(a.y, b.x, a.x * b.y));
}

template <class T> PPR_HD Q4<T> operator+(Q4<T> a, Q4<T> b) { return q4<T>(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
template <class T> PPR_HD Q4<T> operator*(Q4<T> a, T s) { return q4<T>(a.x * s, a.y * s, a.z * s, a.w * s); }
template <class T> PPR_HD void operator+=(Q4<T>& a, Q4<T> b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
template <class T> PPR_HD T qdot(Q4<T> a, Q4<T> b) { return fma_(a.w, b.w, fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x))); }
template <class T> PPR_HD V3<T> qvec(Q4<T> q) { return v3<T>(q.x, q.y, q.z); }
template <class T> PPR_HD Q4<T> qconj(Q4<T> q) { return q4<T>(-q.x, -q.y, -q.z, q.w); }

// Hamilton product (xyzw). Adjoint: adj_a += adj_c * conj(b); adj_b += conj(a) * adj_c.
template <class T> PPR_HD Q4<T> qmul(Q4<T> a, Q4<T> b) {
    return q4<T>(fnma_(a.z, b.y, fma_(a.y, b.z, fma_(b.w, a.x, a.w * b.x))),
                 fnma_(a.x, b.z, fma_(a.z, b.x, fma_(b.w, a.y, a.w * b.y))),
                 fnma_(a.y, b.x, fma_(a.x, b.y, fma_(b.w, a.z, a.w * b.z))),
                 fnma_(a.z, b.z, fnma_(a.y, b.y, fnma_(a.x, b.x, a.w * b.w))));
}

// Warp quat_rotate / quat_rotate_inv (linear in v; transposes of each other).
template <class T> PPR_HD V3<T> qrot(Q4<T> q, V3<T> v) {
    V3<T> u = qvec(q);
    T b = T(2) * q.w, a = fms_(b, q.w, T(1)), c = T(2) * dot(u, v);
    V3<T> uxv = cross(u, v);
    return v3<T>(fma_(u.x, c, fma_(uxv.x, b, v.x * a)), fma_(u.y, c, fma_(uxv.y, b, v.y * a)),
                 fma_(u.z, c, fma_(uxv.z, b, v.z * a)));
}
template <class T> PPR_HD V3<T> qrot_inv(Q4<T> q, V3<T> v) {
    V3<T> u = qvec(q);
    T b = T(2) * q.w, a = fms_(b, q.w, T(1)), c = T(2) * dot(u, v);
    V3<T> uxv = cross(u, v);
    return v3<T>(fma_(u.x, c, fnma_(uxv.x, b, v.x * a)), fma_(u.y, c, fnma_(uxv.y, b, v.y * a)),
                 fma_(u.z, c, fnma_(uxv.z, b, v.z * a)));
}
// d(g . qrot(q,v))/dq
template <class T> PPR_HD Q4<T> qrot_adj_q(Q4<T> q, V3<T> v, V3<T> g) {
    V3<T> u = qvec(q);
    V3<T> uxv = cross(u, v), vxg = cross(v, g);
    T uv = dot(u, v), ug = dot(u, g);
    T tw = T(2) * q.w;
    T aw = T(2) * fma_(tw, dot(v, g), dot(uxv, g));
    T uv2 = T(2) * uv, ug2 = T(2) * ug;
    return q4<T>(fma_(tw, vxg.x, fma_(uv2, g.x, ug2 * v.x)), fma_(tw, vxg.y, fma_(uv2, g.y, ug2 * v.y)),
                 fma_(tw, vxg.z, fma_(uv2, g.z, ug2 * v.z)), aw);
}
// d(g . qrot_inv(q,v))/dq
template <class T> PPR_HD Q4<T> qrotinv_adj_q(Q4<T> q, V3<T> v, V3<T> g) {
    V3<T> u = qvec(q);
    V3<T> uxv = cross(u, v), vxg = cross(v, g);
    T uv = dot(u, v), ug = dot(u, g);
    T tw = T(2) * q.w;
    T aw = T(2) * fms_(tw, dot(v, g), dot(uxv, g));
    T uv2 = T(2) * uv, ug2 = T(2) * ug;
    return q4<T>(fnma_(tw, vxg.x, fma_(uv2, g.x, ug2 * v.x)), fnma_(tw, vxg.y, fma_(uv2, g.y, ug2 * v.y)),
                 fnma_(tw, vxg.z, fma_(uv2, g.z, ug2 * v.z)), aw);
}

// ---- the same rotation as a 3x3 matrix -------------------------------------------------------------------
// quat_rotate(q, v) is LINEAR in v for any (also non-unit) q:  M(q) = (2w^2-1) I + 2w [u]x + 2 u u^T.
// A body that rotates many vectors by its own quaternion per substep builds M once (18 flops) and pays 9 flops per
// rotation instead of ~21; in the adjoint every d/dq of such a rotation becomes a rank-1 update of G = dL/dM
// (9 flops) and ONE conversion G -> dL/dq per body and substep (qmat_adj) instead of ~35 flops per rotation.
template <class T> struct M3 { T m[9]; };  // row-major
template <class T> PPR_HD M3<T> m3_zero() { M3<T> r; PPR_UNROLL for (int i = 0; i < 9; ++i) r.m[i] = T(0); return r; }
template <class T> PPR_HD M3<T> qmat(Q4<T> q) {
    T tw = T(2) * q.w, a = fms_(tw, q.w, T(1));
    T tx = T(2) * q.x, ty = T(2) * q.y, tz = T(2) * q.z;
    T xy = tx * q.y, xz = tx * q.z, yz = ty * q.z;
    M3<T> r;
    r.m[0] = fma_(tx, q.x, a);    r.m[1] = fnma_(tw, q.z, xy); r.m[2] = fma_(tw, q.y, xz);
    r.m[3] = fma_(tw, q.z, xy);   r.m[4] = fma_(ty, q.y, a);   r.m[5] = fnma_(tw, q.x, yz);
    r.m[6] = fnma_(tw, q.y, xz);  r.m[7] = fma_(tw, q.x, yz);  r.m[8] = fma_(tz, q.z, a);
    return r;
}
template <class T> PPR_HD V3<T> mrot(const M3<T>& M, V3<T> v) {   // = quat_rotate(q, v)
    return v3<T>(fma_(M.m[2], v.z, fma_(M.m[1], v.y, M.m[0] * v.x)), fma_(M.m[5], v.z, fma_(M.m[4], v.y, M.m[3] * v.x)),
                 fma_(M.m[8], v.z, fma_(M.m[7], v.y, M.m[6] * v.x)));
}
template <class T> PPR_HD V3<T> mrot_t(const M3<T>& M, V3<T> v) { // = quat_rotate_inv(q, v)
    return v3<T>(fma_(M.m[6], v.z, fma_(M.m[3], v.y, M.m[0] * v.x)), fma_(M.m[7], v.z, fma_(M.m[4], v.y, M.m[1] * v.x)),
                 fma_(M.m[8], v.z, fma_(M.m[5], v.y, M.m[2] * v.x)));
}
// adjoint bookkeeping: y = M v  with adjoint g  ->  G += g v^T ;   y = M^T v with adjoint g  ->  G += v g^T
template <class T> PPR_HD void m3_acc(M3<T>& G, V3<T> a, V3<T> b) {  // G += a b^T
    G.m[0] = fma_(a.x, b.x, G.m[0]); G.m[1] = fma_(a.x, b.y, G.m[1]); G.m[2] = fma_(a.x, b.z, G.m[2]);
    G.m[3] = fma_(a.y, b.x, G.m[3]); G.m[4] = fma_(a.y, b.y, G.m[4]); G.m[5] = fma_(a.y, b.z, G.m[5]);
    G.m[6] = fma_(a.z, b.x, G.m[6]); G.m[7] = fma_(a.z, b.y, G.m[7]); G.m[8] = fma_(a.z, b.z, G.m[8]);
}
// dL/dq from G = dL/dM(q)
template <class T> PPR_HD Q4<T> qmat_adj(Q4<T> q, const M3<T>& G) {
    T ax = G.m[7] - G.m[5], ay = G.m[2] - G.m[6], az = G.m[3] - G.m[1];   // a_k = sum_ij eps_ikj G_ij
    T tr = G.m[0] + G.m[4] + G.m[8];
    T s01 = G.m[1] + G.m[3], s02 = G.m[2] + G.m[6], s12 = G.m[5] + G.m[7];
    T sx = fma_(s02, q.z, fma_(s01, q.y, (G.m[0] + G.m[0]) * q.x));  // ((G + G^T) u)
    T sy = fma_(s12, q.z, fma_(G.m[4] + G.m[4], q.y, s01 * q.x));
    T sz = fma_(G.m[8] + G.m[8], q.z, fma_(s12, q.y, s02 * q.x));
    T tw = T(2) * q.w;
    return q4<T>(fma_(tw, ax, T(2) * sx), fma_(tw, ay, T(2) * sy), fma_(tw, az, T(2) * sz),
                 T(2) * fma_(tw, tr, fma_(q.z, az, fma_(q.y, ay, q.x * ax))));
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
    T inv = sel(len > T(0), T(1) / len, T(0));
    return q * inv;
}
// y = q/|q| (already computed), len = |q|
template <class T> PPR_HD Q4<T> qnormalize_adj(Q4<T> y, T len, Q4<T> g) {
    T inv = sel(len > T(0), T(1) / len, T(0)), yg = qdot(y, g);   // zero adjoint at the singular point
    return q4<T>((g.x - y.x * yg) * inv, (g.y - y.y * yg) * inv, (g.z - y.z * yg) * inv, (g.w - y.w * yg) * inv);
}

template <class T> PPR_HD T clampT(T x, T lo, T hi) { return sel(x < lo, lo, sel(x > hi, hi, x)); }
template <class T> PPR_HD T clamp_mask(T x, T lo, T hi) { return sel((x < lo) || (x > hi), T(0), T(1)); }
template <class T> PPR_HD V3<T> clamp3(V3<T> a, T lim) {
    return v3<T>(clampT(a.x, -lim, lim), clampT(a.y, -lim, lim), clampT(a.z, -lim, lim));
}
template <class T> PPR_HD V3<T> clamp3_mask(V3<T> a, T lim, V3<T> g) {
    return v3<T>(g.x * clamp_mask(a.x, -lim, lim), g.y * clamp_mask(a.y, -lim, lim), g.z * clamp_mask(a.z, -lim, lim));
}
template <class T> PPR_HD T safe_acos(T x) { return acos(clampT(x, T(-1), T(1))); }
template <class T> PPR_HD T safe_acos_adj(T x) {  // d acos / dx, 0 at saturation
    T d = T(1) - x * x;
    return sel(d > T(0), T(-1) / sqrt(d), T(0));
}
template <class T> PPR_HD T safe_asin(T x) { return asin(clampT(x, T(-1), T(1))); }
template <class T> PPR_HD T safe_asin_adj(T x) {
    T d = T(1) - x * x;
    return sel(d > T(0), T(1) / sqrt(d), T(0));
}

// 3x3 row-major helpers (I[3*i+j])
template <class T> PPR_HD V3<T> matvec(const T* M, V3<T> v) {
    return v3<T>(fma_(M[2], v.z, fma_(M[1], v.y, M[0] * v.x)), fma_(M[5], v.z, fma_(M[4], v.y, M[3] * v.x)),
                 fma_(M[8], v.z, fma_(M[7], v.y, M[6] * v.x)));
}
template <class T> PPR_HD V3<T> matTvec(const T* M, V3<T> v) {
    return v3<T>(fma_(M[6], v.z, fma_(M[3], v.y, M[0] * v.x)), fma_(M[7], v.z, fma_(M[4], v.y, M[1] * v.x)),
                 fma_(M[8], v.z, fma_(M[5], v.y, M[2] * v.x)));
}
template <class T> PPR_HD void outer_acc(T* M, V3<T> a, V3<T> b, T s) {  // M += s * a b^T
    M[0] += s * a.x * b.x; M[1] += s * a.x * b.y; M[2] += s * a.x * b.z;
    M[3] += s * a.y * b.x; M[4] += s * a.y * b.y; M[5] += s * a.y * b.z;
    M[6] += s * a.z * b.x; M[7] += s * a.z * b.y; M[8] += s * a.z * b.z;
}

}  // namespace ppr
