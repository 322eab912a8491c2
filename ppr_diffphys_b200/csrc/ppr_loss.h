// se3 pose / twist loss of the imitation objective and its adjoint, one pair per thread
// (reference: diffphys/dp_utils.py:113-138 `se3_loss`, diffphys/geom_utils.py:37-46 `rot_angle`; call sites
// dp_model.py:777-800).  In the reference this is ~330-400 tiny torch kernels per call (forward + backward);
// at the reference's own problem size (10-64 windows x 24 frames x 13 bodies) the three calls per iteration cost as
// much as the rollout.  Scalar-templated like ppr_body.h.
//
//   loss = |p.xyz - g.xyz|^2 + rot_ratio * acos(clamp((tr(R_p R_g^T) - 1) / 2, -1 + eps, 1 - eps)),   0 for NaN rows
//   dim 7: (xyz, quaternion xyzw), R = rotation of the quaternion (any norm: quaternion_to_matrix divides by |q|^2)
//   dim 6: (xyz, axis-angle),      R = rotation of q = (v sin(|v|/2) / |v|, cos(|v|/2))
#pragma once
#include "ppr_math.h"

namespace ppr {

// The rotation angle is evaluated from the RELATIVE QUATERNION r = q_p (x) conj(q_g):
//     angle = 2 atan2(|r.xyz|, |r.w|)   in [0, pi]
// which equals rot_angle(R_p R_g^T) = acos((tr - 1)/2) but is scale invariant (no normalisation of the inputs needed)
// and well conditioned at small angles, where the trace form loses 3 of float32's 7 digits (1 - cos = 1e-4 at the
// clamp) -- the regime the imitation loop lives in.  The clamp of the cosine to [-1 + eps, 1 - eps] becomes a clamp of
// the angle to [acos(1 - eps), pi - acos(1 - eps)], with zero gradient outside, exactly like torch.clamp.
template <class T> struct Se3Quat {   // rotation operand as a quaternion + what the axis-angle adjoint needs
    Q4<T> q;
    V3<T> v;        // axis-angle input (dim 6)
    T k, th;        // q.xyz = v * k, th = |v|
};

template <class T> PPR_HD Se3Quat<T> se3_quat(int dim, const T* x) {
    Se3Quat<T> r;
    if (dim == 7) {
        r.q = q4<T>(x[3], x[4], x[5], x[6]);
        r.v = vzero<T>(); r.k = T(0); r.th = T(0);
    } else {
        r.v = v3<T>(x[3], x[4], x[5]);
        r.th = sqrt(dot(r.v, r.v));
        T h = T(0.5) * r.th;
        r.k = r.th > T(1e-6) ? sin(h) / r.th : T(0.5) - r.th * r.th / T(48);
        r.q = q4<T>(r.v.x * r.k, r.v.y * r.k, r.v.z * r.k, cos(h));
    }
    return r;
}

// gradient w.r.t. the quaternion -> gradient w.r.t. the 4 (dim 7) or 3 (dim 6) rotation inputs
template <class T> PPR_HD void se3_quat_adj(int dim, const Se3Quat<T>& r, Q4<T> gq, T* out) {
    if (dim == 7) {
        out[0] = gq.x; out[1] = gq.y; out[2] = gq.z; out[3] = gq.w;
        return;
    }
    // q = (v k(th), cos(th/2)),  k = sin(th/2)/th
    T h = T(0.5) * r.th, s = sin(h), c = cos(h);
    T vg = dot(r.v, qvec(gq));
    T dk_over_th, dw_over_th;          // (dk/dth)/th and (d cos(th/2)/dth)/th
    if (r.th > T(1e-6)) {
        dk_over_th = (T(0.5) * c * r.th - s) / (r.th * r.th * r.th);
        dw_over_th = -T(0.5) * s / r.th;
    } else {
        dk_over_th = -T(1) / T(24);
        dw_over_th = -T(0.25);
    }
    T coef = dk_over_th * vg + dw_over_th * gq.w;
    out[0] = r.k * gq.x + r.v.x * coef;
    out[1] = r.k * gq.y + r.v.y * coef;
    out[2] = r.k * gq.z + r.v.z * coef;
}

template <class T> PPR_HD bool se3_isnan(int dim, const T* p, const T* g) {
    T sp = T(0), sg = T(0);
    for (int i = 0; i < dim; ++i) { sp += p[i]; sg += g[i]; }
    return sp != sp || sg != sg;
}

// angle limits equivalent to clamping the cosine to [-1 + eps, 1 - eps]
template <class T> PPR_HD T se3_angle_min(T eps) { return T(2) * asin(sqrt(T(0.5) * eps)); }

template <class T> PPR_HD T se3_pair_loss(int dim, const T* p, const T* g, T ratio, T eps) {
    if (se3_isnan(dim, p, g)) return T(0);
    T dx = p[0] - g[0], dy = p[1] - g[1], dz = p[2] - g[2];
    Se3Quat<T> a = se3_quat(dim, p), b = se3_quat(dim, g);
    Q4<T> r = qmul(a.q, qconj(b.q));
    T nv = sqrt(r.x * r.x + r.y * r.y + r.z * r.z), aw = r.w < T(0) ? -r.w : r.w;
    T ang = T(2) * atan2(nv, aw);
    T lo = se3_angle_min(eps), hi = T(3.14159265358979323846) - lo;
    return dx * dx + dy * dy + dz * dz + ratio * clampT(ang, lo, hi);
}

// adj = dL/dloss; adj_p / adj_g receive `dim` values each (adj_g may be null)
template <class T> PPR_HD void se3_pair_loss_adj(int dim, const T* p, const T* g, T ratio, T eps, T adj, T* adj_p,
                                                 T* adj_g) {
    if (se3_isnan(dim, p, g)) {
        for (int i = 0; i < dim; ++i) { adj_p[i] = T(0); if (adj_g) adj_g[i] = T(0); }
        return;
    }
    PPR_UNROLL for (int i = 0; i < 3; ++i) {
        T d = T(2) * adj * (p[i] - g[i]);
        adj_p[i] = d;
        if (adj_g) adj_g[i] = -d;
    }
    Se3Quat<T> a = se3_quat(dim, p), b = se3_quat(dim, g);
    Q4<T> r = qmul(a.q, qconj(b.q));
    T nv2 = r.x * r.x + r.y * r.y + r.z * r.z, nv = sqrt(nv2), aw = r.w < T(0) ? -r.w : r.w;
    T ang = T(2) * atan2(nv, aw);
    T lo = se3_angle_min(eps), hi = T(3.14159265358979323846) - lo;
    Q4<T> gr = qzero<T>();
    if (ang > lo && ang < hi && nv > T(0)) {
        // d ang / d|v| = 2 |w| / (|v|^2 + w^2),  d ang / d|w| = -2 |v| / (|v|^2 + w^2)
        T n2 = nv2 + aw * aw, ga = adj * ratio;
        T cv = ga * T(2) * aw / (n2 * nv);
        gr = q4<T>(cv * r.x, cv * r.y, cv * r.z, -ga * T(2) * nv / n2 * (r.w < T(0) ? T(-1) : T(1)));
    }
    // r = a (x) conj(b):  g_a = g_r (x) b,   g_conj(b) = conj(a) (x) g_r,  g_b = conj(g_conj(b))
    se3_quat_adj(dim, a, qmul(gr, b.q), adj_p + 3);
    if (adj_g) se3_quat_adj(dim, b, qconj(qmul(qconj(a.q), gr)), adj_g + 3);
}

}  // namespace ppr
