// Per-body / per-contact-point stage functions of one simulation substep and of articulation FK, each with
// its hand-written reverse-mode adjoint.  They restate (from scratch) the arithmetic of
//   eval_body_contacts   /root/reference/diffphys/integrator_euler.py:93-179
//   eval_body_joints     /root/reference/diffphys/integrator_euler.py:289-451 (+ quat_twist :235, quat_decompose :246,
//                        eval_joint_force :262)
//   integrate_bodies     /root/reference/diffphys/integrator_euler.py:21-91
//   eval_articulation_fk warp_lang 0.7.2 (third-party, call sites diffphys/dp_model.py:1068,1204)
// and replace the adjoint kernels Warp's wp.Tape generated for them.  Adjoint functions RECOMPUTE the forward
// intermediates from the saved body state; nothing but (body_q, body_qd) is ever stored per substep.
//
// Conventions: transform (p, q xyzw); twist / wrench (angular, linear); ground plane y = 0, normal +Y.
//
// Rotations by a body's OWN quaternion go through its rotation matrix Rb = qmat(b.r) (built once per substep by the
// caller) and, in the adjoint functions, their d/d(b.r) is accumulated as G += (rank-1) with G = dL/dRb; the caller
// converts once per body and substep with qmat_adj(b.r, G) (ppr_math.h).
#pragma once
#include "ppr_math.h"

namespace ppr {

enum JointType { JT_PRISMATIC = 0, JT_REVOLUTE = 1, JT_BALL = 2, JT_FIXED = 3, JT_FREE = 4, JT_COMPOUND = 5,
                 JT_UNIVERSAL = 6 };

template <class T> struct Body {  // state of one rigid body, also used for its adjoint
    V3<T> x; Q4<T> r; V3<T> w; V3<T> v;
};
template <class T> struct Wrench { V3<T> t; V3<T> f; };

template <class T> PPR_HD Body<T> body_zero() {
    Body<T> b; b.x = vzero<T>(); b.r = qzero<T>(); b.w = vzero<T>(); b.v = vzero<T>(); return b;
}
template <class T> PPR_HD Body<T> body_identity() {
    Body<T> b = body_zero<T>(); b.r.w = T(1); return b;
}
template <class T> PPR_HD Wrench<T> wrench_zero() { Wrench<T> w; w.t = vzero<T>(); w.f = vzero<T>(); return w; }
template <class T> PPR_HD void body_acc(Body<T>& a, const Body<T>& b) { a.x += b.x; a.r += b.r; a.w += b.w; a.v += b.v; }

template <class T> struct JointStatic {
    int type;
    V3<T> xpj; Q4<T> qpj;  // joint_X_p
    Q4<T> qoff;            // rot(joint_X_c)
    V3<T> axis;
};
template <class T> struct JointCtl {  // per-dof quantities of one joint (up to 3 dofs)
    T target[3], act[3], ke[3], kd[3];
    T lo[3], hi[3], lke[3], lkd[3];
};
template <class T> struct ContactMat { T ke, kd, kf, mu; };

// Compile-time feature set of an articulation, so that a kernel instance only carries the code its robot needs
// (the generic adjoint is ~18k SASS instructions and thrashes the instruction cache):
//   JM_REVOLUTE: FREE + REVOLUTE joints only (laikago);  JM_COMPOUND: FREE + COMPOUND only (human, quad);
//   JM_ALL: every supported type.  LIMITS: joint limit springs present.  QOFF: some COMPOUND joint has a
//   non-identity joint_X_c rotation.
enum JointMode { JM_REVOLUTE = 0, JM_COMPOUND = 1, JM_ALL = 2 };
template <int JM> PPR_HD bool jm_has(int type) {
    if (type == JT_FREE) return true;
    if (type == JT_REVOLUTE) return JM != JM_COMPOUND;
    if (type == JT_COMPOUND) return JM != JM_REVOLUTE;
    return JM == JM_ALL;
}

// ------------------------------------------------------------------------------------------ contacts (K3)
// Subtracts the ground-contact wrench of one contact point from F. xc = world COM of the body.
template <class T>
PPR_HD bool contact_point_fwd(const Body<T>& b, const M3<T>& Rb, V3<T> xc, V3<T> p, T dist, ContactMat<T> m,
                              Wrench<T>& F) {
    V3<T> cp = b.x + mrot(Rb, p);
    cp.y -= dist;
    T c = cp.y;
    if (c > T(0)) return false;
    V3<T> rr = cp - xc;
    V3<T> u = b.v + cross(b.w, rr);
    T vn = u.y;
    T fn = c * m.ke;
    T fd = (vn < T(0) ? vn : T(0)) * m.kd * (c < T(0) ? T(1) : T(0));
    T lvt = sqrt(u.x * u.x + u.z * u.z);
    T a = m.kf * lvt, cap = -m.mu * (fn + fd);
    T fm = a < cap ? a : cap;
    T inv = lvt > T(0) ? T(1) / lvt : T(0);
    V3<T> f = v3<T>(u.x * inv * fm, fn + fd, u.z * inv * fm);
    f = clamp3(f, T(500));
    F.t -= cross(rr, f);
    F.f -= f;
    return true;
}

// Reverse of the above: adjF = adjoint of the body's wrench; accumulates into adjB (x, r, w, v) and adj_xc.
template <class T>
PPR_HD void contact_point_adj(const Body<T>& b, const M3<T>& Rb, V3<T> xc, V3<T> p, T dist, ContactMat<T> m,
                              const Wrench<T>& adjF, Body<T>& adjB, M3<T>& G, V3<T>& adj_xc) {
    V3<T> cp = b.x + mrot(Rb, p);
    cp.y -= dist;
    T c = cp.y;
    if (c > T(0)) return;
    V3<T> rr = cp - xc;
    V3<T> u = b.v + cross(b.w, rr);
    T vn = u.y;
    T fn = c * m.ke;
    T stepc = c < T(0) ? T(1) : T(0);
    T fd = (vn < T(0) ? vn : T(0)) * m.kd * stepc;
    T lvt = sqrt(u.x * u.x + u.z * u.z);
    T a = m.kf * lvt, cap = -m.mu * (fn + fd);
    bool a_sel = a < cap;
    T fm = a_sel ? a : cap;
    T inv = lvt > T(0) ? T(1) / lvt : T(0);
    T nx = u.x * inv, nz = u.z * inv;
    V3<T> fr = v3<T>(nx * fm, fn + fd, nz * fm);
    V3<T> f = clamp3(fr, T(500));
    // F.t -= rr x f ; F.f -= f
    V3<T> g_tau = -adjF.t;
    V3<T> g_f = -adjF.f + cross(g_tau, rr);
    V3<T> g_rr = cross(f, g_tau);
    g_f = clamp3_mask(fr, T(500), g_f);
    T g_fnfd = g_f.y;
    // ft = n_vt * fm  (x, z components)
    T g_fm = nx * g_f.x + nz * g_f.z;
    T g_nx = g_f.x * fm, g_nz = g_f.z * fm;
    T g_lvt = T(0);
    if (a_sel) g_lvt = m.kf * g_fm; else g_fnfd += -m.mu * g_fm;
    // n_vt = vt/|vt| ; lvt = |vt|
    T ng = nx * g_nx + nz * g_nz;
    T g_ux = (g_nx - nx * ng) * inv + nx * g_lvt;
    T g_uz = (g_nz - nz * ng) * inv + nz * g_lvt;
    if (!(lvt > T(0))) { g_ux = T(0); g_uz = T(0); }
    T g_c = m.ke * g_fnfd;
    T g_uy = (vn < T(0)) ? m.kd * stepc * g_fnfd : T(0);
    V3<T> g_u = v3<T>(g_ux, g_uy, g_uz);
    // u = v + w x rr
    adjB.v += g_u;
    adjB.w += cross(rr, g_u);
    g_rr += cross(g_u, b.w);
    // rr = cp - xc ; c = cp.y
    V3<T> g_cp = g_rr;
    g_cp.y += g_c;
    adj_xc -= g_rr;
    adjB.x += g_cp;
    m3_acc(G, g_cp, p);
}

// ------------------------------------------------------------------------------------------ joints (K4)
template <class T>
PPR_HD T joint_limit_force(T q, T qd, T lo, T hi, T lke, T lkd) {
    T lim = sel(q < lo, lke * (lo - q) - lkd * sel(qd < T(0), qd, T(0)), T(0));
    lim = sel(q > hi, lke * (hi - q) - lkd * sel(qd > T(0), qd, T(0)), lim);
    return lim;
}
// adjoint of sc = ke(q-target) + kd qd + act - lim, given g = adj_sc
template <class T, bool LIMITS = true>
PPR_HD void joint_scalar_adj(T q, T qd, const JointCtl<T>& c, int k, T g, T& g_q, T& g_qd, T* adj_target, T* adj_act,
                             T* adj_ke, T* adj_kd) {
    g_q += c.ke[k] * g;
    g_qd += c.kd[k] * g;
    adj_target[k] += -c.ke[k] * g;
    adj_act[k] += g;
    adj_ke[k] += (q - c.target[k]) * g;
    adj_kd[k] += qd * g;
    if (!LIMITS) return;
    T g_lim = -g;
    auto above = q > c.hi[k];
    auto below = (q < c.lo[k]) && !above;
    g_q += sel(above || below, -c.lke[k] * g_lim, T(0));
    g_qd += sel((above && (qd > T(0))) || (below && (qd < T(0))), -c.lkd[k] * g_lim, T(0));
}

// Twist angle of r_err about `axis` (quat_twist + acos + sign, integrator_euler.py:235-241,398-400):
//   tw = normalize((d a, w)), d = a . r_err.xyz ;  q = 2 acos(tw.w) sign(a . tw.xyz)
// evaluated in the algebraically identical but well-conditioned form q = 2 atan2(d |a|, w).  The literal acos
// form loses all precision in fp32 near q = 0 (acos(1 - eps) = sqrt(2 eps): 7e-4 rad resolution) and its adjoint
// -1/sqrt(1 - w^2) is singular there; atan2 has neither problem and equals it everywhere else.
template <class T> PPR_HD T revolute_angle(V3<T> axis, Q4<T> r_err) {
    T la = sqrt(dot(axis, axis));
    T y = dot(axis, qvec(r_err)) * la;
    return T(2) * atan2(y, r_err.w);
}
template <class T> PPR_HD void revolute_angle_adj(V3<T> axis, Q4<T> r_err, T g_q, Q4<T>& g_rerr) {
    T la = sqrt(dot(axis, axis));
    T y = dot(axis, qvec(r_err)) * la;
    T den = y * y + r_err.w * r_err.w;
    T iden = sel(den > T(0), T(1) / den, T(0));   // zero adjoint at the singular point
    T g_y = T(2) * g_q * r_err.w * iden;
    g_rerr.w += -T(2) * g_q * y * iden;
    T g_d = g_y * la;
    g_rerr.x += g_d * axis.x; g_rerr.y += g_d * axis.y; g_rerr.z += g_d * axis.z;
}

// COMPOUND joint: XYZ-Euler decomposition of the relative rotation q_pc and the reconstructed rotation axes
// (quat_decompose + axis reconstruction, integrator_euler.py:246-258,419-427):
//   (a0,a1,a2) = -(atan2(c2.y,c2.z), asin(-c2.x), atan2(c1.x,c0.x)),  c_i = quat_rotate(q_pc, e_i)
//   e0 = x,  e1 = R(q0) y,  e2 = R(q1 q0) z   with q0 = (x, a0), q1 = (e1, a1).
// The reference rebuilds e1, e2 through sin/cos of the angles it just extracted with atan2/asin; the same functions
// in closed form are  e1 = (0, cos a0, sin a0),  e2 = (sin a1, -sin a0 cos a1, cos a0 cos a1)  with
//   cos a0 = c2.z/rho, sin a0 = -c2.y/rho, rho = |(c2.y,c2.z)|,  sin a1 = clamp(c2.x), cos a1 = sqrt(1 - sin^2 a1)
// (q0, q1 are exactly unit, so R(q1 q0) = R(q1) R(q0)) -- no trigonometry besides the three inverse functions
// that the PD law itself needs.
template <class T> struct CompoundDec {
    T c0x, c1x; V3<T> c2;
    T ang[3];
    T sa0, ca0, sa1, ca1;
    V3<T> e1, e2;
};
// `known` (optional): the three angles as computed by an earlier call on the same q_pc (the forward kernel stores them
// in the checkpoint so that the adjoint kernel does not repeat the two atan2 and the asin).
template <class T> PPR_HD CompoundDec<T> compound_decompose(Q4<T> q_pc, const T* known = nullptr) {
    CompoundDec<T> d;
    V3<T> c0 = qrot(q_pc, v3<T>(T(1), T(0), T(0)));
    V3<T> c1 = qrot(q_pc, v3<T>(T(0), T(1), T(0)));
    d.c2 = qrot(q_pc, v3<T>(T(0), T(0), T(1)));
    d.c0x = c0.x; d.c1x = c1.x;
    if (known) {
        d.ang[0] = known[0]; d.ang[1] = known[1]; d.ang[2] = known[2];
    } else {
        d.ang[0] = -atan2(d.c2.y, d.c2.z);
        d.ang[1] = -safe_asin(-d.c2.x);
        d.ang[2] = -atan2(d.c1x, d.c0x);
    }
    T rho = sqrt(d.c2.y * d.c2.y + d.c2.z * d.c2.z);
    T ir = T(1) / rho;
    d.ca0 = sel(rho > T(0), d.c2.z * ir, T(1));   // atan2(0,0) = 0
    d.sa0 = sel(rho > T(0), -d.c2.y * ir, T(0));
    d.sa1 = clampT(d.c2.x, T(-1), T(1));
    d.ca1 = sqrt(T(1) - d.sa1 * d.sa1);
    d.e1 = v3<T>(T(0), d.ca0, d.sa0);
    d.e2 = v3<T>(d.sa1, -d.sa0 * d.ca1, d.ca0 * d.ca1);
    return d;
}
// adjoint: g_e1, g_e2 = adjoints of the axes, g_ang = adjoints of the angles (in/out: axis terms are added);
// returns the adjoint of q_pc
template <class T> PPR_HD Q4<T> compound_decompose_adj(Q4<T> q_pc, const CompoundDec<T>& d, V3<T> g_e1, V3<T> g_e2,
                                                       T* g_ang) {
    g_ang[0] += -d.sa0 * g_e1.y + d.ca0 * g_e1.z - d.ca0 * d.ca1 * g_e2.y - d.sa0 * d.ca1 * g_e2.z;
    g_ang[1] += d.ca1 * g_e2.x + d.sa0 * d.sa1 * g_e2.y - d.ca0 * d.sa1 * g_e2.z;
    T g_phi = -g_ang[0], g_theta = -g_ang[1], g_psi = -g_ang[2];
    V3<T> g_c0 = vzero<T>(), g_c1 = vzero<T>(), g_c2 = vzero<T>();
    T den = d.c2.y * d.c2.y + d.c2.z * d.c2.z;
    { T id = sel(den > T(0), T(1) / den, T(0)); g_c2.y += g_phi * d.c2.z * id; g_c2.z -= g_phi * d.c2.y * id; }
    g_c2.x += -g_theta * safe_asin_adj(-d.c2.x);
    den = d.c1x * d.c1x + d.c0x * d.c0x;
    { T id = sel(den > T(0), T(1) / den, T(0)); g_c1.x += g_psi * d.c0x * id; g_c0.x -= g_psi * d.c1x * id; }
    return qrot_adj_q(q_pc, v3<T>(T(1), T(0), T(0)), g_c0) + qrot_adj_q(q_pc, v3<T>(T(0), T(1), T(0)), g_c1) +
           qrot_adj_q(q_pc, v3<T>(T(0), T(0), T(1)), g_c2);
}

// Forward joint wrench. P = parent body (identity / zero twist if the joint has no parent), xcp / xcc = world COMs.
// Outputs the joint torque t and force f together with the two moment arms; the caller applies
//   F_parent += (t + arm_p x f, f),  F_child -= (t + arm_c x f, f)      (integrator_euler.py:448-451)
template <class T, int JM = JM_ALL, bool LIMITS = true, bool QOFF = true>
PPR_HD void joint_fwd(const JointStatic<T>& js, const JointCtl<T>& c, T ake, T akd, const Body<T>& P, V3<T> xcp,
                      bool has_parent, const Body<T>& C, const M3<T>& Rc, V3<T> xcc, V3<T>& t_out, V3<T>& f_out,
                      V3<T>& arm_p, V3<T>& arm_c, T* ang_out = nullptr) {
    t_out = vzero<T>(); f_out = vzero<T>();
    if (ang_out) { ang_out[0] = T(0); ang_out[1] = T(0); ang_out[2] = T(0); }
    V3<T> xA = P.x + qrot(P.r, js.xpj);
    Q4<T> qA = qmul(P.r, js.qpj);
    arm_p = has_parent ? xA - xcp : vzero<T>();
    arm_c = C.x - xcc;
    if (js.type == JT_FREE) return;
    V3<T> x_err = C.x - xA;
    Q4<T> r_err = qmul(qconj(qA), C.r);
    V3<T> v_err = C.v - P.v, w_err = C.w - P.w;
    const T ads = T(0.01);
    if (JM != JM_COMPOUND && js.type == JT_REVOLUTE) {
        V3<T> axis_p = qrot(qA, js.axis), axis_c = mrot(Rc, js.axis);
        T q = revolute_angle(js.axis, r_err);
        if (ang_out) ang_out[0] = q;
        T qd = dot(w_err, axis_p);
        T sc = c.ke[0] * (q - c.target[0]) + c.kd[0] * qd + c.act[0] -
               (LIMITS ? joint_limit_force(q, qd, c.lo[0], c.hi[0], c.lke[0], c.lkd[0]) : T(0));
        t_out = axis_p * sc + cross(axis_p, axis_c) * ake + (w_err - axis_p * qd) * (akd * ads);
        f_out = x_err * ake + v_err * akd;
    } else if (JM != JM_REVOLUTE && js.type == JT_COMPOUND) {
        Q4<T> q_pc = QOFF ? qmul(qmul(qconj(js.qoff), r_err), js.qoff) : r_err;
        CompoundDec<T> dec = compound_decompose(q_pc);
        const T* ang = dec.ang;
        if (ang_out) { ang_out[0] = ang[0]; ang_out[1] = ang[1]; ang_out[2] = ang[2]; }
        Q4<T> qw = QOFF ? qmul(qA, js.qoff) : qA;
        V3<T> ax[3] = {v3<T>(T(1), T(0), T(0)), dec.e1, dec.e2};
        const M3<T> Mw = qmat(qw);
        V3<T> t = vzero<T>();
PPR_UNROLL
        for (int k = 0; k < 3; ++k) {
            V3<T> aw = mrot(Mw, ax[k]);
            T qd = dot(aw, w_err);
            T sc = c.ke[k] * (ang[k] - c.target[k]) + c.kd[k] * qd + c.act[k] -
                   (LIMITS ? joint_limit_force(ang[k], qd, c.lo[k], c.hi[k], c.lke[k], c.lkd[k]) : T(0));
            t += aw * sc;
        }
        t_out = clamp3(t, T(1e4));
        f_out = clamp3(x_err * ake + v_err * akd, T(1e4));
    } else if (JM == JM_ALL && js.type == JT_FIXED) {
        // acos(r_err.w) evaluated as atan2(|r_err.xyz|, r_err.w): same value for a unit quaternion, well
        // conditioned near the identity (see revolute_angle)
        V3<T> e = qvec(r_err);
        T l = sqrt(dot(e, e));
        T inv = sel(l > T(0), T(1) / l, T(0));
        V3<T> ang_err = e * (inv * atan2(l, r_err.w) * T(2));
        f_out = x_err * ake + v_err * akd;
        t_out = qrot(qA, ang_err) * ake + w_err * (akd * ads);
    }
}

// Reverse of joint_fwd + the wrench scatter.  adjFp / adjFc = adjoints of the parent's / child's total wrench.
// Accumulates into adjP / adj_xcp (parent state, parent world-COM), adjC / adj_xcc and the per-dof parameter adjoints.
template <class T, int JM = JM_ALL, bool LIMITS = true, bool QOFF = true>
PPR_HD void joint_adj(const JointStatic<T>& js, const JointCtl<T>& c, T ake, T akd, const Body<T>& P, V3<T> xcp,
                      bool has_parent, const Body<T>& C, const M3<T>& Rc, V3<T> xcc, const Wrench<T>& adjFp,
                      const Wrench<T>& adjFc, Body<T>& adjP, V3<T>& adj_xcp, Body<T>& adjC, M3<T>& Gc, V3<T>& adj_xcc,
                      T* adj_target, T* adj_act, T* adj_ke, T* adj_kd, const T* ang_in = nullptr) {
    if (js.type == JT_FREE) return;
    // ---- recompute forward
    V3<T> xA = P.x + qrot(P.r, js.xpj);
    Q4<T> qA = qmul(P.r, js.qpj);
    V3<T> arm_p = has_parent ? xA - xcp : vzero<T>();
    V3<T> arm_c = C.x - xcc;
    V3<T> x_err = C.x - xA;
    Q4<T> r_err = qmul(qconj(qA), C.r);
    V3<T> v_err = C.v - P.v, w_err = C.w - P.w;
    const T ads = T(0.01);
    V3<T> f_out;
    // adjoints of the intermediate quantities
    V3<T> g_xerr = vzero<T>(), g_verr = vzero<T>(), g_werr = vzero<T>();
    Q4<T> g_rerr = qzero<T>(), g_qA = qzero<T>(), g_Cr = qzero<T>();

    // ---- scatter adjoint: gt = adj t, gf = adj f, arms
    //   Fp.t += t + arm_p x f ; Fp.f += f ; Fc.t -= t + arm_c x f ; Fc.f -= f
    V3<T> gtp = has_parent ? adjFp.t : vzero<T>();
    V3<T> gfp = has_parent ? adjFp.f : vzero<T>();
    V3<T> gt = gtp - adjFc.t;
    // forward values of t, f are needed for the arm adjoints and clamp masks -> computed per type below

    if (JM != JM_COMPOUND && js.type == JT_REVOLUTE) {
        V3<T> axis_p = qrot(qA, js.axis), axis_c = mrot(Rc, js.axis);
        T q = ang_in ? ang_in[0] : revolute_angle(js.axis, r_err);
        T qd = dot(w_err, axis_p);
        T sc = c.ke[0] * (q - c.target[0]) + c.kd[0] * qd + c.act[0] -
               (LIMITS ? joint_limit_force(q, qd, c.lo[0], c.hi[0], c.lke[0], c.lkd[0]) : T(0));
        f_out = x_err * ake + v_err * akd;
        V3<T> gf = gfp - adjFc.f + cross(gtp, arm_p) - cross(adjFc.t, arm_c);
        V3<T> g_armp = cross(f_out, gtp), g_armc = -cross(f_out, adjFc.t);
        g_xerr += gf * ake;
        g_verr += gf * akd;
        T c2 = akd * ads;
        T g_sc = dot(gt, axis_p);
        V3<T> g_axp = gt * (sc - c2 * qd) + cross(axis_c, gt) * ake;
        V3<T> g_axc = cross(gt, axis_p) * ake;
        g_werr += gt * c2;
        T g_qd = -c2 * dot(gt, axis_p);
        T g_q = T(0);
        joint_scalar_adj<T, LIMITS>(q, qd, c, 0, g_sc, g_q, g_qd, adj_target, adj_act, adj_ke, adj_kd);
        g_werr += axis_p * g_qd;
        g_axp += w_err * g_qd;
        revolute_angle_adj(js.axis, r_err, g_q, g_rerr);
        g_qA += qrot_adj_q(qA, js.axis, g_axp);
        m3_acc(Gc, g_axc, js.axis);
        // arms
        adjC.x += g_armc; adj_xcc -= g_armc;
        if (has_parent) { adj_xcp -= g_armp; }
        V3<T> g_xA = g_armp - g_xerr;
        adjC.x += g_xerr;
        adjC.v += g_verr; adjC.w += g_werr;
        adjP.v -= g_verr; adjP.w -= g_werr;
        // r_err = conj(qA) * C.r
        g_Cr += qmul(qA, g_rerr);
        g_qA += qconj(qmul(g_rerr, qconj(C.r)));
        adjC.r += g_Cr;
        adjP.x += g_xA;
        adjP.r += qrot_adj_q(P.r, js.xpj, g_xA);
        adjP.r += qmul(g_qA, qconj(js.qpj));
        return;
    }
    if (JM != JM_REVOLUTE && js.type == JT_COMPOUND) {
        Q4<T> q_pc = QOFF ? qmul(qmul(qconj(js.qoff), r_err), js.qoff) : r_err;
        const V3<T> ex = v3<T>(T(1), T(0), T(0));
        CompoundDec<T> dec = compound_decompose(q_pc, ang_in);
        const T* ang = dec.ang;
        Q4<T> qw = QOFF ? qmul(qA, js.qoff) : qA;
        V3<T> ax[3] = {ex, dec.e1, dec.e2};
        V3<T> aw[3];
        T qd[3], sc[3];
        V3<T> traw = vzero<T>();
        const M3<T> Mw = qmat(qw);
PPR_UNROLL
        for (int k = 0; k < 3; ++k) {
            aw[k] = mrot(Mw, ax[k]);
            qd[k] = dot(aw[k], w_err);
            sc[k] = c.ke[k] * (ang[k] - c.target[k]) + c.kd[k] * qd[k] + c.act[k] -
                    (LIMITS ? joint_limit_force(ang[k], qd[k], c.lo[k], c.hi[k], c.lke[k], c.lkd[k]) : T(0));
            traw += aw[k] * sc[k];
        }
        V3<T> fraw = x_err * ake + v_err * akd;
        f_out = clamp3(fraw, T(1e4));
        V3<T> gf = gfp - adjFc.f + cross(gtp, arm_p) - cross(adjFc.t, arm_c);
        V3<T> g_armp = cross(f_out, gtp), g_armc = -cross(f_out, adjFc.t);
        gf = clamp3_mask(fraw, T(1e4), gf);
        g_xerr += gf * ake;
        g_verr += gf * akd;
        V3<T> gtr = clamp3_mask(traw, T(1e4), gt);
        T g_ang[3] = {T(0), T(0), T(0)};
        V3<T> g_ax[3] = {vzero<T>(), vzero<T>(), vzero<T>()};
        M3<T> Gw = m3_zero<T>();
PPR_UNROLL
        for (int k = 0; k < 3; ++k) {
            T g_sc = dot(gtr, aw[k]);
            V3<T> g_aw = gtr * sc[k];
            T g_qd = T(0);
            joint_scalar_adj<T, LIMITS>(ang[k], qd[k], c, k, g_sc, g_ang[k], g_qd, adj_target, adj_act, adj_ke, adj_kd);
            g_aw += w_err * g_qd;
            g_werr += aw[k] * g_qd;
            m3_acc(Gw, g_aw, ax[k]);
            g_ax[k] += mrot_t(Mw, g_aw);
        }
        Q4<T> g_qw = qmat_adj(qw, Gw);
        Q4<T> g_qpc = compound_decompose_adj(q_pc, dec, g_ax[1], g_ax[2], g_ang);
        // q_pc = (conj(qoff) * r_err) * qoff ; qw = qA * qoff
        if (QOFF) {
            g_rerr += qmul(js.qoff, qmul(g_qpc, qconj(js.qoff)));
            g_qA += qmul(g_qw, qconj(js.qoff));
        } else {
            g_rerr += g_qpc;
            g_qA += g_qw;
        }
        adjC.x += g_armc; adj_xcc -= g_armc;
        if (has_parent) { adj_xcp -= g_armp; }
        V3<T> g_xA = g_armp - g_xerr;
        adjC.x += g_xerr;
        adjC.v += g_verr; adjC.w += g_werr;
        adjP.v -= g_verr; adjP.w -= g_werr;
        g_Cr += qmul(qA, g_rerr);
        g_qA += qconj(qmul(g_rerr, qconj(C.r)));
        adjC.r += g_Cr;
        adjP.x += g_xA;
        adjP.r += qrot_adj_q(P.r, js.xpj, g_xA);
        adjP.r += qmul(g_qA, qconj(js.qpj));
        return;
    }
    if (JM == JM_ALL && js.type == JT_FIXED) {
        V3<T> e = qvec(r_err);
        T l = sqrt(dot(e, e));
        T inv = sel(l > T(0), T(1) / l, T(0));
        T ac = atan2(l, r_err.w) * T(2);
        V3<T> nrm = e * inv;
        V3<T> ang_err = nrm * ac;
        f_out = x_err * ake + v_err * akd;
        V3<T> gf = gfp - adjFc.f + cross(gtp, arm_p) - cross(adjFc.t, arm_c);
        V3<T> g_armp = cross(f_out, gtp), g_armc = -cross(f_out, adjFc.t);
        g_xerr += gf * ake;
        g_verr += gf * akd;
        g_werr += gt * (akd * ads);
        V3<T> g_rot = gt * ake;  // adjoint of qrot(qA, ang_err)
        g_qA += qrot_adj_q(qA, ang_err, g_rot);
        V3<T> g_ang = qrot_inv(qA, g_rot);
        T g_ac = dot(nrm, g_ang);
        V3<T> g_n = g_ang * ac;
        T den = l * l + r_err.w * r_err.w;
        T iden = sel(den > T(0), T(1) / den, T(0));
        {   // l = 0: inv = 0 and nrm = 0, every term below vanishes (zero adjoint at the singular point)
            T ng = dot(nrm, g_n);
            T g_l = T(2) * g_ac * r_err.w * iden;  // d(2 atan2(l, w))/dl
            g_rerr.x += (g_n.x - nrm.x * ng) * inv + nrm.x * g_l; g_rerr.y += (g_n.y - nrm.y * ng) * inv + nrm.y * g_l;
            g_rerr.z += (g_n.z - nrm.z * ng) * inv + nrm.z * g_l;
        }
        g_rerr.w += -T(2) * g_ac * l * iden;
        adjC.x += g_armc; adj_xcc -= g_armc;
        if (has_parent) { adj_xcp -= g_armp; }
        V3<T> g_xA = g_armp - g_xerr;
        adjC.x += g_xerr;
        adjC.v += g_verr; adjC.w += g_werr;
        adjP.v -= g_verr; adjP.w -= g_werr;
        g_Cr += qmul(qA, g_rerr);
        g_qA += qconj(qmul(g_rerr, qconj(C.r)));
        adjC.r += g_Cr;
        adjP.x += g_xA;
        adjP.r += qrot_adj_q(P.r, js.xpj, g_xA);
        adjP.r += qmul(g_qA, qconj(js.qpj));
        return;
    }
}

// ------------------------------------------------------------------------------------------ integrate (K5)
// The part of K5 that does not depend on the wrench (body-frame angular velocity and the gyroscopic term): the CUDA
// forward kernel evaluates it while it waits for the wrenches of the body's children.
template <class T> struct IntegratePre { V3<T> wb, gyro; };
template <class T> PPR_HD IntegratePre<T> integrate_pre(const Body<T>& b, const M3<T>& Rb, const T* I) {
    IntegratePre<T> p;
    p.wb = mrot_t(Rb, b.w);
    p.gyro = cross(p.wb, matvec(I, p.wb));
    return p;
}
template <class T>
PPR_HD Body<T> integrate_post(const Body<T>& b, const M3<T>& Rb, V3<T> xc, V3<T> com, const Wrench<T>& F, T inv_m,
                              const T* inv_I, V3<T> g, T dt, const IntegratePre<T>& pre) {
    T nz = sel(inv_m != T(0), T(1), T(0));
    V3<T> v1 = b.v + (F.f * inv_m + g * nz) * dt;
    V3<T> x1c = xc + v1 * dt;
    V3<T> wb = pre.wb;
    V3<T> tb = mrot_t(Rb, F.t) - pre.gyro;
    V3<T> w1 = mrot(Rb, wb + matvec(inv_I, tb) * dt);
    Q4<T> rq = b.r + qmul(q4<T>(w1.x, w1.y, w1.z, T(0)), b.r) * (T(0.5) * dt);
    T len;
    Q4<T> r1 = qnormalize(rq, len);
    Body<T> o;
    o.w = clamp3(w1 * (T(1) - T(0.1) * dt), T(10));
    o.v = clamp3(v1, T(10));
    o.r = r1;
    o.x = x1c - qrot(r1, com);
    return o;
}
template <class T>
PPR_HD Body<T> integrate_fwd(const Body<T>& b, const M3<T>& Rb, V3<T> xc, V3<T> com, const Wrench<T>& F, T inv_m,
                             const T* I, const T* inv_I, V3<T> g, T dt) {
    return integrate_post(b, Rb, xc, com, F, inv_m, inv_I, g, dt, integrate_pre(b, Rb, I));
}

// Core of K5^T. The inertia-parameter adjoints are rank-1 updates; they are returned as their factors
//   adj_I += gI_a (x) gI_b ,  adj_inv_I += giI_a (x) giI_b      (giI_a already carries the factor dt)
// so that a caller can accumulate them wherever it likes (the CUDA adjoint keeps the accumulators in shared memory).
template <class T>
PPR_HD void integrate_adj_core(const Body<T>& b, const M3<T>& Rb, V3<T> xc, V3<T> com, const Wrench<T>& F, T inv_m,
                               const T* I, const T* inv_I, V3<T> g, T dt, const Body<T>& adjO, Body<T>& adjB, M3<T>& G,
                               V3<T>& adj_xc, Wrench<T>& adjF, T& adj_inv_m, V3<T>& gI_a, V3<T>& gI_b, V3<T>& giI_a,
                               V3<T>& giI_b) {
    // ---- recompute
    T nz = sel(inv_m != T(0), T(1), T(0));
    V3<T> v1 = b.v + (F.f * inv_m + g * nz) * dt;
    V3<T> wb = mrot_t(Rb, b.w);
    V3<T> Iwb = matvec(I, wb);
    V3<T> tb = mrot_t(Rb, F.t) - cross(wb, Iwb);
    V3<T> wb2 = wb + matvec(inv_I, tb) * dt;
    V3<T> w1 = mrot(Rb, wb2);
    Q4<T> wq = q4<T>(w1.x, w1.y, w1.z, T(0));
    Q4<T> rq = b.r + qmul(wq, b.r) * (T(0.5) * dt);
    T len;
    Q4<T> r1 = qnormalize(rq, len);
    T damp = T(1) - T(0.1) * dt;
    // ---- reverse
    V3<T> g_x1c = adjO.x;
    Q4<T> g_r1 = adjO.r + qrot_adj_q(r1, com, -adjO.x);
    V3<T> g_v1 = clamp3_mask(v1, T(10), adjO.v) + g_x1c * dt;
    V3<T> g_w1 = clamp3_mask(w1 * damp, T(10), adjO.w) * damp;
    Q4<T> g_rq = qnormalize_adj(r1, len, g_r1);
    T hdt = T(0.5) * dt;
    adjB.r += g_rq + qmul(qconj(wq), g_rq) * hdt;
    Q4<T> g_wq = qmul(g_rq, qconj(b.r)) * hdt;
    g_w1 += qvec(g_wq);
    m3_acc(G, g_w1, wb2);
    V3<T> g_wb2 = mrot_t(Rb, g_w1);
    V3<T> g_wb = g_wb2;
    V3<T> g_tb = matTvec(inv_I, g_wb2) * dt;
    giI_a = g_wb2 * dt; giI_b = tb;
    adjF.t = mrot(Rb, g_tb);
    m3_acc(G, F.t, g_tb);
    V3<T> g_c = -g_tb;  // c = wb x Iwb
    g_wb += cross(Iwb, g_c);
    V3<T> g_Iwb = cross(g_c, wb);
    g_wb += matTvec(I, g_Iwb);
    gI_a = g_Iwb; gI_b = wb;
    adjB.w += mrot(Rb, g_wb);
    m3_acc(G, b.w, g_wb);
    adj_xc += g_x1c;
    adjB.v += g_v1;
    V3<T> g_a = g_v1 * dt;
    adjF.f = g_a * inv_m;
    adj_inv_m += dot(F.f, g_a);
}

template <class T>
PPR_HD void integrate_adj(const Body<T>& b, const M3<T>& Rb, V3<T> xc, V3<T> com, const Wrench<T>& F, T inv_m,
                          const T* I, const T* inv_I, V3<T> g, T dt, const Body<T>& adjO, Body<T>& adjB, M3<T>& G,
                          V3<T>& adj_xc, Wrench<T>& adjF, T& adj_inv_m, T* adj_I, T* adj_inv_I) {
    V3<T> a, bb, c, d;
    integrate_adj_core(b, Rb, xc, com, F, inv_m, I, inv_I, g, dt, adjO, adjB, G, adj_xc, adjF, adj_inv_m, a, bb, c, d);
    outer_acc(adj_I, a, bb, T(1));
    outer_acc(adj_inv_I, c, d, T(1));
}

// ------------------------------------------------------------------------------------------ FK (K1)
// P = parent world pose / twist (identity, zero when the joint has no parent). jq / jqd point at this joint's
// coordinates / dofs.
template <class T, int JM = JM_ALL>
PPR_HD Body<T> fk_joint_fwd(const JointStatic<T>& js, V3<T> com, const Body<T>& P, const T* jq, const T* jqd) {
    V3<T> x_wj = P.x + qrot(P.r, js.xpj);
    Q4<T> r_wj = qmul(P.r, js.qpj);
    V3<T> x_jc = vzero<T>(), w_j = vzero<T>(), v_j = vzero<T>();
    Q4<T> r_jc = q4<T>(T(0), T(0), T(0), T(1));
    if (js.type == JT_FREE) {
        x_jc = v3<T>(jq[0], jq[1], jq[2]);
        r_jc = q4<T>(jq[3], jq[4], jq[5], jq[6]);
        w_j = v3<T>(jqd[0], jqd[1], jqd[2]);
        v_j = v3<T>(jqd[3], jqd[4], jqd[5]);
    } else if (JM != JM_COMPOUND && js.type == JT_REVOLUTE) {
        r_jc = q_axis_angle(js.axis, jq[0]);
        w_j = js.axis * jqd[0];
    } else if (JM != JM_REVOLUTE && js.type == JT_COMPOUND) {
        const V3<T> ex = v3<T>(T(1), T(0), T(0)), ey = v3<T>(T(0), T(1), T(0)), ez = v3<T>(T(0), T(0), T(1));
        Q4<T> q0 = q_axis_angle(ex, jq[0]);
        V3<T> a1 = qrot(q0, ey);
        Q4<T> q1 = q_axis_angle(a1, jq[1]);
        Q4<T> q10 = qmul(q1, q0);
        V3<T> a2 = qrot(q10, ez);
        Q4<T> q2 = q_axis_angle(a2, jq[2]);
        r_jc = qmul(q2, q10);
        w_j = ex * jqd[0] + a1 * jqd[1] + a2 * jqd[2];
    }
    Body<T> o;
    o.x = x_wj + qrot(r_wj, x_jc);
    o.r = qmul(r_wj, r_jc);
    V3<T> w_w = qrot(r_wj, w_j), v_w = qrot(r_wj, v_j);
    o.w = P.w + w_w;
    o.v = P.v + v_w + cross(w_w, com);
    return o;
}

template <class T, int JM = JM_ALL>
PPR_HD void fk_joint_adj(const JointStatic<T>& js, V3<T> com, const Body<T>& P, const T* jq, const T* jqd,
                         const Body<T>& adjO, Body<T>& adjP, T* adj_jq, T* adj_jqd) {
    Q4<T> r_wj = qmul(P.r, js.qpj);
    adjP.w += adjO.w;
    adjP.v += adjO.v;
    V3<T> g_vw = adjO.v;
    V3<T> g_ww = adjO.w + cross(com, adjO.v);
    V3<T> g_wj = qrot_inv(r_wj, g_ww), g_vj = qrot_inv(r_wj, g_vw);
    V3<T> g_xjc = qrot_inv(r_wj, adjO.x);
    Q4<T> g_rwj = qzero<T>();
    if (js.type == JT_FREE) {
        V3<T> x_jc = v3<T>(jq[0], jq[1], jq[2]);
        Q4<T> r_jc = q4<T>(jq[3], jq[4], jq[5], jq[6]);
        V3<T> w_j = v3<T>(jqd[0], jqd[1], jqd[2]), v_j = v3<T>(jqd[3], jqd[4], jqd[5]);
        g_rwj += qrot_adj_q(r_wj, w_j, g_ww) + qrot_adj_q(r_wj, v_j, g_vw) + qrot_adj_q(r_wj, x_jc, adjO.x);
        g_rwj += qmul(adjO.r, qconj(r_jc));
        Q4<T> g_rjc = qmul(qconj(r_wj), adjO.r);
        adj_jq[0] += g_xjc.x; adj_jq[1] += g_xjc.y; adj_jq[2] += g_xjc.z;
        adj_jq[3] += g_rjc.x; adj_jq[4] += g_rjc.y; adj_jq[5] += g_rjc.z; adj_jq[6] += g_rjc.w;
        adj_jqd[0] += g_wj.x; adj_jqd[1] += g_wj.y; adj_jqd[2] += g_wj.z;
        adj_jqd[3] += g_vj.x; adj_jqd[4] += g_vj.y; adj_jqd[5] += g_vj.z;
    } else if (JM != JM_COMPOUND && js.type == JT_REVOLUTE) {
        Q4<T> r_jc = q_axis_angle(js.axis, jq[0]);
        V3<T> w_j = js.axis * jqd[0];
        g_rwj += qrot_adj_q(r_wj, w_j, g_ww);
        g_rwj += qmul(adjO.r, qconj(r_jc));
        Q4<T> g_rjc = qmul(qconj(r_wj), adjO.r);
        V3<T> dummy = vzero<T>();
        adj_jq[0] += q_axis_angle_adj(js.axis, jq[0], g_rjc, dummy);
        adj_jqd[0] += dot(js.axis, g_wj);
    } else if (JM != JM_REVOLUTE && js.type == JT_COMPOUND) {
        const V3<T> ex = v3<T>(T(1), T(0), T(0)), ey = v3<T>(T(0), T(1), T(0)), ez = v3<T>(T(0), T(0), T(1));
        Q4<T> q0 = q_axis_angle(ex, jq[0]);
        V3<T> a1 = qrot(q0, ey);
        Q4<T> q1 = q_axis_angle(a1, jq[1]);
        Q4<T> q10 = qmul(q1, q0);
        V3<T> a2 = qrot(q10, ez);
        Q4<T> q2 = q_axis_angle(a2, jq[2]);
        Q4<T> r_jc = qmul(q2, q10);
        V3<T> w_j = ex * jqd[0] + a1 * jqd[1] + a2 * jqd[2];
        g_rwj += qrot_adj_q(r_wj, w_j, g_ww);
        g_rwj += qmul(adjO.r, qconj(r_jc));
        Q4<T> g_rjc = qmul(qconj(r_wj), adjO.r);
        adj_jqd[0] += dot(ex, g_wj); adj_jqd[1] += dot(a1, g_wj); adj_jqd[2] += dot(a2, g_wj);
        V3<T> g_a1 = g_wj * jqd[1], g_a2 = g_wj * jqd[2];
        Q4<T> g_q2 = qmul(g_rjc, qconj(q10));
        Q4<T> g_q10 = qmul(qconj(q2), g_rjc);
        adj_jq[2] += q_axis_angle_adj(a2, jq[2], g_q2, g_a2);
        g_q10 += qrot_adj_q(q10, ez, g_a2);
        Q4<T> g_q1 = qmul(g_q10, qconj(q0));
        Q4<T> g_q0 = qmul(qconj(q1), g_q10);
        adj_jq[1] += q_axis_angle_adj(a1, jq[1], g_q1, g_a1);
        g_q0 += qrot_adj_q(q0, ey, g_a1);
        V3<T> dummy = vzero<T>();
        adj_jq[0] += q_axis_angle_adj(ex, jq[0], g_q0, dummy);
    } else {  // FIXED: r_jc = identity
        g_rwj += adjO.r;
    }
    // X_wj = X_wp * X_pj
    adjP.x += adjO.x;
    adjP.r += qrot_adj_q(P.r, js.xpj, adjO.x);
    adjP.r += qmul(g_rwj, qconj(js.qpj));
}

}  // namespace ppr
