// Batch-input producer glue of the imitation loop as one function per time sample:
//   target  = rotate_frame(global_q, q)        T_global @ T_mocap            (dp_utils.py:60-72, dp_model.py:631-640)
//   queried = compose_delta(target, delta)     T_delta  @ target, delta = (translation, axis-angle)
//                                                                            (dp_utils.py:21-30, dp_model.py:650-655)
// and its adjoint w.r.t. global_q (per sample; the caller sums over samples) and delta.  The mocap pose q carries no
// gradient.  Conventions of the composed torch code it replaces (ppr_diffphys_b200/imitation.py): positions are
// rotated by the rotation of the NORMALISED quaternion (quaternion_to_matrix), quaternions are multiplied raw.
// ~600 elementwise torch kernels per iteration (forward + backward) become two launches.
#pragma once
#include "ppr_loss.h"

namespace ppr {

template <class T>
PPR_HD void frame_compose(const T* gq, const T* q, const T* d, T* target, T* queried) {
    Q4<T> g = q4<T>(gq[3], gq[4], gq[5], gq[6]);
    T gl;
    M3<T> Rg = qmat(qnormalize(g, gl));
    V3<T> p = mrot(Rg, v3<T>(q[0], q[1], q[2])) + v3<T>(gq[0], gq[1], gq[2]);
    Q4<T> r = qmul(g, q4<T>(q[3], q[4], q[5], q[6]));
    target[0] = p.x; target[1] = p.y; target[2] = p.z;
    target[3] = r.x; target[4] = r.y; target[5] = r.z; target[6] = r.w;
    Se3Quat<T> dq = se3_quat(6, d);
    T dl;
    M3<T> Rd = qmat(qnormalize(dq.q, dl));
    V3<T> p2 = mrot(Rd, p) + v3<T>(d[0], d[1], d[2]);
    Q4<T> r2 = qmul(dq.q, r);
    queried[0] = p2.x; queried[1] = p2.y; queried[2] = p2.z;
    queried[3] = r2.x; queried[4] = r2.y; queried[5] = r2.z; queried[6] = r2.w;
}

// at / aq = adjoints of target / queried (7 each); adj_g (7) and adj_d (6) are overwritten
template <class T>
PPR_HD void frame_compose_adj(const T* gq, const T* q, const T* d, const T* at, const T* aq, T* adj_g, T* adj_d) {
    Q4<T> g = q4<T>(gq[3], gq[4], gq[5], gq[6]);
    T gl, dl;
    Q4<T> gu = qnormalize(g, gl);
    M3<T> Rg = qmat(gu);
    V3<T> qx = v3<T>(q[0], q[1], q[2]);
    Q4<T> qq = q4<T>(q[3], q[4], q[5], q[6]);
    V3<T> p = mrot(Rg, qx) + v3<T>(gq[0], gq[1], gq[2]);
    Q4<T> r = qmul(g, qq);
    Se3Quat<T> dq = se3_quat(6, d);
    Q4<T> du = qnormalize(dq.q, dl);
    M3<T> Rd = qmat(du);
    // queried = (Rd p + d.xyz, dq (x) r)
    V3<T> g_p2 = v3<T>(aq[0], aq[1], aq[2]);
    Q4<T> g_r2 = q4<T>(aq[3], aq[4], aq[5], aq[6]);
    adj_d[0] = g_p2.x; adj_d[1] = g_p2.y; adj_d[2] = g_p2.z;
    M3<T> G = m3_zero<T>();
    m3_acc(G, g_p2, p);
    Q4<T> g_dq = qnormalize_adj(du, dl, qmat_adj(du, G)) + qmul(g_r2, qconj(r));
    se3_quat_adj(6, dq, g_dq, adj_d + 3);
    // target = (p, r) also receives the adjoint that flows back through queried
    V3<T> g_p = v3<T>(at[0], at[1], at[2]) + mrot_t(Rd, g_p2);
    Q4<T> g_r = q4<T>(at[3], at[4], at[5], at[6]) + qmul(qconj(dq.q), g_r2);
    // p = Rg q.xyz + g.xyz ;  r = g (x) q.quat
    adj_g[0] = g_p.x; adj_g[1] = g_p.y; adj_g[2] = g_p.z;
    M3<T> Gg = m3_zero<T>();
    m3_acc(Gg, g_p, qx);
    Q4<T> g_g = qnormalize_adj(gu, gl, qmat_adj(gu, Gg)) + qmul(g_r, qconj(qq));
    adj_g[3] = g_g.x; adj_g[4] = g_g.y; adj_g[5] = g_g.z; adj_g[6] = g_g.w;
}

}  // namespace ppr
