// CPU PORT (test / baseline infrastructure, not product code) -- PARITY UNPINNED (see oracle/sim_oracle.py).
//
// Host-core restatement of the reference's rollout path, structured the way the reference runs it: per substep
// a contact pass over EVERY contact point (eval_body_contacts, diffphys/integrator_euler.py:93-179), a joint
// pass over every body (eval_body_joints :289-451), an integrate pass (integrate_bodies :21-91), all states kept
// for the reverse sweep (dp_model.py:396-399) which replays the adjoints in reverse order like wp.Tape
// (dp_model.py:1275).  It is the timed "cpu_baseline" / "--impl reference" arm of bench.py (kind = "port":
// the true Warp CPU device cannot be installed here) and a float64 cross-check of the hand-written adjoints
// against the autograd oracle.  The per-body arithmetic is the scalar-templated header the CUDA kernels use
// (ppr_diffphys_b200/csrc/ppr_body.h); the INDEPENDENT parity oracle is oracle/sim_oracle.py.
//
// Environments are independent -> OpenMP parallel-for over envs.
#include <stdint.h>
#include <string.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/ppr_b200.h"
#include "../../ppr_diffphys_b200/csrc/ppr_body.h"
#include "../../ppr_diffphys_b200/csrc/ppr_loss.h"
#include "../../ppr_diffphys_b200/csrc/ppr_frame.h"

using namespace ppr;

namespace {

template <class T> struct CpuModel {
    int nb, nq, nqd, nc;
    std::vector<int> type, parent, qs, qds, ndof, cbody;
    std::vector<JointStatic<T>> js;
    std::vector<V3<T>> com, cpoint;
    std::vector<T> cdist, lo, hi, lke, lkd;
    std::vector<ContactMat<T>> cmat;
    V3<T> g;
    T ake, akd;
};

template <class T> CpuModel<T> make_model(const ppr_model_desc* d) {
    CpuModel<T> m;
    m.nb = d->nb; m.nq = d->nq; m.nqd = d->nqd; m.nc = d->nc;
    m.g = v3<T>(d->gravity[0], d->gravity[1], d->gravity[2]);
    m.ake = d->joint_attach_ke; m.akd = d->joint_attach_kd;
    for (int i = 0; i < m.nb; ++i) {
        JointStatic<T> s;
        s.type = d->joint_type[i];
        const float* xp = d->joint_X_p + 7 * i;
        const float* xc = d->joint_X_c + 7 * i;
        s.xpj = v3<T>(xp[0], xp[1], xp[2]);
        s.qpj = q4<T>(xp[3], xp[4], xp[5], xp[6]);
        s.qoff = q4<T>(xc[3], xc[4], xc[5], xc[6]);
        s.axis = v3<T>(d->joint_axis[3 * i], d->joint_axis[3 * i + 1], d->joint_axis[3 * i + 2]);
        m.js.push_back(s);
        m.type.push_back(s.type);
        m.parent.push_back(d->joint_parent[i]);
        m.qs.push_back(d->joint_q_start[i]);
        m.qds.push_back(d->joint_qd_start[i]);
        int nd = (i + 1 < m.nb ? d->joint_qd_start[i + 1] : m.nqd) - d->joint_qd_start[i];
        m.ndof.push_back(nd);
        m.com.push_back(v3<T>(d->body_com[3 * i], d->body_com[3 * i + 1], d->body_com[3 * i + 2]));
    }
    for (int k = 0; k < m.nqd; ++k) {
        m.lo.push_back(d->joint_limit_lower[k]); m.hi.push_back(d->joint_limit_upper[k]);
        m.lke.push_back(d->joint_limit_ke[k]); m.lkd.push_back(d->joint_limit_kd[k]);
    }
    for (int k = 0; k < m.nc; ++k) {
        m.cbody.push_back(d->contact_body[k]);
        m.cpoint.push_back(v3<T>(d->contact_point[3 * k], d->contact_point[3 * k + 1], d->contact_point[3 * k + 2]));
        m.cdist.push_back(d->contact_dist[k]);
        const float* mt = d->shape_materials + 4 * d->contact_material[k];
        ContactMat<T> cm; cm.ke = mt[0]; cm.kd = mt[1]; cm.kf = mt[2]; cm.mu = mt[3];
        m.cmat.push_back(cm);
    }
    return m;
}

template <class T> Body<T> load_body(const T* q, const T* qd) {
    Body<T> b;
    b.x = v3<T>(q[0], q[1], q[2]); b.r = q4<T>(q[3], q[4], q[5], q[6]);
    b.w = v3<T>(qd[0], qd[1], qd[2]); b.v = v3<T>(qd[3], qd[4], qd[5]);
    return b;
}
template <class T> void store_body(const Body<T>& b, T* q, T* qd) {
    q[0] = b.x.x; q[1] = b.x.y; q[2] = b.x.z; q[3] = b.r.x; q[4] = b.r.y; q[5] = b.r.z; q[6] = b.r.w;
    qd[0] = b.w.x; qd[1] = b.w.y; qd[2] = b.w.z; qd[3] = b.v.x; qd[4] = b.v.y; qd[5] = b.v.z;
}
template <class T> void store_wrench(const Wrench<T>& w, T* o) {
    o[0] = w.t.x; o[1] = w.t.y; o[2] = w.t.z; o[3] = w.f.x; o[4] = w.f.y; o[5] = w.f.z;
}

template <class T>
JointCtl<T> load_ctl(const CpuModel<T>& m, int j, const T* refs, const T* act, const T* ke, const T* kd) {
    JointCtl<T> c;
    for (int k = 0; k < 3; ++k) {
        bool on = k < m.ndof[j] && m.type[j] != JT_FREE;
        int d = m.qds[j] + k;
        c.target[k] = on ? refs[d] : T(0); c.act[k] = on && act ? act[d] : T(0);
        c.ke[k] = on ? ke[d] : T(0); c.kd[k] = on ? kd[d] : T(0);
        c.lo[k] = on ? m.lo[d] : T(-1e30); c.hi[k] = on ? m.hi[d] : T(1e30);
        c.lke[k] = on ? m.lke[d] : T(0); c.lkd[k] = on ? m.lkd[d] : T(0);
    }
    return c;
}

// FK for one articulation
template <class T> void fk_env(const CpuModel<T>& m, const T* jq, const T* jqd, Body<T>* out) {
    for (int i = 0; i < m.nb; ++i) {
        Body<T> P = m.parent[i] >= 0 ? out[m.parent[i]] : body_identity<T>();
        out[i] = fk_joint_fwd(m.js[i], m.com[i], P, jq + m.qs[i], jqd + m.qds[i]);
    }
}
template <class T>
void fk_env_adj(const CpuModel<T>& m, const T* jq, const T* jqd, const Body<T>* bodies, Body<T>* adj, T* adj_jq,
                T* adj_jqd) {
    for (int i = m.nb - 1; i >= 0; --i) {
        Body<T> P = m.parent[i] >= 0 ? bodies[m.parent[i]] : body_identity<T>();
        Body<T> adjP = body_zero<T>();
        fk_joint_adj(m.js[i], m.com[i], P, jq + m.qs[i], jqd + m.qds[i], adj[i], adjP, adj_jq + m.qs[i],
                     adj_jqd + m.qds[i]);
        if (m.parent[i] >= 0) body_acc(adj[m.parent[i]], adjP);
    }
}

// forces of one substep for one env; returns F (total), optionally grf / jaf
template <class T>
void forces_env(const CpuModel<T>& m, const Body<T>* s, const M3<T>* R, const V3<T>* xc, const T* res_f, const T* refs,
                const T* act, const T* ke, const T* kd, Wrench<T>* F, T* grf, T* jaf) {
    for (int b = 0; b < m.nb; ++b) {
        if (res_f) { F[b].t = v3<T>(res_f[6 * b], res_f[6 * b + 1], res_f[6 * b + 2]);
                     F[b].f = v3<T>(res_f[6 * b + 3], res_f[6 * b + 4], res_f[6 * b + 5]); }
        else F[b] = wrench_zero<T>();
    }
    for (int k = 0; k < m.nc; ++k) {
        int b = m.cbody[k];
        contact_point_fwd(s[b], R[b], xc[b], m.cpoint[k], m.cdist[k], m.cmat[k], F[b]);
    }
    if (grf) for (int b = 0; b < m.nb; ++b) store_wrench(F[b], grf + 6 * b);
    for (int j = 0; j < m.nb; ++j) {
        int p = m.parent[j];
        JointCtl<T> c = load_ctl(m, j, refs, act, ke, kd);
        V3<T> t, f, ap, ac;
        Body<T> P = p >= 0 ? s[p] : body_identity<T>();
        joint_fwd(m.js[j], c, m.ake, m.akd, P, p >= 0 ? xc[p] : vzero<T>(), p >= 0, s[j], R[j], xc[j], t, f, ap, ac);
        if (m.type[j] == JT_FREE) continue;
        if (p >= 0) { F[p].t += t + cross(ap, f); F[p].f += f; }
        F[j].t -= t + cross(ac, f); F[j].f -= f;
    }
    if (jaf) for (int b = 0; b < m.nb; ++b) {
        T tmp[6]; store_wrench(F[b], tmp);
        for (int i = 0; i < 6; ++i) jaf[6 * b + i] = tmp[i] - (grf ? grf[6 * b + i] : T(0));
    }
}

template <class T>
int rollout_forward(const ppr_model_desc* d, int64_t bs, int64_t T_, int64_t stride, T dt, const T* q_init,
                    const T* qd_init, const T* torques, const T* res_f, const T* refs, const T* ke, const T* kd,
                    const T* inv_m, const T* I, const T* inv_I, T* out_pos, T* out_vel, T* out_grf, T* out_jaf,
                    T* states) {
    CpuModel<T> m = make_model<T>(d);
    const int nb = m.nb, nq = m.nq, nqd = m.nqd;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < bs; ++e) {
        std::vector<Body<T>> s(nb), s1(nb);
        std::vector<V3<T>> xc(nb);
        std::vector<M3<T>> R(nb);
        std::vector<Wrench<T>> F(nb);
        fk_env(m, q_init + e * nq, qd_init + e * nqd, s.data());
        for (int64_t t = 0; t < T_; ++t) {
            T* st = states + ((t * bs + e) * nb) * 13;
            for (int b = 0; b < nb; ++b) store_body(s[b], st + 13 * b, st + 13 * b + 7);
            bool frame = (t % stride) == 0;
            int64_t fidx = t / stride;
            if (frame) for (int b = 0; b < nb; ++b)
                store_body(s[b], out_pos + ((fidx * bs + e) * nb + b) * 7, out_vel + ((fidx * bs + e) * nb + b) * 6);
            for (int b = 0; b < nb; ++b) { R[b] = qmat(s[b].r); xc[b] = s[b].x + mrot(R[b], m.com[b]); }
            forces_env(m, s.data(), R.data(), xc.data(), res_f ? res_f + (t * bs + e) * nb * 6 : nullptr,
                       refs + (t * bs + e) * nqd, torques ? torques + (t * bs + e) * nqd : nullptr, ke + e * nqd,
                       kd + e * nqd, F.data(), (frame && out_grf) ? out_grf + (fidx * bs + e) * nb * 6 : nullptr,
                       (frame && out_jaf) ? out_jaf + (fidx * bs + e) * nb * 6 : nullptr);
            for (int b = 0; b < nb; ++b)
                s1[b] = integrate_fwd(s[b], R[b], xc[b], m.com[b], F[b], inv_m[e * nb + b], I + (e * nb + b) * 9,
                                      inv_I + (e * nb + b) * 9, m.g, dt);
            s.swap(s1);
        }
    }
    return 0;
}

template <class T>
int rollout_backward(const ppr_model_desc* d, int64_t bs, int64_t T_, int64_t stride, T dt, const T* q_init,
                     const T* qd_init, const T* torques, const T* res_f, const T* refs, const T* ke, const T* kd,
                     const T* inv_m, const T* I, const T* inv_I, const T* states, const T* adj_pos, const T* adj_vel,
                     T* adj_q_init, T* adj_qd_init, T* adj_torques, T* adj_res_f, T* adj_refs, T* adj_ke, T* adj_kd,
                     T* adj_inv_m, T* adj_I, T* adj_inv_I) {
    CpuModel<T> m = make_model<T>(d);
    const int nb = m.nb, nq = m.nq, nqd = m.nqd;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < bs; ++e) {
        std::vector<Body<T>> s(nb), adjS(nb), adjN(nb);
        std::vector<M3<T>> R(nb), G(nb);
        std::vector<V3<T>> xc(nb), adj_xc(nb);
        std::vector<Wrench<T>> F(nb), adjF(nb);
        for (int k = 0; k < nqd; ++k) { adj_ke[e * nqd + k] = 0; adj_kd[e * nqd + k] = 0; }
        for (int b = 0; b < nb; ++b) {
            adj_inv_m[e * nb + b] = 0;
            for (int i = 0; i < 9; ++i) { adj_I[(e * nb + b) * 9 + i] = 0; adj_inv_I[(e * nb + b) * 9 + i] = 0; }
        }
        // adjoint of the state after the last differentiated substep = seed of the last frame
        int64_t last = T_ - 1;  // state index of the last frame output (T = stride*(F-1)+1)
        for (int b = 0; b < nb; ++b) adjN[b] = body_zero<T>();
        auto add_seed = [&](int64_t t, std::vector<Body<T>>& a) {
            if (t % stride) return;
            int64_t fidx = t / stride;
            for (int b = 0; b < nb; ++b) {
                Body<T> g = load_body(adj_pos + ((fidx * bs + e) * nb + b) * 7, adj_vel + ((fidx * bs + e) * nb + b) * 6);
                body_acc(a[b], g);
            }
        };
        add_seed(last, adjN);
        // the substep last -> last+1 only feeds the force side channels: zero gradient (dp_model.py:397)
        {
            T* r = adj_refs + (last * bs + e) * nqd;
            for (int k = 0; k < nqd; ++k) r[k] = 0;
            if (adj_torques) for (int k = 0; k < nqd; ++k) adj_torques[(last * bs + e) * nqd + k] = 0;
            if (adj_res_f) for (int k = 0; k < nb * 6; ++k) adj_res_f[(last * bs + e) * nb * 6 + k] = 0;
        }
        for (int64_t t = last - 1; t >= 0; --t) {
            const T* st = states + ((t * bs + e) * nb) * 13;
            for (int b = 0; b < nb; ++b) s[b] = load_body(st + 13 * b, st + 13 * b + 7);
            for (int b = 0; b < nb; ++b) { R[b] = qmat(s[b].r); xc[b] = s[b].x + mrot(R[b], m.com[b]); }
            const T* refs_t = refs + (t * bs + e) * nqd;
            const T* act_t = torques ? torques + (t * bs + e) * nqd : nullptr;
            forces_env(m, s.data(), R.data(), xc.data(), res_f ? res_f + (t * bs + e) * nb * 6 : nullptr, refs_t, act_t,
                       ke + e * nqd, kd + e * nqd, F.data(), (T*)nullptr, (T*)nullptr);
            // K5^T
            for (int b = 0; b < nb; ++b) {
                adjS[b] = body_zero<T>(); adj_xc[b] = vzero<T>(); G[b] = m3_zero<T>();
                integrate_adj(s[b], R[b], xc[b], m.com[b], F[b], inv_m[e * nb + b], I + (e * nb + b) * 9,
                              inv_I + (e * nb + b) * 9, m.g, dt, adjN[b], adjS[b], G[b], adj_xc[b], adjF[b],
                              adj_inv_m[e * nb + b], adj_I + (e * nb + b) * 9, adj_inv_I + (e * nb + b) * 9);
            }
            // K4^T
            T* ar = adj_refs + (t * bs + e) * nqd;
            for (int k = 0; k < nqd; ++k) ar[k] = 0;
            T* at = adj_torques ? adj_torques + (t * bs + e) * nqd : nullptr;
            if (at) for (int k = 0; k < nqd; ++k) at[k] = 0;
            for (int j = 0; j < nb; ++j) {
                if (m.type[j] == JT_FREE) continue;
                int p = m.parent[j];
                JointCtl<T> c = load_ctl(m, j, refs_t, act_t, ke + e * nqd, kd + e * nqd);
                T a_target[3] = {0, 0, 0}, a_act[3] = {0, 0, 0}, a_ke[3] = {0, 0, 0}, a_kd[3] = {0, 0, 0};
                Body<T> P = p >= 0 ? s[p] : body_identity<T>();
                Body<T> adjP = body_zero<T>();
                V3<T> adj_xcp = vzero<T>();
                Wrench<T> aFp = p >= 0 ? adjF[p] : wrench_zero<T>();
                joint_adj(m.js[j], c, m.ake, m.akd, P, p >= 0 ? xc[p] : vzero<T>(), p >= 0, s[j], R[j], xc[j], aFp,
                          adjF[j], adjP, adj_xcp, adjS[j], G[j], adj_xc[j], a_target, a_act, a_ke, a_kd);
                if (p >= 0) { body_acc(adjS[p], adjP); adj_xc[p] += adj_xcp; }
                for (int k = 0; k < m.ndof[j] && k < 3; ++k) {
                    int dd = m.qds[j] + k;
                    ar[dd] += a_target[k];
                    if (at) at[dd] += a_act[k];
                    adj_ke[e * nqd + dd] += a_ke[k];
                    adj_kd[e * nqd + dd] += a_kd[k];
                }
            }
            // K3^T
            for (int k = 0; k < m.nc; ++k) {
                int b = m.cbody[k];
                contact_point_adj(s[b], R[b], xc[b], m.cpoint[k], m.cdist[k], m.cmat[k], adjF[b], adjS[b], G[b], adj_xc[b]);
            }
            // K2^T
            if (adj_res_f) for (int b = 0; b < nb; ++b) store_wrench(adjF[b], adj_res_f + ((t * bs + e) * nb + b) * 6);
            // world-COM adjoint -> (x, r)
            for (int b = 0; b < nb; ++b) {
                adjS[b].x += adj_xc[b];
                m3_acc(G[b], adj_xc[b], m.com[b]);
                adjS[b].r += qmat_adj(s[b].r, G[b]);
            }
            add_seed(t, adjS);
            adjN.swap(adjS);
        }
        // K1^T (state 0 = eval_fk(q_init, qd_init))
        fk_env(m, q_init + e * nq, qd_init + e * nqd, s.data());
        for (int k = 0; k < nq; ++k) adj_q_init[e * nq + k] = 0;
        for (int k = 0; k < nqd; ++k) adj_qd_init[e * nqd + k] = 0;
        fk_env_adj(m, q_init + e * nq, qd_init + e * nqd, s.data(), adjN.data(), adj_q_init + e * nq,
                   adj_qd_init + e * nqd);
    }
    return 0;
}

template <class T>
int fk_forward(const ppr_model_desc* d, int64_t n, const T* jq, const T* jqd, T* body_q, T* body_qd) {
    CpuModel<T> m = make_model<T>(d);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        std::vector<Body<T>> s(m.nb);
        fk_env(m, jq + e * m.nq, jqd + e * m.nqd, s.data());
        for (int b = 0; b < m.nb; ++b) store_body(s[b], body_q + (e * m.nb + b) * 7, body_qd + (e * m.nb + b) * 6);
    }
    return 0;
}
template <class T>
int fk_backward(const ppr_model_desc* d, int64_t n, const T* jq, const T* jqd, const T* adj_q, const T* adj_qd,
                T* adj_jq, T* adj_jqd) {
    CpuModel<T> m = make_model<T>(d);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        std::vector<Body<T>> s(m.nb), a(m.nb);
        fk_env(m, jq + e * m.nq, jqd + e * m.nqd, s.data());
        for (int b = 0; b < m.nb; ++b) a[b] = load_body(adj_q + (e * m.nb + b) * 7, adj_qd + (e * m.nb + b) * 6);
        for (int k = 0; k < m.nq; ++k) adj_jq[e * m.nq + k] = 0;
        for (int k = 0; k < m.nqd; ++k) adj_jqd[e * m.nqd + k] = 0;
        fk_env_adj(m, jq + e * m.nq, jqd + e * m.nqd, s.data(), a.data(), adj_jq + e * m.nq, adj_jqd + e * m.nqd);
    }
    return 0;
}

}  // namespace

#define PPR_CPU_API(SUF, T)                                                                                          \
    extern "C" int ppr_cpu_rollout_forward_##SUF(                                                                    \
        const ppr_model_desc* d, int64_t bs, int64_t nsteps, int64_t stride, double dt, const T* q_init,             \
        const T* qd_init, const T* torques, const T* res_f, const T* refs, const T* ke, const T* kd, const T* inv_m, \
        const T* I, const T* inv_I, T* out_pos, T* out_vel, T* out_grf, T* out_jaf, T* states) {                      \
        return rollout_forward<T>(d, bs, nsteps, stride, (T)dt, q_init, qd_init, torques, res_f, refs, ke, kd,       \
                                  inv_m, I, inv_I, out_pos, out_vel, out_grf, out_jaf, states);                       \
    }                                                                                                                \
    extern "C" int ppr_cpu_rollout_backward_##SUF(                                                                   \
        const ppr_model_desc* d, int64_t bs, int64_t nsteps, int64_t stride, double dt, const T* q_init,             \
        const T* qd_init, const T* torques, const T* res_f, const T* refs, const T* ke, const T* kd, const T* inv_m, \
        const T* I, const T* inv_I, const T* states, const T* adj_pos, const T* adj_vel, T* adj_q_init,              \
        T* adj_qd_init, T* adj_torques, T* adj_res_f, T* adj_refs, T* adj_ke, T* adj_kd, T* adj_inv_m, T* adj_I,     \
        T* adj_inv_I) {                                                                                              \
        return rollout_backward<T>(d, bs, nsteps, stride, (T)dt, q_init, qd_init, torques, res_f, refs, ke, kd,      \
                                   inv_m, I, inv_I, states, adj_pos, adj_vel, adj_q_init, adj_qd_init, adj_torques,  \
                                   adj_res_f, adj_refs, adj_ke, adj_kd, adj_inv_m, adj_I, adj_inv_I);                 \
    }                                                                                                                \
    extern "C" int ppr_cpu_fk_forward_##SUF(const ppr_model_desc* d, int64_t n, const T* jq, const T* jqd, T* bq,    \
                                            T* bqd) {                                                                \
        return fk_forward<T>(d, n, jq, jqd, bq, bqd);                                                                \
    }                                                                                                                \
    extern "C" int ppr_cpu_fk_backward_##SUF(const ppr_model_desc* d, int64_t n, const T* jq, const T* jqd,          \
                                             const T* aq, const T* aqd, T* ajq, T* ajqd) {                           \
        return fk_backward<T>(d, n, jq, jqd, aq, aqd, ajq, ajqd);                                                    \
    }

PPR_CPU_API(f32, float)
PPR_CPU_API(f64, double)

// se3 loss + adjoint in float64 through the SAME scalar-templated functions the CUDA kernels instantiate in float32
// (ppr_loss.h): lets the CPU suite check the hand-written adjoint against torch autograd without a GPU.
#define PPR_CPU_SE3(SUF, T)                                                                                          \
    extern "C" int ppr_cpu_se3_loss_##SUF(int64_t n, int dim, const T* pred, const T* gt, T ratio, T* loss,          \
                                          const T* adj_loss, T* adj_pred, T* adj_gt) {                               \
        for (int64_t i = 0; i < n; ++i) {                                                                            \
            loss[i] = ppr::se3_pair_loss<T>(dim, pred + i * dim, gt + i * dim, ratio, T(1e-4));                      \
            if (adj_loss)                                                                                            \
                ppr::se3_pair_loss_adj<T>(dim, pred + i * dim, gt + i * dim, ratio, T(1e-4), adj_loss[i],            \
                                          adj_pred + i * dim, adj_gt ? adj_gt + i * dim : nullptr);                  \
        }                                                                                                            \
        return 0;                                                                                                    \
    }
PPR_CPU_SE3(f64, double)
PPR_CPU_SE3(f32, float)

// batch-input frame composition + adjoint in float64 (ppr_frame.h), checked against torch autograd on the CPU
extern "C" int ppr_cpu_frame_compose_f64(int64_t n, const double* gq, const double* q, const double* d, double* target,
                                         double* queried, const double* at, const double* aq, double* adj_g,
                                         double* adj_d) {
    for (int64_t i = 0; i < n; ++i) {
        ppr::frame_compose<double>(gq, q + 7 * i, d + 6 * i, target + 7 * i, queried + 7 * i);
        if (at) ppr::frame_compose_adj<double>(gq, q + 7 * i, d + 6 * i, at + 7 * i, aq + 7 * i, adj_g + 7 * i, adj_d + 6 * i);
    }
    return 0;
}

// torchrun exports OMP_NUM_THREADS=1 to its workers; the timed CPU baseline must still use every host core
extern "C" void ppr_cpu_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}
extern "C" int ppr_cpu_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
