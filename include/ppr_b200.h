/* ppr_b200.h -- C ABI of the B200-native rollout path (libppr_b200.so).
 *
 * Drop-in boundary for the two torch.autograd Functions of the reference
 *   ForwardKinematics   /root/reference/diffphys/dp_model.py:1022-1130
 *   ForwardWarp         /root/reference/diffphys/dp_model.py:1145-1400
 * which in the reference drive Warp kernels (diffphys/integrator_euler.py:21-620 and warp.sim.eval_fk) under a
 * wp.Tape.  Everything here takes plain device pointers + sizes + a cudaStream_t (as void*); no torch types.
 *
 * Conventions (the reference's): fp32; quaternions xyzw; transform = (p[3], q[4]); spatial vectors =
 * (angular[3], linear[3]); env e owns bodies [e*nb,(e+1)*nb), coords [e*nq,(e+1)*nq), dofs [e*nqd,(e+1)*nqd)
 * (dp_model.py:563-572,697-699).
 *
 * Error convention: every function returns int: 0 ok, <0 argument / shape error (PPR_E_*), >0 a cudaError_t.
 * Nothing throws, nothing synchronises the device, no thread-local state: re-entrant from the autograd thread.
 */
#ifndef PPR_B200_H
#define PPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPR_E_ARG (-1)      /* null / inconsistent argument */
#define PPR_E_SHAPE (-2)    /* unsupported size (nb > 32, dofs per joint > 3, ...) */
#define PPR_E_HANDLE (-3)   /* bad model handle */
#define PPR_E_WORKSPACE (-4)/* workspace too small */

/* Static arrays of ONE articulation -- what the reference's kernels read from the Warp `Model`
 * (integrator_euler.py:497-504, 519-538, 603-611).  HOST pointers; copied at create time. */
typedef struct ppr_model_desc {
    int32_t nb, nq, nqd, nc, nshape;
    const int32_t* joint_type;      /* [nb]  Warp enum: 1 REVOLUTE, 3 FIXED, 4 FREE, 5 COMPOUND */
    const int32_t* joint_parent;    /* [nb]  -1 = world; parents precede children */
    const int32_t* joint_q_start;   /* [nb] */
    const int32_t* joint_qd_start;  /* [nb] */
    const float* joint_X_p;         /* [nb,7] */
    const float* joint_X_c;         /* [nb,7] */
    const float* joint_axis;        /* [nb,3] */
    const float* joint_limit_lower; /* [nqd] */
    const float* joint_limit_upper; /* [nqd] */
    const float* joint_limit_ke;    /* [nqd] */
    const float* joint_limit_kd;    /* [nqd] */
    const float* body_com;          /* [nb,3] */
    const int32_t* contact_body;    /* [nc] */
    const float* contact_point;     /* [nc,3] body frame */
    const float* contact_dist;      /* [nc] */
    const int32_t* contact_material;/* [nc] row of shape_materials */
    const float* shape_materials;   /* [nshape,4] ke kd kf mu */
    float gravity[3];
    float joint_attach_ke, joint_attach_kd;
} ppr_model_desc;

typedef struct ppr_model* ppr_model_t;

const char* ppr_version(void);

/* Uploads one copy of the static arrays to the current CUDA device (reference: ModelBuilder.finalize +
 * Model.collide, dp_model.py:384-401, which replicate them num_envs times). */
int ppr_model_create(const ppr_model_desc* desc, ppr_model_t* out);
int ppr_model_destroy(ppr_model_t m);
/* lab4d overwrites env.joint_X_p from torch every step (diffphys/dp_interface.py:465): HOST pointer, [nb,7]. */
int ppr_model_set_joint_X_p(ppr_model_t m, const float* joint_X_p, void* stream);
/* Per-environment joint_X_p, the form lab4d actually assigns (dp_interface.py:454-465: one [nb,7] block per env,
 * because every video instance has its own bone lengths): DEVICE pointer to [n_env, nb, 7], read zero-copy by every
 * later FK / rollout call (like wp.from_torch) -- the caller keeps it alive; articulation i of a call uses block
 * i % n_env (FK frames are laid out [T, bs]).  NULL / 0 returns to the shared table above. */
int ppr_model_set_joint_X_p_env(ppr_model_t m, const float* dev_joint_X_p, int64_t n_env);
int ppr_model_set_attach(ppr_model_t m, float attach_ke, float attach_kd);
int ppr_model_set_gravity(ppr_model_t m, const float g[3]);
/* env.ground (dp_model.py:390): 0 skips the ground-contact kernel like compute_forces does when model.ground is False
 * (integrator_euler.py:492-510); grf then equals res_f.  Default 1. */
int ppr_model_set_ground(ppr_model_t m, int32_t ground);
/* Checkpoint policy of the rollout (default 1): the forward pass keeps the per-substep state every `every` substeps;
 * the adjoint re-computes the substeps in between, segment by segment (costs (every-1)/every of a forward pass,
 * shrinks the workspace `every`-fold).  The reference keeps one full Warp State + gradient mirror per substep
 * (dp_model.py:396-399).  Must be set before ppr_rollout_workspace_bytes / _forward and unchanged until _backward. */
int ppr_model_set_checkpoint_every(ppr_model_t m, int32_t every);
/* Rollouts of at most `max_envs` environments use the LATENCY layout (one environment per warp: the substep
 * latency, not the issue rate, bounds a batch that cannot fill the GPU -- the reference's own shapes, 10..64 envs x
 * 760 substeps, dp_model.py:354-367); larger ones use the throughput packing below.  Default 1024
 * (about two warps per scheduler of a B200, the measured break-even); 0 disables.  Same rule as the checkpoint policy: set before workspace_bytes. */
int ppr_model_set_latency_envs(ppr_model_t m, int64_t max_envs);
int64_t ppr_model_latency_envs(ppr_model_t m);
/* Rollouts of at most `max_envs` environments (and within the latency rule above, checkpoint policy 1) use the TEAM
 * layout: one environment per thread block of three warps -- the main warp runs the substep, two helper warps evaluate
 * the ground contacts (eval_body_contacts, integrator_euler.py:93-179; half of a substep's instructions) of their share
 * of the bodies concurrently on the other schedulers of the SM.  1.6-1.9x lower substep latency than one warp per
 * environment on the reference's own shapes.  Default 296 (two blocks per SM of a B200); 0 disables.  Set before
 * workspace_bytes. */
int ppr_model_set_team_envs(ppr_model_t m, int64_t max_envs);
int64_t ppr_model_team_envs(ppr_model_t m);
/* introspection of the THROUGHPUT packing chosen for this articulation: a group is a warp (32 threads) or a
 * thread block (96 / 160 threads); each group hosts floor(threads / nb) environments, one thread per body. */
int ppr_model_envs_per_group(ppr_model_t m);
int ppr_model_group_threads(ppr_model_t m);

/* ---- articulation FK (replaces eval_fk launches at dp_model.py:1068 and its tape adjoint :1101) ------------
 * n = number of independent articulations (reference: T frames x bs envs, one launch per frame).
 * joint_q [n,nq], joint_qd [n,nqd] -> body_q [n,nb,7], body_qd [n,nb,6]. */
int ppr_fk_forward(ppr_model_t m, int64_t n, const float* joint_q, const float* joint_qd, float* body_q,
                   float* body_qd, void* stream);
/* adj_joint_q [n,nq], adj_joint_qd [n,nqd] are OVERWRITTEN. */
int ppr_fk_backward(ppr_model_t m, int64_t n, const float* joint_q, const float* joint_qd, const float* adj_body_q,
                    const float* adj_body_qd, float* adj_joint_q, float* adj_joint_qd, void* stream);

/* ---- rollout (replaces ForwardWarp.forward / .backward, dp_model.py:1147-1400) ----------------------------
 * nsteps = T substeps simulated (reference: len(steps_idx) = spf*(F-1)+1, dp_model.py:357-359);
 * frame_stride = spf; nframes = F outputs at t = k*frame_stride (t < nsteps).
 * The state saved for the adjoint lives in `workspace` (ppr_rollout_workspace_bytes); the same buffer must be
 * handed to ppr_rollout_backward.  Optional pointers may be NULL:
 *   torques, res_f              -> treated as exact zeros (the reference multiplies them by 0, dp_model.py:529,536)
 *   out_grf, out_jaf            -> force side channels at frame steps not written ([F,bs*nb,6] each)
 */
size_t ppr_rollout_workspace_bytes(ppr_model_t m, int64_t bs, int64_t nsteps);

/* shared_params = 0: target_ke/kd, body_inv_mass, body_inertia, body_inv_inertia are per-env replicated exactly as
 * the reference passes them (dp_model.py:723-730).  shared_params = 1: they are ONE copy shared by all envs
 * ([nqd], [nqd], [nb], [nb,3,3], [nb,3,3]) -- what the reference's replication encodes; the adjoint outputs stay
 * per-env (the caller sums them, which is the backward of the replication). */
int ppr_rollout_forward(ppr_model_t m, int64_t bs, int64_t nsteps, int64_t frame_stride, float dt,
                        int32_t shared_params,
                        const float* q_init,          /* [bs*nq] */
                        const float* qd_init,         /* [bs*nqd] */
                        const float* torques,         /* [T, bs*nqd] or NULL */
                        const float* res_f,           /* [T, bs*nb, 6] or NULL */
                        const float* refs,            /* [T, bs*nqd] */
                        const float* target_ke,       /* [bs*nqd]      ([nqd] if shared_params) */
                        const float* target_kd,       /* [bs*nqd]      ([nqd]) */
                        const float* body_inv_mass,   /* [bs*nb]       ([nb]) */
                        const float* body_inertia,    /* [bs*nb,3,3]   ([nb,3,3]) */
                        const float* body_inv_inertia,/* [bs*nb,3,3]   ([nb,3,3]) */
                        float* out_pos,               /* [F, bs*nb, 7] */
                        float* out_vel,               /* [F, bs*nb, 6] */
                        float* out_grf,               /* [F, bs*nb, 6] or NULL */
                        float* out_jaf,               /* [F, bs*nb, 6] or NULL */
                        void* workspace, size_t workspace_bytes, void* stream);

/* All adj_* outputs are OVERWRITTEN (not accumulated).  adj_torques / adj_res_f may be NULL.
 * adj_body_mass of the reference is identically zero (integrate_bodies never uses `m`,
 * integrator_euler.py:43) and is therefore not an output here. */
int ppr_rollout_backward(ppr_model_t m, int64_t bs, int64_t nsteps, int64_t frame_stride, float dt,
                         int32_t shared_params,
                         const float* q_init, const float* qd_init, const float* torques, const float* res_f,
                         const float* refs, const float* target_ke, const float* target_kd,
                         const float* body_inv_mass, const float* body_inertia, const float* body_inv_inertia,
                         const float* adj_out_pos,    /* [F, bs*nb, 7] */
                         const float* adj_out_vel,    /* [F, bs*nb, 6] */
                         float* adj_q_init,           /* [bs*nq] */
                         float* adj_qd_init,          /* [bs*nqd] */
                         float* adj_torques,          /* [T, bs*nqd] or NULL */
                         float* adj_res_f,            /* [T, bs*nb, 6] or NULL */
                         float* adj_refs,             /* [T, bs*nqd] */
                         float* adj_target_ke,        /* [bs*nqd] */
                         float* adj_target_kd,        /* [bs*nqd] */
                         float* adj_body_inv_mass,    /* [bs*nb] */
                         float* adj_body_inertia,     /* [bs*nb,3,3] */
                         float* adj_body_inv_inertia, /* [bs*nb,3,3] */
                         const void* workspace, size_t workspace_bytes, void* stream);

/* Shared-parameter mode with the reduction over environments done ON THE DEVICE ("backward-kernel epilogue
 * block-reduce -> packed buffer; this replaces the repeat() backward at dp_model.py:723-725"): the five parameter inputs
 * are the shared copies ([nqd], [nqd], [nb], [nb,3,3], [nb,3,3]) and their gradients come back already summed over the
 * bs environments in ONE packed buffer
 *     adj_shared = [ target_ke nqd | target_kd nqd | body_inv_mass nb | body_inertia nb*9 | body_inv_inertia nb*9 ]
 * of ppr_rollout_shared_grad_floats(m) floats -- ready for the single all-reduce of the multi-GPU step.  The sum runs in a
 * fixed order (per thread block in the adjoint kernel's epilogue, then over blocks): deterministic.  `scratch` is a device
 * buffer of ppr_rollout_reduce_scratch_bytes(m, bs) bytes (one packed row per thread block). */
int64_t ppr_rollout_shared_grad_floats(ppr_model_t m);
size_t ppr_rollout_reduce_scratch_bytes(ppr_model_t m, int64_t bs);
int ppr_rollout_backward_shared(ppr_model_t m, int64_t bs, int64_t nsteps, int64_t frame_stride, float dt,
                                const float* q_init, const float* qd_init, const float* torques, const float* res_f,
                                const float* refs, const float* target_ke, const float* target_kd,
                                const float* body_inv_mass, const float* body_inertia, const float* body_inv_inertia,
                                const float* adj_out_pos, const float* adj_out_vel, float* adj_q_init, float* adj_qd_init,
                                float* adj_torques /* or NULL */, float* adj_res_f /* or NULL */, float* adj_refs,
                                float* adj_shared, void* scratch, size_t scratch_bytes, const void* workspace,
                                size_t workspace_bytes, void* stream);

/* ---- struct-argument form of the rollout: every option in one place.  Zero-initialise, fill what applies. --------------
 * Adds to the positional entry points above the FUSED POSE LOSS of the imitation objective (SURVEY.md 8f rank 1;
 * se3_loss(sim_position, target_position), dp_model.py:777 with dp_utils.py:113-138): with `target_pos` [F, bs*nb, 7]
 * and `loss_pos` [F, bs*nb] set, the forward kernel evaluates the per-body pose loss at the frame steps while the pose is
 * still in registers; with `adj_loss_pos` [F, bs*nb] (= d objective / d loss_pos, from the caller's mean / clipping) the
 * adjoint kernel seeds itself from it, `target_pos` and the frame poses in `out_pos` -- no adj_pos tensor, no separate loss
 * kernels.  `adj_out_pos` / `adj_out_vel` stay available (either may be NULL) and are added on top. */
typedef struct ppr_rollout_io {
    int64_t bs, nsteps, frame_stride;
    float dt;
    int32_t shared_params;
    const float *q_init, *qd_init, *torques, *res_f, *refs, *target_ke, *target_kd, *body_inv_mass, *body_inertia,
        *body_inv_inertia;                                  /* as ppr_rollout_forward */
    float *out_pos, *out_vel, *out_grf, *out_jaf;           /* forward outputs; out_pos is READ by backward_ex when the loss is fused */
    void* workspace;
    size_t workspace_bytes;
    const float* target_pos;                                /* [F, bs*nb, 7] or NULL */
    float rot_ratio;                                        /* weight of the rotation angle in the pose loss (reference: 0.1) */
    float* loss_pos;                                        /* forward: [F, bs*nb] or NULL */
    const float* adj_loss_pos;                              /* backward: [F, bs*nb] or NULL */
    float* adj_target_pos;                                  /* backward: [F, bs*nb, 7] or NULL -- gradient w.r.t. the targets */
    const float *adj_out_pos, *adj_out_vel;                 /* backward: [F, bs*nb, 7] / [F, bs*nb, 6], each may be NULL */
    float *adj_q_init, *adj_qd_init, *adj_torques, *adj_res_f, *adj_refs;
    float *adj_target_ke, *adj_target_kd, *adj_body_inv_mass, *adj_body_inertia, *adj_body_inv_inertia; /* per-env form */
    float* adj_shared;                                      /* or: packed sums over envs (see ppr_rollout_backward_shared) */
    void* reduce_scratch;
    size_t reduce_scratch_bytes;
} ppr_rollout_io;
int ppr_rollout_forward_ex(ppr_model_t m, const ppr_rollout_io* io, void* stream);
int ppr_rollout_backward_ex(ppr_model_t m, const ppr_rollout_io* io, void* stream);

/* ---- control references from per-frame values (part of the batch-input producer; replaces the
 * host-side scipy interp1d of get_mocap_data for every substep, dp_model.py:421-427,605-609) -------------------------
 * refs[t, c] = lerp(frames[t / stride, c], frames[t / stride + 1, c], (t % stride) / stride) for t < T, c < n;
 * nframes >= (T - 1) / stride + 1 rows are read.  A caller ships nframes x n floats per window instead of T x n.
 * backward: adj_frames [nframes, n] is OVERWRITTEN with the transpose applied to adj_refs [T, n]. */
int ppr_refs_from_frames(int64_t T, int64_t stride, int64_t nframes, int64_t n, const float* frames, float* refs,
                         void* stream);
int ppr_refs_from_frames_backward(int64_t T, int64_t stride, int64_t nframes, int64_t n, const float* adj_refs,
                                  float* adj_frames, void* stream);

/* ---- se3 pose / twist loss (replaces se3_loss, dp_utils.py:113-138 with rot_angle,
 * geom_utils.py:37-46, called at dp_model.py:777,794,800) ------------------------------------------------------
 * n pairs of dim = 7 (xyz + quaternion xyzw) or dim = 6 (xyz + axis-angle) rows:
 *   loss[i] = |p.xyz - g.xyz|^2 + rot_ratio * acos(clamp((tr(R_p R_g^T) - 1)/2, -1 + 1e-4, 1 - 1e-4)),  0 if a row has NaN.
 * backward: adj_pred [n,dim] (and adj_gt [n,dim] unless NULL) are OVERWRITTEN with adj_loss[i] * dloss[i]/d(.). */
int ppr_se3_loss_forward(int64_t n, int32_t dim, const float* pred, const float* gt, float rot_ratio, float* loss,
                         void* stream);
int ppr_se3_loss_backward(int64_t n, int32_t dim, const float* pred, const float* gt, float rot_ratio,
                          const float* adj_loss, float* adj_pred, float* adj_gt, void* stream);

/* ---- batch-input frame composition (replaces rotate_frame + compose_delta, dp_utils.py:60-72,
 * 21-30, inside get_batch_input, dp_model.py:611-662) -------------------------------------------------------------
 * n time samples: target[i] = T(global_q) @ T(q[i]);  queried[i] = T(delta[i]) @ target[i], delta = (xyz, axis-angle).
 * global_q [7] (DEVICE), q [n,7], delta [n,6] -> target [n,7], queried [n,7].
 * backward: adj_global [n,7] holds the PER-SAMPLE gradient w.r.t. global_q (the caller sums over n), adj_delta [n,6];
 * both OVERWRITTEN.  q carries no gradient (mocap data). */
int ppr_frame_compose_forward(int64_t n, const float* global_q, const float* q, const float* delta, float* target,
                              float* queried, void* stream);
int ppr_frame_compose_backward(int64_t n, const float* global_q, const float* q, const float* delta,
                               const float* adj_target, const float* adj_queried, float* adj_global, float* adj_delta,
                               void* stream);

/* Number of kernels the library has launched since load (bench.py's gpu_launches). */
int64_t ppr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PPR_B200_H */
