// sm_100a kernels + C ABI of the rollout hot path (see include/ppr_b200.h for the boundary and the reference
// call sites each entry point replaces).
//
// Mapping: one THREAD per rigid body. Environments are packed either per WARP (floor(32/nb) envs per warp, slot =
// env * nb + body, tree exchange by warp shuffles) or per BLOCK (floor(NT/nb) envs per block of NT threads, body-major
// slots = position * envs + env, tree exchange through float4 areas of shared memory with one split-phase mbarrier and
// one hardware barrier per substep) -- whichever wastes fewer lanes: human has 19 bodies, i.e. 59 % lane use per warp
// but 99 % with 5 envs in a 96-thread block. The two are the `Comm` policy of the kernels below; batches too small
// to fill the GPU run one env per warp (latency layout) or one per block with helper warps (TeamComm; rollout_geometry).
// The body state (13 floats) lives in registers for the whole rollout; parent/child exchange of states and
// wrenches follows the static articulation tree with ordered reads (no atomics -> deterministic, unlike the
// reference's atomic_add/sub at integrator_euler.py:179,449,451).  The time loop is inside the kernel: one launch
// per rollout instead of the reference's 4 launches + 1 memset + 2 clones per substep (dp_model.py:1209-1228).
// Per substep the forward kernel streams the state, the total body wrench, the active-contact record and the joint
// angles (24 floats / body for REVOLUTE-only robots, 28 otherwise) to an HBM checkpoint buffer laid out
// [t][warp][quad][lane] as float4 quads: one STG.128 per quad, 512 contiguous bytes per warp instruction, and a lane
// only ever touches its own quads; the adjoint kernel streams the rows back in reverse (one substep ahead, into
// shared memory) and recomputes every other intermediate.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <vector>

#include "../../include/ppr_b200.h"
#include "ppr_body.h"
#include "ppr_loss.h"
#include "ppr_frame.h"

using namespace ppr;
typedef V3<float> F3;
typedef Q4<float> F4;
typedef Body<float> BodyF;
typedef Wrench<float> WrenchF;
typedef M3<float> M3F;

#define PPR_MAX_CHILD 8
// Checkpoint row of one body and substep = RQ float4 quads (quad q of lane l at float (q * 32 + l) * 4 of the warp's row):
//   Q0 x.xyz r.x | Q1 r.yzw w.x | Q2 w.yz v.xy | Q3 v.z ang0 rec[0..1] | then
//   RQ = 6 (REVOLUTE only): Q4 rec[2..3] F.t.xy | Q5 F.t.z F.f.xyz
//   RQ = 7                : Q4 rec[2..3] ang1 ang2 | Q5 F.t.xyz F.f.x | Q6 F.f.yz - -
// rec = active-contact record, 4 words = 8 halfwords: count, then up to PPR_REC_MAX point indices.
template <int JM> struct RowOf { static constexpr int kQuads = JM == JM_REVOLUTE ? 6 : 7; };
#define PPR_ROW_QUADS_MAX 7
#define PPR_REC_MAX 7
#define PPR_BLOCK 128  // FK kernels (warp layout)
#define PPR_SUP_N 16   // cells per edge of a cube-map face of the support-function table
#define PPR_SUP_FLOATS (6 * (PPR_SUP_N + 1) * (PPR_SUP_N + 1))
#ifndef PPR_CLIST_CAP
#define PPR_CLIST_CAP 16  // penetrating points listed per body before falling back to the cooperative path
#endif
#define PPR_CLIST_STRIDE (PPR_CLIST_CAP + 1)
#ifndef PPR_BODY_MAJOR
#define PPR_BODY_MAJOR 1
#endif
#define FULL 0xffffffffu

static std::atomic<int64_t> g_launches{0};

// ----------------------------------------------------------------------------------------------- device model
struct DevModel {
    int nb, nq, nqd, nc, epw, maxc, maxdepth, big_threshold;
    const int4* jinfo;        // [nb] type, parent, q_start, qd_start
    const int4* jinfo2;       // [nb] ndof, depth, contact begin, contact end
    const unsigned long long* child;  // [nb] 8 x uint8 child body index (0xff = none)
    const float* xpj_env;             // optional DEVICE array [xpj_nenv, nb, 7]: per-environment joint_X_p (else xpj)
    int64_t xpj_nenv;
    const int* order;                 // [nb] block layout: position in the block -> body
    const int* pos;                   // [nb] block layout: body -> position
    const float* xpj;         // [nb,7] joint_X_p
    const float* qoff;        // [nb,4] rot(joint_X_c)
    const float* axis;        // [nb,3]
    const float* com;         // [nb,3]
    const float4* lim;        // [nqd] lo, hi, lke, lkd
    const float4* cpt;        // [nc] body-frame point xyz + dist, sorted by body
    const int* cmat;          // [nc] material row
    const float4* mats;       // [nshape] ke kd kf mu
    const float* aabb;        // [nb,8] lo xyz, hi xyz, max dist, pad
    const float* sup;         // support-function tables of the big bodies (PPR_SUP_FLOATS each), see support_lower_bound
    const int* sup_of;        // [nb] table index of the body or -1
    float g[3], ake, akd;
    int mat_uniform;          // every contact uses material row cmat[0]
    int ground;               // 0: no ground plane, eval_body_contacts is skipped (model.ground, integrator_euler.py:492)
};

struct ppr_model {
    uint32_t magic;
    int device;
    DevModel d;
    void* blob;               // one device allocation holding every array
    size_t xpj_offset;        // byte offset of xpj inside blob
    std::vector<float> h_xpj;
    int variant;              // 0: FREE+REVOLUTE, 1: FREE+COMPOUND (both: no limits, identity q_off), 2: generic
    int ckpt_every;           // checkpoint policy K (1 = every substep, the fast default)
    int64_t latency_envs;     // batches up to this many envs run one env per warp (latency layout)
    int64_t team_envs;        // batches up to this many envs run one env per BLOCK of 1 + PPR_TEAM_H warps (team layout)
    int comm;                 // env packing of the rollout kernels: 0 per warp (128-thread blocks), 1 per 96-thread
                              // block, 2 per 160-thread block
};
#define PPR_TEAM_H 2          // helper warps of the team layout (see TeamComm)
#define PPR_COMM_TEAM 3
#ifndef PPR_NT1
#define PPR_NT1 96
#endif
static const int kCommThreads[4] = {128, PPR_NT1, 160, 32 * (1 + PPR_TEAM_H)};
#define PPR_MAGIC 0x50505231u
#define PPR_MAX_DEVICES 64
// The model's arrays live on the device that was current at ppr_model_create; every entry point that launches runs on
// that device whatever the caller's current device is (and restores it).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int want) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != want) err = cudaSetDevice(want); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ----------------------------------------------------------------------------------------------- lane helpers
__device__ __forceinline__ float shf(float v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ F3 shf3(F3 v, int src) { return v3<float>(shf(v.x, src), shf(v.y, src), shf(v.z, src)); }
__device__ __forceinline__ BodyF shf_body(const BodyF& b, int src) {
    BodyF o;
    o.x = shf3(b.x, src);
    o.r = q4<float>(shf(b.r.x, src), shf(b.r.y, src), shf(b.r.z, src), shf(b.r.w, src));
    o.w = shf3(b.w, src);
    o.v = shf3(b.v, src);
    return o;
}
__device__ __forceinline__ WrenchF shf_wrench(const WrenchF& w, int src) {
    WrenchF o; o.t = shf3(w.t, src); o.f = shf3(w.f, src); return o;
}
// Ampere-style asynchronous global->shared copies (LDGSTS): the adjoint kernel streams the NEXT checkpoint row into
// shared memory while it differentiates the current substep, so the ~700-cycle DRAM latency is never on the
// critical path and no registers are held for the prefetched values.
// 16-byte .cg copies: bypass L1 so that the streamed checkpoint does not evict the contact-point table from it.
__device__ __forceinline__ void cp_async16(volatile float* smem_dst, const float* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared((const void*)smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
// remove_nan of the reference (dp_utils.py:43-57, clip = False) applied at the store: NaN -> 0, everything else
// (including +-inf) untouched. Saves one full pass over every gradient tensor on the host side.
__device__ __forceinline__ float nan0(float v) { return v != v ? 0.f : v; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Split-phase CTA barriers (mbarrier): a thread ARRIVES as soon as it has published its values and only WAITS when
// it needs its partner's, so independent work (the contact pass) between the two hides the skew between the warps
// of a block instead of stalling on a __syncthreads.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
#ifdef PPR_MBAR_HINT_NS
        // suspend-time hint: the thread may sleep in hardware up to this long before try_wait returns false, so a long
        // wait costs a handful of polls instead of hundreds of issue slots taken from the other resident warps
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"((unsigned)PPR_MBAR_HINT_NS) : "memory");
#else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
#endif
    } while (!ok);
}

// TMA (bulk async copy engine): ONE elected lane moves a warp's whole checkpoint row (3 / 3.5 kB, contiguous in HBM) into
// shared memory with a single instruction; completion is signalled on an mbarrier by byte count.  Replaces the six or
// seven per-lane LDGSTS + commit/wait of the Ampere-style path (kept for the recompute instance, whose rows are written
// by the same kernel through the generic proxy).
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(volatile float* smem_dst, const float* gsrc, unsigned bytes, unsigned long long* bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared((const void*)smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// Row stream of one warp of the adjoint kernel: issue() starts the copy of a row into the warp's buffer, wait() blocks
// until the row issued last has landed.  BULK = TMA + mbarrier, else per-lane cp.async.
template <int RQ, bool BULK> struct RowStream {
    volatile float* buf;
    unsigned long long* bar;
    unsigned phase;
    int lane;
    __device__ __forceinline__ RowStream(volatile float* b, unsigned long long* m, int l) : buf(b), bar(m), phase(0), lane(l) {
        if (BULK) {
            if (lane == 0) {
                mbar_init(bar, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
        }
    }
    __device__ __forceinline__ void issue(const float* grow) {
        if (BULK) {
            __syncwarp();   // every lane has read the previous row out of the buffer
            if (lane == 0) {
                mbar_expect_tx(bar, RQ * 512u);
                bulk_load(buf, grow, RQ * 512u, bar);
            }
        } else {
#pragma unroll
            for (int i = 0; i < RQ; ++i) cp_async16(buf + (i * 32 + lane) * 4, grow + (i * 32 + lane) * 4);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    __device__ __forceinline__ void wait() {
        if (BULK) { mbar_wait(bar, phase); phase ^= 1u; }
        else asm volatile("cp.async.wait_all;" ::: "memory");
    }
};

// ----------------------------------------------------------------------------------------------- tree exchange
// Collectives over the articulation tree, executed by EVERY thread of the group (warp or block):
//   parent_*  : each thread obtains values held by the thread of its parent body
//   gather_*  : each thread sums a message held by the threads of its child bodies
// `child` packs up to 8 children as bytes: slot of the child + 1 (0 = none);
// `ps` = slot of the parent (own slot if none).
template <int NT, bool ADJ = true> struct WarpComm {
    static constexpr int kThreads = NT;
    static constexpr bool kBlock = false;
    static constexpr int kHelpers = 0;
    static constexpr int kExFloats = 0;
    __device__ __forceinline__ explicit WarpComm(float*) {}
    static __device__ __forceinline__ int64_t group() { return ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; }
    static __device__ __forceinline__ int slot() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int envs_per_group(const DevModel& M) { return M.epw; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ BodyF parent_body(const BodyF& s, int ps) const { return shf_body(s, ps); }
    __device__ __forceinline__ F3 parent_vec(F3 v, int ps) const { return shf3(v, ps); }
    __device__ __forceinline__ void init() {}
    // split-phase API: post_* publishes (no-op here), get_* / gather_* obtains the partner's values
    __device__ __forceinline__ void post_state(const BodyF&, F3) {}
    __device__ __forceinline__ void post_state_w(const BodyF&, F3, const WrenchF&) {}
    __device__ __forceinline__ void post_wrench(const WrenchF&) {}
    __device__ __forceinline__ void post_body(const BodyF&) {}
    __device__ __forceinline__ void get_parent_state(const BodyF& s, F3 xc, int ps, BodyF& P, F3& xcp) {
        P = shf_body(s, ps); xcp = shf3(xc, ps);
    }
    __device__ __forceinline__ void get_parent_state_w(const BodyF& s, F3 xc, const WrenchF& w, int ps, BodyF& P, F3& xcp,
                                                       WrenchF& wp) {
        P = shf_body(s, ps); xcp = shf3(xc, ps); wp = shf_wrench(w, ps);
    }
    __device__ __forceinline__ void gather_wrench(const WrenchF& mine, unsigned long long child, int maxc,
                                                  WrenchF& acc) const {
#pragma unroll 1
        for (int sl = 0; sl < maxc; ++sl) {
            unsigned c = (unsigned)((child >> (8 * sl)) & 0xffu);
            WrenchF o = shf_wrench(mine, c ? (int)c - 1 : (int)(threadIdx.x & 31));
            if (c) { acc.t += o.t; acc.f += o.f; }
        }
    }
    __device__ __forceinline__ void gather_body(const BodyF& mine, unsigned long long child, int maxc, BodyF& acc) const {
#pragma unroll 1
        for (int sl = 0; sl < maxc; ++sl) {
            unsigned c = (unsigned)((child >> (8 * sl)) & 0xffu);
            BodyF o = shf_body(mine, c ? (int)c - 1 : (int)(threadIdx.x & 31));
            if (c) body_acc(acc, o);
        }
    }
    __device__ __forceinline__ void gather_body_sync(const BodyF& mine, unsigned long long child, int maxc,
                                                     BodyF& acc) const {
        gather_body(mine, child, maxc, acc);
    }
};

// Team layout (batches of a few hundred environments at most, where the GPU is mostly idle and only the LATENCY of a
// substep counts): one environment per BLOCK of 1 + H warps.  Warp 0 (the main warp) runs the substep exactly like the
// warp layout (tree exchange by shuffles) except for the ground contacts, which are half of its instruction stream
// (profiles/README.md): those are evaluated concurrently by the H helper warps, each for its share of the bodies, on the
// other schedulers of the SM.  Hand-off through shared memory and two named hardware barriers per substep:
//   barrier 1  request published (main arrives and goes on with the joints; helpers wait)
//   barrier 2  replies published (helpers arrive; main waits when it needs the contact wrench / adjoint)
// A helper can only pass barrier 1 of substep t+1 after the main warp consumed the replies of t, and the main warp only
// re-writes the request after the helpers arrived at barrier 2 of t (i.e. finished reading it): no double buffering.
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int H, bool ADJ = true> struct TeamComm : WarpComm<32 * (1 + H), ADJ> {
    static constexpr int kThreads = 32 * (1 + H);
    static constexpr int kHelpers = H;
    // forward: request = body state + world COM (4 quads), reply = contact wrench + active-contact record (3 quads)
    // adjoint: request = wrench adjoint (2 quads), reply = body adjoint 13 + dL/dR 9 + world-COM adjoint 3 (7 quads)
    static constexpr int kReqQuads = ADJ ? 2 : 4, kRepQuads = ADJ ? 7 : 3;
    static constexpr int kExFloats = (kReqQuads + H * kRepQuads) * 4 * 32;
    float4* req;   // [kReqQuads][32]
    float4* rep;   // [H][kRepQuads][32]
    int helper;    // helper warp (0..H-1) that owns this lane's body
    __device__ __forceinline__ explicit TeamComm(float* sm)
        : WarpComm<32 * (1 + H), ADJ>(sm), req((float4*)sm), rep((float4*)sm + kReqQuads * 32), helper(0) {}
    static __device__ __forceinline__ int64_t group() { return blockIdx.x; }
    static __device__ __forceinline__ int role() { return threadIdx.x >> 5; }   // 0: main warp, 1..H: helper warps
    // bodies with contact points are dealt round-robin to the helpers (quadruped feet / biped feet split evenly)
    __device__ __forceinline__ void assign(bool has_points) {
        const unsigned m = __ballot_sync(FULL, has_points);
        helper = __popc(m & ((1u << (threadIdx.x & 31)) - 1u)) % H;
    }
    __device__ __forceinline__ bool mine() const { return helper == role() - 1; }
    // ---- forward
    __device__ __forceinline__ void request_contacts(const BodyF& s, F3 xc) {
        const int l = threadIdx.x & 31;
        req[0 * 32 + l] = make_float4(s.x.x, s.x.y, s.x.z, s.r.x);
        req[1 * 32 + l] = make_float4(s.r.y, s.r.z, s.r.w, s.w.x);
        req[2 * 32 + l] = make_float4(s.w.y, s.w.z, s.v.x, s.v.y);
        req[3 * 32 + l] = make_float4(s.v.z, xc.x, xc.y, xc.z);
        named_bar_arrive(1, kThreads);
    }
    __device__ __forceinline__ void await_request(BodyF& s, F3& xc) {
        named_bar_sync(1, kThreads);
        const int l = threadIdx.x & 31;
        const float4 u0 = req[0 * 32 + l], u1 = req[1 * 32 + l], u2 = req[2 * 32 + l], u3 = req[3 * 32 + l];
        s.x = v3<float>(u0.x, u0.y, u0.z); s.r = q4<float>(u0.w, u1.x, u1.y, u1.z);
        s.w = v3<float>(u1.w, u2.x, u2.y); s.v = v3<float>(u2.z, u2.w, u3.x);
        xc = v3<float>(u3.y, u3.z, u3.w);
    }
    __device__ __forceinline__ void reply_contacts(const WrenchF& F, unsigned long long rlo, unsigned long long rhi) {
        float4* r = rep + (role() - 1) * kRepQuads * 32 + (threadIdx.x & 31);
        r[0 * 32] = make_float4(F.t.x, F.t.y, F.t.z, F.f.x);
        r[1 * 32] = make_float4(F.f.y, F.f.z, __uint_as_float((unsigned)rlo), __uint_as_float((unsigned)(rlo >> 32)));
        r[2 * 32] = make_float4(__uint_as_float((unsigned)rhi), __uint_as_float((unsigned)(rhi >> 32)), 0.f, 0.f);
        named_bar_arrive(2, kThreads);
    }
    __device__ __forceinline__ void await_contacts(WrenchF& F, unsigned long long& rlo, unsigned long long& rhi) {
        named_bar_sync(2, kThreads);
        const float4* r = rep + helper * kRepQuads * 32 + (threadIdx.x & 31);
        const float4 a = r[0 * 32], b = r[1 * 32], c = r[2 * 32];
        F.t = v3<float>(a.x, a.y, a.z); F.f = v3<float>(a.w, b.x, b.y);
        rlo = (unsigned long long)__float_as_uint(b.z) | ((unsigned long long)__float_as_uint(b.w) << 32);
        rhi = (unsigned long long)__float_as_uint(c.x) | ((unsigned long long)__float_as_uint(c.y) << 32);
    }
    // ---- adjoint
    __device__ __forceinline__ void request_contacts_adj(const WrenchF& w) {
        const int l = threadIdx.x & 31;
        req[0 * 32 + l] = make_float4(w.t.x, w.t.y, w.t.z, w.f.x);
        req[1 * 32 + l] = make_float4(w.f.y, w.f.z, 0.f, 0.f);
        named_bar_arrive(1, kThreads);
    }
    __device__ __forceinline__ void await_request_adj(WrenchF& w) {
        named_bar_sync(1, kThreads);
        const int l = threadIdx.x & 31;
        const float4 a = req[0 * 32 + l], b = req[1 * 32 + l];
        w.t = v3<float>(a.x, a.y, a.z); w.f = v3<float>(a.w, b.x, b.y);
    }
    __device__ __forceinline__ void reply_contacts_adj(const BodyF& a, const M3F& G, F3 axc) {
        float4* r = rep + (role() - 1) * kRepQuads * 32 + (threadIdx.x & 31);
        r[0 * 32] = make_float4(a.x.x, a.x.y, a.x.z, a.r.x);
        r[1 * 32] = make_float4(a.r.y, a.r.z, a.r.w, a.w.x);
        r[2 * 32] = make_float4(a.w.y, a.w.z, a.v.x, a.v.y);
        r[3 * 32] = make_float4(a.v.z, axc.x, axc.y, axc.z);
        r[4 * 32] = make_float4(G.m[0], G.m[1], G.m[2], G.m[3]);
        r[5 * 32] = make_float4(G.m[4], G.m[5], G.m[6], G.m[7]);
        r[6 * 32] = make_float4(G.m[8], 0.f, 0.f, 0.f);
        named_bar_arrive(2, kThreads);
    }
    __device__ __forceinline__ void await_contacts_adj(BodyF& a, M3F& G, F3& axc) {
        named_bar_sync(2, kThreads);
        const float4* r = rep + helper * kRepQuads * 32 + (threadIdx.x & 31);
        const float4 u0 = r[0 * 32], u1 = r[1 * 32], u2 = r[2 * 32], u3 = r[3 * 32], u4 = r[4 * 32], u5 = r[5 * 32],
                     u6 = r[6 * 32];
        a.x += v3<float>(u0.x, u0.y, u0.z); a.r += q4<float>(u0.w, u1.x, u1.y, u1.z);
        a.w += v3<float>(u1.w, u2.x, u2.y); a.v += v3<float>(u2.z, u2.w, u3.x);
        axc += v3<float>(u3.y, u3.z, u3.w);
        G.m[0] += u4.x; G.m[1] += u4.y; G.m[2] += u4.z; G.m[3] += u4.w;
        G.m[4] += u5.x; G.m[5] += u5.y; G.m[6] += u5.z; G.m[7] += u5.w; G.m[8] += u6.x;
    }
};

// Block-wide variant: values are published in shared memory ([component][thread], conflict-free), one
// __syncthreads, then read at the partner's slot. In the substep loops a parent_state* call always alternates with
// a gather_* call (different areas `ex` / `msg`), which makes the single barrier per call sufficient: every read of
// an area happens before the next barrier, every write to it after one more barrier. parent_body / parent_vec are
// used back to back (FK levels) and therefore carry a trailing barrier.
template <int NT, bool ADJ = true> struct BlockComm {
    static constexpr int kThreads = NT;
    static constexpr bool kBlock = true;
    static constexpr int kHelpers = 0;
    // exchange areas are arrays of float4 [k][NT] (thread t's k-th quad at [k * NT + t]): one LDS.128 / STS.128 moves
    // what four scalar accesses did, conflict-free because consecutive threads touch consecutive 16-byte slots.
    // The forward kernel exchanges less (state down, wrench up) than the adjoint (state + wrench adjoint down, body
    // adjoint up): 6 instead of 10 quads per thread, which is 6 kB per block given back to the L1 cache.
    static constexpr int kExQuads = ADJ ? 6 : 4;   // body 13 + world COM 3 = 4 quads (+ wrench adjoint 6 (+2 pad) = 2 quads)
    static constexpr int kMsgQuads = ADJ ? 4 : 2;  // body adjoint 13 (+3 pad) / wrench 6 (+2 pad)
    static constexpr int kExFloats = (kExQuads + kMsgQuads) * 4 * NT + 4;  // + two 8-byte mbarriers
    float4* ex;
    float4* msg;
    unsigned long long* bar;  // [0]: `ex` published, [1]: `msg` published
    unsigned phA, phB;
    __device__ __forceinline__ explicit BlockComm(float* sm)
        : ex((float4*)sm), msg((float4*)sm + kExQuads * NT),
          bar((unsigned long long*)(sm + (kExQuads + kMsgQuads) * 4 * NT)), phA(0), phB(0) {}
    __device__ __forceinline__ void init() {
        if (threadIdx.x == 0) { mbar_init(bar, NT); mbar_init(bar + 1, NT); }
        __syncthreads();
    }
    static __device__ __forceinline__ int64_t group() { return blockIdx.x; }
    static __device__ __forceinline__ int slot() { return threadIdx.x; }
    static __device__ __forceinline__ int envs_per_group(const DevModel& M) { return NT / M.nb; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    // quads 0..3 of a body; the last three floats of quad 3 carry `tail` (the world COM in the state exchange)
    __device__ __forceinline__ void put_body(float4* a, const BodyF& s, F3 tail) const {
        const int t = threadIdx.x;
        a[0 * NT + t] = make_float4(s.x.x, s.x.y, s.x.z, s.r.x);
        a[1 * NT + t] = make_float4(s.r.y, s.r.z, s.r.w, s.w.x);
        a[2 * NT + t] = make_float4(s.w.y, s.w.z, s.v.x, s.v.y);
        a[3 * NT + t] = make_float4(s.v.z, tail.x, tail.y, tail.z);
    }
    __device__ __forceinline__ BodyF get_body(const float4* a, int t, F3& tail) const {
        const float4 q0 = a[0 * NT + t], q1 = a[1 * NT + t], q2 = a[2 * NT + t], q3 = a[3 * NT + t];
        BodyF o;
        o.x = v3<float>(q0.x, q0.y, q0.z);
        o.r = q4<float>(q0.w, q1.x, q1.y, q1.z);
        o.w = v3<float>(q1.w, q2.x, q2.y);
        o.v = v3<float>(q2.z, q2.w, q3.x);
        tail = v3<float>(q3.y, q3.z, q3.w);
        return o;
    }
    __device__ __forceinline__ void put_wrench(float4* a, const WrenchF& w) const {
        const int t = threadIdx.x;
        a[0 * NT + t] = make_float4(w.t.x, w.t.y, w.t.z, w.f.x);
        *(float2*)&a[1 * NT + t] = make_float2(w.f.y, w.f.z);
    }
    __device__ __forceinline__ WrenchF get_wrench(const float4* a, int t) const {
        const float4 q0 = a[0 * NT + t];
        const float2 q1 = *(const float2*)&a[1 * NT + t];
        WrenchF w;
        w.t = v3<float>(q0.x, q0.y, q0.z);
        w.f = v3<float>(q0.w, q1.x, q1.y);
        return w;
    }
    __device__ __forceinline__ BodyF parent_body(const BodyF& s, int ps) const {
        F3 tail;
        put_body(ex, s, vzero<float>());
        __syncthreads();
        BodyF o = get_body(ex, ps, tail);
        __syncthreads();
        return o;
    }
    __device__ __forceinline__ F3 parent_vec(F3 v, int ps) const {
        ex[threadIdx.x] = make_float4(v.x, v.y, v.z, 0.f);
        __syncthreads();
        const float4 q = ex[ps];
        __syncthreads();
        return v3<float>(q.x, q.y, q.z);
    }
    // ---- in the substep loops a post_state* / get_parent_state* pair (area `ex`, barrier A: split-phase mbarrier,
    // the contact pass runs between arrive and wait) always alternates with a post_* / gather_* pair (area `msg`,
    // barrier B: nothing to overlap, a plain hardware barrier that costs no polling issue slots); a thread can only
    // pass B of substep t after every thread has reached B, i.e. after it finished reading `ex` of substep t, and
    // can only pass wait(A) of t+1 after every thread arrived at A, i.e. finished reading `msg` of t.
    __device__ __forceinline__ void post_state(const BodyF& s, F3 xc) {
        put_body(ex, s, xc);
        mbar_arrive(bar);
    }
    __device__ __forceinline__ void post_state_w(const BodyF& s, F3 xc, const WrenchF& w) {
        put_body(ex, s, xc);
        put_wrench(ex + 4 * NT, w);
        mbar_arrive(bar);
    }
    __device__ __forceinline__ void get_parent_state(const BodyF&, F3, int ps, BodyF& P, F3& xcp) {
        mbar_wait(bar, phA); phA ^= 1u;
        P = get_body(ex, ps, xcp);
    }
    __device__ __forceinline__ void get_parent_state_w(const BodyF&, F3, const WrenchF&, int ps, BodyF& P, F3& xcp,
                                                       WrenchF& wp) {
        mbar_wait(bar, phA); phA ^= 1u;
        P = get_body(ex, ps, xcp);
        wp = get_wrench(ex + 4 * NT, ps);
    }
#ifndef PPR_BAR_B_SYNC   // barrier B split-phase as well: the callers place the work that does not depend on the gathered
                         // values (checkpoint stores, wrench-independent part of K5, own-body part of the pose adjoint)
                         // between post_* and gather_*.  -DPPR_BAR_B_SYNC: plain hardware barrier (round-1 behaviour)
#define PPR_ARRIVE_B() mbar_arrive(bar + 1)
#define PPR_WAIT_B() do { mbar_wait(bar + 1, phB); phB ^= 1u; } while (0)
#else
#define PPR_ARRIVE_B()
#define PPR_WAIT_B() __syncthreads()
#endif
    __device__ __forceinline__ void post_wrench(const WrenchF& mine) {
        put_wrench(msg, mine);
        PPR_ARRIVE_B();
    }
    __device__ __forceinline__ void post_body(const BodyF& mine) {
        put_body(msg, mine, vzero<float>());
        PPR_ARRIVE_B();
    }
    __device__ __forceinline__ void gather_wrench(const WrenchF&, unsigned long long child, int maxc, WrenchF& acc) {
        PPR_WAIT_B();
#pragma unroll 1
        for (int sl = 0; sl < maxc; ++sl) {
            unsigned c = (unsigned)((child >> (8 * sl)) & 0xffu);
            if (c) {
                const WrenchF w = get_wrench(msg, (int)c - 1);
                acc.t += w.t; acc.f += w.f;
            }
        }
    }
    __device__ __forceinline__ void gather_body(const BodyF&, unsigned long long child, int maxc, BodyF& acc) {
        PPR_WAIT_B();
        F3 tail;
#pragma unroll 1
        for (int sl = 0; sl < maxc; ++sl) {
            unsigned c = (unsigned)((child >> (8 * sl)) & 0xffu);
            if (c) body_acc(acc, get_body(msg, (int)c - 1, tail));
        }
    }
    // plain-barrier variant for the FK adjoint (outside the substep loop)
    __device__ __forceinline__ void gather_body_sync(const BodyF& mine, unsigned long long child, int maxc,
                                                     BodyF& acc) const {
        F3 tail;
        put_body(msg, mine, vzero<float>());
        __syncthreads();
#pragma unroll 1
        for (int sl = 0; sl < maxc; ++sl) {
            unsigned c = (unsigned)((child >> (8 * sl)) & 0xffu);
            if (c) body_acc(acc, get_body(msg, (int)c - 1, tail));
        }
        __syncthreads();
    }
};

// Shared-memory residency of everything that is constant over the time loop, so that it does not occupy
// registers for the whole rollout (the kernels are occupancy/latency bound):
//   static per-BODY table  sm_st[PPR_NSTATIC][32]   (indexed by body; lanes of different envs broadcast-read)
//   per-THREAD parameters   par[PPR_NPAR][NT]   inv_m, I[9], inv_I[9]
//   per-THREAD accumulators acc[18][NT]         adj_I, adj_inv_I        (adjoint kernel only)
// Reads go through volatile pointers so that ptxas re-issues the (29-cycle) LDS at the point of use instead of
// hoisting the values back into loop-long registers.
#define PPR_NSTATIC 28  // xpj 3, qpj 4, axis 3, com 3, parent com 3, aabb 7, qoff 4, pad 1
#define PPR_NPAR 20
enum { ST_XPJ = 0, ST_QPJ = 3, ST_AXIS = 7, ST_COM = 10, ST_CPAR = 13, ST_AABB = 16, ST_QOFF = 23 };

struct LaneInfo {
    int env, body, parent_slot, type, ndof, depth, qs, qds, c0, c1;
    int maxc_w;  // largest child count among the lanes of this WARP (trip count of the child-gather loops)
    bool valid, has_parent;
    unsigned long long child;  // up to 8 children as bytes: slot of the child + 1, 0 = none
    JointStatic<float> js;
    F3 com;
    float aabb[7];
};

// group = warp or block index, slot = lane or thread index inside it, epg = environments per group
// body_major = false: slot = env_in_group * nb + body (a warp holds whole environments: what the shuffle exchange needs)
// body_major = true : slot = body * epg + env_in_group (block layout): the lanes of a warp hold the SAME few bodies of
//   all the group's environments, so joint-type / has-parent / has-children branches and the child-gather loops are
//   (nearly) warp-uniform instead of every warp paying for the root's four children.
__device__ __forceinline__ LaneInfo lane_setup(const DevModel& M, int64_t group, int slot, int64_t n_env, int epg,
                                               bool body_major = false) {
    LaneInfo L;
    const int lane = slot;
    int e_in_w, body;
    if (body_major) {
        int p = slot / epg;
        e_in_w = slot - p * epg;
        body = p < M.nb ? M.order[p] : M.nb;   // the order balances the contact-heavy bodies over the block's warps
    } else { e_in_w = slot / M.nb; body = slot - e_in_w * M.nb; }
    int64_t env = group * epg + e_in_w;
    L.valid = (e_in_w < epg) && (body < M.nb) && (env < n_env);
    if (!L.valid) { e_in_w = 0; body = 0; env = group * epg; }
    L.env = (int)env; L.body = body;
    int4 ji = M.jinfo[body], j2 = M.jinfo2[body];
    // slot of body b of this lane's environment
    auto slot_of = [&](int b) { return body_major ? M.pos[b] * epg + e_in_w : e_in_w * M.nb + b; };
    L.type = ji.x; L.has_parent = ji.y >= 0; L.parent_slot = L.has_parent ? slot_of(ji.y) : lane;
    L.qs = ji.z; L.qds = ji.w; L.ndof = j2.x; L.depth = j2.y; L.c0 = j2.z; L.c1 = j2.w;
    unsigned long long ch = M.child[body], out = 0;
#pragma unroll
    for (int s = 0; s < PPR_MAX_CHILD; ++s) {
        unsigned c = (unsigned)((ch >> (8 * s)) & 0xffu);
        unsigned long long v = (c == 0xffu || !L.valid) ? 0ull : (unsigned long long)(slot_of((int)c) + 1);
        out |= v << (8 * s);
    }
    L.child = out;
    int nchild = 0;
#pragma unroll
    for (int s = 0; s < PPR_MAX_CHILD; ++s) if ((out >> (8 * s)) & 0xffull) nchild = s + 1;
    L.maxc_w = __reduce_max_sync(FULL, nchild);
    const float* xp = M.xpj_env ? M.xpj_env + ((env % M.xpj_nenv) * M.nb + body) * 7 : M.xpj + 7 * body;
    L.js.type = L.type;
    L.js.xpj = v3<float>(xp[0], xp[1], xp[2]);
    L.js.qpj = q4<float>(xp[3], xp[4], xp[5], xp[6]);
    L.js.qoff = q4<float>(M.qoff[4 * body], M.qoff[4 * body + 1], M.qoff[4 * body + 2], M.qoff[4 * body + 3]);
    L.js.axis = v3<float>(M.axis[3 * body], M.axis[3 * body + 1], M.axis[3 * body + 2]);
    L.com = v3<float>(M.com[3 * body], M.com[3 * body + 1], M.com[3 * body + 2]);
#pragma unroll
    for (int i = 0; i < 7; ++i) L.aabb[i] = M.aabb[8 * body + i];
    return L;
}

__device__ __forceinline__ void stage_static(volatile float* st, const LaneInfo& L, F3 com_par) {
    int b = L.body;
    st[(ST_AXIS + 0) * 32 + b] = L.js.axis.x; st[(ST_AXIS + 1) * 32 + b] = L.js.axis.y; st[(ST_AXIS + 2) * 32 + b] = L.js.axis.z;
    st[(ST_COM + 0) * 32 + b] = L.com.x; st[(ST_COM + 1) * 32 + b] = L.com.y; st[(ST_COM + 2) * 32 + b] = L.com.z;
    st[(ST_CPAR + 0) * 32 + b] = com_par.x; st[(ST_CPAR + 1) * 32 + b] = com_par.y; st[(ST_CPAR + 2) * 32 + b] = com_par.z;
#pragma unroll
    for (int i = 0; i < 7; ++i) st[(ST_AABB + i) * 32 + b] = L.aabb[i];
    st[(ST_QOFF + 0) * 32 + b] = L.js.qoff.x; st[(ST_QOFF + 1) * 32 + b] = L.js.qoff.y;
    st[(ST_QOFF + 2) * 32 + b] = L.js.qoff.z; st[(ST_QOFF + 3) * 32 + b] = L.js.qoff.w;
}
// per-thread parameters as five float4 quads [5][NT]: (inv_m, I0..I2) (I3..I6) (I7, I8, J0, J1) (J2..J5) (J6..J8, -);
// `asm volatile` loads: re-issued at the point of use in every substep, never hoisted into loop-long registers
__device__ __forceinline__ float4 lds128v(const float4* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ F3 st_vec3(const volatile float* st, int row, int b) {
    return v3<float>(st[row * 32 + b], st[(row + 1) * 32 + b], st[(row + 2) * 32 + b]);
}
// joint_X_p is per THREAD (two float4 quads [2][NT]: (x, y, z, qx) (qy, qz, qw, -)) because lab4d gives every
// environment its own bone lengths (dp_interface.py:454-465); everything else of the joint is per body
template <int NT, bool QOFF>
__device__ __forceinline__ JointStatic<float> st_joint(const volatile float* st, const float4* xpq, int b, int type) {
    JointStatic<float> js;
    js.type = type;
    const float4 a = lds128v(xpq), c = lds128v(xpq + NT);
    js.xpj = v3<float>(a.x, a.y, a.z);
    js.qpj = q4<float>(a.w, c.x, c.y, c.z);
    js.axis = st_vec3(st, ST_AXIS, b);
    js.qoff = QOFF ? q4<float>(st[(ST_QOFF + 0) * 32 + b], st[(ST_QOFF + 1) * 32 + b], st[(ST_QOFF + 2) * 32 + b],
                               st[(ST_QOFF + 3) * 32 + b])
                   : q4<float>(0.f, 0.f, 0.f, 1.f);
    return js;
}
template <int NT>
__device__ __forceinline__ void par_store(float4* par, float inv_m, const float* I, const float* J) {
    par[0 * NT] = make_float4(inv_m, I[0], I[1], I[2]);
    par[1 * NT] = make_float4(I[3], I[4], I[5], I[6]);
    par[2 * NT] = make_float4(I[7], I[8], J[0], J[1]);
    par[3 * NT] = make_float4(J[2], J[3], J[4], J[5]);
    par[4 * NT] = make_float4(J[6], J[7], J[8], 0.f);
}
template <int NT>
__device__ __forceinline__ void par_load(const float4* par, float& inv_m, float* I, float* J) {
    const float4 a = lds128v(par), b = lds128v(par + NT), c = lds128v(par + 2 * NT), d = lds128v(par + 3 * NT),
                 e = lds128v(par + 4 * NT);
    inv_m = a.x;
    I[0] = a.y; I[1] = a.z; I[2] = a.w; I[3] = b.x; I[4] = b.y; I[5] = b.z; I[6] = b.w; I[7] = c.x; I[8] = c.y;
    J[0] = c.z; J[1] = c.w; J[2] = d.x; J[3] = d.y; J[4] = d.z; J[5] = d.w; J[6] = e.x; J[7] = e.y; J[8] = e.z;
}

__device__ __forceinline__ ContactMat<float> load_mat(const DevModel& M, int k) {
    float4 m = M.mats[M.cmat[k]];
    ContactMat<float> c; c.ke = m.x; c.kd = m.y; c.kf = m.z; c.mu = m.w;
    return c;
}
__device__ __forceinline__ ContactMat<float> mat_of(const DevModel& M, const ContactMat<float>& cm0, int k) {
    return M.mat_uniform ? cm0 : load_mat(M, k);
}

// Phase A of K3 / K3^T: which contact points of which bodies can penetrate the ground plane.
//   small bodies (<= big_threshold points, e.g. the 8 box corners of human / quad): the owning lane will just
//     loop over its own points -> returns -2
//   big bodies (laikago's collision meshes, 96..596 vertices): the WHOLE WARP tests the body's points (only the
//     y-row of the rotation and the height are broadcast: 6 shuffles) and the penetrating point indices are
//     compacted into the owner lane's list in shared memory -> returns their count, or -1 if more than
//     PPR_CLIST_CAP penetrate (owner falls back to the cooperative evaluate-and-reduce path).
// The test uses a 1e-6 m margin; the exact `c > 0` rejection of the reference is re-applied per point in phase B.
//
// Second-level cull of a big body: a lower bound of m(a) = min_k a . p_k (a = body-frame image of the ground normal,
// the y-row of the rotation matrix) from a table of m sampled on a cube map of directions.  m is concave and positively
// homogeneous, so for a = sum_i w_i n_i with w_i >= 0 (the bilinear weights of the four nodes of a's cell, which
// reproduce the affine map (u, v) -> direction exactly) m(a) >= sum_i w_i m(n_i): the interpolated table value never
// exceeds the height of the lowest vertex.  The AABB bound is centimetres loose for a rotated foot mesh that hovers
// millimetres above the ground; this one is ~0.5 mm tight, so hovering feet skip the 96-vertex test.
__device__ __forceinline__ float support_lower_bound(const float* __restrict__ T, float m0, float m1, float m2) {
    const float ax = fabsf(m0), ay = fabsf(m1), az = fabsf(m2);
    int face; float d, u, v;
    if (ax >= ay && ax >= az) { face = m0 < 0.f ? 1 : 0; d = ax; u = m1; v = m2; }
    else if (ay >= az) { face = m1 < 0.f ? 3 : 2; d = ay; u = m0; v = m2; }
    else { face = m2 < 0.f ? 5 : 4; d = az; u = m0; v = m1; }
    const float inv = 1.f / fmaxf(d, 1e-30f);
    const float fu = fminf(fmaxf((u * inv + 1.f) * (0.5f * PPR_SUP_N), 0.f), (float)PPR_SUP_N);
    const float fv = fminf(fmaxf((v * inv + 1.f) * (0.5f * PPR_SUP_N), 0.f), (float)PPR_SUP_N);
    const int iu = min((int)fu, PPR_SUP_N - 1), iv = min((int)fv, PPR_SUP_N - 1);
    const float wu = fu - (float)iu, wv = fv - (float)iv;   // in [0, 1]
    const float* t = T + (face * (PPR_SUP_N + 1) + iv) * (PPR_SUP_N + 1) + iu;
    const float t00 = t[0], t01 = t[1], t10 = t[PPR_SUP_N + 1], t11 = t[PPR_SUP_N + 2];
    const float lo = (1.f - wv) * ((1.f - wu) * t00 + wu * t01) + wv * ((1.f - wu) * t10 + wu * t11);
    return d * lo;
}
template <bool SUP>   // SUP: apply the support-function cull (forward pass; the adjoint's rare overflow path skips it)
__device__ __forceinline__ int contact_candidates(const DevModel& M, const LaneInfo& L, int lane, const BodyF& s,
                                                  const volatile float* st, int* __restrict__ clist, float& m0,
                                                  float& m1, float& m2, bool todo = true) {
    float w = s.r.w, ux = s.r.x, uy = s.r.y, uz = s.r.z;
    m0 = 2.f * (w * uz + uy * ux); m1 = 2.f * w * w - 1.f + 2.f * uy * uy; m2 = 2.f * (uy * uz - w * ux);
    const int b = L.body;
    float ylow = s.x.y + fminf(m0 * st[(ST_AABB + 0) * 32 + b], m0 * st[(ST_AABB + 3) * 32 + b]) +
                 fminf(m1 * st[(ST_AABB + 1) * 32 + b], m1 * st[(ST_AABB + 4) * 32 + b]) +
                 fminf(m2 * st[(ST_AABB + 2) * 32 + b], m2 * st[(ST_AABB + 5) * 32 + b]) - st[(ST_AABB + 6) * 32 + b];
    bool maybe = L.valid && M.ground && (L.c1 > L.c0) && !(ylow > 1e-6f) && todo;   // todo: team layout, see TeamComm
    bool big = (L.c1 - L.c0) > M.big_threshold;
#ifndef PPR_NO_SUPPORT_CULL
    if (SUP && maybe && big) {   // (every big body has a table; the index is fetched here so that it costs no register)
        // |rounding of the interpolation| << 1e-6 for bodies of < 1 m; the table itself is rounded down on the host
        const float* T = M.sup + (size_t)M.sup_of[b] * PPR_SUP_FLOATS;
        const float ysup = s.x.y + support_lower_bound(T, m0, m1, m2) - st[(ST_AABB + 6) * 32 + b];
        maybe = !(ysup > 3e-6f);
    }
#endif
    int mine = (maybe && !big) ? -2 : 0;
    unsigned mask = __ballot_sync(FULL, maybe && big);
    if (mask == 0) return mine;
    const unsigned lt = (1u << lane) - 1u;
    while (mask) {
        int a = __ffs(mask) - 1;
        mask &= mask - 1;
        float a0 = shf(m0, a), a1 = shf(m1, a), a2 = shf(m2, a), ay = shf(s.x.y, a);
        int c0 = __shfl_sync(FULL, L.c0, a), c1 = __shfl_sync(FULL, L.c1, a);
        int cnt = 0;
        for (int base = c0; base < c1; base += 128) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // all loads of the trip in flight before the first use
                int k = base + u * 32 + lane;
                p[u] = (k < c1) ? M.cpt[k] : make_float4(0.f, 0.f, 0.f, -1e30f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int k = base + u * 32 + lane;
                float c = ay + a0 * p[u].x + a1 * p[u].y + a2 * p[u].z - p[u].w;
                bool pen = !(c > 1e-6f);
                unsigned bits = __ballot_sync(FULL, pen);
                if (pen) {
                    int pos = cnt + __popc(bits & lt);
                    if (pos < PPR_CLIST_CAP) clist[a * PPR_CLIST_STRIDE + pos] = k;
                }
                cnt += __popc(bits);
            }
        }
        if (lane == a) mine = cnt > PPR_CLIST_CAP ? -1 : cnt;
    }
    __syncwarp();
    return mine;
}

// record of the contact points that actually produced force in a substep (consumed by the adjoint kernel)
struct ContactRec {
    unsigned long long lo, hi; // 8 halfwords: [0] = number of active points (> PPR_REC_MAX = not recorded, the adjoint
                               // re-derives them), [1..7] = their indices
};
__device__ __forceinline__ unsigned rec_cnt(const ContactRec& r) { return (unsigned)r.lo & 0xffffu; }
__device__ __forceinline__ void rec_overflow(ContactRec& r) { r.lo = (r.lo & ~0xffffull) | (unsigned long long)(PPR_REC_MAX + 1); }
__device__ __forceinline__ void rec_push(ContactRec& r, int k) {
    const unsigned h = rec_cnt(r) + 1;   // halfword that takes the index
    if (h < 4) r.lo |= (unsigned long long)(unsigned)k << (16 * h);
    else if (h < 8) r.hi |= (unsigned long long)(unsigned)k << (16 * (h - 4));
    if (h <= PPR_REC_MAX + 1) r.lo += 1ull;   // the count saturates at PPR_REC_MAX + 1
}
__device__ __forceinline__ int rec_get(const ContactRec& r, unsigned i) {
    const unsigned h = i + 1;
    unsigned long long w = h < 4 ? r.lo : r.hi;
    return (int)((w >> (16 * (h & 3))) & 0xffffull);
}
// checkpoint row, part known BEFORE the child wrenches are gathered / part that needs the total wrench
template <int RQ> __device__ __forceinline__ void row_store_pre(float4* c, const BodyF& s, const float* ang, const ContactRec& rec) {
    c[0 * 32] = make_float4(s.x.x, s.x.y, s.x.z, s.r.x);
    c[1 * 32] = make_float4(s.r.y, s.r.z, s.r.w, s.w.x);
    c[2 * 32] = make_float4(s.w.y, s.w.z, s.v.x, s.v.y);
    c[3 * 32] = make_float4(s.v.z, ang[0], __uint_as_float((unsigned)rec.lo), __uint_as_float((unsigned)(rec.lo >> 32)));
    if (RQ == 7) c[4 * 32] = make_float4(__uint_as_float((unsigned)rec.hi), __uint_as_float((unsigned)(rec.hi >> 32)), ang[1], ang[2]);
}
template <int RQ> __device__ __forceinline__ void row_store_post(float4* c, const ContactRec& rec, const WrenchF& F) {
    if (RQ == 6) {
        c[4 * 32] = make_float4(__uint_as_float((unsigned)rec.hi), __uint_as_float((unsigned)(rec.hi >> 32)), F.t.x, F.t.y);
        c[5 * 32] = make_float4(F.t.z, F.f.x, F.f.y, F.f.z);
    } else {
        c[5 * 32] = make_float4(F.t.x, F.t.y, F.t.z, F.f.x);
        c[6 * 32] = make_float4(F.f.y, F.f.z, 0.f, 0.f);
    }
}
// the same row read back from this lane's quads in shared memory
template <int RQ> __device__ __forceinline__ void row_load(const float4* r, BodyF& s, WrenchF& F, ContactRec& rec, float* ang) {
    const float4 a0 = lds128v(r), a1 = lds128v(r + 32), a2 = lds128v(r + 64), a3 = lds128v(r + 96), a4 = lds128v(r + 128),
                 a5 = lds128v(r + 160);
    s.x = v3<float>(a0.x, a0.y, a0.z); s.r = q4<float>(a0.w, a1.x, a1.y, a1.z);
    s.w = v3<float>(a1.w, a2.x, a2.y); s.v = v3<float>(a2.z, a2.w, a3.x);
    ang[0] = a3.y;
    rec.lo = (unsigned long long)__float_as_uint(a3.z) | ((unsigned long long)__float_as_uint(a3.w) << 32);
    rec.hi = (unsigned long long)__float_as_uint(a4.x) | ((unsigned long long)__float_as_uint(a4.y) << 32);
    if (RQ == 6) {
        ang[1] = 0.f; ang[2] = 0.f;
        F.t = v3<float>(a4.z, a4.w, a5.x); F.f = v3<float>(a5.y, a5.z, a5.w);
    } else {
        const float4 a6 = lds128v(r + 192);
        ang[1] = a4.z; ang[2] = a4.w;
        F.t = v3<float>(a5.x, a5.y, a5.z); F.f = v3<float>(a5.w, a6.x, a6.y);
    }
}

// K3 for the whole warp in two phases (they sit on different sides of the state rendezvous in the forward kernel):
//   warp_contacts_search: which points penetrate (needs only this warp's body states)  -> cand / pen
//   warp_contacts_eval  : subtracts their contact wrenches from F (per lane = per body) and records the active ones
__device__ __forceinline__ int warp_contacts_search(const DevModel& M, const LaneInfo& L, int lane, const BodyF& s,
                                                    const volatile float* st, int* __restrict__ clist, unsigned& pen,
                                                    bool todo = true) {
    float m0, m1, m2;
    int cand = contact_candidates<true>(M, L, lane, s, st, clist, m0, m1, m2, todo);
    pen = 0;
    if (cand == -2) {
        // small body (<= big_threshold <= 32 points, e.g. 8 box corners): test all points with four loads in flight and
        // build a bitmask; the evaluation phase visits only the penetrating ones (one inlined copy of the force code)
        for (int k0 = L.c0; k0 < L.c1; k0 += 4) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = (k0 + u < L.c1) ? M.cpt[k0 + u] : make_float4(0.f, 0.f, 0.f, -1e30f);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (!(s.x.y + m0 * p[u].x + m1 * p[u].y + m2 * p[u].z - p[u].w > 1e-6f)) pen |= 1u << (k0 - L.c0 + u);
        }
    }
    return cand;
}
__device__ __forceinline__ void warp_contacts_eval(const DevModel& M, const LaneInfo& L, int lane, const BodyF& s,
                                                   const M3F& Rb, F3 xc, const ContactMat<float>& cm0, int cand, unsigned pen,
                                                   int* __restrict__ clist, WrenchF& F, ContactRec& rec) {
    rec.lo = 0ull; rec.hi = 0ull;
    const bool can_rec = M.nc <= 65535;
    if (cand == -2) {
        while (pen) {
            int k = L.c0 + __ffs(pen) - 1;
            pen &= pen - 1;
            float4 p = M.cpt[k];
            if (contact_point_fwd(s, Rb, xc, v3<float>(p.x, p.y, p.z), p.w, mat_of(M, cm0, k), F)) rec_push(rec, k);
        }
    } else {
        for (int i = 0; i < cand; ++i) {
            int k = clist[lane * PPR_CLIST_STRIDE + i];
            float4 p = M.cpt[k];
            if (contact_point_fwd(s, Rb, xc, v3<float>(p.x, p.y, p.z), p.w, mat_of(M, cm0, k), F)) rec_push(rec, k);
        }
    }
    if (cand == -1 || !can_rec) rec_overflow(rec);
    unsigned mask = __ballot_sync(FULL, cand == -1);
    while (mask) {  // many penetrating points: evaluate cooperatively and reduce
        int a = __ffs(mask) - 1;
        mask &= mask - 1;
        BodyF sa = shf_body(s, a);
        F3 xca = shf3(xc, a);
        int c0 = __shfl_sync(FULL, L.c0, a), c1 = __shfl_sync(FULL, L.c1, a);
        WrenchF W = wrench_zero<float>();
        const M3F Ra = qmat(sa.r);
        for (int k = c0 + lane; k < c1; k += 32) {
            float4 p = M.cpt[k];
            contact_point_fwd(sa, Ra, xca, v3<float>(p.x, p.y, p.z), p.w, mat_of(M, cm0, k), W);
        }
        W.t.x = warp_sum(W.t.x); W.t.y = warp_sum(W.t.y); W.t.z = warp_sum(W.t.z);
        W.f.x = warp_sum(W.f.x); W.f.y = warp_sum(W.f.y); W.f.z = warp_sum(W.f.z);
        if (lane == a) { F.t += W.t; F.f += W.f; }
    }
    __syncwarp();  // clist is reused by the next substep
}

// K3^T for the whole warp. Normally only replays the points the forward pass recorded as active; lanes whose
// record overflowed (> PPR_REC_MAX active points) re-derive them like the forward pass did.
__device__ __forceinline__ void warp_contacts_adj(const DevModel& M, const LaneInfo& L, int lane, const BodyF& s,
                                                  const M3F& Rb, F3 xc, const ContactMat<float>& cm0,
                                                  const volatile float* st, int* __restrict__ clist,
                                                  const ContactRec& rec, const WrenchF& adjF, BodyF& adjS, M3F& G,
                                                  F3& adj_xc) {
    const unsigned nrec = rec_cnt(rec);
    const bool ovf = L.valid && nrec > PPR_REC_MAX;
    if (!ovf && L.valid) {
        for (unsigned i = 0; i < nrec; ++i) {
            int k = rec_get(rec, i);
            float4 p = M.cpt[k];
            contact_point_adj(s, Rb, xc, v3<float>(p.x, p.y, p.z), p.w, mat_of(M, cm0, k), adjF, adjS, G, adj_xc);
        }
    }
    if (!__any_sync(FULL, ovf)) return;
    float m0, m1, m2;
    int cand = contact_candidates<false>(M, L, lane, s, st, clist, m0, m1, m2);
    if (!ovf) cand = 0;
    if (cand == -2) {
        for (int k = L.c0; k < L.c1; ++k) {
            float4 p = M.cpt[k];
            contact_point_adj(s, Rb, xc, v3<float>(p.x, p.y, p.z), p.w, mat_of(M, cm0, k), adjF, adjS, G, adj_xc);
        }
    } else {
        for (int i = 0; i < cand; ++i) {
            int k = clist[lane * PPR_CLIST_STRIDE + i];
            float4 p = M.cpt[k];
            contact_point_adj(s, Rb, xc, v3<float>(p.x, p.y, p.z), p.w, mat_of(M, cm0, k), adjF, adjS, G, adj_xc);
        }
    }
    unsigned mask = __ballot_sync(FULL, cand == -1);
    while (mask) {
        int a = __ffs(mask) - 1;
        mask &= mask - 1;
        BodyF sa = shf_body(s, a);
        F3 xca = shf3(xc, a);
        WrenchF aFa = shf_wrench(adjF, a);
        int c0 = __shfl_sync(FULL, L.c0, a), c1 = __shfl_sync(FULL, L.c1, a);
        BodyF A = body_zero<float>();
        F3 Axc = vzero<float>();
        const M3F Ra = qmat(sa.r);
        M3F Ga = m3_zero<float>();
        for (int k = c0 + lane; k < c1; k += 32) {
            float4 p = M.cpt[k];
            contact_point_adj(sa, Ra, xca, v3<float>(p.x, p.y, p.z), p.w, mat_of(M, cm0, k), aFa, A, Ga, Axc);
        }
        A.r += qmat_adj(sa.r, Ga);
        A.x.x = warp_sum(A.x.x); A.x.y = warp_sum(A.x.y); A.x.z = warp_sum(A.x.z);
        A.r.x = warp_sum(A.r.x); A.r.y = warp_sum(A.r.y); A.r.z = warp_sum(A.r.z); A.r.w = warp_sum(A.r.w);
        A.w.x = warp_sum(A.w.x); A.w.y = warp_sum(A.w.y); A.w.z = warp_sum(A.w.z);
        A.v.x = warp_sum(A.v.x); A.v.y = warp_sum(A.v.y); A.v.z = warp_sum(A.v.z);
        Axc.x = warp_sum(Axc.x); Axc.y = warp_sum(Axc.y); Axc.z = warp_sum(Axc.z);
        if (lane == a) { body_acc(adjS, A); adj_xc += Axc; }
    }
    __syncwarp();
}

// articulation FK across the warp, level by level (parents before children)
template <int JM, class Comm>
__device__ __forceinline__ BodyF warp_fk(const Comm& comm, const DevModel& M, const LaneInfo& L, const float* jq,
                                         const float* jqd) {
    BodyF s = body_identity<float>();
    for (int d = 0; d <= M.maxdepth; ++d) {
        BodyF P = comm.parent_body(s, L.parent_slot);
        if (!L.has_parent) P = body_identity<float>();
        if (L.depth == d) s = fk_joint_fwd<float, JM>(L.js, L.com, P, jq, jqd);
    }
    return s;
}

__device__ __forceinline__ void load_joint_coords(const LaneInfo& L, const float* q, const float* qd, int nq, int nqd,
                                                  float* jq, float* jqd) {
    int ncoord = L.type == JT_FREE ? 7 : L.ndof;
#pragma unroll
    for (int k = 0; k < 7; ++k) jq[k] = (k < ncoord) ? q[(int64_t)L.env * nq + L.qs + k] : 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) jqd[k] = (k < L.ndof) ? qd[(int64_t)L.env * nqd + L.qds + k] : 0.f;
}

// ----------------------------------------------------------------------------------------------- FK kernels
template <int JM>
__global__ void __launch_bounds__(PPR_BLOCK)
fk_forward_kernel(DevModel M, int64_t n, const float* __restrict__ q, const float* __restrict__ qd,
                  float* __restrict__ body_q, float* __restrict__ body_qd) {
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp * M.epw >= n) return;
    const WarpComm<PPR_BLOCK> comm(nullptr);
    LaneInfo L = lane_setup(M, warp, lane, n, M.epw);
    float jq[7], jqd[6];
    load_joint_coords(L, q, qd, M.nq, M.nqd, jq, jqd);
    BodyF s = warp_fk<JM>(comm, M, L, jq, jqd);
    if (L.valid) {
        float* o = body_q + ((int64_t)L.env * M.nb + L.body) * 7;
        o[0] = s.x.x; o[1] = s.x.y; o[2] = s.x.z; o[3] = s.r.x; o[4] = s.r.y; o[5] = s.r.z; o[6] = s.r.w;
        float* v = body_qd + ((int64_t)L.env * M.nb + L.body) * 6;
        v[0] = s.w.x; v[1] = s.w.y; v[2] = s.w.z; v[3] = s.v.x; v[4] = s.v.y; v[5] = s.v.z;
    }
}

// shared by fk_backward_kernel and the tail of rollout_backward_kernel
template <int JM, class Comm>
__device__ __forceinline__ void warp_fk_adjoint(const Comm& comm, const DevModel& M, const LaneInfo& L, const BodyF& s,
                                                BodyF adj, const float* jq, const float* jqd,
                                                float* __restrict__ adj_q, float* __restrict__ adj_qd) {
    float ajq[7] = {0, 0, 0, 0, 0, 0, 0}, ajqd[6] = {0, 0, 0, 0, 0, 0};
    for (int d = M.maxdepth; d >= 0; --d) {
        BodyF P = comm.parent_body(s, L.parent_slot);
        if (!L.has_parent) P = body_identity<float>();
        BodyF adjP = body_zero<float>();
        if (L.depth == d) fk_joint_adj<float, JM>(L.js, L.com, P, jq, jqd, adj, adjP, ajq, ajqd);
        comm.gather_body_sync(adjP, L.child, M.maxc, adj);
    }
    if (L.valid) {
        int ncoord = L.type == JT_FREE ? 7 : L.ndof;
#pragma unroll
        for (int k = 0; k < 7; ++k) if (k < ncoord) adj_q[(int64_t)L.env * M.nq + L.qs + k] = nan0(ajq[k]);
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < L.ndof) adj_qd[(int64_t)L.env * M.nqd + L.qds + k] = nan0(ajqd[k]);
    }
}

template <int JM>
__global__ void __launch_bounds__(PPR_BLOCK)
fk_backward_kernel(DevModel M, int64_t n, const float* __restrict__ q, const float* __restrict__ qd,
                   const float* __restrict__ adj_body_q, const float* __restrict__ adj_body_qd,
                   float* __restrict__ adj_q, float* __restrict__ adj_qd) {
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp * M.epw >= n) return;
    const WarpComm<PPR_BLOCK> comm(nullptr);
    LaneInfo L = lane_setup(M, warp, lane, n, M.epw);
    float jq[7], jqd[6];
    load_joint_coords(L, q, qd, M.nq, M.nqd, jq, jqd);
    BodyF s = warp_fk<JM>(comm, M, L, jq, jqd);
    BodyF adj = body_zero<float>();
    if (L.valid) {
        const float* a = adj_body_q + ((int64_t)L.env * M.nb + L.body) * 7;
        const float* b = adj_body_qd + ((int64_t)L.env * M.nb + L.body) * 6;
        adj.x = v3<float>(a[0], a[1], a[2]); adj.r = q4<float>(a[3], a[4], a[5], a[6]);
        adj.w = v3<float>(b[0], b[1], b[2]); adj.v = v3<float>(b[3], b[4], b[5]);
    }
    warp_fk_adjoint<JM>(comm, M, L, s, adj, jq, jqd, adj_q, adj_qd);
}

// ----------------------------------------------------------------------------------------------- rollout
struct RolloutArgs {
    int64_t bs, nsteps, stride, nwarps, ngroups;
    int64_t ckpt_every;  // K: a checkpoint row is kept every K substeps; the adjoint recomputes the K-1 in between
    float dt;
    int pstride;  // 1: per-env parameter arrays, 0: one shared copy
    const float *q_init, *qd_init, *torques, *res_f, *refs, *ke, *kd, *inv_m, *I, *inv_I;
    float *out_pos, *out_vel, *out_grf, *out_jaf;
    float* ckpt;
    // backward only
    const float *adj_pos, *adj_vel;
    float *adj_q_init, *adj_qd_init, *adj_torques, *adj_res_f, *adj_refs, *adj_ke, *adj_kd, *adj_inv_m, *adj_I,
        *adj_inv_I;
    // shared-parameter mode with in-kernel reduction: one row of P = 2 nqd + 19 nb floats per thread block
    // ([ke nqd][kd nqd][inv_m nb][I nb*9][inv_I nb*9], summed over the block's environments in a fixed order);
    // reduce_partials_kernel then sums the rows.  When set, the five per-env outputs above are not written.
    float* adj_partial;
    // fused pose loss at the frame steps (se3_loss of dp_utils.py:113-138 on sim vs target poses, dp_model.py:777):
    // forward writes loss_pos[F, bs*nb]; the adjoint seeds itself from adj_loss_pos[F, bs*nb], the target poses and the
    // frame poses the forward pass wrote (saved_pos = out_pos) -- no adj_pos tensor, no separate loss kernels
    const float* target_pos;
    float* loss_pos;
    const float* adj_loss_pos;
    const float* saved_pos;
    float* adj_target_pos;     // optional: d objective / d target_pos [F, bs*nb, 7] (the targets depend on global_q)
    float rot_ratio;
};

__device__ __forceinline__ void load_ctl(const DevModel& M, const LaneInfo& L, const RolloutArgs& A, int64_t t,
                                         const float* ke, const float* kd, JointCtl<float>& c) {
    int64_t row = (t * A.bs + L.env) * M.nqd + L.qds;
    bool on = L.type != JT_FREE;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        bool use = on && k < L.ndof;
        c.target[k] = use ? A.refs[row + k] : 0.f;
        c.act[k] = (use && A.torques) ? A.torques[row + k] : 0.f;
        c.ke[k] = ke[k]; c.kd[k] = kd[k];
    }
}

__device__ __forceinline__ void store_wrench_row(float* base, const WrenchF& w) {
    base[0] = w.t.x; base[1] = w.t.y; base[2] = w.t.z; base[3] = w.f.x; base[4] = w.f.y; base[5] = w.f.z;
}

// Fused pose loss at a frame step (ppr_loss.h).  Out of line on purpose: it runs once per frame, not per substep, and
// the time loops of the COMPOUND adjoint already fill the instruction cache.
__device__ __noinline__ float frame_pose_loss(const float* pose, const float* target, float ratio) {
    const float pp[7] = {pose[0], pose[1], pose[2], pose[3], pose[4], pose[5], pose[6]};
    const float gg[7] = {target[0], target[1], target[2], target[3], target[4], target[5], target[6]};
    return se3_pair_loss<float>(7, pp, gg, ratio, 1e-4f);
}
__device__ __noinline__ void frame_pose_loss_adj(const float* pose, const float* target, float ratio, float adj, float* ap,
                                                 float* adj_target) {
    const float pp[7] = {pose[0], pose[1], pose[2], pose[3], pose[4], pose[5], pose[6]};
    const float gg[7] = {target[0], target[1], target[2], target[3], target[4], target[5], target[6]};
    float ag[7];
    se3_pair_loss_adj<float>(7, pp, gg, ratio, 1e-4f, adj, ap, adj_target ? ag : (float*)nullptr);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        ap[k] = nan0(ap[k]);
        if (adj_target) adj_target[k] = nan0(ag[k]);
    }
}

// Forces of one substep in two halves.  forces_pre: contacts + this body's joint, ends by PUBLISHING the wrench the
// joint exerts on the parent; forces_post: adds the wrenches of the body's children -> F = total wrench on this lane's
// body.  Work that does not need the children's wrenches goes between the two (split-phase rendezvous).
// Optionally exports the grf / jaf side channels.
template <int JM, bool LIMITS, bool QOFF, class Comm>
__device__ __forceinline__ WrenchF forces_pre(Comm& comm, const DevModel& M, const LaneInfo& L, int lane, const BodyF& s,
                                              const M3F& Rb, F3 xc, const JointCtl<float>& ctl, const ContactMat<float>& cm0,
                                              const volatile float* st, const float4* xpq, int* clist, const float* res_f_row,
                                              float* grf_row, WrenchF& F, WrenchF& G, ContactRec& rec, float* ang) {
    unsigned pen = 0;
    int cand = 0;
    if constexpr (Comm::kHelpers > 0) comm.request_contacts(s, xc);   // team layout: the helper warps evaluate the contacts
    else {
        comm.post_state(s, xc);   // published before the (warp-dependent) contact search, awaited after it
        cand = warp_contacts_search(M, L, lane, s, st, clist, pen);
    }
    // joints: this lane is the child of its joint
    BodyF P;
    F3 xcp;
    comm.get_parent_state(s, xc, L.parent_slot, P, xcp);
    if (!L.has_parent) { P = body_identity<float>(); xcp = vzero<float>(); }
    F3 t, f, ap, ac;
    joint_fwd<float, JM, LIMITS, QOFF>(st_joint<Comm::kThreads, QOFF>(st, xpq, L.body, L.type), ctl, M.ake, M.akd, P, xcp, L.has_parent, s,
                                       Rb, xc, t, f, ap, ac, ang);
    WrenchF Wp = wrench_zero<float>();
    if (L.type != JT_FREE && L.has_parent) { Wp.t = t + cross(ap, f); Wp.f = f; }
#ifndef PPR_CONTACTS_EARLY
    // The wrench on the parent does not depend on this body's contacts: publish it FIRST, so that the parent is not held
    // up by the evaluation of the contact points (split-phase rendezvous B; the evaluation overlaps its skew).
    comm.post_wrench(Wp);
#endif
    F = wrench_zero<float>();
    if (res_f_row && L.valid) {
        F.t = v3<float>(res_f_row[0], res_f_row[1], res_f_row[2]);
        F.f = v3<float>(res_f_row[3], res_f_row[4], res_f_row[5]);
    }
    if constexpr (Comm::kHelpers > 0) {
        WrenchF Fc;
        comm.await_contacts(Fc, rec.lo, rec.hi);
        F.t += Fc.t; F.f += Fc.f;
    } else warp_contacts_eval(M, L, lane, s, Rb, xc, cm0, cand, pen, clist, F, rec);
    G = F;
    if (grf_row && L.valid) store_wrench_row(grf_row, F);
    if (L.type != JT_FREE) { F.t -= t + cross(ac, f); F.f -= f; }
#ifdef PPR_CONTACTS_EARLY
    comm.post_wrench(Wp);
#endif
    return Wp;
}
template <class Comm>
__device__ __forceinline__ void forces_post(Comm& comm, const LaneInfo& L, const WrenchF& Wp, const WrenchF& G,
                                            float* jaf_row, WrenchF& F) {
    comm.gather_wrench(Wp, L.child, L.maxc_w, F);
    if (jaf_row && L.valid) {
        WrenchF J; J.t = F.t - G.t; J.f = F.f - G.f;
        store_wrench_row(jaf_row, J);
    }
}

// Resident blocks per SM asked of ptxas: forward <= 128 registers/thread, adjoint <= 168. Measured on B200
// (profiles/README.md): issue-slot utilisation rises with resident warps; going further costs spills.
#ifndef PPR_FWD_THREADS_PER_SM
#define PPR_FWD_THREADS_PER_SM 512
#endif
#ifndef PPR_BWD_THREADS_PER_SM
#define PPR_BWD_THREADS_PER_SM 390
#endif
#define PPR_FWD_MINB(NT) (PPR_FWD_THREADS_PER_SM / (NT))
#define PPR_BWD_MINB(NT) (PPR_BWD_THREADS_PER_SM / (NT))

// dynamic shared-memory layout (floats): [rows (adjoint only)] [static table] [params] [acc (adjoint only)] [clist] [comm]
template <class Comm, bool ADJ> struct SmemLayout {
    static constexpr int NT = Comm::kThreads, NW = NT / 32;
    static constexpr int row = 0;
    static constexpr int st = row + (ADJ ? NW * PPR_ROW_QUADS_MAX * 4 * 32 : 0);
    static constexpr int par = st + PPR_NSTATIC * 32;
    static constexpr int xpq = par + PPR_NPAR * NT;
    static constexpr int acc = xpq + 8 * NT;
    static constexpr int clist = acc + (ADJ ? 20 * NT : 0);
    static constexpr int comm = clist + NW * 32 * PPR_CLIST_STRIDE;
    static constexpr int rowbar = comm + Comm::kExFloats + ((Comm::kExFloats & 1) ? 1 : 0);   // NW 8-byte mbarriers
    static constexpr int sbody = rowbar + (ADJ ? 2 * NW : 0);        // adjoint epilogue: body of every thread (-1: none)
    static constexpr int total = sbody + (ADJ ? NT : 0);
    static constexpr size_t bytes = (size_t)total * sizeof(float);
    static_assert(comm % 4 == 0 && par % 4 == 0 && acc % 4 == 0, "float4 areas must be 16-byte aligned");
    static_assert(rowbar % 2 == 0, "mbarriers must be 8-byte aligned");
};

template <class Comm, int JM, bool LIMITS, bool QOFF>
__global__ void __launch_bounds__(Comm::kThreads, PPR_FWD_MINB(Comm::kThreads))
rollout_forward_kernel(DevModel M, RolloutArgs A) {
    constexpr int NT = Comm::kThreads;
    typedef SmemLayout<Comm, false> SL;
    extern __shared__ __align__(16) float smem[];
    Comm comm(smem + SL::comm);
    const int64_t group = Comm::group();
    // global warp: owns checkpoint rows (team layout: one row per block, written by its main warp)
    const int64_t warp = Comm::kHelpers > 0 ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    int* clist = (int*)(smem + SL::clist) + (threadIdx.x >> 5) * 32 * PPR_CLIST_STRIDE;
    volatile float* st = smem + SL::st;
    float4* par = (float4*)(smem + SL::par) + threadIdx.x;
    float4* xpq = (float4*)(smem + SL::xpq) + threadIdx.x;
    if (group >= A.ngroups) return;
    comm.init();
    ContactMat<float> cm0 = {0.f, 0.f, 0.f, 0.f};
    if (M.nc > 0) cm0 = load_mat(M, 0);
    LaneInfo L = lane_setup(M, group, Comm::slot(), A.bs, Comm::envs_per_group(M), Comm::kBlock && PPR_BODY_MAJOR);
    if constexpr (Comm::kHelpers > 0) comm.assign(L.valid && L.c1 > L.c0);
    // per-env parameters of this body / joint -> shared memory
    int64_t ebp = (int64_t)L.env * A.pstride * M.nb + L.body;
    par_store<NT>(par, A.inv_m[ebp], A.I + ebp * 9, A.inv_I + ebp * 9);
    xpq[0] = make_float4(L.js.xpj.x, L.js.xpj.y, L.js.xpj.z, L.js.qpj.x);
    xpq[NT] = make_float4(L.js.qpj.y, L.js.qpj.z, L.js.qpj.w, 0.f);
    JointCtl<float> ctl;
    float ke[3], kd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        bool use = L.type != JT_FREE && k < L.ndof;
        int64_t d = (int64_t)L.env * A.pstride * M.nqd + L.qds + k;
        ke[k] = use ? A.ke[d] : 0.f; kd[k] = use ? A.kd[d] : 0.f;
        float4 lm = use ? M.lim[L.qds + k] : make_float4(-1e30f, 1e30f, 0.f, 0.f);
        ctl.lo[k] = lm.x; ctl.hi[k] = lm.y; ctl.lke[k] = lm.z; ctl.lkd[k] = lm.w;
    }
    F3 g = v3<float>(M.g[0], M.g[1], M.g[2]);
    BodyF s;
    {
        float jq[7], jqd[6];
        load_joint_coords(L, A.q_init, A.qd_init, M.nq, M.nqd, jq, jqd);
        s = warp_fk<JM>(comm, M, L, jq, jqd);
        stage_static(st, L, comm.parent_vec(L.com, L.parent_slot));
    }
    __syncthreads();  // the static table is per block (all warps stage identical values)

    if constexpr (Comm::kHelpers > 0) {
        if (Comm::role() > 0) {
            // ---- helper warp of the team layout: the ground contacts of its share of the bodies, one reply per substep
            // (same trip count as the main warp's loop below)
            const int64_t nsub = (A.out_grf || A.out_jaf) ? A.nsteps : A.nsteps - 1;
            const bool todo = comm.mine();
            for (int64_t t = 0; t < nsub; ++t) {
                BodyF hs;
                F3 hxc;
                comm.await_request(hs, hxc);
                const M3F hR = qmat(hs.r);
                unsigned pen;
                const int cand = warp_contacts_search(M, L, lane, hs, st, clist, pen, todo);
                WrenchF Fc = wrench_zero<float>();
                ContactRec rec;
                warp_contacts_eval(M, L, lane, hs, hR, hxc, cm0, cand, pen, clist, Fc, rec);
                comm.reply_contacts(Fc, rec.lo, rec.hi);
            }
            return;
        }
    }
    constexpr int RQ = RowOf<JM>::kQuads;
    float4* ck = (float4*)A.ckpt + (warp * RQ) * 32 + lane;
    const int64_t ck_step = A.nwarps * RQ * 32;   // in float4
    // frame / checkpoint phases are carried as counters (no 64-bit divisions inside the time loop)
    int64_t fi = -1, fphase = 0, kphase = 0;
    for (int64_t t = 0; t < A.nsteps; ++t) {
        const bool frame = fphase == 0;
        fi += frame ? 1 : 0;
        if (++fphase == A.stride) fphase = 0;
        int64_t frow = (fi * A.bs + L.env) * M.nb + L.body;
        if (frame && L.valid) {
            float* o = A.out_pos + frow * 7;
            o[0] = s.x.x; o[1] = s.x.y; o[2] = s.x.z; o[3] = s.r.x; o[4] = s.r.y; o[5] = s.r.z; o[6] = s.r.w;
            float* v = A.out_vel + frow * 6;
            v[0] = s.w.x; v[1] = s.w.y; v[2] = s.w.z; v[3] = s.v.x; v[4] = s.v.y; v[5] = s.v.z;
            if (A.loss_pos) A.loss_pos[frow] = frame_pose_loss(o, A.target_pos + frow * 7, A.rot_ratio);
        }
        // the substep past the last frame exists only for the force side channels (dp_model.py:397): skip it
        // when nobody asked for them
        if (t == A.nsteps - 1 && !A.out_grf && !A.out_jaf) break;
        const F3 com = st_vec3(st, ST_COM, L.body);
        const M3F Rb = qmat(s.r);
        F3 xc = s.x + mrot(Rb, com);
        load_ctl(M, L, A, t, ke, kd, ctl);
        WrenchF F, Fg;
        ContactRec rec;
        float ang[3];
        const WrenchF Wp = forces_pre<JM, LIMITS, QOFF>(comm, M, L, lane, s, Rb, xc, ctl, cm0, st, xpq, clist,
                    A.res_f ? A.res_f + ((t * A.bs + L.env) * M.nb + L.body) * 6 : nullptr,
                    (frame && A.out_grf) ? A.out_grf + frow * 6 : nullptr, F, Fg, rec, ang);
        // ---- between publishing the parent wrench and gathering the children's: everything that does not need them
        // checkpoint (every K-th substep): state, joint angles, contact record
        const bool keep = kphase == 0;
        if (++kphase == A.ckpt_every) kphase = 0;
        float4* c = ck;
        if (keep) { ck += ck_step; row_store_pre<RQ>(c, s, ang, rec); }
        float inv_m, I[9], inv_I[9];
        par_load<NT>(par, inv_m, I, inv_I);
        const IntegratePre<float> pre = integrate_pre(s, Rb, I);
        // ---- total wrench
        forces_post(comm, L, Wp, Fg, (frame && A.out_jaf) ? A.out_jaf + frow * 6 : nullptr, F);
        if (keep) row_store_post<RQ>(c, rec, F);
        s = integrate_post(s, Rb, xc, com, F, inv_m, inv_I, g, A.dt, pre);
    }
}

// RECOMP = false: every substep has a checkpoint row (K = 1).  RECOMP = true: rows exist every K substeps; before a
// segment of K substeps is differentiated its rows are RE-COMPUTED from the segment's first state with the forward
// step function and parked in a per-warp scratch area of the workspace (K rows), then consumed exactly like
// stored rows.  Costs (K-1)/K of a forward pass, shrinks the checkpoint K-fold.
template <class Comm, int JM, bool LIMITS, bool QOFF, bool RECOMP>
__global__ void __launch_bounds__(Comm::kThreads, PPR_BWD_MINB(Comm::kThreads))
rollout_backward_kernel(DevModel M, RolloutArgs A) {
    constexpr int NT = Comm::kThreads;
    typedef SmemLayout<Comm, true> SL;
    extern __shared__ __align__(16) float smem[];
    Comm comm(smem + SL::comm);
    const int64_t group = Comm::group();
    const int64_t warp = Comm::kHelpers > 0 ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    int* clist = (int*)(smem + SL::clist) + (threadIdx.x >> 5) * 32 * PPR_CLIST_STRIDE;
    volatile float* st = smem + SL::st;
    float4* par = (float4*)(smem + SL::par) + threadIdx.x;
    float4* xpq = (float4*)(smem + SL::xpq) + threadIdx.x;
    float4* acc = (float4*)(smem + SL::acc) + threadIdx.x;
    constexpr int RQ = RowOf<JM>::kQuads;
    constexpr int ROWF = RQ * 4 * 32;    // floats per checkpoint row of a warp
    ((int*)(smem + SL::sbody))[threadIdx.x] = -1;   // (threads of a partially filled last block leave right below)
    volatile float* roww = smem + SL::row + (threadIdx.x >> 5) * ROWF;  // this warp's row buffer
    const float4* row4 = (const float4*)(smem + SL::row + (threadIdx.x >> 5) * ROWF) + lane;   // this lane's quads
    if (group >= A.ngroups) return;
    comm.init();
    ContactMat<float> cm0 = {0.f, 0.f, 0.f, 0.f};
    if (M.nc > 0) cm0 = load_mat(M, 0);
    LaneInfo L = lane_setup(M, group, Comm::slot(), A.bs, Comm::envs_per_group(M), Comm::kBlock && PPR_BODY_MAJOR);
    if constexpr (Comm::kHelpers > 0) comm.assign(L.valid && L.c1 > L.c0);
    int64_t eb = (int64_t)L.env * M.nb + L.body;
    int64_t ebp = (int64_t)L.env * A.pstride * M.nb + L.body;
    par_store<NT>(par, A.inv_m[ebp], A.I + ebp * 9, A.inv_I + ebp * 9);
    xpq[0] = make_float4(L.js.xpj.x, L.js.xpj.y, L.js.xpj.z, L.js.qpj.x);
    xpq[NT] = make_float4(L.js.qpj.y, L.js.qpj.z, L.js.qpj.w, 0.f);
#pragma unroll
    for (int i = 0; i < 5; ++i) acc[i * NT] = make_float4(0.f, 0.f, 0.f, 0.f);
    JointCtl<float> ctl;
    float ke[3], kd[3];
    const bool jon = L.type != JT_FREE;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        bool use = jon && k < L.ndof;
        int64_t d = (int64_t)L.env * A.pstride * M.nqd + L.qds + k;
        ke[k] = use ? A.ke[d] : 0.f; kd[k] = use ? A.kd[d] : 0.f;
        float4 lm = use ? M.lim[L.qds + k] : make_float4(-1e30f, 1e30f, 0.f, 0.f);
        ctl.lo[k] = lm.x; ctl.hi[k] = lm.y; ctl.lke[k] = lm.z; ctl.lkd[k] = lm.w;
    }
    F3 g = v3<float>(M.g[0], M.g[1], M.g[2]);
    // com of the parent body (to fold the parent's world-COM adjoint into its pose adjoint in the child lane)
    stage_static(st, L, comm.parent_vec(L.com, L.parent_slot));
    __syncthreads();  // the static table is per block (all warps stage identical values)

    float a_inv_m = 0.f, a_ke[3] = {0, 0, 0}, a_kd[3] = {0, 0, 0};

    const float* ckw = A.ckpt + warp * ROWF;  // this warp's rows (3 / 3.5 kB each, 16-byte aligned)
    const int64_t ck_step = A.nwarps * ROWF;
    const int64_t last = A.nsteps - 1;
    const int64_t K = RECOMP ? A.ckpt_every : 1;
    // scratch rows of this warp (RECOMP): behind the ceil(nsteps / K) stored rows of all warps
    float* scr = A.ckpt + ((A.nsteps + K - 1) / K) * ck_step + warp * K * ROWF;
    bool prefetched = false;
    // -DPPR_TMA_ROWS: fetch the rows with the bulk-copy engine.  Measured on B200 (profiles/README.md, round 2): 2.5 %
    // SLOWER than per-lane cp.async for these 3 kB per-warp rows (one more warp rendezvous + mbarrier polling per
    // substep, and no load instruction saved that mattered), so the default stays cp.async.
#ifdef PPR_TMA_ROWS
    constexpr bool BULK = !RECOMP;
#else
    constexpr bool BULK = false;
#endif
    RowStream<RQ, BULK> rows(roww, (unsigned long long*)(smem + SL::rowbar) + (threadIdx.x >> 5), lane);

    if constexpr (Comm::kHelpers > 0 && !RECOMP) {
        if (Comm::role() > 0) {
            // ---- helper warp of the team layout: K3^T of its share of the bodies.  It streams the block's checkpoint
            // rows itself (state + active-contact record), so that only the wrench adjoint has to come from the main warp.
            const bool todo = comm.mine();
            for (int64_t tp = last - 1; tp >= 0; --tp) {
                if (!prefetched) rows.issue(ckw + tp * ck_step);
                rows.wait();
                BodyF hs;
                WrenchF hF;
                ContactRec rec;
                float hang[3];
                row_load<RQ>(row4, hs, hF, rec, hang);
                prefetched = tp > 0;
                if (prefetched) rows.issue(ckw + (tp - 1) * ck_step);
                if (!todo) { rec.lo = 0ull; rec.hi = 0ull; }
                const M3F hR = qmat(hs.r);
                const F3 hxc = hs.x + mrot(hR, st_vec3(st, ST_COM, L.body));
                WrenchF aF;
                comm.await_request_adj(aF);
                BodyF aS = body_zero<float>();
                M3F hG = m3_zero<float>();
                F3 axc = vzero<float>();
                warp_contacts_adj(M, L, lane, hs, hR, hxc, cm0, st, clist, rec, aF, aS, hG, axc);
                comm.reply_contacts_adj(aS, hG, axc);
            }
            return;
        }
    }
    BodyF adjN = body_zero<float>();
    // rows of the never-differentiated last substep (dp_model.py:397): zero gradient
    if (L.valid) {
        int64_t row = (last * A.bs + L.env) * M.nqd + L.qds;
        for (int k = 0; k < L.ndof; ++k) {
            A.adj_refs[row + k] = 0.f;
            if (A.adj_torques) A.adj_torques[row + k] = 0.f;
        }
        if (A.adj_res_f) {
            float* r = A.adj_res_f + ((last * A.bs + L.env) * M.nb + L.body) * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) r[k] = 0.f;
        }
    }
    int64_t fi = last / A.stride, fphase = last % A.stride;   // frame index / phase of substep t, counted down
    for (int64_t t = last; t >= 0; --t) {
        // seed from the loss at frame steps
        const bool frame = fphase == 0;
        const int64_t fcur = fi;
        if (frame) { fphase = A.stride - 1; --fi; } else --fphase;
        if (frame && L.valid) {
            int64_t frow = (fcur * A.bs + L.env) * M.nb + L.body;
            if (A.adj_pos) {
                const float* a = A.adj_pos + frow * 7;
                adjN.x += v3<float>(a[0], a[1], a[2]);
                adjN.r += q4<float>(a[3], a[4], a[5], a[6]);
            }
            if (A.adj_vel) {
                const float* b = A.adj_vel + frow * 6;
                adjN.w += v3<float>(b[0], b[1], b[2]);
                adjN.v += v3<float>(b[3], b[4], b[5]);
            }
            if (A.adj_loss_pos) {   // fused pose loss: d loss / d pose from the saved frame pose and the target
                float ap[7];
                frame_pose_loss_adj(A.saved_pos + frow * 7, A.target_pos + frow * 7, A.rot_ratio, A.adj_loss_pos[frow], ap,
                                    A.adj_target_pos ? A.adj_target_pos + frow * 7 : nullptr);
                adjN.x += v3<float>(ap[0], ap[1], ap[2]);
                adjN.r += q4<float>(ap[3], ap[4], ap[5], ap[6]);
            }
        }
        if (t == 0) break;
        int64_t tp = t - 1;  // differentiate substep tp -> t
        const int64_t seg_lo = (tp / K) * K;   // first substep of the segment tp belongs to (K = 1: tp itself)
        if (RECOMP && (tp == last - 1 || tp % K == K - 1)) {
            // ---- re-run substeps seg_lo .. tp from the stored state and park their rows in the scratch area
            // the stored state goes through the row buffer (nothing is in flight there: prefetched == false here)
            rows.issue(ckw + (seg_lo / K) * ck_step);
            rows.wait();
            BodyF sr;
            {
                WrenchF Fd; ContactRec rd; float ad[3];
                row_load<RQ>(row4, sr, Fd, rd, ad);
            }
            for (int64_t tr = seg_lo; tr <= tp; ++tr) {
                const F3 comr = st_vec3(st, ST_COM, L.body);
                const M3F Rr = qmat(sr.r);
                F3 xcr = sr.x + mrot(Rr, comr);
                load_ctl(M, L, A, tr, ke, kd, ctl);
                WrenchF Fr, Fgr;
                ContactRec recr;
                float angr[3];
                const WrenchF Wpr = forces_pre<JM, LIMITS, QOFF>(comm, M, L, lane, sr, Rr, xcr, ctl, cm0, st, xpq, clist,
                                              A.res_f ? A.res_f + ((tr * A.bs + L.env) * M.nb + L.body) * 6 : nullptr,
                                              nullptr, Fr, Fgr, recr, angr);
                float4* c = (float4*)(scr + (tr - seg_lo) * ROWF) + lane;
                row_store_pre<RQ>(c, sr, angr, recr);
                forces_post(comm, L, Wpr, Fgr, nullptr, Fr);
                row_store_post<RQ>(c, recr, Fr);
                if (tr < tp) {
                    float inv_m, I[9], inv_I[9];
                    par_load<NT>(par, inv_m, I, inv_I);
                    sr = integrate_fwd(sr, Rr, xcr, comr, Fr, inv_m, I, inv_I, g, A.dt);
                }
            }
            __threadfence_block();   // the rows are read back by the same lanes through cp.async
            prefetched = false;
        }
        const float* rowsrc = RECOMP ? scr + (tp - seg_lo) * ROWF : ckw + tp * ck_step;
        if (!prefetched) rows.issue(rowsrc);   // first row (of the segment): nothing was prefetched yet
        rows.wait();
        BodyF s;
        WrenchF F;
        ContactRec rec;
        float ang[3];
        row_load<RQ>(row4, s, F, rec, ang);
        // refill the buffer with the next (earlier) row of the segment
        prefetched = RECOMP ? (tp > seg_lo) : (tp > 0);
        if (prefetched) rows.issue(RECOMP ? rowsrc - ROWF : ckw + (tp - 1) * ck_step);
        const F3 com = st_vec3(st, ST_COM, L.body);
        const M3F Rb = qmat(s.r);
        M3F G = m3_zero<float>();   // dL/dRb, converted to the quaternion adjoint once at the end of the substep
        F3 xc = s.x + mrot(Rb, com);
        load_ctl(M, L, A, tp, ke, kd, ctl);
        // K5^T
        BodyF adjS = body_zero<float>();
        F3 adj_xc = vzero<float>();
        WrenchF adjF;
        {
            float inv_m, I[9], inv_I[9];
            par_load<NT>(par, inv_m, I, inv_I);
            F3 ga, gb, gc, gd;
            integrate_adj_core(s, Rb, xc, com, F, inv_m, I, inv_I, g, A.dt, adjN, adjS, G, adj_xc, adjF, a_inv_m, ga,
                               gb, gc, gd);
            // adj_I += ga gb^T, adj_inv_I += gc gd^T: 18 accumulators as five float4 quads in shared memory
            float4 q0 = acc[0 * NT], q1 = acc[1 * NT], q2 = acc[2 * NT], q3 = acc[3 * NT], q4 = acc[4 * NT];
            q0.x += ga.x * gb.x; q0.y += ga.x * gb.y; q0.z += ga.x * gb.z; q0.w += ga.y * gb.x;
            q1.x += ga.y * gb.y; q1.y += ga.y * gb.z; q1.z += ga.z * gb.x; q1.w += ga.z * gb.y;
            q2.x += ga.z * gb.z; q2.y += gc.x * gd.x; q2.z += gc.x * gd.y; q2.w += gc.x * gd.z;
            q3.x += gc.y * gd.x; q3.y += gc.y * gd.y; q3.z += gc.y * gd.z; q3.w += gc.z * gd.x;
            q4.x += gc.z * gd.y; q4.y += gc.z * gd.z;
            acc[0 * NT] = q0; acc[1 * NT] = q1; acc[2 * NT] = q2; acc[3 * NT] = q3; acc[4 * NT] = q4;
        }
        if constexpr (Comm::kHelpers > 0) comm.request_contacts_adj(adjF);   // team layout: helpers replay the contacts
        comm.post_state_w(s, xc, adjF);   // published before the contact replay, awaited after it
#ifdef PPR_CONTACTS_EARLY
        // K3^T (needs only this body's adjF)
        warp_contacts_adj(M, L, lane, s, Rb, xc, cm0, st, clist, rec, adjF, adjS, G, adj_xc);
#endif
        // K2^T
        if (A.adj_res_f && L.valid) {
            float* r = A.adj_res_f + ((tp * A.bs + L.env) * M.nb + L.body) * 6;
            r[0] = nan0(adjF.t.x); r[1] = nan0(adjF.t.y); r[2] = nan0(adjF.t.z);
            r[3] = nan0(adjF.f.x); r[4] = nan0(adjF.f.y); r[5] = nan0(adjF.f.z);
        }
        // K4^T (this lane = child of its joint)
        BodyF P;
        F3 xcp;
        WrenchF adjFp;
        comm.get_parent_state_w(s, xc, adjF, L.parent_slot, P, xcp, adjFp);
        if (!L.has_parent) { P = body_identity<float>(); xcp = vzero<float>(); adjFp = wrench_zero<float>(); }
        BodyF adjP = body_zero<float>();
        F3 adj_xcp = vzero<float>();
        float g_target[3] = {0, 0, 0}, g_act[3] = {0, 0, 0};
        joint_adj<float, JM, LIMITS, QOFF>(st_joint<NT, QOFF>(st, xpq, L.body, L.type), ctl, M.ake, M.akd, P, xcp, L.has_parent, s,
                                           Rb, xc, adjFp, adjF, adjP, adj_xcp, adjS, G, adj_xc, g_target, g_act, a_ke,
                                           a_kd, ang);
        adjP.x += adj_xcp;
        adjP.r += qrot_adj_q(P.r, st_vec3(st, ST_CPAR, L.body), adj_xcp);
        if (!L.has_parent) adjP = body_zero<float>();
        comm.post_body(adjP);
        if (L.valid) {
            int64_t row = (tp * A.bs + L.env) * M.nqd + L.qds;
#pragma unroll
            for (int k = 0; k < 3; ++k) if (jon && k < L.ndof) {
                A.adj_refs[row + k] = nan0(g_target[k]);
                if (A.adj_torques) A.adj_torques[row + k] = nan0(g_act[k]);
            }
            if (!jon) for (int k = 0; k < L.ndof; ++k) {
                A.adj_refs[row + k] = 0.f;
                if (A.adj_torques) A.adj_torques[row + k] = 0.f;
            }
        }
#ifndef PPR_CONTACTS_EARLY
        // K3^T (needs only this body's adjF and feeds only this body's adjoint): after the parent adjoint is published,
        // so that the parent is not held up by the replay of this body's contact points
        if constexpr (Comm::kHelpers > 0) comm.await_contacts_adj(adjS, G, adj_xc);
        else warp_contacts_adj(M, L, lane, s, Rb, xc, cm0, st, clist, rec, adjF, adjS, G, adj_xc);
#endif
        // world COM -> pose: own-body part of the pose adjoint, evaluated while the children publish theirs
        adjS.x += adj_xc;
        m3_acc(G, adj_xc, com);
        adjS.r += qmat_adj(s.r, G);
        comm.gather_body(adjP, L.child, L.maxc_w, adjS);
        adjN = adjS;
    }
    // K1^T: state 0 = eval_fk(q_init, qd_init), recomputed
    {
        float jq[7], jqd[6];
        comm.sync();
        LaneInfo L2 = lane_setup(M, group, Comm::slot(), A.bs, Comm::envs_per_group(M), Comm::kBlock && PPR_BODY_MAJOR);  // re-derived (was in smem)
        load_joint_coords(L2, A.q_init, A.qd_init, M.nq, M.nqd, jq, jqd);
        BodyF s0 = warp_fk<JM>(comm, M, L2, jq, jqd);
        warp_fk_adjoint<JM>(comm, M, L2, s0, adjN, jq, jqd, A.adj_q_init, A.adj_qd_init);
    }
    if (A.adj_partial) {
        // ---- epilogue reduction of the shared-parameter gradients over the environments of this block:
        // every thread parks its 25 values in the (now free) joint_X_p / inertia-accumulator area, then (body, value)
        // pairs are summed over the block's threads of that body in thread order -> deterministic
        float* scratch = smem + SL::xpq;                 // 28 floats per thread are available, 25 used
        float v[25];
        v[0] = a_inv_m;
#pragma unroll
        for (int i = 0; i < 18; ++i) v[1 + i] = ((const float*)&acc[(i >> 2) * NT])[i & 3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { v[19 + k] = a_ke[k]; v[22 + k] = a_kd[k]; }
        __syncthreads();                                  // everyone is done with xpq / acc
#pragma unroll
        for (int c = 0; c < 25; ++c) scratch[c * NT + threadIdx.x] = nan0(v[c]);
        int* sbody = (int*)(smem + SL::sbody);
        sbody[threadIdx.x] = L.valid ? L.body : -1;
        __syncthreads();
        const int P = 2 * M.nqd + 19 * M.nb;
        float* out = A.adj_partial + (int64_t)blockIdx.x * P;
        // (warp layout: the warps of a partially filled last block that own no environment have left the kernel)
        // (team layout: the helper warps have left as well, the main warp reduces alone)
        const int nalive = Comm::kHelpers > 0 ? 32 : Comm::kBlock ? NT
                         : (int)min((int64_t)NT, (A.ngroups - (int64_t)blockIdx.x * (NT / 32)) * 32);
        for (int i = threadIdx.x; i < M.nb * 25; i += nalive) {
            const int b = i / 25, c = i - b * 25;
            float sum = 0.f;
            for (int t = 0; t < NT; ++t) if (sbody[t] == b) sum += scratch[c * NT + t];
            const int4 ji = M.jinfo[b];
            const int nd = M.jinfo2[b].x;
            if (c == 0) out[2 * M.nqd + b] = sum;
            else if (c < 10) out[2 * M.nqd + M.nb + b * 9 + (c - 1)] = sum;
            else if (c < 19) out[2 * M.nqd + M.nb * 10 + b * 9 + (c - 10)] = sum;
            else {
                const int k = (c - 19) % 3, base = (c < 22 ? 0 : M.nqd) + ji.w;
                if (ji.x != JT_FREE) { if (k < nd) out[base + k] = sum; }
                else { out[base + k] = 0.f; out[base + k + 3] = 0.f; }      // no PD on the six root dofs
            }
        }
    } else if (L.valid) {
        A.adj_inv_m[eb] = nan0(a_inv_m);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            A.adj_I[eb * 9 + i] = nan0(((const float*)&acc[(i >> 2) * NT])[i & 3]);
            A.adj_inv_I[eb * 9 + i] = nan0(((const float*)&acc[((9 + i) >> 2) * NT])[(9 + i) & 3]);
        }
        int64_t d = (int64_t)L.env * M.nqd + L.qds;
#pragma unroll
        for (int k = 0; k < 3; ++k) if (jon && k < L.ndof) { A.adj_ke[d + k] = nan0(a_ke[k]); A.adj_kd[d + k] = nan0(a_kd[k]); }
        if (!jon) for (int k = 0; k < L.ndof; ++k) { A.adj_ke[d + k] = 0.f; A.adj_kd[d + k] = 0.f; }
    }
}

// rows[R][P] -> out[P], fixed summation order (deterministic): 32 columns x 32 row lanes per block, four independent
// partial sums per thread (the loop is latency bound: R is ~10^4 rows of ~300 floats for a full batch)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ rows, int64_t R, int P,
                                                               float* __restrict__ out) {
    __shared__ float part[32][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (col < P) {
        int64_t r = threadIdx.y;
        for (; r + 96 < R; r += 128) {
            s0 += rows[r * P + col]; s1 += rows[(r + 32) * P + col];
            s2 += rows[(r + 64) * P + col]; s3 += rows[(r + 96) * P + col];
        }
        for (; r < R; r += 32) s0 += rows[r * P + col];
    }
    part[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.y == 0 && col < P) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += part[k][threadIdx.x];
        out[col] = t;
    }
}

// ----------------------------------------------------------------------------------------------- host side
static ppr_model* check(ppr_model_t m) { return (m && m->magic == PPR_MAGIC) ? m : nullptr; }

extern "C" const char* ppr_version(void) { return "ppr_b200 0.1.0 (sm_100a)"; }
extern "C" int64_t ppr_launch_count(void) { return g_launches.load(); }

extern "C" int ppr_model_create(const ppr_model_desc* D, ppr_model_t* out) {
    if (!D || !out) return PPR_E_ARG;
    if (D->nb < 1 || D->nb > 32 || D->nq < 1 || D->nqd < 0 || D->nc < 0 || D->nshape < 0) return PPR_E_SHAPE;
    if (!D->joint_type || !D->joint_parent || !D->joint_q_start || !D->joint_qd_start || !D->joint_X_p ||
        !D->joint_X_c || !D->joint_axis || !D->body_com)
        return PPR_E_ARG;
    if (D->nqd > 0 && (!D->joint_limit_lower || !D->joint_limit_upper || !D->joint_limit_ke || !D->joint_limit_kd))
        return PPR_E_ARG;
    if (D->nc > 0 && (!D->contact_body || !D->contact_point || !D->contact_dist || !D->contact_material ||
                      !D->shape_materials))
        return PPR_E_ARG;
    const int nb = D->nb;
    std::vector<int4> jinfo(nb), jinfo2(nb);
    std::vector<unsigned long long> child(nb, ~0ull);
    std::vector<int> nchild(nb, 0), depth(nb, 0);
    int maxc = 0, maxdepth = 0;
    for (int i = 0; i < nb; ++i) {
        int p = D->joint_parent[i];
        if (p >= i) return PPR_E_SHAPE;  // parents must precede children
        int ty = D->joint_type[i];
        int nd = (i + 1 < nb ? D->joint_qd_start[i + 1] : D->nqd) - D->joint_qd_start[i];
        if (ty != JT_REVOLUTE && ty != JT_FIXED && ty != JT_FREE && ty != JT_COMPOUND) return PPR_E_SHAPE;
        if ((ty == JT_FREE && nd != 6) || (ty == JT_REVOLUTE && nd != 1) || (ty == JT_COMPOUND && nd != 3) ||
            (ty == JT_FIXED && nd != 0))
            return PPR_E_SHAPE;
        if (p >= 0) {
            if (nchild[p] >= PPR_MAX_CHILD) return PPR_E_SHAPE;
            child[p] &= ~(0xffull << (8 * nchild[p]));
            child[p] |= (unsigned long long)i << (8 * nchild[p]);
            nchild[p]++;
            if (nchild[p] > maxc) maxc = nchild[p];
            depth[i] = depth[p] + 1;
            if (depth[i] > maxdepth) maxdepth = depth[i];
        }
        jinfo[i] = make_int4(ty, p, D->joint_q_start[i], D->joint_qd_start[i]);
        jinfo2[i] = make_int4(nd, depth[i], 0, 0);
    }
    // contacts sorted by body (stable)
    std::vector<float4> cpt;
    std::vector<int> cmat;
    std::vector<float> aabb(nb * 8, 0.f);
    for (int b = 0; b < nb; ++b) {
        int c0 = (int)cpt.size();
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f}, dmax = 0.f;
        for (int k = 0; k < D->nc; ++k) {
            if (D->contact_body[k] != b) continue;
            if (D->contact_material[k] < 0 || D->contact_material[k] >= D->nshape) return PPR_E_ARG;
            const float* p = D->contact_point + 3 * k;
            cpt.push_back(make_float4(p[0], p[1], p[2], D->contact_dist[k]));
            cmat.push_back(D->contact_material[k]);
            for (int i = 0; i < 3; ++i) { lo[i] = p[i] < lo[i] ? p[i] : lo[i]; hi[i] = p[i] > hi[i] ? p[i] : hi[i]; }
            dmax = D->contact_dist[k] > dmax ? D->contact_dist[k] : dmax;
        }
        int c1 = (int)cpt.size();
        jinfo2[b].z = c0; jinfo2[b].w = c1;
        if (c1 > c0) { for (int i = 0; i < 3; ++i) { aabb[8 * b + i] = lo[i]; aabb[8 * b + 3 + i] = hi[i]; } aabb[8 * b + 6] = dmax; }
    }
    for (int k = 0; k < D->nc; ++k) if (D->contact_body[k] < 0 || D->contact_body[k] >= nb) return PPR_E_ARG;
    // support-function tables of the big bodies (support_lower_bound): m(n) = min_k n . p_k at the nodes of a cube map,
    // n = (+-1, u, v) / (u, +-1, v) / (u, v, +-1) with u, v = -1 + 2 i / N; evaluated in double, rounded DOWN to float
    std::vector<float> sup;
    std::vector<int> sup_of(nb, -1);
    const int big_threshold = 16;
    for (int b = 0; b < nb; ++b) {
        const int c0 = jinfo2[b].z, c1 = jinfo2[b].w;
        if (c1 - c0 <= big_threshold) continue;
        sup_of[b] = (int)(sup.size() / PPR_SUP_FLOATS);
        for (int face = 0; face < 6; ++face)
            for (int iv = 0; iv <= PPR_SUP_N; ++iv)
                for (int iu = 0; iu <= PPR_SUP_N; ++iu) {
                    const double u = -1.0 + 2.0 * iu / PPR_SUP_N, v = -1.0 + 2.0 * iv / PPR_SUP_N, sg = (face & 1) ? -1.0 : 1.0;
                    double n[3];
                    if (face < 2) { n[0] = sg; n[1] = u; n[2] = v; }
                    else if (face < 4) { n[0] = u; n[1] = sg; n[2] = v; }
                    else { n[0] = u; n[1] = v; n[2] = sg; }
                    double best = 1e300;
                    for (int k = c0; k < c1; ++k) {
                        const double dd = n[0] * cpt[k].x + n[1] * cpt[k].y + n[2] * cpt[k].z;
                        best = dd < best ? dd : best;
                    }
                    float f = (float)best;
                    if ((double)f > best) f = nextafterf(f, -INFINITY);
                    sup.push_back(f);
                }
    }
    std::vector<float4> lim(D->nqd > 0 ? D->nqd : 1);
    for (int k = 0; k < D->nqd; ++k)
        lim[k] = make_float4(D->joint_limit_lower[k], D->joint_limit_upper[k], D->joint_limit_ke[k], D->joint_limit_kd[k]);
    std::vector<float> qoff(nb * 4);
    for (int i = 0; i < nb; ++i) for (int k = 0; k < 4; ++k) qoff[4 * i + k] = D->joint_X_c[7 * i + 3 + k];

    // pack into one blob
    struct Seg { const void* src; size_t bytes; size_t off; };
    std::vector<Seg> segs;
    size_t total = 0;
    auto add = [&](const void* p, size_t bytes) {
        size_t off = (total + 255) & ~(size_t)255;
        segs.push_back({p, bytes, off});
        total = off + (bytes ? bytes : 16);
        return off;
    };
    size_t o_jinfo = add(jinfo.data(), nb * sizeof(int4)), o_jinfo2 = add(jinfo2.data(), nb * sizeof(int4));
    size_t o_child = add(child.data(), nb * sizeof(unsigned long long));
    std::vector<int> ident(nb);
    for (int i = 0; i < nb; ++i) ident[i] = i;
    size_t o_order = add(ident.data(), nb * sizeof(int)), o_pos = add(ident.data(), nb * sizeof(int));
    size_t o_xpj = add(D->joint_X_p, nb * 7 * sizeof(float)), o_qoff = add(qoff.data(), nb * 4 * sizeof(float));
    size_t o_axis = add(D->joint_axis, nb * 3 * sizeof(float)), o_com = add(D->body_com, nb * 3 * sizeof(float));
    size_t o_lim = add(lim.data(), lim.size() * sizeof(float4));
    size_t o_cpt = add(cpt.data(), cpt.size() * sizeof(float4)), o_cmat = add(cmat.data(), cmat.size() * sizeof(int));
    size_t o_mats = add(D->shape_materials, (size_t)D->nshape * 4 * sizeof(float));
    size_t o_aabb = add(aabb.data(), aabb.size() * sizeof(float));
    size_t o_sup = add(sup.data(), sup.size() * sizeof(float)), o_supof = add(sup_of.data(), nb * sizeof(int));

    ppr_model* m = new (std::nothrow) ppr_model();
    if (!m) return PPR_E_ARG;
    cudaError_t e = cudaGetDevice(&m->device);
    if (e != cudaSuccess) { delete m; return (int)e; }
    e = cudaMalloc(&m->blob, total);
    if (e != cudaSuccess) { delete m; return (int)e; }
    std::vector<char> host(total, 0);
    for (auto& s : segs) if (s.bytes) memcpy(host.data() + s.off, s.src, s.bytes);
    e = cudaMemcpy(m->blob, host.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(m->blob); delete m; return (int)e; }
    char* base = (char*)m->blob;
    DevModel& d = m->d;
    d.nb = nb; d.nq = D->nq; d.nqd = D->nqd; d.nc = (int)cpt.size();
    d.epw = 32 / nb; d.maxc = maxc; d.maxdepth = maxdepth; d.big_threshold = big_threshold;
    d.jinfo = (const int4*)(base + o_jinfo); d.jinfo2 = (const int4*)(base + o_jinfo2);
    d.child = (const unsigned long long*)(base + o_child);
    d.order = (const int*)(base + o_order); d.pos = (const int*)(base + o_pos);
    d.xpj_env = nullptr; d.xpj_nenv = 0;
    d.xpj = (const float*)(base + o_xpj); d.qoff = (const float*)(base + o_qoff);
    d.axis = (const float*)(base + o_axis); d.com = (const float*)(base + o_com);
    d.lim = (const float4*)(base + o_lim); d.cpt = (const float4*)(base + o_cpt); d.cmat = (const int*)(base + o_cmat);
    d.mats = (const float4*)(base + o_mats); d.aabb = (const float*)(base + o_aabb);
    d.sup = (const float*)(base + o_sup); d.sup_of = (const int*)(base + o_supof);
    d.g[0] = D->gravity[0]; d.g[1] = D->gravity[1]; d.g[2] = D->gravity[2];
    d.ake = D->joint_attach_ke; d.akd = D->joint_attach_kd;
    d.mat_uniform = 1;
    d.ground = 1;
    for (size_t k = 1; k < cmat.size(); ++k) if (cmat[k] != cmat[0]) d.mat_uniform = 0;
    bool any_rev = false, any_cmp = false, any_other = false, limits = false, qoffs = false;
    for (int i = 0; i < nb; ++i) {
        int ty = D->joint_type[i];
        if (ty == JT_REVOLUTE) any_rev = true;
        else if (ty == JT_COMPOUND) {
            any_cmp = true;
            const float* q = D->joint_X_c + 7 * i + 3;
            if (q[0] != 0.f || q[1] != 0.f || q[2] != 0.f || q[3] != 1.f) qoffs = true;
        } else if (ty != JT_FREE) any_other = true;
    }
    for (int k = 0; k < D->nqd; ++k) if (D->joint_limit_ke[k] != 0.f || D->joint_limit_kd[k] != 0.f) limits = true;
    m->variant = 2;
    if (!any_other && !limits && !qoffs) {
        if (!any_cmp) m->variant = 0;
        else if (!any_rev) m->variant = 1;
    }
    // env packing: lanes doing useful work per warp-packed warp vs per block-packed block
    {
        double best = (double)((32 / nb) * nb) / 32.0;
        m->comm = 0;
        for (int c = 1; c < 3; ++c) {
            int nt = kCommThreads[c];
            double u = (double)((nt / nb) * nb) / nt;
            if (u > best * 1.12) { best = u; m->comm = c; }  // the barriers must buy at least ~12 % more useful lanes
        }
        const char* ov = getenv("PPR_COMM");
        if (ov && ov[0] >= '0' && ov[0] <= '2') m->comm = ov[0] - '0';
    }
    // Block layout: order of the bodies inside the block (slot = position * envs_per_block + env).  The contact pass is
    // warp-cooperative and serial over a warp's "big" bodies (collision meshes), and the block's warps meet at two
    // barriers per substep, so the big bodies that usually touch the ground -- the leaves of the tree: feet, hands,
    // head (inner links count 1/32) -- are spread evenly over the warps: minimise the largest per-warp sum of their
    // vertices (then the sum of squares) by pairwise swaps from the identity order; deterministic.
    if (m->comm != 0) {
        const int nt = kCommThreads[m->comm], epb = nt / nb, nw = nt / 32;
        std::vector<int> w(nb), order(ident), pos(nb);
        bool any_big = false;
        for (int b = 0; b < nb; ++b) {
            int n = jinfo2[b].w - jinfo2[b].z;
            w[b] = n > d.big_threshold ? (nchild[b] == 0 ? 32 * n : n) : 0;
            any_big |= w[b] > 0;
        }
        auto cost = [&](const std::vector<int>& o) {
            std::vector<long long> load(nw, 0);
            for (int p2 = 0; p2 < nb; ++p2)
                for (int e2 = 0; e2 < epb; ++e2) load[(p2 * epb + e2) / 32] += w[o[p2]];
            long long mx = 0, sq = 0;
            for (long long l : load) { mx = l > mx ? l : mx; sq += l * l; }
            return std::make_pair(mx, sq);
        };
        const char* ord = getenv("PPR_ORDER");
        if (any_big && ord && !strcmp(ord, "cluster")) {
            // experiment: all contact-prone bodies next to each other (one warp does the contact work of the block)
            std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return w[a] > w[b2]; });
        } else if (any_big && !getenv("PPR_NO_BALANCE")) {
            auto best = cost(order);
            for (bool improved = true; improved;) {
                improved = false;
                for (int i = 0; i < nb; ++i)
                    for (int j = i + 1; j < nb; ++j) {
                        std::swap(order[i], order[j]);
                        auto c = cost(order);
                        if (c < best) { best = c; improved = true; }
                        else std::swap(order[i], order[j]);
                    }
            }
        }
        for (int p2 = 0; p2 < nb; ++p2) pos[order[p2]] = p2;
        e = cudaMemcpy(base + o_order, order.data(), nb * sizeof(int), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(base + o_pos, pos.data(), nb * sizeof(int), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(m->blob); delete m; return (int)e; }
    }
    m->ckpt_every = 1;
    m->latency_envs = 1024;   // ~148 SMs x 4 schedulers x 2 warps: measured break-even of the two layouts on B200
    if (const char* e = getenv("PPR_LATENCY_ENVS")) m->latency_envs = atoll(e);
    m->team_envs = 296;       // 2 blocks of 3 warps per SM: every warp still (nearly) has a scheduler to itself
    if (const char* e = getenv("PPR_TEAM_ENVS")) m->team_envs = atoll(e);
    m->xpj_offset = o_xpj;
    m->h_xpj.assign(D->joint_X_p, D->joint_X_p + nb * 7);
    m->magic = PPR_MAGIC;
    *out = m;
    return 0;
}

extern "C" int ppr_model_destroy(ppr_model_t m) {
    if (!check(m)) return PPR_E_HANDLE;
    m->magic = 0;
    cudaFree(m->blob);
    delete m;
    return 0;
}

extern "C" int ppr_model_set_joint_X_p(ppr_model_t m, const float* xp, void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    if (!xp) return PPR_E_ARG;
    m->h_xpj.assign(xp, xp + m->d.nb * 7);
    cudaError_t e = cudaMemcpyAsync((char*)m->blob + m->xpj_offset, m->h_xpj.data(), m->h_xpj.size() * sizeof(float),
                                    cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return (int)e;
}
extern "C" int ppr_model_set_attach(ppr_model_t m, float ke, float kd) {
    if (!check(m)) return PPR_E_HANDLE;
    m->d.ake = ke; m->d.akd = kd;
    return 0;
}
extern "C" int ppr_model_set_ground(ppr_model_t m, int32_t ground) {
    if (!check(m)) return PPR_E_HANDLE;
    m->d.ground = ground ? 1 : 0;
    return 0;
}
extern "C" int ppr_model_set_gravity(ppr_model_t m, const float g[3]) {
    if (!check(m)) return PPR_E_HANDLE;
    if (!g) return PPR_E_ARG;
    m->d.g[0] = g[0]; m->d.g[1] = g[1]; m->d.g[2] = g[2];
    return 0;
}
extern "C" int ppr_model_set_joint_X_p_env(ppr_model_t m, const float* dev_joint_X_p, int64_t n_env) {
    if (!check(m)) return PPR_E_HANDLE;
    if ((dev_joint_X_p != nullptr) != (n_env > 0) || n_env < 0) return PPR_E_ARG;
    m->d.xpj_env = dev_joint_X_p;
    m->d.xpj_nenv = n_env;
    return 0;
}
extern "C" int ppr_model_set_checkpoint_every(ppr_model_t m, int32_t every) {
    if (!check(m)) return PPR_E_HANDLE;
    if (every < 1 || every > 4096) return PPR_E_ARG;
    m->ckpt_every = every;
    return 0;
}
extern "C" int ppr_model_set_latency_envs(ppr_model_t m, int64_t max_envs) {
    if (!check(m)) return PPR_E_HANDLE;
    if (max_envs < 0) return PPR_E_ARG;
    m->latency_envs = max_envs;
    return 0;
}
extern "C" int64_t ppr_model_latency_envs(ppr_model_t m) { return check(m) ? m->latency_envs : PPR_E_HANDLE; }
extern "C" int ppr_model_set_team_envs(ppr_model_t m, int64_t max_envs) {
    if (!check(m)) return PPR_E_HANDLE;
    if (max_envs < 0) return PPR_E_ARG;
    m->team_envs = max_envs;
    return 0;
}
extern "C" int64_t ppr_model_team_envs(ppr_model_t m) { return check(m) ? m->team_envs : PPR_E_HANDLE; }
extern "C" int ppr_model_envs_per_group(ppr_model_t m) {
    if (!check(m)) return PPR_E_HANDLE;
    return m->comm == 0 ? m->d.epw : kCommThreads[m->comm] / m->d.nb;
}
extern "C" int ppr_model_group_threads(ppr_model_t m) {
    if (!check(m)) return PPR_E_HANDLE;
    return m->comm == 0 ? 32 : kCommThreads[m->comm];
}

static inline int64_t nwarps_for(const DevModel& d, int64_t n) { return (n + d.epw - 1) / d.epw; }
// floats per body in a checkpoint row: the kernel instances of variant 0 (JM_REVOLUTE) keep 6 quads, the others 7
static inline size_t row_floats(const ppr_model* m) { return (m->variant == 0 ? RowOf<JM_REVOLUTE>::kQuads : RowOf<JM_ALL>::kQuads) * 4; }
// rollout kernels: groups (warps or blocks) and warps (each owns one checkpoint row per substep)
// Small batches cannot fill 148 SMs x 4 schedulers, so what counts is the LATENCY of one substep, and that grows with
// the number of environments a warp serialises in the contact phase: below `latency_envs` environments every
// environment gets a warp of its own (warp layout, 1 env per warp) whatever the throughput layout of the model is.
// A pure function of (model, bs): the three entry points of one rollout derive the same geometry.
static inline void rollout_geometry(const ppr_model* m, int64_t bs, int64_t& ngroups, int64_t& nwarps, unsigned& grid,
                                    int& comm, int& epw) {
    comm = m->comm;
    epw = m->d.epw;
    if (bs <= m->latency_envs) { comm = 0; epw = 1; }
    // still smaller batches: the contacts of an environment move to helper warps (team layout; stored rows only)
    if (comm == 0 && epw == 1 && bs <= m->team_envs && m->ckpt_every == 1) {
        comm = PPR_COMM_TEAM;
        ngroups = bs; nwarps = bs; grid = (unsigned)bs;
        return;
    }
    const int nt = kCommThreads[comm];
    if (comm == 0) {
        ngroups = (bs + epw - 1) / epw;
        grid = (unsigned)((ngroups * 32 + nt - 1) / nt);
        nwarps = ngroups;
    } else {
        int epb = nt / m->d.nb;
        ngroups = (bs + epb - 1) / epb;
        grid = (unsigned)ngroups;
        nwarps = ngroups * (nt / 32);
    }
}
template <class K> static cudaError_t launch_rollout(K kernel, size_t smem, unsigned grid, int nt, cudaStream_t st,
                                                     const DevModel& d, const RolloutArgs& A) {
    // opt in to > 48 kB of dynamic shared memory once per (device, kernel instantiation): the attribute is per device.
    // All instantiations share this function template (same signature), hence the small pointer table; a slot is
    // claimed with a compare-exchange, a lost race only repeats the (idempotent) attribute call.
    static std::atomic<const void*> done[PPR_MAX_DEVICES][64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    bool found = false;
    if (dev >= 0 && dev < PPR_MAX_DEVICES) {
        for (int i = 0; i < 64 && !found; ++i) {
            const void* p = done[dev][i].load(std::memory_order_acquire);
            if (p == (const void*)kernel) found = true;
            else if (p == nullptr) {
                e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                const void* expected = nullptr;
                done[dev][i].compare_exchange_strong(expected, (const void*)kernel, std::memory_order_acq_rel);
                found = true;
            }
        }
    }
    if (!found) {   // table full or unusual device ordinal: set it every time
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, nt, smem, st>>>(d, A);
    return cudaGetLastError();
}
#ifdef PPR_AB_ONLY
// quick A/B builds (tools/ab.sh): only the block-packed instances of the two shipped feature sets
#define PPR_LAUNCH_ROLLOUT(KERNEL, ADJ, ...)                                                                              \
    do {                                                                                                             \
        cudaError_t e_ = cudaErrorInvalidValue;                                                                      \
        DevModel d_ = m->d;                                                                                          \
        d_.epw = epw_;                                                                                               \
        typedef BlockComm<PPR_NT1, ADJ> C_;                                                                          \
        if (comm_ == 1 && m->variant == 0) e_ = launch_rollout(KERNEL<C_, JM_REVOLUTE, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, PPR_NT1, st, d_, A); \
        else if (comm_ == 1 && m->variant == 1) e_ = launch_rollout(KERNEL<C_, JM_COMPOUND, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, PPR_NT1, st, d_, A); \
        g_launches++;                                                                                                \
        return (int)e_;                                                                                              \
    } while (0)
#else
#define PPR_LAUNCH_ROLLOUT(KERNEL, ADJ, ...)                                                                              \
    do {                                                                                                             \
        cudaError_t e_;                                                                                              \
        DevModel d_ = m->d;                                                                                          \
        d_.epw = epw_;                                                                                               \
        if (comm_ == 0) {                                                                                            \
            typedef WarpComm<128, ADJ> C_;                                                                           \
            if (m->variant == 0) e_ = launch_rollout(KERNEL<C_, JM_REVOLUTE, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, 128, st, d_, A); \
            else if (m->variant == 1) e_ = launch_rollout(KERNEL<C_, JM_COMPOUND, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, 128, st, d_, A); \
            else e_ = launch_rollout(KERNEL<C_, JM_ALL, true, true __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, 128, st, d_, A); \
        } else if (comm_ == PPR_COMM_TEAM) {                                                                         \
            typedef TeamComm<PPR_TEAM_H, ADJ> C_;                                                                    \
            if (m->variant == 0) e_ = launch_rollout(KERNEL<C_, JM_REVOLUTE, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, C_::kThreads, st, d_, A); \
            else if (m->variant == 1) e_ = launch_rollout(KERNEL<C_, JM_COMPOUND, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, C_::kThreads, st, d_, A); \
            else e_ = launch_rollout(KERNEL<C_, JM_ALL, true, true __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, C_::kThreads, st, d_, A); \
        } else if (comm_ == 1) {                                                                                     \
            typedef BlockComm<PPR_NT1, ADJ> C_;                                                                        \
            if (m->variant == 0) e_ = launch_rollout(KERNEL<C_, JM_REVOLUTE, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, PPR_NT1, st, d_, A); \
            else if (m->variant == 1) e_ = launch_rollout(KERNEL<C_, JM_COMPOUND, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, PPR_NT1, st, d_, A); \
            else e_ = launch_rollout(KERNEL<C_, JM_ALL, true, true __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, PPR_NT1, st, d_, A); \
        } else {                                                                                                     \
            typedef BlockComm<160, ADJ> C_;                                                                          \
            if (m->variant == 0) e_ = launch_rollout(KERNEL<C_, JM_REVOLUTE, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, 160, st, d_, A); \
            else if (m->variant == 1) e_ = launch_rollout(KERNEL<C_, JM_COMPOUND, false, false __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, 160, st, d_, A); \
            else e_ = launch_rollout(KERNEL<C_, JM_ALL, true, true __VA_ARGS__>, SmemLayout<C_, ADJ>::bytes, grid, 160, st, d_, A); \
        }                                                                                                            \
        g_launches++;                                                                                                \
        return (int)e_;                                                                                              \
    } while (0)
#endif
static inline unsigned grid_for(int64_t nwarps) { return (unsigned)((nwarps * 32 + PPR_BLOCK - 1) / PPR_BLOCK); }

extern "C" int ppr_fk_forward(ppr_model_t m, int64_t n, const float* q, const float* qd, float* bq, float* bqd,
                              void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    DeviceGuard guard_(m->device);
    if (guard_.err != cudaSuccess) return (int)guard_.err;
    if (n < 0 || !q || !qd || !bq || !bqd) return PPR_E_ARG;
    if (n == 0) return 0;
    dim3 grid(grid_for(nwarps_for(m->d, n)));
    cudaStream_t st = (cudaStream_t)stream;
    if (m->variant == 0) fk_forward_kernel<JM_REVOLUTE><<<grid, PPR_BLOCK, 0, st>>>(m->d, n, q, qd, bq, bqd);
    else if (m->variant == 1) fk_forward_kernel<JM_COMPOUND><<<grid, PPR_BLOCK, 0, st>>>(m->d, n, q, qd, bq, bqd);
    else fk_forward_kernel<JM_ALL><<<grid, PPR_BLOCK, 0, st>>>(m->d, n, q, qd, bq, bqd);
    g_launches++;
    return (int)cudaGetLastError();
}

extern "C" int ppr_fk_backward(ppr_model_t m, int64_t n, const float* q, const float* qd, const float* abq,
                               const float* abqd, float* aq, float* aqd, void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    DeviceGuard guard_(m->device);
    if (guard_.err != cudaSuccess) return (int)guard_.err;
    if (n < 0 || !q || !qd || !abq || !abqd || !aq || !aqd) return PPR_E_ARG;
    if (n == 0) return 0;
    dim3 grid(grid_for(nwarps_for(m->d, n)));
    cudaStream_t st = (cudaStream_t)stream;
    if (m->variant == 0) fk_backward_kernel<JM_REVOLUTE><<<grid, PPR_BLOCK, 0, st>>>(m->d, n, q, qd, abq, abqd, aq, aqd);
    else if (m->variant == 1) fk_backward_kernel<JM_COMPOUND><<<grid, PPR_BLOCK, 0, st>>>(m->d, n, q, qd, abq, abqd, aq, aqd);
    else fk_backward_kernel<JM_ALL><<<grid, PPR_BLOCK, 0, st>>>(m->d, n, q, qd, abq, abqd, aq, aqd);
    g_launches++;
    return (int)cudaGetLastError();
}

extern "C" size_t ppr_rollout_workspace_bytes(ppr_model_t m, int64_t bs, int64_t nsteps) {
    if (!check(m) || bs <= 0 || nsteps <= 0) return 0;
    int64_t ngroups, nwarps; unsigned grid; int comm_, epw_;
    rollout_geometry(m, bs, ngroups, nwarps, grid, comm_, epw_);
    const int64_t K = m->ckpt_every;
    const int64_t rows = (nsteps + K - 1) / K + (K > 1 ? K : 0);   // stored rows + per-warp scratch rows
    return (size_t)nwarps * (size_t)rows * row_floats(m) * 32 * sizeof(float);
}

extern "C" int ppr_rollout_forward(ppr_model_t m, int64_t bs, int64_t nsteps, int64_t stride, float dt,
                                   int32_t shared_params, const float* q_init, const float* qd_init, const float* torques, const float* res_f,
                                   const float* refs, const float* ke, const float* kd, const float* inv_m,
                                   const float* I, const float* inv_I, float* out_pos, float* out_vel, float* out_grf,
                                   float* out_jaf, void* ws, size_t ws_bytes, void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    DeviceGuard guard_(m->device);
    if (guard_.err != cudaSuccess) return (int)guard_.err;
    if (bs < 0 || nsteps < 1 || stride < 1) return PPR_E_ARG;
    if (!q_init || !qd_init || !refs || !ke || !kd || !inv_m || !I || !inv_I || !out_pos || !out_vel || !ws)
        return PPR_E_ARG;
    if (bs == 0) return 0;
    if (ws_bytes < ppr_rollout_workspace_bytes(m, bs, nsteps)) return PPR_E_WORKSPACE;
    RolloutArgs A;
    memset(&A, 0, sizeof(A));
    unsigned grid; int comm_, epw_;
    rollout_geometry(m, bs, A.ngroups, A.nwarps, grid, comm_, epw_);
    A.bs = bs; A.nsteps = nsteps; A.stride = stride; A.dt = dt; A.ckpt_every = m->ckpt_every;
    A.pstride = shared_params ? 0 : 1;
    A.q_init = q_init; A.qd_init = qd_init; A.torques = torques; A.res_f = res_f; A.refs = refs; A.ke = ke; A.kd = kd;
    A.inv_m = inv_m; A.I = I; A.inv_I = inv_I;
    A.out_pos = out_pos; A.out_vel = out_vel; A.out_grf = out_grf; A.out_jaf = out_jaf; A.ckpt = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    PPR_LAUNCH_ROLLOUT(rollout_forward_kernel, false, );
}

extern "C" int ppr_rollout_backward(ppr_model_t m, int64_t bs, int64_t nsteps, int64_t stride, float dt,
                                    int32_t shared_params, const float* q_init, const float* qd_init, const float* torques,
                                    const float* res_f, const float* refs, const float* ke, const float* kd,
                                    const float* inv_m, const float* I, const float* inv_I, const float* adj_pos,
                                    const float* adj_vel, float* adj_q_init, float* adj_qd_init, float* adj_torques,
                                    float* adj_res_f, float* adj_refs, float* adj_ke, float* adj_kd, float* adj_inv_m,
                                    float* adj_I, float* adj_inv_I, const void* ws, size_t ws_bytes, void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    DeviceGuard guard_(m->device);
    if (guard_.err != cudaSuccess) return (int)guard_.err;
    if (bs < 0 || nsteps < 1 || stride < 1) return PPR_E_ARG;
    if (!q_init || !qd_init || !refs || !ke || !kd || !inv_m || !I || !inv_I || !adj_pos || !adj_vel || !adj_q_init ||
        !adj_qd_init || !adj_refs || !adj_ke || !adj_kd || !adj_inv_m || !adj_I || !adj_inv_I || !ws)
        return PPR_E_ARG;
    if (bs == 0) return 0;
    if (ws_bytes < ppr_rollout_workspace_bytes(m, bs, nsteps)) return PPR_E_WORKSPACE;
    RolloutArgs A;
    memset(&A, 0, sizeof(A));
    unsigned grid; int comm_, epw_;
    rollout_geometry(m, bs, A.ngroups, A.nwarps, grid, comm_, epw_);
    A.bs = bs; A.nsteps = nsteps; A.stride = stride; A.dt = dt; A.ckpt_every = m->ckpt_every;
    A.pstride = shared_params ? 0 : 1;
    A.q_init = q_init; A.qd_init = qd_init; A.torques = torques; A.res_f = res_f; A.refs = refs; A.ke = ke; A.kd = kd;
    A.inv_m = inv_m; A.I = I; A.inv_I = inv_I; A.ckpt = (float*)ws;
    A.adj_pos = adj_pos; A.adj_vel = adj_vel; A.adj_q_init = adj_q_init; A.adj_qd_init = adj_qd_init;
    A.adj_torques = adj_torques; A.adj_res_f = adj_res_f; A.adj_refs = adj_refs; A.adj_ke = adj_ke; A.adj_kd = adj_kd;
    A.adj_inv_m = adj_inv_m; A.adj_I = adj_I; A.adj_inv_I = adj_inv_I;
    cudaStream_t st = (cudaStream_t)stream;
#ifndef PPR_AB_ONLY
    if (m->ckpt_every > 1) PPR_LAUNCH_ROLLOUT(rollout_backward_kernel, true, , true);
#endif
    PPR_LAUNCH_ROLLOUT(rollout_backward_kernel, true, , false);
}

extern "C" int64_t ppr_rollout_shared_grad_floats(ppr_model_t m) {
    if (!check(m)) return PPR_E_HANDLE;
    return 2 * (int64_t)m->d.nqd + 19 * (int64_t)m->d.nb;
}
extern "C" size_t ppr_rollout_reduce_scratch_bytes(ppr_model_t m, int64_t bs) {
    if (!check(m) || bs <= 0) return 0;
    int64_t ngroups, nwarps; unsigned grid; int comm_, epw_;
    rollout_geometry(m, bs, ngroups, nwarps, grid, comm_, epw_);
    return (size_t)grid * (size_t)(2 * m->d.nqd + 19 * m->d.nb) * sizeof(float);
}
extern "C" int ppr_rollout_backward_shared(ppr_model_t m, int64_t bs, int64_t nsteps, int64_t stride, float dt,
                                           const float* q_init, const float* qd_init, const float* torques,
                                           const float* res_f, const float* refs, const float* ke, const float* kd,
                                           const float* inv_m, const float* I, const float* inv_I, const float* adj_pos,
                                           const float* adj_vel, float* adj_q_init, float* adj_qd_init, float* adj_torques,
                                           float* adj_res_f, float* adj_refs, float* adj_shared, void* scratch,
                                           size_t scratch_bytes, const void* ws, size_t ws_bytes, void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    DeviceGuard guard_(m->device);
    if (guard_.err != cudaSuccess) return (int)guard_.err;
    if (bs < 0 || nsteps < 1 || stride < 1) return PPR_E_ARG;
    if (!q_init || !qd_init || !refs || !ke || !kd || !inv_m || !I || !inv_I || !adj_pos || !adj_vel || !adj_q_init ||
        !adj_qd_init || !adj_refs || !adj_shared || !scratch || !ws)
        return PPR_E_ARG;
    if (bs == 0) return 0;
    if (ws_bytes < ppr_rollout_workspace_bytes(m, bs, nsteps)) return PPR_E_WORKSPACE;
    if (scratch_bytes < ppr_rollout_reduce_scratch_bytes(m, bs)) return PPR_E_WORKSPACE;
    RolloutArgs A;
    memset(&A, 0, sizeof(A));
    unsigned grid; int comm_, epw_;
    rollout_geometry(m, bs, A.ngroups, A.nwarps, grid, comm_, epw_);
    A.bs = bs; A.nsteps = nsteps; A.stride = stride; A.dt = dt; A.ckpt_every = m->ckpt_every;
    A.pstride = 0;
    A.q_init = q_init; A.qd_init = qd_init; A.torques = torques; A.res_f = res_f; A.refs = refs; A.ke = ke; A.kd = kd;
    A.inv_m = inv_m; A.I = I; A.inv_I = inv_I; A.ckpt = (float*)ws;
    A.adj_pos = adj_pos; A.adj_vel = adj_vel; A.adj_q_init = adj_q_init; A.adj_qd_init = adj_qd_init;
    A.adj_torques = adj_torques; A.adj_res_f = adj_res_f; A.adj_refs = adj_refs;
    A.adj_partial = (float*)scratch;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = [&]() -> int {
#ifndef PPR_AB_ONLY
        if (m->ckpt_every > 1) PPR_LAUNCH_ROLLOUT(rollout_backward_kernel, true, , true);
#endif
        PPR_LAUNCH_ROLLOUT(rollout_backward_kernel, true, , false);
    }();
    if (rc != 0) return rc;
    const int P = 2 * m->d.nqd + 19 * m->d.nb;
    reduce_partials_kernel<<<(unsigned)((P + 31) / 32), dim3(32, 32), 0, st>>>((const float*)scratch, (int64_t)grid, P, adj_shared);
    g_launches++;
    return (int)cudaGetLastError();
}

// ---- struct-argument entry points: every option of the rollout in one place (shared parameters, device-side reduction
// of their gradients, fused pose loss); the positional entry points above are the reference-shaped subset
static int fill_args(ppr_model* m, const ppr_rollout_io* io, RolloutArgs& A, unsigned& grid, int& comm_, int& epw_) {
    if (io->bs < 0 || io->nsteps < 1 || io->frame_stride < 1) return PPR_E_ARG;
    if (!io->q_init || !io->qd_init || !io->refs || !io->target_ke || !io->target_kd || !io->body_inv_mass ||
        !io->body_inertia || !io->body_inv_inertia || !io->out_pos || !io->out_vel || !io->workspace)
        return PPR_E_ARG;
    if ((io->loss_pos || io->adj_loss_pos) && !io->target_pos) return PPR_E_ARG;
    memset(&A, 0, sizeof(A));
    if (io->bs == 0) return 0;
    if (io->workspace_bytes < ppr_rollout_workspace_bytes(m, io->bs, io->nsteps)) return PPR_E_WORKSPACE;
    rollout_geometry(m, io->bs, A.ngroups, A.nwarps, grid, comm_, epw_);
    A.bs = io->bs; A.nsteps = io->nsteps; A.stride = io->frame_stride; A.dt = io->dt; A.ckpt_every = m->ckpt_every;
    A.pstride = io->shared_params ? 0 : 1;
    A.q_init = io->q_init; A.qd_init = io->qd_init; A.torques = io->torques; A.res_f = io->res_f; A.refs = io->refs;
    A.ke = io->target_ke; A.kd = io->target_kd; A.inv_m = io->body_inv_mass; A.I = io->body_inertia;
    A.inv_I = io->body_inv_inertia; A.ckpt = (float*)io->workspace;
    A.target_pos = io->target_pos; A.rot_ratio = io->rot_ratio;
    return 0;
}
extern "C" int ppr_rollout_forward_ex(ppr_model_t m, const ppr_rollout_io* io, void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    if (!io) return PPR_E_ARG;
    DeviceGuard guard_(m->device);
    if (guard_.err != cudaSuccess) return (int)guard_.err;
    RolloutArgs A;
    unsigned grid = 0; int comm_ = 0, epw_ = 1;
    int rc = fill_args(m, io, A, grid, comm_, epw_);
    if (rc != 0 || io->bs == 0) return rc;
    A.out_pos = io->out_pos; A.out_vel = io->out_vel; A.out_grf = io->out_grf; A.out_jaf = io->out_jaf;
    A.loss_pos = io->loss_pos;
    cudaStream_t st = (cudaStream_t)stream;
    PPR_LAUNCH_ROLLOUT(rollout_forward_kernel, false, );
}
extern "C" int ppr_rollout_backward_ex(ppr_model_t m, const ppr_rollout_io* io, void* stream) {
    if (!check(m)) return PPR_E_HANDLE;
    if (!io) return PPR_E_ARG;
    DeviceGuard guard_(m->device);
    if (guard_.err != cudaSuccess) return (int)guard_.err;
    RolloutArgs A;
    unsigned grid = 0; int comm_ = 0, epw_ = 1;
    int rc = fill_args(m, io, A, grid, comm_, epw_);
    if (rc != 0 || io->bs == 0) return rc;
    if (!io->adj_q_init || !io->adj_qd_init || !io->adj_refs) return PPR_E_ARG;
    if (!io->adj_out_pos && !io->adj_out_vel && !io->adj_loss_pos) return PPR_E_ARG;      // nothing to differentiate
    const bool reduce = io->adj_shared != nullptr;
    if (reduce) {
        if (!io->shared_params || !io->reduce_scratch ||
            io->reduce_scratch_bytes < ppr_rollout_reduce_scratch_bytes(m, io->bs)) return PPR_E_WORKSPACE;
    } else if (!io->adj_target_ke || !io->adj_target_kd || !io->adj_body_inv_mass || !io->adj_body_inertia ||
               !io->adj_body_inv_inertia) return PPR_E_ARG;
    A.adj_pos = io->adj_out_pos; A.adj_vel = io->adj_out_vel; A.adj_loss_pos = io->adj_loss_pos; A.saved_pos = io->out_pos;
    A.adj_target_pos = io->adj_target_pos;
    A.adj_q_init = io->adj_q_init; A.adj_qd_init = io->adj_qd_init; A.adj_torques = io->adj_torques;
    A.adj_res_f = io->adj_res_f; A.adj_refs = io->adj_refs; A.adj_ke = io->adj_target_ke; A.adj_kd = io->adj_target_kd;
    A.adj_inv_m = io->adj_body_inv_mass; A.adj_I = io->adj_body_inertia; A.adj_inv_I = io->adj_body_inv_inertia;
    A.adj_partial = reduce ? (float*)io->reduce_scratch : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    rc = [&]() -> int {
#ifndef PPR_AB_ONLY
        if (m->ckpt_every > 1) PPR_LAUNCH_ROLLOUT(rollout_backward_kernel, true, , true);
#endif
        PPR_LAUNCH_ROLLOUT(rollout_backward_kernel, true, , false);
    }();
    if (rc != 0 || !reduce) return rc;
    const int P = 2 * m->d.nqd + 19 * m->d.nb;
    reduce_partials_kernel<<<(unsigned)((P + 31) / 32), dim3(32, 32), 0, st>>>((const float*)io->reduce_scratch, (int64_t)grid, P,
                                                                             io->adj_shared);
    g_launches++;
    return (int)cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------- control references
// refs[t] = lerp(key[t / stride], key[t / stride + 1], (t % stride) / stride): the per-substep control reference from
// per-FRAME values, what get_mocap_data's interp1d does on the host for every substep of every window
// (dp_model.py:421-427,605-609) -- a caller ships F x n floats instead of T x n.  One thread per (t, column).
__global__ void __launch_bounds__(256)
refs_from_frames_kernel(int64_t T, int64_t stride, int64_t nkey, int64_t n, const float* __restrict__ key,
                        float* __restrict__ refs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * n) return;
    const int64_t t = i / n, c = i - t * n;
    int64_t k0 = t / stride;
    float a = __fdiv_rn((float)(t - k0 * stride), (float)stride);
    if (k0 >= nkey - 1) { k0 = nkey - 2 < 0 ? 0 : nkey - 2; a = nkey > 1 ? __fdiv_rn((float)(t - k0 * stride), (float)stride) : 0.f; }
    const float v0 = key[k0 * n + c], v1 = nkey > 1 ? key[(k0 + 1) * n + c] : v0;
    refs[i] = (1.f - a) * v0 + a * v1;    // exact at both ends of a segment
}
// adjoint: adj_key[k] = sum_t w(t, k) adj_refs[t]; one thread per (k, column), fixed order
__global__ void __launch_bounds__(256)
refs_from_frames_adj_kernel(int64_t T, int64_t stride, int64_t nkey, int64_t n, const float* __restrict__ adj_refs,
                            float* __restrict__ adj_key) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nkey * n) return;
    const int64_t k = i / n, c = i - k * n;
    float s = 0.f;
    for (int64_t t = 0; t < T; ++t) {
        int64_t k0 = t / stride;
        if (k0 >= nkey - 1) k0 = nkey - 2 < 0 ? 0 : nkey - 2;
        const float a = nkey > 1 ? __fdiv_rn((float)(t - k0 * stride), (float)stride) : 0.f;
        if (k == k0) s += (1.f - a) * adj_refs[t * n + c];
        else if (k == k0 + 1) s += a * adj_refs[t * n + c];
    }
    adj_key[i] = nan0(s);
}
extern "C" int ppr_refs_from_frames(int64_t T, int64_t stride, int64_t nkey, int64_t n, const float* key, float* refs,
                                    void* stream) {
    if (T < 0 || stride < 1 || nkey < 1 || n < 0 || !key || !refs) return PPR_E_ARG;
    if (T * n == 0) return 0;
    refs_from_frames_kernel<<<(unsigned)((T * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, stride, nkey, n, key, refs);
    g_launches++;
    return (int)cudaGetLastError();
}
extern "C" int ppr_refs_from_frames_backward(int64_t T, int64_t stride, int64_t nkey, int64_t n, const float* adj_refs,
                                             float* adj_key, void* stream) {
    if (T < 0 || stride < 1 || nkey < 1 || n < 0 || !adj_refs || !adj_key) return PPR_E_ARG;
    if (nkey * n == 0) return 0;
    refs_from_frames_adj_kernel<<<(unsigned)((nkey * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, stride, nkey, n,
                                                                                                   adj_refs, adj_key);
    g_launches++;
    return (int)cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------- se3 loss
// One thread per (pred, gt) pair; see ppr_loss.h.  Replaces the ~330-400 torch kernels of one se3_loss call
// (forward + backward) by two launches.
__global__ void __launch_bounds__(256)
se3_loss_forward_kernel(int64_t n, int dim, const float* __restrict__ pred, const float* __restrict__ gt, float ratio,
                        float eps, float* __restrict__ loss) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p[7], g[7];
    for (int k = 0; k < dim; ++k) { p[k] = pred[i * dim + k]; g[k] = gt[i * dim + k]; }
    loss[i] = se3_pair_loss<float>(dim, p, g, ratio, eps);
}
__global__ void __launch_bounds__(256)
se3_loss_backward_kernel(int64_t n, int dim, const float* __restrict__ pred, const float* __restrict__ gt, float ratio,
                         float eps, const float* __restrict__ adj_loss, float* __restrict__ adj_pred,
                         float* __restrict__ adj_gt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p[7], g[7], ap[7], ag[7];
    for (int k = 0; k < dim; ++k) { p[k] = pred[i * dim + k]; g[k] = gt[i * dim + k]; }
    se3_pair_loss_adj<float>(dim, p, g, ratio, eps, adj_loss[i], ap, adj_gt ? ag : nullptr);
    for (int k = 0; k < dim; ++k) {
        adj_pred[i * dim + k] = nan0(ap[k]);
        if (adj_gt) adj_gt[i * dim + k] = nan0(ag[k]);
    }
}

extern "C" int ppr_se3_loss_forward(int64_t n, int32_t dim, const float* pred, const float* gt, float rot_ratio,
                                    float* loss, void* stream) {
    if (n < 0 || (dim != 6 && dim != 7) || !pred || !gt || !loss) return PPR_E_ARG;
    if (n == 0) return 0;
    se3_loss_forward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, dim, pred, gt, rot_ratio,
                                                                                         1e-4f, loss);
    g_launches++;
    return (int)cudaGetLastError();
}
extern "C" int ppr_se3_loss_backward(int64_t n, int32_t dim, const float* pred, const float* gt, float rot_ratio,
                                     const float* adj_loss, float* adj_pred, float* adj_gt, void* stream) {
    if (n < 0 || (dim != 6 && dim != 7) || !pred || !gt || !adj_loss || !adj_pred) return PPR_E_ARG;
    if (n == 0) return 0;
    se3_loss_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, dim, pred, gt, rot_ratio, 1e-4f, adj_loss, adj_pred, adj_gt);
    g_launches++;
    return (int)cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------- frame composition
// One thread per time sample; see ppr_frame.h.
__global__ void __launch_bounds__(256)
frame_compose_forward_kernel(int64_t n, const float* __restrict__ gq, const float* __restrict__ q,
                             const float* __restrict__ d, float* __restrict__ target, float* __restrict__ queried) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g[7], qi[7], di[6], t[7], u[7];
    for (int k = 0; k < 7; ++k) { g[k] = gq[k]; qi[k] = q[i * 7 + k]; }
    for (int k = 0; k < 6; ++k) di[k] = d[i * 6 + k];
    frame_compose<float>(g, qi, di, t, u);
    for (int k = 0; k < 7; ++k) { target[i * 7 + k] = t[k]; queried[i * 7 + k] = u[k]; }
}
__global__ void __launch_bounds__(256)
frame_compose_backward_kernel(int64_t n, const float* __restrict__ gq, const float* __restrict__ q,
                              const float* __restrict__ d, const float* __restrict__ at, const float* __restrict__ aq,
                              float* __restrict__ adj_g, float* __restrict__ adj_d) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g[7], qi[7], di[6], a[7], b[7], og[7], od[6];
    for (int k = 0; k < 7; ++k) { g[k] = gq[k]; qi[k] = q[i * 7 + k]; a[k] = at[i * 7 + k]; b[k] = aq[i * 7 + k]; }
    for (int k = 0; k < 6; ++k) di[k] = d[i * 6 + k];
    frame_compose_adj<float>(g, qi, di, a, b, og, od);
    for (int k = 0; k < 7; ++k) adj_g[i * 7 + k] = nan0(og[k]);
    for (int k = 0; k < 6; ++k) adj_d[i * 6 + k] = nan0(od[k]);
}

extern "C" int ppr_frame_compose_forward(int64_t n, const float* global_q, const float* q, const float* delta,
                                         float* target, float* queried, void* stream) {
    if (n < 0 || !global_q || !q || !delta || !target || !queried) return PPR_E_ARG;
    if (n == 0) return 0;
    frame_compose_forward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, global_q, q, delta,
                                                                                              target, queried);
    g_launches++;
    return (int)cudaGetLastError();
}
extern "C" int ppr_frame_compose_backward(int64_t n, const float* global_q, const float* q, const float* delta,
                                          const float* adj_target, const float* adj_queried, float* adj_global,
                                          float* adj_delta, void* stream) {
    if (n < 0 || !global_q || !q || !delta || !adj_target || !adj_queried || !adj_global || !adj_delta) return PPR_E_ARG;
    if (n == 0) return 0;
    frame_compose_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, global_q, q, delta, adj_target, adj_queried, adj_global, adj_delta);
    g_launches++;
    return (int)cudaGetLastError();
}
