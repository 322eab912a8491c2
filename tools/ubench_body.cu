// Micro-benchmark (B200): per-body stage functions (joint wrench + integrate, and their adjoints) instantiated with
// T = float (one environment per thread) vs T = F2 (two environments per thread, packed FFMA2 arithmetic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -prec-div=false -prec-sqrt=false -ftz=true \
//        -o tools/_bin/ubench_body tools/ubench_body.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include "../ppr_diffphys_b200/csrc/ppr_body.h"
#include "ubench_f2.h"
using namespace ppr;

template <class S> struct Ld;
template <> struct Ld<float> { static __device__ float get(const float* p, int i, int) { return p[i]; } static constexpr int W = 1; };
template <> struct Ld<F2> { static __device__ F2 get(const float* p, int i, int n) { return F2(p[i], p[i + n]); } static constexpr int W = 2; };
__device__ float total(float a) { return a; }
__device__ float total(F2 a) { return a.v.x + a.v.y; }

template <class S, int JM, bool ADJ, int MINB>
__global__ void __launch_bounds__(96, MINB) k(const float* in, float* out, int nsteps, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto L = [&](int c) { return Ld<S>::get(in + c * 2 * n, i * Ld<S>::W % n, 1); };
    Body<S> s, P;
    s.x = v3<S>(L(0), L(1), L(2)); s.r = q4<S>(L(3), L(4), L(5), L(6)); s.w = v3<S>(L(7), L(8), L(9)); s.v = v3<S>(L(10), L(11), L(12));
    P.x = v3<S>(L(13), L(14), L(15)); P.r = q4<S>(L(16), L(17), L(18), L(19)); P.w = v3<S>(L(20), L(21), L(22)); P.v = v3<S>(L(23), L(24), L(25));
    JointStatic<S> js; js.type = JM == JM_REVOLUTE ? JT_REVOLUTE : JT_COMPOUND;
    js.xpj = v3<S>(L(26), L(27), L(28)); js.qpj = q4<S>(S(0.f), S(0.f), S(0.f), S(1.f)); js.qoff = js.qpj; js.axis = v3<S>(S(0.f), S(0.f), S(1.f));
    JointCtl<S> c;
    for (int k2 = 0; k2 < 3; ++k2) { c.target[k2] = L(29 + k2); c.act[k2] = S(0.f); c.ke[k2] = S(220.f); c.kd[k2] = S(2.f); c.lo[k2] = S(-1e3f); c.hi[k2] = S(1e3f); c.lke[k2] = S(0.f); c.lkd[k2] = S(0.f); }
    S I[9], J[9];
    for (int k2 = 0; k2 < 9; ++k2) { I[k2] = S(k2 % 4 == 0 ? 0.02f : 0.001f); J[k2] = S(k2 % 4 == 0 ? 50.f : -1.f); }
    V3<S> com = v3<S>(S(0.01f), S(-0.05f), S(0.f)), g = v3<S>(S(0.f), S(-9.8f), S(0.f)), xcp = P.x;
    S dt = S(5e-4f), inv_m = S(0.5f), acc = S(0.f);
    Body<S> adjN = s;
#pragma unroll 1
    for (int t = 0; t < nsteps; ++t) {
        M3<S> Rb = qmat(s.r);
        V3<S> xc = s.x + mrot(Rb, com);
        if (!ADJ) {
            V3<S> tq, f, ap, ac;
            S ang[3];
            joint_fwd<S, JM, false, false>(js, c, S(16000.f), S(200.f), P, xcp, true, s, Rb, xc, tq, f, ap, ac, ang);
            Wrench<S> F; F.t = -(tq + cross(ac, f)); F.f = -f;
            acc += ang[0];
            s = integrate_fwd(s, Rb, xc, com, F, inv_m, I, J, g, dt);
        } else {
            Wrench<S> F; F.t = s.w; F.f = s.v;
            Body<S> adjS = body_zero<S>(); M3<S> G = m3_zero<S>(); V3<S> adj_xc = vzero<S>(); Wrench<S> adjF;
            S a_inv_m = S(0.f); V3<S> ga, gb, gc, gd;
            integrate_adj_core(s, Rb, xc, com, F, inv_m, I, J, g, dt, adjN, adjS, G, adj_xc, adjF, a_inv_m, ga, gb, gc, gd);
            acc += a_inv_m + ga.x * gb.x + gc.y * gd.z;
            Body<S> adjP = body_zero<S>(); V3<S> adj_xcp = vzero<S>();
            S gt[3] = {S(0.f), S(0.f), S(0.f)}, gact[3] = {S(0.f), S(0.f), S(0.f)}, ake_[3] = {S(0.f), S(0.f), S(0.f)}, akd_[3] = {S(0.f), S(0.f), S(0.f)};
            S ang[3] = {c.target[0], c.target[1], c.target[2]};
            joint_adj<S, JM, false, false>(js, c, S(16000.f), S(200.f), P, xcp, true, s, Rb, xc, adjF, adjF, adjP, adj_xcp, adjS, G, adj_xc, gt, gact, ake_, akd_, ang);
            acc += gt[0] + ake_[0] + akd_[0] + adjP.x.x + adjP.r.w + adj_xcp.y;
            adjS.x += adj_xc; m3_acc(G, adj_xc, com); adjS.r += qmat_adj(s.r, G);
            adjN = adjS;
            s.r = qnormalize(s.r + adjS.r * S(1e-9f), a_inv_m); s.w = s.w * S(0.999f) + adjS.w * S(1e-9f);
        }
    }
    out[i] = total(acc) + total(s.x.x + s.r.w + adjN.v.x);
}
template <class S, int JM, bool ADJ, int MINB> void run(const char* name, const float* in, float* out, int n_env) {
    int nsteps = 256, nthr = n_env / Ld<S>::W;
    cudaFuncAttributes a; cudaFuncGetAttributes(&a, k<S, JM, ADJ, MINB>);
    k<S, JM, ADJ, MINB><<<nthr / 96, 96>>>(in, out, 8, n_env); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<S, JM, ADJ, MINB><<<nthr / 96, 96>>>(in, out, nsteps, n_env);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s regs %3d  %8.3f ms  %8.2f G body-substeps/s  (%s)\n", name, a.numRegs, ms, (double)n_env * nsteps / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int n_env = 96 * 148 * 64;
    float* in; float* out; cudaMalloc(&in, 40 * 2 * n_env * 4); cudaMalloc(&out, n_env * 4);
    float* h = (float*)malloc(40 * 2 * n_env * 4);
    for (size_t i = 0; i < (size_t)40 * 2 * n_env; ++i) h[i] = 0.05f * ((i * 2654435761u >> 8) % 1000) / 1000.f;
    for (int c : {6, 19}) for (int i = 0; i < 2 * n_env; ++i) h[(size_t)c * 2 * n_env + i] = 1.f;
    cudaMemcpy(in, h, 40 * 2 * n_env * 4, cudaMemcpyHostToDevice);
    run<float, JM_REVOLUTE, false, 5>("fwd revolute  float (5 blk/SM)", in, out, n_env);
    run<F2, JM_REVOLUTE, false, 2>("fwd revolute  F2    (2 blk/SM)", in, out, n_env);
    run<F2, JM_REVOLUTE, false, 3>("fwd revolute  F2    (3 blk/SM)", in, out, n_env);
    run<F2, JM_REVOLUTE, false, 4>("fwd revolute  F2    (4 blk/SM)", in, out, n_env);
    run<float, JM_COMPOUND, false, 5>("fwd compound  float (5 blk/SM)", in, out, n_env);
    run<F2, JM_COMPOUND, false, 2>("fwd compound  F2    (2 blk/SM)", in, out, n_env);
    run<F2, JM_COMPOUND, false, 3>("fwd compound  F2    (3 blk/SM)", in, out, n_env);
    run<float, JM_REVOLUTE, true, 4>("adj revolute  float (4 blk/SM)", in, out, n_env);
    run<F2, JM_REVOLUTE, true, 2>("adj revolute  F2    (2 blk/SM)", in, out, n_env);
    run<F2, JM_REVOLUTE, true, 3>("adj revolute  F2    (3 blk/SM)", in, out, n_env);
    run<float, JM_COMPOUND, true, 4>("adj compound  float (4 blk/SM)", in, out, n_env);
    run<F2, JM_COMPOUND, true, 2>("adj compound  F2    (2 blk/SM)", in, out, n_env);
    run<F2, JM_COMPOUND, true, 3>("adj compound  F2    (3 blk/SM)", in, out, n_env);
    return 0;
}
