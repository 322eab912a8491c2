// Micro-benchmark (B200): issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2), alone and mixed with integer ALU
// work, to decide whether packing the per-body FP32 algebra of the rollout kernels can relieve the issue-slot bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_fp32 tools/ubench_fp32.cu
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 4096
#define NCH 8
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
    float a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    float2 x[NCH];
    int ia[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { x[i] = make_float2(in[i] + threadIdx.x, in[i + 8] - threadIdx.x); ia[i] = threadIdx.x + i; }
    float2 a2 = make_float2(a, a + 1.f), b2 = make_float2(b, b - 1.f);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a2.y, b2.y); }          // 2 FFMA
                if (MODE == 1) { x[i] = __ffma2_rn(x[i], a2, b2); }                                           // 1 FFMA2
                if (MODE == 2) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a2.y, b2.y);
                                 ia[i] = (ia[i] ^ it) + i; ia[i] = (ia[i] & 0xffff) + u; }                    // 2 FFMA + ~2-4 ALU
                if (MODE == 3) { x[i] = __ffma2_rn(x[i], a2, b2);
                                 ia[i] = (ia[i] ^ it) + i; ia[i] = (ia[i] & 0xffff) + u; }                    // 1 FFMA2 + same ALU
                if (MODE == 4) { x[i].x = x[i].x * a; x[i].y = x[i].y + b; }                                  // FMUL + FADD
                if (MODE == 5) { x[i] = __fmul2_rn(x[i], a2); x[i] = __fadd2_rn(x[i], b2); }                  // FMUL2 + FADD2
                if (MODE == 6) { x[i] = __fmul2_rn(x[i], a2); }                                               // FMUL2
                if (MODE == 7) { x[i] = __fadd2_rn(x[i], b2); }                                               // FADD2
                if (MODE == 8) { x[i].x = x[i].x * a; x[i].y = x[i].y * a2.y; }                               // 2 FMUL
                if (MODE == 9) { x[i].x = x[i].x + b; x[i].y = x[i].y + b2.y; }                               // 2 FADD
                if (MODE == 10) { x[i] = __ffma2_rn(x[i], a2, x[(i + 1) % NCH]); }                            // FFMA2, 3 distinct regs
                if (MODE == 11) { x[i].x = fmaf(x[i].x, a, x[(i + 1) % NCH].x); x[i].y = fmaf(x[i].y, a2.y, x[(i + 1) % NCH].y); }
            }
        }
    }
    float s = 0.f; int si = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) { s += x[i].x + x[i].y; si += ia[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)si;
}
template <int MODE> void run(const char* name, double fp_per_iter_thread, float* out, float* in) {
    int nsm = 148, blocks = nsm * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(out, in, 16);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, in, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = fp_per_iter_thread * 4.0 * NCH * (double)ITERS * blocks * threads;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s (fp32 flop)  err=%s\n", name, ms, fl / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out, *in; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096);
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + 1e-6f * i; cudaMemcpy(in, h, 256, cudaMemcpyHostToDevice);
    run<0>("2xFFMA", 4, out, in);
    run<1>("1xFFMA2", 4, out, in);
    run<2>("2xFFMA + int ALU", 4, out, in);
    run<3>("1xFFMA2 + int ALU", 4, out, in);
    run<4>("FMUL + FADD", 2, out, in);
    run<5>("FMUL2 + FADD2 (fused)", 4, out, in);
    run<6>("FMUL2", 2, out, in);
    run<7>("FADD2", 2, out, in);
    run<8>("2xFMUL", 2, out, in);
    run<9>("2xFADD", 2, out, in);
    run<10>("FFMA2 (3 varying operands)", 4, out, in);
    run<11>("2xFFMA (3 varying operands)", 4, out, in);
    return 0;
}
