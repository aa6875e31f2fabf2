// Microbenchmark: scalar FFMA orbit vs packed FFMA2 orbit (fma.rn.f32x2) issue rate on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIter = 32;

__global__ void __launch_bounds__(256) orbit_scalar(float* out, int reps) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    float cx[4], zx[4], zy[4];
    float cy = (t & 1023) * (1.0f / 1024.0f) - 0.5f;
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { cx[k] = (t * 4 + k + r) * 1e-7f - 1.0f; zx[k] = 0.f; zy[k] = 0.f; }
#pragma unroll 4
        for (int it = 0; it < kIter; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float tt = fmaf(zx[k], zx[k], cx[k]);
                float nzx = fmaf(-zy[k], zy[k], tt);
                float nzy = fmaf(zx[k] + zx[k], zy[k], cy);
                zx[k] = nzx; zy[k] = nzy;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += zx[k] + zy[k];
    }
    out[t] = acc;
}

__global__ void __launch_bounds__(256) orbit_packed(float* out, int reps) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    float2 cx[2], zx[2], zy[2];
    float cys = (t & 1023) * (1.0f / 1024.0f) - 0.5f;
    const float2 cy = make_float2(cys, cys);
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            cx[k] = make_float2((t * 4 + 2 * k + r) * 1e-7f - 1.0f, (t * 4 + 2 * k + 1 + r) * 1e-7f - 1.0f);
            zx[k] = make_float2(0.f, 0.f); zy[k] = make_float2(0.f, 0.f);
        }
#pragma unroll 4
        for (int it = 0; it < kIter; ++it) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float2 tt = __ffma2_rn(zx[k], zx[k], cx[k]);
                float2 nzy_neg = make_float2(-zy[k].x, -zy[k].y);
                float2 nzx = __ffma2_rn(nzy_neg, zy[k], tt);
                float2 d = __fadd2_rn(zx[k], zx[k]);
                float2 nzy = __ffma2_rn(d, zy[k], cy);
                zx[k] = nzx; zy[k] = nzy;
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) acc += zx[k].x + zx[k].y + zy[k].x + zy[k].y;
    }
    out[t] = acc;
}

int main() {
    const int blocks = 148 * 32, reps = 64;
    float* out;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float h[2][4];
    for (int which = 0; which < 2; ++which) {
        for (int w = 0; w < 3; ++w) { if (which) orbit_packed<<<blocks, 256>>>(out, reps); else orbit_scalar<<<blocks, 256>>>(out, reps); }
        cudaEventRecord(a);
        for (int w = 0; w < 10; ++w) { if (which) orbit_packed<<<blocks, 256>>>(out, reps); else orbit_scalar<<<blocks, 256>>>(out, reps); }
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double orbits = 10.0 * blocks * 256 * 4.0 * reps;
        cudaMemcpy(h[which], out + 12345, 16, cudaMemcpyDeviceToHost);
        printf("%s: %.3f ms, %.1f G orbit-pixels/s, %.2f T fp32-lane-ops/s (4 per iteration)  sample %g\n", which ? "packed f32x2" : "scalar      ", ms / 10,
               orbits / (ms * 1e-3) / 1e9, orbits * kIter * 4 / (ms * 1e-3) / 1e12, h[which][0]);
    }
    printf("bit-identical samples: %s\n", (h[0][0] == h[1][0] && h[0][1] == h[1][1]) ? "yes" : "NO");
    return 0;
}
