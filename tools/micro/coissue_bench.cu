// Microbenchmark: do integer (ALU-pipe) instructions issue for free next to FFMA2 / FFMA orbit math on sm_100a?
// Each thread runs 4 orbits (as 2 packed pairs, or 4 scalars) and, per orbit iteration, ALU independent integer
// LOP3/IADD3-class instructions.  If time does not move with ALU, the issue port has spare slots next to FFMA2.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIter = 32;

template <int ALU, bool PACKED>
__global__ void __launch_bounds__(256) orbit_alu(float* out, unsigned* iout, int reps) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    float cys = (t & 1023) * (1.0f / 1024.0f) - 0.5f;
    const float2 cy = make_float2(cys, cys);
    float acc = 0.f;
    unsigned a0 = t, a1 = t * 3, a2 = t * 5, a3 = t * 7;
    for (int r = 0; r < reps; ++r) {
        float2 cx[2], zx[2], zy[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            cx[k] = make_float2((t * 4 + 2 * k + r) * 1e-7f - 1.0f, (t * 4 + 2 * k + 1 + r) * 1e-7f - 1.0f);
            zx[k] = make_float2(0.f, 0.f); zy[k] = make_float2(0.f, 0.f);
        }
#pragma unroll 8
        for (int it = 0; it < kIter; ++it) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (PACKED) {
                    float2 tt = __ffma2_rn(zx[k], zx[k], cx[k]);
                    float2 nzx = __ffma2_rn(make_float2(-zy[k].x, -zy[k].y), zy[k], tt);
                    float2 nzy = __ffma2_rn(__fadd2_rn(zx[k], zx[k]), zy[k], cy);
                    zx[k] = nzx; zy[k] = nzy;
                } else {
                    float t0 = fmaf(zx[k].x, zx[k].x, cx[k].x), t1 = fmaf(zx[k].y, zx[k].y, cx[k].y);
                    float n0 = fmaf(-zy[k].x, zy[k].x, t0), n1 = fmaf(-zy[k].y, zy[k].y, t1);
                    float y0 = fmaf(zx[k].x + zx[k].x, zy[k].x, cys), y1 = fmaf(zx[k].y + zx[k].y, zy[k].y, cys);
                    zx[k] = make_float2(n0, n1); zy[k] = make_float2(y0, y1);
                }
            }
#pragma unroll
            for (int j = 0; j < ALU; ++j) {  // 4 independent integer chains, xor/add with loop-variant operands
                if ((j & 3) == 0) a0 = (a0 ^ (a1 + it)) + 0x9E37u;
                if ((j & 3) == 1) a1 = (a1 ^ (a2 + it)) + 0x79B9u;
                if ((j & 3) == 2) a2 = (a2 ^ (a3 + it)) + 0x7F4Au;
                if ((j & 3) == 3) a3 = (a3 ^ (a0 + it)) + 0x7C15u;
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) acc += zx[k].x + zx[k].y + zy[k].x + zy[k].y;
    }
    out[t] = acc;
    iout[t] = a0 ^ a1 ^ a2 ^ a3;
}

template <int ALU, bool PACKED>
void run(float* out, unsigned* iout, int blocks, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) orbit_alu<ALU, PACKED><<<blocks, 256>>>(out, iout, reps);
    cudaEventRecord(a);
    for (int w = 0; w < 10; ++w) orbit_alu<ALU, PACKED><<<blocks, 256>>>(out, iout, reps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double px = 10.0 * blocks * 256 * 4.0 * reps;
    printf("%s  alu/iter %2d: %.3f ms  %.1f G orbit-pixels/s\n", PACKED ? "packed" : "scalar", ALU, ms / 10, px / (ms * 1e-3) / 1e9);
}

int main() {
    const int blocks = 148 * 32, reps = 64;
    float* out; unsigned* iout;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaMalloc(&iout, blocks * 256 * sizeof(unsigned));
    run<0, true>(out, iout, blocks, reps);  run<4, true>(out, iout, blocks, reps);  run<8, true>(out, iout, blocks, reps);
    run<16, true>(out, iout, blocks, reps); run<24, true>(out, iout, blocks, reps);
    run<0, false>(out, iout, blocks, reps); run<4, false>(out, iout, blocks, reps); run<8, false>(out, iout, blocks, reps);
    run<16, false>(out, iout, blocks, reps);
    return 0;
}
