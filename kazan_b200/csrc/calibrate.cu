// calibrate.cu — measured denominators for the shader-bound roofline ("calibrate with an FMA chain kernel in
// the same run").  The fill-rate fragment program is fp32-issue bound by construction, and MEASURED_PEAKS.json carries
// only HBM and tensor numbers, so the benchmark measures the fp32 pipe itself: independent packed FMA chains
// (fma.rn.f32x2, SASS FFMA2 — the instruction the shader uses), nothing else in the loop.
#include "kzn_internal.h"

namespace {

constexpr int kChains = 8;      // independent float2 accumulators per thread
constexpr int kInner = 1024;    // FFMA2 per chain per launch

__global__ void __launch_bounds__(256) fp32_issue_kernel(float* sink, float seed) {
    float2 acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = make_float2(seed + c, seed - c);
    const float2 m = make_float2(0.999f, 1.001f), a = make_float2(1e-3f, -1e-3f);
#pragma unroll 1
    for (int i = 0; i < kInner / 8; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) acc[c] = __ffma2_rn(acc[c], m, a);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c].x + acc[c].y;
    if (s == 123.456f) sink[0] = s;  // never true: keeps the chains alive
}

}  // namespace

extern "C" KznResult kzn_calibrate_fp32_issue(KznQueue* queue, float* tera_lane_ops_per_second) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(tera_lane_ops_per_second, "null out pointer");
    KZN_VALID(!queue->recording, "calibration cannot be recorded");
    float* sink = nullptr;
    KZN_CUDA(cudaMalloc(&sink, 4));
    cudaEvent_t t0, t1;
    KZN_CUDA(cudaEventCreate(&t0));
    KZN_CUDA(cudaEventCreate(&t1));
    const int blocks = queue->device->sm_count * 8, reps = 20;
    for (int w = 0; w < 3; ++w) fp32_issue_kernel<<<blocks, 256, 0, queue->stream>>>(sink, 1.0f);
    cudaEventRecord(t0, queue->stream);
    for (int k = 0; k < reps; ++k) fp32_issue_kernel<<<blocks, 256, 0, queue->stream>>>(sink, 1.0f);
    cudaEventRecord(t1, queue->stream);
    cudaError_t e = cudaEventSynchronize(t1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t0, t1);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(sink);
    if (e != cudaSuccess) return kzn_cuda_failed(e, "fp32 calibration", __FILE__, __LINE__);
    queue->launches += reps + 3;
    // one FFMA2 = 2 fp32 lane-operations (each an fma)
    const double lane_ops = (double)reps * blocks * 256.0 * kChains * kInner * 2.0;
    *tera_lane_ops_per_second = (float)(lane_ops / (ms * 1e-3) / 1e12);
    return KZN_SUCCESS;
}
