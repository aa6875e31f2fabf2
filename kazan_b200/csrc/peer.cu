// peer.cu — split-frame plumbing over NVLink peer memory: memory export / import between the per-GPU
// processes of one box, and stream-ordered signal / wait on 64-bit flags that live in (peer) device memory.
//
// With these, a band-rendering rank binds the PRESENTING GPU's image as its colour attachment: the raster kernels' tile
// epilogue (one 16-byte store per lane, 512 B contiguous per warp) writes the band straight across NVLink into the
// presenter's HBM, and the composite collapses into a flag increment — there is no copy step and no collective on the
// data path.  The reference has nothing to compare with: it exposes one physical device (vulkan-driver/src/api_impl.rs:5092-5119).
#include <new>

#include "kzn_internal.h"

namespace {

constexpr unsigned long long kWaitTimeoutNs = 5ull * 1000 * 1000 * 1000;  // a lost peer must not hang the GPU

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Everything this stream did before (kernel boundaries order it) becomes visible system-wide, then the flag moves.
__global__ void signal_flag_kernel(unsigned long long* flag, unsigned long long add) {
    __threadfence_system();
    atomicAdd_system(flag, add);
}

// Spin until *flag >= expected (acquire at system scope), with a timeout that records failure in *status.
__global__ void wait_flag_kernel(const unsigned long long* flag, unsigned long long expected, uint32_t* status) {
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= expected) return;
        if (globaltimer_ns() - t0 > kWaitTimeoutNs) {
            if (status) atomicExch_system(status, 1u);
            return;
        }
        __nanosleep(200);
    }
}

// the same for `count` consecutive flags, one thread each: returns when every one of them is >= expected
__global__ void wait_flags_kernel(const unsigned long long* flags, unsigned long long expected, uint32_t* status) {
    const unsigned long long t0 = globaltimer_ns();
    const unsigned long long* flag = flags + threadIdx.x;
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= expected) return;
        if (globaltimer_ns() - t0 > kWaitTimeoutNs) {
            if (status) atomicExch_system(status, 1u);
            return;
        }
        __nanosleep(200);
    }
}

}  // namespace

extern "C" {

KznResult kzn_memory_export(KznMemory* memory, uint8_t handle[KZN_MEMORY_HANDLE_BYTES]) {
    KZN_VALID(memory && handle, "kzn_memory_export: null argument");
    KZN_VALID(memory->type == KZN_MEMORY_TYPE_DEVICE_LOCAL && !memory->imported, "only DEVICE_LOCAL allocations of this process can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == KZN_MEMORY_HANDLE_BYTES, "handle size");
    KZN_CUDA(cudaSetDevice(memory->device->ordinal));
    cudaIpcMemHandle_t h;
    KZN_CUDA(cudaIpcGetMemHandle(&h, memory->device_ptr));
    memcpy(handle, &h, sizeof h);
    return KZN_SUCCESS;
}

KznResult kzn_memory_import(KznDevice* device, const uint8_t handle[KZN_MEMORY_HANDLE_BYTES], uint64_t size, KznMemory** memory) {
    KZN_VALID(device && handle && memory && size, "kzn_memory_import: bad argument");
    *memory = nullptr;
    KZN_CUDA(cudaSetDevice(device->ordinal));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        kzn_set_error("cudaIpcOpenMemHandle failed: %s (the exporting GPU must be peer-accessible from device %d)", cudaGetErrorString(e),
                      device->ordinal);
        return KZN_ERROR_INITIALIZATION_FAILED;
    }
    KznMemory* m = new (std::nothrow) KznMemory();
    if (!m) {
        cudaIpcCloseMemHandle(p);
        kzn_set_error("out of host memory");
        return KZN_ERROR_OUT_OF_HOST_MEMORY;
    }
    m->device = device;
    m->type = KZN_MEMORY_TYPE_DEVICE_LOCAL;
    m->size = size;
    m->device_ptr = p;
    m->host_ptr = nullptr;
    m->imported = true;
    *memory = m;
    return KZN_SUCCESS;
}

KznResult kzn_cmd_signal_flag(KznQueue* queue, uint64_t flag_address, uint64_t add) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(flag_address && flag_address % 8 == 0, "flag address must be 8-byte aligned");
    signal_flag_kernel<<<1, 1, 0, queue->stream>>>((unsigned long long*)(uintptr_t)flag_address, add);
    return kzn_launched(queue, "signal_flag_kernel");
}

KznResult kzn_cmd_wait_flag(KznQueue* queue, uint64_t flag_address, uint64_t expected, uint64_t status_address) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(flag_address && flag_address % 8 == 0, "flag address must be 8-byte aligned");
    KZN_VALID(status_address % 4 == 0, "status address must be 4-byte aligned");
    wait_flag_kernel<<<1, 1, 0, queue->stream>>>((const unsigned long long*)(uintptr_t)flag_address, expected, (uint32_t*)(uintptr_t)status_address);
    return kzn_launched(queue, "wait_flag_kernel");
}

KznResult kzn_cmd_wait_flags(KznQueue* queue, uint64_t first_flag_address, uint32_t flag_count, uint64_t expected, uint64_t status_address) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(first_flag_address && first_flag_address % 8 == 0, "flag address must be 8-byte aligned");
    KZN_VALID(flag_count >= 1 && flag_count <= 1024, "kzn_cmd_wait_flags: 1 .. 1024 flags");
    KZN_VALID(status_address % 4 == 0, "status address must be 4-byte aligned");
    wait_flags_kernel<<<1, flag_count, 0, queue->stream>>>((const unsigned long long*)(uintptr_t)first_flag_address, expected, (uint32_t*)(uintptr_t)status_address);
    return kzn_launched(queue, "wait_flags_kernel");
}

}  // extern "C"
