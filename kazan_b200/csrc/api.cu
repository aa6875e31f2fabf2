// api.cu — device, memory, queue and fence objects of the kazan_b200 C ABI.
//
// Replaces (paths in the kazan-3d/kazan tree):
//   vkAllocateMemory / vkMapMemory      vulkan-driver/src/api_impl.rs:3434-3500, device_memory.rs:191-273
//   vkQueueSubmit / vkQueueWaitIdle     vulkan-driver/src/api_impl.rs:3414-3431  (unimplemented!())
//   vkCreateFence .. vkWaitForFences    vulkan-driver/src/api_impl.rs:3653-3697  (unimplemented!())
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <new>
#include <vector>

#include <stdlib.h>

#include "kzn_internal.h"

// vulkan-driver/src/handle.rs:315 `CommandBuffer {}` (a placeholder upstream): here a CUDA graph captured from the queue's
// stream between vkBeginCommandBuffer and vkEndCommandBuffer, instantiated once and replayed by vkQueueSubmit.
struct KznCommandBuffer {
    KznQueue* queue = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool recording = false;
    uint64_t launches = 0;            // kernels per replay
    uint64_t launches_at_begin = 0;
    std::vector<void*> host_copies;   // vkCmdUpdateBuffer data, captured at record time
    KznRasterScratch* raster_scratch = nullptr;  // pinned: a recorded draw names its buffers
};

void kzn_command_buffer_uses_raster_scratch(KznCommandBuffer* cb, KznRasterScratch* scratch) {
    if (!cb || !scratch || cb->raster_scratch) return;
    cb->raster_scratch = scratch;
    kzn_raster_scratch_pin(scratch);
}

void kzn_raster_scratch_destroy(KznRasterScratch* s);  // raster.cu

static thread_local char g_error[512];

void kzn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

KznResult kzn_cuda_failed(cudaError_t e, const char* what, const char* file, int line) {
    kzn_set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    if (e == cudaErrorMemoryAllocation) return KZN_ERROR_OUT_OF_DEVICE_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError ||
        e == cudaErrorStubLibrary)
        return KZN_ERROR_INITIALIZATION_FAILED;
    return KZN_ERROR_DEVICE_LOST;
}

extern "C" {

uint32_t kzn_abi_version(void) { return KZN_ABI_VERSION; }
const char* kzn_last_error(void) { return g_error; }

// ---- device ------------------------------------------------------------------------------------

KznResult kzn_device_count(int32_t* count) {
    KZN_VALID(count, "kzn_device_count: null out pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return kzn_cuda_failed(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    *count = n;
    return KZN_SUCCESS;
}

KznResult kzn_device_create(int32_t ordinal, KznDevice** device) {
    KZN_VALID(device, "kzn_device_create: null out pointer");
    *device = nullptr;
    int n = 0;
    KZN_CUDA(cudaGetDeviceCount(&n));
    KZN_REQUIRE(n > 0, KZN_ERROR_INITIALIZATION_FAILED, "no CUDA device: kazan_b200 has no CPU fallback");
    KZN_VALID(ordinal >= 0 && ordinal < n, "device ordinal %d out of range [0,%d)", ordinal, n);
    KZN_CUDA(cudaSetDevice(ordinal));
    KZN_CUDA(cudaFree(0));  // create the primary context
    KznDevice* d = new (std::nothrow) KznDevice();
    KZN_REQUIRE(d, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    d->ordinal = ordinal;
    KZN_CUDA(cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, ordinal));
    KZN_CUDA(cudaDeviceGetAttribute(&d->l2_bytes, cudaDevAttrL2CacheSize, ordinal));
    KZN_CUDA(cudaDeviceGetAttribute(&d->cc_major, cudaDevAttrComputeCapabilityMajor, ordinal));
    KZN_CUDA(cudaDeviceGetAttribute(&d->cc_minor, cudaDevAttrComputeCapabilityMinor, ordinal));
    if (d->cc_major != 10) {
        // the cubin in this library is sm_100a only; anything else cannot run it
        kzn_set_error("device %d is sm_%d%d; kazan_b200 kernels are built for sm_100a only", ordinal,
                      d->cc_major, d->cc_minor);
        delete d;
        return KZN_ERROR_INITIALIZATION_FAILED;
    }
    {   // KZN_L2_PERSIST=1: L2 lines set aside for the triangle records of binned draws between setup and raster (raster.cu
        // draw_common).  Off by default: measured on B200 it takes a third off the DRAM traffic of the 1M-triangle frame
        // (270 -> 180 MB) without making it faster (the draw is issue-bound), and the set-aside costs every other kernel L2
        // (the 4M-triangle frame: 2.10 -> 2.25 ms).
        int max_persist = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ordinal);
        cudaDeviceGetAttribute(&d->l2_window_bytes, cudaDevAttrMaxAccessPolicyWindowSize, ordinal);
        const char* e = getenv("KZN_L2_PERSIST");
        if (!(e && e[0] == '1')) max_persist = 0;
        if (max_persist > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess) d->l2_persist_bytes = max_persist;
        cudaGetLastError();
    }
    *device = d;
    return KZN_SUCCESS;
}

void kzn_device_destroy(KznDevice* device) { delete device; }

KznResult kzn_device_wait_idle(KznDevice* device) {
    KZN_VALID(device, "null device");
    KZN_CUDA(cudaSetDevice(device->ordinal));
    KZN_CUDA(cudaDeviceSynchronize());
    return KZN_SUCCESS;
}

int32_t kzn_device_sm_count(const KznDevice* device) { return device ? device->sm_count : 0; }
int32_t kzn_device_ordinal(const KznDevice* device) { return device ? device->ordinal : -1; }

// ---- memory --------------------------------------------------------------------------------------

KznResult kzn_memory_allocate(KznDevice* device, uint64_t size, uint32_t memory_type, KznMemory** memory) {
    KZN_VALID(device && memory, "kzn_memory_allocate: null argument");
    *memory = nullptr;
    KZN_VALID(size > 0, "allocationSize must be > 0");
    KZN_VALID(memory_type <= KZN_MEMORY_TYPE_DEVICE_LOCAL, "memory type index %u does not exist", memory_type);
    // DeviceMemoryLayout::calculate(size, MIN_MEMORY_MAP_ALIGNMENT) device_memory.rs:166-181
    uint64_t rounded = kzn_round_up(size, KZN_MIN_MEMORY_MAP_ALIGNMENT);
    KZN_REQUIRE(rounded >= size, KZN_ERROR_OUT_OF_DEVICE_MEMORY, "allocation size overflows");
    KZN_CUDA(cudaSetDevice(device->ordinal));
    KznMemory* m = new (std::nothrow) KznMemory();
    KZN_REQUIRE(m, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    m->device = device;
    m->type = memory_type;
    m->size = rounded;
    m->imported = false;
    cudaError_t e;
    if (memory_type == KZN_MEMORY_TYPE_HOST_COHERENT) {
        // the reference's single memory type is DEVICE_LOCAL|HOST_VISIBLE|HOST_COHERENT|HOST_CACHED
        // (device_memory.rs:30-39) and vkMapMemory hands out a raw pointer: pinned, mapped host memory
        e = cudaHostAlloc(&m->host_ptr, rounded, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(&m->device_ptr, m->host_ptr, 0);
    } else {
        e = cudaMalloc(&m->device_ptr, rounded);
    }
    if (e != cudaSuccess) {
        delete m;
        cudaGetLastError();
        kzn_set_error("allocation of %llu bytes (type %u) failed: %s", (unsigned long long)rounded, memory_type,
                      cudaGetErrorString(e));
        return KZN_ERROR_OUT_OF_DEVICE_MEMORY;  // api_impl.rs:3459-3471
    }
    *memory = m;
    return KZN_SUCCESS;
}

void kzn_memory_free(KznMemory* memory) {
    if (!memory) return;
    cudaSetDevice(memory->device->ordinal);
    if (memory->imported) cudaIpcCloseMemHandle(memory->device_ptr);
    else if (memory->type == KZN_MEMORY_TYPE_HOST_COHERENT) cudaFreeHost(memory->host_ptr);
    else cudaFree(memory->device_ptr);
    delete memory;
}

uint64_t kzn_memory_device_address(const KznMemory* memory) {
    return memory ? (uint64_t)(uintptr_t)memory->device_ptr : 0;
}
uint64_t kzn_memory_size(const KznMemory* memory) { return memory ? memory->size : 0; }

KznResult kzn_memory_map(KznMemory* memory, uint64_t offset, void** host_ptr) {
    KZN_VALID(memory && host_ptr, "kzn_memory_map: null argument");
    KZN_VALID(memory->type == KZN_MEMORY_TYPE_HOST_COHERENT, "memory type %u is not HOST_VISIBLE", memory->type);
    KZN_VALID(offset < memory->size, "map offset %llu beyond allocation of %llu bytes",
              (unsigned long long)offset, (unsigned long long)memory->size);
    *host_ptr = (uint8_t*)memory->host_ptr + offset;  // api_impl.rs:3488-3500
    return KZN_SUCCESS;
}

// ---- queue -----------------------------------------------------------------------------------------

}  // extern "C"

KznResult kzn_queue_enter_transfer(KznQueue* q) {
    KZN_VALID(q && q->device, "null queue");
    KZN_CUDA(cudaSetDevice(q->device->ordinal));
    return q->pending_clear.active ? kzn_flush_render_pass_clear(q) : KZN_SUCCESS;  // LOAD_OP_CLEAR that no draw took over
}

KznResult kzn_queue_enter(KznQueue* q) {
    KznResult r = kzn_queue_enter_transfer(q);
    if (r) return r;
    return q->n_pending ? kzn_flush_transfers(q) : KZN_SUCCESS;
}

KznResult kzn_queue_enter_draw(KznQueue* q) {
    KZN_VALID(q && q->device, "null queue");
    KZN_CUDA(cudaSetDevice(q->device->ordinal));
    return q->n_pending ? kzn_flush_transfers(q) : KZN_SUCCESS;
}

const void* kzn_recorded_host_copy(KznQueue* q, const void* data, size_t size) {
    if (!q->recording) return data;
    void* copy = malloc(size ? size : 1);
    if (!copy) return nullptr;
    memcpy(copy, data, size);
    q->recording->host_copies.push_back(copy);
    return copy;
}

bool kzn_pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("KZN_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

void kzn_timing_begin(KznQueue* q) {
    if (!q->timing || q->recording) return;
    cudaEventRecord(q->time_begin, q->stream);
    q->timing_armed = true;
}

KznResult kzn_launched(KznQueue* q, const char* kernel) {
    q->launches++;
    if (q->timing_armed) {
        cudaEventRecord(q->time_end, q->stream);
        q->timing_armed = false;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        kzn_set_error("launch of %s failed: %s", kernel, cudaGetErrorString(e));
        return KZN_ERROR_DEVICE_LOST;
    }
    return KZN_SUCCESS;
}

extern "C" {

static KznResult make_queue(KznDevice* device, cudaStream_t stream, bool owns, KznQueue** queue) {
    KznQueue* q = new (std::nothrow) KznQueue();
    KZN_REQUIRE(q, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    q->device = device;
    q->stream = stream;
    q->owns_stream = owns;
    q->launches = 0;
    q->stats_address = 0;
    q->raster = nullptr;
    q->timing = q->timing_armed = false;
    q->time_begin = q->time_end = nullptr;
    q->recording = nullptr;
    q->n_pending = 0;
    q->merged_commands = 0;
    *queue = q;
    return KZN_SUCCESS;
}

KznResult kzn_queue_create(KznDevice* device, KznQueue** queue) {
    KZN_VALID(device && queue, "kzn_queue_create: null argument");
    *queue = nullptr;
    KZN_CUDA(cudaSetDevice(device->ordinal));
    cudaStream_t s;
    KZN_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return make_queue(device, s, true, queue);
}

KznResult kzn_queue_wrap_stream(KznDevice* device, void* cuda_stream, KznQueue** queue) {
    KZN_VALID(device && queue, "kzn_queue_wrap_stream: null argument");
    *queue = nullptr;
    return make_queue(device, (cudaStream_t)cuda_stream, false, queue);
}


void kzn_queue_destroy(KznQueue* queue) {
    if (!queue) return;
    cudaSetDevice(queue->device->ordinal);
    cudaStreamSynchronize(queue->stream);
    kzn_raster_scratch_destroy(queue->raster);
    if (queue->time_begin) cudaEventDestroy(queue->time_begin);
    if (queue->time_end) cudaEventDestroy(queue->time_end);
    if (queue->owns_stream) cudaStreamDestroy(queue->stream);
    delete queue;
}

KznResult kzn_queue_wait_idle(KznQueue* queue) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(!queue->recording, "vkQueueWaitIdle while a command buffer is being recorded on this queue");
    KZN_CUDA(cudaStreamSynchronize(queue->stream));
    return KZN_SUCCESS;
}

void* kzn_queue_stream(const KznQueue* queue) { return queue ? (void*)queue->stream : nullptr; }
uint64_t kzn_queue_launch_count(const KznQueue* queue) { return queue ? queue->launches : 0; }

KznResult kzn_queue_set_kernel_timing(KznQueue* queue, int32_t enable) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    if (enable && !queue->time_begin) {
        KZN_CUDA(cudaEventCreate(&queue->time_begin));
        KZN_CUDA(cudaEventCreate(&queue->time_end));
    }
    queue->timing = enable != 0;
    queue->timing_armed = false;
    return KZN_SUCCESS;
}

KznResult kzn_queue_last_kernel_ms(KznQueue* queue, float* ms) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(ms, "null out pointer");
    KZN_VALID(queue->timing && queue->time_begin, "kernel timing is not enabled on this queue");
    cudaError_t e = cudaEventSynchronize(queue->time_end);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, queue->time_begin, queue->time_end);
    if (e != cudaSuccess) {
        cudaGetLastError();
        kzn_set_error("no timed kernel has been launched on this queue yet (%s)", cudaGetErrorString(e));
        return KZN_NOT_READY;
    }
    return KZN_SUCCESS;
}

KznResult kzn_queue_set_draw_statistics(KznQueue* queue, uint64_t stats_address) {
    KZN_VALID(queue, "null queue");
    KZN_VALID(stats_address % 8 == 0, "statistics address must be 8-byte aligned");
    queue->stats_address = stats_address;
    return KZN_SUCCESS;
}

KznResult kzn_queue_upload(KznQueue* queue, uint64_t dst_address, const void* src, uint64_t size) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(dst_address && src, "kzn_queue_upload: null address");
    if (size == 0) return KZN_SUCCESS;
    KZN_CUDA(cudaMemcpyAsync((void*)(uintptr_t)dst_address, src, size, cudaMemcpyHostToDevice, queue->stream));
    return KZN_SUCCESS;
}

KznResult kzn_queue_download(KznQueue* queue, void* dst, uint64_t src_address, uint64_t size) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(dst && src_address, "kzn_queue_download: null address");
    if (size == 0) return KZN_SUCCESS;
    KZN_CUDA(cudaMemcpyAsync(dst, (const void*)(uintptr_t)src_address, size, cudaMemcpyDeviceToHost, queue->stream));
    return KZN_SUCCESS;
}

KznResult kzn_host_alloc_pinned(uint64_t size, void** ptr) {
    KZN_VALID(ptr && size, "kzn_host_alloc_pinned: bad argument");
    KZN_CUDA(cudaHostAlloc(ptr, size, cudaHostAllocPortable));
    return KZN_SUCCESS;
}

void kzn_host_free_pinned(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

// ---- fences ---------------------------------------------------------------------------------------

KznResult kzn_fence_create(KznDevice* device, int32_t signaled, KznFence** fence) {
    KZN_VALID(device && fence, "kzn_fence_create: null argument");
    *fence = nullptr;
    KZN_CUDA(cudaSetDevice(device->ordinal));
    KznFence* f = new (std::nothrow) KznFence();
    KZN_REQUIRE(f, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    f->device = device;
    f->signaled_without_work = signaled != 0;
    f->pending = false;
    cudaError_t e = cudaEventCreateWithFlags(&f->event, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete f;
        return kzn_cuda_failed(e, "cudaEventCreateWithFlags", __FILE__, __LINE__);
    }
    *fence = f;
    return KZN_SUCCESS;
}

void kzn_fence_destroy(KznFence* fence) {
    if (!fence) return;
    cudaSetDevice(fence->device->ordinal);
    cudaEventDestroy(fence->event);
    delete fence;
}

KznResult kzn_fence_reset(KznFence* fence) {
    KZN_VALID(fence, "null fence");
    fence->signaled_without_work = false;
    fence->pending = false;
    return KZN_SUCCESS;
}

KznResult kzn_queue_signal_fence(KznQueue* queue, KznFence* fence) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(fence, "null fence");
    KZN_VALID(!queue->recording, "a fence is signalled by vkQueueSubmit, not from inside a command buffer");
    KZN_VALID(!fence->pending && !fence->signaled_without_work, "fence must be unsignaled when submitted");
    KZN_CUDA(cudaEventRecord(fence->event, queue->stream));
    fence->pending = true;
    return KZN_SUCCESS;
}

KznResult kzn_fence_status(KznFence* fence) {
    KZN_VALID(fence, "null fence");
    if (fence->signaled_without_work) return KZN_SUCCESS;
    if (!fence->pending) return KZN_NOT_READY;
    cudaError_t e = cudaEventQuery(fence->event);
    if (e == cudaSuccess) return KZN_SUCCESS;
    if (e == cudaErrorNotReady) return KZN_NOT_READY;
    return kzn_cuda_failed(e, "cudaEventQuery", __FILE__, __LINE__);
}

KznResult kzn_fence_wait(KznFence* fence, uint64_t timeout_ns) {
    KZN_VALID(fence, "null fence");
    if (fence->signaled_without_work) return KZN_SUCCESS;
    if (!fence->pending) {
        // never submitted: it can only time out (Vulkan: the wait returns VK_TIMEOUT)
        if (timeout_ns == ~0ull) {
            kzn_set_error("waiting forever on a fence that was never submitted");
            return KZN_ERROR_VALIDATION_FAILED;
        }
        return KZN_TIMEOUT;
    }
    if (timeout_ns == ~0ull) {
        KZN_CUDA(cudaEventSynchronize(fence->event));
        return KZN_SUCCESS;
    }
    struct timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (;;) {
        cudaError_t e = cudaEventQuery(fence->event);
        if (e == cudaSuccess) return KZN_SUCCESS;
        if (e != cudaErrorNotReady) return kzn_cuda_failed(e, "cudaEventQuery", __FILE__, __LINE__);
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        uint64_t el = (uint64_t)(t1.tv_sec - t0.tv_sec) * 1000000000ull + (uint64_t)(t1.tv_nsec - t0.tv_nsec);
        if (el >= timeout_ns) return KZN_TIMEOUT;
    }
}

// ---- image layout ------------------------------------------------------------------------------------

KznResult kzn_image_memory_requirements(const KznImageDesc* image, uint64_t* size, uint64_t* alignment) {
    KZN_VALID(image, "null image");
    KZN_REQUIRE(kzn_texel_size(image->format) != 0, KZN_ERROR_FORMAT_NOT_SUPPORTED,
                "format %u: ImageProperties::computed_properties is unimplemented for it (image.rs:86)",
                image->format);
    KZN_REQUIRE(image->mip_levels == 1, KZN_ERROR_FEATURE_NOT_PRESENT, "mip_levels must be 1 (image.rs:65)");
    KZN_VALID(image->tiling <= KZN_TILING_TILED, "bad tiling %u", image->tiling);
    KZN_VALID(image->width && image->height && image->depth && image->array_layers, "zero image extent");
    KZN_VALID(kzn_image_size_fits(image), "image byte size overflows (checked_mul in image.rs:74-82 panics)");
    uint64_t bytes = kzn_slice_pitch(image) * image->depth * image->array_layers;
    if (size) *size = kzn_round_up(bytes, KZN_IMAGE_ALIGNMENT);
    if (alignment) *alignment = KZN_IMAGE_ALIGNMENT;
    return KZN_SUCCESS;
}

KznResult kzn_cmd_pipeline_barrier(KznQueue* queue) {
    // in-order stream: nothing to enqueue (see the header)
    return kzn_queue_enter(queue);
}

// ---- semaphores ------------------------------------------------------------------------------------------------

KznResult kzn_semaphore_create(KznDevice* device, KznSemaphore** semaphore) {
    KZN_VALID(device && semaphore, "kzn_semaphore_create: null argument");
    *semaphore = nullptr;
    KZN_CUDA(cudaSetDevice(device->ordinal));
    KznSemaphore* s = new (std::nothrow) KznSemaphore();
    KZN_REQUIRE(s, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    s->device = device;
    s->signaled = false;
    cudaError_t e = cudaEventCreateWithFlags(&s->event, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete s;
        return kzn_cuda_failed(e, "cudaEventCreateWithFlags", __FILE__, __LINE__);
    }
    *semaphore = s;
    return KZN_SUCCESS;
}

void kzn_semaphore_destroy(KznSemaphore* semaphore) {
    if (!semaphore) return;
    cudaSetDevice(semaphore->device->ordinal);
    cudaEventDestroy(semaphore->event);
    delete semaphore;
}

KznResult kzn_queue_signal_semaphore(KznQueue* queue, KznSemaphore* semaphore) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(semaphore, "null semaphore");
    KZN_VALID(!queue->recording, "semaphores belong to vkQueueSubmit, not to a command buffer");
    KZN_CUDA(cudaEventRecord(semaphore->event, queue->stream));
    semaphore->signaled = true;
    return KZN_SUCCESS;
}

KznResult kzn_queue_wait_semaphore(KznQueue* queue, KznSemaphore* semaphore) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(semaphore, "null semaphore");
    KZN_VALID(!queue->recording, "semaphores belong to vkQueueSubmit, not to a command buffer");
    KZN_VALID(semaphore->signaled, "semaphore wait without a pending signal (binary semaphore rules)");
    KZN_CUDA(cudaStreamWaitEvent(queue->stream, semaphore->event, 0));
    semaphore->signaled = false;
    return KZN_SUCCESS;
}

KznResult kzn_queue_copy_dma(KznQueue* queue, uint64_t dst_address, uint64_t src_address, uint64_t size) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(dst_address && src_address, "kzn_queue_copy_dma: null address");
    if (size == 0) return KZN_SUCCESS;
    KZN_CUDA(cudaMemcpyAsync((void*)(uintptr_t)dst_address, (const void*)(uintptr_t)src_address, size, cudaMemcpyDeviceToDevice, queue->stream));
    return KZN_SUCCESS;
}

// ---- command buffers: vkAllocateCommandBuffers :4464, vkBeginCommandBuffer :4483, vkEndCommandBuffer :4491, ------------
// ---- vkQueueSubmit :3414 (all unimplemented!() upstream) ---------------------------------------------------------------

KznResult kzn_command_buffer_begin(KznQueue* queue, KznCommandBuffer** command_buffer) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(command_buffer, "null out pointer");
    *command_buffer = nullptr;
    KZN_VALID(!queue->recording, "another command buffer is already being recorded on this queue");
    KznCommandBuffer* cb = new (std::nothrow) KznCommandBuffer();
    KZN_REQUIRE(cb, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    cudaError_t e = cudaStreamBeginCapture(queue->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        delete cb;
        return kzn_cuda_failed(e, "cudaStreamBeginCapture", __FILE__, __LINE__);
    }
    cb->queue = queue;
    cb->recording = true;
    cb->launches_at_begin = queue->launches;
    queue->recording = cb;
    *command_buffer = cb;
    return KZN_SUCCESS;
}

KznResult kzn_command_buffer_end(KznCommandBuffer* cb) {
    KZN_VALID(cb && cb->recording, "vkEndCommandBuffer: the command buffer is not in the recording state");
    KznQueue* queue = cb->queue;
    KznResult r = kzn_queue_enter(queue);  // (a render pass clear that no draw took over is recorded here at the latest)
    if (r) return r;
    queue->in_render_pass = false;  // vkEndCommandBuffer inside a render pass is invalid usage (VUID 00060): the pass ends with its command buffer
    queue->recording = nullptr;
    cb->recording = false;
    cb->launches = queue->launches - cb->launches_at_begin;
    queue->launches = cb->launches_at_begin;  // recording launches nothing
    cudaError_t e = cudaStreamEndCapture(queue->stream, &cb->graph);
    if (e != cudaSuccess || !cb->graph) {
        cudaGetLastError();
        kzn_set_error("recording failed: %s (a command that cannot be recorded was issued)", cudaGetErrorString(e));
        return KZN_ERROR_VALIDATION_FAILED;
    }
    KZN_CUDA(cudaGraphInstantiate(&cb->exec, cb->graph, 0));
    return KZN_SUCCESS;
}

KznResult kzn_queue_submit(KznQueue* queue, uint32_t count, KznCommandBuffer* const* command_buffers, KznFence* fence) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(!queue->recording, "vkQueueSubmit while a command buffer is being recorded on this queue");
    KZN_VALID(count == 0 || command_buffers, "null command buffer array");
    for (uint32_t i = 0; i < count; ++i) {
        KznCommandBuffer* cb = command_buffers[i];
        KZN_VALID(cb && cb->exec && !cb->recording, "command buffer %u is not in the executable state", i);
        KZN_VALID(cb->queue->device == queue->device, "command buffer %u was recorded for another device", i);
        // a recorded draw works in the recording queue's raster scratch and work counters: replayed on another queue it
        // would race with that queue's own draws
        KZN_VALID(!cb->raster_scratch || cb->queue == queue, "command buffer %u records a draw: it can only be submitted to the queue it was recorded on", i);
    }
    // the whole submission is validated before any of it is enqueued
    if (fence) KZN_VALID(!fence->pending && !fence->signaled_without_work, "fence must be unsignaled when submitted");
    for (uint32_t i = 0; i < count; ++i) {
        KZN_CUDA(cudaGraphLaunch(command_buffers[i]->exec, queue->stream));
        queue->launches += command_buffers[i]->launches;
    }
    if (fence) return kzn_queue_signal_fence(queue, fence);
    return KZN_SUCCESS;
}

uint64_t kzn_command_buffer_launch_count(const KznCommandBuffer* cb) { return cb ? cb->launches : 0; }

void kzn_command_buffer_destroy(KznCommandBuffer* cb) {
    if (!cb) return;
    cudaSetDevice(cb->queue->device->ordinal);
    if (cb->recording) {  // abandoned mid-recording
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(cb->queue->stream, &g);
        if (g) cudaGraphDestroy(g);
        cb->queue->recording = nullptr;
        cb->queue->n_pending = 0;
        cb->queue->launches = cb->launches_at_begin;
        cudaGetLastError();
    }
    if (cb->exec) cudaGraphExecDestroy(cb->exec);
    if (cb->graph) cudaGraphDestroy(cb->graph);
    if (cb->raster_scratch) {
        cudaStreamSynchronize(cb->queue->stream);  // a replay may still be running in the retired buffers
        kzn_raster_scratch_unpin(cb->raster_scratch);
    }
    for (void* p : cb->host_copies) free(p);
    delete cb;
}

}  // extern "C"

KznResult kzn_check_image(const KznImageDesc* im, const char* what) {
    KZN_VALID(im, "%s: null image", what);
    KZN_REQUIRE(kzn_texel_size(im->format) != 0, KZN_ERROR_FORMAT_NOT_SUPPORTED, "%s: unsupported format %u", what,
                im->format);
    KZN_REQUIRE(im->mip_levels == 1, KZN_ERROR_FEATURE_NOT_PRESENT, "%s: mip_levels must be 1 (image.rs:65)", what);
    KZN_VALID(im->tiling <= KZN_TILING_TILED, "%s: bad tiling %u", what, im->tiling);
    KZN_VALID(im->width && im->height && im->depth && im->array_layers, "%s: zero image extent", what);
    KZN_VALID(kzn_image_size_fits(im), "%s: image byte size overflows (checked_mul in image.rs:74-82 panics)", what);
    KZN_VALID(im->address != 0, "%s: image has no memory bound", what);
    KZN_VALID(im->address % 4 == 0, "%s: image address must be texel aligned", what);
    return KZN_SUCCESS;
}

uint32_t kzn_pack_rgba8_host(const float rgba[4]) {
    uint32_t w = 0;
    for (int i = 0; i < 4; ++i) {
        float f = rgba[i];
        float c = !(f == f) ? 0.0f : (f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f));
        w |= (uint32_t)nearbyintf(c * 255.0f) << (8 * i);  // E.1: RNE, single multiply
    }
    return w;
}
