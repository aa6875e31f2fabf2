// present.cu — the data path of WSI presentation.
//
// The reference sets a swapchain up over X11 MIT-SHM — window pixmaps of 32 bits per pixel in R8G8B8A8 or B8G8R8A8 byte
// order, scanlines padded to the pixmap format's scanline_pad, 2..16 images, SysV shared memory segments that the X server
// attaches (vulkan-driver/src/xcb_swapchain.rs:117-136,258-296, shm.rs:22-58) — and stops there: XcbSwapchain::new,
// vkGetSwapchainImagesKHR, vkAcquireNextImageKHR and vkQueuePresentKHR are `unimplemented!()` (xcb_swapchain.rs:342-349,
// api_impl.rs:5656-5682).  What those need from the GPU side is here: presentable images in HBM, one SysV segment per image
// (registered with CUDA, so the copy engines or the SMs write it directly), and a present that — ordered after the queue's
// rendering — puts the image into its segment in the window's byte order and scanline pitch.  The X11 calls themselves
// (xcb_shm_attach once per segment, xcb_shm_put_image per present) stay in the host: INTEGRATION.md.
#include <sys/ipc.h>
#include <sys/shm.h>

#include <new>

#include "kzn_internal.h"

struct KznSwapchain {
    KznDevice* device = nullptr;
    KznSwapchainCreateInfo info = {};
    uint32_t bytes_per_line = 0;
    uint64_t segment_size = 0, image_size = 0;
    uint32_t next = 0;  // acquire hands the images out in order
    struct Image {
        void* device_image = nullptr;   // HBM, LINEAR R8G8B8A8_UNORM: the render target
        int shm_id = -1;
        void* host = nullptr;           // the attached segment
        void* host_as_device = nullptr;  // the same memory as the GPU addresses it
        cudaEvent_t presented = nullptr;
        bool registered = false, acquired = false, in_flight = false;
    } images[KZN_MAX_SWAPCHAIN_IMAGES];
};

namespace {

// window byte order B8G8R8A8 (the common X11 visual): swap bytes 0 and 2 of every texel on the way out; 4 texels per
// thread, 16-byte loads from HBM and 16-byte stores into the mapped segment (posted PCIe writes)
__global__ void __launch_bounds__(256) present_swizzle_kernel(const uint4* __restrict__ src, unsigned char* __restrict__ dst, uint32_t quads_per_row,
                                                              uint32_t rows, uint32_t bytes_per_line) {
    kzn_pdl_wait();
    const uint64_t n = (uint64_t)quads_per_row * rows;
    for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256u) {
        const uint32_t y = (uint32_t)(i / quads_per_row), q = (uint32_t)(i - (uint64_t)y * quads_per_row);
        uint4 v = __ldg(src + i);
        v.x = __byte_perm(v.x, 0, 0x3012);
        v.y = __byte_perm(v.y, 0, 0x3012);
        v.z = __byte_perm(v.z, 0, 0x3012);
        v.w = __byte_perm(v.w, 0, 0x3012);
        *reinterpret_cast<uint4*>(dst + (uint64_t)y * bytes_per_line + 16ull * q) = v;
    }
}
// widths that are not a multiple of 4 texels: one texel per thread
__global__ void __launch_bounds__(256) present_swizzle_texel_kernel(const uint32_t* __restrict__ src, unsigned char* __restrict__ dst, uint32_t width,
                                                                    uint32_t rows, uint32_t bytes_per_line) {
    kzn_pdl_wait();
    const uint64_t n = (uint64_t)width * rows;
    for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256u) {
        const uint32_t y = (uint32_t)(i / width), x = (uint32_t)(i - (uint64_t)y * width);
        *reinterpret_cast<uint32_t*>(dst + (uint64_t)y * bytes_per_line + 4ull * x) = __byte_perm(__ldg(src + i), 0, 0x3012);
    }
}

void release(KznSwapchain* sc) {
    for (auto& im : sc->images) {
        if (im.presented) cudaEventDestroy(im.presented);
        if (im.registered) cudaHostUnregister(im.host);
        if (im.host) shmdt(im.host);
        if (im.shm_id != -1) shmctl(im.shm_id, IPC_RMID, nullptr);  // (shm.rs:60-66: the segment goes with its owner)
        if (im.device_image) cudaFree(im.device_image);
    }
    delete sc;
}

}  // namespace

extern "C" {

KznResult kzn_swapchain_create(KznDevice* device, const KznSwapchainCreateInfo* info, KznSwapchain** swapchain) {
    KZN_VALID(device && info && swapchain, "vkCreateSwapchainKHR: null argument");
    KZN_VALID(info->width && info->height && info->width <= 32768u && info->height <= 32768u, "vkCreateSwapchainKHR: bad image extent");
    KZN_VALID(info->image_count >= 2 && info->image_count <= KZN_MAX_SWAPCHAIN_IMAGES,
              "vkCreateSwapchainKHR: minImageCount is 2, maxImageCount %u (xcb_swapchain.rs:136,307-308)", KZN_MAX_SWAPCHAIN_IMAGES);
    KZN_REQUIRE(info->window_byte_order == KZN_WINDOW_R8G8B8A8 || info->window_byte_order == KZN_WINDOW_B8G8R8A8, KZN_ERROR_FORMAT_NOT_SUPPORTED,
                "vkCreateSwapchainKHR: the window's pixmap format is neither R8G8B8A8 nor B8G8R8A8 (xcb_swapchain.rs:262-289)");
    KZN_VALID(info->scanline_alignment == 1 || info->scanline_alignment == 2 || info->scanline_alignment == 4,
              "vkCreateSwapchainKHR: scanline_pad is 8, 16 or 32 bits (xcb_swapchain.rs:290-295)");
    KZN_CUDA(cudaSetDevice(device->ordinal));
    KznSwapchain* sc = new (std::nothrow) KznSwapchain();
    KZN_REQUIRE(sc, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    sc->device = device;
    sc->info = *info;
    sc->bytes_per_line = (uint32_t)kzn_round_up((uint64_t)info->width * 4u, info->scanline_alignment);  // 4-byte texels: already aligned
    sc->segment_size = (uint64_t)sc->bytes_per_line * info->height;
    sc->image_size = kzn_round_up((uint64_t)info->width * info->height * 4u, KZN_IMAGE_ALIGNMENT);
    for (uint32_t i = 0; i < info->image_count; ++i) {
        auto& im = sc->images[i];
        cudaError_t e = cudaMalloc(&im.device_image, sc->image_size);
        if (e != cudaSuccess) {
            release(sc);
            return kzn_cuda_failed(e, "cudaMalloc(swapchain image)", __FILE__, __LINE__);
        }
        im.shm_id = shmget(IPC_PRIVATE, sc->segment_size, IPC_CREAT | IPC_EXCL | 0666);  // shm.rs:35-44
        void* at = im.shm_id == -1 ? (void*)-1 : shmat(im.shm_id, nullptr, 0);
        if (at == (void*)-1) {
            release(sc);
            kzn_set_error("vkCreateSwapchainKHR: no SysV shared memory segment of %llu bytes (shmget / shmat failed)", (unsigned long long)sc->segment_size);
            return KZN_ERROR_OUT_OF_HOST_MEMORY;
        }
        im.host = at;
        e = cudaHostRegister(im.host, sc->segment_size, cudaHostRegisterMapped | cudaHostRegisterPortable);
        if (e == cudaSuccess) {
            im.registered = true;
            e = cudaHostGetDevicePointer(&im.host_as_device, im.host, 0);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&im.presented, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            release(sc);
            return kzn_cuda_failed(e, "swapchain segment registration", __FILE__, __LINE__);
        }
    }
    *swapchain = sc;
    return KZN_SUCCESS;
}

void kzn_swapchain_destroy(KznSwapchain* swapchain) {
    if (!swapchain) return;
    cudaSetDevice(swapchain->device->ordinal);
    for (uint32_t i = 0; i < swapchain->info.image_count; ++i)
        if (swapchain->images[i].in_flight) cudaEventSynchronize(swapchain->images[i].presented);
    release(swapchain);
}

KznResult kzn_swapchain_get_image(const KznSwapchain* swapchain, uint32_t index, KznImageDesc* image) {
    KZN_VALID(swapchain && image && index < swapchain->info.image_count, "vkGetSwapchainImagesKHR: bad argument");
    *image = KznImageDesc{(uint64_t)(uintptr_t)swapchain->images[index].device_image, KZN_FORMAT_R8G8B8A8_UNORM, KZN_TILING_LINEAR,
                          swapchain->info.width, swapchain->info.height, 1, 1, 1, 0};
    return KZN_SUCCESS;
}

KznResult kzn_swapchain_get_segment(const KznSwapchain* swapchain, uint32_t index, KznSwapchainSegment* segment) {
    KZN_VALID(swapchain && segment && index < swapchain->info.image_count, "kzn_swapchain_get_segment: bad argument");
    segment->shm_id = swapchain->images[index].shm_id;
    segment->bytes_per_line = swapchain->bytes_per_line;
    segment->size = swapchain->segment_size;
    segment->mapped = swapchain->images[index].host;
    return KZN_SUCCESS;
}

KznResult kzn_swapchain_acquire_next_image(KznSwapchain* swapchain, uint64_t timeout_ns, uint32_t* index) {
    KZN_VALID(swapchain && index, "vkAcquireNextImageKHR: null argument");
    const uint32_t i = swapchain->next;
    auto& im = swapchain->images[i];
    if (im.acquired) {  // every image is out with the application
        kzn_set_error("vkAcquireNextImageKHR: all %u images are acquired and not presented", swapchain->info.image_count);
        return timeout_ns ? KZN_TIMEOUT : KZN_NOT_READY;
    }
    if (im.in_flight) {  // its last present must have landed in the segment before it is rendered to again
        KZN_CUDA(cudaSetDevice(swapchain->device->ordinal));
        cudaError_t e = cudaEventQuery(im.presented);
        if (e == cudaErrorNotReady) {
            if (timeout_ns == 0) return KZN_NOT_READY;
            e = cudaEventSynchronize(im.presented);  // (bounded by the copy of one frame; finer timeouts are the host's business)
        }
        if (e != cudaSuccess) return kzn_cuda_failed(e, "cudaEventSynchronize(presented)", __FILE__, __LINE__);
        im.in_flight = false;
    }
    im.acquired = true;
    swapchain->next = (i + 1) % swapchain->info.image_count;
    *index = i;
    return KZN_SUCCESS;
}

KznResult kzn_queue_present(KznQueue* queue, KznSwapchain* swapchain, uint32_t index) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(swapchain && index < swapchain->info.image_count, "vkQueuePresentKHR: bad image index");
    KZN_VALID(!queue->recording, "vkQueuePresentKHR is a queue operation: it cannot be recorded into a command buffer");
    KZN_VALID(queue->device == swapchain->device, "vkQueuePresentKHR: the swapchain belongs to another device");
    auto& im = swapchain->images[index];
    KZN_VALID(im.acquired, "vkQueuePresentKHR: image %u was not acquired (VUID-VkPresentInfoKHR-pImageIndices-01296)", index);
    const uint32_t w = swapchain->info.width, h = swapchain->info.height;
    if (swapchain->info.window_byte_order == KZN_WINDOW_R8G8B8A8) {
        // same byte order: the copy engine moves the rows (pitch = the window's scanline pitch)
        KZN_CUDA(cudaMemcpy2DAsync(im.host, swapchain->bytes_per_line, im.device_image, (size_t)w * 4u, (size_t)w * 4u, h, cudaMemcpyDeviceToHost, queue->stream));
    } else if (w % 4u == 0 && swapchain->bytes_per_line % 16u == 0) {
        const uint64_t quads = (uint64_t)(w / 4u) * h;
        const uint64_t blocks = (quads + 255) / 256, cap = (uint64_t)queue->device->sm_count * 8;
        kzn_launch_pdl(present_swizzle_kernel, dim3((uint32_t)(blocks < cap ? blocks : cap)), dim3(256), 0, queue->stream,
                       static_cast<const uint4*>(im.device_image), static_cast<unsigned char*>(im.host_as_device), w / 4u, h, swapchain->bytes_per_line);
        if ((r = kzn_launched(queue, "present_swizzle_kernel"))) return r;
    } else {
        const uint64_t texels = (uint64_t)w * h;
        const uint64_t blocks = (texels + 255) / 256, cap = (uint64_t)queue->device->sm_count * 8;
        kzn_launch_pdl(present_swizzle_texel_kernel, dim3((uint32_t)(blocks < cap ? blocks : cap)), dim3(256), 0, queue->stream,
                       static_cast<const uint32_t*>(im.device_image), static_cast<unsigned char*>(im.host_as_device), w, h, swapchain->bytes_per_line);
        if ((r = kzn_launched(queue, "present_swizzle_texel_kernel"))) return r;
    }
    KZN_CUDA(cudaEventRecord(im.presented, queue->stream));
    im.acquired = false;
    im.in_flight = true;
    return KZN_SUCCESS;
}

KznResult kzn_swapchain_wait_presented(KznSwapchain* swapchain, uint32_t index) {
    KZN_VALID(swapchain && index < swapchain->info.image_count, "kzn_swapchain_wait_presented: bad image index");
    auto& im = swapchain->images[index];
    if (!im.in_flight) return KZN_SUCCESS;
    KZN_CUDA(cudaSetDevice(swapchain->device->ordinal));
    KZN_CUDA(cudaEventSynchronize(im.presented));
    return KZN_SUCCESS;
}

}  // extern "C"
