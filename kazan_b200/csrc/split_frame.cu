// split_frame.cu — sort-first split of one frame over the GPUs of a box, behind the C ABI.
//
// The reference exposes exactly one physical device (vulkan-driver/src/api_impl.rs:5092-5119), so a multi-GPU frame is
// internal to the driver: one process opens the N devices (KAZAN_B200_GPUS in the Rust host), every GPU renders the rows of
// its band of the frame (KznFramebuffer::band_y0 / band_y1 clip setup, binning and raster to the band; geometry is
// replicated), and ONE exchange step moves the bands into the presenting GPU's image over NVLink peer memory.  Bands are
// disjoint sets of pixels: seams need no blending and depth is never exchanged.
//
// In one process the whole exchange is stream work: the band's copy queue waits for the band's render queue (a CUDA event),
// pushes the band into the presenter's image with the copy engines (cudaMemcpyPeerAsync) and the presenter's queue waits for
// the N - 1 pushes (events again) — no flag kernels, no host round trip, no collective.  (The multi-process form of the same
// protocol — kzn_memory_export / kzn_memory_import and the flag commands of peer.cu — is what bench.py's torchrun ranks use.)
#include <new>
#include <vector>

#include "kzn_internal.h"

struct KznSplitFrame {
    struct Rank {
        KznDevice* device = nullptr;
        KznQueue* queue = nullptr;       // renders the band
        KznQueue* copy_queue = nullptr;  // pushes it (band ranks only)
        KznMemory* color = nullptr;      // whole-frame image; only the band's rows are ever valid on a band rank
        KznMemory* depth = nullptr;
        cudaEvent_t rendered = nullptr, pushed = nullptr;
        uint32_t y0 = 0, y1 = 0;
    };
    std::vector<Rank> ranks;
    uint32_t width = 0, height = 0, presenter = 0;
    bool with_depth = false;
};

namespace {

// rows [y0, y1) of band `rank`: bands of near-equal height whose starts are multiples of the raster brick height (16), so no
// warp's pixel block is shared between two GPUs
void band_rows(uint32_t height, uint32_t world, uint32_t rank, uint32_t* y0, uint32_t* y1) {
    const uint64_t units = ((uint64_t)height + 15) / 16;
    uint64_t lo = units * rank / world * 16, hi = units * (rank + 1) / world * 16;
    *y0 = (uint32_t)(lo < height ? lo : height);
    *y1 = (uint32_t)(hi < height ? hi : height);
}

}  // namespace

extern "C" {

void kzn_split_frame_destroy(KznSplitFrame* sf) {
    if (!sf) return;
    for (auto& r : sf->ranks) {
        if (!r.device) continue;
        cudaSetDevice(kzn_device_ordinal(r.device));
        if (r.queue) kzn_queue_wait_idle(r.queue);
        if (r.copy_queue) kzn_queue_wait_idle(r.copy_queue);
    }
    for (auto& r : sf->ranks) {
        if (!r.device) continue;
        cudaSetDevice(kzn_device_ordinal(r.device));
        if (r.rendered) cudaEventDestroy(r.rendered);
        if (r.pushed) cudaEventDestroy(r.pushed);
        if (r.color) kzn_memory_free(r.color);
        if (r.depth) kzn_memory_free(r.depth);
        if (r.copy_queue) kzn_queue_destroy(r.copy_queue);
        if (r.queue) kzn_queue_destroy(r.queue);
        kzn_device_destroy(r.device);
    }
    delete sf;
}

KznResult kzn_split_frame_create(uint32_t device_count, const int32_t* device_ordinals, uint32_t width, uint32_t height, uint32_t with_depth,
                                 KznSplitFrame** split_frame) {
    KZN_VALID(split_frame && device_ordinals && device_count >= 1 && device_count <= 16 && width && height, "kzn_split_frame_create: bad argument");
    *split_frame = nullptr;
    KznSplitFrame* sf = new (std::nothrow) KznSplitFrame();
    KZN_REQUIRE(sf, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    sf->width = width;
    sf->height = height;
    sf->with_depth = with_depth != 0;
    sf->ranks.resize(device_count);
    const uint64_t bytes = (uint64_t)width * height * 4;
    KznResult r = KZN_SUCCESS;
    for (uint32_t i = 0; i < device_count && !r; ++i) {
        auto& k = sf->ranks[i];
        band_rows(height, device_count, i, &k.y0, &k.y1);
        if ((r = kzn_device_create(device_ordinals[i], &k.device))) break;
        if ((r = kzn_queue_create(k.device, &k.queue))) break;
        if ((r = kzn_memory_allocate(k.device, bytes, KZN_MEMORY_TYPE_DEVICE_LOCAL, &k.color))) break;
        if (sf->with_depth && (r = kzn_memory_allocate(k.device, bytes, KZN_MEMORY_TYPE_DEVICE_LOCAL, &k.depth))) break;
        if (i != sf->presenter && (r = kzn_queue_create(k.device, &k.copy_queue))) break;
        cudaSetDevice(device_ordinals[i]);
        if (cudaEventCreateWithFlags(&k.rendered, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&k.pushed, cudaEventDisableTiming) != cudaSuccess) {
            kzn_set_error("cudaEventCreate failed on device %d", device_ordinals[i]);
            r = KZN_ERROR_INITIALIZATION_FAILED;
            break;
        }
        if (i != sf->presenter) {  // the band goes straight into the presenter's memory: NVLink peer access
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device_ordinals[i], device_ordinals[sf->presenter]);
            if (!can) {
                kzn_set_error("device %d cannot access device %d's memory (no NVLink / PCIe peer path)", device_ordinals[i], device_ordinals[sf->presenter]);
                r = KZN_ERROR_FEATURE_NOT_PRESENT;
                break;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ordinals[sf->presenter], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                kzn_set_error("cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
                r = KZN_ERROR_INITIALIZATION_FAILED;
                break;
            }
            cudaGetLastError();
        }
    }
    if (r) {
        kzn_split_frame_destroy(sf);
        return r;
    }
    *split_frame = sf;
    return KZN_SUCCESS;
}

uint32_t kzn_split_frame_device_count(const KznSplitFrame* sf) { return sf ? (uint32_t)sf->ranks.size() : 0; }
KznDevice* kzn_split_frame_device(const KznSplitFrame* sf, uint32_t rank) { return sf && rank < sf->ranks.size() ? sf->ranks[rank].device : nullptr; }
KznQueue* kzn_split_frame_queue(const KznSplitFrame* sf, uint32_t rank) { return sf && rank < sf->ranks.size() ? sf->ranks[rank].queue : nullptr; }

KznResult kzn_split_frame_framebuffer(const KznSplitFrame* sf, uint32_t rank, KznFramebuffer* framebuffer) {
    KZN_VALID(sf && framebuffer && rank < sf->ranks.size(), "kzn_split_frame_framebuffer: bad argument");
    const auto& k = sf->ranks[rank];
    memset(framebuffer, 0, sizeof *framebuffer);
    KznImageDesc& c = framebuffer->color;
    c.address = kzn_memory_device_address(k.color);
    c.format = KZN_FORMAT_R8G8B8A8_UNORM;
    c.tiling = KZN_TILING_LINEAR;
    c.width = sf->width;
    c.height = sf->height;
    c.depth = c.array_layers = c.mip_levels = 1;
    if (sf->with_depth) {
        framebuffer->depth = c;
        framebuffer->depth.address = kzn_memory_device_address(k.depth);
        framebuffer->depth.format = KZN_FORMAT_D32_SFLOAT;
    }
    framebuffer->band_y0 = k.y0;
    framebuffer->band_y1 = k.y1;
    return KZN_SUCCESS;
}

// The one exchange step of a frame: every band rank's copy queue, ordered after what its render queue has been given so far,
// pushes the band's rows into the presenter's image; the presenter's queue is ordered after all of them.
KznResult kzn_split_frame_composite(KznSplitFrame* sf) {
    KZN_VALID(sf, "kzn_split_frame_composite: null argument");
    auto& p = sf->ranks[sf->presenter];
    const uint64_t pitch = (uint64_t)sf->width * 4;
    for (uint32_t i = 0; i < sf->ranks.size(); ++i) {
        if (i == sf->presenter) continue;
        auto& k = sf->ranks[i];
        if (k.y1 <= k.y0) continue;
        KZN_CUDA(cudaSetDevice(kzn_device_ordinal(k.device)));
        cudaStream_t render = (cudaStream_t)kzn_queue_stream(k.queue), copy = (cudaStream_t)kzn_queue_stream(k.copy_queue);
        KZN_CUDA(cudaEventRecord(k.rendered, render));
        KZN_CUDA(cudaStreamWaitEvent(copy, k.rendered, 0));
        KZN_CUDA(cudaMemcpyPeerAsync((void*)(uintptr_t)(kzn_memory_device_address(p.color) + k.y0 * pitch), kzn_device_ordinal(p.device),
                                     (const void*)(uintptr_t)(kzn_memory_device_address(k.color) + k.y0 * pitch), kzn_device_ordinal(k.device),
                                     (k.y1 - k.y0) * pitch, copy));
        KZN_CUDA(cudaEventRecord(k.pushed, copy));
        // the band's image may be rendered into again only after the push has read it
        KZN_CUDA(cudaStreamWaitEvent(render, k.pushed, 0));
    }
    KZN_CUDA(cudaSetDevice(kzn_device_ordinal(p.device)));
    for (uint32_t i = 0; i < sf->ranks.size(); ++i)
        if (i != sf->presenter && sf->ranks[i].y1 > sf->ranks[i].y0) KZN_CUDA(cudaStreamWaitEvent((cudaStream_t)kzn_queue_stream(p.queue), sf->ranks[i].pushed, 0));
    return KZN_SUCCESS;
}

uint64_t kzn_split_frame_present_address(const KznSplitFrame* sf) { return sf ? kzn_memory_device_address(sf->ranks[sf->presenter].color) : 0; }
KznQueue* kzn_split_frame_present_queue(const KznSplitFrame* sf) { return sf ? sf->ranks[sf->presenter].queue : nullptr; }

KznResult kzn_split_frame_wait_idle(KznSplitFrame* sf) {
    KZN_VALID(sf, "kzn_split_frame_wait_idle: null argument");
    for (auto& k : sf->ranks) {
        KznResult r = kzn_queue_wait_idle(k.queue);
        if (r) return r;
        if (k.copy_queue && (r = kzn_queue_wait_idle(k.copy_queue))) return r;
    }
    return KZN_SUCCESS;
}

}  // extern "C"
