/* kazan_b200.h — C ABI of the B200-native execution back end for the Kazan Vulkan driver.
 *
 * This is the drop-in boundary for the data-parallel hot path: the entry points a
 * `vulkan-driver` host (Rust `extern "C"` block) binds in place of the
 * `unimplemented!()` bodies of its `vkCmd*` / `vkQueue*` functions.  Every function cites the
 * reference interface it replaces (paths relative to the kazan-3d/kazan tree).
 *
 * Conventions
 *  - plain C, plain pointers and sizes; no CUDA or torch types in any signature
 *    (a CUstream / cudaStream_t crosses the boundary as `void*`);
 *  - structs whose name matches a Vulkan struct (KznBufferImageCopy <-> VkBufferImageCopy, ...)
 *    are layout-identical to Vulkan-Headers 1.1.130 (the version the reference vendors,
 *    external/Vulkan-Headers/include/vulkan/vulkan_core.h:47); tests/test_abi.py checks sizeof /
 *    offsetof against the real header when it is available, so a host may pass its Vk structs as-is;
 *  - enum values are the Vulkan values (VkFormat, VkCompareOp, VkBlendFactor, VkFilter, ...);
 *  - return value: KznResult, numerically a VkResult (0 = VK_SUCCESS, 1 = VK_NOT_READY,
 *    2 = VK_TIMEOUT, negative = error) so the host can return it unchanged.  Invalid usage that the
 *    reference would `assert!`/panic on (vulkan-driver/src/api_impl.rs, e.g. :4992-5003) is reported
 *    as KZN_ERROR_VALIDATION_FAILED; the ICD layer turns that into abort() like the reference;
 *  - all kzn_cmd_* calls are asynchronous: they enqueue sm_100a kernels on the queue's CUDA stream
 *    and return; completion is observed with a fence or kzn_queue_wait_idle;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    KZN_ERROR_INITIALIZATION_FAILED.
 */
#ifndef KAZAN_B200_H
#define KAZAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KZN_API __attribute__((visibility("default")))
#else
#define KZN_API
#endif

/* ------------------------------------------------------------------------------------------ */
/* results                                                                                      */

typedef int32_t KznResult;
#define KZN_SUCCESS 0
#define KZN_NOT_READY 1                            /* VK_NOT_READY */
#define KZN_TIMEOUT 2                              /* VK_TIMEOUT */
#define KZN_ERROR_OUT_OF_HOST_MEMORY (-1)          /* VK_ERROR_OUT_OF_HOST_MEMORY */
#define KZN_ERROR_OUT_OF_DEVICE_MEMORY (-2)        /* VK_ERROR_OUT_OF_DEVICE_MEMORY, api_impl.rs:3459-3471 */
#define KZN_ERROR_INITIALIZATION_FAILED (-3)       /* VK_ERROR_INITIALIZATION_FAILED: no CUDA device */
#define KZN_ERROR_DEVICE_LOST (-4)                 /* VK_ERROR_DEVICE_LOST: any CUresult/cudaError_t failure */
#define KZN_ERROR_FEATURE_NOT_PRESENT (-8)         /* VK_ERROR_FEATURE_NOT_PRESENT */
#define KZN_ERROR_FORMAT_NOT_SUPPORTED (-11)       /* VK_ERROR_FORMAT_NOT_SUPPORTED */
#define KZN_ERROR_VALIDATION_FAILED (-1000011001)  /* VK_ERROR_VALIDATION_FAILED_EXT */

/* library version, bumped on ABI changes */
#define KZN_ABI_VERSION 1u
KZN_API uint32_t kzn_abi_version(void);
/* human-readable text for the last error on the calling thread ("" if none) */
KZN_API const char* kzn_last_error(void);

/* ------------------------------------------------------------------------------------------ */
/* constants of the reference driver (vulkan-driver/src/lib.rs:27-35)                           */

#define KZN_MIN_MEMORY_MAP_ALIGNMENT 128u /* lib.rs:27 */
#define KZN_BUFFER_ALIGNMENT 64u          /* lib.rs:33 */
#define KZN_IMAGE_ALIGNMENT 64u           /* lib.rs:34 */
#define KZN_SUBPIXEL_PRECISION_BITS 4u    /* api_impl.rs:1559 */
#define KZN_WHOLE_SIZE (~0ull)            /* VK_WHOLE_SIZE */
#define KZN_REMAINING_MIP_LEVELS (~0u)    /* VK_REMAINING_MIP_LEVELS */
#define KZN_REMAINING_ARRAY_LAYERS (~0u)  /* VK_REMAINING_ARRAY_LAYERS */

/* VkFormat values handled on the hot path (image.rs:58-88 implements only the first) */
#define KZN_FORMAT_UNDEFINED 0u
#define KZN_FORMAT_R8G8B8A8_UNORM 37u
#define KZN_FORMAT_D32_SFLOAT 126u

/* image.rs:17-22 `enum Tiling`.  TILED is the oracle-defined 16x16-texel tile layout
 * (SURVEY.md Appendix E.10): the reference declares the variant but defines no address function. */
#define KZN_TILING_LINEAR 0u
#define KZN_TILING_TILED 1u
#define KZN_TILE_DIM 16u /* texels per tile edge; one tile = 16*16*4 B = 1 KiB */

#define KZN_IMAGE_ASPECT_COLOR_BIT 1u
#define KZN_IMAGE_ASPECT_DEPTH_BIT 2u

#define KZN_FILTER_NEAREST 0u
#define KZN_FILTER_LINEAR 1u

#define KZN_INDEX_TYPE_UINT16 0u
#define KZN_INDEX_TYPE_UINT32 1u

/* ------------------------------------------------------------------------------------------ */
/* Vulkan-layout structs (vulkan_core.h line numbers of the reference's vendored header)        */

typedef struct KznOffset3D { int32_t x, y, z; } KznOffset3D;                 /* VkOffset3D */
typedef struct KznExtent3D { uint32_t width, height, depth; } KznExtent3D;   /* VkExtent3D */
typedef struct KznOffset2D { int32_t x, y; } KznOffset2D;                    /* VkOffset2D */
typedef struct KznExtent2D { uint32_t width, height; } KznExtent2D;          /* VkExtent2D */
typedef struct KznRect2D { KznOffset2D offset; KznExtent2D extent; } KznRect2D; /* VkRect2D */

typedef struct KznViewport { /* VkViewport */
    float x, y, width, height, minDepth, maxDepth;
} KznViewport;

typedef union KznClearColorValue { /* VkClearColorValue, vulkan_core.h:2915-2919 */
    float float32[4];
    int32_t int32[4];
    uint32_t uint32[4];
} KznClearColorValue;

typedef struct KznClearDepthStencilValue { /* VkClearDepthStencilValue */
    float depth;
    uint32_t stencil;
} KznClearDepthStencilValue;

typedef struct KznImageSubresourceRange { /* VkImageSubresourceRange */
    uint32_t aspectMask, baseMipLevel, levelCount, baseArrayLayer, layerCount;
} KznImageSubresourceRange;

typedef struct KznImageSubresourceLayers { /* VkImageSubresourceLayers */
    uint32_t aspectMask, mipLevel, baseArrayLayer, layerCount;
} KznImageSubresourceLayers;

typedef struct KznBufferCopy { /* VkBufferCopy */
    uint64_t srcOffset, dstOffset, size;
} KznBufferCopy;

typedef struct KznBufferImageCopy { /* VkBufferImageCopy, vulkan_core.h:2906-2913 (56 bytes) */
    uint64_t bufferOffset;
    uint32_t bufferRowLength;
    uint32_t bufferImageHeight;
    KznImageSubresourceLayers imageSubresource;
    KznOffset3D imageOffset;
    KznExtent3D imageExtent;
} KznBufferImageCopy;

typedef struct KznImageCopy { /* VkImageCopy */
    KznImageSubresourceLayers srcSubresource;
    KznOffset3D srcOffset;
    KznImageSubresourceLayers dstSubresource;
    KznOffset3D dstOffset;
    KznExtent3D extent;
} KznImageCopy;

typedef struct KznImageBlit { /* VkImageBlit, vulkan_core.h:2899-2904 */
    KznImageSubresourceLayers srcSubresource;
    KznOffset3D srcOffsets[2];
    KznImageSubresourceLayers dstSubresource;
    KznOffset3D dstOffsets[2];
} KznImageBlit;

typedef struct KznClearRect { /* VkClearRect */
    KznRect2D rect;
    uint32_t baseArrayLayer, layerCount;
} KznClearRect;

/* ------------------------------------------------------------------------------------------ */
/* objects                                                                                      */

typedef struct KznDevice KznDevice; /* one CUDA device + primary context  (api_impl.rs `Device`) */
typedef struct KznQueue KznQueue;   /* one CUDA stream (lib.rs:30-32: 1 family x 1 queue)        */
typedef struct KznMemory KznMemory; /* device_memory.rs:191-273 `DeviceMemory`                   */
typedef struct KznFence KznFence;   /* handle.rs:327 `Fence {}` placeholder -> CUDA event        */

/* Image as the hot path sees it: vulkan-driver/src/image.rs:30-39 `ImageProperties` plus the bound
 * memory (`ImageMemory`, image.rs:92-95) resolved to a device address.
 * Texel (x,y,z,layer) of a LINEAR image lives at
 *     address + texel_size * (x + width * (y + height * (z + depth * layer)))        (image.rs:58-88)
 * A TILED image stores each (z,layer) slice as ceil(w/16) x ceil(h/16) row-major tiles of 16x16
 * row-major texels; slice pitch = texel_size*256*tiles_x*tiles_y. */
typedef struct KznImageDesc {
    uint64_t address; /* device address of texel 0; multiple of KZN_IMAGE_ALIGNMENT */
    uint32_t format;  /* KZN_FORMAT_* */
    uint32_t tiling;  /* KZN_TILING_* */
    uint32_t width, height, depth;
    uint32_t array_layers;
    uint32_t mip_levels; /* must be 1, as in image.rs:65 */
    uint32_t reserved;
} KznImageDesc;

/* size/alignment exactly as ImageProperties::computed_properties (image.rs:58-88) for LINEAR;
 * whole-tile padded for TILED.  Returns KZN_ERROR_FORMAT_NOT_SUPPORTED where the reference hits
 * `unimplemented!` (D32_SFLOAT is additionally accepted). */
KZN_API KznResult kzn_image_memory_requirements(const KznImageDesc* image, uint64_t* size,
                                                uint64_t* alignment);

/* -- device ---------------------------------------------------------------------------------- */
KZN_API KznResult kzn_device_count(int32_t* count);
KZN_API KznResult kzn_device_create(int32_t ordinal, KznDevice** device);
KZN_API void kzn_device_destroy(KznDevice* device);
KZN_API KznResult kzn_device_wait_idle(KznDevice* device); /* vkDeviceWaitIdle api_impl.rs:3429 */
KZN_API int32_t kzn_device_sm_count(const KznDevice* device);
KZN_API int32_t kzn_device_ordinal(const KznDevice* device);

/* -- memory: vkAllocateMemory api_impl.rs:3434-3474, vkMapMemory :3488-3500 ------------------- */
#define KZN_MEMORY_TYPE_HOST_COHERENT 0u /* reference's only type (device_memory.rs:30-39): pinned, mapped */
#define KZN_MEMORY_TYPE_DEVICE_LOCAL 1u  /* added after index 0: HBM only, not mappable */
KZN_API KznResult kzn_memory_allocate(KznDevice* device, uint64_t size, uint32_t memory_type,
                                      KznMemory** memory);
KZN_API void kzn_memory_free(KznMemory* memory);
KZN_API uint64_t kzn_memory_device_address(const KznMemory* memory);
KZN_API uint64_t kzn_memory_size(const KznMemory* memory); /* rounded up to 128, device_memory.rs:166-181 */
/* host pointer = base + offset (api_impl.rs:3488-3500); KZN_ERROR_VALIDATION_FAILED for DEVICE_LOCAL */
KZN_API KznResult kzn_memory_map(KznMemory* memory, uint64_t offset, void** host_ptr);

/* Split-frame rendering (the reference has one physical device, api_impl.rs:5092-5119): a DEVICE_LOCAL
 * allocation can be exported to the other per-GPU processes of the box and imported there as peer memory (NVLink).
 * A rank that binds an imported image as its colour attachment renders its band straight into the presenting GPU. */
#define KZN_MEMORY_HANDLE_BYTES 64u
KZN_API KznResult kzn_memory_export(KznMemory* memory, uint8_t handle[KZN_MEMORY_HANDLE_BYTES]);
KZN_API KznResult kzn_memory_import(KznDevice* device, const uint8_t handle[KZN_MEMORY_HANDLE_BYTES], uint64_t size,
                                    KznMemory** memory); /* release with kzn_memory_free */

/* -- queue: vkGetDeviceQueue :3394, vkQueueSubmit :3414, vkQueueWaitIdle :3424 ----------------
 * A queue is one in-order CUDA stream: a command sees the results of every earlier command of the queue.  The transfer
 * and draw kernels are launched with programmatic dependent launch (each begins with griddepcontrol.wait, which returns
 * once the previous kernel of the stream has completed and its writes are visible), so that order is kept while the launch
 * gap between consecutive kernels is hidden; KZN_PDL=0 in the environment restores ordinary launches. */
KZN_API KznResult kzn_queue_create(KznDevice* device, KznQueue** queue);
/* run on a stream owned by the host (e.g. the framework's current stream); not destroyed with the queue */
KZN_API KznResult kzn_queue_wrap_stream(KznDevice* device, void* cuda_stream, KznQueue** queue);
KZN_API void kzn_queue_destroy(KznQueue* queue);
KZN_API KznResult kzn_queue_wait_idle(KznQueue* queue);
KZN_API void* kzn_queue_stream(const KznQueue* queue);
/* number of kernels this library has launched on the queue since creation (bench `gpu_launches`) */
KZN_API uint64_t kzn_queue_launch_count(const KznQueue* queue);

/* Device-side duration of the kernel that dominates a command (raster kernel of a draw; the fill / copy / box / blit
 * kernel of a transfer), measured with a CUDA event pair on the queue's stream — what the benchmark's roofline uses.
 * kzn_queue_last_kernel_ms waits for the last timed kernel; KZN_NOT_READY if none was launched since enabling. */
KZN_API KznResult kzn_queue_set_kernel_timing(KznQueue* queue, int32_t enable);
KZN_API KznResult kzn_queue_last_kernel_ms(KznQueue* queue, float* ms);

/* Stream-ordered flags in (peer) device memory, the composite step of a split frame: kzn_cmd_signal_flag makes all prior
 * work of the queue visible system-wide and then adds `add` to the 64-bit flag; kzn_cmd_wait_flag holds the queue until
 * the flag is >= expected.  A wait that sees no progress for 5 s gives up and writes 1 to the u32 at status_address
 * (0 = none) instead of hanging the GPU. */
KZN_API KznResult kzn_cmd_signal_flag(KznQueue* queue, uint64_t flag_address, uint64_t add);
KZN_API KznResult kzn_cmd_wait_flag(KznQueue* queue, uint64_t flag_address, uint64_t expected, uint64_t status_address);
/* ... until EVERY one of flag_count consecutive 64-bit flags is >= expected (one flag per band rank: an aggregated counter
 * cannot tell two frames of a fast rank from one frame each of two ranks) */
KZN_API KznResult kzn_cmd_wait_flags(KznQueue* queue, uint64_t first_flag_address, uint32_t flag_count, uint64_t expected,
                                     uint64_t status_address);

/* device-to-device copy on the copy engines (no SM time): local or imported peer addresses; used to push a rendered band
 * to the presenting GPU while the SMs already render the next frame */
KZN_API KznResult kzn_queue_copy_dma(KznQueue* queue, uint64_t dst_address, uint64_t src_address, uint64_t size);

/* Measured fp32 issue rate of the device (tera lane-operations per second, an fma counting once): independent packed
 * fma.rn.f32x2 chains and nothing else.  The denominator of the shader-bound roofline. */
KZN_API KznResult kzn_calibrate_fp32_issue(KznQueue* queue, float* tera_lane_ops_per_second);

/* vkCmdPipelineBarrier api_impl.rs:4863 (unimplemented!() upstream).  A queue is one in-order CUDA stream: every command
 * already executes after all earlier commands of the queue have completed and made their writes visible, so execution
 * and memory barriers need no device work; the entry point exists so that a host can forward the call (and it is
 * recorded as nothing inside a command buffer, where it only ends the current batch of merged transfers — see the
 * command buffer notes below).  Image layout transitions between LINEAR and TILED images are explicit:
 * kzn_cmd_convert_tiling. */
KZN_API KznResult kzn_cmd_pipeline_barrier(KznQueue* queue);

/* plain transfers between host memory and device addresses on the queue (used by the ICD for
 * HOST_COHERENT emulation and by the end-to-end benchmark); host memory should be pinned */
KZN_API KznResult kzn_queue_upload(KznQueue* queue, uint64_t dst_address, const void* src, uint64_t size);
KZN_API KznResult kzn_queue_download(KznQueue* queue, void* dst, uint64_t src_address, uint64_t size);
KZN_API KznResult kzn_host_alloc_pinned(uint64_t size, void** ptr);
KZN_API void kzn_host_free_pinned(void* ptr);

/* -- command buffers: vkAllocateCommandBuffers api_impl.rs:4464, vkBeginCommandBuffer :4483, vkEndCommandBuffer :4491,
 * vkQueueSubmit :3414 (handle.rs:315 `CommandBuffer {}` is a placeholder upstream).
 * Between begin and end every kzn_cmd_* call on the queue is RECORDED (a CUDA graph captured from the queue's stream)
 * instead of executed; kzn_queue_submit replays recorded command buffers in order and then signals `fence` (may be NULL).
 * Replays cost one launch on the host whatever the number of commands.  While recording, consecutive contiguous
 * 16-byte-aligned fills and copies (whole-image clears, tightly packed buffer<->image copies, vkCmdFillBuffer,
 * vkCmdCopyBuffer) whose address ranges are pairwise disjoint are merged into ONE kernel launch of up to 96 commands:
 * Vulkan orders nothing between transfer commands without a barrier, and with disjoint ranges the result is the
 * in-order result anyway.  A command that reads or overwrites a range the batch touches, any other command, or
 * kzn_cmd_pipeline_barrier starts a new launch (KZN_MERGE_TRANSFERS=0 in the environment disables merging).
 * Not recordable: waits on the host, fence
 * signalling, and anything that allocates — a draw recorded into a command buffer needs its raster scratch reserved
 * first with kzn_queue_reserve_draw (largest triangle count and framebuffer it will see). */
typedef struct KznCommandBuffer KznCommandBuffer;
KZN_API KznResult kzn_command_buffer_begin(KznQueue* queue, KznCommandBuffer** command_buffer);
KZN_API KznResult kzn_command_buffer_end(KznCommandBuffer* command_buffer);
KZN_API void kzn_command_buffer_destroy(KznCommandBuffer* command_buffer);
KZN_API uint64_t kzn_command_buffer_launch_count(const KznCommandBuffer* command_buffer); /* kernels per replay */
KZN_API KznResult kzn_queue_submit(KznQueue* queue, uint32_t count, KznCommandBuffer* const* command_buffers, KznFence* fence);
KZN_API KznResult kzn_queue_reserve_draw(KznQueue* queue, uint64_t max_triangles, uint32_t width, uint32_t height);

/* -- semaphores: vkCreateSemaphore api_impl.rs:3700 (unimplemented!() upstream; handle.rs `Semaphore {}`) — binary,
 * device-side ordering between two queues of one device, as vkQueueSubmit's pSignalSemaphores / pWaitSemaphores:
 * work enqueued on `queue` after the wait starts only when everything enqueued before the matching signal has finished.
 * A signal must be enqueued (on the host) before the wait that consumes it. */
typedef struct KznSemaphore KznSemaphore;
KZN_API KznResult kzn_semaphore_create(KznDevice* device, KznSemaphore** semaphore);
KZN_API void kzn_semaphore_destroy(KznSemaphore* semaphore);
KZN_API KznResult kzn_queue_signal_semaphore(KznQueue* queue, KznSemaphore* semaphore);
KZN_API KznResult kzn_queue_wait_semaphore(KznQueue* queue, KznSemaphore* semaphore);

/* -- fences: vkCreateFence..vkWaitForFences api_impl.rs:3653-3697 ----------------------------- */
KZN_API KznResult kzn_fence_create(KznDevice* device, int32_t signaled, KznFence** fence);
KZN_API void kzn_fence_destroy(KznFence* fence);
KZN_API KznResult kzn_fence_reset(KznFence* fence);
KZN_API KznResult kzn_queue_signal_fence(KznQueue* queue, KznFence* fence); /* tail of vkQueueSubmit */
KZN_API KznResult kzn_fence_status(KznFence* fence);                         /* SUCCESS / NOT_READY */
KZN_API KznResult kzn_fence_wait(KznFence* fence, uint64_t timeout_ns);      /* SUCCESS / TIMEOUT */

/* ------------------------------------------------------------------------------------------ */
/* transfer commands                                               */

/* vkCmdClearColorImage  vulkan-driver/src/api_impl.rs:4780-4789 */
KZN_API KznResult kzn_cmd_clear_color_image(KznQueue* queue, const KznImageDesc* image,
                                            const KznClearColorValue* color, uint32_t range_count,
                                            const KznImageSubresourceRange* ranges);
/* vkCmdClearDepthStencilImage  api_impl.rs:4792-4801 (D32_SFLOAT: depth only) */
KZN_API KznResult kzn_cmd_clear_depth_stencil_image(KznQueue* queue, const KznImageDesc* image,
                                                    const KznClearDepthStencilValue* value,
                                                    uint32_t range_count,
                                                    const KznImageSubresourceRange* ranges);
/* vkCmdClearAttachments  api_impl.rs:4804-4812, for one colour or depth attachment image */
KZN_API KznResult kzn_cmd_clear_attachment(KznQueue* queue, const KznImageDesc* image,
                                           const KznClearColorValue* color_or_depth,
                                           uint32_t rect_count, const KznClearRect* rects);
/* vkCmdCopyBufferToImage  api_impl.rs:4734-4743 */
KZN_API KznResult kzn_cmd_copy_buffer_to_image(KznQueue* queue, uint64_t buffer_address,
                                               uint64_t buffer_size, const KznImageDesc* image,
                                               uint32_t region_count, const KznBufferImageCopy* regions);
/* vkCmdCopyImageToBuffer  api_impl.rs:4746-4755 */
KZN_API KznResult kzn_cmd_copy_image_to_buffer(KznQueue* queue, const KznImageDesc* image,
                                               uint64_t buffer_address, uint64_t buffer_size,
                                               uint32_t region_count, const KznBufferImageCopy* regions);
/* vkCmdCopyBuffer  api_impl.rs:4696-4703 */
KZN_API KznResult kzn_cmd_copy_buffer(KznQueue* queue, uint64_t src_address, uint64_t src_size,
                                      uint64_t dst_address, uint64_t dst_size, uint32_t region_count,
                                      const KznBufferCopy* regions);
/* vkCmdFillBuffer  api_impl.rs:4769-4777 (size may be KZN_WHOLE_SIZE) */
KZN_API KznResult kzn_cmd_fill_buffer(KznQueue* queue, uint64_t dst_address, uint64_t dst_size,
                                      uint64_t offset, uint64_t size, uint32_t data);
/* vkCmdUpdateBuffer  api_impl.rs:4758-4766 (host data is captured at call time, <= 65536 B) */
KZN_API KznResult kzn_cmd_update_buffer(KznQueue* queue, uint64_t dst_address, uint64_t dst_size,
                                        uint64_t offset, uint64_t size, const void* data);
/* vkCmdCopyImage  api_impl.rs:4707-4717 */
KZN_API KznResult kzn_cmd_copy_image(KznQueue* queue, const KznImageDesc* src, const KznImageDesc* dst,
                                     uint32_t region_count, const KznImageCopy* regions);
/* vkCmdBlitImage  api_impl.rs:4720-4731 (R8G8B8A8_UNORM <-> R8G8B8A8_UNORM) */
KZN_API KznResult kzn_cmd_blit_image(KznQueue* queue, const KznImageDesc* src, const KznImageDesc* dst,
                                     uint32_t region_count, const KznImageBlit* regions, uint32_t filter);
/* layout swizzle between Tiling::Linear and Tiling::Tiled (image.rs:17-22,48-57): same extents and
 * format, different `tiling`; what a layout transition / present of an OPTIMAL image performs */
KZN_API KznResult kzn_cmd_convert_tiling(KznQueue* queue, const KznImageDesc* src, const KznImageDesc* dst);

/* ------------------------------------------------------------------------------------------ */
/* draw path: vkCmdDraw api_impl.rs:4632, vkCmdDrawIndexed :4643,        */
/* vkCreateGraphicsPipelines :4018 -> GraphicsPipeline::create pipeline.rs:167-183 (all stubs)  */

/* Shader stages.  The reference cannot translate any shader with a body, so the
 * programmable stages are selected from built-in programs defined by the oracle (oracle/raster.c).  Compute entry
 * points go through the PTX back end (kazan_b200_backend.h). */
#define KZN_VS_PASSTHROUGH 0u      /* gl_Position = vec4(pos.xyz, pos.w or 1); colour passed through */
/* The vertex program bound with kzn_cmd_bind_vertex_program (kazan_b200_backend.h): a vertex shader in shader-compiler-ir
 * (for instance the translation of a SPIR-V module) run by the PTX back end, one invocation per vertex, before setup. */
#define KZN_VS_COMPILED 1u
#define KZN_FS_VERTEX_COLOR 0u     /* out = perspective-correct interpolated vertex rgba */
#define KZN_FS_PROCEDURAL_ORBIT 1u /* the fill-rate shader of BASELINE.json configs[2] (fp32 fma orbit, K=32) */
#define KZN_FS_ORBIT_ITERATIONS 32u
/* The fragment program bound with kzn_cmd_bind_fragment_program (kazan_b200_backend.h): a shader-compiler-ir module compiled
 * by the PTX back end into the raster kernels, which is how a GraphicsPipeline's fragment stage runs once the pipeline
 * exists (GraphicsPipeline::create, pipeline.rs:167-183); the two built-in programs above remain as the oracle-defined
 * fixed programs.  Its push constants and descriptor set are the GRAPHICS bind point's (kzn_cmd_push_constants,
 * kzn_cmd_bind_descriptor_set). */
#define KZN_FS_COMPILED 2u
#define KZN_MAX_PUSH_CONSTANTS_SIZE 128u /* maxPushConstantsSize */

typedef struct KznGraphicsPipelineState {
    uint32_t vertex_program;   /* KZN_VS_* */
    uint32_t fragment_program; /* KZN_FS_* */
    uint32_t cull_mode;        /* VkCullModeFlags: 0 none, 1 front, 2 back, 3 both */
    uint32_t front_face;       /* VkFrontFace: 0 counter-clockwise, 1 clockwise */
    uint32_t depth_test_enable;
    uint32_t depth_write_enable;
    uint32_t depth_compare_op; /* VkCompareOp 0..7 */
    uint32_t blend_enable;
    uint32_t src_color_blend_factor; /* VkBlendFactor (0..9 supported: ZERO..ONE_MINUS_DST_ALPHA) */
    uint32_t dst_color_blend_factor;
    uint32_t color_blend_op; /* VkBlendOp 0..4 */
    uint32_t src_alpha_blend_factor;
    uint32_t dst_alpha_blend_factor;
    uint32_t alpha_blend_op;
    uint32_t color_write_mask; /* VkColorComponentFlags, R=1 G=2 B=4 A=8 */
    /* VkPipelineColorBlendStateCreateInfo logicOpEnable / logicOp (the reference advertises the logicOp feature,
     * api_impl.rs:884): 0 = disabled, otherwise 1 + VkLogicOp (CLEAR = 1 ... SET = 16).  Applied to the UNORM8 bits of the
     * converted source and of the destination; when enabled, blending is treated as disabled (Vulkan 1.1 "Logical
     * Operations"); the write mask still applies. */
    uint32_t logic_op;
    /* VkPipelineInputAssemblyStateCreateInfo::topology: VK_PRIMITIVE_TOPOLOGY_TRIANGLE_LIST (3, and 0 is accepted as the
     * default), TRIANGLE_STRIP (4) or TRIANGLE_FAN (5), assembled as the specification's "Primitive Topologies" section
     * defines: strip triangle i = {v[i], v[i+(1+i%2)], v[i+(2-i%2)]}, fan triangle i = {v[i+1], v[i+2], v[0]}
     * (primitive restart is not supported); LINE_LIST (1) {v[2i], v[2i+1]}, LINE_STRIP (2) {v[i], v[i+1]} and POINT_LIST
     * (VK value 0: pass KZN_TOPOLOGY_POINT_LIST, since 0 reads as the default here).  Points are one pixel (pointSizeRange
     * [1, 1], api_impl.rs:1604), lines one pixel wide, non-strict (lineWidthRange [1, 1], strictLines FALSE, :1605-1608): the
     * parallelogram of "Basic Line Segment Rasterization", attributes running along the major axis (rule R9 of
     * oracle/raster.c).  They are never face-culled; a point outside the clip volume is discarded, a segment is cut to it.
     * Adjacency and patch topologies return VK_ERROR_FEATURE_NOT_PRESENT. */
    uint32_t topology;
    uint32_t reserved;
} KznGraphicsPipelineState;
#define KZN_TOPOLOGY_LINE_LIST 1u
#define KZN_TOPOLOGY_LINE_STRIP 2u
#define KZN_TOPOLOGY_TRIANGLE_LIST 3u
#define KZN_TOPOLOGY_TRIANGLE_STRIP 4u
#define KZN_TOPOLOGY_TRIANGLE_FAN 5u
#define KZN_TOPOLOGY_POINT_LIST 0x10u /* VK_PRIMITIVE_TOPOLOGY_POINT_LIST */

typedef struct KznVertexInput { /* one binding, VkVertexInputBindingDescription + 2 attributes */
    uint64_t buffer_address;
    uint64_t buffer_size;
    uint32_t stride;
    uint32_t position_offset;
    uint32_t position_components; /* 3 (R32G32B32_SFLOAT, w=1) or 4 (R32G32B32A32_SFLOAT) */
    uint32_t color_offset;        /* R32G32B32A32_SFLOAT; ~0u = no colour attribute (constant 1,1,1,1) */
} KznVertexInput;

typedef struct KznFramebuffer { /* vkCreateFramebuffer api_impl.rs:4383-4390 + render pass attachments */
    KznImageDesc color; /* R8G8B8A8_UNORM, LINEAR or TILED; format 0 = no colour attachment */
    KznImageDesc depth; /* D32_SFLOAT LINEAR; format 0 = no depth attachment */
    uint32_t color_layer, depth_layer;
    /* sort-first band splitting: only rows [band_y0, band_y1) are rendered;
     * band_y1 == 0 means the whole framebuffer */
    uint32_t band_y0, band_y1;
} KznFramebuffer;

/* vkCmdBeginRenderPass (api_impl.rs:4944-4958) / vkCmdEndRenderPass (:4961-4968) for one subpass with at most one colour
 * and one depth attachment — the part of a render pass that is data-path work: VK_ATTACHMENT_LOAD_OP_CLEAR.  The clear of
 * the render area (cut to the framebuffer's band) is not launched when the render pass begins: the first draw whose clip
 * rectangle (scissor x viewport x framebuffer x band) IS that area, on the same attachments, and that is large enough to
 * be binned, starts every 16 x 16 brick from the clear values instead of reading it and writes every pixel of the brick
 * back — no clear kernel, no read of the framebuffer (what a tiling GPU does with LOAD_OP_CLEAR).  Any other command, a draw
 * that does not qualify, or vkCmdEndRenderPass performs the clear with the transfer kernels first.  The result is the
 * same, byte for byte, as vkCmdClearAttachments on the area followed by the draws. */
#define KZN_ATTACHMENT_LOAD_OP_LOAD 0u
#define KZN_ATTACHMENT_LOAD_OP_CLEAR 1u
typedef struct KznRenderPassBegin {
    KznFramebuffer framebuffer;
    KznRect2D render_area;                 /* VkRenderPassBeginInfo::renderArea: inside the framebuffer */
    uint32_t color_load_op, depth_load_op; /* VkAttachmentDescription::loadOp (DONT_CARE may be passed as LOAD) */
    KznClearColorValue clear_color;        /* pClearValues of the two attachments */
    KznClearDepthStencilValue clear_depth_stencil;
} KznRenderPassBegin;
KZN_API KznResult kzn_cmd_begin_render_pass(KznQueue* queue, const KznRenderPassBegin* begin);
KZN_API KznResult kzn_cmd_end_render_pass(KznQueue* queue);

typedef struct KznDrawStatistics { /* optional, device-written counters (12 x u64) */
    uint64_t triangles_in, triangles_culled, triangles_binned, bin_entries, fragments_covered,
        fragments_depth_passed, tiles_launched, bin_overflow_tiles;
    /* primitive clipping (see "Clipping" above kzn_cmd_draw): triangles that crossed the w = 2^-20 plane or the guard band and
     * were clipped; the triangles the clipper made of them; triangles DISCARDED because a clip coordinate was NaN or infinite;
     * clipped triangles whose extra triangles did not fit the queue's reserve (always 0 unless a draw of more than 65536
     * triangles has more than one extra triangle per primitive on average: those lose the extra triangles) */
    uint64_t triangles_clipped, clip_triangles, triangles_discarded, clip_overflow;
} KznDrawStatistics;

/* Clipping (Vulkan 1.1 "Primitive Clipping"; rule R3 of oracle/raster.c).  The view volume cuts in x / y by scissoring to
 * the viewport rectangle and in z by a per-fragment test against [minDepth, maxDepth] (depthClamp is FALSE,
 * api_impl.rs:875-930) — neither changes a triangle.  A triangle with a vertex at w <= 0, or farther than 2^19 pixels from the
 * origin of the window, is clipped in homogeneous coordinates against w >= 2^-20 and a guard band of +-2^18 pixels and the
 * polygon (at most 8 vertices, colours interpolated with the positions) is drawn as a fan at the primitive's place in
 * primitive order; KznDrawStatistics counts them.  Only a triangle with a NaN or infinite clip coordinate is discarded
 * (triangles_discarded).  first_instance: there are no instance-rate vertex inputs and no built-in that reads the instance
 * index (the vertex stage is KZN_VS_PASSTHROUGH), so instances repeat the same triangles and first_instance has nothing to
 * offset; it is accepted for the signature's sake. */
KZN_API KznResult kzn_cmd_draw(KznQueue* queue, const KznGraphicsPipelineState* state,
                               const KznFramebuffer* framebuffer, const KznViewport* viewport,
                               const KznRect2D* scissor, const KznVertexInput* vertices,
                               uint32_t vertex_count, uint32_t instance_count, uint32_t first_vertex,
                               uint32_t first_instance);
KZN_API KznResult kzn_cmd_draw_indexed(KznQueue* queue, const KznGraphicsPipelineState* state,
                                       const KznFramebuffer* framebuffer, const KznViewport* viewport,
                                       const KznRect2D* scissor, const KznVertexInput* vertices,
                                       uint64_t index_buffer_address, uint64_t index_buffer_size,
                                       uint32_t index_type, uint32_t index_count,
                                       uint32_t instance_count, uint32_t first_index,
                                       int32_t vertex_offset, uint32_t first_instance);
/* ------------------------------------------------------------------------------------------ */
/* split frames: one frame over the GPUs of a box, in one process                */
/* The reference exposes one physical device (api_impl.rs:5092-5119): the split is internal to the driver.  A
 * KznSplitFrame opens the given devices, gives each a render queue, a whole-frame colour (and depth) image in its own HBM
 * and a band of rows (starts on multiples of 16, the raster brick height); rank 0 presents.  The host records the SAME draws on
 * every rank's queue with that rank's framebuffer (kzn_split_frame_framebuffer sets band_y0 / band_y1; geometry is
 * replicated: vertex and index buffers live on every device).  kzn_split_frame_composite is the one exchange step of the
 * frame: every band is pushed into the presenter's image by the copy engines over NVLink peer memory, ordered after the
 * band's rendering, and the presenter's queue is ordered after the pushes — all stream work, no host round trip. */
typedef struct KznSplitFrame KznSplitFrame;
KZN_API KznResult kzn_split_frame_create(uint32_t device_count, const int32_t* device_ordinals, uint32_t width, uint32_t height,
                                         uint32_t with_depth, KznSplitFrame** split_frame);
KZN_API void kzn_split_frame_destroy(KznSplitFrame* split_frame);
KZN_API uint32_t kzn_split_frame_device_count(const KznSplitFrame* split_frame);
KZN_API KznDevice* kzn_split_frame_device(const KznSplitFrame* split_frame, uint32_t rank);
KZN_API KznQueue* kzn_split_frame_queue(const KznSplitFrame* split_frame, uint32_t rank);
KZN_API KznResult kzn_split_frame_framebuffer(const KznSplitFrame* split_frame, uint32_t rank, KznFramebuffer* framebuffer);
KZN_API KznResult kzn_split_frame_composite(KznSplitFrame* split_frame);
KZN_API uint64_t kzn_split_frame_present_address(const KznSplitFrame* split_frame); /* the assembled image, on rank 0's device */
KZN_API KznQueue* kzn_split_frame_present_queue(const KznSplitFrame* split_frame);  /* ordered after every composite */
KZN_API KznResult kzn_split_frame_wait_idle(KznSplitFrame* split_frame);

/* ------------------------------------------------------------------------------------------ */
/* WSI presentation, data path                                            */
/* The reference prepares swapchains over X11 MIT-SHM (vulkan-driver/src/xcb_swapchain.rs:117-136,258-340: 32-bit pixmaps in
 * R8G8B8A8 or B8G8R8A8 byte order, scanline_pad, 2..16 images; SysV segments shm.rs:22-58) and leaves XcbSwapchain::new,
 * vkGetSwapchainImagesKHR, vkAcquireNextImageKHR and vkQueuePresentKHR unimplemented (xcb_swapchain.rs:342-349,
 * api_impl.rs:5656-5682).  A KznSwapchain is their GPU side: `image_count` presentable images (LINEAR R8G8B8A8_UNORM, in HBM:
 * colour attachments / transfer destinations like any image) and one SysV shared memory segment per image, created as
 * shm.rs does (shmget IPC_PRIVATE, IPC_CREAT | IPC_EXCL | 0666) and registered with CUDA.  kzn_queue_present, ordered after
 * the queue's earlier work, writes the image into its segment in the window's byte order with the window's scanline pitch
 * (copy engine when the orders agree, a byte-swapping kernel storing straight into the segment otherwise);
 * kzn_swapchain_wait_presented returns when the segment holds the frame — the point at which the host calls
 * xcb_shm_put_image(shmseg of that segment).  Images are acquired in order; an image whose previous present is still being
 * written is waited for (or VK_NOT_READY with timeout 0).  The xcb calls stay in the host. */
typedef struct KznSwapchain KznSwapchain;
#define KZN_MAX_SWAPCHAIN_IMAGES 16u /* xcb_swapchain.rs:136 MAX_SWAPCHAIN_IMAGE_COUNT */
#define KZN_WINDOW_R8G8B8A8 0u       /* SurfaceFormatGroup::R8G8B8A8 (xcb_swapchain.rs:117-118) */
#define KZN_WINDOW_B8G8R8A8 1u       /* SurfaceFormatGroup::B8G8R8A8 */
typedef struct KznSwapchainCreateInfo {
    uint32_t width, height;      /* VkSwapchainCreateInfoKHR::imageExtent = the window's size (xcb_swapchain.rs:309-311) */
    uint32_t image_count;        /* minImageCount, 2 .. KZN_MAX_SWAPCHAIN_IMAGES */
    uint32_t window_byte_order;  /* KZN_WINDOW_* from the window's pixmap format */
    uint32_t scanline_alignment; /* bytes: scanline_pad / 8 = 1, 2 or 4 (xcb_swapchain.rs:290-295) */
    uint32_t reserved;
} KznSwapchainCreateInfo;
typedef struct KznSwapchainSegment {
    int32_t shm_id;          /* SysV id: xcb_shm_attach(connection, shmseg, shm_id, read_only = 1) */
    uint32_t bytes_per_line; /* width * 4 rounded up to scanline_alignment */
    uint64_t size;           /* bytes_per_line * height */
    void* mapped;            /* the segment attached in this process */
} KznSwapchainSegment;
KZN_API KznResult kzn_swapchain_create(KznDevice* device, const KznSwapchainCreateInfo* info, KznSwapchain** swapchain);
KZN_API void kzn_swapchain_destroy(KznSwapchain* swapchain); /* waits for presents in flight; removes the segments */
KZN_API KznResult kzn_swapchain_get_image(const KznSwapchain* swapchain, uint32_t index, KznImageDesc* image);         /* vkGetSwapchainImagesKHR */
KZN_API KznResult kzn_swapchain_get_segment(const KznSwapchain* swapchain, uint32_t index, KznSwapchainSegment* segment);
KZN_API KznResult kzn_swapchain_acquire_next_image(KznSwapchain* swapchain, uint64_t timeout_ns, uint32_t* index);      /* vkAcquireNextImageKHR */
KZN_API KznResult kzn_queue_present(KznQueue* queue, KznSwapchain* swapchain, uint32_t index);                          /* vkQueuePresentKHR */
KZN_API KznResult kzn_swapchain_wait_presented(KznSwapchain* swapchain, uint32_t index);

/* where (device address of a KznDrawStatistics, zeroed by the caller) the next draws accumulate
 * counters; 0 disables */
KZN_API KznResult kzn_queue_set_draw_statistics(KznQueue* queue, uint64_t stats_address);

#ifdef __cplusplus
}
#endif
#endif /* KAZAN_B200_H */
