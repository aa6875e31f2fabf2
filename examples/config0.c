/* BASELINE.json configs[0] through the C ABI alone (no Python, no torch): vkCmdClearColorImage + vkCmdCopyBufferToImage on
 * a 1024x1024 R8G8B8A8_UNORM image, recorded into a command buffer, submitted with a fence, read back through mapped
 * HOST_COHERENT memory — the calls the reference's Rust host would make in place of its unimplemented!() stubs
 * (vulkan-driver/src/api_impl.rs:4780, :4734, :4483-4491, :3414, :3653-3697, :3434-3500).
 *
 *   gcc -std=c11 -O2 -Iinclude examples/config0.c -Lkazan_b200 -lkazan_b200 -Wl,-rpath,$PWD/kazan_b200 -o examples/config0
 *
 * Exit code 0 and "ok" when the image holds the expected bytes: the clear colour {0.25, 0.5, 0.75, 1} -> 40 80 BF FF
 * outside the copied sub-rectangle (imageOffset (16,8), extent 992x1000), the staging buffer's bytes inside it. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "kazan_b200.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        KznResult r_ = (call);                                                                   \
        if (r_ < 0) {                                                                            \
            fprintf(stderr, "%s failed: %d: %s\n", #call, (int)r_, kzn_last_error());            \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

static uint64_t splitmix64(uint64_t* s) { /* the staging pattern of SURVEY.md §8d */
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(void) {
    const uint32_t W = 1024, H = 1024, RW = 992, RH = 1000, OX = 16, OY = 8;
    int32_t count = 0;
    CHECK(kzn_device_count(&count));
    if (count < 1) {
        fprintf(stderr, "no CUDA device: kazan_b200 has no CPU fallback\n");
        return 2;
    }
    KznDevice* dev = NULL;
    KznQueue* queue = NULL;
    CHECK(kzn_device_create(0, &dev));
    CHECK(kzn_queue_create(dev, &queue));

    KznImageDesc image = {0, KZN_FORMAT_R8G8B8A8_UNORM, KZN_TILING_LINEAR, W, H, 1, 1, 1, 0};
    uint64_t image_size = 0, image_align = 0;
    CHECK(kzn_image_memory_requirements(&image, &image_size, &image_align));
    const uint64_t staging_size = (uint64_t)RW * RH * 4;

    /* memory type 0 = the reference's only type: HOST_VISIBLE | HOST_COHERENT (pinned, mapped); type 1 = HBM */
    KznMemory *image_mem = NULL, *staging_mem = NULL, *readback_mem = NULL;
    CHECK(kzn_memory_allocate(dev, image_size, 1, &image_mem));
    CHECK(kzn_memory_allocate(dev, staging_size, 0, &staging_mem));
    CHECK(kzn_memory_allocate(dev, image_size, 0, &readback_mem));
    image.address = kzn_memory_device_address(image_mem);
    uint8_t *staging = NULL, *readback = NULL;
    CHECK(kzn_memory_map(staging_mem, 0, (void**)&staging));
    CHECK(kzn_memory_map(readback_mem, 0, (void**)&readback));
    uint64_t seed = 0x4B415A414Eull;
    for (uint64_t i = 0; i < staging_size; i += 8) {
        uint64_t v = splitmix64(&seed);
        memcpy(staging + i, &v, 8);
    }

    KznClearColorValue color = {{0.25f, 0.5f, 0.75f, 1.0f}};
    KznImageSubresourceRange range = {KZN_IMAGE_ASPECT_COLOR_BIT, 0, KZN_REMAINING_MIP_LEVELS, 0, KZN_REMAINING_ARRAY_LAYERS};
    KznBufferImageCopy to_image = {0, 0, 0, {KZN_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {(int32_t)OX, (int32_t)OY, 0}, {RW, RH, 1}};
    KznBufferImageCopy to_host = {0, 0, 0, {KZN_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1}, {0, 0, 0}, {W, H, 1}};

    KznCommandBuffer* cb = NULL;
    KznFence* fence = NULL;
    CHECK(kzn_fence_create(dev, 0, &fence));
    CHECK(kzn_command_buffer_begin(queue, &cb)); /* vkBeginCommandBuffer */
    CHECK(kzn_cmd_clear_color_image(queue, &image, &color, 1, &range));
    CHECK(kzn_cmd_pipeline_barrier(queue));
    CHECK(kzn_cmd_copy_buffer_to_image(queue, kzn_memory_device_address(staging_mem), staging_size, &image, 1, &to_image));
    CHECK(kzn_cmd_pipeline_barrier(queue));
    CHECK(kzn_cmd_copy_image_to_buffer(queue, &image, kzn_memory_device_address(readback_mem), image_size, 1, &to_host));
    CHECK(kzn_command_buffer_end(cb)); /* vkEndCommandBuffer */
    CHECK(kzn_queue_submit(queue, 1, &cb, fence)); /* vkQueueSubmit */
    CHECK(kzn_fence_wait(fence, ~0ull));            /* vkWaitForFences */

    const uint8_t clear_bytes[4] = {0x40, 0x80, 0xBF, 0xFF};
    uint64_t bad = 0;
    for (uint32_t y = 0; y < H; ++y)
        for (uint32_t x = 0; x < W; ++x) {
            const uint8_t* got = readback + 4ull * ((uint64_t)y * W + x);
            const uint8_t* want = clear_bytes;
            if (x >= OX && x < OX + RW && y >= OY && y < OY + RH) want = staging + 4ull * ((uint64_t)(y - OY) * RW + (x - OX));
            bad += memcmp(got, want, 4) != 0;
        }
    printf("{\"config\": \"clear + copy 1024x1024 through the C ABI\", \"kernel_launches_per_submit\": %llu, \"mismatching_texels\": %llu}\n",
           (unsigned long long)kzn_command_buffer_launch_count(cb), (unsigned long long)bad);

    kzn_command_buffer_destroy(cb);
    kzn_fence_destroy(fence);
    kzn_memory_free(readback_mem);
    kzn_memory_free(staging_mem);
    kzn_memory_free(image_mem);
    kzn_queue_destroy(queue);
    kzn_device_destroy(dev);
    if (bad) return 3;
    puts("ok");
    return 0;
}
