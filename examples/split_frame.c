/* One frame over N GPUs of a box through the C ABI alone (no Python, no torch, one process): SURVEY.md §8e.
 *
 * A jittered grid mesh (depth LESS + alpha blend, the shape of BASELINE.json configs[3] / [4]) is rendered twice: by ONE
 * device, and split sort-first into N bands over N devices whose bands are composited to device 0 over NVLink peer memory
 * (kzn_split_frame_create / _framebuffer / _composite).  Bands are disjoint rows and every GPU keeps API order inside its
 * band, so the two frames must be identical byte for byte; the program compares them and exits 0 with "ok".
 *
 *   gcc -std=c11 -O2 -Iinclude examples/split_frame.c -Lkazan_b200 -lkazan_b200 -Wl,-rpath,$PWD/kazan_b200 -o examples/split_frame
 *   examples/split_frame [gpus] [width height quads_x quads_y]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kazan_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        KznResult r_ = (call);                                                        \
        if (r_ < 0) {                                                                 \
            fprintf(stderr, "%s failed: %d: %s\n", #call, (int)r_, kzn_last_error()); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static float unit(uint64_t* s) { return (float)(splitmix64(s) >> 40) * (1.0f / 16777216.0f); }

/* renders the scene with `gpus` devices into `out` (W * H * 4 bytes of host memory) */
static int render(uint32_t gpus, uint32_t W, uint32_t H, const float* verts, uint64_t vbytes, const uint32_t* idx, uint64_t ibytes, uint32_t index_count,
                  uint8_t* out) {
    int32_t ordinals[16];
    for (uint32_t i = 0; i < gpus; ++i) ordinals[i] = (int32_t)i;
    KznSplitFrame* sf = NULL;
    CHECK(kzn_split_frame_create(gpus, ordinals, W, H, 1, &sf));
    KznGraphicsPipelineState st;
    memset(&st, 0, sizeof st);
    st.fragment_program = KZN_FS_VERTEX_COLOR;
    st.color_write_mask = 15;
    st.depth_test_enable = st.depth_write_enable = 1;
    st.depth_compare_op = 1; /* VK_COMPARE_OP_LESS */
    st.blend_enable = 1;
    st.src_color_blend_factor = st.src_alpha_blend_factor = 6; /* SRC_ALPHA */
    st.dst_color_blend_factor = st.dst_alpha_blend_factor = 7; /* ONE_MINUS_SRC_ALPHA */
    KznViewport vp = {0, 0, (float)W, (float)H, 0, 1};
    KznRect2D sc = {{0, 0}, {W, H}};
    KznMemory* vmem[16];
    KznMemory* imem[16];
    for (uint32_t i = 0; i < gpus; ++i) { /* geometry is replicated; every rank records the same commands */
        KznDevice* dev = kzn_split_frame_device(sf, i);
        KznQueue* q = kzn_split_frame_queue(sf, i);
        CHECK(kzn_memory_allocate(dev, vbytes, KZN_MEMORY_TYPE_DEVICE_LOCAL, &vmem[i]));
        CHECK(kzn_memory_allocate(dev, ibytes, KZN_MEMORY_TYPE_DEVICE_LOCAL, &imem[i]));
        CHECK(kzn_queue_upload(q, kzn_memory_device_address(vmem[i]), verts, vbytes));
        CHECK(kzn_queue_upload(q, kzn_memory_device_address(imem[i]), idx, ibytes));
        KznFramebuffer fb;
        CHECK(kzn_split_frame_framebuffer(sf, i, &fb));
        /* vkCmdBeginRenderPass with LOAD_OP_CLEAR on both attachments: the framebuffer names this rank's band, so only the
         * band is cleared — and the draw takes the clear over (no clear kernels, no framebuffer read) */
        KznRenderPassBegin rp;
        memset(&rp, 0, sizeof rp);
        rp.framebuffer = fb;
        rp.render_area = sc;
        rp.color_load_op = rp.depth_load_op = KZN_ATTACHMENT_LOAD_OP_CLEAR;
        rp.clear_depth_stencil.depth = 1.0f; /* colour: 0, 0, 0, 0 */
        CHECK(kzn_cmd_begin_render_pass(q, &rp));
        KznVertexInput vi = {kzn_memory_device_address(vmem[i]), vbytes, 28, 0, 3, 12};
        CHECK(kzn_cmd_draw_indexed(q, &st, &fb, &vp, &sc, &vi, kzn_memory_device_address(imem[i]), ibytes, KZN_INDEX_TYPE_UINT32, index_count, 1, 0, 0, 0));
        CHECK(kzn_cmd_end_render_pass(q));
    }
    CHECK(kzn_split_frame_composite(sf)); /* the one exchange step: bands -> device 0 */
    CHECK(kzn_queue_download(kzn_split_frame_present_queue(sf), out, kzn_split_frame_present_address(sf), (uint64_t)W * H * 4));
    CHECK(kzn_split_frame_wait_idle(sf));
    for (uint32_t i = 0; i < gpus; ++i) {
        kzn_memory_free(vmem[i]);
        kzn_memory_free(imem[i]);
    }
    kzn_split_frame_destroy(sf);
    return 0;
}

int main(int argc, char** argv) {
    int32_t count = 0;
    CHECK(kzn_device_count(&count));
    if (count < 1) {
        fprintf(stderr, "no CUDA device: kazan_b200 has no CPU fallback\n");
        return 2;
    }
    uint32_t gpus = argc > 1 ? (uint32_t)atoi(argv[1]) : (uint32_t)(count < 2 ? count : 2);
    const uint32_t W = argc > 5 ? (uint32_t)atoi(argv[2]) : 1280, H = argc > 5 ? (uint32_t)atoi(argv[3]) : 720;
    const uint32_t QX = argc > 5 ? (uint32_t)atoi(argv[4]) : 300, QY = argc > 5 ? (uint32_t)atoi(argv[5]) : 170;
    if (gpus < 1 || (int32_t)gpus > count || gpus > 16) {
        fprintf(stderr, "%u GPUs asked for, %d present\n", gpus, count);
        return 2;
    }
    /* the mesh of SURVEY.md §8d: a grid of quads whose vertices are jittered by +-1.5 cells, so neighbouring triangles overlap */
    const uint32_t vx = QX + 1, vy = QY + 1;
    const uint64_t nv = (uint64_t)vx * vy, ni = 6ull * QX * QY;
    float* verts = malloc(nv * 7 * sizeof(float));
    uint32_t* idx = malloc(ni * sizeof(uint32_t));
    uint8_t *one = malloc((uint64_t)W * H * 4), *split = malloc((uint64_t)W * H * 4);
    if (!verts || !idx || !one || !split) return 1;
    uint64_t seed = 5;
    for (uint32_t j = 0; j < vy; ++j)
        for (uint32_t i = 0; i < vx; ++i) {
            float* v = verts + 7 * ((uint64_t)j * vx + i);
            v[0] = -1.0f + ((float)i + (unit(&seed) - 0.5f) * 3.0f) * (2.0f / (float)QX);
            v[1] = -1.0f + ((float)j + (unit(&seed) - 0.5f) * 3.0f) * (2.0f / (float)QY);
            v[2] = 0.05f + 0.9f * unit(&seed);
            v[3] = unit(&seed);
            v[4] = unit(&seed);
            v[5] = unit(&seed);
            v[6] = 0.25f + 0.75f * unit(&seed);
        }
    for (uint32_t j = 0; j < QY; ++j)
        for (uint32_t i = 0; i < QX; ++i) {
            uint32_t* q = idx + 6 * ((uint64_t)j * QX + i);
            const uint32_t v00 = j * vx + i, v10 = v00 + 1, v01 = v00 + vx, v11 = v01 + 1;
            q[0] = v00; q[1] = v10; q[2] = v11;
            q[3] = v00; q[4] = v11; q[5] = v01;
        }
    if (render(1, W, H, verts, nv * 28, idx, ni * 4, (uint32_t)ni, one)) return 1;
    if (render(gpus, W, H, verts, nv * 28, idx, ni * 4, (uint32_t)ni, split)) return 1;
    uint64_t bad = 0, lit = 0;
    for (uint64_t i = 0; i < (uint64_t)W * H * 4; ++i) {
        bad += one[i] != split[i];
        lit += one[i] != 0;
    }
    printf("{\"frame\": [%u, %u], \"triangles\": %llu, \"gpus\": %u, \"nonzero_bytes\": %llu, \"mismatching_bytes\": %llu}\n", W, H,
           (unsigned long long)(ni / 3), gpus, (unsigned long long)lit, (unsigned long long)bad);
    free(verts); free(idx); free(one); free(split);
    if (bad || !lit) return 3;
    puts("ok");
    return 0;
}
