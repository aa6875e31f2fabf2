/* kzn_oracle.h — CPU oracle for the kazan_b200 hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link, load or call anything declared here.  The product (kazan_b200/libkazan_b200.so) never does.
 *
 * PARITY STATUS — read before trusting:
 *   The reference (kazan-3d/kazan @ f2820899) leaves every function on this path `unimplemented!()`
 *   (vulkan-driver/src/api_impl.rs:4632-4812) and cannot be compiled here (no Rust toolchain, LLVM 7
 *   is fetched over the network by shader-compiler-backend-llvm-7/build.rs:35-41).  There is
 *   therefore NO upstream golden vector for transfer, raster or dispatch RESULTS:
 *       transfer / raster / shading arithmetic  ........  PARITY UNPINNED (oracle is normative)
 *   What the reference does pin, and what this oracle is checked against in tests/:
 *       image/buffer memory layout & sizes  ....  vulkan-driver/src/image.rs:58-88, lib.rs:27-35
 *       struct layouts  ........................  external/Vulkan-Headers vulkan_core.h (1.1.130)
 *       .kazan-ir text goldens, SPIR-V->IR golden  (oracle/ir_*.c, tests/golden/ir_goldens.json)
 *   The rules this file restates are the Vulkan 1.1.130 specification chapters named in
 *   SURVEY.md Appendix E, with every implementation-defined choice written down there.
 *
 * All `address` fields of KznImageDesc are HOST pointers here.
 * Compile with -ffp-contract=off: every fused multiply-add is an explicit fmaf().
 */
#ifndef KZN_ORACLE_H
#define KZN_ORACLE_H

#include "../include/kazan_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* number of worker threads used by the bulk loops (OpenMP); 1 = scalar port */
void kzo_set_threads(int n);
int kzo_get_threads(void);
int kzo_hardware_threads(void);

/* E.1 float <-> UNORM8 */
uint8_t kzo_float_to_unorm8(float f);
float kzo_unorm8_to_float(uint8_t u);
uint32_t kzo_pack_rgba8(const float rgba[4]);

/* layout: vulkan-driver/src/image.rs:58-88 (LINEAR), SURVEY.md E.10 (TILED) */
uint32_t kzo_texel_size(uint32_t format); /* 0 = unsupported */
uint64_t kzo_texel_offset(const KznImageDesc* image, uint32_t x, uint32_t y, uint32_t z, uint32_t layer);
KznResult kzo_image_memory_requirements(const KznImageDesc* image, uint64_t* size, uint64_t* alignment);

/* transfer commands — same argument meaning as the kzn_cmd_* functions of kazan_b200.h */
KznResult kzo_clear_color_image(const KznImageDesc* image, const KznClearColorValue* color,
                                uint32_t range_count, const KznImageSubresourceRange* ranges);
KznResult kzo_clear_depth_stencil_image(const KznImageDesc* image, const KznClearDepthStencilValue* value,
                                        uint32_t range_count, const KznImageSubresourceRange* ranges);
KznResult kzo_clear_attachment(const KznImageDesc* image, const KznClearColorValue* color_or_depth,
                               uint32_t rect_count, const KznClearRect* rects);
KznResult kzo_copy_buffer_to_image(const void* buffer, uint64_t buffer_size, const KznImageDesc* image,
                                   uint32_t region_count, const KznBufferImageCopy* regions);
KznResult kzo_copy_image_to_buffer(const KznImageDesc* image, void* buffer, uint64_t buffer_size,
                                   uint32_t region_count, const KznBufferImageCopy* regions);
KznResult kzo_copy_buffer(const void* src, uint64_t src_size, void* dst, uint64_t dst_size,
                          uint32_t region_count, const KznBufferCopy* regions);
KznResult kzo_fill_buffer(void* dst, uint64_t dst_size, uint64_t offset, uint64_t size, uint32_t data);
KznResult kzo_update_buffer(void* dst, uint64_t dst_size, uint64_t offset, uint64_t size, const void* data);
KznResult kzo_copy_image(const KznImageDesc* src, const KznImageDesc* dst, uint32_t region_count,
                         const KznImageCopy* regions);
KznResult kzo_blit_image(const KznImageDesc* src, const KznImageDesc* dst, uint32_t region_count,
                         const KznImageBlit* regions, uint32_t filter);
KznResult kzo_convert_tiling(const KznImageDesc* src, const KznImageDesc* dst);

/* draw path (oracle/raster.c) — same argument meaning as kzn_cmd_draw / kzn_cmd_draw_indexed;
 * every address is a host pointer.  `stats` may be NULL. */
KznResult kzo_draw(const KznGraphicsPipelineState* state, const KznFramebuffer* framebuffer,
                   const KznViewport* viewport, const KznRect2D* scissor, const KznVertexInput* vertices,
                   uint32_t vertex_count, uint32_t instance_count, uint32_t first_vertex,
                   uint32_t first_instance, KznDrawStatistics* stats);
KznResult kzo_draw_indexed(const KznGraphicsPipelineState* state, const KznFramebuffer* framebuffer,
                           const KznViewport* viewport, const KznRect2D* scissor,
                           const KznVertexInput* vertices, const void* index_buffer,
                           uint64_t index_buffer_size, uint32_t index_type, uint32_t index_count,
                           uint32_t instance_count, uint32_t first_index, int32_t vertex_offset,
                           uint32_t first_instance, KznDrawStatistics* stats);
/* the built-in fragment programs, callable on their own (tests) */
void kzo_fs_procedural_orbit(uint32_t px, uint32_t py, float fb_width, float fb_height, float rgba[4]);

/* deterministic input generators shared by tests and bench */
uint64_t kzo_splitmix64(uint64_t* state);
void kzo_fill_random_bytes(void* dst, uint64_t size, uint64_t seed);
/* grid mesh of quads_x*quads_y*2 triangles, vertex = {x,y,z,r,g,b,a} f32 (28 B), u32 indices
 * (6*quads_x*quads_y); positions jittered by splitmix64(seed).  Buffers are caller-allocated. */
void kzo_make_grid_mesh(uint32_t quads_x, uint32_t quads_y, uint64_t seed, float* vertices, uint32_t* indices);

#ifdef __cplusplus
}
#endif
#endif
