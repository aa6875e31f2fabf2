/* oracle/transfer.c — CPU restatement of the transfer family.  TEST INFRASTRUCTURE (see kzn_oracle.h).
 *
 * PARITY UNPINNED for results: the reference bodies are `unimplemented!()`
 * (vulkan-driver/src/api_impl.rs:4696-4812).  Pinned by the reference and followed here:
 *   - linear layout and size rule      vulkan-driver/src/image.rs:58-88
 *   - IMAGE_ALIGNMENT / BUFFER_ALIGNMENT vulkan-driver/src/lib.rs:33-34
 *   - struct layouts                    external/Vulkan-Headers/include/vulkan/vulkan_core.h:2899-2919
 * Everything else follows Vulkan 1.1.130 "Clear Commands" / "Copy Commands" as written down in
 * SURVEY.md Appendix E.1-E.4, E.10 and the valid-usage rules of Appendix G.
 */
#include "kzn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

void kzo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int kzo_get_threads(void) { return g_threads; }
int kzo_hardware_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* ---- E.1 ---------------------------------------------------------------------------------- */

uint8_t kzo_float_to_unorm8(float f) {
    /* clamp with NaN -> 0, one fp32 multiply, round-to-nearest-even */
    float c;
    if (!(f == f)) c = 0.0f;
    else if (f < 0.0f) c = 0.0f;
    else if (f > 1.0f) c = 1.0f;
    else c = f;
    return (uint8_t)nearbyintf(c * 255.0f);
}

float kzo_unorm8_to_float(uint8_t u) { return (float)u / 255.0f; }

uint32_t kzo_pack_rgba8(const float rgba[4]) {
    /* R8G8B8A8 is a byte-array format: R at the lowest address */
    return (uint32_t)kzo_float_to_unorm8(rgba[0]) | ((uint32_t)kzo_float_to_unorm8(rgba[1]) << 8) |
           ((uint32_t)kzo_float_to_unorm8(rgba[2]) << 16) | ((uint32_t)kzo_float_to_unorm8(rgba[3]) << 24);
}

/* ---- layout ------------------------------------------------------------------------------- */

uint32_t kzo_texel_size(uint32_t format) {
    switch (format) {
    case KZN_FORMAT_R8G8B8A8_UNORM: return 4; /* image.rs:69 pixel_size_in_bytes */
    case KZN_FORMAT_D32_SFLOAT: return 4;     /* SURVEY.md Appendix G extension */
    default: return 0;
    }
}

static uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

static uint64_t slice_pitch(const KznImageDesc* im) {
    uint64_t ts = kzo_texel_size(im->format);
    if (im->tiling == KZN_TILING_TILED) {
        uint64_t tx = ((uint64_t)im->width + KZN_TILE_DIM - 1) / KZN_TILE_DIM;
        uint64_t ty = ((uint64_t)im->height + KZN_TILE_DIM - 1) / KZN_TILE_DIM;
        return tx * ty * KZN_TILE_DIM * KZN_TILE_DIM * ts;
    }
    return ts * im->width * im->height;
}

uint64_t kzo_texel_offset(const KznImageDesc* im, uint32_t x, uint32_t y, uint32_t z, uint32_t layer) {
    uint64_t ts = kzo_texel_size(im->format);
    uint64_t slice = (uint64_t)z + (uint64_t)im->depth * layer;
    if (im->tiling == KZN_TILING_TILED) {
        uint64_t tiles_x = ((uint64_t)im->width + KZN_TILE_DIM - 1) / KZN_TILE_DIM;
        uint64_t tile = (uint64_t)(y / KZN_TILE_DIM) * tiles_x + x / KZN_TILE_DIM;
        uint64_t within = (uint64_t)(y % KZN_TILE_DIM) * KZN_TILE_DIM + x % KZN_TILE_DIM;
        return slice * slice_pitch(im) + (tile * KZN_TILE_DIM * KZN_TILE_DIM + within) * ts;
    }
    /* image.rs:70-80: tightly packed, no row padding */
    return ts * ((uint64_t)x + (uint64_t)im->width * ((uint64_t)y + (uint64_t)im->height * slice));
}

static KznResult check_image(const KznImageDesc* im) {
    if (!im) return KZN_ERROR_VALIDATION_FAILED;
    if (kzo_texel_size(im->format) == 0) return KZN_ERROR_FORMAT_NOT_SUPPORTED;
    if (im->mip_levels != 1) return KZN_ERROR_FEATURE_NOT_PRESENT; /* image.rs:65 */
    if (im->tiling > KZN_TILING_TILED) return KZN_ERROR_VALIDATION_FAILED;
    if (!im->width || !im->height || !im->depth || !im->array_layers) return KZN_ERROR_VALIDATION_FAILED;
    { /* the reference computes the size with checked_mul().unwrap() (image.rs:74-82): an overflowing size is a panic */
        unsigned long long w = im->width, h = im->height, n = 0;
        /* VkOffset3D is signed 32-bit: larger extents could not be addressed by any region */
        if (im->width > 0x7fffffffu || im->height > 0x7fffffffu || im->depth > 0x7fffffffu || im->array_layers > 0x7fffffffu)
            return KZN_ERROR_VALIDATION_FAILED;
        if (im->tiling == KZN_TILING_TILED) {
            w = (unsigned long long)(((uint64_t)im->width + KZN_TILE_DIM - 1) / KZN_TILE_DIM) * KZN_TILE_DIM;
            h = (unsigned long long)(((uint64_t)im->height + KZN_TILE_DIM - 1) / KZN_TILE_DIM) * KZN_TILE_DIM;
        }
        if (__builtin_mul_overflow(w, h, &n) || __builtin_mul_overflow(n, (unsigned long long)im->depth, &n) ||
            __builtin_mul_overflow(n, (unsigned long long)im->array_layers, &n) || __builtin_mul_overflow(n, 4ull, &n) ||
            n >= (1ull << 62))
            return KZN_ERROR_VALIDATION_FAILED;
    }
    return KZN_SUCCESS;
}

KznResult kzo_image_memory_requirements(const KznImageDesc* im, uint64_t* size, uint64_t* alignment) {
    KznResult r = check_image(im);
    if (r) return r;
    uint64_t bytes = slice_pitch(im) * im->depth * im->array_layers;
    /* DeviceMemoryLayout::calculate(size, IMAGE_ALIGNMENT) rounds the size up to the alignment
     * (device_memory.rs:166-181) */
    if (size) *size = round_up(bytes, KZN_IMAGE_ALIGNMENT);
    if (alignment) *alignment = KZN_IMAGE_ALIGNMENT;
    return KZN_SUCCESS;
}

static uint8_t* texel_ptr(const KznImageDesc* im, uint32_t x, uint32_t y, uint32_t z, uint32_t layer) {
    return (uint8_t*)(uintptr_t)im->address + kzo_texel_offset(im, x, y, z, layer);
}

/* ---- E.2 clears --------------------------------------------------------------------------- */

static KznResult resolve_range(const KznImageDesc* im, const KznImageSubresourceRange* r, uint32_t aspect,
                               uint32_t* layer0, uint32_t* layers) {
    if (r->aspectMask != aspect) return KZN_ERROR_VALIDATION_FAILED;
    if (r->baseMipLevel != 0) return KZN_ERROR_VALIDATION_FAILED;
    if (r->levelCount != KZN_REMAINING_MIP_LEVELS && r->levelCount != 1) return KZN_ERROR_VALIDATION_FAILED;
    if (r->baseArrayLayer >= im->array_layers) return KZN_ERROR_VALIDATION_FAILED;
    uint32_t n = r->layerCount == KZN_REMAINING_ARRAY_LAYERS ? im->array_layers - r->baseArrayLayer : r->layerCount;
    if (n == 0 || n > im->array_layers - r->baseArrayLayer) return KZN_ERROR_VALIDATION_FAILED;
    *layer0 = r->baseArrayLayer;
    *layers = n;
    return KZN_SUCCESS;
}

static void fill_layers(const KznImageDesc* im, uint32_t layer0, uint32_t layers, uint32_t word) {
    /* both layouts store whole layers contiguously (padding texels of TILED edge tiles included) */
    uint64_t pitch = slice_pitch(im) * im->depth;
    uint32_t* p = (uint32_t*)((uint8_t*)(uintptr_t)im->address + pitch * layer0);
    int64_t n = (int64_t)(pitch * layers / 4);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = word;
}

KznResult kzo_clear_color_image(const KznImageDesc* im, const KznClearColorValue* color, uint32_t range_count,
                                const KznImageSubresourceRange* ranges) {
    KznResult r = check_image(im);
    if (r) return r;
    if (im->format != KZN_FORMAT_R8G8B8A8_UNORM) return KZN_ERROR_FORMAT_NOT_SUPPORTED;
    if (!color || !range_count || !ranges) return KZN_ERROR_VALIDATION_FAILED;
    uint32_t word = kzo_pack_rgba8(color->float32);
    for (uint32_t i = 0; i < range_count; ++i) {
        uint32_t l0, n;
        if ((r = resolve_range(im, &ranges[i], KZN_IMAGE_ASPECT_COLOR_BIT, &l0, &n))) return r;
    }
    for (uint32_t i = 0; i < range_count; ++i) {
        uint32_t l0 = 0, n = 0;
        resolve_range(im, &ranges[i], KZN_IMAGE_ASPECT_COLOR_BIT, &l0, &n);
        fill_layers(im, l0, n, word);
    }
    return KZN_SUCCESS;
}

KznResult kzo_clear_depth_stencil_image(const KznImageDesc* im, const KznClearDepthStencilValue* value,
                                        uint32_t range_count, const KznImageSubresourceRange* ranges) {
    KznResult r = check_image(im);
    if (r) return r;
    if (im->format != KZN_FORMAT_D32_SFLOAT) return KZN_ERROR_FORMAT_NOT_SUPPORTED;
    if (!value || !range_count || !ranges) return KZN_ERROR_VALIDATION_FAILED;
    uint32_t word;
    memcpy(&word, &value->depth, 4); /* D32 stores the fp32 bits verbatim (E.8: no clamp) */
    for (uint32_t i = 0; i < range_count; ++i) {
        uint32_t l0, n;
        if ((r = resolve_range(im, &ranges[i], KZN_IMAGE_ASPECT_DEPTH_BIT, &l0, &n))) return r;
    }
    for (uint32_t i = 0; i < range_count; ++i) {
        uint32_t l0 = 0, n = 0;
        resolve_range(im, &ranges[i], KZN_IMAGE_ASPECT_DEPTH_BIT, &l0, &n);
        fill_layers(im, l0, n, word);
    }
    return KZN_SUCCESS;
}

KznResult kzo_clear_attachment(const KznImageDesc* im, const KznClearColorValue* v, uint32_t rect_count,
                               const KznClearRect* rects) {
    KznResult r = check_image(im);
    if (r) return r;
    if (!v || !rect_count || !rects) return KZN_ERROR_VALIDATION_FAILED;
    if (im->depth != 1) return KZN_ERROR_VALIDATION_FAILED;
    uint32_t word = im->format == KZN_FORMAT_R8G8B8A8_UNORM ? kzo_pack_rgba8(v->float32) : v->uint32[0];
    for (uint32_t i = 0; i < rect_count; ++i) {
        const KznClearRect* c = &rects[i];
        if (c->rect.offset.x < 0 || c->rect.offset.y < 0) return KZN_ERROR_VALIDATION_FAILED;
        if ((uint64_t)c->rect.offset.x + c->rect.extent.width > im->width) return KZN_ERROR_VALIDATION_FAILED;
        if ((uint64_t)c->rect.offset.y + c->rect.extent.height > im->height) return KZN_ERROR_VALIDATION_FAILED;
        if (c->layerCount == 0 || (uint64_t)c->baseArrayLayer + c->layerCount > im->array_layers)
            return KZN_ERROR_VALIDATION_FAILED;
    }
    for (uint32_t i = 0; i < rect_count; ++i) {
        const KznClearRect* c = &rects[i];
        for (uint32_t l = 0; l < c->layerCount; ++l)
            for (uint32_t y = 0; y < c->rect.extent.height; ++y)
                for (uint32_t x = 0; x < c->rect.extent.width; ++x)
                    memcpy(texel_ptr(im, c->rect.offset.x + x, c->rect.offset.y + y, 0, c->baseArrayLayer + l),
                           &word, 4);
    }
    return KZN_SUCCESS;
}

/* ---- E.3 buffer <-> image ------------------------------------------------------------------ */

static uint32_t aspect_of(uint32_t format) {
    return format == KZN_FORMAT_D32_SFLOAT ? KZN_IMAGE_ASPECT_DEPTH_BIT : KZN_IMAGE_ASPECT_COLOR_BIT;
}

static KznResult check_subresource_box(const KznImageDesc* im, const KznImageSubresourceLayers* s,
                                       const KznOffset3D* off, const KznExtent3D* ext) {
    if (s->aspectMask != aspect_of(im->format)) return KZN_ERROR_VALIDATION_FAILED; /* 00212 */
    if (s->mipLevel != 0) return KZN_ERROR_VALIDATION_FAILED;
    if (s->layerCount == 0 || (uint64_t)s->baseArrayLayer + s->layerCount > im->array_layers)
        return KZN_ERROR_VALIDATION_FAILED;
    if (off->x < 0 || off->y < 0 || off->z < 0) return KZN_ERROR_VALIDATION_FAILED;
    if ((uint64_t)off->x + ext->width > im->width) return KZN_ERROR_VALIDATION_FAILED;   /* 00197 */
    if ((uint64_t)off->y + ext->height > im->height) return KZN_ERROR_VALIDATION_FAILED; /* 00198 */
    if ((uint64_t)off->z + ext->depth > im->depth) return KZN_ERROR_VALIDATION_FAILED;   /* 00200 */
    if (im->depth > 1 && (s->baseArrayLayer != 0 || s->layerCount != 1)) return KZN_ERROR_VALIDATION_FAILED;
    return KZN_SUCCESS;
}

static KznResult check_buffer_image_region(const KznImageDesc* im, uint64_t buffer_size,
                                           const KznBufferImageCopy* c) {
    KznResult r = check_subresource_box(im, &c->imageSubresource, &c->imageOffset, &c->imageExtent);
    if (r) return r;
    uint64_t ts = kzo_texel_size(im->format);
    if (c->bufferOffset % 4 || c->bufferOffset % ts) return KZN_ERROR_VALIDATION_FAILED; /* 00194/01558 */
    if (c->bufferRowLength && c->bufferRowLength < c->imageExtent.width) return KZN_ERROR_VALIDATION_FAILED;   /* 00195 */
    if (c->bufferImageHeight && c->bufferImageHeight < c->imageExtent.height) return KZN_ERROR_VALIDATION_FAILED; /* 00196 */
    uint64_t w = c->imageExtent.width, h = c->imageExtent.height;
    uint64_t slices = (uint64_t)c->imageExtent.depth * c->imageSubresource.layerCount;
    if (w && h && slices) {
        uint64_t row_len = c->bufferRowLength ? c->bufferRowLength : w;
        uint64_t img_h = c->bufferImageHeight ? c->bufferImageHeight : h;
        /* 128-bit: bufferRowLength and bufferImageHeight are free 32-bit values, the product may pass 2^64 */
        unsigned __int128 end = (unsigned __int128)c->bufferOffset +
                                ((((unsigned __int128)(slices - 1) * img_h) + (h - 1)) * row_len + w) * ts;
        if (end > buffer_size) return KZN_ERROR_VALIDATION_FAILED; /* 00171 */
    }
    return KZN_SUCCESS;
}

static void copy_buffer_image(uint8_t* buffer, const KznImageDesc* im, const KznBufferImageCopy* c, int to_image) {
    uint64_t ts = kzo_texel_size(im->format);
    uint64_t w = c->imageExtent.width, h = c->imageExtent.height, d = c->imageExtent.depth;
    uint64_t row_len = c->bufferRowLength ? c->bufferRowLength : w;
    uint64_t img_h = c->bufferImageHeight ? c->bufferImageHeight : h;
    int64_t rows = (int64_t)(h * d * c->imageSubresource.layerCount);
    if (!w || rows == 0) return;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int64_t row = 0; row < rows; ++row) {
        uint64_t y = (uint64_t)row % h;
        uint64_t s = (uint64_t)row / h; /* slice index in the buffer: layer-major, then z */
        uint32_t z = (uint32_t)(s % d), l = (uint32_t)(s / d);
        uint8_t* b = buffer + c->bufferOffset + ((s * img_h + y) * row_len) * ts;
        if (im->tiling == KZN_TILING_LINEAR) {
            uint8_t* t = texel_ptr(im, c->imageOffset.x, c->imageOffset.y + (uint32_t)y, c->imageOffset.z + z,
                                   c->imageSubresource.baseArrayLayer + l);
            if (to_image) memcpy(t, b, w * ts);
            else memcpy(b, t, w * ts);
        } else {
            for (uint64_t x = 0; x < w; ++x) {
                uint8_t* t = texel_ptr(im, c->imageOffset.x + (uint32_t)x, c->imageOffset.y + (uint32_t)y,
                                       c->imageOffset.z + z, c->imageSubresource.baseArrayLayer + l);
                if (to_image) memcpy(t, b + x * ts, ts);
                else memcpy(b + x * ts, t, ts);
            }
        }
    }
}

KznResult kzo_copy_buffer_to_image(const void* buffer, uint64_t buffer_size, const KznImageDesc* im,
                                   uint32_t region_count, const KznBufferImageCopy* regions) {
    KznResult r = check_image(im);
    if (r) return r;
    if (!buffer || !region_count || !regions) return KZN_ERROR_VALIDATION_FAILED;
    for (uint32_t i = 0; i < region_count; ++i)
        if ((r = check_buffer_image_region(im, buffer_size, &regions[i]))) return r;
    for (uint32_t i = 0; i < region_count; ++i) copy_buffer_image((uint8_t*)(uintptr_t)buffer, im, &regions[i], 1);
    return KZN_SUCCESS;
}

KznResult kzo_copy_image_to_buffer(const KznImageDesc* im, void* buffer, uint64_t buffer_size,
                                   uint32_t region_count, const KznBufferImageCopy* regions) {
    KznResult r = check_image(im);
    if (r) return r;
    if (!buffer || !region_count || !regions) return KZN_ERROR_VALIDATION_FAILED;
    for (uint32_t i = 0; i < region_count; ++i)
        if ((r = check_buffer_image_region(im, buffer_size, &regions[i]))) return r;
    for (uint32_t i = 0; i < region_count; ++i) copy_buffer_image((uint8_t*)buffer, im, &regions[i], 0);
    return KZN_SUCCESS;
}

/* ---- buffer commands ------------------------------------------------------------------------ */

KznResult kzo_copy_buffer(const void* src, uint64_t src_size, void* dst, uint64_t dst_size,
                          uint32_t region_count, const KznBufferCopy* regions) {
    if (!src || !dst || !region_count || !regions) return KZN_ERROR_VALIDATION_FAILED;
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznBufferCopy* c = &regions[i];
        if (c->size == 0) return KZN_ERROR_VALIDATION_FAILED; /* VUID-VkBufferCopy-size-01988 */
        if (c->srcOffset > src_size || c->size > src_size - c->srcOffset) return KZN_ERROR_VALIDATION_FAILED;
        if (c->dstOffset > dst_size || c->size > dst_size - c->dstOffset) return KZN_ERROR_VALIDATION_FAILED;
    }
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznBufferCopy* c = &regions[i];
        const uint8_t* s = (const uint8_t*)src + c->srcOffset;
        uint8_t* d = (uint8_t*)dst + c->dstOffset;
        int64_t chunks = (int64_t)((c->size + 65535) / 65536);
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int64_t k = 0; k < chunks; ++k) {
            uint64_t o = (uint64_t)k * 65536;
            uint64_t n = c->size - o < 65536 ? c->size - o : 65536;
            memcpy(d + o, s + o, n);
        }
    }
    return KZN_SUCCESS;
}

KznResult kzo_fill_buffer(void* dst, uint64_t dst_size, uint64_t offset, uint64_t size, uint32_t data) {
    if (!dst) return KZN_ERROR_VALIDATION_FAILED;
    if (offset >= dst_size || offset % 4) return KZN_ERROR_VALIDATION_FAILED; /* 00024/00025 */
    if (size == KZN_WHOLE_SIZE) size = (dst_size - offset) & ~3ull;         /* rounded down to 4 */
    else if (size == 0 || size % 4 || size > dst_size - offset) return KZN_ERROR_VALIDATION_FAILED;
    uint32_t* p = (uint32_t*)((uint8_t*)dst + offset);
    int64_t n = (int64_t)(size / 4);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = data;
    return KZN_SUCCESS;
}

KznResult kzo_update_buffer(void* dst, uint64_t dst_size, uint64_t offset, uint64_t size, const void* data) {
    if (!dst || !data) return KZN_ERROR_VALIDATION_FAILED;
    if (offset % 4 || size % 4 || size == 0 || size > 65536) return KZN_ERROR_VALIDATION_FAILED; /* 00036-00038 */
    if (offset >= dst_size || size > dst_size - offset) return KZN_ERROR_VALIDATION_FAILED;
    memcpy((uint8_t*)dst + offset, data, size);
    return KZN_SUCCESS;
}

/* ---- image <-> image ------------------------------------------------------------------------ */

KznResult kzo_copy_image(const KznImageDesc* src, const KznImageDesc* dst, uint32_t region_count,
                         const KznImageCopy* regions) {
    KznResult r;
    if ((r = check_image(src)) || (r = check_image(dst))) return r;
    if (!region_count || !regions) return KZN_ERROR_VALIDATION_FAILED;
    /* size-compatible formats only (both supported formats are 4 bytes/texel); bytes are copied verbatim */
    if (kzo_texel_size(src->format) != kzo_texel_size(dst->format)) return KZN_ERROR_VALIDATION_FAILED;
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageCopy* c = &regions[i];
        if ((r = check_subresource_box(src, &c->srcSubresource, &c->srcOffset, &c->extent))) return r;
        if ((r = check_subresource_box(dst, &c->dstSubresource, &c->dstOffset, &c->extent))) return r;
        if (c->srcSubresource.layerCount != c->dstSubresource.layerCount) return KZN_ERROR_VALIDATION_FAILED;
    }
    uint64_t ts = kzo_texel_size(src->format);
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageCopy* c = &regions[i];
        uint64_t w = c->extent.width, h = c->extent.height, d = c->extent.depth;
        int64_t rows = (int64_t)(h * d * c->srcSubresource.layerCount);
        if (!w || !rows) continue;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int64_t row = 0; row < rows; ++row) {
            uint32_t y = (uint32_t)((uint64_t)row % h);
            uint64_t s = (uint64_t)row / h;
            uint32_t z = (uint32_t)(s % d), l = (uint32_t)(s / d);
            for (uint32_t x = 0; x < w; ++x)
                memcpy(texel_ptr(dst, c->dstOffset.x + x, c->dstOffset.y + y, c->dstOffset.z + z,
                                 c->dstSubresource.baseArrayLayer + l),
                       texel_ptr(src, c->srcOffset.x + x, c->srcOffset.y + y, c->srcOffset.z + z,
                                 c->srcSubresource.baseArrayLayer + l),
                       ts);
        }
    }
    return KZN_SUCCESS;
}

/* ---- E.4 blit -------------------------------------------------------------------------------- */

static int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

static KznResult check_blit_box(const KznImageDesc* im, const KznImageSubresourceLayers* s, const KznOffset3D o[2]) {
    if (s->aspectMask != KZN_IMAGE_ASPECT_COLOR_BIT || s->mipLevel != 0) return KZN_ERROR_VALIDATION_FAILED;
    if (s->layerCount == 0 || (uint64_t)s->baseArrayLayer + s->layerCount > im->array_layers)
        return KZN_ERROR_VALIDATION_FAILED;
    for (int k = 0; k < 2; ++k) { /* 00243-00252: offsets within [0, dim] */
        if (o[k].x < 0 || (uint32_t)o[k].x > im->width) return KZN_ERROR_VALIDATION_FAILED;
        if (o[k].y < 0 || (uint32_t)o[k].y > im->height) return KZN_ERROR_VALIDATION_FAILED;
        if (o[k].z < 0 || (uint32_t)o[k].z > im->depth) return KZN_ERROR_VALIDATION_FAILED;
    }
    if (im->depth > 1 && (s->baseArrayLayer != 0 || s->layerCount != 1)) return KZN_ERROR_VALIDATION_FAILED;
    return KZN_SUCCESS;
}

/* unnormalised source coordinate of destination texel centre i */
static float blit_coord(int32_t i, int32_t d0, int32_t d1, int32_t s0, int32_t s1) {
    float scale = (float)(s1 - s0) / (float)(d1 - d0);
    return fmaf(((float)i + 0.5f) - (float)d0, scale, (float)s0);
}

KznResult kzo_blit_image(const KznImageDesc* src, const KznImageDesc* dst, uint32_t region_count,
                         const KznImageBlit* regions, uint32_t filter) {
    KznResult r;
    if ((r = check_image(src)) || (r = check_image(dst))) return r;
    if (src->format != KZN_FORMAT_R8G8B8A8_UNORM || dst->format != KZN_FORMAT_R8G8B8A8_UNORM)
        return KZN_ERROR_FORMAT_NOT_SUPPORTED;
    if (!region_count || !regions || filter > KZN_FILTER_LINEAR) return KZN_ERROR_VALIDATION_FAILED;
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageBlit* b = &regions[i];
        if ((r = check_blit_box(src, &b->srcSubresource, b->srcOffsets))) return r;
        if ((r = check_blit_box(dst, &b->dstSubresource, b->dstOffsets))) return r;
        if (b->srcSubresource.layerCount != b->dstSubresource.layerCount) return KZN_ERROR_VALIDATION_FAILED;
    }
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageBlit* b = &regions[i];
        const KznOffset3D *s0 = &b->srcOffsets[0], *s1 = &b->srcOffsets[1];
        const KznOffset3D *d0 = &b->dstOffsets[0], *d1 = &b->dstOffsets[1];
        if (d0->x == d1->x || d0->y == d1->y || d0->z == d1->z) continue; /* empty destination box */
        int32_t x_lo = d0->x < d1->x ? d0->x : d1->x, x_hi = d0->x < d1->x ? d1->x : d0->x;
        int32_t y_lo = d0->y < d1->y ? d0->y : d1->y, y_hi = d0->y < d1->y ? d1->y : d0->y;
        int32_t z_lo = d0->z < d1->z ? d0->z : d1->z, z_hi = d0->z < d1->z ? d1->z : d0->z;
        for (uint32_t l = 0; l < b->srcSubresource.layerCount; ++l)
            for (int32_t k = z_lo; k < z_hi; ++k) {
                /* the third coordinate is always point-sampled */
                float wf = blit_coord(k, d0->z, d1->z, s0->z, s1->z);
                int32_t sz = clampi((int32_t)floorf(wf), 0, (int32_t)src->depth - 1);
#pragma omp parallel for num_threads(g_threads) schedule(static)
                for (int32_t j = y_lo; j < y_hi; ++j) {
                    float v = blit_coord(j, d0->y, d1->y, s0->y, s1->y);
                    for (int32_t ii = x_lo; ii < x_hi; ++ii) {
                        float u = blit_coord(ii, d0->x, d1->x, s0->x, s1->x);
                        uint8_t out[4];
                        uint32_t sl = b->srcSubresource.baseArrayLayer + l;
                        if (filter == KZN_FILTER_NEAREST) {
                            int32_t sx = clampi((int32_t)floorf(u), 0, (int32_t)src->width - 1);
                            int32_t sy = clampi((int32_t)floorf(v), 0, (int32_t)src->height - 1);
                            memcpy(out, texel_ptr(src, sx, sy, sz, sl), 4);
                        } else {
                            float uu = u - 0.5f, vv = v - 0.5f;
                            float fu = floorf(uu), fv = floorf(vv);
                            float a = uu - fu, bb = vv - fv;
                            int32_t x0 = clampi((int32_t)fu, 0, (int32_t)src->width - 1);
                            int32_t x1 = clampi((int32_t)fu + 1, 0, (int32_t)src->width - 1);
                            int32_t y0 = clampi((int32_t)fv, 0, (int32_t)src->height - 1);
                            int32_t y1 = clampi((int32_t)fv + 1, 0, (int32_t)src->height - 1);
                            const uint8_t* t00 = texel_ptr(src, x0, y0, sz, sl);
                            const uint8_t* t10 = texel_ptr(src, x1, y0, sz, sl);
                            const uint8_t* t01 = texel_ptr(src, x0, y1, sz, sl);
                            const uint8_t* t11 = texel_ptr(src, x1, y1, sz, sl);
                            float w00 = (1.0f - a) * (1.0f - bb), w10 = a * (1.0f - bb);
                            float w01 = (1.0f - a) * bb, w11 = a * bb;
                            for (int ch = 0; ch < 4; ++ch) {
                                float acc = w00 * kzo_unorm8_to_float(t00[ch]);
                                acc = acc + w10 * kzo_unorm8_to_float(t10[ch]);
                                acc = acc + w01 * kzo_unorm8_to_float(t01[ch]);
                                acc = acc + w11 * kzo_unorm8_to_float(t11[ch]);
                                out[ch] = kzo_float_to_unorm8(acc);
                            }
                        }
                        memcpy(texel_ptr(dst, ii, j, k, b->dstSubresource.baseArrayLayer + l), out, 4);
                    }
                }
            }
    }
    return KZN_SUCCESS;
}

/* ---- E.10 tiling swizzle ---------------------------------------------------------------------- */

KznResult kzo_convert_tiling(const KznImageDesc* src, const KznImageDesc* dst) {
    KznResult r;
    if ((r = check_image(src)) || (r = check_image(dst))) return r;
    if (src->format != dst->format || src->width != dst->width || src->height != dst->height ||
        src->depth != dst->depth || src->array_layers != dst->array_layers)
        return KZN_ERROR_VALIDATION_FAILED;
    uint64_t ts = kzo_texel_size(src->format);
    int64_t rows = (int64_t)((uint64_t)src->height * src->depth * src->array_layers);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int64_t row = 0; row < rows; ++row) {
        uint32_t y = (uint32_t)((uint64_t)row % src->height);
        uint64_t s = (uint64_t)row / src->height;
        uint32_t z = (uint32_t)(s % src->depth), l = (uint32_t)(s / src->depth);
        for (uint32_t x = 0; x < src->width; ++x) memcpy(texel_ptr(dst, x, y, z, l), texel_ptr(src, x, y, z, l), ts);
    }
    return KZN_SUCCESS;
}

/* ---- deterministic inputs ----------------------------------------------------------------------- */

uint64_t kzo_splitmix64(uint64_t* state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void kzo_fill_random_bytes(void* dst, uint64_t size, uint64_t seed) {
    uint64_t st = seed;
    uint8_t* p = (uint8_t*)dst;
    uint64_t i = 0;
    for (; i + 8 <= size; i += 8) {
        uint64_t v = kzo_splitmix64(&st);
        memcpy(p + i, &v, 8);
    }
    if (i < size) {
        uint64_t v = kzo_splitmix64(&st);
        memcpy(p + i, &v, size - i);
    }
}
