// raster.cu — the draw path as hand-written sm_100a kernels.
//
// Replaces the `unimplemented!()` bodies of vkCmdDraw / vkCmdDrawIndexed
// (vulkan-driver/src/api_impl.rs:4632-4652) and GraphicsPipeline::create (vulkan-driver/src/pipeline.rs:167-183).
// Limits honoured: subPixelPrecisionBits 4 (api_impl.rs:1559), single sample at the pixel centre
// (standardSampleLocations, api_impl.rs:1609), one viewport (api_impl.rs:1566), no depth clamp (api_impl.rs:887).
//
// Pipeline (one stream, no host round trip; the kernels of a draw are chained by programmatic dependent launch):
//   setup_kernel   1 thread / triangle: primitive assembly (list / strip / fan), index + vertex fetch, passthrough VS,
//                  clip->window (R1), snap to 1/16 px (R2), signed area / cull (R4), sign-normalised integer edge equations
//                  + top-left bias (R5), pixel box.  One 144-byte TriRec per triangle that can produce fragments (written
//                  through shared memory as coalesced 16-byte stores) and one packed 8-byte pixel box per triangle; the first
//                  binning pass (tile counts) is fused in.
//   bin_scan / bin_fill   screen-tile binning from the packed boxes: per 128x8-pixel tile, the list of overlapping triangles
//                  (atomics give an unordered list; the raster kernel sorts it so that per-pixel fragment order == primitive
//                  order, R7).
//   raster_kernel  binned draws.  1 CTA / 128x8 tile, warp w = row w; the tile's colour and depth live in shared memory for
//                  the whole list (each framebuffer byte read at most once, written at most once).  After the in-tile radix
//                  sort every row warp streams the list on its own (no CTA barrier): box filter -> pending list -> chunks of
//                  32 triangles copied to private slots by cp.async -> exact spans -> candidate pixels -> fragment queue ->
//                  shading in full 32-lane batches.  Spans of >= 64 pixels take a 4-pixels-per-lane path.
//   raster_direct_kernel  draws of <= 256 triangles (fullscreen passes, quads).  Persistent one-warp CTAs over a work counter,
//                  256x4-pixel blocks, 8 pixels per lane in registers; blocks that the last triangle covers completely and
//                  opaquely are pure shader work (packed FFMA2 orbit, magic-add UNORM8 conversion).  Draws of <= 32
//                  triangles have no setup_kernel: each warp makes the records itself.
// Coverage: int64 edge functions with a bias test (R5); depth: barycentric fp32 (R6), z clip, early-Z (R7); then the fragment
// program, blend or logic op, write mask (R8).
// Numerics are the oracle's rules R1-R8 (oracle/raster.c): every fp32 operation is written out; the file is compiled
// with -fmad=false so only the explicit fmaf() calls fuse.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <cstddef>
#include <new>
#include <vector>

#include "kzn_internal.h"

namespace {

constexpr int kBrick = 16;                     // raster_kernel: a warp owns one 16 x 16 pixel brick of its CTA's tile
constexpr int kTileW = 4 * kBrick, kTileH = 2 * kBrick;  // pixels per raster tile (= binning granularity): 4 x 2 bricks
constexpr int kTileBricks = 8;                 // bricks per tile
constexpr int kSortThreads = 256;              // tile_sort_kernel: one CTA per tile
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRasterWarps = 4;                // raster_kernel: independent warps per CTA, each takes bricks from a work counter
constexpr int kRasterThreads = 32 * kRasterWarps;
#ifndef KZN_RASTER_CTAS
#define KZN_RASTER_CTAS 6
#endif
constexpr int kBrickPx = 8;                    // prologue / epilogue / wide path: lane l = row l >> 1, pixels 8 (l & 1) .. + 7 of the brick
constexpr int kDirectPx = 8;                   // raster_direct_kernel: pixels per lane (two 16-byte vectors)
constexpr int kDirectTileW = 32 * kDirectPx;   // a warp's block is 256 pixels wide ...
#ifndef KZN_DIRECT_ROWS
#define KZN_DIRECT_ROWS 4
#endif
#ifndef KZN_DIRECT_WARPS
#define KZN_DIRECT_WARPS 20
#endif
constexpr int kDirectRows = KZN_DIRECT_ROWS;          // ... and this many rows high
constexpr int kDirectWarpsPerSM = KZN_DIRECT_WARPS;   // persistent one-warp CTAs per SM (bounds the registers per thread)
constexpr int kSortCap = 4096;           // largest binned list sorted in shared memory
constexpr float kGuardBand = 524288.0f;  // 2^19 pixels (R3)
constexpr float kClipGuard = 262144.0f;  // 2^18 pixels: where clipped polygons end (R3)
constexpr float kClipW = 9.5367431640625e-07f;  // 2^-20 (R3)
constexpr uint32_t kFusedMax = 5;        // direct draws of at most this many triangles make their records in the raster kernel (5 x 6 clipped triangles fit 32 slots)
// work_counter slots: 0, 1 block dispenser of the direct kernel and its leavers; 2 brick dispenser of raster_kernel; 3 records
// of a direct draw; 4 extra records allocated by this draw's clipper (zeroed again by its sort kernel); 5 their count for
// raster_kernel; 6 primitives setup_kernel left to the clipper in bin_scan_kernel (zeroed by the sort kernel); 7 CTAs of
// bin_scan_kernel that are done; 8, 9 band_prefilter_kernel's blocks with work and their triangles (zeroed by the sort kernel);
// 16 .. 19 the vertex (0, 0, 0, 1) that out-of-range indices read
constexpr uint32_t kDirectMax = 256;     // draws with at most this many triangles skip binning
// how a kernel instantiation treats the destination colour (template parameter NEED_DST)
constexpr int kDstNone = 0;       // no blending, all channels written: the destination is never read
constexpr int kDstGeneric = 1;    // any blend state / write mask, resolved per fragment
constexpr int kDstAlphaOver = 2;  // SRC_ALPHA, ONE_MINUS_SRC_ALPHA, ADD for colour and alpha, write mask RGBA

struct __align__(16) TriRec {  // expanded record: what the direct kernel and the wide-span path rasterise from
    int32_t minx, maxx, miny, maxy;  // inclusive, scissored; minx > maxx => produces no fragments
    int32_t a[3], b[3];
    uint32_t bias_bits;
    float inv2a;
    long long c[3];
    float z[3], iw[3];
    float col[3][4];  // colour * iw
};
static_assert(sizeof(TriRec) == 144, "TriRec layout");

// Compact record of a binned draw (96 bytes, six 16-byte vectors): the snapped window coordinates and the per-vertex
// interpolants.  Everything else a fragment needs — edge equations, top-left bias, 1/2A, the pixel box — is integer
// arithmetic on X, Y and is re-derived per (triangle, row) by the row warp that rasterises it (derive in raster_chunk,
// expand_record for the wide path): 50-odd instructions per triangle-row against 48 bytes per record that setup would
// write and every row warp would read again.
struct __align__(16) TriRecC {
    int32_t X[3], Y[3];  // window coordinates in 1/16 pixel (R2)
    float z[3], iw[3];
    float col[3][4];  // colour * iw
};
static_assert(sizeof(TriRecC) == 96, "TriRecC layout");

// What lane j of a brick's warp derives for triangle j of its chunk: the three edge functions at the first pixel of the
// triangle's pixel box inside the brick, biased so that "covered" is a sign test, and their steps per pixel in x and y.
// 32-bit arithmetic is exact here because only triangles whose snapped extent is below 2^14 subpixels (1024 pixels) in
// x and y take this path: |a|,|b| < 2^14 and every sample of the pixel box is within the extent, so |E_i| < 2^29
// (larger triangles go through expand_record and the 64-bit path).
struct __align__(16) BrickEdge {
    int32_t f[3];   // F_i = E_i + bias_i - 1 at the box's first pixel: a pixel is covered iff all three are >= 0 (R5)
    uint32_t xw;    // x0 | y0 << 4 (first pixel within the brick) | box width << 8 | (1 - bias_i) << (13 + i) | depth-clip-cannot-reject << 16
    int32_t s[3];   // 16 * a_i: step per pixel in x
    float inv2a;
    int32_t u[3];   // 16 * b_i: step per pixel in y
    uint32_t magic; // ceil(2^16 / width): candidate k of the box is (k % width, k / width), k / width == k * magic >> 16 for k < 256
};
static_assert(sizeof(BrickEdge) == 48, "BrickEdge layout");

struct DrawParams {
    // geometry
    unsigned long long vertex_base, vertex_size;
    unsigned long long index_base;  // 0 => non-indexed
    uint32_t stride, pos_off, pos_comps, col_off;
    uint32_t index_type;
    uint32_t first;  // firstIndex / firstVertex
    int32_t vertex_offset;
    uint32_t tris_per_instance, total_tris;
    uint32_t instances;                         // total_tris / tris_per_instance
    unsigned long long pos_vertices, col_vertices;  // vertices 0 .. n-1 have their position / colour wholly inside the vertex buffer (robust fetch: the others read zeros)
    const float* null_vertex;                       // (0, 0, 0, 1) in device memory: what an out-of-range position reads
    // state
    KznGraphicsPipelineState st;
    float hw, hh, ox, oy, dz, zmin_vp;  // viewport transform (R1)
    float zclip_lo, zclip_hi;
    int32_t cx0, cy0, cx1, cy1;  // inclusive clip rectangle (scissor ∩ viewport ∩ framebuffer ∩ band)
    // framebuffer
    unsigned long long color_base, depth_base;  // layer already applied; 0 => absent
    uint32_t fb_w, fb_h, color_tiled, color_tiles_x;
    float inv_fb_w, inv_fb_h;
    // scratch
    TriRec* tris;            // direct draws: expanded records; binned draws: the same memory holds TriRecC records
    uint2* boxes;            // [total_tris] minx | maxx << 16, miny | maxy << 16 (the records' pixel boxes, packed: the row filter reads only these)
    uint32_t* tile_counts;   // [tiles]  (bin_count)
    uint32_t* tile_offsets;  // [tiles+1] (bin_scan)
    uint32_t* tile_cursor;   // [tiles]  (bin_fill)
    uint32_t* bin_entries;   // [bin_capacity]
    uint32_t* bin_entries_alt;  // [bin_capacity] second buffer of the in-tile sort
    uint8_t* bin_masks;      // [bin_capacity] per sorted entry: the bricks of its tile that the triangle's pixel box meets (tile_sort_kernel)
    uint32_t bin_capacity;
    uint32_t tiles_x, tiles_y, tile_x0, tile_y0;  // tile grid covering the clip rectangle
    uint32_t binned;                               // 0 => every tile walks all triangles
    unsigned long long* stats;                     // KznDrawStatistics or null
    uint32_t* work_counter;                        // direct draws: next block to hand out (zeroed by setup_kernel)
    uint32_t exact_spans;                          // raster_kernel: narrow spans are reduced to the covered interval first
    uint32_t fused_setup;                          // direct draws of <= kFusedMax triangles: raster_direct_kernel makes the records itself
    // vkCmdBeginRenderPass with LOAD_OP_CLEAR taken over by this (binned) draw: bit 0 — every brick's colour starts as
    // clear_color and is written back whole, bit 1 — the same for depth (no clear kernel, no framebuffer read)
    uint32_t load_clear, clear_color;
    float clear_depth;
    // compiled fragment programs (KZN_FS_COMPILED): what vkCmdBindDescriptorSets / vkCmdPushConstants left on the queue, and
    // what the program reads (bit 0: it needs interpolated inputs or the z / w of frag_coord, so no coverage shortcut applies)
    uint32_t fs_flags;
    uint32_t rows_first;  // setup: decide on y before x, z and the colours are fetched (the clip rectangle keeps under half of the viewport's rows)
    // bands of split frames (rows_first, one instance, many triangles): band_prefilter_kernel looks at the rows of every
    // triangle first — index and y / w fetches only, no shared memory, full occupancy — and leaves one byte per block of 256
    // triangles: 0 = none of them can touch the band's rows.  setup_kernel and bin_fill_kernel skip those blocks (nothing is
    // written for their triangles, not even the packed boxes); null: no prefilter ran.
    const uint8_t* block_flags;
    // ... and the list of the blocks that stay (work_counter[8] of them, any order; work_counter[9] = their triangles): CTA i of
    // setup_kernel / bin_fill_kernel takes block live_blocks[i], the CTAs beyond the list exit — after the ones with work were
    // started, whatever part of the index range the band's triangles lie in
    const uint32_t* live_blocks;
    // clipped primitives of binned draws (R3): the first triangle of the clipped polygon takes the primitive's own record, the
    // others ("extra" records, shared by all instances) follow the draw's records at tris / boxes [total_tris + k]
    uint32_t* clip_bits;    // [tris_per_instance / 32] bit t: primitive t has extra records (zeroed with the tile counters)
    uint32_t* clip_link;    // [tris_per_instance] first extra record k | number of extras << 29; valid where the primitive's bit is set
    uint32_t* clip_parent;  // [clip_capacity] primitive << 3 | fan index j (1..5) of each extra record; ~0u: unused
    uint32_t* clip_list;    // [tris_per_instance] the primitives setup_kernel left to the clipper (work_counter[6] of them)
    uint32_t clip_capacity;
    unsigned long long descriptor_table;
    unsigned char push_constants[KZN_MAX_PUSH_CONSTANTS_SIZE];
};

// ---- setup ------------------------------------------------------------------------------------------------

// robust fetch of one vertex's colour (1, 1, 1, 1 without a colour attribute; out of range reads 0, 0, 0, 1)
__device__ __forceinline__ float4 fetch_color(const DrawParams& p, long long vidx) {
    float4 o = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (p.col_off != ~0u) {
        if ((unsigned long long)vidx < p.col_vertices) {
            const float* s = reinterpret_cast<const float*>(p.vertex_base + (unsigned long long)vidx * p.stride + p.col_off);
            o = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), __ldg(s + 3));
        } else {
            o.x = o.y = o.z = 0.0f;
        }
    }
    return o;
}

__device__ __forceinline__ bool finite_f(float f) { return (f - f) == 0.0f; }

__device__ __forceinline__ bool tile_range(const DrawParams& p, const int4 bb, int& tx0, int& tx1, int& ty0, int& ty1) {
    // bb = minx, maxx, miny, maxy
    if (bb.x > bb.y) return false;
    tx0 = bb.x / kTileW - (int)p.tile_x0;
    tx1 = bb.y / kTileW - (int)p.tile_x0;
    ty0 = bb.z / kTileH - (int)p.tile_y0;
    ty1 = bb.w / kTileH - (int)p.tile_y0;
    return true;
}

// Emit the (tile, triangle) pairs of one triangle PER LANE; must be called by all 32 lanes of a converged warp.
// FILL == false: add `weight` to the tile counters; FILL == true: append triangle id `t` to the tile segments.
// Triangles touching few tiles are handled by their own lane; wide ones are spread over the whole warp.
template <bool FILL>
__device__ __forceinline__ void bin_emit(const DrawParams& p, bool valid, uint32_t t, uint32_t weight, int tx0, int tx1, int ty0,
                                         int ty1) {
    constexpr int kSmall = 8;
    const uint32_t lane = threadIdx.x & 31u;
    const int w = tx1 - tx0 + 1;
    const int n = valid ? w * (ty1 - ty0 + 1) : 0;
    auto emit = [&](uint32_t tile, uint32_t tri, uint32_t wgt) {
        if (!FILL) {
            atomicAdd(&p.tile_counts[tile], wgt);
        } else {
            uint32_t pos = p.tile_offsets[tile] + atomicAdd(&p.tile_cursor[tile], 1u);
            if (pos < p.bin_capacity) p.bin_entries[pos] = tri;
        }
    };
    // small triangles: lane-per-triangle, but lanes that hit the same tile in the same step share ONE atomic
    // (neighbouring triangles of a mesh mostly fall into the same tile, which would otherwise serialise at the L2)
    const int n_small = n <= kSmall ? n : 0;
    const int steps = __reduce_max_sync(0xffffffffu, n_small);
    for (int k = 0; k < steps; ++k) {
        const bool act = k < n_small;
        const uint32_t tile = act ? (uint32_t)(ty0 + k / w) * p.tiles_x + (uint32_t)(tx0 + k % w) : (0xffffffffu - lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, tile);
        const int leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (act && (int)lane == leader) {
            if (!FILL) atomicAdd(&p.tile_counts[tile], weight * (uint32_t)__popc(peers));
            else base = p.tile_offsets[tile] + atomicAdd(&p.tile_cursor[tile], (uint32_t)__popc(peers));
        }
        if (FILL) {
            base = __shfl_sync(0xffffffffu, base, leader);
            if (act && base + rank < p.bin_capacity) p.bin_entries[base + rank] = t;
        }
    }
    uint32_t big = __ballot_sync(0xffffffffu, n > kSmall);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int s_tx0 = __shfl_sync(0xffffffffu, tx0, src), s_ty0 = __shfl_sync(0xffffffffu, ty0, src);
        const int s_w = __shfl_sync(0xffffffffu, w, src), s_n = __shfl_sync(0xffffffffu, n, src);
        const uint32_t s_t = __shfl_sync(0xffffffffu, t, src), s_wgt = __shfl_sync(0xffffffffu, weight, src);
        for (int k = (int)lane; k < s_n; k += 32) emit((uint32_t)(s_ty0 + k / s_w) * p.tiles_x + (uint32_t)(s_tx0 + k % s_w), s_t, s_wgt);
    }
}

// bit 3 of TriRec::bias_bits / bit 11 of RowEdge::xw: the interpolated depth of every covered fragment is certainly
// inside the depth-clip range, so the per-fragment z clip (R3) can be skipped.  Inside the triangle the barycentrics are in
// [0,1] and the R6 expression is within a few ulp of a convex combination of the vertex depths: either all three are
// exactly 0 (then every product and sum is exactly 0), or they keep a margin of 2^-16 from both ends.
__device__ __forceinline__ bool depth_clip_cannot_reject(const DrawParams& p, float z0, float z1, float z2) {
    const float zlo = fminf(z0, fminf(z1, z2)), zhi = fmaxf(z0, fmaxf(z1, z2));
    const float margin = 1.52587890625e-05f;
    const bool all_zero = zlo == 0.0f && zhi == 0.0f && p.zclip_lo <= 0.0f && p.zclip_hi >= 0.0f;
    const bool inside = zlo >= p.zclip_lo + margin && zhi <= p.zclip_hi - margin && fabsf(zlo) <= 1.0f && fabsf(zhi) <= 1.0f;
    return all_zero || inside;
}

// Triangle t of an instance -> its compact record (R1-R4: index + position fetch, passthrough VS, viewport, snap, cull)
// and its scissored pixel box.  false: the triangle produces no fragments (box left empty; the record's colours are
// not fetched — on a band of a split frame that is most of the scene).
// ROWS_FIRST: y and w of the three vertices decide first whether the triangle can touch the clip rectangle's rows at all; on a
// band of a split frame most of the scene ends there, before x, z and the colours are fetched.  (Same tests, same
// arithmetic, another order; a draw whose clip rectangle keeps most rows takes the single pass, which is shorter.)
// vertex k of triangle t of an instance: its index into the vertex buffer (negative or huge: out of range)
__device__ __forceinline__ long long triangle_vertex(const DrawParams& p, uint32_t t, int k) {
    // Vulkan 1.1 "Primitive Topologies": list {3t, 3t+1, 3t+2}; strip {t, t+(1+t%2), t+(2-t%2)}; fan {t+1, t+2, 0}
    // points {t}; line list {2t, 2t+1}; line strip {t, t+1}
    unsigned long long rel;
    if (p.st.topology == KZN_TOPOLOGY_TRIANGLE_STRIP) rel = (unsigned long long)t + (k == 0 ? 0u : k == 1 ? 1u + (t & 1u) : 2u - (t & 1u));
    else if (p.st.topology == KZN_TOPOLOGY_TRIANGLE_FAN) rel = k == 2 ? 0ull : (unsigned long long)t + 1u + k;
    else if (p.st.topology == KZN_TOPOLOGY_POINT_LIST) rel = t;
    else if (p.st.topology == KZN_TOPOLOGY_LINE_LIST) rel = 2ull * t + k;
    else if (p.st.topology == KZN_TOPOLOGY_LINE_STRIP) rel = (unsigned long long)t + k;
    else rel = 3ull * t + k;
    unsigned long long i = (unsigned long long)p.first + rel;
    if (p.index_base) {
        unsigned long long raw = p.index_type == KZN_INDEX_TYPE_UINT16
                                     ? (unsigned long long)__ldg(reinterpret_cast<const uint16_t*>(p.index_base) + i)
                                     : (unsigned long long)__ldg(reinterpret_cast<const uint32_t*>(p.index_base) + i);
        return (long long)raw + p.vertex_offset;
    }
    return (long long)i;
}
// where its position is: the robust fetch reads (0, 0, 0, 1) out of range, from a vertex of those values in the scratch
// (p.null_vertex), so that fetching needs no branch
__device__ __forceinline__ const float* vertex_position(const DrawParams& p, long long vidx) {
    return (unsigned long long)vidx < p.pos_vertices ? reinterpret_cast<const float*>(p.vertex_base + (unsigned long long)vidx * p.stride + p.pos_off) : p.null_vertex;
}

// `slow`: the triangle has a vertex behind w = 0 or outside the guard band (or a coordinate that is not finite): it is
// clipped (clip_primitive) or discarded, R3 — false is returned and nothing else is decided here.
// the snapped window y of one vertex (R1, R2) and 1 / w; y_ok goes false for a vertex behind w = 0 or outside the guard band (R3)
__device__ __forceinline__ int32_t vertex_row(const DrawParams& p, const float* pos, float& iw, bool& y_ok) {
    const float y = __ldg(pos + 1);
    iw = 1.0f;  // positions without a w component: w = 1, and 1.0f / 1.0f needs no division
    if (p.pos_comps == 4) {
        const float w = __ldg(pos + 3);
        iw = 1.0f / w;
        if (!(w > 0.0f) || !finite_f(iw)) y_ok = false;  // R3
    }
    const float yf = fmaf(p.hh, y * iw, p.oy);
    if (!(fabsf(yf) <= kGuardBand)) y_ok = false;  // (false for NaN and infinity too)
    return y_ok ? __float2int_rn(yf * 16.0f) : 0;  // R2
}

// true: triangle t certainly produces no fragments because its rows miss the clip rectangle — exactly the triangles
// make_compact_t<true> ends at its first test (not `slow`, box empty); same arithmetic, nothing else fetched
__device__ __forceinline__ bool rows_miss(const DrawParams& p, uint32_t t) {
    bool y_ok = true;
    int32_t Y[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float iw;
        Y[k] = vertex_row(p, vertex_position(p, triangle_vertex(p, t, k)), iw, y_ok);
    }
    if (!y_ok) return false;
    const int32_t minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
    return max((minY - 8 + 15) >> 4, p.cy0) > min((maxY - 8) >> 4, p.cy1);
}

template <bool ROWS_FIRST>
__device__ __forceinline__ bool make_compact_t(const DrawParams& p, uint32_t t, TriRecC& r, int4& box, bool& slow) {
    box = make_int4(1, 0, 1, 0);  // minx maxx miny maxy
    slow = false;
    bool y_ok = true, x_ok = true;
    long long vidx[3];
    const float* pos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vidx[k] = triangle_vertex(p, t, k);
        pos[k] = vertex_position(p, vidx[k]);
        float x = 0.0f, z = 0.0f;
        if (!ROWS_FIRST) {
            x = __ldg(pos[k]);
            z = __ldg(pos[k] + 2);
        }
        float iw;
        const int32_t Yk = vertex_row(p, pos[k], iw, y_ok);
        if (!ROWS_FIRST) {
            const float xf = fmaf(p.hw, x * iw, p.ox);
            const float zf = fmaf(p.dz, z * iw, p.zmin_vp);
            if (!(fabsf(xf) <= kGuardBand) || !finite_f(zf)) x_ok = false;
            r.X[k] = y_ok && x_ok ? __float2int_rn(xf * 16.0f) : 0;  // R2
            r.z[k] = zf;
        }
        r.Y[k] = Yk;
        r.iw[k] = iw;
    }
    bool ok = true;
    if (y_ok) {  // (R3: rows that miss the clip rectangle end the triangle before anything else is looked at)
        const int32_t minY = min(r.Y[0], min(r.Y[1], r.Y[2])), maxY = max(r.Y[0], max(r.Y[1], r.Y[2]));
        // pixels whose sample 16p + 8 lies in [min, max]; the arithmetic shift is a floor division
        box.z = max((minY - 8 + 15) >> 4, p.cy0);
        box.w = min((maxY - 8) >> 4, p.cy1);
        if (box.z > box.w) ok = false;
    }
    if (ROWS_FIRST) {
        if (!ok || !y_ok) {
            box = make_int4(1, 0, 1, 0);
            slow = ok && !y_ok;
            return false;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float x = __ldg(pos[k]), z = __ldg(pos[k] + 2);
            const float xf = fmaf(p.hw, x * r.iw[k], p.ox);
            const float zf = fmaf(p.dz, z * r.iw[k], p.zmin_vp);
            if (!(fabsf(xf) <= kGuardBand) || !finite_f(zf)) x_ok = false;
            r.X[k] = x_ok ? __float2int_rn(xf * 16.0f) : 0;  // R2
            r.z[k] = zf;
        }
    }
    if (ok && !(y_ok && x_ok)) {
        slow = true;
        ok = false;
    }
    const long long two_a = (long long)(r.X[1] - r.X[0]) * (r.Y[2] - r.Y[0]) - (long long)(r.X[2] - r.X[0]) * (r.Y[1] - r.Y[0]);  // R4
    if (two_a == 0) ok = false;
    const bool front = p.st.front_face == 0 ? (two_a < 0) : (two_a > 0);
    if ((front && (p.st.cull_mode & 1u)) || (!front && (p.st.cull_mode & 2u))) ok = false;
    if (ok) {
        const int32_t minX = min(r.X[0], min(r.X[1], r.X[2])), maxX = max(r.X[0], max(r.X[1], r.X[2]));
        box.x = max((minX - 8 + 15) >> 4, p.cx0);
        box.y = min((maxX - 8) >> 4, p.cx1);
        if (box.x > box.y) ok = false;
    }
    if (!ok) {
        box = make_int4(1, 0, 1, 0);
        return false;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 c = fetch_color(p, vidx[k]);
        r.col[k][0] = c.x * r.iw[k];
        r.col[k][1] = c.y * r.iw[k];
        r.col[k][2] = c.z * r.iw[k];
        r.col[k][3] = c.w * r.iw[k];
    }
    return true;
}
__device__ __forceinline__ bool is_point_or_line(const DrawParams& p) {
    return p.st.topology == KZN_TOPOLOGY_POINT_LIST || p.st.topology == KZN_TOPOLOGY_LINE_LIST || p.st.topology == KZN_TOPOLOGY_LINE_STRIP;
}
// pl == false (a compile-time constant at the call): the caller only ever sees triangles — the fused setup of raster_direct_kernel —
// and carries none of the point / line code (that kernel's speed depends on its size: DESIGN.md "raster_direct_kernel")
__device__ __forceinline__ bool make_compact(const DrawParams& p, uint32_t t, TriRecC& r, int4& box, bool& slow, const bool pl = true) {
    if (pl && is_point_or_line(p)) {  // R9: every point / line becomes two triangles, made where clipped polygons are made
        box = make_int4(1, 0, 1, 0);
        slow = true;
        return false;
    }
    return p.rows_first ? make_compact_t<true>(p, t, r, box, slow) : make_compact_t<false>(p, t, r, box, slow);
}

// ---- clipping (R3, oracle/raster.c clip_triangle): the rare triangles make_compact calls `slow` ----------------------------

__device__ __forceinline__ int4 record_box(const DrawParams& p, const TriRecC& c);

__device__ __forceinline__ float clip_distance(const DrawParams& p, const float* c, int plane) {
    if (plane == 0) return c[3] - kClipW;
    const float gw = kClipGuard * c[3];
    if (plane <= 2) {
        const float wx = fmaf(p.hw, c[0], p.ox * c[3]);
        return plane == 1 ? gw - wx : gw + wx;
    }
    const float wy = fmaf(p.hh, c[1], p.oy * c[3]);
    return plane == 3 ? gw - wy : gw + wy;
}

// Triangle t of an instance, clipped: `poly` receives the polygon's vertices (x y z w r g b a each), the result is their
// number (0, 3..8) or -1 when a clip coordinate is not finite (the triangle is discarded).
// R9 (oracle/raster.c expand_point_or_line): a clip-space vertex (x y z w r g b a) in window form {X, Y, z_window, 1/w, r, g, b, a}
// — snapped coordinates as floats (exact up to 2^23); false: behind w, outside the guard band or not finite
__device__ __forceinline__ bool to_window(const DrawParams& p, const float* v, float* out) {
    if (!(v[3] > 0.0f)) return false;
    const float iw = 1.0f / v[3];
    const float xf = fmaf(p.hw, v[0] * iw, p.ox);
    const float yf = fmaf(p.hh, v[1] * iw, p.oy);
    const float zf = fmaf(p.dz, v[2] * iw, p.zmin_vp);
    if (!finite_f(xf) || !finite_f(yf) || !finite_f(zf) || !finite_f(iw) || fabsf(xf) > kGuardBand || fabsf(yf) > kGuardBand) return false;
    out[0] = (float)__float2int_rn(xf * 16.0f);
    out[1] = (float)__float2int_rn(yf * 16.0f);
    out[2] = zf;
    out[3] = iw;
#pragma unroll
    for (int k = 4; k < 8; ++k) out[k] = v[k];
    return true;
}
// point / line t of an instance -> the four corners of its square / parallelogram, in window form, in poly[0..3]; the result is
// 4, 0 (nothing to draw) or -1 (a clip coordinate is not finite)
__device__ __forceinline__ int expand_point_or_line(const DrawParams& p, uint32_t t, float (*poly)[8]) {
    const int n = p.st.topology == KZN_TOPOLOGY_POINT_LIST ? 1 : 2;
    float e[2][8];
    for (int i = 0; i < n; ++i) {
        const long long vidx = triangle_vertex(p, t, i);
        const float* pos = vertex_position(p, vidx);
        const float x = __ldg(pos), y = __ldg(pos + 1), z = __ldg(pos + 2), w = p.pos_comps == 4 ? __ldg(pos + 3) : 1.0f;
        if (!finite_f(x) || !finite_f(y) || !finite_f(z) || !finite_f(w)) return -1;
        const float4 c = fetch_color(p, vidx);
        e[i][0] = x; e[i][1] = y; e[i][2] = z; e[i][3] = w;
        e[i][4] = c.x; e[i][5] = c.y; e[i][6] = c.z; e[i][7] = c.w;
    }
    float w0[8], w1[8];
    if (n == 1) {
        for (int plane = 0; plane < 5; ++plane)
            if (!(clip_distance(p, e[0], plane) >= 0.0f)) return 0;
        if (!to_window(p, e[0], w0)) return 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int c = 0; c < 8; ++c) poly[k][c] = w0[c];
            poly[k][0] = w0[0] + ((k == 1 || k == 2) ? 8.0f : -8.0f);
            poly[k][1] = w0[1] + (k >= 2 ? 8.0f : -8.0f);
        }
        return 4;
    }
    for (int plane = 0; plane < 5; ++plane) {  // the segment is cut plane by plane, from the inside end (R3's arithmetic)
        const float da = clip_distance(p, e[0], plane), db = clip_distance(p, e[1], plane);
        const bool in_a = da >= 0.0f, in_b = db >= 0.0f;
        if (!in_a && !in_b) return 0;
        if (in_a != in_b) {
            const float* I = in_a ? e[0] : e[1];
            const float* O = in_a ? e[1] : e[0];
            const float di = in_a ? da : db, d_o = in_a ? db : da;
            const float tt = __fdiv_rn(di, di - d_o);
            float cut[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) cut[k] = __fmaf_rn(tt, O[k] - I[k], I[k]);
            if (plane == 0) cut[3] = kClipW;
            float* dst = in_a ? e[1] : e[0];
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[k] = cut[k];
        }
    }
    if (!to_window(p, e[0], w0) || !to_window(p, e[1], w1)) return 0;
    const bool x_major = fabsf(w1[0] - w0[0]) >= fabsf(w1[1] - w0[1]);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        poly[0][c] = w0[c]; poly[1][c] = w1[c]; poly[2][c] = w1[c]; poly[3][c] = w0[c];
    }
    const int ax = x_major ? 1 : 0;
    poly[0][ax] -= 8.0f; poly[1][ax] -= 8.0f; poly[2][ax] += 8.0f; poly[3][ax] += 8.0f;
    return 4;
}

__device__ __forceinline__ int clip_primitive(const DrawParams& p, uint32_t t, float (*poly)[8], const bool pl = true) {
    if (pl && is_point_or_line(p)) return expand_point_or_line(p, t, poly);
    float other[8][8];
    float(*src)[8] = other;
    float(*dst)[8] = poly;  // five planes: other -> poly -> other -> poly -> other -> poly
    for (int k = 0; k < 3; ++k) {
        const long long vidx = triangle_vertex(p, t, k);
        const float* pos = vertex_position(p, vidx);
        const float x = __ldg(pos), y = __ldg(pos + 1), z = __ldg(pos + 2), w = p.pos_comps == 4 ? __ldg(pos + 3) : 1.0f;
        if (!finite_f(x) || !finite_f(y) || !finite_f(z) || !finite_f(w)) return -1;
        const float4 c = fetch_color(p, vidx);
        src[k][0] = x; src[k][1] = y; src[k][2] = z; src[k][3] = w;
        src[k][4] = c.x; src[k][5] = c.y; src[k][6] = c.z; src[k][7] = c.w;
    }
    int n = 3;
    for (int plane = 0; plane < 5; ++plane) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const float* cur = src[i];
            const float* nxt = src[i + 1 == n ? 0 : i + 1];
            const float dc = clip_distance(p, cur, plane), dn = clip_distance(p, nxt, plane);
            const bool in_c = dc >= 0.0f, in_n = dn >= 0.0f;
            if (in_c) {
#pragma unroll
                for (int k = 0; k < 8; ++k) dst[m][k] = cur[k];
                ++m;
            }
            if (in_c != in_n) {  // the edge is cut from its inside end to its outside end, whichever way the polygon runs
                const float* I = in_c ? cur : nxt;
                const float* O = in_c ? nxt : cur;
                const float di = in_c ? dc : dn, d_o = in_c ? dn : dc;
                const float tt = __fdiv_rn(di, di - d_o);
#pragma unroll
                for (int k = 0; k < 8; ++k) dst[m][k] = __fmaf_rn(tt, O[k] - I[k], I[k]);
                if (plane == 0) dst[m][3] = kClipW;
                ++m;
            }
        }
        float(*tmp)[8] = src;
        src = dst;
        dst = tmp;
        n = m;
        if (n < 3) return 0;
    }
    return n;  // five swaps: the result is in `poly`
}

// One triangle of a clipped polygon -> its compact record and scissored pixel box, by the rules of make_compact_t (R1, R2,
// R4; oracle/raster.c setup_triangle); false: it produces no fragments.
__device__ __forceinline__ bool make_compact_from(const DrawParams& p, const float* v0, const float* v1, const float* v2, TriRecC& r, int4& box,
                                                  const bool pl = true) {
    box = make_int4(1, 0, 1, 0);
    if (pl && is_point_or_line(p)) {  // R9: corners in window form {X, Y, z, 1/w, colour}; no face culling
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float* v = k == 0 ? v0 : k == 1 ? v1 : v2;
            r.X[k] = (int32_t)v[0];
            r.Y[k] = (int32_t)v[1];
            r.z[k] = v[2];
            r.iw[k] = v[3];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) r.col[k][ch] = v[4 + ch] * v[3];
        }
        const long long area2 = (long long)(r.X[1] - r.X[0]) * (r.Y[2] - r.Y[0]) - (long long)(r.X[2] - r.X[0]) * (r.Y[1] - r.Y[0]);
        if (area2 == 0) return false;
        const int4 b = record_box(p, r);
        if (b.x > b.y || b.z > b.w) return false;
        box = b;
        return true;
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* v = k == 0 ? v0 : k == 1 ? v1 : v2;
        if (!(v[3] > 0.0f)) ok = false;
        const float iw = 1.0f / v[3];
        const float xf = fmaf(p.hw, v[0] * iw, p.ox);
        const float yf = fmaf(p.hh, v[1] * iw, p.oy);
        const float zf = fmaf(p.dz, v[2] * iw, p.zmin_vp);
        if (!finite_f(xf) || !finite_f(yf) || !finite_f(zf) || !finite_f(iw) || fabsf(xf) > kGuardBand || fabsf(yf) > kGuardBand) ok = false;
        r.X[k] = ok ? __float2int_rn(xf * 16.0f) : 0;
        r.Y[k] = ok ? __float2int_rn(yf * 16.0f) : 0;
        r.z[k] = zf;
        r.iw[k] = iw;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) r.col[k][ch] = v[4 + ch] * iw;
    }
    if (!ok) return false;
    const long long two_a = (long long)(r.X[1] - r.X[0]) * (r.Y[2] - r.Y[0]) - (long long)(r.X[2] - r.X[0]) * (r.Y[1] - r.Y[0]);  // R4
    if (two_a == 0) return false;
    const bool front = p.st.front_face == 0 ? (two_a < 0) : (two_a > 0);
    if ((front && (p.st.cull_mode & 1u)) || (!front && (p.st.cull_mode & 2u))) return false;
    const int4 b = record_box(p, r);
    if (b.x > b.y || b.z > b.w) return false;
    box = b;
    return true;
}

__device__ __forceinline__ uint2 pack_box(const int4 box) {
    return make_uint2((uint32_t)box.x | ((uint32_t)box.y << 16), (uint32_t)box.z | ((uint32_t)box.w << 16));
}
__device__ __forceinline__ int4 unpack_box(const uint2 b) {  // minx maxx miny maxy
    return make_int4((int)(b.x & 0xffffu), (int)(b.x >> 16), (int)(b.y & 0xffffu), (int)(b.y >> 16));
}

__device__ __forceinline__ void clip_statistics(const DrawParams& p, int m, unsigned long long n_inst, const bool pl = true) {
    if (!p.stats) return;
    if (pl && is_point_or_line(p) && m >= 0) return;  // (their two triangles are the primitive itself, not the clipper's work)
    if (m < 0) {
        atomicAdd(&p.stats[10], n_inst);
    } else {
        atomicAdd(&p.stats[8], n_inst);
        if (m >= 3) atomicAdd(&p.stats[9], (unsigned long long)(m - 2) * n_inst);
    }
}

// compact -> expanded record (R5: sign-normalised integer edge equations, top-left bias, 1/2A); `c` must be the record of a
// triangle that produces fragments (2A != 0) and `box` its scissored pixel box
__device__ __forceinline__ void expand_record(const DrawParams& p, const TriRecC& c, const int4 box, TriRec& r) {
    r.minx = box.x;
    r.maxx = box.y;
    r.miny = box.z;
    r.maxy = box.w;
    const long long two_a = (long long)(c.X[1] - c.X[0]) * (c.Y[2] - c.Y[0]) - (long long)(c.X[2] - c.X[0]) * (c.Y[1] - c.Y[0]);
    const long long s = two_a < 0 ? -1 : 1;
    r.bias_bits = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ia = (i + 1) % 3, ib = (i + 2) % 3;
        const long long dx = c.X[ib] - c.X[ia], dy = c.Y[ib] - c.Y[ia];
        const long long a = -dy * s, b = dx * s;
        r.a[i] = (int32_t)a;
        r.b[i] = (int32_t)b;
        r.c[i] = s * (dy * c.X[ia] - dx * c.Y[ia]);
        if (a > 0 || (a == 0 && b > 0)) r.bias_bits |= 1u << i;
    }
    r.inv2a = 1.0f / (float)(two_a * s);
    if (depth_clip_cannot_reject(p, c.z[0], c.z[1], c.z[2])) r.bias_bits |= 8u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.z[k] = c.z[k];
        r.iw[k] = c.iw[k];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) r.col[k][ch] = c.col[k][ch];
    }
}
// the scissored pixel box of a compact record (what make_compact returned with it)
__device__ __forceinline__ int4 record_box(const DrawParams& p, const TriRecC& c) {
    const int32_t minX = min(c.X[0], min(c.X[1], c.X[2])), maxX = max(c.X[0], max(c.X[1], c.X[2]));
    const int32_t minY = min(c.Y[0], min(c.Y[1], c.Y[2])), maxY = max(c.Y[0], max(c.Y[1], c.Y[2]));
    return make_int4(max((minX - 8 + 15) >> 4, p.cx0), min((maxX - 8) >> 4, p.cx1), max((minY - 8 + 15) >> 4, p.cy0),
                     min((maxY - 8) >> 4, p.cy1));
}

// Triangle t of an instance -> its expanded record; false: it produces no fragments (its box is left empty).
__device__ __forceinline__ void empty_record(TriRec& r) {
    memset(&r, 0, sizeof r);
    r.minx = 1;
    r.miny = 1;
}
__device__ __forceinline__ bool make_record(const DrawParams& p, uint32_t t, TriRec& r, bool& slow, const bool pl = true) {
    TriRecC c;
    int4 box;
    const bool ok = make_compact(p, t, c, box, slow, pl);
    if (ok) expand_record(p, c, box, r);
    else empty_record(r);
    return ok;
}
// triangle j of a clipped polygon -> expanded record (direct draws)
__device__ __forceinline__ bool make_record_from(const DrawParams& p, const float (*poly)[8], int j, TriRec& r, const bool pl = true) {
    TriRecC c;
    int4 box;
    const bool ok = make_compact_from(p, poly[0], poly[j + 1], poly[j + 2], c, box, pl);
    if (ok) expand_record(p, c, box, r);
    else empty_record(r);
    return ok;
}

// A clipped primitive of a binned draw (bin_scan_kernel): the first triangle of its polygon takes the primitive's own record, the
// others get extra records after the draw's (allocated from work_counter[4]); the tiles they overlap are counted here
// (setup_kernel counted nothing for the primitive), bin_fill_kernel files them.
__device__ __forceinline__ void count_tiles(const DrawParams& p, const int4 box, uint32_t n_inst) {
    int tx0, tx1, ty0, ty1;
    if (!tile_range(p, box, tx0, tx1, ty0, ty1)) return;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&p.tile_counts[(uint32_t)ty * p.tiles_x + (uint32_t)tx], n_inst);
}
__device__ __forceinline__ void clip_binned(const DrawParams& p, uint32_t t, uint32_t n_inst) {
    float poly[8][8];
    const int m = clip_primitive(p, t, poly);
    clip_statistics(p, m, n_inst);
    bool any_fragments = false;
    if (m >= 3) {
        TriRecC c;
        int4 box;
        if (make_compact_from(p, poly[0], poly[1], poly[2], c, box)) {
            any_fragments = true;
            for (uint32_t inst = 0; inst < n_inst; ++inst) {
                reinterpret_cast<TriRecC*>(p.tris)[(size_t)inst * p.tris_per_instance + t] = c;
                p.boxes[(size_t)inst * p.tris_per_instance + t] = pack_box(box);
            }
            count_tiles(p, box, n_inst);
        }
        const uint32_t n_extra = (uint32_t)m - 3u;
        if (n_extra) {
            const uint32_t k = atomicAdd(p.work_counter + 4, n_extra);
            if (k + n_extra > p.clip_capacity) {  // no room (KznDrawStatistics::clip_overflow): the slots that exist stay unused
                for (uint32_t i = k; i < p.clip_capacity && i < k + n_extra; ++i) p.clip_parent[i] = 0xffffffffu;
                if (p.stats) atomicAdd(&p.stats[11], (unsigned long long)n_inst);
                any_fragments = true;  // (not "culled": lost)
            } else {
                p.clip_link[t] = k | n_extra << 29;
                atomicOr(&p.clip_bits[t >> 5], 1u << (t & 31u));
                TriRecC* recs = reinterpret_cast<TriRecC*>(p.tris) + p.total_tris + k;
                for (uint32_t j = 1; j <= n_extra; ++j) {
                    TriRecC e;
                    int4 eb;
                    if (make_compact_from(p, poly[0], poly[j + 1], poly[j + 2], e, eb)) {
                        any_fragments = true;
                        recs[j - 1] = e;
                        count_tiles(p, eb, n_inst);
                    }
                    p.boxes[p.total_tris + k + j - 1u] = pack_box(eb);
                    p.clip_parent[k + j - 1u] = t << 3 | j;
                }
            }
        }
    }
    if (!any_fragments && p.stats) atomicAdd(&p.stats[1], (unsigned long long)n_inst);
}

// setup of one block of 256 triangles (block `blk` of the draw) by one CTA
template <bool BINNED>
__device__ __forceinline__ void setup_block(const DrawParams& p, const uint32_t blk) {
    constexpr int kQuads = BINNED ? (int)(sizeof(TriRecC) / 16) : (int)(sizeof(TriRec) / 16);
    const uint32_t t_raw = blk * 256u + threadIdx.x;
    const bool in_range = t_raw < p.tris_per_instance;
    const uint32_t t = in_range ? t_raw : 0u;  // out-of-range lanes redo triangle 0 and write nothing (they stay for the warp-wide binning)
    // The CTA's 256 records are contiguous in memory: they are transposed through shared memory and written with
    // coalesced 16-byte stores (a thread writing its own record makes kQuads store instructions of 32 scattered
    // sectors each: the LSU queue was the top stall of this kernel).
    __shared__ uint4 stage[256 * kQuads];
    __shared__ uint8_t keep[256];
    TriRecC c;
    int4 box;
    bool slow;
    const bool ok = make_compact(p, t, c, box, slow);
    slow = slow && in_range;
    const uint32_t n_inst = p.instances;  // instances repeat the same triangles (no per-instance inputs): record i*tris_per_instance + t
    bool any_fragments = ok;  // some triangle of the (clipped) primitive produces fragments
    if (BINNED && slow) p.clip_list[atomicAdd(p.work_counter + 6, 1u)] = t;  // R3: bin_scan_kernel clips this one (nothing is counted for it here)
    if (!BINNED) {
        // direct draws (one CTA): a clipped primitive takes as many consecutive records as its polygon has triangles, so the
        // records are placed by a prefix sum over the CTA whenever any primitive was clipped
        __shared__ uint32_t warp_sums[8];
        if (__syncthreads_or(slow)) {
            float poly[8][8];
            int m = 0;
            if (slow) {
                m = clip_primitive(p, t, poly);
                clip_statistics(p, m, n_inst);
            }
            const uint32_t cnt = !in_range ? 0u : (slow && m > 3) ? (uint32_t)m - 2u : 1u;
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)(threadIdx.x & 31u) >= d) incl += y;
            }
            if ((threadIdx.x & 31u) == 31u) warp_sums[threadIdx.x >> 5] = incl;
            __syncthreads();
            uint32_t before = 0, per_inst = 0;
#pragma unroll
            for (uint32_t w = 0; w < 8u; ++w) {
                if (w < (threadIdx.x >> 5)) before += warp_sums[w];
                per_inst += warp_sums[w];
            }
            const uint32_t at = before + incl - cnt;
            if (threadIdx.x == 0) p.work_counter[3] = per_inst * n_inst;
            if (in_range) {
                for (uint32_t j = 0; j < cnt; ++j) {
                    TriRec r;
                    bool okj = ok;
                    if (slow) {
                        okj = m >= 3 && make_record_from(p, poly, (int)j, r);
                        if (m < 3) empty_record(r);
                    } else if (ok) {
                        expand_record(p, c, box, r);
                    } else {
                        empty_record(r);
                    }
                    any_fragments = any_fragments || okj;
                    for (uint32_t inst = 0; inst < n_inst; ++inst) p.tris[(size_t)inst * per_inst + at + j] = r;
                }
                if (p.stats) {
                    atomicAdd(&p.stats[0], (unsigned long long)n_inst);
                    if (!any_fragments) atomicAdd(&p.stats[1], (unsigned long long)n_inst);
                }
            }
            return;
        }
        if (threadIdx.x == 0) p.work_counter[3] = p.total_tris;
    }
    // binned draws find their triangles through the packed boxes: a triangle that produces no fragments (culled,
    // degenerate, outside the clip rectangle or this GPU's band) needs no record — and a CTA none of whose triangles
    // does (most of them, on a band of a split frame) has nothing to stage
    const bool keep_mine = BINNED ? (in_range && ok) : in_range;
    if (__syncthreads_or(keep_mine)) {
        if (BINNED) {
            const uint4* src = reinterpret_cast<const uint4*>(&c);
#pragma unroll
            for (int q = 0; q < kQuads; ++q) stage[threadIdx.x * kQuads + q] = src[q];
        } else {
            TriRec r;
            if (ok) expand_record(p, c, box, r);
            else empty_record(r);
            const uint4* src = reinterpret_cast<const uint4*>(&r);
#pragma unroll
            for (int q = 0; q < kQuads; ++q) stage[threadIdx.x * kQuads + q] = src[q];
        }
        keep[threadIdx.x] = keep_mine;
        __syncthreads();
        const uint32_t block_first = blk * 256u;
        for (uint32_t inst = 0; inst < n_inst; ++inst) {
            const size_t first_rec = (size_t)inst * p.tris_per_instance + block_first;
            uint4* dst = BINNED ? reinterpret_cast<uint4*>(reinterpret_cast<TriRecC*>(p.tris) + first_rec) : reinterpret_cast<uint4*>(p.tris + first_rec);
            for (uint32_t i = threadIdx.x; i < 256u * kQuads; i += 256u)
                if (keep[i / kQuads]) dst[i] = stage[i];
        }
    }
    if (in_range) {
        p.boxes[t] = pack_box(box);
        for (uint32_t inst = 1; inst < n_inst; ++inst) p.boxes[(size_t)inst * p.tris_per_instance + t] = pack_box(box);
        if (p.stats) {
            atomicAdd(&p.stats[0], (unsigned long long)n_inst);
            if (!any_fragments && !(BINNED && slow)) atomicAdd(&p.stats[1], (unsigned long long)n_inst);
        }
    }
    if (BINNED) {  // first binning pass fused here: count the tiles each triangle overlaps
        int tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0;
        bool valid = in_range && tile_range(p, box, tx0, tx1, ty0, ty1);
        if (__any_sync(0xffffffffu, valid)) bin_emit<false>(p, valid, t, n_inst, tx0, tx1, ty0, ty1);
    }
}

template <bool BINNED>
__global__ void __launch_bounds__(256) setup_kernel(const __grid_constant__ DrawParams p) {
    kzn_pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) p.work_counter[BINNED ? 2 : 0] = 0u;  // the work dispenser of raster_kernel / raster_direct_kernel
    uint32_t blk = blockIdx.x;
    if (BINNED && p.live_blocks) {  // band_prefilter_kernel ran: only the listed blocks have a triangle that can touch the band's rows
        const uint32_t n_live = p.work_counter[8];
        if (p.stats && blockIdx.x == 0 && threadIdx.x == 0) {  // the triangles of all other blocks: seen, no fragments
            const unsigned long long n = (unsigned long long)(p.tris_per_instance - p.work_counter[9]) * p.instances;
            atomicAdd(&p.stats[0], n);
            atomicAdd(&p.stats[1], n);
        }
        // (a grid-stride loop over the list: the grid is capped, CTAs that only find the list's end cost a launch each —
        // 13 000 of them on a band of the 4M-triangle frame were a third of this kernel's time)
        for (uint32_t i = blockIdx.x; i < n_live; i += gridDim.x) {
            if (i != blockIdx.x) __syncthreads();  // the previous block's staging area is free again
            setup_block<BINNED>(p, p.live_blocks[i]);
        }
        return;
    }
    setup_block<BINNED>(p, blk);
}

// ---- binning --------------------------------------------------------------------------------------------------

#ifndef KZN_FS_TEMPLATE
// second binning pass: one thread per triangle record writes its id into every tile segment it overlaps
__global__ void __launch_bounds__(256) bin_fill_kernel(const __grid_constant__ DrawParams p) {
    kzn_pdl_wait();
    int tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0;
    // (prefiltered draws have one instance and a capped grid: the blocks of setup's list, grid-stride)
    const uint32_t n_mine = p.live_blocks ? p.work_counter[8] : gridDim.x;
    for (uint32_t i = blockIdx.x; i < n_mine; i += gridDim.x) {
        const uint32_t t = (p.live_blocks ? p.live_blocks[i] : i) * 256u + threadIdx.x;
        bool valid = false;
        if (t < p.total_tris) valid = tile_range(p, unpack_box(__ldg(&p.boxes[t])), tx0, tx1, ty0, ty1);
        bin_emit<true>(p, valid, t << 3, 1u, tx0, tx1, ty0, ty1);  // list entries: record << 3 | fan index (0: the primitive's own record)
    }
    // the extra triangles of clipped primitives (R3; rare): entry (instance's record of the primitive) << 3 | fan index, so
    // that sorting the entries puts them at the primitive's place
    const uint32_t n_extra = min(p.work_counter[4], p.clip_capacity);
    for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < n_extra; k += gridDim.x * 256u) {
        const uint32_t parent = p.clip_parent[k];
        if (parent == 0xffffffffu) continue;
        if (!tile_range(p, unpack_box(p.boxes[p.total_tris + k]), tx0, tx1, ty0, ty1)) continue;
        const uint32_t n_inst = p.instances;
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                const uint32_t tile = (uint32_t)ty * p.tiles_x + (uint32_t)tx;
                for (uint32_t inst = 0; inst < n_inst; ++inst) {
                    const uint32_t pos = p.tile_offsets[tile] + atomicAdd(&p.tile_cursor[tile], 1u);
                    if (pos < p.bin_capacity) p.bin_entries[pos] = ((inst * p.tris_per_instance + (parent >> 3)) << 3) | (parent & 7u);
                }
            }
    }
}

// Before setup on a band of a split frame: which blocks of 256 triangles have a triangle whose rows can touch the band.
// Two passes.  band_code_kernel, one thread per vertex of the vertex buffer (coalesced): bit 0 — the vertex lies above the
// band's first sample row, bit 1 — below its last (0 for a vertex behind w = 0 or outside the guard band: nothing is known).
// band_prefilter_kernel, one thread per triangle: three indices and three bytes of the code table; a triangle whose vertices
// share a bit cannot touch the band (conservative: setup_kernel makes the exact test again for the blocks that stay).
// Against fetching y at the vertex stride per triangle (every 32-byte sector of the vertex buffer through the L1 three times
// over) the table is 1 byte per vertex and stays in the L1 / L2.
__global__ void __launch_bounds__(256) band_code_kernel(const __grid_constant__ DrawParams p, uint8_t* __restrict__ codes, uint32_t n_codes) {
    kzn_pdl_wait();
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    if (v > n_codes) return;  // codes[n_codes]: the vertex an out-of-range index reads (robust fetch)
    bool y_ok = true;
    float iw;
    const int32_t Y = vertex_row(p, vertex_position(p, v < n_codes ? (long long)v : -1ll), iw, y_ok);
    uint32_t c = 0;
    if (y_ok) c = (((Y - 8) >> 4) < p.cy0 ? 1u : 0u) | (((Y - 8 + 15) >> 4) > p.cy1 ? 2u : 0u);
    codes[v] = (uint8_t)c;
}

__global__ void __launch_bounds__(256) band_prefilter_kernel(const __grid_constant__ DrawParams p, uint8_t* __restrict__ flags,
                                                              uint32_t* __restrict__ list, const uint8_t* __restrict__ codes, uint32_t n_codes) {
    kzn_pdl_wait();
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    bool live = false;
    if (t < p.tris_per_instance) {
        if (codes) {
            uint32_t c = 3u;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const unsigned long long v = (unsigned long long)triangle_vertex(p, t, k);
                c &= __ldg(codes + (v < n_codes ? v : (unsigned long long)n_codes));
            }
            live = c == 0u;
        } else {
            live = !rows_miss(p, t);
        }
    }
    const int any = __syncthreads_or(live);
    if (threadIdx.x == 0) {
        flags[blockIdx.x] = (uint8_t)(any != 0);
        if (any) {
            list[atomicAdd(p.work_counter + 8, 1u)] = blockIdx.x;
            atomicAdd(p.work_counter + 9, min(256u, p.tris_per_instance - blockIdx.x * 256u));
        }
    }
}

// The same for the common case — TRIANGLE_LIST, 32-bit indices, the first index 16-byte aligned, a code table: a thread
// takes four triangles (three 16-byte index loads, twelve code bytes in flight at once), a warp half a block, a CTA of 256
// threads four blocks.  The one-triangle-per-thread form waits for one index -> code round trip at a time (25 us on a band
// of the 4M-triangle frame, reading 50 MB).
__global__ void __launch_bounds__(256) band_prefilter4_kernel(const __grid_constant__ DrawParams p, uint8_t* __restrict__ flags,
                                                               uint32_t* __restrict__ list, const uint8_t* __restrict__ codes, uint32_t n_codes) {
    kzn_pdl_wait();
    __shared__ uint32_t warp_live[8];
    const uint32_t q = blockIdx.x * 256u + threadIdx.x;  // triangles 4q .. 4q + 3
    const uint32_t n = p.tris_per_instance;
    bool live = false;
    if (4u * q < n) {
        const uint4* idx = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(p.index_base) + p.first) + 3ull * q;
        uint32_t v[12];
        if (4u * q + 3u < n) {
            const uint4 a = __ldg(idx), b = __ldg(idx + 1), c = __ldg(idx + 2);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
        } else {  // the draw's last triangles: index by index (nothing beyond the index range is read)
            const uint32_t* i32 = reinterpret_cast<const uint32_t*>(idx);
#pragma unroll
            for (uint32_t k = 0; k < 12u; ++k) v[k] = 12u * q + k < 3u * n ? __ldg(i32 + k) : 0u;
        }
        uint32_t c[12];
#pragma unroll
        for (uint32_t k = 0; k < 12u; ++k) {
            const unsigned long long vi = (unsigned long long)((long long)v[k] + p.vertex_offset);
            c[k] = __ldg(codes + (vi < n_codes ? vi : (unsigned long long)n_codes));
        }
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j)
            if (4u * q + j < n && (c[3 * j] & c[3 * j + 1] & c[3 * j + 2]) == 0u) live = true;
    }
    // a warp holds 128 triangles = half a block: block j of this CTA = warps 2j and 2j + 1
    const uint32_t ballot = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31u) == 0) warp_live[threadIdx.x >> 5] = ballot;
    __syncthreads();
    if (threadIdx.x < 4u) {
        const uint32_t blk = blockIdx.x * 4u + threadIdx.x;
        if (blk * 256u < n) {
            const bool any = (warp_live[2 * threadIdx.x] | warp_live[2 * threadIdx.x + 1]) != 0u;
            flags[blk] = (uint8_t)any;
            if (any) {
                list[atomicAdd(p.work_counter + 8, 1u)] = blk;
                atomicAdd(p.work_counter + 9, min(256u, n - blk * 256u));
            }
        }
    }
}

// Between the two binning passes: (1) the primitives setup_kernel<true> found crossing the eye plane or the guard band
// (work_counter[6] of them, usually none) are clipped, one thread each over the whole grid; (2) the last CTA to finish that
// makes the exclusive scan of tile_counts -> tile_offsets[0..tiles] (tiles <= a few hundred thousand).
__global__ void __launch_bounds__(1024) bin_scan_kernel(const __grid_constant__ DrawParams p) {
    kzn_pdl_wait();
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    __shared__ uint32_t last_cta;
    const uint32_t n_slow = p.work_counter[6];
    for (uint32_t i = blockIdx.x * 1024u + threadIdx.x; i < n_slow; i += gridDim.x * 1024u) clip_binned(p, p.clip_list[i], p.instances);
    if (gridDim.x > 1) {
        __threadfence();  // my counts and records before my ticket
        __syncthreads();
        if (threadIdx.x == 0) {
            last_cta = atomicAdd(p.work_counter + 7, 1u) == gridDim.x - 1u;
            if (last_cta) p.work_counter[7] = 0u;  // for the next draw
        }
        __syncthreads();
        if (!last_cta) return;
        __threadfence();
    }
    uint32_t tiles = p.tiles_x * p.tiles_y;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < tiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < tiles ? __ldcg(&p.tile_counts[i]) : 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x], xs = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, xs, d);
                if (threadIdx.x >= d) xs += y;
            }
            warp_sums[threadIdx.x] = xs - w;  // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
        if (i < tiles) p.tile_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.tile_offsets[tiles] = carry;
        if (p.stats) atomicAdd(&p.stats[3], (unsigned long long)carry);
    }
}

#endif  // !KZN_FS_TEMPLATE

// ---- per-fragment helpers ------------------------------------------------------------------------------------------

__device__ __forceinline__ float clamp01(float f) { return __saturatef(f); }  // NaN -> 0
__device__ __forceinline__ float unorm8_to_float(uint32_t u) { return __fdiv_rn((float)u, 255.0f); }
// E.1: clamp to [0, 1], one multiply by 255, round to nearest even.  The rounding is an fp32 add of 1.5 * 2^23: every fp32 in
// [2^23, 2^24) is an integer, so for t in [0, 255] the sum is 1.5 * 2^23 + rint(t), exactly what cvt.rni gives, and the integer
// is the low byte of the sum's bit pattern (an FADD on the fp32 pipe where F2I goes through the conversion unit).
// float_to_unorm8_bits leaves the other bytes of the word as they come (0x4B40 00 xx).
__device__ __forceinline__ uint32_t float_to_unorm8_bits(float f) { return __float_as_uint(__fadd_rn(__fmul_rn(clamp01(f), 255.0f), 12582912.0f)); }
__device__ __forceinline__ uint32_t float_to_unorm8(float f) { return float_to_unorm8_bits(f) & 0xffu; }

// VkCompareOp 0..7 = {NEVER, LESS, EQUAL, LESS_OR_EQUAL, GREATER, NOT_EQUAL, GREATER_OR_EQUAL, ALWAYS}: bit 0 = "less passes",
// bit 1 = "equal passes", bit 2 = "greater passes"; an unordered pair (NaN in the depth buffer) passes NOT_EQUAL and ALWAYS
// only.  Branch-free form of the oracle's switch (oracle/raster.c depth_compare), same truth table.
__device__ __forceinline__ bool depth_compare(uint32_t op, float z, float zs) {
    const bool lt = z < zs, eq = z == zs, gt = z > zs;
    const bool un = !(lt || eq || gt);
    return ((op & 1u) && lt) || ((op & 2u) && eq) || ((op & 4u) && gt) || (un && (op == 5u || op == 7u));
}

// VkBlendFactor 0..9 = {ZERO, ONE, SRC_COLOR, 1-SRC_COLOR, DST_COLOR, 1-DST_COLOR, SRC_ALPHA, 1-SRC_ALPHA, DST_ALPHA,
// 1-DST_ALPHA}: bit 0 selects "one minus", the upper bits the operand — written as selects (the state is warp-uniform but a
// switch per channel per fragment costs a branch chain).  Same values as oracle/raster.c: ONE is 1.0f - 0.0f.
__device__ __forceinline__ float blend_factor(uint32_t f, int ch, const float s[4], const float d[4]) {
    const uint32_t sel = f >> 1;
    float x = 0.0f;
    x = sel == 1u ? s[ch] : x;
    x = sel == 2u ? d[ch] : x;
    x = sel == 3u ? s[3] : x;
    x = sel == 4u ? d[3] : x;
    return (f & 1u) ? 1.0f - x : x;
}

// VkBlendOp 0..4 = {ADD, SUBTRACT, REVERSE_SUBTRACT, MIN, MAX}
__device__ __forceinline__ float blend_op(uint32_t op, float cs, float sf, float cd, float df) {
    const float t0 = cs * sf, t1 = cd * df;
    float r = t0 + t1;
    r = op == 1u ? t0 - t1 : r;
    r = op == 2u ? t1 - t0 : r;
    r = op == 3u ? (cs < cd ? cs : cd) : r;
    r = op == 4u ? (cs > cd ? cs : cd) : r;
    return r;
}

__device__ __forceinline__ void fs_procedural_orbit(int px, int py, float inv_w, float inv_h, float rgba[4]) {
    float u = ((float)px + 0.5f) * inv_w;
    float v = ((float)py + 0.5f) * inv_h;
    float cx = fmaf(u, 3.5f, -2.5f);
    float cy = fmaf(v, 2.0f, -1.0f);
    float zx = 0.0f, zy = 0.0f;
#pragma unroll 8
    for (uint32_t k = 0; k < KZN_FS_ORBIT_ITERATIONS; ++k) {
        float t = fmaf(zx, zx, cx);
        float nzx = fmaf(-zy, zy, t);
        float nzy = fmaf(zx + zx, zy, cy);
        zx = nzx;
        zy = nzy;
    }
    float m = fmaf(zx, zx, zy * zy);
    rgba[0] = 0.25f * m;
    rgba[1] = fmaf(zx, 0.5f, 0.5f);
    rgba[2] = fmaf(zy, 0.5f, 0.5f);
    rgba[3] = 1.0f;
}

__device__ __forceinline__ unsigned long long color_addr(const DrawParams& p, int px, int py) {
    if (p.color_tiled) {
        unsigned long long tile = (unsigned long long)(py >> 4) * p.color_tiles_x + (px >> 4);
        return p.color_base + ((tile << 8) + ((py & 15) << 4) + (px & 15)) * 4ull;
    }
    return p.color_base + ((unsigned long long)py * p.fb_w + px) * 4ull;
}

// ---- raster ------------------------------------------------------------------------------------------------------------

// 4 flag bits <-> 4 flag bytes (0 / 1) of a lane's pixel quad
__device__ __forceinline__ uint32_t bytes_to_bits(uint32_t w) { return (w & 1u) | ((w >> 7) & 2u) | ((w >> 14) & 4u) | ((w >> 21) & 8u); }

struct BrickState {  // private to one warp of raster_kernel: the brick it is working on
    TriRecC recs[32];     // the records of the chunk being rasterised
    BrickEdge edges[32];  // what lane j derived for triangle j of the chunk
    uint32_t pend[160];   // ids of the triangles whose pixel box meets the brick, waiting for a full chunk (32 + one filter step of 128)
    uint32_t queue[64];   // covered fragments waiting for a full shading batch: chunk slot << 8 | y << 4 | x
    uint32_t color[kBrick * kBrick];  // the brick's colour and depth live here for the whole triangle list
    float depth[kBrick * kBrick];
    uint8_t color_dirty[kBrick * kBrick];  // one byte per pixel (plain stores: a brick belongs to one warp)
    uint8_t depth_dirty[kBrick * kBrick];
};
struct RasterShared {
    BrickState brick[kRasterWarps];
    float unorm_lut[256];  // u8 / 255.0f, the exact quotients (E.1), so blending needs no division per fragment
};
struct SortShared {  // tile_sort_cta_kernel
    uint32_t list[kSortCap];
    uint32_t second[kSortCap];
    uint32_t table[256 * kSortWarps];  // [digit][warp]
    uint32_t warp_counts[kSortWarps];
};

// Stable LSD radix sort (8-bit digits) of the n <= kSortCap triangle ids in sh.list, by the whole CTA.  The binning
// atomics deliver each tile's list in arbitrary order; primitive order (R7) is restored here.  Only the bits in which the
// ids of this tile differ are sorted (ids of one tile come from a narrow range of a mesh: typically 2 passes).
// Warp w owns a contiguous slice of the list and walks it 32 ids at a time; match.any ranks equal digits inside a batch
// in lane order, a [digit][warp] table of running offsets orders batches and warps, so every pass is stable.
__device__ __forceinline__ const uint32_t* sort_tile_list(SortShared& sh, uint32_t n) {
    constexpr uint32_t kFull = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t* a = sh.list;
    uint32_t* b = sh.second;
    uint32_t* table = sh.table;  // [256 digits][kSortWarps warps]
    // range of the ids
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (uint32_t i = tid; i < n; i += kSortThreads) {
        const uint32_t v = a[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(kFull, lo);
    hi = __reduce_max_sync(kFull, hi);
    if (lane == 0) {
        sh.warp_counts[warp] = lo;
        table[warp] = hi;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
        lo = min(lo, sh.warp_counts[w]);
        hi = max(hi, table[w]);
    }
    lo &= ~7u;  // entries are record << 3 | fan index: sorted by record (the triangles of one primitive do not overlap)
    __syncthreads();
    const uint32_t span = hi - lo;
    const uint32_t bits = span ? 32u - (uint32_t)__clz(span) : 0u;
    const uint32_t per = ((n + kSortThreads - 1) / kSortThreads) * 32u;  // ids per warp, whole batches
    const uint32_t w_beg = min(warp * per, n), w_end = min(w_beg + per, n);
    for (uint32_t shift = 3; shift < bits; shift += 8) {
        for (uint32_t i = tid; i < 256u * kSortWarps; i += kSortThreads) table[i] = 0;
        __syncthreads();
        // digit counts per warp
        for (uint32_t base = w_beg; base < w_end; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < w_end;
            const uint32_t d = valid ? ((a[i] - lo) >> shift) & 255u : 256u + lane;
            const uint32_t peers = __match_any_sync(kFull, d);
            if (valid && (int)lane == __ffs(peers) - 1) table[d * kSortWarps + warp] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // exclusive scan in (digit, warp) order: thread t owns digit t
        {
            uint32_t v[kSortWarps], total = 0;
#pragma unroll
            for (int k = 0; k < kSortWarps; ++k) {
                v[k] = total;
                total += table[tid * kSortWarps + k];
            }
            uint32_t incl = total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, d);
                if ((int)lane >= d) incl += y;
            }
            if (lane == 31) sh.warp_counts[warp] = incl;
            __syncthreads();
            uint32_t before = incl - total;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w)
                if ((uint32_t)w < warp) before += sh.warp_counts[w];
#pragma unroll
            for (int k = 0; k < kSortWarps; ++k) table[tid * kSortWarps + k] = before + v[k];
        }
        __syncthreads();
        // stable scatter
        for (uint32_t base = w_beg; base < w_end; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < w_end;
            const uint32_t v = valid ? a[i] : 0u;
            const uint32_t d = valid ? ((v - lo) >> shift) & 255u : 256u + lane;
            const uint32_t peers = __match_any_sync(kFull, d);
            const int leader = __ffs(peers) - 1;
            uint32_t start = 0;
            if (valid && (int)lane == leader) {
                start = table[d * kSortWarps + warp];
                table[d * kSortWarps + warp] = start + __popc(peers);
            }
            start = __shfl_sync(kFull, start, leader);
            if (valid) b[start + __popc(peers & ((1u << lane) - 1u))] = v;
            __syncwarp();
        }
        __syncthreads();
        uint32_t* t = a;
        a = b;
        b = t;
    }
    return a;  // the buffer holding the sorted ids
}

// VkLogicOp 0..15 on the UNORM8 bit patterns, all four channels at once (oracle/raster.c logic_op); the state is warp-uniform
__device__ __forceinline__ uint32_t logic_op(uint32_t op, uint32_t s, uint32_t d) {
    switch (op) {
    case 0: return 0u;
    case 1: return s & d;
    case 2: return s & ~d;
    case 3: return s;
    case 4: return ~s & d;
    case 5: return d;
    case 6: return s ^ d;
    case 7: return s | d;
    case 8: return ~(s | d);
    case 9: return ~(s ^ d);
    case 10: return ~d;
    case 11: return s | ~d;
    case 12: return ~s;
    case 13: return ~s | d;
    case 14: return ~(s & d);
    default: return 0xffffffffu;
    }
}

// blend with the destination held as UNORM8 in shared memory; same operation order as oracle/raster.c output_merge
template <int NEED_DST>
__device__ __forceinline__ uint32_t output_merge_lut(const KznGraphicsPipelineState& st, const float* lut, const float src[4],
                                                     uint32_t dst_word) {
    float s[4], o[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) s[ch] = clamp01(src[ch]);
    if (!NEED_DST) {  // no blending, all channels written
        // four low bytes -> one word: two byte permutes for the pairs, one for the word
        const uint32_t rg = __byte_perm(float_to_unorm8_bits(s[0]), float_to_unorm8_bits(s[1]), 0x4440);
        const uint32_t ba = __byte_perm(float_to_unorm8_bits(s[2]), float_to_unorm8_bits(s[3]), 0x4440);
        return __byte_perm(rg, ba, 0x5410);
    }
    if (NEED_DST == kDstAlphaOver) {
        // SRC_ALPHA / ONE_MINUS_SRC_ALPHA / ADD on colour and alpha, all channels written (the usual "over" blend):
        // the generic expression of R8 with the factors resolved at compile time — same operations, same order
        const float sf = s[3], df = 1.0f - s[3];
        uint32_t out = 0;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const float t0 = s[ch] * sf, t1 = lut[(dst_word >> (8 * ch)) & 255u] * df;
            out |= float_to_unorm8(t0 + t1) << (8 * ch);
        }
        return out;
    }
    if (st.logic_op) {  // logicOpEnable: no blending; the write mask still applies (kDstGeneric only)
        uint32_t sw = 0, mask = 0;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            sw |= float_to_unorm8(s[ch]) << (8 * ch);
            if (st.color_write_mask & (1u << ch)) mask |= 0xffu << (8 * ch);
        }
        return (dst_word & ~mask) | (logic_op(st.logic_op - 1u, sw, dst_word) & mask);
    }
    if (st.blend_enable) {
        float d[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) d[ch] = lut[(dst_word >> (8 * ch)) & 255u];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            o[ch] = blend_op(st.color_blend_op, s[ch], blend_factor(st.src_color_blend_factor, ch, s, d), d[ch],
                             blend_factor(st.dst_color_blend_factor, ch, s, d));
        o[3] = blend_op(st.alpha_blend_op, s[3], blend_factor(st.src_alpha_blend_factor, 3, s, d), d[3],
                        blend_factor(st.dst_alpha_blend_factor, 3, s, d));
    } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) o[ch] = s[ch];
    }
    uint32_t out = dst_word;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
        if (st.color_write_mask & (1u << ch)) out = (out & ~(0xffu << (8 * ch))) | (float_to_unorm8(o[ch]) << (8 * ch));
    return out;
}

// perspective-correct vertex colour (R6): one reciprocal per fragment, one multiply per channel
__device__ __forceinline__ void interpolate_color(const float iw[3], const float col[3][4], float l0, float l1, float l2, float src[4]) {
    const float den = fmaf(l2, iw[2], fmaf(l1, iw[1], l0 * iw[0]));
    const float rden = __frcp_rn(den);  // IEEE 1.0f / den
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) src[ch] = __fmul_rn(fmaf(l2, col[2][ch], fmaf(l1, col[1][ch], l0 * col[0][ch])), rden);
}

#ifdef KZN_FS_TEMPLATE
// The compiled fragment program: this translation unit is also compiled to PTX with KZN_FS_TEMPLATE (build.py), where the
// fragment program is a call to this external function; at vkCreateGraphicsPipelines time (kzn_fragment_program_create,
// backend.cpp) every call is replaced by the shader's own PTX, generated from its shader-compiler-ir module by ptx_emit.cpp,
// and the result is loaded with cuModuleLoadDataEx.  Arguments: frag_coord (x, y = pixel centre, z = depth, w = 1/w), the
// four interpolated inputs, the push constant block, the descriptor table.
extern "C" __device__ float4 kzn_fs_main(float x, float y, float z, float w, float in0, float in1, float in2, float in3, const void* push_constants,
                                         const void* descriptor_table);
// ... and for two fragments at once (horizontally adjacent pixels of one triangle): where the shader's control flow does not
// depend on the fragment, the back end emits ONE copy of its code for both, with packed f32x2 arithmetic
struct FsPair {
    float c[8];
};
extern "C" __device__ FsPair kzn_fs_main2(float x0, float y0, float z0, float w0, float a0, float a1, float a2, float a3, float x1, float y1, float z1,
                                          float w1, float b0, float b1, float b2, float b3, const void* push_constants, const void* descriptor_table);
#endif

// does fragment program FS read anything that needs the fragment's barycentrics (interpolated inputs, depth, 1/w)?  When it
// does not, fully covered spans and blocks need no per-pixel coverage arithmetic at all.
template <uint32_t FS>
__device__ __forceinline__ bool fs_needs_interpolants(const DrawParams& p) {
    return FS == KZN_FS_VERTEX_COLOR ? true : FS == KZN_FS_PROCEDURAL_ORBIT ? false : (p.fs_flags & 1u) != 0;
}

// one invocation of fragment program FS: pixel (px, py), depth z, barycentrics l0..l2 over the triangle's per-vertex 1/w and
// colour/w (R6)
template <uint32_t FS>
__device__ __forceinline__ void run_fragment_program(const DrawParams& p, int px, int py, float z, const float iw[3], const float col[3][4],
                                                     float l0, float l1, float l2, float src[4]) {
    if (FS == KZN_FS_PROCEDURAL_ORBIT) {
        fs_procedural_orbit(px, py, p.inv_fb_w, p.inv_fb_h, src);
    } else if (FS == KZN_FS_VERTEX_COLOR) {
        interpolate_color(iw, col, l0, l1, l2, src);
    } else {
#ifdef KZN_FS_TEMPLATE
        float in[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float w = 1.0f;
        if (fs_needs_interpolants<FS>(p)) {
            interpolate_color(iw, col, l0, l1, l2, in);
            w = fmaf(l2, iw[2], fmaf(l1, iw[1], l0 * iw[0]));
        }
        const float4 c = kzn_fs_main((float)px + 0.5f, (float)py + 0.5f, z, w, in[0], in[1], in[2], in[3], p.push_constants,
                                     reinterpret_cast<const void*>(p.descriptor_table));
        src[0] = c.x;
        src[1] = c.y;
        src[2] = c.z;
        src[3] = c.w;
#endif
    }
}

// two invocations: pixels (px, py) and (px + 1, py) of one triangle
template <uint32_t FS>
__device__ __forceinline__ void run_fragment_program2(const DrawParams& p, int px, int py, float z0, float z1, const float iw[3], const float col[3][4],
                                                      const float la[3], const float lb[3], float src0[4], float src1[4]) {
    if (FS != KZN_FS_COMPILED) {
        run_fragment_program<FS>(p, px, py, z0, iw, col, la[0], la[1], la[2], src0);
        run_fragment_program<FS>(p, px + 1, py, z1, iw, col, lb[0], lb[1], lb[2], src1);
    } else {
#ifdef KZN_FS_TEMPLATE
        float a[4] = {0.0f, 0.0f, 0.0f, 0.0f}, b[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float w0 = 1.0f, w1 = 1.0f;
        if (fs_needs_interpolants<FS>(p)) {
            interpolate_color(iw, col, la[0], la[1], la[2], a);
            interpolate_color(iw, col, lb[0], lb[1], lb[2], b);
            w0 = fmaf(la[2], iw[2], fmaf(la[1], iw[1], la[0] * iw[0]));
            w1 = fmaf(lb[2], iw[2], fmaf(lb[1], iw[1], lb[0] * iw[0]));
        }
        const float y = (float)py + 0.5f;
        const FsPair c = kzn_fs_main2((float)px + 0.5f, y, z0, w0, a[0], a[1], a[2], a[3], (float)(px + 1) + 0.5f, y, z1, w1, b[0], b[1], b[2], b[3],
                                      p.push_constants, reinterpret_cast<const void*>(p.descriptor_table));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            src0[k] = c.c[k];
            src1[k] = c.c[4 + k];
        }
#endif
    }
}

// Depth test, fragment program and blend for the first n (<= 32) fragments queued for a warp's brick, one per lane.  A
// queue entry names the chunk slot of its triangle and the pixel; the edge functions at the pixel are rebuilt from the
// triangle's BrickEdge (exact 32-bit integers), then barycentrics and depth (R6).  Fragments of the batch that hit the same
// pixel are applied in lane order (= primitive order, R7) using match.any.  Returns 1 on the lanes whose fragment passed
// the depth test.  Not inlined: it has three call sites in the chunk loop.
template <uint32_t FS, int NEED_DST>
__device__ __noinline__ uint32_t shade_queue(const DrawParams& p, BrickState& bs, const float* lut, uint32_t lane, uint32_t n, int brick_px0,
                                             int brick_py0, bool depth_test, bool depth_write, bool has_color) {
    uint32_t n_passed = 0;
    const bool have = lane < n;
    const uint32_t key = have ? bs.queue[lane] : 0u;
    const uint32_t slot = key >> 8, pix = key & 255u;
    const BrickEdge& e = bs.edges[slot];
    const int4 ef = *reinterpret_cast<const int4*>(&e.f[0]);
    const int4 es = *reinterpret_cast<const int4*>(&e.s[0]);
    const int4 eu = *reinterpret_cast<const int4*>(&e.u[0]);
    const uint32_t xw = (uint32_t)ef.w;
    const int32_t dx = (int32_t)(pix & 15u) - (int32_t)(xw & 15u), dy = (int32_t)(pix >> 4) - (int32_t)(xw >> 4 & 15u);
    const int32_t F0 = ef.x + dx * es.x + dy * eu.x, F1 = ef.y + dx * es.y + dy * eu.y, F2 = ef.z + dx * es.z + dy * eu.z;
    const float inv2a = __int_as_float(es.w);
    const float l0 = (float)(F0 + (int32_t)(xw >> 13 & 1u)) * inv2a;  // R6: E_i = F_i + 1 - bias_i
    const float l1 = (float)(F1 + (int32_t)(xw >> 14 & 1u)) * inv2a;
    const float l2 = (float)(F2 + (int32_t)(xw >> 15 & 1u)) * inv2a;
    const TriRecC& t = bs.recs[slot];
    const float z = fmaf(l2, t.z[2], fmaf(l1, t.z[1], l0 * t.z[0]));
    const int px = brick_px0 + (int)(pix & 15u), py = brick_py0 + (int)(pix >> 4);
    const uint32_t kx = have ? pix : (0x80000000u | lane);  // unique key for idle lanes
    const uint32_t peers = __match_any_sync(0xffffffffu, kx);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    const uint32_t max_rank = __reduce_max_sync(0xffffffffu, have ? rank : 0u);
    float src[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (max_rank == 0) {  // warp-uniform, the usual case: every fragment of the batch has its pixel to itself
        bool pass = have;
        if (depth_test && have) pass = depth_compare(p.st.depth_compare_op, z, bs.depth[pix]);  // R7: early-Z
        if (pass) {
            n_passed = 1;
            if (depth_write) {
                bs.depth[pix] = z;
                bs.depth_dirty[pix] = 1;
            }
            if (has_color) {
                run_fragment_program<FS>(p, px, py, z, t.iw, t.col, l0, l1, l2, src);
                const uint32_t dst = NEED_DST ? bs.color[pix] : 0u;
                bs.color[pix] = output_merge_lut<NEED_DST>(p.st, lut, src, dst);
                bs.color_dirty[pix] = 1;
            }
        }
        __syncwarp();
        return n_passed;
    }
    if (have && has_color) {  // several fragments per pixel in this batch: shade up front, all lanes together
        run_fragment_program<FS>(p, px, py, z, t.iw, t.col, l0, l1, l2, src);
    }
    for (uint32_t rk = 0; rk <= max_rank; ++rk) {
        if (have && rank == rk) {
            bool pass = true;
            if (depth_test) pass = depth_compare(p.st.depth_compare_op, z, bs.depth[pix]);  // R7: early-Z
            if (pass) {
                n_passed = 1;
                if (depth_write) {
                    bs.depth[pix] = z;
                    bs.depth_dirty[pix] = 1;
                }
                if (has_color) {
                    const uint32_t dst = NEED_DST ? bs.color[pix] : 0u;
                    bs.color[pix] = output_merge_lut<NEED_DST>(p.st, lut, src, dst);
                    bs.color_dirty[pix] = 1;
                }
            }
        }
        __syncwarp();
    }
    return n_passed;
}

// PX interleaved orbits as PX/2 packed pairs: fma.rn.f32x2 / add.rn.f32x2 (FFMA2 / FADD2) do two IEEE fp32 operations per
// instruction with per-lane results identical to the scalar instructions, so the shader is still the oracle's sequence.
// orbit_cx: the row-independent part (c.x of the PX pixels), hoisted out of row loops by the direct kernel.
template <int PX>
__device__ __forceinline__ void orbit_cx(int px0, float inv_w, float2 cx[PX / 2]) {
    static_assert(PX % 2 == 0, "packed pairs");
    const float base = (float)px0;  // px0 + k + 0.5 is exact in fp32 (|px| < 2^22): same value as (float)(px0 + k) + 0.5f
#pragma unroll
    for (int k = 0; k < PX / 2; ++k) {
        const float u0 = (base + ((float)(2 * k) + 0.5f)) * inv_w, u1 = (base + ((float)(2 * k + 1) + 0.5f)) * inv_w;
        cx[k] = make_float2(fmaf(u0, 3.5f, -2.5f), fmaf(u1, 3.5f, -2.5f));
    }
}

template <int PX>
__device__ __forceinline__ void orbit_run(const float2 cx[PX / 2], int py, float inv_h, float2 zx[PX / 2], float2 zy[PX / 2]) {
    const float v = ((float)py + 0.5f) * inv_h;
    const float cys = fmaf(v, 2.0f, -1.0f);
    const float2 cy = make_float2(cys, cys);
    // Iteration 1 from z = (+0, +0) is the identity z <- c: fma(0,0,cx) = cx, fma(-0,0,cx) = cx, fma(0+0,0,cy) = cy, exactly,
    // because cx and cy are fma results and an exact-zero fma result is +0 under round-to-nearest (never -0).  So the
    // chains start at z = c and run the remaining ITERATIONS-1 steps: bit-identical to the oracle's loop.
#pragma unroll
    for (int k = 0; k < PX / 2; ++k) {
        zx[k] = cx[k];
        zy[k] = cy;
    }
#pragma unroll
    for (uint32_t it = 1; it < KZN_FS_ORBIT_ITERATIONS; ++it) {
#pragma unroll
        for (int k = 0; k < PX / 2; ++k) {
            const float2 t = __ffma2_rn(zx[k], zx[k], cx[k]);
            const float2 nzx = __ffma2_rn(make_float2(-zy[k].x, -zy[k].y), zy[k], t);
            const float2 nzy = __ffma2_rn(__fadd2_rn(zx[k], zx[k]), zy[k], cy);
            zx[k] = nzx;
            zy[k] = nzy;
        }
    }
}
template <int PX>
__device__ __forceinline__ void fs_procedural_orbit_n(const float2 cx[PX / 2], int py, float inv_h, float rgba[PX][4]) {
    float2 zx[PX / 2], zy[PX / 2];
    orbit_run<PX>(cx, py, inv_h, zx, zy);
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const float x = (k & 1) ? zx[k >> 1].y : zx[k >> 1].x, y = (k & 1) ? zy[k >> 1].y : zy[k >> 1].x;
        float m = fmaf(x, x, y * y);
        rgba[k][0] = 0.25f * m;
        rgba[k][1] = fmaf(x, 0.5f, 0.5f);
        rgba[k][2] = fmaf(y, 0.5f, 0.5f);
        rgba[k][3] = 1.0f;
    }
}
// The same fragment program with the opaque output merge (no blend, all channels written) folded in: PX packed RGBA8
// words.  Each channel is the oracle's sequence clamp01 -> * 255.0f -> round to nearest even; the rounding is done by
// adding 1.5 * 2^23 (every fp32 in [2^23, 2^24) is an integer, so the add rounds t in [0, 255] exactly like cvt.rni) and the
// integer is the low byte of the sum's bit pattern.  Blue is added to 1.5 * 2^23 + 0xFF00 instead, which leaves 0xFF — the
// alpha of this shader, 1.0 -> 255 — in the byte above it, so two byte permutes assemble the word: 3 FADD + 2 PRMT per
// pixel where F2I + shifts + ors were 8 issue slots.
template <int PX>
__device__ __forceinline__ void fs_procedural_orbit_rgba8(const float2 cx[PX / 2], int py, float inv_h, uint32_t out[PX]) {
    float2 zx[PX / 2], zy[PX / 2];
    orbit_run<PX>(cx, py, inv_h, zx, zy);
    const float kRound = 12582912.0f, kRoundAlpha = 12648192.0f;  // 1.5 * 2^23 (+ 0xFF00)
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const float x = (k & 1) ? zx[k >> 1].y : zx[k >> 1].x, y = (k & 1) ? zy[k >> 1].y : zy[k >> 1].x;
        const float m = fmaf(x, x, y * y);
        const float r = clamp01(0.25f * m), g = clamp01(fmaf(x, 0.5f, 0.5f)), b = clamp01(fmaf(y, 0.5f, 0.5f));
        const uint32_t rb = __float_as_uint(__fadd_rn(__fmul_rn(r, 255.0f), kRound));
        const uint32_t gb = __float_as_uint(__fadd_rn(__fmul_rn(g, 255.0f), kRound));
        const uint32_t bb = __float_as_uint(__fadd_rn(__fmul_rn(b, 255.0f), kRoundAlpha));
        out[k] = __byte_perm(__byte_perm(rb, gb, 0x4440), bb, 0x5410);  // [r, g] then [r, g, b, 0xFF]
    }
}
template <int PX>
__device__ __forceinline__ void fs_procedural_orbit_n(int px0, int py, float inv_w, float inv_h, float rgba[PX][4]) {
    float2 cx[PX / 2];
    orbit_cx<PX>(px0, inv_w, cx);
    fs_procedural_orbit_n<PX>(cx, py, inv_h, rgba);
}

// bits k of the pixel group px0..px0+PX-1 whose pixel lies in [xa, xb]
template <int PX>
__device__ __forceinline__ uint32_t quad_mask(int px0, int xa, int xb) {
    const int lo = max(xa - px0, 0), hi = min(xb - px0, PX - 1);
    return hi >= lo ? ((2u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
}

// One triangle against the PX adjacent pixels px0..px0+PX-1 of row py that a lane owns, destination values in registers:
// edge functions evaluated once and stepped by +16a per pixel (exact, R5), barycentric depth (R6), z clip (R3), early-Z
// (R7), fragment program, output merge (R8).  `live` / `zwrite` return which of the PX pixels got a new colour / depth.
// [xa, xb] is the triangle's pixel span in this row (bbox, already scissored) clipped to the warp's block; with UNIFORM_ROW
// all lanes of the warp work on the same row and span, and so the trivial-accept branch is warp-uniform (it is compiled out
// when every lane has its own row: raster_kernel's bricks).
template <uint32_t FS, int NEED_DST, int PX, bool UNIFORM_ROW = true>
__device__ __forceinline__ void span_quad(const DrawParams& p, const TriRec& t, const float* lut, int px0, int py, int32_t sy, int xa, int xb,
                                          bool depth_test, bool depth_write, bool has_color, uint32_t col[PX], float depth[PX],
                                          uint32_t& live, uint32_t& zwrite, uint32_t& n_covered, uint32_t& n_passed,
                                          const float2* orbit_cx_pre = nullptr) {
    const uint32_t bias = t.bias_bits;
    live = 0;
    zwrite = 0;
    bool trivial = false;
    if (UNIFORM_ROW && !fs_needs_interpolants<FS>(p) && !depth_test && (bias & 8u)) {
        // Trivial accept of the whole span: the covered pixels of a row are an interval (each biased edge test is a
        // half-plane), so if both end pixels of the span are covered, every pixel between them is; the depth clip cannot
        // reject (bit 3, set at setup) and this fragment program reads no interpolant — no per-pixel coverage work at all.
        // The six (edge, end pixel) tests are spread over lanes 0..5 — lane j tests edge j%3 at xa (j < 3) or xb — and
        // combined with one vote; all 32 lanes of the warp are converged here (span bounds are warp-uniform).
        const uint32_t lane = threadIdx.x & 31u, e = lane % 3u;
        const int32_t sx_t = 16 * (lane < 3u ? xa : xb) + 8;
        const bool ok = (long long)t.a[e] * sx_t + ((long long)t.b[e] * sy + t.c[e]) + (long long)((bias >> e) & 1u) > 0;
        trivial = (__ballot_sync(0xffffffffu, ok) & 0x3fu) == 0x3fu;
        if (trivial) {
            live = quad_mask<PX>(px0, xa, xb);
            n_covered += __popc(live);
            n_passed += __popc(live);
        }
    }
    float l0[PX], l1[PX], l2[PX], zf[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) l0[k] = l1[k] = l2[k] = zf[k] = 0.0f;  // a trivially accepted span has no barycentrics (and its program reads none)
    if (!trivial) {
        const long long b0 = bias & 1u, b1 = (bias >> 1) & 1u, b2 = (bias >> 2) & 1u;
        const long long r0 = (long long)t.b[0] * sy + t.c[0], r1 = (long long)t.b[1] * sy + t.c[1], r2 = (long long)t.b[2] * sy + t.c[2];
        const int32_t sx = 16 * px0 + 8;
        const long long e0 = (long long)t.a[0] * sx + r0, e1 = (long long)t.a[1] * sx + r1, e2 = (long long)t.a[2] * sx + r2;  // R5, exact
        const long long s0 = 16ll * t.a[0], s1 = 16ll * t.a[1], s2 = 16ll * t.a[2];
        const float inv2a = t.inv2a;
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const long long f0 = e0 + k * s0, f1 = e1 + k * s1, f2 = e2 + k * s2;
            bool cov = px0 + k >= xa && px0 + k <= xb && f0 + b0 > 0 && f1 + b1 > 0 && f2 + b2 > 0;
            l0[k] = (float)f0 * inv2a;  // R6
            l1[k] = (float)f1 * inv2a;
            l2[k] = (float)f2 * inv2a;
            const float z = fmaf(l2[k], t.z[2], fmaf(l1[k], t.z[1], l0[k] * t.z[0]));
            zf[k] = z;
            cov = cov && z >= p.zclip_lo && z <= p.zclip_hi;  // R3
            if (cov) {
                ++n_covered;
                if (!depth_test || depth_compare(p.st.depth_compare_op, z, depth[k])) {  // R7: early-Z
                    ++n_passed;
                    live |= 1u << k;
                    if (depth_write) {
                        depth[k] = z;
                        zwrite |= 1u << k;
                    }
                }
            }
        }
    }
    if (!has_color) live = 0;
    if (live) {
        float src[PX][4];
        if (FS == KZN_FS_PROCEDURAL_ORBIT) {
            if (orbit_cx_pre) fs_procedural_orbit_n<PX>(orbit_cx_pre, py, p.inv_fb_h, src);
            else fs_procedural_orbit_n<PX>(px0, py, p.inv_fb_w, p.inv_fb_h, src);
        } else {
#pragma unroll
            for (int k = 0; k < PX; k += 2) {
                const float la[3] = {l0[k], l1[k], l2[k]}, lb[3] = {l0[k + 1], l1[k + 1], l2[k + 1]};
                run_fragment_program2<FS>(p, px0 + k, py, zf[k], zf[k + 1], t.iw, t.col, la, lb, src[k], src[k + 1]);
            }
        }
#pragma unroll
        for (int k = 0; k < PX; ++k) {  // branch-free: pixels not live keep their value
            const uint32_t merged = output_merge_lut<NEED_DST>(p.st, lut, src[k], col[k]);
            col[k] = (live >> k & 1u) ? merged : col[k];
        }
    }
}

// A triangle of the chunk too large for the 32-bit brick arithmetic (1024 pixels or more across): lane l handles the 8
// pixels 8 (l & 1) .. + 7 of row l >> 1 of the brick directly from the expanded record (64-bit edge functions, 16-byte
// shared-memory accesses).  Every lane expands the same record (warp-uniform); not inlined, so that the registers of the
// expansion stay out of the streaming loop.
template <uint32_t FS, int NEED_DST>
__device__ __noinline__ void shade_wide(const DrawParams& p, BrickState& bs, const float* lut, uint32_t lane, uint32_t slot,
                                        int brick_px0, int brick_py0, bool depth_test, bool depth_write, bool has_color,
                                        uint32_t& n_covered, uint32_t& n_passed) {
    TriRec t;
    {
        const TriRecC& c = bs.recs[slot];
        expand_record(p, c, record_box(p, c), t);
    }
    const uint32_t at = lane * kBrickPx;  // (lane >> 1) * 16 + (lane & 1) * 8
    const int px0 = brick_px0 + (int)(lane & 1u) * kBrickPx, py = brick_py0 + (int)(lane >> 1);
    int xa = max(t.minx, brick_px0), xb = min(t.maxx, brick_px0 + kBrick - 1);
    if (py < t.miny || py > t.maxy) {  // my row is outside the pixel box: an empty span
        xa = 1;
        xb = 0;
    }
    float depth[kBrickPx];
    uint32_t col[kBrickPx];
#pragma unroll
    for (int h = 0; h < kBrickPx / 4; ++h) {
        float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (depth_test) dq = *reinterpret_cast<const float4*>(&bs.depth[at + 4 * h]);
        depth[4 * h] = dq.x; depth[4 * h + 1] = dq.y; depth[4 * h + 2] = dq.z; depth[4 * h + 3] = dq.w;
        uint4 cq = make_uint4(0u, 0u, 0u, 0u);
        if (has_color) cq = *reinterpret_cast<const uint4*>(&bs.color[at + 4 * h]);
        col[4 * h] = cq.x; col[4 * h + 1] = cq.y; col[4 * h + 2] = cq.z; col[4 * h + 3] = cq.w;
    }
    uint32_t live, zwrite;
    span_quad<FS, NEED_DST, kBrickPx, false>(p, t, lut, px0, py, 16 * py + 8, xa, xb, depth_test, depth_write, has_color, col, depth,
                                             live, zwrite, n_covered, n_passed);
#pragma unroll
    for (int k = 0; k < kBrickPx; ++k) {
        if (zwrite >> k & 1u) {
            bs.depth[at + k] = depth[k];
            bs.depth_dirty[at + k] = 1;
        }
        if (live >> k & 1u) {
            bs.color[at + k] = col[k];
            bs.color_dirty[at + k] = 1;
        }
    }
    __syncwarp();
}

// Draws of a few triangles (total <= kDirectMax: fullscreen passes, quads, UI): no binning, no staging, no fragment queue.
// Persistent one-warp CTAs (kDirectWarpsPerSM per SM).  A warp takes *blocks* of 256 x kDirectRows pixels from a global
// work counter (zeroed by setup_kernel) until none are left, so all SM sub-partitions stay busy to the end whatever the
// frame size — with a static grid the last wave of equal-cost blocks lands on a subset of the SMs.  Blocks are numbered
// column-major, so a warp mostly stays in one 256-pixel column and everything that depends on the column only (clip masks,
// the c.x of the fill-rate shader) is recomputed only when the column changes.  Within a block the warp walks its rows one
// after the other, lane l = pixels 8l..8l+7 of the row, colour and depth of the row held in REGISTERS while the lane walks
// the triangle records in primitive order (uniform loads: every lane reads the same record) — per-pixel fragment order is
// the primitive order by construction and nothing is shared between warps but the UNORM8 table.
// *Block-level trivial accept*: lane i tests triangle i (the first 32) at the corner of the block where each edge function
// is smallest; if all three are positive the whole block is inside the triangle and its rows skip coverage entirely.
// Direct draws of at most kFusedMax triangles, one of them clipped (R3; rare): the calling warp makes the draw's records
// again, lane i = triangle i, a clipped one taking as many of the 32 slots of `recs` as its polygon has triangles; returns
// their number.  Kept out of line: nothing of it (the clipper's polygons least of all) should cost the raster loop a register.
__device__ __noinline__ uint32_t fused_clip_records(const DrawParams& p, uint32_t lane, TriRec* recs) {
    TriRec r;
    bool ok = false, slow = false;
    const bool mine = lane < p.total_tris;
    if (mine) ok = make_record(p, lane % p.tris_per_instance, r, slow, false);  // (fused draws are triangles: draw_common)
    uint32_t n_recs = p.total_tris;
    __syncwarp();
    {  // records placed by a prefix sum over the polygon sizes
        float poly[8][8];
        int m = 0;
        if (slow) {
            m = clip_primitive(p, lane % p.tris_per_instance, poly, false);
            if (blockIdx.x == 0) clip_statistics(p, m, 1ull, false);
        }
        const uint32_t cnt = !mine ? 0u : (slow && m > 3) ? (uint32_t)m - 2u : 1u;
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += y;
        }
        n_recs = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t at = incl - cnt;
        for (uint32_t j = 0; j < cnt; ++j) {
            if (slow) {
                if (m >= 3) ok = make_record_from(p, poly, (int)j, r, false) || ok;
                else empty_record(r);
            }
            recs[at + j] = r;
        }
    }
    if (mine && p.stats && blockIdx.x == 0) {
        atomicAdd(&p.stats[0], 1ull);
        if (!ok) atomicAdd(&p.stats[1], 1ull);
    }
    return n_recs;
}

template <uint32_t FS, int NEED_DST>
__device__ __forceinline__ void raster_direct_body(const DrawParams& p) {
    kzn_pdl_wait();
    __shared__ float unorm_lut[NEED_DST ? 256 : 1];
    const uint32_t lane = threadIdx.x;
    const bool has_color = p.color_base != 0, has_depth = p.depth_base != 0;
    const bool depth_test = has_depth && p.st.depth_test_enable;
    const bool depth_write = depth_test && p.st.depth_write_enable;
    if (NEED_DST) {
#pragma unroll
        for (int i = 0; i < 8; ++i) unorm_lut[lane + 32 * i] = unorm8_to_float(lane + 32 * i);
        __syncwarp();
    }
    // Draws of at most kFusedMax triangles (fullscreen passes, quads) have no setup_kernel launch: every warp makes the records
    // itself, lane i = triangle i, into its own shared memory — a few hundred instructions per warp against a dependent
    // kernel launch per draw.  The work counter is then handed back by the last warp to leave (see the end of the kernel).
    __shared__ TriRec own_recs[32];
    __shared__ uint32_t n_recs_s;  // records of the draw: its triangles, a clipped one counting as many as its polygon has triangles (R3)
    uint32_t n_recs = p.total_tris;
    if (p.fused_setup) {
        bool ok = false, slow = false;
        if (lane < p.total_tris) {
            TriRec r;
            ok = make_record(p, lane % p.tris_per_instance, r, slow, false);
            own_recs[lane] = r;
        }
        if (__builtin_expect(__any_sync(0xffffffffu, slow), 0)) {
            n_recs = fused_clip_records(p, lane, own_recs);
        } else if (lane < p.total_tris && p.stats && blockIdx.x == 0) {
            atomicAdd(&p.stats[0], 1ull);
            if (!ok) atomicAdd(&p.stats[1], 1ull);
        }
    } else {
        n_recs = p.work_counter[3];  // setup_kernel<false>
    }
    if (lane == 0) n_recs_s = n_recs;  // (read from shared memory where it is needed: a register for the whole kernel costs the row loop a spill)
    __syncwarp();
#define n_recs (*(volatile uint32_t*)&n_recs_s)
    const TriRec* recs = p.fused_setup ? own_recs : p.tris;
    const uint32_t n_blocks = p.tiles_x * p.tiles_y;
    const unsigned long long pitch = (unsigned long long)p.fb_w * 4ull;
    uint32_t n_covered = 0, n_passed = 0;
    int cur_col = -1, tile_px0 = 0, px0 = 0;
    uint32_t quad_clip = 0;
    bool cvec = false, dvec = false;
    bool col_whole = false;  // every lane's 8 pixels of this column are inside the clip rectangle and go out as vectors
    float2 cxp[kDirectPx / 2];
    uint32_t col[kDirectPx];
    float depth[kDirectPx];
#pragma unroll
    for (int k = 0; k < kDirectPx; ++k) {  // values of pixels that are never live are never stored either
        col[k] = 0u;
        depth[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < kDirectPx / 2; ++k) cxp[k] = make_float2(0.f, 0.f);

    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(p.work_counter, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_blocks) break;
        const int bcol = (int)(u / p.tiles_y), brow = (int)(u - (uint32_t)bcol * p.tiles_y);
        if (bcol != cur_col) {  // per-column state
            cur_col = bcol;
            tile_px0 = (int)(p.tile_x0 + (uint32_t)bcol) * kDirectTileW;
            px0 = tile_px0 + (int)lane * kDirectPx;
            quad_clip = quad_mask<kDirectPx>(px0, p.cx0, p.cx1);  // pixels of my group inside the clip rectangle
            const bool vec_ok = (p.fb_w & 3u) == 0 && px0 + kDirectPx <= (int)p.fb_w;
            cvec = vec_ok && (p.color_base & 15ull) == 0;
            dvec = vec_ok && (p.depth_base & 15ull) == 0;
            col_whole = __all_sync(0xffffffffu, quad_clip == (1u << kDirectPx) - 1u && cvec);
            if (FS == KZN_FS_PROCEDURAL_ORBIT) orbit_cx<kDirectPx>(px0, p.inv_fb_w, cxp);
        }
        const int row0 = (int)(p.tile_y0 + (uint32_t)brow) * kDirectRows;
        uint32_t full = 0;  // bit i: triangle i covers every pixel centre of this block (and needs no per-fragment work)
        if (!fs_needs_interpolants<FS>(p) && !depth_test) {
            bool ok = false;
            if (lane < n_recs) {
                const TriRec& t = recs[lane];
                const uint32_t bias = t.bias_bits;
                ok = t.minx <= t.maxx && (bias & 8u);  // produces fragments; the depth clip cannot reject (setup)
                const int32_t sx_lo = 16 * tile_px0 + 8, sx_hi = 16 * (tile_px0 + kDirectTileW - 1) + 8;
                const int32_t sy_lo = 16 * row0 + 8, sy_hi = 16 * (row0 + kDirectRows - 1) + 8;
#pragma unroll
                for (int e = 0; e < 3; ++e) {  // the edge function is linear: its minimum over the block is at one corner
                    const int32_t a = t.a[e], b = t.b[e];
                    const long long v = (long long)a * (a >= 0 ? sx_lo : sx_hi) + (long long)b * (b >= 0 ? sy_lo : sy_hi) + t.c[e] +
                                        (long long)((bias >> e) & 1u);
                    ok = ok && v > 0;
                }
            }
            full = __ballot_sync(0xffffffffu, ok);
        }
        const int r_lo = max(row0, p.cy0), r_hi = min(row0 + kDirectRows - 1, p.cy1);
        if (FS != KZN_FS_VERTEX_COLOR && NEED_DST == kDstNone) {
            // The block is nothing but shader work when the LAST triangle of the draw covers it completely and opaquely
            // (whatever earlier triangles produced is overwritten; statistics still need their fragments, so with
            // statistics on this is limited to single-triangle draws) and the block lies inside the clip rectangle:
            // four rows of orbit + packed conversion + two 16-byte stores per lane, no per-row bookkeeping.
            const uint32_t last = n_recs - 1u;
            if (last < 32u && (full >> last & 1u) && (last == 0u || !p.stats) && has_color && !p.color_tiled && col_whole &&
                r_lo == row0 && r_hi == row0 + kDirectRows - 1) {
                unsigned long long a = p.color_base + ((unsigned long long)row0 * p.fb_w + px0) * 4ull;
#pragma unroll 1
                for (int r = 0; r < kDirectRows; ++r, a += pitch) {
                    uint32_t out[kDirectPx];
                    if (FS == KZN_FS_PROCEDURAL_ORBIT) {
                        fs_procedural_orbit_rgba8<kDirectPx>(cxp, row0 + r, p.inv_fb_h, out);
                    } else {  // a compiled program that reads no interpolant (full is zero otherwise): its colour, converted (E.1)
                        const float none[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                        for (int k = 0; k < kDirectPx; k += 2) {
                            float src0[4], src1[4];
                            run_fragment_program2<FS>(p, px0 + k, row0 + r, 0.0f, 0.0f, nullptr, nullptr, none, none, src0, src1);
                            out[k] = output_merge_lut<kDstNone>(p.st, unorm_lut, src0, 0u);
                            out[k + 1] = output_merge_lut<kDstNone>(p.st, unorm_lut, src1, 0u);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < kDirectPx / 4; ++h)
                        *reinterpret_cast<uint4*>(a + 16 * h) = make_uint4(out[4 * h], out[4 * h + 1], out[4 * h + 2], out[4 * h + 3]);
                }
                n_covered += (last + 1u) * kDirectPx * kDirectRows;  // last > 0 only without statistics
                n_passed += (last + 1u) * kDirectPx * kDirectRows;
                if (p.stats && lane == 0) atomicAdd(&p.stats[6], 1ull);
                continue;
            }
        }
        unsigned long long caddr = 0, daddr = 0;
        if (has_color) caddr = p.color_tiled ? 0ull : p.color_base + ((unsigned long long)r_lo * p.fb_w + px0) * 4ull;
        if (has_depth) daddr = p.depth_base + ((unsigned long long)r_lo * p.fb_w + px0) * 4ull;
#pragma unroll 1
        for (int py = r_lo; py <= r_hi; ++py, caddr += pitch, daddr += pitch) {
            if (has_color && p.color_tiled) caddr = color_addr(p, px0, py);
            if (quad_clip) {
                if (NEED_DST && has_color) {
#pragma unroll
                    for (int h = 0; h < kDirectPx / 4; ++h) {  // 16-byte vectors where the whole group of 4 is inside
                        const uint32_t m4 = (quad_clip >> (4 * h)) & 15u;
                        if (m4 == 15u && cvec) {
                            const uint4 q = *reinterpret_cast<const uint4*>(caddr + 16 * h);
                            col[4 * h] = q.x; col[4 * h + 1] = q.y; col[4 * h + 2] = q.z; col[4 * h + 3] = q.w;
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (m4 >> k & 1u) col[4 * h + k] = *reinterpret_cast<const uint32_t*>(caddr + 16 * h + 4 * k);
                        }
                    }
                }
                if (depth_test) {
#pragma unroll
                    for (int h = 0; h < kDirectPx / 4; ++h) {
                        const uint32_t m4 = (quad_clip >> (4 * h)) & 15u;
                        if (m4 == 15u && dvec) {
                            const float4 q = *reinterpret_cast<const float4*>(daddr + 16 * h);
                            depth[4 * h] = q.x; depth[4 * h + 1] = q.y; depth[4 * h + 2] = q.z; depth[4 * h + 3] = q.w;
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (m4 >> k & 1u) depth[4 * h + k] = *reinterpret_cast<const float*>(daddr + 16 * h + 4 * k);
                        }
                    }
                }
            }
            const int32_t sy = 16 * py + 8;  // fixed-point sample row (R2)
            uint32_t cdirty = 0, zdirty = 0;
            for (uint32_t i = 0; i < n_recs; ++i) {
                if (FS != KZN_FS_VERTEX_COLOR && i < 32u && (full >> i & 1u)) {  // warp-uniform
                    // block-level trivial accept: every pixel of the row inside the clip rectangle is covered and passes
                    n_covered += __popc(quad_clip);
                    n_passed += __popc(quad_clip);
                    if (has_color && quad_clip) {
                        float src[kDirectPx][4];
                        if (FS == KZN_FS_PROCEDURAL_ORBIT) {
                            fs_procedural_orbit_n<kDirectPx>(cxp, py, p.inv_fb_h, src);
                        } else {
                            const float none[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                            for (int k = 0; k < kDirectPx; k += 2)
                                run_fragment_program2<FS>(p, px0 + k, py, 0.0f, 0.0f, nullptr, nullptr, none, none, src[k], src[k + 1]);
                        }
#pragma unroll
                        for (int k = 0; k < kDirectPx; ++k) {
                            const uint32_t merged = output_merge_lut<NEED_DST>(p.st, unorm_lut, src[k], col[k]);
                            // without a destination read the pixels outside the clip mask are never stored: no select needed
                            col[k] = (NEED_DST == kDstNone || (quad_clip >> k & 1u)) ? merged : col[k];
                        }
                        cdirty |= quad_clip;
                    }
                    continue;
                }
                const TriRec& t = recs[i];
                const int4 bb = *reinterpret_cast<const int4*>(&t);  // minx maxx miny maxy, already scissored
                if (bb.x > bb.y || py < bb.z || py > bb.w) continue;
                const int xa = max(bb.x, tile_px0), xb = min(bb.y, tile_px0 + kDirectTileW - 1);
                if (xa > xb) continue;
                uint32_t live, zwrite;
                span_quad<FS, NEED_DST, kDirectPx>(p, t, unorm_lut, px0, py, sy, xa, xb, depth_test, depth_write, has_color, col, depth, live,
                                                   zwrite, n_covered, n_passed, FS == KZN_FS_PROCEDURAL_ORBIT ? cxp : nullptr);
                cdirty |= live;
                zdirty |= zwrite;
            }
            cdirty &= quad_clip;
            zdirty &= quad_clip;
#pragma unroll
            for (int h = 0; h < kDirectPx / 4; ++h) {
                const uint32_t c4 = (cdirty >> (4 * h)) & 15u, z4 = (zdirty >> (4 * h)) & 15u;
                if (c4 == 15u && cvec) {
                    *reinterpret_cast<uint4*>(caddr + 16 * h) = make_uint4(col[4 * h], col[4 * h + 1], col[4 * h + 2], col[4 * h + 3]);
                } else if (c4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (c4 >> k & 1u) *reinterpret_cast<uint32_t*>(caddr + 16 * h + 4 * k) = col[4 * h + k];
                }
                if (z4 == 15u && dvec) {
                    *reinterpret_cast<float4*>(daddr + 16 * h) = make_float4(depth[4 * h], depth[4 * h + 1], depth[4 * h + 2], depth[4 * h + 3]);
                } else if (z4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (z4 >> k & 1u) *reinterpret_cast<float*>(daddr + 16 * h + 4 * k) = depth[4 * h + k];
                }
            }
        }
        if (p.stats && lane == 0) atomicAdd(&p.stats[6], 1ull);
    }
    if (p.stats) {
        for (int d = 16; d > 0; d >>= 1) {
            n_covered += __shfl_xor_sync(0xffffffffu, n_covered, d);
            n_passed += __shfl_xor_sync(0xffffffffu, n_passed, d);
        }
        if (lane == 0) {
            if (n_covered) atomicAdd(&p.stats[4], (unsigned long long)n_covered);
            if (n_passed) atomicAdd(&p.stats[5], (unsigned long long)n_passed);
        }
    }
    if (lane == 0) {
        // The block dispenser must be zero again for the next direct draw, which may have no setup_kernel to zero it: the
        // last warp to leave (every warp has made its last request by then) resets it and the count of leavers
        __threadfence();
        if (atomicAdd(p.work_counter + 1, 1u) == gridDim.x - 1u) {
            p.work_counter[0] = 0u;
            p.work_counter[1] = 0u;
        }
    }
}

#undef n_recs

__constant__ uint32_t kWidthMagic[17] = {0,     65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192,
                                         7282,  6554,  5958,  5462,  5042,  4682,  4370, 4096};  // ceil(2^16 / w)

// One chunk of a brick's work: the first n (<= 32) slots of bs.recs hold the compact records of triangles whose
// pixel box meets the brick, in primitive order.
//  1. Lane j derives triangle j for this brick (BrickEdge): its pixel box inside the brick, the three biased edge
//     functions at the box's first pixel and their steps in x and y, 1/2A — exact 32-bit integer arithmetic for triangles
//     of less than 1024 pixels extent.
//  2. The candidate pixels of all boxes are streamed 32 per step, one per lane, whatever triangle they belong to (owner
//     found by binary search over the prefix sums of the box sizes): coverage is a sign test of F_i + dx s_i + dy u_i.
//     Covered fragments are appended to the warp's queue in candidate order (= primitive order) and shaded in full batches
//     of 32 (shade_queue), so small triangles fill the lanes as well as large ones in both stages.
//  3. Triangles too large for 32-bit arithmetic go through shade_wide; they split the chunk into segments so that primitive
//     order is kept.  The queue is drained before the chunk's records are replaced.
template <uint32_t FS, int NEED_DST>
__device__ __forceinline__ void raster_chunk(const DrawParams& p, BrickState& bs, const float* lut, uint32_t lane, uint32_t n, int brick_px0,
                                             int brick_py0, bool depth_test, bool depth_write, bool has_color, uint32_t& n_covered,
                                             uint32_t& n_passed) {
    constexpr uint32_t kFull = 0xffffffffu;
    const bool have = lane < n;
    uint32_t count = 0;
    bool wide = false;
    if (have) {
        const TriRecC& r = bs.recs[lane];
        const int4 q0 = *reinterpret_cast<const int4*>(&r.X[0]);  // X0 X1 X2 Y0
        const int2 q1 = *reinterpret_cast<const int2*>(&r.Y[1]);  // Y1 Y2
        const int32_t X0 = q0.x, X1 = q0.y, X2 = q0.z, Y0 = q0.w, Y1 = q1.x, Y2 = q1.y;
        const int32_t minX = min(X0, min(X1, X2)), maxX = max(X0, max(X1, X2));
        const int32_t minY = min(Y0, min(Y1, Y2)), maxY = max(Y0, max(Y1, Y2));
        // the triangle's scissored pixel box, cut to the brick
        const int x0 = max(max((minX - 8 + 15) >> 4, p.cx0), brick_px0), x1 = min(min((maxX - 8) >> 4, p.cx1), brick_px0 + kBrick - 1);
        const int y0 = max(max((minY - 8 + 15) >> 4, p.cy0), brick_py0), y1 = min(min((maxY - 8) >> 4, p.cy1), brick_py0 + kBrick - 1);
        const int w = x1 - x0 + 1, h = y1 - y0 + 1;
        if (w > 0 && h > 0) {
            const bool fits32 = maxX - minX < 16384 && maxY - minY < 16384;
            wide = !fits32;
            if (fits32) {
                count = (uint32_t)(w * h);
                const int32_t two_a = (X1 - X0) * (Y2 - Y0) - (X2 - X0) * (Y1 - Y0);  // R4; |.| < 2^29
                const int32_t sg = two_a < 0 ? -1 : 1;
                const int32_t sx = 16 * x0 + 8, sy = 16 * y0 + 8;  // fixed-point sample position (R2)
                // edge i is opposite vertex i: from vertex (i+1)%3 to vertex (i+2)%3 (R5);  E_i(S) = a_i (Sx - X_ia) + b_i (Sy - Y_ia)
                const int32_t a0 = -(Y2 - Y1) * sg, b0 = (X2 - X1) * sg;
                const int32_t a1 = -(Y0 - Y2) * sg, b1 = (X0 - X2) * sg;
                const int32_t a2 = -(Y1 - Y0) * sg, b2 = (X1 - X0) * sg;
                const int32_t t0 = (a0 > 0 || (a0 == 0 && b0 > 0)) ? 1 : 0;
                const int32_t t1 = (a1 > 0 || (a1 == 0 && b1 > 0)) ? 1 : 0;
                const int32_t t2 = (a2 > 0 || (a2 == 0 && b2 > 0)) ? 1 : 0;
                const int32_t f0 = a0 * (sx - X1) + b0 * (sy - Y1) + t0 - 1;
                const int32_t f1 = a1 * (sx - X2) + b1 * (sy - Y2) + t1 - 1;
                const int32_t f2 = a2 * (sx - X0) + b2 * (sy - Y0) + t2 - 1;
                uint32_t xw = (uint32_t)(x0 - brick_px0) | (uint32_t)(y0 - brick_py0) << 4 | (uint32_t)w << 8 | (uint32_t)(1 - t0) << 13 |
                              (uint32_t)(1 - t1) << 14 | (uint32_t)(1 - t2) << 15;
                if (depth_clip_cannot_reject(p, r.z[0], r.z[1], r.z[2])) xw |= 1u << 16;
                BrickEdge& e = bs.edges[lane];
                *reinterpret_cast<int4*>(&e.f[0]) = make_int4(f0, f1, f2, (int)xw);
                *reinterpret_cast<int4*>(&e.s[0]) = make_int4(16 * a0, 16 * a1, 16 * a2, __float_as_int(1.0f / (float)(two_a * sg)));
                *reinterpret_cast<int4*>(&e.u[0]) = make_int4(16 * b0, 16 * b1, 16 * b2, (int)kWidthMagic[w]);
            }
        }
    }
    __syncwarp();
    uint32_t wide_mask = __ballot_sync(kFull, wide);
    uint32_t incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(kFull, incl, d);
        if ((int)lane >= d) incl += y;
    }
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    const uint32_t excl = incl - count;
    uint32_t* queue = bs.queue;
    uint32_t q_n = 0;  // fragments waiting in the queue (warp-uniform)
    uint32_t seg_begin = 0;
    for (;;) {
        const int wide_lane = wide_mask ? __ffs(wide_mask) - 1 : -1;
        const uint32_t seg_end = wide_lane >= 0 ? __shfl_sync(kFull, excl, wide_lane) : total;
        for (uint32_t c0 = seg_begin; c0 < seg_end; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool active = c < seg_end;
            // owner = the last lane whose exclusive prefix is <= c (binary search over the 32 prefixes)
            uint32_t pos = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                uint32_t e = __shfl_sync(kFull, excl, (pos + step) & 31u);
                if (e <= c) pos += step;
            }
            const uint32_t k = c - __shfl_sync(kFull, excl, pos);
            const BrickEdge& e = bs.edges[pos];
            const int4 ef = *reinterpret_cast<const int4*>(&e.f[0]);
            const int4 es = *reinterpret_cast<const int4*>(&e.s[0]);
            const int4 eu = *reinterpret_cast<const int4*>(&e.u[0]);
            const uint32_t xw = (uint32_t)ef.w;
            const uint32_t dy = (k * (uint32_t)eu.w) >> 16, dx = k - dy * (xw >> 8 & 31u);
            const int32_t F0 = ef.x + (int32_t)dx * es.x + (int32_t)dy * eu.x;  // R5, exact
            const int32_t F1 = ef.y + (int32_t)dx * es.y + (int32_t)dy * eu.y;
            const int32_t F2 = ef.z + (int32_t)dx * es.z + (int32_t)dy * eu.z;
            bool cov = active && (F0 | F1 | F2) >= 0;
            if (cov && !(xw >> 16 & 1u)) {  // the depth clip may reject (R3): evaluate the depth here as shade_queue will
                const float inv2a = __int_as_float(es.w);
                const float l0 = (float)(F0 + (int32_t)(xw >> 13 & 1u)) * inv2a;
                const float l1 = (float)(F1 + (int32_t)(xw >> 14 & 1u)) * inv2a;
                const float l2 = (float)(F2 + (int32_t)(xw >> 15 & 1u)) * inv2a;
                const TriRecC& t = bs.recs[pos];
                const float z = fmaf(l2, t.z[2], fmaf(l1, t.z[1], l0 * t.z[0]));
                cov = z >= p.zclip_lo && z <= p.zclip_hi;  // false for NaN
            }
            const uint32_t ballot = __ballot_sync(kFull, cov);
            if (!ballot) continue;
            if (cov) {
                ++n_covered;
                queue[q_n + __popc(ballot & ((1u << lane) - 1u))] = pos << 8 | ((xw >> 4 & 15u) + dy) << 4 | ((xw & 15u) + dx);
            }
            q_n += __popc(ballot);
            __syncwarp();
            if (q_n >= 32u) {
                n_passed += shade_queue<FS, NEED_DST>(p, bs, lut, lane, 32u, brick_px0, brick_py0, depth_test, depth_write, has_color);
                const uint32_t rest = lane < q_n - 32u ? queue[32u + lane] : 0u;  // move the remainder (< 32 entries) to the front
                __syncwarp();
                if (lane < q_n - 32u) queue[lane] = rest;
                q_n -= 32u;
                __syncwarp();
            }
        }
        if (q_n) {  // before a wide triangle, and before the chunk's records are replaced
            n_passed += shade_queue<FS, NEED_DST>(p, bs, lut, lane, q_n, brick_px0, brick_py0, depth_test, depth_write, has_color);
            q_n = 0;
        }
        if (wide_lane < 0) break;
        shade_wide<FS, NEED_DST>(p, bs, lut, lane, (uint32_t)wide_lane, brick_px0, brick_py0, depth_test, depth_write, has_color, n_covered, n_passed);
        wide_mask &= wide_mask - 1;
        seg_begin = seg_end;
    }
}

// The bricks of the tile at (tile_px0, tile_py0) that a packed pixel box meets: bit 4 by + bx.
__device__ __forceinline__ uint32_t brick_mask_of(uint2 box, int tile_px0, int tile_py0) {
    const int minx = (int)(box.x & 0xffffu), maxx = (int)(box.x >> 16), miny = (int)(box.y & 0xffffu), maxy = (int)(box.y >> 16);
    if (minx > maxx) return 0u;
    const int bx0 = max(minx - tile_px0, 0) >> 4, bx1 = min(maxx - tile_px0, kTileW - 1) >> 4;
    const int by0 = max(miny - tile_py0, 0) >> 4, by1 = min(maxy - tile_py0, kTileH - 1) >> 4;
    if (bx0 > bx1 || by0 > by1) return 0u;
    const uint32_t xm = (2u << bx1) - (1u << bx0);
    return (by0 == 0 ? xm : 0u) | (by1 >= 1 ? xm << 4 : 0u);
}

// where the brick masks of a tile's list live: 4-byte aligned so that a lane can fetch four entries at once (the gap of
// 4 bytes per tile keeps consecutive tiles' regions apart after rounding up)
// the record (and box) a list entry names: record << 3 | fan index j; j > 0: extra record j - 1 of the clipped primitive
__device__ __forceinline__ uint32_t entry_record(const DrawParams& p, uint32_t e) {
    const uint32_t rec = e >> 3, j = e & 7u;
    if (j == 0u) return rec;
    return p.total_tris + (p.clip_link[rec % p.tris_per_instance] & 0x1fffffffu) + j - 1u;
}

__device__ __forceinline__ uint8_t* tile_masks(const DrawParams& p, uint32_t tile, uint32_t beg) {
    return p.bin_masks + (((size_t)beg + 3u) & ~(size_t)3u) + 4u * (size_t)tile;
}

#ifndef KZN_FS_TEMPLATE
// (tile_sort_warp_kernel: frames of many tiles with short lists; tile_sort_cta_kernel below: fewer tiles with long lists.)
// One warp per 64 x 32-pixel screen tile: the binning atomics deliver the tile's triangle ids in arbitrary order; they are
// sorted here (ascending = primitive order, R7) and each entry's pixel box is looked at once, for all eight bricks of the
// tile: one byte per entry that raster_kernel's warps filter the list by.
// Stable LSD radix sort, 8-bit digits, only over the bits in which the ids of this tile differ (ids of one tile come from a
// narrow range of a mesh: typically 2 passes).  The warp walks the list 32 ids at a time: match.any ranks equal digits
// inside a batch in lane order, a 256-entry table of running offsets in shared memory orders the batches.  The lists
// themselves stay in global memory (the tile's segment of the bins and of a second buffer of the same size, ping-pong;
// a few KiB that live in L2), so a warp needs 1 KiB of shared memory and no CTA barrier — 64 sorting warps per SM, where a
// CTA per tile with the lists in shared memory spent its time in barriers (9 us per tile whatever its length).
// The list of one tile, n ids in the tile's segment of the bins, sorted by ONE warp in global memory (ping-pong with the
// second buffer; `table`: 256 words of shared memory), then written back with its brick masks.  Any length: lists beyond
// kSortCap, which the CTA form cannot hold in shared memory, come here too (a tile with tens of thousands of triangles takes
// a warp a millisecond — against every brick of the tile walking every triangle of the draw, which is what used to happen).
__device__ __forceinline__ void sort_tile_in_global(const DrawParams& p, uint32_t tile, uint32_t beg, uint32_t n, uint32_t* table, uint32_t lane,
                                                    int tile_px0, int tile_py0) {
    constexpr uint32_t kFull = 0xffffffffu;
    uint32_t* a = p.bin_entries + beg;
    uint32_t* b = p.bin_entries_alt + beg;
    // range of the ids
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (uint32_t i = lane; i < n; i += 32u) {
        const uint32_t v = a[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(kFull, lo) & ~7u;  // entries are record << 3 | fan index: the triangles of one primitive do not overlap, their order is free
    hi = __reduce_max_sync(kFull, hi);
    const uint32_t span = hi - lo;
    const uint32_t bits = span ? 32u - (uint32_t)__clz(span) : 0u;
    for (uint32_t shift = 3; shift < bits; shift += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) table[lane + 32 * k] = 0;
        __syncwarp();
        // digit counts (four batches of 32 ids in flight: the lists live in L2, a batch per round trip was the whole kernel)
        for (uint32_t base = 0; base < n; base += 128u) {
            uint32_t v[4];
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) v[k] = base + 32u * k + lane < n ? __ldcg(a + base + 32u * k + lane) : 0u;
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                if (base + 32u * k >= n) break;
                const bool valid = base + 32u * k + lane < n;
                const uint32_t d = valid ? ((v[k] - lo) >> shift) & 255u : 256u + lane;
                const uint32_t peers = __match_any_sync(kFull, d);
                if (valid && (int)lane == __ffs(peers) - 1) table[d] += __popc(peers);
                __syncwarp();
            }
        }
        // exclusive scan of the 256 counts: lane l owns digits 8l .. 8l + 7
        {
            uint32_t v[8], total = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = total;
                total += table[lane * 8 + k];
            }
            uint32_t incl = total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, d);
                if ((int)lane >= d) incl += y;
            }
            const uint32_t before = incl - total;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) table[lane * 8 + k] = before + v[k];
        }
        __syncwarp();
        // stable scatter
        for (uint32_t base = 0; base < n; base += 128u) {
            uint32_t vv[4];
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) vv[k] = base + 32u * k + lane < n ? __ldcg(a + base + 32u * k + lane) : 0u;
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                if (base + 32u * k >= n) break;
                const bool valid = base + 32u * k + lane < n;
                const uint32_t v = vv[k];
                const uint32_t d = valid ? ((v - lo) >> shift) & 255u : 256u + lane;
                const uint32_t peers = __match_any_sync(kFull, d);
                const int leader = __ffs(peers) - 1;
                uint32_t start = 0;
                if (valid && (int)lane == leader) {
                    start = table[d];
                    table[d] = start + __popc(peers);
                }
                start = __shfl_sync(kFull, start, leader);
                if (valid) __stcg(b + start + __popc(peers & ((1u << lane) - 1u)), v);
                __syncwarp();
            }
        }
        __threadfence_block();  // the next pass (other lanes of this warp) reads what this one wrote
        __syncwarp();
        uint32_t* t = a;
        a = b;
        b = t;
    }
    // the sorted list goes to the tile's segment of the bins (if the last pass left it in the other buffer) with its masks
    uint32_t* out = p.bin_entries + beg;
    uint8_t* masks = tile_masks(p, tile, beg);
    for (uint32_t i = lane; i < n; i += 32u) {
        const uint32_t id = __ldcg(a + i);
        if (a != out) out[i] = id;
        masks[i] = (uint8_t)brick_mask_of(__ldg(&p.boxes[entry_record(p, id)]), tile_px0, tile_py0);
    }
}

__global__ void __launch_bounds__(kSortThreads) tile_sort_warp_kernel(const __grid_constant__ DrawParams p) {
    kzn_pdl_wait();
    constexpr uint32_t kFull = 0xffffffffu;
    __shared__ uint32_t tables[kSortWarps][256];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // bin_fill_kernel was the last to need the allocator of extra records: zero for the next draw
        p.work_counter[5] = p.work_counter[4];
        p.work_counter[4] = 0u;
        p.work_counter[6] = 0u;
        p.work_counter[8] = 0u;  // band_prefilter_kernel's list of blocks and their triangle count
        p.work_counter[9] = 0u;
    }
    const uint32_t tile = blockIdx.x * kSortWarps + warp;
    if (tile >= p.tiles_x * p.tiles_y) return;
    const uint32_t tile_x = tile % p.tiles_x, tile_y = tile / p.tiles_x;
    const int tile_px0 = (int)(p.tile_x0 + tile_x) * kTileW, tile_py0 = (int)(p.tile_y0 + tile_y) * kTileH;
    const uint32_t beg = p.tile_offsets[tile], end = p.tile_offsets[tile + 1], n = end - beg;
    if (n == 0) return;
    if (p.stats && lane == 0) atomicAdd(&p.stats[6], 1ull);
    if (end > p.bin_capacity) {  // the bins overflowed: raster_kernel walks every triangle of the draw for this tile
        if (p.stats && lane == 0) atomicAdd(&p.stats[7], 1ull);
        return;
    }
    sort_tile_in_global(p, tile, beg, n, tables[warp], lane, tile_px0, tile_py0);
}

// The same with one CTA per tile and the lists in shared memory (per-warp slices, a [digit][warp] table orders batches and
// warps, so every pass is stable): the better form when tiles are few and their lists long (a warp alone walks a list of a
// thousand ids in 35 dependent steps per phase).
__global__ void __launch_bounds__(kSortThreads) tile_sort_cta_kernel(const __grid_constant__ DrawParams p) {
    kzn_pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortShared& sh = *reinterpret_cast<SortShared*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t tile = blockIdx.y * p.tiles_x + blockIdx.x;
    if (tile == 0 && tid == 0) {  // (as in tile_sort_warp_kernel)
        p.work_counter[5] = p.work_counter[4];
        p.work_counter[4] = 0u;
        p.work_counter[6] = 0u;
        p.work_counter[8] = 0u;  // band_prefilter_kernel's list of blocks and their triangle count
        p.work_counter[9] = 0u;
    }
    const int tile_px0 = (int)(p.tile_x0 + blockIdx.x) * kTileW, tile_py0 = (int)(p.tile_y0 + blockIdx.y) * kTileH;
    const uint32_t beg = p.tile_offsets[tile], end = p.tile_offsets[tile + 1], list_len = end - beg;
    if (list_len == 0) return;
    if (p.stats && tid == 0) atomicAdd(&p.stats[6], 1ull);
    if (end > p.bin_capacity) {  // the bins overflowed: raster_kernel walks every triangle of the draw for this tile
        if (p.stats && tid == 0) atomicAdd(&p.stats[7], 1ull);
        return;
    }
    if (list_len > (uint32_t)kSortCap) {  // too long for shared memory: one warp sorts it in global memory
        if (tid < 32u) sort_tile_in_global(p, tile, beg, list_len, sh.table, tid, tile_px0, tile_py0);
        return;
    }
    for (uint32_t i = tid; i < list_len; i += kSortThreads) sh.list[i] = p.bin_entries[beg + i];
    __syncthreads();
    const uint32_t* sorted = sort_tile_list(sh, list_len);
    uint8_t* masks = tile_masks(p, tile, beg);
    for (uint32_t i = tid; i < list_len; i += kSortThreads) {
        const uint32_t id = sorted[i];
        p.bin_entries[beg + i] = id;
        masks[i] = (uint8_t)brick_mask_of(__ldg(&p.boxes[entry_record(p, id)]), tile_px0, tile_py0);
    }
}

#endif  // !KZN_FS_TEMPLATE

// Persistent, independent warps (kRasterWarps per CTA share nothing but the UNORM8 table): a warp takes one 16 x 16 brick
// at a time from a global work counter — unit u = brick u % 8 of tile u / 8 — until none are left, so no warp waits for
// the slowest brick of its tile and the last wave of a frame is bricks, not tiles.  The brick's colour and depth live in
// the warp's shared memory for the whole triangle list (each framebuffer byte read at most once, written at most once),
// and all its fragments are applied by this warp alone, in primitive order.  Bricks rather than rows: a small triangle
// meets one or two bricks but three or four rows, and everything that is done per (triangle, warp) — record copy, edge
// derivation, prefix sums — is paid that many times.
template <uint32_t FS, int NEED_DST>
__device__ __forceinline__ void raster_body(const DrawParams& p) {
    kzn_pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RasterShared& sh = *reinterpret_cast<RasterShared*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    BrickState& bs = sh.brick[warp];
    const float* lut = sh.unorm_lut;
    const bool has_color = p.color_base != 0, has_depth = p.depth_base != 0;
    const bool depth_test = has_depth && p.st.depth_test_enable;
    const bool depth_write = depth_test && p.st.depth_write_enable;
    if (NEED_DST) {
        for (uint32_t i = tid; i < 256u; i += kRasterThreads) sh.unorm_lut[i] = unorm8_to_float(i);
        __syncthreads();
    }
    const uint32_t n_units = p.tiles_x * p.tiles_y * (uint32_t)kTileBricks;
    const TriRecC* recs_g = reinterpret_cast<const TriRecC*>(p.tris);
    const uint32_t recs_s = (uint32_t)__cvta_generic_to_shared(bs.recs);
    constexpr uint32_t kQ = sizeof(TriRecC) / 16;  // 16-byte vectors per record
    uint32_t n_covered = 0, n_passed = 0;

    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(p.work_counter + 2, 1u);  // slot 2: zeroed by setup_kernel<true> (slots 0, 1 belong to the direct kernel)
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        const uint32_t tile = unit / kTileBricks, brick = unit % kTileBricks;
        const uint32_t tile_x = tile % p.tiles_x, tile_y = tile / p.tiles_x;
        const int tile_px0 = (int)(p.tile_x0 + tile_x) * kTileW, tile_py0 = (int)(p.tile_y0 + tile_y) * kTileH;
        const int brick_px0 = tile_px0 + (int)(brick & 3u) * kBrick, brick_py0 = tile_py0 + (int)(brick >> 2) * kBrick;
        if (!(brick_px0 <= p.cx1 && brick_px0 + kBrick - 1 >= p.cx0 && brick_py0 <= p.cy1 && brick_py0 + kBrick - 1 >= p.cy0)) continue;
        // triangle source: the tile's sorted id list, or — a tile whose list did not fit — every triangle of the draw
        const uint32_t beg = p.tile_offsets[tile], end = p.tile_offsets[tile + 1];
        if (end == beg && !p.load_clear) continue;  // (with LOAD_OP_CLEAR an empty brick is still written: the clear values)
        const bool direct = end > p.bin_capacity;  // the bins overflowed (more than 8 tiles per triangle on average): no list for this tile
        // (walking all triangles: with clipped primitives in the draw every primitive stands for eight list entries, its own
        // record and up to five extra ones, so that the extra triangles come at the primitive's place)
        const uint32_t walk_shift = direct && p.work_counter[5] ? 3u : 0u;
        const uint32_t n_source = direct ? p.total_tris << walk_shift : end - beg;

        // ---- brick prologue: lane l moves the 8 pixels 8 (l & 1) .. + 7 of row l >> 1, two 16-byte vectors when the row
        // pitch allows it: destination values the result depends on, dirty flags ----
        const int qx0 = brick_px0 + (int)(lane & 1u) * kBrickPx, qy = brick_py0 + (int)(lane >> 1);
        const uint32_t qat = lane * kBrickPx;
        const bool vec_ok = (p.fb_w & 3u) == 0 && qx0 + kBrickPx <= (int)p.fb_w;
        const bool cvec = vec_ok && (p.color_base & 15ull) == 0, dvec = vec_ok && (p.depth_base & 15ull) == 0;
        uint32_t quad_clip = 0;
#pragma unroll
        for (int k = 0; k < kBrickPx; ++k)
            if (qy >= p.cy0 && qy <= p.cy1 && qx0 + k >= p.cx0 && qx0 + k <= p.cx1) quad_clip |= 1u << k;
        unsigned long long caddr = 0, daddr = 0;
        if (quad_clip) {
            if (has_color) caddr = color_addr(p, qx0, qy);
            if (has_depth) daddr = p.depth_base + ((unsigned long long)qy * p.fb_w + qx0) * 4ull;
        }
        if (p.load_clear & 1u) {
            const uint32_t w = p.clear_color;
            *reinterpret_cast<uint4*>(&bs.color[qat]) = make_uint4(w, w, w, w);
            *reinterpret_cast<uint4*>(&bs.color[qat + 4]) = make_uint4(w, w, w, w);
        } else if (NEED_DST && quad_clip && has_color) {
#pragma unroll
            for (int h = 0; h < kBrickPx / 4; ++h) {
                const uint32_t m4 = quad_clip >> (4 * h) & 15u;
                if (m4 == 15u && cvec) {
                    *reinterpret_cast<uint4*>(&bs.color[qat + 4 * h]) = *reinterpret_cast<const uint4*>(caddr + 16 * h);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (m4 >> k & 1u) bs.color[qat + 4 * h + k] = *reinterpret_cast<const uint32_t*>(caddr + 16 * h + 4 * k);
                }
            }
        }
        if (p.load_clear & 2u) {
            const float z = p.clear_depth;
            *reinterpret_cast<float4*>(&bs.depth[qat]) = make_float4(z, z, z, z);
            *reinterpret_cast<float4*>(&bs.depth[qat + 4]) = make_float4(z, z, z, z);
        } else if (depth_test && quad_clip) {
#pragma unroll
            for (int h = 0; h < kBrickPx / 4; ++h) {
                const uint32_t m4 = quad_clip >> (4 * h) & 15u;
                if (m4 == 15u && dvec) {
                    *reinterpret_cast<float4*>(&bs.depth[qat + 4 * h]) = *reinterpret_cast<const float4*>(daddr + 16 * h);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (m4 >> k & 1u) bs.depth[qat + 4 * h + k] = *reinterpret_cast<const float*>(daddr + 16 * h + 4 * k);
                }
            }
        }
        *reinterpret_cast<uint2*>(&bs.color_dirty[qat]) = make_uint2(0u, 0u);  // my 8 pixels' flags
        *reinterpret_cast<uint2*>(&bs.depth_dirty[qat]) = make_uint2(0u, 0u);
        __syncwarp();

        // ---- the triangle list: lanes test 4 x 32 list entries per step (one byte each; without a sorted list the packed
        // pixel boxes themselves); the ids that meet my brick collect in a pending list, in list order, and are rasterised
        // 32 at a time, so chunks are full whatever the density.  A chunk's 32 compact records are copied to the warp's
        // slots in shared memory first (one L2 round trip per chunk, 6 x 16 bytes per lane).
        {
            uint32_t* pend = bs.pend;
            uint32_t pend_n = 0;
            const uint32_t* ids = p.bin_entries + beg;
            const uint8_t* masks = tile_masks(p, tile, beg);
            const uint32_t n_steps = (n_source + 127u) / 128u;
            uint32_t step = 0;
            for (;;) {
                while (pend_n < 32u && step < n_steps) {
                    const uint32_t i0 = step * 128u + lane * 4u;  // my four entries: i0 .. i0 + 3
                    uint32_t hits = 0;
                    if (!direct) {
                        uint32_t m4 = 0;
                        if (i0 + 3u < n_source) {
                            m4 = *reinterpret_cast<const uint32_t*>(masks + i0);
                        } else {
#pragma unroll
                            for (uint32_t k = 0; k < 4u; ++k)
                                if (i0 + k < n_source) m4 |= (uint32_t)masks[i0 + k] << (8u * k);
                        }
                        m4 >>= brick;
                        hits = (m4 & 1u) | (m4 >> 7 & 2u) | (m4 >> 14 & 4u) | (m4 >> 21 & 8u);
                    } else {
#pragma unroll
                        for (uint32_t k = 0; k < 4u; ++k) {
                            if (i0 + k >= n_source) break;
                            const uint32_t rec = (i0 + k) >> walk_shift, j = (i0 + k) & ((1u << walk_shift) - 1u);
                            uint2 b;
                            if (j) {  // extra record j - 1 of a clipped primitive, if it has that many
                                const uint32_t prim = rec % p.tris_per_instance;
                                const uint32_t link = p.clip_bits[prim >> 5] >> (prim & 31u) & 1u ? p.clip_link[prim] : 0u;
                                if (j > link >> 29) continue;
                                b = __ldg(&p.boxes[p.total_tris + (link & 0x1fffffffu) + j - 1u]);
                            } else {
                                if (p.block_flags && !p.block_flags[rec >> 8]) continue;  // (prefiltered block: its boxes were never written)
                                b = __ldg(&p.boxes[rec]);
                            }
                            if (brick_mask_of(b, tile_px0, tile_py0) >> brick & 1u) hits |= 1u << k;
                        }
                    }
                    // ordered append: entry (lane, k) goes after all hits of lower lanes and my own lower k
                    uint32_t incl = __popc(hits);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                        if ((int)lane >= d) incl += y;
                    }
                    uint32_t at = pend_n + incl - __popc(hits);
#pragma unroll
                    for (uint32_t k = 0; k < 4u; ++k)
                        if (hits >> k & 1u) pend[at++] = entry_record(p, direct ? (i0 + k) << (3u - walk_shift) : ids[i0 + k]);
                    pend_n += __shfl_sync(0xffffffffu, incl, 31);
                    ++step;
                }
                if (pend_n == 0) break;
                const uint32_t n = min(pend_n, 32u);
                __syncwarp();
                // the chunk's records: global -> my warp's slots, asynchronous copies (LDGSTS, no registers)
#pragma unroll
                for (uint32_t i = 0; i < kQ; ++i) {
                    const uint32_t q = lane + 32u * i, r = q / kQ;
                    if (r < n) {
                        const uint4* src = reinterpret_cast<const uint4*>(&recs_g[pend[r]]) + (q - kQ * r);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(recs_s + q * 16u), "l"(src) : "memory");
                    }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                raster_chunk<FS, NEED_DST>(p, bs, lut, lane, n, brick_px0, brick_py0, depth_test, depth_write, has_color, n_covered, n_passed);
                // move the remainder (< 128 + 32 entries) to the front
                const uint32_t rest = pend_n - n;
                uint32_t v[5];
#pragma unroll
                for (uint32_t k = 0; k < 5u; ++k) v[k] = lane + 32u * k < rest ? pend[n + lane + 32u * k] : 0u;
                __syncwarp();
#pragma unroll
                for (uint32_t k = 0; k < 5u; ++k)
                    if (lane + 32u * k < rest) pend[lane + 32u * k] = v[k];
                pend_n = rest;
                __syncwarp();
            }
        }

        // ---- brick epilogue: each lane writes back its 8 pixels where they changed ----
        __syncwarp();
        if (quad_clip) {
            const uint2 cdb = *reinterpret_cast<const uint2*>(&bs.color_dirty[qat]);
            const uint2 ddb = *reinterpret_cast<const uint2*>(&bs.depth_dirty[qat]);
            // (LOAD_OP_CLEAR: every pixel of the brick inside the clip rectangle is written, changed or not)
            const uint32_t cd = ((p.load_clear & 1u) ? 0xffu : bytes_to_bits(cdb.x) | bytes_to_bits(cdb.y) << 4) & quad_clip;
            const uint32_t dd = ((p.load_clear & 2u) ? 0xffu : bytes_to_bits(ddb.x) | bytes_to_bits(ddb.y) << 4) & quad_clip;
#pragma unroll
            for (int h = 0; h < kBrickPx / 4; ++h) {
                const uint32_t c4 = cd >> (4 * h) & 15u, d4 = dd >> (4 * h) & 15u;
                if (c4 == 15u && cvec) {
                    *reinterpret_cast<uint4*>(caddr + 16 * h) = *reinterpret_cast<const uint4*>(&bs.color[qat + 4 * h]);
                } else if (c4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (c4 >> k & 1u) *reinterpret_cast<uint32_t*>(caddr + 16 * h + 4 * k) = bs.color[qat + 4 * h + k];
                }
                if (d4 == 15u && dvec) {
                    *reinterpret_cast<float4*>(daddr + 16 * h) = *reinterpret_cast<const float4*>(&bs.depth[qat + 4 * h]);
                } else if (d4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (d4 >> k & 1u) *reinterpret_cast<float*>(daddr + 16 * h + 4 * k) = bs.depth[qat + 4 * h + k];
                }
            }
        }
        __syncwarp();
    }
    if (p.stats) {
        // warp-reduce then one atomic per warp
        for (int d = 16; d > 0; d >>= 1) {
            n_covered += __shfl_xor_sync(0xffffffffu, n_covered, d);
            n_passed += __shfl_xor_sync(0xffffffffu, n_passed, d);
        }
        if (lane == 0) {
            if (n_covered) atomicAdd(&p.stats[4], (unsigned long long)n_covered);
            if (n_passed) atomicAdd(&p.stats[5], (unsigned long long)n_passed);
        }
    }
}

#ifndef KZN_FS_TEMPLATE
template <uint32_t FS, int NEED_DST>
__global__ void __launch_bounds__(32, kDirectWarpsPerSM) raster_direct_kernel(const __grid_constant__ DrawParams p) {
    raster_direct_body<FS, NEED_DST>(p);
}
template <uint32_t FS, int NEED_DST>
__global__ void __launch_bounds__(kRasterThreads, KZN_RASTER_CTAS) raster_kernel(const __grid_constant__ DrawParams p) {
    raster_body<FS, NEED_DST>(p);
}
#endif

}  // namespace

#ifdef KZN_FS_TEMPLATE
// The PTX templates of the two raster kernels for compiled fragment programs, one per destination mode (see
// run_fragment_program): extern "C" names, so that backend.cpp finds them in the module it builds.
#define KZN_FS_TEMPLATE_KERNELS(MODE)                                                                                              \
    extern "C" __global__ void __launch_bounds__(32, kDirectWarpsPerSM) kzn_fs_raster_direct_##MODE(const __grid_constant__ DrawParams p) { \
        raster_direct_body<KZN_FS_COMPILED, MODE>(p);                                                                              \
    }                                                                                                                              \
    extern "C" __global__ void __launch_bounds__(kRasterThreads, KZN_RASTER_CTAS) kzn_fs_raster_##MODE(const __grid_constant__ DrawParams p) { \
        raster_body<KZN_FS_COMPILED, MODE>(p);                                                                                     \
    }
KZN_FS_TEMPLATE_KERNELS(0)
KZN_FS_TEMPLATE_KERNELS(1)
KZN_FS_TEMPLATE_KERNELS(2)
#else  // everything below is host code and the kernels with built-in fragment programs

extern "C" const uint32_t kzn_draw_params_push_constants_offset = (uint32_t)offsetof(DrawParams, push_constants);

// ---- scratch --------------------------------------------------------------------------------------------------------------

struct KznRasterScratch {
    TriRec* tris = nullptr;
    size_t tri_capacity = 0;
    uint2* boxes = nullptr;  // packed pixel boxes, one per record
    size_t box_capacity = 0;
    uint32_t* tile_words = nullptr;  // counts | offsets | cursor
    size_t tile_capacity = 0;
    uint32_t* bin_entries = nullptr;
    size_t bin_capacity = 0;
    uint8_t* bin_masks = nullptr;
    size_t mask_capacity = 0;
    uint32_t* bin_entries_alt = nullptr;
    size_t alt_capacity = 0;
    uint32_t* counters = nullptr;  // [8]: the work_counter slots (see kFusedMax)
    size_t counter_capacity = 0;
    uint32_t* clip_link = nullptr;  // clipped primitives of binned draws: see DrawParams
    size_t link_capacity = 0;
    uint32_t* clip_parent = nullptr;
    size_t parent_capacity = 0;
    uint32_t* clip_list = nullptr;
    size_t list_capacity = 0;
    uint8_t* block_flags = nullptr;  // band_prefilter_kernel: one byte per block of 256 triangles
    size_t flag_capacity = 0;
    uint32_t* live_blocks = nullptr;  // ... and the list of the blocks whose flag is set
    size_t live_capacity = 0;
    uint8_t* band_codes = nullptr;   // band_code_kernel: one byte per vertex (+ 1)
    size_t code_capacity = 0;
    unsigned char* vs_out = nullptr;  // what a compiled vertex stage writes and setup reads (KZN_VS_COMPILED)
    size_t vs_capacity = 0;
    bool attrs_set = false;
    void* window_base = nullptr;  // the access policy window last set on the queue's stream (the records of binned draws)
    size_t window_bytes = 0;
    // Executable command buffers that recorded a draw hold the buffer addresses of that time in their CUDA graphs: while any
    // such command buffer exists (pins > 0) a buffer that is outgrown is retired instead of freed, and freed when the last
    // of them is destroyed.
    int pins = 0;
    std::vector<void*> retired;
};

void kzn_raster_scratch_pin(KznRasterScratch* s) {
    if (s) ++s->pins;
}
void kzn_raster_scratch_unpin(KznRasterScratch* s) {
    if (!s || --s->pins > 0) return;
    for (void* p : s->retired) cudaFree(p);
    s->retired.clear();
}

void kzn_raster_scratch_destroy(KznRasterScratch* s) {
    if (!s) return;
    for (void* p : s->retired) cudaFree(p);
    cudaFree(s->tris);
    cudaFree(s->boxes);
    cudaFree(s->tile_words);
    cudaFree(s->bin_entries);
    cudaFree(s->bin_masks);
    cudaFree(s->bin_entries_alt);
    cudaFree(s->counters);
    cudaFree(s->clip_link);
    cudaFree(s->clip_parent);
    cudaFree(s->clip_list);
    cudaFree(s->block_flags);
    cudaFree(s->band_codes);
    cudaFree(s->live_blocks);
    cudaFree(s->vs_out);
    delete s;
}

namespace {

// Scratch grows monotonically; growth happens before any kernel of this draw is enqueued and the old buffer
// is freed with cudaFree, which synchronises with work still using it — unless an executable command buffer still names
// it (KznRasterScratch::pins): then it is retired and outlives that command buffer.
template <typename T>
KznResult grow(const KznQueue* q, T** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return KZN_SUCCESS;
    KZN_VALID(!q->recording, "draw needs more raster scratch than is reserved and allocation cannot be recorded: call "
                             "kzn_queue_reserve_draw(queue, triangles, width, height) before vkBeginCommandBuffer");
    size_t n = need + need / 2;
    T* fresh = nullptr;
    KZN_CUDA(cudaMalloc(&fresh, n * sizeof(T)));
    if (*ptr) {
        if (q->raster && q->raster->pins > 0) q->raster->retired.push_back(*ptr);
        else cudaFree(*ptr);
    }
    *ptr = fresh;
    *cap = n;
    return KZN_SUCCESS;
}

bool finite_host(float f) { return (f - f) == 0.0f; }

KznResult check_draw(const KznGraphicsPipelineState* st, const KznFramebuffer* fb, const KznViewport* vp, const KznRect2D* sc,
                     const KznVertexInput* vi, uint32_t* W, uint32_t* H) {
    KZN_VALID(st && fb && vp && sc && vi, "draw: null argument");
    KZN_REQUIRE(st->vertex_program <= KZN_VS_COMPILED, KZN_ERROR_FEATURE_NOT_PRESENT, "draw: unknown vertex program %u", st->vertex_program);
    KZN_REQUIRE(st->fragment_program <= KZN_FS_COMPILED, KZN_ERROR_FEATURE_NOT_PRESENT, "draw: unknown fragment program %u",
                st->fragment_program);
    KZN_VALID(st->cull_mode <= 3 && st->front_face <= 1 && st->depth_compare_op <= 7, "draw: bad rasterization / depth state");
    KZN_REQUIRE(st->src_color_blend_factor <= 9 && st->dst_color_blend_factor <= 9 && st->src_alpha_blend_factor <= 9 &&
                    st->dst_alpha_blend_factor <= 9,
                KZN_ERROR_FEATURE_NOT_PRESENT, "draw: blend factor not supported (dualSrcBlend is FALSE, api_impl.rs:883)");
    KZN_VALID(st->color_blend_op <= 4 && st->alpha_blend_op <= 4, "draw: bad blend op");
    KZN_VALID(st->logic_op <= 16, "draw: logic_op is 0 (disabled) or 1 + VkLogicOp");
    KZN_VALID(st->topology <= 10 || st->topology == KZN_TOPOLOGY_POINT_LIST, "draw: bad primitive topology");
    KZN_REQUIRE(st->topology <= KZN_TOPOLOGY_TRIANGLE_FAN || st->topology == KZN_TOPOLOGY_POINT_LIST, KZN_ERROR_FEATURE_NOT_PRESENT,
                "draw: adjacency and patch topologies are not rasterised");
    bool has_color = fb->color.format != 0, has_depth = fb->depth.format != 0;
    KZN_VALID(has_color || has_depth, "draw: framebuffer has no attachment");
    if (has_color) {
        KZN_REQUIRE(fb->color.format == KZN_FORMAT_R8G8B8A8_UNORM, KZN_ERROR_FORMAT_NOT_SUPPORTED, "draw: colour attachment format %u",
                    fb->color.format);
        KZN_VALID(fb->color.mip_levels == 1 && fb->color.depth == 1 && fb->color.tiling <= 1, "draw: bad colour attachment");
        KZN_VALID(fb->color_layer < fb->color.array_layers, "draw: colour layer out of range");
        KZN_VALID(fb->color.address && fb->color.address % 4 == 0, "draw: colour attachment has no memory");
        *W = fb->color.width;
        *H = fb->color.height;
    }
    if (has_depth) {
        KZN_REQUIRE(fb->depth.format == KZN_FORMAT_D32_SFLOAT, KZN_ERROR_FORMAT_NOT_SUPPORTED, "draw: depth attachment format %u",
                    fb->depth.format);
        KZN_VALID(fb->depth.mip_levels == 1 && fb->depth.depth == 1 && fb->depth.tiling == 0, "draw: bad depth attachment");
        KZN_VALID(fb->depth_layer < fb->depth.array_layers, "draw: depth layer out of range");
        KZN_VALID(fb->depth.address && fb->depth.address % 4 == 0, "draw: depth attachment has no memory");
        KZN_VALID(!has_color || (fb->depth.width == fb->color.width && fb->depth.height == fb->color.height),
                  "draw: attachment sizes differ");
        *W = fb->depth.width;
        *H = fb->depth.height;
    }
    KZN_VALID(!fb->band_y1 || (fb->band_y0 < fb->band_y1 && fb->band_y1 <= *H), "draw: bad band");
    KZN_VALID(sc->offset.x >= 0 && sc->offset.y >= 0, "draw: negative scissor offset (VUID 00598)");
    KZN_VALID(vp->width > 0.0f && vp->height != 0.0f && finite_host(vp->height), "draw: bad viewport size");
    // viewportBoundsRange: everything below converts these to integers, which is undefined for NaN / infinity / huge values
    KZN_VALID(finite_host(vp->x) && finite_host(vp->y) && finite_host(vp->width) && fabsf(vp->x) <= 1048576.0f &&
                  fabsf(vp->y) <= 1048576.0f && vp->width <= 1048576.0f && fabsf(vp->height) <= 1048576.0f,
              "draw: viewport outside the +-2^20 bounds range");
    KZN_VALID(vp->minDepth >= 0.0f && vp->minDepth <= 1.0f && vp->maxDepth >= 0.0f && vp->maxDepth <= 1.0f, "draw: bad viewport depth range");
    KZN_VALID(vi->position_components == 3 || vi->position_components == 4, "draw: position must have 3 or 4 components");
    KZN_VALID(vi->stride % 4 == 0 && vi->position_offset % 4 == 0 && (vi->color_offset == ~0u || vi->color_offset % 4 == 0),
              "draw: vertex layout must be 4-byte aligned");
    KZN_VALID(vi->buffer_address && vi->buffer_address % 4 == 0, "draw: bad vertex buffer address");
    // index (< 2^32 + |vertexOffset|) x stride must stay far from 2^64: the robust-fetch bounds test works on that product
    KZN_VALID(vi->stride < (1u << 28), "draw: vertex stride of 2^28 bytes or more");
    return KZN_SUCCESS;
}

void make_clip(const KznFramebuffer* fb, const KznViewport* vp, const KznRect2D* sc, uint32_t W, uint32_t H, DrawParams* p) {
    float vx0 = vp->x, vx1 = vp->x + vp->width;
    float vy0 = vp->height < 0.0f ? vp->y + vp->height : vp->y;
    float vy1 = vp->height < 0.0f ? vp->y : vp->y + vp->height;
    long long x0 = (long long)ceilf(vx0 - 0.5f), x1 = (long long)floorf(vx1 - 0.5f);
    long long y0 = (long long)ceilf(vy0 - 0.5f), y1 = (long long)floorf(vy1 - 0.5f);
    long long sx1 = (long long)sc->offset.x + sc->extent.width - 1, sy1 = (long long)sc->offset.y + sc->extent.height - 1;
    if (x0 < sc->offset.x) x0 = sc->offset.x;
    if (y0 < sc->offset.y) y0 = sc->offset.y;
    if (x1 > sx1) x1 = sx1;
    if (y1 > sy1) y1 = sy1;
    long long by0 = fb->band_y1 ? fb->band_y0 : 0, by1 = fb->band_y1 ? fb->band_y1 : H;
    if (x0 < 0) x0 = 0;
    if (y0 < by0) y0 = by0;
    if (x1 > (long long)W - 1) x1 = (long long)W - 1;
    if (y1 > by1 - 1) y1 = by1 - 1;
    if (x1 < x0 || y1 < y0) { x0 = y0 = 0; x1 = y1 = -1; }
    p->cx0 = (int32_t)x0; p->cy0 = (int32_t)y0; p->cx1 = (int32_t)x1; p->cy1 = (int32_t)y1;
    p->zclip_lo = vp->minDepth < vp->maxDepth ? vp->minDepth : vp->maxDepth;
    p->zclip_hi = vp->minDepth < vp->maxDepth ? vp->maxDepth : vp->minDepth;
}

// extra records a binned draw of `per_instance` triangles can give its clipped primitives (R3): five each (the most a
// polygon needs) up to 16384 primitives, one each on average beyond that (KznDrawStatistics::clip_overflow counts the rest)
size_t clip_reserve(size_t per_instance) { return per_instance * 5 < per_instance + 65536 ? per_instance * 5 : per_instance + 65536; }

// the queue's raster scratch, grown to hold the records of `total` triangles (`per_instance` of them distinct) and (binned
// draws) the tile tables
KznResult ensure_scratch(KznQueue* q, size_t total, size_t per_instance, size_t tiles, bool binned) {
    KznResult r;
    if (!q->raster) {
        q->raster = new (std::nothrow) KznRasterScratch();
        KZN_REQUIRE(q->raster, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    }
    KznRasterScratch* s = q->raster;
    // direct draws: expanded records, up to six per primitive; binned draws: compact records, the extra ones after the draw's
    const size_t extras = binned ? clip_reserve(per_instance) : 0;
    const size_t tri_slots = binned ? ((total + extras) * sizeof(TriRecC) + sizeof(TriRec) - 1) / sizeof(TriRec) : total * 6;
    if ((r = grow(q, &s->tris, &s->tri_capacity, tri_slots))) return r;
    if ((r = grow(q, &s->boxes, &s->box_capacity, total + extras))) return r;
    if (binned) {
        if ((r = grow(q, &s->clip_link, &s->link_capacity, per_instance))) return r;
        if ((r = grow(q, &s->clip_parent, &s->parent_capacity, extras))) return r;
        if ((r = grow(q, &s->clip_list, &s->list_capacity, per_instance))) return r;
        if ((r = grow(q, &s->block_flags, &s->flag_capacity, (per_instance + 255) / 256))) return r;
        if ((r = grow(q, &s->live_blocks, &s->live_capacity, (per_instance + 255) / 256))) return r;
    }
    if (!s->counters) {  // the dispensers and the allocator of extra records start at zero and are handed back at zero by their users
        if ((r = grow(q, &s->counters, &s->counter_capacity, (size_t)32))) return r;
        KZN_CUDA(cudaMemsetAsync(s->counters, 0, s->counter_capacity * sizeof(uint32_t), q->stream));  // ordered before the first draw kernel of this queue
        static const float one = 1.0f;  // words 16..19: the vertex (0, 0, 0, 1) out-of-range positions read
        KZN_CUDA(cudaMemcpyAsync(s->counters + 19, &one, sizeof one, cudaMemcpyHostToDevice, q->stream));
    }
    if (binned) {
        if ((r = grow(q, &s->tile_words, &s->tile_capacity, 3 * tiles + 1 + (per_instance + 31) / 32))) return r;
        // capacity for the (tile, triangle) pairs: the exact total is only known on the device; tiles whose
        // segment does not fit fall back to walking the whole triangle list (still exact, only slower)
        size_t want = total * 8 + tiles;
        if (want > 0xfffffff0ull) want = 0xfffffff0ull;
        if ((r = grow(q, &s->bin_entries, &s->bin_capacity, want))) return r;
        if ((r = grow(q, &s->bin_masks, &s->mask_capacity, s->bin_capacity + 4 * tiles + 8))) return r;
        if ((r = grow(q, &s->bin_entries_alt, &s->alt_capacity, s->bin_capacity))) return r;
    }
    if (!s->attrs_set) {
        const int bytes = (int)sizeof(RasterShared);
        KZN_CUDA(cudaFuncSetAttribute(raster_kernel<KZN_FS_VERTEX_COLOR, kDstNone>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        KZN_CUDA(cudaFuncSetAttribute(raster_kernel<KZN_FS_VERTEX_COLOR, kDstGeneric>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        KZN_CUDA(cudaFuncSetAttribute(raster_kernel<KZN_FS_VERTEX_COLOR, kDstAlphaOver>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        KZN_CUDA(cudaFuncSetAttribute(raster_kernel<KZN_FS_PROCEDURAL_ORBIT, kDstNone>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        KZN_CUDA(cudaFuncSetAttribute(raster_kernel<KZN_FS_PROCEDURAL_ORBIT, kDstGeneric>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        KZN_CUDA(cudaFuncSetAttribute(raster_kernel<KZN_FS_PROCEDURAL_ORBIT, kDstAlphaOver>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        KZN_CUDA(cudaFuncSetAttribute(tile_sort_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortShared)));
        s->attrs_set = true;
    }
    return KZN_SUCCESS;
}

// a draw whose clip rectangle is empty rasterises nothing, but its triangles still count as submitted and culled
__global__ void count_unrasterised_kernel(unsigned long long* stats, unsigned long long triangles) {
    kzn_pdl_wait();
    atomicAdd(&stats[0], triangles);
    atomicAdd(&stats[1], triangles);
}

// the kernels of a draw are launched with programmatic dependent launch (kzn_internal.h)
void launch_draw_kernel(void (*kernel)(DrawParams), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const DrawParams& p) {
    kzn_launch_pdl(kernel, grid, block, smem, stream, p);
}

KznResult draw_common(KznQueue* q, const KznGraphicsPipelineState* st, const KznFramebuffer* fb, const KznViewport* vp,
                      const KznRect2D* sc, const KznVertexInput* vi, uint64_t index_base, uint32_t index_type, uint32_t first,
                      uint32_t count, int32_t vertex_offset, uint32_t instance_count) {
    KznResult r = kzn_queue_enter_draw(q);  // (a pending LOAD_OP_CLEAR stays pending: this draw may take it over, below)
    if (r) return r;
    uint32_t W = 0, H = 0;
    if ((r = check_draw(st, fb, vp, sc, vi, &W, &H))) return r;
    KznVertexInput shaded;
    if (st->vertex_program == KZN_VS_COMPILED && (r = kzn_flush_render_pass_clear(q))) return r;  // (the vertex stage is a dispatch of its own)
    if (st->vertex_program == KZN_VS_COMPILED && count && instance_count) {
        // The vertex stage as a dispatch of the PTX back end, one invocation per vertex, into the queue's vertex scratch; the
        // fixed stages then read positions and colours from there (same vertex numbering, so indices and vertexOffset apply
        // unchanged).  Setup waits for it like for any earlier kernel of the stream (griddepcontrol.wait).
        KZN_VALID(q->vertex_program, "draw: vertex_program is KZN_VS_COMPILED but no program is bound (kzn_cmd_bind_vertex_program)");
        const KznVertexLayout lay = kzn_vertex_program_layout(q->vertex_program);
        KZN_VALID(vi->stride == lay.inputs_size, "draw: the vertex binding's stride (%u) is not the vertex program's input block size (%llu)", vi->stride,
                  (unsigned long long)lay.inputs_size);
        const uint64_t in_buffer = vi->buffer_size / vi->stride;  // whole vertices of the binding
        uint64_t v0 = 0, n = in_buffer;
        if (!index_base) {  // a non-indexed draw names vertices first .. first + count - 1
            v0 = first < in_buffer ? first : in_buffer;
            n = (uint64_t)first + count < in_buffer ? count : in_buffer - v0;
        }
        uint64_t out = 0;
        if ((r = kzn_raster_vertex_scratch(q, (v0 + n) * lay.outputs_size, &out))) return r;
        if ((r = kzn_vertex_program_run(q, q->vertex_program, vi->buffer_address, out, v0, n))) return r;
        shaded.buffer_address = out;
        shaded.buffer_size = (v0 + n) * lay.outputs_size;  // vertices beyond the binding read (0, 0, 0, 1) / black (robust fetch)
        shaded.stride = (uint32_t)lay.outputs_size;
        shaded.position_offset = lay.position_offset;
        shaded.position_components = 4;
        shaded.color_offset = lay.color_offset;
        vi = &shaded;
    }
    DrawParams p;
    memset(&p, 0, sizeof p);
    make_clip(fb, vp, sc, W, H, &p);
    const uint32_t topo = st->topology ? st->topology : KZN_TOPOLOGY_TRIANGLE_LIST;
    // primitives of an instance (called triangles below; points and lines become two triangles each where clipped polygons are made, R9)
    uint64_t ntri = topo == KZN_TOPOLOGY_POINT_LIST ? count : topo == KZN_TOPOLOGY_LINE_LIST ? count / 2 :
                    topo == KZN_TOPOLOGY_LINE_STRIP ? (count >= 2 ? count - 1 : 0) :
                    topo == KZN_TOPOLOGY_TRIANGLE_LIST ? count / 3 : (count >= 3 ? count - 2 : 0);
    uint64_t total = ntri * instance_count;
    if (total == 0) return kzn_flush_render_pass_clear(q);
    if (p.cx1 < p.cx0 || p.cy1 < p.cy0) {  // empty clip rectangle (scissor, viewport, framebuffer and band do not meet)
        if ((r = kzn_flush_render_pass_clear(q))) return r;
        if (q->stats_address) {
            kzn_launch_pdl(count_unrasterised_kernel, dim3(1), dim3(1), 0, q->stream, reinterpret_cast<unsigned long long*>(q->stats_address),
                           (unsigned long long)total);
            return kzn_launched(q, "count_unrasterised_kernel");
        }
        return KZN_SUCCESS;
    }
    KZN_REQUIRE(p.cx1 <= 65535 && p.cy1 <= 65535, KZN_ERROR_FEATURE_NOT_PRESENT,
                "draw: render area reaches beyond pixel 65535 (pixel boxes are packed into 16 bits per coordinate)");
    KZN_REQUIRE(total <= 0x1fffffffull, KZN_ERROR_FEATURE_NOT_PRESENT, "draw: more than 2^29 triangles in one draw");

    p.vertex_base = vi->buffer_address;
    p.vertex_size = vi->buffer_size;
    p.index_base = index_base;
    p.stride = vi->stride;
    p.pos_off = vi->position_offset;
    p.pos_comps = vi->position_components;
    p.col_off = vi->color_offset;
    p.index_type = index_type;
    p.first = first;
    p.vertex_offset = vertex_offset;
    p.tris_per_instance = (uint32_t)ntri;
    p.total_tris = (uint32_t)total;
    p.instances = instance_count;
    {   // robust fetch as one comparison: vertex v is fetched iff v * stride + offset + bytes <= size
        auto whole = [&](uint64_t offset, uint64_t bytes) -> unsigned long long {
            if (offset + bytes > vi->buffer_size) return 0ull;
            return vi->stride ? (vi->buffer_size - offset - bytes) / vi->stride + 1ull : 1ull << 63;  // (a negative index, as unsigned, is never below it)
        };
        p.pos_vertices = whole(vi->position_offset, vi->position_components * 4ull);
        p.col_vertices = vi->color_offset != ~0u ? whole(vi->color_offset, 16ull) : 0ull;
    }
    p.st = *st;
    p.hw = vp->width * 0.5f;
    p.hh = vp->height * 0.5f;
    p.ox = vp->x + p.hw;
    p.oy = vp->y + p.hh;
    p.dz = vp->maxDepth - vp->minDepth;
    p.zmin_vp = vp->minDepth;
    bool has_color = fb->color.format != 0, has_depth = fb->depth.format != 0;
    if (has_color) {
        p.color_base = fb->color.address + kzn_slice_pitch(&fb->color) * fb->color_layer;
        p.color_tiled = fb->color.tiling == KZN_TILING_TILED;
        p.color_tiles_x = kzn_tiles(fb->color.width);
    }
    if (has_depth) p.depth_base = fb->depth.address + kzn_slice_pitch(&fb->depth) * fb->depth_layer;
    p.fb_w = W;
    p.fb_h = H;
    p.inv_fb_w = 1.0f / (float)W;
    p.inv_fb_h = 1.0f / (float)H;
    p.tile_x0 = (uint32_t)p.cx0 / kTileW;
    p.tile_y0 = (uint32_t)p.cy0 / kTileH;
    p.tiles_x = (uint32_t)p.cx1 / kTileW - p.tile_x0 + 1;
    p.tiles_y = (uint32_t)p.cy1 / kTileH - p.tile_y0 + 1;
    KZN_REQUIRE(p.tiles_y <= 65535u, KZN_ERROR_FEATURE_NOT_PRESENT, "draw: framebuffer taller than 524280 rows");
    p.stats = reinterpret_cast<unsigned long long*>(q->stats_address);
    {
        const double view_rows = fabs((double)vp->height);
        p.rows_first = (double)(p.cy1 - p.cy0 + 1) < 0.5 * view_rows;
    }
    if (st->fragment_program == KZN_FS_COMPILED) {
        KZN_VALID(q->fragment_program, "draw: fragment_program is KZN_FS_COMPILED but no program is bound (kzn_cmd_bind_fragment_program)");
        p.fs_flags = kzn_fragment_program_flags(q->fragment_program);
        p.descriptor_table = q->descriptor_table[0];  // VK_PIPELINE_BIND_POINT_GRAPHICS
        memcpy(p.push_constants, q->push_constants, sizeof p.push_constants);
    }

    size_t tiles = (size_t)p.tiles_x * p.tiles_y;
    p.binned = total > kDirectMax;
    if (q->pending_clear.active) {
        // LOAD_OP_CLEAR of the render pass this draw opens: a binned draw over exactly the cleared rectangle of the same
        // attachments starts its bricks from the clear values (raster_body); anything else has the clear performed first
        const KznPendingClear& c = q->pending_clear;
        const bool same_color = !(c.load & 1u) || (has_color && fb->color.address == c.fb.color.address && fb->color_layer == c.fb.color_layer &&
                                                   fb->color.tiling == c.fb.color.tiling && fb->color.width == c.fb.color.width);
        const bool same_depth = !(c.load & 2u) || (has_depth && fb->depth.address == c.fb.depth.address && fb->depth_layer == c.fb.depth_layer &&
                                                   fb->depth.width == c.fb.depth.width);
        static const bool fuse_enabled = [] { const char* e = getenv("KZN_FUSE_LOAD_CLEAR"); return !(e && e[0] == '0'); }();
        if (fuse_enabled && p.binned && same_color && same_depth && c.x0 == p.cx0 && c.y0 == p.cy0 && c.x1 == p.cx1 && c.y1 == p.cy1) {
            p.load_clear = c.load;
            p.clear_color = c.color_word;
            memcpy(&p.clear_depth, &c.depth_bits, 4);
            q->pending_clear.active = false;
        } else if ((r = kzn_flush_render_pass_clear(q))) {
            return r;
        }
    }
    if ((r = ensure_scratch(q, (size_t)total, (size_t)ntri, tiles, p.binned != 0))) return r;
    if (q->recording) kzn_command_buffer_uses_raster_scratch(q->recording, q->raster);  // its graph will name these buffers
    KznRasterScratch* s = q->raster;
    p.tris = s->tris;
    p.boxes = s->boxes;
    p.work_counter = s->counters;
    p.null_vertex = reinterpret_cast<const float*>(s->counters + 16);
    {
        static const bool exact = [] {
            const char* e = getenv("KZN_RASTER_EXACT_SPANS");
            return !(e && e[0] == '0');
        }();
        p.exact_spans = exact;
    }
    if (p.binned) {
        p.tile_counts = s->tile_words;  // counts | cursor | clip_bits: zeroed together below
        p.tile_cursor = s->tile_words + tiles;
        p.clip_bits = s->tile_words + 2 * tiles;
        p.tile_offsets = p.clip_bits + (ntri + 31) / 32;
        p.bin_entries = s->bin_entries;
        p.bin_masks = s->bin_masks;
        p.bin_entries_alt = s->bin_entries_alt;
        p.bin_capacity = (uint32_t)(s->bin_capacity > 0xfffffff0ull ? 0xfffffff0ull : s->bin_capacity);
        p.clip_link = s->clip_link;
        p.clip_parent = s->clip_parent;
        p.clip_list = s->clip_list;
        p.clip_capacity = (uint32_t)clip_reserve((size_t)ntri);
    }

    if (p.binned && q->device->l2_persist_bytes > 0 && !q->recording) {
        // The compact records are written once by setup_kernel and read once by raster_kernel a few kernels later.  When they
        // (nearly) fit the part of L2 that can be set aside for persisting lines they stay there — the record round trip
        // through HBM is the larger part of the frame's DRAM traffic (1M-triangle frame: 270 -> 180 MB); larger draws are left
        // alone (the window would thrash: the 4M-triangle frame got slower).
        size_t bytes = (size_t)total * sizeof(TriRecC);
        if (bytes > (size_t)q->device->l2_window_bytes || bytes > (size_t)q->device->l2_persist_bytes * 3 / 2) bytes = 0;
        if (s->window_base != (void*)s->tris || s->window_bytes != bytes) {
            cudaStreamAttrValue attr = {};
            attr.accessPolicyWindow.base_ptr = s->tris;
            attr.accessPolicyWindow.num_bytes = bytes;
            const double share = bytes ? (double)q->device->l2_persist_bytes / (double)bytes : 0.0;
            attr.accessPolicyWindow.hitRatio = share < 1.0 ? (float)share : 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(q->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) {
                s->window_base = s->tris;
                s->window_bytes = bytes;
            } else {
                cudaGetLastError();
            }
        }
    }
    if (p.binned) KZN_CUDA(cudaMemsetAsync(p.tile_counts, 0, (2 * tiles + (ntri + 31) / 32) * sizeof(uint32_t), q->stream));
    // (points and lines are never fused: raster_direct_kernel's own setup is compiled for triangles only)
    p.fused_setup = !p.binned && total <= kFusedMax && topo >= KZN_TOPOLOGY_TRIANGLE_LIST && topo <= KZN_TOPOLOGY_TRIANGLE_FAN;
    uint32_t list_grid = 0;
    {
        static const long min_tris = [] { const char* e = getenv("KZN_BAND_PREFILTER_MIN"); return e ? atol(e) : 65536l; }();  // (negative: never)
        static const bool use_codes = [] { const char* e = getenv("KZN_BAND_CODES"); return !(e && e[0] == '0'); }();
        const bool triangles = topo >= KZN_TOPOLOGY_TRIANGLE_LIST && topo <= KZN_TOPOLOGY_TRIANGLE_FAN;
        if (p.binned && p.rows_first && triangles && instance_count == 1 && min_tris >= 0 && (long)ntri >= min_tris) {
            // the code table pays when the vertex buffer is not much larger than what the draw can name (an indexed mesh)
            const bool codes = use_codes && index_base && p.pos_vertices <= 2ull * ntri && p.pos_vertices < 0xffffffffull && !q->recording;
            uint32_t n_codes = 0;
            if (codes) {
                n_codes = (uint32_t)p.pos_vertices;
                if ((r = grow(q, &s->band_codes, &s->code_capacity, (size_t)n_codes + 1))) return r;
                kzn_launch_pdl(band_code_kernel, dim3(n_codes / 256u + 1u), dim3(256), 0, q->stream, p, s->band_codes, n_codes);
                if ((r = kzn_launched(q, "band_code_kernel"))) return r;
            }
            static const bool four = [] { const char* e = getenv("KZN_BAND_PREFILTER4"); return !(e && e[0] == '0'); }();
            if (four && codes && topo == KZN_TOPOLOGY_TRIANGLE_LIST && index_type == KZN_INDEX_TYPE_UINT32 && (index_base + 4ull * first) % 16 == 0)
                kzn_launch_pdl(band_prefilter4_kernel, dim3((uint32_t)((ntri + 1023) / 1024)), dim3(256), 0, q->stream, p, s->block_flags,
                               s->live_blocks, (const uint8_t*)s->band_codes, n_codes);
            else
                kzn_launch_pdl(band_prefilter_kernel, dim3((uint32_t)((ntri + 255) / 256)), dim3(256), 0, q->stream, p, s->block_flags,
                               s->live_blocks, codes ? (const uint8_t*)s->band_codes : (const uint8_t*)nullptr, n_codes);
            if ((r = kzn_launched(q, "band_prefilter_kernel"))) return r;
            p.block_flags = s->block_flags;
            p.live_blocks = s->live_blocks;
            // 2.5 x as many CTAs as the band's share of the viewport's rows suggests blocks on the list, at least 8 per SM: too
            // few make some CTAs loop while others have left (measured: 1.25 x 0.346 ms, 3 x 0.332 ms on a band of the 8K frame),
            // all 15 625 of them mostly find the list's end and cost a launch each (0.350 ms)
            static const double slack = [] { const char* e = getenv("KZN_BAND_LIST_SLACK"); const double v = e ? atof(e) : 2.5; return v > 0.0 ? v : 2.5; }();
            const uint64_t blocks = (ntri + 255) / 256;
            const double share = (double)(p.cy1 - p.cy0 + 1) / fabs((double)vp->height);
            uint64_t want = (uint64_t)((double)blocks * share * slack) + 1, floor_ctas = (uint64_t)q->device->sm_count * 8;
            if (want < floor_ctas) want = floor_ctas;
            list_grid = (uint32_t)(blocks < want ? blocks : want);
        }
    }
    if (!p.fused_setup) {
        // (after the band prefilter the grid is capped: the kernels loop over the list of blocks with work)
        if (p.binned) launch_draw_kernel(setup_kernel<true>, dim3(p.live_blocks ? list_grid : (uint32_t)((ntri + 255) / 256)), dim3(256), 0, q->stream, p);
        else launch_draw_kernel(setup_kernel<false>, dim3((uint32_t)((ntri + 255) / 256)), dim3(256), 0, q->stream, p);
        if ((r = kzn_launched(q, "setup_kernel"))) return r;
    }
    if (p.binned) {
        {   // (as many CTAs as could have work if every primitive needed clipping; usually none does and the last one scans)
            const uint64_t blocks = (ntri + 1023) / 1024, cap = (uint64_t)q->device->sm_count;
            launch_draw_kernel(bin_scan_kernel, dim3((uint32_t)(blocks < cap ? blocks : cap)), dim3(1024), 0, q->stream, p);
        }
        if ((r = kzn_launched(q, "bin_scan_kernel"))) return r;
        launch_draw_kernel(bin_fill_kernel, dim3(p.live_blocks ? list_grid : (uint32_t)((total + 255) / 256)), dim3(256), 0, q->stream, p);
        if ((r = kzn_launched(q, "bin_fill_kernel"))) return r;
        // few tiles with long lists: a CTA per tile; many tiles with short lists: a warp per tile (see the kernels).  The
        // length is estimated from the viewport (a band of a split frame sees 1 / N of the triangles in 1 / N of the tiles)
        const double viewport_tiles = ceil(fabs((double)vp->width) / kTileW) * ceil(fabs((double)vp->height) / kTileH);
        if ((double)total / (viewport_tiles > 1.0 ? viewport_tiles : 1.0) >= 512.0) launch_draw_kernel(tile_sort_cta_kernel, dim3(p.tiles_x, p.tiles_y), dim3(kSortThreads), sizeof(SortShared), q->stream, p);
        else launch_draw_kernel(tile_sort_warp_kernel, dim3((uint32_t)((tiles + kSortWarps - 1) / kSortWarps)), dim3(kSortThreads), 0, q->stream, p);
        if ((r = kzn_launched(q, "tile_sort_kernel"))) return r;
    }
    dim3 grid(p.tiles_x, p.tiles_y);
    // the destination colour is only read when the result depends on it
    const bool need_dst = has_color && (st->blend_enable || (st->color_write_mask & 15u) != 15u || st->logic_op != 0);
    const size_t smem = sizeof(RasterShared);
    kzn_timing_begin(q);  // the raster kernel dominates a draw
    const bool alpha_over = need_dst && !st->logic_op && st->blend_enable && (st->color_write_mask & 15u) == 15u && st->src_color_blend_factor == 6 &&
                            st->dst_color_blend_factor == 7 && st->color_blend_op == 0 && st->src_alpha_blend_factor == 6 &&
                            st->dst_alpha_blend_factor == 7 && st->alpha_blend_op == 0;
    const int mode = !need_dst ? kDstNone : alpha_over ? kDstAlphaOver : kDstGeneric;
    const bool orbit = st->fragment_program == KZN_FS_PROCEDURAL_ORBIT;
#define KZN_RASTER_LAUNCH(KERNEL, THREADS, SMEM)                                                                      \
    do {                                                                                                     \
        if (orbit) {                                                                                         \
            if (mode == kDstNone) launch_draw_kernel(KERNEL<KZN_FS_PROCEDURAL_ORBIT, kDstNone>, grid, dim3(THREADS), SMEM, q->stream, p);             \
            else if (mode == kDstAlphaOver) launch_draw_kernel(KERNEL<KZN_FS_PROCEDURAL_ORBIT, kDstAlphaOver>, grid, dim3(THREADS), SMEM, q->stream, p); \
            else launch_draw_kernel(KERNEL<KZN_FS_PROCEDURAL_ORBIT, kDstGeneric>, grid, dim3(THREADS), SMEM, q->stream, p);   \
        } else {                                                                                             \
            if (mode == kDstNone) launch_draw_kernel(KERNEL<KZN_FS_VERTEX_COLOR, kDstNone>, grid, dim3(THREADS), SMEM, q->stream, p);                 \
            else if (mode == kDstAlphaOver) launch_draw_kernel(KERNEL<KZN_FS_VERTEX_COLOR, kDstAlphaOver>, grid, dim3(THREADS), SMEM, q->stream, p);     \
            else launch_draw_kernel(KERNEL<KZN_FS_VERTEX_COLOR, kDstGeneric>, grid, dim3(THREADS), SMEM, q->stream, p);       \
        }                                                                                                    \
    } while (0)
    if (!p.binned) {  // a few triangles: register-resident rows, no staging; one warp per 256 x kDirectRows block
        p.tile_x0 = (uint32_t)p.cx0 / kDirectTileW;
        p.tiles_x = (uint32_t)p.cx1 / kDirectTileW - p.tile_x0 + 1;
        p.tile_y0 = (uint32_t)p.cy0 / kDirectRows;
        p.tiles_y = (uint32_t)p.cy1 / kDirectRows - p.tile_y0 + 1;
        const uint64_t blocks = (uint64_t)p.tiles_x * p.tiles_y, resident = (uint64_t)q->device->sm_count * kDirectWarpsPerSM;
        KZN_REQUIRE(blocks <= 0xffffffffull, KZN_ERROR_FEATURE_NOT_PRESENT, "draw: framebuffer too large");
        grid = dim3((uint32_t)(blocks < resident ? blocks : resident));
        if (st->fragment_program == KZN_FS_COMPILED) {
            if ((r = kzn_fragment_program_launch(q->fragment_program, 0, mode, grid.x, 32, 0, q->stream, &p, sizeof p))) return r;
        } else {
            KZN_RASTER_LAUNCH(raster_direct_kernel, 32, 0);
        }
        return kzn_launched(q, "raster_direct_kernel");
    }
    {   // persistent warps over a brick counter: as many CTAs as fit on the device, or fewer when the frame is small
        const uint64_t units = (uint64_t)tiles * kTileBricks, resident = (uint64_t)q->device->sm_count * KZN_RASTER_CTAS;
        KZN_REQUIRE(units <= 0xffffffffull, KZN_ERROR_FEATURE_NOT_PRESENT, "draw: framebuffer too large");
        const uint64_t ctas = (units + kRasterWarps - 1) / kRasterWarps;
        grid = dim3((uint32_t)(ctas < resident ? ctas : resident));
    }
    if (st->fragment_program == KZN_FS_COMPILED) {
        if ((r = kzn_fragment_program_launch(q->fragment_program, 1, mode, grid.x, kRasterThreads, smem, q->stream, &p, sizeof p))) return r;
    } else {
        KZN_RASTER_LAUNCH(raster_kernel, kRasterThreads, smem);
    }
#undef KZN_RASTER_LAUNCH
    return kzn_launched(q, "raster_kernel");
}

}  // namespace

KznResult kzn_raster_vertex_scratch(KznQueue* q, uint64_t bytes, uint64_t* address) {
    if (!q->raster) {
        q->raster = new (std::nothrow) KznRasterScratch();
        KZN_REQUIRE(q->raster, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    }
    KznRasterScratch* s = q->raster;
    KznResult r = grow(q, &s->vs_out, &s->vs_capacity, (size_t)(bytes ? bytes : 16));
    if (r) return r;
    if (q->recording) kzn_command_buffer_uses_raster_scratch(q->recording, q->raster);
    *address = (uint64_t)(uintptr_t)s->vs_out;
    return KZN_SUCCESS;
}

extern "C" {

KznResult kzn_queue_reserve_draw(KznQueue* queue, uint64_t max_triangles, uint32_t width, uint32_t height) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(!queue->recording, "kzn_queue_reserve_draw allocates: call it before vkBeginCommandBuffer");
    KZN_VALID(max_triangles > 0 && max_triangles <= 0x1fffffffull && width && height, "kzn_queue_reserve_draw: bad argument");
    size_t tiles = (size_t)((width + kTileW - 1) / kTileW) * ((height + kTileH - 1) / kTileH);
    return ensure_scratch(queue, (size_t)max_triangles, (size_t)max_triangles, tiles, max_triangles > kDirectMax);
}

KznResult kzn_cmd_draw(KznQueue* queue, const KznGraphicsPipelineState* state, const KznFramebuffer* framebuffer,
                       const KznViewport* viewport, const KznRect2D* scissor, const KznVertexInput* vertices, uint32_t vertex_count,
                       uint32_t instance_count, uint32_t first_vertex, uint32_t first_instance) {
    (void)first_instance;
    return draw_common(queue, state, framebuffer, viewport, scissor, vertices, 0, 0, first_vertex, vertex_count, 0, instance_count);
}

KznResult kzn_cmd_draw_indexed(KznQueue* queue, const KznGraphicsPipelineState* state, const KznFramebuffer* framebuffer,
                               const KznViewport* viewport, const KznRect2D* scissor, const KznVertexInput* vertices,
                               uint64_t index_buffer_address, uint64_t index_buffer_size, uint32_t index_type, uint32_t index_count,
                               uint32_t instance_count, uint32_t first_index, int32_t vertex_offset, uint32_t first_instance) {
    (void)first_instance;
    KZN_VALID(index_buffer_address && index_type <= KZN_INDEX_TYPE_UINT32, "vkCmdDrawIndexed: bad index buffer");
    uint64_t isz = index_type == KZN_INDEX_TYPE_UINT16 ? 2 : 4;
    KZN_VALID(index_buffer_address % isz == 0, "vkCmdDrawIndexed: index buffer misaligned");
    KZN_VALID(((uint64_t)first_index + index_count) * isz <= index_buffer_size, "vkCmdDrawIndexed: indices outside the index buffer");
    return draw_common(queue, state, framebuffer, viewport, scissor, vertices, index_buffer_address, index_type, first_index, index_count,
                       vertex_offset, instance_count);
}

}  // extern "C"

#endif  // !KZN_FS_TEMPLATE
