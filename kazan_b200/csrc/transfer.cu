// transfer.cu — the transfer family as HBM-bound sm_100a kernels.
//
// Replaces the `unimplemented!()` bodies of (vulkan-driver/src/api_impl.rs):
//   vkCmdCopyBuffer :4696  vkCmdCopyImage :4707  vkCmdBlitImage :4720  vkCmdCopyBufferToImage :4734
//   vkCmdCopyImageToBuffer :4746  vkCmdUpdateBuffer :4758  vkCmdFillBuffer :4769
//   vkCmdClearColorImage :4780  vkCmdClearDepthStencilImage :4792  vkCmdClearAttachments :4804
// Image addressing follows vulkan-driver/src/image.rs:58-88 (LINEAR: tightly packed) and the
// 16x16-texel tile layout of SURVEY.md E.10 (TILED).
//
// Kernels
//   fill_contig_kernel   contiguous word fill: 128-bit st.global, grid = k x SM count, grid-stride
//   fill_bulk_kernel     same job through the TMA engine: the CTA paints 32 KiB of shared memory once,
//                        then one thread streams it out with cp.async.bulk.global.shared::cta
//   copy_bulk_kernel     contiguous copy through the TMA engine: cp.async.bulk global->shared (mbarrier
//                        complete_tx) then shared->global (bulk_group), 4-stage ring per CTA, no registers
//   box_kernel<...>      general box copy/fill between LINEAR (pitched) and TILED surfaces,
//                        16-byte vectors when alignment allows, 4-byte texels otherwise
//   blit_kernel          scaled RGBA8 copy, NEAREST (integer path) or LINEAR (fp32, no contraction)
//   copy_realign_kernel  contiguous copy whose source is only 4-byte aligned: aligned 16-byte loads, the neighbour's
//                        vector by shuffle, realigned 16-byte stores
//   merged_transfer_kernel  up to 96 recorded contiguous fills / copies with disjoint ranges as one launch
// Except for the two bulk (TMA) kernels, launches use programmatic dependent launch (kzn_internal.h).
// All of them are bandwidth bound; algorithmic bytes per texel are 4 (fill) or 8 (copy).
#include <cuda.h>  // CUtensorMap and its enumerations (the encode function is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>
#include <string.h>

#include "kzn_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// ---- device helpers -----------------------------------------------------------------------------

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t ld_stream4(const void* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// TMA (bulk async-proxy) copies; SASS: UBLKCP
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- contiguous fill -------------------------------------------------------------------------------

// dst is 4-byte aligned, `words` 32-bit words.  Head/tail words are peeled so the body is 16-byte stores.
__global__ void __launch_bounds__(kThreads) fill_contig_kernel(uint32_t* __restrict__ dst, uint64_t words, uint32_t value) {
    kzn_pdl_wait();
    uint64_t head = ((16 - ((uintptr_t)dst & 15)) & 15) >> 2;
    if (head > words) head = words;
    uint64_t body = (words - head) >> 2;  // uint4 count
    uint64_t tail = words - head - (body << 2);
    uint64_t tid = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    uint64_t nthreads = (uint64_t)gridDim.x * kThreads;
    if (tid < head) dst[tid] = value;
    if (tid < tail) dst[head + (body << 2) + tid] = value;
    uint4* b = reinterpret_cast<uint4*>(dst + head);
    const uint4 v = make_uint4(value, value, value, value);
    uint64_t i = tid;
    for (; i + (kUnroll - 1) * nthreads < body; i += kUnroll * nthreads) {
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) st_stream16(b + i + k * nthreads, v);
    }
    for (; i < body; i += nthreads) st_stream16(b + i, v);
}

constexpr uint32_t kBulkChunk = 32768;  // bytes per TMA bulk operation

// dst 16-byte aligned, bytes a multiple of 16.  One CTA per SM slot; CTA c handles chunks c, c+grid, ...
__global__ void __launch_bounds__(kThreads) fill_bulk_kernel(uint8_t* __restrict__ dst, uint64_t bytes, uint32_t value) {
    kzn_pdl_wait();
    extern __shared__ __align__(128) uint8_t smem[];
    uint4* s4 = reinterpret_cast<uint4*>(smem);
    const uint4 v = make_uint4(value, value, value, value);
    for (uint32_t i = threadIdx.x; i < kBulkChunk / 16; i += kThreads) s4[i] = v;
    fence_proxy_async();  // generic-proxy writes -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t chunks = (bytes + kBulkChunk - 1) / kBulkChunk;
        for (uint64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
            uint64_t off = c * kBulkChunk;
            uint32_t n = (uint32_t)(bytes - off < kBulkChunk ? bytes - off : kBulkChunk);
            bulk_s2g(dst + off, smem, n);
        }
        bulk_commit();
        bulk_wait_read<0>();  // shared memory must stay live until every bulk read of it is done
    }
}

// ---- contiguous copy through TMA ---------------------------------------------------------------------

constexpr int kCopyStages = 4;
constexpr uint32_t kCopyChunk = 32768;

// src/dst 16-byte aligned, bytes a multiple of 16.  Thread 0 of each CTA drives a 4-stage ring:
//   wait(stage free: its last store has finished reading smem) -> bulk load (mbarrier tx) ...
//   wait(load landed) -> bulk store -> commit
__global__ void __launch_bounds__(32) copy_bulk_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint64_t bytes) {
    kzn_pdl_wait();
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kCopyStages];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kCopyStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const uint64_t chunks = (bytes + kCopyChunk - 1) / kCopyChunk;
    // chunks owned by this CTA: blockIdx.x, +grid, ...
    uint64_t n_mine = chunks > blockIdx.x ? (chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto chunk_off = [&](uint64_t k) { return (blockIdx.x + k * (uint64_t)gridDim.x) * kCopyChunk; };
    auto chunk_len = [&](uint64_t off) { return (uint32_t)(bytes - off < kCopyChunk ? bytes - off : kCopyChunk); };
    // prologue: fill the ring
    for (uint64_t j = 0; j < n_mine && j < kCopyStages; ++j) {
        uint64_t off = chunk_off(j);
        uint32_t n = chunk_len(off);
        mbar_expect_tx(&full[j], n);
        bulk_g2s(smem + j * kCopyChunk, src + off, n, &full[j]);
    }
    for (uint64_t k = 0; k < n_mine; ++k) {
        int s = (int)(k % kCopyStages);
        mbar_wait(&full[s], (uint32_t)((k / kCopyStages) & 1));
        uint64_t off = chunk_off(k);
        bulk_s2g(dst + off, smem + s * kCopyChunk, chunk_len(off));
        bulk_commit();
        // refill the stage drained one iteration ago: all but the newest store group have finished
        // reading shared memory after wait_group.read 1
        uint64_t j = k + kCopyStages - 1;
        if (k >= 1 && j < n_mine) {
            bulk_wait_read<1>();
            int sj = (int)(j % kCopyStages);
            uint64_t off2 = chunk_off(j);
            uint32_t n2 = chunk_len(off2);
            mbar_expect_tx(&full[sj], n2);
            bulk_g2s(smem + sj * kCopyChunk, src + off2, n2, &full[sj]);
        }
    }
    bulk_wait_read<0>();
}

// ---- tiled <-> linear layout swizzle through TMA tensor copies ------------------------------------------------------------

// The TILED layout is 16 x 16-texel tiles of 1 KiB, row-major inside and across (oracle/transfer.c).  The LINEAR image is
// described to the TMA unit as a 4-D tensor (texel in tile row: 16 x 4 B | row in tile: 16, stride = row pitch | tile
// column: stride 64 B | tile row: stride 16 row pitches), so a box of {16, 16, kSwzTiles, 1} lands in shared memory as
// kSwzTiles whole tiles in exactly the TILED byte order: the other side of the copy is ONE contiguous bulk copy of
// kSwzTiles KiB.  No thread computes an address or touches a texel (SASS: UTMALDG / UTMASTG for the tensor side, UBLKCP for the
// contiguous side); thread 0 of each CTA drives a ring of stages as in copy_bulk_kernel.
// Measured on B200 (profiles/r02l, bench.py extra.transfer): 4K 0.71 / 0.81 of the HBM peak (to TILED / to LINEAR), 8K 0.82 /
// 0.86 — box_kernel<4> with per-thread addresses does 0.77 / 0.81 and 0.93 / 0.96 (the tensor box has 64-byte rows, 128 row
// segments per 8 KiB operation).  So the default path is box_kernel; this one runs with KZN_TRANSFER_PATH=tma.
constexpr int kSwzTiles = 8;                       // tiles per stage: 8 KiB
constexpr int kSwzStages = 6;
constexpr uint32_t kSwzBytes = kSwzTiles * 1024u;

__device__ __forceinline__ void tensor_g2s_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tensor_s2g_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3), "r"(smem_u32(smem_src))
                 : "memory");
}

// groups of kSwzTiles consecutive tiles of a tile row; `groups_x` per tile row (the last one of a row may be partial: the
// tensor side clips it, the contiguous side is cut to the tiles that exist)
template <bool TO_TILED>
__global__ void __launch_bounds__(32) swizzle_tma_kernel(const __grid_constant__ CUtensorMap linear_map, uint8_t* __restrict__ tiled,
                                                         uint32_t tiles_x, uint32_t groups_x, uint64_t groups) {
    kzn_pdl_wait();
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kSwzStages];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSwzStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const uint64_t n_mine = groups > blockIdx.x ? (groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto load = [&](uint64_t k) {  // group k of this CTA -> its stage
        const uint64_t g = blockIdx.x + k * (uint64_t)gridDim.x;
        const uint32_t ty = (uint32_t)(g / groups_x), tx0 = (uint32_t)(g % groups_x) * kSwzTiles;
        const uint32_t nt = min((uint32_t)kSwzTiles, tiles_x - tx0);
        const int st = (int)(k % kSwzStages);
        // the tensor box is always kSwzTiles wide: tiles past the image edge are zero-filled in shared memory (and counted)
        mbar_expect_tx(&full[st], TO_TILED ? kSwzBytes : nt * 1024u);
        if (TO_TILED) tensor_g2s_4d(smem + st * kSwzBytes, &linear_map, 0, 0, (int)tx0, (int)ty, &full[st]);
        else bulk_g2s(smem + st * kSwzBytes, tiled + ((uint64_t)ty * tiles_x + tx0) * 1024u, nt * 1024u, &full[st]);
    };
    for (uint64_t j = 0; j < n_mine && j < kSwzStages; ++j) load(j);
    for (uint64_t k = 0; k < n_mine; ++k) {
        const int st = (int)(k % kSwzStages);
        mbar_wait(&full[st], (uint32_t)((k / kSwzStages) & 1));
        const uint64_t g = blockIdx.x + k * (uint64_t)gridDim.x;
        const uint32_t ty = (uint32_t)(g / groups_x), tx0 = (uint32_t)(g % groups_x) * kSwzTiles;
        const uint32_t nt = min((uint32_t)kSwzTiles, tiles_x - tx0);
        if (TO_TILED) bulk_s2g(tiled + ((uint64_t)ty * tiles_x + tx0) * 1024u, smem + st * kSwzBytes, nt * 1024u);
        else tensor_s2g_4d(&linear_map, 0, 0, (int)tx0, (int)ty, smem + st * kSwzBytes);  // clipped to the tensor's extent
        bulk_commit();
        const uint64_t j = k + kSwzStages - 1;
        if (k >= 1 && j < n_mine) {
            bulk_wait_read<1>();
            load(j);
        }
    }
    bulk_wait_read<0>();
}

// ---- contiguous copy from a source that is only 4-byte aligned --------------------------------------------------

// vkCmdCopyBufferToImage allows any bufferOffset that is a multiple of the texel size (4): the destination image row is
// 16-byte aligned but the source is not.  Instead of falling back to 4-byte accesses, every lane loads the ALIGNED
// 16-byte vector under its destination vector, takes the next one from its neighbour with four shuffles (lane 31 and
// the end of the run load it themselves; the last, partial vector is read word by word so that nothing beyond the source
// range is touched) and stores one realigned 16-byte vector.  shift = (src & 15) / 4 in {1, 2, 3}.
__global__ void __launch_bounds__(kThreads) copy_realign_kernel(uint4* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t vecs) {
    kzn_pdl_wait();
    const uint32_t shift = (uint32_t)(((uintptr_t)src & 15u) >> 2);
    const uint4* a_base = reinterpret_cast<const uint4*>((uintptr_t)src & ~(uintptr_t)15);
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t nthreads = (uint64_t)gridDim.x * kThreads;
    const uint64_t first = (uint64_t)blockIdx.x * kThreads + (threadIdx.x & ~31u);  // my warp's first vector
    for (uint64_t w0 = first; w0 < vecs; w0 += nthreads) {  // uniform per warp
        const uint64_t i = w0 + lane;
        const bool ok = i < vecs;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        if (ok) a = ld_stream16(a_base + i);
        b.x = __shfl_down_sync(0xffffffffu, a.x, 1);
        b.y = __shfl_down_sync(0xffffffffu, a.y, 1);
        b.z = __shfl_down_sync(0xffffffffu, a.z, 1);
        if (ok && (lane == 31u || i + 1 >= vecs)) {
            if (i + 1 < vecs) {
                const uint4 n = ld_stream16(a_base + i + 1);
                b.x = n.x; b.y = n.y; b.z = n.z;
            } else {  // the vector after the last one holds only `shift` valid words
                const uint32_t* t = reinterpret_cast<const uint32_t*>(a_base + i + 1);
                b.x = ld_stream4(t);
                if (shift > 1) b.y = ld_stream4(t + 1);
                if (shift > 2) b.z = ld_stream4(t + 2);
            }
        }
        if (ok) {
            uint4 v;
            if (shift == 1) v = make_uint4(a.y, a.z, a.w, b.x);
            else if (shift == 2) v = make_uint4(a.z, a.w, b.x, b.y);
            else v = make_uint4(a.w, b.x, b.y, b.z);
            st_stream16(dst + i, v);
        }
    }
}

// ---- merged fills / copies (command buffers) ---------------------------------------------------------------

// Up to KZN_MAX_MERGED recorded commands as one launch.  Work is cut into 16 KiB units numbered across the commands
// (first_unit = prefix sum); CTA b takes units b, b + grid, ...; a unit is 4 x 16 bytes per thread, loads issued before
// the stores.  A command is a contiguous run (wv == 0) or a pitched box of LINEAR memory (wv vectors per row at the two
// row pitches: padded bufferRowLength, sub-rectangles); destinations are 16-byte aligned and lengths multiples of 16
// (checked at record time).  Copies whose source is only 4-byte aligned stay with copy_realign_kernel: realigning inside
// this kernel measured 0.71 of HBM at 4K against 0.79 for the kernel of its own (one vector per thread, 28 registers).
constexpr uint32_t kUnitVecs = 1024;  // uint4 per unit
struct MergedOps {
    uint32_t n, total_units;
    uint32_t first_unit[KZN_MAX_MERGED + 1];
    uint32_t vecs[KZN_MAX_MERGED];   // length in uint4
    uint32_t value[KZN_MAX_MERGED];  // fills
    uint32_t wv[KZN_MAX_MERGED];     // pitched boxes: uint4 per row (0: contiguous)
    uint32_t dst_pitch[KZN_MAX_MERGED], src_pitch[KZN_MAX_MERGED];  // bytes
    unsigned long long dst[KZN_MAX_MERGED];
    unsigned long long src[KZN_MAX_MERGED];  // 0: fill
};
static_assert(sizeof(MergedOps) <= 4096, "kernel parameter space");

// BOXES: some command is a pitched box (the plain form keeps its 32 registers and full occupancy)
template <bool BOXES>
__global__ void __launch_bounds__(kThreads) merged_transfer_kernel(const __grid_constant__ MergedOps m) {
    kzn_pdl_wait();
    constexpr uint32_t K = kUnitVecs / kThreads;
    for (uint32_t u = blockIdx.x; u < m.total_units; u += gridDim.x) {
        uint32_t lo = 0, hi = m.n;  // the command this unit belongs to: last i with first_unit[i] <= u
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (m.first_unit[mid] <= u) lo = mid;
            else hi = mid;
        }
        const uint32_t v0 = (u - m.first_unit[lo]) * kUnitVecs;
        const uint32_t vecs = m.vecs[lo];
        const uint32_t cnt = min(kUnitVecs, vecs - v0);
        const unsigned long long src = m.src[lo], dst = m.dst[lo];
        const uint32_t wv = m.wv[lo];
        if (BOXES && wv) {  // pitched box
            const uint32_t dp = m.dst_pitch[lo], sp = m.src_pitch[lo];
            unsigned long long da[K];
            uint4 v[K];
#pragma unroll
            for (uint32_t k = 0; k < K; ++k) {
                const uint32_t i = v0 + threadIdx.x + k * kThreads;
                const uint32_t row = i / wv, x = i - row * wv;
                da[k] = dst + (unsigned long long)row * dp + 16ull * x;
                if (src == 0) v[k] = make_uint4(m.value[lo], m.value[lo], m.value[lo], m.value[lo]);
                else if (i < vecs) v[k] = ld_stream16((const void*)(src + (unsigned long long)row * sp + 16ull * x));
            }
#pragma unroll
            for (uint32_t k = 0; k < K; ++k)
                if (v0 + threadIdx.x + k * kThreads < vecs) st_stream16((void*)da[k], v[k]);
            continue;
        }
        uint4* d = reinterpret_cast<uint4*>(dst) + v0;
        if (src == 0) {
            const uint32_t w = m.value[lo];
            const uint4 v = make_uint4(w, w, w, w);
#pragma unroll
            for (uint32_t k = 0; k < K; ++k) {
                const uint32_t i = threadIdx.x + k * kThreads;
                if (i < cnt) st_stream16(d + i, v);
            }
        } else {
            const uint4* s = reinterpret_cast<const uint4*>(src) + v0;
            uint4 v[K];
#pragma unroll
            for (uint32_t k = 0; k < K; ++k) {
                const uint32_t i = threadIdx.x + k * kThreads;
                if (i < cnt) v[k] = ld_stream16(s + i);
            }
#pragma unroll
            for (uint32_t k = 0; k < K; ++k) {
                const uint32_t i = threadIdx.x + k * kThreads;
                if (i < cnt) st_stream16(d + i, v[k]);
            }
        }
    }
}

// ---- general box copy / fill ----------------------------------------------------------------------------

struct Surf {
    unsigned long long base;         // byte address of (x=0,y=0) of the box's first slice
    unsigned long long row_pitch;    // bytes, LINEAR
    unsigned long long slice_pitch;  // bytes
    uint32_t x0, y0;                 // TILED: box origin in texels (LINEAR folds it into base)
    uint32_t tiles_x;                // TILED
};

template <bool TILED>
__device__ __forceinline__ unsigned long long surf_addr(const Surf& s, uint32_t x, uint32_t y, uint32_t slice) {
    if (TILED) {
        uint32_t X = s.x0 + x, Y = s.y0 + y;
        unsigned long long tile = (unsigned long long)(Y >> 4) * s.tiles_x + (X >> 4);
        return s.base + slice * s.slice_pitch + ((tile << 8) + ((Y & 15u) << 4) + (X & 15u)) * 4ull;
    }
    return s.base + slice * s.slice_pitch + y * s.row_pitch + x * 4ull;
}

struct BoxParams {
    Surf src, dst;
    uint32_t wv;      // items per row (texels / VEC)
    uint32_t h;       // rows per slice
    uint32_t slices;
    uint32_t fill_value;
    unsigned long long total;  // wv*h*slices
};

template <int VEC, bool SRC_TILED, bool DST_TILED, bool FILL>
__global__ void __launch_bounds__(kThreads) box_kernel(const __grid_constant__ BoxParams p) {
    kzn_pdl_wait();
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kThreads;
    unsigned long long i0 = (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
    const bool narrow = p.total <= 0xffffffffull;
    for (; i0 < p.total; i0 += kUnroll * nthreads) {
        unsigned long long sa[kUnroll], da[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            unsigned long long i = i0 + k * nthreads;
            ok[k] = i < p.total;
            uint32_t xv, y, s;
            if (!SRC_TILED && !DST_TILED && p.h == 1 && p.slices == 1) {
                // one contiguous run of LINEAR memory (a single row, or rows collapsed by the host): wv may exceed 32 bits of
                // bytes, keep 64-bit x.  A TILED side always goes through surf_addr, also for a single row.
                sa[k] = p.src.base + i * (4ull * VEC);
                da[k] = p.dst.base + i * (4ull * VEC);
                continue;
            }
            if (narrow) {
                uint32_t ii = (uint32_t)i;
                uint32_t row = ii / p.wv;
                xv = ii - row * p.wv;
                s = row / p.h;
                y = row - s * p.h;
            } else {
                unsigned long long row = i / p.wv;
                xv = (uint32_t)(i - row * p.wv);
                s = (uint32_t)(row / p.h);
                y = (uint32_t)(row - (unsigned long long)s * p.h);
            }
            if (!FILL) sa[k] = surf_addr<SRC_TILED>(p.src, xv * VEC, y, s);
            da[k] = surf_addr<DST_TILED>(p.dst, xv * VEC, y, s);
        }
        if (VEC == 4) {
            uint4 v[kUnroll];
#pragma unroll
            for (int k = 0; k < kUnroll; ++k) {
                if (FILL) v[k] = make_uint4(p.fill_value, p.fill_value, p.fill_value, p.fill_value);
                else if (ok[k]) v[k] = ld_stream16((const void*)sa[k]);
            }
#pragma unroll
            for (int k = 0; k < kUnroll; ++k)
                if (ok[k]) st_stream16((void*)da[k], v[k]);
        } else {
            uint32_t v[kUnroll];
#pragma unroll
            for (int k = 0; k < kUnroll; ++k) {
                if (FILL) v[k] = p.fill_value;
                else if (ok[k]) v[k] = ld_stream4((const void*)sa[k]);
            }
#pragma unroll
            for (int k = 0; k < kUnroll; ++k)
                if (ok[k]) st_stream4((void*)da[k], v[k]);
        }
    }
}

// ---- blit ------------------------------------------------------------------------------

struct BlitParams {
    Surf src, dst;  // whole image surfaces: base = image address (+ layer), x0=y0=0
    uint32_t src_tiled, dst_tiled;
    int32_t src_w, src_h;
    int32_t dx0, dx1, dy0, dy1;  // destination box corners as given (may be reversed)
    int32_t sx0, sx1, sy0, sy1;  // source box corners as given
    int32_t x_lo, y_lo;          // min corner of the destination box
    uint32_t w, h;               // destination box size
    uint32_t filter;
    float scale_x, scale_y;      // (s1 - s0) / (d1 - d0) per axis: one IEEE division each, done on the host
};

__device__ __forceinline__ float blit_coord(int32_t i, int32_t d0, int32_t d1, int32_t s0, int32_t s1) {
    float scale = __fdiv_rn((float)(s1 - s0), (float)(d1 - d0));
    return __fmaf_rn(__fsub_rn(__fadd_rn((float)i, 0.5f), (float)d0), scale, (float)s0);
}
__device__ __forceinline__ int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ unsigned long long surf_addr_rt(const Surf& s, uint32_t tiled, uint32_t x, uint32_t y) {
    return tiled ? surf_addr<true>(s, x, y, 0) : surf_addr<false>(s, x, y, 0);
}
__device__ __forceinline__ float unorm8_to_float(uint32_t u) { return __fdiv_rn((float)u, 255.0f); }
__device__ __forceinline__ uint32_t float_to_unorm8(float f) {
    // E.1: NaN -> 0, clamp, single multiply, round to nearest even
    return __float2uint_rn(__fmul_rn(__saturatef(f), 255.0f));
}

// One thread makes 4 horizontally adjacent destination texels (one 16-byte store where the row allows it); a warp covers
// 128 texels of a row, the CTA's 8 warps 8 rows.  The per-axis scale is computed once on the host (the same IEEE division
// blit_coord would repeat per texel), UNORM8 -> float is the division-free exact form unorm8_exact (16 IEEE divisions per
// LINEAR texel before), source texels are read through the read-only path (neighbouring
// threads share them in L1).  Same operations in the same order as oracle/transfer.c: results are unchanged.
__device__ __forceinline__ float blit_coord_scaled(int32_t i, int32_t d0, float scale, int32_t s0) {
    return __fmaf_rn(__fsub_rn(__fadd_rn((float)i, 0.5f), (float)d0), scale, (float)s0);
}
__device__ __forceinline__ uint32_t ld_texel(unsigned long long a) { return __ldg(reinterpret_cast<const uint32_t*>(a)); }

// u / 255.0f, correctly rounded, without a division: q0 = u * fl(1/255), one fused correction step (the remainder
// u - 255 q0 is exact in an fma).  Equal to __fdiv_rn(u, 255.0f) for all 256 values of u (tests/test_oracle_transfer.py
// checks the sequence in exact arithmetic): the 256-entry table this replaces cost 16 shared-memory lookups with bank
// conflicts per LINEAR texel, which was the kernel's bound.
// byte `ch` of a texel as a float: the byte is placed under the exponent of 2^23 with one byte permute (0x4B0000uu = 2^23 + u)
// and 2^23 subtracted — exact, and on the ALU / fp32 pipes where the compiler's I2F.U8 goes through the conversion unit (16
// lanes per clock: 16 conversions per LINEAR texel were the kernel's bound after the table was gone)
template <int CH>
__device__ __forceinline__ float texel_byte(uint32_t t) {
    return __fsub_rn(__uint_as_float(__byte_perm(t, 0x4B000000u, 0x7540u | CH)), 8388608.0f);
}
// packed fp32 pairs (FMUL2 / FADD2): two IEEE operations per instruction.  NOTE: ptxas 12.9 contracts add.rn.f32x2 of a
// mul.rn.f32x2 result into FFMA2 even though both carry .rn (it leaves mul.rn.f32 / add.rn.f32 alone): keep packed products
// away from packed adds.
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
    return *reinterpret_cast<float2*>(&r);
}
// ... bytes CH and CH + 1 as a pair
template <int CH>
__device__ __forceinline__ float2 texel_bytes(uint32_t t) {
    const float2 b = make_float2(__uint_as_float(__byte_perm(t, 0x4B000000u, 0x7540u | CH)), __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7540u | (CH + 1))));
    return add2(b, make_float2(-8388608.0f, -8388608.0f));
}
__device__ __forceinline__ float2 unorm8_exact2(float2 x) {
    const float2 r = make_float2(0.003921568859368563f, 0.003921568859368563f), k = make_float2(255.0f, 255.0f);
    const float2 q0 = mul2(x, r);
    return __ffma2_rn(__ffma2_rn(make_float2(-q0.x, -q0.y), k, x), r, q0);
}
__device__ __forceinline__ float unorm8_exact(float x) {
    const float r = 0.003921568859368563f;  // fl(1 / 255)
    const float q0 = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-q0, 255.0f, x), r, q0);
}

template <bool SRC_TILED, bool DST_TILED, bool LINEAR_FILTER>
__global__ void __launch_bounds__(kThreads) blit_kernel(const __grid_constant__ BlitParams p) {
    kzn_pdl_wait();
    const uint32_t i0 = (blockIdx.x * 32u + (threadIdx.x & 31u)) * 4u;
    const uint32_t j = blockIdx.y * (kThreads / 32) + (threadIdx.x >> 5);
    if (i0 >= p.w || j >= p.h) return;
    const int32_t dy = p.y_lo + (int32_t)j;
    const float v = blit_coord_scaled(dy, p.dy0, p.scale_y, p.sy0);
    uint32_t out[4];
    if (!LINEAR_FILTER) {
        const int32_t sy = clampi((int32_t)floorf(v), 0, p.src_h - 1);
        const unsigned long long row = SRC_TILED ? 0ull : surf_addr<false>(p.src, 0, sy, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float u = blit_coord_scaled(p.x_lo + (int32_t)(i0 + k), p.dx0, p.scale_x, p.sx0);
            const int32_t sx = clampi((int32_t)floorf(u), 0, p.src_w - 1);
            out[k] = ld_texel(SRC_TILED ? surf_addr<true>(p.src, sx, sy, 0) : row + 4ull * (uint32_t)sx);
        }
    } else {
        const float vv = __fsub_rn(v, 0.5f), fv = floorf(vv), b = __fsub_rn(vv, fv), nb = __fsub_rn(1.0f, b);
        const int32_t y0 = clampi((int32_t)fv, 0, p.src_h - 1), y1 = clampi((int32_t)fv + 1, 0, p.src_h - 1);
        const unsigned long long row0 = SRC_TILED ? 0ull : surf_addr<false>(p.src, 0, y0, 0), row1 = SRC_TILED ? 0ull : surf_addr<false>(p.src, 0, y1, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float u = blit_coord_scaled(p.x_lo + (int32_t)(i0 + k), p.dx0, p.scale_x, p.sx0);
            const float uu = __fsub_rn(u, 0.5f), fu = floorf(uu), a = __fsub_rn(uu, fu), na = __fsub_rn(1.0f, a);
            const int32_t x0 = clampi((int32_t)fu, 0, p.src_w - 1), x1 = clampi((int32_t)fu + 1, 0, p.src_w - 1);
            const uint32_t t00 = ld_texel(SRC_TILED ? surf_addr<true>(p.src, x0, y0, 0) : row0 + 4ull * (uint32_t)x0);
            const uint32_t t10 = ld_texel(SRC_TILED ? surf_addr<true>(p.src, x1, y0, 0) : row0 + 4ull * (uint32_t)x1);
            const uint32_t t01 = ld_texel(SRC_TILED ? surf_addr<true>(p.src, x0, y1, 0) : row1 + 4ull * (uint32_t)x0);
            const uint32_t t11 = ld_texel(SRC_TILED ? surf_addr<true>(p.src, x1, y1, 0) : row1 + 4ull * (uint32_t)x1);
            const float w00 = __fmul_rn(na, nb), w10 = __fmul_rn(a, nb), w01 = __fmul_rn(na, b), w11 = __fmul_rn(a, b);
            // two channels per instruction: packed fmul / fadd / fma (.f32x2: FMUL2, FADD2, FFMA2) do two IEEE operations
            // with the results of the scalar instructions — the same sequence per channel as oracle/transfer.c
            uint32_t bits[4];
            auto channels = [&](float2 c00, float2 c10, float2 c01, float2 c11, uint32_t& lo, uint32_t& hi) {
                // products packed, sums scalar: ptxas contracts an add.rn.f32x2 of a mul.rn.f32x2 into one FFMA2 (one rounding
                // where the oracle has two) although it leaves the scalar forms alone, so a packed product never meets a
                // packed add here
                const float2 p00 = mul2(make_float2(w00, w00), unorm8_exact2(c00)), p10 = mul2(make_float2(w10, w10), unorm8_exact2(c10));
                const float2 p01 = mul2(make_float2(w01, w01), unorm8_exact2(c01)), p11 = mul2(make_float2(w11, w11), unorm8_exact2(c11));
                const float ax = __fadd_rn(__fadd_rn(__fadd_rn(p00.x, p10.x), p01.x), p11.x);
                const float ay = __fadd_rn(__fadd_rn(__fadd_rn(p00.y, p10.y), p01.y), p11.y);
                // E.1 with the rounding done by an fp32 add of 1.5 * 2^23 (exactly cvt.rni for [0, 255]): the byte is the low
                // byte of the sum
                const float2 m = mul2(make_float2(__saturatef(ax), __saturatef(ay)), make_float2(255.0f, 255.0f));
                lo = __float_as_uint(__fadd_rn(m.x, 12582912.0f));
                hi = __float_as_uint(__fadd_rn(m.y, 12582912.0f));
            };
            channels(texel_bytes<0>(t00), texel_bytes<0>(t10), texel_bytes<0>(t01), texel_bytes<0>(t11), bits[0], bits[1]);
            channels(texel_bytes<2>(t00), texel_bytes<2>(t10), texel_bytes<2>(t01), texel_bytes<2>(t11), bits[2], bits[3]);
            out[k] = __byte_perm(__byte_perm(bits[0], bits[1], 0x4440), __byte_perm(bits[2], bits[3], 0x4440), 0x5410);
        }
    }
    const int32_t dx = p.x_lo + (int32_t)i0;
    const unsigned long long da = surf_addr<DST_TILED>(p.dst, dx, dy, 0);
    // the 4 texels are contiguous in memory when the row is LINEAR, or TILED and they do not cross a tile (dx % 4 == 0)
    const bool contiguous = !DST_TILED || (dx & 3) == 0;
    if (i0 + 4u <= p.w && contiguous && (da & 15ull) == 0) {
        st_stream16((void*)da, make_uint4(out[0], out[1], out[2], out[3]));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < p.w) *(uint32_t*)surf_addr<DST_TILED>(p.dst, dx + k, dy, 0) = out[k];
    }
}

typedef void (*BlitKernel)(BlitParams);
BlitKernel blit_kernel_for(bool src_tiled, bool dst_tiled, bool linear) {
    static const BlitKernel table[8] = {blit_kernel<false, false, false>, blit_kernel<false, false, true>, blit_kernel<false, true, false>,
                                        blit_kernel<false, true, true>,   blit_kernel<true, false, false>, blit_kernel<true, false, true>,
                                        blit_kernel<true, true, false>,   blit_kernel<true, true, true>};
    return table[(src_tiled ? 4 : 0) | (dst_tiled ? 2 : 0) | (linear ? 1 : 0)];
}

// ---- host-side launch logic ------------------------------------------------------------------------------

enum class Path { Auto, LdSt, Tma };
Path transfer_path() {
    static Path p = [] {
        const char* e = getenv("KZN_TRANSFER_PATH");
        if (e && !strcmp(e, "ldst")) return Path::LdSt;
        if (e && !strcmp(e, "tma")) return Path::Tma;
        return Path::Auto;
    }();
    return p;
}

int grid_for(const KznQueue* q, unsigned long long items) {
    unsigned long long blocks = (items + (unsigned long long)kThreads * kUnroll - 1) / ((unsigned long long)kThreads * kUnroll);
    // A multiple of the SM count.  The cap is what balances a medium-sized command: with 8 CTAs per SM a 4K image (2025
    // blocks of work) ran as 1.7 grid-stride sweeps — every CTA waits for the ones that do two — and lost 15 %; with the
    // blocks handed out by the hardware one by one the last wave is a fraction of a sweep.  (KZN_GRID_PER_SM overrides.)
    static const unsigned long long per_sm = [] {
        const char* e = getenv("KZN_GRID_PER_SM");
        const long v = e ? atol(e) : 0;
        return (unsigned long long)(v > 0 ? v : 64);
    }();
    unsigned long long cap = (unsigned long long)q->device->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    return blocks ? (int)blocks : 1;
}

// ---- merging of recorded commands ---------------------------------------------------------------------------

bool merge_enabled() {
    static bool on = [] {
        const char* e = getenv("KZN_MERGE_TRANSFERS");
        return !(e && !strcmp(e, "0"));
    }();
    return on;
}

bool ranges_overlap(uint64_t a, uint64_t an, uint64_t b, uint64_t bn) { return a < b + bn && b < a + an; }

KznResult launch_fill_contig(KznQueue* q, uint64_t address, uint64_t bytes, uint32_t value, bool may_merge = true);

// Queue a fill (src == 0) or copy for the next merged launch: one contiguous run of `bytes` (row_bytes == 0), or `rows`
// rows of row_bytes at the two pitches.  false: the caller launches it itself.
bool merge_transfer(KznQueue* q, uint64_t dst, uint64_t src, uint64_t bytes, uint32_t value, KznResult* result, uint64_t row_bytes = 0,
                    uint64_t rows = 0, uint64_t dst_pitch = 0, uint64_t src_pitch = 0) {
    *result = KZN_SUCCESS;
    if (!q->recording || !merge_enabled()) return false;
    if (dst % 16 || bytes % 16 || bytes == 0 || bytes > (64ull << 20)) return false;
    if (row_bytes) {
        if (src % 16 || row_bytes % 16 || dst_pitch % 16 || src_pitch % 16 || dst_pitch > 0xffffffffull || src_pitch > 0xffffffffull) return false;
    } else if (src % 16) {
        return false;
    }
    // the byte ranges the command touches: for a pitched box everything between its first and its last byte (conservative)
    const uint64_t dspan = row_bytes ? (rows - 1) * dst_pitch + row_bytes : bytes, sspan = row_bytes ? (rows - 1) * src_pitch + row_bytes : bytes;
    bool conflict = q->n_pending == KZN_MAX_MERGED;
    for (uint32_t i = 0; i < q->n_pending && !conflict; ++i) {
        const KznPendingTransfer& o = q->pending[i];
        const uint64_t odspan = o.row_bytes ? (uint64_t)(o.rows - 1) * o.dst_pitch + o.row_bytes : o.bytes;
        const uint64_t osspan = o.row_bytes ? (uint64_t)(o.rows - 1) * o.src_pitch + o.row_bytes : o.bytes;
        conflict = ranges_overlap(dst, dspan, o.dst, odspan) || (o.src && ranges_overlap(dst, dspan, o.src, osspan)) ||
                   (src && ranges_overlap(src, sspan, o.dst, odspan));
    }
    if (conflict && (*result = kzn_flush_transfers(q))) return true;
    KznPendingTransfer& t = q->pending[q->n_pending++];
    t.dst = dst;
    t.src = src;
    t.bytes = bytes;
    t.value = value;
    t.row_bytes = (uint32_t)row_bytes;
    t.rows = (uint32_t)rows;
    t.dst_pitch = (uint32_t)dst_pitch;
    t.src_pitch = (uint32_t)src_pitch;
    return true;
}

KznResult launch_fill_contig(KznQueue* q, uint64_t address, uint64_t bytes, uint32_t value, bool may_merge) {
    if (bytes == 0) return KZN_SUCCESS;
    KznResult r;
    if (may_merge && merge_transfer(q, address, 0, bytes, value, &r)) return r;
    if (may_merge && (r = kzn_flush_transfers(q))) return r;
    bool aligned = (address % 16 == 0) && (bytes % 16 == 0);
    Path path = transfer_path();
    bool tma = aligned && (path == Path::Tma || (path == Path::Auto && bytes >= (64ull << 20)));
    if (tma) {
        static bool attr_set = false;
        if (!attr_set) {
            KZN_CUDA(cudaFuncSetAttribute(fill_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkChunk));
            attr_set = true;
        }
        uint64_t chunks = (bytes + kBulkChunk - 1) / kBulkChunk;
        uint64_t cap = (uint64_t)q->device->sm_count * 4;
        int grid = (int)(chunks < cap ? chunks : cap);
        kzn_timing_begin(q);
        // ordinary launch: with programmatic dependent launch back-to-back bulk fills measured slower (8K clear 92 -> 77 % of
        // HBM per submit), the staged CTAs of the next fill compete with the running one's bulk stores
        fill_bulk_kernel<<<grid, kThreads, kBulkChunk, q->stream>>>((uint8_t*)(uintptr_t)address, bytes, value);
        return kzn_launched(q, "fill_bulk_kernel");
    }
    uint64_t words = bytes / 4;
    kzn_timing_begin(q);
    kzn_launch_pdl(fill_contig_kernel, dim3(grid_for(q, words / 4 + 1)), dim3(kThreads), 0, q->stream, (uint32_t*)(uintptr_t)address, words, value);
    return kzn_launched(q, "fill_contig_kernel");
}

KznResult launch_copy_bulk(KznQueue* q, uint64_t dst, uint64_t src, uint64_t bytes) {
    KznResult r = kzn_flush_transfers(q);
    if (r) return r;
    static bool attr_set = false;
    const int smem = kCopyStages * kCopyChunk;
    if (!attr_set) {
        KZN_CUDA(cudaFuncSetAttribute(copy_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    uint64_t chunks = (bytes + kCopyChunk - 1) / kCopyChunk;
    uint64_t cap = q->device->sm_count;  // one 128 KiB ring per SM
    int grid = (int)(chunks < cap ? chunks : cap);
    kzn_timing_begin(q);
    copy_bulk_kernel<<<grid, 32, smem, q->stream>>>((uint8_t*)(uintptr_t)dst, (const uint8_t*)(uintptr_t)src, bytes);  // ordinary launch, as fill_bulk_kernel
    return kzn_launched(q, "copy_bulk_kernel");
}

struct Box {
    // one side of a box transfer, before lowering to Surf
    uint64_t base;         // surface origin (image address or buffer address + bufferOffset)
    bool tiled;
    uint64_t row_pitch;    // LINEAR bytes
    uint64_t slice_pitch;  // bytes
    uint32_t x0, y0;       // texel origin inside the surface
    uint64_t slice0;
    uint32_t tiles_x;
};

Surf lower(const Box& b) {
    Surf s;
    s.slice_pitch = b.slice_pitch;
    s.row_pitch = b.row_pitch;
    s.tiles_x = b.tiles_x;
    if (b.tiled) {
        s.base = b.base + b.slice0 * b.slice_pitch;
        s.x0 = b.x0;
        s.y0 = b.y0;
    } else {
        s.base = b.base + b.slice0 * b.slice_pitch + (uint64_t)b.y0 * b.row_pitch + (uint64_t)b.x0 * 4;
        s.x0 = s.y0 = 0;
    }
    return s;
}

bool vec4_ok(const Box& b, const Surf& s, uint32_t w) {
    if (w % 4) return false;
    if (s.base % 16 || s.slice_pitch % 16) return false;
    if (b.tiled) return b.x0 % 4 == 0;
    return s.row_pitch % 16 == 0;
}

template <int VEC, bool FILL>
KznResult launch_box_t(KznQueue* q, const BoxParams& p, bool st, bool dt) {
    KznResult r = kzn_flush_transfers(q);
    if (r) return r;
    int grid = grid_for(q, p.total);
#define KZN_BOX(ST, DT)                                                          \
    kzn_timing_begin(q);                                                          \
    kzn_launch_pdl(box_kernel<VEC, ST, DT, FILL>, dim3(grid), dim3(kThreads), 0, q->stream, p); \
    return kzn_launched(q, "box_kernel");
    if (st && dt) { KZN_BOX(true, true) }
    if (st) { KZN_BOX(true, false) }
    if (dt) { KZN_BOX(false, true) }
    KZN_BOX(false, false)
#undef KZN_BOX
}

// copy (fill == false) or fill a w x h x slices box
KznResult launch_box(KznQueue* q, const Box* src, const Box& dst, uint32_t w, uint32_t h, uint32_t slices,
                     bool fill, uint32_t fill_value) {
    if (!w || !h || !slices) return KZN_SUCCESS;
    BoxParams p;
    memset(&p, 0, sizeof p);
    p.dst = lower(dst);
    bool st = false, dt = dst.tiled;
    bool v4 = vec4_ok(dst, p.dst, w);
    if (!fill) {
        p.src = lower(*src);
        st = src->tiled;
        v4 = v4 && vec4_ok(*src, p.src, w);
    }
    uint64_t wt = w;  // texels per row after collapsing
    uint32_t hh = h, ss = slices;
    if (!st && !dt) {
        // collapse contiguous rows / slices into longer runs
        uint64_t row_bytes = (uint64_t)w * 4;
        bool rows_contig = hh == 1 || (p.dst.row_pitch == row_bytes && (fill || p.src.row_pitch == row_bytes));
        if (rows_contig && hh > 1) {
            wt *= hh;
            hh = 1;
        }
        if (hh == 1) {
            uint64_t run_bytes = wt * 4;
            bool slices_contig = ss == 1 || (p.dst.slice_pitch == run_bytes && (fill || p.src.slice_pitch == run_bytes));
            if (slices_contig && ss > 1) {
                wt *= ss;
                ss = 1;
            }
        }
        if (hh == 1 && ss == 1) {
            // one contiguous run
            uint64_t bytes = wt * 4;
            if (fill) return launch_fill_contig(q, p.dst.base, bytes, fill_value);
            Path path = transfer_path();
            bool aligned = p.dst.base % 16 == 0 && p.src.base % 16 == 0 && bytes % 16 == 0;
            KznResult mr;
            if (aligned && merge_transfer(q, p.dst.base, p.src.base, bytes, 0, &mr)) return mr;
            if (aligned && path == Path::Tma) return launch_copy_bulk(q, p.dst.base, p.src.base, bytes);
            if (!aligned && p.dst.base % 16 == 0 && p.src.base % 16 != 0 && p.src.base % 4 == 0 && bytes % 16 == 0) {
                KznResult fr = kzn_flush_transfers(q);
                if (fr) return fr;
                const uint64_t vecs = bytes / 16;
                kzn_timing_begin(q);
                kzn_launch_pdl(copy_realign_kernel, dim3(grid_for(q, vecs * kUnroll)), dim3(kThreads), 0, q->stream,
                               (uint4*)(uintptr_t)p.dst.base, (const uint32_t*)(uintptr_t)p.src.base, vecs);
                return kzn_launched(q, "copy_realign_kernel");
            }
        }
    }
    if (!st && !dt && v4 && ss == 1 && hh > 1 && wt * 4 <= 0xffffffffull) {  // pitched rows of LINEAR memory: may join a merged launch
        KznResult mr;
        if (merge_transfer(q, p.dst.base, fill ? 0 : p.src.base, wt * 4 * hh, fill_value, &mr, wt * 4, hh, p.dst.row_pitch,
                           fill ? p.dst.row_pitch : p.src.row_pitch))
            return mr;
    }
    p.h = hh;
    p.slices = ss;
    p.fill_value = fill_value;
    if (v4) {
        KZN_REQUIRE(wt / 4 <= 0xffffffffull || (hh == 1 && ss == 1), KZN_ERROR_FEATURE_NOT_PRESENT, "row too long");
        p.wv = (uint32_t)((wt / 4) & 0xffffffffull);
        p.total = (wt / 4) * hh * ss;
        return fill ? launch_box_t<4, true>(q, p, st, dt) : launch_box_t<4, false>(q, p, st, dt);
    }
    KZN_REQUIRE(wt <= 0xffffffffull || (hh == 1 && ss == 1), KZN_ERROR_FEATURE_NOT_PRESENT, "row too long");
    p.wv = (uint32_t)(wt & 0xffffffffull);
    p.total = wt * hh * ss;
    return fill ? launch_box_t<1, true>(q, p, st, dt) : launch_box_t<1, false>(q, p, st, dt);
}

Box image_box(const KznImageDesc* im, int32_t x, int32_t y, int32_t z, uint32_t base_layer) {
    Box b;
    b.base = im->address;
    b.tiled = im->tiling == KZN_TILING_TILED;
    b.row_pitch = kzn_row_pitch(im);
    b.slice_pitch = kzn_slice_pitch(im);
    b.x0 = (uint32_t)x;
    b.y0 = (uint32_t)y;
    b.slice0 = (uint64_t)z + (uint64_t)im->depth * base_layer;
    b.tiles_x = kzn_tiles(im->width);
    return b;
}

uint32_t aspect_of(uint32_t format) {
    return format == KZN_FORMAT_D32_SFLOAT ? KZN_IMAGE_ASPECT_DEPTH_BIT : KZN_IMAGE_ASPECT_COLOR_BIT;
}

KznResult resolve_range(const KznImageDesc* im, const KznImageSubresourceRange* r, uint32_t aspect, uint32_t* l0, uint32_t* n) {
    KZN_VALID(r->aspectMask == aspect, "clear: aspectMask %#x must be %#x (VUID 02498)", r->aspectMask, aspect);
    KZN_VALID(r->baseMipLevel == 0, "clear: baseMipLevel %u out of range (mip_levels is 1)", r->baseMipLevel);
    KZN_VALID(r->levelCount == KZN_REMAINING_MIP_LEVELS || r->levelCount == 1, "clear: levelCount %u out of range", r->levelCount);
    KZN_VALID(r->baseArrayLayer < im->array_layers, "clear: baseArrayLayer %u out of range", r->baseArrayLayer);
    uint32_t cnt = r->layerCount == KZN_REMAINING_ARRAY_LAYERS ? im->array_layers - r->baseArrayLayer : r->layerCount;
    KZN_VALID(cnt != 0 && cnt <= im->array_layers - r->baseArrayLayer, "clear: layerCount %u out of range", r->layerCount);
    *l0 = r->baseArrayLayer;
    *n = cnt;
    return KZN_SUCCESS;
}

KznResult clear_image(KznQueue* q, const KznImageDesc* im, uint32_t word, uint32_t aspect, uint32_t range_count,
                      const KznImageSubresourceRange* ranges) {
    for (uint32_t i = 0; i < range_count; ++i) {
        uint32_t l0, n;
        KznResult r = resolve_range(im, &ranges[i], aspect, &l0, &n);
        if (r) return r;
    }
    uint64_t layer_pitch = kzn_slice_pitch(im) * im->depth;
    for (uint32_t i = 0; i < range_count; ++i) {
        uint32_t l0 = 0, n = 0;
        resolve_range(im, &ranges[i], aspect, &l0, &n);
        // whole layers are contiguous in both layouts (E.2; TILED padding texels are written too)
        KznResult r = launch_fill_contig(q, im->address + layer_pitch * l0, layer_pitch * n, word);
        if (r) return r;
    }
    return KZN_SUCCESS;
}

KznResult check_subresource_box(const KznImageDesc* im, const KznImageSubresourceLayers* s, const KznOffset3D* off,
                                const KznExtent3D* ext, const char* what) {
    KZN_VALID(s->aspectMask == aspect_of(im->format), "%s: aspectMask %#x does not match the format (VUID 00212)", what, s->aspectMask);
    KZN_VALID(s->mipLevel == 0, "%s: mipLevel %u out of range", what, s->mipLevel);
    KZN_VALID(s->layerCount != 0 && (uint64_t)s->baseArrayLayer + s->layerCount <= im->array_layers, "%s: layers out of range", what);
    KZN_VALID(off->x >= 0 && off->y >= 0 && off->z >= 0, "%s: negative image offset", what);
    KZN_VALID((uint64_t)off->x + ext->width <= im->width, "%s: x range outside the image (VUID 00197)", what);
    KZN_VALID((uint64_t)off->y + ext->height <= im->height, "%s: y range outside the image (VUID 00198)", what);
    KZN_VALID((uint64_t)off->z + ext->depth <= im->depth, "%s: z range outside the image (VUID 00200)", what);
    KZN_VALID(!(im->depth > 1 && (s->baseArrayLayer != 0 || s->layerCount != 1)), "%s: 3D images have one layer", what);
    return KZN_SUCCESS;
}

KznResult check_buffer_image_region(const KznImageDesc* im, uint64_t buffer_size, const KznBufferImageCopy* c, const char* what) {
    KznResult r = check_subresource_box(im, &c->imageSubresource, &c->imageOffset, &c->imageExtent, what);
    if (r) return r;
    uint64_t ts = kzn_texel_size(im->format);
    KZN_VALID(c->bufferOffset % 4 == 0 && c->bufferOffset % ts == 0, "%s: bufferOffset %llu not a multiple of 4 (VUID 00194)", what,
              (unsigned long long)c->bufferOffset);
    KZN_VALID(!c->bufferRowLength || c->bufferRowLength >= c->imageExtent.width, "%s: bufferRowLength < width (VUID 00195)", what);
    KZN_VALID(!c->bufferImageHeight || c->bufferImageHeight >= c->imageExtent.height, "%s: bufferImageHeight < height (VUID 00196)", what);
    uint64_t w = c->imageExtent.width, h = c->imageExtent.height;
    uint64_t slices = (uint64_t)c->imageExtent.depth * c->imageSubresource.layerCount;
    if (w && h && slices) {
        uint64_t row_len = c->bufferRowLength ? c->bufferRowLength : w;
        uint64_t img_h = c->bufferImageHeight ? c->bufferImageHeight : h;
        // 128-bit: bufferRowLength and bufferImageHeight are free 32-bit values, the product may pass 2^64
        const unsigned __int128 end = (unsigned __int128)c->bufferOffset +
                                      ((((unsigned __int128)(slices - 1) * img_h) + (h - 1)) * row_len + w) * ts;
        KZN_VALID(end <= buffer_size, "%s: region needs more bytes than the buffer has (%llu) (VUID 00171)", what,
                  (unsigned long long)buffer_size);
    }
    return KZN_SUCCESS;
}

Box buffer_box(uint64_t buffer_address, const KznBufferImageCopy* c) {
    uint64_t row_len = c->bufferRowLength ? c->bufferRowLength : c->imageExtent.width;
    uint64_t img_h = c->bufferImageHeight ? c->bufferImageHeight : c->imageExtent.height;
    Box b;
    b.base = buffer_address + c->bufferOffset;
    b.tiled = false;
    b.row_pitch = row_len * 4;
    b.slice_pitch = row_len * img_h * 4;
    b.x0 = b.y0 = 0;
    b.slice0 = 0;
    b.tiles_x = 0;
    return b;
}

KznResult copy_buffer_image(KznQueue* q, uint64_t buffer_address, uint64_t buffer_size, const KznImageDesc* im,
                            uint32_t region_count, const KznBufferImageCopy* regions, bool to_image, const char* what) {
    KznResult r = kzn_queue_enter_transfer(q);
    if (r) return r;
    if ((r = kzn_check_image(im, what))) return r;
    KZN_VALID(buffer_address && buffer_address % 4 == 0, "%s: buffer address must be 4-byte aligned", what);
    KZN_VALID(region_count && regions, "%s: regionCount must be > 0", what);
    for (uint32_t i = 0; i < region_count; ++i)
        if ((r = check_buffer_image_region(im, buffer_size, &regions[i], what))) return r;
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznBufferImageCopy* c = &regions[i];
        Box ib = image_box(im, c->imageOffset.x, c->imageOffset.y, c->imageOffset.z, c->imageSubresource.baseArrayLayer);
        Box bb = buffer_box(buffer_address, c);
        uint32_t slices = c->imageExtent.depth * c->imageSubresource.layerCount;
        r = to_image ? launch_box(q, &bb, ib, c->imageExtent.width, c->imageExtent.height, slices, false, 0)
                     : launch_box(q, &ib, bb, c->imageExtent.width, c->imageExtent.height, slices, false, 0);
        if (r) return r;
    }
    return KZN_SUCCESS;
}

}  // namespace

// Capture the pending batch: one command goes down its usual path, several become one merged launch.
KznResult kzn_flush_transfers(KznQueue* q) {
    const uint32_t n = q->n_pending;
    if (n == 0) return KZN_SUCCESS;
    q->n_pending = 0;  // first: the launch helpers below call back into this function
    if (n == 1) {
        const KznPendingTransfer t = q->pending[0];
        if (t.row_bytes) {  // a pitched box on its own: the box kernel, as if it had not been queued
            BoxParams p;
            memset(&p, 0, sizeof p);
            p.src.base = t.src;
            p.dst.base = t.dst;
            p.src.row_pitch = t.src_pitch;
            p.dst.row_pitch = t.dst_pitch;
            p.h = t.rows;
            p.slices = 1;
            p.fill_value = t.value;
            p.wv = t.row_bytes / 16;
            p.total = (unsigned long long)p.wv * t.rows;
            return t.src ? launch_box_t<4, false>(q, p, false, false) : launch_box_t<4, true>(q, p, false, false);
        }
        if (t.src == 0) return launch_fill_contig(q, t.dst, t.bytes, t.value, false);
        if (transfer_path() == Path::Tma) return launch_copy_bulk(q, t.dst, t.src, t.bytes);
        Box sb = {t.src, false, t.bytes, t.bytes, 0, 0, 0, 0}, db = {t.dst, false, t.bytes, t.bytes, 0, 0, 0, 0};
        BoxParams p;
        memset(&p, 0, sizeof p);
        p.src = lower(sb);
        p.dst = lower(db);
        p.h = p.slices = 1;
        p.wv = (uint32_t)(t.bytes / 16);
        p.total = t.bytes / 16;
        return launch_box_t<4, false>(q, p, false, false);
    }
    MergedOps m;
    memset(&m, 0, sizeof m);
    m.n = n;
    uint32_t units = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const KznPendingTransfer& t = q->pending[i];
        m.first_unit[i] = units;
        m.vecs[i] = (uint32_t)(t.bytes / 16);
        m.value[i] = t.value;
        m.wv[i] = t.row_bytes / 16;
        m.dst_pitch[i] = t.dst_pitch;
        m.src_pitch[i] = t.src_pitch;
        m.dst[i] = t.dst;
        m.src[i] = t.src;
        units += (m.vecs[i] + kUnitVecs - 1) / kUnitVecs;
    }
    m.first_unit[n] = units;
    m.total_units = units;
    const uint32_t cap = (uint32_t)q->device->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
    bool boxes = false;
    for (uint32_t i = 0; i < n; ++i) boxes = boxes || m.wv[i];
    const dim3 grid(units < cap ? units : cap);
    if (boxes) kzn_launch_pdl(merged_transfer_kernel<true>, grid, dim3(kThreads), 0, q->stream, m);
    else kzn_launch_pdl(merged_transfer_kernel<false>, grid, dim3(kThreads), 0, q->stream, m);
    q->merged_commands += n;
    return kzn_launched(q, "merged_transfer_kernel");
}

// vkCmdBeginRenderPass left a LOAD_OP_CLEAR pending and no draw took it over: the clear as box fills (what
// vkCmdClearAttachments on the same rectangle does).
KznResult kzn_flush_render_pass_clear(KznQueue* q) {
    if (!q->pending_clear.active) return KZN_SUCCESS;
    const KznPendingClear c = q->pending_clear;
    q->pending_clear.active = false;  // first: the launch helpers enter the queue again
    if (c.x1 < c.x0 || c.y1 < c.y0) return KZN_SUCCESS;
    const uint32_t w = (uint32_t)(c.x1 - c.x0 + 1), h = (uint32_t)(c.y1 - c.y0 + 1);
    KznResult r;
    if (c.load & 1u) {
        Box b = image_box(&c.fb.color, c.x0, c.y0, 0, c.fb.color_layer);
        if ((r = launch_box(q, nullptr, b, w, h, 1, true, c.color_word))) return r;
    }
    if (c.load & 2u) {
        Box b = image_box(&c.fb.depth, c.x0, c.y0, 0, c.fb.depth_layer);
        if ((r = launch_box(q, nullptr, b, w, h, 1, true, c.depth_bits))) return r;
    }
    return KZN_SUCCESS;
}

// ---- C ABI ----------------------------------------------------------------------------------------------------

extern "C" {

KznResult kzn_cmd_clear_color_image(KznQueue* queue, const KznImageDesc* image, const KznClearColorValue* color,
                                    uint32_t range_count, const KznImageSubresourceRange* ranges) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    if ((r = kzn_check_image(image, "vkCmdClearColorImage"))) return r;
    KZN_REQUIRE(image->format == KZN_FORMAT_R8G8B8A8_UNORM, KZN_ERROR_FORMAT_NOT_SUPPORTED,
                "vkCmdClearColorImage: format %u is not a colour format (VUID 00007)", image->format);
    KZN_VALID(color && range_count && ranges, "vkCmdClearColorImage: null argument or rangeCount == 0");
    return clear_image(queue, image, kzn_pack_rgba8_host(color->float32), KZN_IMAGE_ASPECT_COLOR_BIT, range_count, ranges);
}

KznResult kzn_cmd_clear_depth_stencil_image(KznQueue* queue, const KznImageDesc* image, const KznClearDepthStencilValue* value,
                                            uint32_t range_count, const KznImageSubresourceRange* ranges) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    if ((r = kzn_check_image(image, "vkCmdClearDepthStencilImage"))) return r;
    KZN_REQUIRE(image->format == KZN_FORMAT_D32_SFLOAT, KZN_ERROR_FORMAT_NOT_SUPPORTED,
                "vkCmdClearDepthStencilImage: format %u is not a depth format", image->format);
    KZN_VALID(value && range_count && ranges, "vkCmdClearDepthStencilImage: null argument or rangeCount == 0");
    uint32_t word;
    memcpy(&word, &value->depth, 4);
    return clear_image(queue, image, word, KZN_IMAGE_ASPECT_DEPTH_BIT, range_count, ranges);
}

KznResult kzn_cmd_clear_attachment(KznQueue* queue, const KznImageDesc* image, const KznClearColorValue* v,
                                   uint32_t rect_count, const KznClearRect* rects) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    if ((r = kzn_check_image(image, "vkCmdClearAttachments"))) return r;
    KZN_VALID(v && rect_count && rects, "vkCmdClearAttachments: null argument or rectCount == 0");
    KZN_VALID(image->depth == 1, "vkCmdClearAttachments: attachment must be 2D");
    uint32_t word = image->format == KZN_FORMAT_R8G8B8A8_UNORM ? kzn_pack_rgba8_host(v->float32) : v->uint32[0];
    for (uint32_t i = 0; i < rect_count; ++i) {
        const KznClearRect* c = &rects[i];
        KZN_VALID(c->rect.offset.x >= 0 && c->rect.offset.y >= 0, "vkCmdClearAttachments: negative rect offset");
        KZN_VALID((uint64_t)c->rect.offset.x + c->rect.extent.width <= image->width, "vkCmdClearAttachments: rect outside attachment");
        KZN_VALID((uint64_t)c->rect.offset.y + c->rect.extent.height <= image->height, "vkCmdClearAttachments: rect outside attachment");
        KZN_VALID(c->layerCount != 0 && (uint64_t)c->baseArrayLayer + c->layerCount <= image->array_layers,
                  "vkCmdClearAttachments: layers out of range");
    }
    for (uint32_t i = 0; i < rect_count; ++i) {
        const KznClearRect* c = &rects[i];
        Box b = image_box(image, c->rect.offset.x, c->rect.offset.y, 0, c->baseArrayLayer);
        if ((r = launch_box(queue, nullptr, b, c->rect.extent.width, c->rect.extent.height, c->layerCount, true, word))) return r;
    }
    return KZN_SUCCESS;
}

KznResult kzn_cmd_begin_render_pass(KznQueue* queue, const KznRenderPassBegin* begin) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    KZN_VALID(begin, "vkCmdBeginRenderPass: null argument");
    KZN_VALID(!queue->in_render_pass, "vkCmdBeginRenderPass: a render pass is already active on this queue (VUID-vkCmdBeginRenderPass-renderpass)");
    const KznFramebuffer* fb = &begin->framebuffer;
    const bool has_color = fb->color.format != 0, has_depth = fb->depth.format != 0;
    KZN_VALID(has_color || has_depth, "vkCmdBeginRenderPass: the framebuffer has no attachment");
    KZN_VALID(begin->color_load_op <= KZN_ATTACHMENT_LOAD_OP_CLEAR && begin->depth_load_op <= KZN_ATTACHMENT_LOAD_OP_CLEAR, "vkCmdBeginRenderPass: bad loadOp");
    uint32_t W = 0, H = 0;
    if (has_color) {
        if ((r = kzn_check_image(&fb->color, "vkCmdBeginRenderPass"))) return r;
        KZN_REQUIRE(fb->color.format == KZN_FORMAT_R8G8B8A8_UNORM, KZN_ERROR_FORMAT_NOT_SUPPORTED, "vkCmdBeginRenderPass: colour attachment must be R8G8B8A8_UNORM");
        KZN_VALID(fb->color.depth == 1 && fb->color_layer < fb->color.array_layers, "vkCmdBeginRenderPass: colour attachment must be a layer of a 2D image");
        W = fb->color.width;
        H = fb->color.height;
    }
    if (has_depth) {
        if ((r = kzn_check_image(&fb->depth, "vkCmdBeginRenderPass"))) return r;
        KZN_REQUIRE(fb->depth.format == KZN_FORMAT_D32_SFLOAT && fb->depth.tiling == KZN_TILING_LINEAR, KZN_ERROR_FORMAT_NOT_SUPPORTED,
                    "vkCmdBeginRenderPass: depth attachment must be LINEAR D32_SFLOAT");
        KZN_VALID(fb->depth.depth == 1 && fb->depth_layer < fb->depth.array_layers, "vkCmdBeginRenderPass: depth attachment must be a layer of a 2D image");
        KZN_VALID(!has_color || (fb->depth.width == W && fb->depth.height == H), "vkCmdBeginRenderPass: attachments differ in size");
        W = fb->depth.width;
        H = fb->depth.height;
    }
    const KznRect2D* a = &begin->render_area;
    KZN_VALID(a->offset.x >= 0 && a->offset.y >= 0 && (uint64_t)a->offset.x + a->extent.width <= W && (uint64_t)a->offset.y + a->extent.height <= H,
              "vkCmdBeginRenderPass: renderArea outside the framebuffer (VUID-VkRenderPassBeginInfo-renderArea-02848)");
    KZN_VALID(fb->band_y1 == 0 || (fb->band_y0 < fb->band_y1 && fb->band_y1 <= H), "vkCmdBeginRenderPass: bad band");
    queue->in_render_pass = true;
    KznPendingClear& c = queue->pending_clear;
    c.load = (has_color && begin->color_load_op == KZN_ATTACHMENT_LOAD_OP_CLEAR ? 1u : 0u) |
             (has_depth && begin->depth_load_op == KZN_ATTACHMENT_LOAD_OP_CLEAR ? 2u : 0u);
    if (!c.load) return KZN_SUCCESS;
    c.fb = *fb;
    c.x0 = a->offset.x;
    c.x1 = a->offset.x + (int32_t)a->extent.width - 1;
    c.y0 = a->offset.y;
    c.y1 = a->offset.y + (int32_t)a->extent.height - 1;
    if (fb->band_y1) {  // a band of a split frame owns rows [band_y0, band_y1) only
        if (c.y0 < (int32_t)fb->band_y0) c.y0 = (int32_t)fb->band_y0;
        if ((int64_t)c.y1 > (int64_t)fb->band_y1 - 1) c.y1 = (int32_t)((int64_t)fb->band_y1 - 1);
    }
    c.color_word = kzn_pack_rgba8_host(begin->clear_color.float32);
    memcpy(&c.depth_bits, &begin->clear_depth_stencil.depth, 4);
    c.active = true;
    return KZN_SUCCESS;
}

KznResult kzn_cmd_end_render_pass(KznQueue* queue) {
    KznResult r = kzn_queue_enter_transfer(queue);  // (performs a clear that is still pending: a render pass without a qualifying draw)
    if (r) return r;
    KZN_VALID(queue->in_render_pass, "vkCmdEndRenderPass: no render pass is active (VUID-vkCmdEndRenderPass-renderpass)");
    queue->in_render_pass = false;
    return KZN_SUCCESS;
}

KznResult kzn_cmd_copy_buffer_to_image(KznQueue* queue, uint64_t buffer_address, uint64_t buffer_size, const KznImageDesc* image,
                                       uint32_t region_count, const KznBufferImageCopy* regions) {
    return copy_buffer_image(queue, buffer_address, buffer_size, image, region_count, regions, true, "vkCmdCopyBufferToImage");
}

KznResult kzn_cmd_copy_image_to_buffer(KznQueue* queue, const KznImageDesc* image, uint64_t buffer_address, uint64_t buffer_size,
                                       uint32_t region_count, const KznBufferImageCopy* regions) {
    return copy_buffer_image(queue, buffer_address, buffer_size, image, region_count, regions, false, "vkCmdCopyImageToBuffer");
}

KznResult kzn_cmd_copy_buffer(KznQueue* queue, uint64_t src_address, uint64_t src_size, uint64_t dst_address, uint64_t dst_size,
                              uint32_t region_count, const KznBufferCopy* regions) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    KZN_VALID(src_address && dst_address && region_count && regions, "vkCmdCopyBuffer: null argument or regionCount == 0");
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznBufferCopy* c = &regions[i];
        KZN_VALID(c->size != 0, "vkCmdCopyBuffer: region size must be > 0 (VUID 01988)");
        KZN_VALID(c->srcOffset <= src_size && c->size <= src_size - c->srcOffset, "vkCmdCopyBuffer: source range outside the buffer");
        KZN_VALID(c->dstOffset <= dst_size && c->size <= dst_size - c->dstOffset, "vkCmdCopyBuffer: destination range outside the buffer");
    }
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznBufferCopy* c = &regions[i];
        uint64_t s = src_address + c->srcOffset, d = dst_address + c->dstOffset;
        if (s % 4 == 0 && d % 4 == 0 && c->size % 4 == 0) {
            Box sb = {s, false, c->size, c->size, 0, 0, 0, 0}, db = {d, false, c->size, c->size, 0, 0, 0, 0};
            // a buffer is a 1-row box of size/4 "texels"
            uint64_t texels = c->size / 4;
            KZN_REQUIRE(texels <= 0xffffffffull, KZN_ERROR_FEATURE_NOT_PRESENT, "vkCmdCopyBuffer: region larger than 16 GiB");
            if ((r = launch_box(queue, &sb, db, (uint32_t)texels, 1, 1, false, 0))) return r;
        } else {
            // byte-granular regions are legal for vkCmdCopyBuffer: the copy engine handles them
            if ((r = kzn_flush_transfers(queue))) return r;
            KZN_CUDA(cudaMemcpyAsync((void*)(uintptr_t)d, (const void*)(uintptr_t)s, c->size, cudaMemcpyDeviceToDevice, queue->stream));
        }
    }
    return KZN_SUCCESS;
}

KznResult kzn_cmd_fill_buffer(KznQueue* queue, uint64_t dst_address, uint64_t dst_size, uint64_t offset, uint64_t size, uint32_t data) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    KZN_VALID(dst_address && dst_address % 4 == 0, "vkCmdFillBuffer: bad buffer address");
    KZN_VALID(offset < dst_size && offset % 4 == 0, "vkCmdFillBuffer: dstOffset must be < size and a multiple of 4 (VUID 00024/00025)");
    if (size == KZN_WHOLE_SIZE) size = (dst_size - offset) & ~3ull;
    else KZN_VALID(size != 0 && size % 4 == 0 && size <= dst_size - offset, "vkCmdFillBuffer: bad size (VUID 00026-00028)");
    return launch_fill_contig(queue, dst_address + offset, size, data);
}

KznResult kzn_cmd_update_buffer(KznQueue* queue, uint64_t dst_address, uint64_t dst_size, uint64_t offset, uint64_t size, const void* data) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(dst_address && data, "vkCmdUpdateBuffer: null argument");
    KZN_VALID(offset % 4 == 0 && size % 4 == 0 && size != 0 && size <= 65536, "vkCmdUpdateBuffer: dstOffset/dataSize rules (VUID 00036-00038)");
    KZN_VALID(offset < dst_size && size <= dst_size - offset, "vkCmdUpdateBuffer: range outside the buffer");
    // "the data is copied at record time": outside a command buffer the runtime stages the pageable source before
    // returning; inside one the command buffer keeps its own copy for the replays
    const void* src = kzn_recorded_host_copy(queue, data, size);
    KZN_REQUIRE(src, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    KZN_CUDA(cudaMemcpyAsync((void*)(uintptr_t)(dst_address + offset), src, size, cudaMemcpyHostToDevice, queue->stream));
    return KZN_SUCCESS;
}

KznResult kzn_cmd_copy_image(KznQueue* queue, const KznImageDesc* src, const KznImageDesc* dst, uint32_t region_count, const KznImageCopy* regions) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    if ((r = kzn_check_image(src, "vkCmdCopyImage(src)")) || (r = kzn_check_image(dst, "vkCmdCopyImage(dst)"))) return r;
    KZN_VALID(region_count && regions, "vkCmdCopyImage: regionCount must be > 0");
    KZN_VALID(kzn_texel_size(src->format) == kzn_texel_size(dst->format), "vkCmdCopyImage: formats are not size-compatible");
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageCopy* c = &regions[i];
        if ((r = check_subresource_box(src, &c->srcSubresource, &c->srcOffset, &c->extent, "vkCmdCopyImage(src)"))) return r;
        if ((r = check_subresource_box(dst, &c->dstSubresource, &c->dstOffset, &c->extent, "vkCmdCopyImage(dst)"))) return r;
        KZN_VALID(c->srcSubresource.layerCount == c->dstSubresource.layerCount, "vkCmdCopyImage: layerCount mismatch");
    }
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageCopy* c = &regions[i];
        Box sb = image_box(src, c->srcOffset.x, c->srcOffset.y, c->srcOffset.z, c->srcSubresource.baseArrayLayer);
        Box db = image_box(dst, c->dstOffset.x, c->dstOffset.y, c->dstOffset.z, c->dstSubresource.baseArrayLayer);
        if ((r = launch_box(queue, &sb, db, c->extent.width, c->extent.height, c->extent.depth * c->srcSubresource.layerCount, false, 0))) return r;
    }
    return KZN_SUCCESS;
}

static KznResult check_blit_box(const KznImageDesc* im, const KznImageSubresourceLayers* s, const KznOffset3D o[2], const char* what) {
    KZN_VALID(s->aspectMask == KZN_IMAGE_ASPECT_COLOR_BIT && s->mipLevel == 0, "%s: bad subresource", what);
    KZN_VALID(s->layerCount != 0 && (uint64_t)s->baseArrayLayer + s->layerCount <= im->array_layers, "%s: layers out of range", what);
    for (int k = 0; k < 2; ++k) {
        KZN_VALID(o[k].x >= 0 && (uint32_t)o[k].x <= im->width, "%s: x offset outside [0,width] (VUID 00243)", what);
        KZN_VALID(o[k].y >= 0 && (uint32_t)o[k].y <= im->height, "%s: y offset outside [0,height] (VUID 00244)", what);
        KZN_VALID(o[k].z >= 0 && (uint32_t)o[k].z <= im->depth, "%s: z offset outside [0,depth] (VUID 00246)", what);
    }
    KZN_VALID(!(im->depth > 1 && (s->baseArrayLayer != 0 || s->layerCount != 1)), "%s: 3D images have one layer", what);
    return KZN_SUCCESS;
}

static float blit_coord_host(int32_t i, int32_t d0, int32_t d1, int32_t s0, int32_t s1) {
    float scale = (float)(s1 - s0) / (float)(d1 - d0);
    return __builtin_fmaf(((float)i + 0.5f) - (float)d0, scale, (float)s0);
}

KznResult kzn_cmd_blit_image(KznQueue* queue, const KznImageDesc* src, const KznImageDesc* dst, uint32_t region_count,
                             const KznImageBlit* regions, uint32_t filter) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    if ((r = kzn_check_image(src, "vkCmdBlitImage(src)")) || (r = kzn_check_image(dst, "vkCmdBlitImage(dst)"))) return r;
    KZN_REQUIRE(src->format == KZN_FORMAT_R8G8B8A8_UNORM && dst->format == KZN_FORMAT_R8G8B8A8_UNORM, KZN_ERROR_FORMAT_NOT_SUPPORTED,
                "vkCmdBlitImage: only R8G8B8A8_UNORM has BLIT_SRC/BLIT_DST features (VUID 01999/02000)");
    KZN_VALID(region_count && regions && filter <= KZN_FILTER_LINEAR, "vkCmdBlitImage: bad argument");
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageBlit* b = &regions[i];
        if ((r = check_blit_box(src, &b->srcSubresource, b->srcOffsets, "vkCmdBlitImage(src)"))) return r;
        if ((r = check_blit_box(dst, &b->dstSubresource, b->dstOffsets, "vkCmdBlitImage(dst)"))) return r;
        KZN_VALID(b->srcSubresource.layerCount == b->dstSubresource.layerCount, "vkCmdBlitImage: layerCount mismatch");
    }
    for (uint32_t i = 0; i < region_count; ++i) {
        const KznImageBlit* b = &regions[i];
        const KznOffset3D *s0 = &b->srcOffsets[0], *s1 = &b->srcOffsets[1], *d0 = &b->dstOffsets[0], *d1 = &b->dstOffsets[1];
        if (d0->x == d1->x || d0->y == d1->y || d0->z == d1->z) continue;
        BlitParams p;
        memset(&p, 0, sizeof p);
        p.src_tiled = src->tiling == KZN_TILING_TILED;
        p.dst_tiled = dst->tiling == KZN_TILING_TILED;
        p.src_w = (int32_t)src->width;
        p.src_h = (int32_t)src->height;
        p.dx0 = d0->x; p.dx1 = d1->x; p.dy0 = d0->y; p.dy1 = d1->y;
        p.sx0 = s0->x; p.sx1 = s1->x; p.sy0 = s0->y; p.sy1 = s1->y;
        p.x_lo = d0->x < d1->x ? d0->x : d1->x;
        p.y_lo = d0->y < d1->y ? d0->y : d1->y;
        p.w = (uint32_t)(d0->x < d1->x ? d1->x - d0->x : d0->x - d1->x);
        p.h = (uint32_t)(d0->y < d1->y ? d1->y - d0->y : d0->y - d1->y);
        p.filter = filter;
        p.scale_x = (float)(p.sx1 - p.sx0) / (float)(p.dx1 - p.dx0);
        p.scale_y = (float)(p.sy1 - p.sy0) / (float)(p.dy1 - p.dy0);
        int32_t z_lo = d0->z < d1->z ? d0->z : d1->z, z_hi = d0->z < d1->z ? d1->z : d0->z;
        for (uint32_t l = 0; l < b->srcSubresource.layerCount; ++l)
            for (int32_t k = z_lo; k < z_hi; ++k) {
                // third coordinate point-sampled on the host: one launch per destination slice
                float wf = blit_coord_host(k, d0->z, d1->z, s0->z, s1->z);
                int32_t sz = (int32_t)__builtin_floorf(wf);
                sz = sz < 0 ? 0 : (sz > (int32_t)src->depth - 1 ? (int32_t)src->depth - 1 : sz);
                if (filter == KZN_FILTER_NEAREST && src->address != dst->address && d1->x > d0->x && d1->y > d0->y && d1->x - d0->x == s1->x - s0->x &&
                    d1->y - d0->y == s1->y - s0->y) {
                    // same size, not mirrored, point-sampled: texel (i, j) of the destination box reads floor(i + 0.5 - d0 + s0)
                    // = i - d0 + s0 — a box copy (16-byte vectors where the boxes allow; joins a merged launch when recorded)
                    Box sb = image_box(src, s0->x, s0->y, sz, b->srcSubresource.baseArrayLayer + l);
                    Box db = image_box(dst, d0->x, d0->y, k, b->dstSubresource.baseArrayLayer + l);
                    if ((r = launch_box(queue, &sb, db, p.w, p.h, 1, false, 0))) return r;
                    continue;
                }
                p.src = lower(image_box(src, 0, 0, sz, b->srcSubresource.baseArrayLayer + l));
                p.dst = lower(image_box(dst, 0, 0, k, b->dstSubresource.baseArrayLayer + l));
                dim3 grid((p.w + 127) / 128, (p.h + kThreads / 32 - 1) / (kThreads / 32));
                kzn_timing_begin(queue);
                kzn_launch_pdl(blit_kernel_for(p.src_tiled != 0, p.dst_tiled != 0, filter != KZN_FILTER_NEAREST), grid, dim3(kThreads), 0, queue->stream, p);
                if ((r = kzn_launched(queue, "blit_kernel"))) return r;
            }
    }
    return KZN_SUCCESS;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// LINEAR <-> TILED copy of one whole slice through the TMA unit; false: not applicable (the caller uses box_kernel)
bool swizzle_tma(KznQueue* q, const KznImageDesc* linear, uint64_t linear_base, uint64_t tiled_base, bool to_tiled, KznResult* result) {
    *result = KZN_SUCCESS;
    const uint32_t w = linear->width, h = linear->height;
    // whole tiles only: a partial tile would have its padding written by the tensor copy (zero fill), which box_kernel never
    // touches; the tensor map needs 16-byte aligned bases and pitches
    if (transfer_path() != Path::Tma || w % 16 || h % 16 || linear_base % 16 || tiled_base % 16 || (uint64_t)w * h < (1u << 16)) return false;
    EncodeTiledFn encode = encode_tiled();
    if (!encode) return false;
    const uint32_t tiles_x = w / 16, tiles_y = h / 16;
    const uint64_t pitch = (uint64_t)w * 4;
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[4] = {16, 16, tiles_x, tiles_y};
    const cuuint64_t strides[3] = {pitch, 64, 16 * pitch};  // bytes, of dimensions 1..3
    const cuuint32_t box[4] = {16, 16, (cuuint32_t)(tiles_x < (uint32_t)kSwzTiles ? tiles_x : kSwzTiles), 1};
    const cuuint32_t elem[4] = {1, 1, 1, 1};
    if (pitch >= (1ull << 40) / 16) return false;
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)(uintptr_t)linear_base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (tiles_x < (uint32_t)kSwzTiles) return false;  // (the kernel's stages are kSwzTiles wide)
    if ((*result = kzn_flush_transfers(q))) return true;
    static bool attr_set = false;
    const int smem = kSwzStages * (int)kSwzBytes;
    if (!attr_set) {
        cudaFuncSetAttribute(swizzle_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(swizzle_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    const uint32_t groups_x = (tiles_x + kSwzTiles - 1) / kSwzTiles;
    const uint64_t groups = (uint64_t)groups_x * tiles_y;
    const uint64_t cap = (uint64_t)q->device->sm_count * 4;  // 4 rings of 48 KiB per SM
    const int grid = (int)(groups < cap ? groups : cap);
    kzn_timing_begin(q);
    if (to_tiled) swizzle_tma_kernel<true><<<grid, 32, smem, q->stream>>>(map, (uint8_t*)(uintptr_t)tiled_base, tiles_x, groups_x, groups);
    else swizzle_tma_kernel<false><<<grid, 32, smem, q->stream>>>(map, (uint8_t*)(uintptr_t)tiled_base, tiles_x, groups_x, groups);
    *result = kzn_launched(q, "swizzle_tma_kernel");
    return true;
}

}  // namespace

KznResult kzn_cmd_convert_tiling(KznQueue* queue, const KznImageDesc* src, const KznImageDesc* dst) {
    KznResult r = kzn_queue_enter_transfer(queue);
    if (r) return r;
    if ((r = kzn_check_image(src, "convert_tiling(src)")) || (r = kzn_check_image(dst, "convert_tiling(dst)"))) return r;
    KZN_VALID(src->format == dst->format && src->width == dst->width && src->height == dst->height && src->depth == dst->depth &&
                  src->array_layers == dst->array_layers,
              "convert_tiling: images differ in more than tiling");
    if (src->tiling != dst->tiling) {
        // every slice is one LINEAR <-> TILED copy of whole tiles: the TMA unit does the swizzle
        const KznImageDesc* lin = src->tiling == KZN_TILING_LINEAR ? src : dst;
        const KznImageDesc* til = src->tiling == KZN_TILING_LINEAR ? dst : src;
        const uint64_t slices = (uint64_t)src->depth * src->array_layers;
        bool all = true;
        for (uint64_t sl = 0; sl < slices && all; ++sl) {
            KznResult tr;
            all = swizzle_tma(queue, lin, lin->address + sl * kzn_slice_pitch(lin), til->address + sl * kzn_slice_pitch(til), dst->tiling == KZN_TILING_TILED, &tr);
            if (all && tr) return tr;
            if (!all && sl != 0) return KZN_ERROR_DEVICE_LOST;  // (applicability does not depend on the slice)
        }
        if (all) return KZN_SUCCESS;
    }
    Box sb = image_box(src, 0, 0, 0, 0), db = image_box(dst, 0, 0, 0, 0);
    return launch_box(queue, &sb, db, src->width, src->height, src->depth * src->array_layers, false, 0);
}

}  // extern "C"
