/* oracle/raster.c — CPU restatement of the draw path.  TEST INFRASTRUCTURE (see kzn_oracle.h).
 *
 * PARITY UNPINNED: the reference has no draw path at all (vkCmdDraw / vkCmdDrawIndexed are
 * `unimplemented!()`, vulkan-driver/src/api_impl.rs:4632-4652; GraphicsPipeline::create panics,
 * vulkan-driver/src/pipeline.rs:167-183).  What the reference pins and this file follows:
 *   subPixelPrecisionBits = 4              vulkan-driver/src/api_impl.rs:1559
 *   standardSampleLocations, 1 sample      vulkan-driver/src/api_impl.rs:1609
 *   depthClamp / depthBounds / independentBlend / dualSrcBlend = FALSE   api_impl.rs:875-930
 *   maxViewports = 1                       vulkan-driver/src/api_impl.rs:1566
 * Everything else is Vulkan 1.1.130 ("Fixed-Function Vertex Post-Processing", "Rasterization",
 * "Fragment Operations", "The Framebuffer") with the implementation-defined choices of SURVEY.md
 * Appendix E.5-E.9 fixed as written below.  These choices are normative for the CUDA kernels:
 *
 *  R1  clip->window: iw = 1/w; xd = x*iw; xf = fma(W/2, xd, X + W/2) (same for y, z with
 *      fma(maxDepth-minDepth, zd, minDepth)).
 *  R2  snapping: Xfix = (int32) rint(xf * 16) (RNE), sample of pixel p = 16p + 8.
 *  R3  clipping.  In x/y the view volume is cut by scissoring to the viewport rectangle, in z by a per-fragment
 *      test against [minDepth,maxDepth]; neither changes a triangle.  A triangle whose vertices all have
 *      w > 0 and finite window coordinates with |xf|,|yf| <= 2^19 (guard band) is rasterised as it is.  Any
 *      other triangle with finite clip coordinates is CLIPPED (Sutherland-Hodgman, homogeneous
 *      coordinates, colours interpolated with the positions) against five planes, in this order:
 *          w >= 2^-20,   G*w - wx >= 0,   G*w + wx >= 0,   G*w - wy >= 0,   G*w + wy >= 0
 *      with G = 2^18 pixels, wx = fma(W/2, x, (X + W/2)*w), wy likewise (window coordinate times w).  A
 *      crossing edge is cut from its inside end I to its outside end O at t = dI / (dI - dO), every
 *      component c = fma(t, cO - cI, cI) (both triangles sharing an edge compute the same point); on the
 *      w plane the new vertex gets w = 2^-20 exactly.  The polygon (<= 8 vertices) is drawn as the fan
 *      (P0, Pj+1, Pj+2), j = 0.. (<= 6 triangles, at the primitive's place in primitive order), each set
 *      up by R1-R5 like any triangle.  A triangle with a NaN / infinite clip coordinate is discarded.
 *      One exception keeps bands of a split frame cheap: when all three w > 0 and all three yf pass the
 *      tests and the snapped rows miss the clip rectangle, the triangle produces nothing (it is not
 *      looked at further, whatever its x).
 *  R4  2A = (X1-X0)(Y2-Y0) - (X2-X0)(Y1-Y0) in int64; 2A == 0 -> no fragments; with
 *      FRONT_FACE_COUNTER_CLOCKWISE a triangle is front facing iff 2A < 0 (Vulkan's area a = -A).
 *  R5  edge i is opposite vertex i, E_i(S) = a_i*Sx + b_i*Sy + c_i, sign-normalised so the interior
 *      is E_i > 0; top-left rule: E_i == 0 is inside iff a_i > 0 (left edge) or a_i == 0 && b_i > 0
 *      (top edge); implemented as E_i + bias_i > 0.
 *  R6  barycentric l_i = (float)E_i * (1.0f / (float)|2A|); z = fma(l2, z2, fma(l1, z1, l0*z0));
 *      attribute f = fma(l2, f2*iw2, fma(l1, f1*iw1, l0*(f0*iw0))) * (1.0f / fma(l2, iw2, fma(l1, iw1, l0*iw0)))
 *      (one correctly rounded reciprocal per fragment shared by the four channels, then one multiply each: within
 *      1.5 ulp of the exact quotient, inside the 2-ulp bar of the shaded output).
 *  R7  fragment order per pixel = primitive order; depth test before shading (no shader writes depth).
 *  R8  blending: source clamped to [0,1]; Cd = u8/255; C = Cs*Sf (op) Cd*Df with separate roundings; E.1.
 */
#include "kzn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GUARD_BAND 524288.0f /* 2^19 pixels */
#define CLIP_GUARD 262144.0f /* 2^18 pixels: where clipped polygons end (R3) */
#define CLIP_W 9.5367431640625e-07f /* 2^-20 */
#define MAX_SUB 6 /* triangles a clipped triangle can become: 3 + 5 planes = 8 vertices */

typedef struct {
    int32_t minx, maxx, miny, maxy; /* inclusive pixel bbox, already scissored; empty when minx > maxx */
    int64_t a[3], b[3], c[3];
    int64_t bias[3];
    float inv2a;
    float z[3], iw[3];
    float col[3][4]; /* colour * iw */
} TriSetup;

typedef struct {
    float x, y, z, w;
    float rgba[4];
} VsOut;

typedef struct {
    int32_t x0, y0, x1, y1; /* inclusive clip rectangle in pixels */
    float zmin, zmax;
} ClipRect;

/* ---- vertex fetch + passthrough vertex shader ------------------------------------------------- */

static void fetch_vertex(const KznVertexInput* vi, uint64_t index, VsOut* o) {
    const uint8_t* base = (const uint8_t*)(uintptr_t)vi->buffer_address;
    uint64_t off = index * vi->stride;
    uint32_t pos_bytes = vi->position_components * 4;
    o->x = o->y = o->z = 0.0f;
    o->w = 1.0f;
    /* out-of-range fetches read (0,0,0,1) like robustBufferAccess */
    if (off + vi->position_offset + pos_bytes <= vi->buffer_size) {
        float p[4] = {0, 0, 0, 1};
        memcpy(p, base + off + vi->position_offset, pos_bytes);
        o->x = p[0]; o->y = p[1]; o->z = p[2]; o->w = p[3];
    }
    o->rgba[0] = o->rgba[1] = o->rgba[2] = o->rgba[3] = 1.0f;
    if (vi->color_offset != ~0u) {
        if (off + vi->color_offset + 16 <= vi->buffer_size) memcpy(o->rgba, base + off + vi->color_offset, 16);
        else { o->rgba[0] = o->rgba[1] = o->rgba[2] = 0.0f; o->rgba[3] = 1.0f; }
    }
}

/* ---- triangle setup (R1-R5) --------------------------------------------------------------------- */

static int is_finite(float f) { return (f - f) == 0.0f; }

/* returns 1 when the triangle can produce fragments */
static int setup_triangle(const VsOut v[3], const KznGraphicsPipelineState* st, const KznViewport* vp,
                          const ClipRect* clip, TriSetup* t) {
    int32_t X[3], Y[3];
    float hw = vp->width * 0.5f, hh = vp->height * 0.5f;
    float ox = vp->x + hw, oy = vp->y + hh;
    float dz = vp->maxDepth - vp->minDepth;
    for (int i = 0; i < 3; ++i) {
        if (!(v[i].w > 0.0f)) return 0; /* R3 (also rejects NaN w) */
        float iw = 1.0f / v[i].w;
        float xf = fmaf(hw, v[i].x * iw, ox);
        float yf = fmaf(hh, v[i].y * iw, oy);
        float zf = fmaf(dz, v[i].z * iw, vp->minDepth);
        if (!is_finite(xf) || !is_finite(yf) || !is_finite(zf) || !is_finite(iw)) return 0;
        if (fabsf(xf) > GUARD_BAND || fabsf(yf) > GUARD_BAND) return 0;
        X[i] = (int32_t)nearbyintf(xf * 16.0f); /* R2 */
        Y[i] = (int32_t)nearbyintf(yf * 16.0f);
        t->z[i] = zf;
        t->iw[i] = iw;
        for (int ch = 0; ch < 4; ++ch) t->col[i][ch] = v[i].rgba[ch] * iw;
    }
    int64_t two_a = (int64_t)(X[1] - X[0]) * (Y[2] - Y[0]) - (int64_t)(X[2] - X[0]) * (Y[1] - Y[0]); /* R4 */
    if (two_a == 0) return 0;
    int front = st->front_face == 0 ? (two_a < 0) : (two_a > 0);
    if ((front && (st->cull_mode & 1u)) || (!front && (st->cull_mode & 2u))) return 0;
    int64_t s = two_a < 0 ? -1 : 1;
    for (int i = 0; i < 3; ++i) { /* R5 */
        int ia = (i + 1) % 3, ib = (i + 2) % 3;
        int64_t dx = X[ib] - X[ia], dy = Y[ib] - Y[ia];
        t->a[i] = -dy * s;
        t->b[i] = dx * s;
        t->c[i] = s * (dy * X[ia] - dx * Y[ia]);
        t->bias[i] = (t->a[i] > 0 || (t->a[i] == 0 && t->b[i] > 0)) ? 1 : 0;
    }
    t->inv2a = 1.0f / (float)(two_a * s);
    int32_t minX = X[0], maxX = X[0], minY = Y[0], maxY = Y[0];
    for (int i = 1; i < 3; ++i) {
        if (X[i] < minX) minX = X[i];
        if (X[i] > maxX) maxX = X[i];
        if (Y[i] < minY) minY = Y[i];
        if (Y[i] > maxY) maxY = Y[i];
    }
    /* pixels whose sample 16p+8 lies in [min,max]:  p >= ceil((min-8)/16), p <= floor((max-8)/16);
     * the arithmetic shift is a floor division */
    t->minx = (minX - 8 + 15) >> 4;
    t->maxx = (maxX - 8) >> 4;
    t->miny = (minY - 8 + 15) >> 4;
    t->maxy = (maxY - 8) >> 4;
    if (t->minx < clip->x0) t->minx = clip->x0;
    if (t->maxx > clip->x1) t->maxx = clip->x1;
    if (t->miny < clip->y0) t->miny = clip->y0;
    if (t->maxy > clip->y1) t->maxy = clip->y1;
    return t->minx <= t->maxx && t->miny <= t->maxy;
}

/* ---- clipping (R3) ---------------------------------------------------------------------------------------- */

typedef struct {
    float hw, hh, ox, oy;
} WindowScale;

static WindowScale window_scale(const KznViewport* vp) {
    WindowScale s;
    s.hw = vp->width * 0.5f;
    s.hh = vp->height * 0.5f;
    s.ox = vp->x + s.hw;
    s.oy = vp->y + s.hh;
    return s;
}

/* 0: rasterise as it is; 1: clip (or discard); 2: produces nothing (rows) */
static int classify_triangle(const VsOut v[3], const KznViewport* vp, const ClipRect* clip) {
    WindowScale s = window_scale(vp);
    float dz = vp->maxDepth - vp->minDepth;
    int y_ok = 1, x_ok = 1;
    int32_t Y[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {
        float iw = 1.0f / v[i].w;
        float yf = fmaf(s.hh, v[i].y * iw, s.oy);
        if (!(v[i].w > 0.0f) || !is_finite(yf) || !is_finite(iw) || fabsf(yf) > GUARD_BAND) y_ok = 0;
        float xf = fmaf(s.hw, v[i].x * iw, s.ox);
        float zf = fmaf(dz, v[i].z * iw, vp->minDepth);
        if (!is_finite(xf) || !is_finite(zf) || fabsf(xf) > GUARD_BAND) x_ok = 0;
        if (y_ok) Y[i] = (int32_t)nearbyintf(yf * 16.0f);
    }
    if (y_ok) {
        int32_t minY = Y[0], maxY = Y[0];
        for (int i = 1; i < 3; ++i) {
            if (Y[i] < minY) minY = Y[i];
            if (Y[i] > maxY) maxY = Y[i];
        }
        int32_t r0 = (minY - 8 + 15) >> 4, r1 = (maxY - 8) >> 4;
        if (r0 < clip->y0) r0 = clip->y0;
        if (r1 > clip->y1) r1 = clip->y1;
        if (r0 > r1) return 2;
    }
    return y_ok && x_ok ? 0 : 1;
}

static float clip_distance(const float c[8], int plane, const WindowScale* s) {
    if (plane == 0) return c[3] - CLIP_W;
    float gw = CLIP_GUARD * c[3];
    if (plane <= 2) {
        float wx = fmaf(s->hw, c[0], s->ox * c[3]);
        return plane == 1 ? gw - wx : gw + wx;
    }
    float wy = fmaf(s->hh, c[1], s->oy * c[3]);
    return plane == 3 ? gw - wy : gw + wy;
}

/* poly: in = the triangle (x y z w r g b a per vertex), out = the clipped polygon; returns its vertex count (0, 3..8),
 * or -1 when a clip coordinate is not finite */
static int clip_triangle(const VsOut v[3], const KznViewport* vp, float poly[8][8]) {
    WindowScale s = window_scale(vp);
    float a[8][8], b[8][8];
    for (int i = 0; i < 3; ++i) {
        if (!is_finite(v[i].x) || !is_finite(v[i].y) || !is_finite(v[i].z) || !is_finite(v[i].w)) return -1;
        a[i][0] = v[i].x; a[i][1] = v[i].y; a[i][2] = v[i].z; a[i][3] = v[i].w;
        for (int ch = 0; ch < 4; ++ch) a[i][4 + ch] = v[i].rgba[ch];
    }
    int n = 3;
    float (*src)[8] = a, (*dst)[8] = b;
    for (int plane = 0; plane < 5; ++plane) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const float* cur = src[i];
            const float* nxt = src[i + 1 == n ? 0 : i + 1];
            float dc = clip_distance(cur, plane, &s), dn = clip_distance(nxt, plane, &s);
            int in_c = dc >= 0.0f, in_n = dn >= 0.0f;
            if (in_c) memcpy(dst[m++], cur, sizeof(float) * 8);
            if (in_c != in_n) {
                const float* I = in_c ? cur : nxt;
                const float* O = in_c ? nxt : cur;
                float di = in_c ? dc : dn, dO = in_c ? dn : dc;
                float t = di / (di - dO);
                for (int k = 0; k < 8; ++k) dst[m][k] = fmaf(t, O[k] - I[k], I[k]);
                if (plane == 0) dst[m][3] = CLIP_W;
                ++m;
            }
        }
        float (*tmp)[8] = src; src = dst; dst = tmp;
        n = m;
        if (n < 3) return 0;
    }
    memcpy(poly, src, sizeof(float) * 8 * (size_t)n);
    return n;
}

/* ---- points and lines (R9) -------------------------------------------------------------------------------------
 * Vulkan 1.1 "Basic Point Rasterization" / "Basic Line Segment Rasterization" with the reference's limits: pointSizeRange
 * and lineWidthRange [1, 1], strictLines FALSE (api_impl.rs:1604-1608).  Both are rasterised as two triangles over four
 * corners in SNAPPED window coordinates (1/16 pixel), by the triangle rules R4-R8 (top-left rule: no pixel twice, none lost):
 *   point  (X, Y):          the square X-8 .. X+8, Y-8 .. Y+8 — exactly one pixel centre;
 *   line   (Xa,Ya)-(Xb,Yb): the non-strict parallelogram — x-major (|dX| >= |dY|): (Xa, Ya-8) (Xb, Yb-8) (Xb, Yb+8) (Xa, Ya+8),
 *                           y-major: (Xa-8, Ya) (Xb-8, Yb) (Xb+8, Yb) (Xa+8, Ya); the corners of an end carry that end's depth,
 *                           1/w and colour, so attributes run along the major axis, perspective-correct (R6); consecutive
 *                           segments of a strip share no pixel; a segment of zero length draws nothing.
 * Clipping: a point outside w >= 2^-20 or the guard band is discarded; a segment is cut to them first (same planes and
 * intersection arithmetic as R3, from the inside end).  No face culling.  A corner is handed on as the 8 floats
 * {X, Y, z_window, 1/w, r, g, b, a} ("window form": snapped coordinates up to 2^23 are exact in a float). */
static int setup_window_triangle(const float c0[8], const float c1[8], const float c2[8], const ClipRect* clip, TriSetup* t) {
    const float* c[3] = {c0, c1, c2};
    int32_t X[3], Y[3];
    for (int i = 0; i < 3; ++i) {
        X[i] = (int32_t)c[i][0];
        Y[i] = (int32_t)c[i][1];
        t->z[i] = c[i][2];
        t->iw[i] = c[i][3];
        for (int ch = 0; ch < 4; ++ch) t->col[i][ch] = c[i][4 + ch] * c[i][3];
    }
    int64_t two_a = (int64_t)(X[1] - X[0]) * (Y[2] - Y[0]) - (int64_t)(X[2] - X[0]) * (Y[1] - Y[0]);
    if (two_a == 0) return 0;
    int64_t s = two_a < 0 ? -1 : 1;
    for (int i = 0; i < 3; ++i) {
        int ia = (i + 1) % 3, ib = (i + 2) % 3;
        int64_t dx = X[ib] - X[ia], dy = Y[ib] - Y[ia];
        t->a[i] = -dy * s;
        t->b[i] = dx * s;
        t->c[i] = s * (dy * X[ia] - dx * Y[ia]);
        t->bias[i] = (t->a[i] > 0 || (t->a[i] == 0 && t->b[i] > 0)) ? 1 : 0;
    }
    t->inv2a = 1.0f / (float)(two_a * s);
    int32_t minX = X[0], maxX = X[0], minY = Y[0], maxY = Y[0];
    for (int i = 1; i < 3; ++i) {
        if (X[i] < minX) minX = X[i];
        if (X[i] > maxX) maxX = X[i];
        if (Y[i] < minY) minY = Y[i];
        if (Y[i] > maxY) maxY = Y[i];
    }
    t->minx = (minX - 8 + 15) >> 4;
    t->maxx = (maxX - 8) >> 4;
    t->miny = (minY - 8 + 15) >> 4;
    t->maxy = (maxY - 8) >> 4;
    if (t->minx < clip->x0) t->minx = clip->x0;
    if (t->maxx > clip->x1) t->maxx = clip->x1;
    if (t->miny < clip->y0) t->miny = clip->y0;
    if (t->maxy > clip->y1) t->maxy = clip->y1;
    return t->minx <= t->maxx && t->miny <= t->maxy;
}

/* clip-space vertex (x y z w r g b a) -> window form; 0: not representable (behind w, outside the guard band, not finite) */
static int to_window(const float v[8], const KznViewport* vp, float out[8]) {
    WindowScale s = window_scale(vp);
    float dz = vp->maxDepth - vp->minDepth;
    if (!(v[3] > 0.0f)) return 0;
    float iw = 1.0f / v[3];
    float xf = fmaf(s.hw, v[0] * iw, s.ox);
    float yf = fmaf(s.hh, v[1] * iw, s.oy);
    float zf = fmaf(dz, v[2] * iw, vp->minDepth);
    if (!is_finite(xf) || !is_finite(yf) || !is_finite(zf) || !is_finite(iw) || fabsf(xf) > GUARD_BAND || fabsf(yf) > GUARD_BAND) return 0;
    out[0] = (float)(int32_t)nearbyintf(xf * 16.0f);
    out[1] = (float)(int32_t)nearbyintf(yf * 16.0f);
    out[2] = zf;
    out[3] = iw;
    memcpy(out + 4, v + 4, 16);
    return 1;
}

/* the four corners (window form) of a point (n = 1) or a line segment (n = 2); returns 4, 0 (nothing to draw) or -1 (a clip
 * coordinate is not finite: discarded) */
static int expand_point_or_line(const VsOut* v, int n, const KznViewport* vp, float quad[4][8]) {
    float e[2][8];
    for (int i = 0; i < n; ++i) {
        if (!is_finite(v[i].x) || !is_finite(v[i].y) || !is_finite(v[i].z) || !is_finite(v[i].w)) return -1;
        e[i][0] = v[i].x; e[i][1] = v[i].y; e[i][2] = v[i].z; e[i][3] = v[i].w;
        memcpy(e[i] + 4, v[i].rgba, 16);
    }
    WindowScale s = window_scale(vp);
    float w0[8], w1[8];
    if (n == 1) {
        for (int plane = 0; plane < 5; ++plane)
            if (!(clip_distance(e[0], plane, &s) >= 0.0f)) return 0;
        if (!to_window(e[0], vp, w0)) return 0;
        const float dx[4] = {-8.0f, 8.0f, 8.0f, -8.0f}, dy[4] = {-8.0f, -8.0f, 8.0f, 8.0f};
        for (int k = 0; k < 4; ++k) {
            memcpy(quad[k], w0, 32);
            quad[k][0] = w0[0] + dx[k];
            quad[k][1] = w0[1] + dy[k];
        }
        return 4;
    }
    for (int plane = 0; plane < 5; ++plane) { /* the segment is cut plane by plane, from the inside end (R3's arithmetic) */
        float da = clip_distance(e[0], plane, &s), db = clip_distance(e[1], plane, &s);
        int in_a = da >= 0.0f, in_b = db >= 0.0f;
        if (!in_a && !in_b) return 0;
        if (in_a != in_b) {
            const float* I = in_a ? e[0] : e[1];
            const float* O = in_a ? e[1] : e[0];
            float di = in_a ? da : db, dO = in_a ? db : da;
            float t = di / (di - dO);
            float cut[8];
            for (int k = 0; k < 8; ++k) cut[k] = fmaf(t, O[k] - I[k], I[k]);
            if (plane == 0) cut[3] = CLIP_W;
            memcpy(in_a ? e[1] : e[0], cut, 32);
        }
    }
    if (!to_window(e[0], vp, w0) || !to_window(e[1], vp, w1)) return 0;
    float adx = fabsf(w1[0] - w0[0]), ady = fabsf(w1[1] - w0[1]);
    int x_major = adx >= ady;
    memcpy(quad[0], w0, 32); memcpy(quad[1], w1, 32); memcpy(quad[2], w1, 32); memcpy(quad[3], w0, 32);
    if (x_major) { quad[0][1] -= 8.0f; quad[1][1] -= 8.0f; quad[2][1] += 8.0f; quad[3][1] += 8.0f; }
    else { quad[0][0] -= 8.0f; quad[1][0] -= 8.0f; quad[2][0] += 8.0f; quad[3][0] += 8.0f; }
    return 4;
}

/* ---- fragment programs ------------------------------------------------------------------------------ */

void kzo_fs_procedural_orbit(uint32_t px, uint32_t py, float fb_width, float fb_height, float rgba[4]) {
    /* BASELINE.json configs[2]: fixed 32-iteration z <- z^2 + c orbit, fp32, explicit fma, no early exit */
    float inv_w = 1.0f / fb_width, inv_h = 1.0f / fb_height;
    float u = ((float)px + 0.5f) * inv_w;
    float v = ((float)py + 0.5f) * inv_h;
    float cx = fmaf(u, 3.5f, -2.5f);
    float cy = fmaf(v, 2.0f, -1.0f);
    float zx = 0.0f, zy = 0.0f;
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

/* ---- per-fragment operations -------------------------------------------------------------------------- */

static int depth_compare(uint32_t op, float z, float zs) {
    switch (op) {
    case 0: return 0;
    case 1: return z < zs;
    case 2: return z == zs;
    case 3: return z <= zs;
    case 4: return z > zs;
    case 5: return z != zs;
    case 6: return z >= zs;
    default: return 1;
    }
}

static float clamp01(float f) {
    if (!(f == f)) return 0.0f;
    return f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
}

static float blend_factor(uint32_t factor, int ch, const float s[4], const float d[4]) {
    switch (factor) {
    case 0: return 0.0f;
    case 1: return 1.0f;
    case 2: return s[ch];
    case 3: return 1.0f - s[ch];
    case 4: return d[ch];
    case 5: return 1.0f - d[ch];
    case 6: return s[3];
    case 7: return 1.0f - s[3];
    case 8: return d[3];
    case 9: return 1.0f - d[3];
    default: return 0.0f;
    }
}

static float blend_op(uint32_t op, float cs, float sf, float cd, float df) {
    float t0 = cs * sf, t1 = cd * df;
    switch (op) {
    case 0: return t0 + t1;
    case 1: return t0 - t1;
    case 2: return t1 - t0;
    case 3: return cs < cd ? cs : cd;
    default: return cs > cd ? cs : cd;
    }
}

/* VkLogicOp 0..15 on the UNORM8 bit patterns (all four channels at once: the operations are bitwise) */
static uint32_t logic_op(uint32_t op, uint32_t s, uint32_t d) {
    switch (op) {
    case 0: return 0u;            /* CLEAR */
    case 1: return s & d;         /* AND */
    case 2: return s & ~d;        /* AND_REVERSE */
    case 3: return s;             /* COPY */
    case 4: return ~s & d;        /* AND_INVERTED */
    case 5: return d;             /* NO_OP */
    case 6: return s ^ d;         /* XOR */
    case 7: return s | d;         /* OR */
    case 8: return ~(s | d);      /* NOR */
    case 9: return ~(s ^ d);      /* EQUIVALENT */
    case 10: return ~d;           /* INVERT */
    case 11: return s | ~d;       /* OR_REVERSE */
    case 12: return ~s;           /* COPY_INVERTED */
    case 13: return ~s | d;       /* OR_INVERTED */
    case 14: return ~(s & d);     /* NAND */
    default: return 0xffffffffu;  /* SET */
    }
}

static uint32_t output_merge(const KznGraphicsPipelineState* st, const float src_in[4], uint32_t dst_word) {
    float s[4], d[4], o[4];
    if (st->logic_op) { /* logicOpEnable: blending is treated as disabled; the write mask still applies */
        uint32_t sw = 0, mask = 0;
        for (int ch = 0; ch < 4; ++ch) {
            sw |= (uint32_t)kzo_float_to_unorm8(clamp01(src_in[ch])) << (8 * ch);
            if (st->color_write_mask & (1u << ch)) mask |= 0xffu << (8 * ch);
        }
        return (dst_word & ~mask) | (logic_op(st->logic_op - 1u, sw, dst_word) & mask);
    }
    for (int ch = 0; ch < 4; ++ch) {
        s[ch] = clamp01(src_in[ch]);
        d[ch] = kzo_unorm8_to_float((uint8_t)(dst_word >> (8 * ch)));
    }
    if (st->blend_enable) {
        for (int ch = 0; ch < 3; ++ch)
            o[ch] = blend_op(st->color_blend_op, s[ch], blend_factor(st->src_color_blend_factor, ch, s, d), d[ch],
                             blend_factor(st->dst_color_blend_factor, ch, s, d));
        o[3] = blend_op(st->alpha_blend_op, s[3], blend_factor(st->src_alpha_blend_factor, 3, s, d), d[3],
                        blend_factor(st->dst_alpha_blend_factor, 3, s, d));
    } else {
        memcpy(o, s, sizeof o);
    }
    uint32_t out = dst_word;
    for (int ch = 0; ch < 4; ++ch)
        if (st->color_write_mask & (1u << ch))
            out = (out & ~(0xffu << (8 * ch))) | ((uint32_t)kzo_float_to_unorm8(o[ch]) << (8 * ch));
    return out;
}

/* ---- validation shared by draw / draw_indexed ----------------------------------------------------------- */

static KznResult check_draw(const KznGraphicsPipelineState* st, const KznFramebuffer* fb, const KznViewport* vp,
                            const KznRect2D* sc, const KznVertexInput* vi, uint32_t* W, uint32_t* H) {
    if (!st || !fb || !vp || !sc || !vi) return KZN_ERROR_VALIDATION_FAILED;
    if (st->vertex_program != KZN_VS_PASSTHROUGH) return KZN_ERROR_FEATURE_NOT_PRESENT;
    if (st->fragment_program > KZN_FS_PROCEDURAL_ORBIT) return KZN_ERROR_FEATURE_NOT_PRESENT;
    if (st->cull_mode > 3 || st->front_face > 1 || st->depth_compare_op > 7) return KZN_ERROR_VALIDATION_FAILED;
    if (st->src_color_blend_factor > 9 || st->dst_color_blend_factor > 9 || st->src_alpha_blend_factor > 9 ||
        st->dst_alpha_blend_factor > 9)
        return KZN_ERROR_FEATURE_NOT_PRESENT;
    if (st->color_blend_op > 4 || st->alpha_blend_op > 4) return KZN_ERROR_VALIDATION_FAILED;
    if (st->logic_op > 16) return KZN_ERROR_VALIDATION_FAILED;
    if (st->topology > 10 && st->topology != KZN_TOPOLOGY_POINT_LIST) return KZN_ERROR_VALIDATION_FAILED;
    if (st->topology > KZN_TOPOLOGY_TRIANGLE_FAN && st->topology != KZN_TOPOLOGY_POINT_LIST)
        return KZN_ERROR_FEATURE_NOT_PRESENT; /* adjacency, patches */
    int has_color = fb->color.format != 0, has_depth = fb->depth.format != 0;
    if (!has_color && !has_depth) return KZN_ERROR_VALIDATION_FAILED;
    if (has_color) {
        if (fb->color.format != KZN_FORMAT_R8G8B8A8_UNORM) return KZN_ERROR_FORMAT_NOT_SUPPORTED;
        if (fb->color.mip_levels != 1 || fb->color.depth != 1 || fb->color.tiling > 1) return KZN_ERROR_VALIDATION_FAILED;
        if (fb->color_layer >= fb->color.array_layers) return KZN_ERROR_VALIDATION_FAILED;
        *W = fb->color.width;
        *H = fb->color.height;
    }
    if (has_depth) {
        if (fb->depth.format != KZN_FORMAT_D32_SFLOAT) return KZN_ERROR_FORMAT_NOT_SUPPORTED;
        if (fb->depth.mip_levels != 1 || fb->depth.depth != 1 || fb->depth.tiling != 0) return KZN_ERROR_VALIDATION_FAILED;
        if (fb->depth_layer >= fb->depth.array_layers) return KZN_ERROR_VALIDATION_FAILED;
        if (has_color && (fb->depth.width != fb->color.width || fb->depth.height != fb->color.height))
            return KZN_ERROR_VALIDATION_FAILED;
        *W = fb->depth.width;
        *H = fb->depth.height;
    }
    if (fb->band_y1 && (fb->band_y0 >= fb->band_y1 || fb->band_y1 > *H)) return KZN_ERROR_VALIDATION_FAILED;
    if (sc->offset.x < 0 || sc->offset.y < 0) return KZN_ERROR_VALIDATION_FAILED; /* VUID 00598/00599 */
    if (!(vp->width > 0.0f) || vp->height == 0.0f || !is_finite(vp->height)) return KZN_ERROR_VALIDATION_FAILED;
    /* viewportBoundsRange: the clip rectangle converts these to integers (undefined for NaN / infinity / huge values) */
    if (!is_finite(vp->x) || !is_finite(vp->y) || !is_finite(vp->width) || fabsf(vp->x) > 1048576.0f ||
        fabsf(vp->y) > 1048576.0f || vp->width > 1048576.0f || fabsf(vp->height) > 1048576.0f)
        return KZN_ERROR_VALIDATION_FAILED;
    if (!(vp->minDepth >= 0.0f && vp->minDepth <= 1.0f && vp->maxDepth >= 0.0f && vp->maxDepth <= 1.0f))
        return KZN_ERROR_VALIDATION_FAILED;
    if (vi->position_components != 3 && vi->position_components != 4) return KZN_ERROR_VALIDATION_FAILED;
    if (vi->stride % 4 || vi->position_offset % 4 || (vi->color_offset != ~0u && vi->color_offset % 4))
        return KZN_ERROR_VALIDATION_FAILED;
    if (!vi->buffer_address || vi->buffer_address % 4) return KZN_ERROR_VALIDATION_FAILED;
    if (vi->stride >= (1u << 28)) return KZN_ERROR_VALIDATION_FAILED; /* index x stride stays far from 2^64 */
    return KZN_SUCCESS;
}

static void make_clip(const KznFramebuffer* fb, const KznViewport* vp, const KznRect2D* sc, uint32_t W, uint32_t H,
                      ClipRect* c) {
    /* view-volume clip in x/y == scissor to the viewport rectangle: pixels whose centre lies inside it */
    float vx0 = vp->x, vx1 = vp->x + vp->width;
    float vy0 = vp->height < 0.0f ? vp->y + vp->height : vp->y;
    float vy1 = vp->height < 0.0f ? vp->y : vp->y + vp->height;
    int64_t x0 = (int64_t)ceilf(vx0 - 0.5f), x1 = (int64_t)floorf(vx1 - 0.5f);
    int64_t y0 = (int64_t)ceilf(vy0 - 0.5f), y1 = (int64_t)floorf(vy1 - 0.5f);
    int64_t sx1 = (int64_t)sc->offset.x + sc->extent.width - 1, sy1 = (int64_t)sc->offset.y + sc->extent.height - 1;
    if (x0 < sc->offset.x) x0 = sc->offset.x;
    if (y0 < sc->offset.y) y0 = sc->offset.y;
    if (x1 > sx1) x1 = sx1;
    if (y1 > sy1) y1 = sy1;
    int64_t by0 = fb->band_y1 ? fb->band_y0 : 0, by1 = fb->band_y1 ? fb->band_y1 : H;
    if (x0 < 0) x0 = 0;
    if (y0 < by0) y0 = by0;
    if (x1 > (int64_t)W - 1) x1 = (int64_t)W - 1;
    if (y1 > by1 - 1) y1 = by1 - 1;
    if (x1 < x0 || y1 < y0) { x0 = y0 = 0; x1 = y1 = -1; }
    c->x0 = (int32_t)x0; c->y0 = (int32_t)y0; c->x1 = (int32_t)x1; c->y1 = (int32_t)y1;
    c->zmin = vp->minDepth < vp->maxDepth ? vp->minDepth : vp->maxDepth;
    c->zmax = vp->minDepth < vp->maxDepth ? vp->maxDepth : vp->minDepth;
}

/* ---- the rasteriser ------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t covered, depth_passed;
} FragCounts;

static void raster_rows(const KznGraphicsPipelineState* st, const KznFramebuffer* fb, const ClipRect* clip,
                        const TriSetup* tris, const uint8_t* valid, uint64_t ntris, int32_t row0, int32_t row1,
                        uint32_t W, uint32_t H, FragCounts* counts) {
    int has_color = fb->color.format != 0, has_depth = fb->depth.format != 0;
    int depth_test = has_depth && st->depth_test_enable;
    int depth_write = depth_test && st->depth_write_enable;
    for (uint64_t ti = 0; ti < ntris; ++ti) {
        if (!valid[ti]) continue;
        const TriSetup* t = &tris[ti];
        int32_t y0 = t->miny > row0 ? t->miny : row0, y1 = t->maxy < row1 ? t->maxy : row1;
        for (int32_t py = y0; py <= y1; ++py) {
            int64_t sy = 16 * (int64_t)py + 8;
            for (int32_t px = t->minx; px <= t->maxx; ++px) {
                int64_t sx = 16 * (int64_t)px + 8;
                int64_t e0 = t->a[0] * sx + t->b[0] * sy + t->c[0];
                int64_t e1 = t->a[1] * sx + t->b[1] * sy + t->c[1];
                int64_t e2 = t->a[2] * sx + t->b[2] * sy + t->c[2];
                if (e0 + t->bias[0] <= 0 || e1 + t->bias[1] <= 0 || e2 + t->bias[2] <= 0) continue;
                float l0 = (float)e0 * t->inv2a, l1 = (float)e1 * t->inv2a, l2 = (float)e2 * t->inv2a; /* R6 */
                float z = fmaf(l2, t->z[2], fmaf(l1, t->z[1], l0 * t->z[0]));
                if (!(z >= clip->zmin && z <= clip->zmax)) continue; /* R3: z clip */
                counts->covered++;
                float* dp = NULL;
                if (has_depth)
                    dp = (float*)((uint8_t*)(uintptr_t)fb->depth.address +
                                  kzo_texel_offset(&fb->depth, (uint32_t)px, (uint32_t)py, 0, fb->depth_layer));
                if (depth_test && !depth_compare(st->depth_compare_op, z, *dp)) continue; /* R7 */
                counts->depth_passed++;
                if (depth_write) *dp = z;
                if (!has_color) continue;
                float src[4];
                if (st->fragment_program == KZN_FS_PROCEDURAL_ORBIT) {
                    kzo_fs_procedural_orbit((uint32_t)px, (uint32_t)py, (float)W, (float)H, src);
                } else {
                    float den = fmaf(l2, t->iw[2], fmaf(l1, t->iw[1], l0 * t->iw[0]));
                    float rden = 1.0f / den; /* R6: one reciprocal per fragment */
                    for (int ch = 0; ch < 4; ++ch) {
                        float num = fmaf(l2, t->col[2][ch], fmaf(l1, t->col[1][ch], l0 * t->col[0][ch]));
                        src[ch] = num * rden;
                    }
                }
                uint32_t* cp = (uint32_t*)((uint8_t*)(uintptr_t)fb->color.address +
                                           kzo_texel_offset(&fb->color, (uint32_t)px, (uint32_t)py, 0, fb->color_layer));
                *cp = output_merge(st, src, *cp);
            }
        }
    }
}

static KznResult draw_common(const KznGraphicsPipelineState* st, const KznFramebuffer* fb, const KznViewport* vp,
                             const KznRect2D* sc, const KznVertexInput* vi, const void* indices, uint32_t index_type,
                             uint64_t first, uint64_t count, int64_t vertex_offset, uint32_t instance_count,
                             KznDrawStatistics* stats) {
    uint32_t W = 0, H = 0;
    KznResult r = check_draw(st, fb, vp, sc, vi, &W, &H);
    if (r) return r;
    ClipRect clip;
    make_clip(fb, vp, sc, W, H, &clip);
    const uint32_t topo = st->topology ? st->topology : KZN_TOPOLOGY_TRIANGLE_LIST;
    const int prim_verts = topo == KZN_TOPOLOGY_POINT_LIST ? 1 : (topo == KZN_TOPOLOGY_LINE_LIST || topo == KZN_TOPOLOGY_LINE_STRIP) ? 2 : 3;
    uint64_t ntri = topo == KZN_TOPOLOGY_POINT_LIST ? count : topo == KZN_TOPOLOGY_LINE_LIST ? count / 2 :
                    topo == KZN_TOPOLOGY_LINE_STRIP ? (count >= 2 ? count - 1 : 0) :
                    topo == KZN_TOPOLOGY_TRIANGLE_LIST ? count / 3 : (count >= 3 ? count - 2 : 0); /* primitives */
    uint64_t total = ntri * instance_count;
    if (total == 0) return KZN_SUCCESS;
    /* MAX_SUB record slots per primitive: slot 0 of an unclipped one, slots 0 .. n-1 of one clipped into n triangles */
    TriSetup* tris = (TriSetup*)malloc(total * MAX_SUB * sizeof(TriSetup));
    uint8_t* valid = (uint8_t*)calloc(total * MAX_SUB, 1);
    uint8_t* fate = (uint8_t*)calloc(ntri, 1); /* 1: clipped, 2: discarded (not finite) */
    uint8_t* nsub = (uint8_t*)calloc(ntri, 1);
    if (!tris || !valid || !fate || !nsub) { free(tris); free(valid); free(fate); free(nsub); return KZN_ERROR_OUT_OF_HOST_MEMORY; }
    int threads = kzo_get_threads();
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t t = 0; t < (int64_t)ntri; ++t) {
        VsOut v[3];
        for (int k = 0; k < prim_verts; ++k) {
            /* Vulkan 1.1 "Primitive Topologies": list {3t, 3t+1, 3t+2}; strip {t, t+(1+t%2), t+(2-t%2)}; fan {t+1, t+2, 0};
             * points {t}; line list {2t, 2t+1}; line strip {t, t+1} */
            uint64_t rel = (uint64_t)prim_verts * (uint64_t)t + k;
            if (topo == KZN_TOPOLOGY_LINE_STRIP) rel = (uint64_t)t + k;
            if (topo == KZN_TOPOLOGY_TRIANGLE_STRIP) rel = (uint64_t)t + (k == 0 ? 0 : k == 1 ? 1 + (uint64_t)t % 2 : 2 - (uint64_t)t % 2);
            if (topo == KZN_TOPOLOGY_TRIANGLE_FAN) rel = k == 2 ? 0 : (uint64_t)t + 1 + k;
            uint64_t i = first + rel;
            int64_t vidx;
            if (indices) {
                uint64_t raw = index_type == KZN_INDEX_TYPE_UINT16 ? ((const uint16_t*)indices)[i] : ((const uint32_t*)indices)[i];
                vidx = (int64_t)raw + vertex_offset;
            } else {
                vidx = (int64_t)i;
            }
            if (vidx < 0) vidx = (int64_t)1 << 40; /* out of range -> robust fetch */
            fetch_vertex(vi, (uint64_t)vidx, &v[k]);
        }
        if (prim_verts < 3) { /* R9: two triangles over the four corners, in window form */
            float quad[4][8];
            int m = expand_point_or_line(v, prim_verts, vp, quad);
            if (m < 0) fate[t] = 2;
            if (m == 4) {
                valid[t * MAX_SUB] = (uint8_t)setup_window_triangle(quad[0], quad[1], quad[2], &clip, &tris[t * MAX_SUB]);
                valid[t * MAX_SUB + 1] = (uint8_t)setup_window_triangle(quad[0], quad[2], quad[3], &clip, &tris[t * MAX_SUB + 1]);
            }
            continue;
        }
        int cls = classify_triangle(v, vp, &clip);
        if (cls == 0) {
            valid[t * MAX_SUB] = (uint8_t)setup_triangle(v, st, vp, &clip, &tris[t * MAX_SUB]);
        } else if (cls == 1) { /* R3 */
            float poly[8][8];
            int m = clip_triangle(v, vp, poly);
            fate[t] = m < 0 ? 2 : 1;
            for (int j = 0; j + 2 < m; ++j) {
                VsOut s3[3];
                const int src[3] = {0, j + 1, j + 2};
                for (int k = 0; k < 3; ++k) {
                    const float* c = poly[src[k]];
                    s3[k].x = c[0]; s3[k].y = c[1]; s3[k].z = c[2]; s3[k].w = c[3];
                    memcpy(s3[k].rgba, c + 4, 16);
                }
                valid[t * MAX_SUB + j] = (uint8_t)setup_triangle(s3, st, vp, &clip, &tris[t * MAX_SUB + j]);
                nsub[t]++;
            }
        }
    }
    for (uint32_t inst = 1; inst < instance_count; ++inst) { /* no per-instance inputs: instances repeat */
        memcpy(tris + (uint64_t)inst * ntri * MAX_SUB, tris, ntri * MAX_SUB * sizeof(TriSetup));
        memcpy(valid + (uint64_t)inst * ntri * MAX_SUB, valid, ntri * MAX_SUB);
    }
    uint64_t culled = 0, clipped = 0, clip_out = 0, discarded = 0;
    for (uint64_t t = 0; t < ntri; ++t) {
        int any = 0;
        for (int j = 0; j < MAX_SUB; ++j) any |= valid[t * MAX_SUB + j];
        culled += (uint64_t)!any * instance_count;
        clipped += (uint64_t)(fate[t] == 1) * instance_count;
        discarded += (uint64_t)(fate[t] == 2) * instance_count;
        clip_out += (uint64_t)nsub[t] * instance_count;
    }
    /* bands of rows are independent (disjoint pixels); every band walks all triangles in order */
    int32_t band = 8;
    int32_t nb = clip.y1 >= clip.y0 ? (clip.y1 - clip.y0 + band) / band : 0;
    uint64_t covered = 0, passed = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : covered, passed)
    for (int32_t bnd = 0; bnd < nb; ++bnd) {
        FragCounts fc = {0, 0};
        int32_t r0 = clip.y0 + bnd * band, r1 = r0 + band - 1;
        if (r1 > clip.y1) r1 = clip.y1;
        raster_rows(st, fb, &clip, tris, valid, total * MAX_SUB, r0, r1, W, H, &fc);
        covered += fc.covered;
        passed += fc.depth_passed;
    }
    if (stats) {
        stats->triangles_in += total;
        stats->triangles_culled += culled;
        stats->fragments_covered += covered;
        stats->fragments_depth_passed += passed;
        stats->triangles_clipped += clipped;
        stats->clip_triangles += clip_out;
        stats->triangles_discarded += discarded;
    }
    free(tris);
    free(valid);
    free(fate);
    free(nsub);
    return KZN_SUCCESS;
}

KznResult kzo_draw(const KznGraphicsPipelineState* st, const KznFramebuffer* fb, const KznViewport* vp,
                   const KznRect2D* sc, const KznVertexInput* vi, uint32_t vertex_count, uint32_t instance_count,
                   uint32_t first_vertex, uint32_t first_instance, KznDrawStatistics* stats) {
    (void)first_instance;
    return draw_common(st, fb, vp, sc, vi, NULL, 0, first_vertex, vertex_count, 0, instance_count, stats);
}

KznResult kzo_draw_indexed(const KznGraphicsPipelineState* st, const KznFramebuffer* fb, const KznViewport* vp,
                           const KznRect2D* sc, const KznVertexInput* vi, const void* index_buffer,
                           uint64_t index_buffer_size, uint32_t index_type, uint32_t index_count,
                           uint32_t instance_count, uint32_t first_index, int32_t vertex_offset, uint32_t first_instance,
                           KznDrawStatistics* stats) {
    (void)first_instance;
    if (!index_buffer || index_type > KZN_INDEX_TYPE_UINT32) return KZN_ERROR_VALIDATION_FAILED;
    uint64_t isz = index_type == KZN_INDEX_TYPE_UINT16 ? 2 : 4;
    if (((uint64_t)first_index + index_count) * isz > index_buffer_size) return KZN_ERROR_VALIDATION_FAILED;
    return draw_common(st, fb, vp, sc, vi, index_buffer, index_type, first_index, index_count, vertex_offset,
                       instance_count, stats);
}

/* ---- synthetic mesh -------------------------------------------------------- */

static float unit_float(uint64_t* st) { return (float)(kzo_splitmix64(st) >> 40) * (1.0f / 16777216.0f); }

void kzo_make_grid_mesh(uint32_t quads_x, uint32_t quads_y, uint64_t seed, float* vertices, uint32_t* indices) {
    uint64_t st = seed;
    uint32_t vx = quads_x + 1, vy = quads_y + 1;
    float cell_x = 2.0f / (float)quads_x, cell_y = 2.0f / (float)quads_y;
    for (uint32_t j = 0; j < vy; ++j)
        for (uint32_t i = 0; i < vx; ++i) {
            float* v = vertices + 7 * ((uint64_t)j * vx + i);
            /* jitter of +-1.5 cells makes neighbouring quads overlap (depth complexity ~4) */
            float jx = (unit_float(&st) - 0.5f) * 3.0f, jy = (unit_float(&st) - 0.5f) * 3.0f;
            v[0] = -1.0f + ((float)i + jx) * cell_x;
            v[1] = -1.0f + ((float)j + jy) * cell_y;
            v[2] = 0.05f + 0.9f * unit_float(&st);
            v[3] = unit_float(&st);
            v[4] = unit_float(&st);
            v[5] = unit_float(&st);
            v[6] = 0.25f + 0.75f * unit_float(&st);
        }
    for (uint32_t j = 0; j < quads_y; ++j)
        for (uint32_t i = 0; i < quads_x; ++i) {
            uint32_t* q = indices + 6 * ((uint64_t)j * quads_x + i);
            uint32_t v00 = j * vx + i, v10 = v00 + 1, v01 = v00 + vx, v11 = v01 + 1;
            q[0] = v00; q[1] = v10; q[2] = v11;
            q[3] = v00; q[4] = v11; q[5] = v01;
        }
}
