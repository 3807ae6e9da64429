// ellipse_mask.h -- which 16x16-px tiles of a splat's tile rect can contain a covered pixel (host + device, no CUDA dependency).
//
// A splat covers pixel p iff q(p) = (g1.(p-c))^2 + (g2.(p-c))^2 <= 1 at the pixel CENTRE (the fragment shader's `A > 8 -> discard`,
// SplatMaterial3D.js:240-243, with A = 8 q).  q is a convex quadratic, so its minimum over a tile's rectangle of pixel centres is 0
// when the rectangle contains c and otherwise lies on one of the four edges, where it is a clamped 1-D minimisation: exact, no sampling.
// A tile is kept when that minimum is <= 1 + kEllipseSlack; the slack covers the rounding of this test and of the blend kernel's own
// evaluation of q (forward differences in f32), so a tile holding a covered pixel is never dropped.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD static inline
#endif

constexpr float kEllipseSlack = 2e-3f;
constexpr int kMaskTile = 16;          // fine tile edge in pixels (kTile)
constexpr int kMaskSpan = 8;           // the bitmap covers rects of up to 8 x 8 tiles

GS_HD float gs_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// minimum of q(x, y) = qxx x^2 + 2 qxy x y + qyy y^2 over the rectangle [x0, x1] x [y0, y1] (coordinates relative to the centre)
GS_HD float ellipse_min_q(float x0, float x1, float y0, float y1, float qxx, float qxy, float qyy) {
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return 0.f;
    float best = 3.0e38f;
    for (int e = 0; e < 2; ++e) {                       // edges x = x0, x = x1
        const float X = e ? x1 : x0;
        const float y = gs_clampf(-qxy * X / qyy, y0, y1);
        const float q = qxx * X * X + 2.f * qxy * X * y + qyy * y * y;
        best = q < best ? q : best;
    }
    for (int e = 0; e < 2; ++e) {                       // edges y = y0, y = y1
        const float Y = e ? y1 : y0;
        const float x = gs_clampf(-qxy * Y / qxx, x0, x1);
        const float q = qxx * x * x + 2.f * qxy * x * Y + qyy * Y * Y;
        best = q < best ? q : best;
    }
    return best;
}

// Bit (ty - ry0) * 8 + (tx - rx0) is set when tile (tx, ty) of the rect [rx0, rx1] x [ry0, ry1] (tile units, inclusive) can hold a
// covered pixel.  Rects wider or taller than 8 tiles get all ones (no pruning).  (cx, cy): splat centre in pixels; g1, g2: the
// rows of the pixel -> quad-local map stored in the splat record.
GS_HD unsigned long long ellipse_tile_bitmap(int rx0, int ry0, int rx1, int ry1, float cx, float cy, float g1x, float g1y, float g2x, float g2y) {
    if (rx1 - rx0 >= kMaskSpan || ry1 - ry0 >= kMaskSpan) return ~0ull;
    const float qxx = g1x * g1x + g2x * g2x, qxy = g1x * g1y + g2x * g2y, qyy = g1y * g1y + g2y * g2y;
    unsigned long long bits = 0;
    for (int ty = ry0; ty <= ry1; ++ty) {
        const float y0 = (float)(ty * kMaskTile) + 0.5f - cy, y1 = y0 + (float)(kMaskTile - 1);
        for (int tx = rx0; tx <= rx1; ++tx) {
            const float x0 = (float)(tx * kMaskTile) + 0.5f - cx, x1 = x0 + (float)(kMaskTile - 1);
            if (ellipse_min_q(x0, x1, y0, y1, qxx, qxy, qyy) <= 1.0f + kEllipseSlack) bits |= 1ull << ((ty - ry0) * kMaskSpan + (tx - rx0));
        }
    }
    return bits;
}

// The 8 x 4 fine-tile mask (bit = ly * 8 + lx) of coarse tile (ccx, ccy) -- 8 x 4 fine tiles -- taken from a splat's bitmap.
GS_HD uint32_t coarse_mask_from_bitmap(int rx0, int ry0, int ry1, unsigned long long bitmap, int ccx, int ccy) {
    uint32_t m = 0;
    const int shift = rx0 - ccx * 8;
    for (int ly = 0; ly < 4; ++ly) {
        const int fy = ccy * 4 + ly;
        if (fy < ry0 || fy > ry1 || fy - ry0 >= kMaskSpan) continue;
        const uint32_t row = (uint32_t)(bitmap >> ((fy - ry0) * kMaskSpan)) & 0xffu;       // bits for tx = rx0 .. rx0 + 7
        const uint32_t local = shift >= 0 ? (shift < 8 ? (row << shift) : 0u) : (-shift < 8 ? (row >> -shift) : 0u);
        m |= (local & 0xffu) << (8 * ly);
    }
    return m;
}
