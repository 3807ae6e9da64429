// oracle/ellipse_mask_check.cpp -- TEST INFRASTRUCTURE ONLY.  Brute-force check of gaussiansplats3d_b200/csrc/ellipse_mask.h on the
// host: for random splats, every 16x16 tile that contains a pixel centre with q <= 1 must have its bit set (never drop coverage),
// and the bits set must stay close to the exact set (the point of the test is pruning).  Also checks coarse_mask_from_bitmap
// against a per-tile lookup.  Exit code 0 = ok; prints the statistics.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../gaussiansplats3d_b200/csrc/ellipse_mask.h"

int main(int argc, char **argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 200000;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long long tiles_rect = 0, tiles_kept = 0, tiles_exact = 0, dropped = 0, coarse_bad = 0;
    for (int t = 0; t < trials; ++t) {
        const float cx = (float)(U(rng) * 1920.0), cy = (float)(U(rng) * 1080.0);
        const double l1 = std::exp(U(rng) * 4.5), l2 = std::exp(U(rng) * std::log(l1 + 1e-9));   // axis half-lengths in px, 1..90, l2 <= l1
        const double th = U(rng) * 6.283185307179586;
        const float ex = (float)std::cos(th), ey = (float)std::sin(th);
        const float s1 = (float)l1, s2 = (float)std::fmax(l2, 0.6);
        const float g1x = ex / s1, g1y = ey / s1, g2x = ey / s2, g2y = -ex / s2;
        const float b1x = ex * s1, b1y = ey * s1, b2x = ey * s2, b2y = -ex * s2;
        const float hx = std::sqrt(b1x * b1x + b2x * b2x) * 1.0005f + 0.01f, hy = std::sqrt(b1y * b1y + b2y * b2y) * 1.0005f + 0.01f;
        const float fx0 = std::ceil(cx - hx - 0.5f), fx1 = std::floor(cx + hx - 0.5f), fy0 = std::ceil(cy - hy - 0.5f), fy1 = std::floor(cy + hy - 0.5f);
        if (fx1 < 0 || fy1 < 0 || fx0 > 1919 || fy0 > 1079 || fx0 > fx1 || fy0 > fy1) continue;
        const int px0 = (int)std::fmax(fx0, 0.f), px1 = (int)std::fmin(fx1, 1919.f), py0 = (int)std::fmax(fy0, 0.f), py1 = (int)std::fmin(fy1, 1079.f);
        const int rx0 = px0 >> 4, rx1 = px1 >> 4, ry0 = py0 >> 4, ry1 = py1 >> 4;
        const unsigned long long bm = ellipse_tile_bitmap(rx0, ry0, rx1, ry1, cx, cy, g1x, g1y, g2x, g2y);
        const bool big = (rx1 - rx0 >= 8) || (ry1 - ry0 >= 8);
        for (int ty = ry0; ty <= ry1; ++ty)
            for (int tx = rx0; tx <= rx1; ++tx) {
                bool hit = false;     // brute force over the tile's pixel centres, in double
                for (int y = ty * 16; y < ty * 16 + 16 && !hit; ++y)
                    for (int x = tx * 16; x < tx * 16 + 16; ++x) {
                        const double dx = x + 0.5 - cx, dy = y + 0.5 - cy;
                        const double u = dx * g1x + dy * g1y, w = dx * g2x + dy * g2y;
                        if (u * u + w * w <= 1.0 + 1e-4) { hit = true; break; }
                    }
                ++tiles_rect;
                tiles_exact += hit;
                const bool kept = big ? true : ((bm >> ((ty - ry0) * 8 + (tx - rx0))) & 1ull) != 0;
                tiles_kept += kept;
                if (hit && !kept) ++dropped;
            }
        if (!big) {   // coarse extraction: every coarse tile (8 x 4 fine tiles) touching the rect
            for (int ccy = ry0 >> 2; ccy <= ry1 >> 2; ++ccy)
                for (int ccx = rx0 >> 3; ccx <= rx1 >> 3; ++ccx) {
                    const uint32_t m = coarse_mask_from_bitmap(rx0, ry0, ry1, bm, ccx, ccy);
                    for (int ly = 0; ly < 4; ++ly)
                        for (int lx = 0; lx < 8; ++lx) {
                            const int tx = ccx * 8 + lx, ty = ccy * 4 + ly;
                            const bool in_rect = tx >= rx0 && tx <= rx1 && ty >= ry0 && ty <= ry1;
                            const bool want = in_rect && ((bm >> ((ty - ry0) * 8 + (tx - rx0))) & 1ull);
                            if ((((m >> (ly * 8 + lx)) & 1u) != 0) != want) ++coarse_bad;
                        }
                }
        }
    }
    printf("{\"tiles_in_rects\": %lld, \"kept\": %lld, \"exact\": %lld, \"dropped_hits\": %lld, \"coarse_mismatch\": %lld}\n", tiles_rect, tiles_kept, tiles_exact, dropped,
           coarse_bad);
    return (dropped == 0 && coarse_bad == 0) ? 0 : 1;
}
