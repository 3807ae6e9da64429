// oracle/ellipse_mask_check.cpp -- TEST INFRASTRUCTURE ONLY.  Brute-force check of gaussiansplats3d_b200/csrc/ellipse_mask.h on the
// host, used the way the blend kernel uses it: for random splats and every 8x8-px block their AABB reaches, a block that contains a
// pixel centre with q <= 1 must pass ellipse_min_q(...) <= 1 + kEllipseSlack (coverage is never dropped), and the set of blocks kept
// must stay close to the exact set (the point of the test is pruning AABB corners).  Exit code 0 = ok; prints the statistics.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../gaussiansplats3d_b200/csrc/ellipse_mask.h"

int main(int argc, char **argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 200000;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long long blocks_aabb = 0, kept = 0, exact = 0, dropped = 0;
    for (int t = 0; t < trials; ++t) {
        const float cx = (float)(U(rng) * 1920.0), cy = (float)(U(rng) * 1080.0);
        const double l1 = std::exp(U(rng) * 4.5), l2 = std::exp(U(rng) * std::log(l1 + 1e-9));   // axis half-lengths in px, 1..90, l2 <= l1
        const double th = U(rng) * 6.283185307179586;
        const float ex = (float)std::cos(th), ey = (float)std::sin(th);
        const float s1 = (float)l1, s2 = (float)std::fmax(l2, 0.6);
        const float g1x = ex / s1, g1y = ey / s1, g2x = ey / s2, g2y = -ex / s2;
        const float qxx = g1x * g1x + g2x * g2x, qxy = g1x * g1y + g2x * g2y, qyy = g1y * g1y + g2y * g2y;
        const float b1x = ex * s1, b1y = ey * s1, b2x = ey * s2, b2y = -ex * s2;
        const float hx = std::sqrt(b1x * b1x + b2x * b2x) * 1.0005f + 0.01f, hy = std::sqrt(b1y * b1y + b2y * b2y) * 1.0005f + 0.01f;
        const int bx0 = (int)std::floor((cx - hx) / 8.0f) - 1, bx1 = (int)std::floor((cx + hx) / 8.0f) + 1;
        const int by0 = (int)std::floor((cy - hy) / 8.0f) - 1, by1 = (int)std::floor((cy + hy) / 8.0f) + 1;
        for (int by = by0; by <= by1; ++by)
            for (int bx = bx0; bx <= bx1; ++bx) {
                bool hit = false;     // brute force over the block's pixel centres, in double
                for (int y = by * 8; y < by * 8 + 8 && !hit; ++y)
                    for (int x = bx * 8; x < bx * 8 + 8; ++x) {
                        const double dx = x + 0.5 - cx, dy = y + 0.5 - cy;
                        const double u = dx * g1x + dy * g1y, w = dx * g2x + dy * g2y;
                        if (u * u + w * w <= 1.0 + 1e-4) { hit = true; break; }
                    }
                const float x0 = (float)(bx * 8) + 0.5f - cx, y0 = (float)(by * 8) + 0.5f - cy;
                const bool in_aabb = !(x0 > hx || x0 + 7.0f < -hx || y0 > hy || y0 + 7.0f < -hy);
                const bool keep = in_aabb && ellipse_min_q(x0, x0 + 7.0f, y0, y0 + 7.0f, qxx, qxy, qyy) <= 1.0f + kEllipseSlack;
                blocks_aabb += in_aabb;
                kept += keep;
                exact += hit;
                if (hit && !keep) ++dropped;
            }
    }
    printf("{\"blocks_in_aabb\": %lld, \"kept\": %lld, \"exact\": %lld, \"dropped_hits\": %lld}\n", blocks_aabb, kept, exact, dropped);
    return dropped == 0 ? 0 : 1;
}
