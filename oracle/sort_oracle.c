/* oracle/sort_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * CPU restatement of the reference's depth-sort contract, written from its semantics (SURVEY.md
 * Appendix B), not from its text.  Follows /root/reference/src/worker/sorter.cpp:17-168 (scalar twin
 * sorter_no_simd.cpp:17-156):
 *   - distance stage, six branches                       sorter.cpp:29-140
 *   - range map + bucket                                  sorter.cpp:142-149
 *   - output order = reverse(stable ascending by bucket)  sorter.cpp:151-167
 * plus the two host-side pre-steps the sorter's inputs depend on:
 *   - integer centres  round(f32*1000)  w=1000            src/splatmesh/SplatMesh.js:1912-1926
 *   - float centres    w=1                                src/splatmesh/SplatMesh.js:1935-1948
 *
 * Pinned against the natively compiled reference (oracle/_ref/libsorter_ref.so) by
 * tests/test_oracle_sort.py and the committed vectors in tests/golden/.
 *
 * Arithmetic rules (all matter for bit-exactness):
 *   int32 multiply/add wrap (wasm i32; build with -fwrapv); f32 products/sums are unfused and
 *   evaluated left to right (-ffp-contract=off); `(int)(f * 1000.0)` / `* 4096.0` is an f64 product
 *   truncated toward zero; int->f32 conversions round to nearest even.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define GS_ORACLE_API __attribute__((visibility("default")))

enum { GSO_OK = 0, GSO_BAD_ARG = 1, GSO_BUCKET_RANGE = 2 };

/* third row of (mvp * T), both column-major 4x4, f32, left-to-right unfused (sorter.cpp:11-15) */
static void third_row_of_product(const float *mvp, const float *t, float out[4]) {
    for (int c = 0; c < 4; ++c) {
        float acc = mvp[2] * t[4 * c + 0];
        acc = acc + mvp[6] * t[4 * c + 1];
        acc = acc + mvp[10] * t[4 * c + 2];
        acc = acc + mvp[14] * t[4 * c + 3];
        out[c] = acc;
    }
}

static int32_t trunc_scaled(float f, double scale) { return (int32_t)((double)f * scale); }

/* Distance of one splat under the selected mode.  `row_i`/`row_f` hold the (scene-specific) third row. */
static int32_t dist_int3(const int32_t *c, const int32_t r[4]) {
    return (int32_t)((uint32_t)r[0] * (uint32_t)c[0] + (uint32_t)r[1] * (uint32_t)c[1] + (uint32_t)r[2] * (uint32_t)c[2]);
}
static int32_t dist_int4(const int32_t *c, const int32_t r[4]) {
    return (int32_t)((uint32_t)r[0] * (uint32_t)c[0] + (uint32_t)r[1] * (uint32_t)c[1] + (uint32_t)r[2] * (uint32_t)c[2] +
                     (uint32_t)r[3] * (uint32_t)c[3]);
}

/* Stage 1: per-position distances for positions [s0, rc).  Returns min/max with the reference's seeds. */
GS_ORACLE_API int gso_distances(const uint32_t *indexes, const void *centers, const void *precomputed, const float *mvp,
                                const uint32_t *scene_indexes, const float *transforms, uint32_t sort_count,
                                uint32_t render_count, int use_precomputed, int integer_sort, int dynamic_mode,
                                int32_t *dist_out, int32_t *min_out, int32_t *max_out) {
    if (sort_count > render_count) return GSO_BAD_ARG;
    const uint32_t s0 = render_count - sort_count;
    int32_t dmin = 2147483640, dmax = -2147483640; /* sorter.cpp:24-25 */
    const int32_t *ci = (const int32_t *)centers;
    const float *cf = (const float *)centers;

    /* per-scene rows are a pure function of the scene index, so they can be tabulated lazily */
    int32_t cached_scene = -1;
    int32_t row_i[4] = {0, 0, 0, 1};
    float row_f[4] = {0, 0, 0, 0};
    if (!use_precomputed && !dynamic_mode) {
        if (integer_sort) {
            row_i[0] = trunc_scaled(mvp[2], 1000.0);
            row_i[1] = trunc_scaled(mvp[6], 1000.0);
            row_i[2] = trunc_scaled(mvp[10], 1000.0);
        } else {
            row_f[0] = mvp[2]; row_f[1] = mvp[6]; row_f[2] = mvp[10];
        }
    }
    for (uint32_t i = s0; i < render_count; ++i) {
        const uint32_t g = indexes[i];
        int32_t d;
        if (use_precomputed) {
            d = integer_sort ? ((const int32_t *)precomputed)[g] : trunc_scaled(((const float *)precomputed)[g], 4096.0);
        } else if (dynamic_mode) {
            const int32_t sc = (int32_t)scene_indexes[g];
            if (sc != cached_scene) {
                third_row_of_product(mvp, transforms + 16 * (size_t)sc, row_f);
                if (integer_sort)
                    for (int k = 0; k < 4; ++k) row_i[k] = trunc_scaled(row_f[k], 1000.0);
                cached_scene = sc;
            }
            if (integer_sort) {
                d = dist_int4(ci + 4 * (size_t)g, row_i);
            } else {
                const float *c = cf + 4 * (size_t)g;
                float acc = row_f[0] * c[0];
                acc = acc + row_f[1] * c[1];
                acc = acc + row_f[2] * c[2];
                acc = acc + row_f[3] * c[3];
                d = trunc_scaled(acc, 4096.0);
            }
        } else if (integer_sort) {
            d = dist_int3(ci + 4 * (size_t)g, row_i);
        } else {
            const float *c = cf + 4 * (size_t)g;
            float acc = row_f[0] * c[0];
            acc = acc + row_f[1] * c[1];
            acc = acc + row_f[2] * c[2];
            d = trunc_scaled(acc, 4096.0);
        }
        dist_out[i] = d;
        if (d > dmax) dmax = d;
        if (d < dmin) dmin = d;
    }
    *min_out = dmin;
    *max_out = dmax;
    return GSO_OK;
}

/* Stage 2: the f32 range map (sorter.cpp:142-146).  Bucket of distance d. */
GS_ORACLE_API float gso_range_map(int32_t dmin, int32_t dmax, uint32_t range) {
    const float span = (float)dmax - (float)dmin;
    return (float)(range - 1u) / span;
}
GS_ORACLE_API int32_t gso_bucket(int32_t d, int32_t dmin, float range_map) {
    const int32_t rel = (int32_t)((uint32_t)d - (uint32_t)dmin);
    return (int32_t)((float)rel * range_map);
}

/* Whole sort.  out[0..s0) = indexes[0..s0); out[s0..rc) = reverse(stable ascending by bucket).
 * Written as a forward counting sort into a descending cursor so that it is visibly NOT the reference's loop
 * structure while producing the same permutation. */
GS_ORACLE_API int gso_sort_indexes(const uint32_t *indexes, const void *centers, const void *precomputed, const float *mvp,
                                   const uint32_t *scene_indexes, const float *transforms, uint32_t range,
                                   uint32_t sort_count, uint32_t render_count, int use_precomputed, int integer_sort,
                                   int dynamic_mode, uint32_t *out, int32_t *buckets_out /* optional, render_count */) {
    if (sort_count > render_count || range < 2) return GSO_BAD_ARG;
    const uint32_t s0 = render_count - sort_count;
    memcpy(out, indexes, (size_t)s0 * sizeof(uint32_t));
    if (sort_count == 0) return GSO_OK;
    int32_t *dist = (int32_t *)malloc((size_t)render_count * sizeof(int32_t));
    uint32_t *start = (uint32_t *)calloc((size_t)range + 1, sizeof(uint32_t));
    if (!dist || !start) { free(dist); free(start); return GSO_BAD_ARG; }
    int32_t dmin, dmax;
    int rc = gso_distances(indexes, centers, precomputed, mvp, scene_indexes, transforms, sort_count, render_count,
                           use_precomputed, integer_sort, dynamic_mode, dist, &dmin, &dmax);
    if (rc != GSO_OK) { free(dist); free(start); return rc; }
    const float rm = gso_range_map(dmin, dmax, range);
    for (uint32_t i = s0; i < render_count; ++i) {
        int32_t b = gso_bucket(dist[i], dmin, rm);
        const int32_t rel = (int32_t)((uint32_t)dist[i] - (uint32_t)dmin);
        /* the reference has no defined behaviour here (it writes frequencies[>= R]); the engine's documented choice:
           int32 wrap-around of (d - min) is an error, an f32-rounding overshoot to exactly R clamps to R-1 */
        if (dmax == dmin) b = 0;
        if (rel < 0 || b < 0) { free(dist); free(start); return GSO_BUCKET_RANGE; }
        if ((uint32_t)b >= range) b = (int32_t)range - 1;
        dist[i] = b;
        start[b]++;
    }
    /* first output slot of bucket b when buckets are laid out in DESCENDING order */
    uint32_t run = s0;
    for (int64_t b = (int64_t)range - 1; b >= 0; --b) {
        const uint32_t n = start[b];
        start[b] = run;
        run += n;
    }
    /* within a bucket, larger input positions come first */
    for (uint32_t i = render_count; i-- > s0;) out[start[dist[i]]++] = indexes[i];
    if (buckets_out) memcpy(buckets_out + s0, dist + s0, (size_t)sort_count * sizeof(int32_t));
    free(dist);
    free(start);
    return GSO_OK;
}

/* Host pre-step S0: integer centres (SplatMesh.js:1912-1926): Math.round(f32 * 1000.0) in f64, w = 1000. */
GS_ORACLE_API void gso_integer_centers(const float *xyz, uint32_t count, int32_t *out4) {
    for (uint32_t i = 0; i < count; ++i) {
        for (int k = 0; k < 3; ++k) out4[4 * (size_t)i + k] = (int32_t)floor((double)xyz[3 * (size_t)i + k] * 1000.0 + 0.5);
        out4[4 * (size_t)i + 3] = 1000;
    }
}
/* Host pre-step S0': float centres (SplatMesh.js:1935-1948): w = 1. */
GS_ORACLE_API void gso_float_centers(const float *xyz, uint32_t count, float *out4) {
    for (uint32_t i = 0; i < count; ++i) {
        for (int k = 0; k < 3; ++k) out4[4 * (size_t)i + k] = xyz[3 * (size_t)i + k];
        out4[4 * (size_t)i + 3] = 1.0f;
    }
}
