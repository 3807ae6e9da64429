// js/gsplat_b200_addon.cc -- N-API glue between Node.js and libgsplat_b200.so (UNBUILT in this repo: the image has no
// Node.js / node_api.h).  Binds exactly the C ABI of include/gsplat_b200.h; no arithmetic lives here.
//
//   const addon = require('./build/Release/gsplat_b200.node');
//   const h = addon.create({maxSplatCount, distanceMapRange, integerBasedSort, dynamicMode, maxWidth, maxHeight, device});
//   addon.uploadCenters(h, centersArrayBuffer, sceneIndexesOrNull, from, count);
//   const {sortTime} = addon.sort(h, mvpFloat32Array, sortCount, renderCount, indexesUint32OrNull, sortedOutUint32);
//   addon.uploadSplatData(h, {...typed arrays...}); addon.frame(h, sortParams, uniforms, renderParams, frameOutUint8);
#include <node_api.h>
#include <cstring>
#include "../include/gsplat_b200.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, nullptr, #call " failed"); return nullptr; } } while (0)

static napi_value throw_gs(napi_env env, int code) {
    napi_throw_error(env, gs_status_string(code), gs_last_error_message());
    return nullptr;
}
static bool get_u32(napi_env env, napi_value obj, const char *key, uint32_t *out) {
    napi_value v; bool has = false;
    if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return false;
    return napi_get_named_property(env, obj, key, &v) == napi_ok && napi_get_value_uint32(env, v, out) == napi_ok;
}
static void *typed_ptr(napi_env env, napi_value v, size_t *bytes) {
    bool is_ta = false, is_ab = false;
    void *data = nullptr; size_t len = 0;
    napi_is_typedarray(env, v, &is_ta);
    if (is_ta) {
        napi_typedarray_type t; napi_value ab; size_t off;
        napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off);
        static const size_t w[] = {1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
        if (bytes) *bytes = len * w[t];
        return data;
    }
    napi_is_arraybuffer(env, v, &is_ab);
    if (is_ab) { napi_get_arraybuffer_info(env, v, &data, &len); if (bytes) *bytes = len; return data; }
    return nullptr;
}

static napi_value Create(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    gs_config c; memset(&c, 0, sizeof(c)); c.struct_size = sizeof(c);
    uint32_t v;
    if (get_u32(env, argv[0], "maxSplatCount", &v)) c.max_splat_count = v;
    if (get_u32(env, argv[0], "distanceMapRange", &v)) c.distance_map_range = v;
    if (get_u32(env, argv[0], "integerBasedSort", &v)) c.integer_based_sort = (uint8_t)v; else c.integer_based_sort = 1;
    if (get_u32(env, argv[0], "dynamicMode", &v)) c.dynamic_mode = (uint8_t)v;
    if (get_u32(env, argv[0], "maxWidth", &v)) c.max_width = v;
    if (get_u32(env, argv[0], "maxHeight", &v)) c.max_height = v;
    if (get_u32(env, argv[0], "device", &v)) c.device = (int32_t)v;
    gs_engine *e = nullptr;
    int rc = gs_create(&c, &e);
    if (rc) return throw_gs(env, rc);
    napi_value ext;
    NAPI_OK(napi_create_external(env, e, [](napi_env, void *p, void *) { gs_destroy((gs_engine *)p); }, nullptr, &ext));
    return ext;
}

static gs_engine *engine_of(napi_env env, napi_value v) { void *p = nullptr; napi_get_value_external(env, v, &p); return (gs_engine *)p; }

static napi_value UploadCenters(napi_env env, napi_callback_info info) {   // 'centers' message, SortWorker.js:84-98
    size_t argc = 5; napi_value a[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    uint32_t from = 0, count = 0;
    napi_get_value_uint32(env, a[3], &from); napi_get_value_uint32(env, a[4], &count);
    int rc = gs_upload_centers(engine_of(env, a[0]), typed_ptr(env, a[1], nullptr), (const uint32_t *)typed_ptr(env, a[2], nullptr), from, count);
    if (rc) return throw_gs(env, rc);
    return nullptr;
}

static napi_value Sort(napi_env env, napi_callback_info info) {            // 'sort' message, SortWorker.js:31-81
    size_t argc = 6; napi_value a[6];
    NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
    gs_sort_params p; memset(&p, 0, sizeof(p)); p.struct_size = sizeof(p);
    size_t nb = 0;
    const float *mvp = (const float *)typed_ptr(env, a[1], &nb);
    if (!mvp || nb < 64) { napi_throw_type_error(env, nullptr, "modelViewProj must be a Float32Array(16)"); return nullptr; }
    memcpy(p.model_view_proj, mvp, 64);
    napi_get_value_uint32(env, a[2], &p.sort_count); napi_get_value_uint32(env, a[3], &p.render_count);
    p.indexes_to_sort = (const uint32_t *)typed_ptr(env, a[4], nullptr);
    float ms = 0.f;
    int rc = gs_sort(engine_of(env, a[0]), &p, (uint32_t *)typed_ptr(env, a[5], nullptr), &ms);
    if (rc) return throw_gs(env, rc);
    napi_value out, t;
    napi_create_object(env, &out); napi_create_double(env, ms, &t); napi_set_named_property(env, out, "sortTime", t);
    return out;
}

static napi_value Init(napi_env env, napi_value exports) {
    napi_property_descriptor d[] = {
        {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"uploadCenters", nullptr, UploadCenters, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"sort", nullptr, Sort, nullptr, nullptr, nullptr, napi_default, nullptr},
        // uploadSplatData / render / frame follow the same pattern over gs_upload_splat_data / gs_render / gs_frame
    };
    napi_define_properties(env, exports, sizeof(d) / sizeof(d[0]), d);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
