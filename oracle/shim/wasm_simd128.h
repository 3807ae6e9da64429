/* Build shim: the 4 wasm SIMD intrinsics used by the reference's sorter.cpp, expressed with GCC
 * vector extensions so the SIMD spelling of the dot product can also be compiled natively.
 * Test infrastructure only. */
#pragma once
#include <string.h>
typedef int v128_t __attribute__((vector_size(16)));
static inline v128_t wasm_v128_load(const void* p) { v128_t v; memcpy(&v, p, 16); return v; }
static inline void wasm_v128_store(void* p, v128_t v) { memcpy(p, &v, 16); }
static inline v128_t wasm_i32x4_mul(v128_t a, v128_t b) { return a * b; }
