/* Build shim: lets the reference's sorter sources compile with a native g++ (no emscripten here).
 * This is NOT reference code; it only defines the one macro the reference sources use. */
#pragma once
#define EMSCRIPTEN_KEEPALIVE __attribute__((used, visibility("default")))
