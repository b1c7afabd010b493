/* oracle/ref_pystream.cpp — TEST INFRASTRUCTURE (not product code).
 *
 * Linked into oracle/_ref/MARLEnv.so, the reference's own pybind11 module (bindings.cpp:11-95) built by `make pyref`.
 * TrafficFlow.cpp is compiled with ref_rng_shim.h, so its two RNG objects pull 32-bit words from isx_ref_next_u32();
 * this file supplies that stream for the Python-driven module: the harness calls isxpy_seed(seed, env, tick) (through
 * ctypes on the same .so) before every env.step(), exactly as ref_driver.cpp positions the stream for its C ABI. */
#include <stdint.h>

#include "philox.h"

namespace {
struct Stream { uint64_t seed = 0; uint32_t env = 0, tick = 0, j = 0; };
thread_local Stream g_stream;
}  // namespace

extern "C" {
uint32_t isx_ref_next_u32(void) {
    const uint32_t w = isx_traffic_word(g_stream.seed, g_stream.env, g_stream.tick, g_stream.j);
    g_stream.j++;
    return w;
}
void isxpy_seed(uint64_t seed, uint32_t env, uint32_t tick) { g_stream = Stream{seed, env, tick, 0}; }
uint32_t isxpy_words_drawn(void) { return g_stream.j; }
}
