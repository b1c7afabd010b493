/* oracle/ref_rng_shim.h — TEST INFRASTRUCTURE (not product code).
 *
 * Force-included in front of /root/reference/cpp/TrafficFlow.cpp ONLY.  The
 * reference draws NPC randomness from two unseedable
 * `static thread_local std::mt19937 rng{std::random_device{}()}`
 * (TrafficFlow.cpp:278 and :324).  After <random> has been fully parsed we
 * rename those two identifiers so that the very same source text instantiates
 * a 32-bit URBG that pulls words from the harness-controlled per-env Philox
 * stream (oracle/philox.h).  libstdc++'s own uniform_real_distribution<float>
 * / uniform_int_distribution<size_t> still do the float / Lemire mapping, so
 * the distribution arithmetic is the reference's, only the bit source changes.
 */
#pragma once
#include <random>
#include <cstdint>
extern "C" uint32_t isx_ref_next_u32(void);
namespace std {
struct isx_ref_urbg {
    using result_type = uint32_t;
    explicit isx_ref_urbg(uint32_t = 0) {}
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xffffffffu; }
    result_type operator()() { return isx_ref_next_u32(); }
};
struct isx_ref_rd {
    uint32_t operator()() { return 0u; }
};
}  // namespace std
#define mt19937 isx_ref_urbg
#define random_device isx_ref_rd
