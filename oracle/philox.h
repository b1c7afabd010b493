/* oracle/philox.h — TEST INFRASTRUCTURE (not product code).
 *
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
 * as 1, 2, 3", SC'11), restated from the paper.  The reference itself is
 * unseedable (two `static thread_local std::mt19937` seeded from
 * std::random_device, /root/reference/cpp/TrafficFlow.cpp:278,324); parity on
 * "identical seeds" therefore needs a defined stream.  The stream layout below
 * is THIS repo's definition; the CUDA product implements the same layout
 * independently (csrc/isx_rng.cuh) and tests compare the two.
 *
 *   key      = (seed_lo, seed_hi)
 *   traffic  : word j of (env g, tick t) = philox(ctr=(g, t, j>>2, ISX_TAG_TRAFFIC))[j&3]
 *   actions  : (throttle, steer) of (env g, tick t, agent a)
 *              = u2f(philox(ctr=(g, t, a, ISX_TAG_ACTION))[0..1]),
 *              u2f(u) = (float)(u >> 8) * 2^-23 - 1   (exact in f32, in [-1, 1))
 *   tick     = per-env counter, +1 at the start of every step(), never reset by reset().
 */
#ifndef ISX_ORACLE_PHILOX_H
#define ISX_ORACLE_PHILOX_H
#include <stdint.h>

#define ISX_TAG_TRAFFIC 0x54524146u /* 'TRAF' */
#define ISX_TAG_ACTION  0x41435431u /* 'ACT1' */

static inline void isx_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float isx_u32_to_action(uint32_t u) {
    return (float)(u >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

static inline void isx_action_for(uint64_t seed, uint32_t env, uint32_t tick, uint32_t agent,
                                  float *throttle, float *steer) {
    uint32_t ctr[4] = { env, tick, agent, ISX_TAG_ACTION };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    isx_philox4x32_10(ctr, key, o);
    *throttle = isx_u32_to_action(o[0]);
    *steer = isx_u32_to_action(o[1]);
}

static inline uint32_t isx_traffic_word(uint64_t seed, uint32_t env, uint32_t tick, uint32_t j) {
    uint32_t ctr[4] = { env, tick, j >> 2, ISX_TAG_TRAFFIC };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    isx_philox4x32_10(ctr, key, o);
    return o[j & 3];
}
#endif
