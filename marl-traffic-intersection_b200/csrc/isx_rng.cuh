// isx_rng.cuh — counter-based per-env randomness (__host__ __device__).
//
// The reference draws NPC randomness from two unseedable thread_local std::mt19937
// (TrafficFlow.cpp:278,324).  Here every env owns a Philox4x32-10 stream (Salmon et al., SC'11) keyed by
// (seed) and indexed by (global env id, tick, word index), so results do not depend on how envs are
// sharded over GPUs.  The float / integer mappings are the ones libstdc++ 13 applies to a 32-bit URBG:
// generate_canonical<float,24> (bits/random.tcc:3349-3381) and Lemire's bounded integers
// (bits/uniform_int_dist.h:257-281,323-329), so that `u < spawn_prob` and the drawn route are bit-exact
// against the oracle build, whose URBG reads the same stream.
#pragma once
#include "isx_math.cuh"

namespace isx {

constexpr uint32_t TAG_TRAFFIC = 0x54524146u;  // 'TRAF'
constexpr uint32_t TAG_ACTION = 0x41435431u;   // 'ACT1'

struct U4 { uint32_t x, y, z, w; };

ISX_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    const uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

ISX_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c.x, h0, l0);
        mulhilo(0xCD9E8D57u, c.z, h1, l1);
        U4 n;
        n.x = h1 ^ c.y ^ k0;
        n.y = l1;
        n.z = h0 ^ c.w ^ k1;
        n.w = l0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// Word j of the traffic stream of (env, tick).
struct TrafficStream {
    uint32_t k0, k1, env, tick;
    uint32_t j;       // next word index == words consumed so far
    U4 blk;
    uint32_t blk_id;
    ISX_HDM void init(uint64_t seed, uint32_t env_, uint32_t tick_) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); env = env_; tick = tick_; j = 0; blk_id = 0xffffffffu;
    }
    ISX_HDM uint32_t next() {
        const uint32_t b = j >> 2;
        if (b != blk_id) { blk = philox4x32_10(U4{env, tick, b, TAG_TRAFFIC}, k0, k1); blk_id = b; }
        const uint32_t lane = j & 3u;
        ++j;
        return lane == 0 ? blk.x : lane == 1 ? blk.y : lane == 2 ? blk.z : blk.w;
    }
    // std::uniform_real_distribution<float>(0,1) over a 32-bit URBG
    ISX_HDM float uniform01() {
        float r = (float)next() * (1.0f / 4294967296.0f);
        if (r >= 1.0f) r = u2f(0x3f7fffffu);
        return r;
    }
    // std::uniform_int_distribution<size_t>(0, n-1) over a 32-bit URBG
    ISX_HDM uint32_t below(uint32_t n) {
        uint32_t hi, lo;
        mulhilo(next(), n, hi, lo);
        if (lo < n) {
            const uint32_t thr = (0u - n) % n;
            while (lo < thr) mulhilo(next(), n, hi, lo);
        }
        return hi;
    }
};

ISX_HD float action_from_word(uint32_t u) { return (float)(u >> 8) * (1.0f / 8388608.0f) - 1.0f; }

ISX_HD void philox_action(uint64_t seed, uint32_t env, uint32_t tick, uint32_t agent, float& throttle, float& steer) {
    const U4 o = philox4x32_10(U4{env, tick, agent, TAG_ACTION}, (uint32_t)seed, (uint32_t)(seed >> 32));
    throttle = action_from_word(o.x);
    steer = action_from_word(o.y);
}

}  // namespace isx
