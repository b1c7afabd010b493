// isx_device.cuh — device-side view of a handle: configuration + struct-of-arrays buffers in HBM.
// Passed to every kernel by value (a few hundred bytes of kernel parameter space).
#pragma once
#include "../../include/isx.h"
#include "isx_rng.cuh"
#include "isx_sim.cuh"

namespace isx {

struct RouteMeta {
    float spawn_x, spawn_y, spawn_h;
    int intent;
    F2 goal;        // path[159]
    F2 goal_prev;   // path[158]
};

constexpr int STAT_SLOTS = 16;   // per-env u32 counters; slot 14/15 hold reward_sum (double)
enum { ST_HIST0 = 0, ST_SPAWNED = 6, ST_REMOVED = 7, ST_COLLIDED = 8, ST_OVERFLOW = 9, ST_RESETS = 10, ST_STEPS = 11, ST_TIESORT = 12, ST_RSUM = 14 };

struct Dev {
    // ---- configuration
    int E, N, M, R, lanes;
    int use_team, respawn, max_steps, traffic, T, auto_reset;
    int traffic_lanes;           // k_traffic lanes per env: 0 = by batch size, or 8 / 16 / 32 (ISX_TRAFFIC_LANES at isx_create; tests)
    RewardCfg rc;
    float max_progress;          // hypotf(750, 750)
    uint64_t seed;
    long long env_base;
    // ---- constant tables
    const F2* route_path;        // [(N + T)][160]   ego slot i -> i, traffic route r -> N + r
    const RouteMeta* route_meta; // [(N + T)]
    const float* route_far2;     // [(N + T)][160]   far-window bound of Car::update_path_index (isx_sim.cuh path_far_table)
    const uint32_t* road_bits;   // [ROAD_ROWS][ROAD_WORDS]
    const uint8_t* road_skip;    // [SKIP_DIM][SKIP_DIM]
    int box_lo, box_hi;          // strip interior [box_lo, box_hi] x [0,749] (and transpose) is all road
    RoadAna ana;                 // analytic off-road bound of the beam march (isx_sim.cuh ray_safe_samples)
    const float* rel_angle;      // [R]
    // ---- ego state [E][N]
    float *ex, *ey, *ev, *eh, *esteer, *eacc, *epd, *epa0, *epa1;
    int* epidx;
    uint8_t* ealive;
    // ---- NPC state [E][M], per-env [E]
    float *nx, *ny, *nv, *nh, *nsteer;
    int* npidx;
    int* nroute;
    uint32_t* nuid;
    int* ncount;
    uint32_t* next_uid;
    int* step_count;
    uint32_t* tick;
    // ---- k_features -> k_lidar_obs scratch
    void* agent_rec;             // [E][N] AgentRec {x, y, heading, rect_base | -1}
    void* car_rect;              // [E][N+M] PixRect (lidar pixel rectangle of every car)
    uint32_t* cand;              // [E][N][N+M] packed beam candidates: car | first beam << 8 | span << 16 | kmin << 24
    int* cand_n;                 // [E][N]
    unsigned* ray_counter;       // [1] dynamic work counter of k_lidar_obs
    // ---- k_traffic_order -> k_traffic scratch (null: envs are stepped in index order)
    int* order;                  // [5][order_stride] env ids (local to this view) by NPC count 0, 1, 2, 3, >= 4
    unsigned* order_cnt;         // [8] fill of the five lists (zeroed by k_ego, which always follows k_traffic)
    int order_stride;            // envs of the whole handle (list k starts at order + k * order_stride)
    // ---- outputs
    float* obs_c;                // [E][N][32]  compact record for the host-buffer step: obs[0..30] + alive flag (isx_host_expand.cpp)
    uint8_t* hit_c;              // [E][N][R]   lidar hit indices, dense stride R (what crosses PCIe instead of the float lidar columns)
    float* obs;                  // [E][N][127]
    float* reward;               // [E][N]
    uint8_t *done, *status;      // [E][N]
    uint8_t *terminated, *truncated;   // [E]
    int* agents_alive;           // [E]
    uint8_t* lidar_hit;          // [E][N][96]
    isx_traffic_events* events;  // [E]
    uint32_t* env_stats;         // [E][STAT_SLOTS]
    unsigned long long* stats;   // [16] reduced counters (slot 15 = reward_sum as double bits)
    long long* trace;            // optional [E][16] clock64() phase stamps (ISX_TRACE=1 at create; tuning aid), else null
};

}  // namespace isx
