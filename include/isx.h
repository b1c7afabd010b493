/* include/isx.h — C ABI of libisx_b200.so, the B200-native batched stepper for the intersection env.
 *
 * This is the drop-in boundary.  The reference has no C ABI: its boundary is the pybind11 module
 * `MARLEnv` (/root/reference/cpp/bindings.cpp:11-95) that /root/reference/env.py drives.  Each entry
 * point below names the reference interface it replaces.  Signatures are plain C (pointers, sizes,
 * no torch / STL types); all *_dev pointers are CUDA device pointers on the handle's device; every
 * call that takes a `stream` is asynchronous and stream-ordered (`stream` is a cudaStream_t, 0 = the
 * legacy default stream).  Every function returns 0 on success or a negative ISX_E_* code; the text of
 * the last error on the calling thread is isx_last_error().  There is NO CPU fallback: isx_create
 * fails (ISX_E_CUDA) when no CUDA device is usable.
 *
 * One handle = E env instances stepping in lockstep on one GPU (struct-of-arrays device buffers).
 * Multi-GPU = one process and one handle per GPU with env_id_base = the shard's first global env id;
 * results are independent of the sharding because all randomness is keyed by the GLOBAL env id.
 */
#ifndef ISX_H
#define ISX_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISX_ABI_VERSION 2
#define ISX_OBS_DIM 127          /* IntersectionEnv.cpp:424 : 6 ego + 5x5 neighbours + 96 lidar */
#define ISX_PATH_LEN 160         /* RouteGen.cpp:111-205: 50 + 60 + 50 way-points */
#define ISX_MAX_AGENTS 32        /* egos per env */
#define ISX_MAX_GROUPS 64         /* config groups of one heterogeneous batch (isx_create_groups) */
#define ISX_MAX_NPC 32           /* NPC slots per env (reference is unbounded; observed max 13) */
#define ISX_MAX_RAYS 96
#define ISX_MAX_ROUTES 64        /* traffic routes */

enum { ISX_OK = 0, ISX_E_ARG = -1, ISX_E_CUDA = -2, ISX_E_ROUTE_START = -3, ISX_E_ROUTE_END = -4, ISX_E_STATE = -5 };

/* status codes = the strings of IntersectionEnv.cpp:147,169,205,227,240,282,302 */
enum { ISX_ALIVE = 0, ISX_DEAD = 1, ISX_SUCCESS = 2, ISX_CRASH_WALL = 3, ISX_CRASH_LINE = 4, ISX_CRASH_CAR = 5 };

typedef struct isx_handle isx_handle;
typedef struct isx_snapshot isx_snapshot;

/* Everything env.py passes through IntersectionEnv(num_lanes) / configure / configure_traffic /
 * configure_routes / reward_config.* / add_car_with_route (env.py:111-131,147-152). */
typedef struct isx_config {
    int32_t abi_version;         /* = ISX_ABI_VERSION */
    int32_t device;              /* CUDA device ordinal */
    int32_t num_envs;            /* E on this device */
    int32_t num_agents;          /* N egos per env (= number of ego routes) */
    int32_t num_lanes;           /* IntersectionEnv(num_lanes), bindings.cpp:59 */
    int32_t lidar_rays;          /* 96 = add_car_with_route (IntersectionEnv.cpp:113); 72 = default Lidar (Lidar.h:11) */
    int32_t npc_capacity;        /* <= ISX_MAX_NPC; a spawn into a full env is dropped and counted */
    int32_t use_team_reward;     /* configure(use_team, respawn, max_steps), IntersectionEnv.cpp:50-54 */
    int32_t respawn_enabled;
    int32_t max_steps;
    int32_t traffic_flow;        /* configure_traffic(enabled, density), :56-60 */
    float traffic_density;
    float reward[8];             /* k_prog, v_min_ms, k_stuck, k_cv, k_co, k_succ, k_sm, alpha (Reward.h:5-14) */
    const char *const *ego_start;     /* N lane ids, add_car_with_route(start,end) :78 */
    const char *const *ego_end;
    int32_t num_traffic_routes;       /* configure_routes(routes) :62-64 */
    const char *const *traffic_start;
    const char *const *traffic_end;
    uint64_t seed;               /* Philox key (the reference is unseedable, TrafficFlow.cpp:278,324) */
    int64_t env_id_base;         /* global id of local env 0 */
    int32_t auto_reset;          /* The reference has no auto-reset: a caller of env.py resets after terminated|truncated.
                                  * 0 = the same here (isx_reset with a mask).
                                  * 1 = the env is reset at the START of the step that follows a terminated|truncated step, and that
                                  *     step's action already drives the new episode (it was chosen from the terminal observation;
                                  *     the reset observation itself is never returned).  What isx_rollout's random-action benchmark uses.
                                  * 2 = next-step reset as Gymnasium's vector envs do it: the call after a terminated|truncated step only
                                  *     resets — action ignored, nothing simulated (step 0, RNG tick unchanged), reward 0, done 0,
                                  *     terminated = truncated = 0, and obs is what reset() returns (lidar part 1.0). */
    int32_t reserved;
} isx_config;

/* Same fields, same meaning, same layout as oracle/isx_state.h (the checkers use this record too). */
typedef struct isx_car_state {
    float x, y, v, heading;      /* Car.h:9-14 */
    float acc, steer;            /* Car.h:23-24 */
    float prev_dist, prev_a0, prev_a1; /* Car.h:36-37 */
    int32_t path_index;          /* Car.h:34 */
    int32_t route;               /* ego: slot index; NPC: traffic-route index */
    int32_t alive;               /* Car.h:27 */
    uint32_t uid;                /* NPC spawn serial within its env */
    int32_t intention;           /* Car.h:32 */
} isx_car_state;

typedef struct isx_traffic_events {
    int32_t rng_draws;           /* 32-bit words consumed from the env's traffic stream this step */
    int32_t spawn_route;         /* traffic-route index drawn, -1 if no attempt */
    int32_t spawned;             /* 1 if an NPC was appended */
    uint32_t removed_mask;       /* bit i: NPC at list position i (after append, before erase) was erased */
    uint32_t collided_mask;      /* subset erased because of an NPC-NPC collision */
    int32_t npc_count;           /* NPCs after the step */
} isx_traffic_events;

/* Non-owning device views, valid until isx_destroy.  Replaces StepResult (Reward.h:16-29) and the
 * by-value `cars` / `traffic_cars` / `lidars` snapshots of bindings.cpp:60-62. */
typedef struct isx_buffers {
    float *obs;                  /* [E][N][127] f32 */
    float *reward;               /* [E][N] */
    uint8_t *done;               /* [E][N] */
    uint8_t *status;             /* [E][N] ISX_* status code */
    uint8_t *terminated;         /* [E] */
    uint8_t *truncated;          /* [E] */
    int32_t *agents_alive;       /* [E] */
    int32_t *step;               /* [E] step_count */
    uint8_t *lidar_hit;          /* [E][N][96] first-hit sample index k (distance = 4k px), 0 = no hit (250 px) */
    /* ego state, SoA [E][N] */
    float *ego_x, *ego_y, *ego_v, *ego_heading, *ego_steer, *ego_acc, *ego_prev_dist, *ego_prev_a0, *ego_prev_a1;
    int32_t *ego_path_index;
    uint8_t *ego_alive;
    /* NPC state, SoA [E][npc_capacity]; first npc_count[e] slots are live, in list order */
    float *npc_x, *npc_y, *npc_v, *npc_heading, *npc_steer;
    int32_t *npc_path_index;
    int32_t *npc_route;
    uint32_t *npc_uid;
    int32_t *npc_count;          /* [E] */
    isx_traffic_events *events;  /* [E] */
    uint32_t *tick;              /* [E] RNG tick (never reset) */
} isx_buffers;

/* Counters accumulated on the device since create / isx_stats_reset; one all-reduce(sum) over ranks
 * gives job totals (that is the only collective this path needs). */
typedef struct isx_stats {
    int64_t agent_steps;
    int64_t status_hist[6];      /* per agent-step, index = ISX_* status */
    int64_t npc_spawned;
    int64_t npc_removed;
    int64_t npc_collided;
    int64_t npc_overflow;        /* spawns dropped because the env's NPC slots were full */
    int64_t env_resets;          /* auto-resets performed */
    double reward_sum;
    int64_t neighbor_tie_sorts;  /* observations whose neighbour order needed the exact std::sort replay (equal distances, > 16 neighbours) */
} isx_stats;

/* Environment switches read at isx_create (tuning / bisecting aids, all optional): ISX_NO_ORDER=1 (k_traffic steps its envs in index order instead of by NPC count), ISX_NO_PDL=1 (fully serialised kernel
 * launches instead of programmatic dependent launch), ISX_NO_GRAPH=1 (host-buffer step on the stream path instead of the
 * captured CUDA graph), ISX_PIPE_PLAN="w0,w1,..." (env ranges of the host-step pipeline), ISX_LIDAR_CTAS_PER_SM=n (cap of
 * the persistent beam-kernel grid), ISX_HOST_THREADS=n (host threads completing obs rows), ISX_TRACE=1 (per-env phase stamps, isx_trace_read),
 * ISX_TRAFFIC_LANES=8|16|32 (force a k_traffic instance; default: by batch size), ISX_GUARD=1 (red zones around every device
 * buffer, isx_debug_check_guards). */
const char *isx_last_error(void);
int isx_abi_version(void);

/* IntersectionEnv(num_lanes) + configure* + reset() + add_car_with_route x N  (env.py:111-136).
 * Unknown ego start id: the reference silently adds no car (IntersectionEnv.cpp:79-82) — a batched SoA
 * cannot have ragged N, so here that is ISX_E_ROUTE_START.  Unknown end id: std::out_of_range ->
 * IndexError in the reference (RouteGen.cpp:120) -> ISX_E_ROUTE_END here. */
int isx_create(const isx_config *cfg, isx_handle **out);
/* Heterogeneous batch (per-env routes / density / reward weights / lane count in one batch):
 * group g owns envs [first_g, first_g + cfgs[g].num_envs) of ONE set of buffers (obs [sum E][N][127], ...) and is
 * stepped with cfgs[g]'s own settings — what N separately configured IntersectionEnv objects are in the reference
 * (IntersectionEnv.h:27-35, env.py:111-131).  Each group keeps its own seed / env_id_base, so it evolves exactly
 * like a stand-alone batch created from cfgs[g].  device, num_agents, lidar_rays and (for groups with traffic)
 * npc_capacity must agree across groups.  isx_create(cfg) == isx_create_groups(cfg, 1). */
int isx_create_groups(const isx_config *cfgs, int32_t n_groups, isx_handle **out);
int isx_num_groups(isx_handle *h);
int isx_group_range(isx_handle *h, int32_t group, int32_t *first_env, int32_t *num_envs);
int isx_destroy(isx_handle *h);

/* reset() + add_car_with_route (IntersectionEnv.cpp:66-131) for the envs whose mask byte is non-zero
 * (all envs if env_mask_dev is NULL); refreshes obs like get_observations() at reset (lidar part = 1.0). */
int isx_reset(isx_handle *h, const uint8_t *env_mask_dev, void *stream);

/* step(throttles, steerings, dt)  (IntersectionEnv.cpp:133-392, bindings.cpp:76).
 * actions_dev = [E][N][2] f32 (throttle, steer).  Results land in the isx_buffers views. */
int isx_step(isx_handle *h, const float *actions_dev, float dt, void *stream);

/* Same call with HOST buffers (what env.py does around env.step, env.py:167-208): copies actions in,
 * steps, copies obs / reward / done / status / terminated / truncated out and synchronises.
 * Any output pointer may be NULL. */
int isx_step_host(isx_handle *h, const float *actions, float dt, float *obs, float *reward, uint8_t *done,
                  uint8_t *status, uint8_t *terminated, uint8_t *truncated, void *stream);

/* Zero-copy variant of isx_step_host: the library owns PINNED host staging buffers (isx_host_views); the caller writes
 * actions into the `actions` view, calls isx_step_pinned, and reads results from the other views (valid until the next
 * host-buffer step).  The device->host copies are pipelined behind the kernels shard by shard.  Synchronous. */
int isx_step_pinned(isx_handle *h, float dt, void *stream);
int isx_host_views(isx_handle *h, float **actions, float **obs, float **reward, uint8_t **done, uint8_t **status,
                   uint8_t **terminated, uint8_t **truncated);
/* What one host-buffer step moves and who completes the rows: bytes host->device (actions), bytes device->host (per agent a
 * compact obs record of 32 floats + lidar_rays hit-index bytes instead of the 127-float row — the lidar columns are exactly
 * float(4k)/250 of the hit index k and the columns behind the last beam are 0, so the rows are rebuilt bit-identically in
 * host memory by `host_threads` worker threads while later env ranges are still in flight; 0 threads = inline — plus the
 * reward/done/status/flags block), and the number of pipeline ranges.  ISX_HOST_THREADS=n overrides the thread count
 * (default: the CPUs this process may run on, at most 32). */
int isx_host_step_info(isx_handle *h, int64_t *h2d_bytes, int64_t *d2h_bytes, int32_t *host_threads, int32_t *ranges);
/* The host half of that transport on its own (pure host code, no device needed): n compact records (32 floats: obs[0..30]
 * + alive flag) and n x lidar_rays hit indices -> n obs rows of 127 floats, exactly as the device writes them. */
int isx_expand_obs_rows(const float *records32, const uint8_t *hit_index, int32_t lidar_rays, float *obs_rows, int64_t n_agents);
/* Two more pinned views filled by the same step: agents_alive [E] and step [E] (int32), StepResult.agents_alive / .step
 * of bindings.cpp:27-36. */
int isx_host_views_aux(isx_handle *h, int32_t **agents_alive, int32_t **step);

/* `steps` consecutive steps with actions drawn on the device from the Philox action stream
 * (DESIGN.md "RNG streams"); with auto_reset this is the random-action rollout BASELINE.json quotes. */
int isx_rollout(isx_handle *h, int32_t steps, float dt, void *stream);

/* isx_rollout with a CUDA-event pair around each kernel launch (on `stream`); returns the summed device
 * milliseconds of the dynamics kernel and of the lidar+observation kernel.  Synchronous.  Measurement aid. */
int isx_rollout_timed(isx_handle *h, int32_t steps, float dt, void *stream, float *ms_dynamics, float *ms_lidar_obs);
/* same, per kernel: ms4 = {k_traffic, k_ego, k_features, k_lidar_obs} */
int isx_rollout_timed4(isx_handle *h, int32_t steps, float dt, void *stream, float *ms4);

int isx_get_buffers(isx_handle *h, isx_buffers *out);
int isx_num_envs(isx_handle *h);
int isx_num_agents(isx_handle *h);

/* Snapshot / injection of one env (get_state / set_state, IntersectionEnv.cpp:394-416, EnvState.h:9-15).
 * Synchronous.  egos: N records; npcs: up to cap records, *n_npcs = live count. */
int isx_get_env_state(isx_handle *h, int32_t env, isx_car_state *egos, isx_car_state *npcs, int32_t cap,
                      int32_t *n_npcs, int32_t *step_count, uint32_t *tick);
int isx_set_env_state(isx_handle *h, int32_t env, const isx_car_state *egos, const isx_car_state *npcs,
                      int32_t n_npcs, int32_t step_count, uint32_t tick);
/* Whole-batch device snapshots: get_state()/set_state() of the reference ("fast MCTS rollbacks", EnvState.h:3-15,
 * IntersectionEnv.cpp:394-416) for all E envs at once, device to device, stream-ordered.  A snapshot holds the full
 * simulation state (egos, NPCs, counters, RNG tick) AND the last outputs, so the observation after a restore equals
 * the one at save time; the reference's set_state instead swaps in default 72-beam lidars (:411-415) — that quirk is
 * deliberately not reproduced.  restore with a mask rolls back only the envs whose byte is non-zero. */
int isx_snapshot_create(isx_handle *h, isx_snapshot **out);
int isx_snapshot_save(isx_handle *h, isx_snapshot *s, void *stream);
int isx_snapshot_restore(isx_handle *h, isx_snapshot *s, const uint8_t *env_mask_dev, void *stream);
int isx_snapshot_destroy(isx_snapshot *s);

/* Beam count of the handle's lidars at run time (Lidar.h:11-14: `rays` is a field of each Lidar object; the reference
 * swaps the objects — add_car_with_route builds 96-beam lidars, IntersectionEnv.cpp:112-128, set_state leaves default
 * 72-beam ones behind, :411-415).  Clears every stored hit (a fresh Lidar reads 250 px on all beams), zeroes the obs
 * columns beyond the new count and refreshes the observation.  Synchronous. */
int isx_set_lidar_rays(isx_handle *h, int32_t rays);
int isx_lidar_rays(isx_handle *h);

/* What a caller may change on a live reference env: reward_config.* (def_readwrite, bindings.cpp:33-42,63), configure() and
 * configure_traffic()'s density (IntersectionEnv.cpp:50-60).  group = config group of a heterogeneous batch, < 0 = all.
 * Effective from the next step; no buffers are rebuilt.  (Routes, agent count and traffic on/off shape the buffers: those
 * need a new handle.) */
int isx_set_reward(isx_handle *h, int32_t group, const float *k8);
int isx_configure_episode(isx_handle *h, int32_t group, int32_t use_team, int32_t respawn, int32_t max_steps);
int isx_set_traffic_density(isx_handle *h, int32_t group, float density);

/* Car.update(throttle, steer_input, dt) / Car.check_collision(other) of bindings.cpp:30-31 for detached car records,
 * evaluated on the GPU by the very device functions the step kernels use (Car.cpp:9-40, 105-141; 54x24 px cars).
 * isx_car_update rewrites x, y, v, heading, acc, steer of *car in place.  Synchronous, unit-level (not a fast path). */
int isx_car_update(int32_t device, isx_car_state *car, float throttle, float steer_input, float dt);
int isx_car_check_collision(int32_t device, const isx_car_state *a, const isx_car_state *b, int32_t *collide);

/* get_observations() (IntersectionEnv.cpp:418-520) for the current state, without stepping. */
int isx_observe(isx_handle *h, void *stream);

/* Headless debug picture of one env (stands in for the Windows-only window of
 * Renderer.cpp:520-646): rgb_dev = DEVICE uint8 [750][750][3].  Road / grass / centre lines from the tables the
 * simulation itself uses, alive egos in a six-colour palette with a head marker, NPCs grey, and the lidar beams that
 * hit something (green) with their exact hit pixel (red), as draw_lidar does.  Stream-ordered. */
int isx_render(isx_handle *h, int32_t env, uint8_t *rgb_dev, void *stream);

int isx_stats_read(isx_handle *h, isx_stats *out);     /* synchronous */
int isx_stats_reset(isx_handle *h);
/* Tuning aid: with ISX_TRACE=1 in the environment at isx_create, k_traffic stamps clock64() at its phase boundaries
 * per env; this copies the [E][16] stamps out (slots 0..5 = phase boundaries, 6 = NPC count).  ISX_E_STATE when off. */
int isx_trace_read(isx_handle *h, long long *out16_per_env);
/* Debug aid: with ISX_GUARD=1 in the environment at isx_create every device buffer of the handle sits between two 4 KB red
 * zones holding a canary byte; this counts the red zones that were written to (0 = no kernel stored outside its buffers).
 * Synchronous.  ISX_E_STATE when guards are off. */
int isx_debug_check_guards(isx_handle *h, int64_t *violations);
/* Tuning aid: one host-buffer step (stream path) with CUDA events around every pipeline range; ms[4*i+0..3] = kernels
 * begin / kernels end / copy begin / copy end of range i, ms since the step began.  Returns the number of ranges. */
int isx_pipe_timeline(isx_handle *h, float dt, void *stream, float *ms, int32_t cap_ranges);
/* Device pointers for the one collective this path has: the reduced episode counters as int64[ISX_STATS_COUNTERS]
 * (indices ISX_STAT_*; integers only, so an all-reduce(sum) over int64 is exact) and, SEPARATELY, reward_sum as float64[1]
 * (all-reduce it as a double: a double's bit pattern must never ride in an integer sum).  Reduces the per-env counters on
 * `stream` and synchronises it; the pointers stay valid until isx_destroy and are refreshed by every call. */
#define ISX_STATS_COUNTERS 15
enum { ISX_STAT_HIST0 = 0, ISX_STAT_SPAWNED = 6, ISX_STAT_REMOVED = 7, ISX_STAT_COLLIDED = 8, ISX_STAT_OVERFLOW = 9,
       ISX_STAT_RESETS = 10, ISX_STAT_AGENT_STEPS = 11, ISX_STAT_TIE_SORTS = 12 };
int isx_stats_device_ptrs(isx_handle *h, void **counters_i64, int32_t *n_counters, void **reward_sum_f64, void *stream);

/* Host-side route table probe (RouteGen.cpp:7-205 restated in the library): returns the path length,
 * ISX_E_ROUTE_START / ISX_E_ROUTE_END for unknown ids.  path_xy = 320 floats. */
int isx_route(int32_t num_lanes, const char *start_id, const char *end_id, float *path_xy, int32_t *intent,
              float *spawn_x, float *spawn_y, float *spawn_heading);

/* Device self-test of the restated libm (csrc/isx_math.cuh): evaluates sincosf/tanf/atan2f/hypotf/wrap on
 * n host-supplied inputs ON THE GPU and returns the results, so tests can compare with the host libm. */
int isx_math_probe(int32_t device, int32_t n, const float *a, const float *b, float *sin_a, float *cos_a,
                   float *tan_a, float *atan2_ab, float *hypot_ab, float *wrap_a);

#ifdef __cplusplus
}
#endif
#endif /* ISX_H */
