/* oracle/isx_state.h — TEST INFRASTRUCTURE (not product code).
 *
 * Plain-C exchange records shared by the two CPU checkers in this directory
 * (ref_driver.cpp = the reference's own C++ behind a C ABI, and isx_oracle.c =
 * the C restatement).  The CUDA product declares its own, layout-identical
 * records in include/isx.h; tests compare the two through ctypes.
 */
#ifndef ISX_ORACLE_STATE_H
#define ISX_ORACLE_STATE_H
#include <stdint.h>

#define ISX_OBS_DIM 127
#define ISX_PATH_LEN 160

/* status codes for the strings of /root/reference/cpp/IntersectionEnv.cpp:147,169,205,227,240,282,302 */
enum {
    ISX_ALIVE = 0,
    ISX_DEAD = 1,
    ISX_SUCCESS = 2,
    ISX_CRASH_WALL = 3,
    ISX_CRASH_LINE = 4,
    ISX_CRASH_CAR = 5
};

/* One car (ego or NPC).  Mirrors the public members of Car (/root/reference/cpp/Car.h:16-46)
 * that influence the simulation; `route` indexes the handle's route table (egos: slot i -> i,
 * NPCs: index into the traffic-route list) and stands for the car's path + intention. */
typedef struct isx_car_state {
    float x, y, v, heading;          /* State            Car.h:9-14  */
    float acc, steer;                /* acc, steering_angle  :23-24  */
    float prev_dist, prev_a0, prev_a1; /* prev_dist_to_goal, prev_action  :36-37 */
    int32_t path_index;              /* :34 */
    int32_t route;
    int32_t alive;                   /* :27 */
    uint32_t uid;                    /* NPC spawn serial within its env (0 for egos) */
    int32_t intention;               /* :32 (derived from route; reported for checks) */
} isx_car_state;

/* Per-env, per-step NPC event record ("spawn/removal events must be bit-exact"). */
typedef struct isx_traffic_events {
    int32_t rng_draws;      /* 32-bit words consumed from the env's traffic stream this step */
    int32_t spawn_route;    /* traffic-route index drawn this step, -1 if no attempt */
    int32_t spawned;        /* 1 if the attempt was not blocked and an NPC was appended */
    uint32_t removed_mask;  /* bit i: NPC at list position i (after the append, before the erase) was erased */
    uint32_t collided_mask; /* subset of removed_mask erased because of an NPC-NPC collision */
    int32_t npc_count;      /* NPCs alive after the step */
} isx_traffic_events;

#endif
