/* oracle/isx_oracle.h — TEST INFRASTRUCTURE (not product code).  C ABI of the plain-C restatement
 * (isx_oracle.c); same surface as ref_driver.cpp with the prefix isxo_ instead of isxref_. */
#ifndef ISX_ORACLE_H
#define ISX_ORACLE_H
#include <stdint.h>
#include "isx_state.h"
#ifdef __cplusplus
extern "C" {
#endif
void *isxo_create(int num_lanes);
void isxo_destroy(void *h);
void isxo_configure(void *h, int use_team, int respawn, int max_steps);
void isxo_configure_traffic(void *h, int enabled, float density);
void isxo_configure_routes(void *h, int n, const char *const *starts, const char *const *ends);
void isxo_set_reward(void *h, const float *k8);
void isxo_set_ego_routes(void *h, int n, const char *const *starts, const char *const *ends);
void isxo_set_lidar_rays(void *h, int rays);
int isxo_reset(void *h);
void isxo_seed(void *h, uint64_t seed, uint32_t env_id, uint32_t tick);
uint32_t isxo_tick(void *h);
int isxo_num_agents(void *h);
int isxo_num_npcs(void *h);
int isxo_step_count(void *h);
void isxo_set_step_count(void *h, int s);
void isxo_get_obs(void *h, float *obs);
int isxo_step(void *h, const float *throttle, const float *steer, int n_actions, float dt, float *obs, float *reward,
              int32_t *done, int32_t *status, int32_t *terminated, int32_t *truncated, int32_t *agents_alive);
void isxo_get_events(void *h, isx_traffic_events *ev);
int isxo_get_egos(void *h, isx_car_state *out);
int isxo_get_npcs(void *h, isx_car_state *out, int cap);
int isxo_get_lidar(void *h, int agent, float *dist, int cap);
void isxo_set_egos(void *h, const isx_car_state *s, int n);
void isxo_set_npcs(void *h, const isx_car_state *s, int n);
int isxo_std_sort(const float *keys, int n, int32_t *perm_out);
long long isxo_rollout(void *h, int steps, float dt, int32_t *status_hist6, double *reward_sum);
#ifdef __cplusplus
}
#endif
#endif
