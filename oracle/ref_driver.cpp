/* oracle/ref_driver.cpp — TEST INFRASTRUCTURE (not product code).
 *
 * A C ABI around the UNMODIFIED reference simulation core, compiled from the
 * sources where they lie under /root/reference/cpp (see oracle/Makefile; the
 * output goes to oracle/_ref/libisx_ref.so, which is git-ignored).  This is the
 * golden checker: the C restatement (isx_oracle.c) and the CUDA product are both
 * compared with it.  It replaces, for test purposes, the pybind11 module of
 * /root/reference/cpp/bindings.cpp:11-95, and additionally exposes the Car
 * members that module hides (acc, steering_angle, prev_dist_to_goal,
 * prev_action; bindings.cpp:21-31), which state-injection parity needs.
 *
 * Everything here is the builder's own code; it only *calls* the reference.
 */
#include "IntersectionEnv.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstring>
#include <stdexcept>
#include <map>
#include <mutex>

#include "isx_state.h"
#include "philox.h"

namespace {

struct Stream {
    uint64_t seed = 0;
    uint32_t env = 0;
    uint32_t tick = 0;
    uint32_t j = 0;
};
thread_local Stream g_stream;

int status_code(const std::string &s) {
    if (s == "ALIVE") return ISX_ALIVE;
    if (s == "DEAD") return ISX_DEAD;
    if (s == "SUCCESS") return ISX_SUCCESS;
    if (s == "CRASH_WALL") return ISX_CRASH_WALL;
    if (s == "CRASH_LINE") return ISX_CRASH_LINE;
    if (s == "CRASH_CAR") return ISX_CRASH_CAR;
    return -1;
}

using Route = std::pair<std::string, std::string>;

struct RefEnv {
    IntersectionEnv env;
    std::vector<Route> ego_routes;
    std::vector<Route> traffic_routes;
    std::vector<std::vector<std::pair<float, float>>> traffic_paths;
    int lidar_rays = 96;
    uint64_t seed = 0;
    uint32_t env_id = 0;
    uint32_t tick = 0;
    uint32_t next_uid = 1;
    isx_traffic_events ev{};
    explicit RefEnv(int lanes) : env(lanes) {}

    void rebuild_traffic_paths() {
        traffic_paths.clear();
        for (const auto &r : traffic_routes) {
            int intent = determine_intent(env.lane_layout, r.first, r.second);
            traffic_paths.push_back(generate_path_cpp(env.lane_layout, env.num_lanes, intent, r.first, r.second));
        }
    }
    int match_route(const Car &c) const {
        for (size_t r = 0; r < traffic_paths.size(); ++r) {
            if (traffic_paths[r] == c.path) return int(r);
        }
        return -1;
    }
    void apply_lidar_variant() {
        if (lidar_rays == 96) return;
        /* 72-ray variant = default-constructed Lidar (Lidar.h:11, Lidar.cpp:4-14), exactly what
         * IntersectionEnv::set_state leaves behind (IntersectionEnv.cpp:411-415). */
        for (auto &l : env.lidars) l = Lidar();
    }
};

void fill_state(const Car &c, int route, uint32_t uid, isx_car_state *o) {
    o->x = c.state.x; o->y = c.state.y; o->v = c.state.v; o->heading = c.state.heading;
    o->acc = c.acc; o->steer = c.steering_angle;
    o->prev_dist = c.prev_dist_to_goal; o->prev_a0 = c.prev_action.first; o->prev_a1 = c.prev_action.second;
    o->path_index = c.path_index; o->route = route; o->alive = c.alive ? 1 : 0;
    o->uid = uid; o->intention = c.intention;
}

/* NPC uid is parked in Car::prev_dist_to_goal, which no NPC code path reads
 * (it is only touched by compute_progress for egos, IntersectionEnv.cpp:15-28). */
uint32_t npc_uid(const Car &c) { return uint32_t(c.prev_dist_to_goal); }

}  // namespace

extern "C" {

uint32_t isx_ref_next_u32(void) {
    uint32_t w = isx_traffic_word(g_stream.seed, g_stream.env, g_stream.tick, g_stream.j);
    g_stream.j++;
    return w;
}

void *isxref_create(int num_lanes) {
    RefEnv *r = new RefEnv(num_lanes);
    r->traffic_routes = r->env.traffic_routes; /* init_traffic_routes() default, TrafficFlow.cpp:198-238 */
    r->rebuild_traffic_paths();
    return r;
}
void isxref_destroy(void *h) { delete static_cast<RefEnv *>(h); }

void isxref_configure(void *h, int use_team, int respawn, int max_steps) {
    static_cast<RefEnv *>(h)->env.configure(use_team != 0, respawn != 0, max_steps);
}
void isxref_configure_traffic(void *h, int enabled, float density) {
    static_cast<RefEnv *>(h)->env.configure_traffic(enabled != 0, density);
}
void isxref_configure_routes(void *h, int n, const char *const *starts, const char *const *ends) {
    RefEnv *r = static_cast<RefEnv *>(h);
    std::vector<Route> routes;
    for (int i = 0; i < n; ++i) routes.emplace_back(starts[i], ends[i]);
    r->env.configure_routes(routes);
    r->traffic_routes = routes;
    r->rebuild_traffic_paths();
}
int isxref_num_traffic_routes(void *h) { return int(static_cast<RefEnv *>(h)->traffic_routes.size()); }
void isxref_get_traffic_route(void *h, int i, char *start, char *end, int cap) {
    RefEnv *r = static_cast<RefEnv *>(h);
    std::strncpy(start, r->traffic_routes[size_t(i)].first.c_str(), size_t(cap));
    std::strncpy(end, r->traffic_routes[size_t(i)].second.c_str(), size_t(cap));
}
/* k = {k_prog, v_min_ms, k_stuck, k_cv, k_co, k_succ, k_sm, alpha}  (Reward.h:5-14) */
void isxref_set_reward(void *h, const float *k) {
    RewardConfig &rc = static_cast<RefEnv *>(h)->env.reward_config;
    rc.k_prog = k[0]; rc.v_min_ms = k[1]; rc.k_stuck = k[2]; rc.k_cv = k[3];
    rc.k_co = k[4]; rc.k_succ = k[5]; rc.k_sm = k[6]; rc.alpha = k[7];
}
void isxref_set_ego_routes(void *h, int n, const char *const *starts, const char *const *ends) {
    RefEnv *r = static_cast<RefEnv *>(h);
    r->ego_routes.clear();
    for (int i = 0; i < n; ++i) r->ego_routes.emplace_back(starts[i], ends[i]);
}
void isxref_set_lidar_rays(void *h, int rays) { static_cast<RefEnv *>(h)->lidar_rays = rays; }
/* Mid-episode switch to the 72-beam variant: what IntersectionEnv::set_state does to the lidars (IntersectionEnv.cpp:411-415). */
void isxref_swap_lidars(void *h) { static_cast<RefEnv *>(h)->apply_lidar_variant(); }

/* env.reset() followed by add_car_with_route per ego, as env.py:147-152 does. */
int isxref_reset(void *h) {
    RefEnv *r = static_cast<RefEnv *>(h);
    r->env.reset();
    try {
        for (const auto &rt : r->ego_routes) r->env.add_car_with_route(rt.first, rt.second);
    } catch (const std::out_of_range &) {
        return -1; /* unknown end id -> std::out_of_range from .at(), RouteGen.cpp:120 */
    }
    r->apply_lidar_variant();
    r->next_uid = 1;
    return int(r->env.cars.size());
}

void isxref_seed(void *h, uint64_t seed, uint32_t env_id, uint32_t tick) {
    RefEnv *r = static_cast<RefEnv *>(h);
    r->seed = seed; r->env_id = env_id; r->tick = tick;
}
uint32_t isxref_tick(void *h) { return static_cast<RefEnv *>(h)->tick; }

int isxref_num_agents(void *h) { return int(static_cast<RefEnv *>(h)->env.cars.size()); }
int isxref_num_npcs(void *h) { return int(static_cast<RefEnv *>(h)->env.traffic_cars.size()); }
int isxref_step_count(void *h) { return static_cast<RefEnv *>(h)->env.step_count; }
void isxref_set_step_count(void *h, int s) { static_cast<RefEnv *>(h)->env.step_count = s; }

void isxref_get_obs(void *h, float *obs) {
    RefEnv *r = static_cast<RefEnv *>(h);
    auto o = r->env.get_observations();
    for (size_t i = 0; i < o.size(); ++i) std::memcpy(obs + i * ISX_OBS_DIM, o[i].data(), sizeof(float) * ISX_OBS_DIM);
}

/* One step.  Missing actions default to 0 inside the reference (IntersectionEnv.cpp:153-154);
 * n_actions lets tests exercise that.  Returns the step number. */
int isxref_step(void *h, const float *throttle, const float *steer, int n_actions, float dt,
                float *obs, float *reward, int32_t *done, int32_t *status,
                int32_t *terminated, int32_t *truncated, int32_t *agents_alive) {
    RefEnv *r = static_cast<RefEnv *>(h);
    std::vector<float> th(throttle, throttle + n_actions), st(steer, steer + n_actions);

    r->tick += 1;
    g_stream.seed = r->seed; g_stream.env = r->env_id; g_stream.tick = r->tick; g_stream.j = 0;

    std::vector<uint32_t> before;
    for (const auto &c : r->env.traffic_cars) before.push_back(npc_uid(c));

    StepResult res = r->env.step(th, st, dt);

    /* --- NPC event inference (uids tracked through an unused Car field) --- */
    isx_traffic_events ev{};
    ev.rng_draws = int32_t(g_stream.j);
    ev.spawn_route = -1;
    if (r->env.traffic_flow) {
        bool spawned = false;
        for (auto &c : r->env.traffic_cars) {
            if (npc_uid(c) == 0u) { /* fresh append: try_spawn_traffic_car sets prev_dist_to_goal = 0 */
                spawned = true;
                ev.spawn_route = r->match_route(c);
                c.prev_dist_to_goal = float(r->next_uid);
                before.push_back(r->next_uid);
                r->next_uid++;
            }
        }
        ev.spawned = spawned ? 1 : 0;
        size_t k = 0;
        for (size_t i = 0; i < before.size(); ++i) {
            if (k < r->env.traffic_cars.size() && npc_uid(r->env.traffic_cars[k]) == before[i]) { ++k; }
            else ev.removed_mask |= (1u << i);
        }
        if (!spawned && ev.rng_draws > 1) {
            /* an attempt was drawn but blocked: recover the drawn route from the stream
             * (Lemire's method as libstdc++ uniform_int_dist.h:257-281 applies it). */
            uint32_t n = uint32_t(r->traffic_routes.size());
            uint32_t j = 1;
            uint64_t prod = uint64_t(isx_traffic_word(r->seed, r->env_id, r->tick, j)) * n;
            uint32_t low = uint32_t(prod);
            if (low < n) {
                uint32_t thr = (0u - n) % n;
                while (low < thr) { ++j; prod = uint64_t(isx_traffic_word(r->seed, r->env_id, r->tick, j)) * n; low = uint32_t(prod); }
            }
            ev.spawn_route = int32_t(prod >> 32);
        }
    }
    ev.npc_count = int32_t(r->env.traffic_cars.size());
    r->ev = ev;

    const size_t n = r->env.cars.size();
    for (size_t i = 0; i < n; ++i) {
        if (obs) std::memcpy(obs + i * ISX_OBS_DIM, res.obs[i].data(), sizeof(float) * ISX_OBS_DIM);
        if (reward) reward[i] = res.rewards[i];
        if (done) done[i] = res.done[i];
        if (status) status[i] = status_code(res.status[i]);
    }
    if (terminated) *terminated = res.terminated ? 1 : 0;
    if (truncated) *truncated = res.truncated ? 1 : 0;
    if (agents_alive) *agents_alive = res.agents_alive;
    return res.step;
}

void isxref_get_events(void *h, isx_traffic_events *ev) { *ev = static_cast<RefEnv *>(h)->ev; }

/* n envs stepped by one call on `threads` host threads (the census tests compare hundreds of envs per step; one ctypes
 * call per env and step costs more than the simulation).  All arrays are dense over the env index; N = egos per env;
 * lidar_k[e][a][96] = hit sample index (distance / 4, 0 = none) and npc_pose[e][cap][4] = x, y, v, heading in list order
 * are optional.  Every env steps exactly as in isxref_step (same function, same stream positioning). */
void isxref_step_batch(void **handles, int n, const float *actions, int N, float dt, float *obs, float *reward, int32_t *done,
                       int32_t *status, int32_t *terminated, int32_t *truncated, int32_t *agents_alive, int32_t *step,
                       isx_traffic_events *events, uint8_t *lidar_k, float *npc_pose, int cap, int threads) {
    if (threads < 1) threads = 1;
    std::atomic<int> next{0};
    auto work = [&]() {
        std::vector<float> th((size_t)N), st((size_t)N);
        for (;;) {
            const int e = next.fetch_add(1);
            if (e >= n) return;
            const size_t eo = (size_t)e * (size_t)N;
            for (int a = 0; a < N; ++a) { th[(size_t)a] = actions[(eo + a) * 2]; st[(size_t)a] = actions[(eo + a) * 2 + 1]; }
            step[e] = isxref_step(handles[e], th.data(), st.data(), N, dt, obs + eo * ISX_OBS_DIM, reward + eo, done + eo, status + eo,
                                  terminated + e, truncated + e, agents_alive + e);
            RefEnv *r = static_cast<RefEnv *>(handles[e]);
            if (events) events[e] = r->ev;
            if (lidar_k) {
                for (int a = 0; a < N && size_t(a) < r->env.lidars.size(); ++a) {
                    const auto &d = r->env.lidars[size_t(a)].distances;
                    uint8_t *o = lidar_k + (eo + a) * 96;
                    for (size_t i = 0; i < 96; ++i) o[i] = (i < d.size() && d[i] < 250.0f) ? uint8_t(d[i] / 4.0f) : uint8_t(0);
                }
            }
            if (npc_pose) {
                int k = 0;
                for (const auto &c : r->env.traffic_cars) {
                    if (k >= cap) break;
                    float *o = npc_pose + ((size_t)e * cap + k) * 4;
                    o[0] = c.state.x; o[1] = c.state.y; o[2] = c.state.v; o[3] = c.state.heading;
                    ++k;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

int isxref_get_egos(void *h, isx_car_state *out) {
    RefEnv *r = static_cast<RefEnv *>(h);
    for (size_t i = 0; i < r->env.cars.size(); ++i) fill_state(r->env.cars[i], int(i), 0u, out + i);
    return int(r->env.cars.size());
}
int isxref_get_npcs(void *h, isx_car_state *out, int cap) {
    RefEnv *r = static_cast<RefEnv *>(h);
    int n = 0;
    for (const auto &c : r->env.traffic_cars) {
        if (n >= cap) break;
        fill_state(c, r->match_route(c), npc_uid(c), out + n);
        out[n].prev_dist = 0.0f;
        ++n;
    }
    return int(r->env.traffic_cars.size());
}
/* raw lidar distances (px), returns ray count */
int isxref_get_lidar(void *h, int agent, float *dist, int cap) {
    RefEnv *r = static_cast<RefEnv *>(h);
    const auto &d = r->env.lidars[size_t(agent)].distances;
    for (size_t i = 0; i < d.size() && int(i) < cap; ++i) dist[i] = d[i];
    return int(d.size());
}

/* State injection (resync parity).  Ego paths/routes stay those of the slot. */
void isxref_set_egos(void *h, const isx_car_state *s, int n) {
    RefEnv *r = static_cast<RefEnv *>(h);
    for (int i = 0; i < n && size_t(i) < r->env.cars.size(); ++i) {
        Car &c = r->env.cars[size_t(i)];
        c.state.x = s[i].x; c.state.y = s[i].y; c.state.v = s[i].v; c.state.heading = s[i].heading;
        c.acc = s[i].acc; c.steering_angle = s[i].steer;
        c.prev_dist_to_goal = s[i].prev_dist; c.prev_action = {s[i].prev_a0, s[i].prev_a1};
        c.path_index = s[i].path_index; c.alive = s[i].alive != 0;
    }
}
void isxref_set_npcs(void *h, const isx_car_state *s, int n) {
    RefEnv *r = static_cast<RefEnv *>(h);
    r->env.traffic_cars.clear();
    r->env.traffic_lidars.clear();
    for (int i = 0; i < n; ++i) {
        const Route &rt = r->traffic_routes[size_t(s[i].route)];
        Car c;
        c.intention = determine_intent(r->env.lane_layout, rt.first, rt.second);
        c.path = r->traffic_paths[size_t(s[i].route)];
        c.state.x = s[i].x; c.state.y = s[i].y; c.state.v = s[i].v; c.state.heading = s[i].heading;
        c.acc = s[i].acc; c.steering_angle = s[i].steer;
        c.path_index = s[i].path_index; c.alive = true;
        const auto sp = r->env.lane_layout.points.at(rt.first);
        c.spawn_state.x = sp.first; c.spawn_state.y = sp.second;
        c.prev_dist_to_goal = float(s[i].uid);
        if (s[i].uid >= r->next_uid) r->next_uid = s[i].uid + 1;
        r->env.traffic_cars.push_back(std::move(c));
        r->env.traffic_lidars.emplace_back();
    }
}

/* ---- unit probes of individual reference functions ---- */

/* RouteGen.cpp:7-205 + spawn heading as IntersectionEnv.cpp:87-92.  Returns path length, -1 on unknown id. */
int isxref_route(int num_lanes, const char *start, const char *end, float *path_xy, int *intent,
                 float *spawn_x, float *spawn_y, float *spawn_heading) {
    LaneLayout lay = build_lane_layout_cpp(num_lanes);
    auto it = lay.points.find(start);
    if (it == lay.points.end()) return -1;
    if (lay.points.find(end) == lay.points.end()) return -2;
    int in = determine_intent(lay, start, end);
    auto p = generate_path_cpp(lay, num_lanes, in, start, end);
    for (size_t i = 0; i < p.size(); ++i) { path_xy[2 * i] = p[i].first; path_xy[2 * i + 1] = p[i].second; }
    *intent = in;
    *spawn_x = it->second.first; *spawn_y = it->second.second;
    float hd = 0.0f;
    if (p.size() >= 2) hd = std::atan2(-(p[1].second - p[0].second), p[1].first - p[0].first);
    *spawn_heading = hd;
    return int(p.size());
}
int isxref_lane_point(int num_lanes, const char *id, float *x, float *y) {
    LaneLayout lay = build_lane_layout_cpp(num_lanes);
    auto it = lay.points.find(id);
    if (it == lay.points.end()) return -1;
    *x = it->second.first; *y = it->second.second;
    return 0;
}
int isxref_on_road(int num_lanes, float x, float y) { return RoadGeometry(num_lanes).is_on_road(x, y) ? 1 : 0; }
int isxref_yellow(int num_lanes, float x, float y) { return RoadGeometry(num_lanes).hits_yellow_line(x, y) ? 1 : 0; }
int isxref_is_line(int num_lanes, int x, int y) {
    static std::mutex mu;
    static std::map<int, LineMask *> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(num_lanes);
    if (it == cache.end()) it = cache.emplace(num_lanes, new LineMask(num_lanes)).first;
    return it->second->is_line(x, y) ? 1 : 0;
}
/* whole 750x750 maps at once (row-major, 1 byte per pixel) */
void isxref_road_map(int num_lanes, uint8_t *out) {
    RoadGeometry g(num_lanes);
    for (int y = 0; y < HEIGHT; ++y)
        for (int x = 0; x < WIDTH; ++x) out[y * WIDTH + x] = g.is_on_road(float(x), float(y)) ? 1 : 0;
}
void isxref_line_map(int num_lanes, uint8_t *out) {
    LineMask m(num_lanes);
    for (int y = 0; y < HEIGHT; ++y)
        for (int x = 0; x < WIDTH; ++x) out[y * WIDTH + x] = m.is_line(x, y) ? 1 : 0;
}
/* s = {x,y,v,heading,acc,steer} in/out; Car.cpp:9-40 */
void isxref_car_update(float *s, float throttle, float steer, float dt) {
    Car c;
    c.state.x = s[0]; c.state.y = s[1]; c.state.v = s[2]; c.state.heading = s[3]; c.acc = s[4]; c.steering_angle = s[5];
    c.update(throttle, steer, dt);
    s[0] = c.state.x; s[1] = c.state.y; s[2] = c.state.v; s[3] = c.state.heading; s[4] = c.acc; s[5] = c.steering_angle;
}
/* a,b = {x,y,heading}; Car.cpp:105-141 */
int isxref_collide(const float *a, const float *b) {
    Car c1, c2;
    c1.state.x = a[0]; c1.state.y = a[1]; c1.state.heading = a[2];
    c2.state.x = b[0]; c2.state.y = b[1]; c2.state.heading = b[2];
    return c1.check_collision(c2) ? 1 : 0;
}
/* corners of a car {x,y,heading} -> 8 floats; Car.cpp:86-103 */
void isxref_corners(const float *a, float *out8) {
    Car c; c.state.x = a[0]; c.state.y = a[1]; c.state.heading = a[2];
    auto cs = c.corners();
    for (int i = 0; i < 4; ++i) { out8[2 * i] = cs[size_t(i)].first; out8[2 * i + 1] = cs[size_t(i)].second; }
}
/* Lidar.cpp:16-90 on a synthetic scene: self={x,y,heading}, others = n x {x,y,heading}; rays in {72,96} */
void isxref_lidar(int num_lanes, int rays, const float *self_pose, const float *others, int n_others, float *dist) {
    RoadGeometry g(num_lanes);
    Lidar l;
    if (rays != l.rays) {
        l.rays = rays;
        l.distances.assign(size_t(rays), l.max_dist);
        l.rel_angles.clear();
        const float start_angle_deg = -l.fov_deg * 0.5f;
        const float step_deg = (rays > 1) ? (l.fov_deg / float(rays - 1)) : 0.0f;
        constexpr float PI_F2 = 3.14159265358979323846f;
        for (int i = 0; i < rays; ++i) { float deg = start_angle_deg + i * step_deg; l.rel_angles.push_back(deg * PI_F2 / 180.0f); }
    }
    Car self; self.state.x = self_pose[0]; self.state.y = self_pose[1]; self.state.heading = self_pose[2];
    std::vector<Car> cars;
    cars.push_back(self);
    for (int i = 0; i < n_others; ++i) {
        Car c; c.state.x = others[3 * i]; c.state.y = others[3 * i + 1]; c.state.heading = others[3 * i + 2];
        cars.push_back(c);
    }
    l.update(cars[0], cars, g, WIDTH, HEIGHT);
    for (int i = 0; i < rays; ++i) dist[i] = l.distances[size_t(i)];
}

/* libm probes: the exact libm entry points the reference binds (nm -D: sincosf, tanf, atan2f,
 * hypotf, fmodf), evaluated in THIS process so tests can compare device math with them. */
void isxref_libm_sincosf(const float *x, int n, float *s, float *c) { for (int i = 0; i < n; ++i) { s[i] = std::sin(x[i]); c[i] = std::cos(x[i]); } }
void isxref_libm_tanf(const float *x, int n, float *o) { for (int i = 0; i < n; ++i) o[i] = std::tan(x[i]); }
void isxref_libm_atan2f(const float *y, const float *x, int n, float *o) { for (int i = 0; i < n; ++i) o[i] = std::atan2(y[i], x[i]); }
void isxref_libm_hypotf(const float *y, const float *x, int n, float *o) { for (int i = 0; i < n; ++i) o[i] = std::hypot(y[i], x[i]); }
void isxref_libm_fmodf(const float *y, const float *x, int n, float *o) { for (int i = 0; i < n; ++i) o[i] = std::fmod(y[i], x[i]); }

/* The toolchain's std::sort on the element type and comparator of IntersectionEnv.cpp:459-462,490 (a float key and a
 * pointer payload).  perm_out[i] = original list position of the element that ends at rank i.  Used to pin the
 * oracle's and the product's restatement of libstdc++ introsort for lists with exactly-equal distances. */
void isxref_std_sort(const float *keys, int n, int32_t *perm_out) {
    struct Item { float dist; const void *car; };
    std::vector<Item> v;
    static const char base[1] = {0};
    for (int i = 0; i < n; ++i) v.push_back({keys[i], base + i});
    std::sort(v.begin(), v.end(), [](const Item &a, const Item &b) { return a.dist < b.dist; });
    for (int i = 0; i < n; ++i) perm_out[i] = int32_t(static_cast<const char *>(v[size_t(i)].car) - base);
}
/* McIlroy's adversary ("A Killer Adversary for Quicksort", 1999) played against this std::sort: returns keys for which
 * every pivot is nearly the minimum, so that the depth limit is hit and the heap-sort fallback runs. */
void isxref_sort_adversary(int n, float *keys_out) {
    const int gas = n;
    std::vector<int> val(size_t(n), gas), item(static_cast<size_t>(n));
    int nsolid = 0, candidate = 0;
    for (int i = 0; i < n; ++i) item[size_t(i)] = i;
    std::sort(item.begin(), item.end(), [&](int x, int y) {
        if (val[size_t(x)] == gas && val[size_t(y)] == gas) { if (x == candidate) val[size_t(x)] = nsolid++; else val[size_t(y)] = nsolid++; }
        if (val[size_t(x)] == gas) candidate = x; else if (val[size_t(y)] == gas) candidate = y;
        return val[size_t(x)] < val[size_t(y)];
    });
    for (int i = 0; i < n; ++i) keys_out[i] = float(val[size_t(i)]);
}

/* CPU-baseline rollout: `steps` steps with the Philox action stream, reset on terminated|truncated
 * (what a user of env.py does).  Returns agent-steps executed.  Thread-safe per handle. */
long long isxref_rollout(void *h, int steps, float dt, int32_t *status_hist6, double *reward_sum) {
    RefEnv *r = static_cast<RefEnv *>(h);
    const int n = int(r->env.cars.size());
    const size_t nn = size_t(n);
    std::vector<float> th(nn), st(nn), rew(nn);
    std::vector<int32_t> done(nn), stat(nn);
    long long agent_steps = 0;
    for (int s = 0; s < steps; ++s) {
        for (int a = 0; a < n; ++a) isx_action_for(r->seed, r->env_id, r->tick + 1, uint32_t(a), &th[size_t(a)], &st[size_t(a)]);
        int32_t term = 0, trunc = 0, alive = 0;
        isxref_step(h, th.data(), st.data(), n, dt, nullptr, rew.data(), done.data(), stat.data(), &term, &trunc, &alive);
        for (int a = 0; a < n; ++a) {
            if (status_hist6) status_hist6[stat[size_t(a)]]++;
            if (reward_sum) *reward_sum += double(rew[size_t(a)]);
        }
        agent_steps += n;
        if (term || trunc) isxref_reset(h);
    }
    return agent_steps;
}

}  /* extern "C" */
