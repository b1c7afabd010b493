// isx_host_units.cpp — HOST build of the entity-level product arithmetic (isx_sim.cuh, isx_tables.h,
// isx_rng.cuh) behind a C ABI.  Test hook only (tests/test_host_units.py compares it with the reference
// build and the oracle on the CPU, where iteration is cheap); the product path never loads this library
// and has no CPU fallback.  Built with g++ -O2 -ffp-contract=off, no -march.
#include "isx_rng.cuh"
#include "isx_sim.cuh"
#include "isx_tables.h"

#include <map>
#include <mutex>

using namespace isx;

namespace {
std::mutex g_mu;
std::map<int, RoadTables*> g_tables;
const RoadTables* tables_for(int lanes) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_tables.find(lanes);
    if (it == g_tables.end()) {
        RoadTables* t = new RoadTables();
        if (!build_road_tables(lanes, t)) { std::fprintf(stderr, "road map not symmetric\n"); std::abort(); }
        it = g_tables.emplace(lanes, t).first;
    }
    return it->second;
}
}  // namespace

extern "C" {

int isxh_route(int lanes, const char* start, const char* end, float* path_xy, int* intent, float* sx, float* sy, float* sh) {
    RouteHost r;
    const int rc = build_route(lanes, start, end, &r);
    if (rc) return rc;
    for (int i = 0; i < PATH_LEN; ++i) { path_xy[2 * i] = r.path[i].x; path_xy[2 * i + 1] = r.path[i].y; }
    *intent = r.intent; *sx = r.spawn_x; *sy = r.spawn_y; *sh = r.spawn_h;
    return PATH_LEN;
}
int isxh_on_road(int lanes, float x, float y) { return on_road(lanes, x, y) ? 1 : 0; }
int isxh_yellow(int lanes, float x, float y) { return hits_yellow(lanes, x, y) ? 1 : 0; }
int isxh_is_line(int lanes, int x, int y) { return is_line_px(lanes, x, y) ? 1 : 0; }
// unfolded views of the folded tables, for exhaustive comparison with the reference maps
void isxh_road_map(int lanes, uint8_t* out) {
    const RoadTables* t = tables_for(lanes);
    for (int y = 0; y < HEIGHT; ++y) for (int x = 0; x < WIDTH; ++x) out[y * WIDTH + x] = road_bit(t->bits.data(), x, y) ? 1 : 0;
}
void isxh_line_map(int lanes, uint8_t* out) {
    for (int y = 0; y < HEIGHT; ++y) for (int x = 0; x < WIDTH; ++x) out[y * WIDTH + x] = is_line_px(lanes, x, y) ? 1 : 0;
}
void isxh_skip_map(int lanes, uint8_t* out) {
    const RoadTables* t = tables_for(lanes);
    for (int y = 0; y < HEIGHT; ++y) for (int x = 0; x < WIDTH; ++x) out[y * WIDTH + x] = (uint8_t)road_skip(t->skip.data(), x, y);
}
void isxh_car_update(float* s, float thr, float st, float dt) {
    Pose p{s[0], s[1], s[2], s[3]};
    float acc = s[4], steer = s[5];
    car_update(p, steer, acc, thr, st, dt);
    s[0] = p.x; s[1] = p.y; s[2] = p.v; s[3] = p.h; s[4] = acc; s[5] = steer;
}
int isxh_collide(const float* a, const float* b) { return cars_collide(a[0], a[1], a[2], b[0], b[1], b[2]) ? 1 : 0; }
void isxh_corners(const float* a, float* out8) {
    float s, c, cx[4], cy[4];
    sincosf_(a[2], &s, &c);
    car_corners(a[0], a[1], s, c, cx, cy);
    for (int i = 0; i < 4; ++i) { out8[2 * i] = cx[i]; out8[2 * i + 1] = cy[i]; }
}
// The product lidar for one ego: road sphere-trace + per-car slab/verify, as the kernel evaluates it.
void isxh_lidar(int lanes, int rays, const float* self_pose, const float* others, int n_others, float* dist) {
    const RoadTables* t = tables_for(lanes);
    std::vector<float> rel((size_t)rays);
    lidar_rel_angles(rays, rel.data());
    const float cx = self_pose[0], cy = self_pose[1], h = self_pose[2];
    std::vector<PixRect> rects;
    std::vector<BeamWindow> wins;
    for (int i = 0; i < n_others; ++i) {
        const float ox = others[3 * i], oy = others[3 * i + 1], oh = others[3 * i + 2];
        if (fabsf(ox - cx) < 1e-3f && fabsf(oy - cy) < 1e-3f && fabsf(oh - h) < 1e-3f) continue;   // Lidar.cpp:58-63
        rects.push_back(car_pixel_rect(ox, oy, oh));
        wins.push_back(beam_window(rects.back(), cx, cy, h, rays));
    }
    for (int i = 0; i < rays; ++i) {
        float s, c;
        sincosf_(h + rel[i], &s, &c);
        const Ray ray = make_ray(cx, cy, c, -s);
        bool hit;
        const int ke = ray_road_event_ana(t->ana, t->bits.data(), ray, &hit);   // as k_lidar_obs marches (the skip-table march: isxh_road_events)
        int best = hit ? ke : 0;
        const int kmax = hit ? ke - 1 : ke - 1;     // cars only count strictly before the road event
        if (kmax >= 1) {
            for (size_t q = 0; q < rects.size(); ++q) {
                const PixRect& r = rects[q];
                const int lim = best ? best - 1 : kmax;
                if (lim < 1) break;
                if (wins[q].kmin > lim) continue;                     // range pruning, as the kernel does
                if (!beam_in_window(wins[q], i, rays)) continue;      // angular pruning, as the kernel does
                const int k = ray_rect_first_hit(r, ray, lim);
                if (k) best = k;
            }
        }
        dist[i] = best ? (float)(4 * best) : LIDAR_MAX_DIST;
    }
}
// The angular beam window k_features computes for a car seen from an ego (ia, span; span 255 = all beams).
void isxh_beam_window(const float* car_pose, float cx, float cy, float heading, int rays, int* ia, int* span) {
    const BeamWindow w = beam_window(car_pixel_rect(car_pose[0], car_pose[1], car_pose[2]), cx, cy, heading, rays);
    *ia = w.ia; *span = w.span;
}
// Tuning probe: per beam of one ego, the march state after `lockstep` accelerated steps (k, or -1 when already done)
// and the final road event index — what the cooperative tail of k_lidar_obs has left to do.
void isxh_march_stats(int lanes, int rays, int lockstep, const float* self_pose, int* k_after, int* k_event) {
    const RoadTables* t = tables_for(lanes);
    std::vector<float> rel((size_t)rays);
    lidar_rel_angles(rays, rel.data());
    const RoadView rv{t->bits.data(), t->skip.data(), t->box_lo, t->box_hi};
    for (int i = 0; i < rays; ++i) {
        float s, c;
        sincosf_(self_pose[2] + rel[i], &s, &c);
        const Ray ray = make_ray(self_pose[0], self_pose[1], c, -s);
        March m;
        march_init(ray, m);
        for (int it = 0; it < lockstep; ++it) if (!m.done) march_step(rv, ray, m);
        k_after[i] = m.done ? -1 : m.k;
        bool hit;
        k_event[i] = ray_road_event(rv, ray, &hit);
    }
}
// Ray-level comparison of the two road marches on n rays (origin + absolute beam angle, as Lidar.cpp:24-26 forms them):
// skip-table / strip-box march (ray_road_event) against analytic jump + exact samples (ray_road_event_ana).  out6 per ray:
// ke_old, hit_old, ke_new, hit_new, exact samples tested after the jump, K of the jump.
void isxh_road_events(int lanes, int n, const float* cx, const float* cy, const float* angle, int* out6) {
    const RoadTables* t = tables_for(lanes);
    const RoadView rv{t->bits.data(), t->skip.data(), t->box_lo, t->box_hi};
    for (int i = 0; i < n; ++i) {
        float s, c;
        sincosf_(angle[i], &s, &c);
        const Ray ray = make_ray(cx[i], cy[i], c, -s);
        bool h0, h1;
        int tests = 0;
        const int k0 = ray_road_event(rv, ray, &h0);
        const int k1 = ray_road_event_ana(t->ana, t->bits.data(), ray, &h1, &tests);
        int* o = out6 + 6 * (size_t)i;
        o[0] = k0; o[1] = h0; o[2] = k1; o[3] = h1; o[4] = tests; o[5] = ray_safe_samples(t->ana, ray);
    }
}
// Car::update_path_index with the near/far shortcut against the plain 50-point scan, on n (index, position) samples of one route.
int isxh_path_index(int lanes, const char* start, const char* end, int n, const int* idx, const float* x, const float* y, int* fast, int* full) {
    RouteHost r;
    const int rc = build_route(lanes, start, end, &r);
    if (rc) return rc;
    float far2[PATH_LEN];
    path_far_table(r.path, far2);
    for (int i = 0; i < n; ++i) {
        fast[i] = path_index_update(r.path, far2, idx[i], x[i], y[i]);
        full[i] = path_index_update(r.path, nullptr, idx[i], x[i], y[i]);
    }
    return 0;
}
int isxh_ana_enabled(int lanes) { return tables_for(lanes)->ana.enabled; }
int isxh_self_status(int lanes, float x, float y, float h, float gx, float gy, float px, float py) {
    return ego_self_status(lanes, x, y, h, F2{gx, gy}, F2{px, py});
}
void isxh_actions(uint64_t seed, uint32_t env, uint32_t tick, int n, float* out) {
    for (int a = 0; a < n; ++a) philox_action(seed, env, tick, (uint32_t)a, out[2 * a], out[2 * a + 1]);
}
// first `n` words of a traffic stream + the u01 / below(12) views of a fresh stream
void isxh_traffic_words(uint64_t seed, uint32_t env, uint32_t tick, int n, uint32_t* out) {
    TrafficStream s; s.init(seed, env, tick);
    for (int i = 0; i < n; ++i) out[i] = s.next();
}
// neighbour ordering (isx_sim.cuh stdsort): perm_out[i] = list position ranked i; returns the heap-fallback count
int isxh_std_sort(const float* keys, int n, int32_t* perm_out) {
    if (n > 255) return -1;
    uint8_t p[256];
    for (int i = 0; i < n; ++i) p[i] = (uint8_t)i;
    const int heaps = stdsort::sort(keys, p, n);
    for (int i = 0; i < n; ++i) perm_out[i] = p[i];
    return heaps;
}
float isxh_u01_of(uint64_t seed, uint32_t env, uint32_t tick) { TrafficStream s; s.init(seed, env, tick); return s.uniform01(); }

}  // extern "C"
