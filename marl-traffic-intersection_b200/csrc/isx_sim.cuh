// isx_sim.cuh — entity-level simulation arithmetic, __host__ __device__.
//
// Every function here is the float32 arithmetic of one reference function, written so that the SAME
// sequence of IEEE operations happens (no contraction: build with -fmad=false / -ffp-contract=off) and
// with the libm calls replaced by the bit-identical restatements of isx_math.cuh.  The kernels
// (isx_kernels.cu) decide which thread evaluates what; nothing here knows about warps.
// The host build is exercised against the reference in tests/test_host_units.py.
#pragma once
#include "isx_math.cuh"

namespace isx {

// constants.h:4-20
constexpr int WIDTH = 750;
constexpr int HEIGHT = 750;
constexpr float CAR_LENGTH = 54.0f;
constexpr float CAR_WIDTH = 24.0f;
constexpr float LANE_WIDTH_PX = 42.0f;
constexpr float CORNER_RADIUS = 84.0f;
constexpr float MAX_ACC = 15.0f;
constexpr float MAX_STEERING_ANGLE = 0.6108652381980153f;
constexpr float PHYSICS_MAX_SPEED = 8.0f;
constexpr float FPS = 60.0f;
constexpr float SCALE = 12.0f;
constexpr int PATH_LEN = 160;
constexpr int LIDAR_MAX_K = 62;          // samples at 4,8,...,248 px (Lidar.cpp:33: dist < 250)
constexpr float LIDAR_STEP = 4.0f;
constexpr float LIDAR_MAX_DIST = 250.0f;

struct Pose { float x, y, v, h; };
struct alignas(8) F2 { float x, y; };   // way-point / goal (8-byte vector load on the device)

// ---------------------------------------------------------------- road geometry
// RoadGeometry::is_on_road (RoadGeometry.h:19-58): (two strips U four corner squares) \ four grass discs.
ISX_HD bool on_road(int lanes, float x, float y) {
    const float C = WIDTH * 0.5f;
    const float rw = (float)lanes * LANE_WIDTH_PX;
    const float cr = CORNER_RADIUS;
    const float lo = C - rw - cr, hi = C + rw + cr;     // grass-disc centres, also outer edge of the squares
    const float r2 = cr * cr;
    const float dxl = x - lo, dxh = x - hi, dyl = y - lo, dyh = y - hi;
    if (dxl * dxl + dyl * dyl <= r2) return false;
    if (dxh * dxh + dyl * dyl <= r2) return false;
    if (dxl * dxl + dyh * dyh <= r2) return false;
    if (dxh * dxh + dyh * dyh <= r2) return false;
    const float a = C - rw, b = C + rw;
    const bool in_v = (x >= a) && (x <= b);
    const bool in_h = (y >= a) && (y <= b);
    if (in_v || in_h) return true;
    const bool xl = (x >= lo) && (x <= a), xr = (x >= b) && (x <= hi);
    const bool yt = (y >= lo) && (y <= a), yb = (y >= b) && (y <= hi);
    return (xl || xr) && (yt || yb);
}

// RoadGeometry::hits_yellow_line (RoadGeometry.h:60-67)
ISX_HD bool hits_yellow(int lanes, float x, float y) {
    const float C = WIDTH * 0.5f;
    const float rw = (float)lanes * LANE_WIDTH_PX;
    const float ax = fabsf(x - C), ay = fabsf(y - C);
    return (ax <= 2.0f && ay > rw) || (ay <= 2.0f && ax > rw);
}

// LineMask (LineMask.cpp:47-72, LineMask.h:15-18) in closed form: double lines at 373/377 +-1 px, drawn
// from each screen edge to the stop line at centre -+ (lanes*42 + 84).
ISX_HD bool is_line_px(int lanes, int x, int y) {
    if ((unsigned)x >= (unsigned)WIDTH || (unsigned)y >= (unsigned)HEIGHT) return false;
    const int c = WIDTH / 2;
    const int stop = lanes * (int)LANE_WIDTH_PX + (int)CORNER_RADIUS;
    const int ux = x - c, uy = y - c;
    const int ax = ux < 0 ? -ux : ux, ay = uy < 0 ? -uy : uy;
    const bool band_x = (ax >= 1 && ax <= 3);      // columns 372..374, 376..378
    const bool band_y = (ay >= 1 && ay <= 3);
    const bool far_y = (uy <= -stop) || (uy >= stop);
    const bool far_x = (ux <= -stop) || (ux >= stop);
    return (band_x && far_y) || (band_y && far_x);
}

// ---------------------------------------------------------------- car
// Car::update (Car.cpp:9-40).  Kinematic bicycle; NOTE the pose step has no dt (px per frame).
// Split in two so that k_traffic can evaluate the throttle-independent half (steering low-pass + its tangent) for all
// NPCs in parallel before the sequential planner loop; car_update() below is the plain composition of the halves.
ISX_HD float car_steer_update(float steer, float steer_in) {                // Car.cpp:14-15
    const float target = steer_in * MAX_STEERING_ANGLE;
    return steer + (target - steer) * 0.2f;
}
// tan_steer = tanf(new steering angle); only used when |v| > 0.1 (Car.cpp:27-30)
// (sin_h, cos_h): sine / cosine of the NEW heading, for callers that need them again (status corners, SAT, pixel rectangle)
ISX_HD void car_motion_update(Pose& p, float& acc, float throttle, float tan_steer, float dt, float* sin_h = nullptr, float* cos_h = nullptr) {   // Car.cpp:12,17-39
    acc = throttle * MAX_ACC;
    float v = p.v;
    if (throttle == 0.0f) v = v * 0.95f;
    v = v + acc * dt;
    if (v < 0.0f) v = 0.0f;
    if (v > PHYSICS_MAX_SPEED) v = PHYSICS_MAX_SPEED;
    float h = p.h;
    if (fabsf(v) > 0.1f) {
        const float yaw = (v / CAR_LENGTH) * tan_steer;
        h = h + yaw;
    }
    h = fmodf_(h + PI_F, TWO_PI_F);
    if (h < 0.0f) h = h + TWO_PI_F;
    h = h - PI_F;
    float s, c;
    sincosf_nc(h, &s, &c);
    p.x = p.x + v * c;
    p.y = p.y - v * s;
    p.v = v;
    p.h = h;
    if (sin_h) *sin_h = s;
    if (cos_h) *cos_h = c;
}
ISX_HD_NOINL void car_update(Pose& p, float& steer, float& acc, float throttle, float steer_in, float dt) {
    steer = car_steer_update(steer, steer_in);
    car_motion_update(p, acc, throttle, tanf_nc(steer), dt);
}
ISX_HD void car_update_sc(Pose& p, float& steer, float& acc, float throttle, float steer_in, float dt, float* sin_h, float* cos_h) {
    steer = car_steer_update(steer, steer_in);
    car_motion_update(p, acc, throttle, tanf_nc(steer), dt, sin_h, cos_h);
}

// Car::corners (Car.cpp:86-103): (+-27, +-12) rotated by +heading (no y flip) — order FL, FR, RR, RL.
ISX_HD void car_corners(float x, float y, float s, float c, float cx[4], float cy[4]) {
    const float hl = CAR_LENGTH * 0.5f, hw = CAR_WIDTH * 0.5f;
    cx[0] = x + hl * c - hw * s;        cy[0] = y + hl * s + hw * c;
    cx[1] = x + hl * c - (-hw) * s;     cy[1] = y + hl * s + (-hw) * c;
    cx[2] = x + (-hl) * c - (-hw) * s;  cy[2] = y + (-hl) * s + (-hw) * c;
    cx[3] = x + (-hl) * c - hw * s;     cy[3] = y + (-hl) * s + hw * c;
}

// Car::check_collision (Car.cpp:105-141): separating-axis test over the 2+2 edge directions, touching = hit.
ISX_HD bool sat_overlap(const float ax[4], const float ay[4], float as, float ac,
                        const float bx[4], const float by[4], float bs, float bc) {
    const float ux[4] = {ac, -as, bc, -bs};
    const float uy[4] = {as, ac, bs, bc};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float lo1 = ax[0] * ux[k] + ay[0] * uy[k], hi1 = lo1;
        float lo2 = bx[0] * ux[k] + by[0] * uy[k], hi2 = lo2;
#pragma unroll
        for (int i = 1; i < 4; ++i) {
            const float p = ax[i] * ux[k] + ay[i] * uy[k];
            lo1 = fminf(lo1, p); hi1 = fmaxf(hi1, p);
            const float q = bx[i] * ux[k] + by[i] * uy[k];
            lo2 = fminf(lo2, q); hi2 = fmaxf(hi2, q);
        }
        if (hi1 < lo2 || hi2 < lo1) return false;
    }
    return true;
}

// Conservative reject before the SAT: two 54x24 rectangles whose centres are > 70 px apart are
// separated by > 10 px, and then one of the four SAT axes shows a gap of > 7 px — far above float
// rounding — so the exact SAT would also answer "no".  Exactness-preserving by construction.
ISX_HD bool cars_far_apart(float x1, float y1, float x2, float y2) {
    const float dx = x1 - x2, dy = y1 - y2;
    return dx * dx + dy * dy > 4900.0f;
}

// the same with the sines / cosines of both headings supplied (each car's heading is evaluated once per step)
ISX_HD bool cars_collide_sc(float x1, float y1, float s1, float c1, float x2, float y2, float s2, float c2) {
    if (cars_far_apart(x1, y1, x2, y2)) return false;
    float ax[4], ay[4], bx[4], by[4];
    car_corners(x1, y1, s1, c1, ax, ay);
    car_corners(x2, y2, s2, c2, bx, by);
    return sat_overlap(ax, ay, s1, c1, bx, by, s2, c2);
}
ISX_HD bool cars_collide(float x1, float y1, float h1, float x2, float y2, float h2) {
    if (cars_far_apart(x1, y1, x2, y2)) return false;
    float s1, c1, s2, c2;
    sincosf_nc(h1, &s1, &c1);
    sincosf_nc(h2, &s2, &c2);
    float ax[4], ay[4], bx[4], by[4];
    car_corners(x1, y1, s1, c1, ax, ay);
    car_corners(x2, y2, s2, c2, bx, by);
    return sat_overlap(ax, ay, s1, c1, bx, by, s2, c2);
}

// Car::update_path_index (Car.cpp:47-74): first minimum of squared distance over path[idx, min(idx+50,160)).
// Written as fixed iterations over a clamped index so that the loads are independent of the running minimum and
// can be issued in batches (the window tail repeats point 159, which can never win a strict `<` against itself).
// Exactness-preserving shortcut: the window is scanned in two parts, the first PATH_NEAR points and the rest.  far2[start]
// (path_far_table below, built on the host) is (r - 0.01)^2 with r the smallest distance from path[start] to any point of
// the far part; a far point P_j is at least r - |car - path[start]| away from the car, so when
// 2 (d_start + best_near) < far2[start]  —  which implies  r - 0.01 > sqrt(d_start) + sqrt(best_near)  —  every far point
// is farther than the near minimum by more than 0.01 px (rounding of the squared distances is ~1e-7 relative) and cannot
// win the strict `<`: the far part is skipped.  far2 == nullptr scans everything.
constexpr int PATH_WINDOW = 50, PATH_NEAR = 16;
ISX_HD int path_index_update(const F2* path, const float* far2, int idx, float x, float y) {
    const int start = idx < 0 ? 0 : idx;
    float best = INFINITY, d0 = 0.0f;
    int bi = start;
#pragma unroll
    for (int j = 0; j < PATH_NEAR; ++j) {
        const int i = (start + j < PATH_LEN) ? start + j : PATH_LEN - 1;
        const F2 p = path[i];
        const float dx = p.x - x, dy = p.y - y;
        const float d = dx * dx + dy * dy;
        if (j == 0) d0 = d;
        if (d < best) { best = d; bi = i; }
    }
    if (far2 && 2.0f * (d0 + best) < far2[start < PATH_LEN ? start : PATH_LEN - 1]) return bi;
#pragma unroll 17
    for (int j = PATH_NEAR; j < PATH_WINDOW; ++j) {
        const int i = (start + j < PATH_LEN) ? start + j : PATH_LEN - 1;
        const F2 p = path[i];
        const float dx = p.x - x, dy = p.y - y;
        const float d = dx * dx + dy * dy;
        if (d < best) { best = d; bi = i; }
    }
    return bi;
}
// The near part alone (the kernels finish the far part warp-cooperatively for the few lanes that need it).
ISX_HD void path_index_near(const F2* path, int start, float x, float y, float& best, int& bi, float& d0) {
    best = INFINITY; bi = start; d0 = 0.0f;
#pragma unroll
    for (int j = 0; j < PATH_NEAR; ++j) {
        const int i = (start + j < PATH_LEN) ? start + j : PATH_LEN - 1;
        const F2 p = path[i];
        const float dx = p.x - x, dy = p.y - y;
        const float d = dx * dx + dy * dy;
        if (j == 0) d0 = d;
        if (d < best) { best = d; bi = i; }
    }
}
// far2[i] of one 160-point path (host side; double arithmetic, rounded down)
inline void path_far_table(const F2* path, float* far2) {
    for (int i = 0; i < PATH_LEN; ++i) {
        double r = INFINITY;
        for (int j = i + PATH_NEAR; j < i + PATH_WINDOW && j < PATH_LEN; ++j) {
            const double dx = (double)path[j].x - path[i].x, dy = (double)path[j].y - path[i].y;
            r = fmin(r, sqrt(dx * dx + dy * dy));
        }
        const double m = r - 0.01;
        far2[i] = (r == INFINITY) ? INFINITY : (m > 0.0 ? (float)(m * m * (1.0 - 1e-6)) : 0.0f);
    }
}

// ---------------------------------------------------------------- ego status (IntersectionEnv.cpp:166-290)
// goal = path[159], prev = path[158].  Returns ISX status code of the car on its own (before car-car).
ISX_HD int ego_self_status_sc(int lanes, float x, float y, float s, float c, F2 goal, F2 prev);
ISX_HD_NOINL int ego_self_status(int lanes, float x, float y, float h, F2 goal, F2 prev) {
    float s, c;
    sincosf_nc(h, &s, &c);
    return ego_self_status_sc(lanes, x, y, s, c, goal, prev);
}
// (s, c) = sine / cosine of the heading; they are only used when the car is not in its success zone, as in the reference
ISX_HD int ego_self_status_sc(int lanes, float x, float y, float s, float c, F2 goal, F2 prev) {
    const float dxr = goal.x - prev.x, dyr = goal.y - prev.y;
    bool ok;
    if (fabsf(dxr) > fabsf(dyr)) ok = (fabsf(y - goal.y) < 15.0f) && (fabsf(x - goal.x) < 40.0f);
    else                         ok = (fabsf(x - goal.x) < 15.0f) && (fabsf(y - goal.y) < 40.0f);
    if (ok) return 2;  // SUCCESS
    float cx[4], cy[4];
    car_corners(x, y, s, c, cx, cy);
    const float lo = -100.0f, hi = (float)WIDTH + 100.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (cx[k] < lo || cx[k] > hi || cy[k] < lo || cy[k] > hi) return 3;  // CRASH_WALL (left the screen)
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (!on_road(lanes, cx[k], cy[k])) return 3;                          // CRASH_WALL (off road)
    bool line = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) line = line || hits_yellow(lanes, cx[k], cy[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int k2 = (k + 1) & 3;
        const float mx = 0.5f * (cx[k] + cx[k2]), my = 0.5f * (cy[k] + cy[k2]);
        line = line || is_line_px(lanes, f2i_rz(mx), f2i_rz(my));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) line = line || is_line_px(lanes, f2i_rz(cx[k]), f2i_rz(cy[k]));
    return line ? 4 : 0;  // CRASH_LINE / ALIVE
}

// ---------------------------------------------------------------- reward pieces (IntersectionEnv.cpp:15-46)
struct RewardCfg { float k_prog, v_min_ms, k_stuck, k_cv, k_co, k_succ, k_sm, alpha; };

ISX_HD float reward_base(const RewardCfg& rc, float x, float y, float v, float acc, float steer, F2 goal,
                         float max_progress, float& prev_dist, float& pa0, float& pa1) {
    const float cur = hypotf_nc(x - goal.x, y - goal.y);
    float r_prog = 0.0f;
    if (prev_dist > 0.0f) {
        const float progress = prev_dist - cur;
        const float norm = (max_progress > 0.0f) ? (progress / max_progress) : 0.0f;
        r_prog = rc.k_prog * norm;
    }
    prev_dist = cur;
    const float speed_ms = (v * FPS) / SCALE;
    const float r_stuck = (speed_ms < rc.v_min_ms) ? rc.k_stuck : 0.0f;
    const float an = acc / MAX_ACC, sn = steer / MAX_STEERING_ANGLE;
    const float d0 = an - pa0, d1 = sn - pa1;
    const float r_smooth = rc.k_sm * (d0 * d0 + d1 * d1);
    pa0 = an; pa1 = sn;
    return r_prog + r_stuck + r_smooth;
}

// ---------------------------------------------------------------- lidar (Lidar.cpp:16-90)
// Integer pixel rectangle equivalent to the float AABB test of Lidar.cpp:65-78:
//   float(px) >= c.x - ex  &&  float(px) <= c.x + ex   <=>   ceil(c.x - ex) <= px <= floor(c.x + ex)
// The record also carries what every beam test needs from it (ray_rect_first_hit): the rectangle CLAMPED to the screen —
// only on-screen pixels can be hit, the march breaks off screen first — and the slab bounds of the real sample positions
// that truncate into it (0.01 px of slack; truncation toward zero maps every value in (-1, 1) to pixel 0, hence the wider
// lower bound when the clamped edge is 0).  A rectangle that is empty after clamping gets bounds no ray satisfies.
struct alignas(16) PixRect { int x0, x1, y0, y1; float lox, hix, loy, hiy; };
ISX_HD PixRect make_pix_rect(int x0, int x1, int y0, int y1) {
    PixRect r;
    r.x0 = x0 < 0 ? 0 : x0; r.x1 = x1 > WIDTH - 1 ? WIDTH - 1 : x1;
    r.y0 = y0 < 0 ? 0 : y0; r.y1 = y1 > HEIGHT - 1 ? HEIGHT - 1 : y1;
    if (r.x0 > r.x1 || r.y0 > r.y1) { r.lox = r.hix = r.loy = r.hiy = 1e30f; return r; }
    r.lox = r.x0 == 0 ? -1.01f : (float)r.x0 - 0.01f; r.hix = (float)r.x1 + 1.01f;
    r.loy = r.y0 == 0 ? -1.01f : (float)r.y0 - 0.01f; r.hiy = (float)r.y1 + 1.01f;
    return r;
}
ISX_HD bool pix_rect_empty(const PixRect& r) { return r.x0 > r.x1 || r.y0 > r.y1; }
ISX_HD PixRect car_pixel_rect_sc(float x, float y, float s, float c) {
    const float hl = CAR_LENGTH * 0.5f, hw = CAR_WIDTH * 0.5f;
    const float ex = fabsf(c) * hl + fabsf(s) * hw;
    const float ey = fabsf(s) * hl + fabsf(c) * hw;
    return make_pix_rect((int)ceilf(x - ex), (int)floorf(x + ex), (int)ceilf(y - ey), (int)floorf(y + ey));
}
ISX_HD PixRect car_pixel_rect(float x, float y, float h) {
    float s, c;
    sincosf_nc(h, &s, &c);
    const float hl = CAR_LENGTH * 0.5f, hw = CAR_WIDTH * 0.5f;
    const float ex = fabsf(c) * hl + fabsf(s) * hw;
    const float ey = fabsf(s) * hl + fabsf(c) * hw;
    return make_pix_rect((int)ceilf(x - ex), (int)floorf(x + ex), (int)ceilf(y - ey), (int)floorf(y + ey));
}

// Pixel of sample k on a ray (Lidar.cpp:34-35): mul and add rounded separately, truncation toward 0.
ISX_HD void ray_pixel(float cx, float cy, float dx, float dy, int k, int& px, int& py) {
    const float dist = (float)(4 * k);
    px = f2i_rz(cx + dx * dist);
    py = f2i_rz(cy + dy * dist);
}

// Folded road tables (built on the host from on_road(), see isx_tables.cpp):
//   bits : (HALF+1) rows x ROAD_WORDS u32, bit (u,v) = on_road(375+-u, 375+-v)  (the map is mirror-symmetric)
//   skip : (SKIP_DIM x SKIP_DIM) u8 per 4x4 block of (u,v): samples that can be skipped for sure
constexpr int ROAD_HALF = 375;               // u,v in [0,375]
constexpr int ROAD_WORDS = 12;               // 376 bits -> 12 words (18 KB per CTA of k_lidar_obs)
constexpr int ROAD_ROWS = ROAD_HALF + 1;
constexpr int SKIP_DIM = 94;                 // ceil(376/4)

ISX_HD bool road_bit(const uint32_t* bits, int px, int py) {
    int u = px - ROAD_HALF; u = u < 0 ? -u : u;
    int v = py - ROAD_HALF; v = v < 0 ? -v : v;
    return (bits[v * ROAD_WORDS + (u >> 5)] >> (u & 31)) & 1u;
}
ISX_HD int road_skip(const uint8_t* skip, int px, int py) {
    int u = px - ROAD_HALF; u = u < 0 ? -u : u;
    int v = py - ROAD_HALF; v = v < 0 ? -v : v;
    return skip[(v >> 2) * SKIP_DIM + (u >> 2)];
}

#if defined(ISX_ITER_STATS) && !defined(__CUDA_ARCH__)
static long long isx_iter_stats_count = 0;   // host-only instrumentation (tools/): march iterations
#endif
ISX_HD float approx_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
// Per-ray constants of the march: direction, reciprocals (for the analytic strip exits and the car slabs).
struct Ray {
    float cx, cy, dx, dy, inv_dx, inv_dy;
};
ISX_HD Ray make_ray(float cx, float cy, float dx, float dy) {
    Ray r;
    r.cx = cx; r.cy = cy; r.dx = dx; r.dy = dy;
    // only used for conservative bounds (0.01 px of slack against ~1e-7 relative error): approximate is fine.  A component
    // of (almost) zero gets a finite reciprocal of its sign (the coordinate then moves < 3e-4 px over the whole ray).
    r.inv_dx = (fabsf(dx) > 1e-6f) ? approx_rcp(dx) : copysignf(1e6f, dx);
    r.inv_dy = (fabsf(dy) > 1e-6f) ? approx_rcp(dy) : copysignf(1e6f, dy);
    return r;
}

// Largest t >= 0 for which  c + d*t  stays inside [lo + 0.01, hi + 0.99]  (pixel in [lo, hi] with 0.01 px slack),
// given that it is inside at the current sample.  +inf if the coordinate does not move.
ISX_HD float axis_exit(float c, float d, float inv_d, int lo, int hi) {
    if (d > 0.0f) return (((float)hi + 0.99f) - c) * inv_d;
    if (d < 0.0f) return (((float)lo + 0.01f) - c) * inv_d;
    return INFINITY;
}

// ---- road march.  (march_step / ray_road_event below are the round-1 skip-table + strip-box march: no kernel uses them any
// more; the host build keeps them as the independent reference the analytic march is validated against — isxh_road_events,
// tests/test_host_units.py.)  State of one ray: the last sample known NOT to be an event (k, px, py), or the result.
//   done && hit   : off-road pixel at sample ke  (Lidar.cpp:44-48)
//   done && !hit  : left the screen at sample ke (:38-40), or ke == 63: nothing within range
//   ke == 0       : the origin pixel itself is off screen (the reference breaks at dist 0)
struct March {
    int k, px, py, ke;
    bool done, hit;
};
struct RoadView {
    const uint32_t* bits;
    const uint8_t* skip;
    int box_lo, box_hi;
};

ISX_HD void march_init(const Ray& r, March& m) {
    m.k = 0; m.ke = LIDAR_MAX_K + 1; m.done = false; m.hit = false;
    ray_pixel(r.cx, r.cy, r.dx, r.dy, 0, m.px, m.py);
    if ((unsigned)m.px >= (unsigned)WIDTH || (unsigned)m.py >= (unsigned)HEIGHT) { m.ke = 0; m.done = true; }
}

// Exact test of sample k: 0 = nothing, 1 = off screen (march breaks, no hit), 2 = off-road pixel (hit).
// `bits`: where the folded bitmap lives — a plain pointer, or (device) a shared-memory address read with ld.shared, which
// spares every test the generic-to-shared window arithmetic a plain pointer into shared memory costs.
struct RoadBitsPtr {
    const uint32_t* p;
    ISX_HDM uint32_t word(int i) const { return p[i]; }
};
#if defined(__CUDACC__)
struct RoadBitsShared {
    uint32_t addr;                                   // __cvta_generic_to_shared of the table
    __device__ __forceinline__ uint32_t word(int i) const {
        uint32_t w;
        asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr + 4u * (uint32_t)i));
        return w;
    }
};
#endif
ISX_HD int iabs_(int x) { return x < 0 ? -x : x; }
template <class B>
ISX_HD int sample_event(const B& bits, const Ray& r, int k, int& px, int& py) {
    ray_pixel(r.cx, r.cy, r.dx, r.dy, k, px, py);
    if ((unsigned)px >= (unsigned)WIDTH || (unsigned)py >= (unsigned)HEIGHT) return 1;
#if defined(__CUDA_ARCH__)
    const int u = abs(px - ROAD_HALF), v = abs(py - ROAD_HALF);
#else
    const int u = iabs_(px - ROAD_HALF), v = iabs_(py - ROAD_HALF);
#endif
    return ((bits.word(v * ROAD_WORDS + (u >> 5)) >> (u & 31)) & 1u) ? 0 : 2;
}
ISX_HD int sample_event(const uint32_t* bits, const Ray& r, int k, int& px, int& py) {
    return sample_event(RoadBitsPtr{bits}, r, k, px, py);
}

// One accelerated step from a non-event sample.  Exactness-preserving: two sources of "these samples cannot be
// events", both verified against the full-resolution map when the tables are built (isx_tables.h):
//   * skip table: a sample whose 4x4 block has skip count j guarantees the next j samples are on-road, on-screen;
//   * strip boxes: every pixel of the open vertical strip [box_lo, box_hi] x [0,749] (and of the horizontal one)
//     is road, so while a ray stays inside a strip — a linear condition in k — nothing can happen.
// The sample that ends the skip is tested with the exact pixel arithmetic of Lidar.cpp:34-46.
ISX_HD void march_step(const RoadView& rv, const Ray& r, March& m) {
#if defined(ISX_ITER_STATS) && !defined(__CUDA_ARCH__)
    ++isx_iter_stats_count;
#endif
    int u = m.px - ROAD_HALF; u = u < 0 ? -u : u;
    int v = m.py - ROAD_HALF; v = v < 0 ? -v : v;
    int j = rv.skip[(v >> 2) * SKIP_DIM + (u >> 2)];
    const bool in_v = (m.px >= rv.box_lo) && (m.px <= rv.box_hi);
    const bool in_h = (m.py >= rv.box_lo) && (m.py <= rv.box_hi);
    if (in_v | in_h) {
        // exit time of the strip the sample is in: strip walls on one axis, the screen on the other
        const float tv = in_v ? fminf(axis_exit(r.cx, r.dx, r.inv_dx, rv.box_lo, rv.box_hi), axis_exit(r.cy, r.dy, r.inv_dy, 0, HEIGHT - 1)) : 0.0f;
        const float th = in_h ? fminf(axis_exit(r.cy, r.dy, r.inv_dy, rv.box_lo, rv.box_hi), axis_exit(r.cx, r.dx, r.inv_dx, 0, WIDTH - 1)) : 0.0f;
        const float t = fminf(fmaxf(tv, th), 1000.0f);
        const int kb = (int)(t * 0.25f - 0.01f);          // last sample index certainly inside the strip
        j = (kb - m.k > j) ? (kb - m.k) : j;
    }
    m.k += j + 1;
    if (m.k > LIDAR_MAX_K) { m.ke = LIDAR_MAX_K + 1; m.done = true; return; }
    const int e = sample_event(rv.bits, r, m.k, m.px, m.py);
    if (e) { m.ke = m.k; m.hit = (e == 2); m.done = true; }
}

// First road event on a ray (plain sequential form; the kernel runs two lock-step steps per lane and then
// finishes the few stragglers warp-cooperatively, see isx_kernels.cu).
ISX_HD int ray_road_event(const RoadView& rv, const Ray& r, bool* hit) {
    March m;
    march_init(r, m);
    while (!m.done) march_step(rv, r, m);
    *hit = m.hit;
    return m.ke;
}

// ---- analytic lower bound of the first road event (exactness-preserving; replaces the strip boxes and the first skip steps)
// RoadGeometry::is_on_road is "two strips U four corner squares minus four grass discs": its complement is, per screen
// quadrant, the Minkowski sum of the quarter plane {sx(x-C) >= U, sy(y-C) >= U} (U = lanes*42 + 84) with a disc of radius
// 84 — two straight walls joined by a quarter circle.  An off-road PIXEL therefore lies within 84 of one of the four
// quarter planes (checked pixel by pixel when the tables are built, isx_tables.h), and the real sample position lies within
// (1,1) of its pixel, so no sample whose real position is farther than rho = 84 + 1.5 from all four quarter planes can be a
// road event.  ray_safe_samples returns K such that samples 1..K are certainly not road events; the samples after K are
// then tested with the exact pixel arithmetic, one by one.  Off-screen samples need no bound: a ray that leaves the screen
// never comes back, reports "no hit" (Lidar.cpp:38-40), and cars are only ever tested against on-screen pixels.
//   x side: tR / tL = time at which the ray is inside the half plane a >= g / a <= -g (a = x - C, g = U - rho), inf if never;
//   same for y; the blob (quadrant) entered first is (earlier x side, earlier y side) at ts = max of the two times —
//   always a lower bound of the true entry.  If the ray is, at ts, in the notch between the two walls and the quarter
//   circle, the bound is refined with the circle (one square root); the other three blobs are bounded by their slab times.
// Only bounds: approximate reciprocal / square root and explicit FMAs are fine (0.08 px of slack against ~1e-4 px of error).
struct RoadAna { float g, U, rho2; int enabled; };
ISX_HD float approx_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
// Per axis, with af = the coordinate along the ray's own direction of travel: the wall ahead (af >= g) is reached at
// (g - af) / |d| (0 if already behind it); the wall behind (-af >= g) only counts if the origin is already past it.
ISX_HD void slab_times(float a, float d, float inv_d, float g, float& t_first, float& t_other, bool& first_is_positive_side) {
    const bool fwd = d > 0.0f;
    const float af = fwd ? a : -a;
    const float t_f = fmaxf((g - af) * fabsf(inv_d), 0.0f);      // d == 0: |inv_d| = 1e6, i.e. never unless already there
    const float t_b = (-af >= g) ? 0.0f : INFINITY;
    const bool f_first = t_f <= t_b;
    t_first = f_first ? t_f : t_b; t_other = f_first ? t_b : t_f;
    first_is_positive_side = f_first == fwd;
}
ISX_HD int ray_safe_samples(const RoadAna& ra, const Ray& r) {
    const float a = r.cx - (float)ROAD_HALF, b = r.cy - (float)ROAD_HALF;
    float tx, txo, ty, tyo;
    bool right, down;
    slab_times(a, r.dx, r.inv_dx, ra.g, tx, txo, right);
    slab_times(b, r.dy, r.inv_dy, ra.g, ty, tyo, down);
    const float ts = fmaxf(tx, ty);
    const float t2 = fminf(fmaxf(txo, ty), fmaxf(tx, tyo));                // earliest slab time of the other three blobs
    // straight-line on purpose: nearly every warp has a lane in the notch, so a branch would only add its own overhead
    // (ts = inf makes p1 / p2 inf or NaN; every comparison with them is false and the refinement is not selected)
    const float w1 = (right ? a : -a) - ra.U, w2 = (down ? b : -b) - ra.U;
    const float d1 = right ? r.dx : -r.dx, d2 = down ? r.dy : -r.dy;
    const float p1 = fmaf(d1, ts, w1), p2 = fmaf(d2, ts, w2);
    const bool notch = ts < 1000.0f && p1 < -1e-3f && p2 < -1e-3f;         // in the notch: only the quarter circle can be hit
    const float B = fmaf(w1, d1, w2 * d2);
    const float Cq = fmaf(w1, w1, fmaf(w2, w2, -ra.rho2));
    const float D = fmaf(B, B, -Cq);
    const float refined = (D < -0.05f) ? INFINITY : fmaxf(ts, -B - approx_sqrt(fmaxf(D, 0.0f)));
    float entry = notch ? refined : ts;
    entry = fminf(fminf(entry, t2), 1000.0f);
    const int k = (int)floorf((entry - 0.01f) * 0.25f);
    return !ra.enabled ? 0 : (k < 0 ? 0 : (k > LIDAR_MAX_K ? LIDAR_MAX_K : k));
}

// One exact step of the march after the analytic jump: the next sample is tested with the exact arithmetic of
// Lidar.cpp:34-46.
template <class B>
ISX_HD void march_next(const B& bits, const Ray& r, March& m) {
    m.k += 1;
    if (m.k > LIDAR_MAX_K) { m.ke = LIDAR_MAX_K + 1; m.done = true; return; }
    const int e = sample_event(bits, r, m.k, m.px, m.py);
    if (e) { m.ke = m.k; m.hit = (e == 2); m.done = true; }
}
ISX_HD void march_next(const uint32_t* bits, const Ray& r, March& m) { march_next(RoadBitsPtr{bits}, r, m); }
// Road event of one ray, sequential form of what the kernel does (analytic jump, then exact samples).
ISX_HD int ray_road_event_ana(const RoadAna& ra, const uint32_t* bits, const Ray& r, bool* hit, int* tests = nullptr) {
    March m;
    march_init(r, m);
    int n = 0;
    if (!m.done) m.k = ray_safe_samples(ra, r);
    while (!m.done) { march_next(bits, r, m); ++n; }
    if (tests) *tests = n;
    *hit = m.hit;
    return m.ke;
}

// First sample k in [1, kmax] whose pixel lies inside the rectangle, or 0.  A slab test in real arithmetic
// (0.01 px slack, >100x the float rounding of the sample positions) brackets the candidate k range; the
// candidates are then checked with the exact integer test, in order.  Only on-screen pixels can be hit
// (the march breaks off screen first), so the rectangle is clamped to the screen; truncation toward zero
// maps every value in (-1, 1) to pixel 0, hence the wider lower bound when the clamped edge is 0.
ISX_HD int f2i_ceil(float x) {
#if defined(__CUDA_ARCH__)
    return __float2int_ru(x);
#else
    return (int)ceilf(x);
#endif
}
ISX_HD int f2i_floor(float x) {
#if defined(__CUDA_ARCH__)
    return __float2int_rd(x);
#else
    return (int)floorf(x);
#endif
}
ISX_HD int ray_rect_first_hit(const PixRect& r, const Ray& ray, int kmax) {
    const float cx = ray.cx, cy = ray.cy, dx = ray.dx, dy = ray.dy;
    // Slab bounds per axis (prepared with the rectangle, see PixRect).  A direction component of (almost) zero needs no
    // special case: make_ray gives it a finite reciprocal of magnitude >= 1e6, so an origin outside the bounds by more than
    // 0.01 px maps to |t| >= 1e4 (empty bracket), one inside to an unconstrained bracket, and the 0.01 px in between is
    // decided by the exact verification.  An empty rectangle has bounds of 1e30: the bracket is empty for every ray.
    const float ax = (r.lox - cx) * ray.inv_dx, bx = (r.hix - cx) * ray.inv_dx;
    const float ay = (r.loy - cy) * ray.inv_dy, by = (r.hiy - cy) * ray.inv_dy;
    const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), 0.0f);
    const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), (float)(4 * kmax) + 0.5f);
    if (!(t0 <= t1)) return 0;
    // a sample inside the rectangle sits at least 0.0099 px inside the slab bounds (their 0.01 px of slack minus the
    // rounding of the sample position), so 4k lies strictly inside [t0, t1]: no extra candidate on either side
    int ka = f2i_ceil(t0 * 0.25f);
    int kb = f2i_floor(t1 * 0.25f);
    if (ka < 1) ka = 1;
    if (kb > kmax) kb = kmax;
    for (int k = ka; k <= kb; ++k) {
        int px, py;
        ray_pixel(cx, cy, dx, dy, k, px, py);
        if ((unsigned)(px - r.x0) <= (unsigned)(r.x1 - r.x0) && (unsigned)(py - r.y0) <= (unsigned)(r.y1 - r.y0)) return k;
    }
    return 0;
}

// Angular window of beams that can possibly touch a car: the samples that hit the rectangle lie (as real points)
// inside the rectangle grown by 1.01 px, which sits inside a disc of radius rho about its centre; seen from the
// ego at distance D > rho that disc spans +-asin(rho/D) <= +-(x + 0.5708 x^3), x = rho/D.  Beam i (i < R-1) points
// at relative angle -pi + i*2pi/(R-1); beam R-1 duplicates beam 0.  The window is widened by one beam on each
// side (0.065..0.088 rad, against ~1e-6 rad of rounding).  span = 255 means "every beam".
struct BeamWindow { int ia, span, kmin; };
// Conservative set of beams that can touch a car's pixel rectangle, as a circular index interval [ia, ia + span]
// (mod R-1; beam R-1 duplicates beam 0; span 255 = every beam).  The rectangle, grown by 1.01 px for the truncation of
// sample positions to pixels, is seen from the origin under the angles of its silhouette corners: with a = centre
// direction, the offset of corner b is atan(a x b / a . b) (the origin is outside the bounding circle, so |offset| <
// 90 deg and a . b > 0).  Approximate libm / reciprocal are fine: the interval is widened by MARGIN beams (0.25 deg at
// 72 beams, >1000x their error) and every candidate sample is verified exactly by ray_rect_first_hit.
// kmin: no sample before index kmin can lie in the rectangle — a sample sits 4k px from the origin and within 1 px per
// axis of its pixel, so 4k >= (distance to the centre) - (half diagonal of the grown rectangle); half a pixel of slack.
// atan / atan2 for the window only: odd polynomial of degree 9 on [0, 1] (max error 1.2e-5 rad, fitted and measured in
// float32), reciprocal for the rest of the range.  Two of them add up to < 3e-5 rad = 4e-4 beams at 96 beams, against the
// MARGIN of 0.05 beams below (and every candidate sample is verified exactly anyway).  The library atan2f / atanf cost
// ~55 / ~30 instructions each with their exact division and special cases; these cost ~14 / ~10.
ISX_HD float fast_atan01(float a) {
    const float s = a * a;
    return a * fmaf(s, fmaf(s, fmaf(s, fmaf(s, 0.020845098f, -0.08515632f), 0.18015927f), -0.33030477f), 0.9998663f);
}
ISX_HD float fast_atan(float t) {
    const float at = fabsf(t);
    const bool big = at > 1.0f;
    float r = fast_atan01(big ? approx_rcp(at) : at);
    if (big) r = 1.57079632679f - r;
    return copysignf(r, t);
}
ISX_HD float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float r = (mx > 0.0f) ? fast_atan01(mn * approx_rcp(mx)) : 0.0f;
    if (ay > ax) r = 1.57079632679f - r;
    if (x < 0.0f) r = 3.14159265359f - r;
    return copysignf(r, y);
}
ISX_HD BeamWindow beam_window(const PixRect& r, float cx, float cy, float heading, int R) {
    BeamWindow w;
    w.ia = 0; w.span = 255; w.kmin = 0;
    if (R < 4) return w;
    const float hx = 0.5f * (float)(r.x1 - r.x0) + 1.01f, hy = 0.5f * (float)(r.y1 - r.y0) + 1.01f;
    const float ccx = 0.5f * (float)(r.x0 + r.x1), ccy = 0.5f * (float)(r.y0 + r.y1);
    const float X = ccx - cx, Y = ccy - cy;
    const float D2 = X * X + Y * Y, rho2 = hx * hx + hy * hy;
    if (!(D2 > rho2 * 1.05f + 1.0f)) return w;
    {
        const int km = (int)floorf((approx_sqrt(D2) - approx_sqrt(rho2) - 0.5f) * 0.25f);    // 1e-4 px of error against the 0.5
        w.kmin = km < 0 ? 0 : (km > 255 ? 255 : km);
    }
    float tmin = 0.0f, tmax = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float px = X + ((k & 1) ? hx : -hx), py = Y + ((k & 2) ? hy : -hy);
        const float cross = Y * px - X * py;              // (X, -Y) x (px, -py): lidar angles are atan2(-y, x)
        const float dot = X * px + Y * py;
        const float t = cross * approx_rcp(dot);
        tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
    }
    const float MARGIN = 0.05f;
    const float inv_step = (float)(R - 1) * (1.0f / 6.28318530718f);
    float phi = fast_atan2(-Y, X) - heading;
    phi = phi - 6.28318530718f * floorf(phi * (1.0f / 6.28318530718f) + 0.5f);   // to [-pi, pi]
    const float f_lo = (phi + fast_atan(tmin) * 1.0001f + 3.14159265359f) * inv_step;
    const float f_hi = (phi + fast_atan(tmax) * 1.0001f + 3.14159265359f) * inv_step;
    int ia = (int)floorf(f_lo - MARGIN);
    const int ib = (int)ceilf(f_hi + MARGIN);
    const int span = ib - ia;
    if (span >= R - 2) return w;
    const int m = R - 1;
    ia %= m;
    if (ia < 0) ia += m;
    w.ia = ia; w.span = span;
    return w;
}
ISX_HD bool beam_in_window(const BeamWindow& w, int i, int R) {
    if (w.span >= 255) return true;
    const int ii = (i == R - 1) ? 0 : i;
    int dlt = ii - w.ia;
    if (dlt < 0) dlt += R - 1;
    return dlt <= w.span;
}

// ---------------------------------------------------------------- NPC planner pieces (TrafficFlow.cpp:22-196)
// Pairwise, ghost-index-independent part of the yield logic for NPC `me` against NPC `ot`:
//   bit0 : `ot` can conflict at all (not same-direction < 60 deg :103-104, not a stable side-by-side car :107-159)
//   bit1 : yield rules 2-4 hold (:167-176); rule 1 (dist_to_crash < 15) depends on the ghost point.
// What NPC `me` needs to know about one other NPC `ot` when it plans (TrafficFlow.cpp): the other's contribution to
// get_front_car_dist_tf (:28-44) in *front (its distance, or 1e9), and the conflict flags of :91-185 as the return value
// (bit 0: eligible for the ghost-path scan, bit 1: me yields to it).  One function because both start from the same
// centre distance hypot(dx, dy) and the same |wrap(me.h - ot.h)| — on the list-order critical path of k_traffic a
// second hypot + fmod per NPC is ~10% of the step of the busiest env.
ISX_HD int npc_pair_eval(const Pose& me, const Pose& ot, float me_sin, float me_cos, float me_dc, bool me_before_ot, float* front) {
    const float dxt = ot.x - me.x, dyt = ot.y - me.y;
    const float dto = hypotf_(dxt, dyt);
    const float ad = fabsf(wrap_angle(me.h - ot.h));
    const float mx = me_cos, my = -me_sin;
    float f = 1e9f;
    if (!(dto > 80.0f)) {
        const float dot = (dxt * mx + dyt * my) / (dto + 1e-5f);
        if (dot > 0.8f && ad < (45.0f * PI_F / 180.0f)) f = dto;
    }
    *front = f;
    if (ad < (60.0f * PI_F / 180.0f)) return 0;
    if (dto > 1e-5f) {
        const float adn = fminf(ad, 2.0f * PI_F - ad);
        const bool parallel = (adn < (30.0f * PI_F / 180.0f)) || (adn > (150.0f * PI_F / 180.0f));
        if (parallel) {
            const float lon = dxt * mx + dyt * my;
            const float lat = fsqrt_rn(fmaxf(0.0f, dto * dto - lon * lon));
            if (fabsf(lat) < (LANE_WIDTH_PX * 1.5f) && fabsf(lon) < (CAR_LENGTH * 2.0f)) {
                const float fd = 20.0f;
                const float mfx = me.x + mx * fd, mfy = me.y + my * fd;
                float so, co;
                sincosf_nc(ot.h, &so, &co);
                const float ofx = ot.x + co * fd, ofy = ot.y + (-so) * fd;
                const float fdx = ofx - mfx, fdy = ofy - mfy;
                const float fmag = hypotf_nc(fdx, fdy);
                if (fmag > 1e-5f) {
                    const float flon = fdx * mx + fdy * my;
                    const float flat = fsqrt_rn(fmaxf(0.0f, fmag * fmag - flon * flon));
                    if (fabsf(flat - lat) < (LANE_WIDTH_PX * 0.5f)) return 0;
                }
            }
        }
    }
    const float ot_dc = hypotf_nc(ot.x - WIDTH * 0.5f, ot.y - HEIGHT * 0.5f);
    bool y = false;
    if (me.v < 1.0f && ot.v > 3.0f && ot_dc < me_dc + 25.0f) y = true;
    else if (ot_dc < me_dc - 5.0f) y = true;
    else if (fabsf(ot_dc - me_dc) <= 5.0f) y = me_before_ot;
    return 1 | (y ? 2 : 0);
}

// Lateral P-control + cruise thresholds + front-car braking (TrafficFlow.cpp:50-75).
ISX_HD float npc_steer_cmd(const Pose& me, F2 target) {
    const float dx = target.x - me.x, dy = target.y - me.y;
    const float err = wrap_angle(atan2f_nc(-dy, dx) - me.h);
    return fmaxf(-1.0f, fminf(1.0f, err * 3.0f));
}
ISX_HD float npc_cruise_throttle(float v, float front_dist) {
    const float target = PHYSICS_MAX_SPEED * 0.4f;
    float thr = 0.0f;
    if (v < target) thr = 0.5f;
    else if (v > target + 1.0f) thr = -0.1f;
    if (front_dist < 30.0f) thr = -1.0f;
    else if (front_dist < 50.0f) thr = fminf(thr, -0.2f);
    return thr;
}
ISX_HD float npc_final_throttle(float thr, bool conflict, float min_conflict_dist) {
    if (!conflict) return thr;
    if (min_conflict_dist < 35.0f) return -1.0f;
    if (min_conflict_dist < 60.0f) return -0.8f;
    return fminf(thr, 0.0f);
}

// ---------------------------------------------------------------- observation pieces (IntersectionEnv.cpp:431-458, 494-507)
ISX_HD void obs_ego_features(const Pose& p, F2 target, float* o6) {
    o6[0] = p.x / (float)WIDTH;
    o6[1] = p.y / (float)HEIGHT;
    o6[2] = p.v / PHYSICS_MAX_SPEED;
    o6[3] = p.h / PI_F;
    const float dx = target.x - p.x, dy = target.y - p.y;
    o6[4] = fsqrt_rn(dx * dx + dy * dy) / (float)WIDTH;
    o6[5] = wrap_angle(atan2f_nc(-dy, dx) - p.h) / PI_F;
}
ISX_HD void obs_neighbor_features(const Pose& me, const Pose& ot, int intent, float* o5) {
    o5[0] = (ot.x - me.x) / (float)WIDTH;
    o5[1] = (ot.y - me.y) / (float)HEIGHT;
    o5[2] = (ot.v - me.v) / PHYSICS_MAX_SPEED;
    o5[3] = wrap_angle(ot.h - me.h) / PI_F;
    o5[4] = (float)intent;
}

// ------------------------------------------------------------------------------------------------ neighbour order
// IntersectionEnv.cpp:490 orders the neighbour list with std::sort and a `dist <` comparator.  std::sort is not stable:
// when two distances are EXACTLY equal and the list is longer than 16, which of the two comes first is decided by the
// library's introsort.  The reference "as run here" is libstdc++ 13 (bits/stl_algo.h __sort / __introsort_loop /
// __final_insertion_sort, bits/stl_heap.h for the depth-limit fallback), so that algorithm is restated below on a
// permutation p[0..n) of list positions ordered by key[p[i]].  For n <= 16 it degenerates to a stable insertion sort,
// which is what the fast path of k_features computes; the kernel only comes here when a tie can change the answer.
namespace stdsort {
constexpr int THRESHOLD = 16;
ISX_HD bool lt(const float* key, uint8_t a, uint8_t b) { return key[a] < key[b]; }
ISX_HD void swp(uint8_t* p, int i, int j) { const uint8_t t = p[i]; p[i] = p[j]; p[j] = t; }
ISX_HD void sift_up(const float* key, uint8_t* a, int hole, int top, uint8_t val) {
    int parent = (hole - 1) / 2;
    while (hole > top && lt(key, a[parent], val)) { a[hole] = a[parent]; hole = parent; parent = (hole - 1) / 2; }
    a[hole] = val;
}
ISX_HD void sift_hole(const float* key, uint8_t* a, int hole, int len, uint8_t val) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lt(key, a[child], a[child - 1])) --child;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1];
        hole = child - 1;
    }
    sift_up(key, a, hole, top, val);
}
ISX_HD void heap_sort(const float* key, uint8_t* a, int len) {     // __partial_sort(first, last, last)
    if (len >= 2)
        for (int parent = (len - 2) / 2;; --parent) {
            sift_hole(key, a, parent, len, a[parent]);
            if (parent == 0) break;
        }
    for (int last = len - 1; last >= 1; --last) {
        const uint8_t val = a[last];
        a[last] = a[0];
        sift_hole(key, a, 0, last, val);
    }
}
ISX_HD int partition_pivot(const float* key, uint8_t* p, int first, int last) {
    const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
    int med;                                        // median of (a, b, c) under `<`, moved to `first`
    if (lt(key, p[a], p[b])) med = lt(key, p[b], p[c]) ? b : (lt(key, p[a], p[c]) ? c : a);
    else med = lt(key, p[a], p[c]) ? a : (lt(key, p[b], p[c]) ? c : b);
    swp(p, first, med);
    int lo = first + 1, hi = last;
    for (;;) {
        while (lo < last && lt(key, p[lo], p[first])) ++lo;        // bounds only matter for NaN keys (UB in the reference)
        --hi;
        while (hi > first && lt(key, p[first], p[hi])) --hi;
        if (!(lo < hi)) return lo;
        swp(p, lo, hi);
        ++lo;
    }
}
ISX_HD void linear_insert(const float* key, uint8_t* p, int i, int floor_) {
    const uint8_t val = p[i];
    int next = i - 1;
    while (next >= floor_ && lt(key, val, p[next])) { p[i] = p[next]; i = next; --next; }
    p[i] = val;
}
ISX_HD void insertion(const float* key, uint8_t* p, int first, int last) {
    for (int i = first + 1; i < last; ++i) {
        if (lt(key, p[i], p[first])) {
            const uint8_t val = p[i];
            for (int j = i; j > first; --j) p[j] = p[j - 1];
            p[first] = val;
        } else {
            linear_insert(key, p, i, first);
        }
    }
}
// Sorts p[0..n) (caller fills it, normally with the identity); returns how many ranges fell back to the heap sort.
ISX_HD int sort(const float* key, uint8_t* p, int n) {
    if (n <= 0) return 0;
    int heaps = 0;
    int lg = 0;
    while ((n >> (lg + 1)) != 0) ++lg;
    int sf[24], sl[24], sd[24], sp = 0;             // pending (first, last, depth) ranges; they are disjoint, order is free
    sf[0] = 0; sl[0] = n; sd[0] = 2 * lg; sp = 1;
    while (sp) {
        --sp;
        const int first = sf[sp];
        int last = sl[sp], depth = sd[sp];
        while (last - first > THRESHOLD) {
            if (depth == 0) { heap_sort(key, p + first, last - first); ++heaps; break; }
            --depth;
            const int cut = partition_pivot(key, p, first, last);
            sf[sp] = cut; sl[sp] = last; sd[sp] = depth; ++sp;
            last = cut;
        }
    }
    if (n > THRESHOLD) {
        insertion(key, p, 0, THRESHOLD);
        for (int i = THRESHOLD; i < n; ++i) linear_insert(key, p, i, 0);
    } else {
        insertion(key, p, 0, n);
    }
    return heaps;
}
}  // namespace stdsort

}  // namespace isx
