// isx_tables.h — host-side precompute uploaded once per handle: lane layout and route way-point LUT
// (RouteGen.cpp:7-205), folded road bitmap + sphere-tracing skip table (from RoadGeometry::is_on_road),
// lidar beam angles (Lidar.cpp:4-14).  Header-only so the host unit-test library can use it too.
// Uses the restated libm of isx_math.cuh, never the system libm, so the LUT is a pure function of this
// source (and is checked bit-for-bit against the reference's own generate_path_cpp in the tests).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "isx_sim.cuh"

namespace isx {

struct RouteHost {
    F2 path[PATH_LEN];
    int intent = 1;
    float spawn_x = 0, spawn_y = 0, spawn_h = 0;
};

// RouteGen.cpp:7-53.  "IN_k"/"OUT_k", k = dir*lanes + lane + 1, dir in N,E,S,W.  false = unknown id.
inline bool lane_point(int lanes, const char* id, float* x, float* y, int* dir) {
    bool is_in;
    const char* num;
    if (std::strncmp(id, "IN_", 3) == 0) { is_in = true; num = id + 3; }
    else if (std::strncmp(id, "OUT_", 4) == 0) { is_in = false; num = id + 4; }
    else return false;
    if (*num == '\0' || *num == '0') return false;
    int k = 0;
    for (const char* p = num; *p; ++p) {
        if (*p < '0' || *p > '9') return false;
        k = k * 10 + (*p - '0');
        if (k > 100000) return false;
    }
    if (k < 1 || k > 4 * lanes) return false;
    const int d = (k - 1) / lanes, j = (k - 1) % lanes;
    const float C = WIDTH * 0.5f, margin = 30.0f;
    const float off = LANE_WIDTH_PX * (0.5f + (float)j);
    float ix, iy, ox, oy;
    switch (d) {
        case 0: ix = C - off; iy = margin; ox = C + off; oy = margin; break;                                   // N
        case 1: ix = (float)WIDTH - margin; iy = C - off; ox = (float)WIDTH - margin; oy = C + off; break;     // E
        case 2: ix = C + off; iy = (float)HEIGHT - margin; ox = C - off; oy = (float)HEIGHT - margin; break;   // S
        default: ix = margin; iy = C + off; ox = margin; oy = C - off; break;                                  // W
    }
    *x = is_in ? ix : ox;
    *y = is_in ? iy : oy;
    *dir = d;
    return true;
}

// RouteGen.cpp:55-87
inline int route_intent(int s, int e) {
    if (e == ((s + 2) & 3)) return 0;   // opposite  -> STRAIGHT
    if (e == ((s + 1) & 3)) return 1;   // N->E, E->S, S->W, W->N -> LEFT
    if (e == ((s + 3) & 3)) return 2;   // N->W, E->N, S->E, W->S -> RIGHT
    return 1;
}

inline F2 box_projection(int lanes, F2 p) {   // RouteGen.cpp:89-101
    const float C = WIDTH * 0.5f, tb = (float)lanes * LANE_WIDTH_PX;
    if (p.y < C - tb) return F2{p.x, C - tb};
    if (p.y > C + tb) return F2{p.x, C + tb};
    if (p.x < C - tb) return F2{C - tb, p.y};
    return F2{C + tb, p.y};
}

inline void lerp_points(F2* out, int count, F2 a, F2 b) {
    for (int i = 0; i < count; ++i) {
        const float t = (float)i / (float)count;
        out[i] = F2{a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t};
    }
}

// 0 ok, -1 unknown start, -2 unknown end.  RouteGen.cpp:111-205; spawn heading IntersectionEnv.cpp:87-92.
inline int build_route(int lanes, const char* start, const char* end, RouteHost* r) {
    F2 ps, pe;
    int ds, de;
    if (!lane_point(lanes, start, &ps.x, &ps.y, &ds)) return -1;
    if (!lane_point(lanes, end, &pe.x, &pe.y, &de)) return -2;
    const float C = WIDTH * 0.5f;
    r->intent = route_intent(ds, de);
    if (r->intent != 2) {
        const F2 en = box_projection(lanes, ps), ex = box_projection(lanes, pe);
        lerp_points(r->path, 50, ps, en);
        if (r->intent == 0) lerp_points(r->path + 50, 60, en, ex);
        else {
            for (int i = 0; i < 60; ++i) {   // quadratic Bezier with the intersection centre as control point
                const float t = (float)i / 60.0f;
                const float a = (1 - t) * (1 - t), b = 2 * (1 - t) * t, c = t * t;
                r->path[50 + i] = F2{a * en.x + b * C + c * ex.x, a * en.y + b * C + c * ex.y};
            }
        }
        lerp_points(r->path + 110, 50, ex, pe);
    } else {
        const float rh = (float)lanes * LANE_WIDTH_PX;
        const float lo = C - rh - CORNER_RADIUS, hi = C + rh + CORNER_RADIUS;
        float ccx, ccy, t0, t1;
        switch (ds) {
            case 0: ccx = lo; ccy = lo; t0 = 0.0f; t1 = PI_F / 2.0f; break;
            case 1: ccx = hi; ccy = lo; t0 = PI_F / 2.0f; t1 = PI_F; break;
            case 2: ccx = hi; ccy = hi; t0 = PI_F; t1 = 3.0f * PI_F / 2.0f; break;
            default: ccx = lo; ccy = hi; t0 = -PI_F / 2.0f; t1 = 0.0f; break;
        }
        const float rad = CORNER_RADIUS + 0.5f * LANE_WIDTH_PX;
        float s0, c0, s1, c1;
        sincosf_(t0, &s0, &c0);
        sincosf_(t1, &s1, &c1);
        const F2 a0{ccx + rad * c0, ccy + rad * s0}, a1{ccx + rad * c1, ccy + rad * s1};
        lerp_points(r->path, 50, ps, a0);
        for (int i = 0; i < 60; ++i) {
            const float t = (float)i / 60.0f;
            const float th = t0 + (t1 - t0) * t;
            float s, c;
            sincosf_(th, &s, &c);
            r->path[50 + i] = F2{ccx + rad * c, ccy + rad * s};
        }
        lerp_points(r->path + 110, 50, a1, pe);
    }
    r->spawn_x = ps.x;
    r->spawn_y = ps.y;
    r->spawn_h = atan2f_(-(r->path[1].y - r->path[0].y), r->path[1].x - r->path[0].x);
    return 0;
}

// Lidar beam offsets (Lidar.cpp:7-13 / IntersectionEnv.cpp:119-127): -180 + i*360/(R-1) degrees, in radians.
inline void lidar_rel_angles(int rays, float* rel) {
    const float fov = 360.0f;
    const float start = -fov * 0.5f;
    const float step = (rays > 1) ? (fov / (float)(rays - 1)) : 0.0f;
    for (int i = 0; i < rays; ++i) {
        const float deg = start + i * step;
        rel[i] = deg * PI_F / 180.0f;
    }
}

struct RoadTables {
    std::vector<uint32_t> bits;   // ROAD_ROWS x ROAD_WORDS
    std::vector<uint8_t> skip;    // SKIP_DIM x SKIP_DIM
    int box_lo = 1, box_hi = 0;   // every pixel of [box_lo, box_hi] x [0,749] and of its transpose is road (empty if lo > hi)
    RoadAna ana{0.0f, 0.0f, 0.0f, 0};   // analytic off-road bound (ray_safe_samples); enabled only if verified below
};

// Builds the folded bitmap and skip table from on_road() evaluated at every integer pixel, exactly as the
// lidar calls it (Lidar.cpp:44: is_on_road(float(px), float(py))).  Returns false if the full-resolution
// map is not mirror-symmetric about 375 (it always is; the check keeps the folding honest).
inline bool build_road_tables(int lanes, RoadTables* t) {
    const int W = WIDTH, H = HEIGHT;
    std::vector<uint8_t> road((size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) road[(size_t)y * W + x] = on_road(lanes, (float)x, (float)y) ? 1 : 0;
    // Chebyshev clearance to the nearest event pixel (off-road or off-screen), two-pass chamfer (exact for L-inf)
    std::vector<int> clr((size_t)W * H);
    const int BIG = 1 << 20;
    auto at = [&](int x, int y) -> int { return (x < 0 || y < 0 || x >= W || y >= H) ? 0 : clr[(size_t)y * W + x]; };
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int c = road[(size_t)y * W + x] ? BIG : 0;
            if (c) {
                int m = at(x - 1, y);
                m = std::min(m, at(x - 1, y - 1)); m = std::min(m, at(x, y - 1)); m = std::min(m, at(x + 1, y - 1));
                c = std::min(c, m + 1);
            }
            clr[(size_t)y * W + x] = c;
        }
    for (int y = H - 1; y >= 0; --y)
        for (int x = W - 1; x >= 0; --x) {
            int c = clr[(size_t)y * W + x];
            if (c) {
                int m = at(x + 1, y);
                m = std::min(m, at(x + 1, y + 1)); m = std::min(m, at(x, y + 1)); m = std::min(m, at(x - 1, y + 1));
                c = std::min(c, m + 1);
            }
            clr[(size_t)y * W + x] = c;
        }
    // strip interior: columns strictly inside the vertical strip are road over the full screen height (the grass
    // discs only touch the strip walls).  Verified pixel by pixel; an unexpected geometry simply disables the box.
    {
        const int rw = lanes * (int)LANE_WIDTH_PX;
        int lo = WIDTH / 2 - rw + 1, hi = WIDTH / 2 + rw - 1;
        bool ok = lo <= hi && lo >= 0 && hi < W;
        for (int y = 0; ok && y < H; ++y)
            for (int x = lo; ok && x <= hi; ++x) ok = road[(size_t)y * W + x] && road[(size_t)x * W + y];
        if (ok) { t->box_lo = lo; t->box_hi = hi; } else { t->box_lo = 1; t->box_hi = 0; }
    }
    // analytic bound (isx_sim.cuh ray_safe_samples): every off-road pixel must lie within CORNER_RADIUS of one of the four
    // quarter planes {sx (x - 375) >= U, sy (y - 375) >= U}; checked here for every pixel of the screen, in double.
    {
        const double U = (double)lanes * LANE_WIDTH_PX + CORNER_RADIUS, cr = CORNER_RADIUS;
        bool ok = (double)lanes * LANE_WIDTH_PX - 1.5 > 0.0;
        for (int y = 0; ok && y < H; ++y)
            for (int x = 0; ok && x < W; ++x) {
                if (road[(size_t)y * W + x]) continue;
                const double a = std::fabs((double)x - ROAD_HALF), b = std::fabs((double)y - ROAD_HALF);   // nearest quarter plane = own quadrant
                const double e1 = std::max(U - a, 0.0), e2 = std::max(U - b, 0.0);
                ok = std::sqrt(e1 * e1 + e2 * e2) <= cr + 1e-6;
            }
        const float rho = CORNER_RADIUS + 1.5f;
        t->ana = RoadAna{(float)lanes * LANE_WIDTH_PX - 1.5f, (float)lanes * LANE_WIDTH_PX + CORNER_RADIUS, rho * rho, ok ? 1 : 0};
    }
    t->bits.assign((size_t)ROAD_ROWS * ROAD_WORDS, 0u);
    std::vector<int> fclr((size_t)ROAD_ROWS * ROAD_ROWS, BIG);
    for (int v = 0; v <= ROAD_HALF; ++v)
        for (int u = 0; u <= ROAD_HALF; ++u) {
            int bit = -1, cmin = BIG;
            const int xs[2] = {ROAD_HALF - u, ROAD_HALF + u}, ys[2] = {ROAD_HALF - v, ROAD_HALF + v};
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    const int x = xs[a], y = ys[b];
                    if (x < 0 || x >= W || y < 0 || y >= H) { cmin = 0; continue; }   // mirror image off screen: no skipping
                    const int r = road[(size_t)y * W + x];
                    if (bit < 0) bit = r;
                    else if (bit != r) return false;
                    cmin = std::min(cmin, clr[(size_t)y * W + x]);
                }
            if (bit > 0) t->bits[(size_t)v * ROAD_WORDS + (u >> 5)] |= (1u << (u & 31));
            fclr[(size_t)v * ROAD_ROWS + u] = cmin;
        }
    // Sample k at pixel P with clearance C: sample k+j lies within Chebyshev distance 4j+1 of P (|d| <= 1 per
    // axis, truncation adds < 1), so it cannot be an event while 4j+1 < C.  One more pixel of slack is kept:
    // j_max = floor((C-3)/4).  Per 4x4 block: the minimum over the block.
    t->skip.assign((size_t)SKIP_DIM * SKIP_DIM, 0);
    for (int bv = 0; bv < SKIP_DIM; ++bv)
        for (int bu = 0; bu < SKIP_DIM; ++bu) {
            int cmin = BIG;
            for (int v = bv * 4; v < bv * 4 + 4 && v <= ROAD_HALF; ++v)
                for (int u = bu * 4; u < bu * 4 + 4 && u <= ROAD_HALF; ++u) cmin = std::min(cmin, fclr[(size_t)v * ROAD_ROWS + u]);
            int j = (cmin - 3) / 4;
            if (cmin < 3) j = 0;
            if (j > LIDAR_MAX_K) j = LIDAR_MAX_K;
            t->skip[(size_t)bv * SKIP_DIM + bu] = (uint8_t)j;
        }
    return true;
}

}  // namespace isx
