/* oracle/isx_oracle.c — TEST INFRASTRUCTURE (not product code).
 *
 * Plain-C restatement of the per-step simulation of ShamG1/marl-traffic-intersection
 * (the hot path IntersectionEnv::step, /root/reference/cpp/IntersectionEnv.cpp:133-392, and
 * everything it calls).  One env at a time, sequential, float32, calling the same libm
 * entry points the reference binds (sincosf, tanf, atan2f, hypotf, fmodf, expf, sqrtf).
 * Every function cites the reference lines it follows.  Nothing here is copied from the
 * reference: it is a from-scratch restatement over flat structs.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference ITSELF: oracle/_ref/libisx_ref.so (the
 * unmodified reference sources behind ref_driver.cpp) run in this container, both live
 * (tests/test_oracle_vs_ref.py, skipped where _ref is absent) and through committed
 * fixtures generated from it (tests/golden/, tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 * Compile: gcc -O2 -std=c11 -ffp-contract=off (no -march, no -ffast-math); see Makefile.
 */
#define _GNU_SOURCE
#include "isx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "philox.h"

/* constants.h:4-20 */
#define W_PX 750
#define H_PX 750
#define CAR_LENGTH 54.0f
#define CAR_WIDTH 24.0f
#define LANE_W 42.0f
#define CORNER_R 84.0f
#define MAX_ACC 15.0f
#define MAX_STEER 0.6108652381980153f
#define MAX_SPEED 8.0f
#define FPS_F 60.0f
#define SCALE_F 12.0f
static const float PI_F = 3.14159265358979323846f; /* Car.cpp:7 */

#define MAX_AGENTS 32
#define MAX_NPC 64
#define MAX_ROUTES 64
#define MAX_RAYS 128

typedef struct {
    float px[ISX_PATH_LEN], py[ISX_PATH_LEN];
    int n;          /* path length (160) or 0 when the start id is unknown */
    int intent;
    float sx, sy, sh; /* spawn pose */
    char start[16], end[16];
} route_t;

typedef struct {
    float x, y, v, h, acc, steer, prev_dist, pa0, pa1;
    int path_index, alive, route;
    uint32_t uid;
} car_t;

struct isxo_env {
    int lanes, use_team, respawn, max_steps, traffic, rays;
    float density;
    float k_prog, v_min, k_stuck, k_cv, k_co, k_succ, k_sm, alpha;
    route_t ego_routes[MAX_AGENTS]; int n_ego_routes;
    route_t traffic_routes[MAX_ROUTES]; int n_traffic_routes;
    car_t cars[MAX_AGENTS]; int n; int slot_route[MAX_AGENTS];
    float lidar[MAX_AGENTS][MAX_RAYS];
    float rel[MAX_RAYS];
    car_t npcs[MAX_NPC]; int n_npc;
    int step_count;
    uint64_t seed; uint32_t env_id, tick, next_uid, draw;
    uint8_t *line_mask;
    isx_traffic_events ev;
};

/* libm calls go through these so the compiler can never fold them with MPFR */
static inline void sc(float a, float *s, float *c) { volatile float va = a; sincosf(va, s, c); }
static inline float wrap_angle(float a) { /* IntersectionEnv.cpp:9-13, TrafficFlow.cpp:8-12 */
    a = fmodf(a + PI_F, 2.0f * PI_F);
    if (a < 0) a += 2.0f * PI_F;
    return a - PI_F;
}

/* ------------------------------------------------------------------ lane layout / routes */
/* RouteGen.cpp:7-53.  id = "IN_k" / "OUT_k", k = d*lanes + j + 1, d in N,E,S,W */
static int lane_lookup(int lanes, const char *id, float *x, float *y, int *dir) {
    int is_in, k;
    if (strncmp(id, "IN_", 3) == 0) { is_in = 1; k = atoi(id + 3); }
    else if (strncmp(id, "OUT_", 4) == 0) { is_in = 0; k = atoi(id + 4); }
    else return 0;
    if (k < 1 || k > 4 * lanes) return 0;
    { /* reject trailing junk such as "IN_3x" */
        char tmp[24]; snprintf(tmp, sizeof tmp, "%s%d", is_in ? "IN_" : "OUT_", k);
        if (strcmp(tmp, id) != 0) return 0;
    }
    const int d = (k - 1) / lanes, j = (k - 1) % lanes;
    const float CX = W_PX * 0.5f, CY = H_PX * 0.5f, MARGIN = 30.0f;
    const float off = LANE_W * (0.5f + (float)j);
    float ix, iy, ox, oy;
    if (d == 0)      { ix = CX - off; iy = MARGIN;        ox = CX + off; oy = MARGIN; }
    else if (d == 2) { ix = CX + off; iy = H_PX - MARGIN; ox = CX - off; oy = H_PX - MARGIN; }
    else if (d == 1) { ix = W_PX - MARGIN; iy = CY - off; ox = W_PX - MARGIN; oy = CY + off; }
    else             { ix = MARGIN; iy = CY + off;        ox = MARGIN; oy = CY - off; }
    *x = is_in ? ix : ox; *y = is_in ? iy : oy; *dir = d;
    return 1;
}

/* RouteGen.cpp:55-87; directions 0..3 = N,E,S,W */
static int intent_of(int s, int e) {
    static const int opp[4] = {2, 3, 0, 1};
    static const int left[4] = {1, 2, 3, 0};   /* N->E, E->S, S->W, W->N */
    static const int right[4] = {3, 0, 1, 2};  /* N->W, E->N, S->E, W->S */
    if (e == opp[s]) return 0;
    if (e == left[s]) return 1;
    if (e == right[s]) return 2;
    return 1;
}

/* RouteGen.cpp:89-101 */
static void project_to_box(int lanes, float x, float y, float *ox, float *oy) {
    const float CX = W_PX * 0.5f, CY = H_PX * 0.5f;
    const float tb = lanes * LANE_W;
    const float bl = CX - tb, br = CX + tb, bt = CY - tb, bb = CY + tb;
    if (y < bt) { *ox = x; *oy = bt; return; }
    if (y > bb) { *ox = x; *oy = bb; return; }
    if (x < bl) { *ox = bl; *oy = y; return; }
    *ox = br; *oy = y;
}

static void seg(route_t *r, int *n, int cnt, float ax, float ay, float bx, float by) {
    for (int i = 0; i < cnt; ++i) {
        const float t = (float)i / (float)cnt;
        r->px[*n] = ax + (bx - ax) * t;
        r->py[*n] = ay + (by - ay) * t;
        (*n)++;
    }
}

/* RouteGen.cpp:111-205 + spawn pose IntersectionEnv.cpp:83-99.  Returns 0 ok, -1 unknown start, -2 unknown end */
static int build_route(int lanes, const char *start, const char *end, route_t *r) {
    memset(r, 0, sizeof *r);
    snprintf(r->start, sizeof r->start, "%s", start);
    snprintf(r->end, sizeof r->end, "%s", end);
    float sx, sy, ex, ey; int sd, ed;
    if (!lane_lookup(lanes, start, &sx, &sy, &sd)) return -1;
    if (!lane_lookup(lanes, end, &ex, &ey, &ed)) return -2;
    const float CX = W_PX * 0.5f, CY = H_PX * 0.5f;
    const int intent = intent_of(sd, ed);
    float enx, eny, exx, exy;
    project_to_box(lanes, sx, sy, &enx, &eny);
    project_to_box(lanes, ex, ey, &exx, &exy);
    int n = 0;
    if (intent == 0 || intent == 1) {
        seg(r, &n, 50, sx, sy, enx, eny);
        if (intent == 0) seg(r, &n, 60, enx, eny, exx, exy);
        else {
            for (int i = 0; i < 60; ++i) { /* quadratic Bezier through the centre, :103-109 */
                const float t = (float)i / 60.0f;
                r->px[n] = (1 - t) * (1 - t) * enx + 2 * (1 - t) * t * CX + t * t * exx;
                r->py[n] = (1 - t) * (1 - t) * eny + 2 * (1 - t) * t * CY + t * t * exy;
                n++;
            }
        }
        seg(r, &n, 50, exx, exy, ex, ey);
    } else {
        const float rh = lanes * LANE_W;
        float ccx, ccy, ts, te;
        if (sd == 0)      { ccx = CX - rh - CORNER_R; ccy = CY - rh - CORNER_R; ts = 0.0f; te = PI_F / 2.0f; }
        else if (sd == 1) { ccx = CX + rh + CORNER_R; ccy = CY - rh - CORNER_R; ts = PI_F / 2.0f; te = PI_F; }
        else if (sd == 2) { ccx = CX + rh + CORNER_R; ccy = CY + rh + CORNER_R; ts = PI_F; te = 3.0f * PI_F / 2.0f; }
        else              { ccx = CX - rh - CORNER_R; ccy = CY + rh + CORNER_R; ts = -PI_F / 2.0f; te = 0.0f; }
        const float rad = CORNER_R + 0.5f * LANE_W;
        float s0, c0, s1, c1;
        sc(ts, &s0, &c0); sc(te, &s1, &c1);
        const float asx = ccx + rad * c0, asy = ccy + rad * s0;
        const float aex = ccx + rad * c1, aey = ccy + rad * s1;
        seg(r, &n, 50, sx, sy, asx, asy);
        for (int i = 0; i < 60; ++i) {
            const float t = (float)i / 60.0f;
            const float th = ts + (te - ts) * t;
            float s, c; sc(th, &s, &c);
            r->px[n] = ccx + rad * c; r->py[n] = ccy + rad * s; n++;
        }
        seg(r, &n, 50, aex, aey, ex, ey);
    }
    r->n = n; r->intent = intent; r->sx = sx; r->sy = sy;
    { volatile float dy = -(r->py[1] - r->py[0]); volatile float dx = r->px[1] - r->px[0]; r->sh = atan2f(dy, dx); }
    return 0;
}

/* ------------------------------------------------------------------ geometry */
/* RoadGeometry.h:19-58 */
static int on_road(int lanes, float x, float y) {
    const float CX = W_PX * 0.5f, CY = H_PX * 0.5f;
    const float rw = lanes * LANE_W, cr = CORNER_R, r2 = cr * cr;
    const float gx[4] = {CX - rw - cr, CX + rw + cr, CX - rw - cr, CX + rw + cr};
    const float gy[4] = {CY - rw - cr, CY - rw - cr, CY + rw + cr, CY + rw + cr};
    for (int i = 0; i < 4; ++i) {
        const float dx = x - gx[i], dy = y - gy[i];
        if (dx * dx + dy * dy <= r2) return 0;
    }
    if ((x >= CX - rw && x <= CX + rw) || (y >= CY - rw && y <= CY + rw)) return 1;
    if (x >= CX - rw - cr && x <= CX - rw && y >= CY - rw - cr && y <= CY - rw) return 1;
    if (x >= CX + rw && x <= CX + rw + cr && y >= CY - rw - cr && y <= CY - rw) return 1;
    if (x >= CX - rw - cr && x <= CX - rw && y >= CY + rw && y <= CY + rw + cr) return 1;
    if (x >= CX + rw && x <= CX + rw + cr && y >= CY + rw && y <= CY + rw + cr) return 1;
    return 0;
}
/* RoadGeometry.h:60-67 */
static int hits_yellow(int lanes, float x, float y) {
    const float cx = W_PX * 0.5f, cy = H_PX * 0.5f, gap = 2.0f, rw = lanes * LANE_W;
    if (fabsf(x - cx) <= gap && fabsf(y - cy) > rw) return 1;
    if (fabsf(y - cy) <= gap && fabsf(x - cx) > rw) return 1;
    return 0;
}
/* LineMask.cpp:8-72: 8 axis-aligned segments, thickness 2 -> +-1 px */
static uint8_t *make_line_mask(int lanes) {
    uint8_t *g = (uint8_t *)calloc((size_t)W_PX * H_PX, 1);
    const int cx = W_PX / 2, cy = H_PX / 2;
    const int stop = lanes * (int)LANE_W + (int)CORNER_R;
    const int vx[2] = {cx - 2, cx + 2};
    for (int l = 0; l < 2; ++l) {
        for (int d = -1; d <= 1; ++d) {
            const int c = vx[l] + d; /* column (vertical lines) / row (horizontal lines) */
            for (int t = 0; t < W_PX; ++t) {
                if (t <= cy - stop || t >= cy + stop) {
                    g[(size_t)t * W_PX + c] = 1; /* vertical line pixel (x=c, y=t) */
                    g[(size_t)c * W_PX + t] = 1; /* horizontal line pixel (x=t, y=c) */
                }
            }
        }
    }
    return g;
}
static int is_line(const struct isxo_env *e, int x, int y) { /* LineMask.h:15-18 */
    if (x < 0 || x >= W_PX || y < 0 || y >= H_PX) return 0;
    return e->line_mask[(size_t)y * W_PX + x] != 0;
}

/* ------------------------------------------------------------------ car */
/* Car.cpp:9-40 */
static void car_update(car_t *c, float thr, float st, float dt) {
    c->acc = thr * MAX_ACC;
    const float target = st * MAX_STEER;
    c->steer += (target - c->steer) * 0.2f;
    if (thr == 0.0f) c->v *= 0.95f;
    c->v += c->acc * dt;
    if (c->v < 0.0f) c->v = 0.0f;
    if (c->v > MAX_SPEED) c->v = MAX_SPEED;
    if (fabsf(c->v) > 0.1f) {
        volatile float sa = c->steer;
        const float ang = (c->v / CAR_LENGTH) * tanf(sa);
        c->h += ang;
    }
    c->h = fmodf(c->h + PI_F, 2.0f * PI_F);
    if (c->h < 0) c->h += 2.0f * PI_F;
    c->h -= PI_F;
    float s, co; sc(c->h, &s, &co);
    c->x += c->v * co;
    c->y -= c->v * s;
}
/* Car.cpp:47-74 */
static void update_path_index(car_t *c, const route_t *r) {
    if (r->n == 0) { c->path_index = 0; return; }
    int start = c->path_index; if (start < 0) start = 0;
    int end = start + 50; if (end > r->n) end = r->n;
    float best = INFINITY; int bi = start;
    for (int i = start; i < end; ++i) {
        const float dx = r->px[i] - c->x, dy = r->py[i] - c->y;
        const float d = dx * dx + dy * dy;
        if (d < best) { best = d; bi = i; }
    }
    c->path_index = bi;
}
/* Car.cpp:86-103: world(lx,ly) = (x + lx cos - ly sin, y + lx sin + ly cos), (+-27, +-12) */
static void corners(const car_t *c, float cx[4], float cy[4]) {
    const float hx = CAR_WIDTH * 0.5f, hy = CAR_LENGTH * 0.5f;
    float s, co; sc(c->h, &s, &co);
    const float lx[4] = {hy, hy, -hy, -hy}, ly[4] = {hx, -hx, -hx, hx};
    for (int i = 0; i < 4; ++i) {
        cx[i] = c->x + lx[i] * co - ly[i] * s;
        cy[i] = c->y + lx[i] * s + ly[i] * co;
    }
}
/* Car.cpp:105-141: SAT over the 2+2 edge normals; touching counts as colliding */
static int collide(const car_t *a, const car_t *b) {
    float ax[4], ay[4], bx[4], by[4];
    corners(a, ax, ay); corners(b, bx, by);
    float s1, c1, s2, c2; sc(a->h, &s1, &c1); sc(b->h, &s2, &c2);
    const float ux[4] = {c1, -s1, c2, -s2}, uy[4] = {s1, c1, s2, c2};
    for (int k = 0; k < 4; ++k) {
        float mn1 = INFINITY, mx1 = -INFINITY, mn2 = INFINITY, mx2 = -INFINITY;
        for (int i = 0; i < 4; ++i) {
            const float p = ax[i] * ux[k] + ay[i] * uy[k];
            mn1 = fminf(mn1, p); mx1 = fmaxf(mx1, p);
        }
        for (int i = 0; i < 4; ++i) {
            const float p = bx[i] * ux[k] + by[i] * uy[k];
            mn2 = fminf(mn2, p); mx2 = fmaxf(mx2, p);
        }
        if (mx1 < mn2 || mx2 < mn1) return 0;
    }
    return 1;
}
/* Car.cpp:76-84 */
static void respawn(car_t *c, const route_t *r) {
    c->x = r->sx; c->y = r->sy; c->v = 0.0f; c->h = r->sh;
    c->alive = 1; c->path_index = 0; c->prev_dist = 0.0f; c->pa0 = c->pa1 = 0.0f; c->acc = 0.0f; c->steer = 0.0f;
}

/* ------------------------------------------------------------------ lidar */
/* Lidar.cpp:4-14 / IntersectionEnv.cpp:119-127 */
static void make_rel_angles(struct isxo_env *e) {
    const float fov = 360.0f;
    const float start = -fov * 0.5f;
    const float step = (e->rays > 1) ? (fov / (float)(e->rays - 1)) : 0.0f;
    for (int i = 0; i < e->rays; ++i) {
        const float deg = start + i * step;
        e->rel[i] = deg * PI_F / 180.0f;
    }
}
/* Lidar.cpp:16-90.  cars = all candidate obstacles, self_i = index of self in that list */
static void lidar_scan(int lanes, int rays, const float *rel, const car_t *self, const car_t *cars, int ncars,
                       const car_t *self_ptr, float *dist_out) {
    const float max_dist = 250.0f, step = 4.0f;
    for (int i = 0; i < rays; ++i) {
        const float ang = self->h + rel[i];
        float s, c; sc(ang, &s, &c);
        const float dx = c, dy = -s;
        int hit = 0; float fd = max_dist;
        for (float d = 0.0f; d < max_dist; d += step) {
            const int px = (int)(self->x + dx * d);
            const int py = (int)(self->y + dy * d);
            if (px < 0 || px >= W_PX || py < 0 || py >= H_PX) break;
            if (d > 0.0f && !on_road(lanes, (float)px, (float)py)) { hit = 1; fd = d; break; }
            if (d > 0.0f) {
                int col = 0;
                for (int k = 0; k < ncars; ++k) {
                    const car_t *o = &cars[k];
                    if (o == self_ptr) continue;
                    if (fabsf(o->x - self->x) < 1e-3f && fabsf(o->y - self->y) < 1e-3f && fabsf(o->h - self->h) < 1e-3f) continue;
                    float so, co; sc(o->h, &so, &co);
                    const float hl = CAR_LENGTH * 0.5f, hw = CAR_WIDTH * 0.5f;
                    const float ex = fabsf(co) * hl + fabsf(so) * hw;
                    const float ey = fabsf(so) * hl + fabsf(co) * hw;
                    if ((float)px >= o->x - ex && (float)px <= o->x + ex && (float)py >= o->y - ey && (float)py <= o->y + ey) { col = 1; break; }
                }
                if (col) { hit = 1; fd = d; break; }
            }
        }
        dist_out[i] = hit ? fd : max_dist;
    }
}

/* ------------------------------------------------------------------ neighbour ordering
 * IntersectionEnv.cpp:490 is std::sort with `a.dist < b.dist`; with exactly equal distances and more than 16
 * neighbours the order of the equal ones is whatever the library's introsort leaves.  The reference run here is
 * libstdc++ 13: bits/stl_algo.h:1918-1952 (__introsort_loop, depth 2*floor(log2 n), threshold 16, median of
 * first+1 / mid / last-1 swapped to first, __unguarded_partition, then __final_insertion_sort) and
 * bits/stl_heap.h:135-262,340-430 for the __partial_sort fallback.  Restated on (key, position) records;
 * pinned against the real std::sort by tests/test_oracle_vs_ref.py::test_neighbor_sort_*. */
typedef struct { float d; int pos; } nb_t;
static void nb_push_heap(nb_t *a, int hole, int top, nb_t v) {
    int parent = (hole - 1) / 2;
    while (hole > top && a[parent].d < v.d) { a[hole] = a[parent]; hole = parent; parent = (hole - 1) / 2; }
    a[hole] = v;
}
static void nb_adjust_heap(nb_t *a, int hole, int len, nb_t v) {
    const int top = hole; int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (a[second].d < a[second - 1].d) second--;
        a[hole] = a[second]; hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) { second = 2 * (second + 1); a[hole] = a[second - 1]; hole = second - 1; }
    nb_push_heap(a, hole, top, v);
}
static void nb_heapsort(nb_t *a, int len) {
    if (len >= 2) { int parent = (len - 2) / 2; for (;;) { nb_adjust_heap(a, parent, len, a[parent]); if (parent == 0) break; parent--; } }
    while (len > 1) { --len; nb_t v = a[len]; a[len] = a[0]; nb_adjust_heap(a, 0, len, v); }
}
static void nb_swap(nb_t *x, nb_t *y) { nb_t t = *x; *x = *y; *y = t; }
static nb_t *nb_partition_pivot(nb_t *first, nb_t *last) {
    nb_t *a = first + 1, *b = first + (last - first) / 2, *c = last - 1;
    if (a->d < b->d) { if (b->d < c->d) nb_swap(first, b); else if (a->d < c->d) nb_swap(first, c); else nb_swap(first, a); }
    else if (a->d < c->d) nb_swap(first, a);
    else if (b->d < c->d) nb_swap(first, c);
    else nb_swap(first, b);
    nb_t *lo = first + 1, *hi = last;
    for (;;) {
        while (lo->d < first->d) ++lo;
        --hi;
        while (first->d < hi->d) --hi;
        if (!(lo < hi)) return lo;
        nb_swap(lo, hi);
        ++lo;
    }
}
static int g_nb_heaps;
static void nb_introsort(nb_t *first, nb_t *last, int depth) {
    while (last - first > 16) {
        if (depth == 0) { nb_heapsort(first, (int)(last - first)); g_nb_heaps++; return; }
        --depth;
        nb_t *cut = nb_partition_pivot(first, last);
        nb_introsort(cut, last, depth);
        last = cut;
    }
}
static void nb_linear_insert(nb_t *last) {
    nb_t v = *last; nb_t *next = last - 1;
    while (v.d < next->d) { *last = *next; last = next; --next; }
    *last = v;
}
static void nb_insertion(nb_t *first, nb_t *last) {
    if (first == last) return;
    for (nb_t *i = first + 1; i != last; ++i) {
        if (i->d < first->d) { nb_t v = *i; memmove(first + 1, first, (size_t)(i - first) * sizeof(nb_t)); *first = v; }
        else nb_linear_insert(i);
    }
}
static void nb_sort(nb_t *a, int n) {
    if (n == 0) return;
    int lg = 0; while ((n >> (lg + 1)) != 0) ++lg;
    nb_introsort(a, a + n, 2 * lg);
    if (n > 16) { nb_insertion(a, a + 16); for (nb_t *i = a + 16; i != a + n; ++i) nb_linear_insert(i); }
    else nb_insertion(a, a + n);
}
int isxo_std_sort(const float *keys, int n, int32_t *perm_out) {
    nb_t *a = (nb_t *)malloc(sizeof(nb_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { a[i].d = keys[i]; a[i].pos = i; }
    g_nb_heaps = 0;
    nb_sort(a, n);
    for (int i = 0; i < n; ++i) perm_out[i] = a[i].pos;
    free(a);
    return g_nb_heaps;
}


/* ------------------------------------------------------------------ observations */
/* IntersectionEnv.cpp:418-520 */
static void observe(const struct isxo_env *e, float *obs) {
    memset(obs, 0, sizeof(float) * ISX_OBS_DIM * (size_t)e->n);
    for (int i = 0; i < e->n; ++i) {
        float *o = obs + (size_t)i * ISX_OBS_DIM;
        const car_t *c = &e->cars[i];
        if (!c->alive) continue;
        const route_t *r = &e->ego_routes[e->slot_route[i]];
        o[0] = c->x / (float)W_PX; o[1] = c->y / (float)H_PX; o[2] = c->v / MAX_SPEED; o[3] = c->h / PI_F;
        if (r->n > 0) {
            int ti = c->path_index + 10; if (ti > r->n - 1) ti = r->n - 1;
            const float dxd = r->px[ti] - c->x, dyd = r->py[ti] - c->y;
            o[4] = sqrtf(dxd * dxd + dyd * dyd) / (float)W_PX;
            volatile float ny = -dyd, nx = dxd;
            o[5] = wrap_angle(atan2f(ny, nx) - c->h) / PI_F;
        }
        /* neighbours: other alive egos then NPCs, ordered as libstdc++ std::sort leaves them (nb_sort above) */
        float nd[MAX_AGENTS + MAX_NPC]; const car_t *nc[MAX_AGENTS + MAX_NPC]; int nint[MAX_AGENTS + MAX_NPC]; int m = 0;
        for (int j = 0; j < e->n; ++j) {
            if (j == i || !e->cars[j].alive) continue;
            const float dx = e->cars[j].x - c->x, dy = e->cars[j].y - c->y;
            nd[m] = sqrtf(dx * dx + dy * dy); nc[m] = &e->cars[j]; nint[m] = e->ego_routes[e->slot_route[j]].intent; m++;
        }
        if (e->traffic) for (int j = 0; j < e->n_npc; ++j) {
            const float dx = e->npcs[j].x - c->x, dy = e->npcs[j].y - c->y;
            nd[m] = sqrtf(dx * dx + dy * dy); nc[m] = &e->npcs[j]; nint[m] = e->traffic_routes[e->npcs[j].route].intent; m++;
        }
        nb_t ord[MAX_AGENTS + MAX_NPC];
        for (int a = 0; a < m; ++a) { ord[a].d = nd[a]; ord[a].pos = a; }
        nb_sort(ord, m);
        const int take = m < 5 ? m : 5;
        for (int k = 0; k < take; ++k) {
            float *q = o + 6 + 5 * k;
            q[0] = (nc[ord[k].pos]->x - c->x) / (float)W_PX;
            q[1] = (nc[ord[k].pos]->y - c->y) / (float)H_PX;
            q[2] = (nc[ord[k].pos]->v - c->v) / MAX_SPEED;
            q[3] = wrap_angle(nc[ord[k].pos]->h - c->h) / PI_F;
            q[4] = (float)nint[ord[k].pos];
        }
        const float inv = 1.0f / 250.0f; /* Lidar.cpp:92-97 */
        for (int k = 0; k < e->rays && 31 + k < ISX_OBS_DIM; ++k) o[31 + k] = e->lidar[i][k] * inv;
    }
}

/* ------------------------------------------------------------------ traffic flow */
static uint32_t next_u32(struct isxo_env *e) { return isx_traffic_word(e->seed, e->env_id, e->tick, e->draw++); }
/* libstdc++ 13 generate_canonical<float,24> over a 32-bit URBG (bits/random.tcc:3349-3381) */
static float u01(struct isxo_env *e) {
    float r = (float)next_u32(e) / 4294967296.0f;
    if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
    return r;
}
/* libstdc++ 13 uniform_int_distribution over a 32-bit URBG: Lemire (bits/uniform_int_dist.h:257-281,323-329) */
static uint32_t uint_below(struct isxo_env *e, uint32_t n) {
    uint64_t prod = (uint64_t)next_u32(e) * n;
    uint32_t low = (uint32_t)prod;
    if (low < n) {
        const uint32_t thr = (0u - n) % n;
        while (low < thr) { prod = (uint64_t)next_u32(e) * n; low = (uint32_t)prod; }
    }
    return (uint32_t)(prod >> 32);
}

/* TrafficFlow.cpp:22-47 */
static float front_dist(const struct isxo_env *e, int self) {
    const car_t *s = &e->npcs[self];
    float mn = 1e9f, sn, cs; sc(s->h, &sn, &cs);
    const float vx = cs, vy = -sn;
    for (int j = 0; j < e->n_npc; ++j) {
        if (j == self) continue;
        const car_t *o = &e->npcs[j];
        const float dx = o->x - s->x, dy = o->y - s->y;
        const float d = hypotf(dx, dy);
        if (d > 80.0f) continue;
        const float dot = (dx * vx + dy * vy) / (d + 1e-5f);
        if (dot > 0.8f) {
            const float ad = fabsf(wrap_angle(s->h - o->h));
            if (ad < (45.0f * PI_F / 180.0f)) { if (d < mn) mn = d; }
        }
    }
    return mn;
}

/* TrafficFlow.cpp:49-196 */
static void plan_npc(const struct isxo_env *e, int self, float *thr_out, float *steer_out) {
    const car_t *c = &e->npcs[self];
    const route_t *r = &e->traffic_routes[c->route];
    float steer = 0.0f;
    if (r->n > 0) {
        int ti = c->path_index + 12; if (ti > r->n - 1) ti = r->n - 1;
        const float dx = r->px[ti] - c->x, dy = r->py[ti] - c->y;
        volatile float ny = -dy, nx = dx;
        const float err = wrap_angle(atan2f(ny, nx) - c->h);
        steer = fmaxf(-1.0f, fminf(1.0f, err * 3.0f));
    }
    const float target = MAX_SPEED * 0.4f;
    float thr = 0.0f;
    if (c->v < target) thr = 0.5f;
    else if (c->v > target + 1.0f) thr = -0.1f;
    const float fd = front_dist(e, self);
    if (fd < 30.0f) thr = -1.0f;
    else if (fd < 50.0f) thr = fminf(thr, -0.2f);

    int conflict = 0; float min_conf = 1e9f;
    const float safe_sq = (CAR_WIDTH * 2.0f) * (CAR_WIDTH * 2.0f);
    const float my_dc = hypotf(c->x - W_PX * 0.5f, c->y - H_PX * 0.5f);
    const int s0 = c->path_index;
    int s1 = s0 + 120; if (s1 > r->n) s1 = r->n;
    for (int i = s0; i < s1; ++i) {
        const float gx = r->px[i], gy = r->py[i];
        for (int j = 0; j < e->n_npc; ++j) {
            if (j == self) continue;
            const car_t *o = &e->npcs[j];
            const float dxo = o->x - gx, dyo = o->y - gy;
            if (!(dxo * dxo + dyo * dyo < safe_sq)) continue;
            const float ad = fabsf(wrap_angle(c->h - o->h));
            if (ad < (60.0f * PI_F / 180.0f)) continue;
            { /* side-by-side exclusion, :107-159 */
                const float dxt = o->x - c->x, dyt = o->y - c->y;
                const float dto = hypotf(dxt, dyt);
                int skip = 0;
                if (dto > 1e-5f) {
                    float sn, cs; sc(c->h, &sn, &cs);
                    const float mx = cs, my = -sn;
                    const float adn = fminf(ad, 2.0f * PI_F - ad);
                    const int par = (adn < (30.0f * PI_F / 180.0f)) || (adn > (150.0f * PI_F / 180.0f));
                    if (par) {
                        const float lon = dxt * mx + dyt * my;
                        const float latsq = fmaxf(0.0f, dto * dto - lon * lon);
                        const float lat = sqrtf(latsq);
                        if (fabsf(lat) < (LANE_W * 1.5f) && fabsf(lon) < (CAR_LENGTH * 2.0f)) {
                            const float fdist = 20.0f;
                            const float mfx = c->x + mx * fdist, mfy = c->y + my * fdist;
                            float so, co; sc(o->h, &so, &co);
                            const float ofx = o->x + co * fdist, ofy = o->y + (-so) * fdist;
                            const float fdx = ofx - mfx, fdy = ofy - mfy;
                            const float fmag = hypotf(fdx, fdy);
                            if (fmag > 1e-5f) {
                                const float flon = fdx * mx + fdy * my;
                                const float flatsq = fmaxf(0.0f, fmag * fmag - flon * flon);
                                const float flat = sqrtf(flatsq);
                                if (fabsf(flat - lat) < (LANE_W * 0.5f)) skip = 1;
                            }
                        }
                    }
                }
                if (skip) continue;
            }
            int yield = 0;
            const float o_dc = hypotf(o->x - W_PX * 0.5f, o->y - H_PX * 0.5f);
            const float dtc = hypotf(gx - c->x, gy - c->y);
            if (dtc < 15.0f) yield = 1;
            else if (c->v < 1.0f && o->v > 3.0f && o_dc < my_dc + 25.0f) yield = 1;
            else if (o_dc < my_dc - 5.0f) yield = 1;
            else if (fabsf(o_dc - my_dc) <= 5.0f) { if (self < j) yield = 1; } /* address order == list order, :173 */
            if (yield) { conflict = 1; if (dtc < min_conf) min_conf = dtc; }
        }
        if (conflict) break;
    }
    float fin = thr;
    if (conflict) {
        if (min_conf < 35.0f) fin = -1.0f;
        else if (min_conf < 60.0f) fin = -0.8f;
        else fin = fminf(fin, 0.0f);
    }
    *thr_out = fin; *steer_out = steer;
}

/* TrafficFlow.cpp:240-259, 275-315 */
static void try_spawn(struct isxo_env *e) {
    if (e->n_traffic_routes == 0) return;
    const uint32_t ri = uint_below(e, (uint32_t)e->n_traffic_routes);
    e->ev.spawn_route = (int32_t)ri;
    const route_t *r = &e->traffic_routes[ri];
    if (r->n == 0) return; /* unknown start id */
    const float md = CAR_LENGTH * 2.5f, md2 = md * md;
    for (int i = 0; i < e->n; ++i) {
        const float dx = e->cars[i].x - r->sx, dy = e->cars[i].y - r->sy;
        if (dx * dx + dy * dy < md2) return;
    }
    for (int i = 0; i < e->n_npc; ++i) {
        const float dx = e->npcs[i].x - r->sx, dy = e->npcs[i].y - r->sy;
        if (dx * dx + dy * dy < md2) return;
    }
    if (e->n_npc >= MAX_NPC) { fprintf(stderr, "isx_oracle: NPC capacity exceeded\n"); abort(); }
    car_t *c = &e->npcs[e->n_npc++];
    memset(c, 0, sizeof *c);
    c->x = r->sx; c->y = r->sy; c->v = 0.0f; c->h = r->sh; c->alive = 1; c->route = (int)ri; c->uid = e->next_uid++;
    e->ev.spawned = 1;
}

/* TrafficFlow.cpp:317-367 */
static void update_traffic(struct isxo_env *e, float dt) {
    volatile float arg = -e->density * dt;
    const float p = 1.0f - expf(arg);
    if (u01(e) < p) try_spawn(e);
    for (int i = 0; i < e->n_npc; ++i) {
        car_t *c = &e->npcs[i];
        const route_t *r = &e->traffic_routes[c->route];
        update_path_index(c, r);
        float th, st; plan_npc(e, i, &th, &st);
        car_update(c, th, st, dt);
        update_path_index(c, r);
    }
    for (int i = 0; i < e->n_npc; ++i) {
        if (!e->npcs[i].alive) continue;
        for (int j = i + 1; j < e->n_npc; ++j) {
            if (!e->npcs[j].alive) continue;
            if (collide(&e->npcs[i], &e->npcs[j])) { e->npcs[i].alive = 0; e->npcs[j].alive = 0; }
        }
    }
    int w = 0;
    for (int i = 0; i < e->n_npc; ++i) {
        const car_t *c = &e->npcs[i];
        const route_t *r = &e->traffic_routes[c->route];
        const int arrived = (r->n > 0) && hypotf(c->x - r->px[r->n - 1], c->y - r->py[r->n - 1]) < 20.0f;
        const int oos = c->x < -100.0f || c->x > (float)W_PX + 100.0f || c->y < -100.0f || c->y > (float)H_PX + 100.0f;
        if (!c->alive || arrived || oos) {
            e->ev.removed_mask |= (1u << i);
            if (!c->alive) e->ev.collided_mask |= (1u << i);
            continue;
        }
        if (w != i) e->npcs[w] = *c;
        w++;
    }
    e->n_npc = w;
}

/* ------------------------------------------------------------------ step */
/* IntersectionEnv.cpp:133-392 */
static int do_step(struct isxo_env *e, const float *thr, const float *st, int n_act, float dt,
                   float *reward, int32_t *done, int32_t *status, int32_t *terminated, int32_t *truncated,
                   int32_t *agents_alive) {
    const int n = e->n;
    e->step_count++;
    e->tick++; e->draw = 0;
    memset(&e->ev, 0, sizeof e->ev); e->ev.spawn_route = -1;
    if (e->traffic) update_traffic(e, dt);
    e->ev.rng_draws = (int32_t)e->draw; e->ev.npc_count = e->n_npc;

    float rew[MAX_AGENTS]; int dn[MAX_AGENTS], stt[MAX_AGENTS];
    for (int i = 0; i < n; ++i) { rew[i] = 0.0f; dn[i] = 0; stt[i] = ISX_ALIVE; }
    volatile float w750 = (float)W_PX, h750 = (float)H_PX;
    const float max_prog = hypotf(w750, h750);

    for (int i = 0; i < n; ++i) { /* :151-163 */
        car_t *c = &e->cars[i];
        if (!c->alive) continue;
        const route_t *r = &e->ego_routes[e->slot_route[i]];
        car_update(c, i < n_act ? thr[i] : 0.0f, i < n_act ? st[i] : 0.0f, dt);
        update_path_index(c, r);
        float rp = 0.0f; /* :15-28 */
        if (r->n > 0) {
            const float cur = hypotf(c->x - r->px[r->n - 1], c->y - r->py[r->n - 1]);
            if (c->prev_dist > 0.0f) {
                const float prog = c->prev_dist - cur;
                const float norm = (max_prog > 0.0f) ? (prog / max_prog) : 0.0f;
                rp = e->k_prog * norm;
            }
            c->prev_dist = cur;
        }
        const float ms = (c->v * FPS_F) / SCALE_F; /* :30-33 */
        const float rs = (ms < e->v_min) ? e->k_stuck : 0.0f;
        const float an = c->acc / MAX_ACC, sn = c->steer / MAX_STEER; /* :35-46 */
        const float d0 = an - c->pa0, d1 = sn - c->pa1;
        const float rsm = e->k_sm * (d0 * d0 + d1 * d1);
        c->pa0 = an; c->pa1 = sn;
        rew[i] = rp + rs + rsm;
    }
    for (int i = 0; i < n; ++i) { /* :166-290 */
        const car_t *c = &e->cars[i];
        if (!c->alive) { dn[i] = 1; stt[i] = ISX_DEAD; continue; }
        const route_t *r = &e->ego_routes[e->slot_route[i]];
        int d = 0, s = ISX_ALIVE;
        if (r->n >= 2) {
            const float ex = r->px[r->n - 1], ey = r->py[r->n - 1];
            const float dxr = ex - r->px[r->n - 2], dyr = ey - r->py[r->n - 2];
            int ok;
            if (fabsf(dxr) > fabsf(dyr)) ok = fabsf(c->y - ey) < 15.0f && fabsf(c->x - ex) < 40.0f;
            else ok = fabsf(c->x - ex) < 15.0f && fabsf(c->y - ey) < 40.0f;
            if (ok) { d = 1; s = ISX_SUCCESS; }
        }
        if (!d) {
            float cx[4], cy[4]; corners(c, cx, cy);
            int oos = 0;
            for (int k = 0; k < 4; ++k)
                if (cx[k] < -100.0f || cx[k] > (float)W_PX + 100.0f || cy[k] < -100.0f || cy[k] > (float)H_PX + 100.0f) { oos = 1; break; }
            if (oos) { d = 1; s = ISX_CRASH_WALL; }
            else {
                int off = 0;
                for (int k = 0; k < 4; ++k) if (!on_road(e->lanes, cx[k], cy[k])) { off = 1; break; }
                if (off) { d = 1; s = ISX_CRASH_WALL; }
                else {
                    int hl = 0;
                    for (int k = 0; k < 4; ++k) if (hits_yellow(e->lanes, cx[k], cy[k])) { hl = 1; break; }
                    if (!hl) {
                        for (int k = 0; k < 4; ++k) {
                            const int k2 = (k + 1) & 3;
                            const float mx = 0.5f * (cx[k] + cx[k2]), my = 0.5f * (cy[k] + cy[k2]);
                            if (is_line(e, (int)mx, (int)my)) { hl = 1; break; }
                        }
                    }
                    if (!hl) for (int k = 0; k < 4; ++k) if (is_line(e, (int)cx[k], (int)cy[k])) { hl = 1; break; }
                    if (hl) { d = 1; s = ISX_CRASH_LINE; }
                }
            }
        }
        dn[i] = d; stt[i] = s;
    }
    for (int i = 0; i < n; ++i) { /* :293-318 */
        if (!e->cars[i].alive || dn[i]) continue;
        for (int j = i + 1; j < n; ++j) {
            if (!e->cars[j].alive || dn[j]) continue;
            if (collide(&e->cars[i], &e->cars[j])) { dn[i] = dn[j] = 1; stt[i] = stt[j] = ISX_CRASH_CAR; }
        }
        if (e->traffic) for (int j = 0; j < e->n_npc; ++j) {
            if (collide(&e->cars[i], &e->npcs[j])) { dn[i] = 1; stt[i] = ISX_CRASH_CAR; break; }
        }
    }
    for (int i = 0; i < n; ++i) { /* :321-326 */
        if (!dn[i]) continue;
        if (stt[i] == ISX_CRASH_CAR) rew[i] += e->k_cv;
        else if (stt[i] == ISX_CRASH_WALL || stt[i] == ISX_CRASH_LINE) rew[i] += e->k_co;
        else if (stt[i] == ISX_SUCCESS) rew[i] += e->k_succ;
    }
    if (e->use_team && n > 0) { /* :329-336 */
        float avg = 0.0f;
        for (int i = 0; i < n; ++i) avg += rew[i];
        avg /= (float)n;
        for (int i = 0; i < n; ++i) rew[i] = (1.0f - e->alpha) * rew[i] + e->alpha * avg;
    }
    int term = 0;
    if (e->respawn) { /* :339-368 */
        for (int i = 0; i < n; ++i) {
            if (!e->cars[i].alive || !dn[i]) continue;
            if (stt[i] == ISX_CRASH_CAR || stt[i] == ISX_CRASH_WALL || stt[i] == ISX_CRASH_LINE)
                respawn(&e->cars[i], &e->ego_routes[e->slot_route[i]]);
        }
    } else {
        for (int i = 0; i < n; ++i) if (dn[i]) { term = 1; break; }
    }
    int alive = 0, succ = 0;
    for (int i = 0; i < n; ++i) {
        if (!e->cars[i].alive) continue;
        alive++;
        if (dn[i] && stt[i] == ISX_SUCCESS) succ++;
    }
    if (e->respawn && succ > 0 && succ == alive) term = 1;
    const int trunc = (e->max_steps > 0 && e->step_count >= e->max_steps);

    car_t all[MAX_AGENTS + MAX_NPC]; int na = 0; /* :374-388 */
    for (int i = 0; i < n; ++i) all[na++] = e->cars[i];
    if (e->traffic) for (int j = 0; j < e->n_npc; ++j) all[na++] = e->npcs[j];
    for (int i = 0; i < n; ++i) {
        if (!e->cars[i].alive) continue;
        /* the reference passes a COPY of self inside the list in traffic mode; the 1e-3 pose test
         * (Lidar.cpp:58-63) is what skips it there, pointer identity otherwise — same outcome */
        lidar_scan(e->lanes, e->rays, e->rel, &e->cars[i], all, na, e->traffic ? NULL : &all[i], e->lidar[i]);
    }
    for (int i = 0; i < n; ++i) {
        if (reward) reward[i] = rew[i];
        if (done) done[i] = dn[i];
        if (status) status[i] = stt[i];
    }
    if (terminated) *terminated = term;
    if (truncated) *truncated = trunc;
    if (agents_alive) *agents_alive = alive;
    return e->step_count;
}

/* ================================================================== C ABI (mirrors ref_driver.cpp) */
void *isxo_create(int lanes) {
    struct isxo_env *e = (struct isxo_env *)calloc(1, sizeof *e);
    e->lanes = lanes; e->respawn = 1; e->max_steps = 2000; e->rays = 96; e->density = 0.5f;
    e->k_prog = 10.0f; e->v_min = 1.0f; e->k_stuck = -0.01f; e->k_cv = -10.0f; e->k_co = -5.0f; e->k_succ = 10.0f;
    e->k_sm = -0.02f; e->alpha = 0.2f;
    e->line_mask = make_line_mask(lanes);
    e->next_uid = 1;
    make_rel_angles(e);
    /* default traffic routes, TrafficFlow.cpp:198-238: per direction, per in-lane j: straight and left to out-lane j */
    static const int opp[4] = {2, 3, 0, 1}, left[4] = {1, 2, 3, 0};
    for (int d = 0; d < 4; ++d) for (int j = 0; j < lanes; ++j) {
        char s[16], o[16];
        snprintf(s, sizeof s, "IN_%d", d * lanes + j + 1);
        snprintf(o, sizeof o, "OUT_%d", opp[d] * lanes + j + 1);
        build_route(lanes, s, o, &e->traffic_routes[e->n_traffic_routes++]);
        snprintf(o, sizeof o, "OUT_%d", left[d] * lanes + j + 1);
        build_route(lanes, s, o, &e->traffic_routes[e->n_traffic_routes++]);
    }
    return e;
}
void isxo_destroy(void *h) { struct isxo_env *e = h; if (e) { free(e->line_mask); free(e); } }
void isxo_configure(void *h, int use_team, int respawn, int max_steps) {
    struct isxo_env *e = h; e->use_team = use_team != 0; e->respawn = respawn != 0; e->max_steps = max_steps;
}
void isxo_configure_traffic(void *h, int enabled, float density) {
    struct isxo_env *e = h; e->traffic = enabled != 0; e->density = density < 0.0f ? 0.0f : density;
}
void isxo_configure_routes(void *h, int n, const char *const *starts, const char *const *ends) {
    struct isxo_env *e = h; e->n_traffic_routes = 0;
    for (int i = 0; i < n && i < MAX_ROUTES; ++i) build_route(e->lanes, starts[i], ends[i], &e->traffic_routes[e->n_traffic_routes++]);
}
void isxo_set_reward(void *h, const float *k) {
    struct isxo_env *e = h;
    e->k_prog = k[0]; e->v_min = k[1]; e->k_stuck = k[2]; e->k_cv = k[3]; e->k_co = k[4]; e->k_succ = k[5]; e->k_sm = k[6]; e->alpha = k[7];
}
static int g_route_err;
void isxo_set_ego_routes(void *h, int n, const char *const *starts, const char *const *ends) {
    struct isxo_env *e = h; e->n_ego_routes = 0; g_route_err = 0;
    for (int i = 0; i < n && i < MAX_AGENTS; ++i) {
        int rc = build_route(e->lanes, starts[i], ends[i], &e->ego_routes[e->n_ego_routes++]);
        if (rc == -2) { e->ego_routes[e->n_ego_routes - 1].n = -1; } /* unknown end: reset() reports it */
    }
}
void isxo_set_lidar_rays(void *h, int rays) { struct isxo_env *e = h; e->rays = rays; make_rel_angles(e); }
/* IntersectionEnv.cpp:66-131 driven as env.py:147-152 does */
int isxo_reset(void *h) {
    struct isxo_env *e = h;
    e->n = 0; e->n_npc = 0; e->step_count = 0; e->next_uid = 1;
    for (int i = 0; i < e->n_ego_routes; ++i) {
        const route_t *r = &e->ego_routes[i];
        if (r->n == 0) continue;  /* unknown start: silent no-op, :79-82 */
        if (r->n < 0) return -1;  /* unknown end: std::out_of_range, RouteGen.cpp:120 */
        car_t *c = &e->cars[e->n];
        memset(c, 0, sizeof *c);
        respawn(c, r);
        e->slot_route[e->n] = i; c->route = i;
        for (int k = 0; k < e->rays; ++k) e->lidar[e->n][k] = 250.0f;
        e->n++;
    }
    return e->n;
}
void isxo_seed(void *h, uint64_t seed, uint32_t env_id, uint32_t tick) { struct isxo_env *e = h; e->seed = seed; e->env_id = env_id; e->tick = tick; }
uint32_t isxo_tick(void *h) { return ((struct isxo_env *)h)->tick; }
int isxo_num_agents(void *h) { return ((struct isxo_env *)h)->n; }
int isxo_num_npcs(void *h) { return ((struct isxo_env *)h)->n_npc; }
int isxo_step_count(void *h) { return ((struct isxo_env *)h)->step_count; }
void isxo_set_step_count(void *h, int s) { ((struct isxo_env *)h)->step_count = s; }
void isxo_get_obs(void *h, float *obs) { observe(h, obs); }
int isxo_step(void *h, const float *thr, const float *st, int n_act, float dt, float *obs, float *reward, int32_t *done,
              int32_t *status, int32_t *terminated, int32_t *truncated, int32_t *agents_alive) {
    struct isxo_env *e = h;
    const int s = do_step(e, thr, st, n_act, dt, reward, done, status, terminated, truncated, agents_alive);
    float tmp[MAX_AGENTS * ISX_OBS_DIM];
    observe(e, obs ? obs : tmp); /* obs assembly is part of the reference's step (:390) */
    return s;
}
void isxo_get_events(void *h, isx_traffic_events *ev) { *ev = ((struct isxo_env *)h)->ev; }
static void fill(const struct isxo_env *e, const car_t *c, int is_npc, isx_car_state *o) {
    o->x = c->x; o->y = c->y; o->v = c->v; o->heading = c->h; o->acc = c->acc; o->steer = c->steer;
    o->prev_dist = c->prev_dist; o->prev_a0 = c->pa0; o->prev_a1 = c->pa1; o->path_index = c->path_index;
    o->route = c->route; o->alive = c->alive; o->uid = c->uid;
    o->intention = is_npc ? e->traffic_routes[c->route].intent : e->ego_routes[c->route].intent;
}
int isxo_get_egos(void *h, isx_car_state *out) { struct isxo_env *e = h; for (int i = 0; i < e->n; ++i) fill(e, &e->cars[i], 0, out + i); return e->n; }
int isxo_get_npcs(void *h, isx_car_state *out, int cap) {
    struct isxo_env *e = h;
    for (int i = 0; i < e->n_npc && i < cap; ++i) fill(e, &e->npcs[i], 1, out + i);
    return e->n_npc;
}
int isxo_get_lidar(void *h, int agent, float *dist, int cap) {
    struct isxo_env *e = h;
    for (int i = 0; i < e->rays && i < cap; ++i) dist[i] = e->lidar[agent][i];
    return e->rays;
}
void isxo_set_egos(void *h, const isx_car_state *s, int n) {
    struct isxo_env *e = h;
    for (int i = 0; i < n && i < e->n; ++i) {
        car_t *c = &e->cars[i];
        c->x = s[i].x; c->y = s[i].y; c->v = s[i].v; c->h = s[i].heading; c->acc = s[i].acc; c->steer = s[i].steer;
        c->prev_dist = s[i].prev_dist; c->pa0 = s[i].prev_a0; c->pa1 = s[i].prev_a1; c->path_index = s[i].path_index;
        c->alive = s[i].alive != 0;
    }
}
void isxo_set_npcs(void *h, const isx_car_state *s, int n) {
    struct isxo_env *e = h; e->n_npc = 0;
    for (int i = 0; i < n && i < MAX_NPC; ++i) {
        car_t *c = &e->npcs[e->n_npc++];
        memset(c, 0, sizeof *c);
        c->x = s[i].x; c->y = s[i].y; c->v = s[i].v; c->h = s[i].heading; c->acc = s[i].acc; c->steer = s[i].steer;
        c->path_index = s[i].path_index; c->alive = 1; c->route = s[i].route; c->uid = s[i].uid;
        if (c->uid >= e->next_uid) e->next_uid = c->uid + 1;
    }
}
long long isxo_rollout(void *h, int steps, float dt, int32_t *hist6, double *reward_sum) {
    struct isxo_env *e = h;
    float th[MAX_AGENTS], st[MAX_AGENTS], rew[MAX_AGENTS]; int32_t dn[MAX_AGENTS], stt[MAX_AGENTS];
    long long total = 0;
    for (int s = 0; s < steps; ++s) {
        for (int a = 0; a < e->n; ++a) isx_action_for(e->seed, e->env_id, e->tick + 1, (uint32_t)a, &th[a], &st[a]);
        int32_t term = 0, trunc = 0, alive = 0;
        isxo_step(h, th, st, e->n, dt, NULL, rew, dn, stt, &term, &trunc, &alive);
        for (int a = 0; a < e->n; ++a) { if (hist6) hist6[stt[a]]++; if (reward_sum) *reward_sum += (double)rew[a]; }
        total += e->n;
        if (term || trunc) isxo_reset(h);
    }
    return total;
}

/* ---- unit probes ---- */
int isxo_route(int lanes, const char *start, const char *end, float *path_xy, int *intent, float *sx, float *sy, float *sh) {
    route_t r; int rc = build_route(lanes, start, end, &r);
    if (rc) return rc;
    for (int i = 0; i < r.n; ++i) { path_xy[2 * i] = r.px[i]; path_xy[2 * i + 1] = r.py[i]; }
    *intent = r.intent; *sx = r.sx; *sy = r.sy; *sh = r.sh;
    return r.n;
}
int isxo_lane_point(int lanes, const char *id, float *x, float *y) { int d; return lane_lookup(lanes, id, x, y, &d) ? 0 : -1; }
int isxo_on_road(int lanes, float x, float y) { return on_road(lanes, x, y); }
int isxo_yellow(int lanes, float x, float y) { return hits_yellow(lanes, x, y); }
int isxo_is_line(int lanes, int x, int y) {
    struct isxo_env e; e.line_mask = make_line_mask(lanes);
    int r = is_line(&e, x, y); free(e.line_mask); return r;
}
void isxo_road_map(int lanes, uint8_t *out) {
    for (int y = 0; y < H_PX; ++y) for (int x = 0; x < W_PX; ++x) out[y * W_PX + x] = (uint8_t)on_road(lanes, (float)x, (float)y);
}
void isxo_line_map(int lanes, uint8_t *out) { uint8_t *g = make_line_mask(lanes); memcpy(out, g, (size_t)W_PX * H_PX); free(g); }
void isxo_car_update(float *s, float thr, float st, float dt) {
    car_t c; memset(&c, 0, sizeof c);
    c.x = s[0]; c.y = s[1]; c.v = s[2]; c.h = s[3]; c.acc = s[4]; c.steer = s[5];
    car_update(&c, thr, st, dt);
    s[0] = c.x; s[1] = c.y; s[2] = c.v; s[3] = c.h; s[4] = c.acc; s[5] = c.steer;
}
int isxo_collide(const float *a, const float *b) {
    car_t c1, c2; memset(&c1, 0, sizeof c1); memset(&c2, 0, sizeof c2);
    c1.x = a[0]; c1.y = a[1]; c1.h = a[2]; c2.x = b[0]; c2.y = b[1]; c2.h = b[2];
    return collide(&c1, &c2);
}
void isxo_corners(const float *a, float *out8) {
    car_t c; memset(&c, 0, sizeof c); c.x = a[0]; c.y = a[1]; c.h = a[2];
    float cx[4], cy[4]; corners(&c, cx, cy);
    for (int i = 0; i < 4; ++i) { out8[2 * i] = cx[i]; out8[2 * i + 1] = cy[i]; }
}
void isxo_lidar(int lanes, int rays, const float *self_pose, const float *others, int n_others, float *dist) {
    struct isxo_env e; e.rays = rays; make_rel_angles(&e);
    car_t cars[1 + MAX_AGENTS + MAX_NPC]; memset(cars, 0, sizeof cars);
    cars[0].x = self_pose[0]; cars[0].y = self_pose[1]; cars[0].h = self_pose[2];
    int n = 1;
    for (int i = 0; i < n_others && n < 1 + MAX_AGENTS + MAX_NPC; ++i) { cars[n].x = others[3 * i]; cars[n].y = others[3 * i + 1]; cars[n].h = others[3 * i + 2]; n++; }
    lidar_scan(lanes, rays, e.rel, &cars[0], cars, n, &cars[0], dist);
}
void isxo_libm_sincosf(const float *x, int n, float *s, float *c) { for (int i = 0; i < n; ++i) sc(x[i], s + i, c + i); }
void isxo_libm_tanf(const float *x, int n, float *o) { for (int i = 0; i < n; ++i) { volatile float v = x[i]; o[i] = tanf(v); } }
void isxo_libm_atan2f(const float *y, const float *x, int n, float *o) { for (int i = 0; i < n; ++i) { volatile float a = y[i], b = x[i]; o[i] = atan2f(a, b); } }
void isxo_libm_hypotf(const float *y, const float *x, int n, float *o) { for (int i = 0; i < n; ++i) { volatile float a = y[i], b = x[i]; o[i] = hypotf(a, b); } }
void isxo_libm_fmodf(const float *y, const float *x, int n, float *o) { for (int i = 0; i < n; ++i) { volatile float a = y[i], b = x[i]; o[i] = fmodf(a, b); } }
