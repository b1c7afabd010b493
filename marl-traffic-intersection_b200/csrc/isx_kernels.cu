// isx_kernels.cu — the step kernels (sm_100a).
//
//   k_dynamics  : one WARP per env.  NPC traffic flow (spawn draw, Gauss-Seidel planner/integrator in list
//                 order, NPC-NPC SAT, ordered erase) with the warp's lanes spread over ghost-path points /
//                 other cars / path-window points; then ego physics, reward, status, car-car override,
//                 bonuses, team mix, respawn and termination with one LANE per ego.
//                 Replaces TrafficFlow.cpp:317-367 and IntersectionEnv.cpp:137-370.
//   k_lidar_obs : persistent CTAs; folded road bitmap + skip table staged in shared memory once per CTA;
//                 one THREAD per (ego, beam): sphere-traced road march + slab/verify against the other cars'
//                 pixel rectangles; then the 31 ego/neighbour features.  Writes obs rows coalesced.
//                 Replaces Lidar.cpp:16-90 and IntersectionEnv.cpp:374-390, 418-520.
//
// No tensor cores: nothing here is a dense contraction.  Build with -fmad=false (see isx_math.cuh).
#include <cuda_runtime.h>

#include "isx_device.cuh"

namespace isx {

constexpr unsigned FULL = 0xffffffffu;
constexpr int DYN_WARPS = 4;              // envs per CTA in k_dynamics
constexpr int LID_THREADS = 256;
constexpr int LID_AGENTS = 32;            // egos per CTA iteration in k_lidar_obs

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// Car::update_path_index with the 50-point window spread over the lanes; first minimum wins (Car.cpp:62-70).
__device__ __forceinline__ int warp_path_index(const F2* __restrict__ path, int idx, float x, float y, int lane) {
    const int start = idx < 0 ? 0 : idx;
    const int end = min(start + 50, PATH_LEN);
    float best = INFINITY;
    int bi = start;
    for (int i = start + lane; i < end; i += 32) {
        const F2 p = path[i];
        const float dx = p.x - x, dy = p.y - y;
        const float d = dx * dx + dy * dy;
        if (d < best) { best = d; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    return bi;
}

struct NpcSmem {
    float x[ISX_MAX_NPC], y[ISX_MAX_NPC], v[ISX_MAX_NPC], h[ISX_MAX_NPC], steer[ISX_MAX_NPC];
    int pidx[ISX_MAX_NPC], route[ISX_MAX_NPC];
    uint32_t uid[ISX_MAX_NPC];
    uint32_t coll[ISX_MAX_NPC];
};

// ------------------------------------------------------------------------------------------------ k_dynamics
__global__ void __launch_bounds__(DYN_WARPS * 32)
k_dynamics(const Dev d, const float* __restrict__ actions, float dt, float spawn_prob) {
    __shared__ NpcSmem sm_all[DYN_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * DYN_WARPS + warp;
    if (env >= d.E) return;                       // whole warp leaves; only warp-level sync below
    NpcSmem& sm = sm_all[warp];
    const int N = d.N;
    const uint32_t genv = (uint32_t)(d.env_base + env);
    const bool is_ego = lane < N;
    const int ai = env * N + (is_ego ? lane : 0);

    // ---- ego state in registers (lane = ego slot)
    Pose p{0, 0, 0, 0};
    float steer = 0, acc = 0, pd = 0, pa0 = 0, pa1 = 0;
    int pidx = 0;
    bool alive = false;
    if (is_ego) {
        p.x = d.ex[ai]; p.y = d.ey[ai]; p.v = d.ev[ai]; p.h = d.eh[ai];
        steer = d.esteer[ai]; acc = d.eacc[ai]; pd = d.epd[ai]; pa0 = d.epa0[ai]; pa1 = d.epa1[ai];
        pidx = d.epidx[ai]; alive = d.ealive[ai] != 0;
    }
    int step_count = d.step_count[env];
    int c = d.traffic ? d.ncount[env] : 0;
    uint32_t next_uid = d.next_uid[env];
    uint32_t resets = 0;

    // ---- auto-reset: what a caller of env.py does after terminated|truncated (reset(), env.py:147-152)
    if (d.auto_reset && (d.terminated[env] | d.truncated[env])) {
        if (is_ego) {
            const RouteMeta m = d.route_meta[lane];
            p.x = m.spawn_x; p.y = m.spawn_y; p.v = 0.0f; p.h = m.spawn_h;
            steer = 0; acc = 0; pd = 0; pa0 = 0; pa1 = 0; pidx = 0; alive = true;
        }
        step_count = 0; c = 0; next_uid = 1; resets = 1;
    }
    step_count += 1;                               // IntersectionEnv.cpp:137
    const uint32_t tick = d.tick[env] + 1;

    isx_traffic_events evt;
    evt.rng_draws = 0; evt.spawn_route = -1; evt.spawned = 0; evt.removed_mask = 0; evt.collided_mask = 0; evt.npc_count = 0;
    uint32_t overflow = 0;

    // ================================================================ traffic flow (TrafficFlow.cpp:317-367)
    if (d.traffic) {
        if (lane < c) {
            const int ni = env * d.M + lane;
            sm.x[lane] = d.nx[ni]; sm.y[lane] = d.ny[ni]; sm.v[lane] = d.nv[ni]; sm.h[lane] = d.nh[ni];
            sm.steer[lane] = d.nsteer[ni]; sm.pidx[lane] = d.npidx[ni]; sm.route[lane] = d.nroute[ni]; sm.uid[lane] = d.nuid[ni];
        }
        __syncwarp();
        // -- spawn draw (:321-329) and try_spawn_traffic_car (:275-315); every lane runs the same stream
        TrafficStream ts;
        ts.init(d.seed, genv, tick);
        if (ts.uniform01() < spawn_prob && d.T > 0) {
            const int r = (int)ts.below((uint32_t)d.T);
            evt.spawn_route = r;
            const RouteMeta m = d.route_meta[N + r];
            const float md = CAR_LENGTH * 2.5f, md2 = md * md;       // is_spawn_blocked (:240-259)
            bool blk = false;
            if (is_ego) { const float dx = p.x - m.spawn_x, dy = p.y - m.spawn_y; blk = dx * dx + dy * dy < md2; }
            if (lane < c) { const float dx = sm.x[lane] - m.spawn_x, dy = sm.y[lane] - m.spawn_y; blk = blk || (dx * dx + dy * dy < md2); }
            if (!__any_sync(FULL, blk)) {
                if (c < d.M) {
                    if (lane == 0) {
                        sm.x[c] = m.spawn_x; sm.y[c] = m.spawn_y; sm.v[c] = 0.0f; sm.h[c] = m.spawn_h; sm.steer[c] = 0.0f;
                        sm.pidx[c] = 0; sm.route[c] = r; sm.uid[c] = next_uid;
                    }
                    next_uid += 1; c += 1; evt.spawned = 1;
                    __syncwarp();
                } else overflow = 1;                                   // reference list is unbounded; counted
            }
        }
        evt.rng_draws = (int)ts.j;

        // -- NPC controller, sequential in list order: NPC i sees the already-updated state of NPCs < i (:337-344)
        for (int i = 0; i < c; ++i) {
            Pose me{sm.x[i], sm.y[i], sm.v[i], sm.h[i]};
            float msteer = sm.steer[i], macc = 0.0f;
            const F2* path = d.route_path + (size_t)(N + sm.route[i]) * PATH_LEN;
            int mp = warp_path_index(path, sm.pidx[i], me.x, me.y, lane);
            const float steer_cmd = npc_steer_cmd(me, path[min(mp + 12, PATH_LEN - 1)]);
            float ms, mc;
            sincosf_(me.h, &ms, &mc);
            const float me_dc = hypotf_(me.x - WIDTH * 0.5f, me.y - HEIGHT * 0.5f);
            float fc = 1e9f;
            int flags = 0;
            Pose ot{0, 0, 0, 0};
            if (lane < c && lane != i) {
                ot = Pose{sm.x[lane], sm.y[lane], sm.v[lane], sm.h[lane]};
                fc = npc_front_candidate(me, ot, ms, mc);
                flags = npc_pair_flags(me, ot, ms, mc, me_dc, i < lane);
            }
            const float thr0 = npc_cruise_throttle(me.v, warp_min_f(fc));
            const unsigned elig = __ballot_sync(FULL, flags & 1);
            const unsigned yld = __ballot_sync(FULL, flags & 2);
            bool conflict = false;
            float min_conf = 1e9f;
            if (elig) {                                                // ghost-path scan (:91-185), 32 points per pass
                const float safe_sq = (CAR_WIDTH * 2.0f) * (CAR_WIDTH * 2.0f);
                const int s1 = min(mp + 120, PATH_LEN);
                for (int base = mp; base < s1 && !conflict; base += 32) {
                    const int g = base + lane;
                    bool hit = false;
                    float dtc = 0.0f;
                    if (g < s1) {
                        const F2 gp = path[g];
                        unsigned near_yield = 0, near_any = 0;
                        for (unsigned m = elig; m; m &= m - 1) {
                            const int o = __ffs(m) - 1;
                            const float dx = sm.x[o] - gp.x, dy = sm.y[o] - gp.y;
                            if (dx * dx + dy * dy < safe_sq) { near_any = 1; near_yield |= (yld >> o) & 1u; }
                        }
                        if (near_any) {
                            dtc = hypotf_(gp.x - me.x, gp.y - me.y);
                            hit = near_yield || (dtc < 15.0f);
                        }
                    }
                    const unsigned hb = __ballot_sync(FULL, hit);
                    if (hb) { conflict = true; min_conf = __shfl_sync(FULL, dtc, __ffs(hb) - 1); }
                }
            }
            const float thr = npc_final_throttle(thr0, conflict, min_conf);
            car_update(me, msteer, macc, thr, steer_cmd, dt);
            mp = warp_path_index(path, mp, me.x, me.y, lane);
            __syncwarp();
            if (lane == 0) { sm.x[i] = me.x; sm.y[i] = me.y; sm.v[i] = me.v; sm.h[i] = me.h; sm.steer[i] = msteer; sm.pidx[i] = mp; }
            __syncwarp();
        }

        // -- NPC-NPC collisions (:347-356): lane j tests the pair (i, j), j > i
        unsigned alive_m = c >= 32 ? FULL : ((1u << c) - 1u);
        const unsigned all_m = alive_m;
        for (int i = 0; i + 1 < c; ++i) {
            bool hit = false;
            if (lane > i && lane < c) hit = cars_collide(sm.x[i], sm.y[i], sm.h[i], sm.x[lane], sm.y[lane], sm.h[lane]);
            const unsigned hm = __ballot_sync(FULL, hit);
            if ((alive_m >> i) & 1u) {
                const unsigned m = hm & alive_m;
                if (m) alive_m &= ~(m | (1u << i));
            }
        }
        // -- ordered erase of dead / arrived / out-of-screen (:359-366)
        bool rem = false;
        float mx = 0, my = 0, mv = 0, mh = 0, mst = 0; int mpi = 0, mr = 0; uint32_t mu = 0;
        if (lane < c) {
            mx = sm.x[lane]; my = sm.y[lane]; mv = sm.v[lane]; mh = sm.h[lane]; mst = sm.steer[lane];
            mpi = sm.pidx[lane]; mr = sm.route[lane]; mu = sm.uid[lane];
            const F2 goal = d.route_meta[N + mr].goal;
            const bool arrived = hypotf_(mx - goal.x, my - goal.y) < 20.0f;
            const bool oos = mx < -100.0f || mx > (float)WIDTH + 100.0f || my < -100.0f || my > (float)HEIGHT + 100.0f;
            rem = !((alive_m >> lane) & 1u) || arrived || oos;
        }
        const unsigned rem_m = __ballot_sync(FULL, rem);
        evt.removed_mask = rem_m;
        evt.collided_mask = all_m & ~alive_m;
        const unsigned keep_m = all_m & ~rem_m;
        __syncwarp();
        if (lane < c && !rem) {
            const int pos = __popc(keep_m & ((1u << lane) - 1u));
            sm.x[pos] = mx; sm.y[pos] = my; sm.v[pos] = mv; sm.h[pos] = mh; sm.steer[pos] = mst;
            sm.pidx[pos] = mpi; sm.route[pos] = mr; sm.uid[pos] = mu;
            const int ni = env * d.M + pos;
            d.nx[ni] = mx; d.ny[ni] = my; d.nv[ni] = mv; d.nh[ni] = mh; d.nsteer[ni] = mst;
            d.npidx[ni] = mpi; d.nroute[ni] = mr; d.nuid[ni] = mu;
        }
        c = __popc(keep_m);
        evt.npc_count = c;
        __syncwarp();
    }

    // ================================================================ egos (IntersectionEnv.cpp:144-370)
    float rew = 0.0f;
    int status = ISX_ALIVE;
    bool done = false;
    if (is_ego && alive) {                             // Car::update, :151-156
        float thr, st;
        if (actions) { thr = actions[2 * ai]; st = actions[2 * ai + 1]; }
        else philox_action(d.seed, genv, tick, (uint32_t)lane, thr, st);
        car_update(p, steer, acc, thr, st, dt);
    }
    {   // Car::update_path_index (:157) with the 50-point window of every ego spread over 32/NP lanes
        const int NP = N <= 1 ? 1 : N <= 2 ? 2 : N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32;   // lanes per "row" of egos
        const int rows = 32 / NP;
        const int ego = lane & (NP - 1), part = lane / NP;
        const float ex = __shfl_sync(FULL, p.x, ego), ey = __shfl_sync(FULL, p.y, ego);
        const int ep = __shfl_sync(FULL, pidx, ego);
        const int start = ep < 0 ? 0 : ep, end = min(start + 50, PATH_LEN);
        float best = INFINITY;
        int bi = start;
        if (ego < N) {
            const F2* path = d.route_path + (size_t)ego * PATH_LEN;
            for (int i = start + part; i < end; i += rows) {
                const F2 q = path[i];
                const float dx = q.x - ex, dy = q.y - ey;
                const float dd = dx * dx + dy * dy;
                if (dd < best) { best = dd; bi = i; }
            }
        }
        for (int o = NP; o < 32; o <<= 1) {            // first minimum wins: lexicographic (distance, index)
            const float ob = __shfl_xor_sync(FULL, best, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (is_ego && alive) pidx = bi;                // lanes 0..N-1 are part 0 of their own ego
    }
    if (is_ego) {
        if (alive) {                                   // :159-163
            const RouteMeta m = d.route_meta[lane];
            rew = reward_base(d.rc, p.x, p.y, p.v, acc, steer, m.goal, d.max_progress, pd, pa0, pa1);
            status = ego_self_status(d.lanes, p.x, p.y, p.h, m.goal, m.goal_prev);     // :166-290
            done = status != ISX_ALIVE;
        } else { status = ISX_DEAD; done = true; }
    }
    // -- car-car override (:293-318)
    unsigned cmask = 0;                                // bit j: ego `lane` collides with ego j > lane
    for (int dlt = 1; dlt < N; ++dlt) {
        const float ox = __shfl_down_sync(FULL, p.x, dlt), oy = __shfl_down_sync(FULL, p.y, dlt), oh = __shfl_down_sync(FULL, p.h, dlt);
        if (is_ego && lane + dlt < N && cars_collide(p.x, p.y, p.h, ox, oy, oh)) cmask |= 1u << (lane + dlt);
    }
    bool npc_hit = false;
    if (d.traffic && is_ego) {
        for (int k = 0; k < c && !npc_hit; ++k) npc_hit = cars_collide(p.x, p.y, p.h, sm.x[k], sm.y[k], sm.h[k]);
    }
    {
        const unsigned alive_m = __ballot_sync(FULL, is_ego && alive);
        unsigned done_m = __ballot_sync(FULL, is_ego && done);
        const unsigned npc_m = __ballot_sync(FULL, npc_hit);
        unsigned crash_m = 0;
        for (int i = 0; i < N; ++i) {
            const unsigned ci = __shfl_sync(FULL, cmask, i);
            if (!((alive_m >> i) & 1u) || ((done_m >> i) & 1u)) continue;
            const unsigned m = ci & alive_m & ~done_m;
            if (m) { done_m |= m | (1u << i); crash_m |= m | (1u << i); }
            if ((npc_m >> i) & 1u) { done_m |= 1u << i; crash_m |= 1u << i; }
        }
        if ((crash_m >> lane) & 1u) { done = true; status = ISX_CRASH_CAR; }
    }
    // -- terminal bonuses (:321-326)
    if (is_ego && done) {
        if (status == ISX_CRASH_CAR) rew += d.rc.k_cv;
        else if (status == ISX_CRASH_WALL || status == ISX_CRASH_LINE) rew += d.rc.k_co;
        else if (status == ISX_SUCCESS) rew += d.rc.k_succ;
    }
    // -- team mix (:329-336): sum in index order, then blend
    if (d.use_team && N > 0) {
        float avg = 0.0f;
        for (int i = 0; i < N; ++i) avg += __shfl_sync(FULL, rew, i);
        avg /= (float)N;
        rew = (1.0f - d.rc.alpha) * rew + d.rc.alpha * avg;
    }
    // -- respawn / termination (:339-370)
    const unsigned alive_m = __ballot_sync(FULL, is_ego && alive);
    const unsigned done_m = __ballot_sync(FULL, is_ego && done);
    const unsigned succ_m = __ballot_sync(FULL, is_ego && alive && done && status == ISX_SUCCESS);
    bool term = false;
    if (d.respawn) {
        if (is_ego && alive && done && (status == ISX_CRASH_CAR || status == ISX_CRASH_WALL || status == ISX_CRASH_LINE)) {
            const RouteMeta m = d.route_meta[lane];           // Car::respawn (Car.cpp:76-84)
            p.x = m.spawn_x; p.y = m.spawn_y; p.v = 0.0f; p.h = m.spawn_h;
            pidx = 0; pd = 0.0f; pa0 = 0.0f; pa1 = 0.0f; acc = 0.0f; steer = 0.0f;
        }
        const int na = __popc(alive_m), ns = __popc(succ_m);
        term = ns > 0 && ns == na;
    } else term = done_m != 0;
    const bool trunc = d.max_steps > 0 && step_count >= d.max_steps;

    // ---- write back
    if (is_ego) {
        d.ex[ai] = p.x; d.ey[ai] = p.y; d.ev[ai] = p.v; d.eh[ai] = p.h;
        d.esteer[ai] = steer; d.eacc[ai] = acc; d.epd[ai] = pd; d.epa0[ai] = pa0; d.epa1[ai] = pa1;
        d.epidx[ai] = pidx; d.ealive[ai] = alive ? 1 : 0;
        d.reward[ai] = rew; d.done[ai] = done ? 1 : 0; d.status[ai] = (uint8_t)status;
    }
    if (lane == 0) {
        d.step_count[env] = step_count; d.tick[env] = tick; d.next_uid[env] = next_uid;
        d.terminated[env] = term ? 1 : 0; d.truncated[env] = trunc ? 1 : 0; d.agents_alive[env] = __popc(alive_m);
        if (d.traffic) d.ncount[env] = c;
        d.events[env] = evt;
    }
    // ---- per-env counters (no atomics; reduced on demand by k_reduce_stats)
    {
        double rs = is_ego ? (double)rew : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(FULL, rs, o);
        uint32_t* st = d.env_stats + (size_t)env * STAT_SLOTS;
        unsigned hist[6];
#pragma unroll
        for (int s = 0; s < 6; ++s) hist[s] = __popc(__ballot_sync(FULL, is_ego && status == s));
        uint32_t inc = 0;
        if (lane < 6) inc = hist[lane];
        else if (lane == ST_SPAWNED) inc = (uint32_t)evt.spawned;
        else if (lane == ST_REMOVED) inc = (uint32_t)__popc(evt.removed_mask);
        else if (lane == ST_COLLIDED) inc = (uint32_t)__popc(evt.collided_mask);
        else if (lane == ST_OVERFLOW) inc = overflow;
        else if (lane == ST_RESETS) inc = resets;
        else if (lane == ST_STEPS) inc = (uint32_t)N;
        if (lane < 14) { if (inc) st[lane] += inc; }
        else if (lane == 14) { double* ps = reinterpret_cast<double*>(st + ST_RSUM); *ps += rs; }
    }
}

// ------------------------------------------------------------------------------------------------ k_lidar_obs
// Dynamic shared memory of k_lidar_obs, carved at run time: G envs per group, CE = N + M car slots per env.
struct LidSmem {
    uint32_t* bits;            // [ROAD_ROWS * ROAD_WORDS]
    float* rel;                // [ISX_MAX_RAYS]
    PixRect* rect;             // [G * CE]   lidar pixel rectangle of every car (egos then NPCs)
    int* ncand;                // [LID_AGENTS] per ego: number of cars a beam can possibly hit
    int* acb;                  // [LID_AGENTS] first car slot of the ego's env
    int* aself;                // [LID_AGENTS] ego's slot inside its env
    uint8_t *ccar, *cia, *cspan;   // [LID_AGENTS * CE] candidate car slot, first beam of its window, window span
    float *cx, *cy, *cv, *ch;  // [G * CE]
    int* cintent;              // [G * CE]
    int* ncars;                // [LID_AGENTS]
    uint8_t* skip;             // [SKIP_DIM * SKIP_DIM]
    uint8_t* ealive;           // [LID_AGENTS]
};
__host__ __device__ inline size_t lid_carve(unsigned char* base, int G, int CE, LidSmem* s) {
    size_t o = 0;
    const size_t nc = (size_t)G * CE;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
    const size_t o_bits = take(sizeof(uint32_t) * ROAD_ROWS * ROAD_WORDS);
    const size_t o_rel = take(sizeof(float) * ISX_MAX_RAYS);
    const size_t o_rect = take(sizeof(PixRect) * nc);
    const size_t o_ncand = take(sizeof(int) * LID_AGENTS), o_acb = take(sizeof(int) * LID_AGENTS), o_aself = take(sizeof(int) * LID_AGENTS);
    const size_t o_ccar = take((size_t)LID_AGENTS * CE), o_cia = take((size_t)LID_AGENTS * CE), o_cspan = take((size_t)LID_AGENTS * CE);
    const size_t o_cx = take(sizeof(float) * nc), o_cy = take(sizeof(float) * nc), o_cv = take(sizeof(float) * nc), o_ch = take(sizeof(float) * nc);
    const size_t o_int = take(sizeof(int) * nc);
    const size_t o_nc = take(sizeof(int) * LID_AGENTS);
    const size_t o_skip = take(SKIP_DIM * SKIP_DIM);
    const size_t o_alive = take(LID_AGENTS);
    if (s) {
        s->bits = reinterpret_cast<uint32_t*>(base + o_bits); s->rel = reinterpret_cast<float*>(base + o_rel);
        s->rect = reinterpret_cast<PixRect*>(base + o_rect);
        s->ncand = reinterpret_cast<int*>(base + o_ncand); s->acb = reinterpret_cast<int*>(base + o_acb); s->aself = reinterpret_cast<int*>(base + o_aself);
        s->ccar = base + o_ccar; s->cia = base + o_cia; s->cspan = base + o_cspan;
        s->cx = reinterpret_cast<float*>(base + o_cx); s->cy = reinterpret_cast<float*>(base + o_cy);
        s->cv = reinterpret_cast<float*>(base + o_cv); s->ch = reinterpret_cast<float*>(base + o_ch);
        s->cintent = reinterpret_cast<int*>(base + o_int); s->ncars = reinterpret_cast<int*>(base + o_nc);
        s->skip = base + o_skip; s->ealive = base + o_alive;
    }
    return o;
}

enum { LIDAR_MARCH = 0, LIDAR_FROM_HITS = 1 };

// Road march for the 32 rays of a warp (must be called by all 32 lanes, converged).  Every lane first takes two
// accelerated steps of its own ray in lock-step (that finishes ~90% of all rays: mean 1.65 steps/ray); the rays
// still open are then finished ONE AT A TIME by the whole warp, lane j testing sample k+1+j with the exact pixel
// arithmetic — so a warp never idles 31 lanes while one grazing ray crawls along a wall.
__device__ __forceinline__ int warp_road_event(bool active, const RoadView& rv, const Ray& r, bool* hit, int lane) {
    March m;
    m.k = 0; m.px = 0; m.py = 0; m.ke = LIDAR_MAX_K + 1; m.done = true; m.hit = false;
    if (active) march_init(r, m);
#pragma unroll
    for (int it = 0; it < 2; ++it)
        if (!m.done) march_step(rv, r, m);
    unsigned pend = __ballot_sync(FULL, !m.done);
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1;
        Ray o;
        o.cx = __shfl_sync(FULL, r.cx, src); o.cy = __shfl_sync(FULL, r.cy, src);
        o.dx = __shfl_sync(FULL, r.dx, src); o.dy = __shfl_sync(FULL, r.dy, src);
        const int k0 = __shfl_sync(FULL, m.k, src);
        int found = LIDAR_MAX_K + 1, fe = 0;
        for (int base = k0 + 1; base <= LIDAR_MAX_K; base += 32) {
            const int kk = base + lane;
            int px, py, e = 0;
            if (kk <= LIDAR_MAX_K) e = sample_event(rv.bits, o, kk, px, py);
            const unsigned b = __ballot_sync(FULL, e != 0);
            if (b) { const int f = __ffs(b) - 1; found = base + f; fe = __shfl_sync(FULL, e, f); break; }
        }
        if (lane == src) { m.ke = found; m.hit = (fe == 2); m.done = true; }
    }
    *hit = m.hit;
    return m.ke;
}

template <int RT>   // RT = beam count known at compile time (72, 96) or 0 = run-time d.R
__global__ void __launch_bounds__(LID_THREADS, 6)
k_lidar_obs(const Dev d, int mode, int num_groups, int G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int N = d.N, R = RT ? RT : d.R;
    const int CE = N + d.M;
    const int lane = tid & 31;
    const RoadView rv{nullptr, nullptr, d.box_lo, d.box_hi};
    LidSmem s;
    lid_carve(smem_raw, G, CE, &s);
    for (int i = tid; i < ROAD_ROWS * ROAD_WORDS; i += LID_THREADS) s.bits[i] = d.road_bits[i];
    for (int i = tid; i < SKIP_DIM * SKIP_DIM; i += LID_THREADS) s.skip[i] = d.road_skip[i];
    for (int i = tid; i < R; i += LID_THREADS) s.rel[i] = d.rel_angle[i];

    for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
        const int env0 = grp * G;
        const int ng = min(G, d.E - env0);           // envs in this group
        const int na = ng * N;                        // egos in this group
        __syncthreads();
        // ---- stage every car of the group's envs: egos then NPCs, with its lidar pixel rectangle
        for (int t = tid; t < ng * CE; t += LID_THREADS) {
            const int g = t / CE, k = t - g * CE;
            const int env = env0 + g;
            const int nn = d.traffic ? d.ncount[env] : 0;
            if (k == 0) s.ncars[g] = N + nn;
            float x, y, v, h; int intent;
            if (k < N) {
                const int ai = env * N + k;
                x = d.ex[ai]; y = d.ey[ai]; v = d.ev[ai]; h = d.eh[ai]; intent = d.route_meta[k].intent;
                s.ealive[g * N + k] = d.ealive[ai];
            } else if (k < N + nn) {
                const int ni = env * d.M + (k - N);
                x = d.nx[ni]; y = d.ny[ni]; v = d.nv[ni]; h = d.nh[ni]; intent = d.route_meta[N + d.nroute[ni]].intent;
            } else continue;
            s.cx[t] = x; s.cy[t] = y; s.cv[t] = v; s.ch[t] = h; s.cintent[t] = intent;
            s.rect[t] = car_pixel_rect(x, y, h);
        }
        __syncthreads();
        // ---- per ego: candidate set for the beams + the 31 ego/neighbour features (IntersectionEnv.cpp:431-508)
        if (tid < na) {
            const int a = tid, g = a / N, self = a - g * N;
            const int nc = s.ncars[g];
            const int cb = g * CE;                 // first car slot of this env
            const Pose me{s.cx[cb + self], s.cy[cb + self], s.cv[cb + self], s.ch[cb + self]};
            float* orow = d.obs + ((size_t)env0 * N + a) * ISX_OBS_DIM;      // the 31 feature floats go straight to the row
            int nb = 0;
            int ncand = 0;
            s.acb[a] = cb; s.aself[a] = self;
            if (s.ealive[a]) {
                const int ipx = f2i_rz(me.x), ipy = f2i_rz(me.y);
                for (int k = 0; k < nc; ++k) {
                    // Lidar.cpp:57-63: the ego itself, and anything within 1e-3 of its pose, is transparent
                    if (fabsf(s.cx[cb + k] - me.x) < 1e-3f && fabsf(s.cy[cb + k] - me.y) < 1e-3f && fabsf(s.ch[cb + k] - me.h) < 1e-3f) continue;
                    const PixRect r = s.rect[cb + k];
                    // beams reach at most 248 px (+1 px truncation) from the origin pixel
                    if (r.x0 > ipx + 250 || r.x1 < ipx - 250 || r.y0 > ipy + 250 || r.y1 < ipy - 250) continue;
                    const BeamWindow w = beam_window(r, me.x, me.y, me.h, R);
                    const int q = a * CE + ncand++;
                    s.ccar[q] = (uint8_t)k; s.cia[q] = (uint8_t)w.ia; s.cspan[q] = (uint8_t)w.span;
                }
                const F2* path = d.route_path + (size_t)self * PATH_LEN;
                const int pidx = d.epidx[(env0 + g) * N + self];
                float f6[6];
                obs_ego_features(me, path[min(pidx + 10, PATH_LEN - 1)], f6);
#pragma unroll
                for (int i = 0; i < 6; ++i) orow[i] = f6[i];
                // five nearest other alive cars, ascending distance, ties by list order (stable, :490)
                float bd[5]; int bk[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) { bd[q] = INFINITY; bk[q] = 0; }
                for (int k = 0; k < nc; ++k) {
                    if (k == self) continue;
                    if (k < N && !s.ealive[g * N + k]) continue;
                    const float dx = s.cx[cb + k] - me.x, dy = s.cy[cb + k] - me.y;
                    const float dist = fsqrt_rn(dx * dx + dy * dy);
                    int pos = 0;                       // stable insertion slot = number of kept entries <= dist
#pragma unroll
                    for (int q = 0; q < 5; ++q) pos += (bd[q] <= dist) ? 1 : 0;
#pragma unroll
                    for (int q = 4; q >= 1; --q) if (q > pos) { bd[q] = bd[q - 1]; bk[q] = bk[q - 1]; }
#pragma unroll
                    for (int q = 0; q < 5; ++q) if (q == pos) { bd[q] = dist; bk[q] = k; }
                    if (nb < 5) ++nb;
                }
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    if (q >= nb) break;
                    const int k = bk[q];
                    const Pose ot{s.cx[cb + k], s.cy[cb + k], s.cv[cb + k], s.ch[cb + k]};
                    float f5[5];
                    obs_neighbor_features(me, ot, s.cintent[cb + k], f5);
#pragma unroll
                    for (int i = 0; i < 5; ++i) orow[6 + 5 * q + i] = f5[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 6; ++i) orow[i] = 0.0f;
            }
            for (int i = 6 + 5 * nb; i < 31; ++i) orow[i] = 0.0f;          // unused neighbour slots stay zero (:424)
            s.ncand[a] = ncand;
        }
        __syncthreads();
        // ---- beams: one thread per (ego, beam); the trip count is uniform so that warps stay converged for the
        //      cooperative part of the road march
        const RoadView road{s.bits, s.skip, rv.box_lo, rv.box_hi};
        const int items = na * R;
        for (int t0 = 0; t0 < items; t0 += LID_THREADS) {
            const int t = t0 + tid;
            const bool valid = t < items;
            const int tc = valid ? t : 0;
            const int a = tc / R, i = tc - a * R;
            const int cb = s.acb[a], self = s.aself[a];
            const size_t ga = (size_t)env0 * N + a;
            const bool alive = valid && s.ealive[a];
            float out = 0.0f;                                   // dead ego: all-zero row (:426-429)
            if (mode == LIDAR_FROM_HITS) {
                if (alive) {
                    const int k = d.lidar_hit[ga * ISX_MAX_RAYS + i];
                    out = (k ? (float)(4 * k) : LIDAR_MAX_DIST) * (1.0f / LIDAR_MAX_DIST);
                }
            } else {
                Ray ray = make_ray(0.0f, 0.0f, 1.0f, 0.0f);
                if (alive) {
                    float sn, cs;
                    sincosf_(s.ch[cb + self] + s.rel[i], &sn, &cs);
                    ray = make_ray(s.cx[cb + self], s.cy[cb + self], cs, -sn);
                }
                bool hit;
                const int ke = warp_road_event(alive, road, ray, &hit, lane);
                if (alive) {
                    int best = hit ? ke : 0;
                    int lim = ke - 1;                           // cars only count strictly before the road event
                    const int nc = s.ncand[a];
                    const int iw = (i == R - 1) ? 0 : i;         // beam R-1 duplicates beam 0
                    for (int j = 0; j < nc && lim >= 1; ++j) {
                        const int q = a * CE + j;
                        const int span = s.cspan[q];
                        if (span < 255) {                        // angular window of this car (beam_window)
                            int dlt = iw - (int)s.cia[q];
                            if (dlt < 0) dlt += R - 1;
                            if (dlt > span) continue;
                        }
                        const int kh = ray_rect_first_hit(s.rect[cb + s.ccar[q]], ray, lim);
                        if (kh) { best = kh; lim = kh - 1; }
                    }
                    d.lidar_hit[ga * ISX_MAX_RAYS + i] = (uint8_t)best;
                    out = (best ? (float)(4 * best) : LIDAR_MAX_DIST) * (1.0f / LIDAR_MAX_DIST);
                }
            }
            if (valid) d.obs[ga * ISX_OBS_DIM + 31 + i] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------ small kernels
// reset() + add_car_with_route (IntersectionEnv.cpp:66-131) for masked envs
__global__ void k_reset(const Dev d, const uint8_t* __restrict__ mask) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.E * d.N) return;
    const int env = t / d.N, a = t - env * d.N;
    if (mask && !mask[env]) return;
    const RouteMeta m = d.route_meta[a];
    d.ex[t] = m.spawn_x; d.ey[t] = m.spawn_y; d.ev[t] = 0.0f; d.eh[t] = m.spawn_h;
    d.esteer[t] = 0.0f; d.eacc[t] = 0.0f; d.epd[t] = 0.0f; d.epa0[t] = 0.0f; d.epa1[t] = 0.0f;
    d.epidx[t] = 0; d.ealive[t] = 1;
    d.reward[t] = 0.0f; d.done[t] = 0; d.status[t] = ISX_ALIVE;
    for (int i = 0; i < ISX_MAX_RAYS; ++i) d.lidar_hit[(size_t)t * ISX_MAX_RAYS + i] = 0;
    if (a == 0) {
        d.ncount[env] = 0; d.next_uid[env] = 1; d.step_count[env] = 0;
        d.terminated[env] = 0; d.truncated[env] = 0; d.agents_alive[env] = d.N;
        isx_traffic_events e; e.rng_draws = 0; e.spawn_route = -1; e.spawned = 0; e.removed_mask = 0; e.collided_mask = 0; e.npc_count = 0;
        d.events[env] = e;
    }
}

__global__ void k_reduce_stats(const Dev d) {
    // one CTA; 64-bit sums of the per-env u32 counters, reward_sum in double
    __shared__ unsigned long long acc[STAT_SLOTS];
    __shared__ double racc;
    if (threadIdx.x < STAT_SLOTS) acc[threadIdx.x] = 0;
    if (threadIdx.x == 0) racc = 0.0;
    __syncthreads();
    unsigned long long loc[14];
    double r = 0.0;
    for (int i = 0; i < 14; ++i) loc[i] = 0;
    for (int e = threadIdx.x; e < d.E; e += blockDim.x) {
        const uint32_t* st = d.env_stats + (size_t)e * STAT_SLOTS;
        for (int i = 0; i < 14; ++i) loc[i] += st[i];
        r += *reinterpret_cast<const double*>(st + ST_RSUM);
    }
    for (int i = 0; i < 14; ++i) atomicAdd(&acc[i], loc[i]);
    atomicAdd(&racc, r);
    __syncthreads();
    if (threadIdx.x < 14) d.stats[threadIdx.x] = acc[threadIdx.x];
    if (threadIdx.x == 14) d.stats[14] = 0;
    if (threadIdx.x == 15) d.stats[15] = (unsigned long long)__double_as_longlong(racc);
}

// contraction canary: (a*b + c) with operands chosen so that a fused multiply-add gives a different float
__global__ void k_canary(float a, float b, float c, float* out) { out[0] = a * b + c; double x = a, y = b, z = c; out[1] = (float)(x * y + z); }

__global__ void k_math_probe(int n, const float* a, const float* b, float* sn, float* cs, float* tn, float* at, float* hy, float* wr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sincosf_(a[i], &sn[i], &cs[i]);
    tn[i] = tanf_(a[i]);
    at[i] = atan2f_(a[i], b[i]);
    hy[i] = hypotf_(a[i], b[i]);
    wr[i] = wrap_angle(a[i]);
}

// ------------------------------------------------------------------------------------------------ launchers
size_t lidar_smem_bytes(const Dev& d) { const int G = max(1, LID_AGENTS / d.N); return lid_carve(nullptr, G, d.N + d.M, nullptr); }

cudaError_t launch_dynamics(const Dev& d, const float* actions, float dt, float spawn_prob, cudaStream_t st) {
    const int blocks = (d.E + DYN_WARPS - 1) / DYN_WARPS;
    k_dynamics<<<blocks, DYN_WARPS * 32, 0, st>>>(d, actions, dt, spawn_prob);
    return cudaGetLastError();
}
cudaError_t launch_lidar_obs(const Dev& d, int mode, int grid_cap, cudaStream_t st) {
    const int G = max(1, LID_AGENTS / d.N);
    const int groups = (d.E + G - 1) / G;
    const int grid = min(groups, grid_cap);
    const size_t sm = lidar_smem_bytes(d);
    if (d.R == 72) k_lidar_obs<72><<<grid, LID_THREADS, sm, st>>>(d, mode, groups, G);
    else if (d.R == 96) k_lidar_obs<96><<<grid, LID_THREADS, sm, st>>>(d, mode, groups, G);
    else k_lidar_obs<0><<<grid, LID_THREADS, sm, st>>>(d, mode, groups, G);
    return cudaGetLastError();
}
cudaError_t launch_reset(const Dev& d, const uint8_t* mask, cudaStream_t st) {
    const int n = d.E * d.N;
    k_reset<<<(n + 255) / 256, 256, 0, st>>>(d, mask);
    return cudaGetLastError();
}
cudaError_t launch_reduce_stats(const Dev& d, cudaStream_t st) {
    k_reduce_stats<<<1, 1024, 0, st>>>(d);
    return cudaGetLastError();
}
cudaError_t launch_canary(float a, float b, float c, float* out, cudaStream_t st) {
    k_canary<<<1, 1, 0, st>>>(a, b, c, out);
    return cudaGetLastError();
}
cudaError_t launch_math_probe(int n, const float* a, const float* b, float* sn, float* cs, float* tn, float* at, float* hy, float* wr, cudaStream_t st) {
    k_math_probe<<<(n + 255) / 256, 256, 0, st>>>(n, a, b, sn, cs, tn, at, hy, wr);
    return cudaGetLastError();
}
cudaError_t lidar_set_smem_attr(const Dev& d) {
    const int b = (int)lidar_smem_bytes(d);
    cudaError_t e = cudaFuncSetAttribute(k_lidar_obs<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lidar_obs<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lidar_obs<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    return e;
}
cudaError_t lidar_occupancy(const Dev& d, int* ctas_per_sm) {
    if (d.R == 72) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<72>, LID_THREADS, lidar_smem_bytes(d));
    if (d.R == 96) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<96>, LID_THREADS, lidar_smem_bytes(d));
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<0>, LID_THREADS, lidar_smem_bytes(d));
}

}  // namespace isx
