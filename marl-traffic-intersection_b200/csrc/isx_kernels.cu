// isx_kernels.cu — the step kernels (sm_100a).
//
//   k_traffic   : 8 lanes per env, four envs per warp.  NPC traffic flow (spawn draw, Gauss-Seidel planner/integrator in list order,
//                 NPC-NPC SAT, ordered erase), lanes spread over ghost-path points / other NPCs / path-window
//                 points.  Replaces TrafficFlow.cpp:317-367.  Only launched with traffic flow on.  Envs that share a warp are
//                 taken from lists filed by NPC count (k_traffic_order, one thread per env, just before).
//   k_ego       : one LANE per ego, an env = a sub-warp of 2^ceil(log2 N) lanes: physics, reward, status, car-car
//                 override, bonuses, team mix, respawn, termination.  Replaces IntersectionEnv.cpp:137-370.
//   k_features  : FOUR lanes per ego: the 31 ego/neighbour observation features + the per-ego list of cars its
//                 beams can touch (with angular beam windows).  Replaces IntersectionEnv.cpp:418-508.
//   k_lidar_obs : one THREAD per (ego, beam), persistent CTAs with the folded road bitmap in shared
//                 memory, warps claim 32-beam pieces dynamically: analytic road bound + exact sample tests, slab/verify
//                 against the other cars' pixel rectangles.  Writes the lidar part of the obs rows coalesced.
//                 Replaces Lidar.cpp:16-90 and IntersectionEnv.cpp:374-390, 510-514.
//
// No tensor cores: nothing here is a dense contraction.  Build with -fmad=false (see isx_math.cuh).
#include <cuda_runtime.h>

#include "isx_device.cuh"

namespace isx {

constexpr unsigned FULL = 0xffffffffu;
#ifndef ISX_DYN_WARPS
#define ISX_DYN_WARPS 1     // k_traffic: one env (warp) per CTA — a CTA's slot is freed as soon as ITS env is done, instead of
#endif                      // waiting for the slowest of four (measured 39.1 -> 36.5 us at 8192 envs; 2: 38.6, 8: 40.5)
#ifndef ISX_EGO_THREADS
#define ISX_EGO_THREADS 128
#endif
#ifndef ISX_FEAT_THREADS
#define ISX_FEAT_THREADS 128
#endif
#ifndef ISX_TRAFFIC_MINB
#define ISX_TRAFFIC_MINB 48   // 32-thread CTAs; a value above the 32-CTA/SM hardware limit leaves the register count to ptxas: 72 registers,
                              // 28 CTAs/SM, no spills (103 us at 65536 envs; capped at 64 registers = MINB 32: spills, 107 us)
#endif
#ifndef ISX_EGO_MINB
#define ISX_EGO_MINB 8        // 64 registers, no spills (76.3 -> 75.1 us at 65536 envs; tighter caps spill and lose)
#endif
#ifndef ISX_FEAT_MINB
#define ISX_FEAT_MINB 12    // k_features is latency-bound: 40 registers (11 words of spill) let 12 CTAs instead of 5 share an SM
#endif                      // (8192x8: 30.9 -> 25 us; 65536x8: 151 -> 139 us; 16 CTAs / 32 registers spill too much: 148 us)
constexpr int DYN_WARPS = ISX_DYN_WARPS;   // envs (warps) per CTA in k_traffic
constexpr int EGO_THREADS = ISX_EGO_THREADS;
constexpr int FEAT_THREADS = ISX_FEAT_THREADS;
#ifndef ISX_TEST_BRANCHFREE
#define ISX_TEST_BRANCHFREE 1
#endif
#ifndef ISX_LID_MINB
#define ISX_LID_MINB 4      // CTAs per SM: the kernel is issue-bound, not latency-bound, so registers beat occupancy (round-2 history: 8 -> 844 us, 5 -> 788; after the instruction trims 5 -> 716, 4 -> 710, 3 -> 712, 6 -> 770)
#endif
constexpr int LID_THREADS = 256;
#ifndef ISX_TRAFFIC_LANES
#define ISX_TRAFFIC_LANES 0     // 0: chosen per launch by batch size (launch_traffic)
#endif
constexpr int TRAFFIC_LANES = ISX_TRAFFIC_LANES;   // lanes per env in k_traffic: 8 (four envs per warp), 16, 32, or 0 = by batch size
#ifndef ISX_GHOST_PTS
#define ISX_GHOST_PTS 2     // k_traffic ghost-path scan: blocks of L points a group tests per pass
#endif
constexpr int GHOST_PTS = ISX_GHOST_PTS;
#ifndef ISX_WARP_GRAB
#define ISX_WARP_GRAB 3
#endif
#ifndef ISX_LOCKSTEP
#define ISX_LOCKSTEP 2
#endif
#ifndef ISX_LOCKSTEP_EXTRA
#define ISX_LOCKSTEP_EXTRA 2
#endif
#ifndef ISX_LOCKSTEP_MIN_OPEN
#define ISX_LOCKSTEP_MIN_OPEN 2
#endif
constexpr int WARP_GRAB = ISX_WARP_GRAB;   // 32-beam pieces a warp claims per atomic in k_lidar_obs
constexpr int LOCKSTEP = ISX_LOCKSTEP;     // exact sample tests every lane takes after the analytic jump ...
constexpr int LOCKSTEP_EXTRA = ISX_LOCKSTEP_EXTRA;        // ... up to this many more while at least
constexpr int LOCKSTEP_MIN_OPEN = ISX_LOCKSTEP_MIN_OPEN;  // this many rays of the warp are still open; then the cooperative tail
constexpr int ROAD_BITS_BYTES = ((ROAD_ROWS * ROAD_WORDS * 4 + 15) / 16) * 16;
constexpr int ROAD_SKIP_BYTES = ((SKIP_DIM * SKIP_DIM + 15) / 16) * 16;

#define ISX_STAMP(slot) do { if (d.trace && env_ok && lane == 0) d.trace[(size_t)env * 16 + (slot)] = clock64(); } while (0)

struct NpcSmem {
    float x[ISX_MAX_NPC], y[ISX_MAX_NPC], v[ISX_MAX_NPC], h[ISX_MAX_NPC], steer[ISX_MAX_NPC];
    int pidx[ISX_MAX_NPC], route[ISX_MAX_NPC];
    uint32_t uid[ISX_MAX_NPC];
    uint32_t coll[ISX_MAX_NPC];
};

// ------------------------------------------------------------------------------------------------ launch chaining
// Programmatic dependent launch (sm_90+): the four step kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the CTAs of kernel k+1 may become resident while kernel k drains
// its last wave.  Every kernel first lets ITS successor start (launch_dependents), runs whatever does not depend on the
// predecessor (k_lidar_obs: staging 18 KB of constant tables per CTA), and only then waits for the predecessor grid to
// complete and flush (wait) — nothing a predecessor writes is touched before that.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ k_traffic
// NPC traffic flow (TrafficFlow.cpp:317-367), one group of L lanes per env: spawn draw, Gauss-Seidel planner/integrator
// in list order, NPC-NPC SAT, ordered erase.  The lanes spread over path-window points / other NPCs / ghost-path points.
// Reads the egos' PRE-step positions (spawn blocking, :244-249), so it runs before k_ego.  Only launched when traffic
// flow is enabled.
// L = lanes per env.  L = 32: one env per warp, any NPC count.  L = 8 (the default) / 16: FOUR / two envs share a warp,
// one per group of L lanes — the mean env has ~1 NPC, so a whole warp per env runs its list-order chain on one or two
// lanes; packing more envs into the same instruction stream divides the warps the latency-bound kernel has to retire.
// Every warp-wide primitive below is group-wide: shuffles of width L, ballots masked to the group's lanes, loop bounds
// made warp-uniform (max over the groups) with the body predicated per group.
template <int L>
struct Grp {
    static constexpr unsigned MASK = L == 32 ? FULL : ((1u << (L & 31)) - 1u);
    int shift;                                                        // first lane of my group within the warp
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(FULL, p) >> shift) & MASK; }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(FULL, v, src, L); }
    __device__ __forceinline__ float min_f(float v) const {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o, L));
        return v;
    }
    __device__ __forceinline__ int umax(int v) const {                                                                // warp-uniform bound
        if (L <= 16) v = max(v, __shfl_xor_sync(FULL, v, 16));
        if (L <= 8) v = max(v, __shfl_xor_sync(FULL, v, 8));
        return v;
    }
};

// Car::update_path_index with the 50-point window spread over the L lanes of a group; first minimum wins (Car.cpp:62-70).
template <int L>
__device__ __forceinline__ int group_path_index(const F2* __restrict__ path, int idx, float x, float y, int lane) {
    const int start = idx < 0 ? 0 : idx;
    const int end = min(start + 50, PATH_LEN);
    float best = INFINITY;
    int bi = start;
#pragma unroll
    for (int r = 0; r < (50 + L - 1) / L; ++r) {
        const int i = start + lane + r * L;
        if (i < end) {
            const F2 p = path[i];
            const float dx = p.x - x, dy = p.y - y;
            const float d = dx * dx + dy * dy;
            if (d < best) { best = d; bi = i; }
        }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, o, L);
        const int oi = __shfl_xor_sync(FULL, bi, o, L);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    return bi;
}

// One env's traffic update on a group of L lanes (wl = lane within the warp; all 32 lanes of the warp call this together,
// each group with its own env).  env_raw >= d.E: an idle group (c = 0, no writes).
template <int L>
__device__ __forceinline__ void traffic_env(const Dev& d, float dt, float spawn_prob, int env_raw, int wl, NpcSmem& sm) {
    const int sub = wl / L, lane = wl % L;                            // `lane`: my lane within the env's group
    const Grp<L> g{sub * L};
    const bool env_ok = env_raw < d.E;                                // idle groups: c = 0, no writes
    const int env = env_ok ? env_raw : d.E - 1;
    const int N = d.N;
    const uint32_t genv = (uint32_t)(d.env_base + env);
    int c = env_ok ? d.ncount[env] : 0;
    uint32_t next_uid = d.next_uid[env];
    const bool reset_now = d.auto_reset && (d.terminated[env] | d.truncated[env]);   // env.py:147-152 after a done step
    if (reset_now) { c = 0; next_uid = 1; }
    // auto_reset == 2 (next-step reset, see k_ego): the call that resets an env does not simulate it — no spawn draw, no tick
    const bool frozen = reset_now && d.auto_reset == 2;
    const uint32_t tick = d.tick[env] + 1;
    ISX_STAMP(0);

    isx_traffic_events evt;
    evt.rng_draws = 0; evt.spawn_route = -1; evt.spawned = 0; evt.removed_mask = 0; evt.collided_mask = 0; evt.npc_count = 0;
    uint32_t overflow = 0;

    if (lane < c) {
        const int ni = env * d.M + lane;
        sm.x[lane] = d.nx[ni]; sm.y[lane] = d.ny[ni]; sm.v[lane] = d.nv[ni]; sm.h[lane] = d.nh[ni];
        sm.steer[lane] = d.nsteer[ni]; sm.pidx[lane] = d.npidx[ni]; sm.route[lane] = d.nroute[ni]; sm.uid[lane] = d.nuid[ni];
    }
    __syncwarp();
    // -- spawn draw (:321-329) and try_spawn_traffic_car (:275-315); every lane of the group runs the same stream
    TrafficStream ts;
    ts.init(d.seed, genv, tick);
    {
        const bool want = env_ok && !frozen && ts.uniform01() < spawn_prob && d.T > 0;
        int r = 0;
        bool blk = false;
        RouteMeta m{};
        if (want) {
            r = (int)ts.below((uint32_t)d.T);
            evt.spawn_route = r;
            m = d.route_meta[N + r];
            const float md = CAR_LENGTH * 2.5f, md2 = md * md;   // is_spawn_blocked (:240-259)
            for (int a = lane; a < N; a += L) {
                float ex, ey;
                if (reset_now) { const RouteMeta em = d.route_meta[a]; ex = em.spawn_x; ey = em.spawn_y; }
                else { ex = d.ex[env * N + a]; ey = d.ey[env * N + a]; }
                const float dx = ex - m.spawn_x, dy = ey - m.spawn_y;
                blk = blk || (dx * dx + dy * dy < md2);
            }
            if (lane < c) { const float dx = sm.x[lane] - m.spawn_x, dy = sm.y[lane] - m.spawn_y; blk = blk || (dx * dx + dy * dy < md2); }
        }
        const bool blocked = g.any(blk);
        if (want && !blocked) {
            if (c < d.M) {
                if (lane == 0) {
                    sm.x[c] = m.spawn_x; sm.y[c] = m.spawn_y; sm.v[c] = 0.0f; sm.h[c] = m.spawn_h; sm.steer[c] = 0.0f;
                    sm.pidx[c] = 0; sm.route[c] = r; sm.uid[c] = next_uid;
                }
                next_uid += 1; c += 1; evt.spawned = 1;
            } else overflow = 1;                                   // reference list is unbounded; counted
        }
        __syncwarp();
    }
    evt.rng_draws = frozen ? 0 : (int)ts.j;
    ISX_STAMP(1);

    // -- NPC controller (:337-344).  The reference updates NPCs one after the other, NPC i seeing the already-updated
    //    NPCs < i.  Everything that depends only on an NPC's OWN pre-update state — first path-index update, steering
    //    command, steering low-pass and its tangent, heading sin/cos, distance to the centre — is evaluated for all NPCs
    //    at once (lane = NPC); only the parts that look at the other NPCs run in list order.  That cuts the serial
    //    chain per NPC to about a quarter (the kernel time is the critical path of the env with the most NPCs).
    Pose cur{0, 0, 0, 0};                       // lane i < c: current state of NPC i
    float my_steer = 0.0f, my_tan = 0.0f, my_sin = 0.0f, my_cos = 1.0f, my_dc = 0.0f;
    int my_pidx = 0, my_route = 0;
    if (lane < c) {
        cur = Pose{sm.x[lane], sm.y[lane], sm.v[lane], sm.h[lane]};
        my_route = sm.route[lane];
        my_pidx = sm.pidx[lane];
    }
    // first path-index update of every NPC (Car.cpp:47-74), the 50-point window spread over the lanes of the group: a lane
    // scanning its own NPC's window alone kept 2.4 of 32 lanes busy for a quarter of the kernel's instructions
    const int cmax = g.umax(c);
    for (int i = 0; i < cmax; ++i) {
        const F2* path_i = d.route_path + (size_t)(N + g.shfl(my_route, i)) * PATH_LEN;
        const int pi = group_path_index<L>(path_i, g.shfl(my_pidx, i), g.shfl(cur.x, i), g.shfl(cur.y, i), lane);
        if (lane == i) my_pidx = pi;
    }
    if (lane < c) {
        const F2* path = d.route_path + (size_t)(N + my_route) * PATH_LEN;
        const float steer_cmd = npc_steer_cmd(cur, path[min(my_pidx + 12, PATH_LEN - 1)]);
        my_steer = car_steer_update(sm.steer[lane], steer_cmd);
        // inlined on purpose: heading sin/cos and the centre distance do not depend on the steering chain
        // (path point -> atan2 -> low-pass -> tan), so the scheduler can interleave the three
        sincosf_(cur.h, &my_sin, &my_cos);
        my_dc = hypotf_(cur.x - WIDTH * 0.5f, cur.y - HEIGHT * 0.5f);
        my_tan = tanf_nc(my_steer);
    }
    ISX_STAMP(2);
    for (int i = 0; i < cmax; ++i) {
        const bool on = i < c;                                     // my group still has an NPC i
        Pose me;
        me.x = g.shfl(cur.x, i); me.y = g.shfl(cur.y, i);
        me.v = g.shfl(cur.v, i); me.h = g.shfl(cur.h, i);
        const float ms = g.shfl(my_sin, i), mc = g.shfl(my_cos, i);
        const float me_dc = g.shfl(my_dc, i), tan_s = g.shfl(my_tan, i);
        const int mp0 = g.shfl(my_pidx, i);
        const F2* path = d.route_path + (size_t)(N + g.shfl(my_route, i)) * PATH_LEN;
        float fc = 1e9f;
        int flags = 0;
        if (on && lane < c && lane != i) {
            flags = npc_pair_eval(me, cur, ms, mc, me_dc, i < lane, &fc);
        }
        if (i == 0) ISX_STAMP(8);
        const float thr0 = npc_cruise_throttle(me.v, g.min_f(fc));
        const unsigned elig = g.ballot(flags & 1);
        const unsigned yld = g.ballot(flags & 2);
        bool conflict = false;
        float min_conf = 1e9f;
        if (__any_sync(FULL, elig != 0u)) {                        // ghost-path scan (:91-185), L points per pass
            const float safe_sq = (CAR_WIDTH * 2.0f) * (CAR_WIDTH * 2.0f);
            const int s1 = min(mp0 + 120, PATH_LEN);
            // GHOST_PTS blocks of L points per pass: the pass is a dependent chain (load, test, ballot, branch), and most scans
            // run all 120 points without a conflict — fewer, wider passes; the first conflict in index order still wins
            for (int base = mp0;; base += GHOST_PTS * L) {
                const bool go = elig != 0u && base < s1 && !conflict;
                if (!__any_sync(FULL, go)) break;
                bool hit[GHOST_PTS];
                float dtc[GHOST_PTS];
#pragma unroll
                for (int jb = 0; jb < GHOST_PTS; ++jb) {
                    const int gp_i = base + jb * L + lane;
                    hit[jb] = false; dtc[jb] = 0.0f;
                    if (go && gp_i < s1) {
                        const F2 gp = path[gp_i];
                        unsigned near_yield = 0, near_any = 0;
                        for (unsigned mm = elig; mm; mm &= mm - 1) {
                            const int o = __ffs(mm) - 1;
                            const float dx = sm.x[o] - gp.x, dy = sm.y[o] - gp.y;
                            if (dx * dx + dy * dy < safe_sq) { near_any = 1; near_yield |= (yld >> o) & 1u; }
                        }
                        if (near_any) {
                            dtc[jb] = hypotf_nc(gp.x - me.x, gp.y - me.y);
                            hit[jb] = near_yield || (dtc[jb] < 15.0f);
                        }
                    }
                }
#pragma unroll
                for (int jb = 0; jb < GHOST_PTS; ++jb) {
                    const unsigned hb = g.ballot(hit[jb]);
                    const float first_dtc = g.shfl(dtc[jb], hb ? __ffs(hb) - 1 : 0);
                    if (hb && !conflict) { conflict = true; min_conf = first_dtc; }
                }
            }
        }
        if (i == 0) ISX_STAMP(9);
        const float thr = npc_final_throttle(thr0, conflict, min_conf);
        float macc = 0.0f;
        car_motion_update(me, macc, thr, tan_s, dt);               // every lane of the group, same operands: uniform
        if (i == 0) ISX_STAMP(10);
        const int mp = group_path_index<L>(path, mp0, me.x, me.y, lane);
        if (i == 0) ISX_STAMP(11);
        if (on && lane == i) { cur = me; my_pidx = mp; sm.x[i] = me.x; sm.y[i] = me.y; }
        __syncwarp();
    }
    if (lane < c) { sm.v[lane] = cur.v; sm.h[lane] = cur.h; sm.steer[lane] = my_steer; sm.pidx[lane] = my_pidx; }
    __syncwarp();

    ISX_STAMP(3);
    // -- NPC-NPC collisions (:347-356): lane j tests the pair (i, j), j > i
    unsigned alive_m = c >= 32 ? FULL : ((1u << c) - 1u);
    const unsigned all_m = alive_m;
    for (int i = 0; i + 1 < cmax; ++i) {
        bool hit = false;
        if (lane > i && lane < c) hit = cars_collide(sm.x[i], sm.y[i], sm.h[i], sm.x[lane], sm.y[lane], sm.h[lane]);
        const unsigned hm = g.ballot(hit);
        if (i + 1 < c && ((alive_m >> i) & 1u)) {
            const unsigned m = hm & alive_m;
            if (m) alive_m &= ~(m | (1u << i));
        }
    }
    ISX_STAMP(4);
    // -- ordered erase of dead / arrived / out-of-screen (:359-366)
    bool rem = false;
    float mx = 0, my = 0, mv = 0, mh = 0, mst = 0; int mpi = 0, mr = 0; uint32_t mu = 0;
    if (lane < c) {
        mx = sm.x[lane]; my = sm.y[lane]; mv = sm.v[lane]; mh = sm.h[lane]; mst = sm.steer[lane];
        mpi = sm.pidx[lane]; mr = sm.route[lane]; mu = sm.uid[lane];
        const F2 goal = d.route_meta[N + mr].goal;
        const bool arrived = hypotf_nc(mx - goal.x, my - goal.y) < 20.0f;
        const bool oos = mx < -100.0f || mx > (float)WIDTH + 100.0f || my < -100.0f || my > (float)HEIGHT + 100.0f;
        rem = !((alive_m >> lane) & 1u) || arrived || oos;
    }
    const unsigned rem_m = g.ballot(rem);
    evt.removed_mask = rem_m;
    evt.collided_mask = all_m & ~alive_m;
    const unsigned keep_m = all_m & ~rem_m;
    if (lane < c && !rem) {
        const int pos = __popc(keep_m & ((1u << lane) - 1u));
        const int ni = env * d.M + pos;
        d.nx[ni] = mx; d.ny[ni] = my; d.nv[ni] = mv; d.nh[ni] = mh; d.nsteer[ni] = mst;
        d.npidx[ni] = mpi; d.nroute[ni] = mr; d.nuid[ni] = mu;
        // lidar pixel rectangle of this NPC for k_features / k_lidar_obs
        reinterpret_cast<PixRect*>(d.car_rect)[(size_t)env * (N + d.M) + N + pos] = car_pixel_rect(mx, my, mh);
    }
    c = __popc(keep_m);
    evt.npc_count = c;
    if (env_ok && lane == 0) { d.ncount[env] = c; d.next_uid[env] = next_uid; d.events[env] = evt; }
    ISX_STAMP(5);
    if (d.trace && env_ok && lane == 0) d.trace[(size_t)env * 16 + 6] = c;
    if (env_ok && lane == 0) {
        uint32_t* st = d.env_stats + (size_t)env * STAT_SLOTS;
        if (evt.spawned) st[ST_SPAWNED] += 1u;
        if (evt.removed_mask) st[ST_REMOVED] += (uint32_t)__popc(evt.removed_mask);
        if (evt.collided_mask) st[ST_COLLIDED] += (uint32_t)__popc(evt.collided_mask);
        if (overflow) st[ST_OVERFLOW] += overflow;
    }
}

// L lanes per env.  L = 8 (the default): FOUR envs share a warp — the mean env holds about one NPC, so the list-order chain
// of an env runs on one or two lanes whatever the group width, and packing more envs into one instruction stream divides the
// warps this latency-bound kernel has to retire (2 envs per warp: 102 -> 77 us at 32768 envs; 4 per warp: see DESIGN.md).  A group
// of 8 lanes holds at most 8 NPCs: when an env of the warp already has 8 and room for a ninth (capacity > 8), the warp
// steps its four envs one after the other on all 32 lanes instead (rare: the largest population observed is 13).
// Envs that share a warp in k_traffic<8 / 16> run in lock-step for as many list-order rounds as the FULLEST of them holds
// NPCs: in the steady state of BASELINE's C5 an env holds 0 / 1 / 2 / 3 / 4+ NPCs with probability .27 / .44 / .23 / .06 /
// .002, so four envs in index order make a warp run 1.97 rounds on average where the mean env needs 1.07.  This kernel
// (thread per env, just before k_traffic) files every env of the view into one of five lists by the NPC count it starts the
// step with; k_traffic then takes its envs from the lists — fullest first (longest chains start first), equal counts side by
// side.  Envs are independent, so the order in which they are stepped (and the arbitrary order inside a list: one
// CTA-aggregated atomic per list) changes no result, only which envs share an instruction stream.
constexpr int ORDER_LISTS = 5;
constexpr int ORDER_THREADS = 1024;
__global__ void __launch_bounds__(ORDER_THREADS)
k_traffic_order(const Dev d) {
    pdl_launch_dependents();
    pdl_wait();                                                       // keeps the launch chain transitive (k_traffic waits for THIS grid only)
    // one atomic per list and CTA (a warp-aggregated one per warp put 2048 atomics on the same word at 65,536 envs: the
    // kernel took as long as those took to serialise)
    __shared__ unsigned s_cnt[ORDER_THREADS / 32][ORDER_LISTS];       // per warp: envs it files into list k, then its offset in the CTA's share
    __shared__ unsigned s_base[ORDER_LISTS];
    const int env = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b = -1;
    if (env < d.E) {
        const bool reset_now = d.auto_reset && (d.terminated[env] | d.truncated[env]);
        b = reset_now ? 0 : min(d.ncount[env], ORDER_LISTS - 1);
    }
    unsigned mine = 0;                                                // lanes below me that go to my list
#pragma unroll
    for (int k = 0; k < ORDER_LISTS; ++k) {
        const unsigned m = __ballot_sync(FULL, b == k);
        if (lane == 0) s_cnt[warp][k] = (unsigned)__popc(m);
        if (b == k) mine = (unsigned)__popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < ORDER_LISTS) {
        unsigned tot = 0;
        for (int w = 0; w < ORDER_THREADS / 32; ++w) { const unsigned c = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = tot; tot += c; }
        s_base[threadIdx.x] = tot ? atomicAdd(d.order_cnt + threadIdx.x, tot) : 0u;
    }
    __syncthreads();
    if (b >= 0) d.order[(size_t)b * d.order_stride + s_base[b] + s_cnt[warp][b] + mine] = env;
}
// slot -> env: the lists laid end to end, fullest first (slots past the last env stay idle: >= d.E)
__device__ __forceinline__ int ordered_env(const Dev& d, int slot) {
    if (d.order == nullptr || slot >= d.E) return slot;
    int s = slot;
#pragma unroll
    for (int k = ORDER_LISTS - 1; k >= 0; --k) {
        const int n = (int)d.order_cnt[k];
        if (s < n) return d.order[(size_t)k * d.order_stride + s];
        s -= n;
    }
    return d.E;                                                       // not reached: the lists hold every env exactly once
}

template <int L>
__global__ void __launch_bounds__(DYN_WARPS * 32, ISX_TRAFFIC_MINB)
k_traffic(const Dev d, float dt, float spawn_prob) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int EPW = 32 / L;                                       // envs per warp
    __shared__ NpcSmem sm_all[DYN_WARPS * EPW];
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    const int env0 = (blockIdx.x * DYN_WARPS + warp) * EPW;           // first env slot of this warp
    if (env0 >= d.E) return;                                          // whole warp leaves; only warp-level sync below
    if (L < 32) {
        const int e = ordered_env(d, env0 + wl / L);
        bool wide = false;
        if (e < d.E && d.M > L) {
            const bool reset_now = d.auto_reset && (d.terminated[e] | d.truncated[e]);
            wide = !reset_now && d.ncount[e] >= L;
        }
        if (__any_sync(FULL, wide)) {
            for (int k = 0; k < EPW; ++k) { traffic_env<32>(d, dt, spawn_prob, __shfl_sync(FULL, e, k * L), wl, sm_all[warp * EPW]); __syncwarp(); }
            return;
        }
        traffic_env<L>(d, dt, spawn_prob, e, wl, sm_all[warp * EPW + wl / L]);
        return;
    }
    traffic_env<L>(d, dt, spawn_prob, env0 + wl / L, wl, sm_all[warp * EPW + wl / L]);
}

// ------------------------------------------------------------------------------------------------ k_ego
// Ego physics, reward, status, car-car override, bonuses, team mix, respawn, termination (IntersectionEnv.cpp:137-370).
// One LANE per ego; an env occupies a sub-warp of NP = 2^ceil(log2 N) lanes, so a warp serves 32/NP envs and the
// lanes stay busy (8 egos/env -> 4 envs per warp).  All exchanges between the egos of an env are sub-warp shuffles /
// ballot slices.  Runs after k_traffic (ego-NPC collisions see the post-update NPCs, :307-317).
template <int NP>
__global__ void __launch_bounds__(EGO_THREADS, ISX_EGO_MINB)
k_ego(const Dev d, const float* __restrict__ actions, float dt) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int EPW = 32 / NP;                                  // envs per warp
    constexpr unsigned LOW = NP == 32 ? 0xffffffffu : ((1u << NP) - 1u);
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / NP, a = lane % NP;
    const int env_raw = wg * EPW + sub;
    if (d.order_cnt != nullptr && blockIdx.x == 0 && threadIdx.x < 8) d.order_cnt[threadIdx.x] = 0u;   // k_traffic is done with its env lists
    const bool env_ok = env_raw < d.E;
    const int env = env_ok ? env_raw : d.E - 1;
    const int N = d.N;
    const bool is_ego = env_ok && a < N;
    const int ai = env * N + (a < N ? a : 0);
    const int shift = sub * NP;
    const uint32_t genv = (uint32_t)(d.env_base + env);

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
    uint32_t resets = 0;
    // ---- auto-reset: what a caller of env.py does after terminated|truncated (reset(), env.py:147-152)
    if (d.auto_reset && (d.terminated[env] | d.truncated[env])) {
        if (is_ego) {
            const RouteMeta m = d.route_meta[a];
            p.x = m.spawn_x; p.y = m.spawn_y; p.v = 0.0f; p.h = m.spawn_h;
            steer = 0; acc = 0; pd = 0; pa0 = 0; pa1 = 0; pidx = 0; alive = true;
        }
        step_count = 0; resets = 1;
    }
    // auto_reset == 1: the reset happens at the start of the step and this call's action already drives the new episode
    //   (the caller chose it looking at the terminal observation).
    // auto_reset == 2: next-step reset as in Gymnasium's vector envs — the call after a terminated|truncated step only resets:
    //   the action is ignored, nothing is simulated (step 0, tick unchanged, no traffic), reward 0, done 0, and the
    //   observation returned is the one reset() returns (env.py:147-165: lidar part 1.0).
    const bool frozen = resets != 0 && d.auto_reset == 2;
    step_count += frozen ? 0 : 1;                  // IntersectionEnv.cpp:137
    const uint32_t tick = d.tick[env] + (frozen ? 0u : 1u);
    const int c = d.traffic ? d.ncount[env] : 0;   // NPCs after this step's traffic update

    float rew = 0.0f;
    int status = ISX_ALIVE;
    bool done = false;
    float sn = 0.0f, cs = 1.0f;                    // sine / cosine of p.h, evaluated once per ego and step
    bool have_sc = false;
    if (is_ego && !frozen) {
        if (alive) {                               // :151-163
            float thr, st;
            if (actions) { thr = actions[2 * ai]; st = actions[2 * ai + 1]; }
            else philox_action(d.seed, genv, tick, (uint32_t)a, thr, st);
            car_update_sc(p, steer, acc, thr, st, dt, &sn, &cs);       // (sn, cs) of the new heading: also the corners, the SAT, the rectangle
            have_sc = true;
            const RouteMeta m = d.route_meta[a];
            rew = reward_base(d.rc, p.x, p.y, p.v, acc, steer, m.goal, d.max_progress, pd, pa0, pa1);
            status = ego_self_status_sc(d.lanes, p.x, p.y, sn, cs, m.goal, m.goal_prev);     // :166-290
            done = status != ISX_ALIVE;
        } else { status = ISX_DEAD; done = true; }
    }
    // -- Car::update_path_index (Car.cpp:47-74; :163 of the step; nothing above reads the new index).  Every lane scans the
    //    first PATH_NEAR points of its window; the far part is only needed when the bound of path_index_update fails (a car
    //    far off its path) and is then scanned by the whole warp for that one lane — one straggler no longer makes 32 lanes
    //    walk 34 more points each.
    {
        const bool upd = is_ego && !frozen && alive;
        const int start = pidx < 0 ? 0 : pidx;
        float best = INFINITY, d0 = 0.0f;
        int bi = start;
        if (upd) path_index_near(d.route_path + (size_t)a * PATH_LEN, start, p.x, p.y, best, bi, d0);
        const bool need = upd && !(2.0f * (d0 + best) < d.route_far2[(size_t)a * PATH_LEN + min(start, PATH_LEN - 1)]);
        unsigned todo = __ballot_sync(FULL, need);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int s0 = __shfl_sync(FULL, start, src) + PATH_NEAR, s1 = min(s0 - PATH_NEAR + PATH_WINDOW, PATH_LEN);
            const float sx = __shfl_sync(FULL, p.x, src), sy = __shfl_sync(FULL, p.y, src);
            const F2* path = d.route_path + (size_t)__shfl_sync(FULL, a, src) * PATH_LEN;
            float bd = INFINITY;
            int bidx = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < (PATH_WINDOW - PATH_NEAR + 31) / 32; ++r) {
                const int i = s0 + lane + 32 * r;
                if (i < s1) {
                    const F2 q = path[i];
                    const float dx = q.x - sx, dy = q.y - sy;
                    const float dd = dx * dx + dy * dy;
                    if (dd < bd) { bd = dd; bidx = i; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(FULL, bd, o);
                const int oi = __shfl_xor_sync(FULL, bidx, o);
                if (ob < bd || (ob == bd && oi < bidx)) { bd = ob; bidx = oi; }
            }
            if (lane == src && bd < best) { best = bd; bi = bidx; }
        }
        if (upd) pidx = bi;
    }
    if (is_ego && !have_sc) sincosf_nc(p.h, &sn, &cs);             // dead / frozen egos (rare): their pose still takes part below
    // -- car-car override (:293-318)
    unsigned cmask = 0;                            // bit j: ego a collides with ego j > a (sub-warp numbering)
    for (int dlt = 1; dlt < N; ++dlt) {
        const float ox = __shfl_down_sync(FULL, p.x, dlt, NP), oy = __shfl_down_sync(FULL, p.y, dlt, NP);
        const float os = __shfl_down_sync(FULL, sn, dlt, NP), oc = __shfl_down_sync(FULL, cs, dlt, NP);
        // the cheap distance reject inline (it decides nearly every pair); only close pairs pay for the out-of-line SAT
        if (is_ego && !frozen && a + dlt < N && !cars_far_apart(p.x, p.y, ox, oy) && cars_collide_sc(p.x, p.y, sn, cs, ox, oy, os, oc)) cmask |= 1u << (a + dlt);
    }
    bool npc_hit = false;
    if (is_ego && !frozen) {
        for (int k = 0; k < c && !npc_hit; ++k) {
            const int ni = env * d.M + k;
            const float qx = d.nx[ni], qy = d.ny[ni];
            if (!cars_far_apart(p.x, p.y, qx, qy)) {
                float ns, ncs;
                sincosf_nc(d.nh[ni], &ns, &ncs);
                npc_hit = cars_collide_sc(p.x, p.y, sn, cs, qx, qy, ns, ncs);
            }
        }
    }
    {
        const unsigned alive_m = (__ballot_sync(FULL, is_ego && alive) >> shift) & LOW;
        unsigned done_m = (__ballot_sync(FULL, is_ego && done) >> shift) & LOW;
        const unsigned npc_m = (__ballot_sync(FULL, npc_hit) >> shift) & LOW;
        unsigned crash_m = 0;
        for (int i = 0; i < N; ++i) {
            const unsigned ci = __shfl_sync(FULL, cmask, i, NP);
            if (!((alive_m >> i) & 1u) || ((done_m >> i) & 1u)) continue;
            const unsigned m = ci & alive_m & ~done_m;
            if (m) { done_m |= m | (1u << i); crash_m |= m | (1u << i); }
            if ((npc_m >> i) & 1u) { done_m |= 1u << i; crash_m |= 1u << i; }
        }
        if (is_ego && ((crash_m >> a) & 1u)) { done = true; status = ISX_CRASH_CAR; }
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
        for (int i = 0; i < N; ++i) avg += __shfl_sync(FULL, rew, i, NP);
        avg /= (float)N;
        rew = (1.0f - d.rc.alpha) * rew + d.rc.alpha * avg;
    }
    // -- respawn / termination (:339-370)
    const unsigned alive_m = (__ballot_sync(FULL, is_ego && alive) >> shift) & LOW;
    const unsigned done_m = (__ballot_sync(FULL, is_ego && done) >> shift) & LOW;
    const unsigned succ_m = (__ballot_sync(FULL, is_ego && alive && done && status == ISX_SUCCESS) >> shift) & LOW;
    bool term = false;
    if (d.respawn) {
        if (is_ego && alive && done && (status == ISX_CRASH_CAR || status == ISX_CRASH_WALL || status == ISX_CRASH_LINE)) {
            const RouteMeta m = d.route_meta[a];                  // Car::respawn (Car.cpp:76-84)
            p.x = m.spawn_x; p.y = m.spawn_y; p.v = 0.0f; p.h = m.spawn_h;
            pidx = 0; pd = 0.0f; pa0 = 0.0f; pa1 = 0.0f; acc = 0.0f; steer = 0.0f;
            sincosf_nc(p.h, &sn, &cs);                              // the rectangle below is the respawned car's
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
        if (frozen) {                              // a fresh Lidar reads max range on every beam (Lidar.cpp:4-14)
            uint4* hz = reinterpret_cast<uint4*>(d.lidar_hit + (size_t)ai * ISX_MAX_RAYS);
#pragma unroll
            for (int i = 0; i < ISX_MAX_RAYS / 16; ++i) hz[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        // lidar pixel rectangle of the (possibly respawned) ego for k_features / k_lidar_obs
        reinterpret_cast<PixRect*>(d.car_rect)[(size_t)env * (N + d.M) + a] = car_pixel_rect_sc(p.x, p.y, sn, cs);
    }
    if (env_ok && a == 0) {
        d.step_count[env] = step_count; d.tick[env] = tick;
        d.terminated[env] = term ? 1 : 0; d.truncated[env] = trunc ? 1 : 0; d.agents_alive[env] = __popc(alive_m);
        if (!d.traffic && resets) { d.ncount[env] = 0; }
    }
    // ---- per-env counters (no atomics; reduced on demand by k_reduce_stats)
    {
        double rs = is_ego ? (double)rew : 0.0;
#pragma unroll
        for (int o = NP / 2; o > 0; o >>= 1) rs += __shfl_xor_sync(FULL, rs, o, NP);
        unsigned hist[6];
#pragma unroll
        for (int s6 = 0; s6 < 6; ++s6) hist[s6] = __popc((__ballot_sync(FULL, is_ego && !frozen && status == s6) >> shift) & LOW);
        if (env_ok && a == 0) {
            uint32_t* st = d.env_stats + (size_t)env * STAT_SLOTS;
#pragma unroll
            for (int s6 = 0; s6 < 6; ++s6) if (hist[s6]) st[ST_HIST0 + s6] += hist[s6];
            if (resets) st[ST_RESETS] += resets;
            if (!frozen) st[ST_STEPS] += (uint32_t)N;
            double* ps = reinterpret_cast<double*>(st + ST_RSUM);
            *ps += rs;
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_features
// One THREAD per ego.  Writes the 31 ego / neighbour features of its obs row (IntersectionEnv.cpp:431-508) and
// prepares the beam kernel's inputs: the ego's pose record, the lidar pixel rectangle of every car of the env
// (Lidar.cpp:65-78 as integer bounds) and, per ego, the packed list of cars a beam can possibly touch together with
// the angular beam window of each (beam_window).  Cheap (<2% of the step): thread-per-ego, no shared memory.
enum { LIDAR_MARCH = 0, LIDAR_FROM_HITS = 1 };
struct alignas(16) AgentRec { float x, y, h; int rect_base; };     // (one 16-byte load) rect_base = env * CE; -1: dead ego (zero row); -2: beams come from the stored hits;
                                                       // -3: the origin pixel is off screen (every beam breaks at distance 0, Lidar.cpp:38-40)

// Rare path of k_features: the reference's neighbour list of one ego (other alive egos in index order, then the NPCs in
// list order, :466-488), ordered by the restated libstdc++ std::sort; returns the car index (ego j, or N + NPC j; 0xff = none) of ranks 0..4,
// one byte each (by value, so the caller's rank registers never need an address).
// (plain pointers, not `const Dev&`: taking the address of the kernel parameter would copy it to local memory for everyone)
__device__ __noinline__ unsigned long long exact_neighbor_top5(const float* __restrict__ ex, const float* __restrict__ ey, const uint8_t* __restrict__ ealive,
                                                 const float* __restrict__ nx, const float* __restrict__ ny, int N, int M,
                                                 int env, int self, float mx, float my, int nn) {
    float key[ISX_MAX_AGENTS + ISX_MAX_NPC];
    uint8_t src[ISX_MAX_AGENTS + ISX_MAX_NPC], p[ISX_MAX_AGENTS + ISX_MAX_NPC];
    int m = 0;
    for (int j = 0; j < N; ++j) {
        const int g = env * N + j;
        if (j == self || !ealive[g]) continue;
        const float dx = ex[g] - mx, dy = ey[g] - my;
        key[m] = fsqrt_rn(dx * dx + dy * dy); src[m] = (uint8_t)j; p[m] = (uint8_t)m; ++m;
    }
    for (int j = 0; j < nn; ++j) {
        const int g = env * M + j;
        const float dx = nx[g] - mx, dy = ny[g] - my;
        key[m] = fsqrt_rn(dx * dx + dy * dy); src[m] = (uint8_t)(N + j); p[m] = (uint8_t)m; ++m;
    }
    stdsort::sort(key, p, m);
    unsigned long long packed = 0;
    for (int i = 0; i < 5; ++i) packed |= (unsigned long long)(i < m ? src[p[i]] : 0xffu) << (8 * i);
    return packed;
}


__global__ void __launch_bounds__(FEAT_THREADS, ISX_FEAT_MINB)
k_features(const Dev d, int mode) {
    // FOUR lanes per ego (a "quad"): sub-lane q handles the cars q, q+4, ... of the env, so the dependent chain per
    // thread is a quarter as long and there are 4x more warps in flight (thread-per-ego ran at 20% occupancy).
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, q = lane & 3, qshift = lane & ~3;
    const int N = d.N, R = d.R;
    const int agents = d.E * N;
    if (t == 0) *d.ray_counter = 0u;               // work counter of the k_lidar_obs launch that follows in the stream
    const bool ok = (t >> 2) < agents;
    const int ga = ok ? (t >> 2) : agents - 1;     // out-of-range quads still take part in the shuffles
    const int env = ga / N, self = ga - env * N;
    const int CE = N + d.M;
    const int nn = d.traffic ? d.ncount[env] : 0;
    const int nc = N + nn;
    const Pose me{d.ex[ga], d.ey[ga], d.ev[ga], d.eh[ga]};
    const bool alive = d.ealive[ga] != 0;
    // next-step auto-reset (k_ego): an env that was only reset by this call has step 0 — its beams are not marched, the
    // lidar part comes from the (cleared) stored hits, exactly like the observation reset() returns
    const bool fresh = d.auto_reset == 2 && mode == LIDAR_MARCH && d.step_count[env] == 0;
    if (ok && q == 0) {
        AgentRec rec;
        const bool off = (unsigned)f2i_rz(me.x) >= (unsigned)WIDTH || (unsigned)f2i_rz(me.y) >= (unsigned)HEIGHT;   // int(cx + dx * 0) of every beam
        rec.x = me.x; rec.y = me.y; rec.h = me.h; rec.rect_base = alive ? (fresh ? -2 : (off ? -3 : env * CE)) : -1;
        reinterpret_cast<AgentRec*>(d.agent_rec)[ga] = rec;
    }
    float bd[5]; int bk[5];
    int nbr = 0;                                   // neighbours this sub-lane saw (the quad total is the list length of :490)
#pragma unroll
    for (int i = 0; i < 5; ++i) { bd[i] = INFINITY; bk[i] = 0x7fffffff; }
    const int ipx = f2i_rz(me.x), ipy = f2i_rz(me.y);
    uint32_t* cand = d.cand + (size_t)ga * CE;
    const PixRect* rects = reinterpret_cast<const PixRect*>(d.car_rect) + (size_t)env * CE;
    int ncand = 0;
    const int nc_warp = __reduce_max_sync(FULL, nc);   // the ballot below needs every lane of the warp in the loop
    for (int k0 = 0; k0 < nc_warp; k0 += 4) {
        const int k = k0 + q;
        const bool have = ok && k < nc && alive;   // padding quads (!ok) alias the last agent: they must not touch its lists
        float ox = 0, oy = 0, oh = 0;
        bool k_alive = true;
        if (have) {
            if (k < N) { const int j = env * N + k; ox = d.ex[j]; oy = d.ey[j]; oh = d.eh[j]; k_alive = d.ealive[j] != 0; }
            else { const int j = env * d.M + (k - N); ox = d.nx[j]; oy = d.ny[j]; oh = d.nh[j]; }
        }
        // ---- beam candidates.  Lidar.cpp:57-63: the ego itself, and anything within 1e-3 of its pose, is transparent;
        //      beams reach at most 248 px (+1 px truncation) from the origin pixel
        bool is_cand = false;
        if (have && mode == LIDAR_MARCH && !fresh) {
            const bool same = fabsf(ox - me.x) < 1e-3f && fabsf(oy - me.y) < 1e-3f && fabsf(oh - me.h) < 1e-3f;
            if (!same) {
                const PixRect r = rects[k];        // written by k_ego / k_traffic earlier in this step
                is_cand = !(r.x0 > ipx + 250 || r.x1 < ipx - 250 || r.y0 > ipy + 250 || r.y1 < ipy - 250) && !pix_rect_empty(r);   // (clamped to the screen)
            }
        }
        const unsigned qb = (__ballot_sync(FULL, is_cand) >> qshift) & 0xFu;      // this quad's votes, in car order
        if (is_cand) cand[ncand + __popc(qb & ((1u << q) - 1u))] = (uint32_t)k;   // the beam window is added below
        ncand += __popc(qb);
        // ---- own part of the five nearest other alive cars (ascending distance, ties by list order; stable, :466-492)
        if (have && k != self && k_alive) {
            const float dx = ox - me.x, dy = oy - me.y;
            const float dist = fsqrt_rn(dx * dx + dy * dy);
            ++nbr;
            int pos = 0;                           // cars arrive in increasing k, so equal distances keep list order
#pragma unroll
            for (int i = 0; i < 5; ++i) pos += (bd[i] <= dist) ? 1 : 0;
#pragma unroll
            for (int i = 4; i >= 1; --i) if (i > pos) { bd[i] = bd[i - 1]; bk[i] = bk[i - 1]; }
#pragma unroll
            for (int i = 0; i < 5; ++i) if (i == pos) { bd[i] = dist; bk[i] = k; }
        }
    }
    if (ok && q == 0) d.cand_n[ga] = ncand;
    // Beam windows of the compacted candidates, four at a time: ~300 instructions each (atan2 + two atan), so they run on
    // the dense list — every lane of the quad busy — rather than inside the sparse in-range test above.
    __syncwarp();
    if (ok) {
        for (int e = q; e < ncand; e += 4) {
            const uint32_t k = cand[e];
            const BeamWindow w = beam_window(rects[k], me.x, me.y, me.h, R);
            cand[e] = k | ((uint32_t)w.ia << 8) | ((uint32_t)w.span << 16) | ((uint32_t)w.kmin << 24);
        }
    }
    // The 31 features (+ alive flag) of the warp's 8 egos are staged in shared memory and leave as whole-row stores: 31
    // scattered 4-byte stores per ego and copy cost a 32-byte L2 sector each (1.4 GB of L2 traffic per launch at 65,536 envs).
    __shared__ float s_out[FEAT_THREADS / 32][8][33];
    float* so = s_out[threadIdx.x >> 5][lane >> 2];
    // ---- merge the four sorted partial lists: five rounds of a quad-wide lexicographic (distance, list index) minimum,
    //      plus a sixth minimum that is only looked at for the tie test below
    float fd[6]; int fk[5];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        float md = bd[0]; int mk = bk[0];
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            const float od = __shfl_xor_sync(FULL, md, o, 4);
            const int okk = __shfl_xor_sync(FULL, mk, o, 4);
            if (od < md || (od == md && okk < mk)) { md = od; mk = okk; }
        }
        fd[r] = md;
        if (r < 5) {
            fk[r] = mk;
            if (mk == bk[0] && mk != 0x7fffffff) {     // this lane's head won: pop it
#pragma unroll
                for (int i = 0; i < 4; ++i) { bd[i] = bd[i + 1]; bk[i] = bk[i + 1]; }
                bd[4] = INFINITY; bk[4] = 0x7fffffff;
            }
        }
    }
    // The list order above is what a STABLE sort gives.  std::sort (:490) is stable only up to 16 elements; beyond that,
    // exactly equal distances among the first six ranks are ordered by libstdc++'s introsort — replay it (rare).
    nbr += __shfl_xor_sync(FULL, nbr, 1, 4);
    nbr += __shfl_xor_sync(FULL, nbr, 2, 4);
    if (nbr > stdsort::THRESHOLD) {
        bool tie = false;
#pragma unroll
        for (int r = 0; r < 5; ++r) tie |= (fd[r] == fd[r + 1]) && fd[r + 1] < INFINITY;
        if (tie) {                                 // quad-uniform; every lane of the quad replays it (no shuffles inside)
            const unsigned long long ex5 = exact_neighbor_top5(d.ex, d.ey, d.ealive, d.nx, d.ny, N, d.M, env, self, me.x, me.y, nn);
#pragma unroll
            for (int r = 0; r < 5; ++r) { const int kk = (int)((ex5 >> (8 * r)) & 0xffu); fk[r] = kk == 0xff ? 0x7fffffff : kk; }
            if (ok && q == 0) atomicAdd(d.env_stats + (size_t)env * STAT_SLOTS + ST_TIESORT, 1u);
        }
    }
    // neighbour slots 0..3 are written by sub-lanes 0..3 in ONE converged pass (their pose loads overlap), slot 4 by
    // sub-lane 3 in a second; unused slots stay zero (:424)
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        const int r = it == 0 ? q : (q == 3 ? 4 : -1);
        if (r < 0) continue;
        const int mk = r == 0 ? fk[0] : r == 1 ? fk[1] : r == 2 ? fk[2] : r == 3 ? fk[3] : fk[4];
        float f5[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (ok && mk != 0x7fffffff) {
            Pose ot; int intent;
            if (mk < N) { const int j = env * N + mk; ot = Pose{d.ex[j], d.ey[j], d.ev[j], d.eh[j]}; intent = d.route_meta[mk].intent; }
            else { const int j = env * d.M + (mk - N); ot = Pose{d.nx[j], d.ny[j], d.nv[j], d.nh[j]}; intent = d.route_meta[N + d.nroute[j]].intent; }
            obs_neighbor_features(me, ot, intent, f5);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) so[6 + 5 * r + i] = f5[i];
    }
    if (q == 1) {                                  // the six ego features (:431-458); zeros for a dead ego (:426-429)
        float f6[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (ok && alive) {
            const F2* path = d.route_path + (size_t)self * PATH_LEN;
            obs_ego_features(me, path[min(d.epidx[ga] + 10, PATH_LEN - 1)], f6);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) so[i] = f6[i];
        so[31] = alive ? 1.0f : 0.0f;
    }
    __syncwarp();
    {
        const int ga0 = (int)((blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) >> 2);   // first ego of this warp
        const int nag = min(8, agents - ga0);                                            // <= 0: nothing in range
        const float(*sw)[33] = s_out[threadIdx.x >> 5];
        float* crow = d.obs_c + (size_t)ga0 * 32;  // compact record for the host-buffer step: 32 floats per ego, contiguous over egos
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            if (a < nag) {
                const float v = sw[a][lane];
                crow[a * 32 + lane] = v;
                if (lane < 31) d.obs[(size_t)(ga0 + a) * ISX_OBS_DIM + lane] = v;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------ k_lidar_obs
// Road march for the 32 rays of a warp (must be called by all 32 lanes, converged).  Every lane first jumps over the
// samples the analytic bound proves free of road events (ray_safe_samples: more than half of all rays are finished by the
// jump alone — nothing within range), then tests the next samples with the exact pixel arithmetic in lock-step: 76% of
// the rays are done after one test, 98.6% after two, 99.6% after three (tools/march_stats.py).  The few rays still open
// (grazing a wall) are finished ONE AT A TIME by the whole warp, lane j testing sample k+1+j — so a warp never idles 31
// lanes while one grazing ray crawls along a wall.
__device__ __forceinline__ int warp_road_event(bool active, const RoadBitsShared bits, const RoadAna& ra, const Ray& r, bool* hit, int lane) {
    March m;                                         // the origin pixel is on screen (k_features sorts the other egos out)
    m.px = 0; m.py = 0; m.ke = LIDAR_MAX_K + 1; m.done = !active; m.hit = false;
    m.k = ray_safe_samples(ra, r);
#pragma unroll
    for (int it = 0; it < LOCKSTEP; ++it) {
#if ISX_TEST_BRANCHFREE
        // every lane evaluates the next sample (nearly every warp has an open ray anyway); done lanes discard the result.
        // The sample index and the folded pixel are clamped so that the table read stays in range for discarded lanes.
        const int k = min(m.k + 1, LIDAR_MAX_K);
        int px, py;
        ray_pixel(r.cx, r.cy, r.dx, r.dy, k, px, py);
        const bool off = (unsigned)px >= (unsigned)WIDTH || (unsigned)py >= (unsigned)HEIGHT;
        const int u = min(abs(px - ROAD_HALF), ROAD_HALF), v = min(abs(py - ROAD_HALF), ROAD_HALF);
        const bool road = (bits.word(v * ROAD_WORDS + (u >> 5)) >> (u & 31)) & 1u;
        const bool beyond = m.k + 1 > LIDAR_MAX_K;
        const bool ev = beyond || off || !road;
        if (!m.done) {
            m.k += 1;
            if (ev) { m.ke = beyond ? LIDAR_MAX_K + 1 : m.k; m.hit = !beyond && !off; m.done = true; }
        }
#else
        if (!m.done) march_next(bits, r, m);
#endif
    }
    unsigned pend = __ballot_sync(FULL, !m.done);
#pragma unroll 1
    for (int it = 0; it < LOCKSTEP_EXTRA && __popc(pend) >= LOCKSTEP_MIN_OPEN; ++it) {
        if (!m.done) march_next(bits, r, m);
        pend = __ballot_sync(FULL, !m.done);
    }
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
            if (kk <= LIDAR_MAX_K) e = sample_event(bits, o, kk, px, py);
            const unsigned b = __ballot_sync(FULL, e != 0);
            if (b) { const int f = __ffs(b) - 1; found = base + f; fe = __shfl_sync(FULL, e, f); break; }
        }
        if (lane == src) { m.ke = found; m.hit = (fe == 2); m.done = true; }
    }
    *hit = m.hit;
    return m.ke;
}

// One THREAD per (ego, beam), beams of all egos laid end to end; persistent CTAs walk 256-beam chunks, so the only
// block-level synchronisation is the one after the road tables are staged in shared memory (18 KB per CTA: folded
// bitmap + beam angles).  Everything per ego comes from k_features through L1/L2.
// Rare path of k_lidar_obs, out of line so that its address arithmetic is not predicated into every piece: a beam whose
// value comes from the stored hit (refresh modes), or — off-screen origin — is "nothing within range" by definition.
__device__ __noinline__ int unmarched_hit(const Dev& d, bool special, bool offscreen, unsigned hit_index) {
    if (!special) return -1;
    if (offscreen) { d.lidar_hit[hit_index] = 0; return 0; }
    return d.lidar_hit[hit_index];
}

template <int RT, bool WHOLE>   // RT = beam count known at compile time (72, 96) or 0 = run-time d.R; WHOLE: the beam total is a multiple of 32 (no padding lanes)
__global__ void __launch_bounds__(LID_THREADS, ISX_LID_MINB)
k_lidar_obs(const Dev d, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem_raw);
    float* s_rel = reinterpret_cast<float*>(smem_raw + ROAD_BITS_BYTES);
    const int tid = threadIdx.x, lane = tid & 31;
    const int R = RT ? RT : d.R;
    pdl_launch_dependents();
    {   // 16-byte copies; the global tables are padded to ROAD_*_BYTES (constant since isx_create: safe before pdl_wait)
        const uint4* gb = reinterpret_cast<const uint4*>(d.road_bits);
        uint4* sb = reinterpret_cast<uint4*>(s_bits);
        for (int i = tid; i < ROAD_BITS_BYTES / 16; i += LID_THREADS) sb[i] = gb[i];
        for (int i = tid; i < R; i += LID_THREADS) s_rel[i] = d.rel_angle[i];
    }
    pdl_wait();                                      // k_features' records, candidate lists and the zeroed work counter
    __syncthreads();
    const RoadBitsShared road_bits{(uint32_t)__cvta_generic_to_shared(s_bits)};
    const uint32_t rel_addr = (uint32_t)__cvta_generic_to_shared(s_rel);
    const int CE = d.N + d.M;
    const int total = d.E * d.N * R;                 // < 2^31, checked by isx_create
    const AgentRec* recs = reinterpret_cast<const AgentRec*>(d.agent_rec);
    const PixRect* rects = reinterpret_cast<const PixRect*>(d.car_rect);
    // Dynamic work distribution at WARP granularity.  A warp's total time is a sum of very uneven 32-beam pieces
    // (open road vs. wall vs. off screen); with static striding the slowest of ~9,500 warps — a +3.7 sigma outlier —
    // set the kernel time at 67% average occupancy.  Each warp now claims WARP_GRAB consecutive 32-beam pieces at a
    // time from one global counter (zeroed by k_features, which always runs just before) until the beams run out.
    const int pieces = (total + 31) / 32;
    // pieces per claim: WARP_GRAB when every warp gets many claims; small batches (a few pieces per warp) claim one at a time,
    // otherwise a third of the warps would do all the work
    const int warps_total = (int)gridDim.x * (LID_THREADS / 32);
    const int grab = pieces >= 12 * warps_total ? WARP_GRAB : (pieces >= 4 * warps_total ? 2 : 1);
    while (true) {
        int p0 = 0;
        if (lane == 0) p0 = (int)atomicAdd(d.ray_counter, (unsigned)grab);
        p0 = __shfl_sync(FULL, p0, 0);
        if (p0 >= pieces) break;
        const int p1 = min(p0 + grab, pieces);
        for (int pc = p0; pc < p1; ++pc) {
            // all global indices below fit 32 unsigned bits (isx_create: E*N*96 < 2^31), so addresses are one
            // base + 32-bit offset multiply-add each instead of 64-bit index arithmetic
            const int id = pc * 32 + lane;
            const bool valid = WHOLE || id < total;
            const unsigned ga = valid ? (unsigned)(id / R) : 0u;
            const int i = valid ? id - (int)ga * R : 0;
            const AgentRec rec = recs[ga];
            const bool alive = valid && rec.rect_base >= 0;
            float out = 0.0f;                                   // dead ego: all-zero row (:426-429)
            int kout = 0;                                       // hit index behind `out` (compact copy for the host-buffer step)
            // beams that are not marched (rare, so behind a warp-uniform branch): isx_observe / set_state refresh every live ego
            // from the stored hits; so does an env that this call only reset in next-step auto-reset mode (rect_base == -2); an
            // ego whose origin pixel is off screen reads \"nothing within range\" on every beam (-3)
            const bool special = valid && (mode == LIDAR_FROM_HITS ? rec.rect_base != -1 : rec.rect_base < -1);
            if (__any_sync(FULL, special)) {
                const int k = unmarched_hit(d, special, mode != LIDAR_FROM_HITS && rec.rect_base == -3, ga * (unsigned)ISX_MAX_RAYS + (unsigned)i);
                if (k >= 0) { out = (k ? (float)(4 * k) : LIDAR_MAX_DIST) * (1.0f / LIDAR_MAX_DIST); kout = k; }
            }
            if (mode != LIDAR_FROM_HITS) {
                // every lane builds its ray from the record it loaded (lanes past the end read ego 0): dead or padding
                // lanes are simply not marched, and no second, constant ray has to be materialised
                float sn, cs, rel;
                asm("ld.shared.f32 %0, [%1];" : "=f"(rel) : "r"(rel_addr + 4u * (unsigned)i));
                sincosf_(rec.h + rel, &sn, &cs);
                const Ray ray = make_ray(rec.x, rec.y, cs, -sn);
                bool hit;
                const int ke = warp_road_event(alive, road_bits, d.ana, ray, &hit, lane);
                if (alive) {
                    int best = hit ? ke : 0;
                    int lim = ke - 1;                           // cars only count strictly before the road event
                    const int nc = d.cand_n[ga];
                    const unsigned cbase = ga * (unsigned)CE;
                    const int iw = (i == R - 1) ? 0 : i;        // beam R-1 duplicates beam 0
                    for (int j = 0; j < nc && lim >= 1; ++j) {
                        const uint32_t ci = d.cand[cbase + (unsigned)j];
                        if ((int)(ci >> 24) > lim) continue;     // the car lies beyond what this beam can still see (kmin)
                        int dlt = iw - (int)__byte_perm(ci, 0u, 0x4441u);  // angular window of this car (beam_window); span 255 = all
                        dlt += (dlt >> 31) & (R - 1);
                        if (dlt > (int)__byte_perm(ci, 0u, 0x4442u)) continue;
                        const int kh = ray_rect_first_hit(rects[(unsigned)rec.rect_base + (ci & 255u)], ray, lim);
                        if (kh) { best = kh; lim = kh - 1; }
                    }
                    d.lidar_hit[ga * (unsigned)ISX_MAX_RAYS + (unsigned)i] = (uint8_t)best;
                    out = (best ? (float)(4 * best) : LIDAR_MAX_DIST) * (1.0f / LIDAR_MAX_DIST);
                    kout = best;
                }
            }
            if (valid) {
                d.obs[ga * (unsigned)ISX_OBS_DIM + 31u + (unsigned)i] = out;
                d.hit_c[(unsigned)id] = (uint8_t)kout;         // id == ga * R + i: one coalesced byte per lane
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_render
// Headless debug picture of ONE env (stands in for the Windows-only window of Renderer.cpp:520-646):
// thread per pixel of a 750x750 RGB image.  Road surface / grass come from the same folded bitmap the beams march,
// centre-line pixels from is_line_px, egos (alive ones) in a six-colour palette with a dark head marker, NPCs grey,
// and — as draw_lidar does — only the beams that hit something: a green segment with a red end point.  The CTA first
// stages car frames and beam end points (exact hit pixel: ray_pixel of sample k) in shared memory.
struct RenderCar { float x, y, c, s; uint32_t rgb; int npc; };
struct RenderBeam { short x0, y0, x1, y1; };
constexpr int RENDER_MAX_CARS = ISX_MAX_AGENTS + ISX_MAX_NPC;
constexpr int RENDER_MAX_BEAMS = 1024;     // per staging round

__device__ __forceinline__ uint32_t rgb_pack(int r, int g, int b) { return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16); }

__global__ void __launch_bounds__(256)
k_render(const Dev d, int env, uint8_t* __restrict__ rgb) {
    __shared__ RenderCar s_car[RENDER_MAX_CARS];
    __shared__ RenderBeam s_beam[RENDER_MAX_BEAMS];
    __shared__ int s_ncar, s_nbeam;
    const int tid = threadIdx.x;
    const int pix = blockIdx.x * blockDim.x + tid;
    const int x = pix % WIDTH, y = pix / WIDTH;
    const bool inside = pix < WIDTH * HEIGHT;
    const int N = d.N, R = d.R;
    const int nn = d.traffic ? d.ncount[env] : 0;
    if (tid == 0) { s_ncar = 0; s_nbeam = 0; }
    __syncthreads();
    for (int k = tid; k < N + nn; k += blockDim.x) {
        RenderCar c;
        bool draw = true;
        if (k < N) {
            const int j = env * N + k;
            draw = d.ealive[j] != 0;
            const uint32_t pal[6] = {rgb_pack(231, 76, 60), rgb_pack(52, 152, 219), rgb_pack(46, 204, 113),
                                     rgb_pack(155, 89, 182), rgb_pack(241, 196, 15), rgb_pack(230, 126, 34)};
            float sn, cs;
            sincosf_(d.eh[j], &sn, &cs);
            c = RenderCar{d.ex[j], d.ey[j], cs, sn, pal[k % 6], 0};
        } else {
            const int j = env * d.M + (k - N);
            float sn, cs;
            sincosf_(d.nh[j], &sn, &cs);
            c = RenderCar{d.nx[j], d.ny[j], cs, sn, rgb_pack(128, 128, 128), 1};
        }
        if (!draw) c.npc = -1;
        s_car[k] = c;                              // list order = paint order (egos, then NPCs), as draw_cars
    }
    if (tid == 0) s_ncar = N + nn;
    // base colour
    uint32_t col = 0;
    if (inside) {
        int u = x - ROAD_HALF; u = u < 0 ? -u : u;
        int v = y - ROAD_HALF; v = v < 0 ? -v : v;
        const bool road = (d.road_bits[v * ROAD_WORDS + (u >> 5)] >> (u & 31)) & 1u;
        col = road ? rgb_pack(70, 70, 74) : rgb_pack(58, 125, 68);
        if (road && is_line_px(d.lanes, x, y)) col = rgb_pack(240, 200, 40);
    }
    // beams that hit, RENDER_MAX_BEAMS at a time
    const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
    bool on_beam = false, on_tip = false;
    for (int base = 0; base < N * R; base += RENDER_MAX_BEAMS) {
        __syncthreads();
        if (tid == 0) s_nbeam = 0;
        __syncthreads();
        for (int b = base + tid; b < N * R && b < base + RENDER_MAX_BEAMS; b += blockDim.x) {
            const int a = b / R, i = b - a * R;
            const int j = env * N + a;
            const int k = d.ealive[j] ? d.lidar_hit[(size_t)j * ISX_MAX_RAYS + i] : 0;
            if (k) {
                float sn, cs;
                sincosf_(d.eh[j] + d.rel_angle[i], &sn, &cs);
                int px, py;
                ray_pixel(d.ex[j], d.ey[j], cs, -sn, k, px, py);
                s_beam[atomicAdd(&s_nbeam, 1)] = RenderBeam{(short)f2i_rz(d.ex[j]), (short)f2i_rz(d.ey[j]), (short)px, (short)py};
            }
        }
        __syncthreads();
        const int nb = s_nbeam;
        for (int b = 0; b < nb && inside; ++b) {
            const RenderBeam q = s_beam[b];
            const float ax = (float)q.x0 + 0.5f, ay = (float)q.y0 + 0.5f, bx = (float)q.x1 + 0.5f, by = (float)q.y1 + 0.5f;
            const float ex = fx - bx, ey = fy - by;
            if (ex * ex + ey * ey <= 2.5f * 2.5f) on_tip = true;
            const float vx = bx - ax, vy = by - ay, wx = fx - ax, wy = fy - ay;
            const float len2 = vx * vx + vy * vy;
            float t = len2 > 0.0f ? (wx * vx + wy * vy) / len2 : 0.0f;
            t = fminf(fmaxf(t, 0.0f), 1.0f);
            const float qx = wx - t * vx, qy = wy - t * vy;
            if (qx * qx + qy * qy <= 0.75f * 0.75f) on_beam = true;
        }
    }
    __syncthreads();
    if (!inside) return;
    if (on_beam) col = rgb_pack(60, 220, 90);
    // cars on top of the beams, end points on top of everything (later cars overwrite earlier ones, like the painter)
    const float hl = CAR_LENGTH * 0.5f, hw = CAR_WIDTH * 0.5f;
    for (int k = 0; k < s_ncar; ++k) {
        const RenderCar c = s_car[k];
        if (c.npc < 0) continue;
        const float dx = fx - c.x, dy = fy - c.y;
        const float lx = dx * c.c - dy * c.s, ly = dx * c.s + dy * c.c;      // screen -> car frame (screen y points down)
        if (fabsf(lx) <= hl && fabsf(ly) <= hw) {
            col = c.rgb;
            if (lx >= -hl + 0.70f * CAR_LENGTH && lx <= -hl + 0.95f * CAR_LENGTH && fabsf(ly) <= hw - 2.0f)
                col = c.npc ? rgb_pack(20, 20, 20) : rgb_pack(250, 250, 250);
        }
    }
    if (on_tip) col = rgb_pack(220, 30, 30);
    uint8_t* o = rgb + (size_t)pix * 3;
    o[0] = (uint8_t)(col & 255u); o[1] = (uint8_t)((col >> 8) & 255u); o[2] = (uint8_t)((col >> 16) & 255u);
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
    // one CTA of 1024 threads; 64-bit sums of the per-env u32 counters, reward_sum in double.  Fixed summation order
    // (strided partials, then a shared-memory tree), so that the totals — reward_sum included — are bit-reproducible.
    __shared__ unsigned long long acc[STAT_SLOTS];
    __shared__ double rpart[1024];
    if (threadIdx.x < STAT_SLOTS) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long loc[14];
    double r = 0.0;
    for (int i = 0; i < 14; ++i) loc[i] = 0;
    for (int e = threadIdx.x; e < d.E; e += blockDim.x) {
        const uint32_t* st = d.env_stats + (size_t)e * STAT_SLOTS;
        for (int i = 0; i < 14; ++i) loc[i] += st[i];
        r += *reinterpret_cast<const double*>(st + ST_RSUM);
    }
    for (int i = 0; i < 14; ++i) atomicAdd(&acc[i], loc[i]);       // integer sums: order does not matter
    rpart[threadIdx.x] = r;
    __syncthreads();
    for (int half = 512; half > 0; half >>= 1) {
        if ((int)threadIdx.x < half && (int)threadIdx.x + half < (int)blockDim.x) rpart[threadIdx.x] += rpart[threadIdx.x + half];
        __syncthreads();
    }
    if (threadIdx.x < 14) d.stats[threadIdx.x] = acc[threadIdx.x];
    if (threadIdx.x == 14) d.stats[14] = 0;
    if (threadIdx.x == 15) d.stats[15] = (unsigned long long)__double_as_longlong(rpart[0]);
}

// Masked restore of a snapshot (isx_snapshot_restore): thread per (env, array); copies the env's slice of the array.
struct SnapArray { unsigned char* live; const unsigned char* saved; unsigned bytes_per_env; };
__global__ void k_snapshot_restore(const SnapArray* __restrict__ tab, int n_arrays, const uint8_t* __restrict__ mask, int E) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * n_arrays) return;
    const int env = t / n_arrays, a = t - env * n_arrays;
    if (mask && !mask[env]) return;
    const SnapArray s = tab[a];
    const size_t off = (size_t)env * s.bytes_per_env;
    unsigned char* dst = s.live + off;
    const unsigned char* src = s.saved + off;
    if ((s.bytes_per_env & 3u) == 0 && ((reinterpret_cast<size_t>(dst) | reinterpret_cast<size_t>(src)) & 3u) == 0) {
        for (unsigned i = 0; i < s.bytes_per_env; i += 4) *reinterpret_cast<uint32_t*>(dst + i) = *reinterpret_cast<const uint32_t*>(src + i);
    } else {
        for (unsigned i = 0; i < s.bytes_per_env; ++i) dst[i] = src[i];
    }
}

// contraction canary: (a*b + c) with operands chosen so that a fused multiply-add gives a different float
__global__ void k_canary(float a, float b, float c, float* out) { out[0] = a * b + c; double x = a, y = b, z = c; out[1] = (float)(x * y + z); }

// Car::update / Car::check_collision on detached car records (the unit-level methods bindings.cpp:30-31 exposes on `Car`):
// one thread, the same device functions the step kernels call.  io6 = {x, y, v, heading, acc, steering_angle}.
__global__ void k_car_unit(int op, float* io6, const float* other3, float thr, float st, float dt, int* flag) {
    if (op == 0) {
        Pose p{io6[0], io6[1], io6[2], io6[3]};
        float acc = io6[4], steer = io6[5];
        car_update(p, steer, acc, thr, st, dt);
        io6[0] = p.x; io6[1] = p.y; io6[2] = p.v; io6[3] = p.h; io6[4] = acc; io6[5] = steer;
    } else {
        // no far-apart shortcut semantics change: cars_collide is exact (see cars_far_apart)
        *flag = cars_collide(io6[0], io6[1], io6[3], other3[0], other3[1], other3[2]) ? 1 : 0;
    }
}

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
size_t lidar_smem_bytes(const Dev& d) { (void)d; return (size_t)ROAD_BITS_BYTES + sizeof(float) * ISX_MAX_RAYS; }
size_t road_bits_bytes() { return ROAD_BITS_BYTES; }
size_t road_skip_bytes() { return ROAD_SKIP_BYTES; }

// Launch with the programmatic-stream-serialization attribute (see pdl_wait above).  ISX_NO_PDL=1 in the environment
// falls back to fully serialised launches (A/B and bisecting aid).
static bool pdl_enabled() {
    static const bool on = getenv("ISX_NO_PDL") == nullptr;
    return on;
}
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

cudaError_t launch_traffic(const Dev& d, float dt, float spawn_prob, cudaStream_t st) {
    if (!d.traffic) return cudaSuccess;
    // lanes per env by batch size (ISX_TRAFFIC_LANES = 8 / 16 / 32 forces one): small batches are bound by the list-order chain
    // of the busiest env, which is shortest with a whole warp per env; large batches by the warps to retire, and the packed
    // instances take their envs from the lists of k_traffic_order.  Measured with the lists (us at 32 / 16 / 8 lanes; profiles/r02/
    // ab_traffic_lanes_by_batch_ordered.log): 8192 envs 27.9 / 33.6 / 44.6, 12288: 33.6 / 37.4 / 49.1, 16384: 40.4 / 40.9 / 52.6,
    // 24576: 54.0 / 45.1 / 56.9, 32768: 66.4 / 50.3 / 59.4, 65536: - / 82.8 / 70.8
    const int forced = TRAFFIC_LANES ? TRAFFIC_LANES : d.traffic_lanes;
    const int lanes = forced ? forced : (d.E <= 12288 ? 32 : d.E <= 49152 ? 16 : 8);
    if (lanes < 32 && d.order != nullptr) {           // envs that share a warp: filed by NPC count first (k_traffic_order)
        const cudaError_t e = launch_pdl(k_traffic_order, (d.E + ORDER_THREADS - 1) / ORDER_THREADS, ORDER_THREADS, 0, st, d);
        if (e != cudaSuccess) return e;
    }
    if (lanes == 8) {                                 // four envs per warp (wide fallback inside the kernel)
        const int blocks = (d.E + 4 * DYN_WARPS - 1) / (4 * DYN_WARPS);
        return launch_pdl(k_traffic<8>, blocks, DYN_WARPS * 32, 0, st, d, dt, spawn_prob);
    }
    if (lanes == 16) {                                // two envs per warp
        const int blocks = (d.E + 2 * DYN_WARPS - 1) / (2 * DYN_WARPS);
        return launch_pdl(k_traffic<16>, blocks, DYN_WARPS * 32, 0, st, d, dt, spawn_prob);
    }
    const int blocks = (d.E + DYN_WARPS - 1) / DYN_WARPS;
    return launch_pdl(k_traffic<32>, blocks, DYN_WARPS * 32, 0, st, d, dt, spawn_prob);
}
cudaError_t launch_ego(const Dev& d, const float* actions, float dt, cudaStream_t st) {
    const int NP = d.N <= 1 ? 1 : d.N <= 2 ? 2 : d.N <= 4 ? 4 : d.N <= 8 ? 8 : d.N <= 16 ? 16 : 32;
    const long long threads = ((long long)d.E * NP + 31) / 32 * 32;
    const int blocks = (int)((threads + EGO_THREADS - 1) / EGO_THREADS);
    switch (NP) {
        case 1: return launch_pdl(k_ego<1>, blocks, EGO_THREADS, 0, st, d, actions, dt);
        case 2: return launch_pdl(k_ego<2>, blocks, EGO_THREADS, 0, st, d, actions, dt);
        case 4: return launch_pdl(k_ego<4>, blocks, EGO_THREADS, 0, st, d, actions, dt);
        case 8: return launch_pdl(k_ego<8>, blocks, EGO_THREADS, 0, st, d, actions, dt);
        case 16: return launch_pdl(k_ego<16>, blocks, EGO_THREADS, 0, st, d, actions, dt);
        default: return launch_pdl(k_ego<32>, blocks, EGO_THREADS, 0, st, d, actions, dt);
    }
}
cudaError_t launch_dynamics(const Dev& d, const float* actions, float dt, float spawn_prob, cudaStream_t st) {
    cudaError_t e = launch_traffic(d, dt, spawn_prob, st);
    if (e != cudaSuccess) return e;
    return launch_ego(d, actions, dt, st);
}
cudaError_t launch_features(const Dev& d, int mode, cudaStream_t st) {
    const int agents = d.E * d.N;
    return launch_pdl(k_features, (agents * 4 + FEAT_THREADS - 1) / FEAT_THREADS, FEAT_THREADS, 0, st, d, mode);
}
cudaError_t launch_rays(const Dev& d, int mode, int grid_cap, cudaStream_t st) {
    const long long total = (long long)d.E * d.N * d.R;
    const long long chunks = (total + LID_THREADS - 1) / LID_THREADS;
    const int grid = (int)(chunks < grid_cap ? chunks : grid_cap);
    const size_t sm = lidar_smem_bytes(d);
    if (total % 32 == 0) {
        if (d.R == 72) return launch_pdl(k_lidar_obs<72, true>, grid, LID_THREADS, sm, st, d, mode);
        if (d.R == 96) return launch_pdl(k_lidar_obs<96, true>, grid, LID_THREADS, sm, st, d, mode);
    }
    if (d.R == 72) return launch_pdl(k_lidar_obs<72, false>, grid, LID_THREADS, sm, st, d, mode);
    if (d.R == 96) return launch_pdl(k_lidar_obs<96, false>, grid, LID_THREADS, sm, st, d, mode);
    return launch_pdl(k_lidar_obs<0, false>, grid, LID_THREADS, sm, st, d, mode);
}
cudaError_t launch_lidar_obs(const Dev& d, int mode, int grid_cap, cudaStream_t st) {
    cudaError_t e = launch_features(d, mode, st);
    if (e != cudaSuccess) return e;
    return launch_rays(d, mode, grid_cap, st);
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
cudaError_t launch_snapshot_restore(const void* tab, int n_arrays, const uint8_t* mask, int E, cudaStream_t st) {
    const long long n = (long long)E * n_arrays;
    k_snapshot_restore<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const SnapArray*>(tab), n_arrays, mask, E);
    return cudaGetLastError();
}
cudaError_t launch_render(const Dev& d, int env, uint8_t* rgb, cudaStream_t st) {
    k_render<<<(WIDTH * HEIGHT + 255) / 256, 256, 0, st>>>(d, env, rgb);
    return cudaGetLastError();
}
cudaError_t launch_canary(float a, float b, float c, float* out, cudaStream_t st) {
    k_canary<<<1, 1, 0, st>>>(a, b, c, out);
    return cudaGetLastError();
}
cudaError_t launch_car_unit(int op, float* io6, const float* other3, float thr, float st, float dt, int* flag, cudaStream_t stm) {
    k_car_unit<<<1, 1, 0, stm>>>(op, io6, other3, thr, st, dt, flag);
    return cudaGetLastError();
}
cudaError_t launch_math_probe(int n, const float* a, const float* b, float* sn, float* cs, float* tn, float* at, float* hy, float* wr, cudaStream_t st) {
    k_math_probe<<<(n + 255) / 256, 256, 0, st>>>(n, a, b, sn, cs, tn, at, hy, wr);
    return cudaGetLastError();
}
cudaError_t lidar_set_smem_attr(const Dev& d) {
    const int b = (int)lidar_smem_bytes(d);
    cudaError_t e = cudaFuncSetAttribute(k_lidar_obs<72, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lidar_obs<96, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lidar_obs<72, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lidar_obs<96, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lidar_obs<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    return e;
}
cudaError_t lidar_occupancy(const Dev& d, int* ctas_per_sm) {
    const bool whole = ((long long)d.E * d.N * d.R) % 32 == 0;     // the instance launch_rays picks
    if (whole && d.R == 72) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<72, true>, LID_THREADS, lidar_smem_bytes(d));
    if (whole && d.R == 96) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<96, true>, LID_THREADS, lidar_smem_bytes(d));
    if (d.R == 72) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<72, false>, LID_THREADS, lidar_smem_bytes(d));
    if (d.R == 96) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<96, false>, LID_THREADS, lidar_smem_bytes(d));
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_lidar_obs<0, false>, LID_THREADS, lidar_smem_bytes(d));
}

}  // namespace isx
