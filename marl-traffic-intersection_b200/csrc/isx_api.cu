// isx_api.cu — C ABI of libisx_b200.so (include/isx.h): handle lifetime, table upload, launches,
// host<->device staging.  All simulation work happens in isx_kernels.cu on the GPU; nothing here steps an env.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>

#include "isx_device.cuh"
#include "isx_host_expand.h"
#include "isx_tables.h"

namespace isx {
size_t lidar_smem_bytes(const Dev& d);
size_t road_bits_bytes();
size_t road_skip_bytes();
cudaError_t launch_dynamics(const Dev& d, const float* actions, float dt, float spawn_prob, cudaStream_t st);
cudaError_t launch_lidar_obs(const Dev& d, int mode, int grid_cap, cudaStream_t st);
cudaError_t launch_traffic(const Dev& d, float dt, float spawn_prob, cudaStream_t st);
cudaError_t launch_ego(const Dev& d, const float* actions, float dt, cudaStream_t st);
cudaError_t launch_features(const Dev& d, int mode, cudaStream_t st);
cudaError_t launch_rays(const Dev& d, int mode, int grid_cap, cudaStream_t st);
cudaError_t launch_reset(const Dev& d, const uint8_t* mask, cudaStream_t st);
cudaError_t launch_reduce_stats(const Dev& d, cudaStream_t st);
cudaError_t launch_canary(float a, float b, float c, float* out, cudaStream_t st);
cudaError_t launch_render(const Dev& d, int env, uint8_t* rgb, cudaStream_t st);
cudaError_t launch_snapshot_restore(const void* tab, int n_arrays, const uint8_t* mask, int E, cudaStream_t st);
cudaError_t launch_math_probe(int n, const float* a, const float* b, float* sn, float* cs, float* tn, float* at, float* hy, float* wr, cudaStream_t st);
cudaError_t launch_car_unit(int op, float* io6, const float* other3, float thr, float st, float dt, int* flag, cudaStream_t st_);
cudaError_t lidar_set_smem_attr(const Dev& d);
cudaError_t lidar_occupancy(const Dev& d, int* ctas_per_sm);
}  // namespace isx

using namespace isx;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(ISX_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ---- host threads that complete the obs rows of a host-buffer step (isx_host_expand.cpp) while later env ranges are still
// being simulated / copied.  The batch is cut into copy CHUNKS; behind the device->host copies of a chunk the copy stream
// writes the step's sequence number into the chunk's flag word in pinned host memory (a 4-byte device->host copy: same
// stream, so it lands after the data).  Worker w owns a fixed slice of every chunk and polls the flag — no driver callback
// thread, no queues, no locks on the data path.
struct ExpandPool {
    struct Range { size_t a0, a1; };                 // agent range of one copy chunk
    std::vector<std::thread> threads;
    std::vector<Range> ranges;
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    uint32_t seq = 0;                                // step sequence number (guarded by mu), bumped by the caller to start a step
    volatile uint32_t* flag = nullptr;               // pinned [chunks]: flag[c] == seq once chunk c's compact records are in host memory
    int done = 0;
    bool quit = false;
    // the job of the current step
    const float* rec = nullptr; const uint8_t* hits = nullptr; float* dst = nullptr; int R = 0;

    void worker(int w, int T) {
        uint32_t seen = 0, cp_seen = 0;
        for (;;) {
            bool do_copy = false;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_start.wait(lk, [&] { return quit || seq != seen || cp_seq != cp_seen; });
                if (quit) return;
                if (cp_seq != cp_seen) { cp_seen = cp_seq; do_copy = true; } else seen = seq;
            }
            if (do_copy) {
                const size_t per = ((cp_bytes / (size_t)T) + 63) & ~(size_t)63, b0 = per * (size_t)w, b1 = b0 + per < cp_bytes ? b0 + per : cp_bytes;
                if (b0 < cp_bytes) std::memcpy(cp_to + b0, cp_from + b0, (w == T - 1 ? cp_bytes : b1) - b0);
                std::lock_guard<std::mutex> lk(mu);
                if (++done == T) cv_done.notify_one();
                continue;
            }
            for (size_t c = 0; c < ranges.size(); ++c) {
                int spins = 0;
                while (flag[c] != seen) {              // brief pauses, then give the core away: the caller's thread must keep running
                    if (++spins < 256) __builtin_ia32_pause(); else { std::this_thread::yield(); spins = 0; }
                }
                std::atomic_thread_fence(std::memory_order_acquire);
                // slices are cut at multiples of 8 rows so that each starts 32-byte aligned (non-temporal store path)
                const size_t n = ranges[c].a1 - ranges[c].a0, blocks = (n + 7) / 8;
                const size_t b0 = blocks * (size_t)w / (size_t)T, b1 = blocks * (size_t)(w + 1) / (size_t)T;
                const size_t r0 = ranges[c].a0 + b0 * 8, r1 = ranges[c].a0 + (b1 * 8 < n ? b1 * 8 : n);
                if (r1 > r0) expand_obs_rows(rec + r0 * 32, hits + r0 * (size_t)R, R, dst + r0 * ISX_OBS_DIM, r1 - r0);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (++done == T) cv_done.notify_one();
            }
        }
    }
    void start(int T) {
        for (int w = 0; w < T; ++w) threads.emplace_back([this, w, T] { worker(w, T); });
    }
    uint32_t begin_step(const float* rec_, const uint8_t* hits_, float* dst_, int R_) {
        std::lock_guard<std::mutex> lk(mu);
        rec = rec_; hits = hits_; dst = dst_; R = R_; done = 0; ++seq;
        if (seq == 0) ++seq;
        cv_start.notify_all();
        return seq;
    }
    // a plain parallel memcpy on the same threads (stages the caller's actions into the pinned upload buffer)
    void copy(void* to, const void* from, size_t bytes) {
        const size_t T = threads.size();
        if (T < 2 || bytes < (1u << 20)) { std::memcpy(to, from, bytes); return; }
        {
            std::lock_guard<std::mutex> lk(mu);
            cp_to = static_cast<char*>(to); cp_from = static_cast<const char*>(from); cp_bytes = bytes; done = 0; ++cp_seq;
            cv_start.notify_all();
        }
        wait_step();
    }
    char* cp_to = nullptr; const char* cp_from = nullptr; size_t cp_bytes = 0; uint32_t cp_seq = 0;
    void release_all() { for (size_t c = 0; c < ranges.size(); ++c) flag[c] = seq; }   // error path: let the workers finish
    void wait_step() {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return done == (int)threads.size(); });
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; cv_start.notify_all(); }
        for (auto& t : threads) t.join();
        threads.clear();
    }
};

struct isx_handle {
    // One homogeneous slice of the batch (isx_create_groups): its own settings / tables over an env range of the buffers.
    struct Group {
        Dev d{};
        isx_config cfg{};
        std::vector<RouteHost> routes;
        int first = 0;                                   // first env of the group in the whole batch
        float last_dt = -1.0f, last_prob = 0.0f;
    };
    struct Piece { int group, e0, cnt; };                // env range (relative to its group) of the host-step pipeline
    Dev d{};                                             // the whole batch (buffers; group 0's settings)
    std::vector<Group> groups;
    std::vector<Piece> pipe;
    int device = 0;
    int lidar_grid = 0;
    std::vector<void*> allocs;
    bool guard = false;                                  // ISX_GUARD=1: red zones around every device buffer
    std::vector<std::pair<unsigned char*, size_t>> guarded;   // (user pointer, user bytes) of every guarded buffer
    // pinned staging for isx_step_host
    float* h_actions = nullptr; float* d_actions = nullptr;
    float* h_obs = nullptr; float* h_reward = nullptr;
    float* h_rec = nullptr; uint8_t* h_hitc = nullptr;   // pinned landing zone of the compact obs records (isx_host_expand.cpp)
    ExpandPool* pool = nullptr;                          // null: rows are completed inline after the step (small batches)
    struct Chunk { int piece; size_t a0, a1; };          // copy chunk: agent range [a0, a1) of the whole batch, inside pipeline piece `piece`
    std::vector<Chunk> chunks;
    uint32_t* h_seq = nullptr;                           // pinned: [0] = this step's sequence number, [1 + c] = flag of chunk c
    uint32_t* d_seq = nullptr;
    float* expand_dst = nullptr;                         // where this step's rows go (h_obs or the caller's buffer)
    int host_threads = 0;
    uint8_t *h_done = nullptr, *h_status = nullptr, *h_term = nullptr, *h_trunc = nullptr;   // views into h_small
    uint8_t *d_small = nullptr, *h_small = nullptr;      // reward | done | status | terminated | truncated, one block each side
    size_t small_off[7] = {0, 0, 0, 0, 0, 0, 0}, small_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> ev_shard;                   // one per pipeline piece
    cudaStream_t pipe_stream = nullptr;                  // the captured host step replays here
    cudaEvent_t ev_pipe_in = nullptr;
    cudaGraphExec_t pipe_exec = nullptr;
    float pipe_dt = 0.0f;
    bool use_graph = true;
    cudaEvent_t ev_copy_done = nullptr;
};
// A view of envs [e0, e0+cnt) of the same buffers: every per-env array is contiguous per env, so a shard is the same
// struct with advanced pointers; RNG stays keyed by the global env id through env_base.
static Dev shard_of(const Dev& d, int e0, int cnt, int shard_idx) {
    Dev s = d;
    const size_t oE = (size_t)e0, oEN = oE * d.N, oEM = oE * d.M, oEC = oE * (size_t)(d.N + d.M);
    s.E = cnt; s.env_base = d.env_base + e0;
    s.ex += oEN; s.ey += oEN; s.ev += oEN; s.eh += oEN; s.esteer += oEN; s.eacc += oEN; s.epd += oEN; s.epa0 += oEN; s.epa1 += oEN;
    s.epidx += oEN; s.ealive += oEN;
    s.nx += oEM; s.ny += oEM; s.nv += oEM; s.nh += oEM; s.nsteer += oEM; s.npidx += oEM; s.nroute += oEM; s.nuid += oEM;
    s.ncount += oE; s.next_uid += oE; s.step_count += oE; s.tick += oE;
    s.agent_rec = static_cast<char*>(d.agent_rec) + oEN * 16; s.car_rect = static_cast<char*>(d.car_rect) + oEC * sizeof(PixRect);
    s.cand += oEN * (size_t)(d.N + d.M); s.cand_n += oEN; s.ray_counter = d.ray_counter + shard_idx;
    if (d.order) { s.order = d.order + oE; s.order_cnt = d.order_cnt + 8 * shard_idx; }
    s.obs_c += oEN * 32; s.hit_c += oEN * (size_t)d.R;
    s.obs += oEN * ISX_OBS_DIM; s.reward += oEN; s.done += oEN; s.status += oEN;
    s.terminated += oE; s.truncated += oE; s.agents_alive += oE; s.lidar_hit += oEN * ISX_MAX_RAYS;
    s.events += oE; s.env_stats += oE * STAT_SLOTS;
    if (d.trace) s.trace = d.trace + oE * 16;
    return s;
}

static isx_handle::Group& group_of(isx_handle* h, int env) {
    for (auto& g : h->groups) if (env < g.first + g.d.E) return g;
    return h->groups.back();
}

// Every device buffer of a handle.  With ISX_GUARD=1 in the environment at isx_create (debug aid; compute-sanitizer is not
// available everywhere) each buffer sits between two 4 KB red zones filled with a canary byte; isx_debug_check_guards
// reports how many red zones were written to — an out-of-bounds store of any kernel shows up there.
constexpr size_t GUARD_BYTES = 4096;
constexpr int GUARD_CANARY = 0xA5;
template <class T>
static int dev_alloc(isx_handle* h, T** p, size_t n, bool zero = true) {
    void* q = nullptr;
    const size_t bytes = sizeof(T) * (n ? n : 1);
    if (!h->guard) {
        CK(cudaMalloc(&q, bytes));
        if (zero) CK(cudaMemset(q, 0, bytes));
        h->allocs.push_back(q);
        *p = static_cast<T*>(q);
        return 0;
    }
    const size_t padded = (bytes + 255) & ~(size_t)255;              // keep the user pointer 256-byte aligned
    CK(cudaMalloc(&q, padded + 2 * GUARD_BYTES));
    CK(cudaMemset(q, GUARD_CANARY, padded + 2 * GUARD_BYTES));
    unsigned char* user = static_cast<unsigned char*>(q) + GUARD_BYTES;
    if (zero) CK(cudaMemset(user, 0, bytes));
    h->allocs.push_back(q);
    h->guarded.push_back({user, bytes});
    *p = reinterpret_cast<T*>(user);
    return 0;
}

// ---- device snapshots (get_state / set_state of the reference, IntersectionEnv.cpp:394-416, for all envs at once)
struct SnapEntry { unsigned char* live; unsigned char* saved; unsigned bytes_per_env; };
struct isx_snapshot {
    isx_handle* owner = nullptr;
    unsigned char* store = nullptr;       // one allocation holding every saved array
    void* dev_table = nullptr;            // SnapArray[n] for the masked-restore kernel
    std::vector<SnapEntry> entries;
};
static std::vector<std::pair<void*, unsigned>> snapshot_arrays(const Dev& d) {
    const unsigned N = (unsigned)d.N, M = (unsigned)d.M;
    std::vector<std::pair<void*, unsigned>> v;
    for (float* p : {d.ex, d.ey, d.ev, d.eh, d.esteer, d.eacc, d.epd, d.epa0, d.epa1}) v.push_back({p, 4u * N});
    v.push_back({d.epidx, 4u * N}); v.push_back({d.ealive, N});
    for (float* p : {d.nx, d.ny, d.nv, d.nh, d.nsteer}) v.push_back({p, 4u * M});
    v.push_back({d.npidx, 4u * M}); v.push_back({d.nroute, 4u * M}); v.push_back({d.nuid, 4u * M});
    v.push_back({d.ncount, 4u}); v.push_back({d.next_uid, 4u}); v.push_back({d.step_count, 4u}); v.push_back({d.tick, 4u});
    // outputs too, so that the observation after a restore is the observation at save time (the reference instead
    // resets its lidars to default 72-beam ones in set_state, IntersectionEnv.cpp:411-415 — a quirk not reproduced)
    v.push_back({d.obs, 4u * N * ISX_OBS_DIM}); v.push_back({d.reward, 4u * N}); v.push_back({d.done, N}); v.push_back({d.status, N});
    v.push_back({d.terminated, 1u}); v.push_back({d.truncated, 1u}); v.push_back({d.agents_alive, 4u});
    v.push_back({d.lidar_hit, N * ISX_MAX_RAYS}); v.push_back({d.events, (unsigned)sizeof(isx_traffic_events)});
    v.push_back({d.car_rect, (unsigned)sizeof(PixRect) * (N + M)});
    return v;
}

template <class T>
static cudaError_t pull(std::vector<T>& v, const T* dev, size_t off, size_t n) {
    v.resize(n ? n : 1);
    return cudaMemcpy(v.data(), dev + off, sizeof(T) * n, cudaMemcpyDeviceToHost);
}
template <class T>
static cudaError_t push(const std::vector<T>& v, T* dev, size_t off, size_t n) {
    return cudaMemcpy(dev + off, v.data(), sizeof(T) * n, cudaMemcpyHostToDevice);
}

extern "C" {

int isx_rollout_timed4(isx_handle* h, int32_t steps, float dt, void* stream, float* ms4);

const char* isx_last_error(void) { return g_err.c_str(); }
int isx_abi_version(void) { return ISX_ABI_VERSION; }

int isx_route(int32_t lanes, const char* start, const char* end, float* path_xy, int32_t* intent, float* sx, float* sy, float* sh) {
    if (!start || !end) return fail(ISX_E_ARG, "null lane id");
    RouteHost r;
    const int rc = build_route(lanes, start, end, &r);
    if (rc == -1) return fail(ISX_E_ROUTE_START, "unknown start lane id '%s'", start);
    if (rc == -2) return fail(ISX_E_ROUTE_END, "unknown end lane id '%s'", end);
    if (path_xy) for (int i = 0; i < PATH_LEN; ++i) { path_xy[2 * i] = r.path[i].x; path_xy[2 * i + 1] = r.path[i].y; }
    if (intent) *intent = r.intent;
    if (sx) *sx = r.spawn_x;
    if (sy) *sy = r.spawn_y;
    if (sh) *sh = r.spawn_h;
    return PATH_LEN;
}

// Validates one config and fills the configuration half of a Dev (no pointers yet).
static int config_to_dev(const isx_config* cfg, Dev& d) {
    if (cfg->abi_version != ISX_ABI_VERSION) return fail(ISX_E_ARG, "abi_version %d != %d", cfg->abi_version, ISX_ABI_VERSION);
    if (cfg->num_envs < 1) return fail(ISX_E_ARG, "num_envs must be >= 1");
    if (cfg->num_agents < 1 || cfg->num_agents > ISX_MAX_AGENTS) return fail(ISX_E_ARG, "num_agents must be in [1,%d]", ISX_MAX_AGENTS);
    if (cfg->num_lanes < 1 || cfg->num_lanes > 4) return fail(ISX_E_ARG, "num_lanes must be in [1,4]");
    if (cfg->lidar_rays < 1 || cfg->lidar_rays > ISX_MAX_RAYS) return fail(ISX_E_ARG, "lidar_rays must be in [1,%d]", ISX_MAX_RAYS);
    if (cfg->npc_capacity < 0 || cfg->npc_capacity > ISX_MAX_NPC) return fail(ISX_E_ARG, "npc_capacity must be in [0,%d]", ISX_MAX_NPC);
    if (cfg->auto_reset < 0 || cfg->auto_reset > 2) return fail(ISX_E_ARG, "auto_reset must be 0, 1 or 2");
    if (cfg->num_traffic_routes < 0 || cfg->num_traffic_routes > ISX_MAX_ROUTES) return fail(ISX_E_ARG, "num_traffic_routes must be in [0,%d]", ISX_MAX_ROUTES);
    if (!cfg->ego_start || !cfg->ego_end) return fail(ISX_E_ARG, "ego routes missing");
    d.E = cfg->num_envs; d.N = cfg->num_agents; d.M = cfg->traffic_flow ? (cfg->npc_capacity > 0 ? cfg->npc_capacity : 16) : 1;
    d.R = cfg->lidar_rays; d.lanes = cfg->num_lanes;
    d.use_team = cfg->use_team_reward != 0; d.respawn = cfg->respawn_enabled != 0; d.max_steps = cfg->max_steps;
    d.traffic = cfg->traffic_flow != 0; d.T = cfg->num_traffic_routes; d.auto_reset = cfg->auto_reset;
    d.rc = RewardCfg{cfg->reward[0], cfg->reward[1], cfg->reward[2], cfg->reward[3], cfg->reward[4], cfg->reward[5], cfg->reward[6], cfg->reward[7]};
    d.max_progress = hypotf_((float)WIDTH, (float)HEIGHT);
    d.seed = cfg->seed; d.env_base = cfg->env_id_base;
    if (d.traffic && d.T > 0 && (!cfg->traffic_start || !cfg->traffic_end)) return fail(ISX_E_ARG, "traffic routes missing");
    d.traffic_lanes = 0;
    if (const char* tl = getenv("ISX_TRAFFIC_LANES")) { const int v = atoi(tl); if (v == 8 || v == 16 || v == 32) d.traffic_lanes = v; }
    return ISX_OK;
}

int isx_create(const isx_config* cfg, isx_handle** out) { return isx_create_groups(cfg, 1, out); }

// A heterogeneous batch: group g owns the envs [first_g, first_g + cfgs[g].num_envs) of ONE set of buffers and is stepped
// with its own routes / lanes / traffic / reward / episode settings (its own constant tables and kernel launches).
// Every group keeps its own seed and env_id_base, so a group evolves bit-for-bit like a stand-alone batch created from
// the same config.  The groups must agree on what shapes the shared buffers: device, num_agents, lidar_rays, and — for
// the groups with traffic — npc_capacity.
int isx_create_groups(const isx_config* cfgs, int32_t n_groups, isx_handle** out) {
    if (!cfgs || !out) return fail(ISX_E_ARG, "null argument");
    *out = nullptr;
    if (n_groups < 1 || n_groups > ISX_MAX_GROUPS) return fail(ISX_E_ARG, "n_groups must be in [1,%d]", ISX_MAX_GROUPS);
    std::vector<Dev> gd((size_t)n_groups);
    long long total_envs = 0;
    int M_all = 1;
    for (int g = 0; g < n_groups; ++g) {
        const int rc = config_to_dev(&cfgs[g], gd[(size_t)g]);
        if (rc) return rc;
        const isx_config& c = cfgs[g];
        if (c.device != cfgs[0].device || c.num_agents != cfgs[0].num_agents || c.lidar_rays != cfgs[0].lidar_rays)
            return fail(ISX_E_ARG, "group %d: device, num_agents and lidar_rays must be the same in every group", g);
        if (gd[(size_t)g].traffic) {
            if (M_all != 1 && gd[(size_t)g].M != M_all) return fail(ISX_E_ARG, "group %d: npc_capacity must be the same in every group with traffic", g);
            M_all = gd[(size_t)g].M;
        }
        total_envs += c.num_envs;
    }
    if (total_envs * cfgs[0].num_agents * ISX_MAX_RAYS >= (1ll << 31)) return fail(ISX_E_ARG, "num_envs * num_agents too large (beam index must fit 31 bits)");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ISX_E_CUDA, "no CUDA device: this library has no CPU path");
    if (cfgs[0].device < 0 || cfgs[0].device >= ndev) return fail(ISX_E_ARG, "device %d out of range (%d devices)", cfgs[0].device, ndev);
    CK(cudaSetDevice(cfgs[0].device));

    isx_handle* h = new isx_handle();
    { const char* g_ = getenv("ISX_GUARD"); h->guard = g_ && g_[0] == '1'; }
    h->device = cfgs[0].device;
    Dev& d = h->d;                      // the whole batch: group 0's settings, every env, the shared stride M
    d = gd[0];
    d.E = (int)total_envs; d.M = M_all;

#define ALLOC(ptr, n)                                     \
    do {                                                  \
        int rc_ = dev_alloc(h, &(ptr), (size_t)(n));      \
        if (rc_) { isx_destroy(h); return rc_; }          \
    } while (0)
#define UPLOAD(dst, src, bytes)                                                                  \
    do {                                                                                         \
        cudaError_t e_ = cudaMemcpy((void*)(dst), (src), (bytes), cudaMemcpyHostToDevice);       \
        if (e_ != cudaSuccess) { isx_destroy(h); return fail(ISX_E_CUDA, "upload failed: %s", cudaGetErrorString(e_)); } \
    } while (0)

    // ---- per-group constant tables: route LUT (ego slots then traffic routes), folded road map of its lane count
    h->groups.resize((size_t)n_groups);
    float* d_rel = nullptr;
    {
        std::vector<float> rel((size_t)ISX_MAX_RAYS, 0.0f);
        lidar_rel_angles(d.R, rel.data());
        ALLOC(d_rel, rel.size()); UPLOAD(d_rel, rel.data(), sizeof(float) * rel.size());
    }
    for (int g = 0; g < n_groups; ++g) {
        isx_handle::Group& grp = h->groups[(size_t)g];
        const isx_config* cfg = &cfgs[g];
        Dev& q = gd[(size_t)g];
        grp.cfg = *cfg;
        const int nroutes = q.N + q.T;
        grp.routes.resize((size_t)nroutes);
        for (int i = 0; i < nroutes; ++i) {
            const char* s = i < q.N ? cfg->ego_start[i] : cfg->traffic_start[i - q.N];
            const char* e = i < q.N ? cfg->ego_end[i] : cfg->traffic_end[i - q.N];
            if (!s || !e) { isx_destroy(h); return fail(ISX_E_ARG, "null lane id in route %d", i); }
            const int rc = build_route(q.lanes, s, e, &grp.routes[(size_t)i]);
            if (rc == -1) { isx_destroy(h); return fail(ISX_E_ROUTE_START, "unknown start lane id '%s' (route %d)", s, i); }
            if (rc == -2) { isx_destroy(h); return fail(ISX_E_ROUTE_END, "unknown end lane id '%s' (route %d)", e, i); }
        }
        std::vector<F2> paths((size_t)nroutes * PATH_LEN);
        std::vector<RouteMeta> meta((size_t)nroutes);
        for (int i = 0; i < nroutes; ++i) {
            const RouteHost& r = grp.routes[(size_t)i];
            std::memcpy(&paths[(size_t)i * PATH_LEN], r.path, sizeof(F2) * PATH_LEN);
            meta[(size_t)i] = RouteMeta{r.spawn_x, r.spawn_y, r.spawn_h, r.intent, r.path[PATH_LEN - 1], r.path[PATH_LEN - 2]};
        }
        RoadTables rt;
        if (!build_road_tables(q.lanes, &rt)) { isx_destroy(h); return fail(ISX_E_STATE, "road map is not mirror-symmetric"); }
        F2* d_paths; RouteMeta* d_meta; uint32_t* d_bits; uint8_t* d_skip;
        ALLOC(d_paths, paths.size()); UPLOAD(d_paths, paths.data(), sizeof(F2) * paths.size());
        std::vector<float> far2(paths.size());
        for (size_t r0 = 0; r0 + PATH_LEN <= paths.size(); r0 += PATH_LEN) path_far_table(paths.data() + r0, far2.data() + r0);
        float* d_far2;
        ALLOC(d_far2, far2.size()); UPLOAD(d_far2, far2.data(), sizeof(float) * far2.size());
        q.route_far2 = d_far2;
        ALLOC(d_meta, meta.size()); UPLOAD(d_meta, meta.data(), sizeof(RouteMeta) * meta.size());
        rt.bits.resize(road_bits_bytes() / 4, 0u);        // padded to 16 B multiples for the kernel's vector copy
        rt.skip.resize(road_skip_bytes(), 0);
        ALLOC(d_bits, rt.bits.size()); UPLOAD(d_bits, rt.bits.data(), sizeof(uint32_t) * rt.bits.size());
        ALLOC(d_skip, rt.skip.size()); UPLOAD(d_skip, rt.skip.data(), rt.skip.size());
        q.route_path = d_paths; q.route_meta = d_meta; q.road_bits = d_bits; q.road_skip = d_skip; q.rel_angle = d_rel;
        q.box_lo = rt.box_lo; q.box_hi = rt.box_hi; q.ana = rt.ana;
    }
    d.route_path = gd[0].route_path; d.route_far2 = gd[0].route_far2; d.route_meta = gd[0].route_meta; d.road_bits = gd[0].road_bits; d.road_skip = gd[0].road_skip;
    d.rel_angle = d_rel; d.box_lo = gd[0].box_lo; d.box_hi = gd[0].box_hi; d.ana = gd[0].ana;

    // ---- state, scratch and output buffers, shared by all groups
    const size_t EN = (size_t)d.E * d.N, EM = (size_t)d.E * d.M, E = (size_t)d.E;
    ALLOC(d.ex, EN); ALLOC(d.ey, EN); ALLOC(d.ev, EN); ALLOC(d.eh, EN); ALLOC(d.esteer, EN); ALLOC(d.eacc, EN);
    ALLOC(d.epd, EN); ALLOC(d.epa0, EN); ALLOC(d.epa1, EN); ALLOC(d.epidx, EN); ALLOC(d.ealive, EN);
    ALLOC(d.nx, EM); ALLOC(d.ny, EM); ALLOC(d.nv, EM); ALLOC(d.nh, EM); ALLOC(d.nsteer, EM);
    ALLOC(d.npidx, EM); ALLOC(d.nroute, EM); ALLOC(d.nuid, EM);
    ALLOC(d.ncount, E); ALLOC(d.next_uid, E); ALLOC(d.tick, E);
    ALLOC(d.obs, EN * ISX_OBS_DIM);
    ALLOC(d.obs_c, EN * 32); ALLOC(d.hit_c, EN * ISX_MAX_RAYS);     // hit_c is used densely at stride R (<= 96)
    {
        // reward | done | status | terminated | truncated | agents_alive | step live in ONE block (sub-arrays 256 B
        // aligned), mirrored by one pinned host block, so the host-buffer step brings all of them back with a single copy
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        h->small_off[0] = 0;
        h->small_off[1] = up(sizeof(float) * EN);
        h->small_off[2] = h->small_off[1] + up(EN);
        h->small_off[3] = h->small_off[2] + up(EN);
        h->small_off[4] = h->small_off[3] + up(E);
        h->small_off[5] = h->small_off[4] + up(E);
        h->small_off[6] = h->small_off[5] + up(sizeof(int) * E);
        h->small_bytes = h->small_off[6] + up(sizeof(int) * E);
        ALLOC(h->d_small, h->small_bytes);
        d.agents_alive = reinterpret_cast<int*>(h->d_small + h->small_off[5]);
        d.step_count = reinterpret_cast<int*>(h->d_small + h->small_off[6]);
        d.reward = reinterpret_cast<float*>(h->d_small + h->small_off[0]);
        d.done = h->d_small + h->small_off[1]; d.status = h->d_small + h->small_off[2];
        d.terminated = h->d_small + h->small_off[3]; d.truncated = h->d_small + h->small_off[4];
    }
    ALLOC(d.lidar_hit, EN * ISX_MAX_RAYS); ALLOC(d.events, E);
    ALLOC(d.env_stats, E * STAT_SLOTS); ALLOC(d.stats, 16);
    ALLOC(h->d_actions, EN * 2);
    if (getenv("ISX_TRACE")) ALLOC(d.trace, E * 16); else d.trace = nullptr;
    {
        float4* rec; PixRect* rc;
        ALLOC(rec, EN); ALLOC(rc, E * (size_t)(d.N + d.M)); ALLOC(d.cand, EN * (size_t)(d.N + d.M)); ALLOC(d.cand_n, EN);
        ALLOC(d.ray_counter, ISX_MAX_GROUPS + 8);
        // env lists by NPC count for k_traffic (ISX_NO_ORDER=1: envs in index order; tuning / bisecting aid)
        d.order = nullptr; d.order_cnt = nullptr; d.order_stride = (int)E;
        if (getenv("ISX_NO_ORDER") == nullptr) { ALLOC(d.order, 5 * E); ALLOC(d.order_cnt, (ISX_MAX_GROUPS + 8) * 8); }
        d.agent_rec = rec; d.car_rect = rc;
    }
#undef ALLOC
#undef UPLOAD

    // ---- group views: the env range of the shared buffers + the group's own settings and tables
    {
        int first = 0;
        for (int g = 0; g < n_groups; ++g) {
            isx_handle::Group& grp = h->groups[(size_t)g];
            const Dev& q = gd[(size_t)g];
            Dev v = shard_of(d, first, q.E, g);
            v.lanes = q.lanes; v.use_team = q.use_team; v.respawn = q.respawn; v.max_steps = q.max_steps;
            v.traffic = q.traffic; v.T = q.T; v.auto_reset = q.auto_reset; v.rc = q.rc; v.seed = q.seed; v.env_base = q.env_base;
            v.route_path = q.route_path; v.route_far2 = q.route_far2; v.route_meta = q.route_meta; v.road_bits = q.road_bits; v.road_skip = q.road_skip;
            v.box_lo = q.box_lo; v.box_hi = q.box_hi; v.ana = q.ana;
            grp.d = v; grp.first = first;
            first += q.E;
        }
        if (n_groups == 1) h->d.env_base = gd[0].env_base;
    }

    // ---- contraction canary: the build must not fuse a*b+c (isx_math.cuh); fail loudly if it does
    {
        float* d_out = reinterpret_cast<float*>(d.stats);
        const float a = 1.0f + 0x1p-12f, b = 1.0f + 0x1p-12f, c = -1.0f;   // a*b rounds; fma keeps the 2^-24 term
        cudaError_t e = launch_canary(a, b, c, d_out, 0);
        float res[2] = {0, 0};
        if (e == cudaSuccess) e = cudaMemcpy(res, d_out, sizeof res, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { isx_destroy(h); return fail(ISX_E_CUDA, "self-test launch failed: %s", cudaGetErrorString(e)); }
        const volatile float prod = a * b;
        const float expect = prod + c;
        if (std::memcmp(&res[0], &expect, 4) != 0) { isx_destroy(h); return fail(ISX_E_STATE, "library was built with FMA contraction on; rebuild with -fmad=false"); }
        cudaMemset(d.stats, 0, 16 * sizeof(unsigned long long));
    }
    {
        cudaError_t e = lidar_set_smem_attr(d);
        int per_sm = 0, sms = 0;
        if (e == cudaSuccess) e = lidar_occupancy(d, &per_sm);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        if (e != cudaSuccess || per_sm < 1) { isx_destroy(h); return fail(ISX_E_CUDA, "lidar kernel cannot be scheduled: %s", cudaGetErrorString(e)); }
        if (const char* cap = getenv("ISX_LIDAR_CTAS_PER_SM")) {   // tuning aid: leave room for other kernels to co-reside
            const int c = atoi(cap);
            if (c >= 1 && c < per_sm) per_sm = c;
        }
        h->lidar_grid = per_sm * sms;
    }
    // pinned staging
    if (cudaMallocHost((void**)&h->h_actions, sizeof(float) * EN * 2) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_obs, sizeof(float) * EN * ISX_OBS_DIM) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_small, h->small_bytes) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_rec, sizeof(float) * EN * 32) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_hitc, EN * ISX_MAX_RAYS) != cudaSuccess) {
        isx_destroy(h);
        return fail(ISX_E_CUDA, "pinned host allocation failed");
    }
    std::memset(h->h_obs, 0, sizeof(float) * EN * ISX_OBS_DIM);
    h->h_reward = reinterpret_cast<float*>(h->h_small + h->small_off[0]);
    h->h_done = h->h_small + h->small_off[1]; h->h_status = h->h_small + h->small_off[2];
    h->h_term = h->h_small + h->small_off[3]; h->h_trunc = h->h_small + h->small_off[4];
    {
        // Pipeline plan of the host-buffer step: env ranges whose kernels run while the copy engine drains the previous
        // range's obs rows.  Measured on B200 + PCIe gen5 (tools/e2e_probe.py): the D2H of obs is 589 us of a 729 us step
        // at 8192x8; each range costs ~55 us of fixed kernel latency, so few equal ranges beat many or geometric ones.
        // One group: 4 equal ranges, 8 from 32768 envs up where the fixed cost per range no longer shows
        // (ISX_PIPE_PLAN="w0,w1,..." (<= 16 weights) overrides, for tuning).  Several groups:
        // every group is one range.
        if (n_groups == 1) {
            // 65,536 envs (compact transport, 12 host threads): 1,2,3,4,6 -> 3.33 ms; 4 equal 3.59; 8 equal 3.57; 1,2,2,3 3.44
            // below 16,384 agents every kernel sits at its latency floor (one wave): cutting the batch only repeats that floor per range
            int w[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, nw = d.E >= 32768 ? 5 : (d.E >= 1024 && EN >= 16384) ? 4 : 1;
            if (d.E >= 32768) { w[0] = 1; w[1] = 2; w[2] = 3; w[3] = 4; w[4] = 6; }   // a small first range starts the copy engine early
            if (const char* plan = getenv("ISX_PIPE_PLAN")) {
                int k = 0;
                for (const char* c = plan; *c && k < 16;) {
                    char* endp = nullptr;
                    const long v = std::strtol(c, &endp, 10);
                    if (endp == c) break;
                    if (v > 0) w[k++] = (int)v;
                    c = (*endp == ',') ? endp + 1 : endp;
                }
                if (k > 0) nw = k;
            }
            long long tot = 0, acc = 0;
            for (int i = 0; i < nw; ++i) tot += w[i];
            for (int i = 0; i < nw; ++i) {
                const int e0 = (int)((long long)d.E * acc / tot);
                acc += w[i];
                const int e1 = (int)((long long)d.E * acc / tot);
                if (e1 > e0) h->pipe.push_back(isx_handle::Piece{0, e0, e1 - e0});
            }
        } else {
            for (int g = 0; g < n_groups; ++g) h->pipe.push_back(isx_handle::Piece{g, 0, h->groups[(size_t)g].d.E});
        }
        h->ev_shard.resize(h->pipe.size(), nullptr);
        // Host threads completing the obs rows: as many as this process may run on (the caller binds ranks to disjoint CPU
        // sets), ISX_HOST_THREADS overrides; small batches finish their rows inline, without callbacks or thread wake-ups.
        int T = 0;
        cpu_set_t cs;
        if (sched_getaffinity(0, sizeof cs, &cs) == 0) T = CPU_COUNT(&cs);
        if (T < 1) T = (int)std::thread::hardware_concurrency();
        T = T >= 4 ? (T * 3) / 4 : T;      // leave cores to the caller's thread and the driver (16 CPUs: 12 -> 3.42 ms, 14 -> 3.76, 8 -> 3.69)
        if (T > 32) T = 32;
        if (const char* ht = getenv("ISX_HOST_THREADS")) { const int v = atoi(ht); if (v >= 0 && v <= 256) T = v; }
        if (EN < 16384) T = 0;             // small batches ship the rows themselves (enqueue_pinned_step): no thread wake-ups
        // copy chunks: every pipeline piece is drained in `sub` chunks so that the host threads start on a piece while its
        // tail is still crossing PCIe (ISX_PIPE_CHUNKS overrides; about 1M compact-record bytes per chunk at the least)
        {
            int sub = 4;
            if (const char* pcs = getenv("ISX_PIPE_CHUNKS")) { const int v = atoi(pcs); if (v >= 1 && v <= 16) sub = v; }
            for (size_t c = 0; c < h->pipe.size(); ++c) {
                const isx_handle::Piece& pc = h->pipe[c];
                const size_t a0 = (size_t)(h->groups[(size_t)pc.group].first + pc.e0) * d.N, n = (size_t)pc.cnt * d.N;
                int k = sub;
                while (k > 1 && n / (size_t)k < 4096) --k;
                for (int i = 0; i < k; ++i) {
                    const size_t b0 = (n * (size_t)i / (size_t)k) & ~(size_t)7, b1 = i + 1 == k ? n : ((n * (size_t)(i + 1) / (size_t)k) & ~(size_t)7);
                    if (b1 > b0) h->chunks.push_back(isx_handle::Chunk{(int)c, a0 + b0, a0 + b1});
                }
            }
            if (h->chunks.size() > 256) T = 0;
        }
        h->host_threads = T;
        if (cudaMallocHost((void**)&h->h_seq, sizeof(uint32_t) * (h->chunks.size() + 1)) != cudaSuccess) { isx_destroy(h); return fail(ISX_E_CUDA, "pinned host allocation failed"); }
        std::memset(h->h_seq, 0, sizeof(uint32_t) * (h->chunks.size() + 1));
        { int rc_ = dev_alloc(h, &h->d_seq, 1); if (rc_) { isx_destroy(h); return rc_; } }
        if (T > 0) {
            h->pool = new ExpandPool();
            h->pool->flag = h->h_seq + 1;
            for (const auto& ch : h->chunks) h->pool->ranges.push_back(ExpandPool::Range{ch.a0, ch.a1});
            h->pool->start(T);
        }
    }
    {
        cudaError_t e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
        for (size_t i = 0; i < h->ev_shard.size() && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_shard[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_copy_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->pipe_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pipe_in, cudaEventDisableTiming);
        h->use_graph = getenv("ISX_NO_GRAPH") == nullptr;
        if (e != cudaSuccess) { isx_destroy(h); return fail(ISX_E_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e)); }
    }
    *out = h;
    const int rc = isx_reset(h, nullptr, nullptr);
    if (rc) { isx_destroy(h); *out = nullptr; return rc; }
    CK(cudaDeviceSynchronize());
    return ISX_OK;
}

int isx_destroy(isx_handle* h) {
    if (!h) return ISX_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->pool) { h->pool->stop(); delete h->pool; h->pool = nullptr; }
    for (void* p : h->allocs) cudaFree(p);
    if (h->pipe_exec) cudaGraphExecDestroy(h->pipe_exec);
    if (h->pipe_stream) cudaStreamDestroy(h->pipe_stream);
    if (h->ev_pipe_in) cudaEventDestroy(h->ev_pipe_in);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (cudaEvent_t e : h->ev_shard) if (e) cudaEventDestroy(e);
    if (h->ev_copy_done) cudaEventDestroy(h->ev_copy_done);
    if (h->h_actions) cudaFreeHost(h->h_actions);
    if (h->h_obs) cudaFreeHost(h->h_obs);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->h_rec) cudaFreeHost(h->h_rec);
    if (h->h_hitc) cudaFreeHost(h->h_hitc);
    if (h->h_seq) cudaFreeHost(h->h_seq);
    delete h;
    return ISX_OK;
}

int isx_num_envs(isx_handle* h) { return h ? h->d.E : fail(ISX_E_ARG, "null handle"); }
int isx_num_groups(isx_handle* h) { return h ? (int)h->groups.size() : fail(ISX_E_ARG, "null handle"); }
int isx_group_range(isx_handle* h, int32_t group, int32_t* first_env, int32_t* num_envs) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (group < 0 || group >= (int)h->groups.size()) return fail(ISX_E_ARG, "group %d out of range", group);
    if (first_env) *first_env = h->groups[(size_t)group].first;
    if (num_envs) *num_envs = h->groups[(size_t)group].d.E;
    return ISX_OK;
}
int isx_num_agents(isx_handle* h) { return h ? h->d.N : fail(ISX_E_ARG, "null handle"); }

int isx_reset(isx_handle* h, const uint8_t* env_mask_dev, void* stream) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    for (const auto& g : h->groups) {
        CK(launch_reset(g.d, env_mask_dev ? env_mask_dev + g.first : nullptr, st));
        CK(launch_lidar_obs(g.d, /*LIDAR_FROM_HITS*/ 1, h->lidar_grid, st));
    }
    return ISX_OK;
}

int isx_observe(isx_handle* h, void* stream) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    CK(cudaSetDevice(h->device));
    for (const auto& g : h->groups) CK(launch_lidar_obs(g.d, 1, h->lidar_grid, static_cast<cudaStream_t>(stream)));
    return ISX_OK;
}

// spawn_prob = 1 - exp(-density * dt) (TrafficFlow.cpp:321-322), evaluated on the host with the host libm's
// expf — the same entry point the reference calls — once per distinct dt.
static float spawn_prob_for(isx_handle::Group& g, float dt) {
    if (dt != g.last_dt) {
        float dens = g.cfg.traffic_density;
        if (dens < 0.0f) dens = 0.0f;                       // configure_traffic clamps (IntersectionEnv.cpp:59)
        volatile float arg = -dens * dt;
        g.last_prob = 1.0f - expf(arg);
        g.last_dt = dt;
    }
    return g.last_prob;
}

int isx_step(isx_handle* h, const float* actions_dev, float dt, void* stream) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (!actions_dev) return fail(ISX_E_ARG, "actions_dev is null (use isx_rollout for on-device actions)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    for (auto& g : h->groups) {
        CK(launch_dynamics(g.d, actions_dev + (size_t)g.first * h->d.N * 2, dt, spawn_prob_for(g, dt), st));
        CK(launch_lidar_obs(g.d, 0, h->lidar_grid, st));
    }
    return ISX_OK;
}

int isx_rollout(isx_handle* h, int32_t steps, float dt, void* stream) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (steps < 0) return fail(ISX_E_ARG, "steps < 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    for (int s = 0; s < steps; ++s)
        for (auto& g : h->groups) {
            CK(launch_dynamics(g.d, nullptr, dt, spawn_prob_for(g, dt), st));
            CK(launch_lidar_obs(g.d, 0, h->lidar_grid, st));
        }
    return ISX_OK;
}

// Same loop as isx_rollout, with a CUDA-event pair around every kernel on the launching stream; returns the
// summed device time of each kernel (ms).  Synchronises.  Used by bench.py for the roofline line.
int isx_rollout_timed(isx_handle* h, int32_t steps, float dt, void* stream, float* ms_dynamics, float* ms_lidar_obs) {
    float ms[4] = {0, 0, 0, 0};
    const int rc = isx_rollout_timed4(h, steps, dt, stream, ms);
    if (ms_dynamics) *ms_dynamics = ms[0] + ms[1];
    if (ms_lidar_obs) *ms_lidar_obs = ms[2] + ms[3];
    return rc;
}

// CUDA-event pair around EVERY kernel launch: ms4 = {k_traffic, k_ego, k_features, k_lidar_obs} summed over the steps.
int isx_rollout_timed4(isx_handle* h, int32_t steps, float dt, void* stream, float* ms4) {
    if (!h || !ms4) return fail(ISX_E_ARG, "null argument");
    if (steps < 1 || steps > 4096) return fail(ISX_E_ARG, "steps must be in [1,4096]");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    if (h->groups.size() != 1) return fail(ISX_E_STATE, "per-kernel timing is for single-group batches");
    const Dev& gd0 = h->groups[0].d;
    const float prob = spawn_prob_for(h->groups[0], dt);
    std::vector<cudaEvent_t> ev((size_t)steps * 5);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    for (int s = 0; s < steps; ++s) {
        cudaEvent_t* e = &ev[(size_t)s * 5];
        CK(cudaEventRecord(e[0], st));
        CK(launch_traffic(gd0, dt, prob, st));
        CK(cudaEventRecord(e[1], st));
        CK(launch_ego(gd0, nullptr, dt, st));
        CK(cudaEventRecord(e[2], st));
        CK(launch_features(gd0, 0, st));
        CK(cudaEventRecord(e[3], st));
        CK(launch_rays(gd0, 0, h->lidar_grid, st));
        CK(cudaEventRecord(e[4], st));
    }
    CK(cudaStreamSynchronize(st));
    double acc[4] = {0, 0, 0, 0};
    for (int s = 0; s < steps; ++s)
        for (int k = 0; k < 4; ++k) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, ev[(size_t)s * 5 + k], ev[(size_t)s * 5 + k + 1]));
            acc[k] += t;
        }
    for (auto& e : ev) cudaEventDestroy(e);
    for (int k = 0; k < 4; ++k) ms4[k] = (float)acc[k];
    return ISX_OK;
}

// Enqueue one pipelined host-buffer step on `st` (+ the handle's copy stream, forked and joined through events).
static int enqueue_pinned_step(isx_handle* h, float dt, cudaStream_t st, std::vector<cudaEvent_t>* tl = nullptr) {
    const Dev& d = h->d;
    if (tl) CK(cudaEventRecord((*tl)[0], st));
    const size_t EN = (size_t)d.E * d.N;
    CK(cudaMemcpyAsync(h->d_actions, h->h_actions, sizeof(float) * EN * 2, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_seq, h->h_seq, sizeof(uint32_t), cudaMemcpyHostToDevice, st));      // this step's sequence number
    size_t next_chunk = 0;
    for (size_t c = 0; c < h->pipe.size(); ++c) {
        const isx_handle::Piece& pc = h->pipe[c];
        isx_handle::Group& grp = h->groups[(size_t)pc.group];
        const float prob = spawn_prob_for(grp, dt);
        const Dev sd = (pc.e0 == 0 && pc.cnt == grp.d.E) ? grp.d : shard_of(grp.d, pc.e0, pc.cnt, (int)c);
        const size_t aoff = (size_t)(grp.first + pc.e0) * d.N, an = (size_t)pc.cnt * d.N;
        if (tl) CK(cudaEventRecord((*tl)[1 + 4 * c], st));
        CK(launch_dynamics(sd, h->d_actions + aoff * 2, dt, prob, st));
        CK(launch_lidar_obs(sd, 0, h->lidar_grid, st));
        if (tl) CK(cudaEventRecord((*tl)[2 + 4 * c], st));
        CK(cudaEventRecord(h->ev_shard[c], st));
        CK(cudaStreamWaitEvent(h->copy_stream, h->ev_shard[c], 0));
        if (tl) CK(cudaEventRecord((*tl)[3 + 4 * c], h->copy_stream));
        // the compact obs records of the range (32 floats + R bytes per agent instead of 127 floats), chunk by chunk, each
        // followed by its flag word; host threads rebuild the rows of a chunk as soon as its flag shows this step's number
        (void)an;
        for (; next_chunk < h->chunks.size() && h->chunks[next_chunk].piece == (int)c; ++next_chunk) {
            const isx_handle::Chunk& ch = h->chunks[next_chunk];
            const size_t n = ch.a1 - ch.a0;
            if (!h->pool) {
                // small batches (no host threads): the rows themselves cross PCIe, straight into the pinned obs view — a few
                // MB take less time on the copy engine than one host thread needs to rebuild them (C2: 266 -> see DESIGN.md)
                CK(cudaMemcpyAsync(h->h_obs + ch.a0 * ISX_OBS_DIM, d.obs + ch.a0 * ISX_OBS_DIM, sizeof(float) * n * ISX_OBS_DIM, cudaMemcpyDeviceToHost, h->copy_stream));
                continue;
            }
            CK(cudaMemcpyAsync(h->h_rec + ch.a0 * 32, d.obs_c + ch.a0 * 32, sizeof(float) * n * 32, cudaMemcpyDeviceToHost, h->copy_stream));
            CK(cudaMemcpyAsync(h->h_hitc + ch.a0 * (size_t)d.R, d.hit_c + ch.a0 * (size_t)d.R, n * (size_t)d.R, cudaMemcpyDeviceToHost, h->copy_stream));
            CK(cudaMemcpyAsync(h->h_seq + 1 + next_chunk, h->d_seq, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->copy_stream));
        }
        if (tl) CK(cudaEventRecord((*tl)[4 + 4 * c], h->copy_stream));
    }
    // the stream order of copy_stream puts this after the last shard's kernels (its wait on ev_shard[last])
    CK(cudaMemcpyAsync(h->h_small, h->d_small, h->small_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaEventRecord(h->ev_copy_done, h->copy_stream));
    CK(cudaStreamWaitEvent(st, h->ev_copy_done, 0));        // join: `st` is ordered after the copies
    return ISX_OK;
}

// Host-buffer step with the device->host copy PIPELINED behind the kernels: the env range is cut into shards; the
// kernels of shard c+1 run while the copy engine drains the obs rows of shard c (obs is 127 floats per agent — the
// copy, not the simulation, bounds the end-to-end rate).  Results land in the handle's pinned staging buffers
// (isx_host_views); actions are taken from the pinned `actions` view.  Synchronous on return.
// The whole step (1 + shards copies in, 4–5 kernels and 1 copy out per shard, the fork/join events) is captured ONCE per
// dt into a CUDA graph and replayed with a single launch: issuing ~35 runtime calls per step from the host costs more
// than the first shards take to run.  ISX_NO_GRAPH=1 at isx_create keeps the plain stream path.
// One host-buffer step whose obs rows land in `dst` ([E*N][127] floats in host memory).  Synchronous.
static int host_step_into(isx_handle* h, float dt, cudaStream_t st, float* dst, std::vector<cudaEvent_t>* tl = nullptr) {
    CK(cudaSetDevice(h->device));
    if (h->pool) h->h_seq[0] = h->pool->begin_step(h->h_rec, h->h_hitc, dst, h->d.R);
    auto run = [&]() -> int {
        if (!h->use_graph || tl) {
            const int rc = enqueue_pinned_step(h, dt, st, tl);
            if (rc) return rc;
            CK(cudaEventSynchronize(h->ev_copy_done));
            CK(cudaStreamSynchronize(h->copy_stream));           // host callbacks of the last range included
            return ISX_OK;
        }
        if (!h->pipe_exec || std::memcmp(&h->pipe_dt, &dt, sizeof dt) != 0) {
            if (h->pipe_exec) { cudaGraphExecDestroy(h->pipe_exec); h->pipe_exec = nullptr; }
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(h->pipe_stream, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue_pinned_step(h, dt, h->pipe_stream);
            const cudaError_t ce = cudaStreamEndCapture(h->pipe_stream, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            CK(ce);
            const cudaError_t ie = cudaGraphInstantiate(&h->pipe_exec, g, 0);
            cudaGraphDestroy(g);
            CK(ie);
            h->pipe_dt = dt;
        }
        CK(cudaEventRecord(h->ev_pipe_in, st));                  // order the replay after whatever the caller queued on `st`
        CK(cudaStreamWaitEvent(h->pipe_stream, h->ev_pipe_in, 0));
        CK(cudaGraphLaunch(h->pipe_exec, h->pipe_stream));
        CK(cudaStreamSynchronize(h->pipe_stream));
        return ISX_OK;
    };
    const int rc = run();
    if (h->pool) {
        if (rc) h->pool->release_all();
        h->pool->wait_step();                                    // every row of `dst` is complete (and fenced) after this
    } else if (!rc && dst != h->h_obs) {
        std::memcpy(dst, h->h_obs, sizeof(float) * (size_t)h->d.E * h->d.N * ISX_OBS_DIM);   // the copy engine filled the pinned view
    }
    return rc;
}

int isx_step_pinned(isx_handle* h, float dt, void* stream) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    return host_step_into(h, dt, static_cast<cudaStream_t>(stream), h->h_obs);
}

// Tuning aid: ONE host-buffer step on the plain stream path with timing events around every range's kernels and copy.
// ms[4*i + 0..3] = kernels begin, kernels end, copy begin, copy end of range i, in ms since the step's first operation.
// Returns the number of ranges (or an error).  The step itself is a normal step (results in the pinned views).
int isx_pipe_timeline(isx_handle* h, float dt, void* stream, float* ms, int32_t cap_ranges) {
    if (!h || !ms) return fail(ISX_E_ARG, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    const size_t n = h->pipe.size();
    if ((size_t)cap_ranges < n) return fail(ISX_E_ARG, "need room for %d ranges", (int)n);
    std::vector<cudaEvent_t> ev(1 + 4 * n);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    const int rc = host_step_into(h, dt, st, h->h_obs, &ev);
    if (rc) return rc;
    for (size_t i = 0; i < 4 * n; ++i) CK(cudaEventElapsedTime(&ms[i], ev[0], ev[1 + i]));
    for (auto& e : ev) cudaEventDestroy(e);
    return (int)n;
}

int isx_host_views(isx_handle* h, float** actions, float** obs, float** reward, uint8_t** done, uint8_t** status,
                   uint8_t** terminated, uint8_t** truncated) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (actions) *actions = h->h_actions;
    if (obs) *obs = h->h_obs;
    if (reward) *reward = h->h_reward;
    if (done) *done = h->h_done;
    if (status) *status = h->h_status;
    if (terminated) *terminated = h->h_term;
    if (truncated) *truncated = h->h_trunc;
    return ISX_OK;
}

int isx_expand_obs_rows(const float* records32, const uint8_t* hit_index, int32_t lidar_rays, float* obs_rows, int64_t n_agents) {
    if (!records32 || !hit_index || !obs_rows) return fail(ISX_E_ARG, "null argument");
    if (lidar_rays < 1 || lidar_rays > ISX_MAX_RAYS || n_agents < 0) return fail(ISX_E_ARG, "bad sizes");
    expand_obs_rows(records32, hit_index, lidar_rays, obs_rows, (size_t)n_agents);
    return ISX_OK;
}

int isx_host_step_info(isx_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes, int32_t* host_threads, int32_t* ranges) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    const int64_t EN = (int64_t)h->d.E * h->d.N;
    if (h2d_bytes) *h2d_bytes = EN * 2 * (int64_t)sizeof(float);
    if (d2h_bytes) *d2h_bytes = (h->pool ? EN * (32 * (int64_t)sizeof(float) + h->d.R) : EN * (int64_t)sizeof(float) * ISX_OBS_DIM) + (int64_t)h->small_bytes;
    if (host_threads) *host_threads = h->host_threads;
    if (ranges) *ranges = (int32_t)h->pipe.size();
    return ISX_OK;
}

int isx_host_views_aux(isx_handle* h, int32_t** agents_alive, int32_t** step) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (agents_alive) *agents_alive = reinterpret_cast<int32_t*>(h->h_small + h->small_off[5]);
    if (step) *step = reinterpret_cast<int32_t*>(h->h_small + h->small_off[6]);
    return ISX_OK;
}

int isx_step_host(isx_handle* h, const float* actions, float dt, float* obs, float* reward, uint8_t* done, uint8_t* status,
                  uint8_t* terminated, uint8_t* truncated, void* stream) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (!actions) return fail(ISX_E_ARG, "actions is null");
    const Dev& d = h->d;
    const size_t EN = (size_t)d.E * d.N, E = (size_t)d.E;
    if (h->pool) h->pool->copy(h->h_actions, actions, sizeof(float) * EN * 2); else std::memcpy(h->h_actions, actions, sizeof(float) * EN * 2);
    // the obs rows are completed straight into the caller's buffer (no staging copy of 127 floats per agent)
    const int rc = host_step_into(h, dt, static_cast<cudaStream_t>(stream), obs ? obs : h->h_obs);
    if (rc) return rc;
    if (reward) std::memcpy(reward, h->h_reward, sizeof(float) * EN);
    if (done) std::memcpy(done, h->h_done, EN);
    if (status) std::memcpy(status, h->h_status, EN);
    if (terminated) std::memcpy(terminated, h->h_term, E);
    if (truncated) std::memcpy(truncated, h->h_trunc, E);
    return ISX_OK;
}

int isx_render(isx_handle* h, int32_t env, uint8_t* rgb_dev, void* stream) {
    if (!h || !rgb_dev) return fail(ISX_E_ARG, "null argument");
    if (env < 0 || env >= h->d.E) return fail(ISX_E_ARG, "env %d out of range", env);
    const isx_handle::Group& grp = group_of(h, env);
    CK(cudaSetDevice(h->device));
    CK(launch_render(grp.d, env - grp.first, rgb_dev, static_cast<cudaStream_t>(stream)));
    return ISX_OK;
}

int isx_get_buffers(isx_handle* h, isx_buffers* b) {
    if (!h || !b) return fail(ISX_E_ARG, "null argument");
    const Dev& d = h->d;
    b->obs = d.obs; b->reward = d.reward; b->done = d.done; b->status = d.status;
    b->terminated = d.terminated; b->truncated = d.truncated; b->agents_alive = d.agents_alive; b->step = d.step_count;
    b->lidar_hit = d.lidar_hit;
    b->ego_x = d.ex; b->ego_y = d.ey; b->ego_v = d.ev; b->ego_heading = d.eh; b->ego_steer = d.esteer; b->ego_acc = d.eacc;
    b->ego_prev_dist = d.epd; b->ego_prev_a0 = d.epa0; b->ego_prev_a1 = d.epa1; b->ego_path_index = d.epidx; b->ego_alive = d.ealive;
    b->npc_x = d.nx; b->npc_y = d.ny; b->npc_v = d.nv; b->npc_heading = d.nh; b->npc_steer = d.nsteer;
    b->npc_path_index = d.npidx; b->npc_route = d.nroute; b->npc_uid = d.nuid; b->npc_count = d.ncount;
    b->events = d.events; b->tick = d.tick;
    return ISX_OK;
}

int isx_get_env_state(isx_handle* h, int32_t env, isx_car_state* egos, isx_car_state* npcs, int32_t cap, int32_t* n_npcs,
                      int32_t* step_count, uint32_t* tick) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (env < 0 || env >= h->d.E) return fail(ISX_E_ARG, "env %d out of range", env);
    const isx_handle::Group& grp = group_of(h, env);
    const Dev& d = grp.d;                                 // the group's view; `env` becomes group-relative
    env -= grp.first;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    const size_t N = (size_t)d.N, M = (size_t)d.M, eo = (size_t)env * N, no = (size_t)env * M;
    std::vector<float> x, y, v, hd, st, ac, pdv, a0, a1;
    std::vector<int> pi, rt;
    std::vector<uint8_t> al;
    std::vector<uint32_t> uid;
    if (egos) {
        CK(pull(x, d.ex, eo, N)); CK(pull(y, d.ey, eo, N)); CK(pull(v, d.ev, eo, N)); CK(pull(hd, d.eh, eo, N));
        CK(pull(st, d.esteer, eo, N)); CK(pull(ac, d.eacc, eo, N)); CK(pull(pdv, d.epd, eo, N)); CK(pull(a0, d.epa0, eo, N));
        CK(pull(a1, d.epa1, eo, N)); CK(pull(pi, d.epidx, eo, N)); CK(pull(al, d.ealive, eo, N));
        for (size_t i = 0; i < N; ++i)
            egos[i] = isx_car_state{x[i], y[i], v[i], hd[i], ac[i], st[i], pdv[i], a0[i], a1[i], pi[i], (int32_t)i, al[i] ? 1 : 0, 0u, grp.routes[i].intent};
    }
    int cnt = 0;
    if (d.traffic) CK(cudaMemcpy(&cnt, d.ncount + env, sizeof(int), cudaMemcpyDeviceToHost));
    if (n_npcs) *n_npcs = cnt;
    if (npcs && cnt > 0) {
        CK(pull(x, d.nx, no, M)); CK(pull(y, d.ny, no, M)); CK(pull(v, d.nv, no, M)); CK(pull(hd, d.nh, no, M));
        CK(pull(st, d.nsteer, no, M)); CK(pull(pi, d.npidx, no, M)); CK(pull(rt, d.nroute, no, M)); CK(pull(uid, d.nuid, no, M));
        for (int i = 0; i < cnt && i < cap; ++i) {
            const size_t k = (size_t)i;
            npcs[i] = isx_car_state{x[k], y[k], v[k], hd[k], 0.0f, st[k], 0.0f, 0.0f, 0.0f, pi[k], rt[k], 1, uid[k], grp.routes[N + (size_t)rt[k]].intent};
        }
    }
    if (step_count) CK(cudaMemcpy(step_count, d.step_count + env, sizeof(int), cudaMemcpyDeviceToHost));
    if (tick) CK(cudaMemcpy(tick, d.tick + env, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return ISX_OK;
}

int isx_set_env_state(isx_handle* h, int32_t env, const isx_car_state* egos, const isx_car_state* npcs, int32_t n_npcs,
                      int32_t step_count, uint32_t tick) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (env < 0 || env >= h->d.E) return fail(ISX_E_ARG, "env %d out of range", env);
    const isx_handle::Group& grp = group_of(h, env);
    const Dev& d = grp.d;
    env -= grp.first;
    if (n_npcs < 0 || n_npcs > d.M || (n_npcs > 0 && !d.traffic)) return fail(ISX_E_ARG, "n_npcs %d exceeds capacity %d", n_npcs, d.traffic ? d.M : 0);
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    const size_t N = (size_t)d.N, eo = (size_t)env * N, no = (size_t)env * d.M;
    if (egos) {
        std::vector<float> x(N), y(N), v(N), hd(N), st(N), ac(N), pdv(N), a0(N), a1(N);
        std::vector<int> pi(N);
        std::vector<uint8_t> al(N);
        for (size_t i = 0; i < N; ++i) {
            x[i] = egos[i].x; y[i] = egos[i].y; v[i] = egos[i].v; hd[i] = egos[i].heading; st[i] = egos[i].steer; ac[i] = egos[i].acc;
            pdv[i] = egos[i].prev_dist; a0[i] = egos[i].prev_a0; a1[i] = egos[i].prev_a1; pi[i] = egos[i].path_index; al[i] = egos[i].alive ? 1 : 0;
            if (pi[i] < 0 || pi[i] >= PATH_LEN) return fail(ISX_E_ARG, "path_index out of range");
        }
        CK(push(x, d.ex, eo, N)); CK(push(y, d.ey, eo, N)); CK(push(v, d.ev, eo, N)); CK(push(hd, d.eh, eo, N));
        CK(push(st, d.esteer, eo, N)); CK(push(ac, d.eacc, eo, N)); CK(push(pdv, d.epd, eo, N)); CK(push(a0, d.epa0, eo, N));
        CK(push(a1, d.epa1, eo, N)); CK(push(pi, d.epidx, eo, N)); CK(push(al, d.ealive, eo, N));
    }
    if (d.traffic) {
        const size_t K = (size_t)n_npcs;
        if (K > 0) {
            if (!npcs) return fail(ISX_E_ARG, "npcs is null");
            std::vector<float> x(K), y(K), v(K), hd(K), st(K);
            std::vector<int> pi(K), rt(K);
            std::vector<uint32_t> uid(K);
            uint32_t max_uid = 0;
            for (size_t i = 0; i < K; ++i) {
                x[i] = npcs[i].x; y[i] = npcs[i].y; v[i] = npcs[i].v; hd[i] = npcs[i].heading; st[i] = npcs[i].steer;
                pi[i] = npcs[i].path_index; rt[i] = npcs[i].route; uid[i] = npcs[i].uid;
                if (rt[i] < 0 || rt[i] >= d.T) return fail(ISX_E_ARG, "npc route %d out of range", rt[i]);
                if (pi[i] < 0 || pi[i] >= PATH_LEN) return fail(ISX_E_ARG, "path_index out of range");
                if (uid[i] > max_uid) max_uid = uid[i];
            }
            CK(push(x, d.nx, no, K)); CK(push(y, d.ny, no, K)); CK(push(v, d.nv, no, K)); CK(push(hd, d.nh, no, K));
            CK(push(st, d.nsteer, no, K)); CK(push(pi, d.npidx, no, K)); CK(push(rt, d.nroute, no, K)); CK(push(uid, d.nuid, no, K));
            uint32_t nu = 0;
            CK(cudaMemcpy(&nu, d.next_uid + env, 4, cudaMemcpyDeviceToHost));
            if (max_uid >= nu) { nu = max_uid + 1; CK(cudaMemcpy(d.next_uid + env, &nu, 4, cudaMemcpyHostToDevice)); }
        }
        CK(cudaMemcpy(d.ncount + env, &n_npcs, sizeof(int), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(d.step_count + env, &step_count, sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.tick + env, &tick, sizeof(uint32_t), cudaMemcpyHostToDevice));
    // An injected state is a live episode: forget the end-of-episode flags of whatever ran before, or auto_reset would
    // throw the injected state away at the start of the next step.
    CK(cudaMemset(d.terminated + env, 0, 1));
    CK(cudaMemset(d.truncated + env, 0, 1));
    return ISX_OK;
}

// Lidar.h:11-14 / IntersectionEnv.cpp:112-128,411-415: the beam count is a property of the Lidar objects, which the
// reference swaps at run time (add_car_with_route builds 96-beam lidars, set_state leaves default 72-beam ones behind).
// Here it is one setting of the handle: new relative angles, every stored hit cleared (a fresh Lidar reads 250 px on
// every beam), obs columns beyond the new beam count zeroed, observation refreshed.  Synchronous.
int isx_set_lidar_rays(isx_handle* h, int32_t rays) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (rays < 1 || rays > ISX_MAX_RAYS) return fail(ISX_E_ARG, "lidar_rays must be in [1,%d]", ISX_MAX_RAYS);
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    std::vector<float> rel((size_t)ISX_MAX_RAYS, 0.0f);
    lidar_rel_angles(rays, rel.data());
    CK(cudaMemcpy(const_cast<float*>(h->d.rel_angle), rel.data(), sizeof(float) * rel.size(), cudaMemcpyHostToDevice));
    h->d.R = rays;
    for (auto& g : h->groups) { g.d.R = rays; g.cfg.lidar_rays = rays; }
    const size_t EN = (size_t)h->d.E * h->d.N;
    CK(cudaMemset(h->d.lidar_hit, 0, EN * ISX_MAX_RAYS));
    CK(cudaMemset(h->d.obs, 0, sizeof(float) * EN * ISX_OBS_DIM));
    if (h->pipe_exec) { cudaGraphExecDestroy(h->pipe_exec); h->pipe_exec = nullptr; }   // the captured launches carry the old Dev
    const int rc = isx_observe(h, nullptr);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    return ISX_OK;
}
// Run-time setters of what the reference lets a caller change on a live env object (reward_config.* are def_readwrite,
// bindings.cpp:33-42,63; configure / configure_traffic may be called at any time, IntersectionEnv.cpp:50-60).  group < 0 = all.
// The settings travel to the kernels by value with every launch, so they take effect at the next step.
static int for_groups(isx_handle* h, int32_t group, void (*fn)(isx_handle::Group&, const void*), const void* arg) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    if (group >= (int)h->groups.size()) return fail(ISX_E_ARG, "group %d out of range", group);
    for (int g = 0; g < (int)h->groups.size(); ++g) if (group < 0 || group == g) fn(h->groups[(size_t)g], arg);
    if (group <= 0) { const Dev& q = h->groups[0].d; h->d.rc = q.rc; h->d.use_team = q.use_team; h->d.respawn = q.respawn; h->d.max_steps = q.max_steps; }
    if (h->pipe_exec) { cudaGraphExecDestroy(h->pipe_exec); h->pipe_exec = nullptr; }   // re-capture with the new settings
    return ISX_OK;
}
int isx_set_reward(isx_handle* h, int32_t group, const float* k8) {
    if (!k8) return fail(ISX_E_ARG, "null argument");
    return for_groups(h, group, [](isx_handle::Group& g, const void* a) {
        const float* k = static_cast<const float*>(a);
        g.d.rc = RewardCfg{k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7]};
        for (int i = 0; i < 8; ++i) g.cfg.reward[i] = k[i];
    }, k8);
}
int isx_configure_episode(isx_handle* h, int32_t group, int32_t use_team, int32_t respawn, int32_t max_steps) {
    const int v[3] = {use_team, respawn, max_steps};
    return for_groups(h, group, [](isx_handle::Group& g, const void* a) {
        const int* q = static_cast<const int*>(a);
        g.d.use_team = q[0] != 0; g.d.respawn = q[1] != 0; g.d.max_steps = q[2];
        g.cfg.use_team_reward = q[0]; g.cfg.respawn_enabled = q[1]; g.cfg.max_steps = q[2];
    }, v);
}
int isx_set_traffic_density(isx_handle* h, int32_t group, float density) {
    return for_groups(h, group, [](isx_handle::Group& g, const void* a) {
        g.cfg.traffic_density = *static_cast<const float*>(a);
        g.last_dt = -1.0f;                                    // spawn probability is cached per dt
    }, &density);
}
int isx_lidar_rays(isx_handle* h) { return h ? h->d.R : fail(ISX_E_ARG, "null handle"); }

int isx_snapshot_create(isx_handle* h, isx_snapshot** out) {
    if (!h || !out) return fail(ISX_E_ARG, "null argument");
    CK(cudaSetDevice(h->device));
    isx_snapshot* s = new isx_snapshot();
    s->owner = h;
    const auto arrs = snapshot_arrays(h->d);
    size_t total = 0;
    for (auto& a : arrs) total += (((size_t)a.second * h->d.E) + 255) & ~(size_t)255;
    if (cudaMalloc((void**)&s->store, total) != cudaSuccess) { delete s; return fail(ISX_E_CUDA, "snapshot allocation of %zu bytes failed", total); }
    size_t off = 0;
    struct Row { unsigned char* live; const unsigned char* saved; unsigned bytes_per_env; };
    std::vector<Row> rows;
    for (auto& a : arrs) {
        s->entries.push_back(SnapEntry{static_cast<unsigned char*>(a.first), s->store + off, a.second});
        rows.push_back(Row{static_cast<unsigned char*>(a.first), s->store + off, a.second});
        off += (((size_t)a.second * h->d.E) + 255) & ~(size_t)255;
    }
    if (cudaMalloc(&s->dev_table, sizeof(Row) * rows.size()) != cudaSuccess ||
        cudaMemcpy(s->dev_table, rows.data(), sizeof(Row) * rows.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(s->store); if (s->dev_table) cudaFree(s->dev_table); delete s;
        return fail(ISX_E_CUDA, "snapshot table upload failed");
    }
    *out = s;
    return ISX_OK;
}
int isx_snapshot_destroy(isx_snapshot* s) {
    if (!s) return ISX_OK;
    cudaSetDevice(s->owner->device);
    cudaDeviceSynchronize();
    cudaFree(s->store);
    cudaFree(s->dev_table);
    delete s;
    return ISX_OK;
}
int isx_snapshot_save(isx_handle* h, isx_snapshot* s, void* stream) {
    if (!h || !s || s->owner != h) return fail(ISX_E_ARG, "snapshot does not belong to this handle");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    for (auto& e : s->entries) CK(cudaMemcpyAsync(e.saved, e.live, (size_t)e.bytes_per_env * h->d.E, cudaMemcpyDeviceToDevice, st));
    return ISX_OK;
}
int isx_snapshot_restore(isx_handle* h, isx_snapshot* s, const uint8_t* env_mask_dev, void* stream) {
    if (!h || !s || s->owner != h) return fail(ISX_E_ARG, "snapshot does not belong to this handle");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    if (!env_mask_dev) {
        for (auto& e : s->entries) CK(cudaMemcpyAsync(e.live, e.saved, (size_t)e.bytes_per_env * h->d.E, cudaMemcpyDeviceToDevice, st));
    } else {
        CK(launch_snapshot_restore(s->dev_table, (int)s->entries.size(), env_mask_dev, h->d.E, st));
    }
    return ISX_OK;
}

int isx_stats_read(isx_handle* h, isx_stats* out) {
    if (!h || !out) return fail(ISX_E_ARG, "null argument");
    CK(cudaSetDevice(h->device));
    CK(launch_reduce_stats(h->d, 0));
    unsigned long long raw[16];
    CK(cudaMemcpy(raw, h->d.stats, sizeof raw, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; ++i) out->status_hist[i] = (int64_t)raw[ST_HIST0 + i];
    out->npc_spawned = (int64_t)raw[ST_SPAWNED]; out->npc_removed = (int64_t)raw[ST_REMOVED];
    out->npc_collided = (int64_t)raw[ST_COLLIDED]; out->npc_overflow = (int64_t)raw[ST_OVERFLOW];
    out->env_resets = (int64_t)raw[ST_RESETS]; out->agent_steps = (int64_t)raw[ST_STEPS];
    std::memcpy(&out->reward_sum, &raw[15], 8);
    out->neighbor_tie_sorts = (int64_t)raw[ST_TIESORT];
    return ISX_OK;
}
// Debug aid (ISX_GUARD=1 at isx_create): counts the red zones around the handle's device buffers that no longer hold the
// canary.  *violations = 0 means no kernel stored outside its buffers since isx_create.  ISX_E_STATE when guards are off.
int isx_debug_check_guards(isx_handle* h, int64_t* violations) {
    if (!h || !violations) return fail(ISX_E_ARG, "null argument");
    if (!h->guard) return fail(ISX_E_STATE, "guards are off (set ISX_GUARD=1 before isx_create)");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    std::vector<unsigned char> zone(GUARD_BYTES);
    int64_t bad = 0;
    for (const auto& g : h->guarded) {
        const size_t padded = (g.second + 255) & ~(size_t)255;
        const unsigned char* zones[2] = {g.first - GUARD_BYTES, g.first + padded};
        for (int z = 0; z < 2; ++z) {
            CK(cudaMemcpy(zone.data(), zones[z], GUARD_BYTES, cudaMemcpyDeviceToHost));
            bool hit = false;
            for (size_t i = 0; i < GUARD_BYTES && !hit; ++i) hit = zone[i] != (unsigned char)GUARD_CANARY;
            bad += hit ? 1 : 0;
        }
        // the alignment padding between the end of the buffer and the upper red zone is canary, too
        if (padded > g.second) {
            std::vector<unsigned char> pad(padded - g.second);
            CK(cudaMemcpy(pad.data(), g.first + g.second, pad.size(), cudaMemcpyDeviceToHost));
            bool hit = false;
            for (unsigned char c : pad) hit = hit || c != (unsigned char)GUARD_CANARY;
            bad += hit ? 1 : 0;
        }
    }
    *violations = bad;
    return ISX_OK;
}

int isx_trace_read(isx_handle* h, long long* out16_per_env) {
    if (!h || !out16_per_env) return fail(ISX_E_ARG, "null argument");
    if (!h->d.trace) return fail(ISX_E_STATE, "tracing is off (set ISX_TRACE=1 before isx_create)");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out16_per_env, h->d.trace, sizeof(long long) * 16 * (size_t)h->d.E, cudaMemcpyDeviceToHost));
    return ISX_OK;
}
int isx_stats_reset(isx_handle* h) {
    if (!h) return fail(ISX_E_ARG, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(h->d.env_stats, 0, sizeof(uint32_t) * (size_t)h->d.E * STAT_SLOTS));
    CK(cudaMemset(h->d.stats, 0, sizeof(unsigned long long) * 16));
    return ISX_OK;
}
int isx_stats_device_ptrs(isx_handle* h, void** counters_i64, int32_t* n_counters, void** reward_sum_f64, void* stream) {
    if (!h || !counters_i64 || !reward_sum_f64) return fail(ISX_E_ARG, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    CK(launch_reduce_stats(h->d, st));
    CK(cudaStreamSynchronize(st));
    *counters_i64 = h->d.stats;                       // int64[ISX_STATS_COUNTERS]: integers only, safe to all-reduce(sum) as int64
    *reward_sum_f64 = h->d.stats + ISX_STATS_COUNTERS; // float64[1]: reduce separately, as a double
    if (n_counters) *n_counters = ISX_STATS_COUNTERS;
    return ISX_OK;
}

// Car::update (Car.cpp:9-40) / Car::check_collision (Car.cpp:105-141) for detached car records, evaluated ON THE GPU by the
// device functions the step kernels use (one thread; this is a unit-level convenience of the binding, not a fast path).
static int car_unit(int32_t device, int op, isx_car_state* a, const isx_car_state* b, float thr, float st, float dt, int32_t* out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ISX_E_CUDA, "no CUDA device: this library has no CPU path");
    if (device < 0 || device >= ndev) return fail(ISX_E_ARG, "device %d out of range", device);
    CK(cudaSetDevice(device));
    float* buf = nullptr;
    CK(cudaMalloc((void**)&buf, sizeof(float) * 16));
    float host[16] = {a->x, a->y, a->v, a->heading, a->acc, a->steer, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (b) { host[8] = b->x; host[9] = b->y; host[10] = b->heading; }
    cudaError_t e = cudaMemcpy(buf, host, sizeof host, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_car_unit(op, buf, buf + 8, thr, st, dt, reinterpret_cast<int*>(buf + 12), 0);
    if (e == cudaSuccess) e = cudaMemcpy(host, buf, sizeof host, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(ISX_E_CUDA, "car unit kernel failed: %s", cudaGetErrorString(e));
    if (op == 0) { a->x = host[0]; a->y = host[1]; a->v = host[2]; a->heading = host[3]; a->acc = host[4]; a->steer = host[5]; }
    else if (out) { int f; std::memcpy(&f, &host[12], 4); *out = f; }
    return ISX_OK;
}
int isx_car_update(int32_t device, isx_car_state* car, float throttle, float steer_input, float dt) {
    if (!car) return fail(ISX_E_ARG, "null argument");
    return car_unit(device, 0, car, nullptr, throttle, steer_input, dt, nullptr);
}
int isx_car_check_collision(int32_t device, const isx_car_state* a, const isx_car_state* b, int32_t* collide) {
    if (!a || !b || !collide) return fail(ISX_E_ARG, "null argument");
    isx_car_state tmp = *a;
    return car_unit(device, 1, &tmp, b, 0.0f, 0.0f, 0.0f, collide);
}

int isx_math_probe(int32_t device, int32_t n, const float* a, const float* b, float* sn, float* cs, float* tn, float* at, float* hy, float* wr) {
    if (n <= 0 || !a || !b) return fail(ISX_E_ARG, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ISX_E_CUDA, "no CUDA device");
    CK(cudaSetDevice(device));
    float* buf = nullptr;
    const size_t N = (size_t)n;
    CK(cudaMalloc((void**)&buf, sizeof(float) * N * 8));
    float *da = buf, *db = buf + N, *d0 = buf + 2 * N;
    cudaError_t e = cudaMemcpy(da, a, sizeof(float) * N, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, sizeof(float) * N, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_math_probe(n, da, db, d0, d0 + N, d0 + 2 * N, d0 + 3 * N, d0 + 4 * N, d0 + 5 * N, 0);
    float* outs[6] = {sn, cs, tn, at, hy, wr};
    for (int i = 0; i < 6 && e == cudaSuccess; ++i)
        if (outs[i]) e = cudaMemcpy(outs[i], d0 + (size_t)i * N, sizeof(float) * N, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(ISX_E_CUDA, "math probe failed: %s", cudaGetErrorString(e));
    return ISX_OK;
}

}  // extern "C"
