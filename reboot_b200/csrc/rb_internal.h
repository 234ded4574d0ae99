// rb_internal.h -- host-side internals shared by rb_world.cu and rb_dist.cu (not part of the ABI).
#pragma once
#include "../../include/reboot_physics.h"
#include "rb_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// launchers defined in rb_kernels.cu
void rb_launch_integrate_bin(const RbParams& P, int do_integrate, int do_bin, cudaStream_t st);
int rb_launch_scan(const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* tile_sums, cudaStream_t st);
uint32_t rb_scan_tiles(uint32_t n);
int rb_launch_scan_lookback(const uint32_t* in, uint32_t n, uint32_t* out, unsigned long long* status, uint32_t* ticket, cudaStream_t st);
void rb_launch_scatter(const RbParams& P, cudaStream_t st);
int rb_launch_collide(const RbParams& P, bool resolve, int variant, cudaStream_t st, cudaStream_t side = nullptr, cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr);
void rb_launch_clear(void* a, size_t bytes_a, uint32_t* b, uint32_t words_b, cudaStream_t st);
void rb_launch_vl_bin(const RbParams& P, cudaStream_t st);
void rb_launch_vl_scan(const RbParams& P, uint32_t ncols, unsigned long long* status, uint32_t* ticket, cudaStream_t st);
void rb_launch_vl_scatter(const RbParams& P, uint32_t ncols, unsigned long long* status, uint32_t* ticket, cudaStream_t st);
int rb_launch_vl_collide(const RbParams& P, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join, bool with_enumerate, uint32_t hot_rows, int phase);
void rb_launch_list_step_early(const RbParams& P, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join, uint32_t hot_rows);
extern "C" uint32_t rb_list_hot_rows(rb_world* w);
extern "C" int rb_build_triangle_grid_device(rb_world* w, uint32_t n);      // rb_extras.cu: static triangle column grid (n columns per axis) from w->h_tri
uint32_t rb_list_tile_bytes();
uint32_t rb_list_partner_slots();   // partners a persistent list holds
uint32_t rb_compound_tri_slots();   // triangle ids the split compound kernels stage per sphere
#ifndef RB_COMPOUND_DEFAULT
#define RB_COMPOUND_DEFAULT 2u       // compound bodies: 0 per-body kernel, 1 group kernel, 2 split enumeration / resolution
#endif
uint32_t rb_list_tile_rows();      // rows per triangle tile (= threads per block of the list step)
void rb_launch_entity_set_position(float4* pos, float4* mt, float4* pt, const float4* wt, const uint32_t* inv, uint32_t body, const float* xyz, cudaStream_t st);
void rb_launch_mt_refresh(const float4* pos, const float4* wt, float4* mt, uint32_t n, cudaStream_t st);
bool rb_have_device_launch();
void rb_launch_vl_gate(const RbParams& P, cudaStream_t st);
void rb_launch_apply_force(const uint32_t* flags, const float4* in, float4* out, const uint32_t* inv, uint32_t first, uint32_t count, int keep_w, cudaStream_t st);
void rb_launch_set_flags(uint32_t* flags, const uint32_t* values, const uint32_t* inv, uint32_t first, uint32_t count, uint32_t mask, cudaStream_t st);
void rb_launch_rows_to_ids_f4(const float4* rows, const uint32_t* gid, float4* out, uint32_t n, int w_one, cudaStream_t st);
void rb_launch_apply_force3(const uint32_t* flags, const float* in3, float4* out, const uint32_t* inv, uint32_t count, cudaStream_t st);
void rb_launch_rows_to_ids_f3(const float4* rows, const uint32_t* gid, float* out3, uint32_t n, cudaStream_t st);
void rb_launch_rows_to_ids_u32(const uint32_t* rows, const uint32_t* gid, uint32_t* out, uint32_t n, cudaStream_t st);
void rb_launch_ids_to_rows_f4(const float4* in, const uint32_t* inv, float4* rows, uint32_t first, uint32_t count, int keep_w, cudaStream_t st);
void rb_launch_reorder(const RbParams& P, const void* R, cudaStream_t st);
void rb_launch_owned_flags(const RbParams& P, uint32_t* out, uint32_t n, cudaStream_t st);
void rb_launch_model_matrices(const float4* w3, const uint32_t* gid, const float4* mt, const float4* pt, float4* cur, float4* prev, uint32_t n, int prev_is_default, cudaStream_t st);
void rb_launch_fill_f4(float4* p, uint32_t n, float4 v, cudaStream_t st);
// multi-GPU (rb_dist.cu)
struct RbDist;
int rb_dist_step_exchange(rb_world* w, int ms);
int rb_dist_list_exchange(rb_world* w, int list_mode);
extern "C" int rb_list_tick_tail(rb_world* w, bool reorder_due);
void rb_dist_destroy(RbDist* d);
int rb_dist_refresh_counts(rb_world* w);
int rb_dist_sticky_error(rb_world* w, unsigned long long imm_dropped, unsigned long long halo_ovf);   // RB_ERR_CAPACITY if a slab world ever overflowed
extern "C" int rb_reorder_rows_now(rb_world* w);
extern "C" bool rb_step_will_use_lists(rb_world* w);
int rb_dist_rows_moved(rb_world* w);   // the per-body arrays were swapped: refresh the device copy of the halo parameters
extern "C" int enqueue_broadphase_tail(rb_world* w);
extern "C" void rb_list_tick_params(rb_world* w);      // per-tick parameters of a persistent-list tick (before K1 is launched)
extern "C" int rb_plain_tick_params(rb_world* w);      // ... of every other path (per-tick enumeration, stand-alone phases): explicit model translations
extern "C" bool rb_list_tick_begin(rb_world* w, bool reorder_due);   // a list tick starts: true if the host makes it a rebuild tick
extern "C" void rb_list_tick_k1_enqueued(rb_world* w);               // ... its K1 is in the stream
extern "C" int rb_ensure_mt(rb_world* w);              // materialise the model translations list ticks keep implicit
#define RB_TB_BLOCKS 8                      // ring of per-tick counter blocks {contacts lo, hi, hot bodies, -}: a tick's block survives the next 6 ticks

extern thread_local std::string g_create_error;

#ifndef RB_PIPE
#define RB_PIPE 4     // depth of the pipelined host round trip (results arrive RB_PIPE - 1 calls later; 3 left the copy
#endif                // engines no slack: a tick's upload was only queued once the download two ticks back had finished)

struct DevBuf { void* p = nullptr; size_t bytes = 0; };

struct rb_world {
    rb_config cfg;
    rb_config cfg_user;                       // as given to rb_world_create (rb_build fills in the automatic values of cfg)
    std::string err;
    bool built = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // ---- host staging (until rb_build) ----
    std::vector<float> h_tri;                 // 9 floats per triangle
    std::vector<uint32_t> h_geom_start;       // CSR geom -> triangles
    std::vector<float> h_obj;                 // 4 floats per sphere (object space)
    std::vector<uint32_t> h_body_start;       // CSR body -> spheres
    std::vector<float> h_W;                   // 12 floats per body: 3x3 row-major + translation
    std::vector<float> h_pos, h_vel;          // 3 per body
    std::vector<uint32_t> h_flags;
    std::vector<uint32_t> h_sph_body;
    bool any_wt = false;
    std::vector<uint8_t> h_presummed;         // per body: its spheres are already W3x3 * centre, W[0] * radius (bodies carried over by a post-build add)
    // ---- device ----
    RbParams P;
    std::vector<DevBuf> allocs;
    size_t device_bytes = 0;
    uint32_t* d_tile_sums = nullptr;
    unsigned long long* d_scan_status = nullptr; uint32_t* d_scan_ticket = nullptr; size_t hist_clear_bytes = 0;
    int scan_variant = 2;
    uint32_t ncols = 0;
    float4* d_stage = nullptr; size_t stage_elems = 0;      // H2D staging for setters
    uint32_t* d_stage_u = nullptr; size_t stage_u_elems = 0;
    unsigned long long* h_counters = nullptr;               // pinned mirror
    // ---- bookkeeping ----
    uint64_t steps = 0, launches = 0;
    uint64_t pose_updates = 0;              // ticks / integrations / collisions so far (model-matrix export: _prevMVP default)
    uint64_t contacts_last = 0;
    // persistent candidate lists (single worlds of single-sphere bodies, rb_step only; RB_VERLET=0 turns them off)
    bool vl_on = false;                      // decided at rb_build
    bool vl_force = true;                    // the next list tick must rebuild (first tick, rows moved, state edited behind the lists' back)
    float vl_skin_cfg = 0.0f;                // 0 = auto
    // adaptive: while the lists get rebuilt on most ticks (free fall, violent scenes) the per-tick kernel is cheaper
    int group_list_mode = 0;                 // rb_step_group: this tick's mode of the world (0 per-tick enumeration, 2 lists, 3 lists + forced rebuild)
    uint64_t vl_decided_for = 0;             // step number (1-based) the use-lists decision below was taken for
    bool vl_active = true; uint64_t vl_resume_at = 0; uint32_t vl_backoff = 32; uint32_t vl_seen_r = 0, vl_seen_t = 0;
    bool vl_gate_ok = true;                  // recent rebuild rate is low enough for the device-launched chain (else the host enqueues it)
    bool vl_cdp = false;                     // K1 launches the rebuild chain from the device (library built with -rdc=true)
    int vl_hot_cfg = 0;                      // capacity of the hot-body table, 0 = auto (bodies / 128)
    uint32_t* vl_host = nullptr;             // mapped pinned {rebuilds, ticks, fast movers, hot bodies, tick of the report, expiry age, generation, age, previous generation: ticks served, its expiry}
    uint32_t* d_tb = nullptr;                // RB_TB_BLOCKS counter blocks, the list words (P.vl) right behind them
    unsigned long long* ncontact_dev = nullptr;   // where the contact count of the last step / detect lives
    uint32_t tick_seq = 0;                   // list ticks enqueued so far (version of the hot table entries)
    bool slim = false;                       // single-sphere bodies, W.t = 0, centred spheres: list ticks keep the model translations implicit
    bool mt_stale = false;                   // ... and the mt array has not been refreshed since
    uint32_t vl_spacing = 16;                // ... the period as applied: the re-ordering interval split evenly
    uint32_t vl_period = 16, vl_since = 0;   // single worlds: the host schedules the list rebuilds (every vl_period list ticks, adapted to the hot counts it sees)
    uint32_t vl_seen_tick = 0;               // last device report (tick number) the schedule has digested
    uint32_t vl_gen_expired = 0xffffffffu;   // the list generation whose expiry the schedule has acted on
    uint32_t vl_gen_seen = 0xffffffffu;      // the generation of the last report digested
    uint32_t vl_gen_flooded = 0xffffffffu;   // the generation whose flood report (lists do not pay right now) has been acted on
    uint32_t vl_judged_at = 0;               // list tick at which the lists-pay-off decision was last taken
    bool vl_tick_rebuilds = false;           // this list tick rebuilds at the host's request
    bool halo_split = false;                 // slab worlds: this list tick launched the early part of the list step before the halo wait
    bool vl_never_off = false;               // RB_VL_NEVER_OFF=1 (tests): never sit a rebuild-heavy phase out on the per-tick kernel
    unsigned long long* h_sticky = nullptr;  // pinned: sticky overflow counters of slab worlds, fetched at the pace points of rb_step
    void* d_cull_out = nullptr; void* d_cull_cubes = nullptr; uint32_t* d_cull_cnt = nullptr; uint32_t cull_cap = 0;   // frustum-culling scratch (rb_extras.cu)
    float* d_cells_box = nullptr; uint32_t* d_cells_n = nullptr;      // rb_get_static_cells scratch, kept with the world
    cudaEvent_t vl_pace[2] = { nullptr, nullptr }; bool vl_pace_armed[2] = { false, false };   // bounded host run-ahead (see rb_step)
    cudaStream_t vl_side = nullptr; cudaEvent_t vl_fork = nullptr, vl_join = nullptr;   // the hot bodies' launch runs next to the list kernel
    int collide_variant = 2;                 // 1 first generation, 2 fused phased kernel (default), 3 enumerate + resolve launches (measured slower: 0.342 vs 0.331 ms)
    // ---- spatial row re-ordering (single-sphere, single-GPU worlds) ----
    int reorder_every = 64;                  // ticks between re-orderings (0 = never); measured on C3: 32 -> 0.1650, 64 -> 0.1611, 128 -> 0.1611 ms per tick
    uint32_t* d_inv = nullptr;               // body id -> row
    uint32_t* d_gid_alt = nullptr; uint32_t* d_flags_alt = nullptr; uint32_t* d_rest = nullptr; uint32_t* d_key_alt = nullptr;
    float4* alt4[RB_REORDER_MAX] = {};
    float4* d_io4 = nullptr; uint32_t* d_io_u = nullptr;   // id-ordered staging for host IO
    float4* d_w3 = nullptr; float4* d_mm_cur = nullptr; float4* d_mm_prev = nullptr; bool any_w3 = false;   // model-matrix export
    uint32_t* d_oflag = nullptr; uint32_t* d_orank = nullptr; uint8_t* d_oscan = nullptr; size_t oscan_bytes = 0;   // slab re-ordering scratch
    bool profile = false;
    float kernel_ms[5] = { 0, 0, 0, 0, 0 };
    cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    // ---- pipelined host round trip (rb_step_host_async): copy engines overlap the tick ----
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t pe_h2d[RB_PIPE] = {}, pe_force[RB_PIPE] = {}, pe_step[RB_PIPE] = {}, pe_d2h[RB_PIPE] = {};
    float4* p_force[RB_PIPE] = {}; float4* p_pos[RB_PIPE] = {};
    unsigned long long* p_cnt = nullptr;     // pinned, RB_PIPE entries
    uint64_t p_calls = 0;
    uint32_t p_rows = 0;                     // rows each staging buffer of the ring holds
    RbDist* dist = nullptr;
    rb_stats_t dist_stats;
    // ---- slab worlds (rb_dist.cu) ----
    uint32_t cap_local = 0;                  // capacity of the per-body / per-sphere arrays
    uint32_t ghost_cap = 0;                  // ghosts per face the halo buffers hold
    float ext_max = 0.f;                     // compound slab worlds: the farthest any body reaches from its model translation along x (+ margin cap)
    uint32_t* d_dn = nullptr;                // device {n_owned, n_ghost}
    uint32_t* d_gid = nullptr;
    std::vector<uint32_t> h_gid;             // ids given before rb_build
    uint32_t n_owned_host = 0;               // last value read back
};

static inline int fail(rb_world* w, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (w) w->err = buf; else g_create_error = buf;
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(w, RB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

template <class T> static inline int dev_alloc(rb_world* w, T** out, size_t count) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(w, RB_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    DevBuf b; b.p = p; b.bytes = bytes; w->allocs.push_back(b); w->device_bytes += bytes;
    *out = (T*)p;
    return RB_OK;
}
#define ALLOC(ptr, count) do { int r_ = dev_alloc(w, &(ptr), (count)); if (r_) return r_; } while (0)

