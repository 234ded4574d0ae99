// rb_world.cu -- host side of libreboot_b200.so: the world object behind the C ABI declared in
// include/reboot_physics.h. Validates inputs, owns all device memory, builds the static triangle
// column grid once, and enqueues the step kernels (rb_kernels.cu) on the world's stream.
// No CPU path: every entry point that computes needs the CUDA device.
#include "rb_internal.h"

thread_local std::string g_create_error;

static bool finite_all(const float* p, size_t n) { for (size_t i = 0; i < n; i++) if (!std::isfinite(p[i])) return false; return true; }


extern "C" {

const char* rb_version(void) { return "reboot_b200 0.1 (sm_100a)"; }

// ghosts per face ~ n * reach / slab width; 16x headroom (overflow is detected and reported)
uint32_t rb_halo_ghost_cap_auto(uint32_t n_bodies, float r_max, float margin_cap, float slab_width) {
    double mc = margin_cap > 0.f ? margin_cap : 0.5 * r_max;
    double frac = 2.0 * (r_max + mc) / std::max(1e-3, (double)slab_width);
    return (uint32_t)std::min<double>(n_bodies / 2 + 8192.0, std::max(8192.0, 16.0 * n_bodies * frac));
}
void rb_config_default(rb_config* c) {
    memset(c, 0, sizeof *c);
    c->world_half = 1000.0f; c->gravity = -9.8f; c->friction = 0.95f; c->pushout = 1.0001f; c->ss_damp = 0.1f;
    c->margin_cap = 0.0f; c->device = 0; c->stream = nullptr; c->max_contacts = 0;
    c->slab_lo = -1000.0f; c->slab_hi = 1000.0f; c->rank = 0; c->nranks = 1; c->rm_max_hint = 0.0f; c->halo_ghost_cap = 0;
    c->sphere_stride = 0; c->body_reach_hint = 0.0f; c->slab_forces = 0;
}

const char* rb_last_error(const rb_world* w) { return w ? w->err.c_str() : g_create_error.c_str(); }

int rb_world_create(const rb_config* cfg, rb_world** out) {
    if (!out) return fail(nullptr, RB_ERR_INVALID, "rb_world_create: out is NULL");
    *out = nullptr;
    rb_config c; if (cfg) c = *cfg; else rb_config_default(&c);
    if (!(c.world_half > 0) || !(c.margin_cap >= 0.f) || c.nranks < 1 || c.rank < 0 || c.rank >= c.nranks)
        return fail(nullptr, RB_ERR_INVALID, "rb_world_create: bad config");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RB_ERR_NO_DEVICE, "no CUDA device (%s): reboot_b200 has no CPU path", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (c.device < 0 || c.device >= ndev) return fail(nullptr, RB_ERR_INVALID, "device %d out of range (%d devices)", c.device, ndev);
    e = cudaSetDevice(c.device);
    if (e != cudaSuccess) return fail(nullptr, RB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    rb_world* w = new rb_world();
    w->cfg = c; w->cfg_user = c;
    memset(&w->P, 0, sizeof w->P);
    memset(&w->dist_stats, 0, sizeof w->dist_stats);
    if (c.stream) w->stream = (cudaStream_t)c.stream;
    else { e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking); if (e != cudaSuccess) { delete w; return fail(nullptr, RB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); } w->own_stream = true; }
    w->h_geom_start.push_back(0);
    w->h_body_start.push_back(0);
    { const char* v = getenv("RB_COLLIDE_V1"); if (v && *v == '1') w->collide_variant = 1; }
    { const char* v = getenv("RB_COLLIDE_VARIANT"); if (v && *v) w->collide_variant = atoi(v); }
    { const char* v = getenv("RB_VERLET"); w->vl_on = !(v && *v == '0'); }          // persistent candidate lists (default on where they apply)
    { const char* v = getenv("RB_VERLET_SKIN"); if (v && *v) w->vl_skin_cfg = (float)atof(v); }
    { const char* v = getenv("RB_VL_NEVER_OFF"); w->vl_never_off = v && *v == '1'; }      // tests: stay on the lists however often they rebuild
    { const char* v = getenv("RB_VERLET_HOT"); if (v && *v) w->vl_hot_cfg = atoi(v); }
    { const char* v = getenv("RB_VERLET_CDP"); w->vl_cdp = rb_have_device_launch() && !(v && *v == '0'); }
    { const char* v = getenv("RB_SCAN_V1"); if (v && *v == '1') w->scan_variant = 1; }
    { const char* v = getenv("RB_REORDER_EVERY"); if (v && *v) w->reorder_every = atoi(v); }
    *out = w;
    return RB_OK;
}

void rb_world_destroy(rb_world* w) {
    if (!w) return;
    cudaSetDevice(w->cfg.device);
    cudaStreamSynchronize(w->stream);
    if (w->dist) rb_dist_destroy(w->dist);
    for (auto& b : w->allocs) cudaFree(b.p);
    if (w->h_counters) cudaFreeHost(w->h_counters);
    if (w->vl_host) cudaFreeHost(w->vl_host);
    if (w->h_sticky) cudaFreeHost(w->h_sticky);
    cudaFree(w->d_cull_out); cudaFree(w->d_cull_cubes); cudaFree(w->d_cull_cnt);
    if (w->vl_side) { cudaStreamDestroy(w->vl_side); cudaEventDestroy(w->vl_fork); cudaEventDestroy(w->vl_join); }
    if (w->vl_pace[0]) { cudaEventDestroy(w->vl_pace[0]); cudaEventDestroy(w->vl_pace[1]); }
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    if (w->s_h2d) {
        cudaStreamSynchronize(w->s_h2d); cudaStreamSynchronize(w->s_d2h);
        for (int k = 0; k < RB_PIPE; k++) { cudaEventDestroy(w->pe_h2d[k]); cudaEventDestroy(w->pe_force[k]); cudaEventDestroy(w->pe_step[k]); cudaEventDestroy(w->pe_d2h[k]); }
        cudaStreamDestroy(w->s_h2d); cudaStreamDestroy(w->s_d2h); cudaFreeHost(w->p_cnt);
    }
    if (w->own_stream) cudaStreamDestroy(w->stream);
    delete w;
}

int rb_add_static_geometry(rb_world* w, uint32_t ntri, const float* xyz9, uint32_t* geom_id) {
    if (!w) return RB_ERR_INVALID;
    if (w->built) return fail(w, RB_ERR_INVALID, "rb_add_static_geometry after rb_build (Physics::addEntity never inserts triangles, Q9)");
    if (ntri && !xyz9) return fail(w, RB_ERR_INVALID, "xyz9 is NULL");
    if (!finite_all(xyz9, (size_t)ntri * 9)) return fail(w, RB_ERR_INVALID, "non-finite triangle coordinate");
    if ((uint64_t)w->h_tri.size() / 9 + ntri >= 0x7FFFFFFFull) return fail(w, RB_ERR_INVALID, "too many triangles");
    w->h_tri.insert(w->h_tri.end(), xyz9, xyz9 + (size_t)ntri * 9);
    w->h_geom_start.push_back((uint32_t)(w->h_tri.size() / 9));
    if (geom_id) *geom_id = (uint32_t)w->h_geom_start.size() - 2;
    return RB_OK;
}

static int add_body(rb_world* w, uint32_t nsph, const float* obj4, const float* W12, const float* pos3, const float* vel3, uint32_t flags, bool presummed = false) {
    if (nsph == 0) return fail(w, RB_ERR_INVALID, "a model needs at least one sphere");
    if (!finite_all(obj4, (size_t)nsph * 4) || !finite_all(pos3, 3) || !finite_all(vel3, 3) || !finite_all(W12, 12))
        return fail(w, RB_ERR_INVALID, "non-finite model input");
    uint32_t b = (uint32_t)w->h_flags.size();
    w->h_obj.insert(w->h_obj.end(), obj4, obj4 + (size_t)nsph * 4);
    for (uint32_t i = 0; i < nsph; i++) w->h_sph_body.push_back(b);
    w->h_body_start.push_back((uint32_t)(w->h_obj.size() / 4));
    w->h_W.insert(w->h_W.end(), W12, W12 + 12);
    if (W12[9] != 0.f || W12[10] != 0.f || W12[11] != 0.f) w->any_wt = true;
    w->h_pos.insert(w->h_pos.end(), pos3, pos3 + 3);
    w->h_vel.insert(w->h_vel.end(), vel3, vel3 + 3);
    w->h_flags.push_back(flags & (RB_ACTIVE | RB_GRAVITY));
    w->h_presummed.push_back(presummed ? 1 : 0);
    return RB_OK;
}
struct GrowBody { uint32_t nsph; const float* obj4; float W12[12]; const float* pos3; const float* vel3; uint32_t flags; };
static int world_grow(rb_world* w, const std::vector<GrowBody>& add, uint32_t* first_body_id);

int rb_add_model(rb_world* w, uint32_t nspheres, const float* obj4, const float* X, const rb_body_init* init, uint32_t* body_id) {
    if (!w || !init || !obj4) return w ? fail(w, RB_ERR_INVALID, "NULL argument") : RB_ERR_INVALID;
    float W12[12] = { 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0 };
    if (X) {
        if (X[12] != 0.f || X[13] != 0.f || X[14] != 0.f || X[15] != 1.f) return fail(w, RB_ERR_INVALID, "world transform must be affine (last row 0 0 0 1)");
        W12[0] = X[0]; W12[1] = X[1]; W12[2] = X[2]; W12[3] = X[4]; W12[4] = X[5]; W12[5] = X[6]; W12[6] = X[8]; W12[7] = X[9]; W12[8] = X[10];
        W12[9] = X[3]; W12[10] = X[7]; W12[11] = X[11];
    }
    if (w->built) {      // Physics::addEntity (physics/src/Physics.cpp:47-49): the entity joins the next tick
        GrowBody g; g.nsph = nspheres; g.obj4 = obj4; memcpy(g.W12, W12, sizeof W12); g.pos3 = init->pos; g.vel3 = init->vel; g.flags = init->flags;
        return world_grow(w, std::vector<GrowBody>(1, g), body_id);
    }
    int r = add_body(w, nspheres, obj4, W12, init->pos, init->vel, init->flags);
    if (r) return r;
    if (body_id) *body_id = (uint32_t)w->h_flags.size() - 1;
    return RB_OK;
}

int rb_add_models(rb_world* w, uint32_t n, const uint32_t* sphere_start, const float* obj4, const float* world_scale,
                  const float* pos3, const float* vel3, const uint32_t* flags, uint32_t* first_body_id) {
    if (!w || !sphere_start || !obj4 || !pos3 || !vel3 || !flags) return w ? fail(w, RB_ERR_INVALID, "NULL argument") : RB_ERR_INVALID;
    if (w->built) {      // a batch of Physics::addEntity calls: the world is re-laid out once
        std::vector<GrowBody> add(n);
        for (uint32_t i = 0; i < n; i++) {
            if (sphere_start[i + 1] < sphere_start[i]) return fail(w, RB_ERR_INVALID, "sphere_start must be non-decreasing");
            const float sc = world_scale ? world_scale[i] : 1.0f;
            GrowBody& g = add[i];
            g.nsph = sphere_start[i + 1] - sphere_start[i]; g.obj4 = obj4 + (size_t)sphere_start[i] * 4;
            const float W12[12] = { sc, 0, 0, 0, sc, 0, 0, 0, sc, 0, 0, 0 }; memcpy(g.W12, W12, sizeof W12);
            g.pos3 = pos3 + 3 * (size_t)i; g.vel3 = vel3 + 3 * (size_t)i; g.flags = flags[i];
        }
        return world_grow(w, add, first_body_id);
    }
    uint32_t first = (uint32_t)w->h_flags.size();
    w->h_obj.reserve(w->h_obj.size() + (size_t)sphere_start[n] * 4);
    for (uint32_t i = 0; i < n; i++) {
        float s = world_scale ? world_scale[i] : 1.0f;
        float W12[12] = { s, 0, 0, 0, s, 0, 0, 0, s, 0, 0, 0 };
        if (sphere_start[i + 1] < sphere_start[i]) return fail(w, RB_ERR_INVALID, "sphere_start must be non-decreasing");
        int r = add_body(w, sphere_start[i + 1] - sphere_start[i], obj4 + (size_t)sphere_start[i] * 4, W12, pos3 + 3 * (size_t)i, vel3 + 3 * (size_t)i, flags[i]);
        if (r) return r;
    }
    if (first_body_id) *first_body_id = first;
    return RB_OK;
}

// ---- static triangle column grid (replaces the per-leaf triangle sets of OSP::generateGeometryOSP,
// physics/src/OSP.cpp:25-60, 242-332; built once, triangles never move) --------------------------
static int build_triangle_grid(rb_world* w) {
    RbParams& P = w->P;
    const uint32_t nt = (uint32_t)(w->h_tri.size() / 9);
    const float half = w->cfg.world_half;
    P.nt = nt;
    // column width = the MEDIAN triangle extent: a terrain lattice with a few large triangles on it (houses) keeps columns that
    // coincide with its cells -- the mean would be off by a fraction of a percent, every column would straddle two cells per
    // axis and hold four times the triangles (C5: 1.001 instead of 1.0 left 3 % of the spheres with more candidates than the
    // fast walk takes)
    std::vector<float> extent(nt);
    for (uint32_t t = 0; t < nt; t++) {
        const float* p = &w->h_tri[(size_t)t * 9];
        float dx = std::max(p[0], std::max(p[3], p[6])) - std::min(p[0], std::min(p[3], p[6]));
        float dz = std::max(p[2], std::max(p[5], p[8])) - std::min(p[2], std::min(p[5], p[8]));
        extent[t] = std::max(dx, dz);
    }
    double mean = 2.0 * half;
    if (nt) { std::nth_element(extent.begin(), extent.begin() + nt / 2, extent.end()); mean = extent[nt / 2]; }
    std::vector<float>().swap(extent);
    if (!(mean > 0)) mean = 2.0 * half;
    long n = std::lround(2.0 * half / mean);
    n = std::max(1L, std::min(4096L, n));
    P.tcx = P.tcz = (uint32_t)n;
    P.t_inv_w = (float)n / (2.0f * half);
    // registration, the device triangle records and the per-column y ranges: on the device (rb_extras.cu)
    return rb_build_triangle_grid_device(w, (uint32_t)n);
}

static int ensure_stage(rb_world* w, size_t elems) {
    if (w->stage_elems >= elems) return RB_OK;
    float4* p; ALLOC(p, elems);
    w->d_stage = p; w->stage_elems = elems;   // the old buffer stays in allocs until destroy
    return RB_OK;
}
static int ensure_stage_u(rb_world* w, size_t elems) {
    if (w->stage_u_elems >= elems) return RB_OK;
    uint32_t* p; ALLOC(p, elems);
    w->d_stage_u = p; w->stage_u_elems = elems;
    return RB_OK;
}

// ---- id-ordered host IO over spatially re-ordered rows ------------------------------------------
static int ensure_io(rb_world* w) {
    if (w->d_io4) return RB_OK;
    ALLOC(w->d_io4, w->P.nb); ALLOC(w->d_io_u, w->P.nb);
    return RB_OK;
}
// w_one: the w lane reads 1 (single-sphere bodies keep their world radius in pos.w; the interface says positions carry w = 1)
static int read_rows_f4(rb_world* w, const float4* rows, float* host, size_t n, int w_one = 0) {
    if (!w->d_inv && !w_one) { CU(cudaMemcpyAsync(host, rows, n * 16, cudaMemcpyDeviceToHost, w->stream)); return RB_OK; }
    int r = ensure_io(w); if (r) return r;
    rb_launch_rows_to_ids_f4(rows, w->d_inv ? w->P.gid : nullptr, w->d_io4, (uint32_t)n, w_one, w->stream); w->launches++;
    CU(cudaMemcpyAsync(host, w->d_io4, n * 16, cudaMemcpyDeviceToHost, w->stream));
    return RB_OK;
}
static int read_rows_u32(rb_world* w, const uint32_t* rows, uint32_t* host, size_t n) {
    if (!w->d_inv) { CU(cudaMemcpyAsync(host, rows, n * 4, cudaMemcpyDeviceToHost, w->stream)); return RB_OK; }
    int r = ensure_io(w); if (r) return r;
    rb_launch_rows_to_ids_u32(rows, w->P.gid, w->d_io_u, (uint32_t)n, w->stream); w->launches++;
    CU(cudaMemcpyAsync(host, w->d_io_u, n * 4, cudaMemcpyDeviceToHost, w->stream));
    return RB_OK;
}
// keep_w: the w lane of the rows survives (single-sphere bodies keep their world radius in pos.w)
static int write_rows_f4(rb_world* w, float4* rows, uint32_t first, uint32_t count, const float* host, int keep_w = 0) {
    if (!w->d_inv && !keep_w) { CU(cudaMemcpyAsync(rows + first, host, (size_t)count * 16, cudaMemcpyHostToDevice, w->stream)); return RB_OK; }
    int r = ensure_stage(w, count); if (r) return r;
    CU(cudaMemcpyAsync(w->d_stage, host, (size_t)count * 16, cudaMemcpyHostToDevice, w->stream));
    rb_launch_ids_to_rows_f4(w->d_stage, w->d_inv, rows, first, count, keep_w, w->stream); w->launches++;
    return RB_OK;
}

// List ticks of slim worlds keep Entity::_mvp's translation implicit (0 + p); everybody else wants the array
int rb_ensure_mt(rb_world* w) {
    if (!w->mt_stale) return RB_OK;
    const uint32_t n = w->d_dn ? w->cap_local : w->P.nb;
    rb_launch_mt_refresh(w->P.pos, nullptr, w->P.mt, n, w->stream); w->launches++;
    CU(cudaGetLastError());
    w->mt_stale = false;
    return RB_OK;
}
void rb_list_tick_params(rb_world* w) {
    RbParams& P = w->P;
    P.tick = ++w->tick_seq; if (!P.tick) P.tick = ++w->tick_seq;
    // single worlds rotate through the ring of counter blocks and K1 zeroes the next one; slab worlds use the last block,
    // which their per-tick clear launch zeroes together with the rebuild flag right behind it
    const uint32_t blk = P.soft ? w->tick_seq % RB_TB_BLOCKS : RB_TB_BLOCKS - 1;
    P.ncontact = (unsigned long long*)(w->d_tb + 4 * blk); P.nhot = w->d_tb + 4 * blk + 2;
    P.tb_next = P.soft ? w->d_tb + 4 * ((blk + 1) % RB_TB_BLOCKS) : nullptr;
    w->ncontact_dev = P.ncontact;
    P.lazy_mt = w->slim ? 1u : 0u; P.mt_in_valid = w->mt_stale ? 0u : 1u;
}
int rb_plain_tick_params(rb_world* w) {
    RbParams& P = w->P;
    int r = rb_ensure_mt(w); if (r) return r;
    P.ncontact = P.counters + RB_C_NCONTACT; w->ncontact_dev = P.ncontact;
    P.tb_next = nullptr; P.lazy_mt = 0u; P.mt_in_valid = 1u;
    return RB_OK;
}

static int reorder_rows(rb_world* w, bool intick = false);
// K7 driver: put the body rows in the order of the last counting sort (needs a valid sort: call right
// after a broadphase). Single-sphere, single-GPU worlds only.
int rb_reorder_rows_now(rb_world* w) { return reorder_rows(w); }
static int reorder_rows(rb_world* w, bool intick) {
    RbParams& P = w->P;
    if (!P.single || P.nb == 0) return RB_OK;
    const bool slab = w->d_dn != nullptr;
    const size_t nb = P.nb;
    // intick: between the sort and the enumeration of a list-rebuild tick -- K1's per-row outputs (world spheres,
    // build poses, keys) travel too and the sorted ids are rewritten, so the lists are built for the new rows
    // (a stale mt array -- list ticks of slim worlds keep it implicit -- does not travel: rb_ensure_mt rewrites all of it)
    float4** cur[RB_REORDER_MAX] = { &P.pos, &P.vel, w->mt_stale ? nullptr : &P.mt, &P.pt, const_cast<float4**>(&P.sph_obj), const_cast<float4**>(&P.wt), &P.force, &P.torque, &P.apos, &P.avel,
                                     intick ? &P.ws : nullptr, intick ? &P.ws_build : nullptr };
    if (!w->d_gid_alt) {
        if (!slab) { ALLOC(w->d_inv, nb); ALLOC(w->d_gid, nb); }
        ALLOC(w->d_gid_alt, nb); ALLOC(w->d_flags_alt, nb); ALLOC(w->d_rest, 4);
        if (slab) {
            size_t tiles = rb_scan_tiles((uint32_t)nb + 1) + 1;
            ALLOC(w->d_oflag, nb + 8); ALLOC(w->d_orank, nb + 8);
            w->oscan_bytes = 16 + tiles * 8; ALLOC(w->d_oscan, w->oscan_bytes);
        }
    }
    RbReorder R; memset(&R, 0, sizeof R);
    int slot_of[RB_REORDER_MAX];
    for (int a = 0; a < RB_REORDER_MAX; a++) {
        slot_of[a] = -1;
        if (!cur[a] || !*cur[a]) continue;
        if (!w->alt4[a]) ALLOC(w->alt4[a], nb);
        R.src4[R.n4] = *cur[a]; R.dst4[R.n4] = w->alt4[a]; slot_of[a] = R.n4++;
    }
    R.src_flags = P.flags; R.dst_flags = w->d_flags_alt;
    if (intick) {
        if (!w->d_key_alt) ALLOC(w->d_key_alt, nb);
        R.src_key = P.key; R.dst_key = w->d_key_alt; R.sorted_id_fix = P.sorted_id;
    }
    uint32_t* gid_cur = const_cast<uint32_t*>(P.gid);
    R.src_gid = P.gid; R.dst_gid = gid_cur == w->d_gid ? w->d_gid_alt : w->d_gid; R.inv = slab ? nullptr : w->d_inv; R.rest_counter = w->d_rest;
    CU(cudaMemsetAsync(w->d_rest, 0, 4, w->stream));
    if (slab && !intick) {
        // owned rows keep their relative order in the sorted array: rank = exclusive count of owned slots
        CU(cudaMemsetAsync(w->d_oscan, 0, w->oscan_bytes, w->stream));
        rb_launch_owned_flags(P, w->d_oflag, (uint32_t)nb, w->stream);
        w->launches += 1 + rb_launch_scan_lookback(w->d_oflag, (uint32_t)nb, w->d_orank, (unsigned long long*)(w->d_oscan + 16), (uint32_t*)w->d_oscan, w->stream);
        R.owned_rank = w->d_orank; R.n_slots = (uint32_t)nb;
    }
    rb_launch_reorder(P, &R, w->stream); w->launches++;
    CU(cudaGetLastError());
    for (int a = 0; a < RB_REORDER_MAX; a++) if (slot_of[a] >= 0) { float4* old = *cur[a]; *cur[a] = w->alt4[a]; w->alt4[a] = old; }
    { uint32_t* old = P.flags; P.flags = w->d_flags_alt; w->d_flags_alt = old; }
    if (intick) { uint32_t* old = P.key; P.key = w->d_key_alt; w->d_key_alt = old; }
    P.gid = R.dst_gid;
    if (slab) { w->d_gid = R.dst_gid; w->d_gid_alt = gid_cur; rb_dist_rows_moved(w); }
    return RB_OK;
}
static int enqueue_broadphase(rb_world* w, int do_integrate);
static int ensure_angular(rb_world* w);
// which kernels step compound bodies (all bit-identical): RB_COMPOUND=legacy|group|split (RB_COMPOUND_GROUP=1 = group)
static uint32_t compound_kernel_choice() {
    const char* v = getenv("RB_COMPOUND");
    if (v && !strcmp(v, "legacy")) return 0u;
    if (v && !strcmp(v, "group")) return 1u;
    if (v && !strcmp(v, "split")) return 2u;
    const char* g = getenv("RB_COMPOUND_GROUP");
    if (g && *g == '1') return 1u;
    return RB_COMPOUND_DEFAULT;
}
int rb_build(rb_world* w) {
    if (!w) return RB_ERR_INVALID;
    if (w->built) return fail(w, RB_ERR_INVALID, "rb_build is call-once (Physics::addEntities, SURVEY Q9)");
    CU(cudaSetDevice(w->cfg.device));
    RbParams& P = w->P;
    const uint32_t nb = (uint32_t)w->h_flags.size();
    const uint32_t ns = (uint32_t)(w->h_obj.size() / 4);
    const bool slab = w->cfg.nranks > 1;
    P.single = (ns == nb) ? 1u : 0u;
    if (slab && w->any_wt) return fail(w, RB_ERR_INVALID, "slab (multi-GPU) worlds need W.t = 0");
    // compound bodies in slab worlds: rows come and go (migration), so a body keeps its spheres at a fixed stride K -- the largest
    // sphere count anywhere, which every rank must agree on (rb_config.sphere_stride) -- instead of a CSR that would have to be compacted
    uint32_t K = 0;
    if (slab && (!P.single || w->cfg.sphere_stride > 1)) {
        uint32_t kmax = 1; for (uint32_t b = 0; b < nb; b++) kmax = std::max(kmax, w->h_body_start[b + 1] - w->h_body_start[b]);
        if (w->cfg.sphere_stride && kmax > w->cfg.sphere_stride) return fail(w, RB_ERR_INVALID, "a body has %u spheres, rb_config.sphere_stride is %u", kmax, w->cfg.sphere_stride);
        K = std::max(kmax, w->cfg.sphere_stride);
        P.single = 0u; P.kstride = K;
        P.grp_lg = 2; while ((1u << P.grp_lg) < K && P.grp_lg < 5) P.grp_lg++;
        P.grp_on = compound_kernel_choice();
        // how far a body reaches from its model translation along x (the ghost criterion of K1): this rank's bodies and the hint
        float ext = 0.f;
        for (uint32_t b = 0; b < nb; b++) {
            const float* W = &w->h_W[(size_t)b * 12];
            for (uint32_t q = w->h_body_start[b]; q < w->h_body_start[b + 1]; q++) {
                const float* o = &w->h_obj[(size_t)q * 4];
                const float x = w->h_presummed[b] ? o[0] : W[0] * o[0] + W[1] * o[1] + W[2] * o[2], r = w->h_presummed[b] ? o[3] : W[0] * o[3];
                ext = std::max(ext, std::fabs(x) + std::fabs(r));
            }
        }
        w->ext_max = std::max(ext, w->cfg.body_reach_hint);
    }
    // slab worlds keep room for immigrants and ghosts; counts live on the device (P.dn)
    // ghosts per face ~ n * reach / slab width; keep 16x headroom (overflow is detected and reported)
    uint32_t ghost_cap = 0;
    if (slab) {
        float rr = 0.f; for (size_t i = 3; i < w->h_obj.size(); i += 4) rr = std::max(rr, std::fabs(w->h_obj[i]));
        rr = std::max(rr, w->cfg.rm_max_hint);
        if (K) rr = std::max(rr, w->ext_max);      // compound bodies are ghosts as far as they reach
        ghost_cap = w->cfg.halo_ghost_cap ? w->cfg.halo_ghost_cap : rb_halo_ghost_cap_auto(nb, rr, w->cfg.margin_cap, w->cfg.slab_hi - w->cfg.slab_lo);
    }
    w->ghost_cap = ghost_cap;
    const uint32_t cap = slab ? nb + std::max<uint32_t>(nb / 4, 4096u) + 3 * ghost_cap : nb;
    w->cap_local = cap;
    P.nb = cap; P.ns = K ? cap * K : (slab ? cap : ns); P.n_owned = nb;
    w->n_owned_host = nb;
    P.half = w->cfg.world_half; P.gravity = w->cfg.gravity; P.friction = w->cfg.friction;
    P.pushout = w->cfg.pushout; P.ss_damp = w->cfg.ss_damp; P.margin_cap = w->cfg.margin_cap; P.dt = 0.f;

    // ---- spheres: pre-sum the W3x3 * object-centre part of GeometryMath::transform (GeometryMath.cpp:81-89) ----
    std::vector<float> sph((size_t)std::max<uint32_t>(K ? cap * K : (slab ? cap : ns), 1) * 4, 0.f);
    float rmax = 0.f;
    for (uint32_t b = 0; b < nb; b++) {
        const float* W = &w->h_W[(size_t)b * 12];
        for (uint32_t q = w->h_body_start[b]; q < w->h_body_start[b + 1]; q++) {
            const float* o = &w->h_obj[(size_t)q * 4];
            const uint32_t s = K ? b * K + (q - w->h_body_start[b]) : q;      // (fixed stride in compound slab worlds)
            if (w->h_presummed[b]) {      // carried over from the world this one replaces (world_grow): already transformed, bit for bit
                for (int k = 0; k < 4; k++) sph[(size_t)s * 4 + k] = o[k];
                rmax = std::max(rmax, std::fabs(o[3]));
                continue;
            }
            volatile float x = W[0] * o[0] + W[1] * o[1] + W[2] * o[2];   // ((a + b) + c), no contraction on the host build
            volatile float y = W[3] * o[0] + W[4] * o[1] + W[5] * o[2];
            volatile float z = W[6] * o[0] + W[7] * o[1] + W[8] * o[2];
            volatile float r = W[0] * o[3];
            sph[(size_t)s * 4] = x; sph[(size_t)s * 4 + 1] = y; sph[(size_t)s * 4 + 2] = z; sph[(size_t)s * 4 + 3] = r;
            rmax = std::max(rmax, std::fabs((float)r));
        }
    }
    // centred spheres (object centre +0 after W): the world radius rides in pos.w and sph_obj is never read on the hot path
    bool obj_zero = P.single != 0;
    for (size_t i = 0; i < (size_t)ns * 4 && obj_zero; i += 4) {
        uint32_t bx, by, bz; memcpy(&bx, &sph[i], 4); memcpy(&by, &sph[i + 1], 4); memcpy(&bz, &sph[i + 2], 4);
        obj_zero = (bx | by | bz) == 0u;
    }
    P.obj_zero = obj_zero ? 1u : 0u;
    w->slim = P.single && obj_zero && !w->any_wt;
    P.soft = slab ? 0u : 1u;
    if (slab && w->cfg.rm_max_hint > rmax) rmax = w->cfg.rm_max_hint;   // every rank must use the same reach
    if (!(w->cfg.margin_cap > 0.f)) w->cfg.margin_cap = std::max(0.5f * rmax, 1e-3f);   // auto: half the largest radius
    P.margin_cap = w->cfg.margin_cap;
    P.rm_max = rmax + w->cfg.margin_cap;
    if (K) w->ext_max += w->cfg.margin_cap;      // (reach of the farthest sphere grown by the largest margin)
    // persistent candidate lists: single worlds of single-sphere bodies stepped by rb_step with the default kernel
    w->vl_on = w->vl_on && P.single && ns > 0 && w->collide_variant == 2;
    if (w->vl_on) {
        float rmin = rmax;
        for (size_t i = 3; i < sph.size() && i < (size_t)ns * 4; i += 4) rmin = std::min(rmin, std::fabs(sph[i]));
        P.vl_skin = w->vl_skin_cfg > 0.f ? w->vl_skin_cfg : std::max(0.08f * rmin, 4e-3f);   // auto: 8 % of the smallest radius (C3: 0.05 -> 0.1650, 0.035 -> 0.1626 ms per tick)
        P.rm_max += P.vl_skin;               // the largest grown radius any enumeration can meet
        P.vl_slack = std::max(1e-2f, 1e-5f * P.half);   // FP32 rounding of the predicate in sphere-centred coordinates is ~1e-4 at |x| = 1000
    }
    // ---- sphere column grid geometry ----
    {
        float wcol = std::max(2.0f * P.rm_max, 2.0f * P.half / 4096.0f);
        long n = (long)std::floor(2.0 * P.half / wcol);
        n = std::max(1L, std::min(4096L, n));
        P.scz = (uint32_t)n;
        P.s_inv_w = (float)n / (2.0f * P.half);
        P.s_x0 = -P.half;
        P.scx = (uint32_t)n;
        if (slab) {
            // the grid only spans this slab plus a pad for ghosts and drift; spheres beyond clamp into the
            // edge columns (the column function is monotone, so this only costs speed, never correctness)
            float cw = 2.0f * P.half / (float)n;
            float pad = 8.0f * cw;
            float lo = std::max(-P.half, w->cfg.slab_lo - pad), hi = std::min(P.half, w->cfg.slab_hi + pad);
            long c0 = (long)std::floor((lo + P.half) / cw);
            P.s_x0 = -P.half + (float)c0 * cw;
            P.scx = (uint32_t)std::max(1L, std::min((long)n, (long)std::ceil((hi - P.s_x0) / cw)));
        }
        w->ncols = P.scx * P.scz;
        P.g_swap = slab ? 1u : 0u;
        P.g_nu = slab ? P.scz : P.scx; P.g_nv = slab ? P.scx : P.scz;
        P.g_off_u = slab ? P.half : -P.s_x0; P.g_off_v = slab ? -P.s_x0 : P.half;
    }
    // ---- device state ----
    float4 *d_pos, *d_vel, *d_mt, *d_pt, *d_sph, *d_ws, *d_sorted; uint32_t *d_flags, *d_key, *d_rank, *d_cc, *d_cs, *d_sid;
    const size_t nba = slab ? cap : nb, nsa = K ? (size_t)cap * K : (slab ? cap : ns);
    ALLOC(d_pos, nba); ALLOC(d_vel, nba); ALLOC(d_mt, nba); ALLOC(d_pt, nba); ALLOC(d_flags, nba);
    ALLOC(d_sph, nsa); ALLOC(d_ws, nsa); ALLOC(d_key, nsa); ALLOC(d_rank, nsa); ALLOC(d_sorted, nsa); ALLOC(d_sid, nsa);
    if (slab) {
        ALLOC(w->d_dn, 16); ALLOC(w->d_gid, nba);
        uint32_t dn0[16] = { nb, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        CU(cudaMemcpyAsync(w->d_dn, dn0, sizeof dn0, cudaMemcpyHostToDevice, w->stream));
        std::vector<uint32_t> ids(nb);
        for (uint32_t i = 0; i < nb; i++) ids[i] = i < w->h_gid.size() ? w->h_gid[i] : i;
        CU(cudaMemcpyAsync(w->d_gid, ids.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        P.dn = w->d_dn; P.gid = w->d_gid;
    }
    {   // histogram | pad | ticket | status  -- one allocation so that one memset clears all of it every tick
        size_t hist = (((size_t)w->ncols + 1) * 4 + 15) / 16 * 16;
        size_t tiles = rb_scan_tiles(w->ncols) + 1;
        uint8_t* blob; ALLOC(blob, hist + 16 + tiles * 8);
        d_cc = (uint32_t*)blob; w->d_scan_ticket = (uint32_t*)(blob + hist); w->d_scan_status = (unsigned long long*)(blob + hist + 16);
        w->hist_clear_bytes = hist + 16 + tiles * 8;
    }
    ALLOC(d_cs, (size_t)w->ncols + 1);
    ALLOC(w->d_tile_sums, rb_scan_tiles(w->ncols) + 1);
    P.pos = d_pos; P.vel = d_vel; P.mt = d_mt; P.pt = d_pt; P.flags = d_flags; P.sph_obj = d_sph; P.ws = d_ws;
    P.key = d_key; P.rank = d_rank; P.sorted = d_sorted; P.sorted_id = d_sid; P.col_count = d_cc; P.col_start = d_cs;
    std::vector<float> p4((size_t)std::max<uint32_t>(nb, 1) * 4, 0.f), v4 = p4, wt4 = p4;
    for (uint32_t b = 0; b < nb; b++) {
        for (int k = 0; k < 3; k++) { p4[(size_t)b * 4 + k] = w->h_pos[(size_t)b * 3 + k]; v4[(size_t)b * 4 + k] = w->h_vel[(size_t)b * 3 + k]; wt4[(size_t)b * 4 + k] = w->h_W[(size_t)b * 12 + 9 + k]; }
        p4[(size_t)b * 4 + 3] = P.single ? sph[(size_t)b * 4 + 3] : 1.f;      // single-sphere bodies: world radius (GeometryMath.cpp:83-87) in pos.w
    }
    CU(cudaMemcpyAsync(d_pos, p4.data(), (size_t)nb * 16, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d_vel, v4.data(), (size_t)nb * 16, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d_flags, w->h_flags.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d_sph, sph.data(), (K ? (size_t)nb * K : (size_t)ns) * 16, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemsetAsync(d_mt, 0, std::max<size_t>(nba, 1) * 16, w->stream));   // MVP() model = identity -> translation 0
    CU(cudaMemsetAsync(d_pt, 0, std::max<size_t>(nba, 1) * 16, w->stream));
    if (w->any_wt) { float4* d_wt; ALLOC(d_wt, nb); CU(cudaMemcpyAsync(d_wt, wt4.data(), (size_t)nb * 16, cudaMemcpyHostToDevice, w->stream)); P.wt = d_wt; }
    if (K) {
        // fixed-stride compound bodies: spheres per row, every sphere row "not in the broadphase" until K1 says otherwise, commit records
        std::vector<uint32_t> cnt(std::max<uint32_t>(nb, 1), 0u);
        for (uint32_t b = 0; b < nb; b++) cnt[b] = w->h_body_start[b + 1] - w->h_body_start[b];
        uint32_t* d_cnt_rows; ALLOC(d_cnt_rows, nba);
        CU(cudaMemsetAsync(d_cnt_rows, 0, nba * 4, w->stream));
        CU(cudaMemcpyAsync(d_cnt_rows, cnt.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        P.body_cnt = d_cnt_rows;
        CU(cudaMemsetAsync(d_key, 0xff, nsa * 4, w->stream));
        uint32_t* cb; float4 *c1, *c2, *c3, *c4; uint32_t* cf;
        ALLOC(cb, nba); ALLOC(c1, nba); ALLOC(c2, nba); ALLOC(c3, nba); ALLOC(c4, nba); ALLOC(cf, nba);
        P.ch_body = cb; P.ch_pos = c1; P.ch_vel = c2; P.ch_mt = c3; P.ch_pt = c4; P.ch_flags = cf;
    } else if (!P.single) {
        uint32_t *d_sb, *d_bs;
        ALLOC(d_sb, ns); ALLOC(d_bs, (size_t)nb + 1);
        CU(cudaMemcpyAsync(d_sb, w->h_sph_body.data(), (size_t)ns * 4, cudaMemcpyHostToDevice, w->stream));
        CU(cudaMemcpyAsync(d_bs, w->h_body_start.data(), ((size_t)nb + 1) * 4, cudaMemcpyHostToDevice, w->stream));
        P.sph_body = d_sb; P.body_start = d_bs;
        // lanes per body in the group kernel: the next power of two above the largest sphere count, 4 .. 32
        uint32_t kmax = 1; for (uint32_t b = 0; b < nb; b++) kmax = std::max(kmax, w->h_body_start[b + 1] - w->h_body_start[b]);
        P.grp_lg = 2; while ((1u << P.grp_lg) < kmax && P.grp_lg < 5) P.grp_lg++;
        { const char* v = getenv("RB_GROUP_LG"); if (v && *v) P.grp_lg = (uint32_t)std::min(5, std::max(1, atoi(v))); }      // (development switch)
        P.grp_on = compound_kernel_choice();
        uint32_t* cb; float4 *c1, *c2, *c3, *c4; uint32_t* cf;
        ALLOC(cb, nb); ALLOC(c1, nb); ALLOC(c2, nb); ALLOC(c3, nb); ALLOC(c4, nb); ALLOC(cf, nb);
        P.ch_body = cb; P.ch_pos = c1; P.ch_vel = c2; P.ch_mt = c3; P.ch_pt = c4; P.ch_flags = cf;
    }
    if (!P.single && P.grp_on == 2) {      // candidate lists of the split compound kernels (slot-major by sphere row)
        uint32_t *ch, *cg, *cl, *ct;
        ALLOC(ch, nsa); ALLOC(cg, (size_t)nsa * 4); ALLOC(cl, (size_t)nsa * 4); ALLOC(ct, (size_t)nsa * rb_compound_tri_slots());
        CU(cudaMemsetAsync(ch, 0, (size_t)nsa * 4, w->stream));
        P.ce_hdr = ch; P.ce_pg = cg; P.ce_pl = cl; P.ce_tri = ct;
        { uint32_t* cp; ALLOC(cp, nsa); CU(cudaMemsetAsync(cp, 0, (size_t)nsa * 4, w->stream)); P.ce_ph = cp; }
        {   // persistent triangle lists: skin = 8 % of the smallest radius like the single-sphere lists (RB_COMPOUND_SKIN=<m>, 0 = off)
            float rmin = rmax;
            for (size_t i = 3; i < sph.size(); i += 4) if (sph[i] != 0.f) rmin = std::min(rmin, std::fabs(sph[i]));
            P.ce_skin = std::max(0.08f * rmin, 4e-3f);
            const char* v = getenv("RB_COMPOUND_SKIN"); if (v && *v) P.ce_skin = (float)atof(v);
            float4* cb4; ALLOC(cb4, nsa); CU(cudaMemsetAsync(cb4, 0xff, (size_t)nsa * 16, w->stream));      // (NaN: no list yet)
            P.ce_build = cb4;
        }
        uint32_t *sn, *sl, *so; float4 *s1, *s2, *s3, *s4;
        ALLOC(sn, nba); ALLOC(sl, nba); ALLOC(so, nba); ALLOC(s1, nba); ALLOC(s2, nba); ALLOC(s3, nba); ALLOC(s4, nba);
        CU(cudaMemsetAsync(sn, 0, (size_t)nba * 4, w->stream)); CU(cudaMemsetAsync(so, 0, (size_t)nba * 4, w->stream));
        P.ss_npart = sn; P.ss_list = sl; P.ss_slot = so; P.ss_pos = s1; P.ss_vel = s2; P.ss_mt = s3; P.ss_pt = s4;
        // the sphere-sphere phase of the listed bodies runs next to the triangle phase of everybody else
        { const char* v = getenv("RB_COMPOUND_SIDE");      // (development switch)
          if (!(v && *v == '0')) {
              int lo = 0, hi = 0; CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
              CU(cudaStreamCreateWithPriority(&w->vl_side, cudaStreamNonBlocking, hi));
              CU(cudaEventCreateWithFlags(&w->vl_fork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&w->vl_join, cudaEventDisableTiming));
          } }
    }
    if (w->vl_on) {
        float4* wb;
        ALLOC(wb, nsa); CU(cudaMemsetAsync(wb, 0, (size_t)nsa * 16, w->stream));
        CU(cudaMemsetAsync(d_key, 0xff, (size_t)nsa * 4, w->stream));                 // "not in the grid" until the first K1
        CU(cudaHostAlloc((void**)&w->vl_host, 16 * sizeof(uint32_t), cudaHostAllocMapped));
        for (int k = 0; k < 16; k++) w->vl_host[k] = 0;
        w->vl_host[10] = 0xffffffffu;        // (no generation has been flooded yet)
        if (!slab) {
            // single worlds start on the per-tick kernel and switch to the lists once the device reports that few bodies would
            // outrun their skin (a scene that starts at rest gets there within a couple of ticks; one that starts in free fall
            // -- where every tick would rebuild -- never pays for lists it cannot use)
            const char* v = getenv("RB_VL_START_ACTIVE");      // (development switch)
            if (!(v && *v == '1')) { w->vl_active = false; w->vl_host[2] = 0xffffffffu; w->vl_resume_at = 1; }
        }
        uint32_t* dv = nullptr; CU(cudaHostGetDevicePointer((void**)&dv, w->vl_host, 0));
        P.ws_build = wb; P.vl_host = dv;
        {   // the hot bodies' launch is a handful of latency-bound blocks next to a kernel that fills the GPU: highest priority, so
            // that they get their SM slots at once
            int lo = 0, hi = 0; CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            const char* v = getenv("RB_SIDE_PRIO");      // (development switch)
            CU(cudaStreamCreateWithPriority(&w->vl_side, cudaStreamNonBlocking, (v && *v == '0') ? lo : hi));
        }
        CU(cudaEventCreateWithFlags(&w->vl_fork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&w->vl_join, cudaEventDisableTiming));
        // coarse cell table of the hot bodies: cells wide enough that any reach box touches at most 2 x 2 of them
        // (as fine as that allows, up to 2048^2 cells: the chains every row walks stay short even when most bodies are hot)
        const float lc = std::max(4.0f * P.rm_max, 2.0f * P.half / 2048.0f);
        P.hcn = (uint32_t)std::max(1L, (long)std::floor(2.0 * P.half / lc));
        P.hc_inv_w = (float)P.hcn / (2.0f * P.half);
        // every row has a slot (the table cannot overflow); hot_soft is the count at which rebuilding the lists pays
        P.hot_cap = (uint32_t)nsa;
        P.hot_soft = w->vl_hot_cfg > 0 ? (uint32_t)w->vl_hot_cfg : std::max<uint32_t>(1024u, ns / 128u);
        P.hot_flood = std::max<uint32_t>(2u * P.hot_soft, ns / 8u);
        unsigned long long* hh; uint32_t *hr, *hn;
        const size_t ncell = (size_t)P.hcn * P.hcn;
        ALLOC(hh, ncell); ALLOC(hr, (size_t)P.hot_cap + 3u * ghost_cap); ALLOC(hn, (size_t)P.hot_cap + 3u * ghost_cap);   // slab worlds: + this tick's ghosts
        CU(cudaMemsetAsync(hh, 0, ncell * sizeof(unsigned long long), w->stream));      // tick 0 never runs: every cell starts empty, and stays versioned
        P.hot_head = hh; P.hot_row = hr; P.hot_next = hn;
        // ring of per-tick counter blocks {contacts lo, hi, hot bodies, -} and, right behind it, the list words (rebuild flag first)
        ALLOC(w->d_tb, 4 * RB_TB_BLOCKS + RB_VL_WORDS);
        CU(cudaMemsetAsync(w->d_tb, 0, (4 * RB_TB_BLOCKS + RB_VL_WORDS) * sizeof(uint32_t), w->stream));
        P.vl = w->d_tb + 4 * RB_TB_BLOCKS;
        P.scan_status = w->d_scan_status; P.scan_ticket = w->d_scan_ticket; P.ncols = w->ncols;      // for the device-launched rebuild chain
        static_assert(RB_VL_REBUILD == 0, "the rebuild flag must be the first list word");
        // candidate lists: header + partners per row, triangles in per-block tiles (sized for 3 per row on average; a block
        // that does not fit steps without lists)
        uint32_t *eh, *eg, *el, *tc; uint4* tp; uint2* bt;
        const size_t nblk = ((size_t)P.ns + rb_list_tile_rows() - 1) / rb_list_tile_rows();
        ALLOC(eh, P.ns); ALLOC(eg, (size_t)P.ns * rb_list_partner_slots()); ALLOC(el, (size_t)P.ns * rb_list_partner_slots());
        P.tile_pool_units = (uint32_t)std::min<size_t>(0xF0000000ull, (size_t)P.ns * 10 + 4096);
        ALLOC(tp, P.tile_pool_units); ALLOC(bt, nblk); ALLOC(tc, 4);
        CU(cudaMemsetAsync(eh, 0, (size_t)P.ns * 4, w->stream)); CU(cudaMemsetAsync(bt, 0, nblk * sizeof(uint2), w->stream)); CU(cudaMemsetAsync(tc, 0, 16, w->stream));
        P.e_hdr = eh; P.e_pg = eg; P.e_pl = el; P.tile_pool = tp; P.blk_tile = bt; P.tile_cursor = tc;
        if (slab) {
            // RB_HALO_SPLIT=1: the list step (and the hot bodies' launch) split around the halo exchange -- the blocks no ghost can
            // reach step before the wait for the neighbours. Measured at N=2 (C3 weak): the wait disappears (halo stage 25 -> 3 us)
            // but the late hot-body launch, a handful of latency-bound warps that used to hide behind the full list step, is now
            // exposed (35 us): 0.227 against 0.201 ms per tick. Off by default. (tick numbers as versions: never cleared)
            const char* v = getenv("RB_HALO_SPLIT");
            if (v && *v == '1') {
                uint32_t *bb, *bd; ALLOC(bb, nblk); ALLOC(bd, nblk);
                CU(cudaMemsetAsync(bb, 0, nblk * 4, w->stream)); CU(cudaMemsetAsync(bd, 0, nblk * 4, w->stream));
                P.blk_bnd = bb; P.blk_done = bd;
            }
        }
    }
    {   // static 3x3 of every _worldSpaceTransform, kept only when some body is not identity (model-matrix export)
        bool any = false;
        for (uint32_t b = 0; b < nb && !any; b++) { const float* W = &w->h_W[(size_t)b * 12]; any = !(W[0] == 1.f && W[1] == 0.f && W[2] == 0.f && W[3] == 0.f && W[4] == 1.f && W[5] == 0.f && W[6] == 0.f && W[7] == 0.f && W[8] == 1.f); }
        if (any) {
            std::vector<float> w3((size_t)nb * 12, 0.f);
            for (uint32_t b = 0; b < nb; b++) for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) w3[(size_t)b * 12 + r * 4 + c] = w->h_W[(size_t)b * 12 + r * 3 + c];
            ALLOC(w->d_w3, (size_t)nb * 3);
            CU(cudaMemcpyAsync(w->d_w3, w3.data(), w3.size() * sizeof(float), cudaMemcpyHostToDevice, w->stream));
            CU(cudaStreamSynchronize(w->stream));
            w->any_w3 = true;
        }
    }
    // ---- contacts + counters ----
    P.contact_cap = w->cfg.max_contacts ? w->cfg.max_contacts : (unsigned long long)(K ? cap * K : (slab ? cap : ns)) * 4 + 1024;
    RbContactRec* d_c; ALLOC(d_c, (size_t)P.contact_cap); P.contacts = d_c;
    unsigned long long* d_cnt; ALLOC(d_cnt, RB_C_COUNT * (RB_STRIPES + 1)); P.counters = d_cnt; P.stripes = d_cnt + RB_C_COUNT;
    P.ncontact = d_cnt + RB_C_NCONTACT; w->ncontact_dev = P.ncontact; P.mt_in_valid = 1u;
    CU(cudaMemsetAsync(d_cnt, 0, RB_C_COUNT * (RB_STRIPES + 1) * sizeof(unsigned long long), w->stream));
    CU(cudaMallocHost((void**)&w->h_counters, (RB_C_COUNT * (RB_STRIPES + 1) + 1) * sizeof(unsigned long long)));
    memset(w->h_counters, 0, (RB_C_COUNT * (RB_STRIPES + 1) + 1) * sizeof(unsigned long long));
    for (auto& e : w->ev) CU(cudaEventCreate(&e));
    CU(cudaStreamSynchronize(w->stream));
    int r = build_triangle_grid(w);
    if (r) return r;
    // slab worlds whose bodies take forces / torques / angular state: the arrays exist from the start (a migrant arriving later
    // must find them), sized for every row this rank can ever hold
    if (slab && w->cfg.slab_forces) { r = ensure_angular(w); if (r) return r; }
    // ---- prime the model transforms: Entity::_updateKinematics(0) on every entity (SURVEY Appendix A;
    // note StateVector::update(0) still applies the friction factor to bodies in contact / without gravity) ----
    P.dt = 0.0f;
    rb_launch_integrate_bin(P, 1, 0, w->stream); w->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(w->stream));
    w->built = true;
    if (P.single && !slab && nb > 0 && w->reorder_every > 0) {
        // one broadphase over the initial poses, then put the body rows in that spatial order
        int rr = enqueue_broadphase(w, 0); if (rr) return rr;
        rr = reorder_rows(w); if (rr) return rr;
        CU(cudaStreamSynchronize(w->stream));
    }
    // host staging is no longer needed
    std::vector<float>().swap(w->h_tri); std::vector<float>().swap(w->h_obj); std::vector<float>().swap(w->h_W);
    std::vector<float>().swap(w->h_pos); std::vector<float>().swap(w->h_vel); std::vector<uint8_t>().swap(w->h_presummed);
    return RB_OK;
}

// ---- the step ---------------------------------------------------------------------------------
int enqueue_broadphase_tail(rb_world* w) {
    RbParams& P = w->P;
    if (w->profile) CU(cudaEventRecord(w->ev[1], w->stream));
    if (w->scan_variant == 1) w->launches += rb_launch_scan(P.col_count, w->ncols, P.col_start, w->d_tile_sums, w->stream);
    else w->launches += rb_launch_scan_lookback(P.col_count, w->ncols, P.col_start, w->d_scan_status, w->d_scan_ticket, w->stream);
    if (w->profile) CU(cudaEventRecord(w->ev[2], w->stream));
    rb_launch_scatter(P, w->stream); w->launches++;
    if (w->profile) CU(cudaEventRecord(w->ev[3], w->stream));
    return RB_OK;
}
static int enqueue_broadphase(rb_world* w, int do_integrate) {
    RbParams& P = w->P;
    if (w->dist) return rb_dist_step_exchange(w, do_integrate);
    { int r = rb_plain_tick_params(w); if (r) return r; }
    rb_launch_clear(P.col_count, w->hist_clear_bytes, nullptr, 0, w->stream); w->launches++;      // histogram + scan ticket / status
    if (w->profile) CU(cudaEventRecord(w->ev[0], w->stream));
    rb_launch_integrate_bin(P, do_integrate, 1, w->stream); w->launches++;
    return enqueue_broadphase_tail(w);
}
static int enqueue_collide(rb_world* w, bool resolve) {
    RbParams& P = w->P;
    // per-step counters: contact count + changed count
    rb_launch_clear(nullptr, 0, (uint32_t*)(P.counters + RB_C_NCONTACT), 6, w->stream); w->launches++;      // contacts, changed bodies, sphere-sphere list
    w->launches += rb_launch_collide(P, resolve, w->collide_variant, w->stream, w->vl_side, w->vl_fork, w->vl_join);
    if (w->profile) {
        CU(cudaEventRecord(w->ev[4], w->stream));
        CU(cudaEventSynchronize(w->ev[4]));
        for (int k = 0; k < 4; k++) { float ms = 0; CU(cudaEventElapsedTime(&ms, w->ev[k], w->ev[k + 1])); w->kernel_ms[k] += ms; }
        if (w->dist) {                                  // slab worlds: split K1 from the halo step (signal, wait, remove, unpack) that follows it
            float ms = 0; CU(cudaEventElapsedTime(&ms, w->ev[5], w->ev[1]));
            w->kernel_ms[0] -= ms; w->kernel_ms[4] += ms;
        }
    }
    CU(cudaGetLastError());
    return RB_OK;
}

// One tick with persistent candidate lists. K1 integrates and checks, per body, whether the sphere is still inside the
// skin its list was built with; a body that is not steps without its list this tick (rb_collide_single_kernel<3>, a small
// launch next to the list kernel). The lists are rebuilt (bin -> scan -> scatter -> enumerate)
//   * single worlds ("soft"): when the HOST says so -- every vl_period list ticks, a period it adapts to the hot-body counts
//     the device reports through mapped memory -- so a reuse tick is three launches (K1, hot bodies, list step) and
//     nothing on the device ever waits for a decision;
//   * slab worlds: when the device raises the rebuild flag (a row moved between ranks, too many hot bodies) or the host
//     forces it; the chain kernels follow as launches that return at once while the flag is down, or are tail-launched by
//     a one-thread gate kernel (library built with device-side launches).
// Results are the ones of the per-tick enumeration: the lists are supersets of it and every decision is taken by the same
// exact tests in the same order.
// (second half of a list tick: slab worlds get here after their halo exchange; K1 ran before it)
static uint32_t list_hot_rows(rb_world* w);
uint32_t rb_list_hot_rows(rb_world* w) { return list_hot_rows(w); }
static uint32_t list_hot_rows(rb_world* w) {
    // grid of the hot bodies' launch: its blocks stride over the table, so any size is correct. Enough blocks to fill the GPU
    // whatever the count turns out to be (the host's view of it is many ticks old); a block without work returns at once
    static const uint32_t blocks = [] { const char* v = getenv("RB_HOT_BLOCKS"); return v && *v ? (uint32_t)atoi(v) : 148u * 3u; }();      // (development switch)
    return std::min<uint32_t>(std::max<uint32_t>(w->P.ns, 1u), blocks * 256u);
}
int rb_list_tick_tail(rb_world* w, bool reorder_due) {
    RbParams& P = w->P;
    const bool rebuild_now = w->vl_tick_rebuilds;           // the host forced (or, single worlds, scheduled) a rebuild this tick
    // slab worlds built with device-side launches: a one-thread gate kernel tail-launches bin -> scan -> scatter -> enumerate
    // when the rebuild flag is up (rb_kernels.cu), so the host only enqueues the chain itself on the ticks it forces
    const bool chain = P.soft ? rebuild_now : (!w->vl_cdp || rebuild_now || !w->vl_gate_ok);
    const bool gate = !P.soft && !chain;
    if (gate) rb_launch_vl_gate(P, w->stream);
    if (chain) rb_launch_vl_bin(P, w->stream);
    if (w->profile) CU(cudaEventRecord(w->ev[1], w->stream));
    if (chain) rb_launch_vl_scan(P, w->ncols, w->d_scan_status, w->d_scan_ticket, w->stream);
    if (w->profile) CU(cudaEventRecord(w->ev[2], w->stream));
    if (chain) rb_launch_vl_scatter(P, w->ncols, w->d_scan_status, w->d_scan_ticket, w->stream);
    if (reorder_due) { int r = reorder_rows(w, true); if (r) return r; }     // rows into sort order before the lists are built for them
    if (w->profile) CU(cudaEventRecord(w->ev[3], w->stream));
    w->launches += (chain ? 3 : 0) + (gate ? 1 : 0) + 1 /* K1 */ + rb_launch_vl_collide(P, w->stream, w->vl_side, w->vl_fork, w->vl_join, chain, list_hot_rows(w), w->halo_split ? 2 : 0);
    w->halo_split = false;
    if (w->profile) {
        CU(cudaEventRecord(w->ev[4], w->stream));
        CU(cudaEventSynchronize(w->ev[4]));
        for (int k = 0; k < 4; k++) { float ms = 0; CU(cudaEventElapsedTime(&ms, w->ev[k], w->ev[k + 1])); w->kernel_ms[k] += ms; }
        if (w->dist) {                                  // slab worlds: split K1 from the halo step that follows it
            float ms = 0; CU(cudaEventElapsedTime(&ms, w->ev[5], w->ev[1]));
            w->kernel_ms[0] -= ms; w->kernel_ms[4] += ms;
        }
    }
    CU(cudaGetLastError());
    w->vl_force = false;
    return RB_OK;
}
// single worlds: digest the device's latest report {tick, generation of lists, their age, the age at which more than hot_soft
// bodies were hot} and set the rebuild period from it. The report is up to ~32 ticks old (the host runs ahead).
//   * a generation that expired at age a: rebuild every a list ticks (the tick that would have been the expensive one rebuilds);
//     a = 1 says lists do not pay at all right now (free fall: everybody outruns the skin at once) -- rb_step_will_use_lists
//     sits such phases out on the per-tick kernel
//   * a generation that was rebuilt before it expired: the period was shorter than it had to be, lengthen it
static void list_schedule_digest(rb_world* w) {
    const uint32_t t1 = w->vl_host[4], expiry = w->vl_host[5], gen = w->vl_host[6], prev_life = w->vl_host[8], prev_expiry = w->vl_host[9], t2 = w->vl_host[4];
    if (t1 != t2 || t1 == w->vl_seen_tick) return;
    w->vl_seen_tick = t1;
    static const uint32_t max_period = [] { const char* v = getenv("RB_VL_MAXPERIOD"); return v && *v ? (uint32_t)atoi(v) : 64u; }();      // (development switch)
    struct Clamp { rb_world* w; uint32_t m; ~Clamp() { if (w->vl_period > m) w->vl_period = m; } } clamp{ w, max_period };
    // (an expiry after one or two ticks is not a reason to rebuild every tick -- a rebuild costs more than a tick of per-tick
    // enumeration: such lists serve a few ticks with a growing share of hot bodies; a FLOOD is what switches lists off)
    if (expiry && gen != w->vl_gen_expired) { w->vl_gen_expired = gen; w->vl_period = std::max<uint32_t>(4u, std::min<uint32_t>(expiry, 64u)); }
    if (w->vl_host[10] == gen && gen != w->vl_gen_flooded) { w->vl_gen_flooded = gen; w->vl_period = 1u; }
    if (gen != w->vl_gen_seen) {
        // a new generation: if the one before served (nearly) the whole period without expiring, the period can grow
        // (one cut short by a forced rebuild -- a row re-ordering, a host edit -- says nothing)
        if (w->vl_gen_seen != 0xffffffffu && !prev_expiry && !expiry && 4u * (prev_life + 1u) >= 3u * std::min(w->vl_period, w->vl_spacing))
            w->vl_period = std::min<uint32_t>(64u, 2u * w->vl_period);      // (too long shows as an expiry, which sets the period outright)
        w->vl_gen_seen = gen;
    }
}
// A list tick starts: does the host make it a rebuild tick? (forced by other activity since the last list tick, a row
// re-ordering, or -- single worlds -- the schedule)
bool rb_list_tick_begin(rb_world* w, bool reorder_due) {
    bool force = w->vl_force || reorder_due;
    if (w->P.soft && !force) {
        list_schedule_digest(w);
        // the row re-ordering rebuilds every reorder_every ticks anyway: the scheduled rebuilds split that interval evenly
        // (a period of 22 with re-orderings every 32 ticks means 16 + 16, not 22 + 10)
        uint32_t spacing = w->vl_period;
        if (w->reorder_every > 0) { const uint32_t re = (uint32_t)w->reorder_every, n = (re + spacing - 1u) / spacing; spacing = (re + n - 1u) / n; }
        w->vl_spacing = spacing;
        if (w->vl_since + 1 >= spacing) force = true;
    }
    w->vl_tick_rebuilds = force;
    return force;
}
// ... and its K1 has been enqueued
void rb_list_tick_k1_enqueued(rb_world* w) {
    if (w->slim) w->mt_stale = true;                     // from here on Entity::_mvp's translation is implicit (0 + p)
    if (w->vl_tick_rebuilds) w->vl_since = 0;
    else w->vl_since++;
}
static int enqueue_list_tick(rb_world* w, bool reorder_due) {
    RbParams& P = w->P;
    const bool force = rb_list_tick_begin(w, reorder_due);
    if (w->dist) { int r = rb_dist_list_exchange(w, force ? 3 : 2); if (r) return r; }      // K1 (+ pack), exchange, remove + unpack
    else {
        if (w->vl_force) { rb_launch_clear(P.col_count, w->hist_clear_bytes, nullptr, 0, w->stream); w->launches++; }   // other paths may have left counts behind (list ticks leave the histogram clean)
        rb_list_tick_params(w);
        if (!P.soft) { rb_launch_clear(w->d_tb + 4 * (RB_TB_BLOCKS - 1), 20, nullptr, 0, w->stream); w->launches++; }
        if (w->profile) CU(cudaEventRecord(w->ev[0], w->stream));
        rb_launch_integrate_bin(P, 1, force ? 3 : 2, w->stream);
    }
    rb_list_tick_k1_enqueued(w);
    return rb_list_tick_tail(w, reorder_due);
}

// Will the next rb_step run from the persistent lists? (decided once per step)
// The device reports {rebuild ticks, list ticks, hot bodies} through mapped host memory; the host only looks at it (never
// waits for it) to notice phases in which nearly every tick would rebuild (free fall: every body outruns its skin
// within a tick or two), and sits those out on the per-tick kernel. Both paths give the same results, so the switch is
// invisible.
bool rb_step_will_use_lists(rb_world* w) {
    if (!w->vl_on) return false;
    if (w->vl_decided_for == w->steps + 1) return w->vl_active;
    w->vl_decided_for = w->steps + 1;
    if (w->vl_active) {
        bool mostly_rebuilding = false, judged = false;
        if (w->P.soft) {
            // the schedule has converged on rebuilding (nearly) every tick
            // (a period of 1 is the schedule's verdict "more hot bodies after one tick than a rebuild is worth": act on it at once)
            if (w->vl_period < 2u || w->tick_seq - w->vl_judged_at >= 16u) { judged = true; w->vl_judged_at = w->tick_seq; mostly_rebuilding = w->vl_period < 3u; }
        } else {
            const uint32_t nr = w->vl_host[0], nt = w->vl_host[1];
            if (nt - w->vl_seen_t >= 16u) {
                judged = true;
                mostly_rebuilding = 2u * (nr - w->vl_seen_r) > (nt - w->vl_seen_t);
                w->vl_gate_ok = 8u * (nr - w->vl_seen_r) <= (nt - w->vl_seen_t);     // device-launched rebuilds pay off while rebuilds are rare
                w->vl_seen_r = nr; w->vl_seen_t = nt;
            }
        }
        if (judged && w->vl_never_off) mostly_rebuilding = false;
        if (judged) {
            if (mostly_rebuilding) { w->vl_active = false; w->vl_host[2] = 0xffffffffu; w->vl_resume_at = w->steps + w->vl_backoff; w->vl_backoff = std::min<uint32_t>(w->vl_backoff * 2u, 256u); }
            else w->vl_backoff = 32;
        }
    } else if (w->steps >= w->vl_resume_at) {
        // come back once few enough bodies outrun the skin (the per-tick kernel keeps counting them)
        if (w->vl_host[2] < w->P.hot_flood / 2u) {
            w->vl_active = true; w->vl_force = true; w->vl_seen_r = w->vl_host[0]; w->vl_seen_t = w->vl_host[1];
            w->vl_period = 8; w->vl_judged_at = w->tick_seq;
        }
        else w->vl_resume_at = w->steps + (w->vl_host[2] == 0xffffffffu ? 2 : 16);      // (no report yet: look again soon)
    }
    return w->vl_active;
}

int rb_step(rb_world* w, int ms) {
    if (!w || !w->built) return w ? fail(w, RB_ERR_INVALID, "rb_step before rb_build") : RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    w->P.dt = (float)ms / 1000.0f;                     // StateVector.cpp:19
    const bool reorder_due = w->reorder_every > 0 && (w->steps + 1) % (uint64_t)w->reorder_every == 0;
    int r;
    rb_step_will_use_lists(w);
    if (w->vl_on && w->vl_active) { r = enqueue_list_tick(w, reorder_due); if (r) return r; }      // a re-ordering tick rebuilds: it needs a fresh sort
    else {
        if (w->vl_on) CU(cudaMemsetAsync(w->P.vl + RB_VL_NFAST, 0, 4, w->stream));
        r = enqueue_broadphase(w, 1); if (r) return r;
        r = enqueue_collide(w, true); if (r) return r;
        w->vl_force = true;
    }
    w->steps++; w->pose_updates++;
    if (reorder_due && !(w->vl_on && w->vl_active)) { r = reorder_rows(w); if (r) return r; }      // (list ticks re-order inside the tick)
    if (w->vl_on && (w->steps & 7u) == 0) {
        // bounded run-ahead: the use-lists decision above feeds on what the device reports, so a host that enqueues
        // hundreds of ticks ahead would decide blind. Every 8 ticks wait for the tick enqueued 8 ticks earlier:
        // the device always has 8+ ticks queued (no bubble), the host is never more than 16 ahead.
        const int k = (int)((w->steps >> 3) & 1u);
        if (!w->vl_pace[0]) { CU(cudaEventCreateWithFlags(&w->vl_pace[0], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&w->vl_pace[1], cudaEventDisableTiming)); }
        if (w->d_dn) {      // slab worlds: the sticky overflow counters ride along (a lost body / missed ghost must not go unnoticed)
            if (!w->h_sticky) { CU(cudaMallocHost((void**)&w->h_sticky, 4 * sizeof(unsigned long long))); memset(w->h_sticky, 0, 4 * sizeof(unsigned long long)); }
            CU(cudaMemcpyAsync(w->h_sticky + 2 * k, w->P.counters + RB_C_IMM_DROPPED, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
        }
        CU(cudaEventRecord(w->vl_pace[k], w->stream));
        if (w->vl_pace_armed[k ^ 1]) {
            CU(cudaEventSynchronize(w->vl_pace[k ^ 1]));
            if (w->h_sticky) { int e = rb_dist_sticky_error(w, w->h_sticky[2 * (k ^ 1)], w->h_sticky[2 * (k ^ 1) + 1]); if (e) return e; }
        }
        w->vl_pace_armed[k] = true;
    }
    return RB_OK;
}
int rb_integrate(rb_world* w, int ms) {
    if (!w || !w->built) return w ? fail(w, RB_ERR_INVALID, "rb_integrate before rb_build") : RB_ERR_INVALID;
    w->vl_force = true;
    CU(cudaSetDevice(w->cfg.device));
    w->P.dt = (float)ms / 1000.0f;
    { int r = rb_plain_tick_params(w); if (r) return r; }
    rb_launch_integrate_bin(w->P, 1, 0, w->stream); w->launches++;
    w->pose_updates++;
    CU(cudaGetLastError());
    return RB_OK;
}
int rb_collide(rb_world* w) {
    if (!w || !w->built) return w ? fail(w, RB_ERR_INVALID, "rb_collide before rb_build") : RB_ERR_INVALID;
    w->vl_force = true;
    CU(cudaSetDevice(w->cfg.device));
    int r = enqueue_broadphase(w, 0); if (r) return r;
    return enqueue_collide(w, true);
}
int rb_detect(rb_world* w) {
    if (!w || !w->built) return w ? fail(w, RB_ERR_INVALID, "rb_detect before rb_build") : RB_ERR_INVALID;
    w->vl_force = true;
    CU(cudaSetDevice(w->cfg.device));
    int r = enqueue_broadphase(w, 0); if (r) return r;
    return enqueue_collide(w, false);
}
int rb_sync(rb_world* w) {
    if (!w) return RB_ERR_INVALID;
    CU(cudaStreamSynchronize(w->stream));
    if (w->s_d2h) { CU(cudaStreamSynchronize(w->s_h2d)); CU(cudaStreamSynchronize(w->s_d2h)); }
    return RB_OK;
}

static int fetch_counters(rb_world* w) {
    CU(cudaMemcpyAsync(w->h_counters, w->P.counters, RB_C_COUNT * (RB_STRIPES + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    if (w->ncontact_dev && w->ncontact_dev != w->P.counters + RB_C_NCONTACT)      // list ticks count their contacts in the tick's own counter block
        CU(cudaMemcpyAsync(w->h_counters + RB_C_COUNT * (RB_STRIPES + 1), w->ncontact_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    if (w->ncontact_dev && w->ncontact_dev != w->P.counters + RB_C_NCONTACT) w->h_counters[RB_C_NCONTACT] = w->h_counters[RB_C_COUNT * (RB_STRIPES + 1)];
    for (int k = 0; k <= RB_C_OVERFLOW; k++) {          // cumulative test / hit counters live in the stripes
        unsigned long long sum = w->h_counters[k];      // base copy (halo capacity overflow is counted there)
        for (int st = 0; st < RB_STRIPES; st++) sum += w->h_counters[(size_t)(st + 1) * RB_C_COUNT + k];
        w->h_counters[k] = sum;
    }
    w->contacts_last = w->h_counters[RB_C_NCONTACT];
    return RB_OK;
}

int rb_contact_count(rb_world* w, uint64_t* n) {
    if (!w || !w->built || !n) return RB_ERR_INVALID;
    int r = fetch_counters(w); if (r) return r;
    *n = w->contacts_last;
    return RB_OK;
}

int rb_query_contacts(rb_world* w, rb_contact* out, uint64_t cap, uint64_t* n, int sorted) {
    if (!w || !w->built || !n) return w ? fail(w, RB_ERR_INVALID, "rb_query_contacts: bad argument") : RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    int r = fetch_counters(w); if (r) return r;
    uint64_t total = w->contacts_last;
    uint64_t have = std::min<uint64_t>(total, w->P.contact_cap);
    *n = have;
    if (total > w->P.contact_cap)
        return fail(w, RB_ERR_CAPACITY, "device contact buffer overflow: %llu contacts, capacity %llu (set rb_config.max_contacts)", (unsigned long long)total, (unsigned long long)w->P.contact_cap);
    if (!out) return RB_OK;
    if (cap < have) return fail(w, RB_ERR_CAPACITY, "rb_query_contacts: need room for %llu contacts", (unsigned long long)have);
    std::vector<RbContactRec> rec(have);
    if (have) CU(cudaMemcpy(rec.data(), w->P.contacts, have * sizeof(RbContactRec), cudaMemcpyDeviceToHost));
    const std::vector<uint32_t>& bs = w->h_body_start; const std::vector<uint32_t>& gs = w->h_geom_start;
    auto body_of = [&](uint32_t s, uint32_t& body, uint32_t& idx) {
        if (w->P.single) { body = s; idx = 0; return; }
        if (w->P.kstride) { body = s / w->P.kstride; idx = s % w->P.kstride; return; }      // slab worlds: global body id * stride + sphere
        body = w->h_sph_body[s]; idx = s - bs[body];
    };
    for (uint64_t i = 0; i < have; i++) {
        const RbContactRec& c = rec[i]; rb_contact& o = out[i];
        body_of(c.a, o.body, o.sphere);
        if (c.b & RB_TRI_BIT) {
            uint32_t t = c.b & ~RB_TRI_BIT;
            uint32_t g = (uint32_t)(std::upper_bound(gs.begin(), gs.end(), t) - gs.begin()) - 1;
            o.kind = 1; o.other = g; o.other_prim = t - gs[g];
        } else { o.kind = 0; body_of(c.b, o.other, o.other_prim); }
        o.depth = c.depth; o.normal[0] = c.nx; o.normal[1] = c.ny; o.normal[2] = c.nz;
    }
    if (sorted) {
        std::sort(out, out + have, [](const rb_contact& a, const rb_contact& b) {
            if (a.kind != b.kind) return a.kind < b.kind;
            if (a.body != b.body) return a.body < b.body;
            if (a.sphere != b.sphere) return a.sphere < b.sphere;
            if (a.other != b.other) return a.other < b.other;
            return a.other_prim < b.other_prim;
        });
    }
    return RB_OK;
}

int rb_get_state(rb_world* w, float* pos4, float* vel4, uint32_t* flags) {
    if (!w || !w->built) return RB_ERR_INVALID;
    if (w->d_dn) { int r = rb_dist_refresh_counts(w); if (r) return r; }
    size_t nb = w->n_owned_host;
    if (pos4) { int r = read_rows_f4(w, w->P.pos, pos4, nb, w->P.single ? 1 : 0); if (r) return r; }
    if (vel4) { int r = read_rows_f4(w, w->P.vel, vel4, nb); if (r) return r; }
    if (flags) { int r = read_rows_u32(w, w->P.flags, flags, nb); if (r) return r; }
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}
int rb_get_model_translations(rb_world* w, float* model_t4, float* prev_t4) {
    if (!w || !w->built) return RB_ERR_INVALID;
    if (w->d_dn) { int r = rb_dist_refresh_counts(w); if (r) return r; }
    { int r = rb_ensure_mt(w); if (r) return r; }
    size_t nb = w->n_owned_host;
    if (model_t4) { int r = read_rows_f4(w, w->P.mt, model_t4, nb); if (r) return r; }
    if (prev_t4) { int r = read_rows_f4(w, w->P.pt, prev_t4, nb); if (r) return r; }
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}
int rb_export_model_matrices(rb_world* w, const float** cur_dev, const float** prev_dev) {
    if (!w || !w->built) return RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    // slab (multi-GPU) worlds: the rows this rank owns right now, in the order of rb_get_state / rb_get_body_ids (bodies with a
    // scaled or rotated W keep their 3x3 in build order, which migration does not carry: identity W only)
    if (w->d_dn) { int r = rb_dist_refresh_counts(w); if (r) return r; if (w->d_w3) return fail(w, RB_ERR_INVALID, "model-matrix export on slab worlds needs W = identity"); }
    const size_t nb = w->n_owned_host;
    const size_t cap = w->d_dn ? w->cap_local : nb;
    if (!w->d_mm_cur) { ALLOC(w->d_mm_cur, std::max<size_t>(cap, 1) * 4); ALLOC(w->d_mm_prev, std::max<size_t>(cap, 1) * 4); }
    { int r = rb_ensure_mt(w); if (r) return r; }
    rb_launch_model_matrices(w->d_w3, (w->d_inv && !w->d_dn) ? w->P.gid : nullptr, w->P.mt, w->P.pt, w->d_mm_cur, w->d_mm_prev, (uint32_t)nb, w->pose_updates == 0 ? 1 : 0, w->stream); w->launches++;
    CU(cudaGetLastError());
    if (cur_dev) *cur_dev = (const float*)w->d_mm_cur;
    if (prev_dev) *prev_dev = (const float*)w->d_mm_prev;
    return RB_OK;
}
int rb_get_model_matrices(rb_world* w, float* model16, float* prev16) {
    int r = rb_export_model_matrices(w, nullptr, nullptr); if (r) return r;
    const size_t nb = w->n_owned_host;
    if (model16) CU(cudaMemcpyAsync(model16, w->d_mm_cur, nb * 64, cudaMemcpyDeviceToHost, w->stream));
    if (prev16) CU(cudaMemcpyAsync(prev16, w->d_mm_prev, nb * 64, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}
int rb_get_angular(rb_world* w, float* ap4, float* av4) {
    if (!w || !w->built) return RB_ERR_INVALID;
    if (w->d_dn) { int r = rb_dist_refresh_counts(w); if (r) return r; }      // slab worlds: the rows this rank owns now
    size_t nb = w->n_owned_host;
    if (!w->P.avel) { if (ap4) memset(ap4, 0, nb * 16); if (av4) memset(av4, 0, nb * 16); return RB_OK; }
    if (ap4) { int r = read_rows_f4(w, w->P.apos, ap4, nb); if (r) return r; }
    if (av4) { int r = read_rows_f4(w, w->P.avel, av4, nb); if (r) return r; }
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}
int rb_get_world_spheres(rb_world* w, float* xyzr4) {
    if (!w || !w->built || !xyzr4) return RB_ERR_INVALID;
    if (w->P.kstride) return fail(w, RB_ERR_INVALID, "rb_get_world_spheres is not available on slab worlds of compound bodies");
    w->vl_force = true;
    // recompute from the current model translations (no integration, no binning side effects on state)
    { int r = rb_plain_tick_params(w); if (r) return r; }
    CU(cudaMemsetAsync(w->P.col_count, 0, w->hist_clear_bytes, w->stream));
    { RbParams Q = w->P; Q.dp = nullptr;      // (slab worlds: no halo records, no signal -- this is not a tick)
      rb_launch_integrate_bin(Q, 0, 1, w->stream); w->launches++; }
    size_t ns = w->d_dn ? w->n_owned_host : w->P.ns;
    std::vector<float> tmp(ns * 4), obj(ns * 4);
    if (w->P.single) {
        int r = read_rows_f4(w, w->P.ws, tmp.data(), ns); if (r) return r;
        r = read_rows_f4(w, w->P.sph_obj, obj.data(), ns); if (r) return r;
    } else {
        CU(cudaMemcpyAsync(tmp.data(), w->P.ws, ns * 16, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaMemcpyAsync(obj.data(), w->P.sph_obj, ns * 16, cudaMemcpyDeviceToHost, w->stream));
    }
    CU(cudaStreamSynchronize(w->stream));
    for (size_t i = 0; i < ns; i++) { xyzr4[4 * i] = tmp[4 * i]; xyzr4[4 * i + 1] = tmp[4 * i + 1]; xyzr4[4 * i + 2] = tmp[4 * i + 2]; xyzr4[4 * i + 3] = obj[4 * i + 3]; }
    return RB_OK;
}

static int range_ok(rb_world* w, uint32_t first, uint32_t count) {
    if (!w || !w->built) return RB_ERR_INVALID;
    if ((uint64_t)first + count > w->n_owned_host) return fail(w, RB_ERR_INVALID, "body range [%u, %u) out of bounds (%u bodies)", first, first + count, w->n_owned_host);
    return RB_OK;
}
int rb_set_state(rb_world* w, uint32_t first, uint32_t count, const float* pos4, const float* vel4) {
    int r = range_ok(w, first, count); if (r) return r;
    w->vl_force = true;
    for (const float* a : { pos4, vel4 })      // NaN / Inf are rejected like at rb_add_model (the w lane is padding)
        if (a) for (size_t i = 0; i < count; i++) if (!finite_all(a + 4 * i, 3)) return fail(w, RB_ERR_INVALID, "rb_set_state: non-finite value for body %u", first + (uint32_t)i);
    r = rb_ensure_mt(w); if (r) return r;       // Entity::_mvp keeps the pose of the last kinematics update, not the edited one
    if (pos4) { r = write_rows_f4(w, w->P.pos, first, count, pos4, w->P.single ? 1 : 0); if (r) return r; CU(cudaStreamSynchronize(w->stream)); }
    if (vel4) { r = write_rows_f4(w, w->P.vel, first, count, vel4); if (r) return r; }
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}
// Slab (multi-GPU) worlds move rows between ranks; force / torque / angular arrays do not travel with them,
// so the setters refuse instead of leaving a force on a row that now holds another body.
static int no_slab(rb_world* w, const char* what) {
    if (w && w->d_dn) return fail(w, RB_ERR_INVALID, "%s is not available on slab (multi-GPU) worlds", what);
    return RB_OK;
}
// force / torque / angular state on slab worlds: only when the world was created for it (rb_config.slab_forces on every rank:
// the arrays exist from the build on, the migrant records carry them)
static int slab_dynamics_ok(rb_world* w, const char* what) {
    if (w && w->d_dn && !w->cfg.slab_forces) return fail(w, RB_ERR_INVALID, "%s on a slab (multi-GPU) world needs rb_config.slab_forces = 1 on every rank (migrating bodies then carry force / torque / angular state)", what);
    return RB_OK;
}
// Entity::setPosition (model/src/Entity.cpp:373-391) from the host, e.g. a teleport by game logic: unlike rb_set_state it
// moves the model matrices along (and reproduces Q1: the stored position is W.t + p)
int rb_entity_set_position(rb_world* w, uint32_t body, const float* pos3) {
    int r = range_ok(w, body, 1); if (r) return r;
    r = no_slab(w, "rb_entity_set_position"); if (r) return r;
    if (!pos3 || !finite_all(pos3, 3)) return fail(w, RB_ERR_INVALID, "rb_entity_set_position: position is NULL or not finite");
    CU(cudaSetDevice(w->cfg.device));
    w->vl_force = true;
    r = rb_ensure_mt(w); if (r) return r;
    rb_launch_entity_set_position(w->P.pos, w->P.mt, w->P.pt, w->P.wt, w->d_inv, body, pos3, w->stream); w->launches++;
    CU(cudaGetLastError());
    return RB_OK;
}
int rb_set_flags(rb_world* w, uint32_t first, uint32_t count, const uint32_t* values, uint32_t mask) {
    int r = range_ok(w, first, count); if (r) return r;
    w->vl_force = true;
    if (!values) return fail(w, RB_ERR_INVALID, "values is NULL");
    r = ensure_stage_u(w, count); if (r) return r;
    CU(cudaMemcpyAsync(w->d_stage_u, values, (size_t)count * 4, cudaMemcpyHostToDevice, w->stream));
    rb_launch_set_flags(w->P.flags, w->d_stage_u, w->d_inv, first, count, mask & (RB_ACTIVE | RB_GRAVITY | RB_CONTACT), w->stream); w->launches++;
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}
static int ensure_force(rb_world* w) {
    if (w->P.force) return RB_OK;
    float4* f; ALLOC(f, w->P.nb);
    rb_launch_fill_f4(f, w->P.nb, make_float4(0.f, 0.f, 0.f, 1.0f), w->stream); w->launches++;   // mass 1, StateVector.cpp:4-9
    w->P.force = f;
    return RB_OK;
}
static int ensure_angular(rb_world* w) {
    if (w->P.avel) return RB_OK;
    float4 *a, *b, *t;
    ALLOC(a, w->P.nb); ALLOC(b, w->P.nb); ALLOC(t, w->P.nb);
    CU(cudaMemsetAsync(a, 0, (size_t)w->P.nb * 16, w->stream));
    CU(cudaMemsetAsync(b, 0, (size_t)w->P.nb * 16, w->stream));
    CU(cudaMemsetAsync(t, 0, (size_t)w->P.nb * 16, w->stream));
    w->P.apos = a; w->P.avel = b; w->P.torque = t;
    return ensure_force(w);
}
int rb_set_force(rb_world* w, uint32_t first, uint32_t count, const float* xyz4) {
    int r = range_ok(w, first, count); if (r) return r;
    r = slab_dynamics_ok(w, "rb_set_force"); if (r) return r;
    if (!xyz4) return fail(w, RB_ERR_INVALID, "xyz4 is NULL");
    r = ensure_force(w); if (r) return r;
    r = ensure_stage(w, count); if (r) return r;
    CU(cudaMemcpyAsync(w->d_stage, xyz4, (size_t)count * 16, cudaMemcpyHostToDevice, w->stream));
    rb_launch_apply_force(w->P.flags, w->d_stage, w->P.force, w->d_inv, first, count, 1, w->stream); w->launches++;
    return RB_OK;
}
int rb_set_torque(rb_world* w, uint32_t first, uint32_t count, const float* xyz4) {
    int r = range_ok(w, first, count); if (r) return r;
    r = slab_dynamics_ok(w, "rb_set_torque"); if (r) return r;
    if (!xyz4) return fail(w, RB_ERR_INVALID, "xyz4 is NULL");
    r = ensure_angular(w); if (r) return r;
    r = ensure_stage(w, count); if (r) return r;
    CU(cudaMemcpyAsync(w->d_stage, xyz4, (size_t)count * 16, cudaMemcpyHostToDevice, w->stream));
    rb_launch_apply_force(w->P.flags, w->d_stage, w->P.torque, w->d_inv, first, count, 0, w->stream); w->launches++;
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}
int rb_set_angular(rb_world* w, uint32_t first, uint32_t count, const float* ap4, const float* av4) {
    int r = range_ok(w, first, count); if (r) return r;
    r = slab_dynamics_ok(w, "rb_set_angular"); if (r) return r;
    r = ensure_angular(w); if (r) return r;
    if (ap4) { r = write_rows_f4(w, w->P.apos, first, count, ap4); if (r) return r; CU(cudaStreamSynchronize(w->stream)); }
    if (av4) { r = write_rows_f4(w, w->P.avel, first, count, av4); if (r) return r; }
    CU(cudaStreamSynchronize(w->stream));
    return RB_OK;
}

uint32_t rb_host_row_capacity(const rb_world* w) {
    if (!w || !w->built) return 0;
    return w->d_dn ? w->cap_local - 3u * w->ghost_cap : w->n_owned_host;
}
int rb_step_host(rb_world* w, int ms, const float* force4, float* pos4, uint64_t* n_contacts) {
    if (!w || !w->built) return RB_ERR_INVALID;
    int r;
    if (force4) { r = rb_set_force(w, 0, w->n_owned_host, force4); if (r) return r; }
    r = rb_step(w, ms); if (r) return r;
    if (pos4) { r = read_rows_f4(w, w->P.pos, pos4, w->n_owned_host, w->P.single ? 1 : 0); if (r) return r; }
    CU(cudaMemcpyAsync(w->h_counters + RB_C_NCONTACT, w->ncontact_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->contacts_last = w->h_counters[RB_C_NCONTACT];
    if (n_contacts) *n_contacts = w->contacts_last;
    return RB_OK;
}

// Pipelined per-tick host round trip: the H2D copy of this tick's forces, the tick itself and the D2H copy
// of its positions run on three streams over a ring of RB_PIPE (4) staging buffers, so the two copy
// engines overlap the compute of neighbouring ticks -- the reference's own threading model (its physics
// thread produces, its render thread reads a little behind under Physics::_lock). The call returns once
// the results of the call made RB_PIPE - 1 = 3 calls earlier have landed in the host buffers passed to
// THAT call; callers rotate four pos4 buffers and finish with rb_sync().
static_assert(RB_PIPE == RB_HOST_PIPE_DEPTH, "header and implementation disagree on the ring depth");
static int step_host_async(rb_world* w, int ms, const float* force4, float* pos4, uint64_t* n_contacts_done, int stride);
int rb_step_host_async(rb_world* w, int ms, const float* force4, float* pos4, uint64_t* n_contacts_done) { return step_host_async(w, ms, force4, pos4, n_contacts_done, 4); }
int rb_step_host_async3(rb_world* w, int ms, const float* force3, float* pos3, uint64_t* n_contacts_done) { return step_host_async(w, ms, force3, pos3, n_contacts_done, 3); }
static int step_host_async(rb_world* w, int ms, const float* force4, float* pos4, uint64_t* n_contacts_done, int stride) {
    if (!w || !w->built) return RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    if (force4) { int e = no_slab(w, "the force upload of rb_step_host_async"); if (e) return e; }
    // slab worlds gain and lose rows: the staging ring is sized for the most rows the world can ever own
    // (rb_host_row_capacity), and a call never moves more rows than that
    if (!w->p_rows) w->p_rows = rb_host_row_capacity(w);
    const size_t nb = std::min<size_t>(w->n_owned_host, w->p_rows);
    if (!w->s_h2d) {
        CU(cudaStreamCreateWithFlags(&w->s_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&w->s_d2h, cudaStreamNonBlocking));
        for (int k = 0; k < RB_PIPE; k++) {
            CU(cudaEventCreateWithFlags(&w->pe_h2d[k], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&w->pe_force[k], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&w->pe_step[k], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&w->pe_d2h[k], cudaEventDisableTiming));
            ALLOC(w->p_force[k], w->p_rows); ALLOC(w->p_pos[k], w->p_rows);
        }
        CU(cudaMallocHost((void**)&w->p_cnt, RB_PIPE * sizeof(unsigned long long)));
        for (int k = 0; k < RB_PIPE; k++) w->p_cnt[k] = 0;
    }
    const int b = (int)(w->p_calls % RB_PIPE);
    int r;
    // list ticks take the forces and leave the positions themselves (no extra pass over the bodies);
    // whether this tick will be one is rb_step's decision, so it is asked first
    const bool fused_io = rb_step_will_use_lists(w) && !w->d_dn;      // (slab worlds hand out rows, not ids)
    if (force4) {
        r = ensure_force(w); if (r) return r;
        if (w->p_calls >= RB_PIPE) CU(cudaStreamWaitEvent(w->s_h2d, w->pe_force[b], 0));   // staging b was consumed RB_PIPE calls ago
        CU(cudaMemcpyAsync(w->p_force[b], force4, nb * 4 * stride, cudaMemcpyHostToDevice, w->s_h2d));
        CU(cudaEventRecord(w->pe_h2d[b], w->s_h2d));
        CU(cudaStreamWaitEvent(w->stream, w->pe_h2d[b], 0));
        if (fused_io) { w->P.force_in = (const float*)w->p_force[b]; w->P.force_in_stride = (uint32_t)stride; }
        else {
            if (stride == 4) rb_launch_apply_force(w->P.flags, w->p_force[b], w->P.force, w->d_inv, 0, (uint32_t)nb, 1, w->stream);
            else rb_launch_apply_force3(w->P.flags, (const float*)w->p_force[b], w->P.force, w->d_inv, (uint32_t)nb, w->stream);
            w->launches++;
            CU(cudaEventRecord(w->pe_force[b], w->stream));
        }
    }
    if (fused_io && pos4) {
        if (w->p_calls >= RB_PIPE) CU(cudaStreamWaitEvent(w->stream, w->pe_d2h[b], 0));     // output staging b was read RB_PIPE calls ago
        w->P.pos_out = (float*)w->p_pos[b]; w->P.pos_out_stride = (uint32_t)stride;
    }
    r = rb_step(w, ms);
    w->P.force_in = nullptr; w->P.pos_out = nullptr;
    if (r) return r;
    if (fused_io && force4) CU(cudaEventRecord(w->pe_force[b], w->stream));
    if (!fused_io && w->p_calls >= RB_PIPE) CU(cudaStreamWaitEvent(w->stream, w->pe_d2h[b], 0));         // output staging b was read RB_PIPE calls ago
    if (pos4 && !fused_io) {
        if (stride == 3) { rb_launch_rows_to_ids_f3(w->P.pos, w->d_inv ? w->P.gid : nullptr, (float*)w->p_pos[b], (uint32_t)nb, w->stream); w->launches++; }
        else if (w->d_inv || w->P.single) { rb_launch_rows_to_ids_f4(w->P.pos, w->d_inv ? w->P.gid : nullptr, w->p_pos[b], (uint32_t)nb, w->P.single ? 1 : 0, w->stream); w->launches++; }
        else CU(cudaMemcpyAsync(w->p_pos[b], w->P.pos, nb * 16, cudaMemcpyDeviceToDevice, w->stream));
    }
    CU(cudaEventRecord(w->pe_step[b], w->stream));
    CU(cudaStreamWaitEvent(w->s_d2h, w->pe_step[b], 0));
    if (pos4) CU(cudaMemcpyAsync(pos4, w->p_pos[b], nb * 4 * stride, cudaMemcpyDeviceToHost, w->s_d2h));
    CU(cudaMemcpyAsync(&w->p_cnt[b], w->ncontact_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->s_d2h));
    CU(cudaEventRecord(w->pe_d2h[b], w->s_d2h));
    // results of the call made RB_PIPE - 1 calls ago
    if (w->p_calls >= RB_PIPE - 1) {
        const int d = (int)((w->p_calls + 1) % RB_PIPE);
        CU(cudaEventSynchronize(w->pe_d2h[d]));
        if (n_contacts_done) *n_contacts_done = w->p_cnt[d];
    } else if (n_contacts_done) *n_contacts_done = 0;
    w->p_calls++;
    return RB_OK;
}

int rb_set_body_ids(rb_world* w, const uint32_t* ids, uint32_t n) {
    if (!w || !ids) return RB_ERR_INVALID;
    if (w->built) return fail(w, RB_ERR_INVALID, "rb_set_body_ids after rb_build");
    w->h_gid.assign(ids, ids + n);
    return RB_OK;
}
int rb_get_body_ids(rb_world* w, uint32_t* out) {
    if (!w || !w->built || !out) return RB_ERR_INVALID;
    if (!w->d_dn) { for (uint32_t i = 0; i < w->n_owned_host; i++) out[i] = i; return RB_OK; }   // rows are reported in id order
    int r = rb_dist_refresh_counts(w); if (r) return r;
    CU(cudaMemcpy(out, w->d_gid, (size_t)w->n_owned_host * 4, cudaMemcpyDeviceToHost));
    return RB_OK;
}

int rb_get_stats(rb_world* w, rb_stats_t* out) {
    if (!w || !out) return RB_ERR_INVALID;
    memset(out, 0, sizeof *out);
    if (w->built) { int r = fetch_counters(w); if (r) return r; }
    const unsigned long long* c = w->h_counters;
    out->steps = w->steps;
    if (c) {
        out->tests_ss = c[RB_C_TESTS_SS]; out->tests_st = c[RB_C_TESTS_ST]; out->cand_ss = c[RB_C_CAND_SS]; out->cand_st = c[RB_C_CAND_ST];
        out->hits_ss = c[RB_C_HITS_SS]; out->hits_st = c[RB_C_HITS_ST]; out->contacts_last = c[RB_C_NCONTACT];
        out->contacts_dropped = c[RB_C_DROPPED]; out->margin_overflow = c[RB_C_OVERFLOW];
    }
    if (w->built && w->d_dn) { int r = rb_dist_refresh_counts(w); if (r) return r; }
    if (w->built && w->vl_on) {
        uint32_t v[RB_VL_WORDS];
        CU(cudaMemcpyAsync(v, w->P.vl, sizeof v, cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        const uint32_t nh = v[RB_VL_NHOT_LAST];
        { static const bool dbg = [] { const char* e = getenv("RB_DEBUG_LISTS"); return e && *e == '1'; }();      // (development switch)
          if (dbg) fprintf(stderr, "[lists] period %u spacing %u floor %u hot %u no-list rows of the last rebuild: crowded %u, wide %u, tile overflow %u\n",
                           w->vl_period, w->vl_spacing, v[RB_VL_FLOOR], nh, v[RB_VL_NOLIST], v[RB_VL_NOLIST + 1], v[RB_VL_NOLIST + 2]); }
        out->list_rebuilds = v[RB_VL_NREBUILD]; out->list_ticks = v[RB_VL_NTICK]; out->list_skin = w->P.vl_skin;
        out->list_hot_last = nh; out->list_rebuild_reason = v[RB_VL_WHY_LAST];
        for (int b = 0; b < 4; b++) out->list_rebuilds_by_reason[b] = v[RB_VL_NWHY + b];
    }
    out->n_bodies = w->built ? w->n_owned_host : (uint32_t)w->h_flags.size();
    out->n_spheres = w->d_dn ? w->n_owned_host : w->P.ns; out->n_geoms = (uint32_t)w->h_geom_start.size() - 1; out->n_tris = w->P.nt;
    out->sphere_cols_x = w->P.scx; out->sphere_cols_z = w->P.scz; out->tri_cols_x = w->P.tcx; out->tri_cols_z = w->P.tcz;
    out->device_bytes = w->device_bytes;
    out->halo_sent_last = w->dist_stats.halo_sent_last; out->halo_recv_last = w->dist_stats.halo_recv_last;
    out->migrated_out_last = w->dist_stats.migrated_out_last; out->migrated_in_last = w->dist_stats.migrated_in_last;
    return RB_OK;
}
int rb_profile(rb_world* w, int on) { if (!w) return RB_ERR_INVALID; w->profile = on != 0; for (auto& k : w->kernel_ms) k = 0; return RB_OK; }
int rb_get_kernel_ms(rb_world* w, float* out5, uint64_t* launches) {
    if (!w || !out5) return RB_ERR_INVALID;
    for (int k = 0; k < 5; k++) out5[k] = w->kernel_ms[k];
    if (launches) *launches = w->launches;
    return RB_OK;
}
uint64_t rb_launch_count(const rb_world* w) { return w ? w->launches : 0; }

// ---- Physics::addEntity after the build (physics/src/Physics.cpp:47-49) ---------------------------
// The reference appends the entity to its list; from the next tick on it is integrated and re-inserted into the tree like
// everybody else (physics/src/OSP.cpp:182-210). Here every device structure is sized and laid out at rb_build (grids by the
// largest radius, lists, tiles, row order), so an add re-lays the world out: the complete state comes back from the device
// -- static triangles, transformed spheres, W, poses, velocities, flags, both model translations, forces, torques, angular
// state -- a fresh world is built from it plus the new bodies, the dynamic state is restored bit for bit, and the handle
// takes over the new internals. Cost O(world) per call (rb_add_models adds a batch in one go); ids of existing bodies and
// geometries are unchanged, the new bodies get the next ids. Triangles stay refused after the build.
static int world_grow(rb_world* w, const std::vector<GrowBody>& add, uint32_t* first_body_id) {
    if (w->d_dn) return fail(w, RB_ERR_INVALID, "adding bodies after rb_build is not available on slab (multi-GPU) worlds");
    for (const GrowBody& g : add) {
        if (g.nsph == 0) return fail(w, RB_ERR_INVALID, "a model needs at least one sphere");
        if (!finite_all(g.obj4, (size_t)g.nsph * 4) || !finite_all(g.pos3, 3) || !finite_all(g.vel3, 3) || !finite_all(g.W12, 12)) return fail(w, RB_ERR_INVALID, "non-finite model input");
    }
    CU(cudaSetDevice(w->cfg.device));
    CU(cudaStreamSynchronize(w->stream));
    RbParams& P = w->P;
    const uint32_t nb = w->n_owned_host, ns = w->h_body_start.back(), nt = P.nt;
    // ---- the world as it stands, in id order ----
    std::vector<float> tri4((size_t)std::max<uint32_t>(nt, 1) * 12), sph((size_t)std::max<uint32_t>(ns, 1) * 4), wt((size_t)std::max<uint32_t>(nb, 1) * 4, 0.f), w3;
    std::vector<float> pos((size_t)std::max<uint32_t>(nb, 1) * 4), vel(pos.size()), mt(pos.size()), pt(pos.size()), force, torque, apos, avel;
    std::vector<uint32_t> flags(std::max<uint32_t>(nb, 1));
    if (nt) CU(cudaMemcpyAsync(tri4.data(), P.tri, (size_t)nt * 48, cudaMemcpyDeviceToHost, w->stream));
    int r;
    if (P.single) { r = read_rows_f4(w, P.sph_obj, sph.data(), nb); if (r) return r; CU(cudaStreamSynchronize(w->stream)); }
    else if (ns) CU(cudaMemcpyAsync(sph.data(), P.sph_obj, (size_t)ns * 16, cudaMemcpyDeviceToHost, w->stream));
    if (P.wt) { r = read_rows_f4(w, P.wt, wt.data(), nb); if (r) return r; CU(cudaStreamSynchronize(w->stream)); }
    if (w->d_w3) { w3.resize((size_t)nb * 12); CU(cudaMemcpyAsync(w3.data(), w->d_w3, (size_t)nb * 48, cudaMemcpyDeviceToHost, w->stream)); }
    r = rb_ensure_mt(w); if (r) return r;
    struct { const float4* rows; std::vector<float>* out; } arrays[] = { { P.pos, &pos }, { P.vel, &vel }, { P.mt, &mt }, { P.pt, &pt },
                                                                         { P.force, &force }, { P.torque, &torque }, { P.apos, &apos }, { P.avel, &avel } };
    for (auto& a : arrays) {
        if (!a.rows) continue;
        a.out->resize((size_t)std::max<uint32_t>(nb, 1) * 4);
        r = read_rows_f4(w, a.rows, a.out->data(), nb); if (r) return r;
        CU(cudaStreamSynchronize(w->stream));      // (the id-ordered staging buffer is reused by the next array)
    }
    r = read_rows_u32(w, P.flags, flags.data(), nb); if (r) return r;
    CU(cudaStreamSynchronize(w->stream));
    // ---- the same world plus the new bodies, laid out afresh ----
    rb_config c = w->cfg_user;
    c.stream = w->own_stream ? nullptr : (void*)w->stream;
    rb_world* f = nullptr;
    r = rb_world_create(&c, &f);
    if (r) return fail(w, r, "re-layout for a post-build add failed: %s", g_create_error.c_str());
    auto bail = [&](int code) { w->err = f->err; rb_world_destroy(f); return code; };
    for (size_t g = 0; g + 1 < w->h_geom_start.size(); g++) {
        const uint32_t t0 = w->h_geom_start[g], t1 = w->h_geom_start[g + 1];
        std::vector<float> xyz9((size_t)(t1 - t0) * 9);
        for (uint32_t t = t0; t < t1; t++) for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) xyz9[(size_t)(t - t0) * 9 + v * 3 + k] = tri4[(size_t)t * 12 + v * 4 + k];
        r = rb_add_static_geometry(f, t1 - t0, xyz9.data(), nullptr); if (r) return bail(r);
    }
    for (uint32_t b = 0; b < nb; b++) {
        float W12[12] = { 1, 0, 0, 0, 1, 0, 0, 0, 1, wt[(size_t)b * 4], wt[(size_t)b * 4 + 1], wt[(size_t)b * 4 + 2] };
        if (!w3.empty()) for (int rr = 0; rr < 3; rr++) for (int cc = 0; cc < 3; cc++) W12[rr * 3 + cc] = w3[(size_t)b * 12 + rr * 4 + cc];
        const uint32_t s0 = w->h_body_start[b], s1 = w->h_body_start[b + 1];
        // (add_body's finiteness check is for user input; a body that has blown up to Inf / NaN keeps stepping like in the reference)
        f->h_obj.insert(f->h_obj.end(), sph.begin() + (size_t)s0 * 4, sph.begin() + (size_t)s1 * 4);
        for (uint32_t i = s0; i < s1; i++) f->h_sph_body.push_back(b);
        f->h_body_start.push_back(s1);
        f->h_W.insert(f->h_W.end(), W12, W12 + 12);
        if (W12[9] != 0.f || W12[10] != 0.f || W12[11] != 0.f) f->any_wt = true;
        f->h_pos.insert(f->h_pos.end(), &pos[(size_t)b * 4], &pos[(size_t)b * 4] + 3);
        f->h_vel.insert(f->h_vel.end(), &vel[(size_t)b * 4], &vel[(size_t)b * 4] + 3);
        f->h_flags.push_back(flags[b] & (RB_ACTIVE | RB_GRAVITY));
        f->h_presummed.push_back(1);
    }
    for (const GrowBody& g : add) { r = add_body(f, g.nsph, g.obj4, g.W12, g.pos3, g.vel3, g.flags); if (r) return bail(r); }
    r = rb_build(f); if (r) return bail(r);
    // ---- the dynamic state of the bodies that were there, bit for bit (new bodies start like at a first build) ----
    if (nb) {
        RbParams& Q = f->P;
        r = write_rows_f4(f, Q.pos, 0, nb, pos.data(), Q.single ? 1 : 0); if (r) return bail(r);
        CU(cudaStreamSynchronize(f->stream));
        struct { float4* rows; const std::vector<float>* in; } back[] = { { Q.vel, &vel }, { Q.mt, &mt }, { Q.pt, &pt } };
        for (auto& a : back) { r = write_rows_f4(f, a.rows, 0, nb, a.in->data()); if (r) return bail(r); CU(cudaStreamSynchronize(f->stream)); }
        if (!force.empty()) { r = ensure_force(f); if (r) return bail(r); r = write_rows_f4(f, Q.force, 0, nb, force.data()); if (r) return bail(r); CU(cudaStreamSynchronize(f->stream)); }
        if (!avel.empty()) {
            r = ensure_angular(f); if (r) return bail(r);
            struct { float4* rows; const std::vector<float>* in; } ang[] = { { Q.torque, &torque }, { Q.apos, &apos }, { Q.avel, &avel } };
            for (auto& a : ang) { r = write_rows_f4(f, a.rows, 0, nb, a.in->data()); if (r) return bail(r); CU(cudaStreamSynchronize(f->stream)); }
        }
        r = ensure_stage_u(f, nb); if (r) return bail(r);
        CU(cudaMemcpyAsync(f->d_stage_u, flags.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, f->stream));
        rb_launch_set_flags(Q.flags, f->d_stage_u, f->d_inv, 0, nb, RB_ACTIVE | RB_GRAVITY | RB_CONTACT, f->stream); f->launches++;
        CU(cudaStreamSynchronize(f->stream));
    }
    f->steps = w->steps; f->pose_updates = w->pose_updates; f->launches += w->launches; f->profile = w->profile;
    f->vl_force = true;
    if (first_body_id) *first_body_id = nb;
    std::swap(*w, *f);          // the caller's handle now owns the new world ...
    rb_world_destroy(f);        // ... and the old device state goes
    return RB_OK;
}

} // extern "C"

int rb_collide_only(rb_world* w, bool resolve) { return enqueue_collide(w, resolve); }

// ---- multi-GPU placeholders until rb_dist.cu lands ---------------------------------------------
#ifndef RB_HAVE_DIST
int rb_dist_step_exchange(rb_world* w, int) { return fail(w, RB_ERR_INVALID, "multi-GPU slabs not built"); }
int rb_dist_list_exchange(rb_world* w, int) { return fail(w, RB_ERR_INVALID, "multi-GPU slabs not built"); }
int rb_dist_refresh_counts(rb_world*) { return RB_OK; }
int rb_dist_sticky_error(rb_world*, unsigned long long, unsigned long long) { return RB_OK; }
int rb_dist_rows_moved(rb_world*) { return RB_OK; }
void rb_dist_destroy(RbDist*) {}
extern "C" int rb_dist_unique_id(const char*, uint8_t*) { g_create_error = "multi-GPU slabs not built"; return RB_ERR_INVALID; }
extern "C" int rb_dist_init(rb_world* w, const char*, const uint8_t*) { return fail(w, RB_ERR_INVALID, "multi-GPU slabs not built"); }
extern "C" int rb_dist_attach_peers(rb_world* w, rb_world*, rb_world*) { return fail(w, RB_ERR_INVALID, "multi-GPU slabs not built"); }
extern "C" int rb_step_group(rb_world**, int, int) { return RB_ERR_INVALID; }
extern "C" int rb_dist_p2p_export(rb_world* w, uint8_t*) { return fail(w, RB_ERR_INVALID, "multi-GPU slabs not built"); }
extern "C" int rb_dist_p2p_connect(rb_world* w, const uint8_t*, const uint8_t*) { return fail(w, RB_ERR_INVALID, "multi-GPU slabs not built"); }
extern "C" int rb_detect_group(rb_world**, int) { return RB_ERR_INVALID; }
#endif
