// rb_dist.cu -- multi-GPU x-slabs: one world per GPU owns the bodies whose centre
// lies in its x-slab; static geometry is replicated. Every tick, between integration and the
// broadphase, each world
//   1. packs (K6a) the owned spheres that can reach a neighbour's slab ("ghosts", 32 B each) and the
//      bodies that crossed a face ("migrants", full state, 96 B each) into two fixed-capacity buffers,
//   2. exchanges them with the left/right neighbour -- NCCL send/recv over NVLink on the world's own
//      stream (one group call, no host sync, no count round trip: the counts ride in the buffer), or a
//      device-to-device copy when the neighbour world lives in the same process (tests, single GPU),
//   3. removes its emigrants (swap-remove), appends immigrants as owned bodies and ghosts behind them
//      (K6b-d), and bins everything into the slab-local sphere grid (K6e).
// Bodies carry a global id; the canonical order of the narrowphase is defined on global ids, so the
// result is bit-identical to a single world holding all bodies (tests/test_dist_*).
// Resolution only writes the sphere's own body (GeometryMath.cpp:489-554), so ghosts are read-only.
#include "rb_internal.h"

#include <dlfcn.h>

#define RB_BLOCK 256

struct RbHaloBuf {
    uint8_t* base = nullptr; size_t bytes = 0;
    RbHaloHeader* hdr() const { return (RbHaloHeader*)base; }
    RbGhostRec* ghosts() const { return (RbGhostRec*)(base + sizeof(RbHaloHeader)); }
    RbMigrRec* migr(uint32_t ghost_cap, uint32_t ghost_stride = sizeof(RbGhostRec)) const { return (RbMigrRec*)(base + sizeof(RbHaloHeader) + (size_t)ghost_cap * ghost_stride); }
};

// ---- NCCL through dlopen (no link-time dependency; the host says which libnccl to use) -----------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const int NCCL_UINT8 = 1;   // ncclUint8 in nccl.h's ncclDataType_t

static int load_nccl(rb_world* w, const char* path) {
    if (g_nccl.h) return RB_OK;
    const char* cands[] = { path, "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* c : cands) { if (c && *c) { h = dlopen(c, RTLD_NOW | RTLD_GLOBAL); if (h) break; } }
    if (!h) return fail(w, RB_ERR_NCCL, "cannot dlopen NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
#define SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return fail(w, RB_ERR_NCCL, "NCCL symbol %s missing", name); } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.h = h;
    return RB_OK;
}
#define NC(call) do { int r_ = (call); if (r_ != 0) return fail(w, RB_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

// Exchange transports:
//   RB_X_NCCL   ncclSend/ncclRecv of the out buffers into the neighbours' in buffers (baseline)
//   RB_X_P2P    fused compute + communication: K1 stores ghost / migrant records straight into the neighbour's
//               in buffer through a CUDA-IPC peer pointer (NVLink stores); a one-thread signal kernel then
//               copies the 16-B header over and raises the neighbour's arrival flag (release, system scope),
//               the neighbour's wait kernel acquires it. In buffers are double-buffered by tick parity.
//   RB_X_LOCAL  the same peer stores between slab worlds of one process (tests on one GPU; the host
//               sequences the phases, so no flag wait)
enum { RB_X_NONE = 0, RB_X_NCCL = 1, RB_X_P2P = 2, RB_X_LOCAL = 3 };

struct RbDist {
    int mode = RB_X_NONE;
    ncclComm_t comm = nullptr;
    rb_world* peer[2] = { nullptr, nullptr };       // in-process neighbours (left, right)
    RbHaloBuf out[2];                                // local staging (NCCL) and local headers (all modes)
    uint8_t* blob = nullptr; size_t blob_bytes = 0;  // in[side][parity] buffers + arrival flags, one cudaMalloc (IPC-exportable)
    size_t buf_bytes = 0;
    uint8_t* peer_blob[2] = { nullptr, nullptr };    // neighbours' blobs (IPC-opened or same-process pointers)
    bool peer_opened[2] = { false, false };
    RbDistParams D;                                  // parity-independent parameters (host copy)
    RbDistParams Dp[2];                              // per-parity parameters (host copies of d_dp)
    bool params_valid = false;
    RbDistParams* d_dp[2] = { nullptr, nullptr };    // device copies, one per parity
    uint32_t ghost_cap = 0, migr_cap = 0;
    uint32_t ghost_stride = sizeof(RbGhostRec), migr_stride = sizeof(RbMigrRec);      // bytes per record (compound bodies: per-world sizes)
    uint32_t halo_step = 0;
    uint32_t* h_dn = nullptr;                        // pinned
    RbHaloHeader* h_hdr = nullptr;                   // pinned, 4 headers (out L/R, in L/R)
    RbHaloBuf in_buf(int side, int parity) const { RbHaloBuf b; b.base = blob + (size_t)(side * 2 + parity) * buf_bytes; b.bytes = buf_bytes; return b; }
    static RbHaloBuf buf_at(uint8_t* blob_base, size_t buf_bytes, int side, int parity) { RbHaloBuf b; b.base = blob_base + (size_t)(side * 2 + parity) * buf_bytes; b.bytes = buf_bytes; return b; }
    uint32_t* arrive(int side) const { return (uint32_t*)(blob + 4 * buf_bytes) + 4 * side; }
    static uint32_t* arrive_at(uint8_t* blob_base, size_t buf_bytes, int side) { return (uint32_t*)(blob_base + 4 * buf_bytes) + 4 * side; }
    uint32_t* timeout_flag() const { return (uint32_t*)(blob + 4 * buf_bytes) + 8; }
};

// header + arrival flag to the neighbour: the records are already there (K1 stored them remotely)
__global__ void rb_halo_signal_kernel(const RbHaloHeader* out_hdr0, const RbHaloHeader* out_hdr1, RbHaloHeader* peer_hdr0, RbHaloHeader* peer_hdr1,
                                      uint32_t* peer_arrive0, uint32_t* peer_arrive1, uint32_t step) {
    if (threadIdx.x || blockIdx.x) return;
    if (peer_hdr0) *peer_hdr0 = *out_hdr0;
    if (peer_hdr1) *peer_hdr1 = *out_hdr1;
    __threadfence_system();
    if (peer_arrive0) *(volatile uint32_t*)peer_arrive0 = step + 1;
    if (peer_arrive1) *(volatile uint32_t*)peer_arrive1 = step + 1;
}
// bounded wait for the neighbours' arrival flags (a dead neighbour raises the timeout flag instead of hanging)
__device__ __forceinline__ void halo_wait(const uint32_t* arrive0, const uint32_t* arrive1, uint32_t step, uint32_t* timeout_flag) {
    const long long t0 = clock64();
    const long long limit = 4000000000ll;      // ~2 s of SM clocks
    for (int side = 0; side < 2; ++side) {
        const uint32_t* a = side ? arrive1 : arrive0;
        if (!a) continue;
        while (*(volatile const uint32_t*)a < step + 1) {
            if (clock64() - t0 > limit) { *timeout_flag = 1u; break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

// K6a (pack ghosts and migrants) is fused into rb_integrate_bin_kernel (rb_kernels.cu).

// ---- K6b: swap-remove the emigrants (few per tick; one thread keeps it simple and ordered) --------
// list_tick: K1 has linked the hot rows into the cell table by ROW (their entry index rides in P.rank): a row that is moved into
// a hole takes its entry along, an emigrant's entry is invalidated
__global__ void rb_halo_remove_kernel(RbParams P, RbDistParams D, const uint32_t* arrive0, const uint32_t* arrive1, uint32_t step, uint32_t* timeout_flag, int list_tick) {
    if (threadIdx.x || blockIdx.x) return;
    if (arrive0 || arrive1) halo_wait(arrive0, arrive1, step, timeout_flag);     // fused-transport ticks: the neighbours' records must have landed
    uint32_t m = *D.emig_cnt;
    uint32_t n = D.dn[0];
    // descending order so that the element moved into a hole is never itself a pending hole
    for (uint32_t i = 1; i < m; ++i) { uint32_t v = D.emig[i]; int j = (int)i - 1; while (j >= 0 && D.emig[j] < v) { D.emig[j + 1] = D.emig[j]; --j; } D.emig[j + 1] = v; }
    float4* sph = D.sph_obj_w;
    for (uint32_t i = 0; i < m; ++i) {
        uint32_t h = D.emig[i], last = n - 1;
        if (list_tick) {
            const uint32_t kh = P.key[h], kl = P.key[last];
            if (kh != RB_NONE && (kh & RB_HOT_BIT)) P.hot_row[P.rank[h]] = RB_NONE;
            if (h != last && kl != RB_NONE && (kl & RB_HOT_BIT)) P.hot_row[P.rank[last]] = h;
        }
        if (h != last && P.kstride) {
            // compound bodies: the K sphere rows travel with the body row
            P.pos[h] = P.pos[last]; P.vel[h] = P.vel[last]; P.mt[h] = P.mt[last]; P.pt[h] = P.pt[last]; P.flags[h] = P.flags[last];
            D.gid[h] = D.gid[last]; P.body_cnt[h] = P.body_cnt[last];
            if (P.force) P.force[h] = P.force[last];
            if (P.avel) { P.torque[h] = P.torque[last]; P.apos[h] = P.apos[last]; P.avel[h] = P.avel[last]; }
            for (uint32_t j = 0; j < P.kstride; ++j) {
                const uint32_t a = h * P.kstride + j, b = last * P.kstride + j;
                sph[a] = sph[b]; P.ws[a] = P.ws[b]; P.key[a] = P.key[b]; P.rank[a] = P.rank[b];
                if (P.ce_build) P.ce_build[a].w = -1.0f;      // (another body's sphere now: its persistent triangle list is not this row's)
            }
        } else if (h != last) {
            P.pos[h] = P.pos[last]; P.vel[h] = P.vel[last]; P.mt[h] = P.mt[last]; P.pt[h] = P.pt[last]; P.flags[h] = P.flags[last];
            sph[h] = sph[last]; D.gid[h] = D.gid[last]; P.ws[h] = P.ws[last]; P.key[h] = P.key[last]; P.rank[h] = P.rank[last];
            if (P.force) P.force[h] = P.force[last];
            if (P.avel) { P.torque[h] = P.torque[last]; P.apos[h] = P.apos[last]; P.avel[h] = P.avel[last]; }
        }
        n = last;
    }
    D.dn[0] = n;
    D.counters[RB_C_MIGRATED] = m;
}

// ---- K6c: one launch unpacks both kinds. Blocks [0, imm_blocks): immigrants become owned rows (appended
// behind the owned rows); the other blocks: ghosts, stored from the END of the arrays downwards so that they
// do not depend on the final owned count. Both are binned on the way. -----------------------------------
// list_tick: persistent-list ticks -- immigrants only get their column (the conditional bin kernel ranks them) and
// force a list rebuild; ghosts are not binned at all but linked into the hot-body cell table behind the own hot
// bodies (entry hot_cap + ghost index), where every owned body looks them up.
__global__ void __launch_bounds__(RB_BLOCK) rb_halo_unpack_kernel(RbParams P, RbDistParams D, uint32_t imm_blocks, int list_tick) {
    const uint32_t ghost_floor = D.cap_local - 3u * D.ghost_cap;      // rows [ghost_floor, cap) are reserved for ghosts
    if (P.kstride) {
        // compound bodies: one thread per record, the body's spheres go to rows [row * K, row * K + spheres)
        const uint32_t K = P.kstride;
        if (blockIdx.x < imm_blocks) {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            const uint32_t n0 = min(D.in_hdr[0]->n_migr, D.migr_cap), n1 = min(D.in_hdr[1]->n_migr, D.migr_cap);
            if (t >= n0 + n1) return;
            const float4* r = reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(t < n0 ? D.in_migr[0] : D.in_migr[1]) + (size_t)(t < n0 ? t : t - n0) * D.migr_stride);
            const uint32_t i = atomicAdd(&D.dn[0], 1u);
            if (i >= ghost_floor) { atomicSub(&D.dn[0], 1u); atomicAdd(&D.counters[RB_C_IMM_DROPPED], 1ull); return; }      // (sticky: see the single-sphere branch)
            const float4 pos = r[0], vel = r[1], mt = r[2], pt = r[3], h = r[4];
            const uint32_t flags = __float_as_uint(h.x), gid = __float_as_uint(h.y), cnt = __float_as_uint(h.z);
            P.pos[i] = pos; P.vel[i] = vel; P.mt[i] = mt; P.pt[i] = pt; P.flags[i] = flags; D.gid[i] = gid; P.body_cnt[i] = cnt;
            if (P.force) P.force[i] = r[5];
            if (P.avel) { P.torque[i] = r[6]; P.apos[i] = r[7]; P.avel[i] = r[8]; }
            const float speed = mag3(xyz(vel));
            for (uint32_t j = 0; j < K; ++j) {
                const uint32_t s = i * K + j;
                uint32_t k = RB_NONE;
                if (j < cnt) {
                    // the world sphere K1 computed on the old owner (same state, same formula)
                    const float4 o = r[9 + j];
                    D.sph_obj_w[s] = o;
                    if (P.ce_build) P.ce_build[s].w = -1.0f;      // a new row: no persistent triangle list yet
                    const f3 c = mk3(o.x + mt.x, o.y + mt.y, o.z + mt.z);
                    const float margin = fminf(P.margin_cap, fmaxf(2.0f * speed * P.dt + 0.01f * o.w, 1e-3f));
                    const bool in = (flags & RB_F_ACTIVE) && sphere_in_world(c, o.w, P.half);
                    P.ws[s] = make_float4(c.x, c.y, c.z, in ? o.w + margin : -1.0f);
                    if (in) { k = rb_sphere_col(P, c.x, c.z); P.rank[s] = atomicAdd(&P.col_count[k], 1u); }
                }
                P.key[s] = k;
            }
            return;
        }
        const uint32_t t = (blockIdx.x - imm_blocks) * blockDim.x + threadIdx.x;
        const uint32_t n0 = min(D.in_hdr[0]->n_ghost, D.ghost_cap), n1 = min(D.in_hdr[1]->n_ghost, D.ghost_cap), n2 = min(*D.self_cnt, D.ghost_cap);
        if (t >= n0 + n1 + n2) return;
        const uint8_t* base = reinterpret_cast<const uint8_t*>(t < n0 ? D.in_ghost[0] : (t < n0 + n1 ? D.in_ghost[1] : D.self_ghost));
        const float4* r = reinterpret_cast<const float4*>(base + (size_t)(t < n0 ? t : (t < n0 + n1 ? t - n0 : t - n0 - n1)) * D.ghost_stride);
        const uint32_t gi = atomicAdd(&D.dn[1], 1u);
        const uint32_t i = D.cap_local - 1u - gi;
        const float4 h0 = r[0], h1 = r[1];
        const uint32_t cnt = __float_as_uint(h1.w);
        P.mt[i] = make_float4(h0.x, h0.y, h0.z, 0.f); P.pt[i] = make_float4(h1.x, h1.y, h1.z, 0.f);
        P.pos[i] = make_float4(h0.x, h0.y, h0.z, 1.f);
        D.gid[i] = __float_as_uint(h0.w); P.body_cnt[i] = cnt; P.flags[i] = RB_F_ACTIVE;
        for (uint32_t j = 0; j < K; ++j) {
            const uint32_t s = i * K + j;
            uint32_t k = RB_NONE;
            if (j < cnt) {
                const float4 w = r[3 + 2 * j];
                D.sph_obj_w[s] = r[2 + 2 * j]; P.ws[s] = w;
                if (w.w >= 0.f) { k = rb_sphere_col(P, w.x, w.z); P.rank[s] = atomicAdd(&P.col_count[k], 1u); }
            }
            P.key[s] = k;
        }
        return;
    }
    if (blockIdx.x < imm_blocks) {
        uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t n0 = min(D.in_hdr[0]->n_migr, D.migr_cap);
        uint32_t n1 = min(D.in_hdr[1]->n_migr, D.migr_cap);
        if (t >= n0 + n1) return;
        const RbMigrRec m = t < n0 ? D.in_migr[0][t] : D.in_migr[1][t - n0];
        uint32_t i = atomicAdd(&D.dn[0], 1u);
        // no room: the body is lost to the simulation (its old owner already removed it). Counted in a STICKY counter the
        // per-tick clears never touch; the next rb_step / rb_step_group / rb_get_state / rb_get_stats fails with RB_ERR_CAPACITY.
        if (i >= ghost_floor) { atomicSub(&D.dn[0], 1u); atomicAdd(&D.counters[RB_C_IMM_DROPPED], 1ull); return; }
        P.pos[i] = m.pos; P.vel[i] = m.vel; P.mt[i] = m.mt; P.pt[i] = m.pt; P.flags[i] = m.flags;
        if (P.force) P.force[i] = m.force;
        if (P.avel) { P.torque[i] = m.torque; P.apos[i] = m.apos; P.avel[i] = m.avel; }
        D.sph_obj_w[i] = m.obj; D.gid[i] = m.gid;
        // same world sphere K1 computed on the old owner (same state, same formula)
        f3 c = mk3(m.obj.x + m.mt.x, m.obj.y + m.mt.y, m.obj.z + m.mt.z);
        float speed = mag3(xyz(m.vel));
        float margin = fminf(P.margin_cap, fmaxf(2.0f * speed * P.dt + 0.01f * m.obj.w, 1e-3f));
        P.ws[i] = make_float4(c.x, c.y, c.z, m.obj.w + margin);
        uint32_t k = RB_NONE;
        if ((m.flags & RB_F_ACTIVE) && sphere_in_world(c, m.obj.w, P.half)) {
            k = rb_sphere_col(P, c.x, c.z);
            if (!list_tick) P.rank[i] = atomicAdd(&P.col_count[k], 1u);
        }
        P.key[i] = k;
        if (list_tick) { P.vl[RB_VL_REBUILD] = 1u; atomicOr(&P.vl[RB_VL_WHY], 8u); }      // new row: the lists are stale
        return;
    }
    uint32_t t = (blockIdx.x - imm_blocks) * blockDim.x + threadIdx.x;
    uint32_t n0 = min(D.in_hdr[0]->n_ghost, D.ghost_cap);
    uint32_t n1 = min(D.in_hdr[1]->n_ghost, D.ghost_cap);
    uint32_t n2 = min(*D.self_cnt, D.ghost_cap);
    if (t >= n0 + n1 + n2) return;
    const RbGhostRec q = t < n0 ? D.in_ghost[0][t] : (t < n0 + n1 ? D.in_ghost[1][t - n0] : D.self_ghost[t - n0 - n1]);
    const uint32_t gi = atomicAdd(&D.dn[1], 1u);
    uint32_t i = D.cap_local - 1u - gi;
    P.ws[i] = q.ws;
    P.pos[i] = make_float4(q.ws.x, q.ws.y, q.ws.z, q.r);      // single-sphere bodies carry their exact world radius in pos.w (the list step reads a partner's there)
    D.sph_obj_w[i] = make_float4(0.f, 0.f, 0.f, q.r);
    D.gid[i] = q.gid;
    P.flags[i] = RB_F_ACTIVE;
    if (list_tick) {
        const uint32_t h = P.hot_cap + gi;          // behind the own hot bodies' entries
        P.hot_row[h] = i;
        hot_link(P, h, q.ws.x, q.ws.z);
        P.key[i] = RB_NONE;
        return;
    }
    uint32_t k = rb_sphere_col(P, q.ws.x, q.ws.z);
    P.key[i] = k;
    P.rank[i] = atomicAdd(&P.col_count[k], 1u);
}

static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// ---- host ---------------------------------------------------------------------------------------
static int dist_alloc(rb_world* w) {
    if (w->dist) return RB_OK;
    if (!w->built || !w->d_dn) return fail(w, RB_ERR_INVALID, "distributed init needs a built world created with rb_config.nranks > 1");
    RbDist* d = new RbDist();
    d->ghost_cap = w->ghost_cap;
    d->migr_cap = std::max<uint32_t>(1024u, w->ghost_cap / 8);
    if (w->P.kstride) { d->ghost_stride = 32u + 32u * w->P.kstride; d->migr_stride = 144u + 16u * w->P.kstride; }
    size_t bytes = sizeof(RbHaloHeader) + (size_t)d->ghost_cap * d->ghost_stride + (size_t)d->migr_cap * d->migr_stride;
    bytes = (bytes + 255) / 256 * 256;
    d->buf_bytes = bytes;
    for (int k = 0; k < 2; k++) {
        uint8_t* o; ALLOC(o, bytes);
        CU(cudaMemsetAsync(o, 0, bytes, w->stream));
        d->out[k].base = o; d->out[k].bytes = bytes;
    }
    d->blob_bytes = 4 * bytes + 256;
    ALLOC(d->blob, d->blob_bytes);
    CU(cudaMemsetAsync(d->blob, 0, d->blob_bytes, w->stream));
    // per-tick counters sit behind {n_owned, n_ghost} so ONE memset clears them: [2] self ghosts, [3] emigrants, [4..7] / [8..11] out headers
    RbGhostRec* sg; uint32_t *em; { uint8_t* sgb; ALLOC(sgb, (size_t)d->ghost_cap * d->ghost_stride); sg = (RbGhostRec*)sgb; } ALLOC(em, d->migr_cap * 2 + 16);
    uint32_t* cnts = w->d_dn + 2;
    CU(cudaMallocHost((void**)&d->h_dn, 4 * sizeof(uint32_t)));
    CU(cudaMallocHost((void**)&d->h_hdr, 4 * sizeof(RbHaloHeader)));
    RbDistParams& D = d->D;
    memset(&D, 0, sizeof D);
    for (int k = 0; k < 2; k++) {
        D.out_hdr[k] = (RbHaloHeader*)(w->d_dn + 4 + 4 * k); D.out_ghost[k] = d->out[k].ghosts(); D.out_migr[k] = d->out[k].migr(d->ghost_cap, d->ghost_stride);
        RbHaloBuf ib = d->in_buf(k, 0);
        D.in_hdr[k] = ib.hdr(); D.in_ghost[k] = ib.ghosts(); D.in_migr[k] = ib.migr(d->ghost_cap, d->ghost_stride);
    }
    D.self_ghost = sg; D.self_cnt = cnts; D.emig = em; D.emig_cnt = cnts + 1;
    D.ghost_cap = d->ghost_cap; D.migr_cap = d->migr_cap; D.cap_local = w->cap_local;
    D.slab_lo = w->cfg.slab_lo; D.slab_hi = w->cfg.slab_hi;
    D.has_left = w->cfg.rank > 0; D.has_right = w->cfg.rank < w->cfg.nranks - 1;
    // list ticks: a body that has crossed a face stays with its owner -- as the neighbour's ghost, the ghost reach is wider by the
    // same band -- until the next rebuild the host schedules anyway, unless it is deeper in than the band (one margin cap = half a
    // radius). RB_MIG_BAND=<caps> widens it (development switch): 4 caps measured neutral on C3 / C5 at N = 2 (1.494 against
    // 1.504 ms for C5, 50 % more ghost records) -- migrations are not what makes slab worlds rebuild more often than single ones
    float band_factor = 1.0f;
    { const char* v = getenv("RB_MIG_BAND"); if (v && *v) band_factor = (float)atof(v); }
    D.mig_band = w->vl_on ? band_factor * w->P.margin_cap : 0.0f;
    D.ghost_stride = d->ghost_stride; D.migr_stride = d->migr_stride; D.ext_max = w->ext_max;
    D.dn = w->d_dn; D.gid = w->d_gid; D.sph_obj_w = const_cast<float4*>(w->P.sph_obj); D.counters = w->P.counters;
    ALLOC(d->d_dp[0], 1); ALLOC(d->d_dp[1], 1);
    CU(cudaMemcpyAsync(d->d_dp[0], &D, sizeof D, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(d->d_dp[1], &D, sizeof D, cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->P.dp = d->d_dp[0];
    w->dist = d;
    return RB_OK;
}

int rb_dist_rows_moved(rb_world* w) {
    if (!w->dist) return RB_OK;
    RbDistParams& D = w->dist->D;
    D.gid = w->d_gid; D.sph_obj_w = const_cast<float4*>(w->P.sph_obj);
    w->dist->params_valid = false;
    return RB_OK;
}

void rb_dist_destroy(RbDist* d) {
    if (!d) return;
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    for (int k = 0; k < 2; k++) if (d->peer_opened[k] && d->peer_blob[k]) cudaIpcCloseMemHandle(d->peer_blob[k]);
    if (d->h_dn) cudaFreeHost(d->h_dn);
    if (d->h_hdr) cudaFreeHost(d->h_hdr);
    delete d;
}

int rb_dist_refresh_counts(rb_world* w) {
    if (!w->d_dn) return RB_OK;
    uint32_t tmp[4];
    CU(cudaMemcpyAsync(tmp, w->d_dn, sizeof tmp, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->n_owned_host = tmp[0];
    if (w->dist) {
        RbDist* d = w->dist;
        CU(cudaMemcpy(&d->h_hdr[0], w->d_dn + 4, 2 * sizeof(RbHaloHeader), cudaMemcpyDeviceToHost));      // both out headers (tick block)
        const int lp = (d->mode == RB_X_P2P || d->mode == RB_X_LOCAL) ? (int)((d->halo_step + 1) & 1) : 0;   // parity of the last tick
        CU(cudaMemcpy(&d->h_hdr[2], d->in_buf(0, lp).base, sizeof(RbHaloHeader), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&d->h_hdr[3], d->in_buf(1, lp).base, sizeof(RbHaloHeader), cudaMemcpyDeviceToHost));
        if (d->mode == RB_X_P2P) {
            uint32_t to = 0; CU(cudaMemcpy(&to, d->timeout_flag(), 4, cudaMemcpyDeviceToHost));
            if (to) return fail(w, RB_ERR_NCCL, "halo exchange timed out waiting for a neighbour's arrival flag");
        }
        w->dist_stats.halo_sent_last = d->h_hdr[0].n_ghost + d->h_hdr[1].n_ghost;
        w->dist_stats.halo_recv_last = (w->dist->D.has_left ? d->h_hdr[2].n_ghost : 0) + (w->dist->D.has_right ? d->h_hdr[3].n_ghost : 0);
        w->dist_stats.migrated_out_last = d->h_hdr[0].n_migr + d->h_hdr[1].n_migr;
        w->dist_stats.migrated_in_last = (w->dist->D.has_left ? d->h_hdr[2].n_migr : 0) + (w->dist->D.has_right ? d->h_hdr[3].n_migr : 0);
        if (d->h_hdr[0].ghost_overflow || d->h_hdr[1].ghost_overflow || d->h_hdr[0].migr_overflow || d->h_hdr[1].migr_overflow)
            return fail(w, RB_ERR_CAPACITY, "halo buffer overflow (ghost cap %u, migrant cap %u)", d->ghost_cap, d->migr_cap);
    }
    // sticky counters: an overflow of ANY earlier tick (the per-tick headers above only show the last one)
    unsigned long long st[2] = { 0, 0 };
    CU(cudaMemcpy(st, w->P.counters + RB_C_IMM_DROPPED, sizeof st, cudaMemcpyDeviceToHost));
    return rb_dist_sticky_error(w, st[0], st[1]);
}
int rb_dist_sticky_error(rb_world* w, unsigned long long imm_dropped, unsigned long long halo_ovf) {
    if (imm_dropped) return fail(w, RB_ERR_CAPACITY, "%llu immigrant bodies were dropped: the slab world is full (capacity %u rows); bodies have been lost", imm_dropped, w->cap_local);
    if (halo_ovf) return fail(w, RB_ERR_CAPACITY, "%llu ghost / migrant records did not fit the halo buffers (ghost cap %u): cross-face contacts may have been missed; raise rb_config.halo_ghost_cap", halo_ovf, w->ghost_cap);
    return RB_OK;
}

// device parameter blocks, one per tick parity: where K1 stores the records (local staging, or the neighbour's
// in buffer of that parity) and which in buffers the unpack kernel reads. Rebuilt only when the transport
// connects or the body rows move.
static int dist_upload_params(rb_world* w) {
    RbDist* d = w->dist;
    const bool remote = d->mode == RB_X_P2P || d->mode == RB_X_LOCAL;
    for (int p = 0; p < 2; p++) {
        RbDistParams D = d->D;
        for (int k = 0; k < 2; k++) {
            if (remote && d->mode == RB_X_LOCAL && d->peer[k] && !d->peer_blob[k]) d->peer_blob[k] = d->peer[k]->dist->blob;
            if (remote && d->peer_blob[k]) {
                RbHaloBuf pb = RbDist::buf_at(d->peer_blob[k], d->buf_bytes, 1 - k, p);      // my left neighbour receives on ITS right side
                D.out_ghost[k] = pb.ghosts(); D.out_migr[k] = pb.migr(d->ghost_cap, d->ghost_stride);
            } else { D.out_ghost[k] = d->out[k].ghosts(); D.out_migr[k] = d->out[k].migr(d->ghost_cap, d->ghost_stride); }
            RbHaloBuf ib = d->in_buf(k, remote ? p : 0);
            D.in_hdr[k] = ib.hdr(); D.in_ghost[k] = ib.ghosts(); D.in_migr[k] = ib.migr(d->ghost_cap, d->ghost_stride);
        }
        d->Dp[p] = D;
        CU(cudaMemcpyAsync(d->d_dp[p], &d->Dp[p], sizeof D, cudaMemcpyHostToDevice, w->stream));
    }
    d->params_valid = true;
    return RB_OK;
}
// phase A: integrate + bin + halo pack (everything before the exchange)
// list_mode: 0 = per-tick enumeration (K1 bins), 2 = persistent-list tick, 3 = the same with a rebuild forced by the host
static int dist_phase_a(rb_world* w, int do_integrate, int list_mode) {
    RbDist* d = w->dist; RbParams& P = w->P;
    if (!d->params_valid) { int r = dist_upload_params(w); if (r) return r; }
    const bool remote = d->mode == RB_X_P2P || d->mode == RB_X_LOCAL;
    const int p = remote ? (int)(d->halo_step & 1) : 0;
    P.dp = d->d_dp[p];
    // clears as kernels (a memset may land on a copy engine and queue behind a host transfer)
    if (list_mode != 2) { rb_launch_clear(P.col_count, w->hist_clear_bytes, nullptr, 0, w->stream); w->launches++; }      // (list ticks leave the histogram clean themselves)
    if (list_mode) {   // contact + hot counters and the rebuild flag (5 adjacent words, see rb_build) | the tick block of the halo step
        rb_list_tick_params(w);
        rb_launch_clear(w->d_tb + 4 * (RB_TB_BLOCKS - 1), 20, w->d_dn + 1, 11, w->stream); w->launches++;
    } else {
        int r = rb_plain_tick_params(w); if (r) return r;
        rb_launch_clear(nullptr, 0, w->d_dn + 1, 11, w->stream); w->launches++;            // n_ghost, self ghosts, emigrants, both out headers
        rb_launch_clear(nullptr, 0, (uint32_t*)(P.counters + RB_C_NCONTACT), 6, w->stream); w->launches++;
    }
    if (w->profile) CU(cudaEventRecord(w->ev[0], w->stream));
    rb_launch_integrate_bin(P, do_integrate, list_mode ? list_mode : 1, w->stream); w->launches++;     // integrate + bin + halo pack
    if (w->profile) CU(cudaEventRecord(w->ev[5], w->stream));
    // headers (and, on the fused transport, the arrival flags) go out
    RbHaloHeader* ph[2] = { d->out[0].hdr(), d->out[1].hdr() }; uint32_t* pa[2] = { nullptr, nullptr };
    if (remote) for (int k = 0; k < 2; k++) {
        ph[k] = nullptr;
        if (d->peer_blob[k]) { ph[k] = RbDist::buf_at(d->peer_blob[k], d->buf_bytes, 1 - k, p).hdr(); pa[k] = RbDist::arrive_at(d->peer_blob[k], d->buf_bytes, 1 - k); }
    }
    rb_halo_signal_kernel<<<1, 32, 0, w->stream>>>(d->Dp[p].out_hdr[0], d->Dp[p].out_hdr[1], ph[0], ph[1], pa[0], pa[1], d->halo_step); w->launches++;
    // list ticks the host does not turn into rebuilds: the blocks no ghost can reach step now, while the neighbours' records
    // are still on their way (phase B waits for them)
    w->halo_split = false;
    if (list_mode == 2 && P.blk_bnd) { rb_launch_list_step_early(P, w->stream, w->vl_side, w->vl_fork, w->vl_join, rb_list_hot_rows(w)); w->launches += 2; w->halo_split = true; }
    return RB_OK;
}
// phase B: unpack + scan + scatter (everything after the exchange)
static int dist_phase_b(rb_world* w, int list_mode) {
    RbDist* d = w->dist; RbParams& P = w->P;
    const bool remote = d->mode == RB_X_P2P || d->mode == RB_X_LOCAL;
    const int p = remote ? (int)(d->halo_step & 1) : 0;
    const bool wait = d->mode == RB_X_P2P;
    rb_halo_remove_kernel<<<1, 32, 0, w->stream>>>(P, d->Dp[p], wait && d->peer_blob[0] ? d->arrive(0) : nullptr, wait && d->peer_blob[1] ? d->arrive(1) : nullptr, d->halo_step, d->timeout_flag(), list_mode ? 1 : 0);
    const uint32_t imm_blocks = cdiv(2 * d->migr_cap, RB_BLOCK);
    rb_halo_unpack_kernel<<<imm_blocks + cdiv(3 * d->ghost_cap, RB_BLOCK), RB_BLOCK, 0, w->stream>>>(P, d->Dp[p], imm_blocks, list_mode ? 1 : 0);
    w->launches += 2;
    CU(cudaGetLastError());
    d->halo_step++;
    return list_mode ? RB_OK : enqueue_broadphase_tail(w);       // (list ticks: the caller continues with the conditional rebuild chain)
}

static int dist_exchange_nccl(rb_world* w) {
    RbDist* d = w->dist;
    int rank = w->cfg.rank, n = w->cfg.nranks;
    NC(g_nccl.GroupStart());
    if (rank > 0) { NC(g_nccl.Send(d->out[0].base, d->out[0].bytes, NCCL_UINT8, rank - 1, d->comm, w->stream)); NC(g_nccl.Recv(d->in_buf(0, 0).base, d->buf_bytes, NCCL_UINT8, rank - 1, d->comm, w->stream)); }
    if (rank < n - 1) { NC(g_nccl.Send(d->out[1].base, d->out[1].bytes, NCCL_UINT8, rank + 1, d->comm, w->stream)); NC(g_nccl.Recv(d->in_buf(1, 0).base, d->buf_bytes, NCCL_UINT8, rank + 1, d->comm, w->stream)); }
    NC(g_nccl.GroupEnd());
    return RB_OK;
}

int rb_dist_step_exchange(rb_world* w, int do_integrate) {
    RbDist* d = w->dist;
    if (d->mode != RB_X_NCCL && d->mode != RB_X_P2P) return fail(w, RB_ERR_INVALID, "this slab world has in-process neighbours: step it with rb_step_group");
    int r = dist_phase_a(w, do_integrate, 0); if (r) return r;
    if (d->mode == RB_X_NCCL) { r = dist_exchange_nccl(w); if (r) return r; }
    return dist_phase_b(w, 0);
}
// the halo part of a persistent-list tick (K1 with the list checks + pack, exchange, remove + unpack)
int rb_dist_list_exchange(rb_world* w, int list_mode) {
    RbDist* d = w->dist;
    if (d->mode != RB_X_NCCL && d->mode != RB_X_P2P) return fail(w, RB_ERR_INVALID, "this slab world has in-process neighbours: step it with rb_step_group");
    int r = dist_phase_a(w, 1, list_mode); if (r) return r;
    if (d->mode == RB_X_NCCL) { r = dist_exchange_nccl(w); if (r) return r; }
    return dist_phase_b(w, list_mode);
}

extern "C" {

int rb_dist_unique_id(const char* nccl_lib_path, uint8_t id[RB_NCCL_ID_BYTES]) {
    int r = load_nccl(nullptr, nccl_lib_path); if (r) return r;
    ncclUniqueId_t u; memset(&u, 0, sizeof u);
    int e = g_nccl.GetUniqueId(&u);
    if (e) { g_create_error = std::string("ncclGetUniqueId failed: ") + g_nccl.GetErrorString(e); return RB_ERR_NCCL; }
    memcpy(id, &u, RB_NCCL_ID_BYTES);
    return RB_OK;
}

int rb_dist_init(rb_world* w, const char* nccl_lib_path, const uint8_t id[RB_NCCL_ID_BYTES]) {
    if (!w) return RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    int r = load_nccl(w, nccl_lib_path); if (r) return r;
    r = dist_alloc(w); if (r) return r;
    ncclUniqueId_t u; memcpy(&u, id, RB_NCCL_ID_BYTES);
    NC(g_nccl.CommInitRank(&w->dist->comm, w->cfg.nranks, u, w->cfg.rank));
    w->dist->mode = RB_X_NCCL;
    return RB_OK;
}

// In-process neighbours (left / right may be NULL at the world edge): the slab worlds of one process,
// on one or several devices, stepped in lockstep by rb_step_group. Used by the single-GPU tests of the
// multi-rank path and by single-process multi-GPU hosts.
int rb_dist_attach_peers(rb_world* w, rb_world* left, rb_world* right) {
    if (!w) return RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    int r = dist_alloc(w); if (r) return r;
    w->dist->peer[0] = left; w->dist->peer[1] = right;
    for (rb_world* pw : { left, right })
        if (pw && pw->ghost_cap != w->ghost_cap)
            return fail(w, RB_ERR_INVALID, "halo buffer layouts differ between slab worlds (ghost capacity %u vs %u): pass the same rb_config.halo_ghost_cap to all of them", w->ghost_cap, pw->ghost_cap);
    w->dist->mode = RB_X_LOCAL;
    return RB_OK;
}

// Fused compute + communication transport (RB_X_P2P): every rank exports the IPC handle of its halo blob,
// the host hands each rank its neighbours' handles, and from then on K1 stores straight into them.
int rb_dist_p2p_export(rb_world* w, uint8_t handle[RB_IPC_HANDLE_BYTES]) {
    if (!w || !handle) return RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    int r = dist_alloc(w); if (r) return r;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, w->dist->blob));
    static_assert(sizeof(h) + 16 == RB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle, &h, sizeof h);
    // the layout the neighbours will store into: they must have computed the same one
    const uint32_t caps[2] = { w->dist->ghost_cap, w->dist->migr_cap }; const uint64_t bb = w->dist->buf_bytes;
    memcpy(handle + sizeof h, caps, 8); memcpy(handle + sizeof h + 8, &bb, 8);
    return RB_OK;
}
int rb_dist_p2p_connect(rb_world* w, const uint8_t* left_handle, const uint8_t* right_handle) {
    if (!w || !w->dist) return w ? fail(w, RB_ERR_INVALID, "rb_dist_p2p_connect before rb_dist_p2p_export") : RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    const uint8_t* hs[2] = { left_handle, right_handle };
    for (int k = 0; k < 2; k++) {
        if (!hs[k]) continue;
        cudaIpcMemHandle_t h; memcpy(&h, hs[k], sizeof h);
        uint32_t caps[2]; uint64_t bb; memcpy(caps, hs[k] + sizeof h, 8); memcpy(&bb, hs[k] + sizeof h + 8, 8);
        if (caps[0] != w->dist->ghost_cap || caps[1] != w->dist->migr_cap || bb != w->dist->buf_bytes)
            return fail(w, RB_ERR_INVALID, "halo buffer layouts differ between ranks (ghost capacity %u here, %u at the %s neighbour): pass the same rb_config.halo_ghost_cap on every rank",
                        w->dist->ghost_cap, caps[0], k ? "right" : "left");
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        w->dist->peer_blob[k] = (uint8_t*)p; w->dist->peer_opened[k] = true;
    }
    w->dist->mode = RB_X_P2P;
    return RB_OK;
}

static int group_collide(rb_world** ws, int n, int ms, int do_integrate, int resolve);

int rb_step_group(rb_world** ws, int n, int ms) { return group_collide(ws, n, ms, 1, 1); }
int rb_detect_group(rb_world** ws, int n) { return group_collide(ws, n, 0, 0, 0); }

} // extern "C"

int rb_collide_only(rb_world* w, bool resolve);   // rb_world.cu

static int group_collide(rb_world** ws, int n, int ms, int do_integrate, int resolve) {
    for (int i = 0; i < n; i++) {
        rb_world* w = ws[i];
        if (!w || !w->dist) return w ? fail(w, RB_ERR_INVALID, "rb_step_group: world %d has no peers attached", i) : RB_ERR_INVALID;
        CU(cudaSetDevice(w->cfg.device));
        if (do_integrate) w->P.dt = (float)ms / 1000.0f;
        // persistent-list tick? (decided per world; stepping only)
        w->group_list_mode = 0;
        if (do_integrate && resolve && rb_step_will_use_lists(w)) {
            const bool reorder_due = w->reorder_every > 0 && (w->steps + 1) % (uint64_t)w->reorder_every == 0;
            w->group_list_mode = rb_list_tick_begin(w, reorder_due) ? 3 : 2;
        }
        if (w->vl_on && !w->group_list_mode) CU(cudaMemsetAsync(w->P.vl + RB_VL_NFAST, 0, 4, w->stream));     // K1 counts the fast movers while the lists sit a phase out
        int r = dist_phase_a(w, do_integrate, w->group_list_mode); if (r) return r;
        if (w->group_list_mode) rb_list_tick_k1_enqueued(w);
    }
    for (int i = 0; i < n; i++) { rb_world* w = ws[i]; CU(cudaSetDevice(w->cfg.device)); CU(cudaStreamSynchronize(w->stream)); }
    // (the records and headers already sit in the neighbours' in buffers: K1 and the signal kernel stored them there)
    for (int i = 0; i < n; i++) {
        rb_world* w = ws[i];
        CU(cudaSetDevice(w->cfg.device));
        int r = dist_phase_b(w, w->group_list_mode); if (r) return r;
        if (w->group_list_mode) {
            const bool reorder_due = w->reorder_every > 0 && (w->steps + 1) % (uint64_t)w->reorder_every == 0;
            r = rb_list_tick_tail(w, reorder_due); if (r) return r;
            w->steps++; w->pose_updates++;
            continue;
        }
        r = rb_collide_only(w, resolve != 0); if (r) return r;
        if (w->vl_on) w->vl_force = true;
        if (do_integrate) {
            w->steps++;
            if (w->reorder_every > 0 && w->steps % (uint64_t)w->reorder_every == 0) { r = rb_reorder_rows_now(w); if (r) return r; }
        }
    }
    return RB_OK;
}
