// rb_device.cuh -- shared declarations of the B200 physics-step kernels (sm_100a).
//
// Arithmetic contract: the whole library is compiled with -fmad=false and nvcc's default
// -prec-div=true -prec-sqrt=true -ftz=false, so every FP32 expression below rounds exactly like the
// reference's x86-64 SSE code (g++ -O2, no FMA). Expression ORDER therefore matters and follows
// the reference: dot = (x*x' + y*y') + z*z' (math/src/Vector4.cpp:119-126), cross as in
// Vector4.cpp:106-117, magnitude = sqrtf((x^2 + y^2) + z^2) (Vector4.cpp:34-38).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RB_F_ACTIVE 1u
#define RB_F_GRAVITY 2u
#define RB_F_CONTACT 4u
#define RB_NONE 0xFFFFFFFFu
#define RB_TRI_BIT 0x80000000u
#define RB_HOT_BIT 0x80000000u   // in RbParams::key: the body steps without its list this tick

// words of RbParams::vl
#define RB_VL_REBUILD 0   // this tick rebuilds the lists (set by K1, cleared by the resolve kernel)
#define RB_VL_WHY     1   // why: 1 a body entered / left the grid (or emigrated), 2 hot table full, 4 host (first tick, rows moved, state edited), 8 immigrants
#define RB_VL_WHY_LAST 4  // the reasons of the most recent rebuild
#define RB_VL_NWHY    8   // 4 words: rebuilds so far by reason bit (grid entry/exit, hot table full, host, immigrants)
#define RB_VL_WORDS   24
#define RB_VL_NOLIST  16  // 3 words: rows the last rebuild left without a list -- more partners than a list holds, a sphere wider than the fast triangle walk, a block whose tile overflowed
#define RB_VL_STORM   6   // the previous list tick rebuilt too
#define RB_VL_NFAST   5   // per-tick kernel only: bodies whose one-tick margin alone exceeds the skin (saturates at hot_cap)
#define RB_VL_NREBUILD 2  // rebuild ticks so far
#define RB_VL_NTICK   3   // list ticks so far
#define RB_VL_NHOT_LAST 7  // hot bodies of the most recent list tick
#define RB_VL_AGE     12  // list ticks since the lists were built (0 on the rebuild tick)
#define RB_VL_EXPIRY  13  // the age at which more than hot_soft bodies were hot for the first time (0: not yet) -- the lists' useful life
#define RB_VL_GEN     14  // list generations so far (rebuild ticks of list ticks)
#define RB_VL_FLOOR   15  // hot bodies on the rebuild tick itself: rows that can have no list (crowded, oversized) -- a rebuild does not cure them

// device contact record, 32 B (two 16-B stores)
struct __align__(16) RbContactRec {
    uint32_t a;        // global sphere id
    uint32_t b;        // global sphere id, or RB_TRI_BIT | global triangle id
    float depth, nx, ny, nz;
    float r;           // world radius of sphere a at the time of the hit
    uint32_t pad;
};

// counters[] layout (unsigned long long)
enum { RB_C_TESTS_SS = 0, RB_C_TESTS_ST, RB_C_CAND_SS, RB_C_CAND_ST, RB_C_HITS_SS, RB_C_HITS_ST,
       RB_C_OVERFLOW, RB_C_DROPPED, RB_C_NCONTACT /* per step, reset */, RB_C_NCHANGED /* per step */,
       RB_C_NSS /* per step: compound bodies with sphere-sphere partners (split kernels) */, RB_C_MIGRATED,
       RB_C_IMM_DROPPED /* sticky: immigrants a full slab could not take */, RB_C_HALO_OVF /* sticky: ghost / migrant records that did not fit a halo buffer */,
       RB_C_COUNT = 16 };

// ---- multi-GPU halo exchange records (rb_dist.cu) ------------------------------------------------
struct __align__(16) RbHaloHeader { uint32_t n_ghost, n_migr, ghost_overflow, migr_overflow; };
struct __align__(16) RbGhostRec { float4 ws; float r; uint32_t gid; uint32_t pad0, pad1; };                 // 32 B
struct __align__(16) RbMigrRec { float4 pos, vel, mt, pt, obj; uint32_t flags, gid, pad0, pad1;            // 160 B
                                 float4 force, torque, apos, avel; };      // StateVector's force | mass, torque, angular position / velocity (defaults when the world has none)

struct RbDistParams {
    RbHaloHeader* out_hdr[2]; RbGhostRec* out_ghost[2]; RbMigrRec* out_migr[2];      // 0 = left, 1 = right
    const RbHaloHeader* in_hdr[2]; const RbGhostRec* in_ghost[2]; const RbMigrRec* in_migr[2];
    RbGhostRec* self_ghost; uint32_t* self_cnt;     // emigrants stay visible to their old owner for this tick
    uint32_t* emig; uint32_t* emig_cnt;             // local indices of emigrants
    uint32_t ghost_cap, migr_cap, cap_local;
    float slab_lo, slab_hi; int has_left, has_right;
    float mig_band;               // list ticks: a body up to this far beyond a face stays with its owner until the next scheduled rebuild
    // compound bodies (P.kstride): records are byte strings of a per-world size -- a ghost body is {mt | gid, pt | spheres,
    // spheres x {obj, ws}}, a migrant {pos, vel, mt, pt, {flags, gid, spheres, -}, force, torque, apos, avel, spheres x obj} -- and a body reaches as far
    // from its model translation as its farthest sphere: ext_max is the largest such reach (grown by the margin cap) anywhere
    uint32_t ghost_stride, migr_stride;
    float ext_max;
    uint32_t* dn;                                   // {n_owned, n_ghost}
    uint32_t* gid; float4* sph_obj_w; unsigned long long* counters;
};


#define RB_STRIPES 64

struct RbParams {
    uint32_t nb, ns, nt;
    uint32_t single;              // every body has exactly one sphere (sphere id == body id)
    // ---- bodies (SoA float4) ----
    float4* pos;                  // linear position xyz
    float4* vel;                  // linear velocity xyz
    float4* mt;                   // model-matrix translation (Entity::_mvp model [3],[7],[11])
    float4* pt;                   // previous model-matrix translation (Entity::_prevMVP)
    uint32_t* flags;
    const float4* wt;             // translation of _worldSpaceTransform, nullable (all zero)
    float4* force;                // nullable: xyz force, w mass
    float4* torque;               // nullable
    float4* apos;                 // nullable angular position
    float4* avel;                 // nullable angular velocity
    // ---- spheres ----
    const float4* sph_obj;        // xyz = W3x3 * object centre (pre-summed), w = W[0] * object radius
    const uint32_t* sph_body;     // nullable when single
    const uint32_t* body_start;   // nullable when single (CSR body -> spheres)
    float4* ws;                   // frozen world sphere after integration: xyz centre, w = radius + margin
    uint32_t* key;                // sphere-grid column of each sphere (RB_NONE = not in the broadphase)
    uint32_t* rank;               // arrival rank inside the column
    // ---- sphere column grid (rebuilt every step) ----
    uint32_t* col_count;          // [ncols + 1]
    uint32_t* col_start;          // [ncols + 1] exclusive scan
    uint32_t scx, scz;
    float s_inv_w;
    float4* sorted;               // ws in column order
    uint32_t* sorted_id;          // sphere id in column order
    // ---- static triangle column grid (built once) ----
    const float4* tri;            // 3 float4 per triangle: A, B, C (w unused)
    const uint32_t* tcol_start;   // [tcols + 1]
    const uint32_t* tcol_tri;     // ascending triangle ids per column
    const float2* tcol_yr;        // per column: (min y, max y) over its triangles (empty: +inf, -inf)
    const float4* tcol_ent;       // 2 float4 per column entry: (minx,maxx,minz,maxz) (miny,maxy,id bits,0)
    uint32_t tcx, tcz;
    float t_inv_w;
    // ---- constants ----
    float half, gravity, friction, pushout, ss_damp, margin_cap, rm_max, dt;
    // ---- outputs ----
    RbContactRec* contacts;
    unsigned long long contact_cap;
    unsigned long long* counters;
    unsigned long long* stripes;  // RB_STRIPES x RB_C_COUNT striped copies of the cumulative test / hit counters
    // ---- host round trip fused into the tick (list ticks of rb_step_host_async*): K1 takes this tick's forces
    // straight from the uploaded id-ordered array, the resolve kernels leave the final positions in id order ----
    const float* force_in; uint32_t force_in_stride;   // NULL: forces come from P.force as usual
    float* pos_out; uint32_t pos_out_stride;            // NULL: nobody asked
    // ---- candidate lists, indexed by body row (single-sphere worlds) ----
    uint32_t* e_hdr;              // per row: partners (bits 0-7) | triangles (8-15) | offset of its first triangle in the block's tile (16-31)
    uint32_t* e_pg; uint32_t* e_pl;   // [RB_SSCAND_MAX][ns] partner global id / local row (slot-major)
    // Triangle tiles: the rows of one block of the list-step kernel (256 consecutive rows) keep their candidate triangles
    // -- the 48-B vertex records themselves, in row order, each row's ascending by global id -- in ONE contiguous region of
    // the pool, followed by the global triangle ids. The kernel stages the region in shared memory with a single bulk
    // copy (TMA, cp.async.bulk) instead of gathering 48 B per test through the cache hierarchy.
    uint4* tile_pool;             // 16-B units; a region of n entries is 3 n units of vertices + ceil(n / 4) units of ids
    uint2* blk_tile;              // per block: { first unit of its region, entries }
    uint32_t* tile_cursor;        // next free unit (reset when a rebuild starts)
    uint32_t tile_pool_units;     // capacity of the pool
    // slab worlds: the list step of a tick runs in two launches, the blocks none of whose rows can reach a neighbour's sphere
    // BEFORE the halo step waits for the neighbours (their records arrive meanwhile), everybody else after it
    uint32_t* blk_bnd;            // per block: the tick in which K1 last saw one of its rows within reach of a slab face (or leaving)
    uint32_t* blk_done;           // per block: the tick in which the early launch stepped it
    // ---- persistent lists ("Verlet lists"): the lists above are built for spheres grown by vl_skin and
    // reused until some body has used its skin up; the sort and the enumeration only run on those ticks ----
    uint32_t* vl;                 // device words, see RB_VL_*
    volatile uint32_t* vl_host;   // mapped host copy of {rebuilds, ticks} (heuristics only)
    float4* ws_build;             // frozen sphere of every row at the last rebuild
    float vl_skin;
    unsigned long long* scan_status; uint32_t* scan_ticket; uint32_t ncols;     // the scan's work words (device-launched rebuild chain)
    float vl_slack;               // added to the grown radius when a list is pruned with the exact predicate (>> its rounding error)
    // "hot" bodies have used their skin up (fast movers): they enumerate for themselves every tick (triangles
    // from the static grid, cold partners from the sort of the last rebuild, which is off by less than the
    // skin) and are published in a coarse (x,z) cell table so that cold bodies and other hot bodies find them
    unsigned long long* hot_head; // [hcn * hcn] chain heads: tick << 32 | index + 1. An entry of an earlier tick is an empty cell, so
                                  // the table is never cleared (the tick number is its version)
    uint32_t* hot_row; uint32_t* hot_next; uint32_t hot_cap;      // hot_cap = every row: the table cannot overflow
    uint32_t hot_soft;            // the number of hot bodies (beyond those that can have no list) at which a rebuild pays (heuristics only)
    uint32_t hot_flood;           // ... at which lists stop paying altogether: a hot body costs a few times a listed one, so about an eighth of the rows
    uint32_t hcn; float hc_inv_w;
    uint32_t tick;                // version of this tick's hot-table entries (never 0)
    uint32_t* nhot;               // this tick's hot-body counter
    unsigned long long* ncontact; // this tick's contact counter (appended records)
    uint32_t* tb_next;            // single worlds: the 4 words {contacts lo, hi, hot bodies, -} of the NEXT tick, zeroed by K1 (no clear launch)
    uint32_t soft;                // single-GPU worlds: the device never asks for a rebuild -- a body that enters the grid steps without a
                                  // list, one that leaves is marked in its world sphere (w < 0), the hot table holds every row -- and
                                  // the host schedules rebuilds from the hot counts it sees
    uint32_t lazy_mt;             // list ticks of worlds with W.t = 0 and centred spheres: the model translation is 0 + p, never loaded or stored
    uint32_t mt_in_valid;         // ... K1: the mt array is valid on entry (first list tick after other activity), else derive it
    uint32_t obj_zero;            // every sphere is centred on its body (sph_obj.xyz = +0): the radius rides in pos.w, sph_obj is never read
    // ---- deferred commit (compound-body path) ----
    uint32_t* ch_body; float4* ch_pos; float4* ch_vel; float4* ch_mt; float4* ch_pt; uint32_t* ch_flags;
    // ---- slab ownership (multi-GPU) ----
    uint32_t n_owned;             // bodies [0, n_owned) are owned (host value; single-GPU worlds)
    const uint32_t* dn;           // device counts {n_owned, n_ghost} of slab worlds (NULL otherwise):
                                  // local bodies [0, dn[0]) are owned, [dn[0], dn[0]+dn[1]) are ghosts
    const uint32_t* gid;          // global body id per local body (NULL: id == local index)
    float s_x0;                   // x origin of the sphere column grid (-half for whole-world grids)
    // the sphere grid's column key is cv * g_nu + cu: u is the axis whose neighbouring columns are contiguous in the sort, v the
    // outer one. Single worlds: (u, v) = (x, z). Slab worlds: (u, v) = (z, x), so that the row order (which follows the sort)
    // runs along x and the rows near the slab faces are the first and the last blocks
    uint32_t g_swap, g_nu, g_nv; float g_off_u, g_off_v;
    const RbDistParams* dp;       // device copy of the halo parameters (slab worlds with neighbours), else NULL
    // ---- compound bodies (multi-sphere entities, physics/src/Geometry.cpp:27-31) ----
    uint32_t grp_on;              // compound bodies: 0 per-body kernel, 1 rb_collide_group_kernel, 2 enumeration per sphere + resolution per body (RB_COMPOUND=legacy|group|split)
    uint32_t* ce_hdr;             // split kernels, per sphere row: triangles (bits 0-7) | RB_CE_SLOW
    uint32_t* ce_ph;              // ... partners (bits 8-15) | RB_CE_OVER
    uint32_t* ce_pg; uint32_t* ce_pl;   // [4][ns] partner body global id / row, ascending (slot-major)
    uint32_t* ce_tri;             // [RB_CAND_MAX][ns] staged triangle ids, ascending (slot-major)
    // the triangle side of those lists is persistent: static geometry does not move, so a sphere's list -- staged for the sphere
    // grown by ce_skin -- stays a superset of the per-tick candidates while displacement since then + this tick's margin <= ce_skin,
    // and every sphere decides that for itself (no partner is involved: nothing like the hot-body table is needed)
    float4* ce_build;             // per sphere row: centre the list was staged at (w < 0 or NaN: no reusable list)
    float ce_skin;
    // bodies with sphere-sphere partners run that phase in a launch of their own (dense warps), their state parked for the triangle phase
    uint32_t* ss_npart; uint32_t* ss_list; uint32_t* ss_slot;   // per body: partners its spheres found (0 = none; reset by the triangle phase) | the list | body -> entry + 1
    float4* ss_pos; float4* ss_vel; float4* ss_mt; float4* ss_pt;   // per entry: state after the sphere-sphere phase (flags in pos.w, "changed" in vel.w)
    uint32_t grp_lg;              // log2 of the lanes one body gets in rb_collide_group_kernel (2..5, from the largest sphere count)
    uint32_t kstride;             // slab worlds: the spheres of body row b sit at rows [b * kstride, b * kstride + body_cnt[b]) (0: CSR through body_start)
    uint32_t* body_cnt;           // kstride worlds: spheres per body row
};

// body -> spheres (CSR in single worlds, fixed stride in slab worlds where rows come and go)
__device__ __forceinline__ uint32_t rb_body_s0(const RbParams& P, uint32_t b) { return P.kstride ? b * P.kstride : P.body_start[b]; }
__device__ __forceinline__ uint32_t rb_body_ns(const RbParams& P, uint32_t b) { return P.kstride ? P.body_cnt[b] : P.body_start[b + 1] - P.body_start[b]; }
__device__ __forceinline__ uint32_t rb_sph_owner(const RbParams& P, uint32_t s) { return P.kstride ? s / P.kstride : P.sph_body[s]; }

__device__ __forceinline__ uint32_t rb_n_owned(const RbParams& P) { return P.dn ? P.dn[0] : P.n_owned; }
__device__ __forceinline__ uint32_t rb_n_local(const RbParams& P) { return P.dn ? P.dn[0] + P.dn[1] : P.ns; }
// non-binding L1 prefetch of the line holding p (no register, no scoreboard entry)
#ifndef RB_PREFETCH_LEVEL
#define RB_PREFETCH_LEVEL 1
#endif
__device__ __forceinline__ void rb_prefetch(const void* p) {
#if RB_PREFETCH_LEVEL == 1
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#elif RB_PREFETCH_LEVEL == 2
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#endif
}
__device__ __forceinline__ uint32_t rb_gid(const RbParams& P, uint32_t local) { return P.gid ? P.gid[local] : local; }
// the id a contact record names sphere i of body row b by: its sphere id in a single world, global body id * stride + i in a slab
// world (rb_query_contacts takes it apart again)
__device__ __forceinline__ uint32_t rb_sph_gid(const RbParams& P, uint32_t b, uint32_t s0, uint32_t i) { return (P.kstride ? rb_gid(P, b) * P.kstride : s0) + i; }
__device__ __forceinline__ float rb_gu(const RbParams& P, float x, float z) { return P.g_swap ? z : x; }
__device__ __forceinline__ float rb_gv(const RbParams& P, float x, float z) { return P.g_swap ? x : z; }

// K7 arguments: every per-body array pair (current rows -> re-ordered rows)
#define RB_REORDER_MAX 12
struct RbReorder {
    const float4* src4[RB_REORDER_MAX]; float4* dst4[RB_REORDER_MAX]; int n4;
    const uint32_t* src_flags; uint32_t* dst_flags;
    const uint32_t* src_gid; uint32_t* dst_gid; uint32_t* inv;
    uint32_t* rest_counter;
    const uint32_t* src_key; uint32_t* dst_key;   // in-tick re-ordering (between the sort and the enumeration of a list rebuild):
    uint32_t* sorted_id_fix;                      // keys travel with the rows and the sorted ids are rewritten to the new rows
    const uint32_t* owned_rank;   // slab worlds: exclusive count of owned entries in front of each sorted slot (+ total at the end)
    uint32_t n_slots;             // capacity of owned_rank - 1
};

struct f3 { float x, y, z; };

__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 mul3(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 neg3(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return mk3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
__device__ __forceinline__ float mag3(f3 a) { return sqrtf((a.x * a.x + a.y * a.y) + a.z * a.z); }
__device__ __forceinline__ f3 normalize3(f3 a) { float m = mag3(a); return mk3(a.x / m, a.y / m, a.z / m); }

// GeometryMath::sphereTriangleDetection (math/src/GeometryMath.cpp:101-209): Ericson's sphere /
// triangle separating-axis test in sphere-centred coordinates; every separation is closed (>=).
__device__ __forceinline__ bool sphere_triangle_detect(f3 P, float r, f3 ta, f3 tb, f3 tc) {
    f3 A = sub3(ta, P), B = sub3(tb, P), C = sub3(tc, P);
    float rr = r * r;
    f3 V = cross3(sub3(B, A), sub3(C, A));
    float d = dot3(A, V), e = dot3(V, V);
    if ((d * d) >= (rr * e)) return false;
    float aa = dot3(A, A), ab = dot3(A, B), ac = dot3(A, C);
    float bb = dot3(B, B), bc = dot3(B, C), cc = dot3(C, C);
    if ((aa >= rr) & (ab >= aa) & (ac >= aa)) return false;
    if ((bb >= rr) & (ab >= bb) & (bc >= bb)) return false;
    if ((cc >= rr) & (ac >= cc) & (bc >= cc)) return false;
    f3 AB = sub3(B, A), BC = sub3(C, B), CA = sub3(A, C);
    float d1 = ab - aa, d2 = bc - bb, d3 = ac - cc;
    float e1 = dot3(AB, AB), e2 = dot3(BC, BC), e3 = dot3(CA, CA);
    f3 Q1 = sub3(mul3(A, e1), mul3(AB, d1));
    f3 Q2 = sub3(mul3(B, e2), mul3(BC, d2));
    f3 Q3 = sub3(mul3(C, e3), mul3(CA, d3));
    f3 QC = sub3(mul3(C, e1), Q1);
    f3 QA = sub3(mul3(A, e2), Q2);
    f3 QB = sub3(mul3(B, e3), Q3);
    if ((dot3(Q1, Q1) >= (rr * e1 * e1)) && (dot3(Q1, QC) >= 0.0f)) return false;
    if ((dot3(Q2, Q2) >= (rr * e2 * e2)) && (dot3(Q2, QA) >= 0.0f)) return false;
    if ((dot3(Q3, Q3) >= (rr * e3 * e3)) && (dot3(Q3, QB) >= 0.0f)) return false;
    return true;
}

// GeometryMath::triangleCubeDetection (math/src/GeometryMath.cpp:233-411): Akenine-Moeller's triangle / AABB separating-axis
// test in box-centred coordinates -- nine cross-product axes, the three box axes, the triangle plane. The reference pairs
// height / 2 with x, width / 2 with y and length / 2 with z (:238-240); strict comparisons (touching = overlap), the plane
// test closed on the far side (:405-409). ctr + (length, height, width) is the reference's Cube.
__host__ __device__ __forceinline__ float rb_fmax2(float a, float b) { return a > b ? a : b; }      // GeometryMath.cpp:556-559
__host__ __device__ __forceinline__ float rb_fmin2(float a, float b) { return a < b ? a : b; }      // :561-564
__device__ __forceinline__ bool tri_cube_axis_separates(f3 v0, f3 v1, f3 v2, f3 axis, float e0, float e1, float e2) {
    const float p0 = dot3(v0, axis), p1 = dot3(v1, axis), p2 = dot3(v2, axis);
    // r = e0 |u0 . a| + e1 |u1 . a| + e2 |u2 . a| with the unit axes spelled out like the reference's dot products
    const float r = e0 * fabsf((1.0f * axis.x + 0.0f * axis.y) + 0.0f * axis.z) + e1 * fabsf((0.0f * axis.x + 1.0f * axis.y) + 0.0f * axis.z) + e2 * fabsf((0.0f * axis.x + 0.0f * axis.y) + 1.0f * axis.z);
    return (rb_fmax2(p2, rb_fmax2(p0, p1)) < -r) || (rb_fmin2(p2, rb_fmin2(p0, p1)) > r);
}
__device__ __forceinline__ bool triangle_cube_detect(f3 ta, f3 tb, f3 tc, f3 ctr, float length, float height, float width) {
    const float e0 = height / 2.0f, e1 = width / 2.0f, e2 = length / 2.0f;
    const f3 v0 = sub3(ta, ctr), v1 = sub3(tb, ctr), v2 = sub3(tc, ctr);
    const f3 f0 = sub3(v1, v0), f1 = sub3(v2, v1), f2 = sub3(v0, v2);
    const f3 u[3] = { mk3(1.f, 0.f, 0.f), mk3(0.f, 1.f, 0.f), mk3(0.f, 0.f, 1.f) };
    const f3 f[3] = { f0, f1, f2 };
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (tri_cube_axis_separates(v0, v1, v2, cross3(u[i], f[j]), e0, e1, e2)) return false;
    if ((rb_fmax2(v0.x, rb_fmax2(v1.x, v2.x)) < -e0) || (rb_fmin2(v0.x, rb_fmin2(v1.x, v2.x)) > e0)) return false;
    if ((rb_fmax2(v0.y, rb_fmax2(v1.y, v2.y)) < -e1) || (rb_fmin2(v0.y, rb_fmin2(v1.y, v2.y)) > e1)) return false;
    if ((rb_fmax2(v0.z, rb_fmax2(v1.z, v2.z)) < -e2) || (rb_fmin2(v0.z, rb_fmin2(v1.z, v2.z)) > e2)) return false;
    const f3 normal = cross3(f0, f1);
    const float d = -dot3(normal, v0);
    f3 vmin, vmax;
    if (normal.x > 0.0f) { vmin.x = -e0; vmax.x = e0; } else { vmin.x = e0; vmax.x = -e0; }
    if (normal.y > 0.0f) { vmin.y = -e1; vmax.y = e1; } else { vmin.y = e1; vmax.y = -e1; }
    if (normal.z > 0.0f) { vmin.z = -e2; vmax.z = e2; } else { vmin.z = e2; vmax.z = -e2; }
    if (dot3(normal, vmin) + d > 0.0f) return false;
    if (dot3(normal, vmax) + d >= 0.0f) return true;
    return false;
}

// GeometryMath::_closestPoint (math/src/GeometryMath.cpp:566-628), Ericson pp.141-142.
__device__ __forceinline__ f3 closest_point(f3 P, f3 A, f3 B, f3 C) {
    f3 ab = sub3(B, A), ac = sub3(C, A), ap = sub3(P, A);
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return A;
    f3 bp = sub3(P, B);
    float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return B;
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        float v = d1 / (d1 - d3);
        return add3(A, mul3(ab, v));
    }
    f3 cp = sub3(P, C);
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return C;
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        float w = d2 / (d2 - d6);
        return add3(A, mul3(ac, w));
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        float w = (d4 - d3) / (d4 - d3 + d5 - d6);
        return add3(B, mul3(sub3(C, B), w));
    }
    float denom = 1.f / (va + vb + vc);
    float v = vb * denom;
    float w = vc * denom;
    return add3(add3(A, mul3(ab, v)), mul3(ac, w));
}

// GeometryMath::sphereCubeDetection (math/src/GeometryMath.cpp:413-487) against the OSP root cube
// (centre 0, edge 2*half): closest point = per-axis clamp, closed test |q-c|^2 <= r^2. With the
// cube at the origin the reference's unit-axis dot products and steps reduce to these exact ops.
__device__ __forceinline__ bool sphere_in_world(f3 c, float r, float half) {
    float qx = fminf(fmaxf(c.x, -half), half);
    float qy = fminf(fmaxf(c.y, -half), half);
    float qz = fminf(fmaxf(c.z, -half), half);
    f3 v = mk3(qx - c.x, qy - c.y, qz - c.z);
    return dot3(v, v) <= r * r;
}

// monotone column index (shared by triangles at build time and spheres every step)
// (x + half) with half = -origin: whole-world grids start at -half, slab grids at s_x0
__host__ __device__ __forceinline__ uint32_t rb_col(float x, float half, float inv_w, uint32_t n) {
    float t = (x + half) * inv_w;
    if (!(t > 0.0f)) return 0;
    if (t >= (float)n) return n - 1;
    return (uint32_t)t;
}

// sphere-grid column of a point
__device__ __forceinline__ uint32_t rb_sphere_col(const RbParams& P, float x, float z) {
    return rb_col(rb_gv(P, x, z), P.g_off_v, P.s_inv_w, P.g_nv) * P.g_nu + rb_col(rb_gu(P, x, z), P.g_off_u, P.s_inv_w, P.g_nu);
}
// ---- hot-body cell table (persistent lists). Entries carry the tick they were written in: an entry of an earlier tick is an
// empty cell, so nothing is cleared between ticks. ----
__device__ __forceinline__ uint32_t hot_cell_head(const RbParams& P, uint32_t cell) {
    const unsigned long long e = P.hot_head[cell];
    return (uint32_t)(e >> 32) == P.tick ? (uint32_t)e : 0u;
}
// link entry h (already holding its row) into the cell of centre (x, z)
__device__ __forceinline__ void hot_link(const RbParams& P, uint32_t h, float x, float z) {
    const uint32_t cell = rb_col(z, P.half, P.hc_inv_w, P.hcn) * P.hcn + rb_col(x, P.half, P.hc_inv_w, P.hcn);
    const unsigned long long old = atomicExch(&P.hot_head[cell], ((unsigned long long)P.tick << 32) | (unsigned long long)(h + 1u));
    P.hot_next[h] = (uint32_t)(old >> 32) == P.tick ? (uint32_t)old : 0u;
}
// a body that steps without its list this tick: its entry in the table, RB_NONE if the table is full (only possible when hot_cap < rows)
__device__ __forceinline__ uint32_t hot_push(const RbParams& P, uint32_t row, float x, float z) {
    const uint32_t h = (*(volatile uint32_t*)P.nhot >= P.hot_cap) ? P.hot_cap : atomicAdd(P.nhot, 1u);
    if (h >= P.hot_cap) return RB_NONE;
    P.hot_row[h] = row;
    hot_link(P, h, x, z);
    return h;
}
