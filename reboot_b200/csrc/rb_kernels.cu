// rb_kernels.cu -- hand-written sm_100a kernels of the ReBoot physics step.
//
//   K1  integrate_bin    StateVector::update + Entity::_updateKinematics + sphere world transform, fused with the
//                        broadphase key build (and, per-tick enumeration, the counting-sort histogram); in slab
//                        worlds the halo pack; on list ticks the per-body validity check of the persistent lists
//   K1b vl_bin           list-rebuild ticks: histogram rank + build poses
//   K2  scan_lookback    single-pass exclusive scan of the column histogram (decoupled look-back)
//   K3  scatter          counting-sort scatter of world spheres into column order (SoA float4)
//   K4  collide_single   canonical-order narrowphase + resolution for single-sphere bodies, four modes:
//                        0 everything per tick, 1 enumerate + prune -> lists, 2 step from lists, 3 hot bodies
//   K4' collide          first-generation per-body kernel (compound bodies, A/B reference)
//   K5  commit           deferred state commit for compound bodies
//   K7  reorder          spatial re-ordering of the body rows
//   gate                 one thread: tail-launches K1b -> K2 -> K3 -> K4<1> when the lists must be rebuilt
//
// All kernels are HBM/L2/latency-bound integer + FP32 work: no tensor cores (nothing here is a contraction).
#include "rb_device.cuh"

// This file is compiled twice. The normal pass holds every kernel and the host launchers. The second pass
// (-rdc=true -DRB_CHAIN_TU, device-linked against cudadevrt) holds only the gate kernel of the persistent-list tick
// and ITS OWN copies of the rebuild-chain kernels it tail-launches (CUDA dynamic parallelism needs relocatable device
// code, and relocatable code made the big per-tick kernels measurably slower: 0.241 -> 0.262 ms per tick), so the
// kernels that run every tick stay in the normal pass.
#ifdef RB_CHAIN_TU
#define RB_KNS rbk_chain
#else
#define RB_KNS rbk
#endif
namespace RB_KNS {

#define RB_BLOCK 256
#ifndef RB_LS_BLOCK
#define RB_LS_BLOCK 256      // rows per triangle tile = threads per block of the list-step kernel (and of the enumeration that fills the tiles)
#endif
#define RB_MAXLOCAL 6   // contacts a thread buffers before the warp-aggregated append

// ------------------------------------------------------------------------------------------
// K1: integrate + world transform + column histogram.  One thread per body.
//   StateVector::update            math/src/StateVector.cpp:11-66
//   Entity::_updateKinematics      model/src/Entity.cpp:131-142
//   GeometryMath::transform(Sphere) math/src/GeometryMath.cpp:81-89
// total = translation(p) * W keeps W's 3x3 and has translation W.t + p (math/src/Matrix.cpp:136-166
// with the translation matrix of :326-338), so a world sphere centre is
// ((W00*ox + W01*oy) + W02*oz) + (W03 + px): the bracket is pre-summed on the host (sph_obj.xyz).
// ------------------------------------------------------------------------------------------
// raise the rebuild flag (many threads may want to in the same tick: only the first few actually store). The plain volatile
// read / write pair races benignly: every writer stores the same value 1, nobody clears the word during K1, and a thread that
// reads a stale 0 merely stores 1 once more; the reason bits go through atomicOr because they accumulate.
__device__ __forceinline__ void vl_request_rebuild(const RbParams& P, uint32_t why) {
    volatile uint32_t* v = P.vl;
    if (!v[RB_VL_REBUILD]) v[RB_VL_REBUILD] = 1u;
    if (!(v[RB_VL_WHY] & why)) atomicOr(&P.vl[RB_VL_WHY], why);
}


// ---- compound bodies in slab (multi-GPU) worlds: body row b keeps its spheres at rows [b * K, b * K + spheres) ----
// halo records of a compound body (byte strings of a per-world size, see RbDistParams)
__device__ __forceinline__ void write_ghost_body(uint8_t* dst, const RbParams& P, uint32_t b, uint32_t s0, uint32_t cnt, f3 m, uint32_t gid) {
    float4* d = reinterpret_cast<float4*>(dst);
    const float4 pt = P.pt[b];
    d[0] = make_float4(m.x, m.y, m.z, __uint_as_float(gid));
    d[1] = make_float4(pt.x, pt.y, pt.z, __uint_as_float(cnt));
    for (uint32_t i = 0; i < cnt; ++i) { d[2 + 2 * i] = P.sph_obj[s0 + i]; d[3 + 2 * i] = P.ws[s0 + i]; }
}
__device__ __forceinline__ void integrate_bin_compound_slab(const RbParams& P, const uint32_t b, const uint32_t f, const float4 p4, const float4 v4,
                                                            const f3 p, const f3 v, const f3 m, const int do_bin) {
    const uint32_t K = P.kstride, s0 = b * K, cnt = P.body_cnt[b];
    const bool active = (f & RB_F_ACTIVE) != 0;
    const float speed = mag3(v), dt = P.dt;
    float ext = 0.f; bool any_in = false;
    for (uint32_t i = 0; i < cnt; ++i) {
        const float4 o = P.sph_obj[s0 + i];
        const f3 c = mk3(o.x + m.x, o.y + m.y, o.z + m.z);
        const float r = o.w;
        const float margin = fminf(P.margin_cap, fmaxf(2.0f * speed * dt + 0.01f * r, 1e-3f));
        const bool in = active && sphere_in_world(c, r, P.half);        // Q3 / Q7, per sphere like OSP::updateOSP
        P.ws[s0 + i] = make_float4(c.x, c.y, c.z, in ? r + margin : -1.0f);
        any_in |= in;
        ext = fmaxf(ext, fabsf(o.x) + r + margin);
    }
    bool binned = true;
    if (P.dp && do_bin && active) {
        const RbDistParams& D = *P.dp;
        const uint32_t gid = D.gid[b];
        // a body belongs to the slab its model translation lies in
        int dir = -1;
        if (m.x < D.slab_lo && D.has_left) dir = 0; else if (m.x >= D.slab_hi && D.has_right) dir = 1;
        if (dir >= 0) {
            const uint32_t kk = atomicAdd(&D.out_hdr[dir]->n_migr, 1u);
            if (kk < D.migr_cap) {
                float4* d = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(D.out_migr[dir]) + (size_t)kk * D.migr_stride);
                d[0] = make_float4(p.x, p.y, p.z, p4.w); d[1] = make_float4(v.x, v.y, v.z, v4.w);
                d[2] = make_float4(m.x, m.y, m.z, 0.f); d[3] = P.pt[b];
                d[4] = make_float4(__uint_as_float(f), __uint_as_float(gid), __uint_as_float(cnt), 0.f);
                d[5] = P.force ? P.force[b] : make_float4(0.f, 0.f, 0.f, 1.0f); d[6] = P.torque ? P.torque[b] : make_float4(0.f, 0.f, 0.f, 0.f);
                d[7] = P.apos ? P.apos[b] : make_float4(0.f, 0.f, 0.f, 0.f); d[8] = P.avel ? P.avel[b] : make_float4(0.f, 0.f, 0.f, 0.f);
                for (uint32_t i = 0; i < cnt; ++i) d[9 + i] = P.sph_obj[s0 + i];
                D.emig[atomicAdd(D.emig_cnt, 1u)] = b;
                if (any_in) {       // stays visible to its old owner for this tick
                    const uint32_t g = atomicAdd(D.self_cnt, 1u);
                    if (g < D.ghost_cap) write_ghost_body(reinterpret_cast<uint8_t*>(D.self_ghost) + (size_t)g * D.ghost_stride, P, b, s0, cnt, m, gid);
                    else atomicAdd(&P.counters[RB_C_HALO_OVF], 1ull);
                }
                binned = false;     // leaves this slab: not binned here
            } else { atomicAdd(&D.out_hdr[dir]->migr_overflow, 1u); atomicAdd(&P.counters[RB_C_HALO_OVF], 1ull); }      // stays owned here one more tick
        } else if (any_in) {
            // a sphere of this body and a sphere of a neighbour's body can only meet if the bodies' reaches overlap across the face
            const float reach = ext + D.ext_max;
            for (int k = 0; k < 2; ++k) {
                if (k == 0 ? !(D.has_left && m.x - reach < D.slab_lo) : !(D.has_right && m.x + reach >= D.slab_hi)) continue;
                const uint32_t kk = atomicAdd(&D.out_hdr[k]->n_ghost, 1u);
                if (kk < D.ghost_cap) write_ghost_body(reinterpret_cast<uint8_t*>(D.out_ghost[k]) + (size_t)kk * D.ghost_stride, P, b, s0, cnt, m, gid);
                else { atomicAdd(&D.out_hdr[k]->ghost_overflow, 1u); atomicAdd(&P.counters[RB_C_HALO_OVF], 1ull); }
            }
        }
    }
    for (uint32_t i = 0; i < K; ++i) {
        uint32_t k = RB_NONE;
        if (i < cnt && binned) {
            const float4 w = P.ws[s0 + i];
            if (w.w >= 0.f) {
                if (do_bin) { k = rb_sphere_col(P, w.x, w.z); P.rank[s0 + i] = atomicAdd(&P.col_count[k], 1u); }
                else k = 0;
            }
        }
        P.key[s0 + i] = k;
    }
}

__device__ __forceinline__ void rb_integrate_bin_body(const RbParams& P, const int do_integrate, const int do_bin) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_owned = rb_n_owned(P);
    if (b == 0 && P.tb_next) {      // the per-tick words of the NEXT tick (nobody touches them during this one)
        P.tb_next[0] = 0u; P.tb_next[1] = 0u; P.tb_next[2] = 0u; P.tb_next[3] = 0u;
        if (do_bin >= 2) P.vl[RB_VL_REBUILD] = do_bin == 3 ? 1u : 0u;      // soft worlds: only the host asks for rebuilds
    }
    if (b >= (P.dn ? n_owned : P.nb)) return;
    float4 p4 = P.pos[b], v4 = P.vel[b];
    uint32_t f = P.flags[b];
    f3 p = xyz(p4), v = xyz(v4);
    const float dt = P.dt;
    // list ticks of worlds whose model translation is always 0 + p (W.t = 0): the mt array is neither read nor written
    const bool lazy = P.lazy_mt && do_bin >= 2 && do_integrate;
    f3 m;
    if (lazy && !P.mt_in_valid) m = mk3(0.f + p.x, 0.f + p.y, 0.f + p.z);      // Entity::_mvp as the previous tick left it
    else m = xyz(P.mt[b]);
    if (do_integrate) {
        float4 F = make_float4(0.f, 0.f, 0.f, 1.0f);
        if (P.force && b < n_owned) {
            F = P.force[b];
            if (P.force_in) {
                // StateVector::setForce for this tick, fused: gated by the contact / gravity flags as they stand before
                // the integration (math/src/StateVector.cpp:147-157); the array is in body-id order
                const float* fi = P.force_in + (size_t)P.force_in_stride * (P.gid ? P.gid[b] : b);
                if ((f & RB_F_CONTACT) || !(f & RB_F_GRAVITY)) F = make_float4(fi[0], fi[1], fi[2], F.w);
                else F = make_float4(0.f, 0.f, 0.f, F.w);
                P.force[b] = F;
            }
        }
        if ((f & RB_F_ACTIVE) && b < n_owned) {
            f3 frc = xyz(F);
            float mass = F.w;
            f3 la = mk3(0.f, 0.f, 0.f);                  // worlds nobody has given a force yet: 0 / 1 = +0, without the three IEEE divisions
            if (P.force) la = mk3(frc.x / mass, frc.y / mass, frc.z / mass);
            float g = (f & RB_F_GRAVITY) ? P.gravity : 0.0f;
            f3 dv = mk3(la.x * dt, (la.y + g) * dt, la.z * dt);
            float fr = ((f & RB_F_CONTACT) || !(f & RB_F_GRAVITY)) ? P.friction : 1.0f;
            v = mk3((v.x + dv.x) * fr, (v.y + dv.y) * fr, (v.z + dv.z) * fr);
            p = mk3(p.x + v.x * dt, p.y + v.y * dt, p.z + v.z * dt);
            P.pos[b] = make_float4(p.x, p.y, p.z, p4.w);
            P.vel[b] = make_float4(v.x, v.y, v.z, v4.w);
            if (P.avel) {
                f3 tq = mk3(0.f, 0.f, 0.f);
                if (P.torque) tq = xyz(P.torque[b]);
                f3 aa = mk3(tq.x / mass, tq.y / mass, tq.z / mass);
                float4 av4 = P.avel[b], ap4 = P.apos[b];
                f3 av = mk3((av4.x + aa.x * dt) * fr, (av4.y + aa.y * dt) * fr, (av4.z + aa.z * dt) * fr);
                P.avel[b] = make_float4(av.x, av.y, av.z, av4.w);
                P.apos[b] = make_float4(ap4.x + av.x * dt, ap4.y + av.y * dt, ap4.z + av.z * dt, ap4.w);
            }
        }
        // Entity::_updateKinematics runs for inactive entities too: prevMVP = mvp; mvp = T(p) * W
        if (b < n_owned) {
            P.pt[b] = make_float4(m.x, m.y, m.z, 0.f);
            f3 w = P.wt ? xyz(P.wt[b]) : mk3(0.f, 0.f, 0.f);
            m = mk3(w.x + p.x, w.y + p.y, w.z + p.z);
            if (!lazy) P.mt[b] = make_float4(m.x, m.y, m.z, 0.f);
        }
    }
    if (P.kstride) { if (b < n_owned) integrate_bin_compound_slab(P, b, f, p4, v4, p, v, m, do_bin); return; }
    if (!P.single) return;      // compound bodies of a single world: their spheres are binned one thread each (rb_bin_spheres_kernel, next launch)
    if (!do_bin && !P.dn) return;
    float speed = mag3(v);
    uint32_t s0 = P.single ? b : P.body_start[b];
    uint32_t s1 = P.single ? b + 1 : P.body_start[b + 1];
    for (uint32_t s = s0; s < s1; ++s) {
        // single-sphere bodies carry their world radius in pos.w; centred spheres (obj_zero) never read sph_obj
        float4 o = (P.single && P.obj_zero) ? make_float4(0.f, 0.f, 0.f, p4.w) : P.sph_obj[s];
        f3 c = mk3(o.x + m.x, o.y + m.y, o.z + m.z);
        float r = o.w;
        // candidate margin: how far this body may be displaced inside one step (pruning only)
        float margin = fminf(P.margin_cap, fmaxf(2.0f * speed * dt + 0.01f * r, 1e-3f));
        uint32_t k = RB_NONE;
        // Q3/Q7: only active bodies are (re)inserted in the broadphase, and only while they touch
        // the OSP root cube (physics/src/OSP.cpp:182-239)
        bool in_grid = (f & RB_F_ACTIVE) && sphere_in_world(c, r, P.half);
        // a sphere outside the broadphase carries w < 0: a list that still names it as a partner skips it
        P.ws[s] = make_float4(c.x, c.y, c.z, in_grid ? r + margin : -1.0f);
        if (P.dp && do_bin && (f & RB_F_ACTIVE)) {
            // K6a fused here: slab worlds hand migrants and ghosts to the halo buffers while the sphere is
            // still in registers (single-sphere bodies: s == b)
            const RbDistParams& D = *P.dp;
            int dir = -1;
            if (c.x < D.slab_lo && D.has_left) dir = 0; else if (c.x >= D.slab_hi && D.has_right) dir = 1;
            // Persistent lists: a migration moves rows, i.e. costs both ranks a list rebuild (and, in lockstep, everybody a
            // wait). A body that has only just crossed the face therefore stays with its owner -- it is the neighbour's
            // ghost either way, and the ghost reach below is wider by the same band -- until the next rebuild the host
            // schedules anyway (do_bin == 3) or until it is further out than the band. Ownership never shows in results.
            if (dir >= 0 && do_bin == 2 && (dir == 0 ? D.slab_lo - c.x : c.x - D.slab_hi) <= D.mig_band) dir = -1;
            if (dir >= 0) {
                if (P.blk_bnd && do_bin >= 2) P.blk_bnd[b / RB_LS_BLOCK] = P.tick;
                uint32_t kk = atomicAdd(&D.out_hdr[dir]->n_migr, 1u);
                if (kk < D.migr_cap) {
                    RbMigrRec mr; mr.pos = make_float4(p.x, p.y, p.z, p4.w); mr.vel = make_float4(v.x, v.y, v.z, v4.w);
                    mr.mt = make_float4(m.x, m.y, m.z, 0.f); mr.pt = P.pt[b]; mr.obj = o; mr.flags = f; mr.gid = D.gid[b]; mr.pad0 = mr.pad1 = 0;
                    mr.force = P.force ? P.force[b] : make_float4(0.f, 0.f, 0.f, 1.0f); mr.torque = P.torque ? P.torque[b] : make_float4(0.f, 0.f, 0.f, 0.f);
                    mr.apos = P.apos ? P.apos[b] : make_float4(0.f, 0.f, 0.f, 0.f); mr.avel = P.avel ? P.avel[b] : make_float4(0.f, 0.f, 0.f, 0.f);
                    D.out_migr[dir][kk] = mr;
                    D.emig[atomicAdd(D.emig_cnt, 1u)] = b;
                    if (do_bin >= 2) vl_request_rebuild(P, 1u);      // rows move (swap-remove): every list is stale, in the grid or not
                    if (in_grid) { uint32_t g = atomicAdd(D.self_cnt, 1u); if (g < D.ghost_cap) { RbGhostRec q; q.ws = make_float4(c.x, c.y, c.z, r + margin); q.r = r; q.gid = D.gid[b]; q.pad0 = q.pad1 = 0; D.self_ghost[g] = q; } else atomicAdd(&P.counters[RB_C_HALO_OVF], 1ull); }
                    in_grid = false;        // leaves this slab: not binned here
                } else { atomicAdd(&D.out_hdr[dir]->migr_overflow, 1u); atomicAdd(&P.counters[RB_C_HALO_OVF], 1ull); }   // stays owned here one more tick
            } else if (in_grid) {
                float reach = r + margin + P.rm_max + D.mig_band;   // own grown radius + the largest grown radius anywhere (+ how far a neighbour's body may sit on our side)
                RbGhostRec q; q.ws = make_float4(c.x, c.y, c.z, r + margin); q.r = r; q.gid = D.gid[b]; q.pad0 = q.pad1 = 0;
                // (a row that is somebody's ghost is exactly a row a ghost can reach: its block waits for the halo)
                if (P.blk_bnd && do_bin >= 2 && ((D.has_left && c.x - reach < D.slab_lo) || (D.has_right && c.x + reach >= D.slab_hi))) P.blk_bnd[b / RB_LS_BLOCK] = P.tick;
                if (D.has_left && c.x - reach < D.slab_lo) {
                    uint32_t kk = atomicAdd(&D.out_hdr[0]->n_ghost, 1u);
                    if (kk < D.ghost_cap) D.out_ghost[0][kk] = q; else { atomicAdd(&D.out_hdr[0]->ghost_overflow, 1u); atomicAdd(&P.counters[RB_C_HALO_OVF], 1ull); }
                }
                if (D.has_right && c.x + reach >= D.slab_hi) {
                    uint32_t kk = atomicAdd(&D.out_hdr[1]->n_ghost, 1u);
                    if (kk < D.ghost_cap) D.out_ghost[1][kk] = q; else { atomicAdd(&D.out_hdr[1]->ghost_overflow, 1u); atomicAdd(&P.counters[RB_C_HALO_OVF], 1ull); }
                }
            }
        }
        if (in_grid) {
            if (do_bin) {
                k = rb_sphere_col(P, c.x, c.z);
                if (do_bin == 1) P.rank[s] = atomicAdd(&P.col_count[k], 1u);
            } else k = 0;
            // lists switched off for now (the host sits out a phase of constant rebuilding on the per-tick kernel):
            // keep reporting how many bodies outrun the skin within one tick, so that it knows when to come back
            // (would a list last this body 8 ticks? displacement + margin after 8 ticks, a body in free fall accelerating on the way)
            // (only a body that IS falling accelerates: one that rests on other spheres carries no contact flag either -- the flag is
            // the triangles' -- but its speed never gets past the few g*dt it gains between two snap-backs. Without the speed test
            // such bodies count as fast movers for good once the skin is small: r = 0.18 in the 8-GPU weak scene, lists never resumed)
            const float fall = ((f & RB_F_GRAVITY) && !(f & RB_F_CONTACT) && speed > 4.0f * fabsf(P.gravity) * dt) ? fabsf(P.gravity) * dt : 0.0f;
            if (do_bin == 1 && P.vl && !(8.0f * (speed + 4.0f * fall) * dt + 2.0f * (speed + 8.0f * fall) * dt + 0.01f * r <= P.vl_skin)) {
                volatile uint32_t* nf = P.vl + RB_VL_NFAST;
                if (*nf < P.hot_flood) atomicAdd(&P.vl[RB_VL_NFAST], 1u);
            }
        }
        if (do_bin >= 2) {
            // persistent lists stay valid while every sphere stays inside the skin it was enumerated with:
            // displacement since the rebuild + this tick's margin <= skin, and nobody entered or left the grid
            const uint32_t k_prev = P.key[s];
            const float4 wb = P.ws_build[s];
            const float ddx = c.x - wb.x, ddy = c.y - wb.y, ddz = c.z - wb.z;
            const float moved = sqrtf((ddx * ddx + ddy * ddy) + ddz * ddz);
            bool hot = false;
            if (do_bin == 3) { /* the host asked for a rebuild: nothing to classify */ }
            else if ((k_prev == RB_NONE) != (k == RB_NONE)) {
                if (!P.soft) vl_request_rebuild(P, 1u);
                else if (k != RB_NONE) { hot = true; P.ws_build[s].w = -1.0f; }     // entered the grid: no list until the next rebuild
                // (left the grid: its world sphere says so, w < 0)
            }
            else if (k != RB_NONE && (!(moved + margin <= P.vl_skin) || wb.w < 0.f)) hot = true;      // w < 0: this row has no usable list
            if (hot) {
                // hot: steps without its list this tick (its own small launch, rb_collide_single_kernel<3>)
                // (slab worlds: once more bodies are hot than a rebuild costs, the device asks for one; single worlds leave that to the host)
                const uint32_t h = hot_push(P, s, c.x, c.z);
                if (h == RB_NONE || (!P.soft && h >= P.hot_soft)) vl_request_rebuild(P, 2u);
                if (P.dn && h != RB_NONE) P.rank[s] = h;      // slab worlds: a swap-remove that moves this row later in the tick repoints its entry (rb_halo_remove_kernel)
                k |= RB_HOT_BIT;
            }
            if (s == 0 && do_bin == 3 && !P.tb_next) { P.vl[RB_VL_REBUILD] = 1u; P.vl[RB_VL_WHY] |= 4u; }
            if (s == 0 && do_bin == 3 && P.tb_next) P.vl[RB_VL_WHY] |= 4u;
            if (k != k_prev) P.key[s] = k;
        } else P.key[s] = k;
    }
}


// ------------------------------------------------------------------------------------------
// K2: exclusive scan of col_count[0..n) into col_start[0..n], three launches.
// ------------------------------------------------------------------------------------------
#define RB_SCAN_ITEMS 8
#define RB_SCAN_TILE (RB_BLOCK * RB_SCAN_ITEMS)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += t;
    }
    return v;
}
// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[RB_BLOCK / 32];
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t nwarps = blockDim.x >> 5;          // (blocks of up to RB_BLOCK threads)
    uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nwarps ? wsum[lane] : 0;
        uint32_t wi = warp_incl_scan(w, lane);
        if (lane < nwarps) wsum[lane] = wi - w;
        if (lane == nwarps - 1) *total = wi;
    }
    __syncthreads();
    return inc - v + wsum[wid];
}
__global__ void __launch_bounds__(RB_BLOCK) rb_scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t total;
    uint32_t base = blockIdx.x * RB_SCAN_TILE + threadIdx.x * RB_SCAN_ITEMS;
    uint32_t s = 0;
    if (base + RB_SCAN_ITEMS <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(in + base);
        uint4 a = p[0], b = p[1];
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (uint32_t i = 0; i < RB_SCAN_ITEMS; ++i) if (base + i < n) s += in[base + i];
    }
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(RB_BLOCK) rb_scan_top_kernel(uint32_t* tile_sums, uint32_t ntiles, uint32_t* grand_total) {
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ntiles; base += RB_BLOCK) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0;
        uint32_t ex = block_excl_scan(v, &total);
        if (i < ntiles) tile_sums[i] = ex + carry;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void __launch_bounds__(RB_BLOCK) rb_scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ tile_offs, uint32_t* __restrict__ out) {
    __shared__ uint32_t total;
    uint32_t base = blockIdx.x * RB_SCAN_TILE + threadIdx.x * RB_SCAN_ITEMS;
    uint32_t v[RB_SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (uint32_t i = 0; i < RB_SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    uint32_t ex = block_excl_scan(s, &total) + tile_offs[blockIdx.x];
#pragma unroll
    for (uint32_t i = 0; i < RB_SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    // out[n] (the grand total) is written by rb_scan_top_kernel
}

// K2 (default): single-pass exclusive scan with decoupled look-back (Merrill & Garland): one launch, each
// tile of 2048 counts is read once and written once. `status[t]` packs (flag << 32 | value): flag 1 =
// tile aggregate available, 2 = inclusive prefix available. Tile ids are handed out by an atomic ticket
// so that a tile only ever waits on tiles that are already running. ticket + status are zeroed by the
// same memset that clears the histogram (they sit right behind it).
// `cond` (optional): the kernel returns at once unless *cond is set; with `clear_in` the histogram is zeroed as
// it is read (persistent-list ticks: no per-tick memset, the next rebuild finds a clean histogram).
__global__ void __launch_bounds__(RB_BLOCK) rb_scan_lookback_kernel(uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out,
                                                                    unsigned long long* status, uint32_t* ticket, const uint32_t* cond, int clear_in) {
    __shared__ uint32_t s_tile, s_total, s_prefix;
    if (cond && !*cond) return;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * RB_SCAN_TILE + threadIdx.x * RB_SCAN_ITEMS;
    uint32_t v[RB_SCAN_ITEMS];
    uint32_t sum = 0;
    if (base + RB_SCAN_ITEMS <= n) {
        uint4* p = reinterpret_cast<uint4*>(in + base);
        uint4 a = p[0], b = p[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        if (clear_in) { p[0] = make_uint4(0u, 0u, 0u, 0u); p[1] = make_uint4(0u, 0u, 0u, 0u); }
    } else {
#pragma unroll
        for (uint32_t i = 0; i < RB_SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; if (clear_in && base + i < n) in[base + i] = 0; }
    }
#pragma unroll
    for (uint32_t i = 0; i < RB_SCAN_ITEMS; ++i) sum += v[i];
    uint32_t ex = block_excl_scan(sum, &s_total);
    if (threadIdx.x < 32) {          // warp 0 publishes the tile aggregate and looks back 32 tiles at a time
        const uint32_t lane = threadIdx.x;
        const uint32_t total = s_total;
        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&status[0], (2ull << 32) | total);
        } else {
            if (lane == 0) atomicExch(&status[tile], (1ull << 32) | total);
            int idx = (int)tile - 1 - (int)lane;          // lane 0 looks at the nearest predecessor
            while (true) {
                unsigned long long st = (2ull << 32);     // virtual inclusive 0 in front of tile 0
                if (idx >= 0) { do { st = atomicAdd(&status[idx], 0ull); } while ((st >> 32) == 0); }
                const uint32_t incl = __ballot_sync(0xffffffffu, (st >> 32) == 2);
                if (incl) {
                    const uint32_t first = __ffs(incl) - 1;       // nearest tile that already holds its inclusive prefix
                    prefix += __reduce_add_sync(0xffffffffu, lane <= first ? (uint32_t)st : 0u);
                    break;
                }
                prefix += __reduce_add_sync(0xffffffffu, (uint32_t)st);
                idx -= 32;
            }
            if (lane == 0) atomicExch(&status[tile], (2ull << 32) | (unsigned long long)(prefix + total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if ((tile + 1) * RB_SCAN_TILE >= n) out[n] = prefix + total;      // grand total behind the last element
        }
    }
    __syncthreads();
    ex += s_prefix;
    if (base + RB_SCAN_ITEMS <= n) {
        uint4 a, b;
        a.x = ex; ex += v[0]; a.y = ex; ex += v[1]; a.z = ex; ex += v[2]; a.w = ex; ex += v[3];
        b.x = ex; ex += v[4]; b.y = ex; ex += v[5]; b.z = ex; ex += v[6]; b.w = ex;
        uint4* q = reinterpret_cast<uint4*>(out + base);
        q[0] = a; q[1] = b;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < RB_SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    }
}

// ------------------------------------------------------------------------------------------
// K3: counting-sort scatter into column order.
// ------------------------------------------------------------------------------------------
// vl != 0: persistent-list rebuild tick -- conditional on the rebuild flag, spheres are stored grown by the
// skin, and the scan's ticket / tile status words are reset for the next rebuild.
__global__ void __launch_bounds__(RB_BLOCK) rb_scatter_kernel(RbParams P, int vl, unsigned long long* scan_status, uint32_t* scan_ticket, uint32_t scan_tiles) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (vl) {
        if (!P.vl[RB_VL_REBUILD]) return;
        for (uint32_t t = s; t < scan_tiles; t += gridDim.x * blockDim.x) scan_status[t] = 0ull;
        if (s == 0) *scan_ticket = 0u;
    }
    if (P.dn) {     // slab worlds: owned rows in front, ghosts at the end (not sorted on list ticks)
        const uint32_t br = P.kstride ? s / P.kstride : s;
        if (br >= P.nb || (br >= P.dn[0] && (vl || br < P.nb - P.dn[1]))) return;
    } else if (s >= P.ns) return;
    uint32_t k = P.key[s];
    if (k == RB_NONE) return;
    k &= ~RB_HOT_BIT;
    uint32_t slot = P.col_start[k] + P.rank[s];
    float4 wsv = P.ws[s];
    if (vl) wsv.w += P.vl_skin;
    P.sorted[slot] = wsv;
    // bit 31: the body's contact flag, so that the collide kernel needs no per-body gather for bodies
    // that touch nothing this tick (Physics.cpp:201-207 only needs "was in contact")
    uint32_t body = P.single ? s : rb_sph_owner(P, s);
    P.sorted_id[slot] = s | ((P.flags[body] & RB_F_CONTACT) ? RB_TRI_BIT : 0u);
}

// Per-tick clears run as a kernel, not as cudaMemsetAsync: a memset may be executed by a copy engine, and then it
// queues behind whatever transfer that engine is busy with -- with the pipelined host round trip that stalled the
// tick for a good part of every 12-MB copy (measured: +60 us per tick).
__global__ void __launch_bounds__(RB_BLOCK) rb_clear_kernel(uint4* __restrict__ a, uint32_t n16a, uint32_t* __restrict__ a_tail, uint32_t ntail,
                                                           uint32_t* __restrict__ b, uint32_t nb4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16a) a[i] = make_uint4(0u, 0u, 0u, 0u);
    if (i < ntail) a_tail[i] = 0u;         // the (< 4) words behind the last full 16 bytes: never a byte more than asked
    if (i < nb4) b[i] = 0u;
}

// K1b (persistent lists, rebuild ticks only): histogram rank of every binned row + the poses the new lists
// are built for. `cond` kernels return at once on the ticks whose lists are still valid.
__global__ void __launch_bounds__(RB_BLOCK) rb_vl_bin_kernel(RbParams P) {
    if (!P.vl[RB_VL_REBUILD]) return;
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0 && P.tile_cursor) { *P.tile_cursor = 0u; P.vl[RB_VL_NOLIST] = P.vl[RB_VL_NOLIST + 1] = P.vl[RB_VL_NOLIST + 2] = 0u; }      // the enumeration at the end of this chain refills the triangle tiles
    if (s >= (P.dn ? P.dn[0] : P.ns)) return;       // slab worlds: owned rows only (ghosts never enter the sort of a list tick)
    P.ws_build[s] = P.ws[s];
    uint32_t k = P.key[s];
    if (k != RB_NONE) P.rank[s] = atomicAdd(&P.col_count[k & ~RB_HOT_BIT], 1u);
}

// ------------------------------------------------------------------------------------------
// K4: collide.  One thread per body, canonical order:
//   1. sphere-sphere against partner bodies in ascending body id, sphere pairs in ascending order
//      with the lower body id outermost; the partner is simulated from its frozen pose.
//        detection  GeometryMath::sphereSphereDetection   math/src/GeometryMath.cpp:211-230
//        resolution GeometryMath::sphereSphereResolution  math/src/GeometryMath.cpp:523-554
//   2. sphere-triangle for own spheres ascending, triangles in ascending global id, each tested at
//      the body's CURRENT pose and resolved immediately (Gauss-Seidel like Physics.cpp:165-178).
//        detection  GeometryMath::sphereTriangleDetection  math/src/GeometryMath.cpp:101-209
//        resolution GeometryMath::sphereTriangleResolution math/src/GeometryMath.cpp:489-522
//   3. contact-flag bookkeeping of Physics.cpp:180,201-207.
// Entity::setPosition (model/src/Entity.cpp:373-391) is the pose update used by both resolutions.
// ------------------------------------------------------------------------------------------
struct BodyState {
    f3 p, v, mt, pt, wt;
    uint32_t flags;
    bool changed;
};

__device__ __forceinline__ void body_set_position(BodyState& S, f3 np) {
    f3 t = mk3(S.wt.x + np.x, S.wt.y + np.y, S.wt.z + np.z);   // translation of T(np) * W
    S.p = t;                                                    // _state.setLinearPosition(total.t)
    S.pt = S.mt;                                                // _prevMVP.setModel(_mvp.getModelMatrix())
    S.mt = t;                                                   // _mvp.setModel(total)
    S.changed = true;
}

struct HitBuf {
    uint32_t n;
    uint32_t a[RB_MAXLOCAL], b[RB_MAXLOCAL];
    float depth[RB_MAXLOCAL], nx[RB_MAXLOCAL], ny[RB_MAXLOCAL], nz[RB_MAXLOCAL], r[RB_MAXLOCAL];
};

__device__ __forceinline__ void store_contact(const RbParams& P, unsigned long long slot, uint32_t a, uint32_t b,
                                              float depth, float nx, float ny, float nz, float r) {
    if (slot < P.contact_cap) {
        float4* dst = reinterpret_cast<float4*>(P.contacts + slot);
        dst[0] = make_float4(__uint_as_float(a), __uint_as_float(b), depth, nx);
        dst[1] = make_float4(ny, nz, r, 0.f);
    }
}
__device__ __forceinline__ void record_hit(const RbParams& P, HitBuf& H, uint32_t a, uint32_t b, float depth, f3 n, float r) {
    if (H.n == RB_MAXLOCAL) {
        // more contacts than the local buffer holds (two overlapping compound bodies hit dozens of times in a tick): the full
        // buffer is appended with ONE atomic -- a dependent global atomic per hit was the longest stall of those bodies' chains
        const unsigned long long base = atomicAdd(P.ncontact, (unsigned long long)RB_MAXLOCAL);
        uint32_t dropped = 0;
#pragma unroll
        for (uint32_t i = 0; i < RB_MAXLOCAL; ++i) {
            store_contact(P, base + i, H.a[i], H.b[i], H.depth[i], H.nx[i], H.ny[i], H.nz[i], H.r[i]);
            if (base + i >= P.contact_cap) dropped++;
        }
        if (dropped) atomicAdd(&P.counters[RB_C_DROPPED], (unsigned long long)dropped);
        H.n = 0;
    }
    const uint32_t i = H.n;
    H.a[i] = a; H.b[i] = b; H.depth[i] = depth; H.nx[i] = n.x; H.ny[i] = n.y; H.nz[i] = n.z; H.r[i] = r;
    H.n++;
}

struct Tally { uint32_t tests_ss, tests_st, cand_ss, cand_st, hits_ss, hits_st, overflow; };

// one triangle against sphere `sa` of the body, at the body's current pose
template <bool RESOLVE>
__device__ __forceinline__ void process_triangle(const RbParams& P, BodyState& S, HitBuf& H, Tally& T,
                                                 uint32_t sa_global, f3 so, float r, f3 fc, float frm,
                                                 uint32_t tid_global, bool& hit_any) {
    const float4* tp = P.tri + 3ull * tid_global;
    float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
    T.cand_st++;
    // prune against the frozen sphere's box grown by the margin (never changes results)
    if (fminf(a4.x, fminf(b4.x, c4.x)) > fc.x + frm || fmaxf(a4.x, fmaxf(b4.x, c4.x)) < fc.x - frm) return;
    if (fminf(a4.y, fminf(b4.y, c4.y)) > fc.y + frm || fmaxf(a4.y, fmaxf(b4.y, c4.y)) < fc.y - frm) return;
    if (fminf(a4.z, fminf(b4.z, c4.z)) > fc.z + frm || fmaxf(a4.z, fmaxf(b4.z, c4.z)) < fc.z - frm) return;
    f3 A = xyz(a4), B = xyz(b4), C = xyz(c4);
    f3 c = mk3(so.x + S.mt.x, so.y + S.mt.y, so.z + S.mt.z);   // Sphere::getPosition at the current pose
    T.tests_st++;
    if (!sphere_triangle_detect(c, r, A, B, C)) return;
    hit_any = true;
    T.hits_st++;
    f3 q = closest_point(c, A, B, C);
    f3 ov = sub3(c, q);
    float dist = mag3(ov);
    f3 o = mk3(ov.x / dist, ov.y / dist, ov.z / dist);          // Vector4::normalize
    record_hit(P, H, sa_global, RB_TRI_BIT | tid_global, r - dist, o, r);
    if (RESOLVE) {
        f3 normal = mk3(a4.w, b4.w, c4.w);          // normalize((C-A)x(B-A)), pre-computed at build
        f3 delta = sub3(add3(mul3(o, r * P.pushout), q), c);
        float nc = dot3(S.v, normal);
        f3 rv = sub3(S.v, mul3(normal, nc));
        body_set_position(S, add3(S.p, delta));
        S.v = rv;
        S.flags |= RB_F_CONTACT;
    }
}

// sphere-triangle pass of one sphere: ascending triangle id over the columns its grown box touches
template <bool RESOLVE>
__device__ __forceinline__ void sphere_vs_triangles(const RbParams& P, BodyState& S, HitBuf& H, Tally& T,
                                                    uint32_t sa_global, f3 so, float r, f3 fc, float frm, bool& hit_any) {
    if (P.nt == 0) return;
    uint32_t cx0 = rb_col(fc.x - frm, P.half, P.t_inv_w, P.tcx), cx1 = rb_col(fc.x + frm, P.half, P.t_inv_w, P.tcx);
    uint32_t cz0 = rb_col(fc.z - frm, P.half, P.t_inv_w, P.tcz), cz1 = rb_col(fc.z + frm, P.half, P.t_inv_w, P.tcz);
    uint32_t nx = cx1 - cx0 + 1, nz = cz1 - cz0 + 1, ncol = nx * nz;
    if (ncol <= 4) {
        uint32_t cur[4], end[4], head[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            cur[k] = end[k] = 0; head[k] = RB_NONE;
            if (k < ncol) {
                uint32_t col = (cz0 + k / nx) * P.tcx + cx0 + k % nx;
                cur[k] = __ldg(P.tcol_start + col); end[k] = __ldg(P.tcol_start + col + 1);
                if (cur[k] < end[k]) head[k] = __ldg(P.tcol_tri + cur[k]);
            }
        }
        while (true) {
            uint32_t m = min(min(head[0], head[1]), min(head[2], head[3]));
            if (m == RB_NONE) break;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                if (head[k] == m) { cur[k]++; head[k] = cur[k] < end[k] ? __ldg(P.tcol_tri + cur[k]) : RB_NONE; }
            process_triangle<RESOLVE>(P, S, H, T, sa_global, so, r, fc, frm, m, hit_any);
        }
    } else {
        // big sphere / fine columns: repeated selection of the next id above the last one processed
        uint32_t lo = 0;   // ids below lo are done
        while (true) {
            uint32_t best = RB_NONE;
            for (uint32_t cz = cz0; cz <= cz1; ++cz)
                for (uint32_t cx = cx0; cx <= cx1; ++cx) {
                    uint32_t col = cz * P.tcx + cx;
                    uint32_t i = __ldg(P.tcol_start + col), e = __ldg(P.tcol_start + col + 1);
                    for (; i < e; ++i) { uint32_t t = __ldg(P.tcol_tri + i); if (t >= lo) { best = min(best, t); break; } }
                }
            if (best == RB_NONE) break;
            process_triangle<RESOLVE>(P, S, H, T, sa_global, so, r, fc, frm, best, hit_any);
            lo = best + 1;
        }
    }
}

// contact append (one atomic per warp) + counters (one atomic per block and counter); every thread of
// the block must call this
__device__ __forceinline__ void collide_epilogue(const RbParams& P, HitBuf& H, Tally& T, const uint32_t lane) {
    // ---- warp-aggregated contact append: one atomic per warp ----
    uint32_t nloc = min(H.n, (uint32_t)RB_MAXLOCAL);
    uint32_t incl = warp_incl_scan(nloc, lane);
    uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
    if (wtotal) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.ncontact, (unsigned long long)wtotal);
        base = __shfl_sync(0xffffffffu, base, 0);
        unsigned long long slot = base + (incl - nloc);
        uint32_t dropped = 0;
#pragma unroll
        for (uint32_t i = 0; i < RB_MAXLOCAL; ++i)
            if (i < nloc) {
                store_contact(P, slot + i, H.a[i], H.b[i], H.depth[i], H.nx[i], H.ny[i], H.nz[i], H.r[i]);
                if (slot + i >= P.contact_cap) dropped++;
            }
        dropped = __reduce_add_sync(0xffffffffu, dropped);
        if (lane == 0 && dropped) atomicAdd(&P.counters[RB_C_DROPPED], (unsigned long long)dropped);
    }
    // ---- counters: warp reduce, then one atomic per warp and counter on a striped copy (no block barrier,
    // little contention); the host sums the RB_STRIPES stripes when it reads the statistics ----
    uint32_t vals[7] = { T.tests_ss, T.tests_st, T.cand_ss, T.cand_st, T.hits_ss, T.hits_st, T.overflow };
    const int map[7] = { RB_C_TESTS_SS, RB_C_TESTS_ST, RB_C_CAND_SS, RB_C_CAND_ST, RB_C_HITS_SS, RB_C_HITS_ST, RB_C_OVERFLOW };
    const uint32_t stripe = ((blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5)) & (RB_STRIPES - 1);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        uint32_t sum = __reduce_add_sync(0xffffffffu, vals[k]);
        if (lane == 0 && sum) atomicAdd(&P.stripes[stripe * RB_C_COUNT + map[k]], (unsigned long long)sum);
    }
}

// scan the sphere grid around frozen sphere (fc, frm) for the smallest partner body id >= lo
// (ids are global body ids; best_local is the partner's local body index)
template <bool SINGLE>
__device__ __forceinline__ void scan_partners(const RbParams& P, uint32_t A, f3 fc, float frm, uint32_t lo,
                                              uint32_t& best, uint32_t& best_local, uint32_t& eligible, Tally& T) {
    float reach = frm + P.rm_max;
    const float fu = rb_gu(P, fc.x, fc.z), fv = rb_gv(P, fc.x, fc.z);
    uint32_t cx0 = rb_col(fu - reach, P.g_off_u, P.s_inv_w, P.g_nu), cx1 = rb_col(fu + reach, P.g_off_u, P.s_inv_w, P.g_nu);
    uint32_t cz0 = rb_col(fv - reach, P.g_off_v, P.s_inv_w, P.g_nv), cz1 = rb_col(fv + reach, P.g_off_v, P.s_inv_w, P.g_nv);
    for (uint32_t cz = cz0; cz <= cz1; ++cz) {
        uint32_t row = cz * P.g_nu;
        uint32_t i = P.col_start[row + cx0], e = P.col_start[row + cx1 + 1];   // u-adjacent columns are contiguous
        for (; i < e; ++i) {
            float4 o = P.sorted[i];
            float dx = o.x - fc.x, dy = o.y - fc.y, dz = o.z - fc.z, rr = frm + o.w;
            if ((dx * dx + dy * dy) + dz * dz > rr * rr) continue;
            uint32_t sb = P.sorted_id[i] & ~RB_TRI_BIT;
            uint32_t Bl = SINGLE ? sb : rb_sph_owner(P, sb);
            if (Bl == A) continue;
            T.cand_ss++;
            uint32_t B = rb_gid(P, Bl);
            if (B < lo) continue;
            eligible++;
            if (B < best) { best = B; best_local = Bl; }
        }
    }
}

// the same over the coarse table of hot bodies and ghosts (persistent lists)
__device__ __forceinline__ void scan_hot_partners(const RbParams& P, uint32_t A, f3 fc, float frm, uint32_t lo,
                                                  uint32_t& best, uint32_t& best_local, uint32_t& eligible) {
    const float reach = frm + P.rm_max;
    const uint32_t hx0 = rb_col(fc.x - reach, P.half, P.hc_inv_w, P.hcn), hx1 = rb_col(fc.x + reach, P.half, P.hc_inv_w, P.hcn);
    const uint32_t hz0 = rb_col(fc.z - reach, P.half, P.hc_inv_w, P.hcn), hz1 = rb_col(fc.z + reach, P.half, P.hc_inv_w, P.hcn);
    for (uint32_t hz = hz0; hz <= hz1; ++hz)
        for (uint32_t hx = hx0; hx <= hx1; ++hx)
            for (uint32_t h = hot_cell_head(P, hz * P.hcn + hx); h; h = P.hot_next[h - 1u]) {
                const uint32_t Bl = P.hot_row[h - 1u];
                if (Bl == A || Bl == RB_NONE) continue;
                const float4 o = P.ws[Bl];
                float dx = o.x - fc.x, dy = o.y - fc.y, dz = o.z - fc.z, rr = frm + o.w;
                if (o.w < 0.f || (dx * dx + dy * dy) + dz * dz > rr * rr) continue;
                const uint32_t B = rb_gid(P, Bl);
                if (B < lo || B == best) continue;
                eligible++;
                if (B < best) { best = B; best_local = Bl; }
            }
}

template <bool SINGLE, bool RESOLVE>
__global__ void __launch_bounds__(RB_BLOCK) rb_collide_kernel(RbParams P) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    HitBuf H; H.n = 0;
    bool valid;
    uint32_t A = 0, s0 = 0, nsph = 0;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (SINGLE) {
        uint32_t n_sorted = P.col_start[P.scx * P.scz];
        valid = tid < n_sorted;
        if (valid) { f0 = P.sorted[tid]; A = s0 = P.sorted_id[tid] & ~RB_TRI_BIT; nsph = 1; valid = A < rb_n_owned(P); }
    } else {
        valid = tid < rb_n_owned(P);
        if (valid) { A = tid; s0 = rb_body_s0(P, A); nsph = rb_body_ns(P, A); valid = (P.flags[A] & RB_F_ACTIVE) != 0; }
    }
    if (valid) {
        BodyState S;
        S.p = xyz(P.pos[A]); S.v = xyz(P.vel[A]); S.mt = xyz(P.mt[A]); S.pt = xyz(P.pt[A]);
        S.wt = P.wt ? xyz(P.wt[A]) : mk3(0.f, 0.f, 0.f);
        S.flags = P.flags[A];
        S.changed = false;
        const uint32_t flags_in = S.flags;
        const f3 mt_frozen = S.mt;

        // ---------------- sphere-sphere ----------------
        const uint32_t gA = rb_gid(P, A);
        uint32_t lo = 0;
        while (true) {
            uint32_t best = RB_NONE, best_local = RB_NONE, eligible = 0;
            for (uint32_t i = 0; i < nsph; ++i) {
                uint32_t sa = s0 + i;
                if (!SINGLE && P.key[sa] == RB_NONE) continue;
                float4 fa = SINGLE ? f0 : P.ws[sa];
                scan_partners<SINGLE>(P, A, xyz(fa), fa.w, lo, best, best_local, eligible, T);
            }
            if (best == RB_NONE) break;
            const uint32_t B = best_local;
            const bool lo_is_a = gA < best;
            if (RESOLVE || lo_is_a) {
                // partner simulated from its frozen pose (it toggles exactly like us on every hit)
                uint32_t t0 = SINGLE ? B : rb_body_s0(P, B);
                uint32_t nb_sph = SINGLE ? 1 : rb_body_ns(P, B);
                f3 bmt = mk3(0.f, 0.f, 0.f), bpt = bmt, bwt = bmt;
                bool b_toggled = false;
                if (!SINGLE) { bmt = xyz(P.mt[B]); bpt = xyz(P.pt[B]); if (P.wt) bwt = xyz(P.wt[B]); }
                uint32_t n_out = lo_is_a ? nsph : nb_sph, n_in = lo_is_a ? nb_sph : nsph;
                for (uint32_t io = 0; io < n_out; ++io)
                    for (uint32_t ii = 0; ii < n_in; ++ii) {
                        uint32_t ia = lo_is_a ? io : ii, ib = lo_is_a ? ii : io;
                        uint32_t sa = s0 + ia, sb = t0 + ib;
                        if (!SINGLE && (P.key[sa] == RB_NONE || P.key[sb] == RB_NONE)) continue;
                        float4 oa = P.sph_obj[sa], ob = P.sph_obj[sb];
                        f3 ca = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
                        f3 cb;
                        if (SINGLE || !b_toggled) { float4 wb = P.ws[sb]; cb = xyz(wb); }   // frozen partner pose
                        else cb = mk3(ob.x + bmt.x, ob.y + bmt.y, ob.z + bmt.z);
                        // the lower body id plays "sphereA" of GeometryMath.cpp:211 / "modelA" of :523
                        f3 cn = lo_is_a ? sub3(ca, cb) : sub3(cb, ca);
                        float dist = mag3(cn);
                        float radii = lo_is_a ? (oa.w + ob.w) : (ob.w + oa.w);
                        if (lo_is_a) T.tests_ss++;
                        if (!(radii - dist >= 0)) continue;
                        f3 n1 = mk3(cn.x / dist, cn.y / dist, cn.z / dist);
                        if (lo_is_a) { T.hits_ss++; record_hit(P, H, SINGLE ? gA : rb_sph_gid(P, A, s0, ia), SINGLE ? best : rb_sph_gid(P, B, t0, ib), radii - dist, n1, oa.w); }
                        if (RESOLVE) {
                            f3 n = n1;
                            if (!lo_is_a) { n = neg3(n); n = normalize3(n); }      // "Opposite reaction", normalised again
                            float sp = mag3(S.v);
                            S.v = mul3(mul3(n, sp), P.ss_damp);
                            body_set_position(S, S.pt);                             // snap back to the previous pose
                            if (!SINGLE) {                                          // partner does the same on its side
                                f3 t = mk3(bwt.x + bpt.x, bwt.y + bpt.y, bwt.z + bpt.z);
                                bpt = bmt; bmt = t; b_toggled = true;
                            }
                        }
                    }
            }
            lo = best + 1;
            if (eligible <= 1 && SINGLE) break;   // nothing else above B in range
        }

        // ---------------- sphere-triangle ----------------
        bool last_hit = false;
        for (uint32_t i = 0; i < nsph; ++i) {
            uint32_t sa = s0 + i;
            bool hit_any = false;
            if (SINGLE || P.key[sa] != RB_NONE) {
                float4 fa = SINGLE ? f0 : P.ws[sa];
                float4 oa = P.sph_obj[sa];
                sphere_vs_triangles<RESOLVE>(P, S, H, T, SINGLE ? gA : rb_sph_gid(P, A, s0, i), xyz(oa), oa.w, xyz(fa), fa.w, hit_any);
            }
            if (i == nsph - 1) last_hit = hit_any;
        }

        if (RESOLVE) {
            // Physics.cpp:201-207: a body that was in contact and whose last sphere found no triangle loses the flag
            if ((flags_in & RB_F_CONTACT) && !last_hit) S.flags &= ~RB_F_CONTACT;
            // margin audit: the candidate sets were built for displacements up to the smallest margin
            f3 d = sub3(S.mt, mt_frozen);
            float disp = mag3(d);
            float4 fa0 = SINGLE ? f0 : P.ws[s0];
            if (disp > fa0.w - P.sph_obj[s0].w) T.overflow++;
            if (SINGLE) {
                if (S.changed) {
                    P.pos[A] = make_float4(S.p.x, S.p.y, S.p.z, P.sph_obj[s0].w);      // single-sphere bodies keep their world radius in pos.w
                    P.vel[A] = make_float4(S.v.x, S.v.y, S.v.z, 0.f);
                    P.mt[A] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f);
                    P.pt[A] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
                }
                if (S.flags != flags_in) P.flags[A] = S.flags;
            } else if (S.changed || S.flags != flags_in) {
                // partners read our frozen pose concurrently: commit after the kernel (K5)
                uint32_t k = (uint32_t)atomicAdd(&P.counters[RB_C_NCHANGED], 1ull);
                P.ch_body[k] = A;
                P.ch_pos[k] = make_float4(S.p.x, S.p.y, S.p.z, 1.f);
                P.ch_vel[k] = make_float4(S.v.x, S.v.y, S.v.z, 0.f);
                P.ch_mt[k] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f);
                P.ch_pt[k] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
                P.ch_flags[k] = S.flags;
            }
        }
    }

    collide_epilogue(P, H, T, lane);
}

// ------------------------------------------------------------------------------------------
// K4 (single-sphere bodies), phased so that the lanes of a warp do the same kind of work at the same
// time (the v1 kernel above ran ~9 of 32 lanes on average):
//   P1  one scan of the sphere grid: partners passing the cheap reject, kept sorted by global id
//   P2  sphere-sphere tests / resolutions in ascending partner id
//   P3  merge walk over the triangle columns (skipping columns whose y range the grown sphere misses),
//       AABB prune against the frozen box, surviving triangle ids staged in shared memory
//   P4  lock-step exact tests of the staged triangles at the body's current pose, resolving inline
//   P5  flags, state write-back, contact append
// Same canonical order and arithmetic as the v1 kernel (bit-identical results).
// ------------------------------------------------------------------------------------------
#define RB_CAND_MAX 12
#define RB_SSCAND_MAX 4       // partners a thread keeps in registers (per-tick enumeration, hot bodies, hot partners of a listed row)
#ifndef RB_SSLIST_MAX
#define RB_SSLIST_MAX 8       // partners a persistent list holds (dense piles: a sphere touches up to 12 neighbours)
#endif
#define RB_TILE_CAP (3 * RB_LS_BLOCK)      // triangles one block's tile holds (39 KB of shared memory in the list-step kernel at 256 rows)
#ifndef RB_COLLIDE_MINBLOCKS
#define RB_COLLIDE_MINBLOCKS 3
#endif
#ifndef RB_COLLIDE_MINBLOCKS_ENUM
#define RB_COLLIDE_MINBLOCKS_ENUM 4
#endif

template <bool RESOLVE>
__device__ __forceinline__ void ss_pair_single(const RbParams& P, BodyState& S, HitBuf& H, Tally& T,
                                               uint32_t A, uint32_t gA, float4 oa, uint32_t Bl, uint32_t gB) {
    const bool lo_is_a = gA < gB;
    if (!RESOLVE && !lo_is_a) return;
    float4 wb = P.ws[Bl];                       // partner at its frozen pose
    if (wb.w < 0.f) return;                     // it has left the broadphase since the list was built (Q3 / Q7)
    float rb = P.sph_obj[Bl].w;
    f3 ca = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
    f3 cb = xyz(wb);
    f3 cn = lo_is_a ? sub3(ca, cb) : sub3(cb, ca);          // lower id plays sphereA / modelA
    float dist = mag3(cn);
    float radii = lo_is_a ? (oa.w + rb) : (rb + oa.w);
    if (lo_is_a) T.tests_ss++;
    if (!(radii - dist >= 0)) return;
    f3 n1 = mk3(cn.x / dist, cn.y / dist, cn.z / dist);
    if (lo_is_a) { T.hits_ss++; record_hit(P, H, gA, gB, radii - dist, n1, oa.w); }
    if (RESOLVE) {
        f3 n = n1;
        if (!lo_is_a) { n = neg3(n); n = normalize3(n); }
        float sp = mag3(S.v);
        S.v = mul3(mul3(n, sp), P.ss_damp);
        body_set_position(S, S.pt);
    }
}

// MODE 0: everything in one launch (per-tick enumeration), one thread per sorted sphere.
// MODE 1: list rebuild -- enumerate only (P1 + P3 + the exact-predicate pruning), one thread per body ROW: partners and the
//         row's slice of its block's triangle tile are parked for rb_list_step_kernel; rows that cannot have a list (more
//         partners than a list holds, spheres wider than the fast walk, a block whose tile would overflow) are marked and
//         handed to MODE 3 from this tick on.
// MODE 3: the hot bodies of a list tick (no usable list: fast movers, new in the grid, marked rows), one thread each,
//         enumerating for themselves like MODE 0 but against the sort of the last rebuild + the hot-body table.
// All modes run the same tests in the same canonical order as rb_list_step_kernel: results are bit-identical.
template <bool RESOLVE, int MODE>
__device__ __forceinline__ void collide_single_rows(const RbParams& P, const uint32_t tid, const int phase) {
    __shared__ uint32_t s_cand[RB_CAND_MAX][RB_BLOCK];
    const uint32_t lane = threadIdx.x & 31;
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    HitBuf H; H.n = 0;
    if (MODE == 1 && !P.vl[RB_VL_REBUILD]) return;       // persistent lists still valid: nothing to enumerate
    bool valid;
    uint32_t A = 0;
    float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f);
    bool had_contact = false;
    if (MODE == 1) {
        // one thread per body ROW (the rows of a block share a triangle tile); the own sphere is grown by the skin like
        // the entries of the sort it is matched against
        A = tid;
        valid = tid < rb_n_owned(P) && P.key[tid] != RB_NONE;
        if (valid) { f0 = P.ws[tid]; f0.w += P.vl_skin; }
    } else if (MODE == 3) {
        // one thread per hot body of this tick: it enumerates for itself -- partners from the sort of the last rebuild,
        // whose entries are off by less than the skin they were grown by, plus the other hot bodies (and, in slab worlds,
        // the neighbours' ghosts), triangles from the static grid -- and then steps exactly like the fused kernel
        const uint32_t nh = min(*P.nhot, P.hot_cap);
        valid = tid < nh;
        if (valid) { A = P.hot_row[tid]; valid = A != RB_NONE; }      // (an entry whose row emigrated after K1 linked it)
        // slab worlds split the tick around the halo exchange like the list step: early, the hot rows of the blocks no ghost can
        // reach; late, the rows of every block the early launches did not step
        if (valid && phase == 1) valid = P.blk_bnd[A / RB_LS_BLOCK] != P.tick;
        if (valid && phase == 2) valid = P.blk_done[A / RB_LS_BLOCK] != P.tick;
        if (valid) { f0 = P.ws[A]; had_contact = (P.flags[A] & RB_F_CONTACT) != 0; }
    } else {
        const uint32_t n_sorted = P.col_start[P.scx * P.scz];
        valid = tid < n_sorted;
        if (valid) {
            f0 = P.sorted[tid];
            uint32_t sid = P.sorted_id[tid];
            A = sid & ~RB_TRI_BIT; had_contact = (sid & RB_TRI_BIT) != 0;
            valid = A < rb_n_owned(P);
        }
    }
    // body state is loaded lazily: a body with no partner and no staged triangle needs none of it
    // (a body that was in contact last tick will almost surely need it: start the gathers now)
    if (MODE != 1 && valid && had_contact) {
        rb_prefetch(P.pos + A); rb_prefetch(P.vel + A); rb_prefetch(P.sph_obj + A);
    }
    const bool lazy = MODE == 3 && P.lazy_mt;      // list ticks of W.t = 0 worlds: model translation = 0 + p, never loaded or stored
    BodyState S; S.changed = false; S.flags = 0;
    S.p = S.v = S.mt = S.pt = S.wt = mk3(0.f, 0.f, 0.f);
    float4 oa = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t gA = valid ? rb_gid(P, A) : 0u;
    const f3 fc = xyz(f0);
    const float frm = f0.w;

    // ---------------- P1: partner scan ----------------
    constexpr int NPG = (MODE == 1) ? RB_SSLIST_MAX : RB_SSCAND_MAX;      // (the enumeration keeps as many as a list holds)
    uint32_t pg[NPG], pl[NPG];
#pragma unroll
    for (int k = 0; k < NPG; ++k) { pg[k] = RB_NONE; pl[k] = RB_NONE; }
    uint32_t npart = 0;
    {
        // every lane walks its (up to 3) z-rows entry by entry; the loop is warp-uniform so the lanes
        // stay in lock step (independent thread scheduling would otherwise let sub-groups run ahead)
        uint32_t cz = 1, cz1 = 0, cx0 = 0, cx1 = 0, i = 0, e = 0, ni = 0, ne = 0;
        if (valid) {
            float reach = frm + P.rm_max;
            const float fu = rb_gu(P, fc.x, fc.z), fv = rb_gv(P, fc.x, fc.z);      // (u: the axis whose columns are contiguous in the sort)
            cx0 = rb_col(fu - reach, P.g_off_u, P.s_inv_w, P.g_nu); cx1 = rb_col(fu + reach, P.g_off_u, P.s_inv_w, P.g_nu);
            cz = rb_col(fv - reach, P.g_off_v, P.s_inv_w, P.g_nv); cz1 = rb_col(fv + reach, P.g_off_v, P.s_inv_w, P.g_nv);
            i = P.col_start[cz * P.g_nu + cx0]; e = P.col_start[cz * P.g_nu + cx1 + 1];
            if (cz < cz1) { ni = P.col_start[(cz + 1) * P.g_nu + cx0]; ne = P.col_start[(cz + 1) * P.g_nu + cx1 + 1]; }   // next row, in flight early
        }
        while (true) {
            while (i >= e && cz < cz1) {
                ++cz; i = ni; e = ne;
                if (cz < cz1) { ni = P.col_start[(cz + 1) * P.g_nu + cx0]; ne = P.col_start[(cz + 1) * P.g_nu + cx1 + 1]; }
            }
            const bool work = i < e;
            if (!__any_sync(0xffffffffu, work)) break;
            if (work) {
                // up to 4 entries of the row per iteration, all requested before the first is looked at
                const uint32_t nb4 = min(e - i, 4u);
                float4 ov[4];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) ov[u] = P.sorted[min(i + u, e - 1)];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    if (u < nb4) {
                        const float4 o = ov[u];
                        float dx = o.x - fc.x, dy = o.y - fc.y, dz = o.z - fc.z, rr = frm + o.w;
                        if (!((dx * dx + dy * dy) + dz * dz > rr * rr)) {
                            uint32_t Bl = P.sorted_id[i + u] & ~RB_TRI_BIT;
                            if (Bl != A) {
                                T.cand_ss++;
                                uint32_t g = rb_gid(P, Bl);
                                if (RESOLVE || g > gA) {           // detection: the lower id reports the pair
                                    npart++;
                                    // keep the NPG smallest ids, ascending
#pragma unroll
                                    for (int k = 0; k < NPG; ++k)
                                        if (g < pg[k]) { uint32_t tg = pg[k], tl = pl[k]; pg[k] = g; pl[k] = Bl; g = tg; Bl = tl; }
                                }
                            }
                        }
                    }
                }
                i += nb4;
            }
            __syncwarp();
        }
    }
    if (MODE == 3) {
        // hot bodies (and, in slab worlds, the neighbours' ghosts) are in nobody's list and may have left the place the
        // old sort has them at: everybody looks them up in the coarse cell table (<= 2 x 2 cells)
        uint32_t hcell[4] = { 0, 0, 0, 0 }, hn = 0;
        if (valid) {
            const float reach = frm + P.rm_max;
            const uint32_t hx0 = rb_col(fc.x - reach, P.half, P.hc_inv_w, P.hcn), hx1 = rb_col(fc.x + reach, P.half, P.hc_inv_w, P.hcn);
            const uint32_t hz0 = rb_col(fc.z - reach, P.half, P.hc_inv_w, P.hcn), hz1 = rb_col(fc.z + reach, P.half, P.hc_inv_w, P.hcn);
            hcell[0] = hz0 * P.hcn + hx0; hn = 1;
            if (hx1 != hx0) hcell[hn++] = hz0 * P.hcn + hx1;
            if (hz1 != hz0) { hcell[hn++] = hz1 * P.hcn + hx0; if (hx1 != hx0) hcell[hn++] = hz1 * P.hcn + hx1; }
        }
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            uint32_t h = (q < hn) ? hot_cell_head(P, hcell[q]) : 0u;
            while (true) {
                if (!__any_sync(0xffffffffu, h != 0u)) break;
                if (h) {
                    const uint32_t Bl = P.hot_row[h - 1u];
                    h = P.hot_next[h - 1u];
                    if (Bl != A && Bl != RB_NONE) {
                        const float4 o = P.ws[Bl];
                        float dx = o.x - fc.x, dy = o.y - fc.y, dz = o.z - fc.z, rr = frm + o.w;
                        if (o.w >= 0.f && !((dx * dx + dy * dy) + dz * dz > rr * rr)) {
                            uint32_t g = rb_gid(P, Bl), bl = Bl;
                            bool dup = false;
#pragma unroll
                            for (int k = 0; k < RB_SSCAND_MAX; ++k) dup |= (pg[k] == g);
                            if (!dup) {
                                T.cand_ss++; npart++;
#pragma unroll
                                for (int k = 0; k < RB_SSCAND_MAX; ++k)
                                    if (g < pg[k]) { uint32_t tg = pg[k], tl = pl[k]; pg[k] = g; pl[k] = bl; g = tg; bl = tl; }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    // ---------------- P3: stage sphere-triangle candidates ----------------
    // Column entries carry the triangle's box (32 B, contiguous per column), so the prune needs no
    // vertex gather: survivors are staged, then sorted + de-duplicated (ascending global id = the
    // canonical order; a triangle registered in two columns shows up twice).
    uint32_t nc = 0;
    bool st_slow = false;
    {
        uint32_t cur[4], end[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) cur[k] = end[k] = 0;
        if (valid && P.nt) {
            uint32_t cx0 = rb_col(fc.x - frm, P.half, P.t_inv_w, P.tcx), cx1 = rb_col(fc.x + frm, P.half, P.t_inv_w, P.tcx);
            uint32_t cz0 = rb_col(fc.z - frm, P.half, P.t_inv_w, P.tcz), cz1 = rb_col(fc.z + frm, P.half, P.t_inv_w, P.tcz);
            uint32_t nx = cx1 - cx0 + 1, ncol = nx * (cz1 - cz0 + 1);
            if (ncol > 4) st_slow = true;
            else {
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                    if (k < ncol) {
                        uint32_t col = (cz0 + k / nx) * P.tcx + cx0 + k % nx;
                        float2 yr = __ldg(P.tcol_yr + col);
                        if (!(yr.x > fc.y + frm || yr.y < fc.y - frm)) {      // column holds something at this height
                            cur[k] = __ldg(P.tcol_start + col); end[k] = __ldg(P.tcol_start + col + 1);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            while (true) {      // warp-uniform walk of column k, one entry per lane and iteration
                const bool work = cur[k] < end[k];
                if (!__any_sync(0xffffffffu, work)) break;
                if (work) {
                    // two entries per iteration (an aligned heightfield column holds exactly two), both requested up front
                    const uint32_t n2 = min(end[k] - cur[k], 2u);
                    const float4* ep = P.tcol_ent + 2ull * cur[k];
                    float4 ea0 = __ldg(ep), ea1 = __ldg(ep + 1);          // (minx,maxx,minz,maxz) (miny,maxy,id,-)
                    float4 eb0 = ea0, eb1 = ea1;
                    if (n2 == 2) { eb0 = __ldg(ep + 2); eb1 = __ldg(ep + 3); }
#pragma unroll
                    for (uint32_t u = 0; u < 2; ++u) {
                        if (u < n2) {
                            const float4 e0 = u ? eb0 : ea0, e1 = u ? eb1 : ea1;
                            T.cand_st++;
                            bool keep = !(e0.x > fc.x + frm || e0.y < fc.x - frm) && !(e0.z > fc.z + frm || e0.w < fc.z - frm) &&
                                        !(e1.x > fc.y + frm || e1.y < fc.y - frm);
                            if (keep) {
                                if (nc == RB_CAND_MAX) st_slow = true;        // too many: the generic walk takes over
                                else {
                                    const uint32_t tcand = __float_as_uint(e1.z);
                                    s_cand[nc++][threadIdx.x] = tcand;
                                    rb_prefetch(P.tri + 3ull * tcand); rb_prefetch(P.tri + 3ull * tcand + 2);     // the exact test gathers these 48 B later
                                }
                            }
                        }
                    }
                    cur[k] = st_slow ? end[k] : cur[k] + n2;
                }
                __syncwarp();
            }
        }
        if (st_slow) nc = 0;
        // insertion sort + de-duplication of the (few) staged ids, in this thread's shared-memory column
        for (uint32_t i = 1; i < nc; ++i) {
            uint32_t v = s_cand[i][threadIdx.x];
            int j = (int)i - 1;
            while (j >= 0 && s_cand[j][threadIdx.x] > v) { s_cand[j + 1][threadIdx.x] = s_cand[j][threadIdx.x]; --j; }
            s_cand[j + 1][threadIdx.x] = v;
        }
        uint32_t nu = 0;
        for (uint32_t i = 0; i < nc; ++i) {
            uint32_t v = s_cand[i][threadIdx.x];
            if (nu == 0 || s_cand[nu - 1][threadIdx.x] != v) s_cand[nu++][threadIdx.x] = v;
        }
        nc = nu;
        __syncwarp();
    }
    if (MODE == 1) {     // (always: a block's tile is sized for pruned lists -- unpruned ones overflow it, the block's rows lose their lists, and that asks for the next rebuild)
        // Persistent lists keep only the triangles the sphere can actually reach while its list lives: the
        // exact predicate, run once at the build pose with the radius grown by the skin (a body steps from
        // its list only while displacement + margin <= skin, and the distance to a triangle is 1-Lipschitz
        // in the centre) plus a slack far above the predicate's rounding error. What is dropped here could
        // not have tested positive, so the hits -- and with them every result -- are unchanged.
        // Pooled over the warp like P4 (no idle lanes).
        __shared__ uint32_t s_keep[RB_BLOCK];
        const uint32_t wbase = threadIdx.x & ~31u;
        const uint32_t incl = warp_incl_scan(nc, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const float prad = frm + P.vl_slack;
        s_keep[threadIdx.x] = 0;
        __syncwarp();
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t item = base + lane;
            uint32_t lo = 0;
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) { uint32_t v = __shfl_sync(0xffffffffu, incl, (lo + st - 1) & 31); if (v <= item) lo += st; }
            const uint32_t owner = lo & 31;
            const uint32_t o_excl = __shfl_sync(0xffffffffu, incl - nc, owner);
            const float ox = __shfl_sync(0xffffffffu, fc.x, owner), oy = __shfl_sync(0xffffffffu, fc.y, owner), oz = __shfl_sync(0xffffffffu, fc.z, owner);
            const float orad = __shfl_sync(0xffffffffu, prad, owner);
            if (item < total) {
                const uint32_t k = item - o_excl;
                const uint32_t t = s_cand[k][wbase + owner];
                const float4* tp = P.tri + 3ull * t;
                float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
                if (sphere_triangle_detect(mk3(ox, oy, oz), orad, xyz(a4), xyz(b4), xyz(c4))) atomicOr(&s_keep[wbase + owner], 1u << k);
            }
            __syncwarp();
        }
        __syncwarp();
        const uint32_t m = s_keep[threadIdx.x];
        uint32_t nu = 0;
        for (uint32_t j = 0; j < nc; ++j) if ((m >> j) & 1u) s_cand[nu++][threadIdx.x] = s_cand[j][threadIdx.x];
        T.tests_st += nc;
        nc = nu;
    }
    if (MODE == 1) {
        // Park the lists. Partners: slot-major arrays by row. Triangles: the rows of this block pack theirs, in row order,
        // into one contiguous region of the tile pool -- vertex records first, global ids behind them -- which the list-step
        // kernel stages with a single bulk copy. A row that cannot have a list steps as a hot body from now on.
        __shared__ uint32_t s_total, s_base, s_over;
        bool no_list = valid && (npart > RB_SSLIST_MAX || st_slow);
        const uint32_t mine = (valid && !no_list) ? nc : 0u;
        const uint32_t off = block_excl_scan(mine, &s_total);
        if (threadIdx.x == 0) {
            const uint32_t total = s_total;
            uint32_t over = total > RB_TILE_CAP ? 1u : 0u, base = 0u;
            if (!over && total) {
                const uint32_t units = 3u * total + ((total + 3u) >> 2);
                base = atomicAdd(P.tile_cursor, units);
                if (base + units > P.tile_pool_units) over = 1u;       // pool exhausted (sized for 3 triangles per row on average)
            }
            s_base = base; s_over = over;
            P.blk_tile[blockIdx.x] = make_uint2(base, over ? 0u : total);
        }
        __syncthreads();
        if (valid) {
            if (no_list) atomicAdd(&P.vl[RB_VL_NOLIST + (npart > RB_SSLIST_MAX ? 0 : 1)], 1u); else if (s_over) atomicAdd(&P.vl[RB_VL_NOLIST + 2], 1u);
            no_list = no_list || s_over != 0u;
            uint32_t hdr = 0u;
            if (!no_list) {
                hdr = npart | (nc << 8) | (off << 16);
                uint4* reg = P.tile_pool + s_base;
                uint32_t* ids = reinterpret_cast<uint32_t*>(reg + 3ull * s_total);
                for (uint32_t j = 0; j < nc; ++j) {
                    const uint32_t t = s_cand[j][threadIdx.x];
                    const uint4* tp = reinterpret_cast<const uint4*>(P.tri + 3ull * t);
                    uint4* dst = reg + 3ull * (off + j);
                    dst[0] = __ldg(tp); dst[1] = __ldg(tp + 1); dst[2] = __ldg(tp + 2);
                    ids[off + j] = t;
                }
#pragma unroll
                for (int k = 0; k < NPG; ++k) if (pg[k] != RB_NONE) { P.e_pg[(size_t)k * P.ns + A] = pg[k]; P.e_pl[(size_t)k * P.ns + A] = pl[k]; }
            } else {
                // no list for this row: hot from this tick on (its own launch enumerates for it every tick)
                P.ws_build[A].w = -1.0f;
                const uint32_t kA = P.key[A];
                if (!(kA & RB_HOT_BIT)) { P.key[A] = kA | RB_HOT_BIT; hot_push(P, A, fc.x, fc.z); }      // (K1 may have pushed it already: slab worlds rebuild on ticks that classify)
            }
            P.e_hdr[A] = hdr;
        } else if (tid < rb_n_owned(P)) P.e_hdr[A] = 0u;      // not in the broadphase: no list (should it enter the grid it steps as a hot body)
        collide_epilogue(P, H, T, lane);
        return;
    }
    // ---------------- lazy state load ----------------
    const bool need = valid && (npart || nc || st_slow);
    if (need) {
        S.p = xyz(P.pos[A]); S.v = xyz(P.vel[A]);
        if (lazy) S.mt = mk3(0.f + S.p.x, 0.f + S.p.y, 0.f + S.p.z); else S.mt = xyz(P.mt[A]);
        if (npart) S.pt = xyz(P.pt[A]);      // only a sphere-sphere hit reads the previous pose (snap-back); every other path overwrites it first
        if (P.wt) S.wt = xyz(P.wt[A]);
        S.flags = P.flags[A];
        oa = P.sph_obj[A];
    }
    const uint32_t flags_in = S.flags;
    const f3 mt_frozen = S.mt;

    // ---------------- P2: sphere-sphere in ascending partner id ----------------
#pragma unroll
    for (int k = 0; k < RB_SSCAND_MAX; ++k) {
        if (!__any_sync(0xffffffffu, pg[k] != RB_NONE)) break;
        if (pg[k] != RB_NONE) ss_pair_single<RESOLVE>(P, S, H, T, A, gA, oa, pl[k], pg[k]);
        __syncwarp();
    }
    if (npart > RB_SSCAND_MAX) {            // crowded: continue above the last staged id by rescanning
        uint32_t lo = pg[RB_SSCAND_MAX - 1] + 1;
        while (true) {
            uint32_t best = RB_NONE, best_local = RB_NONE, eligible = 0;
            scan_partners<true>(P, A, fc, frm, lo, best, best_local, eligible, T);
            if (MODE == 3) scan_hot_partners(P, A, fc, frm, lo, best, best_local, eligible);
            if (best == RB_NONE) break;
            ss_pair_single<RESOLVE>(P, S, H, T, A, gA, oa, best_local, best);
            lo = best + 1;
            if (eligible <= 1) break;
        }
    }

    // ---------------- P4: lock-step exact tests at the current pose ----------------
    bool hit_any = false;
    if (st_slow) {
        T.cand_st = 0;      // the generic walk recounts
        sphere_vs_triangles<RESOLVE>(P, S, H, T, gA, xyz(oa), oa.w, fc, frm, hit_any);
    }
    // Speculative, load-balanced rounds. Until a sphere is hit its pose does not change, so ALL of its
    // remaining staged triangles can be tested at once: every round pools the remaining tests of the
    // warp, deals them out 32 at a time (the owner's pose travels by shuffle), and each owner then takes
    // its FIRST hit in canonical order, resolves it, and re-enters the next round with the triangles
    // behind that hit -- exactly the sequential test/resolve sequence, without idle lanes.
    {
        __shared__ uint32_t s_mask[RB_BLOCK];
        const uint32_t wbase = threadIdx.x & ~31u;
        uint32_t j0 = 0;
        while (true) {
            const uint32_t rem = j0 < nc ? nc - j0 : 0;
            if (!__any_sync(0xffffffffu, rem != 0)) break;
            const uint32_t incl = warp_incl_scan(rem, lane);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const f3 c = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);       // Sphere::getPosition at the current pose
            s_mask[threadIdx.x] = 0;
            __syncwarp();
            for (uint32_t base = 0; base < total; base += 32) {
                const uint32_t item = base + lane;
                uint32_t lo = 0;                                   // owner = number of lanes whose inclusive count is <= item
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1) { uint32_t v = __shfl_sync(0xffffffffu, incl, (lo + st - 1) & 31); if (v <= item) lo += st; }
                const uint32_t owner = lo & 31;
                const uint32_t o_excl = __shfl_sync(0xffffffffu, incl - rem, owner), o_j0 = __shfl_sync(0xffffffffu, j0, owner);
                const float ox = __shfl_sync(0xffffffffu, c.x, owner), oy = __shfl_sync(0xffffffffu, c.y, owner), oz = __shfl_sync(0xffffffffu, c.z, owner);
                const float orad = __shfl_sync(0xffffffffu, oa.w, owner);
                if (item < total) {
                    const uint32_t k = item - o_excl;
                    const uint32_t t = s_cand[o_j0 + k][wbase + owner];
                    const float4* tp = P.tri + 3ull * t;
                    float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
                    if (sphere_triangle_detect(mk3(ox, oy, oz), orad, xyz(a4), xyz(b4), xyz(c4))) atomicOr(&s_mask[wbase + owner], 1u << k);
                }
                __syncwarp();
            }
            __syncwarp();
            uint32_t m = s_mask[threadIdx.x];
            if (rem) {
                // sequential accounting: the tests up to and including the first hit are the ones the
                // one-at-a-time order would have run at this pose
                const uint32_t first = m ? (uint32_t)__ffs(m) - 1 : rem;
                T.tests_st += (RESOLVE && m) ? first + 1 : rem;
                if (!RESOLVE) j0 = nc; else j0 = m ? j0 + first + 1 : nc;
                const uint32_t jbase = RESOLVE ? j0 - (m ? first + 1 : 0) : 0;
                while (m) {                                        // RESOLVE: exactly one pass; detection: every hit
                    const uint32_t k = (uint32_t)__ffs(m) - 1;
                    m = RESOLVE ? 0u : (m & (m - 1));
                    const uint32_t t = s_cand[(RESOLVE ? jbase : 0) + k][threadIdx.x];
                    const float4* tp = P.tri + 3ull * t;
                    float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
                    f3 TA = xyz(a4), TB = xyz(b4), TC = xyz(c4);
                    hit_any = true;
                    T.hits_st++;
                    f3 q = closest_point(c, TA, TB, TC);
                    f3 ov = sub3(c, q);
                    float dist = mag3(ov);
                    f3 o = mk3(ov.x / dist, ov.y / dist, ov.z / dist);
                    record_hit(P, H, gA, RB_TRI_BIT | t, oa.w - dist, o, oa.w);
                    if (RESOLVE) {
                        // triangle normal normalize((C-A)x(B-A)) is static: pre-computed at build (w lanes)
                        f3 normal = mk3(a4.w, b4.w, c4.w);
                        f3 delta = sub3(add3(mul3(o, oa.w * P.pushout), q), c);
                        float ncomp = dot3(S.v, normal);
                        f3 rv = sub3(S.v, mul3(normal, ncomp));
                        body_set_position(S, add3(S.p, delta));
                        S.v = rv;
                        S.flags |= RB_F_CONTACT;
                    }
                }
            }
            __syncwarp();
        }
    }
    // ---------------- P5: flags, write-back ----------------
    if (RESOLVE && valid) {
        if (need) {
            if ((flags_in & RB_F_CONTACT) && !hit_any) S.flags &= ~RB_F_CONTACT;     // Physics.cpp:201-207
            float disp = mag3(sub3(S.mt, mt_frozen));
            if (disp > frm - oa.w) T.overflow++;
            if (S.changed) {
                // (single-sphere bodies keep their world radius in pos.w)
                P.pos[A] = make_float4(S.p.x, S.p.y, S.p.z, oa.w);
                P.vel[A] = make_float4(S.v.x, S.v.y, S.v.z, 0.f);
                if (!lazy) P.mt[A] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f);
                P.pt[A] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
            }
            if (S.flags != flags_in) P.flags[A] = S.flags;
        } else if (had_contact) {
            atomicAnd(&P.flags[A], ~RB_F_CONTACT);       // touched nothing: lose the contact flag, no state gather
        }
    }
    if (MODE == 3 && P.pos_out && valid) {
        // the host asked for this tick's positions in id order: every row is written exactly once, by the launch that
        // owns it this tick (hot rows here, everything else by the list-step kernel)
        f3 pf = S.p;
        if (!need) pf = xyz(P.pos[A]);
        float* po = P.pos_out + (size_t)P.pos_out_stride * rb_gid(P, A);
        po[0] = pf.x; po[1] = pf.y; po[2] = pf.z;
        if (P.pos_out_stride == 4) po[3] = 1.0f;
    }
    collide_epilogue(P, H, T, lane);
    if (MODE == 0 && P.vl && P.vl_host && tid == 0) P.vl_host[2] = P.vl[RB_VL_NFAST];     // lists switched off: K1's count of fast movers
}
template <bool RESOLVE, int MODE>
__global__ void __launch_bounds__(RB_BLOCK, (MODE == 1) ? RB_COLLIDE_MINBLOCKS_ENUM : RB_COLLIDE_MINBLOCKS) rb_collide_single_kernel(const RbParams P, const int phase) {
    if (MODE == 3) {
        if (phase == 1 && P.vl[RB_VL_REBUILD]) return;      // K1 has asked for a rebuild (rows are about to move): everything steps after the exchange
        // the grid is sized from the hot count the host saw last (several ticks old): stride over the table, whatever its size.
        // Everything a pass shares lives per warp (staging columns, hit masks), so passes need no block barrier in between.
        const uint32_t nh = min(*P.nhot, P.hot_cap);
        for (uint32_t first = blockIdx.x * blockDim.x; first < nh; first += gridDim.x * blockDim.x)
            collide_single_rows<RESOLVE, MODE>(P, first + threadIdx.x, phase);
    } else collide_single_rows<RESOLVE, MODE>(P, blockIdx.x * blockDim.x + threadIdx.x, 0);
}


#ifndef RB_CHAIN_TU
// ------------------------------------------------------------------------------------------
// K4G: compound bodies (multi-sphere entities: Physics.cpp:102-138 loops spheresA x spheresB, Geometry.cpp:27-31 moves
// every sphere of an entity by the same model matrix). Each body gets a group of L = 2^lg lanes (L >= its sphere count
// whenever that is <= 32), lane j owns spheres j, j + L, ...: the enumeration -- the partner scan of the sphere grid and the
// walk of the triangle columns, both pointer chasing -- runs per SPHERE, like the single-sphere kernel, while the body's
// state is replicated over the lanes of its group and every resolution is computed by all of them (same operands, same
// result), so nothing has to be sent around. The sequential test / resolve order of the canonical step is kept by
// speculation: until something hits, the pose does not change, so the lanes test everything that is left at once, the FIRST
// hit in canonical order is resolved, what came before it is final and what comes after it is tested again.
// Same canonical order and arithmetic as rb_collide_kernel<false, .> (bit-identical results; RB_COLLIDE_V1=1 runs that one).
// ------------------------------------------------------------------------------------------
#define RB_GRP_PART 4
#define RB_SS_BLOCK 64              // threads per block of the sphere-sphere phase of the listed bodies (a few long chains: spread them over the SMs)
#define RB_CE_SLOW (1u << 16)      // rb_compound_enum_kernel's header word: the sphere is wider than the fast triangle walk
#define RB_CE_OVER (1u << 17)      // ... it found more partner bodies than a list holds

// reductions over the aligned group of L lanes a body owns; every lane of the WARP takes part (xor partners below L stay inside the group)
__device__ __forceinline__ uint32_t group_min(uint32_t v, const uint32_t L) {
    for (uint32_t d = 1; d < L; d <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ uint32_t group_sum(uint32_t v, const uint32_t L) {
    for (uint32_t d = 1; d < L; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// one pass over the sphere grid around frozen sphere (fc, frm): the RB_GRP_PART smallest partner BODY ids >= lo, ascending and
// unique; `over` = some id did not fit. Warp-uniform (every lane of the warp calls it; `work` says who has a sphere).
template <bool RESOLVE>
__device__ __forceinline__ void group_scan_sphere(const RbParams& P, const bool work_in, const uint32_t A, const uint32_t gA, const f3 fc, const float frm,
                                                  const uint32_t lo, uint32_t (&pg)[RB_GRP_PART], uint32_t (&pl)[RB_GRP_PART], bool& over, Tally& T) {
    uint32_t cz = 1, cz1 = 0, cx0 = 0, cx1 = 0, i = 0, e = 0;
    if (work_in) {
        const float reach = frm + P.rm_max;
        const float fu = rb_gu(P, fc.x, fc.z), fv = rb_gv(P, fc.x, fc.z);
        cx0 = rb_col(fu - reach, P.g_off_u, P.s_inv_w, P.g_nu); cx1 = rb_col(fu + reach, P.g_off_u, P.s_inv_w, P.g_nu);
        cz = rb_col(fv - reach, P.g_off_v, P.s_inv_w, P.g_nv); cz1 = rb_col(fv + reach, P.g_off_v, P.s_inv_w, P.g_nv);
        i = P.col_start[cz * P.g_nu + cx0]; e = P.col_start[cz * P.g_nu + cx1 + 1];
    }
    while (true) {
        while (i >= e && cz < cz1) { ++cz; i = P.col_start[cz * P.g_nu + cx0]; e = P.col_start[cz * P.g_nu + cx1 + 1]; }
        const bool work = i < e;
        if (!__any_sync(0xffffffffu, work)) break;
        if (work) {
            const uint32_t nb4 = min(e - i, 4u);
            float4 ov[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) ov[u] = P.sorted[min(i + u, e - 1)];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                if (u < nb4) {
                    const float4 o = ov[u];
                    float dx = o.x - fc.x, dy = o.y - fc.y, dz = o.z - fc.z, rr = frm + o.w;
                    if (!((dx * dx + dy * dy) + dz * dz > rr * rr)) {
                        uint32_t Bl = rb_sph_owner(P, P.sorted_id[i + u] & ~RB_TRI_BIT);
                        if (Bl != A) {
                            T.cand_ss++;
                            uint32_t g = rb_gid(P, Bl);
                            if (g >= lo && (RESOLVE || g > gA)) {          // detection: the lower id reports the pair
                                bool dup = false;
#pragma unroll
                                for (int k = 0; k < RB_GRP_PART; ++k) dup |= (pg[k] == g);
                                if (!dup) {
#pragma unroll
                                    for (int k = 0; k < RB_GRP_PART; ++k)
                                        if (g < pg[k]) { uint32_t tg = pg[k], tl = pl[k]; pg[k] = g; pl[k] = Bl; g = tg; Bl = tl; }
                                    if (g != RB_NONE) over = true;         // pushed off the end (or never fitted)
                                }
                            }
                        }
                    }
                }
            }
            i += nb4;
        }
        __syncwarp();
    }
}

// P3 of the single-sphere kernel for one sphere per lane: walk of the <= 4 triangle columns the grown frozen box touches, box
// prune on the column entries, survivors staged in this thread's shared-memory column, sorted + de-duplicated. Warp-uniform.
__device__ __forceinline__ void group_stage_triangles(const RbParams& P, const bool work_in, const f3 fc, const float frm,
                                                      uint32_t (*s_cand)[RB_BLOCK], uint32_t& nc, bool& st_slow, Tally& T) {
    nc = 0; st_slow = false;
    uint32_t cur[4], end[4];
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) cur[k] = end[k] = 0;
    if (work_in && P.nt) {
        uint32_t cx0 = rb_col(fc.x - frm, P.half, P.t_inv_w, P.tcx), cx1 = rb_col(fc.x + frm, P.half, P.t_inv_w, P.tcx);
        uint32_t cz0 = rb_col(fc.z - frm, P.half, P.t_inv_w, P.tcz), cz1 = rb_col(fc.z + frm, P.half, P.t_inv_w, P.tcz);
        uint32_t nx = cx1 - cx0 + 1, ncol = nx * (cz1 - cz0 + 1);
        if (ncol > 4) st_slow = true;
        else {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                if (k < ncol) {
                    uint32_t col = (cz0 + k / nx) * P.tcx + cx0 + k % nx;
                    float2 yr = __ldg(P.tcol_yr + col);
                    if (!(yr.x > fc.y + frm || yr.y < fc.y - frm)) { cur[k] = __ldg(P.tcol_start + col); end[k] = __ldg(P.tcol_start + col + 1); }
                }
            }
        }
    }
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        while (true) {
            const bool work = cur[k] < end[k];
            if (!__any_sync(0xffffffffu, work)) break;
            if (work) {
                const uint32_t n2 = min(end[k] - cur[k], 2u);
                const float4* ep = P.tcol_ent + 2ull * cur[k];
                float4 ea0 = __ldg(ep), ea1 = __ldg(ep + 1);
                float4 eb0 = ea0, eb1 = ea1;
                if (n2 == 2) { eb0 = __ldg(ep + 2); eb1 = __ldg(ep + 3); }
#pragma unroll
                for (uint32_t u = 0; u < 2; ++u) {
                    if (u < n2) {
                        const float4 e0 = u ? eb0 : ea0, e1 = u ? eb1 : ea1;
                        T.cand_st++;
                        bool keep = !(e0.x > fc.x + frm || e0.y < fc.x - frm) && !(e0.z > fc.z + frm || e0.w < fc.z - frm) &&
                                    !(e1.x > fc.y + frm || e1.y < fc.y - frm);
                        if (keep) {
                            if (nc == RB_CAND_MAX) st_slow = true;
                            else {
                                const uint32_t tcand = __float_as_uint(e1.z);
                                s_cand[nc++][threadIdx.x] = tcand;
                                rb_prefetch(P.tri + 3ull * tcand); rb_prefetch(P.tri + 3ull * tcand + 2);
                            }
                        }
                    }
                }
                cur[k] = st_slow ? end[k] : cur[k] + n2;
            }
            __syncwarp();
        }
    }
    if (st_slow) nc = 0;
    for (uint32_t i = 1; i < nc; ++i) {
        uint32_t v = s_cand[i][threadIdx.x];
        int j = (int)i - 1;
        while (j >= 0 && s_cand[j][threadIdx.x] > v) { s_cand[j + 1][threadIdx.x] = s_cand[j][threadIdx.x]; --j; }
        s_cand[j + 1][threadIdx.x] = v;
    }
    uint32_t nu = 0;
    for (uint32_t i = 0; i < nc; ++i) {
        uint32_t v = s_cand[i][threadIdx.x];
        if (nu == 0 || s_cand[nu - 1][threadIdx.x] != v) s_cand[nu++][threadIdx.x] = v;
    }
    nc = nu;
    __syncwarp();
}

// state of a body from one lane of its group to all of them
__device__ __forceinline__ void group_bcast_state(BodyState& S, const uint32_t gmask, const uint32_t src) {
    S.p.x = __shfl_sync(gmask, S.p.x, src); S.p.y = __shfl_sync(gmask, S.p.y, src); S.p.z = __shfl_sync(gmask, S.p.z, src);
    S.v.x = __shfl_sync(gmask, S.v.x, src); S.v.y = __shfl_sync(gmask, S.v.y, src); S.v.z = __shfl_sync(gmask, S.v.z, src);
    S.mt.x = __shfl_sync(gmask, S.mt.x, src); S.mt.y = __shfl_sync(gmask, S.mt.y, src); S.mt.z = __shfl_sync(gmask, S.mt.z, src);
    S.pt.x = __shfl_sync(gmask, S.pt.x, src); S.pt.y = __shfl_sync(gmask, S.pt.y, src); S.pt.z = __shfl_sync(gmask, S.pt.z, src);
    S.flags = __shfl_sync(gmask, S.flags, src);
    S.changed = __shfl_sync(gmask, (int)S.changed, src) != 0;
}

// The sphere-sphere phase of a body over the group of lanes it owns (warp-uniform: every lane of the warp takes part).
// PRELOAD: the partner lists start from what rb_compound_enum_kernel parked instead of a scan of the sphere grid.
template <bool RESOLVE, bool PRELOAD>
__device__ __forceinline__ void group_sphere_sphere(const RbParams& P, const bool valid, const uint32_t A, const uint32_t gA, const uint32_t s0, const uint32_t nsph,
                                                    BodyState& S, HitBuf& H, Tally& T, const uint32_t lg, const uint32_t nchunk_w) {
    const uint32_t L = 1u << lg, lane = threadIdx.x & 31, sub = lane & (L - 1);
    const bool leader = sub == 0;
    {
        uint32_t pg[RB_GRP_PART], pl[RB_GRP_PART];
#pragma unroll
        for (int k = 0; k < RB_GRP_PART; ++k) { pg[k] = RB_NONE; pl[k] = RB_NONE; }
        bool over = false, need = valid, first_fill = true;
        uint32_t lo = 0;
        // spheres of A inside the broadphase (the lower body id counts the tests: one per pair of such spheres)
        uint32_t nA_in = 0;
        for (uint32_t c = 0; c < nchunk_w; ++c) {
            const uint32_t i = (c << lg) + sub;
            const bool in = valid && i < nsph && P.key[s0 + i] != RB_NONE;
            nA_in += group_sum(in ? 1u : 0u, L);
        }
        while (true) {
            if (__any_sync(0xffffffffu, need)) {
                // (re)fill this lane's list: the smallest partner ids >= lo over the spheres it owns
                if (need) {
#pragma unroll
                    for (int k = 0; k < RB_GRP_PART; ++k) { pg[k] = RB_NONE; pl[k] = RB_NONE; }
                    over = false;
                }
                if (PRELOAD && first_fill) {
                    // the lists the enumeration kernel parked (the 4 smallest partner ids of every sphere, ascending)
                    for (uint32_t c = 0; c < nchunk_w; ++c) {
                        const uint32_t i = (c << lg) + sub;
                        if (!(need && i < nsph)) continue;
                        const uint32_t sa = s0 + i, h = P.ce_ph[sa], np = (h >> 8) & 0xffu;
                        if (h & RB_CE_OVER) over = true;
                        for (uint32_t k = 0; k < np; ++k) {
                            uint32_t g = P.ce_pg[(size_t)k * P.ns + sa], Bl = P.ce_pl[(size_t)k * P.ns + sa];
                            bool dup = false;
#pragma unroll
                            for (int q = 0; q < RB_GRP_PART; ++q) dup |= (pg[q] == g);
                            if (dup) continue;
#pragma unroll
                            for (int q = 0; q < RB_GRP_PART; ++q)
                                if (g < pg[q]) { uint32_t tg = pg[q], tl = pl[q]; pg[q] = g; pl[q] = Bl; g = tg; Bl = tl; }
                            if (g != RB_NONE) over = true;
                        }
                    }
                } else {
                    for (uint32_t c = 0; c < nchunk_w; ++c) {
                        const uint32_t i = (c << lg) + sub;
                        bool w = need && i < nsph && P.key[s0 + i] != RB_NONE;
                        float4 fa = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (w) fa = P.ws[s0 + i];
                        group_scan_sphere<RESOLVE>(P, w, A, gA, xyz(fa), fa.w, lo, pg, pl, over, T);
                    }
                }
                first_fill = false;
                need = false;
            }
            uint32_t mine = RB_NONE, mine_l = RB_NONE;
            if (valid) {
#pragma unroll
                for (int k = RB_GRP_PART - 1; k >= 0; --k) if (pg[k] != RB_NONE && pg[k] >= lo) { mine = pg[k]; mine_l = pl[k]; }
            }
            const uint32_t best = group_min(valid ? mine : RB_NONE, L);
            const bool more = valid && best != RB_NONE;
            if (!__any_sync(0xffffffffu, more)) break;
            // the pair (A, B = next partner): every loop below is warp-uniform (the groups of a warp stay in lock step)
            const uint32_t src = group_min((more && mine == best) ? lane : RB_NONE, L);
            const uint32_t B = __shfl_sync(0xffffffffu, mine_l, more ? src : lane);
            const bool lo_is_a = gA < best;
            const bool act = more && (RESOLVE || lo_is_a);
            uint32_t t0 = 0, nb_sph = 0, n_in = 1, npairs = 0;
            f3 bmt = mk3(0.f, 0.f, 0.f), bpt = bmt, bwt = bmt;
            bool b_toggled = false;
            if (act) {
                t0 = rb_body_s0(P, B); nb_sph = rb_body_ns(P, B);
                bmt = xyz(P.mt[B]); bpt = xyz(P.pt[B]);
                if (P.wt) bwt = xyz(P.wt[B]);
                n_in = lo_is_a ? nb_sph : nsph; npairs = nsph * nb_sph;
                if (lo_is_a && leader) {
                    uint32_t nB_in = 0;
                    for (uint32_t i = 0; i < nb_sph; ++i) nB_in += P.key[t0 + i] != RB_NONE;
                    T.tests_ss += nA_in * nB_in;
                }
            }
            uint32_t q_start = 0;
            while (true) {
                // every lane: its share of the pairs that are left, in canonical order, at the current poses
                uint32_t myq = RB_NONE;
                if (act) {
                    for (uint32_t q = q_start + sub; q < npairs; q += L) {
                        const uint32_t io = q / n_in, ii = q - io * n_in;
                        const uint32_t ia = lo_is_a ? io : ii, ib = lo_is_a ? ii : io;
                        const uint32_t sa = s0 + ia, sb = t0 + ib;
                        if (P.key[sa] == RB_NONE || P.key[sb] == RB_NONE) continue;
                        const float4 oa = P.sph_obj[sa], ob = P.sph_obj[sb];
                        const f3 ca = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
                        f3 cb;
                        if (!b_toggled) cb = xyz(P.ws[sb]); else cb = mk3(ob.x + bmt.x, ob.y + bmt.y, ob.z + bmt.z);
                        const f3 cn = lo_is_a ? sub3(ca, cb) : sub3(cb, ca);
                        const float dist = mag3(cn);
                        const float radii = lo_is_a ? (oa.w + ob.w) : (ob.w + oa.w);
                        if (!(radii - dist >= 0)) continue;
                        if (RESOLVE) { myq = q; break; }
                        // detection only: nothing moves, every hit is final
                        T.hits_ss++;
                        record_hit(P, H, rb_sph_gid(P, A, s0, ia), rb_sph_gid(P, B, t0, ib), radii - dist, mk3(cn.x / dist, cn.y / dist, cn.z / dist), oa.w);
                    }
                }
                if (!RESOLVE) break;
                const uint32_t qs = group_min(myq, L);
                if (!__any_sync(0xffffffffu, qs != RB_NONE)) break;
                if (qs != RB_NONE) {
                    // the first hit in canonical order, resolved by every lane of the group alike
                    const uint32_t io = qs / n_in, ii = qs - io * n_in;
                    const uint32_t ia = lo_is_a ? io : ii, ib = lo_is_a ? ii : io;
                    const uint32_t sa = s0 + ia, sb = t0 + ib;
                    const float4 oa = P.sph_obj[sa], ob = P.sph_obj[sb];
                    const f3 ca = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
                    f3 cb;
                    if (!b_toggled) cb = xyz(P.ws[sb]); else cb = mk3(ob.x + bmt.x, ob.y + bmt.y, ob.z + bmt.z);
                    const f3 cn = lo_is_a ? sub3(ca, cb) : sub3(cb, ca);
                    const float dist = mag3(cn);
                    const float radii = lo_is_a ? (oa.w + ob.w) : (ob.w + oa.w);
                    const f3 n1 = mk3(cn.x / dist, cn.y / dist, cn.z / dist);
                    if (lo_is_a && leader) { T.hits_ss++; record_hit(P, H, rb_sph_gid(P, A, s0, ia), rb_sph_gid(P, B, t0, ib), radii - dist, n1, oa.w); }
                    f3 n = n1;
                    if (!lo_is_a) { n = neg3(n); n = normalize3(n); }      // "Opposite reaction", normalised again
                    const float sp = mag3(S.v);
                    S.v = mul3(mul3(n, sp), P.ss_damp);
                    body_set_position(S, S.pt);                             // snap back to the previous pose
                    const f3 t = mk3(bwt.x + bpt.x, bwt.y + bpt.y, bwt.z + bpt.z);      // the partner does the same on its side
                    bpt = bmt; bmt = t; b_toggled = true;
                    q_start = qs + 1;
                }
                __syncwarp();
            }
            if (more) {
                lo = best + 1;
                need = over && pg[RB_GRP_PART - 1] < lo;      // this lane's list is used up and there was more
            }
            __syncwarp();
        }
    }
}

template <bool RESOLVE>
__global__ void __launch_bounds__(RB_BLOCK, 3) rb_collide_group_kernel(const RbParams P) {
    __shared__ uint32_t s_cand[RB_CAND_MAX][RB_BLOCK];
    const uint32_t lg = P.grp_lg, L = 1u << lg;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, sub = lane & (L - 1), gbase = lane & ~(L - 1);
    const bool leader = sub == 0;
    const uint32_t A = tid >> lg;
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    HitBuf H; H.n = 0;
    bool valid = A < rb_n_owned(P);                 // (uniform over the group)
    uint32_t s0 = 0, nsph = 0, flags_in = 0;
    if (valid) { flags_in = P.flags[A]; valid = (flags_in & RB_F_ACTIVE) != 0; }
    BodyState S; S.changed = false; S.flags = flags_in;
    S.p = S.v = S.mt = S.pt = S.wt = mk3(0.f, 0.f, 0.f);
    uint32_t gA = 0;
    if (valid) {
        s0 = rb_body_s0(P, A); nsph = rb_body_ns(P, A);
        S.p = xyz(P.pos[A]); S.v = xyz(P.vel[A]); S.mt = xyz(P.mt[A]); S.pt = xyz(P.pt[A]);
        if (P.wt) S.wt = xyz(P.wt[A]);
        gA = rb_gid(P, A);
    }
    const f3 mt_frozen = S.mt;
    const uint32_t nchunk_w = __reduce_max_sync(0xffffffffu, (nsph + L - 1) >> lg);      // chunks of L spheres, warp-wide maximum

    // ---------------- sphere-sphere: partner bodies in ascending id ----------------
    group_sphere_sphere<RESOLVE, false>(P, valid, A, gA, s0, nsph, S, H, T, lg, nchunk_w);

    // ---------------- sphere-triangle: own spheres ascending, triangles in ascending id ----------------
    bool last_hit = false;
    for (uint32_t c = 0; c < nchunk_w; ++c) {
        const uint32_t i = (c << lg) + sub;
        const bool in_body = valid && i < nsph;
        const uint32_t sa = s0 + i;
        const bool sv = in_body && P.key[sa] != RB_NONE;
        float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), oa = fa;
        if (sv) { fa = P.ws[sa]; oa = P.sph_obj[sa]; }
        uint32_t nc; bool st_slow;
        group_stage_triangles(P, sv, xyz(fa), fa.w, s_cand, nc, st_slow, T);
        const bool chunk_live = valid && (c << lg) < nsph;          // (uniform over the group)
        const bool slow = group_min(st_slow ? 0u : 1u, L) == 0u && chunk_live;
        bool my_hit = false;
        if (slow) {
            // a sphere wider than the fast walk (or with more candidates than a column of the staging area holds): the
            // leader takes this chunk's spheres one after the other through the generic walk and hands the state back
            if (leader) {
                const uint32_t i1 = min(nsph, (c + 1) << lg);
                for (uint32_t k = c << lg; k < i1; ++k) {
                    bool hk = false;
                    if (P.key[s0 + k] != RB_NONE) {
                        const float4 fk = P.ws[s0 + k], ok = P.sph_obj[s0 + k];
                        sphere_vs_triangles<RESOLVE>(P, S, H, T, rb_sph_gid(P, A, s0, k), xyz(ok), ok.w, xyz(fk), fk.w, hk);
                    }
                    if (k == nsph - 1) last_hit = hk;
                }
            }
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, slow)) {
            BodyState S2 = S; bool lh = last_hit;
            group_bcast_state(S2, 0xffffffffu, gbase);
            lh = __shfl_sync(0xffffffffu, (int)lh, gbase) != 0;
            if (slow) { S = S2; last_hit = lh; }
        }
        // Speculative rounds, warp-uniform (the groups of a warp stay in lock step): every lane tests what its sphere has left
        // at the current pose; per group, the lowest sphere with a hit takes its first one, every lane of the group resolves it
        // alike, the spheres in front of it are final and the ones behind it start over at the new pose.
        {
            __shared__ uint32_t s_mask[RB_BLOCK];
            const uint32_t wbase = threadIdx.x & ~31u;
            const bool live = chunk_live && !slow;
            uint32_t j0 = 0;
            bool done = !(live && sv && nc);
            while (true) {
                // the tests the warp has left are pooled and dealt out 32 at a time (the owner's pose travels by shuffle), like P4
                // of the single-sphere kernel: a body resting on four of its eight spheres keeps few lanes busy by itself
                const uint32_t rem = done ? 0u : nc - j0;
                if (!__any_sync(0xffffffffu, rem != 0u)) break;
                const uint32_t incl = warp_incl_scan(rem, lane);
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                const f3 cc = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);       // Sphere::getPosition at the current pose
                s_mask[threadIdx.x] = 0;
                __syncwarp();
                for (uint32_t base = 0; base < total; base += 32) {
                    const uint32_t item = base + lane;
                    uint32_t lo = 0;                                   // owner = number of lanes whose inclusive count is <= item
#pragma unroll
                    for (int st = 16; st >= 1; st >>= 1) { uint32_t v = __shfl_sync(0xffffffffu, incl, (lo + st - 1) & 31); if (v <= item) lo += st; }
                    const uint32_t owner = lo & 31;
                    const uint32_t o_excl = __shfl_sync(0xffffffffu, incl - rem, owner), o_j0 = __shfl_sync(0xffffffffu, j0, owner);
                    const float ox = __shfl_sync(0xffffffffu, cc.x, owner), oy = __shfl_sync(0xffffffffu, cc.y, owner), oz = __shfl_sync(0xffffffffu, cc.z, owner);
                    const float orad = __shfl_sync(0xffffffffu, oa.w, owner);
                    if (item < total) {
                        const uint32_t k = item - o_excl;
                        const uint32_t t = s_cand[o_j0 + k][wbase + owner];
                        const float4* tp = P.tri + 3ull * t;
                        const float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
                        if (sphere_triangle_detect(mk3(ox, oy, oz), orad, xyz(a4), xyz(b4), xyz(c4))) atomicOr(&s_mask[wbase + owner], 1u << k);
                    }
                    __syncwarp();
                }
                __syncwarp();
                uint32_t m = s_mask[threadIdx.x], my_t = 0;
                if (RESOLVE) { if (m) { m &= 0u - m; my_t = s_cand[j0 + (uint32_t)__ffs(m) - 1u][threadIdx.x]; } }      // (only the first hit counts: the pose changes with it)
                else {
                    for (uint32_t mm = m; mm; mm &= mm - 1u) {         // detection only: nothing moves, every hit is final
                        const uint32_t t = s_cand[j0 + (uint32_t)__ffs(mm) - 1u][threadIdx.x];
                        const float4* tp = P.tri + 3ull * t;
                        const float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
                        const f3 q = closest_point(cc, xyz(a4), xyz(b4), xyz(c4));
                        const f3 ov = sub3(cc, q);
                        const float dist = mag3(ov);
                        T.hits_st++; my_hit = true;
                        record_hit(P, H, rb_sph_gid(P, A, s0, i), RB_TRI_BIT | t, oa.w - dist, mk3(ov.x / dist, ov.y / dist, ov.z / dist), oa.w);
                    }
                }
                const uint32_t hits_all = __ballot_sync(0xffffffffu, m != 0);
                if (!RESOLVE || !hits_all) { if (!done) T.tests_st += nc - j0; break; }
                const uint32_t wmin = group_min(m ? lane : RB_NONE, L);
                const bool hits = wmin != RB_NONE;
                const uint32_t wl = hits ? wmin : lane;                           // the lowest sphere with a hit: first in canonical order
                const uint32_t first = m ? (uint32_t)__ffs(m) - 1u : 0u;
                if (!done && (!hits || lane < wl)) { T.tests_st += nc - j0; done = true; }     // nothing hit in front of it: final
                const uint32_t t = __shfl_sync(0xffffffffu, my_t, wl);
                const float wox = __shfl_sync(0xffffffffu, oa.x, wl), woy = __shfl_sync(0xffffffffu, oa.y, wl), woz = __shfl_sync(0xffffffffu, oa.z, wl), wr = __shfl_sync(0xffffffffu, oa.w, wl);
                if (hits) {
                    // resolved by every lane of the group alike (process_triangle above, same expressions)
                    const float4* tp = P.tri + 3ull * t;
                    const float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
                    const f3 TA = xyz(a4), TB = xyz(b4), TC = xyz(c4);
                    const f3 cw = mk3(wox + S.mt.x, woy + S.mt.y, woz + S.mt.z);
                    const f3 q = closest_point(cw, TA, TB, TC);
                    const f3 ov = sub3(cw, q);
                    const float dist = mag3(ov);
                    const f3 o = mk3(ov.x / dist, ov.y / dist, ov.z / dist);
                    if (lane == wl) {
                        T.tests_st += first + 1; T.hits_st++; my_hit = true;
                        record_hit(P, H, rb_sph_gid(P, A, s0, i), RB_TRI_BIT | t, wr - dist, o, wr);
                        j0 += first + 1;
                        done = j0 >= nc;
                    }
                    const f3 normal = mk3(a4.w, b4.w, c4.w);
                    const f3 delta = sub3(add3(mul3(o, wr * P.pushout), q), cw);
                    const float ncomp = dot3(S.v, normal);
                    const f3 rv = sub3(S.v, mul3(normal, ncomp));
                    body_set_position(S, add3(S.p, delta));
                    S.v = rv;
                    S.flags |= RB_F_CONTACT;
                }
                __syncwarp();
            }
            // did the body's LAST sphere find a triangle (Physics.cpp:201-207)?
            const uint32_t lsrc = gbase + ((nsph - 1) & (L - 1));
            const bool lh = __shfl_sync(0xffffffffu, (int)my_hit, valid ? lsrc : lane) != 0;
            if (live && ((nsph - 1) >> lg) == c) last_hit = lh;
        }
        __syncwarp();
    }

    if (RESOLVE && valid && leader) {
        if ((flags_in & RB_F_CONTACT) && !last_hit) S.flags &= ~RB_F_CONTACT;
        // margin audit: the candidate sets were built for displacements up to the smallest margin
        const float disp = mag3(sub3(S.mt, mt_frozen));
        if (nsph && disp > P.ws[s0].w - P.sph_obj[s0].w) T.overflow++;
        if (S.changed || S.flags != flags_in) {
            // partners read our frozen pose concurrently: commit after the kernel (K5)
            const uint32_t k = (uint32_t)atomicAdd(&P.counters[RB_C_NCHANGED], 1ull);
            P.ch_body[k] = A;
            P.ch_pos[k] = make_float4(S.p.x, S.p.y, S.p.z, 1.f);
            P.ch_vel[k] = make_float4(S.v.x, S.v.y, S.v.z, 0.f);
            P.ch_mt[k] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f);
            P.ch_pt[k] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
            P.ch_flags[k] = S.flags;
        }
    }
    collide_epilogue(P, H, T, lane);
}

// ------------------------------------------------------------------------------------------
// K4E + K4R: compound bodies, the enumeration split from the resolution. What K4G's capture showed: the
// enumeration -- pointer chasing through the sphere grid and the triangle columns -- is what gains from a lane per SPHERE;
// the sequential test / resolve chain of a body does not. So:
//   K4E  one thread per sphere row: the partner scan (the 4 smallest partner body ids, ascending, + "there were more") and the
//        triangle walk (box-pruned ids, ascending, unique) of rb_collide_group_kernel, parked in slot-major global arrays;
//   K4R  one thread per body, the first-generation kernel with every grid walk replaced by a read of those lists: partner
//        bodies in ascending id merged over the body's spheres, then own spheres ascending x triangles ascending, each tested
//        at the current pose and resolved at once. A sphere whose partner list overflowed re-scans above the last id it
//        holds; a sphere wider than the fast triangle walk takes the generic walk.
// Same candidates, same canonical order, same arithmetic as rb_collide_kernel<false, .>: bit-identical results.
// ------------------------------------------------------------------------------------------
template <int PART>
__global__ void __launch_bounds__(RB_BLOCK, 3) rb_compound_enum_kernel(const RbParams P, const uint32_t nrows) {
    // PART 1: the partner scan (what the sphere-sphere phase of the listed bodies waits for); PART 2: the triangle lists. Two
    // launches, so that phase starts after the first and runs next to the second (and next to the others' triangle phase).
    __shared__ uint32_t s_cand[PART == 2 ? RB_CAND_MAX : 1][RB_BLOCK];
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    HitBuf H; H.n = 0;
    const uint32_t n_own_rows = P.kstride ? rb_n_owned(P) * P.kstride : nrows;
    const bool valid = s < n_own_rows && P.key[s] != RB_NONE;      // (inactive bodies, spheres outside the world cube and unused slots carry no key)
    uint32_t A = 0, gA = 0;
    float4 fa = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) { A = rb_sph_owner(P, s); gA = rb_gid(P, A); fa = P.ws[s]; }
    if (PART == 1) {
        uint32_t pg[RB_GRP_PART], pl[RB_GRP_PART];
#pragma unroll
        for (int k = 0; k < RB_GRP_PART; ++k) { pg[k] = RB_NONE; pl[k] = RB_NONE; }
        bool over = false;
        group_scan_sphere<true>(P, valid, A, gA, xyz(fa), fa.w, 0u, pg, pl, over, T);
        if (s < n_own_rows) {
            uint32_t np = 0;
#pragma unroll
            for (int k = 0; k < RB_GRP_PART; ++k) if (pg[k] != RB_NONE) { P.ce_pg[(size_t)k * P.ns + s] = pg[k]; P.ce_pl[(size_t)k * P.ns + s] = pl[k]; np++; }
            P.ce_ph[s] = valid ? ((np << 8) | (over ? RB_CE_OVER : 0u)) : 0u;
            // the first sphere of a body that finds a partner enters the body in the sphere-sphere list
            if (np && atomicAdd(&P.ss_npart[A], np) == 0u) P.ss_list[(uint32_t)atomicAdd(&P.counters[RB_C_NSS], 1ull)] = A;
        }
    } else {
        // the triangle list: reused while the sphere stays inside the skin it was staged with, else staged again -- for the sphere
        // grown by the skin if this tick's margin fits inside it (a fast mover's list would not even be valid now: per-tick radius)
        bool reuse = false, persist = false;
        uint32_t old_hdr = 0;
        float stage_r = fa.w;
        if (valid && P.ce_skin > 0.f) {
            const float4 bb = P.ce_build[s];
            const float r = P.sph_obj[s].w, margin = fa.w - r;
            const float dx = fa.x - bb.x, dy = fa.y - bb.y, dz = fa.z - bb.z;
            if (bb.w >= 0.f && sqrtf((dx * dx + dy * dy) + dz * dz) + margin <= P.ce_skin) { reuse = true; old_hdr = P.ce_hdr[s]; }
            else if (margin <= P.ce_skin) { persist = true; stage_r = r + P.ce_skin; }
        }
        uint32_t nc; bool st_slow;
        group_stage_triangles(P, valid && !reuse, xyz(fa), stage_r, s_cand, nc, st_slow, T);
        if (reuse) { nc = old_hdr & 0xffu; st_slow = (old_hdr & RB_CE_SLOW) != 0u; }
        if (persist && !st_slow && nc) {
            // a list that is going to live is pruned with the exact predicate itself, run once at the build pose with the radius grown
            // by the skin and a slack far above the predicate's rounding error (1e-5 of the world's half width): the distance to a
            // triangle is 1-Lipschitz in the sphere centre, so what is dropped cannot test positive while the list is valid
            const float prad = stage_r + fmaxf(1e-2f, 1e-5f * P.half);
            uint32_t nu = 0;
            for (uint32_t j = 0; j < nc; ++j) {
                const uint32_t t = s_cand[j][threadIdx.x];
                const float4* tp = P.tri + 3ull * t;
                const float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
                if (sphere_triangle_detect(xyz(fa), prad, xyz(a4), xyz(b4), xyz(c4))) s_cand[nu++][threadIdx.x] = t;
            }
            nc = nu;
        }
        if (s < n_own_rows) {
            if (!reuse) {
                for (uint32_t j = 0; j < nc; ++j) P.ce_tri[(size_t)j * P.ns + s] = s_cand[j][threadIdx.x];
                if (valid && P.ce_skin > 0.f) P.ce_build[s] = make_float4(fa.x, fa.y, fa.z, (persist && !st_slow) ? 0.f : -1.0f);
            }
            P.ce_hdr[s] = valid ? (nc | (st_slow ? RB_CE_SLOW : 0u)) : 0u;
            if (!valid && P.ce_skin > 0.f) P.ce_build[s].w = -1.0f;      // (not in the broadphase: the header word above no longer holds its list)
        }
    }
    collide_epilogue(P, H, T, lane);
}

// K4P: the partner scan of a compound body, one thread per BODY. The spheres of a body sit within a radius or two of each other,
// so their eight scans of the sphere grid (rb_compound_enum_kernel<1>) read nearly the same columns -- and find the seven
// siblings every time; here the body walks the union of those column ranges once and tests every foreign sphere against its own
// spheres with the same grown-radius reject, i.e. it keeps exactly the partner bodies the per-sphere scans keep. The list
// (ascending, unique) is parked in the per-sphere slots the resolution reads: entries 0-3 with sphere 0, 4-7 with sphere 1.
#define RB_BODY_PART 8
__global__ void __launch_bounds__(RB_BLOCK) rb_compound_partner_kernel(const RbParams P) {
    const uint32_t A = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    HitBuf H; H.n = 0;
    const bool owned = A < rb_n_owned(P);
    bool valid = owned && (P.flags[A] & RB_F_ACTIVE) != 0;
    uint32_t s0 = 0, nsph = 0;
    if (owned) { s0 = rb_body_s0(P, A); nsph = rb_body_ns(P, A); }
    float ulo = 3.0e38f, uhi = -3.0e38f, vlo = 3.0e38f, vhi = -3.0e38f;
    uint32_t nin = 0;
    if (valid) {
        for (uint32_t i = 0; i < nsph; ++i) {
            if (P.key[s0 + i] == RB_NONE) continue;
            const float4 fa = P.ws[s0 + i];
            const float reach = fa.w + P.rm_max, fu = rb_gu(P, fa.x, fa.z), fv = rb_gv(P, fa.x, fa.z);
            ulo = fminf(ulo, fu - reach); uhi = fmaxf(uhi, fu + reach); vlo = fminf(vlo, fv - reach); vhi = fmaxf(vhi, fv + reach);
            nin++;
        }
    }
    valid = valid && nin != 0u;
    uint32_t pg[RB_BODY_PART], pl[RB_BODY_PART];
#pragma unroll
    for (int k = 0; k < RB_BODY_PART; ++k) { pg[k] = RB_NONE; pl[k] = RB_NONE; }
    bool over = false;
    if (valid) {
        const uint32_t cx0 = rb_col(ulo, P.g_off_u, P.s_inv_w, P.g_nu), cx1 = rb_col(uhi, P.g_off_u, P.s_inv_w, P.g_nu);
        const uint32_t cz0 = rb_col(vlo, P.g_off_v, P.s_inv_w, P.g_nv), cz1 = rb_col(vhi, P.g_off_v, P.s_inv_w, P.g_nv);
        for (uint32_t cz = cz0; cz <= cz1; ++cz) {
            uint32_t i = P.col_start[cz * P.g_nu + cx0];
            const uint32_t e = P.col_start[cz * P.g_nu + cx1 + 1];      // u-adjacent columns are contiguous in the sort
            for (; i < e; ++i) {
                const uint32_t sid = P.sorted_id[i] & ~RB_TRI_BIT;
                if (sid - s0 < nsph && !P.kstride) continue;             // a sibling (CSR: the body's spheres are one id range)
                const uint32_t Bl = rb_sph_owner(P, sid);
                if (Bl == A) continue;
                const float4 o = P.sorted[i];
                uint32_t pass = 0;
                for (uint32_t k = 0; k < nsph; ++k) {
                    if (P.key[s0 + k] == RB_NONE) continue;
                    const float4 fa = P.ws[s0 + k];
                    const float dx = o.x - fa.x, dy = o.y - fa.y, dz = o.z - fa.z, rr = fa.w + o.w;
                    if (!((dx * dx + dy * dy) + dz * dz > rr * rr)) pass++;
                }
                if (!pass) continue;
                T.cand_ss += pass;
                uint32_t g = rb_gid(P, Bl), bl = Bl;
                bool dup = false;
#pragma unroll
                for (int k = 0; k < RB_BODY_PART; ++k) dup |= (pg[k] == g);
                if (dup) continue;
#pragma unroll
                for (int k = 0; k < RB_BODY_PART; ++k)
                    if (g < pg[k]) { const uint32_t tg = pg[k], tl = pl[k]; pg[k] = g; pl[k] = bl; g = tg; bl = tl; }
                if (g != RB_NONE) over = true;          // pushed off the end (or never fitted)
            }
        }
    }
    if (owned) {
        uint32_t count = 0;
#pragma unroll
        for (int k = 0; k < RB_BODY_PART; ++k) count += pg[k] != RB_NONE;
        const uint32_t cap = min((uint32_t)RB_BODY_PART, RB_GRP_PART * nsph);
        if (count > cap) { over = true; count = cap; }
        for (uint32_t i = 0; i < nsph; ++i) {
            const uint32_t np = count > RB_GRP_PART * i ? min(count - RB_GRP_PART * i, (uint32_t)RB_GRP_PART) : 0u;
            P.ce_ph[s0 + i] = (np << 8) | (over ? RB_CE_OVER : 0u);
            if (i < RB_BODY_PART / RB_GRP_PART) {
#pragma unroll
                for (uint32_t k = 0; k < RB_GRP_PART; ++k)
                    if (k < np) { P.ce_pg[(size_t)k * P.ns + s0 + i] = pg[RB_GRP_PART * i + k]; P.ce_pl[(size_t)k * P.ns + s0 + i] = pl[RB_GRP_PART * i + k]; }
            }
        }
        if (count) { P.ss_npart[A] = count; P.ss_list[(uint32_t)atomicAdd(&P.counters[RB_C_NSS], 1ull)] = A; }
    }
    collide_epilogue(P, H, T, lane);
}

// one staged triangle against sphere (so, r) of the body at its current pose: process_triangle without the prune
template <bool RESOLVE>
__device__ __forceinline__ void exact_triangle(const RbParams& P, BodyState& S, HitBuf& H, Tally& T, uint32_t sa_global, f3 so, float r, uint32_t t, bool& hit_any) {
    const float4* tp = P.tri + 3ull * t;
    const float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
    const f3 TA = xyz(a4), TB = xyz(b4), TC = xyz(c4);
    const f3 c = mk3(so.x + S.mt.x, so.y + S.mt.y, so.z + S.mt.z);
    T.tests_st++;
    if (!sphere_triangle_detect(c, r, TA, TB, TC)) return;
    hit_any = true;
    T.hits_st++;
    const f3 q = closest_point(c, TA, TB, TC);
    const f3 ov = sub3(c, q);
    const float dist = mag3(ov);
    const f3 o = mk3(ov.x / dist, ov.y / dist, ov.z / dist);
    record_hit(P, H, sa_global, RB_TRI_BIT | t, r - dist, o, r);
    if (RESOLVE) {
        const f3 normal = mk3(a4.w, b4.w, c4.w);
        const f3 delta = sub3(add3(mul3(o, r * P.pushout), q), c);
        const float nc = dot3(S.v, normal);
        const f3 rv = sub3(S.v, mul3(normal, nc));
        body_set_position(S, add3(S.p, delta));
        S.v = rv;
        S.flags |= RB_F_CONTACT;
    }
}

// PHASE 1: the bodies of the sphere-sphere list (one thread each: warps full of bodies that all have partners), state parked;
// PHASE 2: the triangle phase, flags and deferred commit of the bodies WITHOUT partners -- independent of phase 1, so the two run
//          side by side on two streams (phase 1 is a few long dependent chains, phase 2 fills the GPU);
// PHASE 3: the same for the listed bodies, from the parked state.
template <bool RESOLVE, int PHASE>
__global__ void __launch_bounds__(RB_BLOCK) rb_compound_resolve_kernel(const RbParams P) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    HitBuf H; H.n = 0;
    bool valid;
    uint32_t A = tid;
    if (PHASE == 1 || PHASE == 3) { if (blockIdx.x * blockDim.x >= (uint32_t)P.counters[RB_C_NSS]) return; valid = tid < (uint32_t)P.counters[RB_C_NSS]; if (valid) A = P.ss_list[tid]; }
    else valid = tid < rb_n_owned(P) && P.ss_npart[tid] == 0u;
    uint32_t s0 = 0, nsph = 0, flags_in = 0;
    if (valid) { flags_in = P.flags[A]; valid = (flags_in & RB_F_ACTIVE) != 0; }
    if (valid) {
        s0 = rb_body_s0(P, A); nsph = rb_body_ns(P, A);
        BodyState S;
        const f3 mt_frozen = xyz(P.mt[A]);
        const uint32_t slot = (PHASE == 3 && RESOLVE) ? P.ss_slot[A] : 0u;
        if (slot) {
            const float4 a = P.ss_pos[slot - 1u], b = P.ss_vel[slot - 1u];
            S.p = xyz(a); S.v = xyz(b); S.mt = xyz(P.ss_mt[slot - 1u]); S.pt = xyz(P.ss_pt[slot - 1u]);
            S.flags = __float_as_uint(a.w); S.changed = __float_as_uint(b.w) != 0u;
            P.ss_slot[A] = 0u;
        } else {
            S.p = xyz(P.pos[A]); S.v = xyz(P.vel[A]); S.mt = mt_frozen; S.pt = xyz(P.pt[A]);
            S.flags = flags_in; S.changed = false;
        }
        if (PHASE == 3) P.ss_npart[A] = 0u;       // (ready for the next tick's enumeration)
        S.wt = P.wt ? xyz(P.wt[A]) : mk3(0.f, 0.f, 0.f);
        const uint32_t gA = rb_gid(P, A);
        // ---------------- sphere-sphere: partner bodies in ascending id, merged over the spheres' lists ----------------
        uint32_t any_np = 0;
        if (PHASE == 1) for (uint32_t i = 0; i < nsph; ++i) any_np |= P.ce_ph[s0 + i] & 0xff00u;
        uint32_t lo = 0;
        while (any_np) {
            uint32_t best = RB_NONE, best_local = RB_NONE;
            for (uint32_t i = 0; i < nsph; ++i) {
                const uint32_t sa = s0 + i, h = P.ce_ph[sa], np = (h >> 8) & 0xffu;
                bool found = false;
                for (uint32_t k = 0; k < np; ++k) {
                    const uint32_t g = P.ce_pg[(size_t)k * P.ns + sa];
                    if (g >= lo) { if (g < best) { best = g; best_local = P.ce_pl[(size_t)k * P.ns + sa]; } found = true; break; }
                }
                if (!found && (h & RB_CE_OVER)) {       // its list is used up and there was more: the grid again, above lo
                    const float4 fs = P.ws[sa];
                    uint32_t eligible = 0;
                    scan_partners<false>(P, A, xyz(fs), fs.w, lo, best, best_local, eligible, T);
                }
            }
            if (best == RB_NONE) break;
            const uint32_t B = best_local;
            const bool lo_is_a = gA < best;
            if (RESOLVE || lo_is_a) {
                // partner simulated from its frozen pose (it toggles exactly like us on every hit)
                const uint32_t t0 = rb_body_s0(P, B), nb_sph = rb_body_ns(P, B);
                f3 bmt = xyz(P.mt[B]), bpt = xyz(P.pt[B]), bwt = mk3(0.f, 0.f, 0.f);
                if (P.wt) bwt = xyz(P.wt[B]);
                bool b_toggled = false;
                const uint32_t n_out = lo_is_a ? nsph : nb_sph, n_in = lo_is_a ? nb_sph : nsph;
                for (uint32_t io = 0; io < n_out; ++io)
                    for (uint32_t ii = 0; ii < n_in; ++ii) {
                        const uint32_t ia = lo_is_a ? io : ii, ib = lo_is_a ? ii : io;
                        const uint32_t sa = s0 + ia, sb = t0 + ib;
                        if (P.key[sa] == RB_NONE || P.key[sb] == RB_NONE) continue;
                        const float4 oa = P.sph_obj[sa], ob = P.sph_obj[sb];
                        const f3 ca = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
                        f3 cb;
                        if (!b_toggled) cb = xyz(P.ws[sb]); else cb = mk3(ob.x + bmt.x, ob.y + bmt.y, ob.z + bmt.z);
                        const f3 cn = lo_is_a ? sub3(ca, cb) : sub3(cb, ca);       // the lower body id plays "sphereA" / "modelA"
                        const float dist = mag3(cn);
                        const float radii = lo_is_a ? (oa.w + ob.w) : (ob.w + oa.w);
                        if (lo_is_a) T.tests_ss++;
                        if (!(radii - dist >= 0)) continue;
                        const f3 n1 = mk3(cn.x / dist, cn.y / dist, cn.z / dist);
                        if (lo_is_a) { T.hits_ss++; record_hit(P, H, rb_sph_gid(P, A, s0, ia), rb_sph_gid(P, B, t0, ib), radii - dist, n1, oa.w); }
                        if (RESOLVE) {
                            f3 n = n1;
                            if (!lo_is_a) { n = neg3(n); n = normalize3(n); }
                            const float sp = mag3(S.v);
                            S.v = mul3(mul3(n, sp), P.ss_damp);
                            body_set_position(S, S.pt);
                            const f3 t = mk3(bwt.x + bpt.x, bwt.y + bpt.y, bwt.z + bpt.z);
                            bpt = bmt; bmt = t; b_toggled = true;
                        }
                    }
            }
            lo = best + 1;
        }
        if (PHASE == 1 && RESOLVE) {      // parked for the triangle phase
            P.ss_pos[tid] = make_float4(S.p.x, S.p.y, S.p.z, __uint_as_float(S.flags));
            P.ss_vel[tid] = make_float4(S.v.x, S.v.y, S.v.z, __uint_as_float(S.changed ? 1u : 0u));
            P.ss_mt[tid] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f); P.ss_pt[tid] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
            P.ss_slot[A] = tid + 1u;
        }
        // ---------------- sphere-triangle: own spheres ascending, staged triangles in ascending id ----------------
        bool last_hit = false;
        for (uint32_t i = 0; PHASE >= 2 && i < nsph; ++i) {
            const uint32_t sa = s0 + i;
            bool hit_any = false;
            if (P.key[sa] != RB_NONE) {
                const uint32_t h = P.ce_hdr[sa];
                const float4 oa = P.sph_obj[sa];
                const uint32_t sid = rb_sph_gid(P, A, s0, i);
                if (h & RB_CE_SLOW) { const float4 fs = P.ws[sa]; sphere_vs_triangles<RESOLVE>(P, S, H, T, sid, xyz(oa), oa.w, xyz(fs), fs.w, hit_any); }
                else {
                    const uint32_t nc = h & 0xffu;
                    for (uint32_t j = 0; j < nc; ++j) exact_triangle<RESOLVE>(P, S, H, T, sid, xyz(oa), oa.w, P.ce_tri[(size_t)j * P.ns + sa], hit_any);
                }
            }
            if (i == nsph - 1) last_hit = hit_any;
        }
        if (RESOLVE && PHASE >= 2) {
            if ((flags_in & RB_F_CONTACT) && !last_hit) S.flags &= ~RB_F_CONTACT;       // Physics.cpp:201-207
            const float disp = mag3(sub3(S.mt, mt_frozen));
            if (nsph && disp > P.ws[s0].w - P.sph_obj[s0].w) T.overflow++;
            if (S.changed || S.flags != flags_in) {
                // partners read our frozen pose concurrently: commit after the kernel (K5)
                const uint32_t k = (uint32_t)atomicAdd(&P.counters[RB_C_NCHANGED], 1ull);
                P.ch_body[k] = A;
                P.ch_pos[k] = make_float4(S.p.x, S.p.y, S.p.z, 1.f);
                P.ch_vel[k] = make_float4(S.v.x, S.v.y, S.v.z, 0.f);
                P.ch_mt[k] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f);
                P.ch_pt[k] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
                P.ch_flags[k] = S.flags;
            }
        }
    }
    collide_epilogue(P, H, T, lane);
}

// PHASE 1 with a group of lanes per listed body: the sphere pairs of a partner are dealt over the group and resolved in
// canonical order by speculation (group_sphere_sphere), the state parked for the triangle phase by the group's first lane.
template <bool RESOLVE>
__global__ void __launch_bounds__(RB_BLOCK) rb_compound_ss_group_kernel(const RbParams P) {
    const uint32_t lg = P.grp_lg, L = 1u << lg;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, sub = lane & (L - 1);
    const uint32_t e = tid >> lg;                   // entry of the sphere-sphere list (uniform over the group)
    if (((blockIdx.x * blockDim.x) >> lg) >= (uint32_t)P.counters[RB_C_NSS]) return;      // (the launch is sized for every body: whole blocks beyond the list leave at once)
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    HitBuf H; H.n = 0;
    bool valid = e < (uint32_t)P.counters[RB_C_NSS];
    uint32_t A = 0, s0 = 0, nsph = 0, flags_in = 0, gA = 0;
    if (valid) { A = P.ss_list[e]; flags_in = P.flags[A]; valid = (flags_in & RB_F_ACTIVE) != 0; }
    BodyState S; S.changed = false; S.flags = flags_in;
    S.p = S.v = S.mt = S.pt = S.wt = mk3(0.f, 0.f, 0.f);
    if (valid) {
        s0 = rb_body_s0(P, A); nsph = rb_body_ns(P, A);
        S.p = xyz(P.pos[A]); S.v = xyz(P.vel[A]); S.mt = xyz(P.mt[A]); S.pt = xyz(P.pt[A]);
        if (P.wt) S.wt = xyz(P.wt[A]);
        gA = rb_gid(P, A);
    }
    const uint32_t nchunk_w = __reduce_max_sync(0xffffffffu, (nsph + L - 1) >> lg);
    group_sphere_sphere<RESOLVE, true>(P, valid, A, gA, s0, nsph, S, H, T, lg, nchunk_w);
    if (RESOLVE && valid && sub == 0) {
        P.ss_pos[e] = make_float4(S.p.x, S.p.y, S.p.z, __uint_as_float(S.flags));
        P.ss_vel[e] = make_float4(S.v.x, S.v.y, S.v.z, __uint_as_float(S.changed ? 1u : 0u));
        P.ss_mt[e] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f); P.ss_pt[e] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
        P.ss_slot[A] = e + 1u;
    }
    collide_epilogue(P, H, T, lane);
}
#endif // !RB_CHAIN_TU

// ------------------------------------------------------------------------------------------
// K4L: the list step -- one thread per body ROW, every tick of the persistent-list regime.
// Everything a row needs is addressed by its row number: state (pos | radius, vel), the list header, up to four
// partner rows, and its slice of the block's triangle tile, which one elected thread stages in shared memory with a
// single TMA bulk copy (cp.async.bulk + mbarrier) while the block is still busy with the sphere-sphere phase. The
// exact tests then read shared memory: no per-test 48-B gather, no dependent load behind the list.
// Same canonical order and the same FP32 expression trees as rb_collide_single_kernel: bit-identical results.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// append n contact records of the calling lanes with one atomic (any converged subset of a warp may call this)
__device__ __forceinline__ void append_contact(const RbParams& P, uint32_t a, uint32_t b, float depth, f3 n, float r) {
    const uint32_t active = __activemask();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t leader = __ffs(active) - 1, rank = __popc(active & ((1u << lane) - 1u));
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(P.ncontact, (unsigned long long)__popc(active));
    base = __shfl_sync(active, base, leader);
    const unsigned long long slot = base + rank;
    if (slot < P.contact_cap) store_contact(P, slot, a, b, depth, n.x, n.y, n.z, r);
    else atomicAdd(&P.counters[RB_C_DROPPED], 1ull);
}

#ifndef RB_LIST_MINBLOCKS
#define RB_LIST_MINBLOCKS (1024 / RB_LS_BLOCK)
#endif
#ifndef RB_LS_PREFETCH
#define RB_LS_PREFETCH 1
#endif
// dynamic shared memory of the list step: the tile (vertex records + ids), then one owner word per tile entry and the work
// list of the later test rounds
#define RB_LS_TILE_BYTES ((3u * RB_TILE_CAP + RB_TILE_CAP / 4u) * 16u)
#define RB_LS_SMEM_BYTES (RB_LS_TILE_BYTES + 2u * RB_TILE_CAP * 2u)
// phase 0: every row (single worlds). Slab worlds split the step around the halo exchange: phase 1 = the blocks no ghost can
// reach, launched before the wait for the neighbours (skipped altogether when K1 has asked for a rebuild: rows are about to
// move); phase 2 = all the other blocks, after the ghosts have been linked in, plus the end-of-tick book-keeping.
__device__ __forceinline__ void list_tick_bookkeeping(const RbParams& P);
__global__ void __launch_bounds__(RB_LS_BLOCK, RB_LIST_MINBLOCKS) rb_list_step_kernel(const RbParams P, const int phase) {
    if (phase == 1) { if (P.vl[RB_VL_REBUILD] || P.blk_bnd[blockIdx.x] == P.tick) return; }
    else if (phase == 2 && P.blk_done[blockIdx.x] == P.tick) { if (blockIdx.x == 0 && threadIdx.x == 0) list_tick_bookkeeping(P); return; }
    extern __shared__ __align__(16) uint4 s_tile[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ __align__(16) float4 s_pose[RB_LS_BLOCK];      // current sphere of each row of the block (w < 0: the row does not step here)
    __shared__ uint32_t s_mask[RB_LS_BLOCK];                  // hits of the round, one bit per entry of the row's list
    __shared__ uint32_t s_nwork[2];
    uint16_t* s_own = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_tile) + RB_LS_TILE_BYTES);      // per tile entry: owner thread | index in its list << 8
    uint16_t* s_work = s_own + RB_TILE_CAP;                // tile entries still to be tested (rounds after the first)
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_owned = rb_n_owned(P);
    const uint32_t A = tid;
    uint32_t keyA = RB_NONE, hdr = 0, fl = 0;
    float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = p4;
    if (tid < n_owned) { keyA = P.key[tid]; hdr = P.e_hdr[tid]; fl = P.flags[tid]; p4 = P.pos[tid]; v4 = P.vel[tid]; }     // all in flight together
    // ---- the block's triangle tile: one bulk copy, in flight while the partners are dealt with ----
    const uint2 bt = P.blk_tile[blockIdx.x];
    const uint32_t tile_n = bt.y;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        if (tile_n) {
            const uint32_t bytes = (3u * tile_n + ((tile_n + 3u) >> 2)) * 16u;
            mbar_expect_tx(&s_bar, bytes);
            bulk_g2s(s_tile, P.tile_pool + bt.x, bytes, &s_bar);
        }
        s_nwork[0] = 0u; s_nwork[1] = 0u;
    }
    __syncthreads();        // the barrier word is initialised before anybody polls it (here, where the warps of the block still run together)
    Tally T = { 0, 0, 0, 0, 0, 0, 0 };
    // rows outside the broadphase have nothing to do; hot rows step in their own launch (rb_collide_single_kernel<3>)
    const bool valid = keyA != RB_NONE && !(keyA & RB_HOT_BIT);
    const bool had_contact = valid && (fl & RB_F_CONTACT) != 0;
    BodyState S; S.changed = false; S.flags = fl;
    S.p = xyz(p4); S.v = xyz(v4); S.pt = S.wt = mk3(0.f, 0.f, 0.f);
    float4 oa = make_float4(0.f, 0.f, 0.f, p4.w);            // centred sphere: the world radius rides in pos.w
    if (valid) {
        if (!P.obj_zero) oa = P.sph_obj[A];
        if (P.wt) S.wt = xyz(P.wt[A]);
    }
    if (P.lazy_mt) S.mt = mk3(0.f + S.p.x, 0.f + S.p.y, 0.f + S.p.z);      // Entity::_mvp after K1: translation(p) * W with W.t = 0
    else S.mt = valid ? xyz(P.mt[A]) : mk3(0.f, 0.f, 0.f);
    const uint32_t gA = valid ? rb_gid(P, A) : 0u;
    // the frozen sphere K1 published for this row (same expressions, same bits): centre and grown radius
    const f3 fc = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
    const float frm = oa.w + fminf(P.margin_cap, fmaxf(2.0f * mag3(S.v) * P.dt + 0.01f * oa.w, 1e-3f));
    const f3 mt_frozen = S.mt;

    // ---------------- partners: from the list, plus whoever is in the hot-body table ----------------
    // The list's partners (ascending global id, up to RB_SSLIST_MAX of them, slot-major by row) are streamed one at a time;
    // the hot bodies found around the row are kept sorted in registers and merged in on the way, so the pairs run in
    // ascending partner id whichever source they come from (a partner that is listed AND hot shows up in both: once).
    const uint32_t nl = valid ? (hdr & 0xffu) : 0u;
    const uint32_t nc = valid ? ((hdr >> 8) & 0xffu) : 0u;
    const uint32_t toff = hdr >> 16;
    uint32_t lg = RB_NONE, ll = RB_NONE, li = 0;              // the list partner in hand (global id, row) and its slot
    if (nl) { lg = P.e_pg[A]; ll = P.e_pl[A]; }
#if RB_LS_PREFETCH
    if (nl) rb_prefetch(P.pt + A);                             // (the row's own previous pose: only a sphere-sphere hit reads it)
    if (nl > 1u) { rb_prefetch(P.e_pg + (size_t)P.ns + A); rb_prefetch(P.e_pl + (size_t)P.ns + A); }
#endif
    uint32_t hg[RB_SSCAND_MAX], hl[RB_SSCAND_MAX], nhp = 0;   // hot partners, ascending
#pragma unroll
    for (int k = 0; k < RB_SSCAND_MAX; ++k) { hg[k] = RB_NONE; hl[k] = RB_NONE; }
    {
        // hot bodies (and, in slab worlds, the neighbours' ghosts) are in nobody's list: everybody looks them up in the
        // coarse cell table (<= 2 x 2 cells). Entries of earlier ticks read as empty cells.
        uint32_t hcell[4] = { 0, 0, 0, 0 }, hn = 0;
        if (valid) {
            const float reach = frm + P.rm_max;
            const uint32_t hx0 = rb_col(fc.x - reach, P.half, P.hc_inv_w, P.hcn), hx1 = rb_col(fc.x + reach, P.half, P.hc_inv_w, P.hcn);
            const uint32_t hz0 = rb_col(fc.z - reach, P.half, P.hc_inv_w, P.hcn), hz1 = rb_col(fc.z + reach, P.half, P.hc_inv_w, P.hcn);
            hcell[0] = hz0 * P.hcn + hx0; hn = 1;
            if (hx1 != hx0) hcell[hn++] = hz0 * P.hcn + hx1;
            if (hz1 != hz0) { hcell[hn++] = hz1 * P.hcn + hx0; if (hx1 != hx0) hcell[hn++] = hz1 * P.hcn + hx1; }
        }
        uint32_t hh[4];
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) hh[q] = (q < hn) ? hot_cell_head(P, hcell[q]) : 0u;      // the (<= 4) heads, all requested up front
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            uint32_t h = hh[q];
            while (true) {
                if (!__any_sync(0xffffffffu, h != 0u)) break;
                if (h) {
                    const uint32_t Bl = P.hot_row[h - 1u];
                    h = P.hot_next[h - 1u];
                    if (Bl != A && Bl != RB_NONE) {
                        const float4 o = P.ws[Bl];
                        float dx = o.x - fc.x, dy = o.y - fc.y, dz = o.z - fc.z, rr = frm + o.w;
                        if (o.w >= 0.f && !((dx * dx + dy * dy) + dz * dz > rr * rr)) {
                            uint32_t g = rb_gid(P, Bl), bl = Bl;
                            bool dup = false;           // (a body straddling two cells is linked once, but be safe)
#pragma unroll
                            for (int k = 0; k < RB_SSCAND_MAX; ++k) dup |= (hg[k] == g);
                            if (!dup) {
                                T.cand_ss++; nhp++;
#pragma unroll
                                for (int k = 0; k < RB_SSCAND_MAX; ++k)
                                    if (g < hg[k]) { uint32_t tg = hg[k], tl = hl[k]; hg[k] = g; hl[k] = bl; g = tg; bl = tl; }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }

    // ---------------- sphere-sphere in ascending partner id ----------------
    // GeometryMath::sphereSphereDetection (math/src/GeometryMath.cpp:211-230): hit iff (rA + rB) - sqrt(d2) >= 0, i.e.
    // sqrt_rn(d2) <= radii. A pair with d2 > radii^2 * (1 + 2^-20) cannot pass (sqrt is monotone, the bound is > 3 ulp of
    // radii away), so it is rejected without the square root; everything closer runs the reference expressions.
    bool pt_loaded = false;
    const bool crowded = nhp > RB_SSCAND_MAX;          // more hot neighbours than the registers hold: this row rescans (below)
    if (crowded) { lg = RB_NONE; hg[0] = RB_NONE; }
    while (true) {
        const bool have = lg != RB_NONE || hg[0] != RB_NONE;
        if (!__any_sync(0xffffffffu, have)) break;
        if (have) {
            uint32_t gB, Bl;
            const bool from_hot = hg[0] < lg;
            const bool both = hg[0] == lg;              // listed and hot: one body
            if (from_hot) { gB = hg[0]; Bl = hl[0]; } else { gB = lg; Bl = ll; }
            if (from_hot || both) {
#pragma unroll
                for (int k = 0; k + 1 < RB_SSCAND_MAX; ++k) { hg[k] = hg[k + 1]; hl[k] = hl[k + 1]; }
                hg[RB_SSCAND_MAX - 1] = RB_NONE;
            }
            if (!from_hot) {                            // the next list partner comes in while this pair is dealt with
                ++li;
                if (li < nl) { lg = P.e_pg[(size_t)li * P.ns + A]; ll = P.e_pl[(size_t)li * P.ns + A]; } else lg = RB_NONE;
            }
            const bool lo_is_a = gA < gB;
            const float4 wb = P.ws[Bl];                        // partner at its frozen pose
            if (wb.w >= 0.f) {                                 // (w < 0: it has left the broadphase since the list was built, Q3 / Q7)
                const float rb = P.pos[Bl].w;                  // its exact radius (pos.w never changes)
                const f3 ca = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
                const f3 cb = xyz(wb);
                const f3 cn = lo_is_a ? sub3(ca, cb) : sub3(cb, ca);          // lower id plays sphereA / modelA
                const float radii = lo_is_a ? (oa.w + rb) : (rb + oa.w);
                const float d2 = (cn.x * cn.x + cn.y * cn.y) + cn.z * cn.z;
                if (lo_is_a) T.tests_ss++;
                if (!(d2 > (radii * radii) * 1.000001f)) {
                    const float dist = sqrtf(d2);
                    if (radii - dist >= 0) {
                        const f3 n1 = mk3(cn.x / dist, cn.y / dist, cn.z / dist);
                        if (lo_is_a) { T.hits_ss++; append_contact(P, gA, gB, radii - dist, n1, oa.w); }
                        f3 n = n1;
                        if (!lo_is_a) { n = neg3(n); n = normalize3(n); }
                        const float sp = mag3(S.v);
                        S.v = mul3(mul3(n, sp), P.ss_damp);
                        if (!pt_loaded && !S.changed) S.pt = xyz(P.pt[A]);      // the previous pose K1 left (only a snap-back reads it)
                        pt_loaded = true;
                        body_set_position(S, S.pt);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (crowded) {            // every partner by rescanning (the sort of the last rebuild + the hot-body table), ascending
        HitBuf H; H.n = 0;
        if (!pt_loaded && !S.changed) { S.pt = xyz(P.pt[A]); pt_loaded = true; }
        uint32_t lo = 0;
        const float4 oa_s = make_float4(oa.x, oa.y, oa.z, oa.w);
        while (true) {
            uint32_t best = RB_NONE, best_local = RB_NONE, eligible = 0;
            scan_partners<true>(P, A, fc, frm, lo, best, best_local, eligible, T);
            scan_hot_partners(P, A, fc, frm, lo, best, best_local, eligible);
            if (best == RB_NONE) break;
            ss_pair_single<true>(P, S, H, T, A, gA, oa_s, best_local, best);
            lo = best + 1;
            if (eligible <= 1) break;
        }
        for (uint32_t i = 0; i < min(H.n, (uint32_t)RB_MAXLOCAL); ++i) append_contact(P, H.a[i], H.b[i], H.depth[i], mk3(H.nx[i], H.ny[i], H.nz[i]), H.r[i]);
    }

    // ---------------- sphere-triangle: exact tests out of the staged tile, pooled over the BLOCK ----------------
    // Speculative rounds (see rb_collide_single_kernel P4): until a sphere is hit its pose does not change, so all of its
    // remaining triangles can be tested at once. Round 1 tests every entry of the tile, one entry per thread whoever owns
    // it (the owner's sphere comes from shared memory); each owner then takes its FIRST hit in canonical order, resolves it
    // and queues the entries behind that hit for the next round -- exactly the sequential test / resolve sequence, with
    // full warps in every round.
    const uint32_t nc_list = tid < n_owned ? ((hdr >> 8) & 0xffu) : 0u;      // (a row that is hot this tick still has its slice of the tile)
    for (uint32_t j = 0; j < nc_list; ++j) s_own[toff + j] = (uint16_t)(threadIdx.x | (j << 8));
    {
        const f3 c0 = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);       // Sphere::getPosition at the current pose
        s_pose[threadIdx.x] = make_float4(c0.x, c0.y, c0.z, valid ? oa.w : -1.0f);
    }
    s_mask[threadIdx.x] = 0u;
    if (tile_n) mbar_wait(&s_bar, 0);
    __syncthreads();
    const uint32_t* tile_ids = reinterpret_cast<const uint32_t*>(s_tile + 3u * tile_n);
    bool hit_any = false;
    {
        uint32_t j0 = 0, nwork = tile_n, round = 0;
        while (true) {
            for (uint32_t i = threadIdx.x; i < nwork; i += RB_LS_BLOCK) {
                const uint32_t e = round ? (uint32_t)s_work[i] : i;
                const uint32_t ok = s_own[e];
                const float4 ps = s_pose[ok & 0xffu];
                if (ps.w >= 0.f) {
                    const uint4* tp = s_tile + 3u * e;
                    const uint4 a4 = tp[0], b4 = tp[1], c4 = tp[2];
                    if (sphere_triangle_detect(mk3(ps.x, ps.y, ps.z), ps.w,
                                               mk3(__uint_as_float(a4.x), __uint_as_float(a4.y), __uint_as_float(a4.z)),
                                               mk3(__uint_as_float(b4.x), __uint_as_float(b4.y), __uint_as_float(b4.z)),
                                               mk3(__uint_as_float(c4.x), __uint_as_float(c4.y), __uint_as_float(c4.z))))
                        atomicOr(&s_mask[ok & 0xffu], 1u << (ok >> 8));
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_nwork[(round + 1u) & 1u] = 0u;      // the counter of the round after next (its last value has been read by everybody)
            const uint32_t rem = j0 < nc ? nc - j0 : 0u;
            if (rem) {
                const uint32_t m = s_mask[threadIdx.x];                 // bit k: entry k of this row's list tested positive (only entries >= j0 were tested)
                s_mask[threadIdx.x] = 0u;
                // sequential accounting: the tests up to and including the first hit are the ones the one-at-a-time order runs
                const uint32_t first = m ? (uint32_t)__ffs(m) - 1u : nc;
                T.tests_st += m ? first - j0 + 1u : rem;
                if (m) {
                    const uint32_t ent = toff + first;
                    const uint4* tp = s_tile + 3u * ent;
                    const uint4 a4 = tp[0], b4 = tp[1], c4 = tp[2];
                    const f3 TA = mk3(__uint_as_float(a4.x), __uint_as_float(a4.y), __uint_as_float(a4.z));
                    const f3 TB = mk3(__uint_as_float(b4.x), __uint_as_float(b4.y), __uint_as_float(b4.z));
                    const f3 TC = mk3(__uint_as_float(c4.x), __uint_as_float(c4.y), __uint_as_float(c4.z));
                    const f3 c = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
                    hit_any = true;
                    T.hits_st++;
                    const f3 q = closest_point(c, TA, TB, TC);
                    const f3 ov = sub3(c, q);
                    const float dist = mag3(ov);
                    const f3 o = mk3(ov.x / dist, ov.y / dist, ov.z / dist);
                    // the contact record is parked in the entry's own slot of the tile (its vertices are in registers and
                    // nobody tests an entry in front of a hit again) and appended once per warp at the end: no global atomic
                    // between two block barriers
                    append_contact(P, gA, RB_TRI_BIT | tile_ids[ent], oa.w - dist, o, oa.w);
                    // GeometryMath::sphereTriangleResolution (math/src/GeometryMath.cpp:489-522); the triangle normal
                    // normalize((C-A)x(B-A)) is static: pre-computed at build (w lanes)
                    const f3 normal = mk3(__uint_as_float(a4.w), __uint_as_float(b4.w), __uint_as_float(c4.w));
                    const f3 delta = sub3(add3(mul3(o, oa.w * P.pushout), q), c);
                    const float ncomp = dot3(S.v, normal);
                    const f3 rv = sub3(S.v, mul3(normal, ncomp));
                    body_set_position(S, add3(S.p, delta));
                    S.v = rv;
                    S.flags |= RB_F_CONTACT;
                    j0 = first + 1u;
                    if (j0 < nc) {      // the triangles behind the hit are tested again, at the new pose
                        const f3 c1 = mk3(oa.x + S.mt.x, oa.y + S.mt.y, oa.z + S.mt.z);
                        s_pose[threadIdx.x] = make_float4(c1.x, c1.y, c1.z, oa.w);
                        const uint32_t w0 = atomicAdd(&s_nwork[round & 1u], nc - j0);
                        for (uint32_t k = j0; k < nc; ++k) s_work[w0 + k - j0] = (uint16_t)(toff + k);
                    }
                } else j0 = nc;
            }
            __syncthreads();
            nwork = s_nwork[round & 1u];
            if (!nwork) break;
            ++round;
        }
    }
    // ---------------- flags, write-back ----------------
    if (valid) {
        if ((fl & RB_F_CONTACT) && !hit_any) S.flags &= ~RB_F_CONTACT;     // Physics.cpp:201-207
        if (S.changed) {
            const float disp = mag3(sub3(S.mt, mt_frozen));
            if (disp > frm - oa.w) T.overflow++;
            P.pos[A] = make_float4(S.p.x, S.p.y, S.p.z, p4.w);
            P.vel[A] = make_float4(S.v.x, S.v.y, S.v.z, v4.w);
            if (!P.lazy_mt) P.mt[A] = make_float4(S.mt.x, S.mt.y, S.mt.z, 0.f);
            P.pt[A] = make_float4(S.pt.x, S.pt.y, S.pt.z, 0.f);
        }
        if (S.flags != fl) P.flags[A] = S.flags;
    }
    (void)had_contact;
    if (P.pos_out && tid < n_owned && !(keyA != RB_NONE && (keyA & RB_HOT_BIT))) {
        // the host asked for this tick's positions in id order: every row is written exactly once, by the launch that
        // owns it this tick (hot rows by rb_collide_single_kernel<3>, everything else here)
        float* po = P.pos_out + (size_t)P.pos_out_stride * rb_gid(P, A);
        po[0] = S.p.x; po[1] = S.p.y; po[2] = S.p.z;
        if (P.pos_out_stride == 4) po[3] = 1.0f;
    }
    {   // cumulative counters: two per word (a lane's count stays far below 2^16 / 32), warp reduce, one atomic per warp and
        // non-zero counter on a striped copy
        const uint32_t packed[3] = { T.tests_st | (T.hits_st << 16), T.tests_ss | (T.hits_ss << 16), T.cand_ss | (T.overflow << 16) };
        const int map[3][2] = { { RB_C_TESTS_ST, RB_C_HITS_ST }, { RB_C_TESTS_SS, RB_C_HITS_SS }, { RB_C_CAND_SS, RB_C_OVERFLOW } };
        const uint32_t stripe = ((blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5)) & (RB_STRIPES - 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!__any_sync(0xffffffffu, packed[k] != 0u)) continue;
            const uint32_t sum = __reduce_add_sync(0xffffffffu, packed[k]);
            if (lane == 0) {
                if (sum & 0xffffu) atomicAdd(&P.stripes[stripe * RB_C_COUNT + map[k][0]], (unsigned long long)(sum & 0xffffu));
                if (sum >> 16) atomicAdd(&P.stripes[stripe * RB_C_COUNT + map[k][1]], (unsigned long long)(sum >> 16));
            }
        }
    }
    if (phase == 1) { if (threadIdx.x == 0) P.blk_done[blockIdx.x] = P.tick; }
    else if (tid == 0) list_tick_bookkeeping(P);
}
// end of a persistent-list tick (one thread): the hot counter is final -- K1 and the enumeration are done
__device__ __forceinline__ void list_tick_bookkeeping(const RbParams& P) {        // end of a persistent-list tick: book-keeping (the hot counter is final: K1 and the enumeration are done)
        const uint32_t f = P.vl[RB_VL_REBUILD];
        if (f) {
            const uint32_t why = P.vl[RB_VL_WHY];
            P.vl[RB_VL_WHY_LAST] = why; P.vl[RB_VL_WHY] = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) if (why & (1u << b)) P.vl[RB_VL_NWHY + b]++;      // rebuilds by reason (several may apply)
        }
        P.vl[RB_VL_STORM] = f;
        const uint32_t nr = P.vl[RB_VL_NREBUILD] + (f ? 1u : 0u), nt = P.vl[RB_VL_NTICK] + 1u;
        P.vl[RB_VL_NREBUILD] = nr; P.vl[RB_VL_NTICK] = nt;
        const uint32_t nh = *P.nhot;
        P.vl[RB_VL_NHOT_LAST] = nh;
        // the useful life of this generation of lists: the first age at which more bodies step without their list than a
        // rebuild is worth (the host schedules the rebuilds of single worlds from it)
        const uint32_t age = f ? 0u : P.vl[RB_VL_AGE] + 1u;
        uint32_t gen = P.vl[RB_VL_GEN], expiry = P.vl[RB_VL_EXPIRY];
        if (f) {
            if (P.vl_host) { P.vl_host[8] = P.vl[RB_VL_AGE]; P.vl_host[9] = expiry; }      // how the previous generation ended: ticks served, expired or not
            gen++; expiry = 0u; P.vl[RB_VL_FLOOR] = nh;
        }
        else if (!expiry && nh > P.vl[RB_VL_FLOOR] + P.hot_soft) expiry = age;
        // a flood: after ONE tick more bodies have outrun their skin than lists can carry (free fall) -- the host sits such phases out
        if (age == 1u && nh > P.vl[RB_VL_FLOOR] + P.hot_flood && P.vl_host) P.vl_host[10] = gen;
        P.vl[RB_VL_AGE] = age; P.vl[RB_VL_GEN] = gen; P.vl[RB_VL_EXPIRY] = expiry;
        if (P.vl_host) {
            P.vl_host[0] = nr; P.vl_host[1] = nt; P.vl_host[3] = nh;
            P.vl_host[5] = expiry; P.vl_host[6] = gen; P.vl_host[7] = age;
            P.vl_host[4] = nt;        // (last: the host takes a report whose tick word it reads unchanged before and after)
        }
}

// do_bin: 0 integrate only, 1 integrate + bin (per-tick enumeration), 2 list tick, 3 list tick with a rebuild forced by the host
#ifndef RB_K1_MINBLOCKS
#define RB_K1_MINBLOCKS 6      // 40 registers: measured 1.5 us faster than the unconstrained 48
#endif
__global__ void __launch_bounds__(RB_BLOCK, RB_K1_MINBLOCKS) rb_integrate_bin_kernel(RbParams P, int do_integrate, int do_bin) {
    rb_integrate_bin_body(P, do_integrate, do_bin);
}
// K1s: compound bodies of a single world -- the second half of K1 (world sphere, candidate margin, sphere-grid key and in-column
// rank), one thread per SPHERE instead of a loop over the body's spheres in the body's thread (C4: K1 90 -> 12 + 25 us).
// Same expressions as the loop of rb_integrate_bin_body.
__global__ void __launch_bounds__(RB_BLOCK) rb_bin_spheres_kernel(RbParams P, int do_bin) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.ns) return;
    const uint32_t b = P.sph_body[s];
    const uint32_t f = P.flags[b];
    const f3 m = xyz(P.mt[b]);
    const float speed = mag3(xyz(P.vel[b]));
    const float4 o = P.sph_obj[s];
    const f3 c = mk3(o.x + m.x, o.y + m.y, o.z + m.z);
    const float r = o.w;
    const float margin = fminf(P.margin_cap, fmaxf(2.0f * speed * P.dt + 0.01f * r, 1e-3f));
    const bool in_grid = (f & RB_F_ACTIVE) && sphere_in_world(c, r, P.half);
    P.ws[s] = make_float4(c.x, c.y, c.z, in_grid ? r + margin : -1.0f);
    uint32_t k = RB_NONE;
    if (in_grid) {
        if (do_bin) { k = rb_sphere_col(P, c.x, c.z); P.rank[s] = atomicAdd(&P.col_count[k], 1u); }
        else k = 0;
    }
    P.key[s] = k;
}
#ifdef RB_CHAIN_TU
// The gate of a list tick (one thread, right behind K1): if K1 raised the rebuild flag it tail-launches
// bin -> scan -> scatter -> enumerate (CUDA dynamic parallelism: tail launches run after this grid and before the next
// kernel of the stream, in order). On the other ticks -- nearly all of them -- the rebuild chain costs this one
// 2-us launch instead of four full-grid launches that return at once.
__global__ void rb_vl_gate_kernel(RbParams P) {
    if (threadIdx.x || blockIdx.x || !P.vl[RB_VL_REBUILD]) return;
    const uint32_t g = (P.ns + RB_BLOCK - 1) / RB_BLOCK, gl = (P.ns + RB_LS_BLOCK - 1) / RB_LS_BLOCK, tiles = (P.ncols + RB_SCAN_TILE - 1) / RB_SCAN_TILE;
    rb_vl_bin_kernel<<<g, RB_BLOCK, 0, cudaStreamTailLaunch>>>(P);
    rb_scan_lookback_kernel<<<tiles, RB_BLOCK, 0, cudaStreamTailLaunch>>>(P.col_count, P.ncols, P.col_start, P.scan_status, P.scan_ticket, P.vl + RB_VL_REBUILD, 1);
    rb_scatter_kernel<<<g, RB_BLOCK, 0, cudaStreamTailLaunch>>>(P, 1, P.scan_status, P.scan_ticket, tiles);
    rb_collide_single_kernel<true, 1><<<gl, RB_LS_BLOCK, 0, cudaStreamTailLaunch>>>(P, 0);
}
#endif

// K5: deferred commit for compound bodies
__global__ void __launch_bounds__(RB_BLOCK) rb_commit_kernel(RbParams P) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = (uint32_t)P.counters[RB_C_NCHANGED];
    if (k >= n) return;
    uint32_t A = P.ch_body[k];
    P.pos[A] = P.ch_pos[k]; P.vel[A] = P.ch_vel[k]; P.mt[A] = P.ch_mt[k]; P.pt[A] = P.ch_pt[k]; P.flags[A] = P.ch_flags[k];
}

// StateVector::setForce / setTorque gating (math/src/StateVector.cpp:147-167). `inv` maps a body id to its
// current row (NULL = identity): worlds re-order their rows spatially from time to time.
__global__ void __launch_bounds__(RB_BLOCK) rb_apply_force_kernel(const uint32_t* __restrict__ flags, const float4* __restrict__ in,
                                                                 float4* __restrict__ out, const uint32_t* __restrict__ inv,
                                                                 uint32_t first, uint32_t count, int keep_w) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t row = inv ? inv[first + i] : first + i;
    uint32_t f = flags[row];
    float4 v = in[i];
    float w = keep_w ? out[row].w : 0.f;
    if ((f & RB_F_CONTACT) || !(f & RB_F_GRAVITY)) out[row] = make_float4(v.x, v.y, v.z, w);
    else out[row] = make_float4(0.f, 0.f, 0.f, w);
}
__global__ void __launch_bounds__(RB_BLOCK) rb_set_flags_kernel(uint32_t* flags, const uint32_t* __restrict__ values, const uint32_t* __restrict__ inv,
                                                               uint32_t first, uint32_t count, uint32_t mask) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t row = inv ? inv[first + i] : first + i;
    flags[row] = (flags[row] & ~mask) | (values[i] & mask);
}
__global__ void __launch_bounds__(RB_BLOCK) rb_fill_f4_kernel(float4* p, uint32_t n, float4 v) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// rows -> id order (readback) and id order -> rows (setters)
// w_one: the w lane reads 1 (positions of single-sphere bodies keep their world radius there; the interface says 1)
__global__ void __launch_bounds__(RB_BLOCK) rb_rows_to_ids_f4_kernel(const float4* __restrict__ rows, const uint32_t* __restrict__ gid, float4* __restrict__ out, uint32_t n, int w_one) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = rows[i];
    if (w_one) v.w = 1.0f;
    out[gid ? gid[i] : i] = v;
}
// packed xyz variants of the host round trip (12 B per body each way instead of 16: the round trip is PCIe-bound)
__global__ void __launch_bounds__(RB_BLOCK) rb_apply_force3_kernel(const uint32_t* __restrict__ flags, const float* __restrict__ in3,
                                                                  float4* __restrict__ out, const uint32_t* __restrict__ inv, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t row = inv ? inv[i] : i;
    uint32_t f = flags[row];
    float w = out[row].w;
    if ((f & RB_F_CONTACT) || !(f & RB_F_GRAVITY)) out[row] = make_float4(in3[3ull * i], in3[3ull * i + 1], in3[3ull * i + 2], w);
    else out[row] = make_float4(0.f, 0.f, 0.f, w);
}
__global__ void __launch_bounds__(RB_BLOCK) rb_rows_to_ids_f3_kernel(const float4* __restrict__ rows, const uint32_t* __restrict__ gid, float* __restrict__ out3, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = rows[i];
    const size_t o = 3ull * (gid ? gid[i] : i);
    out3[o] = v.x; out3[o + 1] = v.y; out3[o + 2] = v.z;
}
__global__ void __launch_bounds__(RB_BLOCK) rb_rows_to_ids_u32_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ gid, uint32_t* __restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[gid[i]] = rows[i];
}
// keep_w: the w lane of the destination survives (single-sphere bodies keep their world radius in pos.w)
__global__ void __launch_bounds__(RB_BLOCK) rb_ids_to_rows_f4_kernel(const float4* __restrict__ in, const uint32_t* __restrict__ inv, float4* __restrict__ rows, uint32_t first, uint32_t count, int keep_w) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t row = inv ? inv[first + i] : first + i;
    float4 v = in[i];
    if (keep_w) v.w = rows[row].w;
    rows[row] = v;
}
// Entity::setPosition (model/src/Entity.cpp:373-391) for one body: total = translation(p) * W has the translation W.t + p;
// linear position := that translation (Q1: W.t is added again), _prevMVP := _mvp, _mvp := total. The spheres follow at the
// next kinematics update, like _worldSpaceGeometry (their transform is recomputed from the model translation every tick).
__global__ void rb_entity_set_position_kernel(float4* pos, float4* mt, float4* pt, const float4* wt, const uint32_t* inv, uint32_t body, float x, float y, float z) {
    if (threadIdx.x || blockIdx.x) return;
    const uint32_t row = inv ? inv[body] : body;
    const float4 w = wt ? wt[row] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float tx = w.x + x, ty = w.y + y, tz = w.z + z;
    pos[row] = make_float4(tx, ty, tz, pos[row].w);
    pt[row] = mt[row];
    mt[row] = make_float4(tx, ty, tz, 0.f);
}
// Entity::_mvp's translation for worlds that keep it implicit on list ticks (W.t = 0): after any complete tick it is
// 0 + p -- K1 leaves translation(p) * W, a resolution leaves p = mt = the same sum (rb_collide_single_kernel P5)
__global__ void __launch_bounds__(RB_BLOCK) rb_mt_refresh_kernel(const float4* __restrict__ pos, float4* __restrict__ mt, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    mt[i] = make_float4(0.f + p.x, 0.f + p.y, 0.f + p.z, 0.f);
}

// K8: model-matrix export. total = translation(p) * W keeps W's 3x3 and carries the translation the step
// maintains (mt / pt), so the renderer-facing 4x4 row-major matrices of Entity::_mvp / _prevMVP
// (model/src/Entity.cpp:134-140, 388-390) are composed here, in body-id order.
__global__ void __launch_bounds__(RB_BLOCK) rb_model_matrix_kernel(const float4* __restrict__ w3, const uint32_t* __restrict__ gid,
                                                                  const float4* __restrict__ mt, const float4* __restrict__ pt,
                                                                  float4* __restrict__ cur, float4* __restrict__ prev, uint32_t n, int prev_is_default) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t id = gid ? gid[r] : r;
    float4 r0 = make_float4(1.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 1.f, 0.f, 0.f), r2 = make_float4(0.f, 0.f, 1.f, 0.f);
    if (w3) { r0 = w3[3ull * id]; r1 = w3[3ull * id + 1]; r2 = w3[3ull * id + 2]; }      // static, stored in id order
    const float4 m = mt[r], p = pt[r];
    float4* c = cur + 4ull * id; float4* q = prev + 4ull * id;
    c[0] = make_float4(r0.x, r0.y, r0.z, m.x); c[1] = make_float4(r1.x, r1.y, r1.z, m.y); c[2] = make_float4(r2.x, r2.y, r2.z, m.z); c[3] = make_float4(0.f, 0.f, 0.f, 1.f);
    // before the first tick _prevMVP still holds MVP()'s default model matrix: identity (math/src/Matrix.cpp:7-16)
    if (prev_is_default) { r0 = make_float4(1.f, 0.f, 0.f, 0.f); r1 = make_float4(0.f, 1.f, 0.f, 0.f); r2 = make_float4(0.f, 0.f, 1.f, 0.f); }
    q[0] = make_float4(r0.x, r0.y, r0.z, p.x); q[1] = make_float4(r1.x, r1.y, r1.z, p.y); q[2] = make_float4(r2.x, r2.y, r2.z, p.z); q[3] = make_float4(0.f, 0.f, 0.f, 1.f);
}

// ------------------------------------------------------------------------------------------
// K7: spatial re-ordering of the body rows (single-sphere worlds). Bodies are gathered and scattered by
// body row all over the step; keeping the rows in the order of the last counting sort turns those random
// 32-B sector accesses into streams. Row order is invisible outside: ids travel with the rows (gid / inv).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RB_BLOCK) rb_reorder_kernel(RbParams P, RbReorder R) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_rows = P.dn ? P.dn[0] : P.nb;
    if (i >= n_rows) {
        // in-tick re-ordering of a slab world: this tick's ghost rows (end of the arrays) move to the new buffers as they are
        if (R.dst_key && P.dn && i < P.nb && i >= P.nb - P.dn[1]) {
#pragma unroll 1
            for (int a = 0; a < R.n4; ++a) R.dst4[a][i] = R.src4[a][i];
            R.dst_flags[i] = R.src_flags[i]; R.dst_key[i] = R.src_key[i]; R.dst_gid[i] = R.src_gid[i];
        }
        return;
    }
    uint32_t k = P.key[i];
    if (k != RB_NONE) k &= ~RB_HOT_BIT;
    uint32_t j;
    if (R.owned_rank) {      // slab world: ghosts share the sorted array, owned rows keep their relative order
        const uint32_t n_sorted = P.col_start[P.scx * P.scz];
        if (k != RB_NONE) j = R.owned_rank[P.col_start[k] + P.rank[i]];
        else j = R.owned_rank[n_sorted] + atomicAdd(R.rest_counter, 1u);
    } else {
        if (k != RB_NONE) j = P.col_start[k] + P.rank[i];                                    // its slot in the last sort
        else j = P.col_start[P.scx * P.scz] + atomicAdd(R.rest_counter, 1u);                 // not in the grid: behind
    }
#pragma unroll 1
    for (int a = 0; a < R.n4; ++a) R.dst4[a][j] = R.src4[a][i];
    const uint32_t fl = R.src_flags[i];
    R.dst_flags[j] = fl;
    if (R.dst_key) {
        R.dst_key[j] = R.src_key[i];
        if (k != RB_NONE) R.sorted_id_fix[j] = j | ((fl & RB_F_CONTACT) ? RB_TRI_BIT : 0u);      // slot j now holds row j
    }
    uint32_t g = R.src_gid ? R.src_gid[i] : i;
    R.dst_gid[j] = g;
    if (R.inv) R.inv[g] = j;
}
// slab worlds: 1 for sorted slots holding an owned row, 0 for ghosts (and beyond the sorted range)
__global__ void __launch_bounds__(RB_BLOCK) rb_owned_flags_kernel(RbParams P, uint32_t* out, uint32_t n) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t n_sorted = P.col_start[P.scx * P.scz];
    out[s] = (s < n_sorted && (P.sorted_id[s] & ~RB_TRI_BIT) < P.dn[0]) ? 1u : 0u;
}

} // namespace RB_KNS
using namespace RB_KNS;

// ------------------------------------------------------------------------------------------
// host-callable launchers (called from rb_world.cu)
// ------------------------------------------------------------------------------------------
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

#ifdef RB_CHAIN_TU
extern "C++" {
void rb_launch_vl_gate(const RbParams& P, cudaStream_t st) { rb_vl_gate_kernel<<<1, 32, 0, st>>>(P); }
bool rb_have_device_launch() { return true; }
}
#else
extern "C++" {
#ifndef RB_HAVE_CDP      /* built without the relocatable pass: the host always enqueues the rebuild chain itself */
void rb_launch_vl_gate(const RbParams&, cudaStream_t) {}
bool rb_have_device_launch() { return false; }
#endif
void rb_launch_integrate_bin(const RbParams& P, int do_integrate, int do_bin, cudaStream_t st) {
    if (P.nb) rb_integrate_bin_kernel<<<cdiv(P.nb, RB_BLOCK), RB_BLOCK, 0, st>>>(P, do_integrate, do_bin);
    if (P.nb && !P.single && !P.kstride && P.ns && do_bin) rb_bin_spheres_kernel<<<cdiv(P.ns, RB_BLOCK), RB_BLOCK, 0, st>>>(P, do_bin);      // (without binning a single world's K1 leaves the spheres alone)
}
// scans col_count[0..n) -> col_start[0..n]; tile_sums must hold cdiv(n, RB_SCAN_TILE) entries
int rb_launch_scan(const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* tile_sums, cudaStream_t st) {
    uint32_t ntiles = cdiv(n, RB_SCAN_TILE);
    rb_scan_tile_sums_kernel<<<ntiles, RB_BLOCK, 0, st>>>(in, n, tile_sums);
    rb_scan_top_kernel<<<1, RB_BLOCK, 0, st>>>(tile_sums, ntiles, out + n);
    rb_scan_apply_kernel<<<ntiles, RB_BLOCK, 0, st>>>(in, n, tile_sums, out);
    return 3;
}
uint32_t rb_scan_tiles(uint32_t n) { return cdiv(n, RB_SCAN_TILE); }
// single launch; status/ticket must have been zeroed (they live behind the histogram, see rb_build)
int rb_launch_scan_lookback(const uint32_t* in, uint32_t n, uint32_t* out, unsigned long long* status, uint32_t* ticket, cudaStream_t st) {
    rb_scan_lookback_kernel<<<cdiv(n, RB_SCAN_TILE), RB_BLOCK, 0, st>>>(const_cast<uint32_t*>(in), n, out, status, ticket, nullptr, 0);
    return 1;
}
void rb_launch_scatter(const RbParams& P, cudaStream_t st) {
    if (P.ns) rb_scatter_kernel<<<cdiv(P.ns, RB_BLOCK), RB_BLOCK, 0, st>>>(P, 0, nullptr, nullptr, 0);
}
// persistent lists: the rebuild chain (bin -> scan -> scatter -> enumerate), every kernel conditional on the
// device-side rebuild flag, followed by the unconditional test / resolve kernel over the lists
// zero exactly `bytes_a` bytes at a (16-byte aligned, a multiple of 4 bytes) and `words_b` 32-bit words at b (either may be empty)
void rb_launch_clear(void* a, size_t bytes_a, uint32_t* b, uint32_t words_b, cudaStream_t st) {
    const uint32_t n16 = (uint32_t)(bytes_a / 16), ntail = (uint32_t)((bytes_a % 16) / 4);
    const uint32_t n = n16 > words_b ? (n16 > ntail ? n16 : ntail) : (words_b > ntail ? words_b : ntail);
    if (n) rb_clear_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>((uint4*)a, n16, (uint32_t*)a + 4ull * n16, ntail, b, words_b);
}
void rb_launch_vl_bin(const RbParams& P, cudaStream_t st) {
    if (P.ns) rb_vl_bin_kernel<<<cdiv(P.ns, RB_BLOCK), RB_BLOCK, 0, st>>>(P);
}
void rb_launch_vl_scan(const RbParams& P, uint32_t ncols, unsigned long long* status, uint32_t* ticket, cudaStream_t st) {
    rb_scan_lookback_kernel<<<cdiv(ncols, RB_SCAN_TILE), RB_BLOCK, 0, st>>>(P.col_count, ncols, P.col_start, status, ticket, P.vl + RB_VL_REBUILD, 1);
}
void rb_launch_vl_scatter(const RbParams& P, uint32_t ncols, unsigned long long* status, uint32_t* ticket, cudaStream_t st) {
    if (P.ns) rb_scatter_kernel<<<cdiv(P.ns, RB_BLOCK), RB_BLOCK, 0, st>>>(P, 1, status, ticket, cdiv(ncols, RB_SCAN_TILE));
}
// The hot bodies' launch is small and latency-bound (a few warps walking dependent loads): it runs on a side
// stream next to the big list kernel. Both only write the bodies they own and read frozen data of the others.
uint32_t rb_list_tile_bytes() { return RB_LS_SMEM_BYTES; }
uint32_t rb_list_tile_rows() { return RB_LS_BLOCK; }
uint32_t rb_list_partner_slots() { return RB_SSLIST_MAX; }
uint32_t rb_compound_tri_slots() { return RB_CAND_MAX; }
int rb_launch_vl_collide(const RbParams& P, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join, bool with_enumerate, uint32_t hot_rows, int phase) {
    if (!P.nb) return 0;
    const uint32_t g = cdiv(P.ns, RB_LS_BLOCK);
    if (with_enumerate) rb_collide_single_kernel<true, 1><<<g, RB_LS_BLOCK, 0, st>>>(P, 0);
    static const int side_mode = [] { const char* v = getenv("RB_HOT_SIDE"); return v && *v ? atoi(v) : 1; }();      // (development switch)
    if (side_mode == 0) {       // everything on one stream: hot bodies first
        rb_collide_single_kernel<true, 3><<<cdiv(hot_rows, RB_BLOCK), RB_BLOCK, 0, st>>>(P, phase);
        rb_list_step_kernel<<<g, RB_LS_BLOCK, rb_list_tile_bytes(), st>>>(P, phase);
        return 3;
    }
    if (side_mode == 2) {       // one stream, hot bodies last
        rb_list_step_kernel<<<g, RB_LS_BLOCK, rb_list_tile_bytes(), st>>>(P, phase);
        rb_collide_single_kernel<true, 3><<<cdiv(hot_rows, RB_BLOCK), RB_BLOCK, 0, st>>>(P, phase);
        return 3;
    }
    cudaEventRecord(fork, st);
    cudaStreamWaitEvent(side, fork, 0);
    rb_collide_single_kernel<true, 3><<<cdiv(hot_rows, RB_BLOCK), RB_BLOCK, 0, side>>>(P, phase);     // hot bodies: the blocks stride over the table (a block without work returns at once)
    cudaEventRecord(join, side);
    rb_list_step_kernel<<<g, RB_LS_BLOCK, rb_list_tile_bytes(), st>>>(P, phase);                         // everybody else, from the lists
    cudaStreamWaitEvent(st, join, 0);
    return 3;
}
// slab worlds: the early part of the list step (blocks no ghost can reach), before the halo wait
void rb_launch_list_step_early(const RbParams& P, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join, uint32_t hot_rows) {
    if (!P.nb) return;
    cudaEventRecord(fork, st);
    cudaStreamWaitEvent(side, fork, 0);
    rb_collide_single_kernel<true, 3><<<cdiv(hot_rows, RB_BLOCK), RB_BLOCK, 0, side>>>(P, 1);
    cudaEventRecord(join, side);
    rb_list_step_kernel<<<cdiv(P.ns, RB_LS_BLOCK), RB_LS_BLOCK, rb_list_tile_bytes(), st>>>(P, 1);
    cudaStreamWaitEvent(st, join, 0);
}
int rb_launch_collide(const RbParams& P, bool resolve, int variant, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join) {
    if (!P.nb) return 0;
    if (P.single) {
        uint32_t g = cdiv(P.ns, RB_BLOCK);
        if (variant == 1) {     // first-generation kernel, kept for A/B checks (RB_COLLIDE_V1=1)
            if (resolve) rb_collide_kernel<true, true><<<g, RB_BLOCK, 0, st>>>(P);
            else rb_collide_kernel<true, false><<<g, RB_BLOCK, 0, st>>>(P);
        } else {                                   // fused phases in one launch
            if (resolve) rb_collide_single_kernel<true, 0><<<g, RB_BLOCK, 0, st>>>(P, 0);
            else rb_collide_single_kernel<false, 0><<<g, RB_BLOCK, 0, st>>>(P, 0);
        }
        return 1;
    }
    uint32_t g = cdiv(P.nb, RB_BLOCK);
    // K4G (a group of lanes per body) is bit-identical to the per-body kernel and faster while bodies fall (C4: 0.94 against 1.20 ms)
    // but not once they rest (1.56 against 1.51 ms: the speculative rounds re-test and four groups of a warp wait for each other):
    // opt-in until the enumeration is split from the resolution (RB_COMPOUND_GROUP=1)
    if (variant != 1 && P.grp_on == 2) {        // enumeration per sphere, resolution per body (K4E + K4R)
        const uint32_t nrows = P.kstride ? P.nb * P.kstride : P.ns;
        // RB_COMPOUND_PARTNERS=body (development switch): the partner scan per body (K4P) instead of per sphere row -- fewer column
        // reads, but 8x fewer threads with longer serial walks: measured slower on C4 (collide stage 0.62 against 0.58 ms)
        static const int per_body = [] { const char* v = getenv("RB_COMPOUND_PARTNERS"); return v && !strcmp(v, "body") ? 1 : 0; }();
        if (per_body) rb_compound_partner_kernel<<<g, RB_BLOCK, 0, st>>>(P);
        else rb_compound_enum_kernel<1><<<cdiv(nrows, RB_BLOCK), RB_BLOCK, 0, st>>>(P, nrows);
        // (the sphere-sphere launch is sized for every body: the list's length only exists on the device; blocks beyond it return at once)
        // RB_COMPOUND_SS_GROUP=1 (development switch): a group of lanes per listed body instead of one thread -- measured slower on C4
        // (0.57 against 0.25 ms: bodies in sphere-sphere contact hit many times per tick, every hit costs the group another round)
        static const int ss_group = [] { const char* v = getenv("RB_COMPOUND_SS_GROUP"); return v && *v == '1' ? 1 : 0; }();
        const uint32_t gs = cdiv(P.nb << P.grp_lg, RB_BLOCK);
        cudaStream_t s1 = side ? side : st;
        if (side) { cudaEventRecord(fork, st); cudaStreamWaitEvent(side, fork, 0); }
        if (resolve) {
            if (ss_group) rb_compound_ss_group_kernel<true><<<gs, RB_BLOCK, 0, s1>>>(P);
            else rb_compound_resolve_kernel<true, 1><<<cdiv(P.nb, RB_SS_BLOCK), RB_SS_BLOCK, 0, s1>>>(P);
            if (side) cudaEventRecord(join, side);
            rb_compound_enum_kernel<2><<<cdiv(nrows, RB_BLOCK), RB_BLOCK, 0, st>>>(P, nrows);
            rb_compound_resolve_kernel<true, 2><<<g, RB_BLOCK, 0, st>>>(P);
            if (side) cudaStreamWaitEvent(st, join, 0);
            rb_compound_resolve_kernel<true, 3><<<g, RB_BLOCK, 0, st>>>(P);
            rb_commit_kernel<<<g, RB_BLOCK, 0, st>>>(P);
            return 6;
        }
        if (ss_group) rb_compound_ss_group_kernel<false><<<gs, RB_BLOCK, 0, s1>>>(P);
        else rb_compound_resolve_kernel<false, 1><<<cdiv(P.nb, RB_SS_BLOCK), RB_SS_BLOCK, 0, s1>>>(P);
        if (side) cudaEventRecord(join, side);
        rb_compound_enum_kernel<2><<<cdiv(nrows, RB_BLOCK), RB_BLOCK, 0, st>>>(P, nrows);
        rb_compound_resolve_kernel<false, 2><<<g, RB_BLOCK, 0, st>>>(P);
        if (side) cudaStreamWaitEvent(st, join, 0);
        rb_compound_resolve_kernel<false, 3><<<g, RB_BLOCK, 0, st>>>(P);
        return 5;
    }
    if (variant != 1 && P.grp_on == 1) {
        const uint32_t gg = cdiv(P.nb << P.grp_lg, RB_BLOCK);
        if (resolve) {
            rb_collide_group_kernel<true><<<gg, RB_BLOCK, 0, st>>>(P);
            rb_commit_kernel<<<g, RB_BLOCK, 0, st>>>(P);
            return 2;
        }
        rb_collide_group_kernel<false><<<gg, RB_BLOCK, 0, st>>>(P);
        return 1;
    }
    if (resolve) {              // first-generation kernel, one thread per body (RB_COLLIDE_V1=1)
        rb_collide_kernel<false, true><<<g, RB_BLOCK, 0, st>>>(P);
        rb_commit_kernel<<<g, RB_BLOCK, 0, st>>>(P);
        return 2;
    }
    rb_collide_kernel<false, false><<<g, RB_BLOCK, 0, st>>>(P);
    return 1;
}
void rb_launch_apply_force(const uint32_t* flags, const float4* in, float4* out, const uint32_t* inv, uint32_t first, uint32_t count, int keep_w, cudaStream_t st) {
    if (count) rb_apply_force_kernel<<<cdiv(count, RB_BLOCK), RB_BLOCK, 0, st>>>(flags, in, out, inv, first, count, keep_w);
}
void rb_launch_set_flags(uint32_t* flags, const uint32_t* values, const uint32_t* inv, uint32_t first, uint32_t count, uint32_t mask, cudaStream_t st) {
    if (count) rb_set_flags_kernel<<<cdiv(count, RB_BLOCK), RB_BLOCK, 0, st>>>(flags, values, inv, first, count, mask);
}
void rb_launch_fill_f4(float4* p, uint32_t n, float4 v, cudaStream_t st) {
    if (n) rb_fill_f4_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>(p, n, v);
}
void rb_launch_rows_to_ids_f4(const float4* rows, const uint32_t* gid, float4* out, uint32_t n, int w_one, cudaStream_t st) {
    if (n) rb_rows_to_ids_f4_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>(rows, gid, out, n, w_one);
}
void rb_launch_apply_force3(const uint32_t* flags, const float* in3, float4* out, const uint32_t* inv, uint32_t count, cudaStream_t st) {
    if (count) rb_apply_force3_kernel<<<cdiv(count, RB_BLOCK), RB_BLOCK, 0, st>>>(flags, in3, out, inv, count);
}
void rb_launch_rows_to_ids_f3(const float4* rows, const uint32_t* gid, float* out3, uint32_t n, cudaStream_t st) {
    if (n) rb_rows_to_ids_f3_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>(rows, gid, out3, n);
}
void rb_launch_rows_to_ids_u32(const uint32_t* rows, const uint32_t* gid, uint32_t* out, uint32_t n, cudaStream_t st) {
    if (n) rb_rows_to_ids_u32_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>(rows, gid, out, n);
}
void rb_launch_ids_to_rows_f4(const float4* in, const uint32_t* inv, float4* rows, uint32_t first, uint32_t count, int keep_w, cudaStream_t st) {
    if (count) rb_ids_to_rows_f4_kernel<<<cdiv(count, RB_BLOCK), RB_BLOCK, 0, st>>>(in, inv, rows, first, count, keep_w);
}
void rb_launch_entity_set_position(float4* pos, float4* mt, float4* pt, const float4* wt, const uint32_t* inv, uint32_t body, const float* xyz, cudaStream_t st) {
    rb_entity_set_position_kernel<<<1, 32, 0, st>>>(pos, mt, pt, wt, inv, body, xyz[0], xyz[1], xyz[2]);
}
void rb_launch_mt_refresh(const float4* pos, const float4*, float4* mt, uint32_t n, cudaStream_t st) {
    if (n) rb_mt_refresh_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>(pos, mt, n);
}
void rb_launch_model_matrices(const float4* w3, const uint32_t* gid, const float4* mt, const float4* pt, float4* cur, float4* prev, uint32_t n, int prev_is_default, cudaStream_t st) {
    if (n) rb_model_matrix_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>(w3, gid, mt, pt, cur, prev, n, prev_is_default);
}
void rb_launch_owned_flags(const RbParams& P, uint32_t* out, uint32_t n, cudaStream_t st) {
    if (n) rb_owned_flags_kernel<<<cdiv(n, RB_BLOCK), RB_BLOCK, 0, st>>>(P, out, n);
}
void rb_launch_reorder(const RbParams& P, const void* R, cudaStream_t st) {
    if (P.nb) rb_reorder_kernel<<<cdiv(P.nb, RB_BLOCK), RB_BLOCK, 0, st>>>(P, *(const RbReorder*)R);
}
}
#endif // RB_CHAIN_TU
