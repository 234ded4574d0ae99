/* reboot_physics.h -- C ABI of the B200-native ReBoot physics step (libreboot_b200.so).
 *
 * Drop-in boundary for the reference's physics path. Every entry point names the reference
 * interface it replaces (paths relative to the reference root). Plain C: opaque handle, plain
 * pointers and sizes, int status codes; nothing throws across this boundary. The library is
 * CUDA-only (sm_100a): there is no CPU fallback and rb_world_create fails loudly without a device.
 *
 * Reference surface being replaced
 *   class Physics            physics/include/Physics.h:29-52   (addEntities / addEntity / run / visualize)
 *   Physics::_physicsProcess physics/src/Physics.cpp:81-208    (the step)
 *   Entity::_updateKinematics, setPosition, setVelocity        model/src/Entity.cpp:131-142, 373-395
 *   class StateVector        math/include/StateVector.h:32-73  (state setters / getters, update)
 *   Geometry::addSphere / addTriangle                           physics/include/Geometry.h:42-46
 *
 * Ids are dense uint32 in insertion order: geom ids for static triangle geometry (Triangle-type
 * entities), body ids for sphere models (Sphere-type entities). Threading follows the reference:
 * one caller thread per world; rb_step only enqueues work on the world's CUDA stream, the
 * query/get calls synchronise (the reference's physics thread produces, its render thread reads
 * under Physics::_lock, Physics.h:44). An rb_world is NOT thread-safe: calls on one world must be
 * serialised by the caller (reboot::Physics in reboot/Physics.hpp holds the lock the reference holds);
 * distinct worlds are independent.
 */
#ifndef REBOOT_PHYSICS_H
#define REBOOT_PHYSICS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_world rb_world;

/* status codes (the reference returns void everywhere and fails by UB; SURVEY.md §8b "Errors") */
enum {
    RB_OK = 0,
    RB_ERR_INVALID = 1,      /* bad argument / NaN input / call in the wrong phase */
    RB_ERR_OOM = 2,          /* host or device allocation failed */
    RB_ERR_CAPACITY = 3,     /* output buffer too small; required size reported */
    RB_ERR_CUDA = 4,         /* CUDA runtime error, see rb_last_error */
    RB_ERR_NCCL = 5,         /* NCCL error, see rb_last_error */
    RB_ERR_NO_DEVICE = 6     /* no CUDA device: the product has no CPU path */
};

/* body flag bits (StateVector::_active/_gravity/_contact, math/include/StateVector.h:39-43) */
enum { RB_ACTIVE = 1u, RB_GRAVITY = 2u, RB_CONTACT = 4u };

/* Compile-time constants of the reference exposed as configuration; rb_config_default() fills the
 * reference's values. */
typedef struct rb_config {
    float    world_half;     /* 1000  : OSP root cube is 2000^3 at the origin, physics/src/Physics.cpp:7-9 */
    float    gravity;        /* -9.8  : math/include/StateVector.h:28 */
    float    friction;       /* 0.95  : math/include/StateVector.h:30 */
    float    pushout;        /* 1.0001: math/src/GeometryMath.cpp:509 */
    float    ss_damp;        /* 0.1   : math/src/GeometryMath.cpp:537,549 */
    float    margin_cap;     /* broadphase candidate margin cap in world units (ours); 0 = auto: half the largest radius */
    int32_t  device;         /* CUDA device ordinal */
    void*    stream;         /* cudaStream_t to enqueue on; NULL = the world creates its own */
    uint64_t max_contacts;   /* device contact-buffer capacity; 0 = 4 per sphere + 1024 */
    /* spatial slab owned by this world along x (multi-GPU); defaults cover the whole world */
    float    slab_lo, slab_hi;
    int32_t  rank, nranks;
    float    rm_max_hint;    /* multi-GPU: largest world-space sphere radius over ALL ranks (0 = this rank's own) */
    uint32_t halo_ghost_cap; /* multi-GPU: capacity (records) of each halo buffer. It fixes the buffer layout the neighbours
                              * store into, so it must be the SAME on every rank: pass the maximum over the ranks of
                              * rb_halo_ghost_cap_auto(). 0 = this rank's own auto value (fine when all ranks hold
                              * equally many bodies; connecting ranks with different layouts is refused). */
    /* multi-GPU worlds of compound bodies (entities with several collision spheres, physics/src/Geometry.cpp:27-31): rows
     * migrate between ranks, so every body keeps its spheres at a fixed stride -- the SAME on every rank */
    uint32_t sphere_stride;  /* largest sphere count of any body over ALL ranks (0 = this rank's own; 1 = single-sphere worlds) */
    float    body_reach_hint;/* largest max_i(|centre_i.x| + r_i) of any body over ALL ranks, world units after W (0 = this rank's own) */
    uint32_t slab_forces;    /* multi-GPU: 1 = bodies take forces / torques / angular state (StateVector::setForce / setTorque,
                              * math/src/StateVector.cpp:147-167): the arrays exist from the build on and travel with migrating
                              * bodies. The SAME on every rank. 0 = rb_set_force / rb_set_torque / rb_set_angular are refused */
} rb_config;

/* initial state of a body: StateVector setters, math/include/StateVector.h:48-60 */
typedef struct rb_body_init {
    float    pos[3];
    float    vel[3];
    uint32_t flags;          /* RB_ACTIVE | RB_GRAVITY (RB_CONTACT starts clear) */
} rb_body_init;

/* one contact of the last step / detect call.  Reference payload: Physics::_triangleIntersectionList
 * (map<Entity*, set<Triangle*>>, physics/include/Physics.h:28,36) plus the sphere-sphere hits that
 * the reference resolves but never reports; depth / normal as SURVEY.md §8(a) a15-a16 derives them. */
typedef struct rb_contact {
    uint32_t kind;           /* 0 = sphere-sphere, 1 = sphere-triangle */
    uint32_t body;           /* body id of the sphere (lower body id for sphere-sphere) */
    uint32_t sphere;         /* sphere index inside that body */
    uint32_t other;          /* kind 0: other body id; kind 1: geom id */
    uint32_t other_prim;     /* kind 0: sphere index inside the other body; kind 1: triangle index inside the geom */
    float    depth;          /* kind 0: rA+rB-|cA-cB| ; kind 1: r-|P-q| */
    float    normal[3];      /* kind 0: normalize(cA-cB) ; kind 1: normalize(P-q) */
} rb_contact;

typedef struct rb_stats_t {
    uint64_t steps;               /* rb_step calls so far */
    uint64_t tests_ss, tests_st;  /* exact narrowphase predicate evaluations (cumulative) */
    uint64_t cand_ss, cand_st;    /* broadphase candidates visited (cumulative) */
    uint64_t hits_ss, hits_st;    /* contacts recorded (cumulative) */
    uint64_t contacts_last;       /* contacts produced by the last step / detect (may exceed capacity) */
    uint64_t contacts_dropped;    /* contacts that did not fit the device buffer (cumulative) */
    uint64_t margin_overflow;     /* bodies displaced further than their candidate margin (cumulative) */
    uint32_t n_bodies, n_spheres, n_geoms, n_tris;
    uint32_t sphere_cols_x, sphere_cols_z, tri_cols_x, tri_cols_z;
    uint64_t device_bytes;        /* HBM held by this world */
    uint32_t halo_sent_last, halo_recv_last; /* multi-GPU: boundary spheres exchanged in the last step */
    uint32_t migrated_out_last, migrated_in_last;
    uint32_t list_ticks, list_rebuilds;      /* persistent candidate lists: ticks stepped from lists / ticks that rebuilt them */
    float    list_skin;                      /* the skin the lists are built with (0: lists not in use) */
    uint32_t list_hot_last;                  /* bodies that stepped without their list in the last tick (fast movers, crowded bodies) */
    uint32_t list_rebuild_reason;            /* most recent rebuild: 1 a body entered/left the grid | 2 too many hot bodies | 4 requested by the host side */
    uint32_t list_rebuilds_by_reason[4];     /* rebuild ticks by reason: a body entered/left the grid or emigrated, hot table full, host side (first tick, scheduled re-ordering, state edited), immigrants */
    uint32_t reserved0;
} rb_stats_t;

void        rb_config_default(rb_config* cfg);

/* Physics::Physics() -- physics/src/Physics.cpp:7-12 */
int         rb_world_create(const rb_config* cfg, rb_world** out);
/* Physics::~Physics() -- physics/src/Physics.cpp:14-16 (the reference frees nothing; we do) */
void        rb_world_destroy(rb_world* w);
/* last error text of this world (or of the failed create when w == NULL) */
const char* rb_last_error(const rb_world* w);

/* add-static-geometry: a Triangle-type entity. Geometry::addTriangle (physics/src/Geometry.cpp:11-13)
 * for ntri triangles; xyz9 = ntri * (A.xyz, B.xyz, C.xyz) in WORLD coordinates (the reference never
 * transforms triangles, physics/src/Geometry.cpp:33-37). Static geometry is inactive. */
int         rb_add_static_geometry(rb_world* w, uint32_t ntri, const float* xyz9, uint32_t* geom_id);

/* add-model: a Sphere-type entity. Geometry::addSphere (physics/src/Geometry.cpp:15-17) for each
 * object-space sphere (x,y,z,r) + Entity ctor (model/src/Entity.cpp:24,45): world_xform16 is the
 * row-major _worldSpaceTransform (NULL = identity; last row must be 0 0 0 1).
 * After rb_build this is Physics::addEntity (physics/include/Physics.h:49, physics/src/Physics.cpp:47-49): the body
 * takes the next id and joins the simulation with the next tick, like an entity the reference appends to its list
 * (physics/src/OSP.cpp:182-210 inserts it at the next update). Every device structure is laid out at build time, so a
 * post-build add re-lays the world out -- O(world) per call; add batches with rb_add_models. Ids, state (bit for bit),
 * forces and contacts flags of the existing bodies are kept; pending rb_step_host_async results are not (rb_sync first).
 * Single-GPU worlds only. Triangles cannot be added after the build (the reference never inserts them either, Q9). */
int         rb_add_model(rb_world* w, uint32_t nspheres, const float* obj_center_radius4,
                         const float* world_xform16, const rb_body_init* init, uint32_t* body_id);

/* bulk add-model: n bodies sharing nothing. sphere_start has n+1 CSR offsets into
 * obj_center_radius4; world_scale (nullable) is a per-body uniform scale of W; pos3/vel3 are n*3;
 * flags n. Returns the first body id. */
int         rb_add_models(rb_world* w, uint32_t n, const uint32_t* sphere_start,
                          const float* obj_center_radius4, const float* world_scale,
                          const float* pos3, const float* vel3, const uint32_t* flags,
                          uint32_t* first_body_id);

/* Physics::addEntities -- physics/src/Physics.cpp:26-45 (builds the broadphase; call once, Q9).
 * Uploads the SoA state, builds the static triangle grid, primes the model transforms like
 * Entity::_updateKinematics(0). */
int         rb_build(rb_world* w);

/* One reference tick (engine/src/MasterClock.cpp:36-45): every Entity::_updateKinematics(ms)
 * (model/src/Entity.cpp:131-142) followed by Physics::_physicsProcess(ms)
 * (physics/src/Physics.cpp:81-208). Asynchronous on the world's stream. */
int         rb_step(rb_world* w, int ms);
/* Only the Entity::_updateKinematics half of the tick (integration + sphere world transform). */
int         rb_integrate(rb_world* w, int ms);
/* Only the Physics::_physicsProcess half (broadphase + narrowphase + resolution + contact flags). */
int         rb_collide(rb_world* w);
/* Frozen-state detection: the collision half with both resolutions turned into no-ops; state and
 * contact flags untouched. Fills the contact buffer (SURVEY.md §8c level L1). */
int         rb_detect(rb_world* w);
/* block until everything enqueued so far has finished */
int         rb_sync(rb_world* w);

/* query-contacts: Physics::visualize -- physics/src/Physics.cpp:51-79. Contacts of the last
 * step/detect, de-duplicated, sorted by (kind, body, sphere, other, other_prim) when sorted != 0.
 * *n receives the number available; RB_ERR_CAPACITY if cap is too small (nothing is written). */
int         rb_query_contacts(rb_world* w, rb_contact* out, uint64_t cap, uint64_t* n, int sorted);
/* number of contacts of the last step without copying them (synchronises) */
int         rb_contact_count(rb_world* w, uint64_t* n);

/* StateVector getters, math/include/StateVector.h:62-72; any pointer may be NULL.
 * pos4/vel4: n_bodies * 4 floats (x,y,z,pad), flags: n_bodies. Synchronises. */
int         rb_get_state(rb_world* w, float* pos4, float* vel4, uint32_t* flags);
/* Entity::getMVP / getPrevMVP model-matrix translations ([3],[7],[11]) as n_bodies * 4 floats */
int         rb_get_model_translations(rb_world* w, float* model_t4, float* prev_t4);
/* Entity::getMVP()->getModelBuffer() / getPrevMVP()->getModelBuffer() (model/include/Entity.h:66-68,
 * model/src/Entity.cpp:134-140, 388-390): the full row-major 4x4 model matrices translation(p) * W of every
 * body, current and previous (motion blur), n_bodies * 16 floats each, in body-id order. Either may be NULL. */
int         rb_get_model_matrices(rb_world* w, float* model16, float* prev16);
/* The same buffers left on the device for a renderer to map (CUDA-GL/Vulkan interop): device pointers to
 * n_bodies * 16 floats each, refreshed by this call, valid until the next call or rb_world_destroy. */
int         rb_export_model_matrices(rb_world* w, const float** model16_dev, const float** prev16_dev);
/* angular state, n_bodies * 4 floats each */
int         rb_get_angular(rb_world* w, float* ang_pos4, float* ang_vel4);
/* world-space spheres as Sphere::getPosition/getRadius see them: n_spheres * 4 floats
 * (not available on multi-GPU worlds of compound bodies, whose sphere rows follow their bodies between ranks: RB_ERR_INVALID) */
int         rb_get_world_spheres(rb_world* w, float* xyzr4);

/* StateVector::setLinearPosition / setLinearVelocity (StateVector.cpp:122-135) for `count` bodies
 * starting at `first`; pos4/vel4 are count*4 floats, either may be NULL. Like the reference these
 * do not touch the model matrices until the next step. */
int         rb_set_state(rb_world* w, uint32_t first, uint32_t count, const float* pos4, const float* vel4);
/* Entity::setPosition (model/include/Entity.h:50, model/src/Entity.cpp:373-391) for one body, e.g. a teleport by game
 * logic: total = translation(pos) * W; the linear position becomes total's translation (W.t + pos: with a translated W the
 * offset is added again, like the reference), _prevMVP takes the current model matrix and _mvp becomes total. The spheres
 * follow at the next tick. Asynchronous on the world's stream. Not available on slab (multi-GPU) worlds. */
int         rb_entity_set_position(rb_world* w, uint32_t body, const float* pos3);
/* StateVector::setActive / setGravity / setContact (StateVector.cpp:117-120, 169-177): for each of
 * `count` bodies starting at `first`, flags = (flags & ~mask) | (values[i] & mask). */
int         rb_set_flags(rb_world* w, uint32_t first, uint32_t count, const uint32_t* values, uint32_t mask);
/* StateVector::setForce / setTorque (StateVector.cpp:147-167): stored only while the body is in
 * contact or has gravity off, otherwise zeroed -- evaluated on the device in stream order.
 * xyz4 = count*4 floats in pinned or pageable host memory. */
int         rb_set_force(rb_world* w, uint32_t first, uint32_t count, const float* xyz4);
int         rb_set_torque(rb_world* w, uint32_t first, uint32_t count, const float* xyz4);
int         rb_set_angular(rb_world* w, uint32_t first, uint32_t count, const float* ang_pos4, const float* ang_vel4);

/* Per-tick host round trip of a reference-style caller, with HOST buffers (bench.py "e2e"):
 * upload forces for all bodies (force4 nullable), run one tick, download linear positions (pos4
 * nullable) and the contact count, and synchronise. */
int         rb_step_host(rb_world* w, int ms, const float* force4, float* pos4, uint64_t* n_contacts);
/* The same round trip, pipelined: the H2D copy of this tick's forces, the tick, and the D2H copy of its
 * positions run on three streams over a ring of 4 staging buffers so both copy engines overlap
 * neighbouring ticks -- the reference's own model (physics thread produces, render thread reads a little
 * behind under Physics::_lock, physics/include/Physics.h:44). Returns when the pos4 buffer and contact
 * count of the call made THREE calls earlier are complete; rotate four pinned pos4 (and force4) buffers
 * and finish with rb_sync(). On ticks stepped from the persistent candidate lists the forces are taken
 * and the positions are left by the tick's own kernels (no extra pass over the bodies). */
#define RB_HOST_PIPE_DEPTH 4
int         rb_step_host_async(rb_world* w, int ms, const float* force4, float* pos4, uint64_t* n_contacts_done);
/* The same with packed xyz arrays (3 floats per body each way; masses keep their values): the pipelined
 * round trip is bound by the PCIe copies, so 12 instead of 16 bytes per body is a quarter less time. */
int         rb_step_host_async3(rb_world* w, int ms, const float* force3, float* pos3, uint64_t* n_contacts_done);
/* rows the host buffers of rb_step_host* must hold: n_bodies for a single world; for a slab (multi-GPU) world, whose
 * row count changes as bodies migrate, the most rows it can ever own. Slab worlds hand out ROWS (rb_get_body_ids says
 * which body each row is) and take no force upload (force4 must be NULL there). */
uint32_t    rb_host_row_capacity(const rb_world* w);

int         rb_get_stats(rb_world* w, rb_stats_t* out);
/* cumulative CUDA-event time (ms) of each kernel family since the last reset, when profiling is on:
 * 0 integrate+bin, 1 scan, 2 scatter, 3 collide, 4 halo. rb_profile(w, 1) turns per-kernel events on
 * (adds synchronisation; never on in timed benchmark regions). */
int         rb_profile(rb_world* w, int on);
int         rb_get_kernel_ms(rb_world* w, float* out5, uint64_t* launches);
/* total kernel launches issued by this library so far (bench.py "gpu_launches") */
uint64_t    rb_launch_count(const rb_world* w);

/* ---- multi-GPU x-slabs: one process per GPU, NCCL over NVLink ------------- */
/* size of the opaque NCCL unique id blob */
#define RB_NCCL_ID_BYTES 128
/* rank 0 creates the id, the host language broadcasts the bytes (torch.distributed / MPI / ...) */
int         rb_dist_unique_id(const char* nccl_lib_path, uint8_t id[RB_NCCL_ID_BYTES]);
/* join the communicator; cfg.rank/nranks/slab_lo/slab_hi say which slab this world owns */
int         rb_dist_init(rb_world* w, const char* nccl_lib_path, const uint8_t id[RB_NCCL_ID_BYTES]);

/* Fused compute + communication halo transport: instead of NCCL send/recv, the integrate kernel stores the
 * ghost / migrant records straight into the neighbour's buffer over NVLink (CUDA IPC peer memory) and a
 * flag handshake replaces the collective. Every rank exports its handle, the host passes each rank the
 * handles of its left / right neighbour (NULL at the world edge). Use instead of rb_dist_init. */
#define RB_IPC_HANDLE_BYTES 80   /* cudaIpcMemHandle_t (64) + the exporting rank's halo buffer layout (checked on connect) */
int         rb_dist_p2p_export(rb_world* w, uint8_t handle[RB_IPC_HANDLE_BYTES]);
int         rb_dist_p2p_connect(rb_world* w, const uint8_t* left_handle, const uint8_t* right_handle);

/* ---- callers either side of the step ---------------------------------------------------------
 * Frustum culling on the spatial structure. rb_frustum_planes = GeometryMath::getFrustumPlanes
 * (math/src/GeometryMath.cpp:630-731): six planes (left, right, far, near, bottom, top) x (nx, ny, nz, d) from the
 * inverse view-projection the reference builds as view.inverse() * projection.inverse() (physics/src/OSP.cpp:129-130),
 * row-major 4x4. Pure host function, bit-identical to the reference. */
int         rb_frustum_planes(const float inv_view_proj[16], float planes24[24]);
/* OSP::getVisibleFrustumCulling (physics/src/OSP.cpp:125-180) for n caller-supplied boxes in the reference's Cube form
 * (centre xyz, length = x, height = y, width = z; e.g. the leaves of the caller's render octree): GeometryMath::
 * frustumAABBDetection (math/src/GeometryMath.cpp:733-757) on the device, then the 0-based indices of the boxes that
 * meet the frustum, front to back by |centre| like the reference (equal distances in index order; the reference's
 * std::sort leaves them unspecified). *n_visible receives the count; RB_ERR_CAPACITY if cap is too small. */
int         rb_frustum_cull_cubes(rb_world* w, const float planes24[24], uint32_t n, const float* cubes6,
                                  uint32_t* visible, uint32_t cap, uint32_t* n_visible);
/* the same over the world's own static structure: ids of the non-empty cells of the triangle column grid that meet
 * the frustum, front to back (a renderer draws the triangles registered there). rb_get_static_cells describes the
 * cells: boxes6 = (min xyz, max xyz), ntri = triangles registered, tri_cols_x * tri_cols_z cells (rb_stats_t). */
int         rb_frustum_cull_static(rb_world* w, const float inv_view_proj[16], uint32_t* cells, uint32_t cap, uint32_t* n_visible);
int         rb_get_static_cells(rb_world* w, float* boxes6, uint32_t* ntri);

/* GeometryMath::triangleCubeDetection (math/src/GeometryMath.cpp:233-411), the predicate OSP::generateGeometryOSP hands
 * triangles down its tree with (physics/src/OSP.cpp:242-332), evaluated on the device for n (triangle, cube) pairs:
 * tri9 = n * (A, B, C), cube6 = n * (centre xyz, length, height, width) as in the reference's Cube; out[i] = 1 on overlap.
 * rb_build uses the same device function to register each static triangle in the grid columns its surface crosses. */
int         rb_triangle_cube_test(rb_world* w, uint32_t n, const float* tri9, const float* cube6, uint32_t* out);

/* Collider ingest: FbxLoader::_buildGeometryData (model/src/FbxLoader.cpp:1109-1216) for a neutral indexed mesh
 * (xyz3 vertices, idx). trs9 = node translation, rotation in degrees (x, y, z), scale: the node transform is
 * translation * rotX * rotY * rotZ * scale (:1114-1131). Pure host functions, bit-identical to the reference's
 * arithmetic.
 *   rb_ingest_mesh_triangles  static model (GeometryType::Triangle, :1133-1160): nidx / 3 world-space triangles with
 *                             the TRS baked in -> rb_add_static_geometry (rb_ingest_static_mesh does both)
 *   rb_ingest_mesh_sphere     animated model (GeometryType::Sphere, :1161-1215): ONE sphere per collider mesh,
 *                             r = (max.x - min.x) / 2, centre = min + r on all three axes (sic), max starting at
 *                             FLT_MIN (sic) -> an object-space sphere for rb_add_model */
int         rb_ingest_trs_matrix(const float trs9[9], float out16[16]);
int         rb_ingest_mesh_triangles(uint32_t nverts, const float* xyz3, uint32_t nidx, const uint32_t* idx, const float trs9[9], float* out_xyz9);
int         rb_ingest_mesh_sphere(uint32_t nverts, const float* xyz3, uint32_t nidx, const uint32_t* idx, const float trs9[9], float sphere4[4]);
int         rb_ingest_static_mesh(rb_world* w, uint32_t nverts, const float* xyz3, uint32_t nidx, const uint32_t* idx, const float trs9[9], uint32_t* geom_id);

/* the halo capacity a rank would pick by itself for n_bodies bodies of radius <= r_max in a slab of the given width
 * (pure function; see rb_config.halo_ghost_cap) */
uint32_t    rb_halo_ghost_cap_auto(uint32_t n_bodies, float r_max, float margin_cap, float slab_width);

/* global body ids of a slab world: set before rb_build (default: local insertion index), read back for
 * the bodies currently owned, in the same order as rb_get_state's rows (ownership migrates) */
int         rb_set_body_ids(rb_world* w, const uint32_t* ids, uint32_t n);
int         rb_get_body_ids(rb_world* w, uint32_t* ids_out);

/* Slab worlds of ONE process (any mix of devices; also several slabs on one device for testing):
 * attach the in-process neighbours instead of rb_dist_init, then advance all slabs in lockstep. */
int         rb_dist_attach_peers(rb_world* w, rb_world* left, rb_world* right);
int         rb_step_group(rb_world** worlds, int n, int ms);
int         rb_detect_group(rb_world** worlds, int n);

const char* rb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* REBOOT_PHYSICS_H */
