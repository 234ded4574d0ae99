/* reboot_oracle.c -- CPU restatement of the ReBoot physics step. TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 path: plain C11, single-threaded, FP32 with no FMA
 * contraction (built with -O2 -ffp-contract=off and no -march / -ffast-math, see oracle/Makefile).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it; the product (reboot_b200/csrc) never links, imports or calls anything in oracle/.
 *
 * PINNING. The reference has no tests or golden vectors of its own, so this
 * restatement is pinned against the reference ITSELF: oracle/_ref/libref_physics.so is the
 * reference's unmodified physics/ + math/ sources compiled headless (oracle/Makefile `ref`), and
 * tests/test_oracle_vs_ref.py + the committed fixtures in tests/golden/ (generated from that
 * library by tests/golden/make_golden.py) check every function below bit-for-bit.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 *
 * Two step orders are provided:
 *   ORC_ORDER_REFERENCE  the reference's own octree broadphase and leaf loops (physics/src/OSP.cpp,
 *                        physics/src/Physics.cpp:81-208). The one thing not restated is libstdc++'s
 *                        pointer-hash iteration order of unordered_map<Entity*,...>; entities are
 *                        visited in ascending insertion index instead.
 *   ORC_ORDER_CANONICAL  the same predicates and resolutions applied per body in a fixed,
 *                        broadphase-independent order (DESIGN.md "canonical order"): the order the
 *                        parallel path implements. It coincides with the reference order whenever
 *                        a body's hits of one step fall into one octree leaf and involve at most
 *                        one partner body.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================== */
/* math/src/Vector4.cpp -- xyz-only arithmetic                                              */
/* ======================================================================================== */
typedef struct { float x, y, z; } v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }       /* Vector4.cpp:70-78 */
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }       /* :89-97 */
static inline v3 v3_mulf(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }          /* :49-56 */
static inline v3 v3_divf(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }          /* :40-47 */
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }                              /* :99-106 */
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }      /* :119-126 */
static inline v3 v3_cross(v3 a, v3 b) {                                                          /* :109-117 */
    return v3_make((a.y * b.z) - (a.z * b.y), -(a.x * b.z) + (a.z * b.x), (a.x * b.y) - (a.y * b.x));
}
static inline float v3_mag(v3 a) { return sqrtf((a.x * a.x) + (a.y * a.y) + (a.z * a.z)); }     /* :34-38 */
static inline v3 v3_normalize(v3 a) { float m = v3_mag(a); return v3_make(a.x / m, a.y / m, a.z / m); } /* :128-136 */

/* ======================================================================================== */
/* math/src/Matrix.cpp -- row-major 4x4 (first 16 of the reference's 25 floats)             */
/* ======================================================================================== */
typedef struct { float m[16]; } mat4;

static mat4 mat_identity(void) {                                                                 /* Matrix.cpp:7-16 */
    mat4 r; memset(&r, 0, sizeof r); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r;
}
static mat4 mat_translation(float x, float y, float z) {                                         /* :326-338 */
    mat4 r = mat_identity(); r.m[3] = x; r.m[7] = y; r.m[11] = z; return r;
}
static mat4 mat_mul(const mat4* a, const mat4* b) {                                              /* :136-166 */
    mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[4 * i + j] = a->m[4 * i] * b->m[j] + a->m[4 * i + 1] * b->m[4 + j] +
                             a->m[4 * i + 2] * b->m[8 + j] + a->m[4 * i + 3] * b->m[12 + j];
    return r;
}
/* Matrix * Vector4 with the input's w (Matrix.cpp:124-134); returns xyz */
static v3 mat_mulv(const mat4* a, v3 v, float w) {
    const float* m = a->m;
    return v3_make(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * w,
                   m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * w,
                   m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * w);
}

/* ======================================================================================== */
/* primitives: physics/include/{Sphere,Triangle,Cube}.h                                      */
/* ======================================================================================== */
typedef struct { v3 c; float r; } sphere_t;           /* world-space sphere */
typedef struct { v3 a, b, c; } tri_t;                 /* world-space triangle (never transformed, Geometry.cpp:33-37) */
typedef struct { v3 c; float length, height, width; } cube_t; /* Cube.h:25-39: length=x height=y width=z */

/* GeometryMath::transform(Sphere*, Matrix) -- GeometryMath.cpp:81-89 (+ Sphere.cpp:13-24) */
static sphere_t sphere_world(const mat4* total, v3 obj_c, float obj_r) {
    sphere_t s; s.c = mat_mulv(total, obj_c, 1.0f); s.r = total->m[0] * obj_r; return s;
}

/* GeometryMath::sphereTriangleDetection -- GeometryMath.cpp:101-209 */
static int sphere_triangle_detect(sphere_t s, const tri_t* t) {
    v3 A = v3_sub(t->a, s.c), B = v3_sub(t->b, s.c), C = v3_sub(t->c, s.c);
    v3 temp = v3_sub(B, A);
    float rr = s.r * s.r;
    v3 V = v3_cross(temp, v3_sub(C, A));
    float d = v3_dot(A, V);
    float e = v3_dot(V, V);
    if ((d * d) >= (rr * e)) return 0;
    float aa = v3_dot(A, A), ab = v3_dot(A, B), ac = v3_dot(A, C);
    float bb = v3_dot(B, B), bc = v3_dot(B, C), cc = v3_dot(C, C);
    if ((aa >= rr) & (ab >= aa) & (ac >= aa)) return 0;
    if ((bb >= rr) & (ab >= bb) & (bc >= bb)) return 0;
    if ((cc >= rr) & (ac >= cc) & (bc >= cc)) return 0;
    v3 AB = v3_sub(B, A), BC = v3_sub(C, B), CA = v3_sub(A, C);
    float d1 = ab - aa, d2 = bc - bb, d3 = ac - cc;
    float e1 = v3_dot(AB, AB), e2 = v3_dot(BC, BC), e3 = v3_dot(CA, CA);
    v3 Q1 = v3_sub(v3_mulf(A, e1), v3_mulf(AB, d1));
    v3 Q2 = v3_sub(v3_mulf(B, e2), v3_mulf(BC, d2));
    v3 Q3 = v3_sub(v3_mulf(C, e3), v3_mulf(CA, d3));
    v3 QC = v3_sub(v3_mulf(C, e1), Q1);
    v3 QA = v3_sub(v3_mulf(A, e2), Q2);
    v3 QB = v3_sub(v3_mulf(B, e3), Q3);
    if ((v3_dot(Q1, Q1) >= (rr * e1 * e1)) && (v3_dot(Q1, QC) >= 0.0f)) return 0;
    if ((v3_dot(Q2, Q2) >= (rr * e2 * e2)) && (v3_dot(Q2, QA) >= 0.0f)) return 0;
    if ((v3_dot(Q3, Q3) >= (rr * e3 * e3)) && (v3_dot(Q3, QB) >= 0.0f)) return 0;
    return 1;
}

/* GeometryMath::sphereSphereDetection -- GeometryMath.cpp:211-230 */
static int sphere_sphere_detect(sphere_t a, sphere_t b) {
    float dist = v3_mag(v3_sub(a.c, b.c));
    float radii = a.r + b.r;
    return (radii - dist >= 0) ? 1 : 0;
}

/* GeometryMath::sphereCubeDetection -- GeometryMath.cpp:413-487. The unit-axis dot products and
 * `step * dist` products are kept exactly as written (they round like the reference: a dot with
 * (1,0,0) is (d.x*1 + d.y*0) + d.z*0). */
static int sphere_cube_detect(sphere_t s, const cube_t* cb) {
    v3 rc = cb->c;
    float rectHeight = cb->height / 2.0f, rectWidth = cb->width / 2.0f, rectLength = cb->length / 2.0f;
    v3 nW = v3_make(1, 0, 0), nH = v3_make(0, 1, 0), nL = v3_make(0, 0, 1);
    v3 d = v3_sub(s.c, rc);
    v3 q = rc;
    float dist = v3_dot(d, nW);
    if (dist > rectLength) dist = rectLength;
    if (dist < -rectLength) dist = -rectLength;
    q = v3_add(q, v3_mulf(nW, dist));
    dist = v3_dot(d, nH);
    if (dist > rectHeight) dist = rectHeight;
    if (dist < -rectHeight) dist = -rectHeight;
    q = v3_add(q, v3_mulf(nH, dist));
    dist = v3_dot(d, nL);
    if (dist > rectWidth) dist = rectWidth;
    if (dist < -rectWidth) dist = -rectWidth;
    q = v3_add(q, v3_mulf(nL, dist));
    v3 v = v3_sub(q, s.c);
    return (v3_dot(v, v) <= s.r * s.r) ? 1 : 0;
}

static inline float fmax2(float a, float b) { return a > b ? a : b; }   /* GeometryMath.cpp:556-559 */
static inline float fmin2(float a, float b) { return a < b ? a : b; }   /* :561-564 */

/* one category-3 axis of GeometryMath::triangleCubeDetection (GeometryMath.cpp:268-275 pattern) */
static int tri_cube_axis_separates(v3 v0, v3 v1, v3 v2, v3 axis, float e0, float e1, float e2) {
    v3 u0 = v3_make(1, 0, 0), u1 = v3_make(0, 1, 0), u2 = v3_make(0, 0, 1);
    float p0 = v3_dot(v0, axis), p1 = v3_dot(v1, axis), p2 = v3_dot(v2, axis);
    float r = e0 * fabsf(v3_dot(u0, axis)) + e1 * fabsf(v3_dot(u1, axis)) + e2 * fabsf(v3_dot(u2, axis));
    return (fmax2(p2, fmax2(p0, p1)) < -r) || (fmin2(p2, fmin2(p0, p1)) > r);
}

/* GeometryMath::triangleCubeDetection -- GeometryMath.cpp:233-411 (note e0=height/2 pairs with x) */
static int triangle_cube_detect(const tri_t* t, const cube_t* cb) {
    v3 c = cb->c;
    float e0 = cb->height / 2.0f, e1 = cb->width / 2.0f, e2 = cb->length / 2.0f;
    v3 v0 = v3_sub(t->a, c), v1 = v3_sub(t->b, c), v2 = v3_sub(t->c, c);
    v3 f0 = v3_sub(v1, v0), f1 = v3_sub(v2, v1), f2 = v3_sub(v0, v2);
    v3 u0 = v3_make(1, 0, 0), u1 = v3_make(0, 1, 0), u2 = v3_make(0, 0, 1);
    v3 axes[9] = { v3_cross(u0, f0), v3_cross(u0, f1), v3_cross(u0, f2),
                   v3_cross(u1, f0), v3_cross(u1, f1), v3_cross(u1, f2),
                   v3_cross(u2, f0), v3_cross(u2, f1), v3_cross(u2, f2) };
    for (int i = 0; i < 9; i++)
        if (tri_cube_axis_separates(v0, v1, v2, axes[i], e0, e1, e2)) return 0;
    if ((fmax2(v0.x, fmax2(v1.x, v2.x)) < -e0) || (fmin2(v0.x, fmin2(v1.x, v2.x)) > e0)) return 0;
    if ((fmax2(v0.y, fmax2(v1.y, v2.y)) < -e1) || (fmin2(v0.y, fmin2(v1.y, v2.y)) > e1)) return 0;
    if ((fmax2(v0.z, fmax2(v1.z, v2.z)) < -e2) || (fmin2(v0.z, fmin2(v1.z, v2.z)) > e2)) return 0;
    v3 normal = v3_cross(f0, f1);
    float d = -v3_dot(normal, v0);
    v3 vmin, vmax;
    if (normal.x > 0.0f) { vmin.x = -e0; vmax.x = e0; } else { vmin.x = e0; vmax.x = -e0; }
    if (normal.y > 0.0f) { vmin.y = -e1; vmax.y = e1; } else { vmin.y = e1; vmax.y = -e1; }
    if (normal.z > 0.0f) { vmin.z = -e2; vmax.z = e2; } else { vmin.z = e2; vmax.z = -e2; }
    if (v3_dot(normal, vmin) + d > 0.0f) return 0;
    if (v3_dot(normal, vmax) + d >= 0.0f) return 1;
    return 0;
}

/* GeometryMath::_closestPoint -- GeometryMath.cpp:566-628 (Ericson pp.141-142) */
static v3 closest_point(v3 P, const tri_t* t) {
    v3 A = t->a, B = t->b, C = t->c;
    v3 ab = v3_sub(B, A), ac = v3_sub(C, A), ap = v3_sub(P, A);
    float d1 = v3_dot(ab, ap), d2 = v3_dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return A;
    v3 bp = v3_sub(P, B);
    float d3 = v3_dot(ab, bp), d4 = v3_dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return B;
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        float v = d1 / (d1 - d3);
        return v3_add(A, v3_mulf(ab, v));
    }
    v3 cp = v3_sub(P, C);
    float d5 = v3_dot(ab, cp), d6 = v3_dot(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return C;
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        float w = d2 / (d2 - d6);
        return v3_add(A, v3_mulf(ac, w));
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        float w = (d4 - d3) / (d4 - d3 + d5 - d6);
        return v3_add(B, v3_mulf(v3_sub(C, B), w));
    }
    float denom = 1.f / (va + vb + vc);
    float v = vb * denom;
    float w = vc * denom;
    return v3_add(v3_add(A, v3_mulf(ab, v)), v3_mulf(ac, w));
}

/* ======================================================================================== */
/* math/src/StateVector.cpp                                                                  */
/* ======================================================================================== */
#define ORC_GRAVITY (-9.8f)   /* StateVector.h:28 */
#define ORC_FRICTION (0.95f)  /* StateVector.h:30 */
#define F_ACTIVE 1
#define F_GRAVITY 2
#define F_CONTACT 4

typedef struct {
    v3 lin_pos, lin_vel, ang_pos, ang_vel, force, torque;
    float mass;
    int flags;
} state_t;

/* StateVector::update -- StateVector.cpp:11-66 */
static void state_update(state_t* s, int ms) {
    if (!(s->flags & F_ACTIVE)) return;
    float dt = (float)ms / 1000.0f;
    v3 la = v3_divf(s->force, s->mass);
    v3 aa = v3_divf(s->torque, s->mass);
    float gravity = ORC_GRAVITY;
    if (!(s->flags & F_GRAVITY)) gravity = 0.0;
    v3 dlv = v3_make(la.x * dt, (la.y + gravity) * dt, la.z * dt);
    float fr = ((s->flags & F_CONTACT) || !(s->flags & F_GRAVITY)) ? ORC_FRICTION : 1.0f;
    s->lin_vel = v3_mulf(v3_add(s->lin_vel, dlv), fr);
    v3 dlp = v3_make(s->lin_vel.x * dt, s->lin_vel.y * dt, s->lin_vel.z * dt);
    s->lin_pos = v3_add(s->lin_pos, dlp);
    v3 dav = v3_make(aa.x * dt, aa.y * dt, aa.z * dt);
    s->ang_vel = v3_mulf(v3_add(s->ang_vel, dav), fr);
    v3 dap = v3_make(s->ang_vel.x * dt, s->ang_vel.y * dt, s->ang_vel.z * dt);
    s->ang_pos = v3_add(s->ang_pos, dap);
}

/* ======================================================================================== */
/* world: entities (model/src/Entity.cpp), octree (physics/src/OSP.cpp), step (Physics.cpp)  */
/* ======================================================================================== */
typedef struct {
    int is_geom;                /* GeometryType::Triangle entity (static) vs Sphere entity (body) */
    mat4 W;                     /* Entity::_worldSpaceTransform */
    mat4 model, prev_model;     /* _mvp / _prevMVP model matrices */
    mat4 sphere_xf;             /* Sphere::_modelTransform of every sphere (Geometry::updateTransform) */
    state_t st;
    int n_sph; float* obj;      /* object-space spheres x,y,z,r */
    int n_tri; tri_t* tri;
    int tri_base;               /* global triangle id of tri[0] */
    int sph_base;               /* global sphere id of obj[0] */
} entity_t;

typedef struct { int e, p; } ref_t;  /* (entity, primitive index) */
typedef struct { ref_t* v; int n, cap; } reflist_t;

typedef struct node_s {
    cube_t cube;
    struct node_s* child[8];
    reflist_t tris, sphs;       /* sorted-unique: stands in for unordered_map<Entity*, set<...>> (OctNode.h:37-40) */
} node_t;

typedef struct {
    int32_t kind, a_body, a_sph, b_id, b_prim;   /* kind 0: b = (body, sphere); kind 1: b = (geom, tri) */
    float depth, nx, ny, nz, px, py, pz, r;
} orc_hit;

typedef struct {
    entity_t* ent; int n_ent, cap_ent;
    int n_geoms, n_bodies, n_tris_total, n_sph_total;
    node_t* root;
    node_t** leaves; int n_leaves, cap_leaves;
    float cubic_dim; int max_geoms;              /* Physics.cpp:7-9 -> 2000, 500 */
    orc_hit* hits; long n_hits, cap_hits;
    long tests_ss, tests_st;
    int record;                                  /* log hits */
    int frozen;                                  /* detection only: resolutions are no-ops */
    /* canonical-order acceleration (pure pruning, never changes results) */
    float prune_slack;
} orc_world;

static void reflist_insert(reflist_t* l, int e, int p) {
    int lo = 0, hi = l->n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        ref_t m = l->v[mid];
        if (m.e < e || (m.e == e && m.p < p)) lo = mid + 1; else hi = mid;
    }
    if (lo < l->n && l->v[lo].e == e && l->v[lo].p == p) return;
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 8; l->v = (ref_t*)realloc(l->v, sizeof(ref_t) * l->cap); }
    memmove(l->v + lo + 1, l->v + lo, sizeof(ref_t) * (l->n - lo));
    l->v[lo].e = e; l->v[lo].p = p; l->n++;
}

static node_t* node_new(cube_t c) { node_t* n = (node_t*)calloc(1, sizeof(node_t)); n->cube = c; return n; }

static void push_hit(orc_world* w, orc_hit h) {
    if (!w->record) return;
    if (w->n_hits == w->cap_hits) { w->cap_hits = w->cap_hits ? w->cap_hits * 2 : 1024; w->hits = (orc_hit*)realloc(w->hits, sizeof(orc_hit) * w->cap_hits); }
    w->hits[w->n_hits++] = h;
}

/* ---- Entity (model/src/Entity.cpp) ---- */
static void entity_set_transforms(entity_t* e, const mat4* total) { e->sphere_xf = *total; }  /* Geometry.cpp:27-31 */

/* Entity::_updateKinematics -- Entity.cpp:131-142 */
static void entity_update_kinematics(entity_t* e, int ms) {
    state_update(&e->st, ms);
    e->prev_model = e->model;
    mat4 kt = mat_translation(e->st.lin_pos.x, e->st.lin_pos.y, e->st.lin_pos.z);
    mat4 total = mat_mul(&kt, &e->W);
    entity_set_transforms(e, &total);
    e->model = total;
}
/* Entity::setPosition -- Entity.cpp:373-391 */
static void entity_set_position(entity_t* e, v3 p) {
    mat4 kt = mat_translation(p.x, p.y, p.z);
    mat4 total = mat_mul(&kt, &e->W);
    e->st.lin_pos = v3_make(total.m[3], total.m[7], total.m[11]);
    entity_set_transforms(e, &total);
    e->prev_model = e->model;
    e->model = total;
}
static sphere_t entity_sphere(const entity_t* e, int i) {  /* Sphere::getPosition / getRadius, Sphere.cpp:13-24 */
    const float* o = e->obj + 4 * i;
    return sphere_world(&e->sphere_xf, v3_make(o[0], o[1], o[2]), o[3]);
}

/* GeometryMath::sphereTriangleResolution -- GeometryMath.cpp:489-522 */
static void sphere_triangle_resolve(orc_world* w, int ea, int si, int eb, int ti) {
    entity_t* A = &w->ent[ea]; const tri_t* t = &w->ent[eb].tri[ti];
    sphere_t s = entity_sphere(A, si);
    v3 normal = v3_normalize(v3_cross(v3_sub(t->c, t->a), v3_sub(t->b, t->a)));
    v3 q = closest_point(s.c, t);
    v3 overlap_raw = v3_sub(s.c, q);
    float dist = v3_mag(overlap_raw);
    v3 overlap = v3_normalize(overlap_raw);
    orc_hit h; h.kind = 1; h.a_body = ea - w->n_geoms; h.a_sph = si; h.b_id = eb; h.b_prim = ti;
    h.depth = s.r - dist; h.nx = overlap.x; h.ny = overlap.y; h.nz = overlap.z;
    h.px = s.c.x; h.py = s.c.y; h.pz = s.c.z; h.r = s.r;
    push_hit(w, h);
    if (w->frozen) return;
    v3 delta = v3_sub(v3_add(v3_mulf(overlap, s.r * 1.0001f), q), s.c);
    float nc = v3_dot(A->st.lin_vel, normal);
    v3 n = v3_mulf(normal, nc);
    v3 rv = v3_sub(A->st.lin_vel, n);
    entity_set_position(A, v3_add(A->st.lin_pos, delta));
    A->st.lin_vel = rv;
    A->st.flags |= F_CONTACT;
}

/* GeometryMath::sphereSphereResolution -- GeometryMath.cpp:523-554 */
static void sphere_sphere_resolve(orc_world* w, int ea, int sa, int eb, int sb) {
    entity_t* A = &w->ent[ea]; entity_t* B = &w->ent[eb];
    sphere_t a = entity_sphere(A, sa), b = entity_sphere(B, sb);
    v3 cn = v3_sub(a.c, b.c);
    {
        orc_hit h; h.kind = 0; h.a_body = ea - w->n_geoms; h.a_sph = sa; h.b_id = eb - w->n_geoms; h.b_prim = sb;
        v3 nn = v3_normalize(cn);
        h.depth = (a.r + b.r) - v3_mag(cn); h.nx = nn.x; h.ny = nn.y; h.nz = nn.z;
        h.px = a.c.x; h.py = a.c.y; h.pz = a.c.z; h.r = a.r;
        push_hit(w, h);
    }
    if (w->frozen) return;
    if (A->st.flags & F_ACTIVE) {
        cn = v3_normalize(cn);
        A->st.lin_vel = v3_mulf(v3_mulf(cn, v3_mag(A->st.lin_vel)), 0.1f);
        entity_set_position(A, v3_make(A->prev_model.m[3], A->prev_model.m[7], A->prev_model.m[11]));
    }
    if (B->st.flags & F_ACTIVE) {
        cn = v3_neg(cn);
        cn = v3_normalize(cn);
        B->st.lin_vel = v3_mulf(v3_mulf(cn, v3_mag(B->st.lin_vel)), 0.1f);
        entity_set_position(B, v3_make(B->prev_model.m[3], B->prev_model.m[7], B->prev_model.m[11]));
    }
}

/* ---- OSP (physics/src/OSP.cpp) ---- */
static void leaves_push(orc_world* w, node_t* n) {
    if (w->n_leaves == w->cap_leaves) { w->cap_leaves = w->cap_leaves ? w->cap_leaves * 2 : 64; w->leaves = (node_t**)realloc(w->leaves, sizeof(node_t*) * w->cap_leaves); }
    w->leaves[w->n_leaves++] = n;
}

/* OSP::_buildOctetTree -- OSP.cpp:242-332 */
static void build_octet_tree(orc_world* w, node_t* node) {
    float cd = node->cube.length / 2.0f;
    float dim = cd / 2.0f;
    v3 pos = node->cube.c;
    static const float sg[8][3] = { {-1,-1,-1}, {-1,-1, 1}, {-1, 1, 1}, { 1,-1,-1}, { 1,-1, 1}, { 1, 1,-1}, {-1, 1,-1}, { 1, 1, 1} }; /* OSP.cpp:254-261 */
    int entered[8] = { 0 };
    for (int k = 0; k < 8; k++) {
        cube_t c; c.length = c.height = c.width = cd;
        c.c = v3_add(v3_make(sg[k][0] * dim, sg[k][1] * dim, sg[k][2] * dim), pos);
        node->child[k] = node_new(c);
    }
    for (int i = 0; i < node->tris.n; i++) {
        ref_t r = node->tris.v[i];
        for (int k = 0; k < 8; k++)
            if (triangle_cube_detect(&w->ent[r.e].tri[r.p], &node->child[k]->cube)) { reflist_insert(&node->child[k]->tris, r.e, r.p); entered[k]++; }
    }
    for (int i = 0; i < node->sphs.n; i++) {
        ref_t r = node->sphs.v[i];
        sphere_t s = entity_sphere(&w->ent[r.e], r.p);
        for (int k = 0; k < 8; k++)
            if (sphere_cube_detect(s, &node->child[k]->cube)) { reflist_insert(&node->child[k]->sphs, r.e, r.p); entered[k]++; }
    }
    for (int k = 0; k < 8; k++) {
        if (entered[k] > w->max_geoms) build_octet_tree(w, node->child[k]);
        else leaves_push(w, node->child[k]);
    }
}

/* OSP::generateGeometryOSP -- OSP.cpp:25-60 */
static void generate_geometry_osp(orc_world* w) {
    cube_t rc; rc.length = rc.height = rc.width = w->cubic_dim; rc.c = v3_make(0, 0, 0);
    w->root = node_new(rc);
    for (int e = 0; e < w->n_ent; e++) {
        entity_t* en = &w->ent[e];
        for (int i = 0; i < en->n_tri; i++) if (triangle_cube_detect(&en->tri[i], &rc)) reflist_insert(&w->root->tris, e, i);
        for (int i = 0; i < en->n_sph; i++) if (sphere_cube_detect(entity_sphere(en, i), &rc)) reflist_insert(&w->root->sphs, e, i);
    }
    build_octet_tree(w, w->root);
}

/* OSP::_insertSphereSubspaces -- OSP.cpp:212-239 */
static void insert_sphere_subspaces(orc_world* w, int e, int si, sphere_t s, node_t* node) {
    for (int k = 0; k < 8; k++) {
        node_t* ch = node->child[k];
        if (ch != NULL && sphere_cube_detect(s, &ch->cube)) {
            insert_sphere_subspaces(w, e, si, s, ch);
            reflist_insert(&ch->sphs, e, si);
        }
    }
}
/* OSP::updateOSP -- OSP.cpp:182-210 */
static void update_osp(orc_world* w) {
    for (int e = 0; e < w->n_ent; e++) {
        entity_t* en = &w->ent[e];
        if (!(en->st.flags & F_ACTIVE)) continue;
        for (int i = 0; i < en->n_sph; i++) insert_sphere_subspaces(w, e, i, entity_sphere(en, i), w->root);
    }
}

/* Physics::_physicsProcess -- Physics.cpp:81-208 */
static void physics_process_reference(orc_world* w) {
    update_osp(w);
    int n = w->n_ent;
    char* prev = (char*)malloc((size_t)n + 1); char* newc = (char*)calloc((size_t)n + 1, 1);
    for (int e = 0; e < n; e++) prev[e] = (w->ent[e].st.flags & F_CONTACT) ? 1 : 0;
    for (int li = 0; li < w->n_leaves; li++) {
        node_t* leaf = w->leaves[li];
        reflist_t* S = &leaf->sphs; reflist_t* T = &leaf->tris;
        /* sphere on sphere: entity groups are runs of equal .e in the sorted list */
        for (int a0 = 0; a0 < S->n;) {
            int ea = S->v[a0].e, a1 = a0; while (a1 < S->n && S->v[a1].e == ea) a1++;
            if (w->ent[ea].st.flags & F_ACTIVE) {
                for (int b0 = a1; b0 < S->n;) {
                    int eb = S->v[b0].e, b1 = b0; while (b1 < S->n && S->v[b1].e == eb) b1++;
                    if (w->ent[eb].st.flags & F_ACTIVE) {
                        for (int i = a0; i < a1; i++)
                            for (int j = b0; j < b1; j++) {
                                w->tests_ss++;
                                if (sphere_sphere_detect(entity_sphere(&w->ent[ea], S->v[i].p), entity_sphere(&w->ent[eb], S->v[j].p)))
                                    sphere_sphere_resolve(w, ea, S->v[i].p, eb, S->v[j].p);
                            }
                    }
                    b0 = b1;
                }
            }
            a0 = a1;
        }
        /* sphere on triangle */
        for (int a0 = 0; a0 < S->n;) {
            int ea = S->v[a0].e, a1 = a0; while (a1 < S->n && S->v[a1].e == ea) a1++;
            for (int t0 = 0; t0 < T->n;) {
                int et = T->v[t0].e, t1 = t0; while (t1 < T->n && T->v[t1].e == et) t1++;
                if ((w->ent[ea].st.flags & F_ACTIVE) || (w->ent[et].st.flags & F_ACTIVE)) {
                    for (int i = a0; i < a1; i++) {
                        int si = S->v[i].p;
                        if (sphere_cube_detect(entity_sphere(&w->ent[ea], si), &leaf->cube)) {
                            int nc = 0;
                            for (int j = t0; j < t1; j++) {
                                w->tests_st++;
                                if (sphere_triangle_detect(entity_sphere(&w->ent[ea], si), &w->ent[et].tri[T->v[j].p])) {
                                    sphere_triangle_resolve(w, ea, si, et, T->v[j].p);
                                    nc = 1;
                                }
                            }
                            newc[ea] = (char)nc;
                        }
                    }
                }
                t0 = t1;
            }
            a0 = a1;
        }
    }
    if (!w->frozen)
        for (int e = 0; e < n; e++) if (prev[e] && !newc[e]) w->ent[e].st.flags &= ~F_CONTACT;
    free(prev); free(newc);
}

/* ---- canonical order: per body, against a frozen snapshot of everybody else ---- */
typedef struct { v3 t, prev_t; } pose_t; /* model-matrix translation + previous one */

static float aabb_gap(sphere_t a, sphere_t b) {
    float dx = fabsf(a.c.x - b.c.x), dy = fabsf(a.c.y - b.c.y), dz = fabsf(a.c.z - b.c.z);
    float m = dx > dy ? dx : dy; m = m > dz ? m : dz;
    return m - (a.r + b.r);
}

static void physics_process_canonical(orc_world* w) {
    int n = w->n_ent;
    cube_t rc; rc.length = rc.height = rc.width = w->cubic_dim; rc.c = v3_make(0, 0, 0);
    /* frozen snapshot of every entity (poses after integration) */
    entity_t* snap = (entity_t*)malloc(sizeof(entity_t) * n);
    memcpy(snap, w->ent, sizeof(entity_t) * n);
    /* column grid over frozen spheres / triangles: pruning only */
    float slack = w->prune_slack;
    /* the box around every entity's frozen spheres: a pair of entities whose boxes are further apart than the slack has no
     * sphere pair within it either (the sphere-pair gap below is the Chebyshev gap of the spheres' own boxes): pruning only */
    float* ebox = (float*)malloc(sizeof(float) * 6 * (size_t)n);
    for (int e = w->n_geoms; e < n; e++) {
        float* b = ebox + 6 * (size_t)e;
        b[0] = b[1] = b[2] = 3.0e38f; b[3] = b[4] = b[5] = -3.0e38f;
        for (int i = 0; i < snap[e].n_sph; i++) {
            sphere_t q = entity_sphere(&snap[e], i);
            if (q.c.x - q.r < b[0]) b[0] = q.c.x - q.r; if (q.c.y - q.r < b[1]) b[1] = q.c.y - q.r; if (q.c.z - q.r < b[2]) b[2] = q.c.z - q.r;
            if (q.c.x + q.r > b[3]) b[3] = q.c.x + q.r; if (q.c.y + q.r > b[4]) b[4] = q.c.y + q.r; if (q.c.z + q.r > b[5]) b[5] = q.c.z + q.r;
        }
    }
    for (int ea = w->n_geoms; ea < n; ea++) {
        entity_t* A = &w->ent[ea];
        if (!(A->st.flags & F_ACTIVE)) continue;
        int prev_contact = (A->st.flags & F_CONTACT) ? 1 : 0;
        /* Q7: spheres outside the root cube never enter the tree */
        int any_in = 0;
        char* in_world = (char*)malloc(A->n_sph);
        for (int i = 0; i < A->n_sph; i++) { in_world[i] = (char)sphere_cube_detect(entity_sphere(&snap[ea], i), &rc); any_in |= in_world[i]; }
        /* ---- sphere-sphere: partner bodies ascending ---- */
        for (int eb = w->n_geoms; eb < n && any_in; eb++) {
            if (eb == ea) continue;
            const entity_t* Bf = &snap[eb];
            if (!(Bf->st.flags & F_ACTIVE)) continue;
            {   /* (boxes further apart than the slack, grown by a margin for the rounding of the box corners) */
                const float* ba = ebox + 6 * (size_t)ea; const float* bb = ebox + 6 * (size_t)eb;
                const float m = slack + 1e-3f;
                if (ba[0] - bb[3] > m || bb[0] - ba[3] > m || ba[1] - bb[4] > m || bb[1] - ba[4] > m || ba[2] - bb[5] > m || bb[2] - ba[5] > m) continue;
            }
            /* prune: some sphere pair within slack at the frozen poses */
            int near = 0;
            for (int i = 0; i < A->n_sph && !near; i++)
                for (int j = 0; j < Bf->n_sph && !near; j++)
                    if (aabb_gap(entity_sphere(&snap[ea], i), entity_sphere(Bf, j)) <= slack) near = 1;
            if (!near) continue;
            entity_t Bsim = *Bf;                      /* partner simulated from its frozen pose */
            int lo_is_a = ea < eb;
            int n_out = lo_is_a ? A->n_sph : Bsim.n_sph, n_in = lo_is_a ? Bsim.n_sph : A->n_sph;
            for (int io = 0; io < n_out; io++)
                for (int ii = 0; ii < n_in; ii++) {
                    int sa = lo_is_a ? io : ii, sb = lo_is_a ? ii : io;
                    if (!in_world[sa]) continue;
                    if (!sphere_cube_detect(entity_sphere(Bf, sb), &rc)) continue;
                    sphere_t a = entity_sphere(A, sa), b = entity_sphere(&Bsim, sb);
                    if (lo_is_a) w->tests_ss++;
                    int hit = lo_is_a ? sphere_sphere_detect(a, b) : sphere_sphere_detect(b, a);
                    if (!hit) continue;
                    /* the lower body id plays "modelA" of GeometryMath.cpp:523 */
                    v3 cn = lo_is_a ? v3_sub(a.c, b.c) : v3_sub(b.c, a.c);
                    if (lo_is_a) {
                        orc_hit h; h.kind = 0; h.a_body = ea - w->n_geoms; h.a_sph = sa; h.b_id = eb - w->n_geoms; h.b_prim = sb;
                        v3 nn = v3_normalize(cn);
                        h.depth = (a.r + b.r) - v3_mag(cn); h.nx = nn.x; h.ny = nn.y; h.nz = nn.z;
                        h.px = a.c.x; h.py = a.c.y; h.pz = a.c.z; h.r = a.r;
                        push_hit(w, h);
                    }
                    if (w->frozen) continue;
                    cn = v3_normalize(cn);
                    if (!lo_is_a) { cn = v3_neg(cn); cn = v3_normalize(cn); }
                    A->st.lin_vel = v3_mulf(v3_mulf(cn, v3_mag(A->st.lin_vel)), 0.1f);
                    entity_set_position(A, v3_make(A->prev_model.m[3], A->prev_model.m[7], A->prev_model.m[11]));
                    entity_set_position(&Bsim, v3_make(Bsim.prev_model.m[3], Bsim.prev_model.m[7], Bsim.prev_model.m[11]));
                }
        }
        /* ---- sphere-triangle: sphere ascending, global triangle id ascending ---- */
        int last_hit = 0;
        for (int si = 0; si < A->n_sph; si++) {
            int hit_any = 0;
            if (in_world[si]) {
                sphere_t fs = entity_sphere(&snap[ea], si);
                for (int eg = 0; eg < w->n_geoms; eg++) {
                    entity_t* G = &w->ent[eg];
                    for (int ti = 0; ti < G->n_tri; ti++) {
                        const tri_t* t = &G->tri[ti];
                        /* prune on the frozen sphere's AABB grown by slack (results unchanged) */
                        float R = fs.r + slack;
                        if (fmin2(t->a.x, fmin2(t->b.x, t->c.x)) > fs.c.x + R || fmax2(t->a.x, fmax2(t->b.x, t->c.x)) < fs.c.x - R) continue;
                        if (fmin2(t->a.z, fmin2(t->b.z, t->c.z)) > fs.c.z + R || fmax2(t->a.z, fmax2(t->b.z, t->c.z)) < fs.c.z - R) continue;
                        if (fmin2(t->a.y, fmin2(t->b.y, t->c.y)) > fs.c.y + R || fmax2(t->a.y, fmax2(t->b.y, t->c.y)) < fs.c.y - R) continue;
                        w->tests_st++;
                        if (sphere_triangle_detect(entity_sphere(A, si), t)) { sphere_triangle_resolve(w, ea, si, eg, ti); hit_any = 1; }
                    }
                }
            }
            if (si == A->n_sph - 1) last_hit = hit_any;
        }
        if (!w->frozen && prev_contact && !last_hit) A->st.flags &= ~F_CONTACT;
        free(in_world);
    }
    free(ebox);
    free(snap);
}

/* ======================================================================================== */
/* C API (ctypes)                                                                             */
/* ======================================================================================== */
#define ORC_ORDER_REFERENCE 0
#define ORC_ORDER_CANONICAL 1

orc_world* orc_create(void) {
    orc_world* w = (orc_world*)calloc(1, sizeof(orc_world));
    w->cubic_dim = 2000.0f; w->max_geoms = 500; w->record = 1; w->prune_slack = 1.0f;
    return w;
}
static void node_free(node_t* n) { if (!n) return; for (int k = 0; k < 8; k++) node_free(n->child[k]); free(n->tris.v); free(n->sphs.v); free(n); }
void orc_destroy(orc_world* w) {
    for (int e = 0; e < w->n_ent; e++) { free(w->ent[e].obj); free(w->ent[e].tri); }
    node_free(w->root); free(w->leaves); free(w->ent); free(w->hits); free(w);
}
static entity_t* ent_new(orc_world* w) {
    if (w->n_ent == w->cap_ent) { w->cap_ent = w->cap_ent ? w->cap_ent * 2 : 64; w->ent = (entity_t*)realloc(w->ent, sizeof(entity_t) * w->cap_ent); }
    entity_t* e = &w->ent[w->n_ent++]; memset(e, 0, sizeof *e);
    e->W = mat_identity(); e->model = mat_identity(); e->prev_model = mat_identity(); e->sphere_xf = mat_identity();
    e->st.mass = 1.0f; e->st.flags = F_GRAVITY;   /* StateVector.cpp:4-9: mass 1, inactive, gravity on; _contact forced false */
    return e;
}
/* static triangle geometry; must be added before any body (canonical insertion order) */
int orc_add_geom(orc_world* w, int ntri, const float* xyz9) {
    entity_t* e = ent_new(w); e->is_geom = 1; e->n_tri = ntri; e->tri = (tri_t*)malloc(sizeof(tri_t) * (ntri ? ntri : 1));
    for (int i = 0; i < ntri; i++) { const float* p = xyz9 + 9 * i; e->tri[i].a = v3_make(p[0], p[1], p[2]); e->tri[i].b = v3_make(p[3], p[4], p[5]); e->tri[i].c = v3_make(p[6], p[7], p[8]); }
    e->tri_base = w->n_tris_total; w->n_tris_total += ntri; w->n_geoms++;
    return w->n_geoms - 1;
}
/* sphere body; W16 row-major world transform (NULL = identity) */
int orc_add_body(orc_world* w, int nsph, const float* obj4, const float* W16, const float* pos3, const float* vel3, int active, int gravity) {
    entity_t* e = ent_new(w); e->n_sph = nsph; e->obj = (float*)malloc(sizeof(float) * 4 * (nsph ? nsph : 1));
    memcpy(e->obj, obj4, sizeof(float) * 4 * nsph);
    if (W16) memcpy(e->W.m, W16, sizeof(float) * 16);
    e->st.lin_pos = v3_make(pos3[0], pos3[1], pos3[2]); e->st.lin_vel = v3_make(vel3[0], vel3[1], vel3[2]);
    e->st.flags = (active ? F_ACTIVE : 0) | (gravity ? F_GRAVITY : 0);
    e->sph_base = w->n_sph_total; w->n_sph_total += nsph; w->n_bodies++;
    return w->n_bodies - 1;
}
/* Physics::addEntities (Physics.cpp:26-45): prime transforms, build the octree */
void orc_build(orc_world* w, int build_octree) {
    for (int e = 0; e < w->n_ent; e++) entity_update_kinematics(&w->ent[e], 0);
    if (build_octree) generate_geometry_osp(w);
}
void orc_integrate(orc_world* w, int ms) { for (int e = 0; e < w->n_ent; e++) entity_update_kinematics(&w->ent[e], ms); }
void orc_collide(orc_world* w, int order) {
    w->n_hits = 0; w->frozen = 0;
    if (order == ORC_ORDER_REFERENCE) physics_process_reference(w); else physics_process_canonical(w);
}
/* one tick: MasterClock.cpp:36-45 -> every Entity::_updateKinematics, then Physics::_physicsProcess */
void orc_step(orc_world* w, int ms, int order) { orc_integrate(w, ms); orc_collide(w, order); }
/* frozen-state detection (L1): same loops, resolutions are no-ops, flags untouched */
long orc_detect(orc_world* w, int order) {
    w->n_hits = 0; w->frozen = 1;
    if (order == ORC_ORDER_REFERENCE) physics_process_reference(w); else physics_process_canonical(w);
    w->frozen = 0;
    return w->n_hits;
}
long orc_hit_count(orc_world* w) { return w->n_hits; }
void orc_get_hits(orc_world* w, int32_t* ints5, float* floats8) {
    for (long i = 0; i < w->n_hits; i++) {
        orc_hit* h = &w->hits[i]; int32_t* a = ints5 + 5 * i; float* f = floats8 + 8 * i;
        a[0] = h->kind; a[1] = h->a_body; a[2] = h->a_sph; a[3] = h->b_id; a[4] = h->b_prim;
        f[0] = h->depth; f[1] = h->nx; f[2] = h->ny; f[3] = h->nz; f[4] = h->px; f[5] = h->py; f[6] = h->pz; f[7] = h->r;
    }
}
void orc_set_record(orc_world* w, int on) { w->record = on; }
void orc_set_prune_slack(orc_world* w, float s) { w->prune_slack = s; }
void orc_get_counters(orc_world* w, int64_t* out2) { out2[0] = w->tests_ss; out2[1] = w->tests_st; }
void orc_reset_counters(orc_world* w) { w->tests_ss = w->tests_st = 0; }
int orc_leaf_count(orc_world* w) { return w->n_leaves; }
/* body state: pos3 vel3 flags model_t prev_t (any pointer may be NULL) */
void orc_get_state(orc_world* w, float* pos, float* vel, int32_t* flags, float* model_t, float* prev_t) {
    for (int b = 0; b < w->n_bodies; b++) {
        entity_t* e = &w->ent[w->n_geoms + b];
        if (pos) { pos[3*b] = e->st.lin_pos.x; pos[3*b+1] = e->st.lin_pos.y; pos[3*b+2] = e->st.lin_pos.z; }
        if (vel) { vel[3*b] = e->st.lin_vel.x; vel[3*b+1] = e->st.lin_vel.y; vel[3*b+2] = e->st.lin_vel.z; }
        if (flags) flags[b] = e->st.flags;
        if (model_t) { model_t[3*b] = e->model.m[3]; model_t[3*b+1] = e->model.m[7]; model_t[3*b+2] = e->model.m[11]; }
        if (prev_t) { prev_t[3*b] = e->prev_model.m[3]; prev_t[3*b+1] = e->prev_model.m[7]; prev_t[3*b+2] = e->prev_model.m[11]; }
    }
}
/* Entity::getMVP()->getModelBuffer() / getPrevMVP()->getModelBuffer(): first 16 floats, row-major */
void orc_get_model_matrices(orc_world* w, float* cur16, float* prev16) {
    for (int b = 0; b < w->n_bodies; b++) {
        entity_t* e = &w->ent[w->n_geoms + b];
        memcpy(cur16 + 16 * b, e->model.m, sizeof(float) * 16);
        memcpy(prev16 + 16 * b, e->prev_model.m, sizeof(float) * 16);
    }
}
void orc_get_angular(orc_world* w, float* apos, float* avel) {
    for (int b = 0; b < w->n_bodies; b++) {
        entity_t* e = &w->ent[w->n_geoms + b];
        apos[3*b] = e->st.ang_pos.x; apos[3*b+1] = e->st.ang_pos.y; apos[3*b+2] = e->st.ang_pos.z;
        avel[3*b] = e->st.ang_vel.x; avel[3*b+1] = e->st.ang_vel.y; avel[3*b+2] = e->st.ang_vel.z;
    }
}
void orc_set_state(orc_world* w, int b, const float* pos3, const float* vel3) {
    entity_t* e = &w->ent[w->n_geoms + b];
    if (pos3) e->st.lin_pos = v3_make(pos3[0], pos3[1], pos3[2]);
    if (vel3) e->st.lin_vel = v3_make(vel3[0], vel3[1], vel3[2]);
}
/* Entity::setPosition from the host -- Entity.cpp:373-391 */
void orc_entity_set_position(orc_world* w, int b, const float* pos3) {
    entity_set_position(&w->ent[w->n_geoms + b], v3_make(pos3[0], pos3[1], pos3[2]));
}
void orc_set_flags(orc_world* w, int b, int active, int gravity, int contact) {
    entity_t* e = &w->ent[w->n_geoms + b];
    if (active >= 0) e->st.flags = (e->st.flags & ~F_ACTIVE) | (active ? F_ACTIVE : 0);
    if (gravity >= 0) e->st.flags = (e->st.flags & ~F_GRAVITY) | (gravity ? F_GRAVITY : 0);
    if (contact >= 0) e->st.flags = (e->st.flags & ~F_CONTACT) | (contact ? F_CONTACT : 0);
}
/* StateVector::setForce / setTorque -- StateVector.cpp:147-167 (zeroed unless contact || !gravity) */
void orc_set_force(orc_world* w, int b, const float* f3) {
    entity_t* e = &w->ent[w->n_geoms + b];
    if ((e->st.flags & F_CONTACT) || !(e->st.flags & F_GRAVITY)) e->st.force = v3_make(f3[0], f3[1], f3[2]); else e->st.force = v3_make(0, 0, 0);
}
void orc_set_torque(orc_world* w, int b, const float* t3) {
    entity_t* e = &w->ent[w->n_geoms + b];
    if ((e->st.flags & F_CONTACT) || !(e->st.flags & F_GRAVITY)) e->st.torque = v3_make(t3[0], t3[1], t3[2]); else e->st.torque = v3_make(0, 0, 0);
}
void orc_get_world_spheres(orc_world* w, int b, float* out4) {
    entity_t* e = &w->ent[w->n_geoms + b];
    for (int i = 0; i < e->n_sph; i++) { sphere_t s = entity_sphere(e, i); out4[4*i] = s.c.x; out4[4*i+1] = s.c.y; out4[4*i+2] = s.c.z; out4[4*i+3] = s.r; }
}

/* ---- predicate / arithmetic micro entry points ---- */
static tri_t tri_from9(const float* t) { tri_t r; r.a = v3_make(t[0], t[1], t[2]); r.b = v3_make(t[3], t[4], t[5]); r.c = v3_make(t[6], t[7], t[8]); return r; }
static cube_t cube_from6(const float* c) { cube_t r; r.c = v3_make(c[0], c[1], c[2]); r.length = c[3]; r.height = c[4]; r.width = c[5]; return r; }
void orc_pred_sphere_triangle(long n, const float* c3, const float* r, const float* t9, int32_t* out) {
    for (long i = 0; i < n; i++) { sphere_t s; s.c = v3_make(c3[3*i], c3[3*i+1], c3[3*i+2]); s.r = r[i]; tri_t t = tri_from9(t9 + 9*i); out[i] = sphere_triangle_detect(s, &t); }
}
void orc_pred_sphere_sphere(long n, const float* a4, const float* b4, int32_t* out) {
    for (long i = 0; i < n; i++) { sphere_t a, b; a.c = v3_make(a4[4*i], a4[4*i+1], a4[4*i+2]); a.r = a4[4*i+3]; b.c = v3_make(b4[4*i], b4[4*i+1], b4[4*i+2]); b.r = b4[4*i+3]; out[i] = sphere_sphere_detect(a, b); }
}
void orc_pred_sphere_cube(long n, const float* s4, const float* cube6, int32_t* out) {
    for (long i = 0; i < n; i++) { sphere_t s; s.c = v3_make(s4[4*i], s4[4*i+1], s4[4*i+2]); s.r = s4[4*i+3]; cube_t c = cube_from6(cube6 + 6*i); out[i] = sphere_cube_detect(s, &c); }
}
void orc_pred_triangle_cube(long n, const float* t9, const float* cube6, int32_t* out) {
    for (long i = 0; i < n; i++) { tri_t t = tri_from9(t9 + 9*i); cube_t c = cube_from6(cube6 + 6*i); out[i] = triangle_cube_detect(&t, &c); }
}
void orc_closest_point(long n, const float* c3, const float* t9, float* out3) {
    for (long i = 0; i < n; i++) { tri_t t = tri_from9(t9 + 9*i); v3 q = closest_point(v3_make(c3[3*i], c3[3*i+1], c3[3*i+2]), &t); out3[3*i] = q.x; out3[3*i+1] = q.y; out3[3*i+2] = q.z; }
}
/* io18 per row = pos3 vel3 apos3 avel3 force3 torque3; flags bit0 active bit1 gravity bit2 contact */
void orc_state_update(long n, float* io18, const int32_t* flags, int ms, int nsteps) {
    for (long i = 0; i < n; i++) {
        float* p = io18 + 18 * i; state_t s; memset(&s, 0, sizeof s); s.mass = 1.0f; s.flags = flags[i];
        s.lin_pos = v3_make(p[0], p[1], p[2]); s.lin_vel = v3_make(p[3], p[4], p[5]);
        s.ang_pos = v3_make(p[6], p[7], p[8]); s.ang_vel = v3_make(p[9], p[10], p[11]);
        s.force = v3_make(p[12], p[13], p[14]); s.torque = v3_make(p[15], p[16], p[17]);
        for (int k = 0; k < nsteps; k++) state_update(&s, ms);
        p[0] = s.lin_pos.x; p[1] = s.lin_pos.y; p[2] = s.lin_pos.z; p[3] = s.lin_vel.x; p[4] = s.lin_vel.y; p[5] = s.lin_vel.z;
        p[6] = s.ang_pos.x; p[7] = s.ang_pos.y; p[8] = s.ang_pos.z; p[9] = s.ang_vel.x; p[10] = s.ang_vel.y; p[11] = s.ang_vel.z;
    }
}

/* ======================================================================================== */
/* frustum culling on the spatial structure                              */
/* ======================================================================================== */
/* GeometryMath::getFrustumPlanes (math/src/GeometryMath.cpp:630-731). Planes come out in the reference's
 * order left, right, far, near, bottom, top as (nx, ny, nz, d), all four divided by |n| (Vector4::normalize,
 * Vector4.cpp:128-136). Point = M * p (Matrix.cpp:124-134) then / w (Vector4.cpp:40-47). */
void orc_frustum_planes(const float* inv_vp16, float* planes24) {
    static const float ndc[8][3] = { { -1, -1, 1 }, { -1, 1, 1 }, { -1, -1, -1 }, { -1, 1, -1 }, { 1, -1, 1 }, { 1, 1, 1 }, { 1, -1, -1 }, { 1, 1, -1 } };
    /* (origin, A = pa - origin, B = pb - origin) per plane, GeometryMath.cpp:648-730 */
    static const int pl[6][3] = { { 0, 2, 1 }, { 4, 5, 6 }, { 6, 3, 2 }, { 4, 0, 1 }, { 0, 4, 2 }, { 5, 1, 3 } };
    v3 pt[8];
    const float* m = inv_vp16;
    for (int i = 0; i < 8; i++) {
        float px = ndc[i][0], py = ndc[i][1], pz = ndc[i][2], pw = 1.0f;   /* Vector4(x,y,z) has w = 1 (Vector4.h default) */
        float x = m[0] * px + m[1] * py + m[2] * pz + m[3] * pw;
        float y = m[4] * px + m[5] * py + m[6] * pz + m[7] * pw;
        float z = m[8] * px + m[9] * py + m[10] * pz + m[11] * pw;
        float w = m[12] * px + m[13] * py + m[14] * pz + m[15] * pw;
        pt[i] = v3_make(x / w, y / w, z / w);
    }
    for (int p = 0; p < 6; p++) {
        v3 o = pt[pl[p][0]];
        v3 A = v3_sub(pt[pl[p][1]], o), B = v3_sub(pt[pl[p][2]], o);
        v3 C = v3_cross(B, A);
        float d = -(C.x * -o.x + C.y * -o.y + C.z * -o.z);
        float mag = v3_mag(C);
        planes24[4 * p] = C.x / mag; planes24[4 * p + 1] = C.y / mag; planes24[4 * p + 2] = C.z / mag; planes24[4 * p + 3] = d / mag;
    }
}
/* GeometryMath::frustumAABBDetection (math/src/GeometryMath.cpp:733-757): a box is culled iff all 8 corners are on the
 * positive side of some plane; "distance" is the plane's w, pointOnPlane = n * w (Vector4.cpp:49-56) */
static int frustum_aabb(const float* planes24, v3 mn, v3 mx) {
    const v3 c[8] = { { mn.x, mn.y, mn.z }, { mn.x, mn.y, mx.z }, { mn.x, mx.y, mx.z }, { mx.x, mx.y, mx.z },
                      { mx.x, mx.y, mn.z }, { mx.x, mn.y, mn.z }, { mx.x, mn.y, mx.z }, { mn.x, mx.y, mn.z } };
    int outside = 0;
    for (int p = 0; p < 6; p++) {
        v3 n = v3_make(planes24[4 * p], planes24[4 * p + 1], planes24[4 * p + 2]);
        float dist = planes24[4 * p + 3];
        v3 on = v3_mulf(n, dist);
        int cnt = 0;
        for (int k = 0; k < 8; k++) if (v3_dot(n, v3_sub(c[k], on)) > 0) cnt++;
        if (cnt == 8) outside = 1;
    }
    return !outside;
}
void orc_frustum_aabb(long n, const float* planes24, const float* mins3, const float* maxs3, int32_t* out) {
    for (long i = 0; i < n; i++)
        out[i] = frustum_aabb(planes24, v3_make(mins3[3 * i], mins3[3 * i + 1], mins3[3 * i + 2]), v3_make(maxs3[3 * i], maxs3[3 * i + 1], maxs3[3 * i + 2]));
}
/* the box of a cube exactly as OSP::getVisibleFrustumCulling forms it (physics/src/OSP.cpp:147-156), and its sort key
 * |centre| (:160) */
void orc_cube_box(const float* cube6, float* mins3, float* maxs3, float* key) {
    mins3[0] = cube6[0] - cube6[3] / 2.0f; mins3[1] = cube6[1] - cube6[4] / 2.0f; mins3[2] = cube6[2] - cube6[5] / 2.0f;
    maxs3[0] = cube6[0] + cube6[3] / 2.0f; maxs3[1] = cube6[1] + cube6[4] / 2.0f; maxs3[2] = cube6[2] + cube6[5] / 2.0f;
    *key = v3_mag(v3_make(cube6[0], cube6[1], cube6[2]));
}
/* OSP::getVisibleFrustumCulling (physics/src/OSP.cpp:125-180) on this world's octree: 1-based offsets of the visible
 * triangle-bearing leaves, front to back by |leaf centre| (ties: leaf order -- the reference's std::sort leaves them
 * unspecified). inv_vp16 = view.inverse() * projection.inverse(), computed by the caller. */
long orc_visible_frustum_culling(orc_world* w, const float* inv_vp16, int32_t* out, long cap) {
    float planes[24];
    orc_frustum_planes(inv_vp16, planes);
    long n = 0;
    float* keys = (float*)malloc(sizeof(float) * (size_t)(w->n_leaves > 0 ? w->n_leaves : 1));
    int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)(w->n_leaves > 0 ? w->n_leaves : 1));
    for (int li = 0; li < w->n_leaves; li++) {
        node_t* leaf = w->leaves[li];
        if (leaf->tris.n == 0) continue;
        float c6[6] = { leaf->cube.c.x, leaf->cube.c.y, leaf->cube.c.z, leaf->cube.length, leaf->cube.height, leaf->cube.width };
        float mn[3], mx[3], key;
        orc_cube_box(c6, mn, mx, &key);
        if (frustum_aabb(planes, v3_make(mn[0], mn[1], mn[2]), v3_make(mx[0], mx[1], mx[2]))) { keys[n] = key; ids[n] = li + 1; n++; }
    }
    for (long i = 1; i < n; i++) {      /* stable insertion sort by key */
        float k = keys[i]; int32_t id = ids[i]; long j = i - 1;
        while (j >= 0 && keys[j] > k) { keys[j + 1] = keys[j]; ids[j + 1] = ids[j]; j--; }
        keys[j + 1] = k; ids[j + 1] = id;
    }
    for (long i = 0; i < n && i < cap; i++) out[i] = ids[i];
    free(keys); free(ids);
    return n;
}

/* ======================================================================================== */
/* collider ingest, FbxLoader::_buildGeometryData (model/src/FbxLoader.cpp:1109-1216)   */
/* ======================================================================================== */
#define ORC_PI_OVER_180 (3.14159265f / 180.0f)           /* math/include/Matrix.h:27-28 */
/* translation * (rotX * rotY * rotZ) * scale, angles in degrees (FbxLoader.cpp:1114-1131; Matrix.cpp:261-376).
 * cos / sin are the float overloads (Matrix.cpp is compiled with `using namespace std`). */
static mat4 trs_matrix(const float* trs9) {
    float tx = trs9[3] * ORC_PI_OVER_180, ty = trs9[4] * ORC_PI_OVER_180, tz = trs9[5] * ORC_PI_OVER_180;
    mat4 rx = mat_identity(), ry = mat_identity(), rz = mat_identity(), sc = mat_identity();
    rx.m[5] = cosf(tx); rx.m[6] = sinf(tx); rx.m[9] = -sinf(tx); rx.m[10] = cosf(tx);
    ry.m[0] = cosf(ty); ry.m[2] = -sinf(ty); ry.m[8] = sinf(ty); ry.m[10] = cosf(ty);
    rz.m[0] = cosf(tz); rz.m[1] = sinf(tz); rz.m[4] = -sinf(tz); rz.m[5] = cosf(tz);
    sc.m[0] = trs9[6]; sc.m[5] = trs9[7]; sc.m[10] = trs9[8];
    mat4 t = mat_translation(trs9[0], trs9[1], trs9[2]);
    mat4 rxy = mat_mul(&rx, &ry), rot = mat_mul(&rxy, &rz);
    mat4 tr = mat_mul(&t, &rot);
    return mat_mul(&tr, &sc);
}
void orc_trs_matrix(const float* trs9, float* out16) { mat4 m = trs_matrix(trs9); memcpy(out16, m.m, sizeof m.m); }
/* GeometryType::Triangle (:1133-1160) */
void orc_ingest_triangles(long nidx, const float* xyz3, const int32_t* idx, const float* trs9, float* out_xyz9) {
    mat4 m = trs_matrix(trs9);
    for (long t = 0; t < nidx / 3; t++)
        for (int k = 0; k < 3; k++) {
            const float* v = xyz3 + 3 * (long)idx[3 * t + k];
            v3 q = mat_mulv(&m, v3_make(v[0], v[1], v[2]), 1.0f);
            out_xyz9[9 * t + 3 * k] = q.x; out_xyz9[9 * t + 3 * k + 1] = q.y; out_xyz9[9 * t + 3 * k + 2] = q.z;
        }
}
/* GeometryType::Sphere (:1161-1215): r = (max.x - min.x) / 2, centre = min + r on ALL axes; max starts at FLT_MIN
 * (numeric_limits<float>::min(), the smallest positive float), so a mesh entirely at negative x keeps max.x = FLT_MIN */
void orc_ingest_sphere(long nidx, const float* xyz3, const int32_t* idx, const float* trs9, float* out4) {
    mat4 m = trs_matrix(trs9);
    float mn[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f }, mx[3] = { 1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f };
    for (long i = 0; i < nidx; i++) {
        const float* v = xyz3 + 3 * (long)idx[i];
        v3 q = mat_mulv(&m, v3_make(v[0], v[1], v[2]), 1.0f);
        float p[3] = { q.x, q.y, q.z };
        for (int a = 0; a < 3; a++) { if (p[a] < mn[a]) mn[a] = p[a]; if (p[a] > mx[a]) mx[a] = p[a]; }
    }
    float radius = (mx[0] - mn[0]) / 2.0f;
    out4[0] = mn[0] + radius; out4[1] = mn[1] + radius; out4[2] = mn[2] + radius; out4[3] = radius;
}
