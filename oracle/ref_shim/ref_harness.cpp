/* ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref).
 *
 * Headless driver around the UNMODIFIED reference translation units
 *   math/src/{Vector4,Matrix,StateVector,MVP,GeometryMath}.cpp
 *   physics/src/{Cube,Sphere,Triangle,Geometry,GeometryGraphic,OSP,Physics}.cpp
 *   engine/src/MasterClock.cpp
 * which are compiled where they lie under /root/reference by oracle/Makefile and linked with this
 * file into oracle/_ref/libref_physics.so. Nothing here is product code; only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the resulting library.
 *
 * The only reference logic restated here is the three Entity methods the physics path calls
 * (model/src/Entity.cpp:45-48 geometry copy, :131-142 _updateKinematics, :373-395 setPosition /
 * setVelocity); the real Entity.cpp drags in IO/Frustum/Shader brokers and cannot be compiled.
 *
 * Hit capture uses GNU ld --wrap on the four GeometryMath detection/resolution entry points that
 * Physics.o calls across the object boundary, so reference sources stay untouched.
 * Compiled with -fno-access-control (Physics::_physicsProcess and GeometryMath::_closestPoint are
 * private).
 */
#include "Physics.h"
#include "GeometryMath.h"
#include "ShaderBroker.h"
#include <chrono>
#include <cstdint>
#include <cstring>

/* ---- Entity restatement (model/src/Entity.cpp) ------------------------------------------- */
Entity::Entity(Model* model, Matrix worldSpaceTransform) : _model(model) {
    _worldSpaceTransform = worldSpaceTransform;          /* Entity.cpp:24 */
    _worldSpaceGeometry  = *_model->getGeometry();       /* Entity.cpp:45 */
    /* Entity.cpp:48 subscribes _updateKinematics to the MasterClock; the harness drives the
     * entities itself, in construction order, which is the order the clock would call them. */
    _state.setContact(false);                            /* uninitialised in the reference */
}
void Entity::_updateKinematics(int milliSeconds) {       /* Entity.cpp:131-142 */
    _state.update(milliSeconds);
    _prevMVP.setModel(_mvp.getModelMatrix());
    Vector4 position          = _state.getLinearPosition();
    Matrix kinematicTransform = Matrix::translation(position.getx(), position.gety(), position.getz());
    auto totalTransform       = kinematicTransform * _worldSpaceTransform;
    _worldSpaceGeometry.updateTransform(totalTransform);
    _mvp.setModel(totalTransform);
}
void Entity::setPosition(Vector4 position) {             /* Entity.cpp:373-391 */
    Matrix kinematicTransform = Matrix::translation(position.getx(), position.gety(), position.getz());
    auto totalTransform       = kinematicTransform * _worldSpaceTransform;
    Vector4 pos = Vector4(totalTransform.getFlatBuffer()[3],
                          totalTransform.getFlatBuffer()[7],
                          totalTransform.getFlatBuffer()[11]);
    _state.setLinearPosition(pos);
    _worldSpaceGeometry.updateTransform(totalTransform);
    _prevMVP.setModel(_mvp.getModelMatrix());
    _mvp.setModel(totalTransform);
}
void Entity::setVelocity(Vector4 velocity) {             /* Entity.cpp:393-395 */
    _state.setLinearVelocity(velocity);
}

/* ---- world ------------------------------------------------------------------------------- */
struct Hit {
    int32_t kind, eA, sA, eB, pB;      /* kind 0 = sphere-sphere, 1 = sphere-triangle */
    float depth, nx, ny, nz, px, py, pz, r;
};
struct RefWorld {
    std::vector<Model*>  models;
    std::vector<Entity*> entities;
    std::unordered_map<Entity*, int> index;
    Physics* phys = nullptr;
    std::vector<Hit> hits;
    long tests_ss = 0, tests_st = 0, hits_ss = 0, hits_st = 0;
    bool frozen = false;
    std::vector<int32_t> viz;          /* (sphere entity, triangle entity, triangle idx) triples */
};
static RefWorld* g_cur = nullptr;      /* world whose step is running (single-threaded) */

static int sphere_index(Entity* e, Sphere* s) {
    auto* v = e->getGeometry()->getSpheres();
    return (int)(s - v->data());
}
static int tri_index(Entity* e, Triangle* t) {
    auto* v = e->getGeometry()->getTriangles();
    return (int)(t - v->data());
}

#include <limits>
#include "ModelBroker.h"
extern "C" {
/* mangled: GeometryMath::sphereSphereDetection(Sphere&, Sphere&) etc. */
bool __real__ZN12GeometryMath21sphereSphereDetectionER6SphereS1_(Sphere*, Sphere*);
bool __real__ZN12GeometryMath23sphereTriangleDetectionER6SphereR8Triangle(Sphere*, Triangle*);
void __real__ZN12GeometryMath24sphereTriangleResolutionEP6EntityR6SphereS1_R8Triangle(Entity*, Sphere*, Entity*, Triangle*);
void __real__ZN12GeometryMath22sphereSphereResolutionEP6EntityR6SphereS1_S3_(Entity*, Sphere*, Entity*, Sphere*);

bool __wrap__ZN12GeometryMath21sphereSphereDetectionER6SphereS1_(Sphere* a, Sphere* b) {
    if (g_cur) g_cur->tests_ss++;
    return __real__ZN12GeometryMath21sphereSphereDetectionER6SphereS1_(a, b);
}
bool __wrap__ZN12GeometryMath23sphereTriangleDetectionER6SphereR8Triangle(Sphere* s, Triangle* t) {
    if (g_cur) g_cur->tests_st++;
    return __real__ZN12GeometryMath23sphereTriangleDetectionER6SphereR8Triangle(s, t);
}
void __wrap__ZN12GeometryMath24sphereTriangleResolutionEP6EntityR6SphereS1_R8Triangle(
        Entity* A, Sphere* s, Entity* B, Triangle* t) {
    if (g_cur) {
        Hit h;
        h.kind = 1; h.eA = g_cur->index[A]; h.sA = sphere_index(A, s);
        h.eB = g_cur->index[B]; h.pB = tri_index(B, t);
        Vector4 P = s->getPosition(); float r = s->getRadius();
        Vector4 q = GeometryMath::_closestPoint(s, t);
        Vector4 o = P - q; float d = o.getMagnitude(); o.normalize();
        h.depth = r - d; h.nx = o.getx(); h.ny = o.gety(); h.nz = o.getz();
        h.px = P.getx(); h.py = P.gety(); h.pz = P.getz(); h.r = r;
        g_cur->hits.push_back(h); g_cur->hits_st++;
        if (g_cur->frozen) return;
    }
    __real__ZN12GeometryMath24sphereTriangleResolutionEP6EntityR6SphereS1_R8Triangle(A, s, B, t);
}
void __wrap__ZN12GeometryMath22sphereSphereResolutionEP6EntityR6SphereS1_S3_(
        Entity* A, Sphere* sa, Entity* B, Sphere* sb) {
    if (g_cur) {
        Hit h;
        h.kind = 0; h.eA = g_cur->index[A]; h.sA = sphere_index(A, sa);
        h.eB = g_cur->index[B]; h.pB = sphere_index(B, sb);
        Vector4 ca = sa->getPosition(), cb = sb->getPosition();
        Vector4 n = ca - cb; float d = n.getMagnitude(); n.normalize();
        h.depth = (sa->getRadius() + sb->getRadius()) - d;
        h.nx = n.getx(); h.ny = n.gety(); h.nz = n.getz();
        h.px = ca.getx(); h.py = ca.gety(); h.pz = ca.getz(); h.r = sa->getRadius();
        g_cur->hits.push_back(h); g_cur->hits_ss++;
        if (g_cur->frozen) return;
    }
    __real__ZN12GeometryMath22sphereSphereResolutionEP6EntityR6SphereS1_S3_(A, sa, B, sb);
}

/* ---- C API (ctypes) ---------------------------------------------------------------------- */
void* ref_world_create() { return new RefWorld(); }
void  ref_world_destroy(void* w) { delete (RefWorld*)w; /* entities/octree leak like the reference */ }

/* Sphere-type entity. W = translation(wt) * scale(scale) (wt may be NULL). center_radius is 4n
 * object-space floats (x,y,z,r). */
int ref_add_sphere_entity(void* wp, int n, const float* center_radius, float scale, const float* wt,
                          const float* pos3, const float* vel3, int active, int gravity) {
    RefWorld* w = (RefWorld*)wp;
    Model* m = new Model(GeometryType::Sphere);
    for (int i = 0; i < n; i++) {
        const float* c = center_radius + 4 * i;
        m->getGeometry()->addSphere(Sphere(c[3], Vector4(c[0], c[1], c[2], 1.0f)));
    }
    Matrix W = Matrix::scale(scale);
    if (wt) W = Matrix::translation(wt[0], wt[1], wt[2]) * W;
    Entity* e = new Entity(m, W);
    StateVector* s = e->getStateVector();
    s->setLinearPosition(Vector4(pos3[0], pos3[1], pos3[2], 1.0f));
    s->setLinearVelocity(Vector4(vel3[0], vel3[1], vel3[2], 1.0f));
    s->setActive(active != 0);
    s->setGravity(gravity != 0);
    w->models.push_back(m);
    w->index[e] = (int)w->entities.size();
    w->entities.push_back(e);
    return (int)w->entities.size() - 1;
}
/* Triangle-type (static) entity; xyz9 = ntri * (A.xyz, B.xyz, C.xyz) in world coordinates. */
int ref_add_triangle_entity(void* wp, int ntri, const float* xyz9) {
    RefWorld* w = (RefWorld*)wp;
    Model* m = new Model(GeometryType::Triangle);
    for (int i = 0; i < ntri; i++) {
        const float* p = xyz9 + 9 * i;
        m->getGeometry()->addTriangle(Triangle(Vector4(p[0], p[1], p[2], 1.0f),
                                               Vector4(p[3], p[4], p[5], 1.0f),
                                               Vector4(p[6], p[7], p[8], 1.0f)));
    }
    Entity* e = new Entity(m, Matrix());
    w->models.push_back(m);
    w->index[e] = (int)w->entities.size();
    w->entities.push_back(e);
    return (int)w->entities.size() - 1;
}
/* Physics(); addEntities(all); run(). Sphere transforms are primed with _updateKinematics(0) so
 * the octree is built from the initial poses. Returns build seconds. */
double ref_build(void* wp) {
    RefWorld* w = (RefWorld*)wp;
    for (Entity* e : w->entities) e->_updateKinematics(0);
    auto t0 = std::chrono::steady_clock::now();
    w->phys = new Physics();
    w->phys->addEntities(w->entities);
    w->phys->run();
    auto t1 = std::chrono::steady_clock::now();
    DebugShader* dbg = static_cast<DebugShader*>(ShaderBroker::instance()->getShader("debugShader"));
    (void)dbg;
    return std::chrono::duration<double>(t1 - t0).count();
}
void ref_integrate(void* wp, int ms) {
    RefWorld* w = (RefWorld*)wp;
    for (Entity* e : w->entities) e->_updateKinematics(ms);
}
void ref_collide(void* wp, int ms) {
    RefWorld* w = (RefWorld*)wp;
    g_cur = w; w->frozen = false; w->hits.clear();
    w->phys->_physicsProcess(ms);
    g_cur = nullptr;
}
/* One reference tick (engine/src/MasterClock.cpp:36-45): every Entity::_updateKinematics in
 * construction order, then Physics::_physicsProcess. */
void ref_step(void* wp, int ms) { ref_integrate(wp, ms); ref_collide(wp, ms); }

/* L1 frozen-state detection: the reference leaf loops with both resolutions turned into no-ops;
 * contact flags restored afterwards. Hits land in the hit log (duplicates across leaves kept). */
long ref_detect_frozen(void* wp) {
    RefWorld* w = (RefWorld*)wp;
    std::vector<bool> saved;
    for (Entity* e : w->entities) saved.push_back(e->getStateVector()->getContact());
    g_cur = w; w->frozen = true; w->hits.clear();
    w->phys->_physicsProcess(0);
    g_cur = nullptr; w->frozen = false;
    for (size_t i = 0; i < w->entities.size(); i++) w->entities[i]->getStateVector()->setContact(saved[i]);
    w->phys->visualize();   /* drops _triangleIntersectionList */
    return (long)w->hits.size();
}
long ref_hit_count(void* wp) { return (long)((RefWorld*)wp)->hits.size(); }
void ref_get_hits(void* wp, int32_t* ints5, float* floats8) {
    RefWorld* w = (RefWorld*)wp;
    for (size_t i = 0; i < w->hits.size(); i++) {
        const Hit& h = w->hits[i];
        int32_t* a = ints5 + 5 * i; float* f = floats8 + 8 * i;
        a[0] = h.kind; a[1] = h.eA; a[2] = h.sA; a[3] = h.eB; a[4] = h.pB;
        f[0] = h.depth; f[1] = h.nx; f[2] = h.ny; f[3] = h.nz; f[4] = h.px; f[5] = h.py; f[6] = h.pz; f[7] = h.r;
    }
}
/* The reference's own query channel: Physics::visualize() payload, as (sphere entity, triangle
 * entity, triangle index) triples; clears the list like the reference does. Returns triple count;
 * call with out == NULL first to size (the list is cached until the next call with out != NULL). */
long ref_visualize(void* wp, int32_t* out) {
    RefWorld* w = (RefWorld*)wp;
    if (!out) {
        w->viz.clear();
        DebugShader* dbg = static_cast<DebugShader*>(ShaderBroker::instance()->getShader("debugShader"));
        dbg->sink = [w](MVP* mvp, std::set<Triangle*>& tris) {
            int se = -1;
            for (size_t i = 0; i < w->entities.size(); i++) if (w->entities[i]->getMVP() == mvp) { se = (int)i; break; }
            for (Triangle* t : tris) {
                for (size_t j = 0; j < w->entities.size(); j++) {
                    auto* v = w->entities[j]->getGeometry()->getTriangles();
                    if (!v->empty() && t >= v->data() && t < v->data() + v->size()) {
                        w->viz.push_back(se); w->viz.push_back((int)j); w->viz.push_back((int)(t - v->data()));
                        break;
                    }
                }
            }
        };
        w->phys->visualize();
        dbg->sink = nullptr;
        return (long)(w->viz.size() / 3);
    }
    memcpy(out, w->viz.data(), w->viz.size() * sizeof(int32_t));
    return (long)(w->viz.size() / 3);
}
void ref_get_counters(void* wp, long* out4) {
    RefWorld* w = (RefWorld*)wp;
    out4[0] = w->tests_ss; out4[1] = w->tests_st; out4[2] = w->hits_ss; out4[3] = w->hits_st;
}
void ref_reset_counters(void* wp) {
    RefWorld* w = (RefWorld*)wp; w->tests_ss = w->tests_st = w->hits_ss = w->hits_st = 0;
}
int ref_entity_count(void* wp) { return (int)((RefWorld*)wp)->entities.size(); }
/* per entity: pos3, vel3, flags (bit0 active, bit1 gravity, bit2 contact), model translation (mvp
 * model [3],[7],[11]) and previous model translation */
void ref_get_state(void* wp, float* pos, float* vel, int32_t* flags, float* model_t, float* prev_t) {
    RefWorld* w = (RefWorld*)wp;
    for (size_t i = 0; i < w->entities.size(); i++) {
        Entity* e = w->entities[i]; StateVector* s = e->getStateVector();
        Vector4 p = s->getLinearPosition(), v = s->getLinearVelocity();
        if (pos) { pos[3*i] = p.getx(); pos[3*i+1] = p.gety(); pos[3*i+2] = p.getz(); }
        if (vel) { vel[3*i] = v.getx(); vel[3*i+1] = v.gety(); vel[3*i+2] = v.getz(); }
        if (flags) flags[i] = (s->getActive() ? 1 : 0) | (s->_gravity ? 2 : 0) | (s->getContact() ? 4 : 0);
        float* m = e->getMVP()->getModelBuffer(); float* pm = e->getPrevMVP()->getModelBuffer();
        if (model_t) { model_t[3*i] = m[3]; model_t[3*i+1] = m[7]; model_t[3*i+2] = m[11]; }
        if (prev_t) { prev_t[3*i] = pm[3]; prev_t[3*i+1] = pm[7]; prev_t[3*i+2] = pm[11]; }
    }
}
/* full model matrices (first 16 floats of Matrix::getFlatBuffer) of every entity */
void ref_get_model_matrices(void* wp, float* cur16, float* prev16) {
    RefWorld* w = (RefWorld*)wp;
    for (size_t i = 0; i < w->entities.size(); i++) {
        memcpy(cur16 + 16 * i, w->entities[i]->getMVP()->getModelBuffer(), sizeof(float) * 16);
        memcpy(prev16 + 16 * i, w->entities[i]->getPrevMVP()->getModelBuffer(), sizeof(float) * 16);
    }
}
void ref_set_state(void* wp, int i, const float* pos3, const float* vel3) {
    StateVector* s = ((RefWorld*)wp)->entities[i]->getStateVector();
    if (pos3) s->setLinearPosition(Vector4(pos3[0], pos3[1], pos3[2], 1.0f));
    if (vel3) s->setLinearVelocity(Vector4(vel3[0], vel3[1], vel3[2], 1.0f));
}
void ref_set_flags(void* wp, int i, int active, int gravity, int contact) {
    StateVector* s = ((RefWorld*)wp)->entities[i]->getStateVector();
    if (active >= 0) s->setActive(active != 0);
    if (gravity >= 0) s->setGravity(gravity != 0);
    if (contact >= 0) s->setContact(contact != 0);
}
/* StateVector::setForce / setTorque with the reference's contact gating (StateVector.cpp:147-167) */
void ref_set_force(void* wp, int i, const float* f3) {
    ((RefWorld*)wp)->entities[i]->getStateVector()->setForce(Vector4(f3[0], f3[1], f3[2], 1.0f));
}
void ref_set_torque(void* wp, int i, const float* t3) {
    ((RefWorld*)wp)->entities[i]->getStateVector()->setTorque(Vector4(t3[0], t3[1], t3[2], 1.0f));
}
void ref_get_angular(void* wp, float* apos, float* avel) {
    RefWorld* w = (RefWorld*)wp;
    for (size_t i = 0; i < w->entities.size(); i++) {
        StateVector* s = w->entities[i]->getStateVector();
        Vector4 p = s->getAngularPosition(), v = s->getAngularVelocity();
        apos[3*i] = p.getx(); apos[3*i+1] = p.gety(); apos[3*i+2] = p.getz();
        avel[3*i] = v.getx(); avel[3*i+1] = v.gety(); avel[3*i+2] = v.getz();
    }
}
/* world-space spheres of one entity (x,y,z,r per sphere) exactly as Sphere::getPosition/getRadius see them */
void ref_get_world_spheres(void* wp, int i, float* out4) {
    auto* v = ((RefWorld*)wp)->entities[i]->getGeometry()->getSpheres();
    for (size_t k = 0; k < v->size(); k++) {
        Vector4 p = (*v)[k].getPosition();
        out4[4*k] = p.getx(); out4[4*k+1] = p.gety(); out4[4*k+2] = p.getz(); out4[4*k+3] = (*v)[k].getRadius();
    }
}
long ref_leaf_count(void* wp) { return (long)((RefWorld*)wp)->phys->_octalSpacePartioner.getOSPLeaves()->size(); }
/* leaf cubes (cx,cy,cz,edge) + number of triangles / spheres registered in each */
void ref_get_leaves(void* wp, float* cube4, int32_t* ntri, int32_t* nsph) {
    auto* leaves = ((RefWorld*)wp)->phys->_octalSpacePartioner.getOSPLeaves();
    for (size_t i = 0; i < leaves->size(); i++) {
        Cube* c = (*leaves)[i]->getData(); Vector4 ctr = c->getCenter();
        cube4[4*i] = ctr.getx(); cube4[4*i+1] = ctr.gety(); cube4[4*i+2] = ctr.getz(); cube4[4*i+3] = c->getLength();
        int nt = 0, ns = 0;
        for (auto& kv : *(*leaves)[i]->getTriangles()) nt += (int)kv.second.size();
        for (auto& kv : *(*leaves)[i]->getSpheres()) ns += (int)kv.second.size();
        ntri[i] = nt; nsph[i] = ns;
    }
}

/* ---- predicate / arithmetic micro entry points (golden-vector generation) ---------------- */
int ref_pred_sphere_triangle(const float* c3, float r, const float* t9) {
    Sphere s(r, Vector4(c3[0], c3[1], c3[2], 1.0f));
    Triangle t(Vector4(t9[0], t9[1], t9[2], 1.0f), Vector4(t9[3], t9[4], t9[5], 1.0f), Vector4(t9[6], t9[7], t9[8], 1.0f));
    return __real__ZN12GeometryMath23sphereTriangleDetectionER6SphereR8Triangle(&s, &t) ? 1 : 0;
}
int ref_pred_sphere_sphere(const float* a4, const float* b4) {
    Sphere a(a4[3], Vector4(a4[0], a4[1], a4[2], 1.0f)), b(b4[3], Vector4(b4[0], b4[1], b4[2], 1.0f));
    return __real__ZN12GeometryMath21sphereSphereDetectionER6SphereS1_(&a, &b) ? 1 : 0;
}
/* cube5 = cx,cy,cz, then length(x) height(y) width(z) */
int ref_pred_sphere_cube(const float* s4, const float* cube6) {
    Sphere s(s4[3], Vector4(s4[0], s4[1], s4[2], 1.0f));
    Cube c(cube6[3], cube6[4], cube6[5], Vector4(cube6[0], cube6[1], cube6[2], 1.0f));
    return GeometryMath::sphereCubeDetection(&s, &c) ? 1 : 0;
}
int ref_pred_triangle_cube(const float* t9, const float* cube6) {
    Triangle t(Vector4(t9[0], t9[1], t9[2], 1.0f), Vector4(t9[3], t9[4], t9[5], 1.0f), Vector4(t9[6], t9[7], t9[8], 1.0f));
    Cube c(cube6[3], cube6[4], cube6[5], Vector4(cube6[0], cube6[1], cube6[2], 1.0f));
    return GeometryMath::triangleCubeDetection(&t, &c) ? 1 : 0;
}
void ref_closest_point(const float* c3, const float* t9, float* out3) {
    Sphere s(1.0f, Vector4(c3[0], c3[1], c3[2], 1.0f));
    Triangle t(Vector4(t9[0], t9[1], t9[2], 1.0f), Vector4(t9[3], t9[4], t9[5], 1.0f), Vector4(t9[6], t9[7], t9[8], 1.0f));
    Vector4 q = GeometryMath::_closestPoint(&s, &t);
    out3[0] = q.getx(); out3[1] = q.gety(); out3[2] = q.getz();
}
/* StateVector::update on a bare state. io = pos3 vel3 apos3 avel3 force3 torque3 (18 floats),
 * flags bit0 active bit1 gravity bit2 contact; force/torque are stored raw (no setForce gating). */
void ref_state_update(float* io18, int flags, int ms, int nsteps) {
    StateVector s;
    s.setLinearPosition(Vector4(io18[0], io18[1], io18[2], 1.0f));
    s.setLinearVelocity(Vector4(io18[3], io18[4], io18[5], 1.0f));
    s.setAngularPosition(Vector4(io18[6], io18[7], io18[8], 1.0f));
    s.setAngularVelocity(Vector4(io18[9], io18[10], io18[11], 1.0f));
    s._force  = Vector4(io18[12], io18[13], io18[14], 1.0f);
    s._torque = Vector4(io18[15], io18[16], io18[17], 1.0f);
    s.setActive(flags & 1); s.setGravity((flags & 2) != 0); s.setContact((flags & 4) != 0);
    for (int k = 0; k < nsteps; k++) s.update(ms);
    Vector4 p = s.getLinearPosition(), v = s.getLinearVelocity(), ap = s.getAngularPosition(), av = s.getAngularVelocity();
    io18[0] = p.getx(); io18[1] = p.gety(); io18[2] = p.getz();
    io18[3] = v.getx(); io18[4] = v.gety(); io18[5] = v.getz();
    io18[6] = ap.getx(); io18[7] = ap.gety(); io18[8] = ap.getz();
    io18[9] = av.getx(); io18[10] = av.gety(); io18[11] = av.getz();
}
/* times nsteps reference ticks; returns seconds */
/* ---- frustum culling: the reference's own functions ------------------------------ */
static Matrix mat16(const float* m16) { float b[MATRIX_SIZE] = {0}; for (int i = 0; i < 16; i++) b[i] = m16[i]; b[24] = 1.0f; return Matrix(b); }
/* GeometryMath::getFrustumPlanes (math/src/GeometryMath.cpp:630-731): 6 planes (left,right,far,near,bottom,top) x (nx,ny,nz,d) */
void ref_frustum_planes(const float* inv_vp16, float* planes24) {
    std::vector<Vector4> planes;
    GeometryMath::getFrustumPlanes(mat16(inv_vp16), planes);
    for (int p = 0; p < 6; p++) { float* f = planes[p].getFlatBuffer(); for (int k = 0; k < 4; k++) planes24[4*p+k] = f[k]; }
}
/* GeometryMath::frustumAABBDetection (math/src/GeometryMath.cpp:733-757) over n boxes */
void ref_frustum_aabb(long n, const float* planes24, const float* mins3, const float* maxs3, int32_t* out) {
    std::vector<Vector4> planes;
    for (int p = 0; p < 6; p++) planes.push_back(Vector4(planes24[4*p], planes24[4*p+1], planes24[4*p+2], planes24[4*p+3]));
    for (long i = 0; i < n; i++) {
        Vector4 mn(mins3[3*i], mins3[3*i+1], mins3[3*i+2]), mx(maxs3[3*i], maxs3[3*i+1], maxs3[3*i+2]);
        out[i] = GeometryMath::frustumAABBDetection(planes, mn, mx) ? 1 : 0;
    }
}
/* inverseViewProjection exactly as OSP::getVisibleFrustumCulling builds it (physics/src/OSP.cpp:129-130) */
void ref_inverse_view_projection(const float* view16, const float* proj16, float* out16) {
    Matrix r = mat16(view16).inverse() * mat16(proj16).inverse();
    for (int i = 0; i < 16; i++) out16[i] = r.getFlatBuffer()[i];
}
/* OSP::getVisibleFrustumCulling (physics/src/OSP.cpp:125-180), unmodified, on the world's physics octree:
 * 1-based leaf offsets of the visible triangle-bearing leaves, front to back by |leaf centre| */
long ref_visible_frustum_culling(void* wp, const float* view16, const float* proj16, int32_t* out, long cap) {
    ModelBroker::getViewManager()->view = mat16(view16);
    ModelBroker::getViewManager()->proj = mat16(proj16);
    std::vector<int> v = ((RefWorld*)wp)->phys->_octalSpacePartioner.getVisibleFrustumCulling();
    for (size_t i = 0; i < v.size() && (long)i < cap; i++) out[i] = v[i];
    return (long)v.size();
}

/* ---- collider ingest: FbxLoader::_buildGeometryData (model/src/FbxLoader.cpp:1109-1216) cannot be
 * compiled (FBX SDK), so its loops are restated here on top of the reference's own Matrix / Vector4 arithmetic ---- */
static Matrix trs_matrix(const float* trs9) {
    Matrix rotation = Matrix::rotationAroundX(trs9[3]) * Matrix::rotationAroundY(trs9[4]) * Matrix::rotationAroundZ(trs9[5]);   /* :1118-1120 */
    Matrix translation = Matrix::translation(trs9[0], trs9[1], trs9[2]);                                                          /* :1122-1124 */
    Matrix scale = Matrix::scale(trs9[6], trs9[7], trs9[8]);                                                                      /* :1126-1128 */
    return translation * rotation * scale;                                                                                        /* :1131 */
}
void ref_trs_matrix(const float* trs9, float* out16) { Matrix m = trs_matrix(trs9); for (int i = 0; i < 16; i++) out16[i] = m.getFlatBuffer()[i]; }
/* GeometryType::Triangle branch (:1133-1160): every index triple becomes a triangle with the node TRS baked in */
void ref_ingest_triangles(long nidx, const float* xyz3, const int32_t* idx, const float* trs9, float* out_xyz9) {
    Matrix transformation = trs_matrix(trs9);
    for (long t = 0; t < nidx / 3; t++)
        for (int k = 0; k < 3; k++) {
            const float* v = xyz3 + 3 * (long)idx[3*t+k];
            Vector4 P(v[0], v[1], v[2], 1.0);
            Vector4 q = transformation * P;
            out_xyz9[9*t+3*k] = q.getx(); out_xyz9[9*t+3*k+1] = q.gety(); out_xyz9[9*t+3*k+2] = q.getz();
        }
}
/* GeometryType::Sphere branch (:1161-1215): one sphere per collider mesh, r = x-extent / 2, centre = min + r on all axes
 * (max starts at numeric_limits<float>::min(), the smallest POSITIVE float, as in the reference) */
void ref_ingest_sphere(long nidx, const float* xyz3, const int32_t* idx, const float* trs9, float* out4) {
    Matrix transformation = trs_matrix(trs9);
    float mn[3] = { (std::numeric_limits<float>::max)(), (std::numeric_limits<float>::max)(), (std::numeric_limits<float>::max)() };
    float mx[3] = { (std::numeric_limits<float>::min)(), (std::numeric_limits<float>::min)(), (std::numeric_limits<float>::min)() };
    for (long i = 0; i < nidx; i++) {
        const float* v = xyz3 + 3 * (long)idx[i];
        Vector4 vert = transformation * Vector4(v[0], v[1], v[2], 1.0f);
        float p[3] = { vert.getx(), vert.gety(), vert.getz() };
        for (int a = 0; a < 3; a++) { if (p[a] < mn[a]) mn[a] = p[a]; if (p[a] > mx[a]) mx[a] = p[a]; }
    }
    float radius = (mx[0] - mn[0]) / 2.0f;
    out4[0] = mn[0] + radius; out4[1] = mn[1] + radius; out4[2] = mn[2] + radius; out4[3] = radius;
}

double ref_bench(void* wp, int ms, int nsteps) {
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < nsteps; k++) ref_step(wp, ms);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
} /* extern "C" */
