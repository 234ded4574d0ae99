// reboot/Physics.hpp -- header-only C++ facade over the C ABI (include/reboot_physics.h) that keeps
// the reference's physics-side class and method names so reference-style host code compiles against
// the B200 path unchanged:
//
//     Physics* physics = new Physics();          // physics/include/Physics.h:46
//     physics->addEntities(entityList);          // Physics.h:48 (engine/src/EngineManager.cpp:223)
//     physics->run();                            // Physics.h:51 (EngineManager.cpp:224)
//     ...
//     physics->visualize();                      // Physics.h:50 (EngineManager.cpp:470)
//
// Types mirrored (value types, no GL/FBX baggage): Vector4 (math/include/Vector4.h:25-58), Matrix
// (math/include/Matrix.h:82-117, row-major 4x4), Sphere / Triangle / Geometry / GeometryType
// (physics/include/{Sphere,Triangle,Geometry}.h), StateVector (math/include/StateVector.h:32-73),
// Model and Entity (only the members the physics path touches: model/include/Entity.h:50-68).
// The tick itself runs on the GPU; this file only marshals.
#pragma once
#include "../reboot_physics.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace reboot {

// engine/include/MasterClock.h:30-31
const int KINEMATICS_TIME = 5;   // kinematics time in milliseconds

// MasterClock (engine/include/MasterClock.h:33-74, engine/src/MasterClock.cpp:31-61, 127-131): the fixed-rate
// physics thread. Subscribers are called in subscription order every KINEMATICS_TIME ms with the tick length;
// the remainder of the tick is slept, an overrun is printed like the reference does. With the B200 path the
// per-Entity _updateKinematics subscriptions disappear: Physics::run() subscribes ONE function, the tick.
class MasterClock {
    std::vector<std::function<void(int)>> _kinematicsRateFuncs;
    std::mutex _kinematicsRateLock;
    std::thread* _physicsThread = nullptr;
    std::atomic<bool> _running{false};
    std::atomic<unsigned long> _ticks{0}, _overruns{0};
    bool _reportOverruns = true;
    void _physicsProcess() {
        while (_running) {
            auto start = std::chrono::high_resolution_clock::now();
            {
                std::lock_guard<std::mutex> g(_kinematicsRateLock);
                for (auto& func : _kinematicsRateFuncs) if (func) func(KINEMATICS_TIME);
            }
            _ticks++;
            auto end = std::chrono::high_resolution_clock::now();
            double milliseconds = std::chrono::duration<double, std::milli>(end - start).count();
            double left = static_cast<double>(KINEMATICS_TIME) - milliseconds;
            if (left > 0) std::this_thread::sleep_for(std::chrono::duration<double, std::milli>(left));
            else if (left < 0) { _overruns++; if (_reportOverruns) std::printf("Extra time being used on physics calculations: %f\n", -left); }
        }
    }
public:
    MasterClock() = default;
    ~MasterClock() { stop(); }
    static MasterClock* instance() { static MasterClock c; return &c; }
    // Physics clock time update (MasterClock.cpp:127-131)
    void subscribeKinematicsRate(std::function<void(int)> func) { std::lock_guard<std::mutex> g(_kinematicsRateLock); _kinematicsRateFuncs.push_back(func); }
    void clearKinematicsSubscribers() { std::lock_guard<std::mutex> g(_kinematicsRateLock); _kinematicsRateFuncs.clear(); }
    // Kicks off the physics thread (MasterClock.cpp:23-29 starts three threads; only the physics one is on this path)
    void run() { if (_running.exchange(true)) return; _physicsThread = new std::thread(&MasterClock::_physicsProcess, this); }
    void stop() { if (!_running.exchange(false)) return; _physicsThread->join(); delete _physicsThread; _physicsThread = nullptr; }
    void reportOverruns(bool on) { _reportOverruns = on; }
    unsigned long ticks() const { return _ticks; }
    unsigned long overruns() const { return _overruns; }
};

class Vector4 {
    float _vec[4];
public:
    Vector4() : _vec{0.f, 0.f, 0.f, 1.f} {}
    Vector4(float x, float y, float z, float w = 1.0f) : _vec{x, y, z, w} {}
    float getx() const { return _vec[0]; }
    float gety() const { return _vec[1]; }
    float getz() const { return _vec[2]; }
    float getw() const { return _vec[3]; }
    float* getFlatBuffer() { return _vec; }
};

class Matrix {
    float _matrix[16];
public:
    Matrix() : _matrix{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1} {}
    float* getFlatBuffer() { return _matrix; }
    const float* data() const { return _matrix; }
    static Matrix translation(float x, float y, float z) { Matrix m; m._matrix[3] = x; m._matrix[7] = y; m._matrix[11] = z; return m; }
    static Matrix scale(float s) { Matrix m; m._matrix[0] = m._matrix[5] = m._matrix[10] = s; return m; }
};

class Sphere {
    float _radius; Vector4 _position;
public:
    Sphere(float radius, Vector4 position) : _radius(radius), _position(position) {}
    Vector4 getObjectPosition() const { return _position; }
    float getObjectRadius() const { return _radius; }
};

class Triangle {
    Vector4 _points[3];
public:
    Triangle(Vector4 A, Vector4 B, Vector4 C) : _points{A, B, C} {}
    const Vector4* getTrianglePoints() const { return _points; }
};

enum class GeometryType { Triangle = 0, Sphere = 1 };

class Geometry {
    std::vector<Sphere> _spheres; std::vector<Triangle> _triangles;
public:
    void addTriangle(Triangle t) { _triangles.push_back(t); }
    void addSphere(Sphere s) { _spheres.push_back(s); }
    std::vector<Triangle>* getTriangles() { return &_triangles; }
    std::vector<Sphere>* getSpheres() { return &_spheres; }
};

class Model {
    GeometryType _type; Geometry _geometry;
public:
    explicit Model(GeometryType t) : _type(t) {}
    GeometryType getGeometryType() const { return _type; }
    Geometry* getGeometry() { return &_geometry; }
};

class Physics;

// MVP (math/include/MVP.h): only the model matrix is on the physics path
class MVP {
    Matrix _model;
public:
    Matrix getModelMatrix() const { return _model; }
    float* getModelBuffer() { return _model.getFlatBuffer(); }
    void setModel(Matrix m) { _model = m; }
};

// Host-side mirror of one body's StateVector. Setters before Physics::addEntities define the initial
// state; afterwards they are forwarded to the device (same gating as StateVector.cpp:147-167 for
// setForce/setTorque, applied on the device). Getters return the last Physics::syncState() snapshot.
class StateVector {
    friend class Physics; friend class Entity;
    Vector4 _linearPosition, _linearVelocity, _force, _torque;
    bool _active = false, _gravity = true, _contact = false;
    Physics* _owner = nullptr; uint32_t _body = 0;
    inline void push_pos(); inline void push_vel(); inline void push_flags(uint32_t mask); inline void push_force(); inline void push_torque();
public:
    void setLinearPosition(Vector4 p) { _linearPosition = p; push_pos(); }
    void setLinearVelocity(Vector4 v) { _linearVelocity = v; push_vel(); }
    void setActive(bool a) { _active = a; push_flags(RB_ACTIVE); }
    void setGravity(bool g) { _gravity = g; push_flags(RB_GRAVITY); }
    void setContact(bool c) { _contact = c; push_flags(RB_CONTACT); }
    void setForce(Vector4 f) { _force = f; push_force(); }
    void setTorque(Vector4 t) { _torque = t; push_torque(); }
    Vector4 getLinearPosition() const { return _linearPosition; }
    Vector4 getLinearVelocity() const { return _linearVelocity; }
    bool getActive() const { return _active; }
    bool getContact() const { return _contact; }
    float getMass() const { return 1.0f; }
};

class Entity {
    friend class Physics;
    Model* _model; Matrix _worldSpaceTransform; StateVector _state;
    Vector4 _modelTranslation, _prevModelTranslation;
    MVP _mvp, _prevMVP;     // model matrices at the last syncState()
    uint32_t _id = 0;       // geom id (Triangle type) or body id (Sphere type) once added
public:
    explicit Entity(Model* model, Matrix worldSpaceTransform = Matrix()) : _model(model), _worldSpaceTransform(worldSpaceTransform) {}
    StateVector* getStateVector() { return &_state; }
    Model* getModel() { return _model; }
    Geometry* getGeometry() { return _model->getGeometry(); }
    Matrix getWorldSpaceTransform() const { return _worldSpaceTransform; }
    // translations of Entity::getMVP()/getPrevMVP() model matrices ([3],[7],[11]) at the last syncState()
    Vector4 getModelTranslation() const { return _modelTranslation; }
    Vector4 getPrevModelTranslation() const { return _prevModelTranslation; }
    void setVelocity(Vector4 v) { _state.setLinearVelocity(v); }
    // Entity::setPosition -- model/include/Entity.h:50, model/src/Entity.cpp:373-391: total = translation(p) * W; the linear
    // position becomes total's translation (Q1), _prevMVP takes the current model matrix, _mvp becomes total. On the device
    // once the entity has been added (the mirrors here follow at the next syncState()); before that it only sets the
    // initial position.
    inline void setPosition(Vector4 position);
    // Entity::getMVP / getPrevMVP -- model/include/Entity.h:66-68: the model matrices translation(p) * W as of the last
    // Physics::syncState()
    MVP* getMVP() { return &_mvp; }
    MVP* getPrevMVP() { return &_prevMVP; }
};

// (geometry entity, triangle index) pairs per sphere entity: the reference's
// ModelIntersections = std::map<Entity*, std::set<Triangle*>> (Physics.h:28) with indices for pointers
using ModelIntersections = std::map<Entity*, std::set<std::pair<Entity*, uint32_t>>>;

class Physics {
    friend class StateVector; friend class Entity;
    rb_world* _w = nullptr;
    std::vector<Entity*> _entities, _bodies, _geoms;
    ModelIntersections _triangleIntersectionList;
    bool _built = false;
    // Physics::_lock (physics/include/Physics.h:44) + StateVector::_stateLock (math/include/StateVector.h:46) in one: the
    // clock thread ticks while render / input threads read contacts, sync state and push setters, and an rb_world must
    // be externally serialised (reboot_physics.h "Threading"). Recursive: a consumer may call back into the facade.
    std::recursive_mutex _lock;
    using Guard = std::lock_guard<std::recursive_mutex>;
    static void ck(rb_world* w, int rc) { if (rc) throw std::runtime_error(std::string("reboot_b200: ") + rb_last_error(w)); }
public:
    // Physics::Physics -- physics/src/Physics.cpp:7-12 (OSP 2000^3, leaf cap 500 -> config defaults)
    explicit Physics(const rb_config* cfg = nullptr) { int rc = rb_world_create(cfg, &_w); if (rc) throw std::runtime_error(std::string("reboot_b200: ") + rb_last_error(nullptr)); }
    ~Physics() { rb_world_destroy(_w); }
    Physics(const Physics&) = delete; Physics& operator=(const Physics&) = delete;

    // Physics::addEntities -- physics/src/Physics.cpp:26-45 (call once; builds the broadphase)
    void addEntities(std::vector<Entity*> entities) {
        for (Entity* e : entities) {
            _entities.push_back(e);
            if (e->getModel()->getGeometryType() == GeometryType::Triangle) {
                std::vector<float> xyz; xyz.reserve(e->getGeometry()->getTriangles()->size() * 9);
                for (const Triangle& t : *e->getGeometry()->getTriangles())
                    for (int k = 0; k < 3; k++) { const Vector4& p = t.getTrianglePoints()[k]; xyz.push_back(p.getx()); xyz.push_back(p.gety()); xyz.push_back(p.getz()); }
                ck(_w, rb_add_static_geometry(_w, (uint32_t)(xyz.size() / 9), xyz.data(), &e->_id));
                _geoms.push_back(e);
            } else {
                std::vector<float> obj;
                for (const Sphere& s : *e->getGeometry()->getSpheres()) { Vector4 p = s.getObjectPosition(); obj.push_back(p.getx()); obj.push_back(p.gety()); obj.push_back(p.getz()); obj.push_back(s.getObjectRadius()); }
                rb_body_init init; StateVector& st = e->_state;
                init.pos[0] = st._linearPosition.getx(); init.pos[1] = st._linearPosition.gety(); init.pos[2] = st._linearPosition.getz();
                init.vel[0] = st._linearVelocity.getx(); init.vel[1] = st._linearVelocity.gety(); init.vel[2] = st._linearVelocity.getz();
                init.flags = (st._active ? (uint32_t)RB_ACTIVE : 0u) | (st._gravity ? (uint32_t)RB_GRAVITY : 0u);
                ck(_w, rb_add_model(_w, (uint32_t)(obj.size() / 4), obj.data(), e->_worldSpaceTransform.data(), &init, &e->_id));
                st._owner = this; st._body = e->_id;
                _bodies.push_back(e);
            }
        }
        ck(_w, rb_build(_w));
        _built = true;
    }
    // Physics::addEntity -- physics/include/Physics.h:49, physics/src/Physics.cpp:47-49: one more entity after addEntities.
    // A Sphere-type entity takes part from the next tick on (rb_add_model after rb_build: the world is re-laid out, O(world)).
    // A Triangle-type entity is kept in the list but never collides: the reference only ever builds triangles into the tree in
    // addEntities -- same here.
    void addEntity(Entity* e) {
        Guard g(_lock);
        if (!_built) throw std::runtime_error("reboot_b200: Physics::addEntity before addEntities");
        _entities.push_back(e);
        if (e->getModel()->getGeometryType() == GeometryType::Triangle) return;
        std::vector<float> obj;
        for (const Sphere& s : *e->getGeometry()->getSpheres()) { Vector4 p = s.getObjectPosition(); obj.push_back(p.getx()); obj.push_back(p.gety()); obj.push_back(p.getz()); obj.push_back(s.getObjectRadius()); }
        rb_body_init init; StateVector& st = e->_state;
        init.pos[0] = st._linearPosition.getx(); init.pos[1] = st._linearPosition.gety(); init.pos[2] = st._linearPosition.getz();
        init.vel[0] = st._linearVelocity.getx(); init.vel[1] = st._linearVelocity.gety(); init.vel[2] = st._linearVelocity.getz();
        init.flags = (st._active ? (uint32_t)RB_ACTIVE : 0u) | (st._gravity ? (uint32_t)RB_GRAVITY : 0u);
        ck(_w, rb_sync(_w));
        ck(_w, rb_add_model(_w, (uint32_t)(obj.size() / 4), obj.data(), e->_worldSpaceTransform.data(), &init, &e->_id));
        st._owner = this; st._body = e->_id;
        _bodies.push_back(e);
    }
    // Physics::run -- physics/src/Physics.cpp:18-24: subscribe the step to the 5 ms kinematics tick of a
    // MasterClock (NULL: the caller drives tick() itself). One subscription replaces every Entity's
    // _updateKinematics subscription plus _physicsProcess.
    void run(MasterClock* clock = nullptr) { if (clock) clock->subscribeKinematicsRate([this](int ms) { tick(ms); }); }
    // one kinematics tick: every Entity::_updateKinematics(ms) then Physics::_physicsProcess(ms)
    void tick(int milliseconds = 5) { Guard g(_lock); ck(_w, rb_step(_w, milliseconds)); }
    // refresh every Entity's StateVector / model translations from the device
    void syncState() {
        Guard g(_lock);
        size_t n = _bodies.size(); if (!n) return;
        std::vector<float> p(n * 4), v(n * 4), m(n * 4), pm(n * 4), mm(n * 16), pmm(n * 16); std::vector<uint32_t> f(n);
        ck(_w, rb_get_state(_w, p.data(), v.data(), f.data()));
        ck(_w, rb_get_model_translations(_w, m.data(), pm.data()));
        ck(_w, rb_get_model_matrices(_w, mm.data(), pmm.data()));
        for (size_t i = 0; i < n; i++) {
            Entity* e = _bodies[i]; StateVector& s = e->_state;
            s._linearPosition = Vector4(p[4 * i], p[4 * i + 1], p[4 * i + 2]); s._linearVelocity = Vector4(v[4 * i], v[4 * i + 1], v[4 * i + 2]);
            s._active = f[i] & RB_ACTIVE; s._gravity = f[i] & RB_GRAVITY; s._contact = f[i] & RB_CONTACT;
            e->_modelTranslation = Vector4(m[4 * i], m[4 * i + 1], m[4 * i + 2]); e->_prevModelTranslation = Vector4(pm[4 * i], pm[4 * i + 1], pm[4 * i + 2]);
            for (int k = 0; k < 16; k++) { e->_mvp.getModelBuffer()[k] = mm[16 * i + k]; e->_prevMVP.getModelBuffer()[k] = pmm[16 * i + k]; }
        }
    }
    // Physics::visualize -- physics/src/Physics.cpp:51-79: hands the triangles in contact per entity to a
    // consumer (the reference's DebugShader) and clears the list.
    void visualize(const std::function<void(Entity*, const std::set<std::pair<Entity*, uint32_t>>&)>& consumer = nullptr) {
        Guard g(_lock);
        uint64_t n = 0; ck(_w, rb_contact_count(_w, &n));
        std::vector<rb_contact> c(n ? n : 1);
        ck(_w, rb_query_contacts(_w, c.data(), c.size(), &n, 1));
        for (uint64_t i = 0; i < n; i++) if (c[i].kind == 1) _triangleIntersectionList[_bodies[c[i].body]].insert({_geoms[c[i].other], c[i].other_prim});
        if (consumer) for (auto& kv : _triangleIntersectionList) consumer(kv.first, kv.second);
        _triangleIntersectionList.clear();
    }
    const ModelIntersections& intersections() {   // contacts of the last tick without clearing (the caller holds no lock: copy if threads are ticking)
        Guard g(_lock);
        _triangleIntersectionList.clear(); uint64_t n = 0; ck(_w, rb_contact_count(_w, &n));
        std::vector<rb_contact> c(n ? n : 1); ck(_w, rb_query_contacts(_w, c.data(), c.size(), &n, 1));
        for (uint64_t i = 0; i < n; i++) if (c[i].kind == 1) _triangleIntersectionList[_bodies[c[i].body]].insert({_geoms[c[i].other], c[i].other_prim});
        return _triangleIntersectionList;
    }
    // OSP::getVisibleFrustumCulling -- physics/src/OSP.cpp:125-180, over this world's own static structure: ids of the
    // non-empty static cells that meet the frustum, front to back. inverseViewProjection is what the reference builds
    // as view.inverse() * projection.inverse() (:129-130).
    std::vector<uint32_t> getVisibleFrustumCulling(const Matrix& inverseViewProjection) {
        Guard g(_lock);
        rb_stats_t st; ck(_w, rb_get_stats(_w, &st));
        std::vector<uint32_t> cells((size_t)st.tri_cols_x * st.tri_cols_z + 1); uint32_t n = 0;
        ck(_w, rb_frustum_cull_static(_w, inverseViewProjection.data(), cells.data(), (uint32_t)cells.size(), &n));
        cells.resize(n);
        return cells;
    }
    rb_world* handle() { return _w; }
};

// FbxLoader::_buildGeometryData -- model/src/FbxLoader.cpp:1109-1216 for a neutral indexed mesh: fills the model's
// Geometry exactly like the reference (static model: one triangle per index triple with the node TRS baked in;
// animated model: ONE sphere per collider mesh, r = x-extent / 2, centre = min + r on all axes).
// trs9 = LclTranslation, LclRotation (degrees), LclScaling of the node.
inline void buildGeometryData(Model* model, const std::vector<float>& xyz3, const std::vector<uint32_t>& indices, const float trs9[9]) {
    const uint32_t nv = (uint32_t)(xyz3.size() / 3), ni = (uint32_t)indices.size();
    if (model->getGeometryType() == GeometryType::Triangle) {
        std::vector<float> t((size_t)(ni / 3) * 9);
        if (rb_ingest_mesh_triangles(nv, xyz3.data(), ni, indices.data(), trs9, t.data())) throw std::runtime_error("reboot_b200: bad mesh");
        for (size_t k = 0; k + 8 < t.size(); k += 9)
            model->getGeometry()->addTriangle(Triangle(Vector4(t[k], t[k + 1], t[k + 2]), Vector4(t[k + 3], t[k + 4], t[k + 5]), Vector4(t[k + 6], t[k + 7], t[k + 8])));
    } else {
        float s[4];
        if (rb_ingest_mesh_sphere(nv, xyz3.data(), ni, indices.data(), trs9, s)) throw std::runtime_error("reboot_b200: bad mesh");
        model->getGeometry()->addSphere(Sphere(s[3], Vector4(s[0], s[1], s[2])));
    }
}

inline void Entity::setPosition(Vector4 position) {
    Physics* owner = _state._owner;
    if (!owner || !owner->_built) { _state._linearPosition = position; return; }
    std::lock_guard<std::recursive_mutex> g(owner->_lock);
    const float p[3] = {position.getx(), position.gety(), position.getz()};
    Physics::ck(owner->_w, rb_entity_set_position(owner->_w, _state._body, p));
}
inline void StateVector::push_pos() {
    if (!_owner || !_owner->_built) return;
    std::lock_guard<std::recursive_mutex> g(_owner->_lock);
    float p[4] = {_linearPosition.getx(), _linearPosition.gety(), _linearPosition.getz(), 1.f};
    Physics::ck(_owner->_w, rb_set_state(_owner->_w, _body, 1, p, nullptr));
}
inline void StateVector::push_vel() {
    if (!_owner || !_owner->_built) return;
    std::lock_guard<std::recursive_mutex> g(_owner->_lock);
    float v[4] = {_linearVelocity.getx(), _linearVelocity.gety(), _linearVelocity.getz(), 0.f};
    Physics::ck(_owner->_w, rb_set_state(_owner->_w, _body, 1, nullptr, v));
}
inline void StateVector::push_flags(uint32_t mask) {
    if (!_owner || !_owner->_built) return;
    std::lock_guard<std::recursive_mutex> g(_owner->_lock);
    uint32_t f = (_active ? (uint32_t)RB_ACTIVE : 0u) | (_gravity ? (uint32_t)RB_GRAVITY : 0u) | (_contact ? (uint32_t)RB_CONTACT : 0u);
    Physics::ck(_owner->_w, rb_set_flags(_owner->_w, _body, 1, &f, mask));
}
inline void StateVector::push_force() {
    if (!_owner || !_owner->_built) return;
    std::lock_guard<std::recursive_mutex> g(_owner->_lock);
    float f[4] = {_force.getx(), _force.gety(), _force.getz(), 0.f};
    Physics::ck(_owner->_w, rb_set_force(_owner->_w, _body, 1, f)); Physics::ck(_owner->_w, rb_sync(_owner->_w));
}
inline void StateVector::push_torque() {
    if (!_owner || !_owner->_built) return;
    std::lock_guard<std::recursive_mutex> g(_owner->_lock);
    float t[4] = {_torque.getx(), _torque.gety(), _torque.getz(), 0.f};
    Physics::ck(_owner->_w, rb_set_torque(_owner->_w, _body, 1, t));
}

} // namespace reboot
