/* Shim for model/include/Model.h (real one pulls GL, FBX SDK, shaders, textures).
 * Provides exactly what physics/ and math/GeometryMath use. */
#pragma once
#include "Matrix.h"
#include "StateVector.h"
#include "Geometry.h"
#include "MVP.h"
#include "VAO.h"
#include <set>
#include <map>
#include <vector>
#include <string>
#include <mutex>

class Entity;

struct RenderBuffers {
    std::vector<Vector4> _v;
    std::vector<Vector4>* getVertices() { return &_v; }
};

struct Shader { virtual ~Shader() {} };

/* Contact-capture hook: Physics::visualize() (physics/src/Physics.cpp:51-79) hands
 * _triangleIntersectionList[entity] by value to runShader, exactly the reference's query channel. */
struct DebugShader : Shader {
    std::function<void(MVP*, std::set<Triangle*>&)> sink;
    void runShader(MVP* mvp, VAO*, std::set<Triangle*> tris, float*, GeometryConstruction) {
        if (sink) sink(mvp, tris);
    }
};

class Model {
public:
    GeometryType   _type;
    Geometry       _geometry;
    RenderBuffers  _rb;
    explicit Model(GeometryType t) : _type(t) {}
    GeometryType   getGeometryType() { return _type; }
    Geometry*      getGeometry() { return &_geometry; }
    RenderBuffers* getRenderBuffers() { return &_rb; }
};
