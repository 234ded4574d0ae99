// Reference-style host code against the B200 path: the three lines the reference has commented out
// at engine/src/EngineManager.cpp:222-225 plus the visualize() reader of :468-470, written against
// include/reboot/Physics.hpp. Prints the final state so the pytest can compare it with the C ABI.
#include "reboot/Physics.hpp"
#include <cstdio>
#include <cstdlib>
using namespace reboot;

int main(int argc, char** argv) {
    int nside = argc > 1 ? atoi(argv[1]) : 6, nticks = argc > 2 ? atoi(argv[2]) : 120;
    // static terrain: a flat 16x16 quad grid at y = 0 with the reference's triangulation (ProcIsland.cpp:61-67)
    Model terrain(GeometryType::Triangle);
    const int G = 16; const float c = 2.0f, half = G * c / 2;
    for (int z = 0; z < G; z++) for (int x = 0; x < G; x++) {
        Vector4 b(x * c - half, 0, z * c - half), b1((x + 1) * c - half, 0, z * c - half), bs(x * c - half, 0, (z + 1) * c - half), bs1((x + 1) * c - half, 0, (z + 1) * c - half);
        terrain.getGeometry()->addTriangle(Triangle(bs, b1, b));
        terrain.getGeometry()->addTriangle(Triangle(bs, bs1, b1));
    }
    Model ball(GeometryType::Sphere);
    ball.getGeometry()->addSphere(Sphere(0.5f, Vector4(0, 0, 0)));
    std::vector<Entity*> entityList;
    entityList.push_back(new Entity(&terrain));
    for (int i = 0; i < nside; i++) for (int j = 0; j < nside; j++) {
        Entity* e = new Entity(&ball);
        e->getStateVector()->setLinearPosition(Vector4(-10.0f + 3.1f * i, 1.0f + 0.25f * ((i * 7 + j * 3) % 5), -9.0f + 2.9f * j));
        e->getStateVector()->setActive(true);
        entityList.push_back(e);
    }
    Physics* physics = new Physics();
    physics->addEntities(entityList);
    physics->run();
    size_t hits = 0;
    for (int k = 0; k < nticks; k++) {
        physics->tick(5);
        physics->visualize([&](Entity*, const std::set<std::pair<Entity*, uint32_t>>& tris) { hits += tris.size(); });
    }
    // a force only sticks while in contact (StateVector.cpp:147-156)
    physics->syncState();
    entityList[1]->getStateVector()->setForce(Vector4(3.0f, 0, 0));
    for (int k = 0; k < 10; k++) physics->tick(5);
    physics->syncState();
    for (size_t i = 1; i < entityList.size(); i++) {
        StateVector* s = entityList[i]->getStateVector();
        Vector4 p = s->getLinearPosition(), v = s->getLinearVelocity();
        printf("%zu %.9g %.9g %.9g %.9g %.9g %.9g %d\n", i - 1, p.getx(), p.gety(), p.getz(), v.getx(), v.gety(), v.getz(), (int)s->getContact());
    }
    printf("hits %zu\n", hits);
    // callers either side of the step: culling over the world's static cells (identity inverse view-projection = the
    // NDC cube itself: the four cells around the origin) and collider ingest (FbxLoader::_buildGeometryData)
    std::vector<uint32_t> vis = physics->getVisibleFrustumCulling(Matrix());
    Model collider(GeometryType::Sphere), prop(GeometryType::Triangle);
    const std::vector<float> mesh = { -1, 0, 0,  3, 0.5f, 0,  0, 2, 1,  1, -1, -2 };
    const std::vector<uint32_t> mi = { 0, 1, 2, 0, 2, 3 };
    const float trs[9] = { 10, 0, 0, 0, 0, 0, 2, 2, 2 };
    buildGeometryData(&collider, mesh, mi, trs);
    buildGeometryData(&prop, mesh, mi, trs);
    const Sphere& cs = (*collider.getGeometry()->getSpheres())[0];
    fprintf(stderr, "extras: visible %zu sphere r %.6g cx %.6g cy %.6g tris %zu\n", vis.size(), cs.getObjectRadius(), cs.getObjectPosition().getx(),
            cs.getObjectPosition().gety(), prop.getGeometry()->getTriangles()->size());
    delete physics;
    return 0;
}
