// Reference-style host code for the calls that change the world after Physics::addEntities: Physics::addEntity
// (physics/include/Physics.h:49), Entity::setPosition (model/include/Entity.h:50) and Entity::getMVP()/getPrevMVP()
// (Entity.h:66-68), written against include/reboot/Physics.hpp. Prints per sphere entity: position, velocity, the 16 floats
// of getMVP()->getModelBuffer() and of getPrevMVP()->getModelBuffer().
#include "reboot/Physics.hpp"
#include <cstdio>
using namespace reboot;

int main() {
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
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
        Entity* e = new Entity(&ball);
        e->getStateVector()->setLinearPosition(Vector4(-8.0f + 3.1f * i, 1.0f + 0.25f * ((i * 7 + j * 3) % 5), -7.0f + 2.9f * j));
        e->getStateVector()->setActive(true);
        entityList.push_back(e);
    }
    Physics* physics = new Physics();
    physics->addEntities(entityList);
    physics->run();
    for (int k = 0; k < 60; k++) physics->tick(5);
    // a late entity with a scaled, translated world transform
    Model pebble(GeometryType::Sphere);
    pebble.getGeometry()->addSphere(Sphere(0.4f, Vector4(0, 0, 0)));
    Matrix W = Matrix::scale(1.5f); W.getFlatBuffer()[3] = 0.25f;
    Entity* late = new Entity(&pebble, W);
    late->getStateVector()->setLinearPosition(Vector4(0.5f, 4.0f, 0.25f));
    late->getStateVector()->setLinearVelocity(Vector4(0.0f, -0.5f, 0.0f));
    late->getStateVector()->setActive(true);
    physics->addEntity(late);
    entityList.push_back(late);
    for (int k = 0; k < 40; k++) physics->tick(5);
    entityList[1 + 3]->setPosition(Vector4(2.0f, 3.0f, -1.0f));
    late->setPosition(Vector4(-3.0f, 5.0f, 2.0f));
    for (int k = 0; k < 30; k++) physics->tick(5);
    physics->syncState();
    for (size_t i = 1; i < entityList.size(); i++) {
        Entity* e = entityList[i];
        Vector4 p = e->getStateVector()->getLinearPosition(), v = e->getStateVector()->getLinearVelocity();
        printf("%.9g %.9g %.9g %.9g %.9g %.9g", p.getx(), p.gety(), p.getz(), v.getx(), v.gety(), v.getz());
        for (int k = 0; k < 16; k++) printf(" %.9g", e->getMVP()->getModelBuffer()[k]);
        for (int k = 0; k < 16; k++) printf(" %.9g", e->getPrevMVP()->getModelBuffer()[k]);
        printf("\n");
    }
    delete physics;
    return 0;
}
