/* Shim for model/include/Entity.h (real one pulls EventSubscriber/AnimatedModel/IO/shaders).
 * Method bodies live in ref_harness.cpp and restate model/src/Entity.cpp:45-48,131-142,373-395. */
#pragma once
#include "Model.h"
#include "MasterClock.h"

class Entity {
public:
    explicit Entity(Model* model, Matrix worldSpaceTransform);
    void         setPosition(Vector4 position);
    void         setVelocity(Vector4 velocity);
    StateVector* getStateVector() { return &_state; }
    Geometry*    getGeometry() { return &_worldSpaceGeometry; }
    MVP*         getPrevMVP() { return &_prevMVP; }
    MVP*         getMVP() { return &_mvp; }
    Model*       getModel() { return _model; }
    void         _updateKinematics(int milliSeconds);

    Matrix       _worldSpaceTransform;
    Geometry     _worldSpaceGeometry;
    MVP          _prevMVP;
    StateVector  _state;
    Model*       _model;
    MVP          _mvp;
};
