/* Shim for shading/include/ShaderBroker.h: Physics::Physics() only looks up "debugShader". */
#pragma once
#include "Model.h"
#include <string>
class ShaderBroker {
public:
    static ShaderBroker* instance() { static ShaderBroker b; return &b; }
    Shader* getShader(std::string) { static DebugShader d; return &d; }
};
