/* Shim for drawbuffers VAO (OpenGL). physics/include/GeometryGraphic.h only calls createVAO. */
#pragma once
#include "Cube.h"
enum class GeometryConstruction { LINE_WIREFRAME, TRIANGLE_WIREFRAME, TRIANGLE_MESH };
struct VAO {
    template <class... A> void createVAO(A...) {}
};
