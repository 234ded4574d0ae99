/* Shim: only OSP::getVisibleFrustumCulling (out of scope) uses the view manager. */
#pragma once
#include "Matrix.h"
struct ViewEventDistributor {
    Matrix getFrustumView() { return Matrix(); }
    Matrix getFrustumProjection() { return Matrix(); }
};
