/* Shim: only OSP::getVisibleFrustumCulling uses the view manager. The harness sets the two matrices
 * (ref_visible_frustum_culling) before calling the unmodified reference function. */
#pragma once
#include "Matrix.h"
struct ViewEventDistributor {
    Matrix view, proj;
    Matrix getFrustumView() { return view; }
    Matrix getFrustumProjection() { return proj; }
};
