/* Shim: the real engine/include/EngineManager.h needs HINSTANCE / D3D12. It is only *included*
 * (never used) by math/src/Matrix.cpp:4 and math/src/StateVector.cpp:2. */
#pragma once
