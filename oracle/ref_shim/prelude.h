/* Forced-include prelude for building the reference's physics/ + math/ TUs headless on Linux.
 * TEST INFRASTRUCTURE ONLY (oracle). Fixes two MSVC-isms without touching reference sources:
 *  - sqrtf undeclared in math/src/Vector4.cpp:35
 *  - unqualified `vector` in physics/include/OctTree.h:60 */
#pragma once
#include <cmath>
#include <math.h>
#include <algorithm>
#include <vector>
#include <iostream>
#include <string>
#include <set>
#include <map>
#include <unordered_map>
#include <mutex>
#include <functional>
using std::vector;
