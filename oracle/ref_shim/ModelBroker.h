#pragma once
#include "ViewEventDistributor.h"
struct ModelBroker {
    static ViewEventDistributor* getViewManager() { static ViewEventDistributor v; return &v; }
};
