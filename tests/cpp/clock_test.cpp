// The facade's MasterClock (engine/src/MasterClock.cpp:31-61 semantics) without any GPU: subscribers are
// called in subscription order every 5 ms with the tick length; overruns are counted.
#include "reboot/Physics.hpp"
#include <cstdio>
int main() {
    reboot::MasterClock clock;
    std::vector<int> order; int calls = 0, ms_seen = 0;
    clock.subscribeKinematicsRate([&](int ms) { order.push_back(1); ms_seen = ms; calls++; });
    clock.subscribeKinematicsRate([&](int) { order.push_back(2); });
    clock.reportOverruns(false);
    clock.run();
    std::this_thread::sleep_for(std::chrono::milliseconds(210));
    clock.stop();
    bool alternating = true;
    for (size_t i = 0; i + 1 < order.size(); i += 2) alternating = alternating && order[i] == 1 && order[i + 1] == 2;
    std::printf("ticks %lu calls %d ms %d alternating %d overruns %lu\n", clock.ticks(), calls, ms_seen, (int)alternating, clock.overruns());
    // slow subscriber: every tick overruns and is reported as such
    reboot::MasterClock slow; slow.reportOverruns(false);
    slow.subscribeKinematicsRate([&](int) { std::this_thread::sleep_for(std::chrono::milliseconds(8)); });
    slow.run(); std::this_thread::sleep_for(std::chrono::milliseconds(60)); slow.stop();
    std::printf("slow ticks %lu overruns %lu\n", slow.ticks(), slow.overruns());
    return 0;
}
