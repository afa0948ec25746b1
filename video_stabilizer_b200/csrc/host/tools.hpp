// tools.hpp — the reference's timing helper (tools.hpp:5), used by align_test.cpp:211,227.
#pragma once

#include <cstdint>

// microseconds since boot (CLOCK_BOOTTIME: includes time spent suspended)
uint64_t get_time_since_boot_microseconds();
