// tools.cpp — see tools.hpp.
#include "tools.hpp"

#include <time.h>

uint64_t get_time_since_boot_microseconds()
{
    struct timespec ts;
#ifdef CLOCK_BOOTTIME
    const clockid_t clock = CLOCK_BOOTTIME;
#else
    const clockid_t clock = CLOCK_MONOTONIC;
#endif
    if (clock_gettime(clock, &ts) != 0) return 0;
    return (uint64_t)ts.tv_sec * 1000000u + (uint64_t)ts.tv_nsec / 1000u;
}
