/* dc_timer.cc -- DC_get_time / DC_get_cycles / DC_timer (reference: src/tools.cc:37-111).
 * Not on the data path; kept so that programs using the reference's timers link. */
#include <sys/time.h>
#include "DC.h"

double DC_get_time ()
{
    struct timeval now;
    gettimeofday(&now, nullptr);
    return (double)now.tv_sec + 1.e-6 * (double)now.tv_usec;
}

uint64_t DC_get_cycles ()
{
#if defined(__x86_64__) || defined(__i386__)
    uint32_t lo, hi;
    __asm__ volatile("rdtsc" : "=a"(lo), "=d"(hi));
    return ((uint64_t)hi << 32) | lo;
#else
    return (uint64_t)(DC_get_time() * 1e9);
#endif
}

DC_timer::DC_timer () : avgTime(0), timeCtr(0), avgCycles(0), cyclesCtr(0) {}

double DC_timer::get_avg_time () { return avgTime; }
void DC_timer::reset_time () { avgTime = 0; timeCtr = 0; }
void DC_timer::start_time () { startTime = DC_get_time(); }
void DC_timer::stop_time ()
{
    double elapsed = DC_get_time() - startTime;
    avgTime = (timeCtr * avgTime + elapsed) / (timeCtr + 1);
    timeCtr++;
}

uint64_t DC_timer::get_avg_cycles () { return avgCycles; }
void DC_timer::reset_cycles () { avgCycles = 0; cyclesCtr = 0; }
void DC_timer::start_cycles () { startCycles = DC_get_cycles(); }
void DC_timer::stop_cycles ()
{
    uint64_t elapsed = DC_get_cycles() - startCycles;
    avgCycles = (cyclesCtr * avgCycles + elapsed) / (cyclesCtr + 1);
    cyclesCtr++;
}
