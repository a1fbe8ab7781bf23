{# Host clocks of a b200 project (written to brianlib/clocks.h of the generated project).  Same
   arithmetic as brian2/devices/cpp_standalone/brianlib/clocks.h:16-77 -- t = timestep*dt,
   epsilon-tolerant interval rounding -- but with the end step exposed, because the persistent
   kernel is launched for a known number of steps. #}
#ifndef _BRIAN_CLOCKS_H
#define _BRIAN_CLOCKS_H
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>

class BaseClock
{
public:
    int64_t *timestep;
    double *t;
    int64_t i_end;
    BaseClock() : timestep(0), t(0), i_end(0) {}
    virtual ~BaseClock() {}
    virtual void tick() = 0;
    virtual void advance(int64_t steps) = 0;
    virtual void set_interval(double start, double end) = 0;
    virtual bool regular() const = 0;
    inline bool running() { return timestep[0] < i_end; }
    inline int64_t steps_left() { return i_end - timestep[0]; }
};

class Clock : public BaseClock
{
    static int64_t nearest(double x) { return (int64_t)(x + 0.5); }
public:
    double epsilon;
    double *dt;
    Clock(double _epsilon=1e-14) : epsilon(_epsilon), dt(0) {}
    void tick() { advance(1); }
    void advance(int64_t steps)
    {
        timestep[0] += steps;
        t[0] = timestep[0] * dt[0];
    }
    bool regular() const { return true; }
    // first step at or after `start`, first step at or after `end` (both within epsilon)
    int64_t step_at_or_after(double x)
    {
        const int64_t i = nearest(x / dt[0]);
        const double ti = i * dt[0];
        if (ti == x || fabs(ti - x) <= epsilon * fabs(ti))
            return i;
        return (int64_t)ceil(x / dt[0]);
    }
    void set_interval(double start, double end)
    {
        timestep[0] = step_at_or_after(start);
        i_end = step_at_or_after(end);
    }
};

class EventClock : public BaseClock
{
public:
    double *times;
    size_t n_times;
    EventClock() : times(0), n_times(0) {}
    void tick() { advance(1); }
    void advance(int64_t steps)
    {
        timestep[0] += steps;
        t[0] = times[timestep[0]];
    }
    bool regular() const { return false; }
    void set_interval(double start, double end)
    {
        timestep[0] = std::lower_bound(times, times + n_times, start) - times;
        t[0] = times[timestep[0]];
        i_end = std::lower_bound(times, times + n_times, end) - times;
    }
};
#endif
