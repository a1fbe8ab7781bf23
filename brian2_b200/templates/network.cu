{# Time loop of a b200 project.  Reference: brian2/devices/cpp_standalone/templates/network.cpp
   (run loop :38-122, clock selection :134-159).  Two execution modes:
     * persistent: all code objects of the run share one regular clock -> the whole step loop
       runs inside one cooperative kernel (generated per run() call, see b200_kernels.cu),
       launched in chunks so that stop requests, progress reports, the run-time limit and
       monitor-buffer growth are serviced by the host between chunks;
     * stepwise: one kernel launch per code object per step, host picks the clocks exactly as
       the reference does (multi-clock networks, profiling). #}
{% macro cpp_file() %}
#include "network.h"
#include "b200_objects.h"
#include <stdlib.h>
#include <math.h>
#include <iostream>
#include <chrono>
#include <utility>

#define Clock_epsilon 1e-14

double Network::_last_run_time = 0.0;
double Network::_last_run_completed_fraction = 0.0;
bool Network::_globally_stopped = false;
bool Network::_globally_running = false;
int Network::_b200_mode = 0;
long long Network::_b200_max_chunk = 20000;
long long Network::_b200_steps_run = 0;
std::vector<B200RunRecord> Network::_b200_run_log;

Network::Network() { t = 0.0; }

void Network::clear() { objects.clear(); }

void Network::add(BaseClock* clock, codeobj_func func)
{
    objects.push_back(std::make_pair(clock, func));
}

void Network::compute_clocks()
{
    clocks.clear();
    for (size_t i = 0; i < objects.size(); i++)
        clocks.insert(objects[i].first);
}

// the clock(s) with the smallest t run next; clocks within Clock_epsilon tick together
BaseClock* Network::next_clocks()
{
    if (clocks.empty())
        return NULL;
    BaseClock* minclock = *clocks.begin();
    for (std::set<BaseClock*>::iterator i = clocks.begin(); i != clocks.end(); i++)
        if ((*i)->t[0] < minclock->t[0])
            minclock = *i;
    curclocks.clear();
    const double tmin = minclock->t[0];
    for (std::set<BaseClock*>::iterator i = clocks.begin(); i != clocks.end(); i++)
    {
        const double s = (*i)->t[0];
        if (s == tmin || fabs(s - tmin) <= Clock_epsilon)
            curclocks.insert(*i);
    }
    return minclock;
}

void Network::run(const double duration, void (*report_func)(const double, const double, const double, const double),
                  const double report_period, const B200Plan* plan)
{
    typedef std::chrono::high_resolution_clock hrc;
    const double t_start = t;
    const double t_end = t + duration;
    double next_report_time = report_period;
    compute_clocks();
    for (std::set<BaseClock*>::iterator i = clocks.begin(); i != clocks.end(); i++)
        (*i)->set_interval(t, t_end);

    // host mirrors -> device (not part of the timed loop, like the reference's _load_arrays)
    const double _upload_before = b200::state().upload_seconds;
    _b200_upload();

    // Monitor buffers that must grow for the first launch grow now: allocation and copy of
    // device buffers are host-side housekeeping like the upload (and are charged to it), not
    // part of the step loop.
    {
        BaseClock* first = next_clocks();
        const auto _g0 = std::chrono::high_resolution_clock::now();
        if (plan && plan->run_chunk && clocks.size() == 1 && first && first->regular() && Network::_b200_mode == 0)
            _b200_prepare_steps(std::min<long long>((report_func != NULL) ? 200 : Network::_b200_max_chunk,
                                                    first->steps_left()), true, false);
        b200::state().upload_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - _g0).count();
    }

    // device-side clock of the loop: CUDA events on the launching stream
    static cudaEvent_t _ev_start = 0, _ev_stop = 0;
    if (!_ev_start) { B200_CUDA(cudaEventCreate(&_ev_start)); B200_CUDA(cudaEventCreate(&_ev_stop)); }
    const unsigned long long _events_before = _b200_events_delivered();
    const long long _steps_before = Network::_b200_steps_run;
    B200_CUDA(cudaDeviceSynchronize());
    // several GPUs: the ranks enter the loop together (their uploads take different times, and a
    // rank whose peer starts late would spin on the peer's first spike list inside the timed loop)
    b200::host_barrier();
    B200_CUDA(cudaEventRecord(_ev_start, b200::state().stream));
    hrc::time_point start = hrc::now(), current;
    const double _t0_unix = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
    if (report_func)
        report_func(0.0, 0.0, t_start, duration);

    BaseClock* clock = next_clocks();
    double elapsed_realtime = 0.0;
    bool did_break_early = false;
    {% if maximum_run_time is not none %}
    const bool has_time_limit = true;
    const double time_limit = {{maximum_run_time}};
    {% else %}
    const bool has_time_limit = false;
    const double time_limit = 0.0;
    {% endif %}
    const bool should_check_time = has_time_limit || (report_func != NULL);

    Network::_globally_running = true;
    Network::_globally_stopped = false;
    *b200::state().stop_request = 0;

    const bool persistent = plan && plan->run_chunk && clocks.size() == 1 && clock &&
                            clock->regular() && Network::_b200_mode == 0;
    if (persistent)
    {
        long long chunk = should_check_time ? 200 : Network::_b200_max_chunk;
        while (clock->running() && !Network::_globally_stopped)
        {
            t = clock->t[0];
            const long long want = std::min<long long>(chunk, clock->steps_left());
            _b200_prepare_steps(want, true);
            _b200_prefault_start(want);
            const hrc::time_point c0 = hrc::now();
            const long long done = plan->run_chunk(want);
            clock->advance(done);
            Network::_b200_steps_run += done;
            current = hrc::now();
            elapsed_realtime = std::chrono::duration<double>(current - start).count();
            if (should_check_time)
            {
                // aim for ~0.25 s of wall time per launch
                const double per_step = std::chrono::duration<double>(current - c0).count() / std::max<long long>(done, 1);
                chunk = std::max<long long>(1, std::min<long long>(Network::_b200_max_chunk, (long long)(0.25 / std::max(per_step, 1e-9))));
                if (has_time_limit && elapsed_realtime > time_limit)
                {
                    did_break_early = true;
                    break;
                }
                if (report_func && elapsed_realtime > next_report_time)
                {
                    report_func(elapsed_realtime, (clock->t[0] - t_start) / duration, t_start, duration);
                    next_report_time += report_period;
                }
            }
            if (*b200::state().stop_request)
                Network::_globally_stopped = true;
        }
    }
    else
    {
        while (clock && clock->running() && !Network::_globally_stopped)
        {
            t = clock->t[0];
            if (should_check_time)
            {
                current = hrc::now();
                elapsed_realtime = std::chrono::duration<double>(current - start).count();
                if (has_time_limit && elapsed_realtime > time_limit)
                {
                    did_break_early = true;
                    break;
                }
                if (report_func && elapsed_realtime > next_report_time)
                {
                    report_func(elapsed_realtime, (t - t_start) / duration, t_start, duration);
                    next_report_time += report_period;
                }
            }
            _b200_prepare_steps(1, false);
            for (size_t i = 0; i < objects.size(); i++)
            {
                if (curclocks.find(objects[i].first) != curclocks.end() && objects[i].second)
                    objects[i].second();
            }
            for (std::set<BaseClock*>::iterator i = curclocks.begin(); i != curclocks.end(); i++)
                (*i)->tick();
            Network::_b200_steps_run += 1;
            clock = next_clocks();
            if (*b200::state().stop_request)
                Network::_globally_stopped = true;
        }
    }
    B200_CUDA(cudaEventRecord(_ev_stop, b200::state().stream));
    B200_CUDA(cudaStreamSynchronize(b200::state().stream));
    B200_CUDA(cudaDeviceSynchronize());
    // several GPUs: nobody touches its rings (download, next upload) while a peer may still be
    // storing spikes into them
    b200::host_barrier();
    Network::_globally_running = false;
    current = hrc::now();
    elapsed_realtime = std::chrono::duration<double>(current - start).count();

    if (!did_break_early && !Network::_globally_stopped)
        t = t_end;
    else
        t = clock ? clock->t[0] : t_end;

    _last_run_time = elapsed_realtime;
    if (duration > 0)
        _last_run_completed_fraction = (t - t_start) / duration;
    else
        _last_run_completed_fraction = 1.0;

    // device -> host mirrors (written arrays only)
    const double _download_before = b200::state().download_seconds;
    _b200_download();
    {
        B200RunRecord rec;
        float ms = 0.f;
        B200_CUDA(cudaEventElapsedTime(&ms, _ev_start, _ev_stop));
        rec.device_seconds = 1e-3 * ms;
        rec.wall_seconds = elapsed_realtime;
        rec.t0_unix = _t0_unix;
        rec.steps = Network::_b200_steps_run - _steps_before;
        rec.events = (double)(_b200_events_delivered() - _events_before);
        rec.upload_seconds = b200::state().upload_seconds - _upload_before;
        rec.download_seconds = b200::state().download_seconds - _download_before;
        static double _prepare_seen = 0.0;
        rec.prepare_seconds = b200::state().prepare_seconds - _prepare_seen;
        _prepare_seen = b200::state().prepare_seconds;
        rec.persistent = persistent ? 1 : 0;
        Network::_b200_run_log.push_back(rec);
    }

    if (report_func)
        report_func(elapsed_realtime, _last_run_completed_fraction, t_start, duration);
}
{% endmacro %}

{% macro h_file() %}
#ifndef _BRIAN_NETWORK_H
#define _BRIAN_NETWORK_H
#include <vector>
#include <utility>
#include <set>
#include "brianlib/clocks.h"

typedef void (*codeobj_func)();

// A pre-generated persistent step kernel for one run() call of the script
struct B200Plan {
    long long (*run_chunk)(long long steps);   // returns the number of steps executed
    const char* signature;
};

// one record per Network::run call (bench.py reads them through b200_get_counter)
struct B200RunRecord {
    double device_seconds;     // CUDA-event time of the step loop
    double wall_seconds;       // host clock around the same region
    double t0_unix;            // host time (seconds since the epoch) when the loop started
    double upload_seconds, download_seconds;
    double prepare_seconds;    // host time spent (re)building pathway CSRs since the previous run
    double events;             // synaptic events delivered during this run
    long long steps;
    int persistent;
};

class Network
{
    std::set<BaseClock*> clocks, curclocks;
    void compute_clocks();
    BaseClock* next_clocks();
public:
    std::vector< std::pair< BaseClock*, codeobj_func > > objects;
    double t;
    static double _last_run_time;
    static double _last_run_completed_fraction;
    static bool _globally_stopped;
    static bool _globally_running;
    static int _b200_mode;              // 0: persistent kernel when possible, 1: always stepwise
    static long long _b200_max_chunk;   // steps per persistent launch
    static long long _b200_steps_run;
    static std::vector<B200RunRecord> _b200_run_log;

    Network();
    void clear();
    void add(BaseClock *clock, codeobj_func func);
    void run(const double duration, void (*report_func)(const double, const double, const double, const double),
             const double report_period, const B200Plan* plan = 0);
};
#endif
{% endmacro %}
