{# Entry points of a b200 project.  The reference emits `int main()` and is run as a subprocess
   (templates/main.cpp:47-75, device.py:1311); here the same sequence of main_lines is exposed
   through the C ABI declared in include/brian2_b200.h and called in-process via ctypes. #}
#include <stdlib.h>
#include "objects.h"
#include "b200_objects.h"
#include "b200_plans.h"
#include <csignal>
#include <ctime>
#include <time.h>
#include "run.h"
#include "brianlib/common_math.h"
#include "brian2_b200.h"

{% for codeobj in code_objects | sort(attribute='name') %}
#include "code_objects/{{codeobj.name}}.h"
{% for block in codeobj.before_after_blocks %}
#include "code_objects/{{block}}_{{codeobj.name}}.h"
{% endfor %}
{% endfor %}

{% for name in user_headers | sort %}
#include {{name}}
{% endfor %}

#include <iostream>
#include <fstream>
#include <string>
#include <map>

{{report_func|autoindent}}

extern std::map<std::string, double> _b200_prof_seconds;
extern bool _b200_profiling;
extern int _b200_ctas_per_sm;
extern int _b200_grid_override;
extern int _b200_dyn_smem;
extern bool _b200_allow_d1;
extern bool _b200_allow_tiles;

static std::string _b200_error;

#include <execinfo.h>
#include <unistd.h>
static void _b200_crash_handler(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "b200: fatal signal inside the project library, backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1)) {}
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

void set_from_command_line(const std::vector<std::string> args)
{
    if (!args.empty()) b200::state().host_epoch++;   // host arrays may change: CSRs are rebuilt
    for (const auto& arg : args) {
        size_t equal_sign = arg.find("=");
        auto name = arg.substr(0, equal_sign);
        auto value = arg.substr(equal_sign + 1, arg.length());
        brian::set_variable_by_name(name, value);
    }
}

static void _b200_main(std::vector<std::string> args)
{
    std::random_device _rd;
    if (args.size() >= 2 && args[0] == "--results_dir")
    {
        brian::results_dir = args[1];
        args.erase(args.begin(), args.begin()+2);
    }
    {{'\n'.join(code_lines['before_start'])|autoindent}}
    brian_start();
    {{'\n'.join(code_lines['after_start'])|autoindent}}
    {
        using namespace brian;
        {{main_lines|autoindent}}
    }
    {{'\n'.join(code_lines['before_end'])|autoindent}}
    _b200_write_profiling();
    _write_arrays();
    {{'\n'.join(code_lines['after_end'])|autoindent}}
}

void _b200_write_profiling()
{
    if (!_b200_profiling) return;
    std::ofstream f(brian::results_dir + "profiling_info.txt");
    {% for codeobj in profiled_codeobjects | sort %}
    f << "{{codeobj}}\t" << _b200_prof_seconds["{{codeobj}}"] << std::endl;
    {% endfor %}
}

extern "C" {

int b200_run_main(int argc, const char** argv)
{
    if (getenv("B200_DEBUG")) {
        std::signal(SIGSEGV, _b200_crash_handler);
        std::signal(SIGBUS, _b200_crash_handler);
        std::signal(SIGFPE, _b200_crash_handler);
    }
    try {
        _b200_error.clear();
        std::vector<std::string> args;
        for (int i = 0; i < argc; i++) args.push_back(argv[i]);
        _b200_main(args);
        return 0;
    } catch (const std::exception& e) {
        _b200_error = e.what();
        return 1;
    } catch (...) {
        _b200_error = "unknown C++ exception";
        return 1;
    }
}

const char* b200_last_error() { return _b200_error.c_str(); }
double b200_last_run_time() { return Network::_last_run_time; }
double b200_last_run_completed_fraction() { return Network::_last_run_completed_fraction; }
void b200_request_stop() { if (b200::state().stop_request) *b200::state().stop_request = 1; Network::_globally_stopped = true; }

int b200_set_comm(int rank, int world, b200_allgather_fn allgather)
{
    if (world < 1 || world > b200::kMaxRanks || rank < 0 || rank >= world) return 1;
    if (world > 1 && !allgather) return 1;
    if (b200::state().initialised) return 2;   // must be called before the first run
    b200::state().rank = rank;
    b200::state().world = world;
    b200::state().allgather = allgather;
    return 0;
}
int b200_comm_rank() { return b200::state().rank; }
int b200_comm_world() { return b200::state().world; }

int b200_set_option(const char* key, double value)
{
    const std::string k(key);
    if (k == "mode") Network::_b200_mode = (int)value;
    else if (k == "max_chunk") Network::_b200_max_chunk = (long long)value;
    else if (k == "profile") _b200_profiling = value != 0;
    else if (k == "ctas_per_sm") _b200_ctas_per_sm = (int)value;
    else if (k == "grid") _b200_grid_override = (int)value;
    else if (k == "allow_d1") _b200_allow_d1 = value != 0;
    else if (k == "tiles") _b200_allow_tiles = value != 0;
    else if (k == "forward") b200::Pathway::allow_forward() = value != 0;
    else if (k == "seed") { b200::state().seed = (unsigned long long)value; b200::state().seeded = true; }
    else return 1;
    return 0;
}

double b200_get_counter(const char* key)
{
    const std::string k(key);
    b200::RuntimeState& st = b200::state();
    if (k == "launches") return (double)st.launches;
    if (k == "events") return (double)_b200_events_delivered();
    if (k == "steps") return (double)Network::_b200_steps_run;
    if (k == "h2d_bytes") return (double)st.h2d_bytes;
    if (k == "d2h_bytes") return (double)st.d2h_bytes;
    if (k == "upload_seconds") return st.upload_seconds;
    if (k == "download_seconds") return st.download_seconds;
    if (k == "device_bytes") return (double)st.bytes_allocated;
    if (k == "num_sms") return (double)st.num_sms;
    if (k == "poll_cycles") return st.poll_cycles;
    if (k == "fence_cycles") return st.fence_cycles;
    if (k == "polls") return st.polls;
    if (k == "all_delayed") return st.all_delayed ? 1.0 : 0.0;
    if (k == "dyn_smem") return (double)_b200_dyn_smem;
    if (k == "connect_seconds") return st.connect_seconds;
    if (k == "connect_synapses") return st.connect_synapses;
    if (k == "connect_launches") return (double)st.connect_launches;
    if (k == "prepare_seconds") return st.prepare_seconds;
    if (k == "grid") return st.initialised ? (double)_b200_grid_size() : 0.0;
    if (k == "runs") return (double)Network::_b200_run_log.size();
    if (k.compare(0, 5, "phase") == 0) {   // "phase<i>": cycles CTA 0 spent in phase i (profile_phases)
        const int i = atoi(k.c_str() + 5);
        unsigned long long v = 0;
        if (i < 0 || i >= 4 * 512 || !_A_host._prof) return -1.0;
        cudaMemcpy(&v, _A_host._prof + i, sizeof(v), cudaMemcpyDeviceToHost);
        return (double)v;
    }
    // per-run records: "run<i>.<field>"
    if (k.compare(0, 3, "run") == 0) {
        const size_t dot = k.find('.');
        if (dot != std::string::npos) {
            const size_t i = (size_t)atoi(k.substr(3, dot - 3).c_str());
            const std::string f = k.substr(dot + 1);
            if (i < Network::_b200_run_log.size()) {
                const B200RunRecord& r = Network::_b200_run_log[i];
                if (f == "device_seconds") return r.device_seconds;
                if (f == "wall_seconds") return r.wall_seconds;
                if (f == "t0_unix") return r.t0_unix;
                if (f == "upload_seconds") return r.upload_seconds;
                if (f == "download_seconds") return r.download_seconds;
                if (f == "prepare_seconds") return r.prepare_seconds;
                if (f == "events") return r.events;
                if (f == "steps") return (double)r.steps;
                if (f == "persistent") return (double)r.persistent;
            }
        }
    }
    return -1.0;
}

int b200_profiling(const char** names, double* seconds, int cap)
{
    int n = 0;
    for (std::map<std::string, double>::iterator it = _b200_prof_seconds.begin(); it != _b200_prof_seconds.end() && n < cap; ++it, ++n) {
        names[n] = it->first.c_str();
        seconds[n] = it->second;
    }
    return n;
}

long long b200_get_array_size(const char* name) { return _b200_array_nbytes(name); }
int b200_get_array(const char* name, void* out, size_t nbytes) { return _b200_array_copy_out(name, out, nbytes); }
int b200_set_array(const char* name, const void* data, size_t nbytes) { return _b200_array_copy_in(name, data, nbytes); }

int b200_finalize()
{
    try {
        _dealloc_arrays();
        cudaDeviceSynchronize();
        return 0;
    } catch (...) { return 1; }
}

}  // extern "C"
