{# Device-side object table of a b200 project: the array table `_A` (device pointers, passed to
   kernels through __constant__ memory), spike rings, pathways, monitor buffers, and the
   upload / download that brackets Network::run.  The host mirrors themselves (namespace brian)
   are the reference's own objects.cpp (templates/objects.cpp:151-176). #}
{% macro h_file() %}
#ifndef _B200_OBJECTS_H
#define _B200_OBJECTS_H
#include "b200_host.h"
#include "objects.h"

struct _B200Clocks {
    {% for clock in clocks | sort(attribute='name') %}
    b200::Clk {{clock.name}};
    {% endfor %}
    int _unused;
};

struct _B200Arrays {
    unsigned long long _seed;
    b200::Control* _ctrl;
    int* _stop_request;
    {% for a in b200_arrays %}
    {{a.ctype}}* {{a.name}};
    {% if a.kind != 'static' %}
    size_t _n{{a.name}};
    size_t _cap{{a.name}};
    {% endif %}
    {% endfor %}
    int _rank, _world;
    unsigned long long* _prof;     // per-phase cycle counters of the persistent kernel (optional)
    {% for es in b200_eventspaces %}
    b200::EventSpaceDev _es{{es.name}};
    {% endfor %}
    {% for pw in b200_pathways %}
    b200::PathwayDev _pw_{{pw.name}};
    {% endfor %}
    {% for mon in b200_monitors %}
    long long* _monN_{{mon.name}};
    {% endfor %}
    {% for sv in b200_summed %}
    b200::TargetIndexDev _sv_{{sv}};
    {% endfor %}
};

extern _B200Arrays _A_host;
{% for es in b200_eventspaces %}
extern b200::EventSpace _b200_es{{es.name}};
{% endfor %}
namespace brian {
{% for pw in b200_pathways %}
extern b200::Pathway {{pw.name}};
{% endfor %}
}
{% for sv in b200_summed %}
extern b200::TargetIndex _b200_sv_{{sv}};
{% endfor %}

{% for es in b200_eventspaces %}
void _run_b200_compact{{es.name}}();   // implicit companion of the thresholder (stepwise mode)
{% endfor %}
void _b200_upload();
void _b200_download();
void _b200_sync_constants();                 // defined next to the kernels (owns __constant__ _A)
int _b200_grid_size();                       // CTAs of every kernel of this project (co-resident)
extern int _b200_dyn_smem;                   // dynamic shared memory of every kernel launch
void _b200_tiles_reserve();                  // dense counted pathways: shared-memory budget ...
void _b200_tiles_build();                    // ... and tile tables (csrc/b200_tiles.cuh)
_B200Clocks _b200_clocks_now();
void _b200_prepare_steps(long long steps, bool exact, bool count = true);   // make monitor buffers large enough
void _b200_prefault_start(long long steps);   // host-side: map the pages the next records will land in
void _b200_prefault_join();
void _b200_launch_begin(const char* name);
void _b200_launch_end(const char* name);
unsigned long long _b200_events_delivered();
void _b200_write_profiling();
// host-mirror access by name for the C ABI ("_array_neurongroup_v" or "neurongroup.v")
long long _b200_array_nbytes(const char* name);
int _b200_array_copy_out(const char* name, void* out, size_t nbytes);
int _b200_array_copy_in(const char* name, const void* data, size_t nbytes);
#endif
{% endmacro %}


{% macro cpp_file() %}
#include "b200_objects.h"
#include <thread>
#include "brianlib/clocks.h"
#include <chrono>
#include <map>

_B200Arrays _A_host;
{% for es in b200_eventspaces %}
b200::EventSpace _b200_es{{es.name}};
{% endfor %}
namespace brian {
{% for pw in b200_pathways %}
b200::Pathway {{pw.name}}({{pw.sources}}, {{pw.start}}, {{pw.stop}}, {{pw.hits_n}});
{% endfor %}
}
{% for sv in b200_summed %}
b200::TargetIndex _b200_sv_{{sv}};
{% endfor %}

// per-monitor host bookkeeping: upper bound of the number of recorded entries
{% for mon in b200_monitors %}
static long long _monN_ub_{{mon.name}} = 0;
{% endfor %}
static bool _b200_first_upload = true;
static unsigned long long _b200_uploaded_epoch = 0;   // host_epoch of the last upload (0: none yet)
bool _b200_allow_d1 = true;      // prefs.devices.b200.elide_end_barrier
// monitor records: number of leading elements identical on host and device (see upload_records)
{% for a in b200_arrays %}
{% if a.used and ((a.kind == 'dynamic1d' and a.monitor) or a.kind == 'dynamic2d') %}
static size_t _b200_synced{{a.name}} = 0;
{% endif %}
{% endfor %}

_B200Clocks _b200_clocks_now()
{
    _B200Clocks c;
    {% for clock in clocks | sort(attribute='name') %}
    c.{{clock.name}}.t = brian::{{array_specs[clock.variables['t']]}}[0];
    {% if clock.__class__.__name__ == "EventClock" %}
    c.{{clock.name}}.dt = 0.0;
    {% else %}
    c.{{clock.name}}.dt = brian::{{array_specs[clock.variables['dt']]}}[0];
    {% endif %}
    c.{{clock.name}}.timestep = brian::{{array_specs[clock.variables['timestep']]}}[0];
    {% endfor %}
    c._unused = 0;
    return c;
}

void _b200_upload()
{
    _b200_prefault_join();
    using namespace brian;
    b200::runtime_init();
    b200::RuntimeState& st = b200::state();
    const auto _t0 = std::chrono::high_resolution_clock::now();
    if (_b200_first_upload) {
        memset(&_A_host, 0, sizeof(_A_host));
        _b200_first_upload = false;
    }
    b200::ensure_seed();
    _A_host._seed = st.seed;
    _A_host._ctrl = st.control;
    _A_host._stop_request = st.stop_request_dev;
    const _B200Clocks _now = _b200_clocks_now();
    (void)_now;
    // Event monitors record into device append buffers.  Growing one (allocate, copy, free) costs
    // hundreds of milliseconds once the buffers are large, so they start with a share of the
    // device memory that is free at this point: 1/8 of it, split over all record buffers.
    static size_t _b200_record_entries = 0;
    if (_b200_record_entries == 0) {
        size_t _free = 0, _total = 0, _bytes = 0;
        {% for a in b200_arrays %}
        {% if a.used and a.kind == 'dynamic1d' and a.monitor %}
        _bytes += sizeof({{a.ctype}});
        {% endif %}
        {% endfor %}
        if (_bytes && cudaMemGetInfo(&_free, &_total) == cudaSuccess)
            _b200_record_entries = std::min<size_t>(_free / 8 / _bytes, (size_t)1 << 30);
        _b200_record_entries = std::max<size_t>(_b200_record_entries, 1);
    }
    // After a run the written arrays are downloaded, so host mirrors and device arrays agree;
    // as long as nothing on the host touched an array since (host_epoch, bumped by every
    // host-side write of the generated main()), the next run needs no upload at all.
    const bool _b200_in_sync = _b200_uploaded_epoch != 0 && _b200_uploaded_epoch == st.host_epoch;
    _b200_uploaded_epoch = st.host_epoch;
    {% for a in b200_arrays %}
    {% if a.used %}
    {% if a.kind == 'static' %}
    if (!_b200_in_sync || !_A_host.{{a.name}})
        b200::upload_array(_A_host.{{a.name}}, brian::{{a.name}}, {{a.size}});
    {% elif a.kind == 'dynamic1d' and a.monitor %}
    b200::upload_records(_A_host.{{a.name}}, _A_host._cap{{a.name}}, _A_host._n{{a.name}}, brian::{{a.dyn_name}}, std::max<size_t>((size_t){{a.min_cap}}, _b200_record_entries), _b200_synced{{a.name}});
    {% elif a.kind == 'dynamic1d' %}
    if (!_b200_in_sync || !_A_host.{{a.name}} || _A_host._n{{a.name}} != brian::{{a.dyn_name}}.size())
        b200::upload_vector(_A_host.{{a.name}}, _A_host._cap{{a.name}}, _A_host._n{{a.name}}, brian::{{a.dyn_name}}, (size_t){{a.min_cap}});
    {% elif a.kind == 'dynamic2d' %}
    if (!_b200_in_sync || !_A_host.{{a.name}} || _A_host._n{{a.name}} != brian::{{a.dyn_name}}.n)
    {
        // row-major (rows x {{a.width}}) append buffer
        const size_t _rows = brian::{{a.dyn_name}}.n, _w = {{a.width}};
        const size_t _need = std::max<size_t>(_rows, (size_t){{a.min_cap}});
        if (!_A_host.{{a.name}} || _A_host._cap{{a.name}} < _need) {
            b200::dev_free(_A_host.{{a.name}});
            _A_host._cap{{a.name}} = std::max<size_t>(_need, 16);
            _A_host.{{a.name}} = ({{a.ctype}}*)b200::dev_alloc(_A_host._cap{{a.name}} * _w * sizeof({{a.ctype}}));
        }
        _A_host._n{{a.name}} = _rows;
        if (_rows && brian::{{a.dyn_name}}.m == _w) {
            std::vector<{{a.ctype}}> _tmp(_rows * _w);
            for (size_t i = 0; i < _rows; i++) for (size_t j = 0; j < _w; j++) _tmp[i*_w + j] = brian::{{a.dyn_name}}(i, j);
            B200_CUDA(cudaMemcpy(_A_host.{{a.name}}, _tmp.data(), _tmp.size()*sizeof({{a.ctype}}), cudaMemcpyHostToDevice));
        }
    }
    {% endif %}
    {% endif %}
    {% endfor %}
    // event spaces (history survives between runs); on several GPUs the rings are mapped into
    // every peer once per allocation (CUDA IPC handles travel through the allgather callback)
    if (!_A_host._prof) {
        _A_host._prof = (unsigned long long*)b200::dev_alloc(4 * 512 * sizeof(unsigned long long));
        B200_CUDA(cudaMemset(_A_host._prof, 0, 4 * 512 * sizeof(unsigned long long)));
    }
    _A_host._rank = st.rank;
    _A_host._world = st.world;
    {   // does every pathway deliver at least one step after the spike?  (-> 'd1' step kernels)
        bool _any = false, _all = _b200_allow_d1;
        {% for pw in b200_pathways %}
        if (brian::{{pw.name}}.prepared) {
            _any = true;
            if (brian::{{pw.name}}.bin_delay.empty() || brian::{{pw.name}}.bin_delay.front() < 1) _all = false;
            brian::{{pw.name}}.ensure_hits({{pw.hits_n}});
        }
        {% endfor %}
        {% for es in b200_eventspaces %}
        {% if es.compact_always %}
        _all = false;   // order-dependent synaptic code walks the current step's compacted list
        {% endif %}
        {% endfor %}
        st.all_delayed = _any && _all;
    }
    _b200_tiles_reserve();      // (before the first _b200_grid_size(): occupancy depends on it)
    _b200_tiles_build();
    {% for es in b200_eventspaces %}
    _b200_es{{es.name}}.id = {{loop.index}};
    _b200_es{{es.name}}.compact_always = {{ 'true' if es.compact_always else 'false' }};
    _b200_es{{es.name}}.ensure({{es.size - 1}}, _b200_grid_size(), _now.{{es.clock}}.timestep);
    {% endfor %}
    {% for es in b200_eventspaces %}
    _b200_es{{es.name}}.open_peers();
    _A_host._es{{es.name}} = _b200_es{{es.name}}.view();
    {% endfor %}
    {% for pw in b200_pathways %}
    if (brian::{{pw.name}}.prepared)
        _A_host._pw_{{pw.name}} = brian::{{pw.name}}.view();
    {% endfor %}
    {% for sv in b200_summed %}
    if (_b200_sv_{{sv}}.prepared)
        _A_host._sv_{{sv}} = _b200_sv_{{sv}}.view();
    {% endfor %}
    {% for mon in b200_monitors %}
    {
        if (!_A_host._monN_{{mon.name}}) _A_host._monN_{{mon.name}} = (long long*)b200::dev_alloc(2 * sizeof(long long));
        const long long _n = (long long)brian::{{mon.N_array}}[0];
        const long long _two[2] = {_n, _n};
        B200_CUDA(cudaMemcpy(_A_host._monN_{{mon.name}}, _two, sizeof(_two), cudaMemcpyHostToDevice));
        _monN_ub_{{mon.name}} = _n;
    }
    {% endfor %}
    _b200_sync_constants();
    B200_CUDA(cudaDeviceSynchronize());
    st.upload_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - _t0).count();
}

// Guarantee that every monitor can record `steps` more steps (exact == true: the true entry
// counts are read back first; false: cheap host-side upper bounds are used).
static long long _b200_steps_prepared = 0;   // steps of the earlier launches (event-rate estimate)
void _b200_prepare_steps(long long steps, bool exact, bool count)
{
    bool changed = false;
    {% for mon in b200_monitors %}
    {
        if (exact) {
            long long _two[2];
            B200_CUDA(cudaMemcpy(_two, _A_host._monN_{{mon.name}}, sizeof(_two), cudaMemcpyDeviceToHost));
            _monN_ub_{{mon.name}} = std::max(_two[0], _two[1]);
        }
        {% if mon.kind == 'spike' %}
        // worst case one entry per source neuron per step; the kernel raises `overflow` when
        // fewer than one step's worst case of slots is free
        long long _need = _monN_ub_{{mon.name}} + (long long){{mon.headroom}} * (exact ? 3 : 1);
        // a launch of `steps` steps at the event rate seen so far (x1.5) should fit without
        // interrupting the persistent kernel for a buffer growth
        if (exact && _b200_steps_prepared > 0 && steps > 1)
            _need = std::max(_need, _monN_ub_{{mon.name}} + (long long){{mon.headroom}} * 3 +
                                    (long long)(1.5 * (double)_monN_ub_{{mon.name}} / (double)_b200_steps_prepared * (double)steps));
        {% else %}
        long long _need = _monN_ub_{{mon.name}} + steps;
        {% endif %}
        {% for b in mon.buffers %}
        if ((long long)_A_host._cap{{b.name}} < _need) {
            if (!exact) {   // re-check with the true count before growing
                long long _two[2];
                B200_CUDA(cudaMemcpy(_two, _A_host._monN_{{mon.name}}, sizeof(_two), cudaMemcpyDeviceToHost));
                _monN_ub_{{mon.name}} = std::max(_two[0], _two[1]);
                {% if mon.kind == 'spike' %}
                _need = _monN_ub_{{mon.name}} + (long long){{mon.headroom}} * 3;
                {% else %}
                _need = _monN_ub_{{mon.name}} + steps;
                {% endif %}
            }
            if ((long long)_A_host._cap{{b.name}} < _need) {
                const size_t _newcap = std::max<size_t>((size_t)_need, 2 * _A_host._cap{{b.name}} + 16);
                size_t _cap_rows = _A_host._cap{{b.name}};
                {% if b.ndim == 2 %}
                {
                    {{b.ctype}}* _nd = ({{b.ctype}}*)b200::dev_alloc(_newcap * {{b.width}} * sizeof({{b.ctype}}));
                    if (_A_host.{{b.name}} && _monN_ub_{{mon.name}} > 0)
                        B200_CUDA(cudaMemcpy(_nd, _A_host.{{b.name}}, (size_t)_monN_ub_{{mon.name}} * {{b.width}} * sizeof({{b.ctype}}), cudaMemcpyDeviceToDevice));
                    b200::dev_free(_A_host.{{b.name}});
                    _A_host.{{b.name}} = _nd;
                    _cap_rows = _newcap;
                }
                {% else %}
                b200::grow_buffer(_A_host.{{b.name}}, _cap_rows, (size_t)_monN_ub_{{mon.name}}, _newcap);
                {% endif %}
                _A_host._cap{{b.name}} = _cap_rows;
                changed = true;
            }
        }
        {% endfor %}
        {% if mon.kind == 'spike' %}
        if (!exact) _monN_ub_{{mon.name}} += (long long){{mon.headroom}} * steps;
        {% else %}
        if (!exact) _monN_ub_{{mon.name}} += steps;
        {% endif %}
    }
    {% endfor %}
    if (exact && count) _b200_steps_prepared += steps;
    if (changed) _b200_sync_constants();
}

// While the persistent kernel runs the host has nothing to do: a helper thread reserves the
// capacity the monitors' host vectors will need after the launch (event rate seen so far x1.5)
// and touches those pages, so that the download at the end of run() copies into mapped memory
// instead of paying a first-touch page fault per 4 KB (that was 3/4 of the end-to-end overhead of
// a COBAHH run).  Nothing else touches these vectors between _b200_prefault_start and
// _b200_prefault_join (called by _b200_download / _b200_upload).
static struct _B200PrefaultThread : std::thread {
    using std::thread::operator=;
    ~_B200PrefaultThread() { if (joinable()) join(); }   // never leave a joinable thread behind
} _b200_prefault_thread;
template <typename T>
static void _b200_prefault_vector(std::vector<T>& v, size_t need)
{
    if (v.capacity() < need) v.reserve(std::max(need, 2 * v.capacity()));
    volatile char* p = (volatile char*)v.data();
    const size_t from = (v.size() * sizeof(T)) & ~(size_t)4095, to = v.capacity() * sizeof(T);
    for (size_t off = from + 4096; off < to; off += 4096) p[off] = 0;
}
void _b200_prefault_join()
{
    if (_b200_prefault_thread.joinable()) _b200_prefault_thread.join();
}
void _b200_prefault_start(long long steps)
{
    _b200_prefault_join();
    if (getenv("B200_NO_PREFAULT")) return;
    {% for mon in b200_monitors %}
    {% if mon.kind == 'spike' %}
    const size_t _need_{{mon.name}} = (size_t)(_monN_ub_{{mon.name}} + (_b200_steps_prepared > steps
        ? (long long)(1.5 * (double)_monN_ub_{{mon.name}} / (double)(_b200_steps_prepared - steps) * (double)steps) : 0));
    {% else %}
    const size_t _need_{{mon.name}} = (size_t)(_monN_ub_{{mon.name}} + steps);
    {% endif %}
    {% endfor %}
    _b200_prefault_thread = std::thread([=]() {
        {% for a in b200_arrays %}
        {% if a.used and a.kind == 'dynamic1d' and a.monitor %}
        _b200_prefault_vector(brian::{{a.dyn_name}}, _need_{{a.monitor}});
        {% endif %}
        {% endfor %}
    });
}

void _b200_download()
{
    using namespace brian;
    b200::RuntimeState& st = b200::state();
    _b200_prefault_join();
    const auto _t0 = std::chrono::high_resolution_clock::now();
    B200_CUDA(cudaDeviceSynchronize());
    {% for mon in b200_monitors %}
    long long _monN_{{mon.name}} = 0;
    {
        long long _two[2];
        B200_CUDA(cudaMemcpy(_two, _A_host._monN_{{mon.name}}, sizeof(_two), cudaMemcpyDeviceToHost));
        _monN_{{mon.name}} = std::max(_two[0], _two[1]);
    }
    {% endfor %}
    {% for a in b200_arrays %}
    {% if a.used and a.written %}
    {% if a.eventspace %}
    {
        // host mirror of an event space = the list of the last executed step
        _b200_es{{a.name}}.download_step(brian::{{a.name}}, _b200_clocks_now().{{a.clock}}.timestep - 1);
    }
    {% elif a.kind == 'static' %}
    b200::download_array(brian::{{a.name}}, _A_host.{{a.name}}, {{a.size}});
    {% elif a.kind == 'dynamic1d' %}
    {% if a.monitor %}
    b200::download_records(brian::{{a.dyn_name}}, _A_host.{{a.name}}, (size_t)_monN_{{a.monitor}}, _b200_synced{{a.name}});
    _A_host._n{{a.name}} = (size_t)_monN_{{a.monitor}};
    {% else %}
    b200::download_vector(brian::{{a.dyn_name}}, _A_host.{{a.name}}, _A_host._n{{a.name}});
    {% endif %}
    {% elif a.kind == 'dynamic2d' %}
    {
        // rows recorded by earlier runs are already on the host (append-only records)
        const size_t _rows = (size_t)_monN_{{a.monitor}}, _w = {{a.width}};
        size_t _from = (brian::{{a.dyn_name}}.n == _b200_synced{{a.name}} && brian::{{a.dyn_name}}.m == _w
                        && _rows >= _b200_synced{{a.name}}) ? _b200_synced{{a.name}} : 0;
        std::vector<{{a.ctype}}> _tmp((_rows - _from) * _w);
        if (_rows > _from) B200_CUDA(cudaMemcpy(_tmp.data(), _A_host.{{a.name}} + _from * _w, _tmp.size()*sizeof({{a.ctype}}), cudaMemcpyDeviceToHost));
        brian::{{a.dyn_name}}.resize(_rows, _w);
        for (size_t i = _from; i < _rows; i++) for (size_t j = 0; j < _w; j++) brian::{{a.dyn_name}}(i, j) = _tmp[(i - _from)*_w + j];
        b200::state().d2h_bytes += _tmp.size()*sizeof({{a.ctype}});
        _b200_synced{{a.name}} = _rows;
    }
    {% endif %}
    {% endif %}
    {% endfor %}
    st.download_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - _t0).count();
}

// ---- host mirrors by name (C ABI: b200_get_array / b200_set_array) ----------------------------
#include <functional>
struct _B200HostArray {
    const char* name;
    const char* user_name;
    std::function<size_t()> nbytes;
    std::function<void(void*)> copy_out;
    std::function<void(const void*, size_t)> copy_in;
};
static std::vector<_B200HostArray>& _b200_host_arrays()
{
    static std::vector<_B200HostArray> table;
    if (!table.empty()) return table;
    {% for a in b200_arrays %}
    {% if a.kind == 'static' %}
    table.push_back({"{{a.name}}", "{{a.user_name}}",
        []() -> size_t { return (size_t){{a.size}} * sizeof({{a.ctype}}); },
        [](void* out) { memcpy(out, brian::{{a.name}}, (size_t){{a.size}} * sizeof({{a.ctype}})); },
        [](const void* in, size_t n) { memcpy(brian::{{a.name}}, in, std::min(n, (size_t){{a.size}} * sizeof({{a.ctype}}))); }});
    {% elif a.kind == 'dynamic1d' %}
    table.push_back({"{{a.dyn_name}}", "{{a.user_name}}",
        []() -> size_t { return brian::{{a.dyn_name}}.size() * sizeof({{a.ctype}}); },
        [](void* out) { if (!brian::{{a.dyn_name}}.empty()) memcpy(out, &brian::{{a.dyn_name}}[0], brian::{{a.dyn_name}}.size() * sizeof({{a.ctype}})); },
        [](const void* in, size_t n) { brian::{{a.dyn_name}}.resize(n / sizeof({{a.ctype}})); if (n) memcpy(&brian::{{a.dyn_name}}[0], in, n); }});
    {% elif a.kind == 'dynamic2d' %}
    table.push_back({"{{a.dyn_name}}", "{{a.user_name}}",
        []() -> size_t { return brian::{{a.dyn_name}}.n * brian::{{a.dyn_name}}.m * sizeof({{a.ctype}}); },
        [](void* out) { {{a.ctype}}* o = ({{a.ctype}}*)out; const size_t n = brian::{{a.dyn_name}}.n, m = brian::{{a.dyn_name}}.m;
                        for (size_t i = 0; i < n; i++) for (size_t j = 0; j < m; j++) o[i*m + j] = brian::{{a.dyn_name}}(i, j); },
        [](const void*, size_t) { throw std::runtime_error("b200: cannot set a 2-d monitor array"); }});
    {% endif %}
    {% endfor %}
    return table;
}
static _B200HostArray* _b200_find_array(const char* name)
{
    std::vector<_B200HostArray>& t = _b200_host_arrays();
    for (size_t i = 0; i < t.size(); i++)
        if (!strcmp(t[i].name, name) || !strcmp(t[i].user_name, name)) return &t[i];
    return 0;
}
long long _b200_array_nbytes(const char* name)
{
    _B200HostArray* a = _b200_find_array(name);
    return a ? (long long)a->nbytes() : -1;
}
int _b200_array_copy_out(const char* name, void* out, size_t nbytes)
{
    _B200HostArray* a = _b200_find_array(name);
    if (!a || a->nbytes() > nbytes) return 1;
    a->copy_out(out);
    return 0;
}
int _b200_array_copy_in(const char* name, const void* data, size_t nbytes)
{
    _B200HostArray* a = _b200_find_array(name);
    if (!a) return 1;
    try { a->copy_in(data, nbytes); } catch (...) { return 2; }
    b200::state().host_epoch++;       // the host mirror changed: upload it before the next run
    return 0;
}

unsigned long long _b200_events_delivered()
{
    unsigned long long total = 0;
    {% for pw in b200_pathways %}
    total += brian::{{pw.name}}.events_delivered();
    {% endfor %}
    return total;
}
{% endmacro %}
