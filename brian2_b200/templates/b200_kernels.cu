{# The single CUDA translation unit of a b200 project: the __constant__ array table, every
   in-loop code object as a __device__ function (+ its own kernel for stepwise mode) and one
   persistent cooperative step kernel per run() call. #}
#include "b200_objects.h"
#include "b200_runtime.cuh"
#include "b200_functions.cuh"
#include "b200_connect.cuh"
#include "b200_tiles.cuh"
#include "network.h"
#include "b200_plans.h"
#include <cooperative_groups.h>
#include <chrono>
#include <map>
#include <string>
#include <cmath>
#include <climits>
{% for name in user_headers | sort %}
#include {{name}}
{% endfor %}

__constant__ _B200Arrays _A;

void _b200_sync_constants()
{
    B200_CUDA(cudaMemcpyToSymbol(_A, &_A_host, sizeof(_B200Arrays)));
}

// ---- launch bookkeeping (kernel count for the bench contract, optional per-object timing) ----
static std::map<std::string, std::pair<cudaEvent_t, cudaEvent_t> > _b200_prof_events;
std::map<std::string, double> _b200_prof_seconds;
bool _b200_profiling = {{ 'true' if profiled else 'false' }};
int _b200_ctas_per_sm = {{ctas_per_sm}};
int _b200_grid_override = 0;

void _b200_launch_begin(const char* name)
{
    b200::state().launches++;
    if (_b200_profiling) {
        std::pair<cudaEvent_t, cudaEvent_t>& ev = _b200_prof_events[name];
        if (!ev.first) { B200_CUDA(cudaEventCreate(&ev.first)); B200_CUDA(cudaEventCreate(&ev.second)); }
        B200_CUDA(cudaEventRecord(ev.first, b200::state().stream));
    }
}
void _b200_launch_end(const char* name)
{
    B200_CUDA(cudaGetLastError());
    if (_b200_profiling) {
        std::pair<cudaEvent_t, cudaEvent_t>& ev = _b200_prof_events[name];
        B200_CUDA(cudaEventRecord(ev.second, b200::state().stream));
        B200_CUDA(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        B200_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        _b200_prof_seconds[name] += 1e-3 * ms;
    }
}

// One grid size for every kernel of the project: the thresholder's segments, the element
// partition shared by all per-element code objects and the grid barrier all assume the same
// number of co-resident CTAs.  = min over all kernels of (occupancy-limited CTAs per SM) x SMs.
static std::vector<const void*>& _b200_all_kernels()
{
    static std::vector<const void*> k;
    return k;
}
struct _B200KernelRegistrar { _B200KernelRegistrar(const void* f) { _b200_all_kernels().push_back(f); } };
#define B200_REGISTER_KERNEL(f) static _B200KernelRegistrar _b200_reg_##f((const void*)f);

int _b200_dyn_smem = 0;          // dynamic shared memory of every kernel (tiles of dense counted pathways)
bool _b200_allow_tiles = true;   // prefs.devices.b200.tiled_delivery

int _b200_grid_size()
{
    static int grid = 0, ov = -1, cps = -1, smem = -1;
    if (grid && ov == _b200_grid_override && cps == _b200_ctas_per_sm && smem == _b200_dyn_smem) return grid;
    b200::runtime_init();
    ov = _b200_grid_override; cps = _b200_ctas_per_sm; smem = _b200_dyn_smem;
    int per_sm = _b200_ctas_per_sm;
    for (size_t i = 0; i < _b200_all_kernels().size(); i++) {
        int occ = 0;
        if (_b200_dyn_smem > 0)
            B200_CUDA(cudaFuncSetAttribute(_b200_all_kernels()[i], cudaFuncAttributeMaxDynamicSharedMemorySize, _b200_dyn_smem));
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, _b200_all_kernels()[i], b200::kBlock, _b200_dyn_smem));
        if (occ < 1) throw std::runtime_error("b200: kernel cannot be resident (registers/shared memory)");
        per_sm = std::min(per_sm, occ);
    }
    grid = per_sm * b200::state().num_sms;
    if (_b200_grid_override > 0) grid = std::min(grid, _b200_grid_override);
    grid = std::max(1, std::min(grid, b200::kBlock));
    return grid;
}

// ---- target tiles of dense counted pathways (csrc/b200_tiles.cuh); called by _b200_upload ----
void _b200_tiles_reserve()
{
    bool any = false;
    {% for pw in b200_pathways %}
    {% if pw.tile_n %}
    brian::{{pw.name}}.tile_n = {{pw.tile_n}};
    any = any || b200::tiles_candidate(brian::{{pw.name}}, 2 * b200::state().num_sms);
    {% endif %}
    {% endfor %}
    _b200_dyn_smem = (any && _b200_allow_tiles) ? (int)b200::kTileSmem : 0;
}
void _b200_tiles_build()
{
    const int grid = _b200_grid_size();
    (void)grid;
    {% for pw in b200_pathways %}
    {% if pw.tile_n %}
    if (brian::{{pw.name}}.prepared && (!brian::{{pw.name}}.tiles_tried || brian::{{pw.name}}.tiles_grid != grid)) {
        b200::tiles_build(brian::{{pw.name}}, grid, (size_t)_b200_dyn_smem);
        brian::{{pw.name}}.tiles_tried = true;
        brian::{{pw.name}}.tiles_grid = grid;
    }
    {% endif %}
    {% endfor %}
}

// ---- implicit companions of the thresholders: compaction of an event space (stepwise mode) ----
{% for es in b200_eventspaces %}
__global__ void __launch_bounds__(b200::kBlock, {{ctas_per_sm}})
_kernel_b200_compact{{es.name}}(const _B200Clocks _clks)
{
    const b200::Ctx _ctx{(int)blockIdx.x, (int)gridDim.x, (int)blockIdx.x, (int)gridDim.x, _A._rank, _A._world};
    b200::view_reset();
    b200::compact_segments(_A._es{{es.name}}, _clks.{{es.clock}}.timestep, _ctx, _A._ctrl);
}
B200_REGISTER_KERNEL(_kernel_b200_compact{{es.name}})
void _run_b200_compact{{es.name}}()
{
    _b200_launch_begin("b200_compact{{es.name}}");
    _kernel_b200_compact{{es.name}}<<<_b200_grid_size(), b200::kBlock, 0, b200::state().stream>>>(_b200_clocks_now());
    _b200_launch_end("b200_compact{{es.name}}");
}
{% endfor %}

{% for codeobj in device_code_objects %}
#include "code_objects/{{codeobj.name}}.cuh"
{% endfor %}
{% for name, rep, counted in code_object_aliases %}
void _run_{{name}}() { _run_{{rep}}(); }   // same source as {{rep}} (created by a later run() call)
{% if counted %}
void _run_{{name}}_apply() { _run_{{rep}}_apply(); }
{% endif %}
{% endfor %}

{% for plan in plans %}
{% if plan.clock and plan.alias is not none %}
// run() call #{{plan.index}} has the same schedule as #{{plan.alias}}: shares its kernels
const B200Plan _b200_plan_{{plan.index}} = { _b200_run_chunk_{{plan.alias}}, "{{plan.signature}}" };
{% elif plan.clock %}
// =============================================================================================
// persistent step kernel(s) for run() call #{{plan.index}}
{% for v in plan.variants %}
//   [{{v.tag}}] schedule: {{v.signature}}
//   [{{v.tag}}] grid barriers per step: {{v.n_barriers}}{{ '' if v.end_barrier else ' (no end-of-step barrier)' }}
{% endfor %}
// =============================================================================================
struct _B200Scal_{{plan.index}} {
    {% for item in plan.variants[0].entries if item.kind == 'codeobj' %}
    _co_{{item.name}}::Scal {{item.name}};
    {% endfor %}
    int _unused;
};

{% for v in plan.variants %}
__global__ void __launch_bounds__(b200::kBlock, {{ctas_per_sm}})
_b200_persistent_{{plan.index}}_{{v.tag}}(const _B200Clocks _clks0, const long long _nsteps, const _B200Scal_{{plan.index}} _sc)
{
    const b200::Ctx _ctx{(int)blockIdx.x, (int)gridDim.x, (int)blockIdx.x, (int)gridDim.x, _A._rank, _A._world};
    unsigned long long _bar_target = 0ULL;
    _B200Clocks _clks = _clks0;
    long long _step = 0;
    b200::view_reset();
    {% if profile_phases %}
    long long _pt = clock64();
    // four sampled CTAs (first, 1/4, 3/4, last of the grid): slot k * 512 + phase
    const int _pk = _ctx.gbid == 0 ? 0 : (_ctx.gbid == _ctx.gnb / 4 ? 1 : (_ctx.gbid == (3 * _ctx.gnb) / 4 ? 2 : (_ctx.gbid == _ctx.gnb - 1 ? 3 : -1)));
    #define B200_PHASE(i) if (_pk >= 0 && threadIdx.x == 0) { const long long _now = clock64(); _A._prof[_pk * 512 + (i)] += (unsigned long long)(_now - _pt); _pt = _now; }
    {% else %}
    #define B200_PHASE(i)
    {% endif %}
    while (_step < _nsteps)
    {
        bool _stop = false;
        // the stop flag lives in host memory (one PCIe round trip per poll): look at it every
        // 64 steps only -- a stop request is honoured within a few milliseconds
        if ((_step & 63) == 0 && _ctx.bid == 0 && threadIdx.x == 0 && b200::ld_volatile_s32(_A._stop_request))
            b200::raise_stop(_A._ctrl);
        {% for item in v.entries %}
        {% if item.barrier %}
        _stop |= b200::grid_barrier(&_A._ctrl->barrier, _bar_target, _ctx);
        B200_PHASE({{2 * loop.index0}})
        {% elif item.kind == 'apply' and item.dual %}
        {% elif not loop.first %}
        __syncthreads();
        {% endif %}
        {% if item.share %}
        {   // CTAs [{{item.share[0]}}/{{item.share[2]}}, {{item.share[1]}}/{{item.share[2]}}) of the grid
            int _lo = (int)(((long long)_ctx.gnb * {{item.share[0]}}) / {{item.share[2]}});
            int _hi = (int)(((long long)_ctx.gnb * {{item.share[1]}}) / {{item.share[2]}});
            if (_lo >= _ctx.gnb) _lo = _ctx.gnb - 1;
            if (_hi <= _lo) _hi = _lo + 1;
            if (_ctx.gbid >= _lo && _ctx.gbid < _hi)
            {
                const b200::Ctx _sub{_ctx.gbid - _lo, _hi - _lo, _ctx.gbid, _ctx.gnb, _ctx.rank, _ctx.world};
                {% if item.kind == 'compact' %}
                b200::compact_segments(_A._es{{item.es}}, _clks.{{item.clock}}.timestep, _sub, _A._ctrl);
                {% else %}
                _dev_{{item.name}}(_sub, _clks, _sc.{{item.name}});
                {% endif %}
            }
        }
        {% elif item.kind == 'compact' %}
        b200::compact_segments(_A._es{{item.es}}, _clks.{{item.clock}}.timestep, _ctx, _A._ctrl);
        {% elif item.kind == 'apply' and item.dual %}
        if (_A._pw_{{item.pathway}}.tileptr)       // dense rows only (sparse rows: delivered by reductions)
            _dev_{{item.name}}_apply(_ctx, _clks, _sc.{{item.name}});
        {% elif item.kind == 'apply' %}
        _dev_{{item.name}}_apply(_ctx, _clks, _sc.{{item.name}});
        {% else %}
        _dev_{{item.name}}(_ctx, _clks, _sc.{{item.name}});
        {% endif %}
        B200_PHASE({{2 * loop.index0 + 1}})
        {% endfor %}
        {% if v.end_barrier %}
        _stop |= b200::grid_barrier(&_A._ctrl->barrier, _bar_target, _ctx);
        {% else %}
        // no end-of-step barrier: nothing in the first phase of the next step depends on the
        // last phase of this one across CTAs (B200Device._plan_barriers)
        __syncthreads();
        {% endif %}
        B200_PHASE({{2 * (v.entries | length)}})
        // Clock::tick (brianlib/clocks.h:34-38)
        _clks.{{plan.clock}}.timestep += 1;
        _clks.{{plan.clock}}.t = _clks.{{plan.clock}}.timestep * _clks.{{plan.clock}}.dt;
        ++_step;
        if (_stop) break;
    }
    if (_ctx.bid == 0 && threadIdx.x == 0) _A._ctrl->steps_done = (int)_step;
}
#undef B200_PHASE
B200_REGISTER_KERNEL(_b200_persistent_{{plan.index}}_{{v.tag}})
{% endfor %}

static long long _b200_run_chunk_{{plan.index}}(long long nsteps)
{
    b200::RuntimeState& st = b200::state();
    _B200Scal_{{plan.index}} sc;
    {% for item in plan.variants[0].entries if item.kind == 'codeobj' %}
    _hostscal_{{item.name}}(sc.{{item.name}});
    {% endfor %}
    sc._unused = 0;
    if (nsteps > INT_MAX) nsteps = INT_MAX;
    B200_CUDA(cudaMemsetAsync(st.control, 0, sizeof(b200::Control), st.stream));
    _B200Clocks clks = _b200_clocks_now();
    void* args[] = {(void*)&clks, (void*)&nsteps, (void*)&sc};
    const int grid = _b200_grid_size();
    // 'd1': every pathway delivers at least one step after the spike (decided at upload)
    {% if plan.variants | length > 1 %}
    const void* kernel = st.all_delayed ? (const void*)_b200_persistent_{{plan.index}}_d1 : (const void*)_b200_persistent_{{plan.index}}_d0;
    {% else %}
    const void* kernel = (const void*)_b200_persistent_{{plan.index}}_{{plan.variants[0].tag}};
    {% endif %}
    _b200_launch_begin("persistent_{{plan.index}}");
    B200_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(b200::kBlock), args, _b200_dyn_smem, st.stream));
    _b200_launch_end("persistent_{{plan.index}}");
    B200_CUDA(cudaMemcpyAsync(st.control_host, st.control, sizeof(b200::Control), cudaMemcpyDeviceToHost, st.stream));
    B200_CUDA(cudaStreamSynchronize(st.stream));
    st.poll_cycles += (double)st.control_host->poll_cycles;
    st.fence_cycles += (double)st.control_host->fence_cycles;
    st.polls += (double)st.control_host->polls;
    if (st.control_host->error)
        throw std::runtime_error("b200: a peer GPU did not deliver its spikes in time (multi-GPU run aborted)");
    return (long long)st.control_host->steps_done;
}
const B200Plan _b200_plan_{{plan.index}} = { _b200_run_chunk_{{plan.index}}, "{{plan.signature}}" };
{% else %}
// run() call #{{plan.index}}: stepwise execution (several clocks, profiling, or persistent mode off)
const B200Plan _b200_plan_{{plan.index}} = { 0, "stepwise" };
{% endif %}
{% endfor %}
