/* brian2_b200.h -- C ABI of a built `b200` project (the shared library that replaces the
 * `./main` executable of Brian2's cpp_standalone device).
 *
 * Every generated project exports exactly these symbols; Python binds them with ctypes
 * (brian2_b200/capi.py).  Plain pointers and sizes only -- no C++ / torch types.
 *
 * Reference interfaces replaced (brian-team/brian2, paths relative to brian2/):
 *   b200_run_main ................ `int main(argc, argv)` of the generated project,
 *                                   devices/cpp_standalone/templates/main.cpp:47-75, started with
 *                                   subprocess.call(["./main", "--results_dir", dir] + run_args),
 *                                   devices/cpp_standalone/device.py:1311.  Arguments have the same
 *                                   meaning ("--results_dir <dir>", then "group.var=value|file").
 *   b200_last_run_time /
 *   b200_last_run_completed_fraction  results/last_run_info.txt, templates/objects.cpp:364-374,
 *                                   read back at device.py:1326-1332 (Network::_last_run_time,
 *                                   templates/network.cpp:14-15,114-118).
 *   b200_get_array(_size) ........ results/<array>_<crc32> raw dumps, templates/objects.cpp:270-323,
 *                                   read by CPPStandaloneDevice.get_value, device.py:544-580.
 *   b200_set_array ............... static_arrays/<name> + `name=file` run arguments,
 *                                   templates/objects.cpp:91-150 (set_variable_by_name), device.py:1114.
 *   b200_profiling ............... results/profiling_info.txt, templates/objects.cpp:349-363,
 *                                   device.py:1991-2002.
 *   b200_request_stop ............ SIGINT handler / Network::_globally_stopped,
 *                                   templates/main.cpp:38-51, templates/network.cpp:16-17,65-67.
 *   b200_last_error .............. non-zero exit status + stderr of ./main, device.py:1316-1324.
 *   b200_set_option / b200_get_counter   no reference counterpart (device preferences and the
 *                                   counters bench.py reports: kernel launches, synaptic events,
 *                                   host<->device bytes).
 *
 * Differences from the sketch in SURVEY.md section 8(b): there is no `b200_init(n_gpus,
 * static_dir)` -- the runtime initialises itself on the first device call of b200_run_main (the
 * static arrays are found relative to the working directory exactly like ./main finds them) and
 * the multi-GPU setup is b200_set_comm; b200_get_array takes the capacity of the caller's buffer
 * by value and fails if it is too small (the size is asked with b200_get_array_size first).
 *
 * Conventions: functions returning int return 0 on success, non-zero on failure with the
 * message available from b200_last_error().  The library never calls exit().  Buffers passed in
 * or out are owned by the caller (copy semantics).  Not re-entrant: one run at a time.
 */
#ifndef BRIAN2_B200_H
#define BRIAN2_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Run the whole generated main(): host-side initialisation code objects (reference C++,
 * bit-identical RNG stream), upload, the device time loop(s), download, result files. */
int b200_run_main(int argc, const char** argv);

/* Message of the last failure ("" if none).  Valid until the next call. */
const char* b200_last_error(void);

/* Wall-clock seconds of the last Network::run time loop (uploads/downloads excluded, exactly
 * like the reference excludes file I/O) and the completed fraction of its duration. */
double b200_last_run_time(void);
double b200_last_run_completed_fraction(void);

/* Ask a running simulation to stop after the current step (callable from another thread). */
void b200_request_stop(void);

/* Multi-GPU (one process per GPU of one box).  Must be called before b200_run_main.  The
 * network is partitioned by postsynaptic neuron: rank r owns a contiguous block of every group
 * (its state, thresholding, reset, monitors) and every synapse whose postsynaptic neuron lies
 * in the block.  Inside the step loop the ranks exchange their spike lists by direct NVLink
 * stores into each other's spike rings (CUDA IPC mapped peer memory) -- no host involvement.
 * `allgather` is only used outside the loop (IPC handle exchange, end-of-run barrier): it must
 * gather `nbytes_per_rank` bytes from every rank into `recv` (rank order) and return 0.
 * Reference counterpart: none (brian2 has no multi-process mode). */
typedef int (*b200_allgather_fn)(const void* send, void* recv, size_t nbytes_per_rank);
int b200_set_comm(int rank, int world, b200_allgather_fn allgather);
int b200_comm_rank(void);
int b200_comm_world(void);

/* Options, to be set before b200_run_main:
 *   "mode"        0 = persistent step kernel when possible (default), 1 = one launch per code object
 *   "max_chunk"   steps per persistent launch
 *   "profile"     1 = per-code-object CUDA-event timing (forces mode 1 semantics per launch)
 *   "ctas_per_sm" resident CTAs per SM used to size grids
 *   "grid"        upper bound on the number of CTAs (0 = none)
 *   "seed"        seed of the device RNG streams (default: drawn from std::random_device unless the
 *                 script called seed(); rank 0's draw on several GPUs)
 *   "allow_d1"    1 = use the step kernel without end-of-step barrier when every pathway delivers
 *                 at least one step after the spike (default), 0 = never
 *   "forward"     1 = forward delivery of counted pathways with delays >= 1 step (default), 0 = off
 *   "tiles"       1 = count dense rows of countable pathways in shared memory over target tiles
 *                 (default), 0 = always scatter */
int b200_set_option(const char* key, double value);

/* Counters: "launches", "events" (delivered synaptic events), "steps", "h2d_bytes",
 * "d2h_bytes", "upload_seconds", "download_seconds", "device_bytes", "num_sms", "grid", "runs",
 * "connect_seconds|connect_synapses|connect_launches" (synapse creation on the device),
 * "prepare_seconds" (host time building pathway CSRs), "all_delayed", "dyn_smem", and per
 * Network::run call "run<i>.device_seconds|wall_seconds|t0_unix|upload_seconds|download_seconds|
 * prepare_seconds|events|steps|persistent"; "phase<k*512+i>" = cycles sampled CTA k (first, 1/4, 3/4, last of the
 * grid) spent in phase i of the persistent kernel (profile_phases builds).  -1 if unknown. */
double b200_get_counter(const char* key);

/* Per-code-object device seconds (profile mode).  Fills up to `cap` entries, returns the count.
 * The name pointers stay valid until b200_finalize. */
int b200_profiling(const char** names, double* seconds, int cap);

/* Host mirrors by name: either the array name ("_array_neurongroup_v",
 * "_dynamic_array_spikemonitor_t") or "owner.variable" ("neurongroup.v").
 * b200_get_array_size returns the size in bytes or -1. */
long long b200_get_array_size(const char* name);
int b200_get_array(const char* name, void* out, size_t nbytes);
int b200_set_array(const char* name, const void* data, size_t nbytes);

/* Free host mirrors and device memory. */
int b200_finalize(void);

#ifdef __cplusplus
}
#endif
#endif /* BRIAN2_B200_H */
