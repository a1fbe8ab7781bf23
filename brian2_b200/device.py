"""The ``b200`` simulation device: ``set_device('b200')`` beside ``runtime``/``cpp_standalone``.

Implements Brian2's ``Device`` plugin surface (``brian2/devices/device.py:86``) by deriving from
``CPPStandaloneDevice`` (``brian2/devices/cpp_standalone/device.py:144``), as the reference's
developer documentation recommends for new standalone back-ends
(``docs_sphinx/developer/devices.rst:22-26``).  What is inherited: array bookkeeping and the array
cache, ``main_queue``, run-argument handling, result files.  What is new:

* run-once code objects stay reference host C++ (bit-identical initialisation / connectivity),
  in-loop code objects become sm_100a ``__device__`` functions (`B200CodeObject`);
* the project is built into ONE shared library with ``nvcc`` and driven in-process through the
  C ABI of ``include/brian2_b200.h`` (ctypes) instead of ``subprocess.call(['./main'])``;
* every ``run()`` call gets a *persistent step kernel*: the schedule of code objects is known at
  build time, so the device computes where grid barriers are really needed from the code
  objects' read/write sets (``_plan_barriers``) and emits a cooperative kernel that executes the
  whole time loop on the GPU.
"""
import os
import shutil
import sys
from collections import defaultdict

import numpy as np

from brian2.codegen.generators.cpp_generator import c_data_type
from brian2.core.namespace import get_local_namespace
from brian2.core.preferences import BrianPreference, prefs
from brian2.core.variables import ArrayVariable, Constant, DynamicArrayVariable
from brian2.devices.cpp_standalone.codeobject import CPPStandaloneCodeObject
from brian2.devices.cpp_standalone.device import CPPStandaloneDevice
from brian2.devices.device import all_devices
from brian2.parsing.rendering import CPPNodeRenderer
from brian2.units import second
from brian2.utils.logger import get_logger

from .capi import B200Library
from .codeobject import (
    DEVICE_TEMPLATES,
    HOST_TEMPLATES,
    RUN_ONCE_DEVICE_TEMPLATES,
    B200CodeObject,
    B200ConnectCodeObject,
    B200HostCodeObject,
    B200ShardHostCodeObject,
)
from .cuda_generator import clock_field, is_eventspace

__all__ = ["B200Device", "b200_device"]

#: templates that write a step's spike list into the spike ring (segments of the owning CTAs)
SPIKE_SOURCE_TEMPLATES = ("threshold", "spikegenerator")

logger = get_logger("brian2.devices.b200")

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PACKAGE_DIR, "csrc")
INCLUDE_DIR = os.path.join(os.path.dirname(PACKAGE_DIR), "include")

prefs.register_preferences(
    "devices.b200",
    "B200 (sm_100a) standalone device preferences",
    nvcc=BrianPreference(
        default=os.environ.get("B200_NVCC", "/usr/local/cuda/bin/nvcc"),
        docs="The nvcc binary used to compile the device code.",
    ),
    cuda_home=BrianPreference(
        default=os.environ.get("CUDA_HOME", "/usr/local/cuda"), docs="CUDA toolkit root."
    ),
    arch_flags=BrianPreference(
        default="-gencode arch=compute_100a,code=sm_100a",
        docs="Target architecture flags passed to nvcc (Blackwell B200 only).",
    ),
    fmad=BrianPreference(
        default=False,
        docs="""
        Allow nvcc to contract a*b+c into fused multiply-adds in device code.  Off by default so
        that results are bit-comparable with a C++ build compiled with ``-ffp-contract=off``.
        """,
    ),
    extra_nvcc_flags=BrianPreference(
        default=["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-w"],
        docs="Additional flags for nvcc.",
    ),
    persistent=BrianPreference(
        default=True,
        docs="""
        Run the whole time loop inside one persistent cooperative kernel whenever all code objects
        of a run share one regular clock.  If False, every code object is launched as its own
        kernel every step (also used when profiling).
        """,
    ),
    max_chunk=BrianPreference(
        default=20000, docs="Maximum number of time steps per persistent-kernel launch."
    ),
    ctas_per_sm=BrianPreference(
        default=2, docs="Resident CTAs (512 threads each) per SM used to size all grids."
    ),
    multi_gpu=BrianPreference(
        default=True,
        docs="""
        Shard the network over all ranks of the initialised ``torch.distributed`` process group
        (one process per GPU, partition by postsynaptic neuron, spike lists exchanged by NVLink
        peer stores).  If False -- or if there is no process group -- every process simulates the
        whole network on its own GPU.
        """,
    ),
    fuse_exp_pow=BrianPreference(
        default=True,
        docs="""
        Evaluate ``exp(a)**c`` (|c| <= 1) in device code as one exponential of the double-double
        product ``a*c`` instead of ``pow(exp(a), c)``.  Both are within ~0.5 ulp of the true value
        (the reference's glibc result is, too); the fused form costs a third.
        """,
    ),
    libm=BrianPreference(
        default="cuda",
        validator=lambda v: v in ("cuda", "glibc"),
        docs="""
        Arithmetic of ``exp``, ``expm1`` (``exprel``), ``log``, ``pow``, ``tanh``, ``sinh``, ``cosh``,
        ``sin`` and ``cos`` in double-precision device code.
        ``'cuda'``: CUDA's algorithms (<= 1 ulp from the host's glibc; state variables of a
        Hodgkin-Huxley network agree with ``cpp_standalone`` to rtol 1e-9, spikes are identical).
        ``'glibc'``: the algorithms of the host's glibc, operation by operation, with the lookup
        tables read from the host's libm (`brian2_b200.libm_tables`, csrc/b200_glibc_math.cuh):
        results of these functions -- and with them the state variables -- are bit-identical to a
        ``cpp_standalone`` build compiled with ``-ffp-contract=off``.  Disables ``fuse_exp_pow``.
        """,
    ),
    cse=BrianPreference(
        default=True,
        docs="""
        Evaluate repeated stateless function calls / powers of the vector code once (value
        numbering on the abstract code, see `brian2_b200.cuda_generator._CallCSE`).  Bit-neutral.
        """,
    ),
    split_phases=BrianPreference(
        default=True,
        docs="""
        Inside the persistent kernel, run the mutually independent code objects of one phase
        (monitors, synaptic pathways, compaction) side by side on disjoint sets of CTAs instead
        of one after the other on all CTAs.
        """,
    ),
    profile_phases=BrianPreference(
        default=False,
        docs="""
        Instrument the persistent kernel: CTA 0 accumulates the SM cycles it spends in every code
        object and at every grid barrier (read back with ``device.phase_profile()``).
        """,
    ),
    grid=BrianPreference(
        default=0, docs="Upper bound on the number of CTAs of every kernel (0: no bound)."
    ),
    forward_delivery=BrianPreference(
        default=True,
        docs="""
        Counted pathways whose delays are all at least one time step (and at most 32 distinct
        values): store the synapses by (source, delay) and deliver every spike ONCE, one step
        after it was emitted, into per-target counters of the step in which each synapse is due
        (a ring of ``max_delay + 1`` counter arrays), instead of once per delay value from a spike
        list of the right age.
        """,
    ),
    tiled_delivery=BrianPreference(
        default=True,
        docs="""
        Counted pathways whose rows are dense (on average at least 8 synapses of a row point into
        the block of targets of every CTA): cut the rows at the CTAs' block boundaries when the
        CSR is built and let the owner of every block count its events in shared memory
        (csrc/b200_tiles.cuh) instead of scattering integer reductions through L2.
        """,
    ),
    elide_end_barrier=BrianPreference(
        default=True,
        docs="""
        When every synaptic pathway of the project delivers at least one time step after the
        spike, run the variant of the persistent kernel whose deliveries overlap the state
        update of the same step and that needs no grid barrier at the end of a step.
        """,
    ),
    counted_pathways=BrianPreference(
        default="auto",
        docs="""
        Synaptic code that only touches data of the element at the non-source end of a synapse
        (``v_post += J``, with or without ``(unless refractory)``; ``v_post += c*(E - v_post)``)
        can be executed by COUNTING the events per target (integer reductions, no floating-point
        atomics, no gather of target-side state per event) and letting the owner of every target
        apply the statements that many times in the reference's order: bit-identical to the
        sequential reference whatever the constants, several pathways may deliver into one
        variable inside the same phase of a step, and dense rows are accumulated in shared memory.
        ``'auto'``: code that reads target-side data is counted; plain ``x_post += constant``
        stays a floating-point reduction while its rows are sparse and switches to the counted
        owner-computes form when they are dense (decided when the CSR is built);
        ``'always'``: counted whenever the code qualifies; ``'never'``.
        """,
        validator=lambda v: v in ("always", "auto", "never"),
    ),
    construction=BrianPreference(
        default="reference",
        docs="""
        How ``Synapses.connect`` and the initialisation of synaptic variables are executed.

        ``'reference'``: the reference's own host C++ (one sequential mt19937 stream): connectivity,
        delays and weights are bit-identical to ``cpp_standalone`` for the same ``seed()``; on
        several GPUs every rank builds the whole network and keeps its share.

        ``'sharded'``: ``connect()`` generator expressions run as CUDA kernels (one Philox stream
        per source row, csrc/b200_connect.cuh) and every rank creates ONLY the synapses whose
        postsynaptic neuron it owns; ``rand()``/``randn()`` in expressions assigned to synaptic
        variables are functions of the synapse, not of its position in a stream.  Statistically
        equivalent to the reference, identical on any number of GPUs, and the way to networks
        that do not fit one host (10^9 synapses).
        """,
        validator=lambda v: v in ("reference", "sharded"),
    ),
    gather_synapses_limit=BrianPreference(
        default=50_000_000,
        docs="""
        Several GPUs, sharded construction: after a run the per-rank synapses of a `Synapses`
        object (indices and every synaptic variable) are gathered on all ranks only if there are
        at most this many of them in total; larger objects stay distributed (each rank sees the
        synapses of its own postsynaptic neurons).
        """,
    ),
    csr_l2_evict_last=BrianPreference(
        default=False,
        docs="""
        EXPERIMENTAL (not measured yet): read the packed CSR index stream with an L2 `evict_last`
        cache policy, so that the rows of neurons that fire rarely survive in the 126 MB L2 between
        their spikes instead of being refetched from DRAM (COBAHH-256k: the last load of the
        propagation chain).  A pure cache hint: results do not change.
        """,
    ),
)


class _SubprocessShim:
    """Stands in for the ``subprocess`` module inside ``CPPStandaloneDevice.run`` so that the
    inherited run-argument handling is reused while the project is executed in-process."""

    def __init__(self, runner):
        self._runner = runner

    def call(self, cmd, stdout=None, **kwds):
        return self._runner(cmd, stdout)


class B200Device(CPPStandaloneDevice):
    """Brian2 device that runs the per-timestep hot loop on one NVIDIA B200 (sm_100a)."""

    def __init__(self):
        super().__init__()
        #: access summary (read / write / scattered) of every device code object
        self._b200_access = {}
        #: extra information recorded when a code object is created
        self._b200_info = {}
        #: one entry per `Network.run` call: list of (clock, codeobj)
        self._b200_plans = []
        self._b200_memo_slots = {}      # state monitor name -> per-CTA memo slot (statemonitor.cu)
        self._b200_stream_counter = 0
        self._b200_seed = None
        self._b200_library = None
        self._b200_run_counter = 0
        #: out-of-band communicator of a multi-GPU run (brian2_b200.multigpu.Communicator)
        self._b200_comm = getattr(self, "_b200_comm", None)
        self._b200_written_vars = set()
        #: arrays from function namespaces (TimedArray values ...): name -> (ctype, size)
        self._b200_func_arrays = {}
        #: names of the Synapses objects whose synapses exist per rank only (sharded construction)
        self._b200_sharded_synapses = set()
        self._b200_sharded_objects = {}
        self.cu_source_files = []

    # ------------------------------------------------------------------------------------------
    # code objects
    # ------------------------------------------------------------------------------------------
    def code_object_class(self, codeobj_class=None, fallback_pref=None):
        if codeobj_class is None:
            return B200CodeObject
        return codeobj_class

    def code_object(
        self,
        owner,
        name,
        abstract_code,
        variables,
        template_name,
        variable_indices,
        codeobj_class=None,
        template_kwds=None,
        override_conditional_write=None,
        compiler_kwds=None,
    ):
        sharded = prefs.devices.b200.construction == "sharded"
        synapses_name = getattr(getattr(owner, "synapses", owner), "name", None)
        if sharded and template_name in RUN_ONCE_DEVICE_TEMPLATES:
            codeobj_class = B200ConnectCodeObject
        elif template_name in HOST_TEMPLATES:
            codeobj_class = B200HostCodeObject
            if synapses_name in self._b200_sharded_synapses:
                if template_name.startswith("synapses_create"):
                    raise NotImplementedError(
                        f"b200 sharded construction: '{synapses_name}' already has synapses created on "
                        "the device; connect(i=array, j=array) cannot be mixed with them"
                    )
                codeobj_class = B200ShardHostCodeObject
                compiler_kwds = dict(compiler_kwds or {})
                compiler_kwds["headers"] = list(compiler_kwds.get("headers", [])) + ['"b200_synrng.h"']
        elif template_name in DEVICE_TEMPLATES:
            codeobj_class = B200CodeObject
        else:
            raise NotImplementedError(
                f"The b200 device has no CUDA template for '{template_name}' code objects yet."
            )
        template_kwds = dict(template_kwds) if template_kwds is not None else {}
        if codeobj_class is B200ConnectCodeObject:
            import zlib

            n_calls = sum(1 for info in self._b200_info.values()
                          if info["template"] == template_name and info["owner"] is owner)
            template_kwds["b200_stream_id"] = zlib.crc32(f"{owner.name}.connect.{n_calls}".encode())
            template_kwds["b200_template_name"] = template_name
            post_parent = getattr(owner.target, "source", owner.target)
            template_kwds["b200_post_parent_size"] = int(len(post_parent))
        if codeobj_class is B200CodeObject:
            import zlib

            clock = getattr(owner, "clock", None)
            template_kwds["b200_clock"] = clock.name if clock is not None else "defaultclock"
            # RNG stream of a code object: stable across run() calls (the counter part of the
            # generator carries the time step), distinct between code objects
            template_kwds["b200_stream_id"] = zlib.crc32(
                f"{owner.name}.{template_name}.{name.rstrip('*')}".encode()
            )
            template_kwds["b200_template_name"] = template_name
            if template_name == "synapses_push_spikes":
                # size of the group the postsynaptic indices refer to (partition by post neuron)
                post_group = owner.synapses.target
                post_parent = getattr(post_group, "source", post_group)
                template_kwds["b200_post_parent_size"] = int(len(post_parent))
            if template_name == "spikegenerator":
                template_kwds["eventspace_variable"] = owner.variables["_spikespace"]
            if template_name == "summed_variable":
                from brian2.groups.neurongroup import NeuronGroup

                template_kwds["b200_target_whole_group"] = isinstance(owner.target, NeuronGroup)
            if template_name == "statemonitor":
                from brian2.groups.neurongroup import NeuronGroup

                src = owner.source
                template_kwds["b200_source_size"] = int(len(src)) if isinstance(src, NeuronGroup) else None
                # slot of the per-CTA "records nothing here" memo (csrc/b200_runtime.cuh)
                if owner.name not in self._b200_memo_slots:
                    n = len(self._b200_memo_slots)
                    self._b200_memo_slots[owner.name] = n if n < 8 else None   # 8 memo slots per CTA
                template_kwds["b200_memo_slot"] = self._b200_memo_slots[owner.name]
        # seen by the CUDA generator while it translates this code object (pathway direction)
        self._b200_current_template_kwds = template_kwds
        codeobj = super().code_object(
            owner,
            name,
            abstract_code,
            variables,
            template_name,
            variable_indices,
            codeobj_class=codeobj_class,
            template_kwds=template_kwds,
            override_conditional_write=override_conditional_write,
            compiler_kwds=compiler_kwds,
        )
        if codeobj_class is B200ConnectCodeObject:
            self._b200_sharded_synapses.add(owner.name)
            try:
                self._b200_sharded_objects[owner.name] = owner.__repr__.__self__
            except (AttributeError, ReferenceError):
                pass
        if codeobj_class in (B200CodeObject, B200ConnectCodeObject):
            try:    # a strong reference: `build()` may be called after the script's objects went
                owner = owner.__repr__.__self__     # out of scope (build_on_run=False)
            except (AttributeError, ReferenceError):
                pass
            self._b200_info[codeobj.name] = {
                "template": template_name,
                "owner": owner,
                "template_kwds": template_kwds,
            }
        return codeobj

    def is_device_codeobj(self, codeobj):
        return isinstance(codeobj, B200CodeObject)

    # ------------------------------------------------------------------------------------------
    # seed
    # ------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        super().seed(seed)
        if seed is None:
            seed = int.from_bytes(os.urandom(7), "little")
        # device RNG streams (in-loop rand/randn) are keyed on the same seed
        self.main_queue.append(
            ("insert_code", f"b200::state().seed = {int(seed)}ULL; b200::state().seeded = true;")
        )

    # ------------------------------------------------------------------------------------------
    # run(): record the schedule of this run as a plan for a persistent kernel
    # ------------------------------------------------------------------------------------------
    def network_run(
        self,
        net,
        duration,
        report=None,
        report_period=10 * second,
        namespace=None,
        profile=None,
        level=0,
        **kwds,
    ):
        if namespace is None:
            namespace = get_local_namespace(level=level + 2)
        build_on_run = self.build_on_run
        self.build_on_run = False
        try:
            super().network_run(
                net,
                duration,
                report=report,
                report_period=report_period,
                namespace=namespace,
                profile=profile,
                level=level + 1,
                **kwds,
            )
        finally:
            self.build_on_run = build_on_run
        # attach the plan to the generated `net.run(...)` line
        # (after_run code objects may have been queued behind it: search backwards)
        run_lines = None
        for func, args in reversed(self.main_queue):
            if func == "run_network" and args[0] is net:
                run_lines = args[1]
                break
        assert run_lines is not None
        # schedule of this run in execution order, recovered from the generated
        # `net.add(&clock, _run_<codeobj>);` lines (the objects' code-object lists are cleared
        # again by net.after_run())
        import re as _re

        clocks_by_name = {clock.name: clock for clock in self.clocks}
        entries = []
        compactions = []     # implicit companion of every thresholder: compaction of its event space
        last_add = -1
        for i, line in enumerate(run_lines):
            m = _re.match(rf"\s*{_re.escape(net.name)}\.add\(&(\w+), _run_(\w+)\);", line)
            if m and m.group(2) in self.code_objects:
                last_add = i
                codeobj = self.code_objects[m.group(2)]
                clock = clocks_by_name[m.group(1)]
                entries.append((clock, codeobj))
                info = self._b200_info.get(codeobj.name)
                if info is not None and info["template"] in SPIKE_SOURCE_TEMPLATES:
                    es = info["template_kwds"]["eventspace_variable"]
                    compactions.append((clock, self.get_array_name(es, access_data=False)))
        # Only order-dependent (serial) synaptic code reads the compacted list of the CURRENT
        # step; everything else reads the segments or older steps, so the compaction normally
        # goes to the end of the step where it costs no extra barrier.
        def needs_early(es_name):
            for _, co in entries:
                if isinstance(co, tuple):
                    continue
                info = self._b200_info.get(co.name)
                if info is None or info["template"] != "synapses":
                    continue
                pathway = info["template_kwds"]["pathway"]
                src_es = pathway.source.variables[pathway.eventspace_name]
                if (self.get_array_name(src_es, access_data=False) == es_name
                        and self._b200_access.get(co.name, {}).get("serial")):
                    return True
            return False

        # counted pathways: the apply pass runs right after the delivery (stepwise mode: its own
        # launch; the persistent kernel places it through _plan_barriers)
        for i in range(len(run_lines) - 1, -1, -1):
            m = _re.match(rf"\s*{_re.escape(net.name)}\.add\(&(\w+), _run_(\w+)\);", run_lines[i])
            if m and m.group(2) in self.code_objects and \
                    self._b200_access.get(m.group(2), {}).get("counted"):
                run_lines.insert(i + 1, f"{net.name}.add(&{m.group(1)}, _run_{m.group(2)}_apply);")
                if i <= last_add:
                    last_add += 1
        extra_lines = []
        for clock, es_name in compactions:
            item = (clock, ("compact", es_name, clock.name))
            line = f"{net.name}.add(&{clock.name}, _run_b200_compact{es_name});"
            if needs_early(es_name):
                pos = next(k for k, (_, co) in enumerate(entries)
                           if not isinstance(co, tuple) and self._b200_info.get(co.name, {}).get("template") in SPIKE_SOURCE_TEMPLATES
                           and self.get_array_name(self._b200_info[co.name]["template_kwds"]["eventspace_variable"],
                                                   access_data=False) == es_name)
                entries.insert(pos + 1, item)
                # same position in the generated net.add sequence
                k = [j for j, l in enumerate(run_lines) if f"_run_{entries[pos][1].name});" in l][0]
                run_lines.insert(k + 1, line)
                last_add += 1
            else:
                entries.append(item)
                extra_lines.append(line)
        run_lines[last_add + 1:last_add + 1] = extra_lines
        plan_index = len(self._b200_plans)
        self._b200_plans.append(entries)
        run_call = f"{net.name}.run("
        for i, line in enumerate(run_lines):
            if line.startswith(run_call) and line.rstrip().endswith(");"):
                run_lines[i] = line.rstrip()[:-2] + f", &_b200_plan_{plan_index});"
        if self.build_on_run:
            if self.has_been_run:
                raise RuntimeError(
                    "The network has already been built and run before. Use set_device with "
                    "build_on_run=False and an explicit device.build call to use multiple run "
                    "statements with this device."
                )
            self.build(direct_call=False, **self.build_options)

    # ------------------------------------------------------------------------------------------
    # barrier analysis
    # ------------------------------------------------------------------------------------------
    def _codeobj_access(self, codeobj):
        """Sets of array names: private (owned-partition) reads/writes and shared reads/writes."""
        info = self._b200_info[codeobj.name]
        acc = self._b200_access.get(
            codeobj.name,
            {"read": set(), "write": set(), "scattered_read": set(), "scattered_write": set()},
        )
        template = info["template"]
        kw = info["template_kwds"]
        # "owned": every element is touched only by the CTA that owns it in the common partition
        owned = self._is_owned_type(codeobj)
        priv_r = set(acc["read"]) if owned else set()
        priv_w = set(acc["write"]) if owned else set()
        shared_r = set(acc["scattered_read"]) | (set() if owned else set(acc["read"]))
        shared_w = set(acc["scattered_write"]) | (set() if owned else set(acc["write"]))
        name_of = lambda var: self.get_array_name(var, access_data=False)
        es = kw.get("eventspace_variable")
        tmpl = getattr(codeobj.templater, template)
        for varname in tmpl.writes_read_only:
            if varname in codeobj.variables and isinstance(codeobj.variables[varname], ArrayVariable):
                shared_w.add(name_of(codeobj.variables[varname]))
        if template in SPIKE_SOURCE_TEMPLATES:
            # the CTA's own segment of the event space: other CTAs read it -> shared write
            shared_w.add(name_of(es))
            if kw.get("_uses_refractory"):
                priv_w.add(name_of(codeobj.variables["not_refractory"]))
                priv_w.add(name_of(codeobj.variables["lastspike"]))
        elif template == "reset":
            pass   # reads the CTA's own segment only (written by the same CTA's thresholder)
        elif template == "spikemonitor":
            shared_r.add(name_of(es))
        elif template == "synapses":
            if es is None:
                es = kw["pathway"].source.variables[kw["pathway"].eventspace_name]
            shared_r.add(name_of(es))
            if acc.get("serial"):
                shared_r.add(name_of(es) + "__compact")
        elif template == "ratemonitor":
            shared_r.add(name_of(codeobj.variables["_spikespace"]))
        elif template == "summed_variable":
            # gather by target element: all reads are "somebody else's" elements; the target
            # itself is written by its owner when the target is a whole group
            shared_r |= priv_r
            priv_r = set()
            target = name_of(kw["_target_var"])
            shared_w.discard(target)
            priv_w.discard(target)
            (priv_w if kw.get("b200_target_whole_group") else shared_w).add(target)
        if template in ("spikemonitor", "statemonitor", "ratemonitor"):
            for var in self._monitor_buffers(codeobj):
                shared_w.add(name_of(var))
        if template == "synapses" and acc.get("counted") and not acc["counted"]["dual"]:
            # the delivery only counts events; the target arrays are touched by the apply pass,
            # element-private (listed here so that they reach the device and come back)
            priv_r |= set(acc["counted"]["read"])
            priv_w |= set(acc["counted"]["write"])
        # scalars that never change inside a run cannot create a dependency
        drop = set()
        for var in codeobj.variables.values():
            if isinstance(var, ArrayVariable) and clock_field(var) is not None:
                drop.add(name_of(var))
        return priv_r - drop, priv_w - drop, shared_r - drop, shared_w - drop

    # Resources of the barrier analysis are (name, lo, hi, private): `name` an array (or one of the
    # per-step structures below), [lo, hi] the range of TIME STEPS (relative to the current one)
    # whose instance of the structure is touched -- ordinary arrays exist once: (-inf, +inf) --,
    # `private` = every element is touched only by the CTA that owns it in the common partition.
    #   "<es>"            per-CTA segments of a step's spike list (ring of past steps)
    #   "<es>#own"        a CTA's own segment (thresholder -> resetter, same CTA)
    #   "<es>__compact"   the compacted (reference-layout) list of a step
    #   "hits:<pathway>"  per-target event counters of a counted pathway (double buffered)
    _ALWAYS = (float("-inf"), float("inf"))

    def _item_resources(self, codeobj, variant):
        """(reads, writes) of one in-loop code object as lists of resources (see above).
        `variant`: 'd0' -- a pathway may deliver in the step the spike was emitted in (reads the
        current step's segments, older steps from their compacted lists); 'd1' -- every delay is
        at least one step (reads the previous step's segments and compacted lists at least two
        steps old): what allows a step without an end-of-step barrier."""
        info = self._b200_info[codeobj.name]
        template, kw = info["template"], info["template_kwds"]
        acc = self._b200_access.get(codeobj.name, {})
        pr, pw, sr, sw = self._codeobj_access(codeobj)
        R = [(n, *self._ALWAYS, True) for n in pr] + [(n, *self._ALWAYS, False) for n in sr]
        W = [(n, *self._ALWAYS, True) for n in pw] + [(n, *self._ALWAYS, False) for n in sw]
        name_of = lambda var: self.get_array_name(var, access_data=False)
        es = kw.get("eventspace_variable")
        if template == "synapses":
            es = kw["pathway"].source.variables[kw["pathway"].eventspace_name]
        elif template == "ratemonitor":
            es = codeobj.variables["_spikespace"]
        E = name_of(es) if es is not None else None
        # shared (scalar) variables that some in-loop code writes are nobody's private data
        shared_scalars = set()
        for a in self._b200_access.values():
            shared_scalars |= set(a.get("scalar_write", ()))
        R = [(n, lo, hi, p and n not in shared_scalars) for (n, lo, hi, p) in R]
        W = [(n, lo, hi, p and n not in shared_scalars) for (n, lo, hi, p) in W]
        # the event-space entries of _codeobj_access are replaced by step-resolved ones
        R = [r for r in R if r[0] not in (E, f"{E}__compact")]
        W = [w for w in W if w[0] != E]
        if template in SPIKE_SOURCE_TEMPLATES:
            W += [(E, 0, 0, False), (f"{E}#own", 0, 0, True)]
        elif template == "reset":
            R += [(f"{E}#own", 0, 0, True)]
        elif template in ("spikemonitor", "ratemonitor"):
            R += [(E, 0, 0, False)]
        elif template == "synapses":
            if acc.get("serial"):
                R += [(f"{E}__compact", float("-inf"), 0, False)]
            elif variant == "d1":
                R += [(E, -1, -1, False), (f"{E}__compact", float("-inf"), -2, False)]
            else:
                R += [(E, 0, 0, False), (f"{E}__compact", float("-inf"), -1, False)]
            if acc.get("counted") and not acc["counted"]["dual"]:
                # the delivery itself only counts; the target arrays belong to the apply item
                mine = set(acc["counted"]["read"]) | set(acc["counted"]["write"])
                R = [r for r in R if r[0] not in mine]
                W = [w for w in W if w[0] not in mine] + [(f"hits:{kw['pathway'].name}", 0, 0, False)]
        return R, W

    @staticmethod
    def _resource_conflict(first, second, shift=0):
        """Dependency between an earlier item `first` and a later item `second` (= (R, W)); the
        later one's step ranges are moved by `shift` steps (1: it belongs to the next time step).
        Returns 'hard' (different CTAs may touch the same data: a grid barrier must separate
        them), 'soft' (element-private on both sides: program order inside the owning CTA is
        enough) or None."""
        R1, W1 = first
        R2, W2 = second
        found = None
        for a_list, b_list in ((W1, R2 + W2), (R1, W2)):
            for (na, la, ha, pa) in a_list:
                for (nb, lb, hb, pb) in b_list:
                    if na != nb or ha < lb + shift or hb + shift < la:
                        continue
                    if pa and pb:
                        found = found or "soft"
                    else:
                        return "hard"
        return found

    def _plan_barriers(self, entries, variant="d0"):
        """Phases of one time step.  Every item goes into the earliest phase its dependencies
        allow: phase(i) = max(phase(j) + 1 over earlier items j it shares data with across CTAs,
        phase(j) over earlier items it only shares element-private data with); inside a phase
        the schedule order is kept.  Items of one phase are mutually independent (or
        element-private), so a grid barrier is only needed between phases -- and at the end of
        the step only if an item of the first phase of the NEXT step depends on an item of the
        last phase of this one.  Returns (items, end_barrier)."""
        items = []

        def add(name, kind, res, owned, extra=None, exempt=None):
            item = {"name": name, "kind": kind, "res": res, "owned": owned, "share": None,
                    "barrier": False, "phase": 0, "order": len(items)}
            if extra:
                item.update(extra)
            for other in items:
                if other is exempt:
                    continue
                dep = self._resource_conflict(other["res"], res)
                if dep == "hard":
                    item["phase"] = max(item["phase"], other["phase"] + 1)
                elif dep == "soft":
                    item["phase"] = max(item["phase"], other["phase"])
            items.append(item)

        for clock, codeobj in entries:
            if isinstance(codeobj, tuple):   # ("compact", event space array name, clock name)
                _, es_name, clk = codeobj
                add(f"compact{es_name}", "compact",
                    ([(es_name, 0, 0, False)], [(es_name + "__compact", 0, 0, False)]), False,
                    {"es": es_name, "clock": clk})
                continue
            if not self.is_device_codeobj(codeobj):
                raise NotImplementedError(
                    f"Code object '{codeobj.name}' cannot run inside the simulation loop on the b200 device"
                )
            info = self._b200_info[codeobj.name]
            if info["template"] == "synapses_push_spikes":
                continue
            res = self._item_resources(codeobj, variant)
            add(codeobj.name, "codeobj", res, self._is_owned_type(codeobj),
                {"weight": 24 if info["template"] == "synapses" else 1, "variant": variant})
            counted = self._b200_access.get(codeobj.name, {}).get("counted")
            if info["template"] == "synapses" and counted:
                # the owners of the targets apply the counted events (element-private); with
                # dense rows they also do the counting, straight from the spike lists
                delivery = items[-1]
                E = self.get_array_name(info["template_kwds"]["pathway"].source.variables[
                    info["template_kwds"]["pathway"].eventspace_name], access_data=False)
                lists = [r for r in res[0] if r[0] in (E, f"{E}__compact")]
                R = [(n, *self._ALWAYS, True) for n in counted["read"]] + lists
                W = [(n, *self._ALWAYS, True) for n in counted["write"]]
                if counted["dual"]:
                    # either the delivery (sparse rows: fp reductions) or the apply pass (dense
                    # rows) does the work of this pathway, never both: no ordering between them
                    add(codeobj.name, "apply", (R, W), True,
                        {"variant": variant, "dual": True, "pathway": info["template_kwds"]["pathway"].name},
                        exempt=delivery)
                    items[-1]["phase"] = max(items[-1]["phase"], delivery["phase"])
                else:
                    R.append((f"hits:{info['template_kwds']['pathway'].name}", 0, 0, False))
                    add(codeobj.name, "apply", (R, W), True,
                        {"variant": variant, "dual": False, "pathway": info["template_kwds"]["pathway"].name})
        items.sort(key=lambda it: (it["phase"], it["order"]))
        last = -1
        for it in items:
            it["barrier"] = it["phase"] != last and last >= 0
            last = it["phase"]
        # end-of-step barrier: first phase of the next step against the last phase of this one
        n_phases = (items[-1]["phase"] + 1) if items else 0
        end_barrier = n_phases <= 1
        if not end_barrier:
            first = [it for it in items if it["phase"] == 0]
            later = [it for it in items if it["phase"] > 0]
            for b_item in first:
                for a_item in later:
                    if self._resource_conflict(a_item["res"], b_item["res"], shift=1) == "hard":
                        end_barrier = True
            # (items of phase 0 against each other one step apart: always separated by the
            # barriers in between, of which there is at least one)
        # Side-by-side execution: the code objects of one phase are mutually independent (that is
        # what "no barrier between them" means), and the ones that are not tied to the element
        # partition are short latency chains -- each gets its own share of the CTAs instead of
        # all CTAs walking through them one after the other.
        if prefs.devices.b200.split_phases:
            # Without an end-of-step barrier a CTA that is late out of the last phase is simply
            # late into the first phase of the next step.  The short latency chains of the last
            # phase (monitors, compaction) therefore go to the first 1/16 of the CTAs, and the
            # deliveries of the first phase to the others: the monitors' round trips disappear
            # behind the deliveries instead of adding to them.
            heavy0 = [it for it in items if it["phase"] == 0 and not it["owned"] and it.get("weight", 1) > 1]
            light_last = [it for it in items if it["phase"] == n_phases - 1 and not it["owned"]
                          and it.get("weight", 1) == 1]
            stagger = (not end_barrier and n_phases >= 2 and heavy0 and light_last
                       and len(heavy0) == sum(1 for it in items if it["phase"] == 0 and not it["owned"])
                       and len(light_last) == sum(1 for it in items if it["phase"] == n_phases - 1 and not it["owned"]))
            for ph in range(n_phases):
                free = [it for it in items if it["phase"] == ph and not it["owned"]]
                if stagger and ph in (0, n_phases - 1):
                    # shares in units of 1/(16 * total): [0, total) = first 1/16 of the grid
                    total = sum(it.get("weight", 1) for it in free)
                    acc = 0
                    for it in free:
                        w = it.get("weight", 1)
                        if ph == 0:
                            it["share"] = (total + 15 * acc, total + 15 * (acc + w), 16 * total)
                        else:
                            it["share"] = (acc, acc + w, 16 * total)
                        acc += w
                elif len(free) > 1:
                    total = sum(it.get("weight", 1) for it in free)
                    acc = 0
                    for it in free:
                        w = it.get("weight", 1)
                        it["share"] = (acc, acc + w, total)
                        acc += w
        for it in items:
            del it["res"]
        return items, end_barrier

    def _is_owned_type(self, codeobj):
        info = self._b200_info[codeobj.name]
        if info["template"] == "summed_variable":
            return bool(info["template_kwds"].get("b200_target_whole_group"))
        return info["template"] in ("stateupdate", "threshold", "reset", "spikegenerator") or (
            info["template"] == "statemonitor" and info["template_kwds"].get("b200_source_size") is not None
        )

    # ------------------------------------------------------------------------------------------
    # monitors
    # ------------------------------------------------------------------------------------------
    def _monitor_buffers(self, codeobj):
        info = self._b200_info[codeobj.name]
        owner = info["owner"]
        template = info["template"]
        if template == "spikemonitor":
            names = sorted(info["template_kwds"].get("record_variables", {}).keys())
            return [owner.variables[n] for n in names]
        if template == "statemonitor":
            rec = info["template_kwds"].get("_recorded_variables", {})
            return [owner.variables["t"]] + [rec[k] for k in sorted(rec.keys())]
        if template == "ratemonitor":
            return [owner.variables["rate"], owner.variables["t"]]
        return []

    def _collect_monitors(self):
        monitors = []
        seen = set()
        for codeobj in self.code_objects.values():
            if not self.is_device_codeobj(codeobj):
                continue
            info = self._b200_info[codeobj.name]
            template = info["template"]
            if template not in ("spikemonitor", "statemonitor", "ratemonitor"):
                continue
            owner = info["owner"]
            # every run() call creates a new code object for the same monitor
            if owner.name in seen:
                continue
            seen.add(owner.name)
            kind = {"spikemonitor": "spike", "statemonitor": "state", "ratemonitor": "rate"}[template]
            buffers = []
            width = 1
            if template == "statemonitor":
                width = int(owner.variables["_indices"].size)
            for var in self._monitor_buffers(codeobj):
                ndim = getattr(var, "ndim", 1)
                buffers.append(
                    {
                        "name": self.arrays[var],
                        "ctype": c_data_type(var.dtype),
                        "ndim": ndim,
                        "width": width if ndim == 2 else 1,
                    }
                )
            headroom = 1
            if template == "spikemonitor":
                headroom = max(1, int(owner.source.stop - owner.source.start))
            monitors.append(
                {
                    "name": owner.name,
                    "kind": kind,
                    "N_array": self.arrays[owner.variables["N"]],
                    "buffers": buffers,
                    "headroom": headroom,
                    "width": width,
                }
            )
        return monitors

    # ------------------------------------------------------------------------------------------
    # source generation
    # ------------------------------------------------------------------------------------------
    def _array_table(self, monitors):
        """Description of every array for the device array table template."""
        used, written = set(), set()
        import re as _re

        def referenced(var, code):
            """Does the generated device code touch the array (beyond declaring a pointer to it)?
            Only asked for arrays that are constant during runs: the synaptic index arrays
            `_synaptic_pre/_post` are in every synaptic code object's variables, but the CUDA
            templates take both ends of a synapse from the CSR, so these (large) arrays need
            not be copied to the device at all unless user code reads `i`/`j`."""
            name = _re.escape(self.arrays[var])
            return bool(_re.search(rf"_ptr{name}\s*\[|_A\.{name}\b(?!\s*;)", code))

        for codeobj in self.code_objects.values():
            if not self.is_device_codeobj(codeobj):
                continue
            code = None
            if self._b200_info[codeobj.name]["template"] in RUN_ONCE_DEVICE_TEMPLATES:
                continue    # works on its own scratch buffers and the host mirrors
            if self._b200_info[codeobj.name]["template"] == "synapses_push_spikes":
                # no device code at all (the spike ring makes the push a no-op): its variables
                # -- the per-synapse `delay` array above all -- are read by the HOST-side
                # before_run block that builds the CSR, straight from the host arrays
                continue
            for var in codeobj.variables.values():
                if isinstance(var, ArrayVariable):
                    if (var.constant and var.read_only and var in self.dynamic_arrays
                            and self._b200_info[codeobj.name]["template"] == "synapses"):
                        if code is None:
                            code = str(codeobj.code.cpp_file)
                        if not referenced(var, code):
                            continue
                    used.add(var)
            if self._b200_info[codeobj.name]["template"] == "synapses_push_spikes":
                continue
            _, pw, _, sw = self._codeobj_access(codeobj)
            names = pw | sw
            for var in codeobj.variables.values():
                if isinstance(var, ArrayVariable) and self.get_array_name(var, access_data=False) in names:
                    written.add(var)
        self._b200_written_vars = set(written)
        monitor_of, min_cap = {}, {}
        for mon in monitors:
            for b in mon["buffers"]:
                monitor_of[b["name"]] = mon["name"]
                min_cap[b["name"]] = max(64 * mon["headroom"], 1 << 16) if mon["kind"] == "spike" else 1024
        table = []
        for var, name in sorted(self.arrays.items(), key=lambda kv: kv[1]):
            try:
                owner_name = var.owner.name
            except (ReferenceError, AttributeError):
                owner_name = "_unknown"
            entry = {
                "name": name,
                "ctype": c_data_type(var.dtype),
                "used": var in used,
                "written": var in written,
                "user_name": f"{owner_name}.{var.name}",
                "monitor": monitor_of.get(name),
                "min_cap": min_cap.get(name, 0),
                "eventspace": False,
                "clock": None,
                "width": 1,
                "dyn_name": None,
                "size": 0,
            }
            if var in self.dynamic_arrays:
                entry["kind"] = "dynamic1d"
                entry["dyn_name"] = self.dynamic_arrays[var]
            elif var in self.dynamic_arrays_2d:
                entry["kind"] = "dynamic2d"
                entry["dyn_name"] = self.dynamic_arrays_2d[var]
                width = 1
                for mon in monitors:
                    if mon["name"] == monitor_of.get(name):
                        width = mon["width"]
                entry["width"] = width
            else:
                entry["kind"] = "static"
                entry["size"] = int(var.size)
                if is_eventspace(var):
                    entry["eventspace"] = True
                    entry["clock"] = var.owner.clock.name
                if clock_field(var) is not None:
                    entry["used"] = False   # clocks travel by value
            table.append(entry)
        # read-only arrays of function namespaces (host side: the reference's static arrays,
        # devices/cpp_standalone/device.py:1094-1116)
        for name, (ctype, size) in sorted(self._b200_func_arrays.items()):
            table.append({
                "name": name, "ctype": ctype, "used": True, "written": False,
                "user_name": f"_function.{name}", "monitor": None, "min_cap": 0,
                "eventspace": False, "clock": None, "width": 1, "dyn_name": None,
                "size": size, "kind": "static",
            })
        return table

    def _eventspaces(self):
        spaces = {}
        for codeobj in self.code_objects.values():
            if not self.is_device_codeobj(codeobj):
                continue
            for var in codeobj.variables.values():
                if is_eventspace(var):
                    spaces[self.arrays[var]] = {
                        "name": self.arrays[var],
                        "size": int(var.size),
                        "clock": var.owner.clock.name,
                        "compact_always": False,
                    }
        # The compacted (reference-layout) list of a step is only built when somebody reads it:
        # pathways with a delay (decided at run time from the delays), and -- decided here --
        # order-dependent synaptic code, which walks the compact list of the current step.
        for codeobj in self.code_objects.values():
            info = self._b200_info.get(codeobj.name)
            if info is None or info["template"] != "synapses":
                continue
            if self._b200_access.get(codeobj.name, {}).get("serial"):
                pathway = info["template_kwds"]["pathway"]
                name = self.arrays[pathway.source.variables[pathway.eventspace_name]]
                if name in spaces:
                    spaces[name]["compact_always"] = True
        return [spaces[k] for k in sorted(spaces)]

    def _summed_updaters(self):
        """Names of the SummedVariableUpdater objects (one by-target index each)."""
        names = set()
        for codeobj in self.code_objects.values():
            info = self._b200_info.get(codeobj.name)
            if info is not None and info["template"] == "summed_variable":
                # (one index per summed variable: a Synapses object may define `x_pre` and `x_post`
                # sums of equally named variables of two groups)
                target = self.get_array_name(info["template_kwds"]["_target_var"], access_data=False)
                names.add(f'{info["owner"].name}{target}')
        return sorted(names)

    def _pathways(self, synapses):
        out = []
        for S in sorted(synapses, key=lambda s: s.name):
            for path in sorted(S._pathways, key=lambda p: p.name):
                hits_n = tile_n = 0
                for codeobj in self.code_objects.values():
                    info = self._b200_info.get(codeobj.name)
                    if info is not None and info["template"] == "synapses" \
                            and info["template_kwds"]["pathway"].name == path.name:
                        counted = self._b200_access.get(codeobj.name, {}).get("counted")
                        if counted:
                            tile_n = int(counted["size"])
                            hits_n = 0 if counted["dual"] else tile_n
                out.append(
                    {
                        "name": path.name,
                        "sources": self.dynamic_arrays[path.synapse_sources],
                        "start": int(path.source.start),
                        "stop": int(path.source.stop),
                        "hits_n": hits_n,
                        "tile_n": tile_n,
                    }
                )
        return out

    def generate_objects_source(
        self, writer, arange_arrays, synapses, static_array_specs, networks, timed_arrays
    ):
        self._b200_synapses_objects = list(synapses)
        # host mirrors: the reference's own objects.cpp, minus its SynapticPathway objects
        host_tmp = CPPStandaloneCodeObject.templater.objects(
            None,
            None,
            array_specs=self.arrays,
            dynamic_array_specs=self.dynamic_arrays,
            dynamic_array_2d_specs=self.dynamic_arrays_2d,
            zero_arrays=self.zero_arrays,
            arange_arrays=arange_arrays,
            synapses=[],
            clocks=self.clocks,
            static_array_specs=static_array_specs,
            networks=networks,
            get_array_filename=self.get_array_filename,
            get_array_name=self.get_array_name,
            profiled_codeobjects=[],
            code_objects=list(self.code_objects.values()),
            timed_arrays=timed_arrays,
        )
        writer.write("objects.*", host_tmp)
        monitors = self._collect_monitors()
        dev_tmp = B200CodeObject.templater.b200_objects(
            None,
            None,
            clocks=self.clocks,
            array_specs=self.arrays,
            b200_arrays=self._array_table(monitors),
            b200_eventspaces=self._eventspaces(),
            b200_pathways=self._pathways(synapses),
            b200_monitors=monitors,
            b200_summed=self._summed_updaters(),
        )
        writer.write("b200_objects.h", dev_tmp.h_file)
        writer.write("b200_objects.cpp", dev_tmp.cpp_file)

    def generate_main_source(self, writer):
        # reuse the inherited construction of `main_lines`; the template lookup goes through
        # code_object_class() and therefore renders OUR main template
        B200CodeObject.templater.env.globals["profiled_codeobjects"] = list(self.profiled_codeobjects)
        # Everything that can change a host array between two run() calls bumps the runtime's
        # "host epoch": a pathway whose inputs are untouched since its CSR was built is not
        # rebuilt (and re-uploaded) by the `_before_run_*_push_spikes()` of the next run.
        mutators = {"run_code_object", "set_by_constant", "set_by_array", "set_by_single_value",
                    "set_array_by_array", "resize_array"}
        # (the clocks' t / timestep / dt are set before every run; pathways look at dt themselves)
        clock_arrays = {self.arrays[var] for clock in self.clocks for var in clock.variables.values()
                        if var in self.arrays}
        queue = []
        for func, args in self.main_queue:
            queue.append((func, args))
            if func == "set_by_single_value" and args[0] in clock_arrays:
                continue
            if func in mutators or (func == "insert_code" and "b200::state()" not in str(args)):
                queue.append(("insert_code", "b200::state().host_epoch++;"))
        original, self.main_queue = self.main_queue, queue
        try:
            super().generate_main_source(writer)
        finally:
            self.main_queue = original

    def generate_run_source(self, writer):
        run_tmp = CPPStandaloneCodeObject.templater.run(
            None,
            None,
            run_funcs=self.runfuncs,
            code_objects=list(self.code_objects.values()),
            user_headers=self.headers,
            array_specs=self.arrays,
            clocks=self.clocks,
        )
        writer.write("run.*", run_tmp)

    def _constant_lines(self, codeobj, device_side):
        """The `%CONSTANTS%` block of one code object (cf. device.py:908-940): array sizes and
        namespace constants; on the device side dynamic-array sizes come from the table `_A`."""
        renderer = CPPNodeRenderer()
        lines = []
        for k, v in codeobj.variables.items():
            if isinstance(v, ArrayVariable):
                try:
                    if isinstance(v, DynamicArrayVariable):
                        if v.ndim == 1:
                            arr, dyn = self.arrays[v], self.dynamic_arrays[v]
                            ctype = c_data_type(v.dtype)
                            if device_side:
                                lines.append(f"const size_t _num{k} = _A._n{arr};")
                            else:
                                lines.append(f"{ctype}* const {arr} = {dyn}.empty()? 0 : &{dyn}[0];")
                                lines.append(f"const size_t _num{k} = {dyn}.size();")
                    else:
                        lines.append(f"const size_t _num{k} = {v.size};")
                except TypeError:
                    pass
            elif isinstance(v, Constant):
                value = renderer.render_expr(repr(v.value))
                lines.append(f"const {c_data_type(v.dtype)} {k} = {value};")
        seen, unique = set(), []
        for line in lines:
            if line not in seen:
                seen.add(line)
                unique.append(line)
        return "\n".join(unique)

    def generate_codeobj_source(self, writer):
        device_objs = []
        canonical, alias_of = {}, {}
        for codeobj in self.code_objects.values():
            host_consts = self._constant_lines(codeobj, device_side=False)
            for block in codeobj.before_after_blocks:
                cpp_code = getattr(codeobj.code, f"{block}_cpp_file").replace("%CONSTANTS%", host_consts)
                writer.write(f"code_objects/{block}_{codeobj.name}.cpp", cpp_code)
                writer.write(f"code_objects/{block}_{codeobj.name}.h", getattr(codeobj.code, f"{block}_h_file"))
            if self.is_device_codeobj(codeobj):
                dev_consts = self._constant_lines(codeobj, device_side=True)
                code = codeobj.code.cpp_file.replace("%CONSTANTS_DEV%", dev_consts)
                code = code.replace("%CONSTANTS%", host_consts)
                # Every run() call re-creates the code objects of all its objects: identical
                # source (up to the name) is compiled once and shared.
                key = code.replace(codeobj.name, "\0")
                rep = canonical.setdefault(key, codeobj.name)
                if rep != codeobj.name:
                    alias_of[codeobj.name] = rep
                else:
                    writer.write(f"code_objects/{codeobj.name}.cuh", code)
                    device_objs.append(codeobj)
            else:
                code = codeobj.code.cpp_file.replace("%CONSTANTS%", host_consts)
                writer.write(f"code_objects/{codeobj.name}.cpp", code)
            writer.write(f"code_objects/{codeobj.name}.h", codeobj.code.h_file)

        # persistent kernels: one per run() call (two variants if delayed-only pathways allow a
        # cheaper schedule: chosen at run time, when the delays are known)
        plans = []
        for index, entries in enumerate(self._b200_plans):
            clocks = {clock for clock, _ in entries}
            plan = {"index": index, "variants": [], "clock": None, "signature": "", "n_barriers": 0}
            if len(clocks) == 1 and prefs.devices.b200.persistent and not self.enable_profiling_any:
                clock = next(iter(clocks))
                try:
                    for tag in ("d0", "d1"):
                        items, end_barrier = self._plan_barriers(entries, variant=tag)
                        for it in items:
                            it["name"] = alias_of.get(it["name"], it["name"])
                        signature = " ".join(
                            ("| " if it["barrier"] else "") + it["name"] + ("@apply" if it["kind"] == "apply" else "")
                            for it in items
                        ) + (" |" if end_barrier else "")
                        variant = {"tag": tag, "entries": items, "end_barrier": end_barrier,
                                   "signature": signature,
                                   "n_barriers": sum(1 for it in items if it["barrier"]) + (1 if end_barrier else 0)}
                        if plan["variants"] and plan["variants"][0]["signature"] == signature:
                            continue     # delays make no difference to this schedule
                        plan["variants"].append(variant)
                    plan["clock"] = clock.name
                    plan["signature"] = " // ".join(f"[{v['tag']}] {v['signature']}" for v in plan["variants"])
                    plan["n_barriers"] = plan["variants"][0]["n_barriers"]
                    # (kept for introspection / tests: the conservative variant)
                    plan["entries"] = plan["variants"][0]["entries"]
                except NotImplementedError as ex:
                    plan["variants"] = []
                    logger.warn(f"run #{index} falls back to stepwise execution: {ex}")
            # run() calls with an identical schedule share one kernel
            plan["alias"] = None
            if plan["clock"]:
                for other in plans:
                    if other["clock"] == plan["clock"] and other["signature"] == plan["signature"] and other["alias"] is None:
                        plan["alias"] = other["index"]
                        break
            plans.append(plan)
        self._b200_plan_info = plans
        user_headers = self.headers + prefs["codegen.cpp.headers"]
        kernels = B200CodeObject.templater.b200_kernels(
            None,
            None,
            device_code_objects=device_objs,
            code_object_aliases=[(name, rep, bool(self._b200_access.get(name, {}).get("counted")))
                                 for name, rep in sorted(alias_of.items())],
            b200_eventspaces=self._eventspaces(),
            b200_pathways=self._pathways(self._b200_synapses_objects),
            plans=[p for p in plans],
            user_headers=user_headers,
            profiled=bool(self.enable_profiling_any),
            ctas_per_sm=int(prefs.devices.b200.ctas_per_sm),
            profile_phases=bool(prefs.devices.b200.profile_phases),
        )
        writer.write("b200_kernels.cu", kernels)
        self.cu_source_files = ["b200_kernels.cu"]
        writer.write("b200_plans.h", B200CodeObject.templater.b200_plans(None, None, plans=plans))

    @property
    def enable_profiling_any(self):
        return bool(self.profiled_codeobjects)

    def copy_source_files(self, writer, directory):
        super().copy_source_files(writer, directory)
        # the spike queue of the reference is not compiled into a b200 project
        writer.source_files.discard("brianlib/spikequeue.cpp")
        writer.header_files.discard("brianlib/spikequeue.h")
        clocks = B200CodeObject.templater.b200_clocks(None, None)
        with open(os.path.join(directory, "brianlib", "clocks.h"), "w") as f:
            f.write(clocks)
        for fname in sorted(os.listdir(CSRC_DIR)):
            if fname.endswith((".h", ".cuh")):
                shutil.copy2(os.path.join(CSRC_DIR, fname), os.path.join(directory, fname))
                writer.header_files.add(fname)
        shutil.copy2(os.path.join(INCLUDE_DIR, "brian2_b200.h"), os.path.join(directory, "brian2_b200.h"))
        writer.header_files.add("brian2_b200.h")
        if prefs.devices.b200.libm == "glibc":
            from . import libm_tables

            libm_tables.write_header(directory)
            writer.header_files.add("b200_libm_tables.h")

    @property
    def library_name(self):
        return "libb200_project.so"

    def nvcc_flags(self):
        flags = [prefs.devices.b200.arch_flags] + list(prefs.devices.b200.extra_nvcc_flags)
        flags += ["-ccbin", os.environ.get("CXX", "g++")]
        flags.append("-fmad=true" if prefs.devices.b200.fmad else "-fmad=false")
        if prefs.core.default_float_dtype == np.float32:
            flags.append("-DB200_FLOAT32")
        if prefs.devices.b200.csr_l2_evict_last:
            flags.append("-DB200_CSR_EVICT_LAST")
        if prefs.devices.b200.libm == "glibc":
            flags.append("-DB200_GLIBC_MATH")
            if prefs.devices.b200.fmad:
                logger.warn("devices.b200.libm = 'glibc' makes the libm calls bit-identical to cpp_standalone, "
                            "but devices.b200.fmad = True lets nvcc fuse the a*b+c of the generated code: "
                            "state variables will not be bit-identical.", name_suffix="glibc_fmad", once=True)
        return " ".join(flags)

    def generate_makefile(self, writer, compiler, compiler_flags, linker_flags, nb_threads, debug):
        if nb_threads:
            raise NotImplementedError("The b200 device does not use OpenMP threads")
        sources = sorted(f for f in writer.source_files if f != "brianlib/spikequeue.cpp")
        makefile = B200CodeObject.templater.makefile(
            None,
            None,
            library_name=self.library_name,
            source_files=" ".join(sources),
            cu_source_files=" ".join(self.cu_source_files),
            header_files=" ".join(sorted(writer.header_files)),
            compiler_flags=compiler_flags,
            compiler_debug_flags="-g -DDEBUG" if debug else "",
            linker_debug_flags="-g" if debug else "",
            linker_flags=linker_flags,
            nvcc=prefs.devices.b200.nvcc,
            cuda_home=prefs.devices.b200.cuda_home,
            nvcc_flags=self.nvcc_flags(),
            nvcc_arch=prefs.devices.b200.arch_flags,
            rm_cmd="rm -f $(OBJS) $(CU_OBJS) $(LIBRARY)",
        )
        writer.write("makefile", makefile)

    # ------------------------------------------------------------------------------------------
    # running: in-process through the C ABI
    # ------------------------------------------------------------------------------------------
    def _run_library(self, cmd, stdout):
        """Replacement of ``subprocess.call(['./main', ...])``: load a fresh copy of the project
        library and call ``b200_run_main`` with the same arguments."""
        args = list(cmd[1:])   # cmd[0] is the reference's "./main"
        lib_path = os.path.join(os.getcwd(), self.library_name)
        self._b200_run_counter += 1
        lib = B200Library(lib_path, fresh_copy=self._b200_run_counter > 1)
        lib.set_option("mode", 0 if prefs.devices.b200.persistent else 1)
        lib.set_option("max_chunk", int(prefs.devices.b200.max_chunk))
        lib.set_option("ctas_per_sm", int(prefs.devices.b200.ctas_per_sm))
        lib.set_option("grid", int(prefs.devices.b200.grid))
        lib.set_option("allow_d1", 1 if prefs.devices.b200.elide_end_barrier else 0)
        lib.set_option("tiles", 1 if prefs.devices.b200.tiled_delivery else 0)
        lib.set_option("forward", 1 if prefs.devices.b200.forward_delivery else 0)
        comm = self.communicator()
        if comm.world > 1:
            self._check_multi_gpu_support()
            lib.set_comm(comm.rank, comm.world, comm.allgather)
        self._b200_library = lib
        status = lib.run_main(args, stdout=stdout)
        if status != 0:
            sys.stderr.write(f"b200 run failed: {lib.last_error()}\n")
        return status

    # ------------------------------------------------------------------------------------------
    # multi-GPU (one process per GPU): communicator, support check, merge of the results
    # ------------------------------------------------------------------------------------------
    def set_communicator(self, comm):
        """Use ``comm`` (`brian2_b200.multigpu.Communicator`) instead of the default, which is
        built on the initialised ``torch.distributed`` process group."""
        self._b200_comm = comm

    def communicator(self):
        from .multigpu import Communicator, default_communicator

        if self._b200_comm is not None:
            return self._b200_comm
        if not prefs.devices.b200.multi_gpu:
            return Communicator()
        return default_communicator()

    def _check_multi_gpu_support(self):
        """The partition by postsynaptic neuron keeps every write local as long as synaptic code
        writes synaptic and postsynaptic variables only."""
        from brian2.synapses.synapses import Synapses

        for codeobj in self.code_objects.values():
            info = self._b200_info.get(codeobj.name)
            if info is None:
                continue
            template, owner = info["template"], info["owner"]
            if template == "summed_variable":
                raise NotImplementedError("b200 multi-GPU: summed variables")
            if template == "stateupdate" and isinstance(owner, Synapses):
                raise NotImplementedError("b200 multi-GPU: clock-driven synaptic equations")
            if template == "statemonitor" and info["template_kwds"].get("b200_source_size") is None:
                raise NotImplementedError("b200 multi-GPU: StateMonitor of synapses or subgroups")
            if template == "synapses":
                acc = self._b200_access.get(codeobj.name, {})
                if acc.get("serial"):
                    raise NotImplementedError("b200 multi-GPU: order-dependent synaptic code")
                pathway = info["template_kwds"]["pathway"]
                for name in codeobj.variables:
                    if codeobj.variable_indices[name] == "_presynaptic_idx":
                        var = codeobj.variables[name]
                        if isinstance(var, ArrayVariable) and var in self._b200_written_vars:
                            raise NotImplementedError(
                                f"b200 multi-GPU: synaptic code of '{pathway.name}' accesses the "
                                f"presynaptic variable '{name}' that changes during the run"
                            )

    def _merge_multi_gpu_results(self, comm):
        """Combine the per-rank results (see brian2_b200/multigpu.py) into ``array_cache``."""
        from brian2.groups.neurongroup import NeuronGroup
        from brian2.monitors.ratemonitor import PopulationRateMonitor
        from brian2.monitors.spikemonitor import EventMonitor
        from brian2.monitors.statemonitor import StateMonitor
        from brian2.synapses.synapses import Synapses

        from . import multigpu as mg

        world = comm.world
        local = lambda var: CPPStandaloneDevice.get_value(self, var)
        merged = {}
        by_owner = defaultdict(list)
        for var in sorted(self._b200_written_vars, key=lambda v: self.arrays[v]):
            if clock_field(var) is not None or is_eventspace(var):
                continue
            try:
                by_owner[var.owner.name].append(var)
            except ReferenceError:
                continue
        owners = {}
        for codeobj in self.code_objects.values():
            info = self._b200_info.get(codeobj.name)
            if info is not None:
                owners[info["owner"].name] = info["owner"]
        # objects reach this point in the same (sorted) order on every rank
        for owner_name in sorted(by_owner):
            variables = by_owner[owner_name]
            owner = owners.get(owner_name)
            if owner is None:   # e.g. a NeuronGroup only written through synapses
                owner = variables[0].owner
            if isinstance(owner, EventMonitor):
                rec = {v.name: local(v) for v in variables if v.name not in ("N", "count")}
                parts = comm.allgather_object(rec)
                t_parts = [p["t"] for p in parts]
                t_all, cols = mg.merge_spike_records(t_parts, {k: [p[k] for p in parts] for k in rec if k != "t"})
                for v in variables:
                    if v.name == "t":
                        merged[v] = t_all
                    elif v.name in cols:
                        merged[v] = cols[v.name]
                    elif v.name == "count":
                        merged[v] = np.sum(np.stack(comm.allgather_object(local(v))), axis=0).astype(v.dtype)
                    elif v.name == "N":
                        merged[v] = np.array([len(t_all)], dtype=v.dtype)
            elif isinstance(owner, StateMonitor):
                indices = np.asarray(owner.variables["_indices"].get_value())
                for v in variables:
                    if getattr(v, "ndim", 1) == 2:
                        parts = comm.allgather_object(local(v))
                        merged[v] = mg.merge_state_columns(parts, indices, len(owner.source), world)
            elif isinstance(owner, PopulationRateMonitor):
                for v in variables:
                    if v.name == "rate":
                        parts = comm.allgather_object(local(v))
                        merged[v] = mg.merge_rate(parts, float(owner.clock.dt_), len(owner.source))
            elif isinstance(owner, Synapses) and owner.name in self._b200_sharded_synapses:
                pass    # gathered below (every rank holds different synapses)
            elif isinstance(owner, Synapses):
                post = local(owner.variables["_synaptic_post"])
                target = owner.target
                n_parent = len(getattr(target, "source", target))
                own = mg.owner_of(post, n_parent, world)
                for v in variables:
                    if getattr(v, "scalar", False):
                        continue
                    parts = comm.allgather_object(local(v))
                    merged[v] = mg.merge_by_owner(parts, own)
            elif isinstance(owner, NeuronGroup):
                for v in variables:
                    if getattr(v, "scalar", False):
                        continue
                    parts = comm.allgather_object(local(v))
                    merged[v] = mg.merge_by_block(parts, len(owner), world)
        # Synapses created per rank (sharded construction): the global object is the union of the
        # ranks' synapses in (pre, post) order -- the order a single-GPU run creates them in.
        limit = int(prefs.devices.b200.gather_synapses_limit)
        for name in sorted(self._b200_sharded_synapses):
            S = self._b200_sharded_objects.get(name)
            if S is None:
                continue
            pre_l = np.asarray(local(S.variables["_synaptic_pre"]))
            sizes = comm.allgather_object(int(len(pre_l)))
            if sum(sizes) > limit:
                logger.info(f"'{name}': {sum(sizes)} synapses stay distributed over the ranks "
                            "(prefs.devices.b200.gather_synapses_limit)")
                continue
            post_l = np.asarray(local(S.variables["_synaptic_post"]))
            order = mg.sharded_synapse_order(comm.allgather_object(pre_l), comm.allgather_object(post_l))
            seen = set()
            # (a set of Variable objects: iterate in an order every rank agrees on)
            for var in sorted(S._registered_variables, key=lambda v: self.arrays[v]):
                if var in seen:
                    continue
                seen.add(var)
                value = merged.get(var)
                if value is not None and len(value) == len(order):
                    continue
                parts = comm.allgather_object(np.asarray(local(var)))
                merged[var] = np.concatenate(parts)[order]
            for vname in ("N_incoming", "N_outgoing"):
                var = S.variables[vname]
                merged[var] = np.sum(np.stack(comm.allgather_object(np.asarray(local(var)))), axis=0).astype(var.dtype)
            merged[S.variables["N"]] = np.array([len(order)], dtype=S.variables["N"].dtype)
        for v, value in merged.items():
            self.array_cache[v] = value
            if isinstance(v, DynamicArrayVariable) and getattr(v, "ndim", 1) == 1:
                v.size = len(value)

    def run(self, directory=None, results_directory=None, with_output=True, run_args=None):
        import brian2.devices.cpp_standalone.device as _ref_device_module

        original = _ref_device_module.subprocess
        _ref_device_module.subprocess = _SubprocessShim(self._run_library)
        try:
            super().run(
                directory=directory,
                results_directory=results_directory,
                with_output=with_output,
                run_args=run_args,
            )
        finally:
            _ref_device_module.subprocess = original
        comm = self.communicator()
        if comm.world > 1:
            self._merge_multi_gpu_results(comm)

    def phase_profile(self, plan=0, all_ctas=False):
        """[(phase name, SM cycles spent by CTA 0)] of the persistent kernel of run() call `plan`
        (needs ``prefs.devices.b200.profile_phases = True`` at build time)."""
        info = self._b200_plan_info[plan]
        if info["alias"] is not None:
            info = self._b200_plan_info[info["alias"]]
        variant = info["variants"][-1] if (len(info["variants"]) > 1 and self.counter("all_delayed") > 0) \
            else info["variants"][0]
        names = []
        for it in variant["entries"]:
            names += ["barrier" if it["barrier"] else None, it["name"] + ("@apply" if it["kind"] == "apply" else "")]
        names.append("end-of-step barrier" if variant["end_barrier"] else "end of step (no barrier)")
        if not all_ctas:
            return [(n, self.counter(f"phase{i}")) for i, n in enumerate(names) if n is not None]
        # cycles of the four sampled CTAs (first, 1/4, 3/4, last of the grid)
        return [(n, [self.counter(f"phase{k * 512 + i}") for k in range(4)])
                for i, n in enumerate(names) if n is not None]

    # counters of the last run (for benchmarks)
    def counter(self, key):
        if self._b200_library is None:
            raise RuntimeError("The b200 project has not been run yet")
        return self._b200_library.get_counter(key)


b200_device = B200Device()
all_devices["b200"] = b200_device
