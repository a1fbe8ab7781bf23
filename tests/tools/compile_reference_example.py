"""Development aid (build container only: reads /root/reference): run a reference example script
with `set_device('b200', build_on_run=False)`, matplotlib stubbed, and cross-compile the
generated project for sm_100a with nvcc (no run -- there is no GPU here).  Usage:

    python tests/tools/compile_reference_example.py /root/reference/examples/synapses/STDP.py

Prints one RESULT line: where the script stopped (plotting / reading results before the run)
and whether the project built.  Round 1: 50 of the 53 scripts under examples/*.py,
examples/synapses and examples/frompapers that reach a `run()` build; the exceptions are
`run_regularly` code that writes shared variables (NotImplementedError) and scripts whose
objects are gone before the deferred build (fails in the reference's own objects.cpp too)."""
import sys, os, types, traceback, re, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
# stub matplotlib
mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
class _Any:
    def __getattr__(self, n): return _Any()
    def __call__(self, *a, **k): return _Any()
    def __iter__(self): return iter([_Any(), _Any()])
    def __getitem__(self, k): return _Any()
def _ga(n):
    if n.startswith("__"):
        raise AttributeError(n)
    return _Any()
for m in (mpl, plt):
    m.__getattr__ = _ga
mpl.__version__ = "3.8.0"; plt.__version__ = "3.8.0"
sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
import brian2_b200, brian2 as b
path = sys.argv[1]
name = os.path.basename(path)[:-3]
d = os.path.join(os.environ.get("TMPDIR", "/tmp"), "b200_example_" + name)
src = open(path).read()
b.set_device("b200", directory=d, build_on_run=False)
status = "ok"
t0 = time.time()
try:
    g = {"__name__": "__main__"}
    exec(compile(src, path, "exec"), g)
except NotImplementedError as ex:
    status = "after-run NotImplemented: " + str(ex)[:150]
except Exception as ex:
    status = "script stopped: %s: %s" % (type(ex).__name__, str(ex)[:150]); traceback.print_exc()
try:
    if not b.device.has_been_run and b.device.main_queue:
        b.device.build(directory=d, compile=True, run=False, with_output=False)
        status += " | BUILD OK"
    else:
        status += " | nothing to build"
except NotImplementedError as ex:
    status += " | build NotImplemented: " + str(ex)[:200]
except Exception as ex:
    status += " | BUILD FAILED %s: %s" % (type(ex).__name__, str(ex)[:300]); traceback.print_exc()
print("RESULT %s: %s (%.0f s)" % (name, status, time.time() - t0))
