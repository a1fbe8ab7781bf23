"""Run the REFERENCE's own standalone-compatible tests on the ``b200`` device
(`brian2.test(test_standalone='b200')`, brian2/tests/__init__.py:391-433 -- the conformance
harness the reference provides for new standalone devices; unsupported features raise
NotImplementedError and are reported as skips, brian2/conftest.py:107-127).

    python tests/tools/run_reference_suite.py [--files test_synapses test_monitor ...] [-n WORKERS]
                                              [--junit OUT.xml] [--summary OUT.json]

Without a GPU every test that reaches `device.run()` fails with "Project run failed" (there is no
CPU fallback): useful as a code-generation + nvcc pre-screen; the summary separates those from
genuine failures."""
import argparse
import json
import os
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

#: the hot-path files named by SURVEY.md section 4 / 8c
DEFAULT_FILES = ["test_synapses", "test_refractory", "test_thresholder", "test_monitor", "test_subgroup",
                 "test_spikegenerator", "test_poissongroup"]


def summarise(junit_paths):
    counts = {"passed": 0, "skipped": 0, "failed": 0, "no_gpu": 0}
    failures, skips = [], []
    for path in junit_paths:
        if not os.path.exists(path):
            continue
        for case in ET.parse(path).getroot().iter("testcase"):
            name = f"{case.get('classname', '').split('.')[-1]}::{case.get('name')}"
            bad = case.find("failure") if case.find("failure") is not None else case.find("error")
            skip = case.find("skipped")
            if bad is not None:
                text = (bad.get("message") or "") + (bad.text or "")
                if "no CUDA device available" in text or "Project run failed" in text:
                    counts["no_gpu"] += 1
                else:
                    counts["failed"] += 1
                    last = [ln for ln in text.strip().splitlines() if ln.strip()][-1:] or [""]
                    failures.append({"test": name, "why": (bad.get("message") or last[0])[:300]})
            elif skip is not None:
                counts["skipped"] += 1
                skips.append({"test": name, "why": (skip.get("message") or "")[:200]})
            else:
                counts["passed"] += 1
    return counts, failures, skips


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", nargs="*", default=DEFAULT_FILES)
    ap.add_argument("-n", type=int, default=0, help="pytest-xdist workers (0: in-process)")
    ap.add_argument("--junit", default=os.path.join(ROOT, "gpurun_out", "reference_suite"))
    ap.add_argument("--summary", default=os.path.join(ROOT, "gpurun_out", "reference_suite_summary.json"))
    ap.add_argument("-k", default=None, help="extra pytest -k expression (and-ed)")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.junit), exist_ok=True)

    import brian2_b200  # noqa: F401
    import brian2
    import pytest

    # brian2.test() calls pytest.main twice for a standalone device (single / multiple run
    # statements): give every call its own junit file
    calls = {"n": 0}
    real_main = pytest.main

    def main_with_junit(argv, plugins=None):
        calls["n"] += 1
        return real_main(list(argv) + [f"--junitxml={args.junit}_{calls['n']}.xml"], plugins=plugins)

    pytest.main = main_with_junit
    expr = " or ".join(args.files)
    if args.k:
        expr = f"({expr}) and ({args.k})"
    os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")
    extra = ["-k", expr, "-q", "-p", "no:cacheprovider", "-p", "brian2_b200.pytest_plugin"]
    if args.n:
        extra += ["-n", str(args.n)]
    try:
        brian2.test(codegen_targets=[], test_codegen_independent=False, test_standalone="b200",
                    fail_for_not_implemented=False, reset_preferences=False, additional_args=extra)
    finally:
        pytest.main = real_main
    counts, failures, skips = summarise([f"{args.junit}_{i}.xml" for i in range(1, calls["n"] + 1)])
    out = {"files": args.files, "counts": counts, "failures": failures, "skips": skips}
    with open(args.summary, "w") as f:
        json.dump(out, f, indent=1)
    print("REFERENCE SUITE ON b200:", json.dumps(counts))
    for item in failures:
        print("  FAILED ", item["test"], "--", item["why"])
    return 0 if counts["failed"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
