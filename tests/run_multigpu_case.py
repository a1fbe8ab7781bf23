"""Run golden cases SHARDED over the ranks of a torchrun launch and compare with the reference's
golden vectors on rank 0.  Used by tests/test_parity_gpu.py::test_multi_gpu_* and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_multigpu_case.py cuba_1000 brunel_hetero
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(cases):
    import torch
    import torch.distributed as dist

    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("gloo")
    import brian2_b200  # noqa: F401
    import brian2 as b

    import models
    from golden.make_golden import CASES

    failures = []
    if cases and cases[0] == "--sharded":
        failures = sharded(b, models, CASES, cases[1:], rank, dist)
        cases = []
    for case in cases:
        model, kwds = CASES[case]
        d = os.path.join(ROOT, "brian2_b200", "_prebuilt", f"mgpu_{case}_r{rank}")
        objs, res = models.run_model(b, model, "b200", d, **kwds)
        assert int(b.device.counter("grid")) > 0
        events = b.device.counter("events")
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"{case}.npz"))
        if rank == 0:
            for key in gold.files:
                g, r = gold[key], res[key]
                exact = g.dtype.kind in "iu" or key.endswith("_t") or case not in ("cobahh_1000", "stdp_1000")
                if g.shape != r.shape:
                    failures.append(f"{case}:{key} shape {r.shape} != {g.shape}")
                elif exact and not np.array_equal(g, r):
                    failures.append(f"{case}:{key} differs ({int(np.sum(g != r))} of {g.size})")
                elif not exact and not np.allclose(r, g, rtol=1e-9, atol=1e-15):
                    failures.append(f"{case}:{key} outside rtol 1e-9")
        print(f"[rank {rank}] {case}: local synaptic events {int(events)}", flush=True)
        dist.barrier()
    if rank == 0:
        print("MULTIGPU", "FAIL" if failures else "OK", failures, flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


def sharded(b, models, CASES, cases, rank, dist):
    """Sharded construction: N ranks against ONE rank (rank 0 re-runs the script on its own)."""
    failures = []
    for case in cases:
        model, kwds = CASES[case]
        runs = {}
        for mode in ("ranks", "single"):
            if mode == "single" and rank != 0:
                continue
            d = os.path.join(ROOT, "brian2_b200", "_prebuilt", f"mgpu_sharded_{mode}_{case}_r{rank}")
            objs, res = models.run_model(
                b, model, "b200", d,
                prefs_update={"devices.b200.construction": "sharded",
                              "devices.b200.multi_gpu": mode == "ranks"}, **kwds)
            for key, obj in objs.items():
                if isinstance(obj, b.Synapses):
                    res[f"{key}_i"] = np.asarray(obj.i[:])
                    res[f"{key}_j"] = np.asarray(obj.j[:])
                    if "delay" in obj.variables and len(np.atleast_1d(obj.delay_[:])) == len(obj):
                        res[f"{key}_delay"] = np.asarray(obj.delay_[:])
            runs[mode] = res
            print(f"[rank {rank}] {case} ({mode}): {len(res.get('spikes_i', []))} spikes, "
                  f"local events {int(b.device.counter('events'))}", flush=True)
        b.prefs["devices.b200.construction"] = "reference"
        b.prefs["devices.b200.multi_gpu"] = True
        if rank == 0:
            for key, ref in runs["single"].items():
                got = runs["ranks"][key]
                if key == "last_run_time":
                    continue
                if ref.shape != got.shape:
                    failures.append(f"{case}:{key} shape {got.shape} != {ref.shape}")
                elif not np.array_equal(ref, got):
                    failures.append(f"{case}:{key} differs ({int(np.sum(ref != got))} of {ref.size})")
        dist.barrier()
    return failures


if __name__ == "__main__":
    main(sys.argv[1:])
