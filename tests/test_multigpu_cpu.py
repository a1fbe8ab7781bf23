"""Host-side logic of multi-GPU runs, exercised with world_size-2/3 gloo process groups on CPU:
the communicator, the partition arithmetic shared with the device code, and the merge of the
per-rank results.  (The data path itself -- NVLink peer stores -- needs GPUs: tests/test_parity_gpu.py.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_rank_range_is_a_partition():
    from brian2_b200.multigpu import owner_of, rank_range

    for n in (0, 1, 31, 32, 33, 1000, 4000, 256000, 1000003):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                lo, hi = rank_range(n, r, world)
                assert lo == covered and hi >= lo
                assert lo % 32 == 0 or lo == n
                covered = hi
            assert covered == n
            if n:
                idx = np.arange(n)
                own = owner_of(idx, n, world)
                for r in range(world):
                    lo, hi = rank_range(n, r, world)
                    assert np.all(own[lo:hi] == r)


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from brian2_b200 import multigpu as mg

    comm = mg.TorchCommunicator()
    assert (comm.rank, comm.world) == (rank, world)
    # --- raw allgather (what the C ABI callback uses for the CUDA IPC handles)
    got = comm.allgather(bytes([rank]) * 64)
    assert got == [bytes([r]) * 64 for r in range(world)]
    comm.barrier()

    # --- a synthetic "global truth" every rank can construct, and this rank's share of it
    rng = np.random.RandomState(5)
    N, S, steps = 1000, 7000, 40
    v_true = rng.rand(N)
    post = np.sort(rng.randint(0, N, S)).astype(np.int32)
    w_true = rng.rand(S)
    spikes = [np.sort(rng.choice(N, rng.randint(0, 60), replace=False)) for _ in range(steps)]
    lo, hi = mg.rank_range(N, rank, world)

    # per-neuron array: valid in the owned block only
    v_local = np.full(N, -1.0)
    v_local[lo:hi] = v_true[lo:hi]
    v = mg.merge_by_block(comm.allgather_object(v_local), N, world)
    assert np.array_equal(v, v_true)

    # synaptic array: valid where the postsynaptic neuron is owned
    own = mg.owner_of(post, N, world)
    w_local = np.where(own == rank, w_true, -1.0)
    w = mg.merge_by_owner(comm.allgather_object(w_local), own)
    assert np.array_equal(w, w_true)

    # spike monitor: every rank records its own neurons, step by step
    t_loc, i_loc = [], []
    for k, ids in enumerate(spikes):
        mine = ids[(ids >= lo) & (ids < hi)]
        t_loc += [k * 1e-4] * len(mine)
        i_loc += list(mine)
    parts = comm.allgather_object({"t": np.array(t_loc), "i": np.array(i_loc, dtype=np.int32)})
    t_all, cols = mg.merge_spike_records([p["t"] for p in parts], {"i": [p["i"] for p in parts]})
    t_true = np.concatenate([[k * 1e-4] * len(ids) for k, ids in enumerate(spikes)])
    i_true = np.concatenate(spikes).astype(np.int32)
    assert np.array_equal(t_all, t_true) and np.array_equal(cols["i"], i_true)

    # rate monitor: per-rank counts -> the reference's formula on the exact total
    counts = np.array([np.sum((ids >= lo) & (ids < hi)) for ids in spikes], dtype=float)
    rate = mg.merge_rate(comm.allgather_object(counts), 1e-4, N)
    rate_true = np.array([1.0 * len(ids) / 1e-4 / N for ids in spikes])
    assert np.array_equal(rate, rate_true)

    # state monitor: column j valid on the owner of indices[j]
    indices = np.array([1, 10, 100, 700, 999])
    trace_true = rng.rand(steps, len(indices))
    own_i = mg.owner_of(indices, N, world)
    trace_local = np.where(own_i[None, :] == rank, trace_true, np.nan)
    trace = mg.merge_state_columns(comm.allgather_object(trace_local), indices, N, world)
    assert np.array_equal(trace, trace_true)
    # sharded construction: every rank holds the synapses of ITS postsynaptic neurons, in
    # (pre, post) order; the merged object has the single-GPU order
    pre_g = np.sort(rng.randint(0, N, S)).astype(np.int32)
    post_g = rng.randint(0, N, S).astype(np.int32)
    order_g = np.lexsort((post_g, pre_g))
    pre_g, post_g = pre_g[order_g], post_g[order_g]
    delay_g = rng.rand(S)
    mine = mg.owner_of(post_g, N, world) == rank
    parts = comm.allgather_object({"pre": pre_g[mine], "post": post_g[mine], "delay": delay_g[mine]})
    order = mg.sharded_synapse_order([p["pre"] for p in parts], [p["post"] for p in parts])
    assert np.array_equal(np.concatenate([p["pre"] for p in parts])[order], pre_g)
    assert np.array_equal(np.concatenate([p["post"] for p in parts])[order], post_g)
    pairs = pre_g.astype(np.int64) * N + post_g
    unique = np.concatenate([[True], pairs[1:] != pairs[:-1]]) & np.concatenate([pairs[1:] != pairs[:-1], [True]])
    merged_delay = np.concatenate([p["delay"] for p in parts])[order]
    assert np.array_equal(merged_delay[unique], delay_g[unique])
    with open(os.path.join(tmpdir, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_merge_over_gloo(world, tmp_path):
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_set_comm_through_the_c_abi(brian):
    """b200_set_comm / b200_comm_rank / b200_comm_world work without a GPU (no CUDA call)."""
    import ctypes

    import __graft_entry__ as ge
    from brian2_b200.capi import B200Library

    directory, _ = ge.build_project("cuba_1000", directory=os.path.join(ge.PREBUILT, "cpu_cuba_1000"))
    lib = B200Library(os.path.join(directory, "libb200_project.so"), fresh_copy=True)
    assert lib.lib.b200_comm_world() == 1
    calls = []

    def allgather(payload):
        calls.append(payload)
        return [payload, payload]

    lib.set_comm(1, 2, allgather)
    assert (lib.lib.b200_comm_rank(), lib.lib.b200_comm_world()) == (1, 2)
    with pytest.raises(RuntimeError):
        lib.set_comm(5, 2, allgather)      # rank outside the world
    with pytest.raises(RuntimeError):
        lib.set_comm(0, 9, allgather)      # more ranks than GPUs of one box
