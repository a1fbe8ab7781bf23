"""Host-side logic of multi-GPU runs of the ``b200`` device (one process per GPU of one box).

Partition: rank ``r`` owns a contiguous block of the neurons of every
group -- their state, thresholding, reset and monitors -- and every synapse whose postsynaptic
neuron lies in the block, so all synaptic effects are local writes.  The only data that crosses
GPUs inside the step loop are the spike lists: each rank stores its segment of every step's list
directly into the peers' spike rings over NVLink (CUDA-IPC mapped peer memory, see
``csrc/b200_runtime.cuh``: ``publish_owned`` stores tagged words, ``view_build`` / ``view_load``
spin on the tag of the word they need -- flag and data travel together).

This module holds what happens on the host, outside the loop:

* the communicator used to exchange the CUDA IPC handles (``Communicator``; by default built on
  ``torch.distributed`` -- plumbing only, never in the step loop);
* the same partition arithmetic as the device code (``rank_range``);
* the merge of the per-rank results after a run (``merge_value`` and friends): per-neuron arrays
  by owner block, synaptic arrays by the owner of the postsynaptic neuron, spike monitors by a
  stable (t, rank) merge, state monitors by column owner, rate monitors by summing the per-rank
  spike counts before the reference's ``1.0*n/dt/N`` (ratemonitor.cpp:33).
"""
import pickle

import numpy as np

__all__ = ["Communicator", "TorchCommunicator", "rank_range", "owner_of", "merge_value"]


def rank_range(n, rank, world):
    """Block of ``[0, n)`` owned by ``rank`` -- identical to ``b200::rank_range`` (device) and
    ``EventSpace::rank_range_host``: blocks of ``ceil(n/world)`` rounded up to a multiple of 32."""
    per = -(-int(n) // int(world))
    per = (per + 31) & ~31
    lo = min(int(n), rank * per)
    hi = min(int(n), lo + per)
    return lo, hi


def owner_of(indices, n, world):
    """Rank that owns each element index of a group of size ``n``."""
    per = -(-int(n) // int(world))
    per = (per + 31) & ~31
    return np.minimum(np.asarray(indices, dtype=np.int64) // per, world - 1).astype(np.int32)


class Communicator:
    """Minimal interface: ``rank``, ``world``, ``allgather(bytes) -> [bytes, ...]`` (rank order)."""

    rank = 0
    world = 1

    def allgather(self, payload):
        return [payload]

    def allgather_object(self, obj):
        return [pickle.loads(p) for p in self.allgather(pickle.dumps(obj, protocol=4))]

    def barrier(self):
        self.allgather(b"\0")


class TorchCommunicator(Communicator):
    """Out-of-band channel over an initialised ``torch.distributed`` process group (gloo or nccl)."""

    def __init__(self):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("torch.distributed is not initialised")
        self._dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    def allgather(self, payload):
        out = [None] * self.world
        self._dist.all_gather_object(out, bytes(payload))
        return out


def default_communicator():
    """``TorchCommunicator`` if a process group with more than one rank exists, else single rank."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchCommunicator()
    except ImportError:
        pass
    return Communicator()


# ---------------------------------------------------------------------------------------------
# merging per-rank results
# ---------------------------------------------------------------------------------------------
def merge_by_block(parts, n, world):
    """Per-element array of a group of ``n`` elements: take every block from its owner."""
    out = np.array(parts[0], copy=True)
    for r in range(world):
        lo, hi = rank_range(n, r, world)
        out[lo:hi] = parts[r][lo:hi]
    return out


def merge_by_owner(parts, owners):
    """Array whose element ``k`` is valid on rank ``owners[k]`` (synaptic variables)."""
    out = np.array(parts[0], copy=True)
    for r in range(len(parts)):
        sel = owners == r
        out[sel] = parts[r][sel]
    return out


def merge_spike_records(ts, columns):
    """Spike/event monitor: every rank recorded the events of its own neurons in (t, i) order.
    The reference appends, per step, the ids in ascending order (spikemonitor.cpp:35-47); rank
    blocks are ascending id ranges, so a stable sort of the concatenation by t restores it.

    ``ts``: list (per rank) of the recorded times; ``columns``: dict name -> list (per rank)."""
    t_all = np.concatenate(ts)
    order = np.argsort(t_all, kind="stable")
    merged = {name: np.concatenate(parts)[order] for name, parts in columns.items()}
    return t_all[order], merged


def merge_state_columns(parts, indices, n, world):
    """State monitor (steps x n_rec): column j is valid on the owner of ``indices[j]``."""
    out = np.array(parts[0], copy=True)
    own = owner_of(indices, n, world)
    for r in range(world):
        sel = own == r
        if out.ndim == 2 and out.shape[1] == len(indices):
            out[:, sel] = parts[r][:, sel]
        else:   # Brian exposes recorded values as (n_rec, steps)
            out[sel, :] = parts[r][sel, :]
    return out


def merge_rate(counts, dt, n_source):
    """Rate monitor: ``counts`` = per-rank spike counts per step (stored in ``rate`` by the
    device on several GPUs); the reference's formula is applied once to the exact total."""
    total = np.sum(np.stack(counts), axis=0)
    return 1.0 * total / dt / n_source


def sharded_synapse_order(pre_parts, post_parts):
    """Sharded construction: every rank created the synapses of its own postsynaptic neurons,
    each rank's list in (pre, post) order.  Returns the permutation that puts the concatenation
    of the ranks' lists into the order a single-GPU run creates the same synapses in:
    (pre ascending, post ascending), equal pairs (multiple synapses) in creation order."""
    pre_all = np.concatenate([np.asarray(p) for p in pre_parts])
    post_all = np.concatenate([np.asarray(p) for p in post_parts])
    return np.lexsort((post_all, pre_all))      # lexsort is stable


def merge_value(kind, parts, **kw):
    return {
        "block": merge_by_block,
        "owner": merge_by_owner,
        "state": merge_state_columns,
        "rate": merge_rate,
    }[kind](parts, **kw)
