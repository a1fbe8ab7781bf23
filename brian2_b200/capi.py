"""Thin ctypes binding of the C ABI exported by every built ``b200`` project
(``include/brian2_b200.h``).  No PyTorch, no fallback: if the library cannot be loaded or no
CUDA device is present the calls fail loudly."""
import ctypes
import os
import shutil

import numpy as np

__all__ = ["B200Library", "ABI_SYMBOLS"]

#: every symbol declared in include/brian2_b200.h
ABI_SYMBOLS = [
    "b200_run_main",
    "b200_last_error",
    "b200_last_run_time",
    "b200_last_run_completed_fraction",
    "b200_request_stop",
    "b200_set_comm",
    "b200_comm_rank",
    "b200_comm_world",
    "b200_set_option",
    "b200_get_counter",
    "b200_profiling",
    "b200_get_array_size",
    "b200_get_array",
    "b200_set_array",
    "b200_finalize",
]


#: int allgather(const void* send, void* recv, size_t nbytes_per_rank)
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


class B200Library:
    """One loaded project library.  ``fresh_copy=True`` loads a private copy of the file so that
    a second run of the same project starts from pristine static state (the reference gets this
    for free by starting a new process for every run)."""

    _copies = 0
    #: paths this process has already dlopen()ed.  glibc matches loaded objects by NAME and ctypes
    #: never dlclose()s, so opening a rebuilt library under a path that was loaded before would
    #: silently return the OLD code and static state: every repeat gets a private copy.
    _loaded_paths = set()

    def __init__(self, path, fresh_copy=False):
        path = os.path.abspath(path)
        if not os.path.exists(path):
            raise RuntimeError(f"b200 project library not found: {path} (build failed?)")
        if fresh_copy or path in B200Library._loaded_paths:
            B200Library._copies += 1
            base, ext = os.path.splitext(path)
            copy = f"{base}_run{B200Library._copies}{ext}"
            while copy in B200Library._loaded_paths:
                B200Library._copies += 1
                copy = f"{base}_run{B200Library._copies}{ext}"
            if os.path.exists(copy):
                os.unlink(copy)     # a new inode: never write into a file that may be mapped
            shutil.copy2(path, copy)
            path = copy
        B200Library._loaded_paths.add(path)
        self.path = path
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        missing = [s for s in ABI_SYMBOLS if not hasattr(self.lib, s)]
        if missing:
            raise RuntimeError(f"{path} does not export the b200 C ABI: missing {missing}")
        L = self.lib
        L.b200_run_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p)]
        L.b200_run_main.restype = ctypes.c_int
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_last_run_time.restype = ctypes.c_double
        L.b200_last_run_completed_fraction.restype = ctypes.c_double
        L.b200_request_stop.restype = None
        L.b200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_double]
        L.b200_set_option.restype = ctypes.c_int
        L.b200_get_counter.argtypes = [ctypes.c_char_p]
        L.b200_get_counter.restype = ctypes.c_double
        L.b200_profiling.argtypes = [
            ctypes.POINTER(ctypes.c_char_p),
            ctypes.POINTER(ctypes.c_double),
            ctypes.c_int,
        ]
        L.b200_profiling.restype = ctypes.c_int
        L.b200_get_array_size.argtypes = [ctypes.c_char_p]
        L.b200_get_array_size.restype = ctypes.c_longlong
        L.b200_get_array.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        L.b200_get_array.restype = ctypes.c_int
        L.b200_set_array.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        L.b200_set_array.restype = ctypes.c_int
        L.b200_finalize.restype = ctypes.c_int
        L.b200_set_comm.argtypes = [ctypes.c_int, ctypes.c_int, ALLGATHER_FN]
        L.b200_set_comm.restype = ctypes.c_int
        L.b200_comm_rank.restype = ctypes.c_int
        L.b200_comm_world.restype = ctypes.c_int
        self._allgather_cb = None   # keeps the ctypes callback alive

    # ------------------------------------------------------------------------------------------
    def run_main(self, args, stdout=None):
        argv = (ctypes.c_char_p * max(1, len(args)))(*[a.encode() for a in args])
        saved = None
        if stdout is not None:
            # the generated code prints with std::cout: redirect fd 1 for the duration of the call
            import sys

            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(stdout.fileno(), 1)
        try:
            return int(self.lib.b200_run_main(len(args), argv))
        finally:
            if saved is not None:
                os.dup2(saved, 1)
                os.close(saved)

    def set_comm(self, rank, world, allgather=None):
        """Multi-GPU: ``allgather(bytes) -> list of bytes (one per rank)`` is a Python callable
        (e.g. built on ``torch.distributed.all_gather_object``); it is only used outside the step
        loop."""
        cb = ALLGATHER_FN(0)
        if allgather is not None:
            def _cb(send, recv, nbytes):
                try:
                    parts = allgather(ctypes.string_at(send, nbytes))
                    if len(parts) != world or any(len(p) != nbytes for p in parts):
                        return 2
                    ctypes.memmove(recv, b"".join(parts), nbytes * world)
                    return 0
                except Exception:   # never let an exception cross the C boundary
                    import traceback

                    traceback.print_exc()
                    return 1

            cb = ALLGATHER_FN(_cb)
        self._allgather_cb = cb
        status = self.lib.b200_set_comm(int(rank), int(world), cb)
        if status != 0:
            raise RuntimeError(f"b200_set_comm({rank}, {world}) failed with status {status}")

    def last_error(self):
        msg = self.lib.b200_last_error()
        return msg.decode() if msg else ""

    def last_run_time(self):
        return float(self.lib.b200_last_run_time())

    def last_run_completed_fraction(self):
        return float(self.lib.b200_last_run_completed_fraction())

    def request_stop(self):
        self.lib.b200_request_stop()

    def set_option(self, key, value):
        if self.lib.b200_set_option(key.encode(), float(value)) != 0:
            raise KeyError(f"unknown b200 option '{key}'")

    def get_counter(self, key):
        return float(self.lib.b200_get_counter(key.encode()))

    def profiling(self, cap=1024):
        names = (ctypes.c_char_p * cap)()
        secs = (ctypes.c_double * cap)()
        n = self.lib.b200_profiling(names, secs, cap)
        return [(names[i].decode(), secs[i]) for i in range(n)]

    def get_array(self, name, dtype):
        nbytes = self.lib.b200_get_array_size(name.encode())
        if nbytes < 0:
            raise KeyError(f"unknown array '{name}'")
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        if nbytes and self.lib.b200_get_array(name.encode(), out.ctypes.data_as(ctypes.c_void_p), nbytes) != 0:
            raise RuntimeError(f"could not read array '{name}'")
        return out

    def set_array(self, name, values):
        values = np.ascontiguousarray(values)
        if self.lib.b200_set_array(name.encode(), values.ctypes.data_as(ctypes.c_void_p), values.nbytes) != 0:
            raise RuntimeError(f"could not set array '{name}'")

    def finalize(self):
        return int(self.lib.b200_finalize())
