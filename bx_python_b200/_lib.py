"""
ctypes binding of libbxb200.so (C ABI: include/bxb200.h).  No CPU fallback: load() raises when the library is
missing, and every compute call raises RuntimeError when there is no CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbxb200.so")

HOST, DEVICE = 0, 1
ERR_MISMATCH = -3

vp, i32, i64, cint = C.c_void_p, C.c_int32, C.c_int64, C.c_int
pvp = C.POINTER(C.c_void_p)
pi32, pi64 = C.POINTER(i32), C.POINTER(i64)

# name -> argtypes (restype is int status unless listed in _RESTYPE)
SIGNATURES = {
    "bxg_init": [cint],
    "bxg_device_count": [C.POINTER(cint)],
    "bxg_device_info": [C.c_char_p, cint, C.POINTER(cint), pi64, C.POINTER(cint), C.POINTER(cint)],
    "bxg_device_pci_bus_id": [C.c_char_p, cint],
    "bxg_last_error": [],
    "bxg_version": [],
    "bxg_sync": [],
    "bxg_launch_count": [pi64],
    "bxg_launch_count_reset": [],
    "bxg_host_alloc": [i64, pvp],
    "bxg_host_free": [vp],
    "bxg_dev_alloc": [i64, pvp],
    "bxg_dev_free": [vp],
    "bxg_dev_memset": [vp, cint, i64],
    "bxg_memcpy_h2d": [vp, vp, i64],
    "bxg_memcpy_d2h": [vp, vp, i64],
    "bxg_timer_create": [pvp],
    "bxg_timer_free": [vp],
    "bxg_timer_start": [vp],
    "bxg_timer_stop": [vp],
    "bxg_timer_elapsed_ms": [vp, C.POINTER(C.c_float)],
    "bxg_l2_flush": [],
    "bxg_copy_probe": [i64, cint, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "bxg_profile_enable": [cint],
    "bxg_profile_report": [C.c_char_p, i64],
    "bxg_bits_create": [i32, i32, pvp],
    "bxg_bits_free": [vp],
    "bxg_bits_geometry": [vp, pi32, pi32, pi32],
    "bxg_bits_clone": [vp, pvp],
    "bxg_bits_clear": [vp],
    "bxg_bits_set_ranges": [vp, vp, vp, i64, cint],
    "bxg_bits_set_ranges_multi": [pvp, i32, vp, vp, vp, i64, cint],
    "bxg_bits_set_bits": [vp, vp, i64, cint, cint],
    "bxg_bits_read": [vp, vp, i64, vp, cint],
    "bxg_bits_and": [vp, vp],
    "bxg_bits_or": [vp, vp],
    "bxg_bits_xor": [vp, vp],
    "bxg_bits_not": [vp],
    "bxg_bits_and_count": [vp, vp, pi64],
    "bxg_bits_binop_batch": [cint, pvp, pvp, i32, vp],
    "bxg_bits_count_ranges": [vp, vp, vp, i64, vp, cint, cint],
    "bxg_bits_count_all": [vp, pi64],
    "bxg_bits_count_all_multi": [pvp, i32, vp, i64, cint],
    "bxg_group_stats_i32": [vp, vp, i64, i32, i32, vp, cint],
    "bxg_bits_count_ranges_multi": [pvp, i32, vp, vp, vp, i64, vp, cint, cint],
    "bxg_bits_next": [vp, i32, i32, cint, pi32],
    "bxg_bits_runs_count": [vp, pi64],
    "bxg_bits_runs_fetch": [vp, vp, vp, i64],
    "bxg_bits_runs_in_ranges": [vp, vp, vp, i64, cint, cint, vp, pi64],
    "bxg_bits_runs_in_ranges_fetch": [vp, vp, vp, i64],
    "bxg_bits_states": [vp, vp],
    "bxg_bits_export_words": [vp, vp],
    "bxg_bits_import_words": [vp, vp],
    "bxg_bits_device_words": [vp, pvp, pi64],
    "bxg_itree_create": [pvp],
    "bxg_itree_free": [vp],
    "bxg_itree_build": [vp, vp, vp, vp, i64, i32, cint],
    "bxg_itree_size": [vp, pi64, pi32],
    "bxg_itree_order": [vp, vp, vp],
    "bxg_itree_find": [vp, vp, vp, vp, i64, cint, pi64],
    "bxg_itree_fetch": [vp, vp, vp],
    "bxg_set_find_mode": [cint],
    "bxg_itree_find_host": [vp, vp, vp, vp, i64, pvp, pvp, pi64],
    "bxg_itree_find_host32": [vp, vp, vp, vp, i64, pvp, pvp, pi64],
    "bxg_itree_find_small": [vp, vp, vp, vp, i32, pvp, pvp, pi64],
    "bxg_itree_find1": [vp, i32, i32, i32, pvp],
    "bxg_set_find_server": [cint],
    "bxg_find_server_stats": [pi64, pi64, pi32],
    "bxg_itree_result_dev": [vp, pvp, pvp, pi64, pi64],
    "bxg_itree_count": [vp, vp, vp, vp, i64, cint, vp, pi64],
    "bxg_itree_neighbors": [vp, vp, vp, vp, vp, i64, cint, cint, pi64],
    "bxg_itree_join": [vp, vp, vp, vp, i64, vp, vp, i32, cint, pi64],
    "bxg_itree_join_fetch": [vp, vp, vp],
    "bxg_scores_create": [vp, i64, i32, cint, pvp],
    "bxg_scores_free": [vp],
    "bxg_aggregate": [vp, vp, vp, vp, i64, cint, vp, vp, vp, vp, vp],
    "bxg_aggregate_multi": [pvp, pvp, i32, vp, vp, vp, i64, cint, vp, vp, vp, vp, vp],
    "bxg_scores_alloc": [i64, i32, C.c_float, pvp],
    "bxg_scores_info": [vp, pi64, pi32, C.POINTER(C.c_float)],
    "bxg_scores_reserve": [vp, i64],
    "bxg_scores_set_spans": [vp, vp, vp, vp, i64, cint],
    "bxg_scores_write": [vp, i64, vp, i64, cint],
    "bxg_scores_get": [vp, vp, i64, vp, cint],
    "bxg_scores_get_range": [vp, i64, i64, vp],
    "bxg_scores_device": [vp, pvp, pi64],
    "bxg_summarize": [vp, vp, vp, i64, cint, C.c_uint32, C.c_uint32, i32, vp, vp, vp, vp, vp],
    "bxg_comm_unique_id": [C.c_char_p],
    "bxg_comm_init": [C.c_char_p, cint, cint],
    "bxg_comm_allreduce_i64": [vp, i64],
    "bxg_comm_allreduce_max_f64": [vp, i64],
    "bxg_comm_allreduce_i64_dev": [vp, i64],
    "bxg_comm_barrier": [],
    "bxg_comm_destroy": [],
}
_RESTYPE = {"bxg_last_error": C.c_char_p, "bxg_version": C.c_char_p, "bxg_itree_find1": i64}

_lib = None
_initialised = False


def load():
    """dlopen libbxb200.so and declare every entry point (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m bx_python_b200.build` (needs nvcc). "
            "bx_python_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = _RESTYPE.get(name, cint)
    _lib = L
    return L


def last_error() -> str:
    return load().bxg_last_error().decode(errors="replace")


def check(rc: int):
    if rc == 0:
        return
    msg = last_error()
    if rc == ERR_MISMATCH:
        raise ValueError(msg)
    raise RuntimeError(f"libbxb200: {msg} (status {rc})")


def lib():
    """Library handle with the CUDA context bound (device from BXB200_DEVICE / LOCAL_RANK, default 0)."""
    global _initialised
    L = load()
    if not _initialised:
        dev = int(os.environ.get("BXB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        check(L.bxg_init(dev))
        _initialised = True
    return L


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def as_i32(x) -> np.ndarray:
    a = np.asarray(x)
    if a.dtype != np.int32:
        if a.dtype.kind not in "iu":
            a = a.astype(np.int64)          # float -> truncation toward zero, like Cython's int coercion
        if a.size and (a.max(initial=0) > 0x7FFFFFFF or a.min(initial=0) < -0x80000000):
            raise OverflowError("value too large to convert to int")
        a = a.astype(np.int32)
    return np.ascontiguousarray(a)


def sync():
    check(lib().bxg_sync())


def launch_count() -> int:
    n = i64()
    check(lib().bxg_launch_count(C.byref(n)))
    return n.value


def profile_enable(on: bool):
    check(lib().bxg_profile_enable(1 if on else 0))


def profile_report() -> dict:
    """-> {kernel: (launches, total_ms)} for every launch since profile_enable(True)."""
    buf = C.create_string_buffer(1 << 16)
    check(lib().bxg_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split("\t")
        out[name] = (int(n), float(ms))
    return out


def l2_flush():
    check(lib().bxg_l2_flush())


def device_info() -> dict:
    name = C.create_string_buffer(256)
    sm, mem, maj, mnr = cint(), i64(), cint(), cint()
    check(lib().bxg_device_info(name, 256, C.byref(sm), C.byref(mem), C.byref(maj), C.byref(mnr)))
    return {"name": name.value.decode(), "sm_count": sm.value, "total_mem": mem.value, "cc": (maj.value, mnr.value)}


def _numa_nodes() -> dict:
    """{node: set(cpus)} from sysfs (empty when the kernel exposes no NUMA topology)."""
    import glob
    nodes = {}
    for d in glob.glob("/sys/devices/system/node/node[0-9]*"):
        try:
            cpus = set()
            for part in open(os.path.join(d, "cpulist")).read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            nodes[int(d.rsplit("node", 1)[1])] = cpus
        except (OSError, ValueError):
            continue
    return nodes


def _probe_numa_node(allowed) -> str:
    """sysfs does not say which NUMA node the GPU hangs off (virtualised boxes report -1): measure it.  For every node
    with usable CPUs, run on that node, allocate pinned memory there (cudaMallocHost places pages on the caller's node)
    and time device-to-host copies; stay on the fastest node if it is clearly (> 10 %) faster than the slowest."""
    cand = {k: c & allowed for k, c in _numa_nodes().items() if c & allowed}
    if len(cand) < 2:
        return "NUMA affinity unknown and fewer than two nodes to probe"
    import time
    L = lib()
    n = 16 << 20
    dev = DeviceBuffer(np.zeros(n, np.int32))
    rates = {}
    for k, cpus in sorted(cand.items()):
        os.sched_setaffinity(0, cpus)
        pin = PinnedArray(n, np.int32)
        pin.array[:] = 0
        check(L.bxg_memcpy_d2h(ptr(pin.array), dev.ptr, n * 4))
        sync()
        t0 = time.perf_counter()
        for _ in range(3):
            check(L.bxg_memcpy_d2h(ptr(pin.array), dev.ptr, n * 4))
        sync()
        rates[k] = 3 * n * 4 / (time.perf_counter() - t0) / 1e9
        del pin
    best = max(rates, key=rates.get)
    shown = ", ".join(f"node {k}: {v:.1f} GB/s" for k, v in sorted(rates.items()))
    if rates[best] > 1.1 * min(rates.values()):
        os.sched_setaffinity(0, cand[best])
        return f"probed D2H ({shown}): bound to NUMA node {best} ({len(cand[best])} CPUs)"
    os.sched_setaffinity(0, allowed)
    return f"probed D2H ({shown}): no clear winner, not bound"


def bind_to_gpu_numa_node() -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated afterwards
    (first touch) and the copy threads are local to the GPU's PCIe root.  Matters when several ranks share a two-socket
    host (measured: 39.6 instead of 57 GB/s device-to-host from the far socket).  The node comes from sysfs, or, where
    sysfs reports none, from a short copy-rate probe.  Returns a short description; never raises."""
    allowed = None
    try:
        allowed = os.sched_getaffinity(0)
        buf = C.create_string_buffer(64)
        check(lib().bxg_device_pci_bus_id(buf, 64))
        bus = buf.value.decode().lower()
        try:
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        except (OSError, ValueError):
            node = -1
        if node < 0:
            return f"{bus}: " + _probe_numa_node(allowed)
        cpus = _numa_nodes().get(node, set()) & allowed
        if not cpus:
            return f"{bus}: NUMA node {node} has no usable CPU"
        os.sched_setaffinity(0, cpus)
        return f"{bus}: bound to NUMA node {node} ({len(cpus)} CPUs)"
    except Exception as e:                             # noqa: BLE001 -- a tuning aid must never take the run down
        try:
            if allowed:
                os.sched_setaffinity(0, allowed)
        except OSError:
            pass
        return f"not bound ({type(e).__name__}: {e})"


class Timer:
    """CUDA-event timer on the library stream."""

    def __init__(self):
        self._h = vp()
        check(lib().bxg_timer_create(C.byref(self._h)))

    def start(self):
        check(lib().bxg_timer_start(self._h))

    def stop(self):
        check(lib().bxg_timer_stop(self._h))

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        check(lib().bxg_timer_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.bxg_timer_free(self._h)
            self._h = None


class PinnedArray:
    """numpy view over cudaMallocHost memory (for H2D/D2H at full PCIe rate)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = (shape,) if np.isscalar(shape) else tuple(shape)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = vp()
        check(lib().bxg_host_alloc(max(nbytes, 1), C.byref(self._p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        if getattr(self, "_p", None) and _lib is not None:
            self.array = None
            _lib.bxg_host_free(self._p)
            self._p = None


class DeviceBuffer:
    """Raw device allocation holding a copy of a host array (bench.py uses it for HBM-resident inputs)."""

    def __init__(self, host: np.ndarray):
        host = np.ascontiguousarray(host)
        self.nbytes, self.dtype, self.shape = host.nbytes, host.dtype, host.shape
        self._p = vp()
        check(lib().bxg_dev_alloc(max(self.nbytes, 1), C.byref(self._p)))
        if self.nbytes:
            check(lib().bxg_memcpy_h2d(self._p, ptr(host), self.nbytes))
            sync()

    @property
    def ptr(self):
        return self._p

    def __del__(self):
        if getattr(self, "_p", None) and _lib is not None:
            _lib.bxg_dev_free(self._p)
            self._p = None
