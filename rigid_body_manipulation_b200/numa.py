"""NUMA placement for the host side of the end-to-end path (one process per GPU).

The `*_host` entry points stream from / to pinned host memory over PCIe.  On a two-socket host each GPU hangs off one
socket; a rank whose pinned buffers live on the other socket pays the inter-socket link on every byte, and eight ranks
doing so share that one link.  `bind_to_device` pins the calling process to the CPUs of the GPU's NUMA node and sets the
memory policy to prefer that node, so that every later allocation (pinned buffers included: they are first-touched by
the allocating thread) is node-local.  Call it BEFORE allocating the pinned buffers.

Everything here is best effort: on hosts that do not expose the topology (containers without sysfs, single-node VMs)
the functions report why and change nothing.
"""
from __future__ import annotations

import ctypes
import os

from . import _lib

_MPOL_PREFERRED = 1
_SYS_SET_MEMPOLICY = 238  # x86_64


def _parse_cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def device_topology(device: int) -> dict:
    """{'pci': bus id, 'numa_node': int or None, 'cpus': [...]} of a CUDA device (runtime numbering)."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(32)
    _lib.check(lib.rbm_device_pci_bus_id(int(device), buf, 32), "rbm_device_pci_bus_id")
    pci = buf.value.decode().lower()
    base = f"/sys/bus/pci/devices/{pci}"
    node, cpus = None, []
    try:
        with open(f"{base}/numa_node") as f:
            v = int(f.read().strip())
            node = v if v >= 0 else None
        with open(f"{base}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
    except OSError:
        pass
    return {"pci": pci, "numa_node": node, "cpus": cpus}


def bind_to_device(device: int) -> dict:
    """Pin this process to the CPUs local to `device` and prefer its NUMA node for memory.  Returns what was done."""
    info = device_topology(device)
    info["bound"] = False
    allowed = os.sched_getaffinity(0)
    info["prev_affinity"] = sorted(allowed)  # so that a caller can undo the CPU binding (restore_affinity)
    cpus = sorted(set(info["cpus"]) & allowed)
    if info["numa_node"] is None or not cpus:
        info["why"] = "topology not exposed (numa_node = -1 / no local_cpulist) or no local CPU in the allowed set"
        return info
    try:
        os.sched_setaffinity(0, cpus)
        info["bound"] = True
        info["n_cpus"] = len(cpus)
    except OSError as e:
        info["why"] = f"sched_setaffinity: {e}"
        return info
    try:  # memory policy: prefer the node (never fails allocations, unlike MPOL_BIND)
        node = info["numa_node"]
        nbits = max(64, node + 2)
        mask = (ctypes.c_ulong * ((nbits + 63) // 64))()
        mask[node // 64] |= 1 << (node % 64)
        rc = ctypes.CDLL(None, use_errno=True).syscall(_SYS_SET_MEMPOLICY, _MPOL_PREFERRED, mask, nbits)
        info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
    except Exception as e:  # pragma: no cover - platform dependent
        info["mempolicy"] = f"unavailable: {e}"
    return info


def restore_affinity(info: dict) -> None:
    """Undo the CPU part of `bind_to_device` (e.g. before forking CPU workers that should use every core)."""
    prev = info.get("prev_affinity")
    if info.get("bound") and prev:
        os.sched_setaffinity(0, prev)
