"""CPU / NUMA affinity of a GPU worker process.

On a two-socket box every GPU hangs off one socket.  A worker whose threads (and therefore whose pinned host buffers,
which are allocated on the NUMA node of the thread that creates them) live on the other socket pays for every
host -> device copy with a trip across the inter-socket link - with two or more workers sharing that link the 154 MB
per-step input copy of the ResNet bench went from 3 ms to 15-30 ms and no longer hid behind the step (end-to-end 36 ms
against 21.7 ms device time at N=2).  ``bind_to_gpu`` pins the calling thread (and the threads it creates afterwards) to
the CPUs that sysfs lists as local to the GPU's PCI device.  Disable with ``SPARKDL_NUMA_BIND=0``.
"""
import os

_done = {}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_locality(device_index):
    """``(numa_node, cpus)`` of a CUDA device from sysfs, or ``(None, None)`` when it cannot be determined."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        node = None
        try:
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        except (OSError, ValueError):
            pass
        return node, cpus
    except Exception:  # pylint: disable=broad-except
        return None, None


def bind_to_gpu(device_index):
    """Restrict the calling thread to the CPUs local to ``device_index``.  Returns a dict describing what was done
    (``{"bound": bool, "numa_node": .., "cpus": n}``); never raises."""
    key = (os.getpid(), device_index)
    if key in _done:
        return _done[key]
    info = {"bound": False, "numa_node": None, "cpus": 0}
    if os.environ.get("SPARKDL_NUMA_BIND", "1") != "0" and hasattr(os, "sched_setaffinity"):
        node, cpus = gpu_locality(device_index)
        try:
            allowed = os.sched_getaffinity(0)
            if cpus:
                cpus = cpus & allowed
            # only narrow when the GPU-local set is a strict, non-trivial subset of what we may use
            if cpus and len(cpus) < len(allowed):
                os.sched_setaffinity(0, cpus)
                info = {"bound": True, "numa_node": node, "cpus": len(cpus)}
            else:
                info = {"bound": False, "numa_node": node, "cpus": len(allowed)}
        except OSError:
            pass
    _done[key] = info
    return info
