"""CPU placement for a process that drives one GPU.

The head step is tens of microseconds: where the launching thread and its pinned host buffers live matters.  Zero-copy
reads of q | k from pinned host memory ran at 44 GB/s on one box and at 21 GB/s on another of the same pool (the prep
launch took 6.7 us against 13 us, profiles/r02_notes.md section 12) -- the difference between a host buffer on the GPU's
own NUMA node and one behind the socket interconnect.  `bind_cpu_to_device()` restricts the calling process to the CPUs
NVML reports as local to the GPU (pinned allocations made afterwards are first-touched there); `restore()` undoes it for
every thread of the process.  Best effort: without NVML / sysfs information, or in a cpuset that excludes those CPUs,
nothing changes and the returned record says so.
"""
import os


def _device_bus_id(index):
    import torch
    p = torch.cuda.get_device_properties(index)
    return "%08x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)


def _local_cpus(bus_id):
    """CPUs local to the PCI device `bus_id` ('dddddddd:bb:dd.0'): NVML first, then sysfs.  Empty set if unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        if cpus:
            return cpus, "nvml"
    except Exception:
        pass
    try:
        short = bus_id[-12:] if len(bus_id) > 12 else bus_id                     # sysfs uses a 4-digit domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % short.lower()).read())
        if node >= 0:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            return cpus, "sysfs node %d" % node
    except Exception:
        pass
    return set(), "unknown"


def bind_cpu_to_device(index=0):
    """Restrict the calling process to the CPUs local to CUDA device `index`.  Returns a record for logs:
    {'bound': bool, 'source': ..., 'cpus': n, 'previous': [...]} -- pass it to restore()."""
    rec = {"bound": False, "source": "unsupported", "cpus": 0, "previous": None}
    if not hasattr(os, "sched_setaffinity"):
        return rec
    try:
        prev = os.sched_getaffinity(0)
        rec["previous"] = sorted(prev)
        local, src = _local_cpus(_device_bus_id(index))
        rec["source"] = src
        want = local & prev
        if want and want != prev:
            os.sched_setaffinity(0, want)
            rec["bound"] = True
        rec["cpus"] = len(want) if want else len(prev)
    except Exception as e:                                                       # never fatal: placement is an optimisation
        rec["source"] = "error: %s" % e
    return rec


def restore(rec):
    """Give every thread of this process the affinity it had before bind_cpu_to_device()."""
    if not rec or not rec.get("bound") or not rec.get("previous"):
        return
    prev = set(rec["previous"])
    try:
        tids = [int(t) for t in os.listdir("/proc/self/task")]
    except Exception:
        tids = [0]
    for t in tids:
        try:
            os.sched_setaffinity(t, prev)
        except Exception:
            pass
