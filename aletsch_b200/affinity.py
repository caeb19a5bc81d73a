"""NUMA placement of a rank next to its GPU.

The end-to-end path moves ~1.4 GB per step and GPU between pinned host memory and the device.  Pinned pages are allocated on
the NUMA node of the thread that asks for them, and a rank started by `torch.distributed.run` runs wherever the scheduler puts
it: with eight ranks on a two-socket box about half of the buffers end up on the socket the GPU is NOT attached to and every
byte crosses the socket interconnect.  bind_to_device(device) restricts the calling process (and the threads it starts
afterwards) to the CPUs sysfs lists as local to the GPU's PCIe root, BEFORE any pinned buffer is allocated; Linux's default
local-allocation policy then places the pages on that node.  It changes nothing when the topology cannot be read (no sysfs
entry, a single node, a cpuset that excludes the local CPUs) and reports what it did."""
import os


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


def _pci_bus_id(device):
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if device < len(ids) and ids[device].strip().isdigit():
                index = int(ids[device])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        return bus
    except Exception:       # noqa: BLE001 -- no NVML, no device: nothing to bind to
        return None


def device_locality(device, sysfs="/sys/bus/pci/devices"):
    """{'bus_id', 'numa_node', 'cpus'} of a CUDA device from sysfs, or None"""
    bus = _pci_bus_id(device)
    if not bus:
        return None
    # NVML prints an 8-digit domain, sysfs a 4-digit one
    dom, rest = bus.split(":", 1)
    name = ("%04x:%s" % (int(dom, 16), rest)).lower()
    base = os.path.join(sysfs, name)
    try:
        with open(os.path.join(base, "local_cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        node = -1
        try:
            with open(os.path.join(base, "numa_node")) as f:
                node = int(f.read().strip())
        except (OSError, ValueError):
            pass
        return {"bus_id": name, "numa_node": node, "cpus": cpus}
    except OSError:
        return None


def bind_to_device(device, locality=None):
    """Restrict this process to the CPUs local to CUDA device `device`.  Returns a dict describing the outcome
    ({'bound': bool, 'numa_node', 'cpus': count, 'why'})."""
    if os.environ.get("AGPU_NO_NUMA_BIND"):
        return {"bound": False, "why": "AGPU_NO_NUMA_BIND set"}
    loc = locality if locality is not None else device_locality(device)
    if not loc or not loc["cpus"]:
        return {"bound": False, "why": "no sysfs locality for the device"}
    try:
        allowed = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return {"bound": False, "why": "sched_getaffinity unavailable"}
    local = allowed & loc["cpus"]
    if not local:
        return {"bound": False, "numa_node": loc["numa_node"], "why": "cpuset excludes the device's local CPUs"}
    if local == allowed:
        return {"bound": False, "numa_node": loc["numa_node"], "cpus": len(allowed), "why": "every allowed CPU is local already"}
    try:
        os.sched_setaffinity(0, local)
    except OSError as e:
        return {"bound": False, "numa_node": loc["numa_node"], "why": "sched_setaffinity: %s" % e}
    return {"bound": True, "numa_node": loc["numa_node"], "cpus": len(local), "of": len(allowed)}
