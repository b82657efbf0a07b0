"""Keep a rank's host side next to its GPU: CPU affinity (and with it first-touch placement of the pinned staging
buffers) on the NUMA node the GPU's PCIe root hangs off.  Eight ranks streaming tokens from pinned memory otherwise share
whatever node the launcher left them on, and the host->device copies of the far GPUs cross the socket interconnect."""
from __future__ import annotations

import glob
import os


def _parse_cpulist(s: str):
    cpus = set()
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device: int):
    """NUMA node of CUDA device `device` from sysfs (None if it cannot be determined)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def pin_to_gpu_numa(device: int):
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with the CPUs it is allowed to use).
    Returns a dict describing what was done, for the bench record."""
    info = {"device": device, "node": None, "cpus": None, "pinned": False}
    node = gpu_numa_node(device)
    if node is None:
        nodes = glob.glob("/sys/devices/system/node/node[0-9]*")
        info["note"] = "GPU NUMA node unknown (%d nodes visible)" % len(nodes)
        return info
    info["node"] = node
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info["cpus"] = len(use)
            info["pinned"] = True
        else:
            info["note"] = "node %d has no CPU this process may use" % node
    except Exception as e:      # noqa: BLE001
        info["note"] = repr(e)
    return info
