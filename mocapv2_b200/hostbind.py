"""Host-side placement of a rank next to its GPU (multi-GPU runs): CPU affinity and preferred memory node of the calling
process are set to the NUMA node the GPU's PCIe root hangs on, BEFORE pinned staging buffers are allocated, so that the
frames of a rank travel host memory -> its own root complex without crossing the socket interconnect.

Everything is read from sysfs; a box that does not expose the files (containers, one-node VMs) is left untouched and the
returned record says so.  Nothing here is needed for correctness.
"""
import ctypes
import os


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(pci_domain, pci_bus, pci_device):
    path = f"/sys/bus/pci/devices/{pci_domain:04x}:{pci_bus:02x}:{pci_device:02x}.0/numa_node"
    try:
        return int(open(path).read().strip())
    except (OSError, ValueError):
        return -1


def bind_to_gpu(device_index):
    """Bind the calling process to the NUMA node of CUDA device `device_index`; returns what was done."""
    import torch
    prop = torch.cuda.get_device_properties(device_index)
    node = gpu_numa_node(getattr(prop, "pci_domain_id", 0), getattr(prop, "pci_bus_id", 0), getattr(prop, "pci_device_id", 0))
    rec = {"numa_node": node, "cpus_bound": None, "mempolicy": None}
    if node < 0:
        rec["note"] = "no numa_node in sysfs for this GPU: left as it was"
        return rec
    try:
        cpus = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            rec["cpus_bound"] = len(cpus)
    except OSError as e:
        rec["note"] = f"affinity not changed: {e}"
    try:
        # set_mempolicy(MPOL_PREFERRED, {node}): pages of later allocations (the pinned staging buffers) come from that node
        # when it has room; never makes an allocation fail.  x86-64 syscall 238, aarch64 237.
        nr = {"x86_64": 238, "aarch64": 237}.get(os.uname().machine)
        if nr is not None and node < 64:
            mask = ctypes.c_ulong(1 << node)
            rc = ctypes.CDLL(None, use_errno=True).syscall(nr, 1, ctypes.byref(mask), 65)
            rec["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
    except (OSError, AttributeError) as e:
        rec["mempolicy"] = f"unavailable: {e}"
    return rec
