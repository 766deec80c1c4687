"""Host-memory placement for the host -> device leg of the pipeline (one process per GPU).

On an 8-GPU box every GPU hangs off one CPU socket's PCIe root.  A pinned staging buffer whose pages sit on the OTHER socket is
read across the inter-socket link on every H2D copy, and eight ranks that all allocate on node 0 share one socket's memory
controllers: round 1 measured 53 GB/s per GPU at 1-2 ranks but 28 / 23 GB/s per GPU at 4 / 8 ranks.  `bind_to_gpu_node`
pins the calling process (threads created later inherit it) and its page allocations to the NUMA node of its GPU *before* the
staging buffers are allocated; `page_nodes` reports where a buffer's pages actually are (move_pages query), so bench.py can
print the placement next to the measured H2D rate.

Linux only; every function degrades to "unknown" (None) when /sys or the syscalls are unavailable -- placement is an
optimisation of the copy path, never a correctness condition.
"""
import ctypes
import os
import platform

_SYS_SET_MEMPOLICY = {"x86_64": 238, "aarch64": 237}
_SYS_MOVE_PAGES = {"x86_64": 279, "aarch64": 239}
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0,1,2,3,8,10,11]."""
    cpus = []
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_pci_address(device_index):
    """'0000:1b:00.0' of a CUDA device (sysfs spelling), or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None


def gpu_numa_node(device_index):
    """NUMA node of the GPU's PCIe root, or None when the platform does not say (-1 in sysfs, containers without /sys)."""
    addr = gpu_pci_address(device_index)
    if addr is None:
        return None
    txt = _read("/sys/bus/pci/devices/%s/numa_node" % addr)
    try:
        node = int(txt)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node):
    return parse_cpulist(_read("/sys/devices/system/node/node%d/cpulist" % node))


def set_mempolicy(mode, node=None):
    """set_mempolicy(2) for the calling thread; True on success."""
    nr = _SYS_SET_MEMPOLICY.get(platform.machine())
    if nr is None:
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        if node is None:
            return libc.syscall(nr, mode, None, 0) == 0
        mask = (ctypes.c_ulong * 16)()
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] |= 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        return libc.syscall(nr, mode, mask, 16 * 8 * ctypes.sizeof(ctypes.c_ulong)) == 0
    except Exception:
        return False


def bind_to_gpu_node(device_index, strict=False):
    """Run this process on the CPUs of the GPU's NUMA node and prefer (strict: require) that node for new pages.
    Returns {'node', 'cpus', 'affinity', 'mempolicy'}; node None = nothing done."""
    node = gpu_numa_node(device_index)
    info = {"node": node, "cpus": 0, "affinity": False, "mempolicy": False}
    if node is None:
        return info
    cpus = node_cpus(node)
    allowed = set(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else set()
    use = sorted(allowed.intersection(cpus)) if allowed else cpus
    if use:
        try:
            os.sched_setaffinity(0, use)
            info["affinity"], info["cpus"] = True, len(use)
        except OSError:
            pass
    info["mempolicy"] = set_mempolicy(MPOL_BIND if strict else MPOL_PREFERRED, node)
    return info


def unbind(all_cpus=None):
    """Back to every CPU the process started with and the default memory policy (bench.py's CPU baseline uses all cores)."""
    try:
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)
    except OSError:
        pass
    set_mempolicy(MPOL_DEFAULT)


def page_nodes(tensor, samples=64):
    """{node: pages} over `samples` evenly spaced pages of a CPU tensor (move_pages query mode), or None."""
    nr = _SYS_MOVE_PAGES.get(platform.machine())
    if nr is None or tensor.numel() == 0:
        return None
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        nbytes = tensor.numel() * tensor.element_size()
        page = os.sysconf("SC_PAGE_SIZE")
        n = max(1, min(samples, nbytes // page))
        base = tensor.data_ptr()
        pages = (ctypes.c_void_p * n)(*[(base + (i * (nbytes // n))) // page * page for i in range(n)])
        status = (ctypes.c_int * n)()
        if libc.syscall(nr, 0, ctypes.c_ulong(n), pages, None, status, 0) != 0:
            return None
        out = {}
        for s in status:
            if s >= 0:
                out[int(s)] = out.get(int(s), 0) + 1
        return out or None
    except Exception:
        return None
