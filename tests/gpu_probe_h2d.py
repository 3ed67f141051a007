"""Probe: host->device copy rate of one step's inputs (8.4 MB pinned) vs CPU affinity / NUMA placement of the pinned
buffer.  Prints one JSON line.  python tests/gpu_probe_h2d.py"""
import json
import os
import subprocess

import torch


def rate(nbytes, dev, reps=40):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)      # first touch here, under the current affinity
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        for _ in range(5):
            d.copy_(h, non_blocking=True)
        s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        b.record(s)
        s.synchronize()
    ms = a.elapsed_time(b) / reps
    return nbytes / ms / 1e6   # GB/s


def main():
    dev = torch.device('cuda', 0)
    out = {}
    try:
        out['topo'] = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True, timeout=20).stdout[-1500:]
    except Exception as e:
        out['topo'] = repr(e)
    bdf = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), 'pci_bus_id') else None
    try:
        q = subprocess.run(['nvidia-smi', '-i', '0', '--query-gpu=pci.bus_id', '--format=csv,noheader'], capture_output=True,
                           text=True, timeout=10).stdout.strip()
        bdf = q.lower()
        if bdf.startswith('00000000:'):
            bdf = bdf[4:]
        out['bdf'] = bdf
        base = '/sys/bus/pci/devices/' + bdf
        out['numa_node'] = open(base + '/numa_node').read().strip()
        out['local_cpulist'] = open(base + '/local_cpulist').read().strip()
    except Exception as e:
        out['sysfs_err'] = repr(e)
    out['affinity_default'] = len(os.sched_getaffinity(0))
    nbytes = 4096 * 1024 * 2
    torch.cuda.init()
    out['gbs_default'] = rate(nbytes + 4096, dev)
    # one NUMA node at a time (from /sys/devices/system/node)
    try:
        nodes = sorted(n for n in os.listdir('/sys/devices/system/node') if n.startswith('node'))
        all_cpus = os.sched_getaffinity(0)
        for i, n in enumerate(nodes):
            cl = open('/sys/devices/system/node/%s/cpulist' % n).read().strip()
            cpus = set()
            for part in cl.split(','):
                if '-' in part:
                    a, b = part.split('-')
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            cpus &= all_cpus
            if not cpus:
                continue
            os.sched_setaffinity(0, cpus)
            out['gbs_%s' % n] = rate(nbytes + 8192 * (i + 2), dev)
            out['cpus_%s' % n] = cl
        os.sched_setaffinity(0, all_cpus)
    except Exception as e:
        out['numa_err'] = repr(e)
    out['gbs_64MB_default'] = rate(64 << 20, dev, reps=10)
    out['gbs_default_again'] = rate(nbytes + 4096 * 9, dev)
    # does the number of host threads that first-touch / fill the pinned buffer matter?
    nt = torch.get_num_threads()
    out['threads_default'] = nt
    torch.set_num_threads(1)
    out['gbs_32MB_1thread_fill'] = rate(32 << 20, dev, reps=10)
    torch.set_num_threads(nt)
    out['gbs_48MB_allthreads_fill'] = rate(48 << 20, dev, reps=10)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
