"""Host -> device copy bandwidth of a pinned 160 MB buffer as a function of the CPU set the
allocating thread runs on (NUMA placement of the pinned pages): explains run-to-run spread of the
end-to-end bench number on multi-socket hosts."""
import glob, os, subprocess, sys, time, torch

def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return f'<{e}>'

print('affinity at start:', sorted(os.sched_getaffinity(0))[:8], '... n =', len(os.sched_getaffinity(0)))
print(sh('nvidia-smi topo -m | head -12'))
print(sh('lscpu | grep -i -E "numa|socket|model name" | head'))
torch.cuda.init()
bdf = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), 'pci_bus_id') else None
print('pci_bus_id attr:', bdf)
bus = sh('nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0')
print('nvidia-smi bus id:', bus)
path = '/sys/bus/pci/devices/' + bus.lower().replace('00000000:', '0000:')
for f in ('local_cpulist', 'numa_node'):
    try:
        print(f, open(os.path.join(path, f)).read().strip())
    except Exception as e:
        print(f, '<', e, '>')
nodes = sorted(glob.glob('/sys/devices/system/node/node[0-9]*'))
print('nodes:', [os.path.basename(n) for n in nodes])

def cpus_of(node):
    s = open(os.path.join(node, 'cpulist')).read().strip()
    out = []
    for part in s.split(','):
        if '-' in part:
            a, b = part.split('-'); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out

def bw(label):
    h = torch.empty(160 << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty_like(h, device='cuda')
    s = torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s):
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record()
        for _ in range(10):
            d.copy_(h, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{label:40s} {h.numel() / ms / 1e6:8.1f} GB/s')

allowed = os.sched_getaffinity(0)
bw('default affinity')
for n in nodes:
    c = set(cpus_of(n)) & allowed
    if not c:
        print(os.path.basename(n), 'no allowed cpus')
        continue
    os.sched_setaffinity(0, c)
    time.sleep(0.05)
    bw(f'allocated on {os.path.basename(n)} ({len(c)} cpus)')
os.sched_setaffinity(0, allowed)
