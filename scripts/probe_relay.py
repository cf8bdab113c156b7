"""8-GPU probe (ONE process, one stream per device): does routing the device->host traffic of the GPUs with the
slow host path over NVLink through the GPUs with the fast one raise the aggregate?  Follows
profiles/r1_d2h_probe_8gpu.txt (all eight copying: GPUs 0-3 11.6 GB/s each, GPUs 4-7 18.5 GB/s; 4-7 alone 44 GB/s each).

    python scripts/probe_relay.py [bytes_per_gpu_MiB] [chunk_MiB]

Each GPU holds `bytes_per_gpu` of results.  Mode "direct": every GPU copies its own bytes to pinned host memory.
Mode "relay f": GPU i < n/2 sends the fraction f of its chunks to GPU i + n/2 with a peer copy (copy engine, NVLink);
the partner forwards them to pinned host memory behind its own chunks.  Reported: wall time until every byte is in
host memory, aggregate GB/s, and the time each device's last copy finished."""
import sys
import time

import torch

n = torch.cuda.device_count()
per_gpu = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
chunk = (int(sys.argv[2]) if len(sys.argv) > 2 else 64) << 20
nchunk = per_gpu // chunk
half = n // 2
print(f"{n} GPUs, {per_gpu >> 20} MiB per GPU in {nchunk} chunks of {chunk >> 20} MiB", flush=True)
for i in range(n):
    for j in range(n):
        if i != j and not torch.cuda.can_device_access_peer(i, j):
            print(f"no peer access {i}->{j}")
src = [torch.empty(per_gpu, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
relay = [torch.empty(4 * chunk, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]       # 4-slot ring per device
host = [torch.empty(2 * per_gpu, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
for h in host:
    h.fill_(1)
own_s = [torch.cuda.Stream(device=i) for i in range(n)]       # D2H of the device's own chunks
p2p_s = [torch.cuda.Stream(device=i) for i in range(n)]       # peer sends (issued on the source device)
fwd_s = [torch.cuda.Stream(device=i) for i in range(n)]       # D2H of relayed chunks (on the partner)


def sync_all():
    for i in range(n):
        torch.cuda.synchronize(i)


def run(relayed_of, partner_of, reps=3):
    """relayed_of[i] = set of chunk indices GPU i sends through partner_of[i]."""
    best = None
    for _ in range(reps):
        sync_all()
        done = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        start = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        t0 = time.perf_counter()
        for i in range(n):
            with torch.cuda.device(i):
                start[i].record(own_s[i])
        slot_free = {i: [None] * 4 for i in range(n)}
        k_of = {i: 0 for i in range(n)}
        for c in range(nchunk):
            for i in range(n):
                lo, hi = c * chunk, (c + 1) * chunk
                if c in relayed_of.get(i, ()):
                    p = partner_of[i]
                    k = k_of[p] % 4
                    k_of[p] += 1
                    dst = relay[p][k * chunk:(k + 1) * chunk]
                    with torch.cuda.device(i), torch.cuda.stream(p2p_s[i]):
                        if slot_free[p][k] is not None:
                            p2p_s[i].wait_event(slot_free[p][k])
                        dst.copy_(src[i][lo:hi], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(p2p_s[i])
                    with torch.cuda.device(p), torch.cuda.stream(fwd_s[p]):
                        fwd_s[p].wait_event(ev)
                        host[p][per_gpu + lo: per_gpu + hi].copy_(dst, non_blocking=True)
                        fe = torch.cuda.Event()
                        fe.record(fwd_s[p])
                        slot_free[p][k] = fe
                else:
                    with torch.cuda.device(i), torch.cuda.stream(own_s[i]):
                        host[i][lo:hi].copy_(src[i][lo:hi], non_blocking=True)
        sync_all()
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best = dt
    return best


total = n * per_gpu
t = run({}, {})
print(f"direct, all {n}:                 {t * 1e3:8.1f} ms   {total / t / 1e9:7.1f} GB/s aggregate", flush=True)
if n >= 2:
    partner = {i: i + half for i in range(half)}
    for f in (1.0, 0.75, 0.5):
        k = int(round(f * nchunk))
        rel = {i: set(range(nchunk - k, nchunk)) for i in range(half)}
        t = run(rel, partner)
        print(f"relay {f:4.2f} of GPUs 0-{half - 1} via +{half}:    {t * 1e3:8.1f} ms   {total / t / 1e9:7.1f} GB/s aggregate", flush=True)
    # control: the other direction
    partner = {i + half: i for i in range(half)}
    rel = {i + half: set(range(nchunk)) for i in range(half)}
    t = run(rel, partner)
    print(f"relay 1.00 of GPUs {half}-{n - 1} via -{half}:   {t * 1e3:8.1f} ms   {total / t / 1e9:7.1f} GB/s aggregate", flush=True)
    # peer copies alone (no host traffic)
    sync_all()
    t0 = time.perf_counter()
    for c in range(nchunk):
        for i in range(half):
            with torch.cuda.device(i), torch.cuda.stream(p2p_s[i]):
                relay[i + half][(c % 4) * chunk:(c % 4 + 1) * chunk].copy_(src[i][c * chunk:(c + 1) * chunk], non_blocking=True)
    sync_all()
    t = time.perf_counter() - t0
    print(f"peer copies only, {half} pairs:      {t * 1e3:8.1f} ms   {half * per_gpu / t / 1e9:7.1f} GB/s aggregate "
          f"({per_gpu / t / 1e9:.1f} GB/s per pair)", flush=True)
