"""2+ GPU functional check of the NVLink relay (torchrun --nproc-per-node N scripts/relay_check.py): the relay is
forced on, every rank runs the files pipeline twice, and what the receivers landed in pinned host memory is
compared byte for byte with the senders' device arenas."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from calcium_b200 import relay
from calcium_b200.ensemble import Ensemble, make_specs
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
specs = make_specs(6, size="small", first_seed=rank * 6)
ens = Ensemble(specs, T=2000, frame_steps=[0, 500, 1000, 1999], output_size=(256, 256), device=dev)
ep = relay.setup(rank, world, dev, ens.files_chunk_capacity(2), trial_bytes=32 << 20, force=True)
assert ep.active and ep.role in ("sender", "receiver"), (ep.role, ep.info)
ens.set_relay(ep)
sums = []
for rep in range(2):
    moved = ens.run_files_to_host(chunk_pouches=2, wait=True)
    # sender: byte sums of the three chunks of this call, straight from a fresh encode of the same pouches
    if ep.role == "sender":
        for lo in range(0, 6, 2):
            enc = ens.encode(pouches=(lo, lo + 2))
            used = int(enc["cursor"].item())
            sums.append([used, int(enc["arena"][:used].to(torch.int64).sum().item())])
torch.cuda.synchronize(); dist.barrier()
gathered = [None] * world
dist.all_gather_object(gathered, (ep.role, ep.partner, sums, int(ep.forwarded_bytes)))
if ep.role == "receiver":
    role, partner, psums, _ = gathered[ep.partner]
    assert role == "sender" and partner == rank
    assert ep.forwarded_bytes == sum(u for u, _ in psums), (ep.forwarded_bytes, psums)
    # the ring holds the last SLOTS chunks; chunk j of 6 went to slot j % 3
    for j in range(3, 6):
        used, want = psums[j]
        got = int(ep._h_ring[(j % relay.SLOTS) * ep.cap:][:used].to(torch.int64).sum().item())
        assert got == want, (j, got, want)
    print(f"rank {rank}: forwarded {ep.forwarded_bytes} bytes for rank {ep.partner}, last 3 chunks byte-sums equal", flush=True)
dist.barrier()
ens.set_relay(None); ep.close()
if rank == 0:
    print("relay check ok", ep.info, flush=True)
dist.barrier(); dist.destroy_process_group()
