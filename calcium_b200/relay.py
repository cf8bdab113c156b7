"""NVLink relay of the final host gather (multi-GPU rows of DESIGN.md).

Pouches are independent and ranks never exchange *data they compute with*; the only shared resource of an
8-GPU box is the path from the GPUs into host memory, and on the boxes this was measured on it is lopsided:
with all eight ranks copying their encoded files out, GPUs 0-3 get 11.6 GB/s each and GPUs 4-7 18.5 GB/s,
while GPUs 4-7 alone sustain 44 GB/s each (``profiles/r1_d2h_probe_8gpu.txt``).  Pushing the chunks of a
slow-path GPU over NVLink (573 GB/s per pair, copy engine only) into a staging ring on a fast-path partner,
which forwards them to pinned host memory, raises the aggregate from 93 to 176 GB/s
(``profiles/r2_relay_probe_8gpu.txt``).

``setup(rank, world, device, cap_bytes)`` is collective over ``torch.distributed``: every rank measures its
device->host bandwidth while all ranks copy at once, the slower half is paired with the faster half if the
gap is large, the pairs exchange CUDA IPC handles (``cab200_relay_*``), and a short trial with synthetic
chunks decides -- by measurement on this box, not by assumption -- whether the relay or the direct path is
used.  The result is an :class:`Endpoint` (role ``"sender"``, ``"receiver"`` or ``None``) that
:meth:`Ensemble.run_files_to_host` uses for its chunk copy-out.  One process per GPU; the host-side
handshake (slot free / chunk posted) goes through a small shared-memory mailbox, the stream ordering
between the two processes through interprocess events.  A single-process driver (``devices=[...]``) needs
none of this: there a plain peer copy does the same (not built yet; see DESIGN.md "what comes next").
"""
from __future__ import annotations

import ctypes
import mmap
import os
import sys
import threading
import time
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import _lib

SLOTS = 3
_MB_WORDS = 64            # int64 words in a mailbox: [0] posted, [1] acked, [2] stop, [8 + k] bytes of slot k


def _mailbox_path(tag: str, sender: int) -> str:
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    return os.path.join(d, f"cab200_relay_{tag}_{sender}")


class _Mailbox:
    def __init__(self, path: str, create: bool):
        self.path = path
        if create:
            with open(path, "wb") as fh:
                fh.write(b"\0" * (_MB_WORDS * 8))
        self.fh = open(path, "r+b")
        self.mm = mmap.mmap(self.fh.fileno(), _MB_WORDS * 8)
        self.a = np.frombuffer(self.mm, dtype=np.int64)

    def close(self, unlink: bool = False):
        self.a = None
        for step in (self.mm.close, self.fh.close):
            try:
                step()
            except Exception:  # noqa: BLE001 - e.g. a caller still holds a view of the mapping
                pass
        if unlink and os.path.exists(self.path):
            os.remove(self.path)


def _vp() -> ctypes.c_void_p:
    return ctypes.c_void_p()


class Endpoint:
    """One rank's end of a relay pair.  ``role``: "sender" (this rank's chunks travel through ``partner``),
    "receiver" (forwards the partner's chunks to pinned host memory) or None (direct copies)."""

    def __init__(self, rank: int, device: torch.device):
        self.rank, self.device = rank, device
        self.role: Optional[str] = None
        self.partner: Optional[int] = None
        self.cap = 0
        self.info: Dict[str, Any] = {}
        self.lib = _lib.load()
        self._mb: Optional[_Mailbox] = None
        self._thread: Optional[threading.Thread] = None
        self._events: List[ctypes.c_void_p] = []
        self._ring = None             # receiver: cudaMalloc'ed staging ring; sender: its IPC mapping
        self._h_ring: Optional[torch.Tensor] = None
        self._seq = 0
        self.forwarded_bytes = 0      # receiver: bytes landed in host memory for the partner
        self.active = False           # the trial chose the relay

    # ---- sender -------------------------------------------------------------------------------
    def send(self, src_ptr: int, nbytes: int, stream: torch.cuda.Stream) -> int:
        """Queue the peer copy of ``nbytes`` from device pointer ``src_ptr`` on ``stream`` into the partner's
        next ring slot and post it.  Returns the ticket to pass to :meth:`wait_ack`."""
        assert self.role == "sender" and nbytes <= self.cap
        seq = self._seq
        k = seq % SLOTS
        a = self._mb.a
        while a[1] < seq - SLOTS + 1:             # the slot's previous chunk is still being forwarded
            time.sleep(20e-6)
        sp = ctypes.c_void_p(stream.cuda_stream)
        _lib.check(self.lib.cab200_copy_async(ctypes.c_void_p(self._ring.value + k * self.cap), ctypes.c_void_p(src_ptr),
                                              nbytes, sp), "cab200_copy_async (peer)")
        _lib.check(self.lib.cab200_event_record(self._events[k], sp), "cab200_event_record")
        a[8 + k] = nbytes
        a[0] = seq + 1                            # posted after the record call was issued
        self._seq = seq + 1
        return seq

    def wait_ack(self, ticket: int) -> None:
        a = self._mb.a
        while a[1] <= ticket:
            time.sleep(20e-6)

    # ---- receiver -----------------------------------------------------------------------------
    def _forward_loop(self):
        torch.cuda.set_device(self.device)
        lib, a = self.lib, self._mb.a
        stream = torch.cuda.Stream(device=self.device)
        sp = ctypes.c_void_p(stream.cuda_stream)
        done = _vp()
        _lib.check(lib.cab200_event_create(ctypes.byref(done)), "cab200_event_create")
        handled = 0
        while True:
            if a[0] <= handled:
                if a[2]:
                    break
                time.sleep(20e-6)
                continue
            k = handled % SLOTS
            nbytes = int(a[8 + k])
            _lib.check(lib.cab200_stream_wait_event(sp, self._events[k]), "cab200_stream_wait_event")
            _lib.check(lib.cab200_copy_async(ctypes.c_void_p(self._h_ring.data_ptr() + k * self.cap),
                                             ctypes.c_void_p(self._ring.value + k * self.cap), nbytes, sp),
                       "cab200_copy_async (forward)")
            _lib.check(lib.cab200_event_record(done, sp), "cab200_event_record")
            _lib.check(lib.cab200_event_synchronize(done), "cab200_event_synchronize")
            self.forwarded_bytes += nbytes
            handled += 1
            a[1] = handled
        lib.cab200_event_destroy(done)

    def close(self):
        try:
            if self.role == "receiver" and self._mb is not None:
                # the sender raises the stop flag when it closes; do not wait forever for a dead partner
                if self._thread is not None:
                    self._thread.join(timeout=10.0)
                for ev in self._events:
                    self.lib.cab200_event_destroy(ev)
                if self._ring is not None:
                    self.lib.cab200_relay_free(self._ring)
                self._mb.close(unlink=True)
            elif self.role == "sender" and self._mb is not None:
                self._mb.a[2] = 1
                if self._ring is not None:
                    self.lib.cab200_relay_close(self._ring)
                for ev in self._events:
                    self.lib.cab200_event_destroy(ev)
                self._mb.close()
        finally:
            self.role, self._mb, self._ring, self._events = None, None, None, []


def measure_d2h(device: torch.device, nbytes: int = 128 << 20, reps: int = 3) -> float:
    """GB/s of ``reps`` device->pinned-host copies of ``nbytes`` on the current stream."""
    src = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(device)
    return _timed_d2h(src, dst, reps)


def _timed_d2h(src, dst, reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e.record()
    e.synchronize()
    return src.numel() * reps / (s.elapsed_time(e) * 1e-3) / 1e9


def choose_pairs(bw: List[float], ratio: float = 0.8) -> Dict[int, int]:
    """sender -> receiver.  Ranks sorted by measured concurrent bandwidth; the slower half relays through
    the faster half (slowest with fastest) when the slow half's mean is below ``ratio`` x the fast half's."""
    n = len(bw)
    if n < 2 or n % 2:
        return {}
    order = sorted(range(n), key=lambda r: (bw[r], r))
    slow, fast = order[: n // 2], order[n // 2:]
    if sum(bw[r] for r in slow) / len(slow) >= ratio * sum(bw[r] for r in fast) / len(fast):
        return {}
    return {s: f for s, f in zip(slow, reversed(fast))}


def setup(rank: int, world: int, device: torch.device, cap_bytes: int, trial_bytes: int = 256 << 20,
          force: Optional[bool] = None) -> Endpoint:
    """Collective.  Returns this rank's :class:`Endpoint`; ``endpoint.active`` tells whether the relay is used.
    ``force``: True / False skips the decision by trial (tests)."""
    import torch.distributed as dist
    ep = Endpoint(rank, device)
    if world < 2:
        return ep
    sys.setswitchinterval(0.0005)
    lib = ep.lib
    tag = os.environ.get("MASTER_PORT", "0") + "_" + os.environ.get("TORCHELASTIC_RUN_ID", "x")

    def barrier():
        torch.cuda.synchronize(device)
        dist.barrier()

    def gather(x: float) -> List[float]:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(v) for v in out]

    # 1. concurrent device->host bandwidth per rank
    nb = 128 << 20
    src = torch.empty(nb, dtype=torch.uint8, device=device)
    dst = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    barrier()
    bw = gather(_timed_d2h(src, dst, 4))
    pairs = choose_pairs(bw) if force is not False else {}
    if force and not pairs and world % 2 == 0:      # tests: pair rank i with i + world/2 whatever the bandwidths
        pairs = {i: i + world // 2 for i in range(world // 2)}
    ep.info = {"d2h_GBps_concurrent": [round(v, 1) for v in bw], "pairs": {str(s): r for s, r in pairs.items()}}
    if not pairs:
        ep.info["mode"] = "direct (no lopsided host path measured)"
        return ep
    ep.cap = (int(cap_bytes) + 255) // 256 * 256
    recv_of = {r: s for s, r in pairs.items()}
    mine: Dict[str, Any] = {}

    def give_up(why: str) -> Endpoint:
        """Every rank calls this together (the failure flags are gathered first): plain direct copies."""
        try:
            ep.close()
        except Exception:  # noqa: BLE001
            pass
        ep.active = False
        ep.info["mode"] = f"direct (relay set-up failed: {why})"
        return ep

    with torch.cuda.device(device):
        # A rank that cannot create or open its IPC objects (a container without CUDA IPC, a GPU pair without peer
        # access) must not leave the others waiting in a collective: every phase ends with the ranks agreeing on
        # whether all of them got through, else all fall back to direct copies.
        try:
            if rank in pairs:                                  # sender: events + mailbox
                ep.role, ep.partner = "sender", pairs[rank]
                ep._mb = _Mailbox(_mailbox_path(tag, rank), create=True)
                hs = []
                for _ in range(SLOTS):
                    ev, h = _vp(), (ctypes.c_ubyte * 64)()
                    _lib.check(lib.cab200_ipc_event_create(ctypes.byref(ev), h), "cab200_ipc_event_create")
                    ep._events.append(ev)
                    hs.append(bytes(h))
                mine["events"] = hs
            elif rank in recv_of:                              # receiver: staging ring
                ep.role, ep.partner = "receiver", recv_of[rank]
                ring, h = _vp(), (ctypes.c_ubyte * 64)()
                _lib.check(lib.cab200_relay_alloc(SLOTS * ep.cap, ctypes.byref(ring), h), "cab200_relay_alloc")
                ep._ring = ring
                ep._h_ring = torch.empty(SLOTS * ep.cap, dtype=torch.uint8, pin_memory=True)
                mine["ring"] = bytes(h)
        except Exception as e:  # noqa: BLE001
            mine = {"error": f"rank {rank}: {e}"}
        everyone: List[Any] = [None] * world
        dist.all_gather_object(everyone, mine)
        errors = [m["error"] for m in everyone if isinstance(m, dict) and "error" in m]
        if errors:
            return give_up(errors[0])
        err = ""
        try:
            if ep.role == "sender":
                ring = _vp()
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[ep.partner]["ring"])
                _lib.check(lib.cab200_relay_open(hb, ctypes.byref(ring)), "cab200_relay_open")
                ep._ring = ring
            elif ep.role == "receiver":
                for hbytes in everyone[ep.partner]["events"]:
                    ev = _vp()
                    hb = (ctypes.c_ubyte * 64).from_buffer_copy(hbytes)
                    _lib.check(lib.cab200_ipc_event_open(hb, ctypes.byref(ev)), "cab200_ipc_event_open")
                    ep._events.append(ev)
                ep._mb = _Mailbox(_mailbox_path(tag, ep.partner), create=False)
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        opened: List[Any] = [None] * world
        dist.all_gather_object(opened, err)
        errors = [m for m in opened if m]
        if errors:
            return give_up(errors[0])
        if ep.role == "receiver":
            ep._thread = threading.Thread(target=ep._forward_loop, daemon=True, name="cab200-relay")
            ep._thread.start()
        barrier()

        # 2. trial: the same bytes per rank, direct vs through the relay; max over ranks decides
        chunk = min(ep.cap, 64 << 20)
        n_chunks = max(1, trial_bytes // chunk)
        tsrc = torch.empty(chunk, dtype=torch.uint8, device=device)
        tdst = torch.empty(chunk, dtype=torch.uint8, pin_memory=True)
        cs = torch.cuda.Stream(device=device)

        def run(use_relay: bool) -> float:
            barrier()
            t0 = time.perf_counter()
            if use_relay and ep.role == "sender":
                tickets = [ep.send(tsrc.data_ptr(), chunk, cs) for _ in range(n_chunks)]
                ep.wait_ack(tickets[-1])
            else:
                with torch.cuda.stream(cs):
                    for _ in range(n_chunks):
                        tdst.copy_(tsrc, non_blocking=True)
                cs.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            return max(gather(dt))

        run(True); run(False)                              # warm both paths
        t_direct = min(run(False), run(False))
        t_relay = min(run(True), run(True))
    total = world * n_chunks * chunk
    ep.info.update({"trial_direct_GBps": round(total / t_direct / 1e9, 1), "trial_relay_GBps": round(total / t_relay / 1e9, 1)})
    ep.forwarded_bytes = 0                                  # the trial's chunks do not count
    ep.active = bool(force) if force is not None else (t_relay < 0.9 * t_direct)
    ep.info["mode"] = ("relay: slow-path ranks push their chunks over NVLink to a fast-path partner" if ep.active
                       else "direct (the trial found the relay no faster)")
    if not ep.active and force is None:
        barrier()
        ep.close()
    return ep
