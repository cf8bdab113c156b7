"""Host-side logic of the NVLink relay (calcium_b200/relay.py): pairing from measured bandwidths and the
shared-memory mailbox two processes hand chunks over with.  The data path itself needs two GPUs
(scripts/relay_check.py, run under torchrun; tests/test_gpu_parity.py runs it when two devices are visible)."""
import multiprocessing as mp
import os
import time

import numpy as np
import pytest

relay = pytest.importorskip("calcium_b200.relay")


def test_pairs_follow_the_measured_bandwidths():
    # the box of profiles/r1_d2h_probe_8gpu.txt: GPUs 0-3 slow, 4-7 fast -> every slow rank gets a fast partner
    bw = [11.6, 11.6, 11.6, 11.7, 18.5, 18.5, 18.6, 18.6]
    pairs = relay.choose_pairs(bw)
    assert sorted(pairs) == [0, 1, 2, 3] and sorted(pairs.values()) == [4, 5, 6, 7]
    # symmetric boxes, odd worlds and single ranks stay direct
    assert relay.choose_pairs([50.0] * 8) == {}
    assert relay.choose_pairs([20.0, 21.0, 19.5, 20.5]) == {}
    assert relay.choose_pairs([10.0, 50.0, 50.0]) == {}
    assert relay.choose_pairs([55.0]) == {}
    # the slow ranks need not be the low-numbered ones
    pairs = relay.choose_pairs([44.0, 12.0, 45.0, 11.0])
    assert pairs == {3: 2, 1: 0}


def _receiver(path, n, out):
    mb = relay._Mailbox(path, create=False)
    a = mb.a
    handled, total = 0, 0
    t0 = time.time()
    while handled < n and time.time() - t0 < 20:
        if a[0] <= handled:
            time.sleep(1e-4)
            continue
        total += int(a[8 + handled % relay.SLOTS])
        handled += 1
        a[1] = handled
    out.put((handled, total))
    mb.close()


def test_mailbox_hands_chunks_between_two_processes(tmp_path):
    """Sender and receiver are separate processes (one per GPU under torchrun): posted / acked counters and the
    per-slot byte counts travel through the mapped file; a slot is reused only after its acknowledgement."""
    path = str(tmp_path / "mb")
    mb = relay._Mailbox(path, create=True)
    q = mp.get_context("spawn").Queue()
    p = mp.get_context("spawn").Process(target=_receiver, args=(path, 10, q))
    p.start()
    a = mb.a
    sizes = [1000 + 17 * i for i in range(10)]
    for seq, nb in enumerate(sizes):
        t0 = time.time()
        while a[1] < seq - relay.SLOTS + 1:
            assert time.time() - t0 < 20
            time.sleep(1e-4)
        a[8 + seq % relay.SLOTS] = nb
        a[0] = seq + 1
    handled, total = q.get(timeout=30)
    p.join(timeout=10)
    assert handled == 10 and total == sum(sizes) and a[1] == 10
    mb.close(unlink=True)
    assert not os.path.exists(path)
