"""Batch entry points: ``generate_simulation_batch`` and ``BatchGenerationController``.

Signatures, defaults, result dictionaries and callback protocol follow the reference's
``main.py:115-424`` and ``utils/batch_controller.py:20-403``.  The body is re-plumbed: instead of
constructing, integrating and rasterising one pouch at a time (reference ``main.py:202-390``, plus
~3.3 s of deliberate ``time.sleep`` per pouch), all pouches are drawn on the host first -- with the
same ``random.choice`` / ``np.random`` call sequence -- and then pushed through the GPU in chunks:
K1 integrates a whole chunk in one launch, K2 rasterises every sampled frame, results stream back
through pinned buffers to a pool of writer threads.

Defects (``calcium_b200/defects.py``): a pouch's ``defect_configs`` entry applies to all
its frames (``main.py:273-275``); without ``defect_configs`` every frame gets its own ``generate_random_defect_config()``
drawn from the stdlib generator in the reference's call order (``:276-277``).  Deterministic configurations run on
the GPU (K5); frames whose configuration holds stochastic / OpenCV defects take the host chain with the pouch's own
numpy generator, interleaved with the prompt draws frame by frame exactly like the reference's loop
(``:289-293`` then ``:314-328``).  A frame whose defects raise in the reference (spontaneous luminescence and cell
fragments index channel 2 of the two-channel frame) is skipped like the reference skips it
(``:342-351``) -- ``broken_defects="off"`` switches those two defects off instead.  Pass ``defect_configs=[{}] * n``
for clean frames (the K1 -> K3 fast path).  Label JSON is out of scope.
"""
from __future__ import annotations

import concurrent.futures as cf
import datetime
import gc
import json
import os
import random
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as _lib_status
from . import io as cio
from .defects import (BROKEN_IN_REFERENCE, apply_all_defects_host, apply_defects, host_defects, plan as defect_plan,
                      random_defect_config)
from .ensemble import DEFAULT_FRAME_STEPS, Ensemble, PouchSpec, build_label_maps
from .parameters import SimulationParameters

DEFAULT_SIZES = ["xsmall", "small", "medium", "large"]
DEFAULT_TYPES = ["Single cell spikes", "Intercellular transients", "Intercellular waves", "Fluttering"]


class _NumpyEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return None          # e.g. a saved Pouch object: not serialisable, recorded as null


def _parse_image_size(image_size) -> tuple:
    if isinstance(image_size, str):
        try:
            w, h = map(int, image_size.split("x"))
            return (w, h)
        except (ValueError, AttributeError):
            print(f"Failed to parse image size '{image_size}', using default 512x512")
            return (512, 512)
    return tuple(image_size)


_RUN_ID_LOCK = threading.Lock()
_LAST_RUN_ID: List[datetime.datetime] = []


def _warn_nonfinite(status, part) -> None:
    """A diverged trajectory: the reference's generate_image raises on the first NaN frame and its loop skips that time
    step (main.py:342-351); here the frames are rendered from the clipped values -- said, not silent (not reachable with
    the parameter ranges of SimulationParameters)."""
    for li, d_ in enumerate(part):
        if int(status[li]) & _lib_status.ST_NONFINITE:
            print(f"Warning: simulation {d_['sim_idx']} ({d_['name']}) produced non-finite values; "
                  f"the reference would skip its frames from the first such time step on")


def _unique_run_id(img_dir: str) -> str:
    """The reference stamps a batch's files with ``%Y%m%d%H%M%S`` (``main.py:174``) and never produces two
    batches within one second (it sleeps seconds per simulation); here batches finish in milliseconds, and a
    second batch with the same stamp would overwrite the first one's files (same sim names, same stamp).
    Same format, but strictly increasing within the process and never a stamp already present in ``img_dir``."""
    with _RUN_ID_LOCK:
        t = datetime.datetime.now().replace(microsecond=0)
        if _LAST_RUN_ID and t <= _LAST_RUN_ID[0]:
            t = _LAST_RUN_ID[0] + datetime.timedelta(seconds=1)
        existing = set()
        if os.path.isdir(img_dir):
            existing = {f.rsplit("_", 1)[-1][:-4] for f in os.listdir(img_dir) if f.endswith(".png")}
        while t.strftime("%Y%m%d%H%M%S") in existing:
            t += datetime.timedelta(seconds=1)
        _LAST_RUN_ID[:] = [t]
        return t.strftime("%Y%m%d%H%M%S")


def _resolve_devices(device, devices) -> List[Any]:
    """``devices``: None (the single ``device``), "all" (every visible CUDA device) or a list of indices /
    ``torch.device`` objects."""
    if devices is None:
        return [device]
    if isinstance(devices, str):
        if devices != "all":
            raise ValueError("devices must be None, 'all' or a list of CUDA devices")
        return [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
    out = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
    if not out:
        raise ValueError("devices is empty")
    return out


def generate_simulation_batch(num_simulations, output_dir, pouch_sizes=None, sim_types=None,
                              time_steps=None, defect_configs=None, progress_callback=None,
                              memory_threshold=70, create_stats=True, edge_blur=False,
                              blur_kernel_size=3, blur_type="mean", generate_masks=True,
                              generate_labels=False, image_size="512x512", jpeg_quality=90,
                              save_pouch=False, dataset_split="train", *, geometry_dir=None,
                              device=None, write_files=True, keep_arrays=False,
                              max_chunk_bytes=4 << 30, writer_threads=None, generate_prompts=True,
                              prompt_threads=None, devices=None, broken_defects="skip"):
    """Generate a batch of simulations with images and instance masks (see module docstring).

    Extra keyword-only arguments: ``geometry_dir``, ``device``, ``write_files`` (False: compute only),
    ``keep_arrays`` (return the images / masks in ``results['arrays']``), ``max_chunk_bytes`` (device
    memory budget for one chunk's raster outputs), ``writer_threads``, ``generate_prompts`` (with
    ``generate_masks``: the per-frame prompt JSON of ``process_batch_for_sam2``,
    ``utils/sam2_data_generator.py:325-341`` -- K4 distance transforms + the reference's own draws from
    the global numpy generator, continued from the state the pouch's own draws left it in),
    ``prompt_threads``, ``devices`` (None: the one ``device``; "all" or a list: the pouches are sharded over
    those GPUs -- independent pouches, longest-processing-time-first, one host thread per device, no
    collective; the files and the returned dict do not depend on the sharding), ``broken_defects`` ("skip": a
    frame whose configuration enables spontaneous luminescence / cell fragments is dropped, as the reference's
    loop drops it after the IndexError those raise on two-channel frames; "off": the two are switched off).
    """
    if broken_defects not in ("skip", "off"):
        raise ValueError("broken_defects must be 'skip' or 'off'")
    if generate_labels:
        # deprecated in the reference ("use prompts instead", main.py:185) and broken there: the flag shadows the imported
        # function, so `generate_labels(pouch, time_step)` (main.py:305) raises for every frame.  Said once, not ignored
        # silently; memory_threshold (a gc.collect() trigger, main.py:205) and jpeg_quality (PNG output) change no output.
        print("Note: generate_labels=True writes no label JSON here (outside this build's scope; the reference's own call "
              "at main.py:305 fails because the flag shadows utils.labeling.generate_labels)")
    os.makedirs(output_dir, exist_ok=True)
    if pouch_sizes is None:
        pouch_sizes = list(DEFAULT_SIZES)
    if sim_types is None:
        sim_types = list(DEFAULT_TYPES)
    if time_steps is None:
        time_steps = list(DEFAULT_FRAME_STEPS)
    batch_dir = output_dir
    img_dir = os.path.join(batch_dir, "images", dataset_split)
    batch_run_id = _unique_run_id(img_dir)
    mask_dir = os.path.join(batch_dir, "masks", dataset_split)
    prompt_dir = os.path.join(batch_dir, "prompts", dataset_split)
    for d in (img_dir, mask_dir, prompt_dir):
        os.makedirs(d, exist_ok=True)
    with os.scandir(mask_dir) as it_:
        earlier_batches = any(True for _ in it_)            # an earlier batch already wrote into this directory
    results: Dict[str, Any] = {"simulations": [], "output_dir": batch_dir,
                               "num_simulations": num_simulations, "batch_index": 0}
    output_size = _parse_image_size(image_size)

    # ---- host: draw every pouch (same RNG call order as the reference loop, main.py:202-249) ----
    drawn: List[Dict[str, Any]] = []
    for sim_idx in range(num_simulations):
        pouch_size = random.choice(pouch_sizes)
        sim_type = random.choice(sim_types)
        # no configuration given for this simulation: the reference draws one per frame inside its frame loop
        # (main.py:273-277) -- the next calls of the stdlib generator after the two choices above
        frame_cfgs = None
        if not (defect_configs is not None and sim_idx < len(defect_configs)):
            frame_cfgs = {}
            for t in time_steps:
                if int(t) < int(3600 / 0.2):
                    frame_cfgs[int(t)] = random_defect_config()
        try:
            sim_params = SimulationParameters(sim_type=sim_type).generate_random_params(seed=sim_idx)
            spec = PouchSpec.draw(sim_params, size=pouch_size, sim_number=sim_idx,
                                  geometry_dir=geometry_dir)
            drawn.append({"sim_idx": sim_idx, "spec": spec, "type": sim_type, "size": pouch_size,
                          "name": f"{sim_type.replace(' ', '_')}_{sim_idx}", "params": sim_params,
                          "frame_cfgs": frame_cfgs,
                          # the reference's frame loop continues the generator its Pouch seeded (core/pouch.py:317)
                          "rng_state": np.random.get_state()})
            print(f"Running simulation {sim_idx+1}/{num_simulations}: {sim_type} on {pouch_size} pouch")
        except Exception as e:  # noqa: BLE001 - the reference skips a failing simulation (main.py:388-390)
            print(f"Error in simulation {sim_idx}: {str(e)}")
            continue

    arrays: Dict[int, Dict[str, Any]] = {}
    mappings: List[tuple] = []
    n_wt = writer_threads if writer_threads else min(32, max(8, os.cpu_count() or 8))
    pool = cf.ThreadPoolExecutor(max_workers=max(1, n_wt)) if write_files else None
    pending: List[cf.Future] = []
    T = int(3600 / 0.2)
    steps = sorted({int(t) for t in time_steps if 0 <= int(t) < T})
    w, h = output_size
    per_pouch_bytes = max(1, len(steps)) * h * w * (2 + (2 if generate_masks else 0))
    chunk = max(1, int(max_chunk_bytes // per_pouch_bytes))

    def _write_many(items) -> int:
        for path, data in items:
            with open(path, "wb") as fh:
                fh.write(data)
        return len(items)

    def _write_arena_files(arena: np.ndarray, items) -> int:
        """Files that sit in a host arena (K3 / K3r output) -- ``items``: (path, offset, size) -- written by the native
        writer pool (``cab200_write_files``) straight from the arena."""
        from . import _lib as _clib
        if not items:
            return 0
        paths = [p_.encode() for p_, _, _ in items]
        blob = b"\0".join(paths) + b"\0"
        poff = np.zeros(len(paths), dtype=np.int64)
        np.cumsum([len(p_) + 1 for p_ in paths[:-1]], out=poff[1:])
        o = np.asarray([x[1] for x in items], dtype=np.int64)
        z = np.asarray([x[2] for x in items], dtype=np.int64)
        written = np.zeros(1, dtype=np.int64)
        a_ = np.ascontiguousarray(arena)
        _clib.check(_clib.load().cab200_write_files(len(paths), blob, _clib.as_i64p(poff), a_.ctypes.data, _clib.as_i64p(o),
                                                    _clib.as_i64p(z), int(n_wt), _clib.as_i64p(written)), "cab200_write_files")
        return len(items)

    k3r_writes: Dict[int, cf.Future] = {}       # per device: the write of the last K3r arena (a pinned buffer in reuse)
    k3r_locks: Dict[int, threading.Lock] = {}
    mappings_by_sim: Dict[int, List[tuple]] = {}
    warned_defects: List[bool] = []
    final_states: Dict[int, Any] = {}

    def is_clean(d) -> bool:
        """No defect of any kind for this pouch: its frames stay flat-shaded and take the K1 -> K3 path."""
        if d["frame_cfgs"] is not None:
            return False
        c_ = defect_configs[d["sim_idx"]]
        return not c_ or not (host_defects(c_) or defect_plan(c_, output_size))

    done_count = [0]
    done_lock = threading.Lock()

    def report_done(n_done: int) -> None:
        """``progress_callback(current, total)`` as simulations complete (the reference calls it as each simulation
        of its sequential loop starts, ``main.py:212-213``; its return value is ignored there too)."""
        if not progress_callback:
            return
        with done_lock:
            for _ in range(n_done):
                done_count[0] += 1
                progress_callback(done_count[0], num_simulations)

    def run_on_device(device, drawn):
        """The chunk loop for the pouches assigned to one GPU (all of them when there is one device)."""
        # The raw-array routes (edge_blur, defects, keep_arrays) are chunked by the device memory their uncompressed
        # frames take; the K1 -> K3 route keeps no frame on the device, so one ensemble -- one K1 launch filling
        # every SM -- covers up to 1184 clean pouches.
        fast = bool(steps) and not keep_arrays and not edge_blur and all(is_clean(d) for d in drawn)
        dchunk = 1184 if fast else chunk
        for c0 in range(0, len(drawn), dchunk):
            part = drawn[c0:c0 + dchunk]
            ens = Ensemble([d["spec"] for d in part], output_size=output_size, T=T,
                           frame_steps=steps, device=device)
            tsfx = [f"_t{t:05d}_{batch_run_id}" for t in steps]
            names = {li: [d["name"] + x for x in tsfx] for li, d in enumerate(part)}
            image_files: Dict[int, List[str]] = {li: [] for li in range(len(part))}
            mask_files: Dict[int, List[str]] = {li: [] for li in range(len(part))}
            img = ins = nact = masks_k3 = png_k3r = None
            # configuration of (pouch, frame): the pouch's entry of defect_configs, else the per-frame random draw
            def frame_cfg(li: int, fi: int):
                d_ = part[li]
                c_ = d_["frame_cfgs"][steps[fi]] if d_["frame_cfgs"] is not None else defect_configs[d_["sim_idx"]]
                if c_ and broken_defects == "off" and any(c_.get(k_, False) for k_ in BROKEN_IN_REFERENCE):
                    c_ = dict(c_, **{k_: False for k_ in BROKEN_IN_REFERENCE})
                return c_
            pouch_cfgs = [None if d["frame_cfgs"] is not None else frame_cfg(li, 0) if steps else None
                          for li, d in enumerate(part)]
            per_frame = [d["frame_cfgs"] is not None for d in part]
            # host chain: per-frame random configurations, or a pouch configuration with stochastic / OpenCV defects
            host_chain = [per_frame[li] or bool(host_defects(pouch_cfgs[li])) for li in range(len(part))]
            gpu_defects = [(not host_chain[li]) and bool(defect_plan(pouch_cfgs[li], output_size)) for li in range(len(part))]
            with_defects = any(host_chain) or any(gpu_defects)
            skipped_frames: Dict[int, set] = {li: set() for li in range(len(part))}
            took_k3 = bool(steps and not keep_arrays and not edge_blur and not with_defects)
            if took_k3:
                # K1 -> K3: every frame arrives on the host as the finished .png / .npz byte string
                # (cab200_encode_batch); writing a file is a plain write of those bytes.
                def consumer(ch):
                    ipaths = [[os.path.join(img_dir, b + ".png") for b in names[li]] for li in range(ch.lo, ch.hi)]
                    mpaths = [[os.path.join(mask_dir, b + ".npz") for b in names[li]] for li in range(ch.lo, ch.hi)] \
                        if generate_masks else None
                    for i, li in enumerate(range(ch.lo, ch.hi)):
                        image_files[li] = ipaths[i]
                        if generate_masks:     # no active cell -> no mask (sam2_data_generator.py:318-319)
                            has = ch.n_active[i] > 0
                            mask_files[li] = [p_ for p_, h_ in zip(mpaths[i], has) if h_]
                            sid = part[li]["sim_idx"]
                            mappings_by_sim.setdefault(sid, []).extend(
                                (os.path.basename(a_), os.path.basename(b_)) for a_, b_, h_ in zip(ipaths[i], mpaths[i], has) if h_)
                    n_done = ch.hi - ch.lo
                    if pool is None:
                        report_done(n_done)
                        return None
                    # the chunk's files go to disk straight from the pinned arena on native writer threads
                    # (cab200_write_files); the ring slot is held until they are done
                    def write_chunk():
                        ch.write_files(ipaths, mpaths, threads=n_wt)
                        report_done(n_done)
                    fut = pool.submit(write_chunk)
                    pending.append(fut)
                    return fut
                ens.run_files_to_host(chunk_pouches=min(len(part), 37), image=True, mask=bool(generate_masks),
                                      consumer=consumer)
                status = ens.check_status()
                _warn_nonfinite(status, part)
            else:
                ens.simulate()
                status = ens.check_status()
                _warn_nonfinite(status, part)
                # edge_blur (main.py:281-287): K2b blends the blurred frame in along the cell outlines; these
                # frames are not flat-shaded any more, so they take the raw-array path and the host PNG writer
                # the instance masks are not touched by blur or defects: unless the caller wants the arrays, they
                # still arrive as finished .npz files from K3 (mask stream only)
                if steps and generate_masks and not keep_arrays:
                    from . import _lib as _clib
                    m_arena, m_off, m_size, m_flags = ens.encoded_files(ens.encode(image=False, mask=True))
                    if not np.any(m_flags[:, :, 1] & _clib.ENC_STAGE_FULL):
                        masks_k3 = (m_arena, m_off[:, :, 1], m_size[:, :, 1])
                out = ens.rasterize(image=True, instance=bool(generate_masks) and masks_k3 is None,
                                    edge_blur=bool(edge_blur), blur_kernel_size=int(blur_kernel_size),
                                    blur_type=blur_type) if steps else {}
                if steps and any(gpu_defects):
                    # K5: the deterministic defects of a pouch's configuration (main.py:289-293), in place
                    for li in range(len(part)):
                        if gpu_defects[li]:
                            apply_defects(out["image"][li], pouch_cfgs[li], output_size)
                ens.synchronize()
                if steps and write_files and not keep_arrays and not any(host_chain):
                    # K3r: blurred / defect frames are no longer flat-shaded, but nothing on the host touches them
                    # any more either: they are encoded on the device (cab200_encode_raw_batch) and arrive as
                    # finished .png files, like K3's
                    from .ensemble import encode_raw_frames
                    with k3r_locks.setdefault(ens.device.index, threading.Lock()):
                        # one pinned landing buffer per device: the previous arena's files must be on disk before it
                        # is reused, and this arena's write is queued before any other shard of this device encodes
                        prev = k3r_writes.pop(ens.device.index, None)
                        if prev is not None:
                            prev.result()
                        png_k3r = encode_raw_frames(out["image"], keep_pinned=True)
                        items = [(os.path.join(img_dir, names[li][fi] + ".png"), int(png_k3r[1][li, fi]), int(png_k3r[2][li, fi]))
                                 for li in range(len(part)) for fi in range(len(steps))]
                        k3r_writes[ens.device.index] = pool.submit(_write_arena_files, png_k3r[0], items)
                        pending.append(k3r_writes[ens.device.index])
                img = out["image"].cpu().numpy() if (steps and png_k3r is None) else None
                ins = out["instance"].cpu().numpy().view(np.uint16) if (steps and "instance" in out) else None
                nact = out["n_active"].cpu().numpy() if steps else None
                del out
            prompt_files: Dict[int, List[str]] = {li: [] for li in range(len(part))}
            want_prompts = bool(steps and generate_masks and generate_prompts)
            if steps and (want_prompts or any(host_chain)):
                # The per-pouch frame loop of the reference (main.py:266-351) for everything that draws from numpy's
                # generator: per frame first the defects (:289-293), then the prompts (:314-328).  Pouches are
                # independent: the reference reseeds per pouch (core/pouch.py:317), so every pouch gets a private
                # legacy generator restored to the state its own draws left behind -- same values as the reference's
                # sequential loop -- and the pouches run on a thread pool (numpy's generators, OpenCV and the K4
                # calls release the GIL).
                tables: Dict[int, Any] = {}
                cmask_host: Dict[int, np.ndarray] = {}
                if want_prompts:
                    from .prompts import frame_prompts, pouch_prompts, prompt_tables
                for li in range(len(part)):
                    g = int(ens.geom_id[li])
                    if want_prompts and g not in tables:
                        # the per-geometry label map is cached for the life of the process, so are its tables
                        tables[g] = prompt_tables(build_label_maps(ens.geometries[g], output_size, ens.device)[1],
                                                  int(ens.g_ncell[g]))
                        tables[g].candidate_entries()                 # built before the threads start
                        tables[g].native_tables(5)
                    if host_chain[li] and g not in cmask_host:         # pouch.get_cell_masks() (:292), a host int32 map
                        cmask_host[g] = build_label_maps(ens.geometries[g], output_size, ens.device)[1].cpu().numpy() \
                            .view(np.uint16).astype(np.int32)
                ca_host = [ens.frames_of(li)[:, 0, :].cpu().numpy() for li in range(len(part))] if want_prompts else None

                def pouch_frames(li: int):
                    d = part[li]
                    rs = np.random.RandomState()
                    rs.set_state(d["rng_state"])
                    g = int(ens.geom_id[li])
                    files = []
                    if want_prompts and not host_chain[li]:
                        # clean frames: nothing but the prompts draws from the pouch's generator, so all its frames go
                        # through one native call outside the interpreter lock (cab200_prompts_pouch) -- same draws,
                        # same files, same generator state as the frame loop below
                        ppaths = [os.path.join(prompt_dir, names[li][fi] + ".json") for fi in range(len(steps))]
                        try:
                            nent, _ = pouch_prompts(tables[g], ca_host[li], rs, paths=ppaths if write_files else None)
                            return [ppaths[fi] for fi in range(len(steps)) if nent[fi] >= 0], rs.get_state()
                        except Exception as e:  # noqa: BLE001 - one pouch must not abort the batch: redo it frame by frame
                            print(f"Native prompt path failed for simulation {d['sim_idx']} ({e}); falling back to the frame loop")
                            rs.set_state(d["rng_state"])
                    with torch.cuda.device(ens.device):
                        for fi in range(len(steps)):
                            try:
                                if host_chain[li]:
                                    c_ = frame_cfg(li, fi)
                                    if c_:
                                        img[li, fi] = apply_all_defects_host(img[li, fi], cmask_host[g], c_, rs)
                                if not want_prompts or not (ca_host[li][fi] > 0.1).any():
                                    continue                          # no active cell: nothing is written (:318-319)
                                text, _ = frame_prompts(tables[g], ca_host[li][fi], rng=rs, as_text=True)
                                ppath = os.path.join(prompt_dir, names[li][fi] + ".json")
                                if write_files:
                                    with open(ppath, "w") as fh:        # the text save_prompts writes (prompts.py)
                                        fh.write(text)
                                files.append(ppath)
                            except Exception as e:  # noqa: BLE001 - a failing frame is reported and skipped (main.py:342-351)
                                skipped_frames[li].add(fi)
                                print(f"Error processing time step {steps[fi]} for simulation {d['sim_idx']}: {str(e)}")
                    return files, rs.get_state()

                last_state = None
                n_pt = prompt_threads if prompt_threads else min(32, max(8, os.cpu_count() or 8))
                with cf.ThreadPoolExecutor(max_workers=max(1, n_pt)) as pp:
                    for li, (files, state) in enumerate(pp.map(pouch_frames, range(len(part)))):
                        prompt_files[li] = files
                        last_state = state
                if last_state is not None:
                    final_states[part[-1]["sim_idx"]] = last_state
            if img is not None and png_k3r is None and pool is not None and any(host_chain):
                # the frames the host chain finished (noise, defocus: numpy's generator and OpenCV) go back to the device
                # for their PNG encoding: 0.5 MB up and ~0.2 MB down per frame instead of 8 ms of zlib on a host core
                from .ensemble import encode_raw_frames
                with k3r_locks.setdefault(ens.device.index, threading.Lock()):
                    prev = k3r_writes.pop(ens.device.index, None)
                    if prev is not None:
                        prev.result()
                    with torch.cuda.device(ens.device):
                        png_k3r = encode_raw_frames(torch.from_numpy(img).to(ens.device), keep_pinned=True)
                    items = [(os.path.join(img_dir, names[li][fi] + ".png"), int(png_k3r[1][li, fi]), int(png_k3r[2][li, fi]))
                             for li in range(len(part)) for fi in range(len(steps)) if fi not in skipped_frames[li]]
                    k3r_writes[ens.device.index] = pool.submit(_write_arena_files, png_k3r[0], items)
                    pending.append(k3r_writes[ens.device.index])
            arena_masks: List[tuple] = []
            for li, d in enumerate(part):
                if img is not None or png_k3r is not None:
                    for fi, t in enumerate(steps):
                        if fi in skipped_frames[li]:                   # the reference's loop skips a failing frame entirely
                            continue
                        base = names[li][fi]
                        ipath = os.path.join(img_dir, base + ".png")
                        image_files[li].append(ipath)
                        if png_k3r is not None:
                            pass                                       # already queued with the whole arena (K3r)
                        elif pool is not None:
                            pending.append(pool.submit(cio.save_gray_alpha_png, img[li, fi], ipath, 10))
                        if generate_masks and nact[li, fi] > 0:      # no active cell -> no mask (sam2_data_generator.py:318-319)
                            mpath = os.path.join(mask_dir, base + ".npz")
                            mask_files[li].append(mpath)
                            mappings_by_sim.setdefault(part[li]["sim_idx"], []).append((os.path.basename(ipath), os.path.basename(mpath)))
                            if pool is not None and masks_k3 is not None:
                                arena_masks.append((mpath, int(masks_k3[1][li, fi]), int(masks_k3[2][li, fi])))
                            elif pool is not None:
                                pending.append(pool.submit(cio.save_instance_mask, ins[li, fi], mpath))
                if keep_arrays:
                    arrays[d["sim_idx"]] = {
                        "time_steps": list(steps), "image": img[li].copy() if img is not None else None,
                        "instance": ins[li].copy() if ins is not None else None,
                        "n_active": nact[li].copy() if nact is not None else None,
                        "frames": ens.frames_of(li).cpu().numpy(), "status": int(status[li])}
                pouch_obj = None
                if save_pouch:
                    from .pouch import Pouch
                    pouch_obj = Pouch(params=d["params"], size=d["size"], sim_number=d["sim_idx"],
                                      save=True, save_name=d["name"], geometry_dir=geometry_dir,
                                      output_size=output_size, jpeg_quality=jpeg_quality, device=device)
                    pouch_obj.simulate()
                results["simulations"].append({
                    "simulation_id": d["sim_idx"], "simulation_name": d["name"],
                    "simulation_type": d["type"], "pouch_size": d["size"], "parameters": d["params"],
                    "image_count": len(image_files[li]), "image_dir": img_dir, "mask_dir": mask_dir,
                    "mask_count": len(mask_files[li]), "prompt_count": len(prompt_files[li]), "pouch": pouch_obj})
            if pool is not None and arena_masks:
                pending.append(pool.submit(_write_arena_files, masks_k3[0], arena_masks))
            if not took_k3:
                report_done(len(part))
            ens.release_files_ring()
            del ens

    # ---- shard over the GPUs of the box: independent pouches, no exchange; longest-
    # processing-time-first over n_cells * T, one host thread per device
    dev_list = _resolve_devices(device, devices)
    try:
        if len(dev_list) <= 1:
            run_on_device(dev_list[0] if dev_list else device, drawn)
        else:
            from .parallel import lpt_partition
            parts = lpt_partition([float(d["spec"].geometry.n_cells) * T for d in drawn], len(dev_list))
            with cf.ThreadPoolExecutor(max_workers=len(dev_list)) as dpool:
                futs = [dpool.submit(run_on_device, dev_list[r], [drawn[i] for i in parts[r]])
                        for r in range(len(dev_list)) if parts[r]]
                for f in futs:
                    f.result()
            results["simulations"].sort(key=lambda s_: s_["simulation_id"])
        if final_states:
            np.random.set_state(final_states[max(final_states)])      # where the reference's loop would leave it
        mappings = [m for k in sorted(mappings_by_sim) for m in mappings_by_sim[k]]
    finally:
        if pool is not None:
            for f in pending:
                f.result()
            pool.shutdown()

    try:
        if mappings and write_files:
            # every pair present in the directory (earlier batches included), like the reference's scan
            if earlier_batches:
                on_disk = cio.scan_image_mask_pairs(batch_dir, dataset_split)
                mappings = sorted(set(on_disk) | set(mappings))
            results["image_mask_csv"] = cio.save_csv_mapping(mappings, batch_dir, "train.csv")
            print(f"Created CSV mapping: {results['image_mask_csv']}")
        else:
            results["csv_error"] = "Failed to create CSV mapping"
    except Exception as e:  # noqa: BLE001
        results["csv_error"] = str(e)
    if create_stats:
        results["stats"] = {"total_simulations": len(results["simulations"]),
                            "total_images": sum(s["image_count"] for s in results["simulations"]),
                            "total_masks": sum(s["mask_count"] for s in results["simulations"])}
    try:
        with open(os.path.join(batch_dir, "simulation_results.json"), "w") as fh:
            json.dump(results, fh, indent=4, cls=_NumpyEncoder)
    except Exception as e:  # noqa: BLE001
        print(f"Error saving results: {str(e)}")
    if keep_arrays:
        results["arrays"] = arrays
    # (the reference ends with gc.collect() because it holds (n, 4, T) histories per pouch; nothing of that size lives
    # on the host here, and a full collection over the call's ~10^5 small objects cost 36 ms of a 0.17 s call)
    return results


class BatchGenerationController:
    """Controls batch generation independent of any GUI (reference ``utils/batch_controller.py``)."""

    def __init__(self):
        self.cancel_requested = False

    def cancel(self):
        self.cancel_requested = True

    def run_batch_generation(self, batch_params: Dict[str, Any], output_dir: str,
                             progress_callback: Optional[Callable] = None,
                             status_callback: Optional[Callable[[str], None]] = None) -> Dict[str, Any]:
        self.cancel_requested = False
        try:
            if status_callback:
                status_callback("Initializing batch generation...")
            config = self._parse_batch_config(batch_params)
            if config["enable_multi_batch"]:
                results = self._run_multi_batch_generation(config, output_dir, progress_callback, status_callback)
            else:
                results = self._run_single_batch_generation(config, output_dir, progress_callback, status_callback)
            if self.cancel_requested:
                if status_callback:
                    status_callback("Batch generation cancelled")
                results["success"] = False
                results["cancelled"] = True
                return results
            if config["create_csv"] and results.get("simulations"):
                results["csv_path"] = results.get("image_mask_csv")
                if status_callback and results["csv_path"]:
                    status_callback(f"CSV mapping created: {os.path.basename(results['csv_path'])}")
            if config["enable_video"] and results.get("simulations") and status_callback:
                # the reference calls Pouch.make_animation, which does not exist
                status_callback("Video generation error: 'Pouch' object has no attribute 'make_animation'")
            if status_callback:
                sims = results.get("simulations", [])
                status_callback(f"Batch complete: {len(sims)} simulations, "
                                f"{sum(s.get('image_count', 0) for s in sims)} images")
            results["success"] = True
            return results
        except Exception as e:  # noqa: BLE001 - same conversion as the reference (:134-144)
            msg = f"Batch generation error: {str(e)}"
            if status_callback:
                status_callback(msg)
            return {"success": False, "error": msg, "simulations": [], "output_dir": output_dir}

    def _parse_batch_config(self, bp: Dict[str, Any]) -> Dict[str, Any]:
        cfg: Dict[str, Any] = {"enable_multi_batch": bp.get("enable_multi_batch", False)}
        if cfg["enable_multi_batch"]:
            cfg["num_batches"] = int(bp.get("num_batches", 1))
            cfg["sims_per_batch"] = int(bp.get("sims_per_batch", 10))
            cfg["total_simulations"] = cfg["num_batches"] * cfg["sims_per_batch"]
        else:
            cfg["num_simulations"] = int(bp.get("num_simulations", 5))
            cfg["total_simulations"] = cfg["num_simulations"]
        try:
            w, h = map(int, bp.get("image_size", "512x512").split("x"))
            cfg["output_size"] = (w, h)
        except (ValueError, AttributeError):
            cfg["output_size"] = (512, 512)
        cfg["pouch_sizes"] = bp.get("pouch_sizes", ["small"])
        cfg["sim_types"] = bp.get("sim_types", None)
        cfg["time_steps"] = bp.get("time_steps", None)
        cfg["jpeg_quality"] = bp.get("jpeg_quality", 90)
        cfg["edge_blur"] = bp.get("edge_blur", False)
        cfg["blur_kernel_size"] = int(bp.get("blur_kernel_size", 3))
        cfg["blur_type"] = bp.get("blur_type", "mean")
        cfg["generate_masks"] = bp.get("generate_masks", True)
        cfg["create_csv"] = bp.get("create_csv", False)
        cfg["enable_video"] = bp.get("enable_video", False)
        cfg["defect_config"] = bp.get("defect_config", None)
        cfg["dataset_split"] = bp.get("dataset_split", "train")
        cfg["memory_threshold"] = 60 if ("large" in cfg["pouch_sizes"] or "xlarge" in cfg["pouch_sizes"]) else 70
        for k in ("geometry_dir", "device", "devices", "write_files", "keep_arrays"):
            if k in bp:
                cfg[k] = bp[k]
        return cfg

    def _call(self, config, output_dir, n, progress_callback, save_pouch):
        dc = config.get("defect_config")
        extra = {k: config[k] for k in ("geometry_dir", "device", "devices", "write_files", "keep_arrays") if k in config}
        return generate_simulation_batch(
            num_simulations=n, output_dir=output_dir, pouch_sizes=config["pouch_sizes"],
            sim_types=config["sim_types"], time_steps=config["time_steps"],
            defect_configs=[dc] * n if dc else None, progress_callback=progress_callback,
            create_stats=True, edge_blur=config["edge_blur"], blur_kernel_size=config["blur_kernel_size"],
            blur_type=config["blur_type"], generate_masks=config["generate_masks"],
            image_size=config["output_size"], jpeg_quality=config["jpeg_quality"],
            memory_threshold=config["memory_threshold"], save_pouch=save_pouch,
            dataset_split=config.get("dataset_split", "train"), **extra)

    def _run_single_batch_generation(self, config, output_dir, progress_callback, status_callback):
        if status_callback:
            status_callback(f"Generating {config['num_simulations']} simulations...")
        return self._call(config, output_dir, config["num_simulations"], progress_callback,
                          config["enable_video"])

    def _run_multi_batch_generation(self, config, output_dir, progress_callback, status_callback):
        all_results: Dict[str, Any] = {"simulations": [], "output_dir": output_dir}
        for b in range(config["num_batches"]):
            if self.cancel_requested:
                break
            if status_callback:
                status_callback(f"Batch {b + 1}/{config['num_batches']}: "
                                f"Generating {config['sims_per_batch']} simulations...")

            def batch_progress(current, total, _b=b):
                if progress_callback:
                    info = {"batch_idx": _b, "batch_total": config["num_batches"]}
                    return progress_callback(current + _b * config["sims_per_batch"],
                                             config["total_simulations"], info)
                return False

            r = self._call(config, output_dir, config["sims_per_batch"], batch_progress, False)
            all_results["simulations"].extend(r.get("simulations", []))
            if "image_mask_csv" in r:
                all_results["image_mask_csv"] = r["image_mask_csv"]
        return all_results
