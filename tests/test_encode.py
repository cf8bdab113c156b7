"""K3 (rasterise + encode on the device, csrc/encode.cu) against its oracle and independent decoders.

Content parity (what the reference would store): decoding the PNG gives gray*4 / alpha*257 of
``generate_image`` (utils/image_processing.py:256-265), ``np.load(npz)['mask']`` gives
``create_instance_mask_from_cells`` of the active-cell mask (utils/sam2_data_generator.py:315-335),
no mask file for frames without active cells (:318-319).  Byte parity: GPU files == the numpy
restatement of the same parse (oracle/encode_oracle.py), which uses zlib for every checksum.
"""
import io
import struct
import zipfile
import zlib

import numpy as np
import pytest

import ctypes


class _EncStream(ctypes.Structure):        # mirror of csrc/encode.cu: struct EncStream
    _fields_ = [("lit", ctypes.c_uint32 * 256), ("len", ctypes.c_uint32 * 260), ("hdr", ctypes.c_uint32 * 96),
                ("hdr_bits", ctypes.c_int), ("eob", ctypes.c_uint32),
                ("prefix", ctypes.c_uint8 * 64), ("suffix", ctypes.c_uint8 * 96), ("head", ctypes.c_uint8 * 192),
                ("head_off", ctypes.c_uint16 * 192), ("head_bits", ctypes.c_int),
                ("prefix_len", ctypes.c_int), ("suffix_len", ctypes.c_int), ("head_len", ctypes.c_int),
                ("dist_left", ctypes.c_uint32), ("dist_up", ctypes.c_uint32), ("bpp", ctypes.c_int),
                ("seg_lanes", ctypes.c_int), ("row_lead", ctypes.c_int), ("crc_const", ctypes.c_uint32),
                ("raw_len", ctypes.c_uint32), ("p_len_prefix", ctypes.c_int), ("p_crc_prefix", ctypes.c_int),
                ("p_csize_prefix", ctypes.c_int), ("p_crc_suffix", ctypes.c_int), ("p_csize_suffix", ctypes.c_int),
                ("p_cdoff_suffix", ctypes.c_int)]


class _EncTables(ctypes.Structure):        # struct EncTables
    _fields_ = [("crc", ctypes.c_uint32 * 256), ("pow8", ctypes.c_uint32 * 32), ("rowshift", ctypes.c_uint32 * 2048),
                ("colshift", ctypes.c_uint32 * 128), ("pow64", ctypes.c_uint32 * 4096), ("pow1", ctypes.c_uint32 * 64),
                ("img", _EncStream), ("msk", _EncStream),
                ("H", ctypes.c_int), ("W", ctypes.c_int)]


def _tables(lib, L, H, W, huffman=1):
    assert int(lib.cab200_encode_tables_bytes()) == ctypes.sizeof(_EncTables)
    raw = _EncTables()
    L.check(lib.cab200_encode_tables(H, W, huffman, ctypes.addressof(raw)), "cab200_encode_tables")
    t = dict(crc=np.array(raw.crc), pow8=np.array(raw.pow8), rowshift=np.array(raw.rowshift), colshift=np.array(raw.colshift),
             pow64=np.array(raw.pow64), pow1=np.array(raw.pow1))
    for name in ("img", "msk"):
        st = getattr(raw, name)
        d = {k: getattr(st, k) for k, _ in _EncStream._fields_ if not hasattr(getattr(st, k), "__len__")}
        d["lit"], d["len"], d["hdr"] = np.array(st.lit), np.array(st.len), np.array(st.hdr)
        d["prefix"] = bytes(st.prefix)[: st.prefix_len]
        d["suffix"] = bytes(st.suffix)[: st.suffix_len]
        d["head"] = bytes(st.head)[: st.head_len]
        t[name] = d
    return t


def _mulmod(a, b):
    p = 0
    for i in range(31, -1, -1):
        if (a >> i) & 1:
            p ^= b
        b = (b >> 1) ^ (0xEDB88320 if b & 1 else 0)
    return p


def _crc0_slow(data, c):
    # CRC register started from c, no final xor: crc32(data, c') = ~reg(data, ~c')
    return (zlib.crc32(bytes(data), c ^ 0xFFFFFFFF) ^ 0xFFFFFFFF) & 0xFFFFFFFF


# ---- CPU: oracle round trips, host tables ------------------------------------------------------------

def _random_label_picture(rng, H, W, n_labels):
    # blocky picture with vertical coherence, literals, long runs and one-pixel runs
    ys = np.sort(rng.integers(0, H, 6)); xs = np.sort(rng.integers(0, W, 9))
    lab = (np.searchsorted(ys, np.arange(H))[:, None] * 10 + np.searchsorted(xs, np.arange(W))[None, :]) % n_labels
    noise = rng.random((H, W)) < 0.03
    lab[noise] = rng.integers(0, n_labels, int(noise.sum()))
    return lab


@pytest.mark.parametrize("H,W", [(64, 48), (75, 100), (33, 530), (128, 1030)])
def test_oracle_files_decode_with_independent_readers(H, W):
    from oracle import encode_oracle as EO
    rng = np.random.default_rng(H * 1000 + W)
    lab = _random_label_picture(rng, H, W, 300)
    gray = (lab * 7 % 256).astype(np.uint8)
    img = np.stack([np.where(lab > 0, gray, 0), np.where(lab > 0, 255, 0)], axis=-1).astype(np.uint8)
    inst = np.where(lab % 3 == 0, lab * 211 % 65536, 0).astype(np.uint16)
    png, npz = EO.encode_png(img), EO.encode_npz(inst)
    dec = EO.decode_png(png)
    assert np.array_equal(dec[..., 0], img[..., 0].astype(np.uint16) * 4)
    assert np.array_equal(dec[..., 1], img[..., 1].astype(np.uint16) * 257)
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:
        cvd = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
        assert cvd.dtype == np.uint16 and np.array_equal(cvd[..., 0], dec[..., 0]) and np.array_equal(cvd[..., -1], dec[..., 1])
    assert zipfile.ZipFile(io.BytesIO(npz)).testzip() is None
    m = np.load(io.BytesIO(npz))["mask"]
    assert m.dtype == np.uint16 and np.array_equal(m, inst)
    # the .npy member is what numpy itself writes
    ref = io.BytesIO(); np.save(ref, inst)
    assert EO.npy_bytes(inst) == ref.getvalue()


def test_host_tables_match_the_oracle_and_zlib():
    """cab200_encode_tables is host code: Huffman tables, CRC machinery and container templates."""
    from calcium_b200 import _lib as L, build
    from oracle import encode_oracle as EO
    build.build()
    lib = L.load()
    H, W = 96, 200
    for huffman, hname in ((0, "fixed"), (1, "tuned")):
        t = _tables(lib, L, H, W, huffman)
        for name, bpp, up in (("img", 4, 1 + 4 * W), ("msk", 2, 2 * W)):
            code = EO.TunedCode(name, bpp, up) if hname == "tuned" else None
            for b in range(256):
                v, n = EO.fixed_symbol(b) if code is None else code.ll[b]
                assert t[name]["lit"][b] == (v | n << 24)
            for ln in range(3, 259):
                v, n = EO.length_token(ln, code)
                assert t[name]["len"][ln] == (v | n << 24)
            for key, d in (("dist_left", bpp), ("dist_up", up)):
                v, n = EO.dist_token(d, code)
                assert (t[name][key] & 0xFFFFFFFF) == (v | n << 24)
            v, n = EO.fixed_symbol(256) if code is None else code.ll[256]
            assert t[name]["eob"] == (v | n << 24)
            hdr = [(3, 3)] if code is None else code.header
            packed = EO._pack(hdr)
            want = np.frombuffer(packed + bytes(8 - len(packed) % 4), dtype="<u4")
            nbits = sum(nb for _, nb in hdr)
            assert t[name]["hdr_bits"] == nbits
            assert np.array_equal(t[name]["hdr"][: (nbits + 31) // 32], want[: (nbits + 31) // 32])
    for b in range(256):
        assert t["crc"][b] == _crc0_slow(bytes([b]), 0)
    # x^(8*2^k), and CRC(zeros(n)) = (0xFFFFFFFF * x^(8n)) ^ 0xFFFFFFFF
    x = 0x00800000
    for k in range(20):
        assert int(t["pow8"][k]) == x
        n = 1 << k
        assert (_mulmod(0xFFFFFFFF, x) ^ 0xFFFFFFFF) == (zlib.crc32(bytes(n)) & 0xFFFFFFFF)
        x = _mulmod(x, x)
    # x^(8k) and x^(8*64*j); the IDAT CRC exactly as the kernel assembles it: 64-byte pieces counted from the END of
    # the range (the first may be shorter), piece j's crc0 times x^(8*64*j), all XOR-ed, then the zeros term
    x = 0x80000000
    for k in range(64):
        assert int(t["pow1"][k]) == x
        x = _mulmod(x, 0x00800000)
    assert int(t["pow64"][1]) == x and int(t["pow64"][2]) == _mulmod(x, x) and int(t["pow64"][0]) == 0x80000000
    assert int(t["pow64"][4095]) == _mulmod(int(t["pow64"][4094]), x)
    rngb = np.random.default_rng(11)
    for Lr in (1, 63, 64, 65, 1000, 33333):
        msg = rngb.integers(0, 256, Lr, dtype=np.uint8).tobytes()
        acc, j = 0, 0
        while 64 * j < Lr:
            e = Lr - 64 * j
            acc ^= _mulmod(_crc0_slow(msg[max(0, e - 64):e], 0), int(t["pow64"][j]))
            j += 1
        xl = _mulmod(int(t["pow64"][Lr >> 6]), int(t["pow1"][Lr & 63]))
        assert (acc ^ _mulmod(0xFFFFFFFF, xl) ^ 0xFFFFFFFF) == (zlib.crc32(msg) & 0xFFFFFFFF)
    # the mask CRC exactly as the kernel assembles it: per 16-pixel chunk crc0, shifted into place
    rng = np.random.default_rng(5)
    inst = np.where(rng.random((H, W)) < 0.3, rng.integers(0, 65536, (H, W)), 0).astype("<u2")
    acc = 0
    for y in range(H):
        rowc = 0
        for c in range((W + 15) // 16):
            chunk = inst[y, 16 * c:16 * c + 16].tobytes()
            rowc ^= _mulmod(_crc0_slow(chunk, 0), int(t["colshift"][c]))
        acc ^= _mulmod(rowc, int(t["rowshift"][y]))
    crc = acc ^ (t["msk"]["crc_const"] & 0xFFFFFFFF)
    assert crc == (zlib.crc32(EO.npy_bytes(inst)) & 0xFFFFFFFF)
    # container templates: the oracle's files start / end with them (variable fields zero in the template)
    img = np.zeros((H, W, 2), np.uint8)
    png = EO.encode_png(img)
    p = t["img"]
    assert png[:p["p_len_prefix"]] == p["prefix"][:p["p_len_prefix"]] and png[p["p_len_prefix"] + 4:p["prefix_len"]] == p["prefix"][p["p_len_prefix"] + 4:]
    assert png[-12:] == p["suffix"] and p["raw_len"] == H * (1 + 4 * W)
    npz = EO.encode_npz(inst)
    m = t["msk"]
    assert m["head"] == EO.npy_bytes(inst)[:m["head_len"]] and m["raw_len"] == m["head_len"] + 2 * H * W
    assert npz[:14] == m["prefix"][:14] and npz[22:38] == m["prefix"][22:38]
    tail = npz[-m["suffix_len"]:]
    for a, b in ((0, 16), (24, 54 + 16), (54 + 20, m["suffix_len"])):
        assert tail[a:b] == m["suffix"][a:b]


# ---- GPU -------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def env(cuda_device, oracle):
    import torch
    from calcium_b200 import _lib, build
    build.build()
    from calcium_b200 import ensemble
    from oracle import encode_oracle
    return dict(dev=cuda_device, O=oracle, EO=encode_oracle, lib=_lib.load(), L=_lib, E=ensemble, torch=torch)


def _run(env, size, count, out, T=4000, every=400, huffman="tuned", **kw):
    E = env["E"]
    specs = E.make_specs(count, size=size, first_seed=40)
    ens = E.Ensemble(specs, T=T, frame_steps=list(range(0, T, every)), output_size=out, device=env["dev"],
                     huffman=huffman)
    ens.simulate()
    raw = ens.rasterize(image=True, instance=True)
    enc = ens.encode(**kw)
    ens.synchronize()
    arena, off, size_, flags = ens.encoded_files(enc)
    return ens, raw, enc, arena, off, size_, flags


@pytest.mark.gpu
@pytest.mark.parametrize("parse", ["events", "pixels"])
@pytest.mark.parametrize("size,count,out", [("small", 4, (256, 256)), ("medium", 4, (512, 512)), ("xsmall", 4, (100, 75)),
                                            ("small", 2, (640, 480)), ("small", 1, (1030, 40)), ("large", 1, (512, 512)),
                                            ("large", 1, (200, 160))])
def test_k3_files_decode_to_k2_arrays_and_match_oracle_bytes(env, size, count, out, parse):
    """Both parses (static events: the fast path; every pixel: the general one) write the same bytes."""
    EO, L = env["EO"], env["L"]
    ens, raw, enc, arena, off, size_, flags = _run(env, size, count, out, parse=parse)
    if parse == "events":
        assert (ens.encode_geometry_tables()[2] > 0).all()
    image = raw["image"].cpu().numpy()
    inst = raw["instance"].cpu().numpy().view(np.uint16)
    nact = raw["n_active"].cpu().numpy()
    assert np.array_equal(enc["n_active"].cpu().numpy(), nact)
    n_bytes_checked = 0
    for i in range(ens.P):
        for f in range(ens.F):
            assert flags[i, f, 0] == L.ENC_OK
            png = arena[off[i, f, 0]: off[i, f, 0] + size_[i, f, 0]].tobytes()
            dec = EO.decode_png(png)                      # checks chunk CRCs and the Adler-32
            assert np.array_equal(dec[..., 0], image[i, f, :, :, 0].astype(np.uint16) * 4)
            assert np.array_equal(dec[..., 1], image[i, f, :, :, 1].astype(np.uint16) * 257)
            if nact[i, f] == 0:
                assert flags[i, f, 1] == L.ENC_SKIPPED and size_[i, f, 1] == 0
            else:
                assert flags[i, f, 1] == L.ENC_OK
                npz = arena[off[i, f, 1]: off[i, f, 1] + size_[i, f, 1]].tobytes()
                assert zipfile.ZipFile(io.BytesIO(npz)).testzip() is None       # zip CRC-32 of the .npy
                m = np.load(io.BytesIO(npz))["mask"]
                assert m.dtype == np.uint16 and np.array_equal(m, inst[i, f])
            if (i * ens.F + f) % 7 == 0:                  # byte parity with the oracle's parse on a sample
                assert png == EO.encode_png(image[i, f])
                if nact[i, f]:
                    assert npz == EO.encode_npz(inst[i, f])
                n_bytes_checked += 1
    assert n_bytes_checked >= 2
    assert int(enc["cursor"].item()) == int(((size_ + 15) // 16 * 16).sum())


@pytest.mark.gpu
def test_k3_fixed_huffman_mode(env):
    """huffman="fixed": one fixed-Huffman block (RFC 1951 3.2.6) instead of the static tuned code; same content."""
    EO = env["EO"]
    ens, raw, enc, arena, off, size_, flags = _run(env, "small", 2, (256, 256), huffman="fixed")
    image = raw["image"].cpu().numpy()
    inst = raw["instance"].cpu().numpy().view(np.uint16)
    tuned_bytes = _run(env, "small", 2, (256, 256))[5].sum()
    assert size_.sum() > 1.2 * tuned_bytes                      # the tuned code is what makes the files small
    for i in range(2):
        for f in (1, 5, 9):
            png = arena[off[i, f, 0]: off[i, f, 0] + size_[i, f, 0]].tobytes()
            assert png == EO.encode_png(image[i, f], huffman="fixed")
            if size_[i, f, 1]:
                npz = arena[off[i, f, 1]: off[i, f, 1] + size_[i, f, 1]].tobytes()
                assert npz == EO.encode_npz(inst[i, f], huffman="fixed")
                assert np.array_equal(np.load(io.BytesIO(npz))["mask"], inst[i, f])


@pytest.mark.gpu
def test_k3_png_is_readable_by_opencv(env):
    cv2 = pytest.importorskip("cv2")
    ens, raw, enc, arena, off, size_, flags = _run(env, "small", 1, (256, 256))
    image = raw["image"].cpu().numpy()
    for f in range(ens.F):
        png = arena[off[0, f, 0]: off[0, f, 0] + size_[0, f, 0]]
        cvd = cv2.imdecode(png, cv2.IMREAD_UNCHANGED)
        assert cvd is not None and cvd.dtype == np.uint16
        assert np.array_equal(cvd[..., 0], image[0, f, :, :, 0].astype(np.uint16) * 4)
        assert np.array_equal(cvd[..., -1], image[0, f, :, :, 1].astype(np.uint16) * 257)


@pytest.mark.gpu
def test_k3_flags_frames_that_do_not_fit_the_staging_buffer(env):
    L = env["L"]
    ens, raw, enc, arena, off, size_, flags = _run(env, "small", 1, (256, 256), stage_bytes=4096)
    full = flags[..., 0] == L.ENC_STAGE_FULL
    assert full.any() and (size_[..., 0][full] == 0).all()
    assert ((flags[..., 0] == L.ENC_OK) | full).all()          # the nearly empty first frames still fit
    with pytest.raises(L.Cab200Error):
        e2 = ens.encode(arena_bytes=4096)
        ens.encoded_files(e2)


@pytest.mark.gpu
def test_k3_intended_ids_mode_and_threshold(env):
    """quirk=False (intended rank-of-active-cell ids) and a non-default threshold go through the same LUTs as K2."""
    E = env["E"]
    specs = E.make_specs(2, size="small", first_seed=3)
    ens = E.Ensemble(specs, T=3000, frame_steps=[0, 1500, 2999], output_size=(256, 256), device=env["dev"])
    ens.simulate()
    raw = ens.rasterize(image=False, instance=True, quirk=False, threshold=0.05)
    enc = ens.encode(image=False, quirk=False, threshold=0.05)
    arena, off, size_, flags = ens.encoded_files(enc)
    inst = raw["instance"].cpu().numpy().view(np.uint16)
    assert (flags[..., 0] == env["L"].ENC_SKIPPED).all()
    for i in range(2):
        for f in range(3):
            if flags[i, f, 1] == env["L"].ENC_OK:
                m = np.load(io.BytesIO(arena[off[i, f, 1]: off[i, f, 1] + size_[i, f, 1]].tobytes()))["mask"]
                assert np.array_equal(m, inst[i, f])


@pytest.mark.gpu
@pytest.mark.parametrize("stage_bytes,wait", [(0, True), (0, False), (4096, True)])
def test_files_pipeline_delivers_every_frame_to_the_host(env, stage_bytes, wait):
    """Ensemble.run_files_to_host (host in, host out; three-slot ring, ragged last chunk), called twice
    back to back like the benchmark does; with a tiny staging buffer every frame takes the K2 + host
    encoder fall-back.  Every file decodes to the arrays of the raw path."""
    E, EO, L = env["E"], env["EO"], env["L"]
    specs = E.make_specs(5, size="small", first_seed=11)
    ens = E.Ensemble(specs, T=3000, frame_steps=[0, 700, 1400, 2999], output_size=(256, 256), device=env["dev"])
    got = {}

    def consumer(ch):
        for i in range(ch.hi - ch.lo):
            for f in range(ens.F):
                for kind in (0, 1):
                    b = ch.file(i, f, kind)
                    got[(len(got_calls), ch.lo + i, f, kind)] = None if b is None else bytes(b)
        assert ch.n_active.shape == (ch.hi - ch.lo, ens.F)

    got_calls = []
    for _ in range(2):
        moved = ens.run_files_to_host(chunk_pouches=2, consumer=consumer, wait=wait, stage_bytes=stage_bytes)
        if not wait:
            ens.flush_files()
        got_calls.append(1)
        assert moved["h2d_bytes"] == ens.h2d_bytes and moved["d2h_bytes"] > 0
    raw = ens.rasterize(image=True, instance=True)
    image = raw["image"].cpu().numpy()
    inst = raw["instance"].cpu().numpy().view(np.uint16)
    nact = raw["n_active"].cpu().numpy()
    assert len(got) == 2 * 5 * ens.F * 2
    for (call, i, f, kind), b in got.items():
        if kind == 0:
            dec = EO.decode_png(b)
            assert np.array_equal(dec[..., 0], image[i, f, :, :, 0].astype(np.uint16) * 4)
            assert np.array_equal(dec[..., 1], image[i, f, :, :, 1].astype(np.uint16) * 257)
        elif nact[i, f] == 0:
            assert b is None
        else:
            assert np.array_equal(np.load(io.BytesIO(b))["mask"], inst[i, f])
    assert ens._fring["h_frames"] is None               # default frames="none": the files are the product
    ens.run_files_to_host(chunk_pouches=5, frames="ca")
    ca_host = ens._fring["h_frames"].numpy()            # frames="ca": the calcium state [P, F, n]
    dev_frames = ens.frames.cpu().numpy().reshape(ens.P, ens.F, 4, -1)
    assert np.array_equal(ca_host, dev_frames[:, :, 0, :])
    ens.run_files_to_host(chunk_pouches=5, frames="all")
    assert np.array_equal(ens._fring["h_frames"].numpy(), ens.frames.cpu().numpy())


@pytest.mark.gpu
def test_batch_path_writes_the_encoded_frames(env, tmp_path):
    """generate_simulation_batch stores K3's byte strings: the files decode to the arrays of the raw path."""
    import glob
    import os
    import random
    from calcium_b200 import generate_simulation_batch
    EO = env["EO"]
    kw = dict(pouch_sizes=["xsmall", "small"], time_steps=[0, 5000, 9000], image_size="256x256",
              defect_configs=[{}] * 3)                                 # clean frames: the K1 -> K3 path
    random.seed(9)
    ra = generate_simulation_batch(3, str(tmp_path / "a"), keep_arrays=True, write_files=False, **kw)
    random.seed(9)
    rb = generate_simulation_batch(3, str(tmp_path / "b"), **kw)
    assert [s["simulation_name"] for s in ra["simulations"]] == [s["simulation_name"] for s in rb["simulations"]]
    n_masks = 0
    for sim in rb["simulations"]:
        arr = ra["arrays"][sim["simulation_id"]]
        assert sim["image_count"] == 3
        for fi, t in enumerate(arr["time_steps"]):
            (png,) = glob.glob(os.path.join(sim["image_dir"], f"{sim['simulation_name']}_t{t:05d}_*.png"))
            dec = EO.decode_png(open(png, "rb").read())
            assert np.array_equal(dec[..., 0], arr["image"][fi][..., 0].astype(np.uint16) * 4)
            assert np.array_equal(dec[..., 1], arr["image"][fi][..., 1].astype(np.uint16) * 257)
            npz = glob.glob(os.path.join(sim["mask_dir"], f"{sim['simulation_name']}_t{t:05d}_*.npz"))
            assert len(npz) == (1 if arr["n_active"][fi] > 0 else 0)
            if npz:
                assert np.array_equal(np.load(npz[0])["mask"], arr["instance"][fi])
                n_masks += 1
        assert sim["mask_count"] == int((arr["n_active"] > 0).sum())
        assert sim["prompt_count"] == sim["mask_count"]
    assert rb["stats"]["total_masks"] == n_masks
    # prompt files: the reference's JSON layout (utils/sam2_data_generator.py:227-232), one per mask
    import json
    pj = sorted(glob.glob(os.path.join(str(tmp_path / "b"), "prompts", "train", "*.json")))
    assert len(pj) == n_masks
    for path in pj[:3]:
        data = json.load(open(path))
        assert set(data) == {"positive_prompts", "negative_prompts", "num_cells", "version"}
        assert data["num_cells"] == len(data["positive_prompts"]) and len(data["negative_prompts"]) <= 5


# ---- K3r: arbitrary frames (edge blur, defects) -----------------------------------------------------------------

def test_raw_encoder_oracle_decodes_with_independent_readers():
    """oracle/encode_oracle.encode_raw_png (the CPU restatement of K3r): the file decodes, with zlib (checksums
    checked) and with cv2, to gray * 4 / alpha * 257 of the frame -- random frames, flat frames, odd sizes."""
    import cv2
    from oracle import encode_oracle as EO
    R = np.random.RandomState(3)
    for H, W in ((40, 70), (33, 50), (17, 31), (64, 64), (5, 200)):
        for kind in range(3):
            img = np.zeros((H, W, 2), np.uint8)
            if kind == 0:
                img[..., 0] = R.randint(0, 256, (H, W)); img[..., 1] = R.randint(0, 256, (H, W))
            elif kind == 1:
                img[H // 4: H // 2, W // 5: W // 2] = (90, 255); img[H // 2:, :, 1] = 255
            png = EO.encode_raw_png(img)
            dec = EO.decode_png(png)
            assert np.array_equal(dec[..., 0], img[..., 0].astype(np.uint16) * 4)
            assert np.array_equal(dec[..., 1], img[..., 1].astype(np.uint16) * 257)
            cv = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
            assert cv.dtype == np.uint16 and np.array_equal(cv[..., 0], dec[..., 0]) and np.array_equal(cv[..., 3], dec[..., 1])


@pytest.mark.gpu
def test_raw_encoder_matches_oracle_bytes_and_decodes(cuda_device):
    """K3r on the device: byte for byte the oracle's restatement on small frames of several shapes (random pixels --
    every byte a literal --, flat frames, real edge-blurred frames), and at 512 x 512 the decoded file is the frame."""
    import torch
    from calcium_b200 import build
    build.build()
    from calcium_b200.ensemble import Ensemble, encode_raw_frames, make_specs
    from oracle import encode_oracle as EO
    R = np.random.RandomState(11)
    for H, W in ((40, 70), (33, 50), (17, 31), (64, 64), (5, 200), (96, 130)):
        frames = np.zeros((7, H, W, 2), np.uint8)
        frames[0, ..., 0] = R.randint(0, 256, (H, W)); frames[0, ..., 1] = R.randint(0, 256, (H, W))
        frames[1] = 255
        frames[3, H // 4: H // 2, W // 5: W // 2] = (90, 255); frames[3, H // 2:, :, 1] = 255
        frames[4, ..., 0] = (np.arange(W)[None, :] * 3 + np.arange(H)[:, None]) % 256; frames[4, ..., 1] = 255
        frames[5, ..., 0] = R.randint(0, 3, (H, W)) * 100; frames[5, ..., 1] = 255
        frames[6, ::2] = 200
        arena, off, size, flags = encode_raw_frames(torch.from_numpy(frames).to(cuda_device))
        for i in range(len(frames)):
            got = arena[off[i]: off[i] + size[i]].tobytes()
            assert got == EO.encode_raw_png(frames[i]), (H, W, i)
    # real frames: edge-blurred medium pouches at 512 x 512, content checked with the independent decoder
    specs = make_specs(2, size="medium", first_seed=4)
    ens = Ensemble(specs, device=cuda_device, frame_steps=[0, 6000, 12000, 17999])
    ens.simulate()
    out = ens.rasterize(image=True, instance=False, edge_blur=True, blur_kernel_size=3, blur_type="mean")
    img = out["image"]
    arena, off, size, flags = encode_raw_frames(img)
    host = img.cpu().numpy()
    assert off.shape == (2, 4)
    for i in range(2):
        for f in range(4):
            dec = EO.decode_png(arena[off[i, f]: off[i, f] + size[i, f]].tobytes())
            assert np.array_equal(dec[..., 0], host[i, f, ..., 0].astype(np.uint16) * 4)
            assert np.array_equal(dec[..., 1], host[i, f, ..., 1].astype(np.uint16) * 257)
    assert EO.encode_raw_png(host[1, 2]) == arena[off[1, 2]: off[1, 2] + size[1, 2]].tobytes()
    assert size.mean() < 0.6 * 512 * 512           # blurred frames compress to a fraction of a byte per pixel
