"""Tiny pass over every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
K1 (single CTA, 2- and 4-CTA clusters, FP32, weighted), K0/K0b/K0c, K2, K2b, K3, K4, K5.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs
from calcium_b200.defects import apply_defects
from calcium_b200.prompts import PromptTables, frame_prompts

dev = torch.device("cuda:0")
T, steps = 120, [0, 60, 119]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "k1"):
    for size, cluster, prec in (("xsmall", 1, 64), ("small", 2, 64), ("medium", 4, 64), ("small", 1, 32), ("medium", 1, 64)):
        ens = Ensemble(make_specs(2, size=size), T=T, frame_steps=steps, device=dev, cluster=cluster, precision=prec,
                       output_size=(64, 64))
        ens.simulate(); ens.synchronize()
        assert not ens.check_status().any()
        print("K1", size, cluster, prec, "ok", flush=True)
if which in ("all", "raster"):
    ens = Ensemble(make_specs(2, size="xsmall"), T=T, frame_steps=steps, device=dev, output_size=(96, 64))
    ens.simulate()
    out = ens.rasterize(image=True, instance=True, active=True)
    for kw in (dict(edge_blur=True), dict(edge_blur=True, blur_kernel_size=5, blur_type="motion"), dict(edge_blur=True, blur_kernel_size=4),
               dict(with_border=True), dict(with_border=True, edge_blur=True)):
        ens.rasterize(image=True, instance=False, **kw)
    enc = ens.encode()
    ens.encoded_files(enc)
    apply_defects(out["image"], {"background_fluorescence": True, "non_uniform_background": False, "vignetting": True,
                                 "quantization": True, "bit_depth": 6, "adjust_brightness_contrast": True, "contrast": 1.1}, (96, 64))
    tables = PromptTables(ens.label_maps()[1][0], 200)
    ca = ens.frames_of(0)[:, 0, :].cpu().numpy()
    frame_prompts(tables, ca[2] + 0.2, rng=np.random.RandomState(1))
    ens.run_files_to_host(chunk_pouches=1)
    ens.synchronize()
    print("K0/K0b/K0c/K2/K2b/K3/K4/K5 ok", flush=True)
