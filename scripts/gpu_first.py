"""First GPU contact: parity of K0/K1/K2 against the oracle + first timings."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calcium_b200 import _lib
from calcium_b200.ensemble import Ensemble, make_specs, build_label_maps, _u16_numpy
from calcium_b200.geometry import get_geometry
from oracle import pouch_oracle as O

dev = torch.device("cuda:0")
lib = _lib.load()
print("lib version", lib.cab200_version(), torch.cuda.get_device_name(0))

def parity(size, n_p, T=18000, fs=None):
    specs = make_specs(n_p, size=size)
    ens = Ensemble(specs, T=T, frame_steps=fs, device=dev, output_size=(512, 512))
    t0 = time.time(); ens.simulate(); ens.synchronize(); t1 = time.time()
    st = ens.check_status()
    worst = 0
    for i, s in enumerate(specs):
        g = s.geometry
        z = np.zeros(g.n_cells)
        ref = O.simulate_c(g.csr, O.pack_params(s.params), z, z, s.r0, s.vplc, T, ens.frame_steps)
        got = ens.frames_of(i).cpu().numpy()
        eq = np.array_equal(ref, got)
        d = np.abs(ref - got).max()
        worst = max(worst, d)
        if not eq: print(f"  {size}[{i}] MISMATCH maxabs {d:.3e}")
    print(f"K1 {size} x{n_p}: {t1-t0:.3f}s status {st.tolist()[:4]} worst abs diff vs C oracle {worst:.3e}")
    return ens, specs

ens, specs = parity("xsmall", 4)
# raster parity on xsmall
res = ens.rasterize(image=True, instance=True, active=True)
ens.synchronize()
g = specs[0].geometry
limg, lmsk = [ _u16_numpy(t) for t in build_label_maps(g, (512,512), dev)]
cm = O.cells_mask_oracle(g.vertices, (512,512)).astype(np.uint16)
print("K0 cells_mask == oracle:", np.array_equal(cm, lmsk), "diff px", int((cm != lmsk).sum()))
pl = O.paint_label_map_cv2(g.vertices, (512,512)).astype(np.uint16)
print("label_img == oracle cv2 map:", np.array_equal(pl, limg))
img = res["image"].cpu().numpy(); ins = res["instance"].cpu().numpy().view(np.uint16); act = res["active"].cpu().numpy(); na = res["n_active"].cpu().numpy()
bad = 0
for i in range(len(specs)):
    fr = ens.frames_of(i).cpu().numpy()
    for f in range(ens.F):
        ca = fr[f, 0]
        oi, oa, oin, on = O.raster_frame_c(ca, limg, lmsk, 0.1, True)
        ok = np.array_equal(oi, img[i, f]) and np.array_equal(oa, act[i, f]) and np.array_equal(oin, ins[i, f]) and on == na[i, f]
        bad += (not ok)
print("K2 frames mismatching oracle:", bad, "of", len(specs) * ens.F)
# literal oracle (per-cell fillPoly, per-cell mask loops) on a few frames
fr = ens.frames_of(0).cpu().numpy()
for f in (0, 25, 50, 89):
    ca = fr[f, 0]
    lit = O.generate_image_oracle(g.vertices, ca, (512,512))
    a = O.active_cells_oracle(ca); am = O.active_mask_oracle(cm.astype(np.int32), a); im = O.instance_mask_oracle(am, a)
    print(f"  frame {f}: image literal eq {np.array_equal(lit, img[0,f])}, active eq {np.array_equal(am, act[0,f])}, instance eq {np.array_equal(im, ins[0,f])}, n_active {len(a)}")

parity("small", 4)
parity("medium", 2)
parity("large", 1, T=3000, fs=[0, 1000, 2999])

# timings
tf = np.zeros(1); 
import ctypes
v = ctypes.c_double()
_lib.check(lib.cab200_peak_fp64(200000, ctypes.byref(v), None), "peak"); print("FP64 peak TFLOP/s", v.value)
buf = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
_lib.check(lib.cab200_peak_hbm_write(buf.data_ptr(), buf.numel(), 5, ctypes.byref(v), None), "hbm"); print("HBM write GB/s", v.value)
del buf
for (size, P) in (("medium", 148), ("medium", 296), ("small", 148*2), ("xsmall", 148*4), ("large", 148)):
    specs = make_specs(P, size=size)
    ens = Ensemble(specs, device=dev)
    ens.simulate(); ens.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ens.simulate(upload=False); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"K1 {size} P={P}: {ms:.2f} ms  {ens.cell_timesteps/ms*1e3:.3e} cell-steps/s")
    if size == "medium" and P == 148:
        out = ens.rasterize()
        ens.synchronize()
        e0.record(); ens.rasterize(out=out); e1.record(); e1.synchronize()
        ms2 = e0.elapsed_time(e1)
        print(f"K2 medium P={P}: {ms2:.2f} ms  {ens.raster_out_bytes()/ms2/1e6:.1f} GB/s written")
        del out
    del ens
    torch.cuda.empty_cache()
