"""ctypes binding of ``libcalcium_b200.so`` (the C ABI in ``include/cab200.h``).

There is no CPU fallback: if the library is missing the import of any compute module fails with
instructions to build it, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# CALCIUM_B200_LIB: an alternative build of the same library (kernel experiments, scripts/k1_variants.py)
LIB_PATH = os.environ.get("CALCIUM_B200_LIB") or os.path.join(HERE, "libcalcium_b200.so")

FP64_STRICT = 64
FP32 = 32
NPARAM = 16
PARAM_SLOTS = ["k_1", "k_a", "k_p", "k_2", "V_SERCA", "K_SERCA", "K_PLC", "K_5", "c_tot", "beta",
               "k_tau", "tau_max", "k_i", "D_c", "D_p", "dt"]
ST_NONFINITE = 1
ST_BADGEOM = 2
ST_PROTOCOL = 4
ENC_OK, ENC_STAGE_FULL, ENC_ARENA_FULL, ENC_SKIPPED = 1, 2, 4, 8
HUFFMAN_FIXED, HUFFMAN_TUNED = 0, 1
BLUR_MEAN, BLUR_MOTION = 0, 1
EDGE_BLUR, EDGE_BLUR_BORDER, EDGE_BORDER = 1, 2, 3
DEFECT_MAX_ONLY, DEFECT_LUT_MAX, DEFECT_LUT_FLAG, DEFECT_MUL = 0, 1, 2, 3

#: every symbol include/cab200.h declares (tests check the library exports all of them)
EXPORTS = ["cab200_version", "cab200_last_error", "cab200_device_info",
           "cab200_sim_scratch_bytes", "cab200_sim_batch", "cab200_sim_frames_tmp_bytes", "cab200_sim_set_plan", "cab200_plan_slots", "cab200_plan_slots_entry", "cab200_k1_legacy_rows", "cab200_sim_layout",
           "cab200_cluster_layout", "cab200_cluster_capacity", "cab200_plan_cluster",
           "cab200_raster_scratch_bytes", "cab200_raster_batch", "cab200_cells_mask", "cab200_paint_label_map",
           "cab200_edge_maps", "cab200_raster_edges_batch", "cab200_legacy_choice", "cab200_sample_background", "cab200_prompts_pouch", "cab200_np_percentile80_sqrt", "cab200_legacy_randint", "cab200_defect_pass",
           "cab200_encode_tables_bytes", "cab200_encode_tables", "cab200_encode_scratch_bytes", "cab200_encode_batch",
           "cab200_encode_geometry_bytes", "cab200_encode_geometry", "cab200_encode_raw_scratch_bytes", "cab200_encode_raw_batch", "cab200_publish_to_host", "cab200_edt_scratch_bytes", "cab200_edt_classes",
           "cab200_peak_fp64", "cab200_peak_hbm_write", "cab200_peak_smem", "cab200_selftest_division",
           "cab200_relay_alloc", "cab200_relay_free", "cab200_relay_open", "cab200_relay_close", "cab200_ipc_event_create",
           "cab200_ipc_event_open", "cab200_event_create", "cab200_event_record", "cab200_event_synchronize",
           "cab200_event_destroy", "cab200_stream_wait_event", "cab200_copy_async", "cab200_write_files"]


class Cab200Error(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m calcium_b200.build` "
            f"(needs nvcc, targets sm_100a). calcium_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    i32p, i64p, f64p = POINTER(c_int32), POINTER(c_int64), POINTER(c_double)
    lib.cab200_version.restype = c_int
    lib.cab200_last_error.restype = c_char_p
    lib.cab200_device_info.argtypes = [c_int, i64p]
    lib.cab200_sim_scratch_bytes.restype = c_size_t
    lib.cab200_sim_scratch_bytes.argtypes = [c_int, c_int]
    lib.cab200_sim_batch.restype = c_int
    lib.cab200_sim_batch.argtypes = [
        c_int, i32p, c_int, i32p, i32p, i32p, i64p, i64p,      # pouch/geometry tables (host)
        c_void_p, c_void_p, c_void_p,                          # rowptr, colidx, vals (device)
        i32p, i64p, c_void_p,                                  # g_cluster, g_plan_off (host), slot_plan (device)
        i64p,                                                  # cell_off (host)
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,      # params, c0, p0, r0, vplc (device)
        c_int, c_int, i32p,                                    # T, n_frames, frame_steps (host)
        c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]
    lib.cab200_sim_frames_tmp_bytes.restype = c_size_t
    lib.cab200_sim_frames_tmp_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int]
    lib.cab200_plan_slots.restype = c_int
    lib.cab200_k1_legacy_rows.restype = c_int
    lib.cab200_k1_legacy_rows.argtypes = [c_int, i32p, i32p]
    lib.cab200_plan_slots.argtypes = [c_int, i32p, i32p, c_int, c_int, c_int, ctypes.c_longlong,
                                      ctypes.c_ulonglong, c_int, POINTER(ctypes.c_uint16), i64p]
    lib.cab200_plan_slots_entry.restype = c_int
    lib.cab200_plan_slots_entry.argtypes = [c_int, i32p, i32p, c_int, c_int, c_int, c_int, ctypes.c_longlong,
                                            ctypes.c_ulonglong, c_int, POINTER(ctypes.c_uint16), i64p]
    lib.cab200_cluster_capacity.restype = c_int
    lib.cab200_cluster_capacity.argtypes = [c_int, c_int, c_int, c_int]
    lib.cab200_cluster_layout.restype = c_int
    lib.cab200_cluster_layout.argtypes = [c_int, c_int, c_int, i64p]
    lib.cab200_plan_cluster.restype = ctypes.c_longlong
    lib.cab200_plan_cluster.argtypes = [c_int, i32p, i32p, c_int, c_int, c_int, ctypes.c_longlong,
                                        ctypes.c_ulonglong, POINTER(ctypes.c_uint16), ctypes.c_longlong, i64p]
    lib.cab200_sim_layout.restype = c_int
    lib.cab200_sim_layout.argtypes = [c_int, c_int, c_int, i64p]
    lib.cab200_sim_set_plan.restype = c_int
    lib.cab200_sim_set_plan.argtypes = [c_int, c_int, c_int]
    lib.cab200_raster_scratch_bytes.restype = c_size_t
    lib.cab200_raster_scratch_bytes.argtypes = [c_int]
    lib.cab200_raster_batch.restype = c_int
    lib.cab200_raster_batch.argtypes = [
        c_int, i32p, i32p, i64p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
        c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cab200_encode_tables_bytes.restype = c_size_t
    lib.cab200_encode_tables_bytes.argtypes = []
    lib.cab200_encode_tables.restype = c_int
    lib.cab200_encode_tables.argtypes = [c_int, c_int, c_int, c_void_p]
    lib.cab200_encode_scratch_bytes.restype = c_size_t
    lib.cab200_encode_scratch_bytes.argtypes = [c_int, c_int]
    lib.cab200_encode_geometry_bytes.restype = c_size_t
    lib.cab200_encode_geometry_bytes.argtypes = [c_int, c_int, c_int]
    lib.cab200_encode_geometry.restype = c_int
    lib.cab200_encode_geometry.argtypes = [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, i32p, c_void_p]
    lib.cab200_encode_batch.restype = c_int
    lib.cab200_encode_batch.argtypes = [
        c_int, i32p, c_int, i32p, i64p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
        c_double, c_int, c_int, c_int,                         # threshold, quirk, want_image, want_mask
        c_void_p, c_void_p, i64p, i32p,                        # tables, geom_tabs, g_tab_off, g_max_events
        c_void_p, c_size_t, c_void_p,                          # arena, arena_bytes, cursor
        c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]
    lib.cab200_edt_scratch_bytes.restype = c_size_t
    lib.cab200_edt_scratch_bytes.argtypes = [c_int, c_int, c_int]
    lib.cab200_edt_classes.restype = c_int
    lib.cab200_edt_classes.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cab200_publish_to_host.restype = c_int
    lib.cab200_publish_to_host.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cab200_cells_mask.restype = c_int
    lib.cab200_cells_mask.argtypes = [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.cab200_paint_label_map.restype = c_int
    lib.cab200_paint_label_map.argtypes = [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.cab200_edge_maps.restype = c_int
    lib.cab200_edge_maps.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p]
    lib.cab200_raster_edges_batch.restype = c_int
    lib.cab200_raster_edges_batch.argtypes = [
        c_int, i32p, i32p, i64p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
        c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cab200_defect_pass.restype = c_int
    lib.cab200_defect_pass.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.cab200_encode_raw_scratch_bytes.restype = c_size_t
    lib.cab200_encode_raw_scratch_bytes.argtypes = [c_int, c_int, c_int]
    lib.cab200_encode_raw_batch.restype = c_int
    lib.cab200_encode_raw_batch.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                            c_size_t, c_int, c_void_p]
    lib.cab200_prompts_pouch.restype = c_int
    lib.cab200_prompts_pouch.argtypes = [c_void_p, c_int, c_void_p, c_double, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.cab200_np_percentile80_sqrt.restype = c_int
    lib.cab200_np_percentile80_sqrt.argtypes = [c_void_p, c_int64, c_void_p]
    lib.cab200_legacy_randint.restype = c_int
    lib.cab200_legacy_randint.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.cab200_sample_background.restype = c_int
    lib.cab200_sample_background.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p,
                                             POINTER(c_int32), c_int32, c_void_p, POINTER(c_int64)]
    lib.cab200_legacy_choice.restype = c_int
    lib.cab200_legacy_choice.argtypes = [c_void_p, POINTER(c_int32), c_int64, c_int32, c_void_p]
    lib.cab200_peak_fp64.restype = c_int
    lib.cab200_peak_fp64.argtypes = [c_int, f64p, c_void_p]
    lib.cab200_peak_hbm_write.restype = c_int
    lib.cab200_peak_hbm_write.argtypes = [c_void_p, c_size_t, c_int, f64p, c_void_p]
    vpp = POINTER(c_void_p)
    lib.cab200_relay_alloc.argtypes = [c_size_t, vpp, c_void_p]
    lib.cab200_relay_free.argtypes = [c_void_p]
    lib.cab200_relay_open.argtypes = [c_void_p, vpp]
    lib.cab200_relay_close.argtypes = [c_void_p]
    lib.cab200_ipc_event_create.argtypes = [vpp, c_void_p]
    lib.cab200_ipc_event_open.argtypes = [c_void_p, vpp]
    lib.cab200_event_create.argtypes = [vpp]
    lib.cab200_event_record.argtypes = [c_void_p, c_void_p]
    lib.cab200_event_synchronize.argtypes = [c_void_p]
    lib.cab200_event_destroy.argtypes = [c_void_p]
    lib.cab200_stream_wait_event.argtypes = [c_void_p, c_void_p]
    lib.cab200_copy_async.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    for _n in ("relay_alloc", "relay_free", "relay_open", "relay_close", "ipc_event_create", "ipc_event_open", "event_create",
               "event_record", "event_synchronize", "event_destroy", "stream_wait_event", "copy_async"):
        getattr(lib, "cab200_" + _n).restype = c_int
    lib.cab200_write_files.restype = c_int
    lib.cab200_write_files.argtypes = [c_int64, c_void_p, i64p, c_void_p, i64p, i64p, c_int, i64p]
    lib.cab200_peak_smem.restype = c_int
    lib.cab200_peak_smem.argtypes = [c_int, c_int, f64p, c_void_p]
    lib.cab200_selftest_division.restype = c_int
    lib.cab200_selftest_division.argtypes = [ctypes.c_longlong, c_int, ctypes.c_ulonglong,
                                             POINTER(ctypes.c_ulonglong), c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cab200_last_error().decode("utf-8", "replace")
        raise Cab200Error(f"{what} failed ({rc}): {msg}")


def as_i32p(a):
    return a.ctypes.data_as(POINTER(c_int32))


def as_i64p(a):
    return a.ctypes.data_as(POINTER(c_int64))
