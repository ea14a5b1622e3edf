"""ctypes binding of libratrack_b200.so (include/ratrack_b200.h).

The library is the product: there is NO fallback.  If it is missing, or a call returns an
error code, this module raises -- it never routes to torch ops or to oracle/.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libratrack_b200.so")

_c_f = ctypes.c_void_p  # device pointers travel as raw addresses
_I, _F, _P = ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> argtypes (restype is always int); mirrors include/ratrack_b200.h
SIGNATURES = {
    "rt_ball_query": [_I, _I, _I, _F, _I, _P, _P, _P, _P],
    "rt_group_points": [_I, _I, _I, _I, _I, _P, _P, _P, _P],
    "rt_group_points_grad": [_I, _I, _I, _I, _I, _P, _P, _P, _P],
    "rt_gather_points": [_I, _I, _I, _I, _P, _P, _P, _P],
    "rt_gather_points_grad": [_I, _I, _I, _I, _P, _P, _P, _P],
    "rt_furthest_point_sampling": [_I, _I, _I, _P, _P, _P, _P],
    "rt_knn": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "rt_three_nn": [_I, _I, _I, _P, _P, _P, _P, _P],
    "rt_three_interpolate": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "rt_three_interpolate_grad": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "rt_knn_expanded": [_I, _I, _I, _I, _P, _P, _P, _P],
    "rt_dbscan": [_I, _I, _I, _P, _F, _I, _P, _P],
    "rt_sinkhorn_match": [_I, _I, _I, _P, _F, _I, _P, _P, _P, _P],
    "rt_object_embeddings": [_I, _I, _P, _P, _P, _P],
    "rt_engine_create": [ctypes.POINTER(ctypes.c_void_p), _I, ctypes.POINTER(ctypes.c_void_p), _I],
    "rt_engine_set_profile_events": [_P, _P, _P],
    "rt_engine_set_flags": [_P, _I],
    "rt_engine_stage_profile": [_P, _I],
    "rt_engine_last_status": [_P, ctypes.POINTER(ctypes.c_int)],
    "rt_engine_status_async": [_P, _P, _P],
    "rt_group_rows": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "rt_group_rows_grad": [_I, _I, _I, _I, _I, _P, _P, _P, _P],
    "rt_cv1_forward": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "rt_cv1_backward": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "rt_act_grad": [ctypes.c_longlong, _I, _I, _P, _P, _P, _P, _P, _P],
    "rt_wsum_forward": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "rt_wsum_backward": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "rt_absmax": [_P, ctypes.c_longlong, _P, _P],
    "rt_dense_tc_forward": [ctypes.c_longlong, _I, _I, _P, ctypes.c_longlong, _P, ctypes.c_longlong, ctypes.c_longlong, _P, _P, _P, _I, _P,
                            ctypes.c_longlong, _P],
    "rt_dense_tc_wgrad": [ctypes.c_longlong, _I, _I, _P, ctypes.c_longlong, _P, ctypes.c_longlong, _P, _P, _P, _P],
    "rt_backbone_forward": [_P, _I, _I] + [_P] * 15 + [ctypes.c_longlong, _P],
    "rt_backbone_forward_varlen": [_P, _I, _I] + [_P] * 17 + [ctypes.c_longlong, _P],
}
# entries whose return value is not an error code
OTHER = {
    "rt_engine_num_weights": ([], ctypes.c_int),
    "rt_engine_destroy": ([_P], None),
    "rt_engine_workspace_bytes": ([_P, _I, _I], ctypes.c_longlong),
    "rt_engine_launch_count": ([_P], ctypes.c_longlong),
    "rt_engine_num_lanes": ([_P, _I], ctypes.c_int),
    "rt_engine_stage_times": ([_P, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_char_p), _I], ctypes.c_int),
}

_lib = None


class RatrackError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RatrackError(
                f"{SO_PATH} not found: build it with `python -m ratrack_b200.build` "
                "(there is no CPU or torch fallback for these ops)")
        L = ctypes.CDLL(SO_PATH)
        L.rt_last_error.restype = ctypes.c_char_p
        L.rt_abi_version.restype = ctypes.c_int
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        for name, (args, res) in OTHER.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(rc: int, name: str):
    if rc != 0:
        msg = lib().rt_last_error().decode("utf-8", "replace")
        raise RatrackError(f"{name} failed (code {rc}): {msg}")


# launch accounting for bench.py: every successful call() is exactly one kernel launch of ours
launch_count = 0
# optional device timing of one entry point: set to {"name": "rt_...", "events": []} and every call
# of that entry is bracketed by CUDA events on the stream it is launched on
profile = None


def call(name: str, *args):
    global launch_count
    p = profile
    if p is not None and p["name"] == name:
        import torch

        st = torch.cuda.ExternalStream(args[-1]) if args[-1] else torch.cuda.default_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        check(getattr(lib(), name)(*args), name)
        e1.record(st)
        p["events"].append((e0, e1))
    else:
        check(getattr(lib(), name)(*args), name)
    launch_count += 1
