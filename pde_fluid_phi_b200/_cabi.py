"""ctypes binding of libpdephi_b200.so (include/pdephi_b200.h).

The library is the product: there is no Python / torch fallback.  If the shared object is
missing the first use raises, loudly, with the build command.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

# PDEPHI_LIB: an alternate build of the same library (kernel experiments: `python -m pde_fluid_phi_b200.build --variant NAME -D...`)
_LIB_PATH = os.environ.get("PDEPHI_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libpdephi_b200.so")
_lib = None
_lock = threading.RLock()

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = range(5)
PREC_FP32, PREC_TF32, PREC_TF32X3 = 0, 1, 2
PREC_AUTO = -1    # host-side only: TF32X3 where the plan supports it, FP32 (FFMA) elsewhere -- both meet the 1e-5 gate
FORWARD, ADJOINT = 0, 1
STAGES_ZY, STAGES_X, STAGES_ALL = 1, 2, 3
OPT_TF32_XSTAGE_SPLIT = 0
OPT_BLOCK_SAVES_DERIVATIVE = 1         # option key; values 0 / 1 / 2 (default 2: packed binary16, see include/pdephi_b200.h)
ACTIVATIONS = {"identity": 0, "gelu": 1, "relu": 2, "tanh": 3, "silu": 4, "sigmoid": 5, "leaky_relu": 6, "elu": 7}
ROUTE_NONE, ROUTE_FFMA, ROUTE_FUSED_TC, ROUTE_PER_AXIS_TC = 0, 1, 2, 3
PER_AXIS_SPLIT_MIN_POINTS = 1 << 25      # 'auto': below this many points per call the FFMA kernels beat the per-axis split route
PRECISIONS = {"auto": PREC_AUTO, "fp32": PREC_FP32, "tf32": PREC_TF32, "tf32x3": PREC_TF32X3}

def rational_order_supported(order) -> bool:
    """(p, q) pairs the K2 kernels are instantiated for (ORDER_CASES in csrc/mix.cu)."""
    try:
        p, q = (int(v) for v in order)
    except (TypeError, ValueError):
        return False
    return (2 <= p <= 4 and 2 <= q <= 4) or (p == q and p in (1, 5, 6))


# name -> (restype, argtypes); mirrors include/pdephi_b200.h one to one
SIGNATURES = {
    "pdephi_version": (c_int, []),
    "pdephi_last_error": (c_char_p, []),
    "pdephi_plan_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "pdephi_plan_create_slab": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "pdephi_plan_destroy": (c_int, [c_void_p]),
    "pdephi_plan_workspace_bytes": (c_size_t, [c_void_p, c_longlong]),
    "pdephi_plan_supports": (c_int, [c_void_p, c_int]),
    "pdephi_plan_route": (c_int, [c_void_p, c_int]),
    "pdephi_analysis_r2c": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pdephi_analysis_stages": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pdephi_synthesis_stages": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                        c_longlong, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pdephi_projection_bwd_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong,
                                              c_int, c_void_p]),
    "pdephi_synthesis_c2r": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                     c_longlong, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pdephi_rational_mix_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p,
                                        c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong,
                                        c_void_p]),
    "pdephi_rational_transfer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_int,
                                         c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pdephi_rational_mix_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_longlong, c_void_p]),
    "pdephi_projection_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p]),
    "pdephi_rational_mix_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_longlong]),
    "pdephi_rational_mix_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p, c_size_t,
                                        c_void_p]),
    "pdephi_complex_mix_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_int, c_void_p]),
    "pdephi_complex_mix_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong,
                                       c_int, c_int, c_void_p]),
    "pdephi_pointwise_linear": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_void_p]),
    "pdephi_pointwise_wgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_longlong]),
    "pdephi_pointwise_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p,
                                       c_size_t, c_void_p]),
    "pdephi_block_tail_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong,
                                      c_void_p]),
    "pdephi_block_tail_bwd_workspace_bytes": (c_size_t, [c_int]),
    "pdephi_block_tail_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      c_int, c_longlong, c_void_p, c_size_t, c_void_p]),
    "pdephi_block_spectral_supported": (c_int, [c_void_p, c_int]),
    "pdephi_block_spectral_fwd_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "pdephi_block_spectral_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pdephi_block_tail_proj_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                           c_void_p, c_int, c_int, c_longlong, c_void_p]),
    "pdephi_block_spectral_proj_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                               c_size_t, c_void_p]),
    "pdephi_block_tail_proj_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_longlong, c_void_p, c_size_t, c_void_p]),
    "pdephi_block_tail_lift_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int, c_int, c_longlong, c_void_p]),
    "pdephi_block_spectral_lift_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                               c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t,
                                               c_void_p]),
    "pdephi_block_tail_lift_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_void_p, c_size_t, c_void_p]),
    "pdephi_block_spectral_bwd_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "pdephi_block_spectral_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                          c_void_p, c_size_t, c_void_p]),
    "pdephi_block_spectral_vz_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "pdephi_block_spectral_fwd_z": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_int,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t,
                                            c_void_p]),
    "pdephi_analysis_from_z": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_size_t, c_void_p]),
    "pdephi_wide_linear": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int,
                                   c_void_p]),
    "pdephi_wide_wgrad_workspace_bytes": (c_size_t, [c_int, c_longlong]),
    "pdephi_wide_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p, c_size_t, c_void_p]),
    "pdephi_wide_block_tail_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int,
                                           c_void_p]),
    "pdephi_wide_block_tail_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                           c_longlong, c_int, c_void_p, c_size_t, c_void_p]),
    "pdephi_set_option": (c_int, [c_int, c_int]),
    "pdephi_get_option": (c_int, [c_int]),
    "pdephi_profile_enable": (c_int, [c_int]),
    "pdephi_profile_reset": (c_int, []),
    "pdephi_profile_num_kernels": (c_int, []),
    "pdephi_profile_kernel_name": (c_char_p, [c_int]),
    "pdephi_profile_query": (c_int, [c_int, ctypes.POINTER(c_longlong), ctypes.POINTER(ctypes.c_double)]),
}


class PdephiError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise PdephiError(
                    f"{_LIB_PATH} is missing: the CUDA library has not been built and there is no fallback path. "
                    "Build it with `python -m pde_fluid_phi_b200.build` (needs nvcc, sm_100a).")
            lib = ctypes.CDLL(_LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc == OK:
        return
    msg = load().pdephi_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(f"pdephi {what}: {msg}")
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(f"pdephi {what}: {msg}")
    raise PdephiError(f"pdephi {what} failed (code {rc}): {msg}")


_plans = {}


def get_plan(device_index: int, grid, modes_h, slab=None):
    """Cached plan handle for (device, grid, (mx,my,mzh)[, (Nx_total, x_offset)]); tables live for the process lifetime."""
    grid = tuple(int(v) for v in grid)
    slab = (grid[0], 0) if slab is None else (int(slab[0]), int(slab[1]))
    key = (int(device_index), grid, tuple(int(v) for v in modes_h), slab)
    h = _plans.get(key)
    if h is None:
        with _lock:
            h = _plans.get(key)
            if h is None:
                handle = c_void_p()
                check(load().pdephi_plan_create_slab(ctypes.byref(handle), key[0], *key[1], *key[2], *key[3]), "plan_create")
                h = handle
                _plans[key] = h
    return h


_auto_route = {}
_ws_bytes = {}


def plan_route(plan, precision: int) -> int:
    """pdephi_plan_route(plan, precision), cached (at the small shapes of cfg 1 a layer is launch-latency-bound and every
    ctypes round trip counts)."""
    key = (plan.value, precision)
    r = _auto_route.get(key)
    if r is None:
        r = load().pdephi_plan_route(plan, precision)
        _auto_route[key] = r
    return r


def resolve_precision(plan, precision: int, points: int = 0) -> int:
    """PREC_AUTO -> the fastest route that meets the fp32 parity gate for this plan's shape and `points` (= volumes x grid
    points of the call; 0 = unknown / large).  The fused split-TF32 kernels win wherever they apply; the per-axis split
    kernels run every axis stage as three launches and beat the FFMA kernels only once the call is large enough to hide
    them (measured on the B200: [4,32,32^3] FFMA 0.39 ms vs 0.96 ms per layer, [1,64,128^3] modes 64^3 the other way)."""
    if precision != PREC_AUTO:
        return precision
    route = plan_route(plan, PREC_TF32X3)
    if route == ROUTE_FUSED_TC or (route == ROUTE_PER_AXIS_TC and (points == 0 or points >= PER_AXIS_SPLIT_MIN_POINTS)):
        return PREC_TF32X3
    return PREC_FP32


def workspace_bytes(plan, BC: int) -> int:
    """pdephi_plan_workspace_bytes(plan, BC), cached."""
    key = (plan.value, int(BC))
    n = _ws_bytes.get(key)
    if n is None:
        n = load().pdephi_plan_workspace_bytes(plan, BC)
        _ws_bytes[key] = n
    return n


class _NoDeviceSwitch:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_no_switch = _NoDeviceSwitch()


def on_device(device):
    """Context manager that makes `device` current -- a no-op object when it already is (torch.cuda.device costs a few
    microseconds per use, which adds up over the ~10 library calls of one small layer)."""
    import torch
    return _no_switch if torch.cuda.current_device() == device.index else torch.cuda.device(device)


def set_option(key: int, value: int) -> int:
    """pdephi_set_option; returns the previous value."""
    lib = load()
    old = lib.pdephi_get_option(key)
    check(lib.pdephi_set_option(key, int(value)), "set_option")
    return old


def profile(enable: bool = True, reset: bool = True):
    lib = load()
    if reset:
        lib.pdephi_profile_reset()
    lib.pdephi_profile_enable(1 if enable else 0)


def profile_report(timed: bool = True):
    """{kernel name: (launches, total device ms)} since the last reset (synchronises on recorded events)."""
    lib = load()
    out = {}
    for i in range(lib.pdephi_profile_num_kernels()):
        n = c_longlong(0)
        ms = ctypes.c_double(0.0)
        check(lib.pdephi_profile_query(i, ctypes.byref(n), ctypes.byref(ms) if timed else None), "profile_query")
        if n.value:
            out[lib.pdephi_profile_kernel_name(i).decode()] = (n.value, ms.value)
    return out
