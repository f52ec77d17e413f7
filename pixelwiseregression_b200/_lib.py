"""ctypes binding of libpwr_b200.so (include/pwr.h).

The library is the product: there is no CPU or PyTorch fallback.  Loading
fails loudly when the shared object is missing, and every wrapper refuses
non-CUDA tensors.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# PWR_LIB_PATH selects another build of the same ABI (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("PWR_LIB_PATH") or os.path.join(_PKG, "libpwr_b200.so")

METHOD_SOFTMAX, METHOD_SUM, METHOD_GIVEN = 0, 1, 2
METHODS = {"softmax": METHOD_SOFTMAX, "sum": METHOD_SUM, "given": METHOD_GIVEN}
FRAME_FORMATS = {"f32": 0, "nyu_gb16": 1, "u16": 2}
MAP_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}

_P = ctypes.c_void_p
_I = ctypes.c_int
_D = ctypes.c_double
_F = ctypes.c_float
_LL = ctypes.c_longlong
_SZ = ctypes.c_size_t

# name -> argtypes, exactly the prototypes of include/pwr.h
SIGNATURES = {
    "pwr_version": [],
    "pwr_error_string": [_I],
    "pwr_set_option": [_I, _I],
    "pwr_sfr_com": [_P, _I, _I, _P, _I, _P],
    "pwr_sfr_bb_filter": [_P, _I, _I, _P, _P, _I, _P],
    "pwr_sfr_workspace_bytes": [_I, _I],
    "pwr_sfr_crop": [_P, _I, _I, _I, _P, _P, _D, _D, _I, _D, _D, _D, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _I, _I,
                     _I, _P],
    "pwr_sfr_build": [_P, _I, _I, _I, _P, _P, _P, _P, _D, _D, _I, _D, _D, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                      _P, _P, _SZ, _P, _I, _I, _I, _I, _P],
    "pwr_sfr_fetch": [_P, _I, _I, _I, _P, _P, _P, _D, _D, _D, _D, _D, _P, _I, _I, _P, _P, _P, _I, _P],
    "pwr_decoder_fwd": [_P] * 13 + [_I, _I, _I, _I, _P],
    "pwr_decoder_bwd": [_P] * 13 + [_I, _I, _I, _I, _P],
    "pwr_decoder_bwd_loss": [_P] * 14 + [_F, _F, _F, _F, _P, _I] + [_P] * 4 + [_I, _I, _I, _I, _P],
    "pwr_decoder_fwd_bwd_loss": [_P] * 9 + [_F, _F, _F, _F, _P, _I] + [_P] * 6 + [_I, _I, _I, _I, _P],
    "pwr_reduce_partials": [_P, _P, _I, _I, _I, _P],
    "pwr_stage_loss": [_P, _I, _I, _F, _F, _F, _I, _P, _P, _P, _P],
    "pwr_scale_inplace": [_P, _P, _P, _I, _P, _LL, _I, _P],
    "pwr_recover_uvd": [_P, _P, _P, _P, _D, _D, _D, _D, _P, _P, _I, _I, _P],
    "pwr_joint_error": [_P, _P, _P, _P, _P, _D, _D, _D, _D, _P, _I, _I, _P],
}

ABI_VERSION = 201      # PWR_VERSION of include/pwr.h this binding is written against
OPTIONS = {"bwd_direct": 0, "fwd_direct": 1, "fwd_pipe": 2, "bwd_no_lean": 3, "sfr_staged": 4, "fused_no_lean": 5}
_ENV_OPTIONS = {"PWR_BWD_DIRECT": ("bwd_direct", "1"), "PWR_FWD_DIRECT": ("fwd_direct", "1"),
                "PWR_FWD_PIPE": ("fwd_pipe", "1"), "PWR_BWD_LEAN": ("bwd_no_lean", "0"),
                "PWR_SFR_STAGED": ("sfr_staged", "1")}

_lib = None


class PwrError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if "PWR_LIB_PATH" not in os.environ:
        from . import build as _build
        if _build.needs_build():               # missing, or built from other sources (stale ABI): compile, never fall back
            try:
                _build.build()
            except Exception as exc:
                raise PwrError("libpwr_b200.so at %s is missing or stale and building it failed (%s): run `python "
                               "-m pixelwiseregression_b200.build` (there is no CPU fallback)" % (LIB_PATH, exc))
    if not os.path.isfile(LIB_PATH):
        raise PwrError(
            "libpwr_b200.so not found at %s: build it with `python -m pixelwiseregression_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = {"pwr_error_string": ctypes.c_char_p, "pwr_sfr_workspace_bytes": _SZ}.get(name, _I)
    if lib.pwr_version() != ABI_VERSION:
        raise PwrError("libpwr_b200.so at %s has ABI version %d, this binding is written for %d: rebuild it with "
                       "`python -m pixelwiseregression_b200.build --force`" % (LIB_PATH, lib.pwr_version(), ABI_VERSION))
    # dispatch overrides: the environment is read here, once, never on the launch path
    for env, (opt, on) in _ENV_OPTIONS.items():
        val = os.environ.get(env)
        if val is not None:
            lib.pwr_set_option(OPTIONS[opt], 1 if val[:1] == on else 0)
    _lib = lib
    return lib


def set_option(name, value):
    """Dispatch override of include/pwr.h (A/B measurements, tests): returns the previous value."""
    prev = load().pwr_set_option(OPTIONS[name], int(value))
    if prev < 0:
        raise PwrError("unknown option %r" % (name,))
    return prev


class option:
    """`with _lib.option("bwd_no_lean", 1): ...` — scoped dispatch override."""

    def __init__(self, name, value):
        self.name, self.value = name, value

    def __enter__(self):
        self.prev = set_option(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.name, self.prev)
        return False


LAUNCHES = {}          # entry point -> number of kernels launched through it
KERNELS_PER_CALL = {"pwr_sfr_build": 2, "pwr_sfr_crop": 2, "pwr_sfr_fetch": 2}    # prep + main; every other entry point is one kernel
PROFILE = None         # when a list: (entry point, start event, end event) per launch


def check(rc, what):
    if rc != 0:
        msg = load().pwr_error_string(rc)
        raise PwrError("%s failed: rc=%d (%s)" % (what, rc, msg.decode() if msg else "?"))
    LAUNCHES[what] = LAUNCHES.get(what, 0) + KERNELS_PER_CALL.get(what, 1)


def launch_count():
    return sum(LAUNCHES.values())


class launch:
    """`with launch(device, "pwr_xyz"):` — everything a launch needs around the ctypes call, at the lowest host
    cost: make `device` current only if it is not already (torch.cuda.device() costs ~5 us per use), and, when
    profiling is switched on (bench.py), bracket the launch with CUDA events on the current stream."""
    __slots__ = ("idx", "what", "prev", "start", "end")

    def __init__(self, device, what=None):
        self.idx = device.index
        self.what = what

    def __enter__(self):
        self.prev = None
        if self.idx is not None:
            cur = torch._C._cuda_getDevice()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        if PROFILE is not None and self.what is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and self.what is not None and exc[0] is None:
            self.end.record()
            PROFILE.append((self.what, self.start, self.end))
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def ptr(t):
    """Device pointer of a tensor as an int (None -> NULL); ctypes converts it for the c_void_p parameters."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (raw handle, no Stream object built)."""
    idx = device.index if (device is not None and device.index is not None) else torch._C._cuda_getDevice()
    return torch._C._cuda_getCurrentRawStream(idx)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PwrError("pixelwiseregression_b200 runs on CUDA tensors only (got a %s tensor); "
                           "there is no CPU fallback" % t.device.type)


def as_f32(t):
    """Contiguous float32 view/copy on the same device (fp16/bf16 under autocast)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
