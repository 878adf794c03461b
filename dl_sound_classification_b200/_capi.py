"""ctypes binding of ``include/b200fbank.h`` (the C-ABI drop-in boundary).

The shared library is built in-tree by ``__graft_entry__.build()`` into
``dl_sound_classification_b200/lib/libb200fbank.so``.  There is no CPU
fallback: if the library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200FBANK_LIB") or os.path.join(_HERE, "lib", "libb200fbank.so")   # env: kernel experiments (tools/ktime.py)
ABI_VERSION = 1
MAX_RATES = 8

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 0, -1, -2, -3, -4
WINDOW_TYPES = {"povey": 0, "hanning": 1, "hamming": 2, "rectangular": 3, "blackman": 4}
LAYOUT_BTF, LAYOUT_BFT = 0, 1
FRONTEND_KALDI_FBANK, FRONTEND_MELSPEC_DB = 0, 1
TABLE_WINDOW, TABLE_MEL_DENSE, TABLE_TAPS_DENSE = 0, 1, 2


class Opts(C.Structure):
    """``b200fbank_opts`` -- field order must match include/b200fbank.h."""
    _fields_ = [
        ("blackman_coeff", C.c_double), ("energy_floor", C.c_double), ("frame_length", C.c_double),
        ("frame_shift", C.c_double), ("high_freq", C.c_double), ("low_freq", C.c_double),
        ("preemphasis_coefficient", C.c_double), ("sample_frequency", C.c_double),
        ("vtln_high", C.c_double), ("vtln_low", C.c_double), ("vtln_warp", C.c_double),
        ("num_mel_bins", C.c_int32), ("window_type", C.c_int32), ("htk_compat", C.c_int32),
        ("raw_energy", C.c_int32), ("remove_dc_offset", C.c_int32), ("round_to_power_of_two", C.c_int32),
        ("snip_edges", C.c_int32), ("subtract_mean", C.c_int32), ("use_energy", C.c_int32),
        ("use_log_fbank", C.c_int32), ("use_power", C.c_int32),
        ("n_rates", C.c_int32), ("orig_rates", C.c_int32 * MAX_RATES), ("lowpass_filter_width", C.c_int32),
        ("rolloff", C.c_double),
        ("frontend", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32), ("win_length", C.c_int32),
        ("top_db", C.c_double),
    ]


class B200FbankError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200fbank error {code}: {msg}")
        self.code = code
        self.msg = msg


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "dl_sound_classification_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    P, I64, I32, F32, F64, VP = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_double, C.c_void_p
    lib.b200fbank_abi_version.restype = C.c_int
    lib.b200fbank_default_opts.argtypes = [C.POINTER(Opts)]
    lib.b200fbank_default_opts.restype = None
    lib.b200fbank_plan_create.argtypes = [C.POINTER(Opts), C.c_int, C.POINTER(P)]
    lib.b200fbank_plan_create.restype = C.c_int
    lib.b200fbank_plan_destroy.argtypes = [P]
    lib.b200fbank_plan_destroy.restype = None
    lib.b200fbank_last_error.restype = C.c_char_p
    lib.b200fbank_resampled_length.argtypes = [P, I64, C.c_int]
    lib.b200fbank_resampled_length.restype = I64
    lib.b200fbank_num_frames.argtypes = [P, I64, C.c_int]
    lib.b200fbank_num_frames.restype = I64
    lib.b200fbank_num_cols.argtypes = [P]
    lib.b200fbank_num_cols.restype = C.c_int
    lib.b200fbank_plan_table.argtypes = [P, C.c_int, C.c_int, C.POINTER(F32), I64]
    lib.b200fbank_plan_table.restype = I64
    lib.b200fbank_plan_info.argtypes = [P, C.c_int, C.POINTER(I64)]
    lib.b200fbank_plan_info.restype = C.c_int
    lib.b200fbank_execute.argtypes = [P, VP, VP, I64, VP, C.c_int, VP, VP, VP, C.c_int, F32, F32, C.c_int,
                                      C.c_int, VP, VP, VP]
    lib.b200fbank_execute.restype = C.c_int
    lib.b200fbank_execute_mixup.argtypes = [P, VP, VP, I64, VP, C.c_int, VP, VP, VP, C.c_int, F32, F32, C.c_int,
                                            C.c_int, VP, VP, VP, VP, VP, VP]
    lib.b200fbank_execute_mixup.restype = C.c_int
    lib.b200fbank_melspec_db.argtypes = [P, VP, VP, I64, VP, C.c_int, VP, C.c_int, C.c_int, F32, F32, C.c_int, C.c_int,
                                         VP, VP, VP, VP]
    lib.b200fbank_melspec_db.restype = C.c_int
    lib.b200fbank_clip_normalize.argtypes = [VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, VP, F32, C.c_int, F32, F32, VP, VP]
    lib.b200fbank_clip_normalize.restype = C.c_int
    lib.b200fbank_remove_clip_mean.argtypes = [VP, VP, I64, C.c_int, VP, VP, VP]
    lib.b200fbank_remove_clip_mean.restype = C.c_int
    lib.b200fbank_pcm16_to_float.argtypes = [VP, VP, I64, C.c_int, VP, I64, VP, VP]
    lib.b200fbank_pcm16_to_float.restype = C.c_int
    lib.b200fbank_stats_accumulate.argtypes = [P, VP, VP, I64, VP, C.c_int, C.c_int, VP, VP]
    lib.b200fbank_stats_accumulate.restype = C.c_int
    lib.b200fbank_resample.argtypes = [P, VP, VP, I64, VP, C.c_int, VP, VP, I64, VP]
    lib.b200fbank_resample.restype = C.c_int
    lib.b200fbank_mixup.argtypes = [VP, VP, VP, VP, C.c_int, I64, VP, VP]
    lib.b200fbank_mixup.restype = C.c_int
    lib.b200fbank_mixup_labels.argtypes = [VP, VP, VP, VP, C.c_int, C.c_int, VP, VP]
    lib.b200fbank_mixup_labels.restype = C.c_int
    lib.b200fbank_patch_embed.argtypes = [VP, C.c_int, C.c_int, C.c_int, VP, VP, C.c_int, C.c_int, C.c_int, VP, C.c_int, VP]
    lib.b200fbank_patch_embed.restype = C.c_int
    lib.b200fbank_launch_count.argtypes = [C.c_int]
    lib.b200fbank_launch_count.restype = I64
    lib.b200fbank_sizeof_opts.restype = C.c_int
    v = lib.b200fbank_abi_version()
    if v != ABI_VERSION:
        raise ImportError(f"libb200fbank ABI {v} != expected {ABI_VERSION}; rebuild")
    if lib.b200fbank_sizeof_opts() != C.sizeof(Opts):
        raise ImportError("b200fbank_opts layout mismatch between include/b200fbank.h and _capi.Opts; rebuild")
    return lib


lib = _load()

EXPORTED_SYMBOLS = [
    "b200fbank_abi_version", "b200fbank_sizeof_opts", "b200fbank_default_opts", "b200fbank_plan_create", "b200fbank_plan_destroy",
    "b200fbank_last_error", "b200fbank_mixup", "b200fbank_mixup_labels", "b200fbank_patch_embed", "b200fbank_resampled_length", "b200fbank_num_frames", "b200fbank_num_cols",
    "b200fbank_plan_table", "b200fbank_plan_info", "b200fbank_execute", "b200fbank_execute_mixup", "b200fbank_melspec_db",
    "b200fbank_stats_accumulate", "b200fbank_clip_normalize", "b200fbank_remove_clip_mean", "b200fbank_pcm16_to_float",
    "b200fbank_resample", "b200fbank_launch_count",
]


def check(rc: int) -> int:
    if rc < 0:
        msg = lib.b200fbank_last_error().decode("utf-8", "replace")
        if rc == ERR_INVALID:
            raise ValueError(msg)
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        raise B200FbankError(rc, msg)
    return rc


def default_opts() -> Opts:
    o = Opts()
    lib.b200fbank_default_opts(C.byref(o))
    return o


class Plan:
    """Owning wrapper of a ``b200fbank_plan*``.  ``device=-1`` -> host-only plan."""

    def __init__(self, opts: Opts, device: int):
        self._h = C.c_void_p()
        self.device = device
        self.opts = opts
        check(lib.b200fbank_plan_create(C.byref(opts), device, C.byref(self._h)))
        info = (C.c_int64 * 8)()
        check(lib.b200fbank_plan_info(self._h, 0, info))
        self.window_shift, self.window_size, self.padded_window_size, self.n_rates = (int(info[i]) for i in range(4))
        self.n_cols = int(lib.b200fbank_num_cols(self._h))

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.b200fbank_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rate_info(self, rate_id: int):
        info = (C.c_int64 * 8)()
        check(lib.b200fbank_plan_info(self._h, rate_id, info))
        return dict(orig=int(info[4]), new=int(info[5]), width=int(info[6]), taps_per_phase=int(info[7]))

    def resampled_length(self, n: int, rate_id: int = 0) -> int:
        return int(check(lib.b200fbank_resampled_length(self._h, n, rate_id)))

    def num_frames(self, n: int, rate_id: int = 0) -> int:
        return int(check(lib.b200fbank_num_frames(self._h, n, rate_id)))

    def table(self, which: int, arg: int = 0):
        import numpy as np
        n = int(check(lib.b200fbank_plan_table(self._h, which, arg, None, 0)))
        out = np.empty(n, dtype=np.float32)
        check(lib.b200fbank_plan_table(self._h, which, arg, out.ctypes.data_as(C.POINTER(C.c_float)), n))
        return out
