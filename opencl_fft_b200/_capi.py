"""ctypes binding of include/b200fft.h (libb200fft.so). Nothing here computes: every call goes to the
CUDA engine. If the library is missing this module raises -- there is no fallback path."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2F_LIB_PATH") or os.path.join(_PKG, "lib", "libb200fft.so")  # (override: A/B builds)

# every symbol include/b200fft.h declares: name -> (restype, argtypes)
_vp, _i, _sz, _fp = C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_float)
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "b2f_error_string": (C.c_char_p, [_i]),
    "b2f_last_cuda_error": (C.c_char_p, []),
    "b2f_version": (C.c_char_p, []),
    "b2f_set_option": (_i, [C.c_char_p, C.c_longlong]),
    "b2f_get_option": (_i, [C.c_char_p, C.POINTER(C.c_longlong)]),
    "b2f_device_count": (_i, [C.POINTER(_i)]),
    "b2f_device_name": (_i, [_i, C.c_char_p, _sz]),
    "b2f_cfft_create": (_i, [_pp, _i, _i, _i, _i]),
    "b2f_cfft_destroy": (_i, [_vp]),
    "b2f_cfft_exec_host": (_i, [_vp, _vp, _i]),
    "b2f_cfft_exec_dev": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b2f_rfft_create": (_i, [_pp, _i, _i, _i, _i]),
    "b2f_rfft_destroy": (_i, [_vp]),
    "b2f_rfft_exec_host": (_i, [_vp, _vp, _vp, _i]),
    "b2f_rfft_exec_dev": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b2f_pconv_create": (_i, [_pp, _i, _i, _i, _i]),
    "b2f_pconv_destroy": (_i, [_vp]),
    "b2f_pconv_nparts": (_i, [_vp]),
    "b2f_pconv_reset": (_i, [_vp]),
    "b2f_pconv_push_ir_host": (_i, [_vp, _vp, _sz]),
    "b2f_pconv_push_ir_dev": (_i, [_vp, _vp, _sz, _vp]),
    "b2f_pconv_process_host": (_i, [_vp, _vp, _vp]),
    "b2f_pconv_process_dev": (_i, [_vp, _vp, _vp, _vp]),
    "b2f_pconv_process_tv_host": (_i, [_vp, _vp, _vp, _vp]),
    "b2f_pconv_process_tv_dev": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b2f_pconv_read_spectra": (_i, [_vp, _i, _i, _vp]),
    "b2f_dconv_create": (_i, [_pp, _i, _i, _i, _i, _i]),
    "b2f_dconv_destroy": (_i, [_vp]),
    "b2f_dconv_reset": (_i, [_vp]),
    "b2f_dconv_push_ir_host": (_i, [_vp, _vp, _sz]),
    "b2f_dconv_push_ir_dev": (_i, [_vp, _vp, _sz, _vp]),
    "b2f_dconv_process_host": (_i, [_vp, _vp, _vp, _i]),
    "b2f_dconv_process_dev": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b2f_dconv_process_tv_host": (_i, [_vp, _vp, _vp, _vp]),
    "b2f_dconv_process_tv_dev": (_i, [_vp, _vp, _vp, _vp, _vp]),
    # several GPUs behind one handle
    "b2f_pconv_multi_create": (_i, [_pp, C.POINTER(_i), _i, _i, _i, _i]),
    "b2f_pconv_multi_destroy": (_i, [_vp]),
    "b2f_pconv_multi_nparts": (_i, [_vp]),
    "b2f_pconv_multi_reset": (_i, [_vp]),
    "b2f_pconv_multi_push_ir_host": (_i, [_vp, _vp, _sz]),
    "b2f_pconv_multi_push_ir_shard_host": (_i, [_vp, _i, _vp, _sz]),
    "b2f_pconv_multi_process_host": (_i, [_vp, _vp, _vp]),
    "b2f_pconv_multi_process_tv_host": (_i, [_vp, _vp, _vp, _vp]),
    "b2f_dconv_multi_create": (_i, [_pp, C.POINTER(_i), _i, _i, _i, _i, _i]),
    "b2f_dconv_multi_destroy": (_i, [_vp]),
    "b2f_dconv_multi_reset": (_i, [_vp]),
    "b2f_dconv_multi_push_ir_host": (_i, [_vp, _vp, _sz]),
    "b2f_dconv_multi_process_host": (_i, [_vp, _vp, _vp, _i]),
    "b2f_dconv_multi_process_tv_host": (_i, [_vp, _vp, _vp, _vp]),
    "b2f_cfft_multi_create": (_i, [_pp, C.POINTER(_i), _i, _i, _i, _i]),
    "b2f_cfft_multi_destroy": (_i, [_vp]),
    "b2f_cfft_multi_exec_host": (_i, [_vp, _vp, _i]),
    "b2f_rfft_multi_create": (_i, [_pp, C.POINTER(_i), _i, _i, _i, _i]),
    "b2f_rfft_multi_destroy": (_i, [_vp]),
    "b2f_rfft_multi_exec_host": (_i, [_vp, _vp, _vp, _i]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libb200fft.so and bind every declared symbol. Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA engine first (python -m opencl_fft_b200.build). "
                "opencl_fft_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def error_string(code: int) -> str:
    return lib().b2f_error_string(int(code)).decode()


def last_cuda_error() -> str:
    return lib().b2f_last_cuda_error().decode()


class B2fError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        super().__init__(f"{where}: {error_string(code)} [{code}] {last_cuda_error()}")
