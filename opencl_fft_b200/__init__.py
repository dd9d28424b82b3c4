"""opencl_fft_b200 -- host-side mirror of the opencl_fft class interface over the B200 CUDA engine.

The classes carry the reference's names, constructor arguments and method names
(reference cl_fft.h:29-111, cl_conv.h:124-188, cl_dconv.h:17-66):

    Clcfft(device, size, fwd=True)        .transform(c)            .get_error()
    Clrfft(device, size, fwd)             .transform(c, r=None)    .get_error()
    Clpconv(device, cvs, pts)             .push_ir(ir) .convolution(out, in1[, in2]) .get_cl_err()
    Cldconv(device, cvs, vsize)           .push_ir(ir) .convolution(out, in1[, in2]) .get_cl_err()

with three extensions the reference does not have: `max_batch` / `channels` (many transforms or
convolver channels per object, one launch), `*_dev` methods that take device pointers (anything
with `.data_ptr()`, e.g. torch CUDA tensors, or raw ints) plus a CUDA stream handle, and `devices=[...]`
(the channels / transforms of ONE object sharded over several GPUs in contiguous ranges, one host thread
and stream per device, no communication; host-pointer methods only). Methods return the
engine's status code (0 = success) exactly like the reference's methods return cl_int.

Everything runs in libb200fft.so (hand-written sm_100a kernels). There is no CPU fallback: importing
this package without the built library raises, and constructing an object without a CUDA device
leaves a non-zero get_error()/get_cl_err() like the reference does without an OpenCL device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import B2fError, error_string, last_cuda_error, lib  # noqa: F401

PI = 3.141592653589793  # cl_fft::PI, reference cl_fft.h:24

__all__ = ["Clcfft", "Clrfft", "Clpconv", "Cldconv", "device_count", "device_name", "cl_error_string",
           "B2fError", "PI", "set_option", "get_option"]

INVALID_VALUE = 2  # B2F_ERR_INVALID_VALUE


def set_option(name: str, value: int) -> None:
    """Process-wide default copied by handles created afterwards (include/b200fft.h lists the names)."""
    rc = lib().b2f_set_option(name.encode(), int(value))
    if rc:
        raise B2fError(rc, f"b2f_set_option({name})")


def get_option(name: str) -> int:
    v = C.c_longlong(0)
    rc = lib().b2f_get_option(name.encode(), C.byref(v))
    if rc:
        raise B2fError(rc, f"b2f_get_option({name})")
    return v.value


def cl_error_string(err: int) -> str:
    """cl_fft::cl_error_string / cl_conv::cl_string (reference cl_fft.h:25, cl_conv.h:25)."""
    return error_string(err)


def device_count() -> int:
    n = C.c_int(0)
    lib().b2f_device_count(C.byref(n))
    return n.value


def device_name(device: int = 0) -> str:
    buf = C.create_string_buffer(256)
    rc = lib().b2f_device_name(device, buf, 256)
    if rc:
        raise B2fError(rc, "b2f_device_name")
    return buf.value.decode()


def _devlist(devices):
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    return arr, len(devices)


def _no_multi(obj):
    if obj._multi:
        raise TypeError("device-pointer methods take pointers on ONE device: not available with devices=[...]")


def _host(a, dtype, what: str) -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous:
        raise TypeError(f"{what} must be a C-contiguous numpy array of {np.dtype(dtype).name}")
    return a


def _dptr(x) -> int:
    """device pointer of a torch tensor / cupy array / int"""
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


def _stream(stream) -> int:
    if stream is None:
        try:
            import torch

            return int(torch.cuda.current_stream().cuda_stream)
        except Exception:
            return 0
    if hasattr(stream, "cuda_stream"):
        return int(stream.cuda_stream)
    return int(stream)


class Clcfft:
    """Complex-to-complex FFT (reference cl_fft::Clcfft). forward is scaled by 1/N, inverse unscaled."""

    def __init__(self, device: int, size: int, fwd: bool = True, max_batch: int = 1, devices=None):
        self.N, self.forward, self.max_batch = size, bool(fwd), max_batch
        self._h = C.c_void_p()
        self._multi = devices is not None
        if self._multi:
            arr, n = _devlist(devices)
            self._err = lib().b2f_cfft_multi_create(C.byref(self._h), arr, n, size, int(bool(fwd)), max_batch)
        else:
            self._err = lib().b2f_cfft_create(C.byref(self._h), device, size, int(bool(fwd)), max_batch)
        self._log = "" if self._err == 0 else f"{error_string(self._err)} ({last_cuda_error()})"

    def get_error(self) -> int:
        return self._err

    def get_log(self) -> str:
        return self._log

    def transform(self, c: np.ndarray) -> int:
        """In place on a complex64 array of N (or batch*N, batch <= max_batch) points."""
        c = _host(c, np.complex64, "c")
        if self._err:
            return self._err
        if c.size % self.N:
            return INVALID_VALUE
        fn = lib().b2f_cfft_multi_exec_host if self._multi else lib().b2f_cfft_exec_host
        return fn(self._h, c.ctypes.data, c.size // self.N)

    def transform_dev(self, d_in, d_out, batch: int, stream=None) -> int:
        """[batch][N] complex64 device arrays (may alias); asynchronous on `stream`."""
        _no_multi(self)
        return lib().b2f_cfft_exec_dev(self._h, _dptr(d_in), _dptr(d_out), batch, _stream(stream))

    def close(self):
        if C is None:  # interpreter shutdown: the module globals are already gone
            return
        if getattr(self, "_h", None) and self._h.value:
            (lib().b2f_cfft_multi_destroy if self._multi else lib().b2f_cfft_destroy)(self._h)
            self._h = C.c_void_p()

    __del__ = close


class Clrfft:
    """Real FFT of `size` points (reference cl_fft::Clrfft): size/2 packed complex bins, element 0 =
    (DC, Nyquist)/size, element k = 2 X[k]/size, bin size/4 conjugated (reference quirk)."""

    def __init__(self, device: int, size: int, fwd: bool, max_batch: int = 1, devices=None):
        self.size, self.N, self.forward, self.max_batch = size, size // 2, bool(fwd), max_batch
        self._h = C.c_void_p()
        self._multi = devices is not None
        if self._multi:
            arr, n = _devlist(devices)
            self._err = lib().b2f_rfft_multi_create(C.byref(self._h), arr, n, size, int(bool(fwd)), max_batch)
        else:
            self._err = lib().b2f_rfft_create(C.byref(self._h), device, size, int(bool(fwd)), max_batch)
        self._log = "" if self._err == 0 else f"{error_string(self._err)} ({last_cuda_error()})"

    def get_error(self) -> int:
        return self._err

    def get_log(self) -> str:
        return self._log

    def transform(self, c: np.ndarray, r: np.ndarray | None = None) -> int:
        """c: complex64 [batch*size/2]; r: float32 [batch*size] or None for the in-place form
        (c's memory viewed as reals). Forward reads r, writes c; inverse reads c, writes r and c."""
        c = _host(c, np.complex64, "c")
        if self._err:
            return self._err
        rp = c.ctypes.data if r is None else _host(r, np.float32, "r").ctypes.data
        if c.size % self.N or (r is not None and r.size != 2 * c.size):
            return INVALID_VALUE  # the engine would read / write size reals per transform behind r
        fn = lib().b2f_rfft_multi_exec_host if self._multi else lib().b2f_rfft_exec_host
        return fn(self._h, c.ctypes.data, rp, c.size // self.N)

    def transform_dev(self, d_in, d_out, batch: int, stream=None) -> int:
        _no_multi(self)
        return lib().b2f_rfft_exec_dev(self._h, _dptr(d_in), _dptr(d_out), batch, _stream(stream))

    def close(self):
        if C is None:  # interpreter shutdown: the module globals are already gone
            return
        if getattr(self, "_h", None) and self._h.value:
            (lib().b2f_rfft_multi_destroy if self._multi else lib().b2f_rfft_destroy)(self._h)
            self._h = C.c_void_p()

    __del__ = close


class Clpconv:
    """Uniformly-partitioned convolution (reference cl_conv::Clpconv), `channels` independent convolvers.
    Arrays are [channels][...] C-contiguous float32; channels=1 is the reference object."""

    def __init__(self, device: int, cvs: int, pts: int, errs=None, uData=None, channels: int = 1, devices=None):
        self.cvs, self.pts, self.channels = cvs, pts, channels
        self._errs, self._udata = errs, uData
        self._h = C.c_void_p()
        self._multi = devices is not None
        self._devices = list(devices) if devices is not None else [device]
        L = lib()
        if self._multi:
            arr, n = _devlist(devices)
            self._err = L.b2f_pconv_multi_create(C.byref(self._h), arr, n, cvs, pts, channels)
            self._f = (L.b2f_pconv_multi_push_ir_host, L.b2f_pconv_multi_process_host, L.b2f_pconv_multi_process_tv_host,
                       L.b2f_pconv_multi_reset, L.b2f_pconv_multi_destroy)
            self.nparts = L.b2f_pconv_multi_nparts(self._h) if self._err == 0 else 0
        else:
            self._err = L.b2f_pconv_create(C.byref(self._h), device, cvs, pts, channels)
            self._f = (L.b2f_pconv_push_ir_host, L.b2f_pconv_process_host, L.b2f_pconv_process_tv_host,
                       L.b2f_pconv_reset, L.b2f_pconv_destroy)
            self.nparts = L.b2f_pconv_nparts(self._h) if self._err == 0 else 0
        if self._err:
            self._msg(error_string(self._err))

    def _msg(self, s: str):
        # reference cl_conv.h:137-145: default handler prints when no user data was given
        if self._errs is not None:
            self._errs(s, self._udata)
        elif self._udata is None:
            print(s)

    def get_cl_err(self) -> int:
        return self._err

    def cl_error_string(self, err: int) -> str:
        return error_string(err)

    def push_ir(self, ir: np.ndarray) -> int:
        """ir: [channels][>= nparts*pts] float32 (row stride = ir.shape[-1])."""
        ir = _host(ir, np.float32, "ir")
        stride = ir.shape[-1] if ir.ndim > 1 else ir.size // self.channels
        # the engine reads nparts*pts floats from each of `channels` rows `stride` apart
        if stride < self.nparts * self.pts or ir.size < (self.channels - 1) * stride + self.nparts * self.pts:
            self._err = INVALID_VALUE
            return self._err
        self._err = self._f[0](self._h, ir.ctypes.data, stride)
        return self._err

    def push_ir_shard(self, g: int, ir: np.ndarray) -> int:
        """devices=[...] only: the IRs of the channels that live on devices[g] (contiguous range g of len(devices))."""
        if not self._multi:
            raise TypeError("push_ir_shard needs devices=[...]")
        ir = _host(ir, np.float32, "ir")
        n = len(self._devices)
        count = (g + 1) * self.channels // n - g * self.channels // n
        stride = ir.shape[-1] if ir.ndim > 1 else ir.size // count
        if stride < self.nparts * self.pts or ir.size < (count - 1) * stride + self.nparts * self.pts:
            return INVALID_VALUE
        return lib().b2f_pconv_multi_push_ir_shard_host(self._h, g, ir.ctypes.data, stride)

    def convolution(self, output: np.ndarray, input1: np.ndarray, input2: np.ndarray | None = None) -> int:
        out = _host(output, np.float32, "output")
        a = _host(input1, np.float32, "input1")
        need = self.channels * self.pts  # one block per channel is read and written
        if out.size < need or a.size < need or (input2 is not None and np.size(input2) < need):
            self._err = INVALID_VALUE
            return self._err
        if input2 is None:
            self._err = self._f[1](self._h, out.ctypes.data, a.ctypes.data)
        else:
            b = _host(input2, np.float32, "input2")
            self._err = self._f[2](self._h, out.ctypes.data, a.ctypes.data, b.ctypes.data)
        return self._err

    def push_ir_dev(self, d_ir, ir_stride: int, stream=None) -> int:
        _no_multi(self)
        return lib().b2f_pconv_push_ir_dev(self._h, _dptr(d_ir), ir_stride, _stream(stream))

    def convolution_dev(self, d_out, d_in1, d_in2=None, stream=None) -> int:
        _no_multi(self)
        if d_in2 is None:
            return lib().b2f_pconv_process_dev(self._h, _dptr(d_out), _dptr(d_in1), _stream(stream))
        return lib().b2f_pconv_process_tv_dev(self._h, _dptr(d_out), _dptr(d_in1), _dptr(d_in2), _stream(stream))

    def reset(self) -> int:
        return self._f[3](self._h)

    def read_spectra(self, which: int, channel: int = 0) -> np.ndarray:
        """white-box: FDL (which=1, reference spec1) or IR spectra (which=2, spec2) of one channel"""
        _no_multi(self)
        out = np.empty(self.nparts * self.pts, np.complex64)
        rc = lib().b2f_pconv_read_spectra(self._h, which, channel, out.ctypes.data)
        if rc:
            raise B2fError(rc, "b2f_pconv_read_spectra")
        return out

    def close(self):
        if C is None:  # interpreter shutdown: the module globals are already gone
            return
        if getattr(self, "_h", None) and self._h.value:
            self._f[4](self._h)
            self._h = C.c_void_p()

    __del__ = close


class Cldconv:
    """Direct time-domain convolution (reference cl_conv::Cldconv), `channels` independent convolvers.
    y[t] = sum_c ir[c] x[t-1-c] (one-sample delay, reference quirk SURVEY Q9)."""

    def __init__(self, device: int, cvs: int, vsize: int, errs=None, uData=None, channels: int = 1,
                 max_blocks: int = 1, devices=None):
        self.irsize, self.vsize, self.channels, self.max_blocks = cvs, vsize, channels, max_blocks
        self._errs, self._udata = errs, uData
        self._h = C.c_void_p()
        self._multi = devices is not None
        L = lib()
        if self._multi:
            arr, n = _devlist(devices)
            self._err = L.b2f_dconv_multi_create(C.byref(self._h), arr, n, cvs, vsize, channels, max_blocks)
            self._f = (L.b2f_dconv_multi_push_ir_host, L.b2f_dconv_multi_process_host, L.b2f_dconv_multi_process_tv_host,
                       L.b2f_dconv_multi_reset, L.b2f_dconv_multi_destroy)
        else:
            self._err = L.b2f_dconv_create(C.byref(self._h), device, cvs, vsize, channels, max_blocks)
            self._f = (L.b2f_dconv_push_ir_host, L.b2f_dconv_process_host, L.b2f_dconv_process_tv_host,
                       L.b2f_dconv_reset, L.b2f_dconv_destroy)
        if self._err:
            self._msg(error_string(self._err))

    def _msg(self, s: str):
        if self._errs is not None:
            self._errs(s, self._udata)
        elif self._udata is None:
            print(s)

    def get_cl_err(self) -> int:
        return self._err

    def cl_error_string(self, err: int) -> str:
        return error_string(err)

    def push_ir(self, ir: np.ndarray) -> int:
        ir = _host(ir, np.float32, "ir")
        stride = ir.shape[-1] if ir.ndim > 1 else ir.size // self.channels
        if stride < self.irsize or ir.size < (self.channels - 1) * stride + self.irsize:
            return INVALID_VALUE
        return self._f[0](self._h, ir.ctypes.data, stride)

    def convolution(self, output: np.ndarray, input1: np.ndarray, input2: np.ndarray | None = None,
                    nblocks: int = 1) -> int:
        out = _host(output, np.float32, "output")
        a = _host(input1, np.float32, "input1")
        need = self.channels * (nblocks if input2 is None else 1) * self.vsize
        if nblocks < 1 or out.size < need or a.size < need or (input2 is not None and np.size(input2) < need):
            self._err = INVALID_VALUE
            return self._err
        if input2 is None:
            self._err = self._f[1](self._h, out.ctypes.data, a.ctypes.data, nblocks)
        else:
            b = _host(input2, np.float32, "input2")
            self._err = self._f[2](self._h, out.ctypes.data, a.ctypes.data, b.ctypes.data)
        if self._err:
            self._msg(error_string(self._err))
        return self._err

    def push_ir_dev(self, d_ir, ir_stride: int, stream=None) -> int:
        _no_multi(self)
        return lib().b2f_dconv_push_ir_dev(self._h, _dptr(d_ir), ir_stride, _stream(stream))

    def convolution_dev(self, d_out, d_in1, d_in2=None, nblocks: int = 1, stream=None) -> int:
        _no_multi(self)
        if d_in2 is None:
            return lib().b2f_dconv_process_dev(self._h, _dptr(d_out), _dptr(d_in1), nblocks, _stream(stream))
        return lib().b2f_dconv_process_tv_dev(self._h, _dptr(d_out), _dptr(d_in1), _dptr(d_in2), _stream(stream))

    def reset(self) -> int:
        return self._f[3](self._h)

    def close(self):
        if C is None:  # interpreter shutdown: the module globals are already gone
            return
        if getattr(self, "_h", None) and self._h.value:
            self._f[4](self._h)
            self._h = C.c_void_p()

    __del__ = close
