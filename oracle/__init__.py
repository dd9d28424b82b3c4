"""oracle -- ctypes front-end to the parity oracle. TEST INFRASTRUCTURE ONLY.

Two back-ends with the same Python surface:

* ``port``  -- ``oracle/_build/liboracle.so``: the plain-C restatement ``ref_cpu.c`` of the
  reference's kernels + host sequencing (always buildable: ``make -C oracle port``).
* ``ref``   -- ``oracle/_ref/libclfft_ref.so``: the UNMODIFIED reference C++ sources from
  /root/reference compiled against ``oracle/minicl`` (a host-CPU OpenCL runtime) and driven through
  ``ref_capi.cpp``. Built only where /root/reference exists; the built library travels to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package. The product (``opencl_fft_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_HERE, "_build", "liboracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libclfft_ref.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    """(Re)build the oracle libraries with oracle/Makefile (port always; ref when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _as_c64_floats(x) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.complex64)
    return a.view(np.float32).copy()


class _Port:
    """The C restatement (oracle/ref_cpu.c)."""

    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_LIB):
            build()
        L = self.L = C.CDLL(PORT_LIB)
        L.orc_cfft.argtypes = [_f32p, C.c_int, C.c_int]
        L.orc_rfft.argtypes = [_f32p, C.c_int, C.c_int]
        L.orc_pconv_create.restype = C.c_void_p
        L.orc_pconv_create.argtypes = [C.c_int, C.c_int]
        L.orc_pconv_destroy.argtypes = [C.c_void_p]
        L.orc_pconv_nparts.argtypes = [C.c_void_p]
        L.orc_pconv_push_ir.argtypes = [C.c_void_p, _f32p]
        L.orc_pconv_convolution.argtypes = [C.c_void_p, _f32p, _f32p]
        L.orc_pconv_convolution_tv.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        for n in ("orc_pconv_spec1", "orc_pconv_spec2", "orc_pconv_olap"):
            getattr(L, n).restype = C.POINTER(C.c_float)
            getattr(L, n).argtypes = [C.c_void_p]
        L.orc_dconv_create.restype = C.c_void_p
        L.orc_dconv_create.argtypes = [C.c_int, C.c_int]
        L.orc_dconv_destroy.argtypes = [C.c_void_p]
        L.orc_dconv_push_ir.argtypes = [C.c_void_p, _f32p]
        L.orc_dconv_convolution.argtypes = [C.c_void_p, _f32p, _f32p]
        L.orc_dconv_convolution_tv.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        L.orc_dft64.argtypes = [_f64p, _f64p, C.c_int, C.c_int]
        L.orc_pconv_run.restype = C.c_double
        L.orc_pconv_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_int]
        L.orc_dconv_run.restype = C.c_double
        L.orc_dconv_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_int]
        L.orc_rfft_run.restype = C.c_double
        L.orc_rfft_run.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, C.c_int]
        L.orc_cfft_run.restype = C.c_double
        L.orc_cfft_run.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, C.c_int]

    # -- FFTs -------------------------------------------------------------------------------
    def cfft(self, x, fwd: bool = True) -> np.ndarray:
        """Clcfft::transform on one N-point complex64 vector."""
        buf = _as_c64_floats(x)
        self.L.orc_cfft(buf, buf.size // 2, int(fwd))
        return buf.view(np.complex64)

    def rfft_fwd(self, x) -> np.ndarray:
        """Clrfft forward: `size` float32 -> size/2 packed complex64."""
        buf = np.ascontiguousarray(x, dtype=np.float32).copy()
        self.L.orc_rfft(buf, buf.size, 1)
        return buf.view(np.complex64)

    def rfft_inv(self, s) -> np.ndarray:
        """Clrfft inverse: size/2 packed complex64 -> `size` float32."""
        buf = _as_c64_floats(s)
        self.L.orc_rfft(buf, buf.size, 0)
        return buf

    def dft64(self, x, sign: int = -1) -> np.ndarray:
        a = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty(a.size * 2, dtype=np.float64)
        self.L.orc_dft64(a.view(np.float64).copy(), out, a.size, sign)
        return out.view(np.complex128)

    # -- convolvers -------------------------------------------------------------------------
    def pconv(self, cvs: int, pts: int) -> "PConv":
        return PConv(self, cvs, pts)

    def dconv(self, irsize: int, vsize: int) -> "DConv":
        return DConv(self, irsize, vsize)

    # -- timed loops (CPU baseline) -----------------------------------------------------------
    def pconv_run(self, cvs, pts, ir, x, threads):
        ir = np.ascontiguousarray(ir, np.float32)
        x = np.ascontiguousarray(x, np.float32)  # [channels, nblocks*pts]
        ch = x.shape[0]
        nb = x.shape[1] // pts
        out = np.zeros_like(x)
        secs = self.L.orc_pconv_run(ch, cvs, pts, nb, ir.reshape(-1), x.reshape(-1), out.reshape(-1), threads)
        return secs, out

    def dconv_run(self, irsize, vsize, ir, x, threads):
        ir = np.ascontiguousarray(ir, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        ch = x.shape[0]
        nb = x.shape[1] // vsize
        out = np.zeros_like(x)
        secs = self.L.orc_dconv_run(ch, irsize, vsize, nb, ir.reshape(-1), x.reshape(-1), out.reshape(-1), threads)
        return secs, out

    def rfft_run(self, x, fwd, threads):
        buf = np.ascontiguousarray(x, np.float32).copy()  # [batch, size]
        secs = self.L.orc_rfft_run(buf.shape[1], buf.shape[0], buf.reshape(-1), int(fwd), threads)
        return secs, buf

    def cfft_run(self, x, fwd, threads):
        buf = np.ascontiguousarray(x, np.complex64).copy()  # [batch, N]
        f = buf.view(np.float32)
        secs = self.L.orc_cfft_run(buf.shape[1], buf.shape[0], f.reshape(-1), int(fwd), threads)
        return secs, buf


class PConv:
    """Clpconv semantics on the C restatement."""

    def __init__(self, port: _Port, cvs: int, pts: int):
        self._L = port.L
        self.pts = pts
        self._h = self._L.orc_pconv_create(cvs, pts)
        self.nparts = self._L.orc_pconv_nparts(self._h)

    def push_ir(self, ir):
        ir = np.ascontiguousarray(ir, np.float32)
        assert ir.size >= self.nparts * self.pts
        self._L.orc_pconv_push_ir(self._h, ir)

    def convolution(self, x, x2=None) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(self.pts, np.float32)
        if x2 is None:
            self._L.orc_pconv_convolution(self._h, out, x)
        else:
            self._L.orc_pconv_convolution_tv(self._h, out, x, np.ascontiguousarray(x2, np.float32))
        return out

    def spec1(self) -> np.ndarray:
        n = self.nparts * self.pts * 2
        return np.ctypeslib.as_array(self._L.orc_pconv_spec1(self._h), (n,)).copy().view(np.complex64)

    def spec2(self) -> np.ndarray:
        n = self.nparts * self.pts * 2
        return np.ctypeslib.as_array(self._L.orc_pconv_spec2(self._h), (n,)).copy().view(np.complex64)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_pconv_destroy(self._h)
            self._h = None


class DConv:
    """Cldconv semantics on the C restatement."""

    def __init__(self, port: _Port, irsize: int, vsize: int):
        self._L = port.L
        self.irsize, self.vsize = irsize, vsize
        self._h = self._L.orc_dconv_create(irsize, vsize)

    def push_ir(self, ir):
        self._L.orc_dconv_push_ir(self._h, np.ascontiguousarray(ir, np.float32))

    def convolution(self, x, x2=None) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(self.vsize, np.float32)
        if x2 is None:
            self._L.orc_dconv_convolution(self._h, out, x)
        else:
            self._L.orc_dconv_convolution_tv(self._h, out, x, np.ascontiguousarray(x2, np.float32))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_dconv_destroy(self._h)
            self._h = None


class _Ref:
    """The unmodified reference classes (via oracle/minicl); same surface as _Port."""

    kind = "reference"

    def __init__(self):
        L = self.L = C.CDLL(REF_LIB)
        L.ref_device_name.restype = C.c_char_p
        for name in ("ref_cfft_create", "ref_rfft_create", "ref_pconv_create", "ref_dconv_create"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_int, C.c_int]
        for name in ("ref_cfft_destroy", "ref_rfft_destroy", "ref_pconv_destroy", "ref_dconv_destroy",
                     "ref_cfft_error", "ref_rfft_error", "ref_pconv_error", "ref_dconv_error"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_cfft_transform.argtypes = [C.c_void_p, _f32p]
        L.ref_rfft_transform.argtypes = [C.c_void_p, _f32p]
        L.ref_rfft_transform2.argtypes = [C.c_void_p, _f32p, _f32p]
        L.ref_pconv_push_ir.argtypes = [C.c_void_p, _f32p]
        L.ref_pconv_convolution.argtypes = [C.c_void_p, _f32p, _f32p]
        L.ref_pconv_convolution_tv.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        L.ref_dconv_push_ir.argtypes = [C.c_void_p, _f32p]
        L.ref_dconv_convolution.argtypes = [C.c_void_p, _f32p, _f32p]
        L.ref_dconv_convolution_tv.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        L.ref_pconv_run.restype = C.c_double
        L.ref_pconv_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_int]
        L.ref_dconv_run.restype = C.c_double
        L.ref_dconv_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_int]
        L.ref_rfft_run.restype = C.c_double
        L.ref_rfft_run.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, C.c_int]
        L.ref_cfft_run.restype = C.c_double
        L.ref_cfft_run.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, C.c_int]

    def device_name(self) -> str:
        return self.L.ref_device_name().decode()

    def cfft(self, x, fwd: bool = True) -> np.ndarray:
        buf = _as_c64_floats(x)
        h = self.L.ref_cfft_create(buf.size // 2, int(fwd))
        assert self.L.ref_cfft_error(h) == 0
        self.L.ref_cfft_transform(h, buf)
        self.L.ref_cfft_destroy(h)
        return buf.view(np.complex64)

    def rfft_fwd(self, x, out_of_place: bool = False) -> np.ndarray:
        r = np.ascontiguousarray(x, dtype=np.float32).copy()
        h = self.L.ref_rfft_create(r.size, 1)
        assert self.L.ref_rfft_error(h) == 0
        if out_of_place:
            c = np.zeros(r.size, np.float32)
            self.L.ref_rfft_transform2(h, c, r)
        else:
            c = r
            self.L.ref_rfft_transform(h, c)
        self.L.ref_rfft_destroy(h)
        return c.view(np.complex64)

    def rfft_inv(self, s, out_of_place: bool = False) -> np.ndarray:
        c = _as_c64_floats(s)
        h = self.L.ref_rfft_create(c.size, 0)
        assert self.L.ref_rfft_error(h) == 0
        if out_of_place:
            r = np.zeros(c.size, np.float32)
            self.L.ref_rfft_transform2(h, c, r)
        else:
            r = c
            self.L.ref_rfft_transform(h, c)
        self.L.ref_rfft_destroy(h)
        return r

    def pconv(self, cvs: int, pts: int) -> "RefPConv":
        return RefPConv(self, cvs, pts)

    def dconv(self, irsize: int, vsize: int) -> "RefDConv":
        return RefDConv(self, irsize, vsize)

    def pconv_run(self, cvs, pts, ir, x, threads):
        ir = np.ascontiguousarray(ir, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        ch = x.shape[0]
        nb = x.shape[1] // pts
        out = np.zeros_like(x)
        secs = self.L.ref_pconv_run(ch, cvs, pts, nb, ir.reshape(-1), x.reshape(-1), out.reshape(-1), threads)
        return secs, out

    def dconv_run(self, irsize, vsize, ir, x, threads):
        ir = np.ascontiguousarray(ir, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        ch = x.shape[0]
        nb = x.shape[1] // vsize
        out = np.zeros_like(x)
        secs = self.L.ref_dconv_run(ch, irsize, vsize, nb, ir.reshape(-1), x.reshape(-1), out.reshape(-1), threads)
        return secs, out

    def rfft_run(self, x, fwd, threads):
        buf = np.ascontiguousarray(x, np.float32).copy()
        secs = self.L.ref_rfft_run(buf.shape[1], buf.shape[0], buf.reshape(-1), int(fwd), threads)
        return secs, buf

    def cfft_run(self, x, fwd, threads):
        buf = np.ascontiguousarray(x, np.complex64).copy()
        f = buf.view(np.float32)
        secs = self.L.ref_cfft_run(buf.shape[1], buf.shape[0], f.reshape(-1), int(fwd), threads)
        return secs, buf


class RefPConv:
    def __init__(self, ref: _Ref, cvs: int, pts: int):
        self._L = ref.L
        self.pts = pts
        self.nparts = cvs // pts
        self._h = self._L.ref_pconv_create(cvs, pts)
        assert self._L.ref_pconv_error(self._h) == 0

    def push_ir(self, ir):
        ir = np.ascontiguousarray(ir, np.float32).copy()
        assert ir.size >= self.nparts * self.pts
        assert self._L.ref_pconv_push_ir(self._h, ir) == 0

    def convolution(self, x, x2=None) -> np.ndarray:
        # the reference's 3-arg form documents 2*pts-float host arrays (cl_conv.h:178-181); give it that
        xin = np.zeros(2 * self.pts, np.float32)
        xin[: self.pts] = x
        out = np.zeros(2 * self.pts, np.float32)
        if x2 is None:
            assert self._L.ref_pconv_convolution(self._h, out, xin) == 0
        else:
            x2in = np.zeros(2 * self.pts, np.float32)
            x2in[: self.pts] = x2
            assert self._L.ref_pconv_convolution_tv(self._h, out, xin, x2in) == 0
        return out[: self.pts].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_pconv_destroy(self._h)
            self._h = None


class RefDConv:
    def __init__(self, ref: _Ref, irsize: int, vsize: int):
        self._L = ref.L
        self.irsize, self.vsize = irsize, vsize
        self._h = self._L.ref_dconv_create(irsize, vsize)
        assert self._L.ref_dconv_error(self._h) == 0

    def push_ir(self, ir):
        assert self._L.ref_dconv_push_ir(self._h, np.ascontiguousarray(ir, np.float32).copy()) == 0

    def convolution(self, x, x2=None) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32).copy()
        out = np.zeros(self.vsize, np.float32)
        if x2 is None:
            assert self._L.ref_dconv_convolution(self._h, out, x) == 0
        else:
            assert self._L.ref_dconv_convolution_tv(self._h, out, x, np.ascontiguousarray(x2, np.float32).copy()) == 0
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_dconv_destroy(self._h)
            self._h = None


_port = None
_ref = None


def port() -> _Port:
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref():
    """The compiled reference, or None when oracle/_ref/libclfft_ref.so is absent."""
    global _ref
    if _ref is None and os.path.exists(REF_LIB):
        _ref = _Ref()
    return _ref


def best():
    """The strongest CPU implementation available: the real reference if built, else the port."""
    return ref() or port()
