"""Stand-in for the f2py module ``pyc2ray.lib.libc2ray`` (src/c2ray/Makefile:12-13) as far as the hot
path needs it: ``chemistry.global_pass`` runs on the GPU through the C ABI.  The Fortran CPU ray
tracer (``raytracing.do_all_sources``) is deliberately not provided: this build has no CPU path."""
import numpy as np

from ._cabi import L, check, dptr
import ctypes


class _Chemistry:
    @staticmethod
    def global_pass(dt, ndens, temp, xh, xh_av, xh_intermed, phi_ion, bh00, albpow, colh0, temph0, abu_c):
        """f2py signature of src/c2ray/chemistry.f90:13 -> conv_flag.

        ``xh_av`` and ``xh_intermed`` are intent(inout): float64 arrays updated in place.  All arrays
        must share one shape; they may have different memory orders (evolve.py:200 passes a C-ordered
        ``phi_ion`` next to Fortran-ordered ``xh_av``), in which case the others are re-ordered to
        match ``xh_av`` -- the copy f2py itself would make.
        """
        for nm, a in (("xh_av", xh_av), ("xh_intermed", xh_intermed)):
            if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not (a.flags.f_contiguous or a.flags.c_contiguous):
                raise ValueError(f"failed in converting argument `{nm}' of chemistry.global_pass to C/Fortran array")
        order = "F" if (xh_av.flags.f_contiguous and not xh_av.flags.c_contiguous) else "C"
        if order == "F" and not xh_intermed.flags.f_contiguous or order == "C" and not xh_intermed.flags.c_contiguous:
            raise ValueError("xh_av and xh_intermed must share one memory order")

        def conv(a):
            a = np.asarray(a, dtype=np.float64)
            if a.shape != xh_av.shape:
                raise ValueError("global_pass: all grids must have the same shape")
            return np.asfortranarray(a) if order == "F" else np.ascontiguousarray(a)

        nd, tp, x0, ph = conv(ndens), conv(temp), conv(xh), conv(phi_ion)
        flag = ctypes.c_int(0)
        check(L.asora_global_pass(float(dt), dptr(nd), dptr(tp), dptr(x0), dptr(xh_av), dptr(xh_intermed),
                                  dptr(ph), float(bh00), float(albpow), float(colh0), float(temph0),
                                  float(abu_c), xh_av.size, ctypes.byref(flag)))
        return int(flag.value)


class _Raytracing:
    @staticmethod
    def do_all_sources(*args, **kwargs):
        raise NotImplementedError(
            "libc2ray.raytracing.do_all_sources (Fortran CPU ray tracing, src/c2ray/raytracing.f90:52) is not "
            "part of this build: use use_gpu=True (libasora.do_all_sources).")


chemistry = _Chemistry()
raytracing = _Raytracing()
