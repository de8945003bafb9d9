"""Drop-in for the reference's CPython extension ``pyc2ray.lib.libasora``
(src/asora/python_module.cu:153-161): the same six functions with the same positional signatures,
implemented over the C ABI of libasora_b200.so.

Like the reference wrappers, array arguments are taken as raw float64 / int32 buffers; unlike them,
dtype, contiguity and size are checked, because a mismatch here would silently ray-trace garbage
(python_module.cu:60-63 reads PyArray_DATA unchecked).
"""
import numpy as np

from . import _cabi
from ._cabi import L, check, dptr, iptr

__all__ = ["do_all_sources", "device_init", "device_close", "density_to_device", "photo_table_to_device",
           "source_data_to_device", "heat_table_to_device", "do_all_sources_heat"]

_N = None


def _f64(a, name, size=None):
    if not isinstance(a, np.ndarray) or a.dtype != np.float64:
        raise TypeError(f"{name} must be Array of type double")
    if not (a.flags.c_contiguous or a.flags.f_contiguous):
        raise TypeError(f"{name} must be contiguous")
    if size is not None and a.size < size:
        raise ValueError(f"{name} has {a.size} elements, expected at least {size}")
    return a


def device_init(N, num_src_par):
    """python_module.cu:73-82"""
    global _N
    check(L.asora_device_init(int(N), int(num_src_par)))
    _N = int(N)


def device_close():
    """python_module.cu:87-92"""
    global _N
    check(L.asora_device_close())
    _N = None


def density_to_device(ndens, N):
    """python_module.cu:97-109"""
    _f64(ndens, "ndens", int(N) ** 3)
    check(L.asora_density_to_device(dptr(ndens), int(N)))


def photo_table_to_device(thin_table, thick_table, NumTau):
    """python_module.cu:114-128"""
    _f64(thin_table, "thin_table", int(NumTau))
    _f64(thick_table, "thick_table", int(NumTau))
    check(L.asora_photo_table_to_device(dptr(thin_table), dptr(thick_table), int(NumTau)))


def source_data_to_device(pos, flux, NumSrc):
    """python_module.cu:133-148"""
    if not isinstance(pos, np.ndarray) or pos.dtype != np.int32 or not pos.flags.c_contiguous:
        raise TypeError("pos must be a contiguous Array of type int32")
    if pos.size < 3 * int(NumSrc):
        raise ValueError("pos is shorter than 3*NumSrc")
    _f64(flux, "flux", int(NumSrc))
    check(L.asora_source_data_to_device(iptr(pos), dptr(flux), int(NumSrc)))


def do_all_sources(R, coldensh_out, sig, dr, ndens, xh_av, phi_ion, NumSrc, m1, minlogtau, dlogtau, NumTau, *, group=None,
                   download=True, xh_from=None):
    """python_module.cu:21-68.  ``coldensh_out`` and ``ndens`` are accepted and ignored exactly as the
    reference ignores them (raytracing.cu:116); ``phi_ion`` is overwritten in place.

    ``group`` (keyword, not in the reference): a torch.distributed process group (or True for the default group) whose
    ranks each hold a shard of the sources on their own GPU.  The rates of all ranks are then summed on the devices by
    one NCCL all-reduce between the sweep and the download -- the reference's Reduce + Bcast of host arrays
    (pyc2ray/evolve.py:433-437) -- and ``download=False`` lets a rank skip the device-to-host copy when it does not
    need the grid on the host.  ``xh_from=r``: the ionised fractions are the same on every rank (they are in a
    source-sharded run), so only rank ``r`` copies its ``xh_av`` to its GPU and the others receive it from that GPU over
    NVLink (one NCCL broadcast of 8 N^3 bytes) instead of N uploads through the host's memory system; the ``xh_av``
    argument of the other ranks is not read."""
    if not isinstance(coldensh_out, np.ndarray) or coldensh_out.dtype != np.float64:
        raise TypeError("coldensh_out must be Array of type double")  # python_module.cu:53-57
    n3 = int(m1) ** 3
    if group is None or xh_from is None:
        _f64(xh_av, "xh_av", n3)
    _f64(phi_ion, "phi_ion", n3)
    if not phi_ion.flags.writeable:
        raise ValueError("phi_ion must be writeable")
    if group is None:
        check(L.asora_do_all_sources(float(R), float(sig), float(dr), dptr(xh_av), dptr(phi_ion), int(NumSrc),
                                     int(m1), float(minlogtau), float(dlogtau), int(NumTau)))
        return
    import torch
    import torch.distributed as dist
    from ..parallel import device_tensor
    grp = None if group is True else group
    if xh_from is None:
        xh_ptr = dptr(xh_av)
    else:
        if dist.get_rank(grp) == xh_from:
            check(L.asora_buffer_upload(_cabi.BUF_XH_AV, dptr(_f64(xh_av, "xh_av", n3))))
        xav_t = device_tensor(L.asora_device_buffer(_cabi.BUF_XH_AV), n3)
        dist.broadcast(xav_t, src=dist.get_global_rank(grp, xh_from) if grp is not None else xh_from, group=grp)
        xh_ptr = None
    check(L.asora_do_all_sources_begin(float(R), float(sig), float(dr), xh_ptr, int(NumSrc), int(m1), float(minlogtau),
                                       float(dlogtau), int(NumTau)))
    phi_t = device_tensor(L.asora_device_buffer(_cabi.BUF_PHI_ION), n3)
    torch.cuda.nvtx.range_push("asora:allreduce_phi")
    dist.all_reduce(phi_t, op=dist.ReduceOp.SUM, group=grp)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    check(L.asora_do_all_sources_end(dptr(phi_ion) if download else None))


# ---- photo-heating rates: not in the reference's libasora (TODO at c2ray_base.py:424-426); the signatures extend
# ---- photo_table_to_device / do_all_sources the way the CPU ray tracer takes its heating arguments -------------

def heat_table_to_device(heat_thin_table, heat_thick_table, NumTau):
    """Heating tables (radiation/blackbody.py:79-85) next to the photo tables already on the device."""
    _f64(heat_thin_table, "heat_thin_table", int(NumTau))
    _f64(heat_thick_table, "heat_thick_table", int(NumTau))
    check(L.asora_heat_table_to_device(dptr(heat_thin_table), dptr(heat_thick_table), int(NumTau)))


def do_all_sources_heat(R, coldensh_out, sig, dr, ndens, xh_av, phi_ion, phi_heat, NumSrc, m1, minlogtau, dlogtau,
                        NumTau):
    """do_all_sources with ``phi_heat`` (overwritten in place) after ``phi_ion``, as
    libc2ray.raytracing.do_all_sources orders them (raytracing.f90:52-56)."""
    if not isinstance(coldensh_out, np.ndarray) or coldensh_out.dtype != np.float64:
        raise TypeError("coldensh_out must be Array of type double")
    n3 = int(m1) ** 3
    _f64(xh_av, "xh_av", n3)
    _f64(phi_ion, "phi_ion", n3)
    _f64(phi_heat, "phi_heat", n3)
    if not (phi_ion.flags.writeable and phi_heat.flags.writeable):
        raise ValueError("phi_ion and phi_heat must be writeable")
    check(L.asora_do_all_sources_heat(float(R), float(sig), float(dr), dptr(xh_av), dptr(phi_ion), dptr(phi_heat),
                                      int(NumSrc), int(m1), float(minlogtau), float(dlogtau), int(NumTau)))
