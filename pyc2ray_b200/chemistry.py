"""Stand-alone chemistry (reference: pyc2ray/chemistry.py:43-97)."""
import numpy as np

from .load_extensions import load_c2ray

__all__ = ["hydrogenODE", "TEMPH0_ASTROPY"]

# (13.598 eV / k_B) in K with CODATA-2018 e and k_B -- what astropy's (13.598*u.eV/cst.k_B).cgs.value
# evaluates to (chemistry.py:88)
TEMPH0_ASTROPY = 13.598 * 1.602176634e-19 / 1.380649e-23


def hydrogenODE(dt, ndens, temp, xh, phi_ion, bh00=2.59e-13, albpow=-0.7, colh0=1.3e-8, abu_c=7.1e-7):
    """One chemistry pass on the whole grid, hydrogen only; returns the ionised fraction at the end
    of the step.

    The reference passes the same array as xh, xh_av and xh_intermed (chemistry.py:85,91), which is
    undefined behaviour for Fortran dummy arguments; its own documented result
    (tutorials/chemistry_solver.ipynb cell 5: mean 0.050 -> 0.127 after 100 x 50 yr) is the
    end-of-step fraction, which is what is returned here.  Asserts, like the reference, that fewer
    than 1 % of the cells are unconverged.
    """
    libc2ray = load_c2ray()
    xh = np.asfortranarray(xh, dtype=np.float64)
    ndens = np.asfortranarray(ndens, dtype=np.float64)
    temp = np.asfortranarray(temp, dtype=np.float64)
    phi_ion = np.asfortranarray(phi_ion, dtype=np.float64)
    xh_av = np.array(xh, order="F", copy=True)
    xh_intermed = np.array(xh, order="F", copy=True)
    conv_flag = libc2ray.chemistry.global_pass(dt, ndens, temp, xh, xh_av, xh_intermed, phi_ion, bh00, albpow,
                                               colh0, TEMPH0_ASTROPY, abu_c)
    assert conv_flag / np.size(xh_intermed) < 0.01
    return xh_intermed
