"""Minimal flat LambdaCDM background (photons + massless neutrinos), the four calls pyc2ray makes on
astropy's FlatLambdaCDM (pyc2ray/c2ray_base.py:147-165,229-251,280-299,354-373): age, lookback time, scale
factor, and the inverse age(z) -> z.  Host-side bookkeeping only; nothing here is on the hot path."""
import numpy as np
from scipy.integrate import quad
from scipy.optimize import brentq

__all__ = ["FlatLambdaCDM"]

_MPC_CM = 3.0856775814913673e24
_SIGMA_SB = 5.670374419e-5      # erg cm^-2 s^-1 K^-4
_C = 2.99792458e10              # cm/s
_G = 6.67430e-8                 # cm^3 g^-1 s^-2


class FlatLambdaCDM:
    def __init__(self, H0, Om0, Tcmb0=0.0, Ob0=None, Neff=3.04):
        self.H0 = float(H0)                      # km/s/Mpc
        self.Om0 = float(Om0)
        self.Ob0 = Ob0
        self.Tcmb0 = float(Tcmb0)
        self._H0_s = self.H0 * 1.0e5 / _MPC_CM   # 1/s
        rho_crit = 3.0 * self._H0_s ** 2 / (8.0 * np.pi * _G)
        self.Ogamma0 = 4.0 * _SIGMA_SB * self.Tcmb0 ** 4 / (_C ** 3 * rho_crit)
        self.Onu0 = 0.22710731766 * Neff * self.Ogamma0   # 7/8 (4/11)^(4/3) per species
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0

    def efunc(self, z):
        zp1 = 1.0 + z
        return np.sqrt(self.Om0 * zp1 ** 3 + (self.Ogamma0 + self.Onu0) * zp1 ** 4 + self.Ode0)

    def scale_factor(self, z):
        return 1.0 / (1.0 + z)

    def age(self, z):
        """Age of the universe at redshift z in seconds."""
        # t = 1/H0 * int_0^{a} da / (a E(a)); substitute a = 1/(1+z)
        f = lambda a: 1.0 / (a * self.efunc(1.0 / a - 1.0))
        val, _ = quad(f, 0.0, 1.0 / (1.0 + z), epsabs=0.0, epsrel=1e-12, limit=200)
        return val / self._H0_s

    def lookback_time(self, z):
        """Lookback time to redshift z in seconds."""
        f = lambda zz: 1.0 / ((1.0 + zz) * self.efunc(zz))
        val, _ = quad(f, 0.0, z, epsabs=0.0, epsrel=1e-12, limit=200)
        return val / self._H0_s

    def z_at_age(self, t, zmax=1000.0):
        """Redshift at which the universe has age t (seconds)."""
        return brentq(lambda z: self.age(z) - t, 0.0, zmax, xtol=1e-12, rtol=1e-13)
