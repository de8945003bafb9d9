"""Optical-depth grid and black-body photo-ionisation tables (host side, once per run).

These are the *inputs* of the hot path: the sweep kernel interpolates the two tables built here.
Layout contract (reference: pyc2ray/radiation/common.py:13-37, pyc2ray/radiation/blackbody.py:22-77):
``tau[0] = 0`` and ``tau[1:] = 10**(minlogtau + arange(NumTau)*dlogtau)``, so a table has
``NumTau + 1`` entries; ``thick[m] = int SED(nu) exp(-tau_m s(nu)) dnu`` and
``thin[m] = int SED(nu) s(nu) exp(-tau_m s(nu)) dnu`` with ``s(nu) = (nu/nu0)**-index`` (1 if grey),
the SED normalised to ``S_star_ref`` photons/s over the integration band.
"""
import numpy as np
from scipy.integrate import quad, quad_vec

__all__ = ["make_tau_table", "BlackBodySource", "blackbody_tables", "blackbody_heat_tables", "EV2FR"]

# C2Ray's own constant values (blackbody.py:11-14, c2ray_base.py:76): kept for comparability.
H_OVER_K = 6.6260755e-27 / 1.381e-16
_PI = 3.141592654
_C = 2.997925e+10
TWO_PI_OVER_C2 = 2.0 * _PI / (_C * _C)
EV2FR = 0.241838e15
# heating tables (blackbody.py:3-5,15-16 take these two from astropy: CODATA 2018 values)
HPLANCK = 6.62607015e-27            # erg s
ION_FREQ_HI = 10973731.568160 * 2.99792458e8   # Ryd * c in Hz


def make_tau_table(minlogtau, maxlogtau, NumTau):
    """Return (tau[NumTau+1], dlogtau).  common.py:13-37."""
    dlogtau = (maxlogtau - minlogtau) / NumTau
    tau = np.concatenate(([0.0], 10 ** (minlogtau + np.arange(NumTau) * dlogtau)))
    return tau, dlogtau


class BlackBodySource:
    """Point source with a black-body spectrum (blackbody.py:22-77)."""

    def __init__(self, temp, grey, freq0, pl_index):
        self.temp = temp
        self.grey = grey
        self.freq0 = freq0
        self.pl_index = pl_index
        self.R_star = 1.0

    def SED(self, freq):
        x = freq * H_OVER_K / self.temp
        if x < 700.0:
            return 4 * np.pi * self.R_star ** 2 * TWO_PI_OVER_C2 * freq ** 2 / (np.exp(x) - 1.0)
        return 0.0

    def integrate_SED(self, f1, f2):
        return quad(self.SED, f1, f2)[0]

    def normalize_SED(self, f1, f2, S_star_ref):
        self.R_star = np.sqrt(S_star_ref / self.integrate_SED(f1, f2)) * self.R_star

    def cross_section_freq_dependence(self, freq):
        return 1.0 if self.grey else (freq / self.freq0) ** (-self.pl_index)

    def _thick(self, freq, tau):
        s = self.cross_section_freq_dependence(freq)
        return np.where(tau * s < 700.0, self.SED(freq) * np.exp(-tau * s), 0.0)

    def _thin(self, freq, tau):
        s = self.cross_section_freq_dependence(freq)
        return np.where(tau * s < 700.0, self.SED(freq) * s * np.exp(-tau * s), 0.0)

    def make_heat_table(self, tau, freq_min, freq_max, S_star_ref):
        """Photo-heating tables: the photo integrands weighted by h (nu - nu_HI) (blackbody.py:55-61,79-85)."""
        self.normalize_SED(freq_min, freq_max, S_star_ref)
        w = lambda f: HPLANCK * (f - ION_FREQ_HI)
        thin = quad_vec(lambda f: w(f) * self._thin(f, tau), freq_min, freq_max, epsrel=1e-12)[0]
        thick = quad_vec(lambda f: w(f) * self._thick(f, tau), freq_min, freq_max, epsrel=1e-12)[0]
        return thin, thick

    def make_photo_table(self, tau, freq_min, freq_max, S_star_ref):
        self.normalize_SED(freq_min, freq_max, S_star_ref)
        thin = quad_vec(lambda f: self._thin(f, tau), freq_min, freq_max, epsrel=1e-12)[0]
        thick = quad_vec(lambda f: self._thick(f, tau), freq_min, freq_max, epsrel=1e-12)[0]
        return thin, thick


def blackbody_tables(Teff, grey, minlogtau, maxlogtau, NumTau, pl_index=2.8, eth0=13.598, ethe1=54.416):
    """Tables exactly as C2Ray._radiation_init builds them (c2ray_base.py:375-417).

    Returns (thin, thick, dlogtau).
    """
    tau, dlogtau = make_tau_table(minlogtau, maxlogtau, NumTau)
    f_lo = EV2FR * eth0
    f_hi = 10 * EV2FR * ethe1
    src = BlackBodySource(Teff, grey, f_lo, pl_index)
    thin, thick = src.make_photo_table(tau, f_lo, f_hi, 1e48)
    return thin, thick, dlogtau


def blackbody_heat_tables(Teff, grey, minlogtau, maxlogtau, NumTau, pl_index=2.8, eth0=13.598, ethe1=54.416):
    """Heating tables as C2Ray._radiation_init builds them when compute_heating_rates is set
    (c2ray_base.py:419-432).  Returns (heat_thin, heat_thick)."""
    tau, _ = make_tau_table(minlogtau, maxlogtau, NumTau)
    f_lo = EV2FR * eth0
    f_hi = 10 * EV2FR * ethe1
    return BlackBodySource(Teff, grey, f_lo, pl_index).make_heat_table(tau, f_lo, f_hi, 1e48)
