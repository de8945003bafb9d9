"""C2Ray_244Test: the driver class of the 244 Mpc/h EoR run (reference: pyc2ray/c2ray_244paper.py:30-387), i.e. the
caller on the input side of the hot path for BASELINE config 4 -- halo catalogues and coarse-grained N-body densities in,
evolve3D per time step, C2Ray binary grids out, resume from them.

The reference does its file I/O through tools21cm (``t2c.DensityFile``, ``t2c.SourceFile``, ``t2c.get_dens_redshifts``,
``t2c.save_cbin`` / ``read_cbin``) and h5py.  tools21cm is a dependency the reference does not pin (``pyproject.toml``:
"tools21cm", no version) and that is absent from this image and from the reference checkout; what is restated here is
its published algorithm as far as the reference's call sites fix it (sambit-giri/tools21cm: ``density_file.py``,
``source_file.py``, ``conv.py``, ``const.py``):

* density files ``<z>n_all.dat``: three int32 mesh sizes, float32 grid masses, Fortran order;
  ``cgs_density = raw * rho_crit_0 * (mesh / nbox_fine)^3 * OmegaB`` -- the baryon density in g/cm^3 (comoving) of a
  coarse cell holding ``raw`` fine-grid mass units; ``nbox_fine`` is the fine N-body mesh of the simulation (8000 for the
  244 Mpc/h box), ``rho_crit_0 = 3 H0^2 / (8 pi G)`` with tools21cm's constants (h = 0.7, G = 6.6732e-8, pc = 3.086e18);
* text catalogues ``<z>-coarsest_wsubgrid_sources.dat``: first line the number of sources, then one row per source
  ``i j k m_hm [m_lm]`` with 1-indexed mesh cells and masses in fine-grid mass units, converted to solar masses with
  ``M_grid = rho_matter * (L_box)^3 / nbox_fine^3`` (``conv.M_grid``) and 5.02785431e-34 M_sun/g;
* ``get_dens_redshifts``: the redshifts of all ``*n_all.dat`` files of a directory.

These factors are class attributes (``NBOX_FINE``, ``T2C_H`` ...) so that a run with another N-body box overrides them.
Everything else -- the EdS-style time / redshift relations, the density dilution, the mass -> photon-rate conversion,
the file names -- is the reference's own code, cited line by line.  Host-side glue: nothing here is on the hot path.
"""
import glob
import os

import numpy as np

from .c2ray_base import C2Ray, YEAR, Mpc
from .utils import c2ray_files as cf

__all__ = ["C2Ray_244Test"]

m_p = cf.M_P            # c2ray_244paper.py:22
msun2g = cf.MSUN2G      # c2ray_base.py:80

_TITLE = ("                 _________   ____            \n    ____  __  __/ ____/__ \\ / __ \\____ ___  __\n"
          "   / __ \\/ / / / /    __/ // /_/ / __ `/ / / /\n  / /_/ / /_/ / /___ / __// _, _/ /_/ / /_/ / \n"
          " / .___/\\__, /\\____//____/_/ |_|\\__,_/\\__, /  \n/_/    /____/                        /____/   \n")


class C2Ray_244Test(C2Ray):
    """c2ray_244paper.py:30-387.  Same constructor, attributes and methods as the reference class."""

    # ---- tools21cm's simulation constants for the 244 Mpc/h CubeP3M run (conv.set_sim_constants(244), const.py) ----
    NBOX_FINE = 8000          # fine N-body mesh per dimension
    T2C_H = 0.7
    T2C_OMEGA0 = 0.27
    T2C_OMEGAB = 0.044
    T2C_G = 6.6732e-8         # cm^3 g^-1 s^-2
    T2C_MPC = 3.086e24        # cm
    SOLAR_MASSES_PER_GRAM = 5.02785431e-34

    def __init__(self, paramfile, Nmesh, use_gpu):
        super().__init__(paramfile, Nmesh, use_gpu)
        self.printlog('Running: "C2Ray for 244 Mpc/h test"')

    # ---- tools21cm conventions (module docstring) -----------------------------------------------------------------
    @classmethod
    def _rho_crit_0(cls):
        H0cgs = 100.0 * cls.T2C_H * 1e5 / cls.T2C_MPC
        return 3.0 * H0cgs * H0cgs / (8.0 * np.pi * cls.T2C_G)

    @classmethod
    def gridmass_to_cgs_density(cls, mesh):
        """Factor of ``t2c.DensityFile.cgs_density`` for a coarse mesh of ``mesh`` cells per dimension."""
        return cls._rho_crit_0() * (float(mesh) / float(cls.NBOX_FINE)) ** 3 * cls.T2C_OMEGAB

    def gridmass_to_msun(self):
        """``conv.M_grid * const.solar_masses_per_gram``: solar masses per fine-grid mass unit."""
        LB = self._ld["Grid"]["boxsize"] / self.T2C_H
        M_box = self._rho_crit_0() * self.T2C_OMEGA0 * (LB * self.T2C_MPC) ** 3
        return M_box / float(self.NBOX_FINE) ** 3 * self.SOLAR_MASSES_PER_GRAM

    @staticmethod
    def get_dens_redshifts(dens_dir):
        """``t2c.get_dens_redshifts``: redshifts of the ``<z>n_all.dat`` files of a directory, ascending."""
        zs = []
        for f in glob.glob(os.path.join(dens_dir, "*n_all.dat")):
            try:
                zs.append(float(os.path.basename(f)[:-len("n_all.dat")]))
            except ValueError:
                pass
        return np.sort(np.array(zs))

    def _density_from_file(self, z_file, redshift):
        """c2ray_244paper.py:269-270,325: coarse-grained density file -> number density of atoms at `redshift`."""
        file = "%scoarser_densities/%.3fn_all.dat" % (self.inputs_basename, z_file)
        raw = cf.read_cbin(file, bits=32, order="F")
        cgs = raw.astype(np.float64) * self.gridmass_to_cgs_density(raw.shape[0])
        return file, np.asfortranarray(cgs / (self.mean_molecular * m_p) * (1 + redshift) ** 3)

    # ---- time evolution (c2ray_244paper.py:51-150) -----------------------------------------------------------------
    def set_timestep(self, z1, z2, num_timesteps):
        """c2ray_244paper.py:51-72"""
        return (self.zred2time(z2) - self.zred2time(z1)) / num_timesteps

    def cosmo_evolve(self, dt):
        """c2ray_244paper.py:74-109: redshift at the half point of the step; density and cell size follow the
        dilution factor."""
        t_now = self.time
        t_half = t_now + 0.5 * dt
        t_after = t_now + dt
        self.printlog(" This is time : %f\t %f" % (t_now / YEAR, t_after / YEAR))
        z_half = self.time2zred(t_half)
        if self.cosmological:
            dilution_factor = (1 + z_half) / (1 + self.zred)
            self.ndens *= dilution_factor ** 3
            self.dr /= dilution_factor
            self.printlog(f"zfactor = {1. / dilution_factor : .10f}")
        self.zred = z_half
        self.time = t_after

    def cosmo_evolve_to_now(self):
        """c2ray_244paper.py:111-131"""
        z_now = self.time2zred(self.time)
        if self.cosmological:
            dilution_factor = (1 + z_now) / (1 + self.zred)
            self.ndens *= dilution_factor ** 3
            self.dr /= dilution_factor
            self.printlog(f"zfactor = {1. / dilution_factor : .10f}")
        self.zred = z_now

    def time2zred(self, t):
        """c2ray_244paper.py:136-143 (matter-dominated relation anchored at zred_0, time = age of the universe)"""
        return -1 + (1. + self.zred_0) * (self.age_0 / t) ** (2. / 3.)

    def zred2time(self, z, unit="s"):
        """c2ray_244paper.py:145-159"""
        if unit != "s":
            raise ValueError("only seconds are supported")
        return self.age_0 * (((1.0 + self.zred_0) / (1.0 + z)) ** 1.5)

    # ---- user methods (c2ray_244paper.py:196-290) ------------------------------------------------------------------
    def read_sources(self, file, mass, ts):
        """c2ray_244paper.py:196-237: (srcpos (3, numsrc) 1-indexed, normflux) from an HDF5 catalogue
        (utils/source_converter.py layout) or an original C2Ray text catalogue (``t2c.SourceFile``)."""
        S_star_ref = 1e48
        mass2phot = msun2g * self.fgamma_hm * self.cosmology.Ob0 / (m_p * ts * self.cosmology.Om0)  # :221
        if file.endswith(".hdf5"):
            srcpos, normflux = cf.read_sources_hdf5(file, self.fgamma_hm, self.cosmology.Ob0, self.cosmology.Om0, ts, S_star_ref)
        else:
            with open(file, "r") as f:
                numsrc = int(f.readline().split()[0])
            rows = np.loadtxt(file, skiprows=1, ndmin=2)[:numsrc]
            col = {"hm": 3, "lm": 4}[mass]
            srcpos = rows[:, :3].T.astype(np.int64)
            normflux = rows[:, col] * self.gridmass_to_msun() * mass2phot / S_star_ref
        self.printlog("\n---- Reading source file with total of %d ionizing source:\n%s" % (normflux.size, file))
        self.printlog(" Total Flux : %e" % np.sum(normflux * S_star_ref))
        self.printlog(" Source lifetime : %f Myr" % (ts / (1e6 * YEAR)))
        return srcpos, normflux

    def read_density(self, z):
        """c2ray_244paper.py:239-275: the density file at or above the current redshift, re-read only when it changes."""
        redshift = z if self.cosmological else self.zred_0
        above = self.zred_density[self.zred_density >= redshift]
        high_z = above[np.argmin(np.abs(above - redshift))]
        if high_z != self.prev_zdens:
            file, self.ndens = self._density_from_file(high_z, redshift)
            self.printlog("\n---- Reading density file:\n " + file)
            self.printlog(" min, mean and max density : %.3e  %.3e  %.3e [1/cm3]" % (self.ndens.min(), self.ndens.mean(), self.ndens.max()))
            self.prev_zdens = high_z

    def write_output(self, z):
        """c2ray_244paper.py:277-290: xfrac_<z>.dat (64 bit) and IonRates_<z>.dat (32 bit), cbin, Fortran order."""
        cf.write_output_cbin(self.results_basename, z, self.xh, self.phi_ion)
        self.printlog("\n--- Reionization History ----")
        self.printlog(" min, mean, max xHII : %.5e  %.5e  %.5e" % (self.xh.min(), self.xh.mean(), self.xh.max()))
        self.printlog(" min, mean, max Irate : %.5e  %.5e  %.5e [1/s]" % (self.phi_ion.min(), self.phi_ion.mean(), self.phi_ion.max()))
        self.printlog(" min, mean, max density : %.5e  %.5e  %.5e [1/cm3]" % (self.ndens.min(), self.ndens.mean(), self.ndens.max()))

    # ---- initialisation (c2ray_244paper.py:165-194, 296-387) ---------------------------------------------------------
    def _cosmology_init(self):
        c = self._ld["Cosmology"]
        from .cosmology import FlatLambdaCDM
        self.cosmology = FlatLambdaCDM(100 * c["h"], c["Omega0"], c["cmbtemp"], Ob0=c["Omega_B"])
        self.cosmological = c["cosmological"]
        self.zred_0 = c["zred_0"]
        # age of an Einstein-de Sitter-like universe at zred_0 (c2ray_244paper.py:180)
        self.age_0 = 2. * (1. + self.zred_0) ** (-1.5) / (3. * 100 * c["h"] * 1e5 / Mpc * np.sqrt(c["Omega0"]))
        if self.cosmological:
            self.printlog(f"Cosmology is on, scaling comoving quantities to the initial redshift, which is z0 = {self.zred_0:.3f}...")
            self.dr = self.dr_c / (1 + self.zred_0)
        else:
            self.printlog("Cosmology is off.")

    def _redshift_init(self):
        """c2ray_244paper.py:296-317"""
        self.zred_density = self.get_dens_redshifts(self.inputs_basename + "coarser_densities/")[::-1]
        self.zred_sources = cf.get_source_redshifts(self.inputs_basename + "sources/")[::-1]
        if self.resume:
            self.zred = np.min(cf.get_redshifts_from_output(self.results_basename))
            _, self.prev_zdens = cf.find_bins(self.zred, self.zred_density)
            _, self.prev_zsourc = cf.find_bins(self.zred, self.zred_sources) if self.zred_sources.size else (None, -1)
        else:
            self.prev_zdens = -1
            self.prev_zsourc = -1
            self.zred = self.zred_0
        self.time = self.zred2time(self.zred)

    def _material_init(self):
        """c2ray_244paper.py:319-341"""
        m = self._ld["Material"]
        if self.resume:
            _, self.ndens = self._density_from_file(self.prev_zdens, self.zred)
            self.xh, phi = cf.read_output_cbin(self.results_basename, self.zred)
            self.xh = np.asfortranarray(self.xh)
            self.temp = m["temp0"] * np.ones(self.shape, order="F")
            self.phi_ion = np.asfortranarray(phi.astype(np.float64))
        else:
            self.ndens = m["avg_dens"] * np.ones(self.shape, order="F")
            self.xh = m["xh0"] * np.ones(self.shape, order="F")
            self.temp = m["temp0"] * np.ones(self.shape, order="F")
            self.phi_ion = np.zeros(self.shape, order="F")

    def _output_init(self):
        """c2ray_244paper.py:343-360"""
        self.results_basename = self._ld["Output"]["results_basename"]
        self.inputs_basename = self._ld["Output"]["inputs_basename"]
        os.makedirs(self.results_basename, exist_ok=True)
        self.logfile = self.results_basename + self._ld["Output"]["logfile"]
        if self._ld["Grid"]["resume"] and os.path.exists(self.logfile):
            with open(self.logfile, "a") as f:
                f.write("\n\nResuming" + _TITLE[8:] + "\n\n")
        else:
            with open(self.logfile, "w") as f:
                f.write(_TITLE + "\nLog file for pyC2Ray.\n\n")

    def _sources_init(self):
        """c2ray_244paper.py:362-368"""
        s = self._ld["Sources"]
        self.fgamma_hm = s["fgamma_hm"]
        self.fgamma_lm = s["fgamma_lm"]
        self.ts = s["ts"] * YEAR * 1e6
        self.printlog(f"Using UV model with fgamma_lm = {self.fgamma_lm:.1f} and fgamma_hm = {self.fgamma_hm:.1f}")

    def _grid_init(self):
        """c2ray_244paper.py:370-387: box size in Mpc/h; R_max in cells."""
        g, h = self._ld["Grid"], self._ld["Cosmology"]["h"]
        self.boxsize_c = g["boxsize"] * Mpc / h
        self.dr_c = self.boxsize_c / self.N
        self.printlog(f"Welcome! Mesh size is N = {self.N:n}.")
        self.printlog(f"Simulation Box size (comoving Mpc): {self.boxsize_c / Mpc:.3e}")
        self.dr = self.dr_c
        self.R_max_LLS = self._ld["Photo"]["R_max_cMpc"] * self.N * h / g["boxsize"]
        self.printlog(f"Maximum comoving distance for photons from source (type 3 LLS): {self._ld['Photo']['R_max_cMpc'] : .3e} comoving Mpc")
        self.printlog(f"This corresponds to {self.R_max_LLS : .3f} grid cells.")
        self.resume = g["resume"]
