"""Simulation driver classes with pyc2ray's interface (reference: pyc2ray/c2ray_base.py, c2ray_test.py),
without astropy / tools21cm: YAML parameters, cell size and R_max, radiation tables, cosmological
bookkeeping, and the calls into the hot path (device_init, photo_table_to_device, evolve3D,
do_raytracing).  This is host glue around the path, so that the reference's
driver scripts run with ``import pyc2ray_b200 as pc2r``.

Differences from the reference: GPU only (use_gpu must be true); ``use_mpi`` is either falsy or truthy --
when truthy the ranks are those of the initialised torch.distributed process group (NCCL), not mpi4py."""
import atexit
import pickle as pkl
import re

import numpy as np
import yaml

from .asora_core import device_init, device_close, photo_table_to_device, cuda_is_init
from .cosmology import FlatLambdaCDM
from .evolve import evolve3D, evolve3D_dist
from .radiation import BlackBodySource, make_tau_table, EV2FR
from .raytracing import do_raytracing
from .utils.logutils import printlog
from .utils.sourceutils import read_test_sources

__all__ = ["C2Ray", "C2Ray_Test", "YEAR", "Mpc"]

# C2Ray's own conversion factors (c2ray_base.py:72-81)
pc = 3.086e18
YEAR = 3.15576E+07
ev2k = 1.0 / 8.617e-05
Mpc = 1e6 * pc


def _yaml_loader():
    """SafeLoader that reads 1e4 as a float (c2ray_base.py:493-505)."""
    class Loader(yaml.SafeLoader):
        pass
    Loader.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                       |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                       |\.[0-9_]+(?:[eE][-+][0-9]+)?
                       |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$""", re.X),
        list("-+0123456789."))
    return Loader


class C2Ray:
    """Base class of a C2Ray simulation (c2ray_base.py:83-511)."""

    def __init__(self, paramfile, Nmesh, use_gpu, use_mpi=None):
        if not use_gpu:
            raise NotImplementedError("CPU ray tracing is not part of this build (use_gpu must be True)")
        self.mpi = bool(use_mpi)
        self.rank, self.nprocs = 0, 1
        if self.mpi:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("use_mpi: initialise torch.distributed (backend='nccl') first")
            self.rank, self.nprocs = dist.get_rank(), dist.get_world_size()
        with open(paramfile, "r") as f:
            self._ld = yaml.load(f, _yaml_loader())
        self.N = Nmesh
        self.shape = (Nmesh, Nmesh, Nmesh)
        self.gpu = True
        device_init(Nmesh, self._ld["Raytracing"]["source_batch_size"])  # c2ray_base.py:113-119
        atexit.register(self._gpu_close)
        self._param_init()
        self._output_init()
        self._grid_init()
        self._cosmology_init()
        self._redshift_init()
        self._material_init()
        self._sources_init()
        self._radiation_init()
        if self.rank == 0:
            q_max = np.ceil(1.73205080757 * min(self.R_max_LLS, 1.73205080757 * self.N / 2))
            self.printlog(f"Using ASORA Raytracing ( q_max = {q_max : n} )")
            self.printlog("Starting simulation... \n\n")

    # ---- time evolution -------------------------------------------------------------------------------
    def set_timestep(self, z1, z2, num_timesteps):
        """c2ray_base.py:147-168"""
        return (self.cosmology.lookback_time(z1) - self.cosmology.lookback_time(z2)) / num_timesteps

    def evolve3D(self, dt, src_flux, src_pos):
        """c2ray_base.py:170-226: evolve the grid over one time step."""
        args = (self.temp, self.ndens, self.xh, self.photo_thin_table, self.photo_thick_table, self.minlogtau,
                self.dlogtau, self.R_max_LLS, self.convergence_fraction, self.sig, self.bh00, self.albpow, self.colh0,
                self.temph0, self.abu_c)
        if self.mpi and src_flux.shape[0] >= self.nprocs:
            self.xh, self.phi_ion = evolve3D_dist(dt, self.dr, src_flux, src_pos, *args, self.logfile)
        else:
            self.xh, self.phi_ion = evolve3D(dt, self.dr, src_flux, src_pos, True, self.max_subbox, self.subboxsize,
                                             self.loss_fraction, *args, self.logfile)

    def cosmo_evolve(self, dt):
        """c2ray_base.py:229-256: advance time; dilute density and rescale dr in cosmological runs."""
        t_half = self.time + 0.5 * dt
        z_half = self.time2zred(t_half)
        if self.cosmological:
            self.ndens *= ((1 + z_half) / (1 + self.zred)) ** 3
            self.dr = self.dr_c * self.cosmology.scale_factor(z_half)
        self.zred = z_half
        self.time = self.time + dt

    def do_raytracing(self, src_flux, src_pos):
        """c2ray_base.py:300-323"""
        gamma, heat = do_raytracing(self.dr, src_flux, src_pos, True, self.max_subbox, self.subboxsize, self.loss_fraction,
                                    self.ndens, self.xh, self.photo_thin_table, self.photo_thick_table,
                                    self.heat_thin_table, self.heat_thick_table, self.minlogtau, self.dlogtau,
                                    self.R_max_LLS, self.sig, self.logfile)
        self.phi_ion = gamma
        self.phi_heat = heat  # None unless Photo.compute_heating_rates (the reference's GPU branch has no heating)
        return gamma

    def printlog(self, s, quiet=False):
        if self.logfile is None:
            raise RuntimeError("Please set the log file in output_ini")
        printlog(s, self.logfile, quiet)

    def write_output(self, z):
        pass

    def time2zred(self, t):
        return self.cosmology.z_at_age(t)

    def zred2time(self, z, unit="s"):
        if unit != "s":
            raise ValueError("only seconds are supported")
        return self.cosmology.age(z)

    # ---- initialisation (c2ray_base.py:329-487) -------------------------------------------------------
    def _param_init(self):
        ld = self._ld
        self.eth0, self.ethe0, self.ethe1 = ld["CGS"]["eth0"], ld["CGS"]["ethe0"], ld["CGS"]["ethe1"]
        self.bh00, self.fh0, self.xih0, self.albpow = ld["CGS"]["bh00"], ld["CGS"]["fh0"], ld["CGS"]["xih0"], ld["CGS"]["albpow"]
        self.abu_h, self.abu_he, self.abu_c = ld["Abundances"]["abu_h"], ld["Abundances"]["abu_he"], ld["Abundances"]["abu_c"]
        self.mean_molecular = self.abu_h + 4.0 * self.abu_he
        self.colh0 = ld["CGS"]["colh0_fact"] * self.fh0 * self.xih0 / self.eth0 ** 2
        self.temph0 = self.eth0 * ev2k
        self.sig = ld["Photo"]["sigma_HI_at_ion_freq"]
        self.loss_fraction = ld["Raytracing"]["loss_fraction"]
        self.convergence_fraction = ld["Raytracing"]["convergence_fraction"]
        self.max_subbox = ld["Raytracing"]["max_subbox"]
        self.subboxsize = ld["Raytracing"]["subboxsize"]

    def _cosmology_init(self):
        c = self._ld["Cosmology"]
        self.cosmology = FlatLambdaCDM(100 * c["h"], c["Omega0"], c["cmbtemp"], Ob0=c["Omega_B"])
        self.cosmological = c["cosmological"]
        self.zred_0 = c["zred_0"]
        self.age_0 = self.zred2time(self.zred_0)
        if self.cosmological:
            self.dr = self.cosmology.scale_factor(self.zred_0) * self.dr_c

    def _radiation_init(self):
        ph = self._ld["Photo"]
        self.minlogtau, self.maxlogtau, self.NumTau = ph["minlogtau"], ph["maxlogtau"], ph["NumTau"]
        self.grey = ph["grey"]
        if ph["SourceType"] != "blackbody":
            raise NameError("Unknown source type : ", ph["SourceType"])
        self.tau, self.dlogtau = make_tau_table(self.minlogtau, self.maxlogtau, self.NumTau)
        f_lo, f_hi = EV2FR * self.eth0, 10 * EV2FR * self.ethe1
        self.bb_Teff = self._ld["BlackBodySource"]["Teff"]
        src = BlackBodySource(self.bb_Teff, self.grey, f_lo, self._ld["BlackBodySource"]["cross_section_pl_index"])
        self.photo_thin_table, self.photo_thick_table = src.make_photo_table(self.tau, f_lo, f_hi, 1e48)
        # c2ray_base.py:384,427-433: heating tables only on request; zeros otherwise (do_raytracing then returns no
        # heating rates)
        self.compute_heating_rates = bool(ph.get("compute_heating_rates", False))
        if self.compute_heating_rates:
            self.heat_thin_table, self.heat_thick_table = src.make_heat_table(self.tau, f_lo, f_hi, 1e48)
        else:
            self.heat_thin_table = np.zeros(self.NumTau + 1)
            self.heat_thick_table = np.zeros(self.NumTau + 1)
        photo_table_to_device(self.photo_thin_table, self.photo_thick_table)  # c2ray_base.py:441-443

    def _grid_init(self):
        self.boxsize_c = self._ld["Grid"]["boxsize"] * Mpc
        self.dr_c = self.boxsize_c / self.N
        self.dr = self.dr_c
        self.R_max_LLS = self._ld["Photo"]["R_max_cMpc"] * self.N / self._ld["Grid"]["boxsize"]
        if self.rank == 0:
            self.printlog(f"Welcome! Mesh size is N = {self.N:n}.")
            self.printlog(f"Maximum comoving distance for photons from source (type 3 LLS): {self.R_max_LLS : .3f} grid cells.")

    def _output_init(self):
        self.logfile = None

    def _redshift_init(self):
        pass

    def _material_init(self):
        pass

    def _sources_init(self):
        pass

    def _gpu_close(self):
        if cuda_is_init():
            device_close()


class C2Ray_Test(C2Ray):
    """Test-case simulation: constant density, sources from a text file (c2ray_test.py:14-181)."""

    def read_sources(self, file, numsrc, S_star_ref=1e48):
        return read_test_sources(file, numsrc, S_star_ref)

    def density_init(self, z):
        self.set_constant_average_density(self.avg_dens, z)

    def set_constant_average_density(self, ndens, z):
        redshift = z if self.cosmological else self.zred_0
        self.ndens = ndens * np.ones(self.shape, order="F") * (1 + redshift) ** 3

    def generate_redshift_array(self, num_zred, delta_t):
        step = delta_t * YEAR
        return np.array([self.time2zred(self.age_0 + i * step) for i in range(num_zred)])

    def write_output(self, z):
        suffix = f"_{z:.3f}.pkl"
        with open(self.results_basename + "xfrac" + suffix, "wb") as f:
            pkl.dump(self.xh, f)
        with open(self.results_basename + "IonRates" + suffix, "wb") as f:
            pkl.dump(self.phi_ion, f)

    def _redshift_init(self):
        self.time = self.age_0
        self.zred = self.zred_0

    def _material_init(self):
        m = self._ld["Material"]
        self.ndens = np.empty(self.shape, order="F")
        self.xh = m["xh0"] * np.ones(self.shape, order="F")
        self.temp = m["temp0"] * np.ones(self.shape, order="F")
        self.phi_ion = np.zeros(self.shape, order="F")
        self.avg_dens = m["avg_dens"]

    def _output_init(self):
        import os
        self.results_basename = self._ld["Output"]["results_basename"]
        os.makedirs(self.results_basename, exist_ok=True)
        self.logfile = self.results_basename + self._ld["Output"]["logfile"]
        if self.rank == 0:
            with open(self.logfile, "w") as f:
                f.write("\nLog file for pyC2Ray (asora-b200) \n\n")
