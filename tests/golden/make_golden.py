"""Generate the committed golden fixtures.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Nothing here runs on the GPU box; the tests only read the .npz / .json files written next to this
script.  Three kinds of fixture:

1. ``ref_*``   produced by importing the reference's own pure-Python modules from /root/reference
               (radiation tables, source wire format).  astropy is absent here, so the three astropy
               constants those modules import are stubbed with their CODATA values; they only enter
               the heating tables, which are not generated.
2. ``kat.json`` known answers printed in the reference's notebooks (cited per entry).
3. ``oracle_*`` outputs of the CPU oracle (oracle/c2ray_oracle.c) on small seeded inputs, so the GPU
               parity tests also compare against committed vectors, not only against a live oracle run.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _stub_astropy():
    class Q:
        def __init__(self, v):
            self.value = v
            self.cgs = self

        def __mul__(self, o):
            return Q(self.value * (o.value if isinstance(o, Q) else o))

    astropy = types.ModuleType("astropy")
    const = types.ModuleType("astropy.constants")
    const.h = Q(6.62607015e-27)
    const.Ryd = Q(109737.31568160)       # 1/cm
    const.c = Q(2.99792458e10)
    const.k_B = Q(1.380649e-16)
    astropy.constants = const
    sys.modules["astropy"] = astropy
    sys.modules["astropy.constants"] = const


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_tables():
    _stub_astropy()
    common = _load(os.path.join(REF, "pyc2ray/radiation/common.py"), "ref_common")
    bb = _load(os.path.join(REF, "pyc2ray/radiation/blackbody.py"), "ref_blackbody")
    out = {}
    ev2fr, eth0, ethe1 = 0.241838e15, 13.598, 54.416  # c2ray_base.py:76, parameters.yml
    for tag, Teff, grey, NumTau in (("bb1e5", 1e5, 0, 2000), ("bb5e4", 5e4, 0, 2000), ("grey", 5e4, 1, 2000)):
        tau, dlogtau = common.make_tau_table(-20.0, 4.0, NumTau)
        src = bb.BlackBodySource(Teff, grey, ev2fr * eth0, 2.8)
        thin, thick = src.make_photo_table(tau, ev2fr * eth0, 10 * ev2fr * ethe1, 1e48)
        out[f"{tag}_tau"] = tau
        out[f"{tag}_thin"] = thin
        out[f"{tag}_thick"] = thick
        out[f"{tag}_dlogtau"] = np.array(dlogtau)
    np.savez_compressed(os.path.join(HERE, "ref_tables.npz"), **out)
    print("ref_tables.npz", {k: v.shape for k, v in out.items()})


def ref_heat_tables():
    _stub_astropy()
    common = _load(os.path.join(REF, "pyc2ray/radiation/common.py"), "ref_common")
    bb = _load(os.path.join(REF, "pyc2ray/radiation/blackbody.py"), "ref_blackbody")
    ev2fr, eth0, ethe1 = 0.241838e15, 13.598, 54.416
    tau, _ = common.make_tau_table(-20.0, 4.0, 2000)
    src = bb.BlackBodySource(5e4, 0, ev2fr * eth0, 2.8)
    thin, thick = src.make_heat_table(tau, ev2fr * eth0, 10 * ev2fr * ethe1, 1e48)
    np.savez_compressed(os.path.join(HERE, "ref_heat_tables.npz"), bb5e4_heat_thin=thin, bb5e4_heat_thick=thick)
    print("ref_heat_tables.npz", thin.shape, thin[:2], thick[:2])


RUN_FILES = ["xfrac_10.478.dat", "xfrac_9.938.dat", "xfrac_21.062.dat", "IonRates_9.938.dat", "xfrac_notanumber.dat"]
SRC_FILES = ["10.478-coarsest_wsubgrid_sources.dat", "9.938-coarsest_wsubgrid_sources.dat",
             "21.062-coarsest_wsubgrid_sources.dat", "8.515-coarsest_wsubgrid_sources.dat", "readme.txt"]


def ref_other_utils():
    """Redshift bookkeeping of the 244 Mpc run from the reference's own utils/other_utils.py (plain numpy/glob)."""
    import json
    import tempfile
    ou = _load(os.path.join(REF, "pyc2ray/utils/other_utils.py"), "ref_other_utils")
    out = {}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "results"))
        os.makedirs(os.path.join(d, "sources"))
        for f in RUN_FILES:
            open(os.path.join(d, "results", f), "w").close()
        for f in SRC_FILES:
            open(os.path.join(d, "sources", f), "w").close()
        out["from_output"] = [float(z) for z in ou.get_redshifts_from_output(os.path.join(d, "results") + "/")]
        out["source_redshifts"] = [float(z) for z in ou.get_source_redshifts(os.path.join(d, "sources") + "/")]
        out["source_redshifts_9_11"] = [float(z) for z in ou.get_source_redshifts(os.path.join(d, "sources") + "/", 9.0, 11.0)]
        out["source_redshifts_bracket"] = [float(z) for z in
                                           ou.get_source_redshifts(os.path.join(d, "sources") + "/", 9.0, 11.0, True)]
    edges = [8.515, 9.938, 10.478, 21.062]
    out["find_bins"] = {str(v): [None if b is None else float(b) for b in ou.find_bins(v, np.array(edges))]
                        for v in (7.0, 9.0, 10.0, 15.0)}
    out["run_files"], out["src_files"], out["edges"] = RUN_FILES, SRC_FILES, edges
    with open(os.path.join(HERE, "ref_other_utils.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("ref_other_utils.json", out)


def ref_sources():
    su = _load(os.path.join(REF, "pyc2ray/utils/sourceutils.py"), "ref_sourceutils")
    tmp = "/tmp/_golden_src.txt"
    su.generate_test_sourcefile(tmp, 250, 64, 5e48, seed=100)
    pos, flux = su.read_test_sources(tmp, 64)
    pos_flat, flux_flat = su.format_sources(pos, flux)
    np.savez_compressed(os.path.join(HERE, "ref_sources.npz"), pos=pos, flux=flux, pos_flat=pos_flat,
                        flux_flat=flux_flat, file_text=np.array(open(tmp).read()))
    print("ref_sources.npz", pos.shape, pos_flat[:6], flux_flat[:2])


def kat():
    k = {
        "chemistry_tutorial": {
            "source": "tutorials/chemistry_solver.ipynb cells 3,5 (printed output)",
            "seed": 2023, "shape": [10, 10, 10], "dt_yr": 50, "steps": 100,
            "mean_x_start_3dp": "0.050", "mean_x_end_3dp": "0.127"},
        "test3_multisource_mean_x": {
            "source": "test/paper_tests/test3_multisource/make_plot.ipynb cell 5 (printed output)",
            "order": ["grey", "Teff5e3", "Teff5e4", "Teff1e5"],
            "c2ray": [0.09488065, 0.09503048, 0.09583101, 0.09492813],
            "pyc2ray": [0.09488056, 0.0950304, 0.09583087, 0.09492792]},
        "test1_stromgren": {
            "source": "test/paper_tests/test1_Ifront/make_plot.ipynb cells 5,10",
            "r_S_kpc": 964.377, "t_rec_Myr": 654.27, "band": [0.985, 1.005]},
        "hackathon_test1_tolerances": {
            "source": "test/unit_tests_hackathon/1_single_black_body/run_test.py:91-115",
            "abs": {"mean": 1e-8, "std": 3e-7, "max": 5e-6}, "rel": {"mean": 1e-7, "std": 3e-6, "max": 2e-5}},
        "test2_cosmo_Ifront": {
            "source": "test/paper_tests/test2_Ifront_cosmo/make_plot.ipynb cell 5 (printed output) and cell 9 (plot band)",
            "age_z9_Myr": 563.9825828256307, "lambda": 0.862007470892602,
            "cosmology": {"H0": 70, "Om0": 0.27, "Tcmb0": 2.726, "Ob0": 0.043}, "band": [0.985, 1.005]},
        "asora_asymptote_ns_per_source_cell": {
            "source": "test/paper_tests/raytracing_benchmark/plot_sources.ipynb cell 5", "value": 3.1558841065316494e-09},
    }
    json.dump(k, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    print("kat.json")


def oracle_vectors():
    import oracle
    from tests.fields import make_case
    out = {}
    for name in ("small_r5", "clip_full_n24", "multi_n32"):
        c = make_case(name)
        phi, cdh, n = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(),
                                                  c["pos_flat"], c["flux_flat"], c["N"], c["thin"], c["thick"],
                                                  c["minlogtau"], c["dlogtau"], c["NumTau"])
        out[name + "_phi"] = phi
        out[name + "_n"] = np.array(n)
        if c["flux_flat"].size == 1:
            out[name + "_cdh"] = cdh
    np.savez_compressed(os.path.join(HERE, "oracle_sweep.npz"), **out)
    print("oracle_sweep.npz", list(out))


if __name__ == "__main__":
    if "heat" in sys.argv[1:]:
        ref_heat_tables()
        sys.exit(0)
    if "files" in sys.argv[1:]:
        ref_other_utils()
        sys.exit(0)
    ref_tables()
    ref_heat_tables()
    ref_other_utils()
    ref_sources()
    kat()
    oracle_vectors()
