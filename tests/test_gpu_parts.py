"""GPU parity of the split sweeps (parts = 2, 4, 8: half-spaces, quadrants, octants of one source swept by
separate CTAs, bounding planes recomputed but rated once) against the unsplit sweep and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small_r5", "clip_full_n24", "odd_n15_full", "r_int5", "multi_n32", "mid_n48_r14"])
@pytest.mark.parametrize("parts", [2, 4, 8])
@pytest.mark.parametrize("sphere_only", [0, 1])
def test_split_sweep_matches_oracle(name, parts, sphere_only):
    import oracle
    from pyc2ray_b200.lib import _cabi, libasora
    from tests.fields import make_case
    from tests.test_gpu_parity import _setup, _sweep, _assert_close
    c = make_case(name)
    _setup(libasora, c)
    try:
        _cabi.check(_cabi.L.asora_set_tuning(0, parts << 20))
        _cabi.check(_cabi.L.asora_set_sphere_only(sphere_only))
        phi, used, upd = _sweep(libasora, _cabi, c, 1)
    finally:
        _cabi.L.asora_set_tuning(0, 0)
        _cabi.L.asora_set_sphere_only(0)
        libasora.device_close()
    ref, _, n = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                            c["flux_flat"], c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"],
                                            c["NumTau"])
    assert used == 1
    if not sphere_only:
        assert upd == n
    _assert_close(phi, ref, f"{name} parts={parts} sphere_only={sphere_only}")
