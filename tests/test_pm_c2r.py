"""The generic inverse PM pass: b200_pm_c2r_readout = petapm_force_c2r (petapm.c:326-362) with a caller's
source spectrum and table-driven transfer functions, as MP-GenIC's displacement_fields uses it (libgenic/zeldovich.c:150-253).
Golden tests/golden/ref_pm_c2r.npz = the reference's OWN petapm.c driven that way (generator make_golden_pm_c2r.py);
the oracle restatement, the CUDA source under the CPU emulation, and the GPU path are held to it."""
import os
import subprocess
import sys
import numpy as np
import pytest

import oracle
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pm_c2r_scenarios as SC        # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_pm_c2r.npz"))


def test_oracle_c2r_equals_reference_golden():
    for name, pos, box, nmesh, rho_k, fn in SC.cases():
        res = oracle.pm_c2r_readout(pos, box, nmesh, rho_k, fn)
        for j, a in enumerate(res):
            want = GOLD["%s/out%d" % (name, j)]
            assert np.abs(a - want).max() <= 1e-13 * np.abs(want).max(), (name, j)
        # the displacement is the gradient of the density's potential: components differ, none vanishes
        assert all(np.abs(a).max() > 0 for a in res)


@pytest.mark.skipif(not (R.available() and os.path.exists(R.SO_PM)), reason="oracle/_ref/libref_pm.so not built")
def test_oracle_c2r_equals_reference_live():
    r = R.Ref(arena_gib=1.0, nthreads=1, so=R.SO_PM)
    if not hasattr(r.L, "ref_petapm_c2r"):
        pytest.skip("prebuilt libref_pm.so predates ref_petapm_c2r")
    rng = np.random.default_rng(8)
    nmesh, box = 16, 10.0
    pos = rng.random((300, 3)) * box
    dens, disp, vel = SC.genic_tables(nmesh, box, index=-0.8)
    fn = [(0, dens), (3, disp), (1, vel)]
    rk = SC.hermitian_white_noise(nmesh, 5)
    for a, b in zip(oracle.pm_c2r_readout(pos, box, nmesh, rk, fn), r.petapm_c2r(pos, box, nmesh, rk, fn)):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()


def test_c2r_source_under_emulation():
    """k_pm_apply_radial, the inverse passes and k_pm_readout_field of csrc/pm_fft.cu, source unchanged, on the CPU stand-in."""
    env = dict(os.environ, OMP_WAIT_POLICY="passive", B200_FFT_THREADS="64")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "run_fft_emul.py"), "c2r"], env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "c2r ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_gpu_c2r_equals_reference_golden(b200):
    e = b200.Engine(0)
    try:
        for name, pos, box, nmesh, rho_k, fn in SC.cases():
            e.set_particles(pos, np.ones(len(pos), np.float32))
            e.gravpm_init_periodic(box, 1.5, nmesh, 1.0)
            assert e.pm_transform_kind() == 1
            res = e.pm_c2r_readout(rho_k, fn)
            for j, a in enumerate(res):
                want = GOLD["%s/out%d" % (name, j)]
                assert np.abs(a - want).max() <= 1e-12 * np.abs(want).max(), (name, j)
            # the PM force step on the same engine is unaffected by the pass in between
            g0, _ = e.gravpm_force()
            og, _, _ = oracle.pm_force(pos, np.ones(len(pos), np.float32), box, nmesh, 1.5, 1.0)
            assert np.abs(g0 - og).max() <= 1e-9 * np.abs(og).max()
        # a mesh size the shared-memory passes do not take is refused, not silently mis-computed
        e.gravpm_init_periodic(10.0, 1.5, 14, 1.0)
        with pytest.raises(b200.B200Error):
            e.pm_c2r_readout(SC.hermitian_white_noise(14, 1), [(0, np.ones(3 * 7 * 7 + 1))])
    finally:
        e.close()
