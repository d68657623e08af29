"""Inputs of the generic inverse PM pass (b200_pm_c2r_readout = petapm_force_c2r with a caller's spectrum), shared by the
golden generator (tests/golden/make_golden_pm_c2r.py, the reference's own petapm.c), the oracle test, the emulation run and
the GPU test.  The functions are MP-GenIC's set (libgenic/zeldovich.c:181-190): Density and the three displacement
components, plus a velocity-like component with a second table, for a power-law DeltaSpec(k)."""
import numpy as np


def hermitian_white_noise(nmesh, seed):
    """rfftn of a real Gaussian field: a half spectrum with exactly the symmetry a c2r transform assumes"""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(np.fft.rfftn(rng.standard_normal((nmesh,) * 3), axes=(0, 1, 2)))


def genic_tables(nmesh, box, index=-1.5, growth=0.7):
    """table[k2] of density_transfer (zeldovich.c:276-289), disp_transfer with DeltaSpec and with dlogGrowth (:291-313) for
    DeltaSpec(k) = k^index, dlogGrowth(k) = growth * DeltaSpec(k)"""
    nk2 = 3 * (nmesh // 2) ** 2 + 1
    k2 = np.arange(nk2, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        kmag = np.sqrt(k2) * 2 * np.pi / box
        delta = kmag ** index
        r2 = (1.0 / nmesh) ** 2
        dens = np.exp(-k2 * r2) * delta / np.sqrt(box * box * box)
        disp = 1. / (2 * np.pi) / np.sqrt(box) / k2 * delta
    dens[0] = 0.0; disp[0] = 0.0
    return dens, disp, growth * disp


def cases():
    """(name, pos, box, nmesh, rho_k, [(kind, table), ...])"""
    out = []
    for name, nmesh, box, npart, seed in (("n24", 24, 40.0, 700, 1), ("n40", 40, 25.0, 900, 2)):
        rng = np.random.default_rng(100 + seed)
        pos = rng.random((npart, 3)) * box
        pos[:4] = [[0.0, 0.0, 0.0], [box * (1 - 1e-12), box / 2, box / nmesh], [box / nmesh * 3, box * (1 - 1e-9), 0.5], [1.0, 2.0, box * (1 - 1e-10)]]
        dens, disp, vel = genic_tables(nmesh, box)
        fn = [(0, dens), (1, disp), (2, disp), (3, disp), (2, vel)]
        out.append((name, pos, box, nmesh, hermitian_white_noise(nmesh, seed), fn))
    return out
