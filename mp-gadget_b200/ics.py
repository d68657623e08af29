"""Synthetic initial conditions for tests and bench.

Seeded, self-contained (no reference data files): a cubic lattice displaced by
a Zel'dovich-like Gaussian displacement field with P(k) ~ k^-2, rms 3-D
displacement `rms` lattice spacings ("z9" state), optionally with the
reference's clustered recipe (tests/test_gravity.c:288-302) mixed in.
Also the reference's unit-test particle distributions, restated.
"""
import numpy as np

G_INTERNAL = 43.0071          # tests/test_gravity.c:35
OMEGA0 = 0.288
HUBBLE = 100.0                # km/s/(Mpc/h) with Mpc/h length units


def rho_crit(G=G_INTERNAL):
    return 3 * HUBBLE ** 2 / (8 * np.pi * G)


def default_nmesh(ncbrt):
    """run.c:211-212: Nmesh = 3 * 2^floor(log2(cbrt(N_dm)))."""
    return 3 * 2 ** int(np.floor(np.log2(ncbrt) + 1e-9))


def zeldovich_lattice(ng, box, seed=181170, rms=0.2):
    """ng^3 particles; returns pos[n,3] (f64, in [0,box)), mass[n] (f32)."""
    rng = np.random.default_rng(seed)
    spacing = box / ng
    noise = rng.standard_normal((ng, ng, ng)).astype(np.float32)
    nk = np.fft.rfftn(noise)
    del noise
    kx = np.fft.fftfreq(ng) * ng
    kz = np.fft.rfftfreq(ng) * ng
    k2 = (kx[:, None, None] ** 2 + kx[None, :, None] ** 2 + kz[None, None, :] ** 2).astype(np.float32)
    k2[0, 0, 0] = 1.0
    # delta_k ~ noise / k ; psi_k = i k delta_k / k^2  -> noise * i k / k^3
    amp = nk / (k2 ** 1.5)
    amp[0, 0, 0] = 0
    amp[k2 > (ng / 2) ** 2] = 0
    del nk
    disp = []
    for kk in (kx[:, None, None], kx[None, :, None], kz[None, None, :]):
        disp.append(np.fft.irfftn(1j * kk * amp, s=(ng, ng, ng), axes=(0, 1, 2)).astype(np.float64))
    del amp
    var = sum((x ** 2).mean() for x in disp)
    scale = rms * spacing / np.sqrt(var)
    g = (np.arange(ng) + 0.5) * spacing
    pos = np.empty((ng ** 3, 3), dtype=np.float64)
    pos[:, 0] = (g[:, None, None] + scale * disp[0]).ravel()
    pos[:, 1] = (g[None, :, None] + scale * disp[1]).ravel()
    pos[:, 2] = (g[None, None, :] + scale * disp[2]).ravel()
    np.mod(pos, box, out=pos)
    pos[pos >= box] = 0.0
    mass = np.full(ng ** 3, OMEGA0 * rho_crit() * spacing ** 3, dtype=np.float32)
    return pos, mass


def clustered_mix(n, box, seed=0):
    """tests/test_gravity.c:288-302 (do_random_test): 1/4 uniform, 1/2 in a
    clump at box/2, 1/4 in a clump at 0.1 box.  With seed=0 the stream is the
    one the reference test draws: gsl_rng_mt19937 seeded with 0 (= MT19937
    init_genrand(4357)), gsl_rng_uniform = 32-bit draw / 2^32."""
    bg = np.random.MT19937()
    bg._legacy_seeding(4357 if seed == 0 else seed)
    return clustered_mix_from(bg, n, box)


def clustered_mix_from(bg, n, box):
    u = bg.random_raw(3 * n).astype(np.float64) / 4294967296.0
    u = u.reshape(n, 3)
    pos = np.empty((n, 3))
    a, b = n // 4, 3 * n // 4
    pos[:a] = box * u[:a]
    pos[a:b] = box / 2 + box / 8 * np.exp((u[a:b] - 0.5) ** 2)
    pos[b:] = box * 0.1 + box / 32 * np.exp((u[b:] - 0.5) ** 2)
    return pos


def lattice(nc, box):
    """tests/test_gravity.c:231-235 / test_forcetree.c lattice."""
    i = np.arange(nc ** 3)
    return np.stack([(box / nc) * (i // nc // nc), (box / nc) * ((i // nc) % nc), (box / nc) * (i % nc)], 1).astype(np.float64)


def close_cluster(nc, close=5000.0):
    """tests/test_gravity.c:270-276: tight clump at 4 + idx/close."""
    i = np.arange(nc ** 3)
    return np.stack([4. + (i // nc // nc) / close, 4. + ((i // nc) % nc) / close, 4. + (i % nc) / close], 1)


def tree_params(box, n, treeusebh=0, errtol=0.002, rcut=6.0):
    """Defaults of gadget/params.c:112-116,192 (ErrTolForceAcc, BHOpeningAngle,
    MaxBHOpeningAngle, TreeRcut, GravitySoftening = 1/30 mean spacing)."""
    return dict(ErrTolForceAcc=errtol, BHOpeningAngle=0.175, MaxBHOpeningAngle=0.9, TreeUseBH=treeusebh,
                Rcut=rcut, GravitySoftening=(1 / 30.) * box / np.cbrt(n), rho0=OMEGA0 * rho_crit())


def planewave_lattice(ng, box, xplanes=None, seed=181170, rms=0.2, nmodes=64, kmax=None, device="cpu"):
    """Lattice displaced by a sum of random periodic plane waves (Zel'dovich-like,
    amplitude ~ k^-2, 3-D rms displacement `rms` spacings).  The displacement is a
    closed-form function of the lattice site, so any rank can generate any x-range
    of the SAME global particle set without a global FFT: used by the multi-GPU
    bench.  xplanes = (i0, i1) restricts to lattice planes i0 <= ix < i1 (may
    extend beyond [0, ng): wrapped).  Returns torch tensors pos[n,3] f64, mass[n] f32."""
    import torch
    rng = np.random.default_rng(seed)
    if kmax is None:
        kmax = max(4, ng // 8)
    nvec = rng.integers(-kmax, kmax + 1, size=(4 * nmodes, 3))
    nvec = nvec[(nvec ** 2).sum(1) > 0][:nmodes]
    kn = np.sqrt((nvec ** 2).sum(1))
    amp = rng.standard_normal(len(nvec)) / kn ** 2
    phase = rng.random(len(nvec)) * 2 * np.pi
    # rms of sum_m a_m khat_m sin(.) : sum a_m^2 / 2 (3-D)
    spacing = box / ng
    scale = rms * spacing / np.sqrt((amp ** 2).sum() / 2)
    i0, i1 = (0, ng) if xplanes is None else xplanes
    dev = torch.device(device)
    ix = torch.arange(i0, i1, dtype=torch.float64, device=dev)
    iy = torch.arange(ng, dtype=torch.float64, device=dev)
    qx = ((ix + 0.5) * spacing)[:, None, None].expand(len(ix), ng, ng)
    qy = ((iy + 0.5) * spacing)[None, :, None].expand(len(ix), ng, ng)
    qz = ((iy + 0.5) * spacing)[None, None, :].expand(len(ix), ng, ng)
    dx = torch.zeros((len(ix), ng, ng), dtype=torch.float64, device=dev)
    dy = torch.zeros_like(dx)
    dz = torch.zeros_like(dx)
    tw = 2 * np.pi / box
    for m in range(len(nvec)):
        arg = tw * (nvec[m, 0] * qx + nvec[m, 1] * qy + nvec[m, 2] * qz) + phase[m]
        s = torch.sin(arg) * (scale * amp[m] / kn[m])
        dx += s * float(nvec[m, 0]); dy += s * float(nvec[m, 1]); dz += s * float(nvec[m, 2])
    pos = torch.stack([(qx + dx).reshape(-1), (qy + dy).reshape(-1), (qz + dz).reshape(-1)], dim=1)
    pos = torch.remainder(pos, box)
    pos[pos >= box] = 0.0
    mass = torch.full((pos.shape[0],), OMEGA0 * rho_crit() * spacing ** 3, dtype=torch.float32, device=dev)
    return pos.contiguous(), mass


BENCH_STATES = {
    # lattice + plane-wave displacement, rms 1 spacing: leaf occupancy of the octree no longer depends on how the
    # lattice spacing divides the box (a 2^k lattice puts exactly 8 particles in every leaf), so timings at different
    # box sizes -- the weak-scaling series -- are like for like.  The default bench state.
    "displaced": dict(rms=1.0),
    # SURVEY 8d "z9": rms 0.2 spacing, the particles still sit next to their lattice sites
    "z9": dict(rms=0.2),
}


def bench_ics(state, ng, box, device="cpu", xplanes=None, seed=181170):
    """Synthetic boxes of bench.py / the config-size parity tests, as torch tensors pos[n,3] f64, mass[n] f32.
    'displaced' / 'z9': planewave_lattice at rms 1.0 / 0.2 spacings (any x-range of the same global set: xplanes).
    'clustered': SURVEY 8d's second state = the reference's do_random_test recipe (tests/test_gravity.c:288-302) at ng^3
    particles: 1/4 uniform, 1/2 in a clump at Box/2 of size Box/8, 1/4 in a clump at 0.1 Box of size Box/32 (deep tree)."""
    import torch
    if state in BENCH_STATES:
        return planewave_lattice(ng, box, xplanes=xplanes, seed=seed, device=device, **BENCH_STATES[state])
    if state != "clustered":
        raise ValueError("unknown bench state %r" % (state,))
    if xplanes is not None:
        raise ValueError("the clustered state is generated whole")
    n = ng ** 3
    pos = np.mod(clustered_mix(n, box, seed=seed), box)
    mass = np.full(n, OMEGA0 * rho_crit() * (box / ng) ** 3, dtype=np.float32)
    return torch.from_numpy(pos).to(device), torch.from_numpy(mass).to(device)


class FlatLCDM:
    """What the reference's host takes from cosmology.c / timefac.c / timebinmgr.c, for bench and tool runs of the step
    loop: hubble_function for a flat matter + Lambda background, the drift / kick integrals of timefac.c:12-73
    (Gauss-Legendre instead of gsl_integration_qag) and loga_from_ti over a sync-point table."""
    TIMEBINS = 46

    def __init__(self, sync_a=(0.1, 0.2, 0.5, 1.0), Omega0=OMEGA0, Hubble=0.1):
        self.sync = np.log(np.asarray(sync_a, np.float64)); self.Omega0 = Omega0; self.Hubble = Hubble
        self.gx, self.gw = np.polynomial.legendre.leggauss(16)

    def hubble(self, a):
        return self.Hubble * np.sqrt(self.Omega0 / a ** 3 + 1 - self.Omega0)

    def loga_from_ti(self, ti):
        s = ti >> self.TIMEBINS
        step = 0.0 if s >= len(self.sync) - 1 else (self.sync[s + 1] - self.sync[s]) / (1 << self.TIMEBINS)
        return self.sync[s] + (ti & ((1 << self.TIMEBINS) - 1)) * step

    def factor(self, kind, t0, t1):
        """kind 0 drift, 1 gravkick, 2 hydrokick between two integer times"""
        if t0 == t1:
            return 0.0
        a0, a1 = np.exp(self.loga_from_ti(t0)), np.exp(self.loga_from_ti(t1))
        e = np.linspace(a0, a1, 17)
        a = 0.5 * (e[1:] + e[:-1])[:, None] + 0.5 * (e[1:] - e[:-1])[:, None] * self.gx[None, :]
        w = 0.5 * (e[1:] - e[:-1])[:, None] * self.gw[None, :]
        h = self.hubble(a)
        f = (1 / (h * a ** 3), 1 / (h * a ** 2), 1 / (h * a ** (3 * (5.0 / 3 - 1)) * a))[kind]
        return float((w * f).sum())
