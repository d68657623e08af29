"""mp-gadget_b200/csrc/pm_fft.cu -- kernels and host driver, source unchanged -- on the CPU emulation of tests/emul:
density mesh -> potential mesh against numpy's rfftn / Green's function / irfftn, and the power-spectrum sums of the
fused x pass against a direct numpy restatement of powerspectrum_add_mode (gravpm.c:330-361).  TEST INFRASTRUCTURE ONLY.
Started by tests/test_pm_fft_emul.py in a subprocess (OMP_WAIT_POLICY=passive, B200_FFT_THREADS=64)."""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build as EB      # noqa: E402


def green(N, asmth2, pot_factor, ktab):
    k = np.arange(N)
    k = np.where(k <= N // 2, k, k - N)
    kx, ky, kz = k[:, None, None], k[None, :, None], np.arange(N // 2 + 1)[None, None, :]
    k2 = (kx * kx + ky * ky + kz * kz).astype(float)
    f = ktab[:, None, None] * ktab[None, :, None] * ktab[None, None, :N // 2 + 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        g = pot_factor * np.exp(-k2 * asmth2) / k2 * f * f
    g[0, 0, 0] = 0
    return g, k2, f


def main(sizes):
    L = C.CDLL(EB.build_fft())
    p = lambda a: C.c_void_p(a.ctypes.data)
    for N in sizes:
        assert L.emul_pmfft_supported(C.c_int(N)) == 1, N
        rng = np.random.default_rng(N)
        mesh = rng.random((N, N, N))
        ktab = 1 + rng.random(N)
        asmth2, pf = (2 * np.pi * 1.5 / N) ** 2, -43.0071 / (np.pi * 100.)
        bpu = (N - 1) / np.log(np.sqrt(3.) * N / 2)
        spec = np.fft.rfftn(mesh, axes=(0, 1, 2))
        g, k2, f = green(N, asmth2, pf, ktab)
        want = np.fft.irfftn(spec * g, s=(N, N, N), axes=(0, 1, 2)) * float(N) ** 3
        for power in (False, True):
            got = mesh.copy()
            ps = np.zeros(3 * N + 1)
            rc = L.emul_pmfft_potential(C.c_int(N), p(got), p(ktab), C.c_double(asmth2), C.c_double(pf), C.c_double(bpu), p(ps) if power else None)
            assert rc == 0
            err = np.abs(got - want).max() / np.abs(want).max()
            assert err < 1e-13, (N, power, err)
            if power:
                with np.errstate(divide="ignore"):
                    kint = np.floor(bpu * np.log(k2) / 2.)
                kint[0, 0, 0] = N
                w = np.full(k2.shape, 2.0)
                w[:, :, 0] = 1.0
                w[:, :, N // 2] = 1.0
                sel = kint < N
                b = kint[sel].astype(int)
                opw = np.bincount(b, weights=(w * np.abs(spec) ** 2 * f * f)[sel], minlength=N)
                okk = np.bincount(b, weights=(w * np.sqrt(k2))[sel], minlength=N)
                onm = np.bincount(b, weights=w[sel], minlength=N)
                assert np.array_equal(ps[2 * N:3 * N], onm), N
                assert np.abs(ps[N:2 * N] - okk).max() <= 1e-12 * okk.max()
                assert np.abs(ps[:N] - opw).max() <= 1e-12 * opw.max()
                assert abs(ps[3 * N] - np.abs(spec[0, 0, 0]) ** 2) <= 1e-12 * ps[3 * N]
        print("N %d ok (max rel err %.2g)" % (N, err))
    assert L.emul_pmfft_supported(C.c_int(14)) == 0 and L.emul_pmfft_supported(C.c_int(33)) == 0
    print("fft ok")


def c2r():
    """b200_pm_c2r_readout's device side (k_pm_apply_radial, the three inverse passes, k_pm_readout_field) against the golden
    of the reference's own petapm_force_c2r (tests/golden/ref_pm_c2r.npz)."""
    ROOT = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pm_c2r_scenarios as SC
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_pm_c2r.npz"))
    L = C.CDLL(EB.build_fft())
    p = lambda a: C.c_void_p(a.ctypes.data)
    for name, pos, box, nmesh, rho_k, fn in SC.cases():
        pos = np.ascontiguousarray(pos); rk = np.ascontiguousarray(rho_k)
        kinds = np.array([k for k, _ in fn], np.int32)
        tabs = np.ascontiguousarray(np.stack([t for _, t in fn]))
        out = np.zeros((len(fn), len(pos)))
        rc = L.emul_pm_c2r_readout(C.c_int(nmesh), C.c_double(box), C.c_int64(len(pos)), p(pos), p(rk), C.c_int(len(fn)), p(kinds), p(tabs),
                                   C.c_int64(tabs.shape[1]), p(out))
        assert rc == 0
        for j in range(len(fn)):
            want = G["%s/out%d" % (name, j)]
            err = np.abs(out[j] - want).max() / np.abs(want).max()
            assert err < 1e-12, (name, j, err)
    print("c2r ok")


if __name__ == "__main__":
    if sys.argv[1:] == ["c2r"]:
        c2r()
    else:
        main([int(a) for a in sys.argv[1:]] or [10, 16, 24, 40])
