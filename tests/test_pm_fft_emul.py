"""csrc/pm_fft.cu (the PM step's shared-memory transform passes with the Green's function inside) executed on the CPU
under the CUDA stand-in of tests/emul and compared with numpy: the kernels were written and checked here before their first
hardware run; tests/test_gpu_parity.py::test_pm_own_transforms_equal_cufft and every PM parity test are the hardware tests."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_pm_fft_source_under_emulation():
    """Mesh sizes covering every stage kind: 10 = 2.5 (odd half length, ragged last tile), 16 and 24 (fused radix-16 and
    radix-8 groups, radix 3), 40 (radix 5), 18 / 36 / 50 / 54 (chains of several radix-3 and radix-5 stages)."""
    env = dict(os.environ, OMP_WAIT_POLICY="passive", B200_FFT_THREADS="64")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "run_fft_emul.py"), "10", "16", "18", "20", "24", "32", "36", "40", "48", "50", "54"], env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "fft ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_pm_fft_emulation_under_address_sanitizer():
    """The same with the emulated source compiled -fsanitize=address,undefined: an out-of-range row or line index aborts."""
    import pytest
    libs = [subprocess.run(["gcc", "-print-file-name=" + l], capture_output=True, text=True).stdout.strip() for l in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(l) and os.path.exists(l) for l in libs):
        pytest.skip("libasan / libubsan not available")
    env = dict(os.environ, OMP_WAIT_POLICY="passive", B200_FFT_THREADS="64", EMUL_ASAN="1", LD_PRELOAD=" ".join(libs), ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "run_fft_emul.py"), "10", "16", "24", "40"], env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "fft ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
