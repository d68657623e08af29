"""csrc/steploop.cu -- kernels and host drivers, source unchanged -- executed on the CPU under the CUDA
stand-in of tests/emul (blocks serial, the threads of a block an OpenMP team; tree build and walk
supplied by the oracle) and checked against the golden vectors of the reference's own drift.c /
timestep.c / timebinmgr.c.  This is how the step-loop CUDA code, written in a round whose GPU budget was
already spent, was checked before its first hardware run; tests/test_step_gpu.py is the hardware test."""
import os
import subprocess
import sys
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("which", ["primitives", "hierarchy", "gas", "nonsplit", "dropin_primitives", "dropin_hierarchy", "dropin_gas", "domain", "decompose"])
def test_steploop_source_under_emulation(which):
    """dropin_*: the reference's own loop with its calls redirected (ld --wrap) to host/libgadget_step_shims.c."""
    env = dict(os.environ, OMP_WAIT_POLICY="passive")          # 256 OS threads per emulated block: do not spin
    r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "run_emul.py"), which], env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    if "skip:" in r.stdout:
        pytest.skip(r.stdout.strip().splitlines()[-1])
    assert which + " ok" in r.stdout


def test_steploop_emulation_under_address_sanitizer():
    """The same runs with the emulated CUDA sources compiled -fsanitize=address,undefined: every `device` buffer is a
    red-zoned heap block, so an out-of-range index in a kernel or a host driver aborts here instead of corrupting memory on
    the GPU, and so do shifts / conversions with undefined results."""
    libs = [subprocess.run(["gcc", "-print-file-name=" + l], capture_output=True, text=True).stdout.strip() for l in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(l) and os.path.exists(l) for l in libs):
        pytest.skip("libasan / libubsan not available")
    env = dict(os.environ, OMP_WAIT_POLICY="passive", EMUL_ASAN="1", LD_PRELOAD=" ".join(libs), ASAN_OPTIONS="detect_leaks=0")
    for which in ("primitives", "hierarchy", "domain"):
        r = subprocess.run([sys.executable, os.path.join(HERE, "emul", "run_emul.py"), which], env=env, capture_output=True, text=True, timeout=1800)
        assert r.returncode == 0 and which + " ok" in r.stdout, r.stdout[-2000:] + r.stderr[-6000:]


@pytest.mark.parametrize("world", [2, 3])
def test_domain_decompose_emulated_ranks(world):
    """mp-gadget_b200/domain.py::decompose + exchange with one emulated engine per gloo rank: the device side of the chain
    (subsample keys, keys, top-leaf lookup, counts, exchange plan) is the CUDA source under emulation."""
    env = dict(os.environ, OMP_WAIT_POLICY="passive")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(31500 + os.getpid() % 2000), os.path.join(HERE, "emul", "run_emul.py"), "decompose"],
                       env=env, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("decompose ok") == world
