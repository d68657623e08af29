"""Build tests/emul/_build/libsteploop_emul.so: mp-gadget_b200/csrc/steploop.cu and domain_keys.cu compiled for the HOST
against the CUDA stand-in headers of tests/emul/include, kernels and host drivers unchanged except
that the launch syntax  k<<<grid, block, smem, stream>>>(args);  is rewritten to a macro call.
TEST INFRASTRUCTURE ONLY (see tests/emul/include/cuda_runtime.h)."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRCS = [os.path.join(ROOT, "mp-gadget_b200", "csrc", f) for f in ("steploop.cu", "domain_keys.cu")]
SRC = SRCS[0]
ASAN = os.environ.get("EMUL_ASAN") == "1"     # address-sanitised build: run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
OUT = os.path.join(HERE, "_build_asan" if ASAN else "_build")
SO = os.path.join(OUT, "libsteploop_emul.so")


def rewrite_launches(text):
    out, pos, count = [], 0, 0
    for m in re.finditer(r"(\w+(?:<\w+>)?)<<<", text):
        start = m.start()
        end_cfg = text.index(">>>", m.end())
        cfg = [c.strip() for c in text[m.end():end_cfg].split(",")]
        assert text[end_cfg + 3] == "(", text[start:end_cfg + 10]
        depth, k = 0, end_cfg + 3
        while True:                      # matching parenthesis of the argument list
            depth += text[k] == "("; depth -= text[k] == ")"
            if depth == 0:
                break
            k += 1
        args = text[end_cfg + 4:k]
        out.append(text[pos:start])
        out.append("EMUL_LAUNCH(%s, %s, %s, %s)" % (m.group(1), cfg[0], cfg[1], args))
        pos = k + 1
        count += 1
    out.append(text[pos:])
    return "".join(out), count


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    deps = SRCS + [os.path.join(HERE, "emul_mocks.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "include", "cub", "cub.cuh"),
            os.path.join(ROOT, "mp-gadget_b200", "csrc", "engine.h"), os.path.join(ROOT, "include", "b200force.h"), __file__]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    gens, total = [], 0
    for src in SRCS:
        text, n = rewrite_launches(open(src).read())
        total += n
        gen = os.path.join(OUT, os.path.basename(src).replace(".cu", "_emul.cpp"))
        with open(gen, "w") as f:
            f.write("// GENERATED from mp-gadget_b200/csrc/%s by tests/emul/build.py -- do not edit\n" % os.path.basename(src) + text)
        gens.append(gen)
    assert total >= 12, total
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    inc = ["-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "mp-gadget_b200", "csrc")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off"] + (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if ASAN else []) + [ "-Wall", "-Wno-unknown-pragmas",
                           "-Wno-unused-function", "-DSTEP_BLOCKS=4", "-o", SO] + gens + [os.path.join(HERE, "emul_mocks.cpp")] + inc +
                          ["-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return SO


DROPIN = os.path.join(OUT, "libref_dropin_step_emul.so")
REF = "/root/reference"


def build_dropin(force=False):
    """The reference's own step-loop driver (oracle/ref_driver.c + its timestep.c, drift.c, tree gravity ...) with the
    calls redirected by ld --wrap to mp-gadget_b200/host/libgadget_step_shims.c, linked against the EMULATION
    library instead of libb200force.so.  Needs /root/reference; returns None without it."""
    if not os.path.isdir(os.path.join(REF, "libgadget")):
        return DROPIN if os.path.exists(DROPIN) else None
    emul = build(force)
    host = os.path.join(ROOT, "mp-gadget_b200", "host")
    orc = os.path.join(ROOT, "oracle")
    deps = [emul, os.path.join(host, "libgadget_step_shims.c"), os.path.join(host, "step_wrap.opts"), os.path.join(orc, "ref_driver.c")]
    if not force and os.path.exists(DROPIN) and all(os.path.getmtime(d) <= os.path.getmtime(DROPIN) for d in deps):
        return DROPIN
    L = os.path.join(REF, "libgadget")
    names = ["forcetree.c", "treewalk.c", "gravshort-tree.c", "gravity.c", "partmanager.c", "slotsmanager.c", "walltime.c", "utils/peano.c",
             "utils/memory.c", "utils/mymalloc.c", "utils/openmpsort.c", "utils/endrun.c", "utils/system.c", "utils/string.c", "density.c",
             "densitykernel.c", "hydra.c", "drift.c", "timestep.c", "timebinmgr.c"]         # SRCS_STEP of oracle/Makefile.ref
    subprocess.check_call(["gcc", "-fopenmp", "-O3", "-g", "-ffast-math", "-fPIC", "-std=gnu11", "-w", "-Werror=implicit-function-declaration",
                           "-I" + os.path.join(orc, "stubs"), "-I" + REF, "-I" + L, "-I" + os.path.join(REF, "depends", "bigfile", "src"),
                           "-DREF_WITH_STEP", "-shared", "-o", DROPIN, os.path.join(orc, "ref_driver.c")] + [os.path.join(L, f) for f in names] +
                          [os.path.join(host, "libgadget_step_shims.c"), os.path.join(host, "libgadget_shim_ctx.c"),
                           "@" + os.path.join(host, "step_wrap.opts"), "-L" + OUT, "-lsteploop_emul", "-Wl,-rpath," + OUT, "-lm"])
    return DROPIN


if __name__ == "__main__":
    print(build(force=True))
    print(build_dropin(force=True))


FFT_SO = os.path.join(OUT, "libpmfft_emul.so")


def build_fft(force=False):
    """mp-gadget_b200/csrc/pm_fft.cu (kernels and host driver unchanged) for the host, driven by tests/emul/fft_driver.cpp."""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(ROOT, "mp-gadget_b200", "csrc", "pm_fft.cu")
    drv = os.path.join(HERE, "fft_driver.cpp")
    deps = [src, drv, os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(ROOT, "mp-gadget_b200", "csrc", "engine.h"), __file__]
    if not force and os.path.exists(FFT_SO) and all(os.path.getmtime(d) <= os.path.getmtime(FFT_SO) for d in deps):
        return FFT_SO
    text, n = rewrite_launches(open(src).read())
    assert n >= 6, n
    gen = os.path.join(OUT, "pm_fft_emul.cpp")
    with open(gen, "w") as f:
        f.write("// GENERATED from mp-gadget_b200/csrc/pm_fft.cu by tests/emul/build.py -- do not edit\n" + text)
    inc = ["-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "mp-gadget_b200", "csrc")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off"] +
                          (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if ASAN else []) +
                          ["-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-o", FFT_SO, gen, drv] + inc +
                          (["-DFFT_T=" + os.environ["EMUL_FFT_T"]] if os.environ.get("EMUL_FFT_T") else []))
    return FFT_SO


FOF_SO = os.path.join(OUT, "libfof_emul.so")


def build_fof(force=False):
    """mp-gadget_b200/csrc/fof.cu (kernels and host driver unchanged) for the host, driven by tests/emul/fof_driver.cpp."""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(ROOT, "mp-gadget_b200", "csrc", "fof.cu")
    drv = os.path.join(HERE, "fof_driver.cpp")
    deps = [src, drv, os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "include", "cub", "cub.cuh"),
            os.path.join(ROOT, "mp-gadget_b200", "csrc", "engine.h"), __file__]
    if not force and os.path.exists(FOF_SO) and all(os.path.getmtime(d) <= os.path.getmtime(FOF_SO) for d in deps):
        return FOF_SO
    text, n = rewrite_launches(open(src).read())
    assert n >= 7, n
    gen = os.path.join(OUT, "fof_emul.cpp")
    with open(gen, "w") as f:
        f.write("// GENERATED from mp-gadget_b200/csrc/fof.cu by tests/emul/build.py -- do not edit\n" + text)
    inc = ["-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "mp-gadget_b200", "csrc")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off"] +
                          (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if ASAN else []) +
                          ["-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-o", FOF_SO, gen, drv] + inc)
    return FOF_SO
