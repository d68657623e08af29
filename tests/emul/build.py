"""Build tests/emul/_build/libsteploop_emul.so: mp-gadget_b200/csrc/steploop.cu compiled for the HOST
against the CUDA stand-in headers of tests/emul/include, kernels and host drivers unchanged except
that the launch syntax  k<<<grid, block, smem, stream>>>(args);  is rewritten to a macro call.
TEST INFRASTRUCTURE ONLY (see tests/emul/include/cuda_runtime.h)."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "mp-gadget_b200", "csrc", "steploop.cu")
OUT = os.path.join(HERE, "_build")
SO = os.path.join(OUT, "libsteploop_emul.so")


def rewrite_launches(text):
    out, pos, count = [], 0, 0
    for m in re.finditer(r"(\w+)<<<", text):
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
    deps = [SRC, os.path.join(HERE, "emul_mocks.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "include", "cub", "cub.cuh"),
            os.path.join(ROOT, "mp-gadget_b200", "csrc", "engine.h"), os.path.join(ROOT, "include", "b200force.h"), __file__]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    text, n = rewrite_launches(open(SRC).read())
    assert n >= 10, n
    gen = os.path.join(OUT, "steploop_emul.cpp")
    with open(gen, "w") as f:
        f.write("// GENERATED from mp-gadget_b200/csrc/steploop.cu by tests/emul/build.py -- do not edit\n" + text)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    inc = ["-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "mp-gadget_b200", "csrc")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
                           "-Wno-unused-function", "-DSTEP_BLOCKS=4", "-o", SO, gen, os.path.join(HERE, "emul_mocks.cpp")] + inc +
                          ["-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return SO


if __name__ == "__main__":
    print(build(force=True))
