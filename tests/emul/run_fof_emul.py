"""mp-gadget_b200/csrc/fof.cu -- kernels and host driver, source unchanged -- on the CPU emulation of tests/emul against the
golden labels of the reference's own fof.c (tests/golden/ref_fof.npz) and the oracle on further cases (tiny boxes where the
grid has one cell, garbage particles, every particle its own group, one group).  TEST INFRASTRUCTURE ONLY.  Started by
tests/test_fof.py in a subprocess with OMP_WAIT_POLICY=passive."""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import build as EB                   # noqa: E402
import domain_scenarios as DS        # noqa: E402
import oracle                        # noqa: E402


def run(L, pos, ids, typ, box, ll, mask=2, flags=None):
    pos = np.ascontiguousarray(pos, np.float64); ids = np.ascontiguousarray(ids, np.int64); typ = np.ascontiguousarray(typ, np.uint8)
    n = len(ids)
    out = np.empty(n, np.int64); ng = C.c_int64(0)
    p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    fl = np.ascontiguousarray(flags, np.uint8) if flags is not None else None
    rc = L.emul_fof_primary(C.c_int64(n), p(pos), p(typ), p(fl), p(ids), C.c_int(mask), C.c_double(box), C.c_double(ll), p(out), C.byref(ng))
    assert rc == 0
    return out, ng.value


def main():
    L = C.CDLL(EB.build_fof())
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_fof.npz"))
    for mode in ("clique", "cells"):             # the clique-cell search (default where the grid allows) and the plain cell list
        os.environ["B200_FOF"] = mode
        for k, (pos, ids, typ, box, ll) in enumerate(DS.fof_cases()):
            got, ng = run(L, pos, ids, typ, box, ll)
            assert np.array_equal(got, G["%d/minid" % k]), (mode, k)
            assert ng == len(np.unique(got[typ == 1])), (mode, ng, k)
        for pos, ids, typ, box, ll, mask, flags in DS.fof_edge_cases():
            got, ng = run(L, pos, ids, typ, box, ll, mask, flags)
            keep = np.ones(len(ids), bool) if flags is None else (flags & 3) == 0
            want = ids.copy()
            want[keep] = oracle.fof_primary(pos[keep], ids[keep], typ[keep], box, ll, mask=mask)
            assert np.array_equal(got, want), (mode, box, ll, mask)
    print("fof ok")


if __name__ == "__main__":
    main()
