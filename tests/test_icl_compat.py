"""CPU tier: the Boost.ICL stand-in used to build the reference (oracle/compat/boost/icl/interval_map.hpp) against
(i) the known answers derived by hand from the reference's own print-only smoke functions
    (rnacore/interval_map.cc:320-451) and
(ii) a naive per-base model under random positive additions."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(orclib.REF_SO):
        pytest.skip("oracle/_ref not built")
    L = C.CDLL(orclib.REF_SO)
    L.ref_icl_kat.argtypes = [C.c_void_p, C.c_int]
    L.ref_icl_apply.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return L


def test_known_answers(lib):
    out = np.zeros(256, np.int32)
    n = lib.ref_icl_kat(out.ctypes.data, len(out))
    out = out[:n].tolist()
    cut = out.index(-1)
    assert out[:cut] == [1, 2, 4, 2, 3, 5, 3, 4, 2, 4, 5, 2, 6, 7, 3]
    rest = out[cut + 1:]
    cut2 = rest.index(-1)
    cov = rest[:cut2]
    segs = [(1, 2), (2, 3), (3, 4), (4, 5), (6, 7)]
    want = []
    for i in range(9):
        for j in range(i, 9):
            want.append(sum(r - l for l, r in segs if l >= i and r <= j))          # segments fully inside [i, j)
    assert cov == want
    assert rest[cut2 + 1:] == [4, 2, 2, 5]


@pytest.mark.parametrize("join", [0, 1])
def test_random_positive_additions(lib, join):
    rng = np.random.default_rng(7 + join)
    for trial in range(200):
        n = int(rng.integers(1, 40))
        l = rng.integers(0, 60, n).astype(np.int32)
        r = (l + rng.integers(0, 12, n)).astype(np.int32)          # includes empty intervals
        v = rng.integers(0, 4, n).astype(np.int32)                 # includes the identity value
        out = np.zeros(3 * 400, np.int32)
        k = lib.ref_icl_apply(n, l.ctypes.data, r.ctypes.data, v.ctypes.data, join, out.ctypes.data, len(out))
        got = out[:k].reshape(-1, 3).tolist()
        cov = np.zeros(80, np.int64)
        borders = set()
        for a, b, w in zip(l, r, v):
            if a < b and w != 0:
                cov[a:b] += w
                borders.update((int(a), int(b)))
        want = []
        if join:
            p = 0
            while p < 80:
                if cov[p] == 0:
                    p += 1
                    continue
                q = p
                while q < 80 and cov[q] == cov[p]:
                    q += 1
                want.append([p, q, int(cov[p])])
                p = q
        else:
            bs = sorted(borders)
            for a, b in zip(bs[:-1], bs[1:]):
                if cov[a] > 0:
                    want.append([a, b, int(cov[a])])
        assert got == want, (trial, got, want)
