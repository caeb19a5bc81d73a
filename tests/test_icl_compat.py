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
    L.ref_icl_script.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
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


class SegmentModel:
    """Boost.ICL interval_map<int, int, partial_absorber, ..., inplace_plus> as its documentation defines it ("Addition on interval
    maps", "Subtraction", "partial absorber: identity elements are not stored; subtraction only affects stored keys"), written here
    independently of the stand-in: a sorted list of [l, r, v] segments.  split: every border ever inserted survives while the
    segment around it does; join: touching segments of equal value are merged."""

    def __init__(self, join):
        self.s, self.join = [], join

    def _cut(self, p):
        for i, (l, r, v) in enumerate(self.s):
            if l < p < r:
                self.s[i:i + 1] = [[l, p, v], [p, r, v]]
                return

    def add(self, l, r, v):
        if l >= r or v == 0:
            return
        self._cut(l)
        self._cut(r)
        out, cur = [], l
        for a, b, w in self.s:
            if b <= l or a >= r:
                out.append([a, b, w])
                continue
            if cur < a:
                out.append([cur, a, v])
            out.append([a, b, w + v])
            cur = b
        if cur < r:
            out.append([cur, r, v])
        self.s = sorted(x for x in out if x[2] != 0)
        self._merge()

    def sub(self, l, r, v):
        if l >= r or v == 0:
            return
        self._cut(l)
        self._cut(r)
        self.s = [x for x in ([a, b, w - v] if (a >= l and b <= r) else [a, b, w] for a, b, w in self.s) if x[2] != 0]
        self._merge()

    def _merge(self):
        if not self.join:
            return
        out = []
        for a, b, w in self.s:
            if out and out[-1][1] == a and out[-1][2] == w:
                out[-1][1] = b
            else:
                out.append([a, b, w])
        self.s = out


@pytest.mark.parametrize("join", [0, 1])
def test_random_scripts_with_subtraction_and_map_addition(lib, join):
    """`-=` (rnacore/region.cc, the reference's graph revision) and `map += map` (bundle::combine, meta/bundle.cc:102) next to
    `+=`, against the documented semantics restated independently above"""
    rng = np.random.default_rng(101 + join)
    for trial in range(300):
        n = int(rng.integers(1, 50))
        op = rng.choice([0, 0, 0, 1, 1, 2, 2, 3, 4], n).astype(np.int32)
        l = rng.integers(0, 60, n).astype(np.int32)
        r = (l + rng.integers(0, 12, n)).astype(np.int32)
        v = rng.integers(0, 4, n).astype(np.int32)
        out = np.zeros(3 * 600, np.int32)
        k = lib.ref_icl_script(n, op.ctypes.data, l.ctypes.data, r.ctypes.data, v.ctypes.data, join, out.ctypes.data, len(out))
        got = out[:k].reshape(-1, 3).tolist()
        A, B = SegmentModel(join), SegmentModel(join)
        for o, a, b, w in zip(op.tolist(), l.tolist(), r.tolist(), v.tolist()):
            if o == 0:
                A.add(a, b, w)
            elif o == 1:
                A.sub(a, b, w)
            elif o == 2:
                B.add(a, b, w)
            elif o == 3:
                for x in list(B.s):
                    A.add(*x)
            elif o == 4:
                for x in list(B.s):
                    A.sub(*x)
        assert got == A.s, (trial, got, A.s)


def test_subtraction_known_answers(lib):
    """hand-derived from the ICL documentation: subtraction reaches stored segments only, splits them at the operand's ends,
    and a segment that falls to 0 disappears (with its borders)"""
    def run(script, join=0):
        op, l, r, v = (np.array(x, np.int32) for x in zip(*script))
        out = np.zeros(300, np.int32)
        k = lib.ref_icl_script(len(script), op.ctypes.data, l.ctypes.data, r.ctypes.data, v.ctypes.data, join, out.ctypes.data, len(out))
        return out[:k].reshape(-1, 3).tolist()
    assert run([(0, 10, 20, 2), (1, 12, 15, 1)]) == [[10, 12, 2], [12, 15, 1], [15, 20, 2]]
    assert run([(0, 10, 20, 2), (1, 5, 30, 2)]) == []                                     # nothing left, nothing created in the gaps
    assert run([(0, 10, 20, 1), (0, 15, 25, 1), (1, 15, 20, 2)]) == [[10, 15, 1], [20, 25, 1]]
    assert run([(0, 10, 20, 1), (1, 15, 20, 1), (0, 12, 18, 1)]) == [[10, 12, 1], [12, 15, 2], [15, 18, 1]]
    assert run([(2, 0, 5, 1), (2, 3, 8, 2), (0, 4, 6, 1), (3, 0, 0, 0)]) == [[0, 3, 1], [3, 4, 3], [4, 5, 4], [5, 6, 3], [6, 8, 2]]
    assert run([(0, 0, 10, 3), (1, 0, 10, 1)], join=1) == [[0, 10, 2]]
