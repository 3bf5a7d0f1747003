"""CPU-side checks: the C-ABI library loads, exports every symbol the header declares, fails loudly
without a GPU, and the host arithmetic (ADI shifts, packing, sharding) is right."""
import ctypes
import os
import re

import numpy as np
import pytest

import pop_b200 as P
from oracle import arrowhead_oracle as O
from tests.util import random_general

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pop_arrowhead.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pop_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = ctypes.CDLL(P.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pop_arrowhead.h but not exported"
    assert set(syms) == set(P.EXPORTED), "ctypes prototypes out of sync with the header"
    assert lib.pop_abi_version() == 1


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(P.PopError) as ei:
        P.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "piecewiseorthogonalpolynomials.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace("# oracle", ""), f


@pytest.mark.parametrize("a,b", [(1.9262e-05, 0.405285), (3.2e-10, 4 / np.pi ** 2), (1e-3, 0.4), (0.05, 0.4)])
def test_adi_shifts_match_reference_formulas(a, b):
    """examples/adi_poisson.jl:23-40,47-49; oracle uses scipy.special (cephes dn is only ~5e-9
    accurate near m=1, so the elliptic branch is compared at that level and against mpmath)."""
    import mpmath as mp
    c, d = -b, -a
    J = P.adi_num_iterations(a, b, c, d, 1e-15)
    assert J == O.adi_num_iterations(a, b, c, d, 1e-15)
    p, q = P.adi_shifts(J, a, b, c, d)
    po, qo = O.adi_shifts(J, a, b, c, d)
    np.testing.assert_allclose(p, po, rtol=2e-8)
    np.testing.assert_allclose(q, qo, rtol=2e-8)
    gamma = (c - a) * (d - b) / ((c - b) * (d - a))
    alpha = -1 + 2 * gamma + 2 * np.sqrt(gamma ** 2 - gamma)
    if alpha < 1e7:
        mp.mp.dps = 40
        al = -1 + 2 * mp.mpf(gamma) + 2 * mp.sqrt(mp.mpf(gamma) ** 2 - gamma)
        mm = 1 - 1 / al ** 2
        K = mp.ellipk(mm)
        pm = [float(O.mobius(-al * mp.ellipfun("dn", (2 * i + 1) * K / (2 * J), mm), a, b, c, d, al)) for i in range(J)]
        np.testing.assert_allclose(p, pm, rtol=1e-10)
    assert np.all(p > 0) and np.all(q < 0)


def test_packing_matches_oracle_layout():
    from pop_b200.arrowhead import _pack_blocks
    rng = np.random.default_rng(3)
    for lower in (True, False):
        A = random_general(rng, 6, 5, nB=2, nC=2, lower_B=lower, same_D=False)
        pk = O.pack(A)
        pp = _pack_blocks(A.A, A.B, A.C, A.D)
        for k in ("Ad", "Au", "Al", "B", "C", "D"):
            assert np.array_equal(pp[k], getattr(pk, k)), k
    with pytest.raises(P.StructureError):
        bad = random_general(rng, 6, 5)
        bad.B[0][5, 0] = 1.0          # outside bandwidths (1,1): src/arrowhead.jl:100-105
        _pack_blocks(bad.A, bad.B, bad.C, bad.D)
    with pytest.raises(P.DimensionMismatch):
        bad = random_general(rng, 6, 5)
        bad.B[0] = bad.B[0][:, :-1]
        _pack_blocks(bad.A, bad.B, bad.C, bad.D)


def test_shard_bounds():
    for n, w in ((8191, 8), (8191, 4), (31, 2), (5, 8), (16, 1)):
        b = P.shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_dd_columns_partition():
    """Element decomposition (csrc/pop_row.cu): the ranks' local column sets partition the columns, hats first."""
    for m, nh, q, world in ((128, 127, 63, 8), (16, 17, 5, 3), (9, 8, 4, 2), (4, 3, 2, 4)):
        n = nh + m * q
        allc = np.concatenate([P.dd_columns(m, nh, q, r, world) for r in range(world)])
        assert sorted(allc.tolist()) == list(range(n))
        c0 = P.dd_columns(m, nh, q, 0, world)
        assert c0[0] == 0 and (len(c0) < 2 or c0[1] in (1, nh))
