"""Device-resident dense matrices (iak3d_mat_*, include/iak3d.h) against numpy: the operations
predMatsIAK3D_CLEV_ByBlock (compositeLikIAK3D.R:609-798) is made of."""
import numpy as np
import pytest

import iak3d_b200 as iak
from iak3d_b200 import synth
from iak3d_b200._lib import make_pars
from iak3d_b200.devmat import DeviceMat, cov_mat
from helpers import case_pars, scaled_err
from oracle import iak3d_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return iak.Context(0)


def test_set_get_and_subblocks(ctx):
    rng = np.random.default_rng(1)
    a = rng.standard_normal((150, 71))
    M = DeviceMat.from_host(a, ctx)
    assert np.array_equal(M.to_host(), a)
    assert np.array_equal(M.to_host(3, 5, 130, 60), a[3:133, 5:65])
    M.set(np.ones((7, 9)), 100, 33)
    a[100:107, 33:42] = 1.0
    assert np.array_equal(M.to_host(), a)
    # padding stays zero: a product against ones only sees the logical part
    one = DeviceMat.from_host(np.ones((1, 71)), ctx)
    assert scaled_err(DeviceMat.matmul_nt(M, one).to_host().ravel(), a.sum(axis=1)) <= 1e-14


def test_copy_transpose_scale_accumulate(ctx):
    rng = np.random.default_rng(2)
    a = rng.standard_normal((97, 133)); d = rng.standard_normal((140, 120))
    A = DeviceMat.from_host(a, ctx); D = DeviceMat.from_host(d, ctx)
    D.assign(A, dr0=10, dc0=7, sr0=5, sc0=60, nr=80, nc=70, alpha=2.0, beta=-1.0)
    d[10:90, 7:77] = -d[10:90, 7:77] + 2.0 * a[5:85, 60:130]
    assert np.array_equal(D.to_host(), d)
    D.assign(A, dr0=1, dc0=2, sr0=3, sc0=4, nr=100, nc=90, transpose=True, alpha=0.5, beta=1.0)
    d[1:101, 2:92] += 0.5 * a[3:93, 4:104].T
    assert np.array_equal(D.to_host(), d)
    assert np.array_equal(A.t().to_host(), a.T)
    s = rng.standard_normal((70, 70))
    assert np.array_equal(DeviceMat.from_host(s, ctx).symmetrised().to_host(), 0.5 * (s + s.T))
    g = A.gather_blocks([(90, 7), (0, 33)], [(100, 30), (2, 5)])
    assert np.array_equal(g.to_host(), a[np.ix_(np.r_[90:97, 0:33], np.r_[100:130, 2:7])])


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (130, 45, 70), (128, 64, 32), (300, 257, 129), (5, 700, 1000)])
def test_gemm_nt(ctx, m, n, k):
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((m, k)); b = rng.standard_normal((n, k)); c = rng.standard_normal((m, n))
    A = DeviceMat.from_host(a, ctx); B = DeviceMat.from_host(b, ctx); Cm = DeviceMat.from_host(c, ctx)
    assert scaled_err(DeviceMat.matmul_nt(A, B, alpha=-1.5).to_host(), -1.5 * a @ b.T) <= 1e-13
    Cm.gemm_nt(A, B, alpha=2.0, accumulate=True)
    assert scaled_err(Cm.to_host(), c + 2.0 * a @ b.T) <= 1e-13
    with pytest.raises(iak.Iak3dError):
        Cm.gemm_nt(Cm, B)


@pytest.mark.parametrize("n", [1, 63, 64, 129, 700])
def test_spd_inverse(ctx, n):
    rng = np.random.default_rng(n)
    g = rng.standard_normal((n, n + 5))
    s = g @ g.T / n + np.eye(n)
    M = DeviceMat.from_host(s, ctx).spd_inverse()
    got = M.to_host()
    assert scaled_err(got, np.linalg.inv(s)) <= 1e-11
    assert np.array_equal(got, got.T)
    # the padding is still zero after the in-place inverse
    one = DeviceMat.from_host(np.ones((1, n)), ctx)
    assert scaled_err(DeviceMat.matmul_nt(M, one).to_host().ravel(), got.sum(axis=1)) <= 1e-12


def test_spd_inverse_rejects_indefinite(ctx):
    s = np.eye(40); s[7, 7] = -1.0
    M = DeviceMat.from_host(s, ctx)
    with pytest.raises(np.linalg.LinAlgError):
        M.spd_inverse()
    assert np.array_equal(M.to_host(), s)


def test_coldot_and_cov_mat(ctx):
    rng = np.random.default_rng(9)
    a = rng.standard_normal((333, 77)); b = rng.standard_normal((333, 77))
    assert scaled_err(DeviceMat.from_host(a, ctx).coldot(DeviceMat.from_host(b, ctx)), (a * b).sum(axis=0)) <= 1e-13
    ds = synth.make_dataset(310, 17)
    for kind, nonstat in (("edgeroi", False), ("matern0.5", True)):
        P, modelx, types, cmeOpt = case_pars(kind, nonstat, 0.5)
        sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
        Cd = cov_mat(sm, make_pars(P, modelx, *types, cmeOpt))
        want, _ = orc.setCIAK3D(P, modelx, *types, cmeOpt, orc.setupIAK3D(ds["xData"], ds["dIData"]))
        assert scaled_err(Cd.to_host(), want) <= 1e-10
        assert np.array_equal(Cd.to_host(), iak.setCIAK3D(P, modelx, *types, cmeOpt, sm)["C"])
