"""GPU: the STRICT flavour's arithmetic building blocks are bit-identical to their specification:
  * det sin/cos/log/exp and the Philox normals == the CPU oracle's restatement (oracle/det_math.h);
  * branch-free fdiv / fsqrt / normalize3 == the IEEE intrinsics __fdiv_rn / __fsqrt_rn, including the
    MUFU-free shortcut for almost-unit vectors and the rsqrt-seeded reciprocal."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from mppi_b200 import capi
    return capi.lib()


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_det_math_matches_oracle_bitwise(L, oracle):
    import torch
    rng = np.random.default_rng(0)
    cases = {
        0: np.concatenate([np.linspace(-0.2, 0.2, 400001), rng.uniform(-300, 300, 600000)]).astype(np.float32),
        1: (rng.integers(0, 1 << 24, 1000000).astype(np.float32) * np.float32(2.0 ** -24)),
        2: np.concatenate([(rng.integers(1, 1 << 24, 600000).astype(np.float32) * np.float32(2.0 ** -24)),
                           rng.uniform(1e-30, 1e30, 400000).astype(np.float32)]),
        3: np.concatenate([np.linspace(-90, 10, 600001), -rng.exponential(5.0, 400000)]).astype(np.float32),
        4: np.concatenate([np.linspace(-4, 4, 400001), rng.standard_cauchy(400000) * 10,
                           np.array([0.0, -0.0, np.inf, -np.inf, 1e30, -1e-30])]).astype(np.float32),
    }
    for fn, x in cases.items():
        xd = dev(x)
        y0 = torch.empty_like(xd)
        y1 = torch.empty_like(xd)
        assert L.mppi_test_detmath(fn, xd.data_ptr(), y0.data_ptr(), y1.data_ptr(), x.size, None) == 0
        torch.cuda.synchronize()
        r0, r1 = oracle.detmath(fn, x)
        assert np.array_equal(y0.cpu().numpy().view(np.uint32), r0.view(np.uint32)), fn
        if fn < 2:
            assert np.array_equal(y1.cpu().numpy().view(np.uint32), r1.view(np.uint32)), fn


@pytest.mark.parametrize("K,T,k0", [(4096, 100, 0), (333, 7, 5000), (17, 2, 2 ** 31)])
def test_philox_normals_match_oracle_bitwise(L, oracle, K, T, k0):
    import torch
    e1 = torch.zeros(K * T, device="cuda")
    e2 = torch.zeros(K * T, device="cuda")
    seed, off = 0x1234567890ABCDEF, 0x0000000500000003
    assert L.mppi_test_noise(seed, off, 3, k0, K, T, 0, e1.data_ptr(), e2.data_ptr(), None) == 0
    torch.cuda.synchronize()
    r1, r2 = oracle.philox_normals(seed, off, K, T, rover=3, k0=k0 % (1 << 32))
    assert np.array_equal(e1.cpu().numpy().view(np.uint32), r1.ravel().view(np.uint32))
    assert np.array_equal(e2.cpu().numpy().view(np.uint32), r2.ravel().view(np.uint32))


def test_normalize3_is_ieee_exact(L):
    import torch
    rng = np.random.default_rng(1)
    n = 1 << 22
    g = rng.standard_normal((n, 3)).astype(np.float32)
    unit = g / np.linalg.norm(g.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    near = unit * (1 + rng.uniform(-1e-4, 1e-4, (n, 1))).astype(np.float32)     # straddles the 2^-14 shortcut bound
    tiny = unit * np.float32(1e-2) ** 2                                         # normal_on_grid scale (res^2)
    wide = g * np.exp(rng.uniform(-20, 20, (n, 1))).astype(np.float32)
    sparse = unit.copy()
    sparse[:, 2] = 0.0                                                          # planar headings, exact zeros
    for name, v in (("unit", unit), ("near", near), ("tiny", tiny), ("wide", wide), ("sparse", sparse)):
        vd = dev(v)
        out = torch.empty_like(vd)
        ref = torch.empty_like(vd)
        assert L.mppi_test_normalize(vd.data_ptr(), out.data_ptr(), ref.data_ptr(), n, None) == 0
        torch.cuda.synchronize()
        o, r = out.cpu().numpy(), ref.cpu().numpy()
        bad = int((o != r).sum())          # value comparison: +0 / -0 of a zero component is not significant
        assert bad == 0, f"{name}: {bad} of {o.size} components differ from IEEE sqrt/div"


def test_fdiv_fsqrt_are_ieee_exact(L):
    import torch
    rng = np.random.default_rng(2)
    n = 1 << 23
    a = (rng.standard_normal(n) * np.exp(rng.uniform(-25, 25, n))).astype(np.float32)
    b = (rng.standard_normal(n) * np.exp(rng.uniform(-25, 25, n))).astype(np.float32)
    b[b == 0] = 1.0
    a[: n // 64] = 0.0                       # zero numerators / radicands (a rover that does not move)
    ad, bd = dev(a), dev(b)
    out = torch.empty(2 * n, device="cuda")
    ref = torch.empty(2 * n, device="cuda")
    assert L.mppi_test_divsqrt(ad.data_ptr(), bd.data_ptr(), out.data_ptr(), ref.data_ptr(), n, None) == 0
    torch.cuda.synchronize()
    o, r = out.cpu().numpy(), ref.cpu().numpy()
    assert int((o != r).sum()) == 0
