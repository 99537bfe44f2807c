"""CPU: the oracle's building blocks -- Philox4x32-10 against the Random123 known-answer vectors and the
specified ("det") transcendental functions against float64 libm."""
import numpy as np


def ulp_err(y, ref64):
    u = np.spacing(np.abs(ref64).astype(np.float32)).astype(np.float64)
    return float(np.max(np.abs(y.astype(np.float64) - ref64) / u))


def test_philox4x32_10_random123_kat(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, exp in kats:
        assert list(oracle.philox4x32_10(ctr, key)) == exp


def test_det_sincos_accuracy(oracle):
    x = np.linspace(-0.2, 0.2, 200001).astype(np.float32)          # the rollout's range: |w*dt| <= 0.045
    s, c = oracle.detmath(0, x)
    assert ulp_err(s, np.sin(x.astype(np.float64))) <= 1.0
    assert ulp_err(c, np.cos(x.astype(np.float64))) <= 1.5
    x = np.linspace(-100, 100, 1000001).astype(np.float32)
    s, c = oracle.detmath(0, x)
    assert ulp_err(s, np.sin(x.astype(np.float64))) <= 2.0
    assert ulp_err(c, np.cos(x.astype(np.float64))) <= 2.0


def test_det_sincos2pi_accuracy(oracle):
    u = np.arange(0, 1 << 24, 7).astype(np.float32) * np.float32(2.0 ** -24)
    s, c = oracle.detmath(1, u)
    a = 2 * np.pi * u.astype(np.float64)
    assert np.max(np.abs(s - np.sin(a))) < 2e-7 and np.max(np.abs(c - np.cos(a))) < 2e-7


def test_det_log_exp_accuracy(oracle):
    x = np.arange(1, 1 << 24, 5).astype(np.float32) * np.float32(2.0 ** -24)
    lg, _ = oracle.detmath(2, x)
    assert ulp_err(lg, np.log(x.astype(np.float64))) <= 1.0
    x = np.linspace(-87, 5, 1000001).astype(np.float32)
    e, _ = oracle.detmath(3, x)
    assert ulp_err(e, np.exp(x.astype(np.float64))) <= 1.5
    e, _ = oracle.detmath(3, np.array([-87.5, -1e4, 0.0, -0.0], np.float32))
    assert list(e) == [0.0, 0.0, 1.0, 1.0]          # flush below the normal range, by specification


def test_det_atan_accuracy(oracle):
    rng = np.random.default_rng(3)
    x = np.concatenate([np.linspace(-4, 4, 400001), rng.standard_cauchy(400000) * 10,
                        np.array([0.0, -0.0, np.inf, -np.inf, 1e30, -1e-30])]).astype(np.float32)
    a, _ = oracle.detmath(4, x)
    assert ulp_err(a, np.arctan(x.astype(np.float64))) <= 3.0      # Cephes atanf: ~2.7 ulp just above tan(pi/8)
    assert np.array_equal(np.signbit(a[x == 0]), np.signbit(x[x == 0])) and a[np.isposinf(x)][0] == np.float32(np.pi / 2)
    n, _ = oracle.detmath(4, np.array([np.nan], np.float32))
    assert np.isnan(n[0])


def test_philox_normals_statistics_and_replay(oracle):
    e1, e2 = oracle.philox_normals(42, 3, 4096, 100)
    n = e1.size
    for e in (e1, e2):
        assert abs(e.mean()) < 5 / np.sqrt(n) and abs(e.std() - 1) < 5 / np.sqrt(2 * n)
        assert abs(np.mean(e.astype(np.float64) ** 3)) < 0.02 and abs(np.mean(e.astype(np.float64) ** 4) - 3) < 0.05
    # channels independent (the reference's two channels are shifted copies, SURVEY A.1 -- not reproduced)
    assert abs(np.corrcoef(e1.ravel(), e2.ravel())[0, 1]) < 0.01
    assert abs(np.corrcoef(e1[:-1].ravel(), e2[1:].ravel())[0, 1]) < 0.01
    # successive steps of one sample are independent too (cos / sin halves of one Box-Muller pair)
    assert abs(np.corrcoef(e1[:, 0::2].ravel(), e1[:, 1::2].ravel())[0, 1]) < 0.01
    # exact replay, sharding invariance (global sample id keys the counter), different offsets differ
    a1, a2 = oracle.philox_normals(42, 3, 4096, 100)
    assert np.array_equal(a1, e1) and np.array_equal(a2, e2)
    b1, _ = oracle.philox_normals(42, 3, 1000, 100, k0=2000)
    assert np.array_equal(b1, e1[2000:3000])
    c1, _ = oracle.philox_normals(42, 4, 64, 100)
    assert not np.array_equal(c1, e1[:64])
    # odd horizon: the last pair is half used
    d1, d2 = oracle.philox_normals(42, 3, 8, 7)
    assert np.array_equal(d1, e1[:8, :7]) and np.array_equal(d2, e2[:8, :7])
    # libm Box-Muller agrees with the det one to rounding
    l1, _ = oracle.philox_normals(42, 3, 512, 100, math=oracle.MATH_LIBM)
    assert np.max(np.abs(l1 - e1[:512])) < 5e-6


def test_near_unit_normalisation_shortcut_is_ieee_exact(tmp_path):
    """The STRICT kernels skip both MUFU ops when re-normalising an almost-unit vector (csrc/mppi_device.cuh,
    normalize3).  tests/arith_near_unit.c proves by exhaustive enumeration that the shortcut equals IEEE sqrt and
    division for every squared norm in the window, every resulting divisor and every numerator mantissa."""
    import os
    import subprocess
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arith_near_unit.c")
    exe = str(tmp_path / "arith_near_unit")
    subprocess.run(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-mfma", src, "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    bad_s, bad_q, n = (int(x) for x in r.stdout.split())
    assert r.returncode == 0 and bad_s == 0 and bad_q == 0 and n > 3e9


def test_round_toward_zero_add_truncates_exactly(tmp_path):
    """The pipelined kernel indexes its shared-memory DEM tile with one round-toward-zero FADD (+-2^23) instead of a
    float->int conversion (csrc/mppi_device.cuh, tile_rel).  tests/arith_rz_trunc.c proves, for EVERY binary32 value in
    [0, 32768] and its negative, that the integer in the mantissa equals the C truncation."""
    import os
    import subprocess
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arith_rz_trunc.c")
    exe = str(tmp_path / "arith_rz_trunc")
    subprocess.run(["/usr/bin/gcc", "-O2", "-frounding-math", src, "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    bad, n = (int(x) for x in r.stdout.split())
    assert r.returncode == 0 and bad == 0 and n > 1.19e9
