"""The oracle build's stand-ins for GSL's tail probabilities (oracle/gsl_stub/gsl_stub.c: what the reference's PeakFinder computes
its p-values with in oracle/_ref/genomic_scans) against scipy -- the checker of the peaks driver has to be right itself."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import support

scipy_stats = pytest.importorskip("scipy.stats")


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    src = os.path.join(support.ROOT, "oracle", "gsl_stub")
    so = tmp_path_factory.mktemp("gsl") / "libgslstub.so"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I" + src, os.path.join(src, "gsl_stub.c"), "-o", str(so), "-lm"])
    lib = ctypes.CDLL(str(so))
    lib.gsl_cdf_binomial_Q.restype = ctypes.c_double
    lib.gsl_cdf_binomial_Q.argtypes = [ctypes.c_uint, ctypes.c_double, ctypes.c_uint]
    lib.gsl_cdf_poisson_Q.restype = ctypes.c_double
    lib.gsl_cdf_poisson_Q.argtypes = [ctypes.c_uint, ctypes.c_double]
    lib.gsl_cdf_ugaussian_Q.restype = ctypes.c_double
    lib.gsl_cdf_ugaussian_Q.argtypes = [ctypes.c_double]
    return lib


def test_binomial_tail(stub):
    rng = np.random.default_rng(1)
    for _ in range(400):
        n = int(rng.choice([201, 501, 1001, 50_000, 3_000_000]))
        p = float(rng.choice([1e-4, 0.01, 0.114, 0.5, 0.93]))
        k = int(min(n - 1, max(0, rng.normal(n * p, 4 * np.sqrt(n * p * (1 - p)) + 3))))
        want = scipy_stats.binom.sf(k, n, p)
        got = stub.gsl_cdf_binomial_Q(k, p, n)
        assert got == pytest.approx(want, rel=1e-9, abs=1e-300), (k, n, p)
    assert stub.gsl_cdf_binomial_Q(501, 0.3, 501) == 0.0 and stub.gsl_cdf_binomial_Q(0, 0.0, 10) == 0.0


def test_poisson_and_normal_tails(stub):
    rng = np.random.default_rng(2)
    for _ in range(300):
        mu = float(rng.choice([5.0, 6.0, 17.0, 120.0, 505.0]))
        k = int(max(0, rng.normal(mu, 5 * np.sqrt(mu))))
        assert stub.gsl_cdf_poisson_Q(k, mu) == pytest.approx(scipy_stats.poisson.sf(k, mu), rel=1e-9, abs=1e-300), (k, mu)
    for x in np.linspace(-8, 30, 77):
        assert stub.gsl_cdf_ugaussian_Q(float(x)) == pytest.approx(scipy_stats.norm.sf(x), rel=1e-10, abs=1e-300)
