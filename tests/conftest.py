import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "mfcc-rust_b200", ROOT / "oracle", ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import speechsauce_b200 as sb
        return sb.lib.ss_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- tolerances from BASELINE.json north_star ------------------------------------------------
RTOL_SPEC = 1e-4   # rel 1e-4 on power and mel spectra
ATOL_LOG = 1e-3    # abs 1e-3 on log-mel and MFCC coefficients


def assert_spec_close(got, want, rtol=RTOL_SPEC, atol_frac=2e-6, what=""):
    """|got - want| <= rtol*|want| + atol_frac*max|want|.  The absolute floor covers bins that are
    small only through cancellation inside the f32 FFT (both the oracle and the kernel carry an
    error of ~1e-7 * max|X| there, so a purely relative test would compare noise with noise)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    tol = rtol * np.abs(want) + atol_frac * (np.abs(want).max() if want.size else 0.0)
    err = np.abs(got - want)
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond tolerance, worst err {err.max():.3e} (rel {np.max(err / (np.abs(want) + 1e-30)):.3e})"


def assert_log_close(got, want, atol=ATOL_LOG, what=""):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    err = np.abs(got - want)
    assert err.max(initial=0.0) <= atol, f"{what}: worst abs err {err.max():.3e} > {atol}"
