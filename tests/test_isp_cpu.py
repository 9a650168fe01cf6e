"""Oracle for the raw->sRGB branch (oracle/isp_ref.py) against the golden produced by the unmodified reference
util/process.py `process` (tests/golden/isp_kat.npz, make_golden.py isp)."""
import os

import numpy as np

from tests.conftest import REPO


def _kat():
    return np.load(os.path.join(REPO, 'tests', 'golden', 'isp_kat.npz'))


def test_isp_oracle_matches_reference_golden():
    from oracle import isp_ref
    k = _kat()
    y = isp_ref.process(k['x'], k['wb'], k['ccm'], gamma=2.2)
    assert y.shape == k['y'].shape and y.dtype == np.float32
    # outputs are multiples of 1/255: the only admissible difference is a pow() rounding flipping a truncation
    steps = np.rint(np.abs(y - k['y']) * 255.0)
    assert steps.max() <= 1 and (steps > 0).mean() <= 2e-3, (steps.max(), (steps > 0).mean())
    assert np.allclose(np.rint(y * 255.0) / 255.0, y, atol=1e-7)


def test_isp_oracle_crf_branch_structure():
    """CRF branch (paper-side library torchinterp1d absent: parity unpinned): identity response reproduces the clipped
    linear image quantised to 8 bits; a monotone response stays monotone."""
    from oracle import isp_ref
    k = _kat()
    E = np.linspace(0.0, 1.0, 1024, dtype=np.float32)
    ident = np.tile(E, (3, 1))
    y = isp_ref.process(k['x'], k['wb'], k['ccm'], CRF=(E, ident))
    lin = isp_ref.process(k['x'], k['wb'], k['ccm'], gamma=1.0)
    assert np.abs(y - lin).max() <= 1.0 / 255.0 + 1e-6
    sq = np.tile(np.sqrt(E), (3, 1)).astype(np.float32)
    y2 = isp_ref.process(k['x'], k['wb'], k['ccm'], CRF=(np.tile(E, (3, 1)), sq))
    assert (y2 >= lin - 1.0 / 255.0 - 1e-6).all()


def test_crf_branch_against_the_reference_run_with_its_own_yardstick(golden_dir):
    """tests/golden/isp_crf_kat.npz = the UNMODIFIED util/process.py CRF branch on the reference's EMoR curves, with the
    absent third-party torchinterp1d replaced by scipy.interpolate.interp1d - what the reference's own EMoR/test_EMoR.py
    compares torchinterp1d against.  The oracle restates torchinterp1d's published formula (slope = dy / (eps + dx)): the
    two agree to the 8-bit level everywhere except at exactly saturated pixels, where the eps in the slope leaves the
    restated formula 6e-8 under 1.0 and the reference's `.int()` (process.py:82) truncates 254.99998 to 254."""
    import os
    import numpy as np
    from oracle import isp_ref
    k = np.load(os.path.join(golden_dir, 'isp_crf_kat.npz'))
    y = isp_ref.process(k['x'], k['wb'], k['ccm'], CRF=(k['E'], k['fs']))
    steps = np.rint(np.abs(y - k['y']) * 255.0)
    unsat = k['y'] < 1.0
    assert unsat.mean() > 0.5 and steps[unsat].max() <= 1 and (steps[unsat] > 0).mean() <= 5e-3
    assert steps[~unsat].max() <= 1                                   # the documented one-level truncation, nothing else
