"""TEST INFRASTRUCTURE - CPU restatement (numpy, float32) of the reference's raw->sRGB rendering.

Follows util/process.py line by line: apply_gains :15-19, clamp :56, binning :41-48, apply_ccms :22-31, clamp :61,
gamma_compression :34-39 (pinned by tests/golden/isp_kat.npz, produced by the UNMODIFIED reference `process`), and
camera_response_function :71-84, whose interpolation lives in the third-party `torchinterp1d` package (absent from the
reference tree, no version pinned anywhere in it - README.md:32-34 lists names only): its published algorithm
(searchsorted - 1, clamp, y0 + slope*(x - x0), slope = dy / (eps + dx)) is restated here and PINNED by
tests/golden/isp_crf_kat.npz: the unmodified reference CRF branch run on its own EMoR curves with torchinterp1d replaced by
scipy.interpolate.interp1d, the routine the reference's own EMoR/test_EMoR.py:44-47,62-75 holds torchinterp1d against
(agreement to the 8-bit level on every unsaturated pixel; at exactly saturated pixels the eps of the restated slope leaves
254.99998, which process.py:82's .int() truncates - documented in tests/test_isp_cpu.py).  Only tests/ may import this module.
"""
import numpy as np

F = np.float32


def process(bayer, wbs, ccms, gamma=2.2, CRF=None):
    """bayer [N,4,h,w] f32, wbs [N,4], ccms [N,3,3] -> [N,3,h,w] f32."""
    x = bayer.astype(F) * wbs.astype(F)[:, :, None, None]                      # :15-19
    x = np.clip(x, F(0), F(1))                                                  # :56
    rgb = np.stack([x[:, 0], (x[:, 1] + x[:, 3]) * F(0.5), x[:, 2]], axis=1)     # :41-48 (mean of two = sum * 0.5)
    c = ccms.astype(F)
    out = np.empty_like(rgb)
    for k in range(3):
        # :22-31 torch.sum(images * ccms, dim=-1): fp32 products; torch's CPU reduction of the 3 terms is reproduced
        # exactly by a double-precision accumulator rounded once (checked against the golden, saturated pixels included)
        p = [rgb[:, i] * c[:, k, i, None, None] for i in range(3)]
        out[:, k] = (p[0].astype(np.float64) + p[1] + p[2]).astype(F)
    out = np.clip(out, F(0), F(1))                                              # :61
    if CRF is None:
        out = np.power(np.maximum(out, F(1e-8)), F(1.0 / gamma))                 # :34-36
    else:
        E, fs = CRF
        E = np.asarray(E, F)
        E = E[0] if E.ndim == 2 else E
        fs = np.asarray(fs, F)
        res = np.empty_like(out)
        eps = np.finfo(F).eps
        for k in range(3):                                                      # :71-81 via torchinterp1d.Interp1d
            v = out[:, k]
            ind = np.clip(np.searchsorted(E, v, side='left') - 1, 0, E.shape[0] - 2)
            slope = (fs[k, 1:] - fs[k, :-1]) / (eps + (E[1:] - E[:-1]))
            res[:, k] = fs[k][ind] + slope[ind] * (v - E[ind])
        out = res
    q = np.clip((out * F(255)).astype(np.int32), 0, 255).astype(F) / F(255)      # :38 / :83 (.int() truncates)
    return np.clip(q, F(0), F(1))
