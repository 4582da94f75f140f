"""Shared helpers for the parity tests."""
import numpy as np

# north_star tolerances (BASELINE.json): float stage outputs within 1e-4 relative RMS of the
# reference chain, final s16 audio within +-1 LSB.
REL_RMS_TOL = 1e-4
PCM_TOL_LSB = 1


def rel_rms(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.sqrt(np.mean(np.abs(b.astype(np.complex128 if np.iscomplexobj(b) else np.float64)) ** 2))
    num = np.sqrt(np.mean(np.abs(a.astype(np.complex128 if np.iscomplexobj(a) else np.float64) - b) ** 2))
    return float(num / den) if den > 0 else float(num)


def active_channels(spec_carriers):
    """0-based channel indices that carry a signal (discriminator parity is only meaningful there)."""
    return sorted({c.channel - 1 for c in spec_carriers})
