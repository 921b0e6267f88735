"""Severity levels and their parameter rows (surface of reference shared/disturbance_types.py:8-43).

`SEVERITY_CONFIGS[DisturbanceSeverity.X]` is a dict with the reference's four keys.  NONE has no
row, as in the reference.
"""
from enum import Enum


class DisturbanceSeverity(Enum):
    NONE = "NONE"
    MILD = "MILD"
    MODERATE = "MODERATE"
    HARD = "HARD"
    SEVERE = "SEVERE"


_KEYS = ("gaussian_noise_sigma", "gaussian_blur_sigma", "contrast_range", "cutout_ratio")

#            level      noise-sigma  blur-sigma  contrast (lo, hi)  cutout area ratio
_ROWS = (
    ("MILD",     0.08, 1.0, (0.75, 1.25), 0.10),
    ("MODERATE", 0.12, 2.0, (0.7, 1.3),   0.17),
    ("HARD",     0.13, 2.1, (0.69, 1.31), 0.18),
    ("SEVERE",   0.26, 3.0, (0.6, 1.4),   0.25),
)

SEVERITY_CONFIGS = {DisturbanceSeverity[name]: dict(zip(_KEYS, row)) for name, *row in _ROWS}
