"""Error types of the reference kept under the same names (safe_adaptation_gym/utils.py:6-8)."""


class ResamplingError(AssertionError):
    """Raised when we fail to sample a valid distribution of objects or goals"""
    pass
