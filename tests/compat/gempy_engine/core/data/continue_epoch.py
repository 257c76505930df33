class ContinueEpoch(Exception):
    """Control-flow exception of the nugget optimiser (gempy/modules/optimize_nuggets/_ops.py:24)."""
