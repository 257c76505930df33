"""gempy_b200 -- B200-native backend for GemPy's implicit co-kriging hot path."""
__version__ = "0.1.0"
