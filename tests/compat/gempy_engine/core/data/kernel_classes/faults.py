from .._types import FaultsData, FiniteFaultData                  # noqa: F401
