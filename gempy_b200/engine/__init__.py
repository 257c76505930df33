from .data import *  # noqa: F401,F403
