from ._types import InterpolationOptions, EvaluationOptions, KernelOptions                # noqa: F401
