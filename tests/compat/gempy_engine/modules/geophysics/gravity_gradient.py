from gempy_b200.engine.geophysics import calculate_gravity_gradient   # noqa: F401
