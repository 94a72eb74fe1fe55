from .solver import ClawSolver2D, ClawSolver3D  # noqa: F401
