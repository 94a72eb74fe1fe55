from .solver import ClawSolver, ClawSolver2D, ClawSolver3D  # noqa: F401
