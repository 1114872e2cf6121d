from .linear_solvers import CG, Jacobi, Solver  # noqa: F401
