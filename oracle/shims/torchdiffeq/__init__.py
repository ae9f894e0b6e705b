"""Stand-in for torchdiffeq 0.2.1: the solvers the reference's hot path selects (euler, midpoint, rk4 = 3/8 rule,
dopri5) restated from the published algorithms, with the `_impl` class layout src/early_stop_solver.py subclasses."""
from ._impl.odeint import odeint, odeint_adjoint

__version__ = '0.2.1'
