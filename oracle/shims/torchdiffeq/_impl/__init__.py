from .odeint import odeint, odeint_adjoint
