"""grand_b200 -- B200-native ODE right-hand-side engine behind the reference's ODEFunc / ODEblock API.

The module names mirror the reference's flat `src/` layout (function_transformer_attention, block_constant, ...)
so a maintainer can point the reference's imports at this package.  The directory name contains a
hyphen, so import it through `__graft_entry__.load_package()` which registers it as `graph_neural_pde_b200`.
"""
__version__ = '0.1.0'
