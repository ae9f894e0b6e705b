"""String -> class registry, the reference's plugin boundary (reference src/model_configurations.py:10-44).  The
hot-path blocks / functions are built; the reference's other registry entries are named and refused."""
from .block_constant import ConstantODEblock
from .block_transformer_attention import AttODEblock
from .block_transformer_hard_attention import HardAttODEblock
from .function_GAT_attention import ODEFuncAtt
from .function_laplacian_diffusion import LaplacianODEFunc
from .function_transformer_attention import ODEFuncTransformerAtt


class BlockNotDefined(Exception):
  pass


class FunctionNotDefined(Exception):
  pass


_OUT_OF_SCOPE_BLOCKS = ('mixed', 'rewire_attention')
_OUT_OF_SCOPE_FUNCTIONS = ()


def set_block(opt):
  ode_str = opt['block']
  if ode_str == 'attention':
    return AttODEblock
  if ode_str == 'constant':
    return ConstantODEblock
  if ode_str == 'hard_attention':
    return HardAttODEblock
  if ode_str in _OUT_OF_SCOPE_BLOCKS:
    raise NotImplementedError('block %r is outside the grand_b200 hot-path scope (SURVEY.md section 8)' % ode_str)
  raise BlockNotDefined


def set_function(opt):
  ode_str = opt['function']
  if ode_str == 'laplacian':
    return LaplacianODEFunc
  if ode_str == 'transformer':
    return ODEFuncTransformerAtt
  if ode_str == 'GAT':
    return ODEFuncAtt
  if ode_str in _OUT_OF_SCOPE_FUNCTIONS:
    raise NotImplementedError('function %r is outside the grand_b200 hot-path scope (SURVEY.md section 8)' % ode_str)
  raise FunctionNotDefined
