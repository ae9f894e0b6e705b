"""Record the reference CLI's flags and defaults (src/run_GNN.py:283-436) without importing it: the argparse calls are
read from the source with `ast`.  Output: tests/golden/cli_defaults.json, checked by tests/test_host_logic.py against
graph-neural-pde_b200/run_GNN.py's parser.   python tests/golden/make_cli_defaults.py   (build container only)"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/src/run_GNN.py'


def literal(node):
  try:
    return ast.literal_eval(node)
  except Exception:
    return '<expr>'


def main():
  tree = ast.parse(open(SRC).read())
  flags = {}
  for node in ast.walk(tree):
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == 'add_argument':
      if not node.args or not isinstance(node.args[0], ast.Constant) or not str(node.args[0].value).startswith('--'):
        continue
      name = node.args[0].value[2:]
      kw = {k.arg: k.value for k in node.keywords}
      entry = {}
      if 'action' in kw:
        action = literal(kw['action'])
        entry['action'] = action
        entry['default'] = {'store_true': False, 'store_false': True}.get(action)
      if 'default' in kw:
        entry['default'] = literal(kw['default'])
      elif 'action' not in kw:
        entry['default'] = None
      if 'type' in kw and isinstance(kw['type'], ast.Name):
        entry['type'] = kw['type'].id
      if 'dest' in kw:
        entry['dest'] = literal(kw['dest'])
      flags[name] = entry
  out = os.path.join(HERE, 'cli_defaults.json')
  json.dump(flags, open(out, 'w'), indent=1, sort_keys=True)
  print('wrote %d flags -> %s' % (len(flags), out))
  overlay_cases(tree)


OVERLAY_ARGV = [
  ['--dataset', 'ogbn-arxiv'], ['--dataset', 'ogbn-arxiv', '--block', 'hard_attention'],
  ['--dataset', 'Cora', '--time', '3', '--method', 'rk4', '--step_size', '0.5'],
  ['--dataset', 'Photo', '--attention_type', 'cosine_sim', '--epoch', '7', '--beltrami'],
  ['--dataset', 'Pubmed', '--function', 'transformer', '--not_lcc'], ['--dataset', 'Nowhere'],
]


def overlay_cases(tree):
  """Run the reference's OWN merge_cmd_args (exec'd from its source, it is pure dict logic) and best_params_dict over a
  few command lines; tests/test_host_logic.py holds run_GNN.reference_best_params to the recorded results."""
  import sys
  sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))
  from __graft_entry__ import load_package
  load_package()
  import importlib
  parser = importlib.import_module('graph_neural_pde_b200.run_GNN').build_parser()   # defaults == reference (above)
  ns = {}
  for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name == 'merge_cmd_args':
      exec(compile(ast.Module(body=[node], type_ignores=[]), SRC, 'exec'), ns)
  btree = ast.parse(open('/root/reference/src/best_params.py').read())
  table = [ast.literal_eval(n.value) for n in btree.body
           if isinstance(n, ast.Assign) and getattr(n.targets[0], 'id', '') == 'best_params_dict'][0]
  cases = []
  for argv in OVERLAY_ARGV:
    cmd_opt = vars(parser.parse_args(argv))
    cmd_opt.pop('best_params'), cmd_opt.pop('quiet')
    try:                                   # reference src/run_GNN.py:216-221
      opt = {**cmd_opt, **table[cmd_opt['dataset']]}
      ns['merge_cmd_args'](cmd_opt, opt)
    except KeyError:
      opt = cmd_opt
    cases.append({'argv': argv, 'opt': opt})
  json.dump(cases, open(os.path.join(HERE, 'cli_overlay_cases.json'), 'w'), indent=1, sort_keys=True)
  print('wrote %d overlay cases' % len(cases))


if __name__ == '__main__':
  main()
