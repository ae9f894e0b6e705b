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


if __name__ == '__main__':
  main()
