"""Derive graph-neural-pde_b200/best_params.json from the reference's tuned hyper-parameter table
(src/best_params.py, read with `ast`, not imported): per dataset only the keys whose value differs from the CLI default
of run_GNN.py.  `run_GNN.py --best_params reference` overlays them the way the reference's main() does
(src/run_GNN.py:188-221).   python tools/make_best_params.py   (build container only)"""
import ast
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)


def main():
  from __graft_entry__ import load_package
  load_package()
  import importlib
  parser = importlib.import_module('graph_neural_pde_b200.run_GNN').build_parser()
  defaults = vars(parser.parse_args([]))
  tree = ast.parse(open('/root/reference/src/best_params.py').read())
  table = None
  for node in tree.body:
    if isinstance(node, ast.Assign) and getattr(node.targets[0], 'id', '') == 'best_params_dict':
      table = ast.literal_eval(node.value)
  assert table is not None
  out = {}
  for name, params in table.items():
    out[name] = {k: v for k, v in sorted(params.items()) if k not in defaults or defaults[k] != v}
  path = os.path.join(ROOT, 'graph-neural-pde_b200', 'best_params.json')
  json.dump(out, open(path, 'w'), sort_keys=True, separators=(',', ':'))
  print('wrote %s: %s' % (path, {k: len(v) for k, v in out.items()}))


if __name__ == '__main__':
  main()
