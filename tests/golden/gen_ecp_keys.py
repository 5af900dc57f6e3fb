"""Writes tests/golden/ecp_keys.json: the JSON keys, in order, of the dict that each of the reference's three
bbox_to_ecp_format functions returns (parsed from the reference sources with ast - TensorFlow is not needed).
Run here (the reference tree is not present on the GPU box):  python tests/golden/gen_ecp_keys.py"""
import ast
import json
import os

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ecp_keys.json')
FILES = {'standard': 'inference_standard_yolov3.py', 'aleatoric': 'inference_aleatoric.py', 'epistemic': 'inference_epistemic.py'}

keys = {}
for variant, fname in FILES.items():
    tree = ast.parse(open(os.path.join(REF, fname)).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == 'bbox_to_ecp_format')
    ret = next(n for n in ast.walk(fn) if isinstance(n, ast.Return) and isinstance(n.value, ast.Dict))
    keys[variant] = [k.value for k in ret.value.keys]
json.dump(keys, open(OUT, 'w'), indent=1)
print(keys)
