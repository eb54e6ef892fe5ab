"""Writes tests/golden/reference_model_blocks.json: the `model:` block of the reference's own training recipes
(`/root/reference/yamls/hydra-yamls/SD-2-base-256.yaml:14-23`, `SD-2-base-512.yaml:20-29`) parsed with PyYAML, so that the
drop-in test (tests/test_host_logic.py::test_factory_accepts_the_reference_yaml_model_block) can call the factory with
exactly the keyword arguments `composer run.py` / Hydra would pass, on machines where /root/reference does not exist.
Run from the repo root:  python tests/golden/make_yaml_fixture.py
"""
import json
import os

import yaml

REF = '/root/reference/yamls/hydra-yamls'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_model_blocks.json')


def blocks():
    out = {}
    for name in ('SD-2-base-256.yaml', 'SD-2-base-512.yaml'):
        with open(os.path.join(REF, name)) as f:
            out[name] = yaml.safe_load(f)['model']
    return out


if __name__ == '__main__':
    with open(OUT, 'w') as f:
        json.dump(blocks(), f, indent=1, sort_keys=True)
    print(open(OUT).read())
