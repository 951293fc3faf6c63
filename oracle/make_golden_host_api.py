"""oracle/make_golden_host_api.py -- TEST INFRASTRUCTURE ONLY.

Runs oracle/host_api_cases.py against the UNMODIFIED reference modules (``det3d.utils.registry``,
``det3d.torchie.fileio``, and the argument parser of tools/active_select.py; imported from /root/reference with the shims of reference_loader.py) and records the
outcomes -> tests/golden/host_api.json.  tests/test_host.py replays the same table on this package.

    python oracle/make_golden_host_api.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import host_api_cases  # noqa: E402
from reference_loader import load_reference_selectors  # noqa: E402


def reference_cli_parser():
    """``parse_args`` of the reference's tools/active_select.py (:22-63), lifted unmodified out of the file with
    ``ast`` (the module itself imports the detector stack); it reads ``sys.argv``."""
    import ast
    from reference_loader import REFERENCE_ROOT
    path = os.path.join(REFERENCE_ROOT, "tools", "active_select.py")
    fn = [n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name == "parse_args"]
    ns = {"os": os, "argparse": __import__("argparse")}
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)

    def parse(argv):
        saved = sys.argv
        sys.argv = ["active_select.py"] + argv
        try:
            return ns["parse_args"]()
        finally:
            sys.argv = saved
    return parse


def main():
    load_reference_selectors()                       # installs the shims and puts the reference on sys.path
    from det3d import torchie
    from det3d.utils.registry import Registry, build_from_cfg
    api = {"Registry": Registry, "build_from_cfg": build_from_cfg, "load": torchie.load, "dump": torchie.dump}
    with tempfile.TemporaryDirectory() as wd:
        outcomes = host_api_cases.run_all(api, wd)
    outcomes["cli"] = host_api_cases.cli_cases(reference_cli_parser())
    import numpy
    import yaml
    golden = {"source": "det3d/utils/registry.py, det3d/torchie/fileio", "python": sys.version.split()[0],
              "numpy": numpy.__version__, "yaml": yaml.__version__, "outcomes": outcomes}
    out_file = os.path.join(ROOT, "tests", "golden", "host_api.json")
    with open(out_file, "w") as f:
        json.dump(golden, f, indent=1)
    for group, cases in outcomes.items():
        for k, v in cases.items():
            print(group, k, json.dumps(v)[:150])
    print("wrote", out_file)


if __name__ == "__main__":
    main()
