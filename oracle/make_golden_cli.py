"""Golden description of the reference's predict.py command line (flags, defaults, types), read from its SOURCE with
ast -- the script itself cannot be imported here (torch_geometric is missing).  Run in the build container:
    python oracle/make_golden_cli.py   ->   tests/golden/predict_flags.json
TEST INFRASTRUCTURE: nothing in the product imports this."""
import ast
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/pointstowood/predict.py"


def flags(path):
    out = {}
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            names = [a.value for a in node.args if isinstance(a, ast.Constant)]
            kw = {}
            for k in node.keywords:
                if k.arg in ("default", "nargs", "action"):
                    kw[k.arg] = ast.literal_eval(k.value)
                elif k.arg == "type":
                    kw["type"] = k.value.id
            out[max(names, key=len)] = dict(names=sorted(names), **kw)
    return out


if __name__ == "__main__":
    f = flags(SRC)
    with open(os.path.join(ROOT, "tests", "golden", "predict_flags.json"), "w") as fh:
        json.dump(f, fh, indent=1, sort_keys=True)
    print(len(f), "flags:", ", ".join(sorted(f)))
