"""The reference's public API surface, read from its source with `ast` (nothing is imported or executed):
every top-level function and every class with its methods, each with its positional parameter names and the repr of its
defaults.  Run HERE (needs /root/reference); tests/test_api_surface.py compares the package against the committed JSON.

    python tests/golden/make_api_surface.py
"""
import ast
import json
import os

FILES = {"myolo.model": "/root/reference/myolo/model.py", "myolo.myolo_utils": "/root/reference/myolo/myolo_utils.py",
         "myolo.config": "/root/reference/myolo/config.py"}
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_api_surface.json")


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [ast.unparse(d) for d in a.defaults]
    return {"args": names, "defaults": defaults, "vararg": a.vararg.arg if a.vararg else None,
            "kwarg": a.kwarg.arg if a.kwarg else None, "line": fn.lineno}


def main():
    out = {}
    for mod, path in FILES.items():
        tree = ast.parse(open(path).read())
        m = {"functions": {}, "classes": {}}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef):
                m["functions"][node.name] = sig(node)
            elif isinstance(node, ast.ClassDef):
                m["classes"][node.name] = {"bases": [ast.unparse(b) for b in node.bases], "line": node.lineno,
                                           "methods": {n.name: sig(n) for n in node.body if isinstance(n, ast.FunctionDef)}}
        out[mod] = m
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, {k: (len(v["functions"]), len(v["classes"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
