#!/usr/bin/env python3
"""Writes tests/golden/ref_api_surface.json: the public methods (name + positional argument names) of every class of the
reference's Python package `drc` (/root/reference/drc/**/*.py), read with `ast` -- nothing is imported or executed.
tests/test_api_surface_cpu.py compares the mirror package dyros_robot_controller_b200.drc with this list, so the fixture travels to
machines that do not have the reference tree."""
import ast
import json
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/drc")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "ref_api_surface.json"


def main():
    api = {}
    for f in sorted(REF.rglob("*.py")):
        mod = ".".join(f.relative_to(REF).with_suffix("").parts)
        tree = ast.parse(f.read_text())
        for node in tree.body:
            if isinstance(node, ast.ClassDef):
                methods = {}
                for item in node.body:
                    if isinstance(item, ast.FunctionDef) and (not item.name.startswith("_") or item.name == "__init__"):
                        methods[item.name] = [a.arg for a in item.args.args if a.arg != "self"]
                api[f"{mod}.{node.name}"] = methods
    OUT.write_text(json.dumps(api, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT, {k: len(v) for k, v in api.items()})


def bindings_surface(src=Path("/root/reference/src/bindings.cpp")):
    """module name, classes and their .def / .def_readwrite / enum .value names, read from the text of src/bindings.cpp"""
    import re
    text = src.read_text()
    mod = re.search(r"BOOST_PYTHON_MODULE\((\w+)\)", text).group(1)
    classes, cur = {}, None
    for line in text.split("\n"):
        m = re.search(r'bp::(?:class_|enum_)<[^"]*>\s*\(\s*"(\w+)"', line)
        if m:
            cur = m.group(1)
            classes[cur] = []
            continue
        m = re.search(r'\.(?:def|def_readwrite|value)\(\s*"(\w+)"', line)
        if m and cur and m.group(1) not in classes[cur]:
            classes[cur].append(m.group(1))
    out = OUT.parent / "ref_bindings_surface.json"
    out.write_text(json.dumps({"module": mod, "classes": classes}, indent=1, sort_keys=True) + "\n")
    print("wrote", out, {k: len(v) for k, v in classes.items()})


if __name__ == "__main__":
    main()
    bindings_surface()
