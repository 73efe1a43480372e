#!/usr/bin/env python
"""Minimal static check (no linter is installed in the image): unused imports and names that are
read but never bound anywhere in the module or in builtins.  `python tools/lint.py [paths]`."""
import ast
import builtins
import os
import sys


def check(path):
    src = open(path).read()
    tree = ast.parse(src, path)
    imported, bound, used = {}, set(dir(builtins)), set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                imported[(a.asname or a.name).split(".")[0]] = node.lineno
        elif isinstance(node, ast.ImportFrom):
            for a in node.names:
                imported[a.asname or a.name] = node.lineno
        elif isinstance(node, ast.Name):
            (used if isinstance(node.ctx, ast.Load) else bound).add(node.id)
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(node.name)
            if not isinstance(node, ast.ClassDef):
                a = node.args
                for arg in a.posonlyargs + a.args + a.kwonlyargs + [a.vararg, a.kwarg]:
                    if arg is not None:
                        bound.add(arg.arg)
        elif isinstance(node, ast.Lambda):
            a = node.args
            for arg in a.posonlyargs + a.args + a.kwonlyargs + [a.vararg, a.kwarg]:
                if arg is not None:
                    bound.add(arg.arg)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            bound.add(node.name)
        elif isinstance(node, (ast.Global, ast.Nonlocal)):
            bound.update(node.names)
        elif isinstance(node, ast.alias):
            pass
    exported = set()
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
            exported = {e.value for e in getattr(node.value, "elts", []) if isinstance(e, ast.Constant)}
    out = []
    is_init = os.path.basename(path) == "__init__.py"
    for name, line in sorted(imported.items(), key=lambda kv: kv[1]):
        if name not in used and name not in exported and not is_init and name not in ("*", "annotations") \
                and "noqa" not in src.splitlines()[line - 1]:
            out.append("%s:%d: unused import %s" % (path, line, name))
    for name in sorted(used - bound - set(imported) - {"__file__", "__name__", "__doc__", "__spec__"}):
        out.append("%s: undefined name %s" % (path, name))
    return out


def main(argv):
    roots = argv or ["ddlbench_b200", "bench.py", "__graft_entry__.py", "tools", "bench", "benchmark", "pipedream-fork"]
    problems = []
    for root in roots:
        files = [root] if root.endswith(".py") else [os.path.join(d, f) for d, _, fs in os.walk(root) for f in fs
                                                     if f.endswith(".py")]
        for f in sorted(files):
            problems += check(f)
    print("\n".join(problems) or "clean")
    return 1 if problems else 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
