"""Poor man's undefined-name check (no linter in the image): run before spending GPU minutes."""
import ast
import builtins
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bad = 0
for f in glob.glob(os.path.join(ROOT, "torchctr_b200", "**", "*.py"), recursive=True) + [os.path.join(ROOT, "bench.py"),
                                                                                         os.path.join(ROOT, "__graft_entry__.py")]:
    tree = ast.parse(open(f).read())
    defined = set(dir(builtins)) | {"__file__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            defined.update((a.asname or a.name).split(".")[0] for a in n.names)
        elif isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            defined.add(n.name)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            defined.add(n.id)
        elif isinstance(n, ast.arg):
            defined.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            defined.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            defined.update(n.names)
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in defined:
            print(f"{f}:{n.lineno}: undefined name {n.id}")
            bad += 1
sys.exit(1 if bad else 0)
