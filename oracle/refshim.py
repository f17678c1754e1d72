"""
refshim -- run the UNMODIFIED reference sources under Python 3 with Python 2 semantics.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import this module.

The reference (scienceguyrob/PulsarFeatureExtractor) is Python-2.7-only source
(print statements, classic int/int division, ``from scipy import std``, a hard
matplotlib import).  It cannot be imported by the Python 3.12 interpreter in this
image.  This loader reads the reference ``.py`` files *where they lie* under
``/root/reference`` (nothing is copied into the repo), applies a purely mechanical,
in-memory source transformation and executes the result:

  text level   ``print x, y``  -> ``print(x, y)``         (syntax only)
               ``0L``          -> ``0``                    (syntax only, PFDOperations.py:60)
               ``Num.float``   -> ``float``                (removed numpy alias, PFDOperations.py:498)
               ``d.has_key(k)``-> ``(k in d)``             (PFDFile.py:266)
               ``test[ii] not in '...'`` -> ``chr(test[ii]) ...``, ``test.find('\\0')`` -> ``find(b'\\0')``
                                                           (bytes read from the binary .pfd, PFDFile.py:130-139)
  AST level    ``a / b``       -> ``_py2div(a, b)``        classic division: floor for int/int
                                                           (Python ints *and* numpy integer scalars /
                                                           integer arrays), true division otherwise
               ``x == []``     -> ``_eq_emptylist(x)``     2014-era numpy returned False for
                                                           ``ndarray == []``; numpy 2 raises
               ``range(...)``  -> ``list(range(...))``     Python 2 ``range`` is a list
  environment  stub ``matplotlib`` / ``matplotlib.pyplot`` / ``matplotlib.image`` / ``PIL`` modules
               (debug plotting only),
               ``scipy.std = numpy.std`` (the alias the reference imports, ProfileOperations.py:153),
               ``str(x)`` of a float (Python float or numpy.float64) -> Python 2's ``float.__str__``:
               ``'%.12g'`` plus ``'.0'`` when the result has neither ``'.'`` nor an exponent
               (CPython 2.7 ``float_str``: ``PyOS_double_to_string(v, 'g', 12, Py_DTSF_ADD_DOT_0)``;
               numpy < 1.14 printed ``float64`` the same way).  Only the writer uses it
               (``",".join(map(str, scores))``, DataProcessor.py:321,421,442).

It is used (a) by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and (b) by ``tests/test_oracle_vs_reference.py`` (skipped when
``/root/reference`` is absent, e.g. on the GPU box) to pin ``oracle/pfx_oracle.py``
against the reference's own code executed here.
"""
import ast
import builtins
import os
import re
import sys
import types

import numpy as np

REF_SRC = "/root/reference/PulsarFeatureExtractor/src"

_MODULE_ORDER = [
    "Utilities",
    "ProfileOperationsInterface",
    "ProfileOperations",
    "PHCXOperations",
    "PFDOperations",
    "CandidateFileInterface",
    "PHCXFile",
    "SUPERBPHCXFile",
    "PFDFile",
    "Candidate",
    "DataProcessor",
]


def available(src=REF_SRC):
    return os.path.isfile(os.path.join(src, "ProfileOperations.py"))


# ---------------------------------------------------------------------------
# Python 2 semantics helpers injected into every transformed module
# ---------------------------------------------------------------------------
def _is_intlike(v):
    if isinstance(v, bool):
        return True
    if isinstance(v, (int, np.integer)):
        return True
    if isinstance(v, np.ndarray) and v.dtype.kind in "iub":
        return True
    return False


def _py2div(a, b):
    """Python 2 classic division."""
    if _is_intlike(a) and _is_intlike(b):
        return a // b
    return a / b


def _py2str(*args, **kwargs):
    """Python 2 ``str``: floats print with 12 significant digits (see the module docstring)."""
    if len(args) == 1 and not kwargs and isinstance(args[0], float):
        v = float(args[0])
        if v != v:
            return "nan"
        if v in (float("inf"), float("-inf")):
            return "inf" if v > 0 else "-inf"
        t = "%.12g" % v
        if "." not in t and "e" not in t:
            t += ".0"
        return t
    return builtins.str(*args, **kwargs)


def _eq_emptylist(x):
    """``x == []`` as 2014-era numpy evaluated it (False for any ndarray)."""
    if isinstance(x, np.ndarray):
        return False
    return x == []


class _Py2Transformer(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()),
                         args=[node.left, node.right], keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            load_target = ast.parse(ast.unparse(node.target), mode="eval").body
            return ast.copy_location(
                ast.Assign(targets=[node.target],
                           value=ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()),
                                          args=[load_target, node.value], keywords=[])), node)
        return node

    def visit_Compare(self, node):
        self.generic_visit(node)
        if (len(node.ops) == 1 and isinstance(node.ops[0], ast.Eq)
                and isinstance(node.comparators[0], ast.List)
                and not node.comparators[0].elts):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_eq_emptylist", ctx=ast.Load()),
                         args=[node.left], keywords=[]), node)
        return node

    def visit_Call(self, node):
        self.generic_visit(node)
        if isinstance(node.func, ast.Name) and node.func.id == "range":
            return ast.copy_location(
                ast.Call(func=ast.Name(id="list", ctx=ast.Load()), args=[node], keywords=[]), node)
        return node


_PRINT_RE = re.compile(r"^(\s*)print(\s+(.*?))?\s*$")


def _text_fix(src):
    out = []
    for line in src.splitlines():
        m = _PRINT_RE.match(line)
        if m and not line.lstrip().startswith("print("):
            body = m.group(3) or ""
            line = "%sprint(%s)" % (m.group(1), body)
        line = re.sub(r"\b0L\b", "0", line)
        line = line.replace("Num.float", "float")
        line = re.sub(r"(\w+(?:\.\w+)*)\.has_key\(([^)]*)\)", r"(\2 in \1)", line)
        # Python 2 ``str`` read from a binary file is ``bytes`` in Python 3 (PFDFile.py:130-139):
        # indexing gives an int, and ``find`` wants a bytes needle
        line = line.replace("if test[ii] not in '0123456789:.-\\0':", "if chr(test[ii]) not in '0123456789:.-\\0':")
        line = line.replace("test.find('\\0')", "test.find(b'\\0')")
        out.append(line)
    return "\n".join(out) + "\n"


def _install_stubs():
    created = []
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "PIL", "PIL.Image", "Image"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__["__getattr__"] = lambda attr: (lambda *a, **k: None)
                sys.modules[name] = m
                created.append(name)
    if "matplotlib" in created and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import scipy
    if not hasattr(scipy, "std"):
        scipy.std = np.std
    return created


def load_reference(src=REF_SRC, quiet=True):
    """Return {module name: module} for the transformed reference modules."""
    if not available(src):
        raise FileNotFoundError("reference sources not found under %s" % src)
    _install_stubs()
    mods = {}
    saved = {n: sys.modules.get(n) for n in _MODULE_ORDER}
    try:
        for name in _MODULE_ORDER:
            path = os.path.join(src, name + ".py")
            with open(path, "r", encoding="latin-1") as fh:
                text = _text_fix(fh.read())
            tree = ast.parse(text, filename=path)
            tree = ast.fix_missing_locations(_Py2Transformer().visit(tree))
            code = compile(tree, path, "exec")
            mod = types.ModuleType(name)
            mod.__file__ = path
            mod.__dict__["_py2div"] = _py2div
            mod.__dict__["_eq_emptylist"] = _eq_emptylist
            mod.__dict__["str"] = _py2str
            if quiet:
                mod.__dict__["print"] = lambda *a, **k: None
            sys.modules[name] = mod
            exec(code, mod.__dict__)
            mods[name] = mod
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return mods
