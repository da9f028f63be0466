"""Load the *reference's own* modules by file path (this container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
this is used solely by ``oracle/gen_golden.py`` (fixture generation) and by
``tests/test_oracle_vs_reference.py`` (skipped when the mount is absent).
No reference source is copied into this repository.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NAMASTE_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "namaste", "Crossfield_transit.py"))


def _load(name, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def crossfield():
    """namaste/Crossfield_transit.py, unmodified (needs a ``scipy.misc`` name)."""
    try:
        import scipy.misc  # noqa: F401  (deprecated stub in scipy 1.x)
    except Exception:  # scipy >= 2: provide an empty module so the import line works
        stub = types.ModuleType("scipy.misc")
        sys.modules["scipy.misc"] = stub
        import scipy
        scipy.misc = stub
    return _load("_ref_crossfield_transit", "namaste/Crossfield_transit.py")


def k2flatten():
    """namaste/k2flatten.py, unmodified."""
    return _load("_ref_k2flatten", "namaste/k2flatten.py")


def namaste_function(name):
    """One top-level function of namaste/namaste.py, compiled from the reference's own source text.

    The module itself cannot be imported here (autograd, celerite, emcee, astropy are absent), but its
    pure-NumPy helpers can be run as they are: the function definition is cut out of the file with
    ``ast`` and executed in a namespace that only holds ``numpy as np``.  Build container only."""
    import ast
    import numpy as np
    path = os.path.join(REFERENCE_ROOT, "namaste", "namaste.py")
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {"np": np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns[name]
    raise KeyError(name)
