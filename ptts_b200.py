"""Import shim: the package directory is named `pocket-tts.cpp_b200` (not a Python identifier), so it is loaded by path.
    import ptts_b200 as P;  ctx = P.Context(model_dir, max_slots=8);  s = ctx.stream("cosette", temp=0.0)
"""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pocket-tts.cpp_b200")
_spec = _u.spec_from_file_location("pocket_tts_cpp_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["pocket_tts_cpp_b200"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
