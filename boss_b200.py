"""Import shim: the package lives in ``boss.jl_b200/`` (a directory name Python cannot import
directly because of the dot); this module loads it under the importable name ``boss_b200``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "boss.jl_b200")
_spec = _ilu.spec_from_file_location("boss_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["boss_b200"] = _mod
_spec.loader.exec_module(_mod)
