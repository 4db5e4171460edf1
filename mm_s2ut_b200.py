"""Import alias: ``import mm_s2ut_b200`` loads the package directory ``multimodal-s2ut_b200/``
(whose name is not a valid Python identifier) under an importable name."""
import importlib.util as _u
import sys as _sys
from pathlib import Path as _P

_dir = _P(__file__).resolve().parent / "multimodal-s2ut_b200"
_spec = _u.spec_from_file_location(__name__, _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
