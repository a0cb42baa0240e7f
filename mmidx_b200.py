"""Import shim: the package directory is `multimedia-indexing_b200/` (a hyphen is not importable), so this
module loads it under the name `mmidx_b200` and re-exports its public names."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multimedia-indexing_b200")
_name = "multimedia_indexing_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    try:
        _spec.loader.exec_module(_mod)
    except BaseException:
        del sys.modules[_name]
        raise
pkg = sys.modules[_name]
from multimedia_indexing_b200 import *  # noqa: E402,F401
from multimedia_indexing_b200 import _capi, datastructures, aggregation  # noqa: E402,F401
