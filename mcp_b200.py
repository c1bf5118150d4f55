"""Import shim: the package directory is named `monte-carlo-portfolios_b200` (the
reference's repo name plus the target), which is not a Python identifier.  `import
mcp_b200` loads that directory as the package `mcp_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "monte-carlo-portfolios_b200")
_spec = _u.spec_from_file_location("mcp_b200", _os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["mcp_b200"] = _mod
_spec.loader.exec_module(_mod)
