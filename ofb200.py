"""Import alias: `import ofb200` loads the package directory
`drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200/` (whose name, fixed by the
repository layout rule, is not a valid Python identifier) and registers it under this name, so
`import ofb200.of_library as of`, `from ofb200 import solve_lgs` etc. work."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200")
_spec = importlib.util.spec_from_file_location("ofb200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ofb200"] = _mod
_spec.loader.exec_module(_mod)
