"""Import shim: the package directory is named after the reference repository
(`wiflow-wifi-pose-estimation-with-spatio-temporal-decoupling_b200/`), which is not a valid Python identifier.
`import wiflow_b200` (with the repo root on sys.path) loads that directory as the package `wiflow_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        'wiflow-wifi-pose-estimation-with-spatio-temporal-decoupling_b200')
_spec = importlib.util.spec_from_file_location('wiflow_b200', os.path.join(_PKG_DIR, '__init__.py'),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['wiflow_b200'] = _mod
_spec.loader.exec_module(_mod)
