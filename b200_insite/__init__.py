"""Importable alias of the package directory
``ode-discovery-for-longitudinal-heterogeneous-treatment-effects-inference_b200/`` (its name, fixed by
the project layout, contains hyphens and cannot be imported directly).  ``import b200_insite`` executes
that directory's ``__init__.py`` with this module's ``__path__`` pointing at it, so
``b200_insite.cancer_simulation`` etc. resolve to the files there.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "ode-discovery-for-longitudinal-heterogeneous-treatment-effects-inference_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
