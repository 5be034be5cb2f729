"""Importable alias of the hyphenated package directory ``mgld-vsr_b200/`` (a hyphen is not a valid module name).

``import mgld_vsr_b200`` executes ``mgld-vsr_b200/__init__.py`` with ``__path__`` pointing at that directory, so
``mgld_vsr_b200.ops`` etc. resolve to the files that live there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mgld-vsr_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
