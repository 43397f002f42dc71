"""Importable alias of the `robust-object-detection_b200/` package directory.

The package directory carries the repository's hyphenated name, which Python cannot import;
this shim points the import system at it, so
    import robust_object_detection_b200.augmentations as augmentations
loads robust-object-detection_b200/augmentations.py.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "robust-object-detection_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
