"""Import shim: the package directory is ``poem-v2_b200/`` (not a valid Python identifier).

``import poem_v2_b200`` resolves to this file, which turns itself into a package whose
sub-modules live in ``poem-v2_b200/`` and then executes that directory's ``__init__.py``.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "poem-v2_b200")
__path__ = [_pkg_dir]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
