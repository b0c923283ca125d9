"""Import alias: the package directory is named ``grand-plus_b200/`` (after the reference repo),
which is not a valid Python identifier.  ``import grandplus_b200`` resolves to it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "grand-plus_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
