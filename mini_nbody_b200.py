"""Import shim: the package directory is named `mini-nbody_b200/` (not a valid Python identifier),
so `import mini_nbody_b200` resolves here and executes that directory's __init__.py."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mini-nbody_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
