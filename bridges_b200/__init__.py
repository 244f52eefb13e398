"""Import alias: the product package lives in `bridges-with-reinforcement-learning_b200/`
(a directory name Python cannot import directly); `import bridges_b200` resolves to it."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "bridges-with-reinforcement-learning_b200")
__path__.insert(0, _pkg_dir)
with open(_os.path.join(_pkg_dir, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
