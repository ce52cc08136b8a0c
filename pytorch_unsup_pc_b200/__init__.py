"""Importable alias of the ``pytorch-unsup-pc_b200/`` package directory.

The package directory carries the reference's repository name (with hyphens),
which Python cannot import; this stub points ``__path__`` at it and executes
its ``__init__`` so that ``import pytorch_unsup_pc_b200`` and
``pytorch_unsup_pc_b200.ops`` resolve to the files there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "pytorch-unsup-pc_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
