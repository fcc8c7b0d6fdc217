"""Importable alias of the ``flink-parameter-server_b200/`` package directory.

The framework's sources live in ``flink-parameter-server_b200/`` (a name that is not a valid
Python identifier); this shim makes them importable as ``fps_b200``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "flink-parameter-server_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
