"""Import shim: the package lives in the directory `onnx-transformer_b200/` (the name the project layout
prescribes), which is not a valid Python identifier.  This module makes it importable as
`onnx_transformer_b200` by pointing its __path__ at that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "onnx-transformer_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
