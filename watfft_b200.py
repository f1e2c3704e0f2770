"""Import shim: the package directory is `wat-fft_b200/` (not a valid Python identifier), so
`import watfft_b200` resolves here and re-exports that directory as the package `watfft_b200`."""
import importlib.util
import pathlib
import sys

_dir = pathlib.Path(__file__).resolve().parent / "wat-fft_b200"
_spec = importlib.util.spec_from_file_location(
    "watfft_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["watfft_b200"] = _mod
_spec.loader.exec_module(_mod)
