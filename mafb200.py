"""Import shim: the package directory is named `membranealefem.jl_b200` (with a dot, as the project layout
prescribes), which Python cannot import by name. `import mafb200` loads it under the module name
`membranealefem_jl_b200` and re-exports its public API."""
import importlib.util
import os
import sys

_NAME = "membranealefem_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "membranealefem.jl_b200")


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = _load()
globals().update({k: v for k, v in vars(pkg).items() if not k.startswith("_")})
