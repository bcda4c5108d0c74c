"""Importable alias of the hyphenated package directory ``epc-net_b200/`` (a hyphen is not a valid
identifier, exactly like the reference's ``epc-net.py`` plugins): ``import epc_net_b200`` gives the package,
and ``epc_net_b200.X`` resolves to the very same module objects as ``epc-net_b200.X``."""
import importlib
import sys

_pkg = importlib.import_module("epc-net_b200")
for _sub in ("variables", "tf_bundle", "_lib"):
    importlib.import_module("epc-net_b200." + _sub)


class _AliasFinder(object):
    """Maps ``epc_net_b200.<sub>`` to the already-imported (or importable) ``epc-net_b200.<sub>``."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        if not name.startswith("epc_net_b200."):
            return None
        real = "epc-net_b200." + name[len("epc_net_b200."):]
        mod = importlib.import_module(real)
        sys.modules[name] = mod
        return importlib.util.spec_from_loader(name, loader=_Loader(mod))


class _Loader(object):
    def __init__(self, mod):
        self.mod = mod

    def create_module(self, spec):
        return self.mod

    def exec_module(self, module):
        pass


import importlib.util  # noqa: E402

sys.meta_path.insert(0, _AliasFinder)
sys.modules[__name__] = _pkg
