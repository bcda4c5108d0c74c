"""The reference's model plugins (``MODEL = importlib.import_module(args["ARCH"])``, evaluate.py:119).

``load(arch)`` returns the module for a yaml ``ARCH`` value.  The hyphenated file names are kept and every plugin also
imports as a TOP-LEVEL module: with this directory on ``sys.path`` (the reference appends its own ``models`` directory,
evaluate.py:11-13) the reference's unmodified ``importlib.import_module(args["ARCH"])`` finds them
(tests/test_host.py::test_reference_style_plugin_import).
"""
import importlib

ARCHS = ("epc-net", "epc-net-l", "kd_epc-net", "kd_epc-net-l")


def load(arch: str):
    if arch not in ARCHS:
        raise ImportError("unknown ARCH %r (expected one of %s)" % (arch, ", ".join(ARCHS)))
    return importlib.import_module(__name__ + "." + arch)
