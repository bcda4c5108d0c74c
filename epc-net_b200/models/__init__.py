"""The reference's model plugins (``MODEL = importlib.import_module(args["ARCH"])``, evaluate.py:119).

``load(arch)`` returns the module for a yaml ``ARCH`` value; the hyphenated file names are kept so that
adding this directory to ``sys.path`` lets the reference's own import statement find them.
"""
import importlib

ARCHS = ("epc-net", "epc-net-l", "kd_epc-net", "kd_epc-net-l")


def load(arch: str):
    if arch not in ARCHS:
        raise ImportError("unknown ARCH %r (expected one of %s)" % (arch, ", ".join(ARCHS)))
    return importlib.import_module(__name__ + "." + arch)
