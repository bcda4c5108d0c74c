"""epc-net_b200 -- B200-native (sm_100a) implementation of EPC-Net's embedding-and-retrieval hot path.

The directory name mirrors the reference's hyphenated module names (``importlib.import_module("epc-net")``,
evaluate.py:119); import it with ``importlib.import_module("epc-net_b200")`` or through the top-level alias
module ``epc_net_b200``.

Layout (only what the path needs):
  csrc/            hand-written CUDA (sm_100a) + the C ABI of include/epc_b200.h  -> libepc_b200.so
  _lib.py          ctypes binding (fails loudly if the library is missing; no CPU fallback)
  engine.py        model handle, workspaces, chunking, host-buffer pipeline
  variables.py     TF variable names <-> arrays, synthetic weights;  tf_bundle.py: TF checkpoint reader
  models/          the reference's plugin modules: epc-net, epc-net-l, kd_epc-net, kd_epc-net-l
  loupe.py         NetVLAD / G_VLAD interface;  utils/tf_util.py: operator wrappers
  evaluate.py      get_latent_vectors / get_recall / evaluate;  dist.py: multi-GPU sharding
"""
__version__ = "0.1.0"
