"""Mint tests/golden/graph_*.npz by EXECUTING THE REFERENCE'S SHIPPED GRAPHS (exp/*/saved_model/*.meta)
with oracle/graphdef_interp.py on seeded inputs and seeded synthetic weights.

Runs only where /root/reference is mounted (the build container); the .npz outputs are committed and
travel to the GPU box.  Usage:  python tests/golden/make_golden.py [epc-net epc-net-l kd_epc-net kd_epc-net-l]
"""
from __future__ import annotations

import hashlib
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import graphdef_interp as gi  # noqa: E402
import _data  # noqa: E402

variables = importlib.import_module("epc-net_b200.variables")

REF = "/root/reference"
STUDENT_META = "exp/epc-net-l-d/saved_model/student_model_epoch20_iter18101.ckpt.meta"
GRAPHS = {
    # arch: (meta file, variable scope, placeholder prefix, is_training placeholder)
    "epc-net": ("exp/epc-net/saved_model/model_epoch22_iter18101.ckpt.meta", "query_triplets", "", "Placeholder_4"),
    "epc-net-l": ("exp/epc-net-l/saved_model/model_epoch13_iter18101.ckpt.meta", "query_triplets", "", "Placeholder_4"),
    # the student .meta embeds BOTH graphs
    "kd_epc-net-l": (STUDENT_META, "student/query_triplets", "student/", "student/Placeholder_6"),
    "kd_epc-net": (STUDENT_META, "teacher/query_triplets", "teacher/", "teacher/Placeholder_4"),
}
WEIGHT_SEED = {"epc-net": 11, "epc-net-l": 12, "kd_epc-net-l": 13, "kd_epc-net": 14}
CLOUD_SEED = 1000
SAMPLE_POINTS = np.array([0, 1, 17, 255, 1024, 2047, 4095])


def main(archs):
    for arch in archs:
        meta, scope, ppre, train_ph = GRAPHS[arch]
        t0 = time.time()
        gd = gi.load_meta_graph(os.path.join(REF, meta)).graph_def
        V = variables.synthetic_variables(arch, WEIGHT_SEED[arch], scope)
        clouds = _data.golden_batch(CLOUD_SEED)
        # query 1, positives 2, negatives 14, other_neg 1  (train.py:238-255) -> concat -> 18 clouds
        feed = {ppre + "Placeholder": clouds[0:1][None], ppre + "Placeholder_1": clouds[1:3][None],
                ppre + "Placeholder_2": clouds[3:17][None], ppre + "Placeholder_3": clouds[17:18][None],
                train_ph: np.bool_(False)}
        interp = gi.GraphInterpreter(gd, V, small_k_matmul="muladd")
        bscope = "BACKBONE" if arch == "kd_epc-net-l" else "fastdgcnn"
        fs = scope + "/" + bscope
        out = interp.run(scope + "/VLAD/last_output", feed)                     # (1,18,256)
        kth = interp.run(fs + "/Min", feed)                                     # (18,4096,1)
        mask = interp.run(fs + "/Cast", feed)                                   # (18,4096,4096)
        conv5 = interp.run(fs + "/conv5/Relu", feed)                            # (18,4096,1024)
        concat = interp.run(fs + "/concat", feed)                               # (18,4096,64*nblk)
        res = {
            "arch": arch, "scope": scope, "weight_seed": WEIGHT_SEED[arch], "cloud_seed": CLOUD_SEED,
            "output": out.astype(np.float32),
            "kth": kth[..., 0].astype(np.float32),
            "count": mask.sum(-1).astype(np.int32),
            "mask_sha256": hashlib.sha256(np.packbits(mask.astype(np.bool_)).tobytes()).hexdigest(),
            "sample_points": SAMPLE_POINTS,
            "conv5_rows": conv5[:, SAMPLE_POINTS, :].astype(np.float32),
            "concat_rows": concat[:, SAMPLE_POINTS, :].astype(np.float32),
            "conv5_colmax": conv5.max(axis=1).astype(np.float32),
        }
        if arch in ("epc-net", "kd_epc-net"):
            flat = interp.run(scope + "/VLAD/l2_normalize_2", feed)             # (18,65536) f-major
            res["vlad_flat_sample"] = flat[:, ::257].astype(np.float32)
        if arch.startswith("kd_"):
            fea = interp.run(scope + "/l2_normalize", feed)                     # (18*4096,1024) KD feature
            res["kd_feat_rows"] = fea.reshape(18, 4096, 1024)[:, SAMPLE_POINTS, :].astype(np.float32)
        res["ops_executed"] = np.array(sorted(set(interp.executed_ops)))
        path = os.path.join(HERE, "graph_%s.npz" % arch)
        np.savez_compressed(path, **res)
        print("%s: wrote %s (%.1fs, %d KB) out[0,0,:3]=%s max count per cloud=%s" % (
            arch, os.path.basename(path), time.time() - t0, os.path.getsize(path) // 1024, out[0, 0, :3],
            res["count"].max(axis=1)))


if __name__ == "__main__":
    main(sys.argv[1:] or list(GRAPHS))
