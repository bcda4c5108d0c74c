"""Reader for TensorFlow V2 checkpoint bundles (``X.ckpt.index`` + ``X.ckpt.data-00000-of-00001``) without
TensorFlow ("next" row N1).

``.index`` is a LevelDB-format SSTable: a 48-byte footer (two varint BlockHandles + magic), an index block
pointing at data blocks; every block holds prefix-compressed (key, value) records followed by a restart array.
Values are BundleEntryProto messages: 1 dtype, 2 shape, 3 shard_id, 4 offset, 5 size, 6 crc32c.
``.data`` holds the raw little-endian tensors at those offsets.  A minimal writer exists for round-trip tests.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}
_DTYPE_CODE = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def _varint(b, p):
    r = s = 0
    while True:
        c = b[p]
        p += 1
        r |= (c & 0x7f) << s
        s += 7
        if c < 0x80:
            return r, p


def _put_varint(v):
    out = bytearray()
    while True:
        c = v & 0x7f
        v >>= 7
        if v:
            out.append(c | 0x80)
        else:
            out.append(c)
            return bytes(out)


def _read_block(buf, off, size):
    blk = buf[off:off + size]
    nrestart = struct.unpack("<I", blk[-4:])[0]
    end = len(blk) - 4 - 4 * nrestart
    p, key, out = 0, b"", []
    while p < end:
        shared, p = _varint(blk, p)
        nonshared, p = _varint(blk, p)
        vlen, p = _varint(blk, p)
        key = key[:shared] + blk[p:p + nonshared]
        p += nonshared
        out.append((key, blk[p:p + vlen]))
        p += vlen
    return out


def _parse_proto(v):
    p, d = 0, {}
    while p < len(v):
        tag, p = _varint(v, p)
        f, w = tag >> 3, tag & 7
        if w == 0:
            d[f], p = _varint(v, p)
        elif w == 2:
            ln, p = _varint(v, p)
            d[f] = v[p:p + ln]
            p += ln
        elif w == 5:
            d[f] = struct.unpack("<I", v[p:p + 4])[0]
            p += 4
        elif w == 1:
            d[f] = struct.unpack("<Q", v[p:p + 8])[0]
            p += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % w)
    return d


def _parse_shape(b):
    dims, p = [], 0
    while p < len(b):
        tag, p = _varint(b, p)
        if tag & 7 == 2:
            ln, p = _varint(b, p)
            sub, q = b[p:p + ln], 0
            p += ln
            if tag >> 3 == 2:                      # TensorShapeProto.dim
                size = 0
                while q < len(sub):
                    t, q = _varint(sub, q)
                    if t & 7 == 0:
                        val, q = _varint(sub, q)
                        if t >> 3 == 1:
                            size = val
                    elif t & 7 == 2:
                        l2, q = _varint(sub, q)
                        q += l2
                dims.append(size)
        elif tag & 7 == 0:
            _, p = _varint(b, p)
    return dims


def read_index(index_path):
    """-> OrderedDict name -> dict(dtype, shape, shard_id, offset, size)."""
    buf = open(index_path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _MAGIC:
        raise ValueError("%s is not an SSTable (.index) file" % index_path)
    footer = buf[-48:]
    p = 0
    _, p = _varint(footer, p)
    _, p = _varint(footer, p)
    ioff, p = _varint(footer, p)
    isize, p = _varint(footer, p)
    entries = OrderedDict()
    for _, handle in _read_block(buf, ioff, isize):
        off, q = _varint(handle, 0)
        size, q = _varint(handle, q)
        for key, val in _read_block(buf, off, size):
            if key == b"":
                continue                                # BundleHeaderProto
            e = _parse_proto(val)
            entries[key.decode()] = dict(dtype=e.get(1, 0), shape=_parse_shape(e.get(2, b"")), shard_id=e.get(3, 0),
                                         offset=e.get(4, 0), size=e.get(5, 0))
    return entries


def read_checkpoint(prefix, skip_optimizer_slots=True):
    """``prefix`` = path without the .index/.data suffix (what tf.train.Saver.restore takes)."""
    entries = read_index(prefix + ".index")
    out = OrderedDict()
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        for name, e in entries.items():
            if skip_optimizer_slots and (name.endswith("/Adam") or name.endswith("/Adam_1") or
                                         name in ("beta1_power", "beta2_power")):
                continue
            if e["dtype"] not in _DTYPES:
                continue
            f.seek(e["offset"])
            raw = f.read(e["size"])
            out[name] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


# ---- minimal writer (tests / exporting synthetic weights in the reference's own format) -------------------
def _block(records):
    body = bytearray()
    restarts = []
    for key, val in records:                           # restart interval 1: no prefix sharing
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(val)) + key + val
    for r in restarts or [0]:
        body += struct.pack("<I", r)
    body += struct.pack("<I", max(1, len(restarts)))
    return bytes(body)


def _shape_proto(shape):
    out = bytearray()
    for d in shape:
        dim = b"\x08" + _put_varint(int(d))
        out += b"\x12" + _put_varint(len(dim)) + dim
    return bytes(out)


def write_checkpoint(prefix, tensors):
    """Write ``tensors`` (name -> array) as a one-shard bundle readable by ``read_checkpoint``."""
    names = sorted(tensors.keys())
    records = [(b"", b"\x08\x01")]                     # BundleHeaderProto{num_shards: 1}
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.ascontiguousarray(tensors[n])
            raw = a.tobytes()
            shape = _shape_proto(a.shape)
            val = (b"\x08" + _put_varint(_DTYPE_CODE[a.dtype]) + b"\x12" + _put_varint(len(shape)) + shape +
                   b"\x20" + _put_varint(f.tell()) + b"\x28" + _put_varint(len(raw)))
            records.append((n.encode(), val))
            f.write(raw)
    data = _block(records)
    trailer = b"\x00" + struct.pack("<I", 0)            # no compression, crc unchecked by this reader
    index_blk = _block([(names[-1].encode() + b"\xff" if names else b"\xff", _put_varint(0) + _put_varint(len(data)))])
    meta_blk = _block([])
    out = bytearray()
    out += data + trailer
    meta_off = len(out)
    out += meta_blk + trailer
    idx_off = len(out)
    out += index_blk + trailer
    footer = _put_varint(meta_off) + _put_varint(len(meta_blk)) + _put_varint(idx_off) + _put_varint(len(index_blk))
    footer = footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out += footer
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
