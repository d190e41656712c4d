"""Test utility: write a TF1 Saver V2 "tensor bundle" (<prefix>.index + <prefix>.data-00000-of-00001 + the directory's
`checkpoint` state file) from a {name: array} dict, following the LevelDB table / BundleEntryProto layout that
wct_tf_b200/tf_checkpoint.py reads.  (TensorFlow itself is not installable here.)"""
import os
import struct

import numpy as np

from wct_tf_b200.tf_checkpoint import TABLE_MAGIC, masked_crc32c

_DT = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def _vi(x):
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _field(num, wt, payload):
    return _vi((num << 3) | wt) + payload


def _entry(arr, offset, size, crc):
    shape = b"".join(_field(2, 2, _vi(len(d)) + d) for d in (_field(1, 0, _vi(int(s))) for s in arr.shape))
    return (_field(1, 0, _vi(_DT[arr.dtype])) + _field(2, 2, _vi(len(shape)) + shape) + _field(3, 0, _vi(0)) +
            _field(4, 0, _vi(offset)) + _field(5, 0, _vi(size)) + _field(6, 5, struct.pack("<I", crc)))


def _block(pairs, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        last = k
    for r in restarts or [0]:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts) or 1)
    return bytes(out)


def write_bundle(prefix, tensors, block_entries=5):
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    data, pairs = bytearray(), [(b"", _field(1, 0, _vi(1)) + _field(2, 0, _vi(0)))]          # BundleHeaderProto: 1 shard, little endian
    for name in sorted(tensors):
        a = np.asarray(tensors[name], order="C")          # (ascontiguousarray would turn a scalar into shape (1,))
        raw = a.tobytes()
        pairs.append((name.encode(), _entry(a, len(data), len(raw), masked_crc32c(raw))))
        data += raw
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    f, index = bytearray(), []

    def emit(contents):
        off = len(f)
        f.extend(contents + b"\x00" + struct.pack("<I", masked_crc32c(contents + b"\x00")))
        return _vi(off) + _vi(len(contents))

    for i in range(0, len(pairs), block_entries):                                            # several small data blocks
        chunk = pairs[i:i + block_entries]
        index.append((chunk[-1][0] + b"\x00", emit(_block(chunk))))
    meta = emit(_block([]))
    idx = emit(_block(index, restart_interval=1))
    footer = meta + idx
    f.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    open(prefix + ".index", "wb").write(bytes(f))
    with open(os.path.join(os.path.dirname(prefix), "checkpoint"), "w") as s:
        s.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (os.path.basename(prefix), os.path.basename(prefix)))
