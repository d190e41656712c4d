"""TensorFlow-free reader for TF1 ``tf.train.Saver`` (V2, "tensor bundle") checkpoints: the format of the reference's
decoder checkpoints (wct.py:45-56 restores them with ``saver.restore``; train.py:129,177-185 writes them).

Scope row 8f-1.  A bundle is ``<prefix>.index`` + ``<prefix>.data-0000N-of-0000M``:
  * the index is a LevelDB-style sorted string table: prefix-compressed key/value blocks with restart arrays, a
    per-block trailer (compression byte + masked CRC32C), an index block of block handles and a 48-byte footer
    ending in the magic 0xdb4775248b80fb57;
  * key "" holds a BundleHeaderProto (num_shards, endianness), every other key is a variable name whose value is a
    BundleEntryProto (dtype, shape, shard_id, offset, size, crc32c); tensor bytes sit raw (little endian, row major)
    in the data shard.
TensorFlow is not installable offline and no real checkpoint exists on this machine, so this reader is tested
against files produced by ``tests/tf_bundle_writer.py`` (written from the same format description) -- it has NOT been
run on a checkpoint written by TensorFlow itself; that is said wherever it is used.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           19: np.float16}


class TFCheckpointError(ValueError):
    pass


# ----------------------------------------------------------------------------- primitives
def _varint(b, p):
    x = shift = 0
    while True:
        c = b[p]
        p += 1
        x |= (c & 0x7F) << shift
        if not c & 0x80:
            return x, p
        shift += 7


_CRC_TABLE = None


def crc32c(data, crc=0):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    c = crc ^ 0xFFFFFFFF
    for byte in bytes(data):
        c = _CRC_TABLE[(c ^ byte) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _snappy_decompress(b):
    n, p = _varint(b, 0)
    out = bytearray()
    while p < len(b):
        tag = b[p]
        p += 1
        kind = tag & 3
        if kind == 0:                                  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(b[p:p + nb], "little")
                p += nb
            ln += 1
            out += b[p:p + ln]
            p += ln
            continue
        if kind == 1:
            ln, off = 4 + ((tag >> 2) & 7), ((tag >> 5) << 8) | b[p]
            p += 1
        elif kind == 2:
            ln, off = 1 + (tag >> 2), int.from_bytes(b[p:p + 2], "little")
            p += 2
        else:
            ln, off = 1 + (tag >> 2), int.from_bytes(b[p:p + 4], "little")
            p += 4
        if off == 0 or off > len(out):
            raise TFCheckpointError("corrupt snappy block")
        for _ in range(ln):                            # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise TFCheckpointError("snappy length mismatch")
    return bytes(out)


def _checksum_problem(verify, message):
    """verify=True: raise; verify="warn" (default): warn -- this reader could not be tried on a checkpoint written by TensorFlow
    itself, so a disagreement about what a checksum covers must not make a readable file unreadable; verify=False: ignore."""
    if verify is True:
        raise TFCheckpointError(message)
    if verify:
        import warnings
        warnings.warn("tf_checkpoint: " + message)


def _read_block(f, offset, size, verify):
    raw = f[offset:offset + size + 5]
    if len(raw) < size + 5:
        raise TFCheckpointError("truncated table block")
    contents, ctype = raw[:size], raw[size]
    if verify and struct.unpack("<I", raw[size + 1:size + 5])[0] != masked_crc32c(raw[:size + 1]):
        _checksum_problem(verify, "table block checksum mismatch")
    if ctype == 1:
        contents = _snappy_decompress(contents)
    elif ctype != 0:
        raise TFCheckpointError("unknown block compression %d" % ctype)
    return contents


def _block_entries(block):
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    p, key = 0, b""
    while p < end:
        shared, p = _varint(block, p)
        non_shared, p = _varint(block, p)
        vlen, p = _varint(block, p)
        key = key[:shared] + block[p:p + non_shared]
        p += non_shared
        yield key, block[p:p + vlen]
        p += vlen


def _handle(b, p=0):
    off, p = _varint(b, p)
    size, p = _varint(b, p)
    return off, size, p


def read_table(path, verify="warn"):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    f = open(path, "rb").read()
    if len(f) < 48 or struct.unpack("<Q", f[-8:])[0] != TABLE_MAGIC:
        raise TFCheckpointError("%s: not a tensor-bundle index (bad table magic)" % path)
    footer = f[-48:]
    _, _, p = _handle(footer)                          # metaindex (unused)
    ioff, isize, _ = _handle(footer, p)
    out = []
    for _, hv in _block_entries(_read_block(f, ioff, isize, verify)):
        boff, bsize, _ = _handle(hv)
        out.extend(_block_entries(_read_block(f, boff, bsize, verify)))
    return out


# ----------------------------------------------------------------------------- protobuf (the two messages we need)
def _proto_fields(b):
    p = 0
    while p < len(b):
        tag, p = _varint(b, p)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, p = _varint(b, p)
        elif wt == 1:
            v = b[p:p + 8]
            p += 8
        elif wt == 2:
            ln, p = _varint(b, p)
            v = b[p:p + ln]
            p += ln
        elif wt == 5:
            v = b[p:p + 4]
            p += 4
        else:
            raise TFCheckpointError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _parse_entry(b):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for field, wt, v in _proto_fields(b):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:                            # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = v3
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif field == 7:
            e["sliced"] = True
    return e


def read_bundle(prefix, verify="warn"):
    """{variable name: numpy array} of one checkpoint ``prefix`` (the path without .index / .data-*)."""
    entries = read_table(prefix + ".index", verify)
    if not entries or entries[0][0] != b"":
        raise TFCheckpointError("%s.index: bundle header missing" % prefix)
    num_shards, endian = 1, 0
    for field, _, v in _proto_fields(entries[0][1]):
        if field == 1:
            num_shards = v
        elif field == 2:
            endian = v
    if endian != 0:
        raise TFCheckpointError("big-endian bundles are not supported")
    shards, out = {}, {}
    for key, val in entries[1:]:
        e = _parse_entry(val)
        if e["sliced"]:
            raise NotImplementedError("partitioned variable %r" % key.decode())
        if e["dtype"] not in _DTYPES:
            continue                                   # strings / resources: nothing the decoders need
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb").read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        want = int(np.prod(e["shape"], dtype=np.int64)) * dt.itemsize
        if len(raw) != e["size"] or e["size"] != want:
            raise TFCheckpointError("tensor %r: %d bytes in the shard, shape %s needs %d" % (key.decode(), len(raw), e["shape"], want))
        if verify and e["crc32c"] is not None and e["crc32c"] != masked_crc32c(raw):
            # a tensor whose bytes do not match its own checksum would load garbage weights: always fatal
            raise TFCheckpointError("tensor %r: checksum mismatch (corrupt shard %s.data-%05d-of-%05d)"
                                    % (key.decode(), prefix, sid, num_shards))
        out[key.decode()] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    return out


def latest_checkpoint(path):
    """``tf.train.get_checkpoint_state(dir).model_checkpoint_path`` (wct.py:51-53): a directory holding a ``checkpoint``
    state file, or a checkpoint prefix itself."""
    if os.path.isdir(path):
        state = os.path.join(path, "checkpoint")
        if os.path.exists(state):
            m = re.search(r'^model_checkpoint_path:\s*"(.*)"', open(state).read(), re.M)
            if m:
                p = m.group(1)
                return p if os.path.isabs(p) else os.path.join(path, p)
        return None
    if path.endswith(".index"):
        path = path[:-6]
    return path if os.path.exists(path + ".index") else None


def load_decoder_checkpoint(path, relu_target):
    """Decoder layers of ``relu_target`` from a TF checkpoint dir / prefix, in the engine's format
    [{"name": "<relu>_<count>", "kernel": (kH,kW,I,O) float32, "bias": (O,)} ...].  Variables are matched by their
    layer name ``<relu_target>_<count>`` (model.py:283-298) and the ``kernel`` / ``bias`` leaf, whatever scopes
    precede it (wct.py:47-48 filters on 'decoder_<relu_target>' the same way); optimizer slots are ignored."""
    prefix = latest_checkpoint(path)
    if prefix is None:
        raise Exception("No checkpoint found for target {} in dir {}".format(relu_target, path))     # wct.py:57-58
    layers = {}
    pat = re.compile(r"(?:^|/)%s_(\d+)(?:_\d+)?/(kernel|bias)$" % re.escape(relu_target))
    for name, arr in read_bundle(prefix).items():
        m = pat.search(name)
        if m:
            layers.setdefault(int(m.group(1)), {})[m.group(2)] = np.ascontiguousarray(arr, dtype=np.float32)
    out = []
    for count in sorted(layers):
        l = layers[count]
        if "kernel" not in l or "bias" not in l:
            raise TFCheckpointError("%s: layer %s_%d lacks kernel or bias" % (prefix, relu_target, count))
        out.append(dict(name="%s_%d" % (relu_target, count), kernel=l["kernel"], bias=l["bias"]))
    if not out:
        raise Exception("No checkpoint found for target {} in dir {}".format(relu_target, path))
    # the layer set and every shape must be the decoder model.py:245-304 builds for this target
    from .model import decoder_plan
    want = [op for op in decoder_plan(relu_target) if op.kind == "conv"]
    have = {l["name"]: l for l in out}
    missing = [op.name for op in want if op.name not in have]
    if missing:
        raise TFCheckpointError("%s: decoder %s lacks layer(s) %s (found %s)" % (prefix, relu_target, ", ".join(missing),
                                                                              ", ".join(sorted(have)) or "none"))
    for op in want:
        k, bb = have[op.name]["kernel"], have[op.name]["bias"]
        if tuple(k.shape) != (3, 3, op.cin, op.cout) or tuple(bb.shape) != (op.cout,):
            raise TFCheckpointError("%s: layer %s has kernel %s / bias %s, the decoder needs (3, 3, %d, %d) / (%d,)"
                                    % (prefix, op.name, tuple(k.shape), tuple(bb.shape), op.cin, op.cout, op.cout))
    return [have[op.name] for op in want]
