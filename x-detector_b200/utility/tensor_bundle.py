"""Reader and writer for TensorFlow V2 checkpoints ("tensor bundles": ``<prefix>.index`` + ``<prefix>.data-?????-of-?????``), in
pure Python / numpy -- TensorFlow itself is not a dependency of this package.

Format (TensorFlow r1.x, tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc}, core/lib/io/{table,block,format}.cc,
core/protobuf/tensor_bundle.proto -- restated from the published format, no TF build exists offline to pin it against):
  * ``.index`` is a LevelDB-style immutable table: data blocks of prefix-compressed (key, value) entries
    [shared varint32 | non_shared varint32 | value_len varint32 | key suffix | value], a restart array and its count
    (uint32 LE) at the end of each block, a 5-byte trailer per block (compression type, masked CRC32C), an index
    block whose values are BlockHandles (offset, size varint64) and a 48-byte footer (metaindex handle, index
    handle, padding, magic 0xdb4775248b80fb57).
  * key "" holds BundleHeaderProto {num_shards=1, endianness=2, version=3}; every other key is a tensor name with a
    BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6, slices=7}.
  * tensor bytes live at [offset, offset+size) of data shard ``shard_id``, little-endian, row-major.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_BFLOAT16 = 14


def _crc32c_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _crc32c_table()


def _crc32c_bytes(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


_CHUNK = 4096            # bytes per lane of the vectorised form
_NP_TABLE = np.array(_CRC_TABLE, dtype=np.uint32)
_SHIFT_TABLES = None     # multiplication by x^(8*_CHUNK) as 4 byte-indexed tables


def _shift_tables():
    """The linear map 'append _CHUNK zero bytes' on a finalised CRC (zlib's crc32_combine operator), as four 256-entry
    tables indexed by the bytes of the CRC."""
    global _SHIFT_TABLES
    if _SHIFT_TABLES is None:
        v = (np.arange(256, dtype=np.uint32)[None, :] << (8 * np.arange(4, dtype=np.uint32))[:, None]).reshape(-1)
        for _ in range(_CHUNK):
            v = _NP_TABLE[v & 0xFF] ^ (v >> 8)
        _SHIFT_TABLES = [[int(x) for x in row] for row in v.reshape(4, 256)]
    return _SHIFT_TABLES


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli).  Checkpoint-sized inputs (a ~100 MB tensor) are cut into 4 KB lanes whose CRCs advance
    together as numpy vectors, then folded left to right with the zero-extension operator (crc(A|B) = shift(crc(A),
    |B|) ^ crc(B)): seconds become fractions of a second; small inputs take the byte loop."""
    n = len(data)
    if n < 16 * _CHUNK:
        return _crc32c_bytes(data, crc)
    buf = np.frombuffer(data, dtype=np.uint8)
    head = n % _CHUNK
    c = _crc32c_bytes(bytes(buf[:head]), crc)
    lanes = buf[head:].reshape(-1, _CHUNK)
    v = np.full(lanes.shape[0], 0xFFFFFFFF, dtype=np.uint32)
    cols = np.ascontiguousarray(lanes.T)
    for i in range(_CHUNK):
        v = _NP_TABLE[(v ^ cols[i]) & 0xFF] ^ (v >> 8)
    v ^= np.uint32(0xFFFFFFFF)
    t0, t1, t2, t3 = _shift_tables()
    for x in v.tolist():
        c = t0[c & 0xFF] ^ t1[(c >> 8) & 0xFF] ^ t2[(c >> 16) & 0xFF] ^ t3[c >> 24] ^ x
    return c


def masked_crc32c(data):
    """core/lib/hash/crc32c.h Mask(): rotate right by 15 and add a constant."""
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _proto_fields(buf):
    """Minimal protobuf wire decoder -> list of (field number, wire type, value)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


class BundleEntry(object):
    def __init__(self, buf):
        self.dtype, self.shape, self.shard_id, self.offset, self.size, self.crc32c, self.sliced = 0, [], 0, 0, 0, None, False
        for field, _, v in _proto_fields(buf):
            if field == 1:
                self.dtype = v
            elif field == 2:  # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1; string name = 2 }
                for f2, _, dim in _proto_fields(v):
                    if f2 == 2:
                        size = 0
                        for f3, _, dv in _proto_fields(dim):
                            if f3 == 1:
                                size = _signed64(dv)
                        self.shape.append(size)
            elif field == 3:
                self.shard_id = v
            elif field == 4:
                self.offset = v
            elif field == 5:
                self.size = v
            elif field == 6:
                self.crc32c = v
            elif field == 7:
                self.sliced = True


def _read_block(f, offset, size, verify):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError("truncated table block")
    body, ctype, crc = raw[:size], raw[size], struct.unpack_from("<I", raw, size + 1)[0]
    if verify and masked_crc32c(raw[:size + 1]) != crc:
        raise ValueError("table block checksum mismatch")
    if ctype != 0:
        raise NotImplementedError("compressed table blocks (type %d) are not supported; TF writes bundles uncompressed" % ctype)
    return body


def _block_entries(block):
    """Yield (key, value) of one table block (prefix-compressed keys, restart array at the end)."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


class TensorBundleReader(object):
    """``tf.train.NewCheckpointReader`` for V2 checkpoints: ``has_tensor``, ``get_variable_to_shape_map``,
    ``get_tensor``."""

    def __init__(self, prefix, verify_checksums=True):
        self.prefix = prefix
        self.verify = verify_checksums
        self.entries = {}
        self.num_shards = 1
        with open(prefix + ".index", "rb") as f:
            f.seek(0, os.SEEK_END)
            end = f.tell()
            if end < 48:
                raise ValueError("%s.index is too short to be a table" % prefix)
            f.seek(end - 48)
            footer = f.read(48)
            if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
                raise ValueError("%s.index: bad table magic (not a V2 checkpoint index)" % prefix)
            pos = 0
            _, pos = _varint(footer, pos)  # metaindex handle
            _, pos = _varint(footer, pos)
            ioff, pos = _varint(footer, pos)
            isize, pos = _varint(footer, pos)
            for _, handle in _block_entries(_read_block(f, ioff, isize, self.verify)):
                boff, p2 = _varint(handle, 0)
                bsize, _ = _varint(handle, p2)
                for key, value in _block_entries(_read_block(f, boff, bsize, self.verify)):
                    if key == b"":
                        for field, _, v in _proto_fields(value):
                            if field == 1:
                                self.num_shards = v
                            elif field == 2 and v != 0:
                                raise NotImplementedError("big-endian tensor bundles are not supported")
                    else:
                        self.entries[key.decode("utf-8")] = BundleEntry(value)

    def has_tensor(self, name):
        return name in self.entries

    def get_variable_to_shape_map(self):
        return {k: list(e.shape) for k, e in self.entries.items()}

    def get_tensor(self, name):
        e = self.entries[name]
        if e.sliced:
            raise NotImplementedError("partitioned (sliced) variables are not supported: %s" % name)
        path = "%s.data-%05d-of-%05d" % (self.prefix, e.shard_id, self.num_shards)
        with open(path, "rb") as f:
            f.seek(e.offset)
            raw = f.read(e.size)
        if len(raw) != e.size:
            raise ValueError("%s: truncated data for %s" % (path, name))
        if self.verify and e.crc32c is not None and masked_crc32c(raw) != e.crc32c:
            raise ValueError("%s: checksum mismatch for %s" % (path, name))
        if e.dtype == _DT_BFLOAT16:
            u = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
            return u.view(np.float32).reshape(e.shape)
        if e.dtype not in _DTYPES:
            raise NotImplementedError("dtype enum %d of %s" % (e.dtype, name))
        return np.frombuffer(raw, dtype=np.dtype(_DTYPES[e.dtype]).newbyteorder("<")).reshape(e.shape).copy()


# ---- writer --------------------------------------------------------------------------------------------------------
_DT_OF = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def _vi(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb(num, wt, payload):
    return _vi((num << 3) | wt) + payload


class TensorBundleWriter(object):
    """``tf.train.Saver.save`` for this package's variables: ``add(name, array)`` then ``finish()`` writes
    ``<prefix>.index`` and ``<prefix>.data-00000-of-00001`` in the layout described at the top of this file (sorted keys,
    16-entry restart interval, ~4 KB uncompressed data blocks, masked CRC32C of every block and tensor), i.e. what
    ``TensorBundleReader`` -- and TensorFlow's BundleReader -- read back."""

    def __init__(self, prefix, block_bytes=4096):
        self.prefix, self.block_bytes, self.tensors = prefix, block_bytes, {}

    def add(self, name, array):
        a = np.asarray(array, order="C")  # (ascontiguousarray would turn a scalar into shape [1])
        if a.dtype not in _DT_OF:
            raise TypeError("dtype %s of %s is not supported" % (a.dtype, name))
        self.tensors[name] = a

    @staticmethod
    def _block(entries, restart_interval=16):
        out, restarts, last = bytearray(), [], b""
        for i, (k, v) in enumerate(entries):
            shared = 0
            if i % restart_interval == 0:
                restarts.append(len(out))
            else:
                while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                    shared += 1
            out += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
            last = k
        if not restarts:
            restarts = [0]
        for r in restarts:
            out += struct.pack("<I", r)
        out += struct.pack("<I", len(restarts))
        return bytes(out)

    def finish(self):
        os.makedirs(os.path.dirname(os.path.abspath(self.prefix)), exist_ok=True)
        header = _pb(1, 0, _vi(1)) + _pb(3, 2, _vi(2) + _pb(1, 0, _vi(1)))  # num_shards = 1, little endian, version 1
        kv, offset = [(b"", header)], 0
        with open(self.prefix + ".data-00000-of-00001", "wb") as f:
            for name in sorted(self.tensors):
                a = self.tensors[name]
                raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
                # (proto3: a zero-sized dimension serialises as an empty Dim message)
                dims = b"".join(_pb(2, 2, _vi(len(d)) + d) for d in ((_pb(1, 0, _vi(int(s))) if s else b"") for s in a.shape))
                msg = _pb(1, 0, _vi(_DT_OF[a.dtype])) + _pb(2, 2, _vi(len(dims)) + dims)
                if offset:
                    msg += _pb(4, 0, _vi(offset))
                msg += _pb(5, 0, _vi(len(raw))) + _pb(6, 5, struct.pack("<I", masked_crc32c(raw)))
                kv.append((name.encode("utf-8"), msg))
                f.write(raw)
                offset += len(raw)
        out, index_entries = bytearray(), []

        def emit(block):
            off = len(out)
            out.extend(block + b"\x00")
            out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
            return _vi(off) + _vi(len(block))

        chunk, size = [], 0
        for item in kv:
            chunk.append(item)
            size += len(item[0]) + len(item[1]) + 3
            if size >= self.block_bytes:
                index_entries.append((chunk[-1][0] + b"\x00", emit(self._block(chunk))))
                chunk, size = [], 0
        if chunk:
            index_entries.append((chunk[-1][0] + b"\x00", emit(self._block(chunk))))
        meta = emit(self._block([]))
        index = emit(self._block(index_entries, restart_interval=1))
        footer = meta + index
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
        out.extend(footer)
        with open(self.prefix + ".index", "wb") as f:
            f.write(bytes(out))
        return self.prefix


def update_checkpoint_state(directory, prefix, keep=5):
    """The ``checkpoint`` state file tf.train.Saver maintains (``model_checkpoint_path`` + the kept
    ``all_model_checkpoint_paths``); older checkpoints beyond ``keep`` (keep_checkpoint_max, light_head_rfcn_train.py:
    465-471) are deleted."""
    state = os.path.join(directory, "checkpoint")
    name = os.path.basename(prefix)
    paths = []
    if os.path.isfile(state):
        with open(state) as f:
            for line in f:
                if line.startswith("all_model_checkpoint_paths:"):
                    paths.append(line.split('"')[1])
    paths = [p for p in paths if p != name] + [name]
    for old in paths[:-keep]:
        for suffix in (".index", ".data-00000-of-00001"):
            try:
                os.remove(os.path.join(directory, old + suffix))
            except OSError:
                pass
    paths = paths[-keep:]
    with open(state, "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % name)
        for p in paths:
            f.write('all_model_checkpoint_paths: "%s"\n' % p)
