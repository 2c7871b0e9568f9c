"""Test fixture generator: writes a TensorFlow V2 checkpoint (tensor bundle) the way BundleWriter does -- sorted keys,
prefix-compressed table blocks with restart points every 16 entries, uncompressed, masked CRC32C trailers, 48-byte footer
-- so that the reader (x-detector_b200/utility/tensor_bundle.py) can be exercised without TensorFlow.  Independent of the
reader's code except for the CRC helper."""
import struct

import numpy as np

from xdet_b200.utility.tensor_bundle import masked_crc32c

_DT = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


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


def _field(num, wt, payload):
    return _vi((num << 3) | wt) + payload


def _entry_proto(arr, offset, shard_id=0):
    raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
    dims = b"".join(_field(2, 2, _vi(len(d)) + d) for d in (_field(1, 0, _vi(int(s))) for s in arr.shape))
    msg = _field(1, 0, _vi(_DT[arr.dtype]))
    msg += _field(2, 2, _vi(len(dims)) + dims)
    if shard_id:
        msg += _field(3, 0, _vi(shard_id))
    if offset:
        msg += _field(4, 0, _vi(offset))
    msg += _field(5, 0, _vi(len(raw)))
    msg += _field(6, 5, struct.pack("<I", masked_crc32c(raw)))
    return msg, raw


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
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix, tensors, entries_per_block=7):
    """tensors: {name: numpy array}.  Writes prefix.index and prefix.data-00000-of-00001."""
    names = sorted(tensors)
    data, kv = bytearray(), []
    header = _field(1, 0, _vi(1)) + _field(3, 2, _vi(2) + _field(1, 0, _vi(1)))  # num_shards=1, version{producer=1}
    kv.append((b"", header))
    for n in names:
        msg, raw = _entry_proto(np.asarray(tensors[n], order="C"), len(data))
        kv.append((n.encode(), msg))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    out, index_entries = bytearray(), []

    def emit(block):
        off = len(out)
        out.extend(block + b"\x00")
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return _vi(off) + _vi(len(block))

    for i in range(0, len(kv), entries_per_block):
        chunk = kv[i:i + entries_per_block]
        handle = emit(_block(chunk))
        index_entries.append((chunk[-1][0] + b"\x00", handle))  # a separator >= the block's last key
    meta = emit(_block([]))
    index = emit(_block(index_entries, restart_interval=1))
    footer = meta + index
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
