"""Test-side WRITER of TensorFlow checkpoints (V2 bundle and V1 table), independent of csrc/tfckpt.cpp, following the
published layouts: LevelDB-style SSTable (tensorflow/core/lib/io/table_builder.cc, block_builder.cc, format.cc),
BundleHeaderProto / BundleEntryProto (tensor_bundle.proto), SavedTensorSlices (saved_tensor_slice.proto) and the
ordered-code slice keys (saved_tensor_slice_util.cc).  No TensorFlow exists in this environment, so these files are the
only fixtures the reader can be checked against (parity unpinned)."""
import struct

import numpy as np

DT = {np.dtype('float32'): 1, np.dtype('float64'): 2, np.dtype('int32'): 3, np.dtype('int64'): 9}
MAGIC = 0xdb4775248b80fb57


def _crc_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
        t.append(c)
    return t


_T = _crc_table()


def crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c = (c >> 8) ^ _T[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def mask(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def f_varint(num, v):
    return varint(num << 3) + varint(v)


def f_bytes(num, b):
    return varint((num << 3) | 2) + varint(len(b)) + bytes(b)


def f_fixed32(num, v):
    return varint((num << 3) | 5) + struct.pack('<I', v)


def shape_proto(shape):
    return b''.join(f_bytes(2, f_varint(1, d)) for d in shape)


class BlockBuilder(object):
    def __init__(self, restart_interval):
        self.ri, self.buf, self.restarts, self.counter, self.last = restart_interval, bytearray(), [0], 0, b''

    def add(self, key, value):
        shared = 0
        if self.counter < self.ri:
            n = min(len(key), len(self.last))
            while shared < n and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last = key
        self.counter += 1

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def write_table(path, items, block_size=4096):
    """items: sorted list of (key bytes, value bytes)."""
    assert [k for k, _ in items] == sorted(k for k, _ in items)
    out = bytearray()
    index = BlockBuilder(1)

    def emit(block):
        off = len(out)
        out.extend(block)
        out.extend(b'\x00' + struct.pack('<I', mask(crc32c(block + b'\x00'))))
        return varint(off) + varint(len(block))

    data = BlockBuilder(16)
    for key, value in items:
        data.add(key, value)
        if len(data.buf) >= block_size:
            index.add(key, emit(data.finish()))
            data = BlockBuilder(16)
    if data.buf:
        index.add(data.last, emit(data.finish()))
    meta_handle = emit(BlockBuilder(16).finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC)
    out.extend(footer)
    with open(path, 'wb') as f:
        f.write(bytes(out))


def write_v2(prefix, tensors, block_size=4096):
    """tensors: {name: ndarray}.  Writes <prefix>.index and <prefix>.data-00000-of-00001 like BundleWriter."""
    data = bytearray()
    items = [(b'', f_varint(1, 1) + f_bytes(3, f_varint(1, 1)))]          # BundleHeaderProto{num_shards=1, version{producer=1}}
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        raw = a.tobytes()
        entry = f_varint(1, DT[a.dtype]) + f_bytes(2, shape_proto(a.shape))
        if len(data):
            entry += f_varint(4, len(data))
        entry += f_varint(5, len(raw)) + f_fixed32(6, mask(crc32c(raw)))
        items.append((name.encode(), entry))
        data += raw
    write_table(prefix + '.index', items, block_size)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))


def _ordered_string(s):
    return s.replace(b'\x00', b'\x00\xff').replace(b'\xff', b'\xff\x00') + b'\x00\x01'


def _ordered_num(v):
    b = b'' if v == 0 else v.to_bytes((v.bit_length() + 7) // 8, 'big')
    return bytes([len(b)]) + b


def write_v1(path, tensors, block_size=4096):
    """One table file like TensorSliceWriter (Saver(write_version=1)): meta under key '', one full slice per tensor."""
    meta, items = b'', []
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        full_slice = b''.join(f_bytes(1, b'') for _ in a.shape)           # Extent{} per dim: start 0, no length = whole dim
        meta += f_bytes(1, f_bytes(1, name.encode()) + f_bytes(2, shape_proto(a.shape)) + f_varint(3, DT[a.dtype]) + f_bytes(4, full_slice))
        if a.dtype == np.float32:
            payload = f_bytes(5, a.astype('<f4').tobytes())
        elif a.dtype == np.float64:
            payload = f_bytes(6, a.astype('<f8').tobytes())
        elif a.dtype == np.int32:
            payload = f_bytes(7, b''.join(varint(int(x)) for x in a.reshape(-1)))
        else:
            payload = f_bytes(10, b''.join(varint(int(x)) for x in a.reshape(-1)))
        saved_slice = f_bytes(1, name.encode()) + f_bytes(2, full_slice) + f_bytes(3, payload)
        key = _ordered_num(0) + _ordered_string(name.encode()) + _ordered_num(a.ndim) + b'\x80\x7f' * a.ndim   # (start 0, length -1) per dim
        items.append((key, f_bytes(2, saved_slice)))
    items.sort()
    items.insert(0, (b'', f_bytes(1, meta + f_bytes(2, f_varint(1, 21)))))
    write_table(path, items, block_size)
