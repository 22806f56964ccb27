"""TEST INFRASTRUCTURE: a minimal HDF5 writer, just enough to exercise the parts of pilot_b200/h5ad.py that no real
file in this image reaches -- chunked datasets with a version-1 chunk B-tree, the shuffle + deflate filter pipeline,
edge chunks, several B-tree entries -- next to contiguous ones.  It follows the HDF5 file-format specification
(version-0 superblock, version-1 object headers, symbol-table groups); h5py / libhdf5 are not available here, so the
files it writes are checked only by our own reader (a self-consistency test of the chunk / filter arithmetic, not an
interoperability test).
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * ((-len(b)) % 8)


def _msg(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        # class 1, version 1; little-endian, mantissa normalisation 2 (implied), sign position
        size = dt.itemsize
        if size == 4:
            bits, props = (0x20, 31, 0), struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        else:
            bits, props = (0x20, 63, 0), struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack("<B3BI", 0x11, *bits, size) + props
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<B3BI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    raise ValueError(dt)


def _dataspace_msg(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)


class Writer:
    def __init__(self):
        self.buf = bytearray(b"\0" * 96)  # superblock, filled at the end
        self.entries = []                  # (name, object header address)

    def _alloc(self, data: bytes) -> int:
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _object_header(self, msgs: bytes, nmsgs: int) -> int:
        hdr = struct.pack("<BBHII4x", 1, 0, nmsgs, 1, len(msgs))
        return self._alloc(hdr + msgs)

    def dataset(self, name: str, arr: np.ndarray, chunks=None, shuffle=False, gzip=None):
        arr = np.ascontiguousarray(arr)
        msgs = _msg(0x01, _dataspace_msg(arr.shape)) + _msg(0x03, _dtype_msg(arr.dtype))
        n = 2
        if chunks is None:
            addr = self._alloc(arr.tobytes())
            msgs += _msg(0x08, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes))
            n += 1
        else:
            es = arr.dtype.itemsize
            rank = arr.ndim
            filters = []
            if shuffle:
                filters.append((2, (es,)))
            if gzip is not None:
                filters.append((1, (gzip,)))
            if filters:
                body = struct.pack("<BB6x", 1, len(filters))
                for fid, cvals in filters:
                    body += struct.pack("<HHHH", fid, 0, 0, len(cvals))
                    body += b"".join(struct.pack("<I", c) for c in cvals)
                    if len(cvals) % 2:
                        body += b"\0" * 4
                msgs += _msg(0x0B, body)
                n += 1
            # chunks, row-major over the chunk grid
            keys = []
            grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
            import itertools
            for offs in itertools.product(*grid):
                block = np.zeros(chunks, dtype=arr.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, arr.shape))
                block[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
                data = block.tobytes()
                if shuffle:
                    a = np.frombuffer(data, dtype=np.uint8).reshape(-1, es)
                    data = a.T.tobytes()
                if gzip is not None:
                    data = zlib.compress(data, gzip)
                caddr = self._alloc(data)
                keys.append((len(data), offs, caddr))
            # one leaf node of the chunk B-tree
            node = b"TREE" + struct.pack("<BBH", 1, 0, len(keys)) + struct.pack("<QQ", UNDEF, UNDEF)
            for nbytes, offs, caddr in keys:
                node += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
                node += struct.pack("<Q", caddr)
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
            baddr = self._alloc(node)
            lay = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", baddr)
            lay += b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", es)
            msgs += _msg(0x08, lay)
            n += 1
        self.entries.append((name, self._object_header(msgs, n)))

    def finish(self) -> bytes:
        # local heap with the link names (offset 0 holds the empty string)
        heap_data = bytearray(b"\0" * 8)
        name_off = {}
        for name, _ in self.entries:
            name_off[name] = len(heap_data)
            heap_data += _pad8(name.encode() + b"\0")
        data_addr = self._alloc(bytes(heap_data))
        heap = b"HEAP" + struct.pack("<B3x", 0) + struct.pack("<QQQ", len(heap_data), UNDEF, data_addr)
        heap_addr = self._alloc(heap)
        # one symbol-table node
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(self.entries))
        for name, oaddr in sorted(self.entries):
            snod += struct.pack("<QQII16x", name_off[name], oaddr, 0, 0)
        snod_addr = self._alloc(snod)
        tree = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF)
        tree += struct.pack("<Q", 0) + struct.pack("<Q", snod_addr) + struct.pack("<Q", name_off[sorted(self.entries)[-1][0]])
        tree_addr = self._alloc(tree)
        root = self._object_header(_msg(0x11, struct.pack("<QQ", tree_addr, heap_addr)), 1)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
        self.buf[:len(sb)] = sb
        return bytes(self.buf)
