"""Host ingest, the step before the hot path: reading an ``.h5ad`` file.

The reference's ``load_h5ad`` (/root/reference/pilotpy/tools/Trajectory.py:121-137) is ``scanpy.read_h5ad(path)``.
Neither anndata nor any HDF5 library (h5py, PyTables, libhdf5) exists in this image, so this module carries a small
pure-Python reader of the HDF5 subset that anndata writes -- version-0/1 superblock, version-1 object headers,
symbol-table ("old style") groups and compact link messages, contiguous / compact / chunked (B-tree v1) datasets
with the deflate, shuffle and fletcher32 filters, fixed-point / floating-point / fixed and variable-length string /
enum (bool) / object-reference datatypes, attributes -- and maps anndata's on-disk layout (dense or csr/csc ``X``,
``obs`` / ``var`` dataframes with categorical columns in the 0.7 ``__categories`` and the 0.8 ``categorical``
encodings, ``obsm``, ``uns``) onto a minimal AnnData stand-in that is all ``wasserstein_distance`` needs:
``.X``, ``.obs``, ``.var``, ``.var_names``, ``.obsm``, ``.uns``, ``.to_df()``.

When ``anndata`` is importable ``load_h5ad`` simply calls it, like the reference.  Categorical columns come back as
``pandas.Categorical`` built FROM THE STORED CODES (no string factorisation on the way: the codes go to the GPU as
they are, ``tl._Labels``).  Tested against the reference's own tutorial file
(``Tutorial/Datasets/Kidney_IgAN_G.h5ad``, tests/test_h5ad.py, tests/golden/make_kidney_golden.py).
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class _Datatype:
    """Decoded datatype message."""
    __slots__ = ("cls", "size", "dtype", "vlen_str", "vlen_base", "is_ref", "enum_names", "enum_values", "strpad",
                 "fields")

    def __init__(self):
        self.cls = -1
        self.size = 0
        self.dtype = None        # numpy dtype of one stored element (None for variable-length)
        self.vlen_str = False
        self.vlen_base = None
        self.is_ref = False
        self.enum_names = None
        self.enum_values = None
        self.strpad = 0
        self.fields = None


def _parse_datatype(buf: bytes, off: int = 0) -> Tuple[_Datatype, int]:
    """Datatype message at buf[off:]; returns (type, bytes consumed)."""
    t = _Datatype()
    cv = buf[off]
    t.cls = cv & 0x0F
    version = cv >> 4
    bits = buf[off + 1] | (buf[off + 2] << 8) | (buf[off + 3] << 16)
    t.size = struct.unpack_from("<I", buf, off + 4)[0]
    p = off + 8
    if t.cls == 0:      # fixed point
        order = ">" if bits & 1 else "<"
        signed = bool(bits & 0x08)
        t.dtype = np.dtype(f"{order}{'i' if signed else 'u'}{t.size}")
        p += 4
    elif t.cls == 1:    # floating point
        order = ">" if bits & 1 else "<"
        t.dtype = np.dtype(f"{order}f{t.size}")
        p += 12
    elif t.cls == 3:    # fixed-length string
        t.strpad = bits & 0x0F
        t.dtype = np.dtype(f"S{t.size}")
    elif t.cls == 4:    # bit field
        t.dtype = np.dtype(f"<u{t.size}")
        p += 4
    elif t.cls == 5:    # opaque
        taglen = ((bits & 0xFF) + 7) & ~7
        t.dtype = np.dtype(f"V{t.size}")
        p += taglen
    elif t.cls == 6:    # compound
        n = bits & 0xFFFF
        fields = []
        for _ in range(n):
            e = buf.index(b"\0", p)
            name = buf[p:e].decode("utf-8")
            if version < 3:
                p += ((e - p) // 8 + 1) * 8
            else:
                p = e + 1
            if version == 1:
                boff = struct.unpack_from("<I", buf, p)[0]
                p += 4 + 1 + 3 + 4 + 4 + 16
            elif version == 2:
                boff = struct.unpack_from("<I", buf, p)[0]
                p += 4
            else:
                nb = max(1, (int(t.size).bit_length() + 7) // 8)
                boff = int.from_bytes(buf[p:p + nb], "little")
                p += nb
            ft, used = _parse_datatype(buf, p)
            p += used
            fields.append((name, boff, ft))
        t.fields = fields
        if all(f[2].dtype is not None for f in fields):
            t.dtype = np.dtype({"names": [f[0] for f in fields], "formats": [f[2].dtype for f in fields],
                                "offsets": [f[1] for f in fields], "itemsize": t.size})
    elif t.cls == 7:    # reference
        t.is_ref = True
        t.dtype = np.dtype(f"<u{t.size}") if t.size in (1, 2, 4, 8) else np.dtype(f"V{t.size}")
    elif t.cls == 8:    # enumeration
        n = bits & 0xFFFF
        base, used = _parse_datatype(buf, p)
        p += used
        names = []
        for _ in range(n):
            e = buf.index(b"\0", p)
            names.append(buf[p:e].decode("utf-8"))
            if version < 3:
                p += ((e - p) // 8 + 1) * 8
            else:
                p = e + 1
        vals = np.frombuffer(buf, dtype=base.dtype, count=n, offset=p)
        p += n * base.size
        t.dtype = base.dtype
        t.enum_names, t.enum_values = names, vals.copy()
    elif t.cls == 9:    # variable length
        kind = bits & 0x0F
        base, used = _parse_datatype(buf, p)
        p += used
        t.vlen_str = kind == 1
        t.vlen_base = base
        t.dtype = None
    elif t.cls == 10:   # array
        ndim = buf[p]
        p += 1 if version >= 3 else 4
        dims = struct.unpack_from(f"<{ndim}I", buf, p)
        p += 4 * ndim
        if version < 3:
            p += 4 * ndim
        base, used = _parse_datatype(buf, p)
        p += used
        t.dtype = np.dtype((base.dtype, tuple(dims)))
    else:
        raise H5Error(f"unsupported HDF5 datatype class {t.cls}")
    return t, p - off


def _parse_dataspace(buf: bytes, off: int, L: int) -> Optional[Tuple[int, ...]]:
    """Dataspace message -> shape; () for a scalar, None for a null dataspace."""
    version, rank, flags = buf[off], buf[off + 1], buf[off + 2]
    if version == 1:
        p = off + 8
    elif version == 2:
        if buf[off + 3] == 2:
            return None
        p = off + 4
    else:
        raise H5Error(f"dataspace message version {version}")
    fmt = "<Q" if L == 8 else "<I"
    return tuple(struct.unpack_from(fmt, buf, p + i * L)[0] for i in range(rank))


class _Object:
    """An object header: its messages, decoded lazily."""

    def __init__(self, f: "H5File", addr: int):
        self.f = f
        self.addr = addr
        self.msgs: List[Tuple[int, bytes]] = []
        f._read_object_header(addr, self.msgs)
        self._attrs = None
        self._links = None

    def _first(self, mtype: int) -> Optional[bytes]:
        for t, b in self.msgs:
            if t == mtype:
                return b
        return None

    # ---- groups ----
    @property
    def is_group(self) -> bool:
        return self._first(0x11) is not None or self._first(0x02) is not None or \
            (self._first(0x08) is None and self._first(0x06) is not None) or \
            (self._first(0x08) is None and self._first(0x03) is None)

    def links(self) -> Dict[str, int]:
        if self._links is None:
            out: Dict[str, int] = {}
            st = self._first(0x11)
            if st is not None:
                O = self.f.O
                btree, heap = self.f._unpack_addr(st, 0), self.f._unpack_addr(st, O)
                self.f._walk_group_btree(btree, heap, out)
            for t, b in self.msgs:
                if t == 0x06:
                    name, addr = self.f._parse_link(b)
                    if addr is not None:
                        out[name] = addr
                elif t == 0x02:
                    # link info: dense storage (fractal heap) is not read
                    O = self.f.O
                    flags = b[1]
                    p = 2 + (8 if flags & 1 else 0)
                    fheap = self.f._unpack_addr(b, p)
                    if fheap != UNDEF:
                        raise H5Error("group with dense link storage (fractal heap): not supported by this reader")
            self._links = out
        return self._links

    def __contains__(self, name: str) -> bool:
        return name in self.links()

    def keys(self):
        return list(self.links().keys())

    def __getitem__(self, path: str) -> "_Object":
        obj = self
        for part in path.strip("/").split("/"):
            if not part:
                continue
            lk = obj.links()
            if part not in lk:
                raise KeyError(part)
            obj = self.f.object_at(lk[part])
        return obj

    # ---- attributes ----
    @property
    def attrs(self) -> Dict[str, object]:
        if self._attrs is None:
            out = {}
            for t, b in self.msgs:
                if t == 0x0C:
                    name, val = self.f._parse_attribute(b)
                    out[name] = val
                elif t == 0x15:
                    flags = b[1]
                    p = 2 + (2 if flags & 1 else 0)
                    if self.f._unpack_addr(b, p) != UNDEF:
                        raise H5Error("object with dense attribute storage (fractal heap): not supported")
            self._attrs = out
        return self._attrs

    # ---- datasets ----
    @property
    def is_dataset(self) -> bool:
        return self._first(0x08) is not None and self._first(0x03) is not None

    @property
    def shape(self):
        b = self._first(0x01)
        return _parse_dataspace(b, 0, self.f.L)

    def datatype(self) -> _Datatype:
        return _parse_datatype(self._first(0x03))[0]

    def read(self):
        """The whole dataset as a NumPy array (object array of str for variable-length strings)."""
        return self.f._read_dataset(self)


class H5File:
    """Read-only view of an HDF5 file held in memory."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        base = -1
        for start in (0, 512, 1024, 2048, 4096):
            if b[start:start + 8] == b"\x89HDF\r\n\x1a\n":
                base = start
                break
        if base < 0:
            raise H5Error("not an HDF5 file")
        ver = b[base + 8]
        self._objs: Dict[int, _Object] = {}
        self._gcol: Dict[int, Dict[int, bytes]] = {}
        if ver in (0, 1):
            self.O, self.L = b[base + 13], b[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            self.base = self._unpack_addr(b, p)
            p += 4 * self.O
            # root group symbol-table entry
            root_header = self._unpack_addr(b, p + self.O)
            self.root = self.object_at(root_header)
        elif ver in (2, 3):
            self.O, self.L = b[base + 9], b[base + 10]
            self.base = self._unpack_addr(b, base + 12)
            root_header = self._unpack_addr(b, base + 12 + 3 * self.O)
            self.root = self.object_at(root_header)
        else:
            raise H5Error(f"superblock version {ver}")

    # ---- primitives ----
    def _unpack_addr(self, b: bytes, p: int) -> int:
        return int.from_bytes(b[p:p + self.O], "little")

    def _unpack_len(self, b: bytes, p: int) -> int:
        return int.from_bytes(b[p:p + self.L], "little")

    def object_at(self, addr: int) -> _Object:
        o = self._objs.get(addr)
        if o is None:
            o = _Object(self, addr)
            self._objs[addr] = o
        return o

    def __getitem__(self, path: str) -> _Object:
        return self.root[path]

    # ---- object headers ----
    def _read_object_header(self, addr: int, msgs: List[Tuple[int, bytes]]) -> None:
        b = self.buf
        a = addr + self.base
        if b[a:a + 4] == b"OHDR":
            self._read_object_header_v2(a, msgs)
            return
        if b[a] != 1:
            raise H5Error(f"object header version {b[a]} at {addr}")
        nmsg = struct.unpack_from("<H", b, a + 2)[0]
        hsize = struct.unpack_from("<I", b, a + 8)[0]
        blocks = [(a + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                p += 8 + msize
                seen += 1
                if mtype == 0x10:
                    blocks.append((self._unpack_addr(body, 0) + self.base, self._unpack_len(body, self.O)))
                elif mtype != 0:
                    if mflags & 2:
                        raise H5Error("shared header messages are not supported by this reader")
                    msgs.append((mtype, body))

    def _read_object_header_v2(self, a: int, msgs: List[Tuple[int, bytes]]) -> None:
        b = self.buf
        flags = b[a + 5]
        p = a + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nb = 1 << (flags & 3)
        size0 = int.from_bytes(b[p:p + nb], "little")
        p += nb
        track = bool(flags & 0x04)
        blocks = [(p, size0)]
        while blocks:
            p, size = blocks.pop(0)
            end = p + size
            while p + 4 + (2 if track else 0) <= end:
                mtype = b[p]
                msize = struct.unpack_from("<H", b, p + 1)[0]
                mflags = b[p + 3]
                p += 4 + (2 if track else 0)
                body = b[p:p + msize]
                p += msize
                if mtype == 0x10:
                    ca, cl = self._unpack_addr(body, 0) + self.base, self._unpack_len(body, self.O)
                    blocks.append((ca + 4, cl - 8))          # skip "OCHK", drop the checksum
                elif mtype != 0:
                    if mflags & 2:
                        raise H5Error("shared header messages are not supported by this reader")
                    msgs.append((mtype, body))

    # ---- groups ----
    def _heap_string(self, heap_addr: int, off: int) -> str:
        b = self.buf
        a = heap_addr + self.base
        if b[a:a + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        data = self._unpack_addr(b, a + 8 + 2 * self.L) + self.base
        e = b.index(b"\0", data + off)
        return b[data + off:e].decode("utf-8")

    def _walk_group_btree(self, node: int, heap: int, out: Dict[str, int]) -> None:
        b = self.buf
        a = node + self.base
        if b[a:a + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, a + 6)[0]
            p = a + 8
            for _ in range(n):
                name_off = self._unpack_addr(b, p)
                header = self._unpack_addr(b, p + self.O)
                out[self._heap_string(heap, name_off)] = header
                p += 2 * self.O + 24
            return
        if b[a:a + 4] != b"TREE":
            raise H5Error("bad group B-tree signature")
        used = struct.unpack_from("<H", b, a + 6)[0]
        p = a + 8 + 2 * self.O
        for i in range(used):
            p += self.L                                  # key i
            self._walk_group_btree(self._unpack_addr(b, p), heap, out)
            p += self.O

    def _parse_link(self, body: bytes) -> Tuple[str, Optional[int]]:
        flags = body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nb = 1 << (flags & 3)
        n = int.from_bytes(body[p:p + nb], "little")
        p += nb
        name = body[p:p + n].decode("utf-8")
        p += n
        return name, (self._unpack_addr(body, p) if ltype == 0 else None)

    # ---- global heap / variable-length data ----
    def _gheap_object(self, coll: int, index: int) -> bytes:
        objs = self._gcol.get(coll)
        if objs is None:
            b = self.buf
            a = coll + self.base
            if b[a:a + 4] != b"GCOL":
                raise H5Error("bad global heap signature")
            size = self._unpack_len(b, a + 8)
            p, end = a + 8 + self.L, a + size
            objs = {}
            while p + 8 + self.L <= end:
                idx = struct.unpack_from("<H", b, p)[0]
                osz = self._unpack_len(b, p + 8)
                if idx == 0:
                    break
                objs[idx] = b[p + 8 + self.L:p + 8 + self.L + osz]
                p += 8 + self.L + ((osz + 7) & ~7)
            self._gcol[coll] = objs
        return objs[index]

    def _decode_vlen(self, raw: bytes, count: int, t: _Datatype):
        out = np.empty(count, dtype=object)
        step = 4 + self.O + 4
        for i in range(count):
            p = i * step
            n = struct.unpack_from("<I", raw, p)[0]
            coll = self._unpack_addr(raw, p + 4)
            idx = struct.unpack_from("<I", raw, p + 4 + self.O)[0]
            if coll == 0 or (n == 0 and idx == 0):
                out[i] = "" if t.vlen_str else np.empty(0, dtype=t.vlen_base.dtype)
                continue
            data = self._gheap_object(coll, idx)
            if t.vlen_str:
                out[i] = data[:n].decode("utf-8")
            else:
                out[i] = np.frombuffer(data, dtype=t.vlen_base.dtype, count=n).copy()
        return out

    def _finish(self, raw: bytes, shape, t: _Datatype):
        count = 1
        for s in shape:
            count *= s
        if t.cls == 9:
            arr = self._decode_vlen(raw, count, t)
        elif t.cls == 3:
            a = np.frombuffer(raw, dtype=t.dtype, count=count)
            arr = np.array([x.rstrip(b"\0 ").decode("utf-8") if t.strpad != 1 else x.split(b"\0")[0].decode("utf-8")
                            for x in a], dtype=object)
        else:
            arr = np.frombuffer(raw, dtype=t.dtype, count=count).copy()
            if t.enum_names is not None and set(t.enum_names) == {"FALSE", "TRUE"}:
                true_val = t.enum_values[t.enum_names.index("TRUE")]
                arr = arr == true_val
            elif arr.dtype.byteorder == ">":
                arr = arr.astype(arr.dtype.newbyteorder("<"))
        return arr.reshape(shape)

    # ---- attributes ----
    def _parse_attribute(self, body: bytes):
        version = body[0]
        nsize, tsize, ssize = struct.unpack_from("<HHH", body, 2)
        p = 8 + (1 if version == 3 else 0)
        pad = (lambda n: (n + 7) & ~7) if version == 1 else (lambda n: n)
        name = body[p:p + nsize].split(b"\0")[0].decode("utf-8")
        p += pad(nsize)
        t, _ = _parse_datatype(body, p)
        p += pad(tsize)
        shape = _parse_dataspace(body, p, self.L)
        p += pad(ssize)
        if shape is None:
            return name, None
        count = 1
        for s in shape:
            count *= s
        nbytes = count * (4 + self.O + 4 if t.cls == 9 else t.size)
        val = self._finish(body[p:p + nbytes], shape, t)
        if shape == ():
            val = val.item() if val.dtype != object else val.reshape(-1)[0]
        return name, val

    # ---- datasets ----
    def _filters(self, obj: _Object) -> List[Tuple[int, Tuple[int, ...]]]:
        b = obj._first(0x0B)
        if b is None:
            return []
        version, n = b[0], b[1]
        p = 8 if version == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", b, p)[0]
            if version == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", b, p + 2)[0]
                p += 4
            else:
                nlen = 0
                p += 2
            flags, ncv = struct.unpack_from("<HH", b, p)
            p += 4
            p += ((nlen + 7) & ~7) if version == 1 else nlen
            cvals = struct.unpack_from(f"<{ncv}I", b, p)
            p += 4 * ncv
            if version == 1 and ncv % 2:
                p += 4
            out.append((fid, cvals))
        return out

    @staticmethod
    def _unfilter(data: bytes, filters, mask: int, elsize: int) -> bytes:
        for i in range(len(filters) - 1, -1, -1):
            if mask & (1 << i):
                continue
            fid, cvals = filters[i]
            if fid == 1:
                data = zlib.decompress(data)
            elif fid == 2:
                es = cvals[0] if cvals else elsize
                n = len(data) // es
                a = np.frombuffer(data, dtype=np.uint8, count=n * es).reshape(es, n)
                data = a.T.tobytes() + data[n * es:]
            elif fid == 3:
                data = data[:-4]
            else:
                raise H5Error(f"HDF5 filter {fid} is not supported (deflate, shuffle and fletcher32 are)")
        return data

    def _read_dataset(self, obj: _Object):
        t = obj.datatype()
        shape = obj.shape
        if shape is None:
            return None
        lay = obj._first(0x08)
        version = lay[0]
        elsize = 4 + self.O + 4 if t.cls == 9 else t.size
        count = 1
        for s in shape:
            count *= s
        if version == 3:
            cls = lay[1]
            if cls == 0:
                n = struct.unpack_from("<H", lay, 2)[0]
                raw = lay[4:4 + n]
            elif cls == 1:
                addr = self._unpack_addr(lay, 2)
                size = self._unpack_len(lay, 2 + self.O)
                raw = b"\0" * (count * elsize) if addr == UNDEF else self.buf[addr + self.base:addr + self.base + size]
            elif cls == 2:
                nd = lay[2]
                btree = self._unpack_addr(lay, 3)
                cdims = struct.unpack_from(f"<{nd}I", lay, 3 + self.O)
                raw = self._read_chunked(obj, btree, shape, cdims[:-1], elsize)
            else:
                raise H5Error(f"data layout class {cls}")
        elif version in (1, 2):
            nd, cls = lay[1], lay[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr = self._unpack_addr(lay, p)
                p += self.O
            dims = struct.unpack_from(f"<{nd}I", lay, p)
            p += 4 * nd
            if cls == 0:
                n = struct.unpack_from("<I", lay, p)[0]
                raw = lay[p + 4:p + 4 + n]
            elif cls == 1:
                raw = self.buf[addr + self.base:addr + self.base + count * elsize]
            else:
                raw = self._read_chunked(obj, addr, shape, dims[:-1], elsize)
        else:
            raise H5Error(f"data layout message version {version} (written with a newer libver): not supported")
        return self._finish(raw, shape, t)

    def _read_chunked(self, obj: _Object, btree: int, shape, cdims, elsize: int) -> bytes:
        filters = self._filters(obj)
        rank = len(shape)
        out = np.zeros(tuple(shape) + (elsize,), dtype=np.uint8)
        if btree == UNDEF:
            return out.tobytes()
        csize = elsize
        for c in cdims:
            csize *= c

        def walk(node: int):
            b = self.buf
            a = node + self.base
            if b[a:a + 4] != b"TREE":
                raise H5Error("bad chunk B-tree signature")
            level = b[a + 5]
            used = struct.unpack_from("<H", b, a + 6)[0]
            p = a + 8 + 2 * self.O
            ksize = 8 + 8 * (rank + 1)
            for _ in range(used):
                nbytes, mask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from(f"<{rank}Q", b, p + 8)
                child = self._unpack_addr(b, p + ksize)
                p += ksize + self.O
                if level > 0:
                    walk(child)
                    continue
                data = self._unfilter(b[child + self.base:child + self.base + nbytes], filters, mask, elsize)
                chunk = np.frombuffer(data, dtype=np.uint8, count=csize).reshape(tuple(cdims) + (elsize,))
                sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                sl_in = tuple(slice(0, so.stop - so.start) for so in sl_out)
                out[sl_out] = chunk[sl_in]

        walk(btree)
        return out.tobytes()


# ---------------------------------------------------------------------------------------------------------------
# anndata layout
# ---------------------------------------------------------------------------------------------------------------
class MiniAnnData:
    """The part of an AnnData object that ``wasserstein_distance`` touches."""

    def __init__(self, X, obs: pd.DataFrame, var: pd.DataFrame, obsm: Dict[str, np.ndarray], uns: Dict[str, object]):
        self.X, self.obs, self.var, self.obsm, self.uns = X, obs, var, obsm, uns

    @property
    def var_names(self):
        return self.var.index

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def n_obs(self):
        return self.obs.shape[0]

    @property
    def n_vars(self):
        return self.var.shape[0]

    @property
    def shape(self):
        return (self.n_obs, self.n_vars)

    def __getitem__(self, key) -> "MiniAnnData":
        """``adata[:, names]`` (Trajectory.py:290) and ``adata[rows]``: a view-like copy."""
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        if not (isinstance(cols, slice) and cols == slice(None)):
            cols = [self.var.index.get_loc(c) if not isinstance(c, (int, np.integer)) else int(c)
                    for c in (cols if not isinstance(cols, str) else [cols])]
        ri = rows if isinstance(rows, slice) else np.asarray(rows)
        X = self.X[ri][:, cols] if self.X is not None else None
        obs = self.obs.iloc[ri] if not isinstance(ri, slice) or ri != slice(None) else self.obs
        var = self.var.iloc[cols] if not isinstance(cols, slice) else self.var
        obsm = {k: v[ri] for k, v in self.obsm.items()}
        return MiniAnnData(X, obs, var, obsm, dict(self.uns))

    def to_df(self) -> pd.DataFrame:
        X = self.X.toarray() if hasattr(self.X, "toarray") else self.X
        return pd.DataFrame(X, index=self.obs.index, columns=self.var.index)

    def __repr__(self):
        return (f"MiniAnnData object with n_obs x n_vars = {self.n_obs} x {self.n_vars}\n    obs: {list(self.obs.columns)}"
                f"\n    var: {list(self.var.columns)}\n    obsm: {list(self.obsm)}\n    uns: {list(self.uns)}")


def _as_str(x):
    return x.decode("utf-8") if isinstance(x, bytes) else x


def _read_array(obj: _Object, f: H5File):
    """A dataset or an encoded group (categorical, csr/csc matrix, nullable array) as an in-memory value."""
    if obj.is_dataset:
        arr = obj.read()
        cats_ref = obj.attrs.get("categories")          # anndata 0.7: reference to obs/__categories/<name>
        if cats_ref is not None and obj.datatype().cls == 0:
            ref = int(np.asarray(cats_ref).reshape(-1)[0])
            cats = f.object_at(ref).read()
            ordered = bool(f.object_at(ref).attrs.get("ordered", False))
            return pd.Categorical.from_codes(arr.astype(np.int64), categories=pd.Index(cats), ordered=ordered)
        return arr
    enc = _as_str(obj.attrs.get("encoding-type", ""))
    if enc == "categorical":
        codes = obj["codes"].read()
        cats = obj["categories"].read()
        return pd.Categorical.from_codes(codes.astype(np.int64), categories=pd.Index(cats),
                                         ordered=bool(obj.attrs.get("ordered", False)))
    if enc in ("csr_matrix", "csc_matrix"):
        import scipy.sparse as sps
        shape = tuple(int(s) for s in np.asarray(obj.attrs["shape"]).reshape(-1))
        cls = sps.csr_matrix if enc == "csr_matrix" else sps.csc_matrix
        return cls((obj["data"].read(), obj["indices"].read(), obj["indptr"].read()), shape=shape)
    if enc in ("nullable-integer", "nullable-boolean"):
        vals, mask = obj["values"].read(), obj["mask"].read().astype(bool)
        if enc == "nullable-integer":
            return pd.arrays.IntegerArray(np.where(mask, 0, vals), mask)
        return pd.arrays.BooleanArray(np.where(mask, False, vals).astype(bool), mask)
    if enc == "dataframe" or "_index" in obj.attrs:
        return _read_dataframe(obj, f)
    return _read_mapping(obj, f)


def _read_mapping(obj: _Object, f: H5File) -> Dict[str, object]:
    out = {}
    for name in obj.keys():
        child = obj[name]
        try:
            val = _read_array(child, f)
        except H5Error:
            continue
        if isinstance(val, np.ndarray) and val.shape == ():
            val = val.item()
        out[name] = val
    return out


def _read_dataframe(obj: _Object, f: H5File) -> pd.DataFrame:
    if obj.is_dataset:
        # anndata < 0.7: a compound dataset, first field = index
        rec = obj.read()
        names = list(rec.dtype.names)
        df = pd.DataFrame({n: rec[n] for n in names[1:]})
        idx = rec[names[0]]
        df.index = pd.Index([_as_str(x) for x in idx], name=None)
        return df
    attrs = obj.attrs
    index_key = _as_str(attrs.get("_index", "_index"))
    order = attrs.get("column-order")
    cols = [] if order is None else [_as_str(c) for c in np.asarray(order).reshape(-1)]
    data = {}
    for c in cols:
        data[c] = _read_array(obj[c], f)
    index = obj[index_key].read()
    df = pd.DataFrame(data, index=pd.Index([_as_str(x) for x in index]))
    if index_key not in ("_index", "index", "__index_level_0__"):
        df.index.name = index_key
    return df


def read_h5ad(path: str) -> MiniAnnData:
    """Read ``path`` with the built-in HDF5 reader."""
    f = H5File(path)
    root = f.root
    obs = _read_dataframe(root["obs"], f) if "obs" in root else pd.DataFrame()
    var = _read_dataframe(root["var"], f) if "var" in root else pd.DataFrame()
    X = _read_array(root["X"], f) if "X" in root else None
    obsm = {}
    if "obsm" in root and root["obsm"].is_group:
        for k in root["obsm"].keys():
            obsm[k] = _read_array(root["obsm"][k], f)
    uns = _read_mapping(root["uns"], f) if "uns" in root and root["uns"].is_group else {}
    return MiniAnnData(X, obs, var, obsm, uns)


def load_h5ad(path: str):
    """``pilotpy.tl.load_h5ad`` (Trajectory.py:121-137): the AnnData object of ``path``, or -- like the reference -- a
    printed hint and ``None`` when there is no such file.  anndata when it is importable, the built-in reader
    otherwise."""
    import os
    if os.path.isfile(path):
        try:
            import anndata  # type: ignore
        except ImportError:
            return read_h5ad(path)
        return anndata.read_h5ad(path)
    print("There is no such data, check the path or name")
    return None
