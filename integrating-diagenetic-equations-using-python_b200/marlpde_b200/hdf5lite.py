"""Minimal pure-Python HDF5 reader/writer (no libhdf5, no h5py in this image).

Why it exists: the reference stores its results with h5py
(`marlpde/Evolve_scenario.py:170-178`: datasets `solutions`, `times`,
`event_0..event_6`, all parameters as root attributes) and its regression tests
read HDF5 fixtures (`tests/Regression_test/test_regression.py:9-18`).  The
"HDF5 output layout unchanged" requirement therefore needs a writer, and the
parity tests need a reader for the reference's chunked+deflate fixtures.

Scope (only what those two uses need):

reader  superblock v0/v1, v1 object headers (+continuations), old-style groups
        (v1 B-tree + local heap + SNOD) and compact new-style groups (link
        messages), dataspace v1/v2, fixed/float/string/vlen-string/enum
        datatypes, contiguous / compact / chunked (v1 chunk B-tree) layouts,
        deflate + shuffle filters, attribute messages v1-v3.
writer  superblock v0, one root group (symbol table), contiguous little-endian
        datasets of float64/int64/int8, root attributes (float, int, bool as
        h5py-style enum, vlen UTF-8 str, numeric arrays).

The facade mimics the sliver of h5py the reference touches: ``File(path, mode)``
as a context manager, ``.get(name)``/``[name]`` -> array-like with ``.shape`` and
numpy indexing, ``.create_dataset(name, data=)``, ``.attrs.update(dict)``.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# --------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != _SIG:
            raise OSError("not an HDF5 file (bad signature)")
        ver = buf[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"superblock version {ver} not supported")
        if buf[13] != 8 or buf[14] != 8:
            raise NotImplementedError("only 8-byte offsets/lengths supported")
        off = 24 if ver == 0 else 28
        self.base = struct.unpack_from("<Q", buf, off)[0]
        root_entry = off + 32
        self.root_header = struct.unpack_from("<Q", buf, root_entry + 8)[0]

    # -- object headers ------------------------------------------------------
    def messages(self, addr: int):
        """Yield (type, flags, payload-bytes) of a version-1 object header."""
        b = self.b
        if b[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("version-2 object headers not supported")
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, pos)
                payload = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", payload, 0)
                    blocks.append((caddr, clen))
                out.append((mtype, mflags, payload))
        return out

    # -- groups ----------------------------------------------------------------
    def links(self, addr: int) -> dict:
        res = {}
        for mtype, _f, p in self.messages(addr):
            if mtype == 0x11:  # symbol table: old-style group
                btree, heap = struct.unpack_from("<QQ", p, 0)
                res.update(self._walk_group_btree(btree, heap))
            elif mtype == 0x06:  # link message: compact new-style group
                name, target = self._parse_link(p)
                if target is not None:
                    res[name] = target
        return res

    def _parse_link(self, p: bytes):
        ver, flags = p[0], p[1]
        pos = 2
        ltype = 0
        if flags & 0x08:
            ltype = p[pos]
            pos += 1
        if flags & 0x04:
            pos += 8
        if flags & 0x10:
            pos += 1
        lsz = 1 << (flags & 3)
        nlen = int.from_bytes(p[pos:pos + lsz], "little")
        pos += lsz
        name = p[pos:pos + nlen].decode("utf-8")
        pos += nlen
        if ltype != 0:
            return name, None
        return name, struct.unpack_from("<Q", p, pos)[0]

    def _heap_string(self, heap: int, off: int) -> str:
        b = self.b
        assert b[heap:heap + 4] == b"HEAP"
        data = struct.unpack_from("<Q", b, heap + 24)[0]
        s = data + off
        e = b.index(b"\x00", s)
        return b[s:e].decode("utf-8")

    def _walk_group_btree(self, node: int, heap: int) -> dict:
        b = self.b
        assert b[node:node + 4] == b"TREE", "bad group B-tree node"
        ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
        assert ntype == 0
        res = {}
        pos = node + 24
        for i in range(used):
            child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
            if level > 0:
                res.update(self._walk_group_btree(child, heap))
            else:
                assert b[child:child + 4] == b"SNOD"
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                for k in range(nsym):
                    e = child + 8 + 40 * k
                    noff, ohdr = struct.unpack_from("<QQ", b, e)
                    res[self._heap_string(heap, noff)] = ohdr
        return res

    # -- datatypes ------------------------------------------------------------------
    def parse_dtype(self, p: bytes, pos: int = 0):
        """Return (descr, nbytes_consumed). descr is a numpy dtype, or a tuple
        ('vlen_str',) / ('enum', base_dtype, {value: name})."""
        cv = p[pos]
        cls, ver = cv & 0x0F, cv >> 4
        bits = p[pos + 1] | (p[pos + 2] << 8) | (p[pos + 3] << 16)
        size = struct.unpack_from("<I", p, pos + 4)[0]
        body = pos + 8
        order = ">" if (bits & 1) else "<"
        if cls == 0:  # fixed point
            signed = bool(bits & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), 8 + 4
        if cls == 1:  # float
            return np.dtype(f"{order}f{size}"), 8 + 12
        if cls == 3:  # fixed-length string
            return np.dtype(f"S{size}"), 8
        if cls == 9:  # variable length
            vtype = bits & 0x0F
            base, n = self.parse_dtype(p, body)
            if vtype == 1:
                return ("vlen_str",), 8 + n
            return ("vlen", base), 8 + n
        if cls == 8:  # enum
            nmemb = bits & 0xFFFF
            base, n = self.parse_dtype(p, body)
            q = body + n
            names = []
            for _ in range(nmemb):
                e = p.index(b"\x00", q)
                names.append(p[q:e].decode())
                ln = e - q + 1
                q += ln if ver >= 3 else (ln + 7) // 8 * 8
            vals = np.frombuffer(p, dtype=base, count=nmemb, offset=q)
            q += nmemb * base.itemsize
            return ("enum", base, {int(v): nm for v, nm in zip(vals, names)}), q - pos
        raise NotImplementedError(f"datatype class {cls}")

    @staticmethod
    def parse_space(p: bytes):
        ver, rank, flags = p[0], p[1], p[2]
        if ver == 1:
            pos = 8
        elif ver == 2:
            pos = 4
            if p[3] == 2:  # null dataspace
                return None
        else:
            raise NotImplementedError(f"dataspace version {ver}")
        return tuple(struct.unpack_from(f"<{rank}Q", p, pos)) if rank else ()

    # -- raw data ---------------------------------------------------------------------
    def _decode(self, raw: bytes, dt, shape):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, np.dtype):
            a = np.frombuffer(raw, dtype=dt, count=n).reshape(shape)
            return a.astype(dt.newbyteorder("=")) if dt.kind in "fiu" else a
        if dt[0] == "vlen_str":
            out = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                out.append(self._gheap_object(gaddr, gidx)[:ln].decode("utf-8"))
            return out[0] if shape == () else np.array(out, dtype=object).reshape(shape)
        if dt[0] == "enum":
            base, table = dt[1], dt[2]
            a = np.frombuffer(raw, dtype=base, count=n).reshape(shape)
            if set(table.values()) == {"FALSE", "TRUE"}:
                return a.astype(bool)
            return a
        raise NotImplementedError(str(dt))

    def _gheap_object(self, addr: int, index: int) -> bytes:
        b = self.b
        assert b[addr:addr + 4] == b"GCOL"
        csize = struct.unpack_from("<Q", b, addr + 8)[0]
        pos, end = addr + 16, addr + csize
        while pos + 16 <= end:
            idx, _ref, _r, osz = struct.unpack_from("<HHIQ", b, pos)
            if idx == index:
                return b[pos + 16:pos + 16 + osz]
            if idx == 0:
                break
            pos += 16 + (osz + 7) // 8 * 8
        raise KeyError(f"global heap object {index} not found")

    def attributes(self, addr: int) -> dict:
        res = {}
        for mtype, _f, p in self.messages(addr):
            if mtype != 0x0C:
                continue
            ver = p[0]
            nsz, tsz, ssz = struct.unpack_from("<HHH", p, 2)
            pos = 8 if ver < 3 else 9
            pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
            name = p[pos:pos + nsz].split(b"\x00")[0].decode("utf-8")
            pos += pad(nsz)
            dt, _ = self.parse_dtype(p, pos)
            pos += pad(tsz)
            shape = self.parse_space(p[pos:pos + ssz])
            pos += pad(ssz)
            if shape is None:
                res[name] = None
                continue
            val = self._decode(p[pos:], dt, shape)
            if isinstance(val, np.ndarray) and val.shape == ():
                val = val[()]
            res[name] = val
        return res

    def dataset(self, addr: int) -> np.ndarray:
        dt = shape = layout = None
        filters = []
        for mtype, _f, p in self.messages(addr):
            if mtype == 0x01:
                shape = self.parse_space(p)
            elif mtype == 0x03:
                dt, _ = self.parse_dtype(p)
            elif mtype == 0x08:
                layout = p
            elif mtype == 0x0B:
                filters = self._parse_filters(p)
        if dt is None or shape is None or layout is None:
            raise OSError("object is not a dataset")
        if layout[0] != 3:
            raise NotImplementedError(f"data layout version {layout[0]}")
        cls = layout[1]
        n = int(np.prod(shape)) if shape else 1
        isz = dt.itemsize if isinstance(dt, np.dtype) else 16
        if cls == 1:  # contiguous
            daddr, dsize = struct.unpack_from("<QQ", layout, 2)
            raw = b"" if daddr == _UNDEF else self.b[daddr:daddr + dsize]
            if len(raw) < n * isz:
                raw = raw + b"\x00" * (n * isz - len(raw))
            return self._decode(raw, dt, shape)
        if cls == 0:  # compact
            dsize = struct.unpack_from("<H", layout, 2)[0]
            return self._decode(layout[4:4 + dsize], dt, shape)
        if cls == 2:  # chunked, v1 B-tree index
            nd = layout[2]
            btree = struct.unpack_from("<Q", layout, 3)[0]
            cdims = struct.unpack_from(f"<{nd}I", layout, 11)
            chunk_shape = cdims[:-1]
            assert isinstance(dt, np.dtype)
            out = np.zeros(shape, dtype=dt.newbyteorder("="))
            if btree != _UNDEF:
                for offs, caddr, csize, mask in self._walk_chunk_btree(btree, nd):
                    raw = self.b[caddr:caddr + csize]
                    for k, (fid, cdata) in reversed(list(enumerate(filters))):
                        if mask & (1 << k):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            esz = cdata[0] if cdata else dt.itemsize
                            a = np.frombuffer(raw, dtype=np.uint8)
                            raw = a.reshape(esz, -1).T.tobytes()
                        else:
                            raise NotImplementedError(f"HDF5 filter id {fid}")
                    chunk = np.frombuffer(raw, dtype=dt).reshape(chunk_shape)
                    sl_out, sl_in = [], []
                    for o, c, s in zip(offs, chunk_shape, shape):
                        e = min(o + c, s)
                        sl_out.append(slice(o, e))
                        sl_in.append(slice(0, e - o))
                    out[tuple(sl_out)] = chunk[tuple(sl_in)]
            return out
        raise NotImplementedError(f"layout class {cls}")

    @staticmethod
    def _parse_filters(p: bytes):
        ver, nf = p[0], p[1]
        pos = 8 if ver == 1 else 2
        res = []
        for _ in range(nf):
            fid = struct.unpack_from("<H", p, pos)[0]
            pos += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", p, pos)[0]
                pos += 2
            _flags, ncd = struct.unpack_from("<HH", p, pos)
            pos += 4
            if nlen:
                pos += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = struct.unpack_from(f"<{ncd}I", p, pos)
            pos += 4 * ncd
            if ver == 1 and ncd % 2:
                pos += 4
            res.append((fid, cd))
        return res

    def _walk_chunk_btree(self, node: int, nd: int):
        b = self.b
        assert b[node:node + 4] == b"TREE", "bad chunk B-tree node"
        ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
        assert ntype == 1
        keysz = 8 + 8 * nd
        pos = node + 24
        for i in range(used):
            k = pos + i * (keysz + 8)
            csize, mask = struct.unpack_from("<II", b, k)
            offs = struct.unpack_from(f"<{nd}Q", b, k + 8)[:-1]
            child = struct.unpack_from("<Q", b, k + keysz)[0]
            if level > 0:
                yield from self._walk_chunk_btree(child, nd)
            else:
                yield offs, child, csize, mask


# --------------------------------------------------------------------------- writer
def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dt_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + \
            struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + \
            struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError(f"unsupported dtype {dt}")


def _bool_enum_msg() -> bytes:
    base = _dt_msg(np.dtype("i1"))
    names = _pad8(b"FALSE\x00") + _pad8(b"TRUE\x00")
    return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])


def _vlen_str_msg() -> bytes:
    # class 9, version 1; bits: type=string(1), pad=null-terminate(0), cset=UTF-8(1)
    base = struct.pack("<BBBBI", 0x13, 0x10, 0, 0, 1)  # 1-byte UTF-8 string, nullterm
    return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + base


def _space_msg(shape) -> bytes:
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _msg(mtype: int, payload: bytes, flags: int = 0) -> bytes:
    payload = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(payload), flags) + payload


def _object_header(msgs) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


class _Writer:
    """Accumulates datasets/attributes, then serialises a complete file."""

    def __init__(self):
        self.datasets = {}
        self.attrs = {}

    def serialise(self) -> bytes:
        names = sorted(self.datasets, key=lambda s: s.encode("utf-8"))
        # ---- fixed front: superblock (96) ; root header follows at 96 ------------
        gheap_objs = []  # vlen string payloads

        def attr_msg(name: str, value) -> bytes:
            nm = name.encode("utf-8") + b"\x00"
            if isinstance(value, (bool, np.bool_)):
                dtm, spm, data = _bool_enum_msg(), _space_msg(()), bytes([1 if value else 0])
            elif isinstance(value, str):
                gheap_objs.append(value.encode("utf-8"))
                dtm, spm = _vlen_str_msg(), _space_msg(())
                data = ("VLEN", len(gheap_objs))  # patched once the heap address is known
            else:
                arr = np.asarray(value)
                if arr.dtype.kind == "b":
                    dtm, arr = _bool_enum_msg(), arr.astype("i1")
                elif arr.dtype.kind in "iu":
                    arr = arr.astype("<i8")
                    dtm = _dt_msg(arr.dtype)
                elif arr.dtype.kind == "f":
                    arr = arr.astype("<f8")
                    dtm = _dt_msg(arr.dtype)
                else:
                    raise TypeError(f"attribute {name!r}: unsupported value {value!r}")
                spm, data = _space_msg(arr.shape), arr.tobytes()
            head = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(spm))
            return head, _pad8(nm) + _pad8(dtm) + _pad8(spm), data

        attr_parts = [attr_msg(k, v) for k, v in self.attrs.items()]

        # ---- layout planning -------------------------------------------------------
        # local heap strings: offset 0 = "" (8 bytes), then the names
        heap_data = bytearray(b"\x00" * 8)
        name_off = {}
        for nm in names:
            name_off[nm] = len(heap_data)
            heap_data += _pad8(nm.encode("utf-8") + b"\x00")
        heap_data += b"\x00" * 16  # free block (size filled below)
        LEAF_K, INT_K = 64, 16
        if len(names) > 2 * LEAF_K:
            raise NotImplementedError("too many datasets for one SNOD")

        root_attr_len = sum(8 + len(_pad8(h + b + (b"\x00" * 16 if isinstance(d, tuple) else d)))
                            for h, b, d in attr_parts)
        root_hdr_len = 16 + (8 + 16) + root_attr_len
        pos = 96
        root_addr = pos
        pos += root_hdr_len
        btree_addr = pos
        pos += 24 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8
        heap_addr = pos
        pos += 32
        heap_data_addr = pos
        pos += len(heap_data)
        snod_addr = pos
        pos += 8 + 40 * 2 * LEAF_K
        gheap_addr = pos
        gheap = b""
        if gheap_objs:
            body = b""
            for i, o in enumerate(gheap_objs, start=1):
                body += struct.pack("<HHIQ", i, 1, 0, len(o)) + _pad8(o)
            total = 16 + len(body) + 16
            total = max(total, 4096)
            free = total - 16 - len(body)
            gheap = b"GCOL" + struct.pack("<B3xQ", 1, total) + body + \
                struct.pack("<HHIQ", 0, 0, 0, free) + b"\x00" * (free - 16)
            pos += len(gheap)
        # dataset headers + data
        ds_hdr_addr, ds_data_addr, ds_hdr_bytes, ds_data_bytes = {}, {}, {}, {}
        for nm in names:
            arr = self.datasets[nm]
            raw = arr.tobytes()
            ds_hdr_addr[nm] = pos
            fill = struct.pack("<BBBB", 2, 2, 2, 0)  # v2, alloc late, write if-set, undefined
            hdr_len = 16 + sum(len(_msg(0, p)) for p in (
                _space_msg(arr.shape), _dt_msg(arr.dtype), fill, b"\x00" * 18))
            daddr = pos + hdr_len if raw else _UNDEF
            layout = struct.pack("<BBQQ", 3, 1, daddr, len(raw))
            hdr = _object_header([
                _msg(0x01, _space_msg(arr.shape)),
                _msg(0x03, _dt_msg(arr.dtype), flags=1),
                _msg(0x05, fill),
                _msg(0x08, layout),
            ])
            assert len(hdr) == hdr_len
            ds_hdr_bytes[nm], ds_data_bytes[nm] = hdr, _pad8(raw)
            ds_data_addr[nm] = daddr
            pos += hdr_len + len(ds_data_bytes[nm])
        eof = pos

        # ---- emit ------------------------------------------------------------------
        out = bytearray()
        out += _SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        out += struct.pack("<HHI", LEAF_K, INT_K, 0)
        out += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        out += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
        assert len(out) == 96
        msgs = [_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))]
        for head, body, data in attr_parts:
            if isinstance(data, tuple):
                idx = data[1]
                data = struct.pack("<IQI", len(gheap_objs[idx - 1]), gheap_addr, idx)
            msgs.append(_msg(0x0C, head + body + data))
        hdr = _object_header(msgs)
        assert len(hdr) == root_hdr_len, (len(hdr), root_hdr_len)
        out += hdr
        # group B-tree: one leaf entry
        bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, _UNDEF, _UNDEF)
        bt += struct.pack("<Q", 0)
        if names:
            bt += struct.pack("<QQ", snod_addr, name_off[names[-1]])
        bt += b"\x00" * (24 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8 - len(bt))
        out += bt
        free_off = len(heap_data) - 16
        struct.pack_into("<QQ", heap_data, free_off, 1, 16)  # next-free=1 (none), size
        out += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr)
        out += heap_data
        sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for nm in names:
            sn += struct.pack("<QQII16x", name_off[nm], ds_hdr_addr[nm], 0, 0)
        sn += b"\x00" * (8 + 40 * 2 * LEAF_K - len(sn))
        out += sn
        out += gheap
        for nm in names:
            assert len(out) == ds_hdr_addr[nm]
            out += ds_hdr_bytes[nm] + ds_data_bytes[nm]
        assert len(out) == eof
        return bytes(out)


# --------------------------------------------------------------------------- facade
class Dataset:
    """Array-like view returned by File.get / File[name] (h5py.Dataset look-alike)."""

    def __init__(self, name: str, data: np.ndarray, attrs=None):
        self.name = name
        self._data = data
        self.attrs = attrs or {}

    shape = property(lambda self: self._data.shape)
    dtype = property(lambda self: self._data.dtype)
    ndim = property(lambda self: self._data.ndim)
    size = property(lambda self: self._data.size)

    def __getitem__(self, idx):
        return self._data[idx]

    def __len__(self):
        return len(self._data)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)


class _Attrs(dict):
    def __init__(self, owner, initial=()):
        super().__init__(initial)
        self._owner = owner

    def _check(self):
        if self._owner is not None and self._owner.mode == "r":
            raise OSError("file opened read-only")

    def __setitem__(self, k, v):
        self._check()
        super().__setitem__(k, v)

    def update(self, *a, **kw):
        self._check()
        super().update(*a, **kw)


class File:
    """h5py.File look-alike for modes 'r' and 'w'."""

    def __init__(self, path, mode: str = "r"):
        if mode not in ("r", "w"):
            raise ValueError("hdf5lite.File supports modes 'r' and 'w' only")
        self.filename = str(path)
        self.mode = mode
        self._closed = False
        if mode == "r":
            with open(path, "rb") as fh:
                self._r = _Reader(fh.read())
            self._links = self._r.links(self._r.root_header)
            self.attrs = _Attrs(None, self._r.attributes(self._r.root_header))
            self.attrs._owner = self
            self._cache = {}
        else:
            open(path, "wb").close()  # fail early like h5py on unwritable paths
            self._w = _Writer()
            self.attrs = _Attrs(self)

    # reading -----------------------------------------------------------------
    def keys(self):
        return list(self._links) if self.mode == "r" else list(self._w.datasets)

    def __contains__(self, name):
        return name in self.keys()

    def get(self, name, default=None):
        try:
            return self[name]
        except KeyError:
            return default

    def __getitem__(self, name):
        name = name.lstrip("/")
        if self.mode == "w":
            return Dataset(name, self._w.datasets[name])
        if name not in self._links:
            raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
        if name not in self._cache:
            addr = self._links[name]
            self._cache[name] = Dataset(name, self._r.dataset(addr), self._r.attributes(addr))
        return self._cache[name]

    # writing -----------------------------------------------------------------
    def create_dataset(self, name, data=None, shape=None, dtype=None):
        if self.mode != "w":
            raise OSError("file opened read-only")
        if data is None:
            data = np.zeros(shape, dtype=dtype or "f8")
        arr = np.array(data, dtype=dtype, copy=True, order="C")
        if arr.dtype.kind == "f":
            arr = arr.astype("<f8" if arr.dtype.itemsize == 8 else "<f4")
        elif arr.dtype.kind in "iu":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        elif arr.dtype.kind == "b":
            arr = arr.astype("i1")
        else:
            raise TypeError(f"dataset {name!r}: unsupported dtype {arr.dtype}")
        if name in self._w.datasets:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        self._w.datasets[name] = arr
        return Dataset(name, arr)

    def close(self):
        if self._closed:
            return
        if self.mode == "w":
            self._w.attrs = dict(self.attrs)
            blob = self._w.serialise()
            with open(self.filename, "wb") as fh:
                fh.write(blob)
        self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
