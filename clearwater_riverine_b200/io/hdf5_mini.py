"""Minimal read-only HDF5 reader (pure Python: struct + zlib + numpy).

Why this exists: the reference reads HEC-RAS output with h5py
(reference: src/clearwater_riverine/io/hdf.py:127-145); h5py/libhdf5 are not
available in the build image, and the parity fixtures (tests/golden/*.npz) have
to be extracted from the reference's own HEC-RAS plan files.

Supported subset = what HEC-RAS 6.x "classic" files use:
superblock v0/v1, object header v1 (+ continuation blocks), symbol-table groups
(v1 B-tree + SNOD + local heap), dataspace v1/v2, datatype classes
0 (integer), 1 (float), 3 (string), 6 (compound v1/v2/v3), layout message v3
(compact / contiguous / chunked via v1 chunk B-tree), filter pipeline v1/v2
with deflate (id 1) and shuffle (id 2), attribute messages v1/v2/v3.
Anything else raises NotImplementedError -- never a silent wrong answer.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(NotImplementedError):
    pass


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Node:
    """A group or dataset (object header address + parsed messages)."""

    def __init__(self, file: "File", addr: int, name: str):
        self.file = file
        self.addr = addr
        self.name = name
        self._msgs: Optional[List[Tuple[int, int, bytes]]] = None
        self._links: Optional[Dict[str, int]] = None
        self._attrs: Optional[Dict[str, object]] = None

    # -- object header -----------------------------------------------------
    def _messages(self) -> List[Tuple[int, int, bytes]]:
        if self._msgs is not None:
            return self._msgs
        buf = self.file.buf
        a = self.addr
        version = buf[a]
        if version != 1:
            raise HDF5FormatError(f"object header version {version} at {a:#x}")
        nmsg, = struct.unpack_from("<H", buf, a + 2)
        hsize, = struct.unpack_from("<I", buf, a + 8)
        blocks = [(a + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, pos)
                body = bytes(buf[pos + 8: pos + 8 + msize])
                pos += 8 + msize
                if mtype == 0x10:  # continuation
                    off, length = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off, length))
                out.append((mtype, mflags, body))
        self._msgs = out
        return out

    def _msg(self, mtype: int) -> Optional[bytes]:
        for t, _f, b in self._messages():
            if t == mtype:
                return b
        return None

    # -- group interface -----------------------------------------------------
    def is_group(self) -> bool:
        return self._msg(0x11) is not None

    def links(self) -> Dict[str, int]:
        if self._links is not None:
            return self._links
        st = self._msg(0x11)
        if st is None:
            raise KeyError(f"{self.name!r} is not a group")
        btree, heap = struct.unpack_from("<QQ", st, 0)
        heap_data = self.file._local_heap(heap)
        links: Dict[str, int] = {}
        self.file._walk_group_btree(btree, heap_data, links)
        self._links = links
        return links

    def keys(self):
        return list(self.links().keys())

    def __contains__(self, key: str) -> bool:
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        if path == () or path is Ellipsis:
            return self.read()
        if not isinstance(path, str):
            return self.read()[path]
        node = self
        for part in [p for p in path.split("/") if p]:
            links = node.links()
            if part not in links:
                raise KeyError(f"{part!r} not found under {node.name!r}")
            node = _Node(self.file, links[part], node.name.rstrip("/") + "/" + part)
        return node

    # -- dataset interface ---------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return _parse_dataspace(self._msg(0x01))

    @property
    def dtype(self) -> np.dtype:
        return _parse_datatype(self._msg(0x03), 0)[0]

    def read(self) -> np.ndarray:
        ds = self._msg(0x01)
        dt = self._msg(0x03)
        lay = self._msg(0x08)
        if ds is None or dt is None or lay is None:
            raise KeyError(f"{self.name!r} is not a dataset")
        shape = _parse_dataspace(ds)
        dtype, _ = _parse_datatype(dt, 0)
        filters = _parse_filters(self._msg(0x0B))
        buf = self.file.buf
        version = lay[0]
        if version != 3:
            raise HDF5FormatError(f"layout message version {version}")
        cls = lay[1]
        count = int(np.prod(shape)) if shape else 1
        if cls == 0:  # compact
            size, = struct.unpack_from("<H", lay, 2)
            raw = lay[4:4 + size]
            return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape).copy()
        if cls == 1:  # contiguous
            addr, size = struct.unpack_from("<QQ", lay, 2)
            if addr == _UNDEF:
                return np.zeros(shape, dtype=dtype)
            return np.frombuffer(buf, dtype=dtype, count=count, offset=addr).reshape(shape).copy()
        if cls == 2:  # chunked
            ndim = lay[2]
            btree, = struct.unpack_from("<Q", lay, 3)
            cdims = struct.unpack_from("<" + "I" * ndim, lay, 11)
            chunk_shape = tuple(cdims[:-1])
            if len(chunk_shape) != len(shape):
                raise HDF5FormatError("chunk rank mismatch")
            out = np.zeros(shape, dtype=dtype)
            if btree != _UNDEF:
                self.file._walk_chunk_btree(btree, ndim, chunk_shape, dtype, filters, out)
            return out
        raise HDF5FormatError(f"layout class {cls}")

    # -- attributes ----------------------------------------------------------
    @property
    def attrs(self) -> Dict[str, object]:
        if self._attrs is not None:
            return self._attrs
        out: Dict[str, object] = {}
        for t, _f, b in self._messages():
            if t != 0x0C:
                continue
            version = b[0]
            if version == 1:
                nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
                p = 8
                name = b[p:p + nsz].split(b"\0")[0].decode("ascii", "replace")
                p += _pad8(nsz)
                tbytes = b[p:p + tsz]
                p += _pad8(tsz)
                sbytes = b[p:p + ssz]
                p += _pad8(ssz)
            elif version in (2, 3):
                nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
                p = 8 + (1 if version == 3 else 0)
                name = b[p:p + nsz].split(b"\0")[0].decode("ascii", "replace")
                p += nsz
                tbytes = b[p:p + tsz]
                p += tsz
                sbytes = b[p:p + ssz]
                p += ssz
            else:
                raise HDF5FormatError(f"attribute message version {version}")
            dtype, _ = _parse_datatype(tbytes, 0)
            shape = _parse_dataspace(sbytes)
            count = int(np.prod(shape)) if shape else 1
            val = np.frombuffer(b, dtype=dtype, count=count, offset=p)
            out[name] = val.reshape(shape).copy() if shape else val[0]
        self._attrs = out
        return out


def _parse_dataspace(b: bytes) -> Tuple[int, ...]:
    version, rank, flags = b[0], b[1], b[2]
    if version == 1:
        p = 8
    elif version == 2:
        if b[3] == 2:  # null dataspace
            return (0,)
        p = 4
    else:
        raise HDF5FormatError(f"dataspace version {version}")
    return tuple(struct.unpack_from("<" + "Q" * rank, b, p))


def _parse_datatype(b: bytes, p: int) -> Tuple[np.dtype, int]:
    """Returns (numpy dtype, offset just past this datatype description)."""
    cv = b[p]
    cls, version = cv & 0x0F, cv >> 4
    bits0, bits1, _bits2 = b[p + 1], b[p + 2], b[p + 3]
    size, = struct.unpack_from("<I", b, p + 4)
    q = p + 8
    if cls == 0:  # fixed point
        order = ">" if bits0 & 1 else "<"
        signed = bool(bits0 & 8)
        return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), q + 4
    if cls == 1:  # float
        order = ">" if bits0 & 1 else "<"
        return np.dtype(f"{order}f{size}"), q + 12
    if cls == 3:  # fixed-length string
        return np.dtype(f"S{size}"), q
    if cls == 6:  # compound
        nmemb = bits0 | (bits1 << 8)
        names, formats, offsets = [], [], []
        for _ in range(nmemb):
            e = b.index(b"\0", q)
            name = b[q:e].decode("ascii", "replace")
            if version in (1, 2):
                q += _pad8(e - q + 1)
            else:
                q = e + 1
            if version == 3:
                nb = 1 if size < 256 else 2 if size < 65536 else 4
                off = int.from_bytes(b[q:q + nb], "little")
                q += nb
                dims: Tuple[int, ...] = ()
            else:
                off, = struct.unpack_from("<I", b, q)
                q += 4
                if version == 1:
                    rank = b[q]
                    dimsz = struct.unpack_from("<4I", b, q + 12)
                    dims = tuple(dimsz[:rank])
                    q += 28
                else:
                    dims = ()
            mtype, q = _parse_datatype(b, q)
            names.append(name)
            formats.append((mtype, dims) if dims else mtype)
            offsets.append(off)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), q
    raise HDF5FormatError(f"datatype class {cls}")


def _parse_filters(b: Optional[bytes]) -> List[int]:
    if b is None:
        return []
    version, n = b[0], b[1]
    ids = []
    if version == 1:
        p = 8
        for _ in range(n):
            fid, nlen, _flags, ncd = struct.unpack_from("<HHHH", b, p)
            p += 8 + _pad8(nlen) + 4 * ncd + (4 if ncd % 2 else 0)
            ids.append(fid)
    elif version == 2:
        p = 2
        for _ in range(n):
            fid, = struct.unpack_from("<H", b, p)
            p += 2
            nlen = 0
            if fid >= 256:
                nlen, = struct.unpack_from("<H", b, p)
                p += 2
            _flags, ncd = struct.unpack_from("<HH", b, p)
            p += 4 + nlen + 4 * ncd
            ids.append(fid)
    else:
        raise HDF5FormatError(f"filter pipeline version {version}")
    for fid in ids:
        if fid not in (1, 2):
            raise HDF5FormatError(f"filter id {fid} (only deflate/shuffle supported)")
    return ids


class File(_Node):
    """`File(path)[...]` mimics the small part of h5py.File the reader needs."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = memoryview(fh.read())
        if bytes(self.buf[:8]) != _SIG:
            raise HDF5FormatError("not an HDF5 file (or a git-LFS pointer)")
        version = self.buf[8]
        if version not in (0, 1):
            raise HDF5FormatError(f"superblock version {version}")
        if self.buf[13] != 8 or self.buf[14] != 8:
            raise HDF5FormatError("only 8-byte offsets/lengths supported")
        p = 24 + (4 if version == 1 else 0)
        p += 32  # base, free-space, eof, driver addresses
        _name_off, root_addr = struct.unpack_from("<QQ", self.buf, p)
        super().__init__(self, root_addr, "/")

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def _local_heap(self, addr: int) -> memoryview:
        if bytes(self.buf[addr:addr + 4]) != b"HEAP":
            raise HDF5FormatError("bad local heap signature")
        size, _free, data = struct.unpack_from("<QQQ", self.buf, addr + 8)
        return self.buf[data:data + size]

    def _walk_group_btree(self, addr: int, heap: memoryview, links: Dict[str, int]):
        buf = self.buf
        if bytes(buf[addr:addr + 4]) != b"TREE":
            raise HDF5FormatError("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", buf, addr + 4)
        if ntype != 0:
            raise HDF5FormatError("expected group B-tree")
        p = addr + 24
        for i in range(used):
            child, = struct.unpack_from("<Q", buf, p + 8 + i * 16)
            if level > 0:
                self._walk_group_btree(child, heap, links)
                continue
            if bytes(buf[child:child + 4]) != b"SNOD":
                raise HDF5FormatError("bad SNOD signature")
            nsym, = struct.unpack_from("<H", buf, child + 6)
            for s in range(nsym):
                e = child + 8 + s * 40
                noff, oaddr = struct.unpack_from("<QQ", buf, e)
                end = noff
                while heap[end] != 0:
                    end += 1
                links[bytes(heap[noff:end]).decode("utf-8", "replace")] = oaddr

    def _walk_chunk_btree(self, addr, ndim, chunk_shape, dtype, filters, out):
        buf = self.buf
        if bytes(buf[addr:addr + 4]) != b"TREE":
            raise HDF5FormatError("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", buf, addr + 4)
        if ntype != 1:
            raise HDF5FormatError("expected chunk B-tree")
        keysize = 8 + 8 * ndim
        p = addr + 24
        for i in range(used):
            kp = p + i * (keysize + 8)
            csize, fmask = struct.unpack_from("<II", buf, kp)
            offs = struct.unpack_from("<" + "Q" * ndim, buf, kp + 8)
            child, = struct.unpack_from("<Q", buf, kp + keysize)
            if level > 0:
                self._walk_chunk_btree(child, ndim, chunk_shape, dtype, filters, out)
                continue
            raw = bytes(buf[child:child + csize])
            for j, fid in reversed(list(enumerate(filters))):
                if fmask & (1 << j):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    isz = dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(isz, -1).T.tobytes()
            chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk_shape))).reshape(chunk_shape)
            sl_out, sl_in = [], []
            for d, (o, c) in enumerate(zip(offs[:-1], chunk_shape)):
                hi = min(o + c, out.shape[d])
                sl_out.append(slice(o, hi))
                sl_in.append(slice(0, hi - o))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]
