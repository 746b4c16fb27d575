"""Minimal pure-Python HDF5 reader (TEST INFRASTRUCTURE — part of the oracle, never on the product path).

h5py is not installed in the build container nor on the GPU box, but the
reference's fixtures (``prepared_data/input_*.h5``) and golden outputs
(``output/jax_ref/output_*.h5``) are HDF5.  This reader supports exactly what
those files use: superblock v0, v1 object headers with
continuation blocks, old-style groups (symbol table: v1 B-tree + local heap +
SNOD nodes), contiguous / compact dataset layouts, and fixed-point, float and
compound (v1) datatypes.  Anything else raises ``NotImplementedError``.

It replaces ``h5py.File(fname)['segments'][:]`` of the reference loaders
(reference: optimize/dataio.py:114-115, src/larndsim/sim_jax.py:23-26).
"""
import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Lite:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file: %s" % path)
        if b[8] != 0:
            raise NotImplementedError("superblock version %d" % b[8])
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError("offset/length size != 8")
        # superblock v0: 8 sig, 8 version bytes, 2+2 group leaf/internal k, 4 flags,
        # base addr, free-space addr, eof addr, driver addr, then root symbol table entry
        root_entry = 24 + 4 * 8
        self.root = self._read_symbol_entry(root_entry)

    # ------------------------------------------------------------------ low level
    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _read_symbol_entry(self, off):
        name_off = self._u(off, 8)
        ohdr = self._u(off + 8, 8)
        cache_type = self._u(off + 16, 4)
        scratch = self.b[off + 24:off + 40]
        ent = {"name_off": name_off, "ohdr": ohdr, "cache": cache_type}
        if cache_type == 1:
            ent["btree"] = int.from_bytes(scratch[:8], "little")
            ent["heap"] = int.from_bytes(scratch[8:16], "little")
        return ent

    def _messages(self, ohdr):
        """Yield (type, flags, payload_offset, size) of a v1 object header incl. continuations."""
        b = self.b
        if b[ohdr] != 1:
            raise NotImplementedError("object header version %d" % b[ohdr])
        nmsg = self._u(ohdr + 2, 2)
        hsize = self._u(ohdr + 8, 4)
        blocks = [(ohdr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            start, size = blocks.pop(0)
            p = start
            while p + 8 <= start + size and len(out) < nmsg:
                mtype = self._u(p, 2)
                msize = self._u(p + 2, 2)
                mflags = b[p + 4]
                payload = p + 8
                out.append((mtype, mflags, payload, msize))
                if mtype == 0x10:  # continuation
                    blocks.append((self._u(payload, 8), self._u(payload + 8, 8)))
                p = payload + msize
        return out

    # ------------------------------------------------------------------ groups
    def _heap_data(self, heap_addr):
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        return self._u(heap_addr + 24, 8)

    def _group_entries(self, btree, heap):
        data = self._heap_data(heap)
        out = {}

        def name_at(o):
            s = data + o
            e = self.b.index(b"\x00", s)
            return self.b[s:e].decode()

        def walk(addr):
            sig = self.b[addr:addr + 4]
            if sig == b"TREE":
                level = self.b[addr + 5]
                n = self._u(addr + 6, 2)
                p = addr + 24
                # keys and children alternate: key0, child0, key1, child1 ... key_n
                for i in range(n):
                    child = self._u(p + 8 + i * 16, 8)
                    walk(child)
                _ = level
            elif sig == b"SNOD":
                n = self._u(addr + 6, 2)
                p = addr + 8
                for i in range(n):
                    ent = self._read_symbol_entry(p + i * 40)
                    out[name_at(ent["name_off"])] = ent
            else:
                raise ValueError("bad group node %r" % sig)

        walk(btree)
        return out

    def _children(self, ent):
        if "btree" in ent:
            return self._group_entries(ent["btree"], ent["heap"])
        for mtype, _, payload, _ in self._messages(ent["ohdr"]):
            if mtype == 0x11:  # symbol table message
                return self._group_entries(self._u(payload, 8), self._u(payload + 8, 8))
        return None

    def keys(self, path="/"):
        ent = self._resolve(path)
        ch = self._children(ent)
        if ch is None:
            raise KeyError("%s is not a group" % path)
        return sorted(ch.keys())

    def is_group(self, path):
        return self._children(self._resolve(path)) is not None

    def _resolve(self, path):
        ent = self.root
        for part in [p for p in path.split("/") if p]:
            ch = self._children(ent)
            if ch is None or part not in ch:
                raise KeyError(path)
            ent = ch[part]
        return ent

    # ------------------------------------------------------------------ datatypes
    def _dtype(self, off):
        b = self.b
        cv = b[off]
        cls, ver = cv & 0x0F, cv >> 4
        bits0, bits1, bits2 = b[off + 1], b[off + 2], b[off + 3]
        size = self._u(off + 4, 4)
        prop = off + 8
        if cls == 0:  # fixed point
            order = ">" if bits0 & 1 else "<"
            signed = "i" if bits0 & 8 else "u"
            return np.dtype("%s%s%d" % (order, signed, size)), prop + 4
        if cls == 1:  # floating point
            order = ">" if bits0 & 1 else "<"
            return np.dtype("%sf%d" % (order, size)), prop + 12
        if cls == 6:  # compound
            nmem = bits0 | (bits1 << 8)
            p = prop
            names, fmts, offs = [], [], []
            for _ in range(nmem):
                e = b.index(b"\x00", p)
                name = b[p:e].decode()
                if ver < 3:
                    p += ((e - p + 1) + 7) // 8 * 8
                else:
                    p = e + 1
                if ver == 1:
                    moff = self._u(p, 4)
                    p += 4 + 1 + 3 + 4 + 4 + 16  # offset, dimensionality, reserved, perm, reserved, dims
                elif ver == 2:
                    moff = self._u(p, 4)
                    p += 4
                else:
                    nb = max(1, (size.bit_length() + 7) // 8)
                    moff = self._u(p, nb)
                    p += nb
                mdt, p = self._dtype(p)
                names.append(name)
                fmts.append(mdt)
                offs.append(moff)
            _ = bits2
            return np.dtype({"names": names, "formats": fmts, "offsets": offs, "itemsize": size}), p
        raise NotImplementedError("datatype class %d" % cls)

    # ------------------------------------------------------------------ datasets
    def read(self, path):
        ent = self._resolve(path)
        shape = dtype = None
        data = None
        for mtype, _, payload, msize in self._messages(ent["ohdr"]):
            if mtype == 0x01:  # dataspace
                ver = self.b[payload]
                rank = self.b[payload + 1]
                if ver == 1:
                    p = payload + 8
                elif ver == 2:
                    p = payload + 4
                else:
                    raise NotImplementedError("dataspace version %d" % ver)
                shape = tuple(self._u(p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                dtype, _ = self._dtype(payload)
            elif mtype == 0x08:  # layout
                ver = self.b[payload]
                if ver != 3:
                    raise NotImplementedError("layout version %d" % ver)
                lclass = self.b[payload + 1]
                if lclass == 1:  # contiguous
                    addr = self._u(payload + 2, 8)
                    size = self._u(payload + 10, 8)
                    data = (addr, size)
                elif lclass == 0:  # compact
                    size = self._u(payload + 2, 2)
                    data = (payload + 4, size)
                else:
                    raise NotImplementedError("chunked layout")
        if shape is None or dtype is None or data is None:
            raise KeyError("%s is not a dataset" % path)
        n = int(np.prod(shape)) if shape else 1
        addr, size = data
        if addr == _UNDEF or n == 0:
            return np.zeros(shape, dtype=dtype)
        arr = np.frombuffer(self.b, dtype=dtype, count=n, offset=addr).reshape(shape)
        return arr.copy()

    def walk(self, path="/"):
        """Yield dataset paths below ``path``."""
        for k in self.keys(path):
            p = (path.rstrip("/") + "/" + k)
            if self.is_group(p):
                yield from self.walk(p)
            else:
                yield p


def read_dataset(fname, path):
    return H5Lite(fname).read(path)
