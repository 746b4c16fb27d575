"""HDF5 in and out without h5py (absent from this image): what the reference's loaders read and what its production
driver writes.

* ``read_dataset(path, name)`` / ``H5File(path)``: the ``segments`` table of ``prepared_data/input_*.h5`` (compound records,
  optimize/dataio.py:114-115) and the nested ``batch_<i>/event_<id>/<dataset>`` files the reference writes.
* ``write_h5(path, tree)``: nested groups of N-d numeric arrays in the layout ``optimize/simulate.py:136-165`` produces
  through ``h5py`` with default settings — old-style groups (symbol table, v1 B-tree, local heap), v1 object headers,
  contiguous little-endian datasets — so that ``optimize/comparison.py`` (``h5py.File(...)[batch][event][key]``) reads it.

Format reference: "HDF5 File Format Specification Version 1.1/2.0" (superblock version 0).  With ``h5py`` installed the
caller may of course use it instead; nothing here depends on it.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 8, 16          # symbol-table node holds 2*LEAF_K entries, a B-tree node 2*INTERNAL_K children


# ================================================================================================ reading
class H5File:
    """Eager read-only view: ``f["batch_0/event_63/adc"]`` -> ndarray, ``f.keys("batch_0")`` -> names."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != SIGNATURE:
            raise ValueError("%s: not an HDF5 file" % path)
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise NotImplementedError("only superblock version 0 with 8-byte offsets / lengths is supported")
        self.root = self._group_of_entry(56)

    def _int(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _header_messages(self, addr):
        b = self.buf
        if b[addr] != 1:
            raise NotImplementedError("object header version %d" % b[addr])
        total = self._int(addr + 2, 2)
        spans = [(addr + 16, self._int(addr + 8, 4))]
        msgs = []
        while spans and len(msgs) < total:
            pos, size = spans.pop(0)
            end = pos + size
            while pos + 8 <= end and len(msgs) < total:
                kind, length = self._int(pos, 2), self._int(pos + 2, 2)
                msgs.append((kind, pos + 8, length))
                if kind == 0x10:
                    spans.append((self._int(pos + 8, 8), self._int(pos + 16, 8)))
                pos += 8 + length
        return msgs

    def _group_of_entry(self, entry_off):
        """{'name': object-header address} of the group a symbol-table entry (or its header's 0x11 message) describes."""
        cache = self._int(entry_off + 16, 4)
        if cache == 1:
            return self._walk_group(self._int(entry_off + 24, 8), self._int(entry_off + 32, 8))
        return self._group_of_header(self._int(entry_off + 8, 8))

    def _group_of_header(self, addr):
        for kind, pos, _ in self._header_messages(addr):
            if kind == 0x11:
                return self._walk_group(self._int(pos, 8), self._int(pos + 8, 8))
        return None

    def _walk_group(self, tree, heap):
        b = self.buf
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("corrupt local heap")
        names_at = self._int(heap + 24, 8)
        members = {}
        todo = [tree]
        while todo:
            node = todo.pop()
            tag = b[node:node + 4]
            if tag == b"TREE":
                for i in range(self._int(node + 6, 2)):
                    todo.append(self._int(node + 24 + 8 + 16 * i, 8))
            elif tag == b"SNOD":
                for i in range(self._int(node + 6, 2)):
                    e = node + 8 + 40 * i
                    s = names_at + self._int(e, 8)
                    members[b[s:b.index(b"\x00", s)].decode()] = self._int(e + 8, 8)
            else:
                raise ValueError("corrupt group node")
        return members

    def _lookup(self, path):
        members, addr = self.root, None
        for part in [p for p in path.split("/") if p]:
            if members is None or part not in members:
                raise KeyError(path)
            addr = members[part]
            members = self._group_of_header(addr)
        return addr, members

    def keys(self, path="/"):
        _, members = self._lookup(path)
        if members is None:
            raise KeyError("%s is not a group" % path)
        return sorted(members)

    def __contains__(self, path):
        try:
            self._lookup(path)
            return True
        except KeyError:
            return False

    def _datatype(self, pos):
        b = self.buf
        cls, version = b[pos] & 15, b[pos] >> 4
        flags = b[pos + 1:pos + 4]
        size = self._int(pos + 4, 4)
        if cls == 0:
            return np.dtype("%s%s%d" % (">" if flags[0] & 1 else "<", "i" if flags[0] & 8 else "u", size)), pos + 12
        if cls == 1:
            return np.dtype("%sf%d" % (">" if flags[0] & 1 else "<", size)), pos + 20
        if cls == 6:
            count = flags[0] | (flags[1] << 8)
            cur = pos + 8
            names, formats, offsets = [], [], []
            for _ in range(count):
                end = b.index(b"\x00", cur)
                names.append(b[cur:end].decode())
                cur = cur + (end - cur + 8) // 8 * 8 if version < 3 else end + 1
                if version == 1:
                    offsets.append(self._int(cur, 4))
                    cur += 32
                elif version == 2:
                    offsets.append(self._int(cur, 4))
                    cur += 4
                else:
                    width = max(1, (size.bit_length() + 7) // 8)
                    offsets.append(self._int(cur, width))
                    cur += width
                member, cur = self._datatype(cur)
                formats.append(member)
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), cur
        raise NotImplementedError("HDF5 datatype class %d" % cls)

    def __getitem__(self, path):
        addr, members = self._lookup(path)
        if members is not None or addr is None:
            raise KeyError("%s is a group" % path)
        shape = dtype = where = None
        for kind, pos, _ in self._header_messages(addr):
            if kind == 0x01:
                version, rank = self.buf[pos], self.buf[pos + 1]
                first = pos + (8 if version == 1 else 4)
                shape = tuple(self._int(first + 8 * i, 8) for i in range(rank))
            elif kind == 0x03:
                dtype, _ = self._datatype(pos)
            elif kind == 0x08:
                if self.buf[pos] != 3:
                    raise NotImplementedError("data layout version %d" % self.buf[pos])
                if self.buf[pos + 1] == 1:
                    where = self._int(pos + 2, 8)
                elif self.buf[pos + 1] == 0:
                    where = pos + 4
                else:
                    raise NotImplementedError("chunked datasets")
        if shape is None or dtype is None or where is None:
            raise KeyError("%s is not a dataset" % path)
        count = int(np.prod(shape)) if shape else 1
        if where == UNDEF or count == 0:
            return np.zeros(shape, dtype)
        return np.frombuffer(self.buf, dtype=dtype, count=count, offset=where).reshape(shape).copy()


def read_dataset(path, name):
    return H5File(path)[name]


# ================================================================================================ writing
def _pad8(n):
    return (n + 7) // 8 * 8


def _datatype_message(dt):
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise ValueError("big-endian arrays are not supported")
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize, 0, dt.itemsize * 8, man_bits, exp_bits, 0,
                           man_bits, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    raise ValueError("unsupported dtype %s (numeric little-endian only)" % dt)


def _message(kind, payload, flags=0):
    body = payload + b"\x00" * (_pad8(len(payload)) - len(payload))
    return struct.pack("<HHB3x", kind, len(body), flags) + body


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


class _Writer:
    def __init__(self):
        self.chunks = [b"\x00" * 96]      # superblock placeholder
        self.size = 96

    def put(self, data):
        addr = self.size
        data = data + b"\x00" * (_pad8(len(data)) - len(data))
        self.chunks.append(data)
        self.size += len(data)
        return addr

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        raw = arr.tobytes()
        data_addr = self.put(raw) if raw else UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        return self.put(_object_header([_message(0x01, space), _message(0x03, _datatype_message(arr.dtype), 1), _message(0x05, fill),
                                        _message(0x08, layout)]))

    def group(self, members):
        """members: {name: array | dict}; returns (object header address, B-tree address, heap address)."""
        names = sorted(members, key=lambda s: s.encode())
        addrs = {}
        for n in names:
            child = members[n]
            addrs[n] = self.group(child) if isinstance(child, dict) else (self.dataset(child), None, None)
        # local heap: the empty string at offset 0 (the B-tree's lowest key), then the names
        heap_data, name_off = bytearray(8), {}
        for n in names:
            name_off[n] = len(heap_data)
            raw = n.encode() + b"\x00"
            heap_data += raw + b"\x00" * (_pad8(len(raw)) - len(raw))
        data_addr = self.put(bytes(heap_data))
        heap_addr = self.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))       # 1 = no free block
        # leaves: symbol-table nodes of up to 2*LEAF_K entries (allocated at full size like the library does)
        level, keys = [], []
        for i in range(0, max(len(names), 1), 2 * LEAF_K):
            part = names[i:i + 2 * LEAF_K]
            body = b"SNOD" + struct.pack("<BxH", 1, len(part))
            for n in part:
                hdr, tree, heap = addrs[n]
                if tree is None:
                    body += struct.pack("<QQII16x", name_off[n], hdr, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", name_off[n], hdr, 1, 0, tree, heap)
            body += b"\x00" * (40 * (2 * LEAF_K - len(part)))
            level.append(self.put(body))
            keys.append(name_off[part[-1]] if part else 0)
        depth = 0
        while True:
            nodes, node_keys = [], []
            for i in range(0, len(level), 2 * INTERNAL_K):
                kids, kk = level[i:i + 2 * INTERNAL_K], keys[i:i + 2 * INTERNAL_K]
                first_key = 0 if i == 0 else keys[i - 1]
                body = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(kids), UNDEF, UNDEF) + struct.pack("<Q", first_key)
                for child, key in zip(kids, kk):
                    body += struct.pack("<QQ", child, key)
                body += b"\x00" * (16 * (2 * INTERNAL_K - len(kids)))
                nodes.append(self.put(body))
                node_keys.append(kk[-1])
            # sibling links are left undefined (the library tolerates that for reading; a single root is the common case)
            level, keys, depth = nodes, node_keys, depth + 1
            if len(level) == 1:
                break
        tree_addr = level[0]
        header = self.put(_object_header([_message(0x11, struct.pack("<QQ", tree_addr, heap_addr))]))
        return header, tree_addr, heap_addr

    def finish(self, root):
        header, tree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, self.size, UNDEF)
        sb += struct.pack("<QQIIQQ", 0, header, 1, 0, tree, heap)
        assert len(sb) == 96
        self.chunks[0] = sb
        return b"".join(self.chunks)


def write_h5(path, tree):
    """``tree``: nested dict, leaves are numeric arrays — e.g. {"batch_0": {"event_63": {"adc": ..., "ticks": ...}}}."""
    w = _Writer()
    blob = w.finish(w.group(tree))
    with open(path, "wb") as fh:
        fh.write(blob)
    return len(blob)
