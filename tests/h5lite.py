"""h5lite -- a minimal HDF5 reader for the files MEmilio's save_result writes (io/result_io.cpp:33-100): superblock
version 0, groups as symbol tables (B-tree v1 + local heap + SNOD), object headers version 1, contiguous datasets of
little-endian IEEE doubles. Test infrastructure: it checks the facade's hand-written writer
(include/memilio_b200/h5lite.hpp) and is itself validated on a result file written by the reference with libhdf5
(tests/golden/results_osecirvvs.h5, copied from cpp/tests/data)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert b[8] == 0, "superblock version 0 expected"
        assert b[13] == 8 and b[14] == 8, "8-byte offsets and lengths expected"
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", b, 24)
        assert self.eof == len(b), "end-of-file address does not match the file size"
        self.root = self._symbol_entry(56)

    def _symbol_entry(self, o):
        name_off, ohdr, cache, _ = struct.unpack_from("<QQII", self.b, o)
        e = dict(name_off=name_off, ohdr=ohdr, cache=cache)
        if cache == 1:
            e["btree"], e["heap"] = struct.unpack_from("<QQ", self.b, o + 24)
        return e

    def messages(self, addr):
        """[(type, flags, bytes)] of a version 1 object header, following continuation messages"""
        b = self.b
        assert b[addr] == 1, "object header version 1 expected"
        nmsg, = struct.unpack_from("<H", b, addr + 2)
        size, = struct.unpack_from("<I", b, addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            o, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                t, sz, fl = struct.unpack_from("<HHB", b, o)
                data = b[o + 8:o + 8 + sz]
                out.append((t, fl, data))
                if t == 0x10:
                    ca, cl = struct.unpack_from("<QQ", data, 0)
                    blocks.append((ca, cl))
                o += 8 + sz
                left -= 8 + sz
        return out

    def _heap_string(self, heap, off):
        b = self.b
        assert b[heap:heap + 4] == b"HEAP"
        data_addr, = struct.unpack_from("<Q", b, heap + 24)
        s = data_addr + off
        return b[s:b.index(b"\0", s)].decode()

    def _group_tables(self, entry_or_addr):
        if isinstance(entry_or_addr, dict) and entry_or_addr.get("cache") == 1:
            return entry_or_addr["btree"], entry_or_addr["heap"]
        addr = entry_or_addr["ohdr"] if isinstance(entry_or_addr, dict) else entry_or_addr
        for t, _, d in self.messages(addr):
            if t == 0x11:
                return struct.unpack_from("<QQ", d, 0)
        raise KeyError("not a group (no symbol table message)")

    def _walk(self, node, heap, out):
        b = self.b
        assert b[node:node + 4] == b"TREE" and b[node + 4] == 0
        level = b[node + 5]
        used, = struct.unpack_from("<H", b, node + 6)
        o = node + 24
        for i in range(used):
            child, = struct.unpack_from("<Q", b, o + 8 + 16 * i)
            if level > 0:
                self._walk(child, heap, out)
            else:
                assert b[child:child + 4] == b"SNOD"
                n, = struct.unpack_from("<H", b, child + 6)
                for k in range(n):
                    e = self._symbol_entry(child + 8 + 40 * k)
                    out[self._heap_string(heap, e["name_off"])] = e

    def listdir(self, group=None):
        """{name: symbol table entry} of a group (default: the root group), in B-tree (= name) order"""
        btree, heap = self._group_tables(self.root if group is None else group)
        out = {}
        self._walk(btree, heap, out)
        return out

    def dataset(self, entry):
        shape = dtype_ok = addr = size = None
        for t, _, d in self.messages(entry["ohdr"]):
            if t == 0x01:
                ver, rank, flags = d[0], d[1], d[2]
                off = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, d, off)
            elif t == 0x03:
                cls, ver = d[0] & 0x0F, d[0] >> 4
                sz, = struct.unpack_from("<I", d, 4)
                dtype_ok = cls == 1 and sz == 8 and (d[1] & 1) == 0  # IEEE float, 8 bytes, little endian
            elif t == 0x08:
                assert d[0] == 3 and d[1] == 1, "layout version 3, contiguous, expected"
                addr, size = struct.unpack_from("<QQ", d, 2)
        assert shape is not None and dtype_ok and addr is not None, "dataset header incomplete"
        n = int(np.prod(shape)) if len(shape) else 1
        assert size == 8 * n
        return np.frombuffer(self.b, dtype="<f8", count=n, offset=addr).reshape(shape).copy()


def read_result(path):
    """mio::read_result (io/result_io.cpp:110-190): {region id: dict(Time, Total, Group1..)} """
    f = H5File(path)
    out = {}
    for name, e in f.listdir().items():
        out[int(name)] = {k: f.dataset(d) for k, d in f.listdir(e).items()}
    return out
