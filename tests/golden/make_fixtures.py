"""Generate the committed golden fixtures from the reference's own test data.

Run ONCE in the build container (where /root/reference is mounted):

    python tests/golden/make_fixtures.py

Outputs (committed, tiny):
  tests/golden/meshes.npz       -- the 4 mesh fixtures of the reference's tests
                                   (tests/tests_data/{coarse_square,disk,square_tri,square_quad}.h5)
  tests/golden/golden_tags.npz  -- every golden CSV of tests/tests_data/*.csv
                                   (key "<csv stem>" -> int32 array [2, n]: row 0 indices, row 1 values;
                                   layout as written by tests/test_compute_meshtags.py:182-196)

Nothing under tests/ or the product reads /root/reference at run time; only this
script does.  h5py is not available in the image, so a ~100-line reader for the
subset of HDF5 these four files use (superblock v0, v1 object headers / groups,
contiguous or single-chunk deflate+shuffle datasets) is included here.
"""
import glob
import os
import struct
import zlib

import numpy as np

REF = "/root/reference/tests/tests_data"
HERE = os.path.dirname(os.path.abspath(__file__))


class MiniH5:
    """Just enough HDF5 to read the reference's four mesh files."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n" or b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise ValueError("unsupported HDF5 flavour: " + path)
        # superblock v0: root symbol-table entry sits after 4 addresses at byte 24
        self.root = self._sym(24 + 32)["header"]

    def _sym(self, off):
        name_off, header = struct.unpack_from("<QQ", self.buf, off)
        return {"name_off": name_off, "header": header}

    def _messages(self, addr):
        b = self.buf
        version, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert version == 1
        todo, out = [(addr + 16, hsize)], []
        while todo:
            p, size = todo.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", b, p)
                p += 8
                body = b[p:p + msize]
                if mtype == 0x10:  # continuation block
                    todo.append(struct.unpack_from("<QQ", body, 0))
                out.append((mtype, body))
                p += msize
        return out

    def children(self, header):
        for mtype, body in self._messages(header):
            if mtype == 0x11:  # symbol table message: (btree, heap)
                btree, heap = struct.unpack_from("<QQ", body, 0)
                assert self.buf[heap:heap + 4] == b"HEAP"
                heap_data = struct.unpack_from("<QQQ", self.buf, heap + 8)[2]
                return self._walk(btree, heap_data)
        return None

    def _walk(self, node, heap_data):
        b = self.buf
        assert b[node:node + 4] == b"TREE"
        _, level, nent = struct.unpack_from("<BBH", b, node + 4)
        p = node + 24
        kids = []
        for _ in range(nent):
            p += 8
            kids.append(struct.unpack_from("<Q", b, p)[0])
            p += 8
        out = {}
        for kid in kids:
            if level > 0:
                out.update(self._walk(kid, heap_data))
                continue
            assert b[kid:kid + 4] == b"SNOD"
            n = struct.unpack_from("<H", b, kid + 6)[0]
            for j in range(n):
                ent = self._sym(kid + 8 + 40 * j)
                name = b[heap_data + ent["name_off"]:].split(b"\0", 1)[0].decode()
                out[name] = ent["header"]
        return out

    def get(self, path):
        node = self.root
        for part in path.strip("/").split("/"):
            node = self.children(node)[part]
        return self.read(node)

    def read(self, header):
        b = self.buf
        shape = dtype = None
        layout = filters = None
        for mtype, body in self._messages(header):
            if mtype == 0x1:  # dataspace
                rank = body[1]
                shape = struct.unpack_from("<%dQ" % rank, body, 8 if body[0] == 1 else 4)
            elif mtype == 0x3:  # datatype
                cls, size = body[0] & 0xF, struct.unpack_from("<I", body, 4)[0]
                signed = bool(body[1] & 0x08)
                dtype = {(0, 8): "<i8" if signed else "<u8", (0, 4): "<i4" if signed else "<u4",
                         (1, 8): "<f8", (1, 4): "<f4"}[(cls, size)]
                item = size
            elif mtype == 0x8:  # layout
                layout = body
            elif mtype == 0xB:  # filter pipeline
                filters = body
        n = int(np.prod(shape))
        if layout[0] == 3 and layout[1] == 1:  # contiguous
            addr = struct.unpack_from("<Q", layout, 2)[0]
            return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()
        if layout[0] in (1, 2):  # old-style contiguous
            addr = struct.unpack_from("<Q", layout, 8)[0]
            return np.frombuffer(b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()
        assert layout[0] == 3 and layout[1] == 2, "unsupported layout"
        nd = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<%dI" % nd, layout, 11)
        ids = []
        if filters is not None:
            fver, nfil = filters[0], filters[1]
            p = 8 if fver == 1 else 2
            for _ in range(nfil):
                fid = struct.unpack_from("<H", filters, p)[0]
                assert fver == 1
                nlen, _, ncd = struct.unpack_from("<HHH", filters, p + 2)
                p += 8 + (nlen + 7) // 8 * 8 + 4 * ncd + (4 if ncd % 2 else 0)
                ids.append(fid)
        assert b[btree:btree + 4] == b"TREE"
        ntype, level, nent = struct.unpack_from("<BBH", b, btree + 4)
        assert ntype == 1 and level == 0
        out = np.zeros(shape, dtype=dtype)
        p = btree + 24
        for _ in range(nent):
            csize, _ = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<%dQ" % nd, b, p + 8)
            p += 8 + 8 * nd
            child = struct.unpack_from("<Q", b, p)[0]
            p += 8
            raw = b[child:child + csize]
            for fid in reversed(ids):
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:  # byte shuffle
                    raw = np.frombuffer(raw, np.uint8).reshape(item, -1).T.copy().tobytes()
                else:
                    raise ValueError("filter %d" % fid)
            chunk = np.frombuffer(raw, dtype=dtype).reshape(cdims[:-1])
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:-1], cdims[:-1], shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out


def main():
    meshes = {}
    for name in ("coarse_square", "square_tri", "square_quad"):
        h5 = MiniH5(os.path.join(REF, name + ".h5"))
        meshes[name + "_x"] = h5.get("/Mesh/mesh/geometry").astype(np.float64)
        # file order of the vertices of each cell (XDMF order; quads are stored cyclically)
        meshes[name + "_cells"] = h5.get("/Mesh/mesh/topology").astype(np.int32)
    h5 = MiniH5(os.path.join(REF, "disk.h5"))
    meshes["disk_x"] = h5.get("/data0").astype(np.float64)
    meshes["disk_cells"] = h5.get("/data1").astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "meshes.npz"), **meshes)
    for k, v in meshes.items():
        print(k, v.shape, v.dtype)

    gold = {}
    for path in sorted(glob.glob(os.path.join(REF, "*.csv"))):
        arr = np.loadtxt(path, delimiter=" ", ndmin=2)
        stem = os.path.basename(path)[:-4]
        assert arr.shape[0] == 2 and np.all(arr == np.round(arr))
        gold[stem] = arr.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "golden_tags.npz"), **gold)
    print(len(gold), "golden CSVs packed")


if __name__ == "__main__":
    main()
