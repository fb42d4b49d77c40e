"""Minimal reader of uncompressed Blender 2.5x/2.6x .blend files (file-block + SDNA layout), enough to recover the
fixtures the reference ships as example*.blend without Blender: objects (type, parent, loc / rot / size, parentinv),
meshes (MVert / MFace / material slots), materials and objects' ID properties (the E.A.R. add-on keeps its settings
there), F-curves (keyframed location), scene frame range.  Used once, by tests/golden/make_example2.py, to emit
ear_b200/scenes_example2.npz; nothing at run time depends on it."""
from __future__ import annotations

import struct
from typing import Dict, List


class Blend:
    def __init__(self, path: str):
        self.d = open(path, "rb").read()
        assert self.d[:7] == b"BLENDER", "not a .blend (or compressed)"
        self.psz = 8 if self.d[7:8] == b"-" else 4
        self.end = "<" if self.d[8:9] == b"v" else ">"
        self.version = int(self.d[9:12])
        self.blocks = []          # (code, size, old, sdna, count, offset)
        self.by_ptr: Dict[int, tuple] = {}
        p = 12
        hdr = self.end + "4si" + ("Q" if self.psz == 8 else "I") + "ii"
        hsz = struct.calcsize(hdr)
        while p < len(self.d):
            code, size, old, sdna, count = struct.unpack_from(hdr, self.d, p)
            b = (code.rstrip(b"\0"), size, old, sdna, count, p + hsz)
            self.blocks.append(b)
            self.by_ptr[old] = b
            p += hsz + size
            if code[:4] == b"ENDB":
                break
        self._read_dna()

    # ---- SDNA ----
    def _read_dna(self):
        b = next(b for b in self.blocks if b[0] == b"DNA1")
        d, p = self.d, b[5]
        assert d[p:p + 8] == b"SDNANAME"
        p += 8
        (n,) = struct.unpack_from(self.end + "i", d, p); p += 4
        names = []
        for _ in range(n):
            e = d.index(b"\0", p); names.append(d[p:e].decode()); p = e + 1
        p = (p + 3) & ~3
        assert d[p:p + 4] == b"TYPE"; p += 4
        (n,) = struct.unpack_from(self.end + "i", d, p); p += 4
        types = []
        for _ in range(n):
            e = d.index(b"\0", p); types.append(d[p:e].decode()); p = e + 1
        p = (p + 3) & ~3
        assert d[p:p + 4] == b"TLEN"; p += 4
        tlen = list(struct.unpack_from(self.end + f"{len(types)}h", d, p)); p += 2 * len(types)
        p = (p + 3) & ~3
        assert d[p:p + 4] == b"STRC"; p += 4
        (n,) = struct.unpack_from(self.end + "i", d, p); p += 4
        self.structs = []         # index -> (type name, [(field type, field name)])
        self.struct_by_name = {}
        for _ in range(n):
            t, nf = struct.unpack_from(self.end + "hh", d, p); p += 4
            fields = []
            for _ in range(nf):
                ft, fn = struct.unpack_from(self.end + "hh", d, p); p += 4
                fields.append((types[ft], names[fn]))
            self.structs.append((types[t], fields))
            self.struct_by_name[types[t]] = len(self.structs) - 1
        self.tlen = dict(zip(types, tlen))

    def _field_size(self, ftype, fname):
        n = 1
        base = fname
        while base.endswith("]"):
            i = base.rindex("[")
            n *= int(base[i + 1:-1]); base = base[:i]
        if base.startswith("*") or base.startswith("(*"):
            return self.psz * n, base, n
        return self.tlen[ftype] * n, base, n

    def layout(self, sname):
        out, off = {}, 0
        for ftype, fname in self.structs[self.struct_by_name[sname]][1]:
            size, base, n = self._field_size(ftype, fname)
            out[base.lstrip("*(").rstrip(")")] = (off, ftype, fname, size, n)
            off += size
        return out, off

    def get(self, sname, base_off, path):
        """Reads field `path` ('id.name', 'loc', 'data' ...) of struct `sname` at file offset base_off."""
        parts = path.split(".")
        off, cur = base_off, sname
        for i, part in enumerate(parts):
            lay, _ = self.layout(cur)
            o, ftype, fname, size, n = lay[part]
            off += o
            if i + 1 < len(parts):
                cur = ftype
                continue
            is_ptr = fname.startswith("*") or fname.startswith("(*")
            if is_ptr:
                vals = struct.unpack_from(self.end + ("Q" if self.psz == 8 else "I") * n, self.d, off)
                return vals[0] if n == 1 else list(vals)
            fmt = {"float": "f", "int": "i", "short": "h", "char": "c", "double": "d", "uchar": "B", "ushort": "H", "uint": "I"}.get(ftype)
            if fmt is None:
                return (ftype, off)       # nested struct: caller continues with get(ftype, off, ...)
            if ftype == "char" and n > 1:
                raw = self.d[off:off + n]
                return raw.split(b"\0")[0].decode("latin1")
            vals = struct.unpack_from(self.end + fmt * n, self.d, off)
            return vals[0] if n == 1 else list(vals)

    def blocks_of(self, code):
        return [b for b in self.blocks if b[0] == code]

    def struct_name(self, block):
        return self.structs[block[3]][0]

    def deref(self, ptr):
        return self.by_ptr.get(ptr)
