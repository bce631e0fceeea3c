"""Minimal pure-Python reader of the Keras-2.8 `save_weights(*.h5)` files of the reference
(`gated_ccvae.py:146-159, 391-411`; files `models/params_*/{encoder_model,decoder_model,cond_prior,classifier}_{best,last}.h5`).

h5py is not available, so this parses exactly the subset of HDF5 that h5py 3.x / libhdf5 1.12 emit for such files:
superblock version 0, version-1 object headers (with continuation blocks), old-style groups (symbol-table message ->
version-1 B-tree -> SNOD symbol nodes -> local heap), contiguous or compact dataset layouts (layout message v3),
little-endian IEEE floats / fixed-point integers, and fixed- or variable-length (global heap) string attributes (`layer_names`, `weight_names`,
`backend`, `keras_version`).  Anything else (chunked / filtered datasets, new-style groups, variable-length sequences)
raises `H5Error` — a loud failure, never a guess.

    f = H5File(path)
    f.attrs("/")["layer_names"]                 -> [b"conv2d", ...]
    f.attrs("/conv2d")["weight_names"]          -> [b"conv2d/kernel:0", b"conv2d/bias:0"]
    f["/conv2d/conv2d/kernel:0"]                -> np.ndarray float32 [4,4,3,32]
    keras_weights(path)                         -> [(weight_name, ndarray)] in Keras' own order
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_VLEN_STR = np.dtype("O")      # marker for variable-length strings (global-heap references)


class H5Error(ValueError):
    pass


class _Obj:
    """parsed object header: either a group (btree, heap) or a dataset (shape, dtype, data location)."""
    __slots__ = ("addr", "btree", "heap", "shape", "dtype", "data_addr", "data_size", "compact", "attrs")

    def __init__(self, addr):
        self.addr = addr
        self.btree = self.heap = None
        self.shape = self.dtype = None
        self.data_addr = self.data_size = None
        self.compact = None
        self.attrs = {}


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("{}: not an HDF5 file".format(path))
        if b[8] != 0:
            raise H5Error("superblock version {} not supported (expected 0)".format(b[8]))
        self.size_of_offsets, self.size_of_lengths = b[13], b[14]
        if (self.size_of_offsets, self.size_of_lengths) != (8, 8):
            raise H5Error("only 8-byte offsets/lengths are supported")
        # superblock v0: 8 sig, 8 versions/sizes, 2+2 group k's, 4 flags, then base / freespace / eof / driver
        # addresses (4 x 8 bytes), then the root group's symbol-table entry
        self.base = self._u64(24)
        root_entry = 24 + 32
        self.root_addr = self._u64(root_entry + 8)
        self._cache = {}

    # ---- primitives ---------------------------------------------------------------------------------------------
    def _u8(self, o):
        return self.buf[o]

    def _u16(self, o):
        return struct.unpack_from("<H", self.buf, o)[0]

    def _u32(self, o):
        return struct.unpack_from("<I", self.buf, o)[0]

    def _u64(self, o):
        return struct.unpack_from("<Q", self.buf, o)[0]

    # ---- object headers (version 1) -------------------------------------------------------------------------------
    def _object(self, addr) -> _Obj:
        if addr in self._cache:
            return self._cache[addr]
        a = addr + self.base
        if self._u8(a) != 1:
            raise H5Error("object header version {} at {} not supported (expected 1)".format(self._u8(a), addr))
        n_msgs = self._u16(a + 2)
        hdr_size = self._u32(a + 8)
        obj = _Obj(addr)
        blocks = [(a + 16, hdr_size)]          # 12-byte prefix padded to 16
        seen = 0
        while blocks and seen < n_msgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < n_msgs:
                mtype, msize, mflags = self._u16(pos), self._u16(pos + 2), self._u8(pos + 4)
                body = pos + 8
                seen += 1
                if mflags & 0x02:
                    raise H5Error("shared header messages are not supported")
                if mtype == 0x0010:            # continuation
                    blocks.append((self._u64(body) + self.base, self._u64(body + 8)))
                elif mtype == 0x0011:          # symbol table: old-style group
                    obj.btree, obj.heap = self._u64(body), self._u64(body + 8)
                elif mtype == 0x0001:
                    obj.shape = self._dataspace(body)
                elif mtype == 0x0003:
                    obj.dtype = self._datatype(body)[0]
                elif mtype == 0x0008:
                    self._layout(body, obj)
                elif mtype == 0x000B:
                    raise H5Error("filtered (compressed) datasets are not supported")
                elif mtype == 0x000C:
                    k, v = self._attribute(body)
                    obj.attrs[k] = v
                elif mtype in (0x0002, 0x0006):
                    raise H5Error("new-style groups (link messages) are not supported")
                pos = body + msize
        self._cache[addr] = obj
        return obj

    def _dataspace(self, o):
        ver, rank, flags = self._u8(o), self._u8(o + 1), self._u8(o + 2)
        if ver == 1:
            o += 8
        elif ver == 2:
            if self._u8(o + 3) == 2:          # null dataspace
                return None
            o += 4
        else:
            raise H5Error("dataspace version {}".format(ver))
        return tuple(self._u64(o + 8 * i) for i in range(rank))

    def _datatype(self, o):
        """-> (numpy dtype, bytes consumed by the message body)."""
        cv = self._u8(o)
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = self._u8(o + 1)
        size = self._u32(o + 4)
        if cls == 0:       # fixed point
            if bits0 & 1:
                raise H5Error("big-endian integers are not supported")
            signed = bool(bits0 & 0x08)
            return np.dtype("<{}{}".format("i" if signed else "u", size)), 12
        if cls == 1:       # floating point
            if bits0 & 1:
                raise H5Error("big-endian floats are not supported")
            if size not in (2, 4, 8):
                raise H5Error("float size {}".format(size))
            return np.dtype("<f{}".format(size)), 20
        if cls == 3:       # fixed-length string
            return np.dtype("S{}".format(size)), 8
        if cls == 9:       # variable length: only strings (how h5py stores Keras' layer_names / weight_names)
            if (bits0 & 0x0F) != 1:
                raise H5Error("variable-length sequences are not supported")
            return _VLEN_STR, 8
        raise H5Error("datatype class {} is not supported".format(cls))

    def _layout(self, o, obj):
        ver = self._u8(o)
        if ver != 3:
            raise H5Error("data layout message version {} (expected 3)".format(ver))
        lc = self._u8(o + 1)
        if lc == 1:        # contiguous
            obj.data_addr, obj.data_size = self._u64(o + 2), self._u64(o + 10)
        elif lc == 0:      # compact
            n = self._u16(o + 2)
            obj.compact = bytes(self.buf[o + 4:o + 4 + n])
        else:
            raise H5Error("chunked datasets are not supported")

    def _attribute(self, o):
        ver = self._u8(o)
        if ver not in (1, 2, 3):
            raise H5Error("attribute message version {}".format(ver))
        name_size, dt_size, ds_size = self._u16(o + 2), self._u16(o + 4), self._u16(o + 6)
        p = o + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = bytes(self.buf[p:p + name_size]).split(b"\0")[0].decode("utf8")
        p += pad(name_size)
        try:
            dtype, _ = self._datatype(p)
        except H5Error:
            return name, None          # an attribute this reader cannot decode is skipped, not fatal
        p += pad(dt_size)
        shape = self._dataspace(p)
        p += pad(ds_size)
        if shape is None:
            return name, None
        n = int(np.prod(shape)) if shape else 1
        if dtype is _VLEN_STR:
            vals = [self._global_heap_object(self._u64(p + 16 * i + 4), self._u32(p + 16 * i + 12))[:self._u32(p + 16 * i)]
                    for i in range(n)]
            return name, (vals if shape else vals[0])
        arr = np.frombuffer(self.buf, dtype=dtype, count=n, offset=p)
        if dtype.kind == "S":
            vals = [bytes(v) for v in arr]
            return name, (vals if shape else vals[0])
        arr = arr.reshape(shape)
        return name, (arr.copy() if shape else arr.reshape(()).item())

    def _global_heap_object(self, addr, index):
        a = addr + self.base
        if self.buf[a:a + 4] != b"GCOL":
            raise H5Error("bad global heap signature at {}".format(addr))
        end = a + self._u64(a + 8)
        p = a + 16
        while p + 16 <= end:
            idx, size = self._u16(p), self._u64(p + 8)
            if idx == 0:
                break
            if idx == index:
                return bytes(self.buf[p + 16:p + 16 + size])
            p += 16 + ((size + 7) & ~7)
        raise H5Error("global heap object {} not found in collection at {}".format(index, addr))

    # ---- groups -----------------------------------------------------------------------------------------------------
    def _heap_data(self, heap_addr):
        a = heap_addr + self.base
        if self.buf[a:a + 4] != b"HEAP":
            raise H5Error("bad local heap signature at {}".format(heap_addr))
        return self._u64(a + 24) + self.base

    def _children(self, obj):
        """ordered {name: object header address} of an old-style group (B-tree order = name order)."""
        if obj.btree is None:
            raise H5Error("not a group")
        heap = self._heap_data(obj.heap)
        out = {}

        def name_at(off):
            s = heap + off
            e = self.buf.index(b"\0", s)
            return bytes(self.buf[s:e]).decode("utf8")

        def walk(addr):
            a = addr + self.base
            sig = self.buf[a:a + 4]
            if sig == b"TREE":
                ntype, level, used = self._u8(a + 4), self._u8(a + 5), self._u16(a + 6)
                if ntype != 0:
                    raise H5Error("unexpected B-tree node type {}".format(ntype))
                p = a + 24                       # after sig/type/level/used and the two sibling addresses
                for i in range(used):            # key0, child0, key1, child1, ..., key_n
                    child = self._u64(p + 8 + 16 * i)
                    walk(child)
            elif sig == b"SNOD":
                n = self._u16(a + 6)
                for i in range(n):
                    e = a + 8 + 40 * i
                    out[name_at(self._u64(e))] = self._u64(e + 8)
            else:
                raise H5Error("bad group node signature {!r} at {}".format(sig, addr))

        if obj.btree != UNDEF:
            walk(obj.btree)
        return out

    def _resolve(self, path) -> _Obj:
        obj = self._object(self.root_addr)
        for part in [p for p in path.split("/") if p]:
            kids = self._children(obj)
            if part not in kids:
                raise KeyError(path)
            obj = self._object(kids[part])
        return obj

    # ---- public ---------------------------------------------------------------------------------------------------------
    def keys(self, path="/"):
        return list(self._children(self._resolve(path)).keys())

    def attrs(self, path="/"):
        return dict(self._resolve(path).attrs)

    def is_group(self, path):
        return self._resolve(path).btree is not None

    def __getitem__(self, path) -> np.ndarray:
        obj = self._resolve(path)
        if obj.btree is not None:
            raise H5Error("{} is a group".format(path))
        if obj.dtype is None or obj.shape is None:
            raise H5Error("{}: dataset without datatype/dataspace".format(path))
        n = int(np.prod(obj.shape)) if obj.shape else 1
        if obj.compact is not None:
            arr = np.frombuffer(obj.compact, dtype=obj.dtype, count=n)
        else:
            if obj.data_addr in (None, UNDEF):
                raise H5Error("{}: dataset has no allocated storage".format(path))
            if obj.data_size < n * obj.dtype.itemsize:
                raise H5Error("{}: storage smaller than the dataspace".format(path))
            arr = np.frombuffer(self.buf, dtype=obj.dtype, count=n, offset=obj.data_addr + self.base)
        return arr.reshape(obj.shape).copy()

    def visit(self, path="/"):
        """all dataset paths below `path`, depth first in name order."""
        out = []
        obj = self._resolve(path)
        for name, addr in self._children(obj).items():
            child = path.rstrip("/") + "/" + name
            if self._object(addr).btree is not None:
                out.extend(self.visit(child))
            else:
                out.append(child)
        return out


def keras_weights(path):
    """[(weight_name, ndarray)] in the order Keras' `load_weights` consumes them: for every name in the root attribute
    `layer_names`, the layer group's `weight_names` attribute, each a dataset path below that group
    (keras/saving/hdf5_format.py `load_weights_from_hdf5_group`; plus `top_level_model_weights`, empty here)."""
    f = H5File(path)
    root = f.attrs("/")
    if "layer_names" not in root:
        raise H5Error("{}: no layer_names attribute - not a Keras weights file".format(path))
    out = []
    for lname in root["layer_names"]:
        lname = lname.decode("utf8")
        wnames = f.attrs("/" + lname).get("weight_names")
        if not isinstance(wnames, list):        # weight-less layers (Reshape) carry an empty float array
            continue
        for wn in wnames:
            wn = wn.decode("utf8")
            out.append((wn, f["/" + lname + "/" + wn]))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# writer: the same subset, laid out the way libhdf5 1.12 / h5py lays out the reference's files
# ---------------------------------------------------------------------------------------------------------------------
_LEAF_K, _INT_K = 4, 16                      # group leaf / internal node K of the reference's superblocks
_TREE_BYTES = 24 + (2 * _INT_K + 1) * 8 + 2 * _INT_K * 8      # 544
_SNOD_BYTES = 8 + 2 * _LEAF_K * 40                            # 328
_FREE_NULL = 1                               # H5HL_FREE_NULL: "no free block" in a local heap


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dataspace_v1(shape) -> bytes:
    # version 1, rank, flags (0: no max dims), 5 reserved bytes, then the dimensions
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


_F32_TYPE = bytes.fromhex("11201f000400000000002000170800177f000000")     # IEEE little-endian binary32 (as in the files)


def _string_type(size: int) -> bytes:
    # class 3 (string), version 1, null-padded ASCII - what h5py writes for numpy 'S' data
    return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(1, size))


def _attribute_v1(name: str, dtype_msg: bytes, shape, data: bytes) -> bytes:
    nm = name.encode("utf8") + b"\0"
    ds = _dataspace_v1(shape)
    return (struct.pack("<BxHHH", 1, len(nm), len(dtype_msg), len(ds)) + _pad8(nm) + _pad8(dtype_msg) + _pad8(ds) + data)


def _string_attr(name: str, value) -> bytes:
    """scalar bytes or list of bytes -> attribute message body with a fixed-length string type."""
    if isinstance(value, (bytes, str)):
        v = value.encode("utf8") if isinstance(value, str) else value
        return _attribute_v1(name, _string_type(len(v)), (), v.ljust(max(1, len(v)), b"\0"))
    vals = [v.encode("utf8") if isinstance(v, str) else v for v in value]
    width = max([len(v) for v in vals] + [1])
    return _attribute_v1(name, _string_type(width), (len(vals),), b"".join(v.ljust(width, b"\0") for v in vals))


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)             # superblock + root symbol-table entry, filled in at the end

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, messages) -> int:
        body = b"".join(messages)
        # version 1 prefix: version, reserved, message count, reference count, header size, 4 bytes of alignment
        return self.alloc(struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body)

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self.alloc(arr.tobytes())
        msgs = [
            _msg(0x0001, struct.pack("<BBB5x", 1, arr.ndim, 1) + b"".join(struct.pack("<Q", d) for d in arr.shape) * 2),
            _msg(0x0003, _F32_TYPE, flags=1),
            _msg(0x0005, bytes.fromhex("0202020100000000"), flags=1),      # fill value v2: none defined
            _msg(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)),
        ]
        return self.object_header(msgs)

    def group(self, children, attrs=()):
        """children: {name: (object header address, btree, heap)} (btree/heap None for datasets).
        -> (object header address, btree address, heap address)"""
        names = sorted(children, key=lambda s: s.encode("utf8"))
        heap = bytearray(8)                  # offset 0: the empty string (B-tree key 0, root entry name)
        off = {}
        for n in names:
            off[n] = len(heap)
            heap += _pad8(n.encode("utf8") + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), _FREE_NULL, heap_data))
        # symbol nodes of up to 2 * leaf K entries, in name order
        per = 2 * _LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)]
        if len(chunks) > 2 * _INT_K:
            raise H5Error("group with more than {} entries".format(2 * _INT_K * per))
        keys, kids = [0], []
        for ch in chunks:
            body = b"SNOD" + struct.pack("<BxH", 1, len(ch))
            for n in ch:
                addr, bt, hp = children[n]
                if bt is None:
                    body += struct.pack("<QQII16x", off[n], addr, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", off[n], addr, 1, 0, bt, hp)
            kids.append(self.alloc(body.ljust(_SNOD_BYTES, b"\0")))
            keys.append(off[ch[-1]])
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(kids), UNDEF, UNDEF)
        for i, kid in enumerate(kids):
            tree += struct.pack("<QQ", keys[i], kid)
        tree += struct.pack("<Q", keys[len(kids)])
        btree_addr = self.alloc(tree.ljust(_TREE_BYTES, b"\0"))
        msgs = [_msg(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + [_msg(0x000C, a) for a in attrs]
        return self.object_header(msgs), btree_addr, heap_addr

    def finish(self, root) -> bytes:
        addr, bt, hp = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, _LEAF_K, _INT_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQIIQQ", 0, addr, 1, 0, bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_keras_weights(path, layers, backend=b"tensorflow", keras_version=b"2.8.0"):
    """Write a Keras-2.8 style `save_weights` HDF5 file.  `layers` = [(layer_name, [(weight_name, ndarray), ...])] in
    model order; a weight name is the dataset path below its layer group (e.g. "encoder/conv2d/kernel:0").  Same
    structures as the reference's files (superblock 0, version-1 object headers, symbol-table groups, contiguous
    float32 datasets); string attributes are fixed-length (as Keras <= 2.2 / h5py write numpy 'S' data) instead of
    variable-length, which `load_weights` accepts as well.  Checked against this module's reader and against the
    byte layout of the reference's files; NOT checked against libhdf5 (not available here)."""
    w = _Writer()

    def build(tree, attrs=()):
        kids = {}
        for name, node in tree.items():
            if isinstance(node, dict):
                kids[name] = build(node)
            else:
                kids[name] = (w.dataset(node), None, None)
        return w.group(kids, attrs)

    root = {}
    for lname, weights in layers:
        tree = {}
        for wname, arr in weights:
            node = tree
            parts = wname.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = np.asarray(arr)
        root[lname] = (tree, [wn for wn, _ in weights])
    kids = {}
    for lname, (tree, wnames) in root.items():
        sub = {}
        for name, node in tree.items():
            sub[name] = build(node) if isinstance(node, dict) else (w.dataset(node), None, None)
        kids[lname] = w.group(sub, [_string_attr("weight_names", wnames)])
    kids["top_level_model_weights"] = w.group({}, [_string_attr("weight_names", [])])
    attrs = [_string_attr("layer_names", [ln for ln, _ in layers]), _string_attr("backend", backend),
             _string_attr("keras_version", keras_version)]
    data = w.finish(w.group(kids, attrs))
    with open(path, "wb") as fh:
        fh.write(data)
