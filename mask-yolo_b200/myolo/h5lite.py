"""Minimal pure-Python HDF5 reader / writer for Keras weight files.

The reference saves and loads its weights with Keras' HDF5 routines (myolo/model.py:1047-1060 ModelCheckpoint,
1157-1196 `load_weights(filepath, by_name, exclude)` over `keras.engine.saving.load_weights_from_hdf5_group[_by_name]`),
which need h5py.  h5py is not a dependency of this package; the subset of the HDF5 file format those routines produce
with h5py's defaults (libver "earliest") is small enough to restate:

    superblock version 0 (or 1)                      HDF5 File Format Specification, II.A
    groups as symbol tables: v1 B-tree + SNOD + local heap   III.A.1, III.B, III.D
    version-1 object headers with continuation blocks        IV.A.1.a, message 0x0010
    messages: dataspace v1/v2 (0x0001), datatype v1-3 (0x0003: fixed-point, IEEE float, fixed-length string),
              data layout v3 contiguous / compact (0x0008), attribute v1-3 (0x000C), symbol table (0x0011)

Layout Keras 2.x writes (keras/engine/saving.py `save_weights_to_hdf5_group`): root attribute `layer_names` (array of
fixed-length byte strings); one group per layer with attribute `weight_names` (e.g. b'conv1/kernel:0'; for the nested
model b'conv_dw_7/depthwise_kernel:0' inside group 'yolo_model'); one contiguous float32 dataset per weight name, at the
path the name spells below the layer group.  A full-model file (`model.save`) keeps the same structure under the group
'model_weights'.

Anything outside that subset (new-style groups with link messages, chunked / compressed datasets, variable-length
strings, superblock 2/3) raises H5Unsupported with the feature's name -- convert such a file once with
scripts/h5_to_npz.py where h5py exists.  The writer emits the same subset (used for `save_weights('*.h5')` and for the
test fixtures); files it writes carry h5py's default group B-tree parameters (leaf K 4, internal K 16).

The format is third-party: this restatement is pinned only against its own writer and the byte layouts quoted from the
specification in tests/test_h5lite.py, not against libhdf5 (absent here) -- "parity unpinned" in the sense of DESIGN.md.
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Unsupported(NotImplementedError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------------------ reader
class _Dtype(object):
    def __init__(self, np_dtype, size, kind):
        self.np, self.size, self.kind = np_dtype, size, kind


def _parse_datatype(buf, off=0):
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
    cls, ver = cv & 0x0F, cv >> 4
    if ver not in (1, 2, 3):
        raise H5Unsupported("datatype message version %d" % ver)
    order = ">" if (b0 & 1) else "<"
    if cls == 0:                                   # fixed point
        signed = bool(b0 & 0x08)
        return _Dtype(np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), size, "int")
    if cls == 1:                                   # IEEE floating point
        if size not in (2, 4, 8):
            raise H5Unsupported("%d-byte floating point" % size)
        return _Dtype(np.dtype("%sf%d" % (order, size)), size, "float")
    if cls == 3:                                   # fixed-length string
        return _Dtype(np.dtype("S%d" % size), size, "string")
    if cls == 9:
        raise H5Unsupported("variable-length datatype")
    raise H5Unsupported("datatype class %d" % cls)


def _parse_dataspace(buf, off=0):
    ver, rank, flags = struct.unpack_from("<BBB", buf, off)
    if ver == 1:
        p = off + 8
    elif ver == 2:
        if buf[off + 3] == 2:                       # null dataspace
            return None
        p = off + 4
    else:
        raise H5Unsupported("dataspace message version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, buf, p)) if rank else ()


class _Object(object):
    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.messages = f._read_header(addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for typ, data in self.messages:
                if typ == 0x000C:
                    name, val = self._parse_attribute(data)
                    self._attrs[name] = val
                elif typ == 0x0015:
                    raise H5Unsupported("dense attribute storage (attribute info message)")
        return self._attrs

    def _parse_attribute(self, d):
        ver = d[0]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
            p = 8
            name = d[p:p + nsz].split(b"\0")[0].decode("utf8"); p += _pad8(nsz)
            dt = self._try_dtype(d, p); p += _pad8(tsz)
            shape = _parse_dataspace(d, p); p += _pad8(ssz)
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
            p = 8 + (1 if ver == 3 else 0)
            name = d[p:p + nsz].split(b"\0")[0].decode("utf8"); p += nsz
            dt = self._try_dtype(d, p); p += tsz
            shape = _parse_dataspace(d, p); p += ssz
        else:
            raise H5Unsupported("attribute message version %d" % ver)
        if dt is None or shape is None:
            return name, None                       # e.g. variable-length string attributes (keras_version, backend)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(d, dtype=dt.np, count=n, offset=p).reshape(shape)
        return name, (arr.copy() if shape else arr.reshape(()).copy()[()])

    @staticmethod
    def _try_dtype(d, p):
        try:
            return _parse_datatype(d, p)
        except H5Unsupported:
            return None


class Dataset(_Object):
    def __init__(self, f, addr):
        _Object.__init__(self, f, addr)
        self.dtype = self.shape = self._layout = None
        for typ, data in self.messages:
            if typ == 0x0003:
                self.dtype = _parse_datatype(data)
            elif typ == 0x0001:
                self.shape = _parse_dataspace(data)
            elif typ == 0x0008:
                self._layout = data
            elif typ == 0x000B:
                raise H5Unsupported("filtered (compressed) dataset")

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a

    def read(self):
        d = self._layout
        if d is None or self.dtype is None or self.shape is None:
            raise H5Unsupported("dataset without layout / datatype / dataspace message")
        if d[0] != 3:
            raise H5Unsupported("data layout message version %d" % d[0])
        n = int(np.prod(self.shape)) if self.shape else 1
        if d[1] == 1:                                # contiguous
            addr, size = struct.unpack_from("<QQ", d, 2)
            if addr == UNDEF:
                return np.zeros(self.shape, self.dtype.np)
            raw = self.f._read(addr, size)
        elif d[1] == 0:                              # compact
            (size,) = struct.unpack_from("<H", d, 2)
            raw = bytes(d[4:4 + size])
        else:
            raise H5Unsupported("chunked dataset")
        return np.frombuffer(raw, dtype=self.dtype.np, count=n).reshape(self.shape).copy()


class Group(_Object):
    def __init__(self, f, addr):
        _Object.__init__(self, f, addr)
        self._links = None

    def _load(self):
        if self._links is not None:
            return
        self._links = {}
        for typ, data in self.messages:
            if typ == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                heap_data = self.f._heap_data(heap)
                for name_off, obj_addr in self.f._btree_entries(btree):
                    end = heap_data.index(b"\0", name_off)
                    self._links[heap_data[name_off:end].decode("utf8")] = obj_addr
                return
            if typ in (0x0002, 0x0006):
                raise H5Unsupported("new-style group (link messages)")

    def keys(self):
        self._load()
        return list(self._links)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            node._load()
            if part not in node._links:
                raise KeyError(path)
            node = node.f._open(node._links[part])
        return node


class File(Group):
    """Read-only view of an HDF5 file held in memory: `File(path)['group/dataset'].read()`, `.attrs`, `.keys()`."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != SIGNATURE:
            raise ValueError("%s is not an HDF5 file" % path)
        ver = self.buf[8]
        if ver not in (0, 1):
            raise H5Unsupported("superblock version %d" % ver)
        so, sl = self.buf[13], self.buf[14]
        if (so, sl) != (8, 8):
            raise H5Unsupported("offset / length size %d / %d" % (so, sl))
        p = 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", self.buf, p)[0]
        root_entry = p + 32
        root_addr = struct.unpack_from("<Q", self.buf, root_entry + 8)[0]
        self._cache = {}
        Group.__init__(self, self, root_addr)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- low level
    def _read(self, addr, size):
        a = self.base + addr
        if a + size > len(self.buf):
            raise ValueError("HDF5 file is truncated")
        return self.buf[a:a + size]

    def _read_header(self, addr):
        a = self.base + addr
        ver = self.buf[a]
        if self.buf[a:a + 4] == b"OHDR":
            raise H5Unsupported("version-2 object header")
        if ver != 1:
            raise H5Unsupported("object header version %d" % ver)
        nmsgs, _refs, hsize = struct.unpack_from("<HII", self.buf, a + 2)
        msgs, blocks = [], [(a + 16, hsize)]
        while blocks and len(msgs) < nmsgs:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(msgs) < nmsgs:
                typ, size, _flags = struct.unpack_from("<HHB", self.buf, p)
                data = self.buf[p + 8:p + 8 + size]
                p += 8 + size
                if typ == 0x0010:
                    off, length = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.base + off, length))
                msgs.append((typ, data))
        return msgs

    def _heap_data(self, addr):
        h = self._read(addr, 32)
        if h[:4] != b"HEAP":
            raise ValueError("bad local heap signature")
        size, _free, daddr = struct.unpack_from("<QQQ", h, 8)
        return self._read(daddr, size)

    def _btree_entries(self, addr):
        """(name offset in the local heap, object header address) of every link below a v1 group B-tree node."""
        n = self._read(addr, 24)
        if n[:4] == b"SNOD":
            count = struct.unpack_from("<H", n, 6)[0]
            raw = self._read(addr + 8, 40 * count)
            for i in range(count):
                yield struct.unpack_from("<QQ", raw, 40 * i)
            return
        if n[:4] != b"TREE":
            raise ValueError("bad B-tree signature")
        if n[4] != 0:
            raise H5Unsupported("B-tree node type %d" % n[4])
        used = struct.unpack_from("<H", n, 6)[0]
        body = self._read(addr + 24, 8 + 16 * used)
        for i in range(used):
            child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
            for e in self._btree_entries(child):
                yield e

    def _open(self, addr):
        if addr not in self._cache:
            msgs = self._read_header(addr)
            is_group = any(t in (0x0011, 0x0002, 0x0006) for t, _ in msgs)
            self._cache[addr] = Group(self, addr) if is_group else Dataset(self, addr)
        return self._cache[addr]


def read_keras_weights(path):
    """{variable name: float32 array}, in file order, from a Keras 2.x save_weights / model.save file.  Names are the ones
    Keras stores ('conv1/kernel:0'); the variables of a nested model (layer 'yolo_model', weight_names like
    'conv_dw_7/depthwise_kernel:0') come out as 'yolo_model/conv_dw_7/depthwise_kernel:0', so that the owning top-level
    layer -- what load_weights(exclude=...) filters on -- stays visible."""
    out = {}
    f = File(path)
    g = f["model_weights"] if ("model_weights" in f.keys() and "layer_names" not in f.attrs) else f
    names = g.attrs.get("layer_names")
    if names is None:
        raise ValueError("%s has no 'layer_names' attribute: not a Keras weight file" % path)
    for layer in np.atleast_1d(names):
        layer = layer.decode("utf8") if isinstance(layer, bytes) else str(layer)
        grp = g[layer]
        wn = grp.attrs.get("weight_names")
        if wn is None:
            continue
        for w in np.atleast_1d(wn):
            w = w.decode("utf8") if isinstance(w, bytes) else str(w)
            key = w if w.split("/")[0] == layer else layer + "/" + w       # a nested model's variables keep its name in front
            if key in out:
                raise ValueError("duplicate variable name %s" % key)
            out[key] = np.asarray(grp[w].read(), dtype=np.float32)
    return out


# ------------------------------------------------------------------------------------------------------------ writer
class _Writer(object):
    LEAF_K, INTERNAL_K = 4, 16

    def __init__(self):
        self.buf = bytearray(96)                     # superblock placeholder

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def msg(typ, data):
        data = bytes(data) + b"\0" * (_pad8(len(data)) - len(data))
        return struct.pack("<HHB3x", typ, len(data), 0) + data

    def header(self, msgs):
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    @staticmethod
    def datatype(dt):
        dt = np.dtype(dt)
        if dt.kind == "f":
            props = {4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023)}[dt.itemsize]
            return struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) + struct.pack("<HHBBBBI", *props)
        if dt.kind in "iu":
            return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)          # null-terminated / padded ASCII
        raise H5Unsupported("writing dtype %s" % dt)

    @staticmethod
    def dataspace(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + struct.pack("<%dQ" % len(shape), *shape)

    def attribute(self, name, value):
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf8")
        nm = name.encode("utf8") + b"\0"
        dt, ds = self.datatype(arr.dtype), self.dataspace(arr.shape)
        body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds))
        for part in (nm, dt, ds):
            body += part + b"\0" * (_pad8(len(part)) - len(part))
        return self.msg(0x000C, body + arr.tobytes())

    def dataset(self, arr):
        arr = np.asarray(arr)
        if arr.ndim and not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)             # (ascontiguousarray would turn a 0-d array into 1-d)
        addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)
        return self.header([self.msg(0x0001, self.dataspace(arr.shape)), self.msg(0x0003, self.datatype(arr.dtype)),
                            self.msg(0x0005, fill), self.msg(0x0008, layout)])

    def group(self, links, attrs):
        """links: {name: object header address}.  Returns (object header address, B-tree address, heap address)."""
        names = sorted(links, key=lambda s: s.encode("utf8"))
        heap = bytearray(8)                          # offset 0: the empty string (key of the left-most subtree)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            b = n.encode("utf8") + b"\0"
            heap += b + b"\0" * (_pad8(len(b)) - len(b))
        per = 2 * self.LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(chunks) > 2 * self.INTERNAL_K:
            raise H5Unsupported("more than %d links in one group" % (per * 2 * self.INTERNAL_K))
        snods = []
        for ch in chunks:
            body = struct.pack("<4sBxH", b"SNOD", 1, len(ch))
            for n in ch:
                body += struct.pack("<QQII16x", offs[n], links[n], 0, 0)
            body += b"\0" * (40 * (per - len(ch)))
            snods.append(self.alloc(body))
        tree = struct.pack("<4sBBHQQ", b"TREE", 0, 0, len(chunks), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for ch, a in zip(chunks, snods):
            tree += struct.pack("<QQ", a, offs[ch[-1]] if ch else 0)
        tree += b"\0" * (16 * (2 * self.INTERNAL_K - len(chunks)))
        btree = self.alloc(tree)
        hdata = self.alloc(bytes(heap))
        hp = self.alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), UNDEF, hdata))
        msgs = [self.msg(0x0011, struct.pack("<QQ", btree, hp))] + [self.attribute(k, v) for k, v in attrs.items()]
        return self.header(msgs), btree, hp

    def finish(self, root, btree, heap):
        sb = SIGNATURE + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_tree(path, tree, attrs=None):
    """Write a nested dict {name: array | (dict of children, dict of attributes)} as an HDF5 file."""
    w = _Writer()

    def emit(node, node_attrs):
        links = {}
        for name, child in node.items():
            if isinstance(child, tuple):
                links[name] = emit(child[0], child[1])[0]
            elif isinstance(child, dict):
                links[name] = emit(child, {})[0]
            else:
                links[name] = w.dataset(np.asarray(child))
        return w.group(links, node_attrs)

    root, btree, heap = emit(tree, attrs or {})
    data = w.finish(root, btree, heap)
    with open(path, "wb") as fh:
        fh.write(data)


def write_keras_weights(path, layers, nested=None):
    """Keras 2.x `save_weights` layout.  layers: [(layer name, [(weight name as Keras stores it, array), ...]), ...] in model
    order.  nested: {layer name: True} marks a layer that is itself a Model ('yolo_model'): its variables are written
    below the layer's group at the path their names spell, as Keras does."""
    tree = {}
    for lname, weights in layers:
        sub = {}
        for wname, arr in weights:
            node = sub
            parts = wname.split("/")
            for part in parts[:-1]:
                node = node.setdefault(part, ({}, {}))[0]
            node[parts[-1]] = np.asarray(arr, dtype=np.float32)
        wn = np.array([w.encode("utf8") for w, _ in weights]) if weights else np.zeros((0,), "S1")
        tree[lname] = (sub, {"weight_names": wn})
    attrs = {"layer_names": np.array([l.encode("utf8") for l, _ in layers]), "backend": np.array(b"tensorflow"),
             "keras_version": np.array(b"2.2.4")}
    write_tree(path, tree, attrs)
