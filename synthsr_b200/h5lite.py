"""Minimal pure-Python HDF5 reader / writer for Keras weight files (`model.save_weights('x.h5')`,
`model.load_weights('x.h5', by_name=True)` in the reference: SynthSR/training.py:353-369, 429-439; the shipped
models/SynthSR_v10_210712*.h5).  h5py / libhdf5 are not available in this image, so the subset of the HDF5 file format
that h5py 2.10 + Keras 2.3.1 emit is implemented here from the published format specification (HDF5 File Format
Specification version 2.0, the "classic" layout libhdf5 writes by default):

    superblock version 0                           8-byte offsets / lengths
    object headers version 1                       messages: dataspace v1 (0x01), datatype v1 (0x03), fill value (0x05),
                                                   layout v3 contiguous / compact (0x08), attribute v1 (0x0C),
                                                   continuation (0x10), symbol table (0x11), modification time (0x12)
    groups = symbol tables                         B-tree v1 (TREE, node type 0) + SNOD leaves + local heap (HEAP)
    datatypes                                      fixed-point, IEEE float, fixed-length string (little endian)

Read API (dict-like, a few h5py idioms):          f = H5File(path); f.attrs['layer_names']; f['a/b/kernel:0'][()]
Write API:                                        write_h5(path, tree) with tree = Group(attrs, children) / numpy arrays

Variable-length strings (global heap, GCOL) are read (h5py stores `backend` / `keras_version` that way) and written as
fixed-length strings, which h5py / Keras read back identically.
Not supported (raises): chunked / compressed datasets, variable-length sequences, new-style (link-message) groups,
superblock versions >= 2.  Keras weight files need none of them.
"""
import struct
from collections import OrderedDict

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b'\x89HDF\r\n\x1a\n'


class H5FormatError(ValueError):
    pass


# =====================================================================================================================
# reader
# =====================================================================================================================
def _pad8(n):
    return (n + 7) & ~7


def _parse_datatype(buf, off=0):
    """-> (numpy dtype or ('S', n)), total size of the datatype message body."""
    b0 = buf[off]
    cls, ver = b0 & 0x0F, b0 >> 4
    bits0 = buf[off + 1]
    size = struct.unpack_from('<I', buf, off + 4)[0]
    if ver not in (1, 2, 3):
        raise H5FormatError('datatype version %d' % ver)
    if cls == 0:      # fixed point: byte order bit 0, signed bit 3
        order = '>' if bits0 & 1 else '<'
        signed = bool(bits0 & 8)
        return np.dtype('%s%s%d' % (order, 'i' if signed else 'u', size)), 8 + 4
    if cls == 1:      # floating point
        order = '>' if bits0 & 1 else '<'
        return np.dtype('%sf%d' % (order, size)), 8 + 12
    if cls == 3:      # fixed-length string (padding / charset in the bit field; no properties)
        return np.dtype('S%d' % size), 8
    if cls == 9:      # variable length: type 1 (bits0 & 0xF) = string; elements are (length, global heap address, index)
        if (bits0 & 0x0F) != 1:
            raise H5FormatError('variable-length sequences are not supported (only strings)')
        return _VLEN_STR, size
    raise H5FormatError('datatype class %d is not supported' % cls)


_VLEN_STR = np.dtype([('len', '<u4'), ('addr', '<u8'), ('idx', '<u4')])     # marker dtype of a variable-length string


def _parse_dataspace(buf, off=0):
    ver = buf[off]
    if ver == 1:
        rank, flags = buf[off + 1], buf[off + 2]
        p = off + 8
    elif ver == 2:
        rank, flags, typ = buf[off + 1], buf[off + 2], buf[off + 3]
        p = off + 4
        if typ == 2:      # null dataspace
            return None
    else:
        raise H5FormatError('dataspace version %d' % ver)
    dims = struct.unpack_from('<%dQ' % rank, buf, p) if rank else ()
    return tuple(int(d) for d in dims)


class _Node:
    def __init__(self, f, addr):
        self._f, self._addr = f, addr
        self._msgs = f._read_object_header(addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = OrderedDict()
            for typ, body in self._msgs:
                if typ == 0x000C:
                    name, val = self._f._parse_attribute(body)
                    self._attrs[name] = val
        return self._attrs


class H5Dataset(_Node):
    def __init__(self, f, addr):
        super().__init__(f, addr)
        self.shape = self.dtype = None
        self._layout = None
        for typ, body in self._msgs:
            if typ == 0x0001:
                self.shape = _parse_dataspace(body)
            elif typ == 0x0003:
                self.dtype, _ = _parse_datatype(body)
            elif typ == 0x0008:
                self._layout = body
        if self.dtype is None or self._layout is None:
            raise H5FormatError('object at %d is not a dataset' % addr)

    def __getitem__(self, key):
        arr = self._read()
        return arr if key == () or key is Ellipsis else arr[key]

    def _read(self):
        b = self._layout
        ver = b[0]
        n = int(np.prod(self.shape)) if self.shape else 1
        nbytes = n * self.dtype.itemsize
        if ver == 3:
            cls = b[1]
            if cls == 1:      # contiguous
                addr, size = struct.unpack_from('<QQ', b, 2)
                if addr == UNDEF:
                    raw = b'\0' * nbytes
                else:
                    raw = self._f._read(addr, nbytes)
            elif cls == 0:    # compact
                size = struct.unpack_from('<H', b, 2)[0]
                raw = bytes(b[4:4 + size])
            else:
                raise H5FormatError('chunked datasets are not supported')
        elif ver in (1, 2):
            rank, cls = b[1], b[2]
            if cls != 1:
                raise H5FormatError('layout v%d class %d is not supported' % (ver, cls))
            addr = struct.unpack_from('<Q', b, 8)[0]
            raw = self._f._read(addr, nbytes)
        else:
            raise H5FormatError('data layout version %d' % ver)
        arr = np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape if self.shape else ())
        return arr.astype(self.dtype.newbyteorder('='), copy=True)


class H5Group(_Node):
    def __init__(self, f, addr):
        super().__init__(f, addr)
        self._links = None

    def _load(self):
        if self._links is not None:
            return
        self._links = OrderedDict()
        for typ, body in self._msgs:
            if typ == 0x0011:
                btree, heap = struct.unpack_from('<QQ', body, 0)
                for name, oaddr in self._f._iter_symbol_table(btree, heap):
                    self._links[name] = oaddr
            elif typ in (0x0002, 0x0006):
                raise H5FormatError('new-style groups (link messages) are not supported')

    def keys(self):
        self._load()
        return list(self._links.keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, H5Group):
                raise KeyError(path)
            node._load()
            if part not in node._links:
                raise KeyError(path)
            node = node._f._open(node._links[part])
        return node


class H5File(H5Group):
    """read-only view of a classic-layout HDF5 file (whole file held in memory: weight files are ~50 MB)."""

    def __init__(self, path):
        with open(path, 'rb') as fh:
            self._buf = fh.read()
        b = self._buf
        if b[:8] != SIGNATURE:
            raise H5FormatError('%s is not an HDF5 file' % path)
        if b[8] != 0:
            raise H5FormatError('superblock version %d is not supported (classic version 0 only)' % b[8])
        if b[13] != 8 or b[14] != 8:
            raise H5FormatError('only 8-byte offsets / lengths are supported')
        self._base = struct.unpack_from('<Q', b, 24)[0]
        root_oh = struct.unpack_from('<Q', b, 56 + 8)[0]
        self._cache = {}
        super().__init__(self, root_oh)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- low level ---------------------------------------------------------------------------------------------
    def _read(self, addr, n):
        a = self._base + addr
        if a + n > len(self._buf):
            raise H5FormatError('read beyond the end of the file')
        return self._buf[a:a + n]

    def _open(self, addr):
        if addr not in self._cache:
            msgs = self._read_object_header(addr)
            types = {t for t, _ in msgs}
            self._cache[addr] = H5Dataset(self, addr) if 0x0008 in types else H5Group(self, addr)
        return self._cache[addr]

    def _read_object_header(self, addr):
        b = self._buf
        a = self._base + addr
        ver = b[a]
        if ver != 1:
            raise H5FormatError('object header version %d at %d is not supported' % (ver, addr))
        nmsg = struct.unpack_from('<H', b, a + 2)[0]
        hsize = struct.unpack_from('<I', b, a + 8)[0]
        blocks = [(a + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(msgs) < nmsg:
                typ, size, flags = struct.unpack_from('<HHB', b, p)
                body = b[p + 8:p + 8 + size]
                p += 8 + size
                if typ == 0x0010:
                    coff, clen = struct.unpack_from('<QQ', body, 0)
                    blocks.append((self._base + coff, clen))
                msgs.append((typ, body))
        return msgs

    def _parse_attribute(self, body):
        ver = body[0]
        nsz, tsz, ssz = struct.unpack_from('<HHH', body, 2)
        if ver == 1:
            p = 8
            name = bytes(body[p:p + nsz]).split(b'\0')[0].decode()
            p += _pad8(nsz)
            dt, _ = _parse_datatype(body, p)
            p += _pad8(tsz)
            shape = _parse_dataspace(body, p)
            p += _pad8(ssz)
        elif ver in (2, 3):
            p = 8 if ver == 2 else 9
            name = bytes(body[p:p + nsz]).split(b'\0')[0].decode()
            p += nsz
            dt, _ = _parse_datatype(body, p)
            p += tsz
            shape = _parse_dataspace(body, p)
            p += ssz
        else:
            raise H5FormatError('attribute version %d' % ver)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(bytes(body[p:p + n * dt.itemsize]), dtype=dt, count=n)
        if dt is _VLEN_STR:
            vals = [self._global_heap_object(int(e['addr']), int(e['idx']))[:int(e['len'])] for e in arr]
            arr = np.array(vals, dtype='S%d' % max([len(v) for v in vals] + [1]))
        else:
            arr = arr.astype(dt.newbyteorder('='), copy=True)
        if shape == () or shape is None:
            return name, arr[0]
        return name, arr.reshape(shape)

    def _global_heap_object(self, addr, index):
        b = self._buf
        a = self._base + addr
        if b[a:a + 4] != b'GCOL':
            raise H5FormatError('bad global heap signature')
        size = struct.unpack_from('<Q', b, a + 8)[0]
        p, end = a + 16, a + size
        while p + 16 <= end:
            idx, _ref, osize = struct.unpack_from('<HH4xQ', b, p)
            if idx == 0:
                break
            if idx == index:
                return bytes(b[p + 16:p + 16 + osize])
            p += 16 + _pad8(osize)
        raise H5FormatError('global heap object %d not found' % index)

    def _heap_string(self, heap_addr, off):
        b = self._buf
        a = self._base + heap_addr
        if b[a:a + 4] != b'HEAP':
            raise H5FormatError('bad local heap signature')
        data_addr = struct.unpack_from('<Q', b, a + 24)[0]
        s = self._base + data_addr + off
        e = b.index(b'\0', s)
        return b[s:e].decode()

    def _iter_symbol_table(self, btree_addr, heap_addr):
        b = self._buf
        a = self._base + btree_addr
        if b[a:a + 4] != b'TREE':
            raise H5FormatError('bad B-tree signature')
        ntype, level, used = b[a + 4], b[a + 5], struct.unpack_from('<H', b, a + 6)[0]
        if ntype != 0:
            raise H5FormatError('expected a group B-tree')
        p = a + 24                  # key0, child0, key1, child1, ...
        children = []
        for i in range(used):
            child = struct.unpack_from('<Q', b, p + 8)[0]
            children.append(child)
            p += 16
        for child in children:
            if level > 0:
                yield from self._iter_symbol_table(child, heap_addr)
            else:
                c = self._base + child
                if b[c:c + 4] != b'SNOD':
                    raise H5FormatError('bad symbol table node signature')
                nsym = struct.unpack_from('<H', b, c + 6)[0]
                q = c + 8
                for _ in range(nsym):
                    noff, oaddr = struct.unpack_from('<QQ', b, q)
                    yield self._heap_string(heap_addr, noff), oaddr
                    q += 40


# =====================================================================================================================
# writer
# =====================================================================================================================
class Group:
    """in-memory tree for write_h5: children is an ordered {name: Group | numpy array}, attrs {name: value}."""

    def __init__(self, attrs=None, children=None):
        self.attrs = OrderedDict(attrs or {})
        self.children = OrderedDict(children or {})


def _dt_message(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        size = dt.itemsize
        if size == 4:
            props = struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)          # little endian, mantissa normalisation 2 (implied msb), sign bit 31
        elif size == 8:
            props = struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)
        else:
            raise H5FormatError('float%d' % (8 * size))
        return struct.pack('<BBBBI', 0x11, bits[0], bits[1], bits[2], size) + props
    if dt.kind in 'iu':
        bits0 = 0x08 if dt.kind == 'i' else 0x00
        return struct.pack('<BBBBI', 0x10, bits0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, dt.itemsize)      # null-padded (numpy 'S'), ASCII -- as h5py writes
    raise H5FormatError('dtype %s is not supported by the writer' % dt)


def _ds_message(shape):
    shape = tuple(int(s) for s in shape)
    if not shape:
        return struct.pack('<BBBB4x', 1, 0, 0, 0)
    dims = b''.join(struct.pack('<Q', s) for s in shape)
    return struct.pack('<BBBB4x', 1, len(shape), 1, 0) + dims + dims        # flag 1: maximum dimensions present (= dims)


def _msg(typ, body, flags=0):
    body = body + b'\0' * (_pad8(len(body)) - len(body))
    return struct.pack('<HHB3x', typ, len(body), flags) + body


def _attr_message(name, value):
    if isinstance(value, str):
        value = np.bytes_(value.encode())
    if isinstance(value, (bytes, np.bytes_)):
        value = np.asarray(value, dtype='S%d' % max(len(value), 1))
    value = np.asarray(value)
    if value.dtype.kind == 'U':
        value = np.char.encode(value, 'utf-8')
    if value.dtype.kind == 'O':
        raise H5FormatError('object arrays cannot be stored')
    if value.dtype.kind == 'S' and value.dtype.itemsize == 0:
        value = value.astype('S1')
    nm = name.encode() + b'\0'
    dt = _dt_message(value.dtype)
    ds = _ds_message(value.shape)
    body = struct.pack('<BxHHH', 1, len(nm), len(dt), len(ds))
    body += nm + b'\0' * (_pad8(len(nm)) - len(nm))
    body += dt + b'\0' * (_pad8(len(dt)) - len(dt))
    body += ds + b'\0' * (_pad8(len(ds)) - len(ds))
    body += np.ascontiguousarray(value).astype(value.dtype.newbyteorder('<')).tobytes()
    return _msg(0x000C, body)


class _Writer:
    LEAF_K, INTERNAL_K = 4, 16

    def __init__(self):
        self.buf = bytearray(96)            # superblock (56 bytes) + root symbol table entry (40 bytes)

    def _alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _object_header(self, messages):
        body = b''.join(messages)
        hdr = struct.pack('<BxHII4x', 1, len(messages), 1, len(body))
        return self._alloc(hdr + body)

    def write_dataset(self, arr, attrs):
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == '>':
            arr = arr.astype(arr.dtype.newbyteorder('<'))
        raw = arr.tobytes()
        daddr = self._alloc(raw) if raw else UNDEF
        msgs = [_msg(0x0001, _ds_message(arr.shape)),
                _msg(0x0003, _dt_message(arr.dtype), flags=1),
                _msg(0x0005, struct.pack('<BBBBI', 2, 2, 2, 1, 0)),      # fill value v2: late alloc, write if set, default (size 0)
                _msg(0x0008, struct.pack('<BBQQ', 3, 1, daddr, len(raw)))]
        msgs += [_attr_message(k, v) for k, v in attrs.items()]
        return self._object_header(msgs)

    def write_group(self, group):
        """-> (object header address, btree address, heap address)"""
        entries = []
        for name, child in group.children.items():
            if isinstance(child, Group):
                oaddr, bt, hp = self.write_group(child)
                entries.append((name, oaddr, 1, struct.pack('<QQ', bt, hp)))
            else:
                arr, attrs = (child if isinstance(child, tuple) else (child, {}))
                entries.append((name, self.write_dataset(arr, attrs), 0, b'\0' * 16))
        entries.sort(key=lambda e: e[0].encode())             # B-tree order = strcmp order of the link names
        # local heap: offset 0 holds the empty string (key of the left-most B-tree edge)
        heap = bytearray(b'\0' * 8)
        offs = []
        for name, *_ in entries:
            offs.append(len(heap))
            nb = name.encode() + b'\0'
            heap += nb + b'\0' * (_pad8(len(nb)) - len(nb))
        free_off = len(heap)
        heap += struct.pack('<QQ', 1, 16)                      # one free block: next = 1 (none), size 16
        heap_data = self._alloc(bytes(heap))
        heap_addr = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free_off, heap_data))
        # symbol table nodes (<= 2 * LEAF_K symbols each)
        cap = 2 * self.LEAF_K
        snods, keys = [], [0]
        for i in range(0, len(entries), cap):
            chunk = entries[i:i + cap]
            node = b'SNOD' + struct.pack('<BxH', 1, len(chunk))
            for j, (name, oaddr, ctype, scratch) in enumerate(chunk):
                node += struct.pack('<QQII', offs[i + j], oaddr, ctype, 0) + scratch
            node += b'\0' * (40 * (cap - len(chunk)))
            snods.append(self._alloc(node))
            keys.append(offs[i + len(chunk) - 1])
        if len(snods) > 2 * self.INTERNAL_K:
            raise H5FormatError('too many links in one group for a single-level B-tree (%d)' % len(entries))
        tree = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF)
        for i, s in enumerate(snods):
            tree += struct.pack('<QQ', keys[i], s)
        tree += struct.pack('<Q', keys[len(snods)])
        tree += b'\0' * ((2 * self.INTERNAL_K - len(snods)) * 16)
        bt_addr = self._alloc(tree)
        msgs = [_msg(0x0011, struct.pack('<QQ', bt_addr, heap_addr))]
        msgs += [_attr_message(k, v) for k, v in group.attrs.items()]
        return self._object_header(msgs), bt_addr, heap_addr

    def finish(self, root):
        oaddr, bt, hp = self.write_group(root)
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack('<BBBBBBBxHHI', 0, 0, 0, 0, 0, 8, 8, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQII', 0, oaddr, 1, 0) + struct.pack('<QQ', bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path, root):
    """atomic: a crash mid-write must not leave a truncated checkpoint under the final name"""
    import os
    data = _Writer().finish(root)
    tmp = '%s.tmp%d' % (path, os.getpid())
    with open(tmp, 'wb') as fh:
        fh.write(data)
        fh.flush()
        os.fsync(fh.fileno())
    os.replace(tmp, path)


# =====================================================================================================================
# Keras weight files
# =====================================================================================================================
def load_keras_weights(path):
    """-> (OrderedDict {'<layer>/<weight>': array} with the ':0' suffixes stripped, root attributes).  Handles both
    `save_weights` files (layers at the root) and full `model.save` files (layers under /model_weights)."""
    f = H5File(path)
    g = f['model_weights'] if 'model_weights' in f.keys() else f
    out = OrderedDict()
    names = [n.decode() if isinstance(n, bytes) else str(n) for n in np.atleast_1d(g.attrs['layer_names'])]
    for lname in names:
        lg = g[lname]
        if 'weight_names' not in lg.attrs:
            continue
        for wn in np.atleast_1d(lg.attrs['weight_names']):
            wn = wn.decode() if isinstance(wn, bytes) else str(wn)
            arr = lg[wn][()]
            # keyed by the LAYER name (what load_weights(by_name=True) matches on): TensorFlow may have uniquified the
            # variable scope inside weight_names ('unet_bn_down_1_1/gamma:0' in layer 'unet_bn_down_1')
            out['%s/%s' % (lname, wn.split('/')[-1].split(':')[0])] = arr
    return out, dict(g.attrs)


def keras_weights_group(weights, layer_order=None, keras_version='2.3.1', backend='tensorflow'):
    """{'<layer>/<weight>': array} -> Group laid out like keras.engine.saving.save_weights_to_hdf5_group: attrs
    layer_names / backend / keras_version; per layer a group with attr weight_names and the datasets at
    '<layer>/<layer>/<weight>:0' (weightless layers listed in layer_order get an empty group, as Keras writes them)."""
    layers = OrderedDict()
    for key, arr in weights.items():
        lname, wname = key.split('/', 1)
        layers.setdefault(lname, OrderedDict())[wname] = np.asarray(arr)
    order = list(layer_order) if layer_order is not None else list(layers.keys())
    for l in layers:
        if l not in order:
            order.append(l)
    root = Group()
    width = max(len(n) for n in order)
    root.attrs['layer_names'] = np.array([n.encode() for n in order], dtype='S%d' % width)
    root.attrs['backend'] = np.bytes_(backend.encode())
    root.attrs['keras_version'] = np.bytes_(keras_version.encode())
    for lname in order:
        lg = Group()
        ws = layers.get(lname, OrderedDict())
        wnames = ['%s/%s:0' % (lname, w) for w in ws]
        if wnames:
            lg.attrs['weight_names'] = np.array([w.encode() for w in wnames], dtype='S%d' % max(len(w) for w in wnames))
            inner = Group()
            for w, arr in ws.items():
                inner.children['%s:0' % w] = arr
            lg.children[lname] = inner
        else:
            lg.attrs['weight_names'] = np.zeros((0,), dtype=np.float64)      # what h5py stores for an empty list
        root.children[lname] = lg
    return root


def save_keras_weights(path, weights, layer_order=None, extra=None, full_model=False):
    """weights: {'<layer>/<weight>': array} -> a file `keras.Model.load_weights(path, by_name=True)` understands.
    full_model=False: the `save_weights` layout (layers at the root).  full_model=True: the `model.save` /
    ModelCheckpoint layout the reference's training writes (SynthSR/training.py:429): layers under /model_weights,
    keras_version / backend at the root.  extra: {name: array} stored as datasets of the group /optimizer_weights
    (this engine's flat Adam state; Keras ignores what it does not list in that group's weight_names)."""
    wg = keras_weights_group(weights, layer_order)
    if full_model:
        root = Group()
        root.attrs['keras_version'] = wg.attrs['keras_version']
        root.attrs['backend'] = wg.attrs['backend']
        root.children['model_weights'] = wg
    else:
        root = wg
    if extra:
        og = Group()
        og.attrs['weight_names'] = np.zeros((0,), dtype=np.float64)
        for k, v in extra.items():
            og.children[k] = np.asarray(v)
        root.children['optimizer_weights'] = og
    write_h5(path, root)


def load_extra(path, group='optimizer_weights'):
    """datasets stored by save_keras_weights(extra=...) -> {name: array} ({} when the group is absent)."""
    f = H5File(path)
    if group not in f.keys():
        return {}
    g = f[group]
    return {k: g[k][()] for k in g.keys() if isinstance(g[k], H5Dataset)}
