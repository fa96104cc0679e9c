"""Minimal NIfTI-1 (.nii / .nii.gz) and FreeSurfer .mgz readers/writers (nibabel is not available in this image).

Only what the hot path needs (ext/lab2im/utils.py:76-160 uses nib.load(...).get_fdata()/.affine/.header and
nib.save): single-file NIfTI-1, little or big endian, scalar datatypes, sform/qform/pixdim affines, scl_slope/inter.
"""
import gzip
import struct

import numpy as np

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_NIFTI_CODES = {np.dtype(v).str[1:]: k for k, v in _NIFTI_DTYPES.items()}


class Header(dict):
    """dict-like header: header['pixdim'], header['dim'], ... (the subset of nibabel's mapping the reference reads)."""

    def copy(self):
        return Header({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.items()})

    def get_zooms(self):
        nd = int(self['dim'][0])
        return tuple(float(v) for v in self['pixdim'][1:nd + 1])

    def set_zooms(self, zooms):
        pd = np.array(self['pixdim'], dtype=np.float32)
        pd[1:1 + len(zooms)] = zooms
        self['pixdim'] = pd


def blank_header():
    return Header(dim=np.array([3, 1, 1, 1, 1, 1, 1, 1], np.int16), pixdim=np.ones(8, np.float32), datatype=16,
                  bitpix=32, qform_code=0, sform_code=2, scl_slope=1., scl_inter=0., descrip=b'synthsr_b200')


def _open(path, mode='rb'):
    return gzip.open(path, mode) if path.endswith('.gz') else open(path, mode)


def _quat_to_affine(h):
    b, c, d = h['quatern_b'], h['quatern_c'], h['quatern_d']
    a = np.sqrt(max(0., 1. - (b * b + c * c + d * d)))
    r = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    qfac = -1. if h['pixdim'][0] < 0 else 1.
    zooms = np.array(h['pixdim'][1:4], dtype=np.float64) * [1, 1, qfac]
    aff = np.eye(4)
    aff[:3, :3] = r * zooms
    aff[:3, 3] = [h['qoffset_x'], h['qoffset_y'], h['qoffset_z']]
    return aff


def load_nifti(path):
    """-> (data float64 array in F-order semantics [x,y,z(,t)], affine 4x4, Header)."""
    with _open(path) as f:
        raw = f.read()
    end = '<'
    if struct.unpack('<i', raw[:4])[0] != 348:
        end = '>'
        assert struct.unpack('>i', raw[:4])[0] == 348, 'not a NIfTI-1 file: %s' % path
    h = Header()
    h['dim'] = np.array(struct.unpack(end + '8h', raw[40:56]), np.int16)
    h['datatype'], h['bitpix'] = struct.unpack(end + '2h', raw[70:74])
    h['pixdim'] = np.array(struct.unpack(end + '8f', raw[76:108]), np.float32)
    vox_offset, slope, inter = struct.unpack(end + '3f', raw[108:120])
    h['scl_slope'], h['scl_inter'] = slope, inter
    h['descrip'] = raw[148:228]
    h['qform_code'], h['sform_code'] = struct.unpack(end + '2h', raw[252:256])
    (h['quatern_b'], h['quatern_c'], h['quatern_d'], h['qoffset_x'], h['qoffset_y'],
     h['qoffset_z']) = struct.unpack(end + '6f', raw[256:280])
    srow = np.array(struct.unpack(end + '12f', raw[280:328]), np.float64).reshape(3, 4)
    h['srow'] = srow
    nd = int(h['dim'][0])
    shape = [int(v) for v in h['dim'][1:nd + 1]]
    dt = np.dtype(_NIFTI_DTYPES[int(h['datatype'])]).newbyteorder(end)
    off = int(vox_offset) if vox_offset >= 352 else 352
    data = np.frombuffer(raw, dtype=dt, count=int(np.prod(shape)), offset=off).reshape(shape, order='F')
    data = data.astype(np.float64)
    if slope not in (0., 1.) and np.isfinite(slope):
        data = data * slope + inter
    elif inter not in (0.,) and np.isfinite(inter) and slope == 1.:
        data = data + inter
    if h['sform_code'] > 0:
        aff = np.vstack([srow, [0, 0, 0, 1]])
    elif h['qform_code'] > 0:
        aff = _quat_to_affine(h)
    else:
        aff = np.diag(list(h['pixdim'][1:4]) + [1.]).astype(np.float64)
    return data, aff, h


def save_nifti(path, volume, aff, header=None, dtype=None):
    volume = np.asarray(volume)
    if dtype is not None:
        volume = volume.astype(dtype)
    if volume.dtype == np.float64 and dtype is None:
        volume = volume.astype(np.float32)
    if volume.dtype == bool:
        volume = volume.astype(np.uint8)
    code = _NIFTI_CODES[volume.dtype.str[1:]]
    h = header.copy() if isinstance(header, Header) else blank_header()
    aff = np.eye(4) if aff is None else np.asarray(aff, dtype=np.float64)
    dim = np.ones(8, np.int16)
    dim[0] = volume.ndim
    dim[1:1 + volume.ndim] = volume.shape
    pixdim = np.array(h.get('pixdim', np.ones(8)), dtype=np.float32).copy()
    if 'pixdim' not in (header or {}):
        pixdim[1:4] = np.sqrt((aff[:3, :3] ** 2).sum(0))
    buf = bytearray(352)
    struct.pack_into('<i', buf, 0, 348)
    struct.pack_into('<8h', buf, 40, *[int(v) for v in dim])
    struct.pack_into('<2h', buf, 70, code, volume.dtype.itemsize * 8)
    struct.pack_into('<8f', buf, 76, *[float(v) for v in pixdim])
    struct.pack_into('<3f', buf, 108, 352., 1., 0.)
    struct.pack_into('<B', buf, 123, 2)                    # xyzt_units: mm
    buf[148:148 + 12] = b'synthsr_b200'
    struct.pack_into('<2h', buf, 252, 0, 2)                # sform only (aligned)
    struct.pack_into('<12f', buf, 280, *[float(v) for v in aff[:3].reshape(-1)])
    buf[344:348] = b'n+1\x00'
    payload = bytes(buf) + np.asfortranarray(volume).astype(volume.dtype.newbyteorder('<')).tobytes(order='F')
    with _open(path, 'wb') as f:
        f.write(payload)


def load_mgz(path):
    """FreeSurfer .mgh/.mgz (big endian, 284-byte header)."""
    with _open(path if path.endswith('.gz') or path.endswith('.mgh') else path) as f:
        raw = f.read()
    if raw[:2] == b'\x1f\x8b':
        raw = gzip.decompress(raw)
    ver, w, hh, d, nf, typ, dof = struct.unpack('>7i', raw[:28])
    good = struct.unpack('>h', raw[28:30])[0]
    delta = np.ones(3)
    mdc, c = np.eye(3), np.zeros(3)
    if good:
        delta = np.array(struct.unpack('>3f', raw[30:42]), np.float64)
        mdc = np.array(struct.unpack('>9f', raw[42:78]), np.float64).reshape(3, 3).T
        c = np.array(struct.unpack('>3f', raw[78:90]), np.float64)
    dt = {0: '>u1', 4: '>i2', 1: '>i4', 3: '>f4'}[typ]
    shape = [w, hh, d] + ([nf] if nf > 1 else [])
    data = np.frombuffer(raw, dtype=dt, count=int(np.prod(shape)), offset=284).reshape(shape, order='F').astype(np.float64)
    m = mdc * delta
    aff = np.eye(4)
    aff[:3, :3] = m
    aff[:3, 3] = c - m @ (np.array([w, hh, d]) / 2.)
    h = Header(dim=np.array([len(shape)] + shape + [1] * (7 - len(shape)), np.int16), delta=delta,
               pixdim=np.array([1, *delta, 1, 1, 1, 1], np.float32))
    return data, aff, h


def save_mgz(path, volume, aff, header=None, dtype=None):
    """FreeSurfer .mgh / .mgz writer (what nib.save does for those extensions, ext/lab2im/utils.py:158-160): 284-byte
    big-endian header (version 1, dims, type, goodRASFlag, voxel sizes, direction cosines, centre RAS), voxels in Fortran
    order, gzip for .mgz.  MGH stores uint8 / int16 / int32 / float32 only; other dtypes are converted to float32."""
    volume = np.asarray(volume)
    if dtype is not None:
        volume = volume.astype(dtype)
    if volume.ndim > 4:
        raise ValueError('MGH files hold at most 4 dimensions')
    codes = {'u1': 0, 'i2': 4, 'i4': 1, 'f4': 3}
    if volume.dtype.str[1:] not in codes:
        volume = volume.astype(np.int32 if (volume.dtype.kind in 'iub' and volume.dtype.itemsize <= 4) else np.float32)
    typ = codes[volume.dtype.str[1:]]
    aff = np.eye(4) if aff is None else np.asarray(aff, dtype=np.float64)
    shape3 = list(volume.shape[:3]) + [1] * (3 - min(volume.ndim, 3))
    nf = volume.shape[3] if volume.ndim == 4 else 1
    delta = np.sqrt((aff[:3, :3] ** 2).sum(0))
    delta[delta == 0] = 1.
    mdc = aff[:3, :3] / delta
    c = aff[:3, :3] @ (np.array(shape3, dtype=np.float64) / 2.) + aff[:3, 3]
    buf = bytearray(284)
    struct.pack_into('>7i', buf, 0, 1, shape3[0], shape3[1], shape3[2], nf, typ, 0)
    struct.pack_into('>h', buf, 28, 1)
    struct.pack_into('>3f', buf, 30, *[float(v) for v in delta])
    struct.pack_into('>9f', buf, 42, *[float(v) for v in mdc.T.reshape(-1)])
    struct.pack_into('>3f', buf, 78, *[float(v) for v in c])
    payload = bytes(buf) + np.asfortranarray(volume).astype(volume.dtype.newbyteorder('>')).tobytes(order='F')
    payload += struct.pack('>4f', 0., 0., 0., 0.)            # footer: TR, flip angle, TE, TI
    if path.endswith('.mgz') or path.endswith('.gz'):
        with gzip.open(path, 'wb') as f:
            f.write(payload)
    else:
        with open(path, 'wb') as f:
            f.write(payload)
