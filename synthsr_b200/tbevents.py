"""TensorBoard event files without TensorFlow: the scalars Keras' TensorBoard callback writes for the reference's
training (SynthSR/training.py:425-431: `epoch_loss` per epoch under model_dir/logs).

File format (public, stable): a TFRecord stream -- per record  uint64 length | masked crc32c(length) | payload |
masked crc32c(payload)  -- of serialised `Event` protocol buffers:
    Event   { double wall_time = 1; int64 step = 2; string file_version = 3; Summary summary = 5; }
    Summary { repeated Value value = 1; }      Value { string tag = 1; float simple_value = 2; }
The few fields needed are encoded by hand (varint / fixed64 / fixed32 / length-delimited)."""
import os
import socket
import struct
import time

_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field, payload):                       # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_event(wall_time, step=None, file_version=None, scalars=None):
    ev = _varint((1 << 3) | 1) + struct.pack('<d', wall_time)
    if step is not None:
        ev += _varint((2 << 3) | 0) + _varint(int(step))
    if file_version is not None:
        ev += _ld(3, file_version.encode())
    if scalars:
        summary = b''
        for tag, value in scalars:
            summary += _ld(1, _ld(1, tag.encode()) + _varint((2 << 3) | 5) + struct.pack('<f', float(value)))
        ev += _ld(5, summary)
    return ev


def encode_record(payload):
    head = struct.pack('<Q', len(payload))
    return head + struct.pack('<I', _masked_crc(head)) + payload + struct.pack('<I', _masked_crc(payload))


def read_records(path):
    """-> list of payloads (checks both checksums); used by the tests."""
    out, raw, o = [], open(path, 'rb').read(), 0
    while o < len(raw):
        n, = struct.unpack_from('<Q', raw, o)
        assert struct.unpack_from('<I', raw, o + 8)[0] == _masked_crc(raw[o:o + 8])
        payload = raw[o + 12:o + 12 + n]
        assert struct.unpack_from('<I', raw, o + 12 + n)[0] == _masked_crc(payload)
        out.append(payload)
        o += 16 + n
    return out


class EventWriter:
    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, 'events.out.tfevents.%010d.%s' % (int(time.time()), socket.gethostname()))
        self.f = open(self.path, 'ab')
        self.f.write(encode_record(encode_event(time.time(), file_version='brain.Event:2')))
        self.f.flush()

    def scalar(self, tag, value, step):
        self.f.write(encode_record(encode_event(time.time(), step=step, scalars=[(tag, value)])))
        self.f.flush()

    def close(self):
        self.f.close()
