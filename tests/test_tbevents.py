"""TensorBoard event writer (synthsr_b200/tbevents.py): record framing + the few protobuf fields, no TensorFlow."""
import struct

from synthsr_b200 import tbevents as T


def test_crc32c_known_answer():
    assert T.crc32c(b'123456789') == 0xE3069283            # the CRC-32C (Castagnoli) check value
    assert T.crc32c(b'') == 0


def _fields(buf):
    """minimal protobuf walker -> [(field, wire type, value)]"""
    out, o = [], 0
    while o < len(buf):
        key, sh = 0, 0
        while True:
            b = buf[o]; o += 1
            key |= (b & 0x7F) << sh; sh += 7
            if not b & 0x80:
                break
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, sh = 0, 0
            while True:
                b = buf[o]; o += 1
                v |= (b & 0x7F) << sh; sh += 7
                if not b & 0x80:
                    break
        elif wt == 1:
            v = struct.unpack_from('<d', buf, o)[0]; o += 8
        elif wt == 5:
            v = struct.unpack_from('<f', buf, o)[0]; o += 4
        else:
            n, sh = 0, 0
            while True:
                b = buf[o]; o += 1
                n |= (b & 0x7F) << sh; sh += 7
                if not b & 0x80:
                    break
            v = bytes(buf[o:o + n]); o += n
        out.append((f, wt, v))
    return out


def test_event_file_round_trip(tmp_path):
    w = T.EventWriter(str(tmp_path))
    w.scalar('epoch_loss', 0.125, 0)
    w.scalar('epoch_loss', 0.0625, 300)
    w.close()
    recs = T.read_records(w.path)                            # checks both masked CRCs of every record
    assert len(recs) == 3
    first = dict((f, v) for f, _, v in _fields(recs[0]))
    assert first[3] == b'brain.Event:2' and first[1] > 1e9
    ev = _fields(recs[2])
    assert [(f, wt) for f, wt, _ in ev] == [(1, 1), (2, 0), (5, 2)] and ev[1][2] == 300
    value = _fields(_fields(ev[2][2])[0][2])
    assert value[0] == (1, 2, b'epoch_loss') and value[1][:2] == (2, 5) and value[1][2] == 0.0625
    assert 'events.out.tfevents.' in w.path
