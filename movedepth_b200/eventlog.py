"""Tensorboard event files without the tensorboard / tensorboardX packages.

The reference logs through `tensorboardX.SummaryWriter` (movedepth/trainer.py:147-151, 772-793): scalars per loss and a
few images per log step.  Neither package is in this image, so this module writes the same on-disk format directly:
TFRecord framing (length, masked CRC-32C of the length, payload, masked CRC-32C of the payload) around hand-encoded
`Event` / `Summary` protocol-buffer messages, images as zlib-compressed PNG.  `tensorboard --logdir <log_path>` reads the
result.  `read_events` is the inverse (used by the tests and by `tools/`), so the files can be checked without tensorboard.
"""
import os
import socket
import struct
import time
import zlib

import numpy as np

# ------------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data):
    c = 0xFFFFFFFF
    t = _CRC_TABLE
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------- protobuf wire format
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field_bytes(num, payload):
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def _field_varint(num, v):
    return _varint(num << 3) + _varint(v)


def _field_f32(num, v):
    return _varint((num << 3) | 5) + struct.pack("<f", v)


def _field_f64(num, v):
    return _varint((num << 3) | 1) + struct.pack("<d", v)


def _event(wall_time, step, file_version=None, summary=None):
    ev = _field_f64(1, wall_time) + _field_varint(2, step)
    if file_version is not None:
        ev += _field_bytes(3, file_version.encode())
    if summary is not None:
        ev += _field_bytes(5, summary)
    return ev


def _scalar_summary(tag, value):
    return _field_bytes(1, _field_bytes(1, tag.encode()) + _field_f32(2, float(value)))


def encode_png(img):
    """uint8 [H, W, 3] or [H, W] -> PNG bytes (8-bit, no interlace, filter 0, zlib level 3)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    color = 2 if img.ndim == 3 else 0
    raw = np.concatenate([np.zeros((h, 1), np.uint8), img.reshape(h, -1)], 1).tobytes()

    def chunk(kind, data):
        body = kind + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw, 3)) + chunk(b"IEND", b""))


def _image_summary(tag, img):
    h, w = img.shape[:2]
    image = (_field_varint(1, h) + _field_varint(2, w) + _field_varint(3, 3 if img.ndim == 3 else 1)
             + _field_bytes(4, encode_png(img)))
    return _field_bytes(1, _field_bytes(1, tag.encode()) + _field_bytes(4, image))


# ------------------------------------------------------------------------------------------------- the writer
class SummaryWriter:
    """The subset of tensorboardX.SummaryWriter the reference uses: add_scalar, add_image (CHW float in [0,1] or uint8),
    flush, close."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        name = "events.out.tfevents.%010d.%s.%d" % (int(time.time()), socket.gethostname(), os.getpid())
        self.path = os.path.join(logdir, name)
        self._f = open(self.path, "ab")
        self._record(_event(time.time(), 0, file_version="brain.Event:2"))

    def _record(self, payload):
        head = struct.pack("<Q", len(payload))
        self._f.write(head + struct.pack("<I", masked_crc32c(head)) + payload + struct.pack("<I", masked_crc32c(payload)))

    def add_scalar(self, tag, value, step):
        self._record(_event(time.time(), int(step), summary=_scalar_summary(tag, float(value))))

    def add_image(self, tag, img, step):
        if hasattr(img, "detach"):
            img = img.detach().float().cpu().numpy()
        img = np.asarray(img)
        if img.ndim == 3:
            img = np.moveaxis(img, 0, -1)                           # CHW -> HWC
            if img.shape[-1] == 1:
                img = img[..., 0]
        if img.dtype != np.uint8:
            img = (np.clip(img, 0.0, 1.0) * 255.0).astype(np.uint8)
        self._record(_event(time.time(), int(step), summary=_image_summary(tag, img)))

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()


# ------------------------------------------------------------------------------------------------- the reader
def _read_varint(buf, pos):
    v, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return v, pos


def _parse(buf):
    """One protobuf message -> {field number: [raw values]} (varint -> int, 32/64-bit -> bytes, length-delimited -> bytes)."""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        else:
            raise ValueError("unsupported wire type %d" % wt)
        out.setdefault(num, []).append(v)
    return out


def read_events(path):
    """Yields dicts {wall_time, step, file_version | scalars {tag: value} | images {tag: (h, w, channels, png bytes)}};
    verifies both CRCs of every record."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        head = data[pos:pos + 8]
        (n,) = struct.unpack("<Q", head)
        if struct.unpack("<I", data[pos + 8:pos + 12])[0] != masked_crc32c(head):
            raise ValueError("bad length CRC at %d" % pos)
        payload = data[pos + 12:pos + 12 + n]
        if struct.unpack("<I", data[pos + 12 + n:pos + 16 + n])[0] != masked_crc32c(payload):
            raise ValueError("bad payload CRC at %d" % pos)
        pos += 16 + n
        ev = _parse(payload)
        rec = {"wall_time": struct.unpack("<d", ev[1][0])[0], "step": ev.get(2, [0])[0], "scalars": {}, "images": {}}
        if 3 in ev:
            rec["file_version"] = ev[3][0].decode()
        for summary in ev.get(5, []):
            for value in _parse(summary).get(1, []):
                v = _parse(value)
                tag = v[1][0].decode()
                if 2 in v:
                    rec["scalars"][tag] = struct.unpack("<f", v[2][0])[0]
                if 4 in v:
                    im = _parse(v[4][0])
                    rec["images"][tag] = (im[1][0], im[2][0], im[3][0], im[4][0])
        yield rec


# ------------------------------------------------------------------------------------------------- depth colour map
# movedepth/trainer.py:881-911 colours disparities with matplotlib's 'plasma' (not in this image): piecewise-linear
# interpolation through its eleven decile colours (maximum deviation from the 256-entry table below one 8-bit level in
# practice; visualisation only).
_PLASMA = np.array([(13, 8, 135), (65, 4, 157), (106, 0, 168), (143, 13, 164), (177, 42, 144), (204, 71, 120),
                    (225, 100, 98), (242, 132, 75), (252, 166, 54), (252, 206, 37), (240, 249, 33)], np.float32) / 255.0


def colormap(x, normalize=True):
    """[H, W] tensor / array -> float32 [3, H, W] in [0, 1] (movedepth/trainer.py:883-911 for the 2-D case)."""
    if hasattr(x, "detach"):
        x = x.detach().float().cpu().numpy()
    x = np.asarray(x, np.float32)
    if normalize:
        ma, mi = float(x.max()), float(x.min())
        x = (x - mi) / (ma - mi if ma != mi else 1e5)
    t = np.clip(x, 0.0, 1.0) * (len(_PLASMA) - 1)
    i = np.minimum(t.astype(np.int32), len(_PLASMA) - 2)
    f = (t - i)[..., None]
    rgb = _PLASMA[i] * (1.0 - f) + _PLASMA[i + 1] * f
    return np.moveaxis(rgb, -1, 0)
