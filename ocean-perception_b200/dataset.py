"""EuRoC-layout stereo sequences for the PatchMatch path (host-side mirror of the reference's reader).

Mirrors, for the stereo part only (citations relative to /root/reference):
  dataset::EurocDataset            src/vehicle/dataset/euroc_dataset.cpp:11-18, 119-168
      <top>/mav0/cam0/data.csv, <top>/mav0/cam1/data.csv: a header line, then `timestamp[ns],...`;
      the image of a line is <cam>/data/<timestamp>.png (the reader uses the timestamp, not the file
      name column, :157-159); left and right must have equal counts and timestamps (:124-129).
  DataProvider::Playback           src/vehicle/dataset/data_provider.cpp:167-191
      steps through the items in time order, sleeping (dt / speed) between them, calling the stereo
      callback with the pair.
  MaybeConvertToGray + cv::resize  test/stereo_matching/patchmatch_gpu_test.cpp:124-129
      (the `PatchmatchGpuTest.Sequence` driver: gray, half size, Match).
IMU, pose, depth and range streams of the reference's reader are not on this path.

PNG decoding uses cv2 when it is importable and a small zlib-based decoder otherwise (8-bit gray / RGB /
RGBA, non-interlaced), so the module has no hard dependency beyond numpy.
"""
import os
import struct
import time
import zlib

import numpy as np


# ----------------------------------------------------------------------------- PNG (8-bit, no interlace)

def _paeth(a, b, c):
    p = a.astype(np.int32) + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c)).astype(np.uint8)


def read_png(path):
    """uint8 array [h, w] (gray) or [h, w, 3] in R, G, B order (alpha dropped)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("%s is not a PNG file" % path)
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    w, h, depth, ctype, _, _, interlace = hdr
    ch = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if depth != 8 or ch is None or interlace:
        raise ValueError("%s: only 8-bit non-interlaced gray/RGB(A) PNGs are supported" % path)
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8).reshape(h, 1 + w * ch)
    out = np.zeros((h, w * ch), np.uint8)
    prev = np.zeros(w * ch, np.uint8)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:]
        if ft == 0:
            cur = line.copy()
        elif ft == 2:
            cur = line + prev
        else:  # 1 (Sub), 3 (Average), 4 (Paeth) depend on the pixel to the left: channel-wise scan
            cur = np.zeros(w * ch, np.uint8)
            l = np.zeros(ch, np.uint8)
            ul = np.zeros(ch, np.uint8)
            for x in range(w):
                s = slice(x * ch, (x + 1) * ch)
                if ft == 1:
                    v = line[s] + l
                elif ft == 3:
                    v = line[s] + ((l.astype(np.int32) + prev[s]) >> 1).astype(np.uint8)
                else:
                    v = line[s] + _paeth(l, prev[s], ul)
                cur[s] = v
                l, ul = v, prev[s].copy()
        out[y] = cur
        prev = cur
    img = out.reshape(h, w, ch)
    if ch == 1:
        return img[:, :, 0]
    if ch == 2:
        return img[:, :, 0]
    return img[:, :, :3]


def write_png(path, img):
    """8-bit gray [h, w] or RGB [h, w, 3], filter 0 (test fixtures and tools)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    raw = np.zeros((h, 1 + w * ch), np.uint8)
    raw[:, 1:] = img.reshape(h, w * ch)

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if ch == 1 else 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))


def imread(path):
    """Decoded image as OpenCV would hand it over: gray [h, w] or B, G, R [h, w, 3]."""
    try:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is not None and img.dtype == np.uint8:
            return img[:, :, :3] if img.ndim == 3 and img.shape[2] == 4 else img
    except ImportError:
        pass
    img = read_png(path)
    return img if img.ndim == 2 else np.ascontiguousarray(img[:, :, ::-1])


def maybe_convert_to_gray(img):
    """MaybeConvertToGray (vision_core/image_util.cpp:52-61): cv::cvtColor(BGR2GRAY) on 8-bit images,
    i.e. (B*1868 + G*9617 + R*4899 + 2^13) >> 14 (OpenCV's 14-bit fixed-point coefficients)."""
    if img.ndim == 2:
        return np.ascontiguousarray(img, np.uint8)
    b, g, r = (img[:, :, i].astype(np.int32) for i in range(3))
    return ((b * 1868 + g * 9617 + r * 4899 + 8192) >> 14).astype(np.uint8)


def resize_half(gray):
    """cv::resize(img, size / 2) with the default INTER_LINEAR on u8 at an exact factor 2:
    (a + b + c + d + 2) >> 2 over 2x2 blocks; odd trailing rows/columns are dropped
    only when the size is even-halvable exactly (else the reference's generic bilinear applies)."""
    h, w = gray.shape
    if (h | w) & 1:
        raise ValueError("resize_half: %dx%d is not an exact halving" % (w, h))
    a = gray.astype(np.int32)
    return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)


# ----------------------------------------------------------------------------- EuRoC layout

class StereoDatasetItem:
    def __init__(self, timestamp, path_left, path_right):
        self.timestamp, self.path_left, self.path_right = timestamp, path_left, path_right


def _parse_image_folder(cam_folder):
    """EurocDataset::ParseImageFolder, euroc_dataset.cpp:137-166."""
    csv = os.path.join(cam_folder, "data.csv")
    if not os.path.exists(csv):
        raise FileNotFoundError("Cannot open file: " + csv)
    stamps, files = [], []
    with open(csv) as f:
        f.readline()                                   # header
        for line in f:
            line = line.strip()
            if not line:
                continue
            ts = int(line.split(",", 1)[0])
            stamps.append(ts)
            files.append(os.path.join(cam_folder, "data", "%d.png" % ts))
    return stamps, files


class EurocDataset:
    """Stereo part of dataset::EurocDataset + DataProvider playback (see the module docstring)."""

    def __init__(self, toplevel_path):
        mav0 = os.path.join(toplevel_path, "mav0")
        ls, lf = _parse_image_folder(os.path.join(mav0, "cam0"))
        rs, rf = _parse_image_folder(os.path.join(mav0, "cam1"))
        if not (len(ls) == len(rs) == len(lf) == len(rf)):
            raise ValueError("Different number of left/right images and timestamps")
        self.stereo_data = []
        for i in range(len(ls)):
            if ls[i] != rs[i]:
                raise ValueError("Left/right timestamps don't match!")
            for p in (lf[i], rf[i]):
                if not os.path.exists(p):
                    raise FileNotFoundError(p)
            self.stereo_data.append(StereoDatasetItem(ls[i], lf[i], rf[i]))
        self._callbacks = []

    def __len__(self):
        return len(self.stereo_data)

    def RegisterStereoCallback(self, cb):
        """cb(timestamp_ns, left_image, right_image) with the images as decoded (gray or BGR)."""
        self._callbacks.append(cb)

    def Playback(self, speed=1.0, verbose=False, realtime=True):
        """DataProvider::Playback: items in time order; sleeps (dt / speed) between them when
        `realtime` (the reference always does); returns the number of pairs delivered."""
        if not speed > 0.01:
            raise ValueError("Cannot go slower than 1% speed")
        last = None
        for item in self.stereo_data:
            if realtime and last is not None:
                time.sleep(max(0.0, (item.timestamp - last) * 1e-9 / speed))
            last = item.timestamp
            left, right = imread(item.path_left), imread(item.path_right)
            if verbose:
                print("stereo %d" % item.timestamp)
            for cb in self._callbacks:
                cb(item.timestamp, left, right)
        return len(self.stereo_data)


def write_euroc_sequence(toplevel_path, pairs, t0_ns=1_000_000_000, dt_ns=50_000_000):
    """Writes [(left, right), ...] (uint8 gray or RGB arrays) as a minimal EuRoC tree (tools, tests)."""
    for cam in ("cam0", "cam1"):
        os.makedirs(os.path.join(toplevel_path, "mav0", cam, "data"), exist_ok=True)
    rows = []
    for i, (l, r) in enumerate(pairs):
        ts = t0_ns + i * dt_ns
        rows.append(ts)
        write_png(os.path.join(toplevel_path, "mav0", "cam0", "data", "%d.png" % ts), l)
        write_png(os.path.join(toplevel_path, "mav0", "cam1", "data", "%d.png" % ts), r)
    for cam in ("cam0", "cam1"):
        with open(os.path.join(toplevel_path, "mav0", cam, "data.csv"), "w") as f:
            f.write("#timestamp [ns],filename\n")
            for ts in rows:
                f.write("%d,%d.png\n" % (ts, ts))
    return rows
