"""Host-side ingest for the sequence driver (ocean-perception_b200/dataset.py): the EuRoC layout of the
reference's dataset::EurocDataset (euroc_dataset.cpp:119-166), PNG decoding, MaybeConvertToGray and the
half-size resize of PatchmatchGpuTest.Sequence (patchmatch_gpu_test.cpp:124-129). CPU only."""
import importlib
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ds():
    return importlib.import_module("ocean-perception_b200.dataset")


def _cv2():
    try:
        import cv2
        return cv2
    except ImportError:
        return None


def test_png_roundtrip_and_decoder_against_cv2(ds, tmp_path):
    rng = np.random.default_rng(0)
    gray = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    rgb = rng.integers(0, 256, (21, 40, 3)).astype(np.uint8)
    ds.write_png(str(tmp_path / "g.png"), gray)
    ds.write_png(str(tmp_path / "c.png"), rgb)
    assert np.array_equal(ds.read_png(str(tmp_path / "g.png")), gray)
    assert np.array_equal(ds.read_png(str(tmp_path / "c.png")), rgb)
    cv2 = _cv2()
    if cv2 is None:
        pytest.skip("cv2 not importable: decoder checked against its own writer only")
    # files written by libpng (all five filter types on smooth content), decoded by read_png
    yy, xx = np.mgrid[0:64, 0:96]
    smooth = ((np.sin(xx / 7.0) + np.cos(yy / 5.0)) * 60 + 128).astype(np.uint8)
    bgr = np.stack([smooth, smooth[::-1], np.roll(smooth, 9, 1)], -1)
    cv2.imwrite(str(tmp_path / "s.png"), smooth)
    cv2.imwrite(str(tmp_path / "b.png"), bgr)
    assert np.array_equal(ds.read_png(str(tmp_path / "s.png")), smooth)
    assert np.array_equal(ds.read_png(str(tmp_path / "b.png"))[:, :, ::-1], bgr)     # PNG stores R, G, B
    # MaybeConvertToGray == cv::cvtColor(BGR2GRAY); resize_half == cv::resize(size / 2)
    assert np.array_equal(ds.maybe_convert_to_gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(ds.resize_half(smooth), cv2.resize(smooth, (48, 32)))


def test_euroc_layout_playback(ds, tmp_path):
    rng = np.random.default_rng(1)
    pairs = [(rng.integers(0, 256, (16, 24)).astype(np.uint8), rng.integers(0, 256, (16, 24)).astype(np.uint8))
             for _ in range(4)]
    stamps = ds.write_euroc_sequence(str(tmp_path / "seq"), pairs, dt_ns=1_000_000)
    d = ds.EurocDataset(str(tmp_path / "seq"))
    assert len(d) == 4 and [it.timestamp for it in d.stereo_data] == stamps
    seen = []
    d.RegisterStereoCallback(lambda ts, l, r: seen.append((ts, l.copy(), r.copy())))
    assert d.Playback(speed=100.0) == 4
    for (ts, l, r), want_ts, (wl, wr) in zip(seen, stamps, pairs):
        assert ts == want_ts and np.array_equal(l, wl) and np.array_equal(r, wr)
    with pytest.raises(ValueError):
        d.Playback(speed=0.0)          # CHECK_GT(speed, 0.01f), data_provider.cpp:183


def test_euroc_layout_errors(ds, tmp_path):
    rng = np.random.default_rng(2)
    pairs = [(rng.integers(0, 256, (8, 8)).astype(np.uint8),) * 2 for _ in range(2)]
    top = str(tmp_path / "seq")
    ds.write_euroc_sequence(top, pairs)
    csv = os.path.join(top, "mav0", "cam1", "data.csv")
    lines = open(csv).read().splitlines()
    open(csv, "w").write("\n".join(lines[:-1]) + "\n")
    with pytest.raises(ValueError):       # "Different number of left/right images and timestamps"
        ds.EurocDataset(top)
    open(csv, "w").write("\n".join(lines[:-1] + ["123," + "123.png"]) + "\n")
    with pytest.raises(ValueError):       # "Left/right timestamps don't match!"
        ds.EurocDataset(top)
    with pytest.raises(FileNotFoundError):
        ds.EurocDataset(str(tmp_path / "nothing"))
