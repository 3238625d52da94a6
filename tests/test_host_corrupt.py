"""The reference's corrupt-file cases (tests/corrupt_cases.py) on the CPU
emulation backend; the same table runs on the GPU in test_gpu_corrupt.py."""
import pytest

import cpu_backend
import corrupt_cases as cc


@pytest.fixture(autouse=True)
def backend(monkeypatch):
    cpu_backend.install(monkeypatch)


def _id(v):
    return str(v).replace(' ', '')


@pytest.mark.parametrize('missing', cc.MISSING_FRAMES, ids=_id)
def test_sample_copy_missing_frames(missing):
    cc.sample_copy_missing_frames(missing)


@pytest.mark.parametrize('missing', cc.MISSING_BYTES, ids=_id)
def test_sample_copy_missing_bytes(missing):
    cc.sample_copy_missing_bytes(missing)


@pytest.mark.parametrize('frame_nr', cc.MISSING_FRAMESET, ids=_id)
def test_missing_frameset(frame_nr):
    cc.fake_missing_frameset(frame_nr)


@pytest.mark.parametrize('frame_nr,thread', cc.MISSING_THREAD)
def test_missing_thread(frame_nr, thread):
    cc.fake_missing_thread(frame_nr, thread)


@pytest.mark.parametrize('missing_bytes', cc.MISSING_END, ids=_id)
def test_missing_end(missing_bytes):
    cc.fake_missing_end(missing_bytes)


@pytest.mark.parametrize('missing_bytes,missing_data,kept', cc.MISSING_MIDDLE,
                         ids=_id)
def test_missing_middle(missing_bytes, missing_data, kept):
    cc.fake_missing_middle(missing_bytes, missing_data, kept)


def test_invalid_frame_headers():
    cc.fake_invalid_frame_headers()


@pytest.mark.parametrize('affected,replacement', cc.M5B_BAD_BYTES, ids=_id)
def test_m5b_sample_bad_bytes(affected, replacement):
    cc.m5b_sample_bad_bytes(affected, replacement)


@pytest.mark.parametrize('frame_nr', cc.MISSING_FRAMESET, ids=_id)
def test_m5b_missing_frames(frame_nr):
    cc.m5b_fake_missing_frames(frame_nr)


@pytest.mark.parametrize('missing_bytes,missing_frames',
                         cc.M5B_MISSING_MIDDLE, ids=_id)
def test_m5b_missing_middle(missing_bytes, missing_frames):
    cc.m5b_fake_missing_middle(missing_bytes, missing_frames)


@pytest.mark.parametrize('frame_nr', cc.M4_MISSING_FRAMES, ids=_id)
def test_m4_missing_frames(frame_nr):
    cc.m4_fake_missing_frames(frame_nr)
