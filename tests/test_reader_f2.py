"""the C++ capture-file reader (cross-file byte stream + per-PRT frame parser) against the
oracle restatement of read_continuous_file_stream.m and FrameDataRead_xzr.m.  Host code only: runs without a GPU."""
import os

import numpy as np
import pytest

from oracle import mcode, synth
from radar_signal_process_b200.reader import FrameReader


def _write_capture(tmp_path, blob, cuts):
    """Split the byte stream into files 1.000001.bin, 1.000002.bin, ... at the given offsets."""
    edges = [0] + list(cuts) + [len(blob)]
    for i, (a, b) in enumerate(zip(edges[:-1], edges[1:]), start=1):
        name = "1.00000%d.bin" % i if i < 10 else "1.0000%d.bin" % i
        (tmp_path / name).write_bytes(blob[a:b])


def _stream(n_frames, n_prt, n_range, ch, seed=0):
    rng = np.random.default_rng(seed)
    payloads = rng.integers(-32768, 32768, size=(n_frames * n_prt, n_range, ch, 2), dtype=np.int16)
    blob = b"".join(synth.frame_prt(payloads[i], frame_no=i // n_prt, prt_no=i % n_prt, channel_num=ch, servo=(37 * i) % 3600)
                    for i in range(n_frames * n_prt))
    return payloads, blob


@pytest.mark.parametrize("cuts_kind", ["mid_payload", "single_file", "many_small"])
def test_reader_matches_oracle_across_file_boundaries(tmp_path, cuts_kind):
    n_frames, n_prt, n_range, ch = 3, 5, 37, 16
    payloads, blob = _stream(n_frames, n_prt, n_range, ch)
    prt_bytes = len(blob) // (n_frames * n_prt)
    cuts = {"mid_payload": [prt_bytes * 2 + 300, prt_bytes * 7 + 64 + 5, prt_bytes * 11 + 10],
            "single_file": [],
            "many_small": list(range(2777, len(blob), 2777))}[cuts_kind]      # every read crosses at most one boundary
    _write_capture(tmp_path, blob, cuts)
    rd = FrameReader(tmp_path)
    st = mcode.ContinuousFileStream(str(tmp_path))
    for f in range(n_frames):
        raw, meta, nread, eos = rd.next_frame(n_prt, n_range, ch)
        sig, servo, frame_no, timer, cur, send = mcode.frame_data_read_ddc(st, n_prt, n_range, ch)
        assert nread == cur == n_prt and eos == send
        assert np.array_equal(raw, payloads[f * n_prt:(f + 1) * n_prt])
        assert np.array_equal(raw[..., 0] + 1j * raw[..., 1], sig)
        assert np.array_equal(meta["servo_angle"], servo) and np.array_equal(meta["frame_no"], frame_no)
        assert rd.state() == (st.index, st.pos)
    raw, meta, nread, eos = rd.next_frame(n_prt, n_range, ch)            # stream exhausted
    sig, servo, frame_no, timer, cur, send = mcode.frame_data_read_ddc(st, n_prt, n_range, ch)
    assert (nread, eos) == (cur, send) == (0, True)


def test_reader_files_smaller_than_one_read_end_the_stream_like_the_mcode(tmp_path):
    # read_continuous_file_stream.m:101-133 continues into ONE next file only: a payload spanning three files is
    # a short read -> "frame not completed".  Both implementations stop at the same place.
    n_prt, n_range, ch = 5, 37, 16
    payloads, blob = _stream(1, n_prt, n_range, ch)
    _write_capture(tmp_path, blob, list(range(997, len(blob), 997)))
    raw, meta, nread, eos = FrameReader(tmp_path).next_frame(n_prt, n_range, ch)
    sig, _, _, _, cur, send = mcode.frame_data_read_ddc(mcode.ContinuousFileStream(str(tmp_path)), n_prt, n_range, ch)
    assert (nread, eos) == (cur, send) == (0, True)


def test_reader_reproduces_the_exact_boundary_file_skip(tmp_path):
    """A read that ends exactly at a file's end increments the file index twice (…m:148 then :48): the next
    file is skipped.  The C++ reader does the same so a capture parses identically to the M-code."""
    n_prt, n_range, ch = 2, 16, 4
    payloads, blob = _stream(3, n_prt, n_range, ch, seed=3)
    prt_bytes = len(blob) // 6
    _write_capture(tmp_path, blob, [prt_bytes * 2, prt_bytes * 4])        # files end exactly on PRT (tail) boundaries
    rd = FrameReader(tmp_path)
    st = mcode.ContinuousFileStream(str(tmp_path))
    a = rd.next_frame(n_prt, n_range, ch)
    o = mcode.frame_data_read_ddc(st, n_prt, n_range, ch)
    assert a[2] == o[4] == n_prt and np.array_equal(a[0], payloads[0:2])
    a = rd.next_frame(n_prt, n_range, ch)                                 # file 2 is skipped: frame 2 comes from file 3
    o = mcode.frame_data_read_ddc(st, n_prt, n_range, ch)
    assert a[2] == o[4] == n_prt
    assert np.array_equal(a[0], payloads[4:6]) and np.array_equal(a[0][..., 0] + 1j * a[0][..., 1], o[0])
    assert rd.state() == (st.index, st.pos)


def test_reader_truncated_stream_and_geometry_mismatch(tmp_path):
    n_prt, n_range, ch = 4, 20, 16
    payloads, blob = _stream(1, n_prt, n_range, ch, seed=5)
    _write_capture(tmp_path, blob[: len(blob) - 100], [])                 # last PRT cut short
    rd = FrameReader(tmp_path)
    raw, meta, nread, eos = rd.next_frame(n_prt, n_range, ch)
    sig, _, _, _, cur, send = mcode.frame_data_read_ddc(mcode.ContinuousFileStream(str(tmp_path)), n_prt, n_range, ch)
    assert (nread, eos) == (cur, send) and eos and nread == n_prt - 1
    assert np.array_equal(raw[:nread], payloads[:nread])
    rd2 = FrameReader(tmp_path)
    _, _, nread, eos = rd2.next_frame(n_prt, n_range + 1, ch)             # configured geometry differs from the heads
    assert nread == 0 and eos
    with pytest.raises(Exception):
        FrameReader(tmp_path / "does_not_exist")


def test_reader_corrupt_head_does_not_allocate_terabytes(tmp_path):
    """A PRT head whose pulse_data_num is garbage (0xFFFFFFF0 samples) ends the stream gracefully -- like the short fread of
    FrameDataRead_xzr.m:122-127 -- instead of asking for a multi-terabyte buffer or letting bad_alloc cross the C ABI."""
    n_prt, n_range, ch = 2, 20, 16
    _, blob = _stream(1, n_prt, n_range, ch, seed=7)
    bad = bytearray(blob)
    bad[24:28] = (0xFFFFFFF0).to_bytes(4, "little")                      # h[6] = pulse_data_num of the first PRT head
    _write_capture(tmp_path, bytes(bad), [])
    rd = FrameReader(tmp_path)
    _, _, nread, eos = rd.next_frame(n_prt, n_range, ch)
    assert nread == 0 and eos


def test_reader_dbf24_frames_round_trip_and_type_check(tmp_path):
    """DBF-type captures (data_type 2): the reader hands back the padded 24-bit payloads the device decoder takes; asking for
    DDC frames on such a capture is refused."""
    n_frames, n_prt, n_range, ch = 2, 3, 29, 13
    rng = np.random.default_rng(5)
    sig, pad, osp = mcode.dbf24_payload_size(n_range, ch)
    ncol = ((ch * 6 + osp) // 3) // 2
    lanes = rng.integers(-(1 << 23) + 1, 1 << 23, size=(n_frames, n_prt, n_range, ncol)) + 1j * rng.integers(-(1 << 23) + 1, 1 << 23, size=(n_frames, n_prt, n_range, ncol))
    payloads = np.stack([synth.to_dbf24(lanes[f], ch) for f in range(n_frames)])                 # [frame][prt][bytes]
    blob = b"".join(synth.frame_prt_dbf24(payloads[f, p], n_range, frame_no=f, prt_no=p, channel_num=ch, servo=11 * p)
                    for f in range(n_frames) for p in range(n_prt))
    _write_capture(tmp_path, blob, [len(blob) // 3 + 7])
    rd = FrameReader(tmp_path)
    for f in range(n_frames):
        got, meta, nread, eos = rd.next_frame_dbf24(n_prt, n_range, ch)
        assert nread == n_prt and not eos
        assert np.array_equal(got, payloads[f])
        assert np.array_equal(meta["frame_no"], np.full(n_prt, f)) and np.array_equal(meta["servo_angle"], 11 * np.arange(n_prt))
        for p in range(n_prt):                                                                    # and they decode to the lanes
            assert np.array_equal(mcode.unpack_dbf24(got[p], n_range, ch), lanes[f, p])
    assert rd.next_frame_dbf24(n_prt, n_range, ch)[2:] == (0, True)
    with pytest.raises(Exception):
        FrameReader(tmp_path).next_frame(n_prt, n_range, ch)
