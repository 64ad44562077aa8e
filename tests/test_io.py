"""Input/output layer: HIS reader, DDBVF container, directory/angle helpers, option parser,
against the reference's own compiled I/O chain (oracle/_ref, RefIO), the numpy restatement (oracle/formats.py)
and golden files written by the reference (tests/golden/io_*.{his,ddbvf}, made by tests/golden/make_io_golden.py)."""
from __future__ import annotations

import os

import numpy as np
import pytest

import oracle
from oracle import formats
from paris_b200 import io as pio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def frames_for(number_type: int, n=3, h=5, w=7, seed=20261018):
    rng = np.random.default_rng(seed + number_type)
    if number_type in (2, 4, 32):
        hi = {2: 255, 4: 65535, 32: 2 ** 32 - 1}[number_type]
        return rng.integers(0, hi, size=(n, h, w), endpoint=True).astype(formats.HIS_TYPES[number_type])
    return (rng.standard_normal((n, h, w)) * 1e3).astype(formats.HIS_TYPES[number_type])


# ---- HIS ----------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("number_type", [2, 4, 32, 64, 128])
@pytest.mark.parametrize("image_header_size,ulx,uly", [(32, 0, 0), (0, 3, 9), (100, 1, 0)])
def test_his_reader_equals_reference(tmp_path, number_type, image_header_size, ulx, uly):
    src = frames_for(number_type)
    path = str(tmp_path / "a.his")
    formats.write_his(path, src, number_type, image_header_size, ulx, uly)
    mine = pio.his_read(path)
    assert mine.shape == src.shape and mine.dtype == np.float32
    assert np.array_equal(mine, src.astype(np.float32))          # src/his.cpp:98-99: plain conversion to float
    assert np.array_equal(mine, formats.read_his(path))
    assert pio.his_info(path) == (src.shape[2], src.shape[1], src.shape[0], number_type)
    if oracle.have_ref():
        assert np.array_equal(mine, formats.RefIO().his_load(path))


def test_his_invalid_files(tmp_path):
    src = frames_for(4)
    cases = {"magic": dict(file_type=0x7001), "header_size": dict(header_size=100)}
    for name, kw in cases.items():
        p = str(tmp_path / f"{name}.his")
        formats.write_his(p, src, 4, **kw)
        assert pio.his_read(p).shape[0] == 0 and pio.his_info(p) is None
        if oracle.have_ref():
            assert formats.RefIO().his_load(p).shape[0] == 0
    p = str(tmp_path / "type.his")
    formats.write_his(p, src, 5)                       # unsupported number_type (src/his.cpp:188-190)
    assert pio.his_read(p).shape[0] == 0
    if oracle.have_ref():
        assert formats.RefIO().his_load(p).shape[0] == 0
    short = str(tmp_path / "short.his")
    open(short, "wb").write(b"\x00\x70\x44")             # shorter than the header
    assert pio.his_read(short).shape[0] == 0
    with pytest.raises(pio.IoError):                    # src/his.cpp:107-111: std::system_error
        pio.his_read(str(tmp_path / "missing.his"))
    if oracle.have_ref():
        with pytest.raises(OSError):
            formats.RefIO().his_load(str(tmp_path / "missing.his"))


def test_his_truncated_file_keeps_complete_frames(tmp_path):
    src = frames_for(4, n=4)
    p = str(tmp_path / "cut.his")
    formats.write_his(p, src, 4)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:68 + 2 * (32 + src[0].nbytes) + 40])   # third frame cut short
    got = pio.his_read(p)
    assert got.shape[0] == 2 and np.array_equal(got, src[:2].astype(np.float32))


# ---- the scan index of the group driver (which file and frame is projection i; 16-bit frames as they are) -----------

def _write_scan(tmp_path, types=(4, 4, 4), frames=(3, 5, 2), h=6, w=8):
    d = tmp_path / "scan"
    d.mkdir()
    all_frames = []
    for k, (t, n) in enumerate(zip(types, frames)):
        src = frames_for(t, n=n, h=h, w=w, seed=77 + k)
        formats.write_his(str(d / f"p_{k:02d}.his"), src, t, image_header_size=(0, 32, 100)[k % 3])
        all_frames += [src[i] for i in range(n)]
    (d / "p_00x.his").write_bytes(b"garbage")              # invalid files are skipped (src/source.cpp:97)
    return str(d), all_frames


@pytest.mark.parametrize("quality", [1, 2, 3])
def test_scan_index_numbers_frames_like_the_source(tmp_path, quality):
    d, frames = _write_scan(tmp_path)
    af = tmp_path / "angles.txt"
    af.write_text("\n".join(f"{1.5 * i}" for i in range(7)) + "\n")         # shorter than the scan (10 frames)
    idx, phi, from_file, (dim_x, dim_y, is_u16) = pio.scan_index(d, str(af), quality)
    kept = [i for i in range(len(frames)) if i % quality == 0]               # src/source.cpp:105
    assert list(idx) == kept and (dim_x, dim_y, is_u16) == (8, 6, True)
    assert [bool(f) for f in from_file] == [i < 7 for i in kept]
    assert all(phi[k] == np.float32(1.5 * i) for k, i in enumerate(kept) if i < 7)
    for k, i in enumerate(kept):
        f32, u16 = pio.scan_frame(d, quality, k, dim_x, dim_y)
        assert np.array_equal(u16, frames[i]) and u16.dtype == np.uint16      # the file's own samples
        assert np.array_equal(f32, frames[i].astype(np.float32))              # src/his.cpp:98-99
    assert pio.scan_frame(d, quality, len(kept), dim_x, dim_y) == (None, None)


def test_scan_of_mixed_sample_types_is_not_taken_as_sixteen_bit(tmp_path):
    d, frames = _write_scan(tmp_path, types=(4, 128, 4))
    _, _, _, (dim_x, dim_y, is_u16) = pio.scan_index(d, None, 1)
    assert not is_u16
    f32, u16 = pio.scan_frame(d, 1, 0, dim_x, dim_y)                          # a frame of a 16-bit file
    assert np.array_equal(u16, frames[0]) and np.array_equal(f32, frames[0].astype(np.float32))
    f32, u16 = pio.scan_frame(d, 1, 4, dim_x, dim_y)                          # a frame of the float file
    assert u16 is None and np.array_equal(f32, frames[4])
    # every-other-frame scans whose kept frames all sit in 16-bit files ARE 16-bit scans
    d2, _ = _write_scan(tmp_path / "b", types=(4, 128, 4), frames=(2, 1, 2)) if (tmp_path / "b").mkdir() is None else None
    idx, _, _, (_, _, is_u16) = pio.scan_index(d2, None, 4)
    assert list(idx) == [0, 4] and is_u16


def test_his_golden_written_for_the_reference_reader():
    p = os.path.join(GOLDEN, "io_u16.his")
    want = np.load(os.path.join(GOLDEN, "io_u16_frames.npy"))      # what the reference's his::load returned
    assert np.array_equal(pio.his_read(p), want)
    assert np.array_equal(formats.read_his(p), want)


# ---- DDBVF ----------------------------------------------------------------------------------------------------

def test_ddbvf_header_known_answer(tmp_path):
    d = pio.Ddbvf.create(str(tmp_path / "v"), 3, 4, 5)
    d.close()
    raw = open(str(tmp_path / "v.ddbvf"), "rb").read()
    assert raw == formats.ddbvf_header(3, 4, 5)                   # 32 bytes, nothing else yet
    assert raw[:4] == bytes([0xFA, 0xDA, 0xDD, 0xEF]) and raw[4:8] == bytes([0x10, 0, 0, 0])


@needs_ref
def test_ddbvf_file_is_byte_identical_to_the_reference(tmp_path):
    rng = np.random.default_rng(7)
    vol = rng.standard_normal((6, 4, 5)).astype(np.float32)
    formats.RefIO().ddbvf_create_write(str(tmp_path / "ref"), (5, 4, 6), vol, 0)
    d = pio.Ddbvf.create(str(tmp_path / "mine"), 5, 4, 6)
    d.write(vol, 0)
    d.close()
    a, b = open(str(tmp_path / "ref.ddbvf"), "rb").read(), open(str(tmp_path / "mine.ddbvf"), "rb").read()
    assert a == b
    assert np.array_equal(formats.read_ddbvf(str(tmp_path / "mine.ddbvf")), vol)


@needs_ref
def test_ddbvf_slab_at_offset_equals_reference(tmp_path):
    rng = np.random.default_rng(8)
    slab = rng.standard_normal((2, 3, 4)).astype(np.float32)
    formats.RefIO().ddbvf_create_write(str(tmp_path / "ref"), (4, 3, 7), slab, 3)     # slices [3, 5) of 7
    d = pio.Ddbvf.create(str(tmp_path / "mine"), 4, 3, 7)
    d.write(slab, 3)
    d.close()
    a, b = open(str(tmp_path / "ref.ddbvf"), "rb").read(), open(str(tmp_path / "mine.ddbvf"), "rb").read()
    assert a == b and len(a) == 32 + 4 * 3 * 5 * 4                                     # sparse tail not written


def test_ddbvf_slabs_reassemble_and_reopen(tmp_path):
    rng = np.random.default_rng(9)
    vol = rng.standard_normal((7, 3, 4)).astype(np.float32)
    d = pio.Ddbvf.create(str(tmp_path / "v"), 4, 3, 7)
    for first, count in ((4, 3), (0, 2), (2, 2)):          # any order (devices finish when they finish)
        d.write(vol[first:first + count], first)
    d.close()
    assert np.array_equal(formats.read_ddbvf(str(tmp_path / "v.ddbvf")), vol)
    r = pio.Ddbvf.open(str(tmp_path / "v.ddbvf"))         # (the reference's open() mis-reads its own header, F9)
    assert r.dims == (4, 3, 7)
    assert np.array_equal(r.read(2, 4), vol[2:6])
    r.close()


def test_ddbvf_errors(tmp_path):
    d = pio.Ddbvf.create(str(tmp_path / "v"), 4, 3, 5)
    ok = np.zeros((2, 3, 4), np.float32)
    with pytest.raises(pio.IoError, match="out of bounds"):          # src/ddbvf.cpp:128-129
        d.write(ok, 5)
    with pytest.raises(pio.IoError, match="wrong dimensions"):       # src/ddbvf.cpp:131-132
        d.write(np.zeros((2, 3, 5), np.float32), 0)
    with pytest.raises(pio.IoError, match="wrong dimensions"):
        d.write(np.zeros((6, 3, 4), np.float32), 0)
    with pytest.raises(pio.IoError, match="wrong dimensions"):       # would run past the end of the file's volume
        d.write(ok, 4)
    d.close()
    bad = str(tmp_path / "bad.ddbvf")
    open(bad, "wb").write(b"\x00" * 64)
    with pytest.raises(pio.IoError, match="Not a ddbvf file"):
        pio.Ddbvf.open(bad)


def test_ddbvf_golden_written_by_the_reference():
    raw = open(os.path.join(GOLDEN, "io_ref.ddbvf"), "rb").read()
    want = np.load(os.path.join(GOLDEN, "io_ref_volume.npy"))
    assert np.array_equal(formats.read_ddbvf(os.path.join(GOLDEN, "io_ref.ddbvf")), want)
    r = pio.Ddbvf.open(os.path.join(GOLDEN, "io_ref.ddbvf"))
    assert r.dims == (want.shape[2], want.shape[1], want.shape[0])
    assert np.array_equal(r.read(0, want.shape[0]), want)
    r.close()
    assert raw[:32] == formats.ddbvf_header(*r.dims)


# ---- directory, angles -------------------------------------------------------------------------------------------

def test_read_directory_sorted_canonical(tmp_path):
    d = tmp_path / "scan"
    d.mkdir()
    for name in ("b_010.his", "a_002.his", "a_001.his", "c.his"):
        (d / name).write_bytes(b"x")
    got = pio.read_directory(str(d))
    assert got == sorted(os.path.realpath(str(d / n)) for n in os.listdir(d))
    if oracle.have_ref():
        assert got == formats.RefIO().read_directory(str(d))
    with pytest.raises(pio.IoError, match="is not a directory"):
        pio.read_directory(str(d / "c.his"))
    with pytest.raises(pio.IoError, match="does not exist"):
        pio.read_directory(str(d / "nope"))
    assert pio.create_directory(str(tmp_path / "x" / "y" / "z")) and os.path.isdir(tmp_path / "x" / "y" / "z")
    assert pio.create_directory(str(tmp_path / "x"))
    with pytest.raises(pio.IoError, match="not a directory"):
        pio.create_directory(str(d / "c.his"))


def test_read_angles(tmp_path):
    p = tmp_path / "angles.txt"
    p.write_text("0.0 0.5 1.25\n2.5\n359.75\n")
    assert np.array_equal(pio.read_angles(str(p)), np.float32([0.0, 0.5, 1.25, 2.5, 359.75]))
    p.write_text("0,0\n0,5\n1,25\n")                      # decimal comma (src/source.cpp:57-62)
    assert np.array_equal(pio.read_angles(str(p)), np.float32([0.0, 0.5, 1.25]))
    assert pio.read_angles(str(tmp_path / "missing.txt")).size == 0     # warning, default angles


# ---- command line ----------------------------------------------------------------------------------------------

GEOMETRY = """# detector
n_row = 256
n_col=128
l_px_row = 0.4
l_px_col = 0.5   # mm
delta_s = 1.5
delta_t = -2
d_so = 500
d_od = 400.5
delta_phi = 0.25
"""


def test_options_defaults_and_geometry_file(tmp_path):
    g = tmp_path / "geo.cfg"
    g.write_text(GEOMETRY)
    rc, o, msg = pio.parse_options(["--geometry", str(g)])
    assert rc == 0, msg
    assert (o.det.n_row, o.det.n_col) == (256, 128)
    assert np.float32(o.det.l_px_col) == np.float32(0.5) and np.float32(o.det.delta_t) == np.float32(-2.0)
    assert np.float32(o.det.d_od) == np.float32(400.5) and np.float32(o.det.delta_phi) == np.float32(0.25)
    assert (o.enable_io, o.enable_roi, o.enable_angles, o.quality) == (0, 0, 0, 1)
    assert o.prefix == b"vol"                                                   # src/program_options.cpp:74


def test_options_full_command_line(tmp_path):
    g = tmp_path / "geo.cfg"
    g.write_text(GEOMETRY)
    rc, o, msg = pio.parse_options([f"--geometry={g}", "--input", "/in", "--output=/out", "--name", "scan7",
                                    "--angles", "/a.txt", "--quality", "3", "--roi", "--roi-x1", "1", "--roi-x2", "20",
                                    "--roi-y1=2", "--roi-y2=30", "--roi-z1", "3", "--roi-z2", "40"])
    assert rc == 0, msg
    assert (o.enable_io, o.enable_roi, o.enable_angles, o.quality) == (1, 1, 1, 3)
    assert (o.input_path, o.output_path, o.prefix, o.angle_path) == (b"/in", b"/out", b"scan7", b"/a.txt")
    assert (o.roi.x1, o.roi.x2, o.roi.y1, o.roi.y2, o.roi.z1, o.roi.z2) == (1, 20, 2, 30, 3, 40)


@pytest.mark.parametrize("args,needle", [
    ([], "the option '--geometry' is required but missing"),
    (["--geometry", "G", "--input", "/in"], "the option '--output' is required but missing"),
    (["--geometry", "G", "--output", "/o"], "the option '--input' is required but missing"),
    (["--geometry", "G", "--roi", "--roi-x1", "1"], "the option '--roi-x2' is required but missing"),
    (["--geometry", "G", "--frobnicate"], "unrecognised option '--frobnicate'"),
    (["--geometry", "G", "--quality", "many"], "'--quality' is invalid"),
    (["--geometry"], "the required argument for option '--geometry' is missing"),
    (["--geometry", "/nonexistent/geo.cfg"], "the option '--n_row' is required but missing"),
])
def test_options_errors(tmp_path, args, needle):
    g = tmp_path / "geo.cfg"
    g.write_text(GEOMETRY)
    rc, _, msg = pio.parse_options([str(g) if a == "G" else a for a in args])
    assert rc == 2 and needle in msg


def test_options_geometry_file_errors(tmp_path):
    g = tmp_path / "geo.cfg"
    g.write_text(GEOMETRY.replace("d_so = 500\n", ""))
    rc, _, msg = pio.parse_options(["--geometry", str(g)])
    assert rc == 2 and "the option '--d_so' is required but missing" in msg
    g.write_text(GEOMETRY + "colour = blue\n")
    rc, _, msg = pio.parse_options(["--geometry", str(g)])
    assert rc == 2 and "unrecognised option 'colour'" in msg


def test_options_help():
    rc, _, msg = pio.parse_options(["--help"])
    assert rc == 1 and "--geometry" in msg and "--roi-z2" in msg and "--quality" in msg
    rc, _, msg = pio.parse_options(["--geometry-format"])
    assert rc == 1 and "delta_phi" in msg and "n_row" in msg


@needs_ref
@pytest.mark.parametrize("num,dim_z,rem", [(1, 10, 0), (4, 7, 3), (8, 128, 0)])
def test_reference_make_tasks_shape(num, dim_z, rem):
    ids, dz = formats.RefIO().make_tasks(num, dim_z, rem)        # src/task.cpp:38-48: one task per slab, same geometry
    assert list(ids) == list(range(num)) and set(dz) == {dim_z}
