"""The host front end's OpenEXR reader (csrc/host/exr.cpp), which stands in for the OpenEXR library behind ReadImageEXR
(core/imageio.cpp:146-167).  The reference build this repo pins against has no OpenEXR (oracle/Makefile), so the decoder is
checked against an independent numpy reader and writer of the same file layout below -- scan-line files, HALF colour
channels, no / RLE / ZIPS / ZIP compression -- and, where the reference checkout is present, on its own environment maps."""
import os
import struct
import zlib

import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa

LINES = {0: 1, 1: 1, 2: 1, 3: 16}


def _attr(name, typ, value):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(value)) + value


def _rle(data):
    out, i, n = bytearray(), 0, len(data)
    while i < n:
        j = i
        while j < n and j - i < 127 and data[j] == data[i]:
            j += 1
        if j - i >= 3:
            out += struct.pack("b", j - i - 1) + data[i:i + 1]
            i = j
        else:
            j = i
            while j < n and j - i < 127 and not (j + 2 < n and data[j] == data[j + 1] == data[j + 2]):
                j += 1
            out += struct.pack("b", -(j - i)) + data[i:j]
            i = j
    return bytes(out)


def write_exr(path, channels, compression=3, origin=(0, 0), decreasing=False):
    """channels: {name: (type, array[h, w])}, type 1 = HALF (float16 array), 2 = FLOAT (float32), 0 = UINT (uint32)"""
    names = sorted(channels)
    h, w = channels[names[0]][1].shape
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", channels[n][0], 0, 1, 1) for n in names) + b"\0"
    x0, y0 = origin
    box = struct.pack("<4i", x0, y0, x0 + w - 1, y0 + h - 1)
    head = struct.pack("<II", 20000630, 2) + _attr("channels", "chlist", chlist) + _attr("compression", "compression", bytes([compression]))
    head += _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box)
    head += _attr("lineOrder", "lineOrder", bytes([1 if decreasing else 0])) + _attr("pixelAspectRatio", "float", struct.pack("<f", 1))
    head += _attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0)) + _attr("screenWindowWidth", "float", struct.pack("<f", 1)) + b"\0"
    lpb = LINES[compression]
    chunks = []
    for b0 in range(0, h, lpb):
        raw = b"".join(channels[n][1][y].astype({0: "<u4", 1: "<f2", 2: "<f4"}[channels[n][0]]).tobytes()
                       for y in range(b0, min(h, b0 + lpb)) for n in names)
        data = raw
        if compression:
            a = np.frombuffer(raw, np.uint8)
            t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)
            t[1:] = (t[1:] - t[:-1] + 128 + 256) % 256
            t = t.astype(np.uint8).tobytes()
            comp = _rle(t) if compression == 1 else zlib.compress(t)
            if len(comp) < len(raw):
                data = comp
        chunks.append(struct.pack("<ii", y0 + b0, len(data)) + data)
    order = list(range(len(chunks)))
    if decreasing:
        order.reverse()
    pos, offsets = len(head) + 8 * len(chunks), [0] * len(chunks)
    body = b""
    for i in order:
        offsets[i] = pos + len(body)
        body += chunks[i]
    with open(path, "wb") as f:
        f.write(head + struct.pack("<%dQ" % len(chunks), *offsets) + body)


def read_exr(path):
    """independent reader: R, G, B of a scan-line file with HALF colour channels -> float32 [h, w, 3]"""
    d = open(path, "rb").read()
    assert struct.unpack_from("<I", d)[0] == 20000630
    p, attrs = 8, {}
    while d[p] != 0:
        e = d.index(b"\0", p); name = d[p:e].decode(); p = e + 1
        e = d.index(b"\0", p); p = e + 1
        n = struct.unpack_from("<i", d, p)[0]; p += 4
        attrs[name] = d[p:p + n]; p += n
    p += 1
    chans, q, cl = [], 0, attrs["channels"]
    while cl[q] != 0:
        e = cl.index(b"\0", q)
        chans.append((cl[q:e].decode(), struct.unpack_from("<i", cl, e + 1)[0]))
        q = e + 17
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    comp = attrs["compression"][0]
    lpb = LINES[comp]
    out = np.zeros((h, w, 3), np.float32)
    size = {0: 4, 1: 2, 2: 4}
    line = sum(size[t] * w for _, t in chans)
    for c in range((h + lpb - 1) // lpb):
        off = struct.unpack_from("<Q", d, p + 8 * c)[0]
        y, n = struct.unpack_from("<ii", d, off)
        raw = d[off + 8:off + 8 + n]
        lines = min(lpb, y1 - y + 1)
        if comp and n != line * lines:
            if comp == 1:
                t, i = bytearray(), 0
                while i < len(raw):
                    k = struct.unpack_from("b", raw, i)[0]; i += 1
                    if k < 0:
                        t += raw[i:i - k]; i -= k
                    else:
                        t += raw[i:i + 1] * (k + 1); i += 1
                t = bytes(t)
            else:
                t = zlib.decompress(raw)
            a = np.frombuffer(t, np.uint8).astype(np.int64)
            a = ((np.cumsum(a - 128) + 128) % 256).astype(np.uint8)
            half = (len(a) + 1) // 2
            r = np.empty(len(a), np.uint8)
            r[0::2], r[1::2] = a[:half], a[half:]
            raw = r.tobytes()
        for l in range(lines):
            at = l * line
            for name, t in chans:
                if name in "RGB" and len(name) == 1:
                    out[y - y0 + l, :, "RGB".index(name)] = np.frombuffer(raw, "<f2", w, at).astype(np.float32)
                at += size[t] * w
    return out


def _image(rs, h, w):
    a = np.exp(rs.randn(h, w, 3) * 2).astype(np.float16)
    a[0, 0] = [0, 6e-8, 65504]          # zero, a subnormal half, the largest half
    a[h // 2, :] = 0.25                 # a constant run (RLE / deflate actually compress)
    return a


@pytest.mark.parametrize("compression", [0, 1, 2, 3])
def test_exr_reader_against_independent_writer(tmp_path, compression):
    rs = np.random.RandomState(compression)
    for i, (h, w) in enumerate([(1, 1), (9, 17), (16, 8), (33, 5), (40, 64)]):
        img = _image(rs, h, w)
        ch = {"R": (1, img[..., 0]), "G": (1, img[..., 1]), "B": (1, img[..., 2])}
        if i % 2:   # channels that are not colour: an alpha half and a float depth, before and after the colours in the file
            ch["A"] = (1, np.ones((h, w), np.float16))
            ch["Z"] = (2, rs.rand(h, w).astype(np.float32))
        path = str(tmp_path / ("t%d.exr" % i))
        write_exr(path, ch, compression, origin=(3 * i, -2 * i), decreasing=bool(i & 2))
        got = pb.host_read_image(path)
        assert got.shape == (h, w, 3)
        assert np.array_equal(got.view(np.uint32), img.astype(np.float32).view(np.uint32))
        assert np.array_equal(read_exr(path).view(np.uint32), got.view(np.uint32))


def test_exr_features_outside_the_reader_are_refused(tmp_path, capfd):
    img = _image(np.random.RandomState(1), 4, 4)
    path = str(tmp_path / "f.exr")
    write_exr(path, {"R": (2, img[..., 0].astype(np.float32)), "G": (1, img[..., 1]), "B": (1, img[..., 2])})
    with pytest.raises(pb.PbrtGpuError):
        pb.host_read_image(path)            # a FLOAT colour channel (RgbaInputFile would round it to half)
    write_exr(path, {"Y": (1, img[..., 0])})
    with pytest.raises(pb.PbrtGpuError):
        pb.host_read_image(path)            # luminance only
    raw = bytearray(open(path, "rb").read())
    raw[raw.index(b"compression\0compression\0") + 28] = 4
    open(path, "wb").write(bytes(raw))
    with pytest.raises(pb.PbrtGpuError):
        pb.host_read_image(path)            # PIZ
    open(path, "wb").write(bytes(raw[:60]))
    with pytest.raises(pb.PbrtGpuError):
        pb.host_read_image(path)            # truncated
    assert "not supported" in capfd.readouterr().err


def test_exr_environment_map_is_the_pfm_environment_map(staged, tmp_path):
    """the same texels through either file format: the same light (arrays) and the same film"""
    from pbrt_v2_b200 import scenes
    rs = np.random.RandomState(5)
    img = _image(rs, 7, 12)
    write_exr(os.path.join(scenes.STAGED, "t-env.exr"), {"R": (1, img[..., 0]), "G": (1, img[..., 1]), "B": (1, img[..., 2])})
    with open(os.path.join(scenes.STAGED, "t-env.pfm"), "wb") as f:
        f.write(b"PF\n12 7\n-1\n" + img.astype("<f4").tobytes())
    a = pb.host_envmap_build(os.path.join(scenes.STAGED, "t-env.exr"), [1, 1, 1])
    b = pb.host_envmap_build(os.path.join(scenes.STAGED, "t-env.pfm"), [1, 1, 1])
    assert all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in a)
    films = []
    for ext in ("exr", "pfm"):
        path = scenes.variant("bunny", os.path.join(scenes.STAGED, "t-env-%s.pbrt" % ext), 24, 24, 4)
        text = open(path).read().replace("WorldBegin", 'WorldBegin\nLightSource "infinite" "string mapname" "t-env.%s"\n' % ext)
        open(path, "w").write(text)
        films.append(oa.render(pb.HostScene(path, ncores=2), threads=4)[0])
    assert np.array_equal(films[0].view(np.uint8), films[1].view(np.uint8))


@pytest.mark.skipif(not os.path.isdir("/root/reference/scenes/textures"), reason="no reference checkout here")
@pytest.mark.parametrize("name", ["grace_latlong.exr", "lines.exr", "uffizi_latlong.exr"])
def test_exr_reader_on_the_reference_distribution_s_maps(name):
    path = os.path.join("/root/reference/scenes/textures", name)
    got = pb.host_read_image(path)
    assert np.array_equal(read_exr(path).view(np.uint32), got.view(np.uint32))
    assert np.isfinite(got).all() and got.max() > 0.5


def test_exr_film_writer_round_trips(staged, tmp_path):
    """ImageFilm::WriteImage to an .exr name (WriteImageEXR, core/imageio.cpp:170-193: HALF RGBA over the crop region's data
    window inside the full display window): the file decodes -- by the independent reader and by the front end's own -- to
    the film's RGB rounded to half the way the OpenEXR library's `half` rounds (nearest even = numpy's float16 cast)."""
    from helpers import variant_scene
    hs, _ = variant_scene("nonsquare-crop")
    px, _ = oa.render(hs, threads=4)
    want = hs.pixels_to_rgb(px)
    # values that exercise the rounding: ties, subnormal halfs, overflow to infinity, zeros
    flat = want.reshape(-1)
    special = np.array([0.0, 1.0, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 65504.0, 65519.9, 65520.0, 1e9, 6e-8, 2.98e-8, 2.99e-8,
                        5.96e-8, 6.1e-5, 6.09e-5, 0.1, 1 / 3, 2049.0, 2051.0], np.float32)
    assert len(flat) >= len(special)
    path = str(tmp_path / "film.exr")
    hs.write_image(px, path)
    got = pb.host_read_image(path)
    assert got.shape == want.shape
    with np.errstate(over="ignore"):
        assert np.array_equal(got.view(np.uint32), want.astype(np.float16).astype(np.float32).view(np.uint32))
    assert np.array_equal(read_exr(path).view(np.uint32), got.view(np.uint32))
    d = open(path, "rb").read()
    xs, ys, xc, yc = hs.film_extent()
    at = d.index(b"dataWindow\0box2i\0") + 21
    assert struct.unpack_from("<4i", d, at) == (xs, ys, xs + xc - 1, ys + yc - 1)
    at = d.index(b"displayWindow\0box2i\0") + 24
    rd = hs.render_desc()
    assert struct.unpack_from("<4i", d, at) == (0, 0, rd.xResolution - 1, rd.yResolution - 1)
    # the rounding itself, on the special values, through a film of them
    px2 = hs.new_pixels()
    px2["weightSum"] = 1
    rgb = np.resize(special, px2["Lxyz"].shape).astype(np.float32)
    # XYZ whose RGB is the wanted value: write_image converts XYZ -> RGB, so feed it grey XYZ = v * (white point) instead
    # (simpler: compare against the front end's own conversion of the same pixels)
    px2["Lxyz"] = rgb
    hs.write_image(px2, path)
    want2 = hs.pixels_to_rgb(px2)
    with np.errstate(over="ignore"):
        assert np.array_equal(pb.host_read_image(path).view(np.uint32), want2.astype(np.float16).astype(np.float32).view(np.uint32))
    # ... and over the whole half range: magnitudes from below the smallest subnormal half to beyond the largest half
    rs = np.random.RandomState(3)
    for _ in range(6):
        px2["Lxyz"] = (np.exp(rs.uniform(-22, 13, px2["Lxyz"].shape)) * rs.choice([1, 1, 1, 0], px2["Lxyz"].shape)).astype(np.float32)
        hs.write_image(px2, path)
        want2 = hs.pixels_to_rgb(px2)
        with np.errstate(over="ignore"):
            assert np.array_equal(pb.host_read_image(path).view(np.uint32), want2.astype(np.float16).astype(np.float32).view(np.uint32))


def test_exr_writer_rounds_every_half_neighbourhood_like_the_library(tmp_path):
    """float -> half: every finite half, every tie between two neighbouring halfs (exactly representable as a float) and the
    floats just below and above each tie, with both signs; overflow and non-finite values: numpy's float16 cast is the same
    IEEE round-to-nearest-even the OpenEXR library's `half` implements"""
    h = np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float32)
    mid = ((h[:-1].astype(np.float64) + h[1:].astype(np.float64)) / 2).astype(np.float32)
    vals = np.concatenate([h, mid, np.nextafter(mid, np.float32(0)), np.nextafter(mid, np.float32(1e9)),
                           np.array([65519.996, 65520.0, 70000.0, 1e30, np.inf], np.float32)])
    vals = np.concatenate([vals, -vals, np.zeros((-2 * len(vals)) % 3, np.float32)]).astype(np.float32)
    img = vals.reshape(1, -1, 3)
    path = str(tmp_path / "r.exr")
    pb.host_write_exr(path, img)
    with np.errstate(over="ignore"):
        want = img.astype(np.float16).astype(np.float32)
    assert np.array_equal(pb.host_read_image(path).view(np.uint32), want.view(np.uint32))
    assert np.array_equal(read_exr(path).view(np.uint32), want.view(np.uint32))
