"""CPU: the BMP / QOI / PNG readers and writers of the batch command line (imcvt_b200/host/fmt_{bmp,qoi,png}.c, SURVEY 8(f)
row f4) against the reference's CLI compiled unmodified (oracle/_ref/ImCvt_ref: src/imageio_{bmp,qoi,png}.c + src/uPNG) --
stdout, exit code and every written file byte for byte, on generated pictures of every kind the reference's loaders
distinguish (PNG colour type x bit depth x filters x deflate block kinds, interlaced / palette refusals, truncated files;
BMP 8 / 24 / 32 bit with grey and coloured palettes, odd widths; QOI with every op, 3 and 4 channels).  Without the compiled
reference the files are checked through round trips only."""
import os
import struct
import subprocess
import tempfile
import zlib

import numpy as np
import pytest

from tests._util import GOLDEN, ROOT, load_pnm

CLI = os.path.join(ROOT, "imcvt_b200", "ImCvt_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ImCvt_ref")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        import __graft_entry__
        __graft_entry__.build()


def run(exe, args, cwd):
    p = subprocess.run([exe] + args, cwd=cwd, capture_output=True, timeout=600)
    return p.returncode, p.stdout


# ----------------------------------------------------------------------------------------------------------- generators
def chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))


def png(h, w, ctype, depth, rng, *, filters=None, level=6, interlace=0, extra=b"", split_idat=1, wbits=15, strategy=0):
    """A PNG with random samples: colour type / depth as given, a filter per line (really applied), IDAT split in pieces."""
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpl = (w * ch * depth + 7) // 8
    bpp = max(1, ch * depth // 8)
    rows = rng.integers(0, 256, (h, bpl), dtype=np.uint8)
    if depth < 8:                                            # smooth-ish content so that the compressor finds matches
        rows[:, ::2] = rows[:, :1]
    raw = bytearray()
    prev = np.zeros(bpl, np.int32)
    for y in range(h):
        cur = rows[y].astype(np.int32)
        f = (filters[y % len(filters)] if filters else 0)
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if bpl > bpp else np.zeros(bpl, np.int32)
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if bpl > bpp else np.zeros(bpl, np.int32)
        if f == 0:
            out = cur
        elif f == 1:
            out = cur - left
        elif f == 2:
            out = cur - prev
        elif f == 3:
            out = cur - (left + prev) // 2
        else:
            p = left + prev - ul
            pa, pb, pc = abs(p - left), abs(p - prev), abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            out = cur - pred
        raw.append(f)
        raw += (out & 255).astype(np.uint8).tobytes()
        prev = cur
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    z = co.compress(bytes(raw)) + co.flush()
    body = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if ctype == 3:
        body += chunk(b"PLTE", bytes(range(48)))
    body += extra
    n = max(1, len(z) // split_idat)
    for i in range(0, len(z), n):
        body += chunk(b"IDAT", z[i:i + n])
    return body + chunk(b"IEND", b"")


def bmp(h, w, bits, rng, *, palette="grey", ncolors=256, offset_pad=0, top_down=False, hdr=40):
    bpp = bits // 8
    row = (w * bpp + 3) // 4 * 4
    pal = b""
    if bits == 8:
        if palette == "grey":
            pal = b"".join(bytes((i, i, i, 0)) for i in range(ncolors))
        else:
            pal = rng.integers(0, 256, 4 * ncolors, dtype=np.uint8).tobytes()
    off = 14 + hdr + len(pal) + offset_pad
    pix = rng.integers(0, 256, h * row, dtype=np.uint8).tobytes()
    head = b"BM" + struct.pack("<IHHI", off + len(pix), 0, 0, off)
    info = struct.pack("<IiiHHIIiiII", hdr, w, -h if top_down else h, 1, bits, 0, len(pix), 2835, 2835, ncolors if bits == 8 else 0, 0)
    return head + info + bytes(hdr - 40) + pal + bytes(offset_pad) + pix


def qoi(h, w, channels, rng, n_ops=None, end_marker=True):
    """A QOI stream of random ops (INDEX / DIFF / LUMA / RUN / RGB / RGBA); it need not fill the picture exactly."""
    out = bytearray(b"qoif" + struct.pack(">IIBB", w, h, channels, 0))
    n = n_ops if n_ops is not None else h * w
    for _ in range(n):
        op = int(rng.integers(0, 6))
        if op == 0:
            out.append(int(rng.integers(0, 64)))
        elif op == 1:
            out.append(0x40 | int(rng.integers(0, 64)))
        elif op == 2:
            out += bytes((0x80 | int(rng.integers(0, 64)), int(rng.integers(0, 256))))
        elif op == 3:
            out.append(0xC0 | int(rng.integers(0, 62)))
        elif op == 4:
            out += bytes((0xFE,)) + rng.integers(0, 256, 3, dtype=np.uint8).tobytes()
        else:
            out += bytes((0xFF,)) + rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
    if end_marker:
        out += bytes(7) + b"\x01"
    return bytes(out)


def make_inputs(d, seed=20241):
    rng = np.random.default_rng(seed)
    names = []

    def put(name, data):
        open(os.path.join(d, name), "wb").write(data)
        names.append(name)

    # PNG: every colour type x depth of the specification (the reference accepts LUMA8 / RGB8 / RGBA8, names or numbers the rest)
    for ctype, depths in ((0, (1, 2, 4, 8, 16)), (2, (8, 16)), (3, (1, 2, 4, 8)), (4, (8, 16)), (6, (8, 16))):
        for depth in depths:
            put("t%d_%d.png" % (ctype, depth), png(13, 17, ctype, depth, rng, filters=[0, 1, 2, 3, 4]))
    for i, (h, w, ctype) in enumerate(((1, 1, 0), (1, 1, 2), (2, 300, 6), (97, 3, 2), (64, 64, 0), (200, 311, 2), (5, 1000, 6))):
        put("s%d.png" % i, png(h, w, ctype, 8, rng, filters=[4, 3, 1, 2, 0, 4, 4]))
    put("stored.png", png(40, 50, 2, 8, rng, level=0))                      # stored blocks
    put("fixed.png", png(40, 50, 0, 8, rng, strategy=zlib.Z_FIXED))          # fixed Huffman blocks
    put("rle.png", png(120, 90, 2, 8, rng, strategy=zlib.Z_RLE, filters=[1]))
    put("best.png", png(150, 150, 6, 8, rng, level=9, filters=[2, 4]))
    put("win9.png", png(80, 80, 2, 8, rng, wbits=9, filters=[4]))
    put("idat5.png", png(60, 70, 2, 8, rng, split_idat=5, filters=[3]))
    put("interlaced.png", png(16, 16, 2, 8, rng, interlace=1))
    put("anc.png", png(9, 9, 0, 8, rng, extra=chunk(b"tEXt", b"k\0v") + chunk(b"gAMA", struct.pack(">I", 45455))))
    put("crit.png", png(9, 9, 0, 8, rng, extra=chunk(b"ABCD", b"1234")))
    good = png(30, 30, 2, 8, rng, filters=[4])
    put("trunc_a.png", good[:20])
    put("trunc_b.png", good[:60])
    put("trunc_c.png", good[:len(good) - 40])
    put("nosig.png", b"\x89PNX" + good[4:])
    put("empty.png", b"")
    flat = np.zeros((50, 60), np.uint8)
    flat[10:30, 5:50] = 200
    put("flat.png", b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 60, 50, 8, 0, 0, 0, 0))
        + chunk(b"IDAT", zlib.compress(b"".join(b"\0" + r.tobytes() for r in flat), 9)) + chunk(b"IEND", b""))
    # BMP
    for i, (h, w, bits) in enumerate(((1, 1, 8), (3, 5, 8), (7, 6, 24), (9, 7, 24), (4, 3, 32), (33, 65, 8), (20, 31, 24), (10, 10, 32))):
        put("b%d.bmp" % i, bmp(h, w, bits, rng))
    put("bcol.bmp", bmp(11, 13, 8, rng, palette="colour"))
    put("bpal16.bmp", bmp(11, 13, 8, rng, ncolors=16))
    put("bpal0.bmp", bmp(5, 6, 8, rng, ncolors=0))
    put("boff.bmp", bmp(6, 9, 24, rng, offset_pad=10))
    put("btop.bmp", bmp(6, 9, 24, rng, top_down=True))
    put("bhdr.bmp", bmp(6, 9, 24, rng, hdr=124))
    put("b16.bmp", bmp(6, 8, 16, rng))
    whole = bmp(12, 12, 24, rng)
    put("bshort.bmp", whole[:200])
    put("bhead.bmp", whole[:30])
    put("bnosig.bmp", b"BX" + whole[2:])
    # QOI
    for i, (h, w, c) in enumerate(((1, 1, 3), (5, 7, 3), (5, 7, 4), (40, 33, 3), (40, 33, 4), (64, 64, 4))):
        put("q%d.qoi" % i, qoi(h, w, c, rng))
    put("qshort.qoi", qoi(20, 20, 3, rng, n_ops=15))
    put("qlong.qoi", qoi(4, 4, 4, rng, n_ops=200))
    put("qnoend.qoi", qoi(10, 10, 3, rng, end_marker=False))
    put("qc5.qoi", b"qoif" + struct.pack(">IIBB", 4, 4, 5, 0) + bytes(40))
    put("qsig.qoi", b"qoix" + struct.pack(">IIBB", 4, 4, 3, 0) + bytes(40))
    put("qhead.qoi", b"qoif" + bytes(5))
    # PNM pictures of the repository as sources for the writers
    for f in ("P2", "P3", "P4", "P5", "P6"):
        put(f + ".pnm", open(os.path.join(GOLDEN, f + ".pnm"), "rb").read())
    noise = rng.integers(0, 256, (37, 41, 3), dtype=np.uint8)
    put("noise.ppm", b"P6\n41 37\n255\n" + noise.tobytes())
    runs = np.repeat(rng.integers(0, 256, (30, 4), dtype=np.uint8), 25, axis=1)          # long runs: QOI RUN at its 62 limit
    put("runs.pgm", b"P5\n100 30\n255\n" + runs.tobytes())
    grad = (np.arange(90)[None, :, None] * np.array([1, 2, 3])[None, None, :] + np.arange(70)[:, None, None]).astype(np.uint8)
    put("grad.ppm", b"P6\n90 70\n255\n" + grad.tobytes())                                 # small steps: QOI DIFF / LUMA
    wide = rng.integers(0, 256, (2, 70000), dtype=np.uint8)                               # a PNG line longer than one stored block
    put("wide.pgm", b"P5\n70000 2\n255\n" + wide.tobytes())
    return names


def jobs_for(names):
    """every input -> .ppm / .pgm (what the loader saw), then every decodable picture -> .bmp / .qoi / .png and back."""
    args = ["-f"]
    for n in names:
        stem = n.replace(".", "_")
        args += [n, "-o", stem + ".ppm", n, "-o", stem + ".pgm"]
    for n in names:
        stem = n.replace(".", "_")
        for ext in ("bmp", "qoi", "png"):
            args += [n, "-o", "%s_w.%s" % (stem, ext), "%s_w.%s" % (stem, ext), "-o", "%s_w_%s.ppm" % (stem, ext)]
    return args


@pytest.mark.parametrize("seed", [20241, 7, 99])
def test_formats_against_reference_cli(seed):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/ImCvt_ref not built")
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2:
        names = make_inputs(d, seed)
        assert names == make_inputs(d2, seed)
        args = jobs_for(names)
        rc, out = run(CLI, args, d)
        rc2, out2 = run(REF, args, d2)
        assert rc == rc2
        assert out.decode(errors="replace").split("\n") == out2.decode(errors="replace").split("\n")
        f1, f2 = sorted(os.listdir(d)), sorted(os.listdir(d2))
        assert f1 == f2
        for f in f1:
            a, b = open(os.path.join(d, f), "rb").read(), open(os.path.join(d2, f), "rb").read()
            if f.startswith("qshort"):
                continue        # the reference leaves pixels the data does not reach uninitialised (imageio_qoi.c:238-251)
            assert a == b, f
        assert len(f1) > 400 and out.count(b"failed") + out.count(b"not-yet supported") > 20


def test_format_round_trips():
    """Without the reference: PNM -> {bmp, qoi, png} -> PNM gives the pixels back (gray pictures come back as RGB from QOI)."""
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d)
        for src in ("P5.pnm", "P6.pnm", "noise.ppm", "runs.pgm", "grad.ppm", "wide.pgm"):
            img, is_rgb = load_pnm(os.path.join(d, src))
            for ext in ("bmp", "qoi", "png"):
                rc, out = run(CLI, ["-f", src, "-o", "x." + ext, "x." + ext, "-o", "y.pnm"], d)
                assert rc == 0, out
                back, rgb2 = load_pnm(os.path.join(d, "y.pnm"))
                if rgb2 and not is_rgb:
                    assert (back[..., 0] == back[..., 1]).all() and (back[..., 1] == back[..., 2]).all()
                    back = back[..., 0]
                assert np.array_equal(back, img), (src, ext)


def test_python_zlib_reads_the_png_writer():
    """The stored-deflate PNG the writer makes is a valid zlib stream with correct CRCs (checked by Python's zlib)."""
    with tempfile.TemporaryDirectory() as d:
        make_inputs(d)
        for src in ("P6.pnm", "wide.pgm"):
            img, is_rgb = load_pnm(os.path.join(d, src))
            assert run(CLI, ["-f", src, "-o", "x.png"], d)[0] == 0
            data = open(os.path.join(d, "x.png"), "rb").read()
            assert data[:8] == b"\x89PNG\r\n\x1a\n"
            pos, idat = 8, b""
            while pos < len(data):
                n, t = struct.unpack(">I4s", data[pos:pos + 8])
                body = data[pos + 8:pos + 8 + n]
                assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(t + body)
                if t == b"IDAT":
                    idat += body
                pos += 12 + n
            raw = zlib.decompress(idat)
            h, w = img.shape[:2]
            line = w * (3 if is_rgb else 1) + 1
            assert len(raw) == h * line
            rows = np.frombuffer(raw, np.uint8).reshape(h, line)
            assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(img.shape), img)


def test_format_loaders_survive_damaged_files():
    """Memory safety of the readers: 1500 mutated PNG / BMP / QOI files (bit flips, overwritten header fields, truncations,
    insertions) through the loaders built with AddressSanitizer + UBSan; nothing may be reported.  (20 000 mutations were run
    once when the readers were written; this keeps a sample in the tier.)"""
    import random
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    host = os.path.join(ROOT, "imcvt_b200", "host")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "harness")
        cc = subprocess.run(["gcc", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                             os.path.join(ROOT, "tests", "tools", "fmt_fuzz_harness.c")] + [os.path.join(host, f) for f in ("fmt_png.c", "fmt_bmp.c", "fmt_qoi.c")],
                            capture_output=True)
        if cc.returncode != 0:
            pytest.skip("sanitizer build not available: " + cc.stderr.decode()[-200:])
        seeds = [open(os.path.join(d, n), "rb").read() for n in make_inputs(d) if n.endswith((".png", ".bmp", ".qoi"))]
        rnd = random.Random(77)
        env = dict(os.environ, ASAN_OPTIONS="allocator_may_return_null=1")
        for lot in range(15):
            files = []
            for it in range(100):
                s = bytearray(rnd.choice(seeds))
                for _ in range(rnd.randint(1, 6)):
                    op = rnd.randint(0, 4)
                    if not s:
                        break
                    if op == 0:
                        s[rnd.randrange(len(s))] = rnd.randrange(256)
                    elif op == 1:
                        s[rnd.randrange(len(s))] ^= 1 << rnd.randrange(8)
                    elif op == 2 and len(s) > 4:
                        s = s[:rnd.randrange(len(s))]
                    elif op == 3:
                        i = rnd.randrange(len(s))
                        s[i:i] = bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 8)))
                    elif len(s) > 8:
                        s[rnd.randrange(min(len(s), 64))] = rnd.choice([0, 1, 0x7F, 0x80, 0xFF])
                path = os.path.join(d, "m%d_%d.bin" % (lot, it))
                open(path, "wb").write(bytes(s))
                files.append(path)
            p = subprocess.run([exe] + files, capture_output=True, timeout=600, env=env)
            assert p.returncode == 0 and b"ERROR: AddressSanitizer" not in p.stderr and b"runtime error" not in p.stderr, p.stderr.decode(errors="replace")[-1500:]
            for f in files:
                os.remove(f)


def test_damaged_pnm_and_bmp_files_give_what_the_reference_gives():
    """Differential sample: mutated PNM and BMP files through this command line and through the reference's -- text, exit code and
    every written file equal.  (PNG and QOI are left out here: where the data does not reach, the reference's decoders show
    uninitialised memory; 12 000 such files were compared once by hand)"""
    import random
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/ImCvt_ref not built")
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2, tempfile.TemporaryDirectory() as src:
        names = make_inputs(src)
        seeds = [(".bmp", open(os.path.join(src, n), "rb").read()) for n in names if n.endswith(".bmp")]
        seeds += [(".pnm", open(os.path.join(src, n), "rb").read()[:400]) for n in names if n.endswith(".pnm")]
        seeds += [(".pnm", s) for s in (b"P1\n# c\n3 2\n1 0 1\n0 11 0\n", b"P5 # x\n4 2 # y\n200 trailing\n" + bytes(range(8)), b"P2\n2 2\n255\n0 256 511 7\n",
                                        b"P4\n9 3\n" + bytes(6), b"P3\n2 1\n255\n1 2 3 4 5 6\n", b"P6\n2 2\n255\n" + bytes(range(12)))]
        rnd = random.Random(2024)
        compared = 0
        for it in range(400):
            ext, s0 = rnd.choice(seeds)
            s = bytearray(s0)
            for _ in range(rnd.randint(1, 4)):
                if not s:
                    break
                op = rnd.randint(0, 4)
                if op == 0:
                    s[rnd.randrange(min(len(s), 60))] = rnd.randrange(256)
                elif op == 1:
                    s[rnd.randrange(len(s))] ^= 1 << rnd.randrange(8)
                elif op == 2 and len(s) > 4:
                    s = s[:rnd.randrange(len(s))]
                elif op == 3:
                    i = rnd.randrange(min(len(s), 30))
                    s[i:i] = rnd.choice([b"9", b" ", b"\n", b"#", b"-", b"0", b"1 ", b"\r", b"\t"])
                else:
                    i = rnd.randrange(min(len(s), 30))
                    s[i:i + 1] = rnd.choice([b"", b" ", b"\t", b"\r", b"#"])
            if ext == ".bmp" and len(s) >= 26:
                bw, bh = struct.unpack("<II", bytes(s[18:26]))
                if bw * bh > (1 << 20):
                    continue                                # a damaged header would make both programs write gigabytes
            args = ["-f", "m" + ext, "-o", "o.ppm", "m" + ext, "-o", "o.pgm", "m" + ext, "-o", "o.bmp"]
            for dd in (d, d2):
                open(os.path.join(dd, "m" + ext), "wb").write(bytes(s))
                for o in ("o.ppm", "o.pgm", "o.bmp"):
                    if os.path.exists(os.path.join(dd, o)):
                        os.remove(os.path.join(dd, o))
            try:
                want = subprocess.run([REF] + args, cwd=d2, capture_output=True, timeout=10)
            except subprocess.TimeoutExpired:
                continue                                    # a damaged header made the reference write gigabytes, byte by byte
            if want.returncode < 0:
                continue                                    # ... or crash
            got = subprocess.run([CLI] + args, cwd=d, capture_output=True, timeout=60)
            assert (got.returncode, got.stdout) == (want.returncode, want.stdout), bytes(s)[:80]
            for o in ("o.ppm", "o.pgm", "o.bmp"):
                a, b = os.path.join(d, o), os.path.join(d2, o)
                assert os.path.exists(a) == os.path.exists(b), (o, bytes(s)[:80])
                if os.path.exists(a):
                    assert open(a, "rb").read() == open(b, "rb").read(), (o, bytes(s)[:80])
            compared += 1
        assert compared > 200
