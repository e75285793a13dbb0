"""Corrupt-file fuzzing of the native PNG decoder behind rnb_load_png_rgba16 (run in a subprocess by tests/test_dataset_ingest.py so that a
crash of the C code fails the test instead of killing pytest): truncations, byte flips, hostile IHDR sizes with a valid CRC, corrupted
chunk lengths.  Every file must either decode to an RGBA16 array or be rejected with an error — never abort, never allocate for a
header the compressed data cannot back.  usage: png_fuzz.py <seed> <scratch dir> <iterations>"""
import sys, os, zlib, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, cv2
import rnb_loader
pkg = rnb_loader.load_package()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
tmp = sys.argv[2] if len(sys.argv) > 2 else "/tmp/pngfuzz"; os.makedirs(tmp, exist_ok=True)
bases = []
for i, (dt, ch) in enumerate([(np.uint16, 4), (np.uint16, 3), (np.uint8, 4), (np.uint8, 1), (np.uint16, 1)]):
    img = rng.integers(0, np.iinfo(dt).max, size=(37, 29, ch) if ch > 1 else (37, 29)).astype(dt)
    p = os.path.join(tmp, "base%d.png" % i); cv2.imwrite(p, img); bases.append(open(p, "rb").read())
n_ok = n_err = 0
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 400):
    b = bytearray(bases[it % len(bases)])
    mode = it % 4
    if mode == 0: b = b[: int(rng.integers(0, len(b)))]                        # truncation
    elif mode == 1:
        for _ in range(int(rng.integers(1, 6))): b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))      # byte flips
    elif mode == 2:                                                             # hostile IHDR dimensions with a fixed-up CRC
        w, h = int(rng.choice([0, 1, 2**31 - 1, 2**32 - 1, 65536, 29])), int(rng.choice([0, 1, 2**31 - 1, 2**32 - 1, 65536, 37]))
        ih = bytearray(b[12:29]); ih[4:8] = struct.pack(">I", w); ih[8:12] = struct.pack(">I", h)
        b[12:29] = ih; b[29:33] = struct.pack(">I", zlib.crc32(bytes(ih)) & 0xffffffff)
    else:                                                                       # chunk length field corrupted
        b[33:37] = struct.pack(">I", int(rng.choice([0, 1, 2**31 - 1, 2**32 - 1, len(b) * 2])))
    p = os.path.join(tmp, "f.png"); open(p, "wb").write(bytes(b))
    try:
        a = pkg.load_png_rgba16(p)
        assert a.ndim == 3 and a.shape[2] == 4 and a.dtype == np.uint16
        n_ok += 1
    except pkg.RnbError:
        n_err += 1
print("decoded", n_ok, "rejected", n_err)
