"""Snapshot hand-off: the reference's `.msgpack` snapshot format, read and written on the host side.

Mirrors Testbed::save_snapshot / load_snapshot (reference src/testbed.cu:3280-3314, 3333-3390) and Trainer::serialize /
deserialize (tiny-cuda-nn trainer.h:281-304).  A snapshot is the network config JSON with a "snapshot" object added, encoded
with MessagePack the way nlohmann::json::to_msgpack does it (objects with sorted keys, smallest integer encodings, binary blobs
as bin8/16/32):

    snapshot.n_params               parameter count
    snapshot.params_binary          the INFERENCE (EMA) parameters, binary16, in the order of nerf_network.h:539-583
    snapshot.density_grid_size      128
    snapshot.density_grid_binary    the occupancy (density) grid, binary16, Morton order
    snapshot.training_step / loss
    snapshot.nerf.aabb_scale, snapshot.nerf.rgb.{rays_per_batch, measured_batch_size, measured_batch_size_before_compaction}
    snapshot.nerf.dataset           (optional) dataset metadata; only read when no dataset was loaded
    snapshot.rotation / transition / local_rotation / local_transition
                                    binary16 blobs of the global / local movement buffers (nerf_network.h:989-1081); a static scene
                                    keeps their initial values (identity, zero), which is what this module writes

The codec below is self-contained (no third-party package on the product path); tests cross-check it against the `msgpack`
package and against a snapshot written by the reference build.  Adam moments are not part of a snapshot
(include_optimizer_state is false in src/main.cu:468): training restarts them, in the reference as here."""
import struct

import numpy as np

GRID_SIZE = 128


# ---- MessagePack, the subset nlohmann::json emits -------------------------------------------------------------------------
def _pack_into(o, out):
    if o is None:
        out.append(b"\xc0")
    elif o is True:
        out.append(b"\xc3")
    elif o is False:
        out.append(b"\xc2")
    elif isinstance(o, (int, np.integer)):
        o = int(o)
        if o >= 0:
            if o < 128:
                out.append(struct.pack("B", o))
            elif o < 1 << 8:
                out.append(b"\xcc" + struct.pack("B", o))
            elif o < 1 << 16:
                out.append(b"\xcd" + struct.pack(">H", o))
            elif o < 1 << 32:
                out.append(b"\xce" + struct.pack(">I", o))
            elif o < 1 << 64:
                out.append(b"\xcf" + struct.pack(">Q", o))
            else:
                raise OverflowError("integer too large for MessagePack")
        elif o >= -32:
            out.append(struct.pack("b", o))
        elif o >= -(1 << 7):
            out.append(b"\xd0" + struct.pack("b", o))
        elif o >= -(1 << 15):
            out.append(b"\xd1" + struct.pack(">h", o))
        elif o >= -(1 << 31):
            out.append(b"\xd2" + struct.pack(">i", o))
        else:
            out.append(b"\xd3" + struct.pack(">q", o))
    elif isinstance(o, (float, np.floating)):
        o = float(o)
        # nlohmann::json::write_compact_float: the short form only for values inside the binary32 range that survive the round trip; NaN fails every
        # comparison and infinity the range check, so both are written as binary64 (0xcb)
        f32 = struct.unpack(">f", struct.pack(">f", o))[0] if abs(o) <= 3.4028234663852886e38 else None
        if f32 is not None and f32 == o:
            out.append(b"\xca" + struct.pack(">f", o))
        else:
            out.append(b"\xcb" + struct.pack(">d", o))
    elif isinstance(o, str):
        b = o.encode("utf-8"); n = len(b)
        if n < 32:
            out.append(struct.pack("B", 0xA0 | n))
        elif n < 1 << 8:
            out.append(b"\xd9" + struct.pack("B", n))
        elif n < 1 << 16:
            out.append(b"\xda" + struct.pack(">H", n))
        else:
            out.append(b"\xdb" + struct.pack(">I", n))
        out.append(b)
    elif isinstance(o, (bytes, bytearray, memoryview)):
        n = len(o)
        if n < 1 << 8:
            out.append(b"\xc4" + struct.pack("B", n))
        elif n < 1 << 16:
            out.append(b"\xc5" + struct.pack(">H", n))
        else:
            out.append(b"\xc6" + struct.pack(">I", n))
        out.append(bytes(o))
    elif isinstance(o, (list, tuple)):
        n = len(o)
        if n < 16:
            out.append(struct.pack("B", 0x90 | n))
        elif n < 1 << 16:
            out.append(b"\xdc" + struct.pack(">H", n))
        else:
            out.append(b"\xdd" + struct.pack(">I", n))
        for v in o:
            _pack_into(v, out)
    elif isinstance(o, dict):
        n = len(o)
        if n < 16:
            out.append(struct.pack("B", 0x80 | n))
        elif n < 1 << 16:
            out.append(b"\xde" + struct.pack(">H", n))
        else:
            out.append(b"\xdf" + struct.pack(">I", n))
        for k in sorted(o):                                   # nlohmann::json objects are std::map: keys in byte order
            if not isinstance(k, str):
                raise TypeError("object keys must be strings")
            _pack_into(k, out); _pack_into(o[k], out)
    elif isinstance(o, np.ndarray):
        _pack_into(o.tolist(), out)
    else:
        raise TypeError("cannot encode %r" % type(o))


def packb(o):
    out = []
    _pack_into(o, out)
    return b"".join(out)


class _Reader:
    def __init__(self, b):
        self.b = memoryview(b); self.o = 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated MessagePack data")
        v = self.b[self.o:self.o + n]; self.o += n
        return v

    def num(self, fmt):
        return struct.unpack(fmt, self.take(struct.calcsize(fmt)))[0]

    def value(self):
        t = self.num("B")
        if t < 0x80:
            return t
        if t >= 0xE0:
            return t - 256
        if 0x80 <= t <= 0x8F:
            return self.map(t & 15)
        if 0x90 <= t <= 0x9F:
            return [self.value() for _ in range(t & 15)]
        if 0xA0 <= t <= 0xBF:
            return bytes(self.take(t & 31)).decode("utf-8")
        if t == 0xC0:
            return None
        if t == 0xC2:
            return False
        if t == 0xC3:
            return True
        if t in (0xC4, 0xC5, 0xC6):
            return bytes(self.take(self.num({0xC4: "B", 0xC5: ">H", 0xC6: ">I"}[t])))
        if t in (0xC7, 0xC8, 0xC9):                           # ext 8/16/32 (nlohmann: binary with a subtype): payload only
            n = self.num({0xC7: "B", 0xC8: ">H", 0xC9: ">I"}[t]); self.take(1)
            return bytes(self.take(n))
        if t == 0xCA:
            return self.num(">f")
        if t == 0xCB:
            return self.num(">d")
        if 0xCC <= t <= 0xCF:
            return self.num(("B", ">H", ">I", ">Q")[t - 0xCC])
        if 0xD0 <= t <= 0xD3:
            return self.num(("b", ">h", ">i", ">q")[t - 0xD0])
        if 0xD4 <= t <= 0xD8:                                 # fixext 1..16
            self.take(1)
            return bytes(self.take(1 << (t - 0xD4)))
        if t in (0xD9, 0xDA, 0xDB):
            return bytes(self.take(self.num({0xD9: "B", 0xDA: ">H", 0xDB: ">I"}[t]))).decode("utf-8")
        if t in (0xDC, 0xDD):
            return [self.value() for _ in range(self.num(">H" if t == 0xDC else ">I"))]
        if t in (0xDE, 0xDF):
            return self.map(self.num(">H" if t == 0xDE else ">I"))
        raise ValueError("unsupported MessagePack type byte 0x%02x" % t)

    def map(self, n):
        d = {}
        for _ in range(n):
            k = self.value()
            d[k] = self.value()
        return d


def unpackb(b):
    r = _Reader(b)
    v = r.value()
    if r.o != len(r.b):
        raise ValueError("trailing bytes after the MessagePack value")
    return v


# ---- snapshot object ------------------------------------------------------------------------------------------------------
def _half_blob(values):
    return np.asarray(values, np.float16).tobytes()


def movement_defaults():
    """Initial values of the movement buffers of a static scene, as the reference build writes them (checked against its own
    snapshot, tests/golden/ref_snapshot_small.npz): accumulated rotation = 3x3 identity in 12 binary16 slots and transition = 4
    zeros (nerf_network.h:76-80, 852-905); the per-frame delta network keeps a 6D rotation (1,0,0,0,1,0 in 8 slots) and 4 zeros."""
    rot = np.zeros(12, np.float16); rot[[0, 4, 8]] = 1.0
    lrot = np.zeros(8, np.float16); lrot[[0, 4]] = 1.0
    return {"rotation": rot.tobytes(), "transition": _half_blob(np.zeros(4)), "local_rotation": lrot.tobytes(), "local_transition": _half_blob(np.zeros(4))}


def build_snapshot(network_config, params_fp16, density_grid, training_step, loss, rays_per_batch, measured_batch_size, measured_batch_size_before_compaction,
                   aabb_scale=1, dataset=None, movement=None):
    """The dict Testbed::save_snapshot serialises: `network_config` (already merged with its parents) + "snapshot"."""
    params_fp16 = np.ascontiguousarray(params_fp16, np.float16)
    cfg = {k: v for k, v in network_config.items() if k != "snapshot"}
    snap = {"n_params": int(params_fp16.size), "params_binary": params_fp16.tobytes(),
            "density_grid_size": GRID_SIZE, "density_grid_binary": np.asarray(density_grid, np.float32).astype(np.float16).tobytes(),
            "training_step": int(training_step), "loss": float(np.float32(loss)),
            "nerf": {"aabb_scale": int(aabb_scale),
                     "rgb": {"rays_per_batch": int(rays_per_batch), "measured_batch_size": int(measured_batch_size),
                             "measured_batch_size_before_compaction": int(measured_batch_size_before_compaction)}}}
    if dataset is not None:
        snap["nerf"]["dataset"] = dataset
    snap.update(movement_defaults() if movement is None else movement)
    cfg["snapshot"] = snap
    return cfg


def parse_snapshot(cfg):
    """-> dict(params_fp16, density_grid (float32, may be empty), training_step, loss, rays_per_batch, measured_batch_size,
    measured_batch_size_before_compaction, aabb_scale).  Raises like the reference on a file without a snapshot / wrong grid size."""
    if "snapshot" not in cfg:
        raise ValueError("File does not contain a snapshot.")
    s = cfg["snapshot"]
    if s.get("density_grid_size") != GRID_SIZE:
        raise ValueError("Incompatible grid size.")
    p = np.frombuffer(s["params_binary"], np.float16)
    if "n_params" in s and int(s["n_params"]) != p.size:
        raise ValueError("params_binary does not hold n_params binary16 values")
    g = np.frombuffer(s["density_grid_binary"], np.float16).astype(np.float32)
    if g.size not in (0, GRID_SIZE ** 3):
        raise ValueError("Incompatible number of grid cascades.")
    rgb = s["nerf"]["rgb"]
    return dict(params_fp16=p, density_grid=g, training_step=int(s["training_step"]), loss=float(s["loss"]), rays_per_batch=int(rgb["rays_per_batch"]),
                measured_batch_size=int(rgb["measured_batch_size"]), measured_batch_size_before_compaction=int(rgb["measured_batch_size_before_compaction"]),
                aabb_scale=int(s["nerf"].get("aabb_scale", 1)))


def write_snapshot(path, cfg):
    with open(path, "wb") as f:
        f.write(packb(cfg))


def read_snapshot(path):
    with open(path, "rb") as f:
        return unpackb(f.read())
