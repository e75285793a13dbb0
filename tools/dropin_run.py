#!/usr/bin/env python
"""Drop-in evidence, run on the B200 box: the reference's own CLI (`oracle/_ref/bin/testbed`, unmodified sources) next to the same CLI with the
training / meshing / snapshot hooks of shim/rnb_testbed_shim.h routed through librnb_b200.so (`oracle/_ref/bin/testbed_rnb`), driven exactly
like the reference's orchestration drives it (ref:rnb_neus2/pipeline.py:27-103): stage 1 (`--save-snapshot [--no-albedo]`), then stage 2 from
the stage-1 snapshot (`--opti-lights --snapshot S1 --resolution R --save-mesh --save-snapshot --free-memory`), and — when the reference's
Python package has been staged under oracle/_ref/pipeline by `make -f Makefile.ref data` — `run_pipeline.py` end to end with `--testbed`
pointing at either binary.  Writes one JSON record: return codes, wall times, the `iteration= loss=` lines, snapshot keys and blob sizes,
mesh sizes, mesh-to-mesh and mesh-to-ground-truth distances.  TEST / EVIDENCE TOOLING (not product code).

usage: python tools/dropin_run.py OUT_DIR [--iters 3000] [--res 256] [--views 24] [--albedo] [--pipeline] [--only rnb|stock]
"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader  # noqa: E402
import ref_scene  # noqa: E402

BIN = {"stock": os.path.join(ROOT, "oracle/_ref/bin/testbed"), "rnb": os.path.join(ROOT, "oracle/_ref/bin/testbed_rnb")}


def run(cmd, log_path, timeout, env=None):
    t0 = time.time()
    with open(log_path, "w") as f:
        try:
            rc = subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, timeout=timeout, env=env).returncode
        except subprocess.TimeoutExpired:
            rc = -999
    return rc, time.time() - t0


def loss_lines(log_path):
    out = []
    for line in open(log_path, errors="replace"):
        m = re.search(r"iteration=(\d+) loss=([0-9.eE+-]+)", line)
        if m:
            out.append((int(m.group(1)), float(m.group(2))))
    return out


def obj_vertices(path):
    v = []
    nf = 0
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                nf += 1
    return np.asarray(v, np.float64).reshape(-1, 3), nf


def mesh_stats(path, scene):
    """vertices / faces and the distance of the vertices to the analytic surface the scene was rendered from (world frame of the OBJ:
    p = (v - 0.5) / 0.5 for the identity n2w written by tests/ref_scene.py)"""
    v, nf = obj_vertices(path)
    v = v[np.abs(v).sum(1) > 0]                                   # the vertex array is padded to a multiple of 128 with zeros (marching_cubes.cu:810)
    c = (np.asarray(scene.CENTER, np.float64) - 0.5) / 0.5; ax = np.asarray(scene.AXES, np.float64) / 0.5
    q = (v - c) / ax
    r = np.linalg.norm(q, axis=1)
    d = np.abs(r - 1.0) * ax.min()                                # first-order distance to the ellipsoid (lower bound scale: smallest semi-axis)
    return dict(file=os.path.basename(path), bytes=os.path.getsize(path), n_vertices=int(len(v)), n_faces=int(nf),
                dist_to_analytic_surface_mean=float(d.mean()), dist_to_analytic_surface_p99=float(np.percentile(d, 99)), dist_to_analytic_surface_max=float(d.max())), v


def snapshot_summary(path):
    import importlib
    pkg = rnb_loader.load_package()
    snap = importlib.import_module(pkg.__name__ + ".snapshot")
    cfg = snap.read_snapshot(path)
    s = cfg["snapshot"]
    out = dict(file=os.path.basename(path), bytes=os.path.getsize(path), top_keys=sorted(cfg.keys()), snapshot_keys=sorted(s.keys()))
    for k in ("training_step", "rays_per_batch", "measured_batch_size", "measured_batch_size_before_compaction", "n_params", "density_grid_size", "loss"):
        if k in s:
            out[k] = s[k]
    if "nerf" in s and "rgb" in s["nerf"]:
        out["nerf_rgb"] = dict(s["nerf"]["rgb"])
    for k in ("params_binary", "density_grid_binary"):
        if k in s:
            b = s[k]
            out[k + "_bytes"] = len(b)
            a = np.frombuffer(bytes(b), np.float16).astype(np.float32)
            out[k + "_finite"] = bool(np.isfinite(a).all()); out[k + "_l2"] = float(np.linalg.norm(a))
    return out


def two_stage(name, scene_dir, iters, res, albedo, timeout, rec):
    """ref:rnb_neus2/pipeline.py:56-103 (run_two_stage), argv for argv"""
    it1 = int(iters * 2 / 3)
    common = ["--mask-weight", "1.0"]
    base = [BIN[name], "--scene", scene_dir + "/", "--no-gui"]
    s1 = base + ["--maxiter", str(it1)] + common + ["--save-snapshot"] + ([] if albedo else ["--no-albedo"])
    rc1, w1 = run(s1, os.path.join(rec["dir"], name + "_stage1.log"), timeout)
    snap1 = os.path.join(scene_dir, "output", "snapshot_%d.msgpack" % it1)
    r = dict(stage1=dict(argv=s1[1:], rc=rc1, wall_s=round(w1, 2), loss_lines=loss_lines(os.path.join(rec["dir"], name + "_stage1.log")), snapshot_exists=os.path.exists(snap1)))
    snapshot = snap1
    if rc1 == 0 and os.path.exists(snapshot):
        r["stage1"]["snapshot"] = snapshot_summary(snapshot)
        s2 = base + ["--maxiter", str(iters)] + common + ["--opti-lights", "--snapshot", snapshot, "--resolution", str(res), "--save-mesh", "--save-snapshot", "--free-memory"] + ([] if albedo else ["--no-albedo"])
        rc2, w2 = run(s2, os.path.join(rec["dir"], name + "_stage2.log"), timeout)
        r["stage2"] = dict(argv=s2[1:], rc=rc2, wall_s=round(w2, 2), wall_note="includes the 10 s sleep of --free-memory (src/main.cu:458)", loss_lines=loss_lines(os.path.join(rec["dir"], name + "_stage2.log")))
        snap2 = os.path.join(scene_dir, "output", "snapshot_%d.msgpack" % iters)
        if os.path.exists(snap2):
            r["stage2"]["snapshot"] = snapshot_summary(snap2)
        r["stage2"]["output_files"] = sorted(os.listdir(os.path.join(scene_dir, "output")))
    return r


def write_rnb_input(out_dir, views):
    """The reference's RNb input layout (ref:rnb_neus2/dataloaders/rnb_loader.py:1-112): cameras.npz (world_mat_i = K [R|t] world->pixel,
    scale_mat_i), normal/NNN.png, mask/NNN.png — from the same synthetic views."""
    import cv2
    for sub in ("normal", "mask"):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    cams = {}
    w, h = views[0]["w"], views[0]["h"]
    for i, v in enumerate(views):
        xf = np.asarray(v["xform"], np.float64)
        R = xf[:9].reshape(3, 3).T; t = (xf[9:12] - 0.5) / 0.5
        c2w = np.eye(4); c2w[:3, :3] = R; c2w[:3, 3] = t
        K = np.eye(4); K[0, 0] = v["fx"]; K[1, 1] = v["fy"]; K[0, 2] = v["cx"] * w; K[1, 2] = v["cy"] * h
        cams["world_mat_%d" % i] = (K @ np.linalg.inv(c2w)).astype(np.float64)
        cams["scale_mat_%d" % i] = np.eye(4)
        nm = v["normal"]
        cv2.imwrite(os.path.join(out_dir, "normal", "%03d.png" % i), np.ascontiguousarray(nm[..., [2, 1, 0]]))       # cv2 stores BGR: file order stays RGB
        cv2.imwrite(os.path.join(out_dir, "mask", "%03d.png" % i), (nm[..., 3] > 0).astype(np.uint8) * 255)
        if v.get("albedo") is not None:
            os.makedirs(os.path.join(out_dir, "albedo"), exist_ok=True)
            cv2.imwrite(os.path.join(out_dir, "albedo", "%03d.png" % i), np.ascontiguousarray(v["albedo"][..., [2, 1, 0]]))
    np.savez(os.path.join(out_dir, "cameras.npz"), **cams)


def run_pipeline(name, inp, out, iters, res, timeout, rec, has_albedo=False):
    """has_albedo: BASELINE configs[2] — `run_pipeline.py --has-albedo`: phase 1 geometry (--no-albedo, mesh at 512), the albedo-scaling stage, then the two-stage run with
    albedo (ref:rnb_neus2/pipeline.py:106-175).  The reference's albedo_scaling.py needs trimesh; the staged copy oracle/_ref/pipeline_rnb carries the one-line import
    change of INTEGRATION.md §7, so the stage is served by rnb-neus2_b200/albedo_scaling.py over the GPU ray / mesh queries."""
    pipe = os.path.join(ROOT, "oracle/_ref/pipeline_rnb" if has_albedo else "oracle/_ref/pipeline")
    pylib = os.path.join(rec["dir"], "pylib"); os.makedirs(pylib, exist_ok=True)
    link = os.path.join(pylib, "rnb_neus2_b200")
    if not os.path.exists(link):
        os.symlink(os.path.join(ROOT, "rnb-neus2_b200"), link)
    env = dict(os.environ); env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests/stubs"), pipe, pylib, env.get("PYTHONPATH", "")])
    cmd = [sys.executable, os.path.join(pipe, "run_pipeline.py"), "--input", inp, "--testbed", BIN[name], "--output", out, "--max-steps", str(iters), "--mesh-resolution", str(res),
           "--scaling-mode", "none"] + (["--has-albedo", "--n-samples", "2000"] if has_albedo else [])
    log = os.path.join(rec["dir"], name + ("_pipeline_albedo.log" if has_albedo else "_pipeline.log"))
    rc, wall = run(cmd, log, timeout, env)
    txt = open(log, errors="replace").read()
    r = dict(argv=cmd[1:], rc=rc, wall_s=round(wall, 2), complete="=== Pipeline complete ===" in txt, mesh_exists=os.path.exists(os.path.join(out, "mesh.obj")),
             stage_lines=[line.strip()[:160] for line in txt.split("\n") if re.search(r"Stage \d|Phase \d|completed|Pipeline complete|Mesh exported|failed|Albedo|scale ratio|Scale", line)][:40],
             note="postprocess_mesh imports trimesh, which is not in this image: tests/stubs/trimesh is a stand-in that keeps the mesh as it is (load / split / fix_normals / export)")
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out"); ap.add_argument("--iters", type=int, default=3000); ap.add_argument("--res", type=int, default=256); ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--width", type=int, default=400); ap.add_argument("--height", type=int, default=300)
    ap.add_argument("--albedo", action="store_true"); ap.add_argument("--pipeline", action="store_true"); ap.add_argument("--only", default="")
    ap.add_argument("--pipeline-albedo", action="store_true", help="run_pipeline.py --has-albedo (two-phase with albedo scaling) against testbed_rnb"); ap.add_argument("--skip-two-stage", action="store_true")
    ap.add_argument("--timeout", type=int, default=600)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    scene = rnb_loader.load_scene()
    views = scene.make_scene(a.views, a.width, a.height, with_albedo=a.albedo)
    names = [a.only] if a.only else ["stock", "rnb"]
    rec = dict(dir=a.out, iters=a.iters, mesh_resolution=a.res, n_views=a.views, image=[a.width, a.height], albedo=bool(a.albedo), runs={})
    try:
        rec["gpu"] = subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    except Exception:      # noqa: BLE001
        pass
    meshes = {}
    first_sd = None
    for name in ([] if a.skip_two_stage else names):
        sd = os.path.join(a.out, "scene_" + name)
        if first_sd is None:
            t0 = time.time(); ref_scene.write_scene(sd, views, workers=min(16, os.cpu_count() or 8)); rec["scene_write_s"] = round(time.time() - t0, 1); first_sd = sd
        else:      # same images (symlinked), own transform.json and output/
            os.makedirs(os.path.join(sd, "output"), exist_ok=True)
            for sub in ("normals", "albedos"):
                if not os.path.exists(os.path.join(sd, sub)):
                    os.symlink(os.path.abspath(os.path.join(first_sd, sub)), os.path.join(sd, sub))
            import shutil
            shutil.copyfile(os.path.join(first_sd, "transform.json"), os.path.join(sd, "transform.json"))
        r = two_stage(name, sd, a.iters, a.res, a.albedo, a.timeout, rec)
        m = glob.glob(os.path.join(sd, "output", "mesh_*.obj"))
        if m:
            r["mesh"], meshes[name] = mesh_stats(m[0], scene)
        rec["runs"][name] = r
        print(name, "stage1 rc", r["stage1"]["rc"], "wall", r["stage1"]["wall_s"], "stage2", r.get("stage2", {}).get("rc"), r.get("stage2", {}).get("wall_s"), "mesh", r.get("mesh", {}).get("n_vertices"),
              r.get("mesh", {}).get("dist_to_analytic_surface_mean"), flush=True)
    if len(meshes) == 2:
        from scipy.spatial import cKDTree
        va, vb = meshes["stock"], meshes["rnb"]
        d_ab = cKDTree(vb).query(va)[0]; d_ba = cKDTree(va).query(vb)[0]
        rec["mesh_stock_vs_rnb"] = dict(nearest_vertex_stock_to_rnb_mean=float(d_ab.mean()), nearest_vertex_stock_to_rnb_max=float(d_ab.max()),
                                        nearest_vertex_rnb_to_stock_mean=float(d_ba.mean()), nearest_vertex_rnb_to_stock_max=float(d_ba.max()), lattice_spacing=2.0 / a.res,
                                        note="world frame, scene radius ~0.6; independent training runs of a non-deterministic reference: distances of the order of the lattice spacing mean the same surface")
        print("mesh stock vs rnb:", rec["mesh_stock_vs_rnb"], flush=True)
        # the two binaries must write the same set of files and snapshots with the same keys
        for st in ("stage1", "stage2"):
            sa, sb = rec["runs"]["stock"].get(st, {}).get("snapshot"), rec["runs"]["rnb"].get(st, {}).get("snapshot")
            if sa and sb:
                rec.setdefault("snapshot_keys_equal", {})[st] = (sa["snapshot_keys"] == sb["snapshot_keys"] and sa["top_keys"] == sb["top_keys"] and sa.get("params_binary_bytes") == sb.get("params_binary_bytes")
                                                                 and sa.get("density_grid_binary_bytes") == sb.get("density_grid_binary_bytes"))
        fa, fb = rec["runs"]["stock"].get("stage2", {}).get("output_files"), rec["runs"]["rnb"].get("stage2", {}).get("output_files")
        rec["output_files_equal"] = fa == fb and fa is not None
    if a.pipeline and os.path.exists(os.path.join(ROOT, "oracle/_ref/pipeline/run_pipeline.py")):
        inp = os.path.join(a.out, "rnb_input"); write_rnb_input(inp, views)
        rec["run_pipeline"] = {}
        for name in names:
            out = os.path.join(a.out, "pipeline_" + name)
            rec["run_pipeline"][name] = run_pipeline(name, inp, out, a.iters, a.res, a.timeout * 2, rec)
            mp = os.path.join(out, "mesh.obj")
            if os.path.exists(mp):
                rec["run_pipeline"][name]["mesh"], _ = mesh_stats(mp, scene)
            print("run_pipeline.py via", name, {k: rec["run_pipeline"][name][k] for k in ("rc", "wall_s", "complete", "mesh_exists")}, flush=True)
    if a.pipeline_albedo and os.path.exists(os.path.join(ROOT, "oracle/_ref/pipeline_rnb/run_pipeline.py")):
        views_a = scene.make_scene(a.views, a.width, a.height, with_albedo=True)
        # per-view gains on the albedo maps (what the scaling stage is there to undo): the stage must come back with ratios ~ 1 / gain
        gains = 0.7 + 0.6 * np.random.RandomState(3).rand(a.views)
        for v, gn in zip(views_a, gains):
            al = v["albedo"].astype(np.float64); al[..., :3] = np.clip(al[..., :3] * gn, 0, 65535); v["albedo"] = al.astype(np.uint16)
        inp = os.path.join(a.out, "rnb_input_albedo"); write_rnb_input(inp, views_a)
        out = os.path.join(a.out, "pipeline_albedo_rnb")
        r = run_pipeline("rnb", inp, out, a.iters, a.res, a.timeout * 3, rec, has_albedo=True)
        mp = os.path.join(out, "mesh.obj")
        if os.path.exists(mp):
            r["mesh"], _ = mesh_stats(mp, scene)
        r["albedo_gains_applied"] = [round(float(x), 4) for x in gains]
        rec["run_pipeline_has_albedo"] = r
        print("run_pipeline.py --has-albedo via rnb", {k: r[k] for k in ("rc", "wall_s", "complete", "mesh_exists")}, r.get("mesh", {}).get("dist_to_analytic_surface_mean"), flush=True)
    json.dump(rec, open(os.path.join(a.out, "dropin_record.json"), "w"), indent=1)
    print("record:", os.path.join(a.out, "dropin_record.json"))


if __name__ == "__main__":
    main()
