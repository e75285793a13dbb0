#!/usr/bin/env python
"""Pin the mesh path (SURVEY §8(f) N1 + N2) against the UNMODIFIED reference on the GPU box.  TEST INFRASTRUCTURE.

    python tests/ref_pin_mesh.py --config small|full --steps 60 --mesh 64 --out gpurun_out/refpin_mesh

1. trains the reference (oracle/_ref/bin/ref_harness) for a few steps on a synthetic scene, then lets it run
   get_density_on_grid and compute_and_save_marching_cubes_mesh exactly as src/main.cu:460 does (+ a PLY of the same mesh);
   the harness dumps the SDF lattice, the inference parameters, m_mesh.{verts, vert_normals, vert_colors, indices} and the files;
2. compares, on the reference's own lattice: the oracle's and the CUDA path's marching cubes against the reference mesh as
   SETS (the reference numbers vertices and triangles in atomicAdd arrival order): vertex positions bit-exact, triangles
   (as ordered position triples up to rotation) bit-exact, area-weighted normals to 1e-5 of their norm;
3. the text writers on the reference's own arrays: oracle fprintf and the GPU formatter must reproduce ref_mesh.obj / .ply
   byte for byte;
4. N1: rnb_sdf_on_grid and the oracle's SDF with the dumped inference parameters against the reference lattice; vertex
   colours of the CUDA path at the reference's vertices against vert_colors;
5. writes summary_mesh_<config>.json and, with --golden NAME, golden_mesh_NAME.npz (fixture for tests/golden/).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader                                                   # noqa: E402
import oracle_binding as ob                                         # noqa: E402
from oracle_binding import Oracle                                   # noqa: E402
from common import SMALL, FULL, product_config, rel_err            # noqa: E402
import ref_scene                                                    # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")


def h2f(a):
    return np.asarray(a, np.uint16).view(np.float16).astype(np.float32)


def vertex_rows(V):
    """sortable view of float32 [n,3] rows by bit pattern"""
    return np.ascontiguousarray(V, np.float32).view(np.uint32).reshape(-1, 3)


def sort_rows(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


def canonical_triangles(V, F):
    """[t, 9] uint32: the three vertex positions of each triangle (bit patterns), rotated so that the smallest vertex comes
    first (orientation kept), rows sorted."""
    P = vertex_rows(V)[np.asarray(F, np.int64).reshape(-1, 3)]                    # [t, 3, 3]
    _, inv = np.unique(P.reshape(-1, 3), axis=0, return_inverse=True)
    order = inv.reshape(-1, 3).argmin(1)
    idx = (order[:, None] + np.arange(3)[None, :]) % 3
    R = np.take_along_axis(P, idx[:, :, None], axis=1).reshape(-1, 9)
    return sort_rows(R)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--mesh", type=int, default=60)
    ap.add_argument("--out", default="gpurun_out/refpin_mesh")
    ap.add_argument("--work", default="/tmp/refpin_mesh")
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--golden", default="", help="write golden_mesh_<name>.npz (fixture for tests/golden/)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    work = args.work + "_" + args.config
    scene_dir = os.path.join(work, "scene"); dump = os.path.join(work, "dump")
    os.makedirs(dump, exist_ok=True)
    scene = rnb_loader.load_scene()
    views0 = scene.make_scene(args.views, args.res, args.res, with_albedo=True)
    n2w = np.eye(4); n2w[:3, :3] *= 1.75; n2w[:3, 3] = (0.125, -2.5, 31.0)
    ref_scene.write_scene(scene_dir, views0, n2w=n2w)
    cfgd = SMALL if args.config == "small" else FULL
    net_cfg = ref_scene.small_network_config(os.path.join(work, "small.json")) if args.config == "small" else os.path.join(ROOT, "oracle", "_ref", "configs", "nerf", "base.json")
    cmd = [HARNESS, scene_dir + "/", net_cfg, dump, str(args.steps), "--pin-rays", str(args.rays), "--time-only", "--mesh", str(args.mesh)]
    t0 = time.time()
    log = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(os.path.join(args.out, "ref_harness_mesh_%s.log" % args.config), "w").write(log.stdout)
    if log.returncode != 0:
        print(log.stdout[-3000:]); raise SystemExit("ref_harness failed rc=%d" % log.returncode)
    meta = ref_scene.read_meta(os.path.join(dump, "meta.txt"))
    res = int(meta["mesh_res"])
    mn = tuple(float(x) for x in meta["mesh_aabb_min"].split()); mx = tuple(float(x) for x in meta["mesh_aabb_max"].split())
    scale = float(meta["dataset_scale"]); off = tuple(float(x) for x in meta["dataset_offset"].split())
    n2w_s = float(meta["n2w_s"]); n2w_t = tuple(float(x) for x in meta["n2w_t"].split()); from_na = bool(int(meta["from_na"]))

    def rd(name, dt):
        return np.fromfile(os.path.join(dump, name), dt)

    D = rd("mesh_density.bin", np.float32).reshape(res, res, res)
    rV = rd("mesh_verts.bin", np.float32).reshape(-1, 3); rN = rd("mesh_normals.bin", np.float32).reshape(-1, 3)
    rC = rd("mesh_colors.bin", np.float32).reshape(-1, 3); rI = rd("mesh_indices.bin", np.uint32)
    obj_ref = open(os.path.join(dump, "ref_mesh.obj"), "rb").read(); ply_ref = open(os.path.join(dump, "ref_mesh.ply"), "rb").read()
    summary = {"config": args.config, "ref_meta": meta, "ref_seconds": round(time.time() - t0, 1), "mesh_res": res,
               "ref_verts_padded": int(rV.shape[0]), "ref_indices": int(rI.size), "obj_bytes": len(obj_ref), "ply_bytes": len(ply_ref)}

    # --- marching cubes on the reference's lattice ---
    oV, oN, oI, onv = ob.marching_cubes(D, mn, mx, 0.0)
    n_used = int(rI.max()) + 1 if rI.size else 0
    ref_sorted = sort_rows(vertex_rows(rV[:n_used]))
    ref_tris = canonical_triangles(rV, rI)
    # per-vertex attributes are matched through the vertex positions; positions that occur more than once (an SDF value exactly
    # on the threshold puts the vertices of all edges at that lattice point in the same place) cannot be paired and are left out
    def unique_sorted(V, n):
        rows = vertex_rows(V[:n]); o = np.lexsort(rows.T[::-1]); r = rows[o]
        if n < 2:
            return o, np.ones(n, bool)
        same_prev = np.concatenate([[False], np.all(r[1:] == r[:-1], axis=1)]); same_next = np.concatenate([same_prev[1:], [False]])
        return o, ~(same_prev | same_next)
    r_ord, r_uni = unique_sorted(rV, n_used)
    unique_frac = float(r_uni.mean()) if n_used else 1.0

    def cmp_mesh(V, N, I, nv, C=None):
        out = {"n_verts": int(nv), "n_verts_equal": bool(nv == n_used), "padded_equal": bool(V.shape[0] == rV.shape[0]), "n_indices_equal": bool(I.size == rI.size)}
        out["vertex_set_bit_exact"] = bool(nv == n_used and np.array_equal(sort_rows(vertex_rows(V[:nv])), ref_sorted))
        out["triangle_set_bit_exact"] = bool(I.size == rI.size and np.array_equal(canonical_triangles(V, I), ref_tris))
        if out["vertex_set_bit_exact"]:
            o_ord, o_uni = unique_sorted(V, nv)                       # same sorted positions => same uniqueness mask
            Ns, Nr = np.asarray(N[:nv])[o_ord][o_uni], rN[:n_used][r_ord][r_uni]
            den = np.maximum(np.linalg.norm(Nr, axis=1), 1e-30)
            out["normal_max_rel"] = float((np.linalg.norm(Ns - Nr, axis=1) / den).max())
            if C is not None:
                out["colors_max_abs"] = float(np.abs(np.asarray(C[:nv])[o_ord][o_uni] - rC[:n_used][r_ord][r_uni]).max())
        return out

    summary["oracle_mc_vs_ref"] = cmp_mesh(oV, oN, oI, onv)
    summary["vertex_positions_unique_fraction"] = unique_frac

    # --- text writers on the reference's own arrays ---
    tmp = os.path.join(work, "txt"); os.makedirs(tmp, exist_ok=True)
    ob.save_mesh(os.path.join(tmp, "o.obj"), rV, rN, rC, rI, scale, off, n2w_s, n2w_t, from_na)
    ob.save_mesh(os.path.join(tmp, "o.ply"), rV, rN, rC, rI, scale, off, n2w_s, n2w_t, from_na)
    summary["oracle_writer"] = {"obj_identical": open(os.path.join(tmp, "o.obj"), "rb").read() == obj_ref, "ply_identical": open(os.path.join(tmp, "o.ply"), "rb").read() == ply_ref}

    # --- N1 on the CPU: oracle SDF at the lattice points with the reference's inference parameters ---
    p_inf = h2f(rd("mesh_params_inference_fp16.bin", np.uint16))
    o = Oracle(threads=os.cpu_count() or 4, **cfgd)
    o.set_params(p_inf)
    vl = o.valid_level(int(meta["mesh_training_step"]))
    rs = np.random.RandomState(0); pick = np.sort(rs.choice(res ** 3, min(res ** 3, 60000), replace=False))
    iz, iy, ix = np.unravel_index(pick, (res, res, res))
    idx = np.stack([ix, iy, iz], 1).astype(np.float32)
    pos = ((idx * (np.float32(1.0) / np.float32(res))) * (np.array(mx, np.float32) - np.array(mn, np.float32)) + np.array(mn, np.float32)).astype(np.float32)
    s_or, _ = o.eval_sdf(pos, vl)
    summary["sdf_lattice"] = {"valid_level": int(vl), "oracle_vs_ref_rel": rel_err(s_or, D.ravel()[pick]), "oracle_vs_ref_max_abs": float(np.abs(s_or - D.ravel()[pick]).max())}

    if not args.no_cuda:
        import torch
        pkg = rnb_loader.load_package()
        t = pkg.Testbed(product_config(pkg, cfgd, rays_per_batch=args.rays, pin_rays_per_batch=1))
        t.set_params(p_inf)
        t.set_train_state(int(meta["mesh_training_step"]), args.rays, 0, 0)
        dD = torch.from_numpy(D.ravel().copy()).cuda()
        info = t.marching_cubes_from_density(dD.data_ptr(), (res, res, res), mn, mx, 0.0, with_colors=True, use_ema=False)
        m = t.mesh_download()
        summary["cuda_mc_vs_ref"] = cmp_mesh(m["V"], m["N"], m["F"].ravel(), info["n_verts"], m["C"])
        summary["cuda_stage_ms"] = info["stage_ms"]
        summary["cuda_mc_vs_oracle_bit_exact"] = bool(np.array_equal(m["V"], oV) and np.array_equal(m["N"], oN) and np.array_equal(m["F"].ravel(), oI))
        # GPU text formatter on the reference's arrays
        dv, dn, dc, di = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (rV, rN, rC, rI.view(np.int32)))
        for ext, refb in (("obj", obj_ref), ("ply", ply_ref)):
            pth = os.path.join(tmp, "g." + ext)
            t0 = time.time()
            nb = pkg.save_mesh_device(pth, dv.data_ptr(), dn.data_ptr(), dc.data_ptr(), di.data_ptr(), rV.shape[0], rI.size, scale, off, n2w_s, n2w_t, from_na)
            got = open(pth, "rb").read()
            summary.setdefault("cuda_writer", {})[ext + "_identical"] = bool(got == refb and nb == len(refb))
            summary["cuda_writer"][ext + "_seconds"] = round(time.time() - t0, 4)
        # N1: our lattice sweep with the same parameters
        sd = torch.empty(res ** 3, device="cuda")
        t.sdf_on_grid_device((res, res, res), mn, mx, sd.data_ptr(), use_ema=False)
        torch.cuda.synchronize()
        S = sd.cpu().numpy()
        summary["sdf_lattice"].update(cuda_vs_ref_rel=rel_err(S, D.ravel()), cuda_vs_ref_max_abs=float(np.abs(S - D.ravel()).max()),
                                      sign_mismatch=int(np.count_nonzero((S > 0) != (D.ravel() > 0))))
        # whole pipeline timing on our side at the same resolution
        torch.cuda.synchronize(); t0 = time.time()
        t.compute_and_save_marching_cubes_mesh(os.path.join(tmp, "ours.obj"), args.mesh, mn, mx, 0.0, use_ema=False, nerf_scale=scale, nerf_offset=off, n2w_s=n2w_s, n2w_t=n2w_t, from_na=from_na)
        summary["cuda_pipeline_seconds"] = round(time.time() - t0, 4)
        summary["ref_pipeline_seconds"] = float(meta["mesh_seconds"])
    print(json.dumps(summary))
    json.dump(summary, open(os.path.join(args.out, "summary_mesh_%s_%d.json" % (args.config, res)), "w"), indent=1)
    if args.golden:
        np.savez_compressed(os.path.join(args.out, "golden_mesh_%s.npz" % args.golden), density=D.astype(np.float32), aabb=np.array([mn, mx], np.float32), verts=rV, normals=rN, colors=rC,
                            indices=rI, obj=np.frombuffer(obj_ref, np.uint8), ply=np.frombuffer(ply_ref, np.uint8),
                            writer=np.array([scale, *off, n2w_s, *n2w_t, float(from_na)], np.float64))


if __name__ == "__main__":
    main()
