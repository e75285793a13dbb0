#!/usr/bin/env python
"""Time the UNMODIFIED reference CUDA path (oracle/_ref/bin/ref_harness --time-only: Testbed::train of the reference, sm_100
SASS built by oracle/Makefile.ref) on this box, on the same synthetic scene bench.py uses.  TEST / BENCH INFRASTRUCTURE.

    python tools/ref_time.py --steps 600 [--pin-rays 4096] --out gpurun_out/ref_time.json
"""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_scene   # noqa: E402
import bench       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--pin-rays", type=int, default=4096)
    ap.add_argument("--views", type=int, default=96); ap.add_argument("--width", type=int, default=1600); ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--out", default="gpurun_out/ref_time.json")
    ap.add_argument("--work", default="/tmp/ref_time")
    args = ap.parse_args()
    scene_dir = os.path.join(args.work, "scene"); dump = os.path.join(args.work, "dump"); os.makedirs(dump, exist_ok=True)
    t0 = time.time()
    views, _ = bench.build_views(args.views, args.width, args.height, False)
    ref_scene.write_scene(scene_dir, views, workers=min(16, os.cpu_count() or 4))
    t_scene = time.time() - t0
    res = {"scene_seconds": round(t_scene, 1), "runs": []}
    for pin in ([args.pin_rays, 0] if args.pin_rays else [0]):
        cmd = [os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness"), scene_dir + "/", os.path.join(ROOT, "oracle", "_ref", "configs", "nerf", "base.json"), dump, str(args.steps), "--no-albedo", "--time-only"]
        if pin:
            cmd += ["--pin-rays", str(pin)]
        t1 = time.time()
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        meta = ref_scene.read_meta(os.path.join(dump, "meta.txt")) if p.returncode == 0 else {}
        res["runs"].append({"pin_rays": pin, "rc": p.returncode, "wall_s": round(time.time() - t1, 1), "steps": args.steps, "timed_steps": meta.get("timed_steps"), "timed_ms": meta.get("timed_ms"),
                            "timed_rays": meta.get("timed_rays"), "rays_per_second": meta.get("rays_per_second"), "log_tail": p.stdout[-1500:]})
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res)[:3000])


if __name__ == "__main__":
    main()
