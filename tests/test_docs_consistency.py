"""DESIGN.md / INTEGRATION.md / README.md cite files of this repository as evidence (profiles, tests, tools, sources): every cited path
must exist, and the headline numbers quoted in the README must be the ones in the committed bench record."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ("DESIGN.md", "INTEGRATION.md", "README.md")
PREFIXES = ("profiles/", "tests/", "tools/", "oracle/", "shim/", "include/", "rnb-neus2_b200/")
GENERATED = ("oracle/_ref", "rnb-neus2_b200/librnb_b200.so", "oracle/librnb_oracle.so")          # built, git-ignored


def cited_paths(text):
    out = set()
    for m in re.finditer(r"`([^`\s]+)`", text):
        p = m.group(1).rstrip(".,;:)")
        p = p.split("::")[0]
        if p.startswith(PREFIXES) and not any(p.startswith(g) for g in GENERATED):
            out.add(p)
    return out


def test_cited_files_exist():
    missing = []
    for doc in DOCS:
        for p in sorted(cited_paths(open(os.path.join(ROOT, doc)).read())):
            full = os.path.join(ROOT, p)
            if any(ch in p for ch in "*<>{}…"):
                if "*" in p and not glob.glob(full):
                    missing.append((doc, p))
                continue
            if not os.path.exists(full):
                missing.append((doc, p))
    assert not missing, missing


def test_readme_headline_matches_the_bench_record():
    rec = json.load(open(os.path.join(ROOT, "profiles", "r02_final_bench.json")))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02_final_bench_reference_arm.json")))
    readme = open(os.path.join(ROOT, "README.md")).read()
    assert "%.2f M rays/s" % (rec["value"] / 1e6) in readme
    assert "%.2f M end to end" % (rec["e2e"]["value"] / 1e6) in readme
    assert "%.3f ms/step" % rec["ms_per_step"] in readme
    assert "%.2f M rays/s for the reference" % (ref["value"] / 1e6) in readme
    assert rec["roofline"]["kernel"] == "backward" and 0 < rec["roofline"]["frac"] < 1
    assert rec["cpu_baseline"]["kind"] == "port" and rec["gpu_launches"] > 0 and rec["clocks"]["reasons"] == []
    assert "normals+albedo" in rec["config"]["workload"] and rec["config"]["live_hash_levels"] == 14 and rec["config"]["workload"] == ref["config"]["workload"]
    assert set(rec["records"]) >= {"normals", "supernormal", "mesh_1024", "adaptive_controller"}


def test_readme_multi_gpu_numbers_match_the_records():
    recs = [json.load(open(os.path.join(ROOT, "profiles", "r02_final_bench_n%d.json" % n))) for n in (2, 4, 8)]
    readme = open(os.path.join(ROOT, "README.md")).read()
    assert "%.1f M / %.1f M / %.1f M rays/s on 2 / 4 / 8 GPUs" % tuple(r["value"] / 1e6 for r in recs) in readme
    for n, r in zip((2, 4, 8), recs):
        assert r["n_gpus"] == n and r["scaling"] == "weak" and r["config"]["nccl"]["one_sample_order"] and r["config"]["nccl"]["sharded"]
    old = [json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n%d_per_rank_rule.json" % n))) for n in (2, 4, 8)]
    assert "%.1f M / %.1f M / %.1f M, 90 / 90 / 92 %%" % tuple(r["value"] / 1e6 for r in old) in readme
