#!/usr/bin/env python
"""Applies the six one-line hooks of shim/rnb_testbed_shim.h to a SCRATCH COPY of the reference's src/testbed.cu and
src/testbed_nerf.cu (test infrastructure: proves that the binding documented in INTEGRATION.md compiles and links against the
reference's real code).  Usage: shim_patch.py <reference root> <output dir>.  The reference tree is only read; the patched copies land
under oracle/_ref/ (git-ignored) and are never committed.  Anchors are function signatures / single statements, located by text."""
import os
import sys

HOOK_INCLUDE = '#include <rnb_testbed_shim.h>'


def guard(code):
    return ["#ifdef NGP_USE_RNB_B200", code, "#endif"]


def find(lines, needle, start=0):
    for i in range(start, len(lines)):
        if needle in lines[i]:
            return i
    raise SystemExit("shim_patch: anchor not found: %r" % needle)


def closing_brace_before(lines, i):
    for j in range(i - 1, -1, -1):
        if lines[j].rstrip() == "}":
            return j
    raise SystemExit("shim_patch: no closing brace before line %d" % i)


def patch_testbed(lines):
    out = list(lines)
    # processed bottom-up so that earlier line numbers stay valid
    edits = []
    i = find(out, "void Testbed::load_snapshot_incremental(")
    edits.append((closing_brace_before(out, i), guard("\trnb_shim::pull_state(*this);")))
    i = find(out, "void Testbed::save_snapshot(const std::string& filepath_string, bool include_optimizer_state) {")
    edits.append((i + 1, guard("\trnb_shim::push_state(*this);")))
    i = find(out, "void Testbed::train(uint32_t batch_size) {")
    j = find(out, "m_nerf_network->m_training_step = m_training_step;", i)
    edits.append((j, guard("\tif (m_testbed_mode == ETestbedMode::Nerf && rnb_shim::train(*this)) { update_loss_graph(); return; }")))
    i = find(out, "void Testbed::reset_network_incremental() {")
    edits.append((closing_brace_before(out, i), guard("\trnb_shim::on_reset_network(*this);")))
    i = find(out, "void Testbed::compute_and_save_marching_cubes_mesh(")
    j = find(out, "marching_cubes(res3d, aabb, thresh);", i)
    edits.append((j, guard("\tif (m_testbed_mode == ETestbedMode::Nerf && rnb_shim::compute_and_save_mesh(*this, filename, res3d, aabb, thresh, unwrap_it)) return;")))
    i = find(out, "#include <neural-graphics-primitives/testbed.h>")
    edits.append((i + 1, [HOOK_INCLUDE]))
    for at, block in sorted(edits, key=lambda e: -e[0]):
        out[at:at] = block
    return out


def patch_testbed_nerf(lines):
    out = list(lines)
    edits = []
    i = find(out, "void Testbed::load_nerf(uint32_t frame_time_idx, bool is_downsample) {")
    edits.append((closing_brace_before(out, i), guard("\trnb_shim::on_dataset(*this);")))
    i = find(out, "#include <neural-graphics-primitives/testbed.h>")
    edits.append((i + 1, [HOOK_INCLUDE]))
    for at, block in sorted(edits, key=lambda e: -e[0]):
        out[at:at] = block
    return out


def main():
    ref, dst = sys.argv[1], sys.argv[2]
    os.makedirs(dst, exist_ok=True)
    for name, fn in (("testbed.cu", patch_testbed), ("testbed_nerf.cu", patch_testbed_nerf)):
        lines = open(os.path.join(ref, "src", name)).read().split("\n")
        patched = fn(lines)
        open(os.path.join(dst, name), "w").write("\n".join(patched))
        print("shim_patch: %s: %d hook lines added" % (name, len(patched) - len(lines)))


if __name__ == "__main__":
    main()
