"""CPU: the C-ABI library loads and exports every declared symbol; host-side structure build; sharding logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_npz
from graphite_b200 import binding, synthetic
from graphite_b200.distributed import partition_by_point, point_ranges


def declared_symbols(header="graphite_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef[^;]*\(\*gb_[a-z0-9_]+\)[^;]*;", "", text)  # callback typedefs are not entry points
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    L = binding.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"libgraphite_b200.so does not export {n}"
    assert sorted(binding.SYMBOLS) == names, "binding.SYMBOLS must list exactly the header's entry points"
    assert L.gb_version() == 100
    # the generic factor-graph ABI (include/graphite_b200_graph.h)
    from graphite_b200 import graph
    gnames = declared_symbols("graphite_b200_graph.h")
    assert len(gnames) >= 20
    for n in gnames:
        assert hasattr(L, n), f"libgraphite_b200.so does not export {n}"
    assert sorted(graph.SYMBOLS) == gnames, "graph.SYMBOLS must list exactly the header's entry points"


def test_no_gpu_means_loud_failure(built):
    """No CPU fallback: without a device the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(binding.GraphiteB200Error):
        binding.Context(0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under graphite_b200/ may import, link or execute it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "graphite_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    out = subprocess.run(["ldd", binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_structure_matches_reference_golden(built):
    prob = synthetic.make_named("ladybug-49")
    z = golden_npz("ladybug-49__pcg-schur__FP64-FP64.npz")
    s = binding.host_structure(prob.cam_idx, prob.pt_idx, prob.n_cams, prob.n_pts)
    cp, ri, off = s["hessian"]
    assert np.array_equal(cp, z["H_colptr"]) and np.array_equal(ri, z["H_rowidx"]) and np.array_equal(off, z["H_offsets"])


@pytest.mark.parametrize("case", ["schur-fixture", "ladybug-49"])
def test_schur_structure_matches_reference_golden(built, case):
    """SchurComplement::build_structure: d_col_pointers / d_row_indices of the reference, bit-exact."""
    prob = synthetic.schur_fixture() if case == "schur-fixture" else synthetic.make_named(case)
    z = golden_npz(f"{case}__pcg-schur__FP64-FP64.npz")
    hs = binding.host_structure(prob.cam_idx, prob.pt_idx, prob.n_cams, prob.n_pts)
    cp, ri = hs["schur"]
    assert np.array_equal(cp, z["S_colptr"]) and np.array_equal(ri, z["S_rowidx"])
    # a sparser pattern: cameras that share no point have no block
    ci = np.array([0, 1, 1, 2, 2, 3], dtype=np.int32)
    pi = np.array([0, 0, 1, 1, 2, 2], dtype=np.int32)
    cp, ri = binding.host_structure(ci, pi, 4, 3)["schur"]
    assert cp.tolist() == [0, 1, 3, 5, 7] and ri.tolist() == [0, 0, 1, 1, 2, 2, 3]


@pytest.mark.parametrize("tile,cap,st_obs", [(0, 0, 0), (64, 0, 0), (32, 40, 500), (0, 30, 2000)])
def test_tiles_ranks_segments_and_super_tiles(built, tile, cap, st_obs):
    prob = synthetic.make_named("ladybug-49")
    s = binding.host_structure(prob.cam_idx, prob.pt_idx, prob.n_cams, prob.n_pts, tile, cap, st_obs)
    fill, capv = tile or 256, cap or 192
    to, tp, tm = s["tile_obs"], s["tile_pt"], s["tmeta"]
    m, nt = prob.n_obs, len(s["tile_obs"]) - 1
    assert to[0] == 0 and to[-1] == m and tp[-1] == prob.n_pts
    sizes = np.diff(to)
    assert sizes.min() > 0 and sizes.max() <= fill
    assert np.array_equal(s["pptr"][tp], to), "tiles hold whole points"
    assert s["info"]["storage_slots"] == nt * 256 and len(s["ometa"]) == nt * 256
    # storage slots: tile-padded; inside a tile the slots are in (camera, observation) order
    slot = s["slot_of_obs"]
    assert sorted(slot.tolist()) == sorted(np.concatenate([k * 256 + np.arange(sizes[k]) for k in range(nt)]).tolist())
    st_tile, st_row, row_cam = s["st_tile"], s["st_row"], s["row_cam"]
    assert st_tile[0] == 0 and st_tile[-1] == nt and np.all(np.diff(st_tile) > 0)
    assert np.diff(st_row).max() <= capv
    nseg = 0
    for sidx in range(len(st_tile) - 1):
        rows = row_cam[st_row[sidx]:st_row[sidx + 1]]
        assert np.all(np.diff(rows) > 0), "rows of a super-tile: its distinct cameras, ascending"
        o0, o1 = to[st_tile[sidx]], to[st_tile[sidx + 1]]
        assert np.array_equal(np.unique(prob.cam_idx[o0:o1]), rows)
        for k in range(st_tile[sidx], st_tile[sidx + 1]):
            p0, n, npt, ns, seg_off, pt_off, o0k, _ = tm[k]
            assert (p0, n, npt, o0k) == (tp[k], sizes[k], tp[k + 1] - tp[k], to[k])
            om = s["ometa"][k * 256:(k + 1) * 256]
            cslot, prank, ptl = om >> 16, (om >> 8) & 0xff, om & 0xff
            assert sorted(prank.tolist()) == list(range(256)), "point-order positions are a permutation of the 256 slots"
            # slot u of the tile holds the observation at point-order position prank[u]
            obs_of_slot = to[k] + prank[:n].astype(int)
            assert np.array_equal(slot[obs_of_slot], k * 256 + np.arange(n))
            assert np.array_equal(s["rank"][obs_of_slot], np.arange(n))
            cams_sorted = prob.cam_idx[obs_of_slot]
            assert np.all(np.diff(cams_sorted) >= 0), "slots are sorted by camera"
            same = np.diff(cams_sorted) == 0
            assert np.all(np.diff(prank[:n].astype(int))[same] > 0), "and by observation inside a camera"
            assert np.array_equal(rows[cslot[:n]], cams_sorted)
            assert np.array_equal(ptl[:n] + p0, prob.pt_idx[obs_of_slot])
            seg = s["seg_tab"][seg_off:seg_off + ns + 1]
            begins, slots = seg >> 16, seg & 0xffff
            heads = np.flatnonzero(np.diff(cams_sorted, prepend=-1) != 0)
            assert np.array_equal(begins[:ns], heads) and begins[ns] == n
            assert np.array_equal(rows[slots[:ns]], cams_sorted[heads])
            assert seg_off % 4 == 0 and pt_off % 8 == 0
            pt = s["pt_tab"][pt_off:pt_off + npt + 1]
            assert np.array_equal(pt, s["pptr"][p0:p0 + npt + 1] - to[k])
            nseg += ns
    assert nseg == s["info"]["n_camera_segments"]
    # camera -> partial rows CSR lists every row once, ascending super-tile order per camera
    lst, ptr = s["cam_row_list"], s["cam_row_ptr"]
    assert sorted(lst.tolist()) == list(range(len(row_cam)))
    for c in range(prob.n_cams):
        mine = lst[ptr[c]:ptr[c + 1]]
        assert np.all(row_cam[mine] == c) and np.all(np.diff(mine) > 0)


def test_unsorted_input_is_sorted_with_a_permutation(built):
    prob = synthetic.make_named("ladybug-49")
    rng = np.random.default_rng(3)
    perm = rng.permutation(prob.n_obs)
    s = binding.host_structure(prob.cam_idx[perm], prob.pt_idx[perm], prob.n_cams, prob.n_pts)
    assert np.array_equal(s["cam_idx"], prob.cam_idx) and np.array_equal(s["pt_idx"], prob.pt_idx)
    assert np.array_equal(perm[s["perm"]], np.arange(prob.n_obs))


def test_structure_rejects_what_it_cannot_handle(built):
    prob = synthetic.schur_fixture()
    with pytest.raises(binding.GraphiteB200Error, match="no observation"):
        binding.host_structure(prob.cam_idx, prob.pt_idx, prob.n_cams, prob.n_pts + 1)
    with pytest.raises(binding.GraphiteB200Error, match="duplicate"):
        binding.host_structure(np.array([0, 1, 1], dtype=np.int32), np.array([0, 0, 0], dtype=np.int32), 2, 1)
    with pytest.raises(binding.GraphiteB200Error, match="out of range"):
        binding.host_structure(np.array([0, 5], dtype=np.int32), np.array([0, 0], dtype=np.int32), 2, 1)
    # a track longer than the tile / the slot cap is not an error any more: it is cut into fragment tiles
    assert len(binding.host_structure(np.arange(40, dtype=np.int32), np.zeros(40, dtype=np.int32), 40, 1, 32)["frag_tile"]) == 2
    assert len(binding.host_structure(np.arange(40, dtype=np.int32), np.zeros(40, dtype=np.int32), 40, 1, 0, 16)["frag_tile"]) == 3
    with pytest.raises(binding.GraphiteB200Error, match="slot cap"):
        binding.host_structure(np.arange(40, dtype=np.int32), np.zeros(40, dtype=np.int32), 40, 1, 0, 500)


def test_point_partition_covers_everything(built):
    prob = synthetic.make_named("trafalgar-257")
    for n in (2, 4, 8):
        rng_ = point_ranges(prob.pt_idx, prob.n_pts, n)
        assert rng_[0][0] == 0 and rng_[-1][1] == prob.n_pts
        assert all(a[1] == b[0] for a, b in zip(rng_, rng_[1:]))
        parts = [partition_by_point(prob, n, r) for r in range(n)]
        assert sum(p.n_obs for p in parts) == prob.n_obs and sum(p.n_pts for p in parts) == prob.n_pts
        counts = np.array([p.n_obs for p in parts])
        assert counts.max() / counts.mean() < 1.02  # balanced by observations
        for p in parts:
            assert p.n_cams == prob.n_cams and p.pt_idx.min() == 0 and p.pt_idx.max() == p.n_pts - 1


def test_point_partition_keeps_long_tracks_whole(built):
    """Ranks own contiguous point ranges: a long track (cut into fragment tiles inside its rank) is never split over ranks."""
    prob = synthetic.make_named("long-tracks")
    track = np.bincount(prob.pt_idx)
    for n in (2, 4, 8):
        parts = [partition_by_point(prob, n, r) for r in range(n)]
        assert sum(p.n_obs for p in parts) == prob.n_obs and sum(p.n_pts for p in parts) == prob.n_pts
        local_tracks = np.concatenate([np.bincount(p.pt_idx) for p in parts])
        assert np.array_equal(local_tracks, track)
        for p in parts:  # a rank's share builds (cameras without a local observation are legal in a partition)
            ci, pi = p.cam_idx, p.pt_idx
            assert pi.min() == 0 and pi.max() == p.n_pts - 1 and np.all(np.diff(pi.astype(np.int64) * p.n_cams + ci) > 0)


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from graphite_b200 import synthetic
from graphite_b200.distributed import partition_by_point
from oracle.binding import Oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
prob = synthetic.make_named("ladybug-49")
local = partition_by_point(prob, world, rank)
o = Oracle(local)
chi2, sc, b = o.linearize()
# what the C library all-reduces: cost and the camera part of the (unscaled) gradient / Hessian diagonal
jc, jp = Oracle(local).jacobians()
r, _ = Oracle(local).residuals()
diag = np.zeros(9 * prob.n_cams); g = np.zeros(9 * prob.n_cams)
J = jc.reshape(-1, 9, 2)
np.add.at(diag.reshape(-1, 9), local.cam_idx, (J ** 2).sum(2))
np.add.at(g.reshape(-1, 9), local.cam_idx, -(J * r[:, None, :]).sum(2))
t = torch.from_numpy(np.concatenate([[chi2], diag, g]))
dist.all_reduce(t)
if rank == 0:
    full = Oracle(prob)
    fchi2, fsc, fb = full.linearize()
    tot = t.numpy()
    assert abs(tot[0] - fchi2) <= 1e-12 * fchi2
    n = 9 * prob.n_cams
    scale = 1.0 / (np.finfo(float).eps + np.sqrt(tot[1:1 + n]))
    assert np.allclose(scale, fsc[:n], rtol=1e-12)
    assert np.allclose(scale * tot[1 + n:], fb[:n], rtol=1e-9, atol=1e-9 * np.abs(fb[:n]).max())
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_two_rank_sharding_over_gloo(built, tmp_path):
    """world_size 2 on CPU: the point partition + all-reduce of camera-sized vectors reproduces the full problem."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "GLOO_OK" in res.stdout


def test_bench_reference_arm_contract(built):
    """`bench.py --impl reference` (the CPU arm: oracle port on the host cores) prints one JSON line with the contract's keys."""
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ladybug-49",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "LM it/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["cams"] == 49 and "workload" in d["config"] and "model" not in d["config"]


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/graphite_b200.h must compile as C99 (-pedantic) and as C++17."""
    src = tmp_path / "t.c"
    src.write_text('#include "graphite_b200.h"\n#include "graphite_b200_graph.h"\n'
                   'int main(void) { gb_lm_options o = {0}; gb_pcg_options p = {10, 1.0, 5.0, GB_SOLVER_PCG_SCHUR, 0};\n'
                   '  gb_factor_eval e = {0}; gb_graph_eval g = {0}; gb_vertex_set_desc v = {0}; gb_factor_set_desc f = {0};\n'
                   '  (void)o; (void)p; (void)e; (void)g; (void)v; (void)f; return gb_version() == 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])


def test_roofline_byte_model_matches_the_survey_figures():
    """bench.py's algorithmic-byte model reproduces SURVEY.md section 8(d): Venice FP64, 10 PCG iterations, accepted step."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    # bench.py redirects fd 1 at import time (native banners off stdout): keep the test's stdout intact
    saved = os.dup(1)
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    total = mod.lm_iteration_bytes(1778, 993923, 5001946, 8, 8, 10, 1.0)
    assert abs(total / 1e9 - 15.8) < 0.1                      # "Venice: 15.8 GB implicit" per LM iteration
    product, k4 = mod.algorithmic_bytes(1778, 993923, 5001946, 221455, 19788, 8, 8)
    assert abs(k4 / 1e9 - 1.17) < 0.01                         # "1.17 GB implicit" per PCG iteration
    assert product < k4                                        # the stored-factor layout moves fewer bytes than the E blocks


def test_bal_text_round_trip(tmp_path):
    """BAL text format as parsed by examples/bal.cu:63-147: `<n_cams> <n_pts> <n_obs>`, one `cam pt x y` line per
    observation, then 9 values per camera and 3 per point.  Written with 17 significant digits: the round trip is exact."""
    prob = synthetic.make_bal(12, 300, 1100, seed=3, name="tiny")
    path = str(tmp_path / "tiny.txt")
    synthetic.write_bal_text(prob, path)
    head = open(path).readline().split()
    assert [int(v) for v in head] == [12, 300, 1100]
    back = synthetic.read_bal_text(path)
    assert np.array_equal(back.cam_idx, prob.cam_idx) and np.array_equal(back.pt_idx, prob.pt_idx)
    assert np.array_equal(back.obs, prob.obs) and np.array_equal(back.cams, prob.cams) and np.array_equal(back.pts, prob.pts)
    # the parsed problem builds the same structure (ids as in bal.cu: camera id = index, point id = n_cams + index)
    a = binding.host_structure(prob.cam_idx, prob.pt_idx, prob.n_cams, prob.n_pts)
    b = binding.host_structure(back.cam_idx, back.pt_idx, back.n_cams, back.n_pts)
    for x, y in zip(a["hessian"], b["hessian"]):
        assert np.array_equal(x, y)


def test_long_tracks_are_cut_into_fragment_tiles():
    """A point observed by more cameras than one tile holds (192 rows per super-tile; real BAL landmarks have such
    tracks) is cut into FRAGMENT tiles: consecutive tiles that hold nothing but a part of that point's observations."""
    nc, npts = 200, 3
    cam = np.concatenate([np.arange(193), [0, 1], [2, 3]]).astype(np.int32)
    pt = np.concatenate([np.zeros(193), [1, 1], [2, 2]]).astype(np.int32)
    s = binding.host_structure(cam, pt, 193, npts)
    assert s["info"]["max_track"] == 193
    assert s["frag_tile"].tolist() == [0, 1] and s["hv_pt"].tolist() == [0] and s["hv_ptr"].tolist() == [0, 2]
    assert s["tile_obs"].tolist() == [0, 96, 193, 197] and s["tile_pt"].tolist() == [0, 0, 1, 3]
    # 192 fits in one tile (the remaining cameras are observed by the other points, so no vertex is unused)
    cam2 = np.concatenate([np.arange(192), np.arange(192, 200), [0, 1]]).astype(np.int32)
    pt2 = np.concatenate([np.zeros(192), np.ones(8), [2, 2]]).astype(np.int32)
    s = binding.host_structure(cam2, pt2, nc, npts)
    assert s["info"]["max_track"] == 192 and len(s["frag_tile"]) == 0

    many = synthetic.make_long_tracks(n_cams=420, n_pts=3000, n_obs=20000, tracks=tuple(200 + (i * 37) % 221 for i in range(300)), seed=3)
    for kw, prob in (({}, synthetic.make_named("long-tracks")), (dict(tile_size=24, slot_cap=40), synthetic.make_named("long-tracks")),
                     (dict(tile_size=64), synthetic.make_named("long-tracks")), ({}, many)):
        s = binding.host_structure(prob.cam_idx, prob.pt_idx, prob.n_cams, prob.n_pts, **kw)
        cap = min(kw.get("tile_size", 256), kw.get("slot_cap", 192))
        to, tp, tm, pptr = s["tile_obs"], s["tile_pt"], s["tmeta"], s["pptr"]
        nt = len(to) - 1
        track = np.diff(pptr)
        heavy = np.flatnonzero(track > cap)
        assert np.array_equal(s["hv_pt"], heavy) and (cap > 40 or len(heavy) > 4)
        assert to[0] == 0 and to[-1] == prob.n_obs and np.all(np.diff(to) > 0) and np.diff(to).max() <= kw.get("tile_size", 256)
        frag = np.zeros(nt, bool)
        frag[s["frag_tile"]] = True
        for h, p in enumerate(heavy):
            tiles = s["frag_tile"][s["hv_ptr"][h]:s["hv_ptr"][h + 1]]
            assert np.array_equal(tiles, np.arange(tiles[0], tiles[-1] + 1)), "fragments of a point are consecutive tiles"
            assert to[tiles[0]] == pptr[p] and to[tiles[-1] + 1] == pptr[p + 1], "and cover exactly its observations"
            assert len(tiles) == -(-track[p] // cap) and np.diff(to[tiles[0]:tiles[-1] + 2]).max() <= cap
            for j, k in enumerate(tiles):
                p0, n, npt, ns, seg_off, pt_off, o0, code = tm[k]
                assert (p0, npt, n, ns, o0) == (p, 1, to[k + 1] - to[k], to[k + 1] - to[k], to[k])
                assert code & 0x3fffffff == s["hv_ptr"][h] + j + 1 and bool(code >> 30) == (j == 0)
                assert s["pt_tab"][pt_off:pt_off + 2].tolist() == [0, n]
        for k in np.flatnonzero(~frag):  # every other tile holds whole points
            p0, n, npt, ns, seg_off, pt_off, o0, code = tm[k]
            assert code == 0 and (p0, o0) == (tp[k], to[k]) and to[k] == pptr[p0] and to[k + 1] == pptr[p0 + npt]
            assert np.array_equal(s["pt_tab"][pt_off:pt_off + npt + 1], pptr[p0:p0 + npt + 1] - to[k])
        # slots: a permutation of the tile's positions, sorted by camera; rows of a super-tile = its distinct cameras
        slot = s["slot_of_obs"]
        assert sorted(slot.tolist()) == sorted(np.concatenate([k * 256 + np.arange(to[k + 1] - to[k]) for k in range(nt)]).tolist())
        st_tile, st_row, row_cam = s["st_tile"], s["st_row"], s["row_cam"]
        assert np.diff(st_row).max() <= kw.get("slot_cap", 192)
        for sidx in range(len(st_tile) - 1):
            rows = row_cam[st_row[sidx]:st_row[sidx + 1]]
            assert np.array_equal(np.unique(prob.cam_idx[to[st_tile[sidx]]:to[st_tile[sidx + 1]]]), rows)

def _check_structure(cam, pt, nc, npts, s, tile_size, slot_cap):
    """Invariants the kernels rely on (csrc/structure.hpp): tiles and fragments partition the sorted observations, slots are a
    camera-sorted permutation of the tile, segment / point tables describe them, super-tile rows are its distinct cameras."""
    fill, capv = tile_size or 256, slot_cap or 192
    cap = min(fill, capv)
    to, tp, tm, pptr = s["tile_obs"], s["tile_pt"], s["tmeta"], s["pptr"]
    m, nt = len(cam), len(to) - 1
    assert np.array_equal(s["cam_idx"], cam) and np.array_equal(s["pt_idx"], pt)
    assert to[0] == 0 and to[-1] == m and np.all(np.diff(to) > 0) and np.diff(to).max() <= fill
    track = np.diff(pptr)
    heavy = np.flatnonzero(track > cap)
    assert np.array_equal(s["hv_pt"], heavy) and s["info"]["max_track"] == track.max()
    frag = np.zeros(nt, bool)
    frag[s["frag_tile"]] = True
    assert len(s["hv_ptr"]) == len(heavy) + 1 and s["hv_ptr"][-1] == len(s["frag_tile"])
    for h, p in enumerate(heavy):
        tiles = s["frag_tile"][s["hv_ptr"][h]:s["hv_ptr"][h + 1]]
        assert np.array_equal(tiles, np.arange(tiles[0], tiles[-1] + 1))
        assert to[tiles[0]] == pptr[p] and to[tiles[-1] + 1] == pptr[p + 1] and len(tiles) == -(-track[p] // cap)
        for j, k in enumerate(tiles):
            assert tm[k][0] == p and tm[k][2] == 1 and tm[k][7] & 0x3fffffff == s["hv_ptr"][h] + j + 1 and bool(tm[k][7] >> 30) == (j == 0)
    slot = s["slot_of_obs"]
    assert sorted(slot.tolist()) == sorted(np.concatenate([k * 256 + np.arange(to[k + 1] - to[k]) for k in range(nt)]).tolist())
    st_tile, st_row, row_cam = s["st_tile"], s["st_row"], s["row_cam"]
    assert st_tile[0] == 0 and st_tile[-1] == nt and np.all(np.diff(st_tile) > 0) and np.diff(st_row).max() <= capv
    covered = 0
    for sidx in range(len(st_tile) - 1):
        rows = row_cam[st_row[sidx]:st_row[sidx + 1]]
        assert np.array_equal(np.unique(cam[to[st_tile[sidx]]:to[st_tile[sidx + 1]]]), rows)
        for k in range(st_tile[sidx], st_tile[sidx + 1]):
            p0, n, npt, ns, seg_off, pt_off, o0k, code = tm[k]
            assert (n, o0k) == (to[k + 1] - to[k], to[k]) and 1 <= npt <= 128 and (code != 0) == frag[k]
            covered += n
            om = s["ometa"][k * 256:(k + 1) * 256]
            cslot, prank, ptl = om >> 16, (om >> 8) & 0xff, om & 0xff
            assert sorted(prank.tolist()) == list(range(256))
            obs_of_slot = to[k] + prank[:n].astype(int)
            assert np.array_equal(slot[obs_of_slot], k * 256 + np.arange(n))
            cams_sorted = cam[obs_of_slot]
            assert np.all(np.diff(cams_sorted) >= 0) and np.array_equal(rows[cslot[:n]], cams_sorted)
            assert np.array_equal(ptl[:n] + p0, pt[obs_of_slot])
            seg = s["seg_tab"][seg_off:seg_off + ns + 1]
            heads = np.flatnonzero(np.diff(cams_sorted, prepend=-1) != 0)
            assert ns == len(heads) and np.array_equal(seg[:ns] >> 16, heads) and seg[ns] >> 16 == n
            assert np.array_equal(rows[seg[:ns] & 0xffff], cams_sorted[heads])
            ptab = s["pt_tab"][pt_off:pt_off + npt + 1]
            assert np.array_equal(ptab, np.clip(pptr[p0:p0 + npt + 1] - to[k], 0, n))
            if not frag[k]:
                assert to[k] == pptr[p0] and to[k + 1] == pptr[p0 + npt], "an ordinary tile holds whole points"
    assert covered == m
    # camera -> partial rows: every row once, ascending super-tile order inside a camera
    ptr, lst = s["cam_row_ptr"], s["cam_row_list"]
    assert ptr[0] == 0 and ptr[-1] == len(row_cam) and sorted(lst.tolist()) == list(range(len(row_cam)))
    for c in range(nc):
        mine = lst[ptr[c]:ptr[c + 1]]
        assert np.all(row_cam[mine] == c) and np.all(np.diff(mine) > 0)


def test_structure_invariants_on_random_problems(built):
    """Random small problems with random track lengths (many of them longer than the tile / slot cap in force) in random
    tilings: the cut logic (tiles of whole points, fragments of long tracks, super-tiles) keeps every invariant."""
    rng = np.random.default_rng(2026)
    n_heavy_cases = 0
    for trial in range(60):
        nc = int(rng.integers(3, 70))
        npts = int(rng.integers(2, 150))
        tile_size = int(rng.choice([0, 8, 13, 32, 100]))
        slot_cap = int(rng.choice([0, 4, 7, 16, 50]))
        cams, pts = [], []
        for p in range(npts):
            kind = rng.random()
            t = nc if kind < 0.08 else int(rng.integers(1, max(2, min(nc, 6)) + 1)) if kind < 0.8 else int(rng.integers(1, nc + 1))
            cams.append(np.sort(rng.choice(nc, size=t, replace=False)))
            pts.append(np.full(t, p))
        cam, pt = np.concatenate(cams).astype(np.int32), np.concatenate(pts).astype(np.int32)
        missing = np.setdiff1d(np.arange(nc), cam)  # every camera must be observed: give the missing ones to the last point
        if missing.size:
            keep = pt != npts - 1
            last = np.union1d(cam[~keep], missing)
            cam = np.concatenate([cam[keep], last]).astype(np.int32)
            pt = np.concatenate([pt[keep], np.full(last.size, npts - 1)]).astype(np.int32)
        s = binding.host_structure(cam, pt, nc, npts, tile_size, slot_cap)
        _check_structure(cam, pt, nc, npts, s, tile_size, slot_cap)
        n_heavy_cases += len(s["hv_pt"]) > 0
    assert n_heavy_cases >= 20
