// structure.hpp — one-time host-side structure build.
//
// Replaces the reference's structure discovery: FactorDescriptor::initialize_device_ids
// (include/graphite/factor.hpp:455-467), Graph::initialize_optimization ordering (graph.hpp:92-167),
// Hessian::build_structure (hessian.hpp:257-288) and SchurComplement::build_structure (schur.hpp:194-225).
// For BAL-type graphs (camera block column c = camera index, point block column = Nc + point index) the
// upper block-CSC of J^T J is fully determined by the observations sorted by (point, camera):
//   column c  < Nc : the single diagonal block (c, c)                           -> 81 values
//   column Nc + p  : blocks (cam, Nc+p) for the point's cameras ascending, then (Nc+p, Nc+p)
// so no hash maps or coordinate sorts are needed.
//
// Execution structure built here (see kernels.cuh):
//   TILE        256 storage slots = one CTA iteration; a tile holds whole points (<= tile_fill observations),
//               unused slots are padding with zero Jacobians.  Storage slot = tile * 256 + position.
//   SUPER-TILE  consecutive tiles owned by one CTA; it keeps one shared-memory accumulator row per distinct
//               camera it touches (<= slot_cap rows).  Rows of all super-tiles form the "partial rows" that the
//               per-camera kernels sum in ascending super-tile order (camera -> row CSR).
//   per slot    slots of a tile are in (camera, observation) order, so the threads of one camera segment are
//               adjacent; packed meta  cslot:16 | prank:8 | point-in-tile:8  (prank = position of the observation in
//               the tile's (point, camera) order, where the point-side sums are staged)
//   per tile    segment table (begin:16 | cslot:16) and point offset table.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

namespace gb {

constexpr int TILE = 256;       // storage slots per tile = threads per CTA
constexpr int SLOT_CAP = 192;   // camera accumulator rows per super-tile (bounded by shared memory)
constexpr int TILE_PTS = 128;   // max points per tile (bounds the per-tile W stage)
constexpr int CAM_CHUNK = 1024; // observations per CTA of the camera-major kernels (k_prepare_cams)
// Packed per-tile record (one TMA bulk copy): ometa[256] u32 | seg_tab[260] u32 | pt_tab[136] u16 | TileMeta |
// (p0, np) of the next four tiles
constexpr int REC_OMETA = 0;
constexpr int REC_SEG = REC_OMETA + TILE * 4;            // 1024
constexpr int REC_PT = REC_SEG + (TILE + 4) * 4;         // 2064
constexpr int REC_META = REC_PT + (TILE_PTS + 8) * 2;    // 2336
constexpr int REC_NEXT = REC_META + 32;                  // 2368: (p0, np) of tiles k+1 .. k+4 (TMA issue needs them)
constexpr int REC_BYTES = REC_NEXT + 32;                 // 2400 = 150 * 16

// Packed per-super-tile record (one TMA bulk copy, fetched a super-tile ahead by k_pcg_solve): first tile, tile count, first
// partial row, camera rows | camera of every row [SLOT_CAP] | position of every row in the partial buffers [SLOT_CAP]
constexpr int STREC_CAM = 16;
constexpr int STREC_OUT = STREC_CAM + SLOT_CAP * 4;
constexpr int STREC_BYTES = STREC_OUT + SLOT_CAP * 4;    // 1552 = 97 * 16

struct TileMeta {
  int32_t p0;      // first point
  int32_t n;       // observations in the tile
  int32_t np;      // points in the tile
  int32_t nseg;    // camera segments
  int32_t seg_off; // offset into seg_tab (multiple of 4)
  int32_t pt_off;  // offset into pt_tab (multiple of 8)
  int32_t o0;      // sorted observation index of slot 0
  int32_t frag;    // 0: the tile holds whole points.  Else the tile is a FRAGMENT of one long track (np = 1): bits 0-29 =
                   // index of the tile in the fragment list + 1, bit 30 = first fragment of its point (see build())
};
constexpr int32_t FRAG_FIRST = 1 << 30, FRAG_MASK = FRAG_FIRST - 1;

struct HostStructure {
  int64_t M = 0, Mstore = 0;
  int32_t Nc = 0, Np = 0, tile_fill = TILE, slot_cap = SLOT_CAP;
  // CTAs of the TMA kernels.  Measured on Venice: 148 persistent CTAs (static ranges, ring running across
  // super-tile boundaries) were 12 % slower than one CTA per super-tile scheduled by the hardware (SM-to-SM
  // spread of identical work is >= 10 %), so the default is one super-tile per CTA.
  int32_t persistent_ctas = 1 << 30;
  int32_t sm_count = 148;
  bool identity_perm = true;
  std::vector<int64_t> perm;  // sorted position -> caller's factor index (empty when identity)
  std::vector<int32_t> cam_idx, pt_idx, pptr;  // sorted by (point, camera)
  std::vector<int32_t> tile_obs, tile_pt;      // [ntiles+1] (sorted-observation / point boundaries)
  std::vector<TileMeta> tmeta;
  std::vector<uint32_t> ometa;                  // [Mstore]
  std::vector<uint32_t> seg_tab;
  std::vector<uint16_t> pt_tab;
  std::vector<uint8_t> trec;                    // [ntiles][REC_BYTES] packed copy of the three tables + meta
  std::vector<int32_t> st_tile, st_row, row_cam, cam_row_ptr, cam_row_list, slot_of_obs;
  std::vector<int32_t> row_out;                 // [nrows] super-tile row -> camera-major position in the partial buffers
  std::vector<int32_t> cta_st;                  // [ncta+1] super-tile ranges of the persistent CTAs (balanced by tiles)
  std::vector<uint8_t> strec;                   // [nst][STREC_BYTES] packed per-super-tile records
  std::vector<int32_t> tile_cam;                // [Mstore] camera per storage slot (0 in padding)
  // camera-major view (k_prepare_cams): the observations of camera c in ascending (point) order are entries
  // [cm_ptr(c), cm_ptr(c+1)); they are cut into chunks of <= CAM_CHUNK, one CTA and one partial row each
  std::vector<int32_t> cm_slot, cm_pt;          // [M] storage slot / point of the camera-major observation
  std::vector<int32_t> ch_ptr;                  // [nchunks+1] chunk -> range of camera-major observations
  std::vector<int32_t> cam_ch_ptr;              // [Nc+1] camera -> its contiguous chunk rows
  // per sorted observation (tests / host view): rank in tile order and camera segment count
  std::vector<uint8_t> rank;
  std::vector<int32_t> tile_ncam;               // [ntiles] distinct cameras (= camera segments) of each tile
  std::vector<int32_t> tile_st;                 // [ntiles] super-tile of each tile
  // Long tracks.  A point with more observations than one tile can hold (min(tile_fill, slot_cap)) is cut into FRAGMENT
  // tiles: consecutive tiles that hold nothing but a part of that point's observations (np = 1, all cameras distinct).
  // Its per-point sums then have two levels - the kernels write / read per-fragment values and a tiny kernel adds the
  // fragments of a point in order (kernels.cuh, "long tracks").
  std::vector<int32_t> frag_tile;               // [nfrag] tile of every fragment, ascending
  std::vector<int32_t> hv_pt, hv_ptr;           // [nheavy] long-track points, [nheavy+1] their ranges in frag_tile
  bool tables_on_device = false;
  int32_t max_track = 0;
  int64_t nseg_total = 0;

  // run fn(begin, end) over [0, n) split into contiguous ranges on the host's hardware threads (<= 16)
  static int host_threads(int64_t n) {
    return (int)std::min<int64_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u), std::max<int64_t>(n, 1));
  }
  template <typename F> static void parallel_indexed(int64_t n, int nth, F &&fn) { // fn(thread index, begin, end)
    if (nth <= 1) { fn(0, (int64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < nth; i++) th.emplace_back([&, i]() { fn(i, n * i / nth, n * (i + 1) / nth); });
    for (auto &t : th) t.join();
  }
  template <typename F> static void parallel_ranges(int64_t n, F &&fn) {
    parallel_indexed(n, host_threads(n), [&](int, int64_t b, int64_t e) { fn((int32_t)b, (int32_t)e); });
  }
  // compact per-tile tables (host view for tests / exports; the kernels read the packed records)
  void materialize_tables() {
    if (!seg_tab.empty() || tmeta.empty()) return;
    for (size_t k = 0; k < tmeta.size(); k++) {
      const TileMeta &tm = tmeta[k];
      const uint32_t *sg = reinterpret_cast<const uint32_t *>(trec.data() + k * REC_BYTES + REC_SEG);
      const uint16_t *pt = reinterpret_cast<const uint16_t *>(trec.data() + k * REC_BYTES + REC_PT);
      for (int32_t i = 0; i < (tm.nseg + 1 + 3) / 4 * 4; i++) seg_tab.push_back(sg[i]);
      for (int32_t i = 0; i < (tm.np + 1 + 7) / 8 * 8; i++) pt_tab.push_back(pt[i]);
    }
  }

  // returns an empty string on success, else the reason
  // device_tables = true: the observation-sized tables (ometa, trec contents, tile_cam, slot_of_obs, cm_slot, cm_pt) are
  // left to the kernels of structure_device.cuh; the host only makes the cuts (tiles, super-tiles, rows, chunks)
  std::string build(int64_t nc, int64_t np, int64_t m, const int32_t *ci, const int32_t *pi, int tile_size,
                    int slot_cap_opt = 0, int64_t st_obs_opt = 0, bool partition = false, bool device_tables = false) {
    // GB_STRUCT_TIMING=1: phase times of the build on stderr (profiles/README.md, "structure build")
    const bool timing = getenv("GB_STRUCT_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
      if (!timing) return;
      const auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[structure] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    };
    if (nc <= 0 || np <= 0 || m <= 0) return "empty problem";
    if (m >= (int64_t(1) << 30) || nc + np >= (int64_t(1) << 31)) return "problem too large for 32-bit indices";
    if (tile_size <= 0) tile_size = TILE;
    if (tile_size > TILE) return "tile_size must be <= 256";
    if (slot_cap_opt <= 0) slot_cap_opt = SLOT_CAP;
    if (slot_cap_opt > SLOT_CAP) return "slot cap must be <= 192";
    M = m; Nc = (int32_t)nc; Np = (int32_t)np; tile_fill = tile_size; slot_cap = slot_cap_opt;
    cam_idx.resize(m); pt_idx.resize(m);
    bool sorted = true;
    {
      // range check, order check and copy in one pass on the host threads
      const int nth = host_threads(m / 65536 + 1);
      std::vector<int> bad((size_t)nth, 0), unsorted((size_t)nth, 0);
      parallel_indexed(m, nth, [&](int t, int64_t ob, int64_t oe) {
        for (int64_t i = ob; i < oe; i++) {
          if (ci[i] < 0 || ci[i] >= nc || pi[i] < 0 || pi[i] >= np) bad[t] = 1;
          if (i > 0 && !((pi[i] > pi[i - 1]) || (pi[i] == pi[i - 1] && ci[i] > ci[i - 1]))) unsorted[t] = 1;
          cam_idx[i] = ci[i];
          pt_idx[i] = pi[i];
        }
      });
      for (int t = 0; t < nth; t++) {
        if (bad[t]) return "observation index out of range";
        if (unsorted[t]) sorted = false;
      }
    }
    if (sorted) {
      identity_perm = true;
    } else {
      identity_perm = false;
      perm.resize(m);
      std::iota(perm.begin(), perm.end(), int64_t(0));
      std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) {
        return pi[a] != pi[b] ? pi[a] < pi[b] : ci[a] < ci[b];
      });
      for (int64_t i = 0; i < m; i++) { cam_idx[i] = ci[perm[i]]; pt_idx[i] = pi[perm[i]]; }
      for (int64_t i = 1; i < m; i++)
        if (pt_idx[i] == pt_idx[i - 1] && cam_idx[i] == cam_idx[i - 1])
          return "duplicate (camera, point) observations are not supported";
    }
    lap("copy / order check / sort");
    pptr.assign((size_t)np + 1, 0);
    for (int64_t i = 0; i < m; i++) pptr[pt_idx[i] + 1]++;
    max_track = 0;
    for (int64_t p = 0; p < np; p++) {
      if (pptr[p + 1] == 0) return "a point has no observation (unused vertices are not supported)";
      max_track = std::max(max_track, pptr[p + 1]);
      pptr[p + 1] += pptr[p];
    }
    if (!partition) {
      std::vector<uint8_t> seen((size_t)nc, 0);
      for (int64_t i = 0; i < m; i++) seen[cam_idx[i]] = 1;
      for (int64_t c = 0; c < nc; c++)
        if (!seen[c]) return "a camera has no observation (unused vertices are not supported)";
    }
    lap("point CSR / unused checks");
    // ---- tiles of whole points ----------------------------------------------------------------------
    tile_obs.clear(); tile_pt.clear(); tile_ncam.clear();
    frag_tile.clear(); hv_pt.clear(); hv_ptr.assign(1, 0);
    tile_obs.push_back(0); tile_pt.push_back(0);
    int32_t cur = 0, ncam_tile = 0;
    const int32_t frag_cap = std::min(tile_fill, slot_cap); // longest track one tile can hold
    // the open tile ends here; the next one starts at sorted observation `next_obs`, point `next_pt`
    auto close_tile = [&](int32_t next_obs, int32_t next_pt) {
      tile_ncam.push_back(ncam_tile);
      tile_obs.push_back(next_obs); tile_pt.push_back(next_pt);
      cur = 0; ncam_tile = 0;
    };
    std::vector<int32_t> tstamp((size_t)nc, -1);
    for (int32_t p = 0; p < Np; p++) {
      const int32_t t = pptr[p + 1] - pptr[p];
      if (t > frag_cap) {
        // long track: near-equal fragments, one tile each (tile_pt = p for all of them, so np = max(1, difference))
        if (cur > 0) close_tile(pptr[p], p);
        const int32_t nf = (t + frag_cap - 1) / frag_cap;
        hv_pt.push_back(p);
        for (int32_t j = 0; j < nf; j++) {
          const int32_t e = pptr[p] + (int32_t)(((int64_t)t * (j + 1)) / nf);
          frag_tile.push_back((int32_t)tile_pt.size() - 1);
          cur = ncam_tile = e - tile_obs.back();
          if (j + 1 < nf) close_tile(e, p);
          else if (p + 1 < Np) close_tile(e, p + 1);
        }
        hv_ptr.push_back((int32_t)frag_tile.size());
        continue;
      }
      // cameras this point would add to the tile (a tile may not touch more cameras than a super-tile has rows)
      const int32_t tid = (int32_t)tile_pt.size() - 1;
      int32_t add = 0;
      for (int32_t o = pptr[p]; o < pptr[p + 1]; o++) add += tstamp[cam_idx[o]] != tid;
      if (cur + t > tile_fill || p - tile_pt.back() >= TILE_PTS || ncam_tile + add > slot_cap) close_tile(pptr[p], p);
      const int32_t tid2 = (int32_t)tile_pt.size() - 1;
      for (int32_t o = pptr[p]; o < pptr[p + 1]; o++)
        if (tstamp[cam_idx[o]] != tid2) { tstamp[cam_idx[o]] = tid2; ncam_tile++; }
      cur += t;
    }
    tile_obs.push_back((int32_t)m); tile_pt.push_back(Np);
    tile_ncam.push_back(ncam_tile);
    const int32_t nt = (int32_t)tile_obs.size() - 1;
    Mstore = (int64_t)nt * TILE;
    if (Mstore >= (int64_t(1) << 31)) return "problem too large for 32-bit slot indices";
    lap("tiles");
    // ---- super-tiles: consecutive tiles, bounded observation count and distinct cameras ----------------
    auto partition_tiles = [&](int64_t st_obs, std::vector<int32_t> &o_tile, std::vector<int32_t> &o_row,
                               std::vector<int32_t> &o_cam) -> bool { // re-entrant: the candidate lengths run concurrently
      std::vector<int32_t> stamp((size_t)nc, -1);
      const int32_t stamp_base = 0;
      o_tile.clear(); o_row.clear(); o_cam.clear();
      int32_t k = 0;
      while (k < nt) {
        const int32_t s = stamp_base + (int32_t)o_tile.size();
        o_tile.push_back(k);
        o_row.push_back((int32_t)o_cam.size());
        std::vector<int32_t> cams_here;
        int64_t obs_here = 0;
        while (k < nt) {
          std::vector<int32_t> add; // distinct cameras this tile would add
          for (int32_t o = tile_obs[k]; o < tile_obs[k + 1]; o++) {
            const int32_t c = cam_idx[o];
            if (stamp[c] != s) { stamp[c] = s; add.push_back(c); }
          }
          if (!cams_here.empty() && ((int32_t)(cams_here.size() + add.size()) > slot_cap || obs_here >= st_obs)) {
            for (int32_t c : add) stamp[c] = -1; // undo: the tile starts the next super-tile
            break;
          }
          if ((int32_t)add.size() > slot_cap) return false;
          cams_here.insert(cams_here.end(), add.begin(), add.end());
          obs_here += tile_obs[k + 1] - tile_obs[k];
          k++;
        }
        std::sort(cams_here.begin(), cams_here.end());
        o_cam.insert(o_cam.end(), cams_here.begin(), cams_here.end());
      }
      o_tile.push_back(nt);
      o_row.push_back((int32_t)o_cam.size());
      return true;
    };
    if (st_obs_opt > 0) {
      if (!partition_tiles(st_obs_opt, st_tile, st_row, row_cam)) return "a tile touches more cameras than the slot cap";
    } else {
      // The TMA kernels run one CTA per SM, so the number of super-tiles should fill whole waves of `sm_count`
      // CTAs, and a super-tile should be long enough (~12 tiles) to amortise the CTA's prologue and its row traffic:
      // measured on rank shares of Venice (scripts/partition_probe.py, product + reduction per launch) 2 waves beat
      // 6 by 30 % at 2.5 k tiles, 3 waves beat 8 by 20 % at 5 k tiles, 6-8 waves are level at 20 k tiles.
      // Around w0 = tiles / (12 sm_count), clamped to 2..8 waves, keep the fullest last wave.  Small problems
      // (under two waves of two tiles): one tile per super-tile.
      double best = -1.0;
      std::vector<int32_t> t_tile, t_row, t_cam;
      if (nt < 4 * sm_count) { // small problem: as many CTAs as tiles
        if (!partition_tiles(1, st_tile, st_row, row_cam)) return "a tile touches more cameras than the slot cap";
      } else {
        const int w0 = std::min(8, std::max(2, (int)((nt + 6 * sm_count) / (12 * sm_count))));
        // the (up to three) candidate lengths are independent greedy scans: one host thread each
        struct Trial { int w; int64_t target; std::vector<int32_t> t_tile, t_row, t_cam; bool ok = true; };
        std::vector<Trial> trials;
        for (int w = std::max(2, w0 - 1); w <= w0 + 1; w++) {
          Trial tr;
          tr.w = w;
          tr.target = std::max<int64_t>(TILE, (m + (int64_t)sm_count * w - 1) / ((int64_t)sm_count * w));
          trials.push_back(std::move(tr));
          if (trials.back().target == TILE) break;
        }
        {
          std::vector<std::thread> th;
          for (auto &tr : trials) {
            Trial *tp = &tr;
            th.emplace_back([&, tp]() { tp->ok = partition_tiles(tp->target, tp->t_tile, tp->t_row, tp->t_cam); });
          }
          for (auto &t : th) t.join();
        }
        for (auto &tr : trials) {
          if (!tr.ok) return "a tile touches more cameras than the slot cap";
          const int64_t n = (int64_t)tr.t_tile.size() - 1;
          double eff = (double)n / (double)(((n + sm_count - 1) / sm_count) * sm_count);
          if (tr.w == w0) eff += 0.02; // prefer the target length unless a neighbour fills its last wave clearly better
          if (eff > best + 1e-9) { best = eff; st_tile = tr.t_tile; st_row = tr.t_row; row_cam = tr.t_cam; }
        }
      }
    }
    const int32_t nst = (int32_t)st_tile.size() - 1;
    lap("super-tiles");
    // ---- per-tile tables ----------------------------------------------------------------------------------
    tables_on_device = device_tables;
    tmeta.assign((size_t)nt, TileMeta{});
    tile_st.assign((size_t)nt, 0);
    for (int32_t s = 0; s < nst; s++)
      for (int32_t k = st_tile[s]; k < st_tile[s + 1]; k++) tile_st[k] = s;
    for (int32_t k = 0; k < nt; k++) {
      TileMeta &tm = tmeta[k];
      tm.p0 = tile_pt[k]; tm.n = tile_obs[k + 1] - tile_obs[k]; tm.np = tile_np(k);
      tm.nseg = tile_ncam[k]; tm.o0 = tile_obs[k]; tm.frag = 0;
    }
    for (size_t h = 0; h < hv_pt.size(); h++)
      for (int32_t f = hv_ptr[h]; f < hv_ptr[h + 1]; f++) tmeta[frag_tile[f]].frag = (f + 1) | (f == hv_ptr[h] ? FRAG_FIRST : 0);
    seg_tab.clear(); pt_tab.clear(); // compact copies of the tables are materialised on demand (materialize_tables)
    ometa.clear(); rank.clear(); slot_of_obs.clear(); tile_cam.clear(); trec.clear();
    if (!device_tables) {
    ometa.assign((size_t)Mstore, 0u);
    rank.resize((size_t)m);
    slot_of_obs.resize(m);
    tile_cam.assign((size_t)Mstore, 0);
    trec.assign((size_t)nt * REC_BYTES, 0);
    // Per-tile tables, written straight into the packed records.  Super-tiles are independent: host threads take
    // contiguous ranges of them.  Inside a tile the observations are put in (camera, observation) order by a counting
    // sort on the camera's row slot in the super-tile (<= slot_cap values), which is the ascending camera order.
    parallel_ranges(nst, [&](int32_t s_begin, int32_t s_end) {
      std::vector<int32_t> local_t((size_t)nc, 0);
      int32_t start[SLOT_CAP + 1];
      int32_t cs[TILE], order[TILE];
      for (int32_t s = s_begin; s < s_end; s++) {
        const int32_t nslots = st_row[s + 1] - st_row[s];
        for (int32_t r = st_row[s]; r < st_row[s + 1]; r++) local_t[row_cam[r]] = r - st_row[s];
        for (int32_t k = st_tile[s]; k < st_tile[s + 1]; k++) {
          const int32_t o0 = tile_obs[k], n = tile_obs[k + 1] - o0;
          const int32_t p0 = tile_pt[k], npt = tile_np(k);
          TileMeta &tm = tmeta[k];
          tm.p0 = p0; tm.n = n; tm.np = npt; tm.o0 = o0;
          for (int32_t q = 0; q <= nslots; q++) start[q] = 0;
          for (int32_t u = 0; u < n; u++) { cs[u] = local_t[cam_idx[o0 + u]]; start[cs[u] + 1]++; }
          for (int32_t q = 0; q < nslots; q++) start[q + 1] += start[q];
          for (int32_t u = 0; u < n; u++) order[start[cs[u]]++] = u; // stable: ties keep the (point, camera) order
          uint8_t *rec = trec.data() + (size_t)k * REC_BYTES;
          uint32_t *om = reinterpret_cast<uint32_t *>(rec + REC_OMETA);
          uint32_t *sg = reinterpret_cast<uint32_t *>(rec + REC_SEG);
          int32_t nseg = 0;
          for (int32_t u = 0; u < n; u++) {
            const int32_t pos = order[u];
            rank[o0 + pos] = (uint8_t)u;
            const uint32_t cslot = (uint32_t)cs[pos];
            const uint32_t ptl = (uint32_t)(pt_idx[o0 + pos] - p0);
            // the slot of an observation is its position u in the tile's (camera, observation) order: threads of one
            // camera segment are adjacent (broadcast reads of the camera row, conflict-free staging); the meta keeps
            // the position `pos` in (point, camera) order, where the point-side sums are staged
            const uint32_t w = (cslot << 16) | ((uint32_t)pos << 8) | ptl;
            ometa[(size_t)k * TILE + u] = w;
            om[u] = w;
            slot_of_obs[o0 + pos] = k * TILE + u;
            tile_cam[(size_t)k * TILE + u] = cam_idx[o0 + pos];
            if (u == 0 || cslot != (uint32_t)cs[order[u - 1]]) sg[nseg++] = ((uint32_t)u << 16) | cslot;
          }
          // padding slots: unique point-order positions n..TILE-1 (their rows are never read), camera slot 0, point 0
          for (int32_t u = n; u < TILE; u++) { ometa[(size_t)k * TILE + u] = ((uint32_t)u << 8); om[u] = ((uint32_t)u << 8); }
          tm.nseg = nseg;
          for (int32_t i = nseg; i < TILE + 4; i++) sg[i] = ((uint32_t)n << 16); // sentinel: end of the last segment
          uint16_t *pt = reinterpret_cast<uint16_t *>(rec + REC_PT);
          // (clamped to the tile: a fragment holds a part of its point's observations)
          for (int32_t i = 0; i < TILE_PTS + 8; i++)
            pt[i] = i <= npt ? (uint16_t)std::min(std::max(pptr[p0 + i] - o0, 0), n) : (uint16_t)n;
          int32_t *nx = reinterpret_cast<int32_t *>(rec + REC_NEXT);
          for (int32_t d = 1; d <= 4; d++) {
            nx[2 * (d - 1)] = k + d < nt ? tile_pt[k + d] : 0;
            nx[2 * (d - 1) + 1] = k + d < nt ? tile_np(k + d) : 0;
          }
        }
      }
    });
    } // !device_tables
    lap("per-tile tables");
    // offsets of the compact tables (multiples of 4 / 8 entries per tile) and the tile meta inside the records
    nseg_total = 0;
    {
      int32_t seg_off = 0, pt_off = 0;
      for (int32_t k = 0; k < nt; k++) {
        TileMeta &tm = tmeta[k];
        tm.seg_off = seg_off; tm.pt_off = pt_off;
        seg_off += (tm.nseg + 1 + 3) / 4 * 4;
        pt_off += (tm.np + 1 + 7) / 8 * 8;
        nseg_total += tm.nseg;
        if (!device_tables) memcpy(trec.data() + (size_t)k * REC_BYTES + REC_META, &tm, sizeof(TileMeta));
      }
    }
    // ---- camera -> partial rows, ascending super-tile order ------------------------------------------------
    const int32_t nrows = (int32_t)row_cam.size();
    cam_row_ptr.assign((size_t)nc + 1, 0);
    for (int32_t r = 0; r < nrows; r++) cam_row_ptr[row_cam[r] + 1]++;
    for (int64_t c = 0; c < nc; c++) cam_row_ptr[c + 1] += cam_row_ptr[c];
    cam_row_list.resize(nrows);
    std::vector<int32_t> fill(cam_row_ptr.begin(), cam_row_ptr.end() - 1);
    for (int32_t r = 0; r < nrows; r++) cam_row_list[fill[row_cam[r]]++] = r;
    // partial buffers are camera-major: the rows of one camera are contiguous, in ascending super-tile order
    row_out.resize(nrows);
    for (int32_t i = 0; i < nrows; i++) row_out[cam_row_list[i]] = i;
    // ---- packed per-super-tile records -----------------------------------------------------------------------
    // Work-queue order of k_pcg_solve = record order: longest super-tiles first, so that the items the atomic counter hands
    // out last are the short ones and the CTAs finish the product phase close together (longest-processing-time rule;
    // same-box A/B, gpurun_out/r2m_lpt_ab.log: product phase 240 -> 224 us at Venice, 122 -> 118 us on half of it).
    strec.assign((size_t)nst * STREC_BYTES, 0);
    std::vector<int32_t> order((size_t)nst);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
      return st_tile[a + 1] - st_tile[a] > st_tile[b + 1] - st_tile[b];
    });
    for (int32_t j = 0; j < nst; j++) {
      const int32_t s2 = order[j];
      int32_t *r = reinterpret_cast<int32_t *>(strec.data() + (size_t)j * STREC_BYTES);
      r[0] = st_tile[s2]; r[1] = st_tile[s2 + 1] - st_tile[s2];
      r[2] = st_row[s2];  r[3] = st_row[s2 + 1] - st_row[s2];
      for (int32_t q = 0; q < r[3]; q++) {
        r[STREC_CAM / 4 + q] = row_cam[st_row[s2] + q];
        r[STREC_OUT / 4 + q] = row_out[st_row[s2] + q];
      }
    }
    lap("rows / super-tile records");
    // ---- camera-major view: counting sort by camera (stable: ascending point order inside a camera) ---------
    {
      // stable counting sort by camera on the host threads: per-thread histograms over contiguous observation ranges
      const int nth = host_threads(m / 65536 + 1);
      std::vector<std::vector<int32_t>> hist((size_t)nth, std::vector<int32_t>((size_t)nc, 0));
      parallel_indexed(m, nth, [&](int i, int64_t ob, int64_t oe) {
        int32_t *h = hist[i].data();
        for (int64_t o = ob; o < oe; o++) h[cam_idx[o]]++;
      });
      std::vector<int32_t> cptr((size_t)nc + 1, 0);
      {
        int32_t run = 0;
        for (int64_t c = 0; c < nc; c++) {
          cptr[c] = run;
          for (int i = 0; i < nth; i++) { const int32_t v = hist[i][c]; hist[i][c] = run; run += v; }
        }
        cptr[nc] = run;
      }
      cm_slot.clear(); cm_pt.clear();
      if (!device_tables) {
        cm_slot.resize(m); cm_pt.resize(m);
        parallel_indexed(m, nth, [&](int i, int64_t ob, int64_t oe) {
          int32_t *h = hist[i].data();
          for (int64_t o = ob; o < oe; o++) {
            const int32_t pos = h[cam_idx[o]]++;
            cm_slot[pos] = slot_of_obs[o];
            cm_pt[pos] = pt_idx[o];
          }
        });
      }
      ch_ptr.assign(1, 0);
      cam_ch_ptr.assign((size_t)nc + 1, 0);
      for (int64_t c = 0; c < nc; c++) {
        const int32_t n = cptr[c + 1] - cptr[c];
        const int32_t nch = (n + CAM_CHUNK - 1) / CAM_CHUNK;
        for (int32_t k = 0; k < nch; k++) // near-equal chunks
          ch_ptr.push_back(cptr[c] + (int32_t)(((int64_t)n * (k + 1)) / nch));
        cam_ch_ptr[c + 1] = cam_ch_ptr[c] + nch;
      }
    }
    lap("camera-major view");
    // persistent CTAs: contiguous super-tile ranges with near-equal tile counts
    {
      const int32_t ncta = std::min<int32_t>(nst, persistent_ctas);
      cta_st.assign(1, 0);
      for (int32_t b = 1; b < ncta; b++) {
        const int64_t target = (int64_t)nt * b / ncta; // first tile of CTA b
        int32_t sidx = (int32_t)(std::lower_bound(st_tile.begin(), st_tile.end() - 1, (int32_t)target) - st_tile.begin());
        sidx = std::max(sidx, cta_st.back() + 1);
        sidx = std::min(sidx, nst - (ncta - b));
        cta_st.push_back(sidx);
      }
      cta_st.push_back(nst);
    }
    return "";
  }

  int32_t ntiles() const { return (int32_t)tile_obs.size() - 1; }
  // points of tile k: the fragments of a long track all start at their point, so consecutive tile_pt entries are equal
  int32_t tile_np(int32_t k) const { return std::max(1, tile_pt[k + 1] - tile_pt[k]); }
  int32_t nfrag() const { return (int32_t)frag_tile.size(); }
  int32_t nheavy() const { return (int32_t)hv_pt.size(); }
  int32_t nst() const { return (int32_t)st_tile.size() - 1; }
  int32_t nrows() const { return (int32_t)row_cam.size(); }
  int32_t ncta() const { return (int32_t)cta_st.size() - 1; }
  int32_t nchunks() const { return (int32_t)ch_ptr.size() - 1; }

  // Upper block-CSC of the Schur complement S = B - E C^-1 E^T in the reference's order
  // (SchurComplement::build_structure, schur.hpp:194-225, 397-585; csc_utils.hpp:16-50): block (i, j), i <= j, exists
  // iff i == j or cameras i and j observe a common point; columns ascending, rows ascending inside a column.
  void schur_structure(std::vector<int64_t> &colptr, std::vector<int32_t> &rowidx) const {
    std::vector<std::vector<int32_t>> rows((size_t)Nc);
    for (int32_t c = 0; c < Nc; c++) rows[c].push_back(c);
    for (int32_t p = 0; p < Np; p++)
      for (int32_t a = pptr[p]; a < pptr[p + 1]; a++)
        for (int32_t b = a + 1; b < pptr[p + 1]; b++) rows[cam_idx[b]].push_back(cam_idx[a]); // cam_idx[a] < cam_idx[b]
    colptr.assign((size_t)Nc + 1, 0);
    rowidx.clear();
    for (int32_t c = 0; c < Nc; c++) {
      auto &v = rows[c];
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      rowidx.insert(rowidx.end(), v.begin(), v.end());
      colptr[c + 1] = (int64_t)rowidx.size();
      std::vector<int32_t>().swap(v);
    }
  }

  // Execution structure of the EXPLICIT Schur complement (SchurComplement::build_structure / setup_schur_multiplication,
  // schur.hpp:397-585, where the reference emits one MulOp per (point, camera pair) and scatters with atomics).  Here the
  // (point, pair) tuples are sorted ONCE by the block they contribute to, so that one warp owns a block and sums its
  // tuples in a fixed order (no atomics), and the symmetric matrix gets a row view for the product S p:
  //   tptr[b] .. tptr[b+1]   tuples (slot of the row camera's observation, slot of the column camera's, point) of
  //                          off-diagonal block b (blocks in the order of schur_structure; diagonal blocks have none)
  //   row_ptr[c] ..          entries of block row c of the full symmetric matrix: block index << 1 | transposed, and
  //                          the camera whose part of p the block multiplies; the diagonal block is not listed
  struct ExplicitSchur {
    std::vector<int64_t> colptr;
    std::vector<int32_t> rowidx;
    std::vector<int32_t> diag_block;           // [Nc] index of block (c, c)
    std::vector<int64_t> tptr;                 // [nblocks + 1]
    std::vector<int32_t> tup_a, tup_b, tup_p;  // [ntuples]
    std::vector<int32_t> blk_row, blk_col;     // [nblocks] cameras of the block
    std::vector<int32_t> row_ptr, row_ent, row_other;
    int64_t ntuples = 0;
  };
  void explicit_schur(ExplicitSchur &E) const {
    schur_structure(E.colptr, E.rowidx);
    const int64_t nb = (int64_t)E.rowidx.size();
    E.diag_block.resize((size_t)Nc);
    E.blk_row.resize((size_t)nb); E.blk_col.resize((size_t)nb);
    for (int32_t c = 0; c < Nc; c++) {
      E.diag_block[c] = (int32_t)(E.colptr[c + 1] - 1); // rows ascending: the diagonal block is the last of its column
      for (int64_t k = E.colptr[c]; k < E.colptr[c + 1]; k++) { E.blk_row[k] = E.rowidx[k]; E.blk_col[k] = c; }
    }
    auto find_block = [&](int32_t row, int32_t col) -> int64_t { // row < col
      const int32_t *lo = E.rowidx.data() + E.colptr[col], *hi = E.rowidx.data() + E.colptr[col + 1];
      return (int64_t)(std::lower_bound(lo, hi, row) - E.rowidx.data());
    };
    // count, prefix, fill (points ascending, pairs in (a, b) order inside a point: a fixed summation order per block)
    E.tptr.assign((size_t)nb + 1, 0);
    for (int32_t p = 0; p < Np; p++)
      for (int32_t a = pptr[p]; a < pptr[p + 1]; a++)
        for (int32_t b = a + 1; b < pptr[p + 1]; b++) E.tptr[find_block(cam_idx[a], cam_idx[b]) + 1]++;
    for (int64_t k = 0; k < nb; k++) E.tptr[k + 1] += E.tptr[k];
    E.ntuples = E.tptr[nb];
    E.tup_a.resize((size_t)E.ntuples); E.tup_b.resize((size_t)E.ntuples); E.tup_p.resize((size_t)E.ntuples);
    std::vector<int64_t> fill(E.tptr.begin(), E.tptr.end() - 1);
    for (int32_t p = 0; p < Np; p++)
      for (int32_t a = pptr[p]; a < pptr[p + 1]; a++)
        for (int32_t b = a + 1; b < pptr[p + 1]; b++) {
          const int64_t pos = fill[find_block(cam_idx[a], cam_idx[b])]++;
          E.tup_a[pos] = slot_of_obs[a]; E.tup_b[pos] = slot_of_obs[b]; E.tup_p[pos] = p;
        }
    // row view of the symmetric matrix without the diagonal: for camera c the blocks (r, c), r < c, transposed, then the
    // blocks (c, j), j > c, as stored
    E.row_ptr.assign((size_t)Nc + 1, 0);
    for (int64_t k = 0; k < nb; k++)
      if (E.blk_row[k] != E.blk_col[k]) { E.row_ptr[E.blk_row[k] + 1]++; E.row_ptr[E.blk_col[k] + 1]++; }
    for (int32_t c = 0; c < Nc; c++) E.row_ptr[c + 1] += E.row_ptr[c];
    E.row_ent.resize((size_t)E.row_ptr[Nc]); E.row_other.resize((size_t)E.row_ptr[Nc]);
    std::vector<int32_t> rf(E.row_ptr.begin(), E.row_ptr.end() - 1);
    for (int32_t c = 0; c < Nc; c++) // column c, rows r < c: entry of row c (transposed) -> ascending r
      for (int64_t k = E.colptr[c]; k < E.colptr[c + 1] - 1; k++) {
        const int32_t q = rf[c]++;
        E.row_ent[q] = (int32_t)(k << 1) | 1; E.row_other[q] = E.rowidx[k];
      }
    for (int32_t c = 0; c < Nc; c++) // column c, rows r < c: entry of row r (as stored) -> ascending c
      for (int64_t k = E.colptr[c]; k < E.colptr[c + 1] - 1; k++) {
        const int32_t r = E.rowidx[k];
        const int32_t q = rf[r]++;
        E.row_ent[q] = (int32_t)(k << 1); E.row_other[q] = c;
      }
  }

  // Upper block-CSC of the Hessian in the reference's order (hessian.hpp:59-84, 270-278; csc_utils.hpp:16-50).
  void hessian_structure(int64_t *colptr, int64_t *rowidx, int64_t *offsets) const {
    int64_t k = 0, off = 0;
    for (int64_t c = 0; c < Nc; c++) { colptr[c] = k; rowidx[k] = c; offsets[k] = off; off += 81; k++; }
    for (int64_t p = 0; p < Np; p++) {
      colptr[Nc + p] = k;
      for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) { rowidx[k] = cam_idx[f]; offsets[k] = off; off += 27; k++; }
      rowidx[k] = Nc + p; offsets[k] = off; off += 9; k++;
    }
    colptr[(int64_t)Nc + Np] = k;
  }
};

} // namespace gb
