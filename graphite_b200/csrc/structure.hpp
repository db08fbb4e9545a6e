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
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

namespace gb {

struct HostStructure {
  int64_t M = 0;
  int32_t Nc = 0, Np = 0, tile = 256;
  bool identity_perm = true;
  std::vector<int64_t> perm;  // sorted position -> caller's factor index (empty when identity)
  std::vector<int32_t> cam_idx, pt_idx, pptr;
  std::vector<uint8_t> rank;
  std::vector<int32_t> tile_obs, tile_pt, tile_seg, seg_cam, seg_begin, cam_seg_ptr, cam_seg_list;
  int32_t max_track = 0;

  // returns an empty string on success, else the reason
  std::string build(int64_t nc, int64_t np, int64_t m, const int32_t *ci, const int32_t *pi, int tile_size) {
    if (nc <= 0 || np <= 0 || m <= 0) return "empty problem";
    if (m >= (int64_t(1) << 31) - 1024 || nc + np >= (int64_t(1) << 31)) return "problem too large for 32-bit indices";
    if (tile_size <= 0) tile_size = 256;
    if (tile_size > 256) return "tile_size must be <= 256";
    M = m; Nc = (int32_t)nc; Np = (int32_t)np; tile = tile_size;
    for (int64_t i = 0; i < m; i++)
      if (ci[i] < 0 || ci[i] >= nc || pi[i] < 0 || pi[i] >= np) return "observation index out of range";
    bool sorted = true;
    for (int64_t i = 1; i < m && sorted; i++)
      sorted = (pi[i] > pi[i - 1]) || (pi[i] == pi[i - 1] && ci[i] > ci[i - 1]);
    cam_idx.resize(m); pt_idx.resize(m);
    if (sorted) {
      identity_perm = true;
      std::copy(ci, ci + m, cam_idx.begin());
      std::copy(pi, pi + m, pt_idx.begin());
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
    pptr.assign((size_t)np + 1, 0);
    for (int64_t i = 0; i < m; i++) pptr[pt_idx[i] + 1]++;
    max_track = 0;
    for (int64_t p = 0; p < np; p++) {
      if (pptr[p + 1] == 0) return "a point has no observation (unused vertices are not supported)";
      max_track = std::max(max_track, pptr[p + 1]);
      pptr[p + 1] += pptr[p];
    }
    if (max_track > tile) return "a point has more observations than the tile size (" + std::to_string(max_track) + ")";
    {
      std::vector<uint8_t> seen((size_t)nc, 0);
      for (int64_t i = 0; i < m; i++) seen[cam_idx[i]] = 1;
      for (int64_t c = 0; c < nc; c++)
        if (!seen[c]) return "a camera has no observation (unused vertices are not supported)";
    }
    // tiles of whole points
    tile_obs.clear(); tile_pt.clear();
    tile_obs.push_back(0); tile_pt.push_back(0);
    int32_t cur = 0;
    for (int32_t p = 0; p < Np; p++) {
      const int32_t t = pptr[p + 1] - pptr[p];
      if (cur + t > tile) {
        tile_obs.push_back(pptr[p]); tile_pt.push_back(p);
        cur = 0;
      }
      cur += t;
    }
    tile_obs.push_back((int32_t)m); tile_pt.push_back(Np);
    const int32_t nt = (int32_t)tile_obs.size() - 1;
    // per-tile (camera, observation) order -> rank and camera segments
    rank.resize(m);
    tile_seg.assign((size_t)nt + 1, 0);
    seg_cam.clear(); seg_begin.clear();
    std::vector<std::pair<int32_t, int32_t>> loc;
    for (int32_t k = 0; k < nt; k++) {
      const int32_t o0 = tile_obs[k], n = tile_obs[k + 1] - o0;
      loc.resize(n);
      for (int32_t u = 0; u < n; u++) loc[u] = {cam_idx[o0 + u], u};
      std::sort(loc.begin(), loc.end());
      tile_seg[k] = (int32_t)seg_cam.size();
      for (int32_t u = 0; u < n; u++) {
        rank[o0 + loc[u].second] = (uint8_t)u;
        if (u == 0 || loc[u].first != loc[u - 1].first) {
          seg_cam.push_back(loc[u].first);
          seg_begin.push_back(o0 + u);
        }
      }
    }
    tile_seg[nt] = (int32_t)seg_cam.size();
    seg_begin.push_back((int32_t)m);
    const int32_t ns = (int32_t)seg_cam.size();
    cam_seg_ptr.assign((size_t)nc + 1, 0);
    for (int32_t s = 0; s < ns; s++) cam_seg_ptr[seg_cam[s] + 1]++;
    for (int64_t c = 0; c < nc; c++) cam_seg_ptr[c + 1] += cam_seg_ptr[c];
    cam_seg_list.resize(ns);
    std::vector<int32_t> fill(cam_seg_ptr.begin(), cam_seg_ptr.end() - 1);
    for (int32_t s = 0; s < ns; s++) cam_seg_list[fill[seg_cam[s]]++] = s; // ascending tile order per camera
    return "";
  }

  int32_t ntiles() const { return (int32_t)tile_obs.size() - 1; }
  int32_t nseg() const { return (int32_t)seg_cam.size(); }

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
