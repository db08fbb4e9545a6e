// structure_device.cuh — the data-parallel part of the structure build ON THE GPU.
//
// Replaces the host loops of FactorDescriptor::initialize_device_ids / setup_hessian_computation (factor.hpp:455-467,
// 702-763: one hash lookup per (pair, factor)) and SchurComplement::build_structure's tuple loop (schur.hpp:397-585) for
// this layout: everything whose size grows with the number of OBSERVATIONS is built by kernels - the per-tile records
// (slot order inside a tile, packed slot meta, camera segment table, point offsets), the slot of every observation, the
// camera of every slot, and the camera-major view k_prepare_cams walks (a stable radix sort by camera).  The host keeps
// the greedy cuts, which are sequential by nature and only need the point CSR: tiles (whole points, <= 256 observations,
// <= 128 points, <= slot_cap cameras), super-tiles, camera rows and chunk boundaries (structure.hpp, build()).
// The arrays are bit-identical with the host build (HostStructure::build with device_tables = false), which stays as the
// GPU-less view behind gb_structure_* for the CPU test-suite (tests/test_gpu_parity.py compares the two).
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "structure.hpp"

namespace gb {

// one CTA per tile, one thread per position of the tile's (point, camera) order
__global__ void __launch_bounds__(TILE)
k_build_tile_tables(int nt, const int32_t *__restrict__ cam_idx, const int32_t *__restrict__ pt_idx,
                    const int32_t *__restrict__ pptr, const int32_t *__restrict__ tile_obs,
                    const int32_t *__restrict__ tile_pt, const int32_t *__restrict__ tile_st,
                    const int32_t *__restrict__ st_row, const int32_t *__restrict__ row_cam,
                    const TileMeta *__restrict__ tmeta, uint32_t *__restrict__ ometa, unsigned char *__restrict__ trec,
                    int32_t *__restrict__ tile_cam, int32_t *__restrict__ slot_of_obs) {
  __shared__ int32_t cs[TILE];        // camera row slot of position u (INT_MAX: padding)
  __shared__ int32_t cs_sorted[TILE]; // ... of slot v
  const int k = blockIdx.x, u = threadIdx.x;
  const int32_t o0 = tile_obs[k], n = tile_obs[k + 1] - o0;
  const int32_t p0 = tile_pt[k], npt = max(1, tile_pt[k + 1] - p0); // fragments of a long track: np = 1
  const int32_t s = tile_st[k], r0 = st_row[s], nslots = st_row[s + 1] - r0;
  int32_t c = 0, cslot = 0x7fffffff;
  if (u < n) {
    c = cam_idx[o0 + u];
    int lo = 0, hi = nslots - 1; // the super-tile's camera rows are in ascending camera order
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (row_cam[r0 + mid] < c) lo = mid + 1;
      else hi = mid;
    }
    cslot = lo;
  }
  cs[u] = cslot;
  __syncthreads();
  // slot = rank in (camera row, position) order: a stable counting sort by camera row, written as a rank computation
  int32_t slot = u;
  if (u < n) {
    int32_t r = 0;
    for (int v = 0; v < n; v++) {
      const int32_t x = cs[v];
      r += (x < cslot) | ((x == cslot) & (v < u));
    }
    slot = r;
  }
  cs_sorted[slot] = cslot;
  __syncthreads();
  unsigned char *rec = trec + (size_t)k * REC_BYTES;
  uint32_t *om = reinterpret_cast<uint32_t *>(rec + REC_OMETA);
  uint32_t *sg = reinterpret_cast<uint32_t *>(rec + REC_SEG);
  if (u < n) {
    const uint32_t ptl = (uint32_t)(pt_idx[o0 + u] - p0);
    const uint32_t w = ((uint32_t)cslot << 16) | ((uint32_t)u << 8) | ptl;
    ometa[(size_t)k * TILE + slot] = w;
    om[slot] = w;
    slot_of_obs[o0 + u] = k * TILE + slot;
    tile_cam[(size_t)k * TILE + slot] = c;
  } else { // padding slots: unique point-order positions n..TILE-1, camera row 0, point 0
    ometa[(size_t)k * TILE + u] = (uint32_t)u << 8;
    om[u] = (uint32_t)u << 8;
    tile_cam[(size_t)k * TILE + u] = 0;
  }
  // camera segments: slot v starts one where the camera row changes; its index = number of starts before it
  if (u < n) {
    const int v = u;
    const int32_t mine = cs_sorted[v];
    if (v == 0 || cs_sorted[v - 1] != mine) {
      int32_t idx = 0;
      for (int q = 1; q <= v; q++) idx += cs_sorted[q] != cs_sorted[q - 1];
      sg[idx] = ((uint32_t)v << 16) | (uint32_t)mine;
    }
  }
  const TileMeta tm = tmeta[k];
  for (int i = tm.nseg + u; i < TILE + 4; i += TILE) sg[i] = (uint32_t)n << 16; // sentinel: end of the last segment
  uint16_t *pt = reinterpret_cast<uint16_t *>(rec + REC_PT);
  for (int i = u; i < TILE_PTS + 8; i += TILE) pt[i] = i <= npt ? (uint16_t)min(max(pptr[p0 + i] - o0, 0), n) : (uint16_t)n;
  if (u == 0) *reinterpret_cast<TileMeta *>(rec + REC_META) = tm;
  if (u < 4) {
    int32_t *nx = reinterpret_cast<int32_t *>(rec + REC_NEXT);
    const int d = u + 1;
    nx[2 * u] = k + d < nt ? tile_pt[k + d] : 0;
    nx[2 * u + 1] = k + d < nt ? max(1, tile_pt[k + d + 1] - tile_pt[k + d]) : 0;
  }
}

__global__ void k_iota32(int64_t n, int32_t *__restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}
// camera-major view: position in (camera, point) order -> storage slot / point of that observation
__global__ void k_camera_major(int64_t m, const int32_t *__restrict__ order, const int32_t *__restrict__ slot_of_obs,
                               const int32_t *__restrict__ pt_idx, int32_t *__restrict__ cm_slot, int32_t *__restrict__ cm_pt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int32_t o = order[i];
  cm_slot[i] = slot_of_obs[o];
  cm_pt[i] = pt_idx[o];
}

// Stable sort of the (point, camera)-sorted observations by camera: order_out[i] = observation at camera-major position i.
// keys_tmp / order_in / order_out / keys_out: scratch of M int32 each.  Returns a cudaError_t.
inline cudaError_t camera_major_order(int64_t m, int32_t nc, const int32_t *cam_idx, int32_t *keys_out, int32_t *order_in,
                                      int32_t *order_out, cudaStream_t st) {
  int bits = 1;
  while ((int64_t(1) << bits) < nc) bits++;
  k_iota32<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, order_in);
  size_t tmp_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, cam_idx, keys_out, order_in, order_out, (int)m, 0, bits, st);
  if (e != cudaSuccess) return e;
  void *tmp = nullptr;
  e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
  if (e != cudaSuccess) return e;
  e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, cam_idx, keys_out, order_in, order_out, (int)m, 0, bits, st);
  const cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFree(tmp);
  return e != cudaSuccess ? e : e2;
}

} // namespace gb
