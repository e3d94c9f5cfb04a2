// Deterministic per-codeword statistics of the VQ update: the one-hot^T @ z GEMM and the one-hot column sum of
// vq_gnn_v2/vq.py:177,191 (feature_update) and :243,256 (update), as an ORDERED segmented sum.
//
//   stats[k, m, :W] = sum_{b : code[b,k] == m} z[b, k, :]      stats[k, m, Wp] = #{b : code[b,k] == m}
//
// No floating-point atomics: (1) the (branch, code) keys of all B x nb assignments are radix-sorted (stable, so the
// rows of a codeword stay in ascending order), (2) segment boundaries are marked, (3) one warp per (branch, codeword)
// sums its rows -- lane l takes rows l, l+32, ... in order, then a fixed shuffle tree -- so the result is a pure
// function of the inputs (bit-identical between runs, streams and launch orders).  The counts are segment lengths
// (exact integers).  z is re-whitened on the fly from x / g with the same fmaf as the assignment kernels.
// HBM-bound integer + gather work: B*nb*(8 B keys) sorted in ceil(log2(nb*M)/8) passes + one 16..32 B gather per
// (row, branch).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace vqgnn {

__global__ void segsum_keys_kernel(const int16_t* __restrict__ idx, int64_t n, int nbc, int M,
                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ rows) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t b = i / nbc;
    const int k = static_cast<int>(i - b * nbc);
    keys[i] = static_cast<uint32_t>(k) * M + static_cast<uint32_t>(idx[i]);
    rows[i] = static_cast<uint32_t>(b);
  }
}

// seg_start[s] = first position whose key is >= s, for s in [0, S]
__global__ void segsum_bounds_kernel(const uint32_t* __restrict__ keys, int64_t n, int S,
                                     int32_t* __restrict__ seg_start) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += stride) {
    const int lo = i == 0 ? 0 : static_cast<int>(keys[i - 1]) + 1;
    const int hi = i == n ? S : static_cast<int>(keys[i]);
    for (int s = lo; s <= hi; ++s) seg_start[s] = static_cast<int32_t>(i);
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
    segsum_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g, int64_t ldg,
                  const float* __restrict__ scale, const float* __restrict__ shift,
                  const uint32_t* __restrict__ rows, const int32_t* __restrict__ seg_start, int S, int nbc, int M,
                  int D, int Dg, int Wp, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int seg = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (seg >= S) return;
  const int k = seg / M;
  const int C = nbc * D;
  const int w_use = D + (g ? Dg : 0);
  const int s0 = __ldg(seg_start + seg), s1 = __ldg(seg_start + seg + 1);
  float sc[NV * 4], sh[NV * 4];
#pragma unroll
  for (int w = 0; w < NV * 4; ++w) {
    sc[w] = 0.f, sh[w] = 0.f;
    if (w < D) sc[w] = __ldg(scale + k * D + w), sh[w] = __ldg(shift + k * D + w);
    else if (w < w_use) sc[w] = __ldg(scale + C + k * Dg + (w - D)), sh[w] = __ldg(shift + C + k * Dg + (w - D));
  }
  float acc[NV * 4];
#pragma unroll
  for (int w = 0; w < NV * 4; ++w) acc[w] = 0.f;
  for (int i = s0 + lane; i < s1; i += 32) {
    const int64_t b = __ldg(rows + i);
#pragma unroll
    for (int w = 0; w < NV * 4; ++w) {
      float v = 0.f;
      if (w < D) v = fmaf(__ldg(x + b * ldx + k * D + w), sc[w], sh[w]);
      else if (w < w_use) v = fmaf(__ldg(g + b * ldg + k * Dg + (w - D)), sc[w], sh[w]);
      acc[w] += v;
    }
  }
#pragma unroll
  for (int w = 0; w < NV * 4; ++w) acc[w] = warp_sum(acc[w]);   // fixed xor tree: order independent of timing
  float* dst = stats + static_cast<int64_t>(seg) * (Wp + 4);
  if (lane == 0) {
#pragma unroll
    for (int w = 0; w < NV * 4; ++w)
      if (w < Wp) dst[w] = w < w_use ? acc[w] : 0.f;
    for (int w = NV * 4; w < Wp; ++w) dst[w] = 0.f;
    dst[Wp] = static_cast<float>(s1 - s0);
    dst[Wp + 1] = 0.f, dst[Wp + 2] = 0.f, dst[Wp + 3] = 0.f;
  }
}

static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

static int key_bits(int64_t S) {
  int bits = 1;
  while ((static_cast<int64_t>(1) << bits) < S) ++bits;
  return bits;
}

static size_t cub_sort_bytes(int64_t n, int bits) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), n, 0, bits);
  return tmp;
}

}  // namespace vqgnn

using namespace vqgnn;

extern "C" size_t vqgnn_vq_segsum_workspace_bytes(int64_t B, int nbc, int M) {
  const int64_t n = B * nbc, S = static_cast<int64_t>(nbc) * M;
  return 4 * align256(static_cast<size_t>(n) * 4) + align256(static_cast<size_t>(S + 1) * 4) +
         align256(cub_sort_bytes(n, key_bits(S))) + 256;
}

extern "C" int vqgnn_vq_segsum(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* scale,
                               const float* shift, const int16_t* idx, int64_t B, int nbc, int M, int D, int Dg,
                               int Wp, float* stats, void* ws, size_t ws_bytes, void* stream) {
  VQ_CHECK_ARG(x && scale && shift && idx && stats && ws && B > 0 && nbc > 0 && M > 0 && D > 0,
               "vq_segsum: bad arguments");
  VQ_CHECK_ARG(!g || Dg == D || Dg == D + 1, "vq_segsum: Dg must be D or D+1");
  VQ_CHECK_ARG(Wp % 4 == 0 && Wp >= D + (g ? Dg : 0), "vq_segsum: Wp must be a multiple of 4 covering W");
  const int64_t n = B * nbc, S = static_cast<int64_t>(nbc) * M;
  VQ_CHECK_ARG(n < (1ll << 31) && S < (1ll << 31), "vq_segsum: B*nb and nb*M must fit int32");
  VQ_CHECK_ARG(ws_bytes >= vqgnn_vq_segsum_workspace_bytes(B, nbc, M), "vq_segsum: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  const size_t an = align256(static_cast<size_t>(n) * 4);
  uint32_t* keys_in = reinterpret_cast<uint32_t*>(p);
  uint32_t* keys_out = reinterpret_cast<uint32_t*>(p + an);
  uint32_t* rows_in = reinterpret_cast<uint32_t*>(p + 2 * an);
  uint32_t* rows_out = reinterpret_cast<uint32_t*>(p + 3 * an);
  int32_t* seg_start = reinterpret_cast<int32_t*>(p + 4 * an);
  void* cub_tmp = p + 4 * an + align256(static_cast<size_t>(S + 1) * 4);
  const int bits = key_bits(S);
  size_t cub_bytes = cub_sort_bytes(n, bits);
  const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, 16 * kNumSMs));
  segsum_keys_kernel<<<grid, 256, 0, s>>>(idx, n, nbc, M, keys_in, rows_in);
  VQ_LAUNCH_CHECK();
  VQ_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, rows_in, rows_out, n, 0, bits, s));
  count_launch(2);
  segsum_bounds_kernel<<<grid, 256, 0, s>>>(keys_out, n, static_cast<int>(S), seg_start);
  VQ_LAUNCH_CHECK();
  const int w_use = D + (g ? Dg : 0);
  const int nv = (w_use + 3) / 4;
  const int sgrid = static_cast<int>((S + 7) / 8);
#define VQ_SEGSUM(NV)                                                                                          \
  segsum_kernel<NV><<<sgrid, 256, 0, s>>>(x, ldx, g, ldg, scale, shift, rows_out, seg_start, (int)S, nbc, M, D, \
                                          g ? Dg : 0, Wp, stats)
  switch (nv) {
    case 1: VQ_SEGSUM(1); break;
    case 2: VQ_SEGSUM(2); break;
    case 3: VQ_SEGSUM(3); break;
    case 4: VQ_SEGSUM(4); break;
    case 5: VQ_SEGSUM(5); break;
    default: set_error("vq_segsum: joint width %d > 20 is not supported", w_use); return VQGNN_ERR_ARG;
  }
#undef VQ_SEGSUM
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
