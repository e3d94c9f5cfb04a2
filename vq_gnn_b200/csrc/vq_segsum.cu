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
#include <cub/device/device_scan.cuh>

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

// Long codeword segments (early in training most rows share a few codewords) must not serialise one warp: every
// segment is cut into fixed sub-ranges of kSegSub rows.  nsub[s] = ceil(len / kSegSub) (0 for an empty segment ->
// one task that writes zeros is still needed, so max(1, .)); task_off = exclusive scan.  One warp per task.
constexpr int kSegSub = 256;

__global__ void segsum_nsub_kernel(const int32_t* __restrict__ seg_start, int S, int32_t* __restrict__ nsub) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int len = seg_start[s + 1] - seg_start[s];
  nsub[s] = max(1, (len + kSegSub - 1) / kSegSub);
}

// z of one row (branch k) whitened exactly like the assignment kernels: fmaf(v, scale, shift)
template <int NV>
__device__ __forceinline__ void segsum_load(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g,
                                            int64_t ldg, int64_t b, int k, int D, int Dg, int w_use, bool vec,
                                            const float (&sc)[NV * 4], const float (&sh)[NV * 4],
                                            float (&acc)[NV * 4]) {
  if (vec) {   // D == Dg == 4 (or no g), 16 B aligned rows: two 128-bit loads
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + b * ldx + k * 4));
    acc[0] += fmaf(xv.x, sc[0], sh[0]), acc[1] += fmaf(xv.y, sc[1], sh[1]);
    acc[2] += fmaf(xv.z, sc[2], sh[2]), acc[3] += fmaf(xv.w, sc[3], sh[3]);
    if constexpr (NV >= 2) {
      if (g) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + b * ldg + k * 4));
        acc[4] += fmaf(gv.x, sc[4], sh[4]), acc[5] += fmaf(gv.y, sc[5], sh[5]);
        acc[6] += fmaf(gv.z, sc[6], sh[6]), acc[7] += fmaf(gv.w, sc[7], sh[7]);
      }
    }
    return;
  }
#pragma unroll
  for (int w = 0; w < NV * 4; ++w) {
    float v = 0.f;
    if (w < D) v = fmaf(__ldg(x + b * ldx + k * D + w), sc[w], sh[w]);
    else if (w < w_use) v = fmaf(__ldg(g + b * ldg + k * Dg + (w - D)), sc[w], sh[w]);
    acc[w] += v;
  }
}

// One task = one sub-range (<= kSegSub rows) of one (branch, codeword) segment, handled by a group of LG lanes: lane l of
// the group takes rows l, l+LG, ... in order, then a fixed xor tree over the group -- a pure function of the inputs.
// LG = 32 (a warp per task) when segments are long; LG = 8 when the average segment is a handful of rows (M = 4096:
// ~5 rows per codeword), where a whole warp per segment left 27 lanes idle and the kernel was bound by the latency of
// 131 K tiny warps.
template <int NV, int LG>
__global__ void __launch_bounds__(256)
    segsum_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g, int64_t ldg,
                  const float* __restrict__ scale, const float* __restrict__ shift,
                  const uint32_t* __restrict__ rows, const int32_t* __restrict__ seg_start,
                  const int32_t* __restrict__ task_off, int S, int max_tasks, int nbc, int M, int D, int Dg, int Wp,
                  int vec, float* __restrict__ stats, float* __restrict__ part) {
  constexpr int GPW = 32 / LG;   // groups per warp
  const int lane = threadIdx.x & 31, gl = lane & (LG - 1);
  const int task = (blockIdx.x * 8 + (threadIdx.x >> 5)) * GPW + lane / LG;
  const int n_tasks = __ldg(task_off + S);
  const bool live = task < max_tasks && task < n_tasks;     // dead groups still take part in the shuffles below
  // largest seg with task_off[seg] <= task.  task_off[seg] = seg + (extra sub-ranges before seg), so the answer lies
  // in [task - extra_total, task]: a handful of search steps when few segments are long (the common case)
  int seg = 0, sub = 0, nsub = 1, s0 = 0, s1 = 0;
  if (live) {
    const int extra_total = n_tasks - S;
    int lo = max(0, task - extra_total), hi = min(task, S - 1) + 1;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(task_off + mid) <= task) lo = mid;
      else hi = mid;
    }
    seg = lo, sub = task - __ldg(task_off + seg);
    nsub = __ldg(task_off + seg + 1) - __ldg(task_off + seg);
    s0 = __ldg(seg_start + seg) + sub * kSegSub;
    s1 = min(__ldg(seg_start + seg + 1), s0 + kSegSub);
  }
  const int k = seg / M;
  const int C = nbc * D;
  const int w_use = D + (g ? Dg : 0);
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
  for (int i = s0 + gl; i < s1; i += LG)
    segsum_load<NV>(x, ldx, g, ldg, static_cast<int64_t>(__ldg(rows + i)), k, D, Dg, w_use, vec != 0, sc, sh, acc);
#pragma unroll
  for (int w = 0; w < NV * 4; ++w) {   // fixed xor tree inside the group: order independent of timing
#pragma unroll
    for (int o = LG / 2; o > 0; o >>= 1) acc[w] += __shfl_xor_sync(0xffffffffu, acc[w], o);
  }
  if (gl != 0 || !live) return;
  if (nsub == 1) {
    float* dst = stats + static_cast<int64_t>(seg) * (Wp + 4);
#pragma unroll
    for (int w = 0; w < NV * 4; ++w)
      if (w < Wp) dst[w] = w < w_use ? acc[w] : 0.f;
    for (int w = NV * 4; w < Wp; ++w) dst[w] = 0.f;
    dst[Wp] = static_cast<float>(__ldg(seg_start + seg + 1) - __ldg(seg_start + seg));
    dst[Wp + 1] = 0.f, dst[Wp + 2] = 0.f, dst[Wp + 3] = 0.f;
  } else {
    float* dst = part + static_cast<int64_t>(task) * (NV * 4);
#pragma unroll
    for (int w = 0; w < NV * 4; ++w) dst[w] = acc[w];
  }
}

// segments cut into several sub-ranges: their partial sums added in sub-range order (thread per (segment, column))
template <int NV>
__global__ void segsum_combine_kernel(const int32_t* __restrict__ seg_start, const int32_t* __restrict__ task_off,
                                      int S, int w_use, int Wp, const float* __restrict__ part,
                                      float* __restrict__ stats) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int seg = static_cast<int>(i / (NV * 4)), w = static_cast<int>(i - static_cast<int64_t>(seg) * (NV * 4));
  if (seg >= S) return;
  const int t0 = __ldg(task_off + seg), nsub = __ldg(task_off + seg + 1) - t0;
  if (nsub <= 1) return;
  float t = 0.f;
  for (int j = 0; j < nsub; ++j) t += part[static_cast<int64_t>(t0 + j) * (NV * 4) + w];
  float* dst = stats + static_cast<int64_t>(seg) * (Wp + 4);
  if (w < Wp) dst[w] = w < w_use ? t : 0.f;
  if (w == 0) {
    for (int u = NV * 4; u < Wp; ++u) dst[u] = 0.f;
    dst[Wp] = static_cast<float>(__ldg(seg_start + seg + 1) - __ldg(seg_start + seg));
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

static size_t cub_scan_bytes(int64_t n) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                static_cast<int>(n));
  return tmp;
}
static int64_t segsum_max_tasks(int64_t n, int64_t S) { return n / kSegSub + S + 1; }

extern "C" size_t vqgnn_vq_segsum_workspace_bytes(int64_t B, int nbc, int M) {
  const int64_t n = B * nbc, S = static_cast<int64_t>(nbc) * M;
  return 4 * align256(static_cast<size_t>(n) * 4) + 3 * align256(static_cast<size_t>(S + 2) * 4) +
         align256(std::max(cub_sort_bytes(n, key_bits(S)), cub_scan_bytes(S + 1))) +
         align256(static_cast<size_t>(segsum_max_tasks(n, S)) * 20 * 4) + 256;
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
  const size_t as = align256(static_cast<size_t>(S + 2) * 4);
  int32_t* seg_start = reinterpret_cast<int32_t*>(p + 4 * an);
  int32_t* nsub = reinterpret_cast<int32_t*>(p + 4 * an + as);
  int32_t* task_off = reinterpret_cast<int32_t*>(p + 4 * an + 2 * as);
  void* cub_tmp = p + 4 * an + 3 * as;
  const int bits = key_bits(S);
  size_t cub_bytes = cub_sort_bytes(n, bits);
  const size_t cub_cap = align256(std::max(cub_bytes, cub_scan_bytes(S + 1)));
  float* part = reinterpret_cast<float*>(static_cast<char*>(cub_tmp) + cub_cap);
  const int max_tasks = static_cast<int>(segsum_max_tasks(n, S));
  const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, 16 * kNumSMs));
  segsum_keys_kernel<<<grid, 256, 0, s>>>(idx, n, nbc, M, keys_in, rows_in);
  VQ_LAUNCH_CHECK();
  VQ_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, rows_in, rows_out, n, 0, bits, s));
  count_launch(2);
  segsum_bounds_kernel<<<grid, 256, 0, s>>>(keys_out, n, static_cast<int>(S), seg_start);
  VQ_LAUNCH_CHECK();
  // sub-range tasks: nsub per segment, exclusive scan over S + 1 entries (the last one = total)
  segsum_nsub_kernel<<<ceil_div(S, 256), 256, 0, s>>>(seg_start, (int)S, nsub);
  VQ_LAUNCH_CHECK();
  VQ_CUDA(cudaMemsetAsync(nsub + S, 0, sizeof(int32_t), s));
  size_t scan_b = cub_scan_bytes(S + 1);
  VQ_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, scan_b, nsub, task_off, static_cast<int>(S + 1), s));
  count_launch(1);
  const int w_use = D + (g ? Dg : 0);
  const int nv = (w_use + 3) / 4;
  const int vec = (D == 4 && (!g || Dg == 4) && ldx % 4 == 0 && (!g || ldg % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (!g || (reinterpret_cast<uintptr_t>(g) & 15) == 0))
                      ? 1 : 0;
  // lanes per task: a warp when segments are long, 8 lanes when the average (branch, codeword) segment is short
  const bool narrow = n < 16 * S;
  const int sgrid = narrow ? (max_tasks + 31) / 32 : (max_tasks + 7) / 8;
#define VQ_SEGSUM(NV)                                                                                          \
  do {                                                                                                         \
    if (narrow)                                                                                                \
      segsum_kernel<NV, 8><<<sgrid, 256, 0, s>>>(x, ldx, g, ldg, scale, shift, rows_out, seg_start, task_off,  \
                                                 (int)S, max_tasks, nbc, M, D, g ? Dg : 0, Wp, vec, stats,     \
                                                 part);                                                        \
    else                                                                                                       \
      segsum_kernel<NV, 32><<<sgrid, 256, 0, s>>>(x, ldx, g, ldg, scale, shift, rows_out, seg_start, task_off, \
                                                  (int)S, max_tasks, nbc, M, D, g ? Dg : 0, Wp, vec, stats,    \
                                                  part);                                                       \
    segsum_combine_kernel<NV><<<ceil_div(S * (NV * 4), 256), 256, 0, s>>>(seg_start, task_off, (int)S, w_use,  \
                                                                         Wp, part, stats);                     \
    count_launch(1);                                                                                           \
  } while (0)
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
