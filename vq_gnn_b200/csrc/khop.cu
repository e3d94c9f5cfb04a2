// Mini-batch graph construction on the device from the RESIDENT normalised graph (CSR: rowptr int64, col int32,
// val fp32) and the batch's node ids -- the producers of the hot path's `batch_A` input, so that only node ids cross
// PCIe instead of an int64 COO of the batch graph.
//
//   v2  `_k_hop_subgraph` (vq_gnn_v2/dataloader.py:98-148) + `prepare_batch_input` (utils/misc.py:57-75):
//       subset = [batch nodes ; their out-of-batch 1-hop neighbours B' (ascending node id)], relabelled CSR of the
//       (B+B')^2 sub-adjacency -- train: every edge inside the subset, eval: the batch rows only.
//       khop_mark -> (host reads T) -> khop_count -> (host reads nnz) -> khop_fill
//   v1  `__collate__` tail (vq_gnn_v1/utils/dataloader.py:64-86): the tuple (A_BN (r, c, v), A_BB (r, c, v), A_NB_v)
//       in the reference's own COO format (int64 indices, row-sorted), ready for vqgnn_plan_v1_build.
//       collate_count -> (host reads nnz, nbb) -> collate_fill
//
// HBM-bound integer work: mask lookups in a node-position table, warp-per-row stable compaction (ballot prefix), CUB
// exclusive scans.  Everything is a pure function of the inputs (no order-dependent atomics).
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"

namespace vqgnn {

__global__ void khop_pos_kernel(const int64_t* __restrict__ node_idx, int B, int32_t* __restrict__ pos,
                                int32_t* __restrict__ batch_idx32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    const int64_t n = node_idx[i];
    pos[n] = i;
    if (batch_idx32) batch_idx32[i] = static_cast<int32_t>(n);
  }
}

// neighbours of batch rows that are not batch nodes: pos = -2 (idempotent plain stores)
__global__ void __launch_bounds__(256)
    khop_mark_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const int64_t* __restrict__ node_idx, int B, int32_t* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < B; i += gridDim.x * 8) {
    const int64_t n = node_idx[i];
    const int64_t e0 = rowptr[n], e1 = rowptr[n + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
      const int c = __ldg(col + e);
      if (pos[c] == -1) pos[c] = -2;
    }
  }
}

struct IsTail {
  __host__ __device__ int operator()(int32_t p) const { return p == -2 ? 1 : 0; }
};

// pos[n] = B + rank among the marked nodes (ascending id); tail_node[rank] = n; *T = number of marked nodes
__global__ void khop_assign_kernel(int32_t* __restrict__ pos, const int32_t* __restrict__ off, int64_t N, int B,
                                   int32_t* __restrict__ tail_node, int32_t* __restrict__ T) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const bool tail = pos[n] == -2;
  if (tail) {
    pos[n] = B + off[n];
    tail_node[off[n]] = static_cast<int32_t>(n);
  }
  if (n == N - 1) *T = off[n] + (tail ? 1 : 0);
}

// member[n >> 5] bit (n & 31) = node n is in the subset (pos[n] >= 0).  The membership test of every scanned edge then
// reads a 4 B word of an N/8-byte bitmap that mostly stays in the L1 (306 KB at the products shape) instead of a 32 B
// sector of the 10 MB position table from the L2 -- the count / fill kernels were bound by those gathers (1.5 GB of
// L2 -> SM traffic per pass); the position itself is fetched only for the ~35 % of the edges that are kept.
__global__ void __launch_bounds__(256)
    khop_bitmap_kernel(const int32_t* __restrict__ pos, int64_t N, uint32_t* __restrict__ member) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned m = __ballot_sync(0xffffffffu, n < N && pos[n] >= 0);
  if ((threadIdx.x & 31) == 0 && n < N) member[n >> 5] = m;
}
__device__ __forceinline__ bool khop_member(const uint32_t* __restrict__ member, int c) {
  return (__ldg(member + (c >> 5)) >> (c & 31)) & 1u;
}

// number of kept entries of row r of the batch graph (warp per row)
__global__ void __launch_bounds__(256)
    khop_count_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                      const int64_t* __restrict__ node_idx, const int32_t* __restrict__ tail_node, int B, int R,
                      const uint32_t* __restrict__ member, int32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += gridDim.x * 8) {
    const int64_t n = r < B ? node_idx[r] : tail_node[r - B];
    const int64_t e0 = rowptr[n], e1 = rowptr[n + 1];
    int c = 0;
    if (r < B) {
      c = static_cast<int>(e1 - e0);   // every neighbour of a batch node is in the subset
    } else {
      for (int64_t e = e0 + lane; e < e1; e += 32) c += khop_member(member, __ldg(col + e));
      c = __reduce_add_sync(0xffffffffu, c);
    }
    if (lane == 0) cnt[r] = c;
  }
}

__global__ void __launch_bounds__(256)
    khop_fill_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const float* __restrict__ val, const int64_t* __restrict__ node_idx,
                     const int32_t* __restrict__ tail_node, int B, int R, const int32_t* __restrict__ pos,
                     const uint32_t* __restrict__ member, const int32_t* __restrict__ out_rowptr,
                     int32_t* __restrict__ out_col, float* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += gridDim.x * 8) {
    const int64_t n = r < B ? node_idx[r] : tail_node[r - B];
    const int64_t e0 = rowptr[n], e1 = rowptr[n + 1];
    int base = out_rowptr[r];
    for (int64_t eb = e0; eb < e1; eb += 32) {   // stable compaction: ballot prefix keeps the stored order
      const int64_t e = eb + lane;
      int c = 0;
      bool keep = false;
      if (e < e1) c = __ldg(col + e), keep = khop_member(member, c);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int d = base + __popc(m & ((1u << lane) - 1));
        out_col[d] = pos[c];
        out_val[d] = __ldg(val + e);
      }
      base += __popc(m);
    }
  }
}

// ---- v1 collate -------------------------------------------------------------------------------------------------
// per batch row: degree and number of in-batch neighbours
__global__ void __launch_bounds__(256)
    collate_count_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                         const int64_t* __restrict__ node_idx, int B, const int32_t* __restrict__ pos,
                         int32_t* __restrict__ deg, int32_t* __restrict__ inb) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < B; i += gridDim.x * 8) {
    const int64_t n = node_idx[i];
    const int64_t e0 = rowptr[n], e1 = rowptr[n + 1];
    int c = 0;
    if (pos)
      for (int64_t e = e0 + lane; e < e1; e += 32) c += pos[__ldg(col + e)] >= 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) deg[i] = static_cast<int>(e1 - e0), inb[i] = c;
  }
}

__global__ void __launch_bounds__(256)
    collate_fill_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                        const float* __restrict__ val, const float* __restrict__ gdeg,
                        const float* __restrict__ gdeg_inv, const int64_t* __restrict__ node_idx, int B,
                        const int32_t* __restrict__ pos, const int32_t* __restrict__ off_bn,
                        const int32_t* __restrict__ off_bb, int64_t* __restrict__ bn_r, int64_t* __restrict__ bn_c,
                        float* __restrict__ bn_v, float* __restrict__ nb_v, int64_t* __restrict__ bb_r,
                        int64_t* __restrict__ bb_c, float* __restrict__ bb_v, float* __restrict__ deg_inv_out) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < B; i += gridDim.x * 8) {
    const int64_t n = node_idx[i];
    const int64_t e0 = rowptr[n], e1 = rowptr[n + 1];
    const float dn = gdeg ? __ldg(gdeg + n) : 0.f;
    if (lane == 0 && deg_inv_out) deg_inv_out[i] = __ldg(gdeg_inv + n);
    const int64_t o = off_bn[i];
    int bbase = bb_r ? off_bb[i] : 0;
    for (int64_t eb = e0; eb < e1; eb += 32) {
      const int64_t e = eb + lane;
      int p = -1, c = 0;
      float v = 0.f;
      if (e < e1) {
        c = __ldg(col + e), v = __ldg(val + e);
        const int64_t d = o + (e - e0);
        bn_r[d] = i, bn_c[d] = c, bn_v[d] = v;
        // A_NB_v = deg[batch node] * A_BN * deg_inv[neighbour]   (dataloader.py:77-78)
        if (nb_v) nb_v[d] = dn * v * __ldg(gdeg_inv + c);
        if (pos) p = pos[c];
      }
      if (bb_r) {
        const unsigned m = __ballot_sync(0xffffffffu, p >= 0);
        if (p >= 0) {
          const int d = bbase + __popc(m & ((1u << lane) - 1));
          bb_r[d] = i, bb_c[d] = p, bb_v[d] = v;
        }
        bbase += __popc(m);
      }
    }
  }
}

__global__ void khop_totals_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ cnt, int n,
                                   int32_t* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = n > 0 ? off[n - 1] + cnt[n - 1] : 0;
}

static inline size_t kh_al(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
static size_t scan_bytes(int64_t n) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                static_cast<int>(n));
  return b;
}
static int rows_grid(int64_t rows) { return static_cast<int>(std::min<int64_t>((rows + 7) / 8, 32 * kNumSMs)); }

}  // namespace vqgnn

using namespace vqgnn;

// pos [N] int32 lives in the caller's workspace and is SHARED by the three v2 calls (and the two v1 calls) of one batch.
extern "C" size_t vqgnn_khop_workspace_bytes(int64_t N, int64_t rows) {
  const int64_t m = std::max<int64_t>(N, rows) + 1;
  return kh_al(static_cast<size_t>(N) * 4) + 2 * kh_al(static_cast<size_t>(m) * 4) + kh_al(scan_bytes(m)) + 256;
}

namespace {
struct KhWs {
  int32_t* pos;
  int32_t* a;
  int32_t* b;
  void* scan;
  size_t scan_b;
};
KhWs kh_layout(void* ws, int64_t N, int64_t rows) {
  const int64_t m = std::max<int64_t>(N, rows) + 1;
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  KhWs w;
  w.pos = reinterpret_cast<int32_t*>(p);
  p += kh_al(static_cast<size_t>(N) * 4);
  w.a = reinterpret_cast<int32_t*>(p);
  p += kh_al(static_cast<size_t>(m) * 4);
  w.b = reinterpret_cast<int32_t*>(p);
  p += kh_al(static_cast<size_t>(m) * 4);
  w.scan = p;
  w.scan_b = scan_bytes(m);
  return w;
}
}  // namespace

extern "C" int vqgnn_khop_mark(const int64_t* rowptr, const int32_t* col, const int64_t* node_idx, int64_t B,
                               int64_t N, int32_t* batch_idx32, int32_t* tail_node, int32_t* T, void* ws,
                               void* stream) {
  VQ_CHECK_ARG(rowptr && col && node_idx && tail_node && T && ws && B > 0 && N > 0 && N < (1ll << 31) &&
                   B < (1ll << 31),
               "khop_mark: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KhWs w = kh_layout(ws, N, N);
  VQ_CUDA(cudaMemsetAsync(w.pos, 0xFF, sizeof(int32_t) * N, s));
  khop_pos_kernel<<<ceil_div(B, 256), 256, 0, s>>>(node_idx, (int)B, w.pos, batch_idx32);
  VQ_LAUNCH_CHECK();
  khop_mark_kernel<<<rows_grid(B), 256, 0, s>>>(rowptr, col, node_idx, (int)B, w.pos);
  VQ_LAUNCH_CHECK();
  cub::TransformInputIterator<int32_t, IsTail, const int32_t*> flags(w.pos, IsTail());
  size_t sb = w.scan_b;
  VQ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan, sb, flags, w.a, static_cast<int>(N), s));
  count_launch(1);
  khop_assign_kernel<<<ceil_div(N, 256), 256, 0, s>>>(w.pos, w.a, N, (int)B, tail_node, T);
  VQ_LAUNCH_CHECK();
  khop_bitmap_kernel<<<ceil_div(N, 256), 256, 0, s>>>(w.pos, N, reinterpret_cast<uint32_t*>(w.b));   // b: free in v2
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_khop_count(const int64_t* rowptr, const int32_t* col, const int64_t* node_idx,
                                const int32_t* tail_node, int64_t B, int64_t R, int64_t N, int32_t* out_rowptr,
                                int32_t* nnz, void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && col && node_idx && out_rowptr && nnz && ws && B > 0 && R >= B && (R == B || tail_node),
               "khop_count: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KhWs w = kh_layout(ws, N, std::max<int64_t>(N, R));
  KhWs wm = kh_layout(ws, N, N);   // the bitmap was written by vqgnn_khop_mark with this layout
  khop_count_kernel<<<rows_grid(R), 256, 0, s>>>(rowptr, col, node_idx, tail_node, (int)B, (int)R,
                                                 reinterpret_cast<const uint32_t*>(wm.b), w.a);
  VQ_LAUNCH_CHECK();
  size_t sb = w.scan_b;
  VQ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan, sb, w.a, out_rowptr, static_cast<int>(R), s));
  count_launch(1);
  khop_totals_kernel<<<1, 32, 0, s>>>(out_rowptr, w.a, (int)R, out_rowptr + R);
  VQ_LAUNCH_CHECK();
  VQ_CUDA(cudaMemcpyAsync(nnz, out_rowptr + R, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  return VQGNN_OK;
}

extern "C" int vqgnn_khop_fill(const int64_t* rowptr, const int32_t* col, const float* val, const int64_t* node_idx,
                               const int32_t* tail_node, int64_t B, int64_t R, int64_t N, const int32_t* out_rowptr,
                               int32_t* out_col, float* out_val, void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && col && val && node_idx && out_rowptr && ws && B > 0 && R >= B, "khop_fill: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KhWs w = kh_layout(ws, N, std::max<int64_t>(N, R));
  KhWs wm = kh_layout(ws, N, N);
  khop_fill_kernel<<<rows_grid(R), 256, 0, s>>>(rowptr, col, val, node_idx, tail_node, (int)B, (int)R, w.pos,
                                                reinterpret_cast<const uint32_t*>(wm.b), out_rowptr, out_col, out_val);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

// v1: off_bn / off_bb [B+1] (exclusive offsets; entry B = total), counts[0] = nnz(A_BN), counts[1] = nnz(A_BB)
extern "C" int vqgnn_collate_v1_count(const int64_t* rowptr, const int32_t* col, const int64_t* node_idx, int64_t B,
                                      int64_t N, int with_bb, int32_t* off_bn, int32_t* off_bb, int32_t* counts,
                                      void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && col && node_idx && off_bn && off_bb && counts && ws && B > 0 && N > 0 && N < (1ll << 31),
               "collate_v1_count: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KhWs w = kh_layout(ws, N, N);
  if (with_bb) {
    VQ_CUDA(cudaMemsetAsync(w.pos, 0xFF, sizeof(int32_t) * N, s));
    khop_pos_kernel<<<ceil_div(B, 256), 256, 0, s>>>(node_idx, (int)B, w.pos, nullptr);
    VQ_LAUNCH_CHECK();
  }
  collate_count_kernel<<<rows_grid(B), 256, 0, s>>>(rowptr, col, node_idx, (int)B, with_bb ? w.pos : nullptr, w.a, w.b);
  VQ_LAUNCH_CHECK();
  size_t sb = w.scan_b;
  VQ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan, sb, w.a, off_bn, static_cast<int>(B), s));
  sb = w.scan_b;
  VQ_CUDA(cub::DeviceScan::ExclusiveSum(w.scan, sb, w.b, off_bb, static_cast<int>(B), s));
  count_launch(2);
  khop_totals_kernel<<<1, 32, 0, s>>>(off_bn, w.a, (int)B, off_bn + B);
  khop_totals_kernel<<<1, 32, 0, s>>>(off_bb, w.b, (int)B, off_bb + B);
  VQ_LAUNCH_CHECK();
  VQ_CUDA(cudaMemcpyAsync(counts, off_bn + B, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  VQ_CUDA(cudaMemcpyAsync(counts + 1, off_bb + B, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  return VQGNN_OK;
}

extern "C" int vqgnn_collate_v1_fill(const int64_t* rowptr, const int32_t* col, const float* val, const float* gdeg,
                                     const float* gdeg_inv, const int64_t* node_idx, int64_t B, int64_t N,
                                     const int32_t* off_bn, const int32_t* off_bb, int64_t* bn_r, int64_t* bn_c,
                                     float* bn_v, float* nb_v, int64_t* bb_r, int64_t* bb_c, float* bb_v,
                                     float* deg_inv_out, void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && col && val && node_idx && off_bn && bn_r && bn_c && bn_v && ws && B > 0,
               "collate_v1_fill: bad arguments");
  VQ_CHECK_ARG(!nb_v || (gdeg && gdeg_inv), "collate_v1_fill: A_NB_v needs deg and deg_inv");
  VQ_CHECK_ARG(!deg_inv_out || gdeg_inv, "collate_v1_fill: deg_inv output needs the graph's deg_inv");
  VQ_CHECK_ARG(!bb_r || (bb_c && bb_v && off_bb), "collate_v1_fill: A_BB arrays missing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KhWs w = kh_layout(ws, N, N);
  collate_fill_kernel<<<rows_grid(B), 256, 0, s>>>(rowptr, col, val, gdeg, gdeg_inv, node_idx, (int)B,
                                                   bb_r ? w.pos : nullptr, off_bn, off_bb, bn_r, bn_c, bn_v, nb_v, bb_r,
                                                   bb_c, bb_v, deg_inv_out);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
