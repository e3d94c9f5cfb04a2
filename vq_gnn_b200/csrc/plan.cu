// Device-side construction of the v1 batch plan from the reference's batch tuple
//   (deg_inv[B], A_BN (r, c, v) int64/int64/fp32, A_BB (r, c, v) | None, A_NB_v | None, batch_idx[B])
// (vq_gnn_v1/utils/dataloader.py:64-86; `mapper` :144-192 consumes it once per branch per layer).
// Entirely asynchronous on the caller's stream -- no host synchronisation, no allocation: the number of tail
// entries stays on the device (t_count) and the consumers size their grids by the upper bound nnz(A_BN).
//   tail part     : entries of A_BN whose column is not a batch node, compacted per row (order inside a row is
//                   irrelevant to the sums), as a CSR over the B rows with global node ids
//   in-batch part : A_BB (+ its transpose for GCN = to_symmetric, + self loops deg_inv, doubled for GCN), as a CSR
//                   by row (forward) and by column (backward)
// HBM-bound integer work: warp-aggregated counting + cursor scatter, two small single-CTA scans.
//
// Determinism: the ORDER of the entries inside a CSR row decides the order of the fp32 sums in the message-passing
// kernels, so nothing here may depend on which warp wins an atomic: the tail part is a STABLE compaction of the
// row-sorted A_BN (two-level prefix sum, no cursors), and every cursor-scattered CSR (in-batch forward / transposed,
// v2 transposed) is finished by sorting each row segment by its unique key (seg_sort_*: shuffle / rank sort for
// short segments, shared- or global-memory bitonic network for long ones).  Counts are integer atomics (exact).
#include "common.cuh"

namespace vqgnn {

__global__ void plan_pos_kernel(const int64_t* __restrict__ batch_idx, int B, int32_t* __restrict__ pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) pos[batch_idx[i]] = i;
}

// one warp-aggregated atomic per distinct key among the flagged lanes; returns the lane's slot
__device__ __forceinline__ int warp_slot(int* counters, int key, bool flag) {
  const unsigned active = __ballot_sync(0xffffffffu, flag);
  int slot = 0;
  if (flag) {
    const unsigned peers = __match_any_sync(active, key);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counters + key, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    slot = base + __popc(peers & ((1u << lane) - 1));
  }
  return slot;
}

// pass 1 over A_BN: per-row tail counts (integer atomics) + per-tile tail counts; tile = kTailTile consecutive entries
constexpr int kTailTile = 2048;

__global__ void __launch_bounds__(256)
    plan_tail_count_kernel(const int64_t* __restrict__ r, const int64_t* __restrict__ c, int64_t nnz,
                           const int32_t* __restrict__ pos, int has_bb, int* __restrict__ row_cnt,
                           int* __restrict__ tile_cnt) {
  __shared__ int tot;
  if (threadIdx.x == 0) tot = 0;
  __syncthreads();
  const int64_t e0 = static_cast<int64_t>(blockIdx.x) * kTailTile;
  int mine = 0;
  for (int i = threadIdx.x; i < kTailTile; i += 256) {   // warps stay converged: kTailTile % 256 == 0
    const int64_t e = e0 + i;
    bool tail = false;
    int row = 0;
    if (e < nnz) {
      row = static_cast<int>(r[e]);
      tail = !has_bb || __ldg(pos + c[e]) < 0;
    }
    warp_slot(row_cnt, row, tail);
    mine += tail;
  }
  mine = static_cast<int>(warp_sum(static_cast<float>(mine)));   // <= 2048: exact in fp32
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&tot, mine);
  __syncthreads();
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
}

// pass 2: stable compaction -- entry e lands at tile_off[tile] + (number of earlier tail entries of its tile), so the
// output keeps A_BN's order (row-sorted input => row-grouped output matching t_rowptr)
__global__ void __launch_bounds__(256)
    plan_tail_scatter_kernel(const int64_t* __restrict__ r, const int64_t* __restrict__ c,
                             const float* __restrict__ v, const float* __restrict__ rv, int64_t nnz,
                             const int32_t* __restrict__ pos, int has_bb, const int32_t* __restrict__ tile_off,
                             int32_t* __restrict__ t_node, float* __restrict__ t_val, float* __restrict__ t_rval) {
  __shared__ int wtot[8];
  __shared__ int base_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_sh = __ldg(tile_off + blockIdx.x);
  __syncthreads();
  const int64_t e0 = static_cast<int64_t>(blockIdx.x) * kTailTile;
  for (int i0 = 0; i0 < kTailTile; i0 += 256) {
    const int64_t e = e0 + i0 + threadIdx.x;
    bool tail = false;
    int64_t col = 0;
    if (e < nnz) {
      col = c[e];
      tail = !has_bb || __ldg(pos + col) < 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, tail);
    if (lane == 0) wtot[warp] = __popc(m);
    __syncthreads();
    int before = base_sh;
    for (int w = 0; w < warp; ++w) before += wtot[w];
    if (tail) {
      const int64_t dst = before + __popc(m & ((1u << lane) - 1));
      t_node[dst] = static_cast<int32_t>(col);
      t_val[dst] = v[e];
      t_rval[dst] = rv ? rv[e] : 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += wtot[w];
      base_sh += t;
    }
    __syncthreads();
  }
}

// ---- per-segment sort of (key, val) pairs by key (keys are unique inside a segment): src -> dst -----------------
// short segments: one warp each
__global__ void __launch_bounds__(256)
    seg_sort_short_kernel(const int32_t* __restrict__ ptr, int nseg, const int32_t* __restrict__ ksrc,
                          const float* __restrict__ vsrc, int32_t* __restrict__ kdst, float* __restrict__ vdst,
                          int max_short) {
  const int lane = threadIdx.x & 31;
  for (int sg = blockIdx.x * 8 + (threadIdx.x >> 5); sg < nseg; sg += gridDim.x * 8) {
    const int s0 = __ldg(ptr + sg), n = __ldg(ptr + sg + 1) - s0;
    if (n <= 0 || n > max_short) continue;
    if (n <= 32) {
      const int k = lane < n ? ksrc[s0 + lane] : 0x7fffffff;
      const float v = lane < n ? vsrc[s0 + lane] : 0.f;
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const int kj = __shfl_sync(0xffffffffu, k, j);
        rank += (kj < k) || (kj == k && j < lane);
      }
      if (lane < n) kdst[s0 + rank] = k, vdst[s0 + rank] = v;
    } else {
      for (int i = lane; i < n; i += 32) {
        const int k = ksrc[s0 + i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
          const int kj = __ldg(ksrc + s0 + j);
          rank += (kj < k) || (kj == k && j < i);
        }
        kdst[s0 + rank] = k, vdst[s0 + rank] = vsrc[s0 + i];
      }
    }
  }
}

// long segments: one CTA each, normalised bitonic network (every compare-exchange moves the smaller key down, so the
// virtual +inf padding above n never moves); in shared memory up to kSegSmem pairs, in place in dst beyond that
constexpr int kSegSmem = 8192;

__device__ __forceinline__ void seg_cmpx(int32_t* k, float* v, int i, int l, int n) {
  if (l < n) {
    const int ki = k[i], kl = k[l];
    if (kl < ki) {
      const float vi = v[i];
      k[i] = kl, v[i] = v[l];
      k[l] = ki, v[l] = vi;
    }
  }
}

__global__ void __launch_bounds__(1024)
    seg_sort_long_kernel(const int32_t* __restrict__ ptr, int nseg, const int32_t* __restrict__ ksrc,
                         const float* __restrict__ vsrc, int32_t* __restrict__ kdst, float* __restrict__ vdst,
                         int max_short) {
  extern __shared__ __align__(16) unsigned char seg_smem[];
  int32_t* ks = reinterpret_cast<int32_t*>(seg_smem);
  float* vs = reinterpret_cast<float*>(seg_smem + sizeof(int32_t) * kSegSmem);
  for (int sg = blockIdx.x; sg < nseg; sg += gridDim.x) {
    const int s0 = __ldg(ptr + sg), n = __ldg(ptr + sg + 1) - s0;
    if (n <= max_short) continue;
    const bool in_smem = n <= kSegSmem;
    int32_t* k = in_smem ? ks : kdst + s0;
    float* v = in_smem ? vs : vdst + s0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) k[i] = ksrc[s0 + i], v[i] = vsrc[s0 + i];
    __syncthreads();
    int np = 1;
    while (np < n) np <<= 1;
    for (int kk = 2; kk <= np; kk <<= 1) {
      for (int t = threadIdx.x; t < np / 2; t += blockDim.x) {   // flip step: i <-> i ^ (kk - 1)
        const int blk = t / (kk / 2), o = t - blk * (kk / 2);
        const int i = blk * kk + o;
        seg_cmpx(k, v, i, i ^ (kk - 1), n);
      }
      __syncthreads();
      for (int j = kk >> 2; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < np / 2; t += blockDim.x) {
          const int i = ((t / j) * 2 * j) + (t % j);
          seg_cmpx(k, v, i, i + j, n);
        }
        __syncthreads();
      }
    }
    if (in_smem)
      for (int i = threadIdx.x; i < n; i += blockDim.x) kdst[s0 + i] = ks[i], vdst[s0 + i] = vs[i];
  }
}

constexpr int kSegShortMax = 128;

static int seg_sort(const int32_t* ptr, int nseg, const int32_t* ksrc, const float* vsrc, int32_t* kdst, float* vdst,
                    cudaStream_t s) {
  if (nseg <= 0) return VQGNN_OK;
  const int g1 = std::min((nseg + 7) / 8, 32 * kNumSMs);
  seg_sort_short_kernel<<<g1, 256, 0, s>>>(ptr, nseg, ksrc, vsrc, kdst, vdst, kSegShortMax);
  VQ_LAUNCH_CHECK();
  const size_t smem = static_cast<size_t>(kSegSmem) * 8;
  VQ_CUDA(cudaFuncSetAttribute(seg_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seg_sort_long_kernel<<<std::min(nseg, 4 * kNumSMs), 1024, smem, s>>>(ptr, nseg, ksrc, vsrc, kdst, vdst, kSegShortMax);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

// in-batch entries: A_BB (+ transposes) + self loops.  KEY 0: by row (forward CSR), 1: by column (transposed)
template <bool SCATTER>
__global__ void __launch_bounds__(256)
    plan_inb_kernel(const int64_t* __restrict__ br, const int64_t* __restrict__ bc, const float* __restrict__ bv,
                    int64_t nbb, const float* __restrict__ deg_inv, int B, int symmetric, int self_loops,
                    int* __restrict__ cnt_row, int* __restrict__ cnt_col, const int32_t* __restrict__ i_rowptr,
                    int32_t* __restrict__ i_col, float* __restrict__ i_val, const int32_t* __restrict__ b_rowptr,
                    int32_t* __restrict__ b_row, float* __restrict__ b_val) {
  const int64_t total = nbb * (symmetric ? 2 : 1) + (self_loops ? B : 0);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    int row, col;
    float val;
    if (t < nbb) {
      row = static_cast<int>(br[t]), col = static_cast<int>(bc[t]), val = bv[t];
    } else if (symmetric && t < 2 * nbb) {
      row = static_cast<int>(bc[t - nbb]), col = static_cast<int>(br[t - nbb]), val = bv[t - nbb];
    } else {
      row = col = static_cast<int>(t - nbb * (symmetric ? 2 : 1));
      val = deg_inv[row] * (symmetric ? 2.f : 1.f);  // to_symmetric() doubles the diagonal (dataloader.py:189-190)
    }
    const int s_row = atomicAdd(cnt_row + row, 1);
    const int s_col = atomicAdd(cnt_col + col, 1);
    if (SCATTER) {
      const int d0 = i_rowptr[row] + s_row;
      i_col[d0] = col, i_val[d0] = val;
      const int d1 = b_rowptr[col] + s_col;
      b_row[d1] = row, b_val[d1] = val;
    }
  }
}

// exclusive scan of cnt[0..n) into ptr[0..n] by one CTA; zeroes cnt for reuse as cursors; optional total out
__global__ void __launch_bounds__(1024) plan_scan_kernel(int* __restrict__ cnt, int n, int32_t* __restrict__ ptr,
                                                         int32_t* __restrict__ total) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int x = i < n ? cnt[i] : 0;
    int s = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + s - x;
    if (i < n) ptr[i] = before, cnt[i] = 0;
    __syncthreads();
    if (tid == 1023) carry = before + x;
    __syncthreads();
  }
  if (tid == 0) {
    ptr[n] = carry;
    if (total) *total = carry;
  }
}

// chunk_row for a CSR whose nnz is only known on the device
__global__ void plan_chunk_rows_dev_kernel(const int32_t* __restrict__ rowptr, int R, const int32_t* __restrict__ d_nnz,
                                           int chunk, int max_chunks, int32_t* __restrict__ chunk_row) {
  const int cI = blockIdx.x * blockDim.x + threadIdx.x;
  if (cI >= max_chunks) return;
  const int64_t e = static_cast<int64_t>(cI) * chunk;
  if (e >= *d_nnz) {
    chunk_row[cI] = R - 1;
    return;
  }
  int lo = 0, hi = R;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(rowptr + mid) <= e) lo = mid;
    else hi = mid;
  }
  chunk_row[cI] = lo;
}

// Transposed CSR restricted to columns < B (the backward of the v2 forward: vq_gnn_v2 autograd of adj @ x, batch rows
// only).  One warp per forward row; pass 0 counts per column, pass 1 scatters (row, value) through per-column cursors.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
    plan_transpose_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                          const float* __restrict__ val, int R, int B, int* __restrict__ counters,
                          const int32_t* __restrict__ browptr, int32_t* __restrict__ brow,
                          float* __restrict__ bval) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
    const int e0 = __ldg(rowptr + r), e1 = __ldg(rowptr + r + 1);
    for (int e = e0 + lane; e < e1; e += 32) {
      const int c = __ldg(col + e);
      if (c < B) {
        const int slot = atomicAdd(counters + c, 1);
        if (SCATTER) {
          const int d = __ldg(browptr + c) + slot;
          brow[d] = r;
          bval[d] = __ldg(val + e);
        }
      }
    }
  }
}

static int plan_grid(int64_t n) { return static_cast<int>(std::min<int64_t>((n + 255) / 256, 16 * kNumSMs)); }

}  // namespace vqgnn

using namespace vqgnn;

static inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

extern "C" size_t vqgnn_plan_v1_workspace_bytes(int64_t N, int64_t B, int64_t nnz, int64_t nin) {
  const size_t tiles = static_cast<size_t>((nnz + kTailTile - 1) / kTailTile) + 1;
  return al256(static_cast<size_t>(N) * 4) + al256(static_cast<size_t>(B) * 4 * 3) + 2 * al256(tiles * 4) +
         4 * al256(static_cast<size_t>(nin > 0 ? nin : 1) * 4) + 256;
}

extern "C" int vqgnn_plan_v1_build(const int64_t* r, const int64_t* c, const float* v, const float* rv, int64_t nnz,
                                   const int64_t* bb_r, const int64_t* bb_c, const float* bb_v, int64_t nbb,
                                   const int64_t* batch_idx, const float* deg_inv, int64_t B, int64_t N,
                                   int symmetric, int self_loops, int tail_chunk, int chunk, int32_t* t_rowptr,
                                   int32_t* t_node, float* t_val, float* t_rval, int32_t* t_count,
                                   int32_t* t_chunk_row, int32_t* i_rowptr, int32_t* i_col, float* i_val,
                                   int32_t* i_chunk_row, int32_t* b_rowptr, int32_t* b_row, float* b_val,
                                   int32_t* b_chunk_row, void* ws, void* stream) {
  VQ_CHECK_ARG(r && c && v && batch_idx && t_rowptr && t_node && t_val && t_rval && t_count && t_chunk_row && i_rowptr &&
                   b_rowptr && ws,
               "plan_v1_build: null argument");
  VQ_CHECK_ARG(B > 0 && B < (1ll << 31) && N > 0 && N < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31) && nbb >= 0,
               "plan_v1_build: sizes must fit int32");
  VQ_CHECK_ARG(nbb == 0 || (bb_r && bb_c && bb_v), "plan_v1_build: A_BB arrays missing");
  VQ_CHECK_ARG(!self_loops || deg_inv, "plan_v1_build: self loops need deg_inv");
  VQ_CHECK_ARG(tail_chunk > 0 && tail_chunk % 32 == 0 && chunk > 0 && chunk % 32 == 0, "plan_v1_build: bad chunk sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t nin = nbb * (symmetric ? 2 : 1) + (self_loops ? B : 0);
  const int tiles = static_cast<int>((nnz + kTailTile - 1) / kTailTile);
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  int32_t* pos = reinterpret_cast<int32_t*>(p);
  p += al256(static_cast<size_t>(N) * 4);
  int* cnt_t = reinterpret_cast<int*>(p);
  int* cnt_r = cnt_t + B;
  int* cnt_c = cnt_r + B;
  p += al256(static_cast<size_t>(B) * 4 * 3);
  int* tile_cnt = reinterpret_cast<int*>(p);
  p += al256(static_cast<size_t>(tiles + 1) * 4);
  int32_t* tile_off = reinterpret_cast<int32_t*>(p);
  p += al256(static_cast<size_t>(tiles + 1) * 4);
  const size_t an = al256(static_cast<size_t>(nin > 0 ? nin : 1) * 4);
  int32_t* tmp_ic = reinterpret_cast<int32_t*>(p);
  float* tmp_iv = reinterpret_cast<float*>(p + an);
  int32_t* tmp_br = reinterpret_cast<int32_t*>(p + 2 * an);
  float* tmp_bv = reinterpret_cast<float*>(p + 3 * an);
  const int has_bb = bb_r != nullptr;   // A_BB == None (eval / no recovery): every neighbour goes through its codeword
  if (has_bb) {
    VQ_CUDA(cudaMemsetAsync(pos, 0xFF, sizeof(int32_t) * N, s));
    plan_pos_kernel<<<ceil_div(B, 256), 256, 0, s>>>(batch_idx, (int)B, pos);
    VQ_LAUNCH_CHECK();
  }
  VQ_CUDA(cudaMemsetAsync(cnt_t, 0, sizeof(int) * B * 3, s));
  // ---- tail part: per-row counts -> t_rowptr; per-tile counts -> stable positions
  if (nnz > 0) {
    plan_tail_count_kernel<<<tiles, 256, 0, s>>>(r, c, nnz, pos, has_bb, cnt_t, tile_cnt);
    VQ_LAUNCH_CHECK();
  }
  plan_scan_kernel<<<1, 1024, 0, s>>>(cnt_t, (int)B, t_rowptr, t_count);
  VQ_LAUNCH_CHECK();
  if (nnz > 0) {
    plan_scan_kernel<<<1, 1024, 0, s>>>(tile_cnt, tiles, tile_off, nullptr);
    VQ_LAUNCH_CHECK();
    plan_tail_scatter_kernel<<<tiles, 256, 0, s>>>(r, c, v, rv, nnz, pos, has_bb, tile_off, t_node, t_val, t_rval);
    VQ_LAUNCH_CHECK();
  }
  const int max_chunks = static_cast<int>((nnz + tail_chunk - 1) / tail_chunk);
  if (max_chunks > 0) {
    plan_chunk_rows_dev_kernel<<<ceil_div(max_chunks, 256), 256, 0, s>>>(t_rowptr, (int)B, t_count, tail_chunk,
                                                                        max_chunks, t_chunk_row);
    VQ_LAUNCH_CHECK();
  }
  // ---- in-batch part (forward by row, backward by column): cursor scatter into scratch, then per-row sort
  if (nin > 0) {
    VQ_CHECK_ARG(i_col && i_val && b_row && b_val, "plan_v1_build: in-batch outputs missing");
    const int igrid = plan_grid(nin);
    plan_inb_kernel<false><<<igrid, 256, 0, s>>>(bb_r, bb_c, bb_v, nbb, deg_inv, (int)B, symmetric, self_loops, cnt_r,
                                                 cnt_c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    VQ_LAUNCH_CHECK();
  }
  plan_scan_kernel<<<1, 1024, 0, s>>>(cnt_r, (int)B, i_rowptr, nullptr);
  VQ_LAUNCH_CHECK();
  plan_scan_kernel<<<1, 1024, 0, s>>>(cnt_c, (int)B, b_rowptr, nullptr);
  VQ_LAUNCH_CHECK();
  if (nin > 0) {
    plan_inb_kernel<true><<<plan_grid(nin), 256, 0, s>>>(bb_r, bb_c, bb_v, nbb, deg_inv, (int)B, symmetric, self_loops,
                                                         cnt_r, cnt_c, i_rowptr, tmp_ic, tmp_iv, b_rowptr, tmp_br,
                                                         tmp_bv);
    VQ_LAUNCH_CHECK();
    if (int rc = seg_sort(i_rowptr, (int)B, tmp_ic, tmp_iv, i_col, i_val, s)) return rc;
    if (int rc = seg_sort(b_rowptr, (int)B, tmp_br, tmp_bv, b_row, b_val, s)) return rc;
    if (int rc = vqgnn_mp_chunk_rows(i_rowptr, B, nin, chunk, i_chunk_row, stream)) return rc;
    if (int rc = vqgnn_mp_chunk_rows(b_rowptr, B, nin, chunk, b_chunk_row, stream)) return rc;
  }
  return VQGNN_OK;
}

extern "C" size_t vqgnn_csr_transpose_workspace_bytes(int64_t B, int64_t nnz) {
  return al256(static_cast<size_t>(B) * 4) + 2 * al256(static_cast<size_t>(nnz > 0 ? nnz : 1) * 4) + 256;
}

extern "C" int vqgnn_csr_transpose_lt(const int32_t* rowptr, const int32_t* col, const float* val, int64_t R,
                                      int64_t nnz, int64_t B, int32_t* browptr, int32_t* brow, float* bval,
                                      int32_t* count, void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && browptr && count && ws && R > 0 && B > 0 && B <= R && R < (1ll << 31) && nnz >= 0 &&
                   nnz < (1ll << 31),
               "csr_transpose_lt: bad arguments");
  VQ_CHECK_ARG(nnz == 0 || (col && val && brow && bval), "csr_transpose_lt: null arrays");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  int* cnt = reinterpret_cast<int*>(p);
  p += al256(static_cast<size_t>(B) * 4);
  const size_t an = al256(static_cast<size_t>(nnz > 0 ? nnz : 1) * 4);
  int32_t* tmp_r = reinterpret_cast<int32_t*>(p);
  float* tmp_v = reinterpret_cast<float*>(p + an);
  VQ_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * B, s));
  const int grid = static_cast<int>(std::min<int64_t>((R + 7) / 8, 16 * kNumSMs));
  if (nnz > 0) {
    plan_transpose_kernel<false><<<grid, 256, 0, s>>>(rowptr, col, val, (int)R, (int)B, cnt, nullptr, nullptr, nullptr);
    VQ_LAUNCH_CHECK();
  }
  plan_scan_kernel<<<1, 1024, 0, s>>>(cnt, (int)B, browptr, count);
  VQ_LAUNCH_CHECK();
  if (nnz > 0) {
    plan_transpose_kernel<true><<<grid, 256, 0, s>>>(rowptr, col, val, (int)R, (int)B, cnt, browptr, tmp_r, tmp_v);
    VQ_LAUNCH_CHECK();
    // the cursor scatter leaves every column's entries in atomic order: sort them by source row (unique)
    if (int rc = seg_sort(browptr, (int)B, tmp_r, tmp_v, brow, bval, s)) return rc;
  }
  return VQGNN_OK;
}
