// VQ-approximated message passing, GCN / SAGE-Mean: nnz-balanced gather-SpMM over the batch plan.
//
// Work unit = one warp x (chunk of `chunk` consecutive CSR entries) x (slab of 32*VEC columns).  Chunks ignore
// row boundaries, so a power-law hub row (10^4..10^5 entries in the Reddit-shaped batch) is spread over
// hundreds of warps instead of serialising one (the first version of this kernel was row-per-warp and spent
// 30 ms in its longest row).  Rows that lie wholly inside a chunk are stored directly; a row cut by one chunk
// boundary is accumulated with two vector REDs (red.global.add.v4.f32) into a pre-zeroed output (two additions onto
// zero commute, so the result is order-independent); the pieces of a row spanning three or more chunks go to a piece
// buffer and are summed in chunk order by mp_fixup_kernel.  Results are bit-identical from run to run.
//
// In-batch neighbours read dense rows (coalesced 16 B per lane); out-of-batch neighbours read the node's code
// row (2 B per lane, contiguous over the branches of the slab) and gather their codeword from the L2-resident
// codebook -- one 32 B sector per (entry, branch), fetched with a single 256-bit load when both the feature
// and the gradient half are needed.  HBM/L2-bound integer+float gather work: no tensor cores.
//
// Reference maths: vq_gnn_v2/models.py:161-198, vq_gnn_v2/convs.py:65-101,
// vq_gnn_v1/models.py:170-223 + vq_gnn_v1/utils/dataloader.py:144-192.
#include <algorithm>
#include <cstdlib>

#include "mp_common.cuh"
#include "mp_rows.cuh"

namespace vqgnn {

// chunk c starts at entry c*chunk; chunk_row[c] = the row that entry belongs to
__global__ void mp_chunk_rows_kernel(const int32_t* __restrict__ rowptr, int64_t R, int64_t nnz, int chunk,
                                     int n_chunks, int32_t* __restrict__ chunk_row) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const int64_t e = static_cast<int64_t>(c) * chunk;
  int64_t lo = 0, hi = R;  // largest r with rowptr[r] <= e  (rowptr[0] = 0 <= e < nnz = rowptr[R])
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(rowptr + mid) <= e) lo = mid;
    else hi = mid;
  }
  chunk_row[c] = static_cast<int32_t>(lo);
}

template <int VEC, bool HAS_GQ, bool WIDE>
__global__ void __launch_bounds__(kMpWarps * 32)
    mp_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                  const float* __restrict__ val, const float* __restrict__ rval,
                  const int32_t* __restrict__ chunk_row, int n_chunks, int chunk, int nnz, int64_t R, int B,
                  const float* __restrict__ x, int64_t ldx, Codebook cb, int C, int nslab, float feat_scale,
                  float info_scale, float* __restrict__ y, int64_t ldy, float* __restrict__ gq, int64_t ldgq,
                  float* __restrict__ info, double* ws_part, unsigned int* ws_count, float* __restrict__ py,
                  float* __restrict__ pgq) {
  const int lane = threadIdx.x & 31;
  const MpTask t = mp_task<VEC>(chunk_row, n_chunks, chunk, nnz, nslab, C, cb.D);
  float fpart = 0.f;
  if (t.valid) {
    const int c0 = t.c0, k = t.k, off = t.off;
    const int64_t ch = t.eb / chunk;
    float acc[VEC], gqa[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f, gqa[i] = 0.f;
    auto body = [&](const EntryGroup& g) {
      if (t.active) gather_accumulate<VEC, HAS_GQ, WIDE>(g, B, x, ldx, cb, 0, feat_scale, c0, k, off, acc, gqa);
    };
    auto body_dense = [&](const EntryGroup& g) {
      if constexpr (!HAS_GQ && !WIDE) {
        if (t.active)
          gather_accumulate<VEC, false, false, true>(g, B, x, ldx, cb, 0, feat_scale, c0, k, off, acc, gqa);
      }
    };
    auto flush = [&](int r, bool whole, int rs, int re) {
      if (t.active) {
        if (r < B) {
          const int kind = piece_kind(whole, rs, re, t.eb, chunk);
          const int64_t poff = (ch * 2 + (kind == kPieceHubStart ? 1 : 0)) * C + c0;
          float* yp = y + static_cast<int64_t>(r) * ldy + c0;
          if (kind == kPieceWhole) st_vec<VEC>(yp, acc);
          else if (kind == kPieceRed) red_vec<VEC>(yp, acc);
          else st_vec<VEC>(py + poff, acc);
          if (HAS_GQ) {
            if (gq) {
              float* gp = gq + static_cast<int64_t>(r) * ldgq + c0;
              if (kind == kPieceWhole) st_vec<VEC>(gp, gqa);
              else if (kind == kPieceRed) red_vec<VEC>(gp, gqa);
              else st_vec<VEC>(pgq + poff, gqa);
            }
            if (info) {  // v1: <x[r], gq[r]>  (vq_gnn_v1/models.py:223 rewritten row-wise)
              float xr[VEC];
              ld_vec<VEC>(x + static_cast<int64_t>(r) * ldx + c0, xr);
#pragma unroll
              for (int i = 0; i < VEC; ++i) fpart = fmaf(xr[i], gqa[i], fpart);
            }
          }
        } else if (info) {  // v2: <Y[r], Gq[r]> with Gq the node's own gradient codeword (models.py:198)
          const int node = cb.tail_node ? __ldg(cb.tail_node + (r - B)) : (r - B);
          const int code = __ldg(cb.codes + static_cast<int64_t>(node) * cb.nb + k);
          float gv[VEC];
          ld_vec<VEC>(cb.O + (static_cast<int64_t>(k) * cb.M + code) * cb.Wp + cb.D + off, gv);
#pragma unroll
          for (int i = 0; i < VEC; ++i) fpart = fmaf(acc[i], gv[i], fpart);
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = 0.f, gqa[i] = 0.f;
    };
    PlainWeights pol;
    if (!HAS_GQ && !WIDE && cb.tail_feat != nullptr)   // uniform over the grid
      walk_rows<HAS_GQ>(t.eb, t.ee, t.row0, R, rowptr, col, val, rval, B, cb.tail_node, lane, pol, body_dense, flush);
    else
      walk_rows<HAS_GQ>(t.eb, t.ee, t.row0, R, rowptr, col, val, rval, B, cb.tail_node, lane, pol, body, flush);
  }
  if (info) info_reduce_ordered(static_cast<double>(fpart), ws_part, ws_count, info_scale, info);
}

// Dense-tail forward with a deep asynchronous gather pipeline (the v2 forward of large batch graphs, where the
// kernel is bound by the LATENCY of its 512 B row gathers: ncu showed 34 % occupancy, long-scoreboard stalls and
// < 30 % of the L2 / DRAM throughput).  Same work partition and the same in-order accumulation as mp_fwd_kernel,
// but every entry's row -- x[c] for a batch column, tail_feat[c - B] otherwise -- is copied global -> shared with
// cp.async (16 B per lane, no registers held) kAsyncDepth entries ahead of its use, so each warp keeps kAsyncDepth
// rows in flight instead of kMpUnroll (depth 8: 32 KB of shared memory per CTA, six CTAs = 48 warps per SM).
constexpr int kAsyncDepthDefault = 8;   // measured at the products shape: depth 8 -> 1.36 ms, 12 -> 1.52, 16 -> 1.72 (fewer CTAs per SM)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int kAsyncDepth>
__global__ void __launch_bounds__(kMpWarps * 32)
    mp_fwd_async_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                        const float* __restrict__ val, const int32_t* __restrict__ chunk_row, int n_chunks,
                        int chunk, int nnz, int64_t R, int B, const float* __restrict__ x, int64_t ldx, Codebook cb,
                        int C, int nslab, float info_scale, float* __restrict__ y, int64_t ldy,
                        float* __restrict__ info, double* ws_part, unsigned int* ws_count, float* __restrict__ py) {
  extern __shared__ __align__(16) unsigned char async_smem[];   // [kMpWarps][kAsyncDepth][32 lanes] float4
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MpTask t = mp_task<4>(chunk_row, n_chunks, chunk, nnz, nslab, C, cb.D);
  float fpart = 0.f;
  if (t.valid) {
    const int c0 = t.c0;
    const int64_t ch = t.eb / chunk;
    float4* ring = reinterpret_cast<float4*>(async_smem) + (warp * kAsyncDepth) * 32 + lane;
    const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int r = t.row0, rbase = r;
    int rp_l = __ldg(rowptr + min(static_cast<int64_t>(rbase) + lane, R));
    int rs = __shfl_sync(0xffffffffu, rp_l, 0), re = __shfl_sync(0xffffffffu, rp_l, 1);

    auto flush = [&](bool whole) {
      if (t.active) {
        if (r < B) {
          const int kind = piece_kind(whole, rs, re, t.eb, chunk);
          float* yp = y + static_cast<int64_t>(r) * ldy + c0;
          if (kind == kPieceWhole) st_vec<4>(yp, acc);
          else if (kind == kPieceRed) red_vec<4>(yp, acc);
          else st_vec<4>(py + (ch * 2 + (kind == kPieceHubStart ? 1 : 0)) * C + c0, acc);
        } else if (info) {  // v2: <Y[r], Gq[r]> with Gq the node's own gradient codeword (models.py:198)
          const int node = cb.tail_node ? __ldg(cb.tail_node + (r - B)) : (r - B);
          const int code = __ldg(cb.codes + static_cast<int64_t>(node) * cb.nb + t.k);
          float gv[4];
          ld_vec<4>(cb.O + (static_cast<int64_t>(t.k) * cb.M + code) * cb.Wp + cb.D + t.off, gv);
#pragma unroll
          for (int i = 0; i < 4; ++i) fpart = fmaf(acc[i], gv[i], fpart);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = 0.f;
    };

    // entry e of the chunk lives in lane (e - t.eb) % 32 of batch (e - t.eb) / 32; two batches are held in registers
    int c_cur = -1, c_nxt = -1;
    float v_cur = 0.f, v_nxt = 0.f;
    auto load_batch = [&](int bb, int& c_l, float& v_l) {
      const int e = bb + lane;
      c_l = -1, v_l = 0.f;
      if (e < t.ee) c_l = __ldg(col + e), v_l = __ldg(val + e);
    };
    load_batch(t.eb, c_cur, v_cur);
    load_batch(t.eb + 32, c_nxt, v_nxt);
    // issue the copy of entry `e` (its column id is in c_cur / c_nxt of lane (e - eb) % 32) into ring slot e % depth
    auto issue = [&](int e, int bb_cur) {
      if (e < t.ee) {
        const int src_lane = (e - t.eb) & 31;
        const int c = __shfl_sync(0xffffffffu, e < bb_cur + 32 ? c_cur : c_nxt, src_lane);
        if (t.active) {
          const float* p = c >= B ? cb.tail_feat + static_cast<int64_t>(c - B) * cb.ld_tail + c0
                                  : x + static_cast<int64_t>(c) * ldx + c0;
          cp_async16(ring_s + ((e - t.eb) % kAsyncDepth) * 32 * 16, p);
        }
      }
      cp_async_commit();   // one group per entry, also past the end: the wait below counts groups
    };
    for (int e = t.eb; e < t.eb + kAsyncDepth; ++e) issue(e, t.eb);

    bool pending = false;
    for (int bb = t.eb; bb < t.ee; bb += 32) {
      const int bend = min(bb + 32, t.ee);
      for (int e = bb; e < bend; ++e) {
        const float v = __shfl_sync(0xffffffffu, v_cur, e - bb);
        cp_async_wait<kAsyncDepth - 1>();
        if (t.active) {
          const float4 a = ring[((e - t.eb) % kAsyncDepth) * 32];
          acc[0] = fmaf(v, a.x, acc[0]), acc[1] = fmaf(v, a.y, acc[1]);
          acc[2] = fmaf(v, a.z, acc[2]), acc[3] = fmaf(v, a.w, acc[3]);
        }
        issue(e + kAsyncDepth, bb);
        pending = true;
        if (e + 1 == re) {   // row r complete
          flush(rs >= t.eb);
          pending = false;
          if (e + 1 < t.ee) {
            do {  // next non-empty row
              ++r;
              if (r - rbase >= 31) {
                rbase = r;
                rp_l = __ldg(rowptr + min(static_cast<int64_t>(rbase) + lane, R));
              }
              rs = __shfl_sync(0xffffffffu, rp_l, r - rbase);
              re = __shfl_sync(0xffffffffu, rp_l, r - rbase + 1);
            } while (re <= e + 1);
          }
        }
      }
      c_cur = c_nxt, v_cur = v_nxt;
      load_batch(bb + 64, c_nxt, v_nxt);
    }
    if (pending) flush(false);
    cp_async_wait<0>();
  }
  if (info) info_reduce_ordered(static_cast<double>(fpart), ws_part, ws_count, info_scale, info);
}

// Sums the pieces of every row that spans >= 3 chunks, in chunk order, into out (and outq).  One warp task per
// (chunk, slab): chunk c owns the row that STARTS in c and crosses both of the next two chunk boundaries.
// init (optional, [rows, ldi]): a per-row term added first (the backward's gq_scale * dinfo * gq).
template <int VEC>
__global__ void __launch_bounds__(kMpWarps * 32)
    mp_fixup_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ chunk_row, int n_chunks, int chunk,
                    int64_t rows, int C, int nslab, const float* __restrict__ py, float* __restrict__ out, int64_t ldo,
                    const float* __restrict__ pq, float* __restrict__ outq, int64_t ldq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = static_cast<int64_t>(blockIdx.x) * kMpWarps + warp;
  if (task >= static_cast<int64_t>(n_chunks) * nslab) return;
  const int slab = static_cast<int>(task / n_chunks);
  const int c = static_cast<int>(task - static_cast<int64_t>(slab) * n_chunks);
  if (c + 2 >= n_chunks) return;
  const int r = __ldg(chunk_row + c + 1);   // the row holding entry (c+1)*chunk
  if (r >= rows) return;
  const int64_t rs = __ldg(rowptr + r), re = __ldg(rowptr + r + 1);
  if (rs < static_cast<int64_t>(c) * chunk || rs >= static_cast<int64_t>(c + 1) * chunk) return;  // starts elsewhere
  const int lc = static_cast<int>((re - 1) / chunk);
  if (lc < c + 2) return;
  const int c0 = (slab * 32 + lane) * VEC;
  if (c0 >= C) return;
  float acc[VEC], accq[VEC], t[VEC];
  ld_vec<VEC>(py + (static_cast<int64_t>(c) * 2 + 1) * C + c0, acc);
  if (pq) ld_vec<VEC>(pq + (static_cast<int64_t>(c) * 2 + 1) * C + c0, accq);
  for (int cc = c + 1; cc <= lc; ++cc) {
    ld_vec<VEC>(py + static_cast<int64_t>(cc) * 2 * C + c0, t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += t[i];
    if (pq) {
      ld_vec<VEC>(pq + static_cast<int64_t>(cc) * 2 * C + c0, t);
#pragma unroll
      for (int i = 0; i < VEC; ++i) accq[i] += t[i];
    }
  }
  st_vec<VEC>(out + static_cast<int64_t>(r) * ldo + c0, acc);
  if (pq) st_vec<VEC>(outq + static_cast<int64_t>(r) * ldq + c0, accq);
}

// tail_feat[t, :] / tail_grad[t, :] = feature / gradient codewords of tail entry t, all branches (D == 4, Wp == 8):
// one warp per tail entry, one 256-bit codeword load per lane (= branch)
__global__ void __launch_bounds__(256)
    tail_materialize_kernel(const int32_t* __restrict__ tail_node, int64_t T, const int16_t* __restrict__ codes,
                            const float* __restrict__ O, int nb, int M, float* __restrict__ xt, int64_t ldx,
                            float* __restrict__ gt, int64_t ldg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (t >= T) return;
  const int64_t node = tail_node ? __ldg(tail_node + t) : t;
  for (int k = lane; k < nb; k += 32) {
    const int code = __ldg(codes + node * nb + k);
    float f[4], g[4];
    ld_sector(O + (static_cast<int64_t>(k) * M + code) * 8, f, g);
    if (xt) st_vec<4>(xt + t * ldx + 4 * k, f);
    if (gt) st_vec<4>(gt + t * ldg + 4 * k, g);
  }
}

// dx <- gq_scale * dinfo * gq for rows WITHOUT transposed entries, 0 for the others (their CSR-independent term is
// added by the piece that holds the row start, so that every element sees a fixed order of additions)
template <int VEC>
__global__ void mp_bwd_init_kernel(int64_t B, int C, const int32_t* __restrict__ browptr,
                                   const float* __restrict__ gq, int64_t ldgq, float gq_scale,
                                   const float* __restrict__ dinfo, float* __restrict__ dx, int64_t lddx) {
  const int cv = C / VEC;
  const int64_t n = B * cv;
  const float s = gq ? gq_scale * (dinfo ? __ldg(dinfo) : 1.f) : 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cv;
    const int c = static_cast<int>(i - r * cv) * VEC;
    float q[VEC];
    if (gq && __ldg(browptr + r) == __ldg(browptr + r + 1)) {
      ld_vec<VEC>(gq + r * ldgq + c, q);
#pragma unroll
      for (int j = 0; j < VEC; ++j) q[j] *= s;
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) q[j] = 0.f;
    }
    st_vec<VEC>(dx + r * lddx + c, q);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kMpWarps * 32)
    mp_bwd_kernel(const int32_t* __restrict__ browptr, const int32_t* __restrict__ brow,
                  const float* __restrict__ bval, const int32_t* __restrict__ chunk_row, int n_chunks, int chunk,
                  int nnz, int B, const float* __restrict__ dy, int64_t lddy, Codebook cb, int C, int nslab,
                  float tail_scale, const float* __restrict__ dinfo, const float* __restrict__ gq, int64_t ldgq,
                  float gq_scale, float* __restrict__ dx, int64_t lddx, float* __restrict__ pdx) {
  const int lane = threadIdx.x & 31;
  const MpTask t = mp_task<VEC>(chunk_row, n_chunks, chunk, nnz, nslab, C, cb.D);
  if (!t.valid) return;
  const float ts = tail_scale * (dinfo ? __ldg(dinfo) : 1.f);
  const float gs = gq ? gq_scale * (dinfo ? __ldg(dinfo) : 1.f) : 0.f;
  const int64_t ch = t.eb / chunk;
  float acc[VEC], unused[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f, unused[i] = 0.f;
  auto body = [&](const EntryGroup& g) {
    if (t.active)
      gather_accumulate<VEC, false, false>(g, B, dy, lddy, cb, cb.D, ts, t.c0, t.k, t.off, acc, unused);
  };
  auto body_dense = [&](const EntryGroup& g) {
    if (t.active)
      gather_accumulate<VEC, false, false, true>(g, B, dy, lddy, cb, cb.D, ts, t.c0, t.k, t.off, acc, unused);
  };
  auto flush = [&](int j, bool whole, int rs, int re) {
    if (t.active) {
      const int kind = piece_kind(whole, rs, re, t.eb, chunk);
      if (gq && rs >= t.eb) {  // the piece holding the row start carries the CSR-independent term
        float q[VEC];
        ld_vec<VEC>(gq + static_cast<int64_t>(j) * ldgq + t.c0, q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(gs, q[i], acc[i]);
      }
      float* dp = dx + static_cast<int64_t>(j) * lddx + t.c0;
      if (kind == kPieceWhole) st_vec<VEC>(dp, acc);
      else if (kind == kPieceRed) red_vec<VEC>(dp, acc);      // two REDs onto zero: order-independent
      else st_vec<VEC>(pdx + (ch * 2 + (kind == kPieceHubStart ? 1 : 0)) * C + t.c0, acc);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  };
  PlainWeights pol;
  if (cb.tail_grad != nullptr)
    walk_rows<false>(t.eb, t.ee, t.row0, B, browptr, brow, bval, nullptr, B, cb.tail_node, lane, pol, body_dense, flush);
  else
    walk_rows<false>(t.eb, t.ee, t.row0, B, browptr, brow, bval, nullptr, B, cb.tail_node, lane, pol, body, flush);
}

}  // namespace vqgnn

using namespace vqgnn;

static inline size_t mp_al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// workspace layout: [count (256 B)] [per-block info partials] [y pieces] [second piece buffer (gq)]
struct MpWs {
  unsigned int* count;
  double* part;
  float* p0;
  float* p1;
};
static size_t mp_ws_layout(void* ws, int64_t nnz, int chunk, int C, MpWs* out) {
  const int64_t n_chunks = chunk > 0 ? (nnz + chunk - 1) / chunk : 0;
  const int vec = 4;   // upper bound of the grid: the VEC = 1 variant has more slabs, sized for it below
  (void)vec;
  const int64_t nslab1 = (C + 31) / 32;
  const size_t grid_max = static_cast<size_t>(n_chunks * nslab1) + 1;   // >= the grid of every forward variant
  const size_t pieces = mp_al256(static_cast<size_t>(n_chunks) * 2 * C * sizeof(float));
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  if (out) {
    out->count = reinterpret_cast<unsigned int*>(p);
    out->part = reinterpret_cast<double*>(p + 256);
    out->p0 = reinterpret_cast<float*>(p + 256 + mp_al256(grid_max * 8));
    out->p1 = reinterpret_cast<float*>(p + 256 + mp_al256(grid_max * 8) + pieces);
  }
  return 512 + mp_al256(grid_max * 8) + 2 * pieces;
}

extern "C" size_t vqgnn_mp_workspace_bytes(int64_t nnz, int chunk, int C) {
  return mp_ws_layout(nullptr, nnz, chunk, C, nullptr);
}

extern "C" int64_t vqgnn_mp_num_chunks(int64_t nnz, int chunk) {
  return chunk > 0 ? (nnz + chunk - 1) / chunk : -1;
}

extern "C" int vqgnn_mp_chunk_rows(const int32_t* rowptr, int64_t R, int64_t nnz, int chunk, int32_t* chunk_row,
                                   void* stream) {
  VQ_CHECK_ARG(rowptr && R > 0 && nnz >= 0 && chunk > 0 && chunk % 32 == 0, "mp_chunk_rows: bad arguments");
  VQ_CHECK_ARG(nnz < (1ll << 31), "mp_chunk_rows: nnz must fit int32");
  const int n_chunks = static_cast<int>(vqgnn_mp_num_chunks(nnz, chunk));
  if (n_chunks == 0) return VQGNN_OK;
  VQ_CHECK_ARG(chunk_row, "mp_chunk_rows: null output");
  mp_chunk_rows_kernel<<<ceil_div(n_chunks, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rowptr, R, nnz, chunk, n_chunks, chunk_row);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_tail_materialize(const int32_t* tail_node, int64_t T, const int16_t* codes, const float* O,
                                      int nb, int M, int D, int Wp, float* tail_feat, float* tail_grad,
                                      int64_t ld_tail, void* stream) {
  VQ_CHECK_ARG(codes && O && T >= 0 && nb > 0 && (tail_feat || tail_grad), "tail_materialize: bad arguments");
  VQ_CHECK_ARG(D == 4 && Wp == 8 && aligned32(O) && ld_tail % 4 == 0 && ld_tail >= 4 * nb &&
                   (!tail_feat || aligned16(tail_feat)) && (!tail_grad || aligned16(tail_grad)),
               "tail_materialize: needs D == 4, Wp == 8 and 16 B aligned outputs");
  if (T == 0) return VQGNN_OK;
  tail_materialize_kernel<<<ceil_div(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      tail_node, T, codes, O, nb, M, tail_feat, ld_tail, tail_grad, ld_tail);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_mp_fwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* rval,
                            const int32_t* chunk_row, int chunk, int64_t nnz, int64_t R, int64_t B,
                            const float* x, int64_t ldx, const int32_t* tail_node, const int16_t* codes,
                            const float* O, int nb, int M, int D, int Wp, const float* tail_feat,
                            int64_t ld_tail, int tail_slab, float feat_scale, float info_scale,
                            float* y, int64_t ldy, float* gq, int64_t ldgq, float* info, void* ws,
                            void* stream) {
  VQ_CHECK_ARG(rowptr && x && codes && O && y && (nnz == 0 || (col && val)), "mp_fwd: null argument");
  VQ_CHECK_ARG(R >= B && B > 0 && nb > 0 && D > 0 && Wp >= 2 * D, "mp_fwd: bad sizes");
  VQ_CHECK_ARG(ws || nnz == 0, "mp_fwd: needs a workspace of vqgnn_mp_workspace_bytes(nnz, chunk, nb*D) bytes");
  VQ_CHECK_ARG(B < (1ll << 31) && R < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "mp_fwd: sizes must fit int32");
  VQ_CHECK_ARG(chunk > 0 && chunk % 32 == 0 && (nnz == 0 || chunk_row), "mp_fwd: needs chunk_row (vqgnn_mp_chunk_rows)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = nb * D;
  Codebook cb{tail_node, codes, O, nb, M, D, Wp};
  if (tail_feat) {
    VQ_CHECK_ARG(!rval && (tail_slab > 0 || ld_tail % 4 == 0) && tail_slab % 4 == 0 && aligned16(tail_feat),
                 "mp_fwd: dense tail rows need rval == NULL and 16 B alignment");
    cb.tail_feat = tail_feat, cb.ld_tail = ld_tail, cb.tail_slab = tail_slab;
  }
  MpWs w{nullptr, nullptr, nullptr, nullptr};
  if (ws) mp_ws_layout(ws, nnz, chunk, C, &w);
  if (info && ws) VQ_CUDA(cudaMemsetAsync(w.count, 0, 16, s));
  // rows cut by a chunk boundary accumulate with REDs, empty rows are never visited: start from zero
  if (int rc = zero_rows(y, B, C, ldy, s)) return rc;
  if (gq && rval)
    if (int rc = zero_rows(gq, B, C, ldgq, s)) return rc;
  const int n_chunks = static_cast<int>(vqgnn_mp_num_chunks(nnz, chunk));
  if (n_chunks == 0) {
    if (info) VQ_CUDA(cudaMemsetAsync(info, 0, sizeof(float), s));
    return VQGNN_OK;
  }
  const bool vec4 = (D == 4) && (Wp % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(x) &&
                    aligned16(y) && aligned16(O) && (!gq || (ldgq % 4 == 0 && aligned16(gq)));
  const bool wide = vec4 && rval && Wp == 8 && aligned32(O);
  const int vec = vec4 ? 4 : 1;
  const int nslab = ceil_div(C, 32 * vec);
  const int64_t tasks = static_cast<int64_t>(n_chunks) * nslab;
  const int grid = ceil_div(tasks, kMpWarps);
  static const bool use_async = []() {
    const char* e = getenv("VQGNN_MPFWD_ASYNC");
    return !(e && e[0] == '0');
  }();
  if (use_async && vec4 && tail_feat && tail_slab == 0 && !rval && feat_scale == 1.0f && C >= 64) {
    // large v2 batch graphs with materialised tail rows: deep cp.async gather pipeline
    static const int depth = []() {
      const char* e = getenv("VQGNN_ASYNC_DEPTH");
      const int d = e ? atoi(e) : kAsyncDepthDefault;
      return (d == 12 || d == 16) ? d : kAsyncDepthDefault;
    }();
    const size_t smem = static_cast<size_t>(kMpWarps) * depth * 32 * 16;
#define VQ_ASYNC(DD)                                                                                              \
  do {                                                                                                            \
    VQ_CUDA(cudaFuncSetAttribute(mp_fwd_async_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mp_fwd_async_kernel<DD><<<grid, kMpWarps * 32, smem, s>>>(rowptr, col, val, chunk_row, n_chunks, chunk, (int)nnz, \
                                                              R, (int)B, x, ldx, cb, C, nslab, info_scale, y, ldy, \
                                                              info, w.part, w.count, w.p0);                       \
  } while (0)
    if (depth == 12) VQ_ASYNC(12);
    else if (depth == 16) VQ_ASYNC(16);
    else VQ_ASYNC(8);
#undef VQ_ASYNC
    VQ_LAUNCH_CHECK();
    if (n_chunks > 2) {
      mp_fixup_kernel<4><<<grid, kMpWarps * 32, 0, s>>>(rowptr, chunk_row, n_chunks, chunk, B, C, nslab, w.p0, y, ldy,
                                                        nullptr, gq, ldgq);
      VQ_LAUNCH_CHECK();
    }
    return VQGNN_OK;
  }
#define VQ_MP_FWD(VEC, GQ, WIDE)                                                                              \
  mp_fwd_kernel<VEC, GQ, WIDE><<<grid, kMpWarps * 32, 0, s>>>(rowptr, col, val, rval, chunk_row, n_chunks,    \
                                                              chunk, (int)nnz, R, (int)B, x, ldx, cb, C, nslab, \
                                                              feat_scale, info_scale, y, ldy, gq, ldgq, info, \
                                                              w.part, w.count, w.p0, w.p1)
  if (vec4) {
    if (wide) VQ_MP_FWD(4, true, true);
    else if (rval) VQ_MP_FWD(4, true, false);
    else VQ_MP_FWD(4, false, false);
  } else {
    if (rval) VQ_MP_FWD(1, true, false);
    else VQ_MP_FWD(1, false, false);
  }
#undef VQ_MP_FWD
  VQ_LAUNCH_CHECK();
  if (n_chunks > 2) {   // rows spanning >= 3 chunks: ordered sum of their pieces
    float* pq = (gq && rval) ? w.p1 : nullptr;
    if (vec4)
      mp_fixup_kernel<4><<<grid, kMpWarps * 32, 0, s>>>(rowptr, chunk_row, n_chunks, chunk, B, C, nslab, w.p0, y, ldy,
                                                        pq, gq, ldgq);
    else
      mp_fixup_kernel<1><<<grid, kMpWarps * 32, 0, s>>>(rowptr, chunk_row, n_chunks, chunk, B, C, nslab, w.p0, y, ldy,
                                                        pq, gq, ldgq);
    VQ_LAUNCH_CHECK();
  }
  return VQGNN_OK;
}

extern "C" int vqgnn_mp_fwd_rows(const int32_t* rowptr, const int32_t* col, const float* val,
                                 const int32_t* chunk_row, int chunk, int64_t nnz, int64_t R, int64_t B,
                                 const float* x, int64_t ldx, const float* tail_feat, int64_t T, float tail_scale,
                                 const float* tail_scale_dev, const float* tail_grad, int64_t ld_tail, int C,
                                 float info_scale, float* y, int64_t ldy, float* info, void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && x && y && (nnz == 0 || (col && val)), "mp_fwd_rows: null argument");
  VQ_CHECK_ARG(R >= B && B > 0 && B < (1ll << 31) && R < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31),
               "mp_fwd_rows: bad sizes");
  VQ_CHECK_ARG(R == B || tail_feat, "mp_fwd_rows: out-of-batch columns need the materialised feature rows (tail_feat)");
  VQ_CHECK_ARG(T >= 0 && T >= R - B && T < (1ll << 31) && (T == 0 || tail_feat), "mp_fwd_rows: bad tail row count");
  VQ_CHECK_ARG(!info || R == B || tail_grad, "mp_fwd_rows: info needs the materialised gradient rows (tail_grad)");
  VQ_CHECK_ARG(C >= 16 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ld_tail % 4 == 0 && aligned16(x) &&
                   aligned16(y) && (!tail_feat || aligned16(tail_feat)) && (!tail_grad || aligned16(tail_grad)),
               "mp_fwd_rows: needs C >= 16, C % 4 == 0 and 16 B aligned rows");
  VQ_CHECK_ARG(chunk > 0 && chunk % 32 == 0 && chunk <= kRowsChunkMax && (nnz == 0 || chunk_row),
               "mp_fwd_rows: needs chunk_row (vqgnn_mp_chunk_rows) with chunk <= 256");
  VQ_CHECK_ARG(ws || nnz == 0, "mp_fwd_rows: needs a workspace of vqgnn_mp_workspace_bytes(nnz, chunk, C) bytes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MpWs w{nullptr, nullptr, nullptr, nullptr};
  if (ws) mp_ws_layout(ws, nnz, chunk, C, &w);
  if (info && ws) VQ_CUDA(cudaMemsetAsync(w.count, 0, 16, s));
  if (int rc = zero_rows(y, B, C, ldy, s)) return rc;
  const int n_chunks = static_cast<int>(vqgnn_mp_num_chunks(nnz, chunk));
  if (n_chunks == 0) {
    if (info) VQ_CUDA(cudaMemsetAsync(info, 0, sizeof(float), s));
    return VQGNN_OK;
  }
  const int nslab = ceil_div(C, 128);
  const int64_t tasks = static_cast<int64_t>(n_chunks) * nslab;
  // 32-bit row offsets (16 B units) from a common base below both tables
  const uintptr_t xa = reinterpret_cast<uintptr_t>(x), ta = tail_feat ? reinterpret_cast<uintptr_t>(tail_feat) : xa;
  const uintptr_t base = std::min(xa, ta);
  const uint64_t x_end4 = (xa - base) / 16 + static_cast<uint64_t>(B) * (ldx / 4) + 32;
  const uint64_t t_end4 = (ta - base) / 16 + static_cast<uint64_t>(T) * (ld_tail / 4) + 32;
  VQ_CHECK_ARG(x_end4 < (1ull << 32) && t_end4 < (1ull << 32),
               "mp_fwd_rows: x and tail_feat must lie within 64 GB of each other (32-bit row offsets)");
  // warps per CTA x ring slots, measured at the products shape (16 M entries, C = 128): 4 x 8 -> 0.90 ms per launch,
  // 8 x 8 0.98, 2 x 8 0.97, 1 x 8 1.05, 4 x 16 0.97, 8 x 4 1.11, 4 x 32 1.54 (mp_fwd_async_kernel: 1.36)
  constexpr int NW = kRowsWarps, SLOTS = kRowsSlots;
  const size_t smem = sizeof(RowsWarpSmem<SLOTS>) * NW;
  mp_fwd_rows_kernel<NW, SLOTS><<<ceil_div(tasks, NW), NW * 32, smem, s>>>(
      rowptr, col, val, chunk_row, n_chunks, chunk, (int)nnz, (int)R, (int)B, reinterpret_cast<const float4*>(base),
      static_cast<uint32_t>((xa - base) / 16), static_cast<uint32_t>(ldx / 4), static_cast<uint32_t>((ta - base) / 16),
      static_cast<uint32_t>(ld_tail / 4), tail_scale, tail_scale_dev, tail_grad, ld_tail, C, nslab, info_scale, y, ldy,
      info, w.part, w.count, w.p0);
  VQ_LAUNCH_CHECK();
  if (n_chunks > 2) {   // rows spanning >= 3 chunks: ordered sum of their pieces
    mp_fixup_kernel<4><<<ceil_div(tasks, kMpWarps), kMpWarps * 32, 0, s>>>(rowptr, chunk_row, n_chunks, chunk, B, C,
                                                                         nslab, w.p0, y, ldy, nullptr, nullptr, 0);
    VQ_LAUNCH_CHECK();
  }
  return VQGNN_OK;
}

extern "C" int vqgnn_mp_bwd(const int32_t* browptr, const int32_t* brow, const float* bval,
                            const int32_t* chunk_row, int chunk, int64_t nnz, int64_t B, const float* dy,
                            int64_t lddy, const int32_t* tail_node, const int16_t* codes, const float* O, int nb,
                            int M, int D, int Wp, const float* tail_grad, int64_t ld_tail, int tail_slab,
                            float tail_scale,
                            const float* gq, int64_t ldgq, float gq_scale, const float* dinfo, float* dx,
                            int64_t lddx, void* ws, void* stream) {
  VQ_CHECK_ARG(browptr && dy && codes && O && dx && (nnz == 0 || (brow && bval)), "mp_bwd: null argument");
  VQ_CHECK_ARG(B > 0 && B < (1ll << 31) && nb > 0 && D > 0 && Wp >= 2 * D, "mp_bwd: bad sizes");
  VQ_CHECK_ARG(nnz >= 0 && nnz < (1ll << 31), "mp_bwd: nnz must fit int32");
  VQ_CHECK_ARG(chunk > 0 && chunk % 32 == 0 && (nnz == 0 || chunk_row), "mp_bwd: needs chunk_row (vqgnn_mp_chunk_rows)");
  VQ_CHECK_ARG(ws || nnz == 0, "mp_bwd: needs a workspace of vqgnn_mp_workspace_bytes(nnz, chunk, nb*D) bytes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = nb * D;
  MpWs w{nullptr, nullptr, nullptr, nullptr};
  if (ws) mp_ws_layout(ws, nnz, chunk, C, &w);
  Codebook cb{tail_node, codes, O, nb, M, D, Wp};
  if (tail_grad) {
    VQ_CHECK_ARG((tail_slab > 0 || ld_tail % 4 == 0) && tail_slab % 4 == 0 && aligned16(tail_grad),
                 "mp_bwd: dense tail rows must be 16 B aligned");
    cb.tail_grad = tail_grad, cb.ld_tail = ld_tail, cb.tail_slab = tail_slab;
  }
  const bool vec4 = (D == 4) && (Wp % 4 == 0) && (lddy % 4 == 0) && (lddx % 4 == 0) && aligned16(dy) &&
                    aligned16(dx) && aligned16(O) && (!gq || (ldgq % 4 == 0 && aligned16(gq)));
  const int vec = vec4 ? 4 : 1;
  const int nslab = ceil_div(C, 32 * vec);
  const int n_chunks = static_cast<int>(vqgnn_mp_num_chunks(nnz, chunk));
  const int init_grid = static_cast<int>(std::min<int64_t>((B * (C / vec) + 255) / 256, 8 * kNumSMs));
  if (vec4) mp_bwd_init_kernel<4><<<init_grid, 256, 0, s>>>(B, C, browptr, gq, ldgq, gq_scale, dinfo, dx, lddx);
  else mp_bwd_init_kernel<1><<<init_grid, 256, 0, s>>>(B, C, browptr, gq, ldgq, gq_scale, dinfo, dx, lddx);
  VQ_LAUNCH_CHECK();
  if (n_chunks == 0) return VQGNN_OK;
  const int grid = ceil_div(static_cast<int64_t>(n_chunks) * nslab, kMpWarps);
  if (vec4)
    mp_bwd_kernel<4><<<grid, kMpWarps * 32, 0, s>>>(browptr, brow, bval, chunk_row, n_chunks, chunk, (int)nnz,
                                                    (int)B, dy, lddy, cb, C, nslab, tail_scale, dinfo, gq, ldgq,
                                                    gq_scale, dx, lddx, w.p0);
  else
    mp_bwd_kernel<1><<<grid, kMpWarps * 32, 0, s>>>(browptr, brow, bval, chunk_row, n_chunks, chunk, (int)nnz,
                                                    (int)B, dy, lddy, cb, C, nslab, tail_scale, dinfo, gq, ldgq,
                                                    gq_scale, dx, lddx, w.p0);
  VQ_LAUNCH_CHECK();
  if (n_chunks > 2) {
    if (vec4)
      mp_fixup_kernel<4><<<grid, kMpWarps * 32, 0, s>>>(browptr, chunk_row, n_chunks, chunk, B, C, nslab, w.p0, dx,
                                                        lddx, nullptr, nullptr, 0);
    else
      mp_fixup_kernel<1><<<grid, kMpWarps * 32, 0, s>>>(browptr, chunk_row, n_chunks, chunk, B, C, nslab, w.p0, dx,
                                                        lddx, nullptr, nullptr, 0);
    VQ_LAUNCH_CHECK();
  }
  return VQGNN_OK;
}
