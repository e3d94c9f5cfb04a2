// VQ-approximated GAT message passing, v2 ("B+B'") formulation: vq_gnn_v2/models.py:144-231 with
// OurGATConv (vq_gnn_v2/convs.py:165-266) and vq_softmax == un-normalised exp (utils/vq_softmax.py:41-57).
//
//   Xin = [x ; codeword features of the B' out-of-batch nodes | 1]            (C+1 columns, never materialised)
//   a_l[n] = <Xin[n], att_l>, a_r[n] = <Xin[n], att_r>
//   sigma  = sqrt(max(a_l)^2 + 1) * sqrt(max(a_r)^2 + 1)                       ("Trick 1", convs.py:209-211)
//   w_ij   = adj[i,j] * exp(leaky_relu((a_l[j] + a_r[i]) / sigma))             ("Trick 2", convs.py:264)
//   Y[i]   = sum_j w_ij Xin[j]   ;   den[i] = sum_j w_ij  (the ones column)
//   out[i] = Y[i,:C] / (den[i] + 1e-16)  for i < B ; info = wu * sum_{r>=B} <Y[r,:C], Gq[r]>   (models.py:187-198)
//
// Same nnz-balanced warp tasks as mp.cu (mp_common.cuh); the edge weight is recomputed per entry from the two
// score vectors (8 B per entry instead of materialising nnz x (C+1) messages as PyG does).  The backward is
// three passes: an SDDMM-shaped pass over the forward CSR for the score gradients, the transposed weighted
// SpMM for d x, and the attention-vector gradients.
#include <math.h>

#include <algorithm>

#include "mp_common.cuh"
#include "mp_rows.cuh"

namespace vqgnn {

__device__ __forceinline__ float gat_inv_sigma(const float* __restrict__ stat) {
  const float ml = __ldg(stat), mr = __ldg(stat + 1);
  return 1.f / (sqrtf(ml * ml + 1.f) * sqrtf(mr * mr + 1.f));
}

// w = v * exp(leaky_relu((a_col + a_row) * inv_sigma)); also tracks the row's denominator
struct GatWeights {
  const float* a_col;  // score indexed by the entry's column id
  const float* a_row;  // score indexed by the row id
  float inv_sigma, slope;
  float a_cur = 0.f, den = 0.f;
  __device__ __forceinline__ float load_extra(int c) const { return __ldg(a_col + c); }
  __device__ __forceinline__ void row_begin(int r) { a_cur = __ldg(a_row + r); }
  __device__ __forceinline__ float weight(float v, float a, bool valid) {
    float e = (a + a_cur) * inv_sigma;
    e = e > 0.f ? e : slope * e;
    const float w = v * expf(e);
    if (valid) den += w;
    return w;
  }
};

__global__ void gat_stat_init_kernel(float* stat) {
  if (threadIdx.x < 2) stat[threadIdx.x] = __int_as_float(0xff800000);  // -inf
}

// Xin[n, c0..c0+VEC) for a node of the batch graph (n < B: dense row; else the node's codeword feature)
template <int VEC>
__device__ __forceinline__ void load_xin(int n, int B, const float* __restrict__ x, int64_t ldx,
                                         const Codebook& cb, int c0, int k, int off, float (&v)[VEC]) {
  if (n < B) {
    ld_vec<VEC>(x + static_cast<int64_t>(n) * ldx + c0, v);
  } else if (cb.tail_feat) {   // dense rows of gathered codewords (vqgnn_tail_materialize)
    ld_vec<VEC>(cb.tail_feat + static_cast<int64_t>(n - B) * cb.ld_tail + c0, v);
  } else {
    const int node = cb.tail_node ? __ldg(cb.tail_node + (n - B)) : (n - B);
    const int code = __ldg(cb.codes + static_cast<int64_t>(node) * cb.nb + k);
    ld_vec<VEC>(cb.O + (static_cast<int64_t>(k) * cb.M + code) * cb.Wp + off, v);
  }
}

// (F1) scores: one warp per node, lanes over columns
template <int VEC>
__global__ void __launch_bounds__(256)
    gat_scores_kernel(int R, int B, const float* __restrict__ x, int64_t ldx, Codebook cb, int C,
                      const float* __restrict__ att_l, const float* __restrict__ att_r, float* __restrict__ a_l,
                      float* __restrict__ a_r, float* __restrict__ stat) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  float sl = 0.f, sr = 0.f;
  if (n < R) {
    for (int c0 = lane * VEC; c0 < C; c0 += 32 * VEC) {
      const int k = c0 / cb.D, off = c0 - k * cb.D;
      float v[VEC], wl[VEC], wr[VEC];
      load_xin<VEC>(n, B, x, ldx, cb, c0, k, off, v);
      ld_vec<VEC>(att_l + c0, wl);
      ld_vec<VEC>(att_r + c0, wr);
#pragma unroll
      for (int i = 0; i < VEC; ++i) sl = fmaf(v[i], wl[i], sl), sr = fmaf(v[i], wr[i], sr);
    }
    sl = warp_sum(sl) + __ldg(att_l + C);  // the ones column (models.py:176-177)
    sr = warp_sum(sr) + __ldg(att_r + C);
    if (lane == 0) a_l[n] = sl, a_r[n] = sr;
  } else {
    sl = sr = __int_as_float(0xff800000);
  }
  __shared__ float shl[8], shr[8];
  if (lane == 0) shl[warp] = sl, shr[warp] = sr;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ml = shl[0], mr = shr[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) ml = fmaxf(ml, shl[i]), mr = fmaxf(mr, shr[i]);
    atomic_max_float(stat, ml);
    atomic_max_float(stat + 1, mr);
  }
}

// (F2) un-normalised aggregation + denominators + info_backward
template <int VEC>
__global__ void __launch_bounds__(kMpWarps * 32)
    gat_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                   const float* __restrict__ val, const int32_t* __restrict__ chunk_row, int n_chunks, int chunk,
                   int nnz, int64_t R, int B, const float* __restrict__ x, int64_t ldx, Codebook cb, int C,
                   int nslab, const float* __restrict__ a_l, const float* __restrict__ a_r,
                   const float* __restrict__ stat, float slope, float info_scale, float* __restrict__ y,
                   int64_t ldy, float* __restrict__ den, float* __restrict__ info, double* ws_sum,
                   unsigned int* ws_count) {
  const int lane = threadIdx.x & 31;
  const MpTask t = mp_task<VEC>(chunk_row, n_chunks, chunk, nnz, nslab, C, cb.D);
  float fpart = 0.f;
  if (t.valid) {
    float acc[VEC], unused[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f, unused[i] = 0.f;
    GatWeights pol{a_l, a_r, gat_inv_sigma(stat), slope};
    auto body = [&](const EntryGroup& g) {
      if (t.active) gather_accumulate<VEC, false, false>(g, B, x, ldx, cb, 0, 1.f, t.c0, t.k, t.off, acc, unused);
    };
    auto body_dense = [&](const EntryGroup& g) {
      if (t.active) gather_accumulate<VEC, false, false, true>(g, B, x, ldx, cb, 0, 1.f, t.c0, t.k, t.off, acc, unused);
    };
    auto flush = [&](int r, bool whole) {
      if (t.active) {
        if (r < B) {
          float* yp = y + static_cast<int64_t>(r) * ldy + t.c0;
          if (whole) st_vec<VEC>(yp, acc);
          else red_vec<VEC>(yp, acc);
          if (t.slab == 0 && lane == 0) {
            if (whole) den[r] = pol.den;
            else atomicAdd(den + r, pol.den);
          }
        } else if (info) {  // rows >= B stay un-normalised (models.py:187-198)
          float gv[VEC], dummy[VEC];
          (void)dummy;
          const int node = cb.tail_node ? __ldg(cb.tail_node + (r - B)) : (r - B);
          const int code = __ldg(cb.codes + static_cast<int64_t>(node) * cb.nb + t.k);
          ld_vec<VEC>(cb.O + (static_cast<int64_t>(t.k) * cb.M + code) * cb.Wp + cb.D + t.off, gv);
#pragma unroll
          for (int i = 0; i < VEC; ++i) fpart = fmaf(acc[i], gv[i], fpart);
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
      pol.den = 0.f;
    };
    if (cb.tail_feat != nullptr)
      walk_rows<false>(t.eb, t.ee, t.row0, R, rowptr, col, val, nullptr, B, cb.tail_node, lane, pol, body_dense, flush);
    else
      walk_rows<false>(t.eb, t.ee, t.row0, R, rowptr, col, val, nullptr, B, cb.tail_node, lane, pol, body, flush);
  }
  if (info) info_reduce(static_cast<double>(fpart), ws_sum, ws_count, info_scale, info);
}

// (F3) out = Y / (den + 1e-16) in place
__global__ void gat_normalize_kernel(int64_t B, int C, float* __restrict__ y, int64_t ldy,
                                     const float* __restrict__ den) {
  const int64_t n = B * C;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i - r * C);
    y[r * ldy + c] = y[r * ldy + c] / (__ldg(den + r) + 1e-16f);
  }
}

// (B0) dYn = dOut / (den + eps) ; dden = -<dOut, out> / (den + eps).  One warp per batch row.
__global__ void __launch_bounds__(256)
    gat_bwd_prep_kernel(int B, int C, const float* __restrict__ dout, int64_t lddo, const float* __restrict__ out,
                        int64_t ldo, const float* __restrict__ den, float* __restrict__ dyn, int64_t lddyn,
                        float* __restrict__ dden) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= B) return;
  const float inv = 1.f / (__ldg(den + r) + 1e-16f);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = __ldg(dout + static_cast<int64_t>(r) * lddo + c);
    s = fmaf(d, __ldg(out + static_cast<int64_t>(r) * ldo + c), s);
    dyn[static_cast<int64_t>(r) * lddyn + c] = d * inv;
  }
  s = warp_sum(s);
  if (lane == 0) dden[r] = -s * inv;
}

// (B1) score gradients (SDDMM over the forward CSR):
//   de_ij = <dY'[i], Xin[j]> * w_ij * lrelu'(e_ij);  ds_r[i] += de_ij;  ds_l[j] += de_ij
// with dY'[i] = [dYn[i] | dden[i]] for i < B and [tail_scale * dinfo * Gq[i] | 0] for i >= B.
template <int VEC>
__global__ void __launch_bounds__(kMpWarps * 32)
    gat_bwd_edge_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                        const float* __restrict__ val, const int32_t* __restrict__ chunk_row, int n_chunks,
                        int chunk, int nnz, int64_t R, int B, const float* __restrict__ x, int64_t ldx,
                        Codebook cb, int C, int nslab, const float* __restrict__ a_l,
                        const float* __restrict__ a_r, const float* __restrict__ stat, float slope,
                        const float* __restrict__ dyn, int64_t lddyn, const float* __restrict__ dden,
                        float tail_scale, const float* __restrict__ dinfo, float* __restrict__ ds_l,
                        float* __restrict__ ds_r) {
  constexpr int U = kMpUnroll;
  const int lane = threadIdx.x & 31;
  const MpTask t = mp_task<VEC>(chunk_row, n_chunks, chunk, nnz, nslab, C, cb.D);
  if (!t.valid) return;
  const float inv_sigma = gat_inv_sigma(stat);
  const float ts = tail_scale * (dinfo ? __ldg(dinfo) : 1.f);

  struct RowState : GatWeights {
    // extends the weight policy with the row's dY' slice
    const float* dyn;
    const float* dden;
    int64_t lddyn;
    const Codebook* cb;
    int B, c0, k, off;
    float ts;
    bool active, ones;
    float dy[VEC];
    float dd = 0.f, dsr = 0.f;
    __device__ __forceinline__ void row_begin(int r) {
      a_cur = __ldg(a_row + r);
      dd = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) dy[i] = 0.f;
      if (!active) return;
      if (r < B) {
        ld_vec<VEC>(dyn + static_cast<int64_t>(r) * lddyn + c0, dy);
        if (ones) dd = __ldg(dden + r);
      } else {
        if (cb->tail_grad) {
          ld_vec<VEC>(cb->tail_grad + static_cast<int64_t>(r - B) * cb->ld_tail + c0, dy);
        } else {
          const int node = cb->tail_node ? __ldg(cb->tail_node + (r - B)) : (r - B);
          const int code = __ldg(cb->codes + static_cast<int64_t>(node) * cb->nb + k);
          ld_vec<VEC>(cb->O + (static_cast<int64_t>(k) * cb->M + code) * cb->Wp + cb->D + off, dy);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) dy[i] *= ts;
      }
    }
  };
  RowState pol;
  pol.a_col = a_l, pol.a_row = a_r, pol.inv_sigma = inv_sigma, pol.slope = slope;
  pol.dyn = dyn, pol.dden = dden, pol.lddyn = lddyn, pol.cb = &cb, pol.B = B;
  pol.c0 = t.c0, pol.k = t.k, pol.off = t.off, pol.ts = ts;
  pol.active = t.active, pol.ones = t.slab == 0 && lane == 0;

  auto body = [&](const EntryGroup& g) {
    float part[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      part[u] = 0.f;
      if (t.active && g.c[u] >= 0) {
        float xv[VEC];
        load_xin<VEC>(g.c[u], B, x, ldx, cb, t.c0, t.k, t.off, xv);
#pragma unroll
        for (int i = 0; i < VEC; ++i) part[u] = fmaf(pol.dy[i], xv[i], part[u]);
        part[u] += pol.dd;  // ones column (only lane 0 of slab 0 holds a non-zero dd)
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float dot = warp_sum(part[u]);
      if (g.c[u] >= 0) {
        const float e = (g.xtra[u] + pol.a_cur) * inv_sigma;
        const float de = dot * g.v[u] * (e > 0.f ? 1.f : slope);  // g.v = val * exp(lrelu(e))
        pol.dsr += de;
        if (lane == 0) atomicAdd(ds_l + g.c[u], de);
      }
    }
  };
  auto flush = [&](int r, bool) {
    if (lane == 0) atomicAdd(ds_r + r, pol.dsr);
    pol.dsr = 0.f;
    pol.den = 0.f;
  };
  walk_rows<false>(t.eb, t.ee, t.row0, R, rowptr, col, val, nullptr, B, cb.tail_node, lane, pol, body, flush);
}

// (B1, lean) the same score gradients in the style of mp_rows.cuh, for materialised codeword rows: Xin[j] rows arrive
// through the warp's cp.async ring; per entry the lane's four columns give a partial dot product that is parked in a
// 32 x 33 shared-memory tile, and once per 32-entry batch lane L sums column L (conflict-free) -- two instructions per
// entry instead of a five-step shuffle reduction -- and finishes ITS entry: de = s * w * lrelu', one float atomic onto
// ds_l[col] and one onto ds_r[row].  dY'[i] (the row operand) is prefetched one row ahead.
constexpr int kEdgeWarps = 4;
struct __align__(128) GatEdgeSmem {
  RowsWarpSmem<kRowsSlots> r;
  float red[32 * 33];
};

__global__ void __launch_bounds__(kEdgeWarps * 32)
    gat_bwd_edge_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                             const float* __restrict__ val, const int32_t* __restrict__ chunk_row, int n_chunks,
                             int chunk, int nnz, int R, int B, const float4* __restrict__ base, uint32_t xoff4,
                             uint32_t ldx4, uint32_t toff4, uint32_t ldt4, const float* __restrict__ tail_grad,
                             int64_t ld_tail, int C, int nslab, const float* __restrict__ a_l,
                             const float* __restrict__ a_r, const float* __restrict__ stat, float slope,
                             const float* __restrict__ dyn, int64_t lddyn, const float* __restrict__ dden,
                             float tail_scale, const float* __restrict__ dinfo, float* __restrict__ ds_l,
                             float* __restrict__ ds_r) {
  constexpr int SLOTS = kRowsSlots, G = SLOTS / 4;
  extern __shared__ __align__(128) unsigned char edge_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  GatEdgeSmem& S = reinterpret_cast<GatEdgeSmem*>(edge_smem)[warp];
  const int64_t task = static_cast<int64_t>(blockIdx.x) * kEdgeWarps + warp;
  if (task >= static_cast<int64_t>(n_chunks) * nslab) return;
  const int slab = static_cast<int>(task / n_chunks);
  const int ch = static_cast<int>(task - static_cast<int64_t>(slab) * n_chunks);
  const int c0 = slab * 128 + lane * 4;
  const bool active = c0 < C;
  const int eb = ch * chunk, ee = min(eb + chunk, nnz);
  const int row0 = __ldg(chunk_row + ch);
  const int rowL = ch + 1 < n_chunks ? __ldg(chunk_row + ch + 1) : R - 1;
  const uint32_t slot_s = rows_smem_u32(S.r.ring) + lane * 16;
  const float ml = __ldg(stat), mr = __ldg(stat + 1);
  const float inv_sigma = 1.f / (sqrtf(ml * ml + 1.f) * sqrtf(mr * mr + 1.f));
  const float ts = tail_scale * (dinfo ? __ldg(dinfo) : 1.f);

  if (lane < kRowsChunkMax / 32) S.r.endmask[lane] = 0u;
  __syncwarp();
  for (int r = row0 + lane; r <= rowL; r += 32) {
    const int rs = __ldg(rowptr + r), re = __ldg(rowptr + r + 1);
    if (re > rs && re > eb && re <= ee) {
      const int p = re - 1 - eb;
      atomicOr(&S.r.endmask[p >> 5], 1u << (p & 31));
      S.r.row_of[p] = r;
    }
  }
  __syncwarp();

  // per-lane entry state of two batches: ring offset, column, row, d e / d s factor, ones-column term
  uint32_t o_cur, o_nxt;
  int c_cur, c_nxt, r_cur, r_nxt;
  float f_cur, f_nxt, d_cur, d_nxt;
  const uint32_t lane_off = static_cast<uint32_t>(slab * 32);
  auto load_batch = [&](int bb, uint32_t& o_l, int& c_l, int& r_l, float& f_l, float& d_l) {
    const int e = bb + lane;
    o_l = 0u, c_l = 0, r_l = 0, f_l = 0.f, d_l = 0.f;
    if (e < ee) {
      const int c = __ldg(col + e);
      const float v = __ldg(val + e);
      o_l = (c >= B ? toff4 + static_cast<uint32_t>(c - B) * ldt4 : xoff4 + static_cast<uint32_t>(c) * ldx4) + lane_off;
      const int p = e - eb;
      int w = p >> 5;
      uint32_t m = S.r.endmask[w] >> (p & 31);
      int q = p - 1;
      if (m == 0u) {
        q = (w + 1) * 32 - 1;
        for (++w; w < kRowsChunkMax / 32 && (m = S.r.endmask[w]) == 0u; ++w) q += 32;
      }
      const int row = m ? S.r.row_of[q + __ffs(m)] : rowL;
      float ev = (__ldg(a_l + c) + __ldg(a_r + row)) * inv_sigma;
      const float dl = ev > 0.f ? 1.f : slope;
      ev = ev > 0.f ? ev : slope * ev;
      c_l = c, r_l = row;
      f_l = v * expf(ev) * dl;
      d_l = (slab == 0 && row < B) ? __ldg(dden + row) : 0.f;
    }
  };
  auto refill = [&](auto s0_tag, int j0, uint32_t o_src, int bb) {
    constexpr int S0 = decltype(s0_tag)::value;
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int j = j0 + u + SLOTS;
      const uint32_t o = __shfl_sync(0xffffffffu, o_src, j & 31) + lane;
      if (bb + j < ee && active) rows_cp_async16(slot_s + (S0 + u) * kRowsSlotBytes, base + o);
    }
    rows_cp_commit();
  };
  load_batch(eb, o_cur, c_cur, r_cur, f_cur, d_cur);
  load_batch(eb + 32, o_nxt, c_nxt, r_nxt, f_nxt, d_nxt);
  constexpr_for<0, 4>([&](auto gt) {
    constexpr int g = decltype(gt)::value;
    refill(IntC<g * G>{}, g * G - SLOTS, o_cur, eb);
  });

  // dY'[r] for the lane's four columns: dYn[r] for a batch row, tail_scale * dinfo * Gq[r] otherwise
  auto load_dy = [&](int r) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && r >= 0 && r < R) {
      if (r < B) {
        t = __ldg(reinterpret_cast<const float4*>(dyn + static_cast<int64_t>(r) * lddyn + c0));
      } else {
        t = __ldg(reinterpret_cast<const float4*>(tail_grad + static_cast<int64_t>(r - B) * ld_tail + c0));
        t.x *= ts, t.y *= ts, t.z *= ts, t.w *= ts;
      }
    }
    return t;
  };
  int row_now = __shfl_sync(0xffffffffu, r_cur, 0);
  float4 dy = load_dy(row_now);
  int r_pref = row_now + 1;
  float4 dy_pref = load_dy(r_pref);

  int batch = 0;
  for (int bb = eb; bb < ee; bb += 32, ++batch) {
    const int cnt = min(32, ee - bb);
    const uint32_t em = S.r.endmask[batch];
#pragma unroll 1
    for (int j0 = 0; j0 < 32; j0 += SLOTS) {
      const uint32_t emr = em >> j0;
      const uint32_t o_src = (j0 + SLOTS < 32) ? o_cur : o_nxt;
      constexpr_for<0, 4>([&](auto gt) {
        constexpr int g = decltype(gt)::value;
        rows_cp_wait<3>();
        if (j0 + g * G < cnt) {
#pragma unroll
          for (int u = 0; u < G; ++u) {
            const int sI = g * G + u;
            if (j0 + sI < cnt) {
              float part = 0.f;
              if (active) {
                const float4 a = S.r.ring[sI * 32 + lane];
                part = fmaf(dy.x, a.x, fmaf(dy.y, a.y, fmaf(dy.z, a.z, dy.w * a.w)));
              }
              S.red[(j0 + sI) * 33 + lane] = part;
              if ((emr >> sI) & 1u) {     // the row ends here: the next entry starts the next non-empty row
                const int jn = j0 + sI + 1;
                const int rn = __shfl_sync(0xffffffffu, jn < 32 ? r_cur : r_nxt, jn & 31);
                dy = (rn == r_pref) ? dy_pref : load_dy(rn);
                r_pref = rn + 1;
                dy_pref = load_dy(r_pref);
              }
            }
          }
        }
        refill(IntC<g * G>{}, j0 + g * G, o_src, bb);
      });
    }
    __syncwarp();
    {   // lane L finishes entry L of the batch
      float sdot = 0.f;
#pragma unroll
      for (int l = 0; l < 32; ++l) sdot += S.red[lane * 33 + l];
      if (lane < cnt) {
        const float de = (sdot + d_cur) * f_cur;
        atomicAdd(ds_l + c_cur, de);
        atomicAdd(ds_r + r_cur, de);
      }
    }
    __syncwarp();
    o_cur = o_nxt, c_cur = c_nxt, r_cur = r_nxt, f_cur = f_nxt, d_cur = d_nxt;
    load_batch(bb + 64, o_nxt, c_nxt, r_nxt, f_nxt, d_nxt);
  }
  rows_cp_wait<0>();
}

// (B2) d x from the aggregation: dx[j] = sum_i w_ij dY'[i]  over the transposed CSR (columns j < B)
template <int VEC>
__global__ void __launch_bounds__(kMpWarps * 32)
    gat_bwd_node_kernel(const int32_t* __restrict__ browptr, const int32_t* __restrict__ brow,
                        const float* __restrict__ bval, const int32_t* __restrict__ chunk_row, int n_chunks,
                        int chunk, int nnz, int B, const float* __restrict__ dyn, int64_t lddyn, Codebook cb,
                        int C, int nslab, const float* __restrict__ a_l, const float* __restrict__ a_r,
                        const float* __restrict__ stat, float slope, float tail_scale,
                        const float* __restrict__ dinfo, float* __restrict__ dx, int64_t lddx) {
  const int lane = threadIdx.x & 31;
  const MpTask t = mp_task<VEC>(chunk_row, n_chunks, chunk, nnz, nslab, C, cb.D);
  if (!t.valid) return;
  const float ts = tail_scale * (dinfo ? __ldg(dinfo) : 1.f);
  float acc[VEC], unused[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f, unused[i] = 0.f;
  // transposed: the walker's "row" is the source column j (score a_l), its entries are target rows i (a_r)
  GatWeights pol{a_r, a_l, gat_inv_sigma(stat), slope};
  auto body = [&](const EntryGroup& g) {
    if (t.active)
      gather_accumulate<VEC, false, false>(g, B, dyn, lddyn, cb, cb.D, ts, t.c0, t.k, t.off, acc, unused);
  };
  auto body_dense = [&](const EntryGroup& g) {
    if (t.active)
      gather_accumulate<VEC, false, false, true>(g, B, dyn, lddyn, cb, cb.D, ts, t.c0, t.k, t.off, acc, unused);
  };
  auto flush = [&](int j, bool) {
    if (t.active) red_vec<VEC>(dx + static_cast<int64_t>(j) * lddx + t.c0, acc);  // dx is pre-zeroed
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    pol.den = 0.f;
  };
  if (cb.tail_grad != nullptr)
    walk_rows<false>(t.eb, t.ee, t.row0, B, browptr, brow, bval, nullptr, B, cb.tail_node, lane, pol, body_dense, flush);
  else
    walk_rows<false>(t.eb, t.ee, t.row0, B, browptr, brow, bval, nullptr, B, cb.tail_node, lane, pol, body, flush);
}

// (B3) ds -> da through a = s * sigma, sigma = sqrt(max(a_l)^2+1) sqrt(max(a_r)^2+1).  One CTA.
//   da_l[n] = ds_l[n] / sigma + [n == argmax a_l] * dsigma * ml / sqrt(ml^2+1) * sqrt(mr^2+1)
//   dsigma  = -(sum ds_l a_l + sum ds_r a_r) / sigma^2           (in place: ds_* become da_*)
__global__ void __launch_bounds__(1024)
    gat_score_grad_kernel(int R, const float* __restrict__ a_l, const float* __restrict__ a_r,
                          const float* __restrict__ stat, float* __restrict__ ds_l, float* __restrict__ ds_r) {
  __shared__ double red[32];
  __shared__ int arg[2];
  const int tid = threadIdx.x;
  const float ml = stat[0], mr = stat[1];
  if (tid < 2) arg[tid] = 0x7fffffff;
  __syncthreads();
  double s = 0.0;
  for (int n = tid; n < R; n += blockDim.x) {
    s += static_cast<double>(ds_l[n]) * a_l[n] + static_cast<double>(ds_r[n]) * a_r[n];
    if (a_l[n] == ml) atomicMin(&arg[0], n);
    if (a_r[n] == mr) atomicMin(&arg[1], n);
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid < 32) {
    double v = tid < (blockDim.x >> 5) ? red[tid] : 0.0;
    v = warp_sum(v);
    if (tid == 0) red[0] = v;
  }
  __syncthreads();
  const float ql = sqrtf(ml * ml + 1.f), qr = sqrtf(mr * mr + 1.f);
  const float sigma = ql * qr;
  const float dsigma = static_cast<float>(-red[0] / (static_cast<double>(sigma) * sigma));
  const float inv = 1.f / sigma;
  for (int n = tid; n < R; n += blockDim.x) {
    float dl = ds_l[n] * inv, dr = ds_r[n] * inv;
    if (n == arg[0]) dl += dsigma * (ml / ql) * qr;
    if (n == arg[1]) dr += dsigma * (mr / qr) * ql;
    ds_l[n] = dl, ds_r[n] = dr;
  }
}

// (B4) attention-vector gradients and the score path into d x:
//   datt_l[c] = sum_n da_l[n] Xin[n, c]  (c <= C, Xin[n, C] = 1) ; likewise datt_r
//   dx[n, c] += da_l[n] att_l[c] + da_r[n] att_r[c]   for n < B
// grid-stride over nodes, one warp per node at a time; per-lane column accumulators, reduced per CTA.
template <int VEC, int NSLAB>
__global__ void __launch_bounds__(256)
    gat_att_grad_kernel(int R, int B, const float* __restrict__ x, int64_t ldx, Codebook cb, int C,
                        const float* __restrict__ att_l, const float* __restrict__ att_r,
                        const float* __restrict__ da_l, const float* __restrict__ da_r,
                        float* __restrict__ datt_l, float* __restrict__ datt_r, float* __restrict__ dx,
                        int64_t lddx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gl[NSLAB][VEC], gr[NSLAB][VEC];
  float wl[NSLAB][VEC], wr[NSLAB][VEC];
#pragma unroll
  for (int s = 0; s < NSLAB; ++s) {
    const int c0 = (s * 32 + lane) * VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) gl[s][i] = gr[s][i] = wl[s][i] = wr[s][i] = 0.f;
    if (c0 < C) ld_vec<VEC>(att_l + c0, wl[s]), ld_vec<VEC>(att_r + c0, wr[s]);
  }
  float ones_l = 0.f, ones_r = 0.f;
  for (int n = blockIdx.x * 8 + warp; n < R; n += gridDim.x * 8) {
    const float dl = __ldg(da_l + n), dr = __ldg(da_r + n);
    ones_l += dl, ones_r += dr;
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) {
      const int c0 = (s * 32 + lane) * VEC;
      if (c0 >= C) continue;
      const int k = c0 / cb.D, off = c0 - k * cb.D;
      float v[VEC];
      load_xin<VEC>(n, B, x, ldx, cb, c0, k, off, v);
#pragma unroll
      for (int i = 0; i < VEC; ++i) gl[s][i] = fmaf(dl, v[i], gl[s][i]), gr[s][i] = fmaf(dr, v[i], gr[s][i]);
      if (n < B && dx) {
        float* p = dx + static_cast<int64_t>(n) * lddx + c0;
        float d[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) d[i] = p[i] + dl * wl[s][i] + dr * wr[s][i];
        st_vec<VEC>(p, d);
      }
    }
  }
  // per-CTA reduction through shared memory, then one atomic per column
  __shared__ float sh[8][32 * VEC + 1];
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < VEC; ++i) sh[warp][lane * VEC + i] = pass == 0 ? gl[s][i] : gr[s][i];
      __syncthreads();
      for (int c = threadIdx.x; c < 32 * VEC; c += blockDim.x) {
        const int cc = s * 32 * VEC + c;
        if (cc < C) {
          float tsum = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) tsum += sh[w][c];
          atomicAdd((pass == 0 ? datt_l : datt_r) + cc, tsum);
        }
      }
    }
  }
  __syncthreads();
  if (lane == 0) sh[warp][0] = ones_l, sh[warp][1] = ones_r;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tl = 0.f, tr = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tl += sh[w][0], tr += sh[w][1];
    atomicAdd(datt_l + C, tl);
    atomicAdd(datt_r + C, tr);
  }
}

struct GatShape {
  bool vec4;
  int vec, nslab, C;
};
static GatShape gat_shape(int nb, int D, int Wp, const void* O, std::initializer_list<const void*> ptrs,
                          std::initializer_list<int64_t> lds) {
  bool ok = (D == 4) && (Wp % 4 == 0) && aligned16(O);
  for (const void* p : ptrs) ok = ok && (!p || aligned16(p));
  for (int64_t ld : lds) ok = ok && (ld % 4 == 0);
  GatShape s;
  s.vec4 = ok, s.vec = ok ? 4 : 1, s.C = nb * D, s.nslab = ceil_div(s.C, 32 * s.vec);
  return s;
}

}  // namespace vqgnn

using namespace vqgnn;

// dense tail rows (vqgnn_tail_materialize) are optional everywhere below: NULL = gather codewords per use
static int gat_set_tail(Codebook& cb, const float* tail_feat, const float* tail_grad, int64_t ld_tail) {
  VQ_CHECK_ARG((!tail_feat && !tail_grad) || (ld_tail % 4 == 0 && (!tail_feat || aligned16(tail_feat)) &&
                                               (!tail_grad || aligned16(tail_grad))),
               "gat: dense tail rows must be 16 B aligned");
  cb.tail_feat = tail_feat, cb.tail_grad = tail_grad, cb.ld_tail = ld_tail;
  return VQGNN_OK;
}

extern "C" int vqgnn_gat_scores(int64_t R, int64_t B, const float* x, int64_t ldx, const int32_t* tail_node,
                                const int16_t* codes, const float* O, int nb, int M, int D, int Wp,
                                const float* tail_feat, int64_t ld_tail, const float* att_l, const float* att_r,
                                float* a_l, float* a_r, float* stat, void* stream) {
  VQ_CHECK_ARG(x && codes && O && att_l && att_r && a_l && a_r && stat, "gat_scores: null argument");
  VQ_CHECK_ARG(R >= B && B > 0 && R < (1ll << 31) && nb > 0 && D > 0 && Wp >= 2 * D, "gat_scores: bad sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Codebook cb{tail_node, codes, O, nb, M, D, Wp};
  if (int rc = gat_set_tail(cb, tail_feat, nullptr, ld_tail)) return rc;
  // att vectors hold C+1 floats: rows of the parameter are not 16 B aligned in general -> scalar loads there
  const GatShape g = gat_shape(nb, D, Wp, O, {x, att_l, att_r}, {ldx});
  gat_stat_init_kernel<<<1, 32, 0, s>>>(stat);
  VQ_LAUNCH_CHECK();
  const int grid = ceil_div(R, 8);
  if (g.vec4) gat_scores_kernel<4><<<grid, 256, 0, s>>>((int)R, (int)B, x, ldx, cb, g.C, att_l, att_r, a_l, a_r, stat);
  else gat_scores_kernel<1><<<grid, 256, 0, s>>>((int)R, (int)B, x, ldx, cb, g.C, att_l, att_r, a_l, a_r, stat);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_gat_fwd(const int32_t* rowptr, const int32_t* col, const float* val,
                             const int32_t* chunk_row, int chunk, int64_t nnz, int64_t R, int64_t B,
                             const float* x, int64_t ldx, const int32_t* tail_node, const int16_t* codes,
                             const float* O, int nb, int M, int D, int Wp, const float* tail_feat, int64_t ld_tail,
                             const float* a_l, const float* a_r, const float* stat, float negative_slope,
                             float info_scale, float* y, int64_t ldy, float* den, float* info, void* ws,
                             void* stream) {
  VQ_CHECK_ARG(rowptr && col && val && x && codes && O && a_l && a_r && stat && y && den, "gat_fwd: null argument");
  VQ_CHECK_ARG(R >= B && B > 0 && nb > 0 && D > 0 && Wp >= 2 * D, "gat_fwd: bad sizes");
  VQ_CHECK_ARG(!info || ws, "gat_fwd: info needs a workspace");
  VQ_CHECK_ARG(R < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "gat_fwd: sizes must fit int32");
  VQ_CHECK_ARG(chunk > 0 && chunk % 32 == 0 && (nnz == 0 || chunk_row), "gat_fwd: needs chunk_row");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Codebook cb{tail_node, codes, O, nb, M, D, Wp};
  if (int rc = gat_set_tail(cb, tail_feat, nullptr, ld_tail)) return rc;
  const GatShape g = gat_shape(nb, D, Wp, O, {x, y}, {ldx, ldy});
  double* ws_sum = static_cast<double*>(ws);
  unsigned int* ws_count = ws ? reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 8) : nullptr;
  if (info) VQ_CUDA(cudaMemsetAsync(ws, 0, 16, s));
  if (int rc = zero_rows(y, B, g.C, ldy, s)) return rc;
  VQ_CUDA(cudaMemsetAsync(den, 0, sizeof(float) * B, s));
  const int n_chunks = static_cast<int>((nnz + chunk - 1) / chunk);
  if (n_chunks == 0) {
    if (info) VQ_CUDA(cudaMemsetAsync(info, 0, sizeof(float), s));
    return VQGNN_OK;
  }
  const int grid = ceil_div(static_cast<int64_t>(n_chunks) * g.nslab, kMpWarps);
#define VQ_GAT_FWD(VEC)                                                                                     \
  gat_fwd_kernel<VEC><<<grid, kMpWarps * 32, 0, s>>>(rowptr, col, val, chunk_row, n_chunks, chunk, (int)nnz, R, \
                                                     (int)B, x, ldx, cb, g.C, g.nslab, a_l, a_r, stat,      \
                                                     negative_slope, info_scale, y, ldy, den, info, ws_sum, \
                                                     ws_count)
  if (g.vec4) VQ_GAT_FWD(4);
  else VQ_GAT_FWD(1);
#undef VQ_GAT_FWD
  VQ_LAUNCH_CHECK();
  const int ngrid = static_cast<int>(std::min<int64_t>((B * g.C + 255) / 256, 8 * kNumSMs));
  gat_normalize_kernel<<<ngrid, 256, 0, s>>>(B, g.C, y, ldy, den);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

// The same forward through the lean row-gather kernel of mp_rows.cuh (materialised codeword rows, cp.async rings): the
// GAT weights are computed once per entry by the lane that holds it.
extern "C" int vqgnn_gat_fwd_rows(const int32_t* rowptr, const int32_t* col, const float* val,
                                  const int32_t* chunk_row, int chunk, int64_t nnz, int64_t R, int64_t B,
                                  const float* x, int64_t ldx, const float* tail_feat, int64_t T,
                                  const float* tail_grad, int64_t ld_tail, int C, const float* a_l, const float* a_r,
                                  const float* stat, float negative_slope, float info_scale, float* y, int64_t ldy,
                                  float* den, float* info, void* ws, size_t ws_bytes, void* stream) {
  VQ_CHECK_ARG(rowptr && x && a_l && a_r && stat && y && den && (nnz == 0 || (col && val)), "gat_fwd_rows: null argument");
  VQ_CHECK_ARG(R >= B && B > 0 && R < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "gat_fwd_rows: bad sizes");
  VQ_CHECK_ARG(T >= 0 && (T == 0 || tail_feat) && (!info || R == B || tail_grad),
               "gat_fwd_rows: needs the materialised codeword rows (tail_feat; tail_grad for info)");
  VQ_CHECK_ARG(C >= 16 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ld_tail % 4 == 0 && aligned16(x) &&
                   aligned16(y) && (!tail_feat || aligned16(tail_feat)) && (!tail_grad || aligned16(tail_grad)),
               "gat_fwd_rows: needs C >= 16, C % 4 == 0 and 16 B aligned rows");
  VQ_CHECK_ARG(chunk > 0 && chunk % 32 == 0 && chunk <= kRowsChunkMax && (nnz == 0 || chunk_row),
               "gat_fwd_rows: needs chunk_row with chunk <= 256");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = zero_rows(y, B, C, ldy, s)) return rc;
  VQ_CUDA(cudaMemsetAsync(den, 0, sizeof(float) * B, s));
  const int n_chunks = static_cast<int>((nnz + chunk - 1) / chunk);
  if (n_chunks == 0) {
    if (info) VQ_CUDA(cudaMemsetAsync(info, 0, sizeof(float), s));
    return VQGNN_OK;
  }
  const int nslab = ceil_div(C, 128);
  const int64_t tasks = static_cast<int64_t>(n_chunks) * nslab;
  constexpr int NW = kRowsWarps, SLOTS = kRowsSlots;
  const int grid = ceil_div(tasks, NW);
  // workspace: [count 256 B][per-block info partials]
  VQ_CHECK_ARG(!info || (ws && ws_bytes >= 512 + static_cast<size_t>(grid) * 8), "gat_fwd_rows: workspace too small");
  char* wp = ws ? reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255)) : nullptr;
  unsigned int* ws_count = reinterpret_cast<unsigned int*>(wp);
  double* ws_part = reinterpret_cast<double*>(wp + 256);
  if (info) VQ_CUDA(cudaMemsetAsync(ws_count, 0, 16, s));
  const uintptr_t xa = reinterpret_cast<uintptr_t>(x), ta = tail_feat ? reinterpret_cast<uintptr_t>(tail_feat) : xa;
  const uintptr_t base = std::min(xa, ta);
  const uint64_t x_end4 = (xa - base) / 16 + static_cast<uint64_t>(B) * (ldx / 4) + 32;
  const uint64_t t_end4 = (ta - base) / 16 + static_cast<uint64_t>(T) * (ld_tail / 4) + 32;
  VQ_CHECK_ARG(x_end4 < (1ull << 32) && t_end4 < (1ull << 32),
               "gat_fwd_rows: x and tail_feat must lie within 64 GB of each other (32-bit row offsets)");
  const size_t smem = sizeof(RowsWarpSmem<SLOTS>) * NW;
  RowsGat gp{a_l, a_r, stat, negative_slope, den};
  mp_fwd_rows_kernel<NW, SLOTS, true><<<grid, NW * 32, smem, s>>>(
      rowptr, col, val, chunk_row, n_chunks, chunk, (int)nnz, (int)R, (int)B, reinterpret_cast<const float4*>(base),
      static_cast<uint32_t>((xa - base) / 16), static_cast<uint32_t>(ldx / 4), static_cast<uint32_t>((ta - base) / 16),
      static_cast<uint32_t>(ld_tail / 4), 1.0f, nullptr, tail_grad, ld_tail, C, nslab, info_scale, y, ldy, info, ws_part,
      ws_count, nullptr, gp);
  VQ_LAUNCH_CHECK();
  const int ngrid = static_cast<int>(std::min<int64_t>((B * C + 255) / 256, 8 * kNumSMs));
  gat_normalize_kernel<<<ngrid, 256, 0, s>>>(B, C, y, ldy, den);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_gat_bwd(const int32_t* rowptr, const int32_t* col, const float* val,
                             const int32_t* chunk_row, int64_t nnz, int64_t R, const int32_t* browptr,
                             const int32_t* brow, const float* bval, const int32_t* bchunk_row, int64_t bnnz,
                             int chunk, int64_t B, const float* x, int64_t ldx, const int32_t* tail_node,
                             const int16_t* codes, const float* O, int nb, int M, int D, int Wp,
                             const float* tail_feat, const float* tail_grad, int64_t ld_tail,
                             const float* att_l, const float* att_r, const float* a_l, const float* a_r,
                             const float* stat, float negative_slope, const float* out, int64_t ldo,
                             const float* den, const float* dout, int64_t lddo, float tail_scale,
                             const float* dinfo, float* dyn, int64_t lddyn, float* dden, float* ds_l,
                             float* ds_r, float* dx, int64_t lddx, float* datt_l, float* datt_r,
                             void* stream) {
  VQ_CHECK_ARG(rowptr && col && val && browptr && brow && bval && x && codes && O && att_l && att_r && a_l &&
                   a_r && stat && out && den && dout && dyn && dden && ds_l && ds_r && datt_l && datt_r,
               "gat_bwd: null argument");
  VQ_CHECK_ARG(R >= B && B > 0 && R < (1ll << 31) && nb > 0 && D > 0 && Wp >= 2 * D, "gat_bwd: bad sizes");
  VQ_CHECK_ARG(nnz >= 0 && nnz < (1ll << 31) && bnnz >= 0 && bnnz < (1ll << 31), "gat_bwd: nnz must fit int32");
  VQ_CHECK_ARG(chunk > 0 && chunk % 32 == 0 && (nnz == 0 || chunk_row) && (bnnz == 0 || bchunk_row),
               "gat_bwd: needs chunk rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Codebook cb{tail_node, codes, O, nb, M, D, Wp};
  if (int rc = gat_set_tail(cb, tail_feat, tail_grad, ld_tail)) return rc;
  const GatShape g = gat_shape(nb, D, Wp, O, {x, dyn, dx, att_l, att_r}, {ldx, lddyn, dx ? lddx : 0});
  VQ_CHECK_ARG(g.nslab <= 8, "gat_bwd: at most %d columns are supported", 8 * 32 * g.vec);
  const int C = g.C;
  gat_bwd_prep_kernel<<<ceil_div(B, 8), 256, 0, s>>>((int)B, C, dout, lddo, out, ldo, den, dyn, lddyn, dden);
  VQ_LAUNCH_CHECK();
  VQ_CUDA(cudaMemsetAsync(ds_l, 0, sizeof(float) * R, s));
  VQ_CUDA(cudaMemsetAsync(ds_r, 0, sizeof(float) * R, s));
  VQ_CUDA(cudaMemsetAsync(datt_l, 0, sizeof(float) * (C + 1), s));
  VQ_CUDA(cudaMemsetAsync(datt_r, 0, sizeof(float) * (C + 1), s));
  const int n_chunks = static_cast<int>((nnz + chunk - 1) / chunk);
  // materialised codeword rows within one 64 GB window of x: the lean edge kernel
  bool edge_rows = false;
  if (n_chunks > 0 && g.vec4 && tail_feat && tail_grad && C % 4 == 0 && chunk <= kRowsChunkMax && ldx % 4 == 0 &&
      lddyn % 4 == 0 && ld_tail % 4 == 0 && aligned16(x) && aligned16(dyn) && aligned16(tail_feat) &&
      aligned16(tail_grad)) {
    const uintptr_t xa = reinterpret_cast<uintptr_t>(x), ta = reinterpret_cast<uintptr_t>(tail_feat);
    const uintptr_t base = std::min(xa, ta);
    const uint64_t x_end4 = (xa - base) / 16 + static_cast<uint64_t>(B) * (ldx / 4) + 32;
    const uint64_t t_end4 = (ta - base) / 16 + static_cast<uint64_t>(R - B) * (ld_tail / 4) + 32;
    if (x_end4 < (1ull << 32) && t_end4 < (1ull << 32)) {
      edge_rows = true;
      const int nslab = ceil_div(C, 128);
      const int64_t tasks = static_cast<int64_t>(n_chunks) * nslab;
      const size_t smem = sizeof(GatEdgeSmem) * kEdgeWarps;
      static bool attr_set = false;
      if (!attr_set) {
        VQ_CUDA(cudaFuncSetAttribute(gat_bwd_edge_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      gat_bwd_edge_rows_kernel<<<ceil_div(tasks, kEdgeWarps), kEdgeWarps * 32, smem, s>>>(
          rowptr, col, val, chunk_row, n_chunks, chunk, (int)nnz, (int)R, (int)B, reinterpret_cast<const float4*>(base),
          static_cast<uint32_t>((xa - base) / 16), static_cast<uint32_t>(ldx / 4),
          static_cast<uint32_t>((ta - base) / 16), static_cast<uint32_t>(ld_tail / 4), tail_grad, ld_tail, C, nslab, a_l,
          a_r, stat, negative_slope, dyn, lddyn, dden, tail_scale, dinfo, ds_l, ds_r);
      VQ_LAUNCH_CHECK();
    }
  }
  if (n_chunks > 0 && !edge_rows) {
    const int grid = ceil_div(static_cast<int64_t>(n_chunks) * g.nslab, kMpWarps);
#define VQ_GAT_EDGE(VEC)                                                                                      \
  gat_bwd_edge_kernel<VEC><<<grid, kMpWarps * 32, 0, s>>>(rowptr, col, val, chunk_row, n_chunks, chunk,        \
                                                          (int)nnz, R, (int)B, x, ldx, cb, C, g.nslab, a_l, a_r, \
                                                          stat, negative_slope, dyn, lddyn, dden, tail_scale, \
                                                          dinfo, ds_l, ds_r)
    if (g.vec4) VQ_GAT_EDGE(4);
    else VQ_GAT_EDGE(1);
#undef VQ_GAT_EDGE
    VQ_LAUNCH_CHECK();
  }
  bool node_rows = false;
  if (dx) {
    if (int rc = zero_rows(dx, B, C, lddx, s)) return rc;
    const int bn_chunks = static_cast<int>((bnnz + chunk - 1) / chunk);
    // (B2, lean) the transposed weighted SpMM through the row-gather kernel: rows = batch columns j (score a_l), entries =
    // target rows i (score a_r), operand rows dYn[i] / tail_scale * dinfo * Gq[i]
    if (bn_chunks > 0 && edge_rows && lddx % 4 == 0 && aligned16(dx)) {
      const uintptr_t xa = reinterpret_cast<uintptr_t>(dyn), ta = reinterpret_cast<uintptr_t>(tail_grad);
      const uintptr_t base = std::min(xa, ta);
      const uint64_t x_end4 = (xa - base) / 16 + static_cast<uint64_t>(B) * (lddyn / 4) + 32;
      const uint64_t t_end4 = (ta - base) / 16 + static_cast<uint64_t>(R - B) * (ld_tail / 4) + 32;
      if (x_end4 < (1ull << 32) && t_end4 < (1ull << 32)) {
        node_rows = true;
        constexpr int NW = kRowsWarps, SLOTS = kRowsSlots;
        const int nslab = ceil_div(C, 128);
        const int64_t tasks = static_cast<int64_t>(bn_chunks) * nslab;
        const size_t smem = sizeof(RowsWarpSmem<SLOTS>) * NW;
        RowsGat gp{a_r, a_l, stat, negative_slope, nullptr};
        mp_fwd_rows_kernel<NW, SLOTS, true><<<ceil_div(tasks, NW), NW * 32, smem, s>>>(
            browptr, brow, bval, bchunk_row, bn_chunks, chunk, (int)bnnz, (int)B, (int)B,
            reinterpret_cast<const float4*>(base), static_cast<uint32_t>((xa - base) / 16),
            static_cast<uint32_t>(lddyn / 4), static_cast<uint32_t>((ta - base) / 16),
            static_cast<uint32_t>(ld_tail / 4), tail_scale, dinfo, nullptr, ld_tail, C, nslab, 1.0f, dx, lddx, nullptr,
            nullptr, nullptr, nullptr, gp);
        VQ_LAUNCH_CHECK();
      }
    }
    if (bn_chunks > 0 && !node_rows) {
      const int grid = ceil_div(static_cast<int64_t>(bn_chunks) * g.nslab, kMpWarps);
#define VQ_GAT_NODE(VEC)                                                                                     \
  gat_bwd_node_kernel<VEC><<<grid, kMpWarps * 32, 0, s>>>(browptr, brow, bval, bchunk_row, bn_chunks, chunk,  \
                                                          (int)bnnz, (int)B, dyn, lddyn, cb, C, g.nslab, a_l, \
                                                          a_r, stat, negative_slope, tail_scale, dinfo, dx,  \
                                                          lddx)
      if (g.vec4) VQ_GAT_NODE(4);
      else VQ_GAT_NODE(1);
#undef VQ_GAT_NODE
      VQ_LAUNCH_CHECK();
    }
  }
  gat_score_grad_kernel<<<1, 1024, 0, s>>>((int)R, a_l, a_r, stat, ds_l, ds_r);
  VQ_LAUNCH_CHECK();
  const int agrid = static_cast<int>(std::min<int64_t>(ceil_div(R, 8), 2 * kNumSMs));
#define VQ_GAT_ATT(VEC, NS)                                                                                 \
  gat_att_grad_kernel<VEC, NS><<<agrid, 256, 0, s>>>((int)R, (int)B, x, ldx, cb, C, att_l, att_r, ds_l, ds_r, \
                                                     datt_l, datt_r, dx, lddx)
  if (g.vec4) {
    if (g.nslab <= 1) VQ_GAT_ATT(4, 1);
    else if (g.nslab <= 2) VQ_GAT_ATT(4, 2);
    else if (g.nslab <= 4) VQ_GAT_ATT(4, 4);
    else VQ_GAT_ATT(4, 8);
  } else {
    if (g.nslab <= 1) VQ_GAT_ATT(1, 1);
    else if (g.nslab <= 2) VQ_GAT_ATT(1, 2);
    else if (g.nslab <= 4) VQ_GAT_ATT(1, 4);
    else VQ_GAT_ATT(1, 8);
  }
#undef VQ_GAT_ATT
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
