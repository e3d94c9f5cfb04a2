// Shared device code of the message-passing kernels (mp.cu: GCN / SAGE-Mean, gat.cu: GAT).
//
// Work unit = one warp x (chunk of `chunk` consecutive CSR entries) x (slab of 32*VEC columns).  Chunks ignore
// row boundaries, so a power-law hub row is spread over many warps; `walk_rows` tracks the rows a chunk
// crosses and tells the caller when a row (or the part of it inside the chunk) is complete.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace vqgnn {

constexpr int kMpWarps = 8;
constexpr int kMpUnroll = 4;

struct Codebook {
  const int32_t* tail_node;  // [T] or nullptr (identity)
  const int16_t* codes;      // [N, nb]
  const float* O;            // [nb, M, Wp]
  int nb, M, D, Wp;
  // optional dense copies of the tail entries' codewords (vqgnn_tail_materialize): row t = tail entry t.  When a
  // tail node is referenced by many edges (v2 batch graphs: ~20x), gathering its codewords ONCE and reading a
  // coalesced dense row per edge replaces nb scattered 32 B sectors per edge by C*4 contiguous bytes.
  const float* tail_feat = nullptr;  // [T, ld_tail] feature halves
  const float* tail_grad = nullptr;  // [T, ld_tail] gradient halves
  int64_t ld_tail = 0;
  // tail_slab > 0: the two tables are SLAB-MAJOR, [ceil(C / tail_slab)][T][tail_slab] (vqgnn_tail_materialize_slab),
  // and ld_tail holds T
  int tail_slab = 0;
  __device__ __forceinline__ const float* tail_row(const float* t, int te, int c0) const {
    if (tail_slab == 0) return t + static_cast<int64_t>(te) * ld_tail + c0;
    const int sl = c0 / tail_slab;
    return t + (static_cast<int64_t>(sl) * ld_tail + te) * tail_slab + (c0 - sl * tail_slab);
  }
};

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = __ldg(p + i);
  }
}
template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = v[i];
  }
}
template <int VEC>
__device__ __forceinline__ void red_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3])
                 : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) atomicAdd(p + i, v[i]);
  }
}
// one 32 B sector: feature half -> a, gradient half -> b (Wp == 8, D == 4)
__device__ __forceinline__ void ld_sector(const float* p, float (&a)[4], float (&b)[4]) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3])
               : "l"(p));
}

// Per-entry weight policy.  The plain SpMM uses the stored value; the GAT kernels multiply it by
// exp(leaky_relu(.)) of a per-column score (fetched together with the entry) and a per-row score.
struct PlainWeights {
  __device__ __forceinline__ float load_extra(int) const { return 0.f; }
  __device__ __forceinline__ void row_begin(int) {}
  __device__ __forceinline__ float weight(float v, float, bool) { return v; }
};

// A group of up to U consecutive entries of one row, broadcast to every lane of the warp.
//   c[u]    column id (< B: dense row, >= B: tail entry, -1: padding)      v[u]  policy-weighted value
//   node[u] global node id of a tail entry                                   rv[u] reverse value (HAS_RV)
//   xtra[u] the policy's per-entry extra (e.g. the source score)            raw[u] the stored value
struct EntryGroup {
  int c[kMpUnroll], node[kMpUnroll];
  float v[kMpUnroll], rv[kMpUnroll], xtra[kMpUnroll], raw[kMpUnroll];
};

// Walks the CSR entries [eb, ee) starting in row r.  Calls
//   body(group)            for every group of <= U entries of the current row,
//   flush(row, whole)      when the row ends inside the chunk (whole = it also started inside, so no other
//                          warp touches its output) and once, with whole = false, for a trailing partial row;
//                          a flush taking (row, whole, rs, re) also receives the row's entry range [rs, re),
//   pol.row_begin(row)     whenever the current row changes (before its first entry).
// All control flow is warp-uniform.
template <bool HAS_RV, class Policy, class Body, class Flush>
__device__ __forceinline__ void walk_rows(int eb, int ee, int r, int64_t R, const int32_t* __restrict__ rowptr,
                                          const int32_t* __restrict__ col, const float* __restrict__ val,
                                          const float* __restrict__ rval, int B,
                                          const int32_t* __restrict__ tail_node, int lane, Policy& pol,
                                          Body&& body, Flush&& flush) {
  constexpr int U = kMpUnroll;
  // row boundaries: lane i holds rowptr[rbase + i]; row r is [shfl(r - rbase), shfl(r - rbase + 1))
  int rbase = r;
  int rp_l = __ldg(rowptr + min(static_cast<int64_t>(rbase) + lane, R));
  int rs = __shfl_sync(0xffffffffu, rp_l, 0), re = __shfl_sync(0xffffffffu, rp_l, 1);
  bool pending = false;
  pol.row_begin(r);

  for (int bb = eb; bb < ee; bb += 32) {
    const int e = bb + lane;
    int c_l = -1, node_l = 0;
    float v_l = 0.f, rv_l = 0.f, x_l = 0.f;
    if (e < ee) {
      c_l = __ldg(col + e);
      v_l = __ldg(val + e);
      x_l = pol.load_extra(c_l);
      if (HAS_RV) rv_l = __ldg(rval + e);
      if (c_l >= B) node_l = tail_node ? __ldg(tail_node + (c_l - B)) : (c_l - B);
    }
    auto do_flush = [&](int row, bool whole) {
      if constexpr (std::is_invocable_v<Flush&, int, bool, int, int>) flush(row, whole, rs, re);
      else flush(row, whole);
    };
    const int cnt = min(32, ee - bb);
    int j = 0;
    while (j < cnt) {
      const int jend = min(cnt, re - bb);  // entries of row r inside this batch end here
      for (; j < jend; j += U) {
        EntryGroup g;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int src_lane = min(j + u, 31);
          const bool valid = j + u < jend;
          g.c[u] = valid ? __shfl_sync(0xffffffffu, c_l, src_lane) : -1;
          g.raw[u] = __shfl_sync(0xffffffffu, v_l, src_lane);
          g.xtra[u] = __shfl_sync(0xffffffffu, x_l, src_lane);
          g.v[u] = pol.weight(g.raw[u], g.xtra[u], valid);
          g.node[u] = __shfl_sync(0xffffffffu, node_l, src_lane);
          g.rv[u] = HAS_RV ? __shfl_sync(0xffffffffu, rv_l, src_lane) : 0.f;
        }
        body(g);
      }
      j = jend;
      pending = true;
      if (bb + j == re) {  // row r is complete
        do_flush(r, rs >= eb);
        pending = false;
        if (bb + j >= ee) break;
        do {  // next non-empty row (empty rows keep the pre-initialised output)
          ++r;
          if (r - rbase >= 31) {
            rbase = r;
            rp_l = __ldg(rowptr + min(static_cast<int64_t>(rbase) + lane, R));
          }
          rs = __shfl_sync(0xffffffffu, rp_l, r - rbase);
          re = __shfl_sync(0xffffffffu, rp_l, r - rbase + 1);
        } while (re <= bb + j);
        pol.row_begin(r);
      }
    }
  }
  if (pending) {
    if constexpr (std::is_invocable_v<Flush&, int, bool, int, int>) flush(r, false, rs, re);
    else flush(r, false);
  }
}

// How a (partial) row output must be combined so that the result does not depend on the order in which warps
// retire.  A row cut by ONE chunk boundary has two pieces: two REDs onto a zero-initialised output commute exactly
// ((0 + a) + b == (0 + b) + a).  A row spanning >= 3 chunks would depend on the RED order: its pieces are stored in
// a piece buffer (slot 1: the piece holding the row start, slot 0: any later piece) and summed in chunk order by the
// fix-up kernel.
enum PieceKind { kPieceWhole = 0, kPieceRed = 1, kPieceHubStart = 2, kPieceHubNext = 3 };
__device__ __forceinline__ int piece_kind(bool whole, int rs, int re, int eb, int chunk) {
  if (whole) return kPieceWhole;
  const int fc = rs / chunk, lc = (re - 1) / chunk;
  if (lc - fc <= 1) return kPieceRed;
  return rs >= eb ? kPieceHubStart : kPieceHubNext;
}

// block-level fp64 reduction of the info partials; every block stores its partial, the last block to arrive adds
// them in block order (fixed tree): the scalar is bit-stable
__device__ __forceinline__ void info_reduce_ordered(double part, double* ws_part, unsigned int* ws_count,
                                                    float info_scale, float* info) {
  __shared__ double sh[kMpWarps];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  part = warp_sum(part);
  if (lane == 0) sh[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kMpWarps; ++i) t += sh[i];
    ws_part[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(ws_count, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double t = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(ws_part + i);
  t = warp_sum(t);
  __syncthreads();
  if (lane == 0) sh[warp] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
#pragma unroll
    for (int i = 0; i < kMpWarps; ++i) total += sh[i];
    *info = static_cast<float>(static_cast<double>(info_scale) * total);
  }
}

// The gather-accumulate body shared by every SpMM-shaped kernel: for the lane's VEC columns starting at c0
// (branch k, offset off)
//   acc += v * (c < B ? dense[c, c0..] : tscale * O_k[code, half_off + off ..])
//   gqa += rv * O_k[code, D + off ..]                                   (HAS_GQ, tail entries only)
// WIDE: VEC == 4, HAS_GQ, Wp == 8, D == 4, half_off == 0 -> one 256-bit load per gathered codeword.
// DENSE_TAIL: tail entries read row (c - B) of cb.tail_feat (half_off == 0) / cb.tail_grad instead of gathering.
template <int VEC, bool HAS_GQ, bool WIDE, bool DENSE_TAIL = false>
__device__ __forceinline__ void gather_accumulate(const EntryGroup& g, int B, const float* __restrict__ dense,
                                                  int64_t ldd, const Codebook& cb, int half_off, float tscale,
                                                  int c0, int k, int off, float (&acc)[VEC], float (&gqa)[VEC]) {
  constexpr int U = kMpUnroll;
  const float* p[U];
  if constexpr (DENSE_TAIL) {   // tail rows were materialised (vqgnn_tail_materialize): plain coalesced row reads
    static_assert(!WIDE && !HAS_GQ, "dense tail rows carry one half only");
    const float* tdense = half_off == 0 ? cb.tail_feat : cb.tail_grad;
    float a[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (g.c[u] >= B) ld_vec<VEC>(cb.tail_row(tdense, g.c[u] - B, c0), a[u]);
      else if (g.c[u] >= 0) ld_vec<VEC>(dense + static_cast<int64_t>(g.c[u]) * ldd + c0, a[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (g.c[u] >= 0) {
        const float s = g.c[u] >= B ? g.v[u] * tscale : g.v[u];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(s, a[u][i], acc[i]);
      }
    }
    return;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {  // first level: code loads for tail entries (independent)
    p[u] = nullptr;
    if (g.c[u] >= B) {
      const int code = __ldg(cb.codes + static_cast<int64_t>(g.node[u]) * cb.nb + k);
      p[u] = cb.O + (static_cast<int64_t>(k) * cb.M + code) * cb.Wp + off;
    } else if (g.c[u] >= 0) {
      p[u] = dense + static_cast<int64_t>(g.c[u]) * ldd + c0;
    }
  }
  float a[U][VEC], q[U][VEC];
#pragma unroll
  for (int u = 0; u < U; ++u) {  // second level: the gathers
    if (g.c[u] >= B) {
      if constexpr (WIDE) {
        ld_sector(p[u], a[u], q[u]);
      } else {
        ld_vec<VEC>(p[u] + half_off, a[u]);
        if (HAS_GQ) ld_vec<VEC>(p[u] + cb.D, q[u]);
      }
    } else if (g.c[u] >= 0) {
      ld_vec<VEC>(p[u], a[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (g.c[u] >= B) {
      const float s = g.v[u] * tscale;
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(s, a[u][i], acc[i]);
      if (HAS_GQ) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) gqa[i] = fmaf(g.rv[u], q[u][i], gqa[i]);
      }
    } else if (g.c[u] >= 0) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(g.v[u], a[u][i], acc[i]);
    }
  }
}

// block-level fp64 reduction of the info partials + "last block finishes" epilogue
__device__ __forceinline__ void info_reduce(double part, double* ws_sum, unsigned int* ws_count,
                                            float info_scale, float* info) {
  __shared__ double sh[kMpWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  part = warp_sum(part);
  if (lane == 0) sh[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kMpWarps; ++i) t += sh[i];
    atomicAdd(ws_sum, t);
    __threadfence();
    const unsigned int ticket = atomicAdd(ws_count, 1u);
    if (ticket == gridDim.x - 1) {
      const double total = atomicAdd(ws_sum, 0.0);
      *info = static_cast<float>(static_cast<double>(info_scale) * total);
    }
  }
}

// (chunk, slab) of a warp task; slab-major so the warps of a CTA share a slab (same codebook branches)
struct MpTask {
  int slab, eb, ee, row0, c0, k, off;
  bool valid, active;
};
template <int VEC>
__device__ __forceinline__ MpTask mp_task(const int32_t* __restrict__ chunk_row, int n_chunks, int chunk, int nnz,
                                          int nslab, int C, int D) {
  MpTask t;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = static_cast<int64_t>(blockIdx.x) * kMpWarps + warp;
  t.valid = task < static_cast<int64_t>(n_chunks) * nslab;
  t.slab = t.valid ? static_cast<int>(task / n_chunks) : 0;
  const int ch = t.valid ? static_cast<int>(task - static_cast<int64_t>(t.slab) * n_chunks) : 0;
  t.c0 = (t.slab * 32 + lane) * VEC;
  t.active = t.valid && t.c0 < C;
  t.k = t.active ? t.c0 / D : 0;
  t.off = t.active ? t.c0 - t.k * D : 0;
  t.eb = ch * chunk;
  t.ee = min(t.eb + chunk, nnz);
  t.row0 = t.valid ? __ldg(chunk_row + ch) : 0;
  return t;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

static inline int zero_rows(float* p, int64_t rows, int C, int64_t ld, cudaStream_t s) {
  if (ld == C) {
    VQ_CUDA(cudaMemsetAsync(p, 0, sizeof(float) * rows * C, s));
  } else {
    VQ_CUDA(cudaMemset2DAsync(p, sizeof(float) * ld, 0, sizeof(float) * C, rows, s));
  }
  return VQGNN_OK;
}

}  // namespace vqgnn
