// v2 forward of large batch graphs with materialised codeword rows: lean asynchronous row gathers.
//
// csrc/mp.cu's mp_fwd_async_kernel was ISSUE bound (ncu, products shape: 84 warp instructions per CSR entry, 89 % of the
// issue slots busy, L2 at 45 %, DRAM at 34 %): the sequential row walk, the per-entry address arithmetic and one
// commit / wait pair per entry.  Same work partition here (warp x 256-entry CSR chunk x 128-column slab), same in-order
// accumulation and the same order-independent piece scheme as mp_fwd_kernel, but
//   * a warp owns a small ring of row slots (kRowsSlots = 8 x 512 B) in four cp.async groups; entry n of the chunk uses
//     slot n % 8, and a group is refilled (16 B per lane, one commit per group, one wait per group) as soon as it has
//     been consumed: 6..8 rows in flight per warp, 32 warps per SM.  Measured: small rings with many warps beat deep
//     rings (32 slots x 12 warps: 1.54 ms, 8 slots x 32 warps: 0.82 ms per launch at the products shape);
//   * every lane precomputes the 32-bit offset (in 16 B units from a common base) of ITS entry's row once per batch, so
//     issuing a row is SHFL + IADD + IMAD.WIDE + LDGSTS;
//   * row ends are marked once per task from the row pointers (a 256-bit mask + the row id of every end position in
//     shared memory), so the consuming loop (unrolled over one ring round) costs one predicate test per entry instead
//     of the sequential row walk: SHFL (value) + LDS.128 + 4 FFMA + test;
//   * the gradient codeword row of an out-of-batch row (the other operand of info_backward) is read from the
//     materialised table and prefetched one row ahead.
// The same kernel, with the entry value replaced by the GAT weight (template flag GAT), is the GAT forward and the GAT
// backward's transposed SpMM (csrc/gat.cu).  Now DRAM bound at the products shape (ncu: 63 % of the DRAM peak, the
// 460 MB row table does not fit the L2).
// Tried first and dropped: one `cp.async.bulk` (TMA) per row completing on an mbarrier -- correct but slower than the
// kernel it was to replace (1.59 vs 1.36 ms): 512 B bulk copies are bound by the per-SM TMA request rate (~1 per 29 clk).
//
// Reference maths: vq_gnn_v2/models.py:161-198 (conv over [x ; codewords], info_backward), vq_gnn_v2/convs.py:65-101.
#pragma once
#include "mp_common.cuh"

namespace vqgnn {

constexpr int kRowsChunkMax = 256;   // entries per task (the plan's MP_CHUNK)
constexpr int kRowsSlotBytes = 512;  // one slab of a row: 128 fp32 columns
constexpr int kRowsWarps = 4;        // warps per CTA
constexpr int kRowsSlots = 8;        // ring slots (rows in flight) per warp: four cp.async groups of two

template <int SLOTS>
struct __align__(128) RowsWarpSmem {
  float4 ring[SLOTS * 32];           // SLOTS x 512 B
  int32_t row_of[kRowsChunkMax];     // row id of every row-end position of the task
  uint32_t endmask[kRowsChunkMax / 32];
};

__device__ __forceinline__ uint32_t rows_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void rows_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void rows_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rows_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// compile-time loop: f(IntC<I>) for I in [I0, I1)
template <int V>
struct IntC {
  static constexpr int value = V;
};
template <int I0, int I1, class F>
__device__ __forceinline__ void constexpr_for(F&& f) {
  if constexpr (I0 < I1) {
    f(IntC<I0>{});
    constexpr_for<I0 + 1, I1>(f);
  }
}

// per-block fp64 partial of the info scalar, blocks added in index order by the last block (cf. info_reduce_ordered)
template <int NW>
__device__ __forceinline__ void rows_info_reduce(double part, double* ws_part, unsigned int* ws_count,
                                                 float info_scale, float* info) {
  __shared__ double sh[NW];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  part = warp_sum(part);
  if (lane == 0) sh[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += sh[i];
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
    for (int i = 0; i < NW; ++i) total += sh[i];
    *info = static_cast<float>(static_cast<double>(info_scale) * total);
  }
}

// base: a 16 B aligned address below both tables; xoff4 / toff4: offsets of x / tail_feat from it in 16 B units;
// ldx4 / ldt4: row strides in 16 B units.  Every gathered piece must lie below base + 2^32 * 16 B (checked on the host).
// GAT (vq_gnn_v2/convs.py:165-266 with vq_softmax == un-normalised exp): the value of entry (i, j) becomes
// val * exp(leaky_relu((a_col[j] + a_row[i]) / sigma)), sigma = sqrt(stat[0]^2 + 1) sqrt(stat[1]^2 + 1), and the sum of
// the weights of a batch row goes to den[i] (the ones column of x_input).  Rows cut by a chunk boundary accumulate with
// REDs / float atomics here (the GAT path is not bit-reproducible).
struct RowsGat {
  const float* a_col;
  const float* a_row;
  const float* stat;
  float slope;
  float* den;
};

template <int NW, int SLOTS, bool GAT = false>
__global__ void __launch_bounds__(NW * 32)
    mp_fwd_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const float* __restrict__ val, const int32_t* __restrict__ chunk_row, int n_chunks, int chunk,
                       int nnz, int R, int B, const float4* __restrict__ base, uint32_t xoff4, uint32_t ldx4,
                       uint32_t toff4, uint32_t ldt4, float tail_scale, const float* __restrict__ tail_scale_dev,
                       const float* __restrict__ tail_grad, int64_t ld_tail, int C, int nslab, float info_scale, float* __restrict__ y, int64_t ldy, float* __restrict__ info,
                       double* ws_part, unsigned int* ws_count, float* __restrict__ py, RowsGat gat = RowsGat()) {
  extern __shared__ __align__(128) unsigned char rows_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RowsWarpSmem<SLOTS>& S = reinterpret_cast<RowsWarpSmem<SLOTS>*>(rows_smem)[warp];
  constexpr int G = SLOTS / 4;       // rows per cp.async group; four groups in flight
  const int64_t task = static_cast<int64_t>(blockIdx.x) * NW + warp;
  float fpart = 0.f;
  if (task < static_cast<int64_t>(n_chunks) * nslab) {
    const int slab = static_cast<int>(task / n_chunks);
    const int ch = static_cast<int>(task - static_cast<int64_t>(slab) * n_chunks);
    const int c0 = slab * 128 + lane * 4;        // the lane's four columns
    const bool active = c0 < C;
    const int eb = ch * chunk, ee = min(eb + chunk, nnz);
    const int row0 = __ldg(chunk_row + ch);
    const int rowL = ch + 1 < n_chunks ? __ldg(chunk_row + ch + 1) : R - 1;   // row of the first entry after the chunk
    const uint32_t slot_s = rows_smem_u32(S.ring) + lane * 16;                // the lane's 16 B of slot 0

    if (lane < kRowsChunkMax / 32) S.endmask[lane] = 0u;
    __syncwarp();
    // row ends inside the chunk: bit p of the mask <=> entry eb + p is the last entry of row row_of[p]
    for (int r = row0 + lane; r <= rowL; r += 32) {
      const int rs = __ldg(rowptr + r), re = __ldg(rowptr + r + 1);
      if (re > rs && re > eb && re <= ee) {
        const int p = re - 1 - eb;
        atomicOr(&S.endmask[p >> 5], 1u << (p & 31));
        S.row_of[p] = r;
      }
    }
    const bool row0_starts_here = __ldg(rowptr + row0) >= eb;
    __syncwarp();

    // entries: lane l of batch b holds entry eb + 32 b + l (value + row offset); two batches in registers
    uint32_t o_cur, o_nxt;
    float v_cur, v_nxt;
    const uint32_t lane_off = static_cast<uint32_t>(slab * 32);
    const float ts = tail_scale * (tail_scale_dev ? __ldg(tail_scale_dev) : 1.f);   // weight of the codeword rows
    float inv_sigma = 1.f, den_acc = 0.f;
    if constexpr (GAT) {
      const float ml = __ldg(gat.stat), mr = __ldg(gat.stat + 1);
      inv_sigma = 1.f / (sqrtf(ml * ml + 1.f) * sqrtf(mr * mr + 1.f));
    }
    auto load_batch = [&](int bb, uint32_t& o_l, float& v_l) {
      const int e = bb + lane;
      o_l = 0u, v_l = 0.f;
      if (e < ee) {
        const int c = __ldcs(col + e);      // streamed once: evict-first, the L2 is for the gathered rows
        v_l = __ldcs(val + e);
        if (c >= B) v_l *= ts;
        if constexpr (GAT) {
          // the row of this entry = the row whose end is the first marked position at or after it (else the row that
          // continues past the chunk)
          const int p = e - eb;
          int w = p >> 5;
          uint32_t m = S.endmask[w] >> (p & 31);
          int q = p - 1;
          if (m == 0u) {
            q = (w + 1) * 32 - 1;
            for (++w; w < kRowsChunkMax / 32 && (m = S.endmask[w]) == 0u; ++w) q += 32;
          }
          const int row = m ? S.row_of[q + __ffs(m)] : rowL;
          float ev = (__ldg(gat.a_col + c) + __ldg(gat.a_row + row)) * inv_sigma;
          ev = ev > 0.f ? ev : gat.slope * ev;
          v_l = v_l * expf(ev);
        }
        o_l = (c >= B ? toff4 + static_cast<uint32_t>(c - B) * ldt4 : xoff4 + static_cast<uint32_t>(c) * ldx4) + lane_off;
      }
    };
    // entry n of the chunk uses ring slot n % SLOTS.  refill<S0>(j0, o_src, bb): issue the G entries that sit SLOTS
    // positions after the entries j0 .. j0+G-1 of the current batch (they reuse the slots S0 .. S0+G-1 just consumed);
    // o_src holds their row offsets (the current batch's register, or the next batch's when j0 + SLOTS >= 32).  Always
    // one commit, so that the consumer's wait counts groups.
    auto refill = [&](auto s0_tag, int j0, uint32_t o_src, int bb) {
      constexpr int S0 = decltype(s0_tag)::value;
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int j = j0 + u + SLOTS;      // position relative to the current batch (may reach into the next one)
        const uint32_t o = __shfl_sync(0xffffffffu, o_src, j & 31) + lane;
        if (bb + j < ee && active) rows_cp_async16(slot_s + (S0 + u) * kRowsSlotBytes, base + o);
      }
      rows_cp_commit();
    };
    load_batch(eb, o_cur, v_cur);
    load_batch(eb + 32, o_nxt, v_nxt);
    static_assert(SLOTS == 4 * G && 32 % SLOTS == 0, "ring = four groups; a batch is a whole number of ring rounds");
    {   // prologue: the first SLOTS entries
      constexpr_for<0, 4>([&](auto gt) {
        constexpr int g = decltype(gt)::value;
        refill(IntC<g * G>{}, g * G - SLOTS, o_cur, eb);
      });
    }

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // gradient codeword row of the next out-of-batch row, prefetched (rows >= B are consecutive tail entries)
    int r_pref = row0;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_gv = [&](int r) {
      if (info && active && r >= B && r < R)
        gv = __ldcs(reinterpret_cast<const float4*>(tail_grad + static_cast<int64_t>(r - B) * ld_tail + c0));
    };
    load_gv(r_pref);

    auto flush = [&](int r, bool whole) {
      if (r != r_pref) load_gv(r);
      if constexpr (GAT) {
        if (r < B && slab == 0 && lane == 0 && gat.den) {
          if (whole) gat.den[r] = den_acc;
          else atomicAdd(gat.den + r, den_acc);
        }
        den_acc = 0.f;
      }
      if (active) {
        if (r < B) {
          float* yp = y + static_cast<int64_t>(r) * ldy + c0;
          if (whole) {
            st_vec<4>(yp, acc);
          } else if constexpr (GAT) {
            red_vec<4>(yp, acc);
          } else {
            const int kind = piece_kind(false, __ldg(rowptr + r), __ldg(rowptr + r + 1), eb, chunk);
            if (kind == kPieceRed) red_vec<4>(yp, acc);
            else st_vec<4>(py + (static_cast<int64_t>(ch) * 2 + (kind == kPieceHubStart ? 1 : 0)) * C + c0, acc);
          }
        } else if (info) {   // v2: <Y[r], Gq[r]> with Gq the node's own gradient codeword (models.py:198)
          fpart = fmaf(acc[0], gv.x, fpart), fpart = fmaf(acc[1], gv.y, fpart);
          fpart = fmaf(acc[2], gv.z, fpart), fpart = fmaf(acc[3], gv.w, fpart);
        }
      }
      r_pref = r + 1;
      load_gv(r_pref);
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    };

    // The consuming loop is unrolled over ONE ring round (SLOTS entries: static slot addresses), not over the batch:
    // with the row flush inlined 32 times the kernel was 120-150 KB of SASS and its speed followed the code size
    // (instruction cache), not the algorithm.
    int batch = 0;
    for (int bb = eb; bb < ee; bb += 32, ++batch) {
      const int cnt = min(32, ee - bb);
      const uint32_t em = S.endmask[batch];
#pragma unroll 1
      for (int j0 = 0; j0 < 32; j0 += SLOTS) {     // ring rounds of this batch
        const uint32_t emr = em >> j0;
        const uint32_t o_src = (j0 + SLOTS < 32) ? o_cur : o_nxt;
        constexpr_for<0, 4>([&](auto gt) {
          constexpr int g = decltype(gt)::value;
          rows_cp_wait<3>();       // the oldest of the four groups in flight = this round's group g
          if (j0 + g * G < cnt) {
#pragma unroll
            for (int u = 0; u < G; ++u) {
              const int s = g * G + u;              // ring slot (a constant after unrolling)
              if (j0 + s < cnt) {
                const float v = __shfl_sync(0xffffffffu, v_cur, j0 + s);
                if constexpr (GAT) den_acc += v;
                if (active) {
                  const float4 a = S.ring[s * 32 + lane];
                  acc[0] = fmaf(v, a.x, acc[0]), acc[1] = fmaf(v, a.y, acc[1]);
                  acc[2] = fmaf(v, a.z, acc[2]), acc[3] = fmaf(v, a.w, acc[3]);
                }
                if ((emr >> s) & 1u) {
                  const int r = S.row_of[batch * 32 + j0 + s];
                  flush(r, r != row0 || row0_starts_here);
                }
              }
            }
          }
          refill(IntC<g * G>{}, j0 + g * G, o_src, bb);   // a lane refills only the 16 B it has read itself
        });
      }
      o_cur = o_nxt, v_cur = v_nxt;
      load_batch(bb + 64, o_nxt, v_nxt);
    }
    rows_cp_wait<0>();
    const int pl = ee - 1 - eb;
    if (!((S.endmask[pl >> 5] >> (pl & 31)) & 1u)) flush(rowL, false);   // trailing part of a row cut by the chunk end
  }
  if (info) rows_info_reduce<NW>(static_cast<double>(fpart), ws_part, ws_count, info_scale, info);
}

}  // namespace vqgnn
