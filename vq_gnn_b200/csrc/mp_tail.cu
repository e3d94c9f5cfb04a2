// Out-of-batch ("tail") message passing with the codebooks of a branch GROUP resident in shared memory.
// (round 2: lane = (entry slot, branch) instead of lane = entry -- 8 accumulators per lane, 32 warps per SM.)
//
// The generic kernel (mp.cu) fetches one 32 B codeword sector per (entry, branch) through L1tex/L2 and is bound
// by the L1tex wavefront rate (profiles/r1_mp_fwd_ncu_full.txt: 484 M sectors, 1.2 sectors/clk/SM).  For the v1
// formulation (vq_gnn_v1/utils/dataloader.py:144-192 `mapper` + vq_gnn_v1/models.py:170-223), where every batch
// row has hundreds of out-of-batch neighbours, this kernel instead
//   * keeps one HALF (feature or gradient columns) of the de-whitened codebooks O_k of G = 8 consecutive branches
//     in shared memory (M*128 B <= 192 KB), laid out so that the gathers are bank-conflict free,
//   * reads the G codes of a neighbour with ONE 16 B load from a group-major copy of the code table
//     (codes_g [ceil(nb/G)][N][8] int16), i.e. one global sector per (entry, group) instead of G,
//   * gathers codewords with LDS.128 (16 B chunk index XOR-swizzled so random codes spread over all banks),
//   * maps lane = (entry slot, branch): a warp advances 32 / G entries per sub-step, every lane gathers ITS branch's
//     codeword of its slot's entry and keeps 8 partial sums; a row ends with 32 / G shuffles per column into the
//     slot-0 lanes, which store / RED 16 B vectors of y and gq.
// Work items = (branch group, block of 512-entry chunks), dealt round-robin to one persistent CTA per SM.
//   yt[r, k*4..] = sum_e val[e] * feat_scale * O_k[code_k(node[e]), :4]
//   gq[r, k*4..] = sum_e rval[e]            * O_k[code_k(node[e]), 4:8]
// then (mp_tail_finish_kernel)  y += yt ;  info = info_scale * sum_r <x[r], gq[r]>.
// Deterministic: a row segment wholly inside a chunk is stored, a row cut once is two REDs onto zero (they commute),
// the pieces of longer rows go through the piece buffer + mp_tail_fixup_kernel in chunk order; the info scalar is an
// ordered two-level reduction.  Which warp runs which chunk (dynamic scheduling) therefore cannot change a bit.
#include "mp_common.cuh"

namespace vqgnn {

constexpr int kTailWarps = 32;
constexpr int kTailThreads = kTailWarps * 32;

__device__ __forceinline__ void red_v4(float* p, const float (&v)[4]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
               : "memory");
}

// Lane mapping: G = 8 branches per group; a warp advances 8 entries per sub-step, lane = slot * 4 + q gathers the
// codeword HALVES (4 floats, one LDS.128 each) of branches 2q and 2q+1 of the entry in `slot` -- odd slots in the
// opposite order, so the eight lanes of every LDS.128 phase (two entries x four lanes) still cover the eight bank
// groups.  The two halves are separate passes ("half groups"): features -> y with val * feat_scale, gradients -> gq
// with rval.
// Shared-memory layout of a half group: chunk (code, g) at byte code * 128 + g * 16, i.e. the eight lanes of an
// LDS.128 phase (one entry's eight branches) always hit eight DIFFERENT 16 B bank groups whatever the codes are:
// the random gathers are bank-conflict free by construction (M * 128 B <= 192 KB: M <= 1536).
constexpr int kTailG = 8;
constexpr int kTailEPS = 8;     // entries per sub-step
constexpr int kTailBatch = 32;  // entries staged per batch (one per lane)

__global__ void __launch_bounds__(kTailThreads, 1)
    mp_tail_smem_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ node,
                        const float* __restrict__ val, const float* __restrict__ rval,
                        const int32_t* __restrict__ chunk_row, int n_chunks, int chunk, int nnz,
                        const int32_t* __restrict__ d_nnz, int B, const int16_t* __restrict__ codes_g, int64_t N, const float* __restrict__ O, int nb, int M,
                        float feat_scale, float* __restrict__ y, int64_t ldy, float* __restrict__ gq, int64_t ldgq,
                        float* __restrict__ py, float* __restrict__ pgq, int C, int ng, int cpi,
                        int items_per_group) {
  constexpr int G = kTailG, EPS = kTailEPS, BATCH = kTailBatch;
  extern __shared__ __align__(128) unsigned char cb_smem[];  // [M][8] chunks of 16 B
  __shared__ uint4 stage[kTailWarps][32];                    // per warp: the batch's code vectors (8 int16 each)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* cb_ptr = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cb_smem) + 127) & ~uintptr_t(127));
  const uint32_t cb_base = static_cast<uint32_t>(__cvta_generic_to_shared(cb_ptr));
  const uint32_t st_base = static_cast<uint32_t>(__cvta_generic_to_shared(&stage[warp][0]));
  const int slot = lane >> 2, q = lane & 3;
  const int gA = 2 * q + (slot & 1), gB = 2 * q + 1 - (slot & 1);   // branch gathered first / second
  __shared__ int next_chunk;  // warps of the CTA draw the item's chunks dynamically
  if (d_nnz) {  // entry count only known on the device (vqgnn_plan_v1_build): the host sized the grid by an upper bound
    nnz = __ldg(d_nnz);
    n_chunks = (nnz + chunk - 1) / chunk;
  }
  int loaded = -1;
  const int nhg = 2 * ng;   // half groups: (group, half)
  const int n_items = nhg * items_per_group;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int hg = item / items_per_group;
    const int cblk = item - hg * items_per_group;
    const int gk = hg >> 1, half = hg & 1;
    if (half == 1 && gq == nullptr) continue;   // uniform over the CTA
    const int kbase = gk * G;
    const int gcount = min(G, nb - kbase);
    if (hg != loaded) {  // (re)load the half group's codewords into the conflict-free layout
      __syncthreads();
      const int n16 = gcount * M;
      for (int c = threadIdx.x; c < n16; c += kTailThreads) {
        const int gI = c / M, m = c - gI * M;
        const float4 v = __ldg(reinterpret_cast<const float4*>(O + (static_cast<int64_t>(kbase + gI) * M + m) * 8 + half * 4));
        *reinterpret_cast<float4*>(cb_ptr + (static_cast<size_t>(m) * 8 + gI) * 16) = v;
      }
      loaded = hg;
      __syncthreads();
    }
    const int16_t* cg = codes_g + static_cast<int64_t>(gk) * N * 8;
    const float* wsrc = half ? rval : val;
    const float wscale = half ? 1.f : feat_scale;
    float* out = half ? gq : y;
    const int64_t ldo = half ? ldgq : ldy;
    float* pout = half ? pgq : py;
    const int c_begin = cblk * cpi, c_end = min(c_begin + cpi, n_chunks);
    __syncthreads();  // every warp is done with the previous item (and its chunk counter)
    if (threadIdx.x == 0) next_chunk = c_begin;
    __syncthreads();
    const bool onA = gA < gcount, onB = gB < gcount;
    const uint32_t cbA = cb_base + gA * 16, cbB = cb_base + gB * 16;
    const uint32_t my_st = st_base + slot * 16 + q * 4;    // the u32 holding the codes of branches 2q (low) / 2q+1
    const int shA = (slot & 1) * 16, shB = 16 - shA;        // which half of that word is branch gA / gB
    const int colbase = (kbase + 2 * q) * 4;                 // 8 contiguous columns: branches 2q, 2q+1

    while (true) {
      int ch = 0;
      if (lane == 0) ch = atomicAdd(&next_chunk, 1);
      ch = __shfl_sync(0xffffffffu, ch, 0);
      if (ch >= c_end) break;
      const int eb = ch * chunk, ee = min(eb + chunk, nnz);
      float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};
      // rows: lane i holds rowptr[rbase + i]
      int r = __ldg(chunk_row + ch);
      int rbase = r;
      int rp_l = __ldg(rowptr + min(rbase + lane, B));
      int rs = __shfl_sync(0xffffffffu, rp_l, 0), re = __shfl_sync(0xffffffffu, rp_l, 1);

      // the current row segment is complete: sum the 8 slots of every column into the slot-0 lanes and emit
      auto flush = [&](bool whole) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // t[0..4) = branch 2q, t[4..8) = branch 2q+1 (odd slots hold them swapped)
          t[i] = (slot & 1) ? accB[i] : accA[i];
          t[4 + i] = (slot & 1) ? accA[i] : accB[i];
          accA[i] = 0.f, accB[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i] += __shfl_xor_sync(0xffffffffu, t[i], 4);
          t[i] += __shfl_xor_sync(0xffffffffu, t[i], 8);
          t[i] += __shfl_xor_sync(0xffffffffu, t[i], 16);
        }
        if (slot == 0 && 2 * q < gcount) {
          const int kind = piece_kind(whole, rs, re, eb, chunk);
          const bool two = 2 * q + 1 < gcount;
          float* dst = out + static_cast<int64_t>(r) * ldo + colbase;
          float* pdst = pout + (static_cast<int64_t>(ch) * 2 + (kind == kPieceHubStart ? 1 : 0)) * C + colbase;
          const float lo[4] = {t[0], t[1], t[2], t[3]}, hi[4] = {t[4], t[5], t[6], t[7]};
          if (kind == kPieceWhole) {
            *reinterpret_cast<float4*>(dst) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            if (two) *reinterpret_cast<float4*>(dst + 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          } else if (kind == kPieceRed) {
            red_v4(dst, lo);
            if (two) red_v4(dst + 4, hi);
          } else {
            *reinterpret_cast<float4*>(pdst) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            if (two) *reinterpret_cast<float4*>(pdst + 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          }
        }
      };
      auto next_row = [&](int upto) {   // advance to the row that holds entry `upto`
        do {
          ++r;
          if (r - rbase >= 31) {
            rbase = r;
            rp_l = __ldg(rowptr + min(rbase + lane, B));
          }
          rs = __shfl_sync(0xffffffffu, rp_l, r - rbase);
          re = __shfl_sync(0xffffffffu, rp_l, r - rbase + 1);
        } while (re <= upto);
      };

      // software pipeline: (node, weight) two batches ahead, code vectors one batch ahead
      int node_n = 0;
      float v_n = 0.f;
      uint4 cv_n = make_uint4(0, 0, 0, 0);
      {
        const int e = eb + lane;
        if (e < ee) node_n = __ldg(node + e), v_n = __ldg(wsrc + e);
        cv_n = __ldg(reinterpret_cast<const uint4*>(cg + static_cast<int64_t>(node_n) * 8));
      }
      int node_nn = 0;
      float v_nn = 0.f;
      {
        const int e = eb + BATCH + lane;
        if (e < ee) node_nn = __ldg(node + e), v_nn = __ldg(wsrc + e);
      }
      bool pending = false;

      for (int bb = eb; bb < ee; bb += BATCH) {
        const float v_l = v_n * wscale;
        __syncwarp();
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(st_base + lane * 16), "r"(cv_n.x), "r"(cv_n.y),
                     "r"(cv_n.z), "r"(cv_n.w)
                     : "memory");
        __syncwarp();
        // advance the pipeline
        node_n = node_nn, v_n = v_nn;
        cv_n = __ldg(reinterpret_cast<const uint4*>(cg + static_cast<int64_t>(node_n) * 8));
        {
          const int e = bb + 2 * BATCH + lane;
          node_nn = 0, v_nn = 0.f;
          if (e < ee) node_nn = __ldg(node + e), v_nn = __ldg(wsrc + e);
        }
        const int bend = min(bb + BATCH, ee);
        if (bend - bb == BATCH && bend <= re) {
          // fast path (the usual case: rows are hundreds of entries long): the whole batch belongs to the current row,
          // no row bookkeeping inside -- per sub-step 1 SHFL + 1 LDS.32 + 2 conflict-free LDS.128 + 8 FFMA per lane
          uint32_t code[BATCH / EPS];
#pragma unroll
          for (int j = 0; j < BATCH / EPS; ++j)
            asm("ld.shared.u32 %0, [%1];" : "=r"(code[j]) : "r"(my_st + j * EPS * 16));
#pragma unroll
          for (int j = 0; j < BATCH / EPS; ++j) {
            const float vf = __shfl_sync(0xffffffffu, v_l, j * EPS + slot);
            const uint32_t ca = (code[j] >> shA) & 0xffffu, cb2 = (code[j] >> shB) & 0xffffu;
            if (onA) {
              float4 f;
              asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                  : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w)
                  : "r"(cbA + ca * 128));
              accA[0] = fmaf(vf, f.x, accA[0]), accA[1] = fmaf(vf, f.y, accA[1]), accA[2] = fmaf(vf, f.z, accA[2]), accA[3] = fmaf(vf, f.w, accA[3]);
            }
            if (onB) {
              float4 f;
              asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                  : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w)
                  : "r"(cbB + cb2 * 128));
              accB[0] = fmaf(vf, f.x, accB[0]), accB[1] = fmaf(vf, f.y, accB[1]), accB[2] = fmaf(vf, f.z, accB[2]), accB[3] = fmaf(vf, f.w, accB[3]);
            }
          }
          pending = true;
          if (bend == re) {
            flush(rs >= eb);
            pending = false;
            if (bend < ee) next_row(bend);
          }
          continue;
        }
        for (int q0 = bb; q0 < bend; q0 += EPS) {   // general path: the batch holds a row boundary (or is ragged)
          const int src = q0 - bb + slot;
          const int e = q0 + slot;
          const float vf0 = __shfl_sync(0xffffffffu, v_l, src);
          float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), fb = fa;
          const bool live = e < bend;
          if (live) {
            uint32_t code;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(st_base + src * 16 + q * 4));
            const uint32_t ca = (code >> shA) & 0xffffu, cb2 = (code >> shB) & 0xffffu;
            if (onA)
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(fa.x), "=f"(fa.y), "=f"(fa.z), "=f"(fa.w)
                           : "r"(cbA + ca * 128));
            if (onB)
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(fb.x), "=f"(fb.y), "=f"(fb.z), "=f"(fb.w)
                           : "r"(cbB + cb2 * 128));
          }
          const int qend = min(q0 + EPS, bend);
          int j0 = q0;
          while (j0 < qend) {   // pieces of this sub-step that belong to one row
            const int pend = min(re, qend);
            const bool on = live && e >= j0 && e < pend;
            const float vf = on ? vf0 : 0.f;
            accA[0] = fmaf(vf, fa.x, accA[0]), accA[1] = fmaf(vf, fa.y, accA[1]), accA[2] = fmaf(vf, fa.z, accA[2]), accA[3] = fmaf(vf, fa.w, accA[3]);
            accB[0] = fmaf(vf, fb.x, accB[0]), accB[1] = fmaf(vf, fb.y, accB[1]), accB[2] = fmaf(vf, fb.z, accB[2]), accB[3] = fmaf(vf, fb.w, accB[3]);
            pending = true;
            j0 = pend;
            if (pend == re) {  // row r complete
              flush(rs >= eb);
              pending = false;
              if (j0 >= ee) break;
              next_row(j0);
            }
          }
        }
      }
      if (pending) flush(false);
    }
  }
}

// pieces of rows spanning >= 3 chunks, summed in chunk order: thread = (chunk c, float2 column pair)
__global__ void __launch_bounds__(256)
    mp_tail_fixup_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ chunk_row, int n_chunks,
                         int chunk, const int32_t* __restrict__ d_nnz, int B, int C, const float* __restrict__ py,
                         float* __restrict__ y, int64_t ldy, const float* __restrict__ pgq, float* __restrict__ gq,
                         int64_t ldgq) {
  if (d_nnz) n_chunks = (__ldg(d_nnz) + chunk - 1) / chunk;
  const int half = C / 2;
  const int64_t total = static_cast<int64_t>(n_chunks) * half;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i / half);
    const int col = static_cast<int>(i - static_cast<int64_t>(c) * half) * 2;
    if (c + 2 >= n_chunks) continue;
    const int r = __ldg(chunk_row + c + 1);
    if (r >= B) continue;
    const int64_t rs = __ldg(rowptr + r), re = __ldg(rowptr + r + 1);
    if (rs < static_cast<int64_t>(c) * chunk || rs >= static_cast<int64_t>(c + 1) * chunk) continue;
    const int lc = static_cast<int>((re - 1) / chunk);
    if (lc < c + 2) continue;
    float2 a = *reinterpret_cast<const float2*>(py + (static_cast<int64_t>(c) * 2 + 1) * C + col);
    float2 q = make_float2(0.f, 0.f);
    if (gq) q = *reinterpret_cast<const float2*>(pgq + (static_cast<int64_t>(c) * 2 + 1) * C + col);
    for (int cc = c + 1; cc <= lc; ++cc) {
      const float2 t = *reinterpret_cast<const float2*>(py + static_cast<int64_t>(cc) * 2 * C + col);
      a.x += t.x, a.y += t.y;
      if (gq) {
        const float2 u = *reinterpret_cast<const float2*>(pgq + static_cast<int64_t>(cc) * 2 * C + col);
        q.x += u.x, q.y += u.y;
      }
    }
    *reinterpret_cast<float2*>(y + static_cast<int64_t>(r) * ldy + col) = a;
    if (gq) *reinterpret_cast<float2*>(gq + static_cast<int64_t>(r) * ldgq + col) = q;
  }
}

// y += yt ; info = info_scale * sum <x, gq>  (ordered two-level reduction)
__global__ void __launch_bounds__(256)
    mp_tail_finish_kernel(int64_t B, int C, const float* __restrict__ yt, float* __restrict__ y, int64_t ldy,
                          const float* __restrict__ x, int64_t ldx, const float* __restrict__ gq, int64_t ldgq,
                          float* __restrict__ info, float info_scale, double* ws_part, unsigned int* ws_count) {
  const int half = C / 2;
  const int64_t total = B * half;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  double part = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / half;
    const int col = static_cast<int>(i - r * half) * 2;
    float2* yp = reinterpret_cast<float2*>(y + r * ldy + col);
    const float2 t = *reinterpret_cast<const float2*>(yt + r * C + col);
    float2 v = *yp;
    v.x += t.x, v.y += t.y;
    *yp = v;
    if (info) {
      const float2 xr = __ldg(reinterpret_cast<const float2*>(x + r * ldx + col));
      const float2 g = *reinterpret_cast<const float2*>(gq + r * ldgq + col);
      part += static_cast<double>(fmaf(xr.x, g.x, xr.y * g.y));
    }
  }
  if (info) info_reduce_ordered(part, ws_part, ws_count, info_scale, info);
}

// codes_g[(k / G) * N + node][k % G] = codes[node, k] for the listed nodes (all N when rows == nullptr)
__global__ void codes_group_kernel(const int16_t* __restrict__ codes, int nb, const int32_t* __restrict__ rows,
                                   int64_t n_rows, int64_t N, int G, int16_t* __restrict__ codes_g) {
  const int64_t total = n_rows * nb;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t rI = i / nb;
    const int k = static_cast<int>(i - rI * nb);
    const int64_t nd = rows ? __ldg(rows + rI) : rI;
    codes_g[((static_cast<int64_t>(k / G)) * N + nd) * 8 + (k % G)] = __ldg(codes + nd * nb + k);
  }
}

// Apply a list of (node, codes) updates to the code table with LAST-ENTRY-WINS semantics for repeated nodes
// (multi-GPU: the list is the all-gather of every rank's re-assignments in rank order, so every replica resolves
// a node shared by two ranks' batches identically).  owner[node] = index of the last entry naming the node.
__global__ void codes_owner_kernel(const int32_t* __restrict__ nodes, int64_t n, int32_t* __restrict__ owner,
                                   int phase) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride) {
    const int32_t nd = __ldg(nodes + e);
    if (nd < 0) continue;  // padding of a short batch (dist.allgather_code_updates capacity)
    if (phase == 0) owner[nd] = -1;
    else atomicMax(owner + nd, static_cast<int32_t>(e));
  }
}
__global__ void codes_apply_kernel(const int32_t* __restrict__ nodes, const int16_t* __restrict__ new_codes,
                                   int64_t n, int nbc, int k0, const int32_t* __restrict__ owner,
                                   int16_t* __restrict__ codes, int nb, int64_t N, int16_t* __restrict__ codes_g,
                                   int G) {
  const int64_t total = n * nbc;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i / nbc;
    const int j = static_cast<int>(i - e * nbc);
    const int64_t nd = __ldg(nodes + e);
    if (nd < 0 || __ldg(owner + nd) != e) continue;
    const int k = k0 + j;
    const int16_t cv = __ldg(new_codes + i);
    codes[nd * nb + k] = cv;
    if (codes_g) codes_g[((static_cast<int64_t>(k / G)) * N + nd) * 8 + (k % G)] = cv;
  }
}

}  // namespace vqgnn

using namespace vqgnn;

extern "C" int vqgnn_codes_apply_updates(const int32_t* nodes, const int16_t* new_codes, int64_t n, int nbc, int k0,
                                         int16_t* codes, int nb, int64_t N, int16_t* codes_g, int G,
                                         int32_t* owner_ws, void* stream) {
  VQ_CHECK_ARG(nodes && new_codes && codes && owner_ws && n >= 0 && n < (1ll << 31) && nbc > 0 && k0 >= 0 &&
                   k0 + nbc <= nb && N > 0,
               "codes_apply_updates: bad arguments");
  VQ_CHECK_ARG(!codes_g || (G >= 1 && G <= 8), "codes_apply_updates: bad group size");
  if (n == 0) return VQGNN_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int g1 = static_cast<int>(std::min<int64_t>((n + 255) / 256, 8 * kNumSMs));
  codes_owner_kernel<<<g1, 256, 0, s>>>(nodes, n, owner_ws, 0);
  VQ_LAUNCH_CHECK();
  codes_owner_kernel<<<g1, 256, 0, s>>>(nodes, n, owner_ws, 1);
  VQ_LAUNCH_CHECK();
  const int g2 = static_cast<int>(std::min<int64_t>((n * nbc + 255) / 256, 16 * kNumSMs));
  codes_apply_kernel<<<g2, 256, 0, s>>>(nodes, new_codes, n, nbc, k0, owner_ws, codes, nb, N, codes_g, G);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_mp_tail_group(int M, int D, int Wp) {
  if (D != 4 || Wp != 8 || M <= 0) return 0;
  if (M * 128 <= 196608) return kTailG;   // one HALF of 8 branches' codebooks per pass: M * 8 * 16 B
  return 0;
}

extern "C" int vqgnn_codes_group(const int16_t* codes, int nb, const int32_t* rows, int64_t n_rows, int64_t N,
                                 int G, int16_t* codes_g, void* stream) {
  VQ_CHECK_ARG(codes && codes_g && nb > 0 && n_rows >= 0 && N > 0 && G >= 1 && G <= 8, "codes_group: bad arguments");
  if (n_rows == 0) return VQGNN_OK;
  const int grid = static_cast<int>(std::min<int64_t>((n_rows * nb + 255) / 256, 16 * kNumSMs));
  codes_group_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, nb, rows, n_rows, N, G, codes_g);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

static inline size_t tl_al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
constexpr int kTailFinishGrid = 4 * kNumSMs;

// workspace: [count 256 B][finish-kernel partials][yt B*C][y pieces n_chunks*2*C][gq pieces n_chunks*2*C]
extern "C" size_t vqgnn_mp_fwd_tail_workspace_bytes(int64_t nnz, int chunk, int64_t B, int C) {
  const int64_t n_chunks = chunk > 0 ? (nnz + chunk - 1) / chunk : 0;
  return 512 + tl_al256(static_cast<size_t>(kTailFinishGrid) * 8) + tl_al256(static_cast<size_t>(B) * C * 4) +
         2 * tl_al256(static_cast<size_t>(n_chunks) * 2 * C * 4);
}

extern "C" int vqgnn_mp_fwd_tail(const int32_t* rowptr, const int32_t* node, const float* val, const float* rval,
                                 const int32_t* chunk_row, int chunk, int64_t nnz, const int32_t* d_nnz, int64_t B,
                                 const float* x, int64_t ldx, const int16_t* codes_g, int64_t N, const float* O, int nb, int M,
                                 int D, int Wp, float feat_scale, float info_scale, float* y, int64_t ldy,
                                 float* gq, int64_t ldgq, float* info, void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && node && val && rval && x && codes_g && O && y, "mp_fwd_tail: null argument");
  const int G = vqgnn_mp_tail_group(M, D, Wp);
  VQ_CHECK_ARG(G != 0, "mp_fwd_tail: needs D == 4, Wp == 8 and M <= 1024 (got M=%d D=%d Wp=%d)", M, D, Wp);
  VQ_CHECK_ARG(B > 0 && B < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31) && nb > 0, "mp_fwd_tail: bad sizes");
  VQ_CHECK_ARG(chunk > 0 && chunk % 32 == 0 && (nnz == 0 || chunk_row), "mp_fwd_tail: needs chunk_row");
  VQ_CHECK_ARG(ldy % 2 == 0 && ldx % 2 == 0 && (!gq || ldgq % 4 == 0) && aligned16(O) && aligned16(codes_g) &&
                   (reinterpret_cast<uintptr_t>(y) & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0 &&
                   (!gq || aligned16(gq)),
               "mp_fwd_tail: y / x must be 8 B aligned with even leading dimensions, gq 16 B aligned with ldgq % 4 == 0");
  VQ_CHECK_ARG(ws, "mp_fwd_tail: needs a workspace of vqgnn_mp_fwd_tail_workspace_bytes() bytes");
  VQ_CHECK_ARG(!info || gq, "mp_fwd_tail: info needs gq");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n_chunks = static_cast<int>((nnz + chunk - 1) / chunk);
  const int C = nb * D;
  if (n_chunks == 0) {
    if (gq)
      if (int rc = zero_rows(gq, B, C, ldgq, s)) return rc;
    if (info) VQ_CUDA(cudaMemsetAsync(info, 0, sizeof(float), s));
    return VQGNN_OK;
  }
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  unsigned int* ws_count = reinterpret_cast<unsigned int*>(p);
  double* ws_part = reinterpret_cast<double*>(p + 256);
  float* yt = reinterpret_cast<float*>(p + 256 + tl_al256(static_cast<size_t>(kTailFinishGrid) * 8));
  float* py = yt + tl_al256(static_cast<size_t>(B) * C * 4) / 4;
  float* pgq = py + tl_al256(static_cast<size_t>(n_chunks) * 2 * C * 4) / 4;
  VQ_CUDA(cudaMemsetAsync(ws_count, 0, 16, s));
  VQ_CUDA(cudaMemsetAsync(yt, 0, sizeof(float) * B * C, s));
  if (gq)
    if (int rc = zero_rows(gq, B, C, ldgq, s)) return rc;
  const int ng = (nb + G - 1) / G;
  // items = (group, block of cpi chunks): ~8 items per CTA, dealt round-robin
  const int grid = kNumSMs;
  int items_per_group = std::max(1, (grid * 8 + 2 * ng - 1) / (2 * ng));   // per HALF group
  int cpi = std::max(kTailWarps, (n_chunks + items_per_group - 1) / items_per_group);
  items_per_group = (n_chunks + cpi - 1) / cpi;
  const size_t smem = static_cast<size_t>(M) * 128 + 128;
  VQ_CUDA(cudaFuncSetAttribute(mp_tail_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mp_tail_smem_kernel<<<std::min(grid, 2 * ng * items_per_group), kTailThreads, smem, s>>>(
      rowptr, node, val, rval, chunk_row, n_chunks, chunk, (int)nnz, d_nnz, (int)B, codes_g, N, O, nb, M, feat_scale, yt,
      (int64_t)C, gq, ldgq, py, pgq, C, ng, cpi, items_per_group);
  VQ_LAUNCH_CHECK();
  const int64_t fx = static_cast<int64_t>(n_chunks) * (C / 2);
  mp_tail_fixup_kernel<<<static_cast<int>(std::min<int64_t>((fx + 255) / 256, 8 * kNumSMs)), 256, 0, s>>>(
      rowptr, chunk_row, n_chunks, chunk, d_nnz, (int)B, C, py, yt, C, pgq, gq, ldgq);
  VQ_LAUNCH_CHECK();
  const int64_t tot = B * (C / 2);
  const int fgrid = static_cast<int>(std::min<int64_t>((tot + 255) / 256, kTailFinishGrid));
  mp_tail_finish_kernel<<<fgrid, 256, 0, s>>>(B, C, yt, y, ldy, x, ldx, gq, ldgq, info, info_scale, ws_part, ws_count);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
