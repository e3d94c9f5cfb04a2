// info_backward of the v2 ("B+B'") formulation over the OUT-OF-BATCH rows of the batch graph
// (vq_gnn_v2/models.py:198):   info = wu * sum_{r >= B} < Y[r, :], Gq[r, :] >,   Y[r] = sum_e val[e] * Xin[col[e]]
//   = wu * sum over the entries e of rows r >= B of  val[e] * < Xin[col[e], :], Gq[r, :] >        (an SDDMM-shaped sum)
// with Xin[c] = x[c] for batch columns and the feature codewords F[c - B] of the out-of-batch node otherwise, and
// Gq[r] the gradient codewords of node r.  These rows are ~95 % of the entries of a products-shaped batch graph
// (B = 20 K batch nodes drag in ~10^6 one-hop neighbours and every edge among them) but only feed this scalar, so
// nothing is written per row: every entry contributes one partial dot product.
//
// What bounds it: one gather of C floats per entry.  Read from a row-major [T, C] table (460 MB at the products
// shape) almost every gather misses the 126 MB L2 and the kernel runs at HBM speed on 40x the compulsory bytes.
// Here the gathered tables are SLAB-MAJOR, tfS / tgS [ceil(C / SLAB)][T][SLAB], and the work is ordered slab by slab:
// while a slab is processed its slice (T * SLAB * 4 B = 58 MB at SLAB = 16) is L2-resident, so the gathers are L2
// hits and HBM only streams the edge list once per slab.  Lane = (entry slot, 4 columns of the slab); a slot owns
// 32 CONSECUTIVE entries, so it re-reads Gq only when its row changes and loads col / val as 128-bit vectors.
// Deterministic: static work assignment, per-block partials added in block order (info_reduce_ordered).
#include "mp_common.cuh"

namespace vqgnn {

constexpr int kInfoRun = 32;   // consecutive entries per slot

// L2 residency control: the edge list and the per-row gradient slices are STREAMED (each byte used once per slab) and
// together exceed the slice of the feature table that must stay resident, so they are loaded evict-first while the
// gathered feature rows are loaded evict-last.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float* p, uint64_t pol) {
  float4 v;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int4 ldg_i4_hint(const int32_t* p, uint64_t pol) {
  int4 v;
  asm("ld.global.nc.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

template <int SLAB>
__global__ void __launch_bounds__(kMpWarps * 32, 4)
    mp_info_kernel(const int32_t* __restrict__ erow, const int32_t* __restrict__ col, const float* __restrict__ val,
                   int e_begin, int nnz, int B, int T, const float* __restrict__ x, int64_t ldx,
                   const float* __restrict__ tfS, const float* __restrict__ tgS, int C, int nslab, int n_etasks,
                   float info_scale, float* __restrict__ info, double* ws_part, unsigned int* ws_count) {
  constexpr int LPE = SLAB / 4;          // lanes per entry
  constexpr int SLOTS = 32 / LPE;        // entry slots per warp
  constexpr int EPT = SLOTS * kInfoRun;  // entries per warp task
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = static_cast<int64_t>(blockIdx.x) * kMpWarps + warp;
  float fpart = 0.f;
  if (task < static_cast<int64_t>(n_etasks) * nslab) {
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const int slab = static_cast<int>(task / n_etasks);      // slab-major: concurrently running CTAs share a slab
    const int et = static_cast<int>(task - static_cast<int64_t>(slab) * n_etasks);
    const int slot = lane / LPE, cg = lane - slot * LPE;
    const int colbase = slab * SLAB + cg * 4;
    const bool col_ok = colbase < C;     // false only in the zero padding of the last slab
    const int a0 = e_begin & ~3;
    const int es = a0 + et * EPT + slot * kInfoRun;
    const float* tf = tfS + (static_cast<int64_t>(slab) * T - B) * SLAB + cg * 4;   // indexed by column id c >= B
    const float* tg = tgS + (static_cast<int64_t>(slab) * T - B) * SLAB + cg * 4;   // indexed by row id r >= B
    const float* xb = x + colbase;
#pragma unroll 2
    for (int it = 0; it < kInfoRun / 4; ++it) {
      const int e = es + it * 4;
      if (e >= nnz) break;
      int c[4], r[4];
      float v[4];
      if (e + 4 <= nnz) {
        const int4 cc = ldg_i4_hint(col + e, pol_stream), rr = ldg_i4_hint(erow + e, pol_stream);
        const float4 vv = ldg_f4_hint(val + e, pol_stream);
        c[0] = cc.x, c[1] = cc.y, c[2] = cc.z, c[3] = cc.w;
        r[0] = rr.x, r[1] = rr.y, r[2] = rr.z, r[3] = rr.w;
        v[0] = vv.x, v[1] = vv.y, v[2] = vv.z, v[3] = vv.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = e + u < nnz;
          c[u] = in ? __ldg(col + e + u) : B;
          r[u] = in ? __ldg(erow + e + u) : B;
          v[u] = in ? __ldg(val + e + u) : 0.f;
        }
      }
      float4 xin[4], gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {   // eight independent gathers in flight
        if (e + u < e_begin) c[u] = B, r[u] = B, v[u] = 0.f;      // entries of batch rows sharing the first quad
        const float* pf = c[u] >= B ? tf + static_cast<int64_t>(c[u]) * SLAB
                                    : (col_ok ? xb + static_cast<int64_t>(c[u]) * ldx : tf + static_cast<int64_t>(B) * SLAB);
        if (c[u] < B && !col_ok) v[u] = 0.f;
        xin[u] = ldg_f4_hint(pf, pol_keep);
        gv[u] = ldg_f4_hint(tg + static_cast<int64_t>(r[u]) * SLAB, pol_stream);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d = fmaf(xin[u].x, gv[u].x, fmaf(xin[u].y, gv[u].y, fmaf(xin[u].z, gv[u].z, xin[u].w * gv[u].w)));
        fpart = fmaf(v[u], d, fpart);
      }
    }
  }
  info_reduce_ordered(static_cast<double>(fpart), ws_part, ws_count, info_scale, info);
}

// erow[e] = row of CSR entry e, for the rows [r_begin, R): one warp per row
__global__ void __launch_bounds__(256)
    csr_expand_rows_kernel(const int32_t* __restrict__ rowptr, int r_begin, int R, int32_t* __restrict__ erow) {
  const int lane = threadIdx.x & 31;
  for (int r = r_begin + blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += gridDim.x * 8) {
    const int e0 = __ldg(rowptr + r), e1 = __ldg(rowptr + r + 1);
    for (int e = e0 + lane; e < e1; e += 32) erow[e] = r;
  }
}

// slab-major copies of the tail entries' codewords (D == 4, Wp == 8): warp per tail entry, lane = branch
__global__ void __launch_bounds__(256)
    tail_materialize_slab_kernel(const int32_t* __restrict__ tail_node, int64_t T, const int16_t* __restrict__ codes,
                                 const float* __restrict__ O, int nb, int M, int slab, int nslab,
                                 float* __restrict__ tfS, float* __restrict__ tgS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (t >= T) return;
  const int64_t node = tail_node ? __ldg(tail_node + t) : t;
  const int nq = nslab * slab / 4;   // 4-column groups incl. the zero padding of the last slab
  for (int k = lane; k < nq; k += 32) {
    float f[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
    if (k < nb) {
      const int code = __ldg(codes + node * nb + k);
      ld_sector(O + (static_cast<int64_t>(k) * M + code) * 8, f, g);
    }
    const int c0 = 4 * k, s = c0 / slab, o = c0 - s * slab;
    const int64_t off = (static_cast<int64_t>(s) * T + t) * slab + o;
    if (tfS) st_vec<4>(tfS + off, f);
    if (tgS) st_vec<4>(tgS + off, g);
  }
}

}  // namespace vqgnn

using namespace vqgnn;

extern "C" int vqgnn_tail_materialize_slab(const int32_t* tail_node, int64_t T, const int16_t* codes, const float* O,
                                           int nb, int M, int D, int Wp, int slab, float* tfS, float* tgS,
                                           void* stream) {
  VQ_CHECK_ARG(codes && O && T >= 0 && nb > 0 && (tfS || tgS), "tail_materialize_slab: bad arguments");
  VQ_CHECK_ARG(D == 4 && Wp == 8 && aligned32(O) && (slab == 16 || slab == 32 || slab == 64) &&
                   (!tfS || aligned16(tfS)) && (!tgS || aligned16(tgS)),
               "tail_materialize_slab: needs D == 4, Wp == 8, slab in {16, 32, 64} and 16 B aligned outputs");
  if (T == 0) return VQGNN_OK;
  const int nslab = ceil_div(nb * 4, slab);
  tail_materialize_slab_kernel<<<ceil_div(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      tail_node, T, codes, O, nb, M, slab, nslab, tfS, tgS);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" size_t vqgnn_mp_info_workspace_bytes(int64_t nnz, int C, int slab) {
  const int nslab = (C + slab - 1) / slab;
  const int64_t ept = static_cast<int64_t>(32 / (slab / 4)) * kInfoRun;
  const int64_t tasks = ((nnz + 3 + ept - 1) / ept + 1) * nslab;
  return 512 + static_cast<size_t>((tasks + kMpWarps - 1) / kMpWarps + 1) * 8;
}

extern "C" int vqgnn_csr_expand_rows(const int32_t* rowptr, int64_t r_begin, int64_t R, int32_t* erow, void* stream) {
  VQ_CHECK_ARG(rowptr && erow && r_begin >= 0 && R >= r_begin && R < (1ll << 31), "csr_expand_rows: bad arguments");
  if (R == r_begin) return VQGNN_OK;
  const int grid = static_cast<int>(std::min<int64_t>((R - r_begin + 7) / 8, 32 * kNumSMs));
  csr_expand_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rowptr, (int)r_begin, (int)R, erow);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_mp_info(const int32_t* erow, const int32_t* col, const float* val, int64_t e_begin, int64_t nnz,
                             int64_t B, int64_t R, const float* x, int64_t ldx, const float* tfS, const float* tgS,
                             int C, int slab, float info_scale, float* info, void* ws, void* stream) {
  VQ_CHECK_ARG(erow && col && val && x && tfS && tgS && info && ws, "mp_info: null argument");
  VQ_CHECK_ARG(B > 0 && R >= B && R < (1ll << 31) && e_begin >= 0 && e_begin <= nnz && nnz < (1ll << 31) - 64 && C > 0,
               "mp_info: bad sizes");
  VQ_CHECK_ARG((slab == 16 || slab == 32 || slab == 64) && ldx % 4 == 0 && C % 4 == 0 && aligned16(x) &&
                   aligned16(tfS) && aligned16(tgS) && aligned16(col) && aligned16(val) && aligned16(erow),
               "mp_info: slab must be 16 / 32 / 64 and every operand 16 B aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (R == B || e_begin == nnz) {
    VQ_CUDA(cudaMemsetAsync(info, 0, sizeof(float), s));
    return VQGNN_OK;
  }
  const int nslab = ceil_div(C, slab);
  const int64_t ept = static_cast<int64_t>(32 / (slab / 4)) * kInfoRun;
  const int64_t a0 = e_begin & ~static_cast<int64_t>(3);
  const int64_t n_etasks = (nnz - a0 + ept - 1) / ept;
  const int64_t tasks = n_etasks * nslab;
  const int grid = static_cast<int>((tasks + kMpWarps - 1) / kMpWarps);
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  unsigned int* ws_count = reinterpret_cast<unsigned int*>(p);
  double* ws_part = reinterpret_cast<double*>(p + 256);
  VQ_CUDA(cudaMemsetAsync(ws_count, 0, 16, s));
  const int T = static_cast<int>(R - B);
#define VQ_INFO(SL)                                                                                               \
  mp_info_kernel<SL><<<grid, kMpWarps * 32, 0, s>>>(erow, col, val, (int)e_begin, (int)nnz, (int)B, T, x, ldx, tfS, tgS, \
                                                    C, nslab, (int)n_etasks, info_scale, info, ws_part, ws_count)
  if (slab == 16) VQ_INFO(16);
  else if (slab == 32) VQ_INFO(32);
  else VQ_INFO(64);
#undef VQ_INFO
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
