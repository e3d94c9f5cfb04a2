// VQ-approximated GAT message passing, v1 ("B+M") formulation: vq_gnn_v1/models.py:143-233 with OurGATConv
// (vq_gnn_v1/convs.py, same maths as vq_gnn_v2/convs.py:165-266) on the per-branch (B+M)^2 graph that
// vq_gnn_v1/utils/dataloader.py:144-192 (`mapper`) builds.  Every branch k is an independent little GAT with
// D+1 = 5 columns ([features | 1]) and its own att_l / att_r:
//   nodes        : B batch rows (features x[:, 4k:4k+4]) and M codewords (features wu * O_k[m, :4])
//   scores       : a_l/a_r [B, nb] for batch rows, c_l/c_r [nb, M] for codewords; sigma_k from the maxima over both
//   row i < B    : Y_k[i] = sum_{e tail} val_e E((c_l[code_e] + a_r[i]) / sigma) [wu O_k[code_e,:4] | 1]
//                         + sum_{e in-batch i'} val_e E((a_l[i'] + a_r[i]) / sigma) [x_k[i'] | 1]     E = exp o lrelu
//                  out_k[i] = Y_k[i,:4] / (Y_k[i,4] + 1e-16)
//   rows B + m   : only feed info_backward = wu * sum_m <Y_k[B+m, :5], O_k[m, 4:9]>, evaluated per edge:
//                  wu * sum_{e tail of row i} rval_e E((a_l[i] + c_r[code_e]) / sigma) (<x_k[i], O_k[code_e,4:8]> + O_k[code_e,8])
// The count matrices A*R of `mapper` are never formed: summing per edge over code_e equals summing over m with
// the coalesced counts.  One lane = one branch (D = 4); same nnz-balanced warp tasks as mp.cu.
#include <math.h>

#include "mp_common.cuh"

namespace vqgnn {

constexpr int kG1Wp = 12;  // codeword row: 4 feature | 4 gradient | gradient of the ones column | 3 pad

__device__ __forceinline__ float g1_inv_sigma(const float* __restrict__ stat, int k) {
  const float ml = __ldg(stat + 2 * k), mr = __ldg(stat + 2 * k + 1);
  return 1.f / (sqrtf(ml * ml + 1.f) * sqrtf(mr * mr + 1.f));
}
__device__ __forceinline__ float lrelu(float e, float slope) { return e > 0.f ? e : slope * e; }
__device__ __forceinline__ float dot4(const float (&a)[4], const float (&b)[4]) {
  return fmaf(a[0], b[0], fmaf(a[1], b[1], fmaf(a[2], b[2], a[3] * b[3])));
}

// (F1) scores of batch rows [B, nb] and codewords cs [nb, M, 2] + per-branch maxima stat [nb, 2]
__global__ void gat1_stat_init_kernel(float* stat, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stat[i] = __int_as_float(0xff800000);
}
__global__ void __launch_bounds__(256)
    gat1_scores_kernel(int64_t B, int nb, int M, const float* __restrict__ x, int64_t ldx,
                       const float* __restrict__ O, float wu, const float* __restrict__ att_l,
                       const float* __restrict__ att_r, float* __restrict__ a_l, float* __restrict__ a_r,
                       float* __restrict__ cs, float* __restrict__ stat) {
  // thread -> (item, branch); items [0, B) are batch rows, [B, B + M) codewords
  const int64_t total = (B + M) * nb;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t it = t / nb;
    const int k = static_cast<int>(t - it * nb);
    float v[4], wl[4], wr[4];
    float s = 1.f;
    if (it < B) {
      ld_vec<4>(x + it * ldx + 4 * k, v);
    } else {
      ld_vec<4>(O + (static_cast<int64_t>(k) * M + (it - B)) * kG1Wp, v);
      s = wu;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) wl[i] = __ldg(att_l + 5 * k + i), wr[i] = __ldg(att_r + 5 * k + i);
    const float sl = fmaf(s, dot4(v, wl), __ldg(att_l + 5 * k + 4));
    const float sr = fmaf(s, dot4(v, wr), __ldg(att_r + 5 * k + 4));
    if (it < B) {
      a_l[it * nb + k] = sl, a_r[it * nb + k] = sr;
    } else {
      float* p = cs + (static_cast<int64_t>(k) * M + (it - B)) * 2;
      p[0] = sl, p[1] = sr;
    }
    atomic_max_float(stat + 2 * k, sl);
    atomic_max_float(stat + 2 * k + 1, sr);
  }
}

// per-row state of a lane (= branch) while walking a CSR chunk
struct G1Row : PlainWeights {
  const float* a_l;
  const float* a_r;
  const float* x;
  int64_t ldx;
  int nb, k;
  bool active;
  float al = 0.f, ar = 0.f;
  float xr[4];
  __device__ __forceinline__ void row_begin(int r) {
    if (!active) return;
    al = __ldg(a_l + static_cast<int64_t>(r) * nb + k);
    ar = __ldg(a_r + static_cast<int64_t>(r) * nb + k);
    ld_vec<4>(x + static_cast<int64_t>(r) * ldx + 4 * k, xr);
  }
};

// (F2) aggregation: un-normalised Y [B, C], den [B, nb], info
template <bool HAS_RV>
__global__ void __launch_bounds__(kMpWarps * 32)
    gat1_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const float* __restrict__ val, const float* __restrict__ rval,
                    const int32_t* __restrict__ chunk_row, int n_chunks, int chunk, int nnz, int B,
                    const float* __restrict__ x, int64_t ldx, const int16_t* __restrict__ codes,
                    const float* __restrict__ O, int nb, int M, float wu, const float* __restrict__ a_l,
                    const float* __restrict__ a_r, const float* __restrict__ cs, const float* __restrict__ stat,
                    float slope, int nslab, float* __restrict__ y, int64_t ldy, float* __restrict__ den,
                    float* __restrict__ info, float info_scale, double* ws_sum, unsigned int* ws_count) {
  constexpr int U = kMpUnroll;
  const int lane = threadIdx.x & 31;
  const MpTask t = mp_task<4>(chunk_row, n_chunks, chunk, nnz, nslab, nb * 4, 4);
  float fpart = 0.f;
  if (t.valid) {
    const int k = t.k;
    const float inv = t.active ? g1_inv_sigma(stat, k) : 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float dacc = 0.f;
    G1Row row;
    row.a_l = a_l, row.a_r = a_r, row.x = x, row.ldx = ldx, row.nb = nb, row.k = k, row.active = t.active;
    const float* Ok = O + static_cast<int64_t>(k) * M * kG1Wp;
    const float* csk = cs + static_cast<int64_t>(k) * M * 2;
    auto body = [&](const EntryGroup& g) {
      if (!t.active) return;
      int code[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        code[u] = g.c[u] >= B ? __ldg(codes + static_cast<int64_t>(g.node[u]) * nb + k) : 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (g.c[u] >= B) {
          const float2 sc = __ldg(reinterpret_cast<const float2*>(csk) + code[u]);
          const float* o = Ok + static_cast<int64_t>(code[u]) * kG1Wp;
          float f[4];
          ld_vec<4>(o, f);
          const float w = g.raw[u] * expf(lrelu((sc.x + row.ar) * inv, slope));
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(w * wu, f[i], acc[i]);
          dacc += w;
          if (HAS_RV) {
            float gr[4];
            ld_vec<4>(o + 4, gr);
            const float wr = g.rv[u] * expf(lrelu((row.al + sc.y) * inv, slope));
            fpart = fmaf(wr, dot4(row.xr, gr) + __ldg(o + 8), fpart);
          }
        } else if (g.c[u] >= 0) {
          float xv[4];
          ld_vec<4>(x + static_cast<int64_t>(g.c[u]) * ldx + 4 * k, xv);
          const float sl = __ldg(a_l + static_cast<int64_t>(g.c[u]) * nb + k);
          const float w = g.raw[u] * expf(lrelu((sl + row.ar) * inv, slope));
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(w, xv[i], acc[i]);
          dacc += w;
        }
      }
    };
    auto flush = [&](int r, bool whole) {
      if (t.active) {
        float* yp = y + static_cast<int64_t>(r) * ldy + 4 * k;
        float* dp = den + static_cast<int64_t>(r) * nb + k;
        if (whole) {
          st_vec<4>(yp, acc);
          *dp = dacc;
        } else {
          red_vec<4>(yp, acc);
          atomicAdd(dp, dacc);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = 0.f;
      dacc = 0.f;
    };
    walk_rows<HAS_RV>(t.eb, t.ee, t.row0, B, rowptr, col, val, rval, B, nullptr, lane, row, body, flush);
  }
  if (info) info_reduce(static_cast<double>(fpart), ws_sum, ws_count, info_scale, info);
}

// (F3) out = Y / (den + 1e-16), per branch
__global__ void gat1_normalize_kernel(int64_t B, int nb, float* __restrict__ y, int64_t ldy,
                                      const float* __restrict__ den) {
  const int64_t n = B * nb;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / nb;
    const int k = static_cast<int>(i - r * nb);
    const float inv = 1.f / (__ldg(den + i) + 1e-16f);
    float4* p = reinterpret_cast<float4*>(y + r * ldy + 4 * k);
    float4 v = *p;
    v.x *= inv, v.y *= inv, v.z *= inv, v.w *= inv;
    *p = v;
  }
}

// (B1) dY' = [dOut / (den + eps) | -<dOut, out> / (den + eps)] per branch: also the VQ hook's gradient [B, nb*5]
__global__ void gat1_bwd_prep_kernel(int64_t B, int nb, const float* __restrict__ dout, int64_t lddo,
                                     const float* __restrict__ out, int64_t ldo, const float* __restrict__ den,
                                     float* __restrict__ gy /* [B, nb*5] */) {
  const int64_t n = B * nb;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / nb;
    const int k = static_cast<int>(i - r * nb);
    const float inv = 1.f / (__ldg(den + i) + 1e-16f);
    float d[4], o[4];
    ld_vec<4>(dout + r * lddo + 4 * k, d);
    ld_vec<4>(out + r * ldo + 4 * k, o);
    float* p = gy + i * 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) p[c] = d[c] * inv;
    p[4] = -dot4(d, o) * inv;
  }
}

// (B2) one pass over the forward CSR: score gradients (batch: ds [B, nb, 2], codewords: dcs [nb, M, 2]) and the
// aggregation / info_backward parts of d x
__global__ void __launch_bounds__(kMpWarps * 32)
    gat1_bwd_edge_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                         const float* __restrict__ val, const float* __restrict__ rval,
                         const int32_t* __restrict__ chunk_row, int n_chunks, int chunk, int nnz, int B,
                         const float* __restrict__ x, int64_t ldx, const int16_t* __restrict__ codes,
                         const float* __restrict__ O, int nb, int M, float wu, const float* __restrict__ a_l,
                         const float* __restrict__ a_r, const float* __restrict__ cs,
                         const float* __restrict__ stat, float slope, int nslab, const float* __restrict__ gy,
                         const float* __restrict__ dinfo, float* __restrict__ ds /* [B, nb, 2] (l, r) */,
                         float* __restrict__ dcs, float* __restrict__ dx, int64_t lddx) {
  constexpr int U = kMpUnroll;
  const int lane = threadIdx.x & 31;
  const MpTask t = mp_task<4>(chunk_row, n_chunks, chunk, nnz, nslab, nb * 4, 4);
  if (!t.valid) return;
  const int k = t.k;
  const float inv = t.active ? g1_inv_sigma(stat, k) : 0.f;
  const float di = wu * (dinfo ? __ldg(dinfo) : 1.f);  // d loss / d (sum of reverse terms)

  struct Row : G1Row {
    const float* gy;
    float dy[4];
    float dd = 0.f;
    __device__ __forceinline__ void row_begin(int r) {
      G1Row::row_begin(r);
      if (!active) return;
      const float* p = gy + (static_cast<int64_t>(r) * nb + k) * 5;
#pragma unroll
      for (int c = 0; c < 4; ++c) dy[c] = __ldg(p + c);
      dd = __ldg(p + 4);
    }
  };
  Row row;
  row.a_l = a_l, row.a_r = a_r, row.x = x, row.ldx = ldx, row.nb = nb, row.k = k, row.active = t.active;
  row.gy = gy;
  const float* Ok = O + static_cast<int64_t>(k) * M * kG1Wp;
  const float* csk = cs + static_cast<int64_t>(k) * M * 2;
  float* dcsk = dcs + static_cast<int64_t>(k) * M * 2;
  float dsl = 0.f, dsr = 0.f;           // gradients of this row's own scores s_l[i,k], s_r[i,k]
  float dxa[4] = {0.f, 0.f, 0.f, 0.f};  // info_backward's direct gradient w.r.t. x[i, 4k:4k+4]

  auto body = [&](const EntryGroup& g) {
    if (!t.active) return;
    int code[U];
#pragma unroll
    for (int u = 0; u < U; ++u) code[u] = g.c[u] >= B ? __ldg(codes + static_cast<int64_t>(g.node[u]) * nb + k) : 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (g.c[u] >= B) {
        const float2 sc = __ldg(reinterpret_cast<const float2*>(csk) + code[u]);
        const float* o = Ok + static_cast<int64_t>(code[u]) * kG1Wp;
        float f[4], gr[4];
        ld_vec<4>(o, f);
        ld_vec<4>(o + 4, gr);
        // aggregation term: target i (s_r), source codeword (c_l)
        const float e = (sc.x + row.ar) * inv;
        const float w = g.raw[u] * expf(lrelu(e, slope));
        const float de = (wu * dot4(row.dy, f) + row.dd) * w * (e > 0.f ? 1.f : slope);
        dsr += de;
        // reverse (info_backward) term: source i (s_l), target codeword (c_r)
        const float e2 = (row.al + sc.y) * inv;
        const float wr = di * g.rv[u] * expf(lrelu(e2, slope));
        const float de2 = wr * (dot4(row.xr, gr) + __ldg(o + 8)) * (e2 > 0.f ? 1.f : slope);
        dsl += de2;
#pragma unroll
        for (int i = 0; i < 4; ++i) dxa[i] = fmaf(wr, gr[i], dxa[i]);
        atomicAdd(dcsk + 2 * code[u], de);
        atomicAdd(dcsk + 2 * code[u] + 1, de2);
      } else if (g.c[u] >= 0) {
        float xv[4];
        ld_vec<4>(x + static_cast<int64_t>(g.c[u]) * ldx + 4 * k, xv);
        const float sl = __ldg(a_l + static_cast<int64_t>(g.c[u]) * nb + k);
        const float e = (sl + row.ar) * inv;
        const float w = g.raw[u] * expf(lrelu(e, slope));
        const float de = (dot4(row.dy, xv) + row.dd) * w * (e > 0.f ? 1.f : slope);
        dsr += de;
        atomicAdd(ds + (static_cast<int64_t>(g.c[u]) * nb + k) * 2, de);  // source row's s_l
        if (dx) {
          float t4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) t4[i] = w * row.dy[i];
          red_vec<4>(dx + static_cast<int64_t>(g.c[u]) * lddx + 4 * k, t4);
        }
      }
    }
  };
  auto flush = [&](int r, bool) {
    if (t.active) {
      float* p = ds + (static_cast<int64_t>(r) * nb + k) * 2;
      atomicAdd(p, dsl);
      atomicAdd(p + 1, dsr);
      if (dx) red_vec<4>(dx + static_cast<int64_t>(r) * lddx + 4 * k, dxa);
    }
    dsl = dsr = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) dxa[i] = 0.f;
  };
  walk_rows<true>(t.eb, t.ee, t.row0, B, rowptr, col, val, rval, B, nullptr, lane, row, body, flush);
}

// (B3) per branch (one CTA each): ds -> da through a = s * sigma_k, in place on ds [B, nb, 2] and dcs [nb, M, 2]
__global__ void __launch_bounds__(512)
    gat1_score_grad_kernel(int B, int nb, int M, const float* __restrict__ a_l, const float* __restrict__ a_r,
                           const float* __restrict__ cs, const float* __restrict__ stat, float* __restrict__ ds,
                           float* __restrict__ dcs) {
  const int k = blockIdx.x, tid = threadIdx.x;
  __shared__ double red[16];
  __shared__ int arg[2];
  const float ml = stat[2 * k], mr = stat[2 * k + 1];
  if (tid < 2) arg[tid] = 0x7fffffff;
  __syncthreads();
  double s = 0.0;
  const int total = B + M;
  for (int n = tid; n < total; n += blockDim.x) {
    float al, ar, dl, dr;
    if (n < B) {
      al = a_l[static_cast<int64_t>(n) * nb + k], ar = a_r[static_cast<int64_t>(n) * nb + k];
      dl = ds[(static_cast<int64_t>(n) * nb + k) * 2], dr = ds[(static_cast<int64_t>(n) * nb + k) * 2 + 1];
    } else {
      const int64_t o = (static_cast<int64_t>(k) * M + (n - B)) * 2;
      al = cs[o], ar = cs[o + 1], dl = dcs[o], dr = dcs[o + 1];
    }
    s += static_cast<double>(dl) * al + static_cast<double>(dr) * ar;
    if (al == ml) atomicMin(&arg[0], n);
    if (ar == mr) atomicMin(&arg[1], n);
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
  const float sigma = ql * qr, invs = 1.f / sigma;
  const float dsigma = static_cast<float>(-red[0] / (static_cast<double>(sigma) * sigma));
  for (int n = tid; n < total; n += blockDim.x) {
    float* p = n < B ? ds + (static_cast<int64_t>(n) * nb + k) * 2 : dcs + (static_cast<int64_t>(k) * M + (n - B)) * 2;
    float dl = p[0] * invs, dr = p[1] * invs;
    if (n == arg[0]) dl += dsigma * (ml / ql) * qr;
    if (n == arg[1]) dr += dsigma * (mr / qr) * ql;
    p[0] = dl, p[1] = dr;
  }
}

// (B4) attention-vector gradients [nb, 5] x2 and the score path into d x.  thread -> (item, branch)
__global__ void __launch_bounds__(256)
    gat1_att_grad_kernel(int64_t B, int nb, int M, const float* __restrict__ x, int64_t ldx,
                         const float* __restrict__ O, float wu, const float* __restrict__ att_l,
                         const float* __restrict__ att_r, const float* __restrict__ da /* ds after B3 */,
                         const float* __restrict__ dca /* dcs after B3 */, float* __restrict__ datt_l,
                         float* __restrict__ datt_r, float* __restrict__ dx, int64_t lddx) {
  const int64_t total = (B + M) * nb;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t it = t / nb;
    const int k = static_cast<int>(t - it * nb);
    float v[4];
    float s = 1.f, dl, dr;
    if (it < B) {
      ld_vec<4>(x + it * ldx + 4 * k, v);
      dl = __ldg(da + (it * nb + k) * 2), dr = __ldg(da + (it * nb + k) * 2 + 1);
      if (dx) {
        float* p = dx + it * lddx + 4 * k;
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] += dl * __ldg(att_l + 5 * k + i) + dr * __ldg(att_r + 5 * k + i);
      }
    } else {
      const int64_t o = static_cast<int64_t>(k) * M + (it - B);
      ld_vec<4>(O + o * kG1Wp, v);
      s = wu;
      dl = __ldg(dca + 2 * o), dr = __ldg(dca + 2 * o + 1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(datt_l + 5 * k + i, dl * s * v[i]);
      atomicAdd(datt_r + 5 * k + i, dr * s * v[i]);
    }
    atomicAdd(datt_l + 5 * k + 4, dl);
    atomicAdd(datt_r + 5 * k + 4, dr);
  }
}

static int g1_grid(int64_t n) { return static_cast<int>(std::min<int64_t>((n + 255) / 256, 8 * kNumSMs)); }

}  // namespace vqgnn

using namespace vqgnn;

#define VQ_G1_COMMON_CHECKS(name)                                                                            \
  VQ_CHECK_ARG(D == 4 && Wp == kG1Wp, name ": the v1 GAT kernels need num_D = 4 and the add_flag codebook "  \
                                           "layout Wp = 12 (got D=%d Wp=%d)", D, Wp);                        \
  VQ_CHECK_ARG(B > 0 && B < (1ll << 31) && nb > 0 && M > 0, name ": bad sizes");                            \
  VQ_CHECK_ARG(ldx % 4 == 0 && aligned16(x) && aligned16(O), name ": x and O must be 16 B aligned")

extern "C" int vqgnn_gat1_scores(int64_t B, const float* x, int64_t ldx, const float* O, int nb, int M, int D,
                                 int Wp, float wu, const float* att_l, const float* att_r, float* a_l, float* a_r,
                                 float* cs, float* stat, void* stream) {
  VQ_CHECK_ARG(x && O && att_l && att_r && a_l && a_r && cs && stat, "gat1_scores: null argument");
  VQ_G1_COMMON_CHECKS("gat1_scores");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  gat1_stat_init_kernel<<<ceil_div(2 * nb, 128), 128, 0, s>>>(stat, 2 * nb);
  VQ_LAUNCH_CHECK();
  gat1_scores_kernel<<<g1_grid((B + M) * nb), 256, 0, s>>>(B, nb, M, x, ldx, O, wu, att_l, att_r, a_l, a_r, cs, stat);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_gat1_fwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* rval,
                              const int32_t* chunk_row, int chunk, int64_t nnz, int64_t B, const float* x,
                              int64_t ldx, const int16_t* codes, const float* O, int nb, int M, int D, int Wp,
                              float wu, const float* a_l, const float* a_r, const float* cs, const float* stat,
                              float negative_slope, float* y, int64_t ldy, float* den, float* info, void* ws,
                              void* stream) {
  VQ_CHECK_ARG(rowptr && col && val && x && codes && O && a_l && a_r && cs && stat && y && den,
               "gat1_fwd: null argument");
  VQ_G1_COMMON_CHECKS("gat1_fwd");
  VQ_CHECK_ARG(ldy % 4 == 0 && aligned16(y), "gat1_fwd: y must be 16 B aligned");
  VQ_CHECK_ARG(nnz >= 0 && nnz < (1ll << 31) && chunk > 0 && chunk % 32 == 0 && (nnz == 0 || chunk_row),
               "gat1_fwd: bad partition");
  VQ_CHECK_ARG(!info || (ws && rval), "gat1_fwd: info needs a workspace and reverse values");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = zero_rows(y, B, nb * 4, ldy, s)) return rc;
  VQ_CUDA(cudaMemsetAsync(den, 0, sizeof(float) * B * nb, s));
  double* ws_sum = static_cast<double*>(ws);
  unsigned int* ws_count = ws ? reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 8) : nullptr;
  if (info) VQ_CUDA(cudaMemsetAsync(ws, 0, 16, s));
  const int n_chunks = static_cast<int>((nnz + chunk - 1) / chunk);
  if (n_chunks == 0) {
    if (info) VQ_CUDA(cudaMemsetAsync(info, 0, sizeof(float), s));
    return VQGNN_OK;
  }
  const int nslab = ceil_div(nb, 32);
  const int grid = ceil_div(static_cast<int64_t>(n_chunks) * nslab, kMpWarps);
  if (rval)
    gat1_fwd_kernel<true><<<grid, kMpWarps * 32, 0, s>>>(rowptr, col, val, rval, chunk_row, n_chunks, chunk, (int)nnz,
                                                         (int)B, x, ldx, codes, O, nb, M, wu, a_l, a_r, cs, stat,
                                                         negative_slope, nslab, y, ldy, den, info, wu, ws_sum, ws_count);
  else
    gat1_fwd_kernel<false><<<grid, kMpWarps * 32, 0, s>>>(rowptr, col, val, nullptr, chunk_row, n_chunks, chunk,
                                                          (int)nnz, (int)B, x, ldx, codes, O, nb, M, wu, a_l, a_r, cs,
                                                          stat, negative_slope, nslab, y, ldy, den, nullptr, wu,
                                                          ws_sum, ws_count);
  VQ_LAUNCH_CHECK();
  gat1_normalize_kernel<<<g1_grid(B * nb), 256, 0, s>>>(B, nb, y, ldy, den);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_gat1_bwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* rval,
                              const int32_t* chunk_row, int chunk, int64_t nnz, int64_t B, const float* x,
                              int64_t ldx, const int16_t* codes, const float* O, int nb, int M, int D, int Wp,
                              float wu, const float* att_l, const float* att_r, const float* a_l, const float* a_r,
                              const float* cs, const float* stat, float negative_slope, const float* out,
                              int64_t ldo, const float* den, const float* dout, int64_t lddo, const float* dinfo,
                              float* gy, float* ds, float* dcs, float* dx, int64_t lddx, float* datt_l,
                              float* datt_r, void* stream) {
  VQ_CHECK_ARG(rowptr && col && val && rval && x && codes && O && att_l && att_r && a_l && a_r && cs && stat && out &&
                   den && dout && gy && ds && dcs && datt_l && datt_r,
               "gat1_bwd: null argument");
  VQ_G1_COMMON_CHECKS("gat1_bwd");
  VQ_CHECK_ARG(lddo % 4 == 0 && ldo % 4 == 0 && aligned16(dout) && aligned16(out) && (!dx || (lddx % 4 == 0 && aligned16(dx))),
               "gat1_bwd: operands must be 16 B aligned");
  VQ_CHECK_ARG(nnz >= 0 && nnz < (1ll << 31) && chunk > 0 && chunk % 32 == 0 && (nnz == 0 || chunk_row),
               "gat1_bwd: bad partition");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  gat1_bwd_prep_kernel<<<g1_grid(B * nb), 256, 0, s>>>(B, nb, dout, lddo, out, ldo, den, gy);
  VQ_LAUNCH_CHECK();
  VQ_CUDA(cudaMemsetAsync(ds, 0, sizeof(float) * B * nb * 2, s));
  VQ_CUDA(cudaMemsetAsync(dcs, 0, sizeof(float) * static_cast<size_t>(nb) * M * 2, s));
  VQ_CUDA(cudaMemsetAsync(datt_l, 0, sizeof(float) * nb * 5, s));
  VQ_CUDA(cudaMemsetAsync(datt_r, 0, sizeof(float) * nb * 5, s));
  if (dx)
    if (int rc = zero_rows(dx, B, nb * 4, lddx, s)) return rc;
  const int n_chunks = static_cast<int>((nnz + chunk - 1) / chunk);
  if (n_chunks > 0) {
    const int nslab = ceil_div(nb, 32);
    const int grid = ceil_div(static_cast<int64_t>(n_chunks) * nslab, kMpWarps);
    gat1_bwd_edge_kernel<<<grid, kMpWarps * 32, 0, s>>>(rowptr, col, val, rval, chunk_row, n_chunks, chunk, (int)nnz,
                                                        (int)B, x, ldx, codes, O, nb, M, wu, a_l, a_r, cs, stat,
                                                        negative_slope, nslab, gy, dinfo, ds, dcs, dx, lddx);
    VQ_LAUNCH_CHECK();
  }
  gat1_score_grad_kernel<<<nb, 512, 0, s>>>((int)B, nb, M, a_l, a_r, cs, stat, ds, dcs);
  VQ_LAUNCH_CHECK();
  gat1_att_grad_kernel<<<g1_grid((B + M) * nb), 256, 0, s>>>(B, nb, M, x, ldx, O, wu, att_l, att_r, ds, dcs, datt_l,
                                                             datt_r, dx, lddx);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
