// VQ-approximated message passing, GCN / SAGE-Mean: gather-SpMM over the batch plan.
// One warp per (output row, 32*VEC-column slab).  In-batch neighbours read dense rows (coalesced
// 16 B per lane); out-of-batch neighbours read the node's code row (2 B per lane, one 64 B line at
// nb = 32) and gather their codeword from the L2-resident codebook.  HBM-bound integer/float gather
// work: no tensor cores.  Reference maths: vq_gnn_v2/models.py:161-198, vq_gnn_v2/convs.py:65-101,
// vq_gnn_v1/models.py:170-223 + vq_gnn_v1/utils/dataloader.py:144-192.
#include "common.cuh"

namespace vqgnn {

constexpr int kMpWarps = 8;

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

struct Codebook {
  const int32_t* tail_node;  // [T] or nullptr (identity)
  const int16_t* codes;      // [N, nb]
  const float* O;            // [nb, M, Wp]
  int nb, M, D, Wp;
};

// Accumulate one CSR row for the lane's VEC columns starting at column c0 (branch k, offset off).
//   acc += val * (src < B ? dense[src, c0..] : tscale * O_k[code, half_off + off ..])
//   gqa += rval * O_k[code, D + off ..]     (HAS_GQ, tail entries only)
template <int VEC, bool HAS_GQ>
__device__ __forceinline__ void gather_row(int e0, int e1, const int32_t* __restrict__ col,
                                           const float* __restrict__ val, const float* __restrict__ rval,
                                           int B, const float* __restrict__ dense, int64_t ldd,
                                           const Codebook& cb, int half_off, float tscale, bool active,
                                           int c0, int k, int off, int lane, float (&acc)[VEC],
                                           float (&gqa)[VEC]) {
  constexpr int U = 4;
  for (int eb = e0; eb < e1; eb += 32) {
    const int e = eb + lane;
    int c_l = -1, node_l = 0;
    float v_l = 0.f, rv_l = 0.f;
    if (e < e1) {
      c_l = __ldg(col + e);
      v_l = __ldg(val + e);
      if (HAS_GQ) rv_l = __ldg(rval + e);
      if (c_l >= B) node_l = cb.tail_node ? __ldg(cb.tail_node + (c_l - B)) : (c_l - B);
    }
    const int cnt = min(32, e1 - eb);
    for (int j = 0; j < cnt; j += U) {
      int c[U], node[U];
      float v[U], rv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int src_lane = min(j + u, 31);
        c[u] = __shfl_sync(0xffffffffu, c_l, src_lane);
        v[u] = __shfl_sync(0xffffffffu, v_l, src_lane);
        node[u] = __shfl_sync(0xffffffffu, node_l, src_lane);
        rv[u] = HAS_GQ ? __shfl_sync(0xffffffffu, rv_l, src_lane) : 0.f;
        if (j + u >= cnt) c[u] = -1;
      }
      if (!active) continue;
      const float* p[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // first level: code loads for tail entries (independent)
        p[u] = nullptr;
        if (c[u] >= B) {
          const int code = __ldg(cb.codes + static_cast<int64_t>(node[u]) * cb.nb + k);
          p[u] = cb.O + (static_cast<int64_t>(k) * cb.M + code) * cb.Wp + off;
        } else if (c[u] >= 0) {
          p[u] = dense + static_cast<int64_t>(c[u]) * ldd + c0;
        }
      }
      Vec<VEC> a[U], gq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // second level: the gathers
        if (c[u] >= B) {
          a[u].load(p[u] + half_off);
          if (HAS_GQ) gq[u].load(p[u] + cb.D);
        } else if (c[u] >= 0) {
          a[u].load(p[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c[u] >= B) {
          const float s = v[u] * tscale;
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(s, a[u].v[i], acc[i]);
          if (HAS_GQ) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) gqa[i] = fmaf(rv[u], gq[u].v[i], gqa[i]);
          }
        } else if (c[u] >= 0) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(v[u], a[u].v[i], acc[i]);
        }
      }
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

template <int VEC, bool HAS_GQ>
__global__ void __launch_bounds__(kMpWarps * 32)
    mp_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                  const float* __restrict__ val, const float* __restrict__ rval, int64_t R, int B,
                  const float* __restrict__ x, int64_t ldx, Codebook cb, int C, int nslab, float feat_scale,
                  float info_scale, float* __restrict__ y, int64_t ldy, float* __restrict__ gq, int64_t ldgq,
                  float* __restrict__ info, double* ws_sum, unsigned int* ws_count) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = static_cast<int64_t>(blockIdx.x) * kMpWarps + warp;
  double part = 0.0;
  if (task < R * nslab) {
    const int64_t r = task / nslab;
    const int slab = static_cast<int>(task - r * nslab);
    const int c0 = (slab * 32 + lane) * VEC;
    const bool active = c0 < C;
    const int k = active ? c0 / cb.D : 0, off = active ? c0 - k * cb.D : 0;
    float acc[VEC], gqa[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f, gqa[i] = 0.f;
    gather_row<VEC, HAS_GQ>(__ldg(rowptr + r), __ldg(rowptr + r + 1), col, val, rval, B, x, ldx, cb, 0,
                            feat_scale, active, c0, k, off, lane, acc, gqa);
    if (active) {
      if (r < B) {
        Vec<VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
        o.store(y + r * ldy + c0);
        if (HAS_GQ) {
          Vec<VEC> q, xr;
#pragma unroll
          for (int i = 0; i < VEC; ++i) q.v[i] = gqa[i];
          if (gq) q.store(gq + r * ldgq + c0);
          xr.load(x + r * ldx + c0);
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < VEC; ++i) d = fmaf(xr.v[i], gqa[i], d);
          part = d;
        }
      } else if (info) {  // v2: <Y[r], Gq[r]> with Gq the node's own gradient codeword
        const int node = cb.tail_node ? __ldg(cb.tail_node + (r - B)) : static_cast<int>(r - B);
        const int code = __ldg(cb.codes + static_cast<int64_t>(node) * cb.nb + k);
        Vec<VEC> gv;
        gv.load(cb.O + (static_cast<int64_t>(k) * cb.M + code) * cb.Wp + cb.D + off);
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) d = fmaf(acc[i], gv.v[i], d);
        part = d;
      }
    }
  }
  if (info) info_reduce(part, ws_sum, ws_count, info_scale, info);
}

template <int VEC>
__global__ void __launch_bounds__(kMpWarps * 32)
    mp_bwd_kernel(const int32_t* __restrict__ browptr, const int32_t* __restrict__ brow,
                  const float* __restrict__ bval, int B, const float* __restrict__ dy, int64_t lddy,
                  Codebook cb, int C, int nslab, float tail_scale, const float* __restrict__ gq, int64_t ldgq,
                  float gq_scale, const float* __restrict__ dinfo, float* __restrict__ dx, int64_t lddx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = static_cast<int64_t>(blockIdx.x) * kMpWarps + warp;
  if (task >= static_cast<int64_t>(B) * nslab) return;
  const int64_t j = task / nslab;
  const int slab = static_cast<int>(task - j * nslab);
  const int c0 = (slab * 32 + lane) * VEC;
  const bool active = c0 < C;
  const int k = active ? c0 / cb.D : 0, off = active ? c0 - k * cb.D : 0;
  const float di = dinfo ? __ldg(dinfo) : 1.f;
  float acc[VEC], unused[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f, unused[i] = 0.f;
  gather_row<VEC, false>(__ldg(browptr + j), __ldg(browptr + j + 1), brow, bval, nullptr, B, dy, lddy, cb,
                         cb.D, tail_scale * di, active, c0, k, off, lane, acc, unused);
  if (!active) return;
  if (gq) {
    Vec<VEC> q;
    q.load(gq + j * ldgq + c0);
    const float s = gq_scale * di;
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = fmaf(s, q.v[i], acc[i]);
  }
  Vec<VEC> o;
#pragma unroll
  for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
  o.store(dx + j * lddx + c0);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vqgnn

using namespace vqgnn;

extern "C" size_t vqgnn_mp_workspace_bytes(void) { return 64; }

extern "C" int vqgnn_mp_fwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* rval,
                            int64_t R, int64_t B, const float* x, int64_t ldx, const int32_t* tail_node,
                            const int16_t* codes, const float* O, int nb, int M, int D, int Wp, float feat_scale,
                            float info_scale, float* y, int64_t ldy, float* gq, int64_t ldgq, float* info,
                            void* ws, void* stream) {
  VQ_CHECK_ARG(rowptr && col && val && x && codes && O && y, "mp_fwd: null argument");
  VQ_CHECK_ARG(R >= B && B > 0 && nb > 0 && D > 0 && Wp >= 2 * D, "mp_fwd: bad sizes");
  VQ_CHECK_ARG(!info || ws, "mp_fwd: info needs a workspace");
  VQ_CHECK_ARG(B < (1ll << 31) && R < (1ll << 31), "mp_fwd: too many rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = nb * D;
  Codebook cb{tail_node, codes, O, nb, M, D, Wp};
  double* ws_sum = static_cast<double*>(ws);
  unsigned int* ws_count = ws ? reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 8) : nullptr;
  if (info) VQ_CUDA(cudaMemsetAsync(ws, 0, 16, s));
  const bool vec4 = (D % 4 == 0) && (Wp % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(x) &&
                    aligned16(y) && aligned16(O) && (!gq || (ldgq % 4 == 0 && aligned16(gq)));
  const int vec = vec4 ? 4 : 1;
  const int nslab = ceil_div(C, 32 * vec);
  const int64_t tasks = R * nslab;
  const int grid = ceil_div(tasks, kMpWarps);
#define VQ_MP_FWD(VEC, GQ)                                                                                  \
  mp_fwd_kernel<VEC, GQ><<<grid, kMpWarps * 32, 0, s>>>(rowptr, col, val, rval, R, (int)B, x, ldx, cb, C,   \
                                                        nslab, feat_scale, info_scale, y, ldy, gq, ldgq,    \
                                                        info, ws_sum, ws_count)
  if (vec4) {
    if (rval) VQ_MP_FWD(4, true);
    else VQ_MP_FWD(4, false);
  } else {
    if (rval) VQ_MP_FWD(1, true);
    else VQ_MP_FWD(1, false);
  }
#undef VQ_MP_FWD
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_mp_bwd(const int32_t* browptr, const int32_t* brow, const float* bval, int64_t B,
                            const float* dy, int64_t lddy, const int32_t* tail_node, const int16_t* codes,
                            const float* O, int nb, int M, int D, int Wp, float tail_scale, const float* gq,
                            int64_t ldgq, float gq_scale, const float* dinfo, float* dx, int64_t lddx,
                            void* stream) {
  VQ_CHECK_ARG(browptr && brow && bval && dy && codes && O && dx, "mp_bwd: null argument");
  VQ_CHECK_ARG(B > 0 && B < (1ll << 31) && nb > 0 && D > 0 && Wp >= 2 * D, "mp_bwd: bad sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = nb * D;
  Codebook cb{tail_node, codes, O, nb, M, D, Wp};
  const bool vec4 = (D % 4 == 0) && (Wp % 4 == 0) && (lddy % 4 == 0) && (lddx % 4 == 0) && aligned16(dy) &&
                    aligned16(dx) && aligned16(O) && (!gq || (ldgq % 4 == 0 && aligned16(gq)));
  const int vec = vec4 ? 4 : 1;
  const int nslab = ceil_div(C, 32 * vec);
  const int grid = ceil_div(B * nslab, kMpWarps);
  if (vec4)
    mp_bwd_kernel<4><<<grid, kMpWarps * 32, 0, s>>>(browptr, brow, bval, (int)B, dy, lddy, cb, C, nslab,
                                                    tail_scale, gq, ldgq, gq_scale, dinfo, dx, lddx);
  else
    mp_bwd_kernel<1><<<grid, kMpWarps * 32, 0, s>>>(browptr, brow, bval, (int)B, dy, lddy, cb, C, nslab,
                                                    tail_scale, gq, ldgq, gq_scale, dinfo, dx, lddx);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
