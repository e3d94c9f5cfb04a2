// VectorQuantizerEMA on sm_100a: moments -> whitening affine -> fused assign + segmented sums -> EMA/recovery.
// Reference arithmetic: vq_gnn_v2/vq.py:160-279.  All branches of a layer per launch.
#include "common.cuh"

namespace vqgnn {

// ------------------------------------------------------------------------------------------------
// (1) column moments, fp64 accumulation.  HBM-bound: reads x (and g) once.
// block = (32 columns, 8 row lanes); grid = (row tiles, column tiles).  Two levels, no atomics: every row tile
// writes its partial sums, vq_moments_reduce_kernel adds the tiles in index order (bit-stable results).
// ------------------------------------------------------------------------------------------------
constexpr int kMomRowsPerBlock = 512;

__global__ void __launch_bounds__(256) vq_moments_kernel(const float* __restrict__ x, int64_t ldx,
                                                         const float* __restrict__ g, int64_t ldg,
                                                         int64_t B, int C, int Cg, int tilesC,
                                                         double* __restrict__ part) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ctile = blockIdx.y;
  const float* src;
  int64_t ld;
  int c, width, out_c;
  if (ctile < tilesC) {
    src = x, ld = ldx, c = ctile * 32 + tx, width = C, out_c = c;
  } else {
    src = g, ld = ldg, c = (ctile - tilesC) * 32 + tx, width = Cg, out_c = C + c;
  }
  const int Ctot = C + Cg;
  double s1 = 0.0, s2 = 0.0;
  if (c < width) {
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kMomRowsPerBlock;
    const int64_t r1 = min(r0 + kMomRowsPerBlock, B);
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const double v = static_cast<double>(__ldg(src + r * ld + c));
      s1 += v;
      s2 += v * v;
    }
  }
  __shared__ double sh1[8][33], sh2[8][33];
  sh1[ty][tx] = s1, sh2[ty][tx] = s2;
  __syncthreads();
  if (ty == 0 && c < width) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s1 += sh1[i][tx], s2 += sh2[i][tx];
    double* dst = part + static_cast<int64_t>(blockIdx.x) * 2 * Ctot;
    dst[out_c] = s1;
    dst[Ctot + out_c] = s2;
  }
}

__global__ void vq_moments_reduce_kernel(const double* __restrict__ part, int tiles, int n, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double t = 0.0;
  for (int i = 0; i < tiles; ++i) t += part[static_cast<int64_t>(i) * n + c];
  sums[c] = t;
}

// ------------------------------------------------------------------------------------------------
// (2) whitening affine + BatchNorm running-stat update (one thread per column; tiny)
// ------------------------------------------------------------------------------------------------
__global__ void vq_whiten_kernel(const double* __restrict__ sums, double count, const double* d_count, int nb, int D, int Dg,
                                 int has_grad, float* run_mean_f, float* run_var_f, float* run_mean_g,
                                 float* run_var_g, float eps_f, float mom_f, float eps_g, float mom_g,
                                 float gs0, float gs1, int training, int seed_running,
                                 const int32_t* __restrict__ seed_mask, int64_t* nbt_f, int64_t* nbt_g,
                                 float* __restrict__ scale, float* __restrict__ shift) {
  const int C = nb * D, Cg = has_grad ? nb * Dg : 0, Ctot = C + Cg;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Ctot) return;
  const bool is_g = c >= C;
  float* rm = is_g ? run_mean_g + (c - C) : run_mean_f + c;
  float* rv = is_g ? run_var_g + (c - C) : run_var_f + c;
  const float eps = is_g ? eps_g : eps_f;
  const float mom = is_g ? mom_g : mom_f;
  float gs = 1.f;
  if (is_g) {
    const int d = (c - C) % Dg;
    gs = (d < D) ? gs0 : gs1;
  }
  float mean_f, var_f;  // statistics used to normalise (fp32, like ATen)
  if (seed_mask) seed_running = seed_mask[is_g ? (c - C) / Dg : c / D];  // per branch (one bn_inited per quantiser)
  if (training || seed_running) {
    if (d_count) count = *d_count;
    const double mean = sums[c] / count;
    double var_b = sums[Ctot + c] / count - mean * mean;
    if (var_b < 0.0) var_b = 0.0;
    const double var_u = count > 1.0 ? var_b * (count / (count - 1.0)) : var_b;
    mean_f = static_cast<float>(mean), var_f = static_cast<float>(var_b);
    float m_run = *rm, v_run = *rv;
    if (seed_running) {  // vq.py:216-221: running stats <- batch mean / unbiased var, then BN's own update
      m_run = mean_f, v_run = static_cast<float>(var_u);
    }
    if (!training) {  // eval-mode update() on a fresh quantiser: seeded, then normalised with the running statistics
      *rm = m_run, *rv = v_run;
      mean_f = m_run, var_f = v_run;
    } else {
    *rm = (1.f - mom) * m_run + mom * mean_f;
    *rv = (1.f - mom) * v_run + mom * static_cast<float>(var_u);
    // BatchNorm1d.num_batches_tracked += 1 per training forward (one counter per branch)
    if (!is_g && nbt_f && c % D == 0) nbt_f[c / D] += 1;
    if (is_g && nbt_g && (c - C) % Dg == 0) nbt_g[(c - C) / Dg] += 1;
    }
  } else {
    mean_f = *rm, var_f = *rv;
  }
  const float invstd = 1.0f / sqrtf(var_f + eps);
  scale[c] = invstd * gs;
  shift[c] = -mean_f * invstd * gs;
}

// ------------------------------------------------------------------------------------------------
// (3) exact-fp32 SIMT assignment (the parity anchor).
// CTA = 128 threads, each owning RPT rows of one branch; the branch's codebook streams through
// shared memory in chunks of kChunk codewords (broadcast LDS.128 reads: no bank conflicts).
// NV = float4s per padded codeword row (Wp/4).  Columns >= w_use are masked to zero when staging,
// so the same code serves feature_update (w_use = D) and update (w_use = D+Dg).
// ------------------------------------------------------------------------------------------------
constexpr int kAssignThreads = 128;
constexpr int kChunk = 1024;

template <int NV, int RPT>
__global__ void __launch_bounds__(kAssignThreads)
    vq_assign_simt_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g, int64_t ldg,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ E, int64_t B, int nb, int M, int D, int Dg, int Wp,
                          const int32_t* __restrict__ batch_idx, int16_t* __restrict__ codes,
                          int64_t codes_ld, int16_t* __restrict__ idx, float* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* Es = reinterpret_cast<float4*>(smem_raw);                   // [kChunk][NV]
  float* c2s = reinterpret_cast<float*>(smem_raw + sizeof(float4) * NV * kChunk);  // [kChunk]

  const int k = blockIdx.y;
  const int C = nb * D;
  const int w_use = D + (g != nullptr ? Dg : 0);
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * (kAssignThreads * RPT);

  // ---- whiten this thread's rows into registers --------------------------------------------------
  float z[RPT][NV * 4];
  float x2[RPT], best[RPT];
  int besti[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int64_t b = row0 + threadIdx.x + static_cast<int64_t>(i) * kAssignThreads;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < NV * 4; ++w) {
      float v = 0.f;
      if (b < B) {
        if (w < D) {
          const int c = k * D + w;
          v = fmaf(__ldg(x + b * ldx + c), __ldg(scale + c), __ldg(shift + c));
        } else if (w < w_use) {
          const int cg = k * Dg + (w - D);
          v = fmaf(__ldg(g + b * ldg + cg), __ldg(scale + C + cg), __ldg(shift + C + cg));
        }
      }
      z[i][w] = v;
      acc = fmaf(v, v, acc);
    }
    x2[i] = acc, best[i] = __int_as_float(0x7f800000), besti[i] = 0;
  }

  // ---- stream the codebook --------------------------------------------------------------------------
  const float* Ek = E + static_cast<int64_t>(k) * M * Wp;
  for (int m0 = 0; m0 < M; m0 += kChunk) {
    const int mc = min(kChunk, M - m0);
    __syncthreads();
    for (int m = threadIdx.x; m < mc; m += kAssignThreads) {
      const float4* src = reinterpret_cast<const float4*>(Ek + static_cast<int64_t>(m0 + m) * Wp);
      float c2 = 0.f;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        float4 e = __ldg(src + q);
        if (q * 4 + 0 >= w_use) e.x = 0.f;
        if (q * 4 + 1 >= w_use) e.y = 0.f;
        if (q * 4 + 2 >= w_use) e.z = 0.f;
        if (q * 4 + 3 >= w_use) e.w = 0.f;
        Es[m * NV + q] = e;
        c2 = fmaf(e.x, e.x, c2), c2 = fmaf(e.y, e.y, c2), c2 = fmaf(e.z, e.z, c2), c2 = fmaf(e.w, e.w, c2);
      }
      c2s[m] = c2;
    }
    __syncthreads();
#pragma unroll 2
    for (int m = 0; m < mc; ++m) {
      float4 e[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) e[q] = Es[m * NV + q];
      const float c2 = c2s[m];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          dot = fmaf(z[i][q * 4 + 0], e[q].x, dot);
          dot = fmaf(z[i][q * 4 + 1], e[q].y, dot);
          dot = fmaf(z[i][q * 4 + 2], e[q].z, dot);
          dot = fmaf(z[i][q * 4 + 3], e[q].w, dot);
        }
        const float d = fmaf(-2.f, dot, x2[i] + c2);  // (||z||^2 + ||e||^2) - 2 z.e   (vq.py:230-232)
        if (d < best[i]) best[i] = d, besti[i] = m0 + m;  // strict <: lowest index wins ties
      }
    }
  }

  // ---- outputs: code, code table scatter, per-codeword sums and counts --------------------------
  const int Ws = Wp + 4;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int64_t b = row0 + threadIdx.x + static_cast<int64_t>(i) * kAssignThreads;
    if (b >= B) continue;
    const int code = besti[i];
    if (idx) idx[b * nb + k] = static_cast<int16_t>(code);
    if (codes) codes[static_cast<int64_t>(__ldg(batch_idx + b)) * codes_ld + k] = static_cast<int16_t>(code);
    if (stats) {
      float* dst = stats + (static_cast<int64_t>(k) * M + code) * Ws;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        if (q * 4 < w_use)
          atomicAdd(reinterpret_cast<float4*>(dst) + q,
                    make_float4(z[i][q * 4], z[i][q * 4 + 1], z[i][q * 4 + 2], z[i][q * 4 + 3]));
      }
      atomicAdd(dst + Wp, 1.0f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// (4) EMA + Laplace smoothing + codeword recovery, two launches:
//   vq_finalize_size_kernel : one CTA per branch -- cluster sizes (needs the branch total for the Laplace smoothing)
//   vq_finalize_code_kernel : (branch, tile of codewords) CTAs -- EMA sums, codewords, recovery; a branch whose sizes
//                             contain a zero is left untouched ('Bad Init!', vq.py:188-189 / 253-254)
// ------------------------------------------------------------------------------------------------
constexpr int kFinThreads = 512;
constexpr int kFinTile = 256;   // codewords per CTA of the second launch

__global__ void __launch_bounds__(kFinThreads)
    vq_finalize_size_kernel(const float* __restrict__ stats, int M, int Wp, float decay, float omd, int warm_up,
                            float* __restrict__ ema_size, int32_t* __restrict__ status) {
  const int k = blockIdx.x, tid = threadIdx.x;
  const int Ws = Wp + 4;
  const float* st = stats + static_cast<int64_t>(k) * M * Ws;
  float* size = ema_size + static_cast<int64_t>(k) * M;
  __shared__ float red[kFinThreads / 32];
  __shared__ float total_sh;
  __shared__ int bad_sh;
  if (tid == 0) bad_sh = 0;

  // size <- decay*size + (1-decay)*count ; n = sum(size)
  float part = 0.f;
  for (int m = tid; m < M; m += kFinThreads) {
    const float s = size[m] * decay + omd * st[static_cast<int64_t>(m) * Ws + Wp];
    size[m] = s;
    part += s;
  }
  part = warp_sum(part);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < kFinThreads / 32; ++i) t += red[i];
    total_sh = t;
  }
  __syncthreads();
  const float n = total_sh;
  int bad = 0;
  for (int m = tid; m < M; m += kFinThreads) {
    float s = size[m];
    if (warm_up) {  // Laplace smoothing (vq.py:182-186)
      s = (s + 1e-5f) / (n + static_cast<float>(static_cast<double>(M) * 1e-5)) * n;
      size[m] = s;
    }
    if (s == 0.f) bad = 1;
  }
  if (bad) atomicOr(&bad_sh, 1);
  __syncthreads();
  if (tid == 0 && bad_sh) atomicOr(status, VQGNN_STATUS_BAD_INIT);
}

__global__ void __launch_bounds__(256)
    vq_finalize_code_kernel(const float* __restrict__ stats, int M, int D, int Dg, int Wp, int joint, float decay,
                            float omd, float eps, float gs0, float gs1, const float* __restrict__ run_mean_f,
                            const float* __restrict__ run_var_f, const float* __restrict__ run_mean_g,
                            const float* __restrict__ run_var_g, const float* __restrict__ ema_size,
                            float* __restrict__ ema_w, float* __restrict__ E, float* __restrict__ O) {
  const int k = blockIdx.y, tid = threadIdx.x;
  const int Ws = Wp + 4;
  const float* st = stats + static_cast<int64_t>(k) * M * Ws;
  const float* size = ema_size + static_cast<int64_t>(k) * M;
  // the reference raises before it touches _ema_w / _embedding / _embedding_output: skip the whole branch
  int bad = 0;
  for (int m = tid; m < M; m += 256) bad |= (size[m] == 0.f);
  if (__syncthreads_or(bad)) return;

  const int W = D + Dg;
  const int wlim = joint ? W : D;
  const float div0 = static_cast<float>(static_cast<double>(gs0) + static_cast<double>(eps));
  const float div1 = static_cast<float>(static_cast<double>(gs1) + static_cast<double>(eps));
  const int m0 = blockIdx.x * kFinTile, m1 = min(m0 + kFinTile, M);
  for (int i = tid; i < (m1 - m0) * wlim; i += 256) {
    const int m = m0 + i / wlim, w = i - (i / wlim) * wlim;
    const int64_t off = (static_cast<int64_t>(k) * M + m) * Wp + w;
    const float wm = ema_w[off] * decay + omd * st[static_cast<int64_t>(m) * Ws + w];
    ema_w[off] = wm;
    const float e = wm / size[m];
    E[off] = e;
    float o;
    if (w < D) {
      o = e * sqrtf(run_var_f[k * D + w] + 1e-5f) + run_mean_f[k * D + w];
    } else {
      const int dg = w - D;
      const float t = e / (dg < D ? div0 : div1);
      o = t * sqrtf(run_var_g[k * Dg + dg] + eps) + run_mean_g[k * Dg + dg];
      if (gs0 == 0.f) o = 0.f;
    }
    O[off] = o;
  }
}

}  // namespace vqgnn

using namespace vqgnn;

extern "C" size_t vqgnn_vq_moments_workspace_bytes(int64_t B, int C, int Cg) {
  return static_cast<size_t>(ceil_div(B, kMomRowsPerBlock)) * 2 * (C + Cg) * sizeof(double);
}

extern "C" int vqgnn_vq_moments(const float* x, int64_t ldx, const float* g, int64_t ldg, int64_t B, int C,
                                int Cg, double* sums, void* ws, void* stream) {
  VQ_CHECK_ARG(x && sums && ws && B > 0 && C > 0, "vq_moments: bad arguments");
  if (!g) Cg = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tilesC = ceil_div(C, 32), tilesG = ceil_div(Cg, 32);
  const int tilesR = ceil_div(B, kMomRowsPerBlock);
  dim3 grid(tilesR, tilesC + tilesG), block(32, 8);
  double* part = static_cast<double*>(ws);
  vq_moments_kernel<<<grid, block, 0, s>>>(x, ldx, g, ldg, B, C, Cg, tilesC, part);
  VQ_LAUNCH_CHECK();
  const int n = 2 * (C + Cg);
  vq_moments_reduce_kernel<<<ceil_div(n, 128), 128, 0, s>>>(part, tilesR, n, sums);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

extern "C" int vqgnn_vq_whiten(const double* sums, double count, const double* d_count, int nb, int D, int Dg, int has_grad,
                               float* run_mean_f, float* run_var_f, float* run_mean_g, float* run_var_g,
                               float eps_f, float mom_f, float eps_g, float mom_g, float grad_scale0,
                               float grad_scale1, int training, int seed_running, const int32_t* seed_mask,
                               int64_t* nbt_f, int64_t* nbt_g, float* scale, float* shift, void* stream) {
  VQ_CHECK_ARG(nb > 0 && D > 0 && scale && shift && run_mean_f && run_var_f, "vq_whiten: bad arguments");
  VQ_CHECK_ARG(!has_grad || (run_mean_g && run_var_g && (Dg == D || Dg == D + 1)), "vq_whiten: bad grad args");
  VQ_CHECK_ARG(!(training || seed_running || seed_mask) || sums, "vq_whiten: training / seeding needs moments");
  const int Ctot = nb * D + (has_grad ? nb * Dg : 0);
  vq_whiten_kernel<<<ceil_div(Ctot, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, count, d_count, nb, D, Dg, has_grad, run_mean_f, run_var_f, run_mean_g, run_var_g, eps_f, mom_f, eps_g,
      mom_g, grad_scale0, grad_scale1, training, seed_running, seed_mask, nbt_f, nbt_g, scale, shift);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

namespace vqgnn {
size_t assign_tc_workspace_bytes(int nb, int M);
int launch_assign_tc(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* scale,
                     const float* shift, const float* E, int64_t B, int nb, int M, int D, int Dg, int Wp,
                     const int32_t* batch_idx, int16_t* codes, int64_t codes_ld, int16_t* idx, float* stats,
                     void* ws, size_t ws_bytes, cudaStream_t s);

template <int NV>
static int launch_assign_simt(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* scale,
                              const float* shift, const float* E, int64_t B, int nb, int M, int D, int Dg,
                              int Wp, const int32_t* batch_idx, int16_t* codes, int64_t codes_ld, int16_t* idx,
                              float* stats, cudaStream_t s) {
  const size_t smem = (sizeof(float4) * NV + sizeof(float)) * kChunk;
  // rows per thread: as many as keep >= 2 waves of CTAs in flight
  auto ctas = [&](int rpt) { return static_cast<int64_t>(ceil_div(B, kAssignThreads * rpt)) * nb; };
  int rpt = 4;
  while (rpt > 1 && ctas(rpt) < 4 * kNumSMs) rpt >>= 1;
#define VQ_ASSIGN_LAUNCH(RPT)                                                                               \
  do {                                                                                                      \
    auto kern = vq_assign_simt_kernel<NV, RPT>;                                                             \
    VQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    dim3 grid(ceil_div(B, kAssignThreads * RPT), nb);                                                       \
    kern<<<grid, kAssignThreads, smem, s>>>(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, Dg, Wp, batch_idx, \
                                            codes, codes_ld, idx, stats);                                            \
  } while (0)
  if (rpt == 4) VQ_ASSIGN_LAUNCH(4);
  else if (rpt == 2) VQ_ASSIGN_LAUNCH(2);
  else VQ_ASSIGN_LAUNCH(1);
#undef VQ_ASSIGN_LAUNCH
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
}  // namespace vqgnn

extern "C" size_t vqgnn_vq_assign_workspace_bytes(int nb, int M) { return assign_tc_workspace_bytes(nb, M); }

extern "C" int vqgnn_vq_assign(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* scale,
                               const float* shift, const float* E, int64_t B, int nb, int M, int D, int Dg,
                               int Wp, const int32_t* batch_idx, int16_t* codes, int64_t codes_ld, int16_t* idx,
                               float* stats, int impl, void* ws, size_t ws_bytes, void* stream) {
  VQ_CHECK_ARG(x && scale && shift && E && B > 0 && nb > 0 && M > 0 && D > 0, "vq_assign: bad arguments");
  VQ_CHECK_ARG(M <= 32767, "vq_assign: codes are int16, M must be <= 32767 (got %d)", M);
  VQ_CHECK_ARG(!codes || batch_idx, "vq_assign: codes scatter needs batch_idx");
  VQ_CHECK_ARG(!g || Dg == D || Dg == D + 1, "vq_assign: Dg must be D or D+1");
  VQ_CHECK_ARG(Wp % 4 == 0 && Wp >= D + (g ? Dg : 0), "vq_assign: Wp must be a multiple of 4 covering W");
  VQ_CHECK_ARG(nb <= 65535, "vq_assign: too many branches");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 1)
    return launch_assign_tc(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, Dg, Wp, batch_idx, codes, codes_ld, idx,
                            stats, ws, ws_bytes, s);
  const int w_use = D + (g ? Dg : 0);
  const int nv = (w_use + 3) / 4;
  switch (nv) {
    case 1: return launch_assign_simt<1>(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, Dg, Wp, batch_idx, codes, codes_ld, idx, stats, s);
    case 2: return launch_assign_simt<2>(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, Dg, Wp, batch_idx, codes, codes_ld, idx, stats, s);
    case 3: return launch_assign_simt<3>(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, Dg, Wp, batch_idx, codes, codes_ld, idx, stats, s);
    case 4: return launch_assign_simt<4>(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, Dg, Wp, batch_idx, codes, codes_ld, idx, stats, s);
    case 5: return launch_assign_simt<5>(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, Dg, Wp, batch_idx, codes, codes_ld, idx, stats, s);
    default: break;
  }
  set_error("vq_assign: joint width %d > 20 is not supported", w_use);
  return VQGNN_ERR_ARG;
}

extern "C" int vqgnn_vq_finalize(const float* stats, int nb, int M, int D, int Dg, int Wp, int joint, double decay,
                                 int warm_up, float eps, float grad_scale0, float grad_scale1,
                                 const float* run_mean_f, const float* run_var_f, const float* run_mean_g,
                                 const float* run_var_g, float* ema_size, float* ema_w, float* E, float* O,
                                 int32_t* status, void* stream) {
  VQ_CHECK_ARG(stats && ema_size && ema_w && E && O && status && run_mean_f && run_var_f, "vq_finalize: bad arguments");
  VQ_CHECK_ARG(!joint || (run_mean_g && run_var_g), "vq_finalize: joint update needs gradient statistics");
  // python: `size * decay + (1 - decay) * counts` with decay a double scalar cast to fp32 per operand
  const float decay_f = static_cast<float>(decay);
  const float omd = static_cast<float>(1.0 - decay);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  vq_finalize_size_kernel<<<nb, kFinThreads, 0, s>>>(stats, M, Wp, decay_f, omd, warm_up, ema_size, status);
  VQ_LAUNCH_CHECK();
  dim3 grid(ceil_div(M, kFinTile), nb);
  vq_finalize_code_kernel<<<grid, 256, 0, s>>>(stats, M, D, joint ? Dg : 0, Wp, joint, decay_f, omd, eps, grad_scale0,
                                               grad_scale1, run_mean_f, run_var_f, run_mean_g, run_var_g, ema_size,
                                               ema_w, E, O);
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}
