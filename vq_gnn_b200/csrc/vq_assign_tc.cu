// Codeword assignment on the 5th-gen tensor cores: tcgen05.mma (kind::tf32) with TMEM accumulators, the
// codebook tiles fed by TMA bulk copies, and a fused distance + argmin epilogue read back with tcgen05.ld.
//
// Per branch k the reference computes argmin_m (||z||^2 + ||e_m||^2) - 2 z.e_m in fp32 (vq.py:166-171,
// 230-236) with K = W = 4 / 8 / 9.  ||z||^2 does not change the argmin, so the kernel evaluates
//       d'[b, m] = ||e_m||^2 - 2 z_b . e_m
// as ONE K = 32 contraction in error-compensated 3xTF32:  z = zh + zl, e = eh + el (each TF32-exact),
//       A'[b] = [ zh | zl | zh | 1 | 1 | 0.. ]           (3W + 2 <= 32 columns)
//       B'[m] = [-2eh |-2eh |-2el | ch | cl | 0.. ]      c = ||e_m||^2 = ch + cl
// so d' = A'.B' up to the dropped zl.el term and TF32 rounding of the low parts (~2^-21 relative), which is
// what keeps the argmin on the fp32 answer except at near-ties (tests report the mismatch rate).
//
// Persistent CTAs (one per SM, all 512 TMEM columns = two 128x256 fp32 accumulators).  Work item =
// (branch, 128-row tile); per item the 16 epilogue warps whiten their rows and write the A' tile into shared
// memory in the canonical K-major UMMA layout, then for every 256-codeword tile of the branch:
//   warp 16 (one lane) TMA producer : cp.async.bulk of the pre-packed 32 KB B' tile into a 3-stage ring
//   warp 17 (one lane) MMA issuer   : 4 x tcgen05.mma (128x256x8, tf32) into accumulator buffer (tile & 1),
//                                     tcgen05.commit -> frees the smem stage and publishes the accumulator
//   warps 0-15         epilogue     : 4 warps per TMEM lane quarter, 64 columns each: two tcgen05.ld in flight,
//                                     buffer released as soon as they land, running (min, argmin) per row
// and finally the code is written (idx, code table scatter) and z is added to the per-codeword sums/counts.
#include "common.cuh"

namespace vqgnn {
namespace tc {

constexpr int kTileM = 128;           // rows per work item (= TMEM lanes)
constexpr int kTileN = 256;           // codewords per MMA tile (= TMEM columns per accumulator)
constexpr int kK = 32;                // packed contraction length (tf32 elements)
constexpr int kStages = 4;            // B' smem stages: a ring when streaming, the whole branch when M <= 1024
constexpr int kEpiWarps = 16;         // 4 warps per TMEM lane quarter, 64 accumulator columns each
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 64;
constexpr int kATileBytes = kTileM * kK * 4;   // 16 KB
constexpr int kBTileBytes = kTileN * kK * 4;   // 32 KB
constexpr int kALbo = (kTileM / 8) * 128;      // bytes between the two 16 B K-chunks of one MMA (A)
constexpr int kBLbo = (kTileN / 8) * 128;      // same for B
constexpr int kSbo = 128;                      // bytes between 8-row core matrices
constexpr size_t kSmemBytes = 1024 + 2 * kATileBytes + kStages * kBTileBytes + 4096;  // 165 KB: one CTA per SM
constexpr int kCand = 8;              // the scan tracks the best chunk of kCand codewords; vq_assign_refine_kernel picks inside it

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, no swizzle: core matrix = 8 rows x 16 B; LBO = byte step between K chunks, SBO = between row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);  // version = 1 (Blackwell), layout = SWIZZLE_NONE
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTileN >> 3) << 17) | ((kTileM >> 4) << 24);

__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ void tmem_ld64_issue(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// B' packer: [nb][M_pad/256][K chunk 8][row group 32][row 8][4] fp32 (TF32-exact values), i.e. every
// 256-codeword tile is a 32 KB block already in the shared-memory layout the MMA descriptor expects.
// ------------------------------------------------------------------------------------------------
__global__ void pack_codebook_kernel(const float* __restrict__ E, int nb, int M, int M_pad, int Wp, int w_use,
                                     float* __restrict__ Bp) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(nb) * M_pad) return;
  const int k = static_cast<int>(i / M_pad), m = static_cast<int>(i - static_cast<int64_t>(k) * M_pad);
  float b[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) b[j] = 0.f;
  if (m < M) {
    const float* e = E + (static_cast<int64_t>(k) * M + m) * Wp;
    float c = 0.f;
    for (int w = 0; w < w_use; ++w) {
      const float v = __ldg(e + w);
      c = fmaf(v, v, c);
      const float hi = to_tf32(v), lo = to_tf32(v - hi);
      b[w] = -2.f * hi, b[w_use + w] = -2.f * hi, b[2 * w_use + w] = -2.f * lo;
    }
    const float ch = to_tf32(c);
    b[3 * w_use] = ch, b[3 * w_use + 1] = to_tf32(c - ch);
  } else {
    b[3 * w_use] = 3.0e38f;  // padding codeword: never the minimum
  }
  const int tile = m / kTileN, ml = m - tile * kTileN;
  char* base = reinterpret_cast<char*>(Bp) + (static_cast<int64_t>(k) * (M_pad / kTileN) + tile) * kBTileBytes +
               (ml >> 3) * 128 + (ml & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < kK / 4; ++kc)
    *reinterpret_cast<float4*>(base + kc * kBLbo) = make_float4(b[4 * kc], b[4 * kc + 1], b[4 * kc + 2], b[4 * kc + 3]);
}

// ------------------------------------------------------------------------------------------------
// the assignment kernel
// ------------------------------------------------------------------------------------------------
template <int W, bool VLD>  // W: packed width D + Dg (4: feature only, 8: joint, 9: joint + add_flag); VLD: 128-bit row loads
__global__ void __launch_bounds__(kThreads, 1)
    vq_assign_tc_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g, int64_t ldg,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ Bp, int64_t B, int nb, int M, int M_pad, int D, int Dg, int Wp,
                        const int32_t* __restrict__ batch_idx, int16_t* __restrict__ codes, int64_t codes_ld,
                        int16_t* __restrict__ idx, float* __restrict__ stats) {
  static_assert(3 * W + 2 <= kK, "packed contraction does not fit");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_tile = smem;                         // two A' buffers: item i+1 is staged while item i drains
  unsigned char* b_tiles = smem + 2 * kATileBytes;
  unsigned char* misc = b_tiles + kStages * kBTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);  // [0..2] b_full, [3..5] b_empty, [6..7] acc_full, [8..9] acc_empty, [10] a_ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);
  float* xbest = reinterpret_cast<float*>(misc + 256);           // [3][128] candidates of column quarters 1..3
  int* xidx = reinterpret_cast<int*>(misc + 256 + 3 * 512);      // [3][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_FULL = 0, B_EMPTY = kStages, ACC_FULL = 2 * kStages, ACC_EMPTY = 2 * kStages + 2,
                A_READY = 2 * kStages + 4;  // two barriers, one per A' buffer

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(BAR(B_FULL + s), 1), mbar_init(BAR(B_EMPTY + s), 1);
    for (int b = 0; b < 2; ++b) mbar_init(BAR(ACC_FULL + b), 1), mbar_init(BAR(ACC_EMPTY + b), kEpiWarps);
    mbar_init(BAR(A_READY), 1), mbar_init(BAR(A_READY + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + 1) {  // TMEM: all 512 columns (two 128 x 256 fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_item = M_pad / kTileN;
  const int row_tiles = static_cast<int>((B + kTileM - 1) / kTileM);
  const int64_t n_items = static_cast<int64_t>(nb) * row_tiles;
  const int C = nb * D;
  // every CTA owns a CONTIGUOUS range of (branch-major) items, so consecutive items share the branch; when the
  // branch's whole packed codebook fits the stages (M <= 1024) it stays resident in shared memory and is fetched
  // once per branch instead of once per item (the per-item refetch made the first version L2-bound)
  const int64_t ipc = (n_items + gridDim.x - 1) / gridDim.x;
  const int64_t item_begin = blockIdx.x * ipc, item_end = min(n_items, item_begin + ipc);
  const bool resident = tiles_per_item <= kStages;

  if (warp == kEpiWarps) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t gt = 0, it = 0;
      int loaded = -1;
      for (int64_t item = item_begin; item < item_end; ++item, ++it) {
        const int k = static_cast<int>(item / row_tiles);
        const char* src = reinterpret_cast<const char*>(Bp) + static_cast<int64_t>(k) * tiles_per_item * kBTileBytes;
        for (int j = 0; j < tiles_per_item; ++j, ++gt) {
          const int s = resident ? j : gt % kStages;
          const uint32_t use = resident ? it : gt / kStages;
          mbar_wait(BAR(B_EMPTY + s), (use & 1) ^ 1);
          if (resident && k == loaded) {
            mbar_arrive(BAR(B_FULL + s));  // tile already in shared memory: just pass the token
          } else {
            mbar_expect_tx(BAR(B_FULL + s), kBTileBytes);
            tma_bulk_load(smem_u32(b_tiles + s * kBTileBytes), src + static_cast<int64_t>(j) * kBTileBytes, kBTileBytes,
                          BAR(B_FULL + s));
          }
        }
        loaded = k;
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t gt = 0, it = 0;
      const uint32_t a_addr0 = smem_u32(a_tile);
      for (int64_t item = item_begin; item < item_end; ++item, ++it) {
        const uint32_t a_addr = a_addr0 + (it & 1) * kATileBytes;
        mbar_wait(BAR(A_READY + (it & 1)), (it >> 1) & 1);
        for (int j = 0; j < tiles_per_item; ++j, ++gt) {
          const int s = resident ? j : gt % kStages, buf = gt & 1;
          const uint32_t use = resident ? it : gt / kStages;
          mbar_wait(BAR(ACC_EMPTY + buf), ((gt >> 1) & 1) ^ 1);
          mbar_wait(BAR(B_FULL + s), use & 1);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_tiles + s * kBTileBytes);
#pragma unroll
          for (int i = 0; i < kK / 8; ++i) {
            umma_tf32(tmem_base + buf * kTileN, smem_desc(a_addr + i * 2 * kALbo, kALbo, kSbo),
                      smem_desc(b_addr + i * 2 * kBLbo, kBLbo, kSbo), kIdesc, i > 0);
          }
          umma_commit(BAR(B_EMPTY + s));     // smem stage free once these MMAs have read it
          umma_commit(BAR(ACC_FULL + buf));  // accumulator complete
        }
      }
    }
    __syncwarp();
  } else {
    // ===== A' builders + epilogue (warps 0..15) =====
    const int q = warp & 3, h = warp >> 2;  // TMEM lane quarter, column quarter (64 of the tile's 256)
    const int rl = 32 * q + lane;  // row inside the tile == TMEM lane
    uint32_t gt = 0;
    // RAW row of an item for this thread's row (zeros past B / past the last item).  Nothing here may depend on the
    // loaded values: the loads are issued two items ahead and must stay in flight (an ncu source view of the first
    // pipelined version showed 40 % of the stall samples on the whitening FFMAs that consumed them immediately).
    auto load_raw = [&](int64_t item, float (&z)[W]) {
      const int k = static_cast<int>(item / row_tiles);
      const int64_t b = (item - static_cast<int64_t>(k) * row_tiles) * kTileM + rl;
#pragma unroll
      for (int w = 0; w < W; ++w) z[w] = 0.f;
      if (item >= item_end || b >= B || h != 0) return;   // the h == 0 warps own their rows: load, whiten, stage, emit
      if constexpr (VLD) {   // D == 4, 16 B aligned rows: one 128-bit load per operand instead of four scalar ones
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + b * ldx + k * 4));
        z[0] = xv.x, z[1] = xv.y, z[2] = xv.z, z[3] = xv.w;
        if constexpr (W == 8) {
          const float4 gv = __ldg(reinterpret_cast<const float4*>(g + b * ldg + k * 4));
          z[4] = gv.x, z[5] = gv.y, z[6] = gv.z, z[7] = gv.w;
        }
      } else {
#pragma unroll
        for (int w = 0; w < W; ++w)
          z[w] = (w < D) ? __ldg(x + b * ldx + k * D + w) : __ldg(g + b * ldg + k * Dg + (w - D));
      }
    };
    // z = raw * scale + shift of the item's branch (vq.py:223-227 folded into one affine per column)
    auto whiten = [&](int64_t item, float (&z)[W]) {
      const int k = static_cast<int>(item / row_tiles);
      const int64_t b = (item - static_cast<int64_t>(k) * row_tiles) * kTileM + rl;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const int c = (w < D) ? k * D + w : C + k * Dg + (w - D);
        z[w] = (item < item_end && b < B) ? fmaf(z[w], __ldg(scale + c), __ldg(shift + c)) : 0.f;
      }
    };
    // split into TF32 hi/lo, write this thread's quarter of the A' row into buffer (it & 1), publish it
    auto stage_a = [&](uint32_t it, const float (&z)[W]) {
      if (h == 0) {   // one thread per row writes the whole 128 B A' row (8 K chunks of 16 B)
        float a[kK];
#pragma unroll
        for (int j = 0; j < kK; ++j) a[j] = 0.f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const float hi = to_tf32(z[w]), lo = to_tf32(z[w] - hi);
          a[w] = hi, a[W + w] = lo, a[2 * W + w] = hi;
        }
        a[3 * W] = 1.f, a[3 * W + 1] = 1.f;
        unsigned char* dst = a_tile + (it & 1) * kATileBytes + (rl >> 3) * 128 + (rl & 7) * 16;
#pragma unroll
        for (int kk = 0; kk < kK / 4; ++kk)
          *reinterpret_cast<float4*>(dst + kk * kALbo) = make_float4(a[4 * kk], a[4 * kk + 1], a[4 * kk + 2], a[4 * kk + 3]);
        fence_async_smem();
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (threadIdx.x == 0) mbar_arrive(BAR(A_READY + (it & 1)));
    };
    // software pipeline over items: rows are loaded two items ahead, the A' tile is staged one item ahead (its
    // buffer was last read by the MMAs of item it-1, which this thread has already drained), so the MMAs of item
    // it+1 run while item it's accumulators are being read out of TMEM
    float z[W], z1[W], z2[W];   // z: whitened row of the current item; z1 / z2: raw rows of the next two
    load_raw(item_begin, z);
    load_raw(item_begin + 1, z1);
    whiten(item_begin, z);
    if (item_begin < item_end) stage_a(0, z);
    uint32_t it = 0;
    for (int64_t item = item_begin; item < item_end; ++item, ++it) {
      const int k = static_cast<int>(item / row_tiles);
      const int64_t b = (item - static_cast<int64_t>(k) * row_tiles) * kTileM + rl;
      load_raw(item + 2, z2);
      // the code-table row this thread will write at the end of the item: fetch its index now, not on the tail
      whiten(item + 1, z1);          // its loads were issued one whole item ago
      if (item + 1 < item_end) stage_a(it + 1, z1);

      // ---- scan the accumulator tiles ----
      float best = __int_as_float(0x7f800000);
      int besti = 0;
      for (int j = 0; j < tiles_per_item; ++j, ++gt) {
        const int buf = gt & 1;
        mbar_wait(BAR(ACC_FULL + buf), (gt >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + buf * kTileN + 64 * h;
        uint32_t rr[64];
        tmem_ld64_issue(taddr, rr);       // this warp's 32 lanes x 64 columns in one TMEM load
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));  // accumulator in registers: hand the buffer back early
                                                            // (one arrival per warp: 512 per-thread arrivals on one
                                                            // mbarrier were the top stall in the ncu source view)
        // Only the minimum of every kCand-column chunk and the chunk it came from are tracked here (FMNMX3 chain + one
        // compare / select pair, no branch).  Extracting the winning COLUMN inside the scan -- 31 FSETP/SEL pairs per
        // 32 columns whenever ANY lane of the warp improved, i.e. for ~85 % of the chunks -- made the ALU pipe the
        // co-bottleneck of this kernel (ncu: ALU 71 % busy; with the search stubbed out the kernel ran 30 % faster).
        // The column is picked inside the winning chunk by vq_assign_refine_kernel.
#pragma unroll
        for (int cchunk = 0; cchunk < 64 / kCand; ++cchunk) {
          float m = __uint_as_float(rr[kCand * cchunk]);
#pragma unroll
          for (int i = 1; i < kCand; ++i) m = fminf(m, __uint_as_float(rr[kCand * cchunk + i]));
          const bool better = m < best;   // strict: earlier (lower) chunks win ties
          best = better ? m : best;
          besti = better ? j * kTileN + 64 * h + kCand * cchunk : besti;   // first column of the chunk
        }
      }
      // ---- combine the two column halves of every row, emit code + statistics ----
      if (h > 0) xbest[(h - 1) * kTileM + rl] = best, xidx[(h - 1) * kTileM + rl] = besti;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (h == 0 && b < B) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const float ob = xbest[o * kTileM + rl];
          const int oi = xidx[o * kTileM + rl];
          if (ob < best || (ob == best && oi < besti)) best = ob, besti = oi;
        }
        idx[b * nb + k] = static_cast<int16_t>(besti);   // first codeword of the winning chunk (refined below)
      }
#pragma unroll
      for (int w = 0; w < W; ++w) z[w] = z1[w], z1[w] = z2[w];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// Second step of the tcgen05 assignment: inside the winning chunk of kCand codewords the column is picked by re-scoring
// the chunk in fp32 from the codebook itself (d = ||e||^2 - 2 z.e; ||z||^2 is common to the row, z whitened with the
// same fmaf as everywhere else), lowest index winning ties.  Against the tensor-core values this can only move the choice
// between codewords whose distances agree to ~1e-6 relative -- well inside the near-tie band of the parity tests
// (1e-5).  One thread per (row, branch): 2 x 16 B of the row + kCand x 32 B of the L2-resident codebook; also scatters
// the code into the code table and, when asked, accumulates the (non-deterministic) float-atomic statistics.
template <int W>
__global__ void __launch_bounds__(256)
    vq_assign_refine_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g, int64_t ldg,
                            const float* __restrict__ scale, const float* __restrict__ shift,
                            const float* __restrict__ E, int64_t B, int nb, int M, int D, int Dg, int Wp,
                            const int32_t* __restrict__ batch_idx, int16_t* __restrict__ codes, int64_t codes_ld,
                            int16_t* __restrict__ idx, float* __restrict__ stats) {
  constexpr int NV = (W + 3) / 4;
  const int64_t n = B * nb;
  const int C = nb * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t b = i / nb;
    const int k = static_cast<int>(i - b * nb);
    float z[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int c = (w < D) ? k * D + w : C + k * Dg + (w - D);
      const float raw = (w < D) ? __ldg(x + b * ldx + k * D + w) : __ldg(g + b * ldg + k * Dg + (w - D));
      z[w] = fmaf(raw, __ldg(scale + c), __ldg(shift + c));
    }
    const int base = idx[i];
    const float* e0 = E + (static_cast<int64_t>(k) * M + base) * Wp;
    int code = base;
    float dmin = __int_as_float(0x7f800000);
#pragma unroll
    for (int cI = 0; cI < kCand; ++cI) {
      if (base + cI < M) {
        float ev[NV * 4];
#pragma unroll
        for (int w4 = 0; w4 < NV; ++w4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(e0 + static_cast<int64_t>(cI) * Wp) + w4);
          ev[4 * w4] = t.x, ev[4 * w4 + 1] = t.y, ev[4 * w4 + 2] = t.z, ev[4 * w4 + 3] = t.w;
        }
        float c2 = 0.f, dot = 0.f;
#pragma unroll
        for (int w = 0; w < W; ++w) c2 = fmaf(ev[w], ev[w], c2), dot = fmaf(z[w], ev[w], dot);
        const float d = fmaf(-2.f, dot, c2);
        if (d < dmin) dmin = d, code = base + cI;
      }
    }
    idx[i] = static_cast<int16_t>(code);
    if (codes) codes[static_cast<int64_t>(__ldg(batch_idx + b)) * codes_ld + k] = static_cast<int16_t>(code);
    if (stats) {
      float* dst = stats + (static_cast<int64_t>(k) * M + code) * (Wp + 4);
#pragma unroll
      for (int w4 = 0; w4 < NV; ++w4) {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = (4 * w4 + j < W) ? z[4 * w4 + j] : 0.f;
        atomicAdd(reinterpret_cast<float4*>(dst) + w4, make_float4(t[0], t[1], t[2], t[3]));
      }
      atomicAdd(dst + Wp, 1.0f);
    }
  }
}

}  // namespace tc

size_t assign_tc_workspace_bytes(int nb, int M) {
  const int M_pad = (M + tc::kTileN - 1) / tc::kTileN * tc::kTileN;
  return static_cast<size_t>(nb) * M_pad * tc::kK * sizeof(float);
}

int launch_assign_tc(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* scale,
                     const float* shift, const float* E, int64_t B, int nb, int M, int D, int Dg, int Wp,
                     const int32_t* batch_idx, int16_t* codes, int64_t codes_ld, int16_t* idx, float* stats,
                     void* ws, size_t ws_bytes, cudaStream_t s) {
  using namespace tc;
  const int W = D + (g ? Dg : 0);
  VQ_CHECK_ARG(W == 4 || W == 8 || W == 9, "vq_assign: the tcgen05 path supports packed widths 4, 8, 9 (got %d)", W);
  VQ_CHECK_ARG(ws && ws_bytes >= assign_tc_workspace_bytes(nb, M), "vq_assign: tcgen05 path needs a workspace of "
               "vqgnn_vq_assign_workspace_bytes(nb, M) bytes");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "vq_assign: workspace must be 16 B aligned");
  VQ_CHECK_ARG(idx, "vq_assign: the tcgen05 path needs idx (it carries the winning chunk between its two kernels)");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(E) & 15) == 0 && Wp % 4 == 0, "vq_assign: codebook must be 16 B aligned");
  const int M_pad = (M + kTileN - 1) / kTileN * kTileN;
  float* Bp = static_cast<float*>(ws);
  const int64_t n_pack = static_cast<int64_t>(nb) * M_pad;
  pack_codebook_kernel<<<ceil_div(n_pack, 256), 256, 0, s>>>(E, nb, M, M_pad, Wp, W, Bp);
  VQ_LAUNCH_CHECK();
  const int64_t n_items = static_cast<int64_t>(nb) * ((B + kTileM - 1) / kTileM);
  const int grid = static_cast<int>(std::min<int64_t>(n_items, kNumSMs));
  const bool vld = (D == 4) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   (!g || (Dg == 4 && ldg % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0));
#define VQ_TC_LAUNCH(WW, VV)                                                                               \
  do {                                                                                                     \
    auto kern = vq_assign_tc_kernel<WW, VV>;                                                                \
    VQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));     \
    kern<<<grid, kThreads, kSmemBytes, s>>>(x, ldx, g, ldg, scale, shift, Bp, B, nb, M, M_pad, D, Dg, Wp,  \
                                            batch_idx, codes, codes_ld, idx, stats);                       \
  } while (0)
  if (W == 4) {
    if (vld) VQ_TC_LAUNCH(4, true);
    else VQ_TC_LAUNCH(4, false);
  } else if (W == 8) {
    if (vld) VQ_TC_LAUNCH(8, true);
    else VQ_TC_LAUNCH(8, false);
  } else {
    VQ_TC_LAUNCH(9, false);
  }
#undef VQ_TC_LAUNCH
  VQ_LAUNCH_CHECK();
  const int rgrid = static_cast<int>(std::min<int64_t>((B * nb + 255) / 256, 32 * kNumSMs));
#define VQ_TC_REFINE(WW)                                                                                        \
  vq_assign_refine_kernel<WW><<<rgrid, 256, 0, s>>>(x, ldx, g, ldg, scale, shift, E, B, nb, M, D, g ? Dg : 0, Wp, \
                                                    batch_idx, codes, codes_ld, idx, stats)
  if (W == 4) VQ_TC_REFINE(4);
  else if (W == 8) VQ_TC_REFINE(8);
  else VQ_TC_REFINE(9);
#undef VQ_TC_REFINE
  VQ_LAUNCH_CHECK();
  return VQGNN_OK;
}

}  // namespace vqgnn
