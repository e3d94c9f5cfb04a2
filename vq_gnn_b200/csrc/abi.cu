// Library-level entry points of libvqgnn.so: version / arch gate / error text / small helpers.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace vqgnn {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

__global__ void fill_zero_kernel(uint4* p, size_t n16) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride)
    p[i] = make_uint4(0, 0, 0, 0);
}

__global__ void codes_pack_kernel(const int16_t* __restrict__ table, int k, int nb, int64_t N,
                                  int16_t* __restrict__ codes) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; n < N; n += stride)
    codes[n * nb + k] = table[n];
}
}  // namespace vqgnn

using namespace vqgnn;

extern "C" int vqgnn_abi_version(void) { return VQGNN_ABI_VERSION; }

extern "C" const char* vqgnn_last_error(void) { return g_err; }

extern "C" int64_t vqgnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int vqgnn_arch_check(int device) {
  cudaDeviceProp prop;
  VQ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libvqgnn is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor,
              prop.name);
    return VQGNN_ERR_ARCH;
  }
  return VQGNN_OK;
}

extern "C" int vqgnn_fill_zero(void* ptr, size_t bytes, void* stream) {
  VQ_CHECK_ARG(ptr || bytes == 0, "fill_zero: null pointer");
  if (bytes == 0) return VQGNN_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && bytes % 16 == 0) {
    const size_t n16 = bytes / 16;
    const int grid = static_cast<int>(std::min<size_t>((n16 + 255) / 256, 8 * kNumSMs));
    fill_zero_kernel<<<grid, 256, 0, s>>>(static_cast<uint4*>(ptr), n16);
    VQ_LAUNCH_CHECK();
  } else {
    VQ_CUDA(cudaMemsetAsync(ptr, 0, bytes, s));
  }
  return VQGNN_OK;
}

extern "C" int vqgnn_flush_l2(void* buf, size_t bytes, void* stream) { return vqgnn_fill_zero(buf, bytes, stream); }

extern "C" int vqgnn_codes_pack(const int16_t* const* h_tables, int nb, int64_t N, int16_t* codes, void* stream) {
  VQ_CHECK_ARG(h_tables && codes && nb > 0 && N > 0, "codes_pack: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(std::min<int64_t>((N + 255) / 256, 8 * kNumSMs));
  for (int k = 0; k < nb; ++k) {
    VQ_CHECK_ARG(h_tables[k], "codes_pack: table %d is null", k);
    codes_pack_kernel<<<grid, 256, 0, s>>>(h_tables[k], k, nb, N, codes);
    VQ_LAUNCH_CHECK();
  }
  return VQGNN_OK;
}
