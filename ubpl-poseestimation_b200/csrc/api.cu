// Library plumbing: error text, device info, and the libm pow(v, 0.5) exception table.
#include "common.cuh"
#include <stdarg.h>
#include <math.h>
#include <mutex>
#include <vector>

namespace ubpl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UBPL_ERR_CUDA;
  }
  return UBPL_OK;
}

struct DevInfo {
  int dev = -1, sms = 0, major = 0, minor = 0, smem = 0;
};
static DevInfo g_dev;
static std::mutex g_mu;

static const DevInfo& dev_info() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev.dev != dev) {
    cudaDeviceGetAttribute(&g_dev.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_dev.major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&g_dev.minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&g_dev.smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_dev.dev = dev;
  }
  return g_dev;
}
int sm_count() { return dev_info().sms > 0 ? dev_info().sms : 1; }
int smem_optin() { return dev_info().smem; }

// ---------------------------------------------------------------------------------------------
// CPython evaluates the reference's `(dx**2 + dy**2) ** 0.5` (utils/process.py:53-54) with libm
// pow(v, 0.5), which is NOT correctly rounded: for ~0.1 % of arguments it is one ulp away from
// sqrt(v).  Decoded coordinates are integers, so the radicands on the selection path are
// integers below a small bound; we tabulate (once, on the host, with the SAME libm the python
// reference would call here) the radicands where pow and sqrt disagree.  The device computes
// the IEEE sqrt and patches those cases, which makes the float64 distances -- and therefore the
// quantile threshold and the pseudo-label masks -- bit-identical to the reference.
// ---------------------------------------------------------------------------------------------
static double (*volatile g_pow)(double, double) = pow;   // volatile: no folding into sqrt
static std::vector<int32_t> g_exc_key;
static std::vector<double> g_exc_val;
static int32_t* d_exc_key = nullptr;
static uint32_t* d_exc_bits = nullptr;   // bitmap over radicands: 1 = pow(v,0.5) != sqrt(v)
static double* d_exc_val = nullptr;
static int g_exc_n = 0, g_exc_dev = -1;
static const int kPowTableMax = 1 << 21;   // radicands up to 2*1024^2: coordinates up to +-1024 px

int pow_table(const int32_t** keys, const double** vals, const uint32_t** bits, int* n, int* rmax) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_exc_key.empty() && g_exc_n == 0) {
    for (int r = 0; r <= kPowTableMax; ++r) {
      const double v = (double)r;
      const double pw = g_pow(v, 0.5);
      if (pw != sqrt(v)) {
        g_exc_key.push_back(r);
        g_exc_val.push_back(pw);
      }
    }
    g_exc_n = (int)g_exc_key.size();
  }
  if (g_exc_dev != dev) {
    // one table per process/device; a process drives one GPU (torch.distributed launch model)
    const size_t nn = g_exc_n > 0 ? (size_t)g_exc_n : 1;
    if (cudaMalloc(&d_exc_key, nn * sizeof(int32_t)) != cudaSuccess ||
        cudaMalloc(&d_exc_val, nn * sizeof(double)) != cudaSuccess) {
      set_error("pow_table: cudaMalloc failed");
      return UBPL_ERR_CUDA;
    }
    if (g_exc_n > 0) {
      cudaMemcpy(d_exc_key, g_exc_key.data(), g_exc_n * sizeof(int32_t), cudaMemcpyHostToDevice);
      cudaMemcpy(d_exc_val, g_exc_val.data(), g_exc_n * sizeof(double), cudaMemcpyHostToDevice);
    }
    const size_t words = (size_t)kPowTableMax / 32 + 1;
    std::vector<uint32_t> bm(words, 0u);
    for (int r : g_exc_key) bm[(size_t)r >> 5] |= 1u << (r & 31);
    if (cudaMalloc(&d_exc_bits, words * sizeof(uint32_t)) != cudaSuccess) {
      set_error("pow_table: cudaMalloc failed");
      return UBPL_ERR_CUDA;
    }
    cudaMemcpy(d_exc_bits, bm.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice);
    g_exc_dev = dev;
  }
  *keys = d_exc_key;
  *vals = d_exc_val;
  *bits = d_exc_bits;
  *n = g_exc_n;
  *rmax = kPowTableMax;
  return UBPL_OK;
}

// A ring of device counters for kernels that distribute work dynamically; each launch takes the
// next slot and zeroes it on its stream (so concurrent launches on different streams, and CUDA
// graph replays of one captured launch, never share a live counter).
static unsigned long long* d_counters = nullptr;
static int g_counter_dev = -1;
static unsigned g_counter_next = 0;
static const unsigned kCounters = 256;

unsigned long long* work_counter(cudaStream_t stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_counter_dev != dev) {
    if (cudaMalloc(&d_counters, kCounters * sizeof(unsigned long long)) != cudaSuccess) {
      set_error("work_counter: cudaMalloc failed");
      return nullptr;
    }
    g_counter_dev = dev;
  }
  unsigned long long* c = d_counters + (g_counter_next++ % kCounters);
  if (cudaMemsetAsync(c, 0, sizeof(unsigned long long), stream) != cudaSuccess) {
    set_error("work_counter: cudaMemsetAsync failed");
    return nullptr;
  }
  return c;
}

}  // namespace ubpl

extern "C" const char* ubpl_last_error(void) { return ubpl::g_err; }
extern "C" int ubpl_version(void) { return 202; }   // round 2, second ABI addition (ubpl_warp_decode_k2_ema)
extern "C" int ubpl_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    ubpl::set_error("ubpl_device_info: no CUDA device");
    return UBPL_ERR_CUDA;
  }
  const ubpl::DevInfo& d = ubpl::dev_info();
  if (sm_count) *sm_count = d.sms;
  if (cc_major) *cc_major = d.major;
  if (cc_minor) *cc_minor = d.minor;
  if (smem_optin_bytes) *smem_optin_bytes = d.smem;
  return UBPL_OK;
}
