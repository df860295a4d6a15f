// FedAvg over the flat communicated adapter_1 buffer (reference get_average_net, main.py:50-65):
//   temp = 0; for c: temp += client_c * num_c / total; server = temp        (fp32, in client order)
// The arithmetic order (multiply, IEEE divide, add, clients in sequence) is kept so the result is
// bit-identical to the reference's PyTorch expression.  HBM-bound: (n_clients reads + 1 write) * 4 B
// per element, float4-vectorised grid-stride loop.
#include "feddat_b200.h"
#include "host_common.h"

namespace fd {
namespace {

constexpr int kMaxClients = 64;
struct FedAvgParams {
  const float* src[kMaxClients];
  float num[kMaxClients];
  float total;
  int n_clients;
  float* out;
  int64_t n;
};

__global__ void __launch_bounds__(256) fedavg_kernel(const __grid_constant__ FedAvgParams p) {
  const int64_t n4 = p.n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < p.n_clients; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.src[c]) + i);
      const float w = p.num[c];
      acc.x = __fadd_rn(acc.x, __fdiv_rn(__fmul_rn(v.x, w), p.total));
      acc.y = __fadd_rn(acc.y, __fdiv_rn(__fmul_rn(v.y, w), p.total));
      acc.z = __fadd_rn(acc.z, __fdiv_rn(__fmul_rn(v.z, w), p.total));
      acc.w = __fadd_rn(acc.w, __fdiv_rn(__fmul_rn(v.w, w), p.total));
    }
    reinterpret_cast<float4*>(p.out)[i] = acc;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n;
       i += stride) {
    float acc = 0.f;
    for (int c = 0; c < p.n_clients; ++c)
      acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(p.src[c][i], p.num[c]), p.total));
    p.out[i] = acc;
  }
}

}  // namespace
}  // namespace fd

extern "C" int feddat_fedavg(const float* const* clients, const float* weights, int n_clients,
                             float total_weight, float* out, int64_t n, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(clients && weights && out, FD_ERR_INVALID, "fedavg: null pointer argument");
  FD_REQUIRE(n_clients >= 1 && n_clients <= kMaxClients, FD_ERR_UNSUPPORTED,
             "fedavg: n_clients must be in [1, %d] (got %d)", kMaxClients, n_clients);
  FD_REQUIRE(n >= 0, FD_ERR_INVALID, "fedavg: negative length");
  if (n == 0) return FD_OK;
  FedAvgParams p{};
  float total = 0.f;
  for (int c = 0; c < n_clients; ++c) {
    FD_REQUIRE(clients[c] != nullptr, FD_ERR_INVALID, "fedavg: client %d buffer is null", c);
    FD_REQUIRE((reinterpret_cast<uintptr_t>(clients[c]) & 15) == 0, FD_ERR_INVALID,
               "fedavg: client %d buffer is not 16-byte aligned", c);
    p.src[c] = clients[c];
    p.num[c] = weights[c];
    total += weights[c];
  }
  FD_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, FD_ERR_INVALID,
             "fedavg: out is not 16-byte aligned");
  if (total_weight > 0.f) total = total_weight;  // partial sum of a rank: divide by the GLOBAL total
  FD_REQUIRE(total != 0.f, FD_ERR_INVALID, "fedavg: weights sum to zero");
  p.total = total; p.n_clients = n_clients; p.out = out; p.n = n;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  int64_t blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > static_cast<int64_t>(sms) * 8) blocks = static_cast<int64_t>(sms) * 8;
  fedavg_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
