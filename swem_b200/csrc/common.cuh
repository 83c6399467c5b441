// Shared host/device helpers for libswem_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>
#include <stdio.h>

#include "../../include/swem_b200.h"

namespace swem {

constexpr float kEpsNorm = 1e-6f;     // l2norm epsilon, reference modules.py:8
constexpr float kLog2e = 1.4426950408889634f;

// ---- error plumbing (thread-local message + launch counter) --------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void reset_launch_count();

#define SWEM_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::swem::set_error(__VA_ARGS__);             \
      return SWEM_ERR_INVALID_ARG;                \
    }                                             \
  } while (0)

#define SWEM_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::swem::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                        \
      return SWEM_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

#define SWEM_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      ::swem::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),     \
                        __FILE__, __LINE__);                                             \
      return SWEM_ERR_CUDA;                                                              \
    }                                                                                    \
    ::swem::count_launch();                                                              \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Per-device, thread-safe "done once" state for function attributes and occupancy queries: cudaFuncSetAttribute and
// cudaOccupancyMaxActiveClusters are per device, so a process that drives several GPUs must repeat them on each one.
constexpr int kMaxDevices = 64;
struct PerDevice {
  std::mutex mu;
  int value[kMaxDevices];
  bool done[kMaxDevices];
  PerDevice() {
    for (int i = 0; i < kMaxDevices; ++i) { value[i] = -1; done[i] = false; }
  }
};
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t off = 0;
  explicit Arena(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
};

// ---- entry points of the two kernel families -----------------------------------------------
size_t generic_em_workspace(const SwemDims& d);
int generic_em_forward(const SwemEmArgs& a, cudaStream_t st);
size_t generic_em_backward_workspace(const SwemDims& d);
int generic_em_backward(const SwemEmBwdArgs& a, cudaStream_t st);
size_t generic_readout_backward_workspace(const SwemDims& d);
int generic_readout_backward(const SwemReadBwdArgs& a, cudaStream_t st);
size_t generic_readout_workspace(const SwemDims& d);
int generic_readout_forward(const SwemReadArgs& a, cudaStream_t st);

void set_profile_buffer(void* dev);
long long* get_profile_buffer();
bool fused_em_supported(const SwemDims& d);
size_t fused_em_workspace(const SwemDims& d);
int fused_em_forward(const SwemEmArgs& a, cudaStream_t st);
bool fused_em_res_covers(const SwemDims& d, bool v_pixel_major);   // V-resident kernel (fused_em_res.cu): Ck = 64, L <= 128
int fused_em_res_forward(const SwemEmArgs& a, cudaStream_t st);
bool fused_em_res_emits_images(const SwemEmArgs& a);            // the kernel writes the readout's operand images of its output bases itself
bool fused_readout_supported(const SwemDims& d);
size_t fused_readout_workspace(const SwemDims& d);
int fused_readout_forward(const SwemReadArgs& a, cudaStream_t st);
struct ReadoutImages {        // tensor-core operand images of the memory banks inside a readout workspace (fused_readout.cu)
  uint8_t* kblob;
  uint8_t* vblob;
  size_t end_offset;
};
ReadoutImages readout_image_layout(void* workspace, const SwemDims& d);
int launch_bank_images(const SwemDims& d, const float* const kappa[2], const float* const nu[2], int bank0, int bank1,
                       const ReadoutImages& im, cudaStream_t st);
bool fused_readout_topl_covers(const SwemDims& d);             // readout with the in-kernel top-l feature (fused_readout_topl.cu): Ck = 64, Lt <= 256
int fused_readout_topl_launch(const SwemReadArgs& a, const uint8_t* kblob, const uint8_t* vblob, cudaStream_t st);

// shared between families: sorted top-l prefix feature from normalised attention rows
// P: [U, HW, 2*Lt] fp32 (row per pixel, column = side*Lt + j) -> out channels [s_channel, +2*topl)
int launch_perm_inv(const float* P, int U, int HW, int Lt, int topl, float* out, int out_channels,
                    int s_channel, int pixel_major, cudaStream_t st);

}  // namespace swem
