// Shared host-side helpers of libvbsky_b200: error reporting and handle definitions.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "vbsky_b200.h"

namespace vbsky {

int set_error(int code, const char* fmt, ...);

#define VB_CHECK_ARG(cond, ...)                                      \
  do {                                                               \
    if (!(cond)) return ::vbsky::set_error(VBSKY_EINVAL, __VA_ARGS__); \
  } while (0)

#define VB_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return ::vbsky::set_error(VBSKY_ECUDA, "%s failed: %s (%s:%d)", #expr,           \
                                cudaGetErrorString(_e), __FILE__, __LINE__);            \
  } while (0)

// Handles live on the device that was current when they were created; every compute entry point requires it current.
#define VB_CHECK_DEVICE(lay)                                                                         \
  do {                                                                                               \
    VB_CHECK_ARG((lay)->device >= 0, "layout was created without a CUDA device; there is no CPU fallback"); \
    int _cur = -1;                                                                                   \
    VB_CUDA(cudaGetDevice(&_cur));                                                                   \
    VB_CHECK_ARG(_cur == (lay)->device, "current CUDA device %d is not the handle's device %d", _cur, (lay)->device); \
  } while (0)

// One step of the up (postorder) pass: combine children a and b into parent p.
//   x = a | b<<16, y = p, z = slotA | slotB<<8 | slotOut<<16 (shared-memory slots; 0xff = none),
//   w = flags: bit0 a is a tip, bit1 b is a tip, bit2 p is the root, bit3 a is the result of the previous
//   step (still in registers; the children are ordered so that only a ever is), bit5 the result must
//   also be parked in slotOut, bit6 p is a cherry (both children tips, not the root): its message is NOT
//   streamed to scratch.  The message produced by any other step s is streamed to scratch row s.
// One step of the down (preorder) pass: expand parent p into children a and b.
//   x = a | b<<16, y = p, z = slotP | 0xff<<8 | slotB<<16, w = flags: bit0/bit1 tips as above,
//   bit3 pre_p is carried in registers from the previous step -- for the root step, from the seed pi -- (else slotP).
//   The children are ordered so that an internal a is expanded by the next step (its pre vector is
//   carried) and an internal b is parked in slotB until the walk returns to it.
//   Internal children's messages arrive through the message ring in the order a, b (h_msg_rows), except cherries:
//   bit6 a is a cherry -- its message is rebuilt from the next record (a's own step), bit7 b is a cherry (then a is
//   one too) -- rebuilt from the record after next; the chunking keeps those records in the same ring stage.
struct Op {
  int32_t x, y, z, w;
};

// Geometry shared by the host-side layout builder and the pruning kernel.
constexpr int kTile = 256;        // site patterns per CTA tile (two per consumer thread)
constexpr int kChunkSteps = 3;    // schedule steps delivered per ring stage

}  // namespace vbsky

struct vbsky_layout {
  int32_t T = 0, n = 0, N = 0;
  int32_t depth_up = 0, depth_down = 0, max_levels = 0;
  // host copies
  std::vector<vbsky::Op> h_up, h_down;        // [T][n-1]
  std::vector<int32_t> h_down_nodes;          // [T][2n-2]
  std::vector<int32_t> h_msg_rows;            // [T][n-2] scratch rows of internal children in down-pass consumption order
  std::vector<int32_t> h_msg_start;           // [T][chunk_cap+1] prefix of h_msg_rows per down chunk
  std::vector<int32_t> h_dchunk_start;        // [T][chunk_cap+1] first step of each (variable-length, <= kChunkSteps) down chunk
  std::vector<int32_t> h_n_dchunks;           // [T] number of down chunks
  std::vector<int32_t> h_n_msgs;              // [T] messages that travel through scratch (non-root internal nodes minus cherries)
  std::vector<vbsky::Op> h_chunk_meta;        // [T][chunk_cap] {first step, steps, first message, end message} per down chunk
  int32_t chunk_cap = 0;
  std::vector<int32_t> h_child_parent;        // [T][N]
  std::vector<int32_t> h_parent_child;        // [T][N][2]
  std::vector<int32_t> h_siblings;            // [T][N]
  std::vector<double> h_lst;                  // [T][N]
  std::vector<double> h_sample_times;         // [T][n]
  std::vector<double> h_max_sample_time;      // [T]
  std::vector<int32_t> h_level_nodes;         // [T][n-1] internal nodes ordered by depth (root first)
  std::vector<int32_t> h_level_start;         // [T][max_levels+1]
  // device copies
  vbsky::Op* d_up = nullptr;
  vbsky::Op* d_down = nullptr;
  int32_t* d_down_nodes = nullptr;
  int32_t* d_msg_rows = nullptr;
  int32_t* d_msg_start = nullptr;
  int32_t* d_dchunk_start = nullptr;
  int32_t* d_n_dchunks = nullptr;
  int32_t* d_n_msgs = nullptr;
  vbsky::Op* d_chunk_meta = nullptr;
  int32_t* d_child_parent = nullptr;
  int32_t* d_parent_child = nullptr;
  double* d_lst = nullptr;
  double* d_sample_times = nullptr;
  double* d_max_sample_time = nullptr;
  int32_t* d_level_nodes = nullptr;
  int32_t* d_level_start = nullptr;
  int device = -1;
};

struct vbsky_tips {
  int32_t T = 0, n = 0;
  int64_t S = 0;        // patterns per tree as given
  int64_t S_pad = 0;    // padded to a multiple of kTile
  int32_t tiles = 0;    // S_pad / kTile
  // Code ranks (see tips.cu) in the order the two passes consume them, tile-major:
  // [T][tiles][2(n-1)][kTile]; row 2s / 2s+1 = child a / b of schedule step s (filler for internal children).
  uint8_t* d_codes_up = nullptr;
  uint8_t* d_codes_dn = nullptr;
  double* d_counts = nullptr;  // [T][S_pad], 0 on padding
  int device = -1;
};
