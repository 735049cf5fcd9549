// Interface between csrc/mlp.cu (layer kernels for any fully connected architecture) and
// csrc/mlp_umma.cu (the wide layers' forward / backward-data GEMMs on tcgen05).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sgmcmc {

// The activation operand of a tensor-core layer GEMM lives in the workspace already split into a hi
// and a lo plane and in the MMA's canonical K-major order, MU_CN = 32 minibatch rows (N) per unit k:
//   float index inside a plane = ((k / 4) * MU_CN + n) * 4 + k % 4,   plane = round_up(width, MU_KPAD) * MU_CN floats
// (rows n >= batch and units k >= width are zeros), so that a block of MU_KPAD = 32 k is one contiguous
// 4 KB piece per plane and its staging a plain copy.
constexpr int MU_CN = 32, MU_KPAD = 32;

struct MuArgs {
  const float* theta;      // [n_theta_rows, D]
  int64_t D;
  int theta_div;           // theta row of work item c = c / theta_div
  int64_t oW, ob;          // offsets of W_l and b_l inside a parameter row
  int ldw;                 // row length of W_l = its n_out
  int M, K, Mpad;          // units produced, contraction length, round_up(M, MU_KPAD)
  float* ws;               // [n_items, ws_floats]
  int64_t ws_floats;
  int64_t oBc, b_plane;    // canonical activation operand (hi plane; lo plane b_plane floats further)
  int64_t oHin;            // backward: plain H of the produced units [M][bt]
  int64_t oOut;            // plain result [M][bt]
  int64_t oOutc, out_plane;   // canonical result (-1: not needed)
  int64_t oSq;             // forward: partial sums of squares, one per 128-unit tile
  const int32_t* starts;   // minibatch starts (NULL: row tiles, see item_rows in mlp.cu)
  int64_t n_rows;
  int batch, bt;           // rows per item; row stride of the plain layouts
  int64_t chain0;          // first work item of this launch (set by launch_mlp_gemm_umma)
};

int launch_mlp_gemm_umma(const MuArgs& a, bool fwd, int64_t n_items, cudaStream_t st);

}  // namespace sgmcmc
