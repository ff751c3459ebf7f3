// head_train.cuh -- layouts shared by the fused training step of the head network (head_train_fwd.cu, head_train_bwd.cu):
// NeRFNetwork.forward (nerf/network.py:222-283) and its backward for one batch of march_rays_train samples, replacing the
// reference's ~8 cuBLAS GEMMs + ~150 elementwise / cat / repeat / cast launches per direction.
//
//   forward   head_train_fwd_kernel : head_eval's pipeline (3-D encode -> ambient MLP -> tanh -> 2-D encode -> sigma MLP -> exp / SH ->
//             colour MLP -> sigmoid on tcgen05) that additionally SAVES every layer input as the 128-row fp16 tile it already
//             built in shared memory (interleaved UMMA layout, 928 B/sample) and the 2-D grid's d(enc)/d(coordinate);
//   backward  head_train_bwd1_kernel (colour + sigma nets) and head_train_bwd2_kernel (ambient net): per 128-sample tile the saved
//             tiles come back by TMA bulk copies; the data-gradient chain dZ_{l-1} = (dZ_l W_l) * relu' runs as tcgen05.mma against
//             transposed weight blobs; the weight gradients dW_l = dZ_l^T A_{l-1} run as tcgen05.mma with BOTH operands read
//             MN-major from the very same tiles (samples = the MMA's K dimension) and accumulate in TMEM across all tiles of the
//             persistent CTA; one MMA-issuer warp serves all tile groups, so accumulation order into the shared accumulators is
//             program order.  Between them the existing grid kernels scatter the table gradients.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace rn {
namespace train {

// ---- saved activation tiles: byte offsets inside one 128-sample tile record -------------------------------------------------
constexpr uint32_t T_A0 = 0;                      // enc_x           [128 x 32]
constexpr uint32_t T_HA1 = T_A0 + 128 * 32 * 2;   // relu(ambient L1) [128 x 64]
constexpr uint32_t T_HA2 = T_HA1 + 128 * 64 * 2;  // relu(ambient L2) [128 x 64]
constexpr uint32_t T_EW = T_HA2 + 128 * 64 * 2;   // enc_w           [128 x 32]
constexpr uint32_t T_HS1 = T_EW + 128 * 32 * 2;   // relu(sigma L1)   [128 x 64]
constexpr uint32_t T_HS2 = T_HS1 + 128 * 64 * 2;  // relu(sigma L2)   [128 x 64]
constexpr uint32_t T_CIN = T_HS2 + 128 * 64 * 2;  // [sh | geo_feat]  [128 x 80]
constexpr uint32_t T_HC1 = T_CIN + 128 * 80 * 2;  // relu(colour L1)  [128 x 64]
constexpr uint32_t TILE_RECORD_BYTES = T_HC1 + 128 * 64 * 2;
static_assert(TILE_RECORD_BYTES == 118784, "928 bytes per sample");

// ---- transposed weight blob of the backward (K-major B operands of the data-gradient GEMMs), interleaved fp16 --------------
//   B1 (colour + sigma)                                        B2 (ambient)
//   WT_C2  [64 x 16]  = Wc2^T (3 -> 16)                         WT_A3 [64 x 16] = Wa3^T (2 -> 16)
//   WT_C1G [64 x 64]  = Wc1[:, 16:80]^T                         WT_A2 [64 x 64] = Wa2^T
//   WT_S3  [64 x 80]  = (Ws3 rows permuted geo-first)^T         WT_A1 [32 x 64] = Wa1[:, 0:32]^T
//   WT_S2  [64 x 64]  = Ws2^T
//   WT_S1  [64 x 64]  = Ws1[:, 0:64]^T
constexpr uint32_t BW1_C2 = 0;
constexpr uint32_t BW1_C1G = BW1_C2 + 64 * 16 * 2;
constexpr uint32_t BW1_S3 = BW1_C1G + 64 * 64 * 2;
constexpr uint32_t BW1_S2 = BW1_S3 + 64 * 80 * 2;
constexpr uint32_t BW1_S1 = BW1_S2 + 64 * 64 * 2;
constexpr uint32_t BW1_BYTES = BW1_S1 + 64 * 64 * 2;
constexpr uint32_t BW2_A3 = 0;
constexpr uint32_t BW2_A2 = BW2_A3 + 64 * 16 * 2;
constexpr uint32_t BW2_A1 = BW2_A2 + 64 * 64 * 2;
constexpr uint32_t BW2_BYTES = BW2_A1 + 32 * 64 * 2;
constexpr uint32_t BWD_BLOB_BYTES = BW1_BYTES + BW2_BYTES;   // B1 part first

// ---- weight-gradient accumulators in TMEM: column offsets behind the groups' working accumulators ---------------------------
// rows (TMEM lanes) are the M dimension of the MN-major product: `out` features for the square layers, `in` features for the two
// narrow output layers (their dW is kept transposed, N = 16 padded outputs).
constexpr uint32_t DW1_C2T = 0;      // [in 64  x out 16]
constexpr uint32_t DW1_C1 = 16;      // [out 64 x in 80]   column 0 of CIN is the constant SH band 0 -> column 0 / 0.2820948 = colsum(dZc1)
constexpr uint32_t DW1_S3 = 96;      // [out 80 (geo 0..63, log-density 64) x in 64]
constexpr uint32_t DW1_S2 = 160;     // [out 64 x in 64]
constexpr uint32_t DW1_S1 = 224;     // [out 64 x in 64]   (enc_x 0..31 | enc_w 32..63)
constexpr uint32_t DW1_CS = 288;     // [out 64 x 16]      column 0 = colsum(dZs1)  (the hoisted eye column)
constexpr uint32_t DW1_COLS = 304;
constexpr uint32_t DW2_A3T = 0;      // [in 64 x out 16]
constexpr uint32_t DW2_A2 = 16;      // [out 64 x in 64]
constexpr uint32_t DW2_A1 = 80;      // [out 64 x in 32]
constexpr uint32_t DW2_CS = 112;     // [out 64 x 16]      column 0 = colsum(dZa1)  (the hoisted audio columns)
constexpr uint32_t DW2_COLS = 128;

// ---- reduced weight gradients, flat fp32, nn.Linear [out, in] layouts --------------------------------------------------------
constexpr uint32_t G_WA1X = 0;                    // [64, 32]   ambient L1, encoder columns
constexpr uint32_t G_WA2 = G_WA1X + 64 * 32;      // [64, 64]
constexpr uint32_t G_WA3 = G_WA2 + 64 * 64;       // [2, 64]
constexpr uint32_t G_WS1 = G_WA3 + 2 * 64;        // [64, 64]   sigma L1, encoder columns (enc_x | enc_w)
constexpr uint32_t G_WS2 = G_WS1 + 64 * 64;       // [64, 64]
constexpr uint32_t G_WS3 = G_WS2 + 64 * 64;       // [65, 64]   original row order (row 0 = log-density)
constexpr uint32_t G_WC1 = G_WS3 + 65 * 64;       // [64, 80]   colour L1, [sh | geo] columns
constexpr uint32_t G_WC2 = G_WC1 + 64 * 80;       // [3, 64]
constexpr uint32_t G_CS_A1 = G_WC2 + 3 * 64;      // [64] colsum(dZ ambient L1)
constexpr uint32_t G_CS_S1 = G_CS_A1 + 64;        // [64] colsum(dZ sigma L1)
constexpr uint32_t G_CS_C1 = G_CS_S1 + 64;        // [64] colsum(dZ colour L1)
constexpr uint32_t G_FLOATS = G_CS_C1 + 64;

constexpr int BWD_GROUPS = 3;        // 128-thread tile groups per CTA in the backward kernels (+ one issuer warp)
constexpr uint32_t BWD_WORK_COLS = 64;

}  // namespace train
}  // namespace rn
