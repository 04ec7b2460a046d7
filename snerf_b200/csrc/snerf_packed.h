// snerf_packed.h -- layout of the packed weight images the kernels stream from L2/HBM.
// Shared by the packer (snerf_api.cu) and the kernels.  Host + device.
#pragma once
#include <stdint.h>

namespace snerf {

// ------------------------------------------------------------------------------------
// fp32 image (SNERF_MODE_FP32): a layer table followed by K-major ("transposed") weights
//   wide layer  : Wt[K_total][n_out] fp32, K_total = enc rows + hidden rows + dir rows, each
//                 segment padded with zero rows to a multiple of kFp32ChunkRows; streamed as
//                 chunks of kFp32ChunkRows rows by the bulk-copy engine;
//   narrow layer: w[n_out][K] fp32 row-major (heads with <= 4 outputs), read directly.
// ------------------------------------------------------------------------------------
constexpr int kFp32ChunkRows = 16;
constexpr int kFp32MaxLayers = 16;
constexpr int kEncRows = 64;   // 63 encoded point channels + 1 zero row
constexpr int kDirRows = 32;   // 27 encoded direction channels + zero rows
constexpr uint32_t kFp32Magic = 0x53463332u;  // 'SF32'
constexpr uint32_t kBf16Magic = 0x53423136u;  // 'SB16'
constexpr uint32_t kF16Magic = 0x53483136u;   // 'SH16' (same layout, fp16 operands)
constexpr uint32_t kF16x3Magic = 0x53483378u; // 'SH3x' (fp16 hi + lo chunk pairs, SNERF_MODE_FP16X3)

struct Fp32Layer {  // 44 bytes
  int32_t kind;         // 0 = wide, 1 = narrow
  int32_t n_out;        // wide: multiple of 32, <= 256; narrow: <= 4
  int32_t seg_rows[3];  // wide: padded K rows of the (enc, hidden, dir) segments; narrow: {0, K, 0}
  int32_t relu;
  int32_t src;          // activation buffer holding the hidden segment (0 = X, 1 = Y)
  int32_t dst;          // wide: activation buffer written; narrow: first raw column written
  uint32_t w_off;       // float offset (from image start) of the weights
  uint32_t b_off;       // float offset of bias[n_out]
  int32_t ch_off;       // wide: first channel of the output in the training activation store (see below)
};
struct Fp32Header {  // 64 + 16*44 = 768 bytes, weights start at kFp32DataOffset
  uint32_t magic;
  int32_t n_layers;
  int32_t W;
  int32_t chunks_per_tile;  // bulk copies one 64-row tile consumes
  int32_t pad[12];
  Fp32Layer layers[kFp32MaxLayers];
};
constexpr uint32_t kFp32DataOffset = 2048;  // bytes

// ------------------------------------------------------------------------------------
// training (fp32): activation store of one network pass, channel-major fp32 [channel][R] with R = rays x tiles
// per ray x 64 rows (row = (ray * tiles + tile) * 64 + r; rows past the ray's sample count repeat the last sample
// and receive zero gradient).  Channels: encoded point (64), encoded direction (32), then the outputs of the wide
// layers in table order (Fp32Layer.ch_off).  The gradient store d(loss)/d(pre-activation) uses the same channel
// numbering (input channels unused) and d_raw is [4][R].
// The backward image (SNERF_PACK_FP32_BWD) is a table of BwdStep followed by the un-transposed weight blocks
// W[n_out][hidden inputs], which is the [K][n] layout the tile GEMM needs for dX = dZ . W.
// ------------------------------------------------------------------------------------
constexpr int kSaveEncCh = 0;
constexpr int kSaveDirCh = kEncRows;
constexpr int kSaveActCh = kEncRows + kDirRows;
constexpr uint32_t kFp32BwdMagic = 0x53464257u;  // 'SFBW'
constexpr int kBwdMaxSteps = 20;

struct BwdStep {  // 48 bytes
  int32_t kind;       // 0 = wide (streamed GEMM), 1 = head transpose (d_raw columns -> hidden gradient)
  int32_t K;          // wide: contraction rows (= outputs of the forward layer, multiple of 16); head: #raw columns
  int32_t n_out;      // width of the produced gradient (inputs of the forward layer): multiple of 32
  int32_t src, dst;   // ping-pong buffers (0 = X, 1 = Y)
  int32_t raw_col;    // head: first d_raw column
  int32_t mask_ch;    // saved activation (channel offset) whose sign masks the result (ReLU'), -1 = none
  int32_t dz_ch;      // channel offset the masked result is stored at in the gradient store
  uint32_t w_off;     // float offset of the weights: wide [K][n_out]; head [K][n_out] (= nn.Linear layout)
  int32_t add_col;    // wide: >= 0 adds d_raw[add_col][r] * add_w[k] before masking (the alpha head), -1 = none
  uint32_t add_w_off; // float offset of add_w[n_out]
  int32_t pad;
};
struct Fp32BwdHeader {
  uint32_t magic;
  int32_t n_steps;
  int32_t W;
  int32_t n_channels;  // channels of the activation / gradient stores
  int32_t pad[12];
  BwdStep steps[kBwdMaxSteps];
};

// ------------------------------------------------------------------------------------
// bf16 image (SNERF_MODE_BF16): D=8, W=256, skip=4, 63/27 inputs, viewdirs.
// The MLP is ten tensor-core "steps" per 128-row tile; every step's B operand (weights,
// [N, K] K-major) is cut into chunks of 128 (N) x 64 (K) bf16 = 16 KiB stored as the exact
// 128B-swizzled shared-memory image tcgen05.mma reads, in consumption order:
//   step 0  L0        N=256 K=64(enc)          2 chunks   (n-half major, then k-block)
//   step 1-4 L1..L4   N=256 K=256              8 chunks each
//   step 5  L5        N=256 K=64(enc)+256     10 chunks
//   step 6-7 L6,L7    N=256 K=256              8 chunks each
//   step 8  feature   N=256 K=256              8 chunks
//   step 9  views     N=128 K=256              4 chunks   (direction part folded into a per-ray bias)
// followed by one fp32 parameter packet per step and the direction weights of the views layer.
// ------------------------------------------------------------------------------------
constexpr int kBfSteps = 10;
constexpr int kBfChunkBytes = 128 * 64 * 2;  // 16384
constexpr int kBfChunksPerTile = 2 + 8 * 4 + 10 + 8 * 2 + 8 + 4;  // 72
// packet of one step: [0,256) bias fp32 | [256,512) aux | [512,528) scalars | [528,1040) bias TILE: the step's bias as a
// tensor-core B operand, so that the accumulator starts at the bias instead of the epilogue adding it (non-split modes,
// steps 0..8).  Tile = 128 rows x 8 operand values (16 B per row, no-swizzle K-major core matrices: row n at byte 16 n):
//   row n = [t0, t1, t2 of bias[n] | t0, t1, t2 of bias[128 + n] | 0, 0],  bias = t0 + t1 + t2 in the operand type
// (three terms: fp32-exact for bf16; the matching A operand is a constant row [1,1,1,0,0,0,0,0] / [0,0,0,1,1,1,0,0]).
constexpr int kBfPacketHeadFloats = 528;     // what the split (fp16x3) mode stages in shared memory
constexpr int kBfBiasTileFloats = 512;       // 128 rows x 16 B
constexpr int kBfPacketFloats = kBfPacketHeadFloats + kBfBiasTileFloats;
constexpr int kBfPacketBytes = kBfPacketFloats * 4;  // 4160 (multiple of 16)
constexpr int kBfPacketHeadBytes = kBfPacketHeadFloats * 4;
constexpr uint32_t kBfHeaderBytes = 1024;
constexpr uint32_t kBfChunksOffset = kBfHeaderBytes;
constexpr uint32_t kBfPacketsOffset = kBfChunksOffset + kBfChunksPerTile * kBfChunkBytes;
constexpr uint32_t kBfDirWOffset = kBfPacketsOffset + kBfSteps * kBfPacketBytes;  // Wdir[128][32] fp32
constexpr uint32_t kBfImageBytes = kBfDirWOffset + 128 * 32 * 4;
// The split image (SNERF_MODE_FP16X3) stores every chunk twice, hi part then lo part (w = hi + lo, both fp16), so it
// has 2 x 72 chunks; packets and direction weights follow as above.
template <bool kSplit>
struct BfImage {
  static constexpr int kChunks = (kSplit ? 2 : 1) * kBfChunksPerTile;
  static constexpr uint32_t kPacketsOffset = kBfChunksOffset + (uint32_t)kChunks * kBfChunkBytes;
  static constexpr uint32_t kDirWOffset = kPacketsOffset + kBfSteps * kBfPacketBytes;
  static constexpr uint32_t kBytes = kDirWOffset + 128 * 32 * 4;
};
static_assert(BfImage<false>::kBytes == kBfImageBytes, "image layout");
// fp16 has a narrow exponent range: the lo part of a value below 2^-3 would be a subnormal (absolute resolution 2^-24),
// which costs the split its last bits exactly where NeRF activations (~1e-2) and weights (~1/16) live.  Both operand
// families are therefore pre-scaled by exact powers of two before splitting and the accumulator is scaled back in the
// epilogue: activations (and encoded inputs) x 2^4 (full precision down to |x| = 2^-7, overflow above 4094),
// weights x 2^8 (|w| < 255).
constexpr float kX3ActScale = 16.f;
constexpr float kX3WScale = 256.f;
constexpr float kX3InvScale = 1.f / (kX3ActScale * kX3WScale);

// training in the tensor-core modes (snerf_train_tc.cu): 16-bit activation / gradient stores of [slot][rows][256] values,
// laid out as 4 KiB blocks [slot][rb = row / 32][cb = channel / 64] of [c = 8-channel chunk (8)][r = row % 32][8 channels]:
//   * a warp whose lanes own consecutive rows writes a chunk as ONE 512-byte run (a [row][256] layout would put each
//     lane's 16 bytes into a different 128-byte line: 32 store wavefronts per instruction instead of 4);
//   * a block is exactly the no-swizzle MN-major UMMA canonical layout (core matrix = 8 rows x 16 bytes, 128 bytes between
//     8-row groups, 512 bytes between chunks), so the weight-gradient GEMM bulk-copies blocks and multiplies them as they lie.
// relu' masks: one bit per stored activation, [slot 0..8][rb][cb][r] 64-bit words (slot k = h_k, slot 8 = views);
// bit w of each 32-bit half = column 2 w, bit 16 + w = column 2 w + 1 of the half's 32 columns.
constexpr int kTcSlots = 10;
constexpr int kTcRowBytes = 512;
constexpr int kTcBlockBytes = 4096;
constexpr int kTcMaskSlots = 9;
__host__ __device__ inline long long tc_block_offset(long long rows, int slot, long long rb, int cb) {
  return (((long long)slot * (rows >> 5) + rb) * 4 + cb) * kTcBlockBytes;
}
__host__ __device__ inline long long tc_mask_index(long long rows, int slot, long long rb, int cb, int r) {
  return (((long long)slot * (rows >> 5) + rb) * 4 + cb) * 32 + r;
}

struct Bf16Header {
  uint32_t magic;
  int32_t depth;      // trunk depth of the packed network: 8, or 4 (steps 3..6 of the image are then unused, the network's
                      // layers 0,1,2,3 sit in steps 0,1,2,7: same kinds of step -- first / hidden / last-with-alpha)
  int32_t pad[14];
};

// chunks of step s (see table above)
__host__ __device__ inline int bf_step_chunks(int s) {
  return s == 0 ? 2 : (s == 5 ? 10 : (s == 9 ? 4 : 8));
}
__host__ __device__ inline int bf_step_first_chunk(int s) {
  int c = 0;
  for (int i = 0; i < s; ++i) c += bf_step_chunks(i);
  return c;
}

}  // namespace snerf
