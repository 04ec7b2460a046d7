// snerf_train_tc.h -- shared declarations of the tensor-core training step (snerf_train_tc.cu, snerf_api.cu)
#pragma once
#include "snerf_internal.h"
#include "snerf_packed.h"

namespace snerf {

// ---- backward image (SNERF_PACK_BF16_BWD): header | 68 chunks | alpha_w[256], rgb_w[3][128] fp32
constexpr int kBwSteps = 9;                      // dvp->dfeature, dfeature->dz7, dz7->dz6, ..., dz1->dz0
constexpr int kBwChunks = 4 + 8 * 8;
constexpr uint32_t kBwMagic = 0x53425742u;       // 'SBWB'
constexpr uint32_t kBwChunksOffset = kBfHeaderBytes;
constexpr uint32_t kBwParamsOffset = kBwChunksOffset + kBwChunks * kBfChunkBytes;
constexpr int kBwParamFloats = 256 + 384;
constexpr uint32_t kBwImageBytes = kBwParamsOffset + kBwParamFloats * 4;
__host__ __device__ inline int bw_step_chunks(int s) { return s == 0 ? 4 : 8; }

// ---- stores: [slot][rows][256] 16-bit; rows = 2 * ceil(n_rays / 2) * X (whole 128-row tiles)
// (kTcSlots, kTcRowBytes: snerf_packed.h)

struct BwdTcParams {
  const unsigned char* img[2];          // backward images of the coarse / fine network
  const unsigned long long* bits[2];    // relu' mask stores of the coarse / fine pass
  unsigned char* dz[2];                 // gradient stores
  const float4* draw[2];                // d_raw [n_rays * X][4] fp32 (composite_bwd_kernel)
  long long rows[2];                    // rows per slot
  long long valid_rows[2];              // n_rays * X: rows past it belong to the padding ray (zero gradient)
  int tiles[2];                         // rows / 128
};

struct TrainTcLayout {            // byte offsets into the training workspace
  long long rows_c, rows_f;
  size_t act_c, act_f, dz_c, dz_f, bits_c, bits_f, draw_c, draw_f, raw_c, raw_f, z_c, z_f, total_bytes;
};
TrainTcLayout train_tc_layout(int Nc, int Nf, long long n_rays);

constexpr int kMaxDwTcProblems = 28;
struct DwTcProblem {
  const unsigned char* A;  // column block of the gradient store holding channel 0 of the A operand, row block 0
  const unsigned char* B;  // likewise for the activation store / B operand
  float* C;                // gradient of the weight block, row (m - m_lo), leading dimension ldc
  float* bias;             // += column sums of the A operand for rows [m_lo, m_hi), or null
  long long R;             // rows (multiple of 64)
  int M, m_lo, m_hi;       // MMA rows (128 or 256) and the range of them that exists in C
  int N, Nmma, ldc, vec4;  // wanted columns, MMA columns (multiple of 16), vector reductions allowed
  int weight;              // 64-channel column blocks one 64-row block moves (work measure)
  long long first_block;   // position of the problem's first 64-row block in the linearised (problem, block) space
};
// cut[c] .. cut[c + 1] = the contiguous range of linearised 64-row blocks CTA c works on (host-side cost model:
// launch_dw_tc); `timing`: debug switch, per-CTA cycle counts go to a device-global table (snerf_debug_dw_timing)
constexpr int kMaxDwCtas = 160;
struct DwTcTable {
  int n, n_cta, timing;
  DwTcProblem p[kMaxDwTcProblems];
  long long cut[kMaxDwCtas + 1];
};

int pack_bwd_tc(const SnerfNetF32* src, void* packed, cudaStream_t stream);
int launch_dx_chain_tc(const BwdTcParams& p, cudaStream_t stream);
int launch_dw_tc(const BwdTcParams& p, const unsigned char* const act[2], const SnerfNetGradF32* gc, const SnerfNetGradF32* gf,
                 cudaStream_t stream);
int launch_tc_render_save(const RenderParams& p, cudaStream_t stream);
int debug_dw_timing(long long* out_host, int n_cta);
int debug_dw_cuts(long long rows_c, long long rows_f, int n_cta, long long* out_cut, long long* out_first, double* out_cost,
                  double* makespan);
int launch_composite_bwd_rows(const TrainParams& p, cudaStream_t stream);

}  // namespace snerf
