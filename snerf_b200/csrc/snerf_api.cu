// snerf_api.cu -- C ABI of libsnerf_b200.so (see include/snerf_b200.h), weight packers and the small
// stage kernels (composite, inverse-CDF, encoder, ray generator).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "snerf_common.cuh"
#include "snerf_internal.h"
#include "snerf_train_tc.h"
#include "snerf_packed.h"

namespace snerf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return SNERF_ERR_CUDA;
}
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 1;
}
static int require_sm100() {
  int dev = 0;
  if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return SNERF_ERR_CUDA;
  int major = 0;
  if (check_cuda(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev), "cudaDeviceGetAttribute"))
    return SNERF_ERR_CUDA;
  if (major != 10) {
    set_error("libsnerf_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
    return SNERF_ERR_ARCH;
  }
  return 0;
}

// ------------------------------------------------------------------------------------
// layer plan of the fp32 image (host)
// ------------------------------------------------------------------------------------
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

static bool desc_ok(const SnerfNetDesc* d) {
  if (!d) { set_error("null SnerfNetDesc"); return false; }
  if (d->D < 1 || d->D > 12) { set_error("trunk depth D=%d outside 1..12", d->D); return false; }
  if (d->W != 64 && d->W != 128 && d->W != 256) { set_error("width W=%d not in {64,128,256}", d->W); return false; }
  if (d->input_ch < 1 || d->input_ch > 63) { set_error("input_ch=%d outside 1..63", d->input_ch); return false; }
  if (d->use_viewdirs && (d->input_ch_views < 1 || d->input_ch_views > 27)) {
    set_error("input_ch_views=%d outside 1..27", d->input_ch_views); return false;
  }
  if (!d->use_viewdirs && (d->output_ch < 4 || d->output_ch > 8)) { set_error("output_ch=%d outside 4..8", d->output_ch); return false; }
  if (d->skip >= d->D - 1) { set_error("skip=%d must be < D-1=%d (pass -1 when no trunk layer concatenates)", d->skip, d->D - 1); return false; }
  return true;
}
static bool desc_is_flagship(const SnerfNetDesc* d) {
  return d->D == 8 && d->W == 256 && d->input_ch == 63 && d->input_ch_views == 27 && d->skip == 4 && d->use_viewdirs;
}

// NeRF(D=4, W=256) without a live skip: the coarse network create_nerf builds for the shipped configs (netdepth = 4).  The
// tensor-core renderer takes it as the COARSE network of a pair whose fine network is the flagship 8x256 (snerf_bf16_d4.cu).
static bool desc_is_coarse4(const SnerfNetDesc* d) {
  return d->D == 4 && d->W == 256 && d->input_ch == 63 && d->input_ch_views == 27 && d->skip < 0 && d->use_viewdirs;
}
static bool same_arch(const SnerfNetDesc* a, const SnerfNetDesc* b) {
  return a->D == b->D && a->W == b->W && a->skip == b->skip && a->input_ch == b->input_ch &&
         a->input_ch_views == b->input_ch_views && a->use_viewdirs == b->use_viewdirs && a->output_ch == b->output_ch;
}
// The fine network's descriptor (SnerfOpts.desc_fine), validated against the coarse one; nullptr + error on mismatch.
static const SnerfNetDesc* fine_desc(const SnerfNetDesc* d, const SnerfNetDesc* df) {
  if (!df) return d;
  if (!desc_ok(df)) return nullptr;
  if (df->input_ch != d->input_ch || df->input_ch_views != d->input_ch_views || df->use_viewdirs != d->use_viewdirs ||
      df->output_ch != d->output_ch) {
    set_error("desc_fine: input_ch / input_ch_views / use_viewdirs / output_ch must match the coarse network (both are fed "
              "by the same embedders)");
    return nullptr;
  }
  return df;
}

// Builds the layer table; returns total image bytes.
size_t plan_fp32(const SnerfNetDesc* d, Fp32Header* h, bool with_alpha) {
  memset(h, 0, sizeof(*h));
  h->magic = kFp32Magic;
  h->W = d->W;
  uint32_t off = kFp32DataOffset / 4;  // in floats
  int nl = 0, chunks = 0, buf = 0;     // buf = buffer holding the current hidden state
  int ch = kSaveActCh;                 // next free channel of the training activation store
  auto add_wide = [&](int enc, int hid, int dir, int n_out, int relu) {
    Fp32Layer& L = h->layers[nl++];
    L.ch_off = ch; ch += n_out;
    L.kind = 0; L.n_out = n_out; L.seg_rows[0] = enc; L.seg_rows[1] = hid; L.seg_rows[2] = dir; L.relu = relu;
    L.src = buf; L.dst = buf ^ 1;
    const int K = enc + hid + dir;
    off = (uint32_t)round_up((int)off, n_out);  // whole rows of the [rows][n_out] view the TMA descriptors use
    L.w_off = off; off += (uint32_t)K * n_out;
    L.b_off = off; off += (uint32_t)round_up(n_out, 4);
    chunks += K / kFp32ChunkRows;
    buf ^= 1;
  };
  auto add_narrow = [&](int K, int n_out, int dst_col) {
    Fp32Layer& L = h->layers[nl++];
    L.kind = 1; L.n_out = n_out; L.seg_rows[1] = K; L.src = buf; L.dst = dst_col;
    L.w_off = off; off += (uint32_t)K * n_out;
    L.b_off = off; off += 4;
  };
  for (int i = 0; i < d->D; ++i) {
    const bool has_enc = (i == 0) || (d->skip >= 0 && i - 1 == d->skip);
    add_wide(has_enc ? kEncRows : 0, i == 0 ? 0 : d->W, 0, d->W, 1);
  }
  if (d->use_viewdirs) {
    if (with_alpha) add_narrow(d->W, 1, 3);         // alpha_linear -> raw[...,3] (absent in NeRF_RGB)
    add_wide(0, d->W, 0, d->W, 0);                  // feature_linear (no activation)
    add_wide(0, d->W, kDirRows, d->W / 2, 1);       // views_linears.0 on [feature, dirs]
    add_narrow(d->W / 2, 3, 0);                     // rgb_linear -> raw[...,0:3]
  } else {
    add_narrow(d->W, 4, 0);                         // output_linear[:4]
  }
  h->n_layers = nl;
  h->chunks_per_tile = chunks;
  return (size_t)off * 4;
}

// ------------------------------------------------------------------------------------
// packers (device)
// ------------------------------------------------------------------------------------
struct WideSrc {
  const float* w;   // [n_out, ld]
  int ld;
  int n_out;
  int rows_pad[3];   // padded K rows per segment
  int rows_real[3];  // real K rows per segment
  int col0[3];       // first source column of each segment
};
__global__ void __launch_bounds__(256) pack_jobs_kernel(const __grid_constant__ PackJobs t) {
  const PackJob& J = t.j[blockIdx.y];
  if (J.kind == 0) {
    const int K = J.rows_pad[0] + J.rows_pad[1] + J.rows_pad[2];
    const long long total = (long long)K * J.n_out;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
      int k = (int)(idx / J.n_out);
      const int n = (int)(idx % J.n_out);
      float v = 0.f;
      for (int seg = 0; seg < 3; ++seg) {
        if (k < J.rows_pad[seg]) {
          if (k < J.rows_real[seg]) v = J.w[(long long)n * J.ld + J.col0[seg] + k];
          break;
        }
        k -= J.rows_pad[seg];
      }
      if (J.round_tf32) {  // nearest tf32 value (the tensor core would truncate)
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
        v = __uint_as_float(r);
      }
      J.dst[idx] = v;
    }
  } else {
    const long long total = (long long)J.rows * J.cols;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
      float v = J.w ? J.w[(idx / J.cols) * J.ld + J.col_first + (idx % J.cols)] : 0.f;   // absent head: zeros
      if (J.round_tf32) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
        v = __uint_as_float(r);
      }
      J.dst[idx] = v;
    }
  }
}
int launch_pack_jobs(const PackJobs& jobs, cudaStream_t stream) {
  if (jobs.n <= 0) return SNERF_OK;
  pack_jobs_kernel<<<dim3(32, (unsigned)jobs.n), 256, 0, stream>>>(jobs);
  return check_cuda(cudaGetLastError(), "launch pack_jobs_kernel");
}
__global__ void write_header_kernel(Fp32Header h, Fp32Header* dst) {
  const int n = sizeof(Fp32Header) / 4;
  const int* s = reinterpret_cast<const int*>(&h);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

struct Bf16Src {
  int depth;              // 8, or 4 (pts_w[3..6] null: those steps of the image are unused)
  const float* pts_w[8];
  const float* pts_b[8];
  const float *views_w, *views_b, *feature_w, *feature_b, *alpha_w, *alpha_b, *rgb_w, *rgb_b;
};
// one thread per (chunk, row, 8-wide k group)
// `split` (SNERF_MODE_FP16X3): image chunk 2c holds the fp16 "hi" part of chunk c, chunk 2c + 1 its "lo" part
// (w = hi + lo up to 2^-22 |w|).
// (bid / nblk = this block's index / the number of blocks working on the job: the bodies run both as kernels of their own
//  and as jobs of pack_tc_batch_kernel)
__device__ __forceinline__ void pack_bf16_chunks_body(const Bf16Src& s, unsigned char* __restrict__ img, int f16, int split,
                                                      int bid, int nblk) {
  const int total = (split ? 2 : 1) * kBfChunksPerTile * 128 * 8;
  for (int idx = bid * blockDim.x + threadIdx.x; idx < total; idx += nblk * blockDim.x) {
    const int ichunk = idx / 1024, row = (idx >> 3) & 127, g = idx & 7;
    const int chunk = split ? (ichunk >> 1) : ichunk, lo_part = split ? (ichunk & 1) : 0;
    int step = 0, first = 0;
    while (chunk >= first + bf_step_chunks(step)) { first += bf_step_chunks(step); ++step; }
    const int local = chunk - first;
    const int nkb = step == 0 ? 1 : (step == 5 ? 5 : 4);
    const int nh = local / nkb, kb = local % nkb;
    const int n = nh * 128 + row;
    const float* w;
    int ld, col0, valid;  // valid = number of real columns in this 64-wide k-block
    if (step <= 7) {
      w = s.pts_w[step];
      if (step == 0) { ld = 63; col0 = 0; valid = 63; }
      else if (step == 5) { ld = 319; if (kb == 0) { col0 = 0; valid = 63; } else { col0 = 63 + (kb - 1) * 64; valid = 64; } }
      else { ld = 256; col0 = kb * 64; valid = 64; }
    } else if (step == 8) { w = s.feature_w; ld = 256; col0 = kb * 64; valid = 64; }
    else { w = s.views_w; ld = 283; col0 = kb * 64; valid = 64; }
    if (!w) valid = 0;   // step of a deeper network than this one: zeros (never streamed)
    uint32_t out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k0 = g * 8 + 2 * i, k1 = k0 + 1;
      float a = k0 < valid ? w[(long long)n * ld + col0 + k0] : 0.f;
      float b = k1 < valid ? w[(long long)n * ld + col0 + k1] : 0.f;
      if (split) { a *= kX3WScale; b *= kX3WScale; }
      if (f16) {
        __half2 h = __floats2half2_rn(a, b);
        if (lo_part) {
          const float2 hf = __half22float2(h);
          h = __floats2half2_rn(__fsub_rn(a, hf.x), __fsub_rn(b, hf.y));
        }
        out[i] = *reinterpret_cast<uint32_t*>(&h);
      } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        out[i] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    const uint32_t off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((g ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(img + kBfChunksOffset + (size_t)ichunk * kBfChunkBytes + off) =
        make_uint4(out[0], out[1], out[2], out[3]);
  }
}
__global__ void pack_bf16_chunks_kernel(Bf16Src s, unsigned char* __restrict__ img, int f16, int split) {
  pack_bf16_chunks_body(s, img, f16, split, blockIdx.x, gridDim.x);
}
__device__ __forceinline__ void pack_bf16_params_body(const Bf16Src& s, unsigned char* __restrict__ img, int f16, int split,
                                                      int bid, int nblk) {
  float* pk = reinterpret_cast<float*>(img + (split ? BfImage<true>::kPacketsOffset : BfImage<false>::kPacketsOffset));
  float* dw = reinterpret_cast<float*>(img + (split ? BfImage<true>::kDirWOffset : BfImage<false>::kDirWOffset));
  const int tid = bid * blockDim.x + threadIdx.x, nth = nblk * blockDim.x;
  for (int i = tid; i < kBfSteps * kBfPacketFloats; i += nth) {
    const int step = i / kBfPacketFloats, j = i % kBfPacketFloats;
    float v = 0.f;
    if (step <= 7) {
      if (j < 256) v = s.pts_b[step] ? s.pts_b[step][j] : 0.f;
      else if (step == 7 && j < 512) v = s.alpha_w[j - 256];
      else if (step == 7 && j == 512) v = s.alpha_b[0];
    } else if (step == 8) {
      if (j < 256) v = s.feature_b[j];
    } else {
      if (j < 128) v = s.views_b[j];
      else if (j < 512) v = s.rgb_w[j - 128];  // [3][128] row-major
      else if (j < 515) v = s.rgb_b[j - 512];
    }
    if (j < kBfPacketHeadFloats) pk[i] = v;
    else if (split || step == 9) pk[i] = 0.f;      // no bias tile in the split mode / for the views step
  }
  // bias tiles of steps 0..8 (snerf_packed.h): three-term split of each bias in the operand type
  if (!split) {
    for (int i = tid; i < 9 * 256; i += nth) {
      const int step = i >> 8, c = i & 255, n = c & 127, half = c >> 7;
      const float b = step <= 7 ? (s.pts_b[step] ? s.pts_b[step][c] : 0.f) : s.feature_b[c];
      unsigned short t[3];
      float r = b;
      for (int q = 0; q < 3; ++q) {
        float back;
        if (f16) { const __half h = __float2half_rn(r); t[q] = __half_as_ushort(h); back = __half2float(h); }
        else { const __nv_bfloat16 h = __float2bfloat16_rn(r); t[q] = __bfloat16_as_ushort(h); back = __bfloat162float(h); }
        r = __fsub_rn(r, back);
      }
      unsigned short* row = reinterpret_cast<unsigned short*>(pk + (size_t)step * kBfPacketFloats + kBfPacketHeadFloats) + n * 8;
      row[3 * half + 0] = t[0]; row[3 * half + 1] = t[1]; row[3 * half + 2] = t[2];
      if (half == 0) { row[6] = 0; row[7] = 0; }
    }
  }
  for (int i = tid; i < 128 * 32; i += nth) {
    const int n = i >> 5, k = i & 31;
    dw[i] = k < 27 ? s.views_w[(long long)n * 283 + 256 + k] : 0.f;
  }
  if (tid == 0) {
    reinterpret_cast<Bf16Header*>(img)->magic = split ? kF16x3Magic : (f16 ? kF16Magic : kBf16Magic);
    reinterpret_cast<Bf16Header*>(img)->depth = s.depth;
  }
}
__global__ void pack_bf16_params_kernel(Bf16Src s, unsigned char* __restrict__ img, int f16, int split) {
  pack_bf16_params_body(s, img, f16, split, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------------------------
// backward weight image of the tensor-core training step ('SBWB', snerf_train_tc.h): the B operands of the nine chain
// steps ([N = inputs of the forward layer][K = its outputs], K-major, 128 x 64 chunks pre-swizzled like the forward
// image), then alpha_w[256] and rgb_w[3][128] in fp32
// ------------------------------------------------------------------------------------
struct BwPackSrc {
  const float* pts_w[8];
  const float *views_w, *feature_w, *alpha_w, *rgb_w;
};
__device__ __forceinline__ void pack_bw_body(const BwPackSrc& s, unsigned char* __restrict__ img, int bid, int nblk) {
  const int total = kBwChunks * 128 * 8;
  for (int idx = bid * blockDim.x + threadIdx.x; idx < total; idx += nblk * blockDim.x) {
    const int chunk = idx / 1024, row = (idx >> 3) & 127, g = idx & 7;
    int step = 0, first = 0;
    while (chunk >= first + bw_step_chunks(step)) { first += bw_step_chunks(step); ++step; }
    const int local = chunk - first;
    const int nkb = step == 0 ? 2 : 4;
    const int nh = local / nkb, kb = local % nkb;
    const int j = nh * 128 + row;  // output channel of the chain step = input channel of the forward layer
    const float* w;
    int ld, off;
    if (step == 0) { w = s.views_w; ld = 283; off = 0; }
    else if (step == 1) { w = s.feature_w; ld = 256; off = 0; }
    else { const int l = 9 - step; w = s.pts_w[l]; ld = l == 5 ? 319 : 256; off = l == 5 ? 63 : 0; }
    uint32_t out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k0 = kb * 64 + g * 8 + 2 * i;  // contraction index = output channel of the forward layer
      const __nv_bfloat162 h = __floats2bfloat162_rn(w[(long long)k0 * ld + off + j], w[(long long)(k0 + 1) * ld + off + j]);
      out[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    const uint32_t o = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((g ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(img + kBwChunksOffset + (size_t)chunk * kBfChunkBytes + o) = make_uint4(out[0], out[1], out[2], out[3]);
  }
  const int tid = bid * blockDim.x + threadIdx.x;
  float* pr = reinterpret_cast<float*>(img + kBwParamsOffset);
  for (int i = tid; i < kBwParamFloats; i += nblk * blockDim.x) pr[i] = i < 256 ? s.alpha_w[i] : s.rgb_w[i - 256];
  if (tid == 0) reinterpret_cast<Bf16Header*>(img)->magic = kBwMagic;
}
__global__ void pack_bw_chunks_kernel(BwPackSrc s, unsigned char* __restrict__ img) { pack_bw_body(s, img, blockIdx.x, gridDim.x); }

static int fill_bw_src(const SnerfNetF32* src, BwPackSrc& s) {
  for (int i = 0; i < 8; ++i) s.pts_w[i] = src->pts_w[i];
  s.views_w = src->views_w; s.feature_w = src->feature_w; s.alpha_w = src->alpha_w; s.rgb_w = src->rgb_w;
  if (!s.alpha_w) { set_error("training a network without alpha_linear (NeRF_RGB) is not supported"); return SNERF_ERR_UNSUPPORTED; }
  return SNERF_OK;
}
int pack_bwd_tc(const SnerfNetF32* src, void* packed, cudaStream_t stream) {
  BwPackSrc s;
  if (int e = fill_bw_src(src, s)) return e;
  pack_bw_chunks_kernel<<<272, 256, 0, stream>>>(s, (unsigned char*)packed);
  return check_cuda(cudaGetLastError(), "pack backward image (tensor-core)");
}

// Several tensor-core images in ONE launch (the training step re-packs four per iteration: forward + backward image of
// the coarse and of the fine network): blockIdx.y = job.
constexpr int kMaxTcPackJobs = 4;
constexpr int kPackChunkBlocks = 288, kPackParamBlocks = 16, kPackBwBlocks = 272;
struct TcPackJob {
  int kind;            // 0: forward image (chunks + packets), 1: backward image
  int f16, split;
  unsigned char* img;
  Bf16Src fwd;
  BwPackSrc bw;
};
struct TcPackJobs { int n; TcPackJob j[kMaxTcPackJobs]; };
__global__ void __launch_bounds__(256) pack_tc_batch_kernel(const __grid_constant__ TcPackJobs jobs) {
  const TcPackJob& J = jobs.j[blockIdx.y];
  const int bid = blockIdx.x;
  if (J.kind == 0) {
    if (bid < kPackChunkBlocks) pack_bf16_chunks_body(J.fwd, J.img, J.f16, J.split, bid, kPackChunkBlocks);
    else if (bid < kPackChunkBlocks + kPackParamBlocks) pack_bf16_params_body(J.fwd, J.img, J.f16, J.split, bid - kPackChunkBlocks, kPackParamBlocks);
  } else if (bid < kPackBwBlocks) {
    pack_bw_body(J.bw, J.img, bid, kPackBwBlocks);
  }
}

// ------------------------------------------------------------------------------------
// stage kernels
// ------------------------------------------------------------------------------------
__global__ void composite_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                 const float* __restrict__ rays_d, const float* __restrict__ noise, long long n_rays,
                                 int S, int white, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                                 float* depth_map) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n_rays; r += nwarps) {
    const float dx = rays_d[r * 3], dy = rays_d[r * 3 + 1], dz = rays_d[r * 3 + 2];
    const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const RayCarry c = composite_segment(reinterpret_cast<const float4*>(raw) + r * S, z + r * S, S, 0, S, dn,
                                         noise ? noise + r * S : nullptr, nullptr, weights ? weights + r * S : nullptr,
                                         carry_init(), lane);
    if (lane == 0) {
      const float wb = white ? (1.f - c.acc) : 0.f;
      if (rgb_map) { rgb_map[r * 3] = c.r + wb; rgb_map[r * 3 + 1] = c.g + wb; rgb_map[r * 3 + 2] = c.b + wb; }
      if (disp_map) disp_map[r] = disparity(c.depth, c.acc);
      if (acc_map) acc_map[r] = c.acc;
      if (depth_map) depth_map[r] = c.depth;
    }
  }
}

__global__ void __launch_bounds__(128) sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                                         const float* __restrict__ cdf_in, const float* __restrict__ u,
                                                         int u_per_ray, long long n_rays, int B, int n_out,
                                                         float* samples, long long* inds, float* cdf_out) {
  __shared__ float s_cdf[4][kMaxSamples];
  __shared__ float s_bins[4][kMaxSamples];
  __shared__ float s_w[4][kMaxSamples];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = blockIdx.x * 4LL + wib, nwarps = gridDim.x * 4LL;
  for (long long r = warp; r < n_rays; r += nwarps) {
    for (int i = lane; i < B; i += 32) s_bins[wib][i] = bins[r * B + i];
    if (cdf_in) {
      for (int i = lane; i < B; i += 32) s_cdf[wib][i] = cdf_in[r * B + i];
    } else {
      for (int i = lane; i < B - 1; i += 32) s_w[wib][i] = weights[r * (B - 1) + i];
      __syncwarp();
      build_cdf(s_w[wib], B, s_cdf[wib], lane);
    }
    __syncwarp();
    if (cdf_out) for (int i = lane; i < B; i += 32) cdf_out[r * B + i] = s_cdf[wib][i];
    for (int j = lane; j < n_out; j += 32) {
      const float uu = u_per_ray ? u[r * n_out + j] : u[j];
      int ind;
      const float sv = invert_cdf_one(s_bins[wib], s_cdf[wib], B, uu, &ind);
      if (samples) samples[r * n_out + j] = sv;
      if (inds) inds[r * n_out + j] = ind;
    }
    __syncwarp();
  }
}

__global__ void posenc_kernel(const float* __restrict__ x, long long n_rows, int L, float* __restrict__ out) {
  const int width = 3 + 6 * L;
  const long long total = n_rows * width;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / width;
    const int k = (int)(idx % width);
    float v;
    if (k < 3) v = x[m * 3 + k];
    else {
      const int o = (k - 3) / 6, j = (k - 3) % 6;
      const float a = x[m * 3 + j % 3] * __int_as_float((127 + o) << 23);
      v = j < 3 ? sinf(a) : cosf(a);
    }
    out[idx] = v;
  }
}

struct Cam { float m[12]; };
__global__ void get_rays_kernel(int H, int W, float focal, Cam c, float cx, float cy, float* __restrict__ ro,
                                float* __restrict__ rd) {
  const long long total = (long long)H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % W), j = (int)(idx / W);
    // dirs = ((i+.5-cx)/f, -((j+.5)-cy)/f, -1), rays_d[k] = sum_j dirs[j]*c2w[k][j]  (run_nerf_helpers.py:250-256)
    const float d0 = __fdiv_rn(__fsub_rn(__fadd_rn((float)i, 0.5f), cx), focal);
    const float d1 = -__fdiv_rn(__fsub_rn(__fadd_rn((float)j, 0.5f), cy), focal);
    const float d2 = -1.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rd[idx * 3 + k] = __fadd_rn(__fadd_rn(__fmul_rn(d0, c.m[k * 4 + 0]), __fmul_rn(d1, c.m[k * 4 + 1])),
                                  __fmul_rn(d2, c.m[k * 4 + 2]));
      ro[idx * 3 + k] = c.m[k * 4 + 3];
    }
  }
}

TrainTcLayout train_tc_layout(int Nc, int Nf, long long n_rays) {
  TrainTcLayout L{};
  const long long pairs = (n_rays + 1) / 2, S = Nc + Nf;
  L.rows_c = pairs * 2 * Nc;
  L.rows_f = Nf > 0 ? pairs * 2 * S : 0;
  size_t off = 0;
  auto take = [&](long long bytes) { size_t o_ = off; off += ((size_t)bytes + 1023) / 1024 * 1024; return o_; };
  L.act_c = take(L.rows_c * kTcSlots * kTcRowBytes);
  L.act_f = take(L.rows_f * kTcSlots * kTcRowBytes);
  L.dz_c = take(L.rows_c * kTcSlots * kTcRowBytes);
  L.dz_f = take(L.rows_f * kTcSlots * kTcRowBytes);
  L.bits_c = take(L.rows_c * kTcMaskSlots * 32);
  L.bits_f = take(L.rows_f * kTcMaskSlots * 32);
  L.draw_c = take(n_rays * Nc * 16);
  L.draw_f = take(Nf > 0 ? n_rays * S * 16 : 0);
  L.raw_c = take(n_rays * Nc * 16);
  L.raw_f = take(Nf > 0 ? n_rays * S * 16 : 0);
  L.z_c = take(n_rays * Nc * 4);
  L.z_f = take(Nf > 0 ? n_rays * S * 4 : 0);
  L.total_bytes = off;
  return L;
}


}  // namespace snerf

// ======================================================================================
// C ABI
// ======================================================================================
using namespace snerf;

extern "C" {

int snerf_version(void) { return SNERF_ABI_VERSION; }
const char* snerf_last_error(void) { return g_err; }

int snerf_device_check(int dev) {
  int count = 0;
  if (check_cuda(cudaGetDeviceCount(&count), "cudaGetDeviceCount")) return SNERF_ERR_CUDA;
  if (dev < 0 || dev >= count) { set_error("device %d out of range (have %d)", dev, count); return SNERF_ERR_BAD_ARG; }
  int major = 0;
  if (check_cuda(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev), "cudaDeviceGetAttribute"))
    return SNERF_ERR_CUDA;
  if (major != 10) { set_error("device %d is sm_%dx, need sm_100", dev, major); return SNERF_ERR_ARCH; }
  return SNERF_OK;
}

size_t snerf_packed_bytes(const SnerfNetDesc* desc, int mode) {
  if (!desc_ok(desc)) return 0;
  if (mode == SNERF_MODE_FP32 || mode == SNERF_PACK_TF32_FWD) {
    Fp32Header h;
    return plan_fp32(desc, &h, true);
  }
  if (mode == SNERF_MODE_BF16 || mode == SNERF_MODE_FP16 || mode == SNERF_MODE_FP16X3) {
    if (!desc_is_flagship(desc) && !desc_is_coarse4(desc)) {
      set_error("tensor-core modes support NeRF(D=8, W=256, skips=[4], input_ch=63, input_ch_views=27, use_viewdirs), and "
                "NeRF(D=4, W=256, no live skip) as the coarse network beside it");
      return 0;
    }
    return mode == SNERF_MODE_FP16X3 ? BfImage<true>::kBytes : BfImage<false>::kBytes;
  }
  if (mode == SNERF_PACK_FP32_BWD || mode == SNERF_PACK_TF32_BWD) {
    if (!train_supported(desc, mode == SNERF_PACK_TF32_BWD)) return 0;
    Fp32BwdHeader h;
    return plan_bwd(desc, &h);
  }
  if (mode == SNERF_PACK_BF16_BWD) {
    if (!desc_is_flagship(desc)) {
      set_error("tensor-core training supports NeRF(D=8, W=256, skips=[4], input_ch=63, input_ch_views=27, use_viewdirs)");
      return 0;
    }
    return kBwImageBytes;
  }
  set_error("unknown mode %d", mode);
  return 0;
}

// argument checks shared by snerf_pack_weights and snerf_pack_weights_batch
static int validate_pack(const SnerfNetDesc* d, const SnerfNetF32* src, void* packed, size_t packed_bytes, int mode) {
  if (!desc_ok(d)) return SNERF_ERR_BAD_ARG;
  if (!src || !packed) { set_error("null argument"); return SNERF_ERR_BAD_ARG; }
  if ((reinterpret_cast<uintptr_t>(packed) & 127) != 0) { set_error("packed image must be 128-byte aligned"); return SNERF_ERR_BAD_ARG; }
  const size_t need = snerf_packed_bytes(d, mode);
  if (need == 0) return SNERF_ERR_UNSUPPORTED;
  if (packed_bytes < need) { set_error("packed buffer too small: %zu < %zu", packed_bytes, need); return SNERF_ERR_WORKSPACE; }
  for (int i = 0; i < d->D; ++i)
    if (!src->pts_w[i] || !src->pts_b[i]) { set_error("pts_linears.%d missing", i); return SNERF_ERR_BAD_ARG; }
  if (d->use_viewdirs) {
    if (!src->views_w || !src->views_b || !src->feature_w || !src->feature_b || !src->rgb_w || !src->rgb_b) {
      set_error("view-dependent heads missing"); return SNERF_ERR_BAD_ARG;
    }
    if ((src->alpha_w == nullptr) != (src->alpha_b == nullptr)) { set_error("alpha_linear weight/bias mismatch"); return SNERF_ERR_BAD_ARG; }
    if (!src->alpha_w && mode != SNERF_MODE_FP32 && mode != SNERF_PACK_FP32_BWD) {
      set_error("a network without alpha_linear (NeRF_RGB) is supported in fp32 mode only"); return SNERF_ERR_UNSUPPORTED;
    }
  } else if (!src->output_w || !src->output_b) { set_error("output_linear missing"); return SNERF_ERR_BAD_ARG; }
  return require_sm100();
}
// operands of the forward-image packers (tensor-core modes)
static void fill_fwd_src(const SnerfNetDesc* d, const SnerfNetF32* src, Bf16Src& s) {
  for (int i = 0; i < 8; ++i) { s.pts_w[i] = src->pts_w[i]; s.pts_b[i] = src->pts_b[i]; }
  s.depth = d->D;
  if (d->D == 4) {   // layers 0,1,2,3 -> steps 0,1,2,7 of the image (first / hidden / hidden / last-with-alpha); 3..6 stay empty
    for (int i = 3; i < 8; ++i) { s.pts_w[i] = nullptr; s.pts_b[i] = nullptr; }
    s.pts_w[7] = src->pts_w[3]; s.pts_b[7] = src->pts_b[3];
  }
  s.views_w = src->views_w; s.views_b = src->views_b; s.feature_w = src->feature_w; s.feature_b = src->feature_b;
  s.alpha_w = src->alpha_w; s.alpha_b = src->alpha_b; s.rgb_w = src->rgb_w; s.rgb_b = src->rgb_b;
}
static bool is_tc_mode(int mode) { return mode == SNERF_MODE_BF16 || mode == SNERF_MODE_FP16 || mode == SNERF_MODE_FP16X3; }

int snerf_pack_weights_batch(int32_t n, const SnerfNetDesc* const* descs, const SnerfNetF32* const* srcs, void* const* packed,
                             const size_t* packed_bytes, const int32_t* modes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || (n > 0 && (!descs || !srcs || !packed || !packed_bytes || !modes))) { set_error("null argument"); return SNERF_ERR_BAD_ARG; }
  TcPackJobs jobs{};
  auto flush = [&]() -> int {
    if (jobs.n == 0) return SNERF_OK;
    pack_tc_batch_kernel<<<dim3(kPackChunkBlocks + kPackParamBlocks, (unsigned)jobs.n), 256, 0, stream>>>(jobs);
    jobs.n = 0;
    return check_cuda(cudaGetLastError(), "launch pack_tc_batch_kernel");
  };
  for (int i = 0; i < n; ++i) {
    const int mode = modes[i];
    if (!is_tc_mode(mode) && mode != SNERF_PACK_BF16_BWD) {   // the other images keep their own packers
      if (int e = snerf_pack_weights(descs[i], srcs[i], packed[i], packed_bytes[i], mode, stream_)) return e;
      continue;
    }
    if (int e = validate_pack(descs[i], srcs[i], packed[i], packed_bytes[i], mode)) return e;
    if (jobs.n == kMaxTcPackJobs) if (int e = flush()) return e;
    TcPackJob& J = jobs.j[jobs.n];
    J.img = reinterpret_cast<unsigned char*>(packed[i]);
    if (mode == SNERF_PACK_BF16_BWD) {
      J.kind = 1;
      if (int e = fill_bw_src(srcs[i], J.bw)) return e;
    } else {
      J.kind = 0; J.f16 = mode != SNERF_MODE_BF16 ? 1 : 0; J.split = mode == SNERF_MODE_FP16X3 ? 1 : 0;
      fill_fwd_src(descs[i], srcs[i], J.fwd);
    }
    ++jobs.n;
  }
  return flush();
}

int snerf_pack_weights(const SnerfNetDesc* d, const SnerfNetF32* src, void* packed, size_t packed_bytes, int mode,
                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = validate_pack(d, src, packed, packed_bytes, mode)) return e;
  if (mode == SNERF_PACK_FP32_BWD || mode == SNERF_PACK_TF32_BWD)
    return pack_bwd(d, src, packed, mode == SNERF_PACK_TF32_BWD ? 1 : 0, stream);
  if (mode == SNERF_PACK_BF16_BWD) return pack_bwd_tc(src, packed, stream);

  if (is_tc_mode(mode)) {
    const int f16 = mode != SNERF_MODE_BF16 ? 1 : 0, split = mode == SNERF_MODE_FP16X3 ? 1 : 0;
    Bf16Src s;
    fill_fwd_src(d, src, s);
    pack_bf16_chunks_kernel<<<kPackChunkBlocks, 256, 0, stream>>>(s, (unsigned char*)packed, f16, split);
    pack_bf16_params_kernel<<<kPackParamBlocks, 256, 0, stream>>>(s, (unsigned char*)packed, f16, split);
    return check_cuda(cudaGetLastError(), "pack bf16");
  }

  if (mode != SNERF_MODE_FP32 && mode != SNERF_PACK_TF32_FWD) { set_error("unknown mode %d", mode); return SNERF_ERR_BAD_ARG; }
  const int tf32_fwd = mode == SNERF_PACK_TF32_FWD ? 1 : 0;
  Fp32Header h;
  const bool with_alpha = !d->use_viewdirs || src->alpha_w != nullptr;
  plan_fp32(d, &h, with_alpha);
  float* base = reinterpret_cast<float*>(packed);
  write_header_kernel<<<1, 128, 0, stream>>>(h, reinterpret_cast<Fp32Header*>(packed));
  PackJobs jobs{};
  auto copy = [&](uint32_t off, const float* s_, size_t n) {       // plain copy = one-row block
    if (jobs.n == kMaxPackJobs) { launch_pack_jobs(jobs, stream); jobs.n = 0; }   // table full: flush
    PackJob& J = jobs.j[jobs.n++];
    J.w = s_; J.dst = base + off; J.kind = 1; J.ld = (int)n; J.rows = 1; J.cols = (int)n; J.col_first = 0; J.round_tf32 = 0;
  };
  auto wide = [&](const WideSrc& s_, uint32_t off) {
    if (jobs.n == kMaxPackJobs) { launch_pack_jobs(jobs, stream); jobs.n = 0; }   // table full: flush
    PackJob& J = jobs.j[jobs.n++];
    J.w = s_.w; J.dst = base + off; J.kind = 0; J.ld = s_.ld; J.n_out = s_.n_out; J.round_tf32 = tf32_fwd;
    for (int q = 0; q < 3; ++q) { J.rows_pad[q] = s_.rows_pad[q]; J.rows_real[q] = s_.rows_real[q]; J.col0[q] = s_.col0[q]; }
  };
  int l = 0;
  for (int i = 0; i < d->D; ++i, ++l) {
    const Fp32Layer& L = h.layers[l];
    WideSrc s{};
    const bool has_enc = L.seg_rows[0] > 0;
    s.w = src->pts_w[i];
    s.n_out = L.n_out;
    s.ld = (has_enc ? d->input_ch : 0) + (i == 0 ? 0 : d->W);
    s.rows_pad[0] = L.seg_rows[0]; s.rows_real[0] = has_enc ? d->input_ch : 0; s.col0[0] = 0;
    s.rows_pad[1] = L.seg_rows[1]; s.rows_real[1] = L.seg_rows[1]; s.col0[1] = has_enc ? d->input_ch : 0;
    wide(s, L.w_off);
    copy(L.b_off, src->pts_b[i], L.n_out);
  }
  if (d->use_viewdirs) {
    if (with_alpha) {  // alpha
      const Fp32Layer& L = h.layers[l++];
      copy(L.w_off, src->alpha_w, d->W);
      copy(L.b_off, src->alpha_b, 1);
    }
    {  // feature
      const Fp32Layer& L = h.layers[l++];
      WideSrc s{};
      s.w = src->feature_w; s.n_out = L.n_out; s.ld = d->W;
      s.rows_pad[1] = d->W; s.rows_real[1] = d->W; s.col0[1] = 0;
      wide(s, L.w_off);
      copy(L.b_off, src->feature_b, L.n_out);
    }
    {  // views
      const Fp32Layer& L = h.layers[l++];
      WideSrc s{};
      s.w = src->views_w; s.n_out = L.n_out; s.ld = d->W + d->input_ch_views;
      s.rows_pad[1] = d->W; s.rows_real[1] = d->W; s.col0[1] = 0;
      s.rows_pad[2] = kDirRows; s.rows_real[2] = d->input_ch_views; s.col0[2] = d->W;
      wide(s, L.w_off);
      copy(L.b_off, src->views_b, L.n_out);
    }
    {  // rgb
      const Fp32Layer& L = h.layers[l++];
      copy(L.w_off, src->rgb_w, (size_t)3 * (d->W / 2));
      copy(L.b_off, src->rgb_b, 3);
    }
  } else {
    const Fp32Layer& L = h.layers[l++];
    copy(L.w_off, src->output_w, (size_t)4 * d->W);
    copy(L.b_off, src->output_b, 4);
  }
  if (int e = launch_pack_jobs(jobs, stream)) return e;
  return check_cuda(cudaGetLastError(), "pack fp32");
}

size_t snerf_query_workspace(const SnerfNetDesc*, const SnerfOpts*, int64_t) {
  return 0;  // the fused kernels keep every intermediate on-chip
}

static int fill_common(RenderParams& p, const SnerfNetDesc* d, const SnerfOpts* o) {
  p.L = o->multires < 0 ? 0 : o->multires;
  p.Lv = o->multires_views < 0 ? 0 : o->multires_views;
  if (3 + 6 * p.L != d->input_ch) { set_error("multires=%d does not match input_ch=%d", o->multires, d->input_ch); return SNERF_ERR_BAD_ARG; }
  if (d->use_viewdirs && 3 + 6 * p.Lv != d->input_ch_views) {
    set_error("multires_views=%d does not match input_ch_views=%d", o->multires_views, d->input_ch_views);
    return SNERF_ERR_BAD_ARG;
  }
  return 0;
}

int snerf_render_rays_fwd(const SnerfRays* rays, const SnerfNetDesc* d, const void* packed_coarse,
                          const void* packed_fine, const SnerfOpts* o, const SnerfOut* out, void* workspace,
                          size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!rays || !o || !out || !packed_coarse) { set_error("null argument"); return SNERF_ERR_BAD_ARG; }
  if (!desc_ok(d)) return SNERF_ERR_BAD_ARG;
  const SnerfCamera* cam = o->camera;
  if (cam) {
    if (rays->ray_batch) { set_error("camera mode: rays->ray_batch must be NULL"); return SNERF_ERR_BAD_ARG; }
    if (cam->H <= 0 || cam->W <= 0 || !(cam->focal > 0.f) || cam->first_pixel < 0 || rays->n_rays < 0 ||
        cam->first_pixel + rays->n_rays > (int64_t)cam->H * cam->W) {
      set_error("camera mode: bad image size / focal / pixel range"); return SNERF_ERR_BAD_ARG;
    }
    if (o->save_for_backward) { set_error("camera mode is inference-only (training takes a ray batch)"); return SNERF_ERR_UNSUPPORTED; }
  } else {
    if (rays->n_rays < 0 || (rays->n_rays > 0 && !rays->ray_batch)) { set_error("bad ray batch"); return SNERF_ERR_BAD_ARG; }
    if (rays->width != 8 && rays->width != 9 && rays->width != 11 && rays->width != 12) {
      set_error("ray batch width %d not in {8,9,11,12}", rays->width); return SNERF_ERR_BAD_ARG;
    }
    if (rays->row_stride < rays->width) { set_error("row_stride < width"); return SNERF_ERR_BAD_ARG; }
  }
  const int has_vd = cam ? 1 : rays->width > 9;
  if (d->use_viewdirs && !has_vd) { set_error("network uses view directions but the ray batch has none"); return SNERF_ERR_BAD_ARG; }
  if (o->n_samples < 2 || o->n_importance < 0 || o->n_samples + o->n_importance > kMaxSamples) {
    set_error("n_samples=%d n_importance=%d unsupported (need 2 <= Nc, Nc+Nf <= %d)", o->n_samples, o->n_importance, kMaxSamples);
    return SNERF_ERR_UNSUPPORTED;
  }
  if (o->n_importance > 0 && o->n_samples < 3) { set_error("hierarchical sampling needs n_samples >= 3"); return SNERF_ERR_UNSUPPORTED; }
  if (!o->t_vals) { set_error("t_vals missing"); return SNERF_ERR_BAD_ARG; }
  if (o->n_importance > 0 && !o->u_vals && !o->u_rand) { set_error("u_vals / u_rand missing"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  if (rays->n_rays == 0) return SNERF_OK;

  RenderParams p{};
  if (int e = fill_common(p, d, o)) return e;
  p.ray_batch = rays->ray_batch; p.n_rays = rays->n_rays; p.width = rays->width; p.row_stride = rays->row_stride;
  p.has_vd = has_vd;
  if (cam) {
    p.cam_on = 1; p.cam_W = cam->W; p.cam_focal = cam->focal; p.cam_cx = cam->cx; p.cam_cy = cam->cy;
    p.cam_near = cam->near; p.cam_far = cam->far; p.cam_first = cam->first_pixel;
    for (int i = 0; i < 12; ++i) p.cam_m[i] = cam->c2w[i];
    p.width = 11; p.row_stride = 11;
  }
  p.Nc = o->n_samples; p.Nf = o->n_importance; p.lindisp = o->lindisp; p.white_bkgd = o->white_bkgd;
  p.t_vals = o->t_vals; p.u_vals = o->u_vals; p.t_rand = o->t_rand; p.u_rand = o->u_rand;
  p.noise0 = o->noise0; p.noise1 = o->noise1;
  p.out = *out;
  p.img_coarse = (const unsigned char*)packed_coarse;
  p.img_fine = (const unsigned char*)(packed_fine ? packed_fine : packed_coarse);
  p.img_alpha_coarse = (const unsigned char*)o->packed_alpha_coarse;
  p.img_alpha_fine = (const unsigned char*)o->packed_alpha_fine;
  if ((p.img_alpha_coarse || p.img_alpha_fine) && o->mode != SNERF_MODE_FP32) {
    set_error("frozen sigma networks (NeRF_RGB alpha_model) are supported in fp32 mode only");
    return SNERF_ERR_UNSUPPORTED;
  }
  const SnerfNetDesc* df = fine_desc(d, o->desc_fine);
  if (!df) return SNERF_ERR_BAD_ARG;
  const bool two_archs = !same_arch(d, df);
  const bool tc_mode = o->mode == SNERF_MODE_BF16 || o->mode == SNERF_MODE_FP16 || o->mode == SNERF_MODE_FP16X3;
  // the one mixed pair the tensor-core renderer takes: coarse 4x256 (no skip) under the flagship 8x256 fine network
  const bool tc_pair_4_8 = tc_mode && !o->save_for_backward && desc_is_coarse4(d) && desc_is_flagship(df) && packed_fine;
  if (two_archs && !packed_fine) { set_error("desc_fine given without a fine network image"); return SNERF_ERR_UNSUPPORTED; }
  if (two_archs && o->mode != SNERF_MODE_FP32 && !tc_pair_4_8) {
    set_error("coarse and fine networks of different architectures (desc_fine) run in fp32 mode; the tensor-core modes take "
              "the pair coarse NeRF(D=4, W=256, no live skip) + fine NeRF(D=8, W=256, skips=[4]) (inference)");
    return SNERF_ERR_UNSUPPORTED;
  }
  const int Wmax = d->W > df->W ? d->W : df->W;   // the CTA is sized for the wider network (frozen sigma nets: <= Wmax)

  if (o->save_for_backward && o->mode == SNERF_MODE_FP16) {
    set_error("training needs bf16 stores (tcgen05 kind::f16 rejects the fp16 x bf16 operand pair of the weight-gradient GEMM)");
    return SNERF_ERR_UNSUPPORTED;
  }
  if (o->save_for_backward && o->mode == SNERF_MODE_BF16) {
    // tensor-core training forward: the fused renderer + 16-bit activation stores (snerf_train_tc.cu)
    if (!desc_is_flagship(d) || !has_vd || !bf16_geometry_supported(o->n_samples, o->n_importance)) {
      set_error("tensor-core training runs NeRF(8x256, skips=[4], viewdirs) with (N_samples, N_importance) in "
                "{(64,0),(64,64),(64,128),(64,192),(128,0),(128,128)}; use train precision fp32 / tf32 otherwise");
      return SNERF_ERR_UNSUPPORTED;
    }
    if (p.img_alpha_coarse || p.img_alpha_fine) { set_error("training with a frozen alpha_model (NeRF_RGB) is not supported"); return SNERF_ERR_UNSUPPORTED; }
    const TrainTcLayout L = train_tc_layout(p.Nc, p.Nf, p.n_rays);
    if (!workspace || workspace_bytes < L.total_bytes) {
      set_error("training workspace too small: %zu < %zu bytes", workspace_bytes, L.total_bytes);
      return SNERF_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(workspace) & 127) != 0) { set_error("workspace must be 128-byte aligned"); return SNERF_ERR_BAD_ARG; }
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    p.tc_op = o->mode == SNERF_MODE_FP16 ? 1 : 0;
    p.act_c = ws + L.act_c; p.act_rows_c = L.rows_c;
    p.act_f = ws + L.act_f; p.act_rows_f = L.rows_f;
    p.bits_c = reinterpret_cast<unsigned long long*>(ws + L.bits_c);
    p.bits_f = reinterpret_cast<unsigned long long*>(ws + L.bits_f);
    const long long N = p.n_rays, S = p.Nc + p.Nf;
    float* raw_c = reinterpret_cast<float*>(ws + L.raw_c);
    float* raw_f = reinterpret_cast<float*>(ws + L.raw_f);
    float* z_c = reinterpret_cast<float*>(ws + L.z_c);
    float* z_f = reinterpret_cast<float*>(ws + L.z_f);
    if (p.Nf > 0) { p.out.raw_coarse = raw_c; p.out.raw = raw_f; p.out.z_all = z_f; }
    else { p.out.raw = raw_c; p.out.raw_coarse = nullptr; }
    p.out.z_vals_map = z_c;
    if (const char* dbg = getenv("SNERF_TC_SAVE_SKIP")) {   // timing experiments only (wrong gradients): 1 = no relu' bits, 2 = no activation stores
      const int v = atoi(dbg);
      if (v & 1) { p.bits_c = nullptr; p.bits_f = nullptr; }
      if (v & 2) { p.act_c = nullptr; p.act_f = nullptr; }
    }
    if (int e = launch_tc_render_save(p, stream)) return e;
    auto give = [&](float* user, const float* mine, long long n) {
      if (user) cudaMemcpyAsync(user, mine, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream);
    };
    give(out->z_vals_map, z_c, N * p.Nc);
    if (p.Nf > 0) {
      give(out->raw_coarse, raw_c, N * p.Nc * 4);
      give(out->raw, raw_f, N * S * 4);
      give(out->z_all, z_f, N * S);
    } else {
      give(out->raw, raw_c, N * p.Nc * 4);
      give(out->raw_coarse, raw_c, N * p.Nc * 4);
    }
    return check_cuda(cudaGetLastError(), "training forward (tensor-core)");
  }
  if (o->save_for_backward) {
    // training forward: layer activations, raw and depths of both passes stay in the workspace for the backward
    if (o->mode != SNERF_MODE_FP32 && o->mode != SNERF_MODE_TF32) {
      set_error("save_for_backward needs mode fp32, tf32, bf16 or fp16"); return SNERF_ERR_UNSUPPORTED;
    }
    if (!train_supported(d, o->mode == SNERF_MODE_TF32)) return SNERF_ERR_UNSUPPORTED;
    if ((p.img_alpha_coarse || p.img_alpha_fine) && o->mode != SNERF_MODE_FP32) {
      set_error("training with a frozen alpha_model (NeRF_RGB) runs at train precision fp32 only"); return SNERF_ERR_UNSUPPORTED;
    }
    const TrainLayout L = train_layout(d, df, p.Nc, p.Nf, p.n_rays);
    if (!workspace || workspace_bytes < L.total_floats * 4) {
      set_error("training workspace too small: %zu < %zu bytes", workspace_bytes, L.total_floats * 4);
      return SNERF_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(workspace) & 127) != 0) { set_error("workspace must be 128-byte aligned"); return SNERF_ERR_BAD_ARG; }
    float* ws = reinterpret_cast<float*>(workspace);
    p.save_c = ws + L.save_c; p.Rc = L.Rc;
    if (p.Nf > 0) { p.save_f = ws + L.save_f; p.Rf = L.Rf; }
    const long long N = p.n_rays, S = p.Nc + p.Nf;
    if (p.Nf > 0) {
      p.out.raw_coarse = ws + L.raw_c; p.out.raw = ws + L.raw_f; p.out.z_all = ws + L.z_f;
    } else {
      p.out.raw = ws + L.raw_c; p.out.raw_coarse = nullptr;
    }
    p.out.z_vals_map = ws + L.z_c;
    if (o->mode == SNERF_MODE_TF32) {
      if (p.Nf > 0 && !p.out.weights_fine) p.out.weights_fine = nullptr;
      if (int e = launch_train_forward_tf32(d, p, L, ws, stream)) return e;
    } else if (int e = launch_fp32(FE_RAYS, Wmax, p, stream)) return e;
    auto give = [&](float* user, const float* mine, long long n) {
      if (user) cudaMemcpyAsync(user, mine, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream);
    };
    give(out->z_vals_map, ws + L.z_c, N * p.Nc);
    if (p.Nf > 0) {
      give(out->raw_coarse, ws + L.raw_c, N * p.Nc * 4);
      give(out->raw, ws + L.raw_f, N * S * 4);
      give(out->z_all, ws + L.z_f, N * S);
    } else {
      give(out->raw, ws + L.raw_c, N * p.Nc * 4);
      give(out->raw_coarse, ws + L.raw_c, N * p.Nc * 4);
    }
    return check_cuda(cudaGetLastError(), "training forward");
  }
  if (o->mode == SNERF_MODE_FP32) return launch_fp32(FE_RAYS, Wmax, p, stream);
  if (o->mode == SNERF_MODE_BF16 || o->mode == SNERF_MODE_FP16 || o->mode == SNERF_MODE_FP16X3) {
    p.tc_op = o->mode == SNERF_MODE_FP16X3 ? 2 : (o->mode == SNERF_MODE_FP16 ? 1 : 0);  // OP_BF16 / OP_F16 / OP_F16X3
    // coarse 4x256: with a fine pass it needs the 8x256 fine network's own image (the kernel walks the fine tiles as 8 layers)
    const bool coarse4 = desc_is_coarse4(d) && (o->n_importance == 0 || tc_pair_4_8);
    if (!(desc_is_flagship(d) || coarse4) || !has_vd || !bf16_geometry_supported(o->n_samples, o->n_importance)) {
      set_error("bf16 / fp16 mode runs NeRF(8x256, skips=[4], viewdirs) -- or NeRF(4x256) as the coarse network beside it -- "
                "with (N_samples, N_importance) in {(64,0),(64,64),(64,128),(64,192),(128,0),(128,128)}; use mode fp32 otherwise");
      return SNERF_ERR_UNSUPPORTED;
    }
    p.coarse_depth = coarse4 ? 4 : 8;
    return launch_bf16_render(p, stream);
  }
  set_error("unknown mode %d", o->mode);
  return SNERF_ERR_BAD_ARG;
}

size_t snerf_train_workspace_bytes(const SnerfNetDesc* d, int32_t n_samples, int32_t n_importance, int64_t n_rays) {
  if (!desc_ok(d)) return 0;
  if (n_samples < 2 || n_importance < 0 || n_samples + n_importance > kMaxSamples || n_rays < 0) {
    set_error("bad sample / ray counts"); return 0;
  }
  return train_layout(d, nullptr, n_samples, n_importance, n_rays).total_floats * 4 + 128;
}

size_t snerf_train_workspace_bytes_pair(const SnerfNetDesc* d, const SnerfNetDesc* d_fine, int32_t n_samples,
                                        int32_t n_importance, int64_t n_rays, int32_t mode) {
  if (!desc_ok(d)) return 0;
  const SnerfNetDesc* df = fine_desc(d, d_fine);
  if (!df) return 0;
  if (same_arch(d, df)) return snerf_train_workspace_bytes_mode(d, n_samples, n_importance, n_rays, mode);
  if (mode != SNERF_MODE_FP32) { set_error("coarse and fine networks of different architectures train at the fp32 level only"); return 0; }
  if (n_samples < 2 || n_importance < 0 || n_samples + n_importance > kMaxSamples || n_rays < 0) {
    set_error("bad sample / ray counts"); return 0;
  }
  return train_layout(d, df, n_samples, n_importance, n_rays).total_floats * 4 + 128;
}

size_t snerf_train_workspace_bytes_mode(const SnerfNetDesc* d, int32_t n_samples, int32_t n_importance, int64_t n_rays,
                                        int32_t mode) {
  if (mode != SNERF_MODE_BF16) return snerf_train_workspace_bytes(d, n_samples, n_importance, n_rays);
  if (!desc_ok(d) || !desc_is_flagship(d) || !bf16_geometry_supported(n_samples, n_importance) || n_rays < 0) {
    set_error("tensor-core training: unsupported network / sample counts"); return 0;
  }
  return train_tc_layout(n_samples, n_importance, n_rays).total_bytes + 1024;
}

static int render_rays_bwd_tc(const SnerfRays* rays, const SnerfNetDesc* d, const void* bwd_coarse, const void* bwd_fine,
                              const SnerfOpts* o, const SnerfOutGrad* gout, const SnerfNetGradF32* gc,
                              const SnerfNetGradF32* gf, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!desc_is_flagship(d) || !bf16_geometry_supported(o->n_samples, o->n_importance)) {
    set_error("tensor-core training: unsupported network / sample counts"); return SNERF_ERR_UNSUPPORTED;
  }
  const TrainTcLayout L = train_tc_layout(o->n_samples, o->n_importance, rays->n_rays);
  if (workspace_bytes < L.total_bytes) {
    set_error("training workspace too small: %zu < %zu bytes", workspace_bytes, L.total_bytes);
    return SNERF_ERR_WORKSPACE;
  }
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const int Nc = o->n_samples, Nf = o->n_importance;
  TrainParams p{};
  p.ray_batch = rays->ray_batch; p.n_rays = rays->n_rays; p.width = rays->width; p.row_stride = rays->row_stride;
  p.Nc = Nc; p.Nf = Nf; p.white_bkgd = o->white_bkgd;
  p.noise0 = o->noise0; p.noise1 = o->noise1;
  p.raw_c = reinterpret_cast<float*>(ws + L.raw_c); p.raw_f = reinterpret_cast<float*>(ws + L.raw_f);
  p.z_c = reinterpret_cast<float*>(ws + L.z_c); p.z_f = reinterpret_cast<float*>(ws + L.z_f);
  p.g = *gout;
  p.draw4_c = reinterpret_cast<float4*>(ws + L.draw_c); p.draw4_f = reinterpret_cast<float4*>(ws + L.draw_f);
  if (int e = launch_composite_bwd_rows(p, stream)) return e;
  BwdTcParams b{};
  b.img[0] = (const unsigned char*)bwd_coarse; b.img[1] = (const unsigned char*)(bwd_fine ? bwd_fine : bwd_coarse);
  b.bits[0] = reinterpret_cast<const unsigned long long*>(ws + L.bits_c);
  b.bits[1] = reinterpret_cast<const unsigned long long*>(ws + L.bits_f);
  const unsigned char* const act[2] = {ws + L.act_c, ws + L.act_f};
  b.dz[0] = ws + L.dz_c; b.dz[1] = ws + L.dz_f;
  b.draw[0] = p.draw4_c; b.draw[1] = p.draw4_f;
  b.rows[0] = L.rows_c; b.rows[1] = L.rows_f;
  b.valid_rows[0] = rays->n_rays * Nc; b.valid_rows[1] = Nf > 0 ? rays->n_rays * (long long)(Nc + Nf) : 0;
  b.tiles[0] = (int)(L.rows_c / 128); b.tiles[1] = (int)(L.rows_f / 128);
  if (int e = launch_dx_chain_tc(b, stream)) return e;
  return launch_dw_tc(b, act, gc, bwd_fine ? gf : nullptr, stream);
}

int snerf_render_rays_bwd(const SnerfRays* rays, const SnerfNetDesc* d, const void* bwd_coarse, const void* bwd_fine,
                          const SnerfOpts* o, const SnerfOutGrad* gout, const SnerfNetGradF32* gc,
                          const SnerfNetGradF32* gf, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!rays || !o || !gout || !bwd_coarse || !gc || !workspace) { set_error("null argument"); return SNERF_ERR_BAD_ARG; }
  if (!desc_ok(d) || !train_supported(d, o && o->mode == SNERF_MODE_TF32)) return SNERF_ERR_UNSUPPORTED;
  if (rays->n_rays < 0 || (rays->n_rays > 0 && !rays->ray_batch)) { set_error("bad ray batch"); return SNERF_ERR_BAD_ARG; }
  if (o->n_samples < 2 || o->n_importance < 0 || o->n_samples + o->n_importance > kMaxSamples) {
    set_error("bad sample counts"); return SNERF_ERR_BAD_ARG;
  }
  if (bwd_fine && !gf) { set_error("grad_fine missing although a fine network is given"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  if (rays->n_rays == 0) return SNERF_OK;
  if (o->mode == SNERF_MODE_BF16)
    return render_rays_bwd_tc(rays, d, bwd_coarse, bwd_fine, o, gout, gc, gf, workspace, workspace_bytes, stream);
  const SnerfNetDesc* df = fine_desc(d, o->desc_fine);
  if (!df) return SNERF_ERR_BAD_ARG;
  if (!same_arch(d, df) && (o->mode != SNERF_MODE_FP32 || !bwd_fine)) {
    set_error("coarse and fine networks of different architectures train at the fp32 level only"); return SNERF_ERR_UNSUPPORTED;
  }
  const TrainLayout L = train_layout(d, df, o->n_samples, o->n_importance, rays->n_rays);
  if (workspace_bytes < L.total_floats * 4) {
    set_error("training workspace too small: %zu < %zu bytes", workspace_bytes, L.total_floats * 4);
    return SNERF_ERR_WORKSPACE;
  }
  float* ws = reinterpret_cast<float*>(workspace);
  TrainParams p{};
  p.ray_batch = rays->ray_batch; p.n_rays = rays->n_rays; p.width = rays->width; p.row_stride = rays->row_stride;
  p.Nc = o->n_samples; p.Nf = o->n_importance; p.white_bkgd = o->white_bkgd;
  p.TC = L.TC; p.TF = L.TF; p.Rc = L.Rc; p.Rf = L.Rf;
  p.noise0 = o->noise0; p.noise1 = o->noise1;
  p.raw_c = ws + L.raw_c; p.raw_f = ws + L.raw_f; p.z_c = ws + L.z_c; p.z_f = ws + L.z_f;
  p.g = *gout;
  p.save_c = ws + L.save_c; p.save_f = ws + L.save_f;
  p.dz_c = ws + L.dz_c - (long long)kSaveActCh * L.Rc;   // gradient stores skip the input channels
  p.dz_f = ws + L.dz_f - (long long)kSaveActCh * L.Rf;
  p.draw_c = ws + L.draw_c; p.draw_f = ws + L.draw_f;
  p.bwd_c = (const unsigned char*)bwd_coarse;
  p.bwd_f = (const unsigned char*)(bwd_fine ? bwd_fine : bwd_coarse);
  p.dw_tf32 = o->mode == SNERF_MODE_TF32 ? 1 : 0;
  return launch_train_backward(d, df, p, gc, bwd_fine ? gf : nullptr, stream);
}

int snerf_debug_dw_timing(int64_t* out_host, int32_t n_cta) {
  if (!out_host || n_cta <= 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  return debug_dw_timing(reinterpret_cast<long long*>(out_host), n_cta);
}

int snerf_debug_dw_cuts(int64_t rows_c, int64_t rows_f, int32_t n_cta, int64_t* out_cut, int64_t* out_first,
                        double* out_stream_units, double* makespan) {
  if (!out_cut || !out_first) { set_error("null argument"); return SNERF_ERR_BAD_ARG; }
  return debug_dw_cuts(rows_c, rows_f, n_cta, reinterpret_cast<long long*>(out_cut), reinterpret_cast<long long*>(out_first),
                       out_stream_units, makespan);
}

int snerf_query_network(const SnerfNetDesc* d, const void* packed, int mode, int multires, int multires_views,
                        const float* pts, const float* viewdirs, int64_t n_rays, int32_t n_samples, float* raw,
                        void* stream_) {
  if (!desc_ok(d)) return SNERF_ERR_BAD_ARG;
  if (!packed || !pts || !raw || n_rays < 0 || n_samples < 1) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (d->use_viewdirs && !viewdirs) { set_error("viewdirs missing"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  if (n_rays == 0) return SNERF_OK;
  RenderParams p{};
  SnerfOpts o{};
  o.multires = multires; o.multires_views = multires_views;
  if (int e = fill_common(p, d, &o)) return e;
  p.pts = pts; p.viewdirs = d->use_viewdirs ? viewdirs : nullptr; p.n_rays = n_rays; p.S = n_samples; p.out_raw = raw;
  p.img_coarse = p.img_fine = (const unsigned char*)packed;
  if (mode == SNERF_MODE_FP32) return launch_fp32(FE_QUERY, d->W, p, (cudaStream_t)stream_);
  if (mode == SNERF_MODE_BF16) return launch_bf16_query(p, (cudaStream_t)stream_);
  set_error("unknown mode %d", mode);
  return SNERF_ERR_BAD_ARG;
}

int snerf_nerf_forward(const SnerfNetDesc* d, const void* packed, int mode, const float* x, int64_t n_rows,
                       int32_t row_stride, float* out, void* stream_) {
  if (!desc_ok(d)) return SNERF_ERR_BAD_ARG;
  const int width = d->input_ch + (d->use_viewdirs ? d->input_ch_views : 0);
  if (!packed || !x || !out || n_rows < 0 || row_stride < width) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (mode != SNERF_MODE_FP32) { set_error("NeRF.forward on pre-encoded rows runs in fp32 mode only"); return SNERF_ERR_UNSUPPORTED; }
  if (int e = require_sm100()) return e;
  if (n_rows == 0) return SNERF_OK;
  RenderParams p{};
  p.x = x; p.n_rows = n_rows; p.x_stride = row_stride; p.in_ch = d->input_ch;
  p.in_ch_views = d->use_viewdirs ? d->input_ch_views : 0; p.out_raw = out;
  p.img_coarse = p.img_fine = (const unsigned char*)packed;
  return launch_fp32(FE_ROWS, d->W, p, (cudaStream_t)stream_);
}

int snerf_posenc(const float* x, int64_t n_rows, int32_t n_freqs, float* out, void* stream_) {
  if (!x || !out || n_rows < 0 || n_freqs < 0 || n_freqs > 16) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  if (n_rows == 0) return SNERF_OK;
  posenc_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream_>>>(x, n_rows, n_freqs, out);
  return check_cuda(cudaGetLastError(), "launch posenc_kernel");
}

int snerf_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, const float* noise, int64_t n_rays,
                        int32_t n_samples, int32_t white_bkgd, float* rgb_map, float* disp_map, float* acc_map,
                        float* weights, float* depth_map, void* stream_) {
  if (!raw || !z_vals || !rays_d || n_rays < 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (n_samples < 1 || n_samples > kMaxSamples) { set_error("n_samples=%d outside 1..%d", n_samples, kMaxSamples); return SNERF_ERR_UNSUPPORTED; }
  if ((reinterpret_cast<uintptr_t>(raw) & 15) != 0) { set_error("raw must be 16-byte aligned"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  if (n_rays == 0) return SNERF_OK;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > sm_count() * 16LL) blocks = sm_count() * 16LL;
  composite_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream_>>>(raw, z_vals, rays_d, noise, n_rays, n_samples,
                                                                       white_bkgd, rgb_map, disp_map, acc_map, weights,
                                                                       depth_map);
  return check_cuda(cudaGetLastError(), "launch composite_kernel");
}

int snerf_sample_pdf_fwd(const float* bins, const float* weights, const float* cdf_in, const float* u, int32_t u_per_ray,
                         int64_t n_rays, int32_t n_bins, int32_t n_out, float* samples, int64_t* inds, float* cdf_out,
                         void* stream_) {
  if (!bins || (!weights && !cdf_in) || !u || n_rays < 0 || n_out < 1) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (n_bins < 2 || n_bins > kMaxSamples) { set_error("n_bins=%d outside 2..%d", n_bins, kMaxSamples); return SNERF_ERR_UNSUPPORTED; }
  if (int e = require_sm100()) return e;
  if (n_rays == 0) return SNERF_OK;
  long long blocks = (n_rays + 3) / 4;
  if (blocks > sm_count() * 16LL) blocks = sm_count() * 16LL;
  sample_pdf_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream_>>>(bins, weights, cdf_in, u, u_per_ray, n_rays, n_bins,
                                                                        n_out, samples, (long long*)inds, cdf_out);
  return check_cuda(cudaGetLastError(), "launch sample_pdf_kernel");
}

int snerf_get_rays(int32_t H, int32_t W, float focal, const float* c2w_host, float cx, float cy, float* rays_o,
                   float* rays_d, void* stream_) {
  if (H < 1 || W < 1 || !c2w_host || !rays_o || !rays_d) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  Cam c;
  memcpy(c.m, c2w_host, sizeof(c.m));
  get_rays_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream_>>>(H, W, focal, c, cx, cy, rays_o, rays_d);
  return check_cuda(cudaGetLastError(), "launch get_rays_kernel");
}

// ---- hash-grid encoder (snerf_grid.cu) ----------------------------------------------------------------------
static int grid_layout_ok(const SnerfGridDesc* d, const void* base, int64_t sl, int64_t sb, const char* what) {
  // the kernels move the C channels of one (level, point) as one vector: base and strides must keep that aligned
  const int64_t esz = d->dtype == 0 ? 4 : 2;
  int64_t vec = d->C * esz;
  if (vec > 16) vec = 16;
  if (sl < d->C && d->L > 1) { set_error("%s: level stride %lld smaller than level_dim", what, (long long)sl); return SNERF_ERR_BAD_ARG; }
  if ((reinterpret_cast<uintptr_t>(base) % vec) || (sl * esz) % vec || (sb * esz) % vec) {
    set_error("%s: base pointer / strides must be multiples of %lld bytes", what, (long long)vec);
    return SNERF_ERR_BAD_ARG;
  }
  return SNERF_OK;
}

int snerf_grid_encode_fwd(const SnerfGridDesc* desc, const float* inputs, const void* embeddings, const int32_t* offsets,
                          void* outputs, int64_t out_stride_l, int64_t out_stride_b, void* dy_dx, int64_t n_points,
                          void* stream_) {
  if (int e = grid_check_desc(desc)) return e;
  if (n_points == 0) return SNERF_OK;  // an empty batch has null data pointers: nothing to do
  if (!inputs || !embeddings || !offsets || !outputs || n_points < 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (int e = grid_layout_ok(desc, outputs, out_stride_l, out_stride_b, "snerf_grid_encode_fwd(outputs)")) return e;
  if (reinterpret_cast<uintptr_t>(embeddings) % 16) { set_error("embeddings must be 16-byte aligned"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return grid_fwd(desc, inputs, embeddings, offsets, outputs, out_stride_l, out_stride_b, dy_dx, n_points, (cudaStream_t)stream_);
}

int snerf_grid_encode_bwd(const SnerfGridDesc* desc, const void* grad, int64_t grad_stride_l, int64_t grad_stride_b,
                          const float* inputs, const int32_t* offsets, void* grad_embeddings, const void* dy_dx,
                          void* grad_inputs, int64_t n_points, void* stream_) {
  if (int e = grid_check_desc(desc)) return e;
  if (n_points == 0) return SNERF_OK;
  if (!grad || !inputs || !offsets || !grad_embeddings || n_points < 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if ((dy_dx == nullptr) != (grad_inputs == nullptr)) { set_error("dy_dx and grad_inputs go together"); return SNERF_ERR_BAD_ARG; }
  if (int e = grid_layout_ok(desc, grad, grad_stride_l, grad_stride_b, "snerf_grid_encode_bwd(grad)")) return e;
  if (reinterpret_cast<uintptr_t>(grad_embeddings) % 16) { set_error("grad_embeddings must be 16-byte aligned"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return grid_bwd(desc, grad, grad_stride_l, grad_stride_b, inputs, offsets, grad_embeddings, dy_dx, grad_inputs, n_points,
                  (cudaStream_t)stream_);
}

int snerf_grid_grad_tv(const SnerfGridDesc* desc, const float* inputs, const void* embeddings, void* grad,
                       const int32_t* offsets, float weight, int64_t n_points, void* stream_) {
  if (int e = grid_check_desc(desc)) return e;
  if (n_points == 0) return SNERF_OK;
  if (!inputs || !embeddings || !grad || !offsets || n_points < 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (reinterpret_cast<uintptr_t>(embeddings) % 16 || reinterpret_cast<uintptr_t>(grad) % 16) {
    set_error("embeddings / grad must be 16-byte aligned"); return SNERF_ERR_BAD_ARG;
  }
  if (int e = require_sm100()) return e;
  return grid_tv(desc, inputs, embeddings, grad, offsets, weight, n_points, (cudaStream_t)stream_);
}

int snerf_grid_encode_ms_fwd(const SnerfGridDesc* desc, const float* means, const float* stds, float bound,
                             const void* embeddings, const int32_t* offsets, const int32_t* grid_sizes,
                             const float* level_gain, float* out, int64_t out_stride_n, int64_t n_samples,
                             int32_t n_multi, void* stream_) {
  if (int e = grid_check_desc(desc)) return e;
  if (n_samples == 0) return SNERF_OK;
  if (!means || !stds || !embeddings || !offsets || !grid_sizes || !out || n_samples < 0 || n_multi < 1 || !(bound > 0.f)) {
    set_error("bad argument"); return SNERF_ERR_BAD_ARG;
  }
  const int64_t width = (int64_t)desc->L * desc->C + (level_gain ? desc->L : 0);
  if (out_stride_n < width || (desc->C % 2 == 0 && (out_stride_n % 2 || reinterpret_cast<uintptr_t>(out) % 8))) {
    set_error("snerf_grid_encode_ms_fwd: out stride %lld must be >= %lld (and even, base 8-byte aligned, for even level_dim)",
              (long long)out_stride_n, (long long)width);
    return SNERF_ERR_BAD_ARG;
  }
  if (reinterpret_cast<uintptr_t>(embeddings) % 16) { set_error("embeddings must be 16-byte aligned"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return grid_ms_fwd(desc, means, stds, bound, embeddings, offsets, grid_sizes, level_gain, out, out_stride_n, n_samples,
                     n_multi, (cudaStream_t)stream_);
}

int snerf_grid_encode_ms_bwd(const SnerfGridDesc* desc, const float* grad, int64_t grad_stride_n, const float* means,
                             const float* stds, float bound, const int32_t* offsets, const int32_t* grid_sizes,
                             float* grad_embeddings, int64_t n_samples, int32_t n_multi, void* stream_) {
  if (int e = grid_check_desc(desc)) return e;
  if (n_samples == 0) return SNERF_OK;
  if (!grad || !means || !stds || !offsets || !grid_sizes || !grad_embeddings || n_samples < 0 || n_multi < 1 || !(bound > 0.f)) {
    set_error("bad argument"); return SNERF_ERR_BAD_ARG;
  }
  if (grad_stride_n < (int64_t)desc->L * desc->C || (desc->C % 2 == 0 && (grad_stride_n % 2 || reinterpret_cast<uintptr_t>(grad) % 8))) {
    set_error("snerf_grid_encode_ms_bwd: grad stride %lld must be >= L*C (and even, base 8-byte aligned, for even level_dim)",
              (long long)grad_stride_n);
    return SNERF_ERR_BAD_ARG;
  }
  if (reinterpret_cast<uintptr_t>(grad_embeddings) % 16) { set_error("grad_embeddings must be 16-byte aligned"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return grid_ms_bwd(desc, grad, grad_stride_n, means, stds, bound, offsets, grid_sizes, grad_embeddings, n_samples, n_multi,
                     (cudaStream_t)stream_);
}

int snerf_grid_level_gain(const SnerfGridDesc* desc, const void* embeddings, const int32_t* offsets, float init_std,
                          double* scratch, float* level_gain, void* stream_) {
  if (int e = grid_check_desc(desc)) return e;
  if (!embeddings || !offsets || !scratch || !level_gain) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (reinterpret_cast<uintptr_t>(embeddings) % 16) { set_error("embeddings must be 16-byte aligned"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return grid_level_gain(desc, embeddings, offsets, init_std, scratch, level_gain, (cudaStream_t)stream_);
}

int snerf_loss_fwd(const SnerfLossOpts* o, const float* rgb, const float* rgb0, const float* target, const float* depth,
                   const float* depth0, const float* target_depth, const float* confidence, int64_t n_rays, double* scratch,
                   float* out, void* stream_) {
  if (!o || !rgb || !target || !scratch || !out || n_rays <= 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (depth && (!depth0 || !target_depth)) { set_error("snerf_loss_fwd: the depth term needs depth0 and target_depth"); return SNERF_ERR_BAD_ARG; }
  if (confidence && !depth) { set_error("snerf_loss_fwd: confidence without a depth term"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return loss_fwd(o, rgb, rgb0, target, depth, depth0, target_depth, confidence, n_rays, scratch, out, (cudaStream_t)stream_);
}
int snerf_loss_bwd(const SnerfLossOpts* o, const float* rgb, const float* rgb0, const float* target, const float* depth,
                   const float* depth0, const float* target_depth, const float* confidence, int64_t n_rays,
                   const float* stats, const float* grad_loss, float* g_rgb, float* g_rgb0, float* g_depth, float* g_depth0,
                   float* g_confidence, void* stream_) {
  if (!o || !rgb || !target || !stats || !grad_loss || n_rays <= 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (depth && (!depth0 || !target_depth)) { set_error("snerf_loss_bwd: the depth term needs depth0 and target_depth"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return loss_bwd(o, rgb, rgb0, target, depth, depth0, target_depth, confidence, n_rays, stats, grad_loss, g_rgb, g_rgb0,
                  g_depth, g_depth0, g_confidence, (cudaStream_t)stream_);
}

int snerf_mip_encode(const SnerfMipEncode* e, void* stream_) {
  if (int err = require_sm100()) return err;
  return mip_encode(e, (cudaStream_t)stream_);
}
int snerf_linear_tc(const SnerfLinear* l, void* stream_) {
  if (int err = require_sm100()) return err;
  return linear_tc(l, (cudaStream_t)stream_);
}
int snerf_rows_to_bf16(const float* x, int64_t rows, int32_t row_stride, int32_t col0, int32_t ncols, int32_t repeat, void* out,
                       int32_t out_cols, int64_t m_pad, void* stream_) {
  if (int err = require_sm100()) return err;
  return rows_to_bf16(x, rows, row_stride, col0, ncols, repeat, out, out_cols, m_pad, (cudaStream_t)stream_);
}
int snerf_mip_cond_bias(const float* viewdirs, int64_t n_rays, int32_t deg_view, const float* w, int32_t ldw, int32_t k0,
                        const float* b, int32_t n_out, float* out, void* stream_) {
  if (int err = require_sm100()) return err;
  return mip_cond_bias(viewdirs, n_rays, deg_view, w, ldw, k0, b, n_out, out, (cudaStream_t)stream_);
}
int snerf_mip_composite(const SnerfMipComposite* c, void* stream_) {
  if (int err = require_sm100()) return err;
  return mip_composite(c, (cudaStream_t)stream_);
}

int snerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t* step, void* stream_) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !lr || !step || n < 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return adam_step(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, reinterpret_cast<long long*>(step),
                   (cudaStream_t)stream_);
}

int snerf_proposal_loss(const float* s_vals_f, const float* weights_f, const float* s_vals_c, const float* weights_c,
                        int64_t n_rays, int32_t n_fine, int32_t n_coarse, float weight, double* scratch, float* loss_out,
                        float* grad_weights_c, void* stream_) {
  if (!s_vals_f || !weights_f || !s_vals_c || !weights_c || !scratch || !loss_out || n_rays <= 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return proposal_loss(s_vals_f, weights_f, s_vals_c, weights_c, n_rays, n_fine, n_coarse, weight, scratch, loss_out,
                       grad_weights_c, (cudaStream_t)stream_);
}

int snerf_stepfun_resample(const SnerfStepfunOpts* o, const float* t, const float* w, int64_t n_rays, int32_t n_bins,
                           const float* u_base, const float* jitter, int32_t jitter_cols, int32_t n_samples, float* out,
                           float* centers, float* t_dilate, float* w_dilate, void* stream_) {
  if (!o) { set_error("null options"); return SNERF_ERR_BAD_ARG; }
  if (n_rays == 0) return SNERF_OK;
  if (!t || !w || n_rays < 0) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (n_samples > 0 && (!u_base || (!out && !centers))) { set_error("snerf_stepfun_resample: u_base and an output are required when n_samples > 0"); return SNERF_ERR_BAD_ARG; }
  if (jitter && jitter_cols != 1 && jitter_cols != n_samples) { set_error("snerf_stepfun_resample: jitter must have 1 or n_samples columns"); return SNERF_ERR_BAD_ARG; }
  if ((t_dilate || w_dilate) && !o->dilate) { set_error("snerf_stepfun_resample: dilated outputs requested without dilate"); return SNERF_ERR_BAD_ARG; }
  if (o->dilate && o->weights_are_logits) { set_error("snerf_stepfun_resample: dilation works on weights, not logits"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return stepfun_resample(t, w, n_rays, n_bins, o->dilate, o->renormalize, o->weights_are_logits, o->dilation, o->domain_lo,
                          o->domain_hi, o->anneal, o->resample_padding, u_base, jitter, jitter_cols, o->max_jitter, n_samples,
                          out, centers, t_dilate, w_dilate, (cudaStream_t)stream_);
}

int snerf_selftest_umma(const float* a, const float* b, float* d, int32_t variant, void* stream_) {
  if (!a || !b || !d || variant < 0 || variant > 1) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  if (int e = require_sm100()) return e;
  return launch_selftest_umma(a, b, d, variant, (cudaStream_t)stream_);
}

}  // extern "C"
