// C ABI of the m6anet MIL-inference hot path (declared in include/m6anet_b200.h).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/m6anet_b200.h"
#include "m6a_kernel.h"

using namespace m6a;

static_assert(M6A_MAX_SAMPLES == kMaxSamples && M6A_POOL_PROD == kPoolProd && M6A_POOL_MEAN == kPoolMean &&
                  M6A_POOL_MAX == kPoolMax && M6A_N_SIG == kNSig && M6A_H2 == kH2 && M6A_H1_MAX == kH1Max,
              "include/m6anet_b200.h and the kernel limits disagree");

extern "C" int64_t m6a_mil_workspace_bytes(int64_t total_reads);
constexpr int kHostSlots = 3;
struct HostSlot {   // one stage of the host-buffer pipeline (m6a_mil_infer_host_f32)
  cudaStream_t stream = nullptr;
  void *d_feats = nullptr, *d_off = nullptr, *d_kmer = nullptr, *d_rp = nullptr, *d_sp = nullptr, *d_mc = nullptr, *d_ws = nullptr;
  void* d_bag = nullptr;      // per-pass bag probabilities (m6a_mil_validate_host_f32 only)
  size_t cap_feats = 0, cap_off = 0, cap_kmer = 0, cap_rp = 0, cap_sp = 0, cap_mc = 0, cap_ws = 0, cap_bag = 0;
  int64_t* h_off = nullptr;   // pinned staging for re-based offsets
  size_t cap_hoff = 0;
};

struct m6a_model {
  DeviceModel dev;
  WeightImage host_image;        // packed weights, host copy (kernel-parameter path)
  void* d_image;
  void* d_ctab;
  void* d_tc_image = nullptr;    // tcx::WeightImageTc (tensor-core read encoder), nullptr when the model does not fit it
  int encoder = M6A_ENCODER_DEFAULT;
  int* trap_record = nullptr;    // mapped host memory: which bounded wait of the tensor-core kernel gave up (debug)
  int device;
  int n_sms;
  int tile_reads = 0;            // feature rows per tile; 0 = automatic (m6a_model_set_tile_reads)
  HostSlot slots[kHostSlots];
  std::mutex ws_mutex;
};
void m6a_release_workspace(m6a_model* m);

static thread_local LaunchInfo g_last = {0, 0, 0, 0};
static thread_local int g_last_launches = 0;

extern "C" int m6a_version(void) { return M6A_VERSION; }

extern "C" int m6a_device_count(int32_t* count) {
  if (!count) return M6A_EINVAL;
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  *count = (e == cudaSuccess) ? n : 0;
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}

extern "C" int m6a_set_device(int32_t device) {
  const cudaError_t e = cudaSetDevice(device);
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}

extern "C" const char* m6a_strerror(int status) {
  switch (status) {
    case M6A_OK: return "ok";
    case M6A_EINVAL: return "invalid argument";
    case M6A_EUNSUPPORTED: return "model dimensions not supported by the compiled kernel";
    case M6A_EALIGN: return "buffer alignment";
    case M6A_ERANGE: return "value out of range for this mode";
    case M6A_ENOMEM: return "host allocation failed";
    case M6A_EPARSE: return "data.json site line does not match data.info or the expected format";
    case M6A_EIO: return "file open/read/write failed";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown m6anet_b200 status";
}

#define M6A_CUDA(expr)                                  \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) return static_cast<int>(_e); \
  } while (0)

extern "C" int m6a_model_create(const m6a_weights_t* w, m6a_model_t** out) {
  if (w == nullptr || out == nullptr) return M6A_EINVAL;
  *out = nullptr;
  if (!w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->w3 || !w->b3) return M6A_EINVAL;
  if (w->n_sig != kNSig || w->h2 != kH2 || w->h1 < 1 || w->h1 > kH1Max) return M6A_EUNSUPPORTED;
  if (w->emb_dim < 0 || w->emb_dim > 64) return M6A_EUNSUPPORTED;
  if (w->emb_dim > 0 && (w->emb == nullptr || w->n_kmer < 1 || w->n_kmer > 4096)) return M6A_EINVAL;

  const int h1 = w->h1, E = w->emb_dim, in1 = kNSig + kKmerPos * E;
  const int n_kmer = E > 0 ? w->n_kmer : 1;

  WeightImage* img = static_cast<WeightImage*>(calloc(1, sizeof(WeightImage)));
  if (!img) return M6A_ENOMEM;
  for (int j = 0; j < h1; ++j) {
    float* row = img->pair[j >> 1];
    const int half = j & 1;
    for (int k = 0; k < kNSig; ++k) row[2 * k + half] = w->w1[static_cast<size_t>(j) * in1 + k];
    for (int k = 0; k < kH2; ++k) row[(half ? kW2Off1 : kW2Off0) + k] = w->w2[static_cast<size_t>(k) * h1 + j];
  }
  img->n_pairs = (h1 + 1) / 2;
  for (int k = 0; k < kH2; ++k) {
    img->b2[k] = w->b2[k];
    img->w3[k] = w->w3[k];
  }
  img->b3 = w->b3[0];
  img->h1 = h1;

  // ctab[t][kmer][j]: per k-mer-position contribution of the embedding to Linear-1 (table 0 carries b1)
  std::vector<float> ctab(static_cast<size_t>(kKmerPos) * n_kmer * kH1Max, 0.0f);
  for (int t = 0; t < kKmerPos; ++t)
    for (int k = 0; k < n_kmer; ++k)
      for (int j = 0; j < h1; ++j) {
        float c = (t == 0) ? w->b1[j] : 0.0f;
        for (int d = 0; d < E; ++d)
          c = fmaf(w->w1[static_cast<size_t>(j) * in1 + kNSig + t * E + d], w->emb[static_cast<size_t>(k) * E + d], c);
        ctab[(static_cast<size_t>(t) * n_kmer + k) * kH1Max + j] = c;
      }

  // tensor-core image: RN_tf32 split of [w1 | b1] and of w2 (m6a_layout.h); emb_dim 0..2, h1 <= 160
  tcx::WeightImageTc* tci = nullptr;
  if (E <= 2 && h1 <= tcx::kN1 && n_kmer * std::max(E, 1) <= tcx::kEmbMax) {
    tci = static_cast<tcx::WeightImageTc*>(calloc(1, sizeof(tcx::WeightImageTc)));
    if (!tci) {
      free(img);
      return M6A_ENOMEM;
    }
    auto rn_tf32 = [](float f) {
      uint32_t u;
      memcpy(&u, &f, 4);
      u = (u + 0x1000u) & 0xFFFFE000u;
      memcpy(&f, &u, 4);
      return f;
    };
    for (int j = 0; j < h1; ++j)
      for (int k = 0; k < tcx::kK1; ++k) {
        float v = 0.0f;
        if (k < in1) v = w->w1[static_cast<size_t>(j) * in1 + k];
        else if (k == tcx::kK1 - 1) v = w->b1[j];
        const float hi = rn_tf32(v);
        tci->w1hi[k / 4][j][k % 4] = hi;
        tci->w1lo[k / 4][j][k % 4] = rn_tf32(v - hi);
      }
    for (int n = 0; n < kH2; ++n)
      for (int k = 0; k < h1; ++k) {
        const float v = w->w2[static_cast<size_t>(n) * h1 + k];
        const float hi = rn_tf32(v);
        tci->w2s[k / 4][n][k % 4] = hi;
        tci->w2s[k / 4][kH2 + n][k % 4] = rn_tf32(v - hi);
      }
    for (int k = 0; k < kH2; ++k) {
      tci->b2[k] = w->b2[k];
      tci->w3[k] = w->w3[k];
    }
    tci->b3 = w->b3[0];
    tci->emb_dim = E;
    tci->n_kmer = n_kmer;
    for (int i = 0; i < n_kmer * E; ++i) tci->emb[i] = w->emb[i];
  }

  m6a_model* m = new (std::nothrow) m6a_model();
  if (!m) {
    free(img);
    free(tci);
    return M6A_ENOMEM;
  }
  cudaError_t e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->n_sms, cudaDevAttrMultiProcessorCount, m->device);
  m->d_image = nullptr;
  m->d_ctab = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&m->d_image, sizeof(WeightImage));
  if (e == cudaSuccess) e = cudaMalloc(&m->d_ctab, ctab.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(m->d_image, img, sizeof(WeightImage), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_ctab, ctab.data(), ctab.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && tci) e = cudaMalloc(&m->d_tc_image, sizeof(tcx::WeightImageTc));
  if (e == cudaSuccess && tci) e = cudaMemcpy(m->d_tc_image, tci, sizeof(tcx::WeightImageTc), cudaMemcpyHostToDevice);
  m->host_image = *img;
  free(img);
  free(tci);
  if (const char* env = getenv("M6A_ENCODER")) {      // build-independent A/B switch: "ffma" | "tc"
    if (!strcmp(env, "ffma")) m->encoder = M6A_ENCODER_FFMA;
    else if (!strcmp(env, "tc")) m->encoder = M6A_ENCODER_TC;
  }
  if (e != cudaSuccess) {
    if (m->d_tc_image) cudaFree(m->d_tc_image);
    if (m->d_image) cudaFree(m->d_image);
    if (m->d_ctab) cudaFree(m->d_ctab);
    delete m;
    return static_cast<int>(e);
  }
  m->dev.image = static_cast<const WeightImage*>(m->d_image);
  m->dev.ctab = static_cast<const float*>(m->d_ctab);
  m->dev.n_kmer = n_kmer;
  m->dev.h1 = h1;
  m->dev.emb_dim = E;
  *out = m;
  return M6A_OK;
}

extern "C" int m6a_model_set_tile_reads(m6a_model_t* model, int32_t tile_reads) {
  if (!model) return M6A_EINVAL;
  if (tile_reads != 0 && (tile_reads < 64 || tile_reads > kQCap)) return M6A_EINVAL;   // (both kernels accept 64..4096)
  model->tile_reads = tile_reads;
  return M6A_OK;
}

extern "C" int m6a_model_set_encoder(m6a_model_t* model, int32_t encoder) {
  if (!model || (encoder != M6A_ENCODER_FFMA && encoder != M6A_ENCODER_TC)) return M6A_EINVAL;
  if (encoder == M6A_ENCODER_TC && !model->d_tc_image) return M6A_EUNSUPPORTED;
  model->encoder = encoder;
  return M6A_OK;
}

extern "C" int m6a_model_get_encoder(const m6a_model_t* model) {
  if (!model) return M6A_EINVAL;
  return (model->encoder == M6A_ENCODER_TC && model->d_tc_image) ? M6A_ENCODER_TC : M6A_ENCODER_FFMA;
}

// Debug aid of the tensor-core kernel: every mbarrier wait is bounded; when one gives up it records
// {wait site, block, thread, parity} in mapped host memory before it traps, readable after the context is lost.
extern "C" int m6a_debug_trap_record(m6a_model_t* model, int32_t* out4) {
  if (!model) return M6A_EINVAL;
  if (!model->trap_record) {
    M6A_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&model->trap_record), 256, cudaHostAllocMapped));
    memset(model->trap_record, 0, 256);
    int* dptr = nullptr;
    M6A_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), model->trap_record, 0));
    M6A_CUDA(tc_set_trap_record(dptr));
  }
  if (out4)       // 16 wait sites x {site, block, thread, parity}; the caller's buffer holds 64 int32
    for (int i = 0; i < 64; ++i) out4[i] = model->trap_record[i];
  return M6A_OK;
}

// Debug aid (builds with -DM6A_TC_PROFILE=1): cycles per phase of one thread per role of block 0 of the last launch.
extern "C" int m6a_debug_tc_profile(uint64_t* out32) {
  if (!out32) return M6A_EINVAL;
  M6A_CUDA(cudaDeviceSynchronize());
  M6A_CUDA(tc_read_profile(reinterpret_cast<unsigned long long*>(out32)));
  return M6A_OK;
}

extern "C" int m6a_pinned_alloc(void** out, int64_t bytes) {
  if (!out || bytes < 0) return M6A_EINVAL;
  *out = nullptr;
  if (bytes == 0) return M6A_OK;
  const cudaError_t e = cudaHostAlloc(out, static_cast<size_t>(bytes), cudaHostAllocPortable);
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}

extern "C" int m6a_pinned_free(void* p) {
  if (!p) return M6A_OK;
  const cudaError_t e = cudaFreeHost(p);
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}

extern "C" const char* m6a_build_info(void) {
  return "m6anet_b200 r2: mil_infer_tc_kernel<20> (tcgen05 3xTF32 encoder, warp-specialised, 1 CTA/SM) + "
         "mil_infer_kernel<NS,BAGS> (FFMA2 encoder, dynamic tiles)";
}

extern "C" int m6a_model_destroy(m6a_model_t* model) {
  if (!model) return M6A_OK;
  m6a_release_workspace(model);
  cudaFree(model->d_image);
  cudaFree(model->d_ctab);
  cudaFree(model->d_tc_image);
  if (model->trap_record) cudaFreeHost(model->trap_record);
  delete model;
  return M6A_OK;
}

// Rows per tile when the caller did not fix it.  Tiles are cut at multiples of T rows (tile_bounds_kernel), so for
// constant-depth sites T must be a multiple of the depth or every other tile spills a few rows into an extra chunk
// (1 M x 50 reads: T = 500 or 1000 -> 18.2 ms, T = 512 -> 20.1 ms); uneven sites want ~1000 rows (19.5 ms vs 22.7 ms at
// 512).  Small jobs use ~500 to keep at least ~48 tiles per CTA (tail imbalance of the persistent grid).
static int auto_tile_reads(long long n_sites, long long total_reads, int n_sms) {
  long long base = kChunkReads * 2 - 24;   // 1000 with 512-row chunks
  if (total_reads / base < 48ll * n_sms * kCtasPerSm) base = kChunkReads - 12;   // 500
  if (n_sites > 0 && total_reads % n_sites == 0) {
    const long long depth = total_reads / n_sites;
    if (depth >= 1 && depth <= base) base = (base / depth) * depth;
  }
  return static_cast<int>(std::max<long long>(64, std::min<long long>(base, kQCap)));
}

// Rows per tile for the tensor-core kernel: slabs are slices of <= 64 sites, encoded 128 rows at a time, so a tile of
// 64 sites' worth of rows (a multiple of 128 for every even site depth) wastes no MMA rows; <= 6144 rows (q table).
// Tiles are dealt round-robin to one CTA per SM: when the job is only a few rounds long (one shard of an 8-GPU run: 13.2
// rounds of 3200 rows), the tile is shrunk to total / (ceil(rounds) * SMs), rounded up to whole MMA tiles, so that the last
// round is as full as the others (14 rounds of 3072 rows instead of 14 of 3200).
static int auto_tile_reads_tc(long long n_sites, long long total_reads, int n_sms) {
  long long base = 2048;
  if (n_sites > 0 && total_reads % n_sites == 0) {
    const long long depth = total_reads / n_sites;
    if (depth >= 1 && depth * kSitesPerTileMax <= kTcQCap) base = depth * kSitesPerTileMax;
    else if (depth >= 1 && depth <= kTcQCap) base = (kTcQCap / depth) * depth;
  }
  base = std::max<long long>(64, std::min<long long>(base, kTcQCap));
  const long long grid = std::max(1, n_sms);
  const long long rounds = (total_reads + base * grid - 1) / (base * grid);
  if (rounds >= 1 && rounds < 64) {
    long long t = (total_reads + rounds * grid - 1) / (rounds * grid);
    t = (t + tcx::kTileM - 1) / tcx::kTileM * tcx::kTileM;
    if (t >= 256 && t < base) base = t;
  }
  return static_cast<int>(base);
}

static int infer_device_impl(const m6a_model_t* model, int tile_reads, const float* feats, const int64_t* read_off,
                             const int32_t* kmer_idx, int64_t n_sites, int64_t total_reads, int64_t site_id_base,
                             int32_t n_samples, int32_t n_iters, uint64_t seed, const uint16_t* sample_idx,
                             float read_threshold, float* read_prob, float* site_prob, int32_t* mod_count,
                             void* workspace, int64_t workspace_bytes, void* stream, const BagArgs* bags = nullptr,
                             int site_stride = 1) {
  if (!model || n_sites < 0 || total_reads < 0) return M6A_EINVAL;
  if (n_samples < 1 || n_samples > kMaxSamples || n_iters < 1) return M6A_EINVAL;
  if (bags && (bags->pool < kPoolProd || bags->pool > kPoolMax || (bags->replace != 0 && bags->replace != 1))) return M6A_EINVAL;
  if (n_sites == 0) return M6A_OK;
  if (!read_off || !site_prob || !mod_count) return M6A_EINVAL;
  if (total_reads > 0 && (!feats || !read_prob)) return M6A_EINVAL;
  if (model->dev.emb_dim > 0 && !kmer_idx) return M6A_EINVAL;
  if ((reinterpret_cast<uintptr_t>(feats) & 3u) || (reinterpret_cast<uintptr_t>(read_off) & 7u)) return M6A_EALIGN;

  KernelArgs a;
  memset(&a, 0, sizeof(a));
  a.model = model->dev;
  a.feats = feats;
  a.read_off = read_off;
  a.kmer_idx = kmer_idx;
  a.sample_idx = sample_idx;
  a.read_prob = read_prob;
  a.site_prob = site_prob;
  a.mod_count = mod_count;
  a.n_sites = n_sites;
  // read-balanced tiles: tile t = sites whose first row lies in [t*T, (t+1)*T); boundaries by a prepass into the
  // caller's workspace (the library allocates nothing on this path)
  const bool use_tc = model->encoder == M6A_ENCODER_TC && model->d_tc_image != nullptr && bags == nullptr &&
                      sample_idx == nullptr && n_samples == 20;
  if (tile_reads <= 0 || (use_tc && model->tile_reads <= 0)) {
    tile_reads = use_tc ? auto_tile_reads_tc(n_sites, total_reads, model->n_sms) : auto_tile_reads(n_sites, total_reads, model->n_sms);
  }
  a.tile_reads = tile_reads;
  a.site_stride = site_stride;
  a.n_tiles = total_reads / tile_reads + 1;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 7u)) return workspace ? M6A_EALIGN : M6A_EINVAL;
  if (workspace_bytes < static_cast<int64_t>((a.n_tiles + 2) * sizeof(long long))) return M6A_EINVAL;   // bounds + tile counter
  long long* d_bounds = static_cast<long long*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  a.tile_bounds = d_bounds;
  a.site_id_base = site_id_base;
  a.feats_bytes = static_cast<unsigned long long>(total_reads) * (kNSig * sizeof(float));
  a.seed = seed;
  a.n_samples = n_samples;
  a.n_iters = n_iters;
  block_layout(n_iters, &a.iters_per_lane, &a.n_blocks);
  a.read_threshold = read_threshold;
  a.feats_tma_ok = (reinterpret_cast<uintptr_t>(feats) & 15u) == 0;

  LaunchInfo info;
  cudaError_t e = launch_tile_bounds(read_off, n_sites, a.n_tiles, tile_reads, d_bounds, st);
  if (e == cudaSuccess) {
    e = use_tc ? launch_mil_infer_tc(a, static_cast<const tcx::WeightImageTc*>(model->d_tc_image), model->n_sms, st, &info)
               : launch_mil_infer(a, &model->host_image, bags, model->n_sms, st, &info);
  }
  if (e != cudaSuccess) return static_cast<int>(e);
  g_last = info;
  g_last_launches = 2;   // tile_bounds_kernel + mil_infer_kernel
  return M6A_OK;
}

extern "C" int m6a_mil_infer_f32(const m6a_model_t* model, const float* feats, const int64_t* read_off,
                                 const int32_t* kmer_idx, int64_t n_sites, int64_t total_reads, int64_t site_id_base,
                                 int32_t n_samples, int32_t n_iters, uint64_t seed, const uint16_t* sample_idx,
                                 float read_threshold, float* read_prob, float* site_prob, int32_t* mod_count,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  if (!model) return M6A_EINVAL;
  return infer_device_impl(model, model->tile_reads, feats, read_off, kmer_idx, n_sites, total_reads, site_id_base, n_samples,
                           n_iters, seed, sample_idx, read_threshold, read_prob, site_prob, mod_count, workspace,
                           workspace_bytes, stream);
}

extern "C" int m6a_mil_infer_packed_f32(const m6a_model_t* model, const float* feats, const int64_t* read_off,
                                        const int32_t* kmer_idx, int64_t n_sites, int64_t total_reads, int64_t site_id_base,
                                        int32_t n_samples, int32_t n_iters, uint64_t seed, float read_threshold,
                                        float* read_prob, void* site_out, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  if (!model || (n_sites > 0 && !site_out)) return M6A_EINVAL;
  if (reinterpret_cast<uintptr_t>(site_out) & 7u) return M6A_EALIGN;
  return infer_device_impl(model, model->tile_reads, feats, read_off, kmer_idx, n_sites, total_reads, site_id_base, n_samples,
                           n_iters, seed, nullptr, read_threshold, read_prob, static_cast<float*>(site_out),
                           static_cast<int32_t*>(site_out) + 1, workspace, workspace_bytes, stream, nullptr, 2);
}

// validate()-style literal MIL forward (reference utils/training_utils.py:213-268): same read encoder, phase B pools one
// bag per (site, pass) instead of the Monte-Carlo noisy-OR
extern "C" int m6a_mil_validate_f32(const m6a_model_t* model, const float* feats, const int64_t* read_off,
                                    const int32_t* kmer_idx, int64_t n_sites, int64_t total_reads, int64_t site_id_base,
                                    int32_t n_samples, int32_t n_iters, uint64_t seed, const uint16_t* sample_idx,
                                    int32_t pooling, int32_t replace, float read_threshold, float* read_prob,
                                    float* bag_prob, float* site_prob, int32_t* mod_count, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  if (!model) return M6A_EINVAL;
  const BagArgs bags = {bag_prob, replace, pooling};
  return infer_device_impl(model, model->tile_reads, feats, read_off, kmer_idx, n_sites, total_reads, site_id_base, n_samples,
                           n_iters, seed, sample_idx, read_threshold, read_prob, site_prob, mod_count, workspace,
                           workspace_bytes, stream, &bags);
}

extern "C" int32_t m6a_auto_tile_reads(int64_t n_sites, int64_t total_reads, int32_t n_sms) {
  return auto_tile_reads(n_sites, total_reads, n_sms > 0 ? n_sms : 148);
}

extern "C" int32_t m6a_auto_tile_reads_tc(int64_t n_sites, int64_t total_reads, int32_t n_sms) {
  return auto_tile_reads_tc(n_sites, total_reads, n_sms > 0 ? n_sms : 148);
}

extern "C" int64_t m6a_mil_workspace_bytes(int64_t total_reads) {
  if (total_reads < 0) return 0;
  return (total_reads / 64 + 3) * static_cast<int64_t>(sizeof(long long));   // tiles hold >= 64 rows: n_tiles + 1 bounds + 1 counter
}

extern "C" int m6a_sample_indices(uint64_t seed, int64_t site_id, int32_t n_reads, int32_t n_iters, int32_t n_samples,
                                  int32_t* out, void* stream) {
  if (!out || n_reads < 1 || n_iters < 1 || n_samples < 1) return M6A_EINVAL;
  cudaError_t e = launch_sample_indices(seed, static_cast<uint64_t>(site_id), static_cast<uint32_t>(n_reads), n_iters,
                                        n_samples, false, out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}

extern "C" int m6a_sample_bags(uint64_t seed, int64_t site_id, int32_t n_reads, int32_t n_iters, int32_t n_samples,
                               int32_t* out, void* stream) {
  if (!out || n_iters < 1 || n_samples < 1 || n_samples > kMaxSamples) return M6A_EINVAL;
  if (n_reads < n_samples) return M6A_ERANGE;   // no bag of n_samples distinct reads
  cudaError_t e = launch_sample_indices(seed, static_cast<uint64_t>(site_id), static_cast<uint32_t>(n_reads), n_iters,
                                        n_samples, true, out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? M6A_OK : static_cast<int>(e);
}

extern "C" int m6a_last_launch(int32_t* grid, int32_t* block, int32_t* smem_bytes, int32_t* tile_reads,
                               int32_t* n_launches) {
  if (grid) *grid = g_last.grid;
  if (block) *block = g_last.block;
  if (smem_bytes) *smem_bytes = g_last.smem_bytes;
  if (tile_reads) *tile_reads = g_last.tile_reads;
  if (n_launches) *n_launches = g_last_launches;
  return M6A_OK;
}

// ---- host-buffer path: chunked, stream-pipelined H2D -> kernel -> D2H --------------------------------
// The pipeline owns a small persistent workspace inside the model handle (3 slots of device buffers, one
// stream each, pinned staging for the re-based CSR offsets).  It only grows, so steady-state calls do no
// allocation at all; m6a_model_destroy releases it.
static cudaError_t ensure_bytes(void** p, size_t* cap, size_t need) {
  if (need <= *cap) return cudaSuccess;
  if (*p) {
    cudaError_t e = cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    if (e != cudaSuccess) return e;
  }
  const size_t want = need + need / 8 + 256;
  cudaError_t e = cudaMalloc(p, want);
  if (e == cudaSuccess) *cap = want;
  return e;
}

static cudaError_t slot_reserve(HostSlot& sl, int64_t max_sites, int64_t max_reads, int64_t bag_floats_per_site) {
  cudaError_t e = cudaSuccess;
  if (!sl.stream) e = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = ensure_bytes(&sl.d_feats, &sl.cap_feats, static_cast<size_t>(max_reads) * kNSig * sizeof(float) + 16);
  if (e == cudaSuccess) e = ensure_bytes(&sl.d_rp, &sl.cap_rp, static_cast<size_t>(max_reads) * sizeof(float) + 16);
  if (e == cudaSuccess) e = ensure_bytes(&sl.d_off, &sl.cap_off, static_cast<size_t>(max_sites + 1) * sizeof(int64_t));
  if (e == cudaSuccess) e = ensure_bytes(&sl.d_kmer, &sl.cap_kmer, static_cast<size_t>(max_sites) * kKmerPos * sizeof(int32_t));
  if (e == cudaSuccess) e = ensure_bytes(&sl.d_sp, &sl.cap_sp, static_cast<size_t>(max_sites) * sizeof(float));
  if (e == cudaSuccess) e = ensure_bytes(&sl.d_mc, &sl.cap_mc, static_cast<size_t>(max_sites) * sizeof(int32_t));
  if (e == cudaSuccess) e = ensure_bytes(&sl.d_ws, &sl.cap_ws, static_cast<size_t>(m6a_mil_workspace_bytes(max_reads)));
  if (e == cudaSuccess && bag_floats_per_site > 0)
    e = ensure_bytes(&sl.d_bag, &sl.cap_bag, static_cast<size_t>(max_sites) * bag_floats_per_site * sizeof(float));
  if (e == cudaSuccess && static_cast<size_t>(max_sites + 1) > sl.cap_hoff) {
    if (sl.h_off) cudaFreeHost(sl.h_off);
    sl.h_off = nullptr;
    sl.cap_hoff = 0;
    const size_t want = static_cast<size_t>(max_sites + 1) + static_cast<size_t>(max_sites) / 8 + 16;
    e = cudaMallocHost(reinterpret_cast<void**>(&sl.h_off), want * sizeof(int64_t));
    if (e == cudaSuccess) sl.cap_hoff = want;
  }
  return e;
}

void m6a_release_workspace(m6a_model* m) {
  for (int s = 0; s < kHostSlots; ++s) {
    HostSlot& sl = m->slots[s];
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    cudaFree(sl.d_feats); cudaFree(sl.d_off); cudaFree(sl.d_kmer);
    cudaFree(sl.d_rp); cudaFree(sl.d_sp); cudaFree(sl.d_mc); cudaFree(sl.d_ws); cudaFree(sl.d_bag);
    if (sl.h_off) cudaFreeHost(sl.h_off);
    if (sl.stream) cudaStreamDestroy(sl.stream);
    sl = HostSlot();
  }
}

// host_bags: nullptr = inference; otherwise the validate()-style bags (bag_prob is then a HOST pointer [n_sites, n_iters] or NULL)
static int infer_host_impl(const m6a_model_t* model_c, const float* feats, const int64_t* read_off,
                           const int32_t* kmer_idx, int64_t n_sites, int64_t site_id_base, int32_t n_samples,
                           int32_t n_iters, uint64_t seed, float read_threshold, float* read_prob, float* site_prob,
                           int32_t* mod_count, int32_t n_chunks, const BagArgs* host_bags) {
  m6a_model* model = const_cast<m6a_model*>(model_c);
  if (!model || n_sites < 0) return M6A_EINVAL;
  if (n_sites == 0) return M6A_OK;
  if (!read_off || !site_prob || !mod_count) return M6A_EINVAL;
  if (read_off[0] != 0) return M6A_EINVAL;
  const int64_t total_reads = read_off[n_sites];
  if (total_reads < 0) return M6A_EINVAL;
  if (total_reads > 0 && (!feats || !read_prob)) return M6A_EINVAL;
  if (model->dev.emb_dim > 0 && !kmer_idx) return M6A_EINVAL;
  if (n_samples < 1 || n_samples > kMaxSamples || n_iters < 1) return M6A_EINVAL;
  float* h_bag = host_bags ? host_bags->bag_prob : nullptr;

  // chunk boundaries: balanced by reads, cut at site boundaries (~32 MB of features per chunk by default)
  if (n_chunks <= 0) {
    const int64_t bytes = total_reads * kNSig * 4;
    n_chunks = static_cast<int>(std::min<int64_t>(256, std::max<int64_t>(1, bytes / (32ll << 20))));
    if (n_chunks > 1 && n_chunks < 4) n_chunks = 4;
  }
  if (n_chunks > n_sites) n_chunks = static_cast<int>(n_sites);
  std::vector<int64_t> cut(n_chunks + 1, 0);
  cut[n_chunks] = n_sites;
  for (int c = 1; c < n_chunks; ++c) {
    const int64_t target = total_reads * c / n_chunks;
    const int64_t* p = std::lower_bound(read_off + cut[c - 1], read_off + n_sites, target);
    cut[c] = std::max<int64_t>(cut[c - 1], std::min<int64_t>(p - read_off, n_sites));
  }
  int64_t max_sites = 0, max_reads = 0;
  for (int c = 0; c < n_chunks; ++c) {
    max_sites = std::max(max_sites, cut[c + 1] - cut[c]);
    max_reads = std::max(max_reads, read_off[cut[c + 1]] - read_off[cut[c]]);
  }

  std::lock_guard<std::mutex> guard(model->ws_mutex);
  const int tile_reads = model->tile_reads;
  const int n_slots = std::min(kHostSlots, n_chunks);
  for (int s = 0; s < n_slots; ++s)
    M6A_CUDA(slot_reserve(model->slots[s], max_sites, std::max<int64_t>(1, max_reads), h_bag ? n_iters : 0));

  int launches = 0;
  int rc = M6A_OK;
  for (int c = 0; c < n_chunks && rc == M6A_OK; ++c) {
    HostSlot& sl = model->slots[c % n_slots];
    const int64_t sa = cut[c], sb = cut[c + 1], ns = sb - sa;
    if (ns == 0) continue;
    const int64_t ra = read_off[sa], nr = read_off[sb] - ra;
    cudaError_t e = cudaStreamSynchronize(sl.stream);  // slot (and its pinned offset staging) free again
    if (e == cudaSuccess) {
      for (int64_t i = 0; i <= ns; ++i) sl.h_off[i] = read_off[sa + i] - ra;
      e = cudaMemcpyAsync(sl.d_off, sl.h_off, (ns + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, sl.stream);
    }
    if (e == cudaSuccess && nr > 0)
      e = cudaMemcpyAsync(sl.d_feats, feats + ra * kNSig, nr * kNSig * sizeof(float), cudaMemcpyHostToDevice, sl.stream);
    if (e == cudaSuccess && kmer_idx)
      e = cudaMemcpyAsync(sl.d_kmer, kmer_idx + sa * kKmerPos, ns * kKmerPos * sizeof(int32_t), cudaMemcpyHostToDevice,
                          sl.stream);
    if (e != cudaSuccess) {
      rc = static_cast<int>(e);
      break;
    }
    BagArgs dev_bags = {nullptr, 1, kPoolProd};
    if (host_bags) {
      dev_bags = *host_bags;
      dev_bags.bag_prob = h_bag ? static_cast<float*>(sl.d_bag) : nullptr;
    }
    rc = infer_device_impl(model, tile_reads, static_cast<const float*>(sl.d_feats), static_cast<const int64_t*>(sl.d_off),
                           kmer_idx ? static_cast<const int32_t*>(sl.d_kmer) : nullptr, ns, nr, site_id_base + sa,
                           n_samples, n_iters, seed, nullptr, read_threshold, static_cast<float*>(sl.d_rp),
                           static_cast<float*>(sl.d_sp), static_cast<int32_t*>(sl.d_mc), sl.d_ws,
                           static_cast<int64_t>(sl.cap_ws), sl.stream, host_bags ? &dev_bags : nullptr);
    if (rc != M6A_OK) break;
    ++launches;
    if (h_bag)
      e = cudaMemcpyAsync(h_bag + sa * n_iters, sl.d_bag, static_cast<size_t>(ns) * n_iters * sizeof(float),
                          cudaMemcpyDeviceToHost, sl.stream);
    if (e != cudaSuccess) {
      rc = static_cast<int>(e);
      break;
    }
    if (nr > 0) e = cudaMemcpyAsync(read_prob + ra, sl.d_rp, nr * sizeof(float), cudaMemcpyDeviceToHost, sl.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(site_prob + sa, sl.d_sp, ns * sizeof(float), cudaMemcpyDeviceToHost, sl.stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(mod_count + sa, sl.d_mc, ns * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream);
    if (e != cudaSuccess) rc = static_cast<int>(e);
  }
  for (int s = 0; s < n_slots; ++s) {
    cudaError_t e = cudaStreamSynchronize(model->slots[s].stream);
    if (e != cudaSuccess && rc == M6A_OK) rc = static_cast<int>(e);
  }
  g_last_launches = launches;
  return rc;
}

extern "C" int m6a_mil_infer_host_f32(const m6a_model_t* model, const float* feats, const int64_t* read_off,
                                      const int32_t* kmer_idx, int64_t n_sites, int64_t site_id_base,
                                      int32_t n_samples, int32_t n_iters, uint64_t seed, float read_threshold,
                                      float* read_prob, float* site_prob, int32_t* mod_count, int32_t n_chunks) {
  return infer_host_impl(model, feats, read_off, kmer_idx, n_sites, site_id_base, n_samples, n_iters, seed, read_threshold,
                         read_prob, site_prob, mod_count, n_chunks, nullptr);
}

extern "C" int m6a_mil_validate_host_f32(const m6a_model_t* model, const float* feats, const int64_t* read_off,
                                         const int32_t* kmer_idx, int64_t n_sites, int64_t site_id_base,
                                         int32_t n_samples, int32_t n_iters, uint64_t seed, int32_t pooling,
                                         int32_t replace, float read_threshold, float* read_prob, float* bag_prob,
                                         float* site_prob, int32_t* mod_count, int32_t n_chunks) {
  if (pooling < kPoolProd || pooling > kPoolMax || (replace != 0 && replace != 1)) return M6A_EINVAL;
  const BagArgs bags = {bag_prob, replace, pooling};
  return infer_host_impl(model, feats, read_off, kmer_idx, n_sites, site_id_base, n_samples, n_iters, seed, read_threshold,
                         read_prob, site_prob, mod_count, n_chunks, &bags);
}
