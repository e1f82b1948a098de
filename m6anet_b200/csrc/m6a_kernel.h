// Internal interface between the C-ABI layer (m6a_api.cu) and the kernels (m6a_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "m6a_layout.h"

namespace m6a {

// Tunables (overridable at build time for A/B experiments: make EXTRA="-DM6A_PAIR_UNROLL=3 ..." OUT=...; the measured
// alternatives are listed in profiles/r01_tile_size_ab.txt).  Rows per tile are chosen at run time (auto_tile_reads).
#ifndef M6A_THREADS
#define M6A_THREADS 256
#endif
#ifndef M6A_RPT
#define M6A_RPT 2
#endif
#ifndef M6A_CTAS
#define M6A_CTAS 2
#endif
#ifndef M6A_GMAX
#define M6A_GMAX 64
#endif
#ifndef M6A_QCAP
#define M6A_QCAP 4096
#endif
#ifndef M6A_WEIGHTS_CONST
#define M6A_WEIGHTS_CONST 1   // 1: weight image as a __grid_constant__ kernel parameter (uniform datapath) instead of shared memory
#endif
#ifndef M6A_DYNAMIC_TILES
#define M6A_DYNAMIC_TILES 1   // 1: CTAs take tiles from a global counter (workspace slot n_tiles + 1, zeroed by the prepass)
#endif                        //    instead of striding by gridDim.x -- removes the tail imbalance of small / ragged shards
                              //    (B200, round 2: 1 M x 50 17.53 -> 16.33 ms, 125 k x 50 2.267 -> 2.125 ms, ragged 19.09 -> 17.33 ms)
#ifndef M6A_PAIR_UNROLL
#define M6A_PAIR_UNROLL 5   // pair-loop unroll: deeper LDCU lookahead (1: 19.4 ms, 3: 18.36, 5: 18.35, 15: 17.8 but ragged 24.7)
#endif
constexpr int kPairUnroll = M6A_PAIR_UNROLL;
constexpr int kThreads = M6A_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kReadsPerThread = M6A_RPT;
constexpr int kCtasPerSm = M6A_CTAS;
constexpr int kChunkReads = kThreads * kReadsPerThread;  // feature rows staged per bulk copy
constexpr int kSitesPerTileMax = M6A_GMAX;
constexpr int kQCap = M6A_QCAP;         // q = 1-p entries kept in shared memory per tile
constexpr int kCStride = kH1Max;        // even (float2 loads); 152 mod 32 = 24 keeps neighbouring site rows on distinct banks
constexpr int kTcQCap = 6144;           // q entries of a slab slot of the tensor-core kernel (m6a_kernel_tc.cu)
constexpr int kMcMaxBlocks = 64;        // == kMaxBlocks in m6a_rng.cuh: Monte-Carlo partial sums per site
constexpr int kMcMinItersPerLane = 8;

// Decomposition of a site's iterations into blocks of 32*ipl (oracle/philox.py: block_layout).
inline void block_layout(int n_iters, int* ipl, int* n_blocks) {
  int v = (n_iters + 32 * kMcMaxBlocks - 1) / (32 * kMcMaxBlocks);
  if (v < kMcMinItersPerLane) v = kMcMinItersPerLane;
  *ipl = v;
  *n_blocks = (n_iters + 32 * v - 1) / (32 * v);
}

struct KernelArgs {
  DeviceModel model;
  const float* feats;
  const int64_t* read_off;
  const int32_t* kmer_idx;
  const uint16_t* sample_idx;
  float* read_prob;
  float* site_prob;
  int32_t* mod_count;
  const long long* tile_bounds;   // [n_tiles + 1] first site of every tile (prepass: lower_bound(read_off, t * tile_reads))
  long long n_sites;
  long long n_tiles;
  long long site_id_base;
  unsigned long long feats_bytes;
  unsigned long long seed;
  int tile_reads;      // target feature rows per tile
  int site_stride;     // 32-bit words between consecutive sites in site_prob / mod_count (1; 2 = interleaved in one buffer)
  int n_samples;
  int n_iters;
  int n_blocks;        // ceil(n_iters / (32 * iters_per_lane)) <= 64
  int iters_per_lane;  // max(8, ceil(n_iters / 2048))
  float read_threshold;
  bool feats_tma_ok;   // feats base 16-byte aligned -> cp.async.bulk staging
};

// validate()-style bags (m6a_mil_validate_f32): the literal MIL forward of the reference's evaluation loop
// (utils/training_utils.py:236-256) -- per pass one bag of n_samples reads per site, pooled by the model's pooling block.
// Passed as a third kernel parameter so that the inference instantiations keep their parameter layout.
enum PoolKind : int { kPoolProd = 0, kPoolMean = 1, kPoolMax = 2 };
struct BagArgs {
  float* bag_prob;   // [n_sites, n_iters] per-pass pooled probability (nullptr: only the mean over passes is kept)
  int replace;       // 0: bags without replacement (Floyd draws, m6a_rng.cuh); 1: the inference index stream
  int pool;          // PoolKind
};

struct LaunchInfo {
  int grid, block, smem_bytes, tile_reads;
};

size_t smem_bytes();
// bags == nullptr: inference (Monte-Carlo noisy-OR with replacement); otherwise the validate()-style bags instantiation
cudaError_t launch_mil_infer(const KernelArgs& a, const WeightImage* host_image, const BagArgs* bags, int n_sms,
                             cudaStream_t stream, LaunchInfo* info);
// the tensor-core kernel (m6a_kernel_tc.cu): inference stream, n_samples == 20; n_sms CTAs of 512 threads
cudaError_t launch_mil_infer_tc(const KernelArgs& a, const tcx::WeightImageTc* d_image, int n_sms, cudaStream_t stream,
                                LaunchInfo* info);
cudaError_t tc_set_trap_record(int* mapped_device_ptr);
cudaError_t tc_read_profile(unsigned long long* out32);   // zeros unless built with -DM6A_TC_PROFILE=1
cudaError_t launch_tile_bounds(const int64_t* read_off, long long n_sites, long long n_tiles, int tile_reads,
                               long long* tile_bounds, cudaStream_t stream);
// without_replacement: the Floyd bag stream instead of the inference stream
cudaError_t launch_sample_indices(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples,
                                  bool without_replacement, int32_t* out, cudaStream_t stream);

}  // namespace m6a
