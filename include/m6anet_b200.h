/*
 * m6anet_b200 -- C ABI of the B200-native m6anet MIL-inference hot path.
 *
 * One call scores a batch of DRACH sites: read encoder (k-mer embedding + 2-layer MLP + sigmoid
 * read classifier) over every read, mod_ratio, and the Monte-Carlo noisy-OR site probability
 * averaged over n_iters samplings of n_samples reads.
 *
 * The reference is pure Python and has no FFI; each entry point below names the reference
 * interface it replaces (paths relative to the reference's m6anet/ package):
 *
 *   m6a_model_create        MILModel(toml).load_state_dict(...)            scripts/inference.py:88-90,
 *                            (weights of the Sequential read encoder)        model/model.py:40-69
 *   m6a_mil_infer_f32       the body of run_inference per batch:            utils/inference_utils.py:35-54
 *                              get_read_representation + probability_layer  (:35-37)
 *                              group_results                                (:48-51,107-140)
 *                              mod_ratio                                    (:53)
 *                              calculate_site_proba/_calculate_site_proba   (:54,74-104)
 *   m6a_mil_infer_host_f32  the same, including features.to(device) / probs.cpu()  (:35-36,41)
 *   m6a_sample_indices      np.random.choice index draw                     (:85)
 *   m6a_mil_validate_f32    the evaluation loop of validate(): per pass one bag of min_reads reads per site drawn
 *   m6a_mil_validate_host_f32   WITHOUT replacement (utils/data_utils.py:213-214), MILModel.forward on it
 *                            (model/model.py:155-164: read encoder + pooling block, model_blocks/pooling_blocks.py:96-98,
 *                            127-129,158-160), passes averaged             utils/training_utils.py:236-256
 *   m6a_sample_bags         np.random.choice(n, min_reads, replace=False)   utils/data_utils.py:214
 *   m6a_ingest_parts        NanopolishDS._load_data/__getitem__/get_norm_factor + inference_collate,
 *                            NanopolishReplicateDS.load_data                 utils/data_utils.py:152-248,395-427,498-506
 *   m6a_info_count/_read    pd.read_csv(data.info)                          utils/data_utils.py:118-129
 *   m6a_write_site_csv      the site rows   '%s,%d,%s,%.16f,%s,%.16f'       utils/inference_utils.py:59-60
 *   m6a_write_indiv_csv     the read rows   '%s,%d,%s,%.16f'                utils/inference_utils.py:63-64
 *
 * Conventions: plain C types only; no exceptions cross the boundary; every function returns an
 * int status (0 = ok, <0 = M6A_E*, >0 = a cudaError_t value) readable with m6a_strerror().
 * Device entry points enqueue on `stream` of the CURRENT device and return without
 * synchronising; the caller owns every buffer.
 */
#ifndef M6ANET_B200_H_
#define M6ANET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M6A_VERSION 100 /* 0.1.0 */

#define M6A_OK 0
#define M6A_EINVAL (-1)       /* NULL pointer / negative size / inconsistent argument        */
#define M6A_EUNSUPPORTED (-2) /* model dimensions outside the compiled kernel limits         */
#define M6A_EALIGN (-3)       /* a device buffer is not aligned as documented                */
#define M6A_ERANGE (-4)       /* a site has more reads than the mode allows                  */
#define M6A_ENOMEM (-5)       /* host allocation failed                                      */
#define M6A_EPARSE (-6)       /* a data.json site line does not match data.info / the format  */
#define M6A_EIO (-7)          /* open/read/write failed                                      */

/* Compiled limits of the kernel (see m6anet_b200/csrc/m6a_layout.h). */
#define M6A_N_SIG 9          /* signal features per read: 3 positions x (dwell, sd, mean)   */
#define M6A_H2 32            /* width of the second Linear block                            */
#define M6A_H1_MAX 152       /* max width of the first Linear block (shipped models: 150)   */
#define M6A_MAX_READS_EXPLICIT 65535 /* uint16 explicit indices                             */
#define M6A_MAX_SAMPLES 64   /* reads per bag limit (n_samples)                             */

/* Pooling block of the model, used by the validate()-style entry points (reference model_blocks/pooling_blocks.py). */
#define M6A_POOL_PROD 0      /* SigmoidProdPooling  1 - prod(1 - p)   :127-129              */
#define M6A_POOL_MEAN 1      /* SigmoidMeanPooling  mean(p)           :96-98                */
#define M6A_POOL_MAX 2       /* SigmoidMaxPooling   max(p)            :158-160              */

/*
 * Read-encoder parameters, HOST pointers, row-major float32, eval-mode BatchNorm already folded
 * into w1/b1 by the caller (m6anet_b200/model.py does this in float64).
 * Layout mirrors the reference state_dict:
 *   emb [n_kmer, emb_dim]          read_level_encoder.1.embedding_layer.weight   (NULL when emb_dim == 0)
 *   w1  [h1, n_sig + 3*emb_dim]    read_level_encoder.3.layers.0.weight (x BN scale); signal columns first
 *   b1  [h1]
 *   w2  [h2, h1], b2 [h2]          read_level_encoder.4.layers.0.{weight,bias}
 *   w3  [h2], b3 [1]               pooling_filter.probability_layer.0.{weight,bias}
 */
typedef struct {
  const float *emb;
  const float *w1, *b1;
  const float *w2, *b2;
  const float *w3, *b3;
  int32_t n_kmer, emb_dim, n_sig, h1, h2;
} m6a_weights_t;

typedef struct m6a_model m6a_model_t; /* opaque: packed weight image resident on one device */

/* Where the two Linear blocks of the read encoder run (results agree to float32 round-off, same index streams):
 *   M6A_ENCODER_FFMA  CUDA cores, packed FFMA2 (mil_infer_kernel, m6anet_b200/csrc/m6a_kernel.cu)
 *   M6A_ENCODER_TC    5th-generation tensor cores, tcgen05.mma kind::tf32 with error-compensated 3xTF32 operands and the
 *                     accumulators in TMEM (mil_infer_tc_kernel, m6anet_b200/csrc/m6a_kernel_tc.cu); used for the
 *                     inference entry points with n_samples == 20 and no explicit indices, the other calls (explicit
 *                     indices, validate()-style bags, other bag sizes) run the FFMA kernel whatever the setting. */
#define M6A_ENCODER_FFMA 0
#define M6A_ENCODER_TC 1
#ifndef M6A_ENCODER_DEFAULT
#define M6A_ENCODER_DEFAULT M6A_ENCODER_TC
#endif

int m6a_version(void);
const char *m6a_strerror(int status);
/* Thin wrappers of cudaGetDeviceCount / cudaSetDevice so that a host program needs no other CUDA binding. */
int m6a_device_count(int32_t *count);
int m6a_set_device(int32_t device);

/* Packs the weights into the two images the kernels consume -- the pair-interleaved image of the CUDA-core kernel (passed by
 * value as a __grid_constant__ kernel parameter) and the RN_tf32 hi/lo split UMMA operands of the tensor-core kernel (copied
 * to shared memory once per CTA) -- and uploads them to the current CUDA device (synchronous).  The model may be used from any
 * stream of that device. */
int m6a_model_create(const m6a_weights_t *w, m6a_model_t **out);
int m6a_model_destroy(m6a_model_t *model);
/* Selects the read-encoder implementation of this model (M6A_ENCODER_*); the environment variable M6A_ENCODER=ffma|tc
 * overrides the default at m6a_model_create.  M6A_EUNSUPPORTED when the model does not fit the tensor-core image
 * (emb_dim > 2 or h1 > 160).  m6a_model_get_encoder returns the implementation in effect. */
int m6a_model_set_encoder(m6a_model_t *model, int32_t encoder);
int m6a_model_get_encoder(const m6a_model_t *model);
/* Debug aid: arms (first call) and reads the trap record of the tensor-core kernel -- {wait site, block, thread, parity}
 * of a bounded mbarrier wait that gave up, kept in mapped host memory so that it survives the trap. */
int m6a_debug_trap_record(m6a_model_t *model, int32_t *out64);   /* 16 wait sites x {site, block, thread, parity} */
/* Page-locked host memory (cudaHostAlloc, portable) for the host-buffer entry points: the H2D / D2H copies of
 * m6a_mil_infer_host_f32 only overlap with the kernel when its buffers are pinned. */
int m6a_pinned_alloc(void **out, int64_t bytes);
int m6a_pinned_free(void *p);
/* One-line description of the kernels this build ships. */
const char *m6a_build_info(void);
/* Feature rows per tile of the device call, 64..4096; 0 (default) = automatic: a multiple of the site depth close to
 * 1000 rows (500 for small jobs).  Tiles are read-balanced: tile t holds the sites whose first row is in [t*T, (t+1)*T). */
int m6a_model_set_tile_reads(m6a_model_t *model, int32_t tile_reads);
/* The automatic choice for a job of n_sites sites / total_reads rows on a device with n_sms SMs (introspection). */
int32_t m6a_auto_tile_reads(int64_t n_sites, int64_t total_reads, int32_t n_sms);
/* The same for the tensor-core encoder: 64 sites' worth of rows (<= 6144), shrunk to total / (rounds x SMs) rounded up to whole
 * 128-row MMA tiles when the job is fewer than 64 rounds of tiles long, so that the last round is as full as the others. */
int32_t m6a_auto_tile_reads_tc(int64_t n_sites, int64_t total_reads, int32_t n_sms);

/*
 * Score n_sites sites.  All data pointers are DEVICE pointers.
 *
 *   feats       [total_reads, 9] float32, normalised, site-contiguous rows (4-byte aligned; a
 *               16-byte aligned base enables the TMA bulk-copy path, otherwise plain loads are used)
 *   read_off    [n_sites + 1] int64 CSR offsets into feats rows, non-decreasing, read_off[0] == 0 and
 *               read_off[n_sites] == total_reads (a shard passes its own rows and re-based offsets)
 *   kmer_idx    [n_sites, 3] int32 five-mer ids of the site's 7-mer (ignored when emb_dim == 0; may be NULL then)
 *   site_id_base  global id of site 0: the RNG counter is (site_id_base + s), so results do not
 *               depend on how sites are sharded over GPUs
 *   n_samples   reads per bag (the reference hard-codes 20), 1..M6A_MAX_SAMPLES
 *   n_iters     Monte-Carlo iterations (>= 1)
 *   seed        Philox key of the index streams
 *   sample_idx  optional [n_sites, n_iters, n_samples] uint16 explicit indices (parity / replay
 *               mode, requires every site to have <= 65535 reads); NULL => on-device stream:
 *               Philox4x32-10-seeded MWC64X lane streams, index = (word * n_reads) >> 32
 *               (specification: oracle/philox.py, m6anet_b200/csrc/m6a_rng.cuh)
 *   read_threshold  float32 threshold of mod_ratio (p >= threshold)
 * outputs
 *   read_prob   [total_reads] float32, indexed like feats rows (absolute row index)
 *   site_prob   [n_sites] float32; NaN for a site without reads
 *   mod_count   [n_sites] int32 number of reads with p >= read_threshold
 *               (mod_ratio = mod_count / n_reads, formed by the host in float64 like np.mean)
 * scratch
 *   workspace   DEVICE buffer of at least m6a_mil_workspace_bytes(total_reads) bytes, 8-byte aligned, owned by the
 *               caller and private to this call until it has completed on `stream` (tile boundaries of the prepass and
 *               one tile-counter word).
 *               The library allocates nothing on this path.
 */
int64_t m6a_mil_workspace_bytes(int64_t total_reads);
int m6a_mil_infer_f32(const m6a_model_t *model, const float *feats, const int64_t *read_off,
                      const int32_t *kmer_idx, int64_t n_sites, int64_t total_reads,
                      int64_t site_id_base, int32_t n_samples, int32_t n_iters, uint64_t seed,
                      const uint16_t *sample_idx, float read_threshold, float *read_prob,
                      float *site_prob, int32_t *mod_count, void *workspace, int64_t workspace_bytes,
                      void *stream);

/* m6a_mil_infer_f32 without explicit indices, writing the two per-site outputs INTERLEAVED into one buffer
 * site_out [n_sites][2] 32-bit words (word 0 = site probability float32, word 1 = mod_count int32; 8-byte aligned):
 * the send buffer of the one all-gather of a multi-GPU run, so no pack kernels run. */
int m6a_mil_infer_packed_f32(const m6a_model_t *model, const float *feats, const int64_t *read_off,
                             const int32_t *kmer_idx, int64_t n_sites, int64_t total_reads,
                             int64_t site_id_base, int32_t n_samples, int32_t n_iters, uint64_t seed,
                             float read_threshold, float *read_prob, void *site_out, void *workspace,
                             int64_t workspace_bytes, void *stream);

/*
 * Same computation with HOST buffers (read_off[0] must be 0 here).  Sites are cut into
 * `n_chunks` read-balanced chunks (0 => automatic) that are copied, scored and copied back on
 * rotating CUDA streams so that H2D, kernel and D2H overlap.  Synchronous.  Pinned host buffers
 * overlap best; pageable ones still work.  sample_idx is not supported on this path.
 * The model handle keeps a grow-only device workspace (3 pipeline slots) for this call, so steady-state
 * calls allocate nothing; concurrent host calls on one model are serialised.
 */
int m6a_mil_infer_host_f32(const m6a_model_t *model, const float *feats, const int64_t *read_off,
                           const int32_t *kmer_idx, int64_t n_sites, int64_t site_id_base,
                           int32_t n_samples, int32_t n_iters, uint64_t seed, float read_threshold,
                           float *read_prob, float *site_prob, int32_t *mod_count, int32_t n_chunks);

/*
 * validate()-style literal MIL forward: the evaluation loop of the reference (utils/training_utils.py:236-256) for a
 * batch of sites.  The read encoder runs once per read (eval mode: a read's probability does not depend on its bag);
 * then every pass `it` < n_iters pools ONE bag of n_samples reads per site with the model's pooling block.
 * Arguments as m6a_mil_infer_f32, plus
 *   pooling     M6A_POOL_PROD / _MEAN / _MAX
 *   replace     0: bags WITHOUT replacement, like np.random.choice(n, min_reads, replace=False) of the reference's
 *               evaluation datasets (utils/data_utils.py:214) -- Floyd's algorithm on the (site, block, lane) MWC64X
 *               streams, one word per pick (specification: oracle/philox.py "Floyd bags"); a site with fewer than
 *               n_samples reads has no bag: its outputs are NaN.  1: the inference index stream (with replacement).
 *   sample_idx  optional explicit bags [n_sites, n_iters, n_samples] uint16 (replays the reference's MT19937 draw);
 *               overrides `replace`
 * outputs
 *   bag_prob    [n_sites, n_iters] float32 pooled probability of every pass (validate()'s y_pred, transposed); may be NULL
 *   site_prob   [n_sites] float32 mean over the passes (device summation order; the host mirror re-averages bag_prob in
 *               pass order like np.mean(all_y_pred, axis=0))
 *   read_prob, mod_count  as m6a_mil_infer_f32
 */
int m6a_mil_validate_f32(const m6a_model_t *model, const float *feats, const int64_t *read_off,
                         const int32_t *kmer_idx, int64_t n_sites, int64_t total_reads, int64_t site_id_base,
                         int32_t n_samples, int32_t n_iters, uint64_t seed, const uint16_t *sample_idx,
                         int32_t pooling, int32_t replace, float read_threshold, float *read_prob, float *bag_prob,
                         float *site_prob, int32_t *mod_count, void *workspace, int64_t workspace_bytes, void *stream);
/* The same with HOST buffers, through the chunked H2D / kernel / D2H pipeline of m6a_mil_infer_host_f32. */
int m6a_mil_validate_host_f32(const m6a_model_t *model, const float *feats, const int64_t *read_off,
                              const int32_t *kmer_idx, int64_t n_sites, int64_t site_id_base, int32_t n_samples,
                              int32_t n_iters, uint64_t seed, int32_t pooling, int32_t replace, float read_threshold,
                              float *read_prob, float *bag_prob, float *site_prob, int32_t *mod_count, int32_t n_chunks);
/* Writes the device without-replacement bags of one site: out [n_iters, n_samples] int32 (DEVICE pointer), distinct
 * inside a bag.  Test hook like m6a_sample_indices; M6A_ERANGE when n_reads < n_samples. */
int m6a_sample_bags(uint64_t seed, int64_t site_id, int32_t n_reads, int32_t n_iters, int32_t n_samples,
                    int32_t *out, void *stream);

/* Writes the device index stream of one site: out [n_iters, n_samples] int32 (DEVICE pointer).
 * Test hook proving the device generator equals oracle/philox.py bit for bit. */
int m6a_sample_indices(uint64_t seed, int64_t site_id, int32_t n_reads, int32_t n_iters,
                       int32_t n_samples, int32_t *out, void *stream);

/*
 * ---- host I/O of the path (multi-threaded, no CUDA) ------------------------------------------------
 * One part = one site line of one data.json file: `{"<tx>":{"<pos>":{"<7-mer>":[[f0..f8, read_id], ...]}}}`
 * in the byte range [start, end) (data.info columns start/end; reference utils/dataprep_utils.py:473-485).
 * A site has one part per input directory that contains it (replicates are concatenated in part order).
 */
typedef struct {
  int32_t file;          /* index into paths[]                                              */
  int32_t rep;           /* replicate number (informational)                                */
  int64_t start, end;    /* byte range of the line                                          */
  int64_t row_off;       /* first output row of this part                                   */
  int64_t n_rows;        /* rows expected (data.info n_reads); a mismatch is M6A_EPARSE      */
  int64_t site;          /* output site index (row of kmer_idx)                             */
  int32_t first_of_site; /* 1: this part writes kmer_idx[site]                              */
  int32_t reserved;
} m6a_part_t;

/*
 * Parses the parts with n_threads workers (0 = all cores) into
 *   feats    [rows, 3*(2*n_flank+1)] float32 = (raw - mean) / std evaluated in float64, rounded once
 *   read_ids [rows] int64, kmer_idx [n_sites, 2*n_flank+1] int32
 * norm_mean / norm_std: [1024, 3] float64 indexed by the base-4 code of a five-mer (A,C,G,T = 0..3,
 * first letter most significant), NaN = five-mer absent from the norm factors; kmer_id: [1024] five-mer
 * code -> id of the 66-entry table (-1 = none).  *bad_part receives the index of the first failing part.
 */
int m6a_ingest_parts(const char *const *paths, int32_t n_files, const m6a_part_t *parts, int64_t n_parts,
                     int32_t n_flank, const double *norm_mean, const double *norm_std,
                     const int32_t *kmer_id, float *feats, int64_t *read_ids, int32_t *kmer_idx,
                     int32_t n_threads, int64_t *bad_part);

/* The same, and every part's line must carry the keys of its site: tx_buf/tx_off = concatenated transcript ids + CSR offsets
 * [n_sites + 1], tx_pos [n_sites] (the reference looks the line up by json.loads(line)[tx_id][str(tx_pos)] and raises KeyError
 * on a stale data.info, utils/data_utils.py:185); a mismatch is M6A_EPARSE.  Replicates of a site must also agree on the 7-mer. */
int m6a_ingest_parts_keyed(const char *const *paths, int32_t n_files, const m6a_part_t *parts, int64_t n_parts,
                           int32_t n_flank, const double *norm_mean, const double *norm_std,
                           const int32_t *kmer_id, const char *tx_buf, const int64_t *tx_off, const int64_t *tx_pos,
                           float *feats, int64_t *read_ids, int32_t *kmer_idx, int32_t n_threads, int64_t *bad_part);

/* data.info reader (csv with a header naming transcript_id, transcript_position, start, end, n_reads; reference
 * utils/data_utils.py:118-129 reads it with pandas).  m6a_info_count sizes the buffers, m6a_info_read fills them:
 * tx_buf/tx_off = concatenated transcript ids + CSR offsets [n_rows+1]; the other columns as int64 [n_rows]. */
int m6a_info_count(const char *path, int64_t *n_rows, int64_t *tx_bytes);
int m6a_info_read(const char *path, int64_t n_rows, int64_t tx_bytes, char *tx_buf, int64_t *tx_off,
                  int64_t *tx_pos, int64_t *start, int64_t *end, int64_t *n_reads);

/* Appends the reference-format rows to the open file descriptor fd (rows are formatted by n_threads
 * workers and written in site order).  tx_buf/tx_off: concatenated transcript ids, CSR offsets [n_sites+1];
 * kmer5: [n_sites, 5] chars (centre five-mer); read_rep NULL => integer read_index, else "{id}_{rep}". */
int m6a_write_site_csv(int32_t fd, int64_t n_sites, const char *tx_buf, const int64_t *tx_off,
                       const int64_t *tx_pos, const int64_t *read_off, const float *site_prob,
                       const int32_t *mod_count, const char *kmer5, int32_t n_threads);
int m6a_write_indiv_csv(int32_t fd, int64_t n_sites, const char *tx_buf, const int64_t *tx_off,
                        const int64_t *tx_pos, const int64_t *read_off, const int64_t *read_ids,
                        const int32_t *read_rep, const float *read_prob, int32_t n_threads);

/* Launch geometry of the last m6a_mil_infer_f32 call on this thread (for bench/roofline
 * reporting): grid, block, dynamic smem bytes, feature rows per tile, number of kernel launches. */
int m6a_last_launch(int32_t *grid, int32_t *block, int32_t *smem_bytes, int32_t *tile_reads,
                    int32_t *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* M6ANET_B200_H_ */
