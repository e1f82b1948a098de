/*
 * C restatement of the reference MIL-inference hot path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 * Same algorithm as oracle/mil_oracle.py + oracle/philox.py, fast enough to check every site of a full-size
 * job.  Never linked into or loaded by the product (m6anet_b200).  Paths below are relative to the reference's
 * m6anet/ package.
 *
 *   oracle_read_probs     model.get_read_representation + pooling_filter.probability_layer
 *                         (utils/inference_utils.py:35-37; model_blocks/blocks.py:126,204-205,65,249-255;
 *                          model_blocks/pooling_blocks.py:52) -- UNFOLDED BatchNorm, float32
 *   oracle_sample_indices the shared index stream (oracle/philox.py: Philox4x32-10-seeded MWC64X lane streams)
 *   oracle_site_probs     _calculate_site_proba on that stream: (1 - prod(1 - p[idx], axis=1)).mean()
 *                         (utils/inference_utils.py:85-86), float32 products, mod_count (:53)
 *   oracle_sample_bags    bags WITHOUT replacement (oracle/philox.py "Floyd bags"), the device twin of
 *                         np.random.choice(n, min_reads, replace=False) (utils/data_utils.py:213-214)
 *   oracle_bag_probs      the evaluation loop of validate() (utils/training_utils.py:236-256): per pass one bag per site,
 *                         pooled like the model's pooling block (model_blocks/pooling_blocks.py:96-98,127-129,158-160)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MWC_A 4294883355u

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

static void block_layout(int n_iters, int *ipl, int *n_blocks) {
  int v = (n_iters + 32 * 64 - 1) / (32 * 64);
  if (v < 8) v = 8;
  *ipl = v;
  *n_blocks = (n_iters + 32 * v - 1) / (32 * v);
}

/* out [n_iters, n_samples] int32 */
void oracle_sample_indices(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples, int32_t *out) {
  int ipl, n_blocks;
  block_layout(n_iters, &ipl, &n_blocks);
  for (int b = 0; b < n_blocks; ++b)
    for (int l = 0; l < 32; ++l) {
      uint32_t c[4] = {(uint32_t)l, (uint32_t)b, (uint32_t)site_id, (uint32_t)(site_id >> 32)};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      uint32_t x = c[0], cy = (uint32_t)(((uint64_t)c[1] * (MWC_A - 1u)) >> 32);
      if (x == 0 && cy == 0) x = 1;
      for (int k = 0; k < ipl; ++k) {
        const long long it = ((long long)b * ipl + k) * 32 + l;
        if (it >= n_iters) break;
        for (int s = 0; s < n_samples;) {
          const uint32_t word = x ^ cy;
          const uint64_t t = (uint64_t)MWC_A * x + cy;
          x = (uint32_t)t; cy = (uint32_t)(t >> 32);
          const uint64_t u = (uint64_t)word * n_reads;
          out[it * n_samples + s++] = (int32_t)(u >> 32);
          if (n_reads <= 256u && s < n_samples)        /* paired regime: the low half is a second uniform word */
            out[it * n_samples + s++] = (int32_t)(((u & 0xffffffffu) * n_reads) >> 32);
        }
      }
    }
}

/* Unfolded read encoder.  emb may be NULL (signal-only topology, emb_dim = 0).  kmer_idx is per SITE. */
void oracle_read_probs(const float *feats, const int64_t *read_off, const int32_t *kmer_idx, int64_t s_lo, int64_t s_hi, int h1, int h2,
                       int emb_dim, const float *emb, const float *w1, const float *b1, const float *bn_gamma,
                       const float *bn_beta, const float *bn_mean, const float *bn_var, float bn_eps, const float *w2,
                       const float *b2, const float *w3, const float *b3, float *read_prob) {
  const int in1 = 9 + 3 * emb_dim;
  for (int64_t s = s_lo; s < s_hi; ++s) {   /* callers thread over site ranges (oracle/c_oracle.py) */
    float x[9 + 3 * 16], h[1024], g[256];
    for (int t = 0; t < 3; ++t)
      for (int d = 0; d < emb_dim; ++d) x[9 + t * emb_dim + d] = emb[(int64_t)kmer_idx[3 * s + t] * emb_dim + d];
    for (int64_t r = read_off[s]; r < read_off[s + 1]; ++r) {
      memcpy(x, feats + 9 * r, 9 * sizeof(float));
      for (int j = 0; j < h1; ++j) {
        float a = 0.0f;
        for (int k = 0; k < in1; ++k) a += w1[(int64_t)j * in1 + k] * x[k];
        a += b1[j];
        a = (a - bn_mean[j]) * (1.0f / sqrtf(bn_var[j] + bn_eps)) * bn_gamma[j] + bn_beta[j];   /* BatchNorm1d, eval */
        h[j] = a > 0.0f ? a : 0.0f;
      }
      for (int k = 0; k < h2; ++k) {
        float a = 0.0f;
        for (int j = 0; j < h1; ++j) a += w2[(int64_t)k * h1 + j] * h[j];
        a += b2[k];
        g[k] = a > 0.0f ? a : 0.0f;
      }
      float z = 0.0f;
      for (int k = 0; k < h2; ++k) z += w3[k] * g[k];
      z += b3[0];
      read_prob[r] = 1.0f / (1.0f + expf(-z));
    }
  }
}

/* MC noisy-OR of every site on the shared index stream + mod_count.  site_prob NaN for empty sites. */
void oracle_site_probs(const float *read_prob, const int64_t *read_off, int64_t s_lo, int64_t s_hi, int64_t site_id_base,
                       int n_iters, int n_samples, uint64_t seed, float threshold, float *site_prob, int32_t *mod_count) {
  {
    int32_t *idx = (int32_t *)malloc((size_t)n_iters * n_samples * sizeof(int32_t));
    for (int64_t s = s_lo; s < s_hi; ++s) {
      const float *p = read_prob + read_off[s];
      const int64_t n = read_off[s + 1] - read_off[s];
      int cnt = 0;
      for (int64_t r = 0; r < n; ++r) cnt += p[r] >= threshold;
      mod_count[s] = cnt;
      if (n == 0) { site_prob[s] = NAN; continue; }
      oracle_sample_indices(seed, (uint64_t)(site_id_base + s), (uint32_t)n, n_iters, n_samples, idx);
      double acc = 0.0;   /* mean of float32 terms; accumulated wide so that the checker's own sum adds no error */
      for (int it = 0; it < n_iters; ++it) {
        float prod = 1.0f;
        for (int k = 0; k < n_samples; ++k) prod *= 1.0f - p[idx[it * n_samples + k]];
        acc += (double)(1.0f - prod);
      }
      site_prob[s] = (float)(acc / n_iters);
    }
    free(idx);
  }
}

/* Floyd bags of one site: out [n_iters, n_samples] int32, distinct inside a bag; requires n_reads >= n_samples. */
void oracle_sample_bags(uint64_t seed, uint64_t site_id, uint32_t n_reads, int n_iters, int n_samples, int32_t *out) {
  int ipl, n_blocks;
  block_layout(n_iters, &ipl, &n_blocks);
  for (int b = 0; b < n_blocks; ++b)
    for (int l = 0; l < 32; ++l) {
      uint32_t c[4] = {(uint32_t)l, (uint32_t)b, (uint32_t)site_id, (uint32_t)(site_id >> 32)};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      uint32_t x = c[0], cy = (uint32_t)(((uint64_t)c[1] * (MWC_A - 1u)) >> 32);
      if (x == 0 && cy == 0) x = 1;
      for (int k = 0; k < ipl; ++k) {
        const long long it = ((long long)b * ipl + k) * 32 + l;
        if (it >= n_iters) break;
        int32_t *pick = out + it * n_samples;
        for (int d = 0; d < n_samples; ++d) {
          const uint32_t word = x ^ cy;
          const uint64_t t = (uint64_t)MWC_A * x + cy;
          x = (uint32_t)t; cy = (uint32_t)(t >> 32);
          const uint32_t j = n_reads - (uint32_t)n_samples + (uint32_t)d;
          const uint32_t cand = (uint32_t)(((uint64_t)word * (j + 1u)) >> 32);     /* uniform on [0, j] */
          int dup = 0;
          for (int e = 0; e < d; ++e) dup |= (pick[e] == (int32_t)cand);
          pick[d] = (int32_t)(dup ? j : cand);
        }
      }
    }
}

/* bag_prob [n_sites, n_iters] float32 (NaN for a site with fewer reads than the bag); pool: 0 noisy-OR, 1 mean, 2 max. */
void oracle_bag_probs(const float *read_prob, const int64_t *read_off, int64_t s_lo, int64_t s_hi, int64_t site_id_base,
                      int n_iters, int n_samples, uint64_t seed, int pool, float *bag_prob) {
  int32_t *idx = (int32_t *)malloc((size_t)n_iters * n_samples * sizeof(int32_t));
  for (int64_t s = s_lo; s < s_hi; ++s) {
    const float *p = read_prob + read_off[s];
    const int64_t n = read_off[s + 1] - read_off[s];
    float *out = bag_prob + s * n_iters;
    if (n < n_samples) {
      for (int it = 0; it < n_iters; ++it) out[it] = NAN;
      continue;
    }
    oracle_sample_bags(seed, (uint64_t)(site_id_base + s), (uint32_t)n, n_iters, n_samples, idx);
    for (int it = 0; it < n_iters; ++it) {
      const int32_t *b = idx + it * n_samples;
      float y;
      if (pool == 0) {
        float prod = 1.0f;
        for (int k = 0; k < n_samples; ++k) prod *= 1.0f - p[b[k]];
        y = 1.0f - prod;
      } else if (pool == 1) {
        float sum = 0.0f;
        for (int k = 0; k < n_samples; ++k) sum += p[b[k]];
        y = sum / (float)n_samples;
      } else {
        y = p[b[0]];
        for (int k = 1; k < n_samples; ++k) y = p[b[k]] > y ? p[b[k]] : y;
      }
      out[it] = y;
    }
  }
  free(idx);
}
