// C ABI of libpxr.so (include/pxr.h): handle management, weight folding and
// dispatch to the SIMT / tcgen05 kernels.
#include "pxr_common.cuh"

static char g_create_err[512] = "";

extern "C" int pxr_version(void) { return PXR_VERSION; }

extern "C" const char* pxr_last_error(const pxr_handle* h) { return h ? h->err : g_create_err; }

extern "C" int pxr_create(const pxr_config* cfg, pxr_handle** out) {
#define CFAIL(...) do { snprintf(g_create_err, sizeof(g_create_err), __VA_ARGS__); return PXR_ERR_INVALID; } while (0)
  if (!cfg || !out) CFAIL("pxr_create: NULL argument");
  if (cfg->struct_size != (int32_t)sizeof(pxr_config)) CFAIL("pxr_config size mismatch: got %d, library expects %zu", cfg->struct_size, sizeof(pxr_config));
  if (cfg->fusion < 0 || cfg->fusion > 2) CFAIL("Unknown fusion type %d", cfg->fusion);
  if (cfg->embedding_dim <= 0 || cfg->embedding_dim % 4 || cfg->embedding_dim > 512) CFAIL("embedding_dim must be a multiple of 4 in (0, 512], got %d", cfg->embedding_dim);
  if (cfg->n_hidden < 1 || cfg->n_hidden > PXR_MAX_HIDDEN) CFAIL("fusion_hidden_dims must have 1..%d entries", PXR_MAX_HIDDEN);
  for (int l = 0; l < cfg->n_hidden; ++l) if (cfg->hidden[l] <= 0 || cfg->hidden[l] % 4) CFAIL("fusion_hidden_dims[%d]=%d must be a positive multiple of 4", l, cfg->hidden[l]);
  if (cfg->projection_hidden < 0 || cfg->projection_hidden % 4) CFAIL("projection_hidden_dim must be a multiple of 4");
  if (cfg->fusion == PXR_FUSION_ATTENTION && (cfg->num_heads <= 0 || cfg->embedding_dim % cfg->num_heads)) CFAIL("embed_dim must be divisible by num_heads");
  if (cfg->n_tags <= 0) CFAIL("n_tags must be positive");
  if (cfg->precision != PXR_PRECISION_BF16 && cfg->precision != PXR_PRECISION_FP16) CFAIL("unknown precision %d", cfg->precision);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { snprintf(g_create_err, sizeof(g_create_err), "no CUDA device: libpxr has no CPU fallback"); return PXR_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { snprintf(g_create_err, sizeof(g_create_err), "cudaGetDeviceProperties failed"); return PXR_ERR_CUDA; }
  // libpxr.so carries sm_100a code only (no PTX for other architectures, no second backend)
  if (prop.major != 10) { snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d: libpxr.so is built for sm_100a (B200) only", dev, prop.major, prop.minor); return PXR_ERR_INVALID; }
  pxr_handle* h = new pxr_handle();
  h->cfg = *cfg; h->device = dev; h->n_sm = prop.multiProcessorCount; h->err[0] = 0;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  h->M = 3 + (cfg->vision_dim > 0) + (cfg->language_dim > 0) + (cfg->num_numerical > 0);
  h->has_mod[0] = cfg->vision_dim > 0; h->has_mod[1] = cfg->language_dim > 0; h->has_mod[2] = cfg->num_numerical > 0;
  h->fast_ok = (prop.major == 10) && pxr_tc_supported(h);
  if (cfg->path == PXR_PATH_TCGEN05 && !h->fast_ok) { delete h; CFAIL("tcgen05 path requested but this configuration / device is not supported by it"); }
  h->path = (cfg->path == PXR_PATH_SIMT || !h->fast_ok) ? PXR_PATH_SIMT : PXR_PATH_TCGEN05;
  if (pxr_simt_smem_rows(h, false) == 0 || pxr_simt_smem_rows(h, true) == 0) { delete h; CFAIL("layer dims too large for shared memory"); }
  *out = h;
  return PXR_OK;
#undef CFAIL
}

extern "C" void pxr_destroy(pxr_handle* h) {
  if (!h) return;
  if (h->arena) cudaFree(h->arena);
  if (h->fast_w) cudaFree(h->fast_w);
  if (h->tc_items_w) cudaFree(h->tc_items_w);
  if (h->prof_ev) { for (int i = 0; i < 2 * PXR_PROFILE_SLOTS; ++i) cudaEventDestroy(h->prof_ev[i]); delete[] h->prof_ev; }
  delete h;
}

extern "C" int pxr_profile_enable(pxr_handle* h, int on) {
  if (!h) return PXR_ERR_INVALID;
  if (on && !h->prof_ev) {
    h->prof_ev = new cudaEvent_t[2 * PXR_PROFILE_SLOTS];
    for (int i = 0; i < 2 * PXR_PROFILE_SLOTS; ++i) PXR_CUDA(h, cudaEventCreate(&h->prof_ev[i]));
  }
  h->profile = on != 0; h->prof_n = 0;
  return PXR_OK;
}

extern "C" int pxr_profile_read(pxr_handle* h, double* total_ms, int64_t* n_launches) {
  if (!h || !total_ms || !n_launches) return PXR_ERR_INVALID;
  double tot = 0.0;
  for (int i = 0; i < h->prof_n; ++i) {
    PXR_CUDA(h, cudaEventSynchronize(h->prof_ev[2 * i + 1]));
    float ms = 0.f;
    PXR_CUDA(h, cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
    tot += ms;
  }
  *total_ms = tot; *n_launches = h->prof_n; h->prof_n = 0;
  return PXR_OK;
}

extern "C" int64_t pxr_launch_count(const pxr_handle* h) { return h ? h->launches : 0; }
extern "C" int pxr_active_path(const pxr_handle* h) { return h ? h->path : PXR_ERR_INVALID; }
extern "C" const char* pxr_path_reason(const pxr_handle* h) {
  if (!h) return "";
  if (h->path == PXR_PATH_TCGEN05) return "";
  if (h->cfg.path == PXR_PATH_SIMT) return "the SIMT path was requested";
  const char* r = pxr_tc_unsupported_reason(h);
  return r ? r : "";
}
extern "C" int pxr_set_rescore(pxr_handle* h, int on) {
  if (!h) return PXR_ERR_INVALID;
  h->rescore = on != 0;
  return PXR_OK;
}
extern "C" int pxr_get_rescore(const pxr_handle* h) { return h ? (h->rescore ? 1 : 0) : PXR_ERR_INVALID; }
extern "C" int pxr_set_records_only(pxr_handle* h, int on) {
  if (!h) return PXR_ERR_INVALID;
  h->records_only = on != 0;
  h->items_ready = false;                 // the item workspace layout changes: pxr_precompute_items must run again
  return PXR_OK;
}
extern "C" int pxr_set_path(pxr_handle* h, int path) {
  if (!h) return PXR_ERR_INVALID;
  if (path == PXR_PATH_TCGEN05 && !h->fast_ok) PXR_FAIL(h, PXR_ERR_INVALID, "tcgen05 path not supported for this configuration");
  h->path = (path == PXR_PATH_SIMT) ? PXR_PATH_SIMT : (h->fast_ok ? PXR_PATH_TCGEN05 : PXR_PATH_SIMT);
  return PXR_OK;
}

// ---------------------------------------------------------------------------
// weight preparation
// ---------------------------------------------------------------------------
// One block per output row n.  Optionally folds the PREVIOUS layer's eval-mode
// BatchNorm (which sits after the activation, multimodal.py:371-379) into this
// Linear: s = g/sqrt(var+eps), t = beta - mean*s, W' = W diag(s), b' = b + W t.
__global__ void fold_linear_kernel(const float* __restrict__ W, const float* __restrict__ b, int N, int K,
                                   const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                   const float* __restrict__ bn_mean, const float* __restrict__ bn_var, float eps,
                                   float* __restrict__ w_out, float* __restrict__ wt_out, float* __restrict__ b_out) {
  __shared__ double red[128];
  const int n = blockIdx.x;
  double part = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double w = W[(size_t)n * K + k];
    if (bn_w) {
      const double s = (double)bn_w[k] / sqrt((double)bn_var[k] + (double)eps);
      const double t = (double)bn_b[k] - (double)bn_mean[k] * s;
      part += w * t;
      w *= s;
    }
    w_out[(size_t)n * K + k] = (float)w;
    wt_out[(size_t)k * N + n] = (float)w;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) b_out[n] = (float)((b ? (double)b[n] : 0.0) + red[0]);
}

static float* arena_take(pxr_handle* h, size_t n_floats) {
  const size_t bytes = pxr_align_up(n_floats * sizeof(float), 256);
  float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(h->arena) + h->arena_used);
  h->arena_used += bytes;
  return p;
}

static int make_linear(pxr_handle* h, PxrLinear* L, const float* W, const float* b, int N, int K, const float* bn_w,
                       const float* bn_b, const float* bn_mean, const float* bn_var, float eps, cudaStream_t st) {
  L->n = N; L->k = K;
  L->w = arena_take(h, (size_t)N * K); L->wt = arena_take(h, (size_t)N * K); L->b = arena_take(h, N);
  fold_linear_kernel<<<N, 128, 0, st>>>(W, b, N, K, bn_w, bn_b, bn_mean, bn_var, eps, L->w, L->wt, L->b);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

extern "C" int pxr_load_weights(pxr_handle* h, const pxr_weights* w, pxr_stream stream) {
  if (!h || !w) return PXR_ERR_INVALID;
  if (w->struct_size != (int32_t)sizeof(pxr_weights)) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_weights size mismatch: got %d, expected %zu", w->struct_size, sizeof(pxr_weights));
  cudaStream_t st = (cudaStream_t)stream;
  const pxr_config& c = h->cfg;
  const int D = c.embedding_dim, M = h->M;
  if (!w->tag_embedding) PXR_FAIL(h, PXR_ERR_INVALID, "tag_embedding.weight missing");
  const float* pw0[3] = {w->vision_w0, w->language_w0, w->numerical_w0};
  const float* pb0[3] = {w->vision_b0, w->language_b0, w->numerical_b0};
  const float* pw1[3] = {w->vision_w1, w->language_w1, w->numerical_w1};
  const float* pb1[3] = {w->vision_b1, w->language_b1, w->numerical_b1};
  const int in_dim[3] = {c.vision_dim, c.language_dim, c.num_numerical};
  const char* names[3] = {"vision_projection", "language_projection", "numerical_projection"};
  // size the arena
  size_t floats = (size_t)c.n_tags * D + 64;
  auto lin = [&](int n, int k) { floats += 2 * (size_t)n * k + n + 3 * 64; };
  for (int m = 0; m < 3; ++m) if (h->has_mod[m]) {
    if (!pw0[m] || !pb0[m]) PXR_FAIL(h, PXR_ERR_INVALID, "%s.0 weights missing", names[m]);
    if (c.projection_hidden) { if (!pw1[m] || !pb1[m]) PXR_FAIL(h, PXR_ERR_INVALID, "%s.3 weights missing", names[m]); lin(c.projection_hidden, in_dim[m]); lin(D, c.projection_hidden); }
    else lin(D, in_dim[m]);
  }
  if (c.fusion == PXR_FUSION_GATED) { if (!w->gate_w || !w->gate_b) PXR_FAIL(h, PXR_ERR_INVALID, "fusion_layer.gating_network.0 missing"); lin(M, M * D); }
  if (c.fusion == PXR_FUSION_ATTENTION) {
    if (!w->attn_in_w || !w->attn_in_b || !w->attn_out_w || !w->attn_out_b || !w->attn_ln_w || !w->attn_ln_b) PXR_FAIL(h, PXR_ERR_INVALID, "fusion_layer.attention / norm weights missing");
    lin(3 * D, D); lin(D, D); floats += 2 * D + 128;
  }
  int in = (c.fusion == PXR_FUSION_CONCAT) ? M * D : D;
  for (int l = 0; l < c.n_hidden; ++l) {
    if (!w->mlp_w[l] || !w->mlp_b[l]) PXR_FAIL(h, PXR_ERR_INVALID, "prediction_network hidden layer %d missing", l);
    if (c.use_batch_norm && (!w->bn_w[l] || !w->bn_b[l] || !w->bn_mean[l] || !w->bn_var[l])) PXR_FAIL(h, PXR_ERR_INVALID, "BatchNorm %d tensors missing", l);
    lin(c.hidden[l], in); in = c.hidden[l];
  }
  if (!w->out_w || !w->out_b) PXR_FAIL(h, PXR_ERR_INVALID, "output Linear missing");
  lin(1, in);
  if (h->arena) { cudaFree(h->arena); h->arena = nullptr; }
  h->arena_bytes = floats * sizeof(float) + 4096; h->arena_used = 0;
  PXR_CUDA(h, cudaMalloc(&h->arena, h->arena_bytes));

  h->tag_emb = arena_take(h, (size_t)c.n_tags * D);
  PXR_CUDA(h, cudaMemcpyAsync(h->tag_emb, w->tag_embedding, sizeof(float) * c.n_tags * D, cudaMemcpyDeviceToDevice, st));
  int rc;
  for (int m = 0; m < 3; ++m) {
    h->proj[m][0] = PxrLinear(); h->proj[m][1] = PxrLinear();
    if (!h->has_mod[m]) continue;
    if (c.projection_hidden) {
      if ((rc = make_linear(h, &h->proj[m][0], pw0[m], pb0[m], c.projection_hidden, in_dim[m], 0, 0, 0, 0, 0.f, st))) return rc;
      if ((rc = make_linear(h, &h->proj[m][1], pw1[m], pb1[m], D, c.projection_hidden, 0, 0, 0, 0, 0.f, st))) return rc;
    } else if ((rc = make_linear(h, &h->proj[m][0], pw0[m], pb0[m], D, in_dim[m], 0, 0, 0, 0, 0.f, st))) return rc;
  }
  if (c.fusion == PXR_FUSION_GATED && (rc = make_linear(h, &h->gate, w->gate_w, w->gate_b, M, M * D, 0, 0, 0, 0, 0.f, st))) return rc;
  if (c.fusion == PXR_FUSION_ATTENTION) {
    if ((rc = make_linear(h, &h->attn_in, w->attn_in_w, w->attn_in_b, 3 * D, D, 0, 0, 0, 0, 0.f, st))) return rc;
    if ((rc = make_linear(h, &h->attn_out, w->attn_out_w, w->attn_out_b, D, D, 0, 0, 0, 0, 0.f, st))) return rc;
    h->ln_w = arena_take(h, D); h->ln_b = arena_take(h, D);
    PXR_CUDA(h, cudaMemcpyAsync(h->ln_w, w->attn_ln_w, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
    PXR_CUDA(h, cudaMemcpyAsync(h->ln_b, w->attn_ln_b, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  }
  in = (c.fusion == PXR_FUSION_CONCAT) ? M * D : D;
  const float eps = w->bn_eps > 0.f ? w->bn_eps : 1e-5f;
  for (int l = 0; l <= c.n_hidden; ++l) {
    const bool fold = c.use_batch_norm && l > 0;
    const float* W = l < c.n_hidden ? w->mlp_w[l] : w->out_w;
    const float* B = l < c.n_hidden ? w->mlp_b[l] : w->out_b;
    const int N = l < c.n_hidden ? c.hidden[l] : 1;
    PxrLinear* L = l < c.n_hidden ? &h->mlp[l] : &h->out;
    if ((rc = make_linear(h, L, W, B, N, in, fold ? w->bn_w[l - 1] : 0, fold ? w->bn_b[l - 1] : 0,
                          fold ? w->bn_mean[l - 1] : 0, fold ? w->bn_var[l - 1] : 0, eps, st))) return rc;
    in = N;
  }
  if (h->fast_ok && (rc = pxr_tc_prepare_weights(h, st))) return rc;
  if (h->fast_ok && (rc = pxr_items_tc_prepare_weights(h, st))) return rc;
  h->weights_loaded = true;
  h->item_feats = nullptr; h->n_rows = 0; h->items_ready = false;
  return PXR_OK;
}

// ---------------------------------------------------------------------------
// items
// ---------------------------------------------------------------------------
static int64_t rows_padded(int64_t n) { return (n + 31) / 32 * 32; }

static size_t feats_bytes(const pxr_handle* h, int64_t n_rows) {
  return pxr_align_up((size_t)rows_padded(n_rows) * (h->M - 1) * h->cfg.embedding_dim * sizeof(float), 256);
}

extern "C" size_t pxr_items_bytes(const pxr_handle* h, int64_t n_rows) {
  if (!h || n_rows < 0) return 0;
  return feats_bytes(h, n_rows) + ((h->fast_ok && !h->records_only) ? pxr_tc_item_bytes(h, n_rows) : 0) + 256;
}

extern "C" int pxr_precompute_items(pxr_handle* h, const float* item_embedding, const int64_t* item_idx,
                                    const int64_t* tag_idx, const float* vis, const float* txt, const float* num,
                                    int64_t n_rows, int64_t item_base, void* workspace, size_t workspace_bytes,
                                    pxr_stream stream) {
  if (!h) return PXR_ERR_INVALID;
  if (!h->weights_loaded) PXR_FAIL(h, PXR_ERR_STATE, "pxr_load_weights must be called before pxr_precompute_items");
  if (n_rows < 0 || !item_embedding || (!tag_idx && n_rows)) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_precompute_items: bad arguments");
  if (workspace_bytes < pxr_items_bytes(h, n_rows) || (n_rows && !workspace)) PXR_FAIL(h, PXR_ERR_WORKSPACE, "item workspace too small: %zu < %zu", workspace_bytes, pxr_items_bytes(h, n_rows));
  if (((uintptr_t)workspace) & 255) PXR_FAIL(h, PXR_ERR_INVALID, "item workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  h->item_feats = (float*)workspace; h->n_rows = n_rows; h->item_base = item_base;
  h->item_missing = nullptr; h->items_ready = true;
  h->item_fast = (char*)workspace + feats_bytes(h, n_rows);
  if (n_rows == 0) return PXR_OK;
  // tensor-pipe (3xTF32, fp32-accurate) item path next to the fused scoring kernel; the fp32 SIMT kernel otherwise
  const bool items_tc = h->path == PXR_PATH_TCGEN05 && pxr_items_tc_supported(h);
  int rc = items_tc ? pxr_launch_items_tc(h, item_embedding, item_idx, tag_idx, vis, txt, num, n_rows, item_base, h->item_feats, st)
                    : pxr_launch_items_simt(h, item_embedding, item_idx, tag_idx, vis, txt, num, n_rows, item_base, h->item_feats, st);
  if (rc) return rc;
  if (h->fast_ok && !h->records_only) return pxr_tc_prepare_items(h, n_rows, h->item_fast, st);
  return PXR_OK;
}

extern "C" int pxr_set_missing_items(pxr_handle* h, const uint8_t* flags, int64_t n_rows) {
  if (!h) return PXR_ERR_INVALID;
  if (flags && n_rows != h->n_rows) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_set_missing_items: %lld flags for %lld precomputed item rows", (long long)n_rows, (long long)h->n_rows);
  h->item_missing = flags;
  return PXR_OK;
}

// ---------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------
extern "C" size_t pxr_score_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k) {
  if (!h || n_users <= 0) return 256;
  if (h->path == PXR_PATH_TCGEN05 && !h->records_only && pxr_tc_can_run(h, k)) return pxr_tc_topk_bytes(h, n_users, k) + 256;
  return pxr_simt_topk_bytes(h, n_users, k) + 256;
}

extern "C" int pxr_score_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                              const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k, float* out_scores,
                              int32_t* out_idx, void* workspace, size_t workspace_bytes, pxr_stream stream) {
  if (!h) return PXR_ERR_INVALID;
  if (!h->weights_loaded) PXR_FAIL(h, PXR_ERR_STATE, "weights not loaded");
  if (!h->items_ready) PXR_FAIL(h, PXR_ERR_STATE, "pxr_precompute_items must be called (after pxr_load_weights) before scoring");
  if (k <= 0 || n_users < 0 || (n_users && (!user_embedding || !user_idx || !out_scores || !out_idx))) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_score_topk: bad arguments");
  if (seen_indptr && !seen_idx) PXR_FAIL(h, PXR_ERR_INVALID, "seen_indptr given without seen_idx");
  if (n_users == 0) return PXR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes < pxr_score_topk_bytes(h, n_users, k)) PXR_FAIL(h, PXR_ERR_WORKSPACE, "score workspace too small: %zu < %zu", workspace_bytes, pxr_score_topk_bytes(h, n_users, k));
  if (h->path == PXR_PATH_TCGEN05 && !h->records_only && pxr_tc_can_run(h, k))
    return pxr_tc_score_topk(h, user_embedding, user_idx, n_users, seen_indptr, seen_idx, k, out_scores, out_idx, workspace, workspace_bytes, st);
  if (h->n_rows == 0) {   // empty catalogue shard: every list is padding
    PXR_CUDA(h, cudaMemsetAsync(out_idx, 0xFF, sizeof(int32_t) * n_users * k, st));
    // -inf bit pattern 0xFF800000 cannot be memset byte-wise; reuse the row top-K kernel on zero items
    return pxr_launch_topk_rows(h, (const float*)workspace, n_users, 0, h->item_base, k, out_scores, out_idx, st);
  }
  // generic fp32 path: per-user blocks sweep the catalogue and keep the running top-K in shared memory
  return pxr_launch_score_topk_simt(h, user_embedding, user_idx, n_users, seen_indptr, seen_idx, k, out_scores, out_idx,
                                    workspace, workspace_bytes, st);
}

extern "C" int pxr_score_pairs(pxr_handle* h, const float* user_embedding, const int64_t* user_idx,
                               const int64_t* item_row, int64_t n, float* out, float* out_logit, pxr_stream stream) {
  if (!h) return PXR_ERR_INVALID;
  if (!h->weights_loaded || !h->items_ready || !h->item_feats) PXR_FAIL(h, PXR_ERR_STATE, "weights / items not loaded");
  if (n < 0 || (n && (!user_embedding || !user_idx || !item_row || !out))) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_score_pairs: bad arguments");
  return pxr_launch_score_simt(h, user_embedding, user_idx, item_row, n, out, out_logit, (cudaStream_t)stream);
}

extern "C" int pxr_set_small_batch(pxr_handle* h, int mode) {
  if (!h) return PXR_ERR_INVALID;
  if (mode < -1 || mode > 1) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_set_small_batch: mode must be -1 (auto), 0 (off) or 1 (whenever possible)");
  h->small_batch = mode;
  return PXR_OK;
}

extern "C" size_t pxr_rescore_lists_bytes(int64_t n_users, int32_t list_len) {
  return n_users > 0 && list_len > 0 ? pxr_rescore_list_bytes(n_users, list_len) + 256 : 256;
}
extern "C" size_t pxr_rescore_bytes(int64_t n_users) { return pxr_rescore_lists_bytes(n_users, 64); }

extern "C" int pxr_rescore_lists(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                                 const int32_t* cand_idx, int32_t list_len, int32_t k, float* out_scores, int32_t* out_idx,
                                 void* workspace, size_t workspace_bytes, pxr_stream stream) {
  if (!h) return PXR_ERR_INVALID;
  if (!h->weights_loaded || !h->items_ready || !h->item_feats) PXR_FAIL(h, PXR_ERR_STATE, "weights / items not loaded");
  if (list_len < 64 || list_len > 1024 || list_len % 64) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_rescore_lists: list_len must be a multiple of 64 up to 1024");
  if (k <= 0 || k > list_len || n_users < 0 || (n_users && (!user_embedding || !user_idx || !cand_idx || !out_scores || !out_idx)))
    PXR_FAIL(h, PXR_ERR_INVALID, "pxr_rescore_lists: bad arguments (1 <= k <= list_len)");
  if (n_users == 0) return PXR_OK;
  if (workspace_bytes < pxr_rescore_lists_bytes(n_users, list_len) || !workspace) PXR_FAIL(h, PXR_ERR_WORKSPACE, "re-score workspace too small");
  return pxr_launch_rescore(h, user_embedding, user_idx, n_users, cand_idx, list_len, k, out_scores, out_idx, workspace, (cudaStream_t)stream);
}

extern "C" int pxr_rescore_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                                const int32_t* cand_idx, int32_t k, float* out_scores, int32_t* out_idx, void* workspace,
                                size_t workspace_bytes, pxr_stream stream) {
  return pxr_rescore_lists(h, user_embedding, user_idx, n_users, cand_idx, 64, k, out_scores, out_idx, workspace, workspace_bytes, stream);
}

extern "C" int pxr_merge_topk(const float* scores_in, const int32_t* idx_in, int32_t n_shards, int64_t n_users,
                              int32_t k, float* out_scores, int32_t* out_idx, pxr_stream stream) {
  if (n_shards <= 0 || k <= 0 || n_users < 0) return PXR_ERR_INVALID;
  return pxr_launch_merge(scores_in, idx_in, n_shards, n_users, k, out_scores, out_idx, (cudaStream_t)stream);
}

extern "C" size_t pxr_metrics_bytes(int64_t n_users, int32_t n_ks) {
  (void)n_ks;
  const int64_t blocks = (n_users + 127) / 128;
  return pxr_align_up((size_t)(blocks > 0 ? blocks : 1) * PXR_MAX_KS * PXR_METRIC_COLS * sizeof(double), 256);
}

extern "C" int pxr_metrics(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, const int64_t* gt_indptr,
                           const int32_t* gt_idx, const int32_t* recall_den, const int32_t* ks, int32_t n_ks, const double* discount,
                           const double* ideal, double* out_sums, void* workspace, size_t workspace_bytes,
                           pxr_stream stream) {
  if (n_ks <= 0 || n_ks > PXR_MAX_KS || !ks || !out_sums || n_users < 0) return PXR_ERR_INVALID;
  if (workspace_bytes < pxr_metrics_bytes(n_users, n_ks)) return PXR_ERR_WORKSPACE;
  return pxr_launch_metrics(topk_idx, k_stride, n_users, gt_indptr, gt_idx, recall_den, ks, n_ks, discount, ideal, out_sums,
                            workspace, (cudaStream_t)stream);
}
