// tcgen05 fused pair-scoring path (placeholder until the kernel lands).
#include "pxr_common.cuh"

bool pxr_tc_supported(const pxr_handle* h) { (void)h; return false; }
size_t pxr_tc_weight_bytes(const pxr_handle* h) { (void)h; return 0; }
int pxr_tc_prepare_weights(pxr_handle* h, cudaStream_t st) { (void)h; (void)st; return PXR_OK; }
size_t pxr_tc_item_bytes(const pxr_handle* h, int64_t n_rows) { (void)h; (void)n_rows; return 0; }
int pxr_tc_prepare_items(pxr_handle* h, int64_t n_rows, void* ws, cudaStream_t st) { (void)h; (void)n_rows; (void)ws; (void)st; return PXR_OK; }
size_t pxr_tc_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k) { (void)h; (void)n_users; (void)k; return 0; }
int pxr_tc_score_topk(pxr_handle* h, const float*, const int64_t*, int64_t, const int64_t*, const int32_t*, int32_t,
                      float*, int32_t*, void*, size_t, cudaStream_t) { PXR_FAIL(h, PXR_ERR_INVALID, "tcgen05 path not built"); }
