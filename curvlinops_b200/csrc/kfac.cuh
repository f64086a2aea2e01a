// KFAC factor accumulation (placeholder)
extern "C" int curv_kfac_accumulate_batch(curv_program* prog, const void* const* param_ptrs,
                               const void* const* const_ptrs, const void* X, const int* layer_nodes,
                               int n_layers, float* const* A_ptrs, float* const* G_ptrs,
                               const int* joint_bias, const float* grad_outputs, int V, float wA,
                               float wG, void* workspace, size_t workspace_bytes, void* stream) {
  return fail(CURV_ERR_UNSUPPORTED, "kfac not built yet");
}
