// mc_mlp_* entry points (placeholder: not yet implemented).
struct mc_mlp { int64_t steps = 0; };
extern "C" {
int mc_mlp_create(int32_t, const int32_t*, const float* const*, const float* const*, const float*, float, float, float,
                  float, float, int32_t, mc_mlp**) { return fail(MC_ERR_UNSUPPORTED, "mc_mlp_create: not implemented yet"); }
int mc_mlp_destroy(mc_mlp* h) { delete h; return MC_OK; }
int mc_mlp_partial_fit(mc_mlp*, const float*, const int32_t*, int64_t, int32_t, int32_t, mc_grad_sync_fn, void*, double*,
                       void*) { return fail(MC_ERR_UNSUPPORTED, "mc_mlp_partial_fit: not implemented yet"); }
int mc_mlp_get_params(mc_mlp*, float* const*, float* const*) { return fail(MC_ERR_UNSUPPORTED, "not implemented yet"); }
int64_t mc_mlp_steps(const mc_mlp* h) { return h ? h->steps : 0; }
}
