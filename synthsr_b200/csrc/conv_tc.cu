// placeholder -- replaced by the tcgen05 implementation
#include "common.cuh"
extern "C" {
int ssr_conv3d_pack_weights(const float*, float*, int, int, int, int, void*) { ssr_set_error("tc path not built"); return SSR_ERR_UNSUPPORTED; }
long long ssr_conv3d_packed_size(int, int, int, int) { return 0; }
int ssr_conv3d_fwd_tc(const float*, int, const float*, int, const float*, const float*, float*, int, int, int, int, int, int, void*) { ssr_set_error("tc path not built"); return SSR_ERR_UNSUPPORTED; }
int ssr_conv3d_wgrad_tc(const float*, int, const float*, int, const float*, float*, float*, float*, long long, int, int, int, int, int, void*) { ssr_set_error("tc path not built"); return SSR_ERR_UNSUPPORTED; }
long long ssr_conv3d_wgrad_scratch_bytes(int, int, int, int, int, int, int) { return 0; }
int ssr_tc_selftest(void*) { ssr_set_error("tc path not built"); return SSR_ERR_UNSUPPORTED; }
}
