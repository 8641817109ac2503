// placeholder -- replaced by the tcgen05 flash attention kernels
#include "common.cuh"
using namespace aoz;
extern "C" {
int aoz_attn_fwd(const void*, long long, const void*, long long, const void*, long long, void*, long long, void*, int, int, int, int, float, void*) {
    set_error("aoz_attn_fwd: not built yet"); return AOZ_ERR_UNSUPPORTED; }
long long aoz_attn_bwd_workspace_floats(int B, int H, int Tq) { return (long long)B * H * Tq; }
int aoz_attn_bwd(const void*, long long, const void*, long long, const void*, long long, const void*, long long, const void*, long long,
                 const void*, void*, long long, void*, long long, void*, long long, int, int, int, int, float, void*, void*) {
    set_error("aoz_attn_bwd: not built yet"); return AOZ_ERR_UNSUPPORTED; }
}
