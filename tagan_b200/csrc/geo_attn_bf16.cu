// bf16-storage build of the fused geometric attention kernels (Q, K, V rows in bf16; fp32 arithmetic and outputs):
// exports tagan_geo_attn_fwd_bf16 / tagan_geo_attn_bwd_bf16 (+ the _part variants).  See geo_attn.cu.
#define TAGAN_GEO_BF16 1
#include "geo_attn.cu"
