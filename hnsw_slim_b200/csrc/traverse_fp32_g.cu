// The fp32 traversal kernel with the visited hash in global memory (see traverse_fp32.cu).
#define HS_GHASH 1
#include "traverse_fp32.cu"
