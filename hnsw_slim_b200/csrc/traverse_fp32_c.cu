// The fp32 traversal kernel with the compact 16-bit visited table (traverse_common.cuh cv_test_and_set).
#define HS_CVTAB 1
#include "traverse_fp32.cu"
