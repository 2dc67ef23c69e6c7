// Batched multi-block order statistics (stc_select.cu).
#pragma once
#include "stc_common.cuh"
#include <algorithm>

#define SEL_MAX_COLS 16
// A row-major matrix whose `cols` interleaved columns are all selected at once: element (r, c) = data[r * ld + c].
// A plain (possibly strided) vector is rows = n, cols = 1, ld = stride.
struct SelJob { const float* data; int rows; int cols; int ld; };
// ks_dev[j * SEL_MAX_COLS + c]: 0-based rank wanted in column c of job j (device memory).
// out_dev[(j * SEL_MAX_COLS + c) * 2 + {0, 1}] = the order statistics of rank k and k + 1 (the latter repeats the former
// when k is the last rank).  Ranks at or beyond the number of non-NaN values are undefined.  Asynchronous on ctx->stream.
int select_ranks_dev(stc_ctx* ctx, const SelJob* jobs_host, int njobs, const int* ks_dev, float* out_dev);
