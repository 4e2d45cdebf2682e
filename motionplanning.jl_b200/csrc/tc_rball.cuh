// tc_rball.cuh -- interface of the tcgen05 prefilter (tc_rball.cu) used by brute_rball.cu
#pragma once
#include "common.cuh"

namespace mpb {

struct TcPlan {
    float *opB;        // TF32 operands in the canonical UMMA layout, [tile of 128][chunk 4][row 128][4]
    float *nrm_half;   // |x~|^2 / 2 per (padded) sample
    int64_t Npad;
    float delta;       // error allowance of the tensor-core test quantity
};
// false: the error allowance of the TF32 test is not small against r^2 -- use the FP32 CUDA-core prefilter instead
bool tc_band_is_tight(const mpb200_samples *s, double r);
template <int D> int tc_prepare_operands(mpb200_samples *s, double r, TcPlan *plan);
// sweep the first nq_run query columns of the shard; cap > 0: also append hits to the slabs
// rcounts != NULL: the symmetric sweep of a full-range build with slabs -- each unordered pair (q, j), j > q, is
// multiplied once; row q's thread appends it to the FRONT of its own slab row (counts[q] entries) and to the BACK of
// row j's (rcounts[j] entries, atomic slots).  Finish with the merging / sorting slab conversion.
template <int D> int tc_sweep(mpb200_samples *s, const TcPlan &P, double r, int64_t nq_run, int *counts, int cap,
                              int *slab_j, double *slab_s, int *rcounts = nullptr);

}  // namespace mpb
