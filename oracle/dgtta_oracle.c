/*
 * TEST INFRASTRUCTURE ONLY — plain-C CPU oracle for the DG-TTA input-transform hot path
 * (MIND-SSC, GIN, affine grid-sample).  See dgtta_oracle_impl.h for the per-function citations
 * of the reference (multimodallearning/DG-TTA: dg_tta/mind.py, dg_tta/gin.py, dg_tta/tta/ modules).
 * Built by oracle/Makefile into oracle/_build/libdgtta_oracle.so and loaded by oracle/cform.py.
 * Exports every function twice: *_f32 (reference arithmetic type) and *_f64 (truth).
 */
#include <math.h>
#include <stdlib.h>

/* dg_tta/mind.py:104-136 — (d,h,w) offsets of the one-hot 3x3x3 kernels mshift1 / mshift2,
 * i.e. all ordered pairs (i>j) of the 6-neighbourhood with squared distance 2. */
static const int MIND_SHIFT1[12][3] = {{0, 0, -1}, {0, -1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, 1}, {1, 0, 0},
                                       {1, 0, 0},  {1, 0, 0},  {0, 1, 0},  {0, 1, 0}, {0, 1, 0}, {0, 1, 0}};
static const int MIND_SHIFT2[12][3] = {{-1, 0, 0}, {-1, 0, 0}, {0, 0, -1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1},
                                       {0, -1, 0}, {0, 0, 1},  {-1, 0, 0}, {0, 0, -1}, {0, 0, 1},  {1, 0, 0}};

void oracle_mind_shift_table(int *shift1, int *shift2)
{
    for (int c = 0; c < 12; ++c)
        for (int k = 0; k < 3; ++k) {
            shift1[c * 3 + k] = MIND_SHIFT1[c][k];
            shift2[c * 3 + k] = MIND_SHIFT2[c][k];
        }
}

#define REAL float
#define SUFFIX _f32
#define EXPFN expf
#define FLOORFN floorf
#define NEARBY nearbyintf
#define FMAFN fmaf
#include "dgtta_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef EXPFN
#undef FLOORFN
#undef NEARBY
#undef FMAFN

#define REAL double
#define SUFFIX _f64
#define EXPFN exp
#define FLOORFN floor
#define NEARBY nearbyint
#define FMAFN fma
#include "dgtta_oracle_impl.h"
#undef REAL
#undef SUFFIX

/* chain truth: the same double-precision functions with a double-precision MIND input, so that GIN -> MIND can be
 * evaluated without rounding the intermediate volume to float32 (oracle_mind_ssc_f64x) */
#define REAL double
#define SUFFIX _f64x
#define IMGT double
#include "dgtta_oracle_impl.h"
