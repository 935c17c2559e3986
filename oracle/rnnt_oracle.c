/*
 * rnnt_oracle.c — CPU restatement of the reference RNN-T loss path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker (or the timed CPU baseline).  The
 * product path (warp-transducer_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against
 *   - the known-answer vectors held by the reference's own tests
 *     (tests/test_cpu.cpp:18-26,77-109; tests/test_gpu.cu:102-132;
 *      pytorch_binding/test/test.py:52-74), committed as tests/golden/*.json;
 *   - outputs of the reference itself: oracle/_ref/libwarprnnt_ref_cpu.so
 *     (built from /root/reference by oracle/Makefile) and the reference's
 *     numpy implementation (pytorch_binding/test/transducer_np.py), both run
 *     in the authoring container by tests/golden/make_golden.py.
 *
 * What is restated (reference file:line, all under /root/reference):
 *   log-softmax over the vocabulary axis      pytorch_binding/warprnnt_pytorch/__init__.py:95-98
 *                                             (torch log_softmax on the CPU path), tests/test.h:35-60
 *   blank/label log-prob gather               include/detail/cpu_rnnt.h:115-128   (setup_probs)
 *   two-way log-sum-exp                       include/detail/rnnt_helper.h:16-24
 *   alpha recursion + forward log-likelihood  include/detail/cpu_rnnt.h:175-212
 *   beta recursion                            include/detail/cpu_rnnt.h:214-251
 *   sparse gradient w.r.t. log-probs          include/detail/cpu_rnnt.h:253-267
 *   batch driver, padded-label stride         include/detail/cpu_rnnt.h:272-304
 *   chain rule to logits (what autograd adds) dL/dx_k = g_k - softmax_k * sum_v g_v
 *                                             (equals include/detail/gpu_rnnt_kernel.h:159-177)
 *
 * Layout: activations / gradients are [N, maxT, maxU, V] row-major, labels are
 * [N, maxU-1] padded, exactly as include/rnnt.h:75-89 describes.
 *
 * The file is compiled twice through the REAL macro below: once with float
 * (mirrors the reference's fp32 arithmetic) and once with double ("truth").
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* ------------------------------------------------------------------ f32 */
#define REAL float
#define SUF f32
#define R_EXP expf
#define R_LOG logf
#define R_LOG1P log1pf
#define R_FABS fabsf
#include "rnnt_oracle_impl.inc"
#undef REAL
#undef SUF
#undef R_EXP
#undef R_LOG
#undef R_LOG1P
#undef R_FABS

/* ------------------------------------------------------------------ f64 */
#define REAL double
#define SUF f64
#define R_EXP exp
#define R_LOG log
#define R_LOG1P log1p
#define R_FABS fabs
#include "rnnt_oracle_impl.inc"
#undef REAL
#undef SUF
#undef R_EXP
#undef R_LOG
#undef R_LOG1P
#undef R_FABS

int oracle_version(void) { return 1; }

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
