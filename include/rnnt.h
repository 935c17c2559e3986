/*
 * rnnt.h — C-ABI of the B200-native RNN-Transducer loss (libwarprnnt.so).
 *
 * Binary drop-in for the reference library's interface: the same five exported symbols, the
 * same enum values and the same 32-byte by-value options struct, so anything that links the
 * reference (its PyTorch / TensorFlow bindings, tests/*.cu) links this instead.  Each
 * declaration names the reference interface it replaces (paths relative to the reference
 * checkout).  Only the CUDA device path exists here: there is no CPU implementation behind
 * this ABI (see `loc` below).
 *
 * All tensors are dense, row-major, no padding between dimensions:
 *   activations / gradients  [minibatch, maxT, maxU, alphabet_size]
 *   flat_labels              [minibatch, maxU - 1]   (padded to maxU-1 per utterance)
 *   label_lengths, input_lengths, costs  [minibatch]
 */
#ifndef B200_RNNT_H_
#define B200_RNNT_H_

#ifdef __cplusplus
#include <cstddef>
extern "C" {
#else
#include <stdbool.h>
#include <stddef.h>
#endif

/* Opaque CUDA stream handle (same forward declaration trick as reference include/rnnt.h:14). */
typedef struct CUstream_st* CUstream;

/* Return codes — values fixed by reference include/rnnt.h:16-22. */
typedef enum {
    RNNT_STATUS_SUCCESS = 0,
    RNNT_STATUS_MEMOPS_FAILED = 1,    /* a cudaMemcpy/cudaMemset-class call failed            */
    RNNT_STATUS_INVALID_VALUE = 2,    /* null pointer, non-positive dimension, unknown loc    */
    RNNT_STATUS_EXECUTION_FAILED = 3, /* kernel launch/execution error, or loc == RNNT_CPU    */
    RNNT_STATUS_UNKNOWN_ERROR = 4
} rnntStatus_t;

/* Replaces reference include/rnnt.h:25 (src/rnnt_entrypoint.cpp:14-16).  Returns 1: the
 * reference's tests refuse to run against any other value (tests/test_cpu.cpp:382-385). */
int get_warprnnt_version(void);

/* Replaces reference include/rnnt.h:31 (src/rnnt_entrypoint.cpp:18-35).  Static strings. */
const char* rnntGetStatusString(rnntStatus_t status);

/* Reference include/rnnt.h:33-36. */
typedef enum {
    RNNT_CPU = 0, /* not available in this library: compute calls return EXECUTION_FAILED */
    RNNT_GPU = 1
} rnntComputeLocation;

/*
 * Options, passed BY VALUE.  Layout fixed by reference include/rnnt.h:43-64
 * (x86-64: 32 bytes, offsets 0/4/8/16/20/24/28).  Zero-initialise before filling.
 */
struct rnntOptions {
    rnntComputeLocation loc;  /* must be RNNT_GPU                                              */
    unsigned int num_threads; /* accepted and ignored (no host threading on the device path)  */
    CUstream stream;          /* all device work is enqueued here                              */
    int blank_label;          /* index of the blank symbol                                     */
    int maxT;                 /* time extent of the activation tensor                          */
    int maxU;                 /* label extent of the activation tensor (max label length + 1) */
    bool batch_first;         /* ignored, as the reference's GPU path ignores it
                                 (include/detail/gpu_rnnt_kernel.h:7 always indexes [N,T,U,V]; the
                                 reference's tests/test_gpu.cu:42-50 leave it false with [N,T,U,V]
                                 data).  For [T,U,N,V] tensors use rnnt_b200_loss_async_layout. */
};
#ifndef __cplusplus
typedef struct rnntOptions rnntOptions;
#endif

/*
 * RNN-T negative log-likelihood per utterance and, optionally, its gradient with respect to
 * the raw logits.  Replaces reference include/rnnt.h:104-113 (src/rnnt_entrypoint.cpp:38-93)
 * for loc == RNNT_GPU, with the GPU path's conventions (include/detail/gpu_rnnt.h:82-215):
 *
 *   activations   DEVICE, raw (un-normalised) logits; log-softmax over the last axis is
 *                 computed inside.
 *   gradients     DEVICE, same shape, or NULL for loss only.  When non-NULL every element is
 *                 defined on return: d cost[b] / d logit for valid cells, 0 for padded cells
 *                 (t >= input_lengths[b] or u > label_lengths[b]).  Not scaled or reduced.
 *   flat_labels, label_lengths, input_lengths
 *                 DEVICE pointers, as every caller of the reference's GPU path passes them
 *                 (tests/test_gpu.cu:54-59); HOST pointers are also accepted (detected with
 *                 cudaPointerGetAttributes and staged through the workspace).
 *   costs         HOST pointer (as the reference: D2H copy inside the call,
 *                 include/detail/gpu_rnnt.h:209-213); a DEVICE pointer is accepted too.
 *   workspace     DEVICE scratch of get_workspace_size(..., gpu=true, ...) bytes.
 *
 * The call returns after the stream has finished (costs are readable on return), like the
 * reference.  Errors are reported by return code only; nothing is thrown across the ABI.
 * Limits: minibatch*maxT*maxU < 2^31 cells (elements are indexed with 64 bits).
 */
rnntStatus_t compute_rnnt_loss(const float* const activations, float* gradients,
                               const int* const flat_labels, const int* const label_lengths,
                               const int* const input_lengths, int alphabet_size, int minibatch,
                               float* costs, void* workspace, struct rnntOptions options);

/* Double-precision twin.  Replaces reference include/rnnt.h:115-124
 * (src/rnnt_entrypoint.cpp:130-185). */
rnntStatus_t compute_rnnt_loss_fp64(const double* const activations, double* gradients,
                                    const int* const flat_labels,
                                    const int* const label_lengths,
                                    const int* const input_lengths, int alphabet_size,
                                    int minibatch, double* costs, void* workspace,
                                    struct rnntOptions options);

/*
 * Scratch size in bytes.  Replaces reference include/rnnt.h:139-143
 * (src/rnnt_entrypoint.cpp:96-128): same signature, same INVALID_VALUE rule for non-positive
 * extents.  The byte count differs from the reference's (this library keeps a different
 * lattice: see DESIGN.md "HBM layout"); callers always size through this function.
 * gpu == false returns the reference's CPU formula for source compatibility only.
 */
#ifdef __cplusplus
rnntStatus_t get_workspace_size(int maxT, int maxU, int minibatch, bool gpu, size_t* size_bytes,
                                size_t dtype_size = sizeof(float));
#else
rnntStatus_t get_workspace_size(int maxT, int maxU, int minibatch, bool gpu, size_t* size_bytes,
                                size_t dtype_size);
#endif

/* Same function under the name BASELINE.json's north_star uses (the reference has no such
 * symbol; exported so either spelling links). */
rnntStatus_t get_rnnt_workspace_size(int maxT, int maxU, int minibatch, bool gpu,
                                     size_t* size_bytes, size_t dtype_size);

/* ---------------------------------------------------------------------------------------
 * Extensions (no reference counterpart).  Used by the warprnnt_pytorch operator to drop
 * the framework-side passes SURVEY.md §8(f).1 lists, and by bench.py for device timing.
 * ------------------------------------------------------------------------------------- */

/*
 * Asynchronous variant: identical computation, but
 *   - costs_device [minibatch] is written on the DEVICE and the call does NOT synchronise;
 *   - gradients are multiplied by grad_scale (pass 1.0f for the plain gradient): folds the
 *     'mean' 1/N of warprnnt_pytorch/__init__.py:38-40 into the write-back.
 * flat_labels/label_lengths/input_lengths must be DEVICE pointers here.
 */
rnntStatus_t compute_rnnt_loss_async(const float* const activations, float* gradients,
                                     const int* const flat_labels,
                                     const int* const label_lengths,
                                     const int* const input_lengths, int alphabet_size,
                                     int minibatch, float* costs_device, float grad_scale,
                                     void* workspace, struct rnntOptions options);

rnntStatus_t compute_rnnt_loss_async_fp64(const double* const activations, double* gradients,
                                          const int* const flat_labels,
                                          const int* const label_lengths,
                                          const int* const input_lengths, int alphabet_size,
                                          int minibatch, double* costs_device, double grad_scale,
                                          void* workspace, struct rnntOptions options);

/*
 * Training-step split used by warprnnt_pytorch: the gradient pass runs in
 * autograd's backward with the upstream gradient folded in, so the framework never makes a
 * separate pass over the [N,T,U,V] tensor (the reference multiplies it in place afterwards,
 * pytorch_binding/warprnnt_pytorch/__init__.py:47-50).  All pointers DEVICE, no synchronisation.
 *
 *   rnnt_b200_forward   log-softmax statistics + alpha (+ beta when prepare_backward != 0) lattices
 *                       into `workspace`, costs_device[minibatch] = -log-likelihood.
 *   rnnt_b200_backward  gradient pass only, from the SAME workspace and the same activations:
 *                       gradients[b,...] = grad_scale * grad_costs_device[b] * d cost[b]/d logits
 *                       (grad_costs_device may be NULL = all ones).  Zeros on padding.
 */
rnntStatus_t rnnt_b200_forward(const float* const activations, const int* const flat_labels,
                               const int* const label_lengths, const int* const input_lengths,
                               int alphabet_size, int minibatch, float* costs_device,
                               int prepare_backward, void* workspace, struct rnntOptions options);
rnntStatus_t rnnt_b200_forward_fp64(const double* const activations, const int* const flat_labels,
                                    const int* const label_lengths, const int* const input_lengths,
                                    int alphabet_size, int minibatch, double* costs_device,
                                    int prepare_backward, void* workspace, struct rnntOptions options);
rnntStatus_t rnnt_b200_backward(const float* const activations, float* gradients,
                                const int* const flat_labels, const int* const label_lengths,
                                const int* const input_lengths, int alphabet_size, int minibatch,
                                const float* grad_costs_device, float grad_scale, void* workspace,
                                struct rnntOptions options);
rnntStatus_t rnnt_b200_backward_fp64(const double* const activations, double* gradients,
                                     const int* const flat_labels, const int* const label_lengths,
                                     const int* const input_lengths, int alphabet_size,
                                     int minibatch, const double* grad_costs_device,
                                     double grad_scale, void* workspace, struct rnntOptions options);

/*
 * Explicit activation layout.  RNNT_B200_LAYOUT_TUNV takes activations and
 * gradients as [maxT, maxU, minibatch, alphabet_size] - the layout the reference's CPU path indexes
 * when batch_first == false (include/detail/cpu_rnnt.h:139-144) and that its GPU path never
 * implemented.  Labels, lengths and costs stay [minibatch, ...].  Otherwise identical to
 * compute_rnnt_loss_async (all pointers DEVICE, no synchronisation).
 */
enum { RNNT_B200_LAYOUT_NTUV = 0, RNNT_B200_LAYOUT_TUNV = 1 };
rnntStatus_t rnnt_b200_loss_async_layout(int layout, const float* activations, float* gradients,
                                         const int* flat_labels, const int* label_lengths,
                                         const int* input_lengths, int alphabet_size, int minibatch,
                                         float* costs_device, float grad_scale, void* workspace,
                                         struct rnntOptions options);
rnntStatus_t rnnt_b200_loss_async_layout_fp64(int layout, const double* activations, double* gradients,
                                              const int* flat_labels, const int* label_lengths,
                                              const int* input_lengths, int alphabet_size, int minibatch,
                                              double* costs_device, double grad_scale, void* workspace,
                                              struct rnntOptions options);

/*
 * 16-bit storage variants: logits and gradients in bf16 or fp16, arithmetic,
 * lattice and costs in fp32; 6 B per logit instead of 12.  Same semantics as the async / split
 * entries above; workspace sized with dtype_size = sizeof(float).  dtype: RNNT_B200_BF16 / _FP16.
 * The 16-B fast path needs alphabet_size % 8 == 0 and 16-B aligned tensors (else element-wise).
 */
enum { RNNT_B200_BF16 = 1, RNNT_B200_FP16 = 2 };
rnntStatus_t rnnt_b200_loss_async_16(int dtype, const void* activations, void* gradients,
                                     const int* flat_labels, const int* label_lengths,
                                     const int* input_lengths, int alphabet_size, int minibatch,
                                     float* costs_device, float grad_scale, void* workspace,
                                     struct rnntOptions options);
rnntStatus_t rnnt_b200_forward_16(int dtype, const void* activations, const int* flat_labels,
                                  const int* label_lengths, const int* input_lengths,
                                  int alphabet_size, int minibatch, float* costs_device,
                                  int prepare_backward, void* workspace, struct rnntOptions options);
rnntStatus_t rnnt_b200_backward_16(int dtype, const void* activations, void* gradients,
                                   const int* flat_labels, const int* label_lengths,
                                   const int* input_lengths, int alphabet_size, int minibatch,
                                   const float* grad_costs_device, float grad_scale, void* workspace,
                                   struct rnntOptions options);

/*
 * Additive joint network, logits never materialised: for models whose logits
 * are  h[b,t,u,k] = trans[b,t,k] + pred[b,u,k]  (how the reference's own timing script builds them,
 * pytorch_binding/test/test_time.py:73).  trans [minibatch,maxT,V], pred [minibatch,maxU,V];
 * outputs costs_device [minibatch] and, when both gradient pointers are non-NULL,
 * grad_trans = grad_scale * d cost/d trans, grad_pred likewise (docs/rnnt_notes.tex:147-153).
 * All pointers DEVICE, fp32, no synchronisation.  Workspace from rnnt_b200_add_joint_workspace_size.
 * HBM traffic is O(N (T+U) V) instead of O(N T U V).
 */
rnntStatus_t rnnt_b200_add_joint_loss(const float* trans, const float* pred, float* grad_trans,
                                      float* grad_pred, const int* flat_labels,
                                      const int* label_lengths, const int* input_lengths,
                                      int alphabet_size, int minibatch, float* costs_device,
                                      float grad_scale, void* workspace, struct rnntOptions options);
rnntStatus_t rnnt_b200_add_joint_workspace_size(int maxT, int maxU, int minibatch, int alphabet_size,
                                                size_t* size_bytes);
/* Training-step split of the same (see rnnt_b200_forward / rnnt_b200_backward): the backward half
 * folds grad_costs_device[b] * grad_scale into the factor gradients. */
rnntStatus_t rnnt_b200_add_joint_forward(const float* trans, const float* pred, const int* flat_labels,
                                         const int* label_lengths, const int* input_lengths,
                                         int alphabet_size, int minibatch, float* costs_device,
                                         int prepare_backward, void* workspace, struct rnntOptions options);
rnntStatus_t rnnt_b200_add_joint_backward(const float* trans, const float* pred, float* grad_trans,
                                          float* grad_pred, const int* flat_labels,
                                          const int* label_lengths, const int* input_lengths,
                                          int alphabet_size, int minibatch, const float* grad_costs_device,
                                          float grad_scale, void* workspace, struct rnntOptions options);

/* Debug / test hook: forward and backward log-likelihoods (natural log, as doubles on the host) that
 * the last loss+gradient call left in `workspace`.  The reference checks their agreement in debug
 * builds (include/detail/cpu_rnnt.h:167-170); tests/test_gpu_round2.py does the same.  Synchronises. */
rnntStatus_t rnnt_b200_debug_log_likelihoods(const void* workspace, int maxT, int maxU, int minibatch,
                                             size_t dtype_size, double* llf_host, double* llb_host);

/* Number of kernels the last compute call on this thread launched (bench.py's gpu_launches). */
int rnnt_b200_last_launch_count(void);

/* Per-kernel device timing for bench.py's roofline leg.  With profiling enabled on the calling
 * thread, each compute call records CUDA events on options.stream around its three kernels;
 * rnnt_b200_last_kernel_ms() waits for them and writes {rowstats, lattice, grad} milliseconds
 * (-1 where not run) and returns how many were measured. */
void rnnt_b200_set_profiling(int enabled);
int rnnt_b200_last_kernel_ms(float* ms3);
/* Mean {rowstats, lattice, grad} milliseconds over every call recorded on this thread since the
 * previous collect (waits for their events), returns the number of calls and resets the record.
 * Lets a timed loop run without any host synchronisation between steps. */
int rnnt_b200_profile_collect(float* ms3_mean);

/* Host-side dispatch policy, for tests and tuning notes (no device access).  `what`:
 *   0  lanes sharing one short row in the chunk kernels for alphabet size a and element size b bytes (0: row too long
 *      for the chunk kernels);  1  1 if those lanes are `32 / lanes` apart in the warp (bank-aware mapping), 0 if adjacent;
 *   2  label columns per lane of the fp32 wavefront for maxU = a;  3  its threads per utterance and direction;
 *   4  diagonals of its factor ring (b != 0: next to the streaming passes of other batch groups);
 *   5  split-K slabs of the additive joint's S product for alphabet size a.   Returns -1 for an unknown `what`. */
int rnnt_b200_debug_policy(int what, int a, int b);

/* Build identification string, e.g. "b200-rnnt sm_100a <date>". */
const char* rnnt_b200_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_RNNT_H_ */
