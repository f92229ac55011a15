// kao_inst.cu — explicit instantiations of the search kernels (kao_kernels.cuh) for ONE row width
// (KAO_INST_W words), counter depth (KAO_INST_NPH high planes) and evaluation mode (KAO_INST_DELTA;
// KAO_INST_TRANS: the column-major evaluator).
// The Makefile compiles this file once per combination; the objects build in parallel.
#include "kao_kernels.cuh"

#if !defined(KAO_INST_W) || !defined(KAO_INST_NPH) || !defined(KAO_INST_DELTA)
#error "compile with -DKAO_INST_W=<1|2|4|8> -DKAO_INST_NPH=<3|5> -DKAO_INST_DELTA=<0|1>"
#endif

#define KAO_INST_FULL(W, NPH, R, O)                                      \
    template __global__ void KAO_ROUND_KERNEL(W, NPH, R, O);            \
    template __global__ void KAO_PERSISTENT_KERNEL(W, NPH, R, O, threads_for<W>(), false);
#define KAO_INST_DELTA_K(W, NPH, R, O) template __global__ void KAO_PERSISTENT_KERNEL(W, NPH, R, O, KAO_THREADS_DELTA, true);

#if defined(KAO_INST_TUNE) && KAO_INST_TUNE >= 0
// schedules of the column-major evaluator (kao_set_schedule), one object per barrier form
#define KAO_INST_TUNE_K(S, C, T, U, RL, F) template __global__ void KAO_PERSISTENT_KERNEL_TUNE(S, C, T, U, RL, F);
#if KAO_INST_TUNE == 0
KAO_FOR_TUNE_SYNC_0(KAO_INST_TUNE_K)
#elif KAO_INST_TUNE == 1
KAO_FOR_TUNE_SYNC_1(KAO_INST_TUNE_K)
#elif KAO_INST_TUNE == 2
KAO_FOR_TUNE_SYNC_2(KAO_INST_TUNE_K)
#elif KAO_INST_TUNE == 3
KAO_FOR_TUNE_SYNC_3(KAO_INST_TUNE_K)
#else
KAO_FOR_TUNE_SYNC_4(KAO_INST_TUNE_K)
#endif
#elif defined(KAO_INST_TRANS) && KAO_INST_TRANS
// column-major evaluator (kao_device_t.cuh): rows of up to 64 slots
template __global__ void KAO_PERSISTENT_KERNEL_T(KAO_INST_W, 0);
template __global__ void KAO_PERSISTENT_KERNEL_T(KAO_INST_W, 32);
#elif KAO_INST_DELTA
#if KAO_INST_W > 2
#error "delta evaluation: rows of up to 64 slots"
#endif
KAO_FOR_CFGS_NARROW(KAO_INST_DELTA_K, KAO_INST_W, KAO_INST_NPH)
#elif KAO_INST_W <= 2
KAO_FOR_CFGS_NARROW(KAO_INST_FULL, KAO_INST_W, KAO_INST_NPH)
#else
KAO_FOR_CFGS_WIDE(KAO_INST_FULL, KAO_INST_W, KAO_INST_NPH)
#endif
