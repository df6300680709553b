/*
ffr_params.cuh -- what the host hands to a render kernel, shared by the ahead-of-time
interpreter kernels (ffr_kernels.cuh) and the run-time compiled flame-specialised kernel
(ffr_jit_kernel.cuh).
*/

#pragma once

#include "ffr_device.cuh"

struct DevStats
{
    u64 s_iter, s_plot;
    u64 xf_dist[FFR_MAX_XFORMS];       /* by SORTED xform index; host maps to JSON ids */
    u64 pt_min[3], pt_max[3];          /* f64_to_ordered() keys */
    u64 n_bad;
    uint32_t abort, pad;
    u64 bad_xf[FFR_MAX_BAD_RECORDED];  /* JSON ids */
    double bad_pt[FFR_MAX_BAD_RECORDED][3];
};

struct RenderParams
{
    const void *blob;              /* DevFlameT<T> | DevXFormT<T>[] | DevVarT<T>[] */
    const void *colors;            /* T[] */
    void *buffer;                  /* cells x (1+r) elements of sizeof(T) bytes */
    DevStats *stats;
    unsigned int *work_counter;
    void *rsl_scratch;             /* K1b: randrsl columns, 16*FFR_TPB words per block */
    u64 *trace;                    /* FFR_SCATTER_TRACE: where sample `it` of chain k was scattered to, at
                                      trace[it*chain_count + k]: the buffer cell index, or bit 63 | the
                                      element index in `acc` (K1e's tiles); ~0 when not plotted */
    u64 chain_first, chain_count, chain_len, last_len, base_seed, bv_limit;
    uint32_t blob_bytes, scatter_mode;
    void *acc;                     /* K1e: accumulation tile (or null): scrambled cell order, folded into
                                      `buffer` by fold_acc_kernel after the launch; or compact rows */
    unsigned int *dir;             /* K1e, buffers beyond the TLB reach: row directory of the compact
                                      tile (fold_dir_kernel), and its allocation counter */
    unsigned int *dir_next;
};

#define FFR_DIR_EMPTY  0xffffffffu   /* row not seen yet */
#define FFR_DIR_BUSY   0xfffffffeu   /* being allocated by another thread: scatter into the buffer */
#define FFR_DIR_DIRECT 0xfffffffdu   /* tile full: this row is scattered into the buffer */
#define FFR_DIR_ROW_SHIFT 9          /* 512 cells (4 KiB of u64) per row */

/* (size_t)((pf - lo) * mult_d): truncating conversion, buffer_renderer.hpp:202 */
__device__ __forceinline__ u64 to_index(double v) { return __double2ull_rz(v); }
__device__ __forceinline__ u64 to_index(float v) { return __float2ull_rz(v); }
/* ++hist (buffer_renderer.hpp:211-215) for u64 / u32 counters */
__device__ __forceinline__ void hist_add(u64 *cell, unsigned n) { atomicAdd(cell,(u64)n); }
__device__ __forceinline__ void hist_add(unsigned int *cell, unsigned n) { atomicAdd(cell,n); }

